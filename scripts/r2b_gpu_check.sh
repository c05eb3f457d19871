#!/bin/bash
# One GPU-box call: the GPU test-suite, then device-resident timings of the statistics this round touched
# (k_tilem vs the round-1 gather for mapq; min_baseq / no_del off the generic kernel).  Everything lands in gpurun_out/.
mkdir -p gpurun_out
T=${1:-r2b}
timeout 600 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/${T}_gputests.log 2>&1
echo "pytest rc $?" >> gpurun_out/${T}_gputests.log
tail -3 gpurun_out/${T}_gputests.log
ONLY=mapq,mapq_strand,tlen,tlen_strand
timeout 200 python scripts/stat_sweep.py --only $ONLY > gpurun_out/${T}_sweep_new.txt 2>&1
PSS_MAPQ_GATHER=1 timeout 200 python scripts/stat_sweep.py --only mapq,mapq_strand > gpurun_out/${T}_sweep_gather.txt 2>&1
timeout 200 python scripts/stat_sweep.py --only $ONLY --min-baseq 20 > gpurun_out/${T}_sweep_minq20.txt 2>&1
timeout 200 python scripts/stat_sweep.py --only $ONLY --no-del > gpurun_out/${T}_sweep_nodel.txt 2>&1
timeout 200 python scripts/stat_sweep.py --only $ONLY --no-del --generic > gpurun_out/${T}_sweep_nodel_generic.txt 2>&1
for f in new gather minq20 nodel nodel_generic; do echo "== $f"; grep -E "mapq|tlen" gpurun_out/${T}_sweep_$f.txt | grep -v binned; done
