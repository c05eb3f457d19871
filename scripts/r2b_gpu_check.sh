#!/bin/bash
# One GPU-box call: the GPU test-suite, then device-resident timings of every statistic (default, min_baseq 20, no_del).
# Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2b}
if [ "${SKIP_TESTS:-0}" != 1 ]; then
timeout 600 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/${T}_gputests.log 2>&1
echo "pytest rc $?" >> gpurun_out/${T}_gputests.log
tail -3 gpurun_out/${T}_gputests.log
fi
Q=coverage,variation_strand,tlen,tlen_strand,mapq,mapq_strand,baseq,baseq_strand,baseq_ext,baseq_ext_strand
timeout 300 python scripts/stat_sweep.py > gpurun_out/${T}_sweep.txt 2>&1
timeout 300 python scripts/stat_sweep.py --min-baseq 20 > gpurun_out/${T}_sweep_minq20.txt 2>&1
timeout 300 python scripts/stat_sweep.py --no-del --only $Q > gpurun_out/${T}_sweep_nodel.txt 2>&1
for f in sweep sweep_minq20 sweep_nodel; do echo "== $f"; grep -E "%" gpurun_out/${T}_$f.txt; done
