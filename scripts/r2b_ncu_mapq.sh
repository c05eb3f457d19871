#!/bin/bash
# the fixed piecewise-quality test, then one ncu --set full capture of k_tilem (mapq_strand, 50 Mb)
mkdir -p gpurun_out
T=${1:-r2b}
timeout 300 python -m pytest tests/test_gpu_random.py -m gpu -q -k quality_requests -p no:cacheprovider > gpurun_out/${T}_qpieces.log 2>&1
tail -3 gpurun_out/${T}_qpieces.log
export SWEEP_WARMUP=1
CMD="python scripts/stat_sweep.py --only ${2:-mapq_strand} --reps 1"
timeout 200 $CMD > gpurun_out/${T}_plain.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_tilem -c 2 -f -o gpurun_out/${T}_prof_k_tilem $CMD > gpurun_out/${T}_ncu.log 2>&1
echo "ncu rc $?"; tail -2 gpurun_out/${T}_ncu.log
