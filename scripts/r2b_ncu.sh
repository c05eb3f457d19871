#!/bin/bash
# ncu --set full captures of one kernel family on the 50 Mb interval: r2b_ncu.sh TAG STAT KERNEL_REGEX
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=$1; STAT=$2; K=$3
export SWEEP_WARMUP=1
CMD="python scripts/stat_sweep.py --only $STAT --reps 1"
timeout 200 $CMD > gpurun_out/${T}_plain.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -f -o gpurun_out/${T}_prof $CMD > gpurun_out/${T}_ncu.log 2>&1
echo "ncu rc $?"; tail -2 gpurun_out/${T}_ncu.log
