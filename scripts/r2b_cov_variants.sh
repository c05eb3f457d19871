#!/bin/bash
# coverage-kind launch shapes (build/variants/*, scripts/build_variant.sh): device-resident sweep of the coverage statistics
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=coverage,coverage_strand,coverage_ext,coverage_ext_strand,coverage_gc
echo "== in-tree (192 x 8)"; timeout 120 python scripts/stat_sweep.py --only $S 2>&1 | grep "%"
for d in build/variants/${VGLOB:-c}*/; do echo "== $(basename $d)"; PSS_LIB=$PWD/$d/libpss_b200.so timeout 120 python scripts/stat_sweep.py --only $S 2>&1 | grep "%"; done
