#!/bin/bash
# build_variant.sh NAME "-DTTHREADS=256 -DTCTAS=4 ..."  ->  build/variants/NAME/libpss_b200.so
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
D=$ROOT/build/variants/$1
mkdir -p $D/pysamstats_b200/csrc $D/include
cp $ROOT/pysamstats_b200/csrc/*.cu $ROOT/pysamstats_b200/csrc/*.cuh $ROOT/pysamstats_b200/csrc/*.cpp $ROOT/pysamstats_b200/csrc/*.h $ROOT/pysamstats_b200/csrc/Makefile $D/pysamstats_b200/csrc/
cp $ROOT/include/*.h $D/include/
make -s -C $D/pysamstats_b200/csrc libpss_b200.so EXTRA="$2 -Xptxas -v" 2>&1 | grep -A2 "k_tileILi2ELi5" | grep -E "spill|Used" || true
cp $D/pysamstats_b200/csrc/libpss_b200.so $D/libpss_b200.so
rm -rf $D/pysamstats_b200 $D/include
