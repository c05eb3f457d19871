#!/bin/bash
# run the short device-resident bench for every kernel variant under build/variants (and the in-tree library first)
run() {
  PSS_LIB=$2 python bench.py --interval-mb ${MB:-100} --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-binned --no-strong --no-e2e-api 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('%-16s step %.2f ms  var %.3f ms (%.3f)  ext %.3f ms  chk %s' % ('$1', d['ms_per_step'], r['kernel_ms'], r['frac'], r['coverage_ext_kernel_ms'], d['checksum_rank0']['variation_strand']))
except Exception as e: print('$1', 'FAILED', e)
"
}
run in-tree $PWD/pysamstats_b200/csrc/libpss_b200.so
for d in build/variants/*/; do run $(basename $d) $PWD/$d/libpss_b200.so; done
