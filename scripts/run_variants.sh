#!/bin/bash
# run the short device-resident bench for every kernel variant under build/variants
for d in build/variants/*/; do
  n=$(basename $d)
  PSS_LIB=$PWD/$d/libpss_b200.so python bench.py --interval-mb ${MB:-100} --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('%-16s step %.2f ms  var %.3f ms (%.3f)  ext %.3f ms  chk %s' % ('$n', d['ms_per_step'], r['kernel_ms'], r['frac'], r['coverage_ext_kernel_ms'], d['checksum_rank0']['variation_strand']))
except Exception as e: print('$n', 'FAILED', e)
"
done
