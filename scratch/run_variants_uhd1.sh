#!/bin/bash
# like run_variants.sh, on one 3840x2160 pair per step (single-frame latency of the device path)
out=$1; shift
: > $out
for so in /root/repo/scratch/variants/*.so; do
  n=$(basename $so .so)
  env "$@" MOD_B200_LIB=$so python bench.py --config uhd_3840x2160 --pairs 1 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); r=l['roofline']
print('$n uhd x1', 'sf_ms=%.4f cl_ms=%.4f step_ms=%.3f' % (r['kernel_ms'], l['clusterer']['ms'], l['ms_per_step']))" >> $out 2>&1
done
cat $out
