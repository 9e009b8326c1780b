#!/bin/bash
# usage: run_variants.sh out.log [env assignments...] -- benchmarks every scratch/variants/*.so (kernel times from bench.py)
out=$1; shift
: > $out
for so in /root/repo/scratch/variants/*.so; do
  n=$(basename $so .so)
  env "$@" MOD_B200_LIB=$so python bench.py --steps 200 --warmup 3 --no-e2e --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); r=l['roofline']
print('$n $*', 'sf_ms=%.4f frac=%.3f cl_ms=%.4f step_ms=%.3f value=%.0f' % (r['kernel_ms'], r['frac'], l['clusterer']['ms'], l['ms_per_step'], l['value']))" >> $out 2>&1
done
cat $out
