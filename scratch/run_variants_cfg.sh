#!/bin/bash
# usage: run_variants_cfg.sh out.log config pairs -- like run_variants.sh for one configuration / batch
out=$1; cfg=$2; pairs=$3
for so in /root/repo/scratch/variants/*.so; do
  n=$(basename $so .so)
  MOD_B200_LIB=$so python bench.py --config $cfg --pairs $pairs --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); r=l['roofline']
print('$n $cfg x$pairs', 'sf_ms=%.4f cl_ms=%.4f step_ms=%.3f objs=%d' % (r['kernel_ms'], l['clusterer']['ms'], l['ms_per_step'], l['objects_per_step']))" >> $out 2>&1
done
