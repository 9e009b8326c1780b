#!/bin/bash
# 4K bandwidth sweep (BASELINE configs[4]) on one GPU: pairs per step 1 .. 64, device-resident
out=$1; : > $out
for p in 1 2 4 8 16 32 64; do
python bench.py --config uhd_3840x2160 --pairs $p --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); r=l['roofline']
print('uhd_3840x2160 pairs=$p pairs/s=%.0f Mpix/s=%.0f ms/step=%.3f sf_ms=%.4f sf_frac=%.3f sf_GB/s=%.0f cl_ms=%.3f objs=%d' % (l['value'], l['mpix_per_s'], l['ms_per_step'], r['kernel_ms'], r['frac'], r['achieved'], l['clusterer']['ms'], l['objects_per_step']))" >> $out
done
cat $out
