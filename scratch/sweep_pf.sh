#!/bin/bash
# L2 prefetch distance of the scene-flow kernel (MOD_B200_PF_WAVES, in waves of resident CTAs) per configuration
out=$1; : > $out
for cp in "zed_1280x720 63" "zed2k_2208x1242 32" "kitti_1242x375 256"; do
  set -- $cp
  for wv in 0.5 1.0 1.5 2.0 3.0; do
    MOD_B200_PF_WAVES=$wv python bench.py --config $1 --pairs $2 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); r=l['roofline']
print('$1 x$2 waves=$wv sf_ms=%.4f frac=%.3f' % (r['kernel_ms'], r['frac']))" >> $out
  done
done
cat $out
