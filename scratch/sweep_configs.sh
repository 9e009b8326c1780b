#!/bin/bash
# usage: sweep.sh out.txt -- python bench.py over the BASELINE configurations (device-resident, no e2e / cpu baseline)
out=$1; : > $out
run() { python bench.py --config $1 --pairs $2 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); r=l['roofline']
print('$1 pairs=$2 pairs/s=%.0f Mpix/s=%.0f ms/step=%.3f sf_ms=%.4f sf_frac=%.3f cl_ms=%.3f objs=%d dyn=%d' % (l['value'], l['mpix_per_s'], l['ms_per_step'], r['kernel_ms'], r['frac'], l['clusterer']['ms'], l['objects_per_step'], l['clusterer']['dynamic_pixels_per_step']))" >> $out; }
run gazebo_640x480 128
run kitti_1242x375 1024
run zed2k_2208x1242 32
run uhd_3840x2160 16
run uhd_3840x2160 1
run zed_1280x720 1
run zed_1280x720 8
cat $out
