#!/bin/bash
# per-kernel launch lists (ncu, cold-cache serialised times: compare shares) of the five BASELINE configurations
set -x
for spec in "zed_1280x720 63" "kitti_1242x375 256" "zed2k_2208x1242 32" "uhd_3840x2160 16" "gazebo_640x480 128"; do
  set -- $spec
  CMD="python bench.py --config $1 --pairs $2 --steps 2 --warmup 3 --no-extra --no-cpu-baseline --no-e2e"
  $CMD > gpurun_out/plain_$1.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum \
      --clock-control none -k regex:"cl_|scene_flow|dynamic_map" -s 40 -c 20 --csv --log-file gpurun_out/r02_launches_$1.csv $CMD > /dev/null 2>&1
done
