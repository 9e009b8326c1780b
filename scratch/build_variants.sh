#!/bin/bash
# builds experimental variants of libmod_b200.so into scratch/variants/<name>.so ; usage: build_variants.sh name "flags" ...
set -e
cd /root/repo/moving_object_detector_b200/csrc
mkdir -p /root/repo/scratch/variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
while [ $# -gt 0 ]; do
  name=$1; flags=$2; shift 2
  d=/tmp/t/var_$name; mkdir -p $d
  for f in sceneflow cluster glue hostcopy abi; do
    nvcc $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -fmad=false $flags -c $f.cu -o $d/$f.o &
  done
  wait
  nvcc $ARCH -shared -o /root/repo/scratch/variants/$name.so $d/sceneflow.o $d/cluster.o $d/glue.o $d/hostcopy.o $d/abi.o -cudart static
  echo built $name
done
