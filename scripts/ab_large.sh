#!/bin/bash
# A/B of two builds of the library on the block-Jacobi path (developer aid, GPU box):
#   scripts/ab_large.sh <tag>   ->  gpurun_out/<tag>_ab_large.log
tag=${1:-ab}
out=gpurun_out/${tag}_ab_large.log
mkdir -p gpurun_out
: > $out
for lib in ${LIBS:-mpsim_b200/libmpsim_b200_base.so mpsim_b200/libmpsim_b200.so}; do
  echo "== $lib" >> $out
  for args in "50 512" "1 256" "1 512" "4 2048"; do
    MPSIM_B200_LIB=$PWD/$lib timeout 120 python scripts/prof_large.py $args >> $out 2>&1
  done
  MPSIM_B200_LIB=$PWD/$lib timeout 120 python scripts/time_chi256.py >> $out 2>&1
done
echo "== tests (new build)" >> $out
timeout 300 python -m pytest tests/test_gpu_chi256.py tests/test_gpu_kernels.py -m gpu -x -q >> $out 2>&1
tail -25 $out
