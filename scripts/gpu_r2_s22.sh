#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/exp_slabs4.txt
for e in "GSDF_X=default" "GSDF_B200_LIB=$PWD/gsdf_b200/libgsdfb200_nofence.so"; do
  env $e GSDF_MULTI_DEBUG=1 timeout -k 5 300 python scripts/exp_r2_slabs.py 1 2 3 4 2> gpurun_out/stamps.tmp >> gpurun_out/exp_slabs4.txt
  tail -4 gpurun_out/stamps.tmp >> gpurun_out/exp_slabs4.txt
done
cat gpurun_out/exp_slabs4.txt
