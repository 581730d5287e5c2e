#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/exp_slabs2.txt
for e in "GSDF_PDL=1" "GSDF_PDL=0"; do
  env $e GSDF_MULTI_DEBUG=1 timeout -k 5 300 python scripts/exp_r2_slabs.py 3 2>&1 | tail -12 >> gpurun_out/exp_slabs2.txt
done
cat gpurun_out/exp_slabs2.txt
