#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/exp_slabs.txt
for e in "GSDF_PDL=1" "GSDF_PDL=0" "GSDF_PDL=1 GSDF_NO_GRAPH=1"; do
  env $e timeout -k 5 300 python scripts/exp_r2_slabs.py 1 2 3 4 >> gpurun_out/exp_slabs.txt 2>&1
done
cat gpurun_out/exp_slabs.txt
