#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/exp_slabs3.txt
for e in "GSDF_MULTI_PDL=first" "GSDF_MULTI_PDL=all" "GSDF_MULTI_PDL=none" "GSDF_MULTI_PDL=first GSDF_HINT_SLACK=0"; do
  env $e GSDF_MULTI_DEBUG=1 timeout -k 5 300 python scripts/exp_r2_slabs.py 2 3 4 6 2> gpurun_out/stamps.tmp >> gpurun_out/exp_slabs3.txt
  tail -6 gpurun_out/stamps.tmp >> gpurun_out/exp_slabs3.txt
done
cat gpurun_out/exp_slabs3.txt
