#!/bin/bash
# Final GPU session, part 1: the ncu launch list and captures the bench line's roofline reads (specialised and interpreter kernels).
mkdir -p gpurun_out
timeout -k 5 1200 bash scripts/profile_gpu_r2.sh
ls -la gpurun_out | grep -E "prof_|launches" 
