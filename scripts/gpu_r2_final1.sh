#!/bin/bash
# Final GPU session, part 1: the whole -m gpu suite, then the ncu launch list and captures the bench line's roofline reads.
mkdir -p gpurun_out
timeout -k 5 1800 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/gpu_tests.log 2>&1; tail -4 gpurun_out/gpu_tests.log
timeout -k 5 900 bash scripts/profile_gpu_r2.sh
ls -la gpurun_out | tail -12
