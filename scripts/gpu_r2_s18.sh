#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_gpu_graph.py -m gpu -q -x --timeout 600 > gpurun_out/gpu_tests.log 2>&1; tail -15 gpurun_out/gpu_tests.log
timeout -k 5 1200 python -m pytest tests/test_gpu_parity.py tests/test_multi_device.py tests/test_full_size.py -m gpu -q -x --timeout 900 > gpurun_out/gpu_tests2.log 2>&1; tail -5 gpurun_out/gpu_tests2.log
: > gpurun_out/ab_mcfused.txt
run() { echo "$*" >> gpurun_out/ab_mcfused.txt; env "$@" GSDF_AB_GRAPH=1 timeout -k 5 200 python scripts/ab_eval.py 2>&1 | grep -E "Octree|Error|error" >> gpurun_out/ab_mcfused.txt; }
run GSDF_MC_FUSED=1
run GSDF_MC_FUSED=0
run GSDF_X=default
cat gpurun_out/ab_mcfused.txt
timeout -k 5 300 python scripts/exp_r2_pipeline.py 2>&1 | tail -7
