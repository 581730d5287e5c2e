#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_jit.py tests/test_gpu_graph.py -m gpu -q -x --timeout 800 > gpurun_out/gpu_tests_jit.log 2>&1; tail -4 gpurun_out/gpu_tests_jit.log
: > gpurun_out/ab_jit3.txt
run() { echo "$*" >> gpurun_out/ab_jit3.txt; env "$@" GSDF_AB_GRAPH=1 timeout -k 5 400 python scripts/ab_eval.py 2>&1 | grep -E "Octree|Error|error" >> gpurun_out/ab_jit3.txt; }
run GSDF_AB_SPECIAL=1
run GSDF_X=interp
cat gpurun_out/ab_jit3.txt
timeout -k 5 300 python scripts/exp_r2_pipeline.py 2>&1 | tail -6
