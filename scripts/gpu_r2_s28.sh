#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/ab_jit4.txt
run() { echo "$*" >> gpurun_out/ab_jit4.txt; env "$@" GSDF_AB_GRAPH=1 timeout -k 5 400 python scripts/ab_eval.py 2>&1 | grep -E "Octree|Error|error" >> gpurun_out/ab_jit4.txt; }
run GSDF_AB_SPECIAL=1
run GSDF_X=interp
run GSDF_AB_SPECIAL=1
cat gpurun_out/ab_jit4.txt
timeout -k 5 900 python -m pytest tests/test_jit.py tests/test_gpu_parity.py -m gpu -q -x --timeout 800 2>&1 | tail -3
