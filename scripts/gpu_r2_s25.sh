#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_jit.py -q -x --timeout 800 > gpurun_out/gpu_tests_jit.log 2>&1; tail -8 gpurun_out/gpu_tests_jit.log
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
: > gpurun_out/ab_jit.txt
run() { echo "$*" >> gpurun_out/ab_jit.txt; env "$@" GSDF_AB_GRAPH=1 timeout -k 5 400 python scripts/ab_eval.py 2>&1 | grep -E "Octree|Error|error|special" >> gpurun_out/ab_jit.txt; }
run GSDF_X=default
run GSDF_AB_SPECIAL=1
run GSDF_AB_SPECIAL=1 GSDF_JIT_CTA=256
cat gpurun_out/ab_jit.txt
