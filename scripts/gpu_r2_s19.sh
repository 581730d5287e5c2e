#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/gpu_tests.log 2>&1; tail -5 gpurun_out/gpu_tests.log
: > gpurun_out/ab_cta2.txt
run() { echo "$*" >> gpurun_out/ab_cta2.txt; env "$@" GSDF_AB_GRAPH=1 timeout -k 5 200 python scripts/ab_eval.py 2>&1 | grep -E "Octree|Error|error" >> gpurun_out/ab_cta2.txt; }
run GSDF_X=default
run GSDF_EVAL_CTA=384
run GSDF_EVAL_CTA=192
run GSDF_EVAL_CTA=256
cat gpurun_out/ab_cta2.txt
python scripts/disasm.py bolt | head -3
