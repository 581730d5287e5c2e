#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/gpu_tests.log 2>&1; tail -5 gpurun_out/gpu_tests.log
timeout -k 5 300 python scripts/exp_r2_pipeline.py > gpurun_out/exp_pipeline.txt 2>&1; cat gpurun_out/exp_pipeline.txt
echo "---- GSDF_EVAL_P=4"
GSDF_EVAL_P=4 timeout -k 5 300 python scripts/exp_r2_pipeline.py > gpurun_out/exp_pipeline_p4.txt 2>&1; grep "slab of\|slabs= [1348] reb.*=1" gpurun_out/exp_pipeline_p4.txt
