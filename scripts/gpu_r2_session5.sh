#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/gpu_tests.log 2>&1; tail -8 gpurun_out/gpu_tests.log
echo "---- tile5 (default)"
GSDF_AB_GRAPH=1 timeout -k 5 200 python scripts/ab_eval.py > gpurun_out/ab_tile5.txt 2>&1; cat gpurun_out/ab_tile5.txt
echo "---- GSDF_MC_V1"
GSDF_MC_V1=1 GSDF_AB_GRAPH=1 timeout -k 5 200 python scripts/ab_eval.py > gpurun_out/ab_mcv1.txt 2>&1; cat gpurun_out/ab_mcv1.txt
timeout -k 5 300 python scripts/exp_r2_pipeline.py > gpurun_out/exp_pipeline.txt 2>&1; grep "slab of\|slabs= [134] \|timeline" gpurun_out/exp_pipeline.txt | head -30
