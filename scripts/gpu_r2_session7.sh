#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/gpu_tests.log 2>&1; tail -12 gpurun_out/gpu_tests.log
echo "---- block (default)"
GSDF_AB_GRAPH=1 timeout -k 5 200 python scripts/ab_eval.py > gpurun_out/ab_block.txt 2>&1; cat gpurun_out/ab_block.txt
echo "---- GSDF_MC=v1"
GSDF_MC=v1 GSDF_AB_GRAPH=1 timeout -k 5 200 python scripts/ab_eval.py > gpurun_out/ab_mcv1.txt 2>&1; grep Octree gpurun_out/ab_mcv1.txt
B="python bench.py --steps 3 --warmup 1 --no-cpu-baseline --device-only"
M=$(cat scripts/ncu_metrics.txt)
ncu --metrics $M --clock-control none -k regex:"k_mc_|k_scan|k_mesh|k_finish" -s 10 -c 4 --csv --log-file gpurun_out/mc_metrics_block.csv $B > /dev/null 2>&1
