#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/gpu_tests.log 2>&1; tail -8 gpurun_out/gpu_tests.log
echo "---- tile5 (default)"
GSDF_AB_GRAPH=1 timeout -k 5 200 python scripts/ab_eval.py > gpurun_out/ab_tile5.txt 2>&1; grep Octree gpurun_out/ab_tile5.txt
echo "---- GSDF_MC_V1"
GSDF_MC_V1=1 GSDF_AB_GRAPH=1 timeout -k 5 200 python scripts/ab_eval.py > gpurun_out/ab_mcv1.txt 2>&1; grep Octree gpurun_out/ab_mcv1.txt
B="python bench.py --steps 3 --warmup 1 --no-cpu-baseline --device-only"
M=$(cat scripts/ncu_metrics.txt)
ncu --metrics $M --clock-control none -k regex:"k_mc_|k_scan|k_compact|k_finish" -s 10 -c 4 --csv --log-file gpurun_out/mc_metrics_tile5.csv $B > /dev/null 2>&1
