#!/bin/bash
# scan fused into the emit pass + first tile without a counter round trip: parity tests, then A/B (graph replays, in-graph stamps)
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/gpu_tests.log 2>&1; tail -5 gpurun_out/gpu_tests.log
: > gpurun_out/ab_scanfused.txt
run() { echo "$*" >> gpurun_out/ab_scanfused.txt; env "$@" GSDF_AB_GRAPH=1 timeout -k 5 200 python scripts/ab_eval.py 2>&1 | grep -E "Octree|Error|error" >> gpurun_out/ab_scanfused.txt; }
run GSDF_SCAN_FUSED=1
run GSDF_SCAN_FUSED=0
run GSDF_SCAN_FUSED=1 GSDF_BLK_WAVES=12
run GSDF_SCAN_FUSED=1 GSDF_BLK_WAVES=16
cat gpurun_out/ab_scanfused.txt
python scripts/exp_r2_pipeline.py 2>&1 | tail -8
