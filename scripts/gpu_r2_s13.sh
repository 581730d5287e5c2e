#!/bin/bash
# 2-cell prune level (levels 3 + 2 in one launch): parity tests, then A/B against level 3 only (graph replays, in-graph stamps)
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/gpu_tests.log 2>&1; tail -15 gpurun_out/gpu_tests.log
: > gpurun_out/ab_fine.txt
run() { echo "$*" >> gpurun_out/ab_fine.txt; env "$@" GSDF_AB_GRAPH=1 timeout -k 5 200 python scripts/ab_eval.py 2>&1 | grep -E "Octree|Error|error" >> gpurun_out/ab_fine.txt; }
run GSDF_PRUNE_FINE=1
run GSDF_PRUNE_FINE=0
cat gpurun_out/ab_fine.txt
