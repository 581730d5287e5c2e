#!/bin/bash
# self-listing lattice evaluation: parity tests, then A/B against the list-kernel path (graph replays, in-graph stamps)
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/gpu_tests.log 2>&1; tail -5 gpurun_out/gpu_tests.log
: > gpurun_out/ab_selflist.txt
for v in 1 0; do
  echo "GSDF_SELFLIST=$v" >> gpurun_out/ab_selflist.txt
  GSDF_SELFLIST=$v GSDF_AB_GRAPH=1 timeout -k 5 200 python scripts/ab_eval.py 2>&1 | grep -E "Octree|Error|error" >> gpurun_out/ab_selflist.txt
done
cat gpurun_out/ab_selflist.txt
