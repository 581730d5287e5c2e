#!/bin/bash
# Branch-free atan / one-test atan2 prologue + multiply-shift work-item decode (lattice, prune centres): parity suite, A/B, bench line.
mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log
: > gpurun_out/ab_atan.txt
run() { echo "$*" >> gpurun_out/ab_atan.txt; env "$@" GSDF_AB_GRAPH=1 timeout -k 5 400 python scripts/ab_eval.py 2>&1 | grep -E "Octree|Error|error" >> gpurun_out/ab_atan.txt; }
run GSDF_AB_SPECIAL=1
run GSDF_X=interp
run GSDF_AB_SPECIAL=1
cat gpurun_out/ab_atan.txt
timeout -k 5 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err; cut -c1-600 gpurun_out/bench_n1.json
