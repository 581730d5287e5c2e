#!/bin/bash
# Two corners per thread as the default of the specialised lattice kernel, circular-array Sincos tables, emit-pass row look-up
# once per row + dp4a slot sums: whole parity suite, then A/B (graph replays, in-graph stamps).
mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log
F=gpurun_out/ab_tables.txt
: > $F
run() { echo "$*" >> $F; env "$@" GSDF_AB_GRAPH=1 timeout -k 5 400 python scripts/ab_eval.py 2>&1 | grep -E "Octree|Error|error" >> $F; }
run GSDF_AB_SPECIAL=1
run GSDF_AB_SPECIAL=1 GSDF_EVAL_P=4
run GSDF_X=interp
run GSDF_AB_SPECIAL=1
cat $F
