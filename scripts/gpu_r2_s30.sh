#!/bin/bash
# A/B: two corners per thread in the run-time compiled lattice kernel; emit pass with side-by-side row loads and with a register cap
# for 6 / 7 resident CTAs per SM (variant builds libgsdfb200_e6.so / _e8.so: make OUT=../libgsdfb200_e6.so EXTRA=-DGSDF_EMIT_MINB=6).
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_gpu_graph.py -m gpu -q -x --timeout 800 -k "SPECIALIZE" 2>&1 | tail -3
F=gpurun_out/ab_p2_emit.txt
: > $F
run() { echo "$*" >> $F; env "$@" GSDF_AB_GRAPH=1 timeout -k 5 400 python scripts/ab_eval.py 2>&1 | grep -E "Octree|Error|error" >> $F; }
run GSDF_AB_SPECIAL=1
run GSDF_AB_SPECIAL=1 GSDF_EVAL_P=2
run GSDF_AB_SPECIAL=1 GSDF_EVAL_P=2 GSDF_JIT_CTA=256
run GSDF_X=interp
run GSDF_EMIT_EAGER=1
run GSDF_B200_LIB=$PWD/gsdf_b200/libgsdfb200_e6.so
run GSDF_B200_LIB=$PWD/gsdf_b200/libgsdfb200_e6.so GSDF_EMIT_EAGER=1
run GSDF_B200_LIB=$PWD/gsdf_b200/libgsdfb200_e8.so
run GSDF_B200_LIB=$PWD/gsdf_b200/libgsdfb200_e8.so GSDF_EMIT_EAGER=1
cat $F
