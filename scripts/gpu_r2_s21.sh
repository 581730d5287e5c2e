#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/ab_noinl.txt
run() { echo "$*" >> gpurun_out/ab_noinl.txt; env "$@" GSDF_AB_GRAPH=1 timeout -k 5 200 python scripts/ab_eval.py 2>&1 | grep -E "Octree|Error|error" >> gpurun_out/ab_noinl.txt; }
run GSDF_X=default
run GSDF_B200_LIB=$PWD/gsdf_b200/libgsdfb200_noinl.so
run GSDF_B200_LIB=$PWD/gsdf_b200/libgsdfb200_noinl.so GSDF_EVAL_CTA=384
run GSDF_B200_LIB=$PWD/gsdf_b200/libgsdfb200_noinl.so GSDF_EVAL_CTA=256
cat gpurun_out/ab_noinl.txt
timeout -k 5 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "z_slabs or image or full_size or dual or mesh_bit" 2>&1 | tail -3
