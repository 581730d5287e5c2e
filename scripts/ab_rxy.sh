#!/bin/bash
# A/B of the experimental radius-reuse build (include/gsdf_program.h, "Radius reuse").
#   here (no GPU):   bash scripts/ab_rxy.sh build       -> gsdf_b200/libgsdfb200_rxy.so (git-ignored, travels with gpurun)
#   on the GPU box:  bash scripts/ab_rxy.sh run         -> parity tests and stage timings with the variant, then the default
# The flattener side is already proven on the CPU model (tests/test_progsim.py::test_radius_reuse_programs_are_bit_identical_and_gated);
# the device side (interp.cuh under #ifdef GSDF_RXY) has NOT run on a GPU yet.
set -e
cd "$(dirname "$0")/.."
case "$1" in
build)
  make -s -C gsdf_b200/csrc EXTRA=-DGSDF_RXY OUT=../libgsdfb200_rxy.so
  ls -la gsdf_b200/libgsdfb200_rxy.so ;;
run)
  mkdir -p gpurun_out
  export GSDF_B200_LIB=$PWD/gsdf_b200/libgsdfb200_rxy.so
  GSDF_RXY=1 timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/rxy_tests.log 2>&1 || true
  tail -3 gpurun_out/rxy_tests.log
  GSDF_RXY=1 timeout 120 python scripts/ab_eval.py > gpurun_out/ab_rxy_on.txt 2>&1 || true
  GSDF_RXY=0 timeout 120 python scripts/ab_eval.py > gpurun_out/ab_rxy_off.txt 2>&1 || true
  unset GSDF_B200_LIB
  timeout 120 python scripts/ab_eval.py > gpurun_out/ab_default.txt 2>&1 || true
  grep -h Octree gpurun_out/ab_rxy_on.txt gpurun_out/ab_rxy_off.txt gpurun_out/ab_default.txt ;;
*) echo "usage: $0 build|run"; exit 2 ;;
esac
