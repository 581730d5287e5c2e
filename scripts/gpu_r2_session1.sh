#!/bin/bash
# Round-2 GPU session 1: parity suite on the restructured library, smoke, bench, RXY A/B.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
timeout -k 5 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/gpu_tests.log 2>&1; tail -15 gpurun_out/gpu_tests.log
timeout -k 5 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout -k 5 200 python bench.py --steps 50 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cut -c1-400 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout -k 5 120 python scripts/ab_eval.py > gpurun_out/ab_default.txt 2>&1; cat gpurun_out/ab_default.txt | grep Octree
export GSDF_B200_LIB=$PWD/gsdf_b200/libgsdfb200_rxy.so
GSDF_RXY=1 timeout -k 5 600 python -m pytest tests -m gpu -q -x --timeout 600 --deselect tests/test_full_size.py > gpurun_out/rxy_tests.log 2>&1; tail -3 gpurun_out/rxy_tests.log
GSDF_RXY=1 timeout -k 5 120 python scripts/ab_eval.py > gpurun_out/ab_rxy_on.txt 2>&1; grep Octree gpurun_out/ab_rxy_on.txt
