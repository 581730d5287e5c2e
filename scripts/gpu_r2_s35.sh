#!/bin/bash
# Half-quad work lists for the two-corners-per-thread lattice kernel: whole parity suite; only if green: A/B against
# GSDF_HALF_QUADS=0, an ncu capture of the specialised kernels (warp instructions for roofline_issue), bench.py both arms.
mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/gpu_tests.log 2>&1; rc=$?; tail -3 gpurun_out/gpu_tests.log
[ $rc -ne 0 ] && { tail -40 gpurun_out/gpu_tests.log; exit 1; }
F=gpurun_out/ab_halfquads.txt
: > $F
run() { echo "$*" >> $F; env "$@" GSDF_AB_GRAPH=1 GSDF_AB_SPECIAL=1 timeout -k 5 400 python scripts/ab_eval.py 2>&1 | grep -E "Octree|Error|error" >> $F; }
run GSDF_HALF_QUADS=0
run GSDF_HALF_QUADS=1
cat $F | cut -c1-250
rm -f gpurun_out/prof_k_eval.ncu-rep
ncu --set full --clock-control none --import-source on -k "regex:k_eval|k_jit" -s 6 -c 2 -f -o gpurun_out/prof_k_eval python bench.py --steps 3 --warmup 1 --no-cpu-baseline --device-only > /dev/null 2>&1
timeout -k 5 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-200 gpurun_out/bench_ref.json
timeout -k 5 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err; cut -c1-400 gpurun_out/bench_n1.json
