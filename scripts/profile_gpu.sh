#!/bin/bash
# Run under gpurun (1 GPU): launch list of one bench run + full captures of the two heaviest kernels.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_eval -s 8 -c 2 -f -o gpurun_out/prof_eval \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mc_emit -s 3 -c 1 -f -o gpurun_out/prof_emit \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_mc_count -s 3 -c 1 -f -o gpurun_out/prof_count \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out
