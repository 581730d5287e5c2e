#!/bin/bash
# Run under gpurun (1 GPU): launch list of one bench run + full captures of every kernel of the step.
# Numbers printed by a run under ncu are never bench values. --device-only keeps the launch list to the renders of the
# timed `value` region (full-size renders: stage-timed eager loop, then CUDA-graph replays) without the e2e legs' slab renders.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 1 --no-cpu-baseline --device-only > gpurun_out/bench_under_ncu.log 2>&1
for k in "k_eval" "k_mc_emit" "k_mc_count" "k_compact_quads"; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 2 -f -o gpurun_out/prof_$k \
      python bench.py --steps 3 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
done
ls -la gpurun_out
