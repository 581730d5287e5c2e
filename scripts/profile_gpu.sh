#!/bin/bash
# Run under gpurun (1 GPU): launch list of one bench run, a full capture of the dominant kernel (k_eval: centres + lattice)
# and the summariser's metric list (scripts/ncu_metrics.txt) for every other kernel of the step.
# Numbers printed by a run under ncu are never bench values. --device-only keeps the launch list to the renders of the
# timed `value` region (full-size renders: stage-timed eager loop, then CUDA-graph replays) without the e2e legs' slab renders.
# gpurun merges at most 64 MiB back and every report embeds the module (~17 MB): the top kernel's report travels as is, the
# others are exported to CSV on the box and deleted.
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 1 --no-cpu-baseline --device-only"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_eval -s 8 -c 2 -f -o gpurun_out/prof_k_eval $B > /dev/null 2>&1
M=$(cat scripts/ncu_metrics.txt)
for k in "k_mc_emit" "k_mc_count" "k_compact_quads" "k_scan_lookback" "k_finish_render"; do
  ncu --metrics $M --clock-control none -k regex:$k -s 4 -c 1 -f -o gpurun_out/prof_$k $B > /dev/null 2>&1
  ncu -i gpurun_out/prof_$k.ncu-rep --page raw --csv > gpurun_out/prof_$k.raw.csv 2>/dev/null && rm -f gpurun_out/prof_$k.ncu-rep
done
du -sh gpurun_out
