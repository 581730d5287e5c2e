#!/bin/bash
# Round 2, under gpurun (1 GPU): launch list of one bench run, a full capture of the dominant kernel (k_eval: centres + lattice)
# and the summariser's metric list (scripts/ncu_metrics.txt) for every other kernel of the step.
# Numbers printed by a run under ncu are never bench values. --device-only keeps the launch list to the renders of the
# timed `value` region (CUDA-graph replays of full-size renders).
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/*.raw.csv gpurun_out/launches.csv
B="python bench.py --steps 3 --warmup 1 --no-cpu-baseline --device-only"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_eval|k_jit" -s 6 -c 2 -f -o gpurun_out/prof_k_eval $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_eval -s 6 -c 2 -f -o gpurun_out/prof_k_eval_interp $B --no-specialize > /dev/null 2>&1
M=$(cat scripts/ncu_metrics.txt)
for k in "k_mc_blk_emit" "k_mc_blk_count" "k_mesh_lists"; do
  ncu --metrics $M --clock-control none -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_$k $B > /dev/null 2>&1
  ncu -i gpurun_out/prof_$k.ncu-rep --page raw --csv > gpurun_out/prof_$k.raw.csv 2>/dev/null && rm -f gpurun_out/prof_$k.ncu-rep
done
du -sh gpurun_out
