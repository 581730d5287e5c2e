#!/bin/bash
# ncu on the marching-cubes kernels, both variants (tile5 default, GSDF_MC_V1): launch list + metric list, CSV only.
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 1 --no-cpu-baseline --device-only"
M=$(cat scripts/ncu_metrics.txt)
for v in tile5 v1; do
  if [ $v = v1 ]; then export GSDF_MC_V1=1; else unset GSDF_MC_V1; fi
  ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$v.csv $B > gpurun_out/bench_under_ncu_$v.log 2>&1
  ncu --metrics $M --clock-control none -k regex:"k_mc_|k_scan|k_compact|k_finish" -s 14 -c 6 --csv --log-file gpurun_out/mc_metrics_$v.csv $B > /dev/null 2>&1
done
ls -la gpurun_out | head -30
