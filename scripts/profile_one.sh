#!/bin/bash
# usage: profile_one.sh <kernel-regex> <skip> <outname>   (under gpurun, 1 GPU)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -f -o gpurun_out/$3 \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/$3.ncu-rep
