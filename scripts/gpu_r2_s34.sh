#!/bin/bash
# A/B: the last CTA of a render without its system-scope fence behind the counters it stores to mapped host memory (the host reads
# them only behind the stream's completion event): is that fence what a later slab waits ~50 us for while an earlier slab's DMA runs?
mkdir -p gpurun_out
F=gpurun_out/ab_sysfence.txt
: > $F
run() { env "$@" GSDF_AB_SPECIAL=1 timeout -k 5 300 python scripts/exp_r2_slabs.py 1 2 3 2>&1 | grep -E "slabs=|Error|error" | sed "s/^/$(basename ${GSDF_B200_LIB:-default}) /" >> $F; }
run GSDF_X=fence
GSDF_B200_LIB=$PWD/gsdf_b200/libgsdfb200_nf.so run GSDF_B200_LIB=$PWD/gsdf_b200/libgsdfb200_nf.so
run GSDF_X=fence
GSDF_B200_LIB=$PWD/gsdf_b200/libgsdfb200_nf.so run GSDF_B200_LIB=$PWD/gsdf_b200/libgsdfb200_nf.so
cut -c1-270 $F
