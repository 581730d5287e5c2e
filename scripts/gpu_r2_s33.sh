#!/bin/bash
# A/B continued: fewer copy CTAs for the device-driven read-back (32 CTAs copied at 48 GB/s, 148 at 37, 296 at 26).
mkdir -p gpurun_out
F=gpurun_out/ab_copyk2.txt
: > $F
run() { env "$@" GSDF_AB_SPECIAL=1 timeout -k 5 300 python scripts/exp_r2_slabs.py 1 2 3 2>&1 | grep -E "slabs=|Error|error" >> $F; }
run GSDF_MULTI_COPYK=1 GSDF_MULTI_COPYK_CTAS=16
run GSDF_MULTI_COPYK=1 GSDF_MULTI_COPYK_CTAS=8
run GSDF_MULTI_COPYK=1 GSDF_MULTI_COPYK_CTAS=4
cut -c1-260 $F
