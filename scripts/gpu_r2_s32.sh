#!/bin/bash
# A/B: device-driven read-back of the multi-slab driver (GSDF_MULTI_COPYK=1: copy kernels enqueued up front) against the classic
# host-enqueued copies; parity of the mode first.
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_multi_device.py -m gpu -q -x --timeout 800 -k "device_driven" 2>&1 | tail -3
F=gpurun_out/ab_copyk.txt
: > $F
run() { env "$@" GSDF_AB_SPECIAL=1 timeout -k 5 300 python scripts/exp_r2_slabs.py 1 2 3 4 2>&1 | grep -E "slabs=|Error|error" >> $F; }
run GSDF_MULTI_COPYK=0
run GSDF_MULTI_COPYK=1
run GSDF_MULTI_COPYK=1 GSDF_MULTI_COPYK_CTAS=32
run GSDF_MULTI_COPYK=1 GSDF_MULTI_COPYK_CTAS=296
run GSDF_MULTI_COPYK=0
cut -c1-260 $F
