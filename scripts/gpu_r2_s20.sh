#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_graph.py tests/test_multi_device.py -m gpu -q -x --timeout 600 > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log
GSDF_AB_GRAPH=1 timeout -k 5 200 python scripts/ab_eval.py 2>&1 | grep -E "Octree|Error|error" > gpurun_out/ab_emitprefetch.txt
cat gpurun_out/ab_emitprefetch.txt
