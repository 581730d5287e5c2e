#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/gpu_tests.log 2>&1; tail -8 gpurun_out/gpu_tests.log
timeout -k 5 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
