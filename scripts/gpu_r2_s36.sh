#!/bin/bash
# Last call: bench.py both arms with half-quad lists on large lattices only (records), then the parity files that cover both list forms.
mkdir -p gpurun_out
timeout -k 5 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-200 gpurun_out/bench_ref.json
timeout -k 5 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err; cut -c1-300 gpurun_out/bench_n1.json
timeout -k 5 600 python -m pytest tests/test_gpu_graph.py tests/test_full_size.py tests/test_jit.py -m gpu -q -x --timeout 500 2>&1 | tail -3
