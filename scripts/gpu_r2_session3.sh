#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/gpu_tests.log 2>&1; tail -5 gpurun_out/gpu_tests.log
timeout -k 5 300 python scripts/exp_r2_pipeline.py > gpurun_out/exp_pipeline.txt 2>&1; cat gpurun_out/exp_pipeline.txt
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --no-knurled > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; cut -c1-300 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
