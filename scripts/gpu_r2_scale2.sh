#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_multi_device.py -m gpu -q -x --timeout 500 > gpurun_out/gpu_tests_multi2.log 2>&1; tail -2 gpurun_out/gpu_tests_multi2.log
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; cut -c1-220 gpurun_out/bench_n2.json; tail -2 gpurun_out/bench_n2.err
