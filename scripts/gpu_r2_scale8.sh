#!/bin/bash
# 8-GPU box: multi-device tests, then bench.py at N = 8, 4 (torchrun) as the driver's scaling run does.
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout -k 5 600 python -m pytest tests/test_multi_device.py -m gpu -q -x --timeout 500 > gpurun_out/gpu_tests_multi8.log 2>&1; tail -5 gpurun_out/gpu_tests_multi8.log
for n in 8 4; do
  timeout -k 5 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; cut -c1-220 gpurun_out/bench_n$n.json; tail -2 gpurun_out/bench_n$n.err
done
