#!/bin/bash
# bench.py at N=1 and N=2 (torchrun) + reference arm, as the driver runs them.
mkdir -p gpurun_out
timeout -k 5 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cut -c1-250 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout -k 5 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; cut -c1-250 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
timeout -k 5 200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-200 gpurun_out/bench_ref.json
