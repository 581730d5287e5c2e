#!/bin/bash
# Round-2 GPU session 2 (2 GPUs): parity suite with two devices, bench N=1, N=2 (torchrun), reference arm.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
timeout -k 5 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/gpu_tests.log 2>&1; tail -15 gpurun_out/gpu_tests.log
timeout -k 5 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout -k 5 300 python bench.py --steps 50 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cut -c1-300 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; cut -c1-300 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
timeout -k 5 120 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-200 gpurun_out/bench_ref.json
