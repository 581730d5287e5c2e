#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_multi_device.py tests/test_gpu_parity.py -m gpu -q -x --timeout 500 > gpurun_out/gpu_tests.log 2>&1; tail -5 gpurun_out/gpu_tests.log
timeout -k 5 400 python bench.py --gpus 1 --steps 20 --warmup 5 --no-knurled > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err
timeout -k 5 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-knurled > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -3 gpurun_out/bench_n2.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_n1.json", "gpurun_out/bench_n2.json"):
    d = json.load(open(f))
    print(f, d["ms_per_step"], json.dumps(d["e2e"]["by_slabs_per_device"]))
    if d.get("evaluate_e2e"): print(json.dumps(d["evaluate_e2e"]))
PY
