#!/bin/bash
# One gpurun call that re-establishes the measured state at the start of a GPU session (about 4 minutes of box time):
#   gpurun --timeout 420 -- 'bash scripts/gpu_session_start.sh'
# 1. parity tests, 2. smoke, 3. bench (own arm + reference arm), 4. launch list + ncu summaries (scripts/profile_gpu.sh),
# 5. if gsdf_b200/libgsdfb200_rxy.so was built beforehand (scripts/ab_rxy.sh build): the radius-reuse A/B.
# Afterwards, here: python scripts/summarize_profiles.py rNN ; cp gpurun_out/bench_n1.json profiles/rNN_bench_n1.json ...
mkdir -p gpurun_out
timeout -k 5 240 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log
timeout -k 5 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout -k 5 120 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout -k 5 120 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cut -c1-260 gpurun_out/bench_n1.json
timeout -k 5 240 bash scripts/profile_gpu.sh > gpurun_out/profile.log 2>&1; tail -1 gpurun_out/profile.log
if [ -f gsdf_b200/libgsdfb200_rxy.so ]; then timeout -k 5 420 bash scripts/ab_rxy.sh run; fi
