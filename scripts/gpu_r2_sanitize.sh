#!/bin/bash
# compute-sanitizer on the round-2 kernels: smoke() (block marching cubes, multi-slab driver) and the kernel-variant child.
mkdir -p gpurun_out
timeout -k 5 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"; tail -3 gpurun_out/sanitize_memcheck_smoke.log
timeout -k 5 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"; tail -3 gpurun_out/sanitize_racecheck_smoke.log
GSDF_BLK_GRID=2 timeout -k 5 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/count_pipeline_child.py > gpurun_out/sanitize_memcheck_child.log 2>&1; echo "memcheck child rc=$?"; tail -3 gpurun_out/sanitize_memcheck_child.log
timeout -k 5 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_synccheck_smoke.log 2>&1; echo "synccheck smoke rc=$?"; tail -3 gpurun_out/sanitize_synccheck_smoke.log
