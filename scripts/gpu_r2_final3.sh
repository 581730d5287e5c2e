#!/bin/bash
# Final GPU session of round 2 (1 GPU): ncu launch list + captures, then the whole -m gpu suite, smoke(), bench.py both arms as the
# driver runs them, sanitizer pass.
bash scripts/profile_gpu_r2.sh 2>&1 | tail -3
bash scripts/gpu_r2_final2.sh
