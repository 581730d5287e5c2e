#!/bin/bash
# Host-side memory-safety check (no GPU): builds the library with AddressSanitizer + UBSan on the host code and runs the
# host-only tests (builder, bounds, flattener, thread / text builders, program model) against it.
#   bash scripts/asan_host.sh            (about 3 minutes: ~90 s build, ~60 s tests)
set -e
cd "$(dirname "$0")/.."
OUT=${ASAN_LIB:-/tmp/libgsdf_asan.so}
make -s -C gsdf_b200/csrc EXTRA="-Xcompiler -fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer -g" OUT=$OUT
GSDF_B200_LIB=$OUT ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0 \
  LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) \
  python -m pytest tests/test_host.py tests/test_progsim.py tests/test_textsdf.py tests/test_special_evaluators.py tests/test_slab_dist.py \
  -x -q -m "not gpu" -p no:cacheprovider
