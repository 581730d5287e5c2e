#!/bin/bash
# A/B: interpreter CTA size / register cap variants (graph replays, in-graph stage stamps) and the block-kernel grid cap.
mkdir -p gpurun_out
: > gpurun_out/ab_cta.txt
for lib in libgsdfb200.so libgsdfb200_t384_3.so libgsdfb200_t256_4.so libgsdfb200_t256_5.so libgsdfb200_t256_6.so libgsdfb200_t128_8.so libgsdfb200_t128_10.so; do
  [ -f gsdf_b200/$lib ] || continue
  GSDF_AB_GRAPH=1 GSDF_B200_LIB=$PWD/gsdf_b200/$lib timeout -k 5 200 python scripts/ab_eval.py 2>&1 | grep -E "Octree|Error|error" >> gpurun_out/ab_cta.txt
done
for g in 444 592 740 888; do
  echo "GSDF_BLK_GRID=$g" >> gpurun_out/ab_cta.txt
  GSDF_BLK_GRID=$g GSDF_AB_GRAPH=1 timeout -k 5 200 python scripts/ab_eval.py 2>&1 | grep -E "Octree|Error|error" >> gpurun_out/ab_cta.txt
done
cat gpurun_out/ab_cta.txt
