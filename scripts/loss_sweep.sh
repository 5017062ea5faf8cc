#!/bin/bash
# Rebuild libsam2b200.so with different per-block work decompositions of the mask-loss kernels and time each (GPU box).
set -e
cd "$(dirname "$0")/.."
for cfg in "4 1" "8 1" "4 2" "2 1"; do
  set -- $cfg
  echo "=== FWD_ROUNDS=$1 BWD_ROUNDS=$2"
  SAM2B200_EXTRA_NVCC_FLAGS="-DSAM2B200_LOSS_FWD_ROUNDS=$1 -DSAM2B200_LOSS_BWD_ROUNDS=$2" python -m sam2_video_training_b200.build --force > /dev/null
  python scripts/loss_kernel_bench.py 10,7,384 8,13,512 1,32,1024 4,32,1024
done
python -m sam2_video_training_b200.build --force > /dev/null
