#!/bin/bash
# Build a variant of the library for an A/B run on the GPU box:  tools/ab_build.sh NAME -DMACRO=... -> tools/_alt/NAME.so
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_alt
name=$1; shift
S=steered-mixture-of-experts_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" \
  -o tools/_alt/$name.so $S/pack.cu $S/forward.cu $S/backward.cu $S/exchange.cu $S/metrics.cu $S/ssim_loss.cu
echo built tools/_alt/$name.so
