#!/bin/bash
# Builds the stand-alone CUDA bring-up binaries (run them on a B200 through gpurun): build/test_gemm
set -e
cd "$(dirname "$0")/.."
mkdir -p build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o build/test_gemm \
  sunet_tf_b200/csrc/tests/test_gemm.cu sunet_tf_b200/csrc/gemm_tcgen05.cu sunet_tf_b200/csrc/error.cu -lcuda
echo built build/test_gemm
