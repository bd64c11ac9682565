#!/bin/bash
# Builds the stand-alone CUDA bring-up binaries (run them on a B200 through gpurun): build/test_gemm, build/test_units
set -e
cd "$(dirname "$0")/.."
mkdir -p build
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17"
$NV -o build/test_gemm sunet_tf_b200/csrc/tests/test_gemm.cu sunet_tf_b200/csrc/gemm_tcgen05.cu sunet_tf_b200/csrc/error.cu -lcuda
$NV -o build/test_units sunet_tf_b200/csrc/tests/test_units.cu
echo built build/test_gemm build/test_units
$NV -o build/test_ingest sunet_tf_b200/csrc/tests/test_ingest.cu -lcuda
echo built build/test_ingest
$NV -o build/test_attn_tc sunet_tf_b200/csrc/tests/test_attn_tc.cu sunet_tf_b200/csrc/attn_core_tc.cu sunet_tf_b200/csrc/attn_core.cu sunet_tf_b200/csrc/gemm_tcgen05.cu sunet_tf_b200/csrc/error.cu -lcuda
echo built build/test_attn_tc
