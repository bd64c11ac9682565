#!/bin/bash
# parity of the re-indexed elementwise kernels + in-call A/B: tools/gpu_elem_ab.sh "<pytest -k expr>" <variant> [more variants]
mkdir -p gpurun_out
k="$1"; shift
timeout 900 python -m pytest tests -x -q -m gpu -k "$k" 2>&1 | tail -4
timeout 900 python tools/ab_variants.py --steps 30 base "$@" new2:base 2>&1 | grep -v "^attn_fused\|^mlp_fused\|^gemm" | tee gpurun_out/ab_elem.log
