#!/bin/bash
# whole-row fused MLP at C=384 (mlp_row.cu): parity of every test that touches a C=384 block, then A/B against the two-GEMM path
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "384 or proj_ln or whole_model or reproducible or larger_grids" 2>&1 | tail -15 ) > gpurun_out/c22_pytest.log; cat gpurun_out/c22_pytest.log
timeout 600 python tools/ab_variants.py --steps 20 norow:base:SUNET_NO_ROW_MLP=1 row:base 2>&1 | tee gpurun_out/c22_ab.log | grep -E "^==|mlp_fused|gemm_tcgen05 +(19|4\.8)"
