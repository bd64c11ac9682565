#!/bin/bash
# L2 weight prefetch in the GEMM / attn_fused<384> prologues: parity smoke + A/B
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q -k "gemm or 384 or whole_model or 768" 2>&1 | tail -3 ) > gpurun_out/c27_pytest.log; cat gpurun_out/c27_pytest.log
timeout 900 python tools/ab_variants.py --steps 30 nopf base nopf2:nopf base2:base 2>&1 | tee gpurun_out/c27_ab.log | grep -E "^=="
