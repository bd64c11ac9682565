#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q -k "block or mlp or model or reproducible" 2>&1 | tail -4 ) > gpurun_out/c3_pytest.log; cat gpurun_out/c3_pytest.log
timeout 600 python tools/ab_variants.py --steps 20 nbprod0 base 2>&1 | grep -E "^==|mlp_fused|attn_fused" | tee gpurun_out/c3_ab.log
