#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q -k "block or mlp or model or reproducible" 2>&1 | tail -4 ) > gpurun_out/c6_pytest.log; cat gpurun_out/c6_pytest.log
timeout 600 python tools/ab_variants.py --steps 20 nbprod0 base 2>&1 | grep -E "^==|mlp_fused|attn_fused" | tee gpurun_out/c6_ab.log
SUNET_LIB_PATH=$PWD/sunet_tf_b200/variants/libsunet_timing.so SUNET_MLP_TIMING=1 SUNET_MLP_TRACE=1 timeout 300 python tools/one_forward.py 2> gpurun_out/c6_trace.log | tail -1
grep "mlp_proj_fused" gpurun_out/c6_trace.log | grep -v TRACE | sed -n '1p;9p' | cut -c1-900
