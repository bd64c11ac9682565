#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/c2_pytest.log; cat gpurun_out/c2_pytest.log
SUNET_LIB_PATH=$PWD/sunet_tf_b200/variants/libsunet_timing.so SUNET_MLP_TIMING=1 SUNET_AF_TIMING=1 timeout 300 python tools/one_forward.py 2> gpurun_out/c2_timing.log | tail -1
grep "attn_fused<96>" gpurun_out/c2_timing.log | head -2
grep "attn_fused<192>" gpurun_out/c2_timing.log | head -1
grep "attn_fused<384>" gpurun_out/c2_timing.log | head -1
grep "mlp_proj_fused" gpurun_out/c2_timing.log | sed -n '1p;2p;9p;10p'
