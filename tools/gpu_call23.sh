#!/bin/bash
# mlp_row: parity of the C=384 tests, phase counters inside the whole forward (timing variant), device time of the block kernels
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q -k "384 or proj_ln" 2>&1 | tail -5 ) > gpurun_out/c23_pytest.log; cat gpurun_out/c23_pytest.log
SUNET_LIB_PATH=$PWD/sunet_tf_b200/variants/libsunet_timing.so SUNET_MLP_TIMING=1 timeout 300 python tools/one_forward.py 2> gpurun_out/c23_timing_model.log | tail -1; grep "mlp_row" gpurun_out/c23_timing_model.log | tail -3
timeout 600 python tools/ab_variants.py --steps 20 row:base 2>&1 | tee gpurun_out/c23_ab.log | grep -E "^==|mlp_fused|attn_fused +16|gemm_tcgen05 +4.83 GF +50.6"
