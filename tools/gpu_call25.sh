#!/bin/bash
# mlp_row with proj + norm2 in front: parity of the C=384 tests, block time against the proj_ln + mlp_row pair, phase counters, bench
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q -k "384 or proj_ln or whole_model" 2>&1 | tail -5 ) > gpurun_out/c25_pytest.log; cat gpurun_out/c25_pytest.log
python tools/time_block.py 384 16 100 0; python tools/time_block.py 384 16 100 0 SUNET_NO_ROW_PROJ=1
SUNET_LIB_PATH=$PWD/sunet_tf_b200/variants/libsunet_timing.so SUNET_MLP_TIMING=1 timeout 300 python tools/one_forward.py 2> gpurun_out/c25_timing_model.log | tail -1; grep "mlp_row" gpurun_out/c25_timing_model.log | tail -3
timeout 600 python tools/ab_variants.py --steps 20 noproj:base:SUNET_NO_ROW_PROJ=1 rowproj:base 2>&1 | tee gpurun_out/c25_ab.log | grep -E "^==|mlp_fused +(38|43.49 GF +37)|attn_fused +16|gemm_tcgen05 +4.83 GF +50.6"
