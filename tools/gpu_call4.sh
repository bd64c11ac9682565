#!/bin/bash
SUNET_LIB_PATH=$PWD/sunet_tf_b200/variants/libsunet_timing.so SUNET_MLP_TIMING=1 timeout 300 python tools/one_forward.py 2> gpurun_out/c4_timing.log | tail -1
grep "mlp_proj_fused" gpurun_out/c4_timing.log | sed -n '1p;9p'
