#!/bin/bash
SUNET_LIB_PATH=$PWD/sunet_tf_b200/variants/libsunet_timing.so SUNET_MLP_TIMING=1 SUNET_MLP_TRACE=1 timeout 300 python tools/one_forward.py 2> gpurun_out/c5_trace.log | tail -1
grep -c "^TR " gpurun_out/c5_trace.log
