#!/bin/bash
# ncu --set full of the HBM-bound GEMM launches of one forward (M = 262144 rows): the stage-0 skip-concat GEMM (K = 96 + 96, N = 96)
# and an up-sample 1x1 GEMM (K = 96... N = 192), plus the r = 2 quad combine kernel
mkdir -p gpurun_out
for k in gemm_tn_f16_kernel:gemmcat:49 gemm_tn_f16_kernel:gemmup:46 upsample_combine_quad:upquad:2; do
  IFS=: read -r name short skip <<< "$k"
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:"$name" -s $skip -c 1 \
      -o gpurun_out/prof_r11_$short -f python tools/one_forward.py > gpurun_out/ncu_r11_$short.log 2>&1
done
ls -la gpurun_out/prof_r11_*
