#!/bin/bash
# ncu --set full of the elementwise kernels (one launch each): the largest upsample_combine, the stage-0 PatchMerging gather + LN, norm_up
mkdir -p gpurun_out
for k in upsample_combine:upcombine:2 layernorm_kernel:mergeln:0 layernorm_kernel:normup:20 layernorm_kernel:ln768:5; do
  IFS=: read -r name short skip <<< "$k"
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:"$name" -s $skip -c 1 \
      -o gpurun_out/prof_r10_$short -f python tools/one_forward.py > gpurun_out/ncu_r10_$short.log 2>&1
done
ls -la gpurun_out/prof_r10_*
