#!/bin/bash
# One-GPU profiling pass for a round (run under gpurun after bench.py has exited 0 without ncu):
#   tools/profile_round.sh r03 [launches|full|all]
#     launches -> gpurun_out/launches_dram_<tag>.csv (one forward, per-launch time + DRAM bytes)
#     full     -> gpurun_out/prof_<tag>_<kernel>.ncu-rep (ncu --set full of one launch of each hot kernel)
# ncu matches -k against the function name without template arguments: instances are picked by launch order (-s) inside one
# forward (attn_fused: 8 launches each at C = 96, 192, 384 on the way down; mlp_proj_fused: 8 at 96, 8 at 192; proj_ln_kernel:
# <384,384,LN> then <384,1536> per stage-2 block).
tag=${1:-rXX}
what=${2:-all}
if [ "$what" != full ]; then
  N=$(python tools/one_forward.py | awk '/launches per forward/ {print $4}')
  pat='gemm_tn|proj_ln|attn_|mlp_|layernorm|patch_embed|upsample_combine|tail_'
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"$pat" \
      --launch-skip $N --launch-count $N --csv --log-file gpurun_out/launches_dram_$tag.csv python tools/one_forward.py > gpurun_out/ncu_launches_$tag.log 2>&1
fi
if [ "$what" != launches ]; then
  for k in ${KERNELS:-attn_fused_kernel:attn96:0 attn_fused_kernel:attn192:8 attn_fused_kernel:attn384:16 mlp_proj_fused_kernel:mlp96:0 \
           mlp_proj_fused_kernel:mlp192:8 patch_embed_mma:patchembed:1 \
           tail_finish:tailfinish:1 tail_up_fused:tailup:1 attn_core_tc:attntc:8 mlp_row_kernel:mlprow:16 gemm_tn_f16_kernel:gemm:40 \
           upsample_combine:upcombine:2 layernorm_kernel:mergeln:0 layernorm_kernel:normup:20}; do
    IFS=: read -r name short skip <<< "$k"
    ncu --set full --clock-control none --import-source on -k regex:"$name" -s $skip -c 1 \
        -o gpurun_out/prof_${tag}_$short -f python tools/one_forward.py > gpurun_out/ncu_${tag}_$short.log 2>&1
  done
  ls -la gpurun_out/prof_${tag}_* | head -20
fi
