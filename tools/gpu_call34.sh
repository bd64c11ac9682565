#!/bin/bash
# final profile pass of the round (tag r06): bench line, config-4 microbench, latency, ncu launch list, full captures of the hot kernels
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r06_bench.json 2> gpurun_out/r06_bench.err; tail -c 300 gpurun_out/r06_bench.json; tail -2 gpurun_out/r06_bench.err
timeout 300 python tools/bench_window_attention_f16.py > gpurun_out/r06_wattn.json 2> gpurun_out/r06_wattn.err; tail -2 gpurun_out/r06_wattn.err
timeout 300 python tools/bench_latency.py > gpurun_out/r06_latency.json 2> gpurun_out/r06_latency.err; tail -2 gpurun_out/r06_latency.err
for d in "96 64" "192 32" "384 16" "768 8"; do python tools/time_block.py $d 100 0; python tools/time_block.py $d 100 1; done 2>&1 | tee gpurun_out/r06_block_times.log
timeout 600 bash tools/profile_round.sh r06 launches; tail -1 gpurun_out/ncu_launches_r06.log
KERNELS="attn_fused_kernel:attn96:0 mlp_proj_fused_kernel:mlp96:0 attn_fused_kernel:attn192:8 mlp_proj_fused_kernel:mlp192:8 attn_fused_kernel:attn384:16 mlp_row_kernel:mlprow:0 gemm_tn_f16_kernel:gemm:20 tail_up_fused:tailup:0 tail_stencil:stencil:0" timeout 1200 bash tools/profile_round.sh r06 full
