#!/bin/bash
# round profile pass (tag r05): bench line, ncu launch list of one forward (time + DRAM bytes), full captures of the hot kernels, config-4 microbench
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r05_bench.json 2> gpurun_out/r05_bench.err; tail -c 600 gpurun_out/r05_bench.json; tail -3 gpurun_out/r05_bench.err
timeout 300 python tools/bench_window_attention_f16.py > gpurun_out/r05_wattn.json 2> gpurun_out/r05_wattn.err; tail -3 gpurun_out/r05_wattn.err
timeout 600 bash tools/profile_round.sh r05 launches; tail -2 gpurun_out/ncu_launches_r05.log
KERNELS="attn_fused_kernel:attn96:0 mlp_proj_fused_kernel:mlp96:0 attn_fused_kernel:attn192:8 mlp_proj_fused_kernel:mlp192:8 attn_fused_kernel:attn384:16 mlp_row_kernel:mlprow:0" timeout 900 bash tools/profile_round.sh r05 full
