#!/bin/bash
# final profile pass of a round: tools/gpu_final.sh <tag>
#   bench line, config-4 microbench, latency, per-stage block times, ncu launch list, full ncu captures of the hot kernels
tag=${1:-rXX}
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 300 gpurun_out/${tag}_bench.json; tail -2 gpurun_out/${tag}_bench.err
timeout 300 python tools/bench_window_attention_f16.py > gpurun_out/${tag}_wattn.json 2> gpurun_out/${tag}_wattn.err; tail -2 gpurun_out/${tag}_wattn.err
timeout 300 python tools/bench_latency.py > gpurun_out/${tag}_latency.json 2> gpurun_out/${tag}_latency.err; tail -2 gpurun_out/${tag}_latency.err
for d in "96 64" "192 32" "384 16" "768 8"; do python tools/time_block.py $d 100 0; python tools/time_block.py $d 100 1; done 2>&1 | tee gpurun_out/${tag}_block_times.log
timeout 600 bash tools/profile_round.sh $tag launches; tail -1 gpurun_out/ncu_launches_$tag.log
timeout 1500 bash tools/profile_round.sh $tag full
