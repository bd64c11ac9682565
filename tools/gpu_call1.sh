#!/bin/bash
# round 2, call 1: parity of everything new, A/B of the attention / MLP variants, full bench line, config-4 microbench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/c1_pytest.log
cat gpurun_out/c1_pytest.log | tail -8
if ! grep -q " passed" gpurun_out/c1_pytest.log || grep -q "failed" gpurun_out/c1_pytest.log; then
  echo "== retry without the ping-pong attention kernel"
  ( SUNET_NO_AF_PP=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/c1_pytest_nopp.log
  tail -8 gpurun_out/c1_pytest_nopp.log
fi
timeout 600 python tools/ab_variants.py --steps 20 old:af_old:SUNET_NO_AF_PP=1 remap:base:SUNET_NO_AF_PP=1 pp:base mlp_r24 2>&1 | tee gpurun_out/c1_ab.log
timeout 400 python bench.py > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; tail -c 3000 gpurun_out/c1_bench.json
timeout 300 python tools/bench_window_attention_f16.py > gpurun_out/c1_wattn.json 2> gpurun_out/c1_wattn.err; tail -5 gpurun_out/c1_wattn.err; head -c 1500 gpurun_out/c1_wattn.json
