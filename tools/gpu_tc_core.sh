#!/bin/bash
# bring-up + A/B of the tcgen05 attention core (stage 3): unit test binary, the pytest cases that touch it, block timing with / without
mkdir -p gpurun_out
timeout 120 build/test_attn_tc 64 > gpurun_out/test_attn_tc.log 2>&1; echo "rc=$?" >> gpurun_out/test_attn_tc.log
cat gpurun_out/test_attn_tc.log
if grep -q MISMATCH gpurun_out/test_attn_tc.log || ! grep -q "rc=0" gpurun_out/test_attn_tc.log; then exit 1; fi
timeout 600 python -m pytest tests -x -q -m gpu -k "tcgen05 or whole_model or bit_reproducible or swin_block_vs_reference_golden" 2>&1 | tail -15
for part in 1 0; do
  python tools/time_block.py 768 8 50 $part
  python tools/time_block.py 768 8 50 $part SUNET_NO_TC_CORE=1
done 2>&1 | tee gpurun_out/tc_core_block_times.log
