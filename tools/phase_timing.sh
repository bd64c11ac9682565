# Phase cycle counters of the fused kernels over one forward (timing build: -DSUNET_KERNEL_TIMING=1; rebuild without it afterwards)
set -e
export SUNET_NVCC_EXTRA=-DSUNET_KERNEL_TIMING=1   # (exported: the flag stamp would otherwise trigger a plain rebuild on the next import)
python -m sunet_tf_b200._build --force > /dev/null 2>&1
SUNET_TAIL_TIMING=1 SUNET_MLP_TIMING=1 SUNET_AF_TIMING=1 python tools/one_forward.py 2> gpurun_out/phase_timing.log | tail -1
grep -c . gpurun_out/phase_timing.log
# one line per kernel kind and stage: the first encoder launch of each
grep "attn_fused<96>" gpurun_out/phase_timing.log | head -1
grep "attn_fused<192>" gpurun_out/phase_timing.log | head -1
grep "attn_fused<384>" gpurun_out/phase_timing.log | head -1
grep -i "mlp" gpurun_out/phase_timing.log | head -1
grep -i "mlp" gpurun_out/phase_timing.log | sed -n 9p
grep "tail_up_fused" gpurun_out/phase_timing.log | head -1
