set -e
export SUNET_NVCC_EXTRA=-DSUNET_KERNEL_TIMING=1   # (exported: the flag stamp would otherwise trigger a plain rebuild on the next import)
python -m sunet_tf_b200._build --force > /dev/null 2>&1
SUNET_TAIL_TIMING=1 python tools/one_forward.py 2>&1 | grep -i "tail_up_fused\|launches" | tail -3
SUNET_TAIL_TIMING=1 SUNET_TAIL_NO_SPLIT=1 python tools/one_forward.py 2>&1 | grep -i "tail_up_fused" | tail -1
