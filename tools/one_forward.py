"""Two SUNet forwards of the bench workload (B = 64) and nothing else from this library: the target of the ncu launch-list capture.

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -k regex:'gemm_tn|proj_ln|attn_|mlp_|layernorm|patch_embed|upsample_combine|tail_' --launch-skip N --launch-count N --csv \
      --log-file gpurun_out/launches.csv python tools/one_forward.py          (N = launches per forward, printed by this script)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sunet_tf_b200 import SUNet_model  # noqa: E402
from sunet_tf_b200.default_config import DEFAULT_OPT  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = SUNet_model(DEFAULT_OPT).to(dev).eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
x = torch.rand(B, 3, 256, 256, device=dev)
out = torch.empty(B, 1, 256, 256, device=dev)
for _ in range(2):
    model(x, out=out)
torch.cuda.synchronize()
print("launches per forward:", model.swin_unet.launches_per_forward(B))
