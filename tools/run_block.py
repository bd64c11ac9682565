"""Runs SwinTransformerBlock forwards at the SUNet B=64 stage-0 / stage-1 shapes (for ncu captures of the per-block kernels).

  python tools/run_block.py [reps] [shift]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import weights as Wt  # noqa: E402  (synthetic weights only)
from sunet_tf_b200 import SwinTransformerBlock  # noqa: E402

dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
shift = int(sys.argv[2]) if len(sys.argv) > 2 else 4
for dim, grid in ((96, 64), (192, 32), (384, 16)):
    sd = Wt.synth_state_dict(Wt.block_spec("", dim, grid, grid, shift), seed=dim, style="init")
    blk = SwinTransformerBlock(dim, (grid, grid), 8, window_size=8, shift_size=shift, qk_scale=8)
    blk.load_state_dict(sd, strict=True)
    blk = blk.to(dev).eval()
    x = torch.randn(64, grid * grid, dim, device=dev)
    for _ in range(reps):
        y = blk(x)
    torch.cuda.synchronize()
    print(f"dim {dim} grid {grid} shift {shift}: ok, out mean {y.float().mean().item():.4f}")
