"""BASELINE config 5: demo_any_resolution on a 2048 x 2048 input = 225 overlapping 256 x 256 tiles (stride 128), sharded over the ranks.

  python tools/bench_any_resolution.py [size]                     (1 GPU)
  python -m torch.distributed.run --nproc-per-node N ... tools/bench_any_resolution.py

Times sunet_tf_b200.tiles.denoise_any_resolution (tile extraction from the padded canvas, batched forwards, overlap-add fold, one
reduce over ranks, normalise + crop + clamp) with CUDA events, max over ranks, and prints one JSON line (tiles/s, Mpixel/s).
out_chans = 3 as the reference script assumes.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sunet_tf_b200 import SUNet, shard, tiles  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
rank, world, local = shard.world()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    shard.init_process_group("nccl")
torch.manual_seed(0)
net = SUNet(img_size=256, patch_size=4, in_chans=3, out_chans=3, embed_dim=96, depths=[8] * 4, num_heads=[8] * 4, window_size=8, mlp_ratio=4.0,
            qkv_bias=True, qk_scale=8).to(dev).eval()
g = torch.Generator(device=dev)
g.manual_seed(4)
clean = torch.rand(1, 3, size, size, generator=g, device=dev)
noisy = torch.round(torch.clamp(clean + torch.randn(1, 3, size, size, generator=g, device=dev) * (50 / 255.0), 0, 1) * 255) / 255
X, n = tiles.canvas_geometry(size, size)
for _ in range(2):
    tiles.denoise_any_resolution(net, noisy, tile_batch=64, rank=rank, world_size=world)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0.record()
for _ in range(reps):
    out = tiles.denoise_any_resolution(net, noisy, tile_batch=64, rank=rank, world_size=world)
e1.record()
torch.cuda.synchronize()
ms = shard.max_over_ranks(e0.elapsed_time(e1) / reps, dev)
if rank == 0:
    print(json.dumps({"what": "denoise_any_resolution", "image": [size, size], "tiles": n * n, "kernel": 256, "stride": 128, "n_gpus": world,
                      "ms": ms, "tiles_per_s": n * n / ms * 1e3, "mpixel_per_s": size * size / ms * 1e-3,
                      "out_range": [float(out.min()), float(out.max())]}))
if world > 1:
    dist.destroy_process_group()
