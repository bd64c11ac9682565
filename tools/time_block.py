"""Device time of one SwinTransformerBlock forward (fp16 stream entry, the kernels the whole model runs) at a SUNet B=64 stage shape,
launched back to back (PDL active; the block's weights stay L2-resident, the inputs rotate over > 140 MB).

  python tools/time_block.py dim grid [reps] [part] [ENV=1 ...]     e.g.  python tools/time_block.py 384 16 50 0 SUNET_NO_ROW_MLP=1
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dim, grid = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 50
part = int(sys.argv[4]) if len(sys.argv) > 4 else 0
for kv in sys.argv[5:]:
    k, v = kv.split("=")
    os.environ[k] = v
from oracle import weights as Wt  # noqa: E402  (synthetic weights only)
from sunet_tf_b200 import _lib, modules  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
B = 64
sd = Wt.synth_state_dict(Wt.block_spec("", dim, grid, grid, 4), seed=dim, style="init")
blk = modules.SwinTransformerBlock(dim, (grid, grid), 8, window_size=8, shift_size=4, qk_scale=8)
blk.load_state_dict(sd, strict=True)
blk = blk.to(dev).eval()
h = blk._handle()
M = B * grid * grid
nbuf = max(2, int(140e6 // (M * dim * 2)) + 1)
xs = [torch.randn(M, dim, device=dev).half() for _ in range(nbuf)]
out = torch.empty(M, dim, device=dev, dtype=torch.float16)
nbytes = lib.sunet_swin_block_f16_workspace_bytes(h, B)
ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
st = _lib.stream_ptr(dev)


def run(i):
    _lib.check(lib.sunet_swin_block_f16(h, ctypes.c_void_p(xs[i % nbuf].data_ptr()), B, part, ctypes.c_void_p(out.data_ptr()),
                                        ctypes.c_void_p(ws.data_ptr()), nbytes, st))


for i in range(5):
    run(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(reps):
    run(i)
e1.record()
torch.cuda.synchronize()
print(f"dim {dim} grid {grid} part {part} {' '.join(sys.argv[5:])}: {e0.elapsed_time(e1) / reps * 1e3:.1f} us per block forward")
