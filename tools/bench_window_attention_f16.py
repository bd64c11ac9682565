"""BASELINE config 4: the window-attention part of a Swin block per stage, stand-alone, on the library's own fp16 token stream.

  python tools/bench_window_attention_f16.py [--reps 20] [--once] > profiles/r04_window_attention_per_stage.json

For every stage (head_dim 12 / 24 / 48 / 96) and for the unshifted (W-MSA) and shifted (SW-MSA, attn_mask) block it times
sunet_swin_block_f16(part = 1): norm1 -> cyclic shift -> window partition -> qkv -> QK^T + relative-position bias (+ mask) ->
softmax -> AV -> window reverse -> un-shift (SUNet_detail.py:233-257, :107-135; proj rides in the MLP kernel of the fused design and
is not part of this measurement - FLOPs per window are therefore 384 C^2 + 16384 C, not the 512 C^2 + 16384 C of the whole module).
Inputs rotate over enough distinct fp16 buffers to exceed the 126 MB L2; time = CUDA events on the launching stream around `reps`
back-to-back calls after 3 warm-ups.  hd = 96 is never shifted in the real model (8x8 grid); its shifted row uses a 16x16 grid
(SURVEY 8(d) config 4).  Stages 2-3 are enlarged to >= 2 waves of the 148-SM grid.
`--once` runs each case once without timing (the form profiled under ncu for the tensor-pipe figures, see tools/ncu_wattn.sh).
"""
import argparse
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import weights as Wt  # noqa: E402  (synthetic parameters only; nothing of the oracle runs here)
from sunet_tf_b200 import _lib, modules  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--once", action="store_true")
ap.add_argument("--peaks", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))
args = ap.parse_args()

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
lib = _lib.load()
peak = 1662.8
if os.path.exists(args.peaks):
    peak = json.load(open(args.peaks))["bf16_tflops"]

# (C, grid side, images, shift)
CASES = [(96, 64, 64, 0), (96, 64, 64, 4), (192, 32, 64, 0), (192, 32, 64, 4), (384, 16, 256, 0), (384, 16, 256, 4),
         (768, 8, 512, 0), (768, 16, 128, 4)]
rows = []
for C, G, B, shift in CASES:
    blk = modules.SwinTransformerBlock(C, (G, G), 8, window_size=8, shift_size=shift, qk_scale=8)
    spec = Wt.block_spec("", C, G, G, shift)
    sd = Wt.synth_state_dict(spec, seed=3, style="stress")
    blk.load_state_dict(sd, strict=True)
    blk = blk.to(dev).eval()
    h = blk._handle()
    M = B * G * G
    nbuf = max(2, int(140e6 // (M * C * 2)) + 1)
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    xs = [torch.randn(M, C, generator=g, device=dev).half() for _ in range(nbuf)]
    out = torch.empty(M, C, device=dev, dtype=torch.float16)
    nbytes = lib.sunet_swin_block_f16_workspace_bytes(h, B)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    st = _lib.stream_ptr(dev)

    def run(i):
        _lib.check(lib.sunet_swin_block_f16(h, ctypes.c_void_p(xs[i % nbuf].data_ptr()), B, 1, ctypes.c_void_p(out.data_ptr()),
                                            ctypes.c_void_p(ws.data_ptr()), nbytes, st))

    if args.once:
        run(0)
        torch.cuda.synchronize()
        continue
    for i in range(3):
        run(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(args.reps):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    windows = M // 64
    flops = windows * (384.0 * C * C + 16384.0 * C)
    rows.append({"C": C, "head_dim": C // 8, "grid": G, "images": B, "windows": windows, "mask": "shifted (SW-MSA)" if shift else "none (W-MSA)",
                 "us": ms * 1e3, "windows_per_s": windows / ms * 1e3, "algorithmic_tflops": flops / ms * 1e-9,
                 "frac_of_burst_tensor_peak": flops / ms * 1e-9 / peak, "stream_gbs": 4.0 * M * C / ms * 1e-6,
                 "launches": 1 if C <= 384 else 3,
                 "kernels": "attn_fused" if C <= 384 else ("layernorm + qkv GEMM (tcgen05) + attn_core_tc (QK^T / PV on tcgen05, S / O in TMEM)" if G == 8 else "layernorm + qkv GEMM (tcgen05) + attn_core (mma.sync; 16x16 grid, not a SUNet shape)")})
    del blk, xs, out, ws
if not args.once:
    print(json.dumps({"what": "window-attention part of a Swin block (norm1 .. attention output, without proj), fp16 token stream, stand-alone",
                      "entry": "sunet_swin_block_f16(part=1)", "flops_per_window": "384 C^2 + 16384 C", "peak_tflops_burst": peak,
                      "l2": "inputs rotate over > 140 MB of distinct buffers", "reps": args.reps, "rows": rows}, indent=1))
