"""BASELINE config 4: WindowAttention micro-benchmark per stage (win 8, head_dim 12/24/48/96, shifted and unshifted masks).

Times the public module (sunet_tf_b200.WindowAttention.forward -> sunet_window_attention_fwd through the C ABI: fp32 in/out,
qkv projection + attention core + output projection) with CUDA events, B = 64 images worth of windows per stage (stages 2-3 enlarged
to >= 2 waves as SURVEY 8(d) asks), and reports windows/s and TFLOP/s against 512 C^2 + 16384 C FLOPs per window.
Inside the whole-model forward stages 0-1 use the fused kernel instead (attn_fused.cu); its per-launch figures are in bench.py's
"kernels" breakdown.

  python tools/bench_window_attention.py > profiles/rNN_window_attention.json
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import weights as Wt  # noqa: E402  (synthetic weights and the reference mask builder only)
from sunet_tf_b200 import WindowAttention  # noqa: E402

dev = torch.device("cuda:0")
out = []
for C, nW, grid, windows in ((96, 64, 64, 4096), (192, 16, 32, 1024), (384, 4, 16, 1024), (768, 4, 16, 1024)):
    hd = C // 8
    for shifted in (False, True):
        sd_blk = Wt.synth_state_dict(Wt.block_spec("", C, grid, grid, 4 if shifted else 0), seed=3, style="init")
        sd = {k[len("attn."):]: v for k, v in sd_blk.items() if k.startswith("attn.")}
        att = WindowAttention(C, (8, 8), 8, qk_scale=8)
        att.load_state_dict(sd, strict=True)
        att = att.to(dev).eval()
        g = torch.Generator().manual_seed(3)
        x = torch.randn(windows, 64, C, generator=g).to(dev)
        mask = sd_blk["attn_mask"].to(dev) if shifted else None
        for _ in range(3):
            att(x, mask=mask)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            att(x, mask=mask)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        flops = windows * (512.0 * C * C + 16384.0 * C)
        out.append({"C": C, "head_dim": hd, "windows": windows, "mask": "shifted" if shifted else "none", "ms": ms,
                    "windows_per_s": windows / ms * 1e3, "tflops": flops / ms * 1e-9,
                    "frac_of_sustained_tensor_peak": flops / ms * 1e-9 / 1376.8})
print(json.dumps({"what": "WindowAttention module forward (fp32 in/out through the C ABI), CUDA-event timed", "results": out}, indent=1))
