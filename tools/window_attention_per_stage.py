"""BASELINE config 4 seen inside the whole-model forward: per stage, the device time of the window-attention part of a Swin block
(norm1 + shift/partition + qkv + QK^T/bias/mask/softmax/AV [+ proj where it is a separate launch]) from bench.py's per-launch CUDA-event
records (--profile-json), against the algorithmic FLOPs 8 L C^2 + 256 L C per image.

  python tools/window_attention_per_stage.py gpurun_out/kernels_rNN.json > profiles/rNN_window_attention_per_stage.json
"""
import json
import sys

d = json.load(open(sys.argv[1]))
B = d["batch"]
recs = d["launch_list"]
PEAK = 1376.8
stages = [(96, 4096), (192, 1024), (384, 256), (768, 64)]
# walk the launch list: a Swin block starts at an attn_fused launch (stages 0-1) or at layernorm -> gemm -> attn_core (stages 2-3)
out = {C: {"blocks": 0, "ms": 0.0, "launches_per_block": None} for C, _ in stages}
i = 0
while i < len(recs):
    kind, ms, flops, nbytes = recs[i]
    if kind == "attn_fused":
        C = min((96, 192, 384), key=lambda c: abs((flops / nbytes - 64.0) / 1.5 - c))   # flops / bytes = (6 C^2 + 256 C) / (4 C)
        out[C]["blocks"] += 1
        out[C]["ms"] += ms                    # proj rides in the following mlp_fused launch (not counted here)
        out[C]["launches_per_block"] = "1 (proj fused into the MLP kernel)" if C < 384 else "1 + proj GEMM (not counted)"
        i += 1
    elif kind == "attn_core" and i >= 2 and recs[i - 1][0] == "gemm_tcgen05" and recs[i - 2][0] == "layernorm":
        C = 384 if recs[i][3] / 8.0 / B > 256 * 384 - 1 and recs[i][3] / 8.0 / B < 256 * 384 + 1 else 768
        out[C]["blocks"] += 1
        out[C]["ms"] += recs[i - 2][1] + recs[i - 1][1] + ms + recs[i + 1][1]   # LN1 + qkv GEMM + core + proj GEMM
        out[C]["launches_per_block"] = "4 (layernorm, qkv GEMM, attention core, proj GEMM + residual)"
        i += 2
    else:
        i += 1
res = []
for C, L in stages:
    o = out[C]
    if not o["blocks"]:
        continue
    per_block_ms = o["ms"] / o["blocks"]
    with_proj = C >= 768
    flops = B * ((8.0 if with_proj else 6.0) * L * C * C + 256.0 * L * C)
    res.append({"C": C, "head_dim": C // 8, "tokens_per_image": L, "windows": B * L // 64, "blocks": o["blocks"], "launches_per_block": o["launches_per_block"],
                "ms_per_block": per_block_ms, "algorithmic_gflop_per_block": flops / 1e9, "tflops": flops / per_block_ms * 1e-9,
                "frac_of_sustained_tensor_peak": flops / per_block_ms * 1e-9 / PEAK, "windows_per_s": B * L / 64 / per_block_ms * 1e3})
print(json.dumps({"batch": B, "note": "device time per Swin block of the window-attention part, CUDA events around each launch (bench.py --profile-json)",
                  "stages": res}, indent=1))
