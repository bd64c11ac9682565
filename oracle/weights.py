"""Deterministic synthetic weights and inputs shared by the golden generator, the tests and bench.py.

TEST INFRASTRUCTURE ONLY.  The reference publishes no checkpoint and the full model is 99.7 M parameters
(400 MB), so fixtures cannot carry weights: every tensor is regenerated from (seed, crc32(key)) with a CPU
torch.Generator, which is reproducible across machines for the same torch build.  `state_dict_spec`
mirrors the reference's parameter/buffer tree (model/SUNet_detail.py) key-for-key; oracle/make_golden.py
asserts that against the live reference's own state_dict before any golden is written.

Styles:
  "init"   - the distributions of the reference's random init (SUNet_detail.py:688-695 + torch defaults):
             Linear ~ trunc_normal(std .02) / bias 0, LayerNorm 1/0, bias table ~ trunc_normal(std .02),
             Conv2d ~ U(+-1/sqrt(fan_in)), PReLU 0.25.
  "stress" - non-trivial Linear biases and LayerNorm affines, random PReLU slopes and a std-0.5 bias table, so that
             bias / mask / affine / slope mistakes are visible.  Linear weights keep std .02: with QK_SCALE 8 larger
             q/k weights make softmax near-argmax and the fp32 reference itself chaotic (two fp32 evaluation orders
             of the reference then differ by >1e-2), which would make parity meaningless.
"""
import math
import zlib

import torch

from .sunet_oracle import rel_pos_index, shift_mask

WS = 8


def _block_spec(pre, dim, H, W, shift, ws=WS):
    spec = []
    if min(H, W) <= ws:
        shift = 0
    if shift > 0:
        spec.append((pre + "attn_mask", ((H // ws) * (W // ws), ws * ws, ws * ws), "mask", (H, W, shift)))
    spec += [
        (pre + "norm1.weight", (dim,), "ln_w", None),
        (pre + "norm1.bias", (dim,), "ln_b", None),
        (pre + "attn.relative_position_bias_table", ((2 * ws - 1) ** 2, 8), "table", None),
        (pre + "attn.relative_position_index", (ws * ws, ws * ws), "index", None),
        (pre + "attn.qkv.weight", (3 * dim, dim), "qkv_w", None),
        (pre + "attn.qkv.bias", (3 * dim,), "lin_b", None),
        (pre + "attn.proj.weight", (dim, dim), "lin_w", None),
        (pre + "attn.proj.bias", (dim,), "lin_b", None),
        (pre + "norm2.weight", (dim,), "ln_w", None),
        (pre + "norm2.bias", (dim,), "ln_b", None),
        (pre + "mlp.fc1.weight", (4 * dim, dim), "lin_w", None),
        (pre + "mlp.fc1.bias", (4 * dim,), "lin_b", None),
        (pre + "mlp.fc2.weight", (dim, 4 * dim), "lin_w", None),
        (pre + "mlp.fc2.bias", (dim,), "lin_b", None),
    ]
    return spec


def block_spec(pre, dim, H, W, shift, num_heads=8):
    spec = _block_spec(pre, dim, H, W, shift)
    return [(k, (s[0], num_heads) if kind == "table" else s, kind, extra) for k, s, kind, extra in spec]


def upsample_spec(pre, C, r):
    if r == 2:
        return [
            (pre + "conv.weight", (C // 2, C, 1, 1), "conv_w", None),
            (pre + "up_p.0.weight", (2 * C, C, 1, 1), "conv_w", None),
            (pre + "up_p.1.weight", (1,), "prelu", None),
            (pre + "up_p.3.weight", (C // 2, C // 2, 1, 1), "conv_w", None),
            (pre + "up_b.0.weight", (C, C, 1, 1), "conv_w", None),
            (pre + "up_b.0.bias", (C,), "conv_b", C),
            (pre + "up_b.1.weight", (1,), "prelu", None),
            (pre + "up_b.3.weight", (C // 2, C, 1, 1), "conv_w", None),
        ]
    return [
        (pre + "conv.weight", (C, 2 * C, 1, 1), "conv_w", None),
        (pre + "up_p.0.weight", (16 * C, C, 1, 1), "conv_w", None),
        (pre + "up_p.1.weight", (1,), "prelu", None),
        (pre + "up_p.3.weight", (C, C, 1, 1), "conv_w", None),
        (pre + "up_b.0.weight", (C, C, 1, 1), "conv_w", None),
        (pre + "up_b.0.bias", (C,), "conv_b", C),
        (pre + "up_b.1.weight", (1,), "prelu", None),
        (pre + "up_b.3.weight", (C, C, 1, 1), "conv_w", None),
    ]


def merging_spec(pre, dim):
    return [
        (pre + "reduction.weight", (2 * dim, 4 * dim), "lin_w", None),
        (pre + "norm.weight", (4 * dim,), "ln_w", None),
        (pre + "norm.bias", (4 * dim,), "ln_b", None),
    ]


def patch_embed_spec(pre, in_chans, E, patch=4):
    return [
        (pre + "proj.weight", (E, in_chans, patch, patch), "conv_w", None),
        (pre + "proj.bias", (E,), "conv_b", in_chans * patch * patch),
        (pre + "norm.weight", (E,), "ln_w", None),
        (pre + "norm.bias", (E,), "ln_b", None),
    ]


def sunet_spec(pre="swin_unet.", in_chans=3, out_chans=1, E=96, depths=(8, 8, 8, 8), heads=(8, 8, 8, 8), img=256, patch=4):
    """All state_dict entries of SUNet in registration order (SUNet_detail.py:598-684)."""
    G = img // patch
    nl = len(depths)
    spec = [
        (pre + "prelu.weight", (1,), "prelu", None),
        (pre + "conv_first.weight", (E, in_chans, 3, 3), "conv_w", None),
        (pre + "conv_first.bias", (E,), "conv_b", in_chans * 9),
    ]
    spec += patch_embed_spec(pre + "patch_embed.", E, E, patch)
    for i in range(nl):
        dim, H = E * 2 ** i, G // 2 ** i
        for j in range(depths[i]):
            spec += block_spec(f"{pre}layers.{i}.blocks.{j}.", dim, H, H, 0 if j % 2 == 0 else WS // 2, heads[i])
        if i < nl - 1:
            spec += merging_spec(f"{pre}layers.{i}.downsample.", dim)
    ups, cats = [], []
    for inx in range(nl):
        i = nl - 1 - inx
        dim, H = E * 2 ** i, G // 2 ** i
        if inx == 0:
            ups += upsample_spec(f"{pre}layers_up.0.", dim, 2)
            continue
        cats += [(f"{pre}concat_back_dim.{inx}.weight", (dim, 2 * dim), "lin_w", None),
                 (f"{pre}concat_back_dim.{inx}.bias", (dim,), "lin_b", None)]
        for j in range(depths[i]):
            ups += block_spec(f"{pre}layers_up.{inx}.blocks.{j}.", dim, H, H, 0 if j % 2 == 0 else WS // 2, heads[i])
        if inx < nl - 1:
            ups += upsample_spec(f"{pre}layers_up.{inx}.upsample.", dim, 2)
    spec += ups + cats
    spec += [
        (pre + "norm.weight", (E * 2 ** (nl - 1),), "ln_w", None),
        (pre + "norm.bias", (E * 2 ** (nl - 1),), "ln_b", None),
        (pre + "norm_up.weight", (E,), "ln_w", None),
        (pre + "norm_up.bias", (E,), "ln_b", None),
    ]
    spec += upsample_spec(pre + "up.", E, 4)
    spec += [(pre + "output.weight", (out_chans, E, 3, 3), "conv_w", None)]
    return spec


def _gen(key, seed):
    g = torch.Generator()
    g.manual_seed((int(seed) * 1000003 + zlib.crc32(key.encode())) % (2 ** 63 - 1))
    return g


def _trunc_normal(shape, std, g):
    t = torch.randn(shape, generator=g) * std
    return torch.clamp(t, -2.0, 2.0)  # same absolute cut-offs as trunc_normal_(a=-2, b=2); never active at these stds


def make_tensor(key, shape, kind, extra, seed, style):
    g = _gen(key, seed)
    stress = style == "stress"
    if kind == "lin_w":
        return _trunc_normal(shape, 0.02, g)
    if kind == "qkv_w":
        return _trunc_normal(shape, 0.02, g)
    if kind == "lin_b":
        return torch.randn(shape, generator=g) * 0.02 if stress else torch.zeros(shape)
    if kind == "ln_w":
        return 1.0 + 0.1 * torch.randn(shape, generator=g) if stress else torch.ones(shape)
    if kind == "ln_b":
        return 0.05 * torch.randn(shape, generator=g) if stress else torch.zeros(shape)
    if kind == "table":
        return _trunc_normal(shape, 0.5 if stress else 0.02, g)
    if kind == "conv_w":
        fan_in = shape[1] * shape[2] * shape[3]
        bound = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * bound
    if kind == "conv_b":
        bound = 1.0 / math.sqrt(extra)
        return (torch.rand(shape, generator=g) * 2 - 1) * bound
    if kind == "prelu":
        return torch.full(shape, 0.25) if not stress else 0.1 + 0.3 * torch.rand(shape, generator=g)
    if kind == "index":
        return rel_pos_index(WS)
    if kind == "mask":
        H, W, shift = extra
        return shift_mask(H, W, WS, shift)
    raise ValueError(kind)


OUTLIER_GAIN = 50.0        # a few LayerNorm gains per block, as trained Swin models grow on their massive-activation channels
OUTLIER_CHANNEL = 7
OUTLIER_BIAS = 300.0       # one residual-stream channel carried at O(10^2): fp16 ulp there is 0.25


def synth_state_dict(spec, seed=0, style="init"):
    """style "init": the reference's initialisation; "stress": every parameter class perturbed so that bias / mask / affine
    errors are observable; "outlier": "stress" plus the magnitude pattern of trained checkpoints that random init never
    shows: three channels of every block's norm1 / norm2 gain x50, and stream channel 7 offset by 300 at the
    patch embedding, so that the fp16 token stream and the fp16 MMA operands see large-magnitude values."""
    base = "stress" if style == "outlier" else style
    sd = {k: make_tensor(k, shape, kind, extra, seed, base) for k, shape, kind, extra in spec}
    if style == "outlier":
        for k in sd:
            if k.endswith(("norm1.weight", "norm2.weight")):
                idx = torch.randint(0, sd[k].numel(), (3,), generator=_gen(k + "#outlier", seed))
                sd[k][idx] *= OUTLIER_GAIN
            elif k.endswith("patch_embed.norm.bias"):
                sd[k][OUTLIER_CHANNEL] = OUTLIER_BIAS
    return sd


def awgn_input(batch, seed=1, size=256, sigma=50.0, chans=3, quantize=True):
    """BASELINE.md config 2 input: clean ~ U[0,1], AWGN sigma/255, then the 8-bit round trip of demo.py:70-72.
    Returns (noisy, clean)."""
    g = torch.Generator().manual_seed(seed)
    clean = torch.rand(batch, chans, size, size, generator=g)
    noisy = clean + torch.randn(batch, chans, size, size, generator=g) * (sigma / 255.0)
    if quantize:
        noisy = torch.round(torch.clamp(noisy, 0, 1) * 255.0) / 255.0
    return noisy, clean


def luminance(rgb):
    """train.py:329 target for the 1-channel head."""
    return 0.2989 * rgb[:, 0:1] + 0.5870 * rgb[:, 1:2] + 0.1140 * rgb[:, 2:3]
