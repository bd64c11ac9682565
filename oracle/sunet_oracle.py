"""CPU oracle: a functional fp32 restatement of the SUNet forward (mehrdad78/SUNet_TF).

TEST INFRASTRUCTURE ONLY - imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; the product path (sunet_tf_b200/) never imports it.

Parity pinning: the reference ships no tests or golden vectors, so this file is
pinned against the reference ITSELF: oracle/make_golden.py imports the unmodified reference in the
build container, runs it on seeded weights/inputs and commits the outputs under tests/golden/;
tests/test_oracle.py checks this restatement against those fixtures (and against the live reference
when /root/reference is present).

Every function works on a flat ``state_dict`` (name -> tensor) plus a key prefix and cites the
reference lines it restates.  It deliberately uses explicit index maps (gathers, closed-form bias /
mask indices, hand-written bilinear taps) instead of the view/permute/roll/F.interpolate chain of the
reference, so that it also documents the addressing the CUDA kernels implement.
"""
import math

import torch
import torch.nn.functional as F

WS = 8  # window size of the training.yaml architecture (training.yaml:9)


# ----------------------------------------------------------------------------- small pieces
def layer_norm(x, w, b):
    """nn.LayerNorm over the last dim, biased variance, eps 1e-5 (SUNet_detail.py:192,198,299,544,677-678)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + 1e-5) * w + b


def gelu_erf(x):
    """nn.GELU() default = exact erf form (SUNet_detail.py:9,14)."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def prelu(x, a):
    """nn.PReLU() with one shared slope (SUNet_detail.py:345,350,356,361)."""
    return torch.where(x >= 0, x, a * x)


def rel_pos_index(ws=WS):
    """relative_position_index (SUNet_detail.py:87-97) in closed form: (ri-rj+ws-1)*(2ws-1) + (ci-cj+ws-1)."""
    t = torch.arange(ws * ws)
    r, c = t // ws, t % ws
    return (r[:, None] - r[None, :] + ws - 1) * (2 * ws - 1) + (c[:, None] - c[None, :] + ws - 1)


def shift_mask(H, W, ws=WS, shift=WS // 2):
    """SW-MSA mask (SUNet_detail.py:202-221) in closed form.

    After the roll by -shift, window (wr, wc) holds tokens that wrapped around only if it is in the last
    window row / column; inside such a window the tokens with r >= ws-shift (c >= ws-shift) come from the
    other side of the image.  Pairs from different regions get -100 (NOT -inf, :221).
    """
    nWr, nWc = H // ws, W // ws
    t = torch.arange(ws * ws)
    r, c = t // ws, t % ws
    rb = (r >= ws - shift)
    cb = (c >= ws - shift)
    row_split = rb[:, None] != rb[None, :]
    col_split = cb[:, None] != cb[None, :]
    mask = torch.zeros(nWr, nWc, ws * ws, ws * ws)
    mask[nWr - 1, :, :, :] = torch.where(row_split, -100.0, 0.0)
    last_col = torch.where(col_split, -100.0, 0.0)
    mask[:, nWc - 1, :, :] = torch.minimum(mask[:, nWc - 1, :, :], last_col)
    return mask.reshape(nWr * nWc, ws * ws, ws * ws)


def window_token_index(H, W, shift, ws=WS):
    """Flat token index (into L = H*W) of token t of window w after the cyclic shift.

    Restates roll(-shift) + window_partition (SUNet_detail.py:236-244, :27-39): window (wr, wc), token
    (r, c) reads source ((wr*ws + r + shift) mod H, (wc*ws + c + shift) mod W).  window_reverse + roll(+shift)
    (:250-257) scatters back through the same map.  Returns LongTensor (nW, ws*ws).
    """
    nWr, nWc = H // ws, W // ws
    wr = torch.arange(nWr)[:, None, None, None]
    wc = torch.arange(nWc)[None, :, None, None]
    r = torch.arange(ws)[None, None, :, None]
    c = torch.arange(ws)[None, None, None, :]
    src_r = (wr * ws + r + shift) % H
    src_c = (wc * ws + c + shift) % W
    return (src_r * W + src_c).reshape(nWr * nWc, ws * ws)


# ----------------------------------------------------------------------------- modules
def window_attention(sd, pre, xw, mask, num_heads, scale):
    """WindowAttention.forward (SUNet_detail.py:107-138).  xw: (B_, N, C); mask: (nW, N, N) or None."""
    B_, N, C = xw.shape
    hd = C // num_heads
    qkv = F.linear(xw, sd[pre + "qkv.weight"], sd.get(pre + "qkv.bias"))
    # last dim is ordered [3][heads][hd] (:114)
    q = qkv[..., 0 * C:1 * C].reshape(B_, N, num_heads, hd).transpose(1, 2) * scale  # :117
    k = qkv[..., 1 * C:2 * C].reshape(B_, N, num_heads, hd).transpose(1, 2)
    v = qkv[..., 2 * C:3 * C].reshape(B_, N, num_heads, hd).transpose(1, 2)
    s = q @ k.transpose(-1, -2)                                                      # :118
    ws = int(round(math.sqrt(N)))
    table = sd[pre + "relative_position_bias_table"]                                 # ((2ws-1)^2, heads)
    idx = sd[pre + "relative_position_index"] if pre + "relative_position_index" in sd else rel_pos_index(ws)
    bias = table[idx.reshape(-1).long()].reshape(N, N, num_heads).permute(2, 0, 1)   # :120-122
    s = s + bias[None]
    if mask is not None:                                                             # :125-128
        nW = mask.shape[0]
        s = (s.reshape(B_ // nW, nW, num_heads, N, N) + mask[None, :, None]).reshape(B_, num_heads, N, N)
    p = torch.softmax(s, dim=-1)                                                     # :129-131
    o = (p @ v).transpose(1, 2).reshape(B_, N, C)                                    # :135
    return F.linear(o, sd[pre + "proj.weight"], sd[pre + "proj.bias"])               # :136


def mlp(sd, pre, x):
    """Mlp.forward (SUNet_detail.py:18-24); dropouts are identity in eval."""
    h = gelu_erf(F.linear(x, sd[pre + "fc1.weight"], sd[pre + "fc1.bias"]))
    return F.linear(h, sd[pre + "fc2.weight"], sd[pre + "fc2.bias"])


def swin_block(sd, pre, x, H, W, num_heads, shift, scale, ws=WS):
    """SwinTransformerBlock.forward (SUNet_detail.py:227-264); DropPath is identity in eval."""
    B, L, C = x.shape
    if min(H, W) <= ws:          # :186-189
        shift = 0
    idx = window_token_index(H, W, shift, ws)                   # (nW, 64)
    xn = layer_norm(x, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"])
    xw = xn[:, idx.reshape(-1), :].reshape(B * idx.shape[0], ws * ws, C)
    mask = None
    if shift > 0:
        mask = sd[pre + "attn_mask"] if pre + "attn_mask" in sd else shift_mask(H, W, ws, shift)
    aw = window_attention(sd, pre + "attn.", xw, mask, num_heads, scale)
    attn_out = torch.empty_like(x)
    attn_out[:, idx.reshape(-1), :] = aw.reshape(B, L, C)       # reverse + roll back
    x = x + attn_out                                            # :261
    y = layer_norm(x, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"])
    return x + mlp(sd, pre + "mlp.", y)                         # :262


def patch_merging(sd, pre, x, H, W):
    """PatchMerging.forward (SUNet_detail.py:301-322): 2x2 gather, channel order TL,BL,TR,BR (:312-316)."""
    B, L, C = x.shape
    g = x.reshape(B, H // 2, 2, W // 2, 2, C)                   # (b, h2, dh, w2, dw, c)
    cat = torch.cat([g[:, :, 0, :, 0], g[:, :, 1, :, 0], g[:, :, 0, :, 1], g[:, :, 1, :, 1]], dim=-1)
    cat = cat.reshape(B, (H // 2) * (W // 2), 4 * C)
    cat = layer_norm(cat, sd[pre + "norm.weight"], sd[pre + "norm.bias"])
    return F.linear(cat, sd[pre + "reduction.weight"])


def bilinear_taps(n_in, r):
    """nn.Upsample(scale_factor=r, mode='bilinear', align_corners=False) along one axis (:351,:362).
    src = max((d+0.5)/r - 0.5, 0); i0 = floor(src); i1 = min(i0+1, n-1); lam = src - i0."""
    d = torch.arange(n_in * r, dtype=torch.float32)
    src = torch.clamp((d + 0.5) / r - 0.5, min=0.0)
    i0 = src.floor().long()
    i1 = torch.clamp(i0 + 1, max=n_in - 1)
    lam = src - i0.float()
    return i0, i1, lam


def bilinear_up_nhwc(x, r):
    """x: (B, H, W, C) -> (B, rH, rW, C), separable taps as above."""
    B, H, W, C = x.shape
    i0, i1, lh = bilinear_taps(H, r)
    x = x[:, i0] * (1 - lh)[None, :, None, None] + x[:, i1] * lh[None, :, None, None]
    j0, j1, lw = bilinear_taps(W, r)
    return x[:, :, j0] * (1 - lw)[None, None, :, None] + x[:, :, j1] * lw[None, None, :, None]


def pixel_shuffle_nhwc(x, r):
    """nn.PixelShuffle(r): out[b, h*r+i, w*r+j, c] = in[b, h, w, c*r*r + i*r + j] (:346,:357)."""
    B, H, W, C = x.shape
    co = C // (r * r)
    return x.reshape(B, H, W, co, r, r).permute(0, 1, 4, 2, 5, 3).reshape(B, H * r, W * r, co)


def conv1x1_nhwc(x, w, b=None):
    return F.linear(x, w.reshape(w.shape[0], w.shape[1]), b)


def upsample(sd, pre, x, H, W, r):
    """UpSample.forward, the "Dual up-sample" (SUNet_detail.py:365-386), in NHWC.
    Returns (B, 4L, C/2) for r == 2 and the 4-D (B, 4H, 4W, C) tensor for r == 4 (:383-386)."""
    B, L, C = x.shape
    t = x.reshape(B, H, W, C)
    p = conv1x1_nhwc(t, sd[pre + "up_p.0.weight"])
    p = pixel_shuffle_nhwc(prelu(p, sd[pre + "up_p.1.weight"]), r)
    p = conv1x1_nhwc(p, sd[pre + "up_p.3.weight"])
    b = conv1x1_nhwc(t, sd[pre + "up_b.0.weight"], sd[pre + "up_b.0.bias"])
    b = bilinear_up_nhwc(prelu(b, sd[pre + "up_b.1.weight"]), r)
    b = conv1x1_nhwc(b, sd[pre + "up_b.3.weight"])
    out = conv1x1_nhwc(torch.cat([p, b], dim=-1), sd[pre + "conv.weight"])
    if r == 2:
        return out.reshape(B, -1, C // 2)
    return out


def patch_embed(sd, pre, img):
    """PatchEmbed.forward (SUNet_detail.py:548-556): conv k=s=patch, flatten, transpose, LN."""
    w = sd[pre + "proj.weight"]
    x = F.conv2d(img, w, sd[pre + "proj.bias"], stride=w.shape[-1])
    x = x.flatten(2).transpose(1, 2)
    if pre + "norm.weight" in sd:
        x = layer_norm(x, sd[pre + "norm.weight"], sd[pre + "norm.bias"])
    return x


# ----------------------------------------------------------------------------- whole model
DEFAULT_ARCH = dict(img_size=256, patch_size=4, embed_dim=96, depths=(8, 8, 8, 8), num_heads=(8, 8, 8, 8),
                    window_size=8, qk_scale=8)


def arch_from_yaml(opt):
    s = opt["SWINUNET"]
    return dict(img_size=s["IMG_SIZE"], patch_size=s["PATCH_SIZE"], embed_dim=s["EMB_DIM"], depths=tuple(s["DEPTH_EN"]),
                num_heads=tuple(s["HEAD_NUM"]), window_size=s["WIN_SIZE"], qk_scale=s["QK_SCALE"])


def _scale(arch, dim, heads):
    qs = arch.get("qk_scale")
    return qs if qs else (dim // heads) ** -0.5    # `qk_scale or head_dim ** -0.5` (SUNet_detail.py:80)


def sunet_forward(sd, img, arch=None, pre="", taps=None):
    """SUNet.forward (SUNet_detail.py:748-755) = conv_first -> forward_features (:706-720) ->
    forward_up_features (:723-734) -> up_x4 (:736-746) -> output conv (:753).
    ``taps`` (optional dict) receives intermediate activations for per-module goldens."""
    arch = arch or DEFAULT_ARCH
    E, depths, heads = arch["embed_dim"], arch["depths"], arch["num_heads"]
    G = arch["img_size"] // arch["patch_size"]
    ws = arch["window_size"]
    nl = len(depths)

    def tap(name, t):
        if taps is not None:
            taps[name] = t

    x = F.conv2d(img, sd[pre + "conv_first.weight"], sd[pre + "conv_first.bias"], padding=1)      # :749
    x = patch_embed(sd, pre + "patch_embed.", x)                                                  # :708
    tap("patch_embed", x)
    skips = []
    for i in range(nl):                                                                           # :714-716
        dim, H = E * 2 ** i, G // 2 ** i
        skips.append(x)
        for j in range(depths[i]):
            shift = 0 if j % 2 == 0 else ws // 2                                                  # :423
            x = swin_block(sd, f"{pre}layers.{i}.blocks.{j}.", x, H, H, heads[i], shift, _scale(arch, dim, heads[i]), ws)
            tap(f"layers.{i}.blocks.{j}", x)
        if i < nl - 1:
            x = patch_merging(sd, f"{pre}layers.{i}.downsample.", x, H, H)
            tap(f"layers.{i}.downsample", x)
    x = layer_norm(x, sd[pre + "norm.weight"], sd[pre + "norm.bias"])                             # :718
    for inx in range(nl):                                                                         # :724-730
        i = nl - 1 - inx
        dim, H = E * 2 ** i, G // 2 ** i
        if inx == 0:
            x = upsample(sd, f"{pre}layers_up.0.", x, H, H, 2)
            tap("layers_up.0", x)
            continue
        x = torch.cat([x, skips[i]], dim=-1)                                                      # :728
        x = F.linear(x, sd[f"{pre}concat_back_dim.{inx}.weight"], sd[f"{pre}concat_back_dim.{inx}.bias"])
        tap(f"concat_back_dim.{inx}", x)
        for j in range(depths[i]):
            shift = 0 if j % 2 == 0 else ws // 2                                                  # :493
            x = swin_block(sd, f"{pre}layers_up.{inx}.blocks.{j}.", x, H, H, heads[i], shift, _scale(arch, dim, heads[i]), ws)
            tap(f"layers_up.{inx}.blocks.{j}", x)
        if inx < nl - 1:
            x = upsample(sd, f"{pre}layers_up.{inx}.upsample.", x, H, H, 2)
            tap(f"layers_up.{inx}.upsample", x)
    x = layer_norm(x, sd[pre + "norm_up.weight"], sd[pre + "norm_up.bias"])                       # :732
    tap("norm_up", x)
    x = upsample(sd, pre + "up.", x, G, G, 4)                                                     # :742  (B, 4G, 4G, E)
    tap("up", x)
    x = x.permute(0, 3, 1, 2)                                                                     # :744
    return F.conv2d(x, sd[pre + "output.weight"], None, padding=1)                                # :753


def sunet_model_forward(sd, img, arch=None, taps=None):
    """SUNet_model.forward (model/SUNet.py:26-30): grey input is repeated to 3 channels."""
    if img.shape[1] == 1:
        img = img.repeat(1, 3, 1, 1)
    return sunet_forward(sd, img, arch, pre="swin_unet.", taps=taps)


# ----------------------------------------------------------------------------- callers' maths
def torch_psnr(pred, target):
    """utils/image_utils.py:6-10: 20*log10(1/rmse) after clamping both to [0,1]."""
    d = torch.clamp(pred, 0, 1) - torch.clamp(target, 0, 1)
    return 20.0 * torch.log10(1.0 / torch.sqrt((d ** 2).mean()))


def overlapped_square(img, kernel=256, stride=128):
    """demo_any_resolution.py:35-52: centre the image on a zero canvas of side X = ceil(max(h,w)/k)*k and cut
    row-major overlapping tiles.  The reference's `permute(2,0,1,4,3)` after unfold(3).unfold(2) yields
    tile t = (row t // n, col t % n) with origin (row*stride, col*stride) in (y, x) orientation.
    Returns (tiles (n*n, C, k, k), mask (1,1,X,X), X)."""
    b, c, h, w = img.shape
    X = int(math.ceil(max(h, w) / float(kernel)) * kernel)
    canvas = torch.zeros(1, c, X, X, dtype=img.dtype)
    mask = torch.zeros(1, 1, X, X, dtype=img.dtype)
    oy, ox = (X - h) // 2, (X - w) // 2
    canvas[:, :, oy:oy + h, ox:ox + w] = img
    mask[:, :, oy:oy + h, ox:ox + w] = 1.0
    n = (X - kernel) // stride + 1
    tiles = [canvas[0, :, i * stride:i * stride + kernel, j * stride:j * stride + kernel] for i in range(n) for j in range(n)]
    return torch.stack(tiles), mask, X


def fold_tiles(tiles, X, h, w, kernel=256, stride=128):
    """demo_any_resolution.py:125-139: overlap-add the per-tile outputs, divide by the cover count,
    crop the image region back out and clamp to [0,1].  tiles: (n*n, C, k, k) -> (1, C, h, w)."""
    nt, c, _, _ = tiles.shape
    n = (X - kernel) // stride + 1
    acc = torch.zeros(1, c, X, X, dtype=tiles.dtype)
    cnt = torch.zeros(1, 1, X, X, dtype=tiles.dtype)
    for t in range(nt):
        i, j = t // n, t % n
        acc[0, :, i * stride:i * stride + kernel, j * stride:j * stride + kernel] += tiles[t]
        cnt[0, :, i * stride:i * stride + kernel, j * stride:j * stride + kernel] += 1.0
    acc = acc / cnt
    oy, ox = (X - h) // 2, (X - w) // 2
    return torch.clamp(acc[:, :, oy:oy + h, ox:ox + w], 0, 1)


# ----------------------------------------------------------------------------- demo.py image edge (§8f-3)
def to_tensor_u8(img_u8):
    """demo.py:70-71 `TF.to_tensor(PIL RGB)`: uint8 (B, H, W, C) -> float32 (B, C, H, W) / 255."""
    return img_u8.permute(0, 3, 1, 2).to(torch.float32).div(255)


def to_ubyte(restored):
    """demo.py:76-79: clamp(restored, 0, 1) -> NHWC -> skimage.img_as_ubyte, which for float input is
    rint(x * 255) clipped to [0, 255] (skimage.util.dtype._convert: multiply by imax_out, np.rint, np.clip; skimage is not
    installed here - np.rint and torch.round both round half to even)."""
    x = torch.clamp(restored, 0, 1).permute(0, 2, 3, 1)
    return torch.clamp(torch.round(x * 255.0), 0, 255).to(torch.uint8)


def demo_restore_u8(sd, img_u8, arch=None):
    """demo.py:69-79 for a batch of 8-bit images: (B, H, W, 3) uint8 -> (B, H, W, 1) uint8."""
    return to_ubyte(sunet_model_forward(sd, to_tensor_u8(img_u8), arch))


# ----------------------------------------------------------------------------- train.py validation reductions (§8f-4)
def charbonnier_loss(pred, target, weight=None, eps=1e-3):
    """train.py:187-192."""
    diff = pred - target
    l = torch.sqrt(diff * diff + eps * eps)
    if weight is None:
        return l.mean()
    return (l * weight).sum() / weight.sum().clamp(min=1e-8)


def validation_batch(logits, target, weight=None, eps=1e-3):
    """train.py:437-448 for one batch, given the model output: luminance target, prob = sigmoid(logits), mean squared error,
    weighted squared error and weighted Charbonnier loss (weight None = the unit map).  Returns (prob, dict)."""
    if target.shape[1] == 3:
        target = 0.2989 * target[:, 0:1] + 0.5870 * target[:, 1:2] + 0.1140 * target[:, 2:3]
    prob = torch.sigmoid(logits)
    se = (logits - target) ** 2
    w = weight if weight is not None else torch.ones_like(target)
    return prob, {"mse": se.mean().item(), "mse_weighted": (se * w).sum().item() / max(1e-8, w.sum().item()),
                  "charbonnier": charbonnier_loss(logits, target, weight=w, eps=eps).item()}
