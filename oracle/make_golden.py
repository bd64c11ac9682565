"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE ONLY.  Usage:  python -m oracle.make_golden
Weights and inputs are regenerated from seeds (oracle/weights.py); only OUTPUTS of the reference are
stored, as fp32, together with the seeds/shapes needed to rebuild the inputs.
"""
import os
import sys
import time

import numpy as np
import torch

from . import sunet_oracle as O
from . import weights as Wt
from .reference_loader import load_reference

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def module_input(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def check_spec_against_reference(model):
    ref_sd = model.state_dict()
    spec = Wt.sunet_spec()
    ref_keys = list(ref_sd.keys())
    my_keys = [k for k, *_ in spec]
    assert ref_keys == my_keys, f"key mismatch: {set(ref_keys) ^ set(my_keys)} or order differs"
    for k, shape, kind, extra in spec:
        assert tuple(ref_sd[k].shape) == tuple(shape), (k, ref_sd[k].shape, shape)
        if kind in ("index", "mask"):
            mine = Wt.make_tensor(k, shape, kind, extra, 0, "init")
            assert torch.equal(ref_sd[k].float(), mine.float()), f"closed form differs from reference buffer {k}"
    print(f"spec OK: {len(spec)} state_dict entries, closed-form index/mask buffers identical to the reference")
    return spec


def main():
    torch.set_num_threads(os.cpu_count())
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    SUNet_model, D, cfg = load_reference()
    arch = O.arch_from_yaml(cfg)
    torch.manual_seed(0)
    model = SUNet_model(cfg).eval()
    spec = check_spec_against_reference(model)

    # ---------------- whole model, both weight styles
    with torch.no_grad():
        for style in ("init", "stress"):
            sd = Wt.synth_state_dict(spec, seed=0, style=style)
            model.load_state_dict(sd, strict=True)
            noisy, clean = Wt.awgn_input(2, seed=1)
            t0 = time.time()
            ref_out = model(noisy)
            taps = {}
            ora_out = O.sunet_model_forward(sd, noisy, arch, taps=taps)
            err = (ref_out - ora_out).abs().max().item()
            print(f"[{style}] reference fwd B=2 {time.time() - t0:.1f}s; oracle-vs-reference max-abs {err:.3e}; "
                  f"out range [{ref_out.min():.3f},{ref_out.max():.3f}]")
            assert err < 5e-5, "oracle restatement disagrees with the reference"
            psnr = O.torch_psnr(ref_out, Wt.luminance(clean)).item()
            tap_stats = {k: np.array([v.mean().item(), v.std().item(), v.abs().max().item()], dtype=np.float64)
                         for k, v in taps.items()}
            np.savez_compressed(os.path.join(GOLDEN_DIR, f"sunet_model_{style}.npz"), output=ref_out.numpy(), psnr=np.float64(psnr),
                                seed_weights=0, seed_input=1, batch=2,
                                tap_names=np.array(list(tap_stats.keys())), tap_stats=np.stack(list(tap_stats.values())))
            # grey input path (SUNet.py:27-28)
            if style == "init":
                grey = noisy[:1, :1]
                np.savez_compressed(os.path.join(GOLDEN_DIR, "sunet_model_grey.npz"), output=model(grey).numpy())

        # ---------------- per-module goldens on small grids (reference classes instantiated stand-alone)
        mods = {}
        for dim in (96, 192, 384, 768):
            for shift in (0, 4):
                blk = D.SwinTransformerBlock(dim=dim, input_resolution=(16, 16), num_heads=8, window_size=8, shift_size=shift,
                                             mlp_ratio=4.0, qkv_bias=True, qk_scale=8).eval()
                bspec = Wt.block_spec("", dim, 16, 16, shift)
                assert [k for k, *_ in bspec] == list(blk.state_dict().keys())
                sd = Wt.synth_state_dict(bspec, seed=dim + shift, style="stress")
                blk.load_state_dict(sd, strict=True)
                x = module_input((1, 256, dim), seed=100 + dim + shift)
                y = blk(x)
                yo = O.swin_block(sd, "", x, 16, 16, 8, shift, 8)
                assert (y - yo).abs().max() < 1e-4, (dim, shift, (y - yo).abs().max())
                mods[f"block_{dim}_{shift}"] = y.numpy()
                # stand-alone WindowAttention on the same weights (mask = the block's buffer or None)
                xw = module_input((4, 64, dim), seed=200 + dim + shift)
                mods[f"wattn_{dim}_{shift}"] = blk.attn(xw, mask=blk.attn_mask).numpy()
                if shift == 0 and dim in (96, 768):
                    mods[f"mlp_{dim}"] = blk.mlp(xw[:2]).numpy()
        pm = D.PatchMerging((16, 16), 96).eval()
        sd = Wt.synth_state_dict(Wt.merging_spec("", 96), seed=7, style="stress")
        pm.load_state_dict(sd, strict=True)
        mods["merging_96"] = pm(module_input((2, 256, 96), seed=300)).numpy()
        for C, r, H in ((192, 2, 8), (768, 2, 8), (96, 4, 16)):
            up = D.UpSample((H, H), C, r).eval()
            sd = Wt.synth_state_dict(Wt.upsample_spec("", C, r), seed=11 + C + r, style="stress")
            up.load_state_dict(sd, strict=True)
            x = module_input((2, H * H, C), seed=400 + C + r)
            y = up(x)
            assert (y - O.upsample(sd, "", x, H, H, r)).abs().max() < 1e-4
            mods[f"upsample_{C}_{r}"] = y.numpy()
        pe = D.PatchEmbed(img_size=64, patch_size=4, in_chans=96, embed_dim=96, norm_layer=torch.nn.LayerNorm).eval()
        sd = Wt.synth_state_dict(Wt.patch_embed_spec("", 96, 96), seed=13, style="stress")
        pe.load_state_dict(sd, strict=True)
        mods["patch_embed_96"] = pe(module_input((2, 96, 64, 64), seed=500)).numpy()
        np.savez_compressed(os.path.join(GOLDEN_DIR, "modules.npz"), **mods)
        print("modules.npz:", {k: v.shape for k, v in mods.items()})

        # ---------------- any-resolution tile pipeline (demo_any_resolution.py:35-52,116-139) with a 3-channel head
        # (the fork's out_chans=1 wrapper cannot run this script for >1 tile, SURVEY.md 3.2, so SUNet is built with out_chans=3)
        net3 = D.SUNet(img_size=256, patch_size=4, in_chans=3, out_chans=3, embed_dim=96, depths=[8] * 4, num_heads=[8] * 4,
                       window_size=8, mlp_ratio=4.0, qkv_bias=True, qk_scale=8, drop_rate=0.0, drop_path_rate=0.1, ape=False,
                       patch_norm=True).eval()
        spec3 = Wt.sunet_spec(pre="", out_chans=3)
        assert [k for k, *_ in spec3] == list(net3.state_dict().keys())
        sd3 = Wt.synth_state_dict(spec3, seed=3, style="init")
        net3.load_state_dict(sd3, strict=True)
        h, w = 300, 420
        g = torch.Generator().manual_seed(4)
        clean = torch.rand(1, 3, h, w, generator=g)
        noisy = torch.round(torch.clamp(clean + torch.randn(1, 3, h, w, generator=g) * (50 / 255.0), 0, 1) * 255) / 255

        # the reference's own tiling code, executed from its source lines (the script is not importable: argparse + .cuda())
        import math
        import torch.nn.functional as F
        src = open(os.path.join(os.environ.get("SUNET_REF", "/root/reference"), "demo_any_resolution.py")).read()
        start = src.index("def overlapped_square")
        end = src.index("# Utility: save RGB image")
        ns = {"torch": torch, "math": math}
        exec(src[start:end], ns)
        patches, mask, X = ns["overlapped_square"](noisy, kernel=256, stride=128)
        my_tiles, my_mask, my_X = O.overlapped_square(noisy, 256, 128)
        assert my_X == X and torch.equal(mask, my_mask) and all(torch.equal(p[0], my_tiles[i]) for i, p in enumerate(patches))
        outs = torch.cat([net3(p) for p in patches], dim=0)
        B, C, H, W = outs.shape
        patch = outs.contiguous().view(B, C, -1, 256 * 256).permute(2, 1, 3, 0).contiguous().view(1, C * 256 * 256, -1)
        wm = torch.ones_like(outs).view(B, C, -1, 256 * 256).permute(2, 1, 3, 0).contiguous().view(1, C * 256 * 256, -1)
        restored = F.fold(patch, output_size=(X, X), kernel_size=256, stride=128) / F.fold(wm, output_size=(X, X), kernel_size=256, stride=128)
        restored = torch.clamp(torch.masked_select(restored, mask.bool()).reshape(noisy.shape), 0, 1)
        mine = O.fold_tiles(outs, X, h, w, 256, 128)
        assert (restored - mine).abs().max() < 1e-6
        np.savez_compressed(os.path.join(GOLDEN_DIR, "tiles_300x420.npz"), restored=restored.numpy(), n_tiles=B, X=X, h=h, w=w,
                            seed_weights=3, seed_input=4)
        print(f"tiles: {B} tiles, X={X}, folded output {tuple(restored.shape)}")
    total = sum(os.path.getsize(os.path.join(GOLDEN_DIR, f)) for f in os.listdir(GOLDEN_DIR))
    print(f"golden dir size: {total / 1e6:.1f} MB")


if __name__ == "__main__":
    sys.exit(main())
