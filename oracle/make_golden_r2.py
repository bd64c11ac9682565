"""Round-2 additions to tests/golden (build container only; the fixtures of make_golden.py are left untouched).

TEST INFRASTRUCTURE ONLY.  Usage:  python -m oracle.make_golden_r2
  sunet_model_outlier.npz  whole-model output of the UNMODIFIED reference on the "outlier" weight style (oracle/weights.py:
                           LayerNorm gains x50 on three channels per block, one residual channel at +300), 2 AWGN images
  modules_r2.npz           the reference model's own member modules concat_back_dim[1..3] (nn.Linear over cat([x, skip], -1),
                           SUNet_detail.py:652-654, :728-729), norm (:677, :718) and norm_up (:678, :732) on seeded inputs
"""
import os
import sys

import numpy as np
import torch

from . import sunet_oracle as O
from . import weights as Wt
from .make_golden import GOLDEN_DIR, module_input
from .reference_loader import load_reference


def main():
    torch.set_num_threads(os.cpu_count())
    SUNet_model, D, cfg = load_reference()
    arch = O.arch_from_yaml(cfg)
    torch.manual_seed(0)
    model = SUNet_model(cfg).eval()
    spec = Wt.sunet_spec()
    with torch.no_grad():
        sd = Wt.synth_state_dict(spec, seed=0, style="outlier")
        model.load_state_dict(sd, strict=True)
        noisy, clean = Wt.awgn_input(2, seed=1)
        ref_out = model(noisy)
        taps = {}
        ora_out = O.sunet_model_forward(sd, noisy, arch, taps=taps)
        err = (ref_out - ora_out).abs().max().item()
        smax = max(v.abs().max().item() for k, v in taps.items() if "blocks" in k)
        print(f"[outlier] oracle-vs-reference max-abs {err:.3e}; out range [{ref_out.min():.3f},{ref_out.max():.3f}]; stream max |x| {smax:.1f}")
        assert err < 5e-5 and smax > 250.0
        np.savez_compressed(os.path.join(GOLDEN_DIR, "sunet_model_outlier.npz"), output=ref_out.numpy(),
                            psnr=np.float64(O.torch_psnr(ref_out, Wt.luminance(clean)).item()), stream_max=np.float64(smax),
                            seed_weights=0, seed_input=1, batch=2)

        sd = Wt.synth_state_dict(spec, seed=0, style="stress")
        model.load_state_dict(sd, strict=True)
        net = model.swin_unet
        mods = {}
        for inx, dim, L in ((1, 384, 200), (2, 192, 333), (3, 96, 520)):   # row counts off the 128-row tile grid on purpose
            x = module_input((1, L, dim), seed=600 + inx)
            skip = module_input((1, L, dim), seed=610 + inx)
            y = net.concat_back_dim[inx](torch.cat([x, skip], -1))          # :728-729
            mods[f"concat_back_dim_{inx}"] = y.numpy()
        mods["norm"] = net.norm(module_input((1, 70, 768), seed=620, scale=3.0) + 0.5).numpy()        # :718
        mods["norm_up"] = net.norm_up(module_input((1, 777, 96), seed=621, scale=3.0) - 0.25).numpy()   # :732
        np.savez_compressed(os.path.join(GOLDEN_DIR, "modules_r2.npz"), **mods)
        print("modules_r2.npz:", {k: v.shape for k, v in mods.items()})


if __name__ == "__main__":
    sys.exit(main())
