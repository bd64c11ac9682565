"""Golden fixtures for the callers either side of the forward, generated with the UNMODIFIED
reference model and the reference scripts' own source lines.  TEST INFRASTRUCTURE - run here (needs /root/reference):

    python -m oracle.make_golden_edges

* edge_demo_u8.npz       demo.py:69-79 - PIL RGB -> TF.to_tensor -> model -> clamp -> NHWC -> img_as_ubyte.  PIL and torchvision
                         are present and used as the reference does; skimage is absent, so img_as_ubyte is the restatement in
                         oracle/sunet_oracle.py::to_ubyte (rint(x*255)).  Weights: the seeded "stress" set with output.weight x6 so
                         that the result covers both clamp sides.
* edge_validation.npz    train.py:437-448 - luminance target, sigmoid, squared error, weighted squared error and Charbonnier loss,
                         with `charbonnier_loss` / `mse_loss` exec'ed from train.py:187-197 (train.py itself is a script: argparse,
                         CUDA and dataset paths at import).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import sunet_oracle as O  # noqa: E402
from oracle import weights as Wt  # noqa: E402
from oracle.reference_loader import load_reference, reference_root  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
U8_OUTPUT_GAIN = 6.0


def u8_images(batch, seed, size=256):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (batch, size, size, 3), generator=g, dtype=torch.uint8)


def u8_state_dict(seed=21):
    sd = Wt.synth_state_dict(Wt.sunet_spec(), seed=seed, style="stress")
    sd["swin_unet.output.weight"] = sd["swin_unet.output.weight"] * U8_OUTPUT_GAIN
    return sd


def validation_case(batch, seed, size=256):
    """A smooth-ish binary-like target (3-channel, as the loader yields), a noisy input and a ring-like weight map."""
    g = torch.Generator().manual_seed(seed)
    target = (torch.rand(batch, 3, size, size, generator=g) > 0.8).float() * 0.9
    inp = torch.clamp(target + torch.randn(batch, 3, size, size, generator=g) * 0.2, 0, 1)
    weight = torch.rand(batch, 1, size, size, generator=g) * 3.0
    weight = weight / weight.mean()
    return target, inp, weight


def main():
    import torchvision.transforms.functional as TF
    from PIL import Image
    SUNet_model, D, cfg = load_reference()
    root = reference_root()
    model = SUNet_model(cfg).eval()

    # ---- demo.py edge
    sd = u8_state_dict()
    model.load_state_dict(sd, strict=True)
    imgs = u8_images(2, seed=22)
    outs = []
    with torch.no_grad():
        for i in range(imgs.shape[0]):
            img = Image.fromarray(imgs[i].numpy(), mode="RGB")          # demo.py:70
            input_ = TF.to_tensor(img).unsqueeze(0)                     # :71
            restored = model(input_)                                    # :75
            restored = torch.clamp(restored, 0, 1)                      # :76
            restored = restored.permute(0, 2, 3, 1).cpu().detach().numpy()  # :77
            outs.append(np.clip(np.rint(restored[0] * 255.0), 0, 255).astype(np.uint8))   # :79 img_as_ubyte
    ref_u8 = np.stack(outs)
    mine = O.demo_restore_u8(sd, imgs).numpy()
    # the oracle forward is a restatement with a different op order (2e-5 of the reference, tests/test_oracle.py): levels may
    # differ by one where x*255 sits on a rounding boundary, never by more
    dlev = np.abs(ref_u8.astype(np.int16) - mine.astype(np.int16))
    assert dlev.max() <= 1 and (dlev != 0).mean() < 1e-3, "oracle demo edge differs from the reference lines"
    np.savez_compressed(os.path.join(GOLDEN_DIR, "edge_demo_u8.npz"), output=ref_u8, seed_weights=21, seed_input=22,
                        output_gain=U8_OUTPUT_GAIN)
    print("demo u8:", ref_u8.shape, "levels", int(ref_u8.min()), "..", int(ref_u8.max()), "zeros", float((ref_u8 == 0).mean()),
          "255s", float((ref_u8 == 255).mean()))

    # ---- validation reductions
    src = open(os.path.join(root, "train.py")).read()
    start, end = src.index("def charbonnier_loss"), src.index("def background_adjacent_to_foreground")
    ns = {"torch": torch}
    exec(src[start:end], ns)
    sd = Wt.synth_state_dict(Wt.sunet_spec(), seed=23, style="stress")
    model.load_state_dict(sd, strict=True)
    target, inp, weight = validation_case(2, seed=24)
    with torch.no_grad():
        tgt = 0.2989 * target[:, 0:1] + 0.5870 * target[:, 1:2] + 0.1140 * target[:, 2:3]     # train.py:437-438
        logits = model(inp)                                                                   # :440
        prob = torch.sigmoid(logits)                                                          # :441
        se = (logits - tgt) ** 2                                                              # :444
        mse = se.mean().item()                                                                # :445
        mse_w = (se * weight).sum().item() / max(1e-8, weight.sum().item())                   # :448
        charb = ns["charbonnier_loss"](logits, tgt, weight=weight, eps=1e-3).item()           # :450
        mse_w2 = ns["mse_loss"](logits, tgt, weight=weight).item()
        charb_unit = ns["charbonnier_loss"](logits, tgt, weight=None, eps=1e-3).item()
    my_prob, m = O.validation_batch(logits, target, weight)
    assert torch.equal(my_prob, prob) and abs(m["mse"] - mse) < 1e-9 and abs(m["mse_weighted"] - mse_w) < 1e-9
    assert abs(m["charbonnier"] - charb) < 1e-9 and abs(mse_w2 - mse_w) < 1e-7
    np.savez_compressed(os.path.join(GOLDEN_DIR, "edge_validation.npz"), logits=logits.numpy().astype(np.float16),
                        prob=prob.numpy().astype(np.float16), mse=mse, mse_weighted=mse_w, charbonnier=charb,
                        charbonnier_unit=charb_unit, seed_weights=23, seed_input=24)
    print("validation: mse", mse, "weighted", mse_w, "charbonnier", charb, "unit-weight charbonnier", charb_unit)


if __name__ == "__main__":
    main()
