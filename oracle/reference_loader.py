"""Import the UNMODIFIED reference model (mehrdad78/SUNet_TF) for golden generation.

TEST INFRASTRUCTURE ONLY.  Works where the reference tree exists: $SUNET_REF, /root/reference (the build
container) or baseline/_ref (the verbatim, git-ignored install made by baseline/install_ref.py, which gpurun ships
to the GPU box).  The reference imports two packages
that are absent from this image at module import time (model/SUNet_detail.py:5-6):
  * timm.models.layers.{DropPath, to_2tuple, trunc_normal_}
  * thop.profile (used only under __main__, SUNet_detail.py:786)
so tiny in-memory stand-ins are registered in sys.modules before the import; no reference file is
touched, copied or patched.
"""
import collections.abc
import os
import sys
import types

import torch
import yaml

# $SUNET_REF, the build container's read-only tree, or the verbatim install made by baseline/install_ref.py (git-ignored,
# shipped to the GPU box): the same three unmodified files in every case
REF_CANDIDATES = [os.environ.get("SUNET_REF", ""), "/root/reference",
                  os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")]


def reference_root():
    for cand in REF_CANDIDATES:
        if cand and os.path.isfile(os.path.join(cand, "model", "SUNet_detail.py")):
            return cand
    return None


def _install_stubs():
    if "timm.models.layers" in sys.modules and "thop" in sys.modules:
        return

    def to_2tuple(v):
        if isinstance(v, collections.abc.Iterable) and not isinstance(v, str):
            return tuple(v)
        return (v, v)

    class DropPath(torch.nn.Module):
        """Stochastic depth; identity in eval mode, which is the only mode the goldens use."""

        def __init__(self, drop_prob=0.0, scale_by_keep=True):
            super().__init__()
            self.drop_prob = drop_prob
            self.scale_by_keep = scale_by_keep

        def forward(self, x):
            if not self.training or self.drop_prob == 0.0:
                return x
            keep = 1.0 - self.drop_prob
            shape = (x.shape[0],) + (1,) * (x.ndim - 1)
            gate = x.new_empty(shape).bernoulli_(keep)
            if keep > 0.0 and self.scale_by_keep:
                gate.div_(keep)
            return x * gate

    layers = types.ModuleType("timm.models.layers")
    layers.to_2tuple = to_2tuple
    layers.DropPath = DropPath
    layers.trunc_normal_ = torch.nn.init.trunc_normal_
    models = types.ModuleType("timm.models")
    models.layers = layers
    timm = types.ModuleType("timm")
    timm.models = models
    thop = types.ModuleType("thop")
    thop.profile = lambda *a, **k: (0, 0)
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers, "thop": thop})


def load_reference():
    """Returns (SUNet_model class, SUNet_detail module, yaml config dict) from the live reference tree."""
    root = reference_root()
    if root is None:
        raise FileNotFoundError("reference tree not found (set SUNET_REF); goldens can only be made in the build container")
    _install_stubs()
    if root not in sys.path:
        sys.path.insert(0, root)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from model import SUNet as ref_wrapper  # model/SUNet.py
        from model import SUNet_detail as ref_detail  # model/SUNet_detail.py
    with open(os.path.join(root, "training.yaml")) as fh:
        cfg = yaml.safe_load(fh)
    return ref_wrapper.SUNet_model, ref_detail, cfg
