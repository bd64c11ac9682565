"""Checkpoint format of the reference and its loaders (utils/model_utils.py:20-54, demo.py:33-43,
demo_any_resolution.py:83-93).

A checkpoint is ``torch.save({'epoch': int, 'state_dict': model.state_dict(), 'optimizer': optimizer.state_dict()})``
(train.py:520-535).  Models trained under ``nn.DataParallel`` carry a ``module.`` prefix on every key, which the reference
strips by slicing 7 characters off EVERY key when the strict load fails (model_utils.py:27-37).  Here the prefix is
removed only from keys that have it (a mixed dict still fails loudly in ``load_state_dict``), the load stays strict
(867 keys, identical shapes), and the device-side pre-pack (fp16 / folded weights) is rebuilt by the first forward after
the load - `_Packed._handle` keys the pack on every parameter's (data_ptr, version).
"""
import os
from collections import OrderedDict

import torch

_PREFIX = "module."


def strip_module_prefix(state_dict):
    """``module.swin_unet...`` -> ``swin_unet...`` (model_utils.py:32-36)."""
    out = OrderedDict()
    for k, v in state_dict.items():
        out[k[len(_PREFIX):] if k.startswith(_PREFIX) else k] = v
    return out


def _read(weights, map_location, trusted=False):
    """``weights_only=True`` by default: the reference's {'epoch', 'state_dict', 'optimizer'} checkpoints hold only tensors and
    plain containers, so nothing needs the full (code-executing) unpickler.  ``trusted=True`` opts in to it for checkpoints
    from a source you control that pickled other objects."""
    ckpt = torch.load(weights, map_location=map_location, weights_only=not trusted) if isinstance(weights, (str, os.PathLike)) else weights
    if not isinstance(ckpt, dict) or "state_dict" not in ckpt:
        raise RuntimeError("checkpoint has no 'state_dict' entry (expected {'epoch', 'state_dict', 'optimizer'}, train.py:520-535)")
    return ckpt


def load_checkpoint(model, weights, map_location="cpu", trusted=False):
    """model_utils.py:27-37 / demo.py:33-43: strict load, retrying with the ``module.`` prefix stripped.  ``weights`` is a path
    or an already loaded checkpoint dict.  Returns the checkpoint dict (the reference returns None)."""
    ckpt = _read(weights, map_location, trusted)
    state = ckpt["state_dict"]
    try:
        model.load_state_dict(state)
    except RuntimeError:
        model.load_state_dict(strip_module_prefix(state))
    return ckpt


def load_checkpoint_multigpu(model, weights, map_location="cpu", trusted=False):
    """model_utils.py:40-47: the prefix is always stripped."""
    ckpt = _read(weights, map_location, trusted)
    model.load_state_dict(strip_module_prefix(ckpt["state_dict"]))
    return ckpt


def load_start_epoch(weights, map_location="cpu", trusted=False):
    """model_utils.py:50-53."""
    return _read(weights, map_location, trusted)["epoch"]


def save_checkpoint(model_dir, state, session):
    """model_utils.py:20-24: ``model_epoch_{epoch}_{session}.pth`` holding {'epoch', 'state_dict', 'optimizer'}."""
    path = os.path.join(model_dir, "model_epoch_{}_{}.pth".format(state["epoch"], session))
    torch.save(state, path)
    return path
