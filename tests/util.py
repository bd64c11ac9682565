import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def module_input(shape, seed, scale=1.0):
    """Same generator as oracle/make_golden.py::module_input."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def rel_err(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-6)).item()


def max_abs(got, ref):
    return (got.detach().float().cpu() - ref.detach().float().cpu()).abs().max().item()
