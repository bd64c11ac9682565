"""ctypes binding of libsunet_b200.so (the C ABI declared in include/sunet_b200.h).

There is no CPU or PyTorch-eager fallback: if the library cannot be loaded, or a tensor is not on a CUDA device,
every entry point raises.
"""
import ctypes
import os

import torch

from . import _build

_lib = None

c_i64_p = ctypes.POINTER(ctypes.c_int64)
c_f64_p = ctypes.POINTER(ctypes.c_double)

class ProfRec(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("ms", ctypes.c_float), ("flops", ctypes.c_double), ("bytes", ctypes.c_double)]


KERNEL_KINDS = ("gemm_tcgen05", "attn_core", "layernorm", "merge_gather_ln", "patch_embed_conv", "upsample_combine", "tail_stencil",
                "cast", "im2col", "mlp_fused", "attn_fused", "tail_up_fused")

_SIGNATURES = {
    "sunet_abi_version": (ctypes.c_int, []),
    "sunet_last_error": (ctypes.c_char_p, []),
    "sunet_prepack": (ctypes.c_int, [ctypes.c_char_p, c_i64_p, ctypes.c_int, c_f64_p, ctypes.c_int, ctypes.POINTER(ctypes.c_char_p),
                                     ctypes.POINTER(ctypes.c_void_p), c_i64_p, ctypes.c_int, ctypes.c_void_p,
                                     ctypes.POINTER(ctypes.c_void_p)]),
    "sunet_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "sunet_swin_block_fwd": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "sunet_swin_block_f16_workspace_bytes": (ctypes.c_size_t, [ctypes.c_void_p, ctypes.c_int]),
    "sunet_swin_block_f16": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_size_t, ctypes.c_void_p]),
    "sunet_window_attention_fwd": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int,
                                                  ctypes.c_void_p, ctypes.c_void_p]),
    "sunet_mlp_fwd": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]),
    "sunet_patch_merging_fwd": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "sunet_upsample_fwd": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "sunet_patch_embed_fwd": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                             ctypes.c_void_p]),
    "sunet_workspace_bytes": (ctypes.c_size_t, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "sunet_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "sunet_forward_u8": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "sunet_forward_eval": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_int, ctypes.c_void_p, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "sunet_forward_profile": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.POINTER(ProfRec), ctypes.c_int,
                                             ctypes.POINTER(ctypes.c_int)]),
    "sunet_forward_launches": (ctypes.c_int64, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "sunet_tiles_extract": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "sunet_tiles_fold": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "sunet_tiles_finish": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_void_p, ctypes.c_void_p]),
    "sunet_selftest_umma": (ctypes.c_int, [ctypes.c_void_p]),
    "sunet_ln_mlp_residual_f16": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int] + [ctypes.c_void_p] * 8),
    "sunet_layernorm_f16": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p]),
    "sunet_concat_linear_f16": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_void_p]),
    "sunet_gemm_f16": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib_path():
    return _build.LIB_PATH


def load(build_if_missing=True):
    """Load (building first if the .so is missing or stale and nvcc is available) and type the library."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    override = os.environ.get("SUNET_LIB_PATH")   # A/B experiments (tools/ab_variants.py): another build of the SAME library
    if override:
        path, build_if_missing = override, False
    if build_if_missing:
        try:
            path = _build.build()
        except RuntimeError:
            if not os.path.exists(path):
                raise
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with `python -m sunet_tf_b200._build` (needs nvcc); there is no fallback path")
    lib = ctypes.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().sunet_last_error().decode(errors="replace")
        raise RuntimeError(f"libsunet_b200 error {rc}: {msg}")


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, name="input"):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"sunet_tf_b200: {name} must be a CUDA tensor (this package has no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"sunet_tf_b200: {name} must be float32, got {t.dtype}")
    return t.contiguous()


def prepack(kind, iargs, fargs, named_tensors, device):
    """named_tensors: list of (key, fp32 CUDA tensor).  Returns an opaque handle (int)."""
    lib = load()
    n = len(named_tensors)
    keep = [require_cuda(t.detach().float() if t.dtype != torch.float32 else t.detach(), k) for k, t in named_tensors]
    names = (ctypes.c_char_p * n)(*[k.encode() for k, _ in named_tensors])
    ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in keep])
    numels = (ctypes.c_int64 * n)(*[t.numel() for t in keep])
    ia = (ctypes.c_int64 * max(1, len(iargs)))(*[int(v) for v in iargs])
    fa = (ctypes.c_double * max(1, len(fargs)))(*[float(v) for v in fargs])
    out = ctypes.c_void_p()
    with torch.cuda.device(device):
        check(lib.sunet_prepack(kind.encode(), ia, len(iargs), fa, len(fargs), names, ptrs, numels, n, stream_ptr(device), ctypes.byref(out)))
    del keep
    return out.value


def destroy(handle):
    if handle and _lib is not None:
        _lib.sunet_destroy(ctypes.c_void_p(handle))
