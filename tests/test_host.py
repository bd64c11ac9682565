"""CPU: host logic - module surface / state_dict parity, C-ABI library exports, sharding, 2-process gloo tile reduce."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from oracle import weights as Wt
from sunet_tf_b200 import SUNet_model, _lib, shard
from sunet_tf_b200.default_config import DEFAULT_OPT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_state_dict_surface_matches_reference_spec():
    m = SUNet_model(DEFAULT_OPT)
    sd = m.state_dict()
    spec = Wt.sunet_spec()
    assert list(sd.keys()) == [k for k, *_ in spec]          # 867 entries, reference registration order
    assert len(sd) == 867
    for k, shape, kind, extra in spec:
        assert tuple(sd[k].shape) == tuple(shape), k
        if kind == "index":
            assert sd[k].dtype == torch.int64
        elif kind in ("mask",):
            assert torch.equal(sd[k], Wt.make_tensor(k, shape, kind, extra, 0, "init"))
    assert sum(p.numel() for p in m.parameters()) == 99681993
    # a synthetic "reference checkpoint" loads strictly, also with the DataParallel `module.` prefix stripped by the callers
    m.load_state_dict(Wt.synth_state_dict(spec, seed=0, style="init"), strict=True)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "sunet_b200.h")).read()
    declared = set(re.findall(r"\b(sunet_[a-z0-9_]+)\s*\(", header))
    declared -= {"sunet_handle_s"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/sunet_b200.h but not exported"
    assert set(_lib.EXPORTED_SYMBOLS) == declared
    assert lib.sunet_abi_version() == 1
    assert isinstance(lib.sunet_last_error(), bytes)


def test_library_is_built_for_sm100a_with_tcgen05_and_tma():
    out = subprocess.run(["cuobjdump", "-sass", _lib.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "UTCHMMA" in out and "UTMALDG" in out and "LDTM" in out


def test_cpu_tensor_is_rejected_not_silently_computed():
    m = SUNet_model(DEFAULT_OPT)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 256, 256))


def test_shard_ranges():
    for n, g in ((225, 8), (225, 4), (9, 2), (3, 8), (512, 8), (7, 3)):
        rs = [shard.tile_range(n, r, g) for r in range(g)]
        assert rs[0][0] == 0 and rs[-1][1] == n
        for r, (lo, hi) in enumerate(rs):
            assert all(t * g // n == r for t in range(lo, hi))
        bs = [shard.batch_range(n, r, g) for r in range(g)]
        assert bs[0][0] == 0 and bs[-1][1] == n and all(bs[i][1] == bs[i + 1][0] for i in range(g - 1))
        assert max(hi - lo for lo, hi in bs) - min(hi - lo for lo, hi in bs) <= 1
    assert shard.batch_range(512, 3, 8) == (192, 256)


WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["SUNET_ROOT"])
from oracle import sunet_oracle as O
from sunet_tf_b200 import shard
rank, ws, _ = shard.init_process_group("gloo")
h, w = 300, 420
img = torch.rand(1, 3, h, w, generator=torch.Generator().manual_seed(3))
tiles, mask, X = O.overlapped_square(img)
n = tiles.shape[0]
lo, hi = shard.tile_range(n, rank, ws)
# each rank "denoises" (identity * 0.5) only its tiles and folds them into a canvas; one SUM reduce joins the canvases
acc = torch.zeros(3, X, X)
k, s = 256, 128
per = (X - k) // s + 1
for t in range(lo, hi):
    i, j = t // per, t % per
    acc[:, i*s:i*s+k, j*s:j*s+k] += 0.5 * tiles[t]
dist.reduce(acc, dst=0, op=dist.ReduceOp.SUM)
t_max = shard.max_over_ranks(float(rank + 1))
if rank == 0:
    full = O.fold_tiles(0.5 * tiles, X, h, w)
    cnt = torch.zeros(1, X, X)
    for t in range(n):
        i, j = t // per, t % per
        cnt[:, i*s:i*s+k, j*s:j*s+k] += 1
    oy, ox = (X - h) // 2, (X - w) // 2
    mine = torch.clamp((acc / cnt)[:, oy:oy+h, ox:ox+w], 0, 1)[None]
    assert (mine - full).abs().max().item() < 1e-6, (mine - full).abs().max().item()
    assert t_max == float(ws)
    print("GLOO_OK", lo, hi, n)
dist.destroy_process_group()
'''


def test_two_rank_gloo_tile_shard_and_reduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, SUNET_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GLOO_OK" in r.stdout


# ---------------------------------------------------------------- checkpoint format
def test_checkpoint_round_trip_and_module_prefix(tmp_path):
    from collections import OrderedDict

    from sunet_tf_b200 import checkpoint as ck
    spec = Wt.sunet_spec()
    sd = Wt.synth_state_dict(spec, seed=3, style="init")
    # the training script's format (train.py:520-535): {'epoch', 'state_dict', 'optimizer'}
    path = ck.save_checkpoint(str(tmp_path), {"epoch": 7, "state_dict": sd, "optimizer": {"state": {}, "param_groups": []}}, "bestPSNR")
    assert os.path.basename(path) == "model_epoch_7_bestPSNR.pth"
    m = SUNet_model(DEFAULT_OPT)
    ckpt = ck.load_checkpoint(m, path)
    assert ckpt["epoch"] == 7 and ck.load_start_epoch(path) == 7
    assert all(torch.equal(v, sd[k]) for k, v in m.state_dict().items())
    # nn.DataParallel checkpoints: every key carries `module.` (model_utils.py:32-36)
    dp = OrderedDict(("module." + k, v * 2 if v.dtype == torch.float32 else v) for k, v in sd.items())
    m2 = SUNet_model(DEFAULT_OPT)
    ck.load_checkpoint(m2, {"epoch": 1, "state_dict": dp})
    assert torch.equal(m2.state_dict()["swin_unet.norm.weight"], sd["swin_unet.norm.weight"] * 2)
    m3 = SUNet_model(DEFAULT_OPT)
    ck.load_checkpoint_multigpu(m3, {"epoch": 1, "state_dict": dp})
    assert torch.equal(m3.state_dict()["swin_unet.output.weight"], sd["swin_unet.output.weight"] * 2)
    # strictness is kept: a missing key or a foreign dict fails loudly
    broken = OrderedDict(sd)
    broken.pop("swin_unet.norm.weight")
    with pytest.raises(RuntimeError):
        ck.load_checkpoint(SUNet_model(DEFAULT_OPT), {"state_dict": broken})
    with pytest.raises(RuntimeError):
        ck.load_checkpoint(SUNet_model(DEFAULT_OPT), {"weights": sd})


def test_validation_accumulator_matches_reference_aggregation():
    from sunet_tf_b200.validation import ValidationAccumulator, metrics_from_sums
    # two batches; sums = [sum se, sum se*w, sum w, sum charb*w, count]
    s1 = torch.tensor([4.0, 6.0, 3.0, 9.0, 8.0], dtype=torch.float64)
    s2 = torch.tensor([2.0, 0.0, 0.0, 0.0, 4.0], dtype=torch.float64)      # all-zero weights: clamp(min=1e-8) (train.py:192)
    assert metrics_from_sums(s1) == {"mse": 0.5, "mse_weighted": 2.0, "charbonnier": 3.0}
    acc = ValidationAccumulator()
    acc.update(s1)
    acc.update(s2)
    r = acc.result()
    assert r["batches"] == 2 and r["val_mse"] == 0.5 and r["val_mse_weighted"] == 1.0 and r["val_loss"] == 1.5


def test_u8_and_eval_entry_points_reject_cpu_tensors():
    m = SUNet_model(DEFAULT_OPT)
    with pytest.raises(RuntimeError):
        m.forward_u8(torch.zeros(1, 256, 256, 3, dtype=torch.uint8))
    with pytest.raises(RuntimeError):
        m.forward_eval(torch.zeros(1, 3, 256, 256), torch.zeros(1, 3, 256, 256))


# ---------------------------------------------------------------- the device pack handle is never shared between objects
def test_pack_handle_is_not_copied_by_deepcopy_pickle_or_dataparallel(monkeypatch):
    """copy.deepcopy / pickle / nn.DataParallel replicas clone a module's __dict__; each clone must start with an empty handle
    slot, and dropping a clone must not destroy the source's handle (use-after-free / double free otherwise)."""
    import copy
    import pickle

    from sunet_tf_b200 import _lib, modules
    destroyed = []
    monkeypatch.setattr(_lib, "destroy", lambda h: destroyed.append(h))
    blk = modules.SwinTransformerBlock(96, (8, 8), 8, window_size=8, shift_size=0, qk_scale=8)
    blk.__dict__["_sunet_handle"] = 0xDEAD0   # as if packed
    blk.__dict__["_sunet_key"] = ("k",)
    for clone in (copy.deepcopy(blk), copy.copy(blk), pickle.loads(pickle.dumps(blk)), blk._replicate_for_data_parallel()):
        assert clone.__dict__["_sunet_handle"] is None and clone.__dict__["_sunet_key"] is None
        assert clone.attn.__dict__["_sunet_handle"] is None
        clone._release()
        del clone
    assert destroyed == [], destroyed
    assert blk.__dict__["_sunet_handle"] == 0xDEAD0
    m = SUNet_model(DEFAULT_OPT)
    m.swin_unet.__dict__["_sunet_handle"] = 0xBEEF0
    m.swin_unet.__dict__["_workspaces"][("cuda:0", 1)] = object()
    m2 = copy.deepcopy(m)
    assert m2.swin_unet.__dict__["_sunet_handle"] is None and m2.swin_unet.__dict__["_workspaces"] == {}
    assert set(m2.state_dict()) == set(m.state_dict())
    del m2
    assert destroyed == []
    m.swin_unet._release()
    assert destroyed == [0xBEEF0] and m.swin_unet.__dict__["_sunet_handle"] is None
    blk._release()
    assert destroyed == [0xBEEF0, 0xDEAD0]
