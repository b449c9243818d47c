"""Shared helpers for the parity tests."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    cfg = json.loads(str(d["cfg"]))
    arrs = {k: torch.from_numpy(np.asarray(d[k])) for k in d.files if k != "cfg"}
    return cfg, arrs


def golden_names(prefix):
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith(prefix) and f.endswith(".npz"))


def assert_close(got, want, rtol, atol, what=""):
    got = got.detach().float().cpu()
    want = want.detach().float().cpu()
    assert got.shape == want.shape, f"{what}: shape {tuple(got.shape)} vs {tuple(want.shape)}"
    err = (got - want).abs()
    tol = atol + rtol * want.abs()
    bad = err > tol
    if bad.any():
        i = int(torch.argmax(err - tol))
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.numel()} elements out of tolerance "
                             f"(rtol={rtol}, atol={atol}); worst |err|={float(err.flatten()[i]):.3e} at ref="
                             f"{float(want.flatten()[i]):.4e}; max|ref|={float(want.abs().max()):.3e}")


def rel_l2(got, want):
    got = got.detach().double().cpu()
    want = want.detach().double().cpu()
    return float((got - want).norm() / want.norm().clamp_min(1e-30))
