"""Deterministic synthetic parameters / inputs (no checkpoints exist offline).

The reference zero-initialises every ResBlock ``out_layers[-1]`` conv, every attention
``proj_out`` and the ``out`` / ``shift_out`` convs (model/module.py:48-54,264-266,420), so a
freshly constructed model outputs exactly 0.  Benchmarks and parity tests
therefore overwrite *all* parameters with values drawn from a numpy PCG64 stream keyed by
(seed, crc32(parameter name)) -- stable across machines and torch versions, and independent of
construction order, so the reference modules, the oracle and this package see identical weights.
"""
from __future__ import annotations

import zlib
from typing import Dict, Iterable, Tuple

import numpy as np
import torch


def synth_tensor(name: str, shape: Tuple[int, ...], seed: int) -> torch.Tensor:
    rng = np.random.default_rng([seed, zlib.crc32(name.encode())])
    if len(shape) >= 2:
        fan_in = int(np.prod(shape[1:]))
        a = rng.standard_normal(shape, dtype=np.float32) * np.float32(1.0 / np.sqrt(fan_in))
    elif name.endswith("weight"):  # GroupNorm / LayerNorm gain
        a = np.float32(1.0) + np.float32(0.1) * rng.standard_normal(shape, dtype=np.float32)
    else:  # biases
        a = np.float32(0.05) * rng.standard_normal(shape, dtype=np.float32)
    return torch.from_numpy(np.ascontiguousarray(a))


@torch.no_grad()
def fill_named_tensors_(named: Iterable[Tuple[str, torch.Tensor]], seed: int = 0) -> None:
    """In-place overwrite of every floating tensor in ``named`` (parameters or state_dict items)."""
    for name, p in named:
        if not torch.is_floating_point(p):
            continue
        p.copy_(synth_tensor(name, tuple(p.shape), seed).to(p.device, p.dtype))


@torch.no_grad()
def fill_module_(module: torch.nn.Module, seed: int = 0) -> torch.nn.Module:
    fill_named_tensors_(module.state_dict().items(), seed)
    return module


def synth_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int = 0) -> Dict[str, torch.Tensor]:
    return {k: synth_tensor(k, tuple(s), seed) for k, s in shapes.items()}


def synth_images(batch: int, channels: int, size: int, seed: int) -> torch.Tensor:
    """x_0 ~ U[-1, 1] like normalised images."""
    rng = np.random.default_rng([seed, 1])
    return torch.from_numpy(rng.random((batch, channels, size, size), dtype=np.float32) * 2 - 1)


def synth_normal(shape: Tuple[int, ...], seed: int) -> torch.Tensor:
    rng = np.random.default_rng([seed, 2])
    return torch.from_numpy(rng.standard_normal(shape, dtype=np.float32))


# ---- the weight mode SURVEY.md section 8(d) prescribes ------------------------------------------------------------------
def build_survey_init(factory, seed: int = 0) -> torch.nn.Module:
    """``factory()`` under ``torch.manual_seed(seed)`` -- i.e. the reference's own default initialisation (nn.Conv2d /
    nn.Linear kaiming-uniform, GroupNorm 1/0) -- and then EVERY all-zero floating parameter re-drawn N(0, 0.02^2)
    (the zero-initialised ResBlock / attention / head convs and all the zero biases), from a numpy PCG64
    stream keyed by (seed, crc32(parameter name)).  Deterministic for a given torch version; the oracle and the product
    share the resulting state_dict."""
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(seed)
        m = factory()
    with torch.no_grad():
        for name, p in m.named_parameters():
            if torch.is_floating_point(p) and p.numel() and not bool(p.any()):
                rng = np.random.default_rng([seed, zlib.crc32(name.encode()), 20])
                p.copy_(torch.from_numpy(rng.standard_normal(tuple(p.shape), dtype=np.float32) * np.float32(0.02)))
    return m
