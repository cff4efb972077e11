"""Shared helpers of the -m gpu tests: engine construction with the deterministic synthetic weights."""
import functools

import numpy as np
import torch

from oracle import framediff_oracle as fo


@functools.lru_cache(maxsize=None)
def synthetic_state(seed=0):
    return fo.synthetic_weights(seed)


_ENGINES = {}


def engine(precision="fp32", weights="synth"):
    """One engine per process (the handle owns its workspace); precision switched on demand."""
    from se3_diffusion_b200 import FrameDiffEngine
    key = weights
    if key not in _ENGINES:
        e = FrameDiffEngine(0, precision)
        if weights == "synth":
            e.load_weights(synthetic_state(0))
        else:
            e.load_weights(dict(np.load(weights)))
        _ENGINES[key] = e
    e = _ENGINES[key]
    if e.precision != precision:
        e.set_precision(precision)
    return e


def feats_from_golden(g):
    return {k[3:]: torch.tensor(v) for k, v in g.items() if k.startswith("in_")}


def numpy_noise(seed, B, N, num_t):
    """The reference's np.random draw order for a batch: per sample randn(N,3), rand(N), normal(N,3) for the
    prior; then per step normal(B,N,3) for SO(3) followed by normal(B,N,3) for R^3."""
    np.random.seed(seed)
    za, ua, zt = [], [], []
    for _ in range(B):
        za.append(np.random.randn(N, 3)); ua.append(np.random.rand(N)); zt.append(np.random.normal(size=(N, 3)))
    zr, zx = [], []
    for _ in range(num_t - 1):
        zr.append(np.random.normal(size=(B, N, 3))); zx.append(np.random.normal(size=(B, N, 3)))
    return {"z_axis": np.stack(za), "u_angle": np.stack(ua), "z_trans0": np.stack(zt),
            "z_rot": np.stack(zr) if zr else None, "z_trans": np.stack(zx) if zx else None}
