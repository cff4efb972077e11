"""se3_diffusion_b200 — B200-native FrameDiff hot path (ScoreNetwork.forward + SE3Diffuser + the reverse loop).

Hand-written sm_100a CUDA behind a C ABI (include/framediff_b200.h, lib/libframediff_b200.so); this package is the
Python host side mirroring the reference's interface for that path.  No CPU fallback: importing the engine without
the built library, or constructing it without a CUDA device, raises.
"""
from ._lib import FrameDiffError, LIB_PATH  # noqa: F401

__all__ = ["FrameDiffEngine", "FrameDiffError", "LIB_PATH"]


def __getattr__(name):
    if name == "FrameDiffEngine":
        from .engine import FrameDiffEngine
        return FrameDiffEngine
    raise AttributeError(name)
