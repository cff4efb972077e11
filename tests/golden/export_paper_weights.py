"""Copies the reference checkpoint tensors (an input artefact, not source) into weights/paper_weights.npz.

weights/ is git-ignored (70 MB) but travels to the GPU box with the gpurun snapshot, so the optional
paper-weights parity tests can run there.  Run in the build container:  python tests/golden/export_paper_weights.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

if __name__ == "__main__":
    root = os.path.dirname(os.path.dirname(HERE))
    os.makedirs(os.path.join(root, "weights"), exist_ok=True)
    for name in (("paper_weights", "best_weights") if "--all" in sys.argv else ("paper_weights",)):   # 70 MB each: only what the tests use
        sd = rh.load_reference_checkpoint(name + ".pth")
        out = os.path.join(root, "weights", name + ".npz")
        np.savez(out, **{k: v.numpy() for k, v in sd.items()})
        print("wrote", out, len(sd), "tensors")
