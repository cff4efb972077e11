"""Runs a few ScoreNetwork forwards (synthetic weights) — the target command for ncu captures.
usage: python tools/profile_forward.py [precision] [B] [N] [reps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from se3_diffusion_b200 import synthetic as fo
from se3_diffusion_b200 import FrameDiffEngine  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
N = int(sys.argv[3]) if len(sys.argv) > 3 else 256
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
eng = FrameDiffEngine(0, prec)
eng.load_weights(fo.synthetic_weights(0))
np.random.seed(0)
r7 = fo.random_frames(B, N, seed=0)
f = fo.init_feats(r7)
f["t"] = torch.full((B,), 0.5)
for _ in range(reps):
    eng.forward(f, want_atoms=False)
torch.cuda.synchronize()
print("done", prec, B, N)
