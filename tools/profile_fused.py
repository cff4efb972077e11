"""Developer aid: role-level cycle accounting of the fused EdgeTransition kernel (CTA 0)."""
import ctypes as C
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
eng = FrameDiffEngine(0, prec)
eng.load_weights(fo.synthetic_weights(0))
np.random.seed(0)
r7 = fo.random_frames(B, N, seed=0)
f = fo.init_feats(r7)
f["t"] = torch.full((B,), 0.5)
eng.forward(f, want_atoms=False)
lib = eng.lib
buf = (C.c_longlong * 32)()
lib.fd_debug_tc_profile(eng._h, 1, None)
eng.forward(f, want_atoms=False)
torch.cuda.synchronize()
lib.fd_debug_tc_profile(eng._h, 1, buf)
v = list(buf)
tiles = max(v[3], 1)
print(f"precision {prec} B={B} N={N}: CTA0 processed {tiles} tiles")
names = {0: "producer total", 1: "producer wait w_empty", 2: "producer wait z_empty", 8: "mma total", 9: "mma wait w_full", 10: "mma wait a_full",
         12: "mma wait y_empty+z_full", 16: "epi(warp2) total", 17: "epi wait t1_full", 19: "epi wait h2_full", 20: "epi wait y_full",
         21: "epi LN section"}
for k, n in names.items():
    print(f"  {n:28s} {v[k]:12d} cyc  {v[k] / tiles:10.0f} /tile")
lib.fd_debug_tc_profile(eng._h, 0, None)
