"""Developer aid: per-stage CUDA-event times of one forward.  usage: python tools/stage_times.py prec B N"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from se3_diffusion_b200 import synthetic as fo
from se3_diffusion_b200 import FrameDiffEngine
prec = sys.argv[1]; B = int(sys.argv[2]); N = int(sys.argv[3])
eng = FrameDiffEngine(0, prec); eng.load_weights(fo.synthetic_weights(0))
np.random.seed(0)
r7 = fo.random_frames(B, N, seed=0)
f = fo.init_feats(r7); f["t"] = torch.full((B,), 0.5)
eng.forward(f, want_atoms=False); torch.cuda.synchronize()
t0 = time.time()
for _ in range(3): eng.forward(f, want_atoms=False)
torch.cuda.synchronize(); dt = (time.time() - t0) / 3
print(f"{prec} B={B} N={N}: forward {dt*1e3:.2f} ms, {eng.forward_flops(B,N)/dt/1e12:.1f} TFLOP/s executed, {eng.forward_flops(B,N,False)/dt/1e12:.1f} ref-equivalent")
eng.stage_timing(True); eng.forward(f, want_atoms=False); torch.cuda.synchronize()
for k, (ms, nl) in eng.stage_times().items(): print(f"   {k:16s} {ms:9.3f} ms  {nl:4d} launches")
