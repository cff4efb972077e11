"""Throughput of the native PDB writer next to the reference's Python writer (when /root/reference is importable), on the
trajectory the sampler produces: [T frames, N residues, 37, 3] with the 5 backbone atoms.
usage: python tools/bench_pdb_writer.py [T] [N] [ref_frames]"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from se3_diffusion_b200 import pdb_writer  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 500
N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ref_frames = int(sys.argv[3]) if len(sys.argv) > 3 else 20
rng = np.random.RandomState(0)
pos = np.zeros((T, N, 37, 3), np.float32)
pos[:, :, :5] = (rng.randn(T, N, 5, 3) * 20).astype(np.float32)
with tempfile.TemporaryDirectory() as d:
    pdb_writer.write_prot_to_pdb(pos[:2], os.path.join(d, "w.pdb"), no_indexing=True)     # warm-up (library load)
    t0 = time.perf_counter()
    p = pdb_writer.write_prot_to_pdb(pos, os.path.join(d, "ours.pdb"), no_indexing=True)
    dt = time.perf_counter() - t0
    size = os.path.getsize(p)
    print(f"native: {T} frames x {N} residues -> {size / 1e6:.1f} MB in {dt * 1e3:.1f} ms  ({T / dt:.0f} frames/s, {size / dt / 1e6:.0f} MB/s, 1 core)")
    try:
        import ref_harness as rh
        rh.install_stubs()
        from analysis import utils as au
        t0 = time.perf_counter()
        q = au.write_prot_to_pdb(pos[:ref_frames], os.path.join(d, "ref.pdb"), no_indexing=True)
        dr = time.perf_counter() - t0
        same = open(q, "rb").read() == pdb_writer.format_pdb(pos[:ref_frames])
        print(f"reference (analysis/utils.py, Python loops): {ref_frames} frames in {dr:.2f} s ({ref_frames / dr:.1f} frames/s) -> "
              f"{T} frames would take {dr / ref_frames * T:.1f} s; bytes identical: {same}; speed-up {(T / dt) / (ref_frames / dr):.0f}x")
    except Exception as e:  # reference tree not present (GPU box)
        print("reference writer not available here:", type(e).__name__)
