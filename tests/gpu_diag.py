"""Not a test: prints stage-by-stage CUDA-vs-oracle errors (one gpurun call = maximum information).
usage: python tests/gpu_diag.py [fp32|bf16x3|bf16] > gpurun_out/diag.txt"""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import golden  # noqa: E402
from gpu_common import engine, feats_from_golden  # noqa: E402
from oracle import framediff_oracle as fo  # noqa: E402


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return f"SHAPE {a.shape} vs {b.shape}"
    e = np.abs(a - b)
    return f"max|err| {np.nanmax(e):.3e}  max|ref| {np.abs(b).max():.3e}  rel {np.nanmax(e) / max(np.abs(b).max(), 1e-30):.3e}  nan {int(np.isnan(a).sum())}"


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    print("device", torch.cuda.get_device_name(0), "precision", prec)
    e = engine(prec)
    w = fo.as_torch_weights(fo.synthetic_weights(0))
    for gi in (1, 2):
        g = golden(f"forward_synth_{gi}")
        f = feats_from_golden(g)
        B, N = f["rigids_t"].shape[:2]
        Np = (N + 3) // 4 * 4
        e.set_debug(True)
        try:
            out = e.forward(f)
            torch.cuda.synchronize()
        except Exception:
            traceback.print_exc()
            return
        trace = {}
        with torch.no_grad():
            ref = fo.score_network_forward(w, f, trace=trace)
        print(f"--- golden {gi}: B={B} N={N}")
        print("node_embed ", rel(e.debug_fetch("node_embed", (B, N, 256)), trace["node_embed"].numpy()))
        print("edge_embed ", rel(e.debug_fetch("edge_embed", (B, N, N, 128)), trace["edge_embed"].numpy()))
        for b in range(4):
            print(f"attn_{b}     ", rel(e.debug_fetch(f"attn_{b}", (B, 8, N, Np))[..., :N], trace[f"attn_{b}"].numpy()))
            fe = e.debug_fetch(f"ipa_feats_{b}", (B, N, 2688)); fr = trace[f"ipa_feats_{b}"].numpy()
            print(f"feats_{b} o   ", rel(fe[..., :2048], fr[..., :2048]))
            print(f"feats_{b} pt  ", rel(fe[..., 2048:2432], fr[..., 2048:2432]))
            print(f"feats_{b} pair", rel(fe[..., 2432:], fr[..., 2432:]))
            print(f"node_{b}     ", rel(e.debug_fetch(f"node_{b}", (B, N, 256)), trace[f"node_{b}"].numpy()))
            print(f"quat_{b}     ", rel(e.debug_fetch(f"quat_{b}", (B, N, 4)), trace[f"quat_{b}"].numpy()))
            print(f"trans_{b}    ", rel(e.debug_fetch(f"trans_{b}", (B, N, 3)), trace[f"trans_{b}"].numpy()))
            if b < 3:
                print(f"edge_{b}     ", rel(e.debug_fetch(f"edge_{b}", (B, N, N, 128)), trace[f"edge_{b}"].numpy()))
        for k in ("psi", "trans_score", "rot_score", "rigids", "atom37", "atom14"):
            print(f"out {k:12s}", rel(out[k].cpu().numpy(), ref[k].numpy()), "| vs golden", rel(out[k].cpu().numpy(), g["out_" + k]))
        e.set_debug(False)
    # timing of a mid-size forward + stage split
    np.random.seed(0)
    for (B, N) in ((4, 128), (8, 256)):
        r7 = torch.stack([fo.sample_ref(N) for _ in range(B)])
        f = fo.init_feats(r7); f["t"] = torch.full((B,), 0.5, dtype=torch.float64)
        e.forward(f); torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(3):
            e.forward(f)
        torch.cuda.synchronize()
        dt = (time.time() - t0) / 3
        print(f"forward B={B} N={N}: {dt * 1e3:.2f} ms  ({e.forward_flops(B, N) / dt / 1e12:.2f} TFLOP/s executed)")
        e.stage_timing(True)
        e.forward(f); torch.cuda.synchronize()
        for k, (ms, nl) in e.stage_times().items():
            print(f"   {k:16s} {ms:9.3f} ms  {nl:4d} launches")
        e.stage_timing(False)


if __name__ == "__main__":
    main()
