"""Synthetic weights and inputs for benchmarks, profiling and smoke runs (no dataset / checkpoint needed).

`synthetic_weights` is a deterministic random initialisation of the FrameDiff architecture over the library's own parameter schema
(every term active: no zero-initialised 'final' layers, scales that keep activations O(1) through the four blocks).  The generator is
numpy's legacy MT19937 in schema order, so the arrays are identical on every machine — and identical to the ones the parity tests
feed to both the CUDA path and the CPU oracle (tests/test_host_cpu.py pins that).  `init_feats` mirrors `Sampler.sample`'s feature
dict (experiments/inference_se3_diffusion.py:432-449) for a batch of frames."""
import math
from typing import Dict

import numpy as np
import torch

from .engine import param_schema


def synthetic_weights(seed: int = 0) -> Dict[str, np.ndarray]:
    rs = np.random.RandomState(seed)
    out = {}
    for name, shape in param_schema():
        if name.endswith("head_weights"):
            w = 0.5413 + 0.3 * rs.standard_normal(shape)
        elif len(shape) == 1:
            is_ln_gain = (".ln.weight" in name or "ipa_ln" in name or "norm1.weight" in name or "norm2.weight" in name
                          or "layer_norm.weight" in name or name.endswith("embedder.5.weight"))
            w = 1.0 + 0.1 * rs.standard_normal(shape) if is_ln_gain else 0.1 * rs.standard_normal(shape)
        else:
            gain = 0.3 if "bb_update" in name else (1.5 if ("linear_q_points" in name or "linear_kv_points" in name) else 1.0)
            w = gain * rs.standard_normal(shape) / math.sqrt(shape[1])
        out[name] = w.astype(np.float32)
    return out


def random_frames(B: int, N: int, seed: int = 0, trans_std: float = 10.0) -> torch.Tensor:
    """[B,N,7] frames with uniformly random rotations (unit quaternions, w >= 0) and N(0, trans_std² Å²) translations — the scale of
    the diffuser's prior (data/r3_diffuser.py:39-40)."""
    rs = np.random.RandomState(seed)
    q = rs.standard_normal((B, N, 4))
    q /= np.linalg.norm(q, axis=-1, keepdims=True)
    q *= np.where(q[..., :1] < 0, -1.0, 1.0)
    t = rs.standard_normal((B, N, 3)) * trans_std
    return torch.tensor(np.concatenate([q, t], -1), dtype=torch.float32)


def init_feats(rigids7: torch.Tensor, t: float = 0.5) -> Dict[str, torch.Tensor]:
    r = torch.as_tensor(rigids7, dtype=torch.float32)
    if r.ndim == 2:
        r = r[None]
    B, N = r.shape[:2]
    return {"res_mask": torch.ones(B, N, dtype=torch.float64), "fixed_mask": torch.zeros(B, N, dtype=torch.float64),
            "seq_idx": torch.arange(1, N + 1)[None].repeat(B, 1), "torsion_angles_sin_cos": torch.zeros(B, N, 7, 2, dtype=torch.float64),
            "sc_ca_t": torch.zeros(B, N, 3, dtype=torch.float64), "rigids_t": r, "t": torch.full((B,), float(t))}


def training_batch(engine, B: int, N: int, seed: int = 1, pad_last: int = 0) -> Dict[str, torch.Tensor]:
    """Synthetic training batch of the shape Experiment.loss_fn consumes (BASELINE config 4): CA random walk
    (cumsum N(0, 2.2^2), centred) with uniformly random frames as rigids_0, t ~ U(0.01, 1), noised per example on the GPU with the product's
    own forward_marginal (SE3Diffuser.forward_marginal, data/se3_diffuser.py:43-110; numpy draws like the reference's DataLoader workers),
    random unit psi torsions.  pad_last > 0 pads that many trailing residues of the last example (res_mask 0)."""
    rs = np.random.RandomState(seed)
    feats = {k: [] for k in ("rigids_0", "rigids_t", "rot_score", "trans_score", "rot_score_scaling", "trans_score_scaling", "res_mask", "fixed_mask",
                             "seq_idx", "torsion_angles_sin_cos", "sc_ca_t", "t")}
    for b in range(B):
        q = rs.standard_normal((N, 4)); q /= np.linalg.norm(q, axis=-1, keepdims=True)
        ca = np.cumsum(rs.standard_normal((N, 3)) * 2.2, axis=0); ca -= ca.mean(0)
        r0 = np.concatenate([q, ca], -1)
        t = float(rs.uniform(0.01, 1.0))
        res_mask = np.ones(N)
        if pad_last and b == B - 1:
            res_mask[N - pad_last:] = 0.0
        d = engine.forward_marginal(torch.tensor(r0, dtype=torch.float32), t, rs.standard_normal((N, 3)), rs.uniform(size=N), rs.standard_normal((N, 3)),
                                    diffuse_mask=res_mask)
        tors = rs.standard_normal((N, 7, 2)); tors /= np.linalg.norm(tors, axis=-1, keepdims=True)
        feats["rigids_0"].append(r0); feats["rigids_t"].append(d["rigids_t"].cpu().numpy())
        feats["rot_score"].append(d["rot_score"].cpu().numpy()); feats["trans_score"].append(d["trans_score"].cpu().numpy())
        feats["rot_score_scaling"].append(d["rot_score_scaling"]); feats["trans_score_scaling"].append(d["trans_score_scaling"])
        feats["res_mask"].append(res_mask); feats["fixed_mask"].append(np.zeros(N)); feats["seq_idx"].append(np.arange(1, N + 1) * res_mask.astype(int))
        feats["torsion_angles_sin_cos"].append(tors); feats["sc_ca_t"].append(np.zeros((N, 3))); feats["t"].append(t)
    out = {k: torch.tensor(np.stack(v)) for k, v in feats.items()}
    out["rigids_t"] = out["rigids_t"].float(); out["seq_idx"] = out["seq_idx"].long()
    return out
