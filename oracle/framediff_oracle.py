"""CPU oracle for the FrameDiff hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A from-scratch CPU restatement (PyTorch-CPU / numpy / scipy, the same third-party stack the reference itself
runs on) of the one path this repository accelerates:

    ScoreNetwork.forward                      /root/reference/model/score_network.py:170-215
    IpaScore.forward + IPA + EdgeTransition   /root/reference/model/ipa_pytorch.py:194-672
    SE3Diffuser.{forward_marginal,score,reverse,sample_ref,calc_*_score}
                                              /root/reference/data/se3_diffuser.py:43-268
    SO3Diffuser / R3Diffuser arithmetic       /root/reference/data/so3_diffuser.py, data/r3_diffuser.py
    all_atom.compute_backbone                 /root/reference/data/all_atom.py:152-174
    Experiment.inference_fn (reverse loop)    /root/reference/experiments/train_se3_diffusion.py:718-818

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import
this module, and only as the checker / the timed CPU baseline.  The product package ``se3_diffusion_b200`` never
imports it and fails loudly when its CUDA library is missing.

PARITY PINNING: the reference ships no tests or golden vectors.  This oracle is pinned against
outputs of the unmodified reference executed in the build container (``tests/golden/make_golden.py`` imports
``/root/reference`` through ``tests/golden/ref_harness.py`` and commits the vectors under ``tests/golden/*.npz``);
``tests/test_oracle_golden.py`` checks the oracle against them on every CPU test run.

It is written functionally (plain weight dict + tensors) rather than as nn.Modules; every function cites the
reference lines it follows.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
from scipy.spatial.transform import Rotation as _SciRot

F32 = torch.float32
F64 = torch.float64

# --------------------------------------------------------------------------------------------------------------
# Model hyper-parameters (config/base.yaml:25-67) — compile-time constants of the product kernels too.
# --------------------------------------------------------------------------------------------------------------
C_S, C_Z, C_HID, C_SKIP = 256, 128, 256, 64
N_HEADS, N_QK_PTS, N_V_PTS = 8, 8, 12
N_BLOCKS, TFMR_HEADS, TFMR_LAYERS = 4, 4, 2
IDX_EMBED, NUM_BINS, MIN_BIN, MAX_BIN = 32, 22, 1e-5, 20.0
COORD_SCALE = 0.1
R3_MIN_B, R3_MAX_B = 0.1, 20.0
SO3_MIN_SIGMA, SO3_MAX_SIGMA, SO3_NUM_SIGMA, SO3_NUM_OMEGA = 0.1, 1.5, 1000, 1000
IGSO3_L = 1000


# --------------------------------------------------------------------------------------------------------------
# Parameter schema + deterministic synthetic weights
# --------------------------------------------------------------------------------------------------------------
def param_schema():
    """The 282 state-dict entries (name, shape) of the reference ScoreNetwork."""
    s = []

    def lin(name, o, i):
        s.append((name + ".weight", (o, i)))
        s.append((name + ".bias", (o,)))

    def ln(name, c):
        s.append((name + ".weight", (c,)))
        s.append((name + ".bias", (c,)))

    e = "embedding_layer."
    lin(e + "node_embedder.0", 256, 65); lin(e + "node_embedder.2", 256, 256); lin(e + "node_embedder.4", 256, 256)
    ln(e + "node_embedder.5", 256)
    lin(e + "edge_embedder.0", 128, 120); lin(e + "edge_embedder.2", 128, 128); lin(e + "edge_embedder.4", 128, 128)
    ln(e + "edge_embedder.5", 128)
    for b in range(N_BLOCKS):
        t = "score_model.trunk."
        s.append((t + f"ipa_{b}.head_weights", (N_HEADS,)))
        lin(t + f"ipa_{b}.linear_q", N_HEADS * C_HID, C_S)
        lin(t + f"ipa_{b}.linear_kv", 2 * N_HEADS * C_HID, C_S)
        lin(t + f"ipa_{b}.linear_q_points", N_HEADS * N_QK_PTS * 3, C_S)
        lin(t + f"ipa_{b}.linear_kv_points", N_HEADS * (N_QK_PTS + N_V_PTS) * 3, C_S)
        lin(t + f"ipa_{b}.linear_b", N_HEADS, C_Z)
        lin(t + f"ipa_{b}.down_z", C_Z // 4, C_Z)
        lin(t + f"ipa_{b}.linear_out", C_S, N_HEADS * (C_Z // 4 + C_HID + N_V_PTS * 4))
        lin(t + f"ipa_{b}.linear_rbf", 1, 20)
        ln(t + f"ipa_ln_{b}", C_S)
        lin(t + f"skip_embed_{b}", C_SKIP, C_S)
        d = C_S + C_SKIP
        for l in range(TFMR_LAYERS):
            p = t + f"seq_tfmr_{b}.layers.{l}."
            s.append((p + "self_attn.in_proj_weight", (3 * d, d)))
            s.append((p + "self_attn.in_proj_bias", (3 * d,)))
            lin(p + "self_attn.out_proj", d, d)
            lin(p + "linear1", d, d); lin(p + "linear2", d, d)
            ln(p + "norm1", d); ln(p + "norm2", d)
        lin(t + f"post_tfmr_{b}", C_S, d)
        for k in (1, 2, 3):
            lin(t + f"node_transition_{b}.linear_{k}", C_S, C_S)
        ln(t + f"node_transition_{b}.ln", C_S)
        lin(t + f"bb_update_{b}.linear", 6, C_S)
        if b < N_BLOCKS - 1:
            p = t + f"edge_transition_{b}."
            lin(p + "initial_embed", C_Z, C_S)
            lin(p + "trunk.0", 3 * C_Z, 3 * C_Z); lin(p + "trunk.2", 3 * C_Z, 3 * C_Z)
            lin(p + "final_layer", C_Z, 3 * C_Z)
            ln(p + "layer_norm", C_Z)
    p = "score_model.torsion_pred."
    lin(p + "linear_1", C_S, C_S); lin(p + "linear_2", C_S, C_S); lin(p + "linear_3", C_S, C_S)
    lin(p + "linear_final", 2, C_S)
    return s


def synthetic_weights(seed: int = 0) -> Dict[str, np.ndarray]:
    """Deterministic random weights exercising every term (no zero-initialised 'final' layers).

    Uses numpy's legacy MT19937 RandomState so the same arrays are produced on any machine.  Scales keep
    activations O(1) through the 4 blocks so that softmaxes are neither flat nor saturated.
    """
    rs = np.random.RandomState(seed)
    out = {}
    for name, shape in param_schema():
        if name.endswith("head_weights"):
            w = 0.5413 + 0.3 * rs.standard_normal(shape)
        elif len(shape) == 1:
            is_ln_gain = (".ln.weight" in name or "ipa_ln" in name or "norm1.weight" in name or "norm2.weight" in name
                          or "layer_norm.weight" in name or name.endswith("embedder.5.weight"))
            if is_ln_gain:
                w = 1.0 + 0.1 * rs.standard_normal(shape)
            else:
                w = 0.1 * rs.standard_normal(shape)
        else:
            fan_in = shape[1]
            gain = 1.0
            if "bb_update" in name:
                gain = 0.3
            if "linear_q_points" in name or "linear_kv_points" in name:
                gain = 1.5
            w = gain * rs.standard_normal(shape) / math.sqrt(fan_in)
        out[name] = w.astype(np.float32)
    return out


def as_torch_weights(w) -> Dict[str, torch.Tensor]:
    return {k: torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v).detach().to(F32).contiguous()
            for k, v in w.items()}


# --------------------------------------------------------------------------------------------------------------
# Small building blocks
# --------------------------------------------------------------------------------------------------------------
def _linear(x, w, prefix):
    return torch.nn.functional.linear(x, w[prefix + ".weight"], w[prefix + ".bias"])


def _layer_norm(x, w, prefix):
    c = x.shape[-1]
    return torch.nn.functional.layer_norm(x, (c,), w[prefix + ".weight"], w[prefix + ".bias"], 1e-5)


def timestep_embedding(t, dim=IDX_EMBED, max_positions=10000):
    """model/score_network.py:35-47."""
    half = dim // 2
    freq = torch.exp(torch.arange(half, dtype=F32) * -(math.log(max_positions) / (half - 1)))
    arg = (t * max_positions).float()[:, None] * freq[None, :]
    return torch.cat([torch.sin(arg), torch.cos(arg)], dim=1)


def index_embedding(idx, dim=IDX_EMBED, max_len=2056):
    """model/score_network.py:14-32 (idx is an integer tensor; the division happens in fp32)."""
    k = torch.arange(dim // 2)
    arg = idx[..., None] * math.pi / (max_len ** (2 * k[None] / dim))
    return torch.cat([torch.sin(arg), torch.cos(arg)], dim=-1)


def distogram(ca, min_bin=MIN_BIN, max_bin=MAX_BIN, num_bins=NUM_BINS):
    """data/utils.py:570-580 — strict inequalities, last upper edge 1e8, all-zero row where d == 0."""
    d = torch.linalg.norm(ca[:, :, None, :] - ca[:, None, :, :], dim=-1)[..., None]
    lower = torch.linspace(min_bin, max_bin, num_bins)
    upper = torch.cat([lower[1:], lower.new_tensor([1e8])])
    return ((d > lower) * (d < upper)).type(ca.dtype)


# quaternion / rotation helpers (openfold/utils/rigid_utils.py:185-287) -------------------------------------
def quat_to_rotmat(q):
    a, b, c, d = q.unbind(-1)
    rows = [
        torch.stack([a * a + b * b - c * c - d * d, 2 * b * c - 2 * a * d, 2 * b * d + 2 * a * c], -1),
        torch.stack([2 * b * c + 2 * a * d, a * a - b * b + c * c - d * d, 2 * c * d - 2 * a * b], -1),
        torch.stack([2 * b * d - 2 * a * c, 2 * c * d + 2 * a * b, a * a - b * b - c * c + d * d], -1),
    ]
    return torch.stack(rows, -2)


def quat_mul(p, q):
    a1, b1, c1, d1 = p.unbind(-1)
    a2, b2, c2, d2 = q.unbind(-1)
    return torch.stack([
        a1 * a2 - b1 * b2 - c1 * c2 - d1 * d2,
        a1 * b2 + b1 * a2 + c1 * d2 - d1 * c2,
        a1 * c2 - b1 * d2 + c1 * a2 + d1 * b2,
        a1 * d2 + b1 * c2 - c1 * b2 + d1 * a2,
    ], -1)


def quat_mul_vec(q, v):
    """q ⊗ (0, v)  — rigid_utils.py:266."""
    zero = torch.zeros_like(v[..., :1])
    return quat_mul(q, torch.cat([zero, v], -1))


def quat_invert(q):
    """rigid_utils.py:282."""
    conj = q * q.new_tensor([1.0, -1.0, -1.0, -1.0])
    return conj / torch.sum(q * q, dim=-1, keepdim=True)


def rotmat_to_quat_eigh(rot):
    """rigid_utils.py:208-227 — largest-eigenvalue eigenvector of the 4x4 K matrix (sign is arbitrary)."""
    xx, xy, xz = rot[..., 0, 0], rot[..., 0, 1], rot[..., 0, 2]
    yx, yy, yz = rot[..., 1, 0], rot[..., 1, 1], rot[..., 1, 2]
    zx, zy, zz = rot[..., 2, 0], rot[..., 2, 1], rot[..., 2, 2]
    k = [
        [xx + yy + zz, zy - yz, xz - zx, yx - xy],
        [zy - yz, xx - yy - zz, xy + yx, xz + zx],
        [xz - zx, xy + yx, yy - xx - zz, yz + zy],
        [yx - xy, xz + zx, yz + zy, zz - xx - yy],
    ]
    k = (1.0 / 3.0) * torch.stack([torch.stack(r, -1) for r in k], -2)
    _, vec = torch.linalg.eigh(k)
    return vec[..., -1]


def rot_apply(rot, p):
    return torch.einsum("...ij,...j->...i", rot, p)


def quat_to_rotvec(quat, eps=1e-6):
    """data/utils.py:582-599."""
    flip = (quat[..., :1] < 0).float()
    quat = (-1 * quat) * flip + (1 - flip) * quat
    angle = 2 * torch.atan2(torch.linalg.norm(quat[..., 1:], dim=-1), quat[..., 0])
    a2 = angle * angle
    small = 2 + a2 / 12 + 7 * a2 * a2 / 2880
    large = angle / torch.sin(angle / 2 + eps)
    is_small = (angle <= 1e-3).float()
    scale = small * is_small + (1 - is_small) * large
    return scale[..., None] * quat[..., 1:]


# --------------------------------------------------------------------------------------------------------------
# IGSO(3) + R^3 schedules (data/so3_diffuser.py, data/r3_diffuser.py)
# --------------------------------------------------------------------------------------------------------------
def so3_sigma(t):
    """so3_diffuser.py:192-199 (logarithmic schedule)."""
    t = np.asarray(t, dtype=np.float64)
    if np.any(t < 0) or np.any(t > 1):
        raise ValueError(f"Invalid t={t}")
    return np.log(t * np.exp(SO3_MAX_SIGMA) + (1 - t) * np.exp(SO3_MIN_SIGMA))


_DISCRETE_SIGMA = None


def discrete_sigma():
    global _DISCRETE_SIGMA
    if _DISCRETE_SIGMA is None:
        _DISCRETE_SIGMA = so3_sigma(np.linspace(0.0, 1.0, SO3_NUM_SIGMA))
    return _DISCRETE_SIGMA


def discrete_omega():
    return np.linspace(0, np.pi, SO3_NUM_OMEGA + 1)[1:]


def so3_t_to_idx(t):
    """so3_diffuser.py:187-213: digitize(sigma(t), discrete_sigma) - 1."""
    return np.digitize(so3_sigma(t), discrete_sigma()) - 1


def so3_diffusion_coef(t):
    """so3_diffuser.py:201-209."""
    s = so3_sigma(t)
    return np.sqrt(2 * (np.exp(SO3_MAX_SIGMA) - np.exp(SO3_MIN_SIGMA)) * s / np.exp(s))


def _igso3_expansion_np(omega, eps, L=IGSO3_L):
    """so3_diffuser.py:9-49, numpy 1-D branch."""
    ls = np.arange(L)[None]
    omega = omega[..., None]
    p = (2 * ls + 1) * np.exp(-ls * (ls + 1) * eps ** 2 / 2) * np.sin(omega * (ls + 1 / 2)) / np.sin(omega / 2)
    return p.sum(axis=-1)


def _igso3_score_np(exp, omega, eps, L=IGSO3_L):
    """so3_diffuser.py:71-117, numpy branch."""
    ls = np.arange(L)[None]
    omega = omega[..., None]
    hi = np.sin(omega * (ls + 1 / 2))
    dhi = (ls + 1 / 2) * np.cos(omega * (ls + 1 / 2))
    lo = np.sin(omega / 2)
    dlo = 1 / 2 * np.cos(omega / 2)
    d_sigma = (2 * ls + 1) * np.exp(-ls * (ls + 1) * eps ** 2 / 2) * (lo * dhi - hi * dlo) / lo ** 2
    return d_sigma.sum(axis=-1) / (exp + 1e-4)


_IGSO3_ROWS: Dict[int, dict] = {}


def igso3_row(idx: int) -> dict:
    """One sigma-row of the reference's 1000x1000 cache (so3_diffuser.py:151-180), built on demand."""
    idx = int(idx)
    if idx not in _IGSO3_ROWS:
        om = discrete_omega()
        sig = discrete_sigma()[idx]
        exp_vals = _igso3_expansion_np(om, sig)
        pdf = exp_vals * (1 - np.cos(om)) / np.pi
        cdf = pdf.cumsum() / SO3_NUM_OMEGA * np.pi
        score_norms = _igso3_score_np(exp_vals, om, sig)
        scaling = np.sqrt(np.abs(np.sum(score_norms ** 2 * pdf) / np.sum(pdf))) / np.sqrt(3)
        _IGSO3_ROWS[idx] = dict(pdf=pdf, cdf=cdf, score_norms=score_norms, score_scaling=scaling)
    return _IGSO3_ROWS[idx]


def so3_score_scaling(t):
    return igso3_row(so3_t_to_idx(t))["score_scaling"]


def so3_torch_score(vec, t, use_cached_score=False, eps=1e-6):
    """so3_diffuser.py:274-305.  vec [B,N,3] (fp32), t [B] tensor.  Returns float64 [B,N,3].

    NB the mixed precision of the reference is preserved: sin/cos of (l+1/2)·omega and the quotient-rule
    numerator are evaluated in fp32 (omega is fp32, `ls + 1/2` is a default-dtype tensor), the Gaussian factor and
    the sum over l in fp64 (sigma is a float64 numpy scalar).
    """
    omega = torch.linalg.norm(vec, dim=-1) + eps
    t_np = t.detach().cpu().numpy()
    idx = so3_t_to_idx(t_np)
    if use_cached_score:
        rows = np.stack([igso3_row(i)["score_norms"] for i in np.atleast_1d(idx)])
        score_norms_t = torch.tensor(rows)
        omega_idx = torch.bucketize(omega, torch.tensor(discrete_omega()[:-1]))
        omega_scores = torch.gather(score_norms_t, 1, omega_idx)
    else:
        sigma = torch.tensor(discrete_sigma()[idx])[:, None]  # [B,1] float64
        ls = torch.arange(IGSO3_L)[None, None]
        om = omega[..., None]
        sg = sigma[..., None]
        gauss = (2 * ls + 1) * torch.exp(-ls * (ls + 1) * sg ** 2 / 2)
        p = (gauss * torch.sin(om * (ls + 1 / 2)) / torch.sin(om / 2)).sum(dim=-1)
        hi = torch.sin(om * (ls + 1 / 2))
        dhi = (ls + 1 / 2) * torch.cos(om * (ls + 1 / 2))
        lo = torch.sin(om / 2)
        dlo = 1 / 2 * torch.cos(om / 2)
        d_sigma = (gauss * (lo * dhi - hi * dlo) / lo ** 2).sum(dim=-1)
        omega_scores = d_sigma / (p + 1e-4)
    return omega_scores[..., None] * vec / (omega[..., None] + eps)


def r3_marginal_b_t(t):
    """r3_diffuser.py:42."""
    return t * R3_MIN_B + 0.5 * (t ** 2) * (R3_MAX_B - R3_MIN_B)


def r3_b_t(t):
    if np.any(np.asarray(t) < 0) or np.any(np.asarray(t) > 1):
        raise ValueError(f"Invalid t={t}")
    return R3_MIN_B + t * (R3_MAX_B - R3_MIN_B)


def r3_score(x_t, x_0, t, use_torch=False, scale=False):
    """r3_diffuser.py:158-166."""
    exp_fn = torch.exp if use_torch else np.exp
    if scale:
        x_t = x_t * COORD_SCALE
        x_0 = x_0 * COORD_SCALE
    return -(x_t - exp_fn(-0.5 * r3_marginal_b_t(t)) * x_0) / (1 - exp_fn(-r3_marginal_b_t(t)))


def r3_score_scaling(t):
    return 1 / np.sqrt(1 - np.exp(-r3_marginal_b_t(t)))


def score_scaling(t):
    """se3_diffuser.py:155-158."""
    return so3_score_scaling(t), r3_score_scaling(t)


# --------------------------------------------------------------------------------------------------------------
# Backbone atoms (data/all_atom.py:152-174 + openfold/utils/feats.py:165-228, aatype == ALA)
# --------------------------------------------------------------------------------------------------------------
_ALA_N = (-0.525, 1.363, 0.000)
_ALA_C = (1.526, -0.000, -0.000)
_ALA_CB = (-0.529, -0.774, -1.205)
_ALA_O_PSI = (0.627, 1.062, 0.000)


def compute_backbone(rot, trans, psi):
    """rot [...,3,3], trans [...,3] (Å), psi [...,2]=(sin,cos) -> atom37 [...,37,3], mask, atom14 [...,14,3]."""
    rot = rot.to(F32)
    trans = trans.to(F32)
    psi = psi.to(rot.dtype)
    s, c = psi[..., 0], psi[..., 1]
    one, zero = torch.ones_like(s), torch.zeros_like(s)
    # psi frame -> backbone: default_frame[3] ∘ Rx(psi); default rot = diag(1,-1,-1), trans = C
    rx = torch.stack([torch.stack([one, zero, zero], -1),
                      torch.stack([zero, c, -s], -1),
                      torch.stack([zero, s, c], -1)], -2)
    dflt = rot.new_tensor([[1.0, 0.0, 0.0], [0.0, -1.0, 0.0], [0.0, 0.0, -1.0]])
    psi_rot_bb = dflt @ rx
    psi_rot = rot @ psi_rot_bb
    psi_trans = rot_apply(rot, rot.new_tensor(_ALA_C).expand_as(trans)) + trans

    def place(p):
        return rot_apply(rot, rot.new_tensor(p).expand_as(trans)) + trans

    n, ca, cpos, cb = place(_ALA_N), place((0.0, 0.0, 0.0)), place(_ALA_C), place(_ALA_CB)
    o = rot_apply(psi_rot, rot.new_tensor(_ALA_O_PSI).expand_as(trans)) + psi_trans
    atom14 = torch.zeros(trans.shape[:-1] + (14, 3), dtype=rot.dtype)
    atom14[..., 0, :], atom14[..., 1, :], atom14[..., 2, :], atom14[..., 3, :], atom14[..., 4, :] = n, ca, cpos, o, cb
    atom37 = torch.zeros(trans.shape[:-1] + (37, 3), dtype=rot.dtype)
    atom37[..., :3, :] = atom14[..., :3, :]
    atom37[..., 3, :] = atom14[..., 4, :]
    atom37[..., 4, :] = atom14[..., 3, :]
    mask = torch.any(atom37 != 0, dim=-1)
    return atom37, mask, atom14


# --------------------------------------------------------------------------------------------------------------
# Score network
# --------------------------------------------------------------------------------------------------------------
def embed(w, seq_idx, t, fixed_mask, sc_ca):
    """Embedder.forward, model/score_network.py:103-154."""
    B, N = seq_idx.shape
    t_emb = timestep_embedding(t)[:, None, :].expand(B, N, IDX_EMBED)
    prot = torch.cat([t_emb, fixed_mask[..., None]], dim=-1)  # [B,N,33]
    node_in = torch.cat([prot, index_embedding(seq_idx)], dim=-1).float()  # [B,N,65]
    rel = (seq_idx[:, :, None] - seq_idx[:, None, :]).reshape(B, N * N)
    pair = torch.cat([
        prot[:, :, None, :].expand(B, N, N, 33).reshape(B, N * N, 33),
        prot[:, None, :, :].expand(B, N, N, 33).reshape(B, N * N, 33),
        index_embedding(rel),
        distogram(sc_ca).reshape(B, N * N, NUM_BINS),
    ], dim=-1).float()
    p = "embedding_layer.node_embedder."
    h = torch.relu(_linear(node_in, w, p + "0"))
    h = torch.relu(_linear(h, w, p + "2"))
    node = _layer_norm(_linear(h, w, p + "4"), w, p + "5")
    p = "embedding_layer.edge_embedder."
    h = torch.relu(_linear(pair, w, p + "0"))
    h = torch.relu(_linear(h, w, p + "2"))
    edge = _layer_norm(_linear(h, w, p + "4"), w, p + "5").reshape(B, N, N, C_Z)
    return node, edge


def ipa(w, pre, s, z, quat, trans, mask, trace=None):
    """InvariantPointAttention.forward, model/ipa_pytorch.py:303-471.  trans is in scaled (×0.1) units."""
    B, N, _ = s.shape
    H, C, PQ, PV = N_HEADS, C_HID, N_QK_PTS, N_V_PTS
    rot = quat_to_rotmat(quat)  # [B,N,3,3]
    q = _linear(s, w, pre + "linear_q").view(B, N, H, C)
    kv = _linear(s, w, pre + "linear_kv").view(B, N, H, 2 * C)
    k, v = kv[..., :C], kv[..., C:]

    def points(lin_name, npts):
        p = _linear(s, w, pre + lin_name)  # [B,N,3*H*npts] laid out [x-block | y-block | z-block]
        p = torch.stack(torch.split(p, p.shape[-1] // 3, dim=-1), dim=-1)  # [B,N,H*npts,3]
        p = rot_apply(rot[:, :, None], p) + trans[:, :, None]
        return p.view(B, N, H, npts, 3)

    q_pts = points("linear_q_points", PQ)
    kv_pts = points("linear_kv_points", PQ + PV)
    k_pts, v_pts = kv_pts[..., :PQ, :], kv_pts[..., PQ:, :]

    b = _linear(z, w, pre + "linear_b")  # [B,N,N,H]
    a = torch.einsum("bihc,bjhc->bhij", q, k) * math.sqrt(1.0 / (3 * C))
    a = a + math.sqrt(1.0 / 3) * b.permute(0, 3, 1, 2)
    d2 = ((q_pts[:, :, None] - k_pts[:, None]) ** 2).sum(-1)  # [B,N,N,H,PQ]
    hw = torch.nn.functional.softplus(w[pre + "head_weights"]) * math.sqrt(1.0 / (3 * (PQ * 9.0 / 2)))
    pt = (d2 * hw.view(1, 1, 1, H, 1)).sum(-1) * (-0.5)  # [B,N,N,H]
    a = a + pt.permute(0, 3, 1, 2)
    sq_mask = 1e5 * (mask[:, :, None] * mask[:, None, :] - 1)
    a = torch.softmax(a + sq_mask[:, None], dim=-1)  # [B,H,N,N]

    o = torch.einsum("bhij,bjhc->bihc", a, v).reshape(B, N, H * C)
    o_pt = torch.einsum("bhij,bjhpx->bihpx", a, v_pts)  # global frame
    o_pt = rot_apply(rot.transpose(-1, -2)[:, :, None, None], o_pt - trans[:, :, None, None])
    o_pt_norm = torch.sqrt((o_pt ** 2).sum(-1) + 1e-8).reshape(B, N, H * PV)
    o_pt = o_pt.reshape(B, N, H * PV, 3)
    pair_z = _linear(z, w, pre + "down_z")  # [B,N,N,32]
    o_pair = torch.einsum("bhij,bijc->bihc", a, pair_z).reshape(B, N, H * (C_Z // 4))
    feats = torch.cat([o, o_pt[..., 0], o_pt[..., 1], o_pt[..., 2], o_pt_norm, o_pair], dim=-1)
    if trace is not None:
        trace["attn"] = a
        trace["ipa_feats"] = feats
    return _linear(feats, w, pre + "linear_out")


def seq_transformer(w, pre, x, mask, float_mask_quirk=False):
    """nn.TransformerEncoder(2 × post-norm ReLU layers, d=320, 4 heads, ff=320), eval/no-grad semantics:
    padded keys are excluded and padded query rows come back as exact zeros.
    float_mask_quirk=True restates what torch does whenever autograd is recording (any training step, and loss_fn().backward() even
    on an eval-mode module): the fused fast path is off and the FLOAT `src_key_padding_mask = 1 - mask` the reference passes
    (model/ipa_pytorch.py:636) is ADDED to the attention logits — padded keys stay in the softmax with a +1 bias, nothing is zeroed."""
    B, N, D = x.shape
    H = TFMR_HEADS
    dh = D // H
    if float_mask_quirk:
        key_bias = (1.0 - mask).to(x.dtype)[:, None, None, :]
    else:
        key_bias = torch.where(mask > 0.5, 0.0, float("-inf")).to(x.dtype)[:, None, None, :]
    for l in range(TFMR_LAYERS):
        p = f"{pre}layers.{l}."
        qkv = torch.nn.functional.linear(x, w[p + "self_attn.in_proj_weight"], w[p + "self_attn.in_proj_bias"])
        q, k, v = qkv.split(D, dim=-1)
        q = q.view(B, N, H, dh).transpose(1, 2)
        k = k.view(B, N, H, dh).transpose(1, 2)
        v = v.view(B, N, H, dh).transpose(1, 2)
        att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh) + key_bias, dim=-1)
        att = torch.nan_to_num(att)  # sample with no valid residue
        y = (att @ v).transpose(1, 2).reshape(B, N, D)
        x = _layer_norm(x + _linear(y, w, p + "self_attn.out_proj"), w, p + "norm1")
        ff = _linear(torch.relu(_linear(x, w, p + "linear1")), w, p + "linear2")
        x = _layer_norm(x + ff, w, p + "norm2")
    return x if float_mask_quirk else x * mask[..., None]


def edge_transition(w, pre, node, edge):
    """EdgeTransition.forward, model/ipa_pytorch.py:218-233."""
    B, N, _ = node.shape
    nb = _linear(node, w, pre + "initial_embed")
    x = torch.cat([edge,
                   nb[:, :, None, :].expand(B, N, N, C_Z),
                   nb[:, None, :, :].expand(B, N, N, C_Z)], dim=-1).reshape(B * N * N, 3 * C_Z)
    h = torch.relu(_linear(x, w, pre + "trunk.0"))
    h = torch.relu(_linear(h, w, pre + "trunk.2"))
    y = _linear(h + x, w, pre + "final_layer")
    return _layer_norm(y, w, pre + "layer_norm").reshape(B, N, N, C_Z)


def score_network_forward(w, feats, use_cached_score=False, trace: Optional[dict] = None, float_mask_quirk=False):
    """ScoreNetwork.forward (model/score_network.py:170-215) + IpaScore.forward (model/ipa_pytorch.py:611-672).

    feats: res_mask, fixed_mask, seq_idx, t, sc_ca_t, rigids_t, torsion_angles_sin_cos (tensors, batch first).
    Returns the same dict as the reference: psi, rot_score, trans_score, rigids, atom37, atom14.
    """
    bb_mask = feats["res_mask"].type(F32)
    fixed_mask = feats["fixed_mask"].type(F32)
    edge_mask = bb_mask[..., None] * bb_mask[..., None, :]
    node0, edge = embed(w, feats["seq_idx"], feats["t"], fixed_mask, feats["sc_ca_t"])
    edge = edge * edge_mask[..., None]
    node0 = node0 * bb_mask[..., None]
    if trace is not None:
        trace["node_embed"] = node0
        trace["edge_embed"] = edge

    diffuse_mask = (1 - fixed_mask) * bb_mask
    rig_t = feats["rigids_t"].type(F32)
    quat_t, trans_t = rig_t[..., :4], rig_t[..., 4:]
    quat = quat_t.clone()
    trans = trans_t * COORD_SCALE
    node0 = node0 * bb_mask[..., None]
    node = node0 * bb_mask[..., None]
    T = "score_model.trunk."
    for b in range(N_BLOCKS):
        tr = {} if trace is not None else None
        upd = ipa(w, T + f"ipa_{b}.", node, edge, quat, trans, bb_mask, trace=tr) * bb_mask[..., None]
        node = _layer_norm(node + upd, w, T + f"ipa_ln_{b}")
        x = torch.cat([node, _linear(node0, w, T + f"skip_embed_{b}")], dim=-1)
        x = seq_transformer(w, T + f"seq_tfmr_{b}.", x, bb_mask, float_mask_quirk=float_mask_quirk)
        node = node + _linear(x, w, T + f"post_tfmr_{b}")
        p = T + f"node_transition_{b}."
        h = torch.relu(_linear(node, w, p + "linear_1"))
        h = torch.relu(_linear(h, w, p + "linear_2"))
        node = _layer_norm(_linear(h, w, p + "linear_3") + node, w, p + "ln") * bb_mask[..., None]
        upd6 = _linear(node * diffuse_mask[..., None], w, T + f"bb_update_{b}.linear")
        # Rigid.compose_q_update_vec (rigid_utils.py:1039-1063, 587-616): rotation applied to t uses the OLD quat
        rot_old = quat_to_rotmat(quat)
        dq = quat_mul_vec(quat, upd6[..., :3]) * diffuse_mask[..., None]
        trans = trans + rot_apply(rot_old, upd6[..., 3:]) * diffuse_mask[..., None]
        quat = quat + dq
        quat = quat / torch.linalg.norm(quat, dim=-1, keepdim=True)
        if trace is not None:
            trace[f"ipa_feats_{b}"] = tr["ipa_feats"]
            trace[f"attn_{b}"] = tr["attn"]
            trace[f"node_{b}"] = node
            trace[f"quat_{b}"] = quat
            trace[f"trans_{b}"] = trans
        if b < N_BLOCKS - 1:
            edge = edge_transition(w, T + f"edge_transition_{b}.", node, edge) * edge_mask[..., None]
            if trace is not None:
                trace[f"edge_{b}"] = edge

    # heads (ipa_pytorch.py:650-671, se3_diffuser.py:115-125)
    q_rel = quat_mul(quat_invert(quat), quat_t)
    rot_score = so3_torch_score(quat_to_rotvec(q_rel), feats["t"], use_cached_score) * bb_mask[..., None]
    trans_pred = trans / COORD_SCALE
    trans_score = r3_score(trans_t, trans_pred, feats["t"][:, None, None], use_torch=True, scale=True)
    trans_score = trans_score * bb_mask[..., None]
    p = "score_model.torsion_pred."
    h = _linear(torch.relu(_linear(node, w, p + "linear_1")), w, p + "linear_2") + node
    un = _linear(h, w, p + "linear_final")
    psi = un / torch.sqrt(torch.clamp(torch.sum(un ** 2, dim=-1, keepdim=True), min=1e-8))

    gt_psi = feats["torsion_angles_sin_cos"][..., 2, :]
    dm = 1 - fixed_mask[..., None]
    psi = dm * psi + (1 - dm) * gt_psi
    rigids = torch.cat([quat, trans_pred], dim=-1)
    atom37, _, atom14 = compute_backbone(quat_to_rotmat(quat), trans_pred, psi)
    return {"psi": psi, "rot_score": rot_score, "trans_score": trans_score, "rigids": rigids,
            "atom37": atom37, "atom14": atom14}


# --------------------------------------------------------------------------------------------------------------
# SE(3) diffuser host API (data/se3_diffuser.py) — numpy/scipy like the reference
# --------------------------------------------------------------------------------------------------------------
def _rotvec_from_quat7(rigids7):
    """se3_diffuser.py:11-18: quat -> rot mats (torch fp32) -> scipy rotvec (float64)."""
    r = torch.as_tensor(rigids7)
    if not r.is_floating_point():
        r = r.to(F64)
    rot = quat_to_rotmat(r[..., :4].to(F32)).cpu().numpy()
    shp = rot.shape
    rv = _SciRot.from_matrix(rot.reshape(-1, 3, 3)).as_rotvec().reshape(shp[:-2] + (3,))
    return r[..., 4:].cpu().numpy(), rv


def _assemble_rigid7(rotvec, trans):
    """se3_diffuser.py:20-29 followed by Rigid.to_tensor_7 (rot_to_quat/eigh).  Returns ([...,7] f32, rotmat f32)."""
    shp = rotvec.shape
    rotmat = _SciRot.from_rotvec(rotvec.reshape(-1, 3)).as_matrix().reshape(shp[:-1] + (3, 3))
    rot = torch.Tensor(rotmat)
    quat = rotmat_to_quat_eigh(rot)
    # Rigid.__init__ forces fp32 translations (rigid_utils.py:899), so the 7-vector state is fp32 every step.
    return torch.cat([quat, torch.as_tensor(np.asarray(trans)).to(F32)], dim=-1), rot


def compose_rotvec(r1, r2):
    """data/utils.py:184-189."""
    m = np.einsum("...ij,...jk->...ik", _SciRot.from_rotvec(r1).as_matrix(), _SciRot.from_rotvec(r2).as_matrix())
    return _SciRot.from_matrix(m).as_rotvec()


def sample_ref(n_samples: int):
    """se3_diffuser.py:216-268 with both diffusions on.  RNG order: randn(n,3), rand(n), normal(n,3)."""
    x = np.random.randn(n_samples, 3)
    x /= np.linalg.norm(x, axis=-1, keepdims=True)
    u = np.random.rand(n_samples)
    ang = np.interp(u, igso3_row(so3_t_to_idx(1))["cdf"], discrete_omega())
    rot_ref = x * ang[:, None]
    trans_ref = np.random.normal(size=(n_samples, 3)) / COORD_SCALE
    r7, _ = _assemble_rigid7(rot_ref, trans_ref)
    return r7


def reverse_step(rigids7_t, rot_score, trans_score, t, dt, diffuse_mask=None, center=True, noise_scale=1.0,
                 noise=None):
    """SE3Diffuser.reverse (se3_diffuser.py:160-214) = SO3Diffuser.reverse (so3_diffuser.py:330-366)
    + R3Diffuser.reverse (r3_diffuser.py:106-146).  `noise`=(z_rot, z_trans) injects N(0,1) draws, otherwise the
    global numpy RNG is used in the reference's order (rotation first, then translation)."""
    trans_t, rot_t = _rotvec_from_quat7(rigids7_t)
    if not np.isscalar(t):
        raise ValueError(f"{t} must be a scalar.")
    g = so3_diffusion_coef(t)
    z_r = np.random.normal(size=rot_score.shape) if noise is None else noise[0]
    perturb = (g ** 2) * rot_score * dt + g * np.sqrt(dt) * (noise_scale * z_r)
    n = int(np.prod(rot_t.shape[:-1]))
    rot_1 = compose_rotvec(rot_t.reshape(n, 3), perturb.reshape(n, 3)).reshape(rot_t.shape)

    x = trans_t * COORD_SCALE
    g_t = np.sqrt(r3_b_t(t))
    f_t = -0.5 * r3_b_t(t) * x
    z_x = np.random.normal(size=trans_score.shape) if noise is None else noise[1]
    pert = (f_t - g_t ** 2 * trans_score) * dt + g_t * np.sqrt(dt) * (noise_scale * z_x)
    x1 = x - pert
    if center:
        com = np.sum(x1, axis=-2) / np.sum(np.ones(x.shape[:-1]), axis=-1)[..., None]
        x1 = x1 - com[..., None, :]
    trans_1 = x1 / COORD_SCALE
    if diffuse_mask is not None:
        m = diffuse_mask[..., None]
        trans_1 = m * trans_1 + (1 - m) * trans_t
        rot_1 = m * rot_1 + (1 - m) * rot_t
    r7, rotmat = _assemble_rigid7(rot_1, trans_1)
    return r7, rotmat


def forward_marginal(rigids7_0, t, diffuse_mask=None):
    """SE3Diffuser.forward_marginal (se3_diffuser.py:43-110) for a single example [N,7]."""
    trans_0, rot_0 = _rotvec_from_quat7(rigids7_0)
    n = int(np.prod(rot_0.shape[:-1]))
    # SO3Diffuser.sample (so3_diffuser.py:233-248)
    x = np.random.randn(n, 3)
    x /= np.linalg.norm(x, axis=-1, keepdims=True)
    u = np.random.rand(n)
    sampled = x * np.interp(u, igso3_row(so3_t_to_idx(t))["cdf"], discrete_omega())[:, None]
    rot_score = so3_torch_score(torch.tensor(sampled), torch.tensor(t)[None]).numpy().reshape(rot_0.shape)
    rot_t = compose_rotvec(rot_0, sampled).reshape(rot_0.shape)
    x0 = trans_0 * COORD_SCALE
    x_t = np.random.normal(loc=np.exp(-0.5 * r3_marginal_b_t(t)) * x0, scale=np.sqrt(1 - np.exp(-r3_marginal_b_t(t))))
    trans_score = r3_score(x_t, x0, t)
    trans_t = x_t / COORD_SCALE
    if diffuse_mask is not None:
        m = diffuse_mask[..., None]
        rot_t = m * rot_t + (1 - m) * rot_0
        trans_t = m * trans_t + (1 - m) * trans_0
        trans_score = m * trans_score
        rot_score = m * rot_score
    r7, _ = _assemble_rigid7(rot_t, trans_t)
    return {"rigids_t": r7, "trans_score": trans_score, "rot_score": rot_score,
            "trans_score_scaling": r3_score_scaling(t), "rot_score_scaling": so3_score_scaling(t)}


def diffuser_score(rigids7_0, rigids7_t, t):
    """SE3Diffuser.score (se3_diffuser.py:134-153): scores rot_t itself; translations unscaled (quirk C.8)."""
    tran_0, _ = _rotvec_from_quat7(rigids7_0)
    tran_t, rot_t = _rotvec_from_quat7(rigids7_t)
    rot_score = so3_torch_score(torch.tensor(rot_t), torch.tensor(t)[None]).numpy()
    return r3_score(tran_t, tran_0, t), rot_score


# --------------------------------------------------------------------------------------------------------------
# Reverse-diffusion loop (Experiment.inference_fn, experiments/train_se3_diffusion.py:718-818)
# --------------------------------------------------------------------------------------------------------------
def init_feats(rigids7, B=None):
    """Sampler.sample's init_feats (experiments/inference_se3_diffusion.py:432-449) for a batch [B,N,7]."""
    r = torch.as_tensor(rigids7).to(F32)
    if r.ndim == 2:
        r = r[None]
    B, N = r.shape[:2]
    return {
        "res_mask": torch.ones(B, N, dtype=F64),
        "seq_idx": torch.arange(1, N + 1)[None].repeat(B, 1),
        "fixed_mask": torch.zeros(B, N, dtype=F64),
        "torsion_angles_sin_cos": torch.zeros(B, N, 7, 2, dtype=F64),
        "sc_ca_t": torch.zeros(B, N, 3, dtype=F64),
        "rigids_t": r,
    }


def inference_loop(w, feats, num_t=500, min_t=0.01, center=True, aux_traj=False, self_condition=True,
                   noise_scale=1.0, noise_fn=None, max_steps=None):
    """Restates Experiment.inference_fn.  noise_fn(step, shape) -> (z_rot, z_trans) injects noise; default draws
    from the global numpy RNG in the reference's order.  max_steps truncates the loop (CPU-baseline timing)."""
    f = dict(feats)
    B = f["rigids_t"].shape[0]
    ones = torch.ones(B)
    reverse_steps = np.linspace(min_t, 1.0, num_t)[::-1]
    dt = 1 / num_t
    all_rigids = [f["rigids_t"].numpy().copy()]
    all_bb, all_trans0, all_bb0 = [], [], []
    with torch.no_grad():
        if self_condition:
            f["t"] = reverse_steps[0] * ones
            f["sc_ca_t"] = score_network_forward(w, f)["rigids"][..., 4:]
        rigid_pred = None
        for step, t in enumerate(reverse_steps):
            if max_steps is not None and step >= max_steps:
                break
            fixed_mask = f["fixed_mask"] * f["res_mask"]
            diffuse_mask = (1 - f["fixed_mask"]) * f["res_mask"]
            if t > min_t:
                f["t"] = t * ones
                out = score_network_forward(w, f)
                rigid_pred = out["rigids"]
                f["sc_ca_t"] = rigid_pred[..., 4:]
                noise = None if noise_fn is None else noise_fn(step, tuple(out["rot_score"].shape))
                r7, rotmat = reverse_step(f["rigids_t"], out["rot_score"].numpy(), out["trans_score"].numpy(), t, dt,
                                          diffuse_mask=diffuse_mask.numpy(), center=center, noise_scale=noise_scale,
                                          noise=noise)
            else:
                out = score_network_forward(w, f)
                r7 = out["rigids"]
                rotmat = quat_to_rotmat(r7[..., :4])
            f["rigids_t"] = r7
            if aux_traj:
                all_rigids.append(f["rigids_t"].numpy().copy())
            psi_pred = out["psi"]
            if aux_traj:
                a0 = compute_backbone(quat_to_rotmat(rigid_pred[..., :4]), rigid_pred[..., 4:], psi_pred)[0]
                all_bb0.append(a0.numpy())
                tp = diffuse_mask[..., None] * rigid_pred[..., 4:] + fixed_mask[..., None] * f["rigids_t"][..., 4:]
                all_trans0.append(tp.numpy())
            all_bb.append(compute_backbone(rotmat, f["rigids_t"][..., 4:], psi_pred)[0].numpy())
    flip = lambda x: np.flip(np.stack(x), (0,))
    ret = {"prot_traj": flip(all_bb)}
    if aux_traj:
        ret["rigid_traj"] = flip(all_rigids)
        ret["trans_traj"] = flip(all_trans0)
        ret["psi_pred"] = psi_pred[None]
        ret["rigid_0_traj"] = flip(all_bb0)
    return ret


# --------------------------------------------------------------------------------------------------------------
# DSM training loss, forward values (Experiment.loss_fn, experiments/train_se3_diffusion.py:524-693) — the terms computed from the
# model outputs and the noised batch (the model call and the self-conditioning coin flip stay with the caller).
# --------------------------------------------------------------------------------------------------------------
DEFAULT_EXP_CONF = dict(trans_loss_weight=1.0, rot_loss_weight=0.5, rot_loss_t_threshold=0.2, separate_rot_loss=True,
                        trans_x0_threshold=1.0, coordinate_scaling=0.1, bb_atom_loss_weight=1.0, bb_atom_loss_t_filter=0.25,
                        dist_mat_loss_weight=1.0, dist_mat_loss_t_filter=0.25, aux_loss_weight=0.25)   # config/base.yaml:104-115


def loss_terms(model_out, batch, exp_conf=None, diffuse_trans=True, diffuse_rot=True):
    """Per-sample loss terms [B] and the normalised scalars of loss_fn (train_se3_diffusion.py:538-680)."""
    c = dict(DEFAULT_EXP_CONF, **(exp_conf or {}))
    T = lambda x: torch.as_tensor(x)
    bb_mask = T(batch["res_mask"])
    diffuse_mask = 1 - T(batch["fixed_mask"])
    loss_mask = bb_mask * diffuse_mask
    B, N = bb_mask.shape
    t = T(batch["t"])
    gt_rot, gt_trans = T(batch["rot_score"]), T(batch["trans_score"])
    rot_sc, trans_sc = T(batch["rot_score_scaling"]), T(batch["trans_score_scaling"])
    batch_loss_mask = torch.any(bb_mask.bool(), dim=-1)
    pred_rot = T(model_out["rot_score"]) * diffuse_mask[..., None]
    pred_trans = T(model_out["trans_score"]) * diffuse_mask[..., None]
    denom = loss_mask.sum(dim=-1) + 1e-10
    # translation: score loss above the threshold, x0 loss below (:553-573)
    trans_score_loss = torch.sum((gt_trans - pred_trans) ** 2 * loss_mask[..., None] / trans_sc[:, None, None] ** 2, dim=(-1, -2)) / denom
    gt_x0 = T(batch["rigids_0"])[..., 4:] * c["coordinate_scaling"]
    pred_x0 = T(model_out["rigids"])[..., 4:] * c["coordinate_scaling"]
    trans_x0_loss = torch.sum((gt_x0 - pred_x0) ** 2 * loss_mask[..., None], dim=(-1, -2)) / denom
    trans_loss = trans_score_loss * (t > c["trans_x0_threshold"]) + trans_x0_loss * (t <= c["trans_x0_threshold"])
    trans_loss = trans_loss * c["trans_loss_weight"] * int(diffuse_trans)
    # rotation (:576-607)
    if c["separate_rot_loss"]:
        gt_angle = torch.norm(gt_rot, dim=-1, keepdim=True)
        gt_axis = gt_rot / (gt_angle + 1e-6)
        pr_angle = torch.norm(pred_rot, dim=-1, keepdim=True)
        pr_axis = pred_rot / (pr_angle + 1e-6)
        axis_loss = torch.sum((gt_axis - pr_axis) ** 2 * loss_mask[..., None], dim=(-1, -2)) / denom
        angle_loss = torch.sum((gt_angle - pr_angle) ** 2 * loss_mask[..., None] / rot_sc[:, None, None] ** 2, dim=(-1, -2)) / denom
        angle_loss = angle_loss * c["rot_loss_weight"] * (t > c["rot_loss_t_threshold"])
        rot_loss = angle_loss + axis_loss
    else:
        rot_loss = torch.sum((gt_rot - pred_rot) ** 2 * loss_mask[..., None] / rot_sc[:, None, None] ** 2, dim=(-1, -2)) / denom
        rot_loss = rot_loss * c["rot_loss_weight"] * (t > c["rot_loss_t_threshold"])
    rot_loss = rot_loss * int(diffuse_rot)
    # backbone atoms (:611-630): ground truth through compute_backbone of rigids_0 (fp32) and the psi torsion (index 2)
    pred_a = T(model_out["atom37"])[:, :, :5]
    r0 = T(batch["rigids_0"]).to(F32)
    gt_psi = T(batch["torsion_angles_sin_cos"])[..., 2, :]
    gt_a37, gt_m37, _ = compute_backbone(quat_to_rotmat(r0[..., :4]), r0[..., 4:], gt_psi)
    gt_a, a_mask = gt_a37[:, :, :5], gt_m37[:, :, :5]
    bb_m = a_mask * loss_mask[..., None]
    bb_atom_loss = torch.sum((pred_a - gt_a) ** 2 * bb_m[..., None], dim=(-1, -2, -3)) / (bb_m.sum(dim=(-1, -2)) + 1e-10)
    bb_atom_loss = bb_atom_loss * c["bb_atom_loss_weight"] * (t < c["bb_atom_loss_t_filter"]) * c["aux_loss_weight"]
    # pairwise distances of the 5N backbone atoms, pairs closer than 6 A in the ground truth (:633-660)
    gt_f, pr_f = gt_a.reshape(B, N * 5, 3), pred_a.reshape(B, N * 5, 3)
    gt_d = torch.linalg.norm(gt_f[:, :, None, :] - gt_f[:, None, :, :], dim=-1)
    pr_d = torch.linalg.norm(pr_f[:, :, None, :] - pr_f[:, None, :, :], dim=-1)
    flat_loss = torch.tile(loss_mask[:, :, None], (1, 1, 5)).reshape(B, N * 5)
    flat_res = torch.tile(bb_mask[:, :, None], (1, 1, 5)).reshape(B, N * 5)
    gt_d = gt_d * flat_loss[..., None]
    pr_d = pr_d * flat_loss[..., None]
    pair_mask = flat_loss[..., None] * flat_res[:, None, :]
    pair_mask = pair_mask * (gt_d < 6)
    dist_mat_loss = torch.sum((gt_d - pr_d) ** 2 * pair_mask, dim=(1, 2)) / (torch.sum(pair_mask, dim=(1, 2)) - N)
    dist_mat_loss = dist_mat_loss * c["dist_mat_loss_weight"] * (t < c["dist_mat_loss_t_filter"]) * c["aux_loss_weight"]
    final = rot_loss + trans_loss + bb_atom_loss + dist_mat_loss
    norm = lambda x: x.sum() / (batch_loss_mask.sum() + 1e-10)
    return {"batch_train_loss": final, "batch_rot_loss": rot_loss, "batch_trans_loss": trans_loss, "batch_bb_atom_loss": bb_atom_loss,
            "batch_dist_mat_loss": dist_mat_loss, "total_loss": norm(final), "rot_loss": norm(rot_loss), "trans_loss": norm(trans_loss),
            "bb_atom_loss": norm(bb_atom_loss), "dist_mat_loss": norm(dist_mat_loss)}


# --------------------------------------------------------------------------------------------------------------
# Intermittent eval metrics (analysis/metrics.py:120-132) — the checker of fd_ca_metrics
# --------------------------------------------------------------------------------------------------------------
CA_CA = 3.80209737096          # data/residue_constants.py:29


def ca_ca_distance(ca_pos, tol=0.1):
    """analysis/metrics.py:120-125."""
    d = np.linalg.norm(ca_pos - np.roll(ca_pos, 1, axis=0), axis=-1)[1:]
    return np.mean(np.abs(d - CA_CA)), np.mean(d < (CA_CA + tol))


def ca_ca_clashes(ca_pos, tol=1.5):
    """analysis/metrics.py:127-132."""
    d2 = np.linalg.norm(ca_pos[:, None, :] - ca_pos[None, :, :], axis=-1)
    inter = d2[np.where(np.triu(d2, k=0) > 0)]
    cl = inter < tol
    return np.sum(cl), np.mean(cl)
