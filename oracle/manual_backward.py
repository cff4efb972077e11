"""Hand-derived backward pass of the FrameDiff training step — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The reference has no backward code: `Experiment.update_fn` (/root/reference/experiments/train_se3_diffusion.py:320-326) calls
`loss.backward()` and torch autograd differentiates `loss_fn` (:524-693) and `ScoreNetwork.forward`
(/root/reference/model/score_network.py:170-215, model/ipa_pytorch.py:194-672).  The CUDA training path of this repository
(se3_diffusion_b200/csrc/fd_train.cuh) implements that derivative by hand, kernel by kernel.  This module is the CPU
restatement of exactly that decomposition — the same tape, the same grouping of GEMMs / per-residue formulas, torch used only as
an array library with autograd OFF — so that every formula is checked against autograd through the pinned oracle
(tests/test_oracle_golden.py::test_manual_backward_vs_autograd) before it is transliterated to CUDA, and so that the -m gpu
tests can compare the kernels' intermediate gradients stage by stage.

Parity pinning: `oracle.framediff_oracle` (pinned to the reference's goldens) + autograd is the authority; this file is only
accepted where it reproduces it (gradients of all 272 used parameters, 1e-5 relative in float64 arithmetic).

Semantics are the ones torch uses whenever autograd records (float key-padding mask ADDED to the sequence-attention logits,
SURVEY Appendix C.2 — `float_mask_quirk=True` in the oracle).
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

from . import framediff_oracle as fo

C_S, C_Z, C_HID, C_SKIP = fo.C_S, fo.C_Z, fo.C_HID, fo.C_SKIP
H, PQ, PV = fo.N_HEADS, fo.N_QK_PTS, fo.N_V_PTS
NBLK = fo.N_BLOCKS
TF_D, TF_H, TF_L = C_S + C_SKIP, fo.TFMR_HEADS, fo.TFMR_LAYERS
T_ = "score_model.trunk."


# ---------------------------------------------------------------------------------------------------------------------
# primitives (each one = one CUDA kernel family in fd_train.cuh)
# ---------------------------------------------------------------------------------------------------------------------
def lin_fwd(x, w, p):
    return x @ w[p + ".weight"].T + w[p + ".bias"]


def lin_bwd(g, p, x, dy, w, need_dx=True):
    """dW += dy^T x, db += colsum(dy); returns dx = dy W."""
    dy2, x2 = dy.reshape(-1, dy.shape[-1]), x.reshape(-1, x.shape[-1])
    acc(g, p + ".weight", dy2.T @ x2)
    acc(g, p + ".bias", dy2.sum(0))
    return dy @ w[p + ".weight"] if need_dx else None


def acc(g, name, val):
    g[name] = g[name] + val if name in g else val


def ln_fwd(x, w, p, eps=1e-5):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    rstd = 1.0 / torch.sqrt(var + eps)
    xh = (x - mu) * rstd
    return xh * w[p + ".weight"] + w[p + ".bias"], (xh, rstd)


def ln_bwd(g, p, saved, dy, w):
    xh, rstd = saved
    c = xh.shape[-1]
    acc(g, p + ".weight", (dy * xh).reshape(-1, c).sum(0))
    acc(g, p + ".bias", dy.reshape(-1, c).sum(0))
    dxh = dy * w[p + ".weight"]
    return rstd * (dxh - dxh.mean(-1, keepdim=True) - xh * (dxh * xh).mean(-1, keepdim=True))


def rot_from_quat(q):
    return fo.quat_to_rotmat(q)


def quat_grad_from_rot_grad(q, G):
    """d/dq of the (unnormalised) polynomial quat->rot map (rigid_utils.py:185), contracted with G = dL/dR [...,3,3]."""
    a, b, c, d = q.unbind(-1)
    G00, G01, G02 = G[..., 0, 0], G[..., 0, 1], G[..., 0, 2]
    G10, G11, G12 = G[..., 1, 0], G[..., 1, 1], G[..., 1, 2]
    G20, G21, G22 = G[..., 2, 0], G[..., 2, 1], G[..., 2, 2]
    da = 2 * (a * (G00 + G11 + G22) + d * (G10 - G01) + c * (G02 - G20) + b * (G21 - G12))
    db = 2 * (b * (G00 - G11 - G22) + c * (G01 + G10) + d * (G02 + G20) + a * (G21 - G12))
    dc = 2 * (c * (-G00 + G11 - G22) + b * (G01 + G10) + a * (G02 - G20) + d * (G12 + G21))
    dd = 2 * (d * (-G00 - G11 + G22) + a * (G10 - G01) + b * (G02 + G20) + c * (G12 + G21))
    return torch.stack([da, db, dc, dd], -1)


# ---------------------------------------------------------------------------------------------------------------------
# forward with tape
# ---------------------------------------------------------------------------------------------------------------------
def _pair_features(seq_idx, t, fixed_mask, sc_ca):
    B, N = seq_idx.shape
    t_emb = fo.timestep_embedding(t)[:, None, :].expand(B, N, fo.IDX_EMBED)
    prot = torch.cat([t_emb, fixed_mask[..., None]], dim=-1)
    node_in = torch.cat([prot, fo.index_embedding(seq_idx)], dim=-1)
    rel = (seq_idx[:, :, None] - seq_idx[:, None, :]).reshape(B, N * N)
    pair = torch.cat([prot[:, :, None, :].expand(B, N, N, 33).reshape(B, N * N, 33),
                      prot[:, None, :, :].expand(B, N, N, 33).reshape(B, N * N, 33),
                      fo.index_embedding(rel), fo.distogram(sc_ca).reshape(B, N * N, fo.NUM_BINS)], dim=-1)
    return node_in, pair.reshape(B, N, N, 120)


def _mlp3_ln_fwd(x, w, p):
    h1 = torch.relu(lin_fwd(x, w, p + "0"))
    h2 = torch.relu(lin_fwd(h1, w, p + "2"))
    y = lin_fwd(h2, w, p + "4")
    out, ln = ln_fwd(y, w, p + "5")
    return out, (x, h1, h2, ln)


def _mlp3_ln_bwd(g, p, saved, dout, w):
    x, h1, h2, ln = saved
    dy = ln_bwd(g, p + "5", ln, dout, w)
    dh2 = lin_bwd(g, p + "4", h2, dy, w) * (h2 > 0)
    dh1 = lin_bwd(g, p + "2", h1, dh2, w) * (h1 > 0)
    lin_bwd(g, p + "0", x, dh1, w, need_dx=False)


def _points_fwd(raw, npts, rot, trans):
    """[B,N,3*H*npts] laid out [x-block|y-block|z-block] -> global points [B,N,H*npts,3] (ipa_pytorch.py:332-349)."""
    p = torch.stack(torch.split(raw, raw.shape[-1] // 3, dim=-1), dim=-1)
    return torch.einsum("bnij,bnpj->bnpi", rot, p) + trans[:, :, None], p


def _points_bwd(dP, p_local, rot):
    """dP [B,N,P,3] global -> (d raw [B,N,3P] block layout, dR [B,N,3,3], dt [B,N,3])."""
    dp = torch.einsum("bnij,bnpi->bnpj", rot, dP)             # R^T dP
    dR = torch.einsum("bnpi,bnpj->bnij", dP, p_local)
    dt = dP.sum(2)
    draw = torch.cat([dp[..., 0], dp[..., 1], dp[..., 2]], dim=-1)
    return draw, dR, dt


def ipa_fwd(w, pre, s, z, quat, trans, mask):
    B, N, _ = s.shape
    rot = rot_from_quat(quat)
    q = lin_fwd(s, w, pre + "linear_q").view(B, N, H, C_HID)
    kv = lin_fwd(s, w, pre + "linear_kv").view(B, N, H, 2 * C_HID)
    k, v = kv[..., :C_HID], kv[..., C_HID:]
    qp, qp_loc = _points_fwd(lin_fwd(s, w, pre + "linear_q_points"), PQ, rot, trans)
    kvp, kvp_loc = _points_fwd(lin_fwd(s, w, pre + "linear_kv_points"), PQ + PV, rot, trans)
    qp = qp.view(B, N, H, PQ, 3)
    kvp = kvp.view(B, N, H, PQ + PV, 3)
    kp, vp = kvp[..., :PQ, :], kvp[..., PQ:, :]
    c1, c2 = math.sqrt(1.0 / (3 * C_HID)), math.sqrt(1.0 / 3)
    hw = w[pre + "head_weights"]
    gamma = torch.nn.functional.softplus(hw) * math.sqrt(1.0 / (3 * (PQ * 9.0 / 2)))
    bias = lin_fwd(z, w, pre + "linear_b")                                   # [B,N,N,H]
    L = c1 * torch.einsum("bihc,bjhc->bhij", q, k) + c2 * bias.permute(0, 3, 1, 2)
    d2 = ((qp[:, :, None] - kp[:, None]) ** 2).sum(-1).sum(-1)              # [B,N,N,H]
    L = L - 0.5 * (d2 * gamma).permute(0, 3, 1, 2)
    L = L + (1e5 * (mask[:, :, None] * mask[:, None, :] - 1))[:, None]
    A = torch.softmax(L, dim=-1)
    o = torch.einsum("bhij,bjhc->bihc", A, v)
    optg = torch.einsum("bhij,bjhpx->bihpx", A, vp)
    optl = torch.einsum("bnji,bnhpj->bnhpi", rot, optg - trans[:, :, None, None])     # R^T (g - t)
    nrm = torch.sqrt((optl ** 2).sum(-1) + 1e-8)
    pair_z = lin_fwd(z, w, pre + "down_z")                                  # [B,N,N,32]
    opair = torch.einsum("bhij,bijc->bihc", A, pair_z)
    feats = torch.cat([o.reshape(B, N, -1), optl[..., 0].reshape(B, N, -1), optl[..., 1].reshape(B, N, -1), optl[..., 2].reshape(B, N, -1),
                       nrm.reshape(B, N, -1), opair.reshape(B, N, -1)], dim=-1)
    out = lin_fwd(feats, w, pre + "linear_out")
    tape = dict(s=s, z=z, quat=quat, trans=trans, rot=rot, q=q, k=k, v=v, qp=qp, kp=kp, vp=vp, qp_loc=qp_loc, kvp_loc=kvp_loc, gamma=gamma,
                A=A, optg=optg, optl=optl, nrm=nrm, pair_z=pair_z, feats=feats)
    return out, tape


def ipa_bwd(g, w, pre, tp, dout):
    """Returns (ds, dz, dquat, dtrans)."""
    s, z, quat, trans, rot = tp["s"], tp["z"], tp["quat"], tp["trans"], tp["rot"]
    A, q, k, v, qp, kp, vp = tp["A"], tp["q"], tp["k"], tp["v"], tp["qp"], tp["kp"], tp["vp"]
    B, N, _ = s.shape
    c1, c2 = math.sqrt(1.0 / (3 * C_HID)), math.sqrt(1.0 / 3)
    dfeats = lin_bwd(g, pre + "linear_out", tp["feats"], dout, w)
    n0 = H * C_HID
    do = dfeats[..., :n0].reshape(B, N, H, C_HID)
    dx = dfeats[..., n0:n0 + H * PV].reshape(B, N, H, PV)
    dy = dfeats[..., n0 + H * PV:n0 + 2 * H * PV].reshape(B, N, H, PV)
    dzc = dfeats[..., n0 + 2 * H * PV:n0 + 3 * H * PV].reshape(B, N, H, PV)
    dn = dfeats[..., n0 + 3 * H * PV:n0 + 4 * H * PV].reshape(B, N, H, PV)
    dopair = dfeats[..., n0 + 4 * H * PV:].reshape(B, N, H, C_Z // 4)
    # --- per-residue: o_pt local frame + norm (ipa_finish backward) ---
    doptl = torch.stack([dx, dy, dzc], -1) + dn[..., None] * tp["optl"] / tp["nrm"][..., None]
    doptg = torch.einsum("bnij,bnhpj->bnhpi", rot, doptl)                     # R dl
    gm = tp["optg"] - trans[:, :, None, None]
    dR = torch.einsum("bnhpi,bnhpj->bnij", gm, doptl)                         # optl_k = sum_m R[m,k] gm_m -> dR[m,k] = gm_m dl_k
    dtr = -doptg.sum((2, 3))
    # --- dA from the three aggregations (two batched GEMMs + the edge kernel's pair term) ---
    dA = torch.einsum("bihc,bjhc->bhij", do, v) + torch.einsum("bihpx,bjhpx->bhij", doptg, vp)
    dA = dA + torch.einsum("bihc,bijc->bhij", dopair, tp["pair_z"])
    dv = torch.einsum("bhij,bihc->bjhc", A, do)
    dvp = torch.einsum("bhij,bihpx->bjhpx", A, doptg)
    dpair_z = torch.einsum("bhij,bihc->bijc", A, dopair)
    # --- softmax backward ---
    dL = A * (dA - (A * dA).sum(-1, keepdim=True))
    # --- logits terms ---
    dq = c1 * torch.einsum("bhij,bjhc->bihc", dL, k)
    dk = c1 * torch.einsum("bhij,bihc->bjhc", dL, q)
    dbias = c2 * dL.permute(0, 2, 3, 1)                                       # [B,N,N,H]
    gam = tp["gamma"]
    Gq = torch.einsum("bhij,bjhpx->bihpx", dL, kp)                            # sum_j dL kp_j
    Gk = torch.einsum("bhij,bihpx->bjhpx", dL, qp)                            # sum_i dL qp_i
    colsum = dL.sum(2)                                                        # [B,H,N(j)]
    rowsum = dL.sum(3)                                                        # ~0 analytically; kept (masked rows are exactly representable)
    dqp = gam.view(1, 1, H, 1, 1) * (Gq - qp * rowsum.permute(0, 2, 1)[..., None, None])
    dkp = gam.view(1, 1, H, 1, 1) * (Gk - kp * colsum.permute(0, 2, 1)[..., None, None])
    d2sum = (rowsum.permute(0, 2, 1) * (qp ** 2).sum((-1, -2))).sum((0, 1)) + (colsum.permute(0, 2, 1) * (kp ** 2).sum((-1, -2))).sum((0, 1)) \
        - 2 * (qp * Gq).sum((0, 1, 3, 4))
    dgamma = -0.5 * d2sum
    hw = w[pre + "head_weights"]
    acc(g, pre + "head_weights", dgamma * torch.sigmoid(hw) * math.sqrt(1.0 / (3 * (PQ * 9.0 / 2))))
    # --- edge-tensor terms: [dbias | dpair_z] (40 wide) x [Wb ; Wd] ---
    dz = lin_bwd(g, pre + "linear_b", z, dbias, w) + lin_bwd(g, pre + "down_z", z, dpair_z, w)
    # --- points back to the local frame / projections ---
    dkvp = torch.cat([dkp, dvp], dim=3).reshape(B, N, H * (PQ + PV), 3)
    draw_q, dRq, dtq = _points_bwd(dqp.reshape(B, N, H * PQ, 3), tp["qp_loc"], rot)
    draw_kv, dRk, dtk = _points_bwd(dkvp, tp["kvp_loc"], rot)
    dR = dR + dRq + dRk
    dtr = dtr + dtq + dtk
    dkv = torch.cat([dk, dv], dim=-1).reshape(B, N, H * 2 * C_HID)
    ds = lin_bwd(g, pre + "linear_q", s, dq.reshape(B, N, -1), w) + lin_bwd(g, pre + "linear_kv", s, dkv, w) \
        + lin_bwd(g, pre + "linear_q_points", s, draw_q, w) + lin_bwd(g, pre + "linear_kv_points", s, draw_kv, w)
    return ds, dz, quat_grad_from_rot_grad(quat, dR), dtr


def tfmr_fwd(w, pre, x, mask):
    B, N, D = x.shape
    dh = D // TF_H
    kb = (1.0 - mask)[:, None, None, :]
    tapes = []
    for l in range(TF_L):
        p = f"{pre}layers.{l}."
        qkv = x @ w[p + "self_attn.in_proj_weight"].T + w[p + "self_attn.in_proj_bias"]
        q, k, v = [u.view(B, N, TF_H, dh).transpose(1, 2) for u in qkv.split(D, dim=-1)]
        P = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh) + kb, dim=-1)
        y = (P @ v).transpose(1, 2).reshape(B, N, D)
        x1, ln1 = ln_fwd(x + lin_fwd(y, w, p + "self_attn.out_proj"), w, p + "norm1")
        f1 = torch.relu(lin_fwd(x1, w, p + "linear1"))
        x2, ln2 = ln_fwd(x1 + lin_fwd(f1, w, p + "linear2"), w, p + "norm2")
        tapes.append(dict(x=x, q=q, k=k, v=v, P=P, y=y, ln1=ln1, x1=x1, f1=f1, ln2=ln2))
        x = x2
    return x, tapes


def tfmr_bwd(g, w, pre, tapes, dx):
    for l in reversed(range(TF_L)):
        p = f"{pre}layers.{l}."
        tp = tapes[l]
        B, N, D = tp["x"].shape
        dh = D // TF_H
        ds2 = ln_bwd(g, p + "norm2", tp["ln2"], dx, w)
        df1 = lin_bwd(g, p + "linear2", tp["f1"], ds2, w) * (tp["f1"] > 0)
        dx1 = ds2 + lin_bwd(g, p + "linear1", tp["x1"], df1, w)
        ds1 = ln_bwd(g, p + "norm1", tp["ln1"], dx1, w)
        dy = lin_bwd(g, p + "self_attn.out_proj", tp["y"], ds1, w).view(B, N, TF_H, dh).transpose(1, 2)
        P, q, k, v = tp["P"], tp["q"], tp["k"], tp["v"]
        dP = dy @ v.transpose(-1, -2)
        dv = P.transpose(-1, -2) @ dy
        dS = P * (dP - (P * dP).sum(-1, keepdim=True)) / math.sqrt(dh)
        dq = dS @ k
        dk = dS.transpose(-1, -2) @ q
        dqkv = torch.cat([u.transpose(1, 2).reshape(B, N, D) for u in (dq, dk, dv)], dim=-1)
        acc(g, p + "self_attn.in_proj_weight", dqkv.reshape(-1, 3 * D).T @ tp["x"].reshape(-1, D))
        acc(g, p + "self_attn.in_proj_bias", dqkv.reshape(-1, 3 * D).sum(0))
        dx = ds1 + dqkv @ w[p + "self_attn.in_proj_weight"]
    return dx


def edge_transition_fwd(w, pre, node, z, emask):
    """Separable form of EdgeTransition (ipa_pytorch.py:218-233): x = [z | nb_i | nb_j] is never materialised."""
    B, N, _ = node.shape
    nb = lin_fwd(node, w, pre + "initial_embed")
    W1, b1, Wf, bf = w[pre + "trunk.0.weight"], w[pre + "trunk.0.bias"], w[pre + "final_layer.weight"], w[pre + "final_layer.bias"]
    Pi, Qj = nb @ W1[:, C_Z:2 * C_Z].T + b1, nb @ W1[:, 2 * C_Z:].T
    Ui, Vj = nb @ Wf[:, C_Z:2 * C_Z].T + bf, nb @ Wf[:, 2 * C_Z:].T
    h1 = torch.relu(z @ W1[:, :C_Z].T + Pi[:, :, None] + Qj[:, None])
    h2 = torch.relu(lin_fwd(h1, w, pre + "trunk.2"))
    y = h2 @ Wf.T + z @ Wf[:, :C_Z].T + Ui[:, :, None] + Vj[:, None]
    zo, ln = ln_fwd(y, w, pre + "layer_norm")
    return zo * emask[..., None], dict(node=node, nb=nb, z=z, h1=h1, h2=h2, ln=ln, emask=emask)


def edge_transition_bwd(g, w, pre, tp, dzo):
    node, nb, z, h1, h2 = tp["node"], tp["nb"], tp["z"], tp["h1"], tp["h2"]
    W1, Wf, W2 = w[pre + "trunk.0.weight"], w[pre + "final_layer.weight"], w[pre + "trunk.2.weight"]
    B, N, _ = node.shape
    E = B * N * N
    dy = ln_bwd(g, pre + "layer_norm", tp["ln"], dzo * tp["emask"][..., None], w)          # [B,N,N,128]
    dh2 = (dy @ Wf) * (h2 > 0)                                                          # [B,N,N,384]
    dh1 = lin_bwd(g, pre + "trunk.2", h1, dh2, w) * (h1 > 0)
    # row / column sums of the two edge gradients that meet the node terms
    RSy, CSy = dy.sum(2), dy.sum(1)                                                     # [B,N,128]
    RS1, CS1 = dh1.sum(2), dh1.sum(1)                                                   # [B,N,384]
    dy2, dh12, z2, h22 = dy.reshape(E, C_Z), dh1.reshape(E, 3 * C_Z), z.reshape(E, C_Z), h2.reshape(E, 3 * C_Z)
    nb2 = nb.reshape(B * N, C_Z)
    dWf = dy2.T @ h22
    dWf[:, :C_Z] += dy2.T @ z2
    dWf[:, C_Z:2 * C_Z] += RSy.reshape(-1, C_Z).T @ nb2
    dWf[:, 2 * C_Z:] += CSy.reshape(-1, C_Z).T @ nb2
    acc(g, pre + "final_layer.weight", dWf)
    acc(g, pre + "final_layer.bias", dy2.sum(0))
    dW1 = torch.cat([dh12.T @ z2, RS1.reshape(-1, 3 * C_Z).T @ nb2, CS1.reshape(-1, 3 * C_Z).T @ nb2], dim=1)
    acc(g, pre + "trunk.0.weight", dW1)
    acc(g, pre + "trunk.0.bias", dh12.sum(0))
    dz = dh1 @ W1[:, :C_Z] + dy @ Wf[:, :C_Z]
    dnb = RS1 @ W1[:, C_Z:2 * C_Z] + CS1 @ W1[:, 2 * C_Z:] + RSy @ Wf[:, C_Z:2 * C_Z] + CSy @ Wf[:, 2 * C_Z:]
    dnode = lin_bwd(g, pre + "initial_embed", node, dnb, w)
    return dnode, dz


def backbone_update_fwd(w, pre, node, quat, trans, dmask):
    x = node * dmask[..., None]
    upd = lin_fwd(x, w, pre + "linear")
    rot = rot_from_quat(quat)
    dq = fo.quat_mul_vec(quat, upd[..., :3]) * dmask[..., None]
    trans_new = trans + torch.einsum("bnij,bnj->bni", rot, upd[..., 3:]) * dmask[..., None]
    qun = quat + dq
    nrm = torch.linalg.norm(qun, dim=-1, keepdim=True)
    return qun / nrm, trans_new, dict(x=x, upd=upd, quat=quat, rot=rot, nrm=nrm, qnew=qun / nrm, dmask=dmask)


def backbone_update_bwd(g, w, pre, tp, dqnew, dtnew):
    """Returns (dnode, dquat_old, dtrans_old)."""
    quat, upd, rot, dm = tp["quat"], tp["upd"], tp["rot"], tp["dmask"][..., None]
    qn = tp["qnew"]
    dqun = (dqnew - qn * (qn * dqnew).sum(-1, keepdim=True)) / tp["nrm"]
    gq = dqun * dm
    g0, g1, g2, g3 = gq.unbind(-1)
    p0, p1, p2, p3 = quat.unbind(-1)
    v1, v2, v3 = upd[..., 0], upd[..., 1], upd[..., 2]
    dp = torch.stack([g1 * v1 + g2 * v2 + g3 * v3, -g0 * v1 - g2 * v3 + g3 * v2, -g0 * v2 + g1 * v3 - g3 * v1, -g0 * v3 - g1 * v2 + g2 * v1], -1)
    dv = torch.stack([-g0 * p1 + g1 * p0 + g2 * p3 - g3 * p2, -g0 * p2 - g1 * p3 + g2 * p0 + g3 * p1, -g0 * p3 + g1 * p2 - g2 * p1 + g3 * p0], -1)
    gt = dtnew * dm
    dtv = torch.einsum("bnij,bni->bnj", rot, gt)
    dR = torch.einsum("bni,bnj->bnij", gt, upd[..., 3:])
    dquat = dqun + dp + quat_grad_from_rot_grad(quat, dR)
    dupd = torch.cat([dv, dtv], -1)
    dx = lin_bwd(g, pre + "linear", tp["x"], dupd, w)
    return dx * dm, dquat, dtnew


# ---------------------------------------------------------------------------------------------------------------------
# heads: rotation score (IGSO(3) series), translation score, torsion, backbone atoms
# ---------------------------------------------------------------------------------------------------------------------
def _igso3_score_and_dscore(omega, sigma, L=fo.IGSO3_L):
    """s(omega) = d_sigma / (p + 1e-4) (so3_diffuser.py:71-117) and ds/domega, float64.  omega [...], sigma broadcastable."""
    om = omega.double()[..., None]
    sg = torch.as_tensor(sigma, dtype=torch.float64)[..., None]
    ls = torch.arange(L, dtype=torch.float64)
    a = ls + 0.5
    gauss = (2 * ls + 1) * torch.exp(-ls * (ls + 1) * sg ** 2 / 2)
    hi, chi = torch.sin(a * om), torch.cos(a * om)
    lo, clo = torch.sin(om / 2), torch.cos(om / 2)
    p = (gauss * hi / lo).sum(-1)
    num = lo * a * chi - hi * 0.5 * clo                       # lo*dhi - hi*dlo
    dsig = (gauss * num / lo ** 2).sum(-1)
    dp = dsig                                                 # d/domega of hi/lo is exactly the same quotient
    # d/domega (num / lo^2) = num'/lo^2 - 2 num lo'/lo^3 ; num' = lo*(-a^2 hi) + 0.25 hi lo ... (product rule, cross terms cancel)
    dnum = 0.5 * clo * a * chi - lo * a * a * hi - a * chi * 0.5 * clo + hi * 0.25 * lo
    ddsig = (gauss * (dnum / lo ** 2 - 2 * num * 0.5 * clo / lo ** 3)).sum(-1)
    s = dsig / (p + 1e-4)
    ds = ddsig / (p + 1e-4) - dsig * dp / (p + 1e-4) ** 2
    return s, ds


def heads_fwd(w, node, quat, trans, feats, bb_mask, fixed_mask):
    t = feats["t"]
    rig_t = feats["rigids_t"].float()
    quat_t, trans_t = rig_t[..., :4], rig_t[..., 4:]
    # rotation score: q_rel = inv(quat) (x) quat_t -> rotvec -> IGSO(3) score
    n2 = (quat * quat).sum(-1, keepdim=True)
    qinv = quat * quat.new_tensor([1.0, -1.0, -1.0, -1.0]) / n2
    qrel = fo.quat_mul(qinv, quat_t)
    sgn = torch.where(qrel[..., :1] < 0, -1.0, 1.0)
    qf = qrel * sgn
    vn = torch.linalg.norm(qf[..., 1:], dim=-1)
    ang = 2 * torch.atan2(vn, qf[..., 0])
    small = ang <= 1e-3
    scale = torch.where(small, 2 + ang ** 2 / 12 + 7 * ang ** 4 / 2880, ang / torch.sin(ang / 2 + 1e-6))
    rv = scale[..., None] * qf[..., 1:]
    omega = torch.linalg.norm(rv, dim=-1) + 1e-6
    sigma = torch.tensor(fo.discrete_sigma()[fo.so3_t_to_idx(t.detach().cpu().numpy())])[:, None]
    s, ds = _igso3_score_and_dscore(omega, sigma)
    rot_score = (s[..., None] * rv / (omega[..., None] + 1e-6)) * bb_mask[..., None]
    trans_pred = trans / fo.COORD_SCALE
    tt = t[:, None, None]
    bt = fo.r3_marginal_b_t(tt)
    trans_score = -(trans_t * fo.COORD_SCALE - torch.exp(-0.5 * bt) * trans_pred * fo.COORD_SCALE) / (1 - torch.exp(-bt)) * bb_mask[..., None]
    p = "score_model.torsion_pred."
    a1 = torch.relu(lin_fwd(node, w, p + "linear_1"))
    hh = lin_fwd(a1, w, p + "linear_2") + node
    un = lin_fwd(hh, w, p + "linear_final")
    ssq = (un ** 2).sum(-1, keepdim=True)
    den = torch.sqrt(torch.clamp(ssq, min=1e-8))
    psi_pred = un / den
    dm = 1 - fixed_mask[..., None]
    gt_psi = feats["torsion_angles_sin_cos"][..., 2, :]
    psi = dm * psi_pred + (1 - dm) * gt_psi
    rot = rot_from_quat(quat)
    atom37, _, atom14 = fo.compute_backbone(rot, trans_pred, psi)
    out = {"psi": psi, "rot_score": rot_score, "trans_score": trans_score, "rigids": torch.cat([quat, trans_pred], -1), "atom37": atom37,
           "atom14": atom14}
    tape = dict(node=node, quat=quat, quat_t=quat_t, n2=n2, qinv=qinv, sgn=sgn, qf=qf, vn=vn, ang=ang, small=small, scale=scale, rv=rv, omega=omega,
                s=s, ds=ds, bt=bt, a1=a1, hh=hh, un=un, ssq=ssq, den=den, dm=dm, psi=psi, rot=rot, bb_mask=bb_mask)
    return out, tape


def heads_bwd(g, w, tp, dout):
    """dout: grads of rot_score, trans_score, rigids, atom37 (only atoms 0..4 matter), psi (missing keys = zero).
    Returns (dnode, dquat, dtrans[scaled units])."""
    quat, bb = tp["quat"], tp["bb_mask"][..., None]
    z3 = torch.zeros_like(tp["rv"])
    dquat = torch.zeros_like(quat)
    dtp = torch.zeros_like(z3)                                   # d trans_pred (Angstrom)
    dpsi = dout["psi"].float().clone() if "psi" in dout else torch.zeros_like(tp["un"])
    if "rigids" in dout:
        dquat = dquat + dout["rigids"][..., :4].float()
        dtp = dtp + dout["rigids"][..., 4:].float()
    if "trans_score" in dout:
        bt = tp["bt"]
        dtp = dtp + (dout["trans_score"] * bb * (torch.exp(-0.5 * bt) * fo.COORD_SCALE / (1 - torch.exp(-bt)))).float()
    # ---- atom37 (all_atom.compute_backbone): N, CA, C, CB from (R, t); O through the psi frame ----
    if "atom37" in dout:
        da = dout["atom37"].float()
        rot = tp["rot"]
        dR = torch.zeros_like(rot)
        for idx, pos in ((0, fo._ALA_N), (2, fo._ALA_C), (3, fo._ALA_CB)):
            dR = dR + torch.einsum("bni,j->bnij", da[:, :, idx], rot.new_tensor(pos))
        dtp = dtp + da[:, :, :5].sum(2)
        # O = R (D Rx(psi) o + C) + t with D = diag(1,-1,-1), o = (0.627, 1.062, 0): local = (ox, -(c oy), -(s oy)) + C
        s_, c_ = tp["psi"][..., 0].float(), tp["psi"][..., 1].float()
        ox, oy = fo._ALA_O_PSI[0], fo._ALA_O_PSI[1]
        loc = torch.stack([ox + fo._ALA_C[0] + 0 * c_, -(c_ * oy), -(s_ * oy)], -1)
        dO = da[:, :, 4]
        dR = dR + torch.einsum("bni,bnj->bnij", dO, loc)
        dloc = torch.einsum("bnij,bni->bnj", rot, dO)
        dpsi = dpsi + torch.stack([-oy * dloc[..., 2], -oy * dloc[..., 1]], -1)
        dquat = dquat + quat_grad_from_rot_grad(quat, dR)
    # ---- torsion head ----
    dpp = dpsi * tp["dm"].float()
    un, den, ssq = tp["un"], tp["den"], tp["ssq"]
    dun = dpp / den - torch.where(ssq > 1e-8, un * (un * dpp).sum(-1, keepdim=True) / den ** 3, torch.zeros_like(un))
    p = "score_model.torsion_pred."
    dhh = lin_bwd(g, p + "linear_final", tp["hh"], dun, w)
    da1 = lin_bwd(g, p + "linear_2", tp["a1"], dhh, w) * (tp["a1"] > 0)
    dnode = dhh + lin_bwd(g, p + "linear_1", tp["node"], da1, w)
    # ---- rotation score ----
    if "rot_score" in dout:
        drs = (dout["rot_score"] * bb).double()
        rv, omega, s, ds = tp["rv"].double(), tp["omega"].double(), tp["s"], tp["ds"]
        f = s / (omega + 1e-6)
        dfdom = ds / (omega + 1e-6) - s / (omega + 1e-6) ** 2
        nv = omega - 1e-6
        drv = f[..., None] * drs + (dfdom * (drs * rv).sum(-1))[..., None] * rv / torch.clamp(nv, min=1e-30)[..., None]
        drv = drv.float()
        # rv = scale(ang) * v ; ang = 2 atan2(|v|, w)
        qf, vn, ang, scale = tp["qf"], tp["vn"], tp["ang"], tp["scale"]
        v = qf[..., 1:]
        dv = scale[..., None] * drv
        dscale = (drv * v).sum(-1)
        half = ang / 2 + 1e-6
        dsc_dang = torch.where(tp["small"], ang / 6 + 7 * ang ** 3 / 720, 1 / torch.sin(half) - ang * torch.cos(half) / (2 * torch.sin(half) ** 2))
        dang = dscale * dsc_dang
        den2 = vn ** 2 + qf[..., 0] ** 2
        dvn = dang * 2 * qf[..., 0] / den2
        dw = -dang * 2 * vn / den2
        dv = dv + (dvn / torch.clamp(vn, min=1e-30))[..., None] * v
        dqrel = torch.cat([dw[..., None], dv], -1) * tp["sgn"]
        # qrel = qinv (x) quat_t : d qinv = dqrel (x) conj(quat_t)
        qt = tp["quat_t"]
        dqinv = fo.quat_mul(dqrel, qt * qt.new_tensor([1.0, -1.0, -1.0, -1.0]))
        # qinv = conj(q)/|q|^2
        cj = quat.new_tensor([1.0, -1.0, -1.0, -1.0])
        n2 = tp["n2"]
        dquat = dquat + cj * dqinv / n2 - 2 * quat * (dqinv * tp["qinv"]).sum(-1, keepdim=True) / n2
    return dnode, dquat, dtp / fo.COORD_SCALE


# ---------------------------------------------------------------------------------------------------------------------
# whole network
# ---------------------------------------------------------------------------------------------------------------------
def train_forward(w: Dict[str, torch.Tensor], feats):
    bb = feats["res_mask"].float()
    fixed = feats["fixed_mask"].float()
    emask = bb[:, :, None] * bb[:, None, :]
    node_in, pair = _pair_features(feats["seq_idx"], feats["t"], fixed, feats["sc_ca_t"])
    ne, t_ne = _mlp3_ln_fwd(node_in.float(), w, "embedding_layer.node_embedder.")
    ee, t_ee = _mlp3_ln_fwd(pair.float(), w, "embedding_layer.edge_embedder.")
    node0 = ne * bb[..., None]
    z = ee * emask[..., None]
    dmask = (1 - fixed) * bb
    rig_t = feats["rigids_t"].float()
    quat, trans = rig_t[..., :4].clone(), rig_t[..., 4:] * fo.COORD_SCALE
    node = node0
    blocks = []
    for b in range(NBLK):
        tp = {}
        ipa_out, tp["ipa"] = ipa_fwd(w, T_ + f"ipa_{b}.", node, z, quat, trans, bb)
        n1, tp["ln1"] = ln_fwd(node + ipa_out * bb[..., None], w, T_ + f"ipa_ln_{b}")
        skip = lin_fwd(node0, w, T_ + f"skip_embed_{b}")
        x = torch.cat([n1, skip], dim=-1)
        xo, tp["tf"] = tfmr_fwd(w, T_ + f"seq_tfmr_{b}.", x, bb)
        n2 = n1 + lin_fwd(xo, w, T_ + f"post_tfmr_{b}")
        p = T_ + f"node_transition_{b}."
        a1 = torch.relu(lin_fwd(n2, w, p + "linear_1"))
        a2 = torch.relu(lin_fwd(a1, w, p + "linear_2"))
        n3pre, tp["ln2"] = ln_fwd(lin_fwd(a2, w, p + "linear_3") + n2, w, p + "ln")
        n3 = n3pre * bb[..., None]
        tp.update(xo=xo, n2=n2, a1=a1, a2=a2)
        quat, trans, tp["bbu"] = backbone_update_fwd(w, T_ + f"bb_update_{b}.", n3, quat, trans, dmask)
        node = n3
        if b < NBLK - 1:
            z, tp["et"] = edge_transition_fwd(w, T_ + f"edge_transition_{b}.", node, z, emask)
        blocks.append(tp)
    out, t_heads = heads_fwd(w, node, quat, trans, feats, bb, fixed)
    tape = dict(ne=t_ne, ee=t_ee, blocks=blocks, heads=t_heads, bb=bb, emask=emask, node0=node0)
    return out, tape


def train_backward(w, tape, dout, taps=None):
    g: Dict[str, torch.Tensor] = {}
    bb, emask = tape["bb"], tape["emask"]
    dnode, dquat, dtrans = heads_bwd(g, w, tape["heads"], dout)
    dz = None
    dnode0 = torch.zeros_like(dnode)
    for b in reversed(range(NBLK)):
        tp = tape["blocks"][b]
        if b < NBLK - 1:
            dn_et, dz_in = edge_transition_bwd(g, w, T_ + f"edge_transition_{b}.", tp["et"], dz)
            dnode = dnode + dn_et
            dz = dz_in
        dn_bbu, dquat, dtrans = backbone_update_bwd(g, w, T_ + f"bb_update_{b}.", tp["bbu"], dquat, dtrans)
        dnode = dnode + dn_bbu
        p = T_ + f"node_transition_{b}."
        dy = ln_bwd(g, p + "ln", tp["ln2"], dnode * bb[..., None], w)
        da2 = lin_bwd(g, p + "linear_3", tp["a2"], dy, w) * (tp["a2"] > 0)
        da1 = lin_bwd(g, p + "linear_2", tp["a1"], da2, w) * (tp["a1"] > 0)
        dn2 = dy + lin_bwd(g, p + "linear_1", tp["n2"], da1, w)
        dxo = lin_bwd(g, T_ + f"post_tfmr_{b}", tp["xo"], dn2, w)
        dx = tfmr_bwd(g, w, T_ + f"seq_tfmr_{b}.", tp["tf"], dxo)
        dn1 = dn2 + dx[..., :C_S]
        dnode0 = dnode0 + lin_bwd(g, T_ + f"skip_embed_{b}", tape["node0"], dx[..., C_S:], w)
        ds_ln = ln_bwd(g, T_ + f"ipa_ln_{b}", tp["ln1"], dn1, w)
        ds_ipa, dz_ipa, dq_ipa, dt_ipa = ipa_bwd(g, w, T_ + f"ipa_{b}.", tp["ipa"], ds_ln * bb[..., None])
        dnode = ds_ln + ds_ipa
        dz = dz_ipa if dz is None else dz + dz_ipa
        dquat = dquat + dq_ipa
        dtrans = dtrans + dt_ipa
        if taps is not None:
            taps[f"dnode_{b}"], taps[f"dz_{b}"], taps[f"dquat_{b}"], taps[f"dtrans_{b}"] = dnode, dz, dquat, dtrans
    dnode0 = dnode0 + dnode
    _mlp3_ln_bwd(g, "embedding_layer.node_embedder.", tape["ne"], dnode0 * bb[..., None], w)
    _mlp3_ln_bwd(g, "embedding_layer.edge_embedder.", tape["ee"], dz * emask[..., None], w)
    return g


# ---------------------------------------------------------------------------------------------------------------------
# loss (Experiment.loss_fn, train_se3_diffusion.py:538-680): value + gradient w.r.t. the model outputs
# ---------------------------------------------------------------------------------------------------------------------
def loss_and_grad(out, batch, exp_conf=None):
    """Returns (total_loss, dout) with dout = d total_loss / d {rot_score, trans_score, rigids, atom37} (float64 where the output is)."""
    c = dict(fo.DEFAULT_EXP_CONF, **(exp_conf or {}))
    T = torch.as_tensor
    bb = T(batch["res_mask"]).double()
    dmk = 1 - T(batch["fixed_mask"]).double()
    lm = bb * dmk
    B, N = bb.shape
    t = T(batch["t"]).double()
    nvalid = torch.any(bb.bool(), dim=-1).sum() + 1e-10
    denom = lm.sum(-1) + 1e-10
    rsc, tsc = T(batch["rot_score_scaling"]).double(), T(batch["trans_score_scaling"]).double()
    gt_rot, gt_trans = T(batch["rot_score"]).double(), T(batch["trans_score"]).double()
    pr = out["rot_score"].double() * dmk[..., None]
    pt = out["trans_score"].double() * dmk[..., None]
    w_s = (1.0 / nvalid)                                     # d total / d batch_loss[b]
    # translation
    e_t = gt_trans - pt
    ts_loss = (e_t ** 2 * lm[..., None]).sum((-1, -2)) / tsc ** 2 / denom
    x0g, x0p = T(batch["rigids_0"])[..., 4:].double() * c["coordinate_scaling"], out["rigids"][..., 4:].double() * c["coordinate_scaling"]
    x0_loss = ((x0g - x0p) ** 2 * lm[..., None]).sum((-1, -2)) / denom
    hi_t = (t > c["trans_x0_threshold"]).double()
    trans_loss = (ts_loss * hi_t + x0_loss * (1 - hi_t)) * c["trans_loss_weight"]
    d_ts = (-2 * e_t * lm[..., None] / (tsc ** 2 * denom)[:, None, None]) * (hi_t * c["trans_loss_weight"] * w_s)[:, None, None] * dmk[..., None]
    d_x0 = (-2 * (x0g - x0p) * lm[..., None] / denom[:, None, None]) * ((1 - hi_t) * c["trans_loss_weight"] * w_s)[:, None, None] * c["coordinate_scaling"]
    # rotation (separate axis / angle)
    ga = torch.norm(gt_rot, dim=-1, keepdim=True)
    gax = gt_rot / (ga + 1e-6)
    pa = torch.norm(pr, dim=-1, keepdim=True)
    pax = pr / (pa + 1e-6)
    axis_loss = ((gax - pax) ** 2 * lm[..., None]).sum((-1, -2)) / denom
    wa = c["rot_loss_weight"] * (t > c["rot_loss_t_threshold"]).double()
    angle_loss = ((ga - pa) ** 2 * lm[..., None]).sum((-1, -2)) / rsc ** 2 / denom * wa
    rot_loss = angle_loss + axis_loss
    d_pax = -2 * (gax - pax) * lm[..., None] / denom[:, None, None] * w_s
    d_pa = -2 * (ga - pa) * lm[..., None] / (rsc ** 2 * denom)[:, None, None] * (wa * w_s)[:, None, None]
    unit = pr / torch.clamp(pa, min=1e-30)
    # pax = pr/(pa+eps): d pr = d_pax/(pa+eps) - (d_pax . pr)/(pa+eps)^2 * unit ; plus d_pa * unit
    d_pr = d_pax / (pa + 1e-6) - (d_pax * pr).sum(-1, keepdim=True) / (pa + 1e-6) ** 2 * unit + d_pa * unit
    d_rs = d_pr * dmk[..., None]
    # backbone atoms
    r0 = T(batch["rigids_0"]).float()
    gt37, gtm, _ = fo.compute_backbone(fo.quat_to_rotmat(r0[..., :4]), r0[..., 4:], T(batch["torsion_angles_sin_cos"])[..., 2, :])
    ga5, m5 = gt37[:, :, :5].double(), gtm[:, :, :5].double()
    pa5 = out["atom37"][:, :, :5].double()
    bm = m5 * lm[..., None]
    bden = bm.sum((-1, -2)) + 1e-10
    w_bb = c["bb_atom_loss_weight"] * (t < c["bb_atom_loss_t_filter"]).double() * c["aux_loss_weight"]
    bb_loss = ((pa5 - ga5) ** 2 * bm[..., None]).sum((-1, -2, -3)) / bden * w_bb
    d_a = 2 * (pa5 - ga5) * bm[..., None] / bden[:, None, None, None] * (w_bb * w_s)[:, None, None, None]
    # pairwise distances
    gf, pf = ga5.reshape(B, N * 5, 3), pa5.reshape(B, N * 5, 3)
    fl = lm[:, :, None].expand(B, N, 5).reshape(B, N * 5)
    fr = bb[:, :, None].expand(B, N, 5).reshape(B, N * 5)
    gd = torch.linalg.norm(gf[:, :, None] - gf[:, None], dim=-1) * fl[..., None]
    diff = pf[:, :, None] - pf[:, None]
    pd_raw = torch.linalg.norm(diff, dim=-1)
    pd = pd_raw * fl[..., None]
    pm = fl[..., None] * fr[:, None, :] * (gd < 6)
    pden = pm.sum((1, 2)) - N
    w_dm = c["dist_mat_loss_weight"] * (t < c["dist_mat_loss_t_filter"]).double() * c["aux_loss_weight"]
    dm_loss = ((gd - pd) ** 2 * pm).sum((1, 2)) / pden * w_dm
    d_pd = -2 * (gd - pd) * pm / pden[:, None, None] * (w_dm * w_s)[:, None, None] * fl[..., None]          # wrt pd_raw
    coef = d_pd / torch.clamp(pd_raw, min=1e-30)
    coef = torch.where(pd_raw > 0, coef, torch.zeros_like(coef))
    d_pf = (coef[..., None] * diff).sum(2) - (coef[..., None] * diff).sum(1)
    d_a = d_a + d_pf.reshape(B, N, 5, 3)
    total = ((rot_loss + trans_loss + bb_loss + dm_loss).sum()) / nvalid
    d_atom37 = torch.zeros(B, N, 37, 3, dtype=torch.float64)
    d_atom37[:, :, :5] = d_a
    d_rig = torch.zeros(B, N, 7, dtype=torch.float64)
    d_rig[..., 4:] = d_x0
    return total, {"rot_score": d_rs, "trans_score": d_ts, "rigids": d_rig.float(), "atom37": d_atom37.float()}
