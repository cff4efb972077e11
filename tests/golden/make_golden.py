"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run once in the build container:   python tests/golden/make_golden.py
The GPU box has no /root/reference; tests there read only the committed .npz files.

Weights: `oracle.framediff_oracle.synthetic_weights(seed)` (deterministic MT19937 stream, regenerated at test time,
never stored) and — for the `*_paper` vectors — weights/paper_weights.pth of the reference (tests that need them are
skipped when no copy of the checkpoint is available).

Every vector stores its inputs, the RNG seed used for np.random (the reference draws all diffusion noise from the
global numpy RNG) and the reference outputs.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_harness as rh  # noqa: E402
from oracle import framediff_oracle as fo  # noqa: E402  (only for synthetic_weights/init-feats helpers)


def _np(x):
    return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)


def make_feats(dif, B, N, seed, t, pad_tail=0, fixed=None, sc_scale=8.0):
    np.random.seed(seed)
    r7 = torch.stack([dif.sample_ref(n_samples=N, as_tensor_7=True)["rigids_t"] for _ in range(B)]).float()
    f = {
        "res_mask": torch.ones(B, N, dtype=torch.float64),
        "seq_idx": torch.arange(1, N + 1)[None].repeat(B, 1),
        "fixed_mask": torch.zeros(B, N, dtype=torch.float64),
        "torsion_angles_sin_cos": torch.tensor(np.random.randn(B, N, 7, 2)),
        "sc_ca_t": torch.tensor(np.random.randn(B, N, 3) * sc_scale),
        "rigids_t": r7,
        "t": torch.tensor(t, dtype=torch.float64),
    }
    if pad_tail:
        f["res_mask"][B - 1, N - pad_tail:] = 0
        f["seq_idx"][B - 1, N - pad_tail:] = 0
    if fixed is not None:
        f["fixed_mask"][:, fixed[0]:fixed[1]] = 1
    return f


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: _np(v) for k, v in arrs.items()})
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


def forward_vectors(net, dif, tag, cases):
    for i, (B, N, seed, t, pad, fixed) in enumerate(cases):
        f = make_feats(dif, B, N, seed, t, pad, fixed)
        with torch.no_grad():
            out = net({k: v.clone() for k, v in f.items()})
        save(f"forward_{tag}_{i}", **{"in_" + k: v for k, v in f.items()}, **{"out_" + k: v for k, v in out.items()})


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    w_syn = fo.synthetic_weights(0)
    net, dif = rh.build_reference(w_syn)

    # 1. ScoreNetwork.forward, synthetic weights ------------------------------------------------------------
    forward_vectors(net, dif, "synth", [
        (2, 24, 1, [0.7, 0.45], 0, None),          # plain
        (2, 40, 2, [1.0, 0.55], 5, (3, 7)),        # padded tail + fixed (motif) residues
        (1, 64, 3, [0.85], 0, None),               # N multiple of tile sizes
        (3, 17, 4, [0.9, 0.6, 0.5], 0, None),      # ragged N (not a multiple of anything)
    ])

    # 2. IGSO(3) score known-answer grid: SO3Diffuser.torch_score ----------------------------------------------
    rs = np.random.RandomState(5)
    ts = np.array([0.01, 0.05, 0.1, 0.3, 0.5, 0.8, 1.0])
    om = np.concatenate([np.array([1e-4, 1e-3, 1e-2]), np.linspace(0.05, np.pi - 1e-3, 29)])
    axis = rs.standard_normal((len(ts), len(om), 3))
    axis /= np.linalg.norm(axis, axis=-1, keepdims=True)
    vec = (axis * om[None, :, None]).astype(np.float32)
    sc = dif._so3_diffuser.torch_score(torch.tensor(vec), torch.tensor(ts))
    dif._so3_diffuser.use_cached_score = True
    sc_cached = dif._so3_diffuser.torch_score(torch.tensor(vec), torch.tensor(ts))
    dif._so3_diffuser.use_cached_score = False
    save("igso3_score", t=ts, vec=vec, score=sc, score_cached=sc_cached,
         sigma=dif._so3_diffuser.discrete_sigma[dif._so3_diffuser.t_to_idx(ts)],
         sigma_idx=dif._so3_diffuser.t_to_idx(ts))

    # 3. schedules / scalings ------------------------------------------------------------------------------------
    tt = np.linspace(0.01, 1.0, 500)
    rot_sc = np.array([dif.score_scaling(t)[0] for t in tt])
    trans_sc = np.array([dif.score_scaling(t)[1] for t in tt])
    save("schedules", t=tt, rot_score_scaling=rot_sc, trans_score_scaling=trans_sc,
         so3_g=np.array([dif._so3_diffuser.diffusion_coef(t) for t in tt]),
         so3_sigma_idx=np.array([dif._so3_diffuser.t_to_idx(t) for t in tt]),
         r3_b=np.array([dif._r3_diffuser.b_t(t) for t in tt]),
         cdf_t1=dif._so3_diffuser._cdf[dif._so3_diffuser.t_to_idx(1.0)],
         cdf_t03=dif._so3_diffuser._cdf[dif._so3_diffuser.t_to_idx(0.3)])

    # 4. sample_ref ----------------------------------------------------------------------------------------------
    np.random.seed(11)
    sr = dif.sample_ref(n_samples=48, as_tensor_7=True)["rigids_t"]
    save("sample_ref", seed=11, n=48, rigids_t=sr)

    # 5. SE3Diffuser.reverse, one step, several (t, mask, center, noise_scale) settings -----------------------------
    from openfold.utils import rigid_utils as ru
    B, N = 2, 21
    np.random.seed(21)
    r7 = torch.stack([dif.sample_ref(n_samples=N, as_tensor_7=True)["rigids_t"] for _ in range(B)]).float()
    rot_score = np.random.randn(B, N, 3) * 0.7
    trans_score = np.random.randn(B, N, 3) * 1.3
    mask = np.ones((B, N))
    mask[0, 4:9] = 0
    rev = {}
    for j, (t, use_mask, center, ns) in enumerate([(0.9, False, True, 1.0), (0.31, True, True, 0.1),
                                                   (0.02, True, False, 0.5)]):
        np.random.seed(100 + j)
        out = dif.reverse(rigid_t=ru.Rigid.from_tensor_7(r7), rot_score=rot_score, trans_score=trans_score,
                          t=t, dt=1 / 100, diffuse_mask=mask if use_mask else None, center=center, noise_scale=ns)
        rev[f"rot_{j}"] = out.get_rots().get_rot_mats()
        rev[f"trans_{j}"] = out.get_trans()
        rev[f"cfg_{j}"] = np.array([t, float(use_mask), float(center), ns, 100 + j])
    save("reverse_step", rigids_t=r7, rot_score=rot_score, trans_score=trans_score, mask=mask, dt=1 / 100, **rev)

    # 6. forward_marginal + score ---------------------------------------------------------------------------------
    np.random.seed(31)
    r0 = dif.sample_ref(n_samples=19, as_tensor_7=True)["rigids_t"].float()
    dm = np.ones(19)
    dm[5:8] = 0
    fm = {}
    for j, (t, use_mask) in enumerate([(0.5, False), (0.07, True)]):
        np.random.seed(200 + j)
        o = dif.forward_marginal(ru.Rigid.from_tensor_7(r0), t=t, diffuse_mask=dm if use_mask else None,
                                 as_tensor_7=False)
        fm[f"rot_{j}"] = o["rigids_t"].get_rots().get_rot_mats()
        fm[f"trans_{j}"] = o["rigids_t"].get_trans()
        fm[f"trans_score_{j}"] = o["trans_score"]
        fm[f"rot_score_{j}"] = o["rot_score"]
        fm[f"scal_{j}"] = np.array([o["trans_score_scaling"], o["rot_score_scaling"]])
        fm[f"cfg_{j}"] = np.array([t, float(use_mask), 200 + j])
    np.random.seed(41)
    r1 = dif.sample_ref(n_samples=19, as_tensor_7=True)["rigids_t"].float()
    ts_, rs_ = dif.score(ru.Rigid.from_tensor_7(r0), ru.Rigid.from_tensor_7(r1), 0.4)
    save("forward_marginal", rigids_0=r0, rigids_1=r1, mask=dm, score_trans=ts_, score_rot=rs_, score_t=0.4, **fm)

    # 7. trajectory through the reference's real Experiment.inference_fn, synthetic weights ------------------------
    ex = rh.make_experiment(net, dif)

    def run_traj(N, B, num_t, seed, noise_scale):
        np.random.seed(seed)
        r7 = torch.stack([dif.sample_ref(n_samples=N, as_tensor_7=True)["rigids_t"] for _ in range(B)])
        init = {
            "res_mask": torch.ones(B, N, dtype=torch.float64), "seq_idx": torch.arange(1, N + 1)[None].repeat(B, 1),
            "fixed_mask": torch.zeros(B, N, dtype=torch.float64),
            "torsion_angles_sin_cos": torch.zeros(B, N, 7, 2, dtype=torch.float64),
            "sc_ca_t": torch.zeros(B, N, 3, dtype=torch.float64), "rigids_t": r7,
        }
        out = ex.inference_fn(init, num_t=num_t, min_t=0.01, aux_traj=True, noise_scale=noise_scale)
        return r7, out

    r7, out = run_traj(28, 2, 12, 77, 0.1)
    save("traj_synth", seed=77, N=28, B=2, num_t=12, noise_scale=0.1, rigids_init=r7,
         prot_traj=out["prot_traj"], rigid_traj=out["rigid_traj"], trans_traj=out["trans_traj"],
         psi_pred=out["psi_pred"], rigid_0_traj=out["rigid_0_traj"])

    # 8. paper weights: SURVEY §8(c) KATs + config 1 (1 x N=60 x 50 steps) -----------------------------------------
    sd = rh.load_reference_checkpoint("paper_weights.pth")
    net.load_state_dict(sd, strict=True)
    np.random.seed(0)
    r7 = dif.sample_ref(60, as_tensor_7=True)["rigids_t"][None]
    f = {"res_mask": torch.ones(1, 60, dtype=torch.float64), "fixed_mask": torch.zeros(1, 60, dtype=torch.float64),
         "torsion_angles_sin_cos": torch.zeros(1, 60, 7, 2, dtype=torch.float64),
         "sc_ca_t": torch.zeros(1, 60, 3, dtype=torch.float64), "seq_idx": torch.arange(1, 61)[None],
         "t": torch.tensor([1.0], dtype=torch.float64), "rigids_t": r7}
    with torch.no_grad():
        o = net(f)
    print("KAT rot_score[0,0]", o["rot_score"][0, 0].numpy(), "trans_score[0,0]", o["trans_score"][0, 0].numpy())
    save("forward_paper_0", **{"in_" + k: v for k, v in f.items()}, **{"out_" + k: v for k, v in o.items()})
    np.random.seed(123)
    torch.manual_seed(123)
    r7 = dif.sample_ref(60, as_tensor_7=True)["rigids_t"][None]
    init = {k: v for k, v in f.items() if k != "t"}
    init["rigids_t"] = r7
    out = ex.inference_fn(init, num_t=50, min_t=0.01, aux_traj=True, noise_scale=0.1)
    save("traj_paper_c1", seed=123, N=60, B=1, num_t=50, noise_scale=0.1, rigids_init=r7,
         prot_final=out["prot_traj"][0], rigid_traj=out["rigid_traj"][::10], psi_pred=out["psi_pred"],
         trans_final=out["trans_traj"][0])


if __name__ == "__main__":
    main()
