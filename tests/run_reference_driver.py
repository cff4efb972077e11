"""Runs the UNMODIFIED reference driver code (`Experiment.inference_fn`, experiments/train_se3_diffusion.py:718-818 — the function
`inference_se3_diffusion.py`'s Sampler calls) with this repository's overlay first on sys.path, on a CUDA device, and compares the result
with the trajectory the unmodified reference produced on CPU (tests/golden/traj_paper_c1.npz: paper weights, N=60, 50 steps,
np.random.seed(123)).  Needs the reference tree (FRAMEDIFF_REFERENCE) — it is not part of this repository and is shipped to the GPU
box only as scratch for this run.  Prints one JSON line.

    FRAMEDIFF_REFERENCE=/path/to/se3_diffusion python tests/run_reference_driver.py
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, ROOT)


def main():
    import torch
    import ref_harness as rh
    rh.install_stubs()                                                   # stubs + reference root on sys.path
    sys.path.insert(0, os.path.join(ROOT, "se3_diffusion_b200", "overlay"))   # ... and the overlay BEFORE it
    from data import all_atom, se3_diffuser                              # noqa: E402
    from model import ipa_pytorch, score_network                         # noqa: E402
    from experiments import train_se3_diffusion as tsd                   # noqa: E402
    assert "se3_diffusion_b200/overlay" in score_network.__file__ and "se3_diffusion_b200/overlay" in se3_diffuser.__file__
    assert ipa_pytorch.__file__.startswith(rh.REFERENCE_ROOT) and all_atom.__file__.startswith(rh.REFERENCE_ROOT)
    assert tsd.__file__.startswith(rh.REFERENCE_ROOT)
    mc, dc = rh.default_conf()
    dif = se3_diffuser.SE3Diffuser(dc)
    net = score_network.ScoreNetwork(mc, dif, precision=os.environ.get("FD_PRECISION", "bf16x3"))
    wpath = os.path.join(ROOT, "weights", "paper_weights.npz")
    net.load_state_dict({k: torch.tensor(v) for k, v in np.load(wpath).items()}, strict=True)
    device = torch.device("cuda:0")
    net = net.to(device).eval()
    ex = rh.make_experiment(net, dif)                                    # the reference's own unbound inference_fn / _set_t_feats / _self_conditioning
    g = dict(np.load(os.path.join(HERE, "golden", "traj_paper_c1.npz")))
    np.random.seed(int(g["seed"])); torch.manual_seed(int(g["seed"]))
    N = int(g["N"])
    r7 = dif.sample_ref(N, as_tensor_7=True)["rigids_t"][None]
    prior_err = float(np.abs(r7.numpy()[..., 4:] - g["rigids_init"][..., 4:]).max())
    # Sampler.sample's feature dict (experiments/inference_se3_diffusion.py:432-449), moved to the device like Sampler does
    init = {"res_mask": torch.ones(1, N, dtype=torch.float64), "fixed_mask": torch.zeros(1, N, dtype=torch.float64),
            "torsion_angles_sin_cos": torch.zeros(1, N, 7, 2, dtype=torch.float64), "sc_ca_t": torch.zeros(1, N, 3, dtype=torch.float64),
            "seq_idx": torch.arange(1, N + 1)[None], "rigids_t": r7}
    init = {k: v.to(device) for k, v in init.items()}
    t0 = time.perf_counter()
    out = ex.inference_fn(init, num_t=int(g["num_t"]), min_t=0.01, aux_traj=True, noise_scale=float(g["noise_scale"]))
    dt = time.perf_counter() - t0
    final = out["prot_traj"][0]
    err = float(np.abs(final - g["prot_final"]).max()); scale = float(np.abs(g["prot_final"]).max())
    ca = final[0, :, 1]
    bond = float(np.linalg.norm(ca[1:] - ca[:-1], axis=-1).mean())
    res = {"driver": "Experiment.inference_fn (unmodified reference) through the overlay", "device": torch.cuda.get_device_name(0),
           "precision": net.precision, "N": N, "num_t": int(g["num_t"]), "seconds": dt, "prior_trans_max_err": prior_err,
           "final_atom37_max_err": err, "final_atom37_rel_err": err / scale, "mean_ca_ca": bond,
           "shapes": {k: list(v.shape) for k, v in out.items()},
           "ok": bool(err <= 2e-3 * scale and abs(bond - 3.8088) < 5e-3 and prior_err < 1e-4)}
    print(json.dumps(res))
    ok_inf = res["ok"]
    # ---- training: the reference's own loss_fn (+ _self_conditioning) and update step on the overlay model ----------------------------
    # Experiment.update_fn (experiments/train_se3_diffusion.py:320-326) is `loss, aux = self.loss_fn(feats); optimizer.zero_grad();
    # loss.backward(); optimizer.step()` — run verbatim on a light object carrying the attributes loss_fn touches; compared with the gradient
    # norms the unmodified reference model produced on the same batch (tests/golden/loss_b.npz).
    import collections, random, types
    from oracle import framediff_oracle as fo
    g = dict(np.load(os.path.join(HERE, "golden", "loss_b.npz"), allow_pickle=False))
    net2 = score_network.ScoreNetwork(mc, dif, precision="fp32")
    net2.load_state_dict({k: torch.tensor(v) for k, v in fo.synthetic_weights(0).items()}, strict=True)
    net2 = net2.to(device)
    net2.eval()                       # the golden was produced on an eval-mode module with autograd recording (make_golden_loss.py)
    exp_conf = dict(trans_loss_weight=1.0, rot_loss_weight=0.5, rot_loss_t_threshold=0.2, separate_rot_loss=True, trans_x0_threshold=1.0,
                    coordinate_scaling=0.1, bb_atom_loss_weight=1.0, bb_atom_loss_t_filter=0.25, dist_mat_loss_weight=1.0, dist_mat_loss_t_filter=0.25,
                    aux_loss_weight=0.25)

    class _Exp:
        pass

    ex2 = _Exp()
    ex2._model_conf = mc; ex2._diff_conf = dc; ex2._exp_conf = rh.to_attr(exp_conf)
    ex2.model = ex2._model = net2; ex2.diffuser = ex2._diffuser = dif
    ex2._aux_data_history = collections.deque(maxlen=4)
    for name in ("loss_fn", "_self_conditioning", "_set_t_feats"):
        setattr(ex2, name, types.MethodType(getattr(tsd.Experiment, name), ex2))
    batch = {k[3:]: torch.as_tensor(v).to(device) for k, v in g.items() if k.startswith("in_") and k != "in_sc_ca_t"}
    batch["sc_ca_t"] = torch.zeros_like(batch["rigids_t"][..., 4:])
    opt = torch.optim.Adam(net2.parameters(), lr=1e-4)            # train_se3_diffusion.py:139-141
    random.seed(1)                                                 # the golden's self-conditioning coin flip
    loss, aux = ex2.loss_fn(batch)
    opt.zero_grad(); loss.backward(); opt.step()
    names, norms = [str(n) for n in g["grad_names"]], g["grad_norms"]
    named = dict(net2.named_parameters())
    worst, n_none = 0.0, 0
    for n, rn in zip(names, norms):
        gr = named[n].grad
        if rn < 0:
            n_none += int(gr is None)
            continue
        worst = max(worst, max(abs(float(gr.double().norm()) - rn) - 1e-6, 0.0) / rn)     # 1e-6 absolute floor: linear_b.bias is 0 analytically
    res2 = {"driver": "Experiment.loss_fn + loss.backward() + Adam step (unmodified reference) on the overlay ScoreNetwork", "loss": float(loss),
            "reference_loss_with_autograd_semantics": None, "worst_grad_norm_rel_err_vs_reference": worst, "params_with_grad_none": n_none,
            "ok": bool(worst < 2e-3 and n_none == 10 and np.isfinite(float(loss)))}
    print(json.dumps(res2))
    return 0 if (ok_inf and res2["ok"]) else 1


if __name__ == "__main__":
    sys.exit(main())
