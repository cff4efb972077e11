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
    return 0 if res["ok"] else 1


if __name__ == "__main__":
    sys.exit(main())
