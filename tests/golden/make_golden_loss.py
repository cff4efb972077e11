"""Golden vectors for the DSM training loss (Experiment.loss_fn, experiments/train_se3_diffusion.py:524-693), produced by the
UNMODIFIED reference in the build container:  python tests/golden/make_golden_loss.py
A synthetic batch is noised with the reference's own SE3Diffuser.forward_marginal (as data/pdb_data_loader.py:251-272 does), pushed
through the reference ScoreNetwork (synthetic weights) inside the reference's loss_fn; batch, model outputs and every loss term are stored."""
import collections
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_harness as rh  # noqa: E402

rh.install_stubs()
from oracle import framediff_oracle as fo  # noqa: E402  (synthetic weights + prior sampler only)
from openfold.utils import rigid_utils as ru  # noqa: E402

EXP_CONF = dict(trans_loss_weight=1.0, rot_loss_weight=0.5, rot_loss_t_threshold=0.2, separate_rot_loss=True, trans_x0_threshold=1.0,
                coordinate_scaling=0.1, bb_atom_loss_weight=1.0, bb_atom_loss_t_filter=0.25, dist_mat_loss_weight=1.0,
                dist_mat_loss_t_filter=0.25, aux_loss_weight=0.25)


def make_batch(diffuser, B, N, ts, seed):
    np.random.seed(seed)
    feats = collections.defaultdict(list)
    for b in range(B):
        r0 = fo.sample_ref(N).numpy().astype(np.float64)
        r0[:, 4:] *= 0.35                                      # protein-like extent so some atom pairs fall under 6 A
        res_mask = np.ones(N); fixed_mask = np.zeros(N)
        if b == 1:
            res_mask[N - 5:] = 0.0
            fixed_mask[3:9] = 1.0
        gt = ru.Rigid.from_tensor_7(torch.tensor(r0))
        d = diffuser.forward_marginal(rigids_0=gt, t=float(ts[b]), diffuse_mask=(1 - fixed_mask) * res_mask)
        tors = np.random.randn(N, 7, 2); tors /= np.linalg.norm(tors, axis=-1, keepdims=True)
        feats["rigids_0"].append(r0); feats["rigids_t"].append(np.asarray(d["rigids_t"]))
        feats["rot_score"].append(np.asarray(d["rot_score"])); feats["trans_score"].append(np.asarray(d["trans_score"]))
        feats["rot_score_scaling"].append(d["rot_score_scaling"]); feats["trans_score_scaling"].append(d["trans_score_scaling"])
        feats["res_mask"].append(res_mask); feats["fixed_mask"].append(fixed_mask); feats["seq_idx"].append(np.arange(1, N + 1) * res_mask.astype(int))
        feats["torsion_angles_sin_cos"].append(tors); feats["sc_ca_t"].append(np.zeros((N, 3))); feats["t"].append(float(ts[b]))
    out = {k: torch.tensor(np.stack(v)) for k, v in feats.items()}
    out["rigids_t"] = out["rigids_t"].float(); out["seq_idx"] = out["seq_idx"].long()
    return out


def main():
    net, diffuser = rh.build_reference(fo.synthetic_weights(0))
    exp_cls = rh.load_experiment_class()
    model_conf, diff_conf = rh.default_conf()

    class _Exp:
        pass

    ex = _Exp()
    ex._model_conf = model_conf; ex._diff_conf = diff_conf; ex._exp_conf = rh.to_attr(EXP_CONF)
    ex.model = ex._model = net; ex.diffuser = ex._diffuser = diffuser
    ex._aux_data_history = collections.deque(maxlen=4)
    for name in ("loss_fn", "_self_conditioning", "_set_t_feats"):
        setattr(ex, name, types.MethodType(getattr(exp_cls, name), ex))
    for tag, (B, N, ts, seed, rnd) in {"a": (2, 24, [0.6, 0.15], 11, 0), "b": (3, 40, [0.9, 0.22, 0.05], 12, 1)}.items():
        batch = make_batch(diffuser, B, N, ts, seed)
        random.seed(rnd)            # decides the self-conditioning coin flip inside loss_fn
        with torch.no_grad():
            loss, aux = ex.loss_fn(dict(batch))
        hist = ex._aux_data_history[-1]
        used = hist["batch"]      # includes sc_ca_t if self-conditioning fired
        mo = hist["model_out"]
        save = {"in_" + k: v.numpy() for k, v in used.items()}
        save.update({"out_" + k: v.detach().numpy() for k, v in mo.items() if torch.is_tensor(v)})
        save.update({"aux_" + k: np.asarray(v.detach().numpy() if torch.is_tensor(v) else v) for k, v in aux.items()})
        save["loss"] = np.asarray(loss.detach().numpy())
        # gradients of the normalised loss w.r.t. every parameter (eval-mode module, the semantics this repo implements; same coin flip):
        # the backward KAT for the next round's kernels — all 282 gradient norms and three full gradients
        random.seed(rnd)
        net.zero_grad(set_to_none=True)
        loss_g, _ = ex.loss_fn(dict(batch))
        loss_g.backward()
        names = [n for n, _ in net.named_parameters()]
        save["grad_names"] = np.array(names)
        save["grad_norms"] = np.array([-1.0 if p_.grad is None else float(p_.grad.double().norm()) for _, p_ in net.named_parameters()])
        for n_ in ("score_model.trunk.ipa_0.linear_b.weight", "score_model.trunk.edge_transition_0.final_layer.weight",
                   "embedding_layer.edge_embedder.0.bias"):
            save["grad::" + n_] = dict(net.named_parameters())[n_].grad.numpy()
        np.savez_compressed(os.path.join(HERE, f"loss_{tag}.npz"), **save)
        print(tag, float(loss), {k: float(np.asarray(v).sum()) for k, v in aux.items() if k.startswith("batch_")})


if __name__ == "__main__":
    main()
