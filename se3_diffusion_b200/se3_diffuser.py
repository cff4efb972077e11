"""SE3Diffuser — B200 host-side mirror of the reference's data/se3_diffuser.py public API.

Same class name, method names, argument meaning, return types and error behaviour (ValueError on bad t / missing impute)
as /root/reference/data/se3_diffuser.py:31-268, so the reference's drivers can use it unchanged through the overlay module
se3_diffusion_b200/overlay/data/se3_diffuser.py.  All arithmetic runs in libframediff_b200.so on the GPU (IGSO(3) tables and
scores, reverse step, prior sample, forward noising); numpy is used only where the reference itself defines the contract:
the noise comes from the GLOBAL numpy RNG in the reference's draw order, and results cross the API as numpy arrays / Rigid.

Rigid values: when `openfold.utils.rigid_utils` is importable (i.e. inside the reference tree) methods accept and return
its Rigid objects exactly like the reference; otherwise they accept/return [..., 7] tensors (qw,qx,qy,qz,tx,ty,tz).
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from .engine import FrameDiffEngine

try:  # the openfold fork lives in the reference tree; optional here
    from openfold.utils import rigid_utils as ru  # type: ignore
except Exception:  # pragma: no cover
    ru = None


def _to_tensor7(rigid):
    if torch.is_tensor(rigid):
        return rigid
    if ru is not None and isinstance(rigid, ru.Rigid):
        # rot-mat backed rigids go through rot_to_quat (eigh) in the reference; any quaternion of the rotation is equivalent
        return rigid.to_tensor_7()
    raise TypeError(f"expected a Rigid or a [...,7] tensor, got {type(rigid)}")


def _from_tensor7(t7: torch.Tensor, rotmat: torch.Tensor | None = None):
    if ru is None:
        return t7
    if rotmat is not None:   # like _assemble_rigid: rotation-matrix backed, CPU
        return ru.Rigid(rots=ru.Rotation(rot_mats=rotmat.cpu()), trans=t7[..., 4:].cpu())
    return ru.Rigid.from_tensor_7(t7.cpu())


class _SO3Facade:
    """The few SO3Diffuser attributes drivers poke at (se3_diffuser.py:235, train_se3_diffusion.py:701-706)."""

    def __init__(self, parent, conf):
        self._p = parent
        self.min_sigma, self.max_sigma, self.num_sigma = conf.min_sigma, conf.max_sigma, conf.num_sigma
        self.schedule, self.use_cached_score = conf.schedule, conf.use_cached_score
        if (conf.min_sigma, conf.max_sigma, conf.num_sigma, conf.num_omega, conf.schedule) != (0.1, 1.5, 1000, 1000, "logarithmic"):
            raise ValueError("the B200 kernels are compiled for the shipped IGSO(3) configuration (config/base.yaml:35-43)")

    def sigma(self, t):
        t = np.asarray(t, dtype=np.float64)
        if np.any(t < 0) or np.any(t > 1):
            raise ValueError(f"Invalid t={t}")
        return np.log(t * np.exp(self.max_sigma) + (1 - t) * np.exp(self.min_sigma))

    def diffusion_coef(self, t):
        s = self.sigma(t)
        return np.sqrt(2 * (np.exp(self.max_sigma) - np.exp(self.min_sigma)) * s / np.exp(s))

    def score_scaling(self, t):
        return self._p._engine().score_scaling(float(t))[0]


class _R3Facade:
    def __init__(self, conf):
        self._r3_conf = conf
        self.min_b, self.max_b = conf.min_b, conf.max_b
        if (conf.min_b, conf.max_b, conf.coordinate_scaling) != (0.1, 20.0, 0.1):
            raise ValueError("the B200 kernels are compiled for the shipped R3 schedule (config/base.yaml:28-32)")

    def _scale(self, x):
        return x * self._r3_conf.coordinate_scaling

    def _unscale(self, x):
        return x / self._r3_conf.coordinate_scaling

    def b_t(self, t):
        if np.any(t < 0) or np.any(t > 1):
            raise ValueError(f"Invalid t={t}")
        return self.min_b + t * (self.max_b - self.min_b)

    def marginal_b_t(self, t):
        return t * self.min_b + 0.5 * (t ** 2) * (self.max_b - self.min_b)

    def score_scaling(self, t):
        return 1 / np.sqrt(1 - np.exp(-self.marginal_b_t(t)))


class SE3Diffuser:
    def __init__(self, se3_conf):
        self._log = logging.getLogger(__name__)
        self._se3_conf = se3_conf
        self._diffuse_rot = se3_conf.diffuse_rot
        self._diffuse_trans = se3_conf.diffuse_trans
        if not (self._diffuse_rot and self._diffuse_trans):
            raise ValueError("the B200 path implements the shipped configuration diffuse_rot = diffuse_trans = True")
        self._so3_diffuser = _SO3Facade(self, se3_conf.so3)
        self._r3_diffuser = _R3Facade(se3_conf.r3)
        self._eng = None
        self._device = None

    # the engine is created lazily so that constructing the diffuser (e.g. in a DataLoader worker) needs no CUDA context
    def _engine(self) -> FrameDiffEngine:
        if self._eng is None:
            self._eng = FrameDiffEngine(self._device if self._device is not None else torch.cuda.current_device())
            self._eng.use_cached_score = bool(self._so3_diffuser.use_cached_score)
        return self._eng

    def bind_engine(self, engine: FrameDiffEngine):
        """Share the ScoreNetwork's engine (one handle per process and device)."""
        self._eng = engine
        engine.use_cached_score = bool(self._so3_diffuser.use_cached_score)     # the network's rotation-score head follows the diffuser's switch

    # ---- noising (training data) -------------------------------------------------------------------------------------
    def forward_marginal(self, rigids_0, t: float, diffuse_mask: np.ndarray = None, as_tensor_7: bool = True):
        if not np.isscalar(t):
            raise ValueError(f"{t} must be a scalar.")
        r0 = _to_tensor7(rigids_0).float()
        shp = r0.shape[:-1]
        n = int(np.prod(shp))
        # np.random draw order of the reference: SO3Diffuser.sample -> randn(n,3), rand(n); R3Diffuser.forward_marginal -> normal
        z_axis, u = np.random.randn(n, 3), np.random.rand(n)
        z_trans = np.random.normal(size=(n, 3))
        out = self._engine().forward_marginal(r0.reshape(n, 7), float(t), z_axis, u, z_trans,
                                              None if diffuse_mask is None else np.asarray(diffuse_mask).reshape(n))
        rt = out["rigids_t"].reshape(*shp, 7).cpu()
        return {
            "rigids_t": rt if as_tensor_7 else _from_tensor7(rt),
            "trans_score": out["trans_score"].reshape(*shp, 3).cpu().numpy(),
            "rot_score": out["rot_score"].reshape(*shp, 3).cpu().numpy(),
            "trans_score_scaling": out["trans_score_scaling"],
            "rot_score_scaling": out["rot_score_scaling"],
        }

    def calc_trans_0(self, trans_score, trans_t, t):
        beta_t = self._r3_diffuser.marginal_b_t(t)[..., None, None]
        return (trans_score * (1 - torch.exp(-beta_t)) + trans_t) / torch.exp(-0.5 * beta_t)

    def calc_trans_score(self, trans_t, trans_0, t, use_torch=False, scale=True):
        exp_fn = torch.exp if use_torch else np.exp
        if scale:
            trans_t, trans_0 = self._r3_diffuser._scale(trans_t), self._r3_diffuser._scale(trans_0)
        b = self._r3_diffuser.marginal_b_t(t)
        return -(trans_t - exp_fn(-0.5 * b) * trans_0) / (1 - exp_fn(-b))

    def calc_rot_score(self, rots_t, rots_0, t):
        """Rotation score of R_t given R_0 (se3_diffuser.py:119-125): quaternion algebra here, IGSO(3) series on the GPU."""
        q0 = rots_0.get_quats() if hasattr(rots_0, "get_quats") else rots_0
        qt = rots_t.get_quats() if hasattr(rots_t, "get_quats") else rots_t
        inv = torch.cat([q0[..., :1], -q0[..., 1:]], -1) / torch.sum(q0 ** 2, dim=-1, keepdim=True)
        a1, b1, c1, d1 = inv.unbind(-1)
        a2, b2, c2, d2 = qt.unbind(-1)
        q = torch.stack([a1 * a2 - b1 * b2 - c1 * c2 - d1 * d2, a1 * b2 + b1 * a2 + c1 * d2 - d1 * c2,
                         a1 * c2 - b1 * d2 + c1 * a2 + d1 * b2, a1 * d2 + b1 * c2 - c1 * b2 + d1 * a2], -1)
        flip = (q[..., :1] < 0).float()
        q = (-1 * q) * flip + (1 - flip) * q
        angle = 2 * torch.atan2(torch.linalg.norm(q[..., 1:], dim=-1), q[..., 0])
        a2_ = angle * angle
        scale = torch.where(angle <= 1e-3, 2 + a2_ / 12 + 7 * a2_ * a2_ / 2880, angle / torch.sin(angle / 2 + 1e-6))
        rotvec = scale[..., None] * q[..., 1:]
        t_np = torch.as_tensor(t).detach().cpu().numpy().astype(np.float64).reshape(-1)
        sig = self._so3_diffuser.sigma(t_np)
        grid = self._so3_diffuser.sigma(np.linspace(0.0, 1.0, 1000))
        sig_q = grid[np.digitize(sig, grid) - 1]
        if self._so3_diffuser.use_cached_score:
            # so3_diffuser.py:291-298: bucketize the angle on the omega grid and gather from the precomputed score-norm rows (built on the GPU)
            idx = np.digitize(sig, grid) - 1
            rows = torch.tensor(self._engine().igso3_tables([int(i) for i in idx])["score_norms"]).to(rotvec.device)     # [B,1000]
            omega = torch.linalg.norm(rotvec, dim=-1) + 1e-6
            om_grid = torch.tensor(np.linspace(0, np.pi, 1001)[1:-1]).to(rotvec.device)
            oi = torch.bucketize(omega, om_grid)
            sc = torch.gather(rows, 1, oi.reshape(rows.shape[0], -1)).reshape(omega.shape)
            return sc[..., None] * rotvec / (omega[..., None] + 1e-6)
        sigma = torch.tensor(sig_q).reshape(-1, *([1] * (rotvec.ndim - 2))).expand(rotvec.shape[:-1])
        return self._engine().igso3_score(rotvec, sigma).to(rotvec.device)

    def _apply_mask(self, x_diff, x_fixed, diff_mask):
        return diff_mask * x_diff + (1 - diff_mask) * x_fixed

    def score(self, rigid_0, rigid_t, t: float):
        """se3_diffuser.py:134-153 — scores rot_t itself; translations unscaled (the reference's quirk, kept)."""
        r0, rt = _to_tensor7(rigid_0).float(), _to_tensor7(rigid_t).float()
        q = rt[..., :4] / torch.linalg.norm(rt[..., :4], dim=-1, keepdim=True)
        flip = (q[..., :1] < 0).float()
        q = (-1 * q) * flip + (1 - flip) * q
        vn = torch.linalg.norm(q[..., 1:], dim=-1)
        ang = 2 * torch.atan2(vn, q[..., 0])
        rotvec = q[..., 1:] * (ang / vn.clamp_min(1e-30))[..., None]
        sig = self._so3_diffuser.sigma(float(t))
        grid = self._so3_diffuser.sigma(np.linspace(0.0, 1.0, 1000))
        sig_q = float(grid[np.digitize(sig, grid) - 1])
        rot_score = self._engine().igso3_score(rotvec, torch.full(rotvec.shape[:-1], sig_q, dtype=torch.float64)).cpu().numpy()
        if rot_score.ndim == 2:
            # the reference's SO3Diffuser.score calls torch_score(vec, tensor(t)[None]): its [1,1] sigma broadcasts against the [N] angles,
            # so an [N,7] input comes back as [1,N,3] (so3_diffuser.py:268-272) — kept
            rot_score = rot_score[None]
        trans_score = self.calc_trans_score(rt[..., 4:].cpu().numpy(), r0[..., 4:].cpu().numpy(), t, scale=False)
        return trans_score, rot_score

    def score_scaling(self, t):
        return self._engine().score_scaling(float(t))

    # ---- reverse step ------------------------------------------------------------------------------------------------------
    def reverse(self, rigid_t, rot_score: np.ndarray, trans_score: np.ndarray, t: float, dt: float, diffuse_mask: np.ndarray = None,
                center: bool = True, noise_scale: float = 1.0):
        if not np.isscalar(t):
            raise ValueError(f"{t} must be a scalar.")
        r = _to_tensor7(rigid_t).float()
        lead = r.shape[:-1]
        r3 = r.reshape((1,) + tuple(lead) + (7,)) if r.ndim == 2 else r
        B, N = r3.shape[0], r3.shape[1]
        shape3 = (B, N, 3)
        # np.random draw order: SO3Diffuser.reverse then R3Diffuser.reverse (both `size=score_t.shape`)
        z_rot = np.random.normal(size=np.asarray(rot_score).shape).reshape(shape3)
        z_trans = np.random.normal(size=np.asarray(trans_score).shape).reshape(shape3)
        dm = None if diffuse_mask is None else np.asarray(diffuse_mask, dtype=np.float32).reshape(B, N)
        out, rm = self._engine().reverse_step(r3, np.asarray(rot_score).reshape(shape3), np.asarray(trans_score).reshape(shape3), float(t),
                                              float(dt), diffuse_mask=dm, center=center, noise_scale=noise_scale, z_rot=z_rot,
                                              z_trans=z_trans, want_rotmat=True)
        out, rm = out.reshape(*lead, 7), rm.reshape(*lead, 3, 3)
        return _from_tensor7(out, rm)

    # ---- prior ---------------------------------------------------------------------------------------------------------------
    def sample_ref(self, n_samples: int, impute=None, diffuse_mask: np.ndarray = None, as_tensor_7: bool = False):
        if diffuse_mask is not None and impute is None:
            raise ValueError("Must provide imputation values.")
        if (not self._diffuse_rot) and impute is None:
            raise ValueError("Must provide imputation values.")
        if (not self._diffuse_trans) and impute is None:
            raise ValueError("Must provide imputation values.")
        z_axis, u = np.random.randn(n_samples, 3), np.random.rand(n_samples)      # so3.sample
        z_trans = np.random.normal(size=(n_samples, 3))                           # r3.sample_ref
        r7 = self._engine().sample_ref(n_samples, z_axis, u, z_trans).cpu()
        if diffuse_mask is not None:
            imp = _to_tensor7(impute).float().reshape(n_samples, 7)
            m = torch.as_tensor(np.asarray(diffuse_mask, dtype=np.float32)).reshape(n_samples, 1)
            # rotations are mixed as rotation vectors in the reference; with a 0/1 mask this is a selection
            r7 = torch.where(m > 0.5, r7, imp)
        return {"rigids_t": r7 if as_tensor_7 else _from_tensor7(r7)}
