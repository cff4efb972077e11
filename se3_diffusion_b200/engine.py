"""Host-side driver of the B200 FrameDiff engine: a thin Python layer over the C ABI (include/framediff_b200.h).

PyTorch is used only as the owner of device memory (tensors -> raw pointers) and for streams.  All arithmetic of the
hot path runs in libframediff_b200.so; there is no CPU or eager-PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from ._lib import ForwardIn, ForwardOut, SampleCfg, SampleIn, SampleOut, check

PREC = {"fp32": 0, "bf16x3": 1, "bf16": 2}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _np_ptr(a: Optional[np.ndarray]):
    return None if a is None else C.c_void_p(a.ctypes.data)


def param_schema():
    lib = _lib.load()
    out = []
    for i in range(lib.fd_num_params()):
        nd = lib.fd_param_ndim(i)
        out.append((lib.fd_param_name(i).decode(), tuple(int(lib.fd_param_dim(i, d)) for d in range(nd))))
    return out


class FrameDiffEngine:
    """One engine per (process, CUDA device)."""

    def __init__(self, device: int | torch.device | str = 0, precision: str = "fp32"):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.FrameDiffError("FrameDiffEngine needs a CUDA device (B200); there is no CPU fallback")
        dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if dev.type != "cuda":
            raise _lib.FrameDiffError(f"FrameDiffEngine needs a CUDA device, got {dev}")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        h = C.c_void_p()
        check(self.lib.fd_create(C.byref(h), self.device.index))
        self._h = h
        self.set_precision(precision)
        self._weights_version = None
        self.validate_t = False
        self.use_cached_score = False   # SO3Diffuser.use_cached_score (config/base.yaml: False): table lookup instead of the IGSO(3) series
        self.last_gpu_ms = None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self.lib.fd_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- configuration -------------------------------------------------------------------------------------------
    def set_precision(self, precision: str):
        check(self.lib.fd_set_precision(self._h, PREC[precision]))
        self.precision = precision

    def load_weights(self, state: Dict[str, "np.ndarray | torch.Tensor"]):
        """state: the reference's 282-entry state_dict (a leading 'module.' is stripped)."""
        state = {(k[7:] if k.startswith("module.") else k): v for k, v in state.items()}
        schema = param_schema()
        keep, ptrs = [], (C.c_void_p * len(schema))()
        for i, (name, shape) in enumerate(schema):
            if name not in state:
                raise KeyError(f"missing parameter {name}")
            v = state[name]
            a = v.detach().to("cpu", torch.float32).contiguous().numpy() if torch.is_tensor(v) else np.ascontiguousarray(v, dtype=np.float32)
            if tuple(a.shape) != shape:
                raise ValueError(f"parameter {name}: shape {tuple(a.shape)} != {shape}")
            keep.append(a)
            ptrs[i] = a.ctypes.data
        check(self.lib.fd_load_weights(self._h, ptrs))

    def set_debug(self, on: bool):
        check(self.lib.fd_set_debug(self._h, int(on)))

    def debug_fetch(self, name: str, shape, dtype=np.float32) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        n = self.lib.fd_debug_fetch(self._h, name.encode(), _np_ptr(out), out.nbytes)
        if n < 0:
            check(int(n))
        if n != out.nbytes:
            raise ValueError(f"tap {name}: {n} bytes, expected {out.nbytes}")
        return out

    # ---- ScoreNetwork.forward ------------------------------------------------------------------------------------------
    def forward(self, feats: Dict[str, torch.Tensor], want_atoms: bool = True) -> Dict[str, torch.Tensor]:
        dev = self.device
        rig = feats["rigids_t"]
        B, N = rig.shape[0], rig.shape[1]
        t_in = torch.as_tensor(feats["t"])
        t_is_f32 = 0 if t_in.dtype == torch.float64 else 1
        f32 = lambda x: torch.as_tensor(x).to(dev, torch.float32).contiguous()
        rigids_t = f32(rig)
        # the reference raises ValueError for t outside [0,1] (so3_diffuser.py:194); host tensors are checked for free, device tensors
        # only on request (validate_t=True) because the check forces a device->host sync per call
        if (t_in.device.type == "cpu" or self.validate_t) and bool(((t_in < 0) | (t_in > 1)).any()):
            raise ValueError(f"Invalid t={t_in}")
        t64 = t_in.to(dev, torch.float64).contiguous()
        res_mask, fixed_mask = f32(feats["res_mask"]), f32(feats["fixed_mask"])
        seq_idx = torch.as_tensor(feats["seq_idx"]).to(dev, torch.int32).contiguous()
        sc_ca = f32(feats["sc_ca_t"])
        tors = feats.get("torsion_angles_sin_cos")
        gt_psi = f32(torch.as_tensor(tors)[..., 2, :]) if tors is not None else None
        out = {
            "rot_score": torch.empty(B, N, 3, device=dev, dtype=torch.float64),
            "trans_score": torch.empty(B, N, 3, device=dev, dtype=torch.float64),
            "psi": torch.empty(B, N, 2, device=dev, dtype=torch.float32),
            "rigids": torch.empty(B, N, 7, device=dev, dtype=torch.float32),
        }
        if want_atoms:
            out["atom37"] = torch.empty(B, N, 37, 3, device=dev, dtype=torch.float32)
            out["atom14"] = torch.empty(B, N, 14, 3, device=dev, dtype=torch.float32)
        rows = None
        if self.use_cached_score:
            # so3_diffuser.py:291-298: rows of the precomputed _score_norms table at each sample's sigma index (t goes to the host, as in the
            # reference's `self.t_to_idx(du.move_to_np(t))`); rows are built on the GPU once per index and kept on the device
            tn = t64.cpu().numpy()
            sig = np.log(tn * np.exp(1.5) + (1 - tn) * np.exp(0.1))
            grid = np.log(np.linspace(0.0, 1.0, 1000) * np.exp(1.5) + (1 - np.linspace(0.0, 1.0, 1000)) * np.exp(0.1))
            idx = np.digitize(sig, grid) - 1
            cache = self.__dict__.setdefault("_score_rows", {})
            miss = sorted({int(i) for i in idx if int(i) not in cache})
            if miss:
                tab = self.igso3_tables(miss)["score_norms"]
                for k, i in enumerate(miss):
                    cache[i] = torch.tensor(tab[k], device=dev)
            rows = torch.stack([cache[int(i)] for i in idx]).contiguous()
        fin = ForwardIn(_ptr(rigids_t), _ptr(t64), t_is_f32, None, _ptr(res_mask), _ptr(fixed_mask), _ptr(seq_idx), _ptr(sc_ca),
                        _ptr(gt_psi), _ptr(rows))
        fout = ForwardOut(_ptr(out["rot_score"]), _ptr(out["trans_score"]), _ptr(out["psi"]), _ptr(out["rigids"]),
                          _ptr(out.get("atom37")), _ptr(out.get("atom14")))
        stream = torch.cuda.current_stream(dev)
        check(self.lib.fd_forward(self._h, B, N, C.byref(fin), C.byref(fout), C.c_void_p(stream.cuda_stream)))
        # dtype semantics of the reference: rot_score is float64; trans_score follows t; psi follows
        # the torsion features it is mixed with.
        if t_is_f32:
            out["trans_score"] = out["trans_score"].to(torch.float32)
        if tors is not None and torch.as_tensor(tors).dtype == torch.float64:
            out["psi"] = out["psi"].to(torch.float64)
        self._keep = (rigids_t, t64, res_mask, fixed_mask, seq_idx, sc_ca, gt_psi, rows)
        return out

    # ---- SE3Diffuser pieces ----------------------------------------------------------------------------------------------
    def igso3_score(self, vec: torch.Tensor, sigma: torch.Tensor) -> torch.Tensor:
        v = vec.to(self.device, torch.float32).contiguous().reshape(-1, 3)
        s = sigma.to(self.device, torch.float64).contiguous().reshape(-1)
        out = torch.empty(v.shape[0], 3, device=self.device, dtype=torch.float64)
        st = torch.cuda.current_stream(self.device)
        check(self.lib.fd_igso3_score(self._h, v.shape[0], _ptr(v), _ptr(s), _ptr(out), C.c_void_p(st.cuda_stream)))
        return out.reshape(vec.shape)

    def igso3_tables(self, sigma_idx):
        idx = np.ascontiguousarray(sigma_idx, dtype=np.int32).reshape(-1)
        n = idx.shape[0]
        pdf, cdf, sn = (np.empty((n, 1000)) for _ in range(3))
        sc = np.empty(n)
        check(self.lib.fd_igso3_tables_host(self._h, n, _np_ptr(idx), _np_ptr(pdf), _np_ptr(cdf), _np_ptr(sn), _np_ptr(sc)))
        return {"pdf": pdf, "cdf": cdf, "score_norms": sn, "score_scaling": sc}

    def sample_ref(self, n: int, z_axis=None, u_angle=None, z_trans=None, seed=0, first_sample=0, per_sample=None):
        dev = self.device
        out = torch.empty(n, 7, device=dev, dtype=torch.float32)
        d = lambda a: None if a is None else torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
        za, ua, zt = d(z_axis), d(u_angle), d(z_trans)
        st = torch.cuda.current_stream(dev)
        check(self.lib.fd_sample_ref(self._h, n, _ptr(za), _ptr(ua), _ptr(zt), seed, first_sample, per_sample or n, _ptr(out),
                                     C.c_void_p(st.cuda_stream)))
        torch.cuda.current_stream(dev).synchronize()
        return out

    def reverse_step(self, rigids: torch.Tensor, rot_score, trans_score, t: float, dt: float, diffuse_mask=None, center=True,
                     noise_scale=1.0, z_rot=None, z_trans=None, seed=0, first_sample=0, step=0, want_rotmat=False):
        """In-place on a copy: returns (rigids_{t-1} [B,N,7] fp32, rotmat or None)."""
        dev = self.device
        B, N = rigids.shape[:2]
        r = rigids.to(dev, torch.float32).contiguous().clone()
        d64 = lambda a: None if a is None else torch.as_tensor(a).to(dev, torch.float64).contiguous()
        rs, ts, zr, zt = d64(rot_score), d64(trans_score), d64(z_rot), d64(z_trans)
        dm = None if diffuse_mask is None else torch.as_tensor(diffuse_mask).to(dev, torch.float32).contiguous()
        rm = torch.empty(B, N, 3, 3, device=dev, dtype=torch.float32) if want_rotmat else None
        st = torch.cuda.current_stream(dev)
        check(self.lib.fd_reverse_step(self._h, B, N, _ptr(r), _ptr(rs), _ptr(ts), _ptr(dm), float(t), float(dt), int(center),
                                       float(noise_scale), _ptr(zr), _ptr(zt), seed, first_sample, step, _ptr(rm),
                                       C.c_void_p(st.cuda_stream)))
        st.synchronize()
        return r, rm

    def compute_backbone(self, rigids: torch.Tensor, psi: torch.Tensor):
        dev = self.device
        shp = rigids.shape[:-1]
        r = rigids.to(dev, torch.float32).contiguous().reshape(-1, 7)
        p = psi.to(dev, torch.float32).contiguous().reshape(-1, 2)
        n = r.shape[0]
        a37 = torch.empty(n, 37, 3, device=dev, dtype=torch.float32)
        a14 = torch.empty(n, 14, 3, device=dev, dtype=torch.float32)
        st = torch.cuda.current_stream(dev)
        check(self.lib.fd_compute_backbone(self._h, n, _ptr(r), _ptr(p), _ptr(a37), _ptr(a14), C.c_void_p(st.cuda_stream)))
        st.synchronize()
        return a37.reshape(*shp, 37, 3), a14.reshape(*shp, 14, 3)

    # ---- Experiment.inference_fn ---------------------------------------------------------------------------------------------
    def sample(self, B: int, N: int, num_t: int = 500, min_t: float = 0.01, noise_scale: float = 0.1, center: bool = True,
               self_condition: bool = True, aux_traj: bool = False, seed: int = 123, first_sample: int = 0,
               use_graph: bool = True, rigids_init=None, noise: Optional[dict] = None, res_mask=None, fixed_mask=None,
               seq_idx=None, pinned: bool = True, gt_psi=None) -> dict:
        """Whole reverse loop through fd_sample_host: HOST buffers in / out (H2D + D2H inside the call).

        noise: optional dict of injected numpy draws {z_axis,u_angle,z_trans0,z_rot,z_trans} (float64).
        Returns the reference's inference_fn dict (prot_traj, and rigid_traj/trans_traj/rigid_0_traj/psi_pred when
        aux_traj) plus 'gpu_ms' and 'kernel_launches'.  Without aux_traj, prot_traj has a single frame (the sample).
        """
        if self.use_cached_score:
            raise ValueError("the device-resident loop evaluates the IGSO(3) series (use_cached_score=False, the shipped default); the table "
                             "look-up is available through ScoreNetwork.forward, i.e. the unchanged driver loop")
        cfg = SampleCfg(B, N, num_t, min_t, noise_scale, int(center), int(self_condition), int(aux_traj), seed, first_sample,
                        int(use_graph))
        keep = []

        def host(a, dtype):
            if a is None:
                return None
            x = np.ascontiguousarray(a.detach().cpu().numpy() if torch.is_tensor(a) else a, dtype=dtype)
            keep.append(x)
            return x

        noise = noise or {}
        sin = SampleIn(*[_np_ptr(host(noise.get(k), np.float64)) for k in ("z_axis", "u_angle", "z_trans0", "z_rot", "z_trans")],
                       _np_ptr(host(rigids_init, np.float32)), _np_ptr(host(res_mask, np.float32)),
                       _np_ptr(host(fixed_mask, np.float32)), _np_ptr(host(seq_idx, np.int32)), _np_ptr(host(gt_psi, np.float32)))

        def out_buf(shape):
            t = torch.empty(shape, dtype=torch.float32, pin_memory=pinned)
            keep.append(t)
            return t

        a37, rig, psi = out_buf((B, N, 37, 3)), out_buf((B, N, 7)), out_buf((B, N, 2))
        tr = {}
        if aux_traj:
            tr = {"prot_traj": out_buf((num_t, B, N, 37, 3)), "rigid_traj": out_buf((num_t + 1, B, N, 7)),
                  "trans_traj": out_buf((num_t, B, N, 3)), "rigid_0_traj": out_buf((num_t, B, N, 37, 3))}
        ms, nl = C.c_double(0), C.c_int64(0)
        sout = SampleOut(_ptr(a37), _ptr(rig), _ptr(psi), _ptr(tr.get("prot_traj")), _ptr(tr.get("rigid_traj")),
                         _ptr(tr.get("trans_traj")), _ptr(tr.get("rigid_0_traj")), C.pointer(ms), C.pointer(nl))
        check(self.lib.fd_sample_host(self._h, C.byref(cfg), C.byref(sin), C.byref(sout)))
        ret = {"gpu_ms": ms.value, "kernel_launches": nl.value, "rigids_final": rig.numpy(), "psi_pred": psi.numpy()[None]}
        if aux_traj:
            ret.update({k: v.numpy() for k, v in tr.items()})
        else:
            ret["prot_traj"] = a37.numpy()[None]
        return ret

    def sample_device(self, B: int, N: int, num_t: int = 500, min_t: float = 0.01, noise_scale: float = 0.1, seed: int = 123,
                      first_sample: int = 0, use_graph: bool = True, rigids_init: Optional[torch.Tensor] = None):
        """Device-resident loop (bench `value` leg): returns (atom37 cuda tensor, rigids cuda tensor, gpu_ms, launches)."""
        cfg = SampleCfg(B, N, num_t, min_t, noise_scale, 1, 1, 0, seed, first_sample, int(use_graph))
        a37 = torch.empty(B, N, 37, 3, device=self.device, dtype=torch.float32)
        rig = torch.empty(B, N, 7, device=self.device, dtype=torch.float32)
        ms, nl = C.c_double(0), C.c_int64(0)
        ri = None if rigids_init is None else rigids_init.to(self.device, torch.float32).contiguous()
        torch.cuda.synchronize(self.device)
        check(self.lib.fd_sample_dev(self._h, C.byref(cfg), _ptr(ri), _ptr(a37), _ptr(rig), C.byref(ms), C.byref(nl)))
        return a37, rig, ms.value, nl.value

    # ---- introspection -------------------------------------------------------------------------------------------------------
    def stage_timing(self, on: bool):
        check(self.lib.fd_set_stage_timing(self._h, int(on)))

    def launch_count(self) -> int:
        """Kernels launched by this handle so far (difference it around a region; `sample*` restart it)."""
        return int(self.lib.fd_launch_count(self._h))

    def stage_times(self):
        n = self.lib.fd_num_stages()
        ms = (C.c_double * n)()
        nl = (C.c_int64 * n)()
        check(self.lib.fd_stage_times(self._h, ms, nl))
        return {self.lib.fd_stage_name(i).decode(): (ms[i], nl[i]) for i in range(n)}

    def forward_flops(self, B, N, executed=True):
        return int(self.lib.fd_forward_flops(B, N, int(executed)))


# ---- additional SE3Diffuser entry points (appended: forward_marginal / score_scaling / inference_fn) ------------------------
def _forward_marginal(self, rigids_0: torch.Tensor, t: float, z_axis, u_angle, z_trans, diffuse_mask=None):
    """fd_forward_marginal: noises one example [n,7] at time t with the caller's numpy draws.  Returns a dict of tensors on the
    engine's device plus the two host scalings."""
    dev = self.device
    r0 = rigids_0.to(dev, torch.float32).contiguous().reshape(-1, 7)
    n = r0.shape[0]
    d64 = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    za, ua, zt = d64(z_axis), d64(u_angle), d64(z_trans)
    dm = None if diffuse_mask is None else torch.as_tensor(np.ascontiguousarray(diffuse_mask, dtype=np.float32)).to(dev).reshape(-1)
    rt = torch.empty(n, 7, device=dev, dtype=torch.float32)
    rs = torch.empty(n, 3, device=dev, dtype=torch.float64)
    ts = torch.empty(n, 3, device=dev, dtype=torch.float64)
    a, b = C.c_double(0), C.c_double(0)
    st = torch.cuda.current_stream(dev)
    check(self.lib.fd_forward_marginal(self._h, n, _ptr(r0), float(t), _ptr(za), _ptr(ua), _ptr(zt), _ptr(dm), _ptr(rt), _ptr(rs), _ptr(ts),
                                       C.byref(a), C.byref(b), C.c_void_p(st.cuda_stream)))
    st.synchronize()
    return {"rigids_t": rt, "rot_score": rs, "trans_score": ts, "rot_score_scaling": a.value, "trans_score_scaling": b.value}


def _forward_marginal_batch(self, rigids_0: torch.Tensor, t, z_axis, u_angle, z_trans, res_mask):
    """fd_forward_marginal_batch: noises a padded batch [B,N,7] in one call, example b at time t[b]; padded rows (res_mask 0) come back zero like
    du.pad_feats.  Returns the loader's feature dict for the batch (device tensors; scalings as float64 tensors [B])."""
    dev = self.device
    r0 = torch.as_tensor(rigids_0).to(dev, torch.float32).contiguous()
    B, N = r0.shape[:2]
    d64 = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    za, ua, zt = d64(z_axis), d64(u_angle), d64(z_trans)
    rm = torch.as_tensor(np.ascontiguousarray(res_mask, dtype=np.float32)).to(dev)
    th = np.ascontiguousarray(t, dtype=np.float64).reshape(B)
    rt = torch.empty(B, N, 7, device=dev, dtype=torch.float32)
    rs = torch.empty(B, N, 3, device=dev, dtype=torch.float64)
    ts = torch.empty(B, N, 3, device=dev, dtype=torch.float64)
    a, b = np.empty(B), np.empty(B)
    st = torch.cuda.current_stream(dev)
    check(self.lib.fd_forward_marginal_batch(self._h, B, N, _ptr(r0), _np_ptr(th), _ptr(za), _ptr(ua), _ptr(zt), _ptr(rm), _ptr(rt), _ptr(rs), _ptr(ts),
                                             _np_ptr(a), _np_ptr(b), C.c_void_p(st.cuda_stream)))
    self._fm_keep = (r0, za, ua, zt, rm)
    return {"rigids_t": rt, "rot_score": rs, "trans_score": ts, "rot_score_scaling": torch.tensor(a), "trans_score_scaling": torch.tensor(b),
            "t": torch.tensor(th)}


def _ca_metrics(self, ca: torch.Tensor, n_valid=None, tol_bond: float = 0.1, tol_clash: float = 1.5) -> Dict[str, torch.Tensor]:
    """analysis/metrics.py:120-132 for a batch of CA traces [B,N,3] on the device."""
    dev = self.device
    x = torch.as_tensor(ca).to(dev, torch.float32).contiguous()
    B, N = x.shape[:2]
    nv = None if n_valid is None else torch.as_tensor(n_valid).to(dev, torch.int32).contiguous()
    out = torch.empty(B, 4, device=dev, dtype=torch.float64)
    st = torch.cuda.current_stream(dev)
    check(self.lib.fd_ca_metrics(self._h, B, N, _ptr(x), _ptr(nv), float(tol_bond), float(tol_clash), _ptr(out), C.c_void_p(st.cuda_stream)))
    return {"ca_ca_bond_dev": out[:, 0], "ca_ca_valid_percent": out[:, 1], "num_ca_steric_clashes": out[:, 2], "ca_steric_clash_percent": out[:, 3]}


def _score_scaling(self, t: float):
    a, b = C.c_double(0), C.c_double(0)
    check(self.lib.fd_score_scaling(self._h, float(t), C.byref(a), C.byref(b)))
    return a.value, b.value


def _inference_fn(self, data_init: dict, num_t: int = 500, min_t: float = 0.01, center: bool = True, aux_traj: bool = False,
                  self_condition: bool = True, noise_scale: float = 1.0, noise: str = "numpy", seed: int = 123, first_sample: int = 0):
    """Experiment.inference_fn (experiments/train_se3_diffusion.py:718-818) with the whole loop on the device.

    data_init: the reference's feature dict (rigids_t [B,N,7] or [N,7], res_mask, fixed_mask, seq_idx).  noise="numpy" draws the
    per-step Gaussians from the global numpy RNG in the reference's order (so np.random.seed(...) reproduces the reference's
    trajectory); noise="philox" uses the on-device counter-based generator keyed by (seed, first_sample + b).
    Returns the reference's dict: prot_traj [+ rigid_traj, trans_traj, psi_pred, rigid_0_traj when aux_traj].  Deviation (documented in
    INTEGRATION.md): with aux_traj=False prot_traj holds the final frame only ([1,B,N,37,3]) instead of all num_t frames.
    """
    rig = torch.as_tensor(data_init["rigids_t"]).detach().cpu().float()
    if rig.ndim == 2:
        rig = rig[None]
    B, N = rig.shape[:2]
    opt = lambda k, dt: None if k not in data_init else np.ascontiguousarray(torch.as_tensor(data_init[k]).detach().cpu().numpy().reshape(B, N), dtype=dt)
    tors = data_init.get("torsion_angles_sin_cos")     # psi imputed on fixed (motif) residues: model/score_network.py:196-199
    gt_psi = None if tors is None else np.ascontiguousarray(torch.as_tensor(tors).detach().cpu().numpy().reshape(B, N, 7, 2)[:, :, 2, :], dtype=np.float32)
    nz = None
    if noise == "numpy" and num_t > 1:
        zr = np.empty((num_t - 1, B, N, 3)); zx = np.empty((num_t - 1, B, N, 3))
        for s in range(num_t - 1):           # SO(3) draw first, then R^3 (SE3Diffuser.reverse)
            zr[s] = np.random.normal(size=(B, N, 3)); zx[s] = np.random.normal(size=(B, N, 3))
        nz = {"z_rot": zr, "z_trans": zx}
    out = self.sample(B, N, num_t=num_t, min_t=min_t, noise_scale=noise_scale, center=center, self_condition=self_condition,
                      aux_traj=aux_traj, seed=seed, first_sample=first_sample, rigids_init=rig.numpy(), noise=nz,
                      res_mask=opt("res_mask", np.float32), fixed_mask=opt("fixed_mask", np.float32), seq_idx=opt("seq_idx", np.int32),
                      gt_psi=gt_psi)
    ret = {"prot_traj": out["prot_traj"]}
    if aux_traj:
        ret.update({"rigid_traj": out["rigid_traj"], "trans_traj": out["trans_traj"], "psi_pred": torch.as_tensor(out["psi_pred"]),
                    "rigid_0_traj": out["rigid_0_traj"]})
    self.last_gpu_ms = out["gpu_ms"]      # timing is an attribute, not a key: Sampler.sample tree-maps over the returned dict
    return ret


FrameDiffEngine.forward_marginal = _forward_marginal
FrameDiffEngine.forward_marginal_batch = _forward_marginal_batch
FrameDiffEngine.ca_metrics = _ca_metrics
FrameDiffEngine.score_scaling = _score_scaling
FrameDiffEngine.inference_fn = _inference_fn


# ---- training loss, forward values (Experiment.loss_fn, experiments/train_se3_diffusion.py:524-693) -------------------------------
DEFAULT_EXP_CONF = dict(trans_loss_weight=1.0, rot_loss_weight=0.5, rot_loss_t_threshold=0.2, separate_rot_loss=True, trans_x0_threshold=1.0,
                        coordinate_scaling=0.1, bb_atom_loss_weight=1.0, bb_atom_loss_t_filter=0.25, dist_mat_loss_weight=1.0,
                        dist_mat_loss_t_filter=0.25, aux_loss_weight=0.25)   # config/base.yaml:104-115


def loss_forward(engine: "FrameDiffEngine", model_out: dict, batch: dict, exp_conf: Optional[dict] = None, diffuse_trans: bool = True,
                 diffuse_rot: bool = True) -> dict:
    """The loss values of the reference's loss_fn for given model outputs and noised batch, computed on the device by
    fd_loss_forward: returns aux_data's entries (`batch_*` per sample, the normalised scalars, `examples_per_step`, `res_length`) as
    torch tensors on the engine's device.  No autograd graph: the backward pass is not part of this library yet."""
    from ._lib import LossCfg, LossIn
    dev = engine.device if isinstance(engine.device, torch.device) else torch.device("cuda", int(engine.device))
    conf = dict(DEFAULT_EXP_CONF, **({} if exp_conf is None else dict(exp_conf)))

    def dv(x, dtype):
        return torch.as_tensor(x).to(device=dev, dtype=dtype).contiguous()

    res_mask = dv(batch["res_mask"], torch.float32)
    B, N = res_mask.shape
    keep = dict(
        pred_rot_score=dv(model_out["rot_score"], torch.float64), pred_trans_score=dv(model_out["trans_score"], torch.float64),
        pred_rigids=dv(model_out["rigids"], torch.float32), pred_atom37=dv(model_out["atom37"], torch.float32),
        gt_rot_score=dv(batch["rot_score"], torch.float64), gt_trans_score=dv(batch["trans_score"], torch.float64),
        rot_score_scaling=dv(batch["rot_score_scaling"], torch.float64), trans_score_scaling=dv(batch["trans_score_scaling"], torch.float64),
        rigids_0=dv(batch["rigids_0"], torch.float64), t=dv(batch["t"], torch.float64), res_mask=res_mask,
        fixed_mask=dv(batch["fixed_mask"], torch.float32), gt_psi=dv(torch.as_tensor(batch["torsion_angles_sin_cos"])[..., 2, :], torch.float32))
    assert keep["pred_atom37"].shape == (B, N, 37, 3) and keep["rigids_0"].shape == (B, N, 7)
    lin = LossIn(*[_ptr(keep[k]) for k, _ in LossIn._fields_])
    cfg = LossCfg(*[float(conf[k]) for k, ty in LossCfg._fields_ if ty is C.c_double], int(bool(conf["separate_rot_loss"])),
                  int(bool(diffuse_trans)), int(bool(diffuse_rot)))
    terms = torch.empty((B, 5), dtype=torch.float64, device=dev)
    check(engine.lib.fd_loss_forward(engine._h, B, N, C.byref(lin), C.byref(cfg), _ptr(terms), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    n_valid = torch.any(res_mask > 0, dim=-1).sum() + 1e-10
    out = {"batch_rot_loss": terms[:, 0], "batch_trans_loss": terms[:, 1], "batch_bb_atom_loss": terms[:, 2],
           "batch_dist_mat_loss": terms[:, 3], "batch_train_loss": terms[:, 4]}
    for k in ("rot", "trans", "bb_atom", "dist_mat"):
        out[f"{k}_loss"] = out[f"batch_{k}_loss"].sum() / n_valid
    out["total_loss"] = out["batch_train_loss"].sum() / n_valid
    out["examples_per_step"] = torch.tensor(B)
    out["res_length"] = torch.mean(torch.sum(res_mask.double(), dim=-1))
    return out


FrameDiffEngine.loss_forward = loss_forward


# ---- training step (include/framediff_b200.h fd_train_*) ---------------------------------------------------------
def arena_layout():
    """[(name, shape, float offset)] of the flat parameter / gradient arenas and their total length in floats."""
    lib = _lib.load()
    sch = param_schema()
    return [(n, s, int(lib.fd_train_param_offset(i))) for i, (n, s) in enumerate(sch)], int(lib.fd_train_arena_floats())


def flat_from_state(state: Dict[str, "np.ndarray | torch.Tensor"], device) -> torch.Tensor:
    """Packs a state_dict into a new flat fp32 arena on `device` (padding zero)."""
    lay, total = arena_layout()
    flat = torch.zeros(total, dtype=torch.float32, device=device)
    state = {(k[7:] if k.startswith("module.") else k): v for k, v in state.items()}
    for name, shape, off in lay:
        v = torch.as_tensor(np.asarray(state[name]) if not torch.is_tensor(state[name]) else state[name]).detach().to(device, torch.float32)
        if tuple(v.shape) != tuple(shape):
            raise ValueError(f"parameter {name}: shape {tuple(v.shape)} != {shape}")
        flat[off:off + v.numel()] = v.reshape(-1)
    return flat


def views_of(flat: torch.Tensor) -> Dict[str, torch.Tensor]:
    lay, _ = arena_layout()
    return {name: flat[off:off + int(np.prod(shape))].view(shape) for name, shape, off in lay}


def _train_bind(self, params: torch.Tensor, grads: torch.Tensor):
    _, total = arena_layout()
    for t in (params, grads):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == total and t.device == self.device):
            raise ValueError("fd_train_bind needs two contiguous fp32 CUDA tensors of arena_layout()[1] floats on the engine's device")
    check(self.lib.fd_train_bind(self._h, _ptr(params), _ptr(grads)))
    self._train_arenas = (params, grads)


def _train_forward(self, feats: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """ScoreNetwork.forward in training mode (autograd semantics of the sequence attention); keeps the tape for train_backward."""
    dev = self.device
    rig = feats["rigids_t"]
    B, N = rig.shape[0], rig.shape[1]
    t_in = torch.as_tensor(feats["t"])
    t_is_f32 = 0 if t_in.dtype == torch.float64 else 1
    if t_in.device.type == "cpu" and bool(((t_in < 0) | (t_in > 1)).any()):
        raise ValueError(f"Invalid t={t_in}")
    f32 = lambda x: torch.as_tensor(x).to(dev, torch.float32).contiguous()
    tors = feats.get("torsion_angles_sin_cos")
    keep = dict(rigids_t=f32(rig), t=t_in.to(dev, torch.float64).contiguous(), res_mask=f32(feats["res_mask"]), fixed_mask=f32(feats["fixed_mask"]),
                seq_idx=torch.as_tensor(feats["seq_idx"]).to(dev, torch.int32).contiguous(), sc_ca=f32(feats["sc_ca_t"]),
                gt_psi=f32(torch.as_tensor(tors)[..., 2, :]) if tors is not None else None)
    out = {"rot_score": torch.empty(B, N, 3, device=dev, dtype=torch.float64), "trans_score": torch.empty(B, N, 3, device=dev, dtype=torch.float64),
           "psi": torch.empty(B, N, 2, device=dev, dtype=torch.float32), "rigids": torch.empty(B, N, 7, device=dev, dtype=torch.float32),
           "atom37": torch.empty(B, N, 37, 3, device=dev, dtype=torch.float32), "atom14": torch.empty(B, N, 14, 3, device=dev, dtype=torch.float32)}
    fin = ForwardIn(_ptr(keep["rigids_t"]), _ptr(keep["t"]), t_is_f32, None, _ptr(keep["res_mask"]), _ptr(keep["fixed_mask"]), _ptr(keep["seq_idx"]),
                    _ptr(keep["sc_ca"]), _ptr(keep["gt_psi"]), None)
    fout = ForwardOut(*[_ptr(out[k]) for k in ("rot_score", "trans_score", "psi", "rigids", "atom37", "atom14")])
    st = torch.cuda.current_stream(dev)
    check(self.lib.fd_train_forward(self._h, B, N, C.byref(fin), C.byref(fout), C.c_void_p(st.cuda_stream)))
    self._train_keep = keep            # inputs must outlive the backward
    if t_is_f32:
        out["trans_score"] = out["trans_score"].to(torch.float32)
    if tors is not None and torch.as_tensor(tors).dtype == torch.float64:
        out["psi"] = out["psi"].to(torch.float64)
    return out


def _train_backward(self, dout: Dict[str, Optional[torch.Tensor]], stage_first: int = 0, stage_last: int = 3):
    """Backward of the last train_forward from the gradients w.r.t. its outputs (missing keys = zero); accumulates into the bound gradient
    arena.  Stages 0..3 = gradient buckets (torsion head + block 3 | block 2 | block 1 | block 0 + embedders)."""
    from ._lib import TrainGrads
    dev = self.device
    conv = lambda k, dt: None if dout.get(k) is None else dout[k].to(dev, dt).contiguous()
    keep = [conv("rot_score", torch.float64), conv("trans_score", torch.float64), conv("rigids", torch.float32), conv("atom37", torch.float32),
            conv("atom14", torch.float32), conv("psi", torch.float32)]
    g = TrainGrads(*[_ptr(t) for t in keep])
    st = torch.cuda.current_stream(dev)
    check(self.lib.fd_train_backward(self._h, C.byref(g), int(stage_first), int(stage_last), C.c_void_p(st.cuda_stream)))
    self._train_dout_keep = keep


def _loss_backward(self, model_out: dict, batch: dict, exp_conf: Optional[dict] = None, diffuse_trans: bool = True, diffuse_rot: bool = True) -> dict:
    """d total_loss / d model outputs of the reference's loss_fn, on the device (fd_loss_backward)."""
    from ._lib import LossCfg, LossIn, TrainGradsOut
    dev = self.device
    conf = dict(DEFAULT_EXP_CONF, **({} if exp_conf is None else dict(exp_conf)))
    dv = lambda x, dtype: torch.as_tensor(x).to(device=dev, dtype=dtype).contiguous()
    res_mask = dv(batch["res_mask"], torch.float32)
    B, N = res_mask.shape
    keep = dict(
        pred_rot_score=dv(model_out["rot_score"], torch.float64), pred_trans_score=dv(model_out["trans_score"], torch.float64),
        pred_rigids=dv(model_out["rigids"], torch.float32), pred_atom37=dv(model_out["atom37"], torch.float32),
        gt_rot_score=dv(batch["rot_score"], torch.float64), gt_trans_score=dv(batch["trans_score"], torch.float64),
        rot_score_scaling=dv(batch["rot_score_scaling"], torch.float64), trans_score_scaling=dv(batch["trans_score_scaling"], torch.float64),
        rigids_0=dv(batch["rigids_0"], torch.float64), t=dv(batch["t"], torch.float64), res_mask=res_mask,
        fixed_mask=dv(batch["fixed_mask"], torch.float32), gt_psi=dv(torch.as_tensor(batch["torsion_angles_sin_cos"])[..., 2, :], torch.float32))
    lin = LossIn(*[_ptr(keep[k]) for k, _ in LossIn._fields_])
    cfg = LossCfg(*[float(conf[k]) for k, ty in LossCfg._fields_ if ty is C.c_double], int(bool(conf["separate_rot_loss"])),
                  int(bool(diffuse_trans)), int(bool(diffuse_rot)))
    out = {"rot_score": torch.empty(B, N, 3, device=dev, dtype=torch.float64), "trans_score": torch.empty(B, N, 3, device=dev, dtype=torch.float64),
           "rigids": torch.empty(B, N, 7, device=dev, dtype=torch.float32), "atom37": torch.empty(B, N, 37, 3, device=dev, dtype=torch.float32)}
    go = TrainGradsOut(*[_ptr(out[k]) for k in ("rot_score", "trans_score", "rigids", "atom37")])
    st = torch.cuda.current_stream(dev)
    check(self.lib.fd_loss_backward(self._h, B, N, C.byref(lin), C.byref(cfg), C.byref(go), C.c_void_p(st.cuda_stream)))
    return out


def _adam_step(self, params: torch.Tensor, grads: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step: int, lr: float = 1e-4,
               betas=(0.9, 0.999), eps: float = 1e-8, grad_scale: float = 1.0):
    st = torch.cuda.current_stream(self.device)
    check(self.lib.fd_adam_step(self._h, _ptr(params), _ptr(grads), _ptr(exp_avg), _ptr(exp_avg_sq), params.numel(), float(lr), float(betas[0]),
                                float(betas[1]), float(eps), int(step), float(grad_scale), C.c_void_p(st.cuda_stream)))


def _train_set_gemm(self, mode: str):
    """'fp32' (CUDA cores), 'bf16x3' (split-bf16 mma.sync for every GEMM) or 'tc' (bf16x3 with the edge-tensor forward / dgrad GEMMs on tcgen05)."""
    check(self.lib.fd_train_set_gemm(self._h, {"fp32": 0, "bf16x3": 1, "tc": 2}[mode]))
    self.train_gemm = mode


FrameDiffEngine.train_set_gemm = _train_set_gemm
FrameDiffEngine.train_bind = _train_bind
FrameDiffEngine.train_forward = _train_forward
FrameDiffEngine.train_backward = _train_backward
FrameDiffEngine.loss_backward = _loss_backward
FrameDiffEngine.adam_step = _adam_step
