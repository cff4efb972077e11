"""ctypes binding of libframediff_b200.so (include/framediff_b200.h).  Fails loudly when the library is missing."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libframediff_b200.so")

c_f32p, c_f64p, c_i32p, c_voidp = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_void_p


class ForwardIn(C.Structure):
    _fields_ = [("rigids_t", c_voidp), ("t", c_voidp), ("t_is_f32", C.c_int), ("sigma", c_voidp), ("res_mask", c_voidp),
                ("fixed_mask", c_voidp), ("seq_idx", c_voidp), ("sc_ca_t", c_voidp), ("gt_psi", c_voidp), ("cached_score_rows", c_voidp)]


class LossIn(C.Structure):
    _fields_ = [(k, c_voidp) for k in ("pred_rot_score", "pred_trans_score", "pred_rigids", "pred_atom37", "gt_rot_score", "gt_trans_score",
                                       "rot_score_scaling", "trans_score_scaling", "rigids_0", "t", "res_mask", "fixed_mask", "gt_psi")]


class LossCfg(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("trans_loss_weight", "rot_loss_weight", "rot_loss_t_threshold", "trans_x0_threshold",
                                          "coordinate_scaling", "bb_atom_loss_weight", "bb_atom_loss_t_filter", "dist_mat_loss_weight",
                                          "dist_mat_loss_t_filter", "aux_loss_weight")] + \
               [(k, C.c_int) for k in ("separate_rot_loss", "diffuse_trans", "diffuse_rot")]


class TrainGrads(C.Structure):
    _fields_ = [(k, c_voidp) for k in ("d_rot_score", "d_trans_score", "d_rigids", "d_atom37", "d_atom14", "d_psi")]


class TrainGradsOut(C.Structure):
    _fields_ = [(k, c_voidp) for k in ("d_rot_score", "d_trans_score", "d_rigids", "d_atom37")]


class ForwardOut(C.Structure):
    _fields_ = [("rot_score", c_voidp), ("trans_score", c_voidp), ("psi", c_voidp), ("rigids", c_voidp),
                ("atom37", c_voidp), ("atom14", c_voidp)]


class SampleCfg(C.Structure):
    _fields_ = [("B", C.c_int), ("N", C.c_int), ("num_t", C.c_int), ("min_t", C.c_double), ("noise_scale", C.c_double),
                ("center", C.c_int), ("self_condition", C.c_int), ("aux_traj", C.c_int), ("seed", C.c_uint64),
                ("first_sample", C.c_int64), ("use_graph", C.c_int)]


class SampleIn(C.Structure):
    _fields_ = [("z_axis", c_voidp), ("u_angle", c_voidp), ("z_trans0", c_voidp), ("z_rot", c_voidp), ("z_trans", c_voidp),
                ("rigids_init", c_voidp), ("res_mask", c_voidp), ("fixed_mask", c_voidp), ("seq_idx", c_voidp), ("gt_psi", c_voidp)]


class SampleOut(C.Structure):
    _fields_ = [("atom37_final", c_voidp), ("rigids_final", c_voidp), ("psi_final", c_voidp), ("prot_traj", c_voidp),
                ("rigid_traj", c_voidp), ("trans_traj", c_voidp), ("rigid_0_traj", c_voidp), ("gpu_ms", c_f64p),
                ("kernel_launches", C.POINTER(C.c_int64))]


# every symbol include/framediff_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "fd_last_error": (C.c_char_p, []),
    "fd_version": (C.c_char_p, []),
    "fd_create": (C.c_int, [C.POINTER(c_voidp), C.c_int]),
    "fd_destroy": (C.c_int, [c_voidp]),
    "fd_set_precision": (C.c_int, [c_voidp, C.c_int]),
    "fd_get_precision": (C.c_int, [c_voidp]),
    "fd_num_params": (C.c_int, []),
    "fd_param_name": (C.c_char_p, [C.c_int]),
    "fd_param_ndim": (C.c_int, [C.c_int]),
    "fd_param_dim": (C.c_int64, [C.c_int, C.c_int]),
    "fd_param_numel": (C.c_int64, [C.c_int]),
    "fd_load_weights": (C.c_int, [c_voidp, C.POINTER(c_voidp)]),
    "fd_forward": (C.c_int, [c_voidp, C.c_int, C.c_int, C.POINTER(ForwardIn), C.POINTER(ForwardOut), c_voidp]),
    "fd_set_debug": (C.c_int, [c_voidp, C.c_int]),
    "fd_debug_fetch": (C.c_int64, [c_voidp, C.c_char_p, c_voidp, C.c_int64]),
    "fd_igso3_score": (C.c_int, [c_voidp, C.c_int64, c_voidp, c_voidp, c_voidp, c_voidp]),
    "fd_igso3_tables_host": (C.c_int, [c_voidp, C.c_int, c_voidp, c_voidp, c_voidp, c_voidp, c_voidp]),
    "fd_sample_ref": (C.c_int, [c_voidp, C.c_int64, c_voidp, c_voidp, c_voidp, C.c_uint64, C.c_int64, C.c_int, c_voidp, c_voidp]),
    "fd_reverse_step": (C.c_int, [c_voidp, C.c_int, C.c_int, c_voidp, c_voidp, c_voidp, c_voidp, C.c_double, C.c_double, C.c_int,
                                  C.c_double, c_voidp, c_voidp, C.c_uint64, C.c_int64, C.c_int, c_voidp, c_voidp]),
    "fd_forward_marginal": (C.c_int, [c_voidp, C.c_int64, c_voidp, C.c_double, c_voidp, c_voidp, c_voidp, c_voidp, c_voidp, c_voidp, c_voidp,
                                      c_f64p, c_f64p, c_voidp]),
    "fd_forward_marginal_batch": (C.c_int, [c_voidp, C.c_int, C.c_int, c_voidp, c_voidp, c_voidp, c_voidp, c_voidp, c_voidp, c_voidp, c_voidp, c_voidp,
                                            c_voidp, c_voidp, c_voidp]),
    "fd_ca_metrics": (C.c_int, [c_voidp, C.c_int, C.c_int, c_voidp, c_voidp, C.c_double, C.c_double, c_voidp, c_voidp]),
    "fd_score_scaling": (C.c_int, [c_voidp, C.c_double, c_f64p, c_f64p]),
    "fd_compute_backbone": (C.c_int, [c_voidp, C.c_int64, c_voidp, c_voidp, c_voidp, c_voidp, c_voidp]),
    "fd_sample_host": (C.c_int, [c_voidp, C.POINTER(SampleCfg), C.POINTER(SampleIn), C.POINTER(SampleOut)]),
    "fd_sample_dev": (C.c_int, [c_voidp, C.POINTER(SampleCfg), c_voidp, c_voidp, c_voidp, c_f64p, C.POINTER(C.c_int64)]),
    "fd_num_stages": (C.c_int, []),
    "fd_stage_name": (C.c_char_p, [C.c_int]),
    "fd_set_stage_timing": (C.c_int, [c_voidp, C.c_int]),
    "fd_launch_count": (C.c_int64, [c_voidp]),
    "fd_stage_times": (C.c_int, [c_voidp, c_f64p, C.POINTER(C.c_int64)]),
    "fd_forward_flops": (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    "fd_debug_tc_profile": (C.c_int, [c_voidp, C.c_int, C.POINTER(C.c_longlong)]),
    "fd_loss_forward": (C.c_int, [c_voidp, C.c_int, C.c_int, c_voidp, c_voidp, c_voidp, c_voidp]),
    "fd_train_arena_floats": (C.c_int64, []),
    "fd_train_param_offset": (C.c_int64, [C.c_int]),
    "fd_train_bind": (C.c_int, [c_voidp, c_voidp, c_voidp]),
    "fd_train_forward": (C.c_int, [c_voidp, C.c_int, C.c_int, C.POINTER(ForwardIn), C.POINTER(ForwardOut), c_voidp]),
    "fd_train_backward": (C.c_int, [c_voidp, C.POINTER(TrainGrads), C.c_int, C.c_int, c_voidp]),
    "fd_train_release": (C.c_int, [c_voidp]),
    "fd_workspace_bytes": (C.c_int64, [c_voidp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "fd_debug_alloc_bytes": (C.c_int64, [c_voidp, C.c_int]),
    "fd_train_set_gemm": (C.c_int, [c_voidp, C.c_int]),
    "fd_loss_backward": (C.c_int, [c_voidp, C.c_int, C.c_int, c_voidp, c_voidp, C.POINTER(TrainGradsOut), c_voidp]),
    "fd_adam_step": (C.c_int, [c_voidp, c_voidp, c_voidp, c_voidp, c_voidp, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64,
                               C.c_double, c_voidp]),
    "fd_format_pdb": (C.c_int, [c_voidp, C.c_int, C.POINTER(C.c_ubyte), C.POINTER(C.c_int), c_f64p, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                C.POINTER(C.c_size_t)]),
}

_lib = None


class FrameDiffError(RuntimeError):
    pass


def load():
    """Loads the CUDA library.  No fallback: a missing .so is an error (build with `python -m se3_diffusion_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FrameDiffError(f"{LIB_PATH} not found — the FrameDiff B200 path has no CPU fallback; run "
                             f"`python -m se3_diffusion_b200.build` (needs nvcc) first")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().fd_last_error().decode()
        if rc == -1:
            raise ValueError(msg)      # the reference raises ValueError on bad t / missing impute
        raise FrameDiffError(f"libframediff_b200 error {rc}: {msg}")
