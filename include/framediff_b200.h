/*
 * framediff_b200.h — C ABI of libframediff_b200.so: the B200-native FrameDiff hot path.
 *
 * The reference (jasonkyuyim/se3_diffusion) is pure Python: it has no FFI / operator interface of its own.  Its
 * seam for this path is the import-level API
 *     model.score_network.ScoreNetwork.forward             (/root/reference/model/score_network.py:170-215)
 *     data.se3_diffuser.SE3Diffuser.{sample_ref,reverse,forward_marginal,score,calc_rot_score,calc_trans_score}
 *                                                          (/root/reference/data/se3_diffuser.py:43-268)
 *     data.all_atom.compute_backbone                       (/root/reference/data/all_atom.py:152-174)
 *     experiments.train_se3_diffusion.Experiment.inference_fn (…/experiments/train_se3_diffusion.py:718-818)
 * Each entry point below states which of those it replaces.  The Python overlay modules in
 * se3_diffusion_b200/overlay/ bind these symbols with ctypes and re-export the reference's class names, so the
 * reference drivers run unchanged.
 *
 * Conventions
 *   - plain C types only: raw pointers + sizes; no torch / C++ types cross the boundary;
 *   - every function returns 0 on success, a negative FD_E* code on failure; fd_last_error() gives the message;
 *   - "dev" pointers are CUDA device pointers on the handle's device, "host" pointers are host memory;
 *   - all tensors are dense row-major (C order), batch first; B = backbones in the batch, N = residues;
 *   - `stream` is a cudaStream_t passed as void* (NULL = the CUDA default stream); device-API calls are
 *     asynchronous on it; the whole-loop calls (fd_sample_*) run on the handle's own stream and synchronise
 *     before returning;
 *   - one handle per (process, device); a handle is not thread-safe (the reference is single-threaded Python).
 */
#ifndef FRAMEDIFF_B200_H
#define FRAMEDIFF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fd_context* fd_handle;

enum {
  FD_OK = 0,
  FD_EINVAL = -1,   /* bad argument (shape, null pointer, t outside [0,1], …) — the reference raises ValueError */
  FD_ECUDA = -2,    /* CUDA runtime / driver error */
  FD_ENOMEM = -3,   /* device allocation failed */
  FD_ESTATE = -4    /* call order (weights not loaded, …) */
};

/* GEMM operand precision of the edge-tensor kernels (EdgeTransition / edge embedder).  Everything else is fp32
 * (frames, points, softmax statistics) or fp64 (IGSO(3) series, reverse step). */
enum {
  FD_PREC_FP32 = 0,      /* CUDA-core fp32 everywhere (exact-parity mode) */
  FD_PREC_BF16X3 = 1,    /* tcgen05 bf16 tensor cores, 3-term split (hi·hi + hi·lo + lo·hi), fp32 accumulate */
  FD_PREC_BF16 = 2       /* tcgen05 bf16 tensor cores, single term, fp32 accumulate (throughput mode) */
};

const char* fd_last_error(void);
const char* fd_version(void);

/* ---- lifetime ------------------------------------------------------------------------------------------- */
int fd_create(fd_handle* out, int device);
int fd_destroy(fd_handle h);
int fd_set_precision(fd_handle h, int prec);
int fd_get_precision(fd_handle h);

/* ---- parameters: the reference's 282-entry state_dict, fp32 ------------------------- */
int fd_num_params(void);
const char* fd_param_name(int i);
int fd_param_ndim(int i);
int64_t fd_param_dim(int i, int d);
int64_t fd_param_numel(int i);
/* host_ptrs[i] -> fp32 host array of parameter i (schema order).  Packs / transposes / splits into the device arena.
 * Replaces ScoreNetwork.load_state_dict (experiments/inference_se3_diffusion.py:151-154). */
int fd_load_weights(fd_handle h, const float* const* host_ptrs);

/* ---- ScoreNetwork.forward (model/score_network.py:170-215) -------------------------------------------------- */
typedef struct {
  const float* rigids_t;        /* [B,N,7] (qw,qx,qy,qz,tx,ty,tz), Å */
  const double* t;              /* [B] diffusion time in [0,1], widened to float64 */
  int t_is_f32;                 /* 1: the caller's t tensor was fp32 (Experiment.inference_fn): t*1e4 and the R^3 score are
                                   then evaluated in fp32 like the reference; 0: t was float64 (numpy-born features) */
  const double* sigma;          /* [B] IGSO(3) sigma already quantised to the 1000-point grid (so3_diffuser.py:301);
                                   NULL = quantise t on the device */
  const float* res_mask;        /* [B,N] */
  const float* fixed_mask;      /* [B,N] */
  const int32_t* seq_idx;       /* [B,N] 1-based residue index, 0 on padding */
  const float* sc_ca_t;         /* [B,N,3] self-conditioning CA (Å) */
  const float* gt_psi;          /* [B,N,2] torsion_angles_sin_cos[..., 2, :] (used where fixed_mask = 1); may be NULL */
  const double* cached_score_rows; /* use_cached_score=True (so3_diffuser.py:291-298): [B,1000] fp64, row b = _score_norms[sigma_idx(t[b])]
                                   (fd_igso3_tables_host builds them); the head then does the reference's bucketize + gather instead of
                                   evaluating the series.  NULL (the shipped default) = series.  Inference only. */
} fd_forward_in;

typedef struct {
  double* rot_score;            /* [B,N,3] float64, like the reference */
  double* trans_score;          /* [B,N,3] float64 */
  float* psi;                   /* [B,N,2] (sin, cos) */
  float* rigids;                /* [B,N,7] predicted frames, Å */
  float* atom37;                /* [B,N,37,3]; may be NULL */
  float* atom14;                /* [B,N,14,3]; may be NULL */
} fd_forward_out;

int fd_forward(fd_handle h, int B, int N, const fd_forward_in* in, const fd_forward_out* out, void* stream);

/* Debug taps: after fd_set_debug(h,1) every forward keeps copies of named intermediates
 * ("node_embed","edge_embed","node_<b>","edge_<b>","ipa_feats_<b>","attn_<b>","quat_<b>","trans_<b>").
 * fd_debug_fetch copies one to a HOST buffer (returns its byte size; dst may be NULL to query). */
int fd_set_debug(fd_handle h, int on);
int64_t fd_debug_fetch(fd_handle h, const char* name, void* dst_host, int64_t dst_bytes);

/* ---- SE3Diffuser pieces (data/se3_diffuser.py, so3_diffuser.py, r3_diffuser.py) ------------------------------ */
/* SO3Diffuser.torch_score (so3_diffuser.py:274-305) over n rotation vectors: vec [n,3] fp32, sigma [n] fp64
 * (quantised), score_out [n,3] fp64.  Device pointers. */
int fd_igso3_score(fd_handle h, int64_t n, const float* vec, const double* sigma, double* score_out, void* stream);

/* SO3Diffuser.__init__ cache rows (so3_diffuser.py:151-180) for `nrows` sigma-grid indices: pdf/cdf/score_norms
 * [nrows,1000] fp64 and score_scaling [nrows].  HOST pointers (any may be NULL).  Computed on the GPU. */
int fd_igso3_tables_host(fd_handle h, int nrows, const int32_t* sigma_idx, double* pdf, double* cdf,
                         double* score_norms, double* score_scaling);

/* SE3Diffuser.sample_ref (se3_diffuser.py:216-268): prior frames for n residues.  Either inject the reference's
 * numpy draws (z_axis [n,3] ~N(0,1), u_angle [n] ~U(0,1), z_trans [n,3] ~N(0,1); fp64 device pointers) or pass NULLs
 * and a Philox key (seed, first_sample, residues_per_sample).  rigids_out [n,7] fp32 device. */
int fd_sample_ref(fd_handle h, int64_t n, const double* z_axis, const double* u_angle, const double* z_trans,
                  uint64_t seed, int64_t first_sample, int residues_per_sample, float* rigids_out, void* stream);

/* SE3Diffuser.reverse (se3_diffuser.py:160-214): one reverse-SDE step for B backbones of N residues.
 * rot_score/trans_score [B,N,3] fp64, diffuse_mask [B,N] fp32 or NULL, z_rot/z_trans [B,N,3] fp64 N(0,1) draws or
 * NULL (then Philox(seed, first_sample+b, step, residue)).  t, dt: the reference's python floats.
 * rigids_io [B,N,7] fp32 is updated in place.  rotmat_out [B,N,3,3] fp32 optional (the reference returns a
 * rotation-matrix-backed Rigid). */
int fd_reverse_step(fd_handle h, int B, int N, float* rigids_io, const double* rot_score, const double* trans_score,
                    const float* diffuse_mask, double t, double dt, int center, double noise_scale,
                    const double* z_rot, const double* z_trans, uint64_t seed, int64_t first_sample, int step,
                    float* rotmat_out, void* stream);

/* SE3Diffuser.forward_marginal (se3_diffuser.py:43-110) for the n residues of one training example: noises rigids_0 [n,7]
 * at time t with the caller's numpy draws (z_axis [n,3] = randn, u_angle [n] = rand, z_trans [n,3] = normal; fp64 device
 * pointers), optional diffuse_mask [n].  Outputs rigids_t [n,7] fp32, rot_score / trans_score [n,3] fp64 (device) and the
 * two score scalings (host scalars, may be NULL). */
int fd_forward_marginal(fd_handle h, int64_t n, const float* rigids_0, double t, const double* z_axis, const double* u_angle,
                        const double* z_trans, const float* diffuse_mask, float* rigids_t, double* rot_score, double* trans_score,
                        double* rot_score_scaling, double* trans_score_scaling, void* stream);

/* Batched, padded training-data assembly (data/pdb_data_loader.py:251-272 noises ONE example per DataLoader worker call,
 * data/utils.py:387-399 pads): forward_marginal of B examples padded to N residues, example b at its own time t_host[b] (HOST array),
 * res_mask [B,N] marks real residues (they are the diffused ones, as in the loader); padded rows come back all-zero like du.pad_feats.
 * Draws z_axis/u_angle/z_trans [B,N,*] fp64 device (values on padded rows are ignored).  Scalings: HOST arrays [B] (may be NULL). */
int fd_forward_marginal_batch(fd_handle h, int B, int N, const float* rigids_0, const double* t_host, const double* z_axis,
                              const double* u_angle, const double* z_trans, const float* res_mask, float* rigids_t, double* rot_score,
                              double* trans_score, double* rot_score_scaling_host, double* trans_score_scaling_host, void* stream);

/* Intermittent eval metrics (analysis/metrics.py:120-132 ca_ca_distance, ca_ca_clashes) for B backbones on the device:
 * ca [B,N,3] fp32 (Angstrom), n_valid [B] int32 residues used per backbone (NULL = N); out4 [B,4] fp64 device =
 * { ca_ca_bond_dev, ca_ca_valid_percent, num_ca_steric_clashes, ca_steric_clash_percent } with the reference's tolerances
 * tol_bond = 0.1, tol_clash = 1.5 passed by the caller. */
int fd_ca_metrics(fd_handle h, int B, int N, const float* ca, const int32_t* n_valid, double tol_bond, double tol_clash, double* out4,
                  void* stream);

/* SE3Diffuser.score_scaling (se3_diffuser.py:155-158): host scalars. */
int fd_score_scaling(fd_handle h, double t, double* rot_scaling, double* trans_scaling);

/* all_atom.compute_backbone (data/all_atom.py:152-174): rigids [n,7] fp32 (Å) + psi [n,2] -> atom37 [n,37,3],
 * atom14 [n,14,3] (either may be NULL).  Device pointers. */
int fd_compute_backbone(fd_handle h, int64_t n, const float* rigids, const float* psi, float* atom37, float* atom14,
                        void* stream);

/* ---- Experiment.inference_fn (experiments/train_se3_diffusion.py:718-818): the whole reverse loop ------------ */
typedef struct {
  int B, N;                 /* backbones on this device, residues */
  int num_t;                /* denoise steps (500) */
  double min_t;             /* 0.01 */
  double noise_scale;       /* 0.1 */
  int center;               /* 1 */
  int self_condition;       /* 1 */
  int aux_traj;             /* 0: only final outputs; 1: also per-step trajectories */
  uint64_t seed;            /* Philox key when no noise is injected */
  int64_t first_sample;     /* global index of this device's first backbone (multi-GPU batch sharding) */
  int use_graph;            /* 1: capture one denoise step as a CUDA graph and replay it */
} fd_sample_cfg;

typedef struct {
  /* optional injected noise, HOST fp64; NULL -> Philox on device */
  const double* z_axis;     /* [B,N,3]   prior rotation axes      (np.random.randn) */
  const double* u_angle;    /* [B,N]     prior angle quantiles    (np.random.rand) */
  const double* z_trans0;   /* [B,N,3]   prior translations       (np.random.normal) */
  const double* z_rot;      /* [num_t-1,B,N,3] per-step rotation noise */
  const double* z_trans;    /* [num_t-1,B,N,3] per-step translation noise */
  const float* rigids_init; /* [B,N,7] overrides the prior draw when non-NULL */
  const float* res_mask;    /* [B,N] or NULL (= ones) */
  const float* fixed_mask;  /* [B,N] or NULL (= zeros) */
  const int32_t* seq_idx;   /* [B,N] or NULL (= 1..N) */
  const float* gt_psi;      /* [B,N,2] = torsion_angles_sin_cos[..., 2, :] or NULL (= zeros): the psi imputed on fixed (motif)
                               residues, model/score_network.py:196-199 */
} fd_sample_in;

typedef struct {
  /* HOST buffers; any may be NULL */
  float* atom37_final;      /* [B,N,37,3]  = prot_traj[0] */
  float* rigids_final;      /* [B,N,7] */
  float* psi_final;         /* [B,N,2] */
  float* prot_traj;         /* [num_t,B,N,37,3]   (aux_traj) time-reversed like the reference (index 0 = t≈0) */
  float* rigid_traj;        /* [num_t+1,B,N,7]    (aux_traj) */
  float* trans_traj;        /* [num_t,B,N,3]      (aux_traj) */
  float* rigid_0_traj;      /* [num_t,B,N,37,3]   (aux_traj) */
  double* gpu_ms;           /* device time of the loop (CUDA events), ms */
  int64_t* kernel_launches; /* kernels launched (or graph-replayed) inside the loop */
} fd_sample_out;

int fd_sample_host(fd_handle h, const fd_sample_cfg* cfg, const fd_sample_in* in, const fd_sample_out* out);

/* Device-resident variant used by bench.py's `value` leg: inputs already in HBM, final atom37/rigids stay in HBM.
 * rigids_init_dev [B,N,7] fp32 or NULL.  atom37_dev [B,N,37,3], rigids_dev [B,N,7] (may be NULL). */
int fd_sample_dev(fd_handle h, const fd_sample_cfg* cfg, const float* rigids_init_dev, float* atom37_dev,
                  float* rigids_dev, double* gpu_ms, int64_t* kernel_launches);

/* ---- introspection for bench.py ---------------------------------------------------------------------------- */
/* Times (CUDA events on the handle's stream, ms) of the last forward's stages; names via fd_stage_name. */
int fd_num_stages(void);
const char* fd_stage_name(int i);
int fd_set_stage_timing(fd_handle h, int on);
/* Kernels this handle has launched so far (training and forward calls add to it; fd_sample* restart it at 0): take the difference around a
 * region to count its launches — bench.py's `gpu_launches` of the training step. */
int64_t fd_launch_count(fd_handle h);
int fd_stage_times(fd_handle h, double* ms_out /* [fd_num_stages()] */, int64_t* launches_out);
int64_t fd_forward_flops(int B, int N, int executed);   /* algorithmic FLOPs of one forward */

/* Developer aid: role-level cycle counters of the fused EdgeTransition kernel (CTA 0; tools/profile_fused.py).  on != 0 arms the
 * counters for the following forwards; out32 (host, 32 x int64, may be NULL) receives the last values. */
int fd_debug_tc_profile(fd_handle h, int on, long long* out32);

/* ---- training loss, forward values (first half) ------------------------------------------- */
/* The per-sample terms of Experiment.loss_fn (experiments/train_se3_diffusion.py:538-660) from the model outputs and the
 * noised batch: terms[b] = { rot_loss, trans_loss, bb_atom_loss, dist_mat_loss, sum } (aux_data's `batch_*` entries; the
 * caller divides their sums by the number of non-empty samples, :662).  All pointers device memory.  Scores, scalings,
 * rigids_0 and t are fp64 (as the reference's batch / model outputs are), frames, atoms, masks and psi fp32.
 * Gradients: fd_loss_backward below. */
typedef struct {
  const double* pred_rot_score;    /* [B,N,3]  model_out['rot_score'] */
  const double* pred_trans_score;  /* [B,N,3] */
  const float* pred_rigids;        /* [B,N,7] */
  const float* pred_atom37;        /* [B,N,37,3] */
  const double* gt_rot_score;      /* [B,N,3]  batch['rot_score'] */
  const double* gt_trans_score;    /* [B,N,3] */
  const double* rot_score_scaling; /* [B] */
  const double* trans_score_scaling; /* [B] */
  const double* rigids_0;          /* [B,N,7] */
  const double* t;                 /* [B] */
  const float* res_mask;           /* [B,N] */
  const float* fixed_mask;         /* [B,N] */
  const float* gt_psi;             /* [B,N,2] = batch['torsion_angles_sin_cos'][..., 2, :] */
} fd_loss_in;
typedef struct {                   /* config/base.yaml:104-115 + diffuser.diffuse_{trans,rot} */
  double trans_loss_weight, rot_loss_weight, rot_loss_t_threshold, trans_x0_threshold, coordinate_scaling,
      bb_atom_loss_weight, bb_atom_loss_t_filter, dist_mat_loss_weight, dist_mat_loss_t_filter, aux_loss_weight;
  int separate_rot_loss, diffuse_trans, diffuse_rot;
} fd_loss_cfg;
int fd_loss_forward(fd_handle h, int B, int N, const fd_loss_in* in, const fd_loss_cfg* cfg, double* terms_dev /* [B,5] */, void* stream);

/* ---- training step -------------------------------------------------------------------------- */
/* The reference has no backward code: Experiment.update_fn (experiments/train_se3_diffusion.py:320-326) calls loss.backward() and torch
 * autograd differentiates loss_fn (:524-693) and ScoreNetwork.forward.  Here that derivative is hand-written CUDA.
 * Parameters and gradients live in two flat fp32 device arenas owned by the CALLER, laid out in the state_dict's own order and shapes:
 * parameter i occupies [fd_train_param_offset(i), +fd_param_numel(i)) (offsets are 256-byte aligned; fd_train_arena_floats() floats
 * in total; padding floats are never read).  The nn.Module's parameters / .grad tensors are views into these arenas, so
 * torch.optim.Adam, DDP or a plain NCCL all-reduce over slices of the gradient arena work on them directly.  The backward
 * ACCUMULATES into the gradient arena (zero it per step, like optimizer.zero_grad).  The 10 parameters the reference never uses
 * (linear_rbf x4 blocks, torsion_pred.linear_3) receive no gradient. */
int64_t fd_train_arena_floats(void);
int64_t fd_train_param_offset(int i);
int fd_train_bind(fd_handle h, float* params_dev, float* grads_dev);
/* Training-mode forward: same inputs / outputs as fd_forward, computed from the bound parameter arena with the semantics torch
 * applies whenever autograd records (the float key-padding mask is ADDED to the sequence-attention logits, model/ipa_pytorch.py:636,
 * SURVEY Appendix C.2); every activation the backward needs is kept on a tape owned by the handle (fd_train_release frees it).
 * The input pointers must stay valid until the matching fd_train_backward has run. */
int fd_train_forward(fd_handle h, int B, int N, const fd_forward_in* in, const fd_forward_out* out, void* stream);
/* Gradients w.r.t. the model outputs (device pointers, any may be NULL = zero): what autograd hands to the network's outputs. */
typedef struct {
  const double* d_rot_score;    /* [B,N,3] */
  const double* d_trans_score;  /* [B,N,3] */
  const float* d_rigids;        /* [B,N,7] */
  const float* d_atom37;        /* [B,N,37,3] (atoms 0..4 are functions of the frames / psi, the rest are constant zeros) */
  const float* d_atom14;        /* [B,N,14,3] */
  const float* d_psi;           /* [B,N,2] */
} fd_train_grads;
/* Backward of the last fd_train_forward in four stages — 0: torsion head + trunk block 3, 1: block 2, 2: block 1, 3: block 0 + the
 * embedders — so that the caller can all-reduce a finished gradient bucket (a contiguous slice of the arena) while the next stage
 * computes.  Stages must run in order 0..3; (0,3) runs the whole backward. */
int fd_train_backward(fd_handle h, const fd_train_grads* dout, int stage_first, int stage_last, void* stream);
int fd_train_release(fd_handle h);
/* Workspace query: bytes of device memory the handle allocates for a (B, N) problem (estimate within a few per cent: the
 * arenas add 256-byte alignment per buffer).  which = 0 inference workspace in the current precision mode, 1 sampling-loop buffers for
 * num_t steps (aux != 0: with the per-step trajectories), 2 training tape + backward scratch.  The handle owns these arenas (allocated on
 * first use for a shape, reused until the shape changes); all other buffers are the caller's. */
int64_t fd_workspace_bytes(fd_handle h, int which, int B, int N, int num_t, int aux);
int64_t fd_debug_alloc_bytes(fd_handle h, int which);   /* developer aid: bytes of the arena `which` as currently allocated (0 = none) */
/* Arithmetic of the training path's GEMMs: 0 = fp32 on the CUDA cores (default), 1 = split-bf16 (hi*hi + hi*lo + lo*hi, fp32 accumulate)
 * through mma.sync for every GEMM form, 2 = like 1 with the forward and data-gradient GEMMs over the edge tensor (EdgeTransition, edge
 * embedder layers 2 and 4: 2/3 of the step's FLOPs) on tcgen05 (TMA-staged bf16 planes, TMEM accumulators) — the accuracy class of
 * FD_PREC_BF16X3; all meet the gradient tolerances of tests/test_gpu_train.py. */
int fd_train_set_gemm(fd_handle h, int mode);
/* d total_loss / d model outputs for Experiment.loss_fn (train_se3_diffusion.py:538-680; total_loss = sum_b batch_loss[b] / #non-empty
 * samples): the gradient fd_train_backward starts from when the loss is not computed by torch. */
typedef struct {
  double* d_rot_score;          /* [B,N,3] */
  double* d_trans_score;        /* [B,N,3] */
  float* d_rigids;              /* [B,N,7] */
  float* d_atom37;              /* [B,N,37,3] */
} fd_train_grads_out;
int fd_loss_backward(fd_handle h, int B, int N, const fd_loss_in* in, const fd_loss_cfg* cfg, const fd_train_grads_out* out, void* stream);
/* torch.optim.Adam (defaults of the reference, train_se3_diffusion.py:139-141: no weight decay, no amsgrad) over n contiguous floats;
 * step counts from 1; grads are multiplied by grad_scale first (1/world_size after a SUM all-reduce). */
int fd_adam_step(fd_handle h, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                 double beta2, double eps, int64_t step, double grad_scale, void* stream);

/* ---- downstream data format: PDB text of sampled backbones ---------------------------------- */
/* Host-only (no CUDA call): the text analysis/utils.py:39-77 write_prot_to_pdb + data/protein.py:146-219 to_pdb produce
 * for one chain 'A', residue_index = 0..N-1: per frame `MODEL`, one ATOM line per atom with mask != 0 (atom37 order,
 * serials from 1), `TER`, `ENDMDL`, every line padded to 80 columns + '\n'; after the last frame the 3 bytes `END`.
 * pos [T,N,37,3] (Angstrom) float32 (pos_is_f32 != 0) or float64; mask [T,N,37] (0/1) or NULL = the reference's rule
 * `sum(|pos|, axis=-1) > 1e-7` evaluated in pos's own precision; aatype [N] (0..20, NULL = ALA); b_factors [N,37] (NULL = 0).
 * Writes at most cap bytes to out and the full length to *len; returns FD_EINVAL when cap is too small (len still set)
 * or an aatype is out of range (the reference raises ValueError). */
int fd_format_pdb(const void* pos, int pos_is_f32, const unsigned char* mask, const int* aatype, const double* b_factors, int T, int N,
                  char* out, size_t cap, size_t* len);

#ifdef __cplusplus
}
#endif
#endif /* FRAMEDIFF_B200_H */
