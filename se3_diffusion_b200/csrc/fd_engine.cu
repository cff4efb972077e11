// libframediff_b200.so — context, weight packing, forward orchestration, sampling loop, C ABI (include/framediff_b200.h).
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <string>
#include <vector>

#include "../../include/framediff_b200.h"
#include "fd_common.cuh"
#include "fd_gemm.cuh"
#include "fd_kernels.cuh"
#include "fd_train.cuh"
#include "fd_weights.h"
#include "fd_tc.cuh"
#include "fd_mm3.cuh"

using namespace fd;

// ------------------------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CK(call)                                                                                               \
  do {                                                                                                         \
    cudaError_t _e = (call);                                                                                   \
    if (_e != cudaSuccess)                                                                                     \
      return fail(_e == cudaErrorMemoryAllocation ? FD_ENOMEM : FD_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, \
                  #call, cudaGetErrorString(_e));                                                              \
  } while (0)
#define CKI(call)            \
  do {                       \
    int _r = (call);         \
    if (_r != FD_OK) return _r; \
  } while (0)

// Restores the caller's current device when an entry point returns (the handle's device is only current inside the call).
struct DevGuard {
  int prev = -1, dev;
  explicit DevGuard(int d) : dev(d) { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; if (prev != d) cudaSetDevice(d); }
  ~DevGuard() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
};

extern "C" const char* fd_last_error(void) { return g_err; }
extern "C" const char* fd_version(void) { return "framediff_b200 0.2 (sm_100a)"; }

// ------------------------------------------------------------------------------------------------------------------
// stages (for bench.py's breakdown)
// ------------------------------------------------------------------------------------------------------------------
enum Stage { ST_EMBED_NODE = 0, ST_EMBED_EDGE, ST_IPA_PROJ, ST_IPA_LOGITS, ST_IPA_EDGE, ST_IPA_AV, ST_IPA_OUT, ST_NODE_TFMR,
             ST_NODE_TRANS, ST_EDGE_TRANS, ST_HEADS, ST_COUNT };
static const char* kStageNames[ST_COUNT] = {"embed_node", "embed_edge", "ipa_proj", "ipa_logits", "ipa_edge", "ipa_av",
                                            "ipa_out", "node_tfmr", "node_transition", "edge_transition", "heads"};

// ------------------------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------------------------
struct Workspace {
  int B = 0, N = 0, Np = 0;
  long long rows = 0, edges = 0, chunk = 0;
  char* base = nullptr; size_t bytes = 0;
  float *node_in, *temb, *AC, *node0, *node, *tmpA, *tmpB, *x320, *x320b, *qkv, *S, *y320, *ff, *proj, *qp, *kp, *vp, *optg, *zbar,
      *L, *feats, *quat, *trans, *nb, *pquv, *z, *h1, *h2, *ychunk, *tors;
  // tensor-core path: bf16 hi/lo planes of z and staging
  TcWorkspace tc;
};

struct LoopBufs {   // device state of fd_sample_*
  int B = 0, N = 0, num_t = 0, aux = 0;
  char* base = nullptr; size_t bytes = 0;
  float *rigids, *sc_ca, *res_mask, *fixed_mask, *psi, *rigids_pred, *atom37, *atom37_0, *rigids_snap, *gt_psi;
  bool have_gt_psi = false;
  int* seq_idx;
  double *rot_score, *trans_score, *cur_t, *cur_sigma, *z_rot, *z_trans, *z_axis, *u_angle, *z_trans0;
  StepSched* sched; double* sched_sigma; int* step;
  float *traj_prot, *traj_rigid, *traj_trans0, *traj_bb0;
};

struct fd_context {
  int device = 0;
  cudaStream_t stream = nullptr;
  int precision = FD_PREC_FP32;
  bool weights_loaded = false;
  char* warena = nullptr; size_t warena_bytes = 0;
  Weights W;
  TcWeights tcw;
  Workspace ws;
  LoopBufs lb;
  double* d_sigma_grid = nullptr;    // [1000]
  double* d_cdf_t1 = nullptr;        // [1000] IGSO(3) angle cdf at t = 1
  double* d_omega = nullptr;         // [1000]
  std::vector<double> h_sigma_grid;
  StepSched* d_sched1 = nullptr;     // single-entry schedule for fd_reverse_step
  std::map<int, std::pair<double*, double>> igso3_rows;   // sigma index -> (device cdf row, rot score scaling), built lazily
  double* d_t_tmp = nullptr; size_t t_tmp_n = 0;
  bool debug = false;
  std::map<std::string, std::pair<void*, size_t>> dbg;
  bool stage_timing = false;
  cudaEvent_t ev[ST_COUNT * NBLK * 2 + 8];
  double stage_ms[ST_COUNT];
  long long stage_launches[ST_COUNT];
  long long launches = 0;            // kernels launched since last reset
  int sm_count = 148;
  // one captured denoise step, kept across fd_sample_* calls while every pointer / by-value argument baked into it is unchanged
  struct GraphCache {
    cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr; long long per_step = 0;
    int B = 0, N = 0, num_t = 0, aux = 0, inject = 0, center = 0, precision = -1, have_psi = 0;
    double noise_scale = 0; uint64_t seed = 0; long long first_sample = 0;
    const void *ws_base = nullptr, *lb_base = nullptr, *warena = nullptr;
  } gc;
  double* d_loss_acc = nullptr;            // [4096][10] partial sums of the loss kernels
  int train_gemm = 0;                      // training-path GEMMs: 0 = fp32 CUDA cores, 1 = split-bf16 mma.sync tensor cores (fd_mm3.cuh)
  struct fd_train_state* train = nullptr;   // training step state (bound arenas + tape), fd_train_host.cuh
  cudaEvent_t ev_fwd = nullptr;      // recorded after fd_forward on the caller's stream; the sampling stream waits on it (shared workspace)
  bool fwd_pending = false;
};
static void free_train(fd_context* h);
static void free_graph(fd_context* h) {
  if (h->gc.exec) cudaGraphExecDestroy(h->gc.exec);
  if (h->gc.graph) cudaGraphDestroy(h->gc.graph);
  h->gc = fd_context::GraphCache();
}

struct Launcher {  // counts launches, optional per-stage timing
  fd_context* h; cudaStream_t st;
  int cur = -1; cudaEvent_t e0 = nullptr, e1 = nullptr;
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> spans;
  long long l0 = 0;
  void begin(int stage) {
    if (!h->stage_timing) { cur = stage; l0 = h->launches; return; }
    cur = stage; l0 = h->launches;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
  }
  void end() {
    h->stage_launches[cur] += h->launches - l0;
    if (!h->stage_timing) return;
    cudaEventRecord(e1, st);
    spans.push_back({cur, {e0, e1}});
  }
  void finish() {
    if (!h->stage_timing) return;
    cudaStreamSynchronize(st);
    for (auto& s : spans) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, s.second.first, s.second.second);
      h->stage_ms[s.first] += ms;
      cudaEventDestroy(s.second.first); cudaEventDestroy(s.second.second);
    }
    spans.clear();
  }
};

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

// ------------------------------------------------------------------------------------------------------------------
// schedules (host, double; numpy-equivalent formulas — data/so3_diffuser.py:183-213, r3_diffuser.py:26-30)
// ------------------------------------------------------------------------------------------------------------------
static double so3_sigma_host(double t) { return log(t * exp(SO3_MAX_SIGMA) + (1.0 - t) * exp(SO3_MIN_SIGMA)); }
static int sigma_idx_host(const std::vector<double>& grid, double t) {
  const double s = so3_sigma_host(t);
  int lo = 0, hi = (int)grid.size();
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (grid[mid] <= s) lo = mid + 1; else hi = mid; }
  int idx = lo - 1;
  if (idx < 0) idx += (int)grid.size();
  return idx;
}
static double so3_g_host(double t) {
  const double s = so3_sigma_host(t);
  return sqrt(2.0 * (exp(SO3_MAX_SIGMA) - exp(SO3_MIN_SIGMA)) * s / exp(s));
}
static double r3_b_host(double t) { return R3_MIN_B + t * (R3_MAX_B - R3_MIN_B); }

// ------------------------------------------------------------------------------------------------------------------
// create / destroy
// ------------------------------------------------------------------------------------------------------------------
static int build_igso3_rows(fd_context* h, int nrows, const int* idx, double* pdf, double* cdf, double* sn, double* scal);

extern "C" int fd_create(fd_handle* out, int device) {
  if (!out) return fail(FD_EINVAL, "fd_create: out is NULL");
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(FD_EINVAL, "fd_create: device %d out of range (%d visible)", device, ndev);
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(FD_ECUDA, "fd_create: this library is built for sm_100a (B200); device %d is sm_%d%d", device, prop.major, prop.minor);
  fd_context* h = new fd_context();
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&h->ev_fwd, cudaEventDisableTiming));
  memset(h->stage_ms, 0, sizeof(h->stage_ms));
  memset(h->stage_launches, 0, sizeof(h->stage_launches));
  // constants
  float tf[16], idn[16], dl[NBINS], pi;
  memcpy(tf, FD_TIME_FREQ_BITS, sizeof(tf)); memcpy(idn, FD_IDX_DEN_BITS, sizeof(idn));
  memcpy(dl, FD_DGRAM_LOWER_BITS, sizeof(dl)); memcpy(&pi, &FD_PI_F32_BITS, 4);
  CK(cudaMemcpyToSymbol(c_time_freq, tf, sizeof(tf)));
  CK(cudaMemcpyToSymbol(c_idx_den, idn, sizeof(idn)));
  CK(cudaMemcpyToSymbol(c_dgram_lower, dl, sizeof(dl)));
  CK(cudaMemcpyToSymbol(g_dgram_lower, dl, sizeof(dl)));
  CK(cudaMemcpyToSymbol(c_pi_f32, &pi, sizeof(pi)));
  // sigma grid + omega grid + cdf(t=1)
  h->h_sigma_grid.resize(SO3_NSIGMA);
  for (int k = 0; k < SO3_NSIGMA; ++k) {
    const double t = k == SO3_NSIGMA - 1 ? 1.0 : (double)k * (1.0 / (SO3_NSIGMA - 1));
    h->h_sigma_grid[k] = so3_sigma_host(t);
  }
  CK(cudaMalloc(&h->d_sigma_grid, SO3_NSIGMA * sizeof(double)));
  CK(cudaMemcpy(h->d_sigma_grid, h->h_sigma_grid.data(), SO3_NSIGMA * sizeof(double), cudaMemcpyHostToDevice));
  std::vector<double> om(SO3_NOMEGA);
  for (int w = 0; w < SO3_NOMEGA; ++w) om[w] = (double)(w + 1) * (3.14159265358979323846 / SO3_NOMEGA);
  CK(cudaMalloc(&h->d_omega, SO3_NOMEGA * sizeof(double)));
  CK(cudaMemcpy(h->d_omega, om.data(), SO3_NOMEGA * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&h->d_cdf_t1, SO3_NOMEGA * sizeof(double)));
  CK(cudaMalloc(&h->d_sched1, sizeof(StepSched)));
  {
    std::vector<double> cdf(SO3_NOMEGA);
    const int idx = sigma_idx_host(h->h_sigma_grid, 1.0);
    CKI(build_igso3_rows(h, 1, &idx, nullptr, cdf.data(), nullptr, nullptr));
    CK(cudaMemcpy(h->d_cdf_t1, cdf.data(), SO3_NOMEGA * sizeof(double), cudaMemcpyHostToDevice));
  }
  CK(cudaFuncSetAttribute(ipa_edge_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  CK(cudaFuncSetAttribute(ipa_edge_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  CK(cudaFuncSetAttribute(ipa_edge_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  CK(cudaFuncSetAttribute(reverse_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  if (tc_init(h->sm_count)) return fail(FD_ECUDA, "tensor-core path initialisation failed (cuTensorMapEncodeTiled entry point / smem attribute)");
  *out = h;
  return FD_OK;
}

static void free_ws(fd_context* h) {
  free_graph(h);
  if (h->ws.base) cudaFree(h->ws.base);
  h->ws = Workspace();
}
static void free_lb(fd_context* h) {
  free_graph(h);
  if (h->lb.base) cudaFree(h->lb.base);
  h->lb = LoopBufs();
}

extern "C" int fd_destroy(fd_handle h) {
  if (!h) return FD_OK;
  DevGuard dev_guard(h->device);
  cudaStreamSynchronize(h->stream);
  free_graph(h);
  free_ws(h); free_lb(h);
  for (auto& kv : h->dbg) cudaFree(kv.second.first);
  for (auto& kv : h->igso3_rows) cudaFree(kv.second.first);
  if (h->warena) cudaFree(h->warena);
  tc_free_weights(h->tcw);
  cudaFree(h->d_sigma_grid); cudaFree(h->d_cdf_t1); cudaFree(h->d_omega); cudaFree(h->d_sched1);
  if (h->d_t_tmp) cudaFree(h->d_t_tmp);
  free_train(h);
  if (h->d_loss_acc) cudaFree(h->d_loss_acc);
  if (h->ev_fwd) cudaEventDestroy(h->ev_fwd);
  cudaStreamDestroy(h->stream);
  delete h;
  return FD_OK;
}

extern "C" int fd_set_precision(fd_handle h, int prec) {
  if (!h) return fail(FD_EINVAL, "null handle");
  if (prec != FD_PREC_FP32 && prec != FD_PREC_BF16X3 && prec != FD_PREC_BF16) return fail(FD_EINVAL, "unknown precision %d", prec);
  if (prec != h->precision) { cudaStreamSynchronize(h->stream); free_ws(h); }
  h->precision = prec;
  return FD_OK;
}
extern "C" int fd_get_precision(fd_handle h) { return h ? h->precision : FD_EINVAL; }

// ------------------------------------------------------------------------------------------------------------------
// parameters
// ------------------------------------------------------------------------------------------------------------------
extern "C" int fd_num_params(void) { return (int)param_schema().size(); }
extern "C" const char* fd_param_name(int i) {
  const auto& s = param_schema();
  return (i < 0 || i >= (int)s.size()) ? nullptr : s[i].name.c_str();
}
extern "C" int fd_param_ndim(int i) {
  const auto& s = param_schema();
  return (i < 0 || i >= (int)s.size()) ? FD_EINVAL : s[i].ndim;
}
extern "C" int64_t fd_param_dim(int i, int d) {
  const auto& s = param_schema();
  if (i < 0 || i >= (int)s.size() || d < 0 || d >= s[i].ndim) return FD_EINVAL;
  return s[i].dim[d];
}
extern "C" int64_t fd_param_numel(int i) {
  const auto& s = param_schema();
  return (i < 0 || i >= (int)s.size()) ? FD_EINVAL : s[i].numel();
}

namespace {
struct Packer {   // builds one host image, records offsets, then fixes up device pointers
  std::vector<float> host;
  size_t add(const float* src, size_t n) {
    const size_t off = (host.size() + 63) & ~(size_t)63;   // 256-byte aligned
    host.resize(off + n);
    if (src) memcpy(host.data() + off, src, n * sizeof(float)); else memset(host.data() + off, 0, n * sizeof(float));
    return off;
  }
  float* at(size_t off) { return host.data() + off; }
};
}  // namespace

extern "C" int fd_load_weights(fd_handle h, const float* const* P) {
  if (!h || !P) return fail(FD_EINVAL, "fd_load_weights: null argument");
  DevGuard dev_guard(h->device);
  const auto& schema = param_schema();
  std::map<std::string, const float*> M;
  for (size_t i = 0; i < schema.size(); ++i) {
    if (!P[i]) return fail(FD_EINVAL, "fd_load_weights: parameter %zu (%s) is NULL", i, schema[i].name.c_str());
    M[schema[i].name] = P[i];
  }
  Packer pk;
  std::vector<std::pair<const float**, size_t>> fix;   // (pointer slot, offset)
  auto put = [&](const float** slot, const float* src, size_t n) { fix.push_back({slot, pk.add(src, n)}); return fix.back().second; };
  auto lin = [&](Lin& L, const std::string& n, size_t o, size_t i) {
    put(&L.w, M.at(n + ".weight"), o * i);
    put(&L.b, M.at(n + ".bias"), o);
  };
  auto lnp = [&](LNp& L, const std::string& n, size_t c) { put(&L.g, M.at(n + ".weight"), c); put(&L.b, M.at(n + ".bias"), c); };
  Weights& W = h->W;
  const std::string e = "embedding_layer.";
  {  // node embedder layer 0: K 65 -> 68
    const float* w = M.at(e + "node_embedder.0.weight");
    const size_t off = put(&W.ne0.w, nullptr, 256 * NODE_IN_PAD);
    for (int o = 0; o < 256; ++o) memcpy(pk.at(off) + o * NODE_IN_PAD, w + o * NODE_IN, NODE_IN * sizeof(float));
    put(&W.ne0.b, M.at(e + "node_embedder.0.bias"), 256);
  }
  lin(W.ne2, e + "node_embedder.2", 256, 256); lin(W.ne4, e + "node_embedder.4", 256, 256); lnp(W.ne_ln, e + "node_embedder.5", 256);
  {  // edge embedder layer 0 slices, transposed to [k][c]
    const float* w = M.at(e + "edge_embedder.0.weight");   // [128][120]
    size_t oa = put(&W.ee_w0a, nullptr, 33 * 128), oc = put(&W.ee_w0c, nullptr, 33 * 128), orr = put(&W.ee_w0r, nullptr, 32 * 128),
           od = put(&W.ee_D, nullptr, (NBINS + 1) * 128);
    for (int c = 0; c < 128; ++c) {
      for (int k = 0; k < 33; ++k) { pk.at(oa)[k * 128 + c] = w[c * EDGE_IN + k]; pk.at(oc)[k * 128 + c] = w[c * EDGE_IN + 33 + k]; }
      for (int k = 0; k < 32; ++k) pk.at(orr)[k * 128 + c] = w[c * EDGE_IN + 66 + k];
      for (int k = 0; k < NBINS; ++k) pk.at(od)[k * 128 + c] = w[c * EDGE_IN + 98 + k];
    }
    put(&W.ee_b0, M.at(e + "edge_embedder.0.bias"), 128);
  }
  lin(W.ee2, e + "edge_embedder.2", 128, 128); lin(W.ee4, e + "edge_embedder.4", 128, 128); lnp(W.ee_ln, e + "edge_embedder.5", 128);
  const size_t offT = pk.add(nullptr, (size_t)(2 * REL_DMAX + 1) * 128);
  const std::string t = "score_model.trunk.";
  for (int b = 0; b < NBLK; ++b) {
    BlockW& X = W.blk[b];
    const std::string sb = std::to_string(b), ip = t + "ipa_" + sb + ".";
    {
      const size_t ow = put(&X.proj.w, nullptr, (size_t)PROJ_ALL * C_S), ob = put(&X.proj.b, nullptr, PROJ_ALL);
      size_t r = 0;
      for (auto nm : {std::make_pair("linear_q", PROJ_Q), std::make_pair("linear_kv", PROJ_KV),
                      std::make_pair("linear_q_points", PROJ_QP), std::make_pair("linear_kv_points", PROJ_KVP)}) {
        memcpy(pk.at(ow) + r * C_S, M.at(ip + nm.first + ".weight"), (size_t)nm.second * C_S * sizeof(float));
        memcpy(pk.at(ob) + r, M.at(ip + nm.first + ".bias"), (size_t)nm.second * sizeof(float));
        r += nm.second;
      }
    }
    put(&X.Wb, M.at(ip + "linear_b.weight"), H * C_Z); put(&X.bb, M.at(ip + "linear_b.bias"), H);
    {
      const float* hw = M.at(ip + "head_weights");
      const size_t og = put(&X.gamma, nullptr, H);
      for (int k = 0; k < H; ++k) {
        const double x = hw[k];
        const double sp = x > 20.0 ? x : log1p(exp(x));   // torch.nn.Softplus (beta=1, threshold=20)
        pk.at(og)[k] = (float)((float)sp * (float)sqrt(1.0 / (3 * (PQ * 9.0 / 2))));
      }
      const float* wd = M.at(ip + "down_z.weight");   // [32][128]
      const size_t owd = put(&X.WdT, nullptr, 128 * 32);
      for (int d = 0; d < 32; ++d) for (int c = 0; c < 128; ++c) pk.at(owd)[c * 32 + d] = wd[d * 128 + c];
      put(&X.bd, M.at(ip + "down_z.bias"), 32);
      // block-diagonal image: o_pair[(h,d)] = sum_c Wd[d][c] zbar[(h,c)] for all heads in one GEMM (K = H*128, N = H*32)
      const size_t obd = put(&X.down_bd.w, nullptr, (size_t)(H * 32) * (H * C_Z)), obb = put(&X.down_bd.b, nullptr, H * 32);
      const float* bdz = M.at(ip + "down_z.bias");
      for (int hh = 0; hh < H; ++hh)
        for (int d = 0; d < 32; ++d) {
          memcpy(pk.at(obd) + (size_t)(hh * 32 + d) * (H * C_Z) + hh * C_Z, wd + d * C_Z, C_Z * sizeof(float));
          pk.at(obb)[hh * 32 + d] = bdz[d];
        }
    }
    lin(X.out, ip + "linear_out", C_S, IPA_FEAT);
    lnp(X.ipa_ln, t + "ipa_ln_" + sb, C_S);
    lin(X.skip, t + "skip_embed_" + sb, C_SKIP, C_S);
    for (int l = 0; l < TF_LAYERS; ++l) {
      const std::string p = t + "seq_tfmr_" + sb + ".layers." + std::to_string(l) + ".";
      put(&X.tf[l].in_proj.w, M.at(p + "self_attn.in_proj_weight"), (size_t)3 * TF_D * TF_D);
      put(&X.tf[l].in_proj.b, M.at(p + "self_attn.in_proj_bias"), 3 * TF_D);
      lin(X.tf[l].out_proj, p + "self_attn.out_proj", TF_D, TF_D);
      lin(X.tf[l].lin1, p + "linear1", TF_D, TF_D); lin(X.tf[l].lin2, p + "linear2", TF_D, TF_D);
      lnp(X.tf[l].norm1, p + "norm1", TF_D); lnp(X.tf[l].norm2, p + "norm2", TF_D);
    }
    lin(X.post, t + "post_tfmr_" + sb, C_S, TF_D);
    lin(X.tr1, t + "node_transition_" + sb + ".linear_1", C_S, C_S);
    lin(X.tr2, t + "node_transition_" + sb + ".linear_2", C_S, C_S);
    lin(X.tr3, t + "node_transition_" + sb + ".linear_3", C_S, C_S);
    lnp(X.tr_ln, t + "node_transition_" + sb + ".ln", C_S);
    lin(X.bbu, t + "bb_update_" + sb + ".linear", 6, C_S);
    if (b < NBLK - 1) {
      const std::string p = t + "edge_transition_" + sb + ".";
      lin(X.et_init, p + "initial_embed", C_Z, C_S);
      const float* w1 = M.at(p + "trunk.0.weight");          // [384][384]
      const float* b1 = M.at(p + "trunk.0.bias");
      const float* wf = M.at(p + "final_layer.weight");      // [128][384]
      const float* bf = M.at(p + "final_layer.bias");
      const size_t on = put(&X.et_node.w, nullptr, (size_t)ET_NODE * C_Z), onb = put(&X.et_node.b, nullptr, ET_NODE);
      const size_t o1z = put(&X.et_w1z, nullptr, (size_t)ET_HID * C_Z);
      const size_t ofz = put(&X.et_wfz, nullptr, (size_t)C_Z * C_Z);
      for (int o = 0; o < ET_HID; ++o) {
        memcpy(pk.at(o1z) + o * C_Z, w1 + o * ET_HID, C_Z * sizeof(float));
        memcpy(pk.at(on) + (size_t)o * C_Z, w1 + o * ET_HID + C_Z, C_Z * sizeof(float));                 // P rows
        memcpy(pk.at(on) + (size_t)(ET_HID + o) * C_Z, w1 + o * ET_HID + 2 * C_Z, C_Z * sizeof(float));   // Q rows
        pk.at(onb)[o] = b1[o];
      }
      for (int o = 0; o < C_Z; ++o) {
        memcpy(pk.at(ofz) + o * C_Z, wf + o * ET_HID, C_Z * sizeof(float));
        memcpy(pk.at(on) + (size_t)(2 * ET_HID + o) * C_Z, wf + o * ET_HID + C_Z, C_Z * sizeof(float));          // U rows
        memcpy(pk.at(on) + (size_t)(2 * ET_HID + C_Z + o) * C_Z, wf + o * ET_HID + 2 * C_Z, C_Z * sizeof(float));  // V rows
        pk.at(onb)[2 * ET_HID + o] = bf[o];
      }
      lin(X.et_w2, p + "trunk.2", ET_HID, ET_HID);
      put(&X.et_wfh, wf, (size_t)C_Z * ET_HID);
      lnp(X.et_ln, p + "layer_norm", C_Z);
    }
  }
  const std::string tp = "score_model.torsion_pred.";
  lin(W.tor1, tp + "linear_1", C_S, C_S); lin(W.tor2, tp + "linear_2", C_S, C_S); lin(W.torf, tp + "linear_final", 2, C_S);

  cudaStreamSynchronize(h->stream);
  free_graph(h);
  if (h->warena) { cudaFree(h->warena); h->warena = nullptr; }
  h->warena_bytes = pk.host.size() * sizeof(float);
  CK(cudaMalloc(&h->warena, h->warena_bytes));
  CK(cudaMemcpy(h->warena, pk.host.data(), h->warena_bytes, cudaMemcpyHostToDevice));
  float* dbase = reinterpret_cast<float*>(h->warena);
  for (auto& f : fix) *f.first = dbase + f.second;
  W.ee_T = dbase + offT;
  rel_table_kernel<<<2 * REL_DMAX + 1, 128, 0, h->stream>>>(W.ee_w0r, W.ee_T);
  CK(cudaGetLastError());
  // bf16 hi/lo images for the tensor-core edge kernels
  {
    std::vector<TcLinSpec> lins;
    auto reg = [&](const Lin& L, int rows, int cols) { lins.push_back({L.w, pk.host.data() + (L.w - dbase), rows, cols}); };
    reg(W.ne2, 256, 256); reg(W.ne4, 256, 256);
    for (int b = 0; b < NBLK; ++b) {
      const BlockW& X = W.blk[b];
      reg(X.proj, PROJ_ALL, C_S); reg(X.out, C_S, IPA_FEAT); reg(X.down_bd, H * 32, H * C_Z);
      for (int l = 0; l < TF_LAYERS; ++l) {
        reg(X.tf[l].in_proj, 3 * TF_D, TF_D); reg(X.tf[l].out_proj, TF_D, TF_D); reg(X.tf[l].lin1, TF_D, TF_D); reg(X.tf[l].lin2, TF_D, TF_D);
      }
      reg(X.post, C_S, TF_D); reg(X.tr1, C_S, C_S); reg(X.tr2, C_S, C_S); reg(X.tr3, C_S, C_S);
      if (b < NBLK - 1) { reg(X.et_init, C_Z, C_S); reg(X.et_node, ET_NODE, C_Z); }
    }
    reg(W.tor1, C_S, C_S); reg(W.tor2, C_S, C_S);
    if (tc_pack_weights(h->tcw, M, lins, h->stream)) return fail(FD_ECUDA, "packing bf16 weight planes failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  CK(cudaStreamSynchronize(h->stream));
  h->weights_loaded = true;
  return FD_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------------------------------
static int ensure_ws(fd_context* h, int B, int N) {
  Workspace& w = h->ws;
  if (w.base && w.B == B && w.N == N) return FD_OK;
  cudaStreamSynchronize(h->stream);
  free_ws(h);
  w.B = B; w.N = N; w.Np = (N + 3) & ~3;
  w.rows = (long long)B * N; w.edges = w.rows * N;
  const long long max_chunk = 1LL << 18;
  // chunk = whole samples when possible (keeps (b,i,j) decomposition trivial either way)
  w.chunk = w.edges < max_chunk ? w.edges : max_chunk;
  struct Item { float** p; size_t n; };
  const size_t R = (size_t)w.rows;
  const bool tc = h->precision != FD_PREC_FP32;
  std::vector<Item> items = {
      {&w.node_in, R * NODE_IN_PAD}, {&w.temb, (size_t)B * 32}, {&w.AC, R * 256}, {&w.node0, R * C_S}, {&w.node, R * C_S},
      {&w.tmpA, R * C_S}, {&w.tmpB, R * C_S}, {&w.x320, R * TF_D}, {&w.x320b, R * TF_D}, {&w.qkv, R * 3 * TF_D},
      {&w.S, (size_t)B * TF_H * N * w.Np}, {&w.y320, R * TF_D}, {&w.ff, R * TF_D}, {&w.proj, R * PROJ_ALL},
      {&w.qp, R * H * PQ * 3}, {&w.kp, R * H * PQ * 3}, {&w.vp, R * H * PV * 3}, {&w.optg, R * H * PV * 3}, {&w.zbar, tc ? R * H * C_Z : 0},
      {&w.L, (size_t)B * H * N * w.Np}, {&w.feats, R * IPA_FEAT}, {&w.quat, R * 4}, {&w.trans, R * 3}, {&w.nb, R * C_Z},
      {&w.pquv, R * ET_NODE}, {&w.z, (size_t)w.edges * C_Z}, {&w.tors, R * C_S},
      {&w.h1, tc ? 0 : (size_t)w.chunk * ET_HID}, {&w.h2, tc ? 0 : (size_t)w.chunk * ET_HID},
      {&w.ychunk, tc ? 0 : (size_t)w.chunk * C_Z}};
  size_t total = 0;
  for (auto& it : items) total += al256(it.n * sizeof(float));
  const size_t tc_bytes = tc ? tc_workspace_bytes(B, N) : 0;
  total += al256(tc_bytes);
  if (cudaMalloc(&w.base, total) != cudaSuccess) {
    cudaGetLastError();
    w = Workspace();
    return fail(FD_ENOMEM, "workspace allocation of %.1f MB failed for B=%d N=%d", total / 1048576.0, B, N);
  }
  w.bytes = total;
  char* p = w.base;
  for (auto& it : items) { *it.p = reinterpret_cast<float*>(p); p += al256(it.n * sizeof(float)); }
  if (tc && tc_bind_workspace(w.tc, p, B, N)) return fail(FD_ECUDA, "cuTensorMapEncodeTiled failed for the edge-tensor planes");
  return FD_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// debug taps
// ------------------------------------------------------------------------------------------------------------------
static void snap(fd_context* h, const std::string& name, const void* src, size_t bytes, cudaStream_t st) {
  if (!h->debug) return;
  auto it = h->dbg.find(name);
  if (it == h->dbg.end() || it->second.second != bytes) {
    if (it != h->dbg.end()) cudaFree(it->second.first);
    void* p = nullptr;
    cudaMalloc(&p, bytes);
    h->dbg[name] = {p, bytes};
    it = h->dbg.find(name);
  }
  cudaMemcpyAsync(it->second.first, src, bytes, cudaMemcpyDeviceToDevice, st);
}
extern "C" int fd_set_debug(fd_handle h, int on) { if (!h) return FD_EINVAL; h->debug = on != 0; return FD_OK; }
extern "C" int64_t fd_debug_fetch(fd_handle h, const char* name, void* dst, int64_t dst_bytes) {
  if (!h || !name) return fail(FD_EINVAL, "fd_debug_fetch: null argument");
  auto it = h->dbg.find(name);
  if (it == h->dbg.end()) return fail(FD_EINVAL, "fd_debug_fetch: no tap named '%s'", name);
  if (!dst) return (int64_t)it->second.second;
  if (dst_bytes < (int64_t)it->second.second) return fail(FD_EINVAL, "fd_debug_fetch: buffer too small");
  cudaDeviceSynchronize();
  if (cudaMemcpy(dst, it->second.first, it->second.second, cudaMemcpyDeviceToHost) != cudaSuccess) return fail(FD_ECUDA, "memcpy failed");
  return (int64_t)it->second.second;
}

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
namespace {
struct Fwd {
  fd_context* h; cudaStream_t st; Workspace& w; const Weights& W;
  int err = 0;
  void gemm(GemmArgs g, bool kmajor = true) {
    if (err) return;
    cudaError_t e = launch_gemm(g, kmajor, st);
    h->launches++;
    if (e != cudaSuccess) err = fail(FD_ECUDA, "gemm launch failed: %s", cudaGetErrorString(e));
  }
  // y[M,N] = act(x[M,K] · W^T + b) (+ residual)
  void linear(const float* x, int ldx, const Lin& L, int K, int Nout, float* y, int ldy, long long M, bool relu = false,
              const float* residual = nullptr, int ldr = 0, const float* rowmask = nullptr) {
    if (err) return;
    if (h->precision != FD_PREC_FP32) {   // node-path linears on the tensor cores (split-bf16 operands, fp32 accumulate/output)
      const int rc = tc_linear(h->tcw, w.tc, h->precision, x, ldx, L.w, L.b, K, Nout, y, ldy, M, relu, residual, ldr, rowmask, st, &h->launches);
      if (rc == 0) return;
      if (rc < 0) { err = fail(FD_ECUDA, "tensor-core linear launch failed: %s", cudaGetErrorString(cudaGetLastError())); return; }
    }
    GemmArgs g;
    g.A = x; g.lda = ldx; g.B = L.w; g.ldb = K; g.C = y; g.ldc = ldy; g.M = (int)M; g.N = Nout; g.K = K;
    g.bias = L.b; g.relu = relu; g.residual = residual; g.ldr = ldr; g.rowmask = rowmask;
    gemm(g);
  }
  void ln(int C, const float* x, int ldx, float* out, int ldo, const LNp& p, long long M, const float* rowmask = nullptr,
          float* out2 = nullptr, int ldo2 = 0) {
    if (err) return;
    LnArgs a;
    a.x = x; a.ldx = ldx; a.out = out; a.ldo = ldo; a.out2 = out2; a.ldo2 = ldo2; a.gamma = p.g; a.beta = p.b; a.M = M;
    a.rowmask = rowmask;
    cudaError_t e = launch_layernorm(C, a, st);
    h->launches++;
    if (e != cudaSuccess) err = fail(FD_ECUDA, "layernorm launch failed: %s", cudaGetErrorString(e));
  }
  void check(const char* what) {
    if (err) return;
    cudaError_t e = cudaGetLastError();
    h->launches++;
    if (e != cudaSuccess) err = fail(FD_ECUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
  }
};
}  // namespace

static int forward_impl(fd_context* h, int B, int N, const float* rigids_t, const double* t_dev, int t_is_f32,
                        const double* sigma_dev, const float* res_mask, const float* fixed_mask, const int* seq_idx,
                        const float* sc_ca, const float* gt_psi, const fd_forward_out* out, float* sc_ca_out,
                        cudaStream_t st, const double* cached_rows = nullptr) {
  CKI(ensure_ws(h, B, N));
  Workspace& w = h->ws;
  const Weights& W = h->W;
  Fwd f{h, st, w, W};
  Launcher lc{h, st};
  const long long R = w.rows;
  const int Np = w.Np;
  const bool tc = h->precision != FD_PREC_FP32;

  // ---- embedder: node (model/score_network.py:103-151) ----------------------------------------------------------------
  lc.begin(ST_EMBED_NODE);
  node_feats_kernel<<<(unsigned)((R * 16 + 255) / 256), 256, 0, st>>>(t_dev, t_is_f32, fixed_mask, seq_idx, w.node_in, w.temb, B, N);
  f.check("node_feats");
  f.linear(w.node_in, NODE_IN_PAD, W.ne0, NODE_IN_PAD, 256, w.tmpA, 256, R, true);
  f.linear(w.tmpA, 256, W.ne2, 256, 256, w.tmpB, 256, R, true);
  f.linear(w.tmpB, 256, W.ne4, 256, 256, w.tmpA, 256, R);
  f.ln(256, w.tmpA, 256, w.node0, 256, W.ne_ln, R, res_mask, w.node, 256);     // node0 = node = LN(..)·mask
  lc.end();
  snap(h, "node_embed", w.node0, R * C_S * 4, st);

  // ---- embedder: edge ---------------------------------------------------------------------------------------------
  lc.begin(ST_EMBED_EDGE);
  edge_l0_node_terms_kernel<<<(unsigned)((R * 256 + 255) / 256), 256, 0, st>>>(w.temb, fixed_mask, W.ee_w0a, W.ee_w0c, W.ee_b0, w.AC, R, N);
  f.check("edge_l0_node_terms");
  if (tc) {
    if (!f.err) {
      const int rc = tc_edge_embed(h->tcw, w.tc, h->precision, w.AC, W.ee_T, W.ee_D, W.ee_w0r, seq_idx, sc_ca, res_mask, W.ee2.b, W.ee4.b,
                                   W.ee_ln.g, W.ee_ln.b, B, N, st, &h->launches);
      if (rc) f.err = fail(rc == -1 ? FD_EINVAL : FD_ECUDA, "tensor-core edge embedder launch failed (%d): %s", rc, cudaGetErrorString(cudaGetLastError()));
    }
  } else {
    for (long long r0 = 0; r0 < w.edges && !f.err; r0 += w.chunk) {
      const long long m = (w.edges - r0 < w.chunk) ? w.edges - r0 : w.chunk;
      edge_embed_l0_kernel<0><<<(unsigned)((m + 7) / 8), 256, 0, st>>>(w.AC, W.ee_T, W.ee_D, W.ee_w0r, seq_idx, sc_ca, w.h1, nullptr, nullptr, r0, m, N);
      f.check("edge_embed_l0");
      f.linear(w.h1, 128, W.ee2, 128, 128, w.h2, 128, m, true);
      f.linear(w.h2, 128, W.ee4, 128, 128, w.ychunk, 128, m);
      LnArgs a;
      a.x = w.ychunk; a.ldx = 128; a.out = w.z + r0 * C_Z; a.ldo = 128; a.gamma = W.ee_ln.g; a.beta = W.ee_ln.b; a.M = m;
      a.res_mask = res_mask; a.nres = N; a.row_offset = r0;
      if (!f.err && launch_layernorm(128, a, st) != cudaSuccess) f.err = fail(FD_ECUDA, "edge LN launch failed");
      h->launches++;
    }
  }
  lc.end();
  if (h->debug && !f.err) {
    if (tc) tc_export_z(w.tc, w.z, h->precision, st);
    snap(h, "edge_embed", w.z, (size_t)w.edges * C_Z * 4, st);
  }

  init_frames_kernel<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(rigids_t, w.quat, w.trans, R);
  f.check("init_frames");

  for (int b = 0; b < NBLK && !f.err; ++b) {
    const BlockW& X = W.blk[b];
    const std::string sb = std::to_string(b);
    // ---- IPA (model/ipa_pytorch.py:303-471) -----------------------------------------------------------------------
    lc.begin(ST_IPA_PROJ);
    f.linear(w.node, 256, X.proj, 256, PROJ_ALL, w.proj, PROJ_ALL, R);
    ipa_points_kernel<<<(unsigned)R, 224, 0, st>>>(w.proj, w.quat, w.trans, w.qp, w.kp, w.vp, R);
    f.check("ipa_points");
    lc.end();
    lc.begin(ST_IPA_LOGITS);
    if (tc) {
      if (!f.err && tc_ipa_logits(w.tc, w.proj, w.qp, w.kp, X.gamma, w.L, B, N, Np, (float)sqrt(1.0 / (3 * C_HID)), st, &h->launches))
        f.err = fail(FD_ECUDA, "ipa_logits (tensor-core) launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    } else {
      GemmArgs g;
      g.A = w.proj; g.lda = PROJ_ALL; g.sA0 = (long long)N * PROJ_ALL; g.sA1 = C_HID;
      g.B = w.proj + PROJ_Q; g.ldb = PROJ_ALL; g.sB0 = (long long)N * PROJ_ALL; g.sB1 = 2 * C_HID;
      g.C = w.L; g.ldc = Np; g.sC0 = (long long)H * N * Np; g.sC1 = (long long)N * Np;
      g.M = N; g.N = N; g.K = C_HID; g.nb0 = B; g.nb1 = H; g.alpha = (float)sqrt(1.0 / (3 * C_HID));
      f.gemm(g);
    }
    lc.end();
    lc.begin(ST_IPA_EDGE);
    if (!f.err) {
      const size_t smem = (size_t)(H * Np + H * PQ * 3 + 2 * H * C_Z) * sizeof(float);
      if (tc) {
        const bool one_kernel = tc_ipa_edge_fused_ok(N);
        if (tc_ipa_edge(h->tcw, w.tc, b, w.L, w.qp, w.kp, res_mask, X.Wb, X.bb, X.gamma, X.WdT, X.bd, w.feats, w.zbar, B, N, Np, h->precision, h->debug ? 1 : 0, st, &h->launches))
          f.err = fail(FD_ECUDA, "ipa_edge (planes) launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (one_kernel) f.linear(w.zbar, H * C_Z, X.down_bd, H * C_Z, H * 32, w.feats + (H * C_HID + 4 * H * PV), IPA_FEAT, R);   // o_pair
      } else {
        ZRef zr; zr.f32 = w.z;
        ipa_edge_kernel<0><<<dim3(N, B), 256, smem, st>>>(zr, w.L, w.qp, w.kp, res_mask, X.Wb, X.bb, X.gamma, X.WdT, X.bd, w.feats, N, Np);
        f.check("ipa_edge");
      }
    }
    lc.end();
    if (h->debug) snap(h, "attn_" + sb, w.L, (size_t)B * H * N * Np * 4, st);
    lc.begin(ST_IPA_AV);
    {
      GemmArgs g;   // o = a · v
      g.A = w.L; g.lda = Np; g.sA0 = (long long)H * N * Np; g.sA1 = (long long)N * Np;
      g.B = w.proj + PROJ_Q + C_HID; g.ldb = PROJ_ALL; g.sB0 = (long long)N * PROJ_ALL; g.sB1 = 2 * C_HID;
      g.C = w.feats; g.ldc = IPA_FEAT; g.sC0 = (long long)N * IPA_FEAT; g.sC1 = C_HID;
      g.M = N; g.N = C_HID; g.K = N; g.nb0 = B; g.nb1 = H;
      GemmArgs p = g;   // o_pt (global) = a · v_pts
      p.B = w.vp; p.ldb = H * PV * 3; p.sB0 = (long long)N * H * PV * 3; p.sB1 = PV * 3;
      p.C = w.optg; p.ldc = H * PV * 3; p.sC0 = (long long)N * H * PV * 3; p.sC1 = PV * 3; p.N = PV * 3;
      if (tc) {
        if (!f.err && tc_ipa_av(w.tc, w.proj, w.vp, w.L, w.feats, w.optg, B, N, Np, st, &h->launches))
          f.err = fail(FD_ECUDA, "ipa a.v (tensor-core) launch failed: %s", cudaGetErrorString(cudaGetLastError()));
      } else {
        f.gemm(g, false);
        f.gemm(p, false);
      }
      ipa_finish_kernel<<<(unsigned)R, 96, 0, st>>>(w.optg, w.quat, w.trans, w.feats, R);
      f.check("ipa_finish");
    }
    lc.end();
    if (h->debug) snap(h, "ipa_feats_" + sb, w.feats, R * IPA_FEAT * 4, st);
    lc.begin(ST_IPA_OUT);
    // node = LN(node + mask·linear_out(feats))          (ipa_pytorch.py:626-632)
    f.linear(w.feats, IPA_FEAT, X.out, IPA_FEAT, C_S, w.tmpA, C_S, R, false, w.node, C_S, res_mask);
    f.ln(256, w.tmpA, 256, w.node, 256, X.ipa_ln, R, nullptr, w.x320, TF_D);
    lc.end();
    // ---- sequence transformer (ipa_pytorch.py:633-638) --------------------------------------------------------------
    lc.begin(ST_NODE_TFMR);
    f.linear(w.node0, 256, X.skip, 256, C_SKIP, w.x320 + C_S, TF_D, R);
    float* x = w.x320; float* xo = w.x320b;
    for (int l = 0; l < TF_LAYERS; ++l) {
      const TfLayer& T = X.tf[l];
      f.linear(x, TF_D, T.in_proj, TF_D, 3 * TF_D, w.qkv, 3 * TF_D, R);
      {
        GemmArgs g;   // S = q k^T / sqrt(dh)
        g.A = w.qkv; g.lda = 3 * TF_D; g.sA0 = (long long)N * 3 * TF_D; g.sA1 = TF_DH;
        g.B = w.qkv + TF_D; g.ldb = 3 * TF_D; g.sB0 = (long long)N * 3 * TF_D; g.sB1 = TF_DH;
        g.C = w.S; g.ldc = Np; g.sC0 = (long long)TF_H * N * Np; g.sC1 = (long long)N * Np;
        g.M = N; g.N = N; g.K = TF_DH; g.nb0 = B; g.nb1 = TF_H; g.alpha = (float)(1.0 / sqrt((double)TF_DH));
        const bool one_kernel = tc && tc_tf_attn_fused_ok(N);
        if (one_kernel) {
          if (!f.err && tc_tf_attention(w.qkv, res_mask, w.y320, B, N, st, &h->launches))
            f.err = fail(FD_ECUDA, "transformer attention launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        } else if (tc) {
          if (!f.err && tc_tf_logits(w.tc, w.qkv, w.S, B, N, Np, st, &h->launches))
            f.err = fail(FD_ECUDA, "transformer logits (tensor-core) launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        } else {
          f.gemm(g);
        }
        if (!one_kernel) {
        const long long rows = (long long)B * TF_H * N;
        softmax_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(w.S, Np, N, rows, (long long)TF_H * N, res_mask);
        f.check("softmax_rows");
        GemmArgs v;   // y = S v
        v.A = w.S; v.lda = Np; v.sA0 = (long long)TF_H * N * Np; v.sA1 = (long long)N * Np;
        v.B = w.qkv + 2 * TF_D; v.ldb = 3 * TF_D; v.sB0 = (long long)N * 3 * TF_D; v.sB1 = TF_DH;
        v.C = w.y320; v.ldc = TF_D; v.sC0 = (long long)N * TF_D; v.sC1 = TF_DH;
        v.M = N; v.N = TF_DH; v.K = N; v.nb0 = B; v.nb1 = TF_H;
        if (tc) {
          if (!f.err && tc_tf_values(w.tc, w.qkv, w.S, w.y320, B, N, Np, st, &h->launches))
            f.err = fail(FD_ECUDA, "transformer values (tensor-core) launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        } else {
          f.gemm(v, false);
        }
        }
      }
      f.linear(w.y320, TF_D, T.out_proj, TF_D, TF_D, w.ff, TF_D, R, false, x, TF_D);       // x + attn
      f.ln(320, w.ff, TF_D, xo, TF_D, T.norm1, R);
      f.linear(xo, TF_D, T.lin1, TF_D, TF_D, w.y320, TF_D, R, true);
      f.linear(w.y320, TF_D, T.lin2, TF_D, TF_D, w.ff, TF_D, R, false, xo, TF_D);          // x + ff
      f.ln(320, w.ff, TF_D, x, TF_D, T.norm2, R, l == TF_LAYERS - 1 ? res_mask : nullptr);  // padded rows -> 0 at the end
    }
    f.linear(x, TF_D, X.post, TF_D, C_S, w.tmpA, C_S, R, false, w.node, C_S);               // node + post_tfmr(..)
    lc.end();
    // ---- node transition (ipa_pytorch.py:169-191,639-640) ----------------------------------------------------------------
    lc.begin(ST_NODE_TRANS);
    f.linear(w.tmpA, C_S, X.tr1, C_S, C_S, w.tmpB, C_S, R, true);
    f.linear(w.tmpB, C_S, X.tr2, C_S, C_S, w.node, C_S, R, true);
    f.linear(w.node, C_S, X.tr3, C_S, C_S, w.tmpB, C_S, R, false, w.tmpA, C_S);
    f.ln(256, w.tmpB, 256, w.node, 256, X.tr_ln, R, res_mask);
    backbone_update_kernel<<<(unsigned)((R + 7) / 8), 256, 0, st>>>(w.node, X.bbu.w, X.bbu.b, res_mask, fixed_mask, w.quat, w.trans, R);
    f.check("backbone_update");
    lc.end();
    if (h->debug) {
      snap(h, "node_" + sb, w.node, R * C_S * 4, st);
      snap(h, "quat_" + sb, w.quat, R * 16, st);
      snap(h, "trans_" + sb, w.trans, R * 12, st);
    }
    // ---- edge transition (ipa_pytorch.py:194-233,646-649) ----------------------------------------------------------------
    if (b < NBLK - 1) {
      lc.begin(ST_EDGE_TRANS);
      f.linear(w.node, C_S, X.et_init, C_S, C_Z, w.nb, C_Z, R);
      f.linear(w.nb, C_Z, X.et_node, C_Z, ET_NODE, w.pquv, ET_NODE, R);
      if (tc) {
        if (!f.err) {
          const int rc = tc_edge_transition(h->tcw, w.tc, b, h->precision, w.pquv, X.et_w2.b, X.et_ln.g, X.et_ln.b, res_mask, B, N, st, &h->launches);
          if (rc) f.err = fail(rc == -1 ? FD_EINVAL : FD_ECUDA, "tensor-core edge transition launch failed (%d): %s", rc, cudaGetErrorString(cudaGetLastError()));
        }
      } else {
        for (long long r0 = 0; r0 < w.edges && !f.err; r0 += w.chunk) {
          const long long m = (w.edges - r0 < w.chunk) ? w.edges - r0 : w.chunk;
          GemmArgs g;   // h1 = relu(z·W1z^T + P_i + Q_j)
          g.A = w.z + r0 * C_Z; g.lda = C_Z; g.B = X.et_w1z; g.ldb = C_Z; g.C = w.h1; g.ldc = ET_HID; g.M = (int)m; g.N = ET_HID; g.K = C_Z;
          g.relu = 1; g.rowadd_i = w.pquv; g.rowadd_j = w.pquv + ET_HID; g.ld_rowadd = ET_NODE; g.nres = N; g.row_offset = r0;
          f.gemm(g);
          f.linear(w.h1, ET_HID, X.et_w2, ET_HID, ET_HID, w.h2, ET_HID, m, true);
          GemmArgs y;   // y = z·Wfz^T + U_i + V_j
          y.A = w.z + r0 * C_Z; y.lda = C_Z; y.B = X.et_wfz; y.ldb = C_Z; y.C = w.ychunk; y.ldc = C_Z; y.M = (int)m; y.N = C_Z; y.K = C_Z;
          y.rowadd_i = w.pquv + 2 * ET_HID; y.rowadd_j = w.pquv + 2 * ET_HID + C_Z; y.ld_rowadd = ET_NODE; y.nres = N; y.row_offset = r0;
          f.gemm(y);
          GemmArgs y2;  // y += h2·Wf^T
          y2.A = w.h2; y2.lda = ET_HID; y2.B = X.et_wfh; y2.ldb = ET_HID; y2.C = w.ychunk; y2.ldc = C_Z; y2.M = (int)m; y2.N = C_Z; y2.K = ET_HID;
          y2.accumulate = 1;
          f.gemm(y2);
          LnArgs a;
          a.x = w.ychunk; a.ldx = C_Z; a.out = w.z + r0 * C_Z; a.ldo = C_Z; a.gamma = X.et_ln.g; a.beta = X.et_ln.b; a.M = m;
          a.res_mask = res_mask; a.nres = N; a.row_offset = r0;
          if (!f.err && launch_layernorm(128, a, st) != cudaSuccess) f.err = fail(FD_ECUDA, "edge LN launch failed");
          h->launches++;
        }
      }
      lc.end();
      if (h->debug && !f.err) {
        if (tc) tc_export_z(w.tc, w.z, h->precision, st);
        snap(h, "edge_" + sb, w.z, (size_t)w.edges * C_Z * 4, st);
      }
    }
  }
  // ---- heads ---------------------------------------------------------------------------------------------------------
  lc.begin(ST_HEADS);
  f.linear(w.node, C_S, W.tor1, C_S, C_S, w.tmpA, C_S, R, true);
  f.linear(w.tmpA, C_S, W.tor2, C_S, C_S, w.tors, C_S, R, false, w.node, C_S);
  if (!f.err) {
    HeadArgs a;
    a.tors_s = w.tors; a.Wf = W.torf.w; a.bf = W.torf.b; a.quat = w.quat; a.trans = w.trans; a.rigids_t = rigids_t;
    a.t = t_dev; a.t_is_f32 = t_is_f32; a.sigma = sigma_dev; a.sigma_grid = h->d_sigma_grid;
    a.res_mask = res_mask; a.fixed_mask = fixed_mask; a.gt_psi = gt_psi; a.cached_rows = cached_rows; a.omega_grid = h->d_omega;
    a.rot_score = out->rot_score; a.trans_score = out->trans_score; a.psi = out->psi; a.rigids = out->rigids;
    a.atom37 = out->atom37; a.atom14 = out->atom14; a.sc_ca = sc_ca_out; a.rows = R; a.N = N;
    score_head_kernel<<<(unsigned)((R + 7) / 8), 256, 0, st>>>(a);
    f.check("score_head");
  }
  lc.end();
  lc.finish();
  return f.err;
}

extern "C" int fd_forward(fd_handle h, int B, int N, const fd_forward_in* in, const fd_forward_out* out, void* stream) {
  if (!h || !in || !out) return fail(FD_EINVAL, "fd_forward: null argument");
  if (!h->weights_loaded) return fail(FD_ESTATE, "fd_forward: weights not loaded");
  if (B <= 0 || N <= 0) return fail(FD_EINVAL, "fd_forward: B=%d N=%d", B, N);
  if (!in->rigids_t || !in->t || !in->res_mask || !in->fixed_mask || !in->seq_idx || !in->sc_ca_t)
    return fail(FD_EINVAL, "fd_forward: a required input pointer is NULL");
  if (!out->rot_score || !out->trans_score || !out->psi || !out->rigids) return fail(FD_EINVAL, "fd_forward: a required output pointer is NULL");
  if (out->rigids == in->rigids_t) return fail(FD_EINVAL, "fd_forward: out->rigids must not alias in->rigids_t");
  DevGuard dev_guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int rc = forward_impl(h, B, N, in->rigids_t, in->t, in->t_is_f32, in->sigma, in->res_mask, in->fixed_mask, in->seq_idx, in->sc_ca_t,
                              in->gt_psi, out, nullptr, st, in->cached_score_rows);
  if (rc == FD_OK && st != h->stream) { CK(cudaEventRecord(h->ev_fwd, st)); h->fwd_pending = true; }
  return rc;
}

// ------------------------------------------------------------------------------------------------------------------
// diffuser pieces
// ------------------------------------------------------------------------------------------------------------------
extern "C" int fd_igso3_score(fd_handle h, int64_t n, const float* vec, const double* sigma, double* score_out, void* stream) {
  if (!h || !vec || !sigma || !score_out || n <= 0) return fail(FD_EINVAL, "fd_igso3_score: bad argument");
  DevGuard dev_guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  igso3_score_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(vec, sigma, score_out, n);
  CK(cudaGetLastError());
  return FD_OK;
}

static int build_igso3_rows(fd_context* h, int nrows, const int* idx, double* pdf, double* cdf, double* sn, double* scal) {
  std::vector<double> sig(nrows);
  for (int r = 0; r < nrows; ++r) {
    if (idx[r] < 0 || idx[r] >= SO3_NSIGMA) return fail(FD_EINVAL, "sigma index %d out of range", idx[r]);
    sig[r] = h->h_sigma_grid[idx[r]];
  }
  double *d_sig, *d_exp, *d_ds, *d_pdf, *d_cdf, *d_sn, *d_sc;
  const size_t rw = (size_t)nrows * SO3_NOMEGA * sizeof(double);
  CK(cudaMalloc(&d_sig, nrows * sizeof(double)));
  CK(cudaMalloc(&d_exp, rw)); CK(cudaMalloc(&d_ds, rw)); CK(cudaMalloc(&d_pdf, rw)); CK(cudaMalloc(&d_cdf, rw)); CK(cudaMalloc(&d_sn, rw));
  CK(cudaMalloc(&d_sc, nrows * sizeof(double)));
  CK(cudaMemcpyAsync(d_sig, sig.data(), nrows * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  igso3_series_kernel<<<dim3(SO3_NOMEGA, nrows), 128, 0, h->stream>>>(d_sig, d_exp, d_ds);
  igso3_rows_post_kernel<<<(nrows + 63) / 64, 64, 0, h->stream>>>(d_exp, d_ds, d_pdf, d_cdf, d_sn, d_sc, nrows);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  if (pdf) CK(cudaMemcpy(pdf, d_pdf, rw, cudaMemcpyDeviceToHost));
  if (cdf) CK(cudaMemcpy(cdf, d_cdf, rw, cudaMemcpyDeviceToHost));
  if (sn) CK(cudaMemcpy(sn, d_sn, rw, cudaMemcpyDeviceToHost));
  if (scal) CK(cudaMemcpy(scal, d_sc, nrows * sizeof(double), cudaMemcpyDeviceToHost));
  cudaFree(d_sig); cudaFree(d_exp); cudaFree(d_ds); cudaFree(d_pdf); cudaFree(d_cdf); cudaFree(d_sn); cudaFree(d_sc);
  return FD_OK;
}

extern "C" int fd_igso3_tables_host(fd_handle h, int nrows, const int32_t* sigma_idx, double* pdf, double* cdf, double* score_norms,
                                    double* score_scaling) {
  if (!h || nrows <= 0 || !sigma_idx) return fail(FD_EINVAL, "fd_igso3_tables_host: bad argument");
  DevGuard dev_guard(h->device);
  return build_igso3_rows(h, nrows, sigma_idx, pdf, cdf, score_norms, score_scaling);
}

extern "C" int fd_sample_ref(fd_handle h, int64_t n, const double* z_axis, const double* u_angle, const double* z_trans, uint64_t seed,
                             int64_t first_sample, int residues_per_sample, float* rigids_out, void* stream) {
  if (!h || n <= 0 || !rigids_out) return fail(FD_EINVAL, "fd_sample_ref: bad argument");
  if ((z_axis || u_angle || z_trans) && !(z_axis && u_angle && z_trans)) return fail(FD_EINVAL, "fd_sample_ref: inject all three noise arrays or none");
  if (!z_axis && residues_per_sample <= 0) return fail(FD_EINVAL, "fd_sample_ref: residues_per_sample must be > 0");
  DevGuard dev_guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  sample_ref_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(z_axis, u_angle, z_trans, seed, first_sample, residues_per_sample,
                                                                 h->d_cdf_t1, h->d_omega, rigids_out, n);
  CK(cudaGetLastError());
  return FD_OK;
}

static int launch_reverse(fd_context* h, ReverseArgs a, int B, cudaStream_t st) {
  const size_t smem = (size_t)a.N * 3 * sizeof(double);
  if (smem > 160 * 1024) return fail(FD_EINVAL, "reverse step: N=%d too large", a.N);
  reverse_step_kernel<<<B, 256, smem, st>>>(a);
  h->launches++;
  CK(cudaGetLastError());
  return FD_OK;
}

extern "C" int fd_reverse_step(fd_handle h, int B, int N, float* rigids_io, const double* rot_score, const double* trans_score,
                               const float* diffuse_mask, double t, double dt, int center, double noise_scale, const double* z_rot,
                               const double* z_trans, uint64_t seed, int64_t first_sample, int step, float* rotmat_out, void* stream) {
  if (!h || !rigids_io || !rot_score || !trans_score || B <= 0 || N <= 0) return fail(FD_EINVAL, "fd_reverse_step: bad argument");
  if (!(t >= 0.0 && t <= 1.0)) return fail(FD_EINVAL, "Invalid t=%g", t);   // so3_diffuser.py:194 / r3_diffuser.py:27
  if ((z_rot == nullptr) != (z_trans == nullptr)) return fail(FD_EINVAL, "fd_reverse_step: inject both noise arrays or none");
  DevGuard dev_guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  StepSched sc{t, so3_g_host(t), r3_b_host(t), dt};
  CK(cudaMemcpyAsync(h->d_sched1, &sc, sizeof(sc), cudaMemcpyHostToDevice, st));
  ReverseArgs a{};
  a.rigids = rigids_io; a.rot_score = rot_score; a.trans_score = trans_score; a.diffuse_mask = diffuse_mask; a.use_masks = 0;
  a.z_rot = z_rot; a.z_trans = z_trans; a.sched = h->d_sched1; a.step_ptr = nullptr; a.step_fixed = 0;
  a.noise_stride = 0; a.rng_step_bias = step;   // single-entry schedule at index 0; Philox counters use the caller's step
  a.seed = seed; a.first_sample = first_sample; a.center = center; a.noise_scale = noise_scale; a.rotmat_out = rotmat_out; a.N = N;
  return launch_reverse(h, a, B, st);
}

// SE3Diffuser.forward_marginal (data/se3_diffuser.py:43-110) for n residues of one example.  Device pointers; noise injected
// (the reference draws randn(n,3), rand(n) [SO3Diffuser.sample] then normal(n,3) [R3Diffuser.forward_marginal] from np.random).
extern "C" int fd_forward_marginal(fd_handle h, int64_t n, const float* rigids_0, double t, const double* z_axis, const double* u_angle,
                                   const double* z_trans, const float* diffuse_mask, float* rigids_t, double* rot_score, double* trans_score,
                                   double* rot_score_scaling, double* trans_score_scaling, void* stream) {
  if (!h || n <= 0 || !rigids_0 || !z_axis || !u_angle || !z_trans || !rigids_t || !rot_score || !trans_score)
    return fail(FD_EINVAL, "fd_forward_marginal: bad argument");
  if (!(t >= 0.0 && t <= 1.0)) return fail(FD_EINVAL, "Invalid t=%g", t);
  DevGuard dev_guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int idx = sigma_idx_host(h->h_sigma_grid, t);
  // cdf row + score scaling for this sigma index: built on the GPU once and cached
  auto it = h->igso3_rows.find(idx);
  if (it == h->igso3_rows.end()) {
    std::vector<double> cdf(SO3_NOMEGA);
    double scal = 0;
    CKI(build_igso3_rows(h, 1, &idx, nullptr, cdf.data(), nullptr, &scal));
    double* d = nullptr;
    CK(cudaMalloc(&d, SO3_NOMEGA * sizeof(double)));
    CK(cudaMemcpy(d, cdf.data(), SO3_NOMEGA * sizeof(double), cudaMemcpyHostToDevice));
    it = h->igso3_rows.emplace(idx, std::make_pair(d, scal)).first;
  }
  forward_marginal_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(rigids_0, z_axis, u_angle, z_trans, diffuse_mask, t, h->h_sigma_grid[idx],
                                                                 it->second.first, h->d_omega, rigids_t, rot_score, trans_score, n);
  CK(cudaGetLastError());
  if (rot_score_scaling) *rot_score_scaling = it->second.second;
  if (trans_score_scaling) {
    const double beta = t * R3_MIN_B + 0.5 * (t * t) * (R3_MAX_B - R3_MIN_B);
    *trans_score_scaling = 1.0 / sqrt(1.0 - exp(-beta));
  }
  return FD_OK;
}

// Batched, padded training-data assembly: forward_marginal of B examples at their own times t[b] in one call, with the
// zero padding of du.pad_feats (data/utils.py:387-399) — what data/pdb_data_loader.py:251-272 + length_batching do per example on CPU workers.
extern "C" int fd_forward_marginal_batch(fd_handle h, int B, int N, const float* rigids_0, const double* t_host, const double* z_axis,
                                         const double* u_angle, const double* z_trans, const float* res_mask, float* rigids_t, double* rot_score,
                                         double* trans_score, double* rot_score_scaling_host, double* trans_score_scaling_host, void* stream) {
  if (!h || B < 1 || N < 1 || !rigids_0 || !t_host || !z_axis || !u_angle || !z_trans || !res_mask || !rigids_t || !rot_score || !trans_score)
    return fail(FD_EINVAL, "fd_forward_marginal_batch: bad argument");
  for (int b = 0; b < B; ++b)
    if (!(t_host[b] >= 0.0 && t_host[b] <= 1.0)) return fail(FD_EINVAL, "Invalid t=%g", t_host[b]);
  DevGuard dev_guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  for (int b = 0; b < B; ++b) {
    const double t = t_host[b];
    const int idx = sigma_idx_host(h->h_sigma_grid, t);
    auto it = h->igso3_rows.find(idx);
    if (it == h->igso3_rows.end()) {
      std::vector<double> cdf(SO3_NOMEGA);
      double scal = 0;
      CKI(build_igso3_rows(h, 1, &idx, nullptr, cdf.data(), nullptr, &scal));
      double* d = nullptr;
      CK(cudaMalloc(&d, SO3_NOMEGA * sizeof(double)));
      CK(cudaMemcpy(d, cdf.data(), SO3_NOMEGA * sizeof(double), cudaMemcpyHostToDevice));
      it = h->igso3_rows.emplace(idx, std::make_pair(d, scal)).first;
    }
    const long long o = (long long)b * N;
    forward_marginal_kernel<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(rigids_0 + o * 7, z_axis + o * 3, u_angle + o, z_trans + o * 3, res_mask + o, t,
                                                                   h->h_sigma_grid[idx], it->second.first, h->d_omega, rigids_t + o * 7,
                                                                   rot_score + o * 3, trans_score + o * 3, N, res_mask + o);
    CK(cudaGetLastError());
    if (rot_score_scaling_host) rot_score_scaling_host[b] = it->second.second;
    if (trans_score_scaling_host) {
      const double beta = t * R3_MIN_B + 0.5 * (t * t) * (R3_MAX_B - R3_MIN_B);
      trans_score_scaling_host[b] = 1.0 / sqrt(1.0 - exp(-beta));
    }
  }
  return FD_OK;
}

// analysis/metrics.py:120-132 on the device: per-backbone CA-CA bond deviation / valid fraction / steric clashes.
extern "C" int fd_ca_metrics(fd_handle h, int B, int N, const float* ca, const int32_t* n_valid, double tol_bond, double tol_clash, double* out4,
                             void* stream) {
  if (!h || B < 1 || N < 1 || !ca || !out4) return fail(FD_EINVAL, "fd_ca_metrics: bad argument");
  DevGuard dev_guard(h->device);
  ca_metrics_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(ca, n_valid, tol_bond, tol_clash, out4, N);
  CK(cudaGetLastError());
  return FD_OK;
}

// SE3Diffuser.score_scaling (data/se3_diffuser.py:155-158): (rot, trans) scalings at time t; the IGSO(3) row is built on the GPU.
extern "C" int fd_score_scaling(fd_handle h, double t, double* rot_scaling, double* trans_scaling) {
  if (!h) return fail(FD_EINVAL, "null handle");
  if (!(t >= 0.0 && t <= 1.0)) return fail(FD_EINVAL, "Invalid t=%g", t);
  DevGuard dev_guard(h->device);
  const int idx = sigma_idx_host(h->h_sigma_grid, t);
  auto it = h->igso3_rows.find(idx);
  if (it == h->igso3_rows.end()) {
    std::vector<double> cdf(SO3_NOMEGA);
    double scal = 0;
    CKI(build_igso3_rows(h, 1, &idx, nullptr, cdf.data(), nullptr, &scal));
    double* d = nullptr;
    CK(cudaMalloc(&d, SO3_NOMEGA * sizeof(double)));
    CK(cudaMemcpy(d, cdf.data(), SO3_NOMEGA * sizeof(double), cudaMemcpyHostToDevice));
    it = h->igso3_rows.emplace(idx, std::make_pair(d, scal)).first;
  }
  if (rot_scaling) *rot_scaling = it->second.second;
  if (trans_scaling) {
    const double beta = t * R3_MIN_B + 0.5 * (t * t) * (R3_MAX_B - R3_MIN_B);
    *trans_scaling = 1.0 / sqrt(1.0 - exp(-beta));
  }
  return FD_OK;
}

extern "C" int fd_compute_backbone(fd_handle h, int64_t n, const float* rigids, const float* psi, float* atom37, float* atom14, void* stream) {
  if (!h || n <= 0 || !rigids || !psi) return fail(FD_EINVAL, "fd_compute_backbone: bad argument");
  DevGuard dev_guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  compute_backbone_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(rigids, psi, atom37, atom14, n);
  CK(cudaGetLastError());
  return FD_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// sampling loop (Experiment.inference_fn)
// ------------------------------------------------------------------------------------------------------------------
__global__ void set_step_kernel(const StepSched* __restrict__ sched, const double* __restrict__ sched_sigma, const int* __restrict__ step,
                                double* __restrict__ cur_t, double* __restrict__ cur_sigma, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) { const int s = *step; cur_t[i] = (double)(float)sched[s].t; cur_sigma[i] = sched_sigma[s]; }
}
__global__ void inc_step_kernel(int* step) { *step += 1; }
__global__ void fill_f32_kernel(float* p, float v, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void iota_seq_kernel(int* p, int B, int N) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (long long)B * N) p[i] = (int)(i % N) + 1;
}
// trajectory taps, written time-reversed like the reference's flip (index 0 = last step)
__global__ void traj_tap_kernel(const float* __restrict__ rigids, const float* __restrict__ rigids_pred, const float* __restrict__ psi,
                                const float* __restrict__ res_mask, const float* __restrict__ fixed_mask, const int* __restrict__ step,
                                int num_t, float* __restrict__ traj_prot, float* __restrict__ traj_rigid, float* __restrict__ traj_trans0,
                                float* __restrict__ traj_bb0, long long rows) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int s = *step;
  const long long slot = (long long)(num_t - 1 - s);
  float q[4], R[9], t[3];
#pragma unroll
  for (int k = 0; k < 4; ++k) q[k] = rigids[r * 7 + k];
#pragma unroll
  for (int k = 0; k < 3; ++k) t[k] = rigids[r * 7 + 4 + k];
  quat_to_rot(q, R);
  backbone_atoms(R, t, psi[r * 2], psi[r * 2 + 1], traj_prot + (slot * rows + r) * 111, nullptr);
#pragma unroll
  for (int k = 0; k < 7; ++k) traj_rigid[(slot * rows + r) * 7 + k] = rigids[r * 7 + k];   // rigid_traj has num_t+1 slots; slot num_t = init
  float q0[4], t0[3];
#pragma unroll
  for (int k = 0; k < 4; ++k) q0[k] = rigids_pred[r * 7 + k];
#pragma unroll
  for (int k = 0; k < 3; ++k) t0[k] = rigids_pred[r * 7 + 4 + k];
  quat_to_rot(q0, R);
  backbone_atoms(R, t0, psi[r * 2], psi[r * 2 + 1], traj_bb0 + (slot * rows + r) * 111, nullptr);
  const float fm = fixed_mask[r] * res_mask[r], dm = (1.f - fixed_mask[r]) * res_mask[r];
#pragma unroll
  for (int k = 0; k < 3; ++k) traj_trans0[(slot * rows + r) * 3 + k] = dm * t0[k] + fm * t[k];
}

static int ensure_lb(fd_context* h, int B, int N, int num_t, int aux, bool inject) {
  LoopBufs& L = h->lb;
  if (L.base && L.B == B && L.N == N && L.num_t == num_t && L.aux == aux && (!inject || L.z_rot)) return FD_OK;
  cudaStreamSynchronize(h->stream);
  free_lb(h);
  L.B = B; L.N = N; L.num_t = num_t; L.aux = aux;
  const size_t R = (size_t)B * N;
  struct Item { void** p; size_t bytes; };
  std::vector<Item> items = {
      {(void**)&L.rigids, R * 7 * 4}, {(void**)&L.sc_ca, R * 3 * 4}, {(void**)&L.res_mask, R * 4}, {(void**)&L.fixed_mask, R * 4},
      {(void**)&L.psi, R * 2 * 4}, {(void**)&L.rigids_pred, R * 7 * 4}, {(void**)&L.atom37, R * 111 * 4}, {(void**)&L.atom37_0, R * 111 * 4},
      {(void**)&L.rigids_snap, R * 7 * 4}, {(void**)&L.gt_psi, R * 2 * 4}, {(void**)&L.seq_idx, R * 4}, {(void**)&L.rot_score, R * 3 * 8}, {(void**)&L.trans_score, R * 3 * 8},
      {(void**)&L.cur_t, (size_t)B * 8}, {(void**)&L.cur_sigma, (size_t)B * 8},
      {(void**)&L.z_rot, inject ? (size_t)(num_t > 1 ? num_t - 1 : 1) * R * 3 * 8 : 0},
      {(void**)&L.z_trans, inject ? (size_t)(num_t > 1 ? num_t - 1 : 1) * R * 3 * 8 : 0},
      {(void**)&L.z_axis, inject ? R * 3 * 8 : 0}, {(void**)&L.u_angle, inject ? R * 8 : 0}, {(void**)&L.z_trans0, inject ? R * 3 * 8 : 0},
      {(void**)&L.sched, (size_t)num_t * sizeof(StepSched)}, {(void**)&L.sched_sigma, (size_t)num_t * 8}, {(void**)&L.step, 256},
      {(void**)&L.traj_prot, aux ? (size_t)num_t * R * 111 * 4 : 0}, {(void**)&L.traj_rigid, aux ? (size_t)(num_t + 1) * R * 7 * 4 : 0},
      {(void**)&L.traj_trans0, aux ? (size_t)num_t * R * 3 * 4 : 0}, {(void**)&L.traj_bb0, aux ? (size_t)num_t * R * 111 * 4 : 0}};
  size_t total = 0;
  for (auto& it : items) total += al256(it.bytes);
  if (cudaMalloc(&L.base, total) != cudaSuccess) {
    cudaGetLastError();
    L = LoopBufs();
    return fail(FD_ENOMEM, "loop buffers (%.1f MB) allocation failed", total / 1048576.0);
  }
  L.bytes = total;
  char* p = L.base;
  for (auto& it : items) { *it.p = it.bytes ? (void*)p : nullptr; p += al256(it.bytes); }
  return FD_OK;
}

static int run_loop(fd_context* h, const fd_sample_cfg* cfg, bool inject_steps, double* gpu_ms, int64_t* launches) {
  LoopBufs& L = h->lb;
  cudaStream_t st = h->stream;
  const int B = cfg->B, N = cfg->N, num_t = cfg->num_t;
  const long long R = (long long)B * N;
  // schedule: reverse_steps = linspace(min_t, 1, num_t)[::-1], dt = 1/num_t   (train_se3_diffusion.py:746-747)
  std::vector<StepSched> sched(num_t);
  std::vector<double> ssig(num_t);
  const double stepsz = num_t > 1 ? (1.0 - cfg->min_t) / (double)(num_t - 1) : 0.0;
  for (int s = 0; s < num_t; ++s) {
    const int k = num_t - 1 - s;
    const double t = (k == num_t - 1 && num_t > 1) ? 1.0 : cfg->min_t + (double)k * stepsz;
    sched[s] = StepSched{t, so3_g_host(t), r3_b_host(t), 1.0 / (double)num_t};
    // the network sees t rounded to fp32 (t * ones(fp32)); sigma is quantised from that value (du.move_to_np(t))
    ssig[s] = h->h_sigma_grid[sigma_idx_host(h->h_sigma_grid, (double)(float)t)];
  }
  if (h->fwd_pending) { CK(cudaStreamWaitEvent(st, h->ev_fwd, 0)); h->fwd_pending = false; }
  CK(cudaMemcpyAsync(L.sched, sched.data(), num_t * sizeof(StepSched), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(L.sched_sigma, ssig.data(), num_t * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(L.step, 0, sizeof(int), st));
  CK(cudaMemsetAsync(L.sc_ca, 0, R * 3 * sizeof(float), st));
  CKI(ensure_ws(h, B, N));
  CK(cudaStreamSynchronize(st));   // sched vectors go out of scope only after the copies are done

  fd_forward_out fo{L.rot_score, L.trans_score, L.psi, L.rigids_pred, L.atom37_0, nullptr};
  auto forward = [&](bool write_sc) -> int {
    return forward_impl(h, B, N, L.rigids, L.cur_t, 1, L.cur_sigma, L.res_mask, L.fixed_mask, L.seq_idx, L.sc_ca,
                        L.have_gt_psi ? L.gt_psi : nullptr, &fo, write_sc ? L.sc_ca : nullptr, st);
  };
  auto step_body = [&]() -> int {
    set_step_kernel<<<(B + 127) / 128, 128, 0, st>>>(L.sched, L.sched_sigma, L.step, L.cur_t, L.cur_sigma, B);
    h->launches++;
    CKI(forward(true));
    ReverseArgs a{};
    a.rigids = L.rigids; a.rot_score = L.rot_score; a.trans_score = L.trans_score; a.diffuse_mask = nullptr; a.use_masks = 1;
    a.res_mask = L.res_mask; a.fixed_mask = L.fixed_mask;
    a.z_rot = inject_steps ? L.z_rot : nullptr; a.z_trans = inject_steps ? L.z_trans : nullptr;
    a.sched = L.sched; a.step_ptr = L.step; a.seed = cfg->seed; a.first_sample = cfg->first_sample; a.center = cfg->center;
    a.noise_scale = cfg->noise_scale; a.rotmat_out = nullptr; a.N = N;
    a.noise_stride = R * 3;   // per-step noise slice selected on the device from the step counter
    a.rng_step_bias = 0;
    CKI(launch_reverse(h, a, B, st));
    if (cfg->aux_traj) {
      traj_tap_kernel<<<(unsigned)((R + 127) / 128), 128, 0, st>>>(L.rigids, L.rigids_pred, L.psi, L.res_mask, L.fixed_mask, L.step, num_t,
                                                                   L.traj_prot, L.traj_rigid, L.traj_trans0, L.traj_bb0, R);
      h->launches++;
    }
    inc_step_kernel<<<1, 1, 0, st>>>(L.step);
    h->launches++;
    CK(cudaGetLastError());
    return FD_OK;
  };

  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const bool saved_timing = h->stage_timing;
  const bool saved_debug = h->debug;
  h->launches = 0;
  CK(cudaEventRecord(e0, st));
  if (cfg->aux_traj) {   // rigid_traj[num_t] (last after flip) = initial frames
    CK(cudaMemcpyAsync(L.traj_rigid + (size_t)num_t * R * 7, L.rigids, R * 7 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  // priming forward for self-conditioning at t = reverse_steps[0]   (train_se3_diffusion.py:753-756)
  if (cfg->self_condition) {
    set_step_kernel<<<(B + 127) / 128, 128, 0, st>>>(L.sched, L.sched_sigma, L.step, L.cur_t, L.cur_sigma, B);
    h->launches++;
    CKI(forward(true));
  }
  const int nrev = num_t - 1;   // steps with t > min_t
  long long per_step = 0;
  if (cfg->use_graph && nrev > 0 && !saved_timing && !saved_debug) {
    fd_context::GraphCache& G = h->gc;
    const bool hit = G.exec && G.B == B && G.N == N && G.num_t == num_t && G.aux == cfg->aux_traj && G.inject == (int)inject_steps &&
                     G.center == cfg->center && G.precision == h->precision && G.have_psi == (int)L.have_gt_psi &&
                     G.noise_scale == cfg->noise_scale && G.seed == cfg->seed && G.first_sample == cfg->first_sample &&
                     G.ws_base == h->ws.base && G.lb_base == L.base && G.warena == h->warena;
    if (!hit) {
      free_graph(h);
      h->stage_timing = false; h->debug = false;
      const long long l0 = h->launches;
      CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      int rc = step_body();
      cudaError_t ce = cudaStreamEndCapture(st, &G.graph);
      h->stage_timing = saved_timing; h->debug = saved_debug;
      if (rc != FD_OK) { free_graph(h); return rc; }
      if (ce != cudaSuccess) { free_graph(h); return fail(FD_ECUDA, "graph capture failed: %s", cudaGetErrorString(ce)); }
      G.per_step = h->launches - l0;
      h->launches = l0;
      CK(cudaGraphInstantiate(&G.exec, G.graph, 0));
      G.B = B; G.N = N; G.num_t = num_t; G.aux = cfg->aux_traj; G.inject = inject_steps; G.center = cfg->center; G.precision = h->precision;
      G.have_psi = L.have_gt_psi; G.noise_scale = cfg->noise_scale; G.seed = cfg->seed; G.first_sample = cfg->first_sample;
      G.ws_base = h->ws.base; G.lb_base = L.base; G.warena = h->warena;
    }
    per_step = G.per_step;
    for (int s = 0; s < nrev; ++s) CK(cudaGraphLaunch(G.exec, st));
    h->launches += per_step * nrev;
    CK(cudaStreamSynchronize(st));
  } else {
    for (int s = 0; s < nrev; ++s) CKI(step_body());
  }
  // final step (t == min_t): forward only, reusing the previous step's t features; sample := predicted frames
  // (train_se3_diffusion.py:778-781)
  if (num_t == 1 && !cfg->self_condition) {
    set_step_kernel<<<(B + 127) / 128, 128, 0, st>>>(L.sched, L.sched_sigma, L.step, L.cur_t, L.cur_sigma, B);
    h->launches++;
  }
  CK(cudaMemcpyAsync(L.rigids_snap, L.rigids_pred, R * 7 * sizeof(float), cudaMemcpyDeviceToDevice, st));   // stale rigid_pred for trans_traj
  {
    // rigids_t is both input and output here: the head kernel reads rigids_t rows it then overwrites (same warp, after
    // all its reads), so run the forward into rigids_pred and copy.
    fd_forward_out ftmp{L.rot_score, L.trans_score, L.psi, L.rigids_pred, L.atom37, nullptr};
    CKI(forward_impl(h, B, N, L.rigids, L.cur_t, 1, L.cur_sigma, L.res_mask, L.fixed_mask, L.seq_idx, L.sc_ca,
                     L.have_gt_psi ? L.gt_psi : nullptr, &ftmp, nullptr, st));
    if (cfg->aux_traj) {
      // quirk C.4: rigid_0_traj / trans_traj of the last step reuse the PREVIOUS prediction (rigid_pred not refreshed)
      CK(cudaMemcpyAsync(L.rigids, L.rigids_pred, R * 7 * sizeof(float), cudaMemcpyDeviceToDevice, st));
      traj_tap_kernel<<<(unsigned)((R + 127) / 128), 128, 0, st>>>(L.rigids, num_t > 1 ? L.rigids_snap : L.rigids_pred, L.psi, L.res_mask,
                                                                   L.fixed_mask, L.step, num_t, L.traj_prot, L.traj_rigid, L.traj_trans0,
                                                                   L.traj_bb0, R);
      h->launches++;
    } else {
      CK(cudaMemcpyAsync(L.rigids, L.rigids_pred, R * 7 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
  }
  CK(cudaEventRecord(e1, st));
  CK(cudaStreamSynchronize(st));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (gpu_ms) *gpu_ms = ms;
  if (launches) *launches = h->launches;
  return FD_OK;
}

static int validate_cfg(const fd_sample_cfg* c) {
  if (!c) return fail(FD_EINVAL, "null cfg");
  if (c->B <= 0 || c->N <= 0 || c->num_t <= 0) return fail(FD_EINVAL, "fd_sample: B=%d N=%d num_t=%d", c->B, c->N, c->num_t);
  if (!(c->min_t >= 0.0 && c->min_t <= 1.0)) return fail(FD_EINVAL, "Invalid t=%g", c->min_t);
  return FD_OK;
}

static int init_loop_inputs(fd_context* h, const fd_sample_cfg* cfg, const float* res_mask_h, const float* fixed_mask_h, const int32_t* seq_h,
                            const float* gt_psi_h) {
  LoopBufs& L = h->lb;
  cudaStream_t st = h->stream;
  const long long R = (long long)cfg->B * cfg->N;
  L.have_gt_psi = gt_psi_h != nullptr;
  if (gt_psi_h) CK(cudaMemcpyAsync(L.gt_psi, gt_psi_h, R * 2 * 4, cudaMemcpyHostToDevice, st));
  if (res_mask_h) CK(cudaMemcpyAsync(L.res_mask, res_mask_h, R * 4, cudaMemcpyHostToDevice, st));
  else fill_f32_kernel<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(L.res_mask, 1.f, R);
  if (fixed_mask_h) CK(cudaMemcpyAsync(L.fixed_mask, fixed_mask_h, R * 4, cudaMemcpyHostToDevice, st));
  else CK(cudaMemsetAsync(L.fixed_mask, 0, R * 4, st));
  if (seq_h) CK(cudaMemcpyAsync(L.seq_idx, seq_h, R * 4, cudaMemcpyHostToDevice, st));
  else iota_seq_kernel<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(L.seq_idx, cfg->B, cfg->N);
  CK(cudaGetLastError());
  return FD_OK;
}

extern "C" int fd_sample_host(fd_handle h, const fd_sample_cfg* cfg, const fd_sample_in* in, const fd_sample_out* out) {
  if (!h || !out) return fail(FD_EINVAL, "fd_sample_host: null argument");
  CKI(validate_cfg(cfg));
  if (!h->weights_loaded) return fail(FD_ESTATE, "fd_sample: weights not loaded");
  DevGuard dev_guard(h->device);
  static const fd_sample_in kEmpty = {};
  if (!in) in = &kEmpty;
  const bool inj_prior = in->z_axis != nullptr, inj_steps = in->z_rot != nullptr;
  if (inj_prior && !(in->u_angle && in->z_trans0)) return fail(FD_EINVAL, "fd_sample: inject z_axis, u_angle and z_trans0 together");
  if (inj_steps && !in->z_trans) return fail(FD_EINVAL, "fd_sample: inject z_rot and z_trans together");
  const int B = cfg->B, N = cfg->N, num_t = cfg->num_t;
  const long long R = (long long)B * N;
  CKI(ensure_lb(h, B, N, num_t, cfg->aux_traj, inj_prior || inj_steps));
  LoopBufs& L = h->lb;
  cudaStream_t st = h->stream;
  CKI(init_loop_inputs(h, cfg, in->res_mask, in->fixed_mask, in->seq_idx, in->gt_psi));
  if (in->rigids_init) {
    CK(cudaMemcpyAsync(L.rigids, in->rigids_init, R * 7 * 4, cudaMemcpyHostToDevice, st));
  } else if (inj_prior) {
    CK(cudaMemcpyAsync(L.z_axis, in->z_axis, R * 3 * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(L.u_angle, in->u_angle, R * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(L.z_trans0, in->z_trans0, R * 3 * 8, cudaMemcpyHostToDevice, st));
    CKI(fd_sample_ref(h, R, L.z_axis, L.u_angle, L.z_trans0, 0, 0, N, L.rigids, st));
  } else {
    CKI(fd_sample_ref(h, R, nullptr, nullptr, nullptr, cfg->seed, cfg->first_sample, N, L.rigids, st));
  }
  if (inj_steps && num_t > 1) {
    CK(cudaMemcpyAsync(L.z_rot, in->z_rot, (size_t)(num_t - 1) * R * 3 * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(L.z_trans, in->z_trans, (size_t)(num_t - 1) * R * 3 * 8, cudaMemcpyHostToDevice, st));
  }
  int64_t launches = 0; double ms = 0;
  CKI(run_loop(h, cfg, inj_steps, &ms, &launches));
  if (out->atom37_final) CK(cudaMemcpyAsync(out->atom37_final, L.atom37, R * 111 * 4, cudaMemcpyDeviceToHost, st));
  if (out->rigids_final) CK(cudaMemcpyAsync(out->rigids_final, L.rigids, R * 7 * 4, cudaMemcpyDeviceToHost, st));
  if (out->psi_final) CK(cudaMemcpyAsync(out->psi_final, L.psi, R * 2 * 4, cudaMemcpyDeviceToHost, st));
  if (cfg->aux_traj) {
    if (out->prot_traj) CK(cudaMemcpyAsync(out->prot_traj, L.traj_prot, (size_t)num_t * R * 111 * 4, cudaMemcpyDeviceToHost, st));
    if (out->rigid_traj) CK(cudaMemcpyAsync(out->rigid_traj, L.traj_rigid, (size_t)(num_t + 1) * R * 7 * 4, cudaMemcpyDeviceToHost, st));
    if (out->trans_traj) CK(cudaMemcpyAsync(out->trans_traj, L.traj_trans0, (size_t)num_t * R * 3 * 4, cudaMemcpyDeviceToHost, st));
    if (out->rigid_0_traj) CK(cudaMemcpyAsync(out->rigid_0_traj, L.traj_bb0, (size_t)num_t * R * 111 * 4, cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(st));
  if (out->gpu_ms) *out->gpu_ms = ms;
  if (out->kernel_launches) *out->kernel_launches = launches;
  return FD_OK;
}

extern "C" int fd_sample_dev(fd_handle h, const fd_sample_cfg* cfg, const float* rigids_init_dev, float* atom37_dev, float* rigids_dev,
                             double* gpu_ms, int64_t* kernel_launches) {
  if (!h) return fail(FD_EINVAL, "fd_sample_dev: null handle");
  CKI(validate_cfg(cfg));
  if (!h->weights_loaded) return fail(FD_ESTATE, "fd_sample: weights not loaded");
  if (cfg->aux_traj) return fail(FD_EINVAL, "fd_sample_dev: aux_traj is only available through fd_sample_host");
  DevGuard dev_guard(h->device);
  const long long R = (long long)cfg->B * cfg->N;
  CKI(ensure_lb(h, cfg->B, cfg->N, cfg->num_t, 0, false));
  LoopBufs& L = h->lb;
  cudaStream_t st = h->stream;
  CKI(init_loop_inputs(h, cfg, nullptr, nullptr, nullptr, nullptr));
  if (rigids_init_dev) CK(cudaMemcpyAsync(L.rigids, rigids_init_dev, R * 7 * 4, cudaMemcpyDeviceToDevice, st));
  else CKI(fd_sample_ref(h, R, nullptr, nullptr, nullptr, cfg->seed, cfg->first_sample, cfg->N, L.rigids, st));
  CKI(run_loop(h, cfg, false, gpu_ms, kernel_launches));
  if (atom37_dev) CK(cudaMemcpyAsync(atom37_dev, L.atom37, R * 111 * 4, cudaMemcpyDeviceToDevice, st));
  if (rigids_dev) CK(cudaMemcpyAsync(rigids_dev, L.rigids, R * 7 * 4, cudaMemcpyDeviceToDevice, st));
  CK(cudaStreamSynchronize(st));
  return FD_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// training step (include/framediff_b200.h: fd_train_*)
// ------------------------------------------------------------------------------------------------------------------
#include "fd_train_host.cuh"

static fd_train_state* train_state(fd_context* h) {
  if (!h->train) h->train = new fd_train_state();
  return h->train;
}
static void free_train(fd_context* h) {
  if (!h->train) return;
  free_tape(h->train->tape);
  ttc_free(h->train->tc);
  delete h->train;
  h->train = nullptr;
}
extern "C" int64_t fd_train_arena_floats(void) { return arena_layout().total; }
extern "C" int64_t fd_train_param_offset(int i) {
  const auto& L = arena_layout();
  return (i < 0 || i >= (int)L.off.size()) ? FD_EINVAL : L.off[i];
}
extern "C" int fd_train_bind(fd_handle h, float* params, float* grads) {
  if (!h || !params || !grads) return fail(FD_EINVAL, "fd_train_bind: null argument");
  fd_train_state* S = train_state(h);
  S->P = params; S->G = grads;
  S->tape.valid = false;
  return FD_OK;
}
extern "C" int fd_train_forward(fd_handle h, int B, int N, const fd_forward_in* in, const fd_forward_out* out, void* stream) {
  if (!h || !in || !out) return fail(FD_EINVAL, "fd_train_forward: null argument");
  if (!h->train || !h->train->P) return fail(FD_ESTATE, "fd_train_forward: no parameter arena bound (fd_train_bind)");
  if (B <= 0 || N <= 0) return fail(FD_EINVAL, "fd_train_forward: B=%d N=%d", B, N);
  if (!in->rigids_t || !in->t || !in->res_mask || !in->fixed_mask || !in->seq_idx || !in->sc_ca_t)
    return fail(FD_EINVAL, "fd_train_forward: a required input pointer is NULL");
  if (!out->rot_score || !out->trans_score || !out->psi || !out->rigids) return fail(FD_EINVAL, "fd_train_forward: a required output pointer is NULL");
  if (in->cached_score_rows) return fail(FD_EINVAL, "fd_train_forward: use_cached_score is an inference-only lookup (piecewise constant in the angle)");
  DevGuard dev_guard(h->device);
  return train_forward_impl(h, h->train, B, N, in, out, (cudaStream_t)stream);
}
extern "C" int fd_train_backward(fd_handle h, const fd_train_grads* dout, int stage_first, int stage_last, void* stream) {
  if (!h || !dout) return fail(FD_EINVAL, "fd_train_backward: null argument");
  if (!h->train || !h->train->G) return fail(FD_ESTATE, "fd_train_backward: no gradient arena bound (fd_train_bind)");
  if (stage_first < 0 || stage_last >= NBLK || stage_first > stage_last) return fail(FD_EINVAL, "fd_train_backward: stages %d..%d", stage_first, stage_last);
  DevGuard dev_guard(h->device);
  return train_backward_impl(h, h->train, dout, stage_first, stage_last, (cudaStream_t)stream);
}
// Device memory the handle allocates for a (B, N) problem: which = 0 inference workspace (current precision mode), 1 sampling-loop buffers
// (num_t steps, aux_traj trajectories if aux != 0), 2 training tape.  SURVEY §8(b) lists a workspace query; the handle owns these arenas.
extern "C" int64_t fd_workspace_bytes(fd_handle h, int which, int B, int N, int num_t, int aux) {
  if (!h || B < 1 || N < 1) return FD_EINVAL;
  const size_t R = (size_t)B * N, E = R * N, Np = (size_t)((N + 3) & ~3);
  if (which == 0) {
    const bool tc = h->precision != FD_PREC_FP32;
    const size_t chunk = E < ((size_t)1 << 18) ? E : ((size_t)1 << 18);
    size_t fl = R * (NODE_IN_PAD + 256 + 5 * C_S + 4 * TF_D + 3 * TF_D + PROJ_ALL + 2 * H * PQ * 3 + 2 * H * PV * 3 + IPA_FEAT + 7 + C_Z + ET_NODE + C_S) +
                (size_t)B * 32 + (tc ? R * H * C_Z : 0) + (size_t)B * (TF_H + H) * N * Np + E * C_Z + (tc ? 0 : chunk * (2 * ET_HID + C_Z));
    return (int64_t)(fl * sizeof(float) + (tc ? tc_workspace_bytes(B, N) : 0));
  }
  if (which == 1) {
    const size_t T = (size_t)(num_t > 0 ? num_t : 1);
    size_t by = R * (7 * 4 * 3 + 3 * 4 + 2 * 4 + 2 * 4 + 2 * 4 + 111 * 4 * 2 + 4 + 2 * 3 * 8) + T * (sizeof(StepSched) + 8);
    if (aux) by += T * R * (111 * 4 * 2 + 3 * 4) + (T + 1) * R * 7 * 4;
    return (int64_t)by;
  }
  if (which == 2) {
    const size_t AT = (size_t)B * H * N * Np, PT = (size_t)B * TF_H * N * Np;
    size_t fl = R * (NODE_IN_PAD + 4 * 256 + 7 + 2 * C_S) + E * (EDGE_IN + 3 * C_Z) +                                                   // embedders, heads
                R * (5 * C_S + 3 * TF_D + 3 * TF_D + IPA_FEAT + PROJ_ALL + 7 + 2 * H * C_Z + 4 * H * PQ * 3 + 2 * H * PV * 3 + 2 * ET_HID + 3 * C_Z) + PT + AT +
                E * (4 * C_Z + 2 * ET_HID + H) +                                                                                       // backward scratch
                NBLK * (E * C_Z + R * (PROJ_ALL + 7 + 2 * H * PQ * 3 + 2 * H * PV * 3 + H * C_Z + IPA_FEAT + 6 * C_S + (TF_LAYERS + 1) * TF_D +
                                       TF_LAYERS * (3 * TF_D + 6 * TF_D)) + AT + TF_LAYERS * PT) +
                (NBLK - 1) * (R * (C_Z + ET_NODE) + E * (2 * ET_HID + C_Z));
    return (int64_t)(fl * sizeof(float));
  }
  return FD_EINVAL;
}
// actual size of the arena currently allocated (0 if none): lets tests check fd_workspace_bytes against what the handle really allocated
extern "C" int64_t fd_debug_alloc_bytes(fd_handle h, int which) {
  if (!h) return FD_EINVAL;
  if (which == 0) return (int64_t)h->ws.bytes;
  if (which == 1) return (int64_t)h->lb.bytes;
  if (which == 2) return h->train ? (int64_t)h->train->tape.bytes : 0;
  return FD_EINVAL;
}
extern "C" int fd_train_set_gemm(fd_handle h, int mode) {
  if (!h || mode < 0 || mode > 2) return fail(FD_EINVAL, "fd_train_set_gemm: mode %d", mode);
  h->train_gemm = mode;
  return FD_OK;
}
extern "C" int fd_train_release(fd_handle h) {
  if (!h) return FD_EINVAL;
  DevGuard dev_guard(h->device);
  cudaDeviceSynchronize();
  free_train(h);
  return FD_OK;
}
extern "C" int fd_loss_backward(fd_handle h, int B, int N, const fd_loss_in* in, const fd_loss_cfg* cfg, const fd_train_grads_out* out, void* stream) {
  if (!h || !in || !cfg || !out || B < 1 || N < 1) return fail(FD_EINVAL, "fd_loss_backward: bad arguments");
  if (!out->d_rot_score || !out->d_trans_score || !out->d_rigids || !out->d_atom37) return fail(FD_EINVAL, "fd_loss_backward: null output pointer");
  if (!in->pred_rot_score || !in->pred_trans_score || !in->pred_rigids || !in->pred_atom37 || !in->gt_rot_score || !in->gt_trans_score ||
      !in->rot_score_scaling || !in->trans_score_scaling || !in->rigids_0 || !in->t || !in->res_mask || !in->fixed_mask || !in->gt_psi)
    return fail(FD_EINVAL, "fd_loss_backward: null input pointer");
  const size_t smem = (size_t)240 * N;
  if (smem > 200 * 1024) return fail(FD_EINVAL, "fd_loss_backward: N = %d too long", N);
  DevGuard dev_guard(h->device);
  // number of samples with a non-empty mask (train_se3_diffusion.py:662) — a host value: the masks are inputs the caller just uploaded
  std::vector<float> hm((size_t)B * N);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CK(cudaMemcpyAsync(hm.data(), in->res_mask, hm.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  int nvalid = 0;
  for (int b = 0; b < B; ++b) { bool any = false; for (int n = 0; n < N; ++n) any |= hm[(size_t)b * N + n] != 0.f; nvalid += any; }
  LossBwdArgs a{};
  a.pred_rot = in->pred_rot_score; a.pred_trans = in->pred_trans_score; a.pred_rigids = in->pred_rigids; a.pred_atom37 = in->pred_atom37;
  a.gt_rot = in->gt_rot_score; a.gt_trans = in->gt_trans_score; a.rot_scaling = in->rot_score_scaling; a.trans_scaling = in->trans_score_scaling;
  a.rigids_0 = in->rigids_0; a.t = in->t; a.res_mask = in->res_mask; a.fixed_mask = in->fixed_mask; a.gt_psi = in->gt_psi;
  a.trans_loss_weight = cfg->trans_loss_weight; a.rot_loss_weight = cfg->rot_loss_weight; a.rot_loss_t_threshold = cfg->rot_loss_t_threshold;
  a.trans_x0_threshold = cfg->trans_x0_threshold; a.coordinate_scaling = cfg->coordinate_scaling; a.bb_atom_loss_weight = cfg->bb_atom_loss_weight;
  a.bb_atom_loss_t_filter = cfg->bb_atom_loss_t_filter; a.dist_mat_loss_weight = cfg->dist_mat_loss_weight;
  a.dist_mat_loss_t_filter = cfg->dist_mat_loss_t_filter; a.aux_loss_weight = cfg->aux_loss_weight;
  a.separate_rot_loss = cfg->separate_rot_loss; a.diffuse_trans = cfg->diffuse_trans; a.diffuse_rot = cfg->diffuse_rot;
  a.inv_nvalid = 1.0 / ((double)nvalid + 1e-10);
  a.d_rot = out->d_rot_score; a.d_trans = out->d_trans_score; a.d_rigids = out->d_rigids; a.d_atom37 = out->d_atom37; a.N = N;
  {
    const size_t smem2 = (size_t)32 * N * sizeof(float);
    if (smem2 > 48 * 1024) {
      CK(cudaFuncSetAttribute(loss_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      CK(cudaFuncSetAttribute(loss_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    }
    if (!h->d_loss_acc) CK(cudaMalloc(&h->d_loss_acc, 4096 * 10 * sizeof(double)));
    if (B > 4096) return fail(FD_EINVAL, "fd_loss_backward: B = %d too large", B);
    CK(cudaMemsetAsync(h->d_loss_acc, 0, (size_t)B * 3 * sizeof(double), st));
    const dim3 grid(B, (N + LOSS_RES - 1) / LOSS_RES);
    loss_count_kernel<<<grid, 256, smem2, st>>>(a, h->d_loss_acc);
    CK(cudaGetLastError());
    loss_bwd2_kernel<<<grid, 256, smem2, st>>>(a, h->d_loss_acc);
    CK(cudaGetLastError());
    h->launches += 2;
  }
  return FD_OK;
}
extern "C" int fd_adam_step(fd_handle h, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                            double beta2, double eps, int64_t step, double grad_scale, void* stream) {
  if (!h || !params || !grads || !exp_avg || !exp_avg_sq || n < 1 || step < 1) return fail(FD_EINVAL, "fd_adam_step: bad arguments");
  DevGuard dev_guard(h->device);
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(params, grads, exp_avg, exp_avg_sq, n, (float)lr, (float)beta1,
                                                                                          (float)beta2, (float)eps, (float)bc1, (float)bc2, (float)grad_scale);
  CK(cudaGetLastError());
  h->launches++;
  return FD_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// introspection
// ------------------------------------------------------------------------------------------------------------------
// Developer aid (not in the public header): cycle counters of the fused EdgeTransition kernel's roles (CTA 0, last launch).
extern "C" int fd_debug_tc_profile(fd_handle h, int on, long long* out32) {
  if (!h) return FD_EINVAL;
  DevGuard dev_guard(h->device);
  cudaDeviceSynchronize();
  if (on && !g_tc_prof) { cudaMalloc(&g_tc_prof, 32 * sizeof(long long)); cudaMemset(g_tc_prof, 0, 32 * sizeof(long long)); }
  if (out32 && g_tc_prof) cudaMemcpy(out32, g_tc_prof, 32 * sizeof(long long), cudaMemcpyDeviceToHost);
  if (!on && g_tc_prof) { cudaFree(g_tc_prof); g_tc_prof = nullptr; }
  return FD_OK;
}

extern "C" int64_t fd_launch_count(fd_handle h) { return h ? (int64_t)h->launches : -1; }
extern "C" int fd_num_stages(void) { return ST_COUNT; }
extern "C" const char* fd_stage_name(int i) { return (i < 0 || i >= ST_COUNT) ? nullptr : kStageNames[i]; }
extern "C" int fd_set_stage_timing(fd_handle h, int on) {
  if (!h) return FD_EINVAL;
  h->stage_timing = on != 0;
  memset(h->stage_ms, 0, sizeof(h->stage_ms));
  memset(h->stage_launches, 0, sizeof(h->stage_launches));
  return FD_OK;
}
extern "C" int fd_stage_times(fd_handle h, double* ms_out, int64_t* launches_out) {
  if (!h) return FD_EINVAL;
  for (int i = 0; i < ST_COUNT; ++i) {
    if (ms_out) ms_out[i] = h->stage_ms[i];
    if (launches_out) launches_out[i] = h->stage_launches[i];
  }
  return FD_OK;
}
extern "C" int64_t fd_forward_flops(int B, int N, int executed) {
  // reference-equivalent F(N) = 2,248,960·N² + 32,421,376·N per sample.
  // Executed: EdgeTransition uses the separable first/last layers (524,288 instead of 688,128 FLOP/edge ×3) and the
  // edge embedder's layer 0 is a table lookup (65,536 instead of 96,256 FLOP/edge); IPA pair terms use the
  // Σ_j a·z reordering (2,048+2,048 instead of 2,048+8,192+512).
  const double n = N, n2 = n * n;
  const double ref = 2248960.0 * n2 + 32421376.0 * n;
  const double exe = ref - 3 * (688128.0 - 524288.0) * n2 - (96256.0 - 65536.0) * n2 - 4 * (8192.0 + 512.0 - 2048.0) * n2;
  return (int64_t)((executed ? exe : ref) * B);
}

// ------------------------------------------------------------------------------------------------------------------
// DSM loss, forward values (include/framediff_b200.h: fd_loss_forward)
// ------------------------------------------------------------------------------------------------------------------
extern "C" int fd_loss_forward(fd_handle h, int B, int N, const fd_loss_in* in, const fd_loss_cfg* cfg, double* terms_dev, void* stream) {
  if (!h || !in || !cfg || !terms_dev || B < 1 || N < 1) return fail(FD_EINVAL, "fd_loss_forward: bad arguments");
  if (!in->pred_rot_score || !in->pred_trans_score || !in->pred_rigids || !in->pred_atom37 || !in->gt_rot_score || !in->gt_trans_score ||
      !in->rot_score_scaling || !in->trans_score_scaling || !in->rigids_0 || !in->t || !in->res_mask || !in->fixed_mask || !in->gt_psi)
    return fail(FD_EINVAL, "fd_loss_forward: null input pointer");
  const size_t smem = (size_t)30 * N * sizeof(float);
  if (smem > 200 * 1024) return fail(FD_EINVAL, "fd_loss_forward: N = %d too long (pair-distance tile needs %zu bytes of shared memory)", N, smem);
  DevGuard dev_guard(h->device);
  LossArgs a{};
  a.pred_rot = in->pred_rot_score; a.pred_trans = in->pred_trans_score; a.pred_rigids = in->pred_rigids; a.pred_atom37 = in->pred_atom37;
  a.gt_rot = in->gt_rot_score; a.gt_trans = in->gt_trans_score; a.rot_scaling = in->rot_score_scaling; a.trans_scaling = in->trans_score_scaling;
  a.rigids_0 = in->rigids_0; a.t = in->t; a.res_mask = in->res_mask; a.fixed_mask = in->fixed_mask; a.gt_psi = in->gt_psi;
  a.trans_loss_weight = cfg->trans_loss_weight; a.rot_loss_weight = cfg->rot_loss_weight; a.rot_loss_t_threshold = cfg->rot_loss_t_threshold;
  a.trans_x0_threshold = cfg->trans_x0_threshold; a.coordinate_scaling = cfg->coordinate_scaling; a.bb_atom_loss_weight = cfg->bb_atom_loss_weight;
  a.bb_atom_loss_t_filter = cfg->bb_atom_loss_t_filter; a.dist_mat_loss_weight = cfg->dist_mat_loss_weight;
  a.dist_mat_loss_t_filter = cfg->dist_mat_loss_t_filter; a.aux_loss_weight = cfg->aux_loss_weight;
  a.separate_rot_loss = cfg->separate_rot_loss; a.diffuse_trans = cfg->diffuse_trans; a.diffuse_rot = cfg->diffuse_rot;
  a.terms = terms_dev; a.N = N;
  {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t smem2 = (size_t)32 * N * sizeof(float);
    if (smem2 > 48 * 1024) CK(cudaFuncSetAttribute(loss_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    if (!h->d_loss_acc) CK(cudaMalloc(&h->d_loss_acc, 4096 * 10 * sizeof(double)));
    if (B > 4096) return fail(FD_EINVAL, "fd_loss_forward: B = %d too large", B);
    CK(cudaMemsetAsync(h->d_loss_acc, 0, (size_t)B * 10 * sizeof(double), st));
    loss_fwd2_kernel<<<dim3(B, (N + LOSS_RES - 1) / LOSS_RES), 256, smem2, st>>>(a, h->d_loss_acc);
    CK(cudaGetLastError());
    loss_finalize_kernel<<<(B + 127) / 128, 128, 0, st>>>(a, h->d_loss_acc, B);
    CK(cudaGetLastError());
    h->launches += 2;
  }
  return FD_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// PDB text of sampled backbones (host only).  Restates analysis/utils.py:39-77 (write_prot_to_pdb, create_full_prot) and
// data/protein.py:146-219 (to_pdb) for the single-chain proteins the sampler writes; byte-exact (tests/test_pdb_writer.py).
// ------------------------------------------------------------------------------------------------------------------
namespace {
const char* kAtomTypes[37] = {"N", "CA", "C", "CB", "O", "CG", "CG1", "CG2", "OG", "OG1", "SG", "CD", "CD1", "CD2", "ND1", "ND2", "OD1", "OD2", "SD",
                              "CE", "CE1", "CE2", "CE3", "NE", "NE1", "NE2", "OE1", "OE2", "CH2", "NH1", "NH2", "OH", "CZ", "CZ2", "CZ3", "NZ", "OXT"};
const char* kRes3[21] = {"ALA", "ARG", "ASN", "ASP", "CYS", "GLN", "GLU", "GLY", "HIS", "ILE", "LEU", "LYS", "MET", "PHE", "PRO", "SER", "THR",
                         "TRP", "TYR", "VAL", "UNK"};
struct PdbSink {
  char* out; size_t cap; size_t len;
  void line(const char* s, int n) {            // pad to 80 columns (never truncate: over-wide fields widen the line), add '\n'
    const int pad = n < 80 ? 80 - n : 0;
    if (len + (size_t)n + pad + 1 <= cap) {
      memcpy(out + len, s, n);
      memset(out + len + n, ' ', pad);
      out[len + n + pad] = '\n';
    }
    len += (size_t)n + pad + 1;
  }
};
// printf("%{width}.{dec}f") for dec in {2,3}, right-aligned, never truncated.  Fast path: scale, round to nearest-even, emit digits.
// The scaled product is exact for float32-born values (24 + 10 bits), so ties resolve exactly like printf's correctly rounded
// conversion; anything near a tie, huge or non-finite goes through snprintf itself.
inline int fmt_fixed(char* dst, double x, int width, int dec) {
  const double scale = dec == 3 ? 1000.0 : 100.0;
  const double y = fabs(x) * scale;
  if (!(y < 9.0e15)) return snprintf(dst, 64, dec == 3 ? "%*.3f" : "%*.2f", width, x);   // also NaN / inf
  const double f = nearbyint(y);                 // round-half-even in the default rounding mode
  const double d = fabs(y - f);
  if (d > 0.499999 && d < 0.500001) return snprintf(dst, 64, dec == 3 ? "%*.3f" : "%*.2f", width, x);
  unsigned long long v = (unsigned long long)f;
  char tmp[32];
  int n = 0;
  for (int k = 0; k < dec; ++k) { tmp[n++] = (char)('0' + v % 10); v /= 10; }
  tmp[n++] = '.';
  do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  if (std::signbit(x)) tmp[n++] = '-';
  int w = 0;
  for (int k = n; k < width; ++k) dst[w++] = ' ';
  while (n) dst[w++] = tmp[--n];
  return w;
}
inline int fmt_int(char* dst, long long v, int width) {     // "%{width}d", right-aligned, never truncated
  char tmp[24];
  int n = 0;
  const bool neg = v < 0;
  unsigned long long u = neg ? (unsigned long long)(-v) : (unsigned long long)v;
  do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
  if (neg) tmp[n++] = '-';
  int w = 0;
  for (int k = n; k < width; ++k) dst[w++] = ' ';
  while (n) dst[w++] = tmp[--n];
  return w;
}
}  // namespace

extern "C" int fd_format_pdb(const void* pos_v, int pos_is_f32, const unsigned char* mask, const int* aatype, const double* b_factors, int T,
                             int N, char* out, size_t cap, size_t* len) {
  if (!pos_v || !len || T < 1 || N < 1 || (!out && cap)) return fail(FD_EINVAL, "fd_format_pdb: bad arguments");
  const float* pf = pos_is_f32 ? static_cast<const float*>(pos_v) : nullptr;
  const double* pd = pos_is_f32 ? nullptr : static_cast<const double*>(pos_v);
  for (int i = 0; aatype && i < N; ++i)
    if (aatype[i] < 0 || aatype[i] > 20) return fail(FD_EINVAL, "Invalid aatypes.");
  PdbSink sk{out, cap, 0};
  char buf[256];
  for (int t = 0; t < T; ++t) {
    int n = snprintf(buf, sizeof buf, "MODEL     %d", t + 1);
    sk.line(buf, n);
    long long serial = 1;
    for (int i = 0; i < N; ++i) {
      const char* res = kRes3[aatype ? aatype[i] : 0];
      for (int a = 0; a < 37; ++a) {
        const size_t ia = ((size_t)t * N + i) * 37 + a;
        double p[3];
        bool present;
        if (pf) {      // np.sum(np.abs(pos37), axis=-1) > 1e-7 in float32: a length-3 reduction is a plain left-to-right sum
          const float x = pf[ia * 3], y = pf[ia * 3 + 1], z = pf[ia * 3 + 2];
          present = ((fabsf(x) + fabsf(y)) + fabsf(z)) > 1e-7f;
          p[0] = x; p[1] = y; p[2] = z;
        } else {
          p[0] = pd[ia * 3]; p[1] = pd[ia * 3 + 1]; p[2] = pd[ia * 3 + 2];
          present = ((fabs(p[0]) + fabs(p[1])) + fabs(p[2])) > 1e-7;
        }
        if (mask) present = mask[ia] != 0;
        if (!present) continue;
        const char* nm = kAtomTypes[a];
        char name[8];
        if (strlen(nm) == 4) snprintf(name, sizeof name, "%s", nm); else snprintf(name, sizeof name, " %s", nm);
        const char element[2] = {nm[0], 0};
        // "ATOM  " serial:>5 ' ' name:<4 altloc res:>3 ' ' chain resindex:>4 icode '   ' x y z occupancy b '          ' element:>2 charge:>2
        n = 0;
        memcpy(buf, "ATOM  ", 6); n = 6;
        n += fmt_int(buf + n, serial, 5);
        buf[n++] = ' ';
        { const int l = (int)strlen(name); memcpy(buf + n, name, l); for (int k = l; k < 4; ++k) buf[n + k] = ' '; n += l < 4 ? 4 : l; }
        buf[n++] = ' ';
        memcpy(buf + n, res, 3); n += 3;
        buf[n++] = ' '; buf[n++] = 'A';
        n += fmt_int(buf + n, i, 4);
        buf[n++] = ' '; buf[n++] = ' '; buf[n++] = ' '; buf[n++] = ' ';
        n += fmt_fixed(buf + n, p[0], 8, 3);
        n += fmt_fixed(buf + n, p[1], 8, 3);
        n += fmt_fixed(buf + n, p[2], 8, 3);
        memcpy(buf + n, "  1.00", 6); n += 6;
        n += fmt_fixed(buf + n, b_factors ? b_factors[(size_t)i * 37 + a] : 0.0, 6, 2);
        memset(buf + n, ' ', 10); n += 10;
        buf[n++] = ' '; buf[n++] = element[0];
        buf[n++] = ' '; buf[n++] = ' ';
        sk.line(buf, n);
        ++serial;
      }
    }
    n = snprintf(buf, sizeof buf, "%-6s%5lld      %3s %1s%4d", "TER", serial, kRes3[aatype ? aatype[N - 1] : 0], "A", N - 1);
    sk.line(buf, n);
    sk.line("ENDMDL", 6);
  }
  if (sk.len + 3 <= cap) memcpy(out + sk.len, "END", 3);
  sk.len += 3;
  *len = sk.len;
  if (sk.len > cap) return fail(FD_EINVAL, "fd_format_pdb: output needs %zu bytes, capacity %zu", sk.len, cap);
  return FD_OK;
}
