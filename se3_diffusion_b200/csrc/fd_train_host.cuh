// Training step orchestration (included by fd_engine.cu): the training-mode forward with its tape, the backward pass and the
// C ABI around them.  Reference: Experiment.update_fn / loss_fn (experiments/train_se3_diffusion.py:320-326,524-693) differentiated by
// torch autograd there; here the derivative is written out, stage by stage, exactly as restated and checked on the CPU in
// oracle/manual_backward.py.  Weights are read, and gradients written, in the reference's own state_dict layout (one flat arena in
// schema order, caller-owned — the nn.Module's parameters are views into it), so no packed/fused weight image is involved.
#pragma once

namespace {

// ---- flat parameter arena --------------------------------------------------------------------------------------------------
struct ArenaLayout {
  std::vector<long long> off;      // float offset of parameter i (256-byte aligned)
  std::map<std::string, int> index;
  long long total = 0;
};
inline const ArenaLayout& arena_layout() {
  static ArenaLayout L;
  if (!L.off.empty()) return L;
  const auto& s = param_schema();
  long long o = 0;
  for (size_t i = 0; i < s.size(); ++i) {
    L.off.push_back(o);
    L.index[s[i].name] = (int)i;
    o += (s[i].numel() + 63) / 64 * 64;
  }
  L.total = o;
  return L;
}

struct TBlockTape {
  float *proj, *quat_in, *trans_in, *qp, *kp, *vp, *A, *optg, *zbar, *feats, *ipa_pre, *x320[TF_LAYERS + 1];
  float *qkv[TF_LAYERS], *P[TF_LAYERS], *y[TF_LAYERS], *s1[TF_LAYERS], *x1[TF_LAYERS], *f1[TF_LAYERS], *s2[TF_LAYERS];
  float *n2, *a1, *a2, *n3pre, *node_out, *nb, *pquv, *h1, *h2, *ety, *gamma, *WdT;
};
struct TrainTape {
  int B = 0, N = 0, Np = 0;
  long long rows = 0, edges = 0;
  char* base = nullptr; size_t bytes = 0;
  float *node_in, *temb, *ne_h1, *ne_h2, *ne_y, *node0, *pair, *ee_h1, *ee_h2, *ee_y, *z[NBLK], *quat_fin, *trans_fin, *ha1, *tors;
  TBlockTape blk[NBLK];
  // backward scratch
  float *dnode, *dnode0, *tA, *tB, *tC, *d320a, *d320b, *d320c, *dqkv, *dP, *dfeats, *dproj, *dquat, *dtrans, *dzA, *dzB, *dh384a, *dh384b,
      *dy128, *dA, *dbias, *dzbar, *Gq, *Gk, *dvp, *doptg, *colsum, *dgamma, *RS1, *CS1, *RSy, *CSy, *dnb, *dee;
  // inputs kept for the backward
  const float *rigids_t = nullptr, *res_mask = nullptr, *fixed_mask = nullptr, *gt_psi = nullptr;
  const double* t = nullptr; int t_is_f32 = 0;
  bool valid = false;
};

}  // namespace

// tcgen05 path of the training step's big edge-row GEMMs (fd_train_set_gemm mode 2): forward and dgrad of the EdgeTransition / edge-embedder
// linears run on tc_gemm_kernel (fd_tc.cuh: TMA-staged bf16 hi/lo planes, UMMA, TMEM accumulators, fp32 epilogue).  Activations are split
// into plane scratch right before each GEMM; weight planes (and their transposes for the dgrads) are rebuilt from the fp32 arena each step.
struct TrainTcMat { __nv_bfloat16 *hi = nullptr, *lo = nullptr; CUtensorMap mh, ml; int rows = 0, cols = 0; };
struct TrainTcAct {   // one activation tensor as bf16 hi/lo planes [E, C]: K-major maps (box 64 k x 128 rows) for forward / dgrad, MN-major maps
                      // (box 64 channels x 64 rows) for the weight gradients
  __nv_bfloat16 *hi = nullptr, *lo = nullptr; CUtensorMap kh, kl, nh, nl;
};
struct TrainTcLayer { TrainTcMat w1z, w2, wfx, wft, w2t, dzw; TrainTcAct ph1, ph2, pz; };
struct TrainTc {
  bool weights_ready = false;
  char* warena = nullptr;
  TrainTcLayer et[NBLK - 1];
  TrainTcMat ee2, ee4, ee2t, ee4t;
  // activation plane scratch (sized for the tape's E)
  long long E = 0;
  char* sarena = nullptr;
  __nv_bfloat16 *s0h = nullptr, *s0l = nullptr, *s1h = nullptr, *s1l = nullptr;    // [E,384], [E,128]
  CUtensorMap m0h, m0l, m1h, m1l;
  // weight gradients over the edge rows read the same K-major plane images as MN-major operands (contraction over the rows): second
  // activation scratch for the x operand (h1 / h2 planes [E,384]) and z's planes [E,128], and {64 channels x 64 rows} box maps of all four
  long long Ep = 0;
  __nv_bfloat16 *s2h = nullptr, *s2l = nullptr, *s3h = nullptr, *s3l = nullptr;
  CUtensorMap n0h, n0l, n1h, n1l, n2h, n2l, n3h, n3l, m2h, m2l;
  char* parena = nullptr;          // per-layer persistent planes of h1, h2 [E,384] and z [E,128] (written by the forward, read by the backward)
};

struct fd_train_state {
  float* P = nullptr;      // bound parameter arena (device, caller-owned)
  float* G = nullptr;      // bound gradient arena
  TrainTape tape;
  TrainTc tc;
};

static fd_train_state* train_state(fd_context* h);

namespace {

static void free_tape(TrainTape& T) {
  if (T.base) cudaFree(T.base);
  T = TrainTape();
}

static int ensure_tape(fd_context* h, TrainTape& T, int B, int N) {
  if (T.base && T.B == B && T.N == N) return FD_OK;
  cudaDeviceSynchronize();
  free_tape(T);
  T.B = B; T.N = N; T.Np = (N + 3) & ~3;
  T.rows = (long long)B * N; T.edges = T.rows * N;
  const size_t R = (size_t)T.rows, E = (size_t)T.edges, AT = (size_t)B * H * N * T.Np, PT = (size_t)B * TF_H * N * T.Np;
  struct Item { float** p; size_t n; };
  std::vector<Item> it = {
      {&T.node_in, R * NODE_IN_PAD}, {&T.temb, (size_t)B * 32}, {&T.ne_h1, R * 256}, {&T.ne_h2, R * 256}, {&T.ne_y, R * 256}, {&T.node0, R * 256},
      {&T.pair, E * EDGE_IN}, {&T.ee_h1, E * C_Z}, {&T.ee_h2, E * C_Z}, {&T.ee_y, E * C_Z}, {&T.quat_fin, R * 4}, {&T.trans_fin, R * 3},
      {&T.ha1, R * C_S}, {&T.tors, R * C_S},
      {&T.dnode, R * C_S}, {&T.dnode0, R * C_S}, {&T.tA, R * C_S}, {&T.tB, R * C_S}, {&T.tC, R * C_S}, {&T.d320a, R * TF_D}, {&T.d320b, R * TF_D},
      {&T.d320c, R * TF_D}, {&T.dqkv, R * 3 * TF_D}, {&T.dP, PT}, {&T.dfeats, R * IPA_FEAT}, {&T.dproj, R * PROJ_ALL}, {&T.dquat, R * 4},
      {&T.dtrans, R * 3}, {&T.dzA, E * C_Z}, {&T.dzB, E * C_Z}, {&T.dh384a, E * ET_HID}, {&T.dh384b, E * ET_HID}, {&T.dy128, E * C_Z}, {&T.dA, AT},
      {&T.dbias, E * H}, {&T.dzbar, R * H * C_Z}, {&T.Gq, R * H * PQ * 3}, {&T.Gk, R * H * PQ * 3}, {&T.dvp, R * H * PV * 3},
      {&T.doptg, R * H * PV * 3}, {&T.colsum, (size_t)B * H * N}, {&T.dgamma, 64 + 128}, {&T.RS1, R * ET_HID}, {&T.CS1, R * ET_HID}, {&T.RSy, R * C_Z},
      {&T.CSy, R * C_Z}, {&T.dnb, R * C_Z}, {&T.dee, E * C_Z}};
  for (int b = 0; b < NBLK; ++b) {
    TBlockTape& X = T.blk[b];
    it.push_back({&T.z[b], E * C_Z});
    std::vector<Item> bi = {
        {&X.proj, R * PROJ_ALL}, {&X.quat_in, R * 4}, {&X.trans_in, R * 3}, {&X.qp, R * H * PQ * 3}, {&X.kp, R * H * PQ * 3}, {&X.vp, R * H * PV * 3},
        {&X.A, AT}, {&X.optg, R * H * PV * 3}, {&X.zbar, R * H * C_Z}, {&X.feats, R * IPA_FEAT}, {&X.ipa_pre, R * C_S}, {&X.n2, R * C_S},
        {&X.a1, R * C_S}, {&X.a2, R * C_S}, {&X.n3pre, R * C_S}, {&X.node_out, R * C_S}, {&X.gamma, 64}, {&X.WdT, (size_t)C_Z * 32},
        {&X.nb, b < NBLK - 1 ? R * C_Z : 0}, {&X.pquv, b < NBLK - 1 ? R * ET_NODE : 0}, {&X.h1, b < NBLK - 1 ? E * ET_HID : 0},
        {&X.h2, b < NBLK - 1 ? E * ET_HID : 0}, {&X.ety, b < NBLK - 1 ? E * C_Z : 0}};
    for (auto& x : bi) it.push_back(x);
    for (int l = 0; l <= TF_LAYERS; ++l) it.push_back({&X.x320[l], R * TF_D});
    for (int l = 0; l < TF_LAYERS; ++l) {
      it.push_back({&X.qkv[l], R * 3 * TF_D}); it.push_back({&X.P[l], PT}); it.push_back({&X.y[l], R * TF_D}); it.push_back({&X.s1[l], R * TF_D});
      it.push_back({&X.x1[l], R * TF_D}); it.push_back({&X.f1[l], R * TF_D}); it.push_back({&X.s2[l], R * TF_D});
    }
  }
  size_t total = 0;
  for (auto& x : it) total += al256(x.n * sizeof(float));
  if (cudaMalloc(&T.base, total) != cudaSuccess) {
    cudaGetLastError();
    T = TrainTape();
    return fail(FD_ENOMEM, "training tape allocation of %.1f MB failed for B=%d N=%d", total / 1048576.0, B, N);
  }
  T.bytes = total;
  char* p = T.base;
  for (auto& x : it) { *x.p = x.n ? reinterpret_cast<float*>(p) : nullptr; p += al256(x.n * sizeof(float)); }
  return FD_OK;
}

// ---- launch helpers ------------------------------------------------------------------------------------------------------------
struct TG {
  fd_context* h; cudaStream_t st; const float* P; float* G;
  float* T_tmp128 = nullptr;       // 128-float scratch of the narrow column sums
  int err = 0;
  const float* w(const std::string& n) const { return P + arena_layout().off[arena_layout().index.at(n)]; }
  float* g(const std::string& n) const { return G + arena_layout().off[arena_layout().index.at(n)]; }
  void ck(const char* what) {
    if (err) return;
    h->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) err = fail(FD_ECUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
  }
  void gemm(GemmArgs a, bool b_kmajor, bool a_kmajor = true) {
    if (err) return;
    cudaError_t e = h->train_gemm >= 1 ? launch_gemm_mm3(a, b_kmajor, st, a_kmajor) : launch_gemm(a, b_kmajor, st, a_kmajor);
    h->launches++;
    if (e != cudaSuccess) err = fail(FD_ECUDA, "gemm launch failed: %s", cudaGetErrorString(e));
  }
  // y[M,Nout] = act(x[M,K] W^T + b) (*rowmask) (+ residual)       W [Nout][ldw]
  void lin(const float* x, int ldx, const float* W, int ldw, const float* b, int K, int Nout, float* y, int ldy, long long M, bool relu = false,
           const float* residual = nullptr, int ldr = 0, const float* rowmask = nullptr, bool accumulate = false) {
    GemmArgs a;
    a.A = x; a.lda = ldx; a.B = W; a.ldb = ldw; a.C = y; a.ldc = ldy; a.M = (int)M; a.N = Nout; a.K = K; a.bias = b; a.relu = relu;
    a.residual = residual; a.ldr = ldr; a.rowmask = rowmask; a.accumulate = accumulate;
    gemm(a, true);
  }
  // dx[M,Kin] (+)= dy[M,Nout] W  (W [Nout][ldw]) ; optional ReLU mask by the layer's OUTPUT activations
  void dgrad(const float* dy, int lddy, const float* W, int ldw, int Nout, int Kin, float* dx, int lddx, long long M, bool accumulate = false,
             const float* relumask = nullptr, int ldm = 0) {
    GemmArgs a;
    a.A = dy; a.lda = lddy; a.B = W; a.ldb = ldw; a.C = dx; a.ldc = lddx; a.M = (int)M; a.N = Kin; a.K = Nout; a.accumulate = accumulate;
    a.relumask = relumask; a.ldm = ldm;
    // node-level data gradients with a long reduction (d IPA projections, d linear_out: a few thousand rows, K >= 512) fill a fifth of the
    // SMs with 128x128 tiles: split K over the idle ones and combine the partial tiles with the atomic epilogue
    const long long tiles = ((M + 127) / 128) * ((Kin + 127) / 128);
    if (h->train_gemm >= 1 && !relumask && Nout >= 512 && tiles * 2 <= h->sm_count && !err) {
      int sp = (int)(h->sm_count / tiles);
      if (sp > Nout / 256) sp = Nout / 256;
      if (sp > 1) {
        if (!accumulate) {
          if (cudaMemset2DAsync(dx, (size_t)lddx * sizeof(float), 0, (size_t)Kin * sizeof(float), (size_t)M, st) != cudaSuccess) {
            err = fail(FD_ECUDA, "dgrad: clearing the split-K output failed"); return;
          }
        }
        a.accumulate = false; a.atomic = 1; a.splits = sp;
      }
    }
    gemm(a, false);
  }
  // dW[Nout][lddw] += dy[rows,Nout]^T x[rows,Kin]   (atomic accumulation, split over the rows)
  // mm3 path: the bias gradient can ride on the weight gradient (returns true if it did)
  bool wgrad(const float* dy, int lddy, const float* x, int ldx, float* dW, int lddw, int Nout, int Kin, long long rows, float alpha = 1.f,
             float* db = nullptr) {
    GemmArgs a;
    a.A = dy; a.lda = lddy; a.B = x; a.ldb = ldx; a.C = dW; a.ldc = lddw; a.M = Nout; a.N = Kin; a.K = (int)rows; a.atomic = 1; a.alpha = alpha;
    const bool fused_bias = db != nullptr && h->train_gemm >= 1 && 2.0 * Nout * Kin * (double)rows >= 3.0e7 && Nout >= 32 && Kin >= 16;
    if (fused_bias) a.colsum_out = db;
    const bool big = Nout >= 96 && Kin >= 96;
    const int tile = big ? 128 : 64;
    const long long tiles = (long long)((Nout + tile - 1) / tile) * ((Kin + tile - 1) / tile);
    long long sp = (2LL * h->sm_count + tiles - 1) / tiles;
    const long long maxsp = rows / 256 > 0 ? rows / 256 : 1;
    if (sp > maxsp) sp = maxsp;
    if (sp < 1) sp = 1;
    a.splits = (int)sp;
    gemm(a, false, false);
    return fused_bias;
  }
  void bgrad(const float* dy, int lddy, long long rows, int Nout, float* db) {
    if (err) return;
    const bool narrow = Nout < 64 && lddy == Nout && 128 % Nout == 0 && (rows * Nout) % 128 == 0 && rows >= 4096;
    cudaError_t e = narrow ? launch_colsum_narrow(dy, rows, Nout, db, T_tmp128, st) : launch_colsum(dy, lddy, rows, Nout, db, st);
    h->launches++;
    if (e != cudaSuccess) err = fail(FD_ECUDA, "colsum launch failed: %s", cudaGetErrorString(e));
  }
  // full Linear backward: dW, db accumulate; dx (optional) = dy W
  void lin_bwd(const std::string& name, const float* x, int ldx, const float* dy, int lddy, int Nout, int Kin, long long M, float* dx = nullptr,
               int lddx = 0, bool accumulate = false, const float* relumask = nullptr, int ldm = 0) {
    if (!wgrad(dy, lddy, x, ldx, g(name + ".weight"), Kin, Nout, Kin, M, 1.f, g(name + ".bias"))) bgrad(dy, lddy, M, Nout, g(name + ".bias"));
    if (dx) dgrad(dy, lddy, w(name + ".weight"), Kin, Nout, Kin, dx, lddx, M, accumulate, relumask, ldm);
  }
  void ln(int C, const float* x, int ldx, float* out, int ldo, const std::string& name, long long M, const float* rowmask = nullptr,
          const float* edge_mask = nullptr, int nres = 0) {
    if (err) return;
    LnArgs a;
    a.x = x; a.ldx = ldx; a.out = out; a.ldo = ldo; a.gamma = w(name + ".weight"); a.beta = w(name + ".bias"); a.M = M; a.rowmask = rowmask;
    a.res_mask = edge_mask; a.nres = nres; a.row_offset = 0;
    cudaError_t e = launch_layernorm(C, a, st);
    h->launches++;
    if (e != cudaSuccess) err = fail(FD_ECUDA, "layernorm launch failed: %s", cudaGetErrorString(e));
  }
  void ln_bwd(int C, const float* x, int ldx, const float* dy, int lddy, float* dx, int lddx, const std::string& name, long long M,
              const float* rowmask = nullptr, const float* edge_mask = nullptr, int nres = 0, bool accumulate = false) {
    if (err) return;
    LnBwdArgs a;
    a.x = x; a.ldx = ldx; a.dy = dy; a.lddy = lddy; a.gamma = w(name + ".weight"); a.dx = dx; a.lddx = lddx; a.accumulate = accumulate;
    a.dgamma = g(name + ".weight"); a.dbeta = g(name + ".bias"); a.M = M; a.rowmask = rowmask; a.res_mask = edge_mask; a.nres = nres;
    cudaError_t e = launch_ln_bwd(C, a, h->sm_count, st);
    h->launches++;
    if (e != cudaSuccess) err = fail(FD_ECUDA, "layernorm backward launch failed: %s", cudaGetErrorString(e));
  }
  // ---- tcgen05 path (mode 2) ----
  bool tc_on() const { return h->train_gemm == 2; }
  void pack(TrainTcMat& m, const float* src, int ld, int rows, int cols, bool transpose, int col0 = 0) {
    if (err) return;
    pack_planes_kernel<<<(rows * cols + 255) / 256, 256, 0, st>>>(src, ld, rows, cols, transpose ? 1 : 0, m.hi, m.lo, m.cols, col0);
    ck("pack_planes");
  }
  void split(const float* x, int ld, long long M, int K, __nv_bfloat16* hi, __nv_bfloat16* lo) {
    if (err) return;
    const long long n4 = M * (K / 4);
    split_planes_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(x, ld, M, K, hi, lo);
    ck("split_planes");
  }
  // out[M, W.rows] = epi( [A0 | A1] W^T ) through tc_gemm_kernel; A0 = KB0 k-blocks of map (a0h,a0l), A1 = KB1 k-blocks of (a1h,a1l)
  void tc_gemm(const CUtensorMap& a0h, const CUtensorMap& a0l, int KB0, const CUtensorMap& a1h, const CUtensorMap& a1l, int KB1, const TrainTcMat& W,
               long long M, float* out, int ldo, const float* bias, bool relu, const float* rowadd, int off_i, int off_j, int nres,
               const float* relumask, int ldm) {
    if (err) return;
    TcGemmParams p{};
    p.M = (int)M; p.N = W.rows; p.KB0 = KB0; p.KB1 = KB1; p.planes = 2; p.epi = TC_EPI_F32; p.bias = bias; p.relu = relu ? 1 : 0;
    p.rowadd = rowadd; p.off_i = off_i; p.off_j = off_j; p.ld_rowadd = ET_NODE; p.nres = nres;
    p.m_tiles = (int)((M + TC_BM - 1) / TC_BM);
    const int chunks = W.rows / TC_NC;
    p.nch = 1; p.num_tiles = p.m_tiles * chunks; p.n_valid = W.rows; p.out_f32 = out; p.ldo = ldo; p.relumask = relumask; p.ldm = ldm;
    p.chunk_minor = chunks; p.sa = 4; p.eg = 2;        // A (edge activations) streams from DRAM: four of the six stages
    if (tc_launch_maps(a0h, a0l, a1h, a1l, W.mh, W.ml, p, st, &h->launches)) err = fail(FD_ECUDA, "training tcgen05 GEMM launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  // planes out[M, W.rows] = relu([A0 | A1] W^T + bias + node terms), or — with `mask` — ([A0 | A1] W^T) where mask != 0 and 0 elsewhere (the
  // backward of that ReLU): tc_gemm_kernel's plane-writing epilogue, the edge activations never exist in fp32
  void tc_gemm_planes(const CUtensorMap& a0h, const CUtensorMap& a0l, int KB0, const TrainTcMat& W, long long M, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo,
                      const float* bias, const float* rowadd, int off_i, int off_j, int nres, const __nv_bfloat16* mask) {
    if (err) return;
    TcGemmParams p{};
    p.M = (int)M; p.N = W.rows; p.KB0 = KB0; p.KB1 = 0; p.planes = 2; p.epi = TC_EPI_RELU; p.bias = bias;
    p.rowadd = rowadd; p.off_i = off_i; p.off_j = off_j; p.ld_rowadd = ET_NODE; p.nres = nres; p.out_hi = out_hi; p.out_lo = out_lo; p.maskplane = mask;
    p.sa = 4; p.eg = 2;
    if (tc_launch_maps(a0h, a0l, a0h, a0l, W.mh, W.ml, p, st, &h->launches)) err = fail(FD_ECUDA, "training tcgen05 GEMM (planes) launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  void axis_sum_planes(const __nv_bfloat16* hi, const __nv_bfloat16* lo, float* out, long long R, int N, int Cc, int mode) {
    if (err) return;
    edge_axis_sum_planes_kernel<<<(unsigned)R, 128, 0, st>>>(hi, lo, out, N, Cc, mode);
    ck("edge_axis_sum_planes");
  }
  // dW[Cout][lddw] (first Cin columns) += dy^T x on tcgen05, both operands read IN PLACE as MN-major planes [E rows (K), C columns (MN)]
  // (the K-major images the forward / dgrad GEMMs use): k-slices over the rows = inner batches of tc_gemm_kernel's batched mode, partial
  // tiles combined by the atomic fp32 epilogue
  void tc_wgrad(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl, int Cout, int Cin, long long Ep, float* dW,
                int lddw) {
    if (err) return;
    TcGemmParams p{};
    p.M = Cout; p.N = Cin; p.planes = 2; p.epi = TC_EPI_F32; p.atomic = 1; p.mn_major = 1;
    p.m_tiles = (Cout + TC_BM - 1) / TC_BM; p.nch = Cin / TC_NC;
    p.bat_tiles = p.m_tiles;
    const long long kb_total = Ep / TC_BK;
    long long slices = h->sm_count / p.bat_tiles;
    if (slices > kb_total / 8) slices = kb_total / 8;
    if (slices < 1) slices = 1;
    const long long kb_per = (kb_total + slices - 1) / slices;
    slices = (kb_total + kb_per - 1) / kb_per;
    p.KB0 = (int)kb_per; p.KB1 = 0;
    p.bat_inner = (int)slices; p.num_tiles = p.bat_tiles * (int)slices;
    p.a_row_s0 = 0; p.a_row_s1 = 0; p.a_k_s1 = (int)(kb_per * TC_BK);
    p.b_row_s0 = 0; p.b_row_s1 = 0; p.b_k0 = 0; p.b_k_s1 = (int)(kb_per * TC_BK);
    p.o_s0 = 0; p.o_s1 = 0; p.n_valid = Cin; p.out_f32 = dW; p.ldo = lddw;
    p.sa = p.nch == 1 ? 3 : 2;               // both operands stream from DRAM: one A block per nch B blocks
    if (tc_launch_maps(ah, al, ah, al, bh, bl, p, st, &h->launches)) err = fail(FD_ECUDA, "training tcgen05 weight-gradient launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  void copy(float* dst, int ldd, const float* src, int lds, int C, long long M, const float* rowmask = nullptr, bool accumulate = false) {
    if (err) return;
    copy_cols_kernel<<<(unsigned)((M * C + 255) / 256), 256, 0, st>>>(dst, ldd, src, lds, C, M, rowmask, accumulate ? 1 : 0);
    ck("copy_cols");
  }
  void zero(float* p, long long n) {
    if (err) return;
    zero_f32_kernel<<<(unsigned)((n / 4 + 256) / 256), 256, 0, st>>>(p, n);
    ck("zero");
  }
};

static const std::string kTrunk = "score_model.trunk.";

// ---- tcgen05 helpers of the training path ---------------------------------------------------------------------------------------
static int ttc_alloc_weights(TrainTc& T) {
  if (T.warena) return FD_OK;
  auto bytes = [](int rows, int cols) { return al256((size_t)rows * cols * 2); };
  size_t total = 0;
  const int dims[6][2] = {{ET_HID, C_Z}, {ET_HID, ET_HID}, {C_Z, ET_HID + C_Z}, {ET_HID, C_Z}, {ET_HID, ET_HID}, {C_Z, ET_HID + C_Z}};
  for (int l = 0; l < NBLK - 1; ++l) for (auto& d : dims) total += 2 * bytes(d[0], d[1]);
  total += 4 * 2 * bytes(C_Z, C_Z);
  if (cudaMalloc(&T.warena, total) != cudaSuccess) { cudaGetLastError(); return fail(FD_ENOMEM, "training weight planes allocation failed"); }
  char* p = T.warena;
  auto carve = [&](TrainTcMat& m, int rows, int cols) -> int {
    m.rows = rows; m.cols = cols;
    m.hi = reinterpret_cast<__nv_bfloat16*>(p); p += bytes(rows, cols);
    m.lo = reinterpret_cast<__nv_bfloat16*>(p); p += bytes(rows, cols);
    return tc_make_map(&m.mh, m.hi, rows, cols) | tc_make_map(&m.ml, m.lo, rows, cols);
  };
  int rc = 0;
  for (int l = 0; l < NBLK - 1; ++l) {
    TrainTcLayer& L = T.et[l];
    rc |= carve(L.w1z, ET_HID, C_Z); rc |= carve(L.w2, ET_HID, ET_HID); rc |= carve(L.wfx, C_Z, ET_HID + C_Z);
    rc |= carve(L.wft, ET_HID, C_Z); rc |= carve(L.w2t, ET_HID, ET_HID); rc |= carve(L.dzw, C_Z, ET_HID + C_Z);
  }
  rc |= carve(T.ee2, C_Z, C_Z); rc |= carve(T.ee4, C_Z, C_Z); rc |= carve(T.ee2t, C_Z, C_Z); rc |= carve(T.ee4t, C_Z, C_Z);
  return rc ? fail(FD_ECUDA, "cuTensorMapEncodeTiled failed for the training weight planes") : FD_OK;
}
static int ttc_ensure_scratch(TrainTc& T, long long E) {
  if (T.sarena && T.E == E) return FD_OK;
  if (T.sarena) { cudaDeviceSynchronize(); cudaFree(T.sarena); T.sarena = nullptr; }
  const size_t b0 = al256((size_t)E * ET_HID * 2), b1 = al256((size_t)E * C_Z * 2);
  const long long Ep = (E + 63) / 64 * 64;
  if (cudaMalloc(&T.sarena, 4 * b0 + 4 * b1) != cudaSuccess) { cudaGetLastError(); return fail(FD_ENOMEM, "training plane scratch allocation failed"); }
  T.E = E; T.Ep = Ep;
  char* p = T.sarena;
  T.s0h = reinterpret_cast<__nv_bfloat16*>(p); p += b0; T.s0l = reinterpret_cast<__nv_bfloat16*>(p); p += b0;
  T.s1h = reinterpret_cast<__nv_bfloat16*>(p); p += b1; T.s1l = reinterpret_cast<__nv_bfloat16*>(p); p += b1;
  T.s2h = reinterpret_cast<__nv_bfloat16*>(p); p += b0; T.s2l = reinterpret_cast<__nv_bfloat16*>(p); p += b0;
  T.s3h = reinterpret_cast<__nv_bfloat16*>(p); p += b1; T.s3l = reinterpret_cast<__nv_bfloat16*>(p);
  int rc = tc_make_map(&T.m0h, T.s0h, E, ET_HID) | tc_make_map(&T.m0l, T.s0l, E, ET_HID) | tc_make_map(&T.m1h, T.s1h, E, C_Z) |
           tc_make_map(&T.m1l, T.s1l, E, C_Z) | tc_make_map(&T.m2h, T.s2h, E, ET_HID) | tc_make_map(&T.m2l, T.s2l, E, ET_HID);
  {
    if (T.parena) { cudaFree(T.parena); T.parena = nullptr; }
    if (cudaMalloc(&T.parena, (size_t)(NBLK - 1) * (4 * b0 + 2 * b1)) != cudaSuccess) { cudaGetLastError(); return fail(FD_ENOMEM, "training activation planes allocation failed"); }
    char* q = T.parena;
    auto carve = [&](TrainTcAct& a, int Cc, size_t bytes) {
      a.hi = reinterpret_cast<__nv_bfloat16*>(q); q += bytes; a.lo = reinterpret_cast<__nv_bfloat16*>(q); q += bytes;
      return tc_make_map(&a.kh, a.hi, E, Cc) | tc_make_map(&a.kl, a.lo, E, Cc) | tc_make_map(&a.nh, a.hi, E, Cc, 64) | tc_make_map(&a.nl, a.lo, E, Cc, 64);
    };
    for (int l = 0; l < NBLK - 1; ++l) { rc |= carve(T.et[l].ph1, ET_HID, b0); rc |= carve(T.et[l].ph2, ET_HID, b0); rc |= carve(T.et[l].pz, C_Z, b1); }
  }
  rc |= tc_make_map(&T.n0h, T.s0h, E, ET_HID, 64) | tc_make_map(&T.n0l, T.s0l, E, ET_HID, 64) | tc_make_map(&T.n1h, T.s1h, E, C_Z, 64) |
        tc_make_map(&T.n1l, T.s1l, E, C_Z, 64) | tc_make_map(&T.n2h, T.s2h, E, ET_HID, 64) | tc_make_map(&T.n2l, T.s2l, E, ET_HID, 64) |
        tc_make_map(&T.n3h, T.s3h, E, C_Z, 64) | tc_make_map(&T.n3l, T.s3l, E, C_Z, 64);
  return rc ? fail(FD_ECUDA, "cuTensorMapEncodeTiled failed for the training plane scratch") : FD_OK;
}
static void ttc_free(TrainTc& T) {
  if (T.parena) cudaFree(T.parena);
  if (T.warena) cudaFree(T.warena);
  if (T.sarena) cudaFree(T.sarena);
  T = TrainTc();
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------------------
// training-mode forward
// ---------------------------------------------------------------------------------------------------------------------------------
static int train_forward_impl(fd_context* h, fd_train_state* S, int B, int N, const fd_forward_in* in, const fd_forward_out* out, cudaStream_t st) {
  TrainTape& T = S->tape;
  CKI(ensure_tape(h, T, B, N));
  T.valid = false;
  TG f{h, st, S->P, S->G};
  const long long R = T.rows, E = T.edges;
  const int Np = T.Np;
  const float* res_mask = in->res_mask;
  T.rigids_t = in->rigids_t; T.res_mask = in->res_mask; T.fixed_mask = in->fixed_mask; T.gt_psi = in->gt_psi; T.t = in->t; T.t_is_f32 = in->t_is_f32;
  const std::string e = "embedding_layer.";
  TrainTc& C = S->tc;
  if (f.tc_on()) {
    CKI(ttc_alloc_weights(C));
    CKI(ttc_ensure_scratch(C, E));
    // weight planes of this step (the optimiser changed the fp32 arena): forward forms and the transposed forms the dgrads read
    f.pack(C.ee2, f.w(e + "edge_embedder.2.weight"), C_Z, C_Z, C_Z, false); f.pack(C.ee2t, f.w(e + "edge_embedder.2.weight"), C_Z, C_Z, C_Z, true);
    f.pack(C.ee4, f.w(e + "edge_embedder.4.weight"), C_Z, C_Z, C_Z, false); f.pack(C.ee4t, f.w(e + "edge_embedder.4.weight"), C_Z, C_Z, C_Z, true);
    for (int l = 0; l < NBLK - 1; ++l) {
      const std::string p = kTrunk + "edge_transition_" + std::to_string(l) + ".";
      const float *W1 = f.w(p + "trunk.0.weight"), *W2 = f.w(p + "trunk.2.weight"), *Wf = f.w(p + "final_layer.weight");
      TrainTcLayer& L = C.et[l];
      f.pack(L.w1z, W1, ET_HID, ET_HID, C_Z, false);                  // W1[:, :128]                       [384][128]
      f.pack(L.w2, W2, ET_HID, ET_HID, ET_HID, false);                //                                    [384][384]
      f.pack(L.wfx, Wf, ET_HID, C_Z, ET_HID, false, 0);               // [Wf | Wf[:, :128]]                 [128][512]
      f.pack(L.wfx, Wf, ET_HID, C_Z, C_Z, false, ET_HID);
      f.pack(L.wft, Wf, ET_HID, ET_HID, C_Z, true);                   // Wf^T                               [384][128]
      f.pack(L.w2t, W2, ET_HID, ET_HID, ET_HID, true);                // W2^T                               [384][384]
      f.pack(L.dzw, W1, ET_HID, C_Z, ET_HID, true, 0);                // [W1[:, :128]^T | Wf[:, :128]^T]    [128][512]
      f.pack(L.dzw, Wf, ET_HID, C_Z, C_Z, true, ET_HID);
    }
  }
  // ---- embedders (model/score_network.py:103-154) ----
  node_feats_kernel<<<(unsigned)((R * 16 + 255) / 256), 256, 0, st>>>(in->t, in->t_is_f32, in->fixed_mask, in->seq_idx, T.node_in, T.temb, B, N);
  f.ck("node_feats");
  f.lin(T.node_in, NODE_IN_PAD, f.w(e + "node_embedder.0.weight"), NODE_IN, f.w(e + "node_embedder.0.bias"), NODE_IN, 256, T.ne_h1, 256, R, true);
  f.lin(T.ne_h1, 256, f.w(e + "node_embedder.2.weight"), 256, f.w(e + "node_embedder.2.bias"), 256, 256, T.ne_h2, 256, R, true);
  f.lin(T.ne_h2, 256, f.w(e + "node_embedder.4.weight"), 256, f.w(e + "node_embedder.4.bias"), 256, 256, T.ne_y, 256, R);
  f.ln(256, T.ne_y, 256, T.node0, 256, e + "node_embedder.5", R, res_mask);
  pair_feats_kernel<<<dim3((unsigned)R, (N + 7) / 8), 256, 0, st>>>(T.temb, in->fixed_mask, in->seq_idx, in->sc_ca_t, T.pair, N);
  f.ck("pair_feats");
  f.lin(T.pair, EDGE_IN, f.w(e + "edge_embedder.0.weight"), EDGE_IN, f.w(e + "edge_embedder.0.bias"), EDGE_IN, C_Z, T.ee_h1, C_Z, E, true);
  if (f.tc_on()) {
    f.split(T.ee_h1, C_Z, E, C_Z, C.s1h, C.s1l);
    f.tc_gemm(C.m1h, C.m1l, 2, C.m1h, C.m1l, 0, C.ee2, E, T.ee_h2, C_Z, f.w(e + "edge_embedder.2.bias"), true, nullptr, 0, 0, N, nullptr, 0);
    f.split(T.ee_h2, C_Z, E, C_Z, C.s1h, C.s1l);
    f.tc_gemm(C.m1h, C.m1l, 2, C.m1h, C.m1l, 0, C.ee4, E, T.ee_y, C_Z, f.w(e + "edge_embedder.4.bias"), false, nullptr, 0, 0, N, nullptr, 0);
  } else {
    f.lin(T.ee_h1, C_Z, f.w(e + "edge_embedder.2.weight"), C_Z, f.w(e + "edge_embedder.2.bias"), C_Z, C_Z, T.ee_h2, C_Z, E, true);
    f.lin(T.ee_h2, C_Z, f.w(e + "edge_embedder.4.weight"), C_Z, f.w(e + "edge_embedder.4.bias"), C_Z, C_Z, T.ee_y, C_Z, E);
  }
  f.ln(128, T.ee_y, C_Z, T.z[0], C_Z, e + "edge_embedder.5", E, nullptr, res_mask, N);
  init_frames_kernel<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(in->rigids_t, T.blk[0].quat_in, T.blk[0].trans_in, R);
  f.ck("init_frames");

  for (int b = 0; b < NBLK && !f.err; ++b) {
    TBlockTape& X = T.blk[b];
    const std::string sb = std::to_string(b), ip = kTrunk + "ipa_" + sb + ".";
    const float* s = b == 0 ? T.node0 : T.blk[b - 1].node_out;
    // ---- IPA (model/ipa_pytorch.py:303-471) ----
    ipa_derived_kernel<<<(32 * C_Z + 255) / 256, 256, 0, st>>>(f.w(ip + "head_weights"), f.w(ip + "down_z.weight"), X.gamma, X.WdT);
    f.ck("ipa_derived");
    f.lin(s, C_S, f.w(ip + "linear_q.weight"), C_S, f.w(ip + "linear_q.bias"), C_S, PROJ_Q, X.proj, PROJ_ALL, R);
    f.lin(s, C_S, f.w(ip + "linear_kv.weight"), C_S, f.w(ip + "linear_kv.bias"), C_S, PROJ_KV, X.proj + PROJ_Q, PROJ_ALL, R);
    f.lin(s, C_S, f.w(ip + "linear_q_points.weight"), C_S, f.w(ip + "linear_q_points.bias"), C_S, PROJ_QP, X.proj + PROJ_Q + PROJ_KV, PROJ_ALL, R);
    f.lin(s, C_S, f.w(ip + "linear_kv_points.weight"), C_S, f.w(ip + "linear_kv_points.bias"), C_S, PROJ_KVP, X.proj + PROJ_Q + PROJ_KV + PROJ_QP,
          PROJ_ALL, R);
    ipa_points_kernel<<<(unsigned)R, 224, 0, st>>>(X.proj, X.quat_in, X.trans_in, X.qp, X.kp, X.vp, R);
    f.ck("ipa_points");
    {
      GemmArgs g;   // logits = q k^T sqrt(1/(3C))
      g.A = X.proj; g.lda = PROJ_ALL; g.sA0 = (long long)N * PROJ_ALL; g.sA1 = C_HID;
      g.B = X.proj + PROJ_Q; g.ldb = PROJ_ALL; g.sB0 = (long long)N * PROJ_ALL; g.sB1 = 2 * C_HID;
      g.C = X.A; g.ldc = Np; g.sC0 = (long long)H * N * Np; g.sC1 = (long long)N * Np;
      g.M = N; g.N = N; g.K = C_HID; g.nb0 = B; g.nb1 = H; g.alpha = (float)sqrt(1.0 / (3 * C_HID));
      f.gemm(g, true);
    }
    if (!f.err) {
      const size_t smem = (size_t)(H * Np + H * PQ * 3 + 2 * H * C_Z) * sizeof(float);
      ZRef zr; zr.f32 = T.z[b];
      ipa_edge_kernel<0><<<dim3(N, B), 256, smem, st>>>(zr, X.A, X.qp, X.kp, res_mask, f.w(ip + "linear_b.weight"), f.w(ip + "linear_b.bias"), X.gamma,
                                                        X.WdT, f.w(ip + "down_z.bias"), X.feats, N, Np, X.zbar);
      f.ck("ipa_edge");
    }
    {
      GemmArgs g;   // o = a v
      g.A = X.A; g.lda = Np; g.sA0 = (long long)H * N * Np; g.sA1 = (long long)N * Np;
      g.B = X.proj + PROJ_Q + C_HID; g.ldb = PROJ_ALL; g.sB0 = (long long)N * PROJ_ALL; g.sB1 = 2 * C_HID;
      g.C = X.feats; g.ldc = IPA_FEAT; g.sC0 = (long long)N * IPA_FEAT; g.sC1 = C_HID;
      g.M = N; g.N = C_HID; g.K = N; g.nb0 = B; g.nb1 = H;
      f.gemm(g, false);
      GemmArgs p = g;   // o_pt (global) = a v_pts
      p.B = X.vp; p.ldb = H * PV * 3; p.sB0 = (long long)N * H * PV * 3; p.sB1 = PV * 3;
      p.C = X.optg; p.ldc = H * PV * 3; p.sC0 = (long long)N * H * PV * 3; p.sC1 = PV * 3; p.N = PV * 3;
      f.gemm(p, false);
    }
    ipa_finish_kernel<<<(unsigned)R, 96, 0, st>>>(X.optg, X.quat_in, X.trans_in, X.feats, R);
    f.ck("ipa_finish");
    f.lin(X.feats, IPA_FEAT, f.w(ip + "linear_out.weight"), IPA_FEAT, f.w(ip + "linear_out.bias"), IPA_FEAT, C_S, X.ipa_pre, C_S, R, false, s, C_S, res_mask);
    f.ln(256, X.ipa_pre, C_S, X.x320[0], TF_D, kTrunk + "ipa_ln_" + sb, R);
    f.lin(T.node0, C_S, f.w(kTrunk + "skip_embed_" + sb + ".weight"), C_S, f.w(kTrunk + "skip_embed_" + sb + ".bias"), C_S, C_SKIP, X.x320[0] + C_S, TF_D, R);
    // ---- sequence transformer, autograd semantics (ipa_pytorch.py:584-593,636) ----
    for (int l = 0; l < TF_LAYERS; ++l) {
      const std::string p = kTrunk + "seq_tfmr_" + sb + ".layers." + std::to_string(l) + ".";
      const float* x = X.x320[l];
      f.lin(x, TF_D, f.w(p + "self_attn.in_proj_weight"), TF_D, f.w(p + "self_attn.in_proj_bias"), TF_D, 3 * TF_D, X.qkv[l], 3 * TF_D, R);
      GemmArgs g;
      g.A = X.qkv[l]; g.lda = 3 * TF_D; g.sA0 = (long long)N * 3 * TF_D; g.sA1 = TF_DH;
      g.B = X.qkv[l] + TF_D; g.ldb = 3 * TF_D; g.sB0 = (long long)N * 3 * TF_D; g.sB1 = TF_DH;
      g.C = X.P[l]; g.ldc = Np; g.sC0 = (long long)TF_H * N * Np; g.sC1 = (long long)N * Np;
      g.M = N; g.N = N; g.K = TF_DH; g.nb0 = B; g.nb1 = TF_H; g.alpha = (float)(1.0 / sqrt((double)TF_DH));
      f.gemm(g, true);
      const long long rows = (long long)B * TF_H * N;
      softmax_rows_addmask_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(X.P[l], Np, N, rows, (long long)TF_H * N, res_mask);
      f.ck("softmax_rows_addmask");
      GemmArgs v;
      v.A = X.P[l]; v.lda = Np; v.sA0 = (long long)TF_H * N * Np; v.sA1 = (long long)N * Np;
      v.B = X.qkv[l] + 2 * TF_D; v.ldb = 3 * TF_D; v.sB0 = (long long)N * 3 * TF_D; v.sB1 = TF_DH;
      v.C = X.y[l]; v.ldc = TF_D; v.sC0 = (long long)N * TF_D; v.sC1 = TF_DH;
      v.M = N; v.N = TF_DH; v.K = N; v.nb0 = B; v.nb1 = TF_H;
      f.gemm(v, false);
      f.lin(X.y[l], TF_D, f.w(p + "self_attn.out_proj.weight"), TF_D, f.w(p + "self_attn.out_proj.bias"), TF_D, TF_D, X.s1[l], TF_D, R, false, x, TF_D);
      f.ln(320, X.s1[l], TF_D, X.x1[l], TF_D, p + "norm1", R);
      f.lin(X.x1[l], TF_D, f.w(p + "linear1.weight"), TF_D, f.w(p + "linear1.bias"), TF_D, TF_D, X.f1[l], TF_D, R, true);
      f.lin(X.f1[l], TF_D, f.w(p + "linear2.weight"), TF_D, f.w(p + "linear2.bias"), TF_D, TF_D, X.s2[l], TF_D, R, false, X.x1[l], TF_D);
      f.ln(320, X.s2[l], TF_D, X.x320[l + 1], TF_D, p + "norm2", R);
    }
    f.lin(X.x320[TF_LAYERS], TF_D, f.w(kTrunk + "post_tfmr_" + sb + ".weight"), TF_D, f.w(kTrunk + "post_tfmr_" + sb + ".bias"), TF_D, C_S, X.n2, C_S, R,
          false, X.x320[0], TF_D);
    // ---- node transition ----
    const std::string nt = kTrunk + "node_transition_" + sb + ".";
    f.lin(X.n2, C_S, f.w(nt + "linear_1.weight"), C_S, f.w(nt + "linear_1.bias"), C_S, C_S, X.a1, C_S, R, true);
    f.lin(X.a1, C_S, f.w(nt + "linear_2.weight"), C_S, f.w(nt + "linear_2.bias"), C_S, C_S, X.a2, C_S, R, true);
    f.lin(X.a2, C_S, f.w(nt + "linear_3.weight"), C_S, f.w(nt + "linear_3.bias"), C_S, C_S, X.n3pre, C_S, R, false, X.n2, C_S);
    f.ln(256, X.n3pre, C_S, X.node_out, C_S, nt + "ln", R, res_mask);
    // ---- backbone update (frames of the next block / the final frames) ----
    float* qn = b < NBLK - 1 ? T.blk[b + 1].quat_in : T.quat_fin;
    float* tn = b < NBLK - 1 ? T.blk[b + 1].trans_in : T.trans_fin;
    if (!f.err) {
      cudaMemcpyAsync(qn, X.quat_in, R * 4 * sizeof(float), cudaMemcpyDeviceToDevice, st);
      cudaMemcpyAsync(tn, X.trans_in, R * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st);
      backbone_update_kernel<<<(unsigned)((R + 7) / 8), 256, 0, st>>>(X.node_out, f.w(kTrunk + "bb_update_" + sb + ".linear.weight"),
                                                                      f.w(kTrunk + "bb_update_" + sb + ".linear.bias"), res_mask, in->fixed_mask, qn, tn, R);
      f.ck("backbone_update");
    }
    // ---- edge transition (ipa_pytorch.py:194-233), separable first / last layer ----
    if (b < NBLK - 1) {
      const std::string p = kTrunk + "edge_transition_" + sb + ".";
      const float *W1 = f.w(p + "trunk.0.weight"), *Wf = f.w(p + "final_layer.weight");
      f.lin(X.node_out, C_S, f.w(p + "initial_embed.weight"), C_S, f.w(p + "initial_embed.bias"), C_S, C_Z, X.nb, C_Z, R);
      f.lin(X.nb, C_Z, W1 + C_Z, ET_HID, f.w(p + "trunk.0.bias"), C_Z, ET_HID, X.pquv, ET_NODE, R);                       // P_i (+ b1)
      f.lin(X.nb, C_Z, W1 + 2 * C_Z, ET_HID, nullptr, C_Z, ET_HID, X.pquv + ET_HID, ET_NODE, R);                          // Q_j
      f.lin(X.nb, C_Z, Wf + C_Z, ET_HID, f.w(p + "final_layer.bias"), C_Z, C_Z, X.pquv + 2 * ET_HID, ET_NODE, R);          // U_i (+ bf)
      f.lin(X.nb, C_Z, Wf + 2 * C_Z, ET_HID, nullptr, C_Z, C_Z, X.pquv + 2 * ET_HID + C_Z, ET_NODE, R);                   // V_j
      if (f.tc_on()) {
        TrainTcLayer& L = C.et[b];      // h1, h2 exist only as bf16 hi/lo planes (kept for the backward); z's planes are kept as well
        f.split(T.z[b], C_Z, E, C_Z, L.pz.hi, L.pz.lo);
        f.tc_gemm_planes(L.pz.kh, L.pz.kl, 2, L.w1z, E, L.ph1.hi, L.ph1.lo, nullptr, X.pquv, 0, ET_HID, N, nullptr);                       // h1
        f.tc_gemm_planes(L.ph1.kh, L.ph1.kl, 6, L.w2, E, L.ph2.hi, L.ph2.lo, f.w(p + "trunk.2.bias"), nullptr, 0, 0, N, nullptr);           // h2
        f.tc_gemm(L.ph2.kh, L.ph2.kl, 6, L.pz.kh, L.pz.kl, 2, L.wfx, E, X.ety, C_Z, nullptr, false, X.pquv, 2 * ET_HID, 2 * ET_HID + C_Z, N, nullptr, 0);   // y
      } else {
      GemmArgs g;   // h1 = relu(z W1z^T + P_i + Q_j)
      g.A = T.z[b]; g.lda = C_Z; g.B = W1; g.ldb = ET_HID; g.C = X.h1; g.ldc = ET_HID; g.M = (int)E; g.N = ET_HID; g.K = C_Z; g.relu = 1;
      g.rowadd_i = X.pquv; g.rowadd_j = X.pquv + ET_HID; g.ld_rowadd = ET_NODE; g.nres = N; g.row_offset = 0;
      f.gemm(g, true);
      f.lin(X.h1, ET_HID, f.w(p + "trunk.2.weight"), ET_HID, f.w(p + "trunk.2.bias"), ET_HID, ET_HID, X.h2, ET_HID, E, true);
      GemmArgs y;   // y = z Wfz^T + U_i + V_j
      y.A = T.z[b]; y.lda = C_Z; y.B = Wf; y.ldb = ET_HID; y.C = X.ety; y.ldc = C_Z; y.M = (int)E; y.N = C_Z; y.K = C_Z;
      y.rowadd_i = X.pquv + 2 * ET_HID; y.rowadd_j = X.pquv + 2 * ET_HID + C_Z; y.ld_rowadd = ET_NODE; y.nres = N; y.row_offset = 0;
      f.gemm(y, true);
      f.lin(X.h2, ET_HID, Wf, ET_HID, nullptr, ET_HID, C_Z, X.ety, C_Z, E, false, nullptr, 0, nullptr, true);            // y += h2 Wf^T
      }
      f.ln(128, X.ety, C_Z, T.z[b + 1], C_Z, p + "layer_norm", E, nullptr, res_mask, N);
    }
  }
  // ---- heads ----
  const std::string tp = "score_model.torsion_pred.";
  const float* node = T.blk[NBLK - 1].node_out;
  f.lin(node, C_S, f.w(tp + "linear_1.weight"), C_S, f.w(tp + "linear_1.bias"), C_S, C_S, T.ha1, C_S, R, true);
  f.lin(T.ha1, C_S, f.w(tp + "linear_2.weight"), C_S, f.w(tp + "linear_2.bias"), C_S, C_S, T.tors, C_S, R, false, node, C_S);
  if (!f.err) {
    HeadArgs a;
    a.tors_s = T.tors; a.Wf = f.w(tp + "linear_final.weight"); a.bf = f.w(tp + "linear_final.bias"); a.quat = T.quat_fin; a.trans = T.trans_fin;
    a.rigids_t = in->rigids_t; a.t = in->t; a.t_is_f32 = in->t_is_f32; a.sigma = nullptr; a.sigma_grid = h->d_sigma_grid;
    a.res_mask = res_mask; a.fixed_mask = in->fixed_mask; a.gt_psi = in->gt_psi; a.cached_rows = nullptr; a.omega_grid = h->d_omega;
    a.rot_score = out->rot_score; a.trans_score = out->trans_score; a.psi = out->psi; a.rigids = out->rigids; a.atom37 = out->atom37;
    a.atom14 = out->atom14; a.sc_ca = nullptr; a.rows = R; a.N = N;
    score_head_kernel<<<(unsigned)((R + 7) / 8), 256, 0, st>>>(a);
    f.ck("score_head");
  }
  T.valid = f.err == 0;
  return f.err;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// backward, in four stages (gradient buckets: torsion head + block 3 | block 2 | block 1 | block 0 + embedders), so that the caller can
// start the all-reduce of a finished bucket while the next stage computes
// ---------------------------------------------------------------------------------------------------------------------------------
static void ipa_backward(fd_context* h, TG& f, TrainTape& T, int b, float* d_upd /* [R,256] = d(ipa output), already masked */, float* ds /* [R,256] += */,
                         float* dz, bool accumulate_dz, cudaStream_t st) {
  TBlockTape& X = T.blk[b];
  const long long R = T.rows, E = T.edges;
  const int N = T.N, Np = T.Np, B = T.B;
  const std::string ip = kTrunk + "ipa_" + std::to_string(b) + ".";
  const float c1 = (float)sqrt(1.0 / (3 * C_HID));
  f.lin_bwd(ip + "linear_out", X.feats, IPA_FEAT, d_upd, C_S, C_S, IPA_FEAT, R, T.dfeats, IPA_FEAT);
  if (f.err) return;
  ipa_finish_bwd_kernel<<<(unsigned)R, 96, 0, st>>>(T.dfeats, X.feats, X.optg, X.quat_in, X.trans_in, T.doptg, T.dquat, T.dtrans);
  f.ck("ipa_finish_bwd");
  const float* dopair = T.dfeats + (H * C_HID + 4 * H * PV);
  {
    GemmArgs g;   // dzbar[r,h,:] = dopair[r,h,:] Wd        (down_z.weight [32][128] read as B[k][n])
    g.A = dopair; g.lda = IPA_FEAT; g.sA1 = 32; g.B = f.w(ip + "down_z.weight"); g.ldb = C_Z; g.sB1 = 0; g.C = T.dzbar; g.ldc = H * C_Z; g.sC1 = C_Z;
    g.M = (int)R; g.N = C_Z; g.K = 32; g.nb0 = 1; g.nb1 = H;
    f.gemm(g, false);
    GemmArgs w;   // dWd[d][c] += sum_{r,h} dopair[r,h,d] zbar[r,h,c]
    w.A = dopair; w.lda = IPA_FEAT; w.sA1 = 32; w.B = X.zbar; w.ldb = H * C_Z; w.sB1 = C_Z; w.C = f.g(ip + "down_z.weight"); w.ldc = C_Z; w.sC1 = 0;
    w.M = 32; w.N = C_Z; w.K = (int)R; w.nb0 = 1; w.nb1 = H; w.atomic = 1; w.splits = 4;
    f.gemm(w, false, false);
    for (int hh = 0; hh < H; ++hh) f.bgrad(dopair + hh * 32, IPA_FEAT, R, 32, f.g(ip + "down_z.bias"));      // sum_j a = 1
  }
  {
    GemmArgs g;   // dA = do v^T
    g.A = T.dfeats; g.lda = IPA_FEAT; g.sA0 = (long long)N * IPA_FEAT; g.sA1 = C_HID;
    g.B = X.proj + PROJ_Q + C_HID; g.ldb = PROJ_ALL; g.sB0 = (long long)N * PROJ_ALL; g.sB1 = 2 * C_HID;
    g.C = T.dA; g.ldc = Np; g.sC0 = (long long)H * N * Np; g.sC1 = (long long)N * Np;
    g.M = N; g.N = N; g.K = C_HID; g.nb0 = B; g.nb1 = H;
    f.gemm(g, true);
    GemmArgs p = g;   // dA += doptg vp^T
    p.A = T.doptg; p.lda = H * PV * 3; p.sA0 = (long long)N * H * PV * 3; p.sA1 = PV * 3;
    p.B = X.vp; p.ldb = H * PV * 3; p.sB0 = (long long)N * H * PV * 3; p.sB1 = PV * 3; p.K = PV * 3; p.accumulate = 1;
    f.gemm(p, true);
    GemmArgs v;   // dv[j,h,:] = sum_i a[h,i,j] do[i,h,:]
    v.A = X.A; v.lda = Np; v.sA0 = (long long)H * N * Np; v.sA1 = (long long)N * Np;
    v.B = T.dfeats; v.ldb = IPA_FEAT; v.sB0 = (long long)N * IPA_FEAT; v.sB1 = C_HID;
    v.C = T.dproj + PROJ_Q + C_HID; v.ldc = PROJ_ALL; v.sC0 = (long long)N * PROJ_ALL; v.sC1 = 2 * C_HID;
    v.M = N; v.N = C_HID; v.K = N; v.nb0 = B; v.nb1 = H;
    f.gemm(v, false, false);
    GemmArgs vp = v;  // dvp[j,h,:] = sum_i a doptg[i,h,:]
    vp.B = T.doptg; vp.ldb = H * PV * 3; vp.sB0 = (long long)N * H * PV * 3; vp.sB1 = PV * 3;
    vp.C = T.dvp; vp.ldc = H * PV * 3; vp.sC0 = (long long)N * H * PV * 3; vp.sC1 = PV * 3; vp.N = PV * 3;
    f.gemm(vp, false, false);
  }
  if (f.err) return;
  {
    const size_t smem = (size_t)2 * H * Np * sizeof(float);
    ipa_edge_bwd_kernel<<<dim3(N, B), 256, smem, st>>>(T.z[b], X.A, T.dA, T.dzbar, f.w(ip + "linear_b.weight"), T.dbias, dz, accumulate_dz ? 1 : 0, N, Np);
    f.ck("ipa_edge_bwd");
  }
  f.wgrad(T.dbias, H, T.z[b], C_Z, f.g(ip + "linear_b.weight"), C_Z, H, C_Z, E);
  f.bgrad(T.dbias, H, E, H, f.g(ip + "linear_b.bias"));
  {
    GemmArgs q;   // dq = c1 dL k
    q.A = T.dA; q.lda = Np; q.sA0 = (long long)H * N * Np; q.sA1 = (long long)N * Np;
    q.B = X.proj + PROJ_Q; q.ldb = PROJ_ALL; q.sB0 = (long long)N * PROJ_ALL; q.sB1 = 2 * C_HID;
    q.C = T.dproj; q.ldc = PROJ_ALL; q.sC0 = (long long)N * PROJ_ALL; q.sC1 = C_HID;
    q.M = N; q.N = C_HID; q.K = N; q.nb0 = B; q.nb1 = H; q.alpha = c1;
    f.gemm(q, false);
    GemmArgs k = q;   // dk = c1 dL^T q
    k.B = X.proj; k.sB1 = C_HID; k.C = T.dproj + PROJ_Q; k.sC1 = 2 * C_HID;
    f.gemm(k, false, false);
    GemmArgs gq = q;  // Gq = dL kp
    gq.B = X.kp; gq.ldb = H * PQ * 3; gq.sB0 = (long long)N * H * PQ * 3; gq.sB1 = PQ * 3;
    gq.C = T.Gq; gq.ldc = H * PQ * 3; gq.sC0 = (long long)N * H * PQ * 3; gq.sC1 = PQ * 3; gq.N = PQ * 3; gq.alpha = 1.f;
    f.gemm(gq, false);
    GemmArgs gk = gq; // Gk = dL^T qp
    gk.B = X.qp; gk.C = T.Gk;
    f.gemm(gk, false, false);
  }
  if (f.err) return;
  attn_colsum_kernel<<<B * H, 256, 0, st>>>(T.dA, T.colsum, N, Np);
  f.ck("attn_colsum");
  f.zero(T.dgamma, 64);
  if (f.err) return;
  ipa_points_bwd_kernel<<<(unsigned)R, 224, 0, st>>>(X.proj, X.quat_in, X.qp, X.kp, T.Gq, T.Gk, T.dvp, T.colsum, X.gamma, T.dproj, T.dquat, T.dtrans, T.dgamma, N);
  f.ck("ipa_points_bwd");
  ipa_gamma_bwd_kernel<<<1, 32, 0, st>>>(f.w(ip + "head_weights"), T.dgamma, f.g(ip + "head_weights"));
  f.ck("ipa_gamma_bwd");
  const float* s = b == 0 ? T.node0 : T.blk[b - 1].node_out;
  f.lin_bwd(ip + "linear_q", s, C_S, T.dproj, PROJ_ALL, PROJ_Q, C_S, R, ds, C_S, true);
  f.lin_bwd(ip + "linear_kv", s, C_S, T.dproj + PROJ_Q, PROJ_ALL, PROJ_KV, C_S, R, ds, C_S, true);
  f.lin_bwd(ip + "linear_q_points", s, C_S, T.dproj + PROJ_Q + PROJ_KV, PROJ_ALL, PROJ_QP, C_S, R, ds, C_S, true);
  f.lin_bwd(ip + "linear_kv_points", s, C_S, T.dproj + PROJ_Q + PROJ_KV + PROJ_QP, PROJ_ALL, PROJ_KVP, C_S, R, ds, C_S, true);
}

static void tfmr_backward(fd_context* h, TG& f, TrainTape& T, int b, cudaStream_t st) {
  // in: T.d320a = d x_out ; out: T.d320a = d x_in of layer 0
  TBlockTape& X = T.blk[b];
  const long long R = T.rows;
  const int N = T.N, Np = T.Np, B = T.B;
  for (int l = TF_LAYERS - 1; l >= 0 && !f.err; --l) {
    const std::string p = kTrunk + "seq_tfmr_" + std::to_string(b) + ".layers." + std::to_string(l) + ".";
    f.ln_bwd(320, X.s2[l], TF_D, T.d320a, TF_D, T.d320b, TF_D, p + "norm2", R);                                   // ds2
    f.lin_bwd(p + "linear2", X.f1[l], TF_D, T.d320b, TF_D, TF_D, TF_D, R, T.d320c, TF_D, false, X.f1[l], TF_D);   // df1 (ReLU-masked)
    f.lin_bwd(p + "linear1", X.x1[l], TF_D, T.d320c, TF_D, TF_D, TF_D, R, T.d320b, TF_D, true);                   // dx1 = ds2 + df1 W1
    f.ln_bwd(320, X.s1[l], TF_D, T.d320b, TF_D, T.d320a, TF_D, p + "norm1", R);                                   // ds1
    f.lin_bwd(p + "self_attn.out_proj", X.y[l], TF_D, T.d320a, TF_D, TF_D, TF_D, R, T.d320c, TF_D);               // dy
    const float* qkv = X.qkv[l];
    GemmArgs dp;   // dP = dy v^T
    dp.A = T.d320c; dp.lda = TF_D; dp.sA0 = (long long)N * TF_D; dp.sA1 = TF_DH;
    dp.B = qkv + 2 * TF_D; dp.ldb = 3 * TF_D; dp.sB0 = (long long)N * 3 * TF_D; dp.sB1 = TF_DH;
    dp.C = T.dP; dp.ldc = Np; dp.sC0 = (long long)TF_H * N * Np; dp.sC1 = (long long)N * Np;
    dp.M = N; dp.N = N; dp.K = TF_DH; dp.nb0 = B; dp.nb1 = TF_H;
    f.gemm(dp, true);
    GemmArgs dv;   // dv = P^T dy
    dv.A = X.P[l]; dv.lda = Np; dv.sA0 = (long long)TF_H * N * Np; dv.sA1 = (long long)N * Np;
    dv.B = T.d320c; dv.ldb = TF_D; dv.sB0 = (long long)N * TF_D; dv.sB1 = TF_DH;
    dv.C = T.dqkv + 2 * TF_D; dv.ldc = 3 * TF_D; dv.sC0 = (long long)N * 3 * TF_D; dv.sC1 = TF_DH;
    dv.M = N; dv.N = TF_DH; dv.K = N; dv.nb0 = B; dv.nb1 = TF_H;
    f.gemm(dv, false, false);
    if (f.err) return;
    const long long rows = (long long)B * TF_H * N;
    softmax_rows_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(X.P[l], T.dP, Np, N, rows, (float)(1.0 / sqrt((double)TF_DH)));
    f.ck("softmax_rows_bwd");
    GemmArgs dq;   // dq = dS k
    dq.A = T.dP; dq.lda = Np; dq.sA0 = (long long)TF_H * N * Np; dq.sA1 = (long long)N * Np;
    dq.B = qkv + TF_D; dq.ldb = 3 * TF_D; dq.sB0 = (long long)N * 3 * TF_D; dq.sB1 = TF_DH;
    dq.C = T.dqkv; dq.ldc = 3 * TF_D; dq.sC0 = (long long)N * 3 * TF_D; dq.sC1 = TF_DH;
    dq.M = N; dq.N = TF_DH; dq.K = N; dq.nb0 = B; dq.nb1 = TF_H;
    f.gemm(dq, false);
    GemmArgs dk = dq;   // dk = dS^T q
    dk.B = qkv; dk.C = T.dqkv + TF_D;
    f.gemm(dk, false, false);
    // in_proj
    f.wgrad(T.dqkv, 3 * TF_D, X.x320[l], TF_D, f.g(p + "self_attn.in_proj_weight"), TF_D, 3 * TF_D, TF_D, R);
    f.bgrad(T.dqkv, 3 * TF_D, R, 3 * TF_D, f.g(p + "self_attn.in_proj_bias"));
    f.dgrad(T.dqkv, 3 * TF_D, f.w(p + "self_attn.in_proj_weight"), TF_D, 3 * TF_D, TF_D, T.d320a, TF_D, R, true);   // dx = ds1 + dqkv Win
  }
}

static void edge_transition_backward(fd_context* h, TG& f, TrainTape& T, int b, float* dz_in /* d z[b+1] */, float* dz_out /* d z[b], written */,
                                     cudaStream_t st) {
  TBlockTape& X = T.blk[b];
  const long long R = T.rows, E = T.edges;
  const int N = T.N;
  const std::string p = kTrunk + "edge_transition_" + std::to_string(b) + ".";
  const float *W1 = f.w(p + "trunk.0.weight"), *Wf = f.w(p + "final_layer.weight");
  float *dW1 = f.g(p + "trunk.0.weight"), *dWf = f.g(p + "final_layer.weight");
  TrainTc& C = train_state(h)->tc;
  f.ln_bwd(128, X.ety, C_Z, dz_in, C_Z, T.dy128, C_Z, p + "layer_norm", E, nullptr, T.res_mask, N);                  // dy
  if (f.tc_on()) {
    TrainTcLayer& L = C.et[b];
    f.split(T.dy128, C_Z, E, C_Z, C.s1h, C.s1l);
    f.tc_gemm_planes(C.m1h, C.m1l, 2, L.wft, E, C.s0h, C.s0l, nullptr, nullptr, 0, 0, N, L.ph2.hi);                  // dh2 = (dy Wf) * (h2 > 0) -> planes s0
  } else {
    f.dgrad(T.dy128, C_Z, Wf, ET_HID, C_Z, ET_HID, T.dh384a, ET_HID, E, false, X.h2, ET_HID);
  }
  if (f.tc_on()) {
    // weight gradients over the edge rows on tcgen05: MN-major reads of the plane images (s1 = dy, s0 = dh2, s2 = dh1; h1, h2, z from the forward)
    TrainTcLayer& L = C.et[b];
    f.tc_wgrad(C.n1h, C.n1l, L.ph2.nh, L.ph2.nl, C_Z, ET_HID, C.Ep, dWf, ET_HID);                  // dWf += dy^T h2
    f.tc_wgrad(C.n1h, C.n1l, L.pz.nh, L.pz.nl, C_Z, C_Z, C.Ep, dWf, ET_HID);                       // dWf[:, :128] += dy^T z
    f.axis_sum_planes(C.s0h, C.s0l, T.RS1, R, N, ET_HID, 0);
    f.bgrad(T.RS1, ET_HID, R, ET_HID, f.g(p + "trunk.2.bias"));                                    // sum over the edges of dh2
    f.tc_wgrad(C.n0h, C.n0l, L.ph1.nh, L.ph1.nl, ET_HID, ET_HID, C.Ep, f.g(p + "trunk.2.weight"), ET_HID);   // dW2 += dh2^T h1
    f.tc_gemm_planes(C.m0h, C.m0l, 6, L.w2t, E, C.s2h, C.s2l, nullptr, nullptr, 0, 0, N, L.ph1.hi);          // dh1 = (dh2 W2) * (h1 > 0) -> planes s2
    f.tc_wgrad(C.n2h, C.n2l, L.pz.nh, L.pz.nl, ET_HID, C_Z, C.Ep, dW1, ET_HID);                    // dW1[:, :128] += dh1^T z
  } else {
    if (!f.wgrad(T.dy128, C_Z, X.h2, ET_HID, dWf, ET_HID, C_Z, ET_HID, E, 1.f, f.g(p + "final_layer.bias"))) f.bgrad(T.dy128, C_Z, E, C_Z, f.g(p + "final_layer.bias"));
    f.wgrad(T.dy128, C_Z, T.z[b], C_Z, dWf, ET_HID, C_Z, C_Z, E);
    f.lin_bwd(p + "trunk.2", X.h1, ET_HID, T.dh384a, ET_HID, ET_HID, ET_HID, E, T.dh384b, ET_HID, false, X.h1, ET_HID);   // dh1
  }
  if (f.err) return;
  edge_axis_sum_kernel<<<(unsigned)R, 128, 0, st>>>(T.dy128, T.RSy, N, C_Z, 0); f.ck("edge_axis_sum");
  edge_axis_sum_kernel<<<(unsigned)R, 128, 0, st>>>(T.dy128, T.CSy, N, C_Z, 1); f.ck("edge_axis_sum");
  if (f.tc_on()) {
    f.axis_sum_planes(C.s2h, C.s2l, T.RS1, R, N, ET_HID, 0);
    f.axis_sum_planes(C.s2h, C.s2l, T.CS1, R, N, ET_HID, 1);
  } else {
    edge_axis_sum_kernel<<<(unsigned)R, 128, 0, st>>>(T.dh384b, T.RS1, N, ET_HID, 0); f.ck("edge_axis_sum");
    edge_axis_sum_kernel<<<(unsigned)R, 128, 0, st>>>(T.dh384b, T.CS1, N, ET_HID, 1); f.ck("edge_axis_sum");
  }
  if (!f.tc_on()) f.wgrad(T.dh384b, ET_HID, T.z[b], C_Z, dW1, ET_HID, ET_HID, C_Z, E);
  else f.bgrad(T.RSy, C_Z, R, C_Z, f.g(p + "final_layer.bias"));                                  // sum over the edges = sum over i of the row sums
  f.wgrad(T.RS1, ET_HID, X.nb, C_Z, dW1 + C_Z, ET_HID, ET_HID, C_Z, R);
  f.wgrad(T.CS1, ET_HID, X.nb, C_Z, dW1 + 2 * C_Z, ET_HID, ET_HID, C_Z, R);
  f.bgrad(T.RS1, ET_HID, R, ET_HID, f.g(p + "trunk.0.bias"));
  f.wgrad(T.RSy, C_Z, X.nb, C_Z, dWf + C_Z, ET_HID, C_Z, C_Z, R);
  f.wgrad(T.CSy, C_Z, X.nb, C_Z, dWf + 2 * C_Z, ET_HID, C_Z, C_Z, R);
  if (f.tc_on()) {                                                                         // s2 = dh1 planes, s1 = dy planes (above)
    f.tc_gemm(C.m2h, C.m2l, 6, C.m1h, C.m1l, 2, C.et[b].dzw, E, dz_out, C_Z, nullptr, false, nullptr, 0, 0, N, nullptr, 0);    // dz = dh1 W1[:, :128] + dy Wf[:, :128]
  } else {
    f.dgrad(T.dh384b, ET_HID, W1, ET_HID, ET_HID, C_Z, dz_out, C_Z, E);                    // dz = dh1 W1[:, :128]
    f.dgrad(T.dy128, C_Z, Wf, ET_HID, C_Z, C_Z, dz_out, C_Z, E, true);                     //    + dy Wf[:, :128]
  }
  f.dgrad(T.RS1, ET_HID, W1 + C_Z, ET_HID, ET_HID, C_Z, T.dnb, C_Z, R);
  f.dgrad(T.CS1, ET_HID, W1 + 2 * C_Z, ET_HID, ET_HID, C_Z, T.dnb, C_Z, R, true);
  f.dgrad(T.RSy, C_Z, Wf + C_Z, ET_HID, C_Z, C_Z, T.dnb, C_Z, R, true);
  f.dgrad(T.CSy, C_Z, Wf + 2 * C_Z, ET_HID, C_Z, C_Z, T.dnb, C_Z, R, true);
  f.lin_bwd(p + "initial_embed", X.node_out, C_S, T.dnb, C_Z, C_Z, C_S, R, T.dnode, C_S, true);
}

static int train_backward_impl(fd_context* h, fd_train_state* S, const fd_train_grads* d, int stage_first, int stage_last, cudaStream_t st) {
  TrainTape& T = S->tape;
  if (!T.valid) return fail(FD_ESTATE, "fd_train_backward: no training forward on tape");
  TG f{h, st, S->P, S->G};
  f.T_tmp128 = T.dgamma + 64;
  const long long R = T.rows, E = T.edges;
  const int N = T.N;
  const float* res_mask = T.res_mask;
  for (int stage = stage_first; stage <= stage_last && !f.err; ++stage) {
    const int b = NBLK - 1 - stage;
    TBlockTape& X = T.blk[b];
    const std::string sb = std::to_string(b);
    if (stage == 0) {
      // ---- heads ----
      const std::string tp = "score_model.torsion_pred.";
      HeadBwdArgs a{};
      a.tors_s = T.tors; a.Wf = f.w(tp + "linear_final.weight"); a.bf = f.w(tp + "linear_final.bias"); a.quat = T.quat_fin; a.trans = T.trans_fin;
      a.rigids_t = T.rigids_t; a.t = T.t; a.t_is_f32 = T.t_is_f32; a.sigma_grid = h->d_sigma_grid; a.res_mask = res_mask; a.fixed_mask = T.fixed_mask;
      a.gt_psi = T.gt_psi; a.d_rot = d->d_rot_score; a.d_trans = d->d_trans_score; a.d_rigids = d->d_rigids; a.d_atom37 = d->d_atom37;
      a.d_atom14 = d->d_atom14; a.d_psi = d->d_psi; a.dquat = T.dquat; a.dtrans = T.dtrans; a.dtors = T.tA;
      a.dWf = f.g(tp + "linear_final.weight"); a.dbf = f.g(tp + "linear_final.bias"); a.rows = R; a.N = N;
      const unsigned grid = (unsigned)std::min<long long>((R + 7) / 8, (long long)h->sm_count * 2);
      score_head_bwd_kernel<<<grid, 256, 0, st>>>(a);
      f.ck("score_head_bwd");
      const float* node = X.node_out;
      f.lin_bwd(tp + "linear_2", T.ha1, C_S, T.tA, C_S, C_S, C_S, R, T.tB, C_S, false, T.ha1, C_S);     // da1 (ReLU-masked)
      f.copy(T.dnode, C_S, T.tA, C_S, C_S, R);                                                           // dnode = dhh
      f.lin_bwd(tp + "linear_1", node, C_S, T.tB, C_S, C_S, C_S, R, T.dnode, C_S, true);
      f.zero(T.dnode0, R * C_S);
    }
    // ---- edge transition of this block (its output feeds block b+1) ----
    float* dz_cur = T.dzA;      // holds d z[b+1] on entry of stages > 0 ... swapped below
    if (b < NBLK - 1) {
      // d z[b+1] lives in dzA; write d z[b] (ET part) into dzB, then swap roles by copying the pointer names
      edge_transition_backward(h, f, T, b, T.dzA, T.dzB, st);
      std::swap(T.dzA, T.dzB);
      dz_cur = T.dzA;
    }
    if (f.err) break;
    // ---- backbone update ----
    {
      const unsigned grid = (unsigned)std::min<long long>((R + 7) / 8, (long long)h->sm_count * 2);
      backbone_update_bwd_kernel<<<grid, 256, 0, st>>>(X.node_out, f.w(kTrunk + "bb_update_" + sb + ".linear.weight"),
                                                        f.w(kTrunk + "bb_update_" + sb + ".linear.bias"), res_mask, T.fixed_mask, X.quat_in, T.dquat, T.dtrans,
                                                        T.dnode, f.g(kTrunk + "bb_update_" + sb + ".linear.weight"),
                                                        f.g(kTrunk + "bb_update_" + sb + ".linear.bias"), R);
      f.ck("backbone_update_bwd");
    }
    // ---- node transition ----
    const std::string nt = kTrunk + "node_transition_" + sb + ".";
    f.ln_bwd(256, X.n3pre, C_S, T.dnode, C_S, T.tA, C_S, nt + "ln", R, res_mask);                                  // dy
    f.lin_bwd(nt + "linear_3", X.a2, C_S, T.tA, C_S, C_S, C_S, R, T.tB, C_S, false, X.a2, C_S);                    // da2
    f.lin_bwd(nt + "linear_2", X.a1, C_S, T.tB, C_S, C_S, C_S, R, T.tC, C_S, false, X.a1, C_S);                    // da1
    f.lin_bwd(nt + "linear_1", X.n2, C_S, T.tC, C_S, C_S, C_S, R, T.tA, C_S, true);                                // dn2 = dy + da1 W1
    // ---- post_tfmr + transformer ----
    f.lin_bwd(kTrunk + "post_tfmr_" + sb, X.x320[TF_LAYERS], TF_D, T.tA, C_S, C_S, TF_D, R, T.d320a, TF_D);
    tfmr_backward(h, f, T, b, st);
    f.copy(T.tA, C_S, T.d320a, TF_D, C_S, R, nullptr, true);                                                       // dn1 = dn2 + dx[:, :256]
    f.lin_bwd(kTrunk + "skip_embed_" + sb, T.node0, C_S, T.d320a + C_S, TF_D, C_SKIP, C_S, R, T.dnode0, C_S, true);
    // ---- ipa_ln, IPA ----
    f.ln_bwd(256, X.ipa_pre, C_S, T.tA, C_S, T.tB, C_S, kTrunk + "ipa_ln_" + sb, R);                                // ds_ln -> tB (also the residual path)
    f.copy(T.tC, C_S, T.tB, C_S, C_S, R, res_mask);                                                                // d(ipa out) = ds_ln * mask
    ipa_backward(h, f, T, b, T.tC, T.tB, dz_cur, b < NBLK - 1, st);
    f.copy(T.dnode, C_S, T.tB, C_S, C_S, R);
    if (h->debug && !f.err) {
      snap(h, "dnode_" + sb, T.dnode, R * C_S * 4, st);
      snap(h, "dz_" + sb, dz_cur, (size_t)E * C_Z * 4, st);
      snap(h, "dquat_" + sb, T.dquat, R * 16, st);
      snap(h, "dtrans_" + sb, T.dtrans, R * 12, st);
    }
    if (stage == NBLK - 1 && !f.err) {
      // ---- embedders ----
      const std::string e = "embedding_layer.";
      f.copy(T.dnode0, C_S, T.dnode, C_S, C_S, R, nullptr, true);
      f.ln_bwd(256, T.ne_y, 256, T.dnode0, 256, T.tA, 256, e + "node_embedder.5", R, res_mask);
      f.lin_bwd(e + "node_embedder.4", T.ne_h2, 256, T.tA, 256, 256, 256, R, T.tB, 256, false, T.ne_h2, 256);
      f.lin_bwd(e + "node_embedder.2", T.ne_h1, 256, T.tB, 256, 256, 256, R, T.tC, 256, false, T.ne_h1, 256);
      f.wgrad(T.tC, 256, T.node_in, NODE_IN_PAD, f.g(e + "node_embedder.0.weight"), NODE_IN, 256, NODE_IN, R);
      f.bgrad(T.tC, 256, R, 256, f.g(e + "node_embedder.0.bias"));
      f.ln_bwd(128, T.ee_y, C_Z, dz_cur, C_Z, T.dy128, C_Z, e + "edge_embedder.5", E, nullptr, res_mask, N);
      if (f.tc_on()) {
        TrainTc& C = S->tc;
        f.lin_bwd(e + "edge_embedder.4", T.ee_h2, C_Z, T.dy128, C_Z, C_Z, C_Z, E);                  // weight / bias gradients only
        f.split(T.dy128, C_Z, E, C_Z, C.s1h, C.s1l);
        f.tc_gemm(C.m1h, C.m1l, 2, C.m1h, C.m1l, 0, C.ee4t, E, T.dee, C_Z, nullptr, false, nullptr, 0, 0, N, T.ee_h2, C_Z);
        f.lin_bwd(e + "edge_embedder.2", T.ee_h1, C_Z, T.dee, C_Z, C_Z, C_Z, E);
        f.split(T.dee, C_Z, E, C_Z, C.s1h, C.s1l);
        f.tc_gemm(C.m1h, C.m1l, 2, C.m1h, C.m1l, 0, C.ee2t, E, T.dy128, C_Z, nullptr, false, nullptr, 0, 0, N, T.ee_h1, C_Z);
      } else {
        f.lin_bwd(e + "edge_embedder.4", T.ee_h2, C_Z, T.dy128, C_Z, C_Z, C_Z, E, T.dee, C_Z, false, T.ee_h2, C_Z);
        f.lin_bwd(e + "edge_embedder.2", T.ee_h1, C_Z, T.dee, C_Z, C_Z, C_Z, E, T.dy128, C_Z, false, T.ee_h1, C_Z);
      }
      f.wgrad(T.dy128, C_Z, T.pair, EDGE_IN, f.g(e + "edge_embedder.0.weight"), EDGE_IN, C_Z, EDGE_IN, E);
      f.bgrad(T.dy128, C_Z, E, C_Z, f.g(e + "edge_embedder.0.bias"));
    }
  }
  return f.err;
}
