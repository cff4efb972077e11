// CUDA-core kernels of the FrameDiff forward (everything that is not a big GEMM chain) and of the SE(3) diffuser.
// Each kernel cites the reference lines whose arithmetic it implements (paths relative to /root/reference).
#pragma once
#include <cuda_bf16.h>
#include "fd_common.cuh"
#include "fd_constants.h"

namespace fd {

__constant__ float c_time_freq[16];
__constant__ float c_idx_den[16];
__constant__ float c_dgram_lower[NBINS];
__device__ float g_dgram_lower[32];   // same edges in global memory for per-lane (non-uniform) reads; entries >= NBINS unused
__constant__ float c_pi_f32;

// Edge tensor storage.  fp32 mode: z [E,128] fp32.  Tensor-core modes: bf16 "hi" plane (+ "lo" plane = bf16(z - hi) in
// the 3-term split mode), each [E,128]; z = hi (+ lo).
struct ZRef {
  const float* f32 = nullptr;
  const __nv_bfloat16* hi = nullptr;
  const __nv_bfloat16* lo = nullptr;
};
template <int ZMODE>   // 0 fp32, 1 hi, 2 hi+lo
__device__ __forceinline__ float4 z_load4(const ZRef& z, long long elem /* multiple of 4 */) {
  if (ZMODE == 0) return *reinterpret_cast<const float4*>(z.f32 + elem);
  const uint2 h = *reinterpret_cast<const uint2*>(z.hi + elem);
  float4 r;
  r.x = __uint_as_float(h.x << 16); r.y = __uint_as_float(h.x & 0xffff0000u);
  r.z = __uint_as_float(h.y << 16); r.w = __uint_as_float(h.y & 0xffff0000u);
  if (ZMODE == 2) {
    const uint2 l = *reinterpret_cast<const uint2*>(z.lo + elem);
    r.x += __uint_as_float(l.x << 16); r.y += __uint_as_float(l.x & 0xffff0000u);
    r.z += __uint_as_float(l.y << 16); r.w += __uint_as_float(l.y & 0xffff0000u);
  }
  return r;
}
template <int ZMODE>
__device__ __forceinline__ float z_load1(const ZRef& z, long long elem) {
  if (ZMODE == 0) return z.f32[elem];
  float r = __bfloat162float(z.hi[elem]);
  if (ZMODE == 2) r += __bfloat162float(z.lo[elem]);
  return r;
}
// fp32 -> (hi, lo) bf16 split, round-to-nearest both
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// ----------------------------------------------------------------------------------------------------------------
// LayerNorm over rows of width C (128 / 256 / 320), one warp per row.  torch.nn.LayerNorm, eps 1e-5, biased var.
// Optional row mask (node rows: rowmask[m]; edge rows: res_mask[b,i]*res_mask[b,j]) applied AFTER the norm, and an
// optional second output (the transformer input buffer).
// ----------------------------------------------------------------------------------------------------------------
struct LnArgs {
  const float* x = nullptr; int ldx = 0;
  float* out = nullptr; int ldo = 0;
  float* out2 = nullptr; int ldo2 = 0;
  const float* gamma = nullptr; const float* beta = nullptr;
  long long M = 0;
  const float* rowmask = nullptr;
  const float* res_mask = nullptr; int nres = 0; long long row_offset = 0;
};

template <int C>
__global__ void __launch_bounds__(256) layernorm_kernel(const LnArgs a) {
  constexpr int V = C / 4;                  // float4 per row
  constexpr int PER = (V + 31) / 32;
  const int lane = threadIdx.x & 31;
  const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= a.M) return;
  const float4* xr = reinterpret_cast<const float4*>(a.x + m * a.ldx);
  float4 v[PER];
  float s = 0.f;
#pragma unroll
  for (int p = 0; p < PER; ++p) {
    const int idx = lane + p * 32;
    v[p] = idx < V ? xr[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (v[p].x + v[p].y) + (v[p].z + v[p].w);
  }
  const float mean = warp_sum(s) * (1.f / C);
  float q = 0.f;
#pragma unroll
  for (int p = 0; p < PER; ++p) {
    const int idx = lane + p * 32;
    if (idx < V) {
      const float dx = v[p].x - mean, dy = v[p].y - mean, dz = v[p].z - mean, dw = v[p].w - mean;
      q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + 1e-5f);
  float mk = 1.f;
  if (a.rowmask) mk = a.rowmask[m];
  if (a.res_mask) {
    const long long g = a.row_offset + m, nn = (long long)a.nres * a.nres;
    const long long b = g / nn;
    const int rem = (int)(g - b * nn);
    const int i = rem / a.nres, j = rem - i * a.nres;
    mk = a.res_mask[b * a.nres + i] * a.res_mask[b * a.nres + j];
  }
  const float4* g4 = reinterpret_cast<const float4*>(a.gamma);
  const float4* b4 = reinterpret_cast<const float4*>(a.beta);
  float4* o1 = reinterpret_cast<float4*>(a.out + m * a.ldo);
  float4* o2 = a.out2 ? reinterpret_cast<float4*>(a.out2 + m * a.ldo2) : nullptr;
#pragma unroll
  for (int p = 0; p < PER; ++p) {
    const int idx = lane + p * 32;
    if (idx < V) {
      const float4 g = g4[idx], bt = b4[idx];
      float4 r;
      r.x = ((v[p].x - mean) * rstd * g.x + bt.x) * mk;
      r.y = ((v[p].y - mean) * rstd * g.y + bt.y) * mk;
      r.z = ((v[p].z - mean) * rstd * g.z + bt.z) * mk;
      r.w = ((v[p].w - mean) * rstd * g.w + bt.w) * mk;
      o1[idx] = r;
      if (o2) o2[idx] = r;
    }
  }
}

inline cudaError_t launch_layernorm(int C, const LnArgs& a, cudaStream_t st) {
  const unsigned grid = (unsigned)((a.M + 7) / 8);
  if (C == 128) layernorm_kernel<128><<<grid, 256, 0, st>>>(a);
  else if (C == 256) layernorm_kernel<256><<<grid, 256, 0, st>>>(a);
  else if (C == 320) layernorm_kernel<320><<<grid, 256, 0, st>>>(a);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// ----------------------------------------------------------------------------------------------------------------
// Row softmax with key mask, in place (sequence-transformer attention; model/ipa_pytorch.py:584-593,636).
// S [rows, ld], row r belongs to sample r / rows_per_sample.  Keys with mask <= 0.5 are excluded (eval-mode
// nested-tensor semantics).  Columns [n, ld) are zeroed.  A row with no valid key gives zeros.
// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_rows_kernel(float* S, int ld, int n, long long rows,
                                                           long long rows_per_sample, const float* keymask) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  float* row = S + r * ld;
  const float* km = keymask + (r / rows_per_sample) * n;
  float mx = -INFINITY;
  for (int j = lane; j < n; j += 32)
    if (km[j] > 0.5f) mx = fmaxf(mx, row[j]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < n; j += 32) {
    const float e = km[j] > 0.5f ? expf(row[j] - mx) : 0.f;
    row[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = sum > 0.f ? 1.f / sum : 0.f;
  for (int j = lane; j < ld; j += 32) row[j] = j < n ? row[j] * inv : 0.f;
}

// ----------------------------------------------------------------------------------------------------------------
// Input featurisation (model/score_network.py:14-47,103-136).  One thread per (residue, k<16).
//   node_in [B*N, 68] = [sin t-emb(16) | cos t-emb(16) | fixed | sin idx-emb(16) | cos idx-emb(16) | 0 0 0]
//   temb    [B, 32]   time embedding (shared by the edge embedder's first layer)
// t is passed as double together with a flag saying whether the caller's tensor was fp32 (then t*1e4 is an fp32
// product, as in Experiment.inference_fn) or fp64 (t*1e4 in fp64, then rounded).
// ----------------------------------------------------------------------------------------------------------------
__global__ void node_feats_kernel(const double* __restrict__ t, int t_is_f32, const float* __restrict__ fixed_mask,
                                  const int* __restrict__ seq_idx, float* __restrict__ node_in,
                                  float* __restrict__ temb, int B, int N) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * N * 16;
  if (gid >= total) return;
  const int k = (int)(gid & 15);
  const long long row = gid >> 4;
  const int b = (int)(row / N), i = (int)(row - (long long)b * N);
  const double tt = t[b];
  const float ts = t_is_f32 ? __fmul_rn((float)tt, 10000.f) : (float)(tt * 10000.0);
  const float arg = __fmul_rn(ts, c_time_freq[k]);
  const float st = sinf(arg), ct = cosf(arg);
  float* o = node_in + row * NODE_IN_PAD;
  o[k] = st; o[16 + k] = ct;
  if (i == 0) { temb[b * 32 + k] = st; temb[b * 32 + 16 + k] = ct; }
  const float ia = __fdiv_rn(__fmul_rn((float)seq_idx[row], c_pi_f32), c_idx_den[k]);
  o[33 + k] = sinf(ia); o[49 + k] = cosf(ia);
  if (k == 0) { o[32] = fixed_mask[row]; o[65] = 0.f; o[66] = 0.f; o[67] = 0.f; }
}

// Edge embedder layer-0, node-separable part (model/score_network.py:67-86,133-149):
//   W0·[f_i | f_j | rel | dgram] + b0 = (W0[:,0:33]·f_i + b0) + W0[:,33:66]·f_j + W0[:,66:98]·rel(i-j) + W0[:,98+bin]
// AC [B*N, 256]: [A_i (128, includes b0) | C_j (128)].   w0a/w0c are [33][128] transposed slices.
__global__ void edge_l0_node_terms_kernel(const float* __restrict__ temb, const float* __restrict__ fixed_mask,
                                          const float* __restrict__ w0a, const float* __restrict__ w0c,
                                          const float* __restrict__ b0, float* __restrict__ AC, long long rows, int N) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= rows * 256) return;
  const long long row = gid >> 8;
  const int c2 = (int)(gid & 255), c = c2 & 127;
  const int b = (int)(row / N);
  const float* w = c2 < 128 ? w0a : w0c;
  float acc = c2 < 128 ? b0[c] : 0.f;
  const float* te = temb + b * 32;
#pragma unroll 8
  for (int e = 0; e < 32; ++e) acc = fmaf(w[e * 128 + c], te[e], acc);
  acc = fmaf(w[32 * 128 + c], fixed_mask[row], acc);
  AC[row * 256 + c2] = acc;
}

// rel-offset table  T[d + REL_DMAX][c] = sum_e W0[c, 66+e] * idx_emb(d)[e]   (built once per weight load)
__global__ void rel_table_kernel(const float* __restrict__ w0r /* [32][128] */, float* __restrict__ T) {
  const int d = (int)blockIdx.x - REL_DMAX;
  const int c = threadIdx.x;  // 128
  __shared__ float emb[32];
  if (c < 16) {
    const float ia = __fdiv_rn(__fmul_rn((float)d, c_pi_f32), c_idx_den[c]);
    emb[c] = sinf(ia); emb[16 + c] = cosf(ia);
  }
  __syncthreads();
  float acc = 0.f;
#pragma unroll 8
  for (int e = 0; e < 32; ++e) acc = fmaf(w0r[e * 128 + c], emb[e], acc);
  T[(long long)blockIdx.x * 128 + c] = acc;
}

// distogram bin of a CA–CA distance (data/utils.py:570-580): strict (lower, upper); NBINS = "no bin".
__device__ __forceinline__ int dgram_bin(float d) {
  int bin = NBINS;
#pragma unroll
  for (int k = 0; k < NBINS; ++k) {
    const float lo = c_dgram_lower[k];
    const float hi = k + 1 < NBINS ? c_dgram_lower[k + 1] : 1e8f;
    if (d > lo && d < hi) bin = k;
  }
  return bin;
}

// Edge embedder layer 0 for a chunk of edge rows: h[row][c] = relu(A_i + C_j + T[i-j] + D[bin]).  One warp per row,
// one float4 of channels per lane.  `split` != 0 additionally emits the bf16 hi/lo planes for the tensor-core path.
template <int OMODE>   // 0: fp32 h; 1: bf16 hi plane; 2: hi + lo planes
__global__ void __launch_bounds__(256) edge_embed_l0_kernel(
    const float* __restrict__ AC, const float* __restrict__ T, const float* __restrict__ D /* [NBINS+1][128] */,
    const float* __restrict__ w0r, const int* __restrict__ seq_idx, const float* __restrict__ sc_ca,
    float* __restrict__ h, __nv_bfloat16* __restrict__ h_hi, __nv_bfloat16* __restrict__ h_lo, long long row_offset,
    long long rows, int N) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const long long g = row_offset + r, nn = (long long)N * N;
  const long long b = g / nn;
  const int rem = (int)(g - b * nn);
  const int i = rem / N, j = rem - i * N;
  const long long ri = b * N + i, rj = b * N + j;
  const float dx = sc_ca[ri * 3 + 0] - sc_ca[rj * 3 + 0];
  const float dy = sc_ca[ri * 3 + 1] - sc_ca[rj * 3 + 1];
  const float dz = sc_ca[ri * 3 + 2] - sc_ca[rj * 3 + 2];
  const int bin = dgram_bin(sqrtf(dx * dx + dy * dy + dz * dz));
  const int d = seq_idx[ri] - seq_idx[rj];
  const float4 a = reinterpret_cast<const float4*>(AC + ri * 256)[lane];
  const float4 c = reinterpret_cast<const float4*>(AC + rj * 256 + 128)[lane];
  const float4 dg = reinterpret_cast<const float4*>(D + bin * 128)[lane];
  float4 tr;
  if (d >= -REL_DMAX && d <= REL_DMAX) {
    tr = reinterpret_cast<const float4*>(T + (long long)(d + REL_DMAX) * 128)[lane];
  } else {  // out-of-table offset: evaluate the embedding on the fly
    const int k = lane & 15;
    const float ia = __fdiv_rn(__fmul_rn((float)d, c_pi_f32), c_idx_den[k]);
    const float mine = lane < 16 ? sinf(ia) : cosf(ia);   // lane e holds emb[e]
    tr = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = 0; e < 32; ++e) {
      const float em = __shfl_sync(0xffffffffu, mine, e);
      const float4 w = reinterpret_cast<const float4*>(w0r + e * 128)[lane];
      tr.x = fmaf(w.x, em, tr.x); tr.y = fmaf(w.y, em, tr.y); tr.z = fmaf(w.z, em, tr.z); tr.w = fmaf(w.w, em, tr.w);
    }
  }
  float4 o;
  o.x = fmaxf(((a.x + c.x) + tr.x) + dg.x, 0.f);
  o.y = fmaxf(((a.y + c.y) + tr.y) + dg.y, 0.f);
  o.z = fmaxf(((a.z + c.z) + tr.z) + dg.z, 0.f);
  o.w = fmaxf(((a.w + c.w) + tr.w) + dg.w, 0.f);
  if (OMODE == 0) {
    reinterpret_cast<float4*>(h + r * 128)[lane] = o;
  } else {
    __nv_bfloat16 hi[4], lo[4];
    split_bf16(o.x, hi[0], lo[0]); split_bf16(o.y, hi[1], lo[1]); split_bf16(o.z, hi[2], lo[2]); split_bf16(o.w, hi[3], lo[3]);
    *reinterpret_cast<uint2*>(h_hi + r * 128 + lane * 4) = *reinterpret_cast<const uint2*>(hi);
    if (OMODE == 2) *reinterpret_cast<uint2*>(h_lo + r * 128 + lane * 4) = *reinterpret_cast<const uint2*>(lo);
  }
}

// Same layer for the tensor-core path over the whole edge tensor: grid (B*N query rows, ceil(N/8)), one warp per edge, no integer
// divisions, and the distogram bin found by one ballot (lane k tests bin k) instead of a 22-step scan — the first version of this
// kernel was instruction-issue-bound (68 warp instructions per edge, most of them index arithmetic).
template <int OMODE>   // 1: bf16 hi plane; 2: hi + lo planes
__global__ void __launch_bounds__(256) edge_embed_l0_rows_kernel(
    const float* __restrict__ AC, const float* __restrict__ T, const float* __restrict__ D /* [NBINS+1][128] */,
    const float* __restrict__ w0r, const int* __restrict__ seq_idx, const float* __restrict__ sc_ca,
    __nv_bfloat16* __restrict__ h_hi, __nv_bfloat16* __restrict__ h_lo, int N) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (j >= N) return;
  const long long ri = blockIdx.x;                        // b*N + i
  const long long rj = ri - (ri % N) + j;                 // one 64-bit remainder per warp: b*N + j
  const long long r = ri * N + j;                         // edge row
  const float dx = sc_ca[ri * 3 + 0] - sc_ca[rj * 3 + 0];
  const float dy = sc_ca[ri * 3 + 1] - sc_ca[rj * 3 + 1];
  const float dz = sc_ca[ri * 3 + 2] - sc_ca[rj * 3 + 2];
  const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
  const float lo_e = lane < NBINS ? __ldg(g_dgram_lower + lane) : 3e38f;
  const float hi_e = lane + 1 < NBINS ? __ldg(g_dgram_lower + lane + 1) : 1e8f;
  const unsigned hit = __ballot_sync(0xffffffffu, lane < NBINS && dist > lo_e && dist < hi_e);
  const int bin = hit ? 31 - __clz(hit) : NBINS;          // bins are disjoint; the reference scan keeps the last match
  const int d = seq_idx[ri] - seq_idx[rj];
  const float4 a = reinterpret_cast<const float4*>(AC + ri * 256)[lane];
  const float4 c = reinterpret_cast<const float4*>(AC + rj * 256 + 128)[lane];
  const float4 dg = reinterpret_cast<const float4*>(D + bin * 128)[lane];
  float4 tr;
  if (d >= -REL_DMAX && d <= REL_DMAX) {
    tr = reinterpret_cast<const float4*>(T + (long long)(d + REL_DMAX) * 128)[lane];
  } else {  // out-of-table offset: evaluate the embedding on the fly
    const int k = lane & 15;
    const float ia = __fdiv_rn(__fmul_rn((float)d, c_pi_f32), c_idx_den[k]);
    const float mine = lane < 16 ? sinf(ia) : cosf(ia);   // lane e holds emb[e]
    tr = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = 0; e < 32; ++e) {
      const float em = __shfl_sync(0xffffffffu, mine, e);
      const float4 w = reinterpret_cast<const float4*>(w0r + e * 128)[lane];
      tr.x = fmaf(w.x, em, tr.x); tr.y = fmaf(w.y, em, tr.y); tr.z = fmaf(w.z, em, tr.z); tr.w = fmaf(w.w, em, tr.w);
    }
  }
  const float o0 = fmaxf(((a.x + c.x) + tr.x) + dg.x, 0.f), o1 = fmaxf(((a.y + c.y) + tr.y) + dg.y, 0.f);
  const float o2 = fmaxf(((a.z + c.z) + tr.z) + dg.z, 0.f), o3 = fmaxf(((a.w + c.w) + tr.w) + dg.w, 0.f);
  __nv_bfloat16 hi[4], lo[4];
  split_bf16(o0, hi[0], lo[0]); split_bf16(o1, hi[1], lo[1]); split_bf16(o2, hi[2], lo[2]); split_bf16(o3, hi[3], lo[3]);
  *reinterpret_cast<uint2*>(h_hi + r * 128 + lane * 4) = *reinterpret_cast<const uint2*>(hi);
  if (OMODE == 2) *reinterpret_cast<uint2*>(h_lo + r * 128 + lane * 4) = *reinterpret_cast<const uint2*>(lo);
}

// ----------------------------------------------------------------------------------------------------------------
// IPA: frames applied to the projected points (model/ipa_pytorch.py:346-374; Rigid.apply rigid_utils.py:1104).
// proj [B*N, 6816]; raw points are laid out [x-block | y-block | z-block].  Outputs global-frame points:
//   qp [B*N, H*PQ*3], kp [B*N, H*PQ*3], vp [B*N, H*PV*3]
// ----------------------------------------------------------------------------------------------------------------
__global__ void ipa_points_kernel(const float* __restrict__ proj, const float* __restrict__ quat,
                                  const float* __restrict__ trans, float* __restrict__ qp, float* __restrict__ kp,
                                  float* __restrict__ vp, long long rows) {
  const long long row = blockIdx.x;
  const int p = threadIdx.x;  // 0..223
  if (row >= rows || p >= H * PQ + H * (PQ + PV)) return;
  float q[4], R[9];
#pragma unroll
  for (int k = 0; k < 4; ++k) q[k] = quat[row * 4 + k];
  quat_to_rot(q, R);
  const float tx = trans[row * 3 + 0], ty = trans[row * 3 + 1], tz = trans[row * 3 + 2];
  const float* pr = proj + row * PROJ_ALL;
  float x, y, z;
  float* dst;
  if (p < H * PQ) {
    const float* s = pr + PROJ_Q + PROJ_KV;
    x = s[p]; y = s[H * PQ + p]; z = s[2 * H * PQ + p];
    dst = qp + row * (H * PQ * 3) + p * 3;
  } else {
    const int pp = p - H * PQ;           // h*20 + p'
    const float* s = pr + PROJ_Q + PROJ_KV + PROJ_QP;
    constexpr int NB = H * (PQ + PV);
    x = s[pp]; y = s[NB + pp]; z = s[2 * NB + pp];
    const int hh = pp / (PQ + PV), p2 = pp - hh * (PQ + PV);
    dst = p2 < PQ ? kp + row * (H * PQ * 3) + (hh * PQ + p2) * 3 : vp + row * (H * PV * 3) + (hh * PV + (p2 - PQ)) * 3;
  }
  dst[0] = (R[0] * x + R[1] * y + R[2] * z) + tx;
  dst[1] = (R[3] * x + R[4] * y + R[5] * z) + ty;
  dst[2] = (R[6] * x + R[7] * y + R[8] * z) + tz;
}

// ----------------------------------------------------------------------------------------------------------------
// IPA edge pass for one query residue (b,i), all 8 heads (model/ipa_pytorch.py:376-418,449-456):
//   logits[h][j] = qk[h][j] (pre-scaled, from the batched GEMM) + sqrt(1/3)·(Wb·z_ij + bb)[h]
//                  − ½·γ_h·Σ_p |qp_i − kp_j|² + 1e5·(m_i m_j − 1)
//   a = softmax_j(logits)                     -> written back over L (consumed by the a·v / a·v_pts GEMMs)
//   o_pair[h] = Wd·(Σ_j a[h][j] z_ij) + bd    -> feats[:, 2432 + h*32 + d]   (Σ_j a = 1, so bd factors out)
// z row (N×128 fp32) is streamed twice; the second pass hits L2.
// ----------------------------------------------------------------------------------------------------------------
template <int ZMODE>
__global__ void __launch_bounds__(256) ipa_edge_kernel(
    const ZRef z, float* __restrict__ L, const float* __restrict__ qp, const float* __restrict__ kp,
    const float* __restrict__ res_mask, const float* __restrict__ Wb, const float* __restrict__ bb,
    const float* __restrict__ gamma, const float* __restrict__ WdT, const float* __restrict__ bd,
    float* __restrict__ feats, int N, int Np, float* __restrict__ zbar_out = nullptr /* [B*N,H,128]: kept by the training path */) {
  extern __shared__ __align__(16) float sm[];
  float* lg = sm;                         // [H][Np]
  float* qs = lg + H * Np;                // [192]
  float* zb = qs + H * PQ * 3;            // [2][H][128]
  const int i = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long rowi = (long long)b * N + i;
  const float mi = res_mask[rowi];

  for (int idx = tid; idx < H * Np; idx += 256) {
    const int h = idx / Np, j = idx - h * Np;
    lg[idx] = j < N ? L[(((long long)b * H + h) * N + i) * Np + j] : 0.f;
  }
  if (tid < H * PQ * 3) qs[tid] = qp[rowi * (H * PQ * 3) + tid];
  float wb[H][4];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    const float4 w = reinterpret_cast<const float4*>(Wb + h * C_Z)[lane];
    wb[h][0] = w.x; wb[h][1] = w.y; wb[h][2] = w.z; wb[h][3] = w.w;
  }
  __syncthreads();

  const long long zrow = rowi * N * C_Z;   // element offset of z[b,i,0,0]
  const int myh = lane >> 2;
  const float my_bb = bb[myh], my_g = gamma[myh];
  float q6[6];
#pragma unroll
  for (int e = 0; e < 6; ++e) q6[e] = qs[lane * 6 + e];
  const float k13 = 0.57735026918962576f;  // sqrt(1/3)
  for (int j = warp; j < N; j += 8) {
    const float4 zv = z_load4<ZMODE>(z, zrow + (long long)j * C_Z + lane * 4);
    const float2* kpj = reinterpret_cast<const float2*>(kp + ((long long)b * N + j) * (H * PQ * 3)) + lane * 3;
    const float2 k0 = kpj[0], k1 = kpj[1], k2 = kpj[2];
    float d2;
    {
      const float e0 = q6[0] - k0.x, e1 = q6[1] - k0.y, e2 = q6[2] - k1.x, e3 = q6[3] - k1.y, e4 = q6[4] - k2.x,
                  e5 = q6[5] - k2.y;
      d2 = e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3 + e4 * e4 + e5 * e5;
    }
    d2 += __shfl_xor_sync(0xffffffffu, d2, 1);
    d2 += __shfl_xor_sync(0xffffffffu, d2, 2);
    float pb[H];
#pragma unroll
    for (int h = 0; h < H; ++h) pb[h] = wb[h][0] * zv.x + wb[h][1] * zv.y + wb[h][2] * zv.z + wb[h][3] * zv.w;
    // value-halving butterfly: after xor 16/8/4 each lane owns head (lane>>2)&7
    float v4[4], v2[2], v1;
    {
      const bool up = lane & 16;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float send = up ? pb[k] : pb[k + 4];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
        v4[k] = (up ? pb[k + 4] : pb[k]) + recv;
      }
    }
    {
      const bool up = lane & 8;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float send = up ? v4[k] : v4[k + 2];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 8);
        v2[k] = (up ? v4[k + 2] : v4[k]) + recv;
      }
    }
    {
      const bool up = lane & 4;
      const float send = up ? v2[0] : v2[1];
      const float recv = __shfl_xor_sync(0xffffffffu, send, 4);
      v1 = (up ? v2[1] : v2[0]) + recv;
    }
    v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
    v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
    if ((lane & 3) == 0) {
      const float mj = res_mask[(long long)b * N + j];
      float v = lg[myh * Np + j];
      v += k13 * (v1 + my_bb);
      v += -0.5f * (my_g * d2);
      v += 1e5f * (mi * mj - 1.f);
      lg[myh * Np + j] = v;
    }
  }
  __syncthreads();

  {  // softmax: warp h <-> head h
    float* row = lg + warp * Np;
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) mx = fmaxf(mx, row[j]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int j = lane; j < N; j += 32) {
      const float e = expf(row[j] - mx);
      row[j] = e;
      s += e;
    }
    s = warp_sum(s);
    const float inv = 1.f / s;
    float* Lrow = L + (((long long)b * H + warp) * N + i) * Np;
    for (int j = lane; j < Np; j += 32) {
      const float a = j < N ? row[j] * inv : 0.f;
      row[j] = a;
      Lrow[j] = a;
    }
  }
  __syncthreads();

  {  // zbar[h][c] = sum_j a[h][j] z[i][j][c]
    const int c = tid & 127, half = tid >> 7;
    float acc[H];
#pragma unroll
    for (int h = 0; h < H; ++h) acc[h] = 0.f;
    for (int j0 = half * 4; j0 < N; j0 += 8) {
      float zc[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) zc[jj] = j0 + jj < N ? z_load1<ZMODE>(z, zrow + (long long)(j0 + jj) * C_Z + c) : 0.f;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float4 a4 = *reinterpret_cast<const float4*>(lg + h * Np + j0);
        acc[h] = fmaf(a4.x, zc[0], acc[h]); acc[h] = fmaf(a4.y, zc[1], acc[h]);
        acc[h] = fmaf(a4.z, zc[2], acc[h]); acc[h] = fmaf(a4.w, zc[3], acc[h]);
      }
    }
#pragma unroll
    for (int h = 0; h < H; ++h) zb[(half * H + h) * C_Z + c] = acc[h];
  }
  __syncthreads();
  {  // o_pair = Wd · zbar + bd
    const int h = tid >> 5, d = tid & 31;
    float acc = bd[d];
    const float* z0 = zb + h * C_Z;
    const float* z1 = zb + (H + h) * C_Z;
#pragma unroll 8
    for (int c = 0; c < C_Z; ++c) acc = fmaf(WdT[c * 32 + d], z0[c] + z1[c], acc);
    feats[rowi * IPA_FEAT + (H * C_HID + 4 * H * PV) + h * 32 + d] = acc;
  }
  if (zbar_out) {
    for (int idx = tid; idx < H * C_Z; idx += 256) zbar_out[rowi * (H * C_Z) + idx] = zb[idx] + zb[H * C_Z + idx];
  }
}

// ----------------------------------------------------------------------------------------------------------------
// IPA edge pass, tensor-core modes: the pair bias z·Wb^T (+bb) arrives precomputed from a tcgen05 GEMM over the z planes
// (`pbias` [E, 8] fp32), so z is streamed ONCE here (Σ_j a·z).  One CTA per query residue (b,i), 256 threads:
//   phase 1  thread (j mod 32, h): logits[h][j] = qk + sqrt(1/3)·pbias − ½γ_h·Σ_p|qp_i − kp_j|² + mask     (coalesced kp / pbias)
//   phase 2  warp h: softmax over j, probabilities back to L
//   phase 3  warp w: edges j ≡ w (mod 8), lane = 4 channels: zbar_w[h][c] += a[h][j]·z[i][j][c]; cross-warp sum in smem
//   phase 4  o_pair = Wd·zbar + bd
// ----------------------------------------------------------------------------------------------------------------
template <int ZMODE>
__global__ void __launch_bounds__(256) ipa_edge2_kernel(
    const ZRef z, float* __restrict__ L, const float* __restrict__ pbias, const float* __restrict__ qp, const float* __restrict__ kp,
    const float* __restrict__ res_mask, const float* __restrict__ gamma, const float* __restrict__ WdT, const float* __restrict__ bd,
    float* __restrict__ feats, int N, int Np) {
  extern __shared__ __align__(16) float sm[];
  float* lg = sm;                         // [H][Np]
  float* zb = lg + H * Np;                // [8 warps][H][128]
  const int i = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long rowi = (long long)b * N + i;
  const float mi = res_mask[rowi];
  for (int idx = tid; idx < H * Np; idx += 256) {
    const int h = idx / Np, j = idx - h * Np;
    lg[idx] = j < N ? L[(((long long)b * H + h) * N + i) * Np + j] : 0.f;
  }
  {  // phase 1
    const int h = tid & 7, jl = tid >> 3;
    float q[PQ * 3];
#pragma unroll
    for (int e = 0; e < PQ * 3; ++e) q[e] = qp[rowi * (H * PQ * 3) + h * (PQ * 3) + e];
    const float g = gamma[h];
    __syncthreads();
    for (int j = jl; j < N; j += 32) {
      const float4* kj = reinterpret_cast<const float4*>(kp + ((long long)b * N + j) * (H * PQ * 3) + h * (PQ * 3));
      float d2 = 0.f;
#pragma unroll
      for (int e4 = 0; e4 < 6; ++e4) {
        const float4 k4 = kj[e4];
        const float e0 = q[e4 * 4 + 0] - k4.x, e1 = q[e4 * 4 + 1] - k4.y, e2 = q[e4 * 4 + 2] - k4.z, e3 = q[e4 * 4 + 3] - k4.w;
        d2 += e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
      }
      const float pb = pbias[(rowi * N + j) * H + h];
      const float mj = res_mask[(long long)b * N + j];
      float v = lg[h * Np + j];
      v += 0.57735026918962576f * pb;
      v += -0.5f * (g * d2);
      v += 1e5f * (mi * mj - 1.f);
      lg[h * Np + j] = v;
    }
  }
  __syncthreads();
  {  // phase 2: softmax, warp <-> head
    float* row = lg + warp * Np;
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) mx = fmaxf(mx, row[j]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int j = lane; j < N; j += 32) { const float e = expf(row[j] - mx); row[j] = e; s += e; }
    s = warp_sum(s);
    const float inv = 1.f / s;
    float* Lrow = L + (((long long)b * H + warp) * N + i) * Np;
    for (int j = lane; j < Np; j += 32) { const float a = j < N ? row[j] * inv : 0.f; row[j] = a; Lrow[j] = a; }
  }
  __syncthreads();
  {  // phase 3
    float acc[H][4];
#pragma unroll
    for (int h = 0; h < H; ++h) { acc[h][0] = 0.f; acc[h][1] = 0.f; acc[h][2] = 0.f; acc[h][3] = 0.f; }
    const long long zrow = rowi * N * C_Z;
    for (int j = warp; j < N; j += 8) {
      const float4 zv = z_load4<ZMODE>(z, zrow + (long long)j * C_Z + lane * 4);
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float a = lg[h * Np + j];
        acc[h][0] = fmaf(a, zv.x, acc[h][0]); acc[h][1] = fmaf(a, zv.y, acc[h][1]);
        acc[h][2] = fmaf(a, zv.z, acc[h][2]); acc[h][3] = fmaf(a, zv.w, acc[h][3]);
      }
    }
#pragma unroll
    for (int h = 0; h < H; ++h) *reinterpret_cast<float4*>(zb + (warp * H + h) * C_Z + lane * 4) = make_float4(acc[h][0], acc[h][1], acc[h][2], acc[h][3]);
  }
  __syncthreads();
  // reduce the 8 per-warp partials: thread -> (h, 4 channels)
  {
    const int h = tid >> 5, c4 = (tid & 31) * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const float4 v = *reinterpret_cast<const float4*>(zb + (w * H + h) * C_Z + c4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    __syncthreads();
    *reinterpret_cast<float4*>(zb + h * C_Z + c4) = s;
  }
  __syncthreads();
  {  // phase 4
    const int h = tid >> 5, d = tid & 31;
    float acc = bd[d];
    const float* z0 = zb + h * C_Z;
#pragma unroll 8
    for (int c = 0; c < C_Z; ++c) acc = fmaf(WdT[c * 32 + d], z0[c], acc);
    feats[rowi * IPA_FEAT + (H * C_HID + 4 * H * PV) + h * 32 + d] = acc;
  }
}

// o_pt: global -> local frame, norms (model/ipa_pytorch.py:437-447; Rigid.invert_apply rigid_utils.py:1118).
// optg [B*N, H*PV*3] (global-frame Σ_j a v_pts) -> feats columns [2048 .. 2432): x | y | z | norm, index h*12+p.
__global__ void ipa_finish_kernel(const float* __restrict__ optg, const float* __restrict__ quat,
                                  const float* __restrict__ trans, float* __restrict__ feats, long long rows) {
  const long long row = blockIdx.x;
  const int hp = threadIdx.x;  // 0..95
  if (row >= rows || hp >= H * PV) return;
  float q[4], R[9];
#pragma unroll
  for (int k = 0; k < 4; ++k) q[k] = quat[row * 4 + k];
  quat_to_rot(q, R);
  const float* g = optg + row * (H * PV * 3) + hp * 3;
  const float x = g[0] - trans[row * 3 + 0], y = g[1] - trans[row * 3 + 1], zz = g[2] - trans[row * 3 + 2];
  const float lx = R[0] * x + R[3] * y + R[6] * zz;
  const float ly = R[1] * x + R[4] * y + R[7] * zz;
  const float lz = R[2] * x + R[5] * y + R[8] * zz;
  float* f = feats + row * IPA_FEAT + H * C_HID;
  f[hp] = lx; f[H * PV + hp] = ly; f[2 * H * PV + hp] = lz;
  f[3 * H * PV + hp] = sqrtf((lx * lx + ly * ly + lz * lz) + 1e-8f);
}

// ----------------------------------------------------------------------------------------------------------------
// BackboneUpdate + Rigid.compose_q_update_vec (model/ipa_pytorch.py:530-557,641-644; rigid_utils.py:1039-1063,587-616)
// One warp per residue: upd = Wbb·(node·dm) + b;  q' = normalise(q + dm·(q ⊗ (0,upd[0:3])));  t' = t + dm·R(q)·upd[3:6]
// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) backbone_update_kernel(const float* __restrict__ node, const float* __restrict__ Wbb,
                                                              const float* __restrict__ bbb, const float* __restrict__ res_mask,
                                                              const float* __restrict__ fixed_mask, float* __restrict__ quat,
                                                              float* __restrict__ trans, long long rows) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float dm = (1.f - fixed_mask[row]) * res_mask[row];
  float u[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const float* x = node + row * C_S;
  for (int c = lane; c < C_S; c += 32) {
    const float xv = x[c] * dm;
#pragma unroll
    for (int o = 0; o < 6; ++o) u[o] = fmaf(Wbb[o * C_S + c], xv, u[o]);
  }
#pragma unroll
  for (int o = 0; o < 6; ++o) u[o] = warp_sum(u[o]) + bbb[o];
  if (lane == 0) {
    float q[4], R[9];
#pragma unroll
    for (int k = 0; k < 4; ++k) q[k] = quat[row * 4 + k];
    quat_to_rot(q, R);
    const float v[4] = {0.f, u[0], u[1], u[2]};
    float dq[4];
    quat_mul(q, v, dq);
    float nq[4], n2 = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) { nq[k] = q[k] + dq[k] * dm; n2 += nq[k] * nq[k]; }
    const float nrm = sqrtf(n2);
#pragma unroll
    for (int k = 0; k < 4; ++k) quat[row * 4 + k] = nq[k] / nrm;
    trans[row * 3 + 0] += (R[0] * u[3] + R[1] * u[4] + R[2] * u[5]) * dm;
    trans[row * 3 + 1] += (R[3] * u[3] + R[4] * u[4] + R[5] * u[5]) * dm;
    trans[row * 3 + 2] += (R[6] * u[3] + R[7] * u[4] + R[8] * u[5]) * dm;
  }
}

// split rigids_t [rows,7] -> quat_cur [rows,4], trans_cur [rows,3] (scaled by 0.1; model/ipa_pytorch.py:617-622)
__global__ void init_frames_kernel(const float* __restrict__ rigids, float* __restrict__ quat, float* __restrict__ trans,
                                   long long rows) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
#pragma unroll
  for (int k = 0; k < 4; ++k) quat[r * 4 + k] = rigids[r * 7 + k];
#pragma unroll
  for (int k = 0; k < 3; ++k) trans[r * 3 + k] = rigids[r * 7 + 4 + k] * COORD_SCALE;
}

// ----------------------------------------------------------------------------------------------------------------
// IGSO(3) score of a rotation vector (data/so3_diffuser.py:9-49,71-117,274-305), warp-cooperative.
// Mixed precision exactly as the reference evaluates it: sin/cos((l+½)ω) and the quotient-rule numerator in fp32,
// Gaussian factor and the sums in fp64.  Terms whose Gaussian factor underflows to 0 in fp64 are skipped.
// Returns d/dω log f(ω) (the "omega_scores_t" scalar); all lanes get the result.
// ----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double igso3_score_scalar(float omega, double sigma, int lane) {
  const float half_om = omega * 0.5f;
  const float lo = sinf(half_om);
  const float dlo = 0.5f * cosf(half_om);
  const float lo2 = __fmul_rn(lo, lo);
  const double s2h = sigma * sigma;
  double psum = 0.0, dsum = 0.0;
  for (int l = lane; l < IGSO3_L; l += 32) {
    const double ex = -(double)((long long)l * (l + 1)) * s2h / 2.0;
    if (ex < -745.2) break;   // exp() is exactly 0 below this; l is increasing per lane
    const double gauss = (double)(2 * l + 1) * exp(ex);
    const float lf = (float)l + 0.5f;
    const float arg = __fmul_rn(omega, lf);
    float hi, c;
    sincosf(arg, &hi, &c);
    const float dhi = __fmul_rn(lf, c);
    const float num = __fsub_rn(__fmul_rn(lo, dhi), __fmul_rn(hi, dlo));
    const float quo = __fdiv_rn(num, lo2);
    psum += gauss * (double)hi / (double)lo;
    dsum += gauss * (double)quo;
  }
  psum = warp_sum_d(psum);
  dsum = warp_sum_d(dsum);
  return dsum / (psum + 1e-4);
}

// stand-alone score kernel (SO3Diffuser.torch_score): warp per rotation vector
__global__ void __launch_bounds__(256) igso3_score_kernel(const float* __restrict__ vec, const double* __restrict__ sigma,
                                                          double* __restrict__ out, long long n) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= n) return;
  const float x = vec[r * 3], y = vec[r * 3 + 1], z = vec[r * 3 + 2];
  const float omega = sqrtf(x * x + y * y + z * z) + 1e-6f;
  const double s = igso3_score_scalar(omega, sigma[r], lane);
  if (lane < 3) out[r * 3 + lane] = s * (double)vec[r * 3 + lane] / (double)(omega + 1e-6f);
}

// device-side t -> quantised sigma (so3_diffuser.py:183-213): digitize(sigma(t), grid) - 1
__device__ __forceinline__ double quantise_sigma(double t, const double* __restrict__ grid) {
  const double s = log(t * exp(SO3_MAX_SIGMA) + (1.0 - t) * exp(SO3_MIN_SIGMA));
  int lo = 0, hi = SO3_NSIGMA;  // number of grid points <= s
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (grid[mid] <= s) lo = mid + 1; else hi = mid;
  }
  int idx = lo - 1;
  idx = idx < 0 ? SO3_NSIGMA - 1 : idx;  // numpy's negative index wraps; cannot happen for t in [0,1]
  return grid[idx];
}

// ALA backbone from a frame + psi (data/all_atom.py:152-174, openfold/utils/feats.py:165-228, residue_constants ALA).
// R row-major 3x3, t in Å, (s,c) = (sin psi, cos psi).  atom14 order N,CA,C,O,CB ; atom37 order N,CA,C,CB,O.
__device__ __forceinline__ void backbone_atoms(const float R[9], const float t[3], float s, float c, float* atom37,
                                               float* atom14) {
  auto place = [&](float x, float y, float z, float o[3]) {
    o[0] = (R[0] * x + R[1] * y + R[2] * z) + t[0];
    o[1] = (R[3] * x + R[4] * y + R[5] * z) + t[1];
    o[2] = (R[6] * x + R[7] * y + R[8] * z) + t[2];
  };
  float n[3], ca[3], cc[3], cb[3], o[3];
  place(-0.525f, 1.363f, 0.f, n);
  place(0.f, 0.f, 0.f, ca);
  place(1.526f, -0.f, -0.f, cc);
  place(-0.529f, -0.774f, -1.205f, cb);
  // psi frame: rot = R · (diag(1,-1,-1) · Rx(psi)), trans = R·C + t ; O local = (0.627, 1.062, 0)
  {
    // M = diag(1,-1,-1)·Rx = [[1,0,0],[0,-c,s],[0,-s,-c]];  P = R·M
    float P[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      P[r * 3 + 0] = R[r * 3 + 0];
      P[r * 3 + 1] = R[r * 3 + 1] * (-c) + R[r * 3 + 2] * (-s);
      P[r * 3 + 2] = R[r * 3 + 1] * s + R[r * 3 + 2] * (-c);
    }
    const float x = 0.627f, y = 1.062f, z = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) o[r] = (P[r * 3 + 0] * x + P[r * 3 + 1] * y + P[r * 3 + 2] * z) + cc[r];
  }
  if (atom14) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      atom14[0 + k] = n[k]; atom14[3 + k] = ca[k]; atom14[6 + k] = cc[k]; atom14[9 + k] = o[k]; atom14[12 + k] = cb[k];
    }
    for (int k = 15; k < 42; ++k) atom14[k] = 0.f;
  }
  if (atom37) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      atom37[0 + k] = n[k]; atom37[3 + k] = ca[k]; atom37[6 + k] = cc[k]; atom37[9 + k] = cb[k]; atom37[12 + k] = o[k];
    }
    for (int k = 15; k < 111; ++k) atom37[k] = 0.f;
  }
}

__global__ void compute_backbone_kernel(const float* __restrict__ rigids, const float* __restrict__ psi,
                                        float* __restrict__ atom37, float* __restrict__ atom14, long long n) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  float q[4], R[9], t[3];
#pragma unroll
  for (int k = 0; k < 4; ++k) q[k] = rigids[r * 7 + k];
#pragma unroll
  for (int k = 0; k < 3; ++k) t[k] = rigids[r * 7 + 4 + k];
  quat_to_rot(q, R);
  backbone_atoms(R, t, psi[r * 2], psi[r * 2 + 1], atom37 ? atom37 + r * 111 : nullptr, atom14 ? atom14 + r * 42 : nullptr);
}

// ----------------------------------------------------------------------------------------------------------------
// Score heads (model/ipa_pytorch.py:650-671, model/score_network.py:199-214, data/se3_diffuser.py:115-125,
// data/utils.py:582-599, data/r3_diffuser.py:158-166).  One warp per residue.
//   tors_s [rows,256] = linear_2(relu(linear_1(node))) + node   (two GEMMs upstream)
// ----------------------------------------------------------------------------------------------------------------
struct HeadArgs {
  const float* tors_s; const float* Wf; const float* bf;    // linear_final [2][256], [2]
  const float* quat; const float* trans;                     // predicted frames (trans scaled)
  const float* rigids_t;                                     // input frames [rows,7]
  const double* t; int t_is_f32; const double* sigma;        // per sample; sigma may be null -> quantise on device
  const double* sigma_grid;
  const float* res_mask; const float* fixed_mask; const float* gt_psi;  // gt_psi may be null
  const double* cached_rows;       // use_cached_score (so3_diffuser.py:291-298): [B, 1000] rows of _score_norms at each sample's sigma index, or null
  const double* omega_grid;        // discrete_omega [1000]
  double* rot_score; double* trans_score; float* psi; float* rigids; float* atom37; float* atom14;
  // optional secondary outputs for the sampling loop
  float* sc_ca;                                              // [rows,3] predicted CA (Å) -> next step's self-conditioning
  long long rows; int N;
};

__global__ void __launch_bounds__(256) score_head_kernel(const HeadArgs a) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= a.rows) return;
  const int b = (int)(row / a.N);
  const float m = a.res_mask[row];
  // --- torsion head: unnormalised (2) then L2-normalise with clamp(min=1e-8) --------------------------------------
  float u0 = 0.f, u1 = 0.f;
  const float* s = a.tors_s + row * C_S;
  for (int c = lane; c < C_S; c += 32) {
    const float sv = s[c];
    u0 = fmaf(a.Wf[c], sv, u0);
    u1 = fmaf(a.Wf[C_S + c], sv, u1);
  }
  u0 = warp_sum(u0) + a.bf[0];
  u1 = warp_sum(u1) + a.bf[1];
  const float den = sqrtf(fmaxf(u0 * u0 + u1 * u1, 1e-8f));
  float p0 = u0 / den, p1 = u1 / den;
  {
    const float fm = a.fixed_mask[row];
    const float dmk = 1.f - fm;
    const float g0 = a.gt_psi ? a.gt_psi[row * 2] : 0.f, g1 = a.gt_psi ? a.gt_psi[row * 2 + 1] : 0.f;
    p0 = dmk * p0 + (1.f - dmk) * g0;
    p1 = dmk * p1 + (1.f - dmk) * g1;
  }
  // --- rotation score --------------------------------------------------------------------------------------------
  float q0[4], qt[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { q0[k] = a.quat[row * 4 + k]; qt[k] = a.rigids_t[row * 7 + k]; }
  float rv[3], omega;
  {
    const float n2 = q0[0] * q0[0] + q0[1] * q0[1] + q0[2] * q0[2] + q0[3] * q0[3];
    const float qi[4] = {q0[0] / n2, -q0[1] / n2, -q0[2] / n2, -q0[3] / n2};
    float qr[4];
    quat_mul(qi, qt, qr);
    if (qr[0] < 0.f) { qr[0] = -qr[0]; qr[1] = -qr[1]; qr[2] = -qr[2]; qr[3] = -qr[3]; }
    const float vn = sqrtf(qr[1] * qr[1] + qr[2] * qr[2] + qr[3] * qr[3]);
    const float angle = 2.f * atan2f(vn, qr[0]);
    const float a2 = angle * angle;
    const float small = 2.f + a2 / 12.f + 7.f * a2 * a2 / 2880.f;
    const float large = angle / sinf(angle / 2.f + 1e-6f);
    const float sc = angle <= 1e-3f ? small : large;
    rv[0] = sc * qr[1]; rv[1] = sc * qr[2]; rv[2] = sc * qr[3];
    omega = sqrtf(rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]) + 1e-6f;
  }
  const double tt = a.t[b];
  const double sig = a.sigma ? a.sigma[b] : quantise_sigma(tt, a.sigma_grid);
  double sscal;
  if (a.cached_rows) {
    // torch.bucketize(omega, discrete_omega[:-1]) (right = False): number of the first 999 grid points strictly below omega; then a gather
    int lo = 0, hi = SO3_NOMEGA - 1;
    const double om = (double)omega;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a.omega_grid[mid] < om) lo = mid + 1; else hi = mid; }
    sscal = a.cached_rows[(long long)b * SO3_NOMEGA + lo];
  } else {
    sscal = igso3_score_scalar(omega, sig, lane);
  }
  // --- translation score: -(x_t·0.1 − e^{−β̄/2}·x̂0·0.1) / (1 − e^{−β̄}) ----------------------------------------------
  if (lane < 3) {
    a.rot_score[row * 3 + lane] = sscal * (double)rv[lane] / (double)(omega + 1e-6f) * (double)m;
    const float x0s = a.trans[row * 3 + lane];                       // predicted, scaled units
    const float x0 = x0s / COORD_SCALE;                              // unscale (Å) — this is what `rigids` carries
    const float xt = a.rigids_t[row * 7 + 4 + lane];
    double ts;
    if (a.t_is_f32) {
      const float tf = (float)tt;
      const float beta = tf * (float)R3_MIN_B + 0.5f * (tf * tf) * (float)(R3_MAX_B - R3_MIN_B);
      const float e1 = expf(-0.5f * beta), var = 1.f - expf(-beta);
      ts = (double)(-(__fmul_rn(xt, COORD_SCALE) - e1 * __fmul_rn(x0, COORD_SCALE)) / var);
    } else {
      const double beta = tt * R3_MIN_B + 0.5 * (tt * tt) * (R3_MAX_B - R3_MIN_B);
      const double e1 = exp(-0.5 * beta), var = 1.0 - exp(-beta);
      ts = -((double)__fmul_rn(xt, COORD_SCALE) - e1 * (double)__fmul_rn(x0, COORD_SCALE)) / var;
    }
    a.trans_score[row * 3 + lane] = ts * (double)m;
    a.rigids[row * 7 + 4 + lane] = x0;
    if (a.sc_ca) a.sc_ca[row * 3 + lane] = x0;
  }
  if (lane == 0) {
    a.psi[row * 2] = p0; a.psi[row * 2 + 1] = p1;
#pragma unroll
    for (int k = 0; k < 4; ++k) a.rigids[row * 7 + k] = q0[k];
    if (a.atom37 || a.atom14) {
      float R[9];
      quat_to_rot(q0, R);
      const float t3[3] = {a.trans[row * 3] / COORD_SCALE, a.trans[row * 3 + 1] / COORD_SCALE,
                           a.trans[row * 3 + 2] / COORD_SCALE};
      backbone_atoms(R, t3, p0, p1, a.atom37 ? a.atom37 + row * 111 : nullptr, a.atom14 ? a.atom14 + row * 42 : nullptr);
    }
  }
}

// ----------------------------------------------------------------------------------------------------------------
// One reverse-SDE step (data/se3_diffuser.py:160-214; so3_diffuser.py:330-366; r3_diffuser.py:106-146).
// One CTA per backbone.  State is the fp32 7-vector (the reference rounds to fp32 every step: Rigid.__init__ casts),
// the arithmetic in between is fp64 like numpy/scipy.  Rotation update is quaternion-native: q' = q ⊗ exp(δ/2), equal
// to scipy's rotvec→matrix→product→rotvec round trip up to fp64 rounding.
// ----------------------------------------------------------------------------------------------------------------
struct StepSched {   // per-step scalars, precomputed on the host in double (numpy-equivalent formulas)
  double t, g_so3, b_r3, dt;
};

struct ReverseArgs {
  float* rigids;                       // [B,N,7] in/out
  const double* rot_score; const double* trans_score;   // [B,N,3]
  const float* diffuse_mask;           // [B,N] or null
  const float* res_mask; const float* fixed_mask;       // used when diffuse_mask is null and use_masks != 0
  int use_masks;
  const double* z_rot; const double* z_trans;           // injected N(0,1) [B,N,3] or null
  const StepSched* sched; const int* step_ptr; int step_fixed;   // schedule entry = sched[step_ptr ? *step_ptr : step_fixed]
  long long noise_stride; int rng_step_bias;            // injected noise slice = step*noise_stride; Philox step = step + bias
  uint64_t seed; long long first_sample;
  int center; double noise_scale;
  float* rotmat_out;                   // [B,N,9] or null
  int N;
};

__global__ void __launch_bounds__(256) reverse_step_kernel(const ReverseArgs a) {
  extern __shared__ double xs[];       // [N][3] updated (scaled) translations
  __shared__ double red[3][8];
  const int b = blockIdx.x, tid = threadIdx.x, N = a.N;
  const int step = a.step_ptr ? *a.step_ptr : a.step_fixed;
  const StepSched sc = a.sched[step];
  const int rstep = step + a.rng_step_bias;
  const double* zrp = a.z_rot ? a.z_rot + (long long)step * a.noise_stride : nullptr;
  const double* zxp = a.z_rot ? a.z_trans + (long long)step * a.noise_stride : nullptr;
  const double sdt = sqrt(sc.dt);
  const double g = sc.g_so3, g_r3 = sqrt(sc.b_r3);
  double cs[3] = {0.0, 0.0, 0.0};
  Philox rng(a.seed);
  for (int i = tid; i < N; i += 256) {
    const long long row = (long long)b * N + i;
    double zr[3], zx[3];
    if (zrp) {
#pragma unroll
      for (int k = 0; k < 3; ++k) { zr[k] = zrp[row * 3 + k]; zx[k] = zxp[row * 3 + k]; }
    } else {
      const unsigned long long gs = (unsigned long long)(a.first_sample + b);
      uint32_t r4[4];
      double n0, n1;
      rng((uint32_t)i, (uint32_t)rstep * 4u + 0u, (uint32_t)gs, (uint32_t)(gs >> 32), r4); normal2(r4, n0, n1); zr[0] = n0; zr[1] = n1;
      rng((uint32_t)i, (uint32_t)rstep * 4u + 1u, (uint32_t)gs, (uint32_t)(gs >> 32), r4); normal2(r4, n0, n1); zr[2] = n0; zx[0] = n1;
      rng((uint32_t)i, (uint32_t)rstep * 4u + 2u, (uint32_t)gs, (uint32_t)(gs >> 32), r4); normal2(r4, n0, n1); zx[1] = n0; zx[2] = n1;
    }
    // ---- rotation: geodesic random walk, right-multiply -----------------------------------------------------------
    double q[4];
    {
      double n2 = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) { q[k] = (double)a.rigids[row * 7 + k]; n2 += q[k] * q[k]; }
      const double inv = 1.0 / sqrt(n2);
#pragma unroll
      for (int k = 0; k < 4; ++k) q[k] *= inv;
    }
    double d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) d[k] = (g * g) * a.rot_score[row * 3 + k] * sc.dt + g * sdt * (a.noise_scale * zr[k]);
    const double ang = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    double e[4];
    if (ang < 1e-3) {  // scipy's small-angle branch: sin(a/2)/a ≈ 1/2 − a²/48 + a⁴/3840
      const double a2 = ang * ang;
      const double sc2 = 0.5 - a2 / 48.0 + a2 * a2 / 3840.0;
      e[0] = cos(ang * 0.5); e[1] = sc2 * d[0]; e[2] = sc2 * d[1]; e[3] = sc2 * d[2];
    } else {
      const double sh = sin(ang * 0.5) / ang;
      e[0] = cos(ang * 0.5); e[1] = sh * d[0]; e[2] = sh * d[1]; e[3] = sh * d[2];
    }
    double qn[4];
    quat_mul(q, e, qn);
    {
      const double inv = 1.0 / sqrt(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
#pragma unroll
      for (int k = 0; k < 4; ++k) qn[k] *= inv;
    }
    float dmf = 1.f;
    if (a.diffuse_mask) dmf = a.diffuse_mask[row];
    else if (a.use_masks) dmf = (1.f - a.fixed_mask[row]) * a.res_mask[row];
    const bool upd = dmf > 0.5f;
    const double* qo = upd ? qn : q;
#pragma unroll
    for (int k = 0; k < 4; ++k) a.rigids[row * 7 + k] = (float)qo[k];
    if (a.rotmat_out) {
      double Rd[9];
      quat_to_rot_d(qo, Rd);
#pragma unroll
      for (int k = 0; k < 9; ++k) a.rotmat_out[row * 9 + k] = (float)Rd[k];
    }
    // ---- translation: Euler–Maruyama on the VP-SDE in scaled units ---------------------------------------------------
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double x = (double)__fmul_rn(a.rigids[row * 7 + 4 + k], COORD_SCALE);   // numpy keeps fp32 for x_t·0.1
      const double f = -0.5 * sc.b_r3 * x;
      const double pert = (f - g_r3 * g_r3 * a.trans_score[row * 3 + k]) * sc.dt + g_r3 * sdt * (a.noise_scale * zx[k]);
      const double x1 = x - pert;
      xs[i * 3 + k] = x1;
      cs[k] += x1;
    }
  }
  // centre of mass over ALL residues (mask is not forwarded by the reference)
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double w = warp_sum_d(cs[k]);
    if ((tid & 31) == 0) red[k][tid >> 5] = w;
  }
  __syncthreads();
  double com[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[k][w];
    com[k] = a.center ? s / (double)N : 0.0;
  }
  for (int i = tid; i < N; i += 256) {
    const long long row = (long long)b * N + i;
    float dmf = 1.f;
    if (a.diffuse_mask) dmf = a.diffuse_mask[row];
    else if (a.use_masks) dmf = (1.f - a.fixed_mask[row]) * a.res_mask[row];
    const double dm = (double)dmf;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double x1 = (xs[i * 3 + k] - com[k]) / (double)0.1;
      const double xt = (double)a.rigids[row * 7 + 4 + k];
      a.rigids[row * 7 + 4 + k] = (float)(dm * x1 + (1.0 - dm) * xt);
    }
  }
}

// ----------------------------------------------------------------------------------------------------------------
// Prior sample (data/se3_diffuser.py:216-268; so3_diffuser.py:215-252; r3_diffuser.py:39).  Thread per residue.
// cdf/omega: the IGSO(3) angle CDF row for t = 1 and its abscissae (1000 each).
// ----------------------------------------------------------------------------------------------------------------
__global__ void sample_ref_kernel(const double* __restrict__ z_axis, const double* __restrict__ u_angle,
                                  const double* __restrict__ z_trans, uint64_t seed, long long first_sample,
                                  int per_sample, const double* __restrict__ cdf, const double* __restrict__ omega,
                                  float* __restrict__ rigids, long long n) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double ax[3], u, zt[3];
  if (z_axis) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { ax[k] = z_axis[r * 3 + k]; zt[k] = z_trans[r * 3 + k]; }
    u = u_angle[r];
  } else {
    const long long smp = first_sample + r / per_sample;
    const uint32_t res = (uint32_t)(r % per_sample);
    Philox rng(seed);
    uint32_t r4[4];
    double n0, n1;
    const uint32_t s_lo = (uint32_t)smp, s_hi = (uint32_t)((unsigned long long)smp >> 32);
    rng(res, 0xFFFFFFF0u, s_lo, s_hi, r4); normal2(r4, n0, n1); ax[0] = n0; ax[1] = n1;
    rng(res, 0xFFFFFFF1u, s_lo, s_hi, r4); normal2(r4, n0, n1); ax[2] = n0; zt[0] = n1;
    rng(res, 0xFFFFFFF2u, s_lo, s_hi, r4); normal2(r4, n0, n1); zt[1] = n0; zt[2] = n1;
    rng(res, 0xFFFFFFF3u, s_lo, s_hi, r4); u = u53(r4[0], r4[1]);
  }
  const double an = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
  // np.interp(u, cdf, omega)
  double ang;
  if (u <= cdf[0]) ang = omega[0];
  else if (u >= cdf[SO3_NOMEGA - 1]) ang = omega[SO3_NOMEGA - 1];
  else {
    int lo = 0, hi = SO3_NOMEGA - 1;   // invariant cdf[lo] <= u < cdf[hi]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] <= u) lo = mid; else hi = mid;
    }
    const double slope = (omega[lo + 1] - omega[lo]) / (cdf[lo + 1] - cdf[lo]);
    ang = slope * (u - cdf[lo]) + omega[lo];
  }
  const double sh = sin(0.5 * ang) / an;
  rigids[r * 7 + 0] = (float)cos(0.5 * ang);
  rigids[r * 7 + 1] = (float)(sh * ax[0]);
  rigids[r * 7 + 2] = (float)(sh * ax[1]);
  rigids[r * 7 + 3] = (float)(sh * ax[2]);
#pragma unroll
  for (int k = 0; k < 3; ++k) rigids[r * 7 + 4 + k] = (float)(zt[k] / 0.1);
}

// ----------------------------------------------------------------------------------------------------------------
// Forward noising of one example (SE3Diffuser.forward_marginal, data/se3_diffuser.py:43-110; SO3Diffuser.forward_marginal
// so3_diffuser.py:311-328; R3Diffuser.forward_marginal r3_diffuser.py:81-101).  One warp per residue.
//   rot:   axis = z_axis/|z_axis|, angle = interp(u, cdf_row(sigma_idx(t)), omega); s = axis*angle (fp64);
//          rot_score = igso3 score of s (computed on the fp64->fp32 rounded s like the reference's torch.tensor(vec) path
//          keeps fp64 — the reference calls torch_score on a float64 tensor here, so the whole series runs in fp64);
//          R_t = R_0 · exp(s)
//   trans: x_t ~ N(e^{-b/2}·0.1·x_0, 1 - e^{-b}) using z_trans; score = -(x_t - e^{-b/2} x_0s)/(1 - e^{-b}); x_t/0.1
// ----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double igso3_score_scalar_f64(double omega, double sigma, int lane) {
  const double lo = sin(omega / 2), dlo = 0.5 * cos(omega / 2);
  double psum = 0.0, dsum = 0.0;
  for (int l = lane; l < IGSO3_L; l += 32) {
    const double ex = -(double)((long long)l * (l + 1)) * (sigma * sigma) / 2.0;
    if (ex < -745.2) break;
    const double gauss = (double)(2 * l + 1) * exp(ex);
    double hi, c;
    sincos(omega * ((double)l + 0.5), &hi, &c);
    const double dhi = ((double)l + 0.5) * c;
    psum += gauss * hi / lo;
    dsum += gauss * (lo * dhi - hi * dlo) / (lo * lo);
  }
  psum = warp_sum_d(psum);
  dsum = warp_sum_d(dsum);
  return dsum / (psum + 1e-4);
}

__global__ void __launch_bounds__(256) forward_marginal_kernel(
    const float* __restrict__ rigids0, const double* __restrict__ z_axis, const double* __restrict__ u_angle,
    const double* __restrict__ z_trans, const float* __restrict__ diffuse_mask, double t, double sigma, const double* __restrict__ cdf,
    const double* __restrict__ omega_grid, float* __restrict__ rigids_t, double* __restrict__ rot_score, double* __restrict__ trans_score,
    long long n, const float* __restrict__ pad_mask = nullptr /* batched assembly: rows with pad_mask == 0 are padding -> all outputs zero */) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= n) return;
  if (pad_mask && pad_mask[r] == 0.f) {     // du.pad_feats (data/utils.py:387-399) pads every feature with zeros after the noising
    if (lane < 3) { rot_score[r * 3 + lane] = 0.0; trans_score[r * 3 + lane] = 0.0; }
    if (lane < 7) rigids_t[r * 7 + lane] = 0.f;
    return;
  }
  const double ax0 = z_axis[r * 3], ax1 = z_axis[r * 3 + 1], ax2 = z_axis[r * 3 + 2];
  const double an = sqrt(ax0 * ax0 + ax1 * ax1 + ax2 * ax2);
  const double u = u_angle[r];
  double ang;
  if (u <= cdf[0]) ang = omega_grid[0];
  else if (u >= cdf[SO3_NOMEGA - 1]) ang = omega_grid[SO3_NOMEGA - 1];
  else {
    int lo = 0, hi = SO3_NOMEGA - 1;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (cdf[mid] <= u) lo = mid; else hi = mid; }
    ang = (omega_grid[lo + 1] - omega_grid[lo]) / (cdf[lo + 1] - cdf[lo]) * (u - cdf[lo]) + omega_grid[lo];
  }
  const double s[3] = {ax0 / an * ang, ax1 / an * ang, ax2 / an * ang};
  // score of the sampled rotation vector (torch_score on a float64 tensor: omega = |s| + 1e-6, all fp64)
  const double om = sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]) + 1e-6;
  const double sc = igso3_score_scalar_f64(om, sigma, lane);
  const float dmf = diffuse_mask ? diffuse_mask[r] : 1.f;
  const double dm = (double)dmf;
  if (lane < 3) rot_score[r * 3 + lane] = dm * (sc * s[lane] / (om + 1e-6));
  if (lane == 0) {
    double q[4], n2 = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { q[k] = (double)rigids0[r * 7 + k]; n2 += q[k] * q[k]; }
    const double inv = 1.0 / sqrt(n2);
#pragma unroll
    for (int k = 0; k < 4; ++k) q[k] *= inv;
    double e[4];
    const double sh = ang < 1e-3 ? (0.5 - ang * ang / 48.0) : sin(ang * 0.5) / ang;
    e[0] = cos(ang * 0.5); e[1] = sh * s[0]; e[2] = sh * s[1]; e[3] = sh * s[2];
    double qn[4];
    quat_mul(q, e, qn);
    const double inv2 = 1.0 / sqrt(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
    const bool upd = dmf > 0.5f;
#pragma unroll
    for (int k = 0; k < 4; ++k) rigids_t[r * 7 + k] = (float)(upd ? qn[k] * inv2 : q[k]);
    const double beta = t * R3_MIN_B + 0.5 * (t * t) * (R3_MAX_B - R3_MIN_B);
    const double e1 = exp(-0.5 * beta), var = 1.0 - exp(-beta), sd = sqrt(var);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double x0 = (double)__fmul_rn(rigids0[r * 7 + 4 + k], COORD_SCALE);   // numpy keeps fp32 for x_0 * 0.1
      const double xt = e1 * x0 + sd * z_trans[r * 3 + k];
      const double ts = -(xt - e1 * x0) / var;
      trans_score[r * 3 + k] = dm * ts;
      rigids_t[r * 7 + 4 + k] = (float)(dm * (xt / 0.1) + (1.0 - dm) * (double)rigids0[r * 7 + 4 + k]);
    }
  }
}

// ----------------------------------------------------------------------------------------------------------------
// IGSO(3) cache rows (SO3Diffuser.__init__, data/so3_diffuser.py:151-180), all fp64 like the numpy branch.
// Kernel 1: block per (row, omega) computes the two L=1000 series.  Kernel 2: one thread per row does the
// sequential cumsum (numpy order) and the score-scaling reduction.
// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) igso3_series_kernel(const double* __restrict__ sigmas, double* __restrict__ expv,
                                                           double* __restrict__ dsig) {
  const int w = blockIdx.x;      // omega index
  const int r = blockIdx.y;      // row
  const double omega = (double)(w + 1) * (3.14159265358979323846 / SO3_NOMEGA);
  const double sg = sigmas[r];
  const double lo = sin(omega / 2), dlo = 0.5 * cos(omega / 2);
  double p = 0.0, d = 0.0;
  for (int l = threadIdx.x; l < IGSO3_L; l += 128) {
    const double gauss = (double)(2 * l + 1) * exp(-(double)((long long)l * (l + 1)) * (sg * sg) / 2);
    const double arg = omega * ((double)l + 0.5);
    double hi, c;
    sincos(arg, &hi, &c);
    const double dhi = ((double)l + 0.5) * c;
    p += gauss * hi / lo;
    d += gauss * (lo * dhi - hi * dlo) / (lo * lo);
  }
  __shared__ double sp[4], sd[4];
  p = warp_sum_d(p); d = warp_sum_d(d);
  if ((threadIdx.x & 31) == 0) { sp[threadIdx.x >> 5] = p; sd[threadIdx.x >> 5] = d; }
  __syncthreads();
  if (threadIdx.x == 0) {
    expv[(long long)r * SO3_NOMEGA + w] = (sp[0] + sp[1]) + (sp[2] + sp[3]);
    dsig[(long long)r * SO3_NOMEGA + w] = (sd[0] + sd[1]) + (sd[2] + sd[3]);
  }
}

__global__ void igso3_rows_post_kernel(const double* __restrict__ expv, const double* __restrict__ dsig,
                                       double* __restrict__ pdf, double* __restrict__ cdf, double* __restrict__ snorm,
                                       double* __restrict__ scaling, int nrows) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  double run = 0.0, num = 0.0, den = 0.0;
  for (int w = 0; w < SO3_NOMEGA; ++w) {
    const double omega = (double)(w + 1) * (3.14159265358979323846 / SO3_NOMEGA);
    const double e = expv[(long long)r * SO3_NOMEGA + w];
    const double pd = e * (1.0 - cos(omega)) / 3.14159265358979323846;
    const double sn = dsig[(long long)r * SO3_NOMEGA + w] / (e + 1e-4);
    run += pd;
    pdf[(long long)r * SO3_NOMEGA + w] = pd;
    cdf[(long long)r * SO3_NOMEGA + w] = run / SO3_NOMEGA * 3.14159265358979323846;
    snorm[(long long)r * SO3_NOMEGA + w] = sn;
    num += sn * sn * pd;
    den += pd;
  }
  scaling[r] = sqrt(fabs(num / den)) / sqrt(3.0);
}


// ----------------------------------------------------------------------------------------------------------------
// DSM training loss, forward values (Experiment.loss_fn, experiments/train_se3_diffusion.py:538-660); kernels: loss_fwd2_kernel +
// loss_finalize_kernel in fd_train.cuh.
//   terms[b] = { rot_loss, trans_loss, bb_atom_loss, dist_mat_loss, their sum }   (the per-sample `batch_*` entries of aux_data)
// Scores and their targets are fp64 (the reference's score tensors are), frames / atoms fp32 with fp64 accumulation; the ground-truth
// backbone comes from rigids_0 and the psi torsion exactly as all_atom.compute_backbone builds it.
// ----------------------------------------------------------------------------------------------------------------
struct LossArgs {
  const double* pred_rot; const double* pred_trans; const float* pred_rigids; const float* pred_atom37;
  const double* gt_rot; const double* gt_trans; const double* rot_scaling; const double* trans_scaling; const double* rigids_0;
  const double* t; const float* res_mask; const float* fixed_mask; const float* gt_psi;
  double trans_loss_weight, rot_loss_weight, rot_loss_t_threshold, trans_x0_threshold, coordinate_scaling, bb_atom_loss_weight,
      bb_atom_loss_t_filter, dist_mat_loss_weight, dist_mat_loss_t_filter, aux_loss_weight;
  int separate_rot_loss, diffuse_trans, diffuse_rot;
  double* terms;   // [B,5]
  int N;
};

__device__ __forceinline__ double block_sum_f64(double v, double* red) {   // 256 threads; red[8]
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < 8; ++w) s += red[w];
  return s;
}


}  // namespace fd
