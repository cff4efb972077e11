// Training-step kernels: the hand-written backward of ScoreNetwork.forward + Experiment.loss_fn, Adam.
//
// The reference has no backward code — Experiment.update_fn (experiments/train_se3_diffusion.py:320-326) calls loss.backward()
// and torch autograd differentiates loss_fn (:524-693), model/score_network.py:170-215 and model/ipa_pytorch.py:194-672.
// Every kernel here implements one piece of that derivative; the decomposition (tape contents, grouping of GEMMs, per-residue
// formulas) is the one restated and checked against autograd on the CPU in oracle/manual_backward.py, function by function.
// Semantics are those torch applies whenever autograd records: the float key-padding mask is ADDED to the sequence-attention
// logits.
#pragma once
#include "fd_common.cuh"

namespace fd {

// ---------------------------------------------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------------------------------------------
__global__ void zero_f32_kernel(float* __restrict__ p, long long n) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) *reinterpret_cast<float4*>(p + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  else for (long long k = i; k < n; ++k) p[k] = 0.f;
}
// dst[m][0:C] (ld ldd) (+)= src[m][0:C] (ld lds) * (rowmask ? rowmask[m] : 1)
__global__ void copy_cols_kernel(float* __restrict__ dst, int ldd, const float* __restrict__ src, int lds, int C, long long M,
                                 const float* __restrict__ rowmask, int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * C) return;
  const long long m = i / C; const int c = (int)(i - m * C);
  const float v = src[m * lds + c] * (rowmask ? rowmask[m] : 1.f);
  float* d = dst + m * ldd + c;
  *d = accumulate ? *d + v : v;
}

// out[n] += sum_m X[m][n]   (bias gradients).  grid (ceil(N/128), slabs), 128 threads, one column per thread.
__global__ void __launch_bounds__(128) colsum_kernel(const float* __restrict__ X, int ld, long long M, int N, float* __restrict__ out) {
  const int n = blockIdx.x * 128 + threadIdx.x;
  if (n >= N) return;
  const long long per = (M + gridDim.y - 1) / gridDim.y;
  const long long m0 = (long long)blockIdx.y * per, m1 = min(M, m0 + per);
  float s0 = 0.f, s1 = 0.f;
  long long m = m0;
  for (; m + 1 < m1; m += 2) { s0 += X[m * ld + n]; s1 += X[(m + 1) * ld + n]; }
  if (m < m1) s0 += X[m * ld + n];
  atomicAdd(out + n, s0 + s1);
}
// narrow tensors ([M, N] dense with N | 128, e.g. the [E, 8] pair-bias gradient): viewed as [M*N/128, 128] so that all 128 threads of a block
// work, the 128 partial columns folded onto the N outputs by the last stage
__global__ void colsum_fold_kernel(const float* __restrict__ tmp128, int N, float* __restrict__ out) {
  const int n = threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int k = n; k < 128; k += N) s += tmp128[k];
  atomicAdd(out + n, s);
}
inline cudaError_t launch_colsum(const float* X, int ld, long long M, int N, float* out, cudaStream_t st);
inline cudaError_t launch_colsum_narrow(const float* X, long long M, int N, float* out, float* tmp128, cudaStream_t st) {
  cudaMemsetAsync(tmp128, 0, 128 * sizeof(float), st);
  cudaError_t e = launch_colsum(X, 128, M * N / 128, 128, tmp128, st);
  if (e != cudaSuccess) return e;
  colsum_fold_kernel<<<1, 128, 0, st>>>(tmp128, N, out);
  return cudaGetLastError();
}
inline cudaError_t launch_colsum(const float* X, int ld, long long M, int N, float* out, cudaStream_t st) {
  // >= 32 rows per thread; a node-level tensor (a few thousand rows) still spreads over tens of CTAs instead of eight
  int slabs = (int)min((long long)1024, max((long long)1, M / 32));
  colsum_kernel<<<dim3((N + 127) / 128, slabs), 128, 0, st>>>(X, ld, M, N, out);
  return cudaGetLastError();
}

// Sums over one of the two residue axes of an edge tensor X [B,N,N,C]:  out[b][a][c] = sum_k X[b][.., ..][c]
//   mode 0 (row sums):    a = i, k = j   (stride_a = N*C, stride_k = C)
//   mode 1 (column sums): a = j, k = i   (stride_a = C,   stride_k = N*C)
// grid (B*N), block 128: each thread owns channels c, c+128, ... ; optional edge mask is NOT applied (callers pass masked gradients).
__global__ void __launch_bounds__(128) edge_axis_sum_kernel(const float* __restrict__ X, float* __restrict__ out, int N, int C, int mode) {
  const long long ba = blockIdx.x;
  const long long b = ba / N; const int a = (int)(ba - b * N);
  const long long sa = mode == 0 ? (long long)N * C : C, sk = mode == 0 ? C : (long long)N * C;
  const float* base = X + b * N * N * C + a * sa;
  for (int c = threadIdx.x; c < C; c += 128) {
    float s0 = 0.f, s1 = 0.f;
    int k = 0;
    for (; k + 1 < N; k += 2) { s0 += base[k * sk + c]; s1 += base[(k + 1) * sk + c]; }
    if (k < N) s0 += base[k * sk + c];
    out[ba * C + c] = s0 + s1;
  }
}

// fp32 weight block -> bf16 hi/lo planes for the tcgen05 GEMMs of the training path (rebuilt every step: the weights change).
//   dst[r][dst_col0 + c] = split(transpose ? src[c*ld + r] : src[r*ld + c]),  r < rows, c < cols;  dst row stride dst_ld
__global__ void pack_planes_kernel(const float* __restrict__ src, int ld, int rows, int cols, int transpose, __nv_bfloat16* __restrict__ hi,
                                   __nv_bfloat16* __restrict__ lo, int dst_ld, int dst_col0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int r = i / cols, c = i - r * cols;
  const float v = transpose ? src[(long long)c * ld + r] : src[(long long)r * ld + c];
  __nv_bfloat16 h, l;
  split_bf16(v, h, l);
  hi[(long long)r * dst_ld + dst_col0 + c] = h;
  lo[(long long)r * dst_ld + dst_col0 + c] = l;
}

// Sums over one residue axis of an edge tensor given as bf16 hi/lo planes [B,N,N,C] (value = hi + lo); same modes as edge_axis_sum_kernel.
__global__ void __launch_bounds__(128) edge_axis_sum_planes_kernel(const __nv_bfloat16* __restrict__ Xh, const __nv_bfloat16* __restrict__ Xl,
                                                                   float* __restrict__ out, int N, int C, int mode) {
  const long long ba = blockIdx.x;
  const long long b = ba / N; const int a = (int)(ba - b * N);
  const long long sa = mode == 0 ? (long long)N * C : C, sk = mode == 0 ? C : (long long)N * C;
  const long long base = b * N * N * C + a * sa;
  for (int c = threadIdx.x * 2; c < C; c += 256) {          // two channels per thread (one 32-bit load per plane)
    float s0 = 0.f, s1 = 0.f;
    for (int k = 0; k < N; ++k) {
      const uint32_t h = *reinterpret_cast<const uint32_t*>(Xh + base + k * sk + c), l = *reinterpret_cast<const uint32_t*>(Xl + base + k * sk + c);
      s0 += __uint_as_float(h << 16) + __uint_as_float(l << 16);
      s1 += __uint_as_float(h & 0xffff0000u) + __uint_as_float(l & 0xffff0000u);
    }
    out[ba * C + c] = s0; out[ba * C + c + 1] = s1;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// LayerNorm backward (torch.nn.LayerNorm, eps 1e-5).  One warp per row, persistent grid; dgamma/dbeta accumulated in registers,
// reduced per block in shared memory, one atomicAdd per channel per block.
//   g = dy * mask ;  dxh = g*gamma ;  dx = rstd*(dxh - mean(dxh) - xh*mean(dxh*xh)) ;  dgamma += g*xh ; dbeta += g
// mask: rowmask[m] or res_mask[b,i]*res_mask[b,j] for edge rows (the forward applies the mask AFTER the norm).
// ---------------------------------------------------------------------------------------------------------------------
struct LnBwdArgs {
  const float* x = nullptr; int ldx = 0;          // LayerNorm INPUT (pre-norm), statistics are recomputed
  const float* dy = nullptr; int lddy = 0;
  const float* gamma = nullptr;
  float* dx = nullptr; int lddx = 0; int accumulate = 0;
  float* dgamma = nullptr; float* dbeta = nullptr;
  long long M = 0;
  const float* rowmask = nullptr;
  const float* res_mask = nullptr; int nres = 0;
};
template <int C>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const LnBwdArgs a) {
  constexpr int V = C / 4;
  constexpr int PER = (V + 31) / 32;
  __shared__ float red[2][8][C];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 dg[PER], db[PER];
#pragma unroll
  for (int p = 0; p < PER; ++p) { dg[p] = make_float4(0.f, 0.f, 0.f, 0.f); db[p] = dg[p]; }
  const float4* g4 = reinterpret_cast<const float4*>(a.gamma);
  for (long long m = (long long)blockIdx.x * 8 + warp; m < a.M; m += (long long)gridDim.x * 8) {
    const float4* xr = reinterpret_cast<const float4*>(a.x + m * a.ldx);
    const float4* dr = reinterpret_cast<const float4*>(a.dy + m * a.lddy);
    float mk = 1.f;
    if (a.rowmask) mk = a.rowmask[m];
    if (a.res_mask) {
      const long long nn = (long long)a.nres * a.nres;
      const long long b = m / nn;
      const int rem = (int)(m - b * nn);
      const int i = rem / a.nres, j = rem - i * a.nres;
      mk = a.res_mask[b * a.nres + i] * a.res_mask[b * a.nres + j];
    }
    float4 v[PER], g[PER];
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < PER; ++p) {
      const int idx = lane + p * 32;
      v[p] = idx < V ? xr[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
      g[p] = idx < V ? dr[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
      g[p].x *= mk; g[p].y *= mk; g[p].z *= mk; g[p].w *= mk;
      s += (v[p].x + v[p].y) + (v[p].z + v[p].w);
    }
    const float mean = warp_sum(s) * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int p = 0; p < PER; ++p) {
      const int idx = lane + p * 32;
      if (idx < V) {
        v[p].x -= mean; v[p].y -= mean; v[p].z -= mean; v[p].w -= mean;
        q += (v[p].x * v[p].x + v[p].y * v[p].y) + (v[p].z * v[p].z + v[p].w * v[p].w);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + 1e-5f);
    float s1 = 0.f, s2 = 0.f;
    float4 dxh[PER];
#pragma unroll
    for (int p = 0; p < PER; ++p) {
      const int idx = lane + p * 32;
      dxh[p] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < V) {
        const float4 gm = g4[idx];
        v[p].x *= rstd; v[p].y *= rstd; v[p].z *= rstd; v[p].w *= rstd;       // xh
        dxh[p].x = g[p].x * gm.x; dxh[p].y = g[p].y * gm.y; dxh[p].z = g[p].z * gm.z; dxh[p].w = g[p].w * gm.w;
        s1 += (dxh[p].x + dxh[p].y) + (dxh[p].z + dxh[p].w);
        s2 += (dxh[p].x * v[p].x + dxh[p].y * v[p].y) + (dxh[p].z * v[p].z + dxh[p].w * v[p].w);
        dg[p].x += g[p].x * v[p].x; dg[p].y += g[p].y * v[p].y; dg[p].z += g[p].z * v[p].z; dg[p].w += g[p].w * v[p].w;
        db[p].x += g[p].x; db[p].y += g[p].y; db[p].z += g[p].z; db[p].w += g[p].w;
      }
    }
    s1 = warp_sum(s1) * (1.f / C);
    s2 = warp_sum(s2) * (1.f / C);
    float4* o = reinterpret_cast<float4*>(a.dx + m * a.lddx);
#pragma unroll
    for (int p = 0; p < PER; ++p) {
      const int idx = lane + p * 32;
      if (idx < V) {
        float4 r;
        r.x = rstd * (dxh[p].x - s1 - v[p].x * s2); r.y = rstd * (dxh[p].y - s1 - v[p].y * s2);
        r.z = rstd * (dxh[p].z - s1 - v[p].z * s2); r.w = rstd * (dxh[p].w - s1 - v[p].w * s2);
        if (a.accumulate) { const float4 e = o[idx]; r.x += e.x; r.y += e.y; r.z += e.z; r.w += e.w; }
        o[idx] = r;
      }
    }
  }
#pragma unroll
  for (int p = 0; p < PER; ++p) {
    const int idx = lane + p * 32;
    if (idx < V) {
      *reinterpret_cast<float4*>(&red[0][warp][idx * 4]) = dg[p];
      *reinterpret_cast<float4*>(&red[1][warp][idx * 4]) = db[p];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += 256) {
    const int which = c / C, cc = c - which * C;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[which][w][cc];
    atomicAdd((which ? a.dbeta : a.dgamma) + cc, s);
  }
}
inline cudaError_t launch_ln_bwd(int C, const LnBwdArgs& a, int sm_count, cudaStream_t st) {
  const long long want = (a.M + 7) / 8;
  const unsigned grid = (unsigned)min(want, (long long)sm_count * 4);
  if (C == 128) ln_bwd_kernel<128><<<grid, 256, 0, st>>>(a);
  else if (C == 256) ln_bwd_kernel<256><<<grid, 256, 0, st>>>(a);
  else if (C == 320) ln_bwd_kernel<320><<<grid, 256, 0, st>>>(a);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------------
// Sequence transformer attention, autograd semantics (model/ipa_pytorch.py:636: src_key_padding_mask = 1 - mask passed as FLOAT):
// softmax_j( S[r][j] + (1 - mask[b][j]) ) in place over S [rows, ld]; columns [n, ld) zeroed.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_rows_addmask_kernel(float* S, int ld, int n, long long rows, long long rows_per_sample,
                                                                   const float* keymask) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  float* row = S + r * ld;
  const float* km = keymask + (r / rows_per_sample) * n;
  float mx = -INFINITY;
  for (int j = lane; j < n; j += 32) { const float v = row[j] + (1.f - km[j]); row[j] = v; mx = fmaxf(mx, v); }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < n; j += 32) { const float e = expf(row[j] - mx); row[j] = e; sum += e; }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  for (int j = lane; j < ld; j += 32) row[j] = j < n ? row[j] * inv : 0.f;
}
// dS = P * (dP - sum_j P dP) * scale, in place over dP
__global__ void __launch_bounds__(256) softmax_rows_bwd_kernel(const float* __restrict__ P, float* __restrict__ dP, int ld, int n, long long rows,
                                                               float scale) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* p = P + r * ld;
  float* d = dP + r * ld;
  float s = 0.f;
  for (int j = lane; j < n; j += 32) s += p[j] * d[j];
  s = warp_sum(s);
  for (int j = lane; j < ld; j += 32) d[j] = j < n ? p[j] * (d[j] - s) * scale : 0.f;
}

// ---------------------------------------------------------------------------------------------------------------------
// Edge-embedder input features for the training path (model/score_network.py:97-100,137-149): pair [E,120] =
// [t-emb | fixed]_i (33) | [t-emb | fixed]_j (33) | idx-emb(seq_i - seq_j) (32) | distogram(sc_ca) one-hot (22).  One warp per edge.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pair_feats_kernel(const float* __restrict__ temb, const float* __restrict__ fixed_mask,
                                                         const int* __restrict__ seq_idx, const float* __restrict__ sc_ca,
                                                         float* __restrict__ pair, int N) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (j >= N) return;
  const long long ri = blockIdx.x;
  const long long b = ri / N;
  const long long rj = b * N + j;
  float* o = pair + (ri * N + j) * EDGE_IN;
  const float te = temb[b * 32 + lane];
  o[lane] = te; o[33 + lane] = te;
  if (lane == 0) { o[32] = fixed_mask[ri]; o[65] = fixed_mask[rj]; }
  const int d = seq_idx[ri] - seq_idx[rj];
  const int k = lane & 15;
  const float ia = __fdiv_rn(__fmul_rn((float)d, c_pi_f32), c_idx_den[k]);
  o[66 + lane] = lane < 16 ? sinf(ia) : cosf(ia);
  const float dx = sc_ca[ri * 3 + 0] - sc_ca[rj * 3 + 0];
  const float dy = sc_ca[ri * 3 + 1] - sc_ca[rj * 3 + 1];
  const float dz = sc_ca[ri * 3 + 2] - sc_ca[rj * 3 + 2];
  const int bin = dgram_bin(sqrtf(dx * dx + dy * dy + dz * dz));
  if (lane < NBINS) o[98 + lane] = lane == bin ? 1.f : 0.f;
}

// gamma_h = softplus(head_weights_h) * sqrt(1/(3*PQ*9/2)), WdT = transpose(down_z.weight) — the small derived tensors the IPA edge
// kernel takes (rebuilt every step: the weights change)
__global__ void ipa_derived_kernel(const float* __restrict__ hw, const float* __restrict__ Wd /* [32][128] */, float* __restrict__ gamma,
                                   float* __restrict__ WdT /* [128][32] */) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < H) {
    const float x = hw[t];
    const float sp = x > 20.f ? x : log1pf(expf(x));
    gamma[t] = sp * 0.09622504486493763f;   // sqrt(1/108)
  }
  if (t < 32 * C_Z) { const int d = t / C_Z, c = t - d * C_Z; WdT[c * 32 + d] = Wd[t]; }
}
// d head_weights_h = dgamma_h * sigmoid(hw_h) * sqrt(1/108)
__global__ void ipa_gamma_bwd_kernel(const float* __restrict__ hw, const float* __restrict__ dgamma, float* __restrict__ dhw) {
  const int t = threadIdx.x;
  if (t < H) dhw[t] += dgamma[t] * (1.f / (1.f + expf(-hw[t]))) * 0.09622504486493763f;
}

// d quat (+)= d/dq [ R(q) : G ] for the unnormalised polynomial quat -> rot map (rigid_utils.py:185)
__device__ __forceinline__ void quat_grad_from_rot_grad(const float q[4], const float G[9], float dq[4]) {
  const float a = q[0], b = q[1], c = q[2], d = q[3];
  dq[0] = 2.f * (a * (G[0] + G[4] + G[8]) + d * (G[3] - G[1]) + c * (G[2] - G[6]) + b * (G[7] - G[5]));
  dq[1] = 2.f * (b * (G[0] - G[4] - G[8]) + c * (G[1] + G[3]) + d * (G[2] + G[6]) + a * (G[7] - G[5]));
  dq[2] = 2.f * (c * (-G[0] + G[4] - G[8]) + b * (G[1] + G[3]) + a * (G[2] - G[6]) + d * (G[5] + G[7]));
  dq[3] = 2.f * (d * (-G[0] - G[4] + G[8]) + a * (G[3] - G[1]) + b * (G[2] + G[6]) + c * (G[5] + G[7]));
}

// block-wide sum of NV values per thread (blockDim <= 256), result valid in thread 0
template <int NV>
__device__ __forceinline__ void block_sum_vec(float v[NV], float* red /* [8][NV] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) red[warp * NV + k] = v[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) { float s = 0.f; for (int w = 0; w < nw; ++w) s += red[w * NV + k]; v[k] = s; }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// IPA, o_pt part backward (forward: ipa_finish_kernel; model/ipa_pytorch.py:437-447).  One CTA (96 threads) per residue.
//   in : dfeats [R,2688] (x | y | z | norm columns), feats (saved local points + norms), optg (saved global sums), frames
//   out: doptg [R, H*PV*3] = R * dl ;  dquat += dR(gm (x) dl) ;  dtrans -= sum doptg
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(96) ipa_finish_bwd_kernel(const float* __restrict__ dfeats, const float* __restrict__ feats,
                                                            const float* __restrict__ optg, const float* __restrict__ quat,
                                                            const float* __restrict__ trans, float* __restrict__ doptg,
                                                            float* __restrict__ dquat, float* __restrict__ dtrans) {
  __shared__ float red[8 * 12];
  const long long row = blockIdx.x;
  const int hp = threadIdx.x;
  float q[4], R[9];
#pragma unroll
  for (int k = 0; k < 4; ++k) q[k] = quat[row * 4 + k];
  quat_to_rot(q, R);
  const float* f = feats + row * IPA_FEAT + H * C_HID;
  const float* df = dfeats + row * IPA_FEAT + H * C_HID;
  const float lx = f[hp], ly = f[H * PV + hp], lz = f[2 * H * PV + hp], nr = f[3 * H * PV + hp];
  const float dn = df[3 * H * PV + hp] / nr;
  const float dl[3] = {df[hp] + dn * lx, df[H * PV + hp] + dn * ly, df[2 * H * PV + hp] + dn * lz};
  const float* g = optg + row * (H * PV * 3) + hp * 3;
  const float gm[3] = {g[0] - trans[row * 3 + 0], g[1] - trans[row * 3 + 1], g[2] - trans[row * 3 + 2]};
  float dg[3];
#pragma unroll
  for (int m = 0; m < 3; ++m) dg[m] = R[m * 3 + 0] * dl[0] + R[m * 3 + 1] * dl[1] + R[m * 3 + 2] * dl[2];
  float* o = doptg + row * (H * PV * 3) + hp * 3;
  o[0] = dg[0]; o[1] = dg[1]; o[2] = dg[2];
  float v[12];
#pragma unroll
  for (int m = 0; m < 3; ++m)
#pragma unroll
    for (int k = 0; k < 3; ++k) v[m * 3 + k] = gm[m] * dl[k];
  v[9] = -dg[0]; v[10] = -dg[1]; v[11] = -dg[2];
  block_sum_vec<12>(v, red);
  if (threadIdx.x == 0) {
    float dq[4];
    quat_grad_from_rot_grad(q, v, dq);
#pragma unroll
    for (int k = 0; k < 4; ++k) dquat[row * 4 + k] += dq[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) dtrans[row * 3 + k] += v[9 + k];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// IPA edge pass backward for one query residue (b,i), all 8 heads (forward: ipa_edge_kernel).
//   in : A  [B,H,N,Np]  attention probabilities (tape)
//        dA [B,H,N,Np]  partial d/dA from the value aggregations (do.v^T + doptg.vp^T, two batched GEMMs upstream)
//        dzbar [R,H,128] = dopair . Wd  (o_pair = Wd.zbar + bd with zbar = sum_j a z)
//   does: dA += dzbar[h] . z_ij ;  dL = A*(dA - sum_j A dA)  -> written over dA ;  dbias[b,i,j,h] = sqrt(1/3) dL ;
//         dz[b,i,j,:] (+)= sum_h ( sqrt(1/3) dL[h][j] Wb[h,:] + A[h][j] dzbar[h,:] )
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ipa_edge_bwd_kernel(const float* __restrict__ z, const float* __restrict__ A, float* __restrict__ dA,
                                                           const float* __restrict__ dzbar, const float* __restrict__ Wb,
                                                           float* __restrict__ dbias, float* __restrict__ dz, int accumulate_dz, int N, int Np) {
  extern __shared__ __align__(16) float sm[];
  float* as = sm;                 // [H][Np]
  float* ds = as + H * Np;        // [H][Np]
  const int i = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long rowi = (long long)b * N + i;
  for (int idx = tid; idx < H * Np; idx += 256) {
    const int h = idx / Np, j = idx - h * Np;
    const long long off = (((long long)b * H + h) * N + i) * Np + j;
    as[idx] = j < N ? A[off] : 0.f;
    ds[idx] = j < N ? dA[off] : 0.f;
  }
  float wb[H][4], zb[H][4];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    const float4 w = reinterpret_cast<const float4*>(Wb + h * C_Z)[lane];
    wb[h][0] = w.x; wb[h][1] = w.y; wb[h][2] = w.z; wb[h][3] = w.w;
    const float4 d = reinterpret_cast<const float4*>(dzbar + (rowi * H + h) * C_Z)[lane];
    zb[h][0] = d.x; zb[h][1] = d.y; zb[h][2] = d.z; zb[h][3] = d.w;
  }
  __syncthreads();
  const long long zrow = rowi * N * C_Z;
  const int myh = lane >> 2;
  for (int j = warp; j < N; j += 8) {
    const float4 zv = *reinterpret_cast<const float4*>(z + zrow + (long long)j * C_Z + lane * 4);
    float pb[H];
#pragma unroll
    for (int h = 0; h < H; ++h) pb[h] = zb[h][0] * zv.x + zb[h][1] * zv.y + zb[h][2] * zv.z + zb[h][3] * zv.w;
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
    if ((lane & 3) == 0) ds[myh * Np + j] += v1;
  }
  __syncthreads();
  {  // softmax backward: warp h <-> head h
    const float* ar = as + warp * Np;
    float* dr = ds + warp * Np;
    float s = 0.f;
    for (int j = lane; j < N; j += 32) s += ar[j] * dr[j];
    s = warp_sum(s);
    float* out = dA + (((long long)b * H + warp) * N + i) * Np;
    for (int j = lane; j < Np; j += 32) {
      const float v = j < N ? ar[j] * (dr[j] - s) : 0.f;
      dr[j] = v;
      out[j] = v;
    }
  }
  __syncthreads();
  const float k13 = 0.57735026918962576f;
  for (int j = warp; j < N; j += 8) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float dl = k13 * ds[h * Np + j], aa = as[h * Np + j];
      o.x += dl * wb[h][0] + aa * zb[h][0]; o.y += dl * wb[h][1] + aa * zb[h][1];
      o.z += dl * wb[h][2] + aa * zb[h][2]; o.w += dl * wb[h][3] + aa * zb[h][3];
    }
    float4* dst = reinterpret_cast<float4*>(dz + zrow + (long long)j * C_Z + lane * 4);
    if (accumulate_dz) { const float4 e = *dst; o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w; }
    *dst = o;
    if (lane < H) dbias[(rowi * N + j) * H + lane] = k13 * ds[lane * Np + j];
  }
}

// column sums of dL over the query index: cs[b][h][j] = sum_i dL[b][h][i][j].  grid (B*H), block 256
__global__ void __launch_bounds__(256) attn_colsum_kernel(const float* __restrict__ dL, float* __restrict__ cs, int N, int Np) {
  const float* base = dL + (long long)blockIdx.x * N * Np;
  for (int j = threadIdx.x; j < N; j += 256) {
    float s = 0.f;
    for (int i = 0; i < N; ++i) s += base[(long long)i * Np + j];
    cs[(long long)blockIdx.x * N + j] = s;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// IPA points backward (forward: ipa_points_kernel).  One CTA (224 threads) per residue; thread p = one point.
//   q-points:  dP = gamma_h * Gq                         (Gq = sum_j dL kp_j ; the sum_j dL = 0 term drops out)
//   k-points:  dP = gamma_h * (Gk - kp * colsum_j)       (Gk = sum_i dL qp_i)
//   v-points:  dP = dvp
//   d raw (local, [x-block | y-block | z-block]) = R^T dP ; dR += dP (x) p_local ; dt += dP ; dgamma_h -= 0.5*(colsum |kp|^2 - 2 qp.Gq)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(224) ipa_points_bwd_kernel(const float* __restrict__ proj /* [R,6816] raw projections (tape) */,
                                                             const float* __restrict__ quat, const float* __restrict__ qp,
                                                             const float* __restrict__ kp, const float* __restrict__ Gq,
                                                             const float* __restrict__ Gk, const float* __restrict__ dvp,
                                                             const float* __restrict__ colsum /* [B,H,N] */, const float* __restrict__ gamma,
                                                             float* __restrict__ dproj /* [R,6816]: columns of the point projections */,
                                                             float* __restrict__ dquat, float* __restrict__ dtrans,
                                                             float* __restrict__ dgamma /* [H] */, int N) {
  __shared__ float red[8 * 12];
  __shared__ float gred[H];
  const long long row = blockIdx.x;
  const long long b = row / N; const int n = (int)(row - b * N);
  const int p = threadIdx.x;
  if (p < H) gred[p] = 0.f;
  float q[4], R[9];
#pragma unroll
  for (int k = 0; k < 4; ++k) q[k] = quat[row * 4 + k];
  quat_to_rot(q, R);
  float dP[3], loc[3];
  float* dst; int blk, stride;
  float dgam = 0.f; int hh;
  const float* pr = proj + row * PROJ_ALL;
  if (p < H * PQ) {
    hh = p / PQ;
    const float gm = gamma[hh];
    const float* gq = Gq + row * (H * PQ * 3) + p * 3;
    const float* qq = qp + row * (H * PQ * 3) + p * 3;
    dP[0] = gm * gq[0]; dP[1] = gm * gq[1]; dP[2] = gm * gq[2];
    dgam = qq[0] * gq[0] + qq[1] * gq[1] + qq[2] * gq[2];        // -0.5 * (-2 qp.Gq)
    const float* s = pr + PROJ_Q + PROJ_KV;
    loc[0] = s[p]; loc[1] = s[H * PQ + p]; loc[2] = s[2 * H * PQ + p];
    dst = dproj + row * PROJ_ALL + PROJ_Q + PROJ_KV; blk = p; stride = H * PQ;
  } else {
    const int pp = p - H * PQ;
    hh = pp / (PQ + PV);
    const int p2 = pp - hh * (PQ + PV);
    const float* s = pr + PROJ_Q + PROJ_KV + PROJ_QP;
    constexpr int NB = H * (PQ + PV);
    loc[0] = s[pp]; loc[1] = s[NB + pp]; loc[2] = s[2 * NB + pp];
    if (p2 < PQ) {
      const float gm = gamma[hh];
      const float cs = colsum[(b * H + hh) * N + n];
      const float* gk = Gk + row * (H * PQ * 3) + (hh * PQ + p2) * 3;
      const float* kk = kp + row * (H * PQ * 3) + (hh * PQ + p2) * 3;
      dP[0] = gm * (gk[0] - kk[0] * cs); dP[1] = gm * (gk[1] - kk[1] * cs); dP[2] = gm * (gk[2] - kk[2] * cs);
      dgam = -0.5f * cs * (kk[0] * kk[0] + kk[1] * kk[1] + kk[2] * kk[2]);
    } else {
      const float* dv = dvp + row * (H * PV * 3) + (hh * PV + (p2 - PQ)) * 3;
      dP[0] = dv[0]; dP[1] = dv[1]; dP[2] = dv[2];
    }
    dst = dproj + row * PROJ_ALL + PROJ_Q + PROJ_KV + PROJ_QP; blk = pp; stride = NB;
  }
  // local gradient R^T dP, written in block layout
  dst[blk] = R[0] * dP[0] + R[3] * dP[1] + R[6] * dP[2];
  dst[stride + blk] = R[1] * dP[0] + R[4] * dP[1] + R[7] * dP[2];
  dst[2 * stride + blk] = R[2] * dP[0] + R[5] * dP[1] + R[8] * dP[2];
  float v[12];
#pragma unroll
  for (int m = 0; m < 3; ++m)
#pragma unroll
    for (int k = 0; k < 3; ++k) v[m * 3 + k] = dP[m] * loc[k];
  v[9] = dP[0]; v[10] = dP[1]; v[11] = dP[2];
  __syncthreads();
  if (dgam != 0.f) atomicAdd(&gred[hh], dgam);
  block_sum_vec<12>(v, red);          // contains a __syncthreads: gred complete afterwards
  if (threadIdx.x == 0) {
    float dq[4];
    quat_grad_from_rot_grad(q, v, dq);
#pragma unroll
    for (int k = 0; k < 4; ++k) dquat[row * 4 + k] += dq[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) dtrans[row * 3 + k] += v[9 + k];
  }
  if (p < H) atomicAdd(dgamma + p, gred[p]);
}

// ---------------------------------------------------------------------------------------------------------------------
// BackboneUpdate + compose_q_update_vec backward (forward: backbone_update_kernel; rigid_utils.py:587-616,1039-1063).
// One warp per residue.  quat_in/trans: the frames BEFORE the update (tape); node: the block's output node features (tape).
// In/out: dquat, dtrans hold d/d(new frames) on entry and d/d(old frames) on exit; dnode += ; dWbb, dbbb accumulate (atomics).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) backbone_update_bwd_kernel(const float* __restrict__ node, const float* __restrict__ Wbb,
                                                                  const float* __restrict__ bbb, const float* __restrict__ res_mask,
                                                                  const float* __restrict__ fixed_mask, const float* __restrict__ quat_in,
                                                                  float* __restrict__ dquat, float* __restrict__ dtrans,
                                                                  float* __restrict__ dnode, float* __restrict__ dW /* [6][256] */,
                                                                  float* __restrict__ dB /* [6] */, long long rows) {
  __shared__ float wred[8][6];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float du[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float dwacc[6][8];
#pragma unroll
  for (int o = 0; o < 6; ++o)
#pragma unroll
    for (int k = 0; k < 8; ++k) dwacc[o][k] = 0.f;
  float dbacc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long row = (long long)blockIdx.x * 8 + warp; row < rows; row += (long long)gridDim.x * 8) {
    const float dm = (1.f - fixed_mask[row]) * res_mask[row];
    float u[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float* x = node + row * C_S;
    float xv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      xv[k] = x[lane + k * 32] * dm;
#pragma unroll
      for (int o = 0; o < 6; ++o) u[o] = fmaf(Wbb[o * C_S + lane + k * 32], xv[k], u[o]);
    }
#pragma unroll
    for (int o = 0; o < 6; ++o) u[o] = warp_sum(u[o]) + bbb[o];
    // every lane computes the (tiny) per-residue algebra redundantly
    float q[4], R[9];
#pragma unroll
    for (int k = 0; k < 4; ++k) q[k] = quat_in[row * 4 + k];
    quat_to_rot(q, R);
    const float v4[4] = {0.f, u[0], u[1], u[2]};
    float dq4[4];
    quat_mul(q, v4, dq4);
    float nq[4], n2 = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) { nq[k] = q[k] + dq4[k] * dm; n2 += nq[k] * nq[k]; }
    const float nrm = sqrtf(n2);
    float qn[4], gnew[4], dot = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) { qn[k] = nq[k] / nrm; gnew[k] = dquat[row * 4 + k]; dot += qn[k] * gnew[k]; }
    float dqun[4], g[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { dqun[k] = (gnew[k] - qn[k] * dot) / nrm; g[k] = dqun[k] * dm; }
    const float v1 = u[0], v2 = u[1], v3 = u[2];
    float dp[4];
    dp[0] = g[1] * v1 + g[2] * v2 + g[3] * v3;
    dp[1] = -g[0] * v1 - g[2] * v3 + g[3] * v2;
    dp[2] = -g[0] * v2 + g[1] * v3 - g[3] * v1;
    dp[3] = -g[0] * v3 - g[1] * v2 + g[2] * v1;
    du[0] = -g[0] * q[1] + g[1] * q[0] + g[2] * q[3] - g[3] * q[2];
    du[1] = -g[0] * q[2] - g[1] * q[3] + g[2] * q[0] + g[3] * q[1];
    du[2] = -g[0] * q[3] + g[1] * q[2] - g[2] * q[1] + g[3] * q[0];
    const float gt[3] = {dtrans[row * 3] * dm, dtrans[row * 3 + 1] * dm, dtrans[row * 3 + 2] * dm};
    du[3] = R[0] * gt[0] + R[3] * gt[1] + R[6] * gt[2];
    du[4] = R[1] * gt[0] + R[4] * gt[1] + R[7] * gt[2];
    du[5] = R[2] * gt[0] + R[5] * gt[1] + R[8] * gt[2];
    float G[9];
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
      for (int k = 0; k < 3; ++k) G[m * 3 + k] = gt[m] * u[3 + k];
    float dqr[4];
    quat_grad_from_rot_grad(q, G, dqr);
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) dquat[row * 4 + k] = dqun[k] + dp[k] + dqr[k];
      // dtrans passes through unchanged (t' = t + ...)
    }
    // dnode += (du . Wbb) * dm ; dW += du (x) x*dm ; db += du
    float* dn = dnode + row * C_S;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float acc = 0.f;
#pragma unroll
      for (int o = 0; o < 6; ++o) { acc = fmaf(du[o], Wbb[o * C_S + lane + k * 32], acc); dwacc[o][k] = fmaf(du[o], xv[k], dwacc[o][k]); }
      dn[lane + k * 32] += acc * dm;
    }
#pragma unroll
    for (int o = 0; o < 6; ++o) dbacc[o] += du[o];
  }
#pragma unroll
  for (int o = 0; o < 6; ++o)
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(dW + o * C_S + lane + k * 32, dwacc[o][k]);
  if (lane == 0) {
#pragma unroll
    for (int o = 0; o < 6; ++o) wred[warp][o] = dbacc[o];
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += wred[w][threadIdx.x];
    atomicAdd(dB + threadIdx.x, s);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// IGSO(3) score s(omega) = d_sigma/(p + 1e-4) and ds/domega (so3_diffuser.py:9-49,71-117), float64, warp-cooperative.
// (The forward value keeps the reference's fp32/fp64 mix — igso3_score_scalar; the derivative of that expression with respect to
// omega is evaluated analytically in fp64: autograd differentiates the same formula, its fp32 rounding is not part of the derivative.)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void igso3_score_and_derivative(double omega, double sigma, int lane, double& s, double& ds) {
  const double lo = sin(omega * 0.5), clo = cos(omega * 0.5);
  const double s2h = sigma * sigma;
  double psum = 0.0, dsum = 0.0, ddsum = 0.0;
  for (int l = lane; l < IGSO3_L; l += 32) {
    const double ex = -(double)((long long)l * (l + 1)) * s2h / 2.0;
    if (ex < -745.2) break;
    const double gauss = (double)(2 * l + 1) * exp(ex);
    const double a = (double)l + 0.5;
    double hi, chi;
    sincos(a * omega, &hi, &chi);
    const double num = lo * a * chi - hi * 0.5 * clo;
    const double dnum = -lo * a * a * hi + hi * 0.25 * lo;        // d/domega num (the two cross terms cancel)
    psum += gauss * hi / lo;
    dsum += gauss * num / (lo * lo);
    ddsum += gauss * (dnum / (lo * lo) - num * clo / (lo * lo * lo));
  }
  psum = warp_sum_d(psum); dsum = warp_sum_d(dsum); ddsum = warp_sum_d(ddsum);
  const double den = psum + 1e-4;
  s = dsum / den;
  ds = ddsum / den - dsum * dsum / (den * den);
}

// ---------------------------------------------------------------------------------------------------------------------
// Heads backward (forward: score_head_kernel).  One warp per residue.  Upstream gradients (any may be NULL = zero):
//   d_rot [R,3] f64, d_trans [R,3] f64, d_rigids [R,7] f32, d_atom37 [R,37,3] f32 (atoms 0..4 used), d_atom14 [R,14,3] f32, d_psi [R,2] f32
// Outputs: dquat [R,4], dtrans [R,3] (scaled units) — WRITTEN; d tors_s [R,256] (gradient at the torsion head's hidden) — WRITTEN;
//          dWf [2][256], dbf [2] accumulated.
// ---------------------------------------------------------------------------------------------------------------------
struct HeadBwdArgs {
  const float* tors_s; const float* Wf; const float* bf;
  const float* quat; const float* trans; const float* rigids_t;
  const double* t; int t_is_f32; const double* sigma_grid;
  const float* res_mask; const float* fixed_mask; const float* gt_psi;
  const double* d_rot; const double* d_trans; const float* d_rigids; const float* d_atom37; const float* d_atom14; const float* d_psi;
  float* dquat; float* dtrans; float* dtors; float* dWf; float* dbf;
  long long rows; int N;
};
__global__ void __launch_bounds__(256) score_head_bwd_kernel(const HeadBwdArgs a) {
  __shared__ float wacc[8][2 * C_S + 2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float dw0[8], dw1[8], db0 = 0.f, db1 = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { dw0[k] = 0.f; dw1[k] = 0.f; }
  for (long long row = (long long)blockIdx.x * 8 + warp; row < a.rows; row += (long long)gridDim.x * 8) {
    const int b = (int)(row / a.N);
    const float m = a.res_mask[row];
    const float dmk = 1.f - a.fixed_mask[row];
    // ---- recompute the torsion head ----
    float u0 = 0.f, u1 = 0.f, sv[8];
    const float* s = a.tors_s + row * C_S;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      sv[k] = s[lane + k * 32];
      u0 = fmaf(a.Wf[lane + k * 32], sv[k], u0);
      u1 = fmaf(a.Wf[C_S + lane + k * 32], sv[k], u1);
    }
    u0 = warp_sum(u0) + a.bf[0];
    u1 = warp_sum(u1) + a.bf[1];
    const float ssq = u0 * u0 + u1 * u1;
    const float den = sqrtf(fmaxf(ssq, 1e-8f));
    float p0 = u0 / den, p1 = u1 / den;
    const float g0 = a.gt_psi ? a.gt_psi[row * 2] : 0.f, g1 = a.gt_psi ? a.gt_psi[row * 2 + 1] : 0.f;
    const float ps = dmk * p0 + (1.f - dmk) * g0, pc = dmk * p1 + (1.f - dmk) * g1;     // (sin, cos) that place the O atom
    // ---- frames ----
    float q[4], R[9];
#pragma unroll
    for (int k = 0; k < 4; ++k) q[k] = a.quat[row * 4 + k];
    quat_to_rot(q, R);
    float dq[4] = {0.f, 0.f, 0.f, 0.f}, dtp[3] = {0.f, 0.f, 0.f}, dpsi[2] = {0.f, 0.f};
    if (a.d_psi) { dpsi[0] = a.d_psi[row * 2]; dpsi[1] = a.d_psi[row * 2 + 1]; }
    if (a.d_rigids) {
#pragma unroll
      for (int k = 0; k < 4; ++k) dq[k] += a.d_rigids[row * 7 + k];
#pragma unroll
      for (int k = 0; k < 3; ++k) dtp[k] += a.d_rigids[row * 7 + 4 + k];
    }
    const double tt = a.t[b];
    if (a.d_trans) {
      const double beta = tt * R3_MIN_B + 0.5 * (tt * tt) * (R3_MAX_B - R3_MIN_B);
      const double coef = exp(-0.5 * beta) * (double)COORD_SCALE / (1.0 - exp(-beta)) * (double)m;
#pragma unroll
      for (int k = 0; k < 3; ++k) dtp[k] += (float)(a.d_trans[row * 3 + k] * coef);
    }
    // ---- atoms: N, CA, C, CB rigid in the frame; O through the psi frame ----
    {
      float G[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      bool any = false;
      float dN[3] = {0, 0, 0}, dCA[3] = {0, 0, 0}, dC[3] = {0, 0, 0}, dCB[3] = {0, 0, 0}, dO[3] = {0, 0, 0};
      if (a.d_atom37) {
        const float* d = a.d_atom37 + row * 111;
#pragma unroll
        for (int k = 0; k < 3; ++k) { dN[k] += d[k]; dCA[k] += d[3 + k]; dC[k] += d[6 + k]; dCB[k] += d[9 + k]; dO[k] += d[12 + k]; }
        any = true;
      }
      if (a.d_atom14) {
        const float* d = a.d_atom14 + row * 42;
#pragma unroll
        for (int k = 0; k < 3; ++k) { dN[k] += d[k]; dCA[k] += d[3 + k]; dC[k] += d[6 + k]; dO[k] += d[9 + k]; dCB[k] += d[12 + k]; }
        any = true;
      }
      if (any) {
        const float nl[3] = {-0.525f, 1.363f, 0.f}, cl[3] = {1.526f, 0.f, 0.f}, cbl[3] = {-0.529f, -0.774f, -1.205f};
        // O local in the backbone frame: C + diag(1,-1,-1) Rx(psi) (0.627, 1.062, 0) = (1.526 + 0.627, -c*1.062, -s*1.062)
        const float ol[3] = {1.526f + 0.627f, -pc * 1.062f, -ps * 1.062f};
#pragma unroll
        for (int mm = 0; mm < 3; ++mm)
#pragma unroll
          for (int k = 0; k < 3; ++k) G[mm * 3 + k] = dN[mm] * nl[k] + dC[mm] * cl[k] + dCB[mm] * cbl[k] + dO[mm] * ol[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) dtp[k] += dN[k] + dCA[k] + dC[k] + dCB[k] + dO[k];
        float dloc[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) dloc[k] = R[0 * 3 + k] * dO[0] + R[1 * 3 + k] * dO[1] + R[2 * 3 + k] * dO[2];
        dpsi[0] += -1.062f * dloc[2];
        dpsi[1] += -1.062f * dloc[1];
        float dqa[4];
        quat_grad_from_rot_grad(q, G, dqa);
#pragma unroll
        for (int k = 0; k < 4; ++k) dq[k] += dqa[k];
      }
    }
    // ---- torsion head: psi = dm*un/den + (1-dm)*gt ----
    {
      const float d0 = dpsi[0] * dmk, d1 = dpsi[1] * dmk;
      float du0 = d0 / den, du1 = d1 / den;
      if (ssq > 1e-8f) {
        const float dot = (u0 * d0 + u1 * d1) / (den * den * den);
        du0 -= u0 * dot; du1 -= u1 * dot;
      }
      float* dt_ = a.dtors + row * C_S;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        dt_[lane + k * 32] = du0 * a.Wf[lane + k * 32] + du1 * a.Wf[C_S + lane + k * 32];
        dw0[k] = fmaf(du0, sv[k], dw0[k]);
        dw1[k] = fmaf(du1, sv[k], dw1[k]);
      }
      db0 += du0; db1 += du1;
    }
    // ---- rotation score ----
    if (a.d_rot) {
      float qt[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) qt[k] = a.rigids_t[row * 7 + k];
      const float n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
      const float qi[4] = {q[0] / n2, -q[1] / n2, -q[2] / n2, -q[3] / n2};
      float qr[4];
      quat_mul(qi, qt, qr);
      const float sgn = qr[0] < 0.f ? -1.f : 1.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) qr[k] *= sgn;
      const float vn = sqrtf(qr[1] * qr[1] + qr[2] * qr[2] + qr[3] * qr[3]);
      const float angle = 2.f * atan2f(vn, qr[0]);
      const float a2 = angle * angle;
      const bool small = angle <= 1e-3f;
      const float half = angle / 2.f + 1e-6f;
      const float sh = sinf(half), ch = cosf(half);
      const float sc = small ? 2.f + a2 / 12.f + 7.f * a2 * a2 / 2880.f : angle / sh;
      const float rv[3] = {sc * qr[1], sc * qr[2], sc * qr[3]};
      const float nv = sqrtf(rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]);
      const float omega = nv + 1e-6f;
      const double sig = quantise_sigma(tt, a.sigma_grid);
      double sval, dsval;
      igso3_score_and_derivative((double)omega, sig, lane, sval, dsval);
      const double om2 = (double)omega + 1e-6;
      const double f = sval / om2, dfdom = dsval / om2 - sval / (om2 * om2);
      double drs[3], dotr = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) { drs[k] = a.d_rot[row * 3 + k] * (double)m; dotr += drs[k] * (double)rv[k]; }
      float drv[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) drv[k] = (float)(f * drs[k] + dfdom * dotr * (double)rv[k] / fmax((double)nv, 1e-30));
      float dv[3] = {sc * drv[0], sc * drv[1], sc * drv[2]};
      const float dscale = drv[0] * qr[1] + drv[1] * qr[2] + drv[2] * qr[3];
      const float dsc_dang = small ? angle / 6.f + 7.f * a2 * angle / 720.f : 1.f / sh - angle * ch / (2.f * sh * sh);
      const float dang = dscale * dsc_dang;
      const float den2 = vn * vn + qr[0] * qr[0];
      const float dvn = dang * 2.f * qr[0] / den2;
      const float dw = -dang * 2.f * vn / den2;
      const float ivn = 1.f / fmaxf(vn, 1e-30f);
#pragma unroll
      for (int k = 0; k < 3; ++k) dv[k] += dvn * qr[1 + k] * ivn;
      const float dqrel[4] = {dw * sgn, dv[0] * sgn, dv[1] * sgn, dv[2] * sgn};
      const float qtc[4] = {qt[0], -qt[1], -qt[2], -qt[3]};
      float dqinv[4];
      quat_mul(dqrel, qtc, dqinv);
      const float cj[4] = {1.f, -1.f, -1.f, -1.f};
      float dotq = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) dotq += dqinv[k] * qi[k];
#pragma unroll
      for (int k = 0; k < 4; ++k) dq[k] += cj[k] * dqinv[k] / n2 - 2.f * q[k] * dotq / n2;
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) a.dquat[row * 4 + k] = dq[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) a.dtrans[row * 3 + k] = dtp[k] / COORD_SCALE;
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) { wacc[warp][lane + k * 32] = dw0[k]; wacc[warp][C_S + lane + k * 32] = dw1[k]; }
  if (lane == 0) { wacc[warp][2 * C_S] = db0; wacc[warp][2 * C_S + 1] = db1; }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C_S + 2; c += 256) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += wacc[w][c];
    if (c < 2 * C_S) atomicAdd(a.dWf + c, s); else atomicAdd(a.dbf + (c - 2 * C_S), s);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// DSM loss backward (train_se3_diffusion.py:538-680): gradient of total_loss = sum_b batch_loss[b] /
// (#samples with any residue + 1e-10) w.r.t. the model outputs (kernels: loss_count_kernel + loss_bwd2_kernel below).
//   d_rot [B,N,3] f64, d_trans [B,N,3] f64, d_rigids [B,N,7] f32 (only the translation part is non-zero), d_atom37 [B,N,37,3] f32
// ---------------------------------------------------------------------------------------------------------------------
struct LossBwdArgs {
  const double* pred_rot; const double* pred_trans; const float* pred_rigids; const float* pred_atom37;
  const double* gt_rot; const double* gt_trans; const double* rot_scaling; const double* trans_scaling; const double* rigids_0; const double* t;
  const float* res_mask; const float* fixed_mask; const float* gt_psi;
  double trans_loss_weight, rot_loss_weight, rot_loss_t_threshold, trans_x0_threshold, coordinate_scaling, bb_atom_loss_weight,
      bb_atom_loss_t_filter, dist_mat_loss_weight, dist_mat_loss_t_filter, aux_loss_weight;
  int separate_rot_loss, diffuse_trans, diffuse_rot;
  double inv_nvalid;          // 1 / (number of samples with a non-empty res_mask + 1e-10), computed by the caller
  double* d_rot; double* d_trans; float* d_rigids; float* d_atom37;
  int N;
};
__device__ __forceinline__ double block_sum_f64_t(double v, double* red) {   // 256 threads; red[8]; result broadcast
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum_d(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w];
  return s;
}

// ---------------------------------------------------------------------------------------------------------------------
// DSM loss kernels: the arithmetic of Experiment.loss_fn spread over grid (B, slabs) — with ONE CTA per example the 5N x 5N pair loop
// was 6 % of a training step at B = 8.  Every CTA stages the example's ground-truth and
// predicted backbone atoms in shared memory (cheap, O(N)) and owns LOSS_RES residues: their per-residue terms and their 5*LOSS_RES
// rows of the pair matrix.  Forward: partial sums -> atomicAdd(double) into acc[b][10], a second tiny kernel forms the terms.
// Backward: the pair denominator comes from a counting launch; row p's gradient is sum_q (c_pq + c_qp)(x_p - x_q) — both ordered pairs
// evaluated by the owner of p, so no atomics and no second writer.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int LOSS_RES = 16;
struct LossStage {            // shared-memory staging of one example
  float* gt5; float* pr5; float* fl; float* fr;   // [5N][3], [5N][3], [N] loss mask, [N] res mask
};
__device__ __forceinline__ LossStage loss_stage(float* lsm, int N, int b, const double* rigids_0, const float* gt_psi, const float* pred_atom37,
                                                const float* res_mask, const float* fixed_mask) {
  LossStage S{lsm, lsm + 15 * N, lsm + 30 * N, lsm + 31 * N};
  const long long r0 = (long long)b * N;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float q[4], R[9], t3[3], a37[111];
#pragma unroll
    for (int k = 0; k < 4; ++k) q[k] = (float)rigids_0[(r0 + n) * 7 + k];
#pragma unroll
    for (int k = 0; k < 3; ++k) t3[k] = (float)rigids_0[(r0 + n) * 7 + 4 + k];
    quat_to_rot(q, R);
    backbone_atoms(R, t3, gt_psi[(r0 + n) * 2], gt_psi[(r0 + n) * 2 + 1], a37, nullptr);
#pragma unroll
    for (int k = 0; k < 15; ++k) { S.gt5[n * 15 + k] = a37[k]; S.pr5[n * 15 + k] = pred_atom37[(r0 + n) * 111 + k]; }
    S.fl[n] = res_mask[r0 + n] * (1.f - fixed_mask[r0 + n]);
    S.fr[n] = res_mask[r0 + n];
  }
  __syncthreads();
  return S;
}
__device__ __forceinline__ bool atom_present(const float* g) { return (fabsf(g[0]) + fabsf(g[1]) + fabsf(g[2])) != 0.f; }

// acc[b][0..9] = { mask, trans_score, trans_x0, axis, angle, rot, bb, bbm, pair sum, pair count }
__global__ void __launch_bounds__(256) loss_fwd2_kernel(const LossArgs a, double* __restrict__ acc) {
  extern __shared__ __align__(16) float lsm[];
  __shared__ double red[8];
  const int N = a.N, b = blockIdx.x, tid = threadIdx.x;
  const LossStage S = loss_stage(lsm, N, b, a.rigids_0, a.gt_psi, a.pred_atom37, a.res_mask, a.fixed_mask);
  const int n0 = blockIdx.y * LOSS_RES, n1 = min(N, n0 + LOSS_RES);
  double v[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = n0 + tid; i < n1; i += 256) {
    const long long r = (long long)b * N + i;
    const double dm = 1.0 - (double)a.fixed_mask[r], lm = (double)a.res_mask[r] * dm;
    v[0] += lm;
    double ga = 0.0, pa = 0.0, g[3], q[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double pt = a.pred_trans[r * 3 + k] * dm, dt = a.gt_trans[r * 3 + k] - pt;
      v[1] += dt * dt * lm;
      const double x0 = a.rigids_0[r * 7 + 4 + k] * a.coordinate_scaling - (double)a.pred_rigids[r * 7 + 4 + k] * a.coordinate_scaling;
      v[2] += x0 * x0 * lm;
      g[k] = a.gt_rot[r * 3 + k]; q[k] = a.pred_rot[r * 3 + k] * dm;
      ga += g[k] * g[k]; pa += q[k] * q[k];
      const double dr = g[k] - q[k];
      v[5] += dr * dr * lm;
    }
    ga = sqrt(ga); pa = sqrt(pa);
#pragma unroll
    for (int k = 0; k < 3; ++k) { const double d = g[k] / (ga + 1e-6) - q[k] / (pa + 1e-6); v[3] += d * d * lm; }
    v[4] += (ga - pa) * (ga - pa) * lm;
#pragma unroll
    for (int at = 0; at < 5; ++at) {
      const float* gp = S.gt5 + (i * 5 + at) * 3; const float* pp = S.pr5 + (i * 5 + at) * 3;
      const double mk = atom_present(gp) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) { const float df = pp[k] - gp[k]; v[6] += (double)(df * df) * mk * lm; }
      v[7] += mk * lm;
    }
  }
  // pair rows x of this slab (5 atoms per residue), all columns y
  const int M = 5 * N, x0r = 5 * n0, x1r = 5 * n1;
  for (long long pidx = tid; pidx < (long long)(x1r - x0r) * M; pidx += 256) {
    const int x = x0r + (int)(pidx / M), y = (int)(pidx % M);
    const float lmx = S.fl[x / 5], rmy = S.fr[y / 5];
    const float gx = S.gt5[x * 3] - S.gt5[y * 3], gy = S.gt5[x * 3 + 1] - S.gt5[y * 3 + 1], gz = S.gt5[x * 3 + 2] - S.gt5[y * 3 + 2];
    const float px = S.pr5[x * 3] - S.pr5[y * 3], py = S.pr5[x * 3 + 1] - S.pr5[y * 3 + 1], pz = S.pr5[x * 3 + 2] - S.pr5[y * 3 + 2];
    const float gd = sqrtf(gx * gx + gy * gy + gz * gz) * lmx, pd = sqrtf(px * px + py * py + pz * pz) * lmx;
    const double pm = (double)lmx * (double)rmy * (gd < 6.f ? 1.0 : 0.0);
    const double dd = (double)gd - (double)pd;
    v[8] += dd * dd * pm; v[9] += pm;
  }
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    const double s = block_sum_f64_t(v[k], red);
    if (tid == 0 && s != 0.0) atomicAdd(acc + (long long)b * 10 + k, s);
  }
}
__global__ void loss_finalize_kernel(const LossArgs a, const double* __restrict__ acc, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double* s = acc + (long long)b * 10;
  const double tb = a.t[b], denom = s[0] + 1e-10;
  const double ts2 = a.trans_scaling[b] * a.trans_scaling[b], rs2 = a.rot_scaling[b] * a.rot_scaling[b];
  double trans_loss = (tb > a.trans_x0_threshold ? (s[1] / ts2) / denom : 0.0) + (tb <= a.trans_x0_threshold ? s[2] / denom : 0.0);
  trans_loss *= a.trans_loss_weight * (double)a.diffuse_trans;
  double rot_loss;
  if (a.separate_rot_loss) rot_loss = (s[4] / rs2) / denom * a.rot_loss_weight * (tb > a.rot_loss_t_threshold ? 1.0 : 0.0) + s[3] / denom;
  else rot_loss = (s[5] / rs2) / denom * a.rot_loss_weight * (tb > a.rot_loss_t_threshold ? 1.0 : 0.0);
  rot_loss *= (double)a.diffuse_rot;
  const double bb_loss = s[6] / (s[7] + 1e-10) * a.bb_atom_loss_weight * (tb < a.bb_atom_loss_t_filter ? 1.0 : 0.0) * a.aux_loss_weight;
  const double dm_loss = s[8] / (s[9] - (double)a.N) * a.dist_mat_loss_weight * (tb < a.dist_mat_loss_t_filter ? 1.0 : 0.0) * a.aux_loss_weight;
  double* o = a.terms + (long long)b * 5;
  o[0] = rot_loss; o[1] = trans_loss; o[2] = bb_loss; o[3] = dm_loss; o[4] = ((rot_loss + trans_loss) + bb_loss) + dm_loss;
}

// acc[b][0..2] = { loss-mask sum, backbone-atom mask sum, pair-mask sum }  (denominators of the backward)
__global__ void __launch_bounds__(256) loss_count_kernel(const LossBwdArgs a, double* __restrict__ acc) {
  extern __shared__ __align__(16) float lsm[];
  __shared__ double red[8];
  const int N = a.N, b = blockIdx.x, tid = threadIdx.x;
  const LossStage S = loss_stage(lsm, N, b, a.rigids_0, a.gt_psi, a.pred_atom37, a.res_mask, a.fixed_mask);
  const int n0 = blockIdx.y * LOSS_RES, n1 = min(N, n0 + LOSS_RES);
  double v[3] = {0, 0, 0};
  for (int i = n0 + tid; i < n1; i += 256) {
    v[0] += (double)S.fl[i];
#pragma unroll
    for (int at = 0; at < 5; ++at) if (atom_present(S.gt5 + (i * 5 + at) * 3)) v[1] += (double)S.fl[i];
  }
  const int M = 5 * N, x0r = 5 * n0, x1r = 5 * n1;
  for (long long pidx = tid; pidx < (long long)(x1r - x0r) * M; pidx += 256) {
    const int x = x0r + (int)(pidx / M), y = (int)(pidx % M);
    const float lmx = S.fl[x / 5];
    const float gx = S.gt5[x * 3] - S.gt5[y * 3], gy = S.gt5[x * 3 + 1] - S.gt5[y * 3 + 1], gz = S.gt5[x * 3 + 2] - S.gt5[y * 3 + 2];
    if (sqrtf(gx * gx + gy * gy + gz * gz) * lmx < 6.f) v[2] += (double)lmx * (double)S.fr[y / 5];
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double s = block_sum_f64_t(v[k], red);
    if (tid == 0 && s != 0.0) atomicAdd(acc + (long long)b * 3 + k, s);
  }
}
__global__ void __launch_bounds__(256) loss_bwd2_kernel(const LossBwdArgs a, const double* __restrict__ acc) {
  extern __shared__ __align__(16) float lsm[];
  const int N = a.N, b = blockIdx.x, tid = threadIdx.x;
  const LossStage S = loss_stage(lsm, N, b, a.rigids_0, a.gt_psi, a.pred_atom37, a.res_mask, a.fixed_mask);
  const int n0 = blockIdx.y * LOSS_RES, n1 = min(N, n0 + LOSS_RES);
  const double tt = a.t[b], ws = a.inv_nvalid;
  const double denom = acc[(long long)b * 3] + 1e-10, bm_sum = acc[(long long)b * 3 + 1], pden = acc[(long long)b * 3 + 2] - (double)N;
  const double hi_t = tt > a.trans_x0_threshold ? 1.0 : 0.0;
  const double wa = a.rot_loss_weight * (tt > a.rot_loss_t_threshold ? 1.0 : 0.0);
  const double w_bb = a.bb_atom_loss_weight * (tt < a.bb_atom_loss_t_filter ? 1.0 : 0.0) * a.aux_loss_weight;
  const double w_dm = a.dist_mat_loss_weight * (tt < a.dist_mat_loss_t_filter ? 1.0 : 0.0) * a.aux_loss_weight;
  const double rsc = a.rot_scaling[b], tsc = a.trans_scaling[b];
  for (int n = n0 + tid; n < n1; n += 256) {
    const long long r = (long long)b * N + n;
    const double dmk = 1.0 - (double)a.fixed_mask[r];
    const double lm = (double)a.res_mask[r] * dmk;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double e = a.gt_trans[r * 3 + k] - a.pred_trans[r * 3 + k] * dmk;
      a.d_trans[r * 3 + k] = a.diffuse_trans ? -2.0 * e * lm / (tsc * tsc * denom) * (hi_t * a.trans_loss_weight * ws) * dmk : 0.0;
      const double x0g = a.rigids_0[r * 7 + 4 + k] * a.coordinate_scaling, x0p = (double)a.pred_rigids[r * 7 + 4 + k] * a.coordinate_scaling;
      a.d_rigids[r * 7 + 4 + k] =
          (float)(a.diffuse_trans ? -2.0 * (x0g - x0p) * lm / denom * ((1.0 - hi_t) * a.trans_loss_weight * ws) * a.coordinate_scaling : 0.0);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) a.d_rigids[r * 7 + k] = 0.f;
    double pr[3], gr[3], pa = 0.0, ga = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) { pr[k] = a.pred_rot[r * 3 + k] * dmk; gr[k] = a.gt_rot[r * 3 + k]; pa += pr[k] * pr[k]; ga += gr[k] * gr[k]; }
    pa = sqrt(pa); ga = sqrt(ga);
    double dpr[3];
    if (a.separate_rot_loss) {
      double dpax[3], dotp = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) { dpax[k] = -2.0 * (gr[k] / (ga + 1e-6) - pr[k] / (pa + 1e-6)) * lm / denom * ws; dotp += dpax[k] * pr[k]; }
      const double dpa = -2.0 * (ga - pa) * lm / (rsc * rsc * denom) * (wa * ws);
      const double ipa_ = 1.0 / fmax(pa, 1e-30);
#pragma unroll
      for (int k = 0; k < 3; ++k) dpr[k] = dpax[k] / (pa + 1e-6) - dotp / ((pa + 1e-6) * (pa + 1e-6)) * pr[k] * ipa_ + dpa * pr[k] * ipa_;
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) dpr[k] = -2.0 * (gr[k] - pr[k]) * lm / (rsc * rsc * denom) * (wa * ws);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) a.d_rot[r * 3 + k] = a.diffuse_rot ? dpr[k] * dmk : 0.0;
    for (int k = 15; k < 111; ++k) a.d_atom37[r * 111 + k] = 0.f;
  }
  // rows p of this slab: backbone-atom term + pair term, one thread per row
  const int M5 = 5 * N;
  for (int p = 5 * n0 + tid; p < 5 * n1; p += 256) {
    const int n = p / 5;
    const long long r = (long long)b * N + n;
    const float flp = S.fl[n], frp = S.fr[n];
    const float gx = S.gt5[p * 3], gy = S.gt5[p * 3 + 1], gz = S.gt5[p * 3 + 2];
    const float px = S.pr5[p * 3], py = S.pr5[p * 3 + 1], pz = S.pr5[p * 3 + 2];
    const double bm = atom_present(S.gt5 + p * 3) ? (double)flp : 0.0;
    const double cbb = 2.0 * bm / (bm_sum + 1e-10) * (w_bb * ws);
    double ax = cbb * ((double)px - (double)gx), ay = cbb * ((double)py - (double)gy), az = cbb * ((double)pz - (double)gz);
    if (w_dm != 0.0) {
      for (int qx = 0; qx < M5; ++qx) {
        const float flq = S.fl[qx / 5], frq = S.fr[qx / 5];
        const float ex = gx - S.gt5[qx * 3], ey = gy - S.gt5[qx * 3 + 1], ez = gz - S.gt5[qx * 3 + 2];
        const float gdist = sqrtf(ex * ex + ey * ey + ez * ez);
        const float dxp = px - S.pr5[qx * 3], dyp = py - S.pr5[qx * 3 + 1], dzp = pz - S.pr5[qx * 3 + 2];
        const float pdr = sqrtf(dxp * dxp + dyp * dyp + dzp * dzp);
        if (!(pdr > 0.f)) continue;
        double c = 0.0;
        {   // ordered pair (p, q): mask fl_p fr_q [gd fl_p < 6], distances scaled by fl_p
          const float gd = gdist * flp;
          if (gd < 6.f && flp != 0.f && frq != 0.f)
            c += -2.0 * ((double)gd - (double)pdr * (double)flp) * (double)(flp * frq) / pden * (w_dm * ws) * (double)flp / (double)pdr;
        }
        {   // ordered pair (q, p)
          const float gd = gdist * flq;
          if (gd < 6.f && flq != 0.f && frp != 0.f)
            c += -2.0 * ((double)gd - (double)pdr * (double)flq) * (double)(flq * frp) / pden * (w_dm * ws) * (double)flq / (double)pdr;
        }
        ax += c * dxp; ay += c * dyp; az += c * dzp;
      }
    }
    a.d_atom37[r * 111 + (p - 5 * n) * 3 + 0] = (float)ax;
    a.d_atom37[r * 111 + (p - 5 * n) * 3 + 1] = (float)ay;
    a.d_atom37[r * 111 + (p - 5 * n) * 3 + 2] = (float)az;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Intermittent eval metrics on the device (analysis/metrics.py:120-132 ca_ca_distance / ca_ca_clashes): per backbone
//   out[b] = { mean |d_i - 3.80209737096|, fraction of bonds d_i < 3.802 + tol_bond, number of CA pairs closer than tol_clash,
//              fraction of such pairs among the pairs with distance > 0 }       d_i = |CA_i - CA_{i-1}|, valid residues first (mask prefix)
// ca [B,N,3] fp32 (Angstrom); n_valid[b] residues are used (the reference slices the unpadded chain before calling these functions).
// One CTA per backbone; fp64 accumulation like numpy's float64 means.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ca_metrics_kernel(const float* __restrict__ ca, const int* __restrict__ n_valid, double tol_bond, double tol_clash,
                                                         double* __restrict__ out, int N) {
  __shared__ double red[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = n_valid ? min(n_valid[b], N) : N;
  const float* x = ca + (long long)b * N * 3;
  const double CA_CA = 3.80209737096;
  double dev = 0.0, valid = 0.0;
  for (int i = 1 + tid; i < n; i += 256) {
    const float dx = x[i * 3] - x[(i - 1) * 3], dy = x[i * 3 + 1] - x[(i - 1) * 3 + 1], dz = x[i * 3 + 2] - x[(i - 1) * 3 + 2];
    const float d = sqrtf(dx * dx + dy * dy + dz * dz);        // np.linalg.norm of a float32 array stays float32
    dev += fabs((double)d - CA_CA);
    valid += (double)d < CA_CA + tol_bond ? 1.0 : 0.0;
  }
  double clashes = 0.0, pairs = 0.0;
  for (long long p = tid; p < (long long)n * n; p += 256) {
    const int i = (int)(p / n), j = (int)(p - (long long)i * n);
    if (j < i) continue;                                        // np.triu(k=0); the diagonal has distance 0 and is dropped by `> 0`
    const float dx = x[i * 3] - x[j * 3], dy = x[i * 3 + 1] - x[j * 3 + 1], dz = x[i * 3 + 2] - x[j * 3 + 2];
    const float d = sqrtf(dx * dx + dy * dy + dz * dz);
    if (d > 0.f) { pairs += 1.0; clashes += (double)d < tol_clash ? 1.0 : 0.0; }
  }
  dev = block_sum_f64_t(dev, red); valid = block_sum_f64_t(valid, red);
  clashes = block_sum_f64_t(clashes, red); pairs = block_sum_f64_t(pairs, red);
  if (tid == 0) {
    const double nb = n > 1 ? (double)(n - 1) : 1.0;
    out[b * 4 + 0] = dev / nb; out[b * 4 + 1] = valid / nb; out[b * 4 + 2] = clashes; out[b * 4 + 3] = pairs > 0.0 ? clashes / pairs : 0.0;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam defaults of the reference: betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad;
// experiments/train_se3_diffusion.py:139-141) over the flat parameter arena.  skip[i] != 0 marks elements of parameters that
// received no gradient this step (the reference's unused parameters keep grad None and are skipped by the optimiser).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                            float lr, float b1, float b2, float eps, float bc1, float bc2, float gscale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gr = g[i] * gscale;
  const float mi = b1 * m[i] + (1.f - b1) * gr;
  const float vi = b2 * v[i] + (1.f - b2) * gr * gr;
  m[i] = mi; v[i] = vi;
  const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
  p[i] -= (lr / bc1) * (mi / denom);
}

}  // namespace fd
