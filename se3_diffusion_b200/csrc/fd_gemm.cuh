// Generic fp32 CUDA-core GEMM with a fused epilogue.  Used for the O(N) node path (projections, transformer,
// transitions: ~3 % of the forward's FLOPs) and, in FD_PREC_FP32 mode, for the edge-tensor MLPs.
//   C[z][m][n] = epi( alpha * sum_k A[z][m][k] * Bop[z][k][n] )
//   B_KMAJOR = true : B is W[n][k] (PyTorch Linear weight, row-major, K contiguous)           ("TN")
//   B_KMAJOR = false: B is Bm[k][n] (N contiguous), e.g. attention values                      ("NN")
// Epilogue order: +bias[n]  +rowadd_i/rowadd_j (edge rows -> (b,i,j))  +C (accumulate)  relu  *rowmask[m]  +residual.
#pragma once
#include "fd_common.cuh"

namespace fd {

struct GemmArgs {
  const float* A = nullptr; int lda = 0; long long sA0 = 0, sA1 = 0;
  const float* B = nullptr; int ldb = 0; long long sB0 = 0, sB1 = 0;
  float* C = nullptr; int ldc = 0; long long sC0 = 0, sC1 = 0;
  int M = 0, N = 0, K = 0;
  int nb0 = 1, nb1 = 1;                 // batch = nb0 * nb1, z = z0 * nb1 + z1
  float alpha = 1.f;
  const float* bias = nullptr;
  const float* residual = nullptr; int ldr = 0; long long sR0 = 0, sR1 = 0;
  const float* rowmask = nullptr;       // [M] (not batched)
  int relu = 0, accumulate = 0;
  // edge-row broadcast adds: global row g = row_offset + m -> b = g / (nres*nres), i = (g / nres) % nres, j = g % nres
  const float* rowadd_i = nullptr; const float* rowadd_j = nullptr; int ld_rowadd = 0; int nres = 0; long long row_offset = 0;
  int vecA = 0, vecB = 0;               // 16-byte vector loads allowed (pointer, ld and batch strides 4-float aligned)
  // training-path extensions
  const float* relumask = nullptr; int ldm = 0;   // v = relumask[m][n] > 0 ? v : 0 (backward of a ReLU whose OUTPUT is relumask), applied last
  int splits = 1;                       // split-K: blockIdx.z = batch * splits + split, partial sums combined with atomicAdd (needs atomic = 1)
  int atomic = 0;                       // C[m][n] += alpha*acc through atomicAdd (gradient accumulation); bias/rowadd/relu/residual ignored
  float* colsum_out = nullptr;          // weight-gradient form only (A = dy[k][m]): colsum_out[m] += sum_k A[k][m] — the bias gradient, taken from the
                                        // A tiles the n-tile-0 CTAs stream anyway (mm3_kernel; the CUDA-core kernel ignores it)
};

// A_KMAJOR = true : A is A[m][k] (row-major activations, K contiguous)
// A_KMAJOR = false: A is At[k][m] (M contiguous) — weight gradients dW[m][n] = sum_k dY[k][m] X[k][n] read both operands untransposed
template <int BM, int BN, int TM, int TN, bool B_KMAJOR, bool A_KMAJOR = true>
__global__ void __launch_bounds__(256) gemm_kernel(const GemmArgs g) {
  constexpr int BK = 8;
  constexpr int LDA_S = BM + 4, LDB_S = BN + 4;
  static_assert((BM / TM) * (BN / TN) == 256, "256 threads");
  __shared__ __align__(16) float As[2][BK][LDA_S];
  __shared__ __align__(16) float Bs[2][BK][LDB_S];

  const int tid = threadIdx.x;
  const int zs = blockIdx.z / g.splits, split = blockIdx.z - zs * g.splits;
  const int z = zs, z0 = z / g.nb1, z1 = z % g.nb1;
  const float* __restrict__ A = g.A + z0 * g.sA0 + z1 * g.sA1;
  const float* __restrict__ B = g.B + z0 * g.sB0 + z1 * g.sB1;
  float* __restrict__ C = g.C + z0 * g.sC0 + z1 * g.sC1;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int M = g.M, N = g.N;
  // split-K: this CTA covers k in [kbeg, K)
  int kbeg = 0, K = g.K;
  if (g.splits > 1) {
    const int per = ((g.K + g.splits - 1) / g.splits + BK - 1) / BK * BK;
    kbeg = split * per;
    K = min(g.K, kbeg + per);
    if (kbeg >= K) return;
  }

  constexpr int TX = BN / TN;           // thread columns
  const int tx = tid % TX, ty = tid / TX;

  constexpr int A_F4 = BM * BK / 4, B_F4 = BN * BK / 4;
  constexpr int A_PER = (A_F4 + 255) / 256, B_PER = (B_F4 + 255) / 256;
  float4 ra[A_PER], rb[B_PER];

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int it = 0; it < A_PER; ++it) {
      const int idx = tid + it * 256;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < A_F4) {
        if (A_KMAJOR) {
          const int row = idx / (BK / 4), kq = (idx % (BK / 4)) * 4;
          const int m = m0 + row, k = k0 + kq;
          if (m < M) {
            const float* p = A + (long long)m * g.lda + k;
            if (g.vecA && k + 3 < K) v = *reinterpret_cast<const float4*>(p);
            else {
              if (k < K) v.x = p[0];
              if (k + 1 < K) v.y = p[1];
              if (k + 2 < K) v.z = p[2];
              if (k + 3 < K) v.w = p[3];
            }
          }
        } else {
          const int kr = idx / (BM / 4), mq = (idx % (BM / 4)) * 4;
          const int k = k0 + kr, m = m0 + mq;
          if (k < K) {
            const float* p = A + (long long)k * g.lda + m;
            if (g.vecA && m + 3 < M) v = *reinterpret_cast<const float4*>(p);
            else {
              if (m < M) v.x = p[0];
              if (m + 1 < M) v.y = p[1];
              if (m + 2 < M) v.z = p[2];
              if (m + 3 < M) v.w = p[3];
            }
          }
        }
      }
      ra[it] = v;
    }
#pragma unroll
    for (int it = 0; it < B_PER; ++it) {
      const int idx = tid + it * 256;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < B_F4) {
        if (B_KMAJOR) {
          const int row = idx / (BK / 4), kq = (idx % (BK / 4)) * 4;
          const int n = n0 + row, k = k0 + kq;
          if (n < N) {
            const float* p = B + (long long)n * g.ldb + k;
            if (g.vecB && k + 3 < K) v = *reinterpret_cast<const float4*>(p);
            else {
              if (k < K) v.x = p[0];
              if (k + 1 < K) v.y = p[1];
              if (k + 2 < K) v.z = p[2];
              if (k + 3 < K) v.w = p[3];
            }
          }
        } else {
          const int kr = idx / (BN / 4), nq = (idx % (BN / 4)) * 4;
          const int k = k0 + kr, n = n0 + nq;
          if (k < K) {
            const float* p = B + (long long)k * g.ldb + n;
            if (g.vecB && n + 3 < N) v = *reinterpret_cast<const float4*>(p);
            else {
              if (n < N) v.x = p[0];
              if (n + 1 < N) v.y = p[1];
              if (n + 2 < N) v.z = p[2];
              if (n + 3 < N) v.w = p[3];
            }
          }
        }
      }
      rb[it] = v;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int it = 0; it < A_PER; ++it) {
      const int idx = tid + it * 256;
      if (idx < A_F4) {
        if (A_KMAJOR) {
          const int row = idx / (BK / 4), kq = (idx % (BK / 4)) * 4;
          As[buf][kq + 0][row] = ra[it].x; As[buf][kq + 1][row] = ra[it].y;
          As[buf][kq + 2][row] = ra[it].z; As[buf][kq + 3][row] = ra[it].w;
        } else {
          const int kr = idx / (BM / 4), mq = (idx % (BM / 4)) * 4;
          *reinterpret_cast<float4*>(&As[buf][kr][mq]) = ra[it];
        }
      }
    }
#pragma unroll
    for (int it = 0; it < B_PER; ++it) {
      const int idx = tid + it * 256;
      if (idx < B_F4) {
        if (B_KMAJOR) {
          const int row = idx / (BK / 4), kq = (idx % (BK / 4)) * 4;
          Bs[buf][kq + 0][row] = rb[it].x; Bs[buf][kq + 1][row] = rb[it].y;
          Bs[buf][kq + 2][row] = rb[it].z; Bs[buf][kq + 3][row] = rb[it].w;
        } else {
          const int kr = idx / (BN / 4), nq = (idx % (BN / 4)) * 4;
          *reinterpret_cast<float4*>(&Bs[buf][kr][nq]) = rb[it];
        }
      }
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // Row/col ownership: TM (TN) is split in float4 groups strided by BM/(TM/4) so smem reads are conflict-free.
  constexpr int RG = TM / 4, CG = TN / 4;
  constexpr int RSTRIDE = BM / RG, CSTRIDE = BN / CG;

  const int nk = (K - kbeg + BK - 1) / BK;
  load_tiles(kbeg);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles(kbeg + (kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int r = 0; r < RG; ++r) {
        const float4 v = *reinterpret_cast<const float4*>(&As[buf][k][r * RSTRIDE + ty * 4]);
        a[r * 4 + 0] = v.x; a[r * 4 + 1] = v.y; a[r * 4 + 2] = v.z; a[r * 4 + 3] = v.w;
      }
#pragma unroll
      for (int c = 0; c < CG; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][c * CSTRIDE + tx * 4]);
        b[c * 4 + 0] = v.x; b[c * 4 + 1] = v.y; b[c * 4 + 2] = v.z; b[c * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue
  const float* R = g.residual ? g.residual + z0 * g.sR0 + z1 * g.sR1 : nullptr;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + (i / 4) * RSTRIDE + ty * 4 + (i % 4);
    if (m >= M) continue;
    const float* pi = nullptr; const float* pj = nullptr;
    if (g.rowadd_i) {
      const long long grow = g.row_offset + m;
      const long long nn = (long long)g.nres * g.nres;
      const long long b = grow / nn;
      const int rem = (int)(grow - b * nn);
      const int ri = rem / g.nres, rj = rem - ri * g.nres;
      pi = g.rowadd_i + (b * g.nres + ri) * g.ld_rowadd;
      pj = g.rowadd_j + (b * g.nres + rj) * g.ld_rowadd;
    }
    const float rm = g.rowmask ? g.rowmask[m] : 1.f;
#pragma unroll
    for (int c = 0; c < CG; ++c) {
      const int nb = n0 + c * CSTRIDE + tx * 4;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int n = nb + jj;
        if (n >= N) continue;
        float v = g.alpha * acc[i][c * 4 + jj];
        if (g.atomic) { atomicAdd(C + (long long)m * g.ldc + n, v); continue; }
        if (g.bias) v += g.bias[n];
        if (pi) v += pi[n] + pj[n];
        float* cp = C + (long long)m * g.ldc + n;
        if (g.accumulate) v += *cp;
        if (g.relu) v = fmaxf(v, 0.f);
        v *= rm;
        if (R) v += R[(long long)m * g.ldr + n];
        if (g.relumask && !(g.relumask[(long long)m * g.ldm + n] > 0.f)) v = 0.f;
        *cp = v;
      }
    }
  }
}

// Host-side launcher: picks the 128x128 (8x8 per thread) tile for big problems, 64x64 (4x4) otherwise.
inline cudaError_t launch_gemm(GemmArgs g, bool b_kmajor, cudaStream_t st, bool a_kmajor = true) {
  auto al4 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  g.vecA = al4(g.A) && g.lda % 4 == 0 && g.sA0 % 4 == 0 && g.sA1 % 4 == 0;
  g.vecB = al4(g.B) && g.ldb % 4 == 0 && g.sB0 % 4 == 0 && g.sB1 % 4 == 0;
  if (g.splits < 1) g.splits = 1;
  if (g.splits > 1 && !g.atomic) return cudaErrorInvalidValue;
  const int nbatch = g.nb0 * g.nb1 * g.splits;
  if (!a_kmajor) {   // weight-gradient form (A and B both read along their contiguous dimension)
    if (b_kmajor) return cudaErrorInvalidValue;
    const bool big = g.M >= 96 && g.N >= 96;
    if (big) {
      dim3 grid((g.N + 127) / 128, (g.M + 127) / 128, nbatch);
      gemm_kernel<128, 128, 8, 8, false, false><<<grid, 256, 0, st>>>(g);
    } else {
      dim3 grid((g.N + 63) / 64, (g.M + 63) / 64, nbatch);
      gemm_kernel<64, 64, 4, 4, false, false><<<grid, 256, 0, st>>>(g);
    }
    return cudaGetLastError();
  }
  const bool big = (long long)g.M * g.N >= 128LL * 128 * 64 && g.N >= 96 && g.M >= 128;
  if (big) {
    dim3 grid((g.N + 127) / 128, (g.M + 127) / 128, nbatch);
    if (b_kmajor) gemm_kernel<128, 128, 8, 8, true><<<grid, 256, 0, st>>>(g);
    else gemm_kernel<128, 128, 8, 8, false><<<grid, 256, 0, st>>>(g);
  } else {
    dim3 grid((g.N + 63) / 64, (g.M + 63) / 64, nbatch);
    if (b_kmajor) gemm_kernel<64, 64, 4, 4, true><<<grid, 256, 0, st>>>(g);
    else gemm_kernel<64, 64, 4, 4, false><<<grid, 256, 0, st>>>(g);
  }
  return cudaGetLastError();
}

}  // namespace fd
