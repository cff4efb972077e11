// Split-bf16 tensor-core GEMM on fp32 operands for the TRAINING path: C = epi(alpha * A·B) with the same argument block and
// epilogue as the CUDA-core gemm_kernel (fd_gemm.cuh), in all three operand forms a training step needs —
//   forward  y  = x W^T      A[m][k] (k contiguous)   B = W[n][k] (k contiguous)
//   dgrad    dx = dy W       A[m][k]                  B = W[k][n] (n contiguous)
//   wgrad    dW = dy^T x     A = dy[k][m] (m contig.) B = x[k][n]                     (split-K + atomicAdd)
// — so neither operand ever needs a transposed or pre-split copy in HBM: fp32 tiles are read once, split into bf16 hi/lo on the way
// into shared memory, and multiplied as hi·hi + hi·lo + lo·hi with mma.sync.m16n8k16 (fp32 accumulate): products keep ~2^-17 relative
// accuracy, the same 3-term scheme as the inference path's bf16x3 mode.  Transposed operand forms are fed to the MMA through
// ldmatrix.trans.  CTA tile 128x128x32, 8 warps (2 x 4, 64x32 per warp), register-prefetched double buffering.
// (The inference path's edge GEMMs use tcgen05 with TMA-staged bf16 planes, fd_tc.cuh; mma.sync is used here because all three
// operand forms, incl. the K = rows weight-gradient with split-K, come out of one kernel reading fp32 activations directly.)
#pragma once
#include "fd_gemm.cuh"

namespace fd {

constexpr int MM3_BM = 128, MM3_BN = 128, MM3_BK = 32;
constexpr int MM3_KSTRIDE = 80;     // bytes per smem row of a k-contiguous tile: 32 bf16 + 16 B padding
constexpr int MM3_MSTRIDE = 272;    // bytes per smem row of an m/n-contiguous tile: 128 bf16 + 16 B padding
constexpr int MM3_PLANE = 128 * MM3_KSTRIDE;                 // 10240 B (>= 32 * 272)
constexpr int MM3_STAGE = 4 * MM3_PLANE;                     // A hi | A lo | B hi | B lo
constexpr size_t MM3_SMEM = 2 * MM3_STAGE;

// MT = m16 tiles per warp: 4 -> 8 warps (2 x 4, 64x32 per warp, 256 threads); 2 -> 16 warps (4 x 4, 32x32 per warp, 512 threads: more warps in
// flight per SM, fewer registers per thread)
template <bool A_KMAJOR, bool B_KMAJOR, int MT = 4>
__global__ void __launch_bounds__(1024 / MT) mm3_kernel(const GemmArgs g) {
  constexpr int NTHR = 1024 / MT;          // 256 or 512
  constexpr int NLD = 1024 / NTHR;         // float4 loads per thread and operand: 4 or 2
  extern __shared__ __align__(128) uint8_t mm3_smem[];
  const uint32_t sbase = smem_u32(mm3_smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int zs = blockIdx.z / g.splits, split = blockIdx.z - zs * g.splits;
  const int z0 = zs / g.nb1, z1 = zs % g.nb1;
  const float* __restrict__ A = g.A + z0 * g.sA0 + z1 * g.sA1;
  const float* __restrict__ B = g.B + z0 * g.sB0 + z1 * g.sB1;
  float* __restrict__ C = g.C + z0 * g.sC0 + z1 * g.sC1;
  const int m0 = blockIdx.y * MM3_BM, n0 = blockIdx.x * MM3_BN;
  const int M = g.M, N = g.N;
  int kbeg = 0, K = g.K;
  if (g.splits > 1) {
    const int per = ((g.K + g.splits - 1) / g.splits + MM3_BK - 1) / MM3_BK * MM3_BK;
    kbeg = split * per;
    K = min(g.K, kbeg + per);
    if (kbeg >= K) return;
  }

  // bias gradient riding on the weight gradient: column sums of the A tiles (threads keep their (k row, m quad) slots across k-steps)
  const bool do_colsum = !A_KMAJOR && g.colsum_out != nullptr && blockIdx.x == 0;
  float4 csum[NLD];
#pragma unroll
  for (int it = 0; it < NLD; ++it) csum[it] = make_float4(0.f, 0.f, 0.f, 0.f);

  float4 ra[NLD], rb[NLD];
  // ---- global -> registers (fp32, zero-filled outside the problem) ----
  auto load_one = [&](const float* __restrict__ P, int ld, bool kmajor, int vec, int r0, int R, int k0, int idx) -> float4 {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kmajor) {                                 // tile [128 rows][32 k], 8 float4 per row
      const int row = idx >> 3, kq = (idx & 7) * 4;
      const int r = r0 + row, k = k0 + kq;
      if (r < R) {
        const float* p = P + (long long)r * ld + k;
        if (vec && k + 3 < K) v = *reinterpret_cast<const float4*>(p);
        else {
          if (k < K) v.x = p[0];
          if (k + 1 < K) v.y = p[1];
          if (k + 2 < K) v.z = p[2];
          if (k + 3 < K) v.w = p[3];
        }
      }
    } else {                                      // tile [32 k][128 rows], 32 float4 per k row
      const int kr = idx >> 5, rq = (idx & 31) * 4;
      const int k = k0 + kr, r = r0 + rq;
      if (k < K) {
        const float* p = P + (long long)k * ld + r;
        if (vec && r + 3 < R) v = *reinterpret_cast<const float4*>(p);
        else {
          if (r < R) v.x = p[0];
          if (r + 1 < R) v.y = p[1];
          if (r + 2 < R) v.z = p[2];
          if (r + 3 < R) v.w = p[3];
        }
      }
    }
    return v;
  };
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int it = 0; it < NLD; ++it) {
      ra[it] = load_one(A, g.lda, A_KMAJOR, g.vecA, m0, M, k0, tid + it * NTHR);
      rb[it] = load_one(B, g.ldb, B_KMAJOR, g.vecB, n0, N, k0, tid + it * NTHR);
    }
  };
  // ---- registers -> shared (bf16 hi / lo) ----
  auto store_one = [&](uint32_t hi_s, uint32_t lo_s, bool kmajor, int idx, const float4& v) {
    uint32_t h0, l0, h1, l1;
    split2_bf16(v.x, v.y, h0, l0);
    split2_bf16(v.z, v.w, h1, l1);
    uint32_t off;
    if (kmajor) { const int row = idx >> 3, kq = (idx & 7) * 4; off = (uint32_t)(row * MM3_KSTRIDE + kq * 2); }
    else { const int kr = idx >> 5, rq = (idx & 31) * 4; off = (uint32_t)(kr * MM3_MSTRIDE + rq * 2); }
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(hi_s + off), "r"(h0), "r"(h1) : "memory");
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(lo_s + off), "r"(l0), "r"(l1) : "memory");
  };
  auto store_tiles = [&](int buf) {
    const uint32_t st = sbase + buf * MM3_STAGE;
    if (do_colsum) {
#pragma unroll
      for (int it = 0; it < NLD; ++it) { csum[it].x += ra[it].x; csum[it].y += ra[it].y; csum[it].z += ra[it].z; csum[it].w += ra[it].w; }
    }
#pragma unroll
    for (int it = 0; it < NLD; ++it) {
      store_one(st, st + MM3_PLANE, A_KMAJOR, tid + it * NTHR, ra[it]);
      store_one(st + 2 * MM3_PLANE, st + 3 * MM3_PLANE, B_KMAJOR, tid + it * NTHR, rb[it]);
    }
  };

  float acc[MT][4][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.f; acc[i][j][1] = 0.f; acc[i][j][2] = 0.f; acc[i][j][3] = 0.f; }

  const int wm = (warp >> 2) * (16 * MT), wn = (warp & 3) * 32;       // warp tile origin inside the CTA tile
  const int mi = lane >> 3, lr = lane & 7;

  const int nk = (K - kbeg + MM3_BK - 1) / MM3_BK;
  load_tiles(kbeg);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles(kbeg + (kt + 1) * MM3_BK);
    const uint32_t st = sbase + buf * MM3_STAGE;
    const uint32_t a_hi = st, a_lo = st + MM3_PLANE, b_hi = st + 2 * MM3_PLANE, b_lo = st + 3 * MM3_PLANE;
#pragma unroll
    for (int ks = 0; ks < MM3_BK / 16; ++ks) {
      // B fragments of the warp's four n8 tiles: [nt][0..1] = (b0, b1)
      uint32_t bh[4][2], bl[4][2];
#pragma unroll
      for (int np = 0; np < 2; ++np) {           // n-tile pairs (2np, 2np+1)
        uint32_t r[4], s[4];
        if (B_KMAJOR) {                          // smem [n][k]: matrices (n 0-7,k lo) (n 0-7,k hi) (n 8-15,k lo) (n 8-15,k hi)
          const uint32_t off = (uint32_t)((wn + np * 16 + (mi >> 1) * 8 + lr) * MM3_KSTRIDE + (ks * 2 + (mi & 1)) * 16);
          ldsm_x4(b_hi + off, r); ldsm_x4(b_lo + off, s);
        } else {                                 // smem [k][n]: matrices (k lo,n 0-7) (k hi,n 0-7) (k lo,n 8-15) (k hi,n 8-15), transposed
          const uint32_t off = (uint32_t)((ks * 16 + (mi & 1) * 8 + lr) * MM3_MSTRIDE + (wn + np * 16 + (mi >> 1) * 8) * 2);
          ldsm_x4_t(b_hi + off, r); ldsm_x4_t(b_lo + off, s);
        }
        bh[2 * np][0] = r[0]; bh[2 * np][1] = r[1]; bh[2 * np + 1][0] = r[2]; bh[2 * np + 1][1] = r[3];
        bl[2 * np][0] = s[0]; bl[2 * np][1] = s[1]; bl[2 * np + 1][0] = s[2]; bl[2 * np + 1][1] = s[3];
      }
      uint32_t ah[MT][4], al[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        if (A_KMAJOR) {                          // smem [m][k]: matrices (m 0-7,k lo) (m 8-15,k lo) (m 0-7,k hi) (m 8-15,k hi)
          const uint32_t off = (uint32_t)((wm + mt * 16 + (mi & 1) * 8 + lr) * MM3_KSTRIDE + (ks * 2 + (mi >> 1)) * 16);
          ldsm_x4(a_hi + off, ah[mt]); ldsm_x4(a_lo + off, al[mt]);
        } else {                                 // smem [k][m]: matrices (k lo,m 0-7) (k lo,m 8-15) (k hi,m 0-7) (k hi,m 8-15), transposed
          const uint32_t off = (uint32_t)((ks * 16 + (mi >> 1) * 8 + lr) * MM3_MSTRIDE + (wm + mt * 16 + (mi & 1) * 8) * 2);
          ldsm_x4_t(a_hi + off, ah[mt]); ldsm_x4_t(a_lo + off, al[mt]);
        }
      }
      // three passes, each over ALL of the warp's accumulator tiles: consecutive MMAs never target the same accumulator, so the
      // dependent-issue latency of mma.sync (the top stall of the first version: "wait") is covered by the other tiles
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[mt][nt], ah[mt][0], ah[mt][1], ah[mt][2], ah[mt][3], bh[nt][0], bh[nt][1]);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[mt][nt], ah[mt][0], ah[mt][1], ah[mt][2], ah[mt][3], bl[nt][0], bl[nt][1]);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[mt][nt], al[mt][0], al[mt][1], al[mt][2], al[mt][3], bh[nt][0], bh[nt][1]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  if (do_colsum) {
    // slot it of thread tid covered tile rows k = (tid + 256 it) >> 5 and columns m0 + ((tid + 256 it) & 31) * 4 + {0..3}: (tid & 31) is the
    // same for all four slots, so a thread's four partial sums belong to the same 4 columns; 8 threads (tid >> 5) share them
    __syncthreads();
    float* red = reinterpret_cast<float*>(mm3_smem);            // [NTHR/32][128]
    float4 t4 = csum[0];
#pragma unroll
    for (int it = 1; it < NLD; ++it) { t4.x += csum[it].x; t4.y += csum[it].y; t4.z += csum[it].z; t4.w += csum[it].w; }
    *reinterpret_cast<float4*>(red + (tid >> 5) * 128 + (tid & 31) * 4) = t4;
    __syncthreads();
    if (tid < 128 && m0 + tid < M) {
      float sacc = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < NTHR / 32; ++w8) sacc += red[w8 * 128 + tid];
      atomicAdd(g.colsum_out + m0 + tid, sacc);
    }
  }

  // ---- epilogue (same order as gemm_kernel): +bias, +rowadd, +C, relu, *rowmask, +residual, relu-mask; or atomicAdd ----
  const float* R = g.residual ? g.residual + z0 * g.sR0 + z1 * g.sR1 : nullptr;
  const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int m = m0 + wm + mt * 16 + gq + half * 8;
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
      for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int n = n0 + wn + nt * 8 + tq * 2 + e;
          if (n >= N) continue;
          float v = g.alpha * acc[mt][nt][half * 2 + e];
          float* cp = C + (long long)m * g.ldc + n;
          if (g.atomic) { atomicAdd(cp, v); continue; }
          if (g.bias) v += g.bias[n];
          if (pi) v += pi[n] + pj[n];
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
}

static bool g_mm3_ready = false;
static int g_mm3_mt = 2;          // FD_MM3_MT=4 selects the 8-warp variant (kept for A/B measurements)
inline int mm3_init() {
  if (g_mm3_ready) return 0;
  if (const char* e = getenv("FD_MM3_MT")) g_mm3_mt = atoi(e) == 4 ? 4 : 2;
  cudaError_t e = cudaSuccess;
  e = cudaFuncSetAttribute(mm3_kernel<true, true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MM3_SMEM); if (e != cudaSuccess) return -1;
  e = cudaFuncSetAttribute(mm3_kernel<true, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MM3_SMEM); if (e != cudaSuccess) return -1;
  e = cudaFuncSetAttribute(mm3_kernel<false, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MM3_SMEM); if (e != cudaSuccess) return -1;
  e = cudaFuncSetAttribute(mm3_kernel<true, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MM3_SMEM); if (e != cudaSuccess) return -1;
  e = cudaFuncSetAttribute(mm3_kernel<true, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MM3_SMEM); if (e != cudaSuccess) return -1;
  e = cudaFuncSetAttribute(mm3_kernel<false, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MM3_SMEM); if (e != cudaSuccess) return -1;
  g_mm3_ready = true;
  return 0;
}

// Same contract as launch_gemm.  Small problems stay on the CUDA-core kernel (launch-latency bound anyway).
inline cudaError_t launch_gemm_mm3(GemmArgs g, bool b_kmajor, cudaStream_t st, bool a_kmajor = true) {
  const double work = 2.0 * g.M * g.N * (double)g.K * g.nb0 * g.nb1;
  // (weight-gradient form: even an 8-row output — d linear_b — is worth a tensor-core tile when K is the edge count)
  if (work < 3.0e7 || g.M < (a_kmajor ? 32 : 8) || g.N < 16 || (!a_kmajor && b_kmajor) || mm3_init()) return launch_gemm(g, b_kmajor, st, a_kmajor);
  auto al4 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  g.vecA = al4(g.A) && g.lda % 4 == 0 && g.sA0 % 4 == 0 && g.sA1 % 4 == 0;
  g.vecB = al4(g.B) && g.ldb % 4 == 0 && g.sB0 % 4 == 0 && g.sB1 % 4 == 0;
  if (g.splits < 1) g.splits = 1;
  if (g.splits > 1 && !g.atomic) return cudaErrorInvalidValue;
  dim3 grid((g.N + MM3_BN - 1) / MM3_BN, (g.M + MM3_BM - 1) / MM3_BM, g.nb0 * g.nb1 * g.splits);
  if (g_mm3_mt == 4) {
    if (a_kmajor && b_kmajor) mm3_kernel<true, true, 4><<<grid, 256, MM3_SMEM, st>>>(g);
    else if (a_kmajor) mm3_kernel<true, false, 4><<<grid, 256, MM3_SMEM, st>>>(g);
    else mm3_kernel<false, false, 4><<<grid, 256, MM3_SMEM, st>>>(g);
  } else {
    if (a_kmajor && b_kmajor) mm3_kernel<true, true, 2><<<grid, 512, MM3_SMEM, st>>>(g);
    else if (a_kmajor) mm3_kernel<true, false, 2><<<grid, 512, MM3_SMEM, st>>>(g);
    else mm3_kernel<false, false, 2><<<grid, 512, MM3_SMEM, st>>>(g);
  }
  return cudaGetLastError();
}

}  // namespace fd
