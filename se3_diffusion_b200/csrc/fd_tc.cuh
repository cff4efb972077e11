// tcgen05 tensor-core path of the edge-tensor MLPs (EdgeTransition = 87 % of the reference's FLOPs, edge embedder).
//
//   FD_PREC_BF16X3 : every fp32 operand x is carried as two bf16 planes  x ≈ hi + lo  (hi = bf16(x), lo = bf16(x − hi));
//                    a·b ≈ hi·hi + hi·lo + lo·hi  (three tcgen05.mma passes into the same fp32 TMEM accumulator) — relative
//                    error ~2^-16, which keeps the whole forward within 1e-4 of the fp32 reference;
//   FD_PREC_BF16   : hi planes only, one pass (throughput mode, ~1e-3).
//
// One generic kernel, tc_gemm_kernel:   C[M, N] = epilogue( A[M, K] · W[N, K]^T )
//   * persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocator), warps 2..5 = epilogue;
//   * A and W tiles arrive by TMA (cp.async.bulk.tensor, 128-byte swizzle) into two independent mbarrier rings
//     (A: 128 rows × 64 k; W: 128-row n-chunks × 64 k), operands are K-major, UMMA 128×128×16, accumulators in TMEM
//     (N ≤ 384 → ≤ 384 of the 512 columns);
//   * A may come from two tensors along K (EdgeTransition's last layer: [h2 | z]·[Wf | Wfz]^T);
//   * epilogue (TMEM → registers, thread = row): + bias, + row-broadcast node terms P_i/Q_j (or U_i/V_j), then either
//     ReLU → bf16 hi/lo planes, or LayerNorm(128) → × edge mask → bf16 hi/lo planes (the next layer's TMA source).
//
// Encodings follow CUTLASS cute/arch/mma_sm100_desc.hpp (InstrDescriptor, SmemDescriptor) and
// cute/atom/mma_traits_sm100.hpp::make_umma_desc<Major::K> for the 128B-swizzled K-major canonical layout.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <map>
#include <string>
#include <vector>
#include "fd_common.cuh"
#include "fd_kernels.cuh"

namespace fd {

// ------------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug traps (-> CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && spin > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled smem operand descriptor (rows of 64 bf16 = 128 B, 8-row swizzle atoms of 1024 B)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);          // start address  [0,14)
  d |= (uint64_t)1 << 16;                               // LBO (unused for swizzled K-major; CUTLASS writes 1)
  d |= (uint64_t)(1024 >> 4) << 32;                     // SBO = 1024 B between 8-row groups  [32,46)
  d |= (uint64_t)1 << 46;                               // descriptor version 1 (Blackwell)
  d |= (uint64_t)2 << 61;                               // layout type: SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128, N = 128
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------------------------
// GEMM kernel
// ------------------------------------------------------------------------------------------------------------------
constexpr int TC_BM = 128, TC_BK = 64, TC_NC = 128;           // tile rows, k per stage, n-chunk
constexpr int TC_PLANE_BYTES = TC_BM * TC_BK * 2;             // 16 KB: one bf16 plane of a 128×64 tile
constexpr int TC_SA = 2, TC_SB = 4;                           // ring depths
constexpr int TC_THREADS = 192;
constexpr int TC_TMEM_COLS = 512;

enum { TC_EPI_RELU = 0, TC_EPI_LN = 1 };

struct TcGemmParams {
  int M, N;                 // rows; output width (128 or 384)
  int KB0, KB1;             // 64-wide k-blocks taken from A0 (then A1); weight K = (KB0+KB1)*64
  int planes;               // 2 = hi+lo (3 MMA passes), 1 = hi only
  int epi;
  const float* bias;        // [N] or null
  const float* rowadd;      // node-term table [B*nres, ld_rowadd] or null; cols off_i.. for row i, off_j.. for row j
  int off_i, off_j, ld_rowadd, nres;
  const float* res_mask;    // [B*nres] (LN epilogue: edge mask m_i*m_j) or null
  const float* ln_g; const float* ln_b;
  __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;   // [M,N] planes
  int num_tiles;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mA0h, const __grid_constant__ CUtensorMap mA0l,
               const __grid_constant__ CUtensorMap mA1h, const __grid_constant__ CUtensorMap mA1l,
               const __grid_constant__ CUtensorMap mBh, const __grid_constant__ CUtensorMap mBl, const TcGemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: A ring | B ring | barriers
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_ring = base;
  const uint32_t b_ring = a_ring + TC_SA * 2 * TC_PLANE_BYTES;
  const uint32_t bar0 = b_ring + TC_SB * 2 * TC_PLANE_BYTES;
  // barrier slots (8 B each): a_full[SA], a_empty[SA], b_full[SB], b_empty[SB], tmem_full, tmem_empty, tmem_ptr
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (TC_SA + s); };
  auto b_full = [&](int s) { return bar0 + 8u * (2 * TC_SA + s); };
  auto b_empty = [&](int s) { return bar0 + 8u * (2 * TC_SA + TC_SB + s); };
  const uint32_t tmem_full = bar0 + 8u * (2 * TC_SA + 2 * TC_SB);
  const uint32_t tmem_empty = tmem_full + 8u;
  const uint32_t tmem_ptr_addr = tmem_empty + 8u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = p.KB0 + p.KB1;
  const int NCH = p.N / TC_NC;
  const uint32_t stage_bytes = (uint32_t)p.planes * TC_PLANE_BYTES;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_SA; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < TC_SB; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 4);     // one arrive per epilogue warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&mA0h); tma_prefetch_desc(&mBh);
    if (p.planes == 2) { tma_prefetch_desc(&mA0l); tma_prefetch_desc(&mBl); }
    if (p.KB1 > 0) { tma_prefetch_desc(&mA1h); if (p.planes == 2) tma_prefetch_desc(&mA1l); }
  }
  if (warp == 1) {   // TMEM allocation (whole warp), address lands in smem
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      uint32_t ia = 0, ib = 0;   // running stage counters
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int m0 = tile * TC_BM;
        for (int kb = 0; kb < KB; ++kb) {
          const uint32_t sa = ia % TC_SA, pa = (ia / TC_SA) & 1u;
          mbar_wait(a_empty(sa), pa ^ 1u);
          mbar_expect_tx(a_full(sa), stage_bytes);
          const bool first = kb < p.KB0;
          const int ka = (first ? kb : kb - p.KB0) * TC_BK;
          const uint32_t dstA = a_ring + sa * 2 * TC_PLANE_BYTES;
          tma_load_2d(dstA, first ? &mA0h : &mA1h, a_full(sa), ka, m0);
          if (p.planes == 2) tma_load_2d(dstA + TC_PLANE_BYTES, first ? &mA0l : &mA1l, a_full(sa), ka, m0);
          ++ia;
          for (int c = 0; c < NCH; ++c) {
            const uint32_t sb = ib % TC_SB, pb = (ib / TC_SB) & 1u;
            mbar_wait(b_empty(sb), pb ^ 1u);
            mbar_expect_tx(b_full(sb), stage_bytes);
            const uint32_t dstB = b_ring + sb * 2 * TC_PLANE_BYTES;
            tma_load_2d(dstB, &mBh, b_full(sb), kb * TC_BK, c * TC_NC);
            if (p.planes == 2) tma_load_2d(dstB + TC_PLANE_BYTES, &mBl, b_full(sb), kb * TC_BK, c * TC_NC);
            ++ib;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(TC_BM, TC_NC);
      uint32_t ia = 0, ib = 0, it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        mbar_wait(tmem_empty, (it & 1u) ^ 1u);      // epilogue has drained the previous tile's accumulators
        tc_fence_after();
        for (int kb = 0; kb < KB; ++kb) {
          const uint32_t sa = ia % TC_SA, pa = (ia / TC_SA) & 1u;
          mbar_wait(a_full(sa), pa);
          tc_fence_after();
          const uint32_t aH = a_ring + sa * 2 * TC_PLANE_BYTES, aL = aH + TC_PLANE_BYTES;
          for (int c = 0; c < NCH; ++c) {
            const uint32_t sb = ib % TC_SB, pb = (ib / TC_SB) & 1u;
            mbar_wait(b_full(sb), pb);
            tc_fence_after();
            const uint32_t bH = b_ring + sb * 2 * TC_PLANE_BYTES, bL = bH + TC_PLANE_BYTES;
            const uint32_t d = tmem_base + (uint32_t)(c * TC_NC);
#pragma unroll
            for (int ks = 0; ks < TC_BK / 16; ++ks) {
              const uint32_t ko = ks * 32;   // 16 bf16 = 32 bytes along K inside the swizzle atom
              const uint32_t acc0 = (kb > 0 || ks > 0) ? 1u : 0u;
              umma_bf16(d, make_sw128_desc(aH + ko), make_sw128_desc(bH + ko), idesc, acc0);
              if (p.planes == 2) {
                umma_bf16(d, make_sw128_desc(aH + ko), make_sw128_desc(bL + ko), idesc, 1u);
                umma_bf16(d, make_sw128_desc(aL + ko), make_sw128_desc(bH + ko), idesc, 1u);
              }
            }
            tc_commit(b_empty(sb));     // frees this weight stage once the MMAs above retire
            ++ib;
          }
          tc_commit(a_empty(sa));
          ++ia;
        }
        tc_commit(tmem_full);           // accumulators complete -> epilogue
      }
    }
  } else {
    // ================================ epilogue (warps 2..5) ================================
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
    const int row_in_tile = quad * 32 + lane;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      mbar_wait(tmem_full, it & 1u);
      tc_fence_after();
      const long long m = (long long)tile * TC_BM + row_in_tile;
      const bool valid = m < p.M;
      const float* add_i = nullptr; const float* add_j = nullptr;
      float emask = 1.f;
      if (valid && (p.rowadd || p.res_mask)) {
        const long long nn = (long long)p.nres * p.nres;
        const long long b = m / nn;
        const int rem = (int)(m - b * nn);
        const int ri = rem / p.nres, rj = rem - ri * p.nres;
        if (p.rowadd) {
          add_i = p.rowadd + (b * p.nres + ri) * p.ld_rowadd + p.off_i;
          add_j = p.rowadd + (b * p.nres + rj) * p.ld_rowadd + p.off_j;
        }
        if (p.res_mask) emask = p.res_mask[b * p.nres + ri] * p.res_mask[b * p.nres + rj];
      }
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
      if (p.epi == TC_EPI_RELU) {
        for (int c0 = 0; c0 < p.N; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(trow + (uint32_t)c0, r);
          if (valid) {
            __align__(16) __nv_bfloat16 hi[32];
            __align__(16) __nv_bfloat16 lo[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float4 bi = p.bias ? *reinterpret_cast<const float4*>(p.bias + c0 + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
              if (add_i) {
                const float4 x = *reinterpret_cast<const float4*>(add_i + c0 + q * 4);
                const float4 y = *reinterpret_cast<const float4*>(add_j + c0 + q * 4);
                bi.x += x.x + y.x; bi.y += x.y + y.y; bi.z += x.z + y.z; bi.w += x.w + y.w;
              }
              const float v0 = fmaxf(__uint_as_float(r[q * 4 + 0]) + bi.x, 0.f), v1 = fmaxf(__uint_as_float(r[q * 4 + 1]) + bi.y, 0.f);
              const float v2 = fmaxf(__uint_as_float(r[q * 4 + 2]) + bi.z, 0.f), v3 = fmaxf(__uint_as_float(r[q * 4 + 3]) + bi.w, 0.f);
              split_bf16(v0, hi[q * 4 + 0], lo[q * 4 + 0]); split_bf16(v1, hi[q * 4 + 1], lo[q * 4 + 1]);
              split_bf16(v2, hi[q * 4 + 2], lo[q * 4 + 2]); split_bf16(v3, hi[q * 4 + 3], lo[q * 4 + 3]);
            }
            uint4* oh = reinterpret_cast<uint4*>(p.out_hi + m * p.N + c0);
            const uint4* sh = reinterpret_cast<const uint4*>(hi);
            oh[0] = sh[0]; oh[1] = sh[1]; oh[2] = sh[2]; oh[3] = sh[3];
            if (p.planes == 2) {
              uint4* ol = reinterpret_cast<uint4*>(p.out_lo + m * p.N + c0);
              const uint4* sl = reinterpret_cast<const uint4*>(lo);
              ol[0] = sl[0]; ol[1] = sl[1]; ol[2] = sl[2]; ol[3] = sl[3];
            }
          }
        }
      } else {   // LayerNorm over the 128 columns of this row, then edge mask
        float v[128];
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(trow + (uint32_t)c0, r);
#pragma unroll
          for (int q = 0; q < 32; ++q) v[c0 + q] = __uint_as_float(r[q]);
        }
        if (valid) {
          float s = 0.f;
#pragma unroll
          for (int c = 0; c < 128; c += 4) {
            float4 bi = p.bias ? *reinterpret_cast<const float4*>(p.bias + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (add_i) {
              const float4 x = *reinterpret_cast<const float4*>(add_i + c);
              const float4 y = *reinterpret_cast<const float4*>(add_j + c);
              bi.x += x.x + y.x; bi.y += x.y + y.y; bi.z += x.z + y.z; bi.w += x.w + y.w;
            }
            v[c] += bi.x; v[c + 1] += bi.y; v[c + 2] += bi.z; v[c + 3] += bi.w;
            s += (v[c] + v[c + 1]) + (v[c + 2] + v[c + 3]);
          }
          const float mean = s * (1.f / 128.f);
          float q2 = 0.f;
#pragma unroll
          for (int c = 0; c < 128; ++c) { const float d = v[c] - mean; q2 = fmaf(d, d, q2); }
          const float rstd = rsqrtf(q2 * (1.f / 128.f) + 1e-5f);
#pragma unroll
          for (int c0 = 0; c0 < 128; c0 += 8) {
            __align__(16) __nv_bfloat16 hi[8];
            __align__(16) __nv_bfloat16 lo[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float y = ((v[c0 + q] - mean) * rstd * p.ln_g[c0 + q] + p.ln_b[c0 + q]) * emask;
              split_bf16(y, hi[q], lo[q]);
            }
            *reinterpret_cast<uint4*>(p.out_hi + m * 128 + c0) = *reinterpret_cast<const uint4*>(hi);
            if (p.planes == 2) *reinterpret_cast<uint4*>(p.out_lo + m * 128 + c0) = *reinterpret_cast<const uint4*>(lo);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
}

constexpr size_t TC_SMEM_BYTES = 1024 + (size_t)(TC_SA + TC_SB) * 2 * TC_PLANE_BYTES + 256;

// planes -> fp32 (debug taps / export)
__global__ void planes_to_f32_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, float* __restrict__ out,
                                     long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __bfloat162float(hi[i]) + (lo ? __bfloat162float(lo[i]) : 0.f);
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;
static int g_tc_sms = 148;

inline int tc_init(int sm_count) {
  g_tc_sms = sm_count;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return -2;
  g_encode = (PFN_encodeTiled)fn;
  if (cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES) != cudaSuccess) return -2;
  return 0;
}

// 2-D bf16 row-major [rows, cols] tensor map with a {64 cols, 128 rows} box and 128B swizzle
inline int tc_make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)TC_BM};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

struct TcMat {   // a bf16 hi/lo weight image [rows, cols] + maps
  __nv_bfloat16* hi = nullptr; __nv_bfloat16* lo = nullptr;
  CUtensorMap mh, ml;
  int rows = 0, cols = 0;
};

struct TcWeights {
  bool ready = false;
  char* arena = nullptr;
  TcMat ee2, ee4;                 // edge embedder layers 2 and 4: [128][128]
  TcMat w1z[3], w2[3], wf[3];     // EdgeTransition: [384][128], [384][384], [128][512] = [Wf | Wf[:, :128]]
};

struct TcWorkspace {
  char* base = nullptr;
  long long E = 0;
  __nv_bfloat16 *z_hi, *z_lo, *h1_hi, *h1_lo, *h2_hi, *h2_lo;
  CUtensorMap m_z_h, m_z_l, m_h1_h, m_h1_l, m_h2_h, m_h2_l;      // K = 128 / 384 / 384
  CUtensorMap m_e0_h, m_e0_l, m_e1_h, m_e1_l;                     // embedder staging viewed as [E,128] inside h1 / h2
};

inline void tc_free_weights(TcWeights& w) {
  if (w.arena) cudaFree(w.arena);
  w = TcWeights();
}

static inline uint16_t f2bf(float f) {   // round-to-nearest-even, like __float2bfloat16_rn
  uint32_t u; memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  const uint32_t r = 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)((u + r) >> 16);
}
static inline float bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }

inline int tc_pack_weights(TcWeights& tw, const std::map<std::string, const float*>& M, cudaStream_t st) {
  tc_free_weights(tw);
  struct Item { TcMat* m; std::vector<float> w; int rows, cols; };
  std::vector<Item> items;
  auto add = [&](TcMat* m, const float* src, int rows, int cols) { items.push_back({m, std::vector<float>(src, src + (size_t)rows * cols), rows, cols}); };
  add(&tw.ee2, M.at("embedding_layer.edge_embedder.2.weight"), 128, 128);
  add(&tw.ee4, M.at("embedding_layer.edge_embedder.4.weight"), 128, 128);
  for (int b = 0; b < 3; ++b) {
    const std::string p = "score_model.trunk.edge_transition_" + std::to_string(b) + ".";
    const float* w1 = M.at(p + "trunk.0.weight");      // [384][384]
    const float* wf = M.at(p + "final_layer.weight");  // [128][384]
    std::vector<float> w1z((size_t)ET_HID * C_Z), wfc((size_t)C_Z * (ET_HID + C_Z));
    for (int o = 0; o < ET_HID; ++o) memcpy(&w1z[(size_t)o * C_Z], w1 + (size_t)o * ET_HID, C_Z * sizeof(float));
    for (int o = 0; o < C_Z; ++o) {
      memcpy(&wfc[(size_t)o * (ET_HID + C_Z)], wf + (size_t)o * ET_HID, ET_HID * sizeof(float));           // applied to h2
      memcpy(&wfc[(size_t)o * (ET_HID + C_Z) + ET_HID], wf + (size_t)o * ET_HID, C_Z * sizeof(float));      // applied to z
    }
    items.push_back({&tw.w1z[b], w1z, ET_HID, C_Z});
    add(&tw.w2[b], M.at(p + "trunk.2.weight"), ET_HID, ET_HID);
    items.push_back({&tw.wf[b], wfc, C_Z, ET_HID + C_Z});
  }
  size_t total = 0;
  for (auto& it : items) total += 2 * (((size_t)it.rows * it.cols * 2 + 1023) & ~(size_t)1023);
  if (cudaMalloc(&tw.arena, total) != cudaSuccess) return -3;
  std::vector<uint16_t> host(total / 2);
  size_t off = 0;
  for (auto& it : items) {
    const size_t n = (size_t)it.rows * it.cols, padded = ((n * 2 + 1023) & ~(size_t)1023) / 2;
    uint16_t* hh = host.data() + off; uint16_t* hl = hh + padded;
    for (size_t i = 0; i < n; ++i) { hh[i] = f2bf(it.w[i]); hl[i] = f2bf(it.w[i] - bf2f(hh[i])); }
    it.m->hi = reinterpret_cast<__nv_bfloat16*>(tw.arena) + off;
    it.m->lo = it.m->hi + padded;
    it.m->rows = it.rows; it.m->cols = it.cols;
    off += 2 * padded;
  }
  if (cudaMemcpyAsync(tw.arena, host.data(), total, cudaMemcpyHostToDevice, st) != cudaSuccess) return -2;
  if (cudaStreamSynchronize(st) != cudaSuccess) return -2;
  for (auto& it : items) {
    if (tc_make_map(&it.m->mh, it.m->hi, it.rows, it.cols)) return -2;
    if (tc_make_map(&it.m->ml, it.m->lo, it.rows, it.cols)) return -2;
  }
  tw.ready = true;
  return 0;
}

inline size_t tc_workspace_bytes(int B, int N) {
  const size_t E = (size_t)B * N * N;
  auto al = [](size_t x) { return (x + 1023) & ~(size_t)1023; };
  return 2 * al(E * C_Z * 2) + 4 * al(E * ET_HID * 2) + 1024;
}
inline int tc_bind_workspace(TcWorkspace& w, char* p, int B, int N) {
  const size_t E = (size_t)B * N * N;
  auto al = [](size_t x) { return (x + 1023) & ~(size_t)1023; };
  p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023);
  w.base = p; w.E = (long long)E;
  w.z_hi = (__nv_bfloat16*)p; p += al(E * C_Z * 2);
  w.z_lo = (__nv_bfloat16*)p; p += al(E * C_Z * 2);
  w.h1_hi = (__nv_bfloat16*)p; p += al(E * ET_HID * 2);
  w.h1_lo = (__nv_bfloat16*)p; p += al(E * ET_HID * 2);
  w.h2_hi = (__nv_bfloat16*)p; p += al(E * ET_HID * 2);
  w.h2_lo = (__nv_bfloat16*)p; p += al(E * ET_HID * 2);
  int rc = 0;
  rc |= tc_make_map(&w.m_z_h, w.z_hi, E, C_Z); rc |= tc_make_map(&w.m_z_l, w.z_lo, E, C_Z);
  rc |= tc_make_map(&w.m_h1_h, w.h1_hi, E, ET_HID); rc |= tc_make_map(&w.m_h1_l, w.h1_lo, E, ET_HID);
  rc |= tc_make_map(&w.m_h2_h, w.h2_hi, E, ET_HID); rc |= tc_make_map(&w.m_h2_l, w.h2_lo, E, ET_HID);
  rc |= tc_make_map(&w.m_e0_h, w.h1_hi, E, C_Z); rc |= tc_make_map(&w.m_e0_l, w.h1_lo, E, C_Z);
  rc |= tc_make_map(&w.m_e1_h, w.h2_hi, E, C_Z); rc |= tc_make_map(&w.m_e1_l, w.h2_lo, E, C_Z);
  return rc;
}

inline int tc_launch(const CUtensorMap& a0h, const CUtensorMap& a0l, const CUtensorMap& a1h, const CUtensorMap& a1l, const TcMat& Wt,
                     TcGemmParams p, cudaStream_t st, long long* launches) {
  p.num_tiles = (p.M + TC_BM - 1) / TC_BM;
  const int grid = p.num_tiles < g_tc_sms ? p.num_tiles : g_tc_sms;
  tc_gemm_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(a0h, a0l, a1h, a1l, Wt.mh, Wt.ml, p);
  if (launches) ++*launches;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// Edge embedder (model/score_network.py:79-86): layer 0 = table lookup kernel -> bf16 planes; layers 2, 4 on tensor cores.
inline int tc_edge_embed(const TcWeights& tw, TcWorkspace& w, int prec, const float* AC, const float* T, const float* D, const float* w0r,
                         const int* seq_idx, const float* sc_ca, const float* res_mask, const float* b2, const float* b4,
                         const float* ln_g, const float* ln_b, int B, int N, cudaStream_t st, long long* launches) {
  const long long E = w.E;
  const int planes = prec == 1 ? 2 : 1;
  if (E > 0x7fffffffLL) return -1;
  if (planes == 2) edge_embed_l0_kernel<2><<<(unsigned)((E + 7) / 8), 256, 0, st>>>(AC, T, D, w0r, seq_idx, sc_ca, nullptr, w.h1_hi, w.h1_lo, 0, E, N);
  else edge_embed_l0_kernel<1><<<(unsigned)((E + 7) / 8), 256, 0, st>>>(AC, T, D, w0r, seq_idx, sc_ca, nullptr, w.h1_hi, w.h1_lo, 0, E, N);
  if (launches) ++*launches;
  TcGemmParams p{};
  p.M = (int)E; p.N = 128; p.KB0 = 2; p.KB1 = 0; p.planes = planes; p.epi = TC_EPI_RELU; p.bias = b2; p.nres = N;
  p.out_hi = w.h2_hi; p.out_lo = w.h2_lo;
  if (tc_launch(w.m_e0_h, w.m_e0_l, w.m_e0_h, w.m_e0_l, tw.ee2, p, st, launches)) return -2;
  TcGemmParams q{};
  q.M = (int)E; q.N = 128; q.KB0 = 2; q.KB1 = 0; q.planes = planes; q.epi = TC_EPI_LN; q.bias = b4; q.nres = N; q.res_mask = res_mask;
  q.ln_g = ln_g; q.ln_b = ln_b; q.out_hi = w.z_hi; q.out_lo = w.z_lo;
  if (tc_launch(w.m_e1_h, w.m_e1_l, w.m_e1_h, w.m_e1_l, tw.ee4, q, st, launches)) return -2;
  return 0;
}

// EdgeTransition (model/ipa_pytorch.py:218-233) with the separable first/last layers (node terms P,Q,U,V precomputed):
//   h1 = relu(z·W1z^T + P_i + Q_j);  h2 = relu(h1·W2^T + b2);  z' = LN([h2|z]·[Wf|Wfz]^T + U_i + V_j)·mask
inline int tc_edge_transition(const TcWeights& tw, TcWorkspace& w, int blk, int prec, const float* pquv, const float* b2, const float* ln_g,
                              const float* ln_b, const float* res_mask, int B, int N, cudaStream_t st, long long* launches) {
  const long long E = w.E;
  const int planes = prec == 1 ? 2 : 1;
  if (E > 0x7fffffffLL) return -1;
  TcGemmParams p{};
  p.M = (int)E; p.N = ET_HID; p.KB0 = 2; p.KB1 = 0; p.planes = planes; p.epi = TC_EPI_RELU; p.rowadd = pquv; p.off_i = 0; p.off_j = ET_HID;
  p.ld_rowadd = ET_NODE; p.nres = N; p.out_hi = w.h1_hi; p.out_lo = w.h1_lo;
  if (tc_launch(w.m_z_h, w.m_z_l, w.m_z_h, w.m_z_l, tw.w1z[blk], p, st, launches)) return -2;
  TcGemmParams q{};
  q.M = (int)E; q.N = ET_HID; q.KB0 = 6; q.KB1 = 0; q.planes = planes; q.epi = TC_EPI_RELU; q.bias = b2; q.nres = N;
  q.out_hi = w.h2_hi; q.out_lo = w.h2_lo;
  if (tc_launch(w.m_h1_h, w.m_h1_l, w.m_h1_h, w.m_h1_l, tw.w2[blk], q, st, launches)) return -2;
  TcGemmParams r{};
  r.M = (int)E; r.N = C_Z; r.KB0 = 6; r.KB1 = 2; r.planes = planes; r.epi = TC_EPI_LN; r.rowadd = pquv; r.off_i = 2 * ET_HID;
  r.off_j = 2 * ET_HID + C_Z; r.ld_rowadd = ET_NODE; r.nres = N; r.res_mask = res_mask; r.ln_g = ln_g; r.ln_b = ln_b;
  r.out_hi = w.z_hi; r.out_lo = w.z_lo;
  if (tc_launch(w.m_h2_h, w.m_h2_l, w.m_z_h, w.m_z_l, tw.wf[blk], r, st, launches)) return -2;
  return 0;
}

inline void tc_export_z(TcWorkspace& w, float* z_f32, int prec, cudaStream_t st) {
  const long long n = w.E * C_Z;
  planes_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w.z_hi, prec == 1 ? w.z_lo : nullptr, z_f32, n);
}

inline int tc_ipa_edge(TcWorkspace& w, float* L, const float* qp, const float* kp, const float* res_mask, const float* Wb, const float* bb,
                       const float* gamma, const float* WdT, const float* bd, float* feats, int B, int N, int Np, int prec,
                       cudaStream_t st, long long* launches) {
  const size_t smem = (size_t)(H * Np + H * PQ * 3 + 2 * H * C_Z) * sizeof(float);
  ZRef z; z.hi = w.z_hi; z.lo = w.z_lo;
  if (prec == 1) ipa_edge_kernel<2><<<dim3(N, B), 256, smem, st>>>(z, L, qp, kp, res_mask, Wb, bb, gamma, WdT, bd, feats, N, Np);
  else ipa_edge_kernel<1><<<dim3(N, B), 256, smem, st>>>(z, L, qp, kp, res_mask, Wb, bb, gamma, WdT, bd, feats, N, Np);
  if (launches) ++*launches;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace fd
