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
#include <stdlib.h>
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
// ---- thread-block-cluster helpers (2-CTA weight multicast of the fused EdgeTransition kernel) ----
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(bar), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {   // barrier with remote arrivals: cluster-scope acquire
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && spin > (1u << 22)) __trap();
  }
}
// TMA tile load delivered to the same smem offset (and signalling the same mbarrier offset) in every CTA of `mask`
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// (v0, v1) -> packed bf16x2 hi and lo words (hi = rn(v), lo = rn(v - hi)); low half = first element
__device__ __forceinline__ void split2_bf16(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float r0 = v0 - __uint_as_float(hi << 16), r1 = v1 - __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(r0, r1);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// 256-bit read-only global load (LDG.E.256): halves the LSU instruction / wavefront count of per-row scattered reads
__device__ __forceinline__ void ldg256(const float* p, float (&v)[8]) {
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
// one elected lane of a converged warp (CUTLASS elect_one_sync): operands stay warp-uniform, only the issue is predicated
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred)::"memory");
  return pred != 0;
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

__device__ __forceinline__ void tc_commit_elect(uint32_t bar) {
  if (elect_one()) tc_commit(bar);
  __syncwarp();
}
// One k-step (K = 16) of the split product into accumulator d:  hi·hi (+ hi·lo + lo·hi when three_pass).
// acc0 = 0 makes the first MMA overwrite d.  Descriptors are 64-bit values; advancing 32 B along K adds 2.
__device__ __forceinline__ void umma_kstep(uint32_t d, uint64_t aH, uint64_t aL, uint64_t bH, uint64_t bL, uint32_t idesc, uint32_t acc0,
                                           bool three_pass) {
  if (three_pass) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.eq.b32 q, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %3, %5, p;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %4, %5, q;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %2, %3, %5, q;\n\t}" ::"r"(d),
        "l"(aH), "l"(aL), "l"(bH), "l"(bL), "r"(idesc), "r"(acc0)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
        "l"(aH), "l"(bH), "r"(idesc), "r"(acc0)
        : "memory");
  }
}

// optional 4th term lo·lo (used by the small pair-bias GEMM, where MMA time is free and the logits want the extra bits)
__device__ __forceinline__ void umma_lolo(uint32_t d, uint64_t aL, uint64_t bL, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.eq.b32 q, 0, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n\t}" ::"r"(d),
      "l"(aL), "l"(bL), "r"(idesc)
      : "memory");
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
// MN-major SWIZZLE_128B operand (cute: Swizzle<3,4,3> o ((8,n),(8,k)) : ((1,LBO),(8,SBO)) in 16-byte units): a K row holds 64 contiguous MN
// elements (128 B); the next 64 MN elements live LBO bytes further (here: the second TMA box, 8 KB), the next group of 8 K rows SBO = 1 KB further
__device__ __forceinline__ uint64_t make_sw128_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)(8192 >> 4) << 16;                     // LBO
  d |= (uint64_t)(1024 >> 4) << 32;                     // SBO
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------------------------
// GEMM kernel
// ------------------------------------------------------------------------------------------------------------------
constexpr int TC_BM = 128, TC_BK = 64, TC_NC = 128;           // tile rows, k per stage, n-chunk
constexpr int TC_PLANE_BYTES = TC_BM * TC_BK * 2;             // 16 KB: one bf16 plane of a 128×64 tile
constexpr int TC_SA = 2, TC_SB = 4;                           // ring depths
constexpr int TC_THREADS = 192;                                // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue (EG = 2: warps 2..9)
constexpr int TC_TMEM_COLS = 512;

enum { TC_EPI_RELU = 0, TC_EPI_LN = 1, TC_EPI_F32 = 2 };

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
  // node-path linears (TC_EPI_F32): a work item is (row tile, n-group of `nch` 128-column chunks); fp32 output
  int m_tiles, nch;         // row tiles; chunks per work item (edge layers: nch = N/128, one group)
  int n_valid;              // true output width (weight rows are padded to a multiple of 128)
  float* out_f32; int ldo;
  const float* residual; int ldr;
  const float* rowmask;     // [M]
  int relu;
  int lolo;                 // also accumulate a_lo·b_lo (4-term product)
  int mma_n;                // UMMA N (128, or 16 for the 8-wide pair-bias GEMM: only the first rows of the weight tile are multiplied)
  // batched mode (attention GEMMs, TC_EPI_F32 only): tile -> (batch, local tile); batch -> (outer, inner) = (bat / bat_inner, bat % bat_inner).
  // Both operands are activation planes addressed through 2-D maps with per-batch row / k offsets (TMA coordinates).
  int bat_inner, bat_tiles;           // bat_inner == 0: unbatched
  int a_row_s0, a_row_s1, a_k_s1;     // A coordinate offsets: row += outer*s0 + inner*s1, k += inner*k_s1
  int b_row_s0, b_row_s1, b_k0, b_k_s1;
  long long o_s0, o_s1;               // output element offsets per (outer, inner)
  float alpha;                        // accumulator scale (0 = 1)
  // training path (TC_EPI_F32 only): zero the output where relumask[m][n] <= 0 (backward of a ReLU whose OUTPUT is relumask); `rowadd`
  // (edge-row node terms) is honoured by the fp32 epilogue as well
  const float* relumask; int ldm;
  // TC_EPI_RELU extension (training backward): the planes written are acc (+bias) WITHOUT the ReLU where maskplane[m][n] != 0 and 0 elsewhere —
  // d(pre-activation) = d(post-activation) * [activation > 0], the activation given by its bf16 hi plane ([M, N], same layout as the output)
  const __nv_bfloat16* maskplane;
  int mn_major;                       // both operands are MN-major: planes [K rows, MN columns] (an activation tensor [rows, C] contracted over its
                                      // ROWS — the weight gradient dW = dy^T x — read in place, no transposed copy): TMA boxes {64 MN, 64 K}, two per
                                      // 128-wide tile, UMMA descriptors with LBO = 8 KB (next 64-column block) and SBO = 1 KB (next 8 K rows)
  int atomic;                         // out_f32[m][n] += alpha*acc through atomicAdd (weight gradients: the k-slices of one output tile are
                                      // the batched mode's inner batches with o_s1 = 0); bias / relu / mask / residual are ignored
  int eg;                             // 2: ten-warp variant (two epilogue groups) — pays on edge-sized GEMMs with wide outputs
  int sa;                             // stages of the activation ring (0 = 2; the weight ring takes 6 - sa)
  int chunk_minor;                    // > 0: work item lt -> (row tile lt / chunk_minor, n-group lt % chunk_minor): the n-groups of one row tile run
                                      // on neighbouring CTAs at the same time, so the A tile comes from DRAM once and from L2 afterwards
                                      // (0: lt -> (lt % m_tiles, lt / m_tiles))
};

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]);
__device__ __forceinline__ void tmem_ld_wait();

// Coalesced plane I/O for the plane-writing epilogues: a warp owns 32 consecutive rows, each thread one row.  64 bf16 columns (128 B) of
// every row go through a 4 KB XOR-swizzled shared-memory tile so that each global instruction moves four whole 128-byte row segments
// (row-per-thread 16-byte accesses touch 32 different rows per instruction: half-used sectors, 8x the LSU wavefronts).  `ldn` = row stride.
__device__ __forceinline__ void warp_store_rows64_ld(uint32_t stage, const uint32_t (&w)[32], __nv_bfloat16* out, long long m_warp, int col0,
                                                     long long M, int lane, int ldn) {
#pragma unroll
  for (int u = 0; u < 8; ++u)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + (uint32_t)(lane * 128 + ((u ^ (lane & 7)) << 4))), "r"(w[4 * u]),
                 "r"(w[4 * u + 1]), "r"(w[4 * u + 2]), "r"(w[4 * u + 3])
                 : "memory");
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int r = k * 4 + (lane >> 3), u = lane & 7;
    uint4 val;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                 : "r"(stage + (uint32_t)(r * 128 + ((u ^ (r & 7)) << 4))) : "memory");
    if (m_warp + r < M) *reinterpret_cast<uint4*>(out + (m_warp + r) * ldn + col0 + u * 8) = val;
  }
  __syncwarp();
}
// fp32 variant: 32 fp32 columns (128 B) of every row; `out` already points at the first of those columns of row 0
__device__ __forceinline__ void warp_store_rows32f(uint32_t stage, const uint32_t (&w)[32], float* out, long long m_warp, long long M, int lane,
                                                   long long ldo) {
#pragma unroll
  for (int u = 0; u < 8; ++u)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + (uint32_t)(lane * 128 + ((u ^ (lane & 7)) << 4))), "r"(w[4 * u]),
                 "r"(w[4 * u + 1]), "r"(w[4 * u + 2]), "r"(w[4 * u + 3])
                 : "memory");
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int r = k * 4 + (lane >> 3), u = lane & 7;
    uint4 val;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                 : "r"(stage + (uint32_t)(r * 128 + ((u ^ (r & 7)) << 4))) : "memory");
    if (m_warp + r < M) *reinterpret_cast<uint4*>(out + (m_warp + r) * ldo + u * 4) = val;
  }
  __syncwarp();
}
// the inverse: 64 bf16 columns of the warp's 32 rows, global -> (swizzled smem) -> each thread's own row in registers
__device__ __forceinline__ void warp_load_rows64_ld(uint32_t stage, uint32_t (&w)[32], const __nv_bfloat16* in, long long m_warp, int col0,
                                                    long long M, int lane, int ldn) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int r = k * 4 + (lane >> 3), u = lane & 7;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (m_warp + r < M) val = *reinterpret_cast<const uint4*>(in + (m_warp + r) * ldn + col0 + u * 8);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + (uint32_t)(r * 128 + ((u ^ (r & 7)) << 4))), "r"(val.x), "r"(val.y),
                 "r"(val.z), "r"(val.w)
                 : "memory");
  }
  __syncwarp();
#pragma unroll
  for (int u = 0; u < 8; ++u)
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[4 * u]), "=r"(w[4 * u + 1]), "=r"(w[4 * u + 2]), "=r"(w[4 * u + 3])
                 : "r"(stage + (uint32_t)(lane * 128 + ((u ^ (lane & 7)) << 4))) : "memory");
  __syncwarp();
}

// Row gather for the node terms of edge rows: thread `lane` wants 32 fp32 (128 B) starting at its own pointer `mine` (null = zeros).  Read
// row-per-thread, one instruction touches 32 different lines 16 bytes at a time (every access an L2 round trip, the table rows are 4 KB apart);
// here eight lanes fetch one row's 128 bytes, four rows per instruction, and the tile is transposed through the warp's staging buffer.
__device__ __forceinline__ void warp_gather_rows32f(uint32_t stage, uint32_t (&w)[32], const float* mine, int lane) {
  const unsigned long long pm = reinterpret_cast<unsigned long long>(mine);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int r = k * 4 + (lane >> 3), u = lane & 7;
    const unsigned long long pr = __shfl_sync(0xffffffffu, pm, r);
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (pr) val = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(pr) + u * 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + (uint32_t)(r * 128 + ((u ^ (r & 7)) << 4))), "r"(val.x), "r"(val.y),
                 "r"(val.z), "r"(val.w)
                 : "memory");
  }
  __syncwarp();
#pragma unroll
  for (int u = 0; u < 8; ++u)
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[4 * u]), "=r"(w[4 * u + 1]), "=r"(w[4 * u + 2]), "=r"(w[4 * u + 3])
                 : "r"(stage + (uint32_t)(lane * 128 + ((u ^ (lane & 7)) << 4))) : "memory");
  __syncwarp();
}

// EG = epilogue groups: 1 -> 192 threads (up to 255 registers), 2 -> 320 threads (ten warps: three on one scheduler, so 168 registers)
template <int EG>
__global__ void __launch_bounds__(64 + 128 * EG, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mA0h, const __grid_constant__ CUtensorMap mA0l,
               const __grid_constant__ CUtensorMap mA1h, const __grid_constant__ CUtensorMap mA1l,
               const __grid_constant__ CUtensorMap mBh, const __grid_constant__ CUtensorMap mBl, const TcGemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: A ring | B ring | barriers
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_ring = base;
  // the six 32 KB stages are split between the two rings per launch: SA for the activation operand, the rest for the weights.  A GEMM whose A
  // streams from DRAM (edge-sized M) is bound by (A stages in flight) / (DRAM latency): it takes four; the node-path GEMMs, whose A sits in L2 and
  // which need up to three weight chunks per A block, keep 2 + 4
  constexpr uint32_t TC_S = TC_SA + TC_SB;
  const uint32_t SA = p.sa > 0 ? (uint32_t)p.sa : (uint32_t)TC_SA, SB = TC_S - SA;
  const uint32_t b_ring = a_ring + SA * 2 * TC_PLANE_BYTES;
  const uint32_t bar0 = a_ring + TC_S * 2 * TC_PLANE_BYTES;
  // barrier slots (8 B each): a_full[6], a_empty[6], b_full[6], b_empty[6], tmem_full[2], tmem_empty[2], tmem_ptr
  auto a_full = [&](uint32_t s) { return bar0 + 8u * s; };
  auto a_empty = [&](uint32_t s) { return bar0 + 8u * (TC_S + s); };
  auto b_full = [&](uint32_t s) { return bar0 + 8u * (2 * TC_S + s); };
  auto b_empty = [&](uint32_t s) { return bar0 + 8u * (3 * TC_S + s); };
  // accumulators are double-buffered in tensor memory (two 256-column halves) whenever a work item needs <= 256 columns
  auto tmem_full = [&](uint32_t x) { return bar0 + 8u * (4 * TC_S + x); };
  auto tmem_empty = [&](uint32_t x) { return bar0 + 8u * (4 * TC_S + 2 + x); };
  const uint32_t tmem_ptr_addr = bar0 + 8u * (4 * TC_S + 4);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int KB = p.KB0 + p.KB1;
  const int NCH = p.nch;
  const bool dbuf = NCH * TC_NC <= 256;
  const uint32_t stage_bytes = (uint32_t)p.planes * TC_PLANE_BYTES;

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < SA; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (uint32_t s = 0; s < SB; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (uint32_t x = 0; x < 2; ++x) { mbar_init(tmem_full(x), 1); mbar_init(tmem_empty(x), 4 * EG); }   // empty: one arrive per epilogue warp
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
    {
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0;   // ring positions and phase bits (whole warp runs the loop; one elected lane issues)
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int lt = tile, a_r = 0, a_k = 0, b_r = 0, b_k = 0;
        if (p.bat_inner) {
          const int bat = tile / p.bat_tiles, bo = bat / p.bat_inner, bi = bat - bo * p.bat_inner;
          lt = tile - bat * p.bat_tiles;
          a_r = bo * p.a_row_s0 + bi * p.a_row_s1; a_k = bi * p.a_k_s1;
          b_r = bo * p.b_row_s0 + bi * p.b_row_s1; b_k = p.b_k0 + bi * p.b_k_s1;
        }
        const int mt = p.chunk_minor ? lt / p.chunk_minor : lt % p.m_tiles, ng = p.chunk_minor ? lt - mt * p.chunk_minor : lt / p.m_tiles;
        const int m0 = mt * TC_BM + a_r, n0 = ng * NCH * TC_NC + b_r;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(a_empty(sa), pa ^ 1u);
          const bool first = kb < p.KB0;
          const int ka = (first ? kb : kb - p.KB0) * TC_BK + a_k;
          const uint32_t dstA = a_ring + sa * 2 * TC_PLANE_BYTES;
          if (elect_one()) {
            mbar_expect_tx(a_full(sa), stage_bytes);
            if (p.mn_major) {       // inner coordinate = MN (channel), row coordinate = K (row of the activation tensor)
              tma_load_2d(dstA, &mA0h, a_full(sa), m0, ka); tma_load_2d(dstA + 8192u, &mA0h, a_full(sa), m0 + 64, ka);
              tma_load_2d(dstA + TC_PLANE_BYTES, &mA0l, a_full(sa), m0, ka); tma_load_2d(dstA + TC_PLANE_BYTES + 8192u, &mA0l, a_full(sa), m0 + 64, ka);
            } else {
              tma_load_2d(dstA, first ? &mA0h : &mA1h, a_full(sa), ka, m0);
              if (p.planes == 2) tma_load_2d(dstA + TC_PLANE_BYTES, first ? &mA0l : &mA1l, a_full(sa), ka, m0);
            }
          }
          __syncwarp();
          if (++sa == SA) { sa = 0; pa ^= 1u; }
          for (int c = 0; c < NCH; ++c) {
            mbar_wait(b_empty(sb), pb ^ 1u);
            const uint32_t dstB = b_ring + sb * 2 * TC_PLANE_BYTES;
            if (elect_one()) {
              mbar_expect_tx(b_full(sb), stage_bytes);
              if (p.mn_major) {
                const int kk = kb * TC_BK + b_k, nn = n0 + c * TC_NC;
                tma_load_2d(dstB, &mBh, b_full(sb), nn, kk); tma_load_2d(dstB + 8192u, &mBh, b_full(sb), nn + 64, kk);
                tma_load_2d(dstB + TC_PLANE_BYTES, &mBl, b_full(sb), nn, kk); tma_load_2d(dstB + TC_PLANE_BYTES + 8192u, &mBl, b_full(sb), nn + 64, kk);
              } else {
                tma_load_2d(dstB, &mBh, b_full(sb), kb * TC_BK + b_k, n0 + c * TC_NC);
                if (p.planes == 2) tma_load_2d(dstB + TC_PLANE_BYTES, &mBl, b_full(sb), kb * TC_BK + b_k, n0 + c * TC_NC);
              }
            }
            __syncwarp();
            if (++sb == SB) { sb = 0; pb ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    {
      const uint32_t idesc = make_idesc_bf16(TC_BM, p.mma_n) | (p.mn_major ? ((1u << 15) | (1u << 16)) : 0u);      // a_major / b_major = MN
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0, it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const uint32_t buf = dbuf ? (it & 1u) : 0u, use = dbuf ? (it >> 1) : it;
        mbar_wait(tmem_empty(buf), (use & 1u) ^ 1u);      // epilogue has drained this half's previous accumulators
        tc_fence_after();
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(a_full(sa), pa);
          tc_fence_after();
          const uint32_t aH = a_ring + sa * 2 * TC_PLANE_BYTES, aL = aH + TC_PLANE_BYTES;
          for (int c = 0; c < NCH; ++c) {
            mbar_wait(b_full(sb), pb);
            tc_fence_after();
            const uint32_t bH = b_ring + sb * 2 * TC_PLANE_BYTES, bL = bH + TC_PLANE_BYTES;
            const uint32_t d = tmem_base + 256u * buf + (uint32_t)(c * TC_NC);
            const bool mn = p.mn_major != 0;
            const uint64_t dAH = mn ? make_sw128_desc_mn(aH) : make_sw128_desc(aH), dAL = mn ? make_sw128_desc_mn(aL) : make_sw128_desc(aL);
            const uint64_t dBH = mn ? make_sw128_desc_mn(bH) : make_sw128_desc(bH), dBL = mn ? make_sw128_desc_mn(bL) : make_sw128_desc(bL);
            const uint32_t kadv = mn ? 128u : 2u;        // K-major: 16 bf16 = 32 B inside the swizzle atom; MN-major: 16 K rows x 128 B = 2 KB
            if (elect_one()) {
#pragma unroll
              for (int ks = 0; ks < TC_BK / 16; ++ks) {
                umma_kstep(d, dAH + kadv * ks, dAL + kadv * ks, dBH + kadv * ks, dBL + kadv * ks, idesc, (kb > 0 || ks > 0) ? 1u : 0u, p.planes == 2);
                if (p.lolo && p.planes == 2) umma_lolo(d, dAL + kadv * ks, dBL + kadv * ks, idesc);
              }
              tc_commit(b_empty(sb));     // frees this weight stage once the MMAs above retire
            }
            __syncwarp();
            if (++sb == SB) { sb = 0; pb ^= 1u; }
          }
          tc_commit_elect(a_empty(sa));
          if (++sa == SA) { sa = 0; pa ^= 1u; }
        }
        tc_commit_elect(tmem_full(buf));      // accumulators complete -> epilogue
      }
    }
  } else {
    // ================================ epilogue (warps 2..9) ================================
    // two groups of four warps (one warp per TMEM lane quadrant in each): the groups take alternate column steps of every work item, so each
    // scheduler has two epilogue warps to interleave (one warp alone is bound by its own instruction latencies)
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
    const int grp = EG == 2 ? (warp - 2) >> 2 : 0;
    const int row_in_tile = quad * 32 + lane;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = dbuf ? (it & 1u) : 0u, use = dbuf ? (it >> 1) : it;
      mbar_wait(tmem_full(buf), use & 1u);
      tc_fence_after();
      int lt = tile; long long o_off = 0;
      if (p.bat_inner) {
        const int bat = tile / p.bat_tiles, bo = bat / p.bat_inner, bi = bat - bo * p.bat_inner;
        lt = tile - bat * p.bat_tiles;
        o_off = bo * p.o_s0 + bi * p.o_s1;
      }
      const int mt = p.chunk_minor ? lt / p.chunk_minor : lt % p.m_tiles, ng = p.chunk_minor ? lt - mt * p.chunk_minor : lt / p.m_tiles;
      const long long m = (long long)mt * TC_BM + row_in_tile;
      const int n0 = ng * NCH * TC_NC;
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
      const uint32_t trow = tmem_base + 256u * buf + ((uint32_t)(quad * 32) << 16);
      if (p.epi == TC_EPI_F32) {
        const float rm = (valid && p.rowmask) ? p.rowmask[m] : 1.f;
        const int ncols = min(NCH * TC_NC, ((p.n_valid - n0 + 31) / 32) * 32);     // skip accumulator columns beyond the valid width
        for (int c0 = grp * 32; c0 < ncols; c0 += 32 * EG) {
          const int n = n0 + c0;
          const bool act = valid && n < p.n_valid;
          // bias / residual of this 32-column group are requested before the accumulator read: their latencies overlap instead of
          // serialising behind the stores (the compiler cannot hoist them itself: out_f32 may alias)
          float4 bv[8], rv[8];
          const float* rrow = (act && p.residual) ? p.residual + m * p.ldr + n : nullptr;
          const bool gather_j = p.rowadd != nullptr && n + 32 <= p.n_valid;       // warp-uniform: the j-side node terms come through the staging tile
          uint32_t aj[32];
          if (gather_j) warp_gather_rows32f(bar0 + 1024u + (uint32_t)(warp - 2) * 4096u, aj, add_j ? add_j + n : nullptr, lane);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const bool in = act && n + q * 4 < p.n_valid;     // n_valid is a multiple of 4 for every linear routed here
            bv[q] = (in && p.bias) ? __ldg(reinterpret_cast<const float4*>(p.bias + n + q * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            rv[q] = (in && rrow) ? *reinterpret_cast<const float4*>(rrow + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (in && add_i) {
              const float4 x = *reinterpret_cast<const float4*>(add_i + n + q * 4);
              float4 y;
              if (gather_j) y = make_float4(__uint_as_float(aj[q * 4]), __uint_as_float(aj[q * 4 + 1]), __uint_as_float(aj[q * 4 + 2]), __uint_as_float(aj[q * 4 + 3]));
              else y = *reinterpret_cast<const float4*>(add_j + n + q * 4);
              bv[q].x += x.x + y.x; bv[q].y += x.y + y.y; bv[q].z += x.z + y.z; bv[q].w += x.w + y.w;
            }
          }
          uint32_t r[32];
          tmem_ld32(trow + (uint32_t)c0, r);
          if (p.atomic) {                     // warp-uniform
            if (act) {
              float* orow = p.out_f32 + o_off + m * p.ldo + n;
              const float al = p.alpha != 0.f ? p.alpha : 1.f;
#pragma unroll
              for (int q = 0; q < 32; ++q)
                if (n + q < p.n_valid) atomicAdd(orow + q, al * __uint_as_float(r[q]));
            }
          } else {
            // full 32-column groups leave through the warp's staging tile (whole 128-byte row segments per store instruction); a ragged last
            // group is stored row-per-thread
            const bool full = n + 32 <= p.n_valid && (p.ldo & 3) == 0 && (o_off & 3) == 0;        // warp-uniform
            uint32_t wv[32];
            if (act) {
              float* orow = p.out_f32 + o_off + m * p.ldo + n;
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                if (n + q * 4 < p.n_valid) {
                  float4 v = make_float4(__uint_as_float(r[q * 4 + 0]), __uint_as_float(r[q * 4 + 1]), __uint_as_float(r[q * 4 + 2]), __uint_as_float(r[q * 4 + 3]));
                  if (p.alpha != 0.f) { v.x *= p.alpha; v.y *= p.alpha; v.z *= p.alpha; v.w *= p.alpha; }
                  v.x += bv[q].x; v.y += bv[q].y; v.z += bv[q].z; v.w += bv[q].w;
                  if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                  v.x *= rm; v.y *= rm; v.z *= rm; v.w *= rm;
                  v.x += rv[q].x; v.y += rv[q].y; v.z += rv[q].z; v.w += rv[q].w;
                  if (p.relumask) {
                    const float4 mk = *reinterpret_cast<const float4*>(p.relumask + m * p.ldm + n + q * 4);
                    v.x = mk.x > 0.f ? v.x : 0.f; v.y = mk.y > 0.f ? v.y : 0.f; v.z = mk.z > 0.f ? v.z : 0.f; v.w = mk.w > 0.f ? v.w : 0.f;
                  }
                  if (full) {
                    wv[q * 4 + 0] = __float_as_uint(v.x); wv[q * 4 + 1] = __float_as_uint(v.y); wv[q * 4 + 2] = __float_as_uint(v.z); wv[q * 4 + 3] = __float_as_uint(v.w);
                  } else {
                    *reinterpret_cast<float4*>(orow + q * 4) = v;
                  }
                }
              }
            }
            if (full) warp_store_rows32f(bar0 + 1024u + (uint32_t)(warp - 2) * 4096u, wv, p.out_f32 + o_off + n, m - lane, p.M, lane, p.ldo);
          }
        }
      } else if (p.epi == TC_EPI_RELU) {
        // 64 columns per step: accumulator -> (+bias, +node terms) -> ReLU or ReLU-mask -> bf16 hi/lo -> coalesced stores through the warp's staging tile
        const uint32_t stage = bar0 + 1024u + (uint32_t)(warp - 2) * 4096u;
        const long long m_warp = m - lane;
        if (p.maskplane) {
          // backward of a ReLU: acc (+bias) where the activation plane is non-zero, 0 elsewhere
          for (int c0 = grp * 64; c0 < NCH * TC_NC; c0 += 64 * EG) {      // accumulator column c0 = output column n0 + c0
            uint32_t r0[32], r1[32];
            tmem_ld32_nowait(trow + (uint32_t)c0, r0);
            tmem_ld32_nowait(trow + (uint32_t)(c0 + 32), r1);
            uint32_t mk[32];
            warp_load_rows64_ld(stage, mk, p.maskplane, m_warp, n0 + c0, p.M, lane, p.N);
            tmem_ld_wait();
            uint32_t hw[32], lw[32];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const int cc = n0 + c0 + half * 32 + q * 4;
                const float4 bi = (valid && p.bias) ? *reinterpret_cast<const float4*>(p.bias + cc) : make_float4(0.f, 0.f, 0.f, 0.f);
                const uint32_t* rr = half ? r1 : r0;
                float v0 = __uint_as_float(rr[q * 4 + 0]) + bi.x, v1 = __uint_as_float(rr[q * 4 + 1]) + bi.y;
                float v2 = __uint_as_float(rr[q * 4 + 2]) + bi.z, v3 = __uint_as_float(rr[q * 4 + 3]) + bi.w;
                // word j of mk holds columns 2j, 2j+1 of this 64-column step: non-zero bf16 <=> activation > 0
                const uint32_t ma = mk[half * 16 + q * 2], mb = mk[half * 16 + q * 2 + 1];
                v0 = (ma & 0x00007fffu) ? v0 : 0.f; v1 = (ma & 0x7fff0000u) ? v1 : 0.f;
                v2 = (mb & 0x00007fffu) ? v2 : 0.f; v3 = (mb & 0x7fff0000u) ? v3 : 0.f;
                split2_bf16(v0, v1, hw[half * 16 + q * 2], lw[half * 16 + q * 2]);
                split2_bf16(v2, v3, hw[half * 16 + q * 2 + 1], lw[half * 16 + q * 2 + 1]);
              }
            }
            warp_store_rows64_ld(stage, hw, p.out_hi, m_warp, n0 + c0, p.M, lane, p.N);
            if (p.planes == 2) warp_store_rows64_ld(stage, lw, p.out_lo, m_warp, n0 + c0, p.M, lane, p.N);
          }
        } else {
          // forward: relu(acc + bias + node terms); the j-side node terms of the warp's 32 rows are gathered through the staging tile
          for (int c0 = grp * 64; c0 < NCH * TC_NC; c0 += 64 * EG) {
            uint32_t hw[32], lw[32];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              uint32_t r[32], aj[32];
              tmem_ld32_nowait(trow + (uint32_t)(c0 + half * 32), r);
              if (p.rowadd) warp_gather_rows32f(stage, aj, add_j ? add_j + n0 + c0 + half * 32 : nullptr, lane);
              tmem_ld_wait();
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const int cc = n0 + c0 + half * 32 + q * 4;
                float4 bi = (valid && p.bias) ? *reinterpret_cast<const float4*>(p.bias + cc) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (add_i) {
                  const float4 x = *reinterpret_cast<const float4*>(add_i + cc);
                  bi.x += x.x + __uint_as_float(aj[q * 4]); bi.y += x.y + __uint_as_float(aj[q * 4 + 1]);
                  bi.z += x.z + __uint_as_float(aj[q * 4 + 2]); bi.w += x.w + __uint_as_float(aj[q * 4 + 3]);
                }
                const float v0 = fmaxf(__uint_as_float(r[q * 4 + 0]) + bi.x, 0.f), v1 = fmaxf(__uint_as_float(r[q * 4 + 1]) + bi.y, 0.f);
                const float v2 = fmaxf(__uint_as_float(r[q * 4 + 2]) + bi.z, 0.f), v3 = fmaxf(__uint_as_float(r[q * 4 + 3]) + bi.w, 0.f);
                split2_bf16(v0, v1, hw[half * 16 + q * 2], lw[half * 16 + q * 2]);
                split2_bf16(v2, v3, hw[half * 16 + q * 2 + 1], lw[half * 16 + q * 2 + 1]);
              }
            }
            warp_store_rows64_ld(stage, hw, p.out_hi, m_warp, n0 + c0, p.M, lane, p.N);
            if (p.planes == 2) warp_store_rows64_ld(stage, lw, p.out_lo, m_warp, n0 + c0, p.M, lane, p.N);
          }
        }
      } else if (grp == 0) {   // LayerNorm over the 128 columns of this row, then edge mask (whole rows per thread: one group works)
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
      if (lane == 0) mbar_arrive(tmem_empty(buf));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Fused EdgeTransition kernel (model/ipa_pytorch.py:218-233): one launch per layer, z read once and written once.
//
//   per 128-edge tile:  h1 = relu(z·W1z^T + P_i + Q_j)   6 chunks of 64 hidden units: G1_c (SS MMA, A = z tile in smem) -> TMEM T1[c&1]
//                                                          -> epilogue -> bf16 hi/lo written back IN PLACE into T1[c&1]
//                       H2 += h1_c · W2[:, 64c:64c+64]^T  G2_c: TS MMA, A = T1[c&1] (TMEM), K = 64, N = 384 -> TMEM H2
//                       h2 = relu(H2 + b2)                6 chunks of 64 -> bf16 hi/lo in place over H2's own columns
//                       Y  = z·Wfz^T (G3z, SS) + sum_c h2_c · Wf[:, 64c:64c+64]^T (G3_c, TS)      -> TMEM Y (reuses T1)
//                       z' = LayerNorm(Y + U_i + V_j) · m_i m_j  -> bf16 hi/lo planes (in place: tiles are row-disjoint)
//   TMEM: [0,64) T1a  [64,128) T1b  [128,512) H2 ; Y = [0,128).  A-operand image of a 64-wide chunk inside its 64 fp32 columns:
//         columns [0,32) = hi (two bf16 per 32-bit column, K ascending), columns [32,64) = lo.
//   Activations never touch shared memory: the MMAs of the two big GEMMs read A from TMEM, so shared-memory bandwidth only
//   carries the weight blocks (an SS 128x128x16 MMA needs 128 B/clk of smem — the whole budget — a TS one 64 B/clk).
//   smem: z tile 64 KB (TMA)  |  weight ring 160 KB (TMA; every weight block streams from L2 once per tile)  |  mbarriers.
//   warps: 0 = TMA producer, 1 = MMA issuer (+TMEM alloc), 2..9 = epilogue: two groups of four warps (one per TMEM lane quadrant);
//          group g converts the chunks of parity g, 64 columns per thread-row per step.
//   Ordering between MMAs that reuse TMEM columns (G2_c reads T1[x] then G1_{c+2} overwrites it; G3_5 reads H2 then the next
//   tile's G2_0 overwrites it) relies on tcgen05.mma executing in issue order.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
      "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
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
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// TS k-step: A (bf16, K = 16 -> 8 TMEM columns) from tensor memory, B from shared memory
__device__ __forceinline__ void umma_kstep_ts(uint32_t d, uint32_t aH, uint32_t aL, uint64_t bH, uint64_t bL, uint32_t idesc, uint32_t acc0,
                                              bool three_pass) {
  if (three_pass) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.eq.b32 q, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %3, %5, p;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %4, %5, q;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%2], %3, %5, q;\n\t}" ::"r"(d),
        "r"(aH), "r"(aL), "l"(bH), "l"(bL), "r"(idesc), "r"(acc0)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
        "r"(aH), "l"(bH), "r"(idesc), "r"(acc0)
        : "memory");
  }
}

// Warp-cooperative store of a [32 rows x 64 bf16] block held one row per lane (w = the row's 32 packed words) through a 4 KB
// XOR-swizzled shared-memory staging buffer: every global store instruction then writes four full 128-byte row segments
// instead of 32 scattered 16-byte pieces (8x fewer LSU wavefronts; the LayerNorm epilogues were bound by them).
__device__ __forceinline__ void warp_store_rows64(uint32_t stage, const uint32_t (&w)[32], __nv_bfloat16* out, long long m_warp, int col0,
                                                  long long E, int lane) {
#pragma unroll
  for (int u = 0; u < 8; ++u)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + (uint32_t)(lane * 128 + ((u ^ (lane & 7)) << 4))), "r"(w[4 * u]),
                 "r"(w[4 * u + 1]), "r"(w[4 * u + 2]), "r"(w[4 * u + 3])
                 : "memory");
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int r = k * 4 + (lane >> 3), u = lane & 7;
    uint4 val;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                 : "r"(stage + (uint32_t)(r * 128 + ((u ^ (r & 7)) << 4))) : "memory");
    if (m_warp + r < E) *reinterpret_cast<uint4*>(out + (m_warp + r) * 128 + col0 + u * 8) = val;
  }
  __syncwarp();
}

// Weight-ring helpers as force-inlined members (warp-uniform: the whole warp runs them, one elected lane issues)
struct FuRing {
  uint32_t ringb, bar0, slot_bytes, S, planes;
  uint32_t iw;
  long long c_wait; bool prof_on;
  uint32_t cl, rank;                       // cluster mode (2-CTA weight multicast) and this CTA's rank in the pair
  __device__ __forceinline__ uint32_t slot(uint32_t s, int pl) const { return ringb + s * slot_bytes + (uint32_t)pl * TC_PLANE_BYTES; }
  __device__ __forceinline__ uint32_t full(uint32_t s) const { return bar0 + 8u * s; }
  __device__ __forceinline__ uint32_t empty(uint32_t s) const { return bar0 + 8u * (10 + s); }
  // producer: wait for a free slot, arm it, issue the TMA loads of one weight block.
  // Cluster mode (cl != 0): the two CTAs of a pair consume the same weight sequence, so each block is fetched from L2 ONCE and
  // multicast into both rings.  Per slot: `empty` = this CTA's MMAs retired; `free2` (count 2) = both CTAs' slots are free (each
  // producer arrives locally and on its peer's barrier); loads alternate between the two CTAs as issuers.
  __device__ __forceinline__ uint32_t free2(uint32_t s) const { return bar0 + 8u * (44 + s); }
  __device__ __forceinline__ void load(const CUtensorMap* mh, const CUtensorMap* ml, int k, int n, uint32_t bytes_per_plane) {
    const uint32_t s = iw % S, ph = (iw / S) & 1u;
    if (prof_on) { const long long t0 = clock64(); mbar_wait(empty(s), ph ^ 1u); c_wait += clock64() - t0; } else mbar_wait(empty(s), ph ^ 1u);
    if (cl) {
      if (elect_one()) { mbar_arrive(free2(s)); mbar_arrive_remote(free2(s), rank ^ 1u); }
      __syncwarp();
      mbar_wait_cluster(free2(s), ph);
    }
    if (elect_one()) {
      mbar_expect_tx(full(s), planes * bytes_per_plane);
      if (!cl) {
        tma_load_2d(slot(s, 0), mh, full(s), k, n);
        if (planes == 2) tma_load_2d(slot(s, 1), ml, full(s), k, n);
      } else if ((iw & 1u) == rank) {
        tma_load_2d_mc(slot(s, 0), mh, full(s), k, n, (uint16_t)3);
        if (planes == 2) tma_load_2d_mc(slot(s, 1), ml, full(s), k, n, (uint16_t)3);
      }
    }
    __syncwarp();
    ++iw;
  }
  __device__ __forceinline__ uint32_t acquire() {
    const uint32_t s = iw % S, ph = (iw / S) & 1u;
    if (prof_on) { const long long t0 = clock64(); mbar_wait(full(s), ph); c_wait += clock64() - t0; } else mbar_wait(full(s), ph);
    tc_fence_after();
    return s;
  }
  // MMA issuer, SS: one weight block (K = 64) against A planes in shared memory
  __device__ __forceinline__ void mma_ss(uint32_t aH, uint32_t aL, uint32_t d, uint32_t idesc, bool zero_first) {
    const uint32_t s = acquire();
    const uint64_t dAH = make_sw128_desc(aH), dAL = make_sw128_desc(aL);
    const uint64_t dBH = make_sw128_desc(slot(s, 0)), dBL = make_sw128_desc(slot(s, 1));
    if (elect_one()) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_kstep(d, dAH + 2u * ks, dAL + 2u * ks, dBH + 2u * ks, dBL + 2u * ks, idesc, (zero_first && ks == 0) ? 0u : 1u, planes == 2);
      tc_commit(empty(s));
    }
    __syncwarp();
    ++iw;
  }
  // MMA issuer, TS: A chunk image in tensor memory (hi at a, lo at a + 32 columns)
  __device__ __forceinline__ void mma_ts(uint32_t a, uint32_t d, uint32_t idesc, bool zero_first) {
    const uint32_t s = acquire();
    const uint64_t dBH = make_sw128_desc(slot(s, 0)), dBL = make_sw128_desc(slot(s, 1));
    if (elect_one()) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_kstep_ts(d, a + 8u * ks, a + 32u + 8u * ks, dBH + 2u * ks, dBL + 2u * ks, idesc, (zero_first && ks == 0) ? 0u : 1u, planes == 2);
      tc_commit(empty(s));
    }
    __syncwarp();
    ++iw;
  }
};

constexpr int TC_QKV = PROJ_Q + PROJ_KV;   // q | k,v columns of the fused IPA projection (2048 + 4096); also the width of the logits operand planes
constexpr int FU_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two groups of four)
constexpr int FU_RING_BYTES = 128 * 1024;
constexpr size_t FU_SMEM_BYTES = 64 * 1024 + FU_RING_BYTES + 512 + 2048 + 8 * 4096;   // + LayerNorm statistics + 8 x 4 KB store staging   // dynamic smem starts 1024-aligned (no static smem here)

struct FusedParams {
  int E, planes, nres, num_tiles;
  const float* pquv;        // [B*nres, 1024] = P | Q | U | V node terms
  const float* b2; const float* ln_g; const float* ln_b; const float* res_mask;
  __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;
  long long* prof;          // optional [32] cycle counters written by CTA 0 (developer aid)
  int dbg_noq;              // developer experiment: skip the node-term loads (results wrong; isolates their cost)
};

#define FU_PROF(ctr, stmt)                               \
  do {                                                   \
    if (prof_on) { const long long _t0 = clock64(); stmt; ctr += clock64() - _t0; } else { stmt; } \
  } while (0)

template <bool CL>
__global__ void __launch_bounds__(FU_THREADS, 1)
tc_edge_fused_kernel(const __grid_constant__ CUtensorMap mZh, const __grid_constant__ CUtensorMap mZl,
                     const __grid_constant__ CUtensorMap mW1h, const __grid_constant__ CUtensorMap mW1l,
                     const __grid_constant__ CUtensorMap mW2h, const __grid_constant__ CUtensorMap mW2l,
                     const __grid_constant__ CUtensorMap mWfh, const __grid_constant__ CUtensorMap mWfl, const FusedParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t PL = TC_PLANE_BYTES;   // 16 KB
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t zb = base;                              // z(kb, plane) = zb + (kb*2 + plane)*PL
  const uint32_t ringb = zb + 4 * PL;
  const uint32_t bar0 = ringb + FU_RING_BYTES;
  const int S = p.planes == 2 ? 4 : 8;
  const uint32_t slot_bytes = (uint32_t)p.planes * PL;
  auto zbuf = [&](int kb, int pl) { return zb + (uint32_t)(kb * 2 + pl) * PL; };
  // barriers: ring full[10] (slots 0..9), ring empty[10] (10..19), then the named ones
  const uint32_t z_full = bar0 + 8u * 20, z_empty = bar0 + 8u * 21;
  auto t1_full = [&](int x) { return bar0 + 8u * (22 + x); };
  auto a_full = [&](int x) { return bar0 + 8u * (24 + x); };
  const uint32_t h2_full = bar0 + 8u * 26, y_full = bar0 + 8u * 27, y_empty = bar0 + 8u * 28;
  auto a2_full = [&](int c) { return bar0 + 8u * (29 + c); };      // one per h2 chunk (no back-pressure on epi2: avoid 2-phase run-ahead)
  const uint32_t tmem_ptr_addr = bar0 + 8u * 35;
  const uint32_t stats = bar0 + 512u;     // LayerNorm partial statistics [2][128 rows][2] fp32 = 2 KB

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform by construction
  if (threadIdx.x == 0) {
    for (int s = 0; s < 20; ++s) mbar_init(bar0 + 8u * s, 1);
    mbar_init(z_full, 1); mbar_init(z_empty, 1);
    for (int x = 0; x < 2; ++x) { mbar_init(t1_full(x), 1); mbar_init(a_full(x), 4); }
    for (int c = 0; c < 6; ++c) mbar_init(a2_full(c), 4);
    mbar_init(h2_full, 1); mbar_init(y_full, 1); mbar_init(y_empty, 8);
    if (CL) for (int s2 = 0; s2 < 10; ++s2) mbar_init(bar0 + 8u * (44 + s2), 2);     // free2: both CTAs of the pair released the slot
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&mZh); tma_prefetch_desc(&mW1h); tma_prefetch_desc(&mW2h); tma_prefetch_desc(&mWfh);
    if (p.planes == 2) { tma_prefetch_desc(&mZl); tma_prefetch_desc(&mW1l); tma_prefetch_desc(&mW2l); tma_prefetch_desc(&mWfl); }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (CL) cluster_sync_all();      // the peer's barriers are initialised before any remote arrive / multicast reaches them
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));
  // Tiles of this CTA: blockIdx.x, + gridDim.x, ...  In cluster mode both CTAs of a pair run the same number of iterations (the
  // weight sequence is shared); an iteration past the last tile computes on zero-filled TMA boxes and stores nothing.
  const uint32_t crank = CL ? cluster_ctarank() : 0u;
  const int first = CL ? (int)(blockIdx.x - crank) : (int)blockIdx.x;
  const uint32_t n_it = first < p.num_tiles ? (uint32_t)((p.num_tiles - 1 - first) / (int)gridDim.x + 1) : 0u;

  if (warp == 0) {
    // ============================================ TMA producer ============================================
    uint32_t it = 0;
    const bool prof_on = p.prof != nullptr && blockIdx.x == 0;
    long long c_zempty = 0; const long long c_start = clock64();
    FuRing rg{ringb, bar0, slot_bytes, (uint32_t)S, (uint32_t)p.planes, 0u, 0, prof_on, CL ? 1u : 0u, crank};
    for (int tile = blockIdx.x; it < n_it; tile += gridDim.x, ++it) {
      const int m0 = tile * TC_BM;
      FU_PROF(c_zempty, mbar_wait(z_empty, (it & 1u) ^ 1u));
      if (elect_one()) {
        mbar_expect_tx(z_full, (uint32_t)p.planes * 2 * PL);
        for (int kb = 0; kb < 2; ++kb) {
          tma_load_2d(zbuf(kb, 0), &mZh, z_full, kb * 64, m0);
          if (p.planes == 2) tma_load_2d(zbuf(kb, 1), &mZl, z_full, kb * 64, m0);
        }
      }
      __syncwarp();
      // weight blocks in exactly the MMA issue order: G1_0 G1_1 { G2_c G1_{c+2} } G3z G3_c
      for (int c = 0; c < 2; ++c)
        for (int kb = 0; kb < 2; ++kb) rg.load(&mW1h, &mW1l, kb * 64, c * 64, PL / 2);
      for (int c = 0; c < 6; ++c) {
        for (int n = 0; n < 3; ++n) rg.load(&mW2h, &mW2l, c * 64, n * 128, PL);
        if (c + 2 < 6)
          for (int kb = 0; kb < 2; ++kb) rg.load(&mW1h, &mW1l, kb * 64, (c + 2) * 64, PL / 2);
      }
      for (int kb = 0; kb < 2; ++kb) rg.load(&mWfh, &mWfl, (6 + kb) * 64, 0, PL);   // G3z: Wfz
      for (int c = 0; c < 6; ++c) rg.load(&mWfh, &mWfl, c * 64, 0, PL);             // G3_c: Wf
    }
    if (prof_on && lane == 0) { p.prof[0] = clock64() - c_start; p.prof[1] = rg.c_wait; p.prof[2] = c_zempty; p.prof[3] = it; }
  } else if (warp == 1) {
    // ============================================ MMA issuer ============================================
    const uint32_t idesc64 = make_idesc_bf16(TC_BM, 64), idesc128 = make_idesc_bf16(TC_BM, 128);
    uint32_t it = 0, n_af[2] = {0, 0};
    const bool prof_on = p.prof != nullptr && blockIdx.x == 0;
    long long c_afull = 0, c_tile0 = 0; const long long c_start = clock64();
    FuRing rg{ringb, bar0, slot_bytes, (uint32_t)S, (uint32_t)p.planes, 0u, 0, prof_on, CL ? 1u : 0u, crank};
    // Tensor-memory layout alternates with the tile parity so that the next tile's first GEMMs never wait for the LayerNorm:
    //   even tiles: T1 = [0,128)   (T1a | T1b), H2 = [128,512), Y = T1
    //   odd tiles:  T1 = [384,512),             H2 = [0,384),   Y = T1
    // An odd tile's T1 is the even tile's H2 chunks 4,5 (consumed by its G3_4/G3_5, earlier in issue order) and vice versa; only
    // G2_0, the first write into the new H2 (which covers the previous Y), waits for the LayerNorm's read of Y.
    // G1_c: T1[c&1] = z · W1z[64c:64c+64, :]^T  (two k-blocks of z)
#define FU_G1(c_)                                                                                                            \
  do {                                                                                                                       \
    const int x_ = (c_) & 1;                                                                                                  \
    rg.mma_ss(zbuf(0, 0), zbuf(0, 1), t1b + (uint32_t)(x_ * 64), idesc64, true);                                              \
    rg.mma_ss(zbuf(1, 0), zbuf(1, 1), t1b + (uint32_t)(x_ * 64), idesc64, false);                                             \
    tc_commit_elect(t1_full(x_));                                                                                             \
  } while (0)
    for (int tile = blockIdx.x; it < n_it; tile += gridDim.x, ++it) {
      const uint32_t t1b = tmem_base + ((it & 1u) ? 384u : 0u), h2b = tmem_base + ((it & 1u) ? 0u : 128u);
      FU_PROF(c_tile0, mbar_wait(z_full, it & 1u));
      tc_fence_after();
      FU_G1(0);
      FU_G1(1);
      for (int c = 0; c < 6; ++c) {
        const int x = c & 1;
        if (c == 0) { FU_PROF(c_tile0, mbar_wait(y_empty, (it & 1u) ^ 1u)); }   // previous tile's Y (inside this tile's H2) is in registers
        FU_PROF(c_afull, mbar_wait(a_full(x), n_af[x] & 1u)); ++n_af[x];     // epilogue wrote h1_c (bf16 hi/lo) into T1[x]
        tc_fence_after();
#pragma unroll
        for (int n = 0; n < 3; ++n) rg.mma_ts(t1b + (uint32_t)(x * 64), h2b + (uint32_t)(n * 128), idesc128, c == 0);
        if (c == 5) tc_commit_elect(h2_full);
        if (c + 2 < 6) FU_G1(c + 2);                               // overwrites T1[x]: ordered after G2_c by issue order
      }
      rg.mma_ss(zbuf(0, 0), zbuf(0, 1), t1b, idesc128, true);     // Y = z · Wfz^T   (T1 is free: G2_4/G2_5 precede in issue order)
      rg.mma_ss(zbuf(1, 0), zbuf(1, 1), t1b, idesc128, false);
      tc_commit_elect(z_empty);
      for (int c = 0; c < 6; ++c) {
        FU_PROF(c_afull, mbar_wait(a2_full(c), it & 1u));                    // epilogue wrote h2_c into H2's columns [64c, 64c+64)
        tc_fence_after();
        rg.mma_ts(h2b + (uint32_t)(c * 64), t1b, idesc128, false);
      }
      tc_commit_elect(y_full);
    }
#undef FU_G1
    if (prof_on && lane == 0) { p.prof[8] = clock64() - c_start; p.prof[9] = rg.c_wait; p.prof[10] = c_afull; p.prof[12] = c_tile0; }
  } else {
    // ============================================ epilogue warps 2..9 ============================================
    // Two independent groups of four warps (one warp per TMEM lane quadrant).  Group g owns T1[g]: it converts the h1 chunks
    // c = g, g+2, g+4 and the h2 chunks of the same parity, a whole 64-column chunk per step per thread-row, so the two groups
    // work on consecutive chunks concurrently and the MMA pipe always has the other group's chunk to consume.  Per-step fixed
    // latencies (barrier wake-up, tcgen05.ld/st round trips) are what bound this path; fewer, larger steps amortise them.
    const int quad = warp & 3, grp = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
    uint32_t it = 0, n_t1f = 0;
    const bool prof_on = p.prof != nullptr && blockIdx.x == 0 && warp == 2;
    long long c_t1f = 0, c_h2f = 0, c_yf = 0, c_ln = 0; const long long c_start = clock64();
    // v[64] (fp32, this row's chunk) -> bf16 hi/lo images over the chunk's own 64 TMEM columns: hi words -> columns [0,32),
    // lo words -> columns [32,64).  All reads of the fp32 columns by this thread precede the stores (same thread, in order).
    auto write_a = [&](uint32_t tchunk, const float (&v)[64]) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t h[16], l[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) split2_bf16(v[half * 32 + 2 * e], v[half * 32 + 2 * e + 1], h[e], l[e]);
        tmem_st16(tchunk + (uint32_t)(16 * half), h);
        if (p.planes == 2) tmem_st16(tchunk + 32u + (uint32_t)(16 * half), l);
      }
      tmem_st_wait();
    };
    struct RowCtx { long long m; const float* pP; const float* pQ; float emask; bool valid; };
    auto row_ctx = [&](int tile) {
      RowCtx rc;
      rc.m = (long long)tile * TC_BM + row;
      rc.pP = nullptr; rc.pQ = nullptr; rc.emask = 0.f;
      const bool valid0 = rc.m < p.E;
      if (valid0) {
        const long long nn = (long long)p.nres * p.nres;
        const long long b = rc.m / nn;
        const int rem = (int)(rc.m - b * nn);
        const int ri = rem / p.nres, rj = rem - ri * p.nres;
        rc.pP = p.pquv + (b * p.nres + ri) * ET_NODE;          // P at +0, U at +768
        rc.pQ = p.pquv + (b * p.nres + rj) * ET_NODE + ET_HID;  // Q at +384 (-> +0 here), V at +896 (-> +512 here)
        rc.emask = p.res_mask[b * p.nres + ri] * p.res_mask[b * p.nres + rj];
      }
      rc.valid = valid0 && !p.dbg_noq;
      return rc;
    };
    // ---- epi1 step: h1 chunk c of the current tile (T1[grp] of this tile's layout).  Q_j (L2 latency) is requested before the
    //      wait on the accumulator, P_i (shared by the tile's rows, L1 hits) at use. ----
    auto epi1 = [&](const RowCtx& rc, uint32_t t1b, int c) {
      const bool valid = rc.valid; const float* pP = rc.pP; const float* pQ = rc.pQ;
      float v[64];
      const int col0 = c * 64;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float y8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (valid) ldg256(pQ + col0 + q * 8, y8);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[q * 8 + e] = y8[e];
      }
      FU_PROF(c_t1f, mbar_wait(t1_full(grp), n_t1f & 1u)); ++n_t1f;
      tc_fence_after();
      const uint32_t tchunk = trow + t1b + (uint32_t)(grp * 64);
      {
        uint32_t r0[32], r1[32];
        tmem_ld32_nowait(tchunk, r0);
        tmem_ld32_nowait(tchunk + 32u, r1);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 pa = valid ? __ldg(reinterpret_cast<const float4*>(pP + col0 + q * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 pb = valid ? __ldg(reinterpret_cast<const float4*>(pP + col0 + 32 + q * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
          v[q * 4 + 0] = fmaxf((__uint_as_float(r0[q * 4 + 0]) + pa.x) + v[q * 4 + 0], 0.f); v[q * 4 + 1] = fmaxf((__uint_as_float(r0[q * 4 + 1]) + pa.y) + v[q * 4 + 1], 0.f);
          v[q * 4 + 2] = fmaxf((__uint_as_float(r0[q * 4 + 2]) + pa.z) + v[q * 4 + 2], 0.f); v[q * 4 + 3] = fmaxf((__uint_as_float(r0[q * 4 + 3]) + pa.w) + v[q * 4 + 3], 0.f);
          v[32 + q * 4 + 0] = fmaxf((__uint_as_float(r1[q * 4 + 0]) + pb.x) + v[32 + q * 4 + 0], 0.f); v[32 + q * 4 + 1] = fmaxf((__uint_as_float(r1[q * 4 + 1]) + pb.y) + v[32 + q * 4 + 1], 0.f);
          v[32 + q * 4 + 2] = fmaxf((__uint_as_float(r1[q * 4 + 2]) + pb.z) + v[32 + q * 4 + 2], 0.f); v[32 + q * 4 + 3] = fmaxf((__uint_as_float(r1[q * 4 + 3]) + pb.w) + v[32 + q * 4 + 3], 0.f);
        }
      }
      write_a(tchunk, v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full(grp));
    };
    // ---- epi3 of tile lt (layout parity lt & 1): LayerNorm + mask + store.  Runs AFTER this group's first epi1 step of tile lt + 1, so
    //      the tensor pipe already works on the next tile while these warps normalise and store. ----
    auto epi3 = [&](const RowCtx& rc, uint32_t lt) {
      const bool valid = rc.valid; const float* pP = rc.pP; const float* pQ = rc.pQ; const float emask = rc.emask; const long long m = rc.m;
      const uint32_t ybase = (lt & 1u) ? 384u : 0u;
      // ---- epi3: LayerNorm + mask + store.  Warp (quad, grp) owns 64 of the row's 128 columns; the two partial (sum, sum of
      //      squared deviations) pairs meet in smem.  Y goes to registers first and is released at once (the next tile's G2_0
      //      overwrites it); U_i + V_j are added afterwards. ---------------------------------------------------------------
      {
        float v[64];
        const int cb = grp * 64;
        FU_PROF(c_yf, mbar_wait(y_full, lt & 1u));
        const long long c_ln0 = prof_on ? clock64() : 0;
        tc_fence_after();
        {
          uint32_t r0[32], r1[32];
          tmem_ld32_nowait(trow + ybase + (uint32_t)cb, r0);
          tmem_ld32_nowait(trow + ybase + (uint32_t)(cb + 32), r1);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 32; ++q) { v[q] = __uint_as_float(r0[q]); v[32 + q] = __uint_as_float(r1[q]); }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(y_empty);
        if (valid) {
          const float* pU = pP + 2 * ET_HID + cb;
          const float* pV = pQ + (ET_HID + C_Z) + cb;   // pQ points at +384: V sits at 896 = 384 + 512
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float x8[8], y8[8];
            ldg256(pU + q * 8, x8);
            ldg256(pV + q * 8, y8);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[q * 8 + e] += x8[e] + y8[e];
          }
        } else {
#pragma unroll
          for (int q = 0; q < 64; ++q) v[q] = 0.f;
        }
        // statistics: each half row gets its own (mean, M2) exactly (two local passes), then ONE exchange through smem and Chan's
        // combination  M2 = M2_a + M2_b + 32 (mean_a - mean_b)^2  — no cancellation, half the barrier traffic of two global passes.
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int q = 0; q < 64; q += 4) { s0 += v[q]; s1 += v[q + 1]; s2 += v[q + 2]; s3 += v[q + 3]; }
        const float mloc = ((s0 + s1) + (s2 + s3)) * (1.f / 64.f);
        float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
        for (int q = 0; q < 64; q += 4) {
          const float d0 = v[q] - mloc, d1 = v[q + 1] - mloc, d2 = v[q + 2] - mloc, d3 = v[q + 3] - mloc;
          q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1); q2 = fmaf(d2, d2, q2); q3 = fmaf(d3, d3, q3);
        }
        const float m2loc = (q0 + q1) + (q2 + q3);
        // stats buffer: [half 2][row 128][2] fp32 = 2 KB
        const uint32_t st_mine = stats + (uint32_t)((grp * 128 + row) * 8), st_other = stats + (uint32_t)(((grp ^ 1) * 128 + row) * 8);
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(st_mine), "f"(mloc), "f"(m2loc) : "memory");
        asm volatile("bar.sync %0, 64;" ::"r"(2 + quad) : "memory");     // the two warps of this quadrant
        float omean, om2;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(omean), "=f"(om2) : "r"(st_other) : "memory");
        asm volatile("bar.sync %0, 64;" ::"r"(2 + quad) : "memory");     // slots may be rewritten (next tile) only after both have read; cheap: the pair is aligned here
        const float mean = 0.5f * (mloc + omean);
        const float dm = mloc - omean;
        const float rstd = rsqrtf(((m2loc + om2) + 32.f * dm * dm) * (1.f / 128.f) + 1e-5f);
        {
          uint32_t hw[32], lw[32];
#pragma unroll
          for (int c0 = 0; c0 < 64; c0 += 8) {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.ln_g + cb + c0)), g1v = __ldg(reinterpret_cast<const float4*>(p.ln_g + cb + c0 + 4));
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.ln_b + cb + c0)), b1v = __ldg(reinterpret_cast<const float4*>(p.ln_b + cb + c0 + 4));
            const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1v.x, g1v.y, g1v.z, g1v.w};
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1v.x, b1v.y, b1v.z, b1v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
              split2_bf16(((v[c0 + 2 * e] - mean) * rstd * gg[2 * e] + bb[2 * e]) * emask,
                          ((v[c0 + 2 * e + 1] - mean) * rstd * gg[2 * e + 1] + bb[2 * e + 1]) * emask, hw[c0 / 2 + e], lw[c0 / 2 + e]);
          }
          const long long m_warp = m - lane;
          const uint32_t stage = stats + 2048u + (uint32_t)(warp - 2) * 4096u;
          warp_store_rows64(stage, hw, p.out_hi, m_warp, cb, p.E, lane);
          if (p.planes == 2) warp_store_rows64(stage, lw, p.out_lo, m_warp, cb, p.E, lane);
        }
        if (prof_on) c_ln += clock64() - c_ln0;
      }
    };
    RowCtx prev{};
    for (int tile = blockIdx.x; it < n_it; tile += gridDim.x, ++it) {
      const RowCtx cur = row_ctx(tile);
      const uint32_t t1b = (it & 1u) ? 384u : 0u, h2b = (it & 1u) ? 0u : 128u;
      epi1(cur, t1b, grp);
      if (it > 0) epi3(prev, it - 1);
      for (int c = grp + 2; c < 6; c += 2) epi1(cur, t1b, c);
      // ---- epi2: h2 chunks of this group's parity, in place over H2 -----------------------------------------------------
      FU_PROF(c_h2f, mbar_wait(h2_full, it & 1u));
      tc_fence_after();
      for (int c = grp; c < 6; c += 2) {
        float v[64];
        const int col0 = c * 64;
        const uint32_t tchunk = trow + h2b + (uint32_t)(c * 64);
        {
          uint32_t r0[32], r1[32];
          tmem_ld32_nowait(tchunk, r0);
          tmem_ld32_nowait(tchunk + 32u, r1);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 ba = __ldg(reinterpret_cast<const float4*>(p.b2 + col0 + q * 4));
            const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b2 + col0 + 32 + q * 4));
            v[q * 4 + 0] = fmaxf(__uint_as_float(r0[q * 4 + 0]) + ba.x, 0.f); v[q * 4 + 1] = fmaxf(__uint_as_float(r0[q * 4 + 1]) + ba.y, 0.f);
            v[q * 4 + 2] = fmaxf(__uint_as_float(r0[q * 4 + 2]) + ba.z, 0.f); v[q * 4 + 3] = fmaxf(__uint_as_float(r0[q * 4 + 3]) + ba.w, 0.f);
            v[32 + q * 4 + 0] = fmaxf(__uint_as_float(r1[q * 4 + 0]) + bb.x, 0.f); v[32 + q * 4 + 1] = fmaxf(__uint_as_float(r1[q * 4 + 1]) + bb.y, 0.f);
            v[32 + q * 4 + 2] = fmaxf(__uint_as_float(r1[q * 4 + 2]) + bb.z, 0.f); v[32 + q * 4 + 3] = fmaxf(__uint_as_float(r1[q * 4 + 3]) + bb.w, 0.f);
          }
        }
        write_a(tchunk, v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a2_full(c));
      }
      prev = cur;
    }
    if (it > 0) epi3(prev, it - 1);
    if (prof_on && lane == 0) { p.prof[16] = clock64() - c_start; p.prof[17] = c_t1f; p.prof[19] = c_h2f; p.prof[20] = c_yf; p.prof[21] = c_ln; }
  }
  tc_fence_before();
  __syncthreads();
  if (CL) cluster_sync_all();      // no CTA leaves while its peer may still signal its barriers
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Fused edge-embedder tail (model/score_network.py:64-72 edge_embedder layers 2..4):
//     z = LayerNorm(relu(h0·W2^T + b2)·W4^T + b4) · mask        per 128-edge tile, h1 never leaves the SM.
//   Both weight matrices (hi/lo, 128 KB) stay resident in shared memory for the whole persistent CTA; the only stream is the
//   h0 tile (TMA, 64 KB).  TMEM is split in two 256-column halves used by alternating tiles: H1 (128 fp32 columns, rewritten
//   in place as the bf16 hi/lo A image of the second GEMM) | Y (128 fp32 columns).  The MMA warp issues the first GEMM of
//   tile t+1 before the second GEMM of tile t, so the tensor pipe works under the epilogue of the previous tile.
//   warps: 0 = TMA producer, 1 = MMA issuer (+TMEM alloc), 2..9 = epilogue (quadrant = warp & 3, column half = (warp-2) >> 2).
// ------------------------------------------------------------------------------------------------------------------
constexpr int EF_THREADS = 320;
constexpr size_t EF_SMEM_BYTES = 12 * (size_t)TC_PLANE_BYTES + 512 + 2048 + 8 * 4096;   // dynamic smem starts 1024-aligned (no static smem)

struct EmbedFusedParams {
  int E, planes, nres, num_tiles;
  const float* b2; const float* b4; const float* ln_g; const float* ln_b; const float* res_mask;
  __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;
};

__global__ void __launch_bounds__(EF_THREADS, 1)
tc_embed_fused_kernel(const __grid_constant__ CUtensorMap mAh, const __grid_constant__ CUtensorMap mAl,
                      const __grid_constant__ CUtensorMap mW2h, const __grid_constant__ CUtensorMap mW2l,
                      const __grid_constant__ CUtensorMap mW4h, const __grid_constant__ CUtensorMap mW4l, const EmbedFusedParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t PL = TC_PLANE_BYTES;   // 16 KB: 128 rows x 64 bf16
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto wbuf = [&](int l, int kb, int pl) { return base + (uint32_t)((l * 2 + kb) * 2 + pl) * PL; };
  auto abuf = [&](int kb, int pl) { return base + 8u * PL + (uint32_t)(kb * 2 + pl) * PL; };
  const uint32_t bar0 = base + 12u * PL;
  const uint32_t w_full = bar0, a_full = bar0 + 8u, a_empty = bar0 + 16u;
  auto h1_full = [&](uint32_t x) { return bar0 + 8u * (3 + x); };
  auto c_full = [&](uint32_t x) { return bar0 + 8u * (5 + x); };
  auto y_full = [&](uint32_t x) { return bar0 + 8u * (7 + x); };
  auto t_empty = [&](uint32_t x) { return bar0 + 8u * (9 + x); };
  const uint32_t tmem_ptr_addr = bar0 + 8u * 11;
  const uint32_t stats = bar0 + 512u;     // LayerNorm partial statistics [2][128 rows][2] fp32 = 2 KB, then 8 x 4 KB store staging

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(w_full, 1); mbar_init(a_full, 1); mbar_init(a_empty, 1);
    for (uint32_t x = 0; x < 2; ++x) { mbar_init(h1_full(x), 1); mbar_init(c_full(x), 8); mbar_init(y_full(x), 1); mbar_init(t_empty(x), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&mAh); tma_prefetch_desc(&mW2h); tma_prefetch_desc(&mW4h);
    if (p.planes == 2) { tma_prefetch_desc(&mAl); tma_prefetch_desc(&mW2l); tma_prefetch_desc(&mW4l); }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));
  const uint32_t n_my = blockIdx.x < (unsigned)p.num_tiles ? (uint32_t)((p.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1) : 0u;

  if (warp == 0) {
    // ============================================ TMA producer ============================================
    if (elect_one()) {
      mbar_expect_tx(w_full, (uint32_t)p.planes * 4 * PL);
      for (int kb = 0; kb < 2; ++kb) {
        tma_load_2d(wbuf(0, kb, 0), &mW2h, w_full, kb * 64, 0);
        tma_load_2d(wbuf(1, kb, 0), &mW4h, w_full, kb * 64, 0);
        if (p.planes == 2) {
          tma_load_2d(wbuf(0, kb, 1), &mW2l, w_full, kb * 64, 0);
          tma_load_2d(wbuf(1, kb, 1), &mW4l, w_full, kb * 64, 0);
        }
      }
    }
    __syncwarp();
    for (uint32_t it = 0; it < n_my; ++it) {
      const int m0 = ((int)blockIdx.x + (int)it * (int)gridDim.x) * TC_BM;
      mbar_wait(a_empty, (it & 1u) ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(a_full, (uint32_t)p.planes * 2 * PL);
        for (int kb = 0; kb < 2; ++kb) {
          tma_load_2d(abuf(kb, 0), &mAh, a_full, kb * 64, m0);
          if (p.planes == 2) tma_load_2d(abuf(kb, 1), &mAl, a_full, kb * 64, m0);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ============================================ MMA issuer ============================================
    const uint32_t idesc = make_idesc_bf16(TC_BM, 128);
    const bool x3 = p.planes == 2;
    mbar_wait(w_full, 0u);
    tc_fence_after();
    // first GEMM of tile k: H1[k&1] = h0 · W2^T (A and B from shared memory)
    auto g1 = [&](uint32_t k) {
      const uint32_t x = k & 1u;
      mbar_wait(t_empty(x), ((k >> 1) & 1u) ^ 1u);     // tile k-2 (same TMEM half) fully read by its LayerNorm
      mbar_wait(a_full, k & 1u);
      tc_fence_after();
      if (elect_one()) {
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t dAH = make_sw128_desc(abuf(kb, 0)), dAL = make_sw128_desc(abuf(kb, 1));
          const uint64_t dBH = make_sw128_desc(wbuf(0, kb, 0)), dBL = make_sw128_desc(wbuf(0, kb, 1));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_kstep(tmem_base + 256u * x, dAH + 2u * ks, dAL + 2u * ks, dBH + 2u * ks, dBL + 2u * ks, idesc, (kb == 0 && ks == 0) ? 0u : 1u, x3);
        }
        tc_commit(a_empty);
        tc_commit(h1_full(x));
      }
      __syncwarp();
    };
    if (n_my > 0) g1(0);
    for (uint32_t it = 0; it < n_my; ++it) {
      if (it + 1 < n_my) g1(it + 1);
      const uint32_t x = it & 1u;
      mbar_wait(c_full(x), (it >> 1) & 1u);              // epilogue wrote relu(h1) as bf16 hi/lo images over H1[x]
      tc_fence_after();
      if (elect_one()) {
        for (int c = 0; c < 2; ++c) {
          const uint32_t a = tmem_base + 256u * x + 64u * c;
          const uint64_t dBH = make_sw128_desc(wbuf(1, c, 0)), dBL = make_sw128_desc(wbuf(1, c, 1));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_kstep_ts(tmem_base + 256u * x + 128u, a + 8u * ks, a + 32u + 8u * ks, dBH + 2u * ks, dBL + 2u * ks, idesc, (c == 0 && ks == 0) ? 0u : 1u, x3);
        }
        tc_commit(y_full(x));
      }
      __syncwarp();
    }
  } else {
    // ============================================ epilogue warps 2..9 ============================================
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int cb = half * 64;
    const uint32_t stage = stats + 2048u + (uint32_t)(warp - 2) * 4096u;     // this warp's store-staging buffer
    // Software pipeline: the activation of tile it+1 is converted before the LayerNorm of tile it, so the second GEMM of tile it+1
    // runs on the tensor pipe while these warps normalise and store tile it.
    auto convert = [&](uint32_t it) {
      const uint32_t x = it & 1u, ph = (it >> 1) & 1u;
      // ---- relu(H1 + b2) -> bf16 hi/lo A image over this warp's own 64 columns (hi words [0,32), lo words [32,64)) ----
      mbar_wait(h1_full(x), ph);
      tc_fence_after();
      {
        const uint32_t tchunk = trow + 256u * x + (uint32_t)cb;
        uint32_t r0[32], r1[32];
        tmem_ld32_nowait(tchunk, r0);
        tmem_ld32_nowait(tchunk + 32u, r1);
        tmem_ld_wait();
        uint32_t h[16], l[16];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 ba = __ldg(reinterpret_cast<const float4*>(p.b2 + cb + q * 4));
          split2_bf16(fmaxf(__uint_as_float(r0[q * 4 + 0]) + ba.x, 0.f), fmaxf(__uint_as_float(r0[q * 4 + 1]) + ba.y, 0.f), h[2 * q], l[2 * q]);
          split2_bf16(fmaxf(__uint_as_float(r0[q * 4 + 2]) + ba.z, 0.f), fmaxf(__uint_as_float(r0[q * 4 + 3]) + ba.w, 0.f), h[2 * q + 1], l[2 * q + 1]);
        }
        tmem_st16(tchunk, h);
        if (p.planes == 2) tmem_st16(tchunk + 32u, l);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b2 + cb + 32 + q * 4));
          split2_bf16(fmaxf(__uint_as_float(r1[q * 4 + 0]) + bb.x, 0.f), fmaxf(__uint_as_float(r1[q * 4 + 1]) + bb.y, 0.f), h[2 * q], l[2 * q]);
          split2_bf16(fmaxf(__uint_as_float(r1[q * 4 + 2]) + bb.z, 0.f), fmaxf(__uint_as_float(r1[q * 4 + 3]) + bb.w, 0.f), h[2 * q + 1], l[2 * q + 1]);
        }
        tmem_st16(tchunk + 16u, h);
        if (p.planes == 2) tmem_st16(tchunk + 48u, l);
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(c_full(x));
    };
    auto layernorm = [&](uint32_t it) {
      const uint32_t x = it & 1u, ph = (it >> 1) & 1u;
      const long long m = (long long)((int)blockIdx.x + (int)it * (int)gridDim.x) * TC_BM + row;
      const bool valid = m < p.E;
      float emask = 0.f;
      if (valid) {
        const long long nn = (long long)p.nres * p.nres;
        const long long b = m / nn;
        const int rem = (int)(m - b * nn);
        const int ri = rem / p.nres, rj = rem - ri * p.nres;
        emask = p.res_mask[b * p.nres + ri] * p.res_mask[b * p.nres + rj];
      }
      // ---- LayerNorm(Y + b4) * mask -> z planes.  Two warps per quadrant split the 128 columns; partial statistics meet in smem ----
      float v[64];
      mbar_wait(y_full(x), ph);
      tc_fence_after();
      {
        uint32_t r0[32], r1[32];
        tmem_ld32_nowait(trow + 256u * x + 128u + (uint32_t)cb, r0);
        tmem_ld32_nowait(trow + 256u * x + 128u + (uint32_t)(cb + 32), r1);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 ba = __ldg(reinterpret_cast<const float4*>(p.b4 + cb + q * 4)), bb = __ldg(reinterpret_cast<const float4*>(p.b4 + cb + 32 + q * 4));
          v[q * 4 + 0] = __uint_as_float(r0[q * 4 + 0]) + ba.x; v[q * 4 + 1] = __uint_as_float(r0[q * 4 + 1]) + ba.y;
          v[q * 4 + 2] = __uint_as_float(r0[q * 4 + 2]) + ba.z; v[q * 4 + 3] = __uint_as_float(r0[q * 4 + 3]) + ba.w;
          v[32 + q * 4 + 0] = __uint_as_float(r1[q * 4 + 0]) + bb.x; v[32 + q * 4 + 1] = __uint_as_float(r1[q * 4 + 1]) + bb.y;
          v[32 + q * 4 + 2] = __uint_as_float(r1[q * 4 + 2]) + bb.z; v[32 + q * 4 + 3] = __uint_as_float(r1[q * 4 + 3]) + bb.w;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty(x));            // this TMEM half may be overwritten by tile it+2
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int q = 0; q < 64; q += 4) { s0 += v[q]; s1 += v[q + 1]; s2 += v[q + 2]; s3 += v[q + 3]; }
      const float psum = (s0 + s1) + (s2 + s3);
      const uint32_t st_mine = stats + (uint32_t)((half * 128 + row) * 8), st_other = stats + (uint32_t)(((half ^ 1) * 128 + row) * 8);
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(st_mine), "f"(psum) : "memory");
      asm volatile("bar.sync %0, 64;" ::"r"(2 + quad) : "memory");
      float osum;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(osum) : "r"(st_other) : "memory");
      const float mean = (psum + osum) * (1.f / 128.f);
      float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
      for (int q = 0; q < 64; q += 4) {
        const float d0 = v[q] - mean, d1 = v[q + 1] - mean, d2 = v[q + 2] - mean, d3 = v[q + 3] - mean;
        q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1); q2 = fmaf(d2, d2, q2); q3 = fmaf(d3, d3, q3);
      }
      const float pvar = (q0 + q1) + (q2 + q3);
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(st_mine + 4), "f"(pvar) : "memory");
      asm volatile("bar.sync %0, 64;" ::"r"(2 + quad) : "memory");
      float ovar;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(ovar) : "r"(st_other + 4) : "memory");
      const float rstd = rsqrtf((pvar + ovar) * (1.f / 128.f) + 1e-5f);
      {
        uint32_t hw[32], lw[32];
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 8) {
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.ln_g + cb + c0)), g1v = __ldg(reinterpret_cast<const float4*>(p.ln_g + cb + c0 + 4));
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.ln_b + cb + c0)), b1v = __ldg(reinterpret_cast<const float4*>(p.ln_b + cb + c0 + 4));
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1v.x, g1v.y, g1v.z, g1v.w};
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1v.x, b1v.y, b1v.z, b1v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e)
            split2_bf16(((v[c0 + 2 * e] - mean) * rstd * gg[2 * e] + bb[2 * e]) * emask,
                        ((v[c0 + 2 * e + 1] - mean) * rstd * gg[2 * e + 1] + bb[2 * e + 1]) * emask, hw[c0 / 2 + e], lw[c0 / 2 + e]);
        }
        const long long m_warp = m - lane;
        warp_store_rows64(stage, hw, p.out_hi, m_warp, cb, p.E, lane);
        if (p.planes == 2) warp_store_rows64(stage, lw, p.out_lo, m_warp, cb, p.E, lane);
      }
    };
    if (n_my > 0) convert(0);
    for (uint32_t it = 0; it < n_my; ++it) {
      if (it + 1 < n_my) convert(it + 1);
      layernorm(it);
      // the stats slots are rewritten by this pair only after its next bar.sync pair, i.e. after both have read them
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
}

constexpr size_t TC_SMEM_BYTES = 1024 + (size_t)(TC_SA + TC_SB) * 2 * TC_PLANE_BYTES + 1024 + 8 * 4096;   // rings | barriers (1 KB) | 8 x 4 KB epilogue staging tiles

// fp32 [M, ld] (first K columns) -> dense bf16 hi/lo planes [M, K]  (A operand of a node-path tensor-core linear)
__global__ void split_planes_kernel(const float* __restrict__ x, int ld, long long M, int K, __nv_bfloat16* __restrict__ hi,
                                    __nv_bfloat16* __restrict__ lo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one float4
  const int k4 = K / 4;
  if (i >= M * k4) return;
  const long long m = i / k4;
  const int k = (int)(i - m * k4) * 4;
  const float4 v = *reinterpret_cast<const float4*>(x + m * ld + k);
  uint32_t h0, l0, h1, l1;
  split2_bf16(v.x, v.y, h0, l0);
  split2_bf16(v.z, v.w, h1, l1);
  *reinterpret_cast<uint2*>(hi + m * K + k) = make_uint2(h0, h1);
  if (lo) *reinterpret_cast<uint2*>(lo + m * K + k) = make_uint2(l0, l1);
}

// fp32 [M, ld] (first Kv columns valid) -> bf16 hi/lo planes [M, Kp], zero beyond Kv  (attention probabilities: A operand of a·v)
__global__ void split_pad_planes_kernel(const float* __restrict__ x, int ld, long long M, int Kv, int Kp, __nv_bfloat16* __restrict__ hi,
                                        __nv_bfloat16* __restrict__ lo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // four columns
  const int k4 = Kp / 4;
  if (i >= M * k4) return;
  const long long m = i / k4;
  const int k = (int)(i - m * k4) * 4;
  float v[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) v[e] = (k + e < Kv) ? x[m * ld + k + e] : 0.f;
  uint32_t h0, l0, h1, l1;
  split2_bf16(v[0], v[1], h0, l0);
  split2_bf16(v[2], v[3], h1, l1);
  *reinterpret_cast<uint2*>(hi + m * Kp + k) = make_uint2(h0, h1);
  *reinterpret_cast<uint2*>(lo + m * Kp + k) = make_uint2(l0, l1);
}

// Values transposed per (sample, head) into K-major planes for the probability·value GEMMs:
//   vt[(bh*dhp + c), j] = x[(b*N + j)*ld + col0 + h*hstride + c]  for c < dh, j < N; zero elsewhere (c < dhp, j < Kp).
__global__ void vt_planes_kernel(const float* __restrict__ x, int ld, int col0, int hstride, int nheads, int dh, int dhp, int N, int Kp,
                                 __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  __shared__ float tile[64][33];                  // 64 edges x 32 channels per block; Kp is a multiple of 64
  const int bh = blockIdx.z, b = bh / nheads, hh = bh - b * nheads;
  const int j0 = blockIdx.x * 64, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
  for (int r = ty; r < 64; r += 8) {
    const int j = j0 + r;
    tile[r][tx] = (j < N && c0 + tx < dh) ? x[((long long)b * N + j) * ld + col0 + hh * hstride + c0 + tx] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {              // channel c0 + r: lane tx writes the edge pair (j0 + 2 tx, j0 + 2 tx + 1)
    uint32_t h2, l2;
    split2_bf16(tile[2 * tx][r], tile[2 * tx + 1][r], h2, l2);
    const long long o = ((long long)bh * dhp + c0 + r) * Kp + j0 + 2 * tx;
    *reinterpret_cast<uint32_t*>(hi + o) = h2;
    *reinterpret_cast<uint32_t*>(lo + o) = l2;
  }
}

// Per-head zero-padded split of the sequence transformer's q and k:  out[m, g*dhp + c] = x[m*ld + g*dh + c] (c < dh), 0 otherwise,
// for the ngroups = 2*heads consecutive dh-wide groups (q heads then k heads).  K-blocks of 64 then never straddle two heads.
__global__ void split_heads_pad_kernel(const float* __restrict__ x, int ld, long long M, int ngroups, int dh, int dhp,
                                       __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // four columns
  const int W = ngroups * dhp, k4 = W / 4;
  if (i >= M * k4) return;
  const long long m = i / k4;
  const int k = (int)(i - m * k4) * 4, g = k / dhp, c = k - g * dhp;      // dh and dhp are multiples of 4
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < dh) v = *reinterpret_cast<const float4*>(x + m * ld + g * dh + c);
  uint32_t h0, l0, h1, l1;
  split2_bf16(v.x, v.y, h0, l0);
  split2_bf16(v.z, v.w, h1, l1);
  *reinterpret_cast<uint2*>(hi + m * W + k) = make_uint2(h0, h1);
  *reinterpret_cast<uint2*>(lo + m * W + k) = make_uint2(l0, l1);
}

// Operand planes of the extended IPA logits GEMM (model/ipa_pytorch.py:376-417).  Per head the contraction runs over 384 columns:
//   [0,256)    scalar q_h / k_h
//   [256,384)  point term  -½γ_h Σ|qp_i - kp_j|² = γ_h qp_i·kp_j - ½γ_h|kp_j|² - ½γ_h|qp_i|²  (the last addend is constant over j: it cancels in
//              the softmax and is dropped).  The GEMM scales its accumulator by alpha, so the query side carries s = γ_h/alpha.  With
//              x = x0 + x1 + x2 (three bf16 pieces, 24 bits) and the GEMM's split product (hi·hi + hi·lo + lo·hi) the columns are
//                 +0..23  A (a0, a1)  B (k0, k1)      a0k0 + a0k1 + a1k0        (a = s·qp)
//                +24..47  A (a0, 0)   B (k2, 0)       a0k2
//                +48..71  A (a2, 0)   B (k0, 0)       a2k0
//                +72..95  A (a1, 0)   B (k1, 0)       a1k1
//                +96..98  A (1, 0)    B (n0|n1|n2, 0) n = -(s/2)|kp_j|²  exact to 24 bits
//              i.e. every product term down to 2^-24, the accuracy of the fp32 reference loop.
// Layout [R, 6144]: query side at column h*384, key side at 3072 + h*384.
__global__ void __launch_bounds__(256) ipa_qkx_planes_kernel(const float* __restrict__ proj, const float* __restrict__ qp, const float* __restrict__ kp,
                                                             const float* __restrict__ gamma, float inv_alpha, __nv_bfloat16* __restrict__ hi,
                                                             __nv_bfloat16* __restrict__ lo) {
  const long long row = blockIdx.x;
  const int tid = threadIdx.x;
  const float* pr = proj + row * PROJ_ALL;
  __nv_bfloat16* oh = hi + row * TC_QKV; __nv_bfloat16* ol = lo + row * TC_QKV;
#pragma unroll
  for (int h = 0; h < H; ++h) {
    __nv_bfloat16 a, b2;
    split_bf16(pr[h * C_HID + tid], a, b2);
    oh[h * 384 + tid] = a; ol[h * 384 + tid] = b2;
    split_bf16(pr[PROJ_Q + h * 2 * C_HID + tid], a, b2);
    oh[3072 + h * 384 + tid] = a; ol[3072 + h * 384 + tid] = b2;
  }
  const __nv_bfloat16 zero = __float2bfloat16(0.f);
  for (int idx = tid; idx < 2 * H * 128; idx += 256) {
    const int side = idx >> 10, h = (idx >> 7) & 7, e = idx & 127;      // side 0: query (A operand), 1: key (B operand)
    const float s = gamma[h] * inv_alpha;
    __nv_bfloat16 vh = zero, vl = zero;
    if (e < 96) {
      const int grp = e / 24, c = e - grp * 24;
      const float x = side == 0 ? s * qp[row * (H * PQ * 3) + h * (PQ * 3) + c] : kp[row * (H * PQ * 3) + h * (PQ * 3) + c];
      const __nv_bfloat16 x0 = __float2bfloat16(x);
      const float r1 = x - __bfloat162float(x0);
      const __nv_bfloat16 x1 = __float2bfloat16(r1);
      const __nv_bfloat16 x2 = __float2bfloat16(r1 - __bfloat162float(x1));
      if (side == 0) {
        if (grp == 0) { vh = x0; vl = x1; } else if (grp == 1) vh = x0; else if (grp == 2) vh = x2; else vh = x1;
      } else {
        if (grp == 0) { vh = x0; vl = x1; } else if (grp == 1) vh = x2; else if (grp == 2) vh = x0; else vh = x1;
      }
    } else if (e < 99) {
      if (side == 0) {
        vh = __float2bfloat16(1.f);
      } else {
        float n2 = 0.f;
#pragma unroll
        for (int c = 0; c < PQ * 3; ++c) { const float k = kp[row * (H * PQ * 3) + h * (PQ * 3) + c]; n2 = fmaf(k, k, n2); }
        const float n = -0.5f * s * n2;
        const __nv_bfloat16 n0 = __float2bfloat16(n);
        const float r1 = n - __bfloat162float(n0);
        const __nv_bfloat16 n1 = __float2bfloat16(r1);
        const __nv_bfloat16 n2b = __float2bfloat16(r1 - __bfloat162float(n1));
        vh = e == 96 ? n0 : (e == 97 ? n1 : n2b);
      }
    }
    oh[side * 3072 + h * 384 + 256 + e] = vh; ol[side * 3072 + h * 384 + 256 + e] = vl;
  }
}

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
static int g_tc_sa = 0;             // FD_TC_SA=2|3|4 overrides the activation-ring depth of the launches that set one (A/B runs)
static int g_tc_eg = 0;             // epilogue groups of tc_gemm_kernel: 0 = per launch (TcGemmParams::eg); FD_TC_EG=1|2 forces one variant (A/B runs)

static long long* g_tc_prof = nullptr;   // device [32]; set by fd_debug_tc_profile
// ------------------------------------------------------------------------------------------------------------------
// IPA edge pass, one kernel per block (model/ipa_pytorch.py:376-432): one CTA per query residue (b,i), 8 warps, mma.sync m16n8k16
// with operand fragments built straight from global memory (no shared-memory staging of z):
//   pass A  pair bias  pb[j][h] = z[j,:]·Wb[h,:].  A = z tile (16 edges x 128 channels); the contraction order over channels is free, so
//           the q-th 16-byte load of lane (g,t) takes channels 32q + 8t + {0..7} of rows g and g+8 (a warp instruction reads 64 contiguous
//           bytes of 8 rows, every sector fully used) and k-step ks = 2q + e uses channels 32q + 8t + 4e + {0..3}; the Wb fragments use
//           the same channel permutation.  4-term split product (hi·hi + hi·lo + lo·hi + lo·lo).
//   logits  lg[h][j] = L_in + sqrt(1/3)(pb + bb) + mask, where L_in = qk - ½γ_h Σ_p|qp_i - kp_j|² (up to a per-i constant) comes from the
//           extended-K logits GEMM (tc_ipa_logits);  softmax over j;  probabilities -> L (for a·v) and smem
//   pass B  zbar[h][c] = Σ_j a[h][j] z[j][c].  A = probabilities (heads as rows 0..7), B needs pairs along j: lane (g,t) loads 16 bytes
//           (8 channels) of rows 2t, 2t+1, 2t+8, 2t+9 and interleaves them with byte permutes; warp w owns channels [64(w&1), +64)
//           and the (w>>1)-th quarter of the edges; the four partial sums meet in shared memory.  The second read of z hits L2.
//   zbar -> global; o_pair = Wd·zbar + bd is one block-diagonal tensor-core linear over all heads (caller)
// The separate pair-bias GEMM + attention kernel (ipa_edge2) remain as the cross-check path (FD_IPA_EDGE2=1).
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}

constexpr int E3_ASTRIDE = 528;   // bytes per head row of the probability planes (<= 256 bf16 = 512 B, + 16 B: conflict-free 4-byte reads)
inline size_t ipa_edge3_smem(int Np) {
  return (size_t)H * Np * 4 + 2 * (size_t)H * E3_ASTRIDE + 4 * (size_t)H * C_Z * 4 + 2 * 8 * 32 * 8;
}

template <int PLANES>
__global__ void __launch_bounds__(256, 3) ipa_edge3_kernel(
    const __nv_bfloat16* __restrict__ z_hi, const __nv_bfloat16* __restrict__ z_lo, float* __restrict__ L, const float* __restrict__ res_mask,
    const float* __restrict__ Wb, const float* __restrict__ bb, float* __restrict__ zbar, __nv_bfloat16* __restrict__ at_hi,
    __nv_bfloat16* __restrict__ at_lo, int Kp, int write_L, int N, int Np) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int NT = (N + 15) / 16;
  float* lg = reinterpret_cast<float*>(smem_raw);                                 // [H][Np] logits -> probabilities
  uint8_t* ap = smem_raw + (size_t)H * Np * 4;                                    // probability planes [2][H] rows of E3_ASTRIDE bytes
  float* zbp = reinterpret_cast<float*>(ap + 2 * H * E3_ASTRIDE);                 // [4 edge quarters][H][128] partial zbar
  uint2* wbs = reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(zbp) + 4 * H * C_Z * 4);   // Wb fragments [plane][ks][lane]

  const int i = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const long long rowi = (long long)b * N + i;
  const long long zrow = rowi * N;                // first edge row of this residue in the z planes
  {  // Wb fragments of k-step ks = warp (all warps together cover the 8 k-steps): channels 32(ks>>1) + 8t + 4(ks&1) + {0,1} and {2,3} of head g
    const float4 w4 = *reinterpret_cast<const float4*>(Wb + g * C_Z + 32 * (warp >> 1) + 8 * t + 4 * (warp & 1));
    uint32_t h0, l0, h1, l1;
    split2_bf16(w4.x, w4.y, h0, l0);
    split2_bf16(w4.z, w4.w, h1, l1);
    wbs[(0 * 8 + warp) * 32 + lane] = make_uint2(h0, h1);
    wbs[(1 * 8 + warp) * 32 + lane] = make_uint2(l0, l1);
  }
  const float mi = res_mask[rowi];
  // ---- phase 1: logits from the GEMM (scalar q·k and the point-distance term, see tc_ipa_logits) + mask -> lg ----
  for (int idx = tid; idx < H * Np; idx += 256) {
    const int h = idx / Np, j = idx - h * Np;
    lg[idx] = j < N ? L[(((long long)b * H + h) * N + i) * Np + j] + 1e5f * (mi * res_mask[(long long)b * N + j] - 1.f) : 0.f;
  }
  __syncthreads();
  // ---- pass A: pair bias on the tensor cores, added into lg ----
  for (int jt = warp; jt < NT; jt += 8) {
    const int r0 = jt * 16 + g, r1 = r0 + 8;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int half = 0; half < 2; ++half) {     // loads q = 2 half, 2 half + 1: k-steps 4 half .. 4 half + 3
      uint4 x0h[2], x1h[2], x0l[2], x1l[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const long long o0 = (zrow + r0) * C_Z + 32 * (2 * half + q) + 8 * t, o1 = (zrow + r1) * C_Z + 32 * (2 * half + q) + 8 * t;
        x0h[q] = r0 < N ? *reinterpret_cast<const uint4*>(z_hi + o0) : make_uint4(0u, 0u, 0u, 0u);
        x1h[q] = r1 < N ? *reinterpret_cast<const uint4*>(z_hi + o1) : make_uint4(0u, 0u, 0u, 0u);
        if (PLANES == 2) {
          x0l[q] = r0 < N ? *reinterpret_cast<const uint4*>(z_lo + o0) : make_uint4(0u, 0u, 0u, 0u);
          x1l[q] = r1 < N ? *reinterpret_cast<const uint4*>(z_lo + o1) : make_uint4(0u, 0u, 0u, 0u);
        }
      }
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const int ks = half * 4 + k4;
        const uint2 bh = wbs[(0 * 8 + ks) * 32 + lane], bl = wbs[(1 * 8 + ks) * 32 + lane];
        const uint4 v0h = x0h[k4 >> 1], v1h = x1h[k4 >> 1];
        const uint32_t a0 = (k4 & 1) ? v0h.z : v0h.x, a2 = (k4 & 1) ? v0h.w : v0h.y;
        const uint32_t a1 = (k4 & 1) ? v1h.z : v1h.x, a3 = (k4 & 1) ? v1h.w : v1h.y;
        mma_bf16_16816(acc, a0, a1, a2, a3, bh.x, bh.y);
        mma_bf16_16816(acc, a0, a1, a2, a3, bl.x, bl.y);
        if (PLANES == 2) {
          const uint4 v0l = x0l[k4 >> 1], v1l = x1l[k4 >> 1];
          const uint32_t c0 = (k4 & 1) ? v0l.z : v0l.x, c2 = (k4 & 1) ? v0l.w : v0l.y;
          const uint32_t c1 = (k4 & 1) ? v1l.z : v1l.x, c3 = (k4 & 1) ? v1l.w : v1l.y;
          mma_bf16_16816(acc, c0, c1, c2, c3, bh.x, bh.y);
          mma_bf16_16816(acc, c0, c1, c2, c3, bl.x, bl.y);
        }
      }
    }
    const int h0 = 2 * t, h1 = 2 * t + 1;
    const float b0 = bb[h0], b1 = bb[h1];
    if (r0 < N) { lg[h0 * Np + r0] += 0.57735026918962576f * (acc[0] + b0); lg[h1 * Np + r0] += 0.57735026918962576f * (acc[1] + b1); }
    if (r1 < N) { lg[h0 * Np + r1] += 0.57735026918962576f * (acc[2] + b0); lg[h1 * Np + r1] += 0.57735026918962576f * (acc[3] + b1); }
  }
  __syncthreads();
  // ---- softmax over j (warp <-> head); probabilities to L (fp32, for a·v) and to the bf16 hi/lo planes of pass B ----
  {
    float* row = lg + warp * Np;
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) mx = fmaxf(mx, row[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) { const float e = expf(row[j] - mx); row[j] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    // two adjacent edges per lane: probabilities go (a) to the smem planes of pass B, (b) to the global [B*H*N, Kp] bf16 hi/lo planes the
    // a·v GEMMs read as their A operand (zero-padded to Kp), (c) as fp32 to L only when the debug tap wants them
    float* Lrow = L + (((long long)b * H + warp) * N + i) * Np;
    uint32_t* ph = reinterpret_cast<uint32_t*>(ap + warp * E3_ASTRIDE);
    uint32_t* pl = reinterpret_cast<uint32_t*>(ap + (H + warp) * E3_ASTRIDE);
    uint32_t* gh = reinterpret_cast<uint32_t*>(at_hi + (((long long)b * H + warp) * N + i) * Kp);
    uint32_t* gl = reinterpret_cast<uint32_t*>(at_lo + (((long long)b * H + warp) * N + i) * Kp);
    for (int j = 2 * lane; j < Kp; j += 64) {
      const float a0 = j < N ? row[j] * inv : 0.f, a1 = j + 1 < N ? row[j + 1] * inv : 0.f;
      uint32_t hh, ll;
      split2_bf16(a0, a1, hh, ll);
      if (j < NT * 16) { ph[j >> 1] = hh; pl[j >> 1] = ll; }
      gh[j >> 1] = hh; gl[j >> 1] = ll;
      if (write_L && j < Np) *reinterpret_cast<float2*>(Lrow + j) = make_float2(a0, a1);
    }
  }
  __syncthreads();
  // ---- pass B: zbar partials.  n-tile x (0..7), column n  <->  channel 64 ch + 8 n + x ----
  {
    const int ch = warp & 1, kq = warp >> 1;
    const int KQ = (NT + 3) / 4;
    const int ks_end = (kq + 1) * KQ < NT ? (kq + 1) * KQ : NT;
    float acc[8][4];
#pragma unroll
    for (int x = 0; x < 8; ++x) { acc[x][0] = 0.f; acc[x][1] = 0.f; acc[x][2] = 0.f; acc[x][3] = 0.f; }
    const int cb = 64 * ch + 8 * g;
    for (int ks = kq * KQ; ks < ks_end; ++ks) {
      const int ja = ks * 16 + 2 * t;
      uint4 wh[4], wl[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = ja + (q & 1) + (q >> 1) * 8;     // rows 2t, 2t+1, 2t+8, 2t+9
        const long long o = (zrow + j) * C_Z + cb;
        wh[q] = j < N ? *reinterpret_cast<const uint4*>(z_hi + o) : make_uint4(0u, 0u, 0u, 0u);
        if (PLANES == 2) wl[q] = j < N ? *reinterpret_cast<const uint4*>(z_lo + o) : make_uint4(0u, 0u, 0u, 0u);
      }
      // probabilities: head g, edges ja, ja+1 (a0) and ja+8, ja+9 (a2); MMA rows 8..15 are zero
      const uint32_t ah0 = *reinterpret_cast<const uint32_t*>(ap + g * E3_ASTRIDE + ja * 2);
      const uint32_t ah2 = *reinterpret_cast<const uint32_t*>(ap + g * E3_ASTRIDE + (ja + 8) * 2);
      const uint32_t al0 = *reinterpret_cast<const uint32_t*>(ap + (H + g) * E3_ASTRIDE + ja * 2);
      const uint32_t al2 = *reinterpret_cast<const uint32_t*>(ap + (H + g) * E3_ASTRIDE + (ja + 8) * 2);
      const uint32_t w0h[4] = {wh[0].x, wh[0].y, wh[0].z, wh[0].w}, w1h[4] = {wh[1].x, wh[1].y, wh[1].z, wh[1].w};
      const uint32_t w2h[4] = {wh[2].x, wh[2].y, wh[2].z, wh[2].w}, w3h[4] = {wh[3].x, wh[3].y, wh[3].z, wh[3].w};
#pragma unroll
      for (int x = 0; x < 8; ++x) {
        const uint32_t sel = (x & 1) ? 0x7632u : 0x5410u;
        const uint32_t b0 = prmt(w0h[x >> 1], w1h[x >> 1], sel), b1 = prmt(w2h[x >> 1], w3h[x >> 1], sel);
        mma_bf16_16816(acc[x], ah0, 0u, ah2, 0u, b0, b1);
        mma_bf16_16816(acc[x], al0, 0u, al2, 0u, b0, b1);
      }
      if (PLANES == 2) {
        const uint32_t w0l[4] = {wl[0].x, wl[0].y, wl[0].z, wl[0].w}, w1l[4] = {wl[1].x, wl[1].y, wl[1].z, wl[1].w};
        const uint32_t w2l[4] = {wl[2].x, wl[2].y, wl[2].z, wl[2].w}, w3l[4] = {wl[3].x, wl[3].y, wl[3].z, wl[3].w};
#pragma unroll
        for (int x = 0; x < 8; ++x) {
          const uint32_t sel = (x & 1) ? 0x7632u : 0x5410u;
          const uint32_t b0 = prmt(w0l[x >> 1], w1l[x >> 1], sel), b1 = prmt(w2l[x >> 1], w3l[x >> 1], sel);
          mma_bf16_16816(acc[x], ah0, 0u, ah2, 0u, b0, b1);
        }
      }
    }
    // C rows g = head, columns 2t and 2t+1 of n-tile x  ->  channels 64 ch + 16 t + x and 64 ch + 16 t + 8 + x
    float* dst = zbp + ((size_t)kq * H + g) * C_Z + 64 * ch + 16 * t;
#pragma unroll
    for (int x = 0; x < 8; ++x) { dst[x] = acc[x][0]; dst[8 + x] = acc[x][1]; }
  }
  __syncthreads();
  {  // zbar = sum of the four edge-quarter partials -> global [B*N, H*128]; o_pair = Wd·zbar + bd follows as one block-diagonal GEMM
    const int h = tid >> 5, c4 = (tid & 31) * 4;
    const float* z0 = zbp + h * C_Z + c4;
    const float4 p0 = *reinterpret_cast<const float4*>(z0), p1 = *reinterpret_cast<const float4*>(z0 + H * C_Z);
    const float4 p2 = *reinterpret_cast<const float4*>(z0 + 2 * H * C_Z), p3 = *reinterpret_cast<const float4*>(z0 + 3 * H * C_Z);
    *reinterpret_cast<float4*>(zbar + rowi * (H * C_Z) + h * C_Z + c4) =
        make_float4((p0.x + p1.x) + (p2.x + p3.x), (p0.y + p1.y) + (p2.y + p3.y), (p0.z + p1.z) + (p2.z + p3.z), (p0.w + p1.w) + (p2.w + p3.w));
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Sequence-transformer self-attention, one kernel (torch.nn.TransformerEncoderLayer inside model/ipa_pytorch.py:584-593,636):
// one CTA per (sample, head), N <= 256.  K and V of the head (dh = 80) are staged once in shared memory as bf16 hi/lo planes;
// each of the 8 warps then runs 16-query-row tiles flash-style with mma.sync m16n8k16: S = q·k^T (3-term split product) over 64-key
// blocks, online softmax in fp32 (keys with mask <= 0.5 excluded, a row without valid keys yields zeros — the eval-mode semantics of
// softmax_rows_kernel), probabilities re-split to bf16 hi/lo in registers as the A operand of P·V (V through ldmatrix.trans).
// Replaces six launches (head-padded split, logits GEMM, softmax, probability split, V transpose, values GEMM) and the [B,4,N,N]
// round trip.  FD_TF_ATTN_GEMM=1 keeps the GEMM path (also the N > 256 path).
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}

constexpr int TFA_STRIDE = 176;     // bytes per K / V row in smem: 80 bf16 + 8 padding (ldmatrix rows land in distinct 16-byte bank groups)
inline size_t tf_attn_smem(int N) { return (size_t)4 * ((N + 63) / 64 * 64) * TFA_STRIDE + ((N + 63) / 64 * 64) * 4; }

__global__ void __launch_bounds__(256) tf_attn_kernel(const float* __restrict__ qkv, const float* __restrict__ keymask, float* __restrict__ y,
                                                      int N, float alpha) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int NK = (N + 63) / 64 * 64;                 // keys padded to whole 64-key blocks
  const uint32_t kh_s = smem_u32(smem_raw), kl_s = kh_s + (uint32_t)NK * TFA_STRIDE, vh_s = kl_s + (uint32_t)NK * TFA_STRIDE,
                 vl_s = vh_s + (uint32_t)NK * TFA_STRIDE;
  float* kvalid = reinterpret_cast<float*>(smem_raw + (size_t)4 * NK * TFA_STRIDE);
  const int bh = blockIdx.x, b = bh / TF_H, hh = bh - b * TF_H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const float* base = qkv + (long long)b * N * (3 * TF_D) + hh * TF_DH;
  // ---- stage K and V of this head: fp32 -> bf16 hi/lo rows ----
  for (int idx = tid; idx < NK * (TF_DH / 4); idx += 256) {
    const int j = idx / (TF_DH / 4), c = (idx - j * (TF_DH / 4)) * 4;
    float4 kf = make_float4(0.f, 0.f, 0.f, 0.f), vf = kf;
    if (j < N) {
      kf = *reinterpret_cast<const float4*>(base + (long long)j * (3 * TF_D) + TF_D + c);
      vf = *reinterpret_cast<const float4*>(base + (long long)j * (3 * TF_D) + 2 * TF_D + c);
    }
    uint32_t h0, l0, h1, l1;
    split2_bf16(kf.x, kf.y, h0, l0); split2_bf16(kf.z, kf.w, h1, l1);
    const uint32_t off = (uint32_t)(j * TFA_STRIDE + c * 2);
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(kh_s + off), "r"(h0), "r"(h1) : "memory");
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(kl_s + off), "r"(l0), "r"(l1) : "memory");
    split2_bf16(vf.x, vf.y, h0, l0); split2_bf16(vf.z, vf.w, h1, l1);
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(vh_s + off), "r"(h0), "r"(h1) : "memory");
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(vl_s + off), "r"(l0), "r"(l1) : "memory");
  }
  for (int j = tid; j < NK; j += 256) kvalid[j] = (j < N && keymask[(long long)b * N + j] > 0.5f) ? 1.f : 0.f;
  __syncthreads();
  const int mi = lane >> 3, lr = lane & 7;
  for (int mt = warp; mt * 16 < N; mt += 8) {       // 16-query-row tiles of this warp
    const int i0 = mt * 16 + g, i1 = i0 + 8;
    // q fragments (A operand): rows i0, i1; k-step ks covers channels 16ks + {2t, 2t+1} and {2t+8, 2t+9}
    uint32_t qh[5][4], ql[5][4];
#pragma unroll
    for (int ks = 0; ks < 5; ++ks) {
      const float2 z2 = make_float2(0.f, 0.f);
      const float2 a0 = i0 < N ? *reinterpret_cast<const float2*>(base + (long long)i0 * (3 * TF_D) + 16 * ks + 2 * t) : z2;
      const float2 a1 = i1 < N ? *reinterpret_cast<const float2*>(base + (long long)i1 * (3 * TF_D) + 16 * ks + 2 * t) : z2;
      const float2 a2 = i0 < N ? *reinterpret_cast<const float2*>(base + (long long)i0 * (3 * TF_D) + 16 * ks + 2 * t + 8) : z2;
      const float2 a3 = i1 < N ? *reinterpret_cast<const float2*>(base + (long long)i1 * (3 * TF_D) + 16 * ks + 2 * t + 8) : z2;
      split2_bf16(a0.x, a0.y, qh[ks][0], ql[ks][0]); split2_bf16(a1.x, a1.y, qh[ks][1], ql[ks][1]);
      split2_bf16(a2.x, a2.y, qh[ks][2], ql[ks][2]); split2_bf16(a3.x, a3.y, qh[ks][3], ql[ks][3]);
    }
    float o[10][4];
#pragma unroll
    for (int c = 0; c < 10; ++c) { o[c][0] = 0.f; o[c][1] = 0.f; o[c][2] = 0.f; o[c][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;      // running max / partial row sums of rows i0, i1
    for (int kb = 0; kb < NK; kb += 64) {
      float sc[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        sc[nt][0] = 0.f; sc[nt][1] = 0.f; sc[nt][2] = 0.f; sc[nt][3] = 0.f;
        const uint32_t rowoff = (uint32_t)((kb + nt * 8 + lr) * TFA_STRIDE);
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {            // k-steps 2kp, 2kp+1: 16-byte chunks 4kp .. 4kp+3
          uint32_t bh4[4], bl4[4];
          ldsm_x4(kh_s + rowoff + (uint32_t)((4 * kp + mi) * 16), bh4);
          ldsm_x4(kl_s + rowoff + (uint32_t)((4 * kp + mi) * 16), bl4);
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int ks = 2 * kp + e;
            mma_bf16_16816(sc[nt], qh[ks][0], qh[ks][1], qh[ks][2], qh[ks][3], bh4[2 * e], bh4[2 * e + 1]);
            mma_bf16_16816(sc[nt], qh[ks][0], qh[ks][1], qh[ks][2], qh[ks][3], bl4[2 * e], bl4[2 * e + 1]);
            mma_bf16_16816(sc[nt], ql[ks][0], ql[ks][1], ql[ks][2], ql[ks][3], bh4[2 * e], bh4[2 * e + 1]);
          }
        }
        uint32_t b0h, b1h, b0l, b1l;                 // k-step 4: chunks 8, 9
        ldsm_x2(kh_s + rowoff + (uint32_t)((8 + (mi & 1)) * 16), b0h, b1h);
        ldsm_x2(kl_s + rowoff + (uint32_t)((8 + (mi & 1)) * 16), b0l, b1l);
        mma_bf16_16816(sc[nt], qh[4][0], qh[4][1], qh[4][2], qh[4][3], b0h, b1h);
        mma_bf16_16816(sc[nt], qh[4][0], qh[4][1], qh[4][2], qh[4][3], b0l, b1l);
        mma_bf16_16816(sc[nt], ql[4][0], ql[4][1], ql[4][2], ql[4][3], b0h, b1h);
      }
      // scale, mask, block maxima of the two rows
      float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float2 kv = *reinterpret_cast<const float2*>(kvalid + kb + nt * 8 + 2 * t);
        sc[nt][0] = kv.x > 0.5f ? sc[nt][0] * alpha : -INFINITY; sc[nt][1] = kv.y > 0.5f ? sc[nt][1] * alpha : -INFINITY;
        sc[nt][2] = kv.x > 0.5f ? sc[nt][2] * alpha : -INFINITY; sc[nt][3] = kv.y > 0.5f ? sc[nt][3] * alpha : -INFINITY;
        bm0 = fmaxf(bm0, fmaxf(sc[nt][0], sc[nt][1])); bm1 = fmaxf(bm1, fmaxf(sc[nt][2], sc[nt][3]));
      }
      bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1)); bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
      bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1)); bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
      const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);
      const float c0 = mn0 == -INFINITY ? 1.f : expf(m0 - mn0), c1 = mn1 == -INFINITY ? 1.f : expf(m1 - mn1);   // exp(-inf) = 0 for a fresh row
      m0 = mn0; m1 = mn1;
      l0 *= c0; l1 *= c1;
#pragma unroll
      for (int c = 0; c < 10; ++c) { o[c][0] *= c0; o[c][1] *= c0; o[c][2] *= c1; o[c][3] *= c1; }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        sc[nt][0] = mn0 == -INFINITY ? 0.f : expf(sc[nt][0] - mn0); sc[nt][1] = mn0 == -INFINITY ? 0.f : expf(sc[nt][1] - mn0);
        sc[nt][2] = mn1 == -INFINITY ? 0.f : expf(sc[nt][2] - mn1); sc[nt][3] = mn1 == -INFINITY ? 0.f : expf(sc[nt][3] - mn1);
        l0 += sc[nt][0] + sc[nt][1]; l1 += sc[nt][2] + sc[nt][3];
      }
      // O += P·V over the block's four 16-key steps
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t ph[4], pl[4];
        split2_bf16(sc[2 * kk][0], sc[2 * kk][1], ph[0], pl[0]); split2_bf16(sc[2 * kk][2], sc[2 * kk][3], ph[1], pl[1]);
        split2_bf16(sc[2 * kk + 1][0], sc[2 * kk + 1][1], ph[2], pl[2]); split2_bf16(sc[2 * kk + 1][2], sc[2 * kk + 1][3], ph[3], pl[3]);
        const uint32_t rowoff = (uint32_t)((kb + kk * 16 + (mi & 1) * 8 + lr) * TFA_STRIDE);
#pragma unroll
        for (int cp = 0; cp < 5; ++cp) {            // channel n-tiles 2cp, 2cp+1
          uint32_t vh4[4], vl4[4];
          ldsm_x4_t(vh_s + rowoff + (uint32_t)((2 * cp + (mi >> 1)) * 16), vh4);
          ldsm_x4_t(vl_s + rowoff + (uint32_t)((2 * cp + (mi >> 1)) * 16), vl4);
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            mma_bf16_16816(o[2 * cp + e], ph[0], ph[1], ph[2], ph[3], vh4[2 * e], vh4[2 * e + 1]);
            mma_bf16_16816(o[2 * cp + e], ph[0], ph[1], ph[2], ph[3], vl4[2 * e], vl4[2 * e + 1]);
            mma_bf16_16816(o[2 * cp + e], pl[0], pl[1], pl[2], pl[3], vh4[2 * e], vh4[2 * e + 1]);
          }
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f, inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
    float* y0 = y + ((long long)b * N + i0) * TF_D + hh * TF_DH, *y1 = y + ((long long)b * N + i1) * TF_D + hh * TF_DH;
#pragma unroll
    for (int c = 0; c < 10; ++c) {
      if (i0 < N) *reinterpret_cast<float2*>(y0 + c * 8 + 2 * t) = make_float2(o[c][0] * inv0, o[c][1] * inv0);
      if (i1 < N) *reinterpret_cast<float2*>(y1 + c * 8 + 2 * t) = make_float2(o[c][2] * inv1, o[c][3] * inv1);
    }
  }
}

static int g_tc_tf_fused = 1;   // FD_TF_ATTN_GEMM=1 selects the six-launch GEMM path (cross-check; also used when N > 256)
inline bool tc_tf_attn_fused_ok(int N) { return g_tc_tf_fused && N <= 256; }
inline int tc_tf_attention(const float* qkv, const float* keymask, float* y, int B, int N, cudaStream_t st, long long* launches) {
  tf_attn_kernel<<<B * TF_H, 256, tf_attn_smem(N), st>>>(qkv, keymask, y, N, (float)(1.0 / sqrt((double)TF_DH)));
  if (launches) ++*launches;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

static int g_tc_edge3 = 1;   // FD_IPA_EDGE2=1 selects the two-kernel path (pair-bias GEMM + attention kernel) kept as the cross-check
static int g_tc_cluster = 0;   // FD_TC_CLUSTER=1: fused EdgeTransition as 2-CTA clusters sharing each weight block by TMA multicast
static int g_tc_fused = 1;   // FD_TC_UNFUSED=1 selects the three-launch path (kept as the fused kernel's cross-check)
inline int tc_init(int sm_count) {
  g_tc_sms = sm_count;
  g_tc_fused = getenv("FD_TC_UNFUSED") ? 0 : 1;
  g_tc_edge3 = getenv("FD_IPA_EDGE2") ? 0 : 1;
  g_tc_tf_fused = getenv("FD_TF_ATTN_GEMM") ? 0 : 1;
  if (cudaFuncSetAttribute(tf_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tf_attn_smem(256)) != cudaSuccess) return -2;
  if (cudaFuncSetAttribute(ipa_edge3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ipa_edge3_smem(256)) != cudaSuccess) return -2;
  if (cudaFuncSetAttribute(ipa_edge3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ipa_edge3_smem(256)) != cudaSuccess) return -2;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return -2;
  g_encode = (PFN_encodeTiled)fn;
  if (cudaFuncSetAttribute(tc_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES) != cudaSuccess) return -2;
  if (cudaFuncSetAttribute(tc_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES) != cudaSuccess) return -2;
  if (const char* e = getenv("FD_TC_SA")) { const int v = atoi(e); g_tc_sa = (v >= 2 && v <= 4) ? v : 0; }
  if (const char* e = getenv("FD_TC_EG")) g_tc_eg = atoi(e) == 2 ? 2 : (atoi(e) == 1 ? 1 : 0);
  if (cudaFuncSetAttribute(tc_edge_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FU_SMEM_BYTES) != cudaSuccess) return -2;
  if (cudaFuncSetAttribute(tc_edge_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FU_SMEM_BYTES) != cudaSuccess) return -2;
  g_tc_cluster = getenv("FD_TC_CLUSTER") ? atoi(getenv("FD_TC_CLUSTER")) : g_tc_cluster;
  if (cudaFuncSetAttribute(tc_embed_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EF_SMEM_BYTES) != cudaSuccess) return -2;
  if (cudaFuncSetAttribute(ipa_edge2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return -2;
  if (cudaFuncSetAttribute(ipa_edge2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return -2;
  return 0;
}

// 2-D bf16 row-major [rows, cols] tensor map with a {64 cols, 128 rows} box and 128B swizzle
inline int tc_make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, int box_rows = TC_BM) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

struct TcMat {   // a bf16 hi/lo weight image [rows, cols] + maps
  __nv_bfloat16* hi = nullptr; __nv_bfloat16* lo = nullptr;
  CUtensorMap mh, ml;         // box {64 k, 128 rows}
  CUtensorMap mh64, ml64;     // box {64 k, 64 rows} (fused kernel's 64-unit hidden chunks)
  int rows = 0, cols = 0;
};

struct TcWeights {
  bool ready = false;
  char* arena = nullptr;
  std::map<const float*, TcMat> lin;   // node-path linears, keyed by the fp32 device weight pointer (rows padded to 128)
  TcMat ee2, ee4;                 // edge embedder layers 2 and 4: [128][128]
  TcMat w1z[3], w2[3], wf[3];     // EdgeTransition: [384][128], [384][384], [128][512] = [Wf | Wf[:, :128]]
  TcMat wb[4];                    // IPA linear_b padded to [128][128] (rows 0..7 real): pair bias z·Wb^T on the tensor cores (UMMA N = 16)
};

struct TcWorkspace {
  char* base = nullptr;
  long long E = 0;
  long long R = 0;                                                  // node rows (B*N)
  __nv_bfloat16 *a_hi = nullptr, *a_lo = nullptr;                    // node-linear A planes scratch [R, <= 2688]
  float* pbias = nullptr;                                            // IPA pair bias [E, 8] fp32
  std::map<int, std::pair<CUtensorMap, CUtensorMap>> a_maps;        // per K: maps over the scratch viewed as [R, K]
  __nv_bfloat16 *z_hi, *z_lo, *h1_hi, *h1_lo, *h2_hi, *h2_lo;
  CUtensorMap m_z_h, m_z_l, m_h1_h, m_h1_l, m_h2_h, m_h2_l;      // K = 128 / 384 / 384
  CUtensorMap m_e0_h, m_e0_l, m_e1_h, m_e1_l;                     // embedder staging viewed as [E,128] inside h1 / h2
  // attention GEMMs (always split precision): q|k|v projection planes [R, 6144], probabilities [B*H*N, Kp], values^T [B*H*256, Kp]
  int Kp = 0;
  __nv_bfloat16 *pj_hi = nullptr, *pj_lo = nullptr, *at_hi = nullptr, *at_lo = nullptr, *vt_hi = nullptr, *vt_lo = nullptr;
  CUtensorMap m_pj_h, m_pj_l, m_at_h, m_at_l, m_vt_h, m_vt_l;
  __nv_bfloat16 *vp_hi = nullptr, *vp_lo = nullptr;                 // IPA value points^T [B*H*64, Kp] (36 real rows per head)
  CUtensorMap m_vp_h, m_vp_l;
  CUtensorMap m_tq_h, m_tq_l;                                       // sequence-transformer q|k head-padded planes [R, 1024] (inside pj)
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

struct TcLinSpec { const float* key; const float* host; int rows, cols; };   // fp32 device pointer (key), host image [rows, cols]

inline int tc_pack_weights(TcWeights& tw, const std::map<std::string, const float*>& M, const std::vector<TcLinSpec>& lins, cudaStream_t st) {
  tc_free_weights(tw);
  struct Item { TcMat* m; std::vector<float> w; int rows, cols; };
  std::vector<Item> items;
  for (const auto& L : lins) {      // node-path linears: pad the output dimension to a multiple of 128 with zero rows
    const int prow = (L.rows + 127) / 128 * 128;
    std::vector<float> w((size_t)prow * L.cols, 0.f);
    memcpy(w.data(), L.host, (size_t)L.rows * L.cols * sizeof(float));
    items.push_back({&tw.lin[L.key], std::move(w), prow, L.cols});
  }
  auto add = [&](TcMat* m, const float* src, int rows, int cols) { items.push_back({m, std::vector<float>(src, src + (size_t)rows * cols), rows, cols}); };
  add(&tw.ee2, M.at("embedding_layer.edge_embedder.2.weight"), 128, 128);
  add(&tw.ee4, M.at("embedding_layer.edge_embedder.4.weight"), 128, 128);
  for (int b = 0; b < 4; ++b) {
    const float* wbp = M.at("score_model.trunk.ipa_" + std::to_string(b) + ".linear_b.weight");   // [8][128]
    std::vector<float> w((size_t)128 * C_Z, 0.f);
    memcpy(w.data(), wbp, (size_t)H * C_Z * sizeof(float));
    items.push_back({&tw.wb[b], std::move(w), 128, C_Z});
  }
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
    if (tc_make_map(&it.m->mh64, it.m->hi, it.rows, it.cols, 64)) return -2;
    if (tc_make_map(&it.m->ml64, it.m->lo, it.rows, it.cols, 64)) return -2;
  }
  tw.ready = true;
  return 0;
}

constexpr int TC_AMAX_K = IPA_FEAT;   // widest node-linear input (linear_out: 2688)
inline size_t tc_hidden_width() { return getenv("FD_TC_UNFUSED") ? (size_t)ET_HID : (size_t)C_Z; }   // h1/h2 staging width
inline size_t tc_workspace_bytes(int B, int N) {
  const size_t E = (size_t)B * N * N, R = (size_t)B * N;
  auto al = [](size_t x) { return (x + 1023) & ~(size_t)1023; };
  const size_t Kp = ((size_t)N + 63) / 64 * 64;
  return 2 * al(E * C_Z * 2) + 4 * al(E * tc_hidden_width() * 2) + 2 * al(R * TC_AMAX_K * 2) + al(E * H * 4) + 1024 +
         2 * al(R * TC_QKV * 2) + 2 * al(R * H * Kp * 2) + 2 * al((size_t)B * H * C_HID * Kp * 2) + 2 * al((size_t)B * H * 64 * Kp * 2);
}
inline int tc_bind_workspace(TcWorkspace& w, char* p, int B, int N) {
  const size_t E = (size_t)B * N * N;
  auto al = [](size_t x) { return (x + 1023) & ~(size_t)1023; };
  p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023);
  w.base = p; w.E = (long long)E;
  w.z_hi = (__nv_bfloat16*)p; p += al(E * C_Z * 2);
  w.z_lo = (__nv_bfloat16*)p; p += al(E * C_Z * 2);
  const size_t HW = tc_hidden_width();
  w.h1_hi = (__nv_bfloat16*)p; p += al(E * HW * 2);
  w.h1_lo = (__nv_bfloat16*)p; p += al(E * HW * 2);
  w.h2_hi = (__nv_bfloat16*)p; p += al(E * HW * 2);
  w.h2_lo = (__nv_bfloat16*)p; p += al(E * HW * 2);
  w.R = (long long)B * N;
  w.a_hi = (__nv_bfloat16*)p; p += al((size_t)w.R * TC_AMAX_K * 2);
  w.a_lo = (__nv_bfloat16*)p; p += al((size_t)w.R * TC_AMAX_K * 2);
  w.pbias = (float*)p; p += al(E * H * 4);
  w.Kp = (N + 63) / 64 * 64;
  const size_t R = (size_t)w.R, Kp = (size_t)w.Kp;
  w.pj_hi = (__nv_bfloat16*)p; p += al(R * TC_QKV * 2);
  w.pj_lo = (__nv_bfloat16*)p; p += al(R * TC_QKV * 2);
  w.at_hi = (__nv_bfloat16*)p; p += al(R * H * Kp * 2);
  w.at_lo = (__nv_bfloat16*)p; p += al(R * H * Kp * 2);
  w.vt_hi = (__nv_bfloat16*)p; p += al((size_t)B * H * C_HID * Kp * 2);
  w.vt_lo = (__nv_bfloat16*)p; p += al((size_t)B * H * C_HID * Kp * 2);
  w.vp_hi = (__nv_bfloat16*)p; p += al((size_t)B * H * 64 * Kp * 2);
  w.vp_lo = (__nv_bfloat16*)p; p += al((size_t)B * H * 64 * Kp * 2);
  w.a_maps.clear();
  int rc = 0;
  rc |= tc_make_map(&w.m_vp_h, w.vp_hi, (uint64_t)B * H * 64, Kp); rc |= tc_make_map(&w.m_vp_l, w.vp_lo, (uint64_t)B * H * 64, Kp);
  rc |= tc_make_map(&w.m_tq_h, w.pj_hi, R, 2 * TF_H * 128); rc |= tc_make_map(&w.m_tq_l, w.pj_lo, R, 2 * TF_H * 128);
  rc |= tc_make_map(&w.m_pj_h, w.pj_hi, R, TC_QKV); rc |= tc_make_map(&w.m_pj_l, w.pj_lo, R, TC_QKV);
  rc |= tc_make_map(&w.m_at_h, w.at_hi, R * H, Kp); rc |= tc_make_map(&w.m_at_l, w.at_lo, R * H, Kp);
  rc |= tc_make_map(&w.m_vt_h, w.vt_hi, (uint64_t)B * H * C_HID, Kp); rc |= tc_make_map(&w.m_vt_l, w.vt_lo, (uint64_t)B * H * C_HID, Kp);
  for (int K : {128, 256, 320, 384, H * C_Z, IPA_FEAT}) {
    std::pair<CUtensorMap, CUtensorMap> mp;
    rc |= tc_make_map(&mp.first, w.a_hi, (uint64_t)w.R, (uint64_t)K);
    rc |= tc_make_map(&mp.second, w.a_lo, (uint64_t)w.R, (uint64_t)K);
    w.a_maps[K] = mp;
  }
  rc |= tc_make_map(&w.m_z_h, w.z_hi, E, C_Z); rc |= tc_make_map(&w.m_z_l, w.z_lo, E, C_Z);
  if (HW == (size_t)ET_HID) {   // three-launch cross-check path only
    rc |= tc_make_map(&w.m_h1_h, w.h1_hi, E, ET_HID); rc |= tc_make_map(&w.m_h1_l, w.h1_lo, E, ET_HID);
    rc |= tc_make_map(&w.m_h2_h, w.h2_hi, E, ET_HID); rc |= tc_make_map(&w.m_h2_l, w.h2_lo, E, ET_HID);
  }
  rc |= tc_make_map(&w.m_e0_h, w.h1_hi, E, C_Z); rc |= tc_make_map(&w.m_e0_l, w.h1_lo, E, C_Z);
  rc |= tc_make_map(&w.m_e1_h, w.h2_hi, E, C_Z); rc |= tc_make_map(&w.m_e1_l, w.h2_lo, E, C_Z);
  return rc;
}

inline int tc_launch_maps(const CUtensorMap& a0h, const CUtensorMap& a0l, const CUtensorMap& a1h, const CUtensorMap& a1l, const CUtensorMap& bh,
                          const CUtensorMap& bl, TcGemmParams p, cudaStream_t st, long long* launches) {
  if (p.epi == TC_EPI_RELU) {          // one 128-column chunk per work item (double-buffered accumulators), chunks of a row tile adjacent
    p.m_tiles = (p.M + TC_BM - 1) / TC_BM; p.nch = 1; p.chunk_minor = p.N / TC_NC; p.num_tiles = p.m_tiles * p.chunk_minor; p.n_valid = p.N;
  } else if (p.epi != TC_EPI_F32) {
    p.m_tiles = (p.M + TC_BM - 1) / TC_BM; p.nch = p.N / TC_NC; p.num_tiles = p.m_tiles; p.n_valid = p.N;
  }
  if (p.mma_n == 0) p.mma_n = TC_NC;
  if (g_tc_sa && p.sa > 0) p.sa = g_tc_sa;
  const int grid = p.num_tiles < g_tc_sms ? p.num_tiles : g_tc_sms;
  if (g_tc_eg == 2 || (g_tc_eg == 0 && p.eg == 2)) tc_gemm_kernel<2><<<grid, 320, TC_SMEM_BYTES, st>>>(a0h, a0l, a1h, a1l, bh, bl, p);
  else tc_gemm_kernel<1><<<grid, 192, TC_SMEM_BYTES, st>>>(a0h, a0l, a1h, a1l, bh, bl, p);
  if (launches) ++*launches;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
inline int tc_launch(const CUtensorMap& a0h, const CUtensorMap& a0l, const CUtensorMap& a1h, const CUtensorMap& a1l, const TcMat& Wt,
                     TcGemmParams p, cudaStream_t st, long long* launches) {
  return tc_launch_maps(a0h, a0l, a1h, a1l, Wt.mh, Wt.ml, p, st, launches);
}

// Edge embedder (model/score_network.py:79-86): layer 0 = table lookup kernel -> bf16 planes; layers 2, 4 on tensor cores.
inline int tc_edge_embed(const TcWeights& tw, TcWorkspace& w, int prec, const float* AC, const float* T, const float* D, const float* w0r,
                         const int* seq_idx, const float* sc_ca, const float* res_mask, const float* b2, const float* b4,
                         const float* ln_g, const float* ln_b, int B, int N, cudaStream_t st,
                         long long* launches) {
  const long long E = w.E;
  const int planes = prec == 1 ? 2 : 1;
  if (E > 0x7fffffffLL) return -1;
  const dim3 l0_grid((unsigned)w.R, (unsigned)((N + 7) / 8));
  if (planes == 2) edge_embed_l0_rows_kernel<2><<<l0_grid, 256, 0, st>>>(AC, T, D, w0r, seq_idx, sc_ca, w.h1_hi, w.h1_lo, N);
  else edge_embed_l0_rows_kernel<1><<<l0_grid, 256, 0, st>>>(AC, T, D, w0r, seq_idx, sc_ca, w.h1_hi, w.h1_lo, N);
  if (launches) ++*launches;
  if (g_tc_fused) {   // layers 2..4 in one persistent kernel (h1 stays in tensor memory)
    EmbedFusedParams f{};
    f.E = (int)E; f.planes = planes; f.nres = N; f.num_tiles = (int)((E + TC_BM - 1) / TC_BM);
    f.b2 = b2; f.b4 = b4; f.ln_g = ln_g; f.ln_b = ln_b; f.res_mask = res_mask; f.out_hi = w.z_hi; f.out_lo = w.z_lo;
    const int grid = f.num_tiles < g_tc_sms ? f.num_tiles : g_tc_sms;
    tc_embed_fused_kernel<<<grid, EF_THREADS, EF_SMEM_BYTES, st>>>(w.m_e0_h, w.m_e0_l, tw.ee2.mh, tw.ee2.ml, tw.ee4.mh, tw.ee4.ml, f);
    if (launches) ++*launches;
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
  }
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
                              const float* ln_b, const float* res_mask, int B, int N,
                              cudaStream_t st, long long* launches) {
  const long long E = w.E;
  const int planes = prec == 1 ? 2 : 1;
  if (E > 0x7fffffffLL) return -1;
  if (g_tc_fused) {
    FusedParams f{};
    f.E = (int)E; f.planes = planes; f.nres = N; f.num_tiles = (int)((E + TC_BM - 1) / TC_BM);
    f.pquv = pquv; f.b2 = b2; f.ln_g = ln_g; f.ln_b = ln_b; f.res_mask = res_mask; f.out_hi = w.z_hi; f.out_lo = w.z_lo;
    f.prof = g_tc_prof;
    f.dbg_noq = getenv("FD_FU_NOQ") ? 1 : 0;
    int grid = f.num_tiles < g_tc_sms ? f.num_tiles : g_tc_sms;
    if (g_tc_cluster) {
      grid = (grid + 1) & ~1;                 // whole pairs; a CTA without tiles still mirrors its peer's weight sequence
      if (grid > (g_tc_sms & ~1)) grid = g_tc_sms & ~1;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(FU_THREADS); cfg.dynamicSmemBytes = FU_SMEM_BYTES; cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      if (cudaLaunchKernelEx(&cfg, tc_edge_fused_kernel<true>, w.m_z_h, w.m_z_l, tw.w1z[blk].mh64, tw.w1z[blk].ml64, tw.w2[blk].mh, tw.w2[blk].ml,
                             tw.wf[blk].mh, tw.wf[blk].ml, f) != cudaSuccess) return -2;
    } else {
      tc_edge_fused_kernel<false><<<grid, FU_THREADS, FU_SMEM_BYTES, st>>>(w.m_z_h, w.m_z_l, tw.w1z[blk].mh64, tw.w1z[blk].ml64, tw.w2[blk].mh,
                                                                          tw.w2[blk].ml, tw.wf[blk].mh, tw.wf[blk].ml, f);
    }
    if (launches) ++*launches;
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
  }
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

// y[M, n_out] = act(x[M, K] · W^T + b) (* rowmask) (+ residual): node-path linear on the tensor cores.
// Returns 1 if this linear is not registered / not eligible (caller uses the CUDA-core GEMM), 0 on success, <0 on error.
inline int tc_linear(const TcWeights& tw, TcWorkspace& w, int prec, const float* x, int ldx, const float* wkey, const float* bias, int K,
                     int n_out, float* y, int ldy, long long M, bool relu, const float* residual, int ldr, const float* rowmask,
                     cudaStream_t st, long long* launches) {
  auto it = tw.lin.find(wkey);
  if (it == tw.lin.end() || K % TC_BK != 0 || K > TC_AMAX_K || M > w.R || n_out % 4 != 0 || ldx % 4 != 0) return 1;
  auto mp = w.a_maps.find(K);
  if (mp == w.a_maps.end()) return 1;
  const int planes = prec == 1 ? 2 : 1;
  const long long n4 = M * (K / 4);
  split_planes_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(x, ldx, M, K, w.a_hi, planes == 2 ? w.a_lo : nullptr);
  if (launches) ++*launches;
  const TcMat& Wt = it->second;
  TcGemmParams p{};
  p.M = (int)M; p.N = Wt.rows; p.KB0 = K / TC_BK; p.KB1 = 0; p.planes = planes; p.epi = TC_EPI_F32; p.bias = bias;
  p.m_tiles = (int)((M + TC_BM - 1) / TC_BM);
  const int chunks = Wt.rows / TC_NC;
  p.nch = (chunks % 2 == 0 && p.m_tiles * (chunks / 2) >= g_tc_sms) ? 2 : 1;    // wider work items once the grid is full anyway
  p.num_tiles = p.m_tiles * (chunks / p.nch);
  p.n_valid = n_out; p.out_f32 = y; p.ldo = ldy; p.residual = residual; p.ldr = ldr; p.rowmask = rowmask; p.relu = relu ? 1 : 0;
  return tc_launch(mp->second.first, mp->second.second, mp->second.first, mp->second.second, Wt, p, st, launches);
}

inline void tc_export_z(TcWorkspace& w, float* z_f32, int prec, cudaStream_t st) {
  const long long n = w.E * C_Z;
  planes_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w.z_hi, prec == 1 ? w.z_lo : nullptr, z_f32, n);
}

// IPA edge pass in the tensor-core modes.  Pair bias  pbias[e,h] = z_e·Wb[h]^T + bb[h]  first, as a memory-bound tcgen05 GEMM over
// the z planes (UMMA N = 16, 8 valid columns, fp32 out), then the attention kernel streams z exactly once (Σ_j a·z).
inline bool tc_ipa_edge_fused_ok(int N) { return g_tc_edge3 && N <= 256; }
inline int tc_ipa_edge(const TcWeights& tw, TcWorkspace& w, int blk, float* L, const float* qp, const float* kp, const float* res_mask,
                       const float* Wb, const float* bb, const float* gamma, const float* WdT, const float* bd, float* feats, float* zbar, int B,
                       int N, int Np, int prec, int write_L, cudaStream_t st, long long* launches) {
  const int planes = prec == 1 ? 2 : 1;
  if (tc_ipa_edge_fused_ok(N)) {
    const size_t smem = ipa_edge3_smem(Np);
    if (prec == 1) ipa_edge3_kernel<2><<<dim3(N, B), 256, smem, st>>>(w.z_hi, w.z_lo, L, res_mask, Wb, bb, zbar, w.at_hi, w.at_lo, w.Kp, write_L, N, Np);
    else ipa_edge3_kernel<1><<<dim3(N, B), 256, smem, st>>>(w.z_hi, w.z_lo, L, res_mask, Wb, bb, zbar, w.at_hi, w.at_lo, w.Kp, write_L, N, Np);
    if (launches) ++*launches;
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
  }
  TcGemmParams p{};
  p.M = (int)w.E; p.N = 128; p.KB0 = 2; p.KB1 = 0; p.planes = planes; p.epi = TC_EPI_F32; p.bias = bb; p.mma_n = 16; p.lolo = 1;
  p.m_tiles = (int)((w.E + TC_BM - 1) / TC_BM); p.nch = 1; p.num_tiles = p.m_tiles; p.n_valid = H; p.out_f32 = w.pbias; p.ldo = H;
  if (tc_launch(w.m_z_h, w.m_z_l, w.m_z_h, w.m_z_l, tw.wb[blk], p, st, launches)) return -2;
  const size_t smem = (size_t)(H * Np + 8 * H * C_Z) * sizeof(float);
  ZRef z; z.hi = w.z_hi; z.lo = w.z_lo;
  if (prec == 1) ipa_edge2_kernel<2><<<dim3(N, B), 256, smem, st>>>(z, L, w.pbias, qp, kp, res_mask, gamma, WdT, bd, feats, N, Np);
  else ipa_edge2_kernel<1><<<dim3(N, B), 256, smem, st>>>(z, L, w.pbias, qp, kp, res_mask, gamma, WdT, bd, feats, N, Np);
  if (launches) ++*launches;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// IPA attention logits on the tensor cores (model/ipa_pytorch.py:376-417): per (sample, head) a [N,N] = q·k^T GEMM, both operands from split
// planes, scaled by sqrt(1/(3*c_hidden)), fp32 out into L [B,H,N,Np].  K = 256 (scalar part) or 384 with the point-distance columns.
inline int tc_ipa_logits(TcWorkspace& w, const float* proj, const float* qp, const float* kp, const float* gamma, float* L, int B, int N, int Np,
                         float alpha, cudaStream_t st, long long* launches) {
  const bool ext = tc_ipa_edge_fused_ok(N);     // the one-kernel edge pass takes the point-distance term from this GEMM (K = 384 per head)
  if (ext) {
    ipa_qkx_planes_kernel<<<(unsigned)w.R, 256, 0, st>>>(proj, qp, kp, gamma, 1.f / alpha, w.pj_hi, w.pj_lo);
  } else {
    const long long n4 = (long long)w.R * (TC_QKV / 4);
    split_planes_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(proj, PROJ_ALL, w.R, TC_QKV, w.pj_hi, w.pj_lo);
  }
  if (launches) ++*launches;
  TcGemmParams p{};
  p.M = N; p.N = 128; p.KB0 = (ext ? 384 : C_HID) / TC_BK; p.KB1 = 0; p.planes = 2; p.epi = TC_EPI_F32;
  p.m_tiles = (N + TC_BM - 1) / TC_BM; p.nch = 1;
  const int n_groups = (N + TC_NC - 1) / TC_NC;
  p.bat_inner = H; p.bat_tiles = p.m_tiles * n_groups; p.num_tiles = B * H * p.bat_tiles;
  p.a_row_s0 = N; p.a_row_s1 = 0; p.a_k_s1 = ext ? 384 : C_HID;
  p.b_row_s0 = N; p.b_row_s1 = 0; p.b_k0 = ext ? 3072 : PROJ_Q; p.b_k_s1 = ext ? 384 : 2 * C_HID;
  p.n_valid = N; p.out_f32 = L; p.ldo = Np; p.o_s0 = (long long)H * N * Np; p.o_s1 = (long long)N * Np; p.alpha = alpha;
  return tc_launch_maps(w.m_pj_h, w.m_pj_l, w.m_pj_h, w.m_pj_l, w.m_pj_h, w.m_pj_l, p, st, launches);
}

// IPA o = a·v and o_pt = a·v_pts on the tensor cores (model/ipa_pytorch.py:433-447): per (sample, head) [N, 256] = a [N, N] · v [N, 256] and
// [N, 36] = a · v_pts, K = N padded to a multiple of 64 with zeros in both operands.  Outputs: feats[:, h*256 : (h+1)*256] and optg.
inline int tc_ipa_av(TcWorkspace& w, const float* proj, const float* vp, const float* L, float* feats, float* optg, int B, int N, int Np,
                     cudaStream_t st, long long* launches) {
  const long long M = (long long)w.R * H;
  if (!tc_ipa_edge_fused_ok(N)) {     // the one-kernel edge pass has already written the probability planes
    const long long n4 = M * (w.Kp / 4);
    split_pad_planes_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(L, Np, M, N, w.Kp, w.at_hi, w.at_lo);
    if (launches) ++*launches;
  }
  vt_planes_kernel<<<dim3(w.Kp / 64, C_HID / 32, B * H), dim3(32, 8), 0, st>>>(proj, PROJ_ALL, PROJ_Q + C_HID, 2 * C_HID, H, C_HID, C_HID, N, w.Kp,
                                                                            w.vt_hi, w.vt_lo);
  vt_planes_kernel<<<dim3(w.Kp / 64, 2, B * H), dim3(32, 8), 0, st>>>(vp, H * PV * 3, 0, PV * 3, H, PV * 3, 64, N, w.Kp, w.vp_hi, w.vp_lo);
  if (launches) *launches += 2;
  TcGemmParams p{};
  p.M = N; p.N = C_HID; p.KB0 = w.Kp / TC_BK; p.KB1 = 0; p.planes = 2; p.epi = TC_EPI_F32;
  p.m_tiles = (N + TC_BM - 1) / TC_BM; p.nch = 2;
  p.bat_inner = H; p.bat_tiles = p.m_tiles; p.num_tiles = B * H * p.bat_tiles;
  p.a_row_s0 = H * N; p.a_row_s1 = N; p.a_k_s1 = 0;
  p.b_row_s0 = H * C_HID; p.b_row_s1 = C_HID; p.b_k0 = 0; p.b_k_s1 = 0;
  p.n_valid = C_HID; p.out_f32 = feats; p.ldo = IPA_FEAT; p.o_s0 = (long long)N * IPA_FEAT; p.o_s1 = C_HID;
  if (tc_launch_maps(w.m_at_h, w.m_at_l, w.m_at_h, w.m_at_l, w.m_vt_h, w.m_vt_l, p, st, launches)) return -2;
  TcGemmParams q = p;   // points: 36 valid output columns per head (the 128-row weight box also covers the next head's rows; ignored)
  q.N = 128; q.nch = 1; q.b_row_s0 = H * 64; q.b_row_s1 = 64;
  q.n_valid = PV * 3; q.out_f32 = optg; q.ldo = H * PV * 3; q.o_s0 = (long long)N * H * PV * 3; q.o_s1 = PV * 3;
  return tc_launch_maps(w.m_at_h, w.m_at_l, w.m_at_h, w.m_at_l, w.m_vp_h, w.m_vp_l, q, st, launches);
}

// Sequence-transformer self-attention (torch.nn.TransformerEncoderLayer inside model/ipa_pytorch.py:583-599) on the tensor cores.
//   logits: S[b,h] = q_h·k_h^T / sqrt(80), K = 80 zero-padded to 128 per head;   values: y[:, h*80:(h+1)*80] = P[b,h]·v_h.
inline int tc_tf_logits(TcWorkspace& w, const float* qkv, float* S, int B, int N, int Np, cudaStream_t st, long long* launches) {
  const int W = 2 * TF_H * 128;
  const long long n4 = (long long)w.R * (W / 4);
  split_heads_pad_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(qkv, 3 * TF_D, w.R, 2 * TF_H, TF_DH, 128, w.pj_hi, w.pj_lo);
  if (launches) ++*launches;
  TcGemmParams p{};
  p.M = N; p.N = 128; p.KB0 = 2; p.KB1 = 0; p.planes = 2; p.epi = TC_EPI_F32;
  p.m_tiles = (N + TC_BM - 1) / TC_BM; p.nch = 1;
  const int n_groups = (N + TC_NC - 1) / TC_NC;
  p.bat_inner = TF_H; p.bat_tiles = p.m_tiles * n_groups; p.num_tiles = B * TF_H * p.bat_tiles;
  p.a_row_s0 = N; p.a_row_s1 = 0; p.a_k_s1 = 128;
  p.b_row_s0 = N; p.b_row_s1 = 0; p.b_k0 = TF_H * 128; p.b_k_s1 = 128;
  p.n_valid = N; p.out_f32 = S; p.ldo = Np; p.o_s0 = (long long)TF_H * N * Np; p.o_s1 = (long long)N * Np;
  p.alpha = (float)(1.0 / sqrt((double)TF_DH));
  return tc_launch_maps(w.m_tq_h, w.m_tq_l, w.m_tq_h, w.m_tq_l, w.m_tq_h, w.m_tq_l, p, st, launches);
}
inline int tc_tf_values(TcWorkspace& w, const float* qkv, const float* S, float* y, int B, int N, int Np, cudaStream_t st, long long* launches) {
  const long long M = (long long)w.R * TF_H;
  const long long n4 = M * (w.Kp / 4);
  split_pad_planes_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(S, Np, M, N, w.Kp, w.at_hi, w.at_lo);
  vt_planes_kernel<<<dim3(w.Kp / 64, 4, B * TF_H), dim3(32, 8), 0, st>>>(qkv, 3 * TF_D, 2 * TF_D, TF_DH, TF_H, TF_DH, 128, N, w.Kp, w.vt_hi, w.vt_lo);
  if (launches) *launches += 2;
  TcGemmParams p{};
  p.M = N; p.N = 128; p.KB0 = w.Kp / TC_BK; p.KB1 = 0; p.planes = 2; p.epi = TC_EPI_F32;
  p.m_tiles = (N + TC_BM - 1) / TC_BM; p.nch = 1;
  p.bat_inner = TF_H; p.bat_tiles = p.m_tiles; p.num_tiles = B * TF_H * p.bat_tiles;
  p.a_row_s0 = TF_H * N; p.a_row_s1 = N; p.a_k_s1 = 0;
  p.b_row_s0 = TF_H * 128; p.b_row_s1 = 128; p.b_k0 = 0; p.b_k_s1 = 0;
  p.n_valid = TF_DH; p.out_f32 = y; p.ldo = TF_D; p.o_s0 = (long long)N * TF_D; p.o_s1 = TF_DH;
  return tc_launch_maps(w.m_at_h, w.m_at_l, w.m_at_h, w.m_at_l, w.m_vt_h, w.m_vt_l, p, st, launches);
}

}  // namespace fd
