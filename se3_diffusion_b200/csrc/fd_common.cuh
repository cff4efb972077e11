// Shared constants + device helpers for the FrameDiff B200 kernels.
// Hyper-parameters are compile-time constants (config/base.yaml:25-67 of the reference; both shipped checkpoints
// carry the same values in their pickled conf).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace fd {

constexpr int C_S = 256;        // node channels
constexpr int C_Z = 128;        // edge channels
constexpr int C_HID = 256;      // IPA scalar head dim
constexpr int C_SKIP = 64;
constexpr int H = 8;            // IPA heads
constexpr int PQ = 8;           // query/key points per head
constexpr int PV = 12;          // value points per head
constexpr int NBLK = 4;
constexpr int TF_D = C_S + C_SKIP;   // 320
constexpr int TF_H = 4;
constexpr int TF_DH = TF_D / TF_H;   // 80
constexpr int TF_LAYERS = 2;
constexpr int IDX_EMB = 32;
constexpr int NBINS = 22;
constexpr int NODE_IN = 65, NODE_IN_PAD = 68;
constexpr int EDGE_IN = 120;
constexpr int IPA_FEAT = H * (C_Z / 4 + C_HID + PV * 4);   // 2688
constexpr int PROJ_Q = H * C_HID;                 // 2048
constexpr int PROJ_KV = 2 * H * C_HID;            // 4096
constexpr int PROJ_QP = H * PQ * 3;               // 192
constexpr int PROJ_KVP = H * (PQ + PV) * 3;       // 480
constexpr int PROJ_ALL = PROJ_Q + PROJ_KV + PROJ_QP + PROJ_KVP;   // 6816
constexpr int ET_HID = 3 * C_Z;                   // 384
constexpr int ET_NODE = 2 * ET_HID + 2 * C_Z;     // P(384) Q(384) U(128) V(128) = 1024
constexpr int REL_DMAX = 2056;                    // |seq_idx_i - seq_idx_j| table half-width (max_len of the embedding)
constexpr float COORD_SCALE = 0.1f;
constexpr double R3_MIN_B = 0.1, R3_MAX_B = 20.0;
constexpr double SO3_MIN_SIGMA = 0.1, SO3_MAX_SIGMA = 1.5;
constexpr int SO3_NSIGMA = 1000, SO3_NOMEGA = 1000, IGSO3_L = 1000;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// quaternion (w,x,y,z) -> rotation matrix, openfold/utils/rigid_utils.py:185 (no normalisation, like the reference)
__device__ __forceinline__ void quat_to_rot(const float q[4], float R[9]) {
  const float a = q[0], b = q[1], c = q[2], d = q[3];
  R[0] = a * a + b * b - c * c - d * d; R[1] = 2.f * b * c - 2.f * a * d;       R[2] = 2.f * b * d + 2.f * a * c;
  R[3] = 2.f * b * c + 2.f * a * d;       R[4] = a * a - b * b + c * c - d * d; R[5] = 2.f * c * d - 2.f * a * b;
  R[6] = 2.f * b * d - 2.f * a * c;       R[7] = 2.f * c * d + 2.f * a * b;       R[8] = a * a - b * b - c * c + d * d;
}
__device__ __forceinline__ void quat_to_rot_d(const double q[4], double R[9]) {
  const double a = q[0], b = q[1], c = q[2], d = q[3];
  R[0] = a * a + b * b - c * c - d * d; R[1] = 2. * b * c - 2. * a * d;       R[2] = 2. * b * d + 2. * a * c;
  R[3] = 2. * b * c + 2. * a * d;       R[4] = a * a - b * b + c * c - d * d; R[5] = 2. * c * d - 2. * a * b;
  R[6] = 2. * b * d - 2. * a * c;       R[7] = 2. * c * d + 2. * a * b;       R[8] = a * a - b * b - c * c + d * d;
}
// Hamilton product r = p ⊗ q
template <typename T>
__device__ __forceinline__ void quat_mul(const T p[4], const T q[4], T r[4]) {
  r[0] = p[0] * q[0] - p[1] * q[1] - p[2] * q[2] - p[3] * q[3];
  r[1] = p[0] * q[1] + p[1] * q[0] + p[2] * q[3] - p[3] * q[2];
  r[2] = p[0] * q[2] - p[1] * q[3] + p[2] * q[0] + p[3] * q[1];
  r[3] = p[0] * q[3] + p[1] * q[2] - p[2] * q[1] + p[3] * q[0];
}

// Philox4x32-10 counter-based RNG (Salmon et al. 2011) — results independent of grid shape / GPU count.
struct Philox {
  uint32_t key[2];
  __device__ __forceinline__ Philox(uint64_t seed) { key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32); }
  __device__ __forceinline__ void operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) const {
    uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
};
// two uint32 -> uniform double in (0,1) with 53 bits
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
  const uint64_t x = (((uint64_t)a << 32) | b) >> 11;
  return ((double)x + 0.5) * (1.0 / 9007199254740992.0);
}
// Box–Muller: 4 uint32 -> 2 standard normals (fp64)
__device__ __forceinline__ void normal2(const uint32_t r[4], double& z0, double& z1) {
  const double u1 = u53(r[0], r[1]), u2 = u53(r[2], r[3]);
  const double rad = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  z0 = rad * c; z1 = rad * s;
}

}  // namespace fd
