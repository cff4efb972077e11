// Microbenchmark: peak rate of legacy mma.sync.m16n8k16 (bf16 -> fp32) on this GPU, registers only (no memory traffic).
// Decides whether the training-path GEMM (fd_mm3.cuh, mma.sync) is close to what the instruction can deliver.  nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  for (int blocks_per_sm = 1; blocks_per_sm <= 4; blocks_per_sm *= 2) {
    const int iters = 20000, grid = 148 * blocks_per_sm;
    k<<<grid, 256>>>(d, 100); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<<<grid, 256>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 16 * 8 * 16 * 8.0 * iters * 8 /*warps*/ * grid;
    printf("mma.sync m16n8k16 bf16: %d CTA/SM x 8 warps: %.1f TFLOP/s (%.3f ms)\n", blocks_per_sm, flops / ms / 1e9, ms);
  }
  return 0;
}
