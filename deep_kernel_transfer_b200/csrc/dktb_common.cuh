// Shared definitions for the dktb200 kernels (sm_100a).  The same sources also build with g++
// under -DDKTB_EMU (tests/emu/cuda_emu.h) so the plain-CUDA kernels can be logic-checked on a
// GPU-less box; that build is test infrastructure and is never loaded by the product.
#pragma once

#ifdef DKTB_EMU
#include "cuda_emu.h"
#define DKTB_LAUNCH(kern, grid, block, smem, stream, ...) \
  emu::launch((grid), (block), (smem), [&]() { kern(__VA_ARGS__); })
#define DKTB_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::dyn_smem)
#else
#include <cuda_runtime.h>
#include <stdint.h>
#define DKTB_LAUNCH(kern, grid, block, smem, stream, ...) \
  kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define DKTB_DYN_SMEM(type, name)                                  \
  extern __shared__ __align__(1024) unsigned char dktb_dyn_smem_[]; \
  type* name = reinterpret_cast<type*>(dktb_dyn_smem_)
#endif

#include <math.h>

#define DKTB_EXPORT extern "C" __attribute__((visibility("default")))

#define DKTB_C 64          // channel width of every ConvBlock (backbone.py:253-255)
#define DKTB_BAD_ARG (-1)

// Return the CUDA status of the launch that just happened (0 = ok).
static inline int dktb_launch_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

#define DKTB_CHECK_ARG(cond) \
  do {                       \
    if (!(cond)) return DKTB_BAD_ARG; \
  } while (0)

__device__ __forceinline__ float4 dktb_ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void dktb_st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float dktb_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double dktb_warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float dktb_softplus(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float dktb_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

// Warp-level tensor-core MMA D(16x8) += A(16x8, row) * B(8x8, col), tf32 inputs / fp32 accumulate
// (mma.sync.aligned.m16n8k8).  Fragment ownership with g = lane / 4, t = lane % 4:
//   a0 = A[g][t]   a1 = A[g+8][t]   a2 = A[g][t+4]   a3 = A[g+8][t+4]
//   b0 = B[t][g]   b1 = B[t+4][g]
//   d0 = D[g][2t]  d1 = D[g][2t+1]  d2 = D[g+8][2t]  d3 = D[g+8][2t+1]
// The tensor core reads the upper 19 bits of each operand.  The g++ emulation (tests only) rebuilds the products with
// warp shuffles so that the fragment indexing of the callers is exercised without a GPU.
__device__ __forceinline__ void dktb_mma_m16n8k8_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
#ifdef DKTB_EMU
  const int lane = threadIdx.x % 32, g = lane >> 2, t = lane & 3;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k = 0; k < 8; ++k) {
    const int kr = k & 3, khi = k >> 2;
    // A[g][k], A[g+8][k]: owner lane g*4 + kr, registers (khi ? a2 : a0), (khi ? a3 : a1)
    const unsigned alo = __shfl_sync(0xffffffffu, khi ? a[2] : a[0], g * 4 + kr);
    const unsigned ahi = __shfl_sync(0xffffffffu, khi ? a[3] : a[1], g * 4 + kr);
    // B[k][2t], B[k][2t+1]: owner lanes (2t)*4 + kr, (2t+1)*4 + kr, register (khi ? b1 : b0)
    const unsigned b0v = __shfl_sync(0xffffffffu, khi ? b[1] : b[0], (2 * t) * 4 + kr);
    const unsigned b1v = __shfl_sync(0xffffffffu, khi ? b[1] : b[0], (2 * t + 1) * 4 + kr);
    const float fa0 = __uint_as_float(alo & 0xFFFFE000u), fa1 = __uint_as_float(ahi & 0xFFFFE000u);
    const float fb0 = __uint_as_float(b0v & 0xFFFFE000u), fb1 = __uint_as_float(b1v & 0xFFFFE000u);
    acc[0] = fmaf(fa0, fb0, acc[0]);
    acc[1] = fmaf(fa0, fb1, acc[1]);
    acc[2] = fmaf(fa1, fb0, acc[2]);
    acc[3] = fmaf(fa1, fb1, acc[3]);
  }
  for (int i = 0; i < 4; ++i) d[i] += acc[i];
#else
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
#endif
}

// BatchNorm + ReLU + MaxPool backward gate of one pool window and one channel: which of the (up to) 4 raw conv outputs
// receives the pooled gradient (first maximum of the post-ReLU values in scan order, as torch's max_pool2d), the gated
// gradient and the normalised value at that position.  Shared by bn_pool.cu and the fused first-layer backward.
__device__ __forceinline__ void bn_bwd_route(const float yv[4], int npos, float m, float is, float sc, float bt,
                                             float gout, int& arg, float& gz, float& xhat_arg) {
  float bestz = 0.f;
  arg = 0;
  float zs0 = fmaxf(fmaf(yv[0] - m, sc, bt), 0.f);
  bestz = zs0;
  for (int k = 1; k < npos; ++k) {
    const float z = fmaxf(fmaf(yv[k] - m, sc, bt), 0.f);
    if (z > bestz) { bestz = z; arg = k; }
  }
  gz = bestz > 0.f ? gout : 0.f;
  xhat_arg = (yv[arg] - m) * is;
}

// Padded activation layout helpers: [img][H+2][W+2][64], zero border.
__host__ __device__ __forceinline__ long dktb_prow(int img, int hp, int wp, int Hp, int Wp) {
  return ((long)img * Hp + hp) * Wp + wp;
}
