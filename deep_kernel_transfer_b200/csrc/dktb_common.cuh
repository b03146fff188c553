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
