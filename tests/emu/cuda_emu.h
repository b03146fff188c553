// TEST INFRASTRUCTURE ONLY.  A tiny CPU emulation of the CUDA execution model so that the
// plain-CUDA kernels in deep_kernel_transfer_b200/csrc can be compiled with g++ (-DDKTB_EMU)
// and exercised on the GPU-less authoring box: index math, shared-memory staging, barriers and
// warp shuffles are executed with one OS thread per CUDA thread, blocks run one after another.
// It exists to catch logic bugs before spending B200 minutes; it is never loaded by the product
// (deep_kernel_transfer_b200/_lib.py only opens the nvcc-built library).  tcgen05/TMA kernels
// are outside its reach and are compiled out under DKTB_EMU.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) int2 { int x, y; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
struct alignas(4) uchar4 { unsigned char x, y, z, w; };
static inline uchar4 make_uchar4(unsigned char a, unsigned char b, unsigned char c, unsigned char d) { return uchar4{a, b, c, d}; }
static inline int4 make_int4(int a, int b, int c, int d) { return int4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
static const int cudaSuccess = 0;
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline const char* cudaGetErrorString(int) { return "emu"; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

namespace emu {
struct Ctx {
  dim3 tid, bid;
  int lin = 0;
};
extern thread_local Ctx ctx;
extern dim3 g_block, g_grid;
extern std::barrier<>* g_block_bar;
extern std::vector<std::unique_ptr<std::barrier<>>> g_warp_bar;
extern uint64_t (*g_warp_buf)[32];
extern unsigned char* dyn_smem;
extern std::mutex g_atomic_mu;

template <class F>
void launch(dim3 grid, dim3 block, size_t smem, F&& f);
}  // namespace emu

// threadIdx / blockIdx / blockDim / gridDim as objects with .x .y .z
#define threadIdx (emu::ctx.tid)
#define blockIdx (emu::ctx.bid)
#define blockDim (emu::g_block)
#define gridDim (emu::g_grid)

static inline void __syncthreads() { emu::g_block_bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::g_warp_bar[emu::ctx.lin / 32]->arrive_and_wait(); }
static inline void __threadfence() {}

template <class T>
static inline T emu_shfl(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shfl size");
  int w = emu::ctx.lin / 32, lane = emu::ctx.lin % 32;
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  emu::g_warp_buf[w][lane] = raw;
  emu::g_warp_bar[w]->arrive_and_wait();
  uint64_t got = emu::g_warp_buf[w][src_lane & 31];
  emu::g_warp_bar[w]->arrive_and_wait();
  T out;
  memcpy(&out, &got, sizeof(T));
  return out;
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return emu_shfl(v, (emu::ctx.lin % 32) ^ m); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d, int = 32) {
  int lane = emu::ctx.lin % 32;
  T o = emu_shfl(v, lane + d < 32 ? lane + d : lane);
  return o;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emu_shfl(v, src); }

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
// explicitly rounded operations: volatile keeps g++ from contracting them into fused multiply-adds
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline float atomicAdd(float* p, float v) { std::lock_guard<std::mutex> l(emu::g_atomic_mu); float o = *p; *p = o + v; return o; }
static inline double atomicAdd(double* p, double v) { std::lock_guard<std::mutex> l(emu::g_atomic_mu); double o = *p; *p = o + v; return o; }
static inline int atomicAdd(int* p, int v) { std::lock_guard<std::mutex> l(emu::g_atomic_mu); int o = *p; *p = o + v; return o; }
static inline int atomicMax(int* p, int v) { std::lock_guard<std::mutex> l(emu::g_atomic_mu); int o = *p; *p = std::max(o, v); return o; }
static inline int atomicCAS(int* p, int cmp, int v) { std::lock_guard<std::mutex> l(emu::g_atomic_mu); int o = *p; if (o == cmp) *p = v; return o; }
using std::max;
using std::min;

template <class F>
void emu::launch(dim3 grid, dim3 block, size_t smem, F&& f) {
  const int nthr = int(block.x * block.y * block.z);
  const long nblk = long(grid.x) * grid.y * grid.z;
  if (nthr <= 0 || nblk <= 0) return;
  g_block = block;
  g_grid = grid;
  std::barrier<> bar(nthr);
  g_block_bar = &bar;
  const int nwarp = (nthr + 31) / 32;
  g_warp_bar.clear();
  for (int w = 0; w < nwarp; ++w) g_warp_bar.emplace_back(new std::barrier<>(std::min(32, nthr - 32 * w)));
  std::vector<uint64_t> wb(size_t(nwarp) * 32);
  g_warp_buf = reinterpret_cast<uint64_t(*)[32]>(wb.data());
  void* raw = nullptr;
  if (posix_memalign(&raw, 1024, smem + 1024) != 0) abort();
  dyn_smem = static_cast<unsigned char*>(raw);
  auto body = [&](int t) {
    ctx.lin = t;
    ctx.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
    for (long b = 0; b < nblk; ++b) {
      ctx.bid = dim3(unsigned(b % grid.x), unsigned((b / grid.x) % grid.y), unsigned(b / (long(grid.x) * grid.y)));
      f();
      bar.arrive_and_wait();
    }
  };
  std::vector<std::thread> pool;
  pool.reserve(nthr);
  for (int t = 0; t < nthr; ++t) pool.emplace_back(body, t);
  for (auto& th : pool) th.join();
  free(raw);
  dyn_smem = nullptr;
}

#ifdef DKTB_EMU_IMPL
namespace emu {
thread_local Ctx ctx;
dim3 g_block, g_grid;
std::barrier<>* g_block_bar = nullptr;
std::vector<std::unique_ptr<std::barrier<>>> g_warp_bar;
uint64_t (*g_warp_buf)[32] = nullptr;
unsigned char* dyn_smem = nullptr;
std::mutex g_atomic_mu;
}  // namespace emu
#endif
