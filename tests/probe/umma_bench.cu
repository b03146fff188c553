// Hardware micro-benchmark (test infrastructure): cycles per tcgen05.mma for the tile shapes of the convolution kernels.
// One CTA issues `n` MMAs (M=128) back to back from one elected thread and waits for the final commit; reports
// cycles/MMA for kind::tf32 (K=8) and kind::f16 (K=16), N in {64,128,256}, A from smem (ss) or TMEM (ts), and with the
// MMAs rotating over 1, 2 or 4 independent accumulators (does the accumulate chain serialise?).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "tc_common.cuh"

__device__ __forceinline__ void umma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
               "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int NACC>
__global__ void __launch_bounds__(128) bench_kernel(long long* out, int n, int kind_f16, int N, int ts) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid / 32;
  for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc<512>(&s_tmem);
  tc::fence_proxy_async_smem();
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = s_tmem;
  if (warp == 0) {
    const uint32_t idesc = tc::umma_idesc(kind_f16 ? 1 : 2, 128, N, 0, 0);
    const uint32_t abase = tc::smem_u32(smem), bbase = tc::smem_u32(smem + 16384);
    const uint32_t a_t = tmem + 448;                 // A operand columns for the ts variant
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {              // rep 0 = warm-up
      t0 = clock64();
      if (tc::elect_one()) {
        uint64_t da[4], db[4];
        uint32_t dd[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          da[j] = tc::umma_desc_sw128(abase + j * 32, 16, 1024);
          db[j] = tc::umma_desc_sw128(bbase + j * 32, 16, 1024);
          dd[j] = tmem + (uint32_t)(j % NACC) * (uint32_t)N;
        }
        const int mode = kind_f16 * 2 + ts;
        for (int i = 0; i < n; i += 8) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int j = u & 3;
            if (mode == 0) tc::umma_tf32_ss(dd[j], da[j], db[j], idesc, 1u);
            else if (mode == 1) tc::umma_tf32_ts(dd[j], a_t + j * 8, db[j], idesc, 1u);
            else if (mode == 2) umma_f16_ss(dd[j], da[j], db[j], idesc, 1u);
            else umma_f16_ts(dd[j], a_t + j * 8, db[j], idesc, 1u);
          }
        }
        tc::umma_commit(&bar);
      }
      __syncwarp();
      tc::mbar_wait(&bar, rep & 1);
      t1 = clock64();
    }
    if (tid == 0) out[0] = t1 - t0;
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(bench_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
  cudaFuncSetAttribute(bench_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
  cudaFuncSetAttribute(bench_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
  const int n = 2048;
  printf("kind  N   nacc mode  cycles/MMA  MAC/clk\n");
  for (int kind = 0; kind < 2; ++kind)
    for (int N : {64, 128, 256})
      for (int nacc : {1, 2, 4}) {
        if (N * nacc > 448) continue;
        for (int ts = 0; ts < 2; ++ts) {
          if (nacc == 1) bench_kernel<1><<<1, 128, 66560>>>(d, n, kind, N, ts);
          else if (nacc == 2) bench_kernel<2><<<1, 128, 66560>>>(d, n, kind, N, ts);
          else bench_kernel<4><<<1, 128, 66560>>>(d, n, kind, N, ts);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long c;
          cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
          const double per = (double)c / n;
          const double macs = 128.0 * N * (kind ? 16 : 8);
          printf("%-5s %-3d %-4d %-4s %9.1f  %8.0f\n", kind ? "bf16" : "tf32", N, nacc, ts ? "ts" : "ss", per, macs / per);
        }
      }
  return 0;
}
