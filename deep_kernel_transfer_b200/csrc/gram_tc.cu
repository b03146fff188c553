// Gram / cross-kernel matrix on tcgen05 (the GEMM-shaped part of the GP head, reference methods/DKT.py:161, 177, 187,
// 265 through GPyTorch's LinearKernel):   out[e][m][n] = sum_d x1[e][m][d] * x2[e][n][d]
// with fp32-class accuracy from the 3xTF32 error-compensated split (a_hi b_hi + a_hi b_lo + a_lo b_hi, fp32 accumulation
// in TMEM).  Both operands are K-major as they lie in HBM (feature rows contiguous), so a 128 x 32 tile of either is one
// 2-D TMA box (SWIZZLE_128B); the tf32 hi / lo split happens in shared memory (hi in place, lo into a second tile).
//   grid (ceil(N/128) * KS, ceil(M/128), E), 192 threads: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (one
//   elected lane), warps 2-5 = splitters, then epilogue (tcgen05.ld -> masked, row-contiguous global stores).
// SPLIT-K: the tensor core's fp32 accumulation truncates instead of rounding, so a chain of n accumulating instructions
// carries a bias of ~n/2 ulp of the running sum (measured: 1.7e-5 relative over D = 1600 in one chain).  Every CTA
// therefore covers only kChunk = 64 features (2 stages) with the main term a_hi b_hi and the two correction terms in
// SEPARATE accumulators (8-instruction chains, corrections three orders of magnitude below the main sum), writes its
// partial tile, and a second kernel adds the KS partials in a fixed order with ordinary round-to-nearest fp32 adds:
// 2e-7 relative, like an FFMA Gram -- and KS times more CTAs to fill the 148 SMs with what would be 32 tiles.
// A diagonal tile of a symmetric Gram (x1 == x2, same row block) loads and splits ONE operand and uses it on both sides.
// Pipeline: full[s] (TMA landed) -> split[s] (hi / lo ready) -> MMA -> empty[s] (tcgen05.commit), kStages-deep over D / 32.
#include "dktb_common.cuh"

#ifndef DKTB_EMU
#include "tc_common.cuh"

namespace {

constexpr int kTile = 128;
constexpr int kStages = 3;
constexpr int kOpBytes = 16384;                  // one 128 x 32 fp32 operand tile
constexpr int kStageBytes = 4 * kOpBytes;        // A hi | A lo | B hi | B lo
constexpr int kSmemBytes = kStages * kStageBytes + 1024;
constexpr int kChunk = 64;                       // features per CTA (split-K granule): 2 stages

__global__ void __launch_bounds__(192, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               float* __restrict__ out, int M, int N, int D, int symmetric, int tiles_n, long split_stride,
               int* __restrict__ err) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_full[kStages], bar_split[kStages], bar_empty[kStages], bar_acc;
  __shared__ uint32_t s_tmem;
  __shared__ int s_err;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (*reinterpret_cast<volatile int*>(err) != 0) return;
  const int e = blockIdx.z, m0 = blockIdx.y * kTile, n0 = (blockIdx.x % tiles_n) * kTile;
  const int ks = blockIdx.x / tiles_n;              // split-K index: features [ks * kChunk, +kChunk)
  const bool diag = symmetric && m0 == n0;          // B is A
  const int k_begin = ks * kChunk;
  const int iters = (min(D, k_begin + kChunk) - k_begin + 31) / 32;
  out += (long)ks * split_stride;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      tc::mbar_init(&bar_full[s], 1);
      tc::mbar_init(&bar_split[s], 128);
      tc::mbar_init(&bar_empty[s], 1);
    }
    tc::mbar_init(&bar_acc, 1);
    s_err = 0;
    tc::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { tc::prefetch_tmap(&map_a); tc::prefetch_tmap(&map_b); }
  if (warp == 1) tc::tmem_alloc<256>(&s_tmem);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t d_tmem = s_tmem;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    for (int it = 0; it < iters; ++it) {
      const int s = it % kStages, ph = (it / kStages) & 1;
      if (!tc::mbar_wait(&bar_empty[s], ph ^ 1)) { s_err = 1; break; }
      if (tc::elect_one()) {
        unsigned char* st = smem + s * kStageBytes;
        tc::mbar_expect_tx(&bar_full[s], diag ? kOpBytes : 2 * kOpBytes);
        tc::tma_load_2d(st, &map_a, &bar_full[s], k_begin + it * 32, e * M + m0);
        if (!diag) tc::tma_load_2d(st + 2 * kOpBytes, &map_b, &bar_full[s], k_begin + it * 32, e * N + n0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = tc::umma_idesc(2, 128, 128, 0, 0);
    bool ok = true;
    for (int it = 0; it < iters && ok; ++it) {
      const int s = it % kStages, ph = (it / kStages) & 1;
      ok = tc::mbar_wait(&bar_split[s], ph);
      if (!ok) break;
      tc::tcgen05_fence_after();
      const uint32_t base = tc::smem_u32(smem + s * kStageBytes);
      const uint32_t bb = diag ? base : base + 2 * kOpBytes;
      if (tc::elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t a_hi = tc::umma_desc_sw128(base + k * 32, 16, 1024);
          const uint64_t a_lo = tc::umma_desc_sw128(base + kOpBytes + k * 32, 16, 1024);
          const uint64_t b_hi = tc::umma_desc_sw128(bb + k * 32, 16, 1024);
          const uint64_t b_lo = tc::umma_desc_sw128(bb + kOpBytes + k * 32, 16, 1024);
          tc::umma_tf32_ss(d_tmem + 128, a_lo, b_hi, idesc, (it | k) ? 1u : 0u);     // corrections: their own accumulator
          tc::umma_tf32_ss(d_tmem + 128, a_hi, b_lo, idesc, 1u);
          tc::umma_tf32_ss(d_tmem, a_hi, b_hi, idesc, (it | k) ? 1u : 0u);           // main term
        }
        tc::umma_commit(&bar_empty[s]);
      }
      __syncwarp();
    }
    if (!ok) s_err = 1;
    if (tc::elect_one()) tc::umma_commit(&bar_acc);
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ splitters (128 threads)
    const int ct = tid - 64;
    bool ok = true;
    for (int it = 0; it < iters && ok; ++it) {
      const int s = it % kStages, ph = (it / kStages) & 1;
      ok = tc::mbar_wait(&bar_full[s], ph);
      if (!ok) break;
      unsigned char* st = smem + s * kStageBytes;
      const int nops = diag ? 1 : 2;
      for (int op = 0; op < nops; ++op) {
        unsigned char* raw = st + op * 2 * kOpBytes;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4* pa = reinterpret_cast<float4*>(raw + j * 2048 + ct * 16);
          float4* pl = reinterpret_cast<float4*>(raw + kOpBytes + j * 2048 + ct * 16);
          const float4 v = *pa;
          float4 hi, lo;
          hi.x = tc::to_tf32_rna(v.x); hi.y = tc::to_tf32_rna(v.y); hi.z = tc::to_tf32_rna(v.z); hi.w = tc::to_tf32_rna(v.w);
          lo.x = tc::to_tf32_rna(v.x - hi.x); lo.y = tc::to_tf32_rna(v.y - hi.y);
          lo.z = tc::to_tf32_rna(v.z - hi.z); lo.w = tc::to_tf32_rna(v.w - hi.w);
          *pa = hi;
          *pl = lo;
        }
      }
      tc::fence_proxy_async_smem();
      tc::mbar_arrive(&bar_split[s]);
    }
    if (!ok) s_err = 1;
    // ------------------------------------------------------------------ epilogue: lane = row m0 + r, 128 columns
    ok = ok && tc::mbar_wait(&bar_acc, 0);
    tc::tcgen05_fence_after();
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int m = m0 + r;
    if (ok) {
      float* orow = out + ((long)e * M + m) * N + n0;
      const bool vec = (N % 4) == 0;
#pragma unroll 1
      for (int c = 0; c < kTile; c += 16) {
        uint32_t v[16], u[16];
        tc::tmem_ld16(d_tmem + ((uint32_t)(quarter * 32) << 16) + c, v);
        tc::tmem_ld16(d_tmem + 128 + ((uint32_t)(quarter * 32) << 16) + c, u);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
        if (m < M) {
          if (vec && n0 + c + 16 <= N) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              dktb_st4(orow + c + j, make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                                 __uint_as_float(v[j + 3])));
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (n0 + c + j < N) orow[c + j] = __uint_as_float(v[j]);
          }
        }
      }
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (tid == 0 && s_err) atomicExch(err, 1);
  if (warp == 1) tc::tmem_dealloc<256>(d_tmem);
}

// out[i] = sum_s partial[s][i] in a fixed order (round-to-nearest fp32 adds)
__global__ void gram_tc_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, long n, long stride,
                                      int ks) {
  const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  if (i + 4 <= n) {
    float4 t = dktb_ld4(partial + i);
    for (int s = 1; s < ks; ++s) {
      const float4 v = dktb_ld4(partial + (long)s * stride + i);
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    dktb_st4(out + i, t);
  } else {
    for (long j = i; j < n; ++j) {
      float t = partial[j];
      for (int s = 1; s < ks; ++s) t += partial[(long)s * stride + j];
      out[j] = t;
    }
  }
}

}  // namespace

DKTB_EXPORT int dktb_gram_tc_ok(int E, int M, int N, int D) {
  // TMA: row pitch multiple of 16 B; at least one full 128-row box in each operand tensor
  return D % 4 == 0 && D >= 32 && M > 0 && N > 0 && (long)E * M >= kTile && (long)E * N >= kTile;
}

// scratch floats dktb_gram_tc needs: one partial Gram per 64-feature chunk
DKTB_EXPORT long dktb_gram_tc_scratch_floats(int E, int M, int N, int D) {
  return (long)((D + kChunk - 1) / kChunk) * (((long)E * M * N + 3) / 4 * 4);
}

// Same contract as dktb_gram (out [E][M][N] = x1 [E][M][D] . x2 [E][N][D]^T) on tcgen05 / TMA; scratch: the split-K
// partials (dktb_gram_tc_scratch_floats); err: device int set to 1 if a pipeline wait timed out (zero-initialised by the
// caller).  x1 == x2 (and M == N): diagonal tiles load one operand.
DKTB_EXPORT int dktb_gram_tc(const float* x1, const float* x2, float* out, float* scratch, int* err, int E, int M, int N,
                             int D, cudaStream_t stream) {
  DKTB_CHECK_ARG(x1 && x2 && out && scratch && err && E > 0 && dktb_gram_tc_ok(E, M, N, D) && E <= 65535);
  DKTB_CHECK_ARG((long)E * M < 2147483000L && (long)E * N < 2147483000L);
  CUtensorMap map_a, map_b;
  if (tc_make_tmap_2d(&map_a, x1, (uint64_t)D, (uint64_t)E * M, 32, kTile) != 0) return DKTB_BAD_ARG - 1;
  if (tc_make_tmap_2d(&map_b, x2, (uint64_t)D, (uint64_t)E * N, 32, kTile) != 0) return DKTB_BAD_ARG - 1;
  cudaFuncSetAttribute(gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  const int ks = (D + kChunk - 1) / kChunk;
  const int tiles_n = (N + kTile - 1) / kTile;
  const long n = (long)E * M * N, stride = (n + 3) / 4 * 4;
  dim3 grid(tiles_n * ks, (M + kTile - 1) / kTile, E);
  gram_tc_kernel<<<grid, 192, kSmemBytes, stream>>>(map_a, map_b, scratch, M, N, D, (x1 == x2 && M == N) ? 1 : 0, tiles_n,
                                                    stride, err);
  int rc = dktb_launch_status();
  if (rc != 0) return rc;
  gram_tc_reduce_kernel<<<(unsigned)((n / 4 + 255) / 256 + 1), 256, 0, stream>>>(scratch, out, n, stride, ks);
  return dktb_launch_status();
}

#endif  // DKTB_EMU
