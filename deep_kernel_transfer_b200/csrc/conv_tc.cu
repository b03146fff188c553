// tcgen05 (5th-gen tensor core) 3x3 64->64 convolution over the padded-flat NHWC layout, forward and dgrad,
// with fp32-class accuracy from an error-compensated 3xTF32 split:
//     a = a_hi + a_lo  (a_hi = rna_tf32(a), a_lo = a - a_hi exact),  w likewise (split once per step by a prep kernel)
//     D += a_hi*w_hi + a_lo*w_hi + a_hi*w_lo        (fp32 accumulation in TMEM; dropped term ~2^-22)
// GEMM view per CTA: M = 128 consecutive padded-flat pixels of one image, N = 64 output channels,
// K = 9 taps x 64 input channels; tap (r,s) is the same activation matrix shifted by (r-1)*Wp + (s-1) rows, so
// every operand tile is a plain 2-D TMA box and the zero border of the layout supplies the padding.
//
// Kernels in this file: the persistent forward / dgrad kernel (A operand staged in TMEM by stager warps, weights
// streamed by TMA, dedicated epilogue warps), the tcgen05 weight-gradient kernel and the first-layer (3 -> 64) kernel.
#include "dktb_common.cuh"

#ifndef DKTB_EMU
#include <stdlib.h>
#include <cuda_bf16.h>
#include "tc_common.cuh"

namespace {

constexpr int kRows = 128;

constexpr int kHaloBox = 32;                  // rows per halo TMA box
constexpr int kTsThreads = 64 + 256;          // TMA warp, MMA warp, 8 stager / epilogue warps

// tf32 split with full-rate integer ops: hi = round-to-nearest(-away) to 10 mantissa bits, lo = rounded remainder
// DKTB_SPLIT_ROUND_LO: also round the remainder (2 more integer ops per element).  Default: hand the exact remainder
// v - hi to the tensor core, which truncates it to tf32 -- the operand is then carried to 2^-21 instead of 2^-22
// (|lo| <= 2^-11 |v|, truncation loses < 2^-10 |lo|), well inside the 1e-4 parity bar, and the stager warps (the
// bottleneck of the convolution kernels: tensor pipe 55% busy) issue 3 instead of 5 instructions per element.
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(v) + 0x1000u) & 0xFFFFE000u;
#ifdef DKTB_SPLIT_ROUND_LO
  lo = (__float_as_uint(v - __uint_as_float(hi)) + 0x1000u) & 0xFFFFE000u;
#else
  lo = __float_as_uint(v - __uint_as_float(hi));
#endif
}
// two neighbouring channels: tf32 hi parts + the two remainders packed as bf16 (low half = the first channel) for the
// kind::f16 small-term instruction
__device__ __forceinline__ void split_tf32_bf16(float v0, float v1, uint32_t& hi0, uint32_t& hi1, uint32_t& lo_pair) {
  hi0 = (__float_as_uint(v0) + 0x1000u) & 0xFFFFE000u;
  hi1 = (__float_as_uint(v1) + 0x1000u) & 0xFFFFE000u;
  const float l0 = v0 - __uint_as_float(hi0), l1 = v1 - __uint_as_float(hi1);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo_pair) : "f"(l1), "f"(l0));      // {upper, lower} = {l1, l0}
}
__device__ __forceinline__ void split_tf32_exact(float v, uint32_t& hi, uint32_t& lo) {      // weights (prepared once)
  hi = (__float_as_uint(v) + 0x1000u) & 0xFFFFE000u;
  lo = (__float_as_uint(v - __uint_as_float(hi)) + 0x1000u) & 0xFFFFE000u;
}

// ------------------------------------------------------------------------------------------------------------------
// Persistent A-in-TMEM kernel.  One CTA per SM walks a static list of (image, tile) pairs; the
// weight ring (6 x 16 KB) streams continuously across tiles, the activation halo and the accumulator are double
// buffered, and four dedicated epilogue warps drain accumulator t while the stagers / MMA already work on tile t+1:
// prologue, epilogue and L2 latency all overlap the tensor pipe.
//   warp 0      TMA producer (halo of tile t+1 as soon as its buffer is free; weight ring)
//   warp 1      MMA issuer (elected lane), TMEM owner
//   warps 2-9   stagers: halo row -> tf32 hi/lo split -> tcgen05.st into one of 6 A stages
//   warps 10-13 epilogue: tcgen05.ld accumulator -> +bias -> global store; BatchNorm partial sums by warp shuffles
// TMEM (512 columns): accumulators [0,128), A stages [128, 512).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kP_WStages = 4;                 // half-tap weight slots: [w_hi | w_lo] x 32 channels (tf32, 16 KB) and -- in the
constexpr int kP_WStageBytes = 24576;         // slot of channel half 0 -- the whole tap's w_hi in bf16 (64 x 64, 8 KB)
constexpr int kP_WFloats = 9 * 2 * 64 * 64;   // tf32 part of a prepared weight tensor; the bf16 part follows
constexpr int kP_AStages = 4;                 // half-tap A stages (32 channels): {hi 32 | lo 32} columns each; a ring of
                                              // four gives a stage three MMA groups of slack before it is needed again
constexpr int kP_Taps = 9;
constexpr int kP_Threads = 64 + 256 + 128;

__global__ void __launch_bounds__(kP_Threads, 1)
conv3x3_tc_persistent_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                             const __grid_constant__ CUtensorMap map_wb, const float* __restrict__ bias, float* __restrict__ out, float* __restrict__ partials,
                             int B, int H, int W, int halo_rows_pad, int tiles_per_img, int lo_n64,
                             int* __restrict__ err) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_hfull[2], bar_hempty[2], bar_wfull[kP_WStages], bar_wempty[kP_WStages], bar_afull[kP_AStages],
      bar_aempty[kP_AStages], bar_accfull[2], bar_accempty[2];
  __shared__ uint32_t s_tmem;
  __shared__ int s_err;
  __shared__ float s_stat[2][4][64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (*reinterpret_cast<volatile int*>(err) != 0) return;
  const int Hp = H + 2, Wp = W + 2;
  const int half_bytes = halo_rows_pad * 128;
  unsigned char* s_halo = smem;                               // [2 buffers][2 halves][halo_rows_pad][128 B]
  unsigned char* s_w = smem + 4 * half_bytes;                 // [kP_WStages][24 KB]
  const long ntiles = (long)B * tiles_per_img;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&bar_hfull[s], 1);
      tc::mbar_init(&bar_hempty[s], 256);
      tc::mbar_init(&bar_accfull[s], 1);
      tc::mbar_init(&bar_accempty[s], 128);
    }
    for (int s = 0; s < kP_WStages; ++s) { tc::mbar_init(&bar_wfull[s], 1); tc::mbar_init(&bar_wempty[s], 1); }
    for (int s = 0; s < kP_AStages; ++s) { tc::mbar_init(&bar_afull[s], 128); tc::mbar_init(&bar_aempty[s], 1); }
    s_err = 0;
    tc::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { tc::prefetch_tmap(&map_a); tc::prefetch_tmap(&map_w); tc::prefetch_tmap(&map_wb); }
  if (warp == 1) tc::tmem_alloc<512>(&s_tmem);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t a_tmem = tmem + 256;                         // accumulators [0,256): 2 buffers x (x w_hi | x w_lo)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    int t = 0;
    long wi = 0;
    bool ok = true;
    auto load_halo = [&](long tile, int tt) -> bool {
      const int hb = tt & 1, hp = (tt >> 1) & 1;
      const int img = (int)(tile / tiles_per_img), tix = (int)(tile % tiles_per_img);
      const long img_base = (long)img * Hp * Wp;
      const int q0 = (Wp + 1) + tix * kRows;
      if (!tc::mbar_wait(&bar_hempty[hb], hp ^ 1)) return false;
      if (tc::elect_one()) {
        tc::mbar_expect_tx(&bar_hfull[hb], 2 * half_bytes);
        const int row0 = (int)(img_base + q0 - (Wp + 1));
        for (int h = 0; h < 2; ++h)
          for (int r = 0; r < halo_rows_pad; r += kHaloBox)
            tc::tma_load_2d(s_halo + (hb * 2 + h) * half_bytes + r * 128, &map_a, &bar_hfull[hb], h * 32, row0 + r);
      }
      __syncwarp();
      return true;
    };
    if (blockIdx.x < ntiles) ok = load_halo(blockIdx.x, 0);
    for (long tile = blockIdx.x; tile < ntiles && ok; tile += gridDim.x, ++t) {
      for (int tap = 0; tap < kP_Taps && ok; ++tap) {
        for (int h = 0; h < 2 && ok; ++h, ++wi) {                 // wi counts half-tap weight slots
          const int s = (int)(wi % kP_WStages), ph = (int)((wi / kP_WStages) & 1);
          ok = tc::mbar_wait(&bar_wempty[s], ph ^ 1);
          if (!ok) break;
          if (tc::elect_one()) {
            unsigned char* slot = s_w + s * kP_WStageBytes;
            tc::mbar_expect_tx(&bar_wfull[s], h == 0 ? 16384 + 8192 : 16384);
            tc::tma_load_2d(slot, &map_w, &bar_wfull[s], h * 32, tap * 128);                // [w_hi | w_lo] x 32 channels
            if (h == 0) tc::tma_load_2d(slot + 16384, &map_wb, &bar_wfull[s], 0, tap * 64);  // bf16 w_hi, all 64 channels
          }
          __syncwarp();
        }
        if (tap == 1 && tile + gridDim.x < ntiles) ok = load_halo(tile + gridDim.x, t + 1);
      }
    }
    if (!ok) s_err = 1;
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: 16 x (M128, N128, K8) per tap
    const uint32_t idesc = tc::umma_idesc(2, 128, 128, 0, 0);
    // 3xTF32 needs a_hi*w_hi + a_hi*w_lo + a_lo*w_hi.  The first two are ONE tf32 instruction per 8 channels on the stacked
    // operand [w_hi | w_lo] (N = 128).  The third is small (|a_lo| <= 2^-11 |a|): it runs as kind::f16 on bf16 copies of
    // a_lo and w_hi, K = 16 channels per instruction (N = 64) -- half the instructions of the tf32 form at the same issue
    // cost; its rounding (2^-9 of a term that is 2^-11 of the product) stays at 2^-20 of the result.
    const uint32_t idesc_lo = tc::umma_idesc(1, 128, 64, 0, 0);
    int t = 0;
    long wi = 0;                                               // half-tap counter (weight slots and A stages)
    bool ok = true;
    for (long tile = blockIdx.x; tile < ntiles && ok; tile += gridDim.x, ++t) {
      const int ab = t & 1, ap = (t >> 1) & 1;
      ok = tc::mbar_wait(&bar_accempty[ab], ap ^ 1);
      if (!ok) break;
      tc::tcgen05_fence_after();
      const uint32_t d_tmem = tmem + ab * 128;
      for (int tap = 0; tap < kP_Taps && ok; ++tap) {
        int sw0 = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h, ++wi) {                     // channel halves: separately staged, separately released
          const int sw = (int)(wi % kP_WStages), pw = (int)((wi / kP_WStages) & 1);
          if (h == 0) sw0 = sw;
          ok = tc::mbar_wait(&bar_wfull[sw], pw);
          if (!ok) break;
          const uint32_t wbase = tc::smem_u32(s_w + sw * kP_WStageBytes);
          const uint32_t wb16 = tc::smem_u32(s_w + sw0 * kP_WStageBytes + 16384);     // the tap's bf16 w_hi (slot of half 0)
          const int sa = (int)(wi & 3), pa = (int)((wi >> 2) & 1);
          ok = tc::mbar_wait(&bar_afull[sa], pa);
          if (!ok) break;
          tc::tcgen05_fence_after();
          const uint32_t acol = a_tmem + sa * 64;             // a_hi tf32 [0,32) | a_lo bf16 pairs [32,48)
          if (tc::elect_one()) {
            // the four N = 128 tf32 instructions first, then the two N = 64 bf16 ones (switching shapes costs a bubble)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t w_cat = tc::umma_desc_sw128(wbase + k * 32, 16, 1024);
              tc::umma_tf32_ts(d_tmem, acol + k * 8, w_cat, idesc, (tap | h | k) ? 1u : 0u);   // a_hi * [w_hi | w_lo]
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const uint64_t w16 = tc::umma_desc_sw128(wb16 + (h * 2 + k) * 32, 16, 1024);
              tc::umma_f16_ts(d_tmem, acol + 32 + k * 8, w16, idesc_lo, 1u);                   // a_lo * w_hi (bf16, K = 16)
            }
            tc::umma_commit(&bar_aempty[sa]);
            if (h == 1) {
              tc::umma_commit(&bar_wempty[sw0]);             // half 0's slot also held the bf16 tile used until here
              tc::umma_commit(&bar_wempty[sw]);
              if (tap == kP_Taps - 1) tc::umma_commit(&bar_accfull[ab]);
            }
          }
          __syncwarp();
        }
      }
    }
    if (!ok) s_err = 1;
  } else if (warp < 10) {
    // ------------------------------------------------------------------ stagers (256 threads): 32 channels per thread
    const int ct = tid - 64;
    const int set = ct >> 7;                          // channel half of the tap handled by this thread
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    int t = 0;
    long wi = 0;
    bool ok = true;
    for (long tile = blockIdx.x; tile < ntiles && ok; tile += gridDim.x, ++t) {
      const int hb = t & 1, hp = (t >> 1) & 1;
      ok = tc::mbar_wait(&bar_hfull[hb], hp);
      if (!ok) break;
      for (int tap = 0; tap < kP_Taps && ok; ++tap, ++wi) {
        const int sa = (int)((wi & 1) << 1) | set, pa = (int)((wi >> 1) & 1);      // this channel half's own stage ring
        const int row = r + (tap / 3) * Wp + (tap % 3);
        const unsigned char* src = s_halo + (hb * 2 + set) * half_bytes + row * 128;
        uint32_t hi[32], lo[16];                       // lo: bf16 pairs (low half = even channel), 16 columns
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 v = *reinterpret_cast<const float4*>(src + ((j ^ (row & 7)) << 4));
          split_tf32_bf16(v.x, v.y, hi[4 * j + 0], hi[4 * j + 1], lo[2 * j + 0]);
          split_tf32_bf16(v.z, v.w, hi[4 * j + 2], hi[4 * j + 3], lo[2 * j + 1]);
        }
        ok = tc::mbar_wait(&bar_aempty[sa], pa ^ 1);
        if (!ok) break;
        tc::tcgen05_fence_after();
        const uint32_t dst = a_tmem + sa * 64 + lane_base;
        tc::tmem_st16(dst, hi);
        tc::tmem_st16(dst + 16, hi + 16);
        tc::tmem_st16(dst + 32, lo);
        tc::tmem_st_wait();
        tc::tcgen05_fence_before();
        tc::mbar_arrive(&bar_afull[sa]);
      }
      if (ok) tc::mbar_arrive(&bar_hempty[hb]);       // this thread no longer reads the halo buffer
    }
    if (!ok) s_err = 1;
  } else {
    // ------------------------------------------------------------------ epilogue (128 threads)
    const int et = tid - 320;
    const int quarter = warp & 3;
    const int ew = et >> 5;                           // 0..3 (slot in s_stat)
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    int t = 0;
    bool ok = true;
    for (long tile = blockIdx.x; tile < ntiles && ok; tile += gridDim.x, ++t) {
      const int ab = t & 1, ap = (t >> 1) & 1;
      const int img = (int)(tile / tiles_per_img), tix = (int)(tile % tiles_per_img);
      const long img_base = (long)img * Hp * Wp;
      const int q0 = (Wp + 1) + tix * kRows;
      ok = tc::mbar_wait(&bar_accfull[ab], ap);
      if (!ok) break;
      tc::tcgen05_fence_after();
      const int q = q0 + r;
      const int hp = q / Wp, wp = q - hp * Wp;
      const bool valid = q < Hp * Wp && hp >= 1 && hp <= H && wp >= 1 && wp <= W;
      float* orow = out + (img_base + q) * 64;
#pragma unroll
      for (int c = 0; c < 64; c += 16) {
        uint32_t v[16], u[16];
        tc::tmem_ld16(tmem + ab * 128 + lane_base + c, v);
        tc::tmem_ld16(tmem + ab * 128 + 64 + lane_base + c, u);
        tc::tmem_ld_wait();
        if (c == 48) {                                // accumulator fully read: hand the buffer back to the MMA warp
          tc::tcgen05_fence_before();
          tc::mbar_arrive(&bar_accempty[ab]);
        }
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          f[j] = valid ? (__uint_as_float(v[j]) + __uint_as_float(u[j])) + (bias ? bias[c + j] : 0.f) : 0.f;
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) dktb_st4(orow + c + j, make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]));
        }
        if (partials != nullptr) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float sm = f[j], sq = f[j] * f[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              sm += __shfl_xor_sync(0xffffffffu, sm, o);
              sq += __shfl_xor_sync(0xffffffffu, sq, o);
            }
            if (lane == 0) { s_stat[0][ew][c + j] = sm; s_stat[1][ew][c + j] = sq; }
          }
        }
      }
      if (partials != nullptr) {
        asm volatile("bar.sync 2, 128;" ::: "memory");
        const int which = et >> 6, c = et & 63;
        const float tsum = (s_stat[which][0][c] + s_stat[which][1][c]) + (s_stat[which][2][c] + s_stat[which][3][c]);
        partials[(((long)img * tiles_per_img + tix) * 2 + which) * 64 + c] = tsum;
        asm volatile("bar.sync 2, 128;" ::: "memory");
      }
    }
    if (!ok) s_err = 1;
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (tid == 0 && s_err) atomicExch(err, 1);
  if (warp == 1) tc::tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------------------------
// wgrad on tcgen05:  dW[tap][ci][co] = sum_q X[q + off(tap)][ci] * G[q][co]  over ALL padded-flat rows (both operands
// have zero borders, so images concatenate seamlessly).  GEMM view: M = 128 = 2 taps x 64 ci, N = 64 co, K = rows.
// Both operands are "MN-major" in HBM (channels contiguous), which tcgen05 only supports for 32-bit data through a
// dedicated swizzle mode; instead the stager threads transpose on the fly:
//   A = X^T goes to TMEM (lane = (tap-in-pair, ci), column = row), 3xTF32-split in registers, read straight from the
//       TMA-loaded halo [rows][64 ch] (one scalar LDS per element, conflict-free: a warp reads 32 consecutive channels);
//   B = G is re-laid-out K-major ([co][32 rows], SWIZZLE_128B pattern written by hand) as hi / lo tiles in smem.
// Each persistent CTA owns a strided set of 32-row K-blocks and keeps all five 128x64 fp32 accumulators
// (taps {0,1},{2,3},{4,5},{6,7},{8,-}) in TMEM (320 columns) + three A stages (3 x 64 columns); the partial dW of
// every CTA is written once and reduced by dktb_conv3x3_wgrad_reduce in a fixed order (deterministic).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kKR = 32;                         // rows per K-block
constexpr int kWgAStages = 3;

// 16 consecutive rows (row pitch 128 B, SWIZZLE_128B) of one channel -> tf32 hi / lo; A = swizzle phase of the first row
template <int A>
__device__ __forceinline__ void wg_gather16(const unsigned char* __restrict__ xb, const int (&off)[8], uint32_t (&hi)[16],
                                            uint32_t (&lo)[16]) {
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const float v = *reinterpret_cast<const float*>(xb + k * 128 + off[(A + k) & 7]);
    split_tf32(v, hi[k], lo[k]);
  }
}

__global__ void __launch_bounds__(kTsThreads, 1)
conv3x3_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_g,
                        float* __restrict__ partial, long total_rows, int Wp, int halo_pad, int pstride,
                        int* __restrict__ err) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_raw_full[2], bar_raw_empty[2], bar_bfull[2], bar_afull[kWgAStages], bar_aempty[kWgAStages],
      bar_acc;
  __shared__ uint32_t s_tmem;
  __shared__ int s_err;
  __shared__ float s_bias[4][64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (*reinterpret_cast<volatile int*>(err) != 0) return;   // a time-out was already reported: do not pile up waits
  const int x_half_bytes = halo_pad * 128;
  const int stage_bytes = 2 * x_half_bytes + 8192 + 16384;   // X raw | G raw | B hi | B lo
  const long nkb = (total_rows + kKR - 1) / kKR;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&bar_raw_full[s], 1);
      tc::mbar_init(&bar_raw_empty[s], 257);      // 256 stagers done reading + 1 tcgen05.commit (B consumed)
      tc::mbar_init(&bar_bfull[s], 256);
    }
    for (int s = 0; s < kWgAStages; ++s) { tc::mbar_init(&bar_afull[s], 256); tc::mbar_init(&bar_aempty[s], 1); }
    tc::mbar_init(&bar_acc, 1);
    s_err = 0;
    tc::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { tc::prefetch_tmap(&map_x); tc::prefetch_tmap(&map_g); }
  if (warp == 1) tc::tmem_alloc<512>(&s_tmem);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t a_tmem = tmem + 320;

  if (warp == 0) {
    int n = 0;
    for (long kb = blockIdx.x; kb < nkb; kb += gridDim.x, ++n) {
      const int s = n & 1, ph = (n >> 1) & 1;
      if (!tc::mbar_wait(&bar_raw_empty[s], ph ^ 1)) { s_err = 1; break; }
      if (tc::elect_one()) {
        unsigned char* st = smem + s * stage_bytes;
        tc::mbar_expect_tx(&bar_raw_full[s], 2 * x_half_bytes + 8192);
        const long q0 = kb * kKR;
        const int xrow0 = (int)(q0 - (Wp + 1));
        for (int h = 0; h < 2; ++h) {
          for (int r = 0; r < halo_pad; r += kHaloBox)
            tc::tma_load_2d(st + h * x_half_bytes + r * 128, &map_x, &bar_raw_full[s], h * 32, xrow0 + r);
          tc::tma_load_2d(st + 2 * x_half_bytes + h * 4096, &map_g, &bar_raw_full[s], h * 32, (int)q0);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    {
      const uint32_t idesc = tc::umma_idesc(2, 128, 64, 0, 0);
      bool ok = true;
      int n = 0, ai = 0;
      for (long kb = blockIdx.x; kb < nkb && ok; kb += gridDim.x, ++n) {
        const int s = n & 1, ph = (n >> 1) & 1;
        ok = tc::mbar_wait(&bar_bfull[s], ph);
        if (!ok) break;
        const uint32_t bbase = tc::smem_u32(smem + s * stage_bytes + 2 * x_half_bytes + 8192);
        for (int g = 0; g < 5 && ok; ++g, ++ai) {
          const int sa = ai % kWgAStages, pa = (ai / kWgAStages) & 1;
          ok = tc::mbar_wait(&bar_afull[sa], pa);
          if (!ok) break;
          tc::tcgen05_fence_after();
          const uint32_t acol = a_tmem + sa * 64;
          const uint32_t dcol = tmem + g * 64;
          if (tc::elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t b_hi = tc::umma_desc_sw128(bbase + k * 32, 16, 1024);
            const uint64_t b_lo = tc::umma_desc_sw128(bbase + 8192 + k * 32, 16, 1024);
            tc::umma_tf32_ts(dcol, acol + 32 + k * 8, b_hi, idesc, (n | k) ? 1u : 0u);
            tc::umma_tf32_ts(dcol, acol + k * 8, b_lo, idesc, 1u);
            tc::umma_tf32_ts(dcol, acol + k * 8, b_hi, idesc, 1u);
          }
          tc::umma_commit(&bar_aempty[sa]);
          }
          __syncwarp();
        }
        if (tc::elect_one()) tc::umma_commit(&bar_raw_empty[s]);
        __syncwarp();
      }
      if (!ok) s_err = 1;
      if (tc::elect_one()) tc::umma_commit(&bar_acc);
      __syncwarp();
    }
  } else {
    const int ct = tid - 64;                                  // 0..255
    const int set = ct >> 7;                                  // k range [set*16, +16) of the 32-row K-block
    const int quarter = warp & 3;
    const int L = quarter * 32 + lane;                        // TMEM lane: (tap-in-pair, ci)
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int ci = L & 63, tsel = L >> 6;
    const int co = ct & 63, kq8 = ct >> 6;                    // G transpose: 8 k values [kq8*8, +8) of channel co
    int xoff[8];                                              // byte offset of channel ci inside a row of swizzle phase p
#pragma unroll
    for (int p8 = 0; p8 < 8; ++p8) xoff[p8] = ((((ci & 31) >> 2) ^ p8) << 4) + (ci & 3) * 4;
    float bias_acc = 0.f;
    bool ok = true;
    int n = 0, ai = 0;
    for (long kb = blockIdx.x; kb < nkb && ok; kb += gridDim.x, ++n) {
      const int s = n & 1, ph = (n >> 1) & 1;
      ok = tc::mbar_wait(&bar_raw_full[s], ph);
      if (!ok) break;
      unsigned char* st = smem + s * stage_bytes;
      // ---- B: G[k][co] -> K-major [co][k] hi / lo (8 k values per thread)
      {
        const unsigned char* graw = st + 2 * x_half_bytes + (co >> 5) * 4096;
        unsigned char* bhi = st + 2 * x_half_bytes + 8192 + co * 128;
        unsigned char* blo = bhi + 8192;
        const int cq = (co & 31) >> 2, cr = co & 3;
#pragma unroll
        for (int kq = 0; kq < 2; ++kq) {
          uint32_t h[4], l[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = kq8 * 8 + kq * 4 + j;
            const float v = *reinterpret_cast<const float*>(graw + k * 128 + ((cq ^ (k & 7)) << 4) + cr * 4);
            bias_acc += v;
            split_tf32(v, h[j], l[j]);
          }
          const int kchunk = kq8 * 2 + kq;                   // 16-byte chunk index along K (4 floats)
          const int off = (kchunk ^ (co & 7)) << 4;
          *reinterpret_cast<uint4*>(bhi + off) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(blo + off) = make_uint4(l[0], l[1], l[2], l[3]);
        }
        tc::fence_proxy_async_smem();
        tc::mbar_arrive(&bar_bfull[s]);
      }
      // ---- A: five tap-pair groups into TMEM (this thread: 16 of the 32 K rows of its lane)
      for (int g = 0; g < 5 && ok; ++g, ++ai) {
        const int sa = ai % kWgAStages, pa = (ai / kWgAStages) & 1;
        const int tap = 2 * g + tsel;
        uint32_t hi[16], lo[16];
        if (tap < 9) {
          // rows shift .. shift+15 of channel ci: the 128-byte swizzle phase of row r is r & 7, so with the (warp-uniform)
          // phase of the first row as a template constant every offset below is a register picked at compile time
          const int shift = (tap / 3) * Wp + (tap % 3) + set * 16;
          const unsigned char* xb = st + (ci >> 5) * x_half_bytes + shift * 128;
          switch (shift & 7) {
            case 0: wg_gather16<0>(xb, xoff, hi, lo); break;
            case 1: wg_gather16<1>(xb, xoff, hi, lo); break;
            case 2: wg_gather16<2>(xb, xoff, hi, lo); break;
            case 3: wg_gather16<3>(xb, xoff, hi, lo); break;
            case 4: wg_gather16<4>(xb, xoff, hi, lo); break;
            case 5: wg_gather16<5>(xb, xoff, hi, lo); break;
            case 6: wg_gather16<6>(xb, xoff, hi, lo); break;
            default: wg_gather16<7>(xb, xoff, hi, lo); break;
          }
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) { hi[k] = 0u; lo[k] = 0u; }
        }
        ok = tc::mbar_wait(&bar_aempty[sa], pa ^ 1);
        if (!ok) break;
        tc::tcgen05_fence_after();
        const uint32_t dst = a_tmem + sa * 64 + lane_base + set * 16;
        tc::tmem_st16(dst, hi);
        tc::tmem_st16(dst + 32, lo);
        tc::tmem_st_wait();
        tc::tcgen05_fence_before();
        tc::mbar_arrive(&bar_afull[sa]);
      }
      if (ok) tc::mbar_arrive(&bar_raw_empty[s]);             // this thread is done with the raw stage
    }
    if (!ok) s_err = 1;
    ok = ok && tc::mbar_wait(&bar_acc, 0);
    tc::tcgen05_fence_after();
    float* out = partial + (long)blockIdx.x * pstride;
    if (ok) {
      for (int g = 0; g < 5; ++g) {
        const int tap = 2 * g + tsel;
#pragma unroll
        for (int cc = 0; cc < 32; cc += 16) {
          const int c = set * 32 + cc;
          uint32_t v[16];
          tc::tmem_ld16(tmem + g * 64 + lane_base + c, v);
          tc::tmem_ld_wait();
          if (tap < 9) {
            float* o = out + ((long)tap * 64 + ci) * 64 + c;
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              dktb_st4(o + j, make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                          __uint_as_float(v[j + 3])));
          }
        }
      }
    }
    s_bias[kq8][co] = bias_acc;
    tc::tcgen05_fence_before();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (kq8 == 0) out[9 * 64 * 64 + co] = (s_bias[0][co] + s_bias[1][co]) + (s_bias[2][co] + s_bias[3][co]);
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (tid == 0 && s_err) atomicExch(err, 1);
  if (warp == 1) tc::tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------------------------
// First layer (3 -> 64 channels, K = 27 padded to 32) on tcgen05.  The im2col row of every output pixel is gathered by
// the stager threads from a small NCHW input patch in shared memory, 3xTF32-split and written to TMEM (A operand);
// the weights [64 co][32 k] hi / lo sit in shared memory for the whole CTA.  12 MMAs per 128-pixel tile: the kernel is
// bound by its epilogue traffic, so three epilogues are provided:
//   kC1_Y     : write the pre-BN output y [B,H,W,64] + per-tile BatchNorm partial sums (training forward)
//   kC1_STATS : partial sums only (nothing written)
//   kC1_APPLY : BatchNorm (given mean / invstd) + ReLU + MaxPool2d(2) fused, writes the zero-bordered block output
//               [B,H/2+2,W/2+2,64] directly -- the 1.8 MB/image pre-BN tensor never goes to HBM (eval / monitoring)
// One CTA = one image x one group of 4 rows, looping over the 32-column tiles of that group (same tile numbering as
// the fp32 conv1 kernel, so the BatchNorm finalize kernel is shared).
// ------------------------------------------------------------------------------------------------------------------
enum { kC1_Y = 0, kC1_STATS = 1, kC1_APPLY = 2 };
constexpr int kC1PatchW = 88;                  // patch row pitch (floats): W + 2 <= 88
constexpr int kC1OutLd = 68;                   // epilogue staging pitch: float4 rows, conflict-free for 16 B accesses
constexpr int kC1RG = 2;                       // 4-row groups per CTA (amortises patch load / TMEM / barrier set-up)
constexpr int kC1PR = 4 * kC1RG + 2;           // patch rows per input channel

// im2col slots SET*16 .. SET*16+15 of one output pixel (k = ci*9 + r*3 + s; slots 27..31 are zero): every patch offset is
// a compile-time constant relative to `base` = &patch[py][w0 + px]
template <int SET>
__device__ __forceinline__ void c1_stage_row(const float* __restrict__ base, uint32_t (&hi)[16], uint32_t (&lo)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int k = SET * 16 + j;
    float v = 0.f;
    if (k < 27) v = base[((k / 9) * kC1PR + (k % 9) / 3) * kC1PatchW + (k % 3)];
    split_tf32(v, hi[j], lo[j]);
  }
}

__global__ void __launch_bounds__(kTsThreads, 3)
conv1_tc_kernel(const float* __restrict__ x, const __grid_constant__ CUtensorMap map_w, const float* __restrict__ bias,
                float* __restrict__ y, float* __restrict__ partials, const float* __restrict__ mean,
                const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                float* __restrict__ act, int ipe, int H, int W, int mode, int* __restrict__ err) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_w, bar_afull, bar_acc;
  __shared__ uint32_t s_tmem;
  __shared__ int s_err;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (*reinterpret_cast<volatile int*>(err) != 0) return;
  unsigned char* s_wt = smem;                                        // W1 hi 8 KB | W1 lo 8 KB
  float* s_patch = reinterpret_cast<float*>(smem + 16384);           // [3][kC1PR][kC1PatchW]
  float* s_out = s_patch + 3 * kC1PR * kC1PatchW;                    // [128][kC1OutLd]
  __shared__ float s_stat[2][2][64];
  const int b = blockIdx.y, hbase = blockIdx.x * 4 * kC1RG;
  const int TX = (W + 31) / 32, RGtot = (H + 3) / 4;
  const int ngroups = min(kC1RG, RGtot - (int)blockIdx.x * kC1RG);
  const int ntile = ngroups * TX;                                    // tiles of this CTA: (row group, 32-column tile)

  if (tid == 0) {
    tc::mbar_init(&bar_w, 1);
    tc::mbar_init(&bar_afull, 256);
    tc::mbar_init(&bar_acc, 1);
    s_err = 0;
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<128>(&s_tmem);
  // input patch: rows hbase-1 .. hbase+4*kC1RG, columns -1 .. W, zero outside the image; one warp per (channel, row)
  for (int rowi = warp; rowi < 3 * kC1PR; rowi += kTsThreads / 32) {
    const int ci = rowi / kC1PR, hh = hbase + rowi % kC1PR - 1;
    const bool rok = hh >= 0 && hh < H;
    const float* src = x + (((long)b * 3 + ci) * H + (rok ? hh : 0)) * W - 1;
    for (int c = lane; c < kC1PatchW; c += 32) s_patch[rowi * kC1PatchW + c] = (rok && c >= 1 && c <= W) ? src[c] : 0.f;
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t d_tmem = s_tmem, a_tmem = s_tmem + 64;

  if (warp == 0) {
    if (tc::elect_one()) {
      tc::mbar_expect_tx(&bar_w, 16384);
      tc::tma_load_2d(s_wt, &map_w, &bar_w, 0, 0);
      tc::tma_load_2d(s_wt + 8192, &map_w, &bar_w, 0, 64);
    }
    __syncwarp();
  } else if (warp == 1) {
    const uint32_t idesc = tc::umma_idesc(2, 128, 64, 0, 0);
    bool ok = tc::mbar_wait(&bar_w, 0);
    for (int ti = 0; ti < ntile && ok; ++ti) {
      ok = tc::mbar_wait(&bar_afull, ti & 1);
      if (!ok) break;
      tc::tcgen05_fence_after();
      const uint32_t wbase = tc::smem_u32(s_wt);
      if (tc::elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t w_hi = tc::umma_desc_sw128(wbase + k * 32, 16, 1024);
          const uint64_t w_lo = tc::umma_desc_sw128(wbase + 8192 + k * 32, 16, 1024);
          tc::umma_tf32_ts(d_tmem, a_tmem + 32 + k * 8, w_hi, idesc, k ? 1u : 0u);
          tc::umma_tf32_ts(d_tmem, a_tmem + k * 8, w_lo, idesc, 1u);
          tc::umma_tf32_ts(d_tmem, a_tmem + k * 8, w_hi, idesc, 1u);
        }
        tc::umma_commit(&bar_acc);
      }
      __syncwarp();
    }
    if (!ok) s_err = 1;
  } else {
    const int ct = tid - 64;
    const int set = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int p = quarter * 32 + lane;                       // pixel of the tile = accumulator row = TMEM lane
    const int py = p >> 5, px = p & 31;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int e = (ipe > 0) ? b / ipe : 0;
    const float* prow = s_patch + py * kC1PatchW + px;
    float* orow = s_out + p * kC1OutLd + set * 32;
    const float* brow = bias ? bias + set * 32 : nullptr;
    // epilogue roles (fixed per thread): Y store = 16 threads per pixel row; stats = (sum | sumsq, row half, channel)
    const int st_r0 = ct >> 4, st_c4 = (ct & 15) * 4;
    const int sw = ct >> 7, sh = (ct >> 6) & 1, sc = ct & 63;
    bool ok = true;
    for (int ti = 0; ti < ntile && ok; ++ti) {
      const int rg = ti / TX, tx = ti - rg * TX;
      const int w0 = tx * 32, h0 = hbase + 4 * rg;
      // ---- stage the im2col row (16 of the 32 k-slots per thread)
      uint32_t hi[16], lo[16];
      if (set == 0) c1_stage_row<0>(prow + 4 * rg * kC1PatchW + w0, hi, lo);
      else c1_stage_row<1>(prow + 4 * rg * kC1PatchW + w0, hi, lo);
      tc::tmem_st16(a_tmem + lane_base + set * 16, hi);
      tc::tmem_st16(a_tmem + lane_base + 32 + set * 16, lo);
      tc::tmem_st_wait();
      tc::tcgen05_fence_before();
      tc::mbar_arrive(&bar_afull);
      // ---- accumulator (+ bias) -> smem tile; pixels outside the image are stored as zeros
      ok = tc::mbar_wait(&bar_acc, ti & 1);
      if (!ok) break;
      tc::tcgen05_fence_after();
      const bool valid = (h0 + py < H) && (w0 + px < W);
#pragma unroll
      for (int cc = 0; cc < 32; cc += 16) {
        uint32_t v[16];
        tc::tmem_ld16(d_tmem + lane_base + set * 32 + cc, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                 __uint_as_float(v[j + 3]));
          if (brow) {
            const float4 bv = dktb_ld4(brow + cc + j);
            o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
          }
          if (!valid) o = make_float4(0.f, 0.f, 0.f, 0.f);
          dktb_st4(orow + cc + j, o);
        }
      }
      tc::tcgen05_fence_before();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (mode == kC1_Y) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int oh = h0 + (it >> 1), ow = w0 + st_r0 + 16 * (it & 1);
          if (oh < H && ow < W)
            dktb_st4(y + (((long)b * H + oh) * W + ow) * 64 + st_c4,
                     dktb_ld4(s_out + (st_r0 + 16 * it) * kC1OutLd + st_c4));
        }
      }
      if (mode != kC1_APPLY && partials != nullptr) {
        const float* col = s_out + (sh * 64) * kC1OutLd + sc;
        float t = 0.f;
        if (sw) {
#pragma unroll 16
          for (int rr = 0; rr < 64; ++rr) { const float v = col[rr * kC1OutLd]; t = fmaf(v, v, t); }
        } else {
#pragma unroll 16
          for (int rr = 0; rr < 64; ++rr) t += col[rr * kC1OutLd];
        }
        s_stat[sw][sh][sc] = t;
      }
      if (mode == kC1_APPLY) {
        const int c4 = (ct & 15) * 4, pp = ct >> 4;          // pooled column pp (0..15), both pooled rows
        const int Ho = H / 2, Wo = W / 2;
        const float4 g = dktb_ld4(gamma + c4), bt = dktb_ld4(beta + c4);
        const float4 m = dktb_ld4(mean + e * 64 + c4), is = dktb_ld4(invstd + e * 64 + c4);
        const float sc4[4] = {g.x * is.x, g.y * is.y, g.z * is.z, g.w * is.w};
        const float mm[4] = {m.x, m.y, m.z, m.w}, bb[4] = {bt.x, bt.y, bt.z, bt.w};
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
          const int ho = (h0 >> 1) + pr, wo = (w0 >> 1) + pp;
          if (ho < Ho && wo < Wo) {
            float best[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
              for (int dx = 0; dx < 2; ++dx) {
                const float4 q = dktb_ld4(s_out + ((2 * pr + dy) * 32 + 2 * pp + dx) * kC1OutLd + c4);
                const float src[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) best[j] = fmaxf(best[j], fmaf(src[j] - mm[j], sc4[j], bb[j]));
              }
            dktb_st4(act + (((long)b * (Ho + 2) + ho + 1) * (Wo + 2) + wo + 1) * 64 + c4,
                     make_float4(best[0], best[1], best[2], best[3]));
          }
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (mode != kC1_APPLY && partials != nullptr && ct < 128) {
        const long blk = ((long)b * RGtot + (blockIdx.x * kC1RG + rg)) * TX + tx;
        const int w2 = ct >> 6;                               // 0: sum, 1: sum of squares
        partials[(blk * 2 + w2) * 64 + sc] = s_stat[w2][0][sc] + s_stat[w2][1][sc];
      }
    }
    if (!ok) s_err = 1;
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (tid == 0 && s_err) atomicExch(err, 1);
  if (warp == 1) tc::tmem_dealloc<128>(d_tmem);
}

// w1 [64][3][3][3] -> wb1 [2][64][32]: hi / lo, k = ci*9 + r*3 + s, columns 27..31 zero
__global__ void prep_weights_conv1_tc_kernel(const float* __restrict__ w, float* __restrict__ wb1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 32) return;
  const int co = i / 32, k = i % 32;
  const float v = k < 27 ? w[co * 27 + k] : 0.f;
  uint32_t h, l;
  split_tf32_exact(v, h, l);
  wb1[i] = __uint_as_float(h);
  wb1[64 * 32 + i] = __uint_as_float(l);
}

// w_ref [co][ci][3][3] -> wb_fwd / wb_dgrad [tap][hl][n][k] (hi = tf32-rounded, lo = rounded remainder): the hi and lo
// tiles of one tap are adjacent so that they also form one N = 128 operand [w_hi | w_lo]
__global__ void prep_weights_tc_kernel(const float* __restrict__ w, float* __restrict__ wb_fwd,
                                       float* __restrict__ wb_dgrad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 64 * 9) return;
  const int tap = i % 9, ci = (i / 9) % 64, co = i / (9 * 64);
  const float v = w[i];
  const float hi = tc::to_tf32_rna(v), lo = tc::to_tf32_rna(v - hi);
  const unsigned short hb = __bfloat16_as_ushort(__float2bfloat16_rn(hi));      // bf16 copy of w_hi for the small term
  if (wb_fwd) {
    wb_fwd[((tap * 2 + 0) * 64 + co) * 64 + ci] = hi;
    wb_fwd[((tap * 2 + 1) * 64 + co) * 64 + ci] = lo;
    reinterpret_cast<unsigned short*>(wb_fwd + kP_WFloats)[(tap * 64 + co) * 64 + ci] = hb;
  }
  if (wb_dgrad) {
    wb_dgrad[(((8 - tap) * 2 + 0) * 64 + ci) * 64 + co] = hi;
    wb_dgrad[(((8 - tap) * 2 + 1) * 64 + ci) * 64 + co] = lo;
    reinterpret_cast<unsigned short*>(wb_dgrad + kP_WFloats)[((8 - tap) * 64 + ci) * 64 + co] = hb;
  }
}

}  // namespace

// floats per prepared weight tensor: [9][hi | lo][64][64] tf32 + [9][64][64] bf16 (w_hi)
DKTB_EXPORT long dktb_conv3x3_tc_weight_floats(void) { return (long)kP_WFloats + 9 * 64 * 64 / 2; }

DKTB_EXPORT int dktb_prep_weights_tc(const float* w, float* wb_fwd, float* wb_dgrad, cudaStream_t stream) {
  DKTB_CHECK_ARG(w != nullptr);
  prep_weights_tc_kernel<<<(64 * 64 * 9 + 255) / 256, 256, 0, stream>>>(w, wb_fwd, wb_dgrad);
  return dktb_launch_status();
}

// 64 -> 64 3x3 convolution (forward: wb = the forward tensor of dktb_prep_weights_tc + bias + BatchNorm partial sums;
// dgrad: the flipped / transposed tensor, bias = partials = NULL) over the padded-flat layout [B][H+2][W+2][64].
// Persistent tcgen05 kernel, one CTA per SM.  err: device int, set to 1 if a pipeline wait timed out (zero-initialised
// by the caller; checked by ConvNetEngine.check_tc()).
DKTB_EXPORT int dktb_conv3x3_tc_fwd(const float* a, const float* wb, const float* bias, float* out, float* partials,
                                     int* err, int B, int H, int W, cudaStream_t stream) {
  DKTB_CHECK_ARG(a && wb && out && err && B > 0 && H > 0 && W > 0);
  const int Hp = H + 2, Wp = W + 2;
  const long rows = (long)B * Hp * Wp;
  DKTB_CHECK_ARG(rows < 2147483000L);
  const int halo = kRows + 2 * (Wp + 1);
  const int halo_pad = (halo + kHaloBox - 1) / kHaloBox * kHaloBox;
  const int smem = 4 * halo_pad * 128 + kP_WStages * kP_WStageBytes + 1024;
  DKTB_CHECK_ARG(smem <= 227 * 1024);
  CUtensorMap map_a, map_w, map_wb;
  if (tc_make_tmap_2d(&map_a, a, 64, (uint64_t)rows, 32, kHaloBox) != 0) return DKTB_BAD_ARG - 1;
  if (tc_make_tmap_2d(&map_w, wb, 64, 2 * 9 * 64, 32, 128) != 0) return DKTB_BAD_ARG - 1;
  if (tc_make_tmap_2d_bf16(&map_wb, wb + kP_WFloats, 64, 9 * 64, 64, 64) != 0) return DKTB_BAD_ARG - 1;
  cudaFuncSetAttribute(conv3x3_tc_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int span = Hp * Wp - 2 * (Wp + 1);
  const int tiles_per_img = (span + kRows - 1) / kRows;
  const long ntiles = (long)B * tiles_per_img;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(ntiles < sms ? ntiles : sms);
  const int lo_n64 = 1;      // the a_lo pass covers only the w_hi half of the stacked operand (N = 64)
  conv3x3_tc_persistent_kernel<<<grid, kP_Threads, smem, stream>>>(map_a, map_w, map_wb, bias, out, partials, B, H, W,
                                                                   halo_pad, tiles_per_img, lo_n64, err);
  return dktb_launch_status();
}

DKTB_EXPORT int dktb_prep_weights_conv1_tc(const float* w, float* wb1, cudaStream_t stream) {
  DKTB_CHECK_ARG(w && wb1);
  prep_weights_conv1_tc_kernel<<<8, 256, 0, stream>>>(w, wb1);
  return dktb_launch_status();
}

// First-layer convolution on tcgen05.  mode 0: y + partials (same tile numbering as dktb_conv1_fwd); mode 1: partials only;
// mode 2: fused BatchNorm(mean, invstd [B/ipe or 1][64]) + ReLU + MaxPool2d(2) -> act [B,H/2+2,W/2+2,64] (zero border kept).
DKTB_EXPORT int dktb_conv1_tc(const float* x, const float* wb1, const float* bias, float* y, float* partials,
                              const float* mean, const float* invstd, const float* gamma, const float* beta, float* act,
                              int* err, int B, int H, int W, int ipe, int mode, cudaStream_t stream) {
  DKTB_CHECK_ARG(x && wb1 && err && B > 0 && H > 0 && W > 0 && W + 2 <= kC1PatchW && B <= 65535 && mode >= 0 && mode <= 2);
  DKTB_CHECK_ARG(mode != kC1_Y || y);
  DKTB_CHECK_ARG(mode != kC1_APPLY || (mean && invstd && gamma && beta && act));
  CUtensorMap map_w;
  if (tc_make_tmap_2d(&map_w, wb1, 32, 128, 32, 64) != 0) return DKTB_BAD_ARG - 1;
  const int smem = 16384 + 3 * kC1PR * kC1PatchW * 4 + kRows * kC1OutLd * 4 + 1024;
  cudaFuncSetAttribute(conv1_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  dim3 grid(((H + 3) / 4 + kC1RG - 1) / kC1RG, B);
  conv1_tc_kernel<<<grid, kTsThreads, smem, stream>>>(x, map_w, bias, y, partials, mean, invstd, gamma, beta, act, ipe, H,
                                                       W, mode, err);
  return dktb_launch_status();
}

extern "C" int dktb_conv3x3_wgrad_reduce(const float* partial, int nsplit, float* dw, float* db, cudaStream_t stream);

// tcgen05 wgrad: same contract as dktb_conv3x3_wgrad (scratch >= dktb_conv3x3_wgrad_scratch_floats()), plus err.
DKTB_EXPORT int dktb_conv3x3_wgrad_tc(const float* a, const float* gy, float* dw, float* db, float* scratch, int* err,
                                      int B, int H, int W, cudaStream_t stream) {
  DKTB_CHECK_ARG(a && gy && dw && scratch && err && B > 0);
  const int Hp = H + 2, Wp = W + 2;
  const long rows = (long)B * Hp * Wp;
  DKTB_CHECK_ARG(rows < 2147483000L);
  const int halo = kKR + 2 * (Wp + 1);
  const int halo_pad = (halo + kHaloBox - 1) / kHaloBox * kHaloBox;
  const int stage = 2 * halo_pad * 128 + 8192 + 16384;
  const int smem = 2 * stage + 1024;
  DKTB_CHECK_ARG(smem <= 227 * 1024);
  CUtensorMap map_x, map_g;
  if (tc_make_tmap_2d(&map_x, a, 64, (uint64_t)rows, 32, kHaloBox) != 0) return DKTB_BAD_ARG - 1;
  if (tc_make_tmap_2d(&map_g, gy, 64, (uint64_t)rows, 32, kKR) != 0) return DKTB_BAD_ARG - 1;
  cudaFuncSetAttribute(conv3x3_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const long nkb = (rows + kKR - 1) / kKR;
  const int pstride = 9 * 64 * 64 + 64;
  const int nsplit = (int)(nkb < 148 ? nkb : 148);
  conv3x3_wgrad_tc_kernel<<<nsplit, kTsThreads, smem, stream>>>(map_x, map_g, scratch, rows, Wp, halo_pad, pstride, err);
  int rc = dktb_launch_status();
  if (rc != 0) return rc;
  return dktb_conv3x3_wgrad_reduce(scratch, nsplit, dw, db, stream);
}

#endif  // DKTB_EMU
