// ResNet stem on tcgen05 (reference backbone.py:336-340: Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)):
// implicit GEMM  y[pix][co] = sum_k P[pix][k] * W[co][k],  k = ci*49 + r*7 + s  (147, padded to 160 = 5 K-blocks of 32),
// 3xTF32 error-compensated (a_lo*w_hi + a_hi*w_lo + a_hi*w_hi, fp32-class accuracy).
//   CTA         one image x one band of 8 output rows; loops over the 16-column tiles of the band (tile = 8 x 16 = 128
//               output pixels = the 128 TMEM lanes)
//   patch       input rows 2*oh0-3 .. 2*oh0+17 of the three channels in shared memory, DE-INTERLEAVED by column parity:
//               an output column step is an input column step of 2, so consecutive lanes read consecutive words of one
//               parity plane (no bank conflicts); zero outside the image (the convolution's padding)
//   weights     [kb][w_hi | w_lo][64 co][32 k] (80 KB), TMA-loaded once per CTA, SWIZZLE_128B K-major B operand
//   warps 0-3   epilogue: tcgen05.ld accumulator (+ bias) -> NHWC global store
//   warps 4-11  stagers: 16 im2col slots per thread per K-block (compile-time patch offsets) -> tf32 hi / lo -> tcgen05.st
//               into one of 4 A stages (64 TMEM columns each)
//   warp 12     weight TMA;  warp 13  MMA issuer (12 x (M128, N64, K8) per K-block), TMEM owner
// The previous path (conv2d_fwd_flat_mma, mma.sync tiles fed through registers) took 3.3 ms per launch at B = 420, 224x224.
#include "dktb_common.cuh"

#ifndef DKTB_EMU
#include "tc_common.cuh"

namespace {

constexpr int kSR = 8, kSC = 16;                // output tile: rows x columns
constexpr int kSPL = 120;                       // words per parity plane of a patch row (ceil16(W/2) + 3 <= 120)
constexpr int kSRP = 2 * kSPL;                  // patch row pitch (both planes)
constexpr int kSRows = 2 * kSR + 5;             // 21 input rows per band
constexpr int kSCP = kSRows * kSRP;             // channel pitch
constexpr int kSKB = 5;                         // K-blocks of 32
constexpr int kSWBytes = kSKB * 2 * 64 * 128;   // 81 920
constexpr int kSAStages = 4;
constexpr int kSThreads = 448;

__device__ __forceinline__ void stem_split(float v, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(v) + 0x1000u) & 0xFFFFE000u;
  lo = __float_as_uint(v - __uint_as_float(hi));
}

// im2col slots KB*32 + SET*16 .. +15 of one output pixel: every patch offset is a compile-time constant relative to
// `base` = &patch[0][2*py][plane 0][ow]  (input column 2*ow - 3 + s: s even -> odd plane, word ow + s/2; s odd -> even
// plane, word ow + (s-1)/2)
template <int KB, int SET>
__device__ __forceinline__ void stem_stage(const float* __restrict__ base, uint32_t (&hi)[16], uint32_t (&lo)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int k = KB * 32 + SET * 16 + j;
    float v = 0.f;
    if (k < 147) {
      const int ci = k / 49, r = (k % 49) / 7, s = k % 7;
      v = base[ci * kSCP + r * kSRP + ((s & 1) ? 0 : kSPL) + (s >> 1)];
    }
    stem_split(v, hi[j], lo[j]);
  }
}

template <int SET>
__device__ __forceinline__ void stem_stage_kb(int kb, const float* __restrict__ base, uint32_t (&hi)[16],
                                              uint32_t (&lo)[16]) {
  switch (kb) {
    case 0: stem_stage<0, SET>(base, hi, lo); break;
    case 1: stem_stage<1, SET>(base, hi, lo); break;
    case 2: stem_stage<2, SET>(base, hi, lo); break;
    case 3: stem_stage<3, SET>(base, hi, lo); break;
    default: stem_stage<4, SET>(base, hi, lo); break;
  }
}

__global__ void __launch_bounds__(kSThreads, 1)
stem_tc_kernel(const float* __restrict__ x, const __grid_constant__ CUtensorMap map_w, const float* __restrict__ bias,
               float* __restrict__ y, int H, int W, int Ho, int Wo, int* __restrict__ err) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_w, bar_afull[kSAStages], bar_aempty[kSAStages], bar_accfull[2], bar_accempty[2];
  __shared__ uint32_t s_tmem;
  __shared__ int s_err;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (*reinterpret_cast<volatile int*>(err) != 0) return;
  unsigned char* s_wt = smem;                                          // [5][hi | lo][64][128 B]
  float* s_patch = reinterpret_cast<float*>(smem + kSWBytes);          // [3][21][2][kSPL]
  const int b = blockIdx.y, oh0 = blockIdx.x * kSR;
  const int ntile = (Wo + kSC - 1) / kSC;

  if (tid == 0) {
    tc::mbar_init(&bar_w, 1);
    for (int s = 0; s < kSAStages; ++s) { tc::mbar_init(&bar_afull[s], 256); tc::mbar_init(&bar_aempty[s], 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&bar_accfull[s], 1); tc::mbar_init(&bar_accempty[s], 128); }
    s_err = 0;
    tc::fence_barrier_init();
  }
  if (warp == 12 && lane == 0) tc::prefetch_tmap(&map_w);
  if (warp == 13) tc::tmem_alloc<512>(&s_tmem);
  // input patch (NCHW source): one (channel, row) per warp pass, word j of a parity plane <- input column 2j-2 (even
  // plane) / 2j-3 (odd plane)
  for (int rowi = warp; rowi < 3 * kSRows; rowi += kSThreads / 32) {
    const int ci = rowi / kSRows, hh = 2 * oh0 - 3 + rowi % kSRows;
    const bool rok = hh >= 0 && hh < H;
    const float* src = x + (((long)b * 3 + ci) * H + (rok ? hh : 0)) * W;
    float* dst = s_patch + rowi * kSRP;
    for (int i = lane; i < kSRP; i += 32) {
      const int par = i >= kSPL, j = i - par * kSPL;
      const int c = 2 * j - 2 - par;
      dst[i] = (rok && c >= 0 && c < W) ? src[c] : 0.f;
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t a_tmem = tmem + 128;                        // accumulators [0,128): 2 x 64 columns; A stages [128,384)

  if (warp == 12) {
    if (tc::elect_one()) {
      tc::mbar_expect_tx(&bar_w, kSWBytes);
      for (int i = 0; i < 2 * kSKB; ++i) tc::tma_load_2d(s_wt + i * 8192, &map_w, &bar_w, 0, i * 64);
    }
    __syncwarp();
  } else if (warp == 13) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = tc::umma_idesc(2, 128, 64, 0, 0);
    bool ok = tc::mbar_wait(&bar_w, 0);
    long ai = 0;
    for (int ti = 0; ti < ntile && ok; ++ti) {
      const int ab = ti & 1, ap = (ti >> 1) & 1;
      ok = tc::mbar_wait(&bar_accempty[ab], ap ^ 1);
      if (!ok) break;
      tc::tcgen05_fence_after();
      const uint32_t d_tmem = tmem + ab * 64;
      for (int kb = 0; kb < kSKB && ok; ++kb, ++ai) {
        const int sa = (int)(ai % kSAStages), pa = (int)((ai / kSAStages) & 1);
        ok = tc::mbar_wait(&bar_afull[sa], pa);
        if (!ok) break;
        tc::tcgen05_fence_after();
        const uint32_t acol = a_tmem + sa * 64;              // hi [0,32) | lo [32,64)
        const uint32_t wbase = tc::smem_u32(s_wt + kb * 16384);
        if (tc::elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t w_hi = tc::umma_desc_sw128(wbase + k * 32, 16, 1024);
            const uint64_t w_lo = tc::umma_desc_sw128(wbase + 8192 + k * 32, 16, 1024);
            tc::umma_tf32_ts(d_tmem, acol + 32 + k * 8, w_hi, idesc, (kb | k) ? 1u : 0u);      // a_lo * w_hi (small first)
            tc::umma_tf32_ts(d_tmem, acol + k * 8, w_lo, idesc, 1u);                             // a_hi * w_lo
            tc::umma_tf32_ts(d_tmem, acol + k * 8, w_hi, idesc, 1u);                             // a_hi * w_hi
          }
          tc::umma_commit(&bar_aempty[sa]);
          if (kb == kSKB - 1) tc::umma_commit(&bar_accfull[ab]);
        }
        __syncwarp();
      }
    }
    if (!ok) s_err = 1;
  } else if (warp >= 4 && warp < 12) {
    // ------------------------------------------------------------------ stagers (256 threads)
    const int set = (warp - 4) >> 2;                         // which 16 of a K-block's 32 slots
    const int quarter = warp & 3;
    const int p = quarter * 32 + lane;                       // pixel of the tile = TMEM lane
    const int py = p >> 4, px = p & 15;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const float* prow = s_patch + (2 * py) * kSRP + px;
    long ai = 0;
    bool ok = true;
    for (int ti = 0; ti < ntile && ok; ++ti) {
      const float* base = prow + ti * kSC;
      for (int kb = 0; kb < kSKB && ok; ++kb, ++ai) {
        const int sa = (int)(ai % kSAStages), pa = (int)((ai / kSAStages) & 1);
        uint32_t hi[16], lo[16];
        if (set == 0) stem_stage_kb<0>(kb, base, hi, lo);
        else stem_stage_kb<1>(kb, base, hi, lo);
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
    }
    if (!ok) s_err = 1;
  } else if (warp < 4) {
    // ------------------------------------------------------------------ epilogue (128 threads)
    const int quarter = warp & 3;
    const int p = quarter * 32 + lane;
    const int py = p >> 4, px = p & 15;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int oh = oh0 + py;
    bool ok = true;
    for (int ti = 0; ti < ntile && ok; ++ti) {
      const int ab = ti & 1, ap = (ti >> 1) & 1;
      const int ow = ti * kSC + px;
      const bool valid = oh < Ho && ow < Wo;
      float* orow = y + (((long)b * Ho + oh) * Wo + ow) * 64;
      ok = tc::mbar_wait(&bar_accfull[ab], ap);
      if (!ok) break;
      tc::tcgen05_fence_after();
#pragma unroll
      for (int c = 0; c < 64; c += 16) {
        uint32_t v[16];
        tc::tmem_ld16(tmem + ab * 64 + lane_base + c, v);
        tc::tmem_ld_wait();
        if (c == 48) {                              // accumulator fully read: hand the buffer back to the MMA warp
          tc::tcgen05_fence_before();
          tc::mbar_arrive(&bar_accempty[ab]);
        }
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                   __uint_as_float(v[j + 3]));
            if (bias) { o.x += bias[c + j]; o.y += bias[c + j + 1]; o.z += bias[c + j + 2]; o.w += bias[c + j + 3]; }
            dktb_st4(orow + c + j, o);
          }
        }
      }
    }
    if (!ok) s_err = 1;
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (tid == 0 && s_err) atomicExch(err, 1);
  if (warp == 13) tc::tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------------------------
// Stem weight gradient on tcgen05:  dW[co][k] = sum_pix G[pix][co] * P[pix][k]  (k = ci*49 + r*7 + s).  GEMM with M = the
// k slots (two M = 128 blocks: k = 0..127 | 128..146), N = 64 output channels, K = output pixels.  Persistent CTAs loop
// over (image, band of 8 output rows) work items with the forward kernel's parity-plane patch; a K-block is 32 pixels =
// 2 output rows x 16 columns of the band.
//   A (TMEM, `.ts`)  P^T: lane = k slot, column = pixel: every stager thread OWNS one k slot and reads its 2 x 16 pixels
//                    as two runs of consecutive words of one parity plane -> tf32 hi / lo -> tcgen05.st (3 stages)
//   B (smem)         G^T [64 co][32 pixels] hi / lo, K-major SWIZZLE_128B, transposed by the same threads from coalesced
//                    global loads of dy (issued before the patch gather of the same K-block), double-buffered
//   accumulators     2 x 64 TMEM columns; drained into fp32 registers every kSWgDrain bands (the tensor core's
//                    accumulation truncates: ~2e-5 per 1700 accumulating instructions) and written once per CTA
//   warps 0-7  stagers / transposers / drain;  warp 8  MMA issuer (24 x (M128, N64, K8) per K-block), TMEM owner
// The partials [CTA][256][64] are summed by stem_wgrad_reduce_kernel in a fixed order (deterministic).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kSWgThreads = 288;
constexpr int kSWgAStages = 3;
constexpr int kSWgDrain = 2;                    // bands per accumulator life: 2 x 28 K-blocks x 12 instructions

__global__ void __launch_bounds__(kSWgThreads, 1)
stem_wgrad_tc_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ partial, int B, int H,
                     int W, int Ho, int Wo, int* __restrict__ err) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_afull[kSWgAStages], bar_aempty[kSWgAStages], bar_bfull[2], bar_bempty[2], bar_acc, bar_patch;
  __shared__ uint32_t s_tmem;
  __shared__ int s_err;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (*reinterpret_cast<volatile int*>(err) != 0) return;
  unsigned char* s_b = smem;                                           // [2 stages][hi | lo][64 co][128 B]
  float* s_patch = reinterpret_cast<float*>(smem + 2 * 16384);         // [3][21][2][kSPL]
  const int nband = (Ho + kSR - 1) / kSR, ntile = (Wo + kSC - 1) / kSC;
  const long nwork = (long)B * nband;
  const int nkb_band = (kSR / 2) * ntile;                              // K-blocks per band: 4 row pairs x column tiles

  if (tid == 0) {
    for (int s = 0; s < kSWgAStages; ++s) { tc::mbar_init(&bar_afull[s], 256); tc::mbar_init(&bar_aempty[s], 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&bar_bfull[s], 256); tc::mbar_init(&bar_bempty[s], 1); }
    tc::mbar_init(&bar_acc, 1);
    tc::mbar_init(&bar_patch, 256);
    s_err = 0;
    tc::fence_barrier_init();
  }
  if (warp == 8) tc::tmem_alloc<512>(&s_tmem);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t a_tmem = tmem + 128;                        // accumulators [0,128): M-block 0 | 1; A stages 3 x 128 columns

  if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = tc::umma_idesc(2, 128, 64, 0, 0);
    bool ok = true;
    long n = 0;                                               // K-block counter
    int wcount = 0;
    for (long work = blockIdx.x; work < nwork && ok; work += gridDim.x, ++wcount) {
      const bool fresh = wcount % kSWgDrain == 0;             // first band of an accumulator life: overwrite
      for (int kb = 0; kb < nkb_band && ok; ++kb, ++n) {
        const int sb = (int)(n & 1), pb = (int)((n >> 1) & 1);
        const int sa = (int)(n % kSWgAStages), pa = (int)((n / kSWgAStages) & 1);
        ok = tc::mbar_wait(&bar_bfull[sb], pb);
        if (!ok) break;
        ok = tc::mbar_wait(&bar_afull[sa], pa);
        if (!ok) break;
        tc::tcgen05_fence_after();
        const uint32_t bbase = tc::smem_u32(s_b + sb * 16384);
        if (tc::elect_one()) {
#pragma unroll
          for (int mb = 0; mb < 2; ++mb) {
            const uint32_t acol = a_tmem + sa * 128 + mb * 64;      // hi [0,32) | lo [32,64)
            const uint32_t dcol = tmem + mb * 64;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t b_hi = tc::umma_desc_sw128(bbase + k * 32, 16, 1024);
              const uint64_t b_lo = tc::umma_desc_sw128(bbase + 8192 + k * 32, 16, 1024);
              tc::umma_tf32_ts(dcol, acol + 32 + k * 8, b_hi, idesc, (fresh && kb == 0 && k == 0) ? 0u : 1u);
              tc::umma_tf32_ts(dcol, acol + k * 8, b_lo, idesc, 1u);
              tc::umma_tf32_ts(dcol, acol + k * 8, b_hi, idesc, 1u);
            }
          }
          tc::umma_commit(&bar_aempty[sa]);
          tc::umma_commit(&bar_bempty[sb]);
        }
        __syncwarp();
      }
      // end of an accumulator life (or of this CTA's work): the stagers drain the accumulators
      if (ok && (wcount % kSWgDrain == kSWgDrain - 1 || work + gridDim.x >= nwork)) {
        if (tc::elect_one()) tc::umma_commit(&bar_acc);
        __syncwarp();
      }
    }
    if (!ok) s_err = 1;
  } else {
    // ------------------------------------------------------------------ stagers (256 threads)
    const int mb = warp >> 2;                                // M-block of this thread's k slot
    const int quarter = warp & 3;
    const int L = quarter * 32 + lane;                       // TMEM lane
    const int k = mb * 128 + L;
    const bool kval = k < 147;
    // tcgen05.st / .ld are warp-collective (.sync.aligned): whether a warp stages at all must not depend on the lane.  The
    // warp holding k = 128..159 stages all 32 lanes (zeros beyond slot 146); the warps of k >= 160 stage nothing.
    const bool wval = mb * 128 + quarter * 32 < 147;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    int koff = 0;                                            // patch offset of this k slot (see stem_stage)
    if (kval) {
      const int ci = k / 49, r = (k % 49) / 7, sx = k % 7;
      koff = ci * kSCP + r * kSRP + ((sx & 1) ? 0 : kSPL) + (sx >> 1);
    }
    const int co = tid & 63, kq8 = tid >> 6;                 // G transpose: pixels [kq8*8, +8) of channel co
    float run[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) run[j] = 0.f;
    bool ok = true;
    long n = 0;
    int wcount = 0, ndrain = 0, nsync = 0;
    // the 256 stagers meet on an mbarrier (bounded wait like every other wait of the kernel, never a bar.sync that a
    // thread leaving on a time-out would dead-lock)
    auto stager_sync = [&]() -> bool {
      tc::mbar_arrive(&bar_patch);
      const bool r = tc::mbar_wait(&bar_patch, nsync & 1);
      ++nsync;
      return r;
    };
    for (long work = blockIdx.x; work < nwork && ok; work += gridDim.x, ++wcount) {
      const int b = (int)(work / nband), oh0 = (int)(work % nband) * kSR;
      // ---- this band's input patch (all stagers; the previous band's gathers are complete: they ended in registers)
      ok = stager_sync();
      if (!ok) break;
      for (int rowi = warp; rowi < 3 * kSRows; rowi += 8) {
        const int ci = rowi / kSRows, hh = 2 * oh0 - 3 + rowi % kSRows;
        const bool rok = hh >= 0 && hh < H;
        const float* src = x + (((long)b * 3 + ci) * H + (rok ? hh : 0)) * W;
        float* dst = s_patch + rowi * kSRP;
        for (int i = lane; i < kSRP; i += 32) {
          const int par = i >= kSPL, j = i - par * kSPL;
          const int c = 2 * j - 2 - par;
          dst[i] = (rok && c >= 0 && c < W) ? src[c] : 0.f;
        }
      }
      ok = stager_sync();
      if (!ok) break;
      for (int kb = 0; kb < nkb_band && ok; ++kb, ++n) {
        const int rp = kb / ntile, ti = kb - rp * ntile;     // row pair of the band, column tile
        const int sb = (int)(n & 1), pb = (int)((n >> 1) & 1);
        const int sa = (int)(n % kSWgAStages), pa = (int)((n / kSWgAStages) & 1);
        // ---- G^T tile: pixel p of the K-block = (row rp*2 + (p >> 4), column ti*16 + (p & 15))
        float gv[8];
        {
          const int prow = kq8 >> 1, pc0 = (kq8 & 1) * 8;
          const int oh = oh0 + rp * 2 + prow;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int ow = ti * kSC + pc0 + j;
            gv[j] = (oh < Ho && ow < Wo) ? gy[(((long)b * Ho + oh) * Wo + ow) * 64 + co] : 0.f;
          }
        }
        // ---- P^T: this k slot's 32 pixels
        uint32_t hi[32], lo[32];
        if (wval) {
          const float* base = s_patch + koff + (4 * rp) * kSRP + ti * kSC;
#pragma unroll
          for (int j = 0; j < 32; ++j) stem_split(kval ? base[(j >> 4) * 2 * kSRP + (j & 15)] : 0.f, hi[j], lo[j]);
        }
        ok = tc::mbar_wait(&bar_bempty[sb], pb ^ 1);
        if (!ok) break;
        {
          unsigned char* bhi = s_b + sb * 16384 + co * 128;
          unsigned char* blo = bhi + 8192;
#pragma unroll
          for (int kq = 0; kq < 2; ++kq) {
            uint32_t h[4], l[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) stem_split(gv[kq * 4 + j], h[j], l[j]);
            const int off = ((kq8 * 2 + kq) ^ (co & 7)) << 4;
            *reinterpret_cast<uint4*>(bhi + off) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(blo + off) = make_uint4(l[0], l[1], l[2], l[3]);
          }
          tc::fence_proxy_async_smem();
          tc::mbar_arrive(&bar_bfull[sb]);
        }
        ok = tc::mbar_wait(&bar_aempty[sa], pa ^ 1);
        if (!ok) break;
        tc::tcgen05_fence_after();
        if (wval) {
          const uint32_t dst = a_tmem + sa * 128 + mb * 64 + lane_base;
          tc::tmem_st16(dst, hi);
          tc::tmem_st16(dst + 16, hi + 16);
          tc::tmem_st16(dst + 32, lo);
          tc::tmem_st16(dst + 48, lo + 16);
          tc::tmem_st_wait();
        }
        tc::tcgen05_fence_before();
        tc::mbar_arrive(&bar_afull[sa]);
      }
      if (ok && (wcount % kSWgDrain == kSWgDrain - 1 || work + gridDim.x >= nwork)) {
        ok = tc::mbar_wait(&bar_acc, ndrain & 1);
        ++ndrain;
        if (!ok) break;
        tc::tcgen05_fence_after();
#pragma unroll
        for (int c = 0; c < 64; c += 16) {
          uint32_t v[16];
          tc::tmem_ld16(tmem + mb * 64 + lane_base + c, v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) run[c + j] += __uint_as_float(v[j]);
        }
        tc::tcgen05_fence_before();
      }
    }
    if (!ok) s_err = 1;
    if (kval) {
      float* o = partial + ((long)blockIdx.x * 256 + mb * 128 + L) * 64;
#pragma unroll
      for (int j = 0; j < 64; j += 4) dktb_st4(o + j, make_float4(run[j], run[j + 1], run[j + 2], run[j + 3]));
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (tid == 0 && s_err) atomicExch(err, 1);
  if (warp == 8) tc::tmem_dealloc<512>(tmem);
}

// partial [nsplit][256 k slots][64 co] -> dw [64][147]
__global__ void stem_wgrad_reduce_kernel(const float* __restrict__ partial, int nsplit, float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 147 * 64) return;
  const int co = i % 64, k = i / 64;
  const float* p = partial + (long)k * 64 + co;
  float t = 0.f;
  for (int s = 0; s < nsplit; ++s) t += p[(long)s * 256 * 64];
  dw[co * 147 + k] = t;
}

// w [64][3][7][7] -> wb [5 kb][hi | lo][64 co][32 k]: k = ci*49 + r*7 + s = the flattened weight row, slots 147..159 zero
__global__ void prep_weights_stem_tc_kernel(const float* __restrict__ w, float* __restrict__ wb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 160) return;
  const int co = i / 160, k = i % 160;
  const float v = k < 147 ? w[co * 147 + k] : 0.f;
  const float hi = tc::to_tf32_rna(v), lo = tc::to_tf32_rna(v - hi);
  const int kb = k >> 5, kk = k & 31;
  wb[((kb * 2 + 0) * 64 + co) * 32 + kk] = hi;
  wb[((kb * 2 + 1) * 64 + co) * 32 + kk] = lo;
}

}  // namespace

DKTB_EXPORT int dktb_stem_tc_ok(int Cin, int Cout, int R, int stride, int pad, int dil, int H, int W) {
  return Cin == 3 && Cout == 64 && R == 7 && stride == 2 && pad == 3 && dil == 1 && H % 2 == 0 && W % 2 == 0 && H >= 2 &&
         W >= 2 && (W / 2 + kSC - 1) / kSC * kSC + 3 <= kSPL;      /* the last (ragged) tile's gather stays inside a plane */
}

DKTB_EXPORT long dktb_stem_tc_weight_floats(void) { return (long)kSKB * 2 * 64 * 32; }

DKTB_EXPORT int dktb_prep_weights_stem_tc(const float* w, float* wb, cudaStream_t stream) {
  DKTB_CHECK_ARG(w && wb);
  prep_weights_stem_tc_kernel<<<(64 * 160 + 255) / 256, 256, 0, stream>>>(w, wb);
  return dktb_launch_status();
}

// x [B,3,H,W] (NCHW, as the data loader delivers it) -> y [B,H/2,W/2,64] (NHWC, dense).  wb: dktb_prep_weights_stem_tc.
// err: device int, zero-initialised by the caller, set to 1 when a pipeline wait timed out.
DKTB_EXPORT int dktb_stem_tc(const float* x, const float* wb, const float* bias, float* y, int* err, int B, int H, int W,
                             cudaStream_t stream) {
  DKTB_CHECK_ARG(x && wb && y && err && B > 0 && B <= 65535);
  DKTB_CHECK_ARG(dktb_stem_tc_ok(3, 64, 7, 2, 3, 1, H, W));
  CUtensorMap map_w;
  if (tc_make_tmap_2d(&map_w, wb, 32, (uint64_t)kSKB * 2 * 64, 32, 64) != 0) return DKTB_BAD_ARG - 1;
  const int smem = kSWBytes + 3 * kSCP * 4 + 1024;
  DKTB_CHECK_ARG(smem <= 227 * 1024);
  cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int Ho = H / 2, Wo = W / 2;
  dim3 grid((Ho + kSR - 1) / kSR, B);
  stem_tc_kernel<<<grid, kSThreads, smem, stream>>>(x, map_w, bias, y, H, W, Ho, Wo, err);
  return dktb_launch_status();
}

static int stem_wgrad_nsplit(int B, int Ho) {
  const long nwork = (long)B * ((Ho + kSR - 1) / kSR);
  return (int)(nwork < 148 ? nwork : 148);
}
DKTB_EXPORT long dktb_stem_wgrad_tc_scratch_floats(int B, int H) { return (long)stem_wgrad_nsplit(B, H / 2) * 256 * 64; }

// Weight gradient of the stem: x [B,3,H,W] NCHW, gy [B,H/2,W/2,64] NHWC dense -> dw [64][3][7][7] (overwritten).
// scratch: dktb_stem_wgrad_tc_scratch_floats(B, H) floats.  err as dktb_stem_tc.
DKTB_EXPORT int dktb_stem_wgrad_tc(const float* x, const float* gy, float* dw, float* scratch, int* err, int B, int H, int W,
                                   cudaStream_t stream) {
  DKTB_CHECK_ARG(x && gy && dw && scratch && err && B > 0);
  DKTB_CHECK_ARG(dktb_stem_tc_ok(3, 64, 7, 2, 3, 1, H, W));
  const int smem = 2 * 16384 + 3 * kSCP * 4 + 1024;
  cudaFuncSetAttribute(stem_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int Ho = H / 2, Wo = W / 2;
  const int nsplit = stem_wgrad_nsplit(B, Ho);
  stem_wgrad_tc_kernel<<<nsplit, kSWgThreads, smem, stream>>>(x, gy, scratch, B, H, W, Ho, Wo, err);
  int rc = dktb_launch_status();
  if (rc != 0) return rc;
  stem_wgrad_reduce_kernel<<<(147 * 64 + 255) / 256, 256, 0, stream>>>(scratch, nsplit, dw);
  return dktb_launch_status();
}

#endif  // DKTB_EMU
