// tcgen05 convolution for the ResNet layers (reference backbone.py:135-247 SimpleBlock / BottleneckBlock): any
// Cin, Cout that are multiples of 64 (64 ... 2048), 3x3 / stride 1 / pad 1 or 1x1 / stride 1, forward and dgrad, over the
// padded-flat NHWC layout [B][H+2][W+2][C] (zero border) -- the same GEMM view and the same 3xTF32 error-compensated
// arithmetic as the 64 -> 64 kernel of csrc/conv_tc.cu, generalised:
//   work item = (128-pixel tile, NB-channel output block cb), NB = 64 or 128 (tcg_nb below);  K = (Cin / 64 input groups g)
//   x taps x 64 channels:   out[q][cb*NB + n] = sum_g sum_tap sum_k  A[q + off(tap)][g*64 + k] * W[cb*NB + n][g*64 + k][tap]
// The activation halo of (tile, g) is one pair of 2-D TMA boxes (channel coordinate g*64 + {0, 32}); the weight ring
// streams 32 KB stages of [w_hi | w_lo]; K is cut into chunks (576 for 3x3, 128 for 1x1), each accumulated in a fresh TMEM
// accumulator and summed in fp32 registers by the epilogue.  A 1x1 convolution is the same kernel with one tap (a plain
// GEMM over dense rows).  Warp roles are whole warpgroups so that setmaxnreg can move registers to where they are needed:
//   warps 0-3   epilogue (176 registers): tcgen05.ld accumulator -> NB running sums -> +bias -> global store
//   warps 4-11  stagers (136): halo row -> tf32 hi / lo split -> tcgen05.st into one of 4 A stages
//   warp 12     TMA producer (64): halo of the next (tile, group) as soon as its buffer is free; weight ring
//   warp 13     MMA issuer (elected lane), TMEM owner;  warps 14-15 idle (they only give their registers away)
// Measured at B = 420 (tests/probe/time_conv_tcg.py): NB = 128 is 1.15-1.3x the NB = 64 item on every layer with
// Cin >= 128 (each staged A tile feeds twice the MMA work), e.g. 128 -> 128 at 28x28 0.580 -> 0.488 ms.
// Work items are ordered tile-major / cb-minor so that the CTAs resident at any time share their halos through L2.
#include "dktb_common.cuh"

#ifndef DKTB_EMU
#include "tc_common.cuh"

namespace {

constexpr int kRows = 128;
constexpr int kHaloBox = 32;
constexpr int kWStages = 3;
constexpr int kWStageBytes = 32768;           // [half 0 | half 1] x [w_hi | w_lo] x 64 rows x 128 B
constexpr int kAStages = 4;
// warp roles by warpgroup (setmaxnreg moves registers between whole warpgroups): 0 = epilogue (4 warps), 1-2 = stagers
// (8 warps), 3 = TMA producer (warp 12), MMA issuer (warp 13), two idle warps
constexpr int kThreads = 512;
constexpr int kRegsEpilogue = 176, kRegsStager = 136, kRegsControl = 64;      // (176 + 2 x 136 + 64) x 128 = 65536 registers

// output channels per work item: 128 when they divide and K spans at least two input groups (with one group the kernel is
// bound by writing the tile, and the narrower item spreads that over more CTAs).  Decides the weight tensor's layout too.
__host__ __device__ inline int tcg_nb(int cout, int cin) { return (cout % 128 == 0 && cin >= 128) ? 128 : 64; }

// tf32 split: hi = round-to-nearest(-away) to 10 mantissa bits, lo = the exact remainder (the tensor core truncates it to
// tf32: the operand is carried to 2^-21).  Rounding the remainder instead (2^-23) was measured: no change of the
// end-to-end ResNet50 gradient error, 3 % slower -- the operand split is not what limits the accuracy.
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(v) + 0x1000u) & 0xFFFFE000u;
  lo = __float_as_uint(v - __uint_as_float(hi));
}

// Structural zeros of the weight tensor: the K loop visits tap t of input group g (output block cb) only when bit t of the
// mask of its PLANE is set.  Ordinary layers: one plane, every tap.  A stride-2 3x3 convolution runs as a stride-1
// convolution over the space-to-depth input (4 parity planes = 4 C channels): 9 of the 36 (plane, tap) pairs are non-zero
// (forward: plane of the input group; dgrad: plane of the output block).
struct TapMasks {
  unsigned long long m;      // 9-bit tap mask per plane, 16 bits apart (packed: no dynamically indexed kernel parameter)
  int per_plane;             // input groups (by_cb = 0) or output blocks (by_cb = 1) per plane
  int by_cb;
  __host__ __device__ unsigned of(int g, int cb) const {
    return (unsigned)(m >> (16 * ((by_cb ? cb : g) / per_plane))) & 0x1FFu;
  }
};

// NB = output channels per work item.  NB = 64: the three 3xTF32 products are a_hi x [w_hi | w_lo] (one N = 128
// instruction on the stacked tile) + a_lo x w_hi (N = 64), accumulator = 64 + 64 columns added in the epilogue, one
// accumulator per K chunk summed in registers.  NB = 128 (Cout % 128 == 0): a_hi x w_hi, a_hi x w_lo, a_lo x w_hi are
// three full-rate N = 128 instructions into ONE 128-column accumulator, and every staged A tile serves twice as many
// output channels (the stager / TMEM round trip, not the MMA issue, bounds these kernels).
// Weight stage (32 KB) = [w_hi NB rows | w_lo NB rows] x 32 channels x 4 B for both channel halves of one (group, tap)
// (NB = 64) or for one half (NB = 128).
template <int NB>
__global__ void __launch_bounds__(kThreads, 1)
conv_tcg_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                const float* __restrict__ bias, float* __restrict__ out, int B, int H, int W, int Cout, int G, int nblk,
                int ntaps, int halo_rows_pad, int tiles_per_img, int flat, int gchunk, TapMasks tm,
                int* __restrict__ err) {
  constexpr int kHS = NB == 64 ? 2 : 1;                       // channel halves per 32 KB weight stage
  constexpr int kWRing = kWStages;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_hfull[2], bar_hempty[2], bar_wfull[kWRing], bar_wempty[kWRing], bar_afull[kAStages],
      bar_aempty[kAStages], bar_accfull[2], bar_accempty[2];
  __shared__ uint32_t s_tmem;
  __shared__ int s_err;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (*reinterpret_cast<volatile int*>(err) != 0) return;
  const int Hp = H + 2, Wp = W + 2;
  const int lead = ntaps == 9 ? Wp + 1 : 0;                  // rows of halo in front of the tile
  const int half_bytes = halo_rows_pad * 128;
  unsigned char* s_halo = smem;                               // [2 buffers][2 halves][halo_rows_pad][128 B]
  unsigned char* s_w = smem + 4 * half_bytes;                 // [kWRing][kWStageBytes]
  const long nwork = (long)B * tiles_per_img * nblk;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&bar_hfull[s], 1);
      tc::mbar_init(&bar_hempty[s], 256);
      tc::mbar_init(&bar_accfull[s], 1);
      tc::mbar_init(&bar_accempty[s], 128);
    }
    for (int s = 0; s < kWRing; ++s) { tc::mbar_init(&bar_wfull[s], 1); tc::mbar_init(&bar_wempty[s], 1); }
    for (int s = 0; s < kAStages; ++s) { tc::mbar_init(&bar_afull[s], 128); tc::mbar_init(&bar_aempty[s], 1); }
    s_err = 0;
    tc::fence_barrier_init();
  }
  if (warp == 12 && lane == 0) { tc::prefetch_tmap(&map_a); tc::prefetch_tmap(&map_w); }
  if (warp == 13) tc::tmem_alloc<512>(&s_tmem);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t a_tmem = tmem + 256;                         // accumulators [0,256): 2 buffers x 128 columns
  // the epilogue keeps NB fp32 running sums per thread: it takes the registers the control warpgroup does not need
  if (warp >= 12) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(kRegsControl));
  }
  if (warp == 12) {
    // ------------------------------------------------------------------ TMA producer
    long vt = 0;                                              // (work item, group) counter: halo double buffering
    long wi = 0;                                              // weight stage counter: (group, tap, half)
    bool ok = true;
    auto load_halo = [&](long work, int g, long vv) -> bool {
      const int hb = (int)(vv & 1), hp = (int)((vv >> 1) & 1);
      const long tile = work / nblk;
      const int img = (int)(tile / tiles_per_img), tix = (int)(tile % tiles_per_img);
      const long img_base = flat ? 0 : (long)img * Hp * Wp;
      const int q0 = flat ? tix * kRows : (Wp + 1) + tix * kRows;
      if (!tc::mbar_wait(&bar_hempty[hb], hp ^ 1)) return false;
      if (tc::elect_one()) {
        tc::mbar_expect_tx(&bar_hfull[hb], 2 * half_bytes);
        const int row0 = (int)(img_base + q0 - lead);
        for (int h = 0; h < 2; ++h)
          for (int r = 0; r < halo_rows_pad; r += kHaloBox)
            tc::tma_load_2d(s_halo + (hb * 2 + h) * half_bytes + r * 128, &map_a, &bar_hfull[hb], g * 64 + h * 32, row0 + r);
      }
      __syncwarp();
      return true;
    };
    if (blockIdx.x < nwork) ok = load_halo(blockIdx.x, 0, 0);
    for (long work = blockIdx.x; work < nwork && ok; work += gridDim.x) {
      const int cb = (int)(work % nblk);
      for (int g = 0; g < G && ok; ++g, ++vt) {
        const unsigned mask = tm.of(g, cb);
        const int pre = min(2, __popc(mask));                 // the next halo is requested after this many taps' weights
        int nv = 0;
        for (int tap = 0; tap < ntaps && ok; ++tap) {
          if (!((mask >> tap) & 1)) continue;
          for (int h = 0; h < 2 && ok; h += kHS, ++wi) {
            const int s = (int)(wi % kWRing), ph = (int)((wi / kWRing) & 1);
            ok = tc::mbar_wait(&bar_wempty[s], ph ^ 1);
            if (!ok) break;
            if (tc::elect_one()) {
              const int wrow = (((cb * G + g) * ntaps) + tap) * (2 * NB);
              tc::mbar_expect_tx(&bar_wfull[s], kWStageBytes);
#pragma unroll
              for (int j = 0; j < kHS; ++j)
                tc::tma_load_2d(s_w + s * kWStageBytes + j * (kWStageBytes / kHS), &map_w, &bar_wfull[s], (h + j) * 32, wrow);
            }
            __syncwarp();
          }
          if (++nv == pre && ok) {
            if (g + 1 < G) ok = load_halo(work, g + 1, vt + 1);
            else if (work + gridDim.x < nwork) ok = load_halo(work + gridDim.x, 0, vt + 1);
          }
        }
      }
    }
    if (!ok) s_err = 1;
  } else if (warp == 13) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = tc::umma_idesc(2, 128, 128, 0, 0);
    const uint32_t idesc_lo = tc::umma_idesc(2, 128, 64, 0, 0);      // NB = 64: a_lo only meets the w_hi half of the stacked tile
    int t = 0;                                                // accumulator use counter (one per K chunk)
    long wi = 0;                                              // half-tap counter: A stages
    long ws = 0;                                              // weight stage counter
    bool ok = true;
    for (long work = blockIdx.x; work < nwork && ok; work += gridDim.x) {
      // K is cut into chunks of `gchunk` input groups; every chunk gets a FRESH TMEM accumulator that the epilogue warps
      // drain into NB fp32 running sums in registers: the tensor core's own
      // accumulation does not round to nearest, and over K = 4608 (512 channels x 9 taps) that bias reaches 2e-5
      for (int g0 = 0; g0 < G && ok; g0 += gchunk, ++t) {
        const int ab = t & 1, ap = (t >> 1) & 1;
        ok = tc::mbar_wait(&bar_accempty[ab], ap ^ 1);
        if (!ok) break;
        tc::tcgen05_fence_after();
        const uint32_t d_tmem = tmem + ab * 128;
        int nk = 0;                                           // (group, tap) steps of this chunk that are not masked out
        {
          const int cb = (int)(work % nblk);
          for (int g = g0; g < min(G, g0 + gchunk); ++g) nk += __popc(tm.of(g, cb) & ((1u << ntaps) - 1));
        }
        for (int kk = 0; kk < nk && ok; ++kk) {
          int sw = 0;
#pragma unroll
          for (int h = 0; h < 2; ++h, ++wi) {
            if (h % kHS == 0) {
              sw = (int)(ws % kWRing);
              ok = tc::mbar_wait(&bar_wfull[sw], (int)((ws / kWRing) & 1));
              ++ws;
              if (!ok) break;
            }
            const uint32_t wbase = tc::smem_u32(s_w + sw * kWStageBytes + (h % kHS) * (kWStageBytes / kHS));
            const int sa = (int)(wi & 3), pa = (int)((wi >> 2) & 1);
            ok = tc::mbar_wait(&bar_afull[sa], pa);
            if (!ok) break;
            tc::tcgen05_fence_after();
            const uint32_t acol = a_tmem + sa * 64;             // hi [0,32) | lo [32,64)
            if (tc::elect_one()) {
              if constexpr (NB == 64) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {     // the N = 128 instructions first, then the N = 64 ones (see conv_tc.cu)
                  const uint64_t w_cat = tc::umma_desc_sw128(wbase + k * 32, 16, 1024);
                  tc::umma_tf32_ts(d_tmem, acol + k * 8, w_cat, idesc, (kk | h | k) ? 1u : 0u);    // a_hi * [w_hi | w_lo]
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t w_cat = tc::umma_desc_sw128(wbase + k * 32, 16, 1024);
                  tc::umma_tf32_ts(d_tmem, acol + 32 + k * 8, w_cat, idesc_lo, 1u);                  // a_lo * w_hi
                }
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t w_hi = tc::umma_desc_sw128(wbase + k * 32, 16, 1024);
                  const uint64_t w_lo = tc::umma_desc_sw128(wbase + NB * 128 + k * 32, 16, 1024);
                  tc::umma_tf32_ts(d_tmem, acol + 32 + k * 8, w_hi, idesc, (kk | h | k) ? 1u : 0u);  // a_lo * w_hi (small first)
                  tc::umma_tf32_ts(d_tmem, acol + k * 8, w_lo, idesc, 1u);                            // a_hi * w_lo
                  tc::umma_tf32_ts(d_tmem, acol + k * 8, w_hi, idesc, 1u);                            // a_hi * w_hi
                }
              }
              tc::umma_commit(&bar_aempty[sa]);
              if (h % kHS == kHS - 1) tc::umma_commit(&bar_wempty[sw]);
              if (h == 1 && kk == nk - 1) tc::umma_commit(&bar_accfull[ab]);
            }
            __syncwarp();
          }
        }
      }
    }
    if (!ok) s_err = 1;
  } else if (warp >= 4 && warp < 12) {
    // ------------------------------------------------------------------ stagers (256 threads): 32 channels per thread
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(kRegsStager));
    const int ct = tid - 128;
    const int set = ct >> 7;                          // channel half handled by this thread
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    long vt = 0;
    long wi = 0;
    bool ok = true;
    for (long work = blockIdx.x; work < nwork && ok; work += gridDim.x) {
      for (int g = 0; g < G && ok; ++g, ++vt) {
        const int hb = (int)(vt & 1), hp = (int)((vt >> 1) & 1);
        ok = tc::mbar_wait(&bar_hfull[hb], hp);
        if (!ok) break;
        const unsigned mask = tm.of(g, (int)(work % nblk));
        for (int tap = 0; tap < ntaps && ok; ++tap) {
          if (!((mask >> tap) & 1)) continue;
          const int sa = (int)((wi & 1) << 1) | set, pa = (int)((wi >> 1) & 1);
          ++wi;
          const int row = ntaps == 9 ? r + (tap / 3) * Wp + (tap % 3) : r;
          const unsigned char* src = s_halo + (hb * 2 + set) * half_bytes + row * 128;
          uint32_t hi[32], lo[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 v = *reinterpret_cast<const float4*>(src + ((j ^ (row & 7)) << 4));
            split_tf32(v.x, hi[4 * j + 0], lo[4 * j + 0]);
            split_tf32(v.y, hi[4 * j + 1], lo[4 * j + 1]);
            split_tf32(v.z, hi[4 * j + 2], lo[4 * j + 2]);
            split_tf32(v.w, hi[4 * j + 3], lo[4 * j + 3]);
          }
          ok = tc::mbar_wait(&bar_aempty[sa], pa ^ 1);
          if (!ok) break;
          tc::tcgen05_fence_after();
          const uint32_t dst = a_tmem + sa * 64 + lane_base;
          tc::tmem_st16(dst, hi);
          tc::tmem_st16(dst + 16, hi + 16);
          tc::tmem_st16(dst + 32, lo);
          tc::tmem_st16(dst + 48, lo + 16);
          tc::tmem_st_wait();
          tc::tcgen05_fence_before();
          tc::mbar_arrive(&bar_afull[sa]);
        }
        if (ok) tc::mbar_arrive(&bar_hempty[hb]);     // this thread no longer reads the halo buffer
      }
    }
    if (!ok) s_err = 1;
  } else if (warp < 4) {
    // ------------------------------------------------------------------ epilogue (128 threads)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(kRegsEpilogue));
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    int t = 0;
    bool ok = true;
    for (long work = blockIdx.x; work < nwork && ok; work += gridDim.x) {
      const long tile = work / nblk;
      const int cb = (int)(work % nblk);
      const int img = (int)(tile / tiles_per_img), tix = (int)(tile % tiles_per_img);
      const long img_base = flat ? 0 : (long)img * Hp * Wp;
      const int q0 = flat ? tix * kRows : (Wp + 1) + tix * kRows;
      const int q = q0 + r;
      const int hp = q / Wp, wp = q - hp * Wp;
      // flat (1x1 over a dense [rows][C] tensor: H = number of rows, B = 1): every row below H is an output
      const bool valid = flat ? q < H : (q < Hp * Wp && hp >= 1 && hp <= H && wp >= 1 && wp <= W);
      float run[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) run[j] = 0.f;
      for (int g0 = 0; g0 < G && ok; g0 += gchunk, ++t) {       // one accumulator per K chunk, summed here in fp32
        const int ab = t & 1, ap = (t >> 1) & 1;
        ok = tc::mbar_wait(&bar_accfull[ab], ap);
        if (!ok) break;
        tc::tcgen05_fence_after();
#pragma unroll
        for (int c = 0; c < NB; c += 16) {
          uint32_t v[16];
          tc::tmem_ld16(tmem + ab * 128 + lane_base + c, v);
          if constexpr (NB == 64) {                     // main and small-term accumulators are separate column ranges
            uint32_t u[16];
            tc::tmem_ld16(tmem + ab * 128 + 64 + lane_base + c, u);
            tc::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
          } else {
            tc::tmem_ld_wait();
          }
          if (c == NB - 16) {                           // accumulator fully read: hand the buffer back to the MMA warp
            tc::tcgen05_fence_before();
            tc::mbar_arrive(&bar_accempty[ab]);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) run[c + j] += __uint_as_float(v[j]);
        }
      }
      if (!ok) break;
      if (valid) {
        float* orow = out + (img_base + q) * (long)Cout + cb * NB;
        const float* brow = bias ? bias + cb * NB : nullptr;
#pragma unroll
        for (int j = 0; j < NB; j += 4) {
          float4 o = make_float4(run[j], run[j + 1], run[j + 2], run[j + 3]);
          if (brow) { o.x += brow[j]; o.y += brow[j + 1]; o.z += brow[j + 2]; o.w += brow[j + 3]; }
          dktb_st4(orow + j, o);
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
// Weight gradient on tcgen05 for the same layers:  dW[co][ci][tap] = sum_q X[q + off(tap)][ci] * G[q][co]  over all rows
// (3x3: padded-flat rows, both operands zero-bordered, images concatenate seamlessly; 1x1: dense rows).  The 64 x 64
// scheme of csrc/conv_tc.cu per (input group g, output block cb) pair = blockIdx.y: GEMM M = 128 = 2 taps x 64 ci, N = 64
// co, K = rows; A = X^T gathered into TMEM by the stagers (3xTF32 split in registers), B = G re-laid-out K-major in shared
// memory; accumulators of all tap pairs stay in TMEM; every CTA writes its partial dW once, dktb_wgrad_tcg's second
// kernel reduces them in a fixed order (deterministic).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kKR = 32;                         // rows per K-block
constexpr int kWgAStages = 3;
constexpr int kWgThreads = 64 + 256;
constexpr int kWgPStride = 9 * 64 * 64 + 64;    // one partial: [tap][ci][co] + bias row

template <int A>
__device__ __forceinline__ void wg_gather16(const unsigned char* __restrict__ xb, const int (&off)[8], uint32_t (&hi)[16],
                                            uint32_t (&lo)[16]) {
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const float v = *reinterpret_cast<const float*>(xb + k * 128 + off[(A + k) & 7]);
    split_tf32(v, hi[k], lo[k]);
  }
}

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_tcg_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_g,
                      float* __restrict__ partial, long total_rows, int Wp, int halo_pad, int ntaps, int nblk, int np, int Cs2,
                      int* __restrict__ err) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_raw_full[2], bar_raw_empty[2], bar_bfull[2], bar_afull[kWgAStages], bar_aempty[kWgAStages],
      bar_acc;
  __shared__ uint32_t s_tmem;
  __shared__ int s_err;
  __shared__ float s_bias[4][64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (*reinterpret_cast<volatile int*>(err) != 0) return;
  const int pair = blockIdx.y, g_in = pair / nblk, cb = pair % nblk;
  // np = 1: an ordinary layer, slot = tap.  np = 4: the space-to-depth input of a stride-2 3x3 layer (Cs2 channels per
  // parity plane): this CTA's 64 input channels are loaded from all four planes and the 9 slots are the (plane, tap) pairs
  // that carry weights -- plane 0: tap 4; plane 1: taps 3, 4; plane 2: taps 1, 4; plane 3: taps 0, 1, 3, 4.
  constexpr unsigned kS2Plane = 0u | (1u << 2) | (1u << 4) | (2u << 6) | (2u << 8) | (3u << 10) | (3u << 12) | (3u << 14) | (3u << 16);
  constexpr unsigned long long kS2Tap = 4ull | (3ull << 4) | (4ull << 8) | (1ull << 12) | (4ull << 16) | (0ull << 20) |
                                        (1ull << 24) | (3ull << 28) | (4ull << 32);
  const int ngroups = (ntaps + 1) / 2;                      // slot pairs
  const int lead = ntaps == 9 ? Wp + 1 : 0;
  const int x_half_bytes = halo_pad * 128;
  const int xr_bytes = np * 2 * x_half_bytes;                // X raw: [plane][channel half][halo row][128 B]
  const int stage_bytes = xr_bytes + 8192 + 16384;           // X raw | G raw | B hi | B lo
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
        tc::mbar_expect_tx(&bar_raw_full[s], xr_bytes + 8192);
        const long q0 = kb * kKR;
        const int xrow0 = (int)(q0 - lead);
        for (int h = 0; h < 2; ++h) {
          for (int pl = 0; pl < np; ++pl)
            for (int r = 0; r < halo_pad; r += kHaloBox)
              tc::tma_load_2d(st + (pl * 2 + h) * x_half_bytes + r * 128, &map_x, &bar_raw_full[s],
                              pl * Cs2 + g_in * 64 + h * 32, xrow0 + r);
          tc::tma_load_2d(st + xr_bytes + h * 4096, &map_g, &bar_raw_full[s], cb * 64 + h * 32, (int)q0);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t idesc = tc::umma_idesc(2, 128, 64, 0, 0);
    bool ok = true;
    int n = 0, ai = 0;
    for (long kb = blockIdx.x; kb < nkb && ok; kb += gridDim.x, ++n) {
      const int s = n & 1, ph = (n >> 1) & 1;
      ok = tc::mbar_wait(&bar_bfull[s], ph);
      if (!ok) break;
      const uint32_t bbase = tc::smem_u32(smem + s * stage_bytes + xr_bytes + 8192);
      for (int g = 0; g < ngroups && ok; ++g, ++ai) {
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
  } else {
    const int ct = tid - 64;                                  // 0..255
    const int set = ct >> 7;                                  // k range [set*16, +16) of the 32-row K-block
    const int quarter = warp & 3;
    const int L = quarter * 32 + lane;                        // TMEM lane: (tap-in-pair, ci)
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int ci = L & 63, tsel = L >> 6;
    const int co = ct & 63, kq8 = ct >> 6;                    // G transpose: 8 k values [kq8*8, +8) of channel co
    int xoff[8];
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
      {
        const unsigned char* graw = st + xr_bytes + (co >> 5) * 4096;
        unsigned char* bhi = st + xr_bytes + 8192 + co * 128;
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
          const int kchunk = kq8 * 2 + kq;
          const int off = (kchunk ^ (co & 7)) << 4;
          *reinterpret_cast<uint4*>(bhi + off) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(blo + off) = make_uint4(l[0], l[1], l[2], l[3]);
        }
        tc::fence_proxy_async_smem();
        tc::mbar_arrive(&bar_bfull[s]);
      }
      for (int g = 0; g < ngroups && ok; ++g, ++ai) {
        const int sa = ai % kWgAStages, pa = (ai / kWgAStages) & 1;
        const int slot = 2 * g + tsel;
        uint32_t hi[16], lo[16];
        if (slot < ntaps) {
          const int tap = np == 1 ? slot : (int)((kS2Tap >> (4 * slot)) & 15);
          const int pl = np == 1 ? 0 : (int)((kS2Plane >> (2 * slot)) & 3);
          const int shift = (ntaps == 9 ? (tap / 3) * Wp + (tap % 3) : 0) + set * 16;
          const unsigned char* xb = st + (pl * 2 + (ci >> 5)) * x_half_bytes + shift * 128;
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
      if (ok) tc::mbar_arrive(&bar_raw_empty[s]);
    }
    if (!ok) s_err = 1;
    ok = ok && tc::mbar_wait(&bar_acc, 0);
    tc::tcgen05_fence_after();
    float* out = partial + ((long)pair * gridDim.x + blockIdx.x) * kWgPStride;
    if (ok) {
      for (int g = 0; g < ngroups; ++g) {
        const int tap = 2 * g + tsel;
#pragma unroll
        for (int cc = 0; cc < 32; cc += 16) {
          const int c = set * 32 + cc;
          uint32_t v[16];
          tc::tmem_ld16(tmem + g * 64 + lane_base + c, v);
          tc::tmem_ld_wait();
          if (tap < ntaps) {
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

// partial [pair = g*nblk + cb][split][tap][ci][co] (+ bias row) -> dw [Cout][Cin][taps], db [Cout] (from the g = 0 pairs)
__global__ void conv_wgrad_tcg_reduce_kernel(const float* __restrict__ partial, int nsplit, int nblk, int Cin, int ntaps,
                                             float* __restrict__ dw, float* __restrict__ db) {
  const int pair = blockIdx.y, g = pair / nblk, cb = pair % nblk;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int nw = ntaps * 64 * 64;
  if (i >= nw + 64) return;
  const int src = i < nw ? i : 9 * 64 * 64 + (i - nw);
  const float* p = partial + (long)pair * nsplit * kWgPStride + src;
  float t = 0.f;
  for (int s = 0; s < nsplit; ++s) t += p[(long)s * kWgPStride];
  if (i < nw) {
    const int co = i % 64, ci = (i / 64) % 64, tap = i / 4096;
    dw[((long)(cb * 64 + co) * Cin + g * 64 + ci) * ntaps + tap] = t;
  } else if (db != nullptr && g == 0) {
    db[cb * 64 + (i - nw)] = t;
  }
}

// w [Cout][Cin][R][R] (R = 3 or 1) -> wb_fwd [Cout/NBf][Cin/64][taps][hi NBf | lo NBf][64]  (row = output channel, col =
// input channel) and wb_dgrad [Cin/NBd][Cout/64][taps][hi NBd | lo NBd][64] (taps flipped, roles swapped): hi = tf32-rounded,
// lo = rounded remainder.  NBf / NBd = tcg_nb(output channels, input channels) of the respective direction.
__global__ void prep_weights_tcg_kernel(const float* __restrict__ w, float* __restrict__ wb_fwd,
                                        float* __restrict__ wb_dgrad, int Cout, int Cin, int ntaps) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)Cout * Cin * ntaps) return;
  const int tap = (int)(i % ntaps);
  const int ci = (int)((i / ntaps) % Cin), co = (int)(i / ((long)ntaps * Cin));
  const float v = w[i];
  const float hi = tc::to_tf32_rna(v), lo = tc::to_tf32_rna(v - hi);
  if (wb_fwd) {
    const int NB = tcg_nb(Cout, Cin), G = Cin / 64;
    const long base = ((((long)(co / NB) * G + ci / 64) * ntaps + tap) * (2 * NB) + (co % NB)) * 64 + (ci % 64);
    wb_fwd[base] = hi;
    wb_fwd[base + NB * 64] = lo;
  }
  if (wb_dgrad) {
    const int NB = tcg_nb(Cin, Cout), G = Cout / 64;
    const long base = ((((long)(ci / NB) * G + co / 64) * ntaps + (ntaps - 1 - tap)) * (2 * NB) + (ci % NB)) * 64 + (co % 64);
    wb_dgrad[base] = hi;
    wb_dgrad[base + NB * 64] = lo;
  }
}

// zero the 1-pixel border of a padded NHWC tensor [B][H+2][W+2][C] (after element-wise passes that ran over the whole
// flat tensor); one thread per (border pixel, 4 channels)
__global__ void zero_border_kernel(float* __restrict__ x, int B, int H, int W, int C) {
  const int Hp = H + 2, Wp = W + 2;
  const int nb = 2 * Wp + 2 * H;                      // border pixels per image
  const long total = (long)B * nb * (C / 4);
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c4 = (int)(i % (C / 4)) * 4;
  long t = i / (C / 4);
  const int k = (int)(t % nb);
  const int b = (int)(t / nb);
  int hp, wp;
  if (k < Wp) { hp = 0; wp = k; }
  else if (k < 2 * Wp) { hp = Hp - 1; wp = k - Wp; }
  else { const int j = k - 2 * Wp; hp = 1 + (j >> 1); wp = (j & 1) ? Wp - 1 : 0; }
  dktb_st4(x + (((long)b * Hp + hp) * Wp + wp) * C + c4, make_float4(0.f, 0.f, 0.f, 0.f));
}

// dir 0: padded [B][H+2][W+2][C] interior <- dense [B][H][W][C];  dir 1: dense <- padded interior.  C % 4 == 0
__global__ void pad_copy_kernel(float* __restrict__ dense, float* __restrict__ padded, int B, int H, int W, int C,
                                int dir) {
  const long total = (long)B * H * W * (C / 4);
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c4 = (int)(i % (C / 4)) * 4;
  long t = i / (C / 4);
  const int w = (int)(t % W);
  t /= W;
  const int h = (int)(t % H);
  const int b = (int)(t / H);
  const long po = (((long)b * (H + 2) + h + 1) * (W + 2) + w + 1) * C + c4;
  const long uo = (((long)b * H + h) * W + w) * C + c4;
  if (dir == 0) dktb_st4(padded + po, dktb_ld4(dense + uo));
  else dktb_st4(dense + uo, dktb_ld4(padded + po));
}

}  // namespace

DKTB_EXPORT int dktb_conv_tcg_ok(int Cin, int Cout, int R, int stride, int pad, int dil, int W) {
  if (Cin % 64 != 0 || Cout % 64 != 0 || stride != 1 || dil != 1) return 0;
  if (R == 1 && pad == 0) return 1;
  return R == 3 && pad == 1 && W + 3 <= 64;       // halo (128 + 2 (W + 3) rows) x 4 buffers + weight ring within 227 KB
}

DKTB_EXPORT long dktb_conv_tcg_weight_floats(int Cin, int Cout, int R) { return (long)Cin * Cout * R * R * 2; }

DKTB_EXPORT int dktb_prep_weights_tcg(const float* w, float* wb_fwd, float* wb_dgrad, int Cout, int Cin, int R,
                                      cudaStream_t stream) {
  DKTB_CHECK_ARG(w && Cin % 64 == 0 && Cout % 64 == 0 && (R == 1 || R == 3));
  const long n = (long)Cout * Cin * R * R;
  prep_weights_tcg_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(w, wb_fwd, wb_dgrad, Cout, Cin, R * R);
  return dktb_launch_status();
}

// Launch of conv_tcg_kernel<NB>: geometry as dktb_conv_tcg; tm / gchunk: the K loop's tap masks and chunk length.
static int conv_tcg_launch(const float* a, const float* wb, const float* bias, float* out, int* err, int B, int H, int W,
                           int Cin, int Cout, int R, int NB, int gchunk, const TapMasks& tm, cudaStream_t stream) {
  const int flat = R == 1;
  const int Hp = H + 2, Wp = W + 2;
  const long rows = flat ? (long)B * H * W : (long)B * Hp * Wp;
  DKTB_CHECK_ARG(rows < 2147483000L);
  const int ntaps = R * R;
  const int G = Cin / 64;
  const int halo = kRows + (ntaps == 9 ? 2 * (Wp + 1) : 0);
  const int halo_pad = (halo + kHaloBox - 1) / kHaloBox * kHaloBox;
  const int smem = 4 * halo_pad * 128 + kWStages * kWStageBytes + 1024;
  DKTB_CHECK_ARG(smem <= 227 * 1024);
  const int nb = Cout / NB;
  CUtensorMap map_a, map_w;
  if (tc_make_tmap_2d(&map_a, a, (uint64_t)Cin, (uint64_t)rows, 32, kHaloBox) != 0) return DKTB_BAD_ARG - 1;
  if (tc_make_tmap_2d(&map_w, wb, 64, (uint64_t)nb * G * ntaps * 2 * NB, 32, 2 * NB) != 0) return DKTB_BAD_ARG - 1;
  const long span = flat ? rows : (long)Hp * Wp - 2 * (Wp + 1);
  const int tiles_per_img = (int)((span + kRows - 1) / kRows);
  const int nimg = flat ? 1 : B;
  const long nwork = (long)nimg * tiles_per_img * nb;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(nwork < sms ? nwork : sms);
  // flat: the kernel sees one "image" of `rows` rows (H carries the row count for the validity test)
  if (NB == 128) {
    cudaFuncSetAttribute(conv_tcg_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    conv_tcg_kernel<128><<<grid, kThreads, smem, stream>>>(map_a, map_w, bias, out, nimg, flat ? (int)rows : H, W, Cout, G, nb,
                                                           ntaps, halo_pad, tiles_per_img, flat, gchunk, tm, err);
  } else {
    cudaFuncSetAttribute(conv_tcg_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    conv_tcg_kernel<64><<<grid, kThreads, smem, stream>>>(map_a, map_w, bias, out, nimg, flat ? (int)rows : H, W, Cout, G, nb,
                                                          ntaps, halo_pad, tiles_per_img, flat, gchunk, tm, err);
  }
  return dktb_launch_status();
}

// R = 3: convolution (stride 1, pad 1) over padded-flat NHWC tensors: a [B][H+2][W+2][Cin] (zero border) -> out
// [B][H+2][W+2][Cout] (interior written, border untouched).  R = 1: a plain GEMM over the rows of DENSE tensors:
// a [B*H*W][Cin] -> out [B*H*W][Cout] (no padding involved).  wb: the forward tensor of dktb_prep_weights_tcg (+ bias
// [Cout] or NULL), or -- dgrad -- the dgrad tensor with Cin / Cout swapped by the caller (a = dL/dout [.., Cout], out =
// dL/da [.., Cin]).  err: device int, zero-initialised by the caller, set to 1 when a pipeline wait timed out.
DKTB_EXPORT int dktb_conv_tcg(const float* a, const float* wb, const float* bias, float* out, int* err, int B, int H,
                              int W, int Cin, int Cout, int R, cudaStream_t stream) {
  DKTB_CHECK_ARG(a && wb && out && err && B > 0 && H > 0 && W > 0);
  DKTB_CHECK_ARG(dktb_conv_tcg_ok(Cin, Cout, R, 1, R == 3 ? 1 : 0, 1, W));
  TapMasks tm;
  tm.m = (1ull << (R * R)) - 1;
  tm.per_plane = 1 << 30;
  tm.by_cb = 0;
  // K per accumulator: 576 (3x3) / 128 (1x1); no measurable cost against 1152 / 1024
  return conv_tcg_launch(a, wb, bias, out, err, B, H, W, Cin, Cout, R, tcg_nb(Cout, Cin), R == 3 ? 1 : 2, tm, stream);
}

// ------------------------------------------------------------------------------------------------------------------
// Stride-2 3x3 convolutions (pad 1; reference backbone.py:150-152, 195-197: the first convolution of a down-sampling block)
// as stride-1 convolutions over the SPACE-TO-DEPTH input: xs [B][H/2+2][W/2+2][4 C] padded-flat, channel = plane * C + c,
// plane = (ih & 1) * 2 + (iw & 1), block position (ih >> 1, iw >> 1).  Output pixel (oh, ow) reads input row 2 oh - 1 + r:
// r = 1 -> even plane at block oh; r = 0 / 2 -> odd plane at block oh - 1 / oh; the same along the width.  In the 3x3 tap
// grid of the stride-1 kernel (dh, dw in {-1, 0, 1}) that is dh in {-1, 0} for odd planes and dh = 0 for even ones: 9 of
// the 36 (plane, tap) pairs carry weights, the TapMasks skip the rest.  dgrad: dxs [.., 4 C] from dy [.., Cout], taps
// dh in {0, +1} for odd output planes (r = 2 / 0), dh = 0 for even ones (r = 1), mask by the plane of the output block.
// ------------------------------------------------------------------------------------------------------------------
namespace {

__host__ __device__ inline int s2_nb_fwd(int C, int Cout) { return Cout % 128 == 0 ? 128 : 64; }
__host__ __device__ inline int s2_nb_dgrad(int C, int Cout) { return (C % 128 == 0 && Cout >= 128) ? 128 : 64; }

// w [Cout][C][3][3] -> wb_fwd [Cout/NBf][4C/64][9][hi NBf | lo NBf][64], wb_dgrad [4C/NBd][Cout/64][9][hi NBd | lo NBd][64]
// (only the slots of the 9 non-zero (plane, tap) pairs are written; the kernel never reads the others)
__global__ void prep_weights_tcg_s2_kernel(const float* __restrict__ w, float* __restrict__ wb_fwd,
                                           float* __restrict__ wb_dgrad, int Cout, int C) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)Cout * C * 9) return;
  const int s = (int)(i % 3), r = (int)((i / 3) % 3);
  const int c = (int)((i / 9) % C), co = (int)(i / (9L * C));
  const float v = w[i];
  const float hi = tc::to_tf32_rna(v), lo = tc::to_tf32_rna(v - hi);
  const int ph = r != 1, pw = s != 1;                   // parity plane the tap reads (forward) / writes (dgrad)
  const int plane = ph * 2 + pw;
  if (wb_fwd) {
    const int dh = r == 0 ? -1 : 0, dw = s == 0 ? -1 : 0;
    const int tap = 3 * (dh + 1) + (dw + 1);
    const int NB = s2_nb_fwd(C, Cout), G = 4 * C / 64;
    const int g = plane * (C / 64) + c / 64;
    const long base = ((((long)(co / NB) * G + g) * 9 + tap) * (2 * NB) + (co % NB)) * 64 + (c % 64);
    wb_fwd[base] = hi;
    wb_fwd[base + NB * 64] = lo;
  }
  if (wb_dgrad) {
    const int dh = r == 0 ? 1 : 0, dw = s == 0 ? 1 : 0;
    const int tap = 3 * (dh + 1) + (dw + 1);
    const int NB = s2_nb_dgrad(C, Cout), G = Cout / 64;
    const int n = plane * C + c;
    const long base = ((((long)(n / NB) * G + co / 64) * 9 + tap) * (2 * NB) + (n % NB)) * 64 + (co % 64);
    wb_dgrad[base] = hi;
    wb_dgrad[base + NB * 64] = lo;
  }
}

// dir 0: x [B][H][W][C] (dense, or the interior of a padded-flat buffer when x_pad) -> xs [B][H/2+2][W/2+2][4C] interior;
// dir 1: the reverse.  One float4 per thread.
__global__ void s2d_kernel(float* __restrict__ x, float* __restrict__ xs, int H, int W, int C, int x_pad, int dir,
                           unsigned per_img4) {
  const unsigned i4 = blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 >= per_img4) return;
  const int b = blockIdx.y;
  const unsigned c4n = (unsigned)C >> 2;
  const unsigned p = i4 / c4n;
  const int c = (int)(i4 - p * c4n) * 4;
  const int ih = (int)(p / (unsigned)W), iw = (int)(p - (unsigned)ih * W);
  const int Hs = H / 2 + 2, Ws = W / 2 + 2;
  const long xo = x_pad ? (((long)b * (H + 2) + ih + 1) * (W + 2) + iw + 1) * C + c : (((long)b * H + ih) * W + iw) * C + c;
  const long so = (((long)b * Hs + (ih >> 1) + 1) * Ws + (iw >> 1) + 1) * (4L * C) + ((ih & 1) * 2 + (iw & 1)) * C + c;
  if (dir == 0) dktb_st4(xs + so, dktb_ld4(x + xo));
  else dktb_st4(x + xo, dktb_ld4(xs + so));
}

// Stride-2 1x1 convolutions (the shortcut of a down-sampling block, backbone.py:158-160, 205-207) read every second pixel
// of every second row: dir 0 gathers them into a dense [B][H/2][W/2][C] tensor (then a plain GEMM, dktb_conv_tcg R = 1);
// dir 1 scatters a gradient back (the skipped pixels get zeros).  x: dense, or padded-flat when x_pad.
__global__ void subsample2_kernel(float* __restrict__ x, float* __restrict__ xg, int H, int W, int C, int x_pad, int dir,
                                  unsigned per_img4) {
  const unsigned i4 = blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 >= per_img4) return;
  const int b = blockIdx.y;
  const unsigned c4n = (unsigned)C >> 2;
  const unsigned p = i4 / c4n;
  const int c = (int)(i4 - p * c4n) * 4;
  const int ih = (int)(p / (unsigned)W), iw = (int)(p - (unsigned)ih * W);
  const bool even = !((ih | iw) & 1);
  const long xo = x_pad ? (((long)b * (H + 2) + ih + 1) * (W + 2) + iw + 1) * C + c : (((long)b * H + ih) * W + iw) * C + c;
  const long go = (((long)b * (H / 2) + (ih >> 1)) * (W / 2) + (iw >> 1)) * C + c;
  if (dir == 0) {
    if (even) dktb_st4(xg + go, dktb_ld4(x + xo));
  } else {
    dktb_st4(x + xo, even ? dktb_ld4(xg + go) : make_float4(0.f, 0.f, 0.f, 0.f));
  }
}

TapMasks s2_masks(int dgrad, int per_plane) {
  TapMasks tm;
  tm.m = 0;
  for (int plane = 0; plane < 4; ++plane) {
    const int ph = plane >> 1, pw = plane & 1;
    unsigned m = 0;
    for (int dh = -1; dh <= 1; ++dh)
      for (int dw = -1; dw <= 1; ++dw) {
        const bool vh = dh == 0 || (ph == 1 && dh == (dgrad ? 1 : -1));
        const bool vw = dw == 0 || (pw == 1 && dw == (dgrad ? 1 : -1));
        if (vh && vw) m |= 1u << (3 * (dh + 1) + (dw + 1));
      }
    tm.m |= (unsigned long long)m << (16 * plane);
  }
  tm.per_plane = per_plane;
  tm.by_cb = dgrad;
  return tm;
}

}  // namespace

DKTB_EXPORT int dktb_conv_tcg_s2_ok(int C, int Cout, int H, int W) {
  return C % 64 == 0 && Cout % 64 == 0 && H % 2 == 0 && W % 2 == 0 && H >= 2 && W >= 2 && W / 2 + 3 <= 64;
}
DKTB_EXPORT long dktb_conv_tcg_s2_weight_floats(int C, int Cout) { return 4L * C * Cout * 9 * 2; }

DKTB_EXPORT int dktb_prep_weights_tcg_s2(const float* w, float* wb_fwd, float* wb_dgrad, int Cout, int C,
                                         cudaStream_t stream) {
  DKTB_CHECK_ARG(w && C % 64 == 0 && Cout % 64 == 0);
  const long n = (long)Cout * C * 9;
  prep_weights_tcg_s2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(w, wb_fwd, wb_dgrad, Cout, C);
  return dktb_launch_status();
}

// dir 0: x [B][H][W][C] -> xs [B][H/2+2][W/2+2][4C] (interior; the border of xs must already be zero); dir 1: xs -> x.
// x_pad: x is the base of a padded-flat buffer [B][H+2][W+2][C] instead of a dense tensor.
DKTB_EXPORT int dktb_s2d(float* x, float* xs, int B, int H, int W, int C, int x_pad, int dir, cudaStream_t stream) {
  DKTB_CHECK_ARG(x && xs && B > 0 && B <= 65535 && H % 2 == 0 && W % 2 == 0 && C % 4 == 0 && (long)H * W * C / 4 < 2147483647L);
  const unsigned per_img4 = (unsigned)((long)H * W * C / 4);
  s2d_kernel<<<dim3((per_img4 + 255) / 256, B), 256, 0, stream>>>(x, xs, H, W, C, x_pad, dir, per_img4);
  return dktb_launch_status();
}

// dir 0: xg [B][H/2][W/2][C] <- x[:, ::2, ::2, :]; dir 1: x <- xg at the even pixels, zeros elsewhere.  x_pad as dktb_s2d.
DKTB_EXPORT int dktb_subsample2(float* x, float* xg, int B, int H, int W, int C, int x_pad, int dir, cudaStream_t stream) {
  DKTB_CHECK_ARG(x && xg && B > 0 && B <= 65535 && H % 2 == 0 && W % 2 == 0 && C % 4 == 0 && (long)H * W * C / 4 < 2147483647L);
  const unsigned per_img4 = (unsigned)((long)H * W * C / 4);
  subsample2_kernel<<<dim3((per_img4 + 255) / 256, B), 256, 0, stream>>>(x, xg, H, W, C, x_pad, dir, per_img4);
  return dktb_launch_status();
}

// Ho, Wo: OUTPUT size (input H = 2 Ho, W = 2 Wo).  dgrad = 0: a = xs [B][Ho+2][Wo+2][4C] -> out [B][Ho+2][Wo+2][Cout]
// (+ bias); dgrad = 1: a = dy [B][Ho+2][Wo+2][Cout] -> out = dxs [B][Ho+2][Wo+2][4C].  Both padded-flat, zero borders on
// the input, interior of the output written.  wb: the matching tensor of dktb_prep_weights_tcg_s2.
DKTB_EXPORT int dktb_conv_tcg_s2(const float* a, const float* wb, const float* bias, float* out, int* err, int B, int Ho,
                                 int Wo, int C, int Cout, int dgrad, cudaStream_t stream) {
  DKTB_CHECK_ARG(a && wb && out && err && B > 0 && Ho > 0 && Wo > 0);
  DKTB_CHECK_ARG(dktb_conv_tcg_s2_ok(C, Cout, 2 * Ho, 2 * Wo));
  if (!dgrad) {
    const int NB = s2_nb_fwd(C, Cout);
    // one parity plane per accumulator: K = C x {1, 2, 2, 4} taps
    return conv_tcg_launch(a, wb, bias, out, err, B, Ho, Wo, 4 * C, Cout, 3, NB, C / 64, s2_masks(0, C / 64), stream);
  }
  const int NB = s2_nb_dgrad(C, Cout);
  return conv_tcg_launch(a, wb, nullptr, out, err, B, Ho, Wo, Cout, 4 * C, 3, NB, 2, s2_masks(1, C / NB), stream);
}

// how many K-splits (CTAs per (group, block) pair) dktb_wgrad_tcg uses, and the scratch it needs (floats)
static int wgrad_tcg_nsplit(long rows, int npairs) {
  const long nkb = (rows + kKR - 1) / kKR;
  int per = 148 / npairs;
  if (per < 1) per = 1;
  if (per > 148) per = 148;
  return (int)(nkb < per ? nkb : per);
}
DKTB_EXPORT long dktb_wgrad_tcg_scratch_floats(int B, int H, int W, int Cin, int Cout, int R) {
  const long rows = R == 1 ? (long)B * H * W : (long)B * (H + 2) * (W + 2);
  const int npairs = (Cin / 64) * (Cout / 64);
  return (long)npairs * wgrad_tcg_nsplit(rows, npairs) * kWgPStride;
}

// Weight gradient of the dktb_conv_tcg layers: x [.., Cin], gy [.., Cout] in the layouts of dktb_conv_tcg (R = 3: both
// padded-flat with ZERO borders; R = 1: dense rows) -> dw [Cout][Cin][R][R] (overwritten), db [Cout] or NULL.
DKTB_EXPORT int dktb_wgrad_tcg(const float* x, const float* gy, float* dw, float* db, float* scratch, int* err, int B,
                               int H, int W, int Cin, int Cout, int R, cudaStream_t stream) {
  DKTB_CHECK_ARG(x && gy && dw && scratch && err && B > 0 && H > 0 && W > 0);
  DKTB_CHECK_ARG(dktb_conv_tcg_ok(Cin, Cout, R, 1, R == 3 ? 1 : 0, 1, W));
  const int flat = R == 1;
  const int Wp = W + 2;
  const long rows = flat ? (long)B * H * W : (long)B * (H + 2) * Wp;
  DKTB_CHECK_ARG(rows < 2147483000L);
  const int ntaps = R * R;
  const int G = Cin / 64, nblk = Cout / 64, npairs = G * nblk;
  DKTB_CHECK_ARG(npairs <= 65535);
  const int halo = kKR + (ntaps == 9 ? 2 * (Wp + 1) : 0);
  const int halo_pad = (halo + kHaloBox - 1) / kHaloBox * kHaloBox;
  const int stage = 2 * halo_pad * 128 + 8192 + 16384;
  const int smem = 2 * stage + 1024;
  DKTB_CHECK_ARG(smem <= 227 * 1024);
  CUtensorMap map_x, map_g;
  if (tc_make_tmap_2d(&map_x, x, (uint64_t)Cin, (uint64_t)rows, 32, kHaloBox) != 0) return DKTB_BAD_ARG - 1;
  if (tc_make_tmap_2d(&map_g, gy, (uint64_t)Cout, (uint64_t)rows, 32, kKR) != 0) return DKTB_BAD_ARG - 1;
  cudaFuncSetAttribute(conv_wgrad_tcg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int nsplit = wgrad_tcg_nsplit(rows, npairs);
  conv_wgrad_tcg_kernel<<<dim3(nsplit, npairs), kWgThreads, smem, stream>>>(map_x, map_g, scratch, rows, Wp, halo_pad, ntaps,
                                                                           nblk, 1, 0, err);
  int rc = dktb_launch_status();
  if (rc != 0) return rc;
  conv_wgrad_tcg_reduce_kernel<<<dim3((ntaps * 4096 + 64 + 255) / 256, npairs), 256, 0, stream>>>(scratch, nsplit, nblk, Cin,
                                                                                                 ntaps, dw, db);
  return dktb_launch_status();
}

// partial [pair = cg*nblk + cb][split][slot][ci][co] of a stride-2 layer (slot = the 9 (parity plane, tap) pairs of
// conv_wgrad_tcg_kernel) -> dw [Cout][C][3][3], db [Cout] (from the cg = 0 pairs)
__global__ void conv_wgrad_tcg_s2_reduce_kernel(const float* __restrict__ partial, int nsplit, int nblk, int C,
                                                float* __restrict__ dw, float* __restrict__ db) {
  const int pair = blockIdx.y, cg = pair / nblk, cb = pair % nblk;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int nw = 9 * 64 * 64;
  if (i >= nw + 64) return;
  if (i >= nw && (db == nullptr || cg != 0)) return;
  const float* p = partial + (long)pair * nsplit * kWgPStride + i;
  float t = 0.f;
  for (int s = 0; s < nsplit; ++s) t += p[(long)s * kWgPStride];
  if (i < nw) {
    const int co = i % 64, ci = (i / 64) % 64, slot = i / 4096;
    const int planes[9] = {0, 1, 1, 2, 2, 3, 3, 3, 3}, taps[9] = {4, 3, 4, 1, 4, 0, 1, 3, 4};
    const int ph = planes[slot] >> 1, pw = planes[slot] & 1, dh = taps[slot] / 3 - 1, dv = taps[slot] % 3 - 1;
    const int r = ph == 0 ? 1 : (dh == -1 ? 0 : 2), sx = pw == 0 ? 1 : (dv == -1 ? 0 : 2);
    dw[(((long)(cb * 64 + co) * C + cg * 64 + ci) * 3 + r) * 3 + sx] = t;
  } else {
    db[cb * 64 + (i - nw)] = t;
  }
}

// Weight gradient of a stride-2 3x3 layer from its space-to-depth input: xs [B][Ho+2][Wo+2][4C], gy [B][Ho+2][Wo+2][Cout]
// (both padded-flat, zero borders) -> dw [Cout][C][3][3] (overwritten), db [Cout] or NULL.  One CTA column per (64 input
// channels, 64 output channels) pair: the four parity planes of those channels are staged together, so a K-block feeds
// the same five tap-pair accumulators as an ordinary 3x3 layer.
DKTB_EXPORT int dktb_wgrad_tcg_s2_ok(int C, int Cout, int H, int W) {
  if (!dktb_conv_tcg_s2_ok(C, Cout, H, W)) return 0;
  const int halo_pad = (kKR + (W / 2 + 3) + kHaloBox - 1) / kHaloBox * kHaloBox;
  return 2 * (8 * halo_pad * 128 + 8192 + 16384) + 1024 <= 227 * 1024;
}
DKTB_EXPORT long dktb_wgrad_tcg_s2_scratch_floats(int B, int Ho, int Wo, int C, int Cout) {
  const long rows = (long)B * (Ho + 2) * (Wo + 2);
  const int npairs = (C / 64) * (Cout / 64);
  return (long)npairs * wgrad_tcg_nsplit(rows, npairs) * kWgPStride;
}
DKTB_EXPORT int dktb_wgrad_tcg_s2(const float* xs, const float* gy, float* dw, float* db, float* scratch, int* err, int B,
                                  int Ho, int Wo, int C, int Cout, cudaStream_t stream) {
  DKTB_CHECK_ARG(xs && gy && dw && scratch && err && B > 0 && Ho > 0 && Wo > 0);
  DKTB_CHECK_ARG(dktb_wgrad_tcg_s2_ok(C, Cout, 2 * Ho, 2 * Wo));
  const int Wp = Wo + 2;
  const long rows = (long)B * (Ho + 2) * Wp;
  DKTB_CHECK_ARG(rows < 2147483000L);
  const int nblk = Cout / 64, npairs = (C / 64) * nblk;
  DKTB_CHECK_ARG(npairs <= 65535);
  const int halo_pad = (kKR + (Wp + 1) + kHaloBox - 1) / kHaloBox * kHaloBox;      // taps reach back (dh, dw in {-1, 0}) only
  const int smem = 2 * (8 * halo_pad * 128 + 8192 + 16384) + 1024;
  CUtensorMap map_x, map_g;
  if (tc_make_tmap_2d(&map_x, xs, (uint64_t)(4 * C), (uint64_t)rows, 32, kHaloBox) != 0) return DKTB_BAD_ARG - 1;
  if (tc_make_tmap_2d(&map_g, gy, (uint64_t)Cout, (uint64_t)rows, 32, kKR) != 0) return DKTB_BAD_ARG - 1;
  cudaFuncSetAttribute(conv_wgrad_tcg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int nsplit = wgrad_tcg_nsplit(rows, npairs);
  conv_wgrad_tcg_kernel<<<dim3(nsplit, npairs), kWgThreads, smem, stream>>>(map_x, map_g, scratch, rows, Wp, halo_pad, 9, nblk,
                                                                           4, C, err);
  int rc = dktb_launch_status();
  if (rc != 0) return rc;
  conv_wgrad_tcg_s2_reduce_kernel<<<dim3((9 * 4096 + 64 + 255) / 256, npairs), 256, 0, stream>>>(scratch, nsplit, nblk, C, dw,
                                                                                                db);
  return dktb_launch_status();
}

DKTB_EXPORT int dktb_zero_border(float* x, int B, int H, int W, int C, cudaStream_t stream) {
  DKTB_CHECK_ARG(x && B > 0 && H > 0 && W > 0 && C % 4 == 0);
  const long total = (long)B * (2 * (W + 2) + 2 * H) * (C / 4);
  zero_border_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(x, B, H, W, C);
  return dktb_launch_status();
}

// dir 0: padded [B][H+2][W+2][C] interior <- dense [B][H][W][C] (the border of `padded` is not touched: allocate it
// zeroed);  dir 1: dense <- padded interior
DKTB_EXPORT int dktb_pad_copy(float* dense, float* padded, int B, int H, int W, int C, int dir, cudaStream_t stream) {
  DKTB_CHECK_ARG(dense && padded && B > 0 && H > 0 && W > 0 && C % 4 == 0 && (dir == 0 || dir == 1));
  const long total = (long)B * H * W * (C / 4);
  pad_copy_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(dense, padded, B, H, W, C, dir);
  return dktb_launch_status();
}

#endif  // DKTB_EMU
