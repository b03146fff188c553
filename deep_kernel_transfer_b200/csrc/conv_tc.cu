// tcgen05 (5th-gen tensor core) 3x3 64->64 convolution over the padded-flat NHWC layout, forward and dgrad,
// with fp32-class accuracy from an error-compensated 3xTF32 split:
//     a = a_hi + a_lo  (a_hi = rna_tf32(a), a_lo = a - a_hi exact),  w likewise (split once per step by a prep kernel)
//     D += a_hi*w_hi + a_lo*w_hi + a_hi*w_lo        (fp32 accumulation in TMEM; dropped term ~2^-22)
// GEMM view per CTA: M = 128 consecutive padded-flat pixels of one image, N = 64 output channels,
// K = 9 taps x 64 input channels; tap (r,s) is the same activation matrix shifted by (r-1)*Wp + (s-1) rows, so
// every operand tile is a plain 2-D TMA box and the zero border of the layout supplies the padding.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer,
// warps 2-5 = in-SM splitter (raw fp32 tile -> hi in place + lo copy) and epilogue (TMEM -> regs -> smem ->
// coalesced global stores + per-tile BatchNorm partial sums).  Pipeline: full[s] (TMA landed) -> conv[s]
// (split done) -> MMA -> empty[s] (tcgen05.commit), STAGES-deep ring over the 18 (tap, channel-half) k-blocks.
#include "dktb_common.cuh"

#ifndef DKTB_EMU
#include "tc_common.cuh"

namespace {

constexpr int kRows = 128;
constexpr int kStages = 2;
constexpr int kStageBytes = 49152;            // A(hi) 16K | A lo 16K | W hi 8K | W lo 8K
constexpr int kOffALo = 16384, kOffWHi = 32768, kOffWLo = 40960;
constexpr int kIters = 18;                    // 9 taps x 2 channel halves
constexpr int kOutLd = 65;                    // epilogue staging row stride (floats)
constexpr int kSmemBytes = kStages * kStageBytes + 1024;

struct TcErr { int flag; };

__global__ void __launch_bounds__(192, 2)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                  const float* __restrict__ bias, float* __restrict__ out, float* __restrict__ partials, int B, int H,
                  int W, int* __restrict__ err) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // the dynamic smem base is only guaranteed 16 B aligned by the ABI: align by hand for SWIZZLE_128B
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_full[kStages], bar_conv[kStages], bar_empty[kStages], bar_acc;
  __shared__ uint32_t s_tmem;
  __shared__ int s_err;
  __shared__ float s_valid[kRows];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Hp = H + 2, Wp = W + 2;
  const int img = blockIdx.y;
  const int q0 = (Wp + 1) + blockIdx.x * kRows;
  const long img_base = (long)img * Hp * Wp;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      tc::mbar_init(&bar_full[s], 1);
      tc::mbar_init(&bar_conv[s], 128);
      tc::mbar_init(&bar_empty[s], 1);
    }
    tc::mbar_init(&bar_acc, 1);
    s_err = 0;
    tc::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&map_a);
    tc::prefetch_tmap(&map_w);
  }
  if (warp == 1) tc::tmem_alloc<64>(&s_tmem);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t d_tmem = s_tmem;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int it = 0; it < kIters; ++it) {
        const int s = it % kStages, ph = (it / kStages) & 1;
        const int tap = it >> 1, half = it & 1;
        if (!tc::mbar_wait(&bar_empty[s], ph ^ 1)) { s_err = 1; break; }
        unsigned char* st = smem + s * kStageBytes;
        tc::mbar_expect_tx(&bar_full[s], 16384 + 8192 + 8192);
        const int row = (int)(img_base + q0 - (Wp + 1) + (tap / 3) * Wp + (tap % 3));
        tc::tma_load_2d(st, &map_a, &bar_full[s], half * 32, row);
        tc::tma_load_2d(st + kOffWHi, &map_w, &bar_full[s], half * 32, tap * 64);
        tc::tma_load_2d(st + kOffWLo, &map_w, &bar_full[s], half * 32, (9 + tap) * 64);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = tc::umma_idesc(2, 128, 64, 0, 0);
      bool ok = true;
      for (int it = 0; it < kIters && ok; ++it) {
        const int s = it % kStages, ph = (it / kStages) & 1;
        ok = tc::mbar_wait(&bar_conv[s], ph);
        if (!ok) break;
        tc::tcgen05_fence_after();
        const uint32_t base = tc::smem_u32(smem + s * kStageBytes);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t a_hi = tc::umma_desc_sw128(base + k * 32, 16, 1024);
          const uint64_t a_lo = tc::umma_desc_sw128(base + kOffALo + k * 32, 16, 1024);
          const uint64_t w_hi = tc::umma_desc_sw128(base + kOffWHi + k * 32, 16, 1024);
          const uint64_t w_lo = tc::umma_desc_sw128(base + kOffWLo + k * 32, 16, 1024);
          tc::umma_tf32_ss(d_tmem, a_lo, w_hi, idesc, (it | k) ? 1u : 0u);   // small terms first
          tc::umma_tf32_ss(d_tmem, a_hi, w_lo, idesc, 1u);
          tc::umma_tf32_ss(d_tmem, a_hi, w_hi, idesc, 1u);
        }
        tc::umma_commit(&bar_empty[s]);
      }
      if (!ok) s_err = 1;
      tc::umma_commit(&bar_acc);
    }
  } else {
    // ------------------------------------------------------------------ splitter warps (128 threads)
    const int ct = tid - 64;
    bool ok = true;
    for (int it = 0; it < kIters && ok; ++it) {
      const int s = it % kStages, ph = (it / kStages) & 1;
      ok = tc::mbar_wait(&bar_full[s], ph);
      if (!ok) break;
      unsigned char* st = smem + s * kStageBytes;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4* pa = reinterpret_cast<float4*>(st + j * 2048 + ct * 16);
        float4* pl = reinterpret_cast<float4*>(st + kOffALo + j * 2048 + ct * 16);
        const float4 v = *pa;
        float4 hi, lo;
        hi.x = tc::to_tf32_rna(v.x); hi.y = tc::to_tf32_rna(v.y); hi.z = tc::to_tf32_rna(v.z); hi.w = tc::to_tf32_rna(v.w);
        lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
        *pa = hi;
        *pl = lo;
      }
      tc::fence_proxy_async_smem();
      tc::mbar_arrive(&bar_conv[s]);
    }
    if (!ok) s_err = 1;
    // ------------------------------------------------------------------ epilogue
    ok = ok && tc::mbar_wait(&bar_acc, 0);
    tc::tcgen05_fence_after();
    float* s_out = reinterpret_cast<float*>(smem);          // [128][kOutLd], stage memory is free now
    const int quarter = warp & 3;                             // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;                        // accumulator row = padded-flat pixel q0 + r
    const int q = q0 + r;
    const int hp = q / Wp, wp = q - hp * Wp;
    const bool valid = ok && q < Hp * Wp && hp >= 1 && hp <= H && wp >= 1 && wp <= W;
    s_valid[r] = valid ? 1.f : 0.f;
    if (ok) {
#pragma unroll
      for (int c = 0; c < 64; c += 16) {
        uint32_t v[16];
        tc::tmem_ld16(d_tmem + ((uint32_t)(quarter * 32) << 16) + c, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float b = bias ? bias[c + j] : 0.f;
          s_out[r * kOutLd + c + j] = __uint_as_float(v[j]) + b;
        }
      }
    }
    tc::tcgen05_fence_before();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    // coalesced stores: 16 lanes cover one pixel's 64 channels
    for (int idx = ct; idx < kRows * 16; idx += 128) {
      const int rr = idx >> 4, c4 = (idx & 15) * 4;
      if (s_valid[rr] != 0.f) {
        const float* src = s_out + rr * kOutLd + c4;
        dktb_st4(out + (img_base + q0 + rr) * 64 + c4, make_float4(src[0], src[1], src[2], src[3]));
      }
    }
    if (partials != nullptr) {
      const int which = ct >> 6, c = ct & 63;
      float t = 0.f;
      for (int rr = 0; rr < kRows; ++rr) {
        const float v = s_out[rr * kOutLd + c] * s_valid[rr];
        t += which ? v * v : v;
      }
      const long blk = (long)img * gridDim.x + blockIdx.x;
      partials[(blk * 2 + which) * 64 + c] = t;
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (tid == 0 && s_err) atomicExch(err, 1);
  if (warp == 1) tc::tmem_dealloc<64>(d_tmem);
}

// w_ref [co][ci][3][3] -> wb_fwd / wb_dgrad [hl][tap][n][k] (hi = rna_tf32, lo = exact remainder)
__global__ void prep_weights_tc_kernel(const float* __restrict__ w, float* __restrict__ wb_fwd,
                                       float* __restrict__ wb_dgrad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 64 * 9) return;
  const int tap = i % 9, ci = (i / 9) % 64, co = i / (9 * 64);
  const float v = w[i];
  const float hi = tc::to_tf32_rna(v), lo = v - hi;
  if (wb_fwd) {
    wb_fwd[((0 * 9 + tap) * 64 + co) * 64 + ci] = hi;
    wb_fwd[((1 * 9 + tap) * 64 + co) * 64 + ci] = lo;
  }
  if (wb_dgrad) {
    wb_dgrad[((0 * 9 + (8 - tap)) * 64 + ci) * 64 + co] = hi;
    wb_dgrad[((1 * 9 + (8 - tap)) * 64 + ci) * 64 + co] = lo;
  }
}

}  // namespace

DKTB_EXPORT int dktb_prep_weights_tc(const float* w, float* wb_fwd, float* wb_dgrad, cudaStream_t stream) {
  DKTB_CHECK_ARG(w != nullptr);
  prep_weights_tc_kernel<<<(64 * 64 * 9 + 255) / 256, 256, 0, stream>>>(w, wb_fwd, wb_dgrad);
  return dktb_launch_status();
}

// Same contract as dktb_conv3x3_fwd, with wb = the [2][9][64][64] hi/lo weight tensor of dktb_prep_weights_tc.
// err: device int, set to 1 if a pipeline wait timed out (must be zero-initialised by the caller).
DKTB_EXPORT int dktb_conv3x3_tc_fwd(const float* a, const float* wb, const float* bias, float* out, float* partials,
                                    int* err, int B, int H, int W, cudaStream_t stream) {
  DKTB_CHECK_ARG(a && wb && out && err && B > 0 && H > 0 && W > 0 && B <= 65535);
  const int Hp = H + 2, Wp = W + 2;
  const long rows = (long)B * Hp * Wp;
  DKTB_CHECK_ARG(rows < 2147483000L);
  CUtensorMap map_a, map_w;
  if (tc_make_tmap_2d(&map_a, a, 64, (uint64_t)rows, 32, kRows) != 0) return DKTB_BAD_ARG - 1;
  if (tc_make_tmap_2d(&map_w, wb, 64, 2 * 9 * 64, 32, 64) != 0) return DKTB_BAD_ARG - 1;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    attr_done = true;
  }
  const int span = Hp * Wp - 2 * (Wp + 1);
  dim3 grid((span + kRows - 1) / kRows, B);
  conv3x3_tc_kernel<<<grid, 192, kSmemBytes, stream>>>(map_a, map_w, bias, out, partials, B, H, W, err);
  return dktb_launch_status();
}

#endif  // DKTB_EMU
