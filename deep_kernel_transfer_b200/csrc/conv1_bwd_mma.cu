// Fused first-block backward, tensor-core variant (same contract as dktb_conv1_bwd_fused in conv_fp32.cu; reference:
// loss.backward() through trunk[0] = Conv2d(3,64,3,pad 1) -> BatchNorm2d -> ReLU -> MaxPool2d(2), backbone.py:93-102).
// Per 4 x 32 pixel tile the BatchNorm / ReLU / MaxPool backward of the pooled gradient is evaluated into shared memory
// exactly as in the FFMA version; the weight-gradient GEMM  dW[tap][co] += sum_pix P[pix][tap] * G[pix][co]
// (M = 32 taps incl. padding, N = 64, K = 128 pixels) then runs on warp-level mma.sync.m16n8k8 tf32 with the 3xTF32
// error-compensated split (a_hi b_hi + a_lo b_hi + a_hi b_lo), ~9x fewer issue slots than the FFMA loop.  Padding tap 27
// carries the constant 1 so that row 27 of the product is the bias gradient.  Each warp owns 16 pixels of the tile
// (two K = 8 steps) and the full 32 x 64 accumulator; the 8 warps are summed once at the end in a fixed order.
#include "dktb_common.cuh"

#define CM_TH 4
#define CM_TW 32
#define CM_GLD 72        // gradient tile pitch: B-fragment loads (4 rows x 8 channels per warp) hit 32 distinct banks

__device__ __forceinline__ void cm_split(float v, unsigned& hi, unsigned& lo) {
  hi = (__float_as_uint(v) + 0x1000u) & 0xFFFFE000u;
  lo = __float_as_uint(v - __uint_as_float(hi));          // exact remainder; the tensor core truncates it to tf32
}

__global__ void __launch_bounds__(256, 2) conv1_bwd_mma_kernel(
    const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ gout,
    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
    const float* __restrict__ beta, const float* __restrict__ sums, float* __restrict__ partial, int B, int H, int W,
    int ipe, int out_pad, float inv_count) {
  __shared__ __align__(16) float s_in[3 * (CM_TH + 2) * (CM_TW + 4)];
  __shared__ __align__(16) float s_g[CM_TH * CM_TW * CM_GLD];
  __shared__ __align__(16) float s_k[5][64];       // mean, gamma*invstd, beta, A = -ss*s1/n, Bc = -ss*s2/n*invstd
  constexpr int IW = CM_TW + 4, IPLANE = (CM_TH + 2) * IW;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int TX = (W + CM_TW - 1) / CM_TW, TY = (H + CM_TH - 1) / CM_TH;
  const long ntiles = (long)B * TX * TY;
  const int Ho = H / 2, Wo = W / 2, Hq = Ho + 2 * out_pad, Wq = Wo + 2 * out_pad;
  const int c4 = (tid % 16) * 4;
  // MMA roles: this warp's 16 pixels of the tile and this lane's 4 tap rows (two m-tiles x {g, g+8})
  const int py = warp >> 1, pxb = (warp & 1) * 16;
  int toff[2][2];
  bool tval[2][2], tone[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int tap = 16 * mt + g + 8 * hh;
      tval[mt][hh] = tap < 27;
      tone[mt][hh] = tap == 27;
      const int tp = tap < 27 ? tap : 0;
      toff[mt][hh] = (tp / 9) * IPLANE + ((tp % 9) / 3) * IW + (tp % 3) + py * IW + pxb + t;
    }
  float acc[2][8][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
  int e_cur = -1;
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = (int)(tile / (TX * TY));
    const int rem = (int)(tile % (TX * TY));
    const int h0 = (rem / TX) * CM_TH, w0 = (rem % TX) * CM_TW;
    const int e = b / ipe;
    if (e != e_cur) {                                       // uniform over the CTA
      if (tid < 64) {
        const float is = invstd[e * 64 + tid], ss = gamma[tid] * is;
        s_k[0][tid] = mean[e * 64 + tid];
        s_k[1][tid] = ss;
        s_k[2][tid] = beta[tid];
        s_k[3][tid] = -ss * (sums[(long)e * 128 + tid] * inv_count);
        s_k[4][tid] = -ss * (sums[(long)e * 128 + 64 + tid] * inv_count) * is;
      }
      e_cur = e;
    }
    __syncthreads();
    // input patch rows h0-1 .. h0+4, columns w0-1 .. w0+32 (zero outside the image): loads now, stores after the
    // gradient-tile loads have been issued as well
    float xv[3];
    int xdst[3];
#pragma unroll
    for (int it = 0; it < 3; ++it) {
      const int i = tid + 256 * it;
      xv[it] = 0.f;
      xdst[it] = -1;
      if (i < 3 * (CM_TH + 2) * (CM_TW + 2)) {
        const int ci = i / ((CM_TH + 2) * (CM_TW + 2));
        const int rm = i % ((CM_TH + 2) * (CM_TW + 2));
        const int r = rm / (CM_TW + 2), c = rm % (CM_TW + 2);
        const int hh = h0 + r - 1, ww = w0 + c - 1;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) xv[it] = x[(((long)b * 3 + ci) * H + hh) * W + ww];
        xdst[it] = ci * IPLANE + r * IW + c;
      }
    }
    // all global loads of both windows are issued before any arithmetic: one memory round trip per tile
    float4 yv[2][4], gv[2];
    bool full2[2], inb[2][4];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int wi = tid / 16 + 16 * half;
      const int wy = wi / (CM_TW / 2), wx = wi % (CM_TW / 2);
      const int oh = h0 / 2 + wy, ow = w0 / 2 + wx;
      full2[half] = (oh < Ho) && (ow < Wo);
      const int hh0 = h0 + 2 * wy, ww0 = w0 + 2 * wx;
      const float* yb = y + (((long)b * H + hh0) * W + ww0) * 64 + c4;
      gv[half] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (full2[half]) gv[half] = dktb_ld4(gout + (((long)b * Hq + oh + out_pad) * Wq + ow + out_pad) * 64 + c4);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        inb[half][k] = (hh0 + (k >> 1) < H) && (ww0 + (k & 1) < W);
        yv[half][k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (inb[half][k]) yv[half][k] = dktb_ld4(yb + ((long)(k >> 1) * W + (k & 1)) * 64);
      }
    }
#pragma unroll
    for (int it = 0; it < 3; ++it)
      if (xdst[it] >= 0) s_in[xdst[it]] = xv[it];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int wi = tid / 16 + 16 * half;
      const int wy = wi / (CM_TW / 2), wx = wi % (CM_TW / 2);
      const bool full = full2[half];
      float* sg = s_g + ((2 * wy) * CM_TW + 2 * wx) * CM_GLD + c4;
#pragma unroll
      for (int jp = 0; jp < 2; ++jp) {
        const int c = c4 + 2 * jp;
        const float2 mm = *reinterpret_cast<const float2*>(&s_k[0][c]), sc = *reinterpret_cast<const float2*>(&s_k[1][c]);
        const float2 bb = *reinterpret_cast<const float2*>(&s_k[2][c]), aa = *reinterpret_cast<const float2*>(&s_k[3][c]);
        const float2 bc = *reinterpret_cast<const float2*>(&s_k[4][c]);
        const float2 gg = jp ? make_float2(gv[half].z, gv[half].w) : make_float2(gv[half].x, gv[half].y);
        float2 d[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 v = jp ? make_float2(yv[half][k].z, yv[half][k].w) : make_float2(yv[half][k].x, yv[half][k].y);
          d[k] = make_float2(v.x - mm.x, v.y - mm.y);
        }
        float bx = fmaxf(fmaf(d[0].x, sc.x, bb.x), 0.f), by = fmaxf(fmaf(d[0].y, sc.y, bb.y), 0.f);
        int ax = 0, ay = 0;
#pragma unroll
        for (int k = 1; k < 4; ++k) {
          const float zx = fmaxf(fmaf(d[k].x, sc.x, bb.x), 0.f), zy = fmaxf(fmaf(d[k].y, sc.y, bb.y), 0.f);
          if (zx > bx) { bx = zx; ax = k; }
          if (zy > by) { by = zy; ay = k; }
        }
        const float hx = (full && bx > 0.f) ? sc.x * gg.x : 0.f, hy = (full && by > 0.f) ? sc.y * gg.y : 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float2 r = make_float2(fmaf(d[k].x, bc.x, aa.x + (k == ax ? hx : 0.f)),
                                 fmaf(d[k].y, bc.y, aa.y + (k == ay ? hy : 0.f)));
          if (!inb[half][k]) r = make_float2(0.f, 0.f);
          *reinterpret_cast<float2*>(sg + ((k >> 1) * CM_TW + (k & 1)) * CM_GLD + 2 * jp) = r;
        }
      }
    }
    __syncthreads();
    // ---- dW += P^T G on the tensor cores: two K = 8 pixel steps per warp
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      unsigned ah[2][4], al[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
          for (int kh = 0; kh < 2; ++kh) {
            float v = tone[mt][hh] ? 1.f : 0.f;
            if (tval[mt][hh]) v = s_in[toff[mt][hh] + ks * 8 + kh * 4];
            cm_split(v, ah[mt][hh + 2 * kh], al[mt][hh + 2 * kh]);
          }
      const float* gp = s_g + (py * CM_TW + pxb + ks * 8 + t) * CM_GLD + g;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        unsigned bh[2], bl[2];
        cm_split(gp[nt * 8], bh[0], bl[0]);
        cm_split(gp[4 * CM_GLD + nt * 8], bh[1], bl[1]);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          dktb_mma_m16n8k8_tf32(acc[mt][nt], al[mt], bh);
          dktb_mma_m16n8k8_tf32(acc[mt][nt], ah[mt], bl);
          dktb_mma_m16n8k8_tf32(acc[mt][nt], ah[mt], bh);
        }
      }
    }
  }
  // ---- the 8 warps of the CTA -> one partial [28][64] (fixed order, two rounds of four warps through s_g)
  float tot[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) tot[i] = 0.f;
  float* s_red = s_g;                                       // [4][28][64] floats = 28 KB
#pragma unroll
  for (int round = 0; round < 2; ++round) {
    __syncthreads();
    if ((warp >> 2) == round) {
      float* dst = s_red + (warp & 3) * 28 * 64;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const int r0 = 16 * mt + g, r1 = r0 + 8, cc = 8 * nt + 2 * t;
          if (r0 < 28) { dst[r0 * 64 + cc] = acc[mt][nt][0]; dst[r0 * 64 + cc + 1] = acc[mt][nt][1]; }
          if (r1 < 28) { dst[r1 * 64 + cc] = acc[mt][nt][2]; dst[r1 * 64 + cc + 1] = acc[mt][nt][3]; }
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const int idx = tid + 256 * i;
      tot[i] += (s_red[idx] + s_red[28 * 64 + idx]) + (s_red[2 * 28 * 64 + idx] + s_red[3 * 28 * 64 + idx]);
    }
  }
  float* out = partial + (long)blockIdx.x * (28 * 64);
#pragma unroll
  for (int i = 0; i < 7; ++i) out[tid + 256 * i] = tot[i];
}

extern "C" int dktb_conv1_wgrad_reduce(const float* partial, int nsplit, float* dw, float* db, cudaStream_t stream);

// Same contract as dktb_conv1_bwd_fused (scratch: dktb_conv1_wgrad_nsplit()*28*64 floats).
DKTB_EXPORT int dktb_conv1_bwd_fused_mma(const float* x, const float* y, const float* gout, const float* mean,
                                         const float* invstd, const float* gamma, const float* beta, const float* sums,
                                         float* dw, float* db, float* scratch, int B, int H, int W, int ipe, int out_pad,
                                         cudaStream_t stream) {
  DKTB_CHECK_ARG(x && y && gout && mean && invstd && gamma && beta && sums && dw && scratch);
  DKTB_CHECK_ARG(B > 0 && H > 0 && W > 0 && ipe > 0 && B % ipe == 0);
  const long ntiles = (long)B * ((W + CM_TW - 1) / CM_TW) * ((H + CM_TH - 1) / CM_TH);
  const int nsplit = (int)(ntiles < 296 ? ntiles : 296);      // 2 resident CTAs per SM x 148 SMs
  const float inv_count = 1.0f / ((float)ipe * (float)H * (float)W);
  DKTB_LAUNCH(conv1_bwd_mma_kernel, dim3(nsplit), dim3(256), 0, stream, x, y, gout, mean, invstd, gamma, beta, sums,
              scratch, B, H, W, ipe, out_pad, inv_count);
  return dktb_conv1_wgrad_reduce(scratch, nsplit, dw, db, stream);
}
