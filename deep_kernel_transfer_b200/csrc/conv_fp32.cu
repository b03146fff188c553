// fp32 (CUDA-core FFMA) convolution kernels of the Conv4/Conv6 backbone
// (reference: backbone.py:105-132 ConvBlock, 250-268 ConvNet).
//
//  * conv1 (Cin=3, K=27): bandwidth/issue bound, reads the loader's NCHW images directly and
//    writes NHWC pre-BN activations + per-tile BatchNorm partial sums.
//  * conv3x3 64->64 "flat-row" kernels over the zero-bordered NHWC layout [img][H+2][W+2][64]:
//    out[q][n] = sum_{r,s} sum_k A[q + (r-1)*Wp + (s-1)][k] * Wt[r*3+s][k][n]
//    -- the same kernel serves forward (Wt = w^T per tap) and dgrad (Wt = flipped taps), so the
//    zero border supplies the padding and no boundary test sits in the inner loop.
//  * wgrad for both.
// The 64->64 forward/dgrad/wgrad also exist as tcgen05 3xTF32 kernels (conv_tc.cu); these fp32
// versions handle the layers the tensor-core tiles do not cover and serve as on-device checkers.
#include "dktb_common.cuh"

#define C1_TH 4
#define C1_TW 32

__device__ __forceinline__ int c1_co(int cg, int j) { return j < 4 ? cg * 4 + j : 32 + cg * 4 + (j - 4); }

// ------------------------------------------------------------------------------------------------
// conv1 forward: x [B,3,H,W] (NCHW) -> y [B,H,W,64] (NHWC), bias added; optional BN partial sums
// partials[((b*TY+ty)*TX+tx)][2][64] = {sum, sum of squares} over the tile's valid pixels.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ y,
                                                        float* __restrict__ partials, int H, int W) {
  __shared__ __align__(16) float s_w[27 * 64];
  __shared__ float s_in[3][C1_TH + 2][C1_TW + 4];
  __shared__ float s_red[2][32][64 + 1];
  const int tid = threadIdx.x;
  const int b = blockIdx.z, h0 = blockIdx.y * C1_TH, w0 = blockIdx.x * C1_TW;
  for (int i = tid; i < 27 * 64; i += 256) {
    int co = i / 27, k = i % 27;
    s_w[k * 64 + co] = w[i];
  }
  for (int i = tid; i < 3 * (C1_TH + 2) * (C1_TW + 2); i += 256) {
    int ci = i / ((C1_TH + 2) * (C1_TW + 2));
    int rem = i % ((C1_TH + 2) * (C1_TW + 2));
    int r = rem / (C1_TW + 2), c = rem % (C1_TW + 2);
    int hh = h0 + r - 1, ww = w0 + c - 1;
    float v = 0.f;
    if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = x[(((long)b * 3 + ci) * H + hh) * W + ww];
    s_in[ci][r][c] = v;
  }
  __syncthreads();
  const int cg = tid % 8, pg = tid / 8;
  const int row = pg / 8, col0 = (pg % 8) * 4;
  float acc[4][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float bv = bias ? bias[c1_co(cg, j)] : 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p) acc[p][j] = bv;
  }
#pragma unroll
  for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float v[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) v[c] = s_in[ci][row + r][col0 + c];
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int k = ci * 9 + r * 3 + s;
        const float4 wa = dktb_ld4(&s_w[k * 64 + cg * 4]);
        const float4 wb = dktb_ld4(&s_w[k * 64 + 32 + cg * 4]);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float xv = v[p + s];
          acc[p][0] = fmaf(xv, wa.x, acc[p][0]);
          acc[p][1] = fmaf(xv, wa.y, acc[p][1]);
          acc[p][2] = fmaf(xv, wa.z, acc[p][2]);
          acc[p][3] = fmaf(xv, wa.w, acc[p][3]);
          acc[p][4] = fmaf(xv, wb.x, acc[p][4]);
          acc[p][5] = fmaf(xv, wb.y, acc[p][5]);
          acc[p][6] = fmaf(xv, wb.z, acc[p][6]);
          acc[p][7] = fmaf(xv, wb.w, acc[p][7]);
        }
      }
    }
  }
  float sm[8], sq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sm[j] = sq[j] = 0.f;
  const int hh = h0 + row;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int ww = w0 + col0 + p;
    if (hh < H && ww < W) {
      float* o = y + (((long)b * H + hh) * W + ww) * 64;
      dktb_st4(o + cg * 4, make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]));
      dktb_st4(o + 32 + cg * 4, make_float4(acc[p][4], acc[p][5], acc[p][6], acc[p][7]));
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sm[j] += acc[p][j];
        sq[j] = fmaf(acc[p][j], acc[p][j], sq[j]);
      }
    }
  }
  if (partials != nullptr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s_red[0][pg][c1_co(cg, j)] = sm[j];
      s_red[1][pg][c1_co(cg, j)] = sq[j];
    }
    __syncthreads();
    if (tid < 128) {
      const int which = tid / 64, co = tid % 64;
      float t = 0.f;
      for (int g = 0; g < 32; ++g) t += s_red[which][g][co];
      const long blk = ((long)b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      partials[(blk * 2 + which) * 64 + co] = t;
    }
  }
}

DKTB_EXPORT int dktb_conv1_tiles(int H, int W) { return ((H + C1_TH - 1) / C1_TH) * ((W + C1_TW - 1) / C1_TW); }

DKTB_EXPORT int dktb_conv1_fwd(const float* x, const float* w, const float* bias, float* y, float* partials, int B,
                               int H, int W, cudaStream_t stream) {
  DKTB_CHECK_ARG(x && w && y && B > 0 && H > 0 && W > 0 && B <= 65535);
  dim3 grid((W + C1_TW - 1) / C1_TW, (H + C1_TH - 1) / C1_TH, B);
  DKTB_LAUNCH(conv1_fwd_kernel, grid, dim3(256), 0, stream, x, w, bias, y, partials, H, W);
  return dktb_launch_status();
}

// ------------------------------------------------------------------------------------------------
// conv1 wgrad: dw[co][ci][r][s] = sum x[b][ci][h+r-1][w+s-1] * gy[b][h][w][co];  db[co] = sum gy.
// Persistent CTAs loop over 4x32 pixel tiles and keep their 27x64(+64) partial in registers;
// partial buffer [nsplit][28*64] (row 27 = bias) is reduced by dktb_reduce_partials.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv1_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                                          float* __restrict__ partial, int B, int H, int W) {
  __shared__ float s_in[3][C1_TH + 2][C1_TW + 4];
  __shared__ float s_g[C1_TH * C1_TW][64 + 1];
  const int tid = threadIdx.x;
  const int co = tid % 64, grp = tid / 64;      // grp handles (ci,r) combos grp, grp+4, grp+8
  const int TX = (W + C1_TW - 1) / C1_TW, TY = (H + C1_TH - 1) / C1_TH;
  const long ntiles = (long)B * TX * TY;
  float acc[3][3];
  float accb = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int s = 0; s < 3; ++s) acc[a][s] = 0.f;
  for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int b = (int)(t / (TX * TY));
    const int rem = (int)(t % (TX * TY));
    const int h0 = (rem / TX) * C1_TH, w0 = (rem % TX) * C1_TW;
    __syncthreads();
    for (int i = tid; i < 3 * (C1_TH + 2) * (C1_TW + 2); i += 256) {
      int ci = i / ((C1_TH + 2) * (C1_TW + 2));
      int rm = i % ((C1_TH + 2) * (C1_TW + 2));
      int r = rm / (C1_TW + 2), c = rm % (C1_TW + 2);
      int hh = h0 + r - 1, ww = w0 + c - 1;
      float v = 0.f;
      if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = x[(((long)b * 3 + ci) * H + hh) * W + ww];
      s_in[ci][r][c] = v;
    }
    for (int i = tid; i < C1_TH * C1_TW * 16; i += 256) {
      const int p = i / 16, c4 = (i % 16) * 4;
      const int hh = h0 + p / C1_TW, ww = w0 + p % C1_TW;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (hh < H && ww < W) v = dktb_ld4(gy + (((long)b * H + hh) * W + ww) * 64 + c4);
      s_g[p][c4 + 0] = v.x;
      s_g[p][c4 + 1] = v.y;
      s_g[p][c4 + 2] = v.z;
      s_g[p][c4 + 3] = v.w;
    }
    __syncthreads();
    for (int row = 0; row < C1_TH; ++row) {
      for (int c0 = 0; c0 < C1_TW; c0 += 4) {
        float g[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) g[p] = s_g[row * C1_TW + c0 + p][co];
        if (grp == 0) accb += (g[0] + g[1]) + (g[2] + g[3]);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const int combo = grp + 4 * a;
          if (combo < 9) {
            const int ci = combo / 3, r = combo % 3;
            float v[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) v[c] = s_in[ci][row + r][c0 + c];
#pragma unroll
            for (int s = 0; s < 3; ++s)
#pragma unroll
              for (int p = 0; p < 4; ++p) acc[a][s] = fmaf(v[p + s], g[p], acc[a][s]);
          }
        }
      }
    }
  }
  float* out = partial + (long)blockIdx.x * (28 * 64);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int combo = grp + 4 * a;
    if (combo < 9) {
      const int ci = combo / 3, r = combo % 3;
#pragma unroll
      for (int s = 0; s < 3; ++s) out[(ci * 9 + r * 3 + s) * 64 + co] = acc[a][s];
    }
  }
  if (grp == 0) out[27 * 64 + co] = accb;
}

// partial [nsplit][28*64] -> dw [64][3][3][3] (reference layout), db [64]
__global__ void conv1_wgrad_reduce_kernel(const float* __restrict__ partial, int nsplit, float* __restrict__ dw,
                                          float* __restrict__ db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 28 * 64) return;
  float t = 0.f;
  for (int s = 0; s < nsplit; ++s) t += partial[(long)s * (28 * 64) + i];
  const int k = i / 64, co = i % 64;
  if (k < 27) dw[co * 27 + k] = t;
  else if (db != nullptr) db[co] = t;
}

// ------------------------------------------------------------------------------------------------------------------
// Fused first-layer backward: gy = BatchNorm/ReLU/MaxPool backward of the pooled gradient (second pass: the episode sums
// s1, s2 are already known) evaluated tile by tile in shared memory and consumed on the spot by the weight-gradient
// accumulation -- the 1.8 MB/image gradient of the pre-BN map is never written to (or re-read from) HBM.
// One CTA loops over 4 x 32 pixel tiles (2 x 16 pool windows); thread (co, row) accumulates all 27 taps of its tile row.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 3) conv1_bwd_fused_kernel(
    const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ gout,
    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
    const float* __restrict__ beta, const float* __restrict__ sums, float* __restrict__ partial, int B, int H, int W,
    int ipe, int out_pad, float inv_count) {
  constexpr int GLD = 68;                                  // gradient tile pitch: float4 rows
  __shared__ __align__(16) float s_in[3][C1_TH + 2][C1_TW + 4];
  __shared__ __align__(16) float s_g[C1_TH * C1_TW * GLD];
  // per-episode channel constants: mean, gamma*invstd, beta, A = -ss*s1/n, Bc = -ss*s2/n*invstd, ss
  __shared__ __align__(16) float s_k[6][64];
  float (*s_fin)[28][64] = reinterpret_cast<float (*)[28][64]>(s_g);            // reused after the tile loop
  const int tid = threadIdx.x;
  const int co = tid % 64, row = tid / 64;
  const int TX = (W + C1_TW - 1) / C1_TW, TY = (H + C1_TH - 1) / C1_TH;
  const long ntiles = (long)B * TX * TY;
  const int Ho = H / 2, Wo = W / 2, Hq = Ho + 2 * out_pad, Wq = Wo + 2 * out_pad;
  const int c4 = (tid % 16) * 4;                            // staging role: 4 channels, windows tid/16 and tid/16 + 16
  float acc[3][3][3];
  float accb = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) acc[a][r][s] = 0.f;
  int e_cur = -1;
  for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int b = (int)(t / (TX * TY));
    const int rem = (int)(t % (TX * TY));
    const int h0 = (rem / TX) * C1_TH, w0 = (rem % TX) * C1_TW;
    const int e = b / ipe;
    if (e != e_cur) {                                       // uniform over the CTA (depends on t only)
      if (tid < 64) {
        const float is = invstd[e * 64 + tid], ss = gamma[tid] * is;
        s_k[0][tid] = mean[e * 64 + tid];
        s_k[1][tid] = ss;
        s_k[2][tid] = beta[tid];
        s_k[3][tid] = -ss * (sums[(long)e * 128 + tid] * inv_count);
        s_k[4][tid] = -ss * (sums[(long)e * 128 + 64 + tid] * inv_count) * is;
        s_k[5][tid] = ss;
      }
      e_cur = e;
    }
    __syncthreads();
    for (int i = tid; i < 3 * (C1_TH + 2) * (C1_TW + 2); i += 256) {
      const int ci = i / ((C1_TH + 2) * (C1_TW + 2));
      const int rm = i % ((C1_TH + 2) * (C1_TW + 2));
      const int r = rm / (C1_TW + 2), c = rm % (C1_TW + 2);
      const int hh = h0 + r - 1, ww = w0 + c - 1;
      float v = 0.f;
      if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = x[(((long)b * 3 + ci) * H + hh) * W + ww];
      s_in[ci][r][c] = v;
    }
    // gradient tile: work item = (pool window of the tile, 4 channels), handled as two channel pairs to bound the
    // live registers next to the 28 accumulators; 32 windows x 16 channel groups = 2 items per thread
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int wi = tid / 16 + 16 * half;
      const int wy = wi / (C1_TW / 2), wx = wi % (C1_TW / 2);
      const int oh = h0 / 2 + wy, ow = w0 / 2 + wx;
      const bool full = (oh < Ho) && (ow < Wo);
      const int hh0 = h0 + 2 * wy, ww0 = w0 + 2 * wx;
      const bool inr[2] = {hh0 < H, hh0 + 1 < H}, inc[2] = {ww0 < W, ww0 + 1 < W};
      const float* yb = y + (((long)b * H + hh0) * W + ww0) * 64 + c4;
      const float* gb = gout + (((long)b * Hq + oh + out_pad) * Wq + ow + out_pad) * 64 + c4;
      float* sg = s_g + ((2 * wy) * C1_TW + 2 * wx) * GLD + c4;
#pragma unroll
      for (int jp = 0; jp < 2; ++jp) {
        const int c = c4 + 2 * jp;
        const float2 mm = *reinterpret_cast<const float2*>(&s_k[0][c]), sc = *reinterpret_cast<const float2*>(&s_k[1][c]);
        const float2 bb = *reinterpret_cast<const float2*>(&s_k[2][c]), aa = *reinterpret_cast<const float2*>(&s_k[3][c]);
        const float2 bc = *reinterpret_cast<const float2*>(&s_k[4][c]);
        float2 gg = make_float2(0.f, 0.f);
        if (full) gg = *reinterpret_cast<const float2*>(gb + 2 * jp);
        float2 d[4];                                        // y - mean at the 4 window positions
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float2 v = make_float2(0.f, 0.f);
          if (inr[k >> 1] && inc[k & 1]) v = *reinterpret_cast<const float2*>(yb + ((long)(k >> 1) * W + (k & 1)) * 64 + 2 * jp);
          d[k] = make_float2(v.x - mm.x, v.y - mm.y);
        }
        // first maximum of the post-ReLU values in scan order receives the pooled gradient (if positive)
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
          if (!(inr[k >> 1] && inc[k & 1])) r = make_float2(0.f, 0.f);
          *reinterpret_cast<float2*>(sg + ((k >> 1) * C1_TW + (k & 1)) * GLD + 2 * jp) = r;
        }
      }
    }
    __syncthreads();
    for (int c0 = 0; c0 < C1_TW; c0 += 4) {
      float g[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) g[p] = s_g[(row * C1_TW + c0 + p) * GLD + co];
      accb += (g[0] + g[1]) + (g[2] + g[3]);
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const float4 v0 = dktb_ld4(&s_in[ci][row + r][c0]);
          const float2 v1 = *reinterpret_cast<const float2*>(&s_in[ci][row + r][c0 + 4]);
          const float v[6] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y};
#pragma unroll
          for (int s = 0; s < 3; ++s)
#pragma unroll
            for (int p = 0; p < 4; ++p) acc[ci][r][s] = fmaf(v[p + s], g[p], acc[ci][r][s]);
        }
    }
  }
  // the four tile rows of the CTA -> one partial (fixed order)
  __syncthreads();
#pragma unroll
  for (int ci = 0; ci < 3; ++ci)
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) s_fin[row][ci * 9 + r * 3 + s][co] = acc[ci][r][s];
  s_fin[row][27][co] = accb;
  __syncthreads();
  float* out = partial + (long)blockIdx.x * (28 * 64);
  for (int i = tid; i < 28 * 64; i += 256) {
    const int k = i / 64, c = i % 64;
    out[i] = (s_fin[0][k][c] + s_fin[1][k][c]) + (s_fin[2][k][c] + s_fin[3][k][c]);
  }
}

// partial [nsplit][28*64] -> dw / db (fixed order); shared with the tensor-core first-block backward
DKTB_EXPORT int dktb_conv1_wgrad_reduce(const float* partial, int nsplit, float* dw, float* db, cudaStream_t stream) {
  DKTB_CHECK_ARG(partial && dw && nsplit > 0);
  DKTB_LAUNCH(conv1_wgrad_reduce_kernel, dim3(7), dim3(256), 0, stream, partial, nsplit, dw, db);
  return dktb_launch_status();
}

DKTB_EXPORT int dktb_conv1_wgrad_nsplit(void) { return 592; }

DKTB_EXPORT int dktb_conv1_wgrad(const float* x, const float* gy, float* dw, float* db, float* scratch, int B, int H,
                                 int W, cudaStream_t stream) {
  DKTB_CHECK_ARG(x && gy && dw && scratch && B > 0);
  const int TX = (W + C1_TW - 1) / C1_TW, TY = (H + C1_TH - 1) / C1_TH;
  long ntiles = (long)B * TX * TY;
  int nsplit = (int)(ntiles < 592 ? ntiles : 592);
  DKTB_LAUNCH(conv1_wgrad_kernel, dim3(nsplit), dim3(256), 0, stream, x, gy, scratch, B, H, W);
  DKTB_LAUNCH(conv1_wgrad_reduce_kernel, dim3(7), dim3(256), 0, stream, scratch, nsplit, dw, db);
  return dktb_launch_status();
}

// ------------------------------------------------------------------------------------------------
// weight re-layout for the 64->64 kernels.  w_ref [co][ci][3][3] ->
//   wt_fwd [tap][ci][co]           (forward:  out[q][co] += A[q+off(tap)][ci] * wt_fwd[tap][ci][co])
//   wt_dgrad [tap'][co][ci] with tap' = 8 - tap (dgrad: gx[q][ci] += gy[q+off(tap')][co] * w[co][ci][tap])
// ------------------------------------------------------------------------------------------------
__global__ void prep_weights_kernel(const float* __restrict__ w, float* __restrict__ wt_fwd,
                                    float* __restrict__ wt_dgrad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 64 * 9) return;
  const int tap = i % 9, ci = (i / 9) % 64, co = i / (9 * 64);
  const float v = w[i];
  if (wt_fwd) wt_fwd[(tap * 64 + ci) * 64 + co] = v;
  if (wt_dgrad) wt_dgrad[((8 - tap) * 64 + co) * 64 + ci] = v;
}

DKTB_EXPORT int dktb_prep_weights(const float* w, float* wt_fwd, float* wt_dgrad, cudaStream_t stream) {
  DKTB_CHECK_ARG(w != nullptr);
  DKTB_LAUNCH(prep_weights_kernel, dim3((64 * 64 * 9 + 255) / 256), dim3(256), 0, stream, w, wt_fwd, wt_dgrad);
  return dktb_launch_status();
}

// ------------------------------------------------------------------------------------------------
// conv3x3 64->64 flat-row kernel (forward and dgrad).  One CTA = 128 consecutive padded-flat rows of
// one image starting at the first interior pixel; the A halo (128 + 2*(Wp+1) rows) is staged once in
// shared memory and reused by the 9 taps; weights are staged per tap.
// ------------------------------------------------------------------------------------------------
#define CF_ROWS 128
#define CF_LD 68  // smem row stride in floats (272 B): consecutive rows land in different banks

DKTB_EXPORT int dktb_conv3x3_tiles(int H, int W) {
  const int Hp = H + 2, Wp = W + 2;
  const int span = Hp * Wp - 2 * (Wp + 1);   // first interior .. last interior (inclusive range length)
  return (span + CF_ROWS - 1) / CF_ROWS;
}

__global__ void __launch_bounds__(256) conv3x3_flat_kernel(const float* __restrict__ a, const float* __restrict__ wt,
                                                           const float* __restrict__ bias, float* __restrict__ out,
                                                           float* __restrict__ partials, int B, int H, int W) {
  DKTB_DYN_SMEM(float, smem);
  const int Hp = H + 2, Wp = W + 2;
  const int halo = CF_ROWS + 2 * (Wp + 1);
  float* s_a = smem;                       // [halo][CF_LD]
  float* s_w = smem + (long)halo * CF_LD;  // [64][64]
  float* s_red = s_w + 64 * 64;            // [2][16][64]
  const int tid = threadIdx.x;
  const int img = blockIdx.y;
  const int q0 = (Wp + 1) + blockIdx.x * CF_ROWS;   // image-local flat index of the tile's first row
  const long img_base = (long)img * Hp * Wp;
  const long total_rows = (long)B * Hp * Wp;
  // stage the halo: rows q0-(Wp+1) .. q0+127+(Wp+1)
  for (int i = tid; i < halo * 16; i += 256) {
    const int r = i / 16, c4 = (i % 16) * 4;
    const long grow = img_base + q0 - (Wp + 1) + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (grow >= 0 && grow < total_rows) v = dktb_ld4(a + grow * 64 + c4);
    dktb_st4(&s_a[r * CF_LD + c4], v);
  }
  const int co4 = tid % 16, rg = tid / 16;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int tap = 0; tap < 9; ++tap) {
    __syncthreads();
    for (int i = tid; i < 64 * 16; i += 256) dktb_st4(&s_w[i * 4], dktb_ld4(wt + (long)tap * 4096 + i * 4));
    __syncthreads();
    const int shift = (tap / 3) * Wp + (tap % 3);
#pragma unroll 4
    for (int k4 = 0; k4 < 16; ++k4) {
      const float4 w0 = dktb_ld4(&s_w[(k4 * 4 + 0) * 64 + co4 * 4]);
      const float4 w1 = dktb_ld4(&s_w[(k4 * 4 + 1) * 64 + co4 * 4]);
      const float4 w2 = dktb_ld4(&s_w[(k4 * 4 + 2) * 64 + co4 * 4]);
      const float4 w3 = dktb_ld4(&s_w[(k4 * 4 + 3) * 64 + co4 * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 av = dktb_ld4(&s_a[(rg + 16 * i + shift) * CF_LD + k4 * 4]);
        acc[i][0] = fmaf(av.x, w0.x, acc[i][0]);
        acc[i][1] = fmaf(av.x, w0.y, acc[i][1]);
        acc[i][2] = fmaf(av.x, w0.z, acc[i][2]);
        acc[i][3] = fmaf(av.x, w0.w, acc[i][3]);
        acc[i][0] = fmaf(av.y, w1.x, acc[i][0]);
        acc[i][1] = fmaf(av.y, w1.y, acc[i][1]);
        acc[i][2] = fmaf(av.y, w1.z, acc[i][2]);
        acc[i][3] = fmaf(av.y, w1.w, acc[i][3]);
        acc[i][0] = fmaf(av.z, w2.x, acc[i][0]);
        acc[i][1] = fmaf(av.z, w2.y, acc[i][1]);
        acc[i][2] = fmaf(av.z, w2.z, acc[i][2]);
        acc[i][3] = fmaf(av.z, w2.w, acc[i][3]);
        acc[i][0] = fmaf(av.w, w3.x, acc[i][0]);
        acc[i][1] = fmaf(av.w, w3.y, acc[i][1]);
        acc[i][2] = fmaf(av.w, w3.z, acc[i][2]);
        acc[i][3] = fmaf(av.w, w3.w, acc[i][3]);
      }
    }
  }
  float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias != nullptr) bv = dktb_ld4(bias + co4 * 4);
  float sm[4] = {0.f, 0.f, 0.f, 0.f}, sq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int q = q0 + rg + 16 * i;
    const int hp = q / Wp, wp = q % Wp;
    if (q < Hp * Wp && hp >= 1 && hp <= H && wp >= 1 && wp <= W) {
      const float4 o = make_float4(acc[i][0] + bv.x, acc[i][1] + bv.y, acc[i][2] + bv.z, acc[i][3] + bv.w);
      dktb_st4(out + (img_base + q) * 64 + co4 * 4, o);
      sm[0] += o.x; sm[1] += o.y; sm[2] += o.z; sm[3] += o.w;
      sq[0] = fmaf(o.x, o.x, sq[0]); sq[1] = fmaf(o.y, o.y, sq[1]);
      sq[2] = fmaf(o.z, o.z, sq[2]); sq[3] = fmaf(o.w, o.w, sq[3]);
    }
  }
  if (partials != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s_red[(0 * 16 + rg) * 64 + co4 * 4 + j] = sm[j];
      s_red[(1 * 16 + rg) * 64 + co4 * 4 + j] = sq[j];
    }
    __syncthreads();
    if (tid < 128) {
      const int which = tid / 64, co = tid % 64;
      float t = 0.f;
      for (int g = 0; g < 16; ++g) t += s_red[(which * 16 + g) * 64 + co];
      const long blk = (long)img * gridDim.x + blockIdx.x;
      partials[(blk * 2 + which) * 64 + co] = t;
    }
  }
}

static inline size_t conv3x3_flat_smem(int W) {
  const int Wp = W + 2;
  return ((size_t)(CF_ROWS + 2 * (Wp + 1)) * CF_LD + 64 * 64 + 2 * 16 * 64) * sizeof(float);
}

DKTB_EXPORT int dktb_conv3x3_fwd(const float* a, const float* wt, const float* bias, float* out, float* partials, int B,
                                 int H, int W, cudaStream_t stream) {
  DKTB_CHECK_ARG(a && wt && out && B > 0 && H > 0 && W > 0 && B <= 65535);
  const size_t smem = conv3x3_flat_smem(W);
  DKTB_CHECK_ARG(smem <= 227 * 1024);
  cudaFuncSetAttribute(conv3x3_flat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(dktb_conv3x3_tiles(H, W), B);
  DKTB_LAUNCH(conv3x3_flat_kernel, grid, dim3(256), smem, stream, a, wt, bias, out, partials, B, H, W);
  return dktb_launch_status();
}

// ------------------------------------------------------------------------------------------------
// conv3x3 64->64 wgrad: dW[tap][ci][co] = sum_q A[q+off(tap)][ci] * gy[q][co] over all padded-flat rows
// (gy is zero on the border, so no masking).  grid = (nsplit, 3 filter rows); every CTA strides over
// 64-row chunks; partial [nsplit][9*64*64 + 64]; reduced + transposed to w_ref layout afterwards.
// ------------------------------------------------------------------------------------------------
#define WG_ROWS 64
#define WG_PSTRIDE (9 * 64 * 64 + 64)

__global__ void __launch_bounds__(256) conv3x3_wgrad_kernel(const float* __restrict__ a, const float* __restrict__ gy,
                                                            float* __restrict__ partial, long total_rows, int Wp) {
  __shared__ __align__(16) float s_g[WG_ROWS][CF_LD];
  __shared__ __align__(16) float s_a[WG_ROWS + 2][CF_LD];
  const int tid = threadIdx.x;
  const int r = blockIdx.y;                        // filter row
  const int co4 = tid % 16, ci4 = tid / 16;
  float acc[3][4][4];
  float accb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[s][u][v] = 0.f;
  const long nchunks = (total_rows + WG_ROWS - 1) / WG_ROWS;
  for (long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const long q0 = ch * WG_ROWS;
    __syncthreads();
    for (int i = tid; i < WG_ROWS * 16; i += 256) {
      const int j = i / 16, c4 = (i % 16) * 4;
      const long q = q0 + j;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q < total_rows) v = dktb_ld4(gy + q * 64 + c4);
      dktb_st4(&s_g[j][c4], v);
    }
    for (int i = tid; i < (WG_ROWS + 2) * 16; i += 256) {
      const int j = i / 16, c4 = (i % 16) * 4;
      const long q = q0 + j + (long)(r - 1) * Wp - 1;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q >= 0 && q < total_rows) v = dktb_ld4(a + q * 64 + c4);
      dktb_st4(&s_a[j][c4], v);
    }
    __syncthreads();
#pragma unroll 2
    for (int j = 0; j < WG_ROWS; ++j) {
      const float4 g = dktb_ld4(&s_g[j][co4 * 4]);
      if (ci4 == 0) { accb[0] += g.x; accb[1] += g.y; accb[2] += g.z; accb[3] += g.w; }
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const float4 av = dktb_ld4(&s_a[j + s][ci4 * 4]);
        acc[s][0][0] = fmaf(av.x, g.x, acc[s][0][0]); acc[s][0][1] = fmaf(av.x, g.y, acc[s][0][1]);
        acc[s][0][2] = fmaf(av.x, g.z, acc[s][0][2]); acc[s][0][3] = fmaf(av.x, g.w, acc[s][0][3]);
        acc[s][1][0] = fmaf(av.y, g.x, acc[s][1][0]); acc[s][1][1] = fmaf(av.y, g.y, acc[s][1][1]);
        acc[s][1][2] = fmaf(av.y, g.z, acc[s][1][2]); acc[s][1][3] = fmaf(av.y, g.w, acc[s][1][3]);
        acc[s][2][0] = fmaf(av.z, g.x, acc[s][2][0]); acc[s][2][1] = fmaf(av.z, g.y, acc[s][2][1]);
        acc[s][2][2] = fmaf(av.z, g.z, acc[s][2][2]); acc[s][2][3] = fmaf(av.z, g.w, acc[s][2][3]);
        acc[s][3][0] = fmaf(av.w, g.x, acc[s][3][0]); acc[s][3][1] = fmaf(av.w, g.y, acc[s][3][1]);
        acc[s][3][2] = fmaf(av.w, g.z, acc[s][3][2]); acc[s][3][3] = fmaf(av.w, g.w, acc[s][3][3]);
      }
    }
  }
  float* out = partial + (long)blockIdx.x * WG_PSTRIDE;
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int u = 0; u < 4; ++u)
      dktb_st4(out + (((r * 3 + s) * 64) + ci4 * 4 + u) * 64 + co4 * 4,
               make_float4(acc[s][u][0], acc[s][u][1], acc[s][u][2], acc[s][u][3]));
  if (r == 1 && ci4 == 0) dktb_st4(out + 9 * 64 * 64 + co4 * 4, make_float4(accb[0], accb[1], accb[2], accb[3]));
}

// partial [nsplit][tap][ci][co] (+[64] bias row) -> dw [co][ci][3][3], db [co]
__global__ void conv3x3_wgrad_reduce_kernel(const float* __restrict__ partial, int nsplit, float* __restrict__ dw,
                                            float* __restrict__ db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= WG_PSTRIDE) return;
  float t = 0.f;
  for (int s = 0; s < nsplit; ++s) t += partial[(long)s * WG_PSTRIDE + i];
  if (i < 9 * 64 * 64) {
    const int co = i % 64, ci = (i / 64) % 64, tap = i / 4096;
    dw[(co * 64 + ci) * 9 + tap] = t;
  } else if (db != nullptr) {
    db[i - 9 * 64 * 64] = t;
  }
}

DKTB_EXPORT int dktb_conv3x3_wgrad_reduce(const float* partial, int nsplit, float* dw, float* db,
                                          cudaStream_t stream) {
  DKTB_CHECK_ARG(partial && dw && nsplit > 0);
  DKTB_LAUNCH(conv3x3_wgrad_reduce_kernel, dim3((WG_PSTRIDE + 255) / 256), dim3(256), 0, stream, partial, nsplit, dw,
              db);
  return dktb_launch_status();
}

DKTB_EXPORT int dktb_conv3x3_wgrad_nsplit(void) { return 296; }
DKTB_EXPORT long dktb_conv3x3_wgrad_scratch_floats(void) { return (long)296 * WG_PSTRIDE; }

DKTB_EXPORT int dktb_conv3x3_wgrad(const float* a, const float* gy, float* dw, float* db, float* scratch, int B, int H,
                                   int W, cudaStream_t stream) {
  DKTB_CHECK_ARG(a && gy && dw && scratch && B > 0);
  const int Hp = H + 2, Wp = W + 2;
  const long total_rows = (long)B * Hp * Wp;
  const long nchunks = (total_rows + WG_ROWS - 1) / WG_ROWS;
  const int nsplit = (int)(nchunks < 296 ? nchunks : 296);
  DKTB_LAUNCH(conv3x3_wgrad_kernel, dim3(nsplit, 3), dim3(256), 0, stream, a, gy, scratch, total_rows, Wp);
  DKTB_LAUNCH(conv3x3_wgrad_reduce_kernel, dim3((WG_PSTRIDE + 255) / 256), dim3(256), 0, stream, scratch, nsplit, dw,
              db);
  return dktb_launch_status();
}

// Fused second half of the first block's backward (after dktb_bn_relu_pool_bwd(..., gy = NULL, ...) produced `sums`
// [B/ipe][2][64] and the BatchNorm parameter gradients): BN/ReLU/MaxPool2d(2) backward of `gout` [B,H/2+2p,W/2+2p,64]
// (p = out_pad) fused into the conv1 weight / bias gradient.  y [B,H,W,64] is the pre-BN conv1 output.
// scratch: dktb_conv1_wgrad_nsplit()*28*64 floats.
DKTB_EXPORT int dktb_conv1_bwd_fused(const float* x, const float* y, const float* gout, const float* mean,
                                     const float* invstd, const float* gamma, const float* beta, const float* sums,
                                     float* dw, float* db, float* scratch, int B, int H, int W, int ipe, int out_pad,
                                     cudaStream_t stream) {
  DKTB_CHECK_ARG(x && y && gout && mean && invstd && gamma && beta && sums && dw && scratch);
  DKTB_CHECK_ARG(B > 0 && H > 0 && W > 0 && ipe > 0 && B % ipe == 0);
  const long ntiles = (long)B * ((W + C1_TW - 1) / C1_TW) * ((H + C1_TH - 1) / C1_TH);
  const int nsplit = (int)(ntiles < 444 ? ntiles : 444);      // 3 resident CTAs per SM x 148 SMs, one wave
  const float inv_count = 1.0f / ((float)ipe * (float)H * (float)W);
  DKTB_LAUNCH(conv1_bwd_fused_kernel, dim3(nsplit), dim3(256), 0, stream, x, y, gout, mean, invstd, gamma, beta, sums,
              scratch, B, H, W, ipe, out_pad, inv_count);
  DKTB_LAUNCH(conv1_wgrad_reduce_kernel, dim3(7), dim3(256), 0, stream, (const float*)scratch, nsplit, dw, db);
  return dktb_launch_status();
}
