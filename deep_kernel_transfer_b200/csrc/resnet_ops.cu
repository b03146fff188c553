// Channel-generic NHWC building blocks of the ResNet backbones (reference backbone.py:135-247 SimpleBlock /
// BottleneckBlock, 330-376 ResNet): BatchNorm2d with PER-EPISODE batch statistics for any channel count (+ fused
// residual add and ReLU), MaxPool2d(3, stride 2, pad 1), global AvgPool2d, forward and backward.  HBM-bound
// element-wise / reduction kernels; reductions are two-stage with a fixed order (deterministic).
#include "dktb_common.cuh"

// ------------------------------------------------------------------------------------------------ tensor layouts
// Every NHWC tensor argument of the *_l entry points is either dense [B][H][W][C] or the INTERIOR of a padded-flat buffer
// [B][H+2][W+2][C] (pointer = buffer base; the layout the tcgen05 convolutions of csrc/conv_tcg.cu read and write, so that
// no copy separates them from the element-wise kernels).  `lay` holds one bit per argument; the border of a padded buffer
// is never read or written here.
enum { kLayX = 1, kLayY = 2, kLayGy = 4, kLayGx = 8, kLayRes = 16 };
struct NhwcGeom { int HW, W, HWp; };        // HWp = (H + 2) * (W + 2)
// The element-wise kernels run one grid row (blockIdx.y) per image and decompose the in-image index with 32-bit unsigned
// arithmetic (a 64-bit division per element costs more than the memory access it addresses).
// pixel p of image b -> row index of a dense / padded-flat tensor
__device__ __forceinline__ long dense_row(int b, unsigned p, const NhwcGeom& gm) { return (long)b * gm.HW + p; }
__device__ __forceinline__ long padded_row(int b, unsigned p, const NhwcGeom& gm) {
  const unsigned h = p / (unsigned)gm.W, w = p - h * (unsigned)gm.W;
  return (long)b * gm.HWp + (long)(h + 1) * (gm.W + 2) + (w + 1);
}

// ------------------------------------------------------------------------------------------------ BN statistics
// pixel splits per image: enough CTAs to fill the machine when B x C / 64 alone does not (the 64-channel layers at 56x56 /
// 112x112 hold most of the bytes); a split is one more row of `partial`, summed by the finalize kernels in a fixed order
static int bn2d_splits(int B, int HW, int C) {
  const long ctas = (long)B * ((C + 63) / 64);
  int s = (int)((4 * 148 + ctas - 1) / ctas);
  const int smax = HW / 256 > 0 ? HW / 256 : 1;
  if (s > smax) s = smax;
  if (s > 8) s = 8;
  return s < 1 ? 1 : s;
}
DKTB_EXPORT long dktb_bn2d_partial_floats(int B, int HW, int C) { return (long)B * bn2d_splits(B, HW, C) * C * 2; }

// partial[b * S + split][c][2] = sum, sumsq over the split's pixels of image b.  grid (ceil(C/64), B, S), 256 threads =
// 16 channel quads (float4 loads, C % 4 == 0) x 16 pixel slices
__global__ void __launch_bounds__(256) bn2d_partial_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                           const float* __restrict__ y, const float* __restrict__ mean,
                                                           const float* __restrict__ invstd, float* __restrict__ partial,
                                                           int HW, int C, int ipe, int mode, NhwcGeom gm, int lay) {
  // mode 0: (x, x^2);  mode 1 (backward): g' = g * (y > 0 if y given), (g', g' * xhat)
  // Sums in DOUBLE: dbeta / dgamma (and the batch mean) are sums of up to 12 544 signed terms per image that largely cancel
  // (cond = sum|t| / |sum t| ~ 1e2 .. 1e3 for the first layers' gradients); a sequential float32 chain over HW / 4 pixels
  // left 1e-4 on the stem BatchNorm's gradients of a ResNet50 step, where torch's cascaded float32 sum leaves 1e-5.
  // float32 chains of 8 pixels are folded into the double totals (one DADD per 8 elements: the double pipe is slow).
  __shared__ double s_red[2][16][64];
  const int cq = threadIdx.x & 15, ps = threadIdx.x >> 4;
  const int c = blockIdx.x * 64 + cq * 4;
  const int b = blockIdx.y, S = gridDim.z;
  const int per = (HW + S - 1) / S;
  const int p_lo = blockIdx.z * per, p_hi = min(HW, p_lo + per);
  double a0[4] = {0.0, 0.0, 0.0, 0.0}, a1[4] = {0.0, 0.0, 0.0, 0.0};
  if (c < C) {
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f), is = m;
    if (mode == 1) {
      const int e = b / ipe;
      m = dktb_ld4(mean + (long)e * C + c);
      is = dktb_ld4(invstd + (long)e * C + c);
    }
    const long xb = ((lay & kLayX) ? (long)b * gm.HWp : (long)b * HW) * C + c;
    const long gb = ((lay & kLayGy) ? (long)b * gm.HWp : (long)b * HW) * C + c;
    const long yb = ((lay & kLayY) ? (long)b * gm.HWp : (long)b * HW) * C + c;
    int hh = 0, ww = p_lo + ps;                    // (row, column) of the next pixel, advanced incrementally
    if (lay) {
      hh = ww / gm.W;
      ww -= hh * gm.W;
    }
    for (int p0 = p_lo + ps; p0 < p_hi; p0 += 16 * 8) {
      float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f), f1 = f0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {             // two batches of four pixels: every load of a batch is issued first
        float4 xv[4], gv[4], yv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int p = p0 + 16 * (4 * h + u);
          const bool in = p < p_hi;
          const int pd = in ? p : p_lo;
          int pp = pd;                                      // pixel index inside a padded image
          if (lay) {
            pp = in ? (hh + 1) * (gm.W + 2) + ww + 1 : gm.W + 3;
            ww += 16;
            while (ww >= gm.W) { ww -= gm.W; ++hh; }
          }
          xv[u] = dktb_ld4(x + xb + (long)((lay & kLayX) ? pp : pd) * C);
          if (mode == 1) {
            gv[u] = dktb_ld4(g + gb + (long)((lay & kLayGy) ? pp : pd) * C);
            if (y != nullptr) yv[u] = dktb_ld4(y + yb + (long)((lay & kLayY) ? pp : pd) * C);
          }
          if (!in) {
            xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (mode == 1) gv[u] = xv[u];
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (mode == 0) {
            f0.x += xv[u].x; f0.y += xv[u].y; f0.z += xv[u].z; f0.w += xv[u].w;
            f1.x = fmaf(xv[u].x, xv[u].x, f1.x); f1.y = fmaf(xv[u].y, xv[u].y, f1.y);
            f1.z = fmaf(xv[u].z, xv[u].z, f1.z); f1.w = fmaf(xv[u].w, xv[u].w, f1.w);
          } else {
            float4 q = gv[u];
            if (y != nullptr) {
              if (!(yv[u].x > 0.f)) q.x = 0.f;
              if (!(yv[u].y > 0.f)) q.y = 0.f;
              if (!(yv[u].z > 0.f)) q.z = 0.f;
              if (!(yv[u].w > 0.f)) q.w = 0.f;
            }
            f0.x += q.x; f0.y += q.y; f0.z += q.z; f0.w += q.w;
            f1.x = fmaf(q.x, (xv[u].x - m.x) * is.x, f1.x); f1.y = fmaf(q.y, (xv[u].y - m.y) * is.y, f1.y);
            f1.z = fmaf(q.z, (xv[u].z - m.z) * is.z, f1.z); f1.w = fmaf(q.w, (xv[u].w - m.w) * is.w, f1.w);
          }
        }
      }
      a0[0] += (double)f0.x; a0[1] += (double)f0.y; a0[2] += (double)f0.z; a0[3] += (double)f0.w;
      a1[0] += (double)f1.x; a1[1] += (double)f1.y; a1[2] += (double)f1.z; a1[3] += (double)f1.w;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    s_red[0][ps][cq * 4 + j] = a0[j];
    s_red[1][ps][cq * 4 + j] = a1[j];
  }
  __syncthreads();
  if (threadIdx.x < 128) {                     // fixed-order sum over the 16 pixel slices
    const int st = threadIdx.x >> 6, cc = threadIdx.x & 63;
    if (blockIdx.x * 64 + cc < C) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < 16; ++k) t += s_red[st][k][cc];
      partial[(((long)b * S + blockIdx.z) * C + blockIdx.x * 64 + cc) * 2 + st] = (float)t;
    }
  }
}

// sum of an episode's `rows` partial rows (images x pixel splits) for 32 channels: 8 slices of the block take every 8th row (loads of different
// slices are in flight together), then a fixed-order sum over the slices (deterministic).  Call with all 256 threads.
__device__ __forceinline__ void bn2d_episode_sums(const float* __restrict__ partial, int e, int rows, int C, int c,
                                                  double (&s_red)[2][8][32], double& s, double& q) {
  const int cx = threadIdx.x & 31, sl = threadIdx.x >> 5;
  double a0 = 0.0, a1 = 0.0;
  if (c < C) {
    const float2* pp = reinterpret_cast<const float2*>(partial) + ((long)e * rows) * C + c;
    for (int i = sl; i < rows; i += 8) {
      const float2 v = pp[(long)i * C];
      a0 += (double)v.x;
      a1 += (double)v.y;
    }
  }
  __syncthreads();                       // the previous episode's sums have been read
  s_red[0][sl][cx] = a0;
  s_red[1][sl][cx] = a1;
  __syncthreads();
  s = q = 0.0;
  if (sl == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s += s_red[0][k][cx];
      q += s_red[1][k][cx];
    }
  }
}

// forward statistics: mean / invstd per (episode, channel) + sequential running-stat EMA.  Block = 32 channels x 8 slices.
__global__ void __launch_bounds__(256) bn2d_finalize_kernel(const float* __restrict__ partial, float* __restrict__ mean,
                                                            float* __restrict__ invstd, float* __restrict__ running_mean,
                                                            float* __restrict__ running_var, int E, int ipe, int rows,
                                                            int HW, int C, float momentum, float eps) {
  __shared__ double s_red[2][8][32];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const bool lead = (threadIdx.x >> 5) == 0 && c < C;
  const double n = (double)ipe * HW;
  float rm = (lead && running_mean) ? running_mean[c] : 0.f, rv = (lead && running_var) ? running_var[c] : 0.f;
  for (int e = 0; e < E; ++e) {
    double s, q;
    bn2d_episode_sums(partial, e, rows, C, c, s_red, s, q);
    if (lead) {
      const double m = s / n;
      double var = q / n - m * m;
      if (var < 0.0) var = 0.0;
      mean[(long)e * C + c] = (float)m;
      invstd[(long)e * C + c] = (float)(1.0 / sqrt(var + (double)eps));
      rm = (1.f - momentum) * rm + momentum * (float)m;
      rv = (1.f - momentum) * rv + momentum * (float)(n > 1.0 ? var * n / (n - 1.0) : var);
    }
  }
  if (lead && running_mean) running_mean[c] = rm;
  if (lead && running_var) running_var[c] = rv;
}

// backward sums: sums[e][c][2] = sum over the episode's images; dgamma / dbeta = sum over episodes
__global__ void __launch_bounds__(256) bn2d_bwd_finalize_kernel(const float* __restrict__ partial, float* __restrict__ sums,
                                                                float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                int E, int rows, int C) {
  __shared__ double s_red[2][8][32];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const bool lead = (threadIdx.x >> 5) == 0 && c < C;
  double tg = 0.0, tb = 0.0;
  for (int e = 0; e < E; ++e) {
    double s, q;
    bn2d_episode_sums(partial, e, rows, C, c, s_red, s, q);
    if (lead) {
      sums[((long)e * C + c) * 2 + 0] = (float)s;
      sums[((long)e * C + c) * 2 + 1] = (float)q;
      tb += s;
      tg += q;
    }
  }
  if (lead) {
    dgamma[c] = (float)tg;
    dbeta[c] = (float)tb;
  }
}

static NhwcGeom nhwc_geom(int HW, int W) {
  NhwcGeom gm;
  gm.HW = HW;
  gm.W = W > 0 ? W : HW;
  gm.HWp = (HW / gm.W + 2) * (gm.W + 2);
  return gm;
}

// W, lay: see "tensor layouts" above (lay = 0: every tensor dense, W unused)
DKTB_EXPORT int dktb_bn2d_stats_l(const float* x, float* mean, float* invstd, float* running_mean, float* running_var,
                                  float* partial, int B, int HW, int C, int ipe, float momentum, float eps, int W, int lay,
                                  cudaStream_t stream) {
  DKTB_CHECK_ARG(x && mean && invstd && partial && B > 0 && HW > 0 && C > 0 && ipe > 0 && B % ipe == 0 && B <= 65535);
  DKTB_CHECK_ARG(C % 4 == 0 && (lay == 0 || (W > 0 && HW % W == 0)));
  const int S = bn2d_splits(B, HW, C);
  DKTB_LAUNCH(bn2d_partial_kernel, dim3((C + 63) / 64, B, S), dim3(256), 0, stream, x, (const float*)nullptr,
              (const float*)nullptr, (const float*)nullptr, (const float*)nullptr, partial, HW, C, ipe, 0, nhwc_geom(HW, W),
              lay);
  DKTB_LAUNCH(bn2d_finalize_kernel, dim3((C + 31) / 32), dim3(256), 0, stream, (const float*)partial, mean, invstd,
              running_mean, running_var, B / ipe, ipe, ipe * S, HW, C, momentum, eps);
  return dktb_launch_status();
}
DKTB_EXPORT int dktb_bn2d_stats(const float* x, float* mean, float* invstd, float* running_mean, float* running_var,
                                float* partial, int B, int HW, int C, int ipe, float momentum, float eps,
                                cudaStream_t stream) {
  return dktb_bn2d_stats_l(x, mean, invstd, running_mean, running_var, partial, B, HW, C, ipe, momentum, eps, 0, 0, stream);
}

// y = act( (x - mean) * invstd * gamma + beta (+ res) );  stats row = img / ipe (ipe == 0: row 0, eval mode)
__global__ void __launch_bounds__(256) bn2d_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                         const float* __restrict__ invstd,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         const float* __restrict__ res, float* __restrict__ y,
                                                         unsigned per_img4, int C, int ipe, int relu, NhwcGeom gm, int lay) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;       // float4 index inside image blockIdx.y
  if (i >= per_img4) return;
  const int b = blockIdx.y;
  const unsigned c4n = (unsigned)C >> 2;
  const unsigned p = i / c4n;
  const int c4 = (int)(i - p * c4n) * 4;
  const int e = ipe > 0 ? b / ipe : 0;
  const long drow = dense_row(b, p, gm), prow = lay ? padded_row(b, p, gm) : drow;
  const float4 v = dktb_ld4(x + ((lay & kLayX) ? prow : drow) * C + c4);
  const float4 m = dktb_ld4(mean + (long)e * C + c4), is = dktb_ld4(invstd + (long)e * C + c4);
  const float4 g = dktb_ld4(gamma + c4), bt = dktb_ld4(beta + c4);
  float4 o;
  o.x = fmaf(v.x - m.x, is.x * g.x, bt.x);
  o.y = fmaf(v.y - m.y, is.y * g.y, bt.y);
  o.z = fmaf(v.z - m.z, is.z * g.z, bt.z);
  o.w = fmaf(v.w - m.w, is.w * g.w, bt.w);
  if (res != nullptr) {
    const float4 r = dktb_ld4(res + ((lay & kLayRes) ? prow : drow) * C + c4);
    o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
  }
  if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
  dktb_st4(y + ((lay & kLayY) ? prow : drow) * C + c4, o);
}

DKTB_EXPORT int dktb_bn2d_apply_l(const float* x, const float* mean, const float* invstd, const float* gamma,
                                  const float* beta, const float* res, float* y, int B, int HW, int C, int ipe, int relu,
                                  int W, int lay, cudaStream_t stream) {
  DKTB_CHECK_ARG(x && mean && invstd && gamma && beta && y && B > 0 && HW > 0 && C > 0 && C % 4 == 0);
  DKTB_CHECK_ARG(lay == 0 || (W > 0 && HW % W == 0));
  DKTB_CHECK_ARG(B <= 65535 && (long)HW * C / 4 < 2147483647L);
  const unsigned per_img4 = (unsigned)((long)HW * C / 4);
  DKTB_LAUNCH(bn2d_apply_kernel, dim3((per_img4 + 255) / 256, B), dim3(256), 0, stream, x, mean, invstd, gamma, beta, res, y,
              per_img4, C, ipe, relu, nhwc_geom(HW, W), lay);
  return dktb_launch_status();
}
DKTB_EXPORT int dktb_bn2d_apply(const float* x, const float* mean, const float* invstd, const float* gamma,
                                const float* beta, const float* res, float* y, int B, int HW, int C, int ipe, int relu,
                                cudaStream_t stream) {
  return dktb_bn2d_apply_l(x, mean, invstd, gamma, beta, res, y, B, HW, C, ipe, relu, 0, 0, stream);
}

// gx = gamma*invstd*(g' - s1/n - xhat*s2/n),  g' = gy * (y > 0) when relu;  gres (nullable) = g'
__global__ void __launch_bounds__(256) bn2d_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                             const float* __restrict__ gy, const float* __restrict__ mean,
                                                             const float* __restrict__ invstd,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ sums, float* __restrict__ gx,
                                                             float* __restrict__ gres, unsigned per_img4, int C, int ipe,
                                                             int relu, float inv_n, NhwcGeom gm, int lay) {
  // four channels per thread (C % 4 == 0): 16-byte loads / stores; one grid row per image, 32-bit index arithmetic
  const unsigned i4 = blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 >= per_img4) return;
  const int b = blockIdx.y;
  const unsigned c4n = (unsigned)C >> 2;
  const unsigned p = i4 / c4n;
  const int c = (int)(i4 - p * c4n) * 4;
  const int e = b / ipe;
  const long pix = dense_row(b, p, gm), prow = lay ? padded_row(b, p, gm) : pix;
  float4 g = dktb_ld4(gy + ((lay & kLayGy) ? prow : pix) * C + c);
  if (relu) {
    const float4 yy = dktb_ld4(y + ((lay & kLayY) ? prow : pix) * C + c);
    if (!(yy.x > 0.f)) g.x = 0.f;
    if (!(yy.y > 0.f)) g.y = 0.f;
    if (!(yy.z > 0.f)) g.z = 0.f;
    if (!(yy.w > 0.f)) g.w = 0.f;
  }
  const float4 xv = dktb_ld4(x + ((lay & kLayX) ? prow : pix) * C + c);
  const float4 m = dktb_ld4(mean + (long)e * C + c), is = dktb_ld4(invstd + (long)e * C + c), gw = dktb_ld4(gamma + c);
  const float* sp = sums + ((long)e * C + c) * 2;
  const float4 sa = dktb_ld4(sp), sb = dktb_ld4(sp + 4);      // (s1, s2) pairs of channels c, c+1 | c+2, c+3
  float4 o;
  o.x = gw.x * is.x * (g.x - sa.x * inv_n - (xv.x - m.x) * is.x * (sa.y * inv_n));
  o.y = gw.y * is.y * (g.y - sa.z * inv_n - (xv.y - m.y) * is.y * (sa.w * inv_n));
  o.z = gw.z * is.z * (g.z - sb.x * inv_n - (xv.z - m.z) * is.z * (sb.y * inv_n));
  o.w = gw.w * is.w * (g.w - sb.z * inv_n - (xv.w - m.w) * is.w * (sb.w * inv_n));
  dktb_st4(gx + ((lay & kLayGx) ? prow : pix) * C + c, o);
  if (gres != nullptr) dktb_st4(gres + ((lay & kLayRes) ? prow : pix) * C + c, g);
}

// x: BN input; y: block output after add/ReLU (for the ReLU mask; nullable when relu == 0); gy: gradient w.r.t. y.
// partial: dktb_bn2d_partial_floats(B, HW, C) floats, sums: (B/ipe)*C*2 floats.
DKTB_EXPORT int dktb_bn2d_bwd_l(const float* x, const float* y, const float* gy, const float* mean, const float* invstd,
                                const float* gamma, float* gx, float* gres, float* dgamma, float* dbeta, float* partial,
                                float* sums, int B, int HW, int C, int ipe, int relu, int W, int lay, cudaStream_t stream) {
  DKTB_CHECK_ARG(x && gy && mean && invstd && gamma && gx && dgamma && dbeta && partial && sums && (!relu || y));
  DKTB_CHECK_ARG(B > 0 && ipe > 0 && B % ipe == 0 && B <= 65535);
  DKTB_CHECK_ARG(C % 4 == 0 && (lay == 0 || (W > 0 && HW % W == 0)));
  const int S = bn2d_splits(B, HW, C);
  const NhwcGeom gm = nhwc_geom(HW, W);
  DKTB_LAUNCH(bn2d_partial_kernel, dim3((C + 63) / 64, B, S), dim3(256), 0, stream, x, gy, relu ? y : (const float*)nullptr,
              mean, invstd, partial, HW, C, ipe, 1, gm, lay);
  DKTB_LAUNCH(bn2d_bwd_finalize_kernel, dim3((C + 31) / 32), dim3(256), 0, stream, (const float*)partial, sums, dgamma,
              dbeta, B / ipe, ipe * S, C);
  DKTB_CHECK_ARG((long)HW * C / 4 < 2147483647L);
  const unsigned per_img4 = (unsigned)((long)HW * C / 4);
  DKTB_LAUNCH(bn2d_bwd_apply_kernel, dim3((per_img4 + 255) / 256, B), dim3(256), 0, stream, x, y, gy, mean, invstd, gamma,
              (const float*)sums, gx, gres, per_img4, C, ipe, relu, 1.0f / ((float)ipe * (float)HW), gm, lay);
  return dktb_launch_status();
}
DKTB_EXPORT int dktb_bn2d_bwd(const float* x, const float* y, const float* gy, const float* mean, const float* invstd,
                              const float* gamma, float* gx, float* gres, float* dgamma, float* dbeta, float* partial,
                              float* sums, int B, int HW, int C, int ipe, int relu, cudaStream_t stream) {
  return dktb_bn2d_bwd_l(x, y, gy, mean, invstd, gamma, gx, gres, dgamma, dbeta, partial, sums, B, HW, C, ipe, relu, 0, 0,
                         stream);
}

// ------------------------------------------------------------------------------------------------ pooling
// MaxPool2d(3, stride 2, padding 1); idx stores the arg-max tap (0..8, first maximum in scan order)
__global__ void __launch_bounds__(256) maxpool3_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                           unsigned char* __restrict__ idx, int B, int H, int W, int C,
                                                           int Ho, int Wo) {
  // four channels per thread (C % 4 == 0)
  const long i4 = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c4n = C >> 2;
  const long total4 = (long)B * Ho * Wo * c4n;
  if (i4 >= total4) return;
  const int c = (int)(i4 % c4n) * 4;
  long t = i4 / c4n;
  const int ow = (int)(t % Wo);
  t /= Wo;
  const int oh = (int)(t % Ho);
  const int b = (int)(t / Ho);
  float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int ih = oh * 2 - 1 + r, iw = ow * 2 - 1 + s;
      if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
      const float4 v = dktb_ld4(x + (((long)b * H + ih) * W + iw) * C + c);
      if (v.x > best.x) { best.x = v.x; a0 = r * 3 + s; }
      if (v.y > best.y) { best.y = v.y; a1 = r * 3 + s; }
      if (v.z > best.z) { best.z = v.z; a2 = r * 3 + s; }
      if (v.w > best.w) { best.w = v.w; a3 = r * 3 + s; }
    }
  dktb_st4(y + i4 * 4, best);
  *reinterpret_cast<uchar4*>(idx + i4 * 4) = make_uchar4((unsigned char)a0, (unsigned char)a1, (unsigned char)a2, (unsigned char)a3);
}

__global__ void __launch_bounds__(256) maxpool3_bwd_kernel(const float* __restrict__ gy,
                                                           const unsigned char* __restrict__ idx, float* __restrict__ gx,
                                                           int B, int H, int W, int C, int Ho, int Wo) {
  const long i4 = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c4n = C >> 2;
  const long total4 = (long)B * H * W * c4n;
  if (i4 >= total4) return;
  const int c = (int)(i4 % c4n) * 4;
  long t = i4 / c4n;
  const int iw = (int)(t % W);
  t /= W;
  const int ih = (int)(t % H);
  const int b = (int)(t / H);
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int th = ih + 1 - r, tw = iw + 1 - s;
      if (th < 0 || tw < 0 || (th & 1) || (tw & 1)) continue;
      const int oh = th >> 1, ow = tw >> 1;
      if (oh >= Ho || ow >= Wo) continue;
      const long o = (((long)b * Ho + oh) * Wo + ow) * C + c;
      const uchar4 a = *reinterpret_cast<const uchar4*>(idx + o);
      const float4 v = dktb_ld4(gy + o);
      const unsigned char tap = (unsigned char)(r * 3 + s);
      if (a.x == tap) g.x += v.x;
      if (a.y == tap) g.y += v.y;
      if (a.z == tap) g.z += v.z;
      if (a.w == tap) g.w += v.w;
    }
  dktb_st4(gx + i4 * 4, g);
}

DKTB_EXPORT int dktb_maxpool3_fwd(const float* x, float* y, unsigned char* idx, int B, int H, int W, int C,
                                  cudaStream_t stream) {
  DKTB_CHECK_ARG(x && y && idx && B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0);
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long total = (long)B * Ho * Wo * (C / 4);
  DKTB_LAUNCH(maxpool3_fwd_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, x, y, idx, B, H, W, C, Ho,
              Wo);
  return dktb_launch_status();
}

DKTB_EXPORT int dktb_maxpool3_bwd(const float* gy, const unsigned char* idx, float* gx, int B, int H, int W, int C,
                                  cudaStream_t stream) {
  DKTB_CHECK_ARG(gy && idx && gx && B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0);
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long total = (long)B * H * W * (C / 4);
  DKTB_LAUNCH(maxpool3_bwd_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, gy, idx, gx, B, H, W, C,
              Ho, Wo);
  return dktb_launch_status();
}

// global average pooling over HW (AvgPool2d(7) on the 7x7 ResNet map): y[b][c] = mean_p x[b][p][c]
__global__ void avgpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int HW, int C, long total) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long b = i / C;
  float s = 0.f;
  for (int p = 0; p < HW; ++p) s += x[(b * HW + p) * C + c];
  y[i] = s / (float)HW;
}
__global__ void avgpool_bwd_kernel(const float* __restrict__ gy, float* __restrict__ gx, int HW, int C, long total) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long b = i / ((long)HW * C);
  gx[i] = gy[b * C + c] / (float)HW;
}

DKTB_EXPORT int dktb_avgpool_fwd(const float* x, float* y, int B, int HW, int C, cudaStream_t stream) {
  DKTB_CHECK_ARG(x && y && B > 0 && HW > 0 && C > 0);
  const long total = (long)B * C;
  DKTB_LAUNCH(avgpool_fwd_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, x, y, HW, C, total);
  return dktb_launch_status();
}
DKTB_EXPORT int dktb_avgpool_bwd(const float* gy, float* gx, int B, int HW, int C, cudaStream_t stream) {
  DKTB_CHECK_ARG(gy && gx && B > 0 && HW > 0 && C > 0);
  const long total = (long)B * HW * C;
  DKTB_LAUNCH(avgpool_bwd_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, gy, gx, HW, C, total);
  return dktb_launch_status();
}

// a += b
__global__ void add_inplace_kernel(float* __restrict__ a, const float* __restrict__ b, long n) {
  const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 4 <= n) {
    const float4 x = dktb_ld4(a + i), y = dktb_ld4(b + i);
    dktb_st4(a + i, make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w));
  } else {
    for (long j = i; j < n; ++j) a[j] += b[j];
  }
}
DKTB_EXPORT int dktb_add_inplace(float* a, const float* b, long n, cudaStream_t stream) {
  DKTB_CHECK_ARG(a && b && n > 0);
  DKTB_LAUNCH(add_inplace_kernel, dim3((unsigned)((n / 4 + 256) / 256)), dim3(256), 0, stream, a, b, n);
  return dktb_launch_status();
}

// a += b over NHWC tensors of either layout (lay: kLayX = a padded, kLayY = b padded); C % 4 == 0
__global__ void __launch_bounds__(256) add_inplace_l_kernel(float* __restrict__ a, const float* __restrict__ b,
                                                            unsigned per_img4, int C, NhwcGeom gm, int lay) {
  const unsigned i4 = blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 >= per_img4) return;
  const int img = blockIdx.y;
  const unsigned c4n = (unsigned)C >> 2;
  const unsigned p = i4 / c4n;
  const int c = (int)(i4 - p * c4n) * 4;
  const long drow = dense_row(img, p, gm), prow = padded_row(img, p, gm);
  float* pa = a + ((lay & kLayX) ? prow : drow) * C + c;
  const float4 x = dktb_ld4(pa), y = dktb_ld4(b + ((lay & kLayY) ? prow : drow) * C + c);
  dktb_st4(pa, make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w));
}
DKTB_EXPORT int dktb_add_inplace_l(float* a, const float* b, int B, int HW, int C, int W, int lay, cudaStream_t stream) {
  DKTB_CHECK_ARG(a && b && B > 0 && HW > 0 && C > 0 && C % 4 == 0 && W > 0 && HW % W == 0);
  if (lay == 0) return dktb_add_inplace(a, b, (long)B * HW * C, stream);
  DKTB_CHECK_ARG(B <= 65535 && (long)HW * C / 4 < 2147483647L);
  const unsigned per_img4 = (unsigned)((long)HW * C / 4);
  DKTB_LAUNCH(add_inplace_l_kernel, dim3((per_img4 + 255) / 256, B), dim3(256), 0, stream, a, b, per_img4, C,
              nhwc_geom(HW, W), lay);
  return dktb_launch_status();
}
