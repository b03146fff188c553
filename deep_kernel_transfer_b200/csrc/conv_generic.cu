// Generic NHWC fp32 convolution (any Cin/Cout, RxS filter, stride, padding, dilation), forward / dgrad / wgrad, as
// shared-memory-tiled implicit GEMMs on the CUDA cores.  Used by the backbones whose layers are not the 64->64 3x3
// stride-1 case covered by conv_tc.cu: Conv3 of the regression path (reference backbone.py:379-402: three
// 3x3 stride-2 dilation-2 un-padded convolutions + ReLU) and, later, the ResNet stems / strided / 1x1 layers.
//   forward : out[n,oh,ow,co] = act( bias[co] + sum_{r,s,ci} x[n, oh*st-pad+r*dil, ow*st-pad+s*dil, ci] * w[co,ci,r,s] )
//   dgrad   : gx[n,ih,iw,ci]  = sum_{r,s,co : (ih+pad-r*dil) % st == 0} gy[n,oh,ow,co] * w[co,ci,r,s]
//   wgrad   : dw[co,ci,r,s]   = sum_{n,oh,ow} x[n,ih,iw,ci] * gy[n,oh,ow,co]     (split over pixels, fixed-order reduce)
// Weights stay in the reference layout [Cout][Cin][R][S].  When relu != 0 the forward applies ReLU and the backward
// entry points take the forward OUTPUT to mask gy (ReLU backward fused into the operand load).
#include "dktb_common.cuh"

struct ConvGeo {
  int N, H, W, Cin, Ho, Wo, Cout, R, S, stride, pad, dil;
};

#define CG_TM 64   // output pixels per CTA
#define CG_TN 64   // output channels per CTA
#define CG_TK 16   // reduction chunk

// ---------------------------------------------------------------------------------------------------- forward
__global__ void __launch_bounds__(256) conv2d_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ out,
                                                         ConvGeo g, int relu) {
  __shared__ __align__(16) float s_a[CG_TK][CG_TM + 4];
  __shared__ __align__(16) float s_b[CG_TK][CG_TN + 4];
  const int tid = threadIdx.x;
  const long npix = (long)g.N * g.Ho * g.Wo;
  const long p0 = (long)blockIdx.x * CG_TM;
  const int co0 = blockIdx.y * CG_TN;
  const int ty = tid / 16, tx = tid % 16;          // 4 pixels x 4 channels per thread
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // loader roles
  const int la_p = tid % CG_TM, la_k = tid / CG_TM;       // A: 64 pixels x 4 k per pass (4 passes)
  const int lb_n = tid % CG_TN, lb_k = tid / CG_TN;       // B: 64 channels x 4 k per pass
  const long pa = p0 + la_p;
  int an = 0, aoh = 0, aow = 0;
  const bool pa_ok = pa < npix;
  if (pa_ok) {
    an = (int)(pa / ((long)g.Ho * g.Wo));
    const int rem = (int)(pa % ((long)g.Ho * g.Wo));
    aoh = rem / g.Wo;
    aow = rem % g.Wo;
  }
  for (int r = 0; r < g.R; ++r)
    for (int s = 0; s < g.S; ++s) {
      const int ih = aoh * g.stride - g.pad + r * g.dil, iw = aow * g.stride - g.pad + s * g.dil;
      const bool in_ok = pa_ok && ih >= 0 && ih < g.H && iw >= 0 && iw < g.W;
      const float* xrow = x + (((long)an * g.H + ih) * g.W + iw) * g.Cin;
      for (int c0 = 0; c0 < g.Cin; c0 += CG_TK) {
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int k = la_k + 4 * q, ci = c0 + k;
          s_a[k][la_p] = (in_ok && ci < g.Cin) ? xrow[ci] : 0.f;
          const int kb = lb_k + 4 * q, cib = c0 + kb, co = co0 + lb_n;
          s_b[kb][lb_n] = (cib < g.Cin && co < g.Cout) ? w[(((long)co * g.Cin + cib) * g.R + r) * g.S + s] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < CG_TK; ++k) {
          const float4 a = dktb_ld4(&s_a[k][ty * 4]);
          const float4 b = dktb_ld4(&s_b[k][tx * 4]);
          acc[0][0] = fmaf(a.x, b.x, acc[0][0]); acc[0][1] = fmaf(a.x, b.y, acc[0][1]);
          acc[0][2] = fmaf(a.x, b.z, acc[0][2]); acc[0][3] = fmaf(a.x, b.w, acc[0][3]);
          acc[1][0] = fmaf(a.y, b.x, acc[1][0]); acc[1][1] = fmaf(a.y, b.y, acc[1][1]);
          acc[1][2] = fmaf(a.y, b.z, acc[1][2]); acc[1][3] = fmaf(a.y, b.w, acc[1][3]);
          acc[2][0] = fmaf(a.z, b.x, acc[2][0]); acc[2][1] = fmaf(a.z, b.y, acc[2][1]);
          acc[2][2] = fmaf(a.z, b.z, acc[2][2]); acc[2][3] = fmaf(a.z, b.w, acc[2][3]);
          acc[3][0] = fmaf(a.w, b.x, acc[3][0]); acc[3][1] = fmaf(a.w, b.y, acc[3][1]);
          acc[3][2] = fmaf(a.w, b.z, acc[3][2]); acc[3][3] = fmaf(a.w, b.w, acc[3][3]);
        }
      }
    }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long p = p0 + ty * 4 + i;
    if (p >= npix) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      if (co >= g.Cout) continue;
      float v = acc[i][j] + (bias ? bias[co] : 0.f);
      if (relu) v = fmaxf(v, 0.f);
      out[p * g.Cout + co] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------------- dgrad
// gx[p_in][ci]: M = input pixels, N = Cin, K = (r,s,co)
__global__ void __launch_bounds__(256) conv2d_dgrad_kernel(const float* __restrict__ gy, const float* __restrict__ yout,
                                                           const float* __restrict__ w, float* __restrict__ gx,
                                                           ConvGeo g, int relu) {
  __shared__ __align__(16) float s_a[CG_TK][CG_TM + 4];
  __shared__ __align__(16) float s_b[CG_TK][CG_TN + 4];
  const int tid = threadIdx.x;
  const long npix = (long)g.N * g.H * g.W;
  const long p0 = (long)blockIdx.x * CG_TM;
  const int ci0 = blockIdx.y * CG_TN;
  const int ty = tid / 16, tx = tid % 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int la_p = tid % CG_TM, la_k = tid / CG_TM;
  const int lb_n = tid % CG_TN, lb_k = tid / CG_TN;
  const long pa = p0 + la_p;
  int an = 0, aih = 0, aiw = 0;
  const bool pa_ok = pa < npix;
  if (pa_ok) {
    an = (int)(pa / ((long)g.H * g.W));
    const int rem = (int)(pa % ((long)g.H * g.W));
    aih = rem / g.W;
    aiw = rem % g.W;
  }
  for (int r = 0; r < g.R; ++r)
    for (int s = 0; s < g.S; ++s) {
      const int th = aih + g.pad - r * g.dil, tw = aiw + g.pad - s * g.dil;
      bool ok = pa_ok && th >= 0 && tw >= 0 && (th % g.stride) == 0 && (tw % g.stride) == 0;
      const int oh = th / g.stride, ow = tw / g.stride;
      ok = ok && oh < g.Ho && ow < g.Wo;
      const long orow = (((long)an * g.Ho + oh) * g.Wo + ow) * g.Cout;
      for (int c0 = 0; c0 < g.Cout; c0 += CG_TK) {
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int k = la_k + 4 * q, co = c0 + k;
          float v = 0.f;
          if (ok && co < g.Cout) {
            v = gy[orow + co];
            if (relu && !(yout[orow + co] > 0.f)) v = 0.f;
          }
          s_a[k][la_p] = v;
          const int kb = lb_k + 4 * q, cob = c0 + kb, ci = ci0 + lb_n;
          s_b[kb][lb_n] = (cob < g.Cout && ci < g.Cin) ? w[(((long)cob * g.Cin + ci) * g.R + r) * g.S + s] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < CG_TK; ++k) {
          const float4 a = dktb_ld4(&s_a[k][ty * 4]);
          const float4 b = dktb_ld4(&s_b[k][tx * 4]);
          acc[0][0] = fmaf(a.x, b.x, acc[0][0]); acc[0][1] = fmaf(a.x, b.y, acc[0][1]);
          acc[0][2] = fmaf(a.x, b.z, acc[0][2]); acc[0][3] = fmaf(a.x, b.w, acc[0][3]);
          acc[1][0] = fmaf(a.y, b.x, acc[1][0]); acc[1][1] = fmaf(a.y, b.y, acc[1][1]);
          acc[1][2] = fmaf(a.y, b.z, acc[1][2]); acc[1][3] = fmaf(a.y, b.w, acc[1][3]);
          acc[2][0] = fmaf(a.z, b.x, acc[2][0]); acc[2][1] = fmaf(a.z, b.y, acc[2][1]);
          acc[2][2] = fmaf(a.z, b.z, acc[2][2]); acc[2][3] = fmaf(a.z, b.w, acc[2][3]);
          acc[3][0] = fmaf(a.w, b.x, acc[3][0]); acc[3][1] = fmaf(a.w, b.y, acc[3][1]);
          acc[3][2] = fmaf(a.w, b.z, acc[3][2]); acc[3][3] = fmaf(a.w, b.w, acc[3][3]);
        }
      }
    }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long p = p0 + ty * 4 + i;
    if (p >= npix) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + tx * 4 + j;
      if (ci < g.Cin) gx[p * g.Cin + ci] = acc[i][j];
    }
  }
}

// ---------------------------------------------------------------------------------------------------- wgrad
// For one (r,s): dw_rs[ci][co] = sum_p x[p@(r,s)][ci] * gy[p][co]; M = Cin tile, N = Cout tile, K = output pixels.
// grid (ci tiles, co tiles, R*S*nsplit); partial [nsplit][R*S][Cin][Cout] then a fixed-order reduce into [co][ci][r][s].
__global__ void __launch_bounds__(256) conv2d_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                                           const float* __restrict__ yout, float* __restrict__ partial,
                                                           ConvGeo g, int relu, int nsplit) {
  __shared__ __align__(16) float s_a[CG_TK][CG_TM + 4];
  __shared__ __align__(16) float s_b[CG_TK][CG_TN + 4];
  const int tid = threadIdx.x;
  const int ci0 = blockIdx.x * CG_TM, co0 = blockIdx.y * CG_TN;
  const int rs = blockIdx.z / nsplit, split = blockIdx.z % nsplit;
  const int r = rs / g.S, s = rs % g.S;
  const long npix = (long)g.N * g.Ho * g.Wo;
  const long per = (npix + nsplit - 1) / nsplit;
  const long pbeg = split * per, pend = (pbeg + per < npix) ? pbeg + per : npix;
  const int ty = tid / 16, tx = tid % 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int l_c = tid % 64, l_k = tid / 64;            // loaders: 64 channels x 4 pixels per pass
  for (long pc = pbeg; pc < pend; pc += CG_TK) {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = l_k + 4 * q;
      const long p = pc + k;
      float va = 0.f, vb = 0.f;
      if (p < pend) {
        const int n = (int)(p / ((long)g.Ho * g.Wo));
        const int rem = (int)(p % ((long)g.Ho * g.Wo));
        const int oh = rem / g.Wo, ow = rem % g.Wo;
        const int ih = oh * g.stride - g.pad + r * g.dil, iw = ow * g.stride - g.pad + s * g.dil;
        const int ci = ci0 + l_c, co = co0 + l_c;
        if (ci < g.Cin && ih >= 0 && ih < g.H && iw >= 0 && iw < g.W)
          va = x[(((long)n * g.H + ih) * g.W + iw) * g.Cin + ci];
        if (co < g.Cout) {
          vb = gy[p * g.Cout + co];
          if (relu && !(yout[p * g.Cout + co] > 0.f)) vb = 0.f;
        }
      }
      s_a[k][l_c] = va;
      s_b[k][l_c] = vb;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CG_TK; ++k) {
      const float4 a = dktb_ld4(&s_a[k][ty * 4]);
      const float4 b = dktb_ld4(&s_b[k][tx * 4]);
      acc[0][0] = fmaf(a.x, b.x, acc[0][0]); acc[0][1] = fmaf(a.x, b.y, acc[0][1]);
      acc[0][2] = fmaf(a.x, b.z, acc[0][2]); acc[0][3] = fmaf(a.x, b.w, acc[0][3]);
      acc[1][0] = fmaf(a.y, b.x, acc[1][0]); acc[1][1] = fmaf(a.y, b.y, acc[1][1]);
      acc[1][2] = fmaf(a.y, b.z, acc[1][2]); acc[1][3] = fmaf(a.y, b.w, acc[1][3]);
      acc[2][0] = fmaf(a.z, b.x, acc[2][0]); acc[2][1] = fmaf(a.z, b.y, acc[2][1]);
      acc[2][2] = fmaf(a.z, b.z, acc[2][2]); acc[2][3] = fmaf(a.z, b.w, acc[2][3]);
      acc[3][0] = fmaf(a.w, b.x, acc[3][0]); acc[3][1] = fmaf(a.w, b.y, acc[3][1]);
      acc[3][2] = fmaf(a.w, b.z, acc[3][2]); acc[3][3] = fmaf(a.w, b.w, acc[3][3]);
    }
  }
  float* o = partial + ((long)split * g.R * g.S + rs) * g.Cin * g.Cout;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ci = ci0 + ty * 4 + i;
    if (ci >= g.Cin) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      if (co < g.Cout) o[(long)ci * g.Cout + co] = acc[i][j];
    }
  }
}

__global__ void conv2d_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, ConvGeo g,
                                           int nsplit) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)g.R * g.S * g.Cin * g.Cout;
  if (i >= total) return;
  float t = 0.f;
  for (int sp = 0; sp < nsplit; ++sp) t += partial[(long)sp * total + i];
  const int co = (int)(i % g.Cout);
  const int ci = (int)((i / g.Cout) % g.Cin);
  const int rs = (int)(i / ((long)g.Cout * g.Cin));
  dw[((long)co * g.Cin + ci) * g.R * g.S + rs] = t;
}

// db[co] = sum_p gy[p][co] (masked by ReLU); one CTA per 64 channels, fixed order over pixel slices
__global__ void __launch_bounds__(256) conv2d_bgrad_kernel(const float* __restrict__ gy, const float* __restrict__ yout,
                                                           float* __restrict__ db, long npix, int Cout, int relu) {
  __shared__ float s_red[4][64];
  const int c = blockIdx.x * 64 + threadIdx.x % 64, sl = threadIdx.x / 64;
  float t = 0.f;
  if (c < Cout)
    for (long p = sl; p < npix; p += 4) {
      float v = gy[p * Cout + c];
      if (relu && !(yout[p * Cout + c] > 0.f)) v = 0.f;
      t += v;
    }
  s_red[sl][threadIdx.x % 64] = t;
  __syncthreads();
  if (sl == 0 && c < Cout) db[c] = (s_red[0][threadIdx.x] + s_red[1][threadIdx.x]) + (s_red[2][threadIdx.x] + s_red[3][threadIdx.x]);
}

static inline ConvGeo make_geo(int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad, int dil) {
  ConvGeo g;
  g.N = N; g.H = H; g.W = W; g.Cin = Cin; g.Cout = Cout; g.R = R; g.S = S; g.stride = stride; g.pad = pad; g.dil = dil;
  g.Ho = (H + 2 * pad - dil * (R - 1) - 1) / stride + 1;
  g.Wo = (W + 2 * pad - dil * (S - 1) - 1) / stride + 1;
  return g;
}

DKTB_EXPORT int dktb_conv2d_out_size(int H, int R, int stride, int pad, int dil) {
  return (H + 2 * pad - dil * (R - 1) - 1) / stride + 1;
}

DKTB_EXPORT int dktb_conv2d_fwd(const float* x, const float* w, const float* bias, float* out, int N, int H, int W,
                                int Cin, int Cout, int R, int S, int stride, int pad, int dil, int relu,
                                cudaStream_t stream) {
  DKTB_CHECK_ARG(x && w && out && N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && R > 0 && S > 0 && stride > 0);
  const ConvGeo g = make_geo(N, H, W, Cin, Cout, R, S, stride, pad, dil);
  DKTB_CHECK_ARG(g.Ho > 0 && g.Wo > 0);
  const long npix = (long)N * g.Ho * g.Wo;
  DKTB_LAUNCH(conv2d_fwd_kernel, dim3((unsigned)((npix + CG_TM - 1) / CG_TM), (Cout + CG_TN - 1) / CG_TN), dim3(256), 0,
              stream, x, w, bias, out, g, relu);
  return dktb_launch_status();
}

// gy / yout [N,Ho,Wo,Cout]; yout (the forward output) is only read when relu != 0
DKTB_EXPORT int dktb_conv2d_dgrad(const float* gy, const float* yout, const float* w, float* gx, int N, int H, int W,
                                  int Cin, int Cout, int R, int S, int stride, int pad, int dil, int relu,
                                  cudaStream_t stream) {
  DKTB_CHECK_ARG(gy && w && gx && (!relu || yout) && N > 0);
  const ConvGeo g = make_geo(N, H, W, Cin, Cout, R, S, stride, pad, dil);
  const long npix = (long)N * H * W;
  DKTB_LAUNCH(conv2d_dgrad_kernel, dim3((unsigned)((npix + CG_TM - 1) / CG_TM), (Cin + CG_TN - 1) / CG_TN), dim3(256),
              0, stream, gy, yout, w, gx, g, relu);
  return dktb_launch_status();
}

DKTB_EXPORT int dktb_conv2d_wgrad_nsplit(long npix) {
  long n = npix / 2048;
  return (int)(n < 1 ? 1 : (n > 64 ? 64 : n));
}

// dw [Cout,Cin,R,S], db [Cout] (nullable); scratch: nsplit*R*S*Cin*Cout floats, nsplit = dktb_conv2d_wgrad_nsplit(N*Ho*Wo)
DKTB_EXPORT int dktb_conv2d_wgrad(const float* x, const float* gy, const float* yout, float* dw, float* db,
                                  float* scratch, int N, int H, int W, int Cin, int Cout, int R, int S, int stride,
                                  int pad, int dil, int relu, cudaStream_t stream) {
  DKTB_CHECK_ARG(x && gy && dw && scratch && (!relu || yout) && N > 0);
  const ConvGeo g = make_geo(N, H, W, Cin, Cout, R, S, stride, pad, dil);
  const long npix = (long)N * g.Ho * g.Wo;
  const int nsplit = dktb_conv2d_wgrad_nsplit(npix);
  DKTB_LAUNCH(conv2d_wgrad_kernel, dim3((Cin + CG_TM - 1) / CG_TM, (Cout + CG_TN - 1) / CG_TN, R * S * nsplit),
              dim3(256), 0, stream, x, gy, yout, scratch, g, relu, nsplit);
  const long total = (long)R * S * Cin * Cout;
  DKTB_LAUNCH(conv2d_wgrad_reduce_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream,
              (const float*)scratch, dw, g, nsplit);
  if (db != nullptr)
    DKTB_LAUNCH(conv2d_bgrad_kernel, dim3((Cout + 63) / 64), dim3(256), 0, stream, gy, yout, db, npix, Cout, relu);
  return dktb_launch_status();
}

// [N,C,H,W] -> [N,H,W,C]
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ out, int N, int C, int H, int W) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)N * C * H * W;
  if (i >= total) return;
  const int c = (int)(i % C);
  long t = i / C;
  const int w = (int)(t % W);
  t /= W;
  const int h = (int)(t % H);
  const int n = (int)(t / H);
  out[i] = x[(((long)n * C + c) * H + h) * W + w];
}

DKTB_EXPORT int dktb_nchw_to_nhwc(const float* x, float* out, int N, int C, int H, int W, cudaStream_t stream) {
  DKTB_CHECK_ARG(x && out && N > 0 && C > 0 && H > 0 && W > 0);
  const long total = (long)N * C * H * W;
  DKTB_LAUNCH(nchw_to_nhwc_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, x, out, N, C, H, W);
  return dktb_launch_status();
}
