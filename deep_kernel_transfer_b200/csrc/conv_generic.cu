// Generic NHWC fp32 convolution (any Cin/Cout, RxS filter, stride, padding, dilation), forward / dgrad / wgrad, as
// shared-memory-tiled implicit GEMMs on the CUDA cores.  Used by the backbones whose layers are not the 64->64 3x3
// stride-1 case covered by conv_tc.cu: Conv3 of the regression path (reference backbone.py:379-402: three
// 3x3 stride-2 dilation-2 un-padded convolutions + ReLU) and, later, the ResNet stems / strided / 1x1 layers.
//   forward : out[n,oh,ow,co] = act( bias[co] + sum_{r,s,ci} x[n, oh*st-pad+r*dil, ow*st-pad+s*dil, ci] * w[co,ci,r,s] )
//   dgrad   : gx[n,ih,iw,ci]  = sum_{r,s,co : (ih+pad-r*dil) % st == 0} gy[n,oh,ow,co] * w[co,ci,r,s]
//   wgrad   : dw[co,ci,r,s]   = sum_{n,oh,ow} x[n,ih,iw,ci] * gy[n,oh,ow,co]     (split over pixels, fixed-order reduce)
// Weights stay in the reference layout [Cout][Cin][R][S].  When relu != 0 the forward applies ReLU and the backward
// entry points take the forward OUTPUT to mask gy (ReLU backward fused into the operand load).
#include <stdlib.h>

#include "tile_mma.cuh"

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

// db[co] = sum_p gy[p][co] (masked by ReLU): grid (64-channel groups, pixel slices) -> part[slice][co], then a
// fixed-order sum over the slices (deterministic; one CTA per channel group alone left 140 SMs idle: 2 ms per layer).
__global__ void __launch_bounds__(256) conv2d_bgrad_kernel(const float* __restrict__ gy, const float* __restrict__ yout,
                                                           float* __restrict__ part, long npix, int Cout, int relu,
                                                           int nslice) {
  __shared__ float s_red[4][64];
  const int c = blockIdx.x * 64 + threadIdx.x % 64, sl = threadIdx.x / 64;
  float t = 0.f;
  if (c < Cout)
    for (long p = (long)blockIdx.y * 4 + sl; p < npix; p += 4L * nslice) {
      float v = gy[p * Cout + c];
      if (relu && !(yout[p * Cout + c] > 0.f)) v = 0.f;
      t += v;
    }
  s_red[sl][threadIdx.x % 64] = t;
  __syncthreads();
  if (sl == 0 && c < Cout)
    part[(long)blockIdx.y * Cout + c] =
        (s_red[0][threadIdx.x] + s_red[1][threadIdx.x]) + (s_red[2][threadIdx.x] + s_red[3][threadIdx.x]);
}
__global__ void conv2d_bgrad_final_kernel(const float* __restrict__ part, float* __restrict__ db, int Cout, int nslice) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cout) return;
  float t = 0.f;
  for (int s = 0; s < nslice; ++s) t += part[(long)s * Cout + c];
  db[c] = t;
}

// ------------------------------------------------------------------------------------- tensor-core variants (ResNet)
// The same three products as 64 x 64 x k tiles on the warp-level tensor cores (tile_mma.cuh: mma.sync m16n8k8, 3xTF32
// split, fp32 accumulate, next k-chunk prefetched into registers), for layers whose reduction width (Cin forward,
// Cout dgrad) is a multiple of 32 -- every ResNet layer but the stem (reference backbone.py:135-247, 330-376).
// Operands are read k-contiguous: activations NHWC as they are, weights from two re-laid-out copies the caller refreshes
// once per step with dktb_conv2d_prep_mma:  wf [R*S][Cout][Cin] (forward),  wd [R*S][Cin][Cout] (dgrad).
__global__ void conv2d_prep_mma_kernel(const float* __restrict__ w, float* __restrict__ wf, float* __restrict__ wd,
                                       int Cout, int Cin, int RS) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)Cout * Cin * RS) return;
  const int t = (int)(i % RS);
  const int ci = (int)((i / RS) % Cin);
  const int co = (int)(i / ((long)RS * Cin));
  const float v = w[i];
  wf[((long)t * Cout + co) * Cin + ci] = v;
  wd[((long)t * Cin + ci) * Cout + co] = v;
}

// A operand of forward (MODE 0: x at the tap's input position of each output pixel) and dgrad (MODE 1: gy at the output
// position that used this input pixel through the tap, if any).  tbl[q] = (image base in pixels, h0, w0, valid).
template <int MODE>
struct CmPixels {
  const float* src; const int4* tbl; ConvGeo g; int kc;       // kc = channels per tap on the k axis
  __device__ __forceinline__ void fetch(float (&v)[8], int k0) const {
    int tap, c;
    bool kok = true;
    if (MODE == 2) {                 // narrow input: k runs over the flattened (r, s, ci) index, padded to a multiple of 32
      const int kk = k0 + (threadIdx.x & 31);
      kok = kk < g.R * g.S * g.Cin;
      tap = kk / g.Cin;
      c = kk - tap * g.Cin;
    } else {
      tap = k0 / kc;
      c = k0 - tap * kc + (threadIdx.x & 31);
    }
    const int r = tap / g.S, s = tap - r * g.S;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int4 t = tbl[(threadIdx.x >> 5) + 8 * i];
      float val = 0.f;
      if (MODE == 0 || MODE == 2) {
        const int ih = t.y + r * g.dil, iw = t.z + s * g.dil;
        if (kok && t.w && (unsigned)ih < (unsigned)g.H && (unsigned)iw < (unsigned)g.W)
          val = src[((long)t.x + (long)ih * g.W + iw) * g.Cin + c];
      } else {
        const int th = t.y - r * g.dil, tw = t.z - s * g.dil;
        if (g.stride == 1) {                                    // uniform branch: no divisions on the common layers
          if (t.w && (unsigned)th < (unsigned)g.Ho && (unsigned)tw < (unsigned)g.Wo)
            val = src[((long)t.x + (long)th * g.Wo + tw) * g.Cout + c];
        } else if (t.w && th >= 0 && tw >= 0 && th % g.stride == 0 && tw % g.stride == 0) {
          const int oh = th / g.stride, ow = tw / g.stride;
          if (oh < g.Ho && ow < g.Wo) val = src[((long)t.x + (long)oh * g.Wo + ow) * g.Cout + c];
        }
      }
      v[i] = val;
    }
  }
  __device__ __forceinline__ void store(float* sdst, const float (&v)[8]) const {
    const int k = threadIdx.x & 31, q0 = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 8; ++i) sdst[(q0 + 8 * i) * GL_LDK + k] = v[i];
  }
};
// B operand: re-laid-out weights [tap][nn][kc], rows n0 .. n0 + 63 of the tap the chunk belongs to.
struct CmWeights {
  const float* wt; int nn, n0, kc;
  __device__ __forceinline__ void fetch(float (&v)[8], int k0) const {
    const int tap = k0 / kc, c = k0 - tap * kc + (threadIdx.x & 31);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = n0 + (threadIdx.x >> 5) + 8 * i;
      v[i] = n < nn ? wt[((long)tap * nn + n) * kc + c] : 0.f;
    }
  }
  __device__ __forceinline__ void store(float* sdst, const float (&v)[8]) const {
    const int k = threadIdx.x & 31, q0 = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 8; ++i) sdst[(q0 + 8 * i) * GL_LDK + k] = v[i];
  }
};

// The same two operands fetched as 16-byte vectors (4 consecutive channels of one pixel / filter row per thread, two
// per thread and chunk instead of eight scalars): valid whenever the channel count on the k axis is a multiple of 32,
// i.e. MODE 0 and 1.  The tap of a chunk is tracked incrementally (gl_product asks for consecutive chunks).
template <int MODE>
struct CmPixels4 {
  const float* src; const int4* tbl; ConvGeo g; int kc;
  mutable int next_k0, tap, c0;
  __device__ __forceinline__ void locate(int k0) const {
    if (k0 == next_k0) {
      c0 += GL_KC;
      if (c0 == kc) { c0 = 0; ++tap; }
    } else {
      tap = k0 / kc;
      c0 = k0 - tap * kc;
    }
    next_k0 = k0 + GL_KC;
  }
  __device__ __forceinline__ void fetch(float (&v)[8], int k0) const {
    locate(k0);
    const int r = tap / g.S, s = tap - r * g.S;
    const int c = c0 + 4 * (threadIdx.x & 7);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int4 t = tbl[(threadIdx.x >> 3) + 32 * i];
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (MODE == 0) {
        const int ih = t.y + r * g.dil, iw = t.z + s * g.dil;
        if (t.w && (unsigned)ih < (unsigned)g.H && (unsigned)iw < (unsigned)g.W)
          val = dktb_ld4(src + ((long)t.x + (long)ih * g.W + iw) * g.Cin + c);
      } else {
        const int th = t.y - r * g.dil, tw = t.z - s * g.dil;
        if (g.stride == 1) {
          if (t.w && (unsigned)th < (unsigned)g.Ho && (unsigned)tw < (unsigned)g.Wo)
            val = dktb_ld4(src + ((long)t.x + (long)th * g.Wo + tw) * g.Cout + c);
        } else if (t.w && th >= 0 && tw >= 0 && th % g.stride == 0 && tw % g.stride == 0) {
          const int oh = th / g.stride, ow = tw / g.stride;
          if (oh < g.Ho && ow < g.Wo) val = dktb_ld4(src + ((long)t.x + (long)oh * g.Wo + ow) * g.Cout + c);
        }
      }
      v[4 * i] = val.x; v[4 * i + 1] = val.y; v[4 * i + 2] = val.z; v[4 * i + 3] = val.w;
    }
  }
  __device__ __forceinline__ void store(float* sdst, const float (&v)[8]) const {
#pragma unroll
    for (int i = 0; i < 2; ++i)
      dktb_st4(sdst + ((threadIdx.x >> 3) + 32 * i) * GL_LDK + 4 * (threadIdx.x & 7),
               make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
  }
};
struct CmWeights4 {
  const float* wt; int nn, n0, kc;
  mutable int next_k0, tap, c0;
  __device__ __forceinline__ void fetch(float (&v)[8], int k0) const {
    if (k0 == next_k0) {
      c0 += GL_KC;
      if (c0 == kc) { c0 = 0; ++tap; }
    } else {
      tap = k0 / kc;
      c0 = k0 - tap * kc;
    }
    next_k0 = k0 + GL_KC;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int n = n0 + (threadIdx.x >> 3) + 32 * i;
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < nn) val = dktb_ld4(wt + ((long)tap * nn + n) * kc + c0 + 4 * (threadIdx.x & 7));
      v[4 * i] = val.x; v[4 * i + 1] = val.y; v[4 * i + 2] = val.z; v[4 * i + 3] = val.w;
    }
  }
  __device__ __forceinline__ void store(float* sdst, const float (&v)[8]) const {
#pragma unroll
    for (int i = 0; i < 2; ++i)
      dktb_st4(sdst + ((threadIdx.x >> 3) + 32 * i) * GL_LDK + 4 * (threadIdx.x & 7),
               make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
  }
};

// MODE 0: out[p][co] = bias[co] + sum_{tap,ci} x[p@tap][ci] wf[tap][co][ci]     (p over output pixels)
// MODE 1: gx[p][ci]  =           sum_{tap,co} gy[p@tap][co] wd[tap][ci][co]     (p over input pixels)
template <int MODE>
__global__ void __launch_bounds__(GL_THREADS, 2) conv2d_mma_kernel(const float* __restrict__ src,
                                                                   const float* __restrict__ wt,
                                                                   const float* __restrict__ bias,
                                                                   float* __restrict__ dst, ConvGeo g) {
  __shared__ __align__(16) float s_as[2 * GL_PLANE];
  __shared__ __align__(16) float s_bs[2 * GL_PLANE];
  __shared__ int4 s_tbl[GL_T];
  typedef GlMap<true> M;
  const int tid = threadIdx.x;
  const int Hp = MODE != 1 ? g.Ho : g.H, Wp = MODE != 1 ? g.Wo : g.W;      // the pixel grid this kernel tiles
  const long npix = (long)g.N * Hp * Wp;
  const long p0 = (long)blockIdx.x * GL_T;
  const int nn = MODE != 1 ? g.Cout : g.Cin;
  const int kc = MODE == 0 ? g.Cin : (MODE == 1 ? g.Cout : (g.R * g.S * g.Cin + GL_KC - 1) / GL_KC * GL_KC);
  const int n0 = blockIdx.y * GL_T;
  if (tid < GL_T) {
    const long p = p0 + tid;
    int4 t = make_int4(0, 0, 0, 0);
    if (p < npix) {
      const int n = (int)(p / ((long)Hp * Wp));
      const int rem = (int)(p - (long)n * Hp * Wp);
      const int ph = rem / Wp, pw = rem - ph * Wp;
      if (MODE != 1) t = make_int4(n * g.H * g.W, ph * g.stride - g.pad, pw * g.stride - g.pad, 1);
      else t = make_int4(n * g.Ho * g.Wo, ph + g.pad, pw + g.pad, 1);
    }
    s_tbl[tid] = t;
  }
  __syncthreads();
  float acc[4][4];
  gl_zero(acc);
  if (MODE == 2) {
    CmPixels<MODE> la{src, s_tbl, g, kc};
    CmWeights4 lb{wt, nn, n0, kc, -1, 0, 0};              // wflat rows are kc = Kp floats: one "tap" of width Kp
    gl_product_ps(acc, s_as, s_bs, la, lb, 0, kc);
  } else {
    CmPixels4<MODE> la{src, s_tbl, g, kc, -1, 0, 0};
    CmWeights4 lb{wt, nn, n0, kc, -1, 0, 0};
    gl_product_ps(acc, s_as, s_bs, la, lb, 0, g.R * g.S * kc);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long p = p0 + M::row(i);
    if (p >= npix) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + M::col(j);
      if (n < nn) dst[p * nn + n] = acc[i][j] + ((MODE != 1 && bias != nullptr) ? bias[n] : 0.f);
    }
  }
}

// dgrad through a stride: an input pixel only meets the taps whose offset has its parity class, (ih + pad - r dil) % stride
// == 0 -- a quarter of a 3x3 filter on average at stride 2, none at all for three of the four classes of a 1x1.  So the
// input pixels are tiled class by class ((ih + pad) % stride, (iw + pad) % stride) and each CTA runs only its class's
// taps (the generic kernel multiplied zeros for the rest: 4x the work on the heaviest dgrad launches of a ResNet).
struct CmStrideClasses {
  int tile_begin[17];            // first tile of each class (stride <= 4), [nclass] = total
};
struct CmPixelsS4 {
  const float* src; const int4* tbl; const int* taps; ConvGeo g;
  mutable int next_k0, ti, c0;
  __device__ __forceinline__ void fetch(float (&v)[8], int k0) const {
    if (k0 == next_k0) {
      c0 += GL_KC;
      if (c0 == g.Cout) { c0 = 0; ++ti; }
    } else {
      ti = k0 / g.Cout;
      c0 = k0 - ti * g.Cout;
    }
    next_k0 = k0 + GL_KC;
    const int tap = taps[ti];
    const int r = tap / g.S, s = tap - r * g.S;
    const int c = c0 + 4 * (threadIdx.x & 7);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int4 t = tbl[(threadIdx.x >> 3) + 32 * i];
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      const int th = t.y - r * g.dil, tw = t.z - s * g.dil;          // multiples of the stride by construction
      if (t.w && th >= 0 && tw >= 0) {
        const int oh = th / g.stride, ow = tw / g.stride;
        if (oh < g.Ho && ow < g.Wo) val = dktb_ld4(src + ((long)t.x + (long)oh * g.Wo + ow) * g.Cout + c);
      }
      v[4 * i] = val.x; v[4 * i + 1] = val.y; v[4 * i + 2] = val.z; v[4 * i + 3] = val.w;
    }
  }
  __device__ __forceinline__ void store(float* sdst, const float (&v)[8]) const {
#pragma unroll
    for (int i = 0; i < 2; ++i)
      dktb_st4(sdst + ((threadIdx.x >> 3) + 32 * i) * GL_LDK + 4 * (threadIdx.x & 7),
               make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
  }
};
struct CmWeightsS4 {
  const float* wt; const int* taps; int nn, n0, kc;
  mutable int next_k0, ti, c0;
  __device__ __forceinline__ void fetch(float (&v)[8], int k0) const {
    if (k0 == next_k0) {
      c0 += GL_KC;
      if (c0 == kc) { c0 = 0; ++ti; }
    } else {
      ti = k0 / kc;
      c0 = k0 - ti * kc;
    }
    next_k0 = k0 + GL_KC;
    const int tap = taps[ti];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int n = n0 + (threadIdx.x >> 3) + 32 * i;
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < nn) val = dktb_ld4(wt + ((long)tap * nn + n) * kc + c0 + 4 * (threadIdx.x & 7));
      v[4 * i] = val.x; v[4 * i + 1] = val.y; v[4 * i + 2] = val.z; v[4 * i + 3] = val.w;
    }
  }
  __device__ __forceinline__ void store(float* sdst, const float (&v)[8]) const {
#pragma unroll
    for (int i = 0; i < 2; ++i)
      dktb_st4(sdst + ((threadIdx.x >> 3) + 32 * i) * GL_LDK + 4 * (threadIdx.x & 7),
               make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
  }
};

__global__ void __launch_bounds__(GL_THREADS, 2) conv2d_dgrad_strided_mma_kernel(const float* __restrict__ gy,
                                                                                 const float* __restrict__ wd,
                                                                                 float* __restrict__ gx, ConvGeo g,
                                                                                 CmStrideClasses cls) {
  __shared__ __align__(16) float s_as[2 * GL_PLANE];
  __shared__ __align__(16) float s_bs[2 * GL_PLANE];
  __shared__ int4 s_tbl[GL_T];
  __shared__ int s_out[GL_T];
  __shared__ int s_taps[64];
  __shared__ int s_ntaps;
  typedef GlMap<true> M;
  const int tid = threadIdx.x, st = g.stride;
  int c = 0;
  while (c + 1 < st * st && (int)blockIdx.x >= cls.tile_begin[c + 1]) ++c;
  const int ca = c / st, cb = c - ca * st;                     // (ih + pad) % st, (iw + pad) % st of this CTA's pixels
  const int fh = ((ca - g.pad) % st + st) % st, fw = ((cb - g.pad) % st + st) % st;     // first row / column of the class
  const int Ha = fh < g.H ? (g.H - fh + st - 1) / st : 0, Wb = fw < g.W ? (g.W - fw + st - 1) / st : 0;
  const long npix = (long)g.N * Ha * Wb;
  const long q0 = (long)((int)blockIdx.x - cls.tile_begin[c]) * GL_T;
  const int n0 = blockIdx.y * GL_T;
  if (tid < GL_T) {
    const long q = q0 + tid;
    int4 t = make_int4(0, 0, 0, 0);
    int o = -1;
    if (q < npix) {
      const int n = (int)(q / ((long)Ha * Wb));
      const int rem = (int)(q - (long)n * Ha * Wb);
      const int a = rem / Wb, b = rem - a * Wb;
      const int ih = fh + a * st, iw = fw + b * st;
      t = make_int4(n * g.Ho * g.Wo, ih + g.pad, iw + g.pad, 1);
      o = (n * g.H + ih) * g.W + iw;
    }
    s_tbl[tid] = t;
    s_out[tid] = o;
  }
  if (tid == 0) {
    int nt = 0;
    for (int r = 0; r < g.R; ++r)
      for (int s = 0; s < g.S; ++s)
        if ((r * g.dil) % st == ca && (s * g.dil) % st == cb) s_taps[nt++] = r * g.S + s;
    s_ntaps = nt;
  }
  __syncthreads();
  float acc[4][4];
  gl_zero(acc);
  {
    CmPixelsS4 la{gy, s_tbl, s_taps, g, -1, 0, 0};
    CmWeightsS4 lb{wd, s_taps, g.Cin, n0, g.Cout, -1, 0, 0};
    gl_product_ps(acc, s_as, s_bs, la, lb, 0, s_ntaps * g.Cout);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int o = s_out[M::row(i)];
    if (o < 0) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + M::col(j);
      if (n < g.Cin) gx[(long)o * g.Cin + n] = acc[i][j];
    }
  }
}

// wgrad: partial[split][tap][ci][co] = sum_{p in split} x[p@tap][ci] gy[p][co]; k = output pixels, both operands read
// with the channel along the lanes (coalesced) and the pixel coordinates advanced incrementally from chunk to chunk.
struct CmWgradX {
  const float* x; ConvGeo g; int c0, r, s; long pbeg, pend;
  mutable int n, oh, ow; mutable long pcur;                    // pixel of this thread's first element of the next fetch
  __device__ __forceinline__ void seek(long p) const {
    pcur = p;
    n = (int)(p / ((long)g.Ho * g.Wo));
    const int rem = (int)(p - (long)n * g.Ho * g.Wo);
    oh = rem / g.Wo;
    ow = rem - oh * g.Wo;
  }
  __device__ __forceinline__ void fetch(float (&v)[8], int k0) const {          // called with k0 = 0, 32, 64, ...
    const int q = threadIdx.x & 63, c = c0 + q;
    int nn = n, hh = oh, ww = ow;
    long p = pcur;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float val = 0.f;
      if (p < pend && c < g.Cin) {
        const int ih = hh * g.stride - g.pad + r * g.dil, iw = ww * g.stride - g.pad + s * g.dil;
        if ((unsigned)ih < (unsigned)g.H && (unsigned)iw < (unsigned)g.W)
          val = x[(((long)nn * g.H + ih) * g.W + iw) * g.Cin + c];
      }
      v[i] = val;
      p += 4;
      ww += 4;
      while (ww >= g.Wo) { ww -= g.Wo; if (++hh == g.Ho) { hh = 0; ++nn; } }
    }
    n = nn; oh = hh; ow = ww; pcur = p;                          // 8 elements x 4 pixels = the next chunk's first pixel
  }
  __device__ __forceinline__ void store(float* sdst, const float (&v)[8]) const {
    const int q = threadIdx.x & 63, kk = threadIdx.x >> 6;
#pragma unroll
    for (int i = 0; i < 8; ++i) sdst[q * GL_LDK + kk + 4 * i] = v[i];
  }
};
struct CmWgradG {
  const float* gy; int Cout, c0; long pbeg, pend;
  __device__ __forceinline__ void fetch(float (&v)[8], int k0) const {
    const int q = threadIdx.x & 63, c = c0 + q;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const long p = pbeg + k0 + (threadIdx.x >> 6) + 4 * i;
      v[i] = (p < pend && c < Cout) ? gy[p * Cout + c] : 0.f;
    }
  }
  __device__ __forceinline__ void store(float* sdst, const float (&v)[8]) const {
    const int q = threadIdx.x & 63, kk = threadIdx.x >> 6;
#pragma unroll
    for (int i = 0; i < 8; ++i) sdst[q * GL_LDK + kk + 4 * i] = v[i];
  }
};

// 16-byte versions (channel counts multiples of 4): a thread fetches 4 consecutive channels of one pixel, twice per
// chunk; lane = (channel-group parity, 16 pixels) makes the transposing shared-memory stores conflict-free.
struct CmWgradX4 {
  const float* x; ConvGeo g; int c0, r, s; long pbeg, pend;
  mutable int n, oh, ow; mutable long pcur;
  __device__ __forceinline__ void seek(long p) const {
    pcur = p;
    n = (int)(p / ((long)g.Ho * g.Wo));
    const int rem = (int)(p - (long)n * g.Ho * g.Wo);
    oh = rem / g.Wo;
    ow = rem - oh * g.Wo;
  }
  __device__ __forceinline__ void fetch(float (&v)[8], int k0) const {
    const int c = c0 + 4 * (2 * (threadIdx.x >> 5) + (threadIdx.x & 1));
    int nn = n, hh = oh, ww = ow;
    long p = pcur;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p < pend && c < g.Cin) {
        const int ih = hh * g.stride - g.pad + r * g.dil, iw = ww * g.stride - g.pad + s * g.dil;
        if ((unsigned)ih < (unsigned)g.H && (unsigned)iw < (unsigned)g.W)
          val = dktb_ld4(x + (((long)nn * g.H + ih) * g.W + iw) * g.Cin + c);
      }
      v[4 * i] = val.x; v[4 * i + 1] = val.y; v[4 * i + 2] = val.z; v[4 * i + 3] = val.w;
      p += 16;
      ww += 16;
      while (ww >= g.Wo) { ww -= g.Wo; if (++hh == g.Ho) { hh = 0; ++nn; } }
    }
    n = nn; oh = hh; ow = ww; pcur = p;
  }
  __device__ __forceinline__ void store(float* sdst, const float (&v)[8]) const {
    const int q = 4 * (2 * (threadIdx.x >> 5) + (threadIdx.x & 1)), kk = (threadIdx.x & 31) >> 1;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) sdst[(q + j) * GL_LDK + kk + 16 * i] = v[4 * i + j];
  }
};
struct CmWgradG4 {
  const float* gy; int Cout, c0; long pbeg, pend;
  __device__ __forceinline__ void fetch(float (&v)[8], int k0) const {
    const int c = c0 + 4 * (2 * (threadIdx.x >> 5) + (threadIdx.x & 1));
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const long p = pbeg + k0 + ((threadIdx.x & 31) >> 1) + 16 * i;
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p < pend && c < Cout) val = dktb_ld4(gy + p * Cout + c);
      v[4 * i] = val.x; v[4 * i + 1] = val.y; v[4 * i + 2] = val.z; v[4 * i + 3] = val.w;
    }
  }
  __device__ __forceinline__ void store(float* sdst, const float (&v)[8]) const {
    const int q = 4 * (2 * (threadIdx.x >> 5) + (threadIdx.x & 1)), kk = (threadIdx.x & 31) >> 1;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) sdst[(q + j) * GL_LDK + kk + 16 * i] = v[4 * i + j];
  }
};

// Narrow inputs (the 3-channel stem, backbone.py:343-347): a 64-wide tile over the flattened (r, s, ci) index instead
// of one tap x 64 mostly absent channels; consecutive indices are consecutive addresses along a filter row.
struct CmWgradXFlat {
  const float* x; ConvGeo g; int q0; long pbeg, pend;
  mutable int n, oh, ow; mutable long pcur;
  __device__ __forceinline__ void seek(long p) const {
    pcur = p;
    n = (int)(p / ((long)g.Ho * g.Wo));
    const int rem = (int)(p - (long)n * g.Ho * g.Wo);
    oh = rem / g.Wo;
    ow = rem - oh * g.Wo;
  }
  __device__ __forceinline__ void fetch(float (&v)[8], int k0) const {
    const int qq = q0 + (threadIdx.x & 63);
    const bool qok = qq < g.R * g.S * g.Cin;
    const int r = qq / (g.S * g.Cin), rem = qq - r * g.S * g.Cin, s = rem / g.Cin, ci = rem - s * g.Cin;
    int nn = n, hh = oh, ww = ow;
    long p = pcur;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float val = 0.f;
      if (p < pend && qok) {
        const int ih = hh * g.stride - g.pad + r * g.dil, iw = ww * g.stride - g.pad + s * g.dil;
        if ((unsigned)ih < (unsigned)g.H && (unsigned)iw < (unsigned)g.W)
          val = x[(((long)nn * g.H + ih) * g.W + iw) * g.Cin + ci];
      }
      v[i] = val;
      p += 4;
      ww += 4;
      while (ww >= g.Wo) { ww -= g.Wo; if (++hh == g.Ho) { hh = 0; ++nn; } }
    }
    n = nn; oh = hh; ow = ww; pcur = p;
  }
  __device__ __forceinline__ void store(float* sdst, const float (&v)[8]) const {
    const int q = threadIdx.x & 63, kk = threadIdx.x >> 6;
#pragma unroll
    for (int i = 0; i < 8; ++i) sdst[q * GL_LDK + kk + 4 * i] = v[i];
  }
};

// partial[split][(r, s, ci)][co]: the same memory layout as [split][tap][ci][co]
__global__ void __launch_bounds__(GL_THREADS, 2) conv2d_wgrad_flat_mma_kernel(const float* __restrict__ x,
                                                                              const float* __restrict__ gy,
                                                                              float* __restrict__ partial, ConvGeo g,
                                                                              int nsplit) {
  __shared__ __align__(16) float s_as[2 * GL_PLANE];
  __shared__ __align__(16) float s_bs[2 * GL_PLANE];
  typedef GlMap<true> M;
  const int q0 = blockIdx.x * GL_T, co0 = blockIdx.y * GL_T, split = blockIdx.z;
  const int Q = g.R * g.S * g.Cin;
  const long npix = (long)g.N * g.Ho * g.Wo;
  long per = (npix + nsplit - 1) / nsplit;
  per = (per + GL_KC - 1) / GL_KC * GL_KC;
  const long pbeg = split * per, pend = (pbeg + per < npix) ? pbeg + per : npix;
  float acc[4][4];
  gl_zero(acc);
  if (pbeg < pend) {
    CmWgradXFlat la{x, g, q0, pbeg, pend};
    la.seek(pbeg + (threadIdx.x >> 6));
    if (g.Cout % 4 == 0) {
      CmWgradG4 lb{gy, g.Cout, co0, pbeg, pend};
      gl_product_ps(acc, s_as, s_bs, la, lb, 0, (int)(pend - pbeg));
    } else {
      CmWgradG lb{gy, g.Cout, co0, pbeg, pend};
      gl_product_ps(acc, s_as, s_bs, la, lb, 0, (int)(pend - pbeg));
    }
  }
  float* o = partial + (long)split * Q * g.Cout;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qq = q0 + M::row(i);
    if (qq >= Q) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + M::col(j);
      if (co < g.Cout) o[(long)qq * g.Cout + co] = acc[i][j];
    }
  }
}

__global__ void __launch_bounds__(GL_THREADS, 2) conv2d_wgrad_mma_kernel(const float* __restrict__ x,
                                                                         const float* __restrict__ gy,
                                                                         float* __restrict__ partial, ConvGeo g,
                                                                         int nsplit) {
  __shared__ __align__(16) float s_as[2 * GL_PLANE];
  __shared__ __align__(16) float s_bs[2 * GL_PLANE];
  typedef GlMap<true> M;
  const int ci0 = blockIdx.x * GL_T, co0 = blockIdx.y * GL_T;
  const int rs = blockIdx.z / nsplit, split = blockIdx.z % nsplit;
  const int r = rs / g.S, s = rs % g.S;
  const long npix = (long)g.N * g.Ho * g.Wo;
  long per = (npix + nsplit - 1) / nsplit;
  per = (per + GL_KC - 1) / GL_KC * GL_KC;
  const long pbeg = split * per, pend = (pbeg + per < npix) ? pbeg + per : npix;
  float acc[4][4];
  gl_zero(acc);
  if (pbeg < pend) {
    if (g.Cin % 4 == 0 && g.Cout % 4 == 0) {
      CmWgradX4 la{x, g, ci0, r, s, pbeg, pend};
      la.seek(pbeg + ((threadIdx.x & 31) >> 1));
      CmWgradG4 lb{gy, g.Cout, co0, pbeg, pend};
      gl_product_ps(acc, s_as, s_bs, la, lb, 0, (int)(pend - pbeg));
    } else {
      CmWgradX la{x, g, ci0, r, s, pbeg, pend};
      la.seek(pbeg + (threadIdx.x >> 6));
      CmWgradG lb{gy, g.Cout, co0, pbeg, pend};
      gl_product_ps(acc, s_as, s_bs, la, lb, 0, (int)(pend - pbeg));
    }
  }
  float* o = partial + ((long)split * g.R * g.S + rs) * g.Cin * g.Cout;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ci = ci0 + M::row(i);
    if (ci >= g.Cin) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + M::col(j);
      if (co < g.Cout) o[(long)ci * g.Cout + co] = acc[i][j];
    }
  }
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

static bool conv2d_use_mma() {
#ifdef DKTB_EMU      // tests switch it per test case (the emulated tensor-core tiles are slow at ResNet sizes)
  const char* v = getenv("DKTB_RESNET_CONV");
  return !(v && v[0] == 'f');
#else
  return true;       // device build: the tensor-core tiles, no switch
#endif
}

// 1 if dktb_conv2d_fwd_mma / dktb_conv2d_dgrad_mma serve this layer (reduction widths multiples of 32)
DKTB_EXPORT int dktb_conv2d_mma_ok(int Cin, int Cout) { return conv2d_use_mma() && Cin % GL_KC == 0 && Cout % GL_KC == 0; }

DKTB_EXPORT int dktb_conv2d_prep_mma(const float* w, float* wf, float* wd, int Cout, int Cin, int R, int S,
                                     cudaStream_t stream) {
  DKTB_CHECK_ARG(w && wf && wd && Cout > 0 && Cin > 0 && R > 0 && S > 0);
  const long total = (long)Cout * Cin * R * S;
  DKTB_LAUNCH(conv2d_prep_mma_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, w, wf, wd, Cout, Cin,
              R * S);
  return dktb_launch_status();
}

// narrow inputs (the stem): wflat [Cout][Kp], k = (r * S + s) * Cin + ci, Kp = R*S*Cin rounded up to 32 (zero padded)
__global__ void conv2d_prep_flat_mma_kernel(const float* __restrict__ w, float* __restrict__ wflat, int Cout, int Cin,
                                            int RS, int Kp) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)Cout * Kp) return;
  const int k = (int)(i % Kp), co = (int)(i / Kp);
  float v = 0.f;
  if (k < RS * Cin) {
    const int tap = k / Cin, ci = k - tap * Cin;
    v = w[((long)co * Cin + ci) * RS + tap];
  }
  wflat[i] = v;
}

DKTB_EXPORT int dktb_conv2d_flat_k(int Cin, int R, int S) { return (R * S * Cin + GL_KC - 1) / GL_KC * GL_KC; }

DKTB_EXPORT int dktb_conv2d_prep_flat_mma(const float* w, float* wflat, int Cout, int Cin, int R, int S,
                                          cudaStream_t stream) {
  DKTB_CHECK_ARG(w && wflat && Cout > 0 && Cin > 0 && R > 0 && S > 0);
  const int Kp = dktb_conv2d_flat_k(Cin, R, S);
  const long total = (long)Cout * Kp;
  DKTB_LAUNCH(conv2d_prep_flat_mma_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, w, wflat, Cout,
              Cin, R * S, Kp);
  return dktb_launch_status();
}

// forward of a narrow-input layer from wflat (any Cin; meant for Cin < 32)
DKTB_EXPORT int dktb_conv2d_fwd_flat_mma(const float* x, const float* wflat, const float* bias, float* out, int N, int H,
                                         int W, int Cin, int Cout, int R, int S, int stride, int pad, int dil,
                                         cudaStream_t stream) {
  DKTB_CHECK_ARG(x && wflat && out && N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && R > 0 && S > 0 && stride > 0);
  const ConvGeo g = make_geo(N, H, W, Cin, Cout, R, S, stride, pad, dil);
  DKTB_CHECK_ARG(g.Ho > 0 && g.Wo > 0 && (long)N * H * W < 2147483647L);
  const long npix = (long)N * g.Ho * g.Wo;
  DKTB_LAUNCH(conv2d_mma_kernel<2>, dim3((unsigned)((npix + GL_T - 1) / GL_T), (Cout + GL_T - 1) / GL_T),
              dim3(GL_THREADS), 0, stream, x, wflat, bias, out, g);
  return dktb_launch_status();
}

DKTB_EXPORT int dktb_conv2d_fwd_mma(const float* x, const float* wf, const float* bias, float* out, int N, int H, int W,
                                    int Cin, int Cout, int R, int S, int stride, int pad, int dil, cudaStream_t stream) {
  DKTB_CHECK_ARG(x && wf && out && N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && R > 0 && S > 0 && stride > 0);
  DKTB_CHECK_ARG(Cin % GL_KC == 0);
  const ConvGeo g = make_geo(N, H, W, Cin, Cout, R, S, stride, pad, dil);
  DKTB_CHECK_ARG(g.Ho > 0 && g.Wo > 0 && (long)N * H * W < 2147483647L);
  const long npix = (long)N * g.Ho * g.Wo;
  DKTB_LAUNCH(conv2d_mma_kernel<0>, dim3((unsigned)((npix + GL_T - 1) / GL_T), (Cout + GL_T - 1) / GL_T),
              dim3(GL_THREADS), 0, stream, x, wf, bias, out, g);
  return dktb_launch_status();
}

DKTB_EXPORT int dktb_conv2d_dgrad_mma(const float* gy, const float* wd, float* gx, int N, int H, int W, int Cin,
                                      int Cout, int R, int S, int stride, int pad, int dil, cudaStream_t stream) {
  DKTB_CHECK_ARG(gy && wd && gx && N > 0 && Cout % GL_KC == 0);
  const ConvGeo g = make_geo(N, H, W, Cin, Cout, R, S, stride, pad, dil);
  DKTB_CHECK_ARG(g.Ho > 0 && g.Wo > 0 && (long)N * H * W < 2147483647L);
  const long npix = (long)N * H * W;
  if (stride > 1 && stride <= 4 && R * S <= 64) {        // class-by-class tiles: only the taps a pixel can meet
    CmStrideClasses cls;
    int tiles = 0;
    for (int c = 0; c < stride * stride; ++c) {
      const int ca = c / stride, cb = c % stride;
      const int fh = ((ca - pad) % stride + stride) % stride, fw = ((cb - pad) % stride + stride) % stride;
      const int Ha = fh < H ? (H - fh + stride - 1) / stride : 0, Wb = fw < W ? (W - fw + stride - 1) / stride : 0;
      cls.tile_begin[c] = tiles;
      tiles += (int)(((long)N * Ha * Wb + GL_T - 1) / GL_T);
    }
    for (int c = stride * stride; c < 17; ++c) cls.tile_begin[c] = tiles;
    DKTB_LAUNCH(conv2d_dgrad_strided_mma_kernel, dim3((unsigned)tiles, (Cin + GL_T - 1) / GL_T), dim3(GL_THREADS), 0,
                stream, gy, wd, gx, g, cls);
    return dktb_launch_status();
  }
  DKTB_LAUNCH(conv2d_mma_kernel<1>, dim3((unsigned)((npix + GL_T - 1) / GL_T), (Cin + GL_T - 1) / GL_T),
              dim3(GL_THREADS), 0, stream, gy, wd, (const float*)nullptr, gx, g);
  return dktb_launch_status();
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
  return (int)(n < 1 ? 1 : (n > 98 ? 98 : n));     // 3 x 98 (stem) and 9 x 98 (3x3 layers) CTAs are whole waves of 2 x 148
}

// dw [Cout,Cin,R,S], db [Cout] (nullable); scratch: nsplit*R*S*Cin*Cout floats, nsplit = dktb_conv2d_wgrad_nsplit(N*Ho*Wo)
DKTB_EXPORT int dktb_conv2d_wgrad(const float* x, const float* gy, const float* yout, float* dw, float* db,
                                  float* scratch, int N, int H, int W, int Cin, int Cout, int R, int S, int stride,
                                  int pad, int dil, int relu, cudaStream_t stream) {
  DKTB_CHECK_ARG(x && gy && dw && scratch && (!relu || yout) && N > 0);
  const ConvGeo g = make_geo(N, H, W, Cin, Cout, R, S, stride, pad, dil);
  const long npix = (long)N * g.Ho * g.Wo;
  const int nsplit = dktb_conv2d_wgrad_nsplit(npix);
  if (!relu && conv2d_use_mma() && Cin < 16)         // the stem: tiles over the flattened (r, s, ci) index
    DKTB_LAUNCH(conv2d_wgrad_flat_mma_kernel, dim3((R * S * Cin + GL_T - 1) / GL_T, (Cout + GL_T - 1) / GL_T, nsplit),
                dim3(GL_THREADS), 0, stream, x, gy, scratch, g, nsplit);
  else if (!relu && conv2d_use_mma())                // tensor-core tiles (the ReLU-fused form is only used by Conv3)
    DKTB_LAUNCH(conv2d_wgrad_mma_kernel, dim3((Cin + GL_T - 1) / GL_T, (Cout + GL_T - 1) / GL_T, R * S * nsplit),
                dim3(GL_THREADS), 0, stream, x, gy, scratch, g, nsplit);
  else
    DKTB_LAUNCH(conv2d_wgrad_kernel, dim3((Cin + CG_TM - 1) / CG_TM, (Cout + CG_TN - 1) / CG_TN, R * S * nsplit),
                dim3(256), 0, stream, x, gy, yout, scratch, g, relu, nsplit);
  const long total = (long)R * S * Cin * Cout;
  DKTB_LAUNCH(conv2d_wgrad_reduce_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream,
              (const float*)scratch, dw, g, nsplit);
  if (db != nullptr) {         // the weight-gradient partials are consumed: the scratch doubles as the slice buffer
    long cap = (long)nsplit * R * S * Cin;
    const int nslice = (int)(cap < 128 ? cap : 128);
    DKTB_LAUNCH(conv2d_bgrad_kernel, dim3((Cout + 63) / 64, nslice), dim3(256), 0, stream, gy, yout, scratch, npix, Cout,
                relu, nslice);
    DKTB_LAUNCH(conv2d_bgrad_final_kernel, dim3((Cout + 127) / 128), dim3(128), 0, stream, (const float*)scratch, db, Cout,
                nslice);
  }
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
