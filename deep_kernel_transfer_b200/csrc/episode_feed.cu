// Episode feeder: the per-image transform pipeline of the reference's episode loader, evaluated on the
// device from a resident uint8 image store -- one CTA per output image, the whole pipeline in shared memory, the fp32
// NCHW episode tensor written once.  Replaces, per image, the PIL / torchvision chain composed in
// data/datamgr.py:37-46 (called from data/dataset.py:66-70):
//     aug:    RandomSizedCrop(S) -> ImageJitter (data/additional_transforms.py:24-34) -> RandomHorizontalFlip
//             -> ToTensor -> Normalize
//     plain:  Scale([int(1.15 S)]*2) -> CenterCrop(S) -> ToTensor -> Normalize
// Results are bit-identical to PIL / torchvision for the same crop box, jitter factors and flip flag (the random draws
// are inputs): Pillow's 8-bit resample is a two-pass fixed-point convolution (22-bit coefficients computed in double
// precision, horizontal pass first, intermediate rounded to uint8), ImageEnhance is Image.blend in float32 with a
// truncating cast, ToTensor / Normalize are float32 division, subtraction, division.  All floating-point steps below
// therefore use the explicitly rounded intrinsics (no fused multiply-add contraction).
//
// HBM-bound byte work: algorithmic bytes per image = the crop box read once (3 B / pixel) + 12 S^2 B written.
#include "dktb_common.cuh"

#define FEED_THREADS 512
#define FEED_PRECISION_BITS 22

struct FeedSmem {
  int* bh;            // [S][2] horizontal (xmin, count)
  int* bv;            // [S][2] vertical
  int* kh;            // [S][kstride] horizontal coefficients (kstride odd: conflict-free across output columns)
  int* kv;            // [S][kstride]
  int* band;          // [S] end (exclusive) of the band of output rows starting at each band start
  int* red;           // [32] block reduction + broadcast slots
  float* lut_out;     // [3][256] ((u / 255) - mean_c) / std_c
  unsigned char* lut_b;  // [256] Brightness blend of every byte value
  unsigned char* lut_c;  // [256] Contrast blend of every byte value
  unsigned char* img; // [S*S*3] resized image, HWC
  unsigned char* tmp; // [tmp_rows][S][3] horizontally resampled rows of the current band
};

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for output index `o` of `out_size`, box [0, in_size).
// Returns the tap count (0 and *bad = 1 if it exceeds kmax).
__device__ __forceinline__ void feed_coeffs(int o, int in_size, int out_size, int kmax, int* bounds, int* k, int* bad) {
  const double in1 = (double)(float)in_size;
  const double scale = in1 / (double)out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = filterscale;                       // bilinear: support 1.0 * filterscale
  const double ss = 1.0 / filterscale;
  const double center = __dadd_rn(0.0, __dmul_rn((double)o + 0.5, scale));
  int xmin = (int)(__dadd_rn(__dadd_rn(center, -support), 0.5));
  if (xmin < 0) xmin = 0;
  int xmax = (int)(__dadd_rn(__dadd_rn(center, support), 0.5));
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  if (xmax > kmax) {
    *bad = 1;
    xmax = 0;
  }
  double ww = 0.0;
  for (int x = 0; x < xmax; ++x) {
    double a = __dmul_rn(__dadd_rn(__dadd_rn((double)(x + xmin), -center), 0.5), ss);
    if (a < 0.0) a = -a;
    const double w = a < 1.0 ? __dadd_rn(1.0, -a) : 0.0;
    ww = __dadd_rn(ww, w);
  }
  for (int x = 0; x < xmax; ++x) {
    double a = __dmul_rn(__dadd_rn(__dadd_rn((double)(x + xmin), -center), 0.5), ss);
    if (a < 0.0) a = -a;
    double w = a < 1.0 ? __dadd_rn(1.0, -a) : 0.0;
    if (ww != 0.0) w = w / ww;
    const double s = __dmul_rn(w, (double)(1 << FEED_PRECISION_BITS));
    k[x] = w < 0.0 ? (int)__dadd_rn(-0.5, s) : (int)__dadd_rn(0.5, s);
  }
  bounds[0] = xmin;
  bounds[1] = xmax;
}

__device__ __forceinline__ unsigned char feed_clip8(int acc) {
  acc = (acc + (1 << (FEED_PRECISION_BITS - 1))) >> FEED_PRECISION_BITS;
  return (unsigned char)(acc < 0 ? 0 : (acc > 255 ? 255 : acc));
}

// Blend.c: im1 + alpha * (im2 - im1), float32, truncating cast; clipped when alpha is outside [0, 1].
__device__ __forceinline__ int feed_blend(int a, int b, float alpha) {
  if (alpha == 0.f) return a;
  if (alpha == 1.f) return b;
  const float t = __fadd_rn((float)a, __fmul_rn(alpha, (float)(b - a)));
  if (alpha >= 0.f && alpha <= 1.f) return (int)t;
  return t <= 0.f ? 0 : (t >= 255.f ? 255 : (int)t);
}

__device__ __forceinline__ int feed_L(int r, int g, int b) { return (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16; }

// params [B][8] int: image id, crop top, crop left, crop h, crop w, flip, jitter on/off, unused
// factors [B][3] float: brightness, contrast, color
// desc [n_images][3] long: byte offset into `store`, height, width (HWC uint8, rows contiguous)
// The crop box is resized to RH x RW; the S x S window at (oy, ox) of that result is the output.
__global__ void __launch_bounds__(FEED_THREADS)
episode_transform_kernel(const unsigned char* __restrict__ store, const long* __restrict__ desc, long n_images,
                         const int* __restrict__ params, const float* __restrict__ factors, float* __restrict__ out,
                         int S, int RH, int RW, int oy, int ox, int kmax, int tmp_rows, float m0, float m1, float m2,
                         float s0, float s1, float s2, int* __restrict__ err) {
  DKTB_DYN_SMEM(unsigned char, smem);
  const int kstride = kmax | 1;
  FeedSmem sm;
  sm.bh = reinterpret_cast<int*>(smem);
  sm.bv = sm.bh + 2 * S;
  sm.kh = sm.bv + 2 * S;
  sm.kv = sm.kh + S * kstride;
  sm.band = sm.kv + S * kstride;
  sm.red = sm.band + S;
  sm.lut_out = reinterpret_cast<float*>(sm.red + 32);
  sm.lut_b = reinterpret_cast<unsigned char*>(sm.lut_out + 768);
  sm.lut_c = sm.lut_b + 256;
  sm.img = sm.lut_c + 256;
  sm.tmp = sm.img + ((S * S * 3 + 15) & ~15);

  const int b = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
  const int* p = params + (long)b * 8;
  const long id = p[0];
  const int top = p[1], left = p[2], ch = p[3], cw = p[4], flip = p[5], jitter = p[6];
  if (tid == 0) sm.red[0] = 0;
  __syncthreads();
  long off = 0;
  int H = 0, W = 0;
  bool ok = id >= 0 && id < n_images;
  if (ok) {
    off = desc[id * 3];
    H = (int)desc[id * 3 + 1];
    W = (int)desc[id * 3 + 2];
    ok = ch > 0 && cw > 0 && top >= 0 && left >= 0 && top + ch <= H && left + cw <= W;
  }
  if (!ok) {                                    // uniform over the CTA
    if (tid == 0) atomicMax(err, 3);
    return;
  }
  // -- coefficients of the S output columns / rows that are kept; the ToTensor / Normalize table of every byte value
  for (int t = tid; t < 2 * S; t += nthr) {
    if (t < S) feed_coeffs(ox + t, cw, RW, kmax, sm.bh + 2 * t, sm.kh + t * kstride, sm.red);
    else feed_coeffs(oy + (t - S), ch, RH, kmax, sm.bv + 2 * (t - S), sm.kv + (t - S) * kstride, sm.red);
  }
  for (int t = tid; t < 768; t += nthr) {
    const int c = t >> 8;
    const float mc = c == 0 ? m0 : (c == 1 ? m1 : m2), sc = c == 0 ? s0 : (c == 1 ? s1 : s2);
    sm.lut_out[t] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)(t & 255), 255.f), mc), sc);
  }
  float fb = 1.f, fc = 1.f, fs = 1.f;
  if (jitter) {
    fb = factors[b * 3];
    fc = factors[b * 3 + 1];
    fs = factors[b * 3 + 2];
    for (int t = tid; t < 256; t += nthr) sm.lut_b[t] = (unsigned char)feed_blend(0, t, fb);
  }
  __syncthreads();
  if (sm.red[0]) {
    if (tid == 0) atomicMax(err, 1);
    return;
  }
  // -- bands of output rows whose input rows fit tmp (one thread walks the S rows once)
  if (tid == 0) {
    int a = 0, bad = 0;
    while (a < S) {
      const int r0 = sm.bv[2 * a];
      int e = a + 1;
      while (e < S && sm.bv[2 * e] + sm.bv[2 * e + 1] - r0 <= tmp_rows) ++e;
      if (sm.bv[2 * (e - 1)] + sm.bv[2 * (e - 1) + 1] - r0 > tmp_rows) bad = 1;
      sm.band[a] = e;
      a = e;
    }
    sm.red[1] = bad;
  }
  __syncthreads();
  if (sm.red[1]) {
    if (tid == 0) atomicMax(err, 2);
    return;
  }
  const unsigned char* src = store + off + ((long)top * W + left) * 3;
  const long rstride = (long)W * 3;
  const int S3 = S * 3;
  const bool words = (S3 & 3) == 0;             // rows of tmp / img are whole 32-bit words
  // -- per band: horizontal pass of the input rows it needs into tmp, then its vertical pass into img
  for (int a = 0; a < S; a = sm.band[a]) {
    const int e = sm.band[a];
    const int r0 = sm.bv[2 * a];
    const int r1 = sm.bv[2 * (e - 1)] + sm.bv[2 * (e - 1) + 1];
    // a warp per input row, a lane per output column, all three channels of the pixel in one thread
    for (int r = warp; r < r1 - r0; r += nwarp) {
      const unsigned char* rowp = src + (long)(r0 + r) * rstride;
      unsigned char* trow = sm.tmp + r * S3;
      for (int xx = lane; xx < S; xx += 32) {
        const int n = sm.bh[2 * xx + 1];
        const int* k = sm.kh + xx * kstride;
        const unsigned char* q = rowp + sm.bh[2 * xx] * 3;
        int a0 = 0, a1 = 0, a2 = 0;
#pragma unroll 4
        for (int x = 0; x < n; ++x) {
          const int kk = k[x];
          a0 += (int)__ldg(q + 3 * x) * kk;
          a1 += (int)__ldg(q + 3 * x + 1) * kk;
          a2 += (int)__ldg(q + 3 * x + 2) * kk;
        }
        trow[xx * 3] = feed_clip8(a0);
        trow[xx * 3 + 1] = feed_clip8(a1);
        trow[xx * 3 + 2] = feed_clip8(a2);
      }
    }
    __syncthreads();
    // a warp per output row; a lane per 32-bit word (four bytes) of the row
    for (int yy = a + warp; yy < e; yy += nwarp) {
      const int ymin = sm.bv[2 * yy] - r0, n = sm.bv[2 * yy + 1];
      const int* k = sm.kv + yy * kstride;
      if (words) {
        const int W4 = S3 >> 2;
        const unsigned* tw = reinterpret_cast<const unsigned*>(sm.tmp) + ymin * W4;
        unsigned* iw = reinterpret_cast<unsigned*>(sm.img) + yy * W4;
        for (int w = lane; w < W4; w += 32) {
          int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll 4
          for (int y = 0; y < n; ++y) {
            const unsigned v = tw[y * W4 + w];
            const int kk = k[y];
            a0 += (int)(v & 255u) * kk;
            a1 += (int)((v >> 8) & 255u) * kk;
            a2 += (int)((v >> 16) & 255u) * kk;
            a3 += (int)(v >> 24) * kk;
          }
          iw[w] = (unsigned)feed_clip8(a0) | ((unsigned)feed_clip8(a1) << 8) | ((unsigned)feed_clip8(a2) << 16) |
                  ((unsigned)feed_clip8(a3) << 24);
        }
      } else {
        const unsigned char* tb = sm.tmp + ymin * S3;
        for (int xc = lane; xc < S3; xc += 32) {
          int acc = 0;
          for (int y = 0; y < n; ++y) acc += (int)tb[y * S3 + xc] * k[y];
          sm.img[yy * S3 + xc] = feed_clip8(acc);
        }
      }
    }
    __syncthreads();
  }
  // -- ImageJitter: Brightness (table) -> Contrast (needs the mean of the L image; table) -> Color; a thread owns
  //    whole pixels
  if (jitter) {
    int lsum = 0;
    for (int i = tid; i < S * S; i += nthr) {
      unsigned char* px = sm.img + i * 3;
      const int r = sm.lut_b[px[0]], g = sm.lut_b[px[1]], bl = sm.lut_b[px[2]];
      px[0] = (unsigned char)r;
      px[1] = (unsigned char)g;
      px[2] = (unsigned char)bl;
      lsum += feed_L(r, g, bl);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    if (lane == 0) sm.red[2 + warp] = lsum;
    __syncthreads();
    if (tid == 0) {
      long tot = 0;
      for (int w = 0; w < nwarp; ++w) tot += sm.red[2 + w];
      const long cnt = (long)S * S;
      sm.red[0] = (int)((2 * tot + cnt) / (2 * cnt));        // int(sum / count + 0.5), exact in integers
    }
    __syncthreads();
    const int mean = sm.red[0];
    for (int t = tid; t < 256; t += nthr) sm.lut_c[t] = (unsigned char)feed_blend(mean, t, fc);
    __syncthreads();
    for (int i = tid; i < S * S; i += nthr) {
      unsigned char* px = sm.img + i * 3;
      const int r = sm.lut_c[px[0]], g = sm.lut_c[px[1]], bl = sm.lut_c[px[2]];
      const int L = feed_L(r, g, bl);
      px[0] = (unsigned char)feed_blend(L, r, fs);
      px[1] = (unsigned char)feed_blend(L, g, fs);
      px[2] = (unsigned char)feed_blend(L, bl, fs);
    }
    __syncthreads();
  }
  // -- flip, ToTensor, Normalize (table); fp32 NCHW, written once
  float* o = out + (long)b * 3 * S * S;
  if ((S & 3) == 0) {
    const int S4 = S >> 2;
    for (int c = 0; c < 3; ++c) {
      const float* lut = sm.lut_out + c * 256;
      for (int y = warp; y < S; y += nwarp) {
        const unsigned char* row = sm.img + y * S3 + c;
        for (int x4 = lane; x4 < S4; x4 += 32) {
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int x = flip ? S - 1 - (4 * x4 + j) : 4 * x4 + j;
            v[j] = lut[row[x * 3]];
          }
          dktb_st4(o + ((long)c * S + y) * S + 4 * x4, make_float4(v[0], v[1], v[2], v[3]));
        }
      }
    }
  } else {
    for (int i = tid; i < 3 * S * S; i += nthr) {
      const int c = i / (S * S), rem = i - c * S * S;
      const int y = rem / S, xo = rem - y * S;
      const int x = flip ? S - 1 - xo : xo;
      o[i] = sm.lut_out[c * 256 + sm.img[(y * S + x) * 3 + c]];
    }
  }
}

static long feed_smem_bytes(int S, int kmax, int tmp_rows) {
  return (long)(4 * S + 2 * S * (kmax | 1) + S + 32 + 768) * 4 + 512 + ((S * S * 3 + 15) & ~15) + (long)tmp_rows * S * 3;
}

// Shared memory the transform needs for output size S, at most kmax taps per output pixel and tmp_rows rows of
// horizontally resampled input kept per band.
DKTB_EXPORT long dktb_episode_transform_smem(int S, int kmax, int tmp_rows) {
  if (S <= 0 || kmax <= 0 || tmp_rows <= 0) return DKTB_BAD_ARG;
  return feed_smem_bytes(S, kmax, tmp_rows);
}

// out [B,3,S,S] fp32.  err: device int, zero-initialised by the caller; set to 1 (an image needs more than kmax
// taps), 2 (one output row needs more than tmp_rows input rows) or 3 (bad image id / crop box outside the image).
DKTB_EXPORT int dktb_episode_transform(const unsigned char* store, const long* desc, long n_images, const int* params,
                                       const float* factors, float* out, int B, int S, int RH, int RW, int oy, int ox,
                                       int kmax, int tmp_rows, float mean_r, float mean_g, float mean_b, float std_r,
                                       float std_g, float std_b, int* err, cudaStream_t stream) {
  DKTB_CHECK_ARG(store && desc && params && out && err);
  DKTB_CHECK_ARG(B > 0 && S > 0 && RH >= S + oy && RW >= S + ox && oy >= 0 && ox >= 0 && kmax >= 2 && tmp_rows >= 1);
  const long smem = feed_smem_bytes(S, kmax, tmp_rows);
  DKTB_CHECK_ARG(smem <= 227 * 1024);
  cudaFuncSetAttribute(episode_transform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  DKTB_LAUNCH(episode_transform_kernel, dim3(B), dim3(FEED_THREADS), (size_t)smem, stream, store, desc, n_images, params,
              factors, out, S, RH, RW, oy, ox, kmax, tmp_rows, mean_r, mean_g, mean_b, std_r, std_g, std_b, err);
  return dktb_launch_status();
}
