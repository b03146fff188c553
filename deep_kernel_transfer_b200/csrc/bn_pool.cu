// BatchNorm2d (+ReLU +MaxPool2d(2)) of a ConvBlock, forward and backward, with PER-EPISODE batch
// statistics: E episodes are packed along the image axis and every group of `ipe` consecutive
// images is one BatchNorm batch, exactly what the reference sees when it runs one episode per
// forward (methods/DKT.py:140-141, backbone.py:116-121; momentum 0.1, eps 1e-5, biased variance for
// normalisation, unbiased for the running estimate).  All kernels are HBM-bound element-wise /
// reduction passes over NHWC fp32 tensors with float4 accesses (16 lanes cover one pixel's 64
// channels = 256 contiguous bytes); reductions are two-stage (per-CTA partials, then a fixed-order
// sum in double) so results are run-to-run deterministic.
#include "dktb_common.cuh"

// ------------------------------------------------------------------------------------------------
// partials [B][T][2][64]  ->  sums [E][2][64]  (fixed order, double accumulation), two levels:
// grid (E, ES_CHUNKS) CTAs each sum a contiguous slice into chunk_sums [E][ES_CHUNKS][128] (double),
// then one small CTA per episode adds the ES_CHUNKS slices in order.
// ------------------------------------------------------------------------------------------------
#define ES_CHUNKS 32
__global__ void __launch_bounds__(256) episode_sum_kernel(const float* __restrict__ partials, int T, int ipe,
                                                          double* __restrict__ chunk_sums) {
  __shared__ double s_acc[2][128];
  const int e = blockIdx.x, ch = blockIdx.y;
  const int tid = threadIdx.x;
  const int col = tid % 128, slice = tid / 128;     // col = which*64 + c ; 2 slices
  const long n = (long)ipe * T;
  const long per = (n + ES_CHUNKS - 1) / ES_CHUNKS;
  const long lo = ch * per, hi = (lo + per < n) ? lo + per : n;
  const float* base = partials + (long)e * n * 128;
  double acc0 = 0.0, acc1 = 0.0;
  long i = lo + slice;
  for (; i + 2 < hi; i += 4) {
    acc0 += (double)base[i * 128 + col];
    acc1 += (double)base[(i + 2) * 128 + col];
  }
  for (; i < hi; i += 2) acc0 += (double)base[i * 128 + col];
  s_acc[slice][col] = acc0 + acc1;
  __syncthreads();
  if (tid < 128) chunk_sums[((long)e * ES_CHUNKS + ch) * 128 + tid] = s_acc[0][tid] + s_acc[1][tid];
}

__global__ void episode_sum_final_kernel(const double* __restrict__ chunk_sums, float* __restrict__ sums,
                                         double* __restrict__ sums_d) {
  const int e = blockIdx.x, tid = threadIdx.x;     // 128 threads
  double t = 0.0;
  for (int ch = 0; ch < ES_CHUNKS; ++ch) t += chunk_sums[((long)e * ES_CHUNKS + ch) * 128 + tid];
  if (sums) sums[(long)e * 128 + tid] = (float)t;
  if (sums_d) sums_d[(long)e * 128 + tid] = t;
}

// sums_d [E][2][64] (sum, sumsq) -> mean, invstd [E][64]; sequential running-stat EMA over episodes
__global__ void bn_stats_kernel(const double* __restrict__ sums_d, int E, double count, float* __restrict__ mean,
                                float* __restrict__ invstd, float* __restrict__ running_mean,
                                float* __restrict__ running_var, float momentum, float eps) {
  const int c = threadIdx.x;
  if (c >= 64) return;
  float rm = running_mean ? running_mean[c] : 0.f;
  float rv = running_var ? running_var[c] : 0.f;
  for (int e = 0; e < E; ++e) {
    const double m = sums_d[(long)e * 128 + c] / count;
    double var = sums_d[(long)e * 128 + 64 + c] / count - m * m;
    if (var < 0.0) var = 0.0;
    mean[e * 64 + c] = (float)m;
    invstd[e * 64 + c] = (float)(1.0 / sqrt(var + (double)eps));
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    rm = (1.f - momentum) * rm + momentum * (float)m;
    rv = (1.f - momentum) * rv + momentum * (float)unbiased;
  }
  if (running_mean) running_mean[c] = rm;
  if (running_var) running_var[c] = rv;
}

// eval mode: mean = running_mean, invstd = 1/sqrt(running_var + eps)   (n features)
__global__ void bn_eval_prepare_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                       float* __restrict__ mean, float* __restrict__ invstd, int n, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    mean[i] = running_mean[i];
    invstd[i] = 1.f / sqrtf(running_var[i] + eps);
  }
}

DKTB_EXPORT int dktb_bn_eval_prepare(const float* running_mean, const float* running_var, float* mean, float* invstd,
                                     int n, float eps, cudaStream_t stream) {
  DKTB_CHECK_ARG(running_mean && running_var && mean && invstd && n > 0);
  DKTB_LAUNCH(bn_eval_prepare_kernel, dim3((n + 255) / 256), dim3(256), 0, stream, running_mean, running_var, mean,
              invstd, n, eps);
  return dktb_launch_status();
}

DKTB_EXPORT long dktb_bn_scratch_doubles(int E) { return (long)E * 128 * (1 + ES_CHUNKS); }

// Train-mode statistics from conv partial sums.  scratch_d: dktb_bn_scratch_doubles(E) doubles.
DKTB_EXPORT int dktb_bn_finalize(const float* partials, int B, int T, int ipe, int hw, float* mean, float* invstd,
                                 float* running_mean, float* running_var, double* scratch_d, float momentum, float eps,
                                 cudaStream_t stream) {
  DKTB_CHECK_ARG(partials && mean && invstd && scratch_d && B > 0 && T > 0 && ipe > 0 && B % ipe == 0 && hw > 0);
  const int E = B / ipe;
  double* chunk = scratch_d + (long)E * 128;
  DKTB_LAUNCH(episode_sum_kernel, dim3(E, ES_CHUNKS), dim3(256), 0, stream, partials, T, ipe, chunk);
  DKTB_LAUNCH(episode_sum_final_kernel, dim3(E), dim3(128), 0, stream, (const double*)chunk, (float*)nullptr,
              scratch_d);
  DKTB_LAUNCH(bn_stats_kernel, dim3(1), dim3(64), 0, stream, (const double*)scratch_d, E, (double)ipe * (double)hw,
              mean, invstd, running_mean, running_var, momentum, eps);
  return dktb_launch_status();
}

// ------------------------------------------------------------------------------------------------
// forward: out = maxpool2( relu( (y - mean) * invstd * gamma + beta ) )
// y: [B][H+2ip][W+2ip][64] (ip = in_pad), out: [B][Ho+2op][Wo+2op][64]; stats row = img/ipe (ipe==0: row 0)
// ------------------------------------------------------------------------------------------------
template <int POOL>
__global__ void __launch_bounds__(256) bn_relu_pool_fwd_kernel(const float* __restrict__ y, const float* __restrict__ mean,
                                                               const float* __restrict__ invstd,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, float* __restrict__ out,
                                                               int B, int H, int W, int ipe, int in_pad, int out_pad) {
  const int Ho = POOL ? H / 2 : H, Wo = POOL ? W / 2 : W;
  const long total = (long)B * Ho * Wo * 16;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c4 = (int)(idx % 16) * 4;
  long t = idx / 16;
  const int ow = (int)(t % Wo);
  t /= Wo;
  const int oh = (int)(t % Ho);
  const int img = (int)(t / Ho);
  const int e = ipe > 0 ? img / ipe : 0;
  const int Hi = H + 2 * in_pad, Wi = W + 2 * in_pad;
  constexpr int NP = POOL ? 2 : 1;
  // all window loads in flight before any arithmetic
  const float* src = y + (((long)img * Hi + (POOL ? oh * 2 : oh) + in_pad) * Wi + (POOL ? ow * 2 : ow) + in_pad) * 64 + c4;
  float4 v[NP * NP];
#pragma unroll
  for (int dy = 0; dy < NP; ++dy)
#pragma unroll
    for (int dx = 0; dx < NP; ++dx) v[dy * NP + dx] = dktb_ld4(src + ((long)dy * Wi + dx) * 64);
  const float4 m = dktb_ld4(mean + e * 64 + c4), is = dktb_ld4(invstd + e * 64 + c4);
  const float4 g = dktb_ld4(gamma + c4), bt = dktb_ld4(beta + c4);
  const float4 sc = make_float4(g.x * is.x, g.y * is.y, g.z * is.z, g.w * is.w);
  float4 best = make_float4(0.f, 0.f, 0.f, 0.f);     // relu floor doubles as the max identity
#pragma unroll
  for (int k = 0; k < NP * NP; ++k) {
    best.x = fmaxf(best.x, fmaf(v[k].x - m.x, sc.x, bt.x));
    best.y = fmaxf(best.y, fmaf(v[k].y - m.y, sc.y, bt.y));
    best.z = fmaxf(best.z, fmaf(v[k].z - m.z, sc.z, bt.z));
    best.w = fmaxf(best.w, fmaf(v[k].w - m.w, sc.w, bt.w));
  }
  const int Hq = Ho + 2 * out_pad, Wq = Wo + 2 * out_pad;
  dktb_st4(out + (((long)img * Hq + oh + out_pad) * Wq + ow + out_pad) * 64 + c4, best);
}

DKTB_EXPORT int dktb_bn_relu_pool_fwd(const float* y, const float* mean, const float* invstd, const float* gamma,
                                      const float* beta, float* out, int B, int H, int W, int ipe, int in_pad,
                                      int out_pad, int pool, cudaStream_t stream) {
  DKTB_CHECK_ARG(y && mean && invstd && gamma && beta && out && B > 0 && H > 0 && W > 0);
  const int Ho = pool ? H / 2 : H, Wo = pool ? W / 2 : W;
  const long total = (long)B * Ho * Wo * 16;
  DKTB_CHECK_ARG(total > 0);
  if (pool) {
    DKTB_LAUNCH(bn_relu_pool_fwd_kernel<1>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, y, mean, invstd,
                gamma, beta, out, B, H, W, ipe, in_pad, out_pad);
  } else {
    DKTB_LAUNCH(bn_relu_pool_fwd_kernel<0>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, y, mean, invstd,
                gamma, beta, out, B, H, W, ipe, in_pad, out_pad);
  }
  return dktb_launch_status();
}

// ------------------------------------------------------------------------------------------------
// backward helpers.  For one output (pooled) pixel and 4 channels: recompute z = relu(bn(y)) of the
// window, find the first maximum (PyTorch MaxPool2d tie rule: strict '>' in scan order), route the
// incoming gradient there if the ReLU was active.  Returns g_z and xhat at the arg-max position.
// ------------------------------------------------------------------------------------------------
struct Win4 {
  float4 v[4];     // raw y of the (up to) 4 window positions
};

// pass 1: partial[(b*chunks + chunk)][2][64] = { sum g_z , sum g_z * xhat } over the chunk's pixels
#define BWD_PIX_PER_CHUNK 256
__global__ void __launch_bounds__(256) bn_relu_pool_bwd_reduce_kernel(
    const float* __restrict__ y, const float* __restrict__ gout, const float* __restrict__ mean,
    const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
    float* __restrict__ partial, int H, int W, int ipe, int in_pad, int out_pad, int pool) {
  __shared__ float s_red[2][16][64 + 4];
  const int Ho = pool ? H / 2 : H, Wo = pool ? W / 2 : W;
  const int img = blockIdx.y, chunk = blockIdx.x;
  const int tid = threadIdx.x, c4 = (tid % 16) * 4, lp = tid / 16;
  const int e = ipe > 0 ? img / ipe : 0;
  const float4 m = dktb_ld4(mean + e * 64 + c4), is = dktb_ld4(invstd + e * 64 + c4);
  const float4 g = dktb_ld4(gamma + c4), bt = dktb_ld4(beta + c4);
  const float mm[4] = {m.x, m.y, m.z, m.w}, ii[4] = {is.x, is.y, is.z, is.w};
  const float ss[4] = {g.x * is.x, g.y * is.y, g.z * is.z, g.w * is.w}, bb[4] = {bt.x, bt.y, bt.z, bt.w};
  const int Hi = H + 2 * in_pad, Wi = W + 2 * in_pad, Hq = Ho + 2 * out_pad, Wq = Wo + 2 * out_pad;
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  const int npix = Ho * Wo;
  const int np = pool ? 2 : 1;
  for (int p = chunk * BWD_PIX_PER_CHUNK + lp; p < npix && p < (chunk + 1) * BWD_PIX_PER_CHUNK; p += 16) {
    const int oh = p / Wo, ow = p % Wo;
    const float4 go = dktb_ld4(gout + (((long)img * Hq + oh + out_pad) * Wq + ow + out_pad) * 64 + c4);
    const float gg[4] = {go.x, go.y, go.z, go.w};
    float yv[4][4];
    int k = 0;
    for (int dy = 0; dy < np; ++dy)
      for (int dx = 0; dx < np; ++dx, ++k) {
        const int h = (pool ? oh * 2 : oh) + dy, w = (pool ? ow * 2 : ow) + dx;
        const float4 v = dktb_ld4(y + (((long)img * Hi + h + in_pad) * Wi + w + in_pad) * 64 + c4);
        yv[0][k] = v.x; yv[1][k] = v.y; yv[2][k] = v.z; yv[3][k] = v.w;
      }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int arg; float gz, xh;
      bn_bwd_route(yv[j], np * np, mm[j], ii[j], ss[j], bb[j], gg[j], arg, gz, xh);
      s1[j] += gz;
      s2[j] = fmaf(gz, xh, s2[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    s_red[0][lp][c4 + j] = s1[j];
    s_red[1][lp][c4 + j] = s2[j];
  }
  __syncthreads();
  if (tid < 128) {
    const int which = tid / 64, c = tid % 64;
    float t = 0.f;
    for (int q = 0; q < 16; ++q) t += s_red[which][q][c];
    partial[(((long)img * gridDim.x + chunk) * 2 + which) * 64 + c] = t;
  }
}

// sums [E][2][64] -> dgamma[c] (+)= sum_e sums[e][1][c], dbeta[c] (+)= sum_e sums[e][0][c]
__global__ void bn_param_grad_kernel(const float* __restrict__ sums, int E, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta) {
  const int c = threadIdx.x;
  if (c >= 64) return;
  double a = 0.0, b = 0.0;
  for (int e = 0; e < E; ++e) {
    b += (double)sums[(long)e * 128 + c];
    a += (double)sums[(long)e * 128 + 64 + c];
  }
  dgamma[c] = (float)a;
  dbeta[c] = (float)b;
}

// pass 2: gy = gamma*invstd * ( g_z - s1/n - xhat * s2/n ) for every position of the full-res map
__global__ void __launch_bounds__(256) bn_relu_pool_bwd_apply_kernel(
    const float* __restrict__ y, const float* __restrict__ gout, const float* __restrict__ mean,
    const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ sums, float* __restrict__ gy, int B, int H, int W, int ipe, int in_pad, int out_pad,
    int pool, float inv_count) {
  const int Ho = pool ? H / 2 : H, Wo = pool ? W / 2 : W;
  const int Hc = pool ? (H + 1) / 2 : H, Wc = pool ? (W + 1) / 2 : W;   // windows incl. the dropped odd edge
  const long total = (long)B * Hc * Wc * 16;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c4 = (int)(idx % 16) * 4;
  long t = idx / 16;
  const int ow = (int)(t % Wc);
  t /= Wc;
  const int oh = (int)(t % Hc);
  const int img = (int)(t / Hc);
  const int e = ipe > 0 ? img / ipe : 0;
  const float4 m = dktb_ld4(mean + e * 64 + c4), is = dktb_ld4(invstd + e * 64 + c4);
  const float4 g = dktb_ld4(gamma + c4), bt = dktb_ld4(beta + c4);
  const float4 q1 = dktb_ld4(sums + (long)e * 128 + c4), q2 = dktb_ld4(sums + (long)e * 128 + 64 + c4);
  const float mm[4] = {m.x, m.y, m.z, m.w}, ii[4] = {is.x, is.y, is.z, is.w};
  const float ss[4] = {g.x * is.x, g.y * is.y, g.z * is.z, g.w * is.w}, bb[4] = {bt.x, bt.y, bt.z, bt.w};
  const float a1[4] = {q1.x * inv_count, q1.y * inv_count, q1.z * inv_count, q1.w * inv_count};
  const float a2[4] = {q2.x * inv_count, q2.y * inv_count, q2.z * inv_count, q2.w * inv_count};
  const int Hi = H + 2 * in_pad, Wi = W + 2 * in_pad, Hq = Ho + 2 * out_pad, Wq = Wo + 2 * out_pad;
  const int np = pool ? 2 : 1;
  const bool full = (oh < Ho) && (ow < Wo);
  float gg[4] = {0.f, 0.f, 0.f, 0.f};
  if (full) {
    const float4 go = dktb_ld4(gout + (((long)img * Hq + oh + out_pad) * Wq + ow + out_pad) * 64 + c4);
    gg[0] = go.x; gg[1] = go.y; gg[2] = go.z; gg[3] = go.w;
  }
  float yv[4][4];
  bool inb[4];
  int k = 0;
  for (int dy = 0; dy < np; ++dy)
    for (int dx = 0; dx < np; ++dx, ++k) {
      const int h = (pool ? oh * 2 : oh) + dy, w = (pool ? ow * 2 : ow) + dx;
      inb[k] = (h < H) && (w < W);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (inb[k]) v = dktb_ld4(y + (((long)img * Hi + h + in_pad) * Wi + w + in_pad) * 64 + c4);
      yv[0][k] = v.x; yv[1][k] = v.y; yv[2][k] = v.z; yv[3][k] = v.w;
    }
  float res[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int arg = -1; float gz = 0.f, xh = 0.f;
    if (full) bn_bwd_route(yv[j], np * np, mm[j], ii[j], ss[j], bb[j], gg[j], arg, gz, xh);
    for (int kk = 0; kk < np * np; ++kk) {
      const float xhat = (yv[j][kk] - mm[j]) * ii[j];
      const float gzk = (kk == arg) ? gz : 0.f;
      res[j][kk] = ss[j] * (gzk - a1[j] - xhat * a2[j]);
    }
  }
  k = 0;
  for (int dy = 0; dy < np; ++dy)
    for (int dx = 0; dx < np; ++dx, ++k) {
      if (!inb[k]) continue;
      const int h = (pool ? oh * 2 : oh) + dy, w = (pool ? ow * 2 : ow) + dx;
      dktb_st4(gy + (((long)img * Hi + h + in_pad) * Wi + w + in_pad) * 64 + c4,
               make_float4(res[0][k], res[1][k], res[2][k], res[3][k]));
    }
}

DKTB_EXPORT int dktb_bn_bwd_chunks(int H, int W, int pool) {
  const int Ho = pool ? H / 2 : H, Wo = pool ? W / 2 : W;
  return (Ho * Wo + BWD_PIX_PER_CHUNK - 1) / BWD_PIX_PER_CHUNK;
}

// Full BatchNorm+ReLU+pool backward.  partial: B*chunks*128 floats; sums: E*128 floats (kept: it is
// read by the apply pass); scratch_d: dktb_bn_scratch_doubles(E); dgamma/dbeta [64] are overwritten.
DKTB_EXPORT int dktb_bn_relu_pool_bwd(const float* y, const float* gout, const float* mean, const float* invstd,
                                      const float* gamma, const float* beta, float* gy, float* dgamma, float* dbeta,
                                      float* partial, float* sums, double* scratch_d, int B, int H, int W, int ipe,
                                      int in_pad, int out_pad, int pool, cudaStream_t stream) {
  DKTB_CHECK_ARG(y && gout && mean && invstd && gamma && beta && dgamma && dbeta && partial && sums);
  DKTB_CHECK_ARG(B > 0 && ipe > 0 && B % ipe == 0 && B <= 65535);
  const int chunks = dktb_bn_bwd_chunks(H, W, pool);
  const int E = B / ipe;
  DKTB_LAUNCH(bn_relu_pool_bwd_reduce_kernel, dim3(chunks, B), dim3(256), 0, stream, y, gout, mean, invstd, gamma, beta,
              partial, H, W, ipe, in_pad, out_pad, pool);
  DKTB_CHECK_ARG(scratch_d != nullptr);
  double* chunk = scratch_d + (long)E * 128;
  DKTB_LAUNCH(episode_sum_kernel, dim3(E, ES_CHUNKS), dim3(256), 0, stream, (const float*)partial, chunks, ipe, chunk);
  DKTB_LAUNCH(episode_sum_final_kernel, dim3(E), dim3(128), 0, stream, (const double*)chunk, sums, (double*)nullptr);
  DKTB_LAUNCH(bn_param_grad_kernel, dim3(1), dim3(64), 0, stream, (const float*)sums, E, dgamma, dbeta);
  if (gy == nullptr) return dktb_launch_status();      // sums / parameter gradients only (dktb_conv1_bwd_fused follows)
  const int Hc = pool ? (H + 1) / 2 : H, Wc = pool ? (W + 1) / 2 : W;
  const long total = (long)B * Hc * Wc * 16;
  const float inv_count = 1.0f / ((float)ipe * (float)H * (float)W);
  DKTB_LAUNCH(bn_relu_pool_bwd_apply_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, y, gout, mean,
              invstd, gamma, beta, (const float*)sums, gy, B, H, W, ipe, in_pad, out_pad, pool, inv_count);
  return dktb_launch_status();
}
