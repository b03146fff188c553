// Base-kernel family of the reference's ExactGPLayer (methods/DKT.py:352-372; methods/DKT_regression.py:117-124),
// evaluated as an element-wise epilogue on the Gram matrix of the (mean-centred) features, forward and backward:
//   kind 0 linear   k = v * g                          v = softplus(raw_variance)
//   kind 1 rbf      k = exp(-d2/2)                     d2 = ||x_i - x_j||^2 / l^2,  l = softplus(raw_lengthscale)
//   kind 2 matern   k = (1 + sqrt5 d + 5/3 d2) exp(-sqrt5 d)   (nu = 2.5, GPyTorch default), d = sqrt(max(d2, 1e-30))
//   kind 3 poli1    k = g + off                        off = softplus(raw_offset)
//   kind 4 poli2    k = (g + off)^2
// g is the Gram matrix x1 x2^T (linear / polynomial kernels); rbf / matern take the squared distances from
// dktb_sqdist, which accumulates (a-b)^2 directly: GPyTorch instead expands ||a||^2+||b||^2-2ab after centring both
// inputs on x1's mean and zeroes the diagonal -- the direct form needs neither and has no cancellation.
// dktb_center_rows is kept for callers that want the reference's centring (the kernels are translation invariant).
// One kernel matrix per class (every one-vs-rest model has its own hyper-parameter): kb [E][C][M][N].
// Backward: dkb = dLoss/dKb_c -> dg [E][N][N] (sum over classes, including the terms that reach the diagonal through
// the norms n_i = g_ii) and the raw hyper-parameter gradient per (episode, class).
#include "dktb_common.cuh"

#define KIND_LINEAR 0
#define KIND_RBF 1
#define KIND_MATERN 2
#define KIND_POLI1 3
#define KIND_POLI2 4

// x [E][N][D] -> x - mean_over_rows(ref) ; ref [E][Nr][D] supplies the mean (x itself for the training inputs)
__global__ void __launch_bounds__(256) center_rows_kernel(const float* __restrict__ x, const float* __restrict__ ref,
                                                          float* __restrict__ out, int N, int Nr, int D) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int e = blockIdx.y;
  if (j >= D) return;
  const float* re = ref + (long)e * Nr * D;
  float s = 0.f;
  for (int n = 0; n < Nr; ++n) s += re[(long)n * D + j];
  const float m = s / (float)Nr;
  const float* xe = x + (long)e * N * D;
  float* oe = out + (long)e * N * D;
  for (int n = 0; n < N; ++n) oe[(long)n * D + j] = xe[(long)n * D + j] - m;
}

DKTB_EXPORT int dktb_center_rows(const float* x, const float* ref, float* out, int E, int N, int Nr, int D,
                                 cudaStream_t stream) {
  DKTB_CHECK_ARG(x && ref && out && E > 0 && N > 0 && Nr > 0 && D > 0);
  DKTB_LAUNCH(center_rows_kernel, dim3((D + 255) / 256, E), dim3(256), 0, stream, x, ref, out, N, Nr, D);
  return dktb_launch_status();
}

// sq[e][n] = ||x[e][n]||^2 ; one warp per row
__global__ void __launch_bounds__(256) row_sqnorm_kernel(const float* __restrict__ x, float* __restrict__ sq, long rows,
                                                         int D) {
  const long row = (long)blockIdx.x * 8 + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  float s = 0.f;
  if (row < rows)
    for (int j = lane; j < D; j += 32) s = fmaf(x[row * D + j], x[row * D + j], s);
  s = dktb_warp_sum(s);
  if (row < rows && lane == 0) sq[row] = s;
}

DKTB_EXPORT int dktb_row_sqnorm(const float* x, float* sq, long rows, int D, cudaStream_t stream) {
  DKTB_CHECK_ARG(x && sq && rows > 0 && D > 0);
  DKTB_LAUNCH(row_sqnorm_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, stream, x, sq, rows, D);
  return dktb_launch_status();
}

// d2[e][m][n] = || x1[e][m] - x2[e][n] ||^2 accumulated directly (no ||a||^2 + ||b||^2 - 2ab cancellation; the diagonal
// of a symmetric call is exactly zero).  Same tiling as the Gram kernel.
__global__ void __launch_bounds__(256) sqdist_kernel(const float* __restrict__ x1, const float* __restrict__ x2,
                                                     float* __restrict__ out, int M, int N, int D) {
  __shared__ float s1[32][33];
  __shared__ float s2[32][33];
  const int e = blockIdx.z;
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int tid = threadIdx.x;
  const int lr = tid / 32, lc = tid % 32;
  const int ty = tid / 16, tx = tid % 16;
  const float* a = x1 + (long)e * M * D;
  const float* b = x2 + (long)e * N * D;
  float acc00 = 0.f, acc01 = 0.f, acc10 = 0.f, acc11 = 0.f;
  for (int k0 = 0; k0 < D; k0 += 32) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = lr + 8 * i;
      const int k = k0 + lc;
      s1[r][lc] = (m0 + r < M && k < D) ? a[(long)(m0 + r) * D + k] : 0.f;
      s2[r][lc] = (n0 + r < N && k < D) ? b[(long)(n0 + r) * D + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float a0 = s1[ty * 2][k], a1 = s1[ty * 2 + 1][k];
      const float b0 = s2[tx * 2][k], b1 = s2[tx * 2 + 1][k];
      acc00 = fmaf(a0 - b0, a0 - b0, acc00);
      acc01 = fmaf(a0 - b1, a0 - b1, acc01);
      acc10 = fmaf(a1 - b0, a1 - b0, acc10);
      acc11 = fmaf(a1 - b1, a1 - b1, acc11);
    }
  }
  float* o = out + (long)e * M * N;
  const int m = m0 + ty * 2, n = n0 + tx * 2;
  if (m < M && n < N) o[(long)m * N + n] = acc00;
  if (m < M && n + 1 < N) o[(long)m * N + n + 1] = acc01;
  if (m + 1 < M && n < N) o[(long)(m + 1) * N + n] = acc10;
  if (m + 1 < M && n + 1 < N) o[(long)(m + 1) * N + n + 1] = acc11;
}

DKTB_EXPORT int dktb_sqdist(const float* x1, const float* x2, float* out, int E, int M, int N, int D,
                            cudaStream_t stream) {
  DKTB_CHECK_ARG(x1 && x2 && out && E > 0 && M > 0 && N > 0 && D > 0 && E <= 65535);
  DKTB_LAUNCH(sqdist_kernel, dim3((N + 31) / 32, (M + 31) / 32, E), dim3(256), 0, stream, x1, x2, out, M, N, D);
  return dktb_launch_status();
}

__device__ __forceinline__ float kfam_eval(int kind, float g, float raw_d2, float p) {
  if (kind == KIND_LINEAR) return p * g;
  if (kind == KIND_POLI1) return g + p;
  if (kind == KIND_POLI2) return (g + p) * (g + p);
  const float d2 = fmaxf(raw_d2, 0.f) / (p * p);
  if (kind == KIND_RBF) return expf(-0.5f * d2);
  const float d = sqrtf(fmaxf(d2, 1e-30f));
  const float s5 = 2.2360679774997896f;
  return (s5 * d + 1.f + (5.f / 3.f) * d * d) * expf(-s5 * d);
}

// kb[e][c][m][n] = k_c(g[e][m][n]);  grid (ceil(M*N/256), C, E)
__global__ void __launch_bounds__(256) kernel_fwd_kernel(int kind, const float* __restrict__ g,
                                                         const float* __restrict__ d2m,
                                                         const float* __restrict__ raw_param, float* __restrict__ kb,
                                                         int C, int M, int N) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y, e = blockIdx.z;
  if (idx >= M * N) return;
  const float p = dktb_softplus(raw_param[c]);
  const float gv = g ? g[(long)e * M * N + idx] : 0.f;
  const float dv = d2m ? d2m[(long)e * M * N + idx] : 0.f;
  kb[(((long)e * C + c) * M) * N + idx] = kfam_eval(kind, gv, dv, p);
}

// g: Gram matrix (linear / poli kernels), d2: squared distances from dktb_sqdist (rbf / matern); the other may be NULL
DKTB_EXPORT int dktb_kernel_fwd(int kind, const float* g, const float* d2, const float* raw_param, float* kb, int E,
                                int C, int M, int N, cudaStream_t stream) {
  DKTB_CHECK_ARG(raw_param && kb && E > 0 && C > 0 && M > 0 && N > 0 && kind >= 0 && kind <= 4);
  DKTB_CHECK_ARG((kind == KIND_RBF || kind == KIND_MATERN) ? (d2 != nullptr) : (g != nullptr));
  DKTB_LAUNCH(kernel_fwd_kernel, dim3((M * N + 255) / 256, C, E), dim3(256), 0, stream, kind, g, d2, raw_param, kb, C, M,
              N);
  return dktb_launch_status();
}

// Backward over the symmetric training kernel (M == N, sq = diag).  One CTA per (row i, episode e):
//   dg[e][i][j] = sum_c dkb_c[i][j] * dk/dg_ij           (j != i)
//   dg[e][i][i] = sum_c ( dkb_c[i][i] * dk/dg_ii + 2/l^2 * sum_j A_c[i][j] )   with A = dkb * dk/dd2 (rbf/matern)
//   dparam_rows[e][c][i] = sum_j dkb_c[i][j] * dk/dparam
__global__ void __launch_bounds__(128) kernel_bwd_kernel(int kind, const float* __restrict__ g,
                                                         const float* __restrict__ d2m,
                                                         const float* __restrict__ raw_param,
                                                         const float* __restrict__ dkb, float* __restrict__ dg,
                                                         float* __restrict__ dparam_rows, int C, int N) {
  __shared__ float s_red[4];
  const int i = blockIdx.x, e = blockIdx.y;
  const int tid = threadIdx.x, lane = tid % 32, wid = tid / 32;
  const float* ge = g ? g + (long)e * N * N + (long)i * N : nullptr;
  const float* de = d2m ? d2m + (long)e * N * N + (long)i * N : nullptr;
  float diag_extra = 0.f;
  for (int j = tid; j < N; j += 128) dg[(long)e * N * N + (long)i * N + j] = 0.f;
  __syncthreads();
  for (int c = 0; c < C; ++c) {
    const float p = dktb_softplus(raw_param[c]);
    const float* we = dkb + (((long)e * C + c) * N + i) * N;
    float dp = 0.f, rowA = 0.f;
    for (int j = tid; j < N; j += 128) {
      const float w = we[j], gij = ge ? ge[j] : 0.f;
      float dkdg = 0.f;
      if (kind == KIND_LINEAR) {
        dkdg = p;
        dp = fmaf(w, gij, dp);
      } else if (kind == KIND_POLI1) {
        dkdg = 1.f;
        dp += w;
      } else if (kind == KIND_POLI2) {
        dkdg = 2.f * (gij + p);
        dp = fmaf(w, 2.f * (gij + p), dp);
      } else {
        const float raw = de[j];
        const float d2 = fmaxf(raw, 0.f) / (p * p);
        float dkdd2;
        if (kind == KIND_RBF) {
          dkdd2 = -0.5f * expf(-0.5f * d2);
        } else {
          const float d = sqrtf(fmaxf(d2, 1e-30f));
          const float s5 = 2.2360679774997896f;
          dkdd2 = -(5.f / 6.f) * (1.f + s5 * d) * expf(-s5 * d);
        }
        const float a = (raw > 0.f) ? w * dkdd2 : 0.f;       // clamp_min(0) passes no gradient below zero
        dkdg = 0.f;
        dg[(long)e * N * N + (long)i * N + j] += -2.f * a / (p * p);
        rowA += a;
        dp = fmaf(a, -2.f * d2 / p, dp);
      }
      if (kind == KIND_LINEAR || kind == KIND_POLI1 || kind == KIND_POLI2)
        dg[(long)e * N * N + (long)i * N + j] += w * dkdg;
    }
    // block reductions of dp and rowA
    dp = dktb_warp_sum(dp);
    rowA = dktb_warp_sum(rowA);
    __syncthreads();
    if (lane == 0) s_red[wid] = dp;
    __syncthreads();
    const float dpt = s_red[0] + s_red[1] + s_red[2] + s_red[3];
    __syncthreads();
    if (lane == 0) s_red[wid] = rowA;
    __syncthreads();
    const float rat = s_red[0] + s_red[1] + s_red[2] + s_red[3];
    if (tid == 0) {
      dparam_rows[((long)e * C + c) * N + i] = dpt * dktb_sigmoid(raw_param[c]);
      diag_extra += 2.f * rat / (p * p);
    }
  }
  __syncthreads();
  if (tid == 0 && (kind == KIND_RBF || kind == KIND_MATERN)) dg[(long)e * N * N + (long)i * N + i] += diag_extra;
}

// dparam[c] = sum_e sum_i dparam_rows[e][c][i]   (fixed order)
__global__ void kernel_bwd_reduce_kernel(const float* __restrict__ rows, float* __restrict__ dparam, int E, int C,
                                         int N) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float t = 0.f;
  for (int e = 0; e < E; ++e)
    for (int n = 0; n < N; ++n) t += rows[((long)e * C + c) * N + n];
  dparam[c] = t;
}

// dparam [C] = dLoss/d raw_param summed over episodes; scratch: E*C*N floats
DKTB_EXPORT int dktb_kernel_bwd(int kind, const float* g, const float* d2, const float* raw_param, const float* dkb,
                                float* dg, float* dparam, float* scratch, int E, int C, int N, cudaStream_t stream) {
  DKTB_CHECK_ARG(raw_param && dkb && dg && dparam && scratch && E > 0 && C > 0 && N > 0 && kind >= 0 && kind <= 4);
  DKTB_CHECK_ARG((kind == KIND_RBF || kind == KIND_MATERN) ? (d2 != nullptr) : (g != nullptr));
  DKTB_LAUNCH(kernel_bwd_kernel, dim3(N, E), dim3(128), 0, stream, kind, g, d2, raw_param, dkb, dg, scratch, C, N);
  DKTB_LAUNCH(kernel_bwd_reduce_kernel, dim3((C + 31) / 32), dim3(32), 0, stream, (const float*)scratch, dparam, E, C,
              N);
  return dktb_launch_status();
}

// Predictive variance of likelihood(model(x*)) (methods/DKT_regression.py:90-93, confidence_region):
//   var[e][c][m] = s_c * kss[e][c][m] - || L_c^-1 (s_c kx[e][c][m][:]) ||^2 + noise_c
// linv [E][C][N][N] (lower) from dktb_gp_fit; kx [E][(C)][M][N] base cross kernel; kss [E][(C)][M] base k(x*,x*).
__global__ void __launch_bounds__(128) gp_predict_var_kernel(const float* __restrict__ kx, long kx_class_stride,
                                                             const float* __restrict__ kss, long kss_class_stride,
                                                             const float* __restrict__ linv,
                                                             const float* __restrict__ raw_outputscale,
                                                             const float* __restrict__ raw_noise,
                                                             float* __restrict__ var, int C, int M, int N) {
  __shared__ float s_red[4];
  const int m = blockIdx.x, c = blockIdx.y, e = blockIdx.z;
  const int tid = threadIdx.x, lane = tid % 32, wid = tid / 32;
  const float s = raw_outputscale ? dktb_softplus(raw_outputscale[c]) : 1.f;
  const float noise = dktb_softplus(raw_noise[c]) + 1e-4f;
  const float* row = kx + ((long)e * (kx_class_stride ? C : 1)) * M * N + (long)c * kx_class_stride + (long)m * N;
  const float* L = linv + ((long)e * C + c) * N * N;
  float acc = 0.f;
  for (int i = tid; i < N; i += 128) {       // v_i = sum_{k<=i} Linv[i][k] * s*kx[k]
    float v = 0.f;
    for (int k = 0; k <= i; ++k) v = fmaf(L[(long)i * N + k], row[k], v);
    v *= s;
    acc = fmaf(v, v, acc);
  }
  acc = dktb_warp_sum(acc);
  if (lane == 0) s_red[wid] = acc;
  __syncthreads();
  if (tid == 0) {
    const float q = s_red[0] + s_red[1] + s_red[2] + s_red[3];
    const float kd = kss[((long)e * (kss_class_stride ? C : 1)) * M + (long)c * kss_class_stride + m];
    var[((long)e * C + c) * M + m] = s * kd - q + noise;
  }
}

DKTB_EXPORT int dktb_gp_predict_var(const float* kx, long kx_class_stride, const float* kss, long kss_class_stride,
                                    const float* linv, const float* raw_outputscale, const float* raw_noise, float* var,
                                    int E, int C, int M, int N, cudaStream_t stream) {
  DKTB_CHECK_ARG(kx && kss && linv && raw_noise && var && E > 0 && C > 0 && M > 0 && N > 0);
  DKTB_LAUNCH(gp_predict_var_kernel, dim3(M, C, E), dim3(128), 0, stream, kx, kx_class_stride, kss, kss_class_stride,
              linv, raw_outputscale, raw_noise, var, C, M, N);
  return dktb_launch_status();
}
