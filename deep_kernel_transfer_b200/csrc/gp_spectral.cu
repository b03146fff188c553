// Spectral-mixture kernel with ARD (GPyTorch SpectralMixtureKernel(num_mixtures=Q, ard_num_dims=D); reference
// methods/DKT_regression.py:122 with Q=4, D=2916 and sines/train_DKT.py:132 with Q=4, D=40), forward and backward:
//     k(x, x') = sum_q w_q * prod_d exp(-2 pi^2 tau_d^2 v_qd^2) * cos(2 pi tau_d mu_qd),   tau = x - x'
// with w = softplus(raw_mixture_weights) [Q], mu = softplus(raw_mixture_means) [Q][D], v = softplus(raw_mixture_scales).
// The feature index j runs over the NHWC-flattened features; parameters are stored in the reference's NCHW-flatten
// order (index (j % Cch)*P + j / Cch, identity when P <= 1), like bn_out.
// Direct O(Q * N * M * D) evaluation: one CTA per (pair, episode); N <= 19 in the reference, so this is tiny.
#include "dktb_common.cuh"

#define SM_TWO_PI 6.283185307179586f
#define SM_TWO_PI2 19.739208802178716f   // 2 pi^2
#define SM_MAXQ 8

__device__ __forceinline__ int sm_pidx(int j, int Cch, int P) { return P <= 1 ? j : (j % Cch) * P + j / Cch; }

__device__ __forceinline__ float block_reduce_sum_128(float v, float* s_buf) {
  v = dktb_warp_sum(v);
  __syncthreads();
  if (threadIdx.x % 32 == 0) s_buf[threadIdx.x / 32] = v;
  __syncthreads();
  return (s_buf[0] + s_buf[1]) + (s_buf[2] + s_buf[3]);
}
__device__ __forceinline__ float block_reduce_prod_128(float v, float* s_buf) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (threadIdx.x % 32 == 0) s_buf[threadIdx.x / 32] = v;
  __syncthreads();
  return (s_buf[0] * s_buf[1]) * (s_buf[2] * s_buf[3]);
}

// kb[e][m][n] = k(x1[e][m], x2[e][n]);  ec[e][q][m][n] = E_q * C_q (kept for the backward pass, nullable)
__global__ void __launch_bounds__(128) spectral_fwd_kernel(const float* __restrict__ x1, const float* __restrict__ x2,
                                                           const float* __restrict__ raw_w,
                                                           const float* __restrict__ raw_mu,
                                                           const float* __restrict__ raw_v, float* __restrict__ kb,
                                                           float* __restrict__ ec, int M, int N, int D, int Q, int Cch,
                                                           int P) {
  __shared__ float s_buf[4];
  const int n = blockIdx.x % N, m = blockIdx.x / N, e = blockIdx.y;
  const float* a = x1 + ((long)e * M + m) * D;
  const float* b = x2 + ((long)e * N + n) * D;
  float k = 0.f;
  for (int q = 0; q < Q; ++q) {
    float s = 0.f, pr = 1.f;
    for (int j = threadIdx.x; j < D; j += 128) {
      const int pj = sm_pidx(j, Cch, P);
      const float tau = a[j] - b[j];
      const float v = dktb_softplus(raw_v[(long)q * D + pj]), mu = dktb_softplus(raw_mu[(long)q * D + pj]);
      s = fmaf(tau * tau, v * v, s);
      pr *= cosf(SM_TWO_PI * tau * mu);
    }
    const float st = block_reduce_sum_128(s, s_buf);
    const float pt = block_reduce_prod_128(pr, s_buf);
    const float val = expf(-SM_TWO_PI2 * st) * pt;
    if (threadIdx.x == 0 && ec != nullptr) ec[(((long)e * Q + q) * M + m) * N + n] = val;
    k = fmaf(dktb_softplus(raw_w[q]), val, k);
  }
  if (threadIdx.x == 0) kb[((long)e * M + m) * N + n] = k;
}

DKTB_EXPORT int dktb_spectral_fwd(const float* x1, const float* x2, const float* raw_w, const float* raw_mu,
                                  const float* raw_v, float* kb, float* ec, int E, int M, int N, int D, int Q, int Cch,
                                  int P, cudaStream_t stream) {
  DKTB_CHECK_ARG(x1 && x2 && raw_w && raw_mu && raw_v && kb && E > 0 && M > 0 && N > 0 && D > 0 && Q > 0 && Q <= SM_MAXQ);
  DKTB_LAUNCH(spectral_fwd_kernel, dim3(M * N, E), dim3(128), 0, stream, x1, x2, raw_w, raw_mu, raw_v, kb, ec, M, N, D,
              Q, Cch, P);
  return dktb_launch_status();
}

// d(prod_d' cos_d')/d(cos_d) = product of the other factors: C/cos_d when cos_d is not tiny, else recomputed
__device__ float sm_prod_excluding(const float* a, const float* b, const float* raw_mu_q, int D, int Cch, int P, int skip) {
  float pr = 1.f;
  for (int j = 0; j < D; ++j) {
    if (j == skip) continue;
    const float mu = dktb_softplus(raw_mu_q[sm_pidx(j, Cch, P)]);
    pr *= cosf(SM_TWO_PI * (a[j] - b[j]) * mu);
  }
  return pr;
}

// Parameter gradients: thread per (q, j): dmu[q][pj], dv[q][pj] (already times sigmoid(raw)); dw[q] by thread j == 0.
// dkb [E][N][N] = dLoss/dK of the training kernel (x1 == x2 == x).
__global__ void __launch_bounds__(128) spectral_bwd_param_kernel(const float* __restrict__ x,
                                                                 const float* __restrict__ raw_w,
                                                                 const float* __restrict__ raw_mu,
                                                                 const float* __restrict__ raw_v,
                                                                 const float* __restrict__ dkb,
                                                                 const float* __restrict__ ec, float* __restrict__ dw,
                                                                 float* __restrict__ dmu, float* __restrict__ dv, int E,
                                                                 int N, int D, int Q, int Cch, int P) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int q = blockIdx.y;
  if (j >= D) return;
  const int pj = sm_pidx(j, Cch, P);
  const float w = dktb_softplus(raw_w[q]);
  const float mu = dktb_softplus(raw_mu[(long)q * D + pj]), v = dktb_softplus(raw_v[(long)q * D + pj]);
  float gmu = 0.f, gv = 0.f, gw = 0.f;
  for (int e = 0; e < E; ++e)
    for (int m = 0; m < N; ++m)
      for (int n = 0; n < N; ++n) {
        const float wt = dkb[((long)e * N + m) * N + n];
        const float val = ec[(((long)e * Q + q) * N + m) * N + n];
        if (j == 0) gw = fmaf(wt, val, gw);
        const float* a = x + ((long)e * N + m) * D;
        const float* b = x + ((long)e * N + n) * D;
        const float tau = a[j] - b[j];
        gv = fmaf(wt * w * val, -2.f * SM_TWO_PI2 * tau * tau * v, gv);
        const float arg = SM_TWO_PI * tau * mu;
        const float c = cosf(arg);
        float dC;       // d(E*C)/d mu_j = E * (-2 pi tau sin(arg)) * prod_{d != j} cos
        if (fabsf(c) > 1e-6f) {
          dC = val * (-SM_TWO_PI * tau * sinf(arg) / c);
        } else {
          float s = 0.f;
          for (int d = 0; d < D; ++d) {
            const float vv = dktb_softplus(raw_v[(long)q * D + sm_pidx(d, Cch, P)]);
            s = fmaf((a[d] - b[d]) * (a[d] - b[d]), vv * vv, s);
          }
          dC = expf(-SM_TWO_PI2 * s) * (-SM_TWO_PI * tau * sinf(arg)) *
               sm_prod_excluding(a, b, raw_mu + (long)q * D, D, Cch, P, j);
        }
        gmu = fmaf(wt * w, dC, gmu);
      }
  dmu[(long)q * D + pj] = gmu * dktb_sigmoid(raw_mu[(long)q * D + pj]);
  dv[(long)q * D + pj] = gv * dktb_sigmoid(raw_v[(long)q * D + pj]);
  if (j == 0) dw[q] = gw * dktb_sigmoid(raw_w[q]);
}

// Input gradient: thread per (n, j): dx[e][n][j] = sum_m (dkb[n][m] + dkb[m][n]) * dk(x_n, x_m)/d x_n,j
__global__ void __launch_bounds__(128) spectral_bwd_x_kernel(const float* __restrict__ x, const float* __restrict__ raw_w,
                                                             const float* __restrict__ raw_mu,
                                                             const float* __restrict__ raw_v,
                                                             const float* __restrict__ dkb,
                                                             const float* __restrict__ ec, float* __restrict__ dx, int N,
                                                             int D, int Q, int Cch, int P) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y, e = blockIdx.z;
  if (j >= D) return;
  const int pj = sm_pidx(j, Cch, P);
  const float* a = x + ((long)e * N + n) * D;
  float g = 0.f;
  for (int m = 0; m < N; ++m) {
    if (m == n) continue;                       // tau = 0: both derivative terms vanish
    const float* b = x + ((long)e * N + m) * D;
    const float wt = dkb[((long)e * N + n) * N + m] + dkb[((long)e * N + m) * N + n];
    const float tau = a[j] - b[j];
    for (int q = 0; q < Q; ++q) {
      const float w = dktb_softplus(raw_w[q]);
      const float mu = dktb_softplus(raw_mu[(long)q * D + pj]), v = dktb_softplus(raw_v[(long)q * D + pj]);
      const float val = ec[(((long)e * Q + q) * N + n) * N + m];
      const float arg = SM_TWO_PI * tau * mu;
      const float c = cosf(arg);
      float term = val * (-2.f * SM_TWO_PI2 * tau * v * v);
      if (fabsf(c) > 1e-6f) {
        term += val * (-SM_TWO_PI * mu * sinf(arg) / c);
      } else {
        float s = 0.f;
        for (int d = 0; d < D; ++d) {
          const float vv = dktb_softplus(raw_v[(long)q * D + sm_pidx(d, Cch, P)]);
          s = fmaf((a[d] - b[d]) * (a[d] - b[d]), vv * vv, s);
        }
        term += expf(-SM_TWO_PI2 * s) * (-SM_TWO_PI * mu * sinf(arg)) *
                sm_prod_excluding(a, b, raw_mu + (long)q * D, D, Cch, P, j);
      }
      g = fmaf(wt * w, term, g);
    }
  }
  dx[((long)e * N + n) * D + j] = g;
}

// dkb [E][N][N], ec [E][Q][N][N] from the forward; dw [Q], dmu / dv [Q][D] (reference order), dx [E][N][D]
DKTB_EXPORT int dktb_spectral_bwd(const float* x, const float* raw_w, const float* raw_mu, const float* raw_v,
                                  const float* dkb, const float* ec, float* dw, float* dmu, float* dv, float* dx, int E,
                                  int N, int D, int Q, int Cch, int P, cudaStream_t stream) {
  DKTB_CHECK_ARG(x && raw_w && raw_mu && raw_v && dkb && ec && dw && dmu && dv && dx && E > 0 && N > 0 && D > 0 && Q > 0);
  DKTB_LAUNCH(spectral_bwd_param_kernel, dim3((D + 127) / 128, Q), dim3(128), 0, stream, x, raw_w, raw_mu, raw_v, dkb, ec,
              dw, dmu, dv, E, N, D, Q, Cch, P);
  DKTB_LAUNCH(spectral_bwd_x_kernel, dim3((D + 127) / 128, N, E), dim3(128), 0, stream, x, raw_w, raw_mu, raw_v, dkb, ec,
              dx, N, D, Q, Cch, P);
  return dktb_launch_status();
}
