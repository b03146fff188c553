// Feature head of the bncossim / cossim kernels: "bn_out" BatchNorm1d over the flattened features
// (methods/DKT.py:45-48) with per-episode batch statistics, followed by F.normalize(p=2, dim=1,
// eps=1e-12) (methods/DKT.py:142, 175, 185, 237, 263), forward and backward.
//
// Features are held NHWC-flattened ([pixel][channel], j = p*Cch + c) while the reference flattens
// NCHW (index c*P + p, backbone.py:46-51): the bn_out parameter/buffer of feature j lives at
// pidx(j) = (j % Cch)*P + j / Cch in the reference-ordered tensors, so checkpoints stay compatible
// and the Gram matrix (a sum over features) is unchanged.
#include "dktb_common.cuh"

__device__ __forceinline__ int head_pidx(int j, int Cch, int P) { return P <= 1 ? j : (j % Cch) * P + j / Cch; }

// grid (ceil(D/256), E).  training: batch stats over the N rows of episode e; else running stats.
__global__ void __launch_bounds__(256) bn1d_fwd_kernel(const float* __restrict__ f, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta,
                                                       const float* __restrict__ running_mean,
                                                       const float* __restrict__ running_var, float* __restrict__ z,
                                                       float* __restrict__ mean, float* __restrict__ invstd,
                                                       float* __restrict__ var, int N, int D, int Cch, int P,
                                                       int training, float eps) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int e = blockIdx.y;
  if (j >= D) return;
  const int pj = head_pidx(j, Cch, P);
  const float* fe = f + (long)e * N * D;
  float m, is;
  if (training) {
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += fe[(long)n * D + j];
    m = s / (float)N;
    float v = 0.f;
    for (int n = 0; n < N; ++n) {
      const float d = fe[(long)n * D + j] - m;
      v = fmaf(d, d, v);
    }
    v /= (float)N;
    is = 1.f / sqrtf(v + eps);
    mean[(long)e * D + j] = m;
    invstd[(long)e * D + j] = is;
    var[(long)e * D + j] = v;
  } else {
    m = running_mean[pj];
    is = 1.f / sqrtf(running_var[pj] + eps);
    if (mean) mean[(long)e * D + j] = m;
    if (invstd) invstd[(long)e * D + j] = is;
  }
  const float sc = gamma[pj] * is, bt = beta[pj];
  float* ze = z + (long)e * N * D;
  for (int n = 0; n < N; ++n) ze[(long)n * D + j] = fmaf(fe[(long)n * D + j] - m, sc, bt);
}

__global__ void bn1d_running_kernel(const float* __restrict__ mean, const float* __restrict__ var,
                                    float* __restrict__ running_mean, float* __restrict__ running_var, int E, int N,
                                    int D, int Cch, int P, float momentum) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= D) return;
  const int pj = head_pidx(j, Cch, P);
  float rm = running_mean[pj], rv = running_var[pj];
  const float unb = N > 1 ? (float)N / (float)(N - 1) : 1.f;
  for (int e = 0; e < E; ++e) {
    rm = (1.f - momentum) * rm + momentum * mean[(long)e * D + j];
    rv = (1.f - momentum) * rv + momentum * var[(long)e * D + j] * unb;
  }
  running_mean[pj] = rm;
  running_var[pj] = rv;
}

DKTB_EXPORT int dktb_bn1d_fwd(const float* f, const float* gamma, const float* beta, float* running_mean,
                              float* running_var, float* z, float* mean, float* invstd, float* var, int E, int N,
                              int D, int Cch, int P, int training, int update_running, float momentum, float eps,
                              cudaStream_t stream) {
  DKTB_CHECK_ARG(f && gamma && beta && z && E > 0 && N > 0 && D > 0 && running_mean && running_var);
  DKTB_CHECK_ARG(!training || (mean && invstd && var));
  DKTB_CHECK_ARG(P <= 1 || Cch * P == D);
  DKTB_LAUNCH(bn1d_fwd_kernel, dim3((D + 255) / 256, E), dim3(256), 0, stream, f, gamma, beta,
              (const float*)running_mean, (const float*)running_var, z, mean, invstd, var, N, D, Cch, P, training, eps);
  if (training && update_running)
    DKTB_LAUNCH(bn1d_running_kernel, dim3((D + 255) / 256), dim3(256), 0, stream, (const float*)mean,
                (const float*)var, running_mean, running_var, E, N, D, Cch, P, momentum);
  return dktb_launch_status();
}

// block-wide sum (256 threads); every thread gets the result
__device__ __forceinline__ float block_sum_256(float v, float* s_buf) {
  v = dktb_warp_sum(v);
  const int lane = threadIdx.x % 32, wid = threadIdx.x / 32;
  __syncthreads();
  if (lane == 0) s_buf[wid] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += s_buf[w];
  return t;
}

// one block per row: zhat = z / max(||z||, eps); inv[row] = 1/max(||z||, eps)
__global__ void __launch_bounds__(256) l2norm_fwd_kernel(const float* __restrict__ z, float* __restrict__ zhat,
                                                         float* __restrict__ inv, int D, float eps) {
  __shared__ float s_buf[8];
  const long row = blockIdx.x;
  const float* zr = z + row * D;
  float s = 0.f;
  for (int j = threadIdx.x; j < D; j += 256) s = fmaf(zr[j], zr[j], s);
  const float tot = block_sum_256(s, s_buf);
  const float iv = 1.f / fmaxf(sqrtf(tot), eps);
  for (int j = threadIdx.x; j < D; j += 256) zhat[row * D + j] = zr[j] * iv;
  if (threadIdx.x == 0 && inv) inv[row] = iv;
}

DKTB_EXPORT int dktb_l2norm_fwd(const float* z, float* zhat, float* inv, long rows, int D, float eps,
                                cudaStream_t stream) {
  DKTB_CHECK_ARG(z && zhat && rows > 0 && D > 0);
  DKTB_LAUNCH(l2norm_fwd_kernel, dim3((unsigned)rows), dim3(256), 0, stream, z, zhat, inv, D, eps);
  return dktb_launch_status();
}

// g_z = inv * ( g_zhat - zhat * <zhat, g_zhat> )
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const float* __restrict__ zhat, const float* __restrict__ gzhat,
                                                         const float* __restrict__ inv, float* __restrict__ gz, int D) {
  __shared__ float s_buf[8];
  const long row = blockIdx.x;
  float s = 0.f;
  for (int j = threadIdx.x; j < D; j += 256) s = fmaf(zhat[row * D + j], gzhat[row * D + j], s);
  const float dot = block_sum_256(s, s_buf);
  const float iv = inv[row];
  for (int j = threadIdx.x; j < D; j += 256) gz[row * D + j] = iv * (gzhat[row * D + j] - zhat[row * D + j] * dot);
}

DKTB_EXPORT int dktb_l2norm_bwd(const float* zhat, const float* gzhat, const float* inv, float* gz, long rows, int D,
                                cudaStream_t stream) {
  DKTB_CHECK_ARG(zhat && gzhat && inv && gz && rows > 0 && D > 0);
  DKTB_LAUNCH(l2norm_bwd_kernel, dim3((unsigned)rows), dim3(256), 0, stream, zhat, gzhat, inv, gz, D);
  return dktb_launch_status();
}

// BatchNorm1d backward (train mode), per (episode, feature) column; pgrad[e][2][D] = {sum g, sum g*xhat}
__global__ void __launch_bounds__(256) bn1d_bwd_kernel(const float* __restrict__ f, const float* __restrict__ gz,
                                                       const float* __restrict__ gamma, const float* __restrict__ mean,
                                                       const float* __restrict__ invstd, float* __restrict__ gf,
                                                       float* __restrict__ pgrad, int N, int D, int Cch, int P) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int e = blockIdx.y;
  if (j >= D) return;
  const int pj = head_pidx(j, Cch, P);
  const float m = mean[(long)e * D + j], is = invstd[(long)e * D + j];
  const float* fe = f + (long)e * N * D;
  const float* ge = gz + (long)e * N * D;
  float s1 = 0.f, s2 = 0.f;
  for (int n = 0; n < N; ++n) {
    const float g = ge[(long)n * D + j];
    s1 += g;
    s2 = fmaf(g, (fe[(long)n * D + j] - m) * is, s2);
  }
  pgrad[((long)e * 2 + 0) * D + j] = s1;
  pgrad[((long)e * 2 + 1) * D + j] = s2;
  const float sc = gamma[pj] * is, a1 = s1 / (float)N, a2 = s2 / (float)N;
  float* oe = gf + (long)e * N * D;
  for (int n = 0; n < N; ++n) {
    const float xh = (fe[(long)n * D + j] - m) * is;
    oe[(long)n * D + j] = sc * (ge[(long)n * D + j] - a1 - xh * a2);
  }
}

__global__ void bn1d_param_grad_kernel(const float* __restrict__ pgrad, int E, int D, int Cch, int P,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= D) return;
  const int pj = head_pidx(j, Cch, P);
  float a = 0.f, b = 0.f;
  for (int e = 0; e < E; ++e) {
    b += pgrad[((long)e * 2 + 0) * D + j];
    a += pgrad[((long)e * 2 + 1) * D + j];
  }
  dgamma[pj] = a;
  dbeta[pj] = b;
}

DKTB_EXPORT int dktb_bn1d_bwd(const float* f, const float* gz, const float* gamma, const float* mean,
                              const float* invstd, float* gf, float* dgamma, float* dbeta, float* pgrad, int E, int N,
                              int D, int Cch, int P, cudaStream_t stream) {
  DKTB_CHECK_ARG(f && gz && gamma && mean && invstd && gf && dgamma && dbeta && pgrad && E > 0 && N > 0 && D > 0);
  DKTB_LAUNCH(bn1d_bwd_kernel, dim3((D + 255) / 256, E), dim3(256), 0, stream, f, gz, gamma, mean, invstd, gf, pgrad, N,
              D, Cch, P);
  DKTB_LAUNCH(bn1d_param_grad_kernel, dim3((D + 255) / 256), dim3(256), 0, stream, (const float*)pgrad, E, D, Cch, P,
              dgamma, dbeta);
  return dktb_launch_status();
}
