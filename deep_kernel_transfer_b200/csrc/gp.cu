// Exact-GP arithmetic of DKT (what the reference obtains from GPyTorch at methods/DKT.py:161-162,
// 177, 187, 265 and methods/DKT_regression.py:52-54, 90-93; semantics restated in SURVEY.md App. A):
//   * dktb_gram          batched X1 X2^T (the linear / cosine base kernel, cross-kernel for prediction)
//   * dktb_gp_fit        per (episode, class): K~ = s*Kb + sigma^2 I -> Cholesky -> L^-1 -> alpha, logdet,
//                        -log p/(N C); optionally K~^-1 and the gradients w.r.t. Kb, outputscale,
//                        constant mean and noise -- everything stays in shared memory
//   * dktb_gram_bwd      dZ = (S + S^T) Z with S = sum_c dL/dKb_c
//   * dktb_gp_predict    predictive mean  m_c + s_c Kx alpha_c  and the class arg-max
// One CTA owns one (episode, class) system: N <= 165 keeps L and L^-1 (2 N^2 floats) resident in the
// 227 KB of shared memory, two CTAs per SM at the reference's N = 85 / 105.  Throughput comes from
// batching E*C independent systems over the 148 SMs, not from parallelising one tiny factorisation.
#include "dktb_common.cuh"

#define NOISE_FLOOR 1e-4f   // GaussianLikelihood: noise = softplus(raw) + 1e-4

// ------------------------------------------------------------------------------------------------
// gram: out[e][m][n] = sum_d x1[e][m][d] * x2[e][n][d]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gram_kernel(const float* __restrict__ x1, const float* __restrict__ x2,
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
      acc00 = fmaf(a0, b0, acc00);
      acc01 = fmaf(a0, b1, acc01);
      acc10 = fmaf(a1, b0, acc10);
      acc11 = fmaf(a1, b1, acc11);
    }
  }
  float* o = out + (long)e * M * N;
  const int m = m0 + ty * 2, n = n0 + tx * 2;
  if (m < M && n < N) o[(long)m * N + n] = acc00;
  if (m < M && n + 1 < N) o[(long)m * N + n + 1] = acc01;
  if (m + 1 < M && n < N) o[(long)(m + 1) * N + n] = acc10;
  if (m + 1 < M && n + 1 < N) o[(long)(m + 1) * N + n + 1] = acc11;
}

DKTB_EXPORT int dktb_gram(const float* x1, const float* x2, float* out, int E, int M, int N, int D,
                          cudaStream_t stream) {
  DKTB_CHECK_ARG(x1 && x2 && out && E > 0 && M > 0 && N > 0 && D > 0 && E <= 65535);
  DKTB_LAUNCH(gram_kernel, dim3((N + 31) / 32, (M + 31) / 32, E), dim3(256), 0, stream, x1, x2, out, M, N, D);
  return dktb_launch_status();
}

// ------------------------------------------------------------------------------------------------
// gp_fit
// ------------------------------------------------------------------------------------------------
struct GpFitArgs {
  const float* kbase;        // [E][Cs][N][N], Cs = 1 (shared base kernel) or C
  long kbase_class_stride;   // 0 or N*N
  const float* y;            // targets [.][C][N]
  long y_episode_stride;     // 0 (same targets for every episode) or C*N
  const float* raw_outputscale;  // [C] or nullptr (no ScaleKernel: s = 1)
  const float* constant;     // [C]
  const float* raw_noise;    // [C]
  float* alpha;              // [E][C][N]
  float* linv;               // optional [E][C][N][N] (L^-1, lower) for predictive variances
  float* loss_terms;         // [E][C] : -log p_c / (N*C)
  int* info;                 // [E][C] : 0 ok; -k ok after the k-th jitter retry; > 0 = 1-based index of the first
                             //          non-positive pivot of the LAST attempt (not positive definite)
  float* dkbase;             // optional [E][C][N][N] : grad_scale * dLoss/dKb_c
  float* dhyper;             // optional [E][C][3] : grad_scale * d/d(raw_outputscale, constant, raw_noise)
  float grad_scale;
  float jitter;              // psd_safe_cholesky retry base: attempts add 0, j, 10 j, 100 j to the diagonal (0: one try)
  int N, C;
};

__device__ __forceinline__ float gp_block_sum(float v, float* s_red) {
  v = dktb_warp_sum(v);
  const int lane = threadIdx.x % 32, wid = threadIdx.x / 32;
  __syncthreads();
  if (lane == 0) s_red[wid] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < (int)(blockDim.x / 32); ++w) t += s_red[w];
  return t;
}

__global__ void __launch_bounds__(256) gp_fit_kernel(GpFitArgs p) {
  DKTB_DYN_SMEM(float, smem);
  const int N = p.N, C = p.C;
  const int LD = N | 1;
  float* A = smem;                    // [N][LD]  K~ then L (strict lower + raw pivots on the diagonal)
  float* X = A + (long)N * LD;        // [N][LD]  L^-1
  float* s_diag = X + (long)N * LD;   // [N]
  float* s_r = s_diag + N;            // [N]  y - m
  float* s_u = s_r + N;               // [N]
  float* s_al = s_u + N;              // [N]  alpha
  float* s_red = s_al + N;            // [32]
  __shared__ int s_fail;
  const int c = blockIdx.x, e = blockIdx.y;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const float s = p.raw_outputscale ? dktb_softplus(p.raw_outputscale[c]) : 1.f;
  const float noise = dktb_softplus(p.raw_noise[c]) + NOISE_FLOOR;
  const float mconst = p.constant[c];
  const float* kb = p.kbase + ((long)e * (p.kbase_class_stride ? C : 1)) * N * N + (long)c * p.kbase_class_stride;
  const float* yv = p.y + (long)e * p.y_episode_stride + (long)c * N;
  for (int i = tid; i < N; i += nthr) s_r[i] = yv[i] - mconst;
  // GPyTorch's psd_safe_cholesky (the reference relies on it, README.md:27): plain factorisation first; if a pivot is
  // not positive, start over with 1e-6, 1e-5, 1e-4 (fp32 schedule: jitter * 10^i) added to the diagonal of K~
  const int attempts = p.jitter > 0.f ? 4 : 1;
  int attempt = 0;
  float jit = 0.f;
  for (;; ++attempt) {
    jit = attempt == 0 ? 0.f : p.jitter * (attempt == 1 ? 1.f : attempt == 2 ? 10.f : 100.f);
    if (tid == 0) s_fail = 0;
    for (int i = tid; i < N * N; i += nthr) {
      const int r = i / N, k = i % N;
      float v = s * kb[i];
      if (r == k) v += noise + jit;
      A[r * LD + k] = v;
      X[r * LD + k] = 0.f;
    }
    __syncthreads();
    // ---- Cholesky, left-looking by columns (N <= 256: rows i = tid, tid + nthr, ...)
    for (int j = 0; j < N; ++j) {
      float t[2] = {0.f, 0.f};
      int cnt = 0;
      for (int i = j + tid; i < N; i += nthr, ++cnt) {
        float acc = A[i * LD + j];
        const float* ai = A + i * LD;
        const float* aj = A + j * LD;
        for (int k = 0; k < j; ++k) acc = fmaf(-ai[k], aj[k], acc);
        if (cnt < 2) t[cnt] = acc;
        if (i == j) A[j * LD + j] = acc;      // raw pivot; only this thread touches A[j][j]
      }
      __syncthreads();
      const float piv = A[j * LD + j];
      if (!(piv > 0.f)) {                      // uniform: every thread reads the same pivot
        if (tid == 0) s_fail = j + 1;
        break;
      }
      const float d = sqrtf(piv), invd = 1.f / d;
      cnt = 0;
      for (int i = j + tid; i < N; i += nthr, ++cnt) {
        if (i == j) s_diag[j] = d;
        else A[i * LD + j] = t[cnt] * invd;
      }
      __syncthreads();
    }
    __syncthreads();
    if (s_fail == 0 || attempt + 1 >= attempts) break;
    __syncthreads();                           // everyone has read s_fail before the next attempt resets it
  }
  const int fail = s_fail;
  if (tid == 0) p.info[(long)e * C + c] = fail ? fail : -attempt;
  if (fail) {
    // not positive definite even with jitter: NaN loss, and ZERO gradients / alpha so that nothing stale from an
    // earlier step can reach the optimiser before the host raises (the reference raises right here)
    if (tid == 0) {
      p.loss_terms[(long)e * C + c] = nanf("");
      if (p.dhyper) { float* o = p.dhyper + ((long)e * C + c) * 3; o[0] = o[1] = o[2] = 0.f; }
    }
    for (int k = tid; k < N; k += nthr) p.alpha[((long)e * C + c) * N + k] = 0.f;
    if (p.dkbase) {
      float* dk0 = p.dkbase + ((long)e * C + c) * N * N;
      for (int i = tid; i < N * N; i += nthr) dk0[i] = 0.f;
    }
    return;
  }
  // ---- X = L^-1 : thread c0 owns column c0 (independent forward substitutions, no barriers needed)
  for (int col = tid; col < N; col += nthr) {
    const int kstart = (col / 32) * 32;
    for (int i = kstart; i < N; ++i) {
      float acc = (i == col) ? 1.f : 0.f;
      const float* ai = A + i * LD;
      for (int k = kstart; k < i; ++k) acc = fmaf(-ai[k], X[k * LD + col], acc);
      X[i * LD + col] = (i >= col) ? acc / s_diag[i] : 0.f;
    }
  }
  __syncthreads();
  // ---- u = L^-1 r ; alpha = L^-T u.  sum(alpha) = 1^T K~^-1 r is accumulated as (L^-1 1) . (L^-1 r): summing the alpha_k
  // themselves cancels ~cond(K~) digits when the kernel matrix is close to rank one (RBF on similar features)
  float asum_part = 0.f;
  for (int i = tid; i < N; i += nthr) {
    float acc = 0.f, wsum = 0.f;
    for (int k = 0; k <= i; ++k) {
      acc = fmaf(X[i * LD + k], s_r[k], acc);
      wsum += X[i * LD + k];
    }
    s_u[i] = acc;
    asum_part = fmaf(wsum, acc, asum_part);
  }
  __syncthreads();
  float quad_part = 0.f, logdet_part = 0.f, aa_part = 0.f;
  for (int k = tid; k < N; k += nthr) {
    float acc = 0.f;
    for (int i = k; i < N; ++i) acc = fmaf(X[i * LD + k], s_u[i], acc);
    s_al[k] = acc;
    p.alpha[((long)e * C + c) * N + k] = acc;
    quad_part = fmaf(s_r[k], acc, quad_part);
    logdet_part += logf(s_diag[k]);
    aa_part = fmaf(acc, acc, aa_part);
  }
  const float quad = gp_block_sum(quad_part, s_red);
  const float logdet = 2.f * gp_block_sum(logdet_part, s_red);
  const float asum = gp_block_sum(asum_part, s_red);
  const float aa = gp_block_sum(aa_part, s_red);
  if (tid == 0) {
    const float logp = -0.5f * (quad + logdet + (float)N * 1.8378770664093453f);
    p.loss_terms[(long)e * C + c] = -logp / ((float)N * (float)C);
  }
  if (p.linv != nullptr) {
    float* lo = p.linv + ((long)e * C + c) * N * N;
    for (int i = tid; i < N * N; i += nthr) lo[i] = X[(i / N) * LD + (i % N)];
  }
  if (p.dkbase == nullptr && p.dhyper == nullptr) return;
  // ---- gradients: dLoss/dK~ = (K~^-1 - alpha alpha^T) * coef,  K~^-1 = X^T X
  const float coef = p.grad_scale / (2.f * (float)N * (float)C);
  float* dk = p.dkbase ? p.dkbase + ((long)e * C + c) * N * N : nullptr;
  float trinv_part = 0.f;
  for (int idx = tid; idx < N * N; idx += nthr) {
    const int i = idx / N, k = idx % N;
    const int m0 = i > k ? i : k;
    float acc = 0.f;
    for (int m = m0; m < N; ++m) acc = fmaf(X[m * LD + i], X[m * LD + k], acc);
    const float g = (acc - s_al[i] * s_al[k]) * coef;
    if (dk) dk[idx] = s * g;
    if (i == k) trinv_part += acc;
  }
  // hyper-parameter gradients in closed form from well-conditioned sums (positive terms only): with K~ = s Kb + nz I,
  //   sum_ij dK~_ij Kb_ij = [ N - nz tr(K~^-1) - r.alpha + nz alpha.alpha ] coef / s,   sum_i dK~_ii = [tr(K~^-1) - alpha.alpha] coef
  // (accumulating g_ij * Kb_ij over the N^2 entries instead loses ~cond(K~) digits)
  const float trinv = gp_block_sum(trinv_part, s_red);
  const float nz = noise + jit;
  const float ds = coef * (((float)N - nz * trinv) - (quad - nz * aa)) / s;
  const float tr = coef * (trinv - aa);
  if (tid == 0 && p.dhyper != nullptr) {
    float* o = p.dhyper + ((long)e * C + c) * 3;
    o[0] = p.raw_outputscale ? ds * dktb_sigmoid(p.raw_outputscale[c]) : 0.f;
    o[1] = -asum * p.grad_scale / ((float)N * (float)C);
    o[2] = tr * dktb_sigmoid(p.raw_noise[c]);
  }
}

DKTB_EXPORT int dktb_gp_max_n(void) { return 165; }

DKTB_EXPORT int dktb_gp_fit(const float* kbase, long kbase_class_stride, const float* y, long y_episode_stride,
                            const float* raw_outputscale, const float* constant, const float* raw_noise, float* alpha,
                            float* linv, float* loss_terms, int* info, float* dkbase, float* dhyper, float grad_scale,
                            float jitter, int E, int C, int N, cudaStream_t stream) {
  DKTB_CHECK_ARG(kbase && y && constant && raw_noise && alpha && loss_terms && info);
  DKTB_CHECK_ARG(E > 0 && C > 0 && N > 0 && N <= 165 && E <= 65535);
  GpFitArgs a;
  a.kbase = kbase; a.kbase_class_stride = kbase_class_stride; a.y = y; a.y_episode_stride = y_episode_stride;
  a.raw_outputscale = raw_outputscale; a.constant = constant; a.raw_noise = raw_noise; a.alpha = alpha; a.linv = linv;
  a.loss_terms = loss_terms; a.info = info; a.dkbase = dkbase; a.dhyper = dhyper; a.grad_scale = grad_scale;
  a.jitter = jitter; a.N = N; a.C = C;
  const int LD = N | 1;
  const size_t smem = ((size_t)2 * N * LD + 4 * (size_t)N + 32) * sizeof(float);
  cudaFuncSetAttribute(gp_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  DKTB_LAUNCH(gp_fit_kernel, dim3(C, E), dim3(256), smem, stream, a);
  return dktb_launch_status();
}

// loss[e] = sum_c loss_terms[e][c];  hyper[c][3] = sum_e dhyper[e][c][3]  (fixed order)
__global__ void gp_reduce_kernel(const float* __restrict__ loss_terms, const float* __restrict__ dhyper,
                                 float* __restrict__ loss, float* __restrict__ hyper, int E, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < E && loss != nullptr) {
    float t = 0.f;
    for (int c = 0; c < C; ++c) t += loss_terms[(long)i * C + c];
    loss[i] = t;
  }
  if (i < C * 3 && hyper != nullptr && dhyper != nullptr) {
    float t = 0.f;
    for (int e = 0; e < E; ++e) t += dhyper[(long)e * C * 3 + i];
    hyper[i] = t;
  }
}

DKTB_EXPORT int dktb_gp_reduce(const float* loss_terms, const float* dhyper, float* loss, float* hyper, int E, int C,
                               cudaStream_t stream) {
  DKTB_CHECK_ARG(loss_terms && E > 0 && C > 0);
  const int n = E > C * 3 ? E : C * 3;
  DKTB_LAUNCH(gp_reduce_kernel, dim3((n + 127) / 128), dim3(128), 0, stream, loss_terms, dhyper, loss, hyper, E, C);
  return dktb_launch_status();
}

// sticky[0] = max(sticky[0], max_i info[i]) (a failed factorisation: > 0), sticky[1] = min(sticky[1], min_i info[i])
// (deepest jitter retry: < 0).  One thread: n = E*C is tiny.  Lets the host look at the Cholesky status of MANY steps
// with one read-back instead of one per step (the reference raises at the failing step; here the failing system's
// gradients are zeroed by the kernel and the error surfaces at the next check).
__global__ void gp_info_accumulate_kernel(const int* __restrict__ info, int* __restrict__ sticky, int n) {
  int hi = sticky[0], lo = sticky[1];
  for (int i = 0; i < n; ++i) {
    const int v = info[i];
    hi = v > hi ? v : hi;
    lo = v < lo ? v : lo;
  }
  sticky[0] = hi;
  sticky[1] = lo;
}

DKTB_EXPORT int dktb_gp_info_accumulate(const int* info, int* sticky, int n, cudaStream_t stream) {
  DKTB_CHECK_ARG(info && sticky && n > 0);
  DKTB_LAUNCH(gp_info_accumulate_kernel, dim3(1), dim3(1), 0, stream, info, sticky, n);
  return dktb_launch_status();
}

// ------------------------------------------------------------------------------------------------
// gram_bwd: dZ[e][n][d] = scale * sum_m (S[n][m] + S[m][n]) Z[e][m][d],  S = sum_c W[e][c]
// grid (d groups, ceil(N/32), E); 256 threads: (ty: 2 rows) x (tx: 4 columns).  The symmetrised 32-row block of S is
// built ONCE per CTA in shared memory (it was rebuilt for every 64-wide slice of D before: the class sum and its
// transposed reads dominated the kernel), then the CTA walks its share of the D / 64 slices.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gram_bwd_kernel(const float* __restrict__ w, const float* __restrict__ z,
                                                       float* __restrict__ dz, int C, int N, int D, float scale,
                                                       int Np) {
  DKTB_DYN_SMEM(float, smem);
  float* s_s = smem;                                   // [32][Np + 1]
  float* s_z = smem + ((32 * (Np + 1) + 3) & ~3);      // [32][64]
  const int e = blockIdx.z;
  const int n0 = blockIdx.y * 32;
  const int tid = threadIdx.x;
  const int ty = tid / 16, tx = tid % 16;
  const float* we = w + (long)e * C * N * N;
  const float* ze = z + (long)e * N * D;
  for (int i = tid; i < 32 * Np; i += 256) {
    const int r = i / Np, m = i - r * Np;
    const int n = n0 + r;
    float v = 0.f;
    if (n < N && m < N)
      for (int c = 0; c < C; ++c) v += we[((long)c * N + n) * N + m];
    s_s[r * (Np + 1) + m] = v;
  }
  __syncthreads();
  for (int i = tid; i < 32 * Np; i += 256) {            // + transpose: lanes along the row index -> coalesced reads
    const int m = i / 32, r = i - m * 32;
    const int n = n0 + r;
    if (n < N && m < N) {
      float v = 0.f;
      for (int c = 0; c < C; ++c) v += we[((long)c * N + m) * N + n];
      s_s[r * (Np + 1) + m] += v;
    }
  }
  const int nslice = (D + 63) / 64;
  float* o = dz + (long)e * N * D;
  for (int sl = blockIdx.x; sl < nslice; sl += gridDim.x) {
    const int d0 = sl * 64;
    float acc[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int m0 = 0; m0 < N; m0 += 32) {
      __syncthreads();
      for (int i = tid; i < 32 * 64; i += 256) {
        const int r = i / 64, cc = i % 64;
        const int m = m0 + r, d = d0 + cc;
        s_z[r * 64 + cc] = (m < N && d < D) ? ze[(long)m * D + d] : 0.f;
      }
      __syncthreads();
      const float* sa = s_s + (ty * 2) * (Np + 1) + m0;
#pragma unroll 8
      for (int k = 0; k < 32; ++k) {
        const float a0 = sa[k], a1 = sa[Np + 1 + k];
        const float4 b = dktb_ld4(&s_z[k * 64 + tx * 4]);
        acc[0][0] = fmaf(a0, b.x, acc[0][0]); acc[0][1] = fmaf(a0, b.y, acc[0][1]);
        acc[0][2] = fmaf(a0, b.z, acc[0][2]); acc[0][3] = fmaf(a0, b.w, acc[0][3]);
        acc[1][0] = fmaf(a1, b.x, acc[1][0]); acc[1][1] = fmaf(a1, b.y, acc[1][1]);
        acc[1][2] = fmaf(a1, b.z, acc[1][2]); acc[1][3] = fmaf(a1, b.w, acc[1][3]);
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int n = n0 + ty * 2 + i;
      if (n >= N) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int d = d0 + tx * 4 + j;
        if (d < D) o[(long)n * D + d] = scale * acc[i][j];
      }
    }
  }
}

DKTB_EXPORT int dktb_gram_bwd(const float* w, const float* z, float* dz, int E, int C, int N, int D, float scale,
                              cudaStream_t stream) {
  DKTB_CHECK_ARG(w && z && dz && E > 0 && C > 0 && N > 0 && D > 0 && E <= 65535);
  const int Np = (N + 31) / 32 * 32;
  const size_t smem = (size_t)(((32 * (Np + 1) + 3) & ~3) + 32 * 64) * sizeof(float);
  DKTB_CHECK_ARG(smem <= 200 * 1024);
  // enough CTAs for two per SM, at most one per 64-wide slice of D
  const int rows = (N + 31) / 32, nslice = (D + 63) / 64;
  int groups = (2 * 148 + rows * E - 1) / (rows * E);
  groups = groups < 1 ? 1 : (groups > nslice ? nslice : groups);
  cudaFuncSetAttribute(gram_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  DKTB_LAUNCH(gram_bwd_kernel, dim3(groups, rows, E), dim3(256), smem, stream, w, z, dz, C, N, D, scale, Np);
  return dktb_launch_status();
}

// ------------------------------------------------------------------------------------------------
// predict: mean[e][c][m] = const_c + s_c * sum_n kx[e][(c)][m][n] * alpha[e][c][n]
//          pred[e][m]    = argmax_c sigmoid(mean[e][c][m])   (first index on ties, DKT.py:266-269)
// one block per episode, one warp per test point
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gp_predict_kernel(const float* __restrict__ kx, long kx_class_stride,
                                                         const float* __restrict__ alpha,
                                                         const float* __restrict__ raw_outputscale,
                                                         const float* __restrict__ constant, float* __restrict__ mean,
                                                         int* __restrict__ pred, int C, int M, int N) {
  const int e = blockIdx.x;
  const int lane = threadIdx.x % 32, wid = threadIdx.x / 32, nw = blockDim.x / 32;
  const float* kxe = kx + (long)e * (kx_class_stride ? C : 1) * M * N;
  for (int m = wid; m < M; m += nw) {
    float best = 0.f;
    int arg = 0;
    for (int c = 0; c < C; ++c) {
      const float* row = kxe + (long)c * kx_class_stride + (long)m * N;
      const float* al = alpha + ((long)e * C + c) * N;
      float acc = 0.f;
      for (int n = lane; n < N; n += 32) acc = fmaf(row[n], al[n], acc);
      acc = dktb_warp_sum(acc);
      const float s = raw_outputscale ? dktb_softplus(raw_outputscale[c]) : 1.f;
      const float mu = fmaf(s, acc, constant[c]);
      if (lane == 0) mean[((long)e * C + c) * M + m] = mu;
      const float sg = dktb_sigmoid(mu);
      if (c == 0 || sg > best) { best = sg; arg = c; }
    }
    if (lane == 0 && pred != nullptr) pred[(long)e * M + m] = arg;
  }
}

DKTB_EXPORT int dktb_gp_predict(const float* kx, long kx_class_stride, const float* alpha,
                                const float* raw_outputscale, const float* constant, float* mean, int* pred, int E,
                                int C, int M, int N, cudaStream_t stream) {
  DKTB_CHECK_ARG(kx && alpha && constant && mean && E > 0 && C > 0 && M > 0 && N > 0);
  DKTB_LAUNCH(gp_predict_kernel, dim3(E), dim3(256), 0, stream, kx, kx_class_stride, alpha, raw_outputscale, constant,
              mean, pred, C, M, N);
  return dktb_launch_status();
}
