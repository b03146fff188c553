// dktb_gp_fit_large: the same contract as dktb_gp_fit (gp.cu) for systems that do not fit in shared memory
// (165 < N <= 512: 20-way training episodes, the Gram-N sweep of BASELINE.json configs[4]; reference call sites
// methods/DKT.py:161-163 through GPyTorch's ExactMarginalLogLikelihood, semantics SURVEY.md App. A).
// One CTA per (episode, class) system again, but K~ / L and L^-1 live in a caller-provided global workspace
// (2 N^2 floats per system, L2-resident: 2 MB at N = 500) and the factorisation is blocked left-looking with 32-wide
// column panels: the panel update is a register-tiled (rows x 32 x j0) product whose row operand is read as whole
// 128-byte lines, the diagonal block is factored by one warp in shared memory, the rest of the panel by a row-per-
// thread triangular solve.  L^-1, alpha, K~^-1 and the gradients follow gp_fit's formulas on the global arrays.
#include "dktb_common.cuh"

#define NOISE_FLOOR 1e-4f
#define GL_NB 32
#define GL_THREADS 256
#define GL_MAXN 512

struct GpFitLargeArgs {
  const float* kbase;
  long kbase_class_stride;
  const float* y;
  long y_episode_stride;
  const float* raw_outputscale;
  const float* constant;
  const float* raw_noise;
  float* alpha;
  float* linv;
  float* loss_terms;
  int* info;
  float* dkbase;
  float* dhyper;
  float* work;            // [E][C][2][N][N]
  float grad_scale, jitter;
  int N, C;
};

__device__ __forceinline__ float gl_block_sum(float v, float* s_red) {
  v = dktb_warp_sum(v);
  const int lane = threadIdx.x % 32, wid = threadIdx.x / 32;
  __syncthreads();
  if (lane == 0) s_red[wid] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < GL_THREADS / 32; ++w) t += s_red[w];
  return t;
}

__global__ void __launch_bounds__(GL_THREADS) gp_fit_large_kernel(GpFitLargeArgs p) {
  __shared__ float s_t[GL_NB][GL_NB + 1];      // L[j0.., k0..] tile, then the factored diagonal block
  __shared__ float s_diag[GL_MAXN];
  __shared__ float s_r[GL_MAXN];
  __shared__ float s_u[GL_MAXN];
  __shared__ float s_al[GL_MAXN];
  __shared__ float s_red[32];
  __shared__ int s_fail;
  const int N = p.N, C = p.C;
  const int c = blockIdx.x, e = blockIdx.y;
  const int tid = threadIdx.x;
  const float s = p.raw_outputscale ? dktb_softplus(p.raw_outputscale[c]) : 1.f;
  const float noise = dktb_softplus(p.raw_noise[c]) + NOISE_FLOOR;
  const float mconst = p.constant[c];
  const float* kb = p.kbase + ((long)e * (p.kbase_class_stride ? C : 1)) * N * N + (long)c * p.kbase_class_stride;
  const float* yv = p.y + (long)e * p.y_episode_stride + (long)c * N;
  float* A = p.work + ((long)e * C + c) * 2 * N * N;     // K~ -> L (lower incl. diagonal)
  float* X = A + (long)N * N;                             // L^-1 (lower)
  if (tid == 0) s_fail = 0;
  for (int i = tid; i < N * N; i += GL_THREADS) {
    const int r = i / N, k = i % N;
    float v = s * kb[i];
    if (r == k) v += noise + p.jitter;
    A[i] = v;
  }
  for (int i = tid; i < N; i += GL_THREADS) s_r[i] = yv[i] - mconst;
  __syncthreads();

  // ---- blocked left-looking Cholesky
  for (int j0 = 0; j0 < N; j0 += GL_NB) {
    const int nb = min(GL_NB, N - j0);
    const int R = N - j0;                              // panel rows j0 .. N-1
    // each thread owns panel rows r = tid, tid + 256 (N <= 512): 32 accumulators per row
    float acc[2][GL_NB];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = tid + h * GL_THREADS;
#pragma unroll
      for (int cc = 0; cc < GL_NB; ++cc) acc[h][cc] = (r < R && cc < nb) ? A[(long)(j0 + r) * N + j0 + cc] : 0.f;
    }
    for (int k0 = 0; k0 < j0; k0 += GL_NB) {           // j0 is a multiple of 32, so full tiles
      __syncthreads();
      for (int i = tid; i < GL_NB * GL_NB; i += GL_THREADS) {
        const int rr = i / GL_NB, kk = i % GL_NB;
        s_t[rr][kk] = (rr < nb) ? A[(long)(j0 + rr) * N + k0 + kk] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = tid + h * GL_THREADS;
        if (r >= R) continue;
        const float* arow = A + (long)(j0 + r) * N + k0;
        float reg[GL_NB];
#pragma unroll
        for (int kk = 0; kk < GL_NB; ++kk) reg[kk] = arow[kk];
#pragma unroll
        for (int cc = 0; cc < GL_NB; ++cc) {
          float a = acc[h][cc];
#pragma unroll
          for (int kk = 0; kk < GL_NB; ++kk) a = fmaf(-reg[kk], s_t[cc][kk], a);
          acc[h][cc] = a;
        }
      }
    }
    __syncthreads();
    // diagonal block -> shared memory, factored by warp 0 (lane = row)
    if (tid < GL_NB) {
#pragma unroll
      for (int cc = 0; cc < GL_NB; ++cc) s_t[tid][cc] = acc[0][cc];
    }
    __syncthreads();
    if (tid < 32) {
      const int lane = tid;
      for (int j = 0; j < nb; ++j) {
        float v = 0.f;
        if (lane >= j && lane < nb) {
          v = s_t[lane][j];
          for (int k = 0; k < j; ++k) v = fmaf(-s_t[lane][k], s_t[j][k], v);
        }
        const float piv = __shfl_sync(0xffffffffu, v, j);
        if (!(piv > 0.f)) {
          if (lane == 0) s_fail = j0 + j + 1;
          break;
        }
        const float d = sqrtf(piv);
        if (lane == j) { s_t[j][j] = d; s_diag[j0 + j] = d; }
        else if (lane > j && lane < nb) s_t[lane][j] = v / d;
        __syncwarp();
      }
    }
    __syncthreads();
    if (s_fail) break;
    // panel rows: row r < nb is the diagonal block itself; r >= nb solved against it
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = tid + h * GL_THREADS;
      if (r >= R) continue;
      float* arow = A + (long)(j0 + r) * N + j0;
      if (r < nb) {
        for (int cc = 0; cc <= r; ++cc) arow[cc] = s_t[r][cc];
      } else {
#pragma unroll
        for (int cc = 0; cc < GL_NB; ++cc) {
          if (cc < nb) {
            float v = acc[h][cc];
#pragma unroll
            for (int k = 0; k < GL_NB; ++k)
              if (k < cc) v = fmaf(-acc[h][k], s_t[cc][k], v);
            v = v / s_t[cc][cc];
            acc[h][cc] = v;
            arow[cc] = v;
          }
        }
      }
    }
    __syncthreads();
  }
  __syncthreads();
  const int fail = s_fail;
  if (tid == 0) p.info[(long)e * C + c] = fail;
  if (fail) {
    if (tid == 0) p.loss_terms[(long)e * C + c] = nanf("");
    return;
  }
  // ---- X = L^-1 : thread owns a column (forward substitution; reads of L broadcast, X coalesced)
  for (int col = tid; col < N; col += GL_THREADS) {
    const int kstart = (col / 32) * 32;
    for (int i = 0; i < kstart; ++i) X[(long)i * N + col] = 0.f;
    for (int i = kstart; i < N; ++i) {
      float a = (i == col) ? 1.f : 0.f;
      const float* ai = A + (long)i * N;
      for (int k = kstart; k < i; ++k) a = fmaf(-ai[k], X[(long)k * N + col], a);
      X[(long)i * N + col] = (i >= col) ? a / s_diag[i] : 0.f;
    }
  }
  __syncthreads();
  for (int i = tid; i < N; i += GL_THREADS) {
    float a = 0.f;
    for (int k = 0; k <= i; ++k) a = fmaf(X[(long)i * N + k], s_r[k], a);
    s_u[i] = a;
  }
  __syncthreads();
  float quad_part = 0.f, logdet_part = 0.f, asum_part = 0.f;
  for (int k = tid; k < N; k += GL_THREADS) {
    float a = 0.f;
    for (int i = k; i < N; ++i) a = fmaf(X[(long)i * N + k], s_u[i], a);
    s_al[k] = a;
    p.alpha[((long)e * C + c) * N + k] = a;
    quad_part = fmaf(s_r[k], a, quad_part);
    logdet_part += logf(s_diag[k]);
    asum_part += a;
  }
  const float quad = gl_block_sum(quad_part, s_red);
  const float logdet = 2.f * gl_block_sum(logdet_part, s_red);
  const float asum = gl_block_sum(asum_part, s_red);
  if (tid == 0) {
    const float logp = -0.5f * (quad + logdet + (float)N * 1.8378770664093453f);
    p.loss_terms[(long)e * C + c] = -logp / ((float)N * (float)C);
  }
  if (p.linv != nullptr) {
    float* lo = p.linv + ((long)e * C + c) * N * N;
    for (int i = tid; i < N * N; i += GL_THREADS) lo[i] = X[i];
  }
  if (p.dkbase == nullptr && p.dhyper == nullptr) return;
  const float coef = p.grad_scale / (2.f * (float)N * (float)C);
  float* dk = p.dkbase ? p.dkbase + ((long)e * C + c) * N * N : nullptr;
  float ds_part = 0.f, tr_part = 0.f;
  for (int idx = tid; idx < N * N; idx += GL_THREADS) {
    const int i = idx / N, k = idx % N;
    const int m0 = i > k ? i : k;
    float a = 0.f;
    for (int m = m0; m < N; ++m) a = fmaf(X[(long)m * N + i], X[(long)m * N + k], a);
    const float g = (a - s_al[i] * s_al[k]) * coef;
    if (dk) dk[idx] = s * g;
    ds_part = fmaf(g, kb[idx], ds_part);
    if (i == k) tr_part += g;
  }
  const float ds = gl_block_sum(ds_part, s_red);
  const float tr = gl_block_sum(tr_part, s_red);
  if (tid == 0 && p.dhyper != nullptr) {
    float* o = p.dhyper + ((long)e * C + c) * 3;
    o[0] = p.raw_outputscale ? ds * dktb_sigmoid(p.raw_outputscale[c]) : 0.f;
    o[1] = -asum * p.grad_scale / ((float)N * (float)C);
    o[2] = tr * dktb_sigmoid(p.raw_noise[c]);
  }
}

DKTB_EXPORT int dktb_gp_large_max_n(void) { return GL_MAXN; }
DKTB_EXPORT long dktb_gp_large_work_floats(int E, int C, int N) { return (long)E * C * 2 * N * N; }

// Same arguments as dktb_gp_fit plus `work` (dktb_gp_large_work_floats(E, C, N) floats).
DKTB_EXPORT int dktb_gp_fit_large(const float* kbase, long kbase_class_stride, const float* y, long y_episode_stride,
                                  const float* raw_outputscale, const float* constant, const float* raw_noise,
                                  float* alpha, float* linv, float* loss_terms, int* info, float* dkbase, float* dhyper,
                                  float* work, float grad_scale, float jitter, int E, int C, int N,
                                  cudaStream_t stream) {
  DKTB_CHECK_ARG(kbase && y && constant && raw_noise && alpha && loss_terms && info && work);
  DKTB_CHECK_ARG(E > 0 && C > 0 && N > 0 && N <= GL_MAXN && E <= 65535);
  GpFitLargeArgs a;
  a.kbase = kbase; a.kbase_class_stride = kbase_class_stride; a.y = y; a.y_episode_stride = y_episode_stride;
  a.raw_outputscale = raw_outputscale; a.constant = constant; a.raw_noise = raw_noise; a.alpha = alpha; a.linv = linv;
  a.loss_terms = loss_terms; a.info = info; a.dkbase = dkbase; a.dhyper = dhyper; a.work = work;
  a.grad_scale = grad_scale; a.jitter = jitter; a.N = N; a.C = C;
  DKTB_LAUNCH(gp_fit_large_kernel, dim3(C, E), dim3(GL_THREADS), 0, stream, a);
  return dktb_launch_status();
}
