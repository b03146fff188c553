// dktb_gp_fit_large: the same contract as dktb_gp_fit (gp.cu) for systems that do not fit in shared memory
// (165 < N <= 512: 20-way training episodes, the Gram-N sweep of BASELINE.json configs[4]; reference call sites
// methods/DKT.py:161-163 through GPyTorch's ExactMarginalLogLikelihood, semantics SURVEY.md App. A).
// One CTA per (episode, class) system; K~ / L and L^-1 live in a caller-provided global workspace (2 N^2 floats per
// system, L2-resident: 2 MB at N = 500).  Everything cubic is a 64 x 64 x k tile product run by the whole CTA out of
// shared memory (4 x 4 outputs per thread, operands staged k-contiguous with a 36 / 68-float pitch so the 16-byte
// operand reads are bank-conflict free, the next k-chunk prefetched into registers while the current one is used):
//   per 64-wide block column j:  diagonal tile  D   = K~_jj - L_j,<j L_j,<j^T      (tile product)
//                                factor D = L_jj L_jj^T (one warp, shared memory), W = L_jj^-1 (64 threads)
//                                panel tiles    L_ij = (K~_ij - L_i,<j L_j,<j^T) W^T   (two tile products)
//                                block row j of L^-1:  X_jk = -W sum_{k<=m<j} L_jm X_mk (two tile products)
//   K~^-1 = X^T X tile by tile (lower tiles, mirrored through shared memory), fused with the gradient epilogue.
// Multiplying by the inverted 64 x 64 diagonal block replaces the row-by-row triangular solves (K~ = s K + sigma^2 I
// keeps those blocks well conditioned).  alpha, log-determinant, loss and hyper-parameter gradients follow gp_fit.
#include <stdlib.h>

#include "tile_mma.cuh"

#define NOISE_FLOOR 1e-4f
#define GL_LDD 65            // pitch of the diagonal block while it is factored (row per lane)
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

struct GlSmem {
  float* as;      // [2][64][36]
  float* bs;      // [2][64][36]
  float* d;       // [64][65]  diagonal block / its Cholesky factor (shares the storage of t)
  float* w;       // [64][68]  inverse of the factor
  float* t;       // [64][68]  a tile handed from one product to the next
  float* diag;    // [512]
  float* r;       // [512] y - m
  float* u;       // [512] L^-1 r
  float* al;      // [512] alpha
  float* red;     // [32]
  int* fail;
};
#define GL_SMEM_FLOATS (4 * GL_T * GL_LDK + 2 * GL_T * GL_LDT + 4 * GL_MAXN + 32 + 4)

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

// v - sum_{k < n} a[k * sa] * b[k * sb] out of shared memory: loads issued eight at a time, two independent FMA chains
// (a plain dependent loop pays the shared-memory latency on every step: profiles/r01_gp_fit_large.summary.txt).
__device__ __forceinline__ float gl_dot_sub(float v, const float* a, int sa, const float* b, int sb, int n) {
  float v2 = 0.f;
  int k = 0;
  for (; k + 8 <= n; k += 8) {
    float x[8], y[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { x[q] = a[(k + q) * sa]; y[q] = b[(k + q) * sb]; }
#pragma unroll
    for (int q = 0; q < 8; q += 2) { v = fmaf(-x[q], y[q], v); v2 = fmaf(-x[q + 1], y[q + 1], v2); }
  }
  for (; k < n; ++k) v = fmaf(-a[k * sa], b[k * sb], v);
  return v + v2;
}

template <bool MMA>
__global__ void __launch_bounds__(GL_THREADS, 2) gp_fit_large_kernel(GpFitLargeArgs p) {
  typedef GlMap<MMA> M;
  DKTB_DYN_SMEM(float, smem);
  GlSmem sm;
  sm.as = smem;
  sm.bs = sm.as + 2 * GL_T * GL_LDK;
  sm.w = sm.bs + 2 * GL_T * GL_LDK;
  sm.t = sm.w + GL_T * GL_LDT;
  sm.d = sm.t;                                           // the diagonal block is dead once L_jj / W are written out
  sm.diag = sm.t + GL_T * GL_LDT;
  sm.r = sm.diag + GL_MAXN;
  sm.u = sm.r + GL_MAXN;
  sm.al = sm.u + GL_MAXN;
  sm.red = sm.al + GL_MAXN;
  sm.fail = reinterpret_cast<int*>(sm.red + 32);
  const int N = p.N, C = p.C;
  const int c = blockIdx.x, e = blockIdx.y;
  const int tid = threadIdx.x;
  const float s = p.raw_outputscale ? dktb_softplus(p.raw_outputscale[c]) : 1.f;
  const float noise = dktb_softplus(p.raw_noise[c]) + NOISE_FLOOR;
  const float mconst = p.constant[c];
  const float* kb = p.kbase + ((long)e * (p.kbase_class_stride ? C : 1)) * N * N + (long)c * p.kbase_class_stride;
  const float* yv = p.y + (long)e * p.y_episode_stride + (long)c * N;
  float* A = p.work + ((long)e * C + c) * 2 * N * N;     // K~ -> L (lower incl. diagonal)
  float* X = A + (long)N * N;                             // L^-1 (lower; entries above the diagonal blocks unused)
  const int nT = (N + GL_T - 1) / GL_T;
  float acc[4][4];
  // psd_safe_cholesky (GPyTorch; README.md:27): plain factorisation, then retries with jitter * {1, 10, 100} added to
  // the diagonal (the fp32 schedule 1e-6, 1e-5, 1e-4 when the caller passes 1e-6; jitter == 0: a single attempt)
  const int attempts = p.jitter > 0.f ? 4 : 1;
  int attempt = 0;
  float jit = 0.f;
  for (;; ++attempt) {
  jit = attempt == 0 ? 0.f : p.jitter * (attempt == 1 ? 1.f : attempt == 2 ? 10.f : 100.f);
  if (tid == 0) *sm.fail = 0;
  for (int i0 = tid; i0 < N * N; i0 += 4 * GL_THREADS) {     // four independent loads in flight per thread
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = (i0 + q * GL_THREADS < N * N) ? kb[i0 + q * GL_THREADS] : 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = i0 + q * GL_THREADS;
      if (i < N * N) A[i] = s * v[q];
    }
  }
  __syncthreads();
  for (int i = tid; i < N; i += GL_THREADS) {
    A[(long)i * N + i] += noise + jit;
    sm.r[i] = yv[i] - mconst;
  }
  __syncthreads();

  // ---- blocked Cholesky (left-looking by 64-wide block columns) with the block rows of L^-1 produced on the way
  for (int jb = 0; jb < nT; ++jb) {
    const int j0 = jb * GL_T, nb = min(GL_T, N - j0);
    // diagonal tile
    gl_zero(acc);
    {
      GlRows la{A, N, j0, nb, j0}, lb{A, N, j0, nb, j0};
      gl_product<MMA>(acc, sm.as, sm.bs, la, lb, 0, j0);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = M::row(i), cc = M::col(j);
        sm.d[r * GL_LDD + cc] = (r < nb && cc < nb) ? A[(long)(j0 + r) * N + j0 + cc] - acc[i][j] : (r == cc ? 1.f : 0.f);
      }
    __syncthreads();
    // right-looking Cholesky of the 64 x 64 block by the whole CTA: per column the pivot, the scaled column (one thread
    // per row) and the rank-1 update of the trailing lower triangle (thread = row x 4-way interleaved columns)
    {
      float* colv = sm.as;                                // staging area is idle here
      const int ur = tid >> 2, uq = tid & 3;
      for (int j = 0; j < nb; ++j) {
        const float piv = sm.d[j * GL_LDD + j];
        if (!(piv > 0.f)) {                               // uniform: every thread reads the same pivot
          if (tid == 0) *sm.fail = j0 + j + 1;
          break;
        }
        const float dg = sqrtf(piv);
        if (tid > j && tid < nb) colv[tid] = sm.d[tid * GL_LDD + j] / dg;
        __syncthreads();
        if (tid == j) { sm.d[j * GL_LDD + j] = dg; sm.diag[j0 + j] = dg; }
        else if (tid > j && tid < nb) sm.d[tid * GL_LDD + j] = colv[tid];
        if (ur > j && ur < nb) {
          const float lr = colv[ur];
          float* dr = sm.d + ur * GL_LDD;
          for (int cc = j + 1 + uq; cc <= ur; cc += 4) dr[cc] = fmaf(-lr, colv[cc], dr[cc]);
        }
        __syncthreads();
      }
    }
    __syncthreads();
    if (*sm.fail) break;
    // W = L_jj^-1 in two 32-wide halves: the diagonal halves by forward substitution (a thread per column), the
    // off-diagonal block W21 = -W22 (L21 W11) as two small products by the whole CTA; rows beyond nb form an identity
    if (tid < GL_T) {
      const int col = tid, lo = col & 32;
      for (int i = 0; i < GL_T; ++i) {
        float a = 0.f;
        if (i >= col && i < lo + 32) {
          const float* di = sm.d + i * GL_LDD;
          a = gl_dot_sub((i == col) ? 1.f : 0.f, di + col, 1, sm.w + col * GL_LDT + col, GL_LDT, i - col) / di[i];
        }
        if (i < lo + 32) sm.w[i * GL_LDT + col] = a;      // rows 32.. of the first 32 columns come from the products
      }
    }
    __syncthreads();
    {
      float* tt = sm.as;                                  // T = L21 W11, [32][33]
      const int r = tid >> 3, c0 = tid & 7;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int cc = c0 + 8 * q;
        tt[r * 33 + cc] = -gl_dot_sub(0.f, sm.d + (32 + r) * GL_LDD + cc, 1, sm.w + cc * GL_LDT + cc, GL_LDT, 32 - cc);
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int cc = c0 + 8 * q;
        sm.w[(32 + r) * GL_LDT + cc] = gl_dot_sub(0.f, sm.w + (32 + r) * GL_LDT + 32, 1, tt + cc, 33, r + 1);
      }
    }
    __syncthreads();
    // L_jj and X_jj to global (zeros above the diagonal inside the block)
    for (int i = tid; i < GL_T * GL_T; i += GL_THREADS) {
      const int r = i >> 6, cc = i & 63;
      if (r < nb && cc < nb) {
        if (cc <= r) A[(long)(j0 + r) * N + j0 + cc] = sm.d[r * GL_LDD + cc];
        X[(long)(j0 + r) * N + j0 + cc] = sm.w[r * GL_LDT + cc];
      }
    }
    // panel tiles below the diagonal: L_ij = (K~_ij - L_i,<j L_j,<j^T) W^T
    for (int ib = jb + 1; ib < nT; ++ib) {
      const int i0 = ib * GL_T, nr = min(GL_T, N - i0);
      gl_zero(acc);
      {
        GlRows la{A, N, i0, nr, j0}, lb{A, N, j0, nb, j0};
        gl_product<MMA>(acc, sm.as, sm.bs, la, lb, 0, j0);
      }
      __syncthreads();                                    // sm.t free (previous tile's second product done)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = M::row(i), cc = M::col(j);
          sm.t[r * GL_LDT + cc] = (r < nr && cc < nb) ? A[(long)(i0 + r) * N + j0 + cc] - acc[i][j] : 0.f;
        }
      __syncthreads();
      gl_zero(acc);
      gl_tile<MMA>(acc, sm.t, GL_LDT, sm.w, GL_LDT, GL_T);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = M::row(i), cc = M::col(j);
          if (r < nr && cc < nb) A[(long)(i0 + r) * N + j0 + cc] = acc[i][j];
        }
    }
    __syncthreads();
    // block row jb of X = L^-1: X_jk = -W * sum_{m = k .. j-1} L_jm X_mk   (block indices), k < jb
    for (int kbk = 0; kbk < jb; ++kbk) {
      const int k0 = kbk * GL_T;
      gl_zero(acc);
      {
        GlRows la{A, N, j0, nb, j0};
        GlCols lb{X, N, k0, GL_T, j0};
        gl_product<MMA>(acc, sm.as, sm.bs, la, lb, k0, j0);
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) sm.t[M::col(j) * GL_LDT + M::row(i)] = acc[i][j];     // transposed: t[c][m]
      __syncthreads();
      gl_zero(acc);
      gl_tile<MMA>(acc, sm.w, GL_LDT, sm.t, GL_LDT, GL_T);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = M::row(i), cc = M::col(j);
          if (r < nb) X[(long)(j0 + r) * N + k0 + cc] = -acc[i][j];
        }
    }
    __syncthreads();
  }
  __syncthreads();
  if (*sm.fail == 0 || attempt + 1 >= attempts) break;
  __syncthreads();                                        // everyone has read the flag before the next attempt resets it
  }
  const int fail = *sm.fail;
  if (tid == 0) p.info[(long)e * C + c] = fail ? fail : -attempt;
  if (fail) {      // NaN loss, zero gradients / alpha: nothing stale reaches the optimiser before the host raises
    if (tid == 0) {
      p.loss_terms[(long)e * C + c] = nanf("");
      if (p.dhyper) { float* o = p.dhyper + ((long)e * C + c) * 3; o[0] = o[1] = o[2] = 0.f; }
    }
    for (int k = tid; k < N; k += GL_THREADS) p.alpha[((long)e * C + c) * N + k] = 0.f;
    if (p.dkbase) {
      float* dk0 = p.dkbase + ((long)e * C + c) * N * N;
      for (int i = tid; i < N * N; i += GL_THREADS) dk0[i] = 0.f;
    }
    return;
  }
  // ---- u = L^-1 r (a warp per row), alpha = L^-T u (a thread per column, coalesced down the rows).  sum(alpha) =
  // 1^T K~^-1 r is accumulated as (L^-1 1) . (L^-1 r): summing the alpha_k cancels ~cond(K~) digits (see gp.cu)
  float asum_part = 0.f;
  {
    const int lane = tid & 31, warp = tid >> 5;
    for (int i = warp; i < N; i += GL_THREADS / 32) {
      float a = 0.f, w1 = 0.f;
      for (int k = lane; k <= i; k += 32) {
        const float xv = X[(long)i * N + k];
        a = fmaf(xv, sm.r[k], a);
        w1 += xv;
      }
      a = dktb_warp_sum(a);
      w1 = dktb_warp_sum(w1);
      if (lane == 0) {
        sm.u[i] = a;
        asum_part = fmaf(w1, a, asum_part);
      }
    }
  }
  __syncthreads();
  float quad_part = 0.f, logdet_part = 0.f, aa_part = 0.f;
  for (int k = tid; k < N; k += GL_THREADS) {
    float a = 0.f;
    for (int i = k; i < N; ++i) a = fmaf(X[(long)i * N + k], sm.u[i], a);
    sm.al[k] = a;
    p.alpha[((long)e * C + c) * N + k] = a;
    quad_part = fmaf(sm.r[k], a, quad_part);
    logdet_part += logf(sm.diag[k]);
    aa_part = fmaf(a, a, aa_part);
  }
  const float quad = gl_block_sum(quad_part, sm.red);
  const float logdet = 2.f * gl_block_sum(logdet_part, sm.red);
  const float asum = gl_block_sum(asum_part, sm.red);
  const float aa = gl_block_sum(aa_part, sm.red);
  if (tid == 0) {
    const float logp = -0.5f * (quad + logdet + (float)N * 1.8378770664093453f);
    p.loss_terms[(long)e * C + c] = -logp / ((float)N * (float)C);
  }
  if (p.linv != nullptr) {
    float* lo = p.linv + ((long)e * C + c) * N * N;
    for (int i = tid; i < N * N; i += GL_THREADS) {
      const int r = i / N, k = i - r * N;
      lo[i] = k <= r ? X[i] : 0.f;
    }
  }
  if (p.dkbase == nullptr && p.dhyper == nullptr) return;
  // ---- K~^-1 = X^T X, lower tiles; gradient epilogue on the tile and (through sm.t) on its mirror image
  const float coef = p.grad_scale / (2.f * (float)N * (float)C);
  float* dk = p.dkbase ? p.dkbase + ((long)e * C + c) * N * N : nullptr;
  float trinv_part = 0.f;
  for (int ib = 0; ib < nT; ++ib) {
    const int i0 = ib * GL_T, ni = min(GL_T, N - i0);
    for (int kbk = 0; kbk <= ib; ++kbk) {
      const int k0 = kbk * GL_T;
      gl_zero(acc);
      {
        GlCols la{X, N, i0, ni, N}, lb{X, N, k0, GL_T, N};       // kbk <= ib: the k-tile is full unless it is the last
        lb.nq = min(GL_T, N - k0);
        gl_product<MMA>(acc, sm.as, sm.bs, la, lb, i0, N);
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = M::row(i), cc = M::col(j);
          const int gi = i0 + r, gk = k0 + cc;
          float g = 0.f;
          if (gi < N && gk < N) {
            g = (acc[i][j] - sm.al[gi] * sm.al[gk]) * coef;
            const long idx = (long)gi * N + gk;
            if (dk) dk[idx] = s * g;
            if (gi == gk) trinv_part += acc[i][j];
          }
          if (kbk < ib) sm.t[cc * GL_LDT + r] = g;
        }
      if (kbk < ib && dk) {
        __syncthreads();
        for (int i = tid; i < GL_T * GL_T; i += GL_THREADS) {
          const int cc = i >> 6, r = i & 63;
          const int gi = i0 + r, gk = k0 + cc;
          if (gi < N && gk < N) {
            const float g = sm.t[cc * GL_LDT + r];
            if (dk) dk[(long)gk * N + gi] = s * g;
          }
        }
      }
    }
  }
  // hyper-parameter gradients in closed form from well-conditioned sums (gp.cu: K~ = s Kb + nz I)
  const float trinv = gl_block_sum(trinv_part, sm.red);
  const float nz = noise + jit;
  const float ds = coef * (((float)N - nz * trinv) - (quad - nz * aa)) / s;
  const float trc = coef * (trinv - aa);
  if (tid == 0 && p.dhyper != nullptr) {
    float* o = p.dhyper + ((long)e * C + c) * 3;
    o[0] = p.raw_outputscale ? ds * dktb_sigmoid(p.raw_outputscale[c]) : 0.f;
    o[1] = -asum * p.grad_scale / ((float)N * (float)C);
    o[2] = trc * dktb_sigmoid(p.raw_noise[c]);
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
  const size_t smem = (size_t)GL_SMEM_FLOATS * sizeof(float);
  // Device build: mma.sync 3xTF32 tile products.  (The g++ emulation build of the tests can select the CUDA-core tile
  // product with DKTB_GP_LARGE=ffma.)  Two CTAs per SM: a third (80 registers) was slower at every size tried
  // (profiles/r01_gp_size_sweep.txt).
#ifdef DKTB_EMU
  static const bool ffma = [] { const char* v = getenv("DKTB_GP_LARGE"); return v && v[0] == 'f'; }();
#else
  const bool ffma = false;
#endif
  if (ffma) {
    cudaFuncSetAttribute(gp_fit_large_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    DKTB_LAUNCH(gp_fit_large_kernel<false>, dim3(C, E), dim3(GL_THREADS), smem, stream, a);
  } else {
    cudaFuncSetAttribute(gp_fit_large_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    DKTB_LAUNCH(gp_fit_large_kernel<true>, dim3(C, E), dim3(GL_THREADS), smem, stream, a);
  }
  return dktb_launch_status();
}
