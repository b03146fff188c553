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
#include "dktb_common.cuh"

#define NOISE_FLOOR 1e-4f
#define GL_T 64              // tile edge
#define GL_KC 32             // k-chunk staged per step
#define GL_LDK 36            // pitch of a staged chunk   (36 mod 32 = 4: conflict-free float4 rows)
#define GL_LDT 68            // pitch of a resident tile  (68 mod 32 = 4)
#define GL_LDD 65            // pitch of the diagonal block while it is factored (row per lane)
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

struct GlSmem {
  float* as;      // [64][36]
  float* bs;      // [64][36]
  float* d;       // [64][65]  diagonal block / its Cholesky factor
  float* w;       // [64][68]  inverse of the factor
  float* t;       // [64][68]  a tile handed from one product to the next
  float* diag;    // [512]
  float* r;       // [512] y - m
  float* u;       // [512] L^-1 r
  float* al;      // [512] alpha
  float* red;     // [32]
  int* fail;
};
#define GL_SMEM_FLOATS (2 * GL_T * GL_LDK + GL_T * GL_LDD + 2 * GL_T * GL_LDT + 4 * GL_MAXN + 32 + 4)

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

// acc[i][j] += sum_k A[tr + 16 i][k] * B[tc + 16 j][k], k < kc (kc a multiple of 4), both operands k-contiguous.
__device__ __forceinline__ void gl_core(float (&acc)[4][4], const float* __restrict__ A, int lda,
                                        const float* __restrict__ B, int ldb, int kc, int tr, int tc) {
#pragma unroll 2
  for (int k = 0; k < kc; k += 4) {
    float4 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = dktb_ld4(A + (tr + 16 * i) * lda + k);
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = dktb_ld4(B + (tc + 16 * j) * ldb + k);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = acc[i][j];
        v = fmaf(a[i].x, b[j].x, v);
        v = fmaf(a[i].y, b[j].y, v);
        v = fmaf(a[i].z, b[j].z, v);
        v = fmaf(a[i].w, b[j].w, v);
        acc[i][j] = v;
      }
  }
}

// Operand loaders of one 64 x 32 chunk: element (q, k) of the staged chunk, q = tile row / column, k = reduction index.
//   rows: q-th row of `src` starting at row0, k along the row        (lanes along k: coalesced)
//   cols: q-th column of `src` starting at col0, k down the column   (lanes along q: coalesced)
// Out-of-range q (>= nq) and k (>= klim) read as zero.
struct GlRows {
  const float* src; int ld, row0, nq, klim;
  __device__ __forceinline__ void fetch(float (&v)[8], int k0) const {
    const int k = threadIdx.x & 31, q0 = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int q = q0 + 8 * i;
      v[i] = (q < nq && k0 + k < klim) ? src[(long)(row0 + q) * ld + k0 + k] : 0.f;
    }
  }
  __device__ __forceinline__ void store(float* s, const float (&v)[8]) const {
    const int k = threadIdx.x & 31, q0 = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 8; ++i) s[(q0 + 8 * i) * GL_LDK + k] = v[i];
  }
};
struct GlCols {
  const float* src; int ld, col0, nq, klim;
  __device__ __forceinline__ void fetch(float (&v)[8], int k0) const {
    const int q = threadIdx.x & 63, kk = threadIdx.x >> 6;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = kk + 4 * i;
      v[i] = (q < nq && k0 + k < klim) ? src[(long)(k0 + k) * ld + col0 + q] : 0.f;
    }
  }
  __device__ __forceinline__ void store(float* s, const float (&v)[8]) const {
    const int q = threadIdx.x & 63, kk = threadIdx.x >> 6;
#pragma unroll
    for (int i = 0; i < 8; ++i) s[q * GL_LDK + kk + 4 * i] = v[i];
  }
};

// acc += A-operand x B-operand over k in [kbeg, kend) (kbeg a multiple of 32), staged through sm.as / sm.bs.
template <class LA, class LB>
__device__ __forceinline__ void gl_product(float (&acc)[4][4], const GlSmem& sm, const LA& la, const LB& lb, int kbeg,
                                           int kend, int tr, int tc) {
  float ra[8], rb[8];
  if (kbeg < kend) {
    la.fetch(ra, kbeg);
    lb.fetch(rb, kbeg);
  }
  for (int k0 = kbeg; k0 < kend; k0 += GL_KC) {
    __syncthreads();                       // the previous chunk (or whatever used the staging area) is consumed
    la.store(sm.as, ra);
    lb.store(sm.bs, rb);
    __syncthreads();
    if (k0 + GL_KC < kend) {               // next chunk in flight while this one is multiplied
      la.fetch(ra, k0 + GL_KC);
      lb.fetch(rb, k0 + GL_KC);
    }
    gl_core(acc, sm.as, GL_LDK, sm.bs, GL_LDK, GL_KC, tr, tc);
  }
}

__device__ __forceinline__ void gl_zero(float (&acc)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
}

__global__ void __launch_bounds__(GL_THREADS, 2) gp_fit_large_kernel(GpFitLargeArgs p) {
  DKTB_DYN_SMEM(float, smem);
  GlSmem sm;
  sm.as = smem;
  sm.bs = sm.as + GL_T * GL_LDK;
  sm.d = sm.bs + GL_T * GL_LDK;
  sm.w = sm.d + GL_T * GL_LDD;
  sm.w += (4 - ((sm.w - smem) & 3)) & 3;                 // 16-byte alignment of the float4-read tiles
  sm.t = sm.w + GL_T * GL_LDT;
  sm.diag = sm.t + GL_T * GL_LDT;
  sm.r = sm.diag + GL_MAXN;
  sm.u = sm.r + GL_MAXN;
  sm.al = sm.u + GL_MAXN;
  sm.red = sm.al + GL_MAXN;
  sm.fail = reinterpret_cast<int*>(sm.red + 32);
  const int N = p.N, C = p.C;
  const int c = blockIdx.x, e = blockIdx.y;
  const int tid = threadIdx.x, tr = tid >> 4, tc = tid & 15;
  const float s = p.raw_outputscale ? dktb_softplus(p.raw_outputscale[c]) : 1.f;
  const float noise = dktb_softplus(p.raw_noise[c]) + NOISE_FLOOR;
  const float mconst = p.constant[c];
  const float* kb = p.kbase + ((long)e * (p.kbase_class_stride ? C : 1)) * N * N + (long)c * p.kbase_class_stride;
  const float* yv = p.y + (long)e * p.y_episode_stride + (long)c * N;
  float* A = p.work + ((long)e * C + c) * 2 * N * N;     // K~ -> L (lower incl. diagonal)
  float* X = A + (long)N * N;                             // L^-1 (lower; entries above the diagonal blocks unused)
  if (tid == 0) *sm.fail = 0;
  for (int i = tid; i < N * N; i += GL_THREADS) {
    const int r = i / N, k = i - r * N;
    float v = s * kb[i];
    if (r == k) v += noise + p.jitter;
    A[i] = v;
  }
  for (int i = tid; i < N; i += GL_THREADS) sm.r[i] = yv[i] - mconst;
  __syncthreads();

  const int nT = (N + GL_T - 1) / GL_T;
  float acc[4][4];
  // ---- blocked Cholesky (left-looking by 64-wide block columns) with the block rows of L^-1 produced on the way
  for (int jb = 0; jb < nT; ++jb) {
    const int j0 = jb * GL_T, nb = min(GL_T, N - j0);
    // diagonal tile
    gl_zero(acc);
    {
      GlRows la{A, N, j0, nb, j0}, lb{A, N, j0, nb, j0};
      gl_product(acc, sm, la, lb, 0, j0, tr, tc);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = tr + 16 * i, cc = tc + 16 * j;
        sm.d[r * GL_LDD + cc] = (r < nb && cc < nb) ? A[(long)(j0 + r) * N + j0 + cc] - acc[i][j] : (r == cc ? 1.f : 0.f);
      }
    __syncthreads();
    if (tid < 32) {                                       // lane owns rows lane and lane + 32
      const int lane = tid;
      for (int j = 0; j < nb; ++j) {
        float v0 = 0.f, v1 = 0.f;
        const float* dj = sm.d + j * GL_LDD;
        if (lane >= j && lane < nb) {
          const float* dr = sm.d + lane * GL_LDD;
          v0 = dr[j];
          for (int k = 0; k < j; ++k) v0 = fmaf(-dr[k], dj[k], v0);
        }
        if (lane + 32 >= j && lane + 32 < nb) {
          const float* dr = sm.d + (lane + 32) * GL_LDD;
          v1 = dr[j];
          for (int k = 0; k < j; ++k) v1 = fmaf(-dr[k], dj[k], v1);
        }
        const float piv = __shfl_sync(0xffffffffu, j < 32 ? v0 : v1, j & 31);
        if (!(piv > 0.f)) {
          if (lane == 0) *sm.fail = j0 + j + 1;
          break;
        }
        const float dg = sqrtf(piv);
        if (lane == j) { sm.d[j * GL_LDD + j] = dg; sm.diag[j0 + j] = dg; }
        else if (lane > j && lane < nb) sm.d[lane * GL_LDD + j] = v0 / dg;
        if (lane + 32 == j) { sm.d[j * GL_LDD + j] = dg; sm.diag[j0 + j] = dg; }
        else if (lane + 32 > j && lane + 32 < nb) sm.d[(lane + 32) * GL_LDD + j] = v1 / dg;
        __syncwarp();
      }
    }
    __syncthreads();
    if (*sm.fail) break;
    // W = L_jj^-1: a thread per column (forward substitution); rows beyond nb form an identity block
    if (tid < GL_T) {
      const int col = tid;
      for (int i = 0; i < GL_T; ++i) {
        float a = 0.f;
        if (i >= col) {
          a = (i == col) ? 1.f : 0.f;
          const float* di = sm.d + i * GL_LDD;
          for (int k = col; k < i; ++k) a = fmaf(-di[k], sm.w[k * GL_LDT + col], a);
          a = a / di[i];
        }
        sm.w[i * GL_LDT + col] = a;
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
        gl_product(acc, sm, la, lb, 0, j0, tr, tc);
      }
      __syncthreads();                                    // sm.t free (previous tile's second product done)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = tr + 16 * i, cc = tc + 16 * j;
          sm.t[r * GL_LDT + cc] = (r < nr && cc < nb) ? A[(long)(i0 + r) * N + j0 + cc] - acc[i][j] : 0.f;
        }
      __syncthreads();
      gl_zero(acc);
      gl_core(acc, sm.t, GL_LDT, sm.w, GL_LDT, GL_T, tr, tc);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = tr + 16 * i, cc = tc + 16 * j;
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
        gl_product(acc, sm, la, lb, k0, j0, tr, tc);
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) sm.t[(tc + 16 * j) * GL_LDT + tr + 16 * i] = acc[i][j];     // transposed: t[c][m]
      __syncthreads();
      gl_zero(acc);
      gl_core(acc, sm.w, GL_LDT, sm.t, GL_LDT, GL_T, tr, tc);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = tr + 16 * i, cc = tc + 16 * j;
          if (r < nb) X[(long)(j0 + r) * N + k0 + cc] = -acc[i][j];
        }
    }
    __syncthreads();
  }
  __syncthreads();
  const int fail = *sm.fail;
  if (tid == 0) p.info[(long)e * C + c] = fail;
  if (fail) {
    if (tid == 0) p.loss_terms[(long)e * C + c] = nanf("");
    return;
  }
  // ---- u = L^-1 r (a warp per row), alpha = L^-T u (a thread per column, coalesced down the rows)
  {
    const int lane = tid & 31, warp = tid >> 5;
    for (int i = warp; i < N; i += GL_THREADS / 32) {
      float a = 0.f;
      for (int k = lane; k <= i; k += 32) a = fmaf(X[(long)i * N + k], sm.r[k], a);
      a = dktb_warp_sum(a);
      if (lane == 0) sm.u[i] = a;
    }
  }
  __syncthreads();
  float quad_part = 0.f, logdet_part = 0.f, asum_part = 0.f;
  for (int k = tid; k < N; k += GL_THREADS) {
    float a = 0.f;
    for (int i = k; i < N; ++i) a = fmaf(X[(long)i * N + k], sm.u[i], a);
    sm.al[k] = a;
    p.alpha[((long)e * C + c) * N + k] = a;
    quad_part = fmaf(sm.r[k], a, quad_part);
    logdet_part += logf(sm.diag[k]);
    asum_part += a;
  }
  const float quad = gl_block_sum(quad_part, sm.red);
  const float logdet = 2.f * gl_block_sum(logdet_part, sm.red);
  const float asum = gl_block_sum(asum_part, sm.red);
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
  float ds_part = 0.f, tr_part = 0.f;
  for (int ib = 0; ib < nT; ++ib) {
    const int i0 = ib * GL_T, ni = min(GL_T, N - i0);
    for (int kbk = 0; kbk <= ib; ++kbk) {
      const int k0 = kbk * GL_T;
      gl_zero(acc);
      {
        GlCols la{X, N, i0, ni, N}, lb{X, N, k0, GL_T, N};       // kbk <= ib: the k-tile is full unless it is the last
        lb.nq = min(GL_T, N - k0);
        gl_product(acc, sm, la, lb, i0, N, tr, tc);
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = tr + 16 * i, cc = tc + 16 * j;
          const int gi = i0 + r, gk = k0 + cc;
          float g = 0.f;
          if (gi < N && gk < N) {
            g = (acc[i][j] - sm.al[gi] * sm.al[gk]) * coef;
            const long idx = (long)gi * N + gk;
            if (dk) dk[idx] = s * g;
            ds_part = fmaf(g, kb[idx], ds_part);
            if (gi == gk) tr_part += g;
          }
          if (kbk < ib) sm.t[cc * GL_LDT + r] = g;
        }
      if (kbk < ib) {
        __syncthreads();
        for (int i = tid; i < GL_T * GL_T; i += GL_THREADS) {
          const int cc = i >> 6, r = i & 63;
          const int gi = i0 + r, gk = k0 + cc;
          if (gi < N && gk < N) {
            const float g = sm.t[cc * GL_LDT + r];
            const long idx = (long)gk * N + gi;
            if (dk) dk[idx] = s * g;
            ds_part = fmaf(g, kb[idx], ds_part);
          }
        }
      }
    }
  }
  const float ds = gl_block_sum(ds_part, sm.red);
  const float trc = gl_block_sum(tr_part, sm.red);
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
  cudaFuncSetAttribute(gp_fit_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  DKTB_LAUNCH(gp_fit_large_kernel, dim3(C, E), dim3(GL_THREADS), smem, stream, a);
  return dktb_launch_status();
}
