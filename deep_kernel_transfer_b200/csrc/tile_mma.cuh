// Shared 64 x 64 x k tile-product machinery (used by gp_large.cu and conv_mma.cu): a CTA of 256 threads multiplies two
// k-contiguous operands staged in shared memory, either on the warp-level tensor cores (mma.sync m16n8k8, 3xTF32) or on the
// CUDA cores, with the next 32-wide k-chunk prefetched into registers while the current one is multiplied.
#pragma once
#include "dktb_common.cuh"

#define GL_T 64              // tile edge
#define GL_KC 32             // k-chunk staged per step
#define GL_LDK 36            // pitch of a staged chunk   (36 mod 32 = 4: conflict-free float4 rows)
#define GL_LDT 68            // pitch of a resident tile  (68 mod 32 = 4)
#define GL_THREADS 256

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

// The same tile product on the warp-level tensor cores (mma.sync m16n8k8, 3xTF32: operands split into a tf32 `hi` and
// the exact remainder `lo`, products lo*hi + hi*lo + hi*hi accumulated in fp32): warp w owns rows 32 (w / 4) .. + 31 and
// columns 16 (w % 4) .. + 15 of the tile as 2 x 2 fragments; acc[i][j] = fragment (i / 2, j / 2), register 2 (i % 2) + j % 2.
__device__ __forceinline__ void gl_split(float v, unsigned& hi, unsigned& lo) {
  hi = __float_as_uint(v) & 0xFFFFE000u;
  lo = __float_as_uint(v - __uint_as_float(hi));
}
__device__ __forceinline__ void gl_core_mma(float (&acc)[4][4], const float* __restrict__ A, int lda,
                                            const float* __restrict__ B, int ldb, int kc) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const float* a0 = A + ((warp >> 2) * 32 + g) * lda + t;
  const float* b0 = B + ((warp & 3) * 16 + g) * ldb + t;
  // The tensor core's accumulator does not round to nearest: over a long reduction the loss grows linearly with the
  // number of accumulations (measured 4e-5 relative at k = 4608).  So each call sums its <= 64-wide chunk into a fresh
  // accumulator and the running total is kept by ordinary fp32 additions.
  float part[2][2][4];
#pragma unroll
  for (int mf = 0; mf < 2; ++mf)
#pragma unroll
    for (int nf = 0; nf < 2; ++nf)
#pragma unroll
      for (int q = 0; q < 4; ++q) part[mf][nf][q] = 0.f;
#pragma unroll 2
  for (int k = 0; k < kc; k += 8) {
    unsigned ah[2][4], al[2][4], bh[2][2], bl[2][2];
#pragma unroll
    for (int mf = 0; mf < 2; ++mf) {
      const float* ap = a0 + mf * 16 * lda + k;
      gl_split(ap[0], ah[mf][0], al[mf][0]);
      gl_split(ap[8 * lda], ah[mf][1], al[mf][1]);
      gl_split(ap[4], ah[mf][2], al[mf][2]);
      gl_split(ap[8 * lda + 4], ah[mf][3], al[mf][3]);
    }
#pragma unroll
    for (int nf = 0; nf < 2; ++nf) {
      const float* bp = b0 + nf * 8 * ldb + k;
      gl_split(bp[0], bh[nf][0], bl[nf][0]);
      gl_split(bp[4], bh[nf][1], bl[nf][1]);
    }
#pragma unroll
    for (int mf = 0; mf < 2; ++mf)
#pragma unroll
      for (int nf = 0; nf < 2; ++nf) {
        dktb_mma_m16n8k8_tf32(part[mf][nf], al[mf], bh[nf]);
        dktb_mma_m16n8k8_tf32(part[mf][nf], ah[mf], bl[nf]);
        dktb_mma_m16n8k8_tf32(part[mf][nf], ah[mf], bh[nf]);
      }
  }
#pragma unroll
  for (int mf = 0; mf < 2; ++mf)
#pragma unroll
    for (int nf = 0; nf < 2; ++nf) {
      acc[2 * mf][2 * nf] += part[mf][nf][0];
      acc[2 * mf][2 * nf + 1] += part[mf][nf][1];
      acc[2 * mf + 1][2 * nf] += part[mf][nf][2];
      acc[2 * mf + 1][2 * nf + 1] += part[mf][nf][3];
    }
}

// Which tile row / column acc[i][j] of this thread holds, for the two cores.
template <bool MMA> struct GlMap;
template <> struct GlMap<false> {
  __device__ static __forceinline__ int row(int i) { return (threadIdx.x >> 4) + 16 * i; }
  __device__ static __forceinline__ int col(int j) { return (threadIdx.x & 15) + 16 * j; }
};
template <> struct GlMap<true> {
  __device__ static __forceinline__ int row(int i) {
    return ((threadIdx.x >> 7) & 1) * 32 + (i >> 1) * 16 + ((threadIdx.x & 31) >> 2) + 8 * (i & 1);
  }
  __device__ static __forceinline__ int col(int j) {
    return ((threadIdx.x >> 5) & 3) * 16 + (j >> 1) * 8 + 2 * (threadIdx.x & 3) + (j & 1);
  }
};
template <bool MMA>
__device__ __forceinline__ void gl_tile(float (&acc)[4][4], const float* A, int lda, const float* B, int ldb, int kc) {
  if (MMA) gl_core_mma(acc, A, lda, B, ldb, kc);
  else gl_core(acc, A, lda, B, ldb, kc, threadIdx.x >> 4, threadIdx.x & 15);
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

// acc += A-operand x B-operand over k in [kbeg, kend) (kbeg a multiple of 32), staged through as / bs, each TWO chunks of
// [64][36] floats: chunk c + 1 is written into the other half while chunk c is multiplied, so the loop needs one barrier
// per chunk; chunk c + 2 is in flight in registers meanwhile.
template <bool MMA, class LA, class LB>
__device__ __forceinline__ void gl_product(float (&acc)[4][4], float* as, float* bs, const LA& la, const LB& lb,
                                           int kbeg, int kend) {
  if (kbeg >= kend) return;
  float ra[8], rb[8];
  la.fetch(ra, kbeg);
  lb.fetch(rb, kbeg);
  __syncthreads();                         // whatever used the staging area before is done with it
  la.store(as, ra);
  lb.store(bs, rb);
  if (kbeg + GL_KC < kend) {
    la.fetch(ra, kbeg + GL_KC);
    lb.fetch(rb, kbeg + GL_KC);
  }
  __syncthreads();
  int buf = 0;
  for (int k0 = kbeg; k0 < kend; k0 += GL_KC, buf ^= 1) {
    gl_tile<MMA>(acc, as + buf * GL_T * GL_LDK, GL_LDK, bs + buf * GL_T * GL_LDK, GL_LDK, GL_KC);
    if (k0 + GL_KC < kend) {
      la.store(as + (buf ^ 1) * GL_T * GL_LDK, ra);
      lb.store(bs + (buf ^ 1) * GL_T * GL_LDK, rb);
      if (k0 + 2 * GL_KC < kend) {
        la.fetch(ra, k0 + 2 * GL_KC);
        lb.fetch(rb, k0 + 2 * GL_KC);
      }
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void gl_zero(float (&acc)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
}

// ---- pre-split variant: the staged chunk is kept as two planes (tf32 `hi`, exact remainder `lo`) written once by the
// thread that fetched the element, and the fragments are read with ldmatrix (one x4 per 16 x 8 A fragment / per pair of
// 8 x 8 B fragments: a tf32 fragment is four 8 x 8 matrices of 16-bit pairs).  Per 8-wide k-step a warp issues 6 ldmatrix
// and 12 mma instead of 12 shared loads, 24 split operations and 12 mma.
#define GL_PLANE (GL_T * GL_LDK)     // floats per staged plane

// r[m] = 32-bit element (lane / 4, lane % 4) of the 8 x 4 matrix whose row `lane % 8` address is given by lane 8 m + lane % 8
__device__ __forceinline__ void gl_ldmatrix_x4(unsigned (&r)[4], const float* row) {
#ifdef DKTB_EMU
  const int lane = threadIdx.x & 31;
  const unsigned long long mine = (unsigned long long)(uintptr_t)row;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const unsigned long long p = __shfl_sync(0xffffffffu, mine, m * 8 + (lane >> 2));
    r[m] = __float_as_uint(reinterpret_cast<const float*>((uintptr_t)p)[lane & 3]);
  }
#else
  const unsigned addr = (unsigned)__cvta_generic_to_shared(row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
#endif
}

// acc += A x B^T over one 32-wide chunk held as planes a_hi / a_lo / b_hi / b_lo ([64][GL_LDK] each)
__device__ __forceinline__ void gl_core_mma_ps(float (&acc)[4][4], const float* a_hi, const float* a_lo,
                                               const float* b_hi, const float* b_lo) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // A: matrices (rows 0-7, k 0-3), (rows 8-15, k 0-3), (rows 0-7, k 4-7), (rows 8-15, k 4-7) -> a0..a3
  const int a_off = ((warp >> 2) * 32 + ((lane >> 3) & 1) * 8 + (lane & 7)) * GL_LDK + (lane >> 4) * 4;
  // B: matrices (cols 0-7, k 0-3), (cols 0-7, k 4-7), (cols 8-15, k 0-3), (cols 8-15, k 4-7) -> nf0.b0, nf0.b1, nf1.b0, nf1.b1
  const int b_off = ((warp & 3) * 16 + (lane >> 4) * 8 + (lane & 7)) * GL_LDK + ((lane >> 3) & 1) * 4;
  float part[2][2][4];
#pragma unroll
  for (int mf = 0; mf < 2; ++mf)
#pragma unroll
    for (int nf = 0; nf < 2; ++nf)
#pragma unroll
      for (int q = 0; q < 4; ++q) part[mf][nf][q] = 0.f;
#pragma unroll
  for (int k = 0; k < GL_KC; k += 8) {
    unsigned ah[2][4], al[2][4], bh4[4], bl4[4];
#pragma unroll
    for (int mf = 0; mf < 2; ++mf) {
      gl_ldmatrix_x4(ah[mf], a_hi + a_off + mf * 16 * GL_LDK + k);
      gl_ldmatrix_x4(al[mf], a_lo + a_off + mf * 16 * GL_LDK + k);
    }
    gl_ldmatrix_x4(bh4, b_hi + b_off + k);
    gl_ldmatrix_x4(bl4, b_lo + b_off + k);
#pragma unroll
    for (int nf = 0; nf < 2; ++nf) {
      const unsigned bh[2] = {bh4[2 * nf], bh4[2 * nf + 1]}, bl[2] = {bl4[2 * nf], bl4[2 * nf + 1]};
#pragma unroll
      for (int mf = 0; mf < 2; ++mf) {
        dktb_mma_m16n8k8_tf32(part[mf][nf], al[mf], bh);
        dktb_mma_m16n8k8_tf32(part[mf][nf], ah[mf], bl);
        dktb_mma_m16n8k8_tf32(part[mf][nf], ah[mf], bh);
      }
    }
  }
#pragma unroll
  for (int mf = 0; mf < 2; ++mf)
#pragma unroll
    for (int nf = 0; nf < 2; ++nf) {
      acc[2 * mf][2 * nf] += part[mf][nf][0];
      acc[2 * mf][2 * nf + 1] += part[mf][nf][1];
      acc[2 * mf + 1][2 * nf] += part[mf][nf][2];
      acc[2 * mf + 1][2 * nf + 1] += part[mf][nf][3];
    }
}

// gl_product on split planes: as / bs hold [hi plane | lo plane] (2 * GL_PLANE floats each).  Accumulator ownership is
// GlMap<true>.
template <class LA, class LB>
__device__ __forceinline__ void gl_product_ps(float (&acc)[4][4], float* as, float* bs, const LA& la, const LB& lb,
                                              int kbeg, int kend) {
  float ra[8], rb[8];
  if (kbeg < kend) {
    la.fetch(ra, kbeg);
    lb.fetch(rb, kbeg);
  }
  for (int k0 = kbeg; k0 < kend; k0 += GL_KC) {
    float hi[8], lo[8];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      hi[i] = __uint_as_float(__float_as_uint(ra[i]) & 0xFFFFE000u);
      lo[i] = ra[i] - hi[i];
    }
    la.store(as, hi);
    la.store(as + GL_PLANE, lo);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      hi[i] = __uint_as_float(__float_as_uint(rb[i]) & 0xFFFFE000u);
      lo[i] = rb[i] - hi[i];
    }
    lb.store(bs, hi);
    lb.store(bs + GL_PLANE, lo);
    __syncthreads();
    if (k0 + GL_KC < kend) {
      la.fetch(ra, k0 + GL_KC);
      lb.fetch(rb, k0 + GL_KC);
    }
    gl_core_mma_ps(acc, as, as + GL_PLANE, bs, bs + GL_PLANE);
  }
}
