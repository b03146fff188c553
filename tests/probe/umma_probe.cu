// Hardware probe (test infrastructure): verifies on a real B200 the tcgen05 descriptor conventions that the
// convolution kernels in csrc/conv_tc.cu rely on, one variant per process so a faulting variant cannot hide
// the others:
//   ss_k   : D[128x64] = A[128xK] * B[64xK]^T, both operands K-major, SWIZZLE_128B tiles loaded by TMA
//   ts_k   : same product with A supplied from TMEM (written by tcgen05.st, one row per thread)
//   ss_mn  : D[128x64] = At[Kx128]^T * Bt[Kx64], both operands MN-major (the wgrad shape)
//   ts_bf16: kind::f16 with bf16 operands, A packed two k per TMEM column (tcgen05.st), B K-major bf16 in shared memory
//            with the 128-byte swizzle written by the threads, K = 16 per instruction
// Each variant reports the max error against fp64 CPU products of (i) tf32-truncated, (ii) tf32-rounded (rna)
// and (iii) unrounded inputs, which also tells how the tensor core treats the 13 low mantissa bits.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I deep_kernel_transfer_b200/csrc tests/probe/umma_probe.cu -lcuda
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "tc_common.cuh"

#define CK(x)                                                                     \
  do {                                                                            \
    cudaError_t e_ = (x);                                                         \
    if (e_ != cudaSuccess) {                                                      \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                    \
    }                                                                             \
  } while (0)

enum { MODE_SS_K = 0, MODE_TS_K = 1, MODE_SS_MN = 2, MODE_TS_BF16 = 3 };
#include <cuda_bf16.h>
constexpr int M = 128, N = 64;

// smem: A region 64 KB, B region 32 KB (enough for K = 64 K-major or K = 64 rows MN-major)
__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap map_a,
                                                    const __grid_constant__ CUtensorMap map_b, const float* a_raw,
                                                    float* d_out, int* status, int mode, int K) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sa = smem;
  unsigned char* sb = smem + 65536;
  __shared__ uint64_t bar_full, bar_done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32;
  if (tid == 0) {
    tc::mbar_init(&bar_full, 1);
    tc::mbar_init(&bar_done, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc<256>(&tmem_base_s);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t d_tmem = tmem;          // columns [0,64)
  const uint32_t a_tmem = tmem + 64;     // columns [64, 64+K) for the TS variant
  bool ok = true;
  if (tid == 0 && mode == MODE_TS_BF16) tc::mbar_arrive(&bar_full);
  if (tid == 0 && mode != MODE_TS_BF16) {
    uint32_t bytes = 0;
    if (mode == MODE_SS_K || mode == MODE_TS_K) {
      // K-major: atoms of 32 floats along K; A atom = [128 rows][128 B], B atom = [64 rows][128 B]
      for (int h = 0; h < K / 32; ++h) {
        if (mode == MODE_SS_K) { tc::tma_load_2d(sa + h * 16384, &map_a, &bar_full, h * 32, 0); bytes += 16384; }
        tc::tma_load_2d(sb + h * 8192, &map_b, &bar_full, h * 32, 0);
        bytes += 8192;
      }
    } else {
      // MN-major: A atoms = 32 m-columns x K rows (K*128 B each), 4 atoms; B: 2 atoms
      for (int at = 0; at < 4; ++at) { tc::tma_load_2d(sa + at * K * 128, &map_a, &bar_full, at * 32, 0); bytes += K * 128; }
      for (int at = 0; at < 2; ++at) { tc::tma_load_2d(sb + at * K * 128, &map_b, &bar_full, at * 32, 0); bytes += K * 128; }
    }
    tc::mbar_expect_tx(&bar_full, bytes);
  }
  if (mode == MODE_TS_BF16) {
    // A: thread t owns row t; column j of its lane = (bf16(a[2j]) | bf16(a[2j+1]) << 16)
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int c = 0; c < K / 2; c += 16) {
      uint32_t r[16];
      for (int j = 0; j < 16; ++j) {
        const __nv_bfloat16 lo = __float2bfloat16_rn(a_raw[(size_t)tid * K + 2 * (c + j)]);
        const __nv_bfloat16 hi = __float2bfloat16_rn(a_raw[(size_t)tid * K + 2 * (c + j) + 1]);
        r[j] = (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
      }
      tc::tmem_st16(a_tmem + lane_base + c, r);
    }
    tc::tmem_st_wait();
    tc::tcgen05_fence_before();
    // B: [64 rows n][64 k bf16 = 128 B], 16-byte chunk c of row n stored at chunk c ^ (n & 7); b_raw follows a_raw
    const float* b_raw = a_raw + (size_t)M * K;
    for (int i = tid; i < N * (K / 8); i += 128) {
      const int n = i / (K / 8), ch = i % (K / 8);
      uint32_t w[4];
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat16 lo = __float2bfloat16_rn(b_raw[(size_t)n * K + ch * 8 + 2 * j]);
        const __nv_bfloat16 hi = __float2bfloat16_rn(b_raw[(size_t)n * K + ch * 8 + 2 * j + 1]);
        w[j] = (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
      }
      *reinterpret_cast<uint4*>(sb + n * 128 + ((ch ^ (n & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    tc::fence_proxy_async_smem();
  }
  if (mode == MODE_TS_K) {
    // thread t owns row t: write its K floats into TMEM columns a_tmem + [0,K) of lane t
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int c = 0; c < K; c += 16) {
      uint32_t r[16];
      for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(a_raw[(size_t)tid * K + c + j]);
      tc::tmem_st16(a_tmem + lane_base + c, r);
    }
    tc::tmem_st_wait();
    tc::tcgen05_fence_before();
  }
  __syncthreads();
  if (tid == 0) {
    ok = tc::mbar_wait(&bar_full, 0);
    tc::tcgen05_fence_after();
    if (ok) {
      if (mode == MODE_SS_K) {
        const uint32_t idesc = tc::umma_idesc(2, M, N, 0, 0);
        int first = 1;
        for (int h = 0; h < K / 32; ++h)
          for (int k = 0; k < 4; ++k) {
            uint64_t da = tc::umma_desc_sw128(tc::smem_u32(sa + h * 16384) + k * 32, 16, 1024);
            uint64_t db = tc::umma_desc_sw128(tc::smem_u32(sb + h * 8192) + k * 32, 16, 1024);
            tc::umma_tf32_ss(d_tmem, da, db, idesc, first ? 0u : 1u);
            first = 0;
          }
      } else if (mode == MODE_TS_BF16) {
        const uint32_t idesc = tc::umma_idesc(1, M, N, 0, 0);          // bf16 x bf16 -> f32
        for (int k = 0; k < K / 16; ++k) {
          uint64_t db = tc::umma_desc_sw128(tc::smem_u32(sb) + k * 32, 16, 1024);
          tc::umma_f16_ts(d_tmem, a_tmem + k * 8, db, idesc, k ? 1u : 0u);
        }
      } else if (mode == MODE_TS_K) {
        const uint32_t idesc = tc::umma_idesc(2, M, N, 0, 0);
        int first = 1;
        for (int h = 0; h < K / 32; ++h)
          for (int k = 0; k < 4; ++k) {
            uint64_t db = tc::umma_desc_sw128(tc::smem_u32(sb + h * 8192) + k * 32, 16, 1024);
            tc::umma_tf32_ts(d_tmem, a_tmem + h * 32 + k * 8, db, idesc, first ? 0u : 1u);
            first = 0;
          }
      } else {
        const uint32_t idesc = tc::umma_idesc(2, M, N, 1, 1);
        for (int k = 0; k < K / 8; ++k) {
          uint64_t da = tc::umma_desc_sw128(tc::smem_u32(sa) + k * 1024, K * 128, 1024);
          uint64_t db = tc::umma_desc_sw128(tc::smem_u32(sb) + k * 1024, K * 128, 1024);
          tc::umma_tf32_ss(d_tmem, da, db, idesc, k ? 1u : 0u);
        }
      }
      tc::umma_commit(&bar_done);
      ok = tc::mbar_wait(&bar_done, 0);
    }
    if (!ok) *status = 1;
  }
  __syncthreads();
  tc::tcgen05_fence_after();
  // every warp reads its lane quarter: row = tid
  {
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int c = 0; c < N; c += 16) {
      uint32_t r[16];
      tc::tmem_ld16(d_tmem + lane_base + c, r);
      tc::tmem_ld_wait();
      for (int j = 0; j < 16; ++j) d_out[(size_t)tid * N + c + j] = __uint_as_float(r[j]);
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<256>(tmem);
}

static float trunc_tf32(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}
static float rna_tf32(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0x1000u;
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}

int main(int argc, char** argv) {
  const char* which = argc > 1 ? argv[1] : "ss_k";
  int mode = !strcmp(which, "ss_k") ? MODE_SS_K : !strcmp(which, "ts_k") ? MODE_TS_K : !strcmp(which, "ts_bf16") ? MODE_TS_BF16 : MODE_SS_MN;
  const int K = 64;
  srand(1234);
  std::vector<float> A((size_t)M * K), B((size_t)N * K);   // logical A[m][k], B[n][k]
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  // device layouts
  std::vector<float> ha, hb;
  uint64_t a_inner, a_rows, b_inner, b_rows;
  uint32_t a_box_rows, b_box_rows;
  if (mode == MODE_SS_MN) {
    ha.resize((size_t)K * M);
    hb.resize((size_t)K * N);
    for (int k = 0; k < K; ++k) {
      for (int m = 0; m < M; ++m) ha[(size_t)k * M + m] = A[(size_t)m * K + k];
      for (int n = 0; n < N; ++n) hb[(size_t)k * N + n] = B[(size_t)n * K + k];
    }
    a_inner = M; a_rows = K; b_inner = N; b_rows = K; a_box_rows = K; b_box_rows = K;
  } else {
    ha = A; hb = B;
    a_inner = K; a_rows = M; b_inner = K; b_rows = N; a_box_rows = M; b_box_rows = N;
  }
  float *da, *db, *dd;
  int* dstatus;
  if (mode == MODE_TS_BF16) ha.insert(ha.end(), hb.begin(), hb.end());      // a_raw = [A | B] for the in-kernel conversion
  CK(cudaMalloc(&da, ha.size() * 4));
  CK(cudaMalloc(&db, hb.size() * 4));
  CK(cudaMalloc(&dd, (size_t)M * N * 4));
  CK(cudaMalloc(&dstatus, 4));
  CK(cudaMemcpy(da, ha.data(), ha.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dd, 0xff, (size_t)M * N * 4));
  CK(cudaMemset(dstatus, 0, 4));
  CUtensorMap ma, mb;
  int r1 = tc_make_tmap_2d(&ma, da, a_inner, a_rows, 32, a_box_rows);
  int r2 = tc_make_tmap_2d(&mb, db, b_inner, b_rows, 32, b_box_rows);
  if (r1 || r2) { printf("%s: tensor map encode failed %d %d\n", which, r1, r2); return 3; }
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304 + 1024));
  probe_kernel<<<1, 128, 98304 + 1024>>>(ma, mb, da, dd, dstatus, mode, K);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: kernel failed: %s\n", which, cudaGetErrorString(e)); return 4; }
  std::vector<float> D((size_t)M * N);
  int st = 0;
  CK(cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&st, dstatus, 4, cudaMemcpyDeviceToHost));
  double e_trunc = 0, e_rna = 0, e_full = 0, ref_max = 0;
  auto bf16r = [](float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x7FFFu + ((u >> 16) & 1u); u &= 0xFFFF0000u; memcpy(&x, &u, 4); return x; };
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double st_ = 0, sr = 0, sf = 0;
      for (int k = 0; k < K; ++k) {
        float a = A[(size_t)m * K + k], b = B[(size_t)n * K + k];
        if (mode == MODE_TS_BF16) { a = bf16r(a); b = bf16r(b); }
        st_ += (double)trunc_tf32(a) * trunc_tf32(b);
        sr += (double)rna_tf32(a) * rna_tf32(b);
        sf += (double)a * b;
      }
      double d = D[(size_t)m * N + n];
      e_trunc = fmax(e_trunc, fabs(d - st_));
      e_rna = fmax(e_rna, fabs(d - sr));
      e_full = fmax(e_full, fabs(d - sf));
      ref_max = fmax(ref_max, fabs(sf));
    }
  printf("%s: status=%d ref_max=%.3f  max|D-trunc|=%.3e  max|D-rna|=%.3e  max|D-full|=%.3e  D[0][0]=%f D[127][63]=%f\n",
         which, st, ref_max, e_trunc, e_rna, e_full, D[0], D[(size_t)127 * N + 63]);
  bool pass = (e_trunc < 1e-4 || e_rna < 1e-4) && st == 0;
  printf("%s: %s\n", which, pass ? "PASS" : "FAIL");
  return pass ? 0 : 1;
}
