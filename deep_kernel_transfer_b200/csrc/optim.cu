// Fused flat-buffer Adam (reference: torch.optim.Adam re-created at every DKT.train_loop call with
// GP lr 1e-4 / backbone lr 1e-3, methods/DKT.py:114-115, 164) and small flat-buffer utilities.
// The whole parameter set lives in ONE fp32 buffer so that a single NCCL all-reduce of the matching
// gradient buffer per meta-step is followed by one Adam launch per learning-rate group.
#include "dktb_common.cuh"

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long n, float step_size, float beta1, float beta2, float eps,
                            float bc2_sqrt, float grad_scale) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i] * grad_scale;
  const float mi = m[i] + (gi - m[i]) * (1.f - beta1);
  const float vi = fmaf(beta2, v[i], (1.f - beta2) * gi * gi);
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] -= step_size * (mi / denom);
}

// step is the 1-based Adam step count; grad_scale multiplies the gradient (1/world_size after a sum all-reduce)
DKTB_EXPORT int dktb_adam_step(float* p, const float* g, float* m, float* v, long n, float lr, float beta1,
                               float beta2, float eps, int step, float grad_scale, cudaStream_t stream) {
  DKTB_CHECK_ARG(p && g && m && v && n > 0 && step >= 1);
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  DKTB_LAUNCH(adam_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, stream, p, g, m, v, n, (float)(lr / bc1),
              beta1, beta2, eps, (float)sqrt(bc2), grad_scale);
  return dktb_launch_status();
}

// The same update with the step count read from DEVICE memory, so that a whole meta-step (forward, backward, Adam) can be
// captured once in a CUDA graph and replayed: no scalar that changes from step to step is baked into a launch.
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, long n, float lr, float beta1, float beta2, float eps,
                                const int* __restrict__ step, float grad_scale) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double t = (double)(*step);
  const float step_size = (float)((double)lr / (1.0 - pow((double)beta1, t)));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, t));
  const float gi = g[i] * grad_scale;
  const float mi = m[i] + (gi - m[i]) * (1.f - beta1);
  const float vi = fmaf(beta2, v[i], (1.f - beta2) * gi * gi);
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] -= step_size * (mi / denom);
}

DKTB_EXPORT int dktb_adam_step_dev(float* p, const float* g, float* m, float* v, long n, float lr, float beta1,
                                   float beta2, float eps, const int* step_dev, float grad_scale, cudaStream_t stream) {
  DKTB_CHECK_ARG(p && g && m && v && step_dev && n > 0);
  DKTB_LAUNCH(adam_dev_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, stream, p, g, m, v, n, lr, beta1, beta2,
              eps, step_dev, grad_scale);
  return dktb_launch_status();
}

__global__ void counter_add_kernel(int* __restrict__ c, int delta) { *c += delta; }

DKTB_EXPORT int dktb_counter_add(int* counter, int delta, cudaStream_t stream) {
  DKTB_CHECK_ARG(counter != nullptr);
  DKTB_LAUNCH(counter_add_kernel, dim3(1), dim3(1), 0, stream, counter, delta);
  return dktb_launch_status();
}

__global__ void scale_kernel(float* __restrict__ x, long n, float a) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] *= a;
}

DKTB_EXPORT int dktb_scale(float* x, long n, float a, cudaStream_t stream) {
  DKTB_CHECK_ARG(x && n > 0);
  DKTB_LAUNCH(scale_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, stream, x, n, a);
  return dktb_launch_status();
}

DKTB_EXPORT int dktb_version(void) { return 100; }
