// train_ops.cuh -- the callers either side of the ST-block path: the optimizer step fused over ONE flat
// parameter / gradient buffer (N2: main.py:147-156,169; script/opt.py:34-76) and window construction on the device from
// the resident series (N3: script/dataloader.py:32-48).  HBM-bound elementwise / copy kernels: 16-byte accesses, grids
// sized in multiples of the SM count.
#pragma once
#include "common.cuh"

namespace stgcn {
namespace train {

struct AdamWArgs {
  float* p; const float* g; float* m; float* v;
  long long n;
  float lr, beta1, beta2, eps, wd, grad_scale;
  long long step;                       // 1-based step number used for the bias corrections ...
  const long long* step_dev;            // ... or, when non-null, *step_dev + 1 (a CUDA-graph replay cannot change `step`)
  const float* lr_dev;                  // optional device-side learning rate (StepLR changes it between epochs)
};

// torch.optim.AdamW (decoupled weight decay, amsgrad off, maximize off), the reference's default optimizer
// (main.py:147-148): p *= 1 - lr*wd; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
// p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, float lr, float b1, float b2, float eps,
                                          float wd, float step_size, float inv_bc2_sqrt) {
  p *= 1.f - lr * wd;
  m = fmaf(1.f - b1, g - m, m);            // torch: exp_avg.lerp_(grad, 1 - beta1) = m + (1 - beta1) (g - m)
  v = b2 * v + (1.f - b2) * g * g;        // torch: exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
  const float denom = sqrtf(v) * inv_bc2_sqrt + eps;
  p -= step_size * (m / denom);
}
__global__ void __launch_bounds__(256) adamw_kernel(AdamWArgs a) {
  const long long t = a.step_dev ? *a.step_dev + 1 : a.step;
  const float lr = a.lr_dev ? *a.lr_dev : a.lr;
  const float bc1 = 1.f - powf(a.beta1, (float)t), bc2 = 1.f - powf(a.beta2, (float)t);
  const float step_size = lr / bc1, inv_bc2_sqrt = 1.f / sqrtf(bc2);
  const long long n4 = a.n >> 2, stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 p = reinterpret_cast<float4*>(a.p)[i], m = reinterpret_cast<float4*>(a.m)[i], v = reinterpret_cast<float4*>(a.v)[i];
    float4 g = reinterpret_cast<const float4*>(a.g)[i];
    g.x *= a.grad_scale; g.y *= a.grad_scale; g.z *= a.grad_scale; g.w *= a.grad_scale;
    adamw_one(p.x, g.x, m.x, v.x, lr, a.beta1, a.beta2, a.eps, a.wd, step_size, inv_bc2_sqrt);
    adamw_one(p.y, g.y, m.y, v.y, lr, a.beta1, a.beta2, a.eps, a.wd, step_size, inv_bc2_sqrt);
    adamw_one(p.z, g.z, m.z, v.z, lr, a.beta1, a.beta2, a.eps, a.wd, step_size, inv_bc2_sqrt);
    adamw_one(p.w, g.w, m.w, v.w, lr, a.beta1, a.beta2, a.eps, a.wd, step_size, inv_bc2_sqrt);
    reinterpret_cast<float4*>(a.p)[i] = p; reinterpret_cast<float4*>(a.m)[i] = m; reinterpret_cast<float4*>(a.v)[i] = v;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
    float p = a.p[i], m = a.m[i], v = a.v[i];
    adamw_one(p, a.g[i] * a.grad_scale, m, v, lr, a.beta1, a.beta2, a.eps, a.wd, step_size, inv_bc2_sqrt);
    a.p[i] = p; a.m[i] = m; a.v[i] = v;
  }
}

// Lion (script/opt.py:34-76): p *= 1 - lr*wd; p -= lr * sign(b1 m + (1-b1) g); m = b2 m + (1-b2) g
__global__ void __launch_bounds__(256) lion_kernel(float* p, const float* g, float* m, long long n, float lr,
                                                   const float* lr_dev, float b1, float b2, float wd, float grad_scale) {
  if (lr_dev) lr = *lr_dev;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i] * grad_scale, mi = m[i];
    float pi = p[i] * (1.f - lr * wd);
    const float u = mi * b1 + gi * (1.f - b1);
    pi -= lr * (u > 0.f ? 1.f : (u < 0.f ? -1.f : 0.f));       // torch.sign: sign(0) = 0
    p[i] = pi;
    m[i] = mi * b2 + gi * (1.f - b2);
  }
}

// x[i, 0, t, :] = series[start_i + t, :] (t < n_his); y[i, :] = series[start_i + n_his + n_pred - 1, :]
// (data_transform, script/dataloader.py:32-48, for the windows of one batch).  start_i = starts[i], or start0 + i.
__global__ void __launch_bounds__(256) windows_kernel(const float* series, long long len, int N, int n_his, int n_pred,
                                                      const long long* starts, long long start0, int B, float* x, float* y) {
  const long long per = (long long)(n_his + 1) * N;        // n_his input rows + the target row per window
  const long long total = (long long)B * per, stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int i = (int)(e / per);
    const long long r = e - (long long)i * per;
    const int t = (int)(r / N), n = (int)(r - (long long)t * N);
    const long long s = starts ? starts[i] : start0 + i;
    const long long row = t < n_his ? s + t : s + n_his + n_pred - 1;
    const float v = (row >= 0 && row < len) ? series[row * N + n] : 0.f;
    if (t < n_his) x[((long long)i * n_his + t) * N + n] = v;
    else y[(long long)i * N + n] = v;
  }
}

inline int elementwise_grid(long long work_items) {
  long long blocks = (work_items + 255) / 256;
  const long long cap = 148LL * 8;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace train
}  // namespace stgcn
