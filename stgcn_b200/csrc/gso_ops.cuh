// gso_ops.cuh -- graph-shift-operator preprocessing on the device: the dense-matrix equivalent of
// calc_gso / calc_chebynet_gso (script/utility.py:6-76).  One-time work per dataset; exists so that large operators
// (the N = 2048 sweep and beyond) never take the scipy sparse -> dense -> host -> device detour, and so the rescaled
// Laplacian's largest eigenvalue comes from a device-side power iteration instead of scipy.sparse.linalg.norm(gso, 2).
#pragma once
#include "common.cuh"

namespace stgcn {
namespace gso {

// utility.py:18: adj + adj.T * (adj.T > adj) - adj * (adj.T > adj) == elementwise max(adj, adj.T); :21-23: + I for *_renorm_*
__global__ void symmetrize_kernel(const float* adj, float* a, int N, int renorm) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * N) return;
  const int i = (int)(idx / N), j = (int)(idx - (long long)i * N);
  float v = fmaxf(adj[idx], adj[(long long)j * N + i]);
  if (renorm && i == j) v += 1.f;
  a[idx] = v;
}
// one warp per row: d[i] = sum_j a[i][j]  (utility.py:27,42)
__global__ void rowsum_kernel(const float* a, float* d, int N) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= N) return;
  float s = 0.f;
  for (int j = lane; j < N; j += 32) s += a[(long long)row * N + j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) d[row] = s;
}
// sym: D^-1/2 A D^-1/2 (utility.py:28-32); rw: D^-1 A (:43-47); lap: I - that (:34-36, :49-51); 1/0 -> 0 (:29, :44)
__global__ void normalize_kernel(const float* a, const float* d, float* out, int N, int rw, int lap) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * N) return;
  const int i = (int)(idx / N), j = (int)(idx - (long long)i * N);
  float v;
  if (rw) {
    const float di = d[i] != 0.f ? 1.f / d[i] : 0.f;
    v = di * a[idx];
  } else {
    const float di = d[i] > 0.f ? rsqrtf(d[i]) : 0.f, dj = d[j] > 0.f ? rsqrtf(d[j]) : 0.f;
    v = di * a[idx] * dj;
  }
  out[idx] = lap ? ((i == j ? 1.f : 0.f) - v) : v;
}

// Spectral norm ||G||_2 = sqrt(lambda_max(G^T G)) by power iteration on G^T G, one CTA of 1024 threads (the operator
// stays in L2; per iteration two mat-vecs with fp32 FMAs, norms in fp64).  Stops when the estimate moved by less than
// `tol` relative over 64 iterations, or after max_iter.  result[0] = sigma_max, result[1] = iterations used.
__global__ void __launch_bounds__(1024) spectral_norm_kernel(const float* G, int N, float* v, float* u, float* result,
                                                             int max_iter, float tol) {
  __shared__ double red[32];
  __shared__ double s_norm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  auto block_sum = [&](double x) -> double {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) red[warp] = x;
    __syncthreads();
    if (warp == 0) {
      double y = lane < nwarps ? red[lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) y += __shfl_xor_sync(0xffffffffu, y, o);
      if (lane == 0) s_norm = y;
    }
    __syncthreads();
    return s_norm;
  };
  // deterministic start vector with a component along every eigenvector (a constant vector is an exact null vector of
  // a Laplacian): hashed signs and magnitudes
  double ss = 0.0;
  for (int i = tid; i < N; i += blockDim.x) {
    unsigned h = (unsigned)i * 2654435761u + 12345u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    const float x = ((h & 0xffff) / 65536.f + 0.25f) * ((h >> 16) & 1 ? 1.f : -1.f);
    v[i] = x; ss += (double)x * x;
  }
  double nrm = sqrt(block_sum(ss));
  for (int i = tid; i < N; i += blockDim.x) v[i] = (float)(v[i] / nrm);
  __syncthreads();
  double sigma2 = 0.0, sigma2_prev = -1.0;
  int it = 0;
  for (; it < max_iter; ++it) {
    // u = G v: one warp per row, lanes stride the columns
    for (int r = warp; r < N; r += nwarps) {
      float s = 0.f;
      const float* row = G + (long long)r * N;
      for (int j = lane; j < N; j += 32) s = fmaf(row[j], v[j], s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) u[r] = s;
    }
    __syncthreads();
    // w = G^T u: one thread per column, coalesced across the threads of a warp; w overwrites v after the norm
    double ws = 0.0;
    float wloc[2] = {0.f, 0.f};                       // N <= 2048 = 2 columns per thread
    int k = 0;
    for (int j = tid; j < N; j += blockDim.x, ++k) {
      float s = 0.f;
      for (int r = 0; r < N; ++r) s = fmaf(G[(long long)r * N + j], u[r], s);
      wloc[k] = s; ws += (double)s * s;
    }
    nrm = sqrt(block_sum(ws));                         // ||G^T G v|| -> lambda_max(G^T G) as v converges
    sigma2 = nrm;
    k = 0;
    for (int j = tid; j < N; j += blockDim.x, ++k) v[j] = (float)(wloc[k] / nrm);
    __syncthreads();
    if ((it & 63) == 63) {
      if (sigma2_prev > 0.0 && fabs(sigma2 - sigma2_prev) <= (double)tol * sigma2) { ++it; break; }
      sigma2_prev = sigma2;
    }
  }
  if (tid == 0) { result[0] = (float)sqrt(sigma2); result[1] = (float)it; }
}
// utility.py:69-72: eigval_max >= 2 -> gso - I, else 2 gso / eigval_max - I
__global__ void cheb_rescale_kernel(const float* g, float* out, int N, const float* eig) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * N) return;
  const int i = (int)(idx / N), j = (int)(idx - (long long)i * N);
  const float lam = eig[0], id = i == j ? 1.f : 0.f;
  out[idx] = lam >= 2.f ? g[idx] - id : 2.f * g[idx] / lam - id;
}

}  // namespace gso
}  // namespace stgcn
