// simt_kernels.cuh -- fp32 CUDA-core kernels of the parity path (precision == STGCN_PREC_FP32).
//
// Data layout everywhere: activations are row-major [rows, C] with rows = (b, t, n) and the
// channel axis innermost ("channels-last"), i.e. the memory order of the reference's own
// permuted (B,T,N,C) tensors (layers.py:145,255).  In that layout every stage of the ST block is
// a GEMM over rows:
//   * temporal (Kt,1) convolution  = "tap GEMM": Kt row-shifted copies of the input tile times
//     per-tap weight slices (layers.py:52-57,89);  its data-gradient is the same GEMM with
//     negative shifts and a validity mask; 1x1 align convs / nn.Linear are the 1-tap case
//     (layers.py:16, 270-271); the Chebyshev weight contraction is a Ks-tap GEMM over the stacked
//     x_k tensors (layers.py:163-165).
//   * node contraction  out[g,h,:] = alpha * sum_i L[h,i] x[g,i,:] + beta * aux[g,h,:]
//     (layers.py:154-161), one [N x N] x [N x (groups*C)] GEMM.
//   * weight gradients = GEMMs whose contraction axis is the row axis, split over CTAs and
//     reduced with fp32 atomics.
#pragma once
#include <type_traits>
#include <algorithm>

#include "common.cuh"

namespace stgcn {
namespace simt {
struct RowMap;
template <class TI, class TO = TI> struct TapArgs;
template <class T> struct GsoArgs;
template <class T> struct WgradArgs;
}  // namespace simt
namespace umma {      // umma_x3.cuh: the same three GEMMs on tcgen05 (3xTF32); false = shape not served
inline bool x3_tapgemm(const simt::TapArgs<float, float>& a, cudaStream_t stream);
inline bool x3_gso(const simt::GsoArgs<float>& a, cudaStream_t stream);
inline bool x3_wgrad(const simt::WgradArgs<float>& a, cudaStream_t stream);
}  // namespace umma
namespace simt {

constexpr int BK = 16;
constexpr int NT = 256;

// activation storage type T: float (parity path) or __nv_bfloat16 (throughput path); math is always fp32
using bf16 = __nv_bfloat16;
__device__ __forceinline__ float ldf(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ldf(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(bf16* p, float v) { *p = __float2bfloat16_rn(v); }
// 8-element vector access (16-byte aligned for bf16, 32-byte for float)
__device__ __forceinline__ void store8(float* p, const float* v) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float* v) {
  uint4 a;
  __nv_bfloat162 t0 = __floats2bfloat162_rn(v[0], v[1]), t1 = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 t2 = __floats2bfloat162_rn(v[4], v[5]), t3 = __floats2bfloat162_rn(v[6], v[7]);
  a.x = *reinterpret_cast<uint32_t*>(&t0); a.y = *reinterpret_cast<uint32_t*>(&t1);
  a.z = *reinterpret_cast<uint32_t*>(&t2); a.w = *reinterpret_cast<uint32_t*>(&t3);
  *reinterpret_cast<uint4*>(p) = a;
}
__device__ __forceinline__ void load8(const float* p, float* v) {
  float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float* v) {
  uint4 a = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
    v[2 * i] = __low2float(h);
    v[2 * i + 1] = __high2float(h);
  }
}

// 8 packed bf16 (16 bytes) -> 8 floats (bf16 -> fp32 is a 16-bit shift)
__device__ __forceinline__ void unpack8(const uint4& a, float* v) {
  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}

// (b, rem) = divmod(r, TN_out), t = rem / N with 32-bit arithmetic (64-bit integer division costs ~100 instructions on
// the GPU and dominated the elementwise kernels); the host guarantees rows < 2^31.
__device__ __forceinline__ void row_decode(long long r, int TN_out, int N, long long TN_in, long long& base, int& t) {
  const unsigned ru = (unsigned)r;
  const unsigned b = ru / (unsigned)TN_out;
  const unsigned rem = ru - b * (unsigned)TN_out;
  t = (int)(rem / (unsigned)N);
  base = (long long)b * TN_in + rem;
}

// Row geometry of a tap GEMM.  Output row r = (b, t, n), t < T_out.  Tap k reads input row
// (b, t + t_shift*k, n) of a [B, T_in, N] tensor (invalid -> contributes 0), displaced by
// k * tap_row_stride rows (stacked operands).
struct RowMap {
  int T_out, T_in, N;
  int t_shift;
  long long tap_row_stride;
};

template <class TI, class TO>
struct TapArgs {
  const TI* in;        // [*, Cin]
  const float* wt;     // [ntaps*Cin, Co], Co contiguous
  const float* bias;   // [Co] or nullptr
  TO* out;             // [rows, ldo]
  long long rows;
  int Cin, Co, ntaps, ldo;
  int accumulate;      // out += result
  RowMap map;
};

template <class TI, class TO, int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(NT) tapgemm_kernel(TapArgs<TI, TO> a) {
  constexpr int TX = BN / TN, TY = BM / TM;
  static_assert(TX * TY == NT, "tile/thread mismatch");
  constexpr int A_PER = BM * BK / NT;
  constexpr int B_PER = (BK * BN + NT - 1) / NT;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const long long row0 = (long long)blockIdx.x * BM;
  const int col0 = blockIdx.y * BN;
  const int Ktot = a.ntaps * a.Cin;
  const long long TN_out = (long long)a.map.T_out * a.map.N;
  const long long TN_in = (long long)a.map.T_in * a.map.N;
  const long long tap_step = (long long)a.map.t_shift * a.map.N + a.map.tap_row_stride;

  // per-thread A rows (fixed across K chunks)
  long long a_base[A_PER];
  int a_t[A_PER];
#pragma unroll
  for (int j = 0; j < A_PER; ++j) {
    int e = tid + j * NT;
    int m = e / BK;
    long long r = row0 + m;
    if (r < a.rows) {
      row_decode(r, (int)TN_out, a.map.N, TN_in, a_base[j], a_t[j]);
    } else {
      a_t[j] = -1000000;
      a_base[j] = 0;
    }
  }

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < Ktot; k0 += BK) {
#pragma unroll
    for (int j = 0; j < A_PER; ++j) {
      int e = tid + j * NT;
      int kk = e % BK, m = e / BK;
      int gk = k0 + kk;
      float v = 0.f;
      if (gk < Ktot) {
        int tap = gk / a.Cin;
        int c = gk - tap * a.Cin;
        int ti = a_t[j] + a.map.t_shift * tap;
        if (ti >= 0 && ti < a.map.T_in) v = ldf(a.in + (a_base[j] + tap * tap_step) * a.Cin + c);
      }
      As[kk][m] = v;
    }
#pragma unroll
    for (int j = 0; j < B_PER; ++j) {
      int e = tid + j * NT;
      if (e < BK * BN) {
        int n = e % BN, kk = e / BN;
        int gk = k0 + kk, o = col0 + n;
        Bs[kk][n] = (gk < Ktot && o < a.Co) ? __ldg(a.wt + (long long)gk * a.Co + o) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TM], bv[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) av[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    long long r = row0 + ty * TM + i;
    if (r >= a.rows) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int o = col0 + tx * TN + j;
      if (o >= a.Co) continue;
      float v = acc[i][j] + (a.bias ? __ldg(a.bias + o) : 0.f);
      TO* p = a.out + r * a.ldo + o;
      if (a.accumulate) v += ldf(p);
      stf(p, v);
    }
  }
}

template <class TI, class TO>
inline void launch_tapgemm(const TapArgs<TI, TO>& a, cudaStream_t s) {
  if (a.rows == 0 || a.Co == 0) return;
  if constexpr (std::is_same<TI, float>::value && std::is_same<TO, float>::value) {
    if (g_x3 && umma::x3_tapgemm(a, s)) return;
  }
  if (a.Co <= 16) {
    dim3 grid(ceil_div(a.rows, 128), ceil_div(a.Co, 16));
    STGCN_LAUNCH((tapgemm_kernel<TI, TO, 128, 16, 2, 4>), grid, NT, 0, s, a);
  } else {
    dim3 grid(ceil_div(a.rows, 64), ceil_div(a.Co, 64));
    STGCN_LAUNCH((tapgemm_kernel<TI, TO, 64, 64, 4, 4>), grid, NT, 0, s, a);
  }
}

// ---- node contraction ---------------------------------------------------------------------
template <class T>
struct GsoArgs {
  const float* M;      // [N, N] row-major
  int trans;           // 0: out[h] = sum_i M[h,i] x[i];  1: uses M[i,h]
  const T* in;         // [G, N, C]
  const T* aux;        // [G, N, C] or nullptr
  T* out;              // [G, N, C]
  int N, C;
  long long G;
  float alpha, beta;
};

template <class T, int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(NT) gso_kernel(GsoArgs<T> a) {
  constexpr int TX = BN / TN, TY = BM / TM;
  static_assert(TX * TY == NT, "tile/thread mismatch");
  constexpr int A_PER = BM * BK / NT;
  constexpr int B_PER = BK * BN / NT;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const long long col0 = (long long)blockIdx.x * BN;   // flattened (g, c)
  const int h0 = blockIdx.y * BM;
  const long long Jtot = a.G * a.C;

  long long b_base[B_PER];
#pragma unroll
  for (int j = 0; j < B_PER; ++j) {
    int e = tid + j * NT;
    int n = e % BN;
    long long J = col0 + n;
    if (J < Jtot) {
      long long g = J / a.C;
      int c = (int)(J - g * a.C);
      b_base[j] = g * a.N * a.C + c;
    } else {
      b_base[j] = -1;
    }
  }
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < a.N; k0 += BK) {
#pragma unroll
    for (int j = 0; j < A_PER; ++j) {
      int e = tid + j * NT;
      int kk, m;
      if (a.trans) { m = e % BM; kk = e / BM; } else { kk = e % BK; m = e / BK; }
      int h = h0 + m, i = k0 + kk;
      float v = 0.f;
      if (h < a.N && i < a.N) v = a.trans ? __ldg(a.M + (long long)i * a.N + h) : __ldg(a.M + (long long)h * a.N + i);
      As[kk][m] = v;
    }
#pragma unroll
    for (int j = 0; j < B_PER; ++j) {
      int e = tid + j * NT;
      int n = e % BN, kk = e / BN;
      int i = k0 + kk;
      Bs[kk][n] = (b_base[j] >= 0 && i < a.N) ? ldf(a.in + b_base[j] + (long long)i * a.C) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TM], bv[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) av[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    long long J = col0 + tx * TN + j;
    if (J >= Jtot) continue;
    long long g = J / a.C;
    int c = (int)(J - g * a.C);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      int h = h0 + ty * TM + i;
      if (h >= a.N) continue;
      long long idx = (g * a.N + h) * a.C + c;
      float v = a.alpha * acc[i][j];
      if (a.aux) v += a.beta * ldf(a.aux + idx);
      stf(a.out + idx, v);
    }
  }
}

template <class T>
inline void launch_gso(const GsoArgs<T>& a, cudaStream_t s) {
  if (a.G == 0) return;
  if constexpr (std::is_same<T, float>::value) {
    if (g_x3 && umma::x3_gso(a, s)) return;
  }
  dim3 grid(ceil_div(a.G * a.C, 64), ceil_div(a.N, 64));
  STGCN_LAUNCH((gso_kernel<T, 64, 64, 4, 4>), grid, NT, 0, s, a);
}

// ---- weight gradient: dwt[(tap,c) | bias row][o] += sum_r in[row(r,tap), c] * dz[r, o] ----
template <class T>
struct WgradArgs {
  const T* in;         // [*, Cin]
  const T* dz;         // [rows, ldz]
  float* dwt;          // [ntaps*Cin (+1), Co], pre-zeroed, atomically accumulated
  long long rows;
  int Cin, Co, ntaps, ldz;
  int bias_row;        // 1: extra last row = column sums of dz (bias gradient)
  int rows_per_cta;
  RowMap map;
  float* partial;      // optional [gridDim.z][Mtot*Co] scratch: per-CTA partial sums (two-stage reduction) instead of
                       // same-address atomics from ~1000 CTAs, which serialise in L2
};

template <class T, int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(NT) wgrad_kernel(WgradArgs<T> a) {
  constexpr int TX = BN / TN, TY = BM / TM;
  static_assert(TX * TY == NT, "tile/thread mismatch");
  constexpr int A_PER = BM * BK / NT;
  constexpr int B_PER = (BK * BN + NT - 1) / NT;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  __shared__ long long row_base[BK];   // input row of tap 0 for each K row of the chunk (-1: past the end)
  __shared__ int row_t[BK];
  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int m0 = blockIdx.x * BM, col0 = blockIdx.y * BN;
  const long long r_begin = (long long)blockIdx.z * a.rows_per_cta;
  const long long r_end = min(a.rows, r_begin + a.rows_per_cta);
  const int Kw = a.ntaps * a.Cin;
  const int Mtot = Kw + (a.bias_row ? 1 : 0);
  const long long TN_out = (long long)a.map.T_out * a.map.N;
  const long long TN_in = (long long)a.map.T_in * a.map.N;
  const long long tap_step = (long long)a.map.t_shift * a.map.N + a.map.tap_row_stride;

  // A-operand columns handled by this thread: m = e % BM is fixed per j
  int a_tap[A_PER], a_c[A_PER];
#pragma unroll
  for (int j = 0; j < A_PER; ++j) {
    int e = tid + j * NT;
    int mm = m0 + e % BM;
    if (mm < Kw) { a_tap[j] = mm / a.Cin; a_c[j] = mm - a_tap[j] * a.Cin; }
    else if (mm == Kw && a.bias_row) { a_tap[j] = -1; a_c[j] = 0; }
    else { a_tap[j] = -2; a_c[j] = 0; }
  }
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (long long k0 = r_begin; k0 < r_end; k0 += BK) {
    if (tid < BK) {      // one row decode per K row (not per element): the divisions dominated this kernel
      long long r = k0 + tid;
      if (r < r_end) {
        long long base; int t;
        row_decode(r, (int)TN_out, a.map.N, TN_in, base, t);
        row_t[tid] = t;
        row_base[tid] = base;
      } else {
        row_t[tid] = 0;
        row_base[tid] = -1;
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < A_PER; ++j) {
      int e = tid + j * NT;
      int m = e % BM, kk = e / BM;
      float v = 0.f;
      const long long rb = row_base[kk];
      if (rb >= 0 && a_tap[j] >= -1) {
        if (a_tap[j] == -1) {
          v = 1.f;
        } else {
          int ti = row_t[kk] + a.map.t_shift * a_tap[j];
          if (ti >= 0 && ti < a.map.T_in) v = ldf(a.in + (rb + a_tap[j] * tap_step) * a.Cin + a_c[j]);
        }
      }
      As[kk][m] = v;
    }
#pragma unroll
    for (int j = 0; j < B_PER; ++j) {
      int e = tid + j * NT;
      if (e < BK * BN) {
        int n = e % BN, kk = e / BN;
        long long r = k0 + kk;
        int o = col0 + n;
        Bs[kk][n] = (r < r_end && o < a.Co) ? ldf(a.dz + r * a.ldz + o) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TM], bv[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) av[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int mm = m0 + ty * TM + i;
    if (mm >= Mtot) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int o = col0 + tx * TN + j;
      if (o >= a.Co) continue;
      if (a.partial) a.partial[((long long)blockIdx.z * Mtot + mm) * a.Co + o] = acc[i][j];
      else atomicAdd(a.dwt + (long long)mm * a.Co + o, acc[i][j]);
    }
  }
}

// out[e] += sum_c partial[c][e]: one warp per 8 elements x 4 lanes-groups; lanes stride over the chunks, shuffle-reduce
__global__ void reduce_partials_kernel(const float* partial, float* out, int n, int chunks) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = blockIdx.x * (blockDim.x >> 5) + warp;
  if (e >= n) return;
  float s = 0.f;
  for (int c = lane; c < chunks; c += 32) s += partial[(long long)c * n + e];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[e] += s;
}
inline void launch_reduce_partials(const float* partial, float* out, int n, int chunks, cudaStream_t s) {
  if (n == 0 || chunks == 0) return;
  STGCN_LAUNCH(reduce_partials_kernel, ceil_div(n, 8), 256, 0, s, partial, out, n, chunks);
}

// ---- skinny weight gradient (Co <= 16: align convs, Chebyshev mix, fc2) -------------------------------------------
// dwt[(tap,c) | bias][o] = sum_r in[row(r,tap), c] * dz[r, o].  The output is tiny (<= 1k elements) and the work is
// streaming rows, so: one CTA per row range, a 32-row tile of both operands staged in shared memory with coalesced
// loads, thread (m, o-quad) accumulates 4 outputs in registers (2 LDS + 4 FMA per row), per-CTA partials are written
// out and reduced by reduce_partials_kernel (no same-address atomics).
constexpr int kSkR = 128;     // rows per tile: the three block barriers per tile were the bottleneck at 32
template <class T, bool VEC>
__global__ void __launch_bounds__(320) wgrad_skinny_kernel(WgradArgs<T> a) {
  extern __shared__ __align__(16) float sk[];
  const int Kw = a.ntaps * a.Cin;
  const int Mtot = Kw + (a.bias_row ? 1 : 0);
  const int Mp = (Mtot + 3) & ~3;              // row pitch of the A tile (16-byte aligned rows)
  const int OG = (a.Co + 3) / 4;
  float* As = sk;                              // [kSkR][Mp]
  float* Bs = sk + kSkR * Mp;                  // [kSkR][16]
  long long* rbase = reinterpret_cast<long long*>(Bs + kSkR * 16);   // [kSkR]
  int* rt = reinterpret_cast<int*>(rbase + kSkR);                    // [kSkR]
  short* tap_of = reinterpret_cast<short*>(rt + kSkR);               // [Kw]
  short* c_of = tap_of + Kw;                                         // [Kw]
  const int tid = threadIdx.x;
  for (int i = tid; i < Kw; i += blockDim.x) { tap_of[i] = (short)(i / a.Cin); c_of[i] = (short)(i % a.Cin); }
  if (a.bias_row) for (int r = tid; r < kSkR; r += blockDim.x) As[r * Mp + Kw] = 1.f;   // bias column: constant ones
  const long long r_begin = (long long)blockIdx.x * a.rows_per_cta;
  const long long r_end = min(a.rows, r_begin + a.rows_per_cta);
  const long long TN_in = (long long)a.map.T_in * a.map.N;
  const int TN_out = a.map.T_out * a.map.N;
  const long long tap_step = (long long)a.map.t_shift * a.map.N + a.map.tap_row_stride;
  const int m = tid / OG, og = tid % OG;
  const bool active = m < Mtot;
  const int kw8 = Kw >> 3;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long k0 = r_begin; k0 < r_end; k0 += kSkR) {
    __syncthreads();
    if (tid < kSkR) {
      long long r = k0 + tid;
      if (r < r_end) { row_decode(r, TN_out, a.map.N, TN_in, rbase[tid], rt[tid]); }
      else { rbase[tid] = -1; rt[tid] = 0; }
    }
    __syncthreads();
    if (VEC) {
      // 8 channels (16 bytes of bf16) per load: Cin % 8 == 0, so a vector never straddles a tap
      for (int v = tid; v < kSkR * kw8; v += blockDim.x) {
        const int r = v / kw8, m8 = (v - r * kw8) << 3;
        float x8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x8[i] = 0.f;
        const long long rb = rbase[r];
        if (rb >= 0) {
          const int tap = tap_of[m8];
          const int ti = rt[r] + a.map.t_shift * tap;
          if (ti >= 0 && ti < a.map.T_in) load8(a.in + (rb + tap * tap_step) * a.Cin + c_of[m8], x8);
        }
        float4* dst = reinterpret_cast<float4*>(As + r * Mp + m8);
        dst[0] = make_float4(x8[0], x8[1], x8[2], x8[3]);
        dst[1] = make_float4(x8[4], x8[5], x8[6], x8[7]);
      }
      for (int v = tid; v < kSkR * 2; v += blockDim.x) {      // Co == 16 here: two vectors per row
        const int r = v >> 1, o8 = (v & 1) << 3;
        const long long rr = k0 + r;
        float x8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x8[i] = 0.f;
        if (rr < r_end) load8(a.dz + rr * a.ldz + o8, x8);
        float4* dst = reinterpret_cast<float4*>(Bs + r * 16 + o8);
        dst[0] = make_float4(x8[0], x8[1], x8[2], x8[3]);
        dst[1] = make_float4(x8[4], x8[5], x8[6], x8[7]);
      }
    } else {
      for (int e = tid; e < kSkR * Kw; e += blockDim.x) {
        const int r = e / Kw, mm = e - r * Kw;
        float v = 0.f;
        const long long rb = rbase[r];
        if (rb >= 0) {
          const int tap = tap_of[mm];
          const int ti = rt[r] + a.map.t_shift * tap;
          if (ti >= 0 && ti < a.map.T_in) v = ldf(a.in + (rb + tap * tap_step) * a.Cin + c_of[mm]);
        }
        As[r * Mp + mm] = v;
      }
      for (int e = tid; e < kSkR * 16; e += blockDim.x) {
        const int r = e >> 4, o = e & 15;
        const long long rr = k0 + r;
        Bs[e] = (rr < r_end && o < a.Co) ? ldf(a.dz + rr * a.ldz + o) : 0.f;
      }
    }
    __syncthreads();
    if (active) {
#pragma unroll 8
      for (int r = 0; r < kSkR; ++r) {
        const float av = As[r * Mp + m];
        const float4 bv = *reinterpret_cast<const float4*>(Bs + r * 16 + og * 4);
        acc[0] = fmaf(av, bv.x, acc[0]); acc[1] = fmaf(av, bv.y, acc[1]);
        acc[2] = fmaf(av, bv.z, acc[2]); acc[3] = fmaf(av, bv.w, acc[3]);
      }
    }
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int o = og * 4 + i;
      if (o < a.Co) a.partial[((long long)blockIdx.x * Mtot + m) * a.Co + o] = acc[i];
    }
  }
}
inline size_t wgrad_skinny_smem(int Mtot, int Kw) {
  return (size_t)kSkR * ((Mtot + 3) & ~3) * 4 + kSkR * 16 * 4 + kSkR * 8 + kSkR * 4 + (size_t)Kw * 4 + 16;
}

struct WgradPlanSimt { int mode, tiles, chunks; long long rpc; };
inline WgradPlanSimt plan_wgrad_simt(long long rows, int Mtot, int Co) {
  WgradPlanSimt pl;
  if (Co <= 16 && Mtot * ((Co + 3) / 4) <= 320 && wgrad_skinny_smem(Mtot, Mtot) <= 160 * 1024) {     // skinny kernel: one CTA per row range, ~6 CTAs per SM
    pl.mode = 4; pl.tiles = 1;
    long long rpc = (rows + 148 * 6 - 1) / (148 * 6);
    if (rpc < 4 * kSkR) rpc = 4 * kSkR;
    pl.rpc = (rpc + kSkR - 1) / kSkR * kSkR;
    pl.chunks = rows > 0 ? ceil_div(rows, pl.rpc) : 0;
    return pl;
  }
  if (Co <= 16) { pl.mode = Mtot <= 64 ? 0 : 1; pl.tiles = ceil_div(Mtot, pl.mode == 0 ? 64 : 128); }
  else if (Mtot <= 16) { pl.mode = 2; pl.tiles = ceil_div(Co, 64); }
  else { pl.mode = 3; pl.tiles = ceil_div(Mtot, 64) * ceil_div(Co, 64); }
  // aim for ~8 CTAs per SM in total, at least 256 rows per CTA
  long long want = (148 * 8 + pl.tiles - 1) / pl.tiles;
  long long rpc = (rows + want - 1) / want;
  if (rpc < 256) rpc = 256;
  pl.rpc = (rpc + BK - 1) / BK * BK;
  pl.chunks = rows > 0 ? ceil_div(rows, pl.rpc) : 0;
  return pl;
}
inline size_t wgrad_partial_elems(long long rows, int Mtot, int Co) {
  return (size_t)plan_wgrad_simt(rows, Mtot, Co).chunks * Mtot * Co;
}

template <class T>
inline void launch_wgrad(WgradArgs<T> a, cudaStream_t s) {
  if (a.rows == 0 || a.Co == 0) return;
  if constexpr (std::is_same<T, float>::value) {
    if (g_x3 && umma::x3_wgrad(a, s)) return;
  }
  const int Mtot = a.ntaps * a.Cin + (a.bias_row ? 1 : 0);
  const WgradPlanSimt pl = plan_wgrad_simt(a.rows, Mtot, a.Co);
  a.rows_per_cta = (int)pl.rpc;
  const int chunks = pl.chunks;
  if (pl.mode == 4 && a.partial) {
    const size_t smem = wgrad_skinny_smem(Mtot, a.ntaps * a.Cin);
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    const bool vec = a.Cin % 8 == 0 && a.Co == 16 && a.ldz % 8 == 0 && al16(a.in) && al16(a.dz) &&
                     (a.map.tap_row_stride * a.Cin) % 8 == 0;
    if (vec) {
      STGCN_CUDA(cudaFuncSetAttribute(wgrad_skinny_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      STGCN_LAUNCH((wgrad_skinny_kernel<T, true>), chunks, 320, smem, s, a);
    } else {
      STGCN_CUDA(cudaFuncSetAttribute(wgrad_skinny_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      STGCN_LAUNCH((wgrad_skinny_kernel<T, false>), chunks, 320, smem, s, a);
    }
  } else if (pl.mode == 4 || pl.mode == 0) {
    STGCN_LAUNCH((wgrad_kernel<T, 64, 16, 1, 4>), dim3(ceil_div(Mtot, 64), 1, chunks), NT, 0, s, a);
  } else if (pl.mode == 1) {
    STGCN_LAUNCH((wgrad_kernel<T, 128, 16, 2, 4>), dim3(ceil_div(Mtot, 128), 1, chunks), NT, 0, s, a);
  } else if (pl.mode == 2) {
    STGCN_LAUNCH((wgrad_kernel<T, 16, 64, 1, 4>), dim3(1, ceil_div(a.Co, 64), chunks), NT, 0, s, a);
  } else {
    STGCN_LAUNCH((wgrad_kernel<T, 64, 64, 4, 4>), dim3(ceil_div(Mtot, 64), ceil_div(a.Co, 64), chunks), NT, 0, s, a);
  }
  if (a.partial) launch_reduce_partials(a.partial, a.dwt, Mtot * a.Co, chunks, s);
}

template <bool FAST = false>
__device__ __forceinline__ float act_fwd(int act, float u, float q) {
  if (act == STGCN_ACT_GLU) return u * sigmoid_t<FAST>(q);
  if (act == STGCN_ACT_GTU) return tanh_t<FAST>(u) * sigmoid_t<FAST>(q);
  if (act == STGCN_ACT_RELU) return fmaxf(u, 0.f);
  if (act == STGCN_ACT_SILU) return u * sigmoid_t<FAST>(u);
  return u;
}
// gradients of act_fwd w.r.t. (u, q) times g
template <bool FAST = false>
__device__ __forceinline__ void act_bwd(int act, float u, float q, float g, float& du, float& dq) {
  dq = 0.f;
  if (act == STGCN_ACT_GLU) { float s = sigmoid_t<FAST>(q); du = g * s; dq = g * u * s * (1.f - s); }
  else if (act == STGCN_ACT_GTU) { float s = sigmoid_t<FAST>(q), th = tanh_t<FAST>(u); du = g * s * (1.f - th * th); dq = g * th * s * (1.f - s); }
  else if (act == STGCN_ACT_RELU) { du = u > 0.f ? g : 0.f; }
  else if (act == STGCN_ACT_SILU) { float s = sigmoid_t<FAST>(u); du = g * (s + u * s * (1.f - s)); }
  else { du = g; }
}
// bf16 storage mode uses the one-MUFU sigmoid / tanh (common.cuh)
template <class T> constexpr bool kFastAct = std::is_same<T, bf16>::value;

// ---- gating / activation of the temporal conv (layers.py:92-115) ---------------------------
template <class T>
struct GateArgs {
  const T* z;          // [rows, W]   pre-activation (W = 2*Cout for glu/gtu, Cout otherwise)
  const T* xin;        // [*, Cin]    layer input (residual source), or nullptr when folded
  const T* dy;         // bwd: [rows, Cout]
  T* y;                // fwd: [rows, Cout]
  T* dz;               // bwd: [rows, W]
  long long rows;
  int Cin, Cout, W, Kt;
  int T_out, T_in, N;
  int explicit_res;    // 1: residual = xin[(b,t+Kt-1,n), j] for j < Cin (zero pad / identity)
  // bwd, GLU "q-only" saved state: z holds just the gate half Q as [rows, Cout]; with the layer OUTPUT h = u * sigma(Q)
  // the gradients are du = dy * s, dq = dy * h * (1 - s) -- neither P nor the residual is needed
  const T* h; int q_only;
};
constexpr int kGateLrC = 16;      // channel count of the low-rank expansion kernel below

template <class T>
__device__ __forceinline__ float gate_residual(const GateArgs<T>& a, long long r, int j) {
  if (!a.explicit_res || j >= a.Cin) return 0.f;
  long long base; int t;
  row_decode(r, a.T_out * a.N, a.N, (long long)a.T_in * a.N, base, t);
  return ldf(a.xin + (base + (long long)(a.Kt - 1) * a.N) * a.Cin + j);
}

template <class T, int ACT>
__global__ void gate_fwd_kernel(GateArgs<T> a) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.rows * a.Cout) return;
  long long r = (long long)((unsigned long long)idx / (unsigned)a.Cout);
  int j = (int)(idx - r * a.Cout);
  float res = gate_residual(a, r, j);
  float p = ldf(a.z + r * a.W + j) + res;
  float out;
  if (ACT == STGCN_ACT_GLU) {
    out = p * sigmoidf_(ldf(a.z + r * a.W + a.Cout + j));
  } else if (ACT == STGCN_ACT_GTU) {
    out = tanhf(p) * sigmoidf_(ldf(a.z + r * a.W + a.Cout + j));
  } else if (ACT == STGCN_ACT_RELU) {
    out = fmaxf(p, 0.f);
  } else if (ACT == STGCN_ACT_SILU) {
    out = p * sigmoidf_(p);
  } else {
    out = p;
  }
  stf(a.y + idx, out);
}

template <class T, int ACT>
__global__ void gate_bwd_kernel(GateArgs<T> a) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.rows * a.Cout) return;
  long long r = idx / a.Cout;
  int j = (int)(idx - r * a.Cout);
  float res = gate_residual(a, r, j);
  float p = ldf(a.z + r * a.W + j) + res;
  float g = ldf(a.dy + idx);
  if (ACT == STGCN_ACT_GLU) {
    float s = sigmoidf_(ldf(a.z + r * a.W + a.Cout + j));
    stf(a.dz + r * a.W + j, g * s);
    stf(a.dz + r * a.W + a.Cout + j, g * p * s * (1.f - s));
  } else if (ACT == STGCN_ACT_GTU) {
    float s = sigmoidf_(ldf(a.z + r * a.W + a.Cout + j));
    float th = tanhf(p);
    stf(a.dz + r * a.W + j, g * s * (1.f - th * th));
    stf(a.dz + r * a.W + a.Cout + j, g * th * s * (1.f - s));
  } else if (ACT == STGCN_ACT_RELU) {
    stf(a.dz + r * a.W + j, p > 0.f ? g : 0.f);
  } else if (ACT == STGCN_ACT_SILU) {
    float s = sigmoidf_(p);
    stf(a.dz + r * a.W + j, g * (s + p * s * (1.f - s)));
  } else {
    stf(a.dz + r * a.W + j, g);
  }
}

// dx[(b, t+Kt-1, n), j] += dz[(b,t,n), j]  for j < min(Cin, Cout)  (gradient of the explicit residual)
template <class T>
__global__ void residual_add_kernel(const T* dz, T* dx, long long rows, int Cres, int W, int Cin, int Kt,
                                    int T_out, int T_in, int N) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * Cres) return;
  long long r = idx / Cres;
  int j = (int)(idx - r * Cres);
  long long base; int t;
  row_decode(r, T_out * N, N, (long long)T_in * N, base, t);
  long long row = base + (long long)(Kt - 1) * N;
  stf(dx + row * Cin + j, ldf(dx + row * Cin + j) + ldf(dz + r * W + j));
}

// 8 channels per thread (requires Cout, W and Cin to be multiples of 8 when a residual is read)
template <class T, int ACT>
__global__ void gate_vec_kernel(GateArgs<T> a, int bwd) {
  const int groups = a.Cout / 8;
  const long long total = a.rows * groups;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // one 8-channel chunk per thread
  if (idx >= total) return;
  const long long r = (long long)((unsigned)idx / (unsigned)groups);    // rows * groups < 2^31 (launcher)
  const int j0 = (int)(idx - r * groups) * 8;
  constexpr bool gated = ACT == STGCN_ACT_GLU || ACT == STGCN_ACT_GTU;
  float zp[8], zq[8], res[8];
  const bool q_only = ACT == STGCN_ACT_GLU && a.q_only;
  if (q_only) {
    load8(a.z + r * a.Cout + j0, zq);
    load8(a.h + r * a.Cout + j0, zp);          // zp carries h here
  } else {
    load8(a.z + r * a.W + j0, zp);
    if (gated) load8(a.z + r * a.W + a.Cout + j0, zq);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) res[i] = 0.f;
  if (!q_only && a.explicit_res && j0 < a.Cin) {
    long long base; int t;
    row_decode(r, a.T_out * a.N, a.N, (long long)a.T_in * a.N, base, t);
    load8(a.xin + (base + (long long)(a.Kt - 1) * a.N) * a.Cin + j0, res);      // Cin % 8 == 0 is checked by the launcher
  }
  if (!bwd) {
    float h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = act_fwd<kFastAct<T>>(ACT, zp[i] + res[i], gated ? zq[i] : 0.f);
    store8(a.y + r * a.Cout + j0, h);
  } else {
    float g[8], du[8], dq[8];
    load8(a.dy + r * a.Cout + j0, g);
    if (q_only) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float sg = sigmoid_t<kFastAct<T>>(zq[i]);
        du[i] = g[i] * sg;
        dq[i] = g[i] * zp[i] * (1.f - sg);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) act_bwd<kFastAct<T>>(ACT, zp[i] + res[i], gated ? zq[i] : 0.f, g[i], du[i], dq[i]);
    }
    store8(a.dz + r * a.W + j0, du);
    if (gated) store8(a.dz + r * a.W + a.Cout + j0, dq);
  }
}

// does the 8-channels-per-thread kernel serve these arguments?  (tconv_bwd asks before it commits to the q-only state)
template <class T>
inline bool gate_vec_ok(const GateArgs<T>& a) {
  const long long n = a.rows * a.Cout;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  return a.Cout % 8 == 0 && a.W % 8 == 0 && n / 8 < (1LL << 31) && (!a.explicit_res || a.Cin % 8 == 0) && al16(a.z) && al16(a.xin) &&
         al16(a.dy) && al16(a.y) && al16(a.dz) && al16(a.h);
}
template <class T, int ACT>
inline void launch_gate(bool bwd, const GateArgs<T>& a, cudaStream_t s) {
  long long n = a.rows * a.Cout;
  if (n == 0) return;
  if (gate_vec_ok(a)) {
    STGCN_LAUNCH((gate_vec_kernel<T, ACT>), ceil_div(n / 8, 256), 256, 0, s, a, bwd ? 1 : 0);
    return;
  }
  STGCN_CHECK(!a.q_only, STGCN_E_UNSUPPORTED, "the q-only saved state needs the vectorised gate kernel");
  if (bwd) STGCN_LAUNCH((gate_bwd_kernel<T, ACT>), ceil_div(n, 256), 256, 0, s, a);
  else     STGCN_LAUNCH((gate_fwd_kernel<T, ACT>), ceil_div(n, 256), 256, 0, s, a);
}

// out[r, j0..j0+7] = sum_{o<16} src[r, o] * w[o * Cout + j]: data gradient of a 1x1 conv with 16 output channels
// (the graph-conv layer's align conv, layers.py:16,225) as plain FMAs -- with K = 16 the tensor-core tile kernel spends
// its time in per-tile epilogue bookkeeping (66 us for 75 MB), this is bandwidth work.  8 channels per thread.
template <class T>
__global__ void __launch_bounds__(256) lowrank_expand_kernel(const T* src, const float* w, T* out, long long rows, int Cout) {
  // thread = (8-channel group, row lane); its 16 x 8 weights live in REGISTERS for the whole (persistent) kernel: from
  // shared memory the 32 LDS.128 per 128 FMAs made the kernel shared-memory-bandwidth bound (146 us for 37 MB,
  // profiles/r01_ab_batch_h.md); the rows' 32-byte inputs are broadcast through L1 to the 8 threads that share a row
  const int groups = Cout / 8;
  const int g = threadIdx.x % groups, lane = threadIdx.x / groups, lanes = blockDim.x / groups;
  const int j0 = g * 8;
  float wr[kGateLrC][8];
#pragma unroll
  for (int o = 0; o < kGateLrC; ++o) {
    const float4 w0 = *reinterpret_cast<const float4*>(w + o * Cout + j0);
    const float4 w1 = *reinterpret_cast<const float4*>(w + o * Cout + j0 + 4);
    wr[o][0] = w0.x; wr[o][1] = w0.y; wr[o][2] = w0.z; wr[o][3] = w0.w;
    wr[o][4] = w1.x; wr[o][5] = w1.y; wr[o][6] = w1.z; wr[o][7] = w1.w;
  }
  const long long stride = (long long)gridDim.x * lanes;
  long long r = (long long)blockIdx.x * lanes + lane;
  uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0;                     // next row's 16 inputs, requested one iteration ahead
  if (r < rows) { n0 = reinterpret_cast<const uint4*>(src + r * kGateLrC)[0]; n1 = reinterpret_cast<const uint4*>(src + r * kGateLrC)[1]; }
  for (; r < rows; r += stride) {
    float d[kGateLrC], v[8];
    unpack8(n0, d); unpack8(n1, d + 8);
    if (r + stride < rows) {
      n0 = reinterpret_cast<const uint4*>(src + (r + stride) * kGateLrC)[0];
      n1 = reinterpret_cast<const uint4*>(src + (r + stride) * kGateLrC)[1];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
#pragma unroll
    for (int o = 0; o < kGateLrC; ++o)
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaf(d[o], wr[o][i], v[i]);
    store8(out + r * Cout + j0, v);
  }
}
template <class T>
inline bool lowrank_expand_supported(const T* src, const float* w, const T* out, long long rows, int Csrc, int Cout) {
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  return Csrc == kGateLrC && Cout % 8 == 0 && 256 % (Cout / 8) == 0 && rows > 0 && al16(src) && al16(w) && al16(out);
}
template <class T>
inline void launch_lowrank_expand(const T* src, const float* w, T* out, long long rows, int Cout, cudaStream_t s) {
  const int lanes = 256 / (Cout / 8);
  const int blocks = (int)std::min<long long>(ceil_div(rows, lanes), 148);       // 154 registers: one CTA per SM
  STGCN_LAUNCH(lowrank_expand_kernel<T>, blocks, 256, 0, s, src, w, out, rows, Cout);
}

template <class T>
inline void launch_gate_any(int act, bool bwd, const GateArgs<T>& a, cudaStream_t s) {
  switch (act) {
    case STGCN_ACT_GLU:  launch_gate<T, STGCN_ACT_GLU>(bwd, a, s); break;
    case STGCN_ACT_GTU:  launch_gate<T, STGCN_ACT_GTU>(bwd, a, s); break;
    case STGCN_ACT_RELU: launch_gate<T, STGCN_ACT_RELU>(bwd, a, s); break;
    case STGCN_ACT_SILU: launch_gate<T, STGCN_ACT_SILU>(bwd, a, s); break;
    case STGCN_ACT_LINEAR: launch_gate<T, STGCN_ACT_LINEAR>(bwd, a, s); break;
    default: throw Error(STGCN_E_UNSUPPORTED, "activation not implemented");
  }
}

// ---- single-output linear layer (fc2 of the output block, end_channel = 1; layers.py:271,281) ---------------------
// forward: y[r] = b + sum_c in[r,c] w[c];  data gradient: din[r,c] = dy[r] w[c];  weight gradient: dw[c] = sum_r dy[r] in[r,c],
// db = sum_r dy[r].  All three are bandwidth work over a [rows, C] tensor: 8 channels per thread, 16-byte accesses.
template <class T>
__global__ void __launch_bounds__(256) rowdot_fwd_kernel(const T* in, const float* w, const float* b, float* y, long long rows, int C) {
  const int G = C / 8, g = threadIdx.x % G, rl = threadIdx.x / G, lanes = blockDim.x / G;   // G = power of two <= 32
  float wv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) wv[i] = w[g * 8 + i];
  const float bias = b ? b[0] : 0.f;
  for (long long r = (long long)blockIdx.x * lanes + rl; r < rows; r += (long long)gridDim.x * lanes) {
    float x[8];
    load8(in + r * C + g * 8, x);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(x[i], wv[i], s);
    for (int o = 1; o < G; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (g == 0) y[r] = s + bias;
  }
}
// relu_ref (optional): the ReLU that sat in front of this layer's input is applied on the way out,
// din[r, c] = relu_ref[r, c] > 0 ? dy[r] * w[c] : 0 (no dropout): one pass instead of a [rows, C] round trip through a
// separate ReLU-backward kernel.
template <class T>
__global__ void __launch_bounds__(256) rowouter_bwd_kernel(const float* dy, const float* w, T* din, long long rows, int C,
                                                           const T* relu_ref) {
  const int G = C / 8;
  const long long total = rows * G;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / G;
    const int g = (int)(idx - r * G);
    const float d = dy[r];
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = d * w[g * 8 + i];
    if (relu_ref) {
      float f[8];
      load8(relu_ref + r * C + g * 8, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = f[i] > 0.f ? v[i] : 0.f;
    }
    store8(din + r * C + g * 8, v);
  }
}
// partial[cta][c] = sum over the CTA's rows of dy[r]*in[r,c]; partial[cta][C] = sum dy[r]
template <class T>
__global__ void __launch_bounds__(256) rowdot_wgrad_kernel(const T* in, const float* dy, float* partial, long long rows, int C,
                                                           int rows_per_cta) {
  __shared__ float red[8][257];
  const int G = C / 8, g = threadIdx.x % G, rl = threadIdx.x / G, lanes = blockDim.x / G;
  float acc[8], accb = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  for (long long r = r0 + rl; r < r1; r += lanes) {
    float x[8];
    load8(in + r * C + g * 8, x);
    const float d = dy[r];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(d, x[i], acc[i]);
    if (g == 0) accb += d;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    for (int o = G; o < 32; o <<= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
  for (int o = G; o < 32; o <<= 1) accb += __shfl_xor_sync(0xffffffffu, accb, o);
  if (lane < G) {
#pragma unroll
    for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = acc[i];
    if (lane == 0) red[warp][C] = accb;
  }
  __syncthreads();
  for (int e = threadIdx.x; e <= C; e += blockDim.x) {
    float v = 0.f;
    for (int w2 = 0; w2 < (int)(blockDim.x >> 5); ++w2) v += red[w2][e];
    partial[(long long)blockIdx.x * (C + 1) + e] = v;
  }
}
inline bool rowdot_supported(int C) { return C == 8 || C == 16 || C == 32 || C == 64 || C == 128 || C == 256; }

// ---- first-layer special: temporal conv with a tiny input width (Cin <= 4; the model input has Cin = 1) ------
// The GEMM has K = Kt*Cin <= 16, so it is bandwidth work: one fused pass computes conv + bias + gate and writes the
// saved pre-activation Z and the output H (8 output channels per thread, 16/32-byte stores).
template <class T>
struct SmallCArgs {
  const T* x;          // [B, T_in, N, Cin]
  const float* wt;     // [(k*Cin + c)][W]
  const float* bias;   // [W]
  T* z;                // [rows, W]
  T* h;                // [rows, Cout]
  const T* dh;         // bwd: [rows, Cout]
  T* dz;               // bwd (optional): [rows, W]
  float* dwt;          // bwd: [(Kt*Cin + 1)][W], pre-zeroed
  float* partial;      // bwd, optional: [gridDim.x][(Kt*Cin + 1)*W] per-CTA partial sums (reduced by reduce_partials_kernel)
  long long rows;
  int Cin, Cout, W, Kt, T_out, T_in, N, act, explicit_res, rows_per_cta;
  int skip_z;          // Cin == 1 fast path: z is not stored by the forward and recomputed from x by the backward
};


template <class T>
__global__ void __launch_bounds__(256) smallc_conv_gate_fwd_kernel(SmallCArgs<T> a) {
  extern __shared__ float sm[];          // wt [K][W] then bias [W]
  const int K = a.Kt * a.Cin;
  for (int i = threadIdx.x; i < K * a.W; i += blockDim.x) sm[i] = a.wt[i];
  for (int i = threadIdx.x; i < a.W; i += blockDim.x) sm[K * a.W + i] = a.bias[i];
  __syncthreads();
  const float* w = sm;
  const float* bs = sm + K * a.W;
  const int groups = a.Cout / 8;
  const bool gated = a.W == 2 * a.Cout;
  const long long total = a.rows * groups;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = (long long)((unsigned long long)idx / (unsigned)groups);
    const int j0 = (int)(idx - r * groups) * 8;
    long long in0; int t_unused;
    row_decode(r, a.T_out * a.N, a.N, (long long)a.T_in * a.N, in0, t_unused);
    float zp[8], zq[8], hv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { zp[i] = bs[j0 + i]; zq[i] = gated ? bs[a.Cout + j0 + i] : 0.f; }
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      if (kk < K) {
        const int k = kk / a.Cin, c = kk - k * a.Cin;
        const float xk = ldf(a.x + (in0 + (long long)k * a.N) * a.Cin + c);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          zp[i] = fmaf(xk, w[kk * a.W + j0 + i], zp[i]);
          if (gated) zq[i] = fmaf(xk, w[kk * a.W + a.Cout + j0 + i], zq[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float res = (a.explicit_res && j0 + i < a.Cin) ? ldf(a.x + (in0 + (long long)(a.Kt - 1) * a.N) * a.Cin + j0 + i) : 0.f;
      hv[i] = act_fwd(a.act, zp[i] + res, zq[i]);
    }
    store8(a.z + r * a.W + j0, zp);
    if (gated) store8(a.z + r * a.W + a.Cout + j0, zq);
    if (a.h) store8(a.h + r * a.Cout + j0, hv);
  }
}

// Cin == 1 specialisation: the K = Kt weights of this thread's 8 (+8 gate) channels live in registers; each thread
// walks rows with a grid stride (its channel group never changes), so the loop body is K loads of x, K*16 FMAs, the
// gate, and three 16-byte stores.
template <class T, int K, int ACT>
__global__ void __launch_bounds__(256) smallc1_conv_gate_fwd_kernel(SmallCArgs<T> a) {
  const int groups = a.Cout / 8;
  const bool gated = a.W == 2 * a.Cout;
  const int j0 = (threadIdx.x % groups) * 8;
  const int rl = threadIdx.x / groups, lanes = blockDim.x / groups;
  float wp[K][8], wq[K][8], bp[8], bq[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    bp[i] = a.bias[j0 + i];
    bq[i] = gated ? a.bias[a.Cout + j0 + i] : 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      wp[k][i] = a.wt[k * a.W + j0 + i];
      wq[k][i] = gated ? a.wt[k * a.W + a.Cout + j0 + i] : 0.f;
    }
  }
  for (long long r = (long long)blockIdx.x * lanes + rl; r < a.rows; r += (long long)gridDim.x * lanes) {
    long long in0; int t_unused;
    row_decode(r, a.T_out * a.N, a.N, (long long)a.T_in * a.N, in0, t_unused);
    float xv[K];
#pragma unroll
    for (int k = 0; k < K; ++k) xv[k] = ldf(a.x + in0 + (long long)k * a.N);
    float zp[8], zq[8], hv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float p = bp[i], q = bq[i];
#pragma unroll
      for (int k = 0; k < K; ++k) { p = fmaf(xv[k], wp[k][i], p); q = fmaf(xv[k], wq[k][i], q); }
      zp[i] = p; zq[i] = q;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float res = (a.explicit_res && j0 + i == 0) ? xv[K - 1] : 0.f;
      hv[i] = act_fwd<kFastAct<T>>(ACT, zp[i] + res, zq[i]);
    }
    if (!a.skip_z) {
      store8(a.z + r * a.W + j0, zp);
      if (gated) store8(a.z + r * a.W + a.Cout + j0, zq);
    }
    store8(a.h + r * a.Cout + j0, hv);
  }
}
template <class T>
inline void launch_smallc1_conv_gate_fwd(const SmallCArgs<T>& a, cudaStream_t s) {
  const int lanes = 256 / (a.Cout / 8);
  const int blocks = (int)std::min<long long>(ceil_div(a.rows, lanes), 148 * 8);
  // the activation is a template parameter: a runtime switch inside the 8-wide unrolled gate cost a branch chain per element
#define STGCN_SC1F(ACTV) do { \
    if (a.Kt == 2) STGCN_LAUNCH((smallc1_conv_gate_fwd_kernel<T, 2, ACTV>), blocks, 256, 0, s, a); \
    else if (a.Kt == 3) STGCN_LAUNCH((smallc1_conv_gate_fwd_kernel<T, 3, ACTV>), blocks, 256, 0, s, a); \
    else STGCN_LAUNCH((smallc1_conv_gate_fwd_kernel<T, 4, ACTV>), blocks, 256, 0, s, a); } while (0)
  switch (a.act) {
    case STGCN_ACT_GLU: STGCN_SC1F(STGCN_ACT_GLU); break;
    case STGCN_ACT_GTU: STGCN_SC1F(STGCN_ACT_GTU); break;
    case STGCN_ACT_RELU: STGCN_SC1F(STGCN_ACT_RELU); break;
    case STGCN_ACT_SILU: STGCN_SC1F(STGCN_ACT_SILU); break;
    default: STGCN_SC1F(STGCN_ACT_LINEAR); break;
  }
#undef STGCN_SC1F
}

// Fused gate-backward + weight gradient for the same layer: per CTA a row range, thread = (channel j, row lane);
// dW[(k,c)][o] and db[o] accumulate in registers, are reduced across the row lanes in shared memory and added to the
// global buffer with one atomic per element per CTA.  Optionally also materialises dz (when dx is needed).
template <class T>
__global__ void __launch_bounds__(256) smallc_gate_wgrad_kernel(SmallCArgs<T> a) {
  extern __shared__ float red[];         // [lanes][2][K+1][Cout]
  const int K = a.Kt * a.Cin;
  const bool gated = a.W == 2 * a.Cout;
  const int lanes = blockDim.x / a.Cout;
  const int j = threadIdx.x % a.Cout, rl = threadIdx.x / a.Cout;
  float accp[17], accq[17];
#pragma unroll
  for (int i = 0; i < 17; ++i) { accp[i] = 0.f; accq[i] = 0.f; }
  const long long r0 = (long long)blockIdx.x * a.rows_per_cta;
  const long long r1 = min(a.rows, r0 + a.rows_per_cta);
  if (rl < lanes) {
    for (long long r = r0 + rl; r < r1; r += lanes) {
      long long in0; int t_unused;
      row_decode(r, a.T_out * a.N, a.N, (long long)a.T_in * a.N, in0, t_unused);
      const float res = (a.explicit_res && j < a.Cin) ? ldf(a.x + (in0 + (long long)(a.Kt - 1) * a.N) * a.Cin + j) : 0.f;
      const float u = ldf(a.z + r * a.W + j) + res;
      const float q = gated ? ldf(a.z + r * a.W + a.Cout + j) : 0.f;
      float du, dq;
      act_bwd(a.act, u, q, ldf(a.dh + r * a.Cout + j), du, dq);
      if (a.dz) {
        stf(a.dz + r * a.W + j, du);
        if (gated) stf(a.dz + r * a.W + a.Cout + j, dq);
      }
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        if (kk < K) {
          const int k = kk / a.Cin, c = kk - k * a.Cin;
          const float xk = ldf(a.x + (in0 + (long long)k * a.N) * a.Cin + c);
          accp[kk] = fmaf(xk, du, accp[kk]);
          accq[kk] = fmaf(xk, dq, accq[kk]);
        }
      }
      accp[16] += du;      // bias gradient
      accq[16] += dq;
    }
  }
  // reduce over row lanes
  const int stride = (K + 1) * a.Cout;
  if (rl < lanes) {
#pragma unroll
    for (int kk = 0; kk < 17; ++kk) {
      const int slot = kk == 16 ? K : kk;        // bias row sits right after the K weight rows
      if (kk == 16 || kk < K) {
        red[(rl * 2 + 0) * stride + slot * a.Cout + j] = accp[kk];
        red[(rl * 2 + 1) * stride + slot * a.Cout + j] = accq[kk];
      }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * stride; e += blockDim.x) {
    const int half = e / stride, rest = e - half * stride;
    if (half == 1 && !gated) continue;
    float v = 0.f;
    for (int l = 0; l < lanes; ++l) v += red[(l * 2 + half) * stride + rest];
    const int kk = rest / a.Cout, jj = rest - kk * a.Cout;
    const long long off = (long long)kk * a.W + half * a.Cout + jj;
    if (a.partial) a.partial[(long long)blockIdx.x * (K + 1) * a.W + off] = v;
    else atomicAdd(a.dwt + off, v);
  }
}

// Cin == 1 specialisation of the above (the model input): 8 channels per thread with 16-byte loads, K = Kt taps known
// at compile time, shuffle + shared-memory reduction, per-CTA partials.
// (A variant that requested the next row's operands one iteration ahead measured 2.5 % slower on the whole step --
// profiles/r02_ab_batch_a.md -- and was removed.)
template <class T, int K, int ACT>
__global__ void __launch_bounds__(256) smallc1_gate_wgrad_kernel(SmallCArgs<T> a) {
  __shared__ float red[8][2 * (K + 1) * 64];       // [warp][half][k][<=64 channels per pass]
  const bool gated = a.W == 2 * a.Cout;
  const int G = a.Cout / 8;                         // channel groups (power of two <= 32 checked by the launcher)
  const int g = threadIdx.x % G, rl = threadIdx.x / G, lanes = blockDim.x / G;
  const int j0 = g * 8;
  float accp[K + 1][8], accq[K + 1][8];
#pragma unroll
  for (int k = 0; k <= K; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) { accp[k][i] = 0.f; accq[k][i] = 0.f; }
  // skip_z: the pre-activations were not stored; they are K FMAs per channel away from x.  The weights sit in shared
  // memory (in registers they cost 64 more per thread and halved the occupancy of this bandwidth-bound kernel).
  __shared__ __align__(16) float w_s[(K + 1) * 128];       // [k][W] then bias [W]; W <= 128
  if (a.skip_z) {
    for (int i = threadIdx.x; i < K * a.W; i += blockDim.x) w_s[i] = a.wt[i];
    for (int i = threadIdx.x; i < a.W; i += blockDim.x) w_s[K * a.W + i] = a.bias[i];
    __syncthreads();
  }
  const long long r0 = (long long)blockIdx.x * a.rows_per_cta;
  const long long r1 = min(a.rows, r0 + a.rows_per_cta);
  for (long long r = r0 + rl; r < r1; r += lanes) {
    long long in0; int t_unused;
    row_decode(r, a.T_out * a.N, a.N, (long long)a.T_in * a.N, in0, t_unused);
    float xv[K];
#pragma unroll
    for (int k = 0; k < K; ++k) xv[k] = ldf(a.x + in0 + (long long)k * a.N);
    float zp[8], zq[8], dh[8], du[8], dq[8];
    if (a.skip_z) {
      const float* bs = w_s + K * a.W;
#pragma unroll
      for (int i = 0; i < 8; ++i) { zp[i] = bs[j0 + i]; zq[i] = gated ? bs[a.Cout + j0 + i] : 0.f; }
#pragma unroll
      for (int k = 0; k < K; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          zp[i] = fmaf(xv[k], w_s[k * a.W + j0 + i], zp[i]);
          if (gated) zq[i] = fmaf(xv[k], w_s[k * a.W + a.Cout + j0 + i], zq[i]);
        }
    } else {
      load8(a.z + r * a.W + j0, zp);
      if (gated) load8(a.z + r * a.W + a.Cout + j0, zq);
    }
    load8(a.dh + r * a.Cout + j0, dh);
    if (a.explicit_res && j0 == 0) zp[0] += xv[K - 1];        // residual = zero-padded input: channel 0 only
#pragma unroll
    for (int i = 0; i < 8; ++i) act_bwd<kFastAct<T>>(ACT, zp[i], gated ? zq[i] : 0.f, dh[i], du[i], dq[i]);
    if (a.dz) {
      store8(a.dz + r * a.W + j0, du);
      if (gated) store8(a.dz + r * a.W + a.Cout + j0, dq);
    }
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int i = 0; i < 8; ++i) { accp[k][i] = fmaf(xv[k], du[i], accp[k][i]); accq[k][i] = fmaf(xv[k], dq[i], accq[k][i]); }
#pragma unroll
    for (int i = 0; i < 8; ++i) { accp[K][i] += du[i]; accq[K][i] += dq[i]; }
  }
  // reduce over the rows held by one warp (lanes with equal g), then over warps through shared memory
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k <= K; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float vp = accp[k][i], vq = accq[k][i];
      for (int o = G; o < 32; o <<= 1) { vp += __shfl_xor_sync(0xffffffffu, vp, o); vq += __shfl_xor_sync(0xffffffffu, vq, o); }
      accp[k][i] = vp; accq[k][i] = vq;
    }
  const int Cp = a.Cout;                            // <= 64 here
  if (lane < G) {
#pragma unroll
    for (int k = 0; k <= K; ++k)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        red[warp][(0 * (K + 1) + k) * 64 + j0 + i] = accp[k][i];
        red[warp][(1 * (K + 1) + k) * 64 + j0 + i] = accq[k][i];
      }
  }
  __syncthreads();
  const int nw = blockDim.x >> 5;
  for (int e = threadIdx.x; e < 2 * (K + 1) * Cp; e += blockDim.x) {
    const int half = e / ((K + 1) * Cp), rest = e - half * (K + 1) * Cp;
    if (half == 1 && !gated) continue;
    const int k = rest / Cp, j = rest - k * Cp;
    float v = 0.f;
    for (int w = 0; w < nw; ++w) v += red[w][(half * (K + 1) + k) * 64 + j];
    a.partial[(long long)blockIdx.x * (K + 1) * a.W + (long long)k * a.W + half * a.Cout + j] = v;
  }
}
template <class T>
inline bool smallc1_supported(int Cin, int Cout, int Kt) {
  return Cin == 1 && Kt >= 2 && Kt <= 4 && (Cout == 8 || Cout == 16 || Cout == 32 || Cout == 64);
}
template <class T>
inline void launch_smallc1_gate_wgrad(const SmallCArgs<T>& a, int ctas, cudaStream_t s) {
#define STGCN_SC1B(ACTV) do { \
    if (a.Kt == 2) STGCN_LAUNCH((smallc1_gate_wgrad_kernel<T, 2, ACTV>), ctas, 256, 0, s, a); \
    else if (a.Kt == 3) STGCN_LAUNCH((smallc1_gate_wgrad_kernel<T, 3, ACTV>), ctas, 256, 0, s, a); \
    else STGCN_LAUNCH((smallc1_gate_wgrad_kernel<T, 4, ACTV>), ctas, 256, 0, s, a); } while (0)
  switch (a.act) {
    case STGCN_ACT_GLU: STGCN_SC1B(STGCN_ACT_GLU); break;
    case STGCN_ACT_GTU: STGCN_SC1B(STGCN_ACT_GTU); break;
    case STGCN_ACT_RELU: STGCN_SC1B(STGCN_ACT_RELU); break;
    case STGCN_ACT_SILU: STGCN_SC1B(STGCN_ACT_SILU); break;
    default: STGCN_SC1B(STGCN_ACT_LINEAR); break;
  }
#undef STGCN_SC1B
}

template <class T>
inline bool smallc_supported(int Cin, int Cout, int W, int Kt) {
  return Cin <= 4 && Kt * Cin <= 16 && Cout % 8 == 0 && Cout <= 256 && W <= 512 && (256 % Cout == 0 || Cout == 256);
}

// ---- small elementwise / layout kernels ----------------------------------------------------
// out[i0][i1][i2] (contiguous) (+)= in[off + i0*s0 + i1*s1 + i2*s2]
template <class TO>
__global__ void gather3_kernel(const float* in, TO* out, int d0, int d1, int d2, long long off,
                               long long s0, long long s1, long long s2, int accumulate) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long tot = (long long)d0 * d1 * d2;
  if (idx >= tot) return;
  int i2 = (int)(idx % d2);
  long long q = idx / d2;
  int i1 = (int)(q % d1);
  int i0 = (int)(q / d1);
  float v = in[off + i0 * s0 + i1 * s1 + i2 * s2];
  if (accumulate) v += ldf(out + idx);
  stf(out + idx, v);
}
template <class TO>
inline void launch_gather3(const float* in, TO* out, int d0, int d1, int d2, long long off, long long s0,
                           long long s1, long long s2, int accumulate, cudaStream_t s) {
  long long tot = (long long)d0 * d1 * d2;
  if (tot == 0) return;
  STGCN_LAUNCH(gather3_kernel<TO>, ceil_div(tot, 256), 256, 0, s, in, out, d0, d1, d2, off, s0, s1, s2, accumulate);
}

// Several independent gather3 jobs in ONE launch (the per-call weight re-layouts and gradient scatters are a dozen
// tiny tensors; one launch each was ~8% of the bf16 step).  Jobs must not depend on each other.
struct GatherJob {
  const float* in; void* out; int out_bf16;
  int d0, d1, d2; long long off, s0, s1, s2;
};
constexpr int kMaxGatherJobs = 8;
struct GatherJobs { GatherJob j[kMaxGatherJobs]; int n; };
__global__ void gather3_multi_kernel(GatherJobs jobs) {
  const GatherJob& g = jobs.j[blockIdx.y];
  const long long tot = (long long)g.d0 * g.d1 * g.d2;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < tot; idx += (long long)gridDim.x * blockDim.x) {
    int i2 = (int)(idx % g.d2);
    long long q = idx / g.d2;
    int i1 = (int)(q % g.d1);
    int i0 = (int)(q / g.d1);
    float v = g.in[g.off + i0 * g.s0 + i1 * g.s1 + i2 * g.s2];
    if (g.out_bf16) reinterpret_cast<bf16*>(g.out)[idx] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(g.out)[idx] = v;
  }
}
struct GatherBatch {
  GatherJobs jobs; cudaStream_t stream;
  explicit GatherBatch(cudaStream_t s) : stream(s) { jobs.n = 0; }
  template <class TO>
  void add(const float* in, TO* out, int d0, int d1, int d2, long long off, long long s0, long long s1, long long s2) {
    if ((long long)d0 * d1 * d2 == 0) return;
    if (jobs.n == kMaxGatherJobs) flush();
    jobs.j[jobs.n++] = GatherJob{in, (void*)out, std::is_same<TO, bf16>::value ? 1 : 0, d0, d1, d2, off, s0, s1, s2};
  }
  void flush() {
    if (jobs.n == 0) return;
    long long mx = 0;
    for (int i = 0; i < jobs.n; ++i) mx = std::max(mx, (long long)jobs.j[i].d0 * jobs.j[i].d1 * jobs.j[i].d2);
    int gx = (int)std::min<long long>(ceil_div(mx, 256), 64);
    STGCN_LAUNCH(gather3_multi_kernel, dim3(gx, jobs.n), 256, 0, stream, jobs);
    jobs.n = 0;
  }
  ~GatherBatch() noexcept(false) { flush(); }
};

// out[i*ldo + j] += in[i*si + j*sj]   (strided block accumulate; used to fold 1x1 align weights)
__global__ void add_block_kernel(float* out, int ldo, const float* in, int d0, int d1, long long si, long long sj) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= d0 * d1) return;
  int i = idx / d1, j = idx - i * d1;
  out[(long long)i * ldo + j] += in[i * si + j * sj];
}

// out[r, j] = j < Cin ? in[r*ldi + j] : 0   for j < Cout   (zero-pad or column slice copy)
template <class T>
__global__ void copy_cols_kernel(const T* in, T* out, long long rows, int Cin, int ldi, int Cout, int accumulate) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * Cout) return;
  long long r = idx / Cout;
  int j = (int)(idx - r * Cout);
  float v = j < Cin ? ldf(in + r * ldi + j) : 0.f;
  if (accumulate) v += ldf(out + idx);
  stf(out + idx, v);
}
template <class T>
inline void launch_copy_cols(const T* in, T* out, long long rows, int Cin, int ldi, int Cout, int accumulate,
                             cudaStream_t s) {
  if (rows * Cout == 0) return;
  STGCN_LAUNCH(copy_cols_kernel<T>, ceil_div(rows * Cout, 256), 256, 0, s, in, out, rows, Cin, ldi, Cout, accumulate);
}

// dtype conversion of an activation tensor
template <class TI, class TO>
__global__ void convert_kernel(const TI* in, TO* out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) stf(out + i, ldf(in + i));
}

// y = relu?(g + a)
template <class T>
__global__ void add_relu_kernel(const T* g, const T* a, T* y, long long n, int relu) {
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i >= n) return;
  if (i + 8 <= n && ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    float u[8], w[8];
    load8(g + i, u);
    if (a) { load8(a + i, w);
#pragma unroll
      for (int k = 0; k < 8; ++k) u[k] += w[k]; }
    if (relu) {
#pragma unroll
      for (int k = 0; k < 8; ++k) u[k] = fmaxf(u[k], 0.f); }
    store8(y + i, u);
  } else {
    for (long long e = i + 8 < n ? i + 8 : n; i < e; ++i) {
      float v = ldf(g + i) + (a ? ldf(a + i) : 0.f);
      stf(y + i, relu ? fmaxf(v, 0.f) : v);
    }
  }
}
// dg = relu ? dy * (y > 0) : dy
template <class T>
__global__ void relu_bwd_kernel(const T* dy, const T* y, T* dg, long long n, int relu) {
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i >= n) return;
  if (i + 8 <= n && ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dg)) & 15) == 0) {
    float d[8], v[8];
    load8(dy + i, d); load8(y + i, v);
    if (relu) {
#pragma unroll
      for (int k = 0; k < 8; ++k) d[k] = v[k] > 0.f ? d[k] : 0.f; }
    store8(dg + i, d);
  } else {
    for (long long e = i + 8 < n ? i + 8 : n; i < e; ++i) stf(dg + i, (!relu || ldf(y + i) > 0.f) ? ldf(dy + i) : 0.f);
  }
}
// y += alpha * x
template <class T>
__global__ void axpy_kernel(float alpha, const T* x, T* y, long long n) {
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i >= n) return;
  if (i + 8 <= n && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    float a[8], b[8];
    load8(x + i, a); load8(y + i, b);
#pragma unroll
    for (int k = 0; k < 8; ++k) b[k] += alpha * a[k];
    store8(y + i, b);
  } else {
    for (long long e = i + 8 < n ? i + 8 : n; i < e; ++i) stf(y + i, ldf(y + i) + alpha * ldf(x + i));
  }
}
// y = relu(x), with optional dropout; and its backward
template <class T>
__global__ void relu_dropout_fwd_kernel(const T* x, T* y, long long n, int training, float p, uint64_t seed) {
  seed = live_seed(seed);      // + the device-side step counter, if one is registered (common.cuh)
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = fmaxf(ldf(x + i), 0.f);
  if (training && p > 0.f) v = dropout_keep(seed, (uint64_t)i, p) ? v / (1.f - p) : 0.f;
  stf(y + i, v);
}
template <class T>
__global__ void relu_dropout_bwd_kernel(const T* dy, const T* x, T* dx, long long n, int training, float p,
                                        uint64_t seed) {
  seed = live_seed(seed);      // + the device-side step counter, if one is registered (common.cuh)
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float g = ldf(x + i) > 0.f ? ldf(dy + i) : 0.f;
  if (training && p > 0.f) g = dropout_keep(seed, (uint64_t)i, p) ? g / (1.f - p) : 0.f;
  stf(dx + i, g);
}

// ---- LayerNorm over the joint (N, C) axes of each (b, t) group (layers.py:246,255) ----------
__device__ __forceinline__ float block_sum(float v, float* red) {
  __syncthreads();   // protect red[] reuse
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (l == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

// VEC = 8: 8 elements per thread per step (requires M % 8 == 0 and 16-byte aligned tensors); VEC = 1: scalar.
template <class T, int VEC>
__global__ void __launch_bounds__(512) ln_fwd_kernel(const T* x, const float* w, const float* b, T* y,
                                                     float* mean, float* rstd, int M, float eps, int training,
                                                     float p, uint64_t seed) {
  seed = live_seed(seed);      // + the device-side step counter, if one is registered (common.cuh)
  __shared__ float red[32];
  long long g = blockIdx.x;
  const T* xp = x + g * M;
  float s = 0.f;
  for (int i = threadIdx.x * VEC; i < M; i += blockDim.x * VEC) {
    if (VEC == 8) { float v[8]; load8(xp + i, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) s += v[k]; }
    else s += ldf(xp + i);
  }
  float mu = block_sum(s, red) / M;
  float q = 0.f;
  for (int i = threadIdx.x * VEC; i < M; i += blockDim.x * VEC) {
    if (VEC == 8) { float v[8]; load8(xp + i, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) { float d = v[k] - mu; q += d * d; } }
    else { float d = ldf(xp + i) - mu; q += d * d; }
  }
  float var = block_sum(q, red) / M;
  float rs = rsqrtf(var + eps);
  if (threadIdx.x == 0) { mean[g] = mu; rstd[g] = rs; }
  bool drop = training && p > 0.f;
  float keep_scale = drop ? 1.f / (1.f - p) : 1.f;
  for (int i = threadIdx.x * VEC; i < M; i += blockDim.x * VEC) {
    if (VEC == 8) {
      float v[8], wv[8], bv[8];
      load8(xp + i, v); load8(w + i, wv); load8(b + i, bv);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float o = (v[k] - mu) * rs * wv[k] + bv[k];
        if (drop) o = dropout_keep(seed, (uint64_t)(g * M + i + k), p) ? o * keep_scale : 0.f;
        v[k] = o;
      }
      store8(y + g * M + i, v);
    } else {
      float v = (ldf(xp + i) - mu) * rs * w[i] + b[i];
      if (drop) v = dropout_keep(seed, (uint64_t)(g * M + i), p) ? v * keep_scale : 0.f;
      stf(y + g * M + i, v);
    }
  }
}

// bf16 single-read variant: the group (M <= 512 * 8 * NCH elements) is loaded ONCE into registers (NCH 16-byte chunks
// per thread, all loads in flight together) and the mean / variance / normalise passes run from there -- the generic
// kernel above re-reads the group twice through L1/L2 with a block reduction between the passes (2.4 TB/s measured).
// Same element-to-thread mapping and summation order as ln_fwd_kernel<bf16, 8>: bit-identical results.
template <int NCH>
__global__ void __launch_bounds__(512) ln_fwd_cached_kernel(const bf16* x, const float* w, const float* b, bf16* y,
                                                            float* mean, float* rstd, int M, float eps, int training,
                                                            float p, uint64_t seed) {
  seed = live_seed(seed);      // + the device-side step counter, if one is registered (common.cuh)
  __shared__ float red[32];
  const long long g = blockIdx.x;
  const bf16* xp = x + g * M;
  uint4 xr[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int i = (c * 512 + (int)threadIdx.x) * 8;
    xr[c] = make_uint4(0, 0, 0, 0);
    if (i < M) xr[c] = *reinterpret_cast<const uint4*>(xp + i);
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if ((c * 512 + (int)threadIdx.x) * 8 < M) {
      float v[8]; unpack8(xr[c], v);
#pragma unroll
      for (int k = 0; k < 8; ++k) s += v[k];
    }
  }
  const float mu = block_sum(s, red) / M;
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if ((c * 512 + (int)threadIdx.x) * 8 < M) {
      float v[8]; unpack8(xr[c], v);
#pragma unroll
      for (int k = 0; k < 8; ++k) { const float d = v[k] - mu; q += d * d; }
    }
  }
  const float var = block_sum(q, red) / M;
  const float rs = rsqrtf(var + eps);
  if (threadIdx.x == 0) { mean[g] = mu; rstd[g] = rs; }
  const bool drop = training && p > 0.f;
  const float keep_scale = drop ? 1.f / (1.f - p) : 1.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int i = (c * 512 + (int)threadIdx.x) * 8;
    if (i < M) {
      float v[8], wv[8], bv[8], o[8];
      unpack8(xr[c], v); load8(w + i, wv); load8(b + i, bv);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        o[k] = (v[k] - mu) * rs * wv[k] + bv[k];
        if (drop) o[k] = dropout_keep(seed, (uint64_t)(g * M + i + k), p) ? o[k] * keep_scale : 0.f;
      }
      store8(y + g * M + i, o);
    }
  }
}

template <class T, int VEC>
__global__ void __launch_bounds__(512) ln_bwd_kernel(const T* x, const T* dy, const float* w, const float* mean,
                                                     const float* rstd, T* dx, int M, int training, float p,
                                                     uint64_t seed) {
  seed = live_seed(seed);      // + the device-side step counter, if one is registered (common.cuh)
  __shared__ float red[32];
  long long g = blockIdx.x;
  const T* xp = x + g * M;
  const T* dp = dy + g * M;
  float mu = mean[g], rs = rstd[g];
  bool drop = training && p > 0.f;
  float keep_scale = drop ? 1.f / (1.f - p) : 1.f;
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x * VEC; i < M; i += blockDim.x * VEC) {
    float xv[VEC], dv[VEC], wv[VEC];
    if (VEC == 8) { load8(xp + i, xv); load8(dp + i, dv); load8(w + i, wv); }
    else { xv[0] = ldf(xp + i); dv[0] = ldf(dp + i); wv[0] = w[i]; }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float d = dv[k];
      if (drop) d = dropout_keep(seed, (uint64_t)(g * M + i + k), p) ? d * keep_scale : 0.f;
      float gi = d * wv[k];
      float xh = (xv[k] - mu) * rs;
      s1 += gi; s2 += gi * xh;
    }
  }
  s1 = block_sum(s1, red) / M;
  s2 = block_sum(s2, red) / M;
  for (int i = threadIdx.x * VEC; i < M; i += blockDim.x * VEC) {
    float xv[VEC], dv[VEC], wv[VEC], o[VEC];
    if (VEC == 8) { load8(xp + i, xv); load8(dp + i, dv); load8(w + i, wv); }
    else { xv[0] = ldf(xp + i); dv[0] = ldf(dp + i); wv[0] = w[i]; }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float d = dv[k];
      if (drop) d = dropout_keep(seed, (uint64_t)(g * M + i + k), p) ? d * keep_scale : 0.f;
      float gi = d * wv[k];
      float xh = (xv[k] - mu) * rs;
      o[k] = rs * (gi - s1 - xh * s2);
    }
    if (VEC == 8) store8(dx + g * M + i, o);
    else stf(dx + g * M + i, o[0]);
  }
}

// dw[i] += sum_g dy'[g,i] * xhat[g,i];  db[i] += sum_g dy'[g,i]   (pre-zeroed, atomics over group chunks)
template <class T, int VEC>
__global__ void ln_param_grad_kernel(const T* x, const T* dy, const float* mean, const float* rstd, float* dw,
                                     float* db, int M, long long G, int groups_per_cta, int training, float p,
                                     uint64_t seed) {
  seed = live_seed(seed);      // + the device-side step counter, if one is registered (common.cuh)
  int i = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (i >= M) return;
  long long g0 = (long long)blockIdx.y * groups_per_cta;
  long long g1 = min(G, g0 + groups_per_cta);
  bool drop = training && p > 0.f;
  float keep_scale = drop ? 1.f / (1.f - p) : 1.f;
  float aw[VEC], ab[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) { aw[k] = 0.f; ab[k] = 0.f; }
  for (long long g = g0; g < g1; ++g) {
    float xv[VEC], dv[VEC];
    if (VEC == 8) { load8(x + g * M + i, xv); load8(dy + g * M + i, dv); }
    else { xv[0] = ldf(x + g * M + i); dv[0] = ldf(dy + g * M + i); }
    const float mu = mean[g], rs = rstd[g];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float d = dv[k];
      if (drop) d = dropout_keep(seed, (uint64_t)(g * M + i + k), p) ? d * keep_scale : 0.f;
      aw[k] += d * (xv[k] - mu) * rs;
      ab[k] += d;
    }
  }
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    if (dw) atomicAdd(dw + i + k, aw[k]);
    if (db) atomicAdd(db + i + k, ab[k]);
  }
}

// ---- LayerNorm backward + LayerNorm parameter gradients + gate backward of the producing temporal conv ------------
// (layers.py:255-256 backward through tc2_ln, then layers.py:92-115 backward through the GLU/GTU/... gate.)
// Two launches.  ln_bwd_sums_kernel: the two per-group reductions of the LayerNorm backward (one CTA per (b, t) group,
// read-only).  ln_gate_bwd_kernel: thread = one 8-element chunk (8 channels of one vertex) x a range of groups; for every
// (group, chunk) it forms the LayerNorm data gradient from the group's scalars, sends it straight through the gate
// derivative (P/Q halves of the saved pre-activation z, residual from the conv input) and stores dz -- the 64-channel
// dH never reaches HBM -- while dw/db of its chunk accumulate in registers across the groups (one atomic per element per
// CTA at the end).  Replaces ln_bwd + ln_param_grad + gate_vec: x and dy are read twice instead of three times, dH
// (write + read) disappears.
template <class T>
struct LnGateArgs {
  const T* x; const T* dy; const float* w; const float* mean; const float* rstd;
  float* sums;                          // [G][2] scratch: (sum dy*w, sum dy*w*xhat) / M
  float* dw; float* db;                 // pre-zeroed; may be null
  int M; long long G;
  int training; float p; uint64_t seed;
  // gate
  const T* z; const T* xin; T* dz;
  int N, C, W, Cin, Kt, T_out, T_in, explicit_res;
  int groups_per_cta;
  int q_only;                           // GLU: z holds only Q as [rows, C]; x (the LayerNorm input) is h = u * sigma(Q)
};
__device__ __forceinline__ void block_sum2(float& a, float& b, float* red) {
  __syncthreads();   // protect red[] reuse
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[w] = a; red[32 + w] = b; }
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  if (w == 0) {
    float ta = l < nw ? red[l] : 0.f, tb = l < nw ? red[32 + l] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { ta += __shfl_xor_sync(0xffffffffu, ta, o); tb += __shfl_xor_sync(0xffffffffu, tb, o); }
    if (l == 0) { red[0] = ta; red[32] = tb; }
  }
  __syncthreads();
  a = red[0]; b = red[32];
}
template <class T>
__global__ void __launch_bounds__(256) ln_bwd_sums_kernel(LnGateArgs<T> a) {
  a.seed = live_seed(a.seed);      // + the device-side step counter, if one is registered (common.cuh)
  __shared__ float red[64];
  const long long g = blockIdx.x;
  const T* xp = a.x + g * a.M;
  const T* dp = a.dy + g * a.M;
  const float mu = a.mean[g], rs = a.rstd[g];
  const bool drop = a.training && a.p > 0.f;
  const float keep_scale = drop ? 1.f / (1.f - a.p) : 1.f;
  float s1 = 0.f, s2 = 0.f;
#pragma unroll 4
  for (int i = threadIdx.x * 8; i < a.M; i += 256 * 8) {
    float xv[8], dv[8], wv[8];
    load8(xp + i, xv); load8(dp + i, dv); load8(a.w + i, wv);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float d = dv[k];
      if (drop) d = dropout_keep(a.seed, (uint64_t)(g * a.M + i + k), a.p) ? d * keep_scale : 0.f;
      const float gi = d * wv[k];
      s1 += gi; s2 += gi * (xv[k] - mu) * rs;
    }
  }
  block_sum2(s1, s2, red);
  if (threadIdx.x == 0) { a.sums[2 * g] = s1 / (float)a.M; a.sums[2 * g + 1] = s2 / (float)a.M; }
}
#ifndef STGCN_LNGATE_MINB
#define STGCN_LNGATE_MINB 4
#endif
#ifndef STGCN_LNGATE_UNROLL
#define STGCN_LNGATE_UNROLL 2
#endif
template <class T, int ACT>
__global__ void __launch_bounds__(128, STGCN_LNGATE_MINB) ln_gate_bwd_kernel(LnGateArgs<T> a) {
  a.seed = live_seed(a.seed);      // + the device-side step counter, if one is registered (common.cuh)
  constexpr bool gated = ACT == STGCN_ACT_GLU || ACT == STGCN_ACT_GTU;
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch * 8 >= a.M) return;
  const int i = ch * 8, Cout = a.C;
  const int n = i / Cout, c0 = i - n * Cout;
  const bool has_res = a.explicit_res && c0 < a.Cin;
  const bool drop = a.training && a.p > 0.f;
  const float keep_scale = drop ? 1.f / (1.f - a.p) : 1.f;
  float wv[8], aw[8], ab[8];
  load8(a.w + i, wv);
#pragma unroll
  for (int k = 0; k < 8; ++k) { aw[k] = 0.f; ab[k] = 0.f; }
  const long long g0 = (long long)blockIdx.y * a.groups_per_cta;
  const long long g1 = min(a.G, g0 + a.groups_per_cta);
  constexpr int kUnroll = STGCN_LNGATE_UNROLL;
#pragma unroll kUnroll
  for (long long g = g0; g < g1; ++g) {
    const float mu = a.mean[g], rs = a.rstd[g], s1 = a.sums[2 * g], s2 = a.sums[2 * g + 1];
    const long long r = g * a.N + n;
    float xv[8], dv[8], zp[8], zq[8], res[8], dh[8], du[8], dq[8];
    load8(a.x + g * a.M + i, xv); load8(a.dy + g * a.M + i, dv);
    const bool q_only = ACT == STGCN_ACT_GLU && a.q_only;
    if (q_only) {
      load8(a.z + r * Cout + c0, zq);
    } else {
      load8(a.z + r * a.W + c0, zp);
      if (gated) load8(a.z + r * a.W + Cout + c0, zq);
    }
    if (has_res && !q_only) {
      const long long b = g / a.T_out;
      const int t = (int)(g - b * a.T_out);
      load8(a.xin + ((b * a.T_in + t + a.Kt - 1) * a.N + n) * a.Cin + c0, res);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float d = dv[k];
      if (drop) d = dropout_keep(a.seed, (uint64_t)(g * a.M + i + k), a.p) ? d * keep_scale : 0.f;
      const float xh = (xv[k] - mu) * rs;
      dh[k] = rs * (d * wv[k] - s1 - xh * s2);
      aw[k] += d * xh;
      ab[k] += d;
      if (has_res && !q_only) zp[k] += res[k];
    }
    if (q_only) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float sg = sigmoid_t<kFastAct<T>>(zq[k]);
        du[k] = dh[k] * sg;
        dq[k] = dh[k] * xv[k] * (1.f - sg);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) act_bwd<kFastAct<T>>(ACT, zp[k], gated ? zq[k] : 0.f, dh[k], du[k], dq[k]);
    }
    store8(a.dz + r * a.W + c0, du);
    if (gated) store8(a.dz + r * a.W + Cout + c0, dq);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (a.dw) atomicAdd(a.dw + i + k, aw[k]);
    if (a.db) atomicAdd(a.db + i + k, ab[k]);
  }
}
// One pass over (dY, x) for BOTH reductions of the LayerNorm backward, in front of umma_fb2_kernel (no dropout):
//   * the two sums of every (b, t) group, deterministic: CTA (part, range) covers the columns [part * 8192, +8192) of
//     the groups of its range and writes sums[part][g][2] (already / M) after a block reduction; the consumer adds the
//     parts in a fixed order (atomics would perturb dZ's bf16 rounding from run to run);
//   * the parameter gradients dw / db: thread = two fixed 8-column chunks, accumulated in registers over the CTA's groups,
//     one atomic per element per CTA at the end.
// The next group's operands are requested before the current group is reduced.  (ln_bwd_sums_kernel + a separate
// parameter-gradient pass read the two tensors twice: 26 + 62 us for block 0 at B = 256.)
constexpr int kLnPgThreads = 256, kLnPgCols = kLnPgThreads * 8, kLnPgMaxParts = 16;
constexpr int kLnPgBatch = 4;      // groups per iteration: 4 x 8 KB requested per CTA before anything is consumed
template <class T>
__global__ void __launch_bounds__(kLnPgThreads, 3) ln_bwd_sums_pg_kernel(LnGateArgs<T> a) {
  // (one group per iteration -- 8 KB per CTA in flight, then a block reduction -- was latency bound: 63 us for 120 MB)
  __shared__ float red[kLnPgThreads / 32][2 * kLnPgBatch];
  const int part = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int i0 = part * kLnPgCols + tid * 8;
  const bool act = i0 < a.M;
  float wv[8], aw[8], ab[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { wv[k] = 0.f; aw[k] = 0.f; ab[k] = 0.f; }
  if (act) load8(a.w + i0, wv);
  const long long g0 = (long long)blockIdx.y * a.groups_per_cta;
  const long long g1 = min(a.G, g0 + a.groups_per_cta);
  const float inv_m = 1.f / (float)a.M;
  // operands of the NEXT batch are in flight (cp.async into this thread's own shared-memory slots, two stages) while the
  // current batch is reduced: the first use of directly loaded values was 22 % of this kernel's stall samples
  extern __shared__ uint4 pg_slots[];                     // [2 stages][2 * kLnPgBatch][kLnPgThreads]
  const uint32_t slot0 = (uint32_t)__cvta_generic_to_shared(pg_slots) + (uint32_t)tid * 16u;
  float mu_n[kLnPgBatch], rs_n[kLnPgBatch];
  auto issue = [&](long long gb, int st) {
#pragma unroll
    for (int j = 0; j < kLnPgBatch; ++j) {
      const bool ok = act && gb + j < g1;
      mu_n[j] = 0.f; rs_n[j] = 0.f;
      if (gb + j < g1) { mu_n[j] = a.mean[gb + j]; rs_n[j] = a.rstd[gb + j]; }
      const long long off = ok ? (gb + j) * a.M + i0 : 0;
      const uint32_t d = slot0 + (uint32_t)((st * 2 * kLnPgBatch + 2 * j) * kLnPgThreads) * 16u, nb = ok ? 16u : 0u;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(a.x + off), "r"(nb) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d + kLnPgThreads * 16u), "l"(a.dy + off), "r"(nb) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue(g0, 0);
  int it = 0;
  for (long long gb = g0; gb < g1; gb += kLnPgBatch, ++it) {
    const int st = it & 1;
    float mu[kLnPgBatch], rs[kLnPgBatch];
#pragma unroll
    for (int j = 0; j < kLnPgBatch; ++j) { mu[j] = mu_n[j]; rs[j] = rs_n[j]; }
    const bool more = gb + kLnPgBatch < g1;
    if (more) issue(gb + kLnPgBatch, st ^ 1);
    if (more) asm volatile("cp.async.wait_group 1;" ::: "memory"); else asm volatile("cp.async.wait_group 0;" ::: "memory");
    uint4 xr[kLnPgBatch], dr[kLnPgBatch];
#pragma unroll
    for (int j = 0; j < kLnPgBatch; ++j) {
      const uint32_t d = slot0 + (uint32_t)((st * 2 * kLnPgBatch + 2 * j) * kLnPgThreads) * 16u;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(xr[j].x), "=r"(xr[j].y), "=r"(xr[j].z), "=r"(xr[j].w) : "r"(d) : "memory");
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(dr[j].x), "=r"(dr[j].y), "=r"(dr[j].z), "=r"(dr[j].w) : "r"(d + kLnPgThreads * 16u) : "memory");
    }
    float sv[2 * kLnPgBatch];
#pragma unroll
    for (int j = 0; j < kLnPgBatch; ++j) {
      const uint32_t xw[4] = {xr[j].x, xr[j].y, xr[j].z, xr[j].w}, dw4[4] = {dr[j].x, dr[j].y, dr[j].z, dr[j].w};
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float xv = __uint_as_float((k & 1) ? (xw[k >> 1] & 0xffff0000u) : (xw[k >> 1] << 16));
        const float dv = __uint_as_float((k & 1) ? (dw4[k >> 1] & 0xffff0000u) : (dw4[k >> 1] << 16));
        const float xh = (xv - mu[j]) * rs[j], gi = dv * wv[k];      // inactive threads / groups: dv = 0
        s1 += gi; s2 = fmaf(gi, xh, s2);
        aw[k] = fmaf(dv, xh, aw[k]);
        ab[k] += dv;
      }
      sv[2 * j] = s1; sv[2 * j + 1] = s2;
    }
    // deterministic block reduction of the 2 x kLnPgBatch sums
#pragma unroll
    for (int v = 0; v < 2 * kLnPgBatch; ++v) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sv[v] += __shfl_xor_sync(0xffffffffu, sv[v], o);
    }
    __syncthreads();                                        // red[] of the previous batch has been read
    if (lane == 0) {
#pragma unroll
      for (int v = 0; v < 2 * kLnPgBatch; ++v) red[warp][v] = sv[v];
    }
    __syncthreads();
    if (tid < 2 * kLnPgBatch) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kLnPgThreads / 32; ++w) t += red[w][tid];
      const long long g = gb + (tid >> 1);
      if (g < g1) a.sums[((long long)part * a.G + g) * 2 + (tid & 1)] = t * inv_m;
    }
  }
  if (act) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (a.dw) atomicAdd(a.dw + i0 + k, aw[k]);
      if (a.db) atomicAdd(a.db + i0 + k, ab[k]);
    }
  }
}
inline int ln_pg_parts(int M) { return ceil_div(M, kLnPgCols); }
template <class T>
inline void launch_ln_bwd_sums_pg(LnGateArgs<T> a, int sms, cudaStream_t s) {
  static_assert(std::is_same<T, bf16>::value, "bf16 activations only (the kernel unpacks bf16 pairs)");
  const int parts = ln_pg_parts(a.M);
  int ranges = std::max(1, (3 * sms) / parts);
  if (ranges > a.G) ranges = (int)a.G;
  a.groups_per_cta = ceil_div(a.G, ranges);
  ranges = ceil_div(a.G, a.groups_per_cta);
  const size_t smem = (size_t)2 * 2 * kLnPgBatch * kLnPgThreads * 16;
  STGCN_CUDA(cudaFuncSetAttribute(ln_bwd_sums_pg_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  STGCN_LAUNCH(ln_bwd_sums_pg_kernel<T>, dim3(parts, ranges), kLnPgThreads, smem, s, a);
}

template <class T>
inline bool ln_gate_bwd_supported(const LnGateArgs<T>& a) {
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  return a.G > 0 && a.M % 8 == 0 && a.C % 8 == 0 && a.W % 8 == 0 && a.M == a.N * a.C &&
         (!a.explicit_res || a.Cin % 8 == 0) && al16(a.x) && al16(a.dy) && al16(a.w) && al16(a.z) && al16(a.xin) && al16(a.dz);
}
template <class T>
inline void launch_ln_gate_bwd(int act, LnGateArgs<T> a, int sms, cudaStream_t s) {
  STGCN_LAUNCH(ln_bwd_sums_kernel<T>, (unsigned)a.G, 256, 0, s, a);
  const int xb = ceil_div(a.M / 8, 128);
  int ychunks = (int)std::min<long long>(a.G, std::max<long long>(1, ((long long)sms * 16) / xb));
  a.groups_per_cta = ceil_div(a.G, ychunks);
  ychunks = ceil_div(a.G, a.groups_per_cta);
  const dim3 grid(xb, ychunks);
  switch (act) {
    case STGCN_ACT_GLU: STGCN_LAUNCH((ln_gate_bwd_kernel<T, STGCN_ACT_GLU>), grid, 128, 0, s, a); break;
    case STGCN_ACT_GTU: STGCN_LAUNCH((ln_gate_bwd_kernel<T, STGCN_ACT_GTU>), grid, 128, 0, s, a); break;
    case STGCN_ACT_RELU: STGCN_LAUNCH((ln_gate_bwd_kernel<T, STGCN_ACT_RELU>), grid, 128, 0, s, a); break;
    case STGCN_ACT_SILU: STGCN_LAUNCH((ln_gate_bwd_kernel<T, STGCN_ACT_SILU>), grid, 128, 0, s, a); break;
    default: STGCN_LAUNCH((ln_gate_bwd_kernel<T, STGCN_ACT_LINEAR>), grid, 128, 0, s, a); break;
  }
}

// loss = mean((pred-target)^2); dpred = 2 (pred-target)/n * scale
__global__ void mse_kernel(const float* pred, const float* target, long long n, float scale, float* loss, float* dpred) {
  __shared__ float red[32];
  float s = 0.f;
  float inv = 1.f / (float)n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float d = pred[i] - target[i];
    s += d * d;
    if (dpred) dpred[i] = 2.f * d * inv * scale;
  }
  float t = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(loss, t * inv);
}

}  // namespace simt
}  // namespace stgcn
