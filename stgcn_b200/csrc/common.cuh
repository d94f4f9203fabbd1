// common.cuh -- error handling, launch accounting, workspace carving, small device helpers.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <atomic>
#include <cstdlib>
#include <utility>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include <mutex>

#include "../../include/stgcn_b200.h"

namespace stgcn {

// ---- errors: C++ exceptions inside, int status at the C boundary -------------------------
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

extern thread_local char g_last_error[512];
extern std::atomic<uint64_t> g_launches;

inline void set_error(const char* msg) {
  std::snprintf(g_last_error, sizeof(g_last_error), "%s", msg);
}

#define STGCN_CHECK(cond, code, msg)                                       \
  do {                                                                     \
    if (!(cond)) throw ::stgcn::Error((code), std::string(msg) + " [" #cond "]"); \
  } while (0)

#define STGCN_CUDA(expr)                                                   \
  do {                                                                     \
    cudaError_t _e = (expr);                                               \
    if (_e != cudaSuccess)                                                 \
      throw ::stgcn::Error((int)_e, std::string(#expr ": ") + cudaGetErrorString(_e)); \
  } while (0)

// ---- optional per-launch CUDA-event profiler (stgcn_profile_begin/end) ----------------------
// When enabled every launch is bracketed by two events recorded on the launching stream; the
// records are aggregated per "<op tag>:<kernel>" key when the profile is collected.
struct ProfRec { std::string key; cudaEvent_t a, b; };
struct Profiler {
  std::atomic<bool> on{false};
  std::mutex mu;
  std::vector<ProfRec> recs;
};
extern Profiler g_prof;
extern thread_local const char* g_tag;     // current op label (set by the host-side op code)
// STGCN_PREC_TF32X3: the fp32 chain's GEMM launchers (simt_kernels.cuh: launch_tapgemm / launch_gso / launch_wgrad) hand their
// work to the 3xTF32 tcgen05 kernel (umma_x3.cuh) for the duration of the ABI call that set this
extern thread_local bool g_x3;
struct X3Scope {
  bool prev;
  explicit X3Scope(bool on) : prev(g_x3) { g_x3 = on; }
  ~X3Scope() { g_x3 = prev; }
};

struct Tag {                               // RAII label for the launches of one logical op
  const char* prev;
  explicit Tag(const char* t) : prev(g_tag) { g_tag = t; }
  ~Tag() { g_tag = prev; }
};

inline cudaEvent_t prof_begin(cudaStream_t s) {
  cudaEvent_t a;
  cudaEventCreate(&a);
  cudaEventRecord(a, s);
  return a;
}
inline void prof_end(const char* kernel, cudaEvent_t a, cudaStream_t s) {
  cudaEvent_t b;
  cudaEventCreate(&b);
  cudaEventRecord(b, s);
  std::string key = std::string(g_tag ? g_tag : "-") + ":" + kernel;
  std::lock_guard<std::mutex> lk(g_prof.mu);
  g_prof.recs.push_back(ProfRec{key, a, b});
}

template <class... KArgs, class... Args>
inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  kernel<<<grid, block, smem, stream>>>(std::forward<Args>(args)...);
}

// Every kernel launch goes through this: counts it and checks the launch status.
#define STGCN_LAUNCH(kernel, grid, block, smem, stream, ...)               \
  STGCN_LAUNCH_NAMED(#kernel, kernel, grid, block, smem, stream, __VA_ARGS__)

#define STGCN_LAUNCH_NAMED(name, kernel, grid, block, smem, stream, ...)   \
  do {                                                                     \
    bool _prof = ::stgcn::g_prof.on.load(std::memory_order_relaxed);       \
    cudaEvent_t _ea = nullptr;                                             \
    if (_prof) _ea = ::stgcn::prof_begin(stream);                          \
    ::stgcn::launch_kernel(kernel, dim3(grid), dim3(block), (size_t)(smem), (stream), __VA_ARGS__); \
    if (_prof) ::stgcn::prof_end(name, _ea, stream);                       \
    ::stgcn::g_launches.fetch_add(1, std::memory_order_relaxed);           \
    STGCN_CUDA(cudaPeekAtLastError());                                     \
  } while (0)

template <class F>
int guarded(F&& f) {
  try {
    f();
    return STGCN_OK;
  } catch (const Error& e) {
    set_error(e.what());
    return e.code ? e.code : STGCN_E_INVALID;
  } catch (const std::exception& e) {
    set_error(e.what());
    return STGCN_E_INVALID;
  } catch (...) {
    set_error("unknown error");
    return STGCN_E_INVALID;
  }
}

// ---- bump allocator over a caller-provided buffer ---------------------------------------
struct Arena {
  char* base;
  size_t cap, off, peak = 0;
  bool dry;   // dry run: only measure
  Arena(void* p, size_t bytes) : base((char*)p), cap(bytes), off(0), dry(p == nullptr) {}
  static size_t align_up(size_t v) { return (v + 255) & ~size_t(255); }
  template <class T>
  T* take(size_t n) {
    size_t o = off;
    off = align_up(off + n * sizeof(T));
    if (off > peak) peak = off;
    if (dry) return nullptr;
    if (off > cap) throw Error(STGCN_E_WORKSPACE, "workspace/saved buffer too small");
    return reinterpret_cast<T*>(base + o);
  }
};

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- device helpers ----------------------------------------------------------------------
// Branch-free MUFU math.  (__frcp_rn / IEEE division put a slow-path CALL inside BSSY/BSYNC regions around every
// element, which stops the compiler from interleaving elements: a serialized ~150-cycle chain per sigmoid.)
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoidf_(float v) { return rcp_approx(1.f + ex2_approx(-1.4426950408889634f * v)); }
// bf16 storage mode only: sigma(v) = 0.5 + 0.5 tanh(v / 2) is ONE MUFU op (tanh.approx, |error| < 2^-11, far below the
// bf16 the results are stored in) instead of ex2 + rcp; the gate epilogues are MUFU / issue bound.  The fp32 parity
// path keeps sigmoidf_ (cancellation near sigma -> 0 would eat into its 1e-3 budget).
__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_tanh_(float v) { return fmaf(0.5f, tanh_approx(0.5f * v), 0.5f); }
template <bool FAST>
__device__ __forceinline__ float sigmoid_t(float v) { return FAST ? sigmoid_tanh_(v) : sigmoidf_(v); }
template <bool FAST>
__device__ __forceinline__ float tanh_t(float v) { return FAST ? tanh_approx(v) : tanhf(v); }

// Dropout seeds cross the ABI by value, which a CUDA-graph capture freezes: every replay would draw the mask of the
// capture pass.  A caller that replays graphs registers a device-side 64-bit step counter (stgcn_set_dropout_step) and
// bumps it once per step from inside the graph; every kernel that draws a mask adds it to its by-value seed, so the
// forward and the backward of one step agree and successive replays differ.
__device__ const unsigned long long* g_dropout_step = nullptr;
__device__ __forceinline__ uint64_t live_seed(uint64_t seed) {
  const unsigned long long* p = g_dropout_step;
  return p ? seed + 0x632BE59BD9B4E019ull * (*p + 1) : seed;
}

// Counter-based dropout keep-mask: splitmix64 finaliser over (seed, element index).
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t idx, float p_drop) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  float u = (float)(uint32_t)(z >> 40) * (1.0f / 16777216.0f);   // 24 random bits -> [0,1)
  return u >= p_drop;
}

}  // namespace stgcn
