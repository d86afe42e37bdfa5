// Shared device/host helpers for the sm_100a session-rec kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/sessrec_b200.h"

#define SRK_WARP 32
#define SRK_FULL 0xffffffffu

void srk_set_error(const char* fmt, ...);

#define SRK_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      srk_set_error(__VA_ARGS__);              \
      return SRK_ERR_INVALID;                  \
    }                                          \
  } while (0)

#define SRK_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      srk_set_error("%s:%d CUDA error %s (%s)", __FILE__, __LINE__, cudaGetErrorName(e__), \
                    cudaGetErrorString(e__));                                            \
      return SRK_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

extern long long g_srk_launches;   // kernels launched by this library (bench.py reports it as gpu_launches)
// after a raw `kernel<<<...>>>(...)` launch
#define SRK_LAUNCH_CHECK_RAW()        \
  do {                                \
    ++g_srk_launches;                 \
    SRK_CUDA(cudaGetLastError());     \
  } while (0)

// ---- launch layer --------------------------------------------------------------------------------------------------
// Every kernel of the training step is launched through srk_launch().  Normally that is cudaLaunchKernel.  Inside the
// native step (csrc/step.cu) the same call sites can instead (a) be captured, once, into a CUDA graph whose kernel
// nodes are remembered in launch order, or (b) on later steps only UPDATE the parameters of those nodes
// (cudaGraphExecKernelNodeSetParams, ~1 us, against ~2.7 us per launch plus the event traffic of the multi-stream
// fork / join), after which the whole step is ONE cudaGraphLaunch.  Shapes (grid sizes, pointers) may change from step
// to step; the kernel sequence may not - a mismatch makes the step fall back to plain launches.  SKIP drops the launch
// (used to replay only the host-side bookkeeping of a part of the step that has already been enqueued).
enum { SRK_LAUNCH_DIRECT = 0, SRK_LAUNCH_CAPTURE = 1, SRK_LAUNCH_UPDATE = 2, SRK_LAUNCH_SKIP = 3 };
int srk_launch_mode();             // of the calling thread
// blob / blob_bytes: the launch configuration and every argument, byte for byte (only built in capture / update mode): an
// update pass leaves a kernel node alone when its blob equals the one the node was last given
int srk_launch_raw(const void* func, dim3 grid, dim3 block, size_t smem, cudaStream_t st, void** args,
                   const unsigned char* blob = nullptr, size_t blob_bytes = 0);
extern thread_local int g_srk_launch_rc;      // result of the last srk_launch on this thread

#ifdef __CUDACC__
#include <cstring>
#include <tuple>
#include <utility>
constexpr size_t SRK_BLOB_MAX = 2048;
template <typename T>
inline void srk_blob_put(unsigned char* blob, size_t& n, const T& v) {
  if (n + sizeof(T) <= SRK_BLOB_MAX) memcpy(blob + n, &v, sizeof(T));
  n += sizeof(T);
}
template <typename Tuple, size_t... I>
inline int srk_launch_packed(const void* func, dim3 grid, dim3 block, size_t smem, cudaStream_t st, Tuple& pack,
                             std::index_sequence<I...>) {
  void* ptrs[] = {static_cast<void*>(&std::get<I>(pack))...};
  const int mode = srk_launch_mode();
  if (mode == SRK_LAUNCH_CAPTURE || mode == SRK_LAUNCH_UPDATE) {
    unsigned char blob[SRK_BLOB_MAX];
    size_t n = 0;
    srk_blob_put(blob, n, grid);
    srk_blob_put(blob, n, block);
    srk_blob_put(blob, n, smem);
    (void)std::initializer_list<int>{(srk_blob_put(blob, n, std::get<I>(pack)), 0)...};
    // an argument list that does not fit is never treated as unchanged
    return srk_launch_raw(func, grid, block, smem, st, ptrs, n <= SRK_BLOB_MAX ? blob : nullptr, n <= SRK_BLOB_MAX ? n : 0);
  }
  return srk_launch_raw(func, grid, block, smem, st, ptrs);
}
template <typename... KArgs, typename... Args>
inline void srk_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  static_assert(sizeof...(KArgs) == sizeof...(Args), "srk_launch: argument count differs from the kernel's");
  std::tuple<KArgs...> pack(static_cast<KArgs>(args)...);
  g_srk_launch_rc = srk_launch_packed(reinterpret_cast<const void*>(kernel), grid, block, smem, st, pack,
                                      std::index_sequence_for<KArgs...>{});
}
#endif
// after srk_launch(...)
#define SRK_LAUNCH_CHECK()                                   \
  do {                                                       \
    if (g_srk_launch_rc != SRK_OK) return g_srk_launch_rc;   \
  } while (0)

// stream-ordered zero fill / device copy as kernels (graph-capturable as plain kernel nodes; short-lived CTAs)
int srk_zero_async(void* p, size_t bytes, cudaStream_t st);
int srk_zero2_async(void* p0, size_t bytes0, void* p1, size_t bytes1, cudaStream_t st);      // two regions, one launch
int srk_copy_async(void* dst, const void* src, size_t bytes, cudaStream_t st);
int srk_zero2d_async(float* C, long long ldc, int rows, int cols, cudaStream_t st);

#define SRK_TRY(expr)               \
  do {                              \
    int r__ = (expr);               \
    if (r__ != SRK_OK) return r__;  \
  } while (0)

// Programmatic dependent launch.  Every kernel of this library starts with SRK_PDL(): griddepcontrol.wait blocks until the
// kernels this launch depends on have completed and flushed their writes (a no-op for a launch without the attribute);
// griddepcontrol.launch_dependents then lets the NEXT kernel of the stream be scheduled while this one runs, so that its
// launch latency and prologue overlap this kernel instead of following it - the training steps are chains of 40-70 dependent
// kernels of a few microseconds each.  The wait comes first and is executed by every thread before anything else, so a
// kernel never touches memory before its producers are done, and "kernel N + 1 complete" still implies "kernel N complete".
// srk_launch() sets cudaLaunchAttributeProgrammaticStreamSerialization (SESSREC_PDL=0 turns it off).
#ifdef __CUDACC__
#define SRK_PDL()                                             \
  do {                                                        \
    asm volatile("griddepcontrol.wait;" ::: "memory");        \
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); \
  } while (0)
#endif

static inline int srk_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SRK_FULL, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(SRK_FULL, v, o));
  return v;
}

// Counter-based dropout: 24-bit uniform from a splitmix64 finaliser of (seed, site, flat index).
// Restated bit-for-bit in oracle/models.py::counter_uniform24 so that tests can inject the same masks.
__device__ __forceinline__ uint32_t srk_rand24(uint64_t seed, uint32_t site, uint64_t idx) {
  uint64_t x = seed + 0x9E3779B97F4A7C15ull * (((uint64_t)site << 40) + idx + 1ull);
  x ^= x >> 30;
  x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27;
  x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return (uint32_t)(x >> 40);
}

struct DropCfg {
  uint64_t seed;
  uint32_t site;
  uint32_t thresh;  // keep iff rand24 >= thresh; 0 disables dropout
  float scale;      // 1 / (1 - p)
};

static inline DropCfg make_drop(const srk_dropout* d, uint32_t site_offset = 0) {
  DropCfg c;
  c.seed = d ? d->seed : 0;
  c.site = (d ? d->site : 0) + site_offset;
  float p = d ? d->p : 0.f;
  c.thresh = (p > 0.f) ? (uint32_t)floor((double)p * 16777216.0) : 0u;
  c.scale = (p > 0.f) ? (float)(1.0 / (1.0 - (double)p)) : 1.f;
  return c;
}

__device__ __forceinline__ float drop_mul(const DropCfg& c, uint64_t idx) {
  if (c.thresh == 0u) return 1.f;
  return srk_rand24(c.seed, c.site, idx) >= c.thresh ? c.scale : 0.f;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// internal launchers shared between translation units -------------------------------------------------
struct GemmArgs {
  int M, N, K;
  const float* A; long long sa_m, sa_k;   // element (m, k) at A[row(m or k) ...]; see gemm.cu
  const float* B; long long sb_k, sb_n;
  float* C; long long ldc;
  const int* a_idx;   // optional indirection on A's M index
  const int* b_idx;   // optional indirection on B's K index (only for sb_n == 1 layouts)
  const int* c_idx;   // optional indirection on C's row index
  const float* bias;  // optional [N], added when the split index is 0
  float alpha;
  int accumulate;     // 0: C = result, 1: C += result
  int split_k;        // >= 1; > 1 requires accumulate = 1 (atomicAdd epilogue)
};
int srk_gemm_launch(const GemmArgs& g, cudaStream_t st);
int srk_pick_split_k(int M, int N, int K);
