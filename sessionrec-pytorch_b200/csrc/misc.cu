// Error channel + small elementwise / reduction helpers of the path.
#include <stdarg.h>

#include "rowops.cuh"

static thread_local char g_err[512] = "";

void srk_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* srk_last_error(void) { return g_err; }
extern "C" int srk_version(void) { return 100; }
long long g_srk_launches = 0;
extern "C" long long srk_launch_count(void) { return g_srk_launches; }

namespace {

__global__ void dropout_apply_kernel(const float* __restrict__ X, float* __restrict__ Y, long long n, DropCfg dc,
                                     int accumulate) {
  SRK_PDL();
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = X[i] * drop_mul(dc, (uint64_t)i);
    Y[i] = accumulate ? Y[i] + v : v;
  }
}

// Y = dropout(X + A): the residual add of the GAT destination-copy gradient folded into the mask pass
__global__ void dropout_add_kernel(const float* __restrict__ X, const float* __restrict__ A, float* __restrict__ Y, long long n,
                                   DropCfg dc) {
  SRK_PDL();
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    Y[i] = (X[i] + A[i]) * drop_mul(dc, (uint64_t)i);
}

// Y = dropout(X) together with its TF32 hi / lo split (operands of the tensor-core projection that consumes Y)
__global__ void dropout_split_kernel(const float* __restrict__ X, float* __restrict__ Y, float* __restrict__ Yhi,
                                     float* __restrict__ Ylo, long long n, DropCfg dc) {
  SRK_PDL();
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = X[i] * drop_mul(dc, (uint64_t)i);
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    Y[i] = v;
    Yhi[i] = h;
    Ylo[i] = v - h;
  }
}

__global__ void fill_kernel(float* __restrict__ X, long long n, float value) {
  SRK_PDL();
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) X[i] = value;
}

template <int NC>
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ X, const int* __restrict__ idx, int R,
                                                          int d, float* __restrict__ Y, long long ldy) {
  SRK_PDL();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < R; r += warps) {
    RowVec<NC> x;
    row_load(x, X + (long long)idx[r] * d, d, lane);
    row_store(x, Y + r * ldy, d, lane);
  }
}

// Y[idx[r]] += X[r]; idx entries must be distinct (one "last" node per session).
template <int NC>
__global__ void __launch_bounds__(256) scatter_add_rows_kernel(const float* __restrict__ X, long long ldx,
                                                               const int* __restrict__ idx, int R, int d,
                                                               float* __restrict__ Y) {
  SRK_PDL();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < R; r += warps) {
    RowVec<NC> x;
    row_load(x, X + r * ldx, d, lane);
    row_add_store(x, Y + (long long)idx[r] * d, d, lane);
  }
}

__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, long long ldx, int R, int d,
                                                     int rows_per_block, float* __restrict__ out) {
  SRK_PDL();
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(R, r0 + rows_per_block);
  float s = 0.f;
  if (j < d)
    for (int r = r0 + ty; r < r1; r += 8) s += X[r * ldx + j];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && j < d) {
#pragma unroll
    for (int k = 1; k < 8; ++k) s += red[k][tx];
    atomicAdd(out + j, s);
  }
}

__global__ void __launch_bounds__(1024) mean_kernel(const float* __restrict__ x, int n, float* __restrict__ out) {
  SRK_PDL();
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) out[0] = s / (float)n;
  }
}

__device__ __forceinline__ void adam_update(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                            float* __restrict__ v, long long i, float decay, float lr, float b1, float b2,
                                            float eps, float bc1, float bc2_sqrt, float grad_scale) {
  float grad = g[i] * grad_scale + decay * p[i];
  float mi = b1 * m[i] + (1.f - b1) * grad;
  float vi = b2 * v[i] + (1.f - b2) * grad * grad;
  m[i] = mi;
  v[i] = vi;
  float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = p[i] - (lr / bc1) * (mi / denom);
}

// Fused multi-tensor Adam with torch.optim.Adam semantics (L2 term folded into the gradient, bias correction,
// denom = sqrt(v) / sqrt(bias2) + eps) over one flat parameter buffer; per-segment weight decay.
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, long long n,
                                                   const long long* __restrict__ seg_off, const float* __restrict__ seg_decay,
                                                   int n_seg, float lr, float b1, float b2, float eps, float bc1,
                                                   float bc2_sqrt, float grad_scale) {
  SRK_PDL();
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int lo = 0, hi = n_seg - 1;           // segment of element i (seg_off ascending, seg_off[n_seg] == n)
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (seg_off[mid] <= i) lo = mid; else hi = mid - 1;
    }
    const float dec = seg_decay[lo];
    if (dec < 0.f) continue;              // parameter the reference's forward never reaches (grad is None): Adam skips it
    adam_update(p, g, m, v, i, dec, lr, b1, b2, eps, bc1, bc2_sqrt, grad_scale);
  }
}

__device__ __forceinline__ bool in_sorted(const int* __restrict__ a, int n, int x) {
  int lo = 0, hi = n - 1;
  while (lo <= hi) {
    int mid = (lo + hi) >> 1;
    int y = a[mid];
    if (y == x) return true;
    if (y < x) lo = mid + 1; else hi = mid - 1;
  }
  return false;
}

// The same element-wise update split in two launches around one [rows, d] table that starts at element `tab`:
//   part 0: the table rows NOT listed in rows_sorted[n_rows] (their gradient is final as soon as the catalog backward is,
//           so this launch runs beside the encoder backward);
//   part 1: everything else = the elements outside the table + the listed rows (the rows the batch gathered, whose
//           gradient the scatter-add completes last).  Every element is updated exactly once, by the same arithmetic.
__global__ void __launch_bounds__(256) adam_split_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                         float* __restrict__ m, float* __restrict__ v, long long n,
                                                         const long long* __restrict__ seg_off,
                                                         const float* __restrict__ seg_decay, int n_seg, long long tab,
                                                         int rows, int d, long long tab_span,
                                                         const int* __restrict__ rows_sorted, int n_rows, int part, float lr,
                                                         float b1, float b2, float eps, float bc1, float bc2_sqrt,
                                                         float grad_scale) {
  SRK_PDL();
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (part == 0) {
    // one float4 per thread and NO grid-stride loop: this launch runs beside latency-critical kernels on higher-priority
    // streams, and the block scheduler can only hand an SM slot to them when one of these CTAs retires
    int lo = 0, hi = n_seg - 1;           // the table is one segment
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (seg_off[mid] <= tab) lo = mid; else hi = mid - 1;
    }
    const float decay = seg_decay[lo];
    const long long t4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long t = t4 * 4;
    if (t >= (long long)rows * d) return;
    if (in_sorted(rows_sorted, n_rows, (int)(t / d))) return;
    const long long i = tab + t;
    float4 pp = *reinterpret_cast<const float4*>(p + i), gg = *reinterpret_cast<const float4*>(g + i);
    float4 mm = *reinterpret_cast<const float4*>(m + i), vv = *reinterpret_cast<const float4*>(v + i);
    const float step_size = lr / bc1;
#define SRK_ADAM1(c)                                                   \
    {                                                                  \
      const float grad = gg.c * grad_scale + decay * pp.c;             \
      mm.c = b1 * mm.c + (1.f - b1) * grad;                            \
      vv.c = b2 * vv.c + (1.f - b2) * grad * grad;                     \
      pp.c = pp.c - step_size * (mm.c / (sqrtf(vv.c) / bc2_sqrt + eps)); \
    }
    SRK_ADAM1(x) SRK_ADAM1(y) SRK_ADAM1(z) SRK_ADAM1(w)
#undef SRK_ADAM1
    *reinterpret_cast<float4*>(m + i) = mm;
    *reinterpret_cast<float4*>(v + i) = vv;
    *reinterpret_cast<float4*>(p + i) = pp;
  } else {
    const long long listed = (long long)n_rows * d;
    const long long total = n - tab_span + listed;       // virtual index space: [0, tab) | listed rows | [tab + tab_span, n)
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
      long long i;
      if (t < tab) i = t;
      else if (t < tab + listed) {
        const long long q = t - tab;
        i = tab + (long long)rows_sorted[q / d] * d + q % d;
      } else i = t - listed + tab_span;
      if (i >= tab + (long long)rows * d && i < tab + tab_span) continue;       // alignment padding behind the table
      int lo = 0, hi = n_seg - 1;
      while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (seg_off[mid] <= i) lo = mid; else hi = mid - 1;
      }
      const float dec = seg_decay[lo];
      if (dec < 0.f) continue;            // inactive segment (see adam_kernel)
      adam_update(p, g, m, v, i, dec, lr, b1, b2, eps, bc1, bc2_sqrt, grad_scale);
    }
  }
}

inline int flat_grid(long long n, int threads) {
  long long g = (n + threads - 1) / threads;
  if (g < 1) g = 1;
  if (g > 148LL * 16) g = 148LL * 16;
  return (int)g;
}
inline int row_grid(long long rows) {
  long long g = (rows + 7) / 8;
  if (g < 1) g = 1;
  if (g > 148LL * 64) g = 148LL * 64;
  return (int)g;
}

}  // namespace

extern "C" int srk_dropout_apply(const float* X, float* Y, long long n, const srk_dropout* drop, int accumulate,
                                 void* stream) {
  if (n <= 0) return SRK_OK;
  srk_launch(dropout_apply_kernel, flat_grid(n, 256), 256, 0, (cudaStream_t)stream, X, Y, n, make_drop(drop), accumulate);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_dropout_apply_add(const float* X, const float* A, float* Y, long long n, const srk_dropout* drop, void* stream) {
  if (n <= 0) return SRK_OK;
  srk_launch(dropout_add_kernel, flat_grid(n, 256), 256, 0, (cudaStream_t)stream, X, A, Y, n, make_drop(drop));
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_dropout_apply_split(const float* X, float* Y, float* Yhi, float* Ylo, long long n, const srk_dropout* drop,
                                       void* stream) {
  if (n <= 0) return SRK_OK;
  SRK_REQUIRE(Y != nullptr && Yhi != nullptr && Ylo != nullptr, "dropout_apply_split: all three outputs are required");
  srk_launch(dropout_split_kernel, flat_grid(n, 256), 256, 0, (cudaStream_t)stream, X, Y, Yhi, Ylo, n, make_drop(drop));
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_fill(float* X, long long n, float value, void* stream) {
  if (n <= 0) return SRK_OK;
  srk_launch(fill_kernel, flat_grid(n, 256), 256, 0, (cudaStream_t)stream, X, n, value);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_gather_rows(const float* X, const int* idx, int R, int d, float* Y, long long ldy, void* stream) {
  SRK_TRY(srk_check_dim(d));
  if (R <= 0) return SRK_OK;
  SRK_REQUIRE(ldy % 4 == 0, "gather_rows: ldy must be a multiple of 4");
  SRK_DISPATCH_NC(d, (srk_launch(gather_rows_kernel<NC>, row_grid(R), 256, 0, (cudaStream_t)stream, X, idx, R, d, Y, ldy)));
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_scatter_add_rows(const float* X, long long ldx, const int* idx, int R, int d, float* Y, void* stream) {
  SRK_TRY(srk_check_dim(d));
  if (R <= 0) return SRK_OK;
  SRK_REQUIRE(ldx % 4 == 0, "scatter_add_rows: ldx must be a multiple of 4");
  SRK_DISPATCH_NC(d, (srk_launch(scatter_add_rows_kernel<NC>, row_grid(R), 256, 0, (cudaStream_t)stream, X, ldx, idx, R, d, Y)));
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_colsum(const float* X, long long ldx, int R, int d, float* out, int accumulate, void* stream) {
  if (d <= 0) return SRK_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) SRK_TRY(srk_zero_async(out, sizeof(float) * d, st));
  if (R <= 0) return SRK_OK;
  int by = srk_cdiv(R, 256);
  if (by > 64) by = 64;
  int rows_per_block = srk_cdiv(R, by);
  dim3 grid(srk_cdiv(d, 32), by);
  srk_launch(colsum_kernel, grid, 256, 0, st, X, ldx, R, d, rows_per_block, out);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_mean(const float* x, int n, float* out, void* stream) {
  SRK_REQUIRE(n > 0, "mean: n must be positive");
  srk_launch(mean_kernel, 1, 1024, 0, (cudaStream_t)stream, x, n, out);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_adam_step_split(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                                   const long long* seg_off, const float* seg_decay, int n_seg, long long tab, int rows,
                                   int d, long long tab_span, const int* rows_sorted, int n_rows, int part, float lr,
                                   float beta1, float beta2, float eps, int step, float grad_scale, void* stream) {
  if (n <= 0) return SRK_OK;
  SRK_REQUIRE(n_seg >= 1 && step >= 1 && (part == 0 || part == 1), "adam_split: bad arguments");
  SRK_REQUIRE(tab >= 0 && rows >= 0 && d > 0 && tab_span >= (long long)rows * d && tab + tab_span <= n && n_rows >= 0,
              "adam_split: table [%lld, +%lld) outside the flat buffer", tab, tab_span);
  float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  const long long work = part == 0 ? (long long)rows * d : n - tab_span + (long long)n_rows * d;
  if (work <= 0) return SRK_OK;
  SRK_REQUIRE(part == 1 || (d % 4 == 0 && tab % 4 == 0), "adam_split: table rows must be 16-byte aligned");
  const int grid = part == 0 ? srk_cdiv(work / 4, 256) : flat_grid(work, 256);
  srk_launch(adam_split_kernel, grid, 256, 0, (cudaStream_t)stream, param, grad, exp_avg, exp_avg_sq, n, seg_off,
                                                                            seg_decay, n_seg, tab, rows, d, tab_span,
                                                                            rows_sorted, n_rows, part, lr, beta1, beta2, eps,
                                                                            bc1, bc2_sqrt, grad_scale);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                             const long long* seg_off, const float* seg_decay, int n_seg, float lr, float beta1,
                             float beta2, float eps, int step, float grad_scale, void* stream) {
  if (n <= 0) return SRK_OK;
  SRK_REQUIRE(n_seg >= 1 && step >= 1, "adam: need >= 1 segment and step >= 1");
  float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  srk_launch(adam_kernel, flat_grid(n, 256), 256, 0, (cudaStream_t)stream, param, grad, exp_avg, exp_avg_sq, n,
                                                                   seg_off, seg_decay,
                                                                   n_seg, lr, beta1, beta2, eps, bc1, bc2_sqrt, grad_scale);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}
