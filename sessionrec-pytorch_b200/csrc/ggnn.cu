// K2 / K3: SRGNN's gated graph layer.  Weighted-mean aggregation over the CSR session graph in both
// directions as warp-level segmented reductions (one warp per node, 128-bit row accesses), and the GRUCell
// pointwise gate math.  The dense parts (W1/W2, W_ih, W_hh) go through srk_gemm.
#include "rowops.cuh"

namespace {

template <int NC>
__global__ void __launch_bounds__(256) ggnn_agg_fwd_kernel(const float* __restrict__ X, int N, int d,
                                                           const int* __restrict__ in_ptr, const int* __restrict__ in_src,
                                                           const int* __restrict__ in_eid, const int* __restrict__ out_ptr,
                                                           const int* __restrict__ out_dst, const int* __restrict__ out_eid,
                                                           const float* __restrict__ w, float* __restrict__ NN,
                                                           float* __restrict__ wsum) {
  SRK_PDL();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < N; v += warps) {
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
      const int* ptr = dir ? out_ptr : in_ptr;
      const int* nbr = dir ? out_dst : in_src;
      const int* eid = dir ? out_eid : in_eid;
      RowVec<NC> acc, x;
      row_zero(acc);
      float ws = 0.f;
      for (int s = ptr[v]; s < ptr[v + 1]; ++s) {
        const float we = w[eid[s]];
        row_load(x, X + (long long)nbr[s] * d, d, lane);
        row_axpy(acc, we, x);
        ws += we;
      }
      if (ws > 0.f) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          acc.v[c].x /= ws; acc.v[c].y /= ws; acc.v[c].z /= ws; acc.v[c].w /= ws;
        }
      }
      row_store(acc, NN + (long long)v * 2 * d + dir * d, d, lane);
      if (lane == 0) wsum[2 * v + dir] = ws;
    }
  }
}

template <int NC>
__global__ void __launch_bounds__(256) ggnn_agg_bwd_kernel(const float* __restrict__ dNN, int N, int d,
                                                           const int* __restrict__ in_ptr, const int* __restrict__ in_src,
                                                           const int* __restrict__ in_eid, const int* __restrict__ out_ptr,
                                                           const int* __restrict__ out_dst, const int* __restrict__ out_eid,
                                                           const float* __restrict__ w, const float* __restrict__ wsum,
                                                           float* __restrict__ dX, int accumulate) {
  SRK_PDL();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < N; u += warps) {
    RowVec<NC> acc, g;
    row_zero(acc);
    // x[u] fed the in-mean of every v with an edge u -> v ...
    for (int s = out_ptr[u]; s < out_ptr[u + 1]; ++s) {
      const int v = out_dst[s];
      row_load(g, dNN + (long long)v * 2 * d, d, lane);
      row_axpy(acc, w[out_eid[s]] / wsum[2 * v], g);
    }
    // ... and the out-mean of every t with an edge t -> u
    for (int s = in_ptr[u]; s < in_ptr[u + 1]; ++s) {
      const int t = in_src[s];
      row_load(g, dNN + (long long)t * 2 * d + d, d, lane);
      row_axpy(acc, w[in_eid[s]] / wsum[2 * t + 1], g);
    }
    if (accumulate) row_add_store(acc, dX + (long long)u * d, d, lane);
    else row_store(acc, dX + (long long)u * d, d, lane);
  }
}

__global__ void __launch_bounds__(256) gru_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh,
                                                      const float* __restrict__ h, long long total, int d,
                                                      float* __restrict__ hnew) {
  SRK_PDL();
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    long long i = t / d;
    int j = (int)(t - i * d);
    const float* a = gi + i * 3 * d;
    const float* b = gh + i * 3 * d;
    float r = sigmoidf_(a[j] + b[j]);
    float z = sigmoidf_(a[d + j] + b[d + j]);
    float n = tanhf(a[2 * d + j] + r * b[2 * d + j]);
    hnew[t] = (1.f - z) * n + z * h[t];
  }
}

__global__ void __launch_bounds__(256) gru_bwd_kernel(float* __restrict__ gi, float* __restrict__ gh,
                                                      const float* __restrict__ h, const float* __restrict__ dhnew,
                                                      long long total, int d, float* __restrict__ dh, int accumulate) {
  SRK_PDL();
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    long long i = t / d;
    int j = (int)(t - i * d);
    float* a = gi + i * 3 * d;
    float* b = gh + i * 3 * d;
    const float ghn = b[2 * d + j];
    float r = sigmoidf_(a[j] + b[j]);
    float z = sigmoidf_(a[d + j] + b[d + j]);
    float n = tanhf(a[2 * d + j] + r * ghn);
    const float g = dhnew[t];
    const float dn = g * (1.f - z) * (1.f - n * n);
    const float dz = g * (h[t] - n) * z * (1.f - z);
    const float dr = dn * ghn * r * (1.f - r);
    a[j] = dr; b[j] = dr;
    a[d + j] = dz; b[d + j] = dz;
    a[2 * d + j] = dn; b[2 * d + j] = dn * r;
    const float direct = g * z;
    dh[t] = accumulate ? dh[t] + direct : direct;
  }
}

inline int row_grid(long long rows) {
  long long g = (rows + 7) / 8;
  if (g < 1) g = 1;
  if (g > 148LL * 64) g = 148LL * 64;
  return (int)g;
}
inline int flat_grid(long long n) {
  long long g = (n + 255) / 256;
  if (g < 1) g = 1;
  if (g > 148LL * 16) g = 148LL * 16;
  return (int)g;
}

}  // namespace

extern "C" int srk_ggnn_aggregate_fwd(const float* X, int N, int d, const int* in_ptr, const int* in_src,
                                      const int* in_eid, const int* out_ptr, const int* out_dst, const int* out_eid,
                                      const float* w, float* NN, float* wsum, void* stream) {
  SRK_TRY(srk_check_dim(d));
  if (N <= 0) return SRK_OK;
  SRK_DISPATCH_NC(d, (srk_launch(ggnn_agg_fwd_kernel<NC>, row_grid(N), 256, 0, (cudaStream_t)stream, X, N, d, in_ptr, in_src, in_eid, out_ptr, out_dst, out_eid, w, NN, wsum)));
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_ggnn_aggregate_bwd(const float* dNN, int N, int d, const int* in_ptr, const int* in_src,
                                      const int* in_eid, const int* out_ptr, const int* out_dst, const int* out_eid,
                                      const float* w, const float* wsum, float* dX, int accumulate, void* stream) {
  SRK_TRY(srk_check_dim(d));
  if (N <= 0) return SRK_OK;
  SRK_DISPATCH_NC(d, (srk_launch(ggnn_agg_bwd_kernel<NC>, row_grid(N), 256, 0, (cudaStream_t)stream, dNN, N, d, in_ptr, in_src, in_eid, out_ptr, out_dst, out_eid, w, wsum, dX, accumulate)));
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_gru_pointwise_fwd(const float* gi, const float* gh, const float* h, int N, int d, float* hnew,
                                     void* stream) {
  long long total = (long long)N * d;
  if (total <= 0) return SRK_OK;
  srk_launch(gru_fwd_kernel, flat_grid(total), 256, 0, (cudaStream_t)stream, gi, gh, h, total, d, hnew);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_gru_pointwise_bwd(float* gi, float* gh, const float* h, const float* dhnew, int N, int d, float* dh,
                                     int accumulate, void* stream) {
  long long total = (long long)N * d;
  if (total <= 0) return SRK_OK;
  srk_launch(gru_bwd_kernel, flat_grid(total), 256, 0, (cudaStream_t)stream, gi, gh, h, dhnew, total, d, dh, accumulate);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}
