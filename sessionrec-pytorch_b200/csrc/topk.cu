// Fused evaluation head: per-session top-k item ids straight from the logits, without materialising log-probs
// (`logits.topk(k=cutoff)` in the reference's evaluate(), src/utils/train.py:49; soft-max is monotone, so the ids
// are the same).  One CTA per session row: 3-level radix select (11 + 11 + 10 bits of the order-preserving integer
// image of the float) finds the k-th largest value exactly, one more pass collects the candidates, a bitonic sort in
// shared memory orders them (value descending, index ascending among exact ties).
#include "common.cuh"

namespace {

constexpr int TK_THREADS = 512;
constexpr int TK_CAP = 1024;

__device__ __forceinline__ uint32_t f2key(float x) {
  uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(TK_THREADS) topk_rows_kernel(const float* __restrict__ Z, long long ldz, int V, int k,
                                                               int* __restrict__ out_idx, float* __restrict__ out_val) {
  SRK_PDL();
  __shared__ int hist[2048];
  __shared__ uint32_t cand_u[TK_CAP];
  __shared__ int cand_i[TK_CAP];
  __shared__ uint32_t s_prefix, s_mask;
  __shared__ int s_need, s_count;
  const float* z = Z + (long long)blockIdx.x * ldz;
  const int tid = threadIdx.x;
  if (tid == 0) { s_prefix = 0; s_mask = 0; s_need = k; s_count = 0; }
  const int shifts[3] = {21, 10, 0};
  const int widths[3] = {11, 11, 10};
  for (int lvl = 0; lvl < 3; ++lvl) {
    for (int i = tid; i < 2048; i += TK_THREADS) hist[i] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix, mask = s_mask;
    const int sh = shifts[lvl], nb = 1 << widths[lvl];
    for (int j = tid; j < V; j += TK_THREADS) {
      uint32_t u = f2key(z[j]);
      if ((u & mask) == prefix) atomicAdd(&hist[(u >> sh) & (nb - 1)], 1);
    }
    __syncthreads();
    if (tid == 0) {
      int need = s_need, acc = 0, d = nb - 1;
      for (; d > 0; --d) {
        if (acc + hist[d] >= need) break;
        acc += hist[d];
      }
      s_need = need - acc;                       // how many to take from digit d downwards at the next level
      s_prefix = prefix | ((uint32_t)d << sh);
      s_mask = mask | ((uint32_t)(nb - 1) << sh);
    }
    __syncthreads();
  }
  const uint32_t kth = s_prefix;                 // exact key of the k-th largest value
  for (int j = tid; j < V; j += TK_THREADS) {
    uint32_t u = f2key(z[j]);
    if (u >= kth) {
      int slot = atomicAdd(&s_count, 1);
      if (slot < TK_CAP) { cand_u[slot] = u; cand_i[slot] = j; }
    }
  }
  __syncthreads();
  if (s_count > TK_CAP) {
    // More than TK_CAP - k values tie EXACTLY with the k-th largest (constant / saturated rows): the unordered collection
    // above may have dropped values strictly greater than the k-th.  Exact slow path: every key > kth first (fewer than k
    // of them), then the ties in ascending index order until k candidates are there (what the sort below would pick).
    __syncthreads();
    if (tid == 0) s_count = 0;
    __syncthreads();
    for (int j = tid; j < V; j += TK_THREADS) {
      uint32_t u = f2key(z[j]);
      if (u > kth) {
        int slot = atomicAdd(&s_count, 1);
        cand_u[slot] = u; cand_i[slot] = j;
      }
    }
    __syncthreads();
    int have = s_count;                          // < k
    __syncthreads();
    for (int j0 = 0; j0 < V && have < k; j0 += TK_THREADS) {
      const int j = j0 + tid;
      const bool tie = j < V && f2key(z[j]) == kth;
      const uint32_t bal = __ballot_sync(SRK_FULL, tie);
      if ((tid & 31) == 0) hist[tid >> 5] = __popc(bal);
      __syncthreads();
      int before = 0, total = 0;
      for (int w = 0; w < TK_THREADS / 32; ++w) {
        const int c = hist[w];
        if (w < (tid >> 5)) before += c;
        total += c;
      }
      const int slot = have + before + __popc(bal & ((1u << (tid & 31)) - 1u));
      if (tie && slot < k) { cand_u[slot] = kth; cand_i[slot] = j; }
      have = min(k, have + total);
      __syncthreads();
    }
    if (tid == 0) s_count = have;
    __syncthreads();
  }
  const int n = min(s_count, TK_CAP);
  for (int i = n + tid; i < TK_CAP; i += TK_THREADS) { cand_u[i] = 0u; cand_i[i] = 0x7fffffff; }
  __syncthreads();
  // bitonic sort, descending by key, ascending by index among equal keys
  for (int size = 2; size <= TK_CAP; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < TK_CAP; i += TK_THREADS) {
        int j = i ^ stride;
        if (j > i) {
          bool up = (i & size) == 0;
          uint32_t ui = cand_u[i], uj = cand_u[j];
          int ii = cand_i[i], ij = cand_i[j];
          bool i_before_j = (ui > uj) || (ui == uj && ii < ij);
          if (i_before_j != up) {
            cand_u[i] = uj; cand_u[j] = ui;
            cand_i[i] = ij; cand_i[j] = ii;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < k; i += TK_THREADS) {
    int idx = cand_i[i];
    out_idx[(long long)blockIdx.x * k + i] = idx;
    if (out_val) out_val[(long long)blockIdx.x * k + i] = (idx >= 0 && idx < V) ? z[idx] : 0.f;
  }
}

}  // namespace

extern "C" int srk_topk_rows(const float* Z, long long ldz, int B, int V, int k, int* out_idx, float* out_val, void* stream) {
  if (B <= 0) return SRK_OK;
  SRK_REQUIRE(k >= 1 && k <= 256 && k <= V, "topk: need 1 <= k <= min(256, V), got k=%d V=%d", k, V);
  srk_launch(topk_rows_kernel, B, TK_THREADS, 0, (cudaStream_t)stream, Z, ldz, V, k, out_idx, out_val);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}
