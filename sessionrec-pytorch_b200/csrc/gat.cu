// K4: 8-head graph attention of MSGIFSR's MSHGNN layer over the CSR session graphs.
//
// Forward (gat_agg_fwd): one CTA per destination node, one warp per head.  Each warp runs the edge softmax of
// its head over the node's in-edges for every relation instance (conv1 on the graph + conv2 on the reverse
// graph, all relations that end in this node type), accumulates attention-weighted source projections with
// 128-bit row loads, adds residual + bias per instance, then the CTA takes the max over heads, adds the
// session mean and (optionally) L2-normalises - all without leaving the SM.
// Backward is split by who owns the output: per destination node (softmax/LeakyReLU backward -> dedge, der),
// per source node (dZel gather over out-edges; no atomics, deterministic), plus small reductions.
#include <float.h>

#include "rowops.cuh"

namespace {

constexpr int H = SRK_HEADS;

struct GatParams {
  srk_gat_inst inst[SRK_MAX_GAT_INST];
  int n_inst;
};

__device__ __forceinline__ float leaky(float x) { return x > 0.f ? x : 0.2f * x; }

template <int NC>
__global__ void __launch_bounds__(256) gat_agg_fwd_kernel(const GatParams P, int N, int d, const float* __restrict__ segmean,
                                                          const int* __restrict__ node2seg, DropCfg adrop, int normalize,
                                                          float* __restrict__ Hout, float* __restrict__ rnorm,
                                                          uint8_t* __restrict__ amax) {
  SRK_PDL();
  extern __shared__ float smem[];           // O[H][d]
  __shared__ float red[8];
  const int v = blockIdx.x;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ldz = H * d + H;
  RowVec<NC> acc;
  row_zero(acc);
  for (int q = 0; q < P.n_inst; ++q) {
    const srk_gat_inst& I = P.inst[q];
    const int s0 = I.in_ptr[v], s1 = I.in_ptr[v + 1];
    const float erv = I.er[(long long)v * H + h];
    float m = -FLT_MAX, ssum = 0.f;
    for (int s = s0; s < s1; ++s) {
      float x = leaky(I.Zel[(long long)I.in_src[s] * ldz + H * d + h] + erv);
      float mn = fmaxf(m, x);
      ssum = ssum * expf(m - mn) + expf(x - mn);
      m = mn;
    }
    DropCfg dc = adrop;
    dc.site = I.attn_site;
    for (int s = s0; s < s1; ++s) {
      const int u = I.in_src[s];
      const long long ei = (long long)I.in_eid[s] * H + h;
      float a = expf(leaky(I.Zel[(long long)u * ldz + H * d + h] + erv) - m) / ssum;
      if (lane == 0) I.att[ei] = a;
      a *= drop_mul(dc, (uint64_t)ei);
      RowVec<NC> z;
      row_load(z, I.Zel + (long long)u * ldz + h * d, d, lane);
      row_axpy(acc, a, z);
    }
    RowVec<NC> r;
    row_load(r, I.xdst + (long long)v * d, d, lane);
    row_axpy(acc, 1.f, r);
    row_load(r, I.bias + h * d, d, lane);
    row_axpy(acc, 1.f, r);
  }
  row_store(acc, smem + h * d, d, lane);
  __syncthreads();
  const float* mean = segmean + (long long)node2seg[v] * d;
  float sq = 0.f;
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    float best = 0.f;
    int bh = 0;
    if (P.n_inst > 0) {
      best = smem[j];
#pragma unroll
      for (int k = 1; k < H; ++k) {
        float o = smem[k * d + j];
        if (o > best) { best = o; bh = k; }
      }
    }
    float val = best + mean[j];
    amax[(long long)v * d + j] = (uint8_t)bh;
    smem[j] = val;                       // row 0 now holds Hpre (each j touched by exactly one thread)
    sq += val * val;
  }
  sq = warp_sum(sq);
  if (lane == 0) red[h] = sq;
  __syncthreads();
  float n = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) n += red[k];
  n = sqrtf(n);
  const float inv = normalize ? 1.f / fmaxf(n, 1e-12f) : 1.f;
  for (int j = threadIdx.x; j < d; j += blockDim.x) Hout[(long long)v * d + j] = smem[j] * inv;
  if (threadIdx.x == 0 && rnorm) rnorm[v] = n;
}

// dO of head h at node v: dHpre[v, j] where head h won the max, else 0.
template <int NC>
__device__ __forceinline__ void load_dO(RowVec<NC>& o, const float* __restrict__ dHpre_row,
                                        const uint8_t* __restrict__ amax_row, int h, int d, int lane) {
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    int col = (c * 32 + lane) * 4;
    if (col < d) {
      float4 g = *reinterpret_cast<const float4*>(dHpre_row + col);
      uchar4 a = *reinterpret_cast<const uchar4*>(amax_row + col);
      o.v[c] = make_float4(a.x == h ? g.x : 0.f, a.y == h ? g.y : 0.f, a.z == h ? g.z : 0.f, a.w == h ? g.w : 0.f);
    } else {
      o.v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

template <int NC>
__global__ void __launch_bounds__(256) gat_bwd_dst_kernel(const GatParams P, int N, int d, DropCfg adrop, int normalize,
                                                          const float* __restrict__ Hn, const float* __restrict__ rnorm,
                                                          const uint8_t* __restrict__ amax, const float* __restrict__ dH,
                                                          float* __restrict__ dHpre) {
  SRK_PDL();
  const int v = blockIdx.x;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ldz = H * d + H;
  // normalisation backward (every warp recomputes the row; d is small)
  RowVec<NC> g;
  row_load(g, dH + (long long)v * d, d, lane);
  if (normalize) {
    RowVec<NC> y;
    row_load(y, Hn + (long long)v * d, d, lane);
    const float n = rnorm[v];
    if (n > 1e-12f) {
      float t = row_dot(y, g);
      row_axpy(g, -t, y);
      row_scale(g, 1.f / n);
    } else {
      row_scale(g, 1e12f);
    }
  }
  if (h == 0) row_store(g, dHpre + (long long)v * d, d, lane);
  // keep only the columns whose max came from this head
  RowVec<NC> dO;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    int col = (c * 32 + lane) * 4;
    if (col < d) {
      uchar4 a = *reinterpret_cast<const uchar4*>(amax + (long long)v * d + col);
      dO.v[c] = make_float4(a.x == h ? g.v[c].x : 0.f, a.y == h ? g.v[c].y : 0.f, a.z == h ? g.v[c].z : 0.f,
                            a.w == h ? g.v[c].w : 0.f);
    } else {
      dO.v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  for (int q = 0; q < P.n_inst; ++q) {
    const srk_gat_inst& I = P.inst[q];
    const int s0 = I.in_ptr[v], s1 = I.in_ptr[v + 1];
    DropCfg dc = adrop;
    dc.site = I.attn_site;
    float t2 = 0.f;
    for (int s = s0; s < s1; ++s) {
      const long long ei = (long long)I.in_eid[s] * H + h;
      RowVec<NC> z;
      row_load(z, I.Zel + (long long)I.in_src[s] * ldz + h * d, d, lane);
      t2 += I.att[ei] * drop_mul(dc, (uint64_t)ei) * row_dot(z, dO);
    }
    const float erv = I.er[(long long)v * H + h];
    float dsum = 0.f;
    for (int s = s0; s < s1; ++s) {
      const int u = I.in_src[s];
      const long long ei = (long long)I.in_eid[s] * H + h;
      RowVec<NC> z;
      row_load(z, I.Zel + (long long)u * ldz + h * d, d, lane);
      const float a = I.att[ei];
      float de = a * (drop_mul(dc, (uint64_t)ei) * row_dot(z, dO) - t2);
      const float pre = I.Zel[(long long)u * ldz + H * d + h] + erv;
      de *= pre > 0.f ? 1.f : 0.2f;
      if (lane == 0) I.dedge[ei] = de;
      dsum += de;
    }
    if (lane == 0) I.der[(long long)v * H + h] = dsum;
  }
}

template <int NC>
__global__ void __launch_bounds__(256) gat_bwd_src_kernel(const srk_gat_inst I, int d, DropCfg dc,
                                                          const float* __restrict__ dHpre,
                                                          const uint8_t* __restrict__ amax, float* __restrict__ Zhi,
                                                          float* __restrict__ Zlo) {
  SRK_PDL();
  const int u = blockIdx.x;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ldz = H * d + H;
  RowVec<NC> acc;
  row_zero(acc);
  float del = 0.f;
  for (int s = I.out_ptr[u]; s < I.out_ptr[u + 1]; ++s) {
    const int v = I.out_dst[s];
    const long long ei = (long long)I.out_eid[s] * H + h;
    const float a = I.att[ei] * drop_mul(dc, (uint64_t)ei);
    RowVec<NC> dO;
    load_dO(dO, dHpre + (long long)v * d, amax + (long long)v * d, h, d, lane);
    row_axpy(acc, a, dO);
    del += I.dedge[ei];
  }
  row_store(acc, I.dZel + (long long)u * ldz + h * d, d, lane);
  if (lane == 0) I.dZel[(long long)u * ldz + H * d + h] = del;
  if (Zhi) {                 // TF32 hi / lo copy for the tensor-core GEMMs that consume dZel (saves a separate split pass)
    RowVec<NC> hi, lo;
    row_split_tf32(acc, hi, lo);
    row_store(hi, Zhi + (long long)u * ldz + h * d, d, lane);
    row_store(lo, Zlo + (long long)u * ldz + h * d, d, lane);
    if (lane == 0) {
      const float dh = __uint_as_float(__float_as_uint(del) & 0xFFFFE000u);
      Zhi[(long long)u * ldz + H * d + h] = dh;
      Zlo[(long long)u * ldz + H * d + h] = del - dh;
    }
  }
}

__global__ void __launch_bounds__(256) gat_bias_bwd_kernel(const float* __restrict__ dHpre, const uint8_t* __restrict__ amax,
                                                           int N, int d, int rows_per_block, float* __restrict__ dbias) {
  SRK_PDL();
  __shared__ float red[8][H][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(N, r0 + rows_per_block);
  float bins[H];
#pragma unroll
  for (int k = 0; k < H; ++k) bins[k] = 0.f;
  if (j < d)
    for (int r = r0 + ty; r < r1; r += 8) {
      const float g = dHpre[(long long)r * d + j];
      const int a = amax[(long long)r * d + j];
#pragma unroll
      for (int k = 0; k < H; ++k) bins[k] += (a == k) ? g : 0.f;
    }
#pragma unroll
  for (int k = 0; k < H; ++k) red[ty][k][tx] = bins[k];
  __syncthreads();
  // 8 warps: warp ty finalises head ty
  if (j < d) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][ty][tx];
    atomicAdd(dbias + ty * d + j, s);
  }
}

template <int NC>
__global__ void __launch_bounds__(256) segmean_fwd_kernel(const float* __restrict__ X, const int* __restrict__ seg, int B,
                                                          int d, float* __restrict__ mean) {
  SRK_PDL();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < B; b += warps) {
    RowVec<NC> acc, x;
    row_zero(acc);
    const int s0 = seg[b], s1 = seg[b + 1];
    for (int i = s0; i < s1; ++i) {
      row_load(x, X + (long long)i * d, d, lane);
      row_axpy(acc, 1.f, x);
    }
    const float cnt = (float)max(s1 - s0, 1);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      acc.v[c].x /= cnt; acc.v[c].y /= cnt; acc.v[c].z /= cnt; acc.v[c].w /= cnt;
    }
    row_store(acc, mean + (long long)b * d, d, lane);
  }
}

template <int NC>
__global__ void __launch_bounds__(256) segmean_bwd_kernel(const float* __restrict__ dHpre, const int* __restrict__ seg,
                                                          int B, int d, float* __restrict__ dX, int accumulate) {
  SRK_PDL();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < B; b += warps) {
    RowVec<NC> acc, x;
    row_zero(acc);
    const int s0 = seg[b], s1 = seg[b + 1];
    for (int i = s0; i < s1; ++i) {
      row_load(x, dHpre + (long long)i * d, d, lane);
      row_axpy(acc, 1.f, x);
    }
    const float cnt = (float)max(s1 - s0, 1);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      acc.v[c].x /= cnt; acc.v[c].y /= cnt; acc.v[c].z /= cnt; acc.v[c].w /= cnt;
    }
    for (int i = s0; i < s1; ++i) {
      if (accumulate) row_add_store(acc, dX + (long long)i * d, d, lane);
      else row_store(acc, dX + (long long)i * d, d, lane);
    }
  }
}

// W_aug = [W ; wl] (rows 8d..8d+7 = a_l contracted with W) and wr[8, d], optionally together with the TF32 hi / lo split of
// W_aug the tensor-core projection reads: CTA per (head, 32-column tile); 8 row groups x 32 columns, smem reduce.  One launch
// instead of copy + contraction + split.
__device__ __forceinline__ void tf32_split_store(float x, float* hi, float* lo, long long at) {
  const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  hi[at] = h;
  lo[at] = x - h;
}
__global__ void __launch_bounds__(256) gat_prep_kernel(const float* __restrict__ W, const float* __restrict__ al,
                                                       const float* __restrict__ ar, int d, float* __restrict__ Waug,
                                                       float* __restrict__ wr, float* __restrict__ Whi,
                                                       float* __restrict__ Wlo) {
  SRK_PDL();
  __shared__ float red[2][8][33];
  const int h = blockIdx.y;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + tx;
  float sl = 0.f, sr = 0.f;
  if (i < d)
    for (int j = ty; j < d; j += 8) {
      const long long at = (long long)(h * d + j) * d + i;
      const float w = W[at];
      Waug[at] = w;
      if (Whi) tf32_split_store(w, Whi, Wlo, at);
      sl = fmaf(al[h * d + j], w, sl);
      sr = fmaf(ar[h * d + j], w, sr);
    }
  red[0][ty][tx] = sl;
  red[1][ty][tx] = sr;
  __syncthreads();
  if (ty == 0 && i < d) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      sl += red[0][k][tx];
      sr += red[1][k][tx];
    }
    const long long at = (long long)(H * d + h) * d + i;
    Waug[at] = sl;
    if (Whi) tf32_split_store(sl, Whi, Wlo, at);
    wr[h * d + i] = sr;
  }
}

template <int NC>
__global__ void __launch_bounds__(256) gat_prep_bwd_kernel(const float* __restrict__ W, const float* __restrict__ al,
                                                           const float* __restrict__ ar, const float* __restrict__ dWaug,
                                                           const float* __restrict__ dwr, int d, float* __restrict__ dW,
                                                           float* __restrict__ dal, float* __restrict__ dar) {
  SRK_PDL();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < H * d; r += warps) {
    const int h = r / d;
    RowVec<NC> w, gl, gr, g;
    row_load(w, W + (long long)r * d, d, lane);
    row_load(gl, dWaug + (long long)(H * d + h) * d, d, lane);
    row_load(gr, dwr + (long long)h * d, d, lane);
    row_load(g, dWaug + (long long)r * d, d, lane);
    const float tl = row_dot(w, gl), tr = row_dot(w, gr);
    if (lane == 0) {
      dal[r] += tl;
      dar[r] += tr;
    }
    row_axpy(g, al[r], gl);
    row_axpy(g, ar[r], gr);
    row_add_store(g, dW + (long long)r * d, d, lane);
  }
}

inline int row_grid(long long rows) {
  long long g = (rows + 7) / 8;
  if (g < 1) g = 1;
  if (g > 148LL * 64) g = 148LL * 64;
  return (int)g;
}

int fill_params(GatParams& P, const srk_gat_inst* inst_host, int n_inst) {
  SRK_REQUIRE(n_inst >= 0 && n_inst <= SRK_MAX_GAT_INST, "gat: %d instances (max %d)", n_inst, SRK_MAX_GAT_INST);
  memset(&P, 0, sizeof(P));
  P.n_inst = n_inst;
  for (int i = 0; i < n_inst; ++i) P.inst[i] = inst_host[i];
  return SRK_OK;
}

}  // namespace

extern "C" int srk_gat_prep_split(const float* W, const float* attn_l, const float* attn_r, int d, float* Waug, float* wr,
                                  float* Whi, float* Wlo, void* stream) {
  SRK_TRY(srk_check_dim(d));
  SRK_REQUIRE((Whi == nullptr) == (Wlo == nullptr), "gat_prep: pass both halves of the TF32 split or neither");
  srk_launch(gat_prep_kernel, dim3(srk_cdiv(d, 32), H), 256, 0, (cudaStream_t)stream, W, attn_l, attn_r, d, Waug, wr, Whi, Wlo);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_gat_prep(const float* W, const float* attn_l, const float* attn_r, int d, float* Waug, float* wr,
                            void* stream) {
  return srk_gat_prep_split(W, attn_l, attn_r, d, Waug, wr, nullptr, nullptr, stream);
}

extern "C" int srk_gat_prep_bwd(const float* W, const float* attn_l, const float* attn_r, const float* dWaug,
                                const float* dwr, int d, float* dW, float* dattn_l, float* dattn_r, void* stream) {
  SRK_TRY(srk_check_dim(d));
  SRK_DISPATCH_NC(d, (srk_launch(gat_prep_bwd_kernel<NC>, row_grid((long long)H * d), 256, 0, (cudaStream_t)stream, W, attn_l, attn_r, dWaug, dwr, d, dW, dattn_l, dattn_r)));
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_gat_aggregate_fwd(const srk_gat_inst* inst_host, int n_inst, int N, int d, const float* segmean,
                                     const int* node2seg, const srk_dropout* attn_drop, int normalize, float* Hout,
                                     float* rnorm, uint8_t* amax, void* stream) {
  SRK_TRY(srk_check_dim(d));
  if (N <= 0) return SRK_OK;
  GatParams P;
  SRK_TRY(fill_params(P, inst_host, n_inst));
  DropCfg dc = make_drop(attn_drop);
  size_t smem = sizeof(float) * (size_t)H * d;
  SRK_DISPATCH_NC(d, (srk_launch(gat_agg_fwd_kernel<NC>, N, 256, smem, (cudaStream_t)stream, P, N, d, segmean, node2seg, dc,
                                                                                     normalize, Hout, rnorm, amax)));
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_gat_aggregate_bwd_dst(const srk_gat_inst* inst_host, int n_inst, int N, int d,
                                         const srk_dropout* attn_drop, int normalize, const float* Hn, const float* rnorm,
                                         const uint8_t* amax, const float* dH, float* dHpre, void* stream) {
  SRK_TRY(srk_check_dim(d));
  if (N <= 0) return SRK_OK;
  GatParams P;
  SRK_TRY(fill_params(P, inst_host, n_inst));
  DropCfg dc = make_drop(attn_drop);
  SRK_DISPATCH_NC(d, (srk_launch(gat_bwd_dst_kernel<NC>, N, 256, 0, (cudaStream_t)stream, P, N, d, dc, normalize, Hn, rnorm, amax,
                                                                                  dH, dHpre)));
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_gat_aggregate_bwd_src(const srk_gat_inst* inst_host, int d, const srk_dropout* attn_drop,
                                         const float* dHpre, const uint8_t* amax, void* stream) {
  SRK_TRY(srk_check_dim(d));
  const srk_gat_inst I = *inst_host;
  if (I.n_src <= 0) return SRK_OK;
  DropCfg dc = make_drop(attn_drop);
  dc.site = I.attn_site;
  SRK_DISPATCH_NC(d, (srk_launch(gat_bwd_src_kernel<NC>, I.n_src, 256, 0, (cudaStream_t)stream, I, d, dc, dHpre, amax, nullptr,
                                 nullptr)));
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_gat_aggregate_bwd_src_split(const srk_gat_inst* inst_host, int d, const srk_dropout* attn_drop,
                                               const float* dHpre, const uint8_t* amax, float* dZel_hi, float* dZel_lo,
                                               void* stream) {
  SRK_TRY(srk_check_dim(d));
  SRK_REQUIRE(dZel_hi != nullptr && dZel_lo != nullptr, "gat_aggregate_bwd_src_split: the hi / lo pair is required");
  const srk_gat_inst& I = *inst_host;
  if (I.n_src <= 0) return SRK_OK;
  DropCfg dc = make_drop(attn_drop);
  dc.site = I.attn_site;
  SRK_DISPATCH_NC(d, (srk_launch(gat_bwd_src_kernel<NC>, I.n_src, 256, 0, (cudaStream_t)stream, I, d, dc, dHpre, amax, dZel_hi,
                                 dZel_lo)));
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_gat_bias_bwd(const float* dHpre, const uint8_t* amax, int N, int d, float* dbias, void* stream) {
  if (N <= 0 || d <= 0) return SRK_OK;
  int by = srk_cdiv(N, 256);
  if (by > 64) by = 64;
  int rows_per_block = srk_cdiv(N, by);
  dim3 grid(srk_cdiv(d, 32), by);
  srk_launch(gat_bias_bwd_kernel, grid, 256, 0, (cudaStream_t)stream, dHpre, amax, N, d, rows_per_block, dbias);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_segmean_fwd(const float* X, const int* seg, int B, int d, float* mean, void* stream) {
  SRK_TRY(srk_check_dim(d));
  if (B <= 0) return SRK_OK;
  SRK_DISPATCH_NC(d, (srk_launch(segmean_fwd_kernel<NC>, row_grid(B), 256, 0, (cudaStream_t)stream, X, seg, B, d, mean)));
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_segmean_bwd(const float* dHpre, const int* seg, int B, int d, float* dX, int accumulate, void* stream) {
  SRK_TRY(srk_check_dim(d));
  if (B <= 0) return SRK_OK;
  SRK_DISPATCH_NC(d, (srk_launch(segmean_bwd_kernel<NC>, row_grid(B), 256, 0, (cudaStream_t)stream, dHpre, seg, B, d, dX,
                                                                                            accumulate)));
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}
