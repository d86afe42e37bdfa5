// K5 attention readout (warp per session, online segment softmax) and the row-wise parts of the scoring
// head: log-sum-exp / NLL / log-prob rewrite and their backward.  HBM-bound streaming kernels.
#include <float.h>

#include "rowops.cuh"

namespace {

template <int NC>
__device__ __forceinline__ float attn_score(const RowVec<NC>& u, const RowVec<NC>& v, const RowVec<NC>& we) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    s += we.v[c].x * sigmoidf_(u.v[c].x + v.v[c].x) + we.v[c].y * sigmoidf_(u.v[c].y + v.v[c].y) +
         we.v[c].z * sigmoidf_(u.v[c].z + v.v[c].z) + we.v[c].w * sigmoidf_(u.v[c].w + v.v[c].w);
  }
  return warp_sum(s);
}

template <int NC>
__global__ void __launch_bounds__(256) readout_fwd_kernel(const float* __restrict__ F, const float* __restrict__ u,
                                                          const float* __restrict__ v, const float* __restrict__ we,
                                                          const int* __restrict__ seg, const int* __restrict__ last, int B,
                                                          int d, int with_last, float* __restrict__ e,
                                                          float* __restrict__ ms, float* __restrict__ sr_in) {
  SRK_PDL();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  RowVec<NC> wev;
  row_load(wev, we, d, lane);
  for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < B; b += warps) {
    RowVec<NC> vb, acc;
    row_load(vb, v + (long long)b * d, d, lane);
    row_zero(acc);
    float m = -FLT_MAX, s = 0.f;
    for (int i = seg[b]; i < seg[b + 1]; ++i) {
      RowVec<NC> ui, fi;
      row_load(ui, u + (long long)i * d, d, lane);
      row_load(fi, F + (long long)i * d, d, lane);
      float ei = attn_score(ui, vb, wev);
      if (lane == 0) e[i] = ei;
      float mn = fmaxf(m, ei);
      float corr = expf(m - mn), pi = expf(ei - mn);
      s = s * corr + pi;
      row_scale(acc, corr);
      row_axpy(acc, pi, fi);
      m = mn;
    }
    row_scale(acc, 1.f / s);
    if (lane == 0) {
      ms[2 * b] = m;
      ms[2 * b + 1] = s;
    }
    if (with_last) {
      RowVec<NC> fl;
      row_load(fl, F + (long long)last[b] * d, d, lane);
      row_store(fl, sr_in + (long long)b * 2 * d, d, lane);
    }
    row_store(acc, sr_in + (long long)b * 2 * d + d, d, lane);
  }
}

template <int NC>
__global__ void __launch_bounds__(256) readout_bwd_kernel(const float* __restrict__ F, float* __restrict__ u,
                                                          float* __restrict__ v, const float* __restrict__ we,
                                                          const int* __restrict__ seg, const int* __restrict__ last,
                                                          const float* __restrict__ e, const float* __restrict__ ms,
                                                          const float* __restrict__ sr_in,
                                                          const float* __restrict__ dsr_in, int B, int d,
                                                          int with_last, float* __restrict__ dF,
                                                          float* __restrict__ dwe, float* __restrict__ duh,
                                                          float* __restrict__ dul) {
  SRK_PDL();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  RowVec<NC> wev, dwe_acc;
  row_load(wev, we, d, lane);
  row_zero(dwe_acc);
  for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < B; b += warps) {
    RowVec<NC> vb, g, dg, dl, dv;
    row_load(vb, v + (long long)b * d, d, lane);
    row_load(g, sr_in + (long long)b * 2 * d + d, d, lane);
    row_load(dl, dsr_in + (long long)b * 2 * d, d, lane);
    row_load(dg, dsr_in + (long long)b * 2 * d + d, d, lane);
    row_zero(dv);
    const float t = row_dot(g, dg);
    const float m = ms[2 * b], inv_s = 1.f / ms[2 * b + 1];
    const int lb = with_last ? last[b] : -1;
    for (int i = seg[b]; i < seg[b + 1]; ++i) {
      RowVec<NC> ui, fi, du, df;
      row_load(ui, u + (long long)i * d, d, lane);
      row_load(fi, F + (long long)i * d, d, lane);
      const float alpha = expf(e[i] - m) * inv_s;
      const float de = alpha * (row_dot(fi, dg) - t);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        float sx = sigmoidf_(ui.v[c].x + vb.v[c].x), sy = sigmoidf_(ui.v[c].y + vb.v[c].y);
        float sz = sigmoidf_(ui.v[c].z + vb.v[c].z), sw = sigmoidf_(ui.v[c].w + vb.v[c].w);
        dwe_acc.v[c].x = fmaf(de, sx, dwe_acc.v[c].x); dwe_acc.v[c].y = fmaf(de, sy, dwe_acc.v[c].y);
        dwe_acc.v[c].z = fmaf(de, sz, dwe_acc.v[c].z); dwe_acc.v[c].w = fmaf(de, sw, dwe_acc.v[c].w);
        du.v[c].x = de * wev.v[c].x * sx * (1.f - sx); du.v[c].y = de * wev.v[c].y * sy * (1.f - sy);
        du.v[c].z = de * wev.v[c].z * sz * (1.f - sz); du.v[c].w = de * wev.v[c].w * sw * (1.f - sw);
      }
      row_axpy(dv, 1.f, du);
      row_store(du, u + (long long)i * d, d, lane);
      if (duh) row_store_tf32_split(du, duh + (long long)i * d, dul + (long long)i * d, d, lane);
      df = dg;
      row_scale(df, alpha);
      if (i == lb) row_axpy(df, 1.f, dl);
      row_store(df, dF + (long long)i * d, d, lane);
    }
    row_store(dv, v + (long long)b * d, d, lane);
  }
  // d w_e: one set of atomics per CTA, not per warp.  With a warp per session B = 2048 warps would each fire d atomicAdds at
  // the same d addresses (measured: 98 us of serialised L2 atomics at cfg2); the 8 warps of a CTA are summed in shared memory
  // first and the grid is capped, so an address sees at most a few hundred adds.
  __shared__ float red[8][NC * 128];
  const int wib = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int col = (c * 32 + lane) * 4;
    red[wib][col] = dwe_acc.v[c].x; red[wib][col + 1] = dwe_acc.v[c].y;
    red[wib][col + 2] = dwe_acc.v[c].z; red[wib][col + 3] = dwe_acc.v[c].w;
  }
  __syncthreads();
  for (int col = threadIdx.x; col < d; col += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][col];
    atomicAdd(dwe + col, t);
  }
}

// ---- scoring head row kernels ------------------------------------------------------------------------

struct MS {
  float m, s;
};
__device__ __forceinline__ MS ms_combine(MS a, MS b) {
  MS r;
  r.m = fmaxf(a.m, b.m);
  r.s = a.s * expf(a.m - r.m) + b.s * expf(b.m - r.m);
  return r;
}

// (max, sum of exp) of the whole block; `red` holds one entry per warp and must not be reused before a barrier
__device__ __forceinline__ MS block_ms_reduce(MS t, MS* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MS other;
    other.m = __shfl_xor_sync(SRK_FULL, t.m, o);
    other.s = __shfl_xor_sync(SRK_FULL, t.s, o);
    t = ms_combine(t, other);
  }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  MS r = red[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) r = ms_combine(r, red[w]);
  return r;
}

__device__ __forceinline__ MS block_lse(const float* z, int V, MS* red) {
  MS t;
  t.m = -FLT_MAX;
  t.s = 0.f;
  // rows are 16-byte aligned (ldz % 4 == 0): 128-bit loads, one max-rescale per 4 elements
  const int V4 = ((reinterpret_cast<uintptr_t>(z) & 15u) == 0) ? (V >> 2) : 0;
  const float4* z4 = reinterpret_cast<const float4*>(z);
  for (int j = threadIdx.x; j < V4; j += blockDim.x) {
    float4 x = z4[j];
    float mx = fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w));
    if (mx > t.m) {
      t.s *= expf(t.m - mx);
      t.m = mx;
    }
    t.s += expf(x.x - t.m) + expf(x.y - t.m) + expf(x.z - t.m) + expf(x.w - t.m);
  }
  for (int j = V4 * 4 + threadIdx.x; j < V; j += blockDim.x) {
    float x = z[j];
    if (x > t.m) {
      t.s = t.s * expf(t.m - x) + 1.f;
      t.m = x;
    } else {
      t.s += expf(x - t.m);
    }
  }
  return block_ms_reduce(t, red);
}

__global__ void __launch_bounds__(512) ce_rows_fwd_kernel(float* __restrict__ Z, long long ldz, const int* __restrict__ labels,
                                                          int V, int write_logp, float* __restrict__ lse,
                                                          float* __restrict__ nll) {
  SRK_PDL();
  __shared__ MS red[16];
  float* z = Z + (long long)blockIdx.x * ldz;
  MS r = block_lse(z, V, red);
  const float l = r.m + logf(r.s);
  if (threadIdx.x == 0) {
    lse[blockIdx.x] = l;
    if (labels) {                       // label < 0: the label's column lives on another rank (catalog sharding)
      const int lab = labels[blockIdx.x];
      nll[blockIdx.x] = lab >= 0 ? l - z[lab] : 0.f;
    }
  }
  if (write_logp) {
    __syncthreads();   // thread 0 has read z[label] before anyone rewrites it
    const int V4 = ((reinterpret_cast<uintptr_t>(z) & 15u) == 0) ? (V >> 2) : 0;
    float4* z4 = reinterpret_cast<float4*>(z);
    for (int j = threadIdx.x; j < V4; j += blockDim.x) {
      float4 x = z4[j];
      z4[j] = make_float4(x.x - l, x.y - l, x.z - l, x.w - l);
    }
    for (int j = V4 * 4 + threadIdx.x; j < V; j += blockDim.x) z[j] -= l;
  }
}

__global__ void __launch_bounds__(512) ce_rows_bwd_kernel(float* __restrict__ Z, long long ldz, const int* __restrict__ labels,
                                                          const float* __restrict__ lse, const float* __restrict__ gscale,
                                                          float scale, int B, int V, int z_is_logp,
                                                          float* __restrict__ Zlo, int col0) {
  SRK_PDL();
  // columns [col0, col0 + V) of every row (col0 = 0, V = catalog size for the whole matrix)
  float* z = Z + (long long)blockIdx.x * ldz + col0;
  float* zl = Zlo ? Zlo + (long long)blockIdx.x * ldz + col0 : nullptr;
  const float l = z_is_logp ? 0.f : lse[blockIdx.x];
  const float c = gscale[0] * scale / (float)B;
  const int lab = labels[blockIdx.x] - col0;
  const bool al = (reinterpret_cast<uintptr_t>(z) & 15u) == 0 && (!zl || (reinterpret_cast<uintptr_t>(zl) & 15u) == 0);
  const int V4 = al ? (V >> 2) : 0;
  float4* z4 = reinterpret_cast<float4*>(z);
  float4* zl4 = reinterpret_cast<float4*>(zl);
  for (int j = threadIdx.x; j < V4; j += blockDim.x) {
    float4 x = z4[j];
    const int b = j * 4;
    float4 g = make_float4(c * (expf(x.x - l) - (b == lab ? 1.f : 0.f)), c * (expf(x.y - l) - (b + 1 == lab ? 1.f : 0.f)),
                           c * (expf(x.z - l) - (b + 2 == lab ? 1.f : 0.f)), c * (expf(x.w - l) - (b + 3 == lab ? 1.f : 0.f)));
    if (zl) {                                   // TF32 split for the tcgen05 backward GEMMs
      float4 h = make_float4(__uint_as_float(__float_as_uint(g.x) & 0xFFFFE000u), __uint_as_float(__float_as_uint(g.y) & 0xFFFFE000u),
                             __uint_as_float(__float_as_uint(g.z) & 0xFFFFE000u), __uint_as_float(__float_as_uint(g.w) & 0xFFFFE000u));
      z4[j] = h;
      zl4[j] = make_float4(g.x - h.x, g.y - h.y, g.z - h.z, g.w - h.w);
    } else {
      z4[j] = g;
    }
  }
  for (int j = V4 * 4 + threadIdx.x; j < V; j += blockDim.x) {
    float p = expf(z[j] - l);
    float g = c * (p - (j == lab ? 1.f : 0.f));
    if (zl) {
      float h = __uint_as_float(__float_as_uint(g) & 0xFFFFE000u);
      z[j] = h;
      zl[j] = g - h;
    } else {
      z[j] = g;
    }
  }
}

__global__ void __launch_bounds__(512) logp_bwd_kernel(const float* __restrict__ LP, long long ldlp,
                                                       const float* __restrict__ G, long long ldg, float scale, int V,
                                                       float* __restrict__ DZ, long long lddz, float* __restrict__ DZlo) {
  SRK_PDL();
  __shared__ float red[16];
  __shared__ float total;
  const float* lp = LP + (long long)blockIdx.x * ldlp;
  const float* g = G + (long long)blockIdx.x * ldg;
  float* dz = DZ + (long long)blockIdx.x * lddz;
  float* dzl = DZlo ? DZlo + (long long)blockIdx.x * lddz : nullptr;
  float s = 0.f;
  for (int j = threadIdx.x; j < V; j += blockDim.x) s += g[j];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
    total = t;
  }
  __syncthreads();
  const float rs = total;
  for (int j = threadIdx.x; j < V; j += blockDim.x) {
    float v = scale * (g[j] - expf(lp[j]) * rs);
    if (dzl) {
      float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
      dz[j] = h;
      dzl[j] = v - h;
    } else {
      dz[j] = v;
    }
  }
}

// ---- REnorm head of MSGIFSR (msgifsr.py:281-305) ----------------------------------------------------------------------
// score[b, v] = phi[b, 0] * softmax over the session's own items (v in the session) + phi[b, 1] * softmax over all other
// items; log score[b, v] = z[b, v] + log phi[b, g] - lse_g[b] with g the group of v.  The session's items are exactly its
// order-1 nodes (iid[seg[b] .. seg[b + 1]), unique within the session), so the mask never exists as a matrix.

__device__ __forceinline__ void store_split(float* dz, float* dzl, long long j, float v) {
  if (dzl) {                                    // TF32 hi / lo pair for the tcgen05 backward GEMMs
    float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    dz[j] = h;
    dzl[j] = v - h;
  } else {
    dz[j] = v;
  }
}

__global__ void __launch_bounds__(512) renorm_head_fwd_kernel(float* Z, long long ldz, int V, const int* __restrict__ iid,
                                                              const int* __restrict__ seg, const float* __restrict__ lphi,
                                                              float* __restrict__ zin) {
  SRK_PDL();
  __shared__ MS red_in[16];
  __shared__ MS red_ex[16];
  float* z = Z + (long long)blockIdx.x * ldz;
  const int s0 = seg[blockIdx.x], n = seg[blockIdx.x + 1] - s0;
  MS t;
  t.m = -FLT_MAX;
  t.s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float x = z[iid[s0 + i]];
    zin[s0 + i] = x;
    MS o;
    o.m = x;
    o.s = 1.f;
    t = ms_combine(t, o);
  }
  const MS rin = block_ms_reduce(t, red_in);           // barrier inside: every in-set logit has been read
  for (int i = threadIdx.x; i < n; i += blockDim.x) z[iid[s0 + i]] = -INFINITY;
  __syncthreads();
  const MS rex = block_lse(z, V, red_ex);
  const float c_in = lphi[2 * blockIdx.x] - (rin.m + logf(rin.s));
  const float c_ex = rex.s > 0.f ? lphi[2 * blockIdx.x + 1] - (rex.m + logf(rex.s)) : 0.f;   // no other item: all -inf
  __syncthreads();
  for (int j = threadIdx.x; j < V; j += blockDim.x) z[j] += c_ex;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) z[iid[s0 + i]] = zin[s0 + i] + c_in;
}

// Backward of the map above.  Upstream gradient of the log-probs: G[B, ldg], or (G == NULL) the mean-NLL loss of `labels`
// scaled by gscale[0] (g[b, v] = -gscale / B at the label).  With s_g = sum of g over group g:
//   dZ[b, v] = scale * (g[b, v] - exp(LP[b, v] - log phi[b, g(v)]) * s_g(v)),   d log phi[b, g] = s_g.
// d log phi is returned times dl_scale (callers whose G already carries `scale` pass 1 / scale).  DZ may alias LP or G: the
// session's own entries of both are parked in tmp[2 N] before anything is written.
__global__ void __launch_bounds__(512) renorm_head_bwd_kernel(const float* LP, long long ldlp, const float* __restrict__ G,
                                                              long long ldg, const int* __restrict__ labels,
                                                              const float* __restrict__ gscale, float scale, float dl_scale,
                                                              int B, int V, const int* __restrict__ iid,
                                                              const int* __restrict__ seg, const float* __restrict__ lphi,
                                                              float* __restrict__ tmp, float* DZ, long long lddz, float* DZlo,
                                                              float* __restrict__ dlphi) {
  SRK_PDL();
  __shared__ float red[2][16];
  __shared__ float tot[2];
  const float* lp = LP + (long long)blockIdx.x * ldlp;
  const float* g = G ? G + (long long)blockIdx.x * ldg : nullptr;
  float* dz = DZ + (long long)blockIdx.x * lddz;
  float* dzl = DZlo ? DZlo + (long long)blockIdx.x * lddz : nullptr;
  const int s0 = seg[blockIdx.x], n = seg[blockIdx.x + 1] - s0;
  float* tlp = tmp + 2 * (long long)s0;
  float* tg = tlp + n;
  const int lab = g ? -1 : labels[blockIdx.x];
  const float c = g ? 0.f : gscale[0] / (float)B;
  float s_all = 0.f, s_in = 0.f;
  if (g) {
    for (int j = threadIdx.x; j < V; j += blockDim.x) s_all += g[j];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float x = g[iid[s0 + i]];
      tg[i] = x;
      s_in += x;
    }
  } else {
    if (threadIdx.x == 0) s_all = -c;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      if (iid[s0 + i] == lab) s_in = -c;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) tlp[i] = lp[iid[s0 + i]];
  s_all = warp_sum(s_all);
  s_in = warp_sum(s_in);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s_all;
    red[1][threadIdx.x >> 5] = s_in;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    float t = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[threadIdx.x][w];
    tot[threadIdx.x] = t;
  }
  __syncthreads();
  const float sum_in = tot[1], sum_ex = tot[0] - tot[1];
  const float lp_in = lphi[2 * blockIdx.x], lp_ex = lphi[2 * blockIdx.x + 1];
  for (int j = threadIdx.x; j < V; j += blockDim.x) {
    const float gj = g ? g[j] : (j == lab ? -c : 0.f);
    store_split(dz, dzl, j, scale * (gj - expf(lp[j] - lp_ex) * sum_ex));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int v = iid[s0 + i];
    const float gj = g ? tg[i] : (v == lab ? -c : 0.f);
    store_split(dz, dzl, v, scale * (gj - expf(tlp[i] - lp_in) * sum_in));
  }
  if (threadIdx.x == 0) {
    dlphi[2 * blockIdx.x] = dl_scale * sum_in;
    dlphi[2 * blockIdx.x + 1] = dl_scale * sum_ex;
  }
}

// Gate of the REnorm head: phi = softmax(W2 relu(h)) with h = W1 shat + b1 already in H (msgifsr.py:206,283).
// Warp per session; H is rewritten with relu(h), lphi[b, :] = log phi.
__global__ void __launch_bounds__(256) gate_fwd_kernel(float* __restrict__ H, const float* __restrict__ W2, int B, int d,
                                                       float* __restrict__ lphi) {
  SRK_PDL();
  const int lane = threadIdx.x & 31;
  for (int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < B; b += gridDim.x * (blockDim.x >> 5)) {
    float* h = H + (long long)b * d;
    float a0 = 0.f, a1 = 0.f;
    for (int i = lane; i < d; i += 32) {
      const float r = fmaxf(h[i], 0.f);
      h[i] = r;
      a0 += r * W2[i];
      a1 += r * W2[d + i];
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    const float m = fmaxf(a0, a1);
    const float l = m + logf(expf(a0 - m) + expf(a1 - m));
    if (lane == 0) {
      lphi[2 * b] = a0 - l;
      lphi[2 * b + 1] = a1 - l;
    }
  }
}

// da = dlphi - phi * (dlphi_0 + dlphi_1) (gradient at the two gate logits), dH = relu'(h) * (W2^T da).
__global__ void __launch_bounds__(256) gate_bwd_kernel(const float* __restrict__ Hr, const float* __restrict__ W2,
                                                       const float* __restrict__ lphi, const float* __restrict__ dlphi, int B,
                                                       int d, float* __restrict__ da, float* __restrict__ dH) {
  SRK_PDL();
  const int lane = threadIdx.x & 31;
  for (int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < B; b += gridDim.x * (blockDim.x >> 5)) {
    const float g0 = dlphi[2 * b], g1 = dlphi[2 * b + 1];
    const float d0 = g0 - expf(lphi[2 * b]) * (g0 + g1), d1 = g1 - expf(lphi[2 * b + 1]) * (g0 + g1);
    if (lane == 0) {
      da[2 * b] = d0;
      da[2 * b + 1] = d1;
    }
    for (int i = lane; i < d; i += 32)
      dH[(long long)b * d + i] = Hr[(long long)b * d + i] > 0.f ? d0 * W2[i] + d1 * W2[d + i] : 0.f;
  }
}

inline int row_grid(long long rows) {
  long long g = (rows + 7) / 8;
  if (g < 1) g = 1;
  if (g > 148LL * 64) g = 148LL * 64;
  return (int)g;
}

}  // namespace

extern "C" int srk_readout_fwd(const float* F, const float* u, const float* v, const float* we, const int* seg,
                               const int* last, int B, int d, int with_last, float* e, float* ms, float* sr_in,
                               void* stream) {
  SRK_TRY(srk_check_dim(d));
  if (B <= 0) return SRK_OK;
  SRK_DISPATCH_NC(d, (srk_launch(readout_fwd_kernel<NC>, row_grid(B), 256, 0, (cudaStream_t)stream, F, u, v, we, seg, last, B, d,
                                                                                            with_last, e, ms, sr_in)));
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

int srk_readout_bwd_split(const float* F, float* u, float* v, const float* we, const int* seg, const int* last, const float* e,
                          const float* ms, const float* sr_in, const float* dsr_in, int B, int d, int with_last, float* dF,
                          float* dwe, float* duh, float* dul, void* stream);

extern "C" int srk_readout_bwd(const float* F, float* u, float* v, const float* we, const int* seg, const int* last,
                               const float* e, const float* ms, const float* sr_in, const float* dsr_in, int B, int d,
                               int with_last, float* dF, float* dwe, void* stream) {
  return srk_readout_bwd_split(F, u, v, we, seg, last, e, ms, sr_in, dsr_in, B, d, with_last, dF, dwe, nullptr, nullptr, stream);
}

// duh / dul: also write the TF32 hi / lo pair of du (dense [N, d]) for the tensor-core products that consume it
int srk_readout_bwd_split(const float* F, float* u, float* v, const float* we, const int* seg, const int* last, const float* e,
                          const float* ms, const float* sr_in, const float* dsr_in, int B, int d, int with_last, float* dF,
                          float* dwe, float* duh, float* dul, void* stream) {
  SRK_TRY(srk_check_dim(d));
  if (B <= 0) return SRK_OK;
  int grid = row_grid(B);
  if (grid > 2 * 148) grid = 2 * 148;          // sessions beyond that are taken by the grid-stride loop (fewer atomics on d w_e)
  SRK_DISPATCH_NC(d, (srk_launch(readout_bwd_kernel<NC>, grid, 256, 0, (cudaStream_t)stream, F, u, v, we, seg, last, e, ms, sr_in, dsr_in, B, d, with_last, dF, dwe, duh, dul)));
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_ce_rows_fwd(float* Z, long long ldz, const int* labels, int B, int V, int write_logp, float* lse,
                               float* nll, void* stream) {
  if (B <= 0) return SRK_OK;
  SRK_REQUIRE(V > 0, "ce_rows_fwd: empty catalog");
  srk_launch(ce_rows_fwd_kernel, B, 512, 0, (cudaStream_t)stream, Z, ldz, labels, V, write_logp, lse, nll);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_ce_rows_bwd(float* Z, long long ldz, const int* labels, const float* lse, const float* gscale,
                               float scale, int B, int V, int z_is_logp, float* Zlo, void* stream) {
  if (B <= 0) return SRK_OK;
  srk_launch(ce_rows_bwd_kernel, B, 512, 0, (cudaStream_t)stream, Z, ldz, labels, lse, gscale, scale, B, V, z_is_logp, Zlo, 0);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_logp_bwd(const float* LP, long long ldlp, const float* G, long long ldg, float scale, int B, int V,
                            float* DZ, long long lddz, float* DZlo, void* stream) {
  if (B <= 0) return SRK_OK;
  srk_launch(logp_bwd_kernel, B, 512, 0, (cudaStream_t)stream, LP, ldlp, G, ldg, scale, V, DZ, lddz, DZlo);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_ce_rows_bwd_cols(float* Z, long long ldz, const int* labels, const float* lse, const float* gscale,
                                    float scale, int B, int col0, int ncols, float* Zlo, void* stream) {
  if (B <= 0 || ncols <= 0) return SRK_OK;
  srk_launch(ce_rows_bwd_kernel, B, 512, 0, (cudaStream_t)stream, Z, ldz, labels, lse, gscale, scale, B, ncols, 0, Zlo, col0);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_renorm_head_fwd(float* Z, long long ldz, int B, int V, const int* iid, const int* seg, const float* lphi,
                                   float* zin, void* stream) {
  if (B <= 0) return SRK_OK;
  SRK_REQUIRE(V > 0, "renorm_head_fwd: empty catalog");
  srk_launch(renorm_head_fwd_kernel, B, 512, 0, (cudaStream_t)stream, Z, ldz, V, iid, seg, lphi, zin);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_renorm_head_bwd(const float* LP, long long ldlp, const float* G, long long ldg, const int* labels,
                                   const float* gscale, float scale, float dl_scale, int B, int V, const int* iid,
                                   const int* seg, const float* lphi, float* tmp, float* DZ, long long lddz, float* DZlo,
                                   float* dlphi, void* stream) {
  if (B <= 0) return SRK_OK;
  SRK_REQUIRE(G != nullptr || (labels != nullptr && gscale != nullptr), "renorm_head_bwd: needs G or (labels, gscale)");
  srk_launch(renorm_head_bwd_kernel, B, 512, 0, (cudaStream_t)stream, LP, ldlp, G, ldg, labels, gscale, scale, dl_scale, B, V, iid,
             seg, lphi, tmp, DZ, lddz, DZlo, dlphi);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_gate_fwd(float* H, const float* W2, int B, int d, float* lphi, void* stream) {
  if (B <= 0) return SRK_OK;
  srk_launch(gate_fwd_kernel, row_grid(B), 256, 0, (cudaStream_t)stream, H, W2, B, d, lphi);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_gate_bwd(const float* Hr, const float* W2, const float* lphi, const float* dlphi, int B, int d, float* da,
                            float* dH, void* stream) {
  if (B <= 0) return SRK_OK;
  srk_launch(gate_bwd_kernel, row_grid(B), 256, 0, (cudaStream_t)stream, Hr, W2, lphi, dlphi, B, d, da, dH);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}
