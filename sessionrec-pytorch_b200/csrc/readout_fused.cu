// Fused tails of the attention readout (warp per session, rows in float4 lanes - rowops.cuh):
//   forward : e_i = <w_e, sigmoid(u_i + v_b)> -> segment soft-max -> g_b -> sr_in = [F[last_b] | g_b] -> s = fc_sr(sr_in)
//             -> normalise -> bf16 hi/lo for the scoring head          (srgnn.py:82-91,141-143; msgifsr.py:141-146,269-273)
//   backward: d shat -> d s (normalise backward) -> d sr_in = d s W_sr -> attention backward -> d u, d v, d w_e,
//             d F = alpha d g (+ d l on the last row)
// The N x d x d projections (u = fc_u(F), d F += d u W_u, ...) stay node-parallel GEMMs; what is fused here is the chain of
// per-session steps that used to be one launch each (readout, fc_sr GEMM, row normalisation, bf16 split / row-norm
// backward, d sr_in GEMM, readout backward): every launch on the critical path of this latency-bound encoder costs
// ~5 us whatever its size.  The per-session mat-vecs read W^T / W from global memory through L1 with the lanes on
// consecutive columns; the session vector is broadcast from a per-warp shared-memory row.
// (A first version that also fused the projections with one thread block slot per session / node was measured 5-10x
// slower: it serialises the nodes of a session and is bound by the longest session of the batch.)
#include <float.h>

#include "rowops.cuh"

namespace {

constexpr int RT_THREADS = 256;     // 8 warps = 8 sessions in flight per CTA

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

// out[col .. col + 3] = sum_k x[k] * Wt[k][col0 + col .. + 3] for this lane's columns; Wt is [K, ld] row-major, x a
// shared-memory row (broadcast reads).  The weights come through a cold L1 (one CTA per SM, every line is touched once),
// so the loop is a chain of L2 round trips unless many loads are in flight: 16 rows are fetched before any is used.
constexpr int MV_DEPTH = 16;
template <int NC>
__device__ __forceinline__ void matvec_cols(RowVec<NC>& out, const float* __restrict__ Wt, int ld, int col0, const float* x, int K,
                                            int d, int lane) {
  row_zero(out);
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int col = (c * 32 + lane) * 4;
    if (col >= d) continue;
    const float* w = Wt + col0 + col;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    for (int k0 = 0; k0 < K; k0 += MV_DEPTH) {
      float4 r[MV_DEPTH];
#pragma unroll
      for (int t = 0; t < MV_DEPTH; ++t)
        r[t] = k0 + t < K ? __ldg(reinterpret_cast<const float4*>(w + (long long)(k0 + t) * ld)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int t = 0; t < MV_DEPTH; t += 2) {
        const float x0 = k0 + t < K ? x[k0 + t] : 0.f, x1 = k0 + t + 1 < K ? x[k0 + t + 1] : 0.f;
        a0.x = fmaf(x0, r[t].x, a0.x); a0.y = fmaf(x0, r[t].y, a0.y); a0.z = fmaf(x0, r[t].z, a0.z); a0.w = fmaf(x0, r[t].w, a0.w);
        a1.x = fmaf(x1, r[t + 1].x, a1.x); a1.y = fmaf(x1, r[t + 1].y, a1.y);
        a1.z = fmaf(x1, r[t + 1].z, a1.z); a1.w = fmaf(x1, r[t + 1].w, a1.w);
      }
    }
    out.v[c] = make_float4(a0.x + a1.x, a0.y + a1.y, a0.z + a1.z, a0.w + a1.w);
  }
}

template <int NC>
__global__ void __launch_bounds__(RT_THREADS) readout_tail_fwd_kernel(
    const float* __restrict__ F, const float* __restrict__ u, const float* __restrict__ v, const float* __restrict__ we,
    const float* __restrict__ WsrT, const int* __restrict__ seg, const int* __restrict__ last, int B, int d, int norm_mode,
    float* __restrict__ e, float* __restrict__ ms, float* __restrict__ sr_in, float* __restrict__ s, float* __restrict__ shat,
    float* __restrict__ rn_s, uint16_t* __restrict__ sbh, uint16_t* __restrict__ sbl) {
  SRK_PDL();
  extern __shared__ float xsm[];                    // [warps][2 d]
  const int lane = threadIdx.x & 31;
  float* xs = xsm + (threadIdx.x >> 5) * 2 * d;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  RowVec<NC> wev;
  row_load(wev, we, d, lane);
  for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < B; b += warps) {
    RowVec<NC> vb, acc;
    row_load(vb, v + (long long)b * d, d, lane);
    row_zero(acc);
    float m = -FLT_MAX, sum = 0.f;
    for (int i = seg[b]; i < seg[b + 1]; ++i) {
      RowVec<NC> ui, fi;
      row_load(ui, u + (long long)i * d, d, lane);
      row_load(fi, F + (long long)i * d, d, lane);
      const float ei = attn_score(ui, vb, wev);
      if (lane == 0) e[i] = ei;
      const float mn = fmaxf(m, ei);
      const float corr = expf(m - mn), pi = expf(ei - mn);
      sum = sum * corr + pi;
      row_scale(acc, corr);
      row_axpy(acc, pi, fi);
      m = mn;
    }
    row_scale(acc, 1.f / sum);
    if (lane == 0) {
      ms[2 * b] = m;
      ms[2 * b + 1] = sum;
    }
    RowVec<NC> fl;
    row_load(fl, F + (long long)last[b] * d, d, lane);
    row_store(fl, sr_in + (long long)b * 2 * d, d, lane);
    row_store(acc, sr_in + (long long)b * 2 * d + d, d, lane);
    // s = W_sr [F_last | g]: the 2 d inputs are broadcast from shared memory, the lanes own the output columns
    row_store(fl, xs, d, lane);
    row_store(acc, xs + d, d, lane);
    __syncwarp();
    RowVec<NC> sv, yv;
    matvec_cols<NC>(sv, WsrT, d, 0, xs, 2 * d, d, lane);
    __syncwarp();
    const float n = row_normalize(sv, yv, norm_mode);
    row_store(sv, s + (long long)b * d, d, lane);
    row_store(yv, shat + (long long)b * d, d, lane);
    if (rn_s && lane == 0) rn_s[b] = n;
    if (sbh) row_store_bf16_split(yv, sbh + (long long)b * d, sbl + (long long)b * d, d, lane);
  }
}

template <int NC>
__global__ void __launch_bounds__(RT_THREADS) readout_head_bwd_kernel(
    const float* __restrict__ F, const float* __restrict__ we, const float* __restrict__ Wsr, const int* __restrict__ seg,
    const int* __restrict__ last, int B, int d, int norm_mode, const float* __restrict__ s, const float* __restrict__ shat,
    const float* __restrict__ rn_s, const float* __restrict__ sr_in, const float* __restrict__ e, const float* __restrict__ ms,
    const float* __restrict__ dshat, float* __restrict__ u, float* __restrict__ v, float* __restrict__ ds, float* __restrict__ dF,
    float* __restrict__ dwe) {
  SRK_PDL();
  extern __shared__ float xsm[];                    // [warps][d]
  const int lane = threadIdx.x & 31;
  float* xs = xsm + (threadIdx.x >> 5) * d;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  RowVec<NC> wev, dwe_acc;
  row_load(wev, we, d, lane);
  row_zero(dwe_acc);
  for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < B; b += warps) {
    // d s = normalise-backward(d shat); d sr_in = d s W_sr = [d l | d g]
    RowVec<NC> xv, yv, dy, dsv, dl, dg;
    row_load(xv, s + (long long)b * d, d, lane);
    row_load(yv, shat + (long long)b * d, d, lane);
    row_load(dy, dshat + (long long)b * d, d, lane);
    row_normalize_bwd<NC>(xv, yv, norm_mode == SRK_NORM_NONE ? 0.f : rn_s[b], norm_mode, dy, nullptr, dsv);
    row_store(dsv, ds + (long long)b * d, d, lane);
    row_store(dsv, xs, d, lane);
    __syncwarp();
    matvec_cols<NC>(dl, Wsr, 2 * d, 0, xs, d, d, lane);
    matvec_cols<NC>(dg, Wsr, 2 * d, d, xs, d, d, lane);
    __syncwarp();
    RowVec<NC> vb, g, dv;
    row_load(vb, v + (long long)b * d, d, lane);
    row_load(g, sr_in + (long long)b * 2 * d + d, d, lane);
    row_zero(dv);
    const float t = row_dot(g, dg);
    const float m = ms[2 * b], inv_s = 1.f / ms[2 * b + 1];
    const int lb = last[b];
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
      df = dg;
      row_scale(df, alpha);
      if (i == lb) row_axpy(df, 1.f, dl);
      row_store(df, dF + (long long)i * d, d, lane);
    }
    row_store(dv, v + (long long)b * d, d, lane);
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    int col = (c * 32 + lane) * 4;
    if (col < d) {
      atomicAdd(dwe + col, dwe_acc.v[c].x); atomicAdd(dwe + col + 1, dwe_acc.v[c].y);
      atomicAdd(dwe + col + 2, dwe_acc.v[c].z); atomicAdd(dwe + col + 3, dwe_acc.v[c].w);
    }
  }
}

__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ X, int rows, int cols, float* __restrict__ Y) {
  SRK_PDL();
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = X[(long long)r * cols + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) Y[(long long)c * rows + r] = tile[threadIdx.x][i];
  }
}

inline int session_grid(int B) {
  int g = srk_cdiv(B, RT_THREADS / 32);
  return g < 1 ? 1 : g;
}

}  // namespace

extern "C" int srk_transpose(const float* X, int rows, int cols, float* Y, void* stream) {
  if (rows <= 0 || cols <= 0) return SRK_OK;
  srk_launch(transpose_kernel, dim3(srk_cdiv(cols, 32), srk_cdiv(rows, 32)), dim3(32, 8), 0, (cudaStream_t)stream, X, rows, cols, Y);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_readout_tail_fwd(const float* F, const float* u, const float* v, const float* we, const float* WsrT,
                                    const int* seg, const int* last, int B, int d, int norm_mode, float* e, float* ms,
                                    float* sr_in, float* s, float* shat, float* rn_s, uint16_t* sbh, uint16_t* sbl,
                                    void* stream) {
  SRK_TRY(srk_check_dim(d));
  if (B <= 0) return SRK_OK;
  SRK_REQUIRE(norm_mode == SRK_NORM_NONE || norm_mode == SRK_NORM_L2 || norm_mode == SRK_NORM_EPS,
              "readout_tail_fwd: norm_mode must be NONE, L2 or EPS");
  SRK_REQUIRE((sbh == nullptr) == (sbl == nullptr), "readout_tail_fwd: bf16 outputs come as a pair");
  const size_t smem = (size_t)(RT_THREADS / 32) * 2 * d * sizeof(float);
  SRK_REQUIRE(smem <= 48 * 1024, "readout_tail_fwd: d = %d too wide for the per-warp staging rows", d);
  SRK_DISPATCH_NC(d, (srk_launch(readout_tail_fwd_kernel<NC>, session_grid(B), RT_THREADS, smem, (cudaStream_t)stream, F, u, v, we, WsrT, seg, last, B, d, norm_mode, e, ms, sr_in, s, shat, rn_s, sbh, sbl)));
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_readout_head_bwd(const float* F, const float* we, const float* Wsr, const int* seg, const int* last, int B,
                                    int d, int norm_mode, const float* s, const float* shat, const float* rn_s,
                                    const float* sr_in, const float* e, const float* ms, const float* dshat, float* u, float* v,
                                    float* ds, float* dF, float* dwe, void* stream) {
  SRK_TRY(srk_check_dim(d));
  if (B <= 0) return SRK_OK;
  SRK_REQUIRE(norm_mode == SRK_NORM_NONE || norm_mode == SRK_NORM_L2 || norm_mode == SRK_NORM_EPS,
              "readout_head_bwd: norm_mode must be NONE, L2 or EPS");
  SRK_REQUIRE(norm_mode == SRK_NORM_NONE || rn_s != nullptr, "readout_head_bwd: rn_s is required for a normalised head");
  const size_t smem = (size_t)(RT_THREADS / 32) * d * sizeof(float);
  SRK_REQUIRE(smem <= 48 * 1024, "readout_head_bwd: d = %d too wide for the per-warp staging rows", d);
  SRK_DISPATCH_NC(d, (srk_launch(readout_head_bwd_kernel<NC>, session_grid(B), RT_THREADS, smem, (cudaStream_t)stream, F, we, Wsr, seg, last, B, d, norm_mode, s, shat, rn_s, sr_in, e, ms, dshat, u, v, ds, dF, dwe)));
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}
