// Order-fusion head of MSGIFSR (msgifsr.py:311-315,321): score = sum_k softmax(alpha)_k * softmax(12 sr_k E^T), out =
// log(score).  The K per-order logit matrices live in one buffer Zall[K][B][ldz]; lse[K][B] are their row
// log-sum-exps.  Row kernels (one CTA per session), HBM-bound streaming over K * V floats.
#include <float.h>

#include "common.cuh"

namespace {

constexpr int MAXK = 4;

struct MixArgs {
  float* Z;            // [K][B][ldz]  logits in, gradients out (in place)
  float* Zlo;          // optional TF32 low halves of the gradients
  long long head_stride, ldz;
  const float* lse;    // [K][B]
  const float* alpha;  // [K] un-normalised mixture logits (device): a = softmax(alpha)
  int K, B, V;
};

__device__ __forceinline__ void mix_weights(const MixArgs& m, float* a, float* loga) {
  float mx = -FLT_MAX, s = 0.f;
#pragma unroll
  for (int k = 0; k < MAXK; ++k)
    if (k < m.K) mx = fmaxf(mx, m.alpha[k]);
#pragma unroll
  for (int k = 0; k < MAXK; ++k)
    if (k < m.K) s += expf(m.alpha[k] - mx);
  const float ls = mx + logf(s);
#pragma unroll
  for (int k = 0; k < MAXK; ++k) {
    loga[k] = k < m.K ? m.alpha[k] - ls : 0.f;
    a[k] = k < m.K ? expf(loga[k]) : 0.f;
  }
}

__global__ void __launch_bounds__(512) mix_logp_kernel(MixArgs m, float* __restrict__ out, long long ldo) {
  SRK_PDL();
  const int b = blockIdx.x;
  float wa[MAXK], wl[MAXK];
  mix_weights(m, wa, wl);
  float l[MAXK];
#pragma unroll
  for (int k = 0; k < MAXK; ++k) l[k] = k < m.K ? wl[k] - m.lse[(long long)k * m.B + b] : 0.f;
  for (int v = threadIdx.x; v < m.V; v += blockDim.x) {
    float t[MAXK], mx = -FLT_MAX;
#pragma unroll
    for (int k = 0; k < MAXK; ++k)
      if (k < m.K) {
        t[k] = m.Z[k * m.head_stride + (long long)b * m.ldz + v] + l[k];
        mx = fmaxf(mx, t[k]);
      }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < MAXK; ++k)
      if (k < m.K) s += expf(t[k] - mx);
    out[(long long)b * ldo + v] = mx + logf(s);
  }
}

// loss[b] = -log sum_k a_k exp(-nll_k[b]); mean over b.  Single CTA.
__global__ void __launch_bounds__(1024) mix_loss_kernel(const float* __restrict__ nll, MixArgs m, float* __restrict__ out) {
  SRK_PDL();
  __shared__ float red[32];
  float wa[MAXK], wl[MAXK];
  mix_weights(m, wa, wl);
  float s = 0.f;
  for (int b = threadIdx.x; b < m.B; b += blockDim.x) {
    float mx = -FLT_MAX, t[MAXK];
#pragma unroll
    for (int k = 0; k < MAXK; ++k)
      if (k < m.K) {
        t[k] = wl[k] - nll[(long long)k * m.B + b];
        mx = fmaxf(mx, t[k]);
      }
    float e = 0.f;
#pragma unroll
    for (int k = 0; k < MAXK; ++k)
      if (k < m.K) e += expf(t[k] - mx);
    s -= mx + logf(e);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) out[0] = s / (float)m.B;
  }
}

// dZ_k[b, v] = scale * (G[b, v] q_k[b, v] - P_k[b, v] r_k[b]),  q_k = a_k P_k / sum_j a_j P_j,  r_k[b] = sum_v G q_k.
// G == NULL: G = -(gscale / B) onehot(label) (fused-loss form).  rsum[K][B] receives r_k (for d alpha).
__global__ void __launch_bounds__(512) mix_bwd_kernel(MixArgs m, const float* __restrict__ G, long long ldg,
                                                      const int* __restrict__ labels, const float* __restrict__ gscale,
                                                      float scale, float* __restrict__ rsum) {
  SRK_PDL();
  __shared__ float red[MAXK][16];
  __shared__ float rk[MAXK];
  const int b = blockIdx.x;
  float wa[MAXK], wl[MAXK];
  mix_weights(m, wa, wl);
  float l[MAXK];
#pragma unroll
  for (int k = 0; k < MAXK; ++k) l[k] = k < m.K ? m.lse[(long long)k * m.B + b] : 0.f;
  float r[MAXK];
#pragma unroll
  for (int k = 0; k < MAXK; ++k) r[k] = 0.f;
  const int lab = labels ? labels[b] : -1;
  const float gl = labels ? -gscale[0] / (float)m.B : 0.f;
  for (int v = threadIdx.x; v < m.V; v += blockDim.x) {
    const float g = G ? G[(long long)b * ldg + v] : (v == lab ? gl : 0.f);
    if (g == 0.f) continue;
    float p[MAXK], sc = 0.f;
#pragma unroll
    for (int k = 0; k < MAXK; ++k)
      if (k < m.K) {
        p[k] = wa[k] * expf(m.Z[k * m.head_stride + (long long)b * m.ldz + v] - l[k]);
        sc += p[k];
      }
#pragma unroll
    for (int k = 0; k < MAXK; ++k)
      if (k < m.K) r[k] += g * p[k] / sc;
  }
#pragma unroll
  for (int k = 0; k < MAXK; ++k) {
    float s = warp_sum(r[k]);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = s;
  }
  __syncthreads();
  if (threadIdx.x < MAXK) {
    float s = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) s += red[threadIdx.x][w];
    rk[threadIdx.x] = s;
    if (threadIdx.x < m.K) rsum[(long long)threadIdx.x * m.B + b] = s;
  }
  __syncthreads();
  for (int v = threadIdx.x; v < m.V; v += blockDim.x) {
    const float g = G ? G[(long long)b * ldg + v] : (v == lab ? gl : 0.f);
    float P[MAXK], sc = 0.f;
#pragma unroll
    for (int k = 0; k < MAXK; ++k)
      if (k < m.K) {
        P[k] = expf(m.Z[k * m.head_stride + (long long)b * m.ldz + v] - l[k]);
        sc += wa[k] * P[k];
      }
#pragma unroll
    for (int k = 0; k < MAXK; ++k)
      if (k < m.K) {
        float d = scale * (g * wa[k] * P[k] / sc - P[k] * rk[k]);
        const long long o = k * m.head_stride + (long long)b * m.ldz + v;
        if (m.Zlo) {
          float h = __uint_as_float(__float_as_uint(d) & 0xFFFFE000u);
          m.Z[o] = h;
          m.Zlo[o] = d - h;
        } else {
          m.Z[o] = d;
        }
      }
  }
}

// d alpha_j += a_j (da_j - sum_k a_k da_k),  da_k = (sum_b rsum[k][b]) / a_k.  Single CTA.
__global__ void __launch_bounds__(256) mix_alpha_bwd_kernel(const float* __restrict__ rsum, MixArgs m,
                                                            float* __restrict__ dalpha) {
  SRK_PDL();
  __shared__ float tot[MAXK];
  __shared__ float red[MAXK][8];
  float wa[MAXK], wl[MAXK];
  mix_weights(m, wa, wl);
#pragma unroll
  for (int k = 0; k < MAXK; ++k) {
    float s = 0.f;
    if (k < m.K)
      for (int b = threadIdx.x; b < m.B; b += blockDim.x) s += rsum[(long long)k * m.B + b];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = s;
  }
  __syncthreads();
  if (threadIdx.x < MAXK) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
    tot[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float da[MAXK], dot = 0.f;
    for (int k = 0; k < m.K; ++k) {
      da[k] = tot[k] / wa[k];
      dot += wa[k] * da[k];
    }
    for (int k = 0; k < m.K; ++k) dalpha[k] += wa[k] * (da[k] - dot);
  }
}

int fill_args(MixArgs& m, float* Z, float* Zlo, long long head_stride, long long ldz, const float* lse, const float* alpha,
              int K, int B, int V) {
  SRK_REQUIRE(K >= 1 && K <= MAXK, "mixture: 1..%d orders", MAXK);
  m.Z = Z; m.Zlo = Zlo; m.head_stride = head_stride; m.ldz = ldz; m.lse = lse; m.alpha = alpha; m.K = K; m.B = B; m.V = V;
  return SRK_OK;
}

}  // namespace

extern "C" int srk_mix_logp_fwd(const float* Zall, long long head_stride, long long ldz, const float* lse,
                                const float* alpha, int K, int B, int V, float* out, long long ldo, void* stream) {
  if (B <= 0) return SRK_OK;
  MixArgs m;
  SRK_TRY(fill_args(m, const_cast<float*>(Zall), nullptr, head_stride, ldz, lse, alpha, K, B, V));
  srk_launch(mix_logp_kernel, B, 512, 0, (cudaStream_t)stream, m, out, ldo);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_mix_loss_fwd(const float* nll, const float* alpha, int K, int B, float* loss_out, void* stream) {
  MixArgs m;
  SRK_TRY(fill_args(m, nullptr, nullptr, 0, 0, nullptr, alpha, K, B, 0));
  srk_launch(mix_loss_kernel, 1, 1024, 0, (cudaStream_t)stream, nll, m, loss_out);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_mix_bwd(float* Zall, float* Zlo_all, long long head_stride, long long ldz, const float* lse,
                           const float* alpha, int K, int B, int V, const float* G, long long ldg, const int* labels,
                           const float* gscale, float scale, float* rsum, void* stream) {
  if (B <= 0) return SRK_OK;
  SRK_REQUIRE((G != nullptr) != (labels != nullptr), "mix_bwd: give either G or labels (+ gscale)");
  MixArgs m;
  SRK_TRY(fill_args(m, Zall, Zlo_all, head_stride, ldz, lse, alpha, K, B, V));
  srk_launch(mix_bwd_kernel, B, 512, 0, (cudaStream_t)stream, m, G, ldg, labels, gscale, scale, rsum);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_mix_alpha_bwd(const float* rsum, const float* alpha, int K, int B, float* dalpha, void* stream) {
  MixArgs m;
  SRK_TRY(fill_args(m, nullptr, nullptr, 0, 0, nullptr, alpha, K, B, 0));
  srk_launch(mix_alpha_bwd_kernel, 1, 256, 0, (cudaStream_t)stream, rsum, m, dalpha);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}
