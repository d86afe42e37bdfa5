// Generic fp32 CUDA-core GEMM used for the small dense layers of the path (GAT fc, GGNN GRU, readout,
// fc_sr) and as the exact-fp32 catalog GEMM (the tcgen05 3xTF32 kernel in score_umma.cu takes over the
// catalog contraction when shapes allow).
//
//   C[m, n] (op)= alpha * sum_k A(m, k) * B(k, n)  (+ bias[n])
//   A(m, k) = A[ra(m) * sa_m + k * sa_k],  ra(m) = a_idx ? a_idx[m] : m
//   B(k, n) = B[rb(k) * sb_k + n * sb_n],  rb(k) = b_idx ? b_idx[k] : k
//   C row m is written at C[rc(m) * ldc + n]
//
// 64x64x16 tiles, 256 threads, 4x4 register micro-tile, register-staged prefetch, optional split-K with
// an atomicAdd epilogue (accumulate mode only).
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4, NT = 256;

enum LoadMode { LM_SCALAR = 0, LM_KVEC = 1, LM_MVEC = 2, LM_KSCAL = 3, LM_MSCAL = 4 };

struct Frag {
  float v[4];
};

// Loads this thread's 4 elements of a (64 x 16) operand tile.  `outer` indexes the 64-wide dimension
// (m for A, n for B), `k` the 16-deep one.  so/sk are the element strides, idx the optional indirection
// (on `outer` when idx_on_outer, else on k).
template <bool kIdxOnOuter>
__device__ __forceinline__ void load_frag(Frag& f, const float* __restrict__ P, long long so, long long sk,
                                          const int* __restrict__ idx, int mode, int o0, int k0, int O, int kEnd,
                                          int tid) {
  if (mode == LM_KVEC || mode == LM_KSCAL) {
    int o = o0 + (tid >> 2), k = k0 + (tid & 3) * 4;
    f.v[0] = f.v[1] = f.v[2] = f.v[3] = 0.f;
    if (o < O) {
      long long ro = kIdxOnOuter && idx ? (long long)idx[o] : (long long)o;
      const float* p = P + ro * so;
      if (mode == LM_KVEC && k + 3 < kEnd && !(!kIdxOnOuter && idx)) {
        float4 t = *reinterpret_cast<const float4*>(p + k);
        f.v[0] = t.x; f.v[1] = t.y; f.v[2] = t.z; f.v[3] = t.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (k + j < kEnd) {
            long long rk = (!kIdxOnOuter && idx) ? (long long)idx[k + j] : (long long)(k + j);
            f.v[j] = p[rk * sk];
          }
      }
    }
  } else if (mode == LM_MVEC || mode == LM_MSCAL) {
    int k = k0 + (tid >> 4), o = o0 + (tid & 15) * 4;
    f.v[0] = f.v[1] = f.v[2] = f.v[3] = 0.f;
    if (k < kEnd) {
      long long rk = (!kIdxOnOuter && idx) ? (long long)idx[k] : (long long)k;
      const float* p = P + rk * sk;
      if (mode == LM_MVEC && o + 3 < O && !(kIdxOnOuter && idx)) {
        float4 t = *reinterpret_cast<const float4*>(p + o);
        f.v[0] = t.x; f.v[1] = t.y; f.v[2] = t.z; f.v[3] = t.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (o + j < O) {
            long long ro = (kIdxOnOuter && idx) ? (long long)idx[o + j] : (long long)(o + j);
            f.v[j] = p[ro * so];
          }
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int e = tid + j * NT;
      int o = o0 + (e & 63), k = k0 + (e >> 6);
      float v = 0.f;
      if (o < O && k < kEnd) {
        long long ro = (kIdxOnOuter && idx) ? (long long)idx[o] : (long long)o;
        long long rk = (!kIdxOnOuter && idx) ? (long long)idx[k] : (long long)k;
        v = P[ro * so + rk * sk];
      }
      f.v[j] = v;
    }
  }
}

__device__ __forceinline__ void store_frag(const Frag& f, float (*S)[BM + PAD], int mode, int tid) {
  if (mode == LM_KVEC || mode == LM_KSCAL) {
    int o = tid >> 2, k = (tid & 3) * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) S[k + j][o] = f.v[j];
  } else if (mode == LM_MVEC || mode == LM_MSCAL) {
    int k = tid >> 4, o = (tid & 15) * 4;
    *reinterpret_cast<float4*>(&S[k][o]) = make_float4(f.v[0], f.v[1], f.v[2], f.v[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int e = tid + j * NT;
      S[e >> 6][e & 63] = f.v[j];
    }
  }
}

struct GemmParams {
  GemmArgs g;
  int a_mode, b_mode, k_chunk, k_chunk_pad;
};

__global__ void __launch_bounds__(NT) sgemm_kernel(const GemmParams p) {
  SRK_PDL();
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const GemmArgs& g = p.g;
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kBeg = blockIdx.z * p.k_chunk;
  const int kEnd = min(g.K, kBeg + p.k_chunk);
  const int ty = tid >> 4, tx = tid & 15;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  Frag fa, fb;
  if (kBeg < kEnd) {
    load_frag<true>(fa, g.A, g.sa_m, g.sa_k, g.a_idx, p.a_mode, m0, kBeg, g.M, kEnd, tid);
    load_frag<false>(fb, g.B, g.sb_n, g.sb_k, g.b_idx, p.b_mode, n0, kBeg, g.N, kEnd, tid);
  }
  for (int k0 = kBeg; k0 < kEnd; k0 += BK) {
    store_frag(fa, As, p.a_mode, tid);
    store_frag(fb, Bs, p.b_mode, tid);
    __syncthreads();
    if (k0 + BK < kEnd) {
      load_frag<true>(fa, g.A, g.sa_m, g.sa_k, g.a_idx, p.a_mode, m0, k0 + BK, g.M, kEnd, tid);
      load_frag<false>(fb, g.B, g.sb_n, g.sb_k, g.b_idx, p.b_mode, n0, k0 + BK, g.N, kEnd, tid);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  const bool first_split = blockIdx.z == 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
    long long rc = g.c_idx ? (long long)g.c_idx[m] : (long long)m;
    float* crow = g.C + rc * g.ldc;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = g.alpha * acc[i][j];
      if (g.bias && first_split) v += g.bias[n];
      if (!g.accumulate)
        crow[n] = v;
      else if (gridDim.z > 1)
        atomicAdd(crow + n, v);
      else
        crow[n] += v;
    }
  }
}

// Single-shot variant for short K extents (k_chunk <= 256): the whole K range of the A and B tiles is brought into
// shared memory with ALL global loads in flight at once (one latency exposure, one barrier), then the 64x64 tile is
// computed without further synchronisation.  The dense layers of this path (K = d ~ 100-256, or a split-K chunk of a
// weight-gradient GEMM) are latency-bound in the pipelined kernel above: ~1 us per 16-deep k-iteration.
constexpr int SS_MAXK = 256;
constexpr int SS_BATCH = 8;          // k-tiles (of 16) loaded per register batch

__global__ void __launch_bounds__(NT) sgemm_singleshot_kernel(const GemmParams p) {
  SRK_PDL();
  extern __shared__ __align__(16) float ss_smem[];
  const GemmArgs& g = p.g;
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kBeg = blockIdx.z * p.k_chunk;
  const int kEnd = min(g.K, kBeg + p.k_chunk);
  const int kc = ((kEnd - kBeg) + BK - 1) / BK * BK;              // rows actually used (multiple of 16)
  float (*As)[BM + PAD] = reinterpret_cast<float (*)[BM + PAD]>(ss_smem);
  float (*Bs)[BN + PAD] = reinterpret_cast<float (*)[BN + PAD]>(ss_smem + (size_t)p.k_chunk_pad * (BM + PAD));
  const int ty = tid >> 4, tx = tid & 15;

  for (int kb = 0; kb < kc; kb += BK * SS_BATCH) {
    Frag fa[SS_BATCH], fb[SS_BATCH];
#pragma unroll
    for (int t = 0; t < SS_BATCH; ++t) {
      const int k0 = kBeg + kb + t * BK;
      if (kb + t * BK < kc) {
        load_frag<true>(fa[t], g.A, g.sa_m, g.sa_k, g.a_idx, p.a_mode, m0, k0, g.M, kEnd, tid);
        load_frag<false>(fb[t], g.B, g.sb_n, g.sb_k, g.b_idx, p.b_mode, n0, k0, g.N, kEnd, tid);
      }
    }
#pragma unroll
    for (int t = 0; t < SS_BATCH; ++t) {
      if (kb + t * BK < kc) {
        store_frag(fa[t], As + kb + t * BK, p.a_mode, tid);
        store_frag(fb[t], Bs + kb + t * BK, p.b_mode, tid);
      }
    }
  }
  __syncthreads();

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 8
  for (int k = 0; k < kc; ++k) {
    float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
    float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
    float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }

  const bool first_split = blockIdx.z == 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
    long long rc = g.c_idx ? (long long)g.c_idx[m] : (long long)m;
    float* crow = g.C + rc * g.ldc;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = g.alpha * acc[i][j];
      if (g.bias && first_split) v += g.bias[n];
      if (!g.accumulate)
        crow[n] = v;
      else if (gridDim.z > 1)
        atomicAdd(crow + n, v);
      else
        crow[n] += v;
    }
  }
}

// ---- small-K kernel --------------------------------------------------------------------------------------------
// The dense layers of this path have K = d ~ 100 (or a split-K chunk of that size) and a few thousand rows: in the
// pipelined kernel above they spend their time in ~6 serial k-iterations of ~1 us each (global-load latency that one
// resident tile per CTA cannot hide).  This variant brings the WHOLE K extent (<= 160) of a 32 x 32 tile into static
// shared memory with all global loads in flight at once (one latency exposure, one barrier), then computes.  4x more
// CTAs than with 64 x 64 tiles, which these tiny problems need anyway; static smem (46 KB) so that no shared-memory
// carve-out change is triggered between launches.
constexpr int SB = 32, SKMAX = 160, SPAD = 4;

template <bool kIdxOnOuter>
__device__ __forceinline__ void small_load(float (*S)[SB + SPAD], const float* __restrict__ P, long long so, long long sk,
                                           const int* __restrict__ idx, int mode, int o0, int kBeg, int kEnd, int O, int tid) {
  const int kc = kEnd - kBeg;
  const int kc4 = (kc + 3) >> 2;
  if (mode == LM_KVEC || mode == LM_KSCAL) {
    // K contiguous: one float4 = 4 consecutive k of one row
    float4 reg[5];
    const int total = SB * kc4;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int e = tid + i * NT;
      reg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < total) {
        const int o = o0 + e / kc4, k = kBeg + (e % kc4) * 4;
        if (o < O) {
          const long long ro = (kIdxOnOuter && idx) ? (long long)idx[o] : (long long)o;
          const float* p = P + ro * so;
          if (mode == LM_KVEC && k + 3 < kEnd && !(!kIdxOnOuter && idx)) {
            reg[i] = *reinterpret_cast<const float4*>(p + k);
          } else {
            float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (k + j < kEnd) {
                const long long rk = (!kIdxOnOuter && idx) ? (long long)idx[k + j] : (long long)(k + j);
                t[j] = p[rk * sk];
              }
            reg[i] = make_float4(t[0], t[1], t[2], t[3]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int e = tid + i * NT;
      if (e < total) {
        const int o = e / kc4, k = (e % kc4) * 4;
        S[k][o] = reg[i].x; S[k + 1][o] = reg[i].y; S[k + 2][o] = reg[i].z; S[k + 3][o] = reg[i].w;
      }
    }
  } else {
    // outer dimension contiguous (or generic strides): one float4 = 4 consecutive outer indices of one k
    float4 reg[5];
    const int total = kc * (SB / 4);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int e = tid + i * NT;
      reg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < total) {
        const int k = kBeg + e / (SB / 4), o = o0 + (e % (SB / 4)) * 4;
        const long long rk = (!kIdxOnOuter && idx) ? (long long)idx[k] : (long long)k;
        const float* p = P + rk * sk;
        if (mode == LM_MVEC && o + 3 < O && !(kIdxOnOuter && idx)) {
          reg[i] = *reinterpret_cast<const float4*>(p + o);
        } else {
          float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (o + j < O) {
              const long long ro = (kIdxOnOuter && idx) ? (long long)idx[o + j] : (long long)(o + j);
              t[j] = p[ro * so];
            }
          reg[i] = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int e = tid + i * NT;
      if (e < total) {
        const int k = e / (SB / 4), o = (e % (SB / 4)) * 4;
        *reinterpret_cast<float4*>(&S[k][o]) = reg[i];
      }
    }
  }
}

__global__ void __launch_bounds__(NT) sgemm_small_kernel(const GemmParams p) {
  SRK_PDL();
  __shared__ __align__(16) float As[SKMAX + 4][SB + SPAD];
  __shared__ __align__(16) float Bs[SKMAX + 4][SB + SPAD];
  const GemmArgs& g = p.g;
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * SB, n0 = blockIdx.x * SB;
  const int kBeg = blockIdx.z * p.k_chunk;
  const int kEnd = min(g.K, kBeg + p.k_chunk);
  const int kc = kEnd - kBeg;
  // generic-stride operands go through the scalar branch of the "outer contiguous" loader
  small_load<true>(As, g.A, g.sa_m, g.sa_k, g.a_idx, p.a_mode == LM_SCALAR ? LM_MSCAL : p.a_mode, m0, kBeg, kEnd, g.M, tid);
  small_load<false>(Bs, g.B, g.sb_n, g.sb_k, g.b_idx, p.b_mode == LM_SCALAR ? LM_MSCAL : p.b_mode, n0, kBeg, kEnd, g.N, tid);
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;
  float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
#pragma unroll 8
  for (int k = 0; k < kc; ++k) {
    const float2 a = *reinterpret_cast<const float2*>(&As[k][ty * 2]);
    const float2 b = *reinterpret_cast<const float2*>(&Bs[k][tx * 2]);
    a00 = fmaf(a.x, b.x, a00); a01 = fmaf(a.x, b.y, a01);
    a10 = fmaf(a.y, b.x, a10); a11 = fmaf(a.y, b.y, a11);
  }
  const float acc[2][2] = {{a00, a01}, {a10, a11}};
  const bool first_split = blockIdx.z == 0;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int m = m0 + ty * 2 + i;
    if (m >= g.M) continue;
    const long long rc = g.c_idx ? (long long)g.c_idx[m] : (long long)m;
    float* crow = g.C + rc * g.ldc;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int n = n0 + tx * 2 + j;
      if (n >= g.N) continue;
      float v = g.alpha * acc[i][j];
      if (g.bias && first_split) v += g.bias[n];
      if (!g.accumulate)
        crow[n] = v;
      else if (gridDim.z > 1)
        atomicAdd(crow + n, v);
      else
        crow[n] += v;
    }
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

// The dense layers of this path are small (N ~ 2k rows, d ~ 100-256): with one 64x64 tile per CTA most of them
// launch far fewer CTAs than the 148 SMs and are bound by the latency of their serial K loop.  Split K until about
// two CTAs per SM are in flight, never below 2 k-iterations (32) per CTA.
int srk_pick_split_k(int M, int N, int K) {
  long long tiles = (long long)srk_cdiv(M, BM) * srk_cdiv(N, BN);
  const bool small = 2.0 * M * N * K < 1.0e8;           // handled by the small-K kernel (chunks of <= 160)
  if (small && K <= SKMAX) return 1;
  if (tiles >= 296 || K < 128) return 1;
  long long want = (296 + tiles - 1) / tiles;
  long long fit = small ? (K + SKMAX - 1) / SKMAX : 1;
  if (want < fit) want = fit;
  long long cap = K / 32;
  long long s = want < cap ? want : cap;
  if (s > 128) s = 128;
  return (int)(s < 1 ? 1 : s);
}

int srk_gemm_launch(const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return SRK_OK;
  SRK_REQUIRE(g.split_k >= 1, "gemm: split_k must be >= 1");
  SRK_REQUIRE(g.split_k == 1 || g.accumulate, "gemm: split-K needs accumulate mode");
  if (!g.accumulate && !g.c_idx && g.K > SKMAX && (g.K >= 256 || 2.0 * g.M * g.N * g.K < 1.0e8)) {
    // overwrite mode with a long K loop and few tiles: zero C, then run the split-K accumulate path
    int s = srk_pick_split_k(g.M, g.N, g.K);
    if (s > 1) {
      SRK_TRY(srk_zero2d_async(g.C, g.ldc, g.M, g.N, st));
      GemmArgs h = g;
      h.accumulate = 1;
      h.split_k = s;
      return srk_gemm_launch(h, st);
    }
  }
  SRK_REQUIRE(srk_cdiv(g.M, BM) <= 65535, "gemm: M too large for grid.y");
  if (g.K <= 0) {
    SRK_REQUIRE(g.accumulate, "gemm: K == 0 with overwrite mode is not supported");
    return SRK_OK;
  }
  GemmParams p;
  p.g = g;
  // operand A: outer = m
  if (g.sa_k == 1)
    p.a_mode = (g.sa_m % 4 == 0 && aligned16(g.A)) ? LM_KVEC : LM_KSCAL;
  else if (g.sa_m == 1)
    p.a_mode = (g.sa_k % 4 == 0 && aligned16(g.A) && !g.a_idx) ? LM_MVEC : LM_MSCAL;
  else
    p.a_mode = LM_SCALAR;
  // operand B: outer = n
  if (g.sb_k == 1)
    p.b_mode = (g.sb_n % 4 == 0 && aligned16(g.B) && !g.b_idx) ? LM_KVEC : LM_KSCAL;
  else if (g.sb_n == 1)
    p.b_mode = (g.sb_k % 4 == 0 && aligned16(g.B)) ? LM_MVEC : LM_MSCAL;
  else
    p.b_mode = LM_SCALAR;
  int S = g.split_k;
  int ktiles = srk_cdiv(g.K, BK);
  int per = srk_cdiv(ktiles, S);
  p.k_chunk = per * BK;
  S = srk_cdiv(g.K, p.k_chunk);
  static const bool no_small = getenv("SESSREC_SGEMM_NO_SMALL") && getenv("SESSREC_SGEMM_NO_SMALL")[0] == '1';
  // 32 x 32 tiles lose to 64 x 64 once the problem is big enough to fill the machine (measured: the 261 MFLOP GAT
  // projections run 19 us pipelined vs 30 us here), so the small-K kernel only takes the small problems
  const double flops = 2.0 * g.M * g.N * g.K;
  if (!no_small && p.k_chunk <= SKMAX && flops < 1.0e8 && srk_cdiv(g.M, SB) <= 65535) {
    dim3 sgrid(srk_cdiv(g.N, SB), srk_cdiv(g.M, SB), S);
    srk_launch(sgemm_small_kernel, sgrid, NT, 0, st, p);
    SRK_LAUNCH_CHECK();
    return SRK_OK;
  }
  dim3 grid(srk_cdiv(g.N, BN), srk_cdiv(g.M, BM), S);
  // Measured on B200 (profiles/r1g): every launch that asks for > 48 KB of dynamic shared memory pays ~5 us more than
  // the static-smem pipelined kernel in this stream of tiny kernels, which eats the gain; opt-in only.
  static const bool use_ss = getenv("SESSREC_SGEMM_SINGLESHOT") && getenv("SESSREC_SGEMM_SINGLESHOT")[0] == '1';
  if (use_ss && p.k_chunk <= SS_MAXK) {
    p.k_chunk_pad = p.k_chunk;                      // multiple of 16 by construction
    const size_t smem = sizeof(float) * 2 * (size_t)p.k_chunk_pad * (BM + PAD);
    static bool attr_set = false;
    if (!attr_set) {
      SRK_CUDA(cudaFuncSetAttribute(sgemm_singleshot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(sizeof(float) * 2 * SS_MAXK * (BM + PAD))));
      attr_set = true;
    }
    srk_launch(sgemm_singleshot_kernel, grid, NT, smem, st, p);
  } else {
    srk_launch(sgemm_kernel, grid, NT, 0, st, p);
  }
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_gemm(int M, int N, int K, const float* A, long long sa_m, long long sa_k, const float* B,
                        long long sb_k, long long sb_n, float* C, long long ldc, const int* a_idx, const int* b_idx,
                        const int* c_idx, const float* bias, float alpha, int accumulate, int split_k, void* stream) {
  GemmArgs g;
  g.M = M; g.N = N; g.K = K;
  g.A = A; g.sa_m = sa_m; g.sa_k = sa_k;
  g.B = B; g.sb_k = sb_k; g.sb_n = sb_n;
  g.C = C; g.ldc = ldc;
  g.a_idx = a_idx; g.b_idx = b_idx; g.c_idx = c_idx;
  g.bias = bias; g.alpha = alpha; g.accumulate = accumulate;
  g.split_k = split_k > 0 ? split_k : (accumulate ? srk_pick_split_k(M, N, K) : 1);
  return srk_gemm_launch(g, (cudaStream_t)stream);
}
