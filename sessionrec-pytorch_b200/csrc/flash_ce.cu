// Fused catalog scoring + cross-entropy head on tcgen05 tensor cores ("flash CE"): the (B, V) logit matrix of
//   logits = scale * shat Ehat^T ; loss = mean_b (logsumexp_v logits[b, :] - logits[b, y_b])
// (srgnn.py:145-147, niser.py:149-156, msgifsr.py:276-309 + utils/train.py:99 and their autograd) is never written to
// HBM.  The forward kernel keeps only per-row soft-max statistics; the backward kernel RECOMPUTES each 128 x 128
// logit tile on the tensor cores, turns it into dZ = coef * (softmax - onehot) in registers, stores dZ as a bf16
// hi/lo pair in shared memory and feeds it straight back to the tensor cores for both gradient products:
//   dS[b, :]  += dZ[b, v] Ehat[v, :]      (accumulated in TMEM over the CTA's catalog range)
//   dE[v, :]   = dZ[b, v]^T shat[b, :]    (one 128-session partial per CTA; partials are summed by catalog_prep_bwd)
//
// Arithmetic: every fp32 operand x is split as x = hi + lo with hi = bf16(x), lo = bf16(x - hi) (16 mantissa bits in
// total) and every product is evaluated as hi*hi + hi*lo + lo*hi with fp32 accumulation in TMEM ("bf16 x 3"): relative
// error ~1e-5 per product, inside the 1e-4 parity bar, at twice the tensor rate of the 3xTF32 scheme of umma_gemm.cu
// and with operands that take no more shared memory than one fp32 copy.
//
// Work decomposition: CTA = (128-session tile, contiguous range of 128-row catalog tiles); grid = #session tiles x
// #ranges ~ one CTA per SM.  Roles: warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer, warps 2-9
// = soft-max math on the logit tile (TMEM lane quadrant = warp % 4, column half = (warp - 2) / 4); the backward kernel
// adds warps 10-13, which drain the dE / dS accumulators through a 16 KB staging buffer and TMA tensor stores.
// Shared memory: operand tiles are 128 rows of swizzled 128-byte (64 bf16, SWIZZLE_128B) or 64-byte (32 bf16,
// SWIZZLE_64B, when d is a multiple of 32 but not of 64) column chunks, so the SAME bytes serve as a K-major operand of
// one product and as an MN-major operand of another:
//   S  = shat tile   hi/lo  [128 b x d]   resident         A (K-major) of the logit tile, B (MN-major) of dE
//                                                          (forward: held in TMEM instead, TS-form MMA)
//   E  = Ehat tile   hi/lo  [128 v x d]   1-4 stages       B (K-major) of the logit tile, B (MN-major) of dS
//   D  = dZ tile     hi/lo  [128 b x 128 v]                A (K-major) of dS, A (MN-major) of dE (the math warps keep dZ
//                                                          in registers until the previous tile's products have read D)
// TMEM (512 columns): logit tile x 2 (double buffer) | dS accumulator (d <= 128 columns) | dE accumulator
// (forward: logit tile x 2 | shat hi | shat lo).
// What bounds the kernels (clock-stamp traces, srk_flash_ce_set_trace + scripts/head_trace.py): forward = L2 -> SM
// bandwidth of the catalog tiles (each is pulled by every session-tile CTA); backward = shared-memory operand bandwidth
// of the 66 MMAs per tile (480 KB at 128 B/clk against 3456 cycles of tensor work).
#include <cuda_bf16.h>
#include <stdlib.h>

#include "umma.cuh"

using namespace umma;

namespace {

constexpr int TB = 128;                     // sessions per tile (UMMA M of the logit tile, TMEM lanes)
constexpr int TV = 128;                     // catalog rows per tile
constexpr uint32_t CHUNK = 128 * 128;       // bytes of one 128-row x 128-byte operand chunk (64 bf16 or 32 fp32 per row)
constexpr int EPI_WARPS = 8;
constexpr int EPI_THREADS = 32 * EPI_WARPS;
constexpr int THREADS = 64 + EPI_THREADS;
constexpr int DRAIN_WARPS = 4;                    // backward kernel only: one warp per TMEM lane quadrant drains dE / dS
constexpr int DRAIN_THREADS = 32 * DRAIN_WARPS;
constexpr int THREADS_BWD = THREADS + DRAIN_THREADS;
constexpr uint32_t D_BYTES = 4 * CHUNK;     // dZ hi (2 chunks of 64 columns) + dZ lo
constexpr size_t FCE_MAX_SMEM = 227 * 1024 - 1024;
constexpr int MAX_STAGES = 4;
constexpr uint32_t TM_Z = 0, TM_DS = 256, TM_DE = 384;     // TMEM column offsets (backward)
constexpr uint32_t TM_SA = 256;                            // forward: shat hi at [256, 320), lo at [320, 384) (bf16 pairs)

struct FceParams {
  int B, V, d;
  // Operand (S, E) tiles are stored as column chunks of 128 rows: 64 bf16 per row under SWIZZLE_128B, or 32 bf16 per
  // row under SWIZZLE_64B.  The narrow chunks waste nothing when d is a multiple of 32 but not of 64 (d = 96: 48 KB per
  // hi/lo operand pair instead of 64 KB), which buys a second catalog stage in the backward kernel.
  int nch;                 // chunks per operand row = ceil(d / chunk columns)
  int cw;                  // chunk columns (64 or 32)
  int kpc_log2;            // UMMA K steps (16 columns) per chunk, log2
  uint32_t chunk_bytes;    // 128 rows x row_bytes
  uint32_t row_bytes;      // 128 or 64
  uint64_t okd_hi, omd_hi; // operand descriptor constants: K-major / MN-major
  int ntm, nvr, nvt;       // session tiles, catalog ranges, catalog tiles
  int estages;
  float scale;
  const int* labels;
  float* part;             // fwd: [2 * nvr][B][2] (max, sum exp) per (range, column half)
  float* zlab;             // fwd: [B] label logit
  const float* lse;        // bwd: [B]
  const float* gout;       // bwd: upstream gradient of the mean loss (device scalar) or null
  float* dEpart;           // bwd: [ntm][V][d] partial table gradients, one per session tile
  int de_atomic;           // bwd: dEpart is ONE zero-initialised [V][d] buffer, every session tile adds into it (TMA reduce-add)
  uint32_t idesc_z, idesc_ds, idesc_de;
  int a_tmem;              // fwd: the session operand (A of the logit product) lives in TMEM instead of shared memory
  const uint16_t *Shi, *Slo;   // fwd, a_tmem: bf16 hi / lo of shat in global memory, row pitch lds
  long long lds;
  long long* trace;        // debug: clock64 stamps of CTA 0, [role 0..10][tile < 64][8] (srk_flash_ce_set_trace), else null
  int topk;                // fwd: 0, or K <= TOPK_MAX: write the K largest logits of every (range, half, row) to cand_*
  float* cand_val;         // [2 * nvr][B][K]
  int* cand_idx;
};

// role: 0 = TMA producer, 1 = MMA issuer, 2 + w = epilogue warp w (lane 0)
__device__ __forceinline__ void tr(const FceParams& p, int role, int it, int k) {
  if (p.trace != nullptr && blockIdx.x == 0 && it < 64 && (SRK_ISSUE_MODE == 2 || (threadIdx.x & 31) == 0)) p.trace[(role * 64 + it) * 8 + k] = clock64();
}

// UMMA shared-memory descriptors (SWIZZLE_128B, version 1): constant high part | (address >> 4).  K-major operands step
// 32 bytes inside the 128-byte swizzle row per UMMA K (16 bf16); MN-major operands step 16 rows (2 KB) per UMMA K and find
// the next 64-column chunk at LBO = CHUNK.  The descriptors are built with one shift + one OR per MMA: the issuing
// thread is a single lane and every instruction it spends between two tcgen05.mma shows up as tensor-pipe idle time.
constexpr uint64_t KDESC_HI = (uint64_t(16 >> 4) << 16) | (uint64_t(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
constexpr uint64_t MDESC_HI = (uint64_t(CHUNK >> 4) << 16) | (uint64_t(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
__device__ __forceinline__ uint64_t kdesc(uint32_t addr) { return KDESC_HI | (uint64_t)((addr >> 4) & 0x3FFFu); }
__device__ __forceinline__ uint64_t mdesc(uint32_t addr) { return MDESC_HI | (uint64_t)((addr >> 4) & 0x3FFFu); }

__device__ __forceinline__ uint64_t odesc(uint64_t hi, uint32_t addr) { return hi | (uint64_t)((addr >> 4) & 0x3FFFu); }

// MMA issue: see SRK_ISSUE_MODE in umma.cuh.  In the default mode the wrappers below are plain calls inside an
// `if (elect_one())` block; modes 0 / 1 run the issue loops with all 32 lanes and predicate the instructions.
#if SRK_ISSUE_MODE == 2
#define SRK_ISSUER_BLOCK if (elect_one())
__device__ __forceinline__ bool lane0() { return true; }
#elif SRK_ISSUE_MODE == 1
#define SRK_ISSUER_BLOCK
__device__ __forceinline__ bool lane0() { return elect_one(); }
#else
#define SRK_ISSUER_BLOCK
__device__ __forceinline__ bool lane0() { return (threadIdx.x & 31) == 0; }
#endif
__device__ __forceinline__ void e_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  if (lane0()) umma::umma_bf16(tmem_d, adesc, bdesc, idesc, accum);
}
__device__ __forceinline__ void e_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  if (lane0()) umma::umma_bf16_ts(tmem_d, tmem_a, bdesc, idesc, accum);
}
__device__ __forceinline__ void e_commit(uint64_t* bar) {
  if (lane0()) umma::umma_commit(bar);
}

// acc (+)= A B with both operands split hi/lo: hi*hi + hi*lo + lo*hi
__device__ __forceinline__ void mma3(uint32_t tacc, uint64_t ah, uint64_t al, uint64_t bh, uint64_t bl, uint32_t idesc,
                                     uint32_t accum) {
  e_umma(tacc, ah, bh, idesc, accum);
  e_umma(tacc, ah, bl, idesc, 1u);
  e_umma(tacc, al, bh, idesc, 1u);
}

// logit tile: Z[128 b x 128 v] = S E^T, K = d in steps of 16 (32 bytes inside the swizzled chunk row)
__device__ __forceinline__ void issue_logits(uint32_t tz, uint32_t S, uint32_t E, uint32_t op_bytes, const FceParams& p) {
  const int nks = p.d >> 4;
  const uint64_t sh = odesc(p.okd_hi, S), sl = odesc(p.okd_hi, S + op_bytes), eh = odesc(p.okd_hi, E), el = odesc(p.okd_hi, E + op_bytes);
  const int kmask = (1 << p.kpc_log2) - 1;
#pragma unroll 1
  for (int ks = 0; ks < nks; ++ks) {
    const uint64_t off = (uint64_t)(((uint32_t)(ks >> p.kpc_log2) * p.chunk_bytes + (uint32_t)(ks & kmask) * 32) >> 4);
    mma3(tz, sh + off, sl + off, eh + off, el + off, p.idesc_z, ks ? 1u : 0u);
  }
}

// same product with the A operand (shat hi / lo, K-major bf16 pairs) resident in TMEM: the tensor core then reads only the
// catalog tile from shared memory - with both operands in shared memory the three products per K step run at the
// 128 B/clk shared-memory limit, not at the tensor-pipe rate
__device__ __forceinline__ void issue_logits_ts(uint32_t tz, uint32_t ta, uint32_t E, uint32_t op_bytes, const FceParams& p) {
  const int nks = p.d >> 4;
  const uint64_t eh = odesc(p.okd_hi, E), el = odesc(p.okd_hi, E + op_bytes);
  const int kmask = (1 << p.kpc_log2) - 1;
#pragma unroll 1
  for (int ks = 0; ks < nks; ++ks) {
    const uint64_t off = (uint64_t)(((uint32_t)(ks >> p.kpc_log2) * p.chunk_bytes + (uint32_t)(ks & kmask) * 32) >> 4);
    const uint32_t ah = ta + (uint32_t)ks * 8, al = ah + 64;
    e_umma_ts(tz, ah, eh + off, p.idesc_z, ks ? 1u : 0u);
    e_umma_ts(tz, ah, el + off, p.idesc_z, 1u);
    e_umma_ts(tz, al, eh + off, p.idesc_z, 1u);
  }
}

constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
__device__ __forceinline__ float ex2f(float x) {           // 2^x on the SFU (MUFU.EX2), flush-to-zero
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// label logit of a row whose label column lies in the 32 accumulator columns at `taddr` (rare path: once per row)
__device__ __noinline__ float pick_column(uint32_t taddr, int j) {
  uint32_t r[32];
  tmem_ld32(taddr, r);
  float x = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) x = (k == j) ? __uint_as_float(r[k]) : x;
  return x;
}

// Fused evaluation head (evaluate(), utils/train.py:36-55: `logits.topk(k=20)`): every soft-max thread keeps the K largest
// logits of its (row, column half) over the CTA's catalog range in a small sorted list - one compare per logit against the
// current K-th value, an insertion is rare once the list has warmed up - and writes the list out; fce_topk_merge_kernel
// merges the 2 * nvr lists of a row.  The (B, V) score matrix never exists, in HBM or anywhere else.
constexpr int TOPK_MAX = 32;
struct TopK {
  float v[TOPK_MAX];
  int i[TOPK_MAX];
  float vmin;
  int n;
  __device__ __forceinline__ void init() { n = 0; vmin = -3.0e38f; }
  __device__ __noinline__ void insert(float a, int col, int K) {
    int p = n < K ? n : K - 1;
    while (p > 0 && v[p - 1] < a) {            // strict: among equal values the smaller column (seen first) stays ahead
      v[p] = v[p - 1];
      i[p] = i[p - 1];
      --p;
    }
    v[p] = a;
    i[p] = col;
    if (n < K) ++n;
    vmin = n == K ? v[K - 1] : -3.0e38f;
  }
  // 32 accumulator columns starting at catalog column col0, the first nvalid of them inside the catalog
  __device__ __forceinline__ void scan(const uint32_t* r, int nvalid, int col0, int K) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float a = __uint_as_float(r[j]);
      if (j < nvalid && a > vmin) insert(a, col0 + j, K);
    }
  }
};

struct TileSched {
  int tb, t0, t1;
  __device__ TileSched(const FceParams& p) {
    tb = blockIdx.x % p.ntm;
    const int vr = blockIdx.x / p.ntm;
    t0 = (int)((long long)vr * p.nvt / p.nvr);
    t1 = (int)((long long)(vr + 1) * p.nvt / p.nvr);
  }
};

__device__ __forceinline__ void load_operand(uint8_t* dst, const CUtensorMap* hi, const CUtensorMap* lo, uint64_t* bar,
                                             int nch, uint32_t chunk_bytes, int cw, uint32_t op_bytes, int row0) {
  for (int c = 0; c < nch; ++c) {
    tma_load_2d(dst + c * chunk_bytes, hi, bar, c * cw, row0);
    tma_load_2d(dst + op_bytes + c * chunk_bytes, lo, bar, c * cw, row0);
  }
}

// ---- forward: per-row (max, sum exp) partials + label logit, nothing else leaves the SM ---------------------------------
template <bool TOPK>
__global__ void __launch_bounds__(THREADS, 1)
fce_fwd_kernel(const __grid_constant__ CUtensorMap mSh, const __grid_constant__ CUtensorMap mSl,
               const __grid_constant__ CUtensorMap mEh, const __grid_constant__ CUtensorMap mEl, const FceParams p) {
  SRK_PDL();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_full, e_full[MAX_STAGES], e_empty[MAX_STAGES], z_full[2], z_empty[2];
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t op_bytes = (uint32_t)p.nch * p.chunk_bytes;
  const bool ats = p.a_tmem != 0;
  uint8_t* S = smem;
  uint8_t* E0 = smem + (ats ? 0 : 2 * op_bytes);
  const TileSched ts(p);

  if (threadIdx.x == 0) {
    mbar_init(&s_full, ats ? EPI_WARPS : 1);
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(&e_full[s], 1);
      mbar_init(&e_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&z_full[s], 1);
      mbar_init(&z_empty[s], EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_tc_before();
  __syncthreads();
  fence_tc_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (issue_lane()) {
      if (!ats) {
        mbar_expect_tx(&s_full, 2 * op_bytes);
        load_operand(S, &mSh, &mSl, &s_full, p.nch, p.chunk_bytes, p.cw, op_bytes, ts.tb * TB);
      }
      int it = 0;
      for (int t = ts.t0; t < ts.t1; ++t, ++it) {
        const int s = it % p.estages;
        mbar_wait(&e_empty[s], ((uint32_t)(it / p.estages) & 1u) ^ 1u);
        tr(p, 0, it, 0);
        mbar_expect_tx(&e_full[s], 2 * op_bytes);
        load_operand(E0 + (size_t)s * 2 * op_bytes, &mEh, &mEl, &e_full[s], p.nch, p.chunk_bytes, p.cw, op_bytes, t * TV);
        tr(p, 0, it, 1);
      }
    }
  } else if (warp == 1) {
    SRK_ISSUER_BLOCK {
      mbar_wait(&s_full, 0);
      fence_tc_after();
      int it = 0;
      for (int t = ts.t0; t < ts.t1; ++t, ++it) {
        const int s = it % p.estages, zb = it & 1;
        mbar_wait(&e_full[s], (uint32_t)(it / p.estages) & 1u);
        tr(p, 1, it, 0);
        mbar_wait(&z_empty[zb], ((uint32_t)(it >> 1) & 1u) ^ 1u);
        tr(p, 1, it, 1);
        fence_tc_after();
        if (ats) issue_logits_ts(tmem_base + TM_Z + (uint32_t)zb * TV, tmem_base + TM_SA, smem_u32(E0 + (size_t)s * 2 * op_bytes), op_bytes, p);
        else issue_logits(tmem_base + TM_Z + (uint32_t)zb * TV, smem_u32(S), smem_u32(E0 + (size_t)s * 2 * op_bytes), op_bytes, p);
        e_commit(&e_empty[s]);
        e_commit(&z_full[zb]);
        tr(p, 1, it, 2);
      }
    }
  } else {
    // Per-row online soft-max in base 2: t = acc * (scale * log2 e), running maximum m2 and s = sum 2^(t - m2) in four
    // independent partial sums (the chain of dependent FADDs would otherwise leave the two warps per scheduler idle).
    // Per element: FMNMX, FFMA, MUFU.EX2, FADD.
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int b = ts.tb * TB + q * 32 + lane;
    const int lab = (b < p.B && p.labels != nullptr) ? p.labels[b] : -1;
    const float c2 = p.scale * LOG2E;
    const uint32_t lanebits = (uint32_t)(q * 32) << 16;
    if (ats) {
      // this thread's row of shat (hi for the warps of column half 0, lo for half 1) -> TMEM, 8 packed bf16 pairs at a time
      const uint16_t* src = (half == 0 ? p.Shi : p.Slo) + (long long)b * p.lds;
      const uint32_t dst = tmem_base + lanebits + TM_SA + (uint32_t)half * 64;
      for (int w = 0; w < (p.d >> 1); w += 8) {
        uint4 x0 = make_uint4(0u, 0u, 0u, 0u), x1 = x0;
        if (b < p.B) {
          x0 = *reinterpret_cast<const uint4*>(src + 2 * w);
          x1 = *reinterpret_cast<const uint4*>(src + 2 * w + 8);
        }
        const uint32_t v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        tmem_st8(dst + (uint32_t)w, v);
      }
      tmem_wait_st();
      fence_tc_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_full);
    }
    float m2 = -3.0e38f, s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, zl = 0.f;
    bool has = false;
    TopK tk;
    if (TOPK) tk.init();
    int it = 0;
    for (int t = ts.t0; t < ts.t1; ++t, ++it) {
      const int zb = it & 1;
      mbar_wait(&z_full[zb], (uint32_t)(it >> 1) & 1u);
      if (warp == 2 && lane == 0) tr(p, 2, it, 0);
      fence_tc_after();
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c0 = half * 64 + cc * 32;
        const uint32_t taddr = tmem_base + lanebits + TM_Z + (uint32_t)(zb * TV + c0);
        uint32_t r[32];
        tmem_ld32(taddr, r);
        const int v0 = t * TV + c0;
        const int nvalid = min(32, p.V - v0);
        if (nvalid == 32) {
          float a0 = __uint_as_float(r[0]), a1 = __uint_as_float(r[1]), a2 = __uint_as_float(r[2]), a3 = __uint_as_float(r[3]);
#pragma unroll
          for (int j = 4; j < 32; j += 4) {
            a0 = fmaxf(a0, __uint_as_float(r[j]));
            a1 = fmaxf(a1, __uint_as_float(r[j + 1]));
            a2 = fmaxf(a2, __uint_as_float(r[j + 2]));
            a3 = fmaxf(a3, __uint_as_float(r[j + 3]));
          }
          const float cm2 = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3)) * c2;
          if (cm2 > m2) {
            const float f = ex2f(m2 - cm2);
            s0 *= f; s1 *= f; s2 *= f; s3 *= f;
            m2 = cm2;
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            s0 += ex2f(fmaf(__uint_as_float(r[j]), c2, -m2));
            s1 += ex2f(fmaf(__uint_as_float(r[j + 1]), c2, -m2));
            s2 += ex2f(fmaf(__uint_as_float(r[j + 2]), c2, -m2));
            s3 += ex2f(fmaf(__uint_as_float(r[j + 3]), c2, -m2));
          }
        } else if (nvalid > 0) {                 // ragged last catalog tile
          float cm = -3.0e38f;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nvalid) cm = fmaxf(cm, __uint_as_float(r[j]));
          const float cm2 = cm * c2;
          if (cm2 > m2) {
            const float f = ex2f(m2 - cm2);
            s0 *= f; s1 *= f; s2 *= f; s3 *= f;
            m2 = cm2;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nvalid) s0 += ex2f(fmaf(__uint_as_float(r[j]), c2, -m2));
        }
        if (TOPK && nvalid > 0) tk.scan(r, nvalid, v0, p.topk);
        const bool mine = lab >= v0 && lab < v0 + nvalid;
        if (__any_sync(SRK_FULL, mine)) {          // tcgen05.ld is warp-collective: every lane re-reads, owners keep
          const float x = p.scale * pick_column(taddr, lab - v0);
          if (mine) {
            zl = x;
            has = true;
          }
        }
      }
      fence_tc_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&z_empty[zb]);
      if (warp == 2 && lane == 0) tr(p, 2, it, 1);
    }
    if (b < p.B) {
      const int vr = blockIdx.x / p.ntm;
      float* pp = p.part + ((long long)(vr * 2 + half) * p.B + b) * 2;
      pp[0] = m2 * LN2;                          // natural-log units: max logit, sum exp(logit - max)
      pp[1] = (s0 + s1) + (s2 + s3);
      if (has) p.zlab[b] = zl;
      if (TOPK) {
        float* cv = p.cand_val + ((long long)(vr * 2 + half) * p.B + b) * p.topk;
        int* ci = p.cand_idx + ((long long)(vr * 2 + half) * p.B + b) * p.topk;
        for (int k = 0; k < p.topk; ++k) {
          cv[k] = k < tk.n ? tk.v[k] : -3.0e38f;
          ci[k] = k < tk.n ? tk.i[k] : 0x7fffffff;
        }
      }
    }
  }
  fence_tc_before();
  __syncthreads();
  if (warp == 1) {
    fence_tc_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// lse[b] = log sum over the partials; nll[b] = lse[b] - zlab[b].  Warp per row.
__global__ void __launch_bounds__(256) fce_finalize_kernel(const float* __restrict__ part, const float* __restrict__ zlab,
                                                           const int* __restrict__ labels, int B, int V, int nparts,
                                                           float* __restrict__ lse, float* __restrict__ nll) {
  SRK_PDL();
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  float mx = -3.0e38f, s = 0.f;
  for (int t = lane; t < nparts; t += 32) {
    const float pm = part[((long long)t * B + b) * 2], ps = part[((long long)t * B + b) * 2 + 1];
    const float nm = fmaxf(mx, pm);
    s = s * expf(mx - nm) + ps * expf(pm - nm);
    mx = nm;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(SRK_FULL, mx, o), os = __shfl_xor_sync(SRK_FULL, s, o);
    const float nm = fmaxf(mx, om);
    s = s * expf(mx - nm) + os * expf(om - nm);
    mx = nm;
  }
  if (lane == 0) {
    const float l = mx + logf(s);
    lse[b] = l;
    if (nll) {                          // label outside [0, V): its column lives on another rank (catalog sharding)
      const int lab = labels[b];
      nll[b] = (lab >= 0 && lab < V) ? l - zlab[b] : 0.f;
    }
  }
}

__global__ void __launch_bounds__(256) sum_parts_kernel(const float4* __restrict__ parts, long long stride4, int nparts, long long n4,
                                                        float4* __restrict__ out, int accumulate) {
  SRK_PDL();
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += step) {
    float4 a = accumulate ? out[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int pt = 0; pt < nparts; ++pt) {
      const float4 x = parts[pt * stride4 + i];
      a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
    }
    out[i] = a;
  }
}

// ---- backward ---------------------------------------------------------------------------------------------------------
// 32 consecutive dZ values -> 16 + 16 packed bf16 pairs.  hi = the upper 16 bits of the fp32 value (truncation: one PRMT
// packs two of them), lo = bf16_rn(x - hi) (x - hi is exact).
__device__ __forceinline__ void pack_dz(const float* dz, uint32_t* h, uint32_t* l) {
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const uint32_t x0 = __float_as_uint(dz[2 * k]), x1 = __float_as_uint(dz[2 * k + 1]);
    h[k] = __byte_perm(x0, x1, 0x7632);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(dz[2 * k] - __uint_as_float(x0 & 0xFFFF0000u),
                                                    dz[2 * k + 1] - __uint_as_float(x1 & 0xFFFF0000u));
    l[k] = *reinterpret_cast<const uint32_t*>(&ll);
  }
}

// packed dZ of tile row r, columns c0 .. c0 + 31 -> the 128-byte-swizzled rows of the D tile (hi and lo halves)
__device__ __forceinline__ void store_dz(uint8_t* Dhi, uint8_t* Dlo, int r, int c0, const uint32_t* h, const uint32_t* l) {
  const uint32_t base = (uint32_t)(c0 >> 6) * CHUNK + (uint32_t)r * 128;
  const uint32_t u0 = (uint32_t)(c0 & 63) >> 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t pu = ((u0 + j) ^ ((uint32_t)r & 7u)) * 16;
    *reinterpret_cast<uint4*>(Dhi + base + pu) = make_uint4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
    *reinterpret_cast<uint4*>(Dlo + base + pu) = make_uint4(l[4 * j], l[4 * j + 1], l[4 * j + 2], l[4 * j + 3]);
  }
}

// dZ[r, c] -= x for one element already stored by store_dz (the onehot term of the label column: once per row)
__device__ __noinline__ void fix_dz(uint8_t* Dhi, uint8_t* Dlo, int r, int c, float x) {
  const uint32_t off = (uint32_t)(c >> 6) * CHUNK + (uint32_t)r * 128 + ((((uint32_t)(c & 63) >> 3) ^ ((uint32_t)r & 7u)) * 16) +
                       (uint32_t)(c & 7) * 2;
  uint16_t* ph = reinterpret_cast<uint16_t*>(Dhi + off);
  uint16_t* pl = reinterpret_cast<uint16_t*>(Dlo + off);
  const float g = __uint_as_float((uint32_t)*ph << 16) + __uint_as_float((uint32_t)*pl << 16) - x;
  const uint32_t gb = __float_as_uint(g);
  *ph = (uint16_t)(gb >> 16);
  const __nv_bfloat16 lo = __float2bfloat16_rn(g - __uint_as_float(gb & 0xFFFF0000u));
  *pl = *reinterpret_cast<const uint16_t*>(&lo);
}

// 32 fp32 accumulator columns of tile row r -> swizzled fp32 staging chunk (128 rows x 128 B) for a TMA store
__device__ __forceinline__ void stage_row(uint8_t* chunk, int r, const uint32_t* v) {
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const uint32_t pu = ((uint32_t)u ^ ((uint32_t)r & 7u)) * 16;
    *reinterpret_cast<uint4*>(chunk + (uint32_t)r * 128 + pu) = make_uint4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
  }
}

__global__ void __launch_bounds__(THREADS_BWD, 1)
fce_bwd_kernel(const __grid_constant__ CUtensorMap mSh, const __grid_constant__ CUtensorMap mSl,
               const __grid_constant__ CUtensorMap mEh, const __grid_constant__ CUtensorMap mEl,
               const __grid_constant__ CUtensorMap mdE, const __grid_constant__ CUtensorMap mdS, const FceParams p) {
  SRK_PDL();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_full, e_full[MAX_STAGES], e_empty[MAX_STAGES], z_full[2], z_empty[2], d_full, d_empty, de_free;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t op_bytes = (uint32_t)p.nch * p.chunk_bytes;
  uint8_t* S = smem;
  uint8_t* Dt = smem + 2 * op_bytes;                // dZ hi (2 chunks) | dZ lo (2 chunks); fp32 staging of the final dS drain
  uint8_t* E0 = Dt + D_BYTES;
  uint8_t* Stg = E0 + (size_t)p.estages * 2 * op_bytes;     // one 128-row x 32-column fp32 chunk: staging of the dE stores
  const TileSched ts(p);
  const int ntiles = ts.t1 - ts.t0;

  if (threadIdx.x == 0) {
    mbar_init(&s_full, 1);
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(&e_full[s], 1);
      mbar_init(&e_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&z_full[s], 1);
      mbar_init(&z_empty[s], EPI_WARPS);
    }
    mbar_init(&d_full, EPI_WARPS);
    mbar_init(&d_empty, 1);
    mbar_init(&de_free, DRAIN_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_tc_before();
  __syncthreads();
  fence_tc_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (issue_lane()) {
      mbar_expect_tx(&s_full, 2 * op_bytes);
      load_operand(S, &mSh, &mSl, &s_full, p.nch, p.chunk_bytes, p.cw, op_bytes, ts.tb * TB);
      for (int it = 0; it < ntiles; ++it) {
        const int s = it % p.estages;
        mbar_wait(&e_empty[s], ((uint32_t)(it / p.estages) & 1u) ^ 1u);
        tr(p, 0, it, 0);
        mbar_expect_tx(&e_full[s], 2 * op_bytes);
        load_operand(E0 + (size_t)s * 2 * op_bytes, &mEh, &mEl, &e_full[s], p.nch, p.chunk_bytes, p.cw, op_bytes, (ts.t0 + it) * TV);
        tr(p, 0, it, 1);
      }
    }
  } else if (warp == 1) {
    SRK_ISSUER_BLOCK {
      const uint32_t Sa = smem_u32(S), Da = smem_u32(Dt);
      const uint32_t tds = tmem_base + TM_DS, tde = tmem_base + TM_DE;
      auto logits = [&](int it) {
        const int s = it % p.estages, zb = it & 1;
        mbar_wait(&e_full[s], (uint32_t)(it / p.estages) & 1u);
        tr(p, 1, it, 0);
        mbar_wait(&z_empty[zb], ((uint32_t)(it >> 1) & 1u) ^ 1u);
        fence_tc_after();
        issue_logits(tmem_base + TM_Z + (uint32_t)zb * TV, Sa, smem_u32(E0 + (size_t)s * 2 * op_bytes), op_bytes, p);
        e_commit(&z_full[zb]);
        tr(p, 1, it, 1);
      };
      mbar_wait(&s_full, 0);
      logits(0);
      for (int it = 0; it < ntiles; ++it) {
        // with more than one catalog stage the next logit tile does not depend on this tile's gradient products
        if (p.estages > 1 && it + 1 < ntiles) logits(it + 1);
        const int s = it % p.estages;
        const uint32_t Ea = smem_u32(E0 + (size_t)s * 2 * op_bytes);
        mbar_wait(&d_full, (uint32_t)it & 1u);
        tr(p, 1, it, 2);
        fence_tc_after();
        // dS[128 b x d] += dZ[128 b x 128 v] E[128 v x d]: A = D K-major, B = E MN-major, K = v in steps of 16 rows
        {
          const uint64_t ah = kdesc(Da), al = kdesc(Da + 2 * CHUNK), bh = odesc(p.omd_hi, Ea), bl = odesc(p.omd_hi, Ea + op_bytes);
          const uint32_t kstep = 16 * p.row_bytes;          // 16 operand rows per UMMA K
#pragma unroll
          for (int ks = 0; ks < TV / 16; ++ks) {
            const uint64_t aoff = (uint64_t)(((uint32_t)(ks >> 2) * CHUNK + (uint32_t)(ks & 3) * 32) >> 4);
            const uint64_t boff = (uint64_t)((uint32_t)ks * kstep >> 4);
            mma3(tds, ah + aoff, al + aoff, bh + boff, bl + boff, p.idesc_ds, (it | ks) ? 1u : 0u);
          }
        }
        e_commit(&e_empty[s]);
        // the drain warps have read the previous tile's dE accumulator out of TMEM
        if (it > 0) {
          mbar_wait(&de_free, (uint32_t)(it - 1) & 1u);
          fence_tc_after();
        }
        // dE[128 v x d] = dZ^T[128 v x 128 b] S[128 b x d]: A = D MN-major, B = S MN-major, K = b in steps of 16 rows
        {
          const uint64_t ah = mdesc(Da), al = mdesc(Da + 2 * CHUNK), bh = odesc(p.omd_hi, Sa), bl = odesc(p.omd_hi, Sa + op_bytes);
          const uint32_t kstep = 16 * p.row_bytes;
#pragma unroll
          for (int ks = 0; ks < TB / 16; ++ks) {
            const uint64_t aoff = (uint64_t)((uint32_t)ks * 2048 >> 4), boff = (uint64_t)((uint32_t)ks * kstep >> 4);
            mma3(tde, ah + aoff, al + aoff, bh + boff, bl + boff, p.idesc_de, ks ? 1u : 0u);
          }
        }
        e_commit(&d_empty);
        tr(p, 1, it, 3);
        if (p.estages == 1 && it + 1 < ntiles) logits(it + 1);
      }
    }
  } else if (warp < 2 + EPI_WARPS) {
    // ===== math warps (2..9): logit tile -> dZ = coef * (softmax - onehot) -> bf16 hi/lo pairs -> D tile =====
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int r = q * 32 + lane;                     // row of the tile owned by this thread (= TMEM lane)
    const int b = ts.tb * TB + r;
    const bool bvalid = b < p.B;
    const int lab = bvalid ? p.labels[b] : -1;
    // softmax = 2^(acc * c2 - lse2); rows beyond B get coef = 0
    const float c2 = p.scale * LOG2E;
    // lse * log2(e) is kept as hi + lo (lo = the rounding residual of the product): a 1e-6 relative error of lse2 would
    // be a systematic 1e-5 relative error of every soft-max value of the row; 2^-lo is folded into the row's coefficient
    const float lse_b = bvalid ? p.lse[b] : 0.f;
    const float lse2 = lse_b * LOG2E;
    const float coef = bvalid ? p.scale * (p.gout ? p.gout[0] : 1.f) / (float)p.B : 0.f;
    const float coef_p = coef * ex2f(-fmaf(lse_b, LOG2E, -lse2));
    const uint32_t lanebits = (uint32_t)(q * 32) << 16;
    uint8_t* Dhi = Dt;
    uint8_t* Dlo = Dt + 2 * CHUNK;
    for (int it = 0; it < ntiles; ++it) {
      const int t = ts.t0 + it, zb = it & 1;
      mbar_wait(&z_full[zb], (uint32_t)(it >> 1) & 1u);
      if (lane == 0) tr(p, warp, it, 0);
      fence_tc_after();
      // dZ of this thread's 64 columns, packed to bf16 hi/lo pairs and HELD IN REGISTERS: the D tile in shared memory is
      // still being read by the previous tile's gradient products while this runs
      uint32_t hw[32], lw[32];
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c0 = half * 64 + cc * 32;
        uint32_t acc[32];
        tmem_ld32(tmem_base + lanebits + TM_Z + (uint32_t)(zb * TV + c0), acc);
        const int v0 = t * TV + c0;
        float dz[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) dz[j] = coef_p * ex2f(fmaf(__uint_as_float(acc[j]), c2, -lse2));
        if (v0 + 32 > p.V) {                      // ragged last catalog tile: columns beyond V carry no gradient
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (v0 + j >= p.V) dz[j] = 0.f;
        }
        pack_dz(dz, hw + 16 * cc, lw + 16 * cc);
      }
      fence_tc_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&z_empty[zb]);
      if (lane == 0) tr(p, warp, it, 1);
      // previous tile's gradient products complete: the D tile is free
      if (it > 0) mbar_wait(&d_empty, (uint32_t)(it - 1) & 1u);
      if (lane == 0) tr(p, warp, it, 2);
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c0 = half * 64 + cc * 32;
        store_dz(Dhi, Dlo, r, c0, hw + 16 * cc, lw + 16 * cc);
        const int v0 = t * TV + c0;
        if (lab >= v0 && lab < v0 + 32) fix_dz(Dhi, Dlo, r, lab - t * TV, coef);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&d_full);
      if (lane == 0) tr(p, warp, it, 3);
    }
  } else {
    // ===== drain warps (10..13, TMEM lane quadrant = warp % 4): accumulators -> registers -> 32-column chunks through
    // ONE 16 KB swizzled staging buffer -> TMA tensor store (dE partial of this session tile) / TMA reduce-add (dS).
    // Scattered per-lane global stores would occupy the load/store unit for ~2000 cycles per tile and delay every
    // mbarrier / shared-memory operation of the CTA (measured); the bulk-copy engine reads shared memory on its own. =====
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lanebits = (uint32_t)(q * 32) << 16;
    const int nc32 = (p.d + 31) >> 5;                // 32-column fp32 chunks of the dS / dE accumulators (<= 4)
    const bool elected = (warp == 2 + EPI_WARPS && lane == 0);
    // stage one chunk and hand it to the bulk-copy engine
    auto put_chunk = [&](const uint32_t* v, int cc, int row0, bool reduce) {
      if (elected) tma_wait_group_read0();         // the previous chunk's store has read the staging buffer
      named_bar_sync(2, DRAIN_THREADS);
      stage_row(Stg, r, v);
      fence_async_smem();
      named_bar_sync(2, DRAIN_THREADS);
      if (elected) {
        if (reduce) tma_reduce_add_2d(&mdS, Stg, cc * 32, row0);
        else if (p.de_atomic) tma_reduce_add_3d(&mdE, Stg, cc * 32, row0, 0);
        else tma_store_3d(&mdE, Stg, cc * 32, row0, ts.tb);
        tma_commit_group();
      }
    };
    auto drain = [&](uint32_t tcol, int row0, bool reduce, bool release) {
      uint32_t a0[32], a1[32], a2[32];
      tmem_ld32(tmem_base + lanebits + tcol, a0);
      if (nc32 > 1) tmem_ld32(tmem_base + lanebits + tcol + 32, a1);
      if (nc32 > 2) tmem_ld32(tmem_base + lanebits + tcol + 64, a2);
      if (release && nc32 <= 3) {                  // accumulator is in registers: the next dE product may overwrite it
        fence_tc_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&de_free);
      }
      put_chunk(a0, 0, row0, reduce);
      if (nc32 > 1) put_chunk(a1, 1, row0, reduce);
      if (nc32 > 2) put_chunk(a2, 2, row0, reduce);
      if (nc32 > 3) {
        tmem_ld32(tmem_base + lanebits + tcol + 96, a0);
        if (release) {
          fence_tc_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&de_free);
        }
        put_chunk(a0, 3, row0, reduce);
      }
    };
    for (int it = 0; it < ntiles; ++it) {
      mbar_wait(&d_empty, (uint32_t)it & 1u);      // both gradient products of tile `it` are complete
      fence_tc_after();
      if (lane == 0) tr(p, 2 + EPI_WARPS, it, warp - 2 - EPI_WARPS);
      drain(TM_DE, (ts.t0 + it) * TV, false, true);
    }
    // dS accumulator of this CTA's whole catalog range -> reduce-add into dS[B, d]
    drain(TM_DS, ts.tb * TB, true, false);
    if (elected) tma_wait_group0();
  }
  fence_tc_before();
  __syncthreads();
  if (warp == 1) {
    fence_tc_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// =========================================================================================================================
// Wide embeddings (d = 256: BASELINE configs[2] SRGNN / Yoochoose1/64 and configs[4] MSGIFSR / Yoochoose1/4).
//
// At d = 256 one 128-row operand tile is 128 KB as a bf16 hi / lo pair, so the layout of the kernels above (session tile + catalog
// tile + dZ tile resident in shared memory) does not fit.  What changes:
//   forward   the catalog tile streams through shared memory in 64-column K chunks (one pipeline stage = hi + lo chunk, 32 KB,
//             6-7 stages) and the logit tile accumulates over the chunks; the session operand lives in TMEM (256 columns) as before.
//   backward  the session tile stays resident (128 KB: K-major A of the logits, MN-major A of dE^T); the catalog tile shrinks to
//             64 rows and streams in 64-column chunks of 16 KB TWICE per tile - once as the K-major B of the logit product, once
//             as the MN-major B of dS (an extra L2 -> SM pass of 64 KB per tile against 4608 cycles of tensor work: 28 B/clk);
//             dE is produced transposed, dE^T[d, 64 v] = S^T dZ in two M = 128 halves, so that every product keeps M = 128 and
//             the accumulator lanes are embedding columns - a warp then stores 32 consecutive floats of one catalog row, a
//             coalesced 128-byte store, and the dE partial needs no staging buffer.
// TMEM (512 columns): logit tile 2 x 64 | dS 256 | dE^T 2 x 64.
constexpr int WTV = 64;                          // catalog rows per backward tile
constexpr uint32_t WECHUNK = 64 * 128;           // bytes of one 64-row x 128-byte catalog chunk (backward)
constexpr int W_MAX_STAGES = 8;
constexpr uint32_t TMW_Z = 0, TMW_DS = 128, TMW_DE = 384;

struct FceWide {
  int B, V, d, nch;        // nch = d / 64 chunks of 64 columns
  int ntm, nvr, nvt;       // session tiles, catalog ranges, catalog tiles (128 rows forward, 64 rows backward)
  int tv;                  // catalog rows per tile
  int estages;
  float scale;
  const int* labels;
  float* part;             // fwd
  float* zlab;             // fwd
  const float* lse;        // bwd
  const float* gout;       // bwd
  float* dEpart;           // bwd: [ntm][V][d]
  const uint16_t *Shi, *Slo;   // fwd: bf16 hi / lo of shat in global memory, row pitch lds (staged into TMEM)
  long long lds;
  uint32_t idesc_z, idesc_ds, idesc_de;
  uint32_t idesc_de2;      // bwd: the dS product with N = d (whole catalog tile resident in the stage ring)
  int topk;                // fwd: see FceParams
  float* cand_val;
  int* cand_idx;
  int dz_tmem;             // bwd: dZ also lives in TMEM (A operand of the dS product); the logit tile is then single-buffered
  int de_atomic;           // bwd: dEpart is ONE zero-initialised [V][d] buffer that every session tile adds into (red.global.add)
};

struct WideSched {
  int tb, t0, t1;
  __device__ WideSched(const FceWide& p) {
    tb = blockIdx.x % p.ntm;
    const int vr = blockIdx.x / p.ntm;
    t0 = (int)((long long)vr * p.nvt / p.nvr);
    t1 = (int)((long long)(vr + 1) * p.nvt / p.nvr);
  }
};

template <bool TOPK>
__global__ void __launch_bounds__(THREADS, 1)
fce_fwd_wide_kernel(const __grid_constant__ CUtensorMap mEh, const __grid_constant__ CUtensorMap mEl, const FceWide p) {
  SRK_PDL();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_full, e_full[W_MAX_STAGES], e_empty[W_MAX_STAGES], z_full[2], z_empty[2];
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t lo_off = (uint32_t)p.d >> 1;          // TMEM columns between shat hi and shat lo (packed bf16 pairs)
  const WideSched ts(p);

  if (threadIdx.x == 0) {
    mbar_init(&s_full, EPI_WARPS);
    for (int s = 0; s < W_MAX_STAGES; ++s) {
      mbar_init(&e_full[s], 1);
      mbar_init(&e_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&z_full[s], 1);
      mbar_init(&z_empty[s], EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_tc_before();
  __syncthreads();
  fence_tc_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (issue_lane()) {
      int u = 0;                                       // pipeline-stage uses so far
      for (int t = ts.t0; t < ts.t1; ++t)
        for (int c = 0; c < p.nch; ++c, ++u) {
          const int s = u % p.estages;
          mbar_wait(&e_empty[s], ((uint32_t)(u / p.estages) & 1u) ^ 1u);
          uint8_t* st = smem + (size_t)s * 2 * CHUNK;
          mbar_expect_tx(&e_full[s], 2 * CHUNK);
          tma_load_2d(st, &mEh, &e_full[s], c * 64, t * TV);
          tma_load_2d(st + CHUNK, &mEl, &e_full[s], c * 64, t * TV);
        }
    }
  } else if (warp == 1) {
    SRK_ISSUER_BLOCK {
      mbar_wait(&s_full, 0);
      fence_tc_after();
      int u = 0, it = 0;
      for (int t = ts.t0; t < ts.t1; ++t, ++it) {
        const int zb = it & 1;
        mbar_wait(&z_empty[zb], ((uint32_t)(it >> 1) & 1u) ^ 1u);
        fence_tc_after();
        const uint32_t tz = tmem_base + TM_Z + (uint32_t)zb * TV;
        for (int c = 0; c < p.nch; ++c, ++u) {
          const int s = u % p.estages;
          mbar_wait(&e_full[s], (uint32_t)(u / p.estages) & 1u);
          fence_tc_after();
          const uint32_t Ea = smem_u32(smem + (size_t)s * 2 * CHUNK);
          const uint64_t eh0 = kdesc(Ea), el0 = kdesc(Ea + CHUNK);      // + 2 per UMMA K step (32 bytes inside the 128-byte row)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t kk = (uint32_t)(c * 4 + ks);
            const uint32_t ah = tmem_base + TM_SA + kk * 8, al = ah + lo_off;
            e_umma_ts(tz, ah, eh0 + 2 * ks, p.idesc_z, kk ? 1u : 0u);
            e_umma_ts(tz, ah, el0 + 2 * ks, p.idesc_z, 1u);
            e_umma_ts(tz, al, eh0 + 2 * ks, p.idesc_z, 1u);
          }
          e_commit(&e_empty[s]);
        }
        e_commit(&z_full[zb]);
      }
    }
  } else {
    // soft-max math: identical to fce_fwd_kernel (per-row online max / sum exp in base 2, label logit)
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int b = ts.tb * TB + q * 32 + lane;
    const int lab = (b < p.B && p.labels != nullptr) ? p.labels[b] : -1;
    const float c2 = p.scale * LOG2E;
    const uint32_t lanebits = (uint32_t)(q * 32) << 16;
    {
      const uint16_t* src = (half == 0 ? p.Shi : p.Slo) + (long long)b * p.lds;
      const uint32_t dst = tmem_base + lanebits + TM_SA + (uint32_t)half * lo_off;
      for (int w = 0; w < (p.d >> 1); w += 8) {
        uint4 x0 = make_uint4(0u, 0u, 0u, 0u), x1 = x0;
        if (b < p.B) {
          x0 = *reinterpret_cast<const uint4*>(src + 2 * w);
          x1 = *reinterpret_cast<const uint4*>(src + 2 * w + 8);
        }
        const uint32_t v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        tmem_st8(dst + (uint32_t)w, v);
      }
      tmem_wait_st();
      fence_tc_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_full);
    }
    float m2 = -3.0e38f, s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, zl = 0.f;
    bool has = false;
    TopK tk;
    if (TOPK) tk.init();
    int it = 0;
    for (int t = ts.t0; t < ts.t1; ++t, ++it) {
      const int zb = it & 1;
      mbar_wait(&z_full[zb], (uint32_t)(it >> 1) & 1u);
      fence_tc_after();
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c0 = half * 64 + cc * 32;
        const uint32_t taddr = tmem_base + lanebits + TM_Z + (uint32_t)(zb * TV + c0);
        uint32_t r[32];
        tmem_ld32(taddr, r);
        const int v0 = t * TV + c0;
        const int nvalid = min(32, p.V - v0);
        if (nvalid == 32) {
          float a0 = __uint_as_float(r[0]), a1 = __uint_as_float(r[1]), a2 = __uint_as_float(r[2]), a3 = __uint_as_float(r[3]);
#pragma unroll
          for (int j = 4; j < 32; j += 4) {
            a0 = fmaxf(a0, __uint_as_float(r[j]));
            a1 = fmaxf(a1, __uint_as_float(r[j + 1]));
            a2 = fmaxf(a2, __uint_as_float(r[j + 2]));
            a3 = fmaxf(a3, __uint_as_float(r[j + 3]));
          }
          const float cm2 = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3)) * c2;
          if (cm2 > m2) {
            const float f = ex2f(m2 - cm2);
            s0 *= f; s1 *= f; s2 *= f; s3 *= f;
            m2 = cm2;
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            s0 += ex2f(fmaf(__uint_as_float(r[j]), c2, -m2));
            s1 += ex2f(fmaf(__uint_as_float(r[j + 1]), c2, -m2));
            s2 += ex2f(fmaf(__uint_as_float(r[j + 2]), c2, -m2));
            s3 += ex2f(fmaf(__uint_as_float(r[j + 3]), c2, -m2));
          }
        } else if (nvalid > 0) {
          float cm = -3.0e38f;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nvalid) cm = fmaxf(cm, __uint_as_float(r[j]));
          const float cm2 = cm * c2;
          if (cm2 > m2) {
            const float f = ex2f(m2 - cm2);
            s0 *= f; s1 *= f; s2 *= f; s3 *= f;
            m2 = cm2;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nvalid) s0 += ex2f(fmaf(__uint_as_float(r[j]), c2, -m2));
        }
        if (TOPK && nvalid > 0) tk.scan(r, nvalid, v0, p.topk);
        const bool mine = lab >= v0 && lab < v0 + nvalid;
        if (__any_sync(SRK_FULL, mine)) {
          const float x = p.scale * pick_column(taddr, lab - v0);
          if (mine) {
            zl = x;
            has = true;
          }
        }
      }
      fence_tc_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&z_empty[zb]);
    }
    if (b < p.B) {
      const int vr = blockIdx.x / p.ntm;
      float* pp = p.part + ((long long)(vr * 2 + half) * p.B + b) * 2;
      pp[0] = m2 * LN2;
      pp[1] = (s0 + s1) + (s2 + s3);
      if (has) p.zlab[b] = zl;
      if (TOPK) {
        float* cv = p.cand_val + ((long long)(vr * 2 + half) * p.B + b) * p.topk;
        int* ci = p.cand_idx + ((long long)(vr * 2 + half) * p.B + b) * p.topk;
        for (int k = 0; k < p.topk; ++k) {
          cv[k] = k < tk.n ? tk.v[k] : -3.0e38f;
          ci[k] = k < tk.n ? tk.i[k] : 0x7fffffff;
        }
      }
    }
  }
  fence_tc_before();
  __syncthreads();
  if (warp == 1) {
    fence_tc_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

__global__ void __launch_bounds__(THREADS_BWD, 1)
fce_bwd_wide_kernel(const __grid_constant__ CUtensorMap mSh, const __grid_constant__ CUtensorMap mSl,
                    const __grid_constant__ CUtensorMap mEh, const __grid_constant__ CUtensorMap mEl,
                    const __grid_constant__ CUtensorMap mdS, const FceWide p) {
  SRK_PDL();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_full, e_full[W_MAX_STAGES], e_empty[W_MAX_STAGES], z_full[2], z_empty[2], d_full, d_empty, de_free;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t op_bytes = (uint32_t)p.nch * CHUNK;          // one of shat hi / lo: nch chunks of 128 rows x 128 bytes
  uint8_t* S = smem;                                          // hi chunks | lo chunks
  uint8_t* Dt = smem + 2 * op_bytes;                          // dZ hi (one chunk) | dZ lo; fp32 staging of the final dS drain
  uint8_t* E0 = Dt + 2 * CHUNK;                               // estages x (hi chunk | lo chunk) of 64 rows
  const WideSched ts(p);
  const int ntiles = ts.t1 - ts.t0;

  if (threadIdx.x == 0) {
    mbar_init(&s_full, 1);
    for (int s = 0; s < W_MAX_STAGES; ++s) {
      mbar_init(&e_full[s], 1);
      mbar_init(&e_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&z_full[s], 1);
      mbar_init(&z_empty[s], EPI_WARPS);
    }
    mbar_init(&d_full, EPI_WARPS);
    mbar_init(&d_empty, 1);
    mbar_init(&de_free, DRAIN_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_tc_before();
  __syncthreads();
  fence_tc_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (issue_lane()) {
      // ===== TMA producer.  Stage order = the order the MMA warp consumes: L(0); then per tile L(it + 1), dS(it). =====
      mbar_expect_tx(&s_full, 2 * op_bytes);
      for (int c = 0; c < p.nch; ++c) {
        tma_load_2d(S + c * CHUNK, &mSh, &s_full, c * 64, ts.tb * TB);
        tma_load_2d(S + op_bytes + c * CHUNK, &mSl, &s_full, c * 64, ts.tb * TB);
      }
      int u = 0;
      auto tile_chunks = [&](int x) {
        for (int c = 0; c < p.nch; ++c, ++u) {
          const int s = u % p.estages;
          mbar_wait(&e_empty[s], ((uint32_t)(u / p.estages) & 1u) ^ 1u);
          uint8_t* st = E0 + (size_t)s * 2 * WECHUNK;
          mbar_expect_tx(&e_full[s], 2 * WECHUNK);
          tma_load_2d(st, &mEh, &e_full[s], c * 64, (ts.t0 + x) * WTV);
          tma_load_2d(st + WECHUNK, &mEl, &e_full[s], c * 64, (ts.t0 + x) * WTV);
        }
      };
      if (ntiles > 0) tile_chunks(0);
      for (int it = 0; it < ntiles; ++it) {
        if (it + 1 < ntiles) tile_chunks(it + 1);
        tile_chunks(it);
      }
    }
  } else if (warp == 1) {
    SRK_ISSUER_BLOCK {
      // ===== MMA issuer =====
      // The issuing thread is a single lane and the N = 64 MMAs of this kernel take only 32 cycles each: every instruction
      // between two tcgen05.mma is tensor-pipe idle time (the first version rebuilt four 64-bit descriptors per product
      // triple and was issue-bound at ~100 cycles per MMA).  All descriptors are therefore loop-invariant 64-bit bases built
      // once, plus small immediate offsets: the address field counts 16-byte units, +2 = 32 bytes (one UMMA K inside a
      // K-major 128-byte row), +128 = 2048 bytes (16 rows of an MN-major operand), +1024 = one 16 KB chunk.
      const uint32_t Sa = smem_u32(S), Da = smem_u32(Dt);
      const uint32_t tds = tmem_base + TMW_DS, tde = tmem_base + TMW_DE;
      const uint64_t sKh = kdesc(Sa), sKl = kdesc(Sa + op_bytes);              // shat, K-major (logits A)
      const uint64_t sMh = mdesc(Sa), sMl = mdesc(Sa + op_bytes);              // shat, MN-major (dE^T A)
      const uint64_t dKh = kdesc(Da), dKl = kdesc(Da + CHUNK);                 // dZ, K-major (dS A, shared-memory variant)
      const uint64_t dMh = mdesc(Da), dMl = mdesc(Da + CHUNK);                 // dZ, MN-major (dE^T B)
      int u = 0;
      // dz_tmem: ONE logit buffer (columns 0..63) and dZ as packed bf16 pairs at columns 64..127 (hi 32 | lo 32): the dS product
      // takes its A operand from TMEM and no longer re-reads the 128-row D tile from shared memory for every 64-column slice
      const bool dzt = p.dz_tmem != 0;
      const uint32_t tdz = tmem_base + TMW_Z + WTV;
      auto logits = [&](int x) {
        const int zb = dzt ? 0 : (x & 1);
        mbar_wait(&z_empty[zb], (dzt ? ((uint32_t)x & 1u) : ((uint32_t)(x >> 1) & 1u)) ^ 1u);
        fence_tc_after();
        const uint32_t tz = tmem_base + TMW_Z + (uint32_t)zb * WTV;
        for (int c = 0; c < p.nch; ++c, ++u) {
          const int s = u % p.estages;
          mbar_wait(&e_full[s], (uint32_t)(u / p.estages) & 1u);
          fence_tc_after();
          const uint32_t Ea = smem_u32(E0 + (size_t)s * 2 * WECHUNK);
          const uint64_t eh = kdesc(Ea), el = kdesc(Ea + WECHUNK);
          const uint64_t so = (uint64_t)((uint32_t)c * (CHUNK >> 4));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            mma3(tz, sKh + so + 2 * ks, sKl + so + 2 * ks, eh + 2 * ks, el + 2 * ks, p.idesc_z, (c | ks) ? 1u : 0u);
          e_commit(&e_empty[s]);
        }
        e_commit(&z_full[zb]);
      };
      mbar_wait(&s_full, 0);
      if (ntiles > 0) logits(0);
      for (int it = 0; it < ntiles; ++it) {
        if (it + 1 < ntiles) logits(it + 1);
        mbar_wait(&d_full, (uint32_t)it & 1u);
        fence_tc_after();
        // dS[128 b, d] += dZ[128 b x 64 v] E[64 v x d]: A = D K-major (or TMEM), B = the catalog tile MN-major.  Every pass over a
        // tile takes nch stages, so with a ring of exactly nch stages the chunks of this pass sit in slots 0 .. nch - 1, one
        // stage stride (= CHUNK bytes, the LBO of mdesc) apart: ONE N = d MMA per K step instead of nch N = 64 ones (the
        // issuing thread needs ~50-100 cycles per tcgen05.mma, a 32-cycle N = 64 MMA cannot hide that, a 128-cycle one can).
        if (p.estages == p.nch && u % p.nch == 0 && 2 * WECHUNK == CHUNK) {
          for (int c = 0; c < p.nch; ++c) {
            mbar_wait(&e_full[c], (uint32_t)((u + c) / p.estages) & 1u);
          }
          fence_tc_after();
          const uint32_t Ea = smem_u32(E0);
          const uint64_t bh = mdesc(Ea), bl = mdesc(Ea + WECHUNK);
          if (dzt) {
#pragma unroll
            for (int ks = 0; ks < WTV / 16; ++ks) {
              const uint32_t ah = tdz + (uint32_t)ks * 8, al = ah + 32;
              e_umma_ts(tds, ah, bh + 128 * ks, p.idesc_de2, (it | ks) ? 1u : 0u);
              e_umma_ts(tds, ah, bl + 128 * ks, p.idesc_de2, 1u);
              e_umma_ts(tds, al, bh + 128 * ks, p.idesc_de2, 1u);
            }
          } else {
#pragma unroll
            for (int ks = 0; ks < WTV / 16; ++ks)
              mma3(tds, dKh + 2 * ks, dKl + 2 * ks, bh + 128 * ks, bl + 128 * ks, p.idesc_de2, (it | ks) ? 1u : 0u);
          }
          for (int c = 0; c < p.nch; ++c) e_commit(&e_empty[c]);
          u += p.nch;
        } else {
          for (int c = 0; c < p.nch; ++c, ++u) {
            const int s = u % p.estages;
            mbar_wait(&e_full[s], (uint32_t)(u / p.estages) & 1u);
            fence_tc_after();
            const uint32_t Ea = smem_u32(E0 + (size_t)s * 2 * WECHUNK);
            const uint64_t bh = mdesc(Ea), bl = mdesc(Ea + WECHUNK);
            const uint32_t tacc = tds + (uint32_t)c * 64;
            if (dzt) {
#pragma unroll
              for (int ks = 0; ks < WTV / 16; ++ks) {
                const uint32_t ah = tdz + (uint32_t)ks * 8, al = ah + 32;
                e_umma_ts(tacc, ah, bh + 128 * ks, p.idesc_ds, (it | ks) ? 1u : 0u);
                e_umma_ts(tacc, ah, bl + 128 * ks, p.idesc_ds, 1u);
                e_umma_ts(tacc, al, bh + 128 * ks, p.idesc_ds, 1u);
              }
            } else {
#pragma unroll
              for (int ks = 0; ks < WTV / 16; ++ks)
                mma3(tacc, dKh + 2 * ks, dKl + 2 * ks, bh + 128 * ks, bl + 128 * ks, p.idesc_ds, (it | ks) ? 1u : 0u);
            }
            e_commit(&e_empty[s]);
          }
        }
        if (it > 0) {                             // the drain warps have read the previous tile's dE^T accumulators
          mbar_wait(&de_free, (uint32_t)(it - 1) & 1u);
          fence_tc_after();
        }
        // dE^T[128 embedding columns of half h, 64 v] = S_h^T[128 x 128 b] dZ[128 b x 64 v]: A = S MN-major (two chunks,
        // LBO = CHUNK), B = D MN-major, K = sessions in steps of 16 rows
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int ks = 0; ks < TB / 16; ++ks)
            mma3(tde + (uint32_t)h * WTV, sMh + 2048 * h + 128 * ks, sMl + 2048 * h + 128 * ks, dMh + 128 * ks, dMl + 128 * ks,
                 p.idesc_de, ks ? 1u : 0u);
        }
        e_commit(&d_empty);
      }
    }
  } else if (warp < 2 + EPI_WARPS) {
    // ===== math warps: 64-column logit tile -> dZ = coef (softmax - onehot) -> bf16 hi / lo -> D tile (one 64-column chunk) =====
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int b = ts.tb * TB + r;
    const bool bvalid = b < p.B;
    const int lab = bvalid ? p.labels[b] : -1;
    const float c2 = p.scale * LOG2E;
    const float lse_b = bvalid ? p.lse[b] : 0.f;
    const float lse2 = lse_b * LOG2E;
    const float coef = bvalid ? p.scale * (p.gout ? p.gout[0] : 1.f) / (float)p.B : 0.f;
    const float coef_p = coef * ex2f(-fmaf(lse_b, LOG2E, -lse2));
    const uint32_t lanebits = (uint32_t)(q * 32) << 16;
    uint8_t* Dhi = Dt;
    uint8_t* Dlo = Dt + CHUNK;
    const int c0 = half * 32;
    const bool dzt = p.dz_tmem != 0;
    const uint32_t tdz = tmem_base + lanebits + TMW_Z + WTV + (uint32_t)(c0 >> 1);       // this thread's 16 packed columns (hi; lo at +32)
    for (int it = 0; it < ntiles; ++it) {
      const int t = ts.t0 + it, zb = dzt ? 0 : (it & 1);
      mbar_wait(&z_full[zb], dzt ? ((uint32_t)it & 1u) : ((uint32_t)(it >> 1) & 1u));
      fence_tc_after();
      uint32_t hw[16], lw[16];
      const int v0 = t * WTV + c0;
      {
        uint32_t acc[32];
        tmem_ld32(tmem_base + lanebits + TMW_Z + (uint32_t)(zb * WTV + c0), acc);
        float dz[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) dz[j] = coef_p * ex2f(fmaf(__uint_as_float(acc[j]), c2, -lse2));
        if (v0 + 32 > p.V) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (v0 + j >= p.V) dz[j] = 0.f;
        }
        if (dzt && lab >= v0 && lab < v0 + 32) {        // the onehot term goes into the registers: both copies of dZ get it
          const int lj = lab - v0;
#pragma unroll
          for (int j = 0; j < 32; ++j) dz[j] -= (j == lj) ? coef : 0.f;
        }
        pack_dz(dz, hw, lw);
      }
      fence_tc_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&z_empty[zb]);
      if (it > 0) mbar_wait(&d_empty, (uint32_t)(it - 1) & 1u);       // previous tile's gradient products have read D (and dZ in TMEM)
      store_dz(Dhi, Dlo, r, c0, hw, lw);
      if (dzt) {
        fence_tc_after();
        tmem_st8(tdz, hw);
        tmem_st8(tdz + 8, hw + 8);
        tmem_st8(tdz + 32, lw);
        tmem_st8(tdz + 40, lw + 8);
        tmem_wait_st();
        fence_tc_before();
      } else if (lab >= v0 && lab < v0 + 32) {
        fix_dz(Dhi, Dlo, r, lab - t * WTV, coef);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&d_full);
    }
  } else {
    // ===== drain warps (TMEM lane quadrant = warp % 4).  dE^T: lane = embedding column, register j = catalog row, so one
    // store instruction of a warp writes 32 consecutive floats of ONE catalog row (128 bytes, coalesced). =====
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lanebits = (uint32_t)(q * 32) << 16;
    const bool elected = (warp == 2 + EPI_WARPS && lane == 0);
    // de_atomic: all session tiles add into one [V, d] buffer with coalesced 128-byte reductions in L2 (where a 17 MB buffer
    // stays resident) instead of writing one partial table per session tile to HBM for a later pass to add up
    float* outp = p.de_atomic ? p.dEpart : p.dEpart + (long long)ts.tb * p.V * p.d;
    for (int it = 0; it < ntiles; ++it) {
      mbar_wait(&d_empty, (uint32_t)it & 1u);
      fence_tc_after();
      const int v0 = (ts.t0 + it) * WTV;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {             // (embedding half h, 32 catalog rows): 32 accumulator columns at a time
        const int h = q4 >> 1, vo = (q4 & 1) * 32;
        uint32_t a[32];
        tmem_ld32(tmem_base + lanebits + TMW_DE + (uint32_t)(h * WTV + vo), a);
        if (q4 == 3) {                            // everything is in registers: the next tile may overwrite the accumulators
          fence_tc_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&de_free);
        }
        float* col = outp + (long long)(v0 + vo) * p.d + h * 128 + r;
        if (p.de_atomic) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (v0 + vo + j < p.V) atomicAdd(col + (long long)j * p.d, __uint_as_float(a[j]));
        } else if (v0 + vo + 32 <= p.V) {
#pragma unroll
          for (int j = 0; j < 32; ++j) col[(long long)j * p.d] = __uint_as_float(a[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (v0 + vo + j < p.V) col[(long long)j * p.d] = __uint_as_float(a[j]);
        }
      }
    }
    // dS accumulator of this CTA's whole catalog range -> TMA reduce-add into dS[B, d], 32 columns at a time through a 16 KB
    // staging buffer in the (now idle) D tile
    if (ntiles > 0) {
      for (int cc = 0; cc < (p.d >> 5); ++cc) {
        uint32_t a[32];
        tmem_ld32(tmem_base + lanebits + TMW_DS + (uint32_t)(cc * 32), a);
        if (elected) tma_wait_group_read0();
        named_bar_sync(2, DRAIN_THREADS);
        stage_row(Dt, r, a);
        fence_async_smem();
        named_bar_sync(2, DRAIN_THREADS);
        if (elected) {
          tma_reduce_add_2d(&mdS, Dt, cc * 32, ts.tb * TB);
          tma_commit_group();
        }
      }
      if (elected) tma_wait_group0();
    }
  }
  fence_tc_before();
  __syncthreads();
  if (warp == 1) {
    fence_tc_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

__global__ void split_bf16_kernel(const float* __restrict__ X, long long ldx, int rows, int cols, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, long long ldo) {
  SRK_PDL();
  const long long total = (long long)rows * cols;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const long long r = t / cols;
    const int c = (int)(t - r * cols);
    const float x = X[r * ldx + c];
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi[r * ldo + c] = h;
    lo[r * ldo + c] = __float2bfloat16_rn(x - __bfloat162float(h));
  }
}

long long* g_trace = nullptr;

int bf16_map(CUtensorMap* m, const uint16_t* base, int d, int rows, long long ld, int cw, int box_rows = 128) {
  SRK_REQUIRE(ld % 8 == 0, "flash_ce: bf16 operand pitch must be a multiple of 8 elements");
  cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)cw, (cuuint32_t)box_rows};
  // operand rows are 2 * d bytes apart and every 128-byte box row is consumed whole: L2 promotion beyond the box row only
  // multiplies the L2 -> SM sector traffic (measured: 3x with L2_256B at d = 96).  SESSREC_FCE_L2PROMO = 0..3 overrides.
  static int promo = -1;
  if (promo < 0) {
    const char* e = getenv("SESSREC_FCE_L2PROMO");
    promo = e ? atoi(e) : 0;
    if (promo < 0 || promo > 3) promo = 0;
  }
  const CUtensorMapL2promotion pm[4] = {CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_64B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B};
  return make_map_nd(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, pm[promo],
                     cw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}

int fill_params(FceParams& p, int B, int V, int d, float scale, const int* labels, bool bwd) {
  static int use_ts = -1;
  if (use_ts < 0) {
    const char* e = getenv("SESSREC_FCE_TS");
    use_ts = !(e && e[0] == '0');
  }
  SRK_REQUIRE(d >= 16 && d <= 128 && d % 16 == 0, "flash_ce: d = %d unsupported (multiple of 16 in [16, 128])", d);
  memset(&p, 0, sizeof(p));
  p.B = B; p.V = V; p.d = d;
  // chunk geometry: 32-column chunks (SWIZZLE_64B) when that shrinks the operand tiles, else 64 (SWIZZLE_128B);
  // SESSREC_FCE_CW = 32 / 64 forces one
  static int force_cw = -1;
  if (force_cw < 0) {
    const char* e = getenv("SESSREC_FCE_CW");
    force_cw = e ? atoi(e) : 0;
  }
  p.cw = (force_cw == 32 || force_cw == 64) ? force_cw : ((d % 64 != 0 && d % 32 == 0) ? 32 : 64);
  p.row_bytes = (uint32_t)p.cw * 2;
  p.chunk_bytes = 128 * p.row_bytes;
  p.kpc_log2 = p.cw == 64 ? 2 : 1;
  p.nch = (d + p.cw - 1) / p.cw;
  {
    const uint64_t type = p.cw == 64 ? 2ull : 4ull;          // SWIZZLE_128B : SWIZZLE_64B
    const uint64_t sbo = 8ull * p.row_bytes;                 // 8-row swizzle atom
    p.okd_hi = (uint64_t(16 >> 4) << 16) | ((sbo >> 4) << 32) | (1ull << 46) | (type << 61);
    p.omd_hi = (uint64_t(p.chunk_bytes >> 4) << 16) | ((sbo >> 4) << 32) | (1ull << 46) | (type << 61);
  }
  p.ntm = srk_cdiv(B, TB);
  p.nvt = srk_cdiv(V, TV);
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    SRK_CUDA(cudaGetDevice(&dev));
    SRK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  int nvr = sms / p.ntm;
  if (nvr < 1) nvr = 1;
  if (nvr > p.nvt) nvr = p.nvt;
  p.nvr = nvr;
  p.a_tmem = (!bwd && use_ts) ? 1 : 0;
  const size_t op = (size_t)p.nch * p.chunk_bytes;
  const size_t fixed = (p.a_tmem ? 0 : 2 * op) + (bwd ? D_BYTES + CHUNK : 0);
  int st = (int)((FCE_MAX_SMEM - 1024 - fixed) / (2 * op));
  if (st > (bwd ? 3 : MAX_STAGES)) st = bwd ? 3 : MAX_STAGES;
  SRK_REQUIRE(st >= 1, "flash_ce: shared-memory budget exceeded");
  p.estages = st;
  SRK_REQUIRE(scale > 0.f, "flash_ce: scale must be positive");
  p.scale = scale;
  p.labels = labels;
  p.trace = g_trace;
  // instruction descriptors: D = F32 (1 << 4), A = B = BF16 (1 << 7, 1 << 10), majors (bit 15 / 16: 1 = MN-major), N >> 3, M >> 4
  const uint32_t base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TB >> 4) << 24);
  p.idesc_z = base | ((uint32_t)(TV >> 3) << 17);
  p.idesc_ds = base | (1u << 16) | ((uint32_t)(d >> 3) << 17);
  p.idesc_de = base | (1u << 15) | (1u << 16) | ((uint32_t)(d >> 3) << 17);
  return SRK_OK;
}

size_t smem_bytes(const FceParams& p, bool bwd) {
  const size_t op = (size_t)p.nch * p.chunk_bytes;
  return (p.a_tmem ? 0 : 2 * op) + (bwd ? D_BYTES + CHUNK : 0) + (size_t)p.estages * 2 * op + 1024;
}

}  // namespace

extern "C" int srk_split_bf16(const float* X, long long ldx, int rows, int cols, uint16_t* hi, uint16_t* lo, long long ldo,
                              void* stream) {
  const long long total = (long long)rows * cols;
  if (total <= 0) return SRK_OK;
  long long g = (total + 255) / 256;
  if (g > 148LL * 16) g = 148LL * 16;
  srk_launch(split_bf16_kernel, (int)g, 256, 0, (cudaStream_t)stream, X, ldx, rows, cols, reinterpret_cast<__nv_bfloat16*>(hi),
                                                              reinterpret_cast<__nv_bfloat16*>(lo), ldo);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_flash_ce_supported(int d) {
  static int wide = -1;                   // SESSREC_FCE_WIDE=0 keeps d = 256 on the materialised 3xTF32 head
  if (wide < 0) {
    const char* e = getenv("SESSREC_FCE_WIDE");
    wide = !(e && e[0] == '0');
  }
  return (d >= 16 && d <= 128 && d % 16 == 0) || (d == 256 && wide);
}

extern "C" long long srk_flash_ce_part_floats(int B, int V);

namespace {

int fill_wide(FceWide& p, int B, int V, int d, float scale, const int* labels, bool bwd) {
  SRK_REQUIRE(d == 256, "flash_ce (wide): d = %d unsupported", d);
  SRK_REQUIRE(scale > 0.f, "flash_ce: scale must be positive");
  memset(&p, 0, sizeof(p));
  p.B = B; p.V = V; p.d = d;
  p.nch = d / 64;
  p.tv = bwd ? WTV : TV;
  p.ntm = srk_cdiv(B, TB);
  p.nvt = srk_cdiv(V, p.tv);
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    SRK_CUDA(cudaGetDevice(&dev));
    SRK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  int nvr = sms / p.ntm;
  if (nvr < 1) nvr = 1;
  if (nvr > p.nvt) nvr = p.nvt;
  p.nvr = nvr;
  const size_t fixed = bwd ? (size_t)2 * p.nch * CHUNK + 2 * CHUNK : 0;
  const size_t stage = bwd ? 2 * (size_t)WECHUNK : 2 * (size_t)CHUNK;
  int st = (int)((FCE_MAX_SMEM - 1024 - fixed) / stage);
  if (st > W_MAX_STAGES) st = W_MAX_STAGES;
  SRK_REQUIRE(st >= 2, "flash_ce (wide): shared-memory budget exceeded");
  if (bwd && st >= p.nch) st = p.nch;        // ring = one catalog tile: the dS pass finds its chunks in slots 0 .. nch - 1
  p.estages = st;
  p.scale = scale;
  p.labels = labels;
  // instruction descriptors: D = F32, A = B = BF16, majors (bit 15 / 16: 1 = MN-major), N >> 3, M >> 4
  const uint32_t base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TB >> 4) << 24);
  p.idesc_z = base | ((uint32_t)(p.tv >> 3) << 17);
  p.idesc_ds = base | (1u << 16) | ((uint32_t)(64 >> 3) << 17);
  p.idesc_de2 = base | (1u << 16) | ((uint32_t)(d >> 3) << 17);
  p.idesc_de = base | (1u << 15) | (1u << 16) | ((uint32_t)(WTV >> 3) << 17);
  return SRK_OK;
}

size_t wide_smem(const FceWide& p, bool bwd) {
  return (bwd ? (size_t)2 * p.nch * CHUNK + 2 * CHUNK + (size_t)p.estages * 2 * WECHUNK : (size_t)p.estages * 2 * CHUNK) + 1024;
}

int wide_fwd(int B, int V, int d, const uint16_t* Shi, const uint16_t* Slo, long long lds, const uint16_t* Ehi, const uint16_t* Elo,
             long long lde, float scale, const int* labels, float* lse, float* nll, float* part, cudaStream_t st, int topk = 0,
             int* nvr_out = nullptr) {
  FceWide p;
  SRK_TRY(fill_wide(p, B, V, d, scale, labels, false));
  p.part = part;
  p.zlab = part + 4LL * p.nvr * B;
  if (topk > 0) {
    p.topk = topk;
    p.cand_val = part + srk_flash_ce_part_floats(B, V);
    p.cand_idx = reinterpret_cast<int*>(p.cand_val + 2LL * p.nvr * B * topk);
    *nvr_out = p.nvr;
  }
  p.Shi = Shi; p.Slo = Slo; p.lds = lds;
  SRK_REQUIRE(((reinterpret_cast<uintptr_t>(Shi) | reinterpret_cast<uintptr_t>(Slo)) & 15u) == 0 && lds % 8 == 0,
              "flash_ce_fwd: shat operands must be 16-byte aligned");
  CUtensorMap mEh, mEl;
  SRK_TRY(bf16_map(&mEh, Ehi, d, V, lde, 64));
  SRK_TRY(bf16_map(&mEl, Elo, d, V, lde, 64));
  static bool attr_set = false;
  if (!attr_set) {
    SRK_CUDA(cudaFuncSetAttribute(fce_fwd_wide_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FCE_MAX_SMEM));
    SRK_CUDA(cudaFuncSetAttribute(fce_fwd_wide_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FCE_MAX_SMEM));
    attr_set = true;
  }
  if (topk > 0) {
    srk_launch(fce_fwd_wide_kernel<true>, p.ntm * p.nvr, THREADS, wide_smem(p, false), st, mEh, mEl, p);
    SRK_LAUNCH_CHECK();
    return SRK_OK;
  }
  srk_launch(fce_fwd_wide_kernel<false>, p.ntm * p.nvr, THREADS, wide_smem(p, false), st, mEh, mEl, p);
  SRK_LAUNCH_CHECK();
  srk_launch(fce_finalize_kernel, srk_cdiv((long long)B * 32, 256), 256, 0, st, p.part, p.zlab, labels, B, V, 2 * p.nvr, lse, nll);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

int wide_bwd(int B, int V, int d, const uint16_t* Shi, const uint16_t* Slo, long long lds, const uint16_t* Ehi, const uint16_t* Elo,
             long long lde, float scale, const int* labels, const float* lse, const float* gout, float* dS, float* dEpart,
             bool ds_zeroed, bool de_atomic, cudaStream_t st) {
  FceWide p;
  SRK_TRY(fill_wide(p, B, V, d, scale, labels, true));
  p.de_atomic = de_atomic;
  p.lse = lse;
  p.gout = gout;
  p.dEpart = dEpart;
  static int dzt = -1;                    // SESSREC_FCE_WIDE_DZT=0: dZ only in shared memory, double-buffered logit tile
  if (dzt < 0) {
    const char* e = getenv("SESSREC_FCE_WIDE_DZT");
    dzt = !(e && e[0] == '0');
  }
  p.dz_tmem = dzt;
  CUtensorMap mSh, mSl, mEh, mEl, mdS;
  SRK_TRY(bf16_map(&mSh, Shi, d, B, lds, 64));
  SRK_TRY(bf16_map(&mSl, Slo, d, B, lds, 64));
  SRK_TRY(bf16_map(&mEh, Ehi, d, V, lde, 64, WTV));
  SRK_TRY(bf16_map(&mEl, Elo, d, V, lde, 64, WTV));
  {
    cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)B};
    cuuint64_t strides[1] = {(cuuint64_t)d * 4};
    cuuint32_t box[2] = {32, 128};
    SRK_TRY(make_map_nd(&mdS, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dS, dims, strides, box));
  }
  static bool attr_set = false;
  if (!attr_set) {
    SRK_CUDA(cudaFuncSetAttribute(fce_bwd_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FCE_MAX_SMEM));
    attr_set = true;
  }
  if (!ds_zeroed) SRK_TRY(srk_zero_async(dS, sizeof(float) * (size_t)B * d, st));
  srk_launch(fce_bwd_wide_kernel, p.ntm * p.nvr, THREADS_BWD, wide_smem(p, true), st, mSh, mSl, mEh, mEl, mdS, p);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

}  // namespace

extern "C" long long srk_flash_ce_part_floats(int B, int V) {
  // upper bound that does not depend on the SM count: 2 partial pairs per (catalog tile, row) + label logits
  return 4LL * srk_cdiv(V, TV) * B + B;
}

static int narrow_fwd(int B, int V, int d, const uint16_t* Shi, const uint16_t* Slo, long long lds, const uint16_t* Ehi,
                      const uint16_t* Elo, long long lde, float scale, const int* labels, float* lse, float* nll, float* part,
                      cudaStream_t st, int topk, int* nvr_out) {
  FceParams p;
  SRK_TRY(fill_params(p, B, V, d, scale, labels, false));
  p.part = part;
  p.zlab = part + 4LL * p.nvr * B;
  p.Shi = Shi; p.Slo = Slo; p.lds = lds;
  if (topk > 0) {
    p.topk = topk;
    p.cand_val = part + srk_flash_ce_part_floats(B, V);
    p.cand_idx = reinterpret_cast<int*>(p.cand_val + 2LL * p.nvr * B * topk);
    *nvr_out = p.nvr;
  }
  SRK_REQUIRE(!p.a_tmem || ((reinterpret_cast<uintptr_t>(Shi) | reinterpret_cast<uintptr_t>(Slo)) & 15u) == 0,
              "flash_ce_fwd: shat operands must be 16-byte aligned");
  CUtensorMap mSh, mSl, mEh, mEl;
  SRK_TRY(bf16_map(&mSh, Shi, d, B, lds, p.cw));
  SRK_TRY(bf16_map(&mSl, Slo, d, B, lds, p.cw));
  SRK_TRY(bf16_map(&mEh, Ehi, d, V, lde, p.cw));
  SRK_TRY(bf16_map(&mEl, Elo, d, V, lde, p.cw));
  static bool attr_set = false;
  if (!attr_set) {
    SRK_CUDA(cudaFuncSetAttribute(fce_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FCE_MAX_SMEM));
    SRK_CUDA(cudaFuncSetAttribute(fce_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FCE_MAX_SMEM));
    attr_set = true;
  }
  if (topk > 0) {
    srk_launch(fce_fwd_kernel<true>, p.ntm * p.nvr, THREADS, smem_bytes(p, false), st, mSh, mSl, mEh, mEl, p);
    SRK_LAUNCH_CHECK();
    return SRK_OK;
  }
  srk_launch(fce_fwd_kernel<false>, p.ntm * p.nvr, THREADS, smem_bytes(p, false), st, mSh, mSl, mEh, mEl, p);
  SRK_LAUNCH_CHECK();
  srk_launch(fce_finalize_kernel, srk_cdiv((long long)B * 32, 256), 256, 0, st, p.part, p.zlab, labels, B, V, 2 * p.nvr, lse, nll);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_flash_ce_fwd(int B, int V, int d, const uint16_t* Shi, const uint16_t* Slo, long long lds,
                                const uint16_t* Ehi, const uint16_t* Elo, long long lde, float scale, const int* labels,
                                float* lse, float* nll, float* part, void* stream) {
  if (B <= 0) return SRK_OK;
  SRK_REQUIRE(V > 0 && labels != nullptr && lse != nullptr && part != nullptr, "flash_ce_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (d > 128) return wide_fwd(B, V, d, Shi, Slo, lds, Ehi, Elo, lde, scale, labels, lse, nll, part, st);
  return narrow_fwd(B, V, d, Shi, Slo, lds, Ehi, Elo, lde, scale, labels, lse, nll, part, st, 0, nullptr);
}

namespace {
// one CTA per row: the K best of the row's 2 * nvr candidate lists, best first (ties: smaller item id first)
__global__ void __launch_bounds__(128) fce_topk_merge_kernel(const float* __restrict__ cand_val, const int* __restrict__ cand_idx, int B,
                                                             int nlists, int K, int* __restrict__ out_idx, float* __restrict__ out_val,
                                                             float scale) {
  SRK_PDL();
  extern __shared__ float msm[];
  float* v = msm;
  int* ix = reinterpret_cast<int*>(msm + (size_t)nlists * K);
  __shared__ float rv[4];
  __shared__ int ri[4], rp[4];
  const int b = blockIdx.x, C = nlists * K;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int l = c / K, k = c - l * K;
    v[c] = cand_val[((long long)l * B + b) * K + k];
    ix[c] = cand_idx[((long long)l * B + b) * K + k];
  }
  __syncthreads();
  for (int k = 0; k < K; ++k) {
    float bv = -3.0e38f;
    int bi = 0x7fffffff, bp = -1;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float x = v[c];
      const int id = ix[c];
      if (x > bv || (x == bv && id < bi)) { bv = x; bi = id; bp = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(SRK_FULL, bv, o);
      const int oi = __shfl_xor_sync(SRK_FULL, bi, o), op = __shfl_xor_sync(SRK_FULL, bp, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; bp = op; }
    }
    if ((threadIdx.x & 31) == 0) { rv[threadIdx.x >> 5] = bv; ri[threadIdx.x >> 5] = bi; rp[threadIdx.x >> 5] = bp; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 4; ++w)
        if (rv[w] > bv || (rv[w] == bv && ri[w] < bi)) { bv = rv[w]; bi = ri[w]; bp = rp[w]; }
      out_idx[(long long)b * K + k] = bi;
      if (out_val) out_val[(long long)b * K + k] = scale * bv;
      if (bp >= 0) { v[bp] = -3.0e38f; ix[bp] = 0x7fffffff; }
    }
    __syncthreads();
  }
}
}  // namespace

/* floats of scratch srk_flash_ce_topk needs (soft-max partials + the candidate lists; SM-count independent bound) */
extern "C" long long srk_flash_ce_topk_scratch_floats(int B, int V, int K) {
  return srk_flash_ce_part_floats(B, V) + 2LL * 2 * 148 * B * K + 64;
}

/* Fused evaluation head: ids (best first) and optionally values of the K <= 32 largest logits scale * shat Ehat^T of every
 * row, straight from the tensor-core tiles of the fused scoring kernel - the (B, V) matrix is never written.  Replaces
 * `logits.topk(k=cutoff)` of evaluate() (utils/train.py:49; soft-max is monotone, so the ids are the same). */
extern "C" int srk_flash_ce_topk(int B, int V, int d, const uint16_t* Shi, const uint16_t* Slo, long long lds, const uint16_t* Ehi,
                                 const uint16_t* Elo, long long lde, float scale, int K, int* out_idx, float* out_val, float* scratch,
                                 void* stream) {
  if (B <= 0) return SRK_OK;
  SRK_REQUIRE(V > 0 && K >= 1 && K <= TOPK_MAX && K <= V && out_idx != nullptr && scratch != nullptr, "flash_ce_topk: bad arguments (K <= %d)", TOPK_MAX);
  SRK_REQUIRE(srk_flash_ce_supported(d), "flash_ce_topk: d = %d unsupported", d);
  cudaStream_t st = (cudaStream_t)stream;
  int nvr = 0;
  if (d > 128) SRK_TRY(wide_fwd(B, V, d, Shi, Slo, lds, Ehi, Elo, lde, scale, nullptr, nullptr, nullptr, scratch, st, K, &nvr));
  else SRK_TRY(narrow_fwd(B, V, d, Shi, Slo, lds, Ehi, Elo, lde, scale, nullptr, nullptr, nullptr, scratch, st, K, &nvr));
  SRK_REQUIRE(nvr >= 1 && nvr <= 148 * 2, "flash_ce_topk: unexpected range count %d", nvr);
  const float* cv = scratch + srk_flash_ce_part_floats(B, V);
  const int* ci = reinterpret_cast<const int*>(cv + 2LL * nvr * B * K);
  const size_t smem = (size_t)2 * nvr * K * 8;
  static bool attr_set = false;
  if (!attr_set) {
    SRK_CUDA(cudaFuncSetAttribute(fce_topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 148 * TOPK_MAX * 8 * 2));
    attr_set = true;
  }
  srk_launch(fce_topk_merge_kernel, B, 128, smem, st, cv, ci, B, 2 * nvr, K, out_idx, out_val, scale);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_flash_ce_bwd_parts(int B) { return srk_cdiv(B, TB); }

int srk_flash_ce_bwd_ex(int B, int V, int d, const uint16_t* Shi, const uint16_t* Slo, long long lds, const uint16_t* Ehi,
                        const uint16_t* Elo, long long lde, float scale, const int* labels, const float* lse, const float* gout,
                        float* dS, float* dEpart, int ds_zeroed, void* stream);

extern "C" int srk_flash_ce_bwd(int B, int V, int d, const uint16_t* Shi, const uint16_t* Slo, long long lds,
                                const uint16_t* Ehi, const uint16_t* Elo, long long lde, float scale, const int* labels,
                                const float* lse, const float* gout, float* dS, float* dEpart, void* stream) {
  return srk_flash_ce_bwd_ex(B, V, d, Shi, Slo, lds, Ehi, Elo, lde, scale, labels, lse, gout, dS, dEpart, 0, stream);
}

// ds_zeroed bit 0: the caller has zeroed dS already (the native steps zero it with their scratch pool): no memset launch here.
// bit 1: dEpart is ONE zero-initialised [V, d] buffer - e.g. the table's gradient rows themselves - that the
// kernel adds into, instead of srk_flash_ce_bwd_parts(B) partial tables
int srk_flash_ce_bwd_ex(int B, int V, int d, const uint16_t* Shi, const uint16_t* Slo, long long lds, const uint16_t* Ehi,
                        const uint16_t* Elo, long long lde, float scale, const int* labels, const float* lse, const float* gout,
                        float* dS, float* dEpart, int ds_zeroed, void* stream) {
  if (B <= 0) return SRK_OK;
  SRK_REQUIRE(V > 0 && labels != nullptr && lse != nullptr && dS != nullptr && dEpart != nullptr, "flash_ce_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (d > 128) return wide_bwd(B, V, d, Shi, Slo, lds, Ehi, Elo, lde, scale, labels, lse, gout, dS, dEpart, (ds_zeroed & 1) != 0, (ds_zeroed & 2) != 0, st);
  FceParams p;
  SRK_TRY(fill_params(p, B, V, d, scale, labels, true));
  p.lse = lse;
  p.gout = gout;
  p.dEpart = dEpart;
  p.de_atomic = (ds_zeroed & 2) != 0;
  CUtensorMap mSh, mSl, mEh, mEl, mdE, mdS;
  {
    cuuint64_t dims[3] = {(cuuint64_t)d, (cuuint64_t)V, (cuuint64_t)(p.de_atomic ? 1 : p.ntm)};
    cuuint64_t strides[2] = {(cuuint64_t)d * 4, (cuuint64_t)V * d * 4};
    cuuint32_t box[3] = {32, 128, 1};
    SRK_TRY(make_map_nd(&mdE, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dEpart, dims, strides, box));
  }
  SRK_TRY(bf16_map(&mSh, Shi, d, B, lds, p.cw));
  SRK_TRY(bf16_map(&mSl, Slo, d, B, lds, p.cw));
  SRK_TRY(bf16_map(&mEh, Ehi, d, V, lde, p.cw));
  SRK_TRY(bf16_map(&mEl, Elo, d, V, lde, p.cw));
  {
    cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)B};
    cuuint64_t strides[1] = {(cuuint64_t)d * 4};
    cuuint32_t box[2] = {32, 128};
    SRK_TRY(make_map_nd(&mdS, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dS, dims, strides, box));
  }
  static bool attr_set = false;
  if (!attr_set) {
    SRK_CUDA(cudaFuncSetAttribute(fce_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FCE_MAX_SMEM));
    attr_set = true;
  }
  if (!(ds_zeroed & 1)) SRK_TRY(srk_zero_async(dS, sizeof(float) * (size_t)B * d, st));
  srk_launch(fce_bwd_kernel, p.ntm * p.nvr, THREADS_BWD, smem_bytes(p, true), st, mSh, mSl, mEh, mEl, mdE, mdS, p);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_sum_parts(const float* parts, long long stride, int nparts, long long n, float* out, int accumulate,
                             void* stream) {
  if (n <= 0) return SRK_OK;
  SRK_REQUIRE(n % 4 == 0 && stride % 4 == 0 && ((reinterpret_cast<uintptr_t>(parts) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0,
              "sum_parts: n / stride must be multiples of 4 and the buffers 16-byte aligned");
  long long g = (n / 4 + 255) / 256;
  if (g > 148LL * 8) g = 148LL * 8;
  srk_launch(sum_parts_kernel, (int)g, 256, 0, (cudaStream_t)stream, reinterpret_cast<const float4*>(parts), stride / 4, nparts, n / 4,
                                                             reinterpret_cast<float4*>(out), accumulate);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

/* Debug aid: device buffer of 11 * 64 * 8 int64 that receives clock64() stamps of CTA 0's three roles (NULL = off). */
extern "C" int srk_flash_ce_set_trace(long long* trace_dev) {
  g_trace = trace_dev;
  return SRK_OK;
}
