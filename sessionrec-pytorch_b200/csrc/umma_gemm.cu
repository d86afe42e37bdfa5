// tcgen05 (5th-gen tensor core) GEMM for the catalog contraction, fp32-faithful through a 3xTF32 split.
//
//   C[M, N] (op)= alpha * A * B      with A = A_hi + A_lo, B = B_hi + B_lo (both halves exactly TF32-representable)
//   acc = A_hi B_hi + A_hi B_lo + A_lo B_hi      (fp32 accumulation in TMEM; the dropped A_lo B_lo term is ~2^-22)
//
// Three operand forms cover the scoring head of one training step (srgnn.py:146, niser.py:152, msgifsr.py:308 and
// their autograd):
//   form NT:  A[M, K] row-major, B[N, K] row-major     Z  = s E^T      (both operands K-major)
//   form NN:  A[M, K] row-major, B[K, N] row-major     dS = dZ E       (B is MN-major)
//   form TN:  A[K, M] row-major, B[K, N] row-major     dE = dZ^T s     (A and B are MN-major)
//
// Structure (one 128 x BN output tile per CTA, cta_group::1, UMMA M = 128, N = BN <= 256, K = 8 per MMA):
//   warp 0   TMA producer: cp.async.bulk.tensor.2d boxes of 128-byte rows (SWIZZLE_128B) into a STAGES-deep ring,
//            completion on an mbarrier (expect_tx)
//   warp 1   TMEM allocation + single-thread tcgen05.mma issue (3 MMAs per k-step), tcgen05.commit releases the
//            smem stage and finally signals the epilogue
//   warps 2-5  epilogue: tcgen05.ld (32 lanes x 32 columns per warp), alpha, masked store / atomicAdd (split-K)
// Shared-memory operand layouts are the canonical UMMA SWIZZLE_128B layouts for K-major and MN-major operands
// (8 rows x 128 B atoms, 1024 B apart), which is exactly what a TMA box with a 128-byte inner extent writes.
#include "umma.cuh"

namespace {

constexpr int BM = 128;          // UMMA M (TMEM lanes)
constexpr int KB = 32;           // floats per k-block = one 128-byte swizzle row
constexpr int UK = 8;            // TF32 UMMA K
constexpr int THREADS = 192;

using namespace umma;

// Epilogue store of one 32 x 32 accumulator chunk (lane = row) with coalesced global accesses: the chunk is transposed
// through a padded per-warp shared-memory tile so that every store instruction covers 4 rows x 128 contiguous bytes
// (the naive lane-per-row store touches 32 different cache lines with 16 bytes each and is LSU-bound).
// mode: 0 = store, 1 = atomicAdd (split-K: several CTAs add into one tile), 2 = C += tile with plain loads / stores (the CTA is
// the only writer of its tile: accumulate without split-K)
__device__ __forceinline__ void store_chunk(float* __restrict__ stile, const float* v, float* __restrict__ C, long long ldc,
                                            int m_base, int M, int nvalid, int lane, int mode) {
  const bool atomic = mode == 1;
#pragma unroll
  for (int j = 0; j < 32; ++j) stile[lane * 33 + j] = v[j];
  __syncwarp();
  const int c = (lane & 7) * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = (lane >> 3) + 4 * i;
    const int m = m_base + r;
    if (m < M && c < nvalid) {
      const float x0 = stile[r * 33 + c], x1 = stile[r * 33 + c + 1], x2 = stile[r * 33 + c + 2], x3 = stile[r * 33 + c + 3];
      float* dst = C + (long long)m * ldc + c;
      if (atomic) {
        atomicAdd(dst, x0);
        if (c + 1 < nvalid) atomicAdd(dst + 1, x1);
        if (c + 2 < nvalid) atomicAdd(dst + 2, x2);
        if (c + 3 < nvalid) atomicAdd(dst + 3, x3);
      } else if (c + 3 < nvalid && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
        float4 o = make_float4(x0, x1, x2, x3);
        if (mode == 2) {
          const float4 p = *reinterpret_cast<const float4*>(dst);
          o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
        }
        *reinterpret_cast<float4*>(dst) = o;
      } else if (mode == 2) {
        dst[0] += x0;
        if (c + 1 < nvalid) dst[1] += x1;
        if (c + 2 < nvalid) dst[2] += x2;
        if (c + 3 < nvalid) dst[3] += x3;
      } else {
        dst[0] = x0;
        if (c + 1 < nvalid) dst[1] = x1;
        if (c + 2 < nvalid) dst[2] = x2;
        if (c + 3 < nvalid) dst[3] = x3;
      }
    }
  }
  __syncwarp();
}

constexpr size_t EPI_TILE = 32 * 33 * sizeof(float);          // one padded 32 x 32 transpose tile per epilogue warp
constexpr size_t EPI_SMEM = 4 * EPI_TILE;                     // generic kernel: 4 epilogue warps
constexpr size_t MAX_DYN_SMEM = 227 * 1024 - 512;

struct UmmaParams {
  int M, N, K;            // logical problem
  int a_mn, b_mn;         // operand major-ness (0 = K-major, 1 = MN-major)
  int BN;                 // N tile (multiple of 16, <= 256)
  int kb_per_split;       // k-blocks per grid.z slice
  int stages;
  float* C;
  long long ldc;
  float alpha;
  int atomic;             // epilogue: 0 plain store, 1 atomicAdd (split-K), 2 C += (accumulate, one CTA per tile)
  const float* bias;      // optional [N]: added once (by the first split-K slice)
  uint32_t idesc;
  uint32_t tmem_cols;
};

__global__ void __launch_bounds__(THREADS, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap mAh, const __grid_constant__ CUtensorMap mAl,
                 const __grid_constant__ CUtensorMap mBh, const __grid_constant__ CUtensorMap mBl, const UmmaParams p) {
  SRK_PDL();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8], acc_bar;
  __shared__ uint32_t tmem_slot;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * p.BN;
  const int nkb_total = (p.K + KB - 1) / KB;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int kb1 = min(nkb_total, kb0 + p.kb_per_split);
  const int nkb = kb1 - kb0;

  const uint32_t a_bytes = BM * 128;                 // one of A_hi / A_lo for one k-block
  const uint32_t b_bytes = (uint32_t)p.BN * 128;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;

  if (nkb > 0) {
    if (warp == 0) {
      if (issue_lane()) {
        // ===== TMA producer =====
        for (int i = 0; i < nkb; ++i) {
          const int s = i % p.stages;
          const uint32_t ph = (uint32_t)(i / p.stages) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          uint8_t* st = smem + (size_t)s * stage_bytes;
          uint8_t *ah = st, *al = st + a_bytes, *bh = st + 2 * a_bytes, *bl = st + 2 * a_bytes + b_bytes;
          mbar_expect_tx(&full_bar[s], stage_bytes);
          const int k0 = (kb0 + i) * KB;
          if (!p.a_mn) {
            tma_load_2d(ah, &mAh, &full_bar[s], k0, m0);
            tma_load_2d(al, &mAl, &full_bar[s], k0, m0);
          } else {
            for (int c = 0; c < BM / 32; ++c) {
              tma_load_2d(ah + c * 4096, &mAh, &full_bar[s], m0 + c * 32, k0);
              tma_load_2d(al + c * 4096, &mAl, &full_bar[s], m0 + c * 32, k0);
            }
          }
          if (!p.b_mn) {
            tma_load_2d(bh, &mBh, &full_bar[s], k0, n0);
            tma_load_2d(bl, &mBl, &full_bar[s], k0, n0);
          } else {
            for (int c = 0; c < p.BN / 32; ++c) {
              tma_load_2d(bh + c * 4096, &mBh, &full_bar[s], n0 + c * 32, k0);
              tma_load_2d(bl + c * 4096, &mBl, &full_bar[s], n0 + c * 32, k0);
            }
          }
        }
      }
    } else if (warp == 1) {
      if (issue_lane()) {
        // ===== MMA issuer =====
        for (int i = 0; i < nkb; ++i) {
          const int s = i % p.stages;
          const uint32_t ph = (uint32_t)(i / p.stages) & 1u;
          mbar_wait(&full_bar[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
          const uint32_t ah = st, al = st + a_bytes, bh = st + 2 * a_bytes, bl = st + 2 * a_bytes + b_bytes;
#pragma unroll
          for (int ks = 0; ks < KB / UK; ++ks) {
            // K-major: advance 32 bytes inside the 128-byte swizzle row; MN-major: advance one 8-row atom (1024 B)
            const uint32_t aoff = p.a_mn ? ks * 1024 : ks * 32;
            const uint32_t boff = p.b_mn ? ks * 1024 : ks * 32;
            const uint64_t dah = p.a_mn ? make_desc(ah + aoff, 4096, 512, 1) : make_desc(ah + aoff, 16, 1024, 2);
            const uint64_t dal = p.a_mn ? make_desc(al + aoff, 4096, 512, 1) : make_desc(al + aoff, 16, 1024, 2);
            const uint64_t dbh = p.b_mn ? make_desc(bh + boff, 4096, 512, 1) : make_desc(bh + boff, 16, 1024, 2);
            const uint64_t dbl = p.b_mn ? make_desc(bl + boff, 4096, 512, 1) : make_desc(bl + boff, 16, 1024, 2);
            umma_tf32(tmem_base, dah, dbh, p.idesc, (i | ks) ? 1u : 0u);
            umma_tf32(tmem_base, dah, dbl, p.idesc, 1u);
            umma_tf32(tmem_base, dal, dbh, p.idesc, 1u);
          }
          umma_commit(&empty_bar[s]);          // frees this smem stage once the MMAs above have read it
        }
        umma_commit(&acc_bar);                 // accumulator complete
      }
    } else {
      // ===== epilogue: warps 2..5, TMEM lane quadrant = warp % 4 =====
      const int q = warp & 3;
      mbar_wait(&acc_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float* stile = reinterpret_cast<float*>(smem + (size_t)p.stages * stage_bytes) + (warp - 2) * 32 * 33;
      for (int c0 = 0; c0 < p.BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        const int nvalid = min(32, min(p.BN - c0, p.N - n0 - c0));
        if (nvalid > 0) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = p.alpha * __uint_as_float(r[j]);
          if (p.bias != nullptr && blockIdx.z == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nvalid) v[j] += p.bias[n0 + c0 + j];
          }
          store_chunk(stile, v, p.C + n0 + c0, p.ldc, m0 + q * 32, p.M, nvalid, lane, p.atomic);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ---- persistent forward scoring kernel with a fused log-sum-exp epilogue ------------------------------------------
// Z = alpha * A B^T (form NT) for the whole catalog in ONE persistent launch (one CTA per SM, static tile schedule
// ordered so that CTAs working at the same time share the catalog tile in L2), 128 x 256 tiles, two smem stages and
// TWO TMEM accumulators (2 x 256 columns = the whole TMEM) so that the epilogue of tile t overlaps TMA + MMA of tile
// t + 1.  The epilogue stores Z and, per row, the running (max, sum exp) of its 256 columns plus the label's logit:
// the separate 88 MB log-sum-exp pass over Z disappears (lse_finalize_kernel combines ntn partials per row).
constexpr int FBN = 256;
constexpr int FWD_EPI_WARPS = 8;                       // 2 per TMEM lane quadrant (latency hiding: 2 warps per SMSP)
constexpr int FWD_THREADS = 64 + 32 * FWD_EPI_WARPS;

struct FwdParams {
  int M, N, K, ntm, ntn;
  float* Z;
  long long ldz;
  float alpha;
  uint32_t idesc;
  const int* labels;     // optional
  float* part;           // [2 * ntn][M][2] (max, sum) per half tile
  float* zlab;           // [M] label logit
};


__global__ void __launch_bounds__(FWD_THREADS, 1)
umma_score_fwd_kernel(const __grid_constant__ CUtensorMap mAh, const __grid_constant__ CUtensorMap mAl,
                      const __grid_constant__ CUtensorMap mBh, const __grid_constant__ CUtensorMap mBl, const FwdParams p) {
  SRK_PDL();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[2], empty_bar[2], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (p.K + KB - 1) / KB;
  const int ntiles = p.ntm * p.ntn;
  constexpr uint32_t a_bytes = BM * 128, b_bytes = FBN * 128, stage_bytes = 2 * a_bytes + 2 * b_bytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], FWD_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (issue_lane()) {
      int it = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int m0 = (t % p.ntm) * BM, n0 = (t / p.ntm) * FBN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it & 1;
          mbar_wait(&empty_bar[s], ((uint32_t)(it >> 1) & 1u) ^ 1u);
          uint8_t* st = smem + (size_t)s * stage_bytes;
          mbar_expect_tx(&full_bar[s], stage_bytes);
          tma_load_2d(st, &mAh, &full_bar[s], kb * KB, m0);
          tma_load_2d(st + a_bytes, &mAl, &full_bar[s], kb * KB, m0);
          tma_load_2d(st + 2 * a_bytes, &mBh, &full_bar[s], kb * KB, n0);
          tma_load_2d(st + 2 * a_bytes + b_bytes, &mBl, &full_bar[s], kb * KB, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (issue_lane()) {
      int it = 0, tc = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tc) {
        const int buf = tc & 1;
        mbar_wait(&tempty_bar[buf], ((uint32_t)(tc >> 1) & 1u) ^ 1u);      // epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_base + (uint32_t)buf * FBN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it & 1;
          mbar_wait(&full_bar[s], (uint32_t)(it >> 1) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
          const uint32_t ah = st, al = st + a_bytes, bh = st + 2 * a_bytes, bl = st + 2 * a_bytes + b_bytes;
#pragma unroll
          for (int ks = 0; ks < KB / UK; ++ks) {
            const uint64_t dah = make_desc(ah + ks * 32, 16, 1024, 2), dal = make_desc(al + ks * 32, 16, 1024, 2);
            const uint64_t dbh = make_desc(bh + ks * 32, 16, 1024, 2), dbl = make_desc(bl + ks * 32, 16, 1024, 2);
            umma_tf32(tacc, dah, dbh, p.idesc, (kb | ks) ? 1u : 0u);
            umma_tf32(tacc, dah, dbl, p.idesc, 1u);
            umma_tf32(tacc, dal, dbh, p.idesc, 1u);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tfull_bar[buf]);
      }
    }
  } else {
    // epilogue warps 2..9: TMEM lane quadrant q = warp % 4, column half = (warp - 2) / 4 (128 of the 256 columns each)
    const int q = warp & 3, half = (warp - 2) >> 2;
    float* stile = reinterpret_cast<float*>(smem + 2 * (size_t)stage_bytes) + (warp - 2) * 32 * 33;
    int tc = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tc) {
      const int buf = tc & 1;
      const int mt = t % p.ntm, nt = t / p.ntm;
      const int m = mt * BM + q * 32 + lane, n0 = nt * FBN;
      mbar_wait(&tfull_bar[buf], (uint32_t)(tc >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int lab = (p.labels && m < p.M) ? p.labels[m] - n0 : -1;
      float rmax = -3.0e38f, rsum = 0.f, zl = 0.f;
      bool has = false;
      for (int c0 = half * (FBN / 2); c0 < (half + 1) * (FBN / 2); c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * FBN + c0), r);
        const int nvalid = min(32, p.N - n0 - c0);
        if (m < p.M && nvalid > 0) {
          float v[32];
          float cm = -3.0e38f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = p.alpha * __uint_as_float(r[j]);
            if (j < nvalid) cm = fmaxf(cm, v[j]);
          }
          if (cm > rmax) {
            rsum *= __expf(rmax - cm);
            rmax = cm;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nvalid) rsum += __expf(v[j] - rmax);
          if (lab >= c0 && lab < c0 + 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j == lab - c0) zl = v[j];
            has = true;
          }
        }
        if (nvalid > 0) {        // warp-uniform: coalesced store through the per-warp transpose tile
          float v2[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v2[j] = p.alpha * __uint_as_float(r[j]);
          store_chunk(stile, v2, p.Z + n0 + c0, p.ldz, mt * BM + q * 32, p.M, nvalid, lane, 0);
        }
      }
      // accumulator fully read: hand the TMEM buffer back to the MMA warp
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
      if (m < p.M) {
        float* pp = p.part + ((long long)(nt * 2 + half) * p.M + m) * 2;
        pp[0] = rmax;
        pp[1] = rsum;
        if (has) p.zlab[m] = zl;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// lse[m] = log sum over the ntn tile partials; nll[m] = lse[m] - zlab[m].  Warp per row.
__global__ void __launch_bounds__(256) lse_finalize_kernel(const float* __restrict__ part, const float* __restrict__ zlab,
                                                           int M, int ntn, float* __restrict__ lse, float* __restrict__ nll) {
  SRK_PDL();
  const int lane = threadIdx.x & 31;
  const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (m >= M) return;
  float mx = -3.0e38f, s = 0.f;
  for (int t = lane; t < ntn; t += 32) {
    const float pm = part[((long long)t * M + m) * 2], ps = part[((long long)t * M + m) * 2 + 1];
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
    lse[m] = l;
    if (nll) nll[m] = l - zlab[m];
  }
}

__global__ void split_tf32_kernel(const float* __restrict__ X, long long ldx, int rows, int cols, float* __restrict__ hi,
                                  float* __restrict__ lo, long long ldo) {
  SRK_PDL();
  long long total = (long long)rows * cols;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    long long r = t / cols;
    int c = (int)(t - r * cols);
    float x = X[r * ldx + c];
    float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    hi[r * ldo + c] = h;
    lo[r * ldo + c] = x - h;
  }
}

// TF32 hi / lo split of TWO pitched fp32 matrices into dense copies in ONE launch (the operand pair of srk_tc_gemm)
struct Split2 {
  const float* X[2];
  float *hi[2], *lo[2];
  long long ldx[2], ldo[2];
  int rows[2], cols[2];
  long long n4[2];        // float4 groups per matrix (cols rounded up to 4; ldo is a multiple of 4)
};

__global__ void __launch_bounds__(256) split2_tf32_kernel(const Split2 sp) {
  SRK_PDL();
  const long long total = sp.n4[0] + sp.n4[1];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int w = t >= sp.n4[0];
    const long long q = w ? t - sp.n4[0] : t;
    const int c4 = (sp.cols[w] + 3) >> 2;
    const long long r = q / c4;
    const int c = (int)(q - r * c4) * 4;
    const float* src = sp.X[w] + r * sp.ldx[w] + c;
    float x[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = (c + j < sp.cols[w]) ? src[j] : 0.f;
    float h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h[j] = __uint_as_float(__float_as_uint(x[j]) & 0xFFFFE000u);
      l[j] = x[j] - h[j];
    }
    *reinterpret_cast<float4*>(sp.hi[w] + r * sp.ldo[w] + c) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(sp.lo[w] + r * sp.ldo[w] + c) = make_float4(l[0], l[1], l[2], l[3]);
  }
}

int umma_gemm_impl(int form, int M, int N, int K, const float* Ahi, const float* Alo, long long lda, const float* Bhi,
                   const float* Blo, long long ldb, float* C, long long ldc, const float* bias, float alpha, int accumulate,
                   int split_k, void* stream);

}  // namespace

// form: 0 = NT (A[M,K], B[N,K]), 1 = NN (A[M,K], B[K,N]), 2 = TN (A[K,M], B[K,N]).  All row-major with pitches lda / ldb.
extern "C" int srk_umma_gemm(int form, int M, int N, int K, const float* Ahi, const float* Alo, long long lda,
                             const float* Bhi, const float* Blo, long long ldb, float* C, long long ldc, float alpha,
                             int accumulate, int split_k, void* stream) {
  return umma_gemm_impl(form, M, N, K, Ahi, Alo, lda, Bhi, Blo, ldb, C, ldc, nullptr, alpha, accumulate, split_k, stream);
}

static inline long long r4(long long x) { return (x + 3) / 4 * 4; }

int srk_tc_gemm_pre(int form, int M, int N, int K, const float* A, long long lda, const float* Ahi, const float* Alo,
                    long long ldah, const float* B, long long ldb, const float* Bhi, const float* Blo, long long ldbh, float* C,
                    long long ldc, const float* bias, float alpha, int accumulate, int split_k, float* scratch, void* stream);

// floats of scratch srk_tc_gemm needs: dense hi / lo copies of both operands
extern "C" long long srk_tc_gemm_scratch_floats(int form, int M, int N, int K) {
  const long long a = form == 2 ? (long long)K * r4(M) : (long long)M * r4(K);
  const long long b = form == 0 ? (long long)N * r4(K) : (long long)K * r4(N);
  return 2 * (a + b) + 64;
}

// C (+)= alpha * op(A) op(B) (+ bias) from PLAIN fp32 operands on the tcgen05 tensor cores: one launch splits both operands
// into TF32 hi / lo pairs (scratch), then the 3xTF32 kernel runs.  Same forms as srk_umma_gemm; any N (forms 1 and 2 are
// tiled over 256 output columns).  split_k <= 0 picks a split that fills the SMs (needs accumulate).
extern "C" int srk_tc_gemm(int form, int M, int N, int K, const float* A, long long lda, const float* B, long long ldb,
                           float* C, long long ldc, const float* bias, float alpha, int accumulate, int split_k,
                           float* scratch, void* stream) {
  return srk_tc_gemm_pre(form, M, N, K, A, lda, nullptr, nullptr, 0, B, ldb, nullptr, nullptr, 0, C, ldc, bias, alpha, accumulate,
                         split_k, scratch, stream);
}

// The same with operands that may already be split: Ahi / Alo (pitch ldah) and / or Bhi / Blo (pitch ldbh) non-null = that
// operand's TF32 pair exists (its producer wrote it, or an earlier product of the step split it) and only the other one is
// split here; both given = no split launch at all.  The native steps split every activation and weight once.
int srk_tc_gemm_pre(int form, int M, int N, int K, const float* A, long long lda, const float* Ahi, const float* Alo,
                    long long ldah, const float* B, long long ldb, const float* Bhi, const float* Blo, long long ldbh, float* C,
                    long long ldc, const float* bias, float alpha, int accumulate, int split_k, float* scratch, void* stream) {
  SRK_REQUIRE(form >= 0 && form <= 2, "tc_gemm: bad form %d", form);
  if (M <= 0 || N <= 0) return SRK_OK;
  const bool need_a = Ahi == nullptr, need_b = Bhi == nullptr;
  SRK_REQUIRE(K > 0 && (!(need_a || need_b) || (scratch != nullptr && (reinterpret_cast<uintptr_t>(scratch) & 15u) == 0)),
              "tc_gemm: bad arguments");
  SRK_REQUIRE((need_a || Alo != nullptr) && (need_b || Blo != nullptr), "tc_gemm: a pre-split operand needs both halves");
  Split2 sp;
  memset(&sp, 0, sizeof(sp));
  sp.X[0] = A; sp.ldx[0] = lda;
  sp.X[1] = B; sp.ldx[1] = ldb;
  sp.rows[0] = form == 2 ? K : M; sp.cols[0] = form == 2 ? M : K;
  sp.rows[1] = form == 0 ? N : K; sp.cols[1] = form == 0 ? K : N;
  const float *hi[2] = {Ahi, Bhi}, *lo[2] = {Alo, Blo};
  long long ldo[2] = {ldah, ldbh};
  float* w = scratch;
  for (int i = 0; i < 2; ++i) {
    if (i == 0 ? !need_a : !need_b) {
      sp.rows[i] = 0;                          // nothing to split
      sp.n4[i] = 0;
      sp.cols[i] = 4;
      continue;
    }
    sp.ldo[i] = r4(sp.cols[i]);
    sp.n4[i] = (long long)sp.rows[i] * (sp.ldo[i] / 4);
    sp.hi[i] = w; w += (long long)sp.rows[i] * sp.ldo[i];
    sp.lo[i] = w; w += (long long)sp.rows[i] * sp.ldo[i];
    hi[i] = sp.hi[i]; lo[i] = sp.lo[i]; ldo[i] = sp.ldo[i];
  }
  if (need_a || need_b) {
    long long g = (sp.n4[0] + sp.n4[1] + 255) / 256;
    if (g > 148LL * 8) g = 148LL * 8;
    srk_launch(split2_tf32_kernel, (int)g, 256, 0, (cudaStream_t)stream, sp);
    SRK_LAUNCH_CHECK();
  }
  const int nstep = form == 0 ? N : 256;
  for (int n0 = 0; n0 < N; n0 += nstep) {
    const int nc = N - n0 < nstep ? N - n0 : nstep;
    int S = split_k;
    if (S <= 0) {
      if (!accumulate) S = 1;
      else {
        const int bn = form == 0 ? 128 : nc;
        const long long tiles = (long long)((M + 127) / 128) * ((nc + bn - 1) / bn);
        S = (int)(148 / (tiles < 1 ? 1 : tiles));
        if (S < 1) S = 1;
      }
    }
    const float* bh = form == 0 ? hi[1] : hi[1] + n0;
    const float* bl = form == 0 ? lo[1] : lo[1] + n0;
    SRK_TRY(umma_gemm_impl(form, M, nc, K, hi[0], lo[0], ldo[0], bh, bl, ldo[1], C + n0, ldc, bias ? bias + n0 : nullptr, alpha,
                           accumulate, S, stream));
  }
  return SRK_OK;
}

namespace {
int umma_gemm_impl(int form, int M, int N, int K, const float* Ahi, const float* Alo, long long lda, const float* Bhi,
                   const float* Blo, long long ldb, float* C, long long ldc, const float* bias, float alpha, int accumulate,
                   int split_k, void* stream) {
  SRK_REQUIRE(form >= 0 && form <= 2, "umma_gemm: bad form %d", form);
  if (M <= 0 || N <= 0) return SRK_OK;
  SRK_REQUIRE(K > 0, "umma_gemm: K must be positive");
  UmmaParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  p.a_mn = form == 2;
  p.b_mn = form != 0;
  // N tile: the catalog dimension is tiled by 128; an embedding-dim N must fit one tile (multiple of 32 for the
  // MN-major 32-float chunks, <= 256)
  if (form == 0) {
    // wide tiles halve the A re-reads and the per-CTA prologue/teardown count; they need K small enough that two
    // 96 KB stages cover most of the K loop
    p.BN = N >= 256 && K <= 128 ? 256 : (N >= 128 ? 128 : ((N + 15) / 16) * 16);
  } else {
    p.BN = ((N + 31) / 32) * 32;
  }
  SRK_REQUIRE(p.BN >= 16 && p.BN <= 256 && p.BN % 16 == 0, "umma_gemm: N tile %d unsupported", p.BN);
  SRK_REQUIRE(form == 0 || N <= 256, "umma_gemm: N = %d > 256 needs N tiling for MN-major B (not built)", N);
  const int nkb = (K + KB - 1) / KB;
  int S = split_k < 1 ? 1 : split_k;
  if (S > nkb) S = nkb;
  SRK_REQUIRE(S == 1 || accumulate, "umma_gemm: split-K needs accumulate mode");
  p.kb_per_split = (nkb + S - 1) / S;
  S = (nkb + p.kb_per_split - 1) / p.kb_per_split;
  p.C = C; p.ldc = ldc; p.alpha = alpha;
  p.atomic = accumulate ? (S > 1 ? 1 : 2) : 0;
  p.bias = bias;
  p.tmem_cols = p.BN <= 32 ? 32 : (p.BN <= 64 ? 64 : (p.BN <= 128 ? 128 : 256));
  // instruction descriptor: D = F32, A = B = TF32, majors, N >> 3, M >> 4
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
            ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  const size_t stage_bytes = 2 * (size_t)BM * 128 + 2 * (size_t)p.BN * 128;
  int stages = (int)((MAX_DYN_SMEM - 1024 - EPI_SMEM) / stage_bytes);
  if (stages > 8) stages = 8;
  if (stages > p.kb_per_split) stages = p.kb_per_split;
  if (stages < 1) stages = 1;
  p.stages = stages;
  const size_t smem = stages * stage_bytes + EPI_SMEM + 1024;

  CUtensorMap mAh, mAl, mBh, mBl;
  if (!p.a_mn) {          // A[M, K]: inner = K
    SRK_TRY(make_map(&mAh, Ahi, K, M, lda, BM, false));
    SRK_TRY(make_map(&mAl, Alo, K, M, lda, BM, false));
  } else {                // A[K, M]: inner = M, boxes of 32 (m) x 32 (k)
    SRK_TRY(make_map(&mAh, Ahi, M, K, lda, KB, true));
    SRK_TRY(make_map(&mAl, Alo, M, K, lda, KB, true));
  }
  if (!p.b_mn) {          // B[N, K]
    SRK_TRY(make_map(&mBh, Bhi, K, N, ldb, p.BN, false));
    SRK_TRY(make_map(&mBl, Blo, K, N, ldb, p.BN, false));
  } else {                // B[K, N]
    SRK_TRY(make_map(&mBh, Bhi, N, K, ldb, KB, true));
    SRK_TRY(make_map(&mBl, Blo, N, K, ldb, KB, true));
  }
  static bool attr_set = false;
  if (!attr_set) {
    SRK_CUDA(cudaFuncSetAttribute(umma_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MAX_DYN_SMEM));
    attr_set = true;
  }
  dim3 grid(srk_cdiv(N, p.BN), srk_cdiv(M, BM), S);
  SRK_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "umma_gemm: grid too large");
  srk_launch(umma_gemm_kernel, grid, THREADS, smem, (cudaStream_t)stream, mAh, mAl, mBh, mBl, p);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}
}  // namespace

extern "C" int srk_split_tf32(const float* X, long long ldx, int rows, int cols, float* hi, float* lo, long long ldo,
                              void* stream) {
  long long total = (long long)rows * cols;
  if (total <= 0) return SRK_OK;
  long long g = (total + 255) / 256;
  if (g > 148LL * 16) g = 148LL * 16;
  srk_launch(split_tf32_kernel, (int)g, 256, 0, (cudaStream_t)stream, X, ldx, rows, cols, hi, lo, ldo);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

// Z[M, ldz] = alpha * A B^T with A[M, K], B[N, K] given as TF32 hi/lo pairs; lse[M] = row log-sum-exp of Z, nll[M] =
// lse - Z[m, labels[m]] (labels / nll optional).  part: scratch of 4 * ceil(N / 256) * M + M floats.
extern "C" int srk_umma_score_fwd(int M, int N, int K, const float* Ahi, const float* Alo, long long lda, const float* Bhi,
                                  const float* Blo, long long ldb, float* Z, long long ldz, float alpha, const int* labels,
                                  float* lse, float* nll, float* part, void* stream) {
  if (M <= 0 || N <= 0) return SRK_OK;
  SRK_REQUIRE(K > 0 && (nll == nullptr || labels != nullptr), "umma_score_fwd: bad arguments");
  FwdParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  p.ntm = srk_cdiv(M, BM);
  p.ntn = srk_cdiv(N, FBN);
  p.Z = Z; p.ldz = ldz; p.alpha = alpha; p.labels = labels;
  p.part = part;
  p.zlab = part + 4LL * p.ntn * M;
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(FBN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  CUtensorMap mAh, mAl, mBh, mBl;
  SRK_TRY(make_map(&mAh, Ahi, K, M, lda, BM, false));
  SRK_TRY(make_map(&mAl, Alo, K, M, lda, BM, false));
  SRK_TRY(make_map(&mBh, Bhi, K, N, ldb, FBN, false));
  SRK_TRY(make_map(&mBl, Blo, K, N, ldb, FBN, false));
  const size_t smem = 2 * (2 * (size_t)BM * 128 + 2 * (size_t)FBN * 128) + FWD_EPI_WARPS * EPI_TILE + 1024;
  static_assert(2 * (2 * (size_t)BM * 128 + 2 * (size_t)FBN * 128) + FWD_EPI_WARPS * EPI_TILE + 1024 <= MAX_DYN_SMEM, "smem budget");
  static bool attr_set = false;
  static int sms = 148;
  if (!attr_set) {
    SRK_CUDA(cudaFuncSetAttribute(umma_score_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MAX_DYN_SMEM));
    int dev = 0;
    SRK_CUDA(cudaGetDevice(&dev));
    SRK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    attr_set = true;
  }
  const int ntiles = p.ntm * p.ntn;
  srk_launch(umma_score_fwd_kernel, ntiles < sms ? ntiles : sms, FWD_THREADS, smem, (cudaStream_t)stream, mAh, mAl, mBh, mBl, p);
  SRK_LAUNCH_CHECK();
  srk_launch(lse_finalize_kernel, srk_cdiv((long long)M * 32, 256), 256, 0, (cudaStream_t)stream, p.part, p.zlab, M, 2 * p.ntn, lse, nll);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}
