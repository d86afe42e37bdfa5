// K1 / K8 / K6a: item-embedding gather (+dropout +row normalisation), deterministic scatter-add of the
// embedding gradient, catalog pre-pass (max_norm renorm + row normalisation) and generic row normalisation.
// All kernels are warp-per-row with 128-bit coalesced accesses (rowops.cuh); HBM/L2-bandwidth bound.
#include <stdlib.h>

#include "rowops.cuh"

namespace {

constexpr int ROW_THREADS = 256;  // 8 warps (rows) per CTA

template <int NC>
__global__ void __launch_bounds__(ROW_THREADS) gather_fwd_kernel(const float* __restrict__ E, const int* __restrict__ iid,
                                                                 int P, int d, int mode, DropCfg dc,
                                                                 float* __restrict__ X, float* __restrict__ rnorm,
                                                                 float* __restrict__ x_first) {
  SRK_PDL();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < P; i += warps) {
    RowVec<NC> x, y;
    row_load(x, E + (long long)iid[i] * d, d, lane);
    row_dropout(x, dc, i, d, lane);
    float n;
    if (mode == SRK_NORM_NISER && x_first) {
      n = sqrtf(row_dot(x, x));
      y = x;
      row_scale(y, 1.f / (n + 1e-12f));
      row_store(y, x_first + (long long)i * d, d, lane);
      float n1 = sqrtf(row_dot(y, y));
      row_scale(y, 1.f / n1);
    } else if (mode == SRK_NORM_NONE) {
      n = 0.f;
      y = x;
    } else {
      n = row_normalize(x, y, mode);
    }
    row_store(y, X + (long long)i * d, d, lane);
    if (rnorm && lane == 0) rnorm[i] = n;
  }
}

// The same gather with the table rows STAGED THROUGH SHARED MEMORY BY THE BULK-COPY ENGINE (TMA, `cp.async.bulk`): every warp
// owns a two-deep ring of row groups; its lane 0 issues one bulk copy per row (d * 4 bytes, 16-byte aligned) of the NEXT
// group while the warp normalises the current one out of shared memory.  The copies complete on an mbarrier
// (expect_tx / complete_tx), so a warp keeps 2 x G rows in flight without holding a single register for them - what a
// latency-bound gather of a few thousand L2-resident rows wants - and the normalised rows leave with coalesced 128-bit
// stores.  G is sized on the host so that the CTA's rings fit GATHER_TMA_SMEM bytes.
constexpr int GATHER_TMA_SMEM = 96 * 1024;

__device__ __forceinline__ uint32_t g_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void g_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(g_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void g_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(g_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void g_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "GWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra GDONE;\n\t"
      "bra GWAIT_LOOP;\n\t"
      "GDONE:\n\t"
      "}" ::"r"(g_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void g_bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(g_smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(g_smem_u32(bar))
               : "memory");
}

template <int NC>
__global__ void __launch_bounds__(ROW_THREADS) gather_tma_kernel(const float* __restrict__ E, const int* __restrict__ iid, int P,
                                                                 int d, int mode, DropCfg dc, int G, float* __restrict__ X,
                                                                 float* __restrict__ rnorm, float* __restrict__ x_first) {
  SRK_PDL();
  extern __shared__ __align__(128) uint8_t gsm[];
  __shared__ __align__(8) uint64_t bars[ROW_THREADS / 32][2];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t row_bytes = (uint32_t)d * 4;
  float* ring = reinterpret_cast<float*>(gsm) + (size_t)wib * 2 * G * d;        // this warp's two row groups
  if (lane == 0) {
    g_mbar_init(&bars[wib][0], 1);
    g_mbar_init(&bars[wib][1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int w0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int ngroups = (P + G - 1) / G;
  // issue the bulk copies of group g (rows g * G ...) into ring slot `slot`
  auto issue = [&](int g, int slot) {             // called by the whole warp: lane k issues the copy of the group's k-th row
    const int r0 = g * G, n = min(G, P - r0);
    if (lane == 0) g_mbar_expect_tx(&bars[wib][slot], (uint32_t)n * row_bytes);
    __syncwarp();
    if (lane < n)
      g_bulk_load(ring + ((size_t)slot * G + lane) * d, E + (long long)iid[r0 + lane] * d, row_bytes, &bars[wib][slot]);
  };
  int it = 0;
  if (w0 < ngroups) issue(w0, 0);
  for (int g = w0; g < ngroups; g += warps, ++it) {
    const int slot = it & 1;
    if (g + warps < ngroups) issue(g + warps, slot ^ 1);         // next group in flight while this one is processed
    g_mbar_wait(&bars[wib][slot], (uint32_t)(it >> 1) & 1u);
    const int r0 = g * G, n = min(G, P - r0);
    for (int k = 0; k < n; ++k) {
      const int i = r0 + k;
      RowVec<NC> x, y;
      row_load(x, ring + ((size_t)slot * G + k) * d, d, lane);
      row_dropout(x, dc, i, d, lane);
      float nn;
      if (mode == SRK_NORM_NISER && x_first) {
        nn = sqrtf(row_dot(x, x));
        y = x;
        row_scale(y, 1.f / (nn + 1e-12f));
        row_store(y, x_first + (long long)i * d, d, lane);
        float n1 = sqrtf(row_dot(y, y));
        row_scale(y, 1.f / n1);
      } else if (mode == SRK_NORM_NONE) {
        nn = 0.f;
        y = x;
      } else {
        nn = row_normalize(x, y, mode);
      }
      row_store(y, X + (long long)i * d, d, lane);
      if (rnorm && lane == 0) rnorm[i] = nn;
    }
    __syncwarp();                                                 // every lane has read the slot before it is refilled
  }
}

// Embedding-gradient scatter-add, load-balanced: the occurrence list is sorted by item id (perm / uoff / uid from the
// batch builder); every warp takes SCATTER_CHUNK consecutive occurrences, so a hot item (Zipf head, ~10% of a batch)
// is spread over many warps instead of serialising one.  Runs of one item that lie entirely inside a warp's chunk are
// added with a plain read-modify-write (deterministic); only items cut by a chunk boundary use atomicAdd.
constexpr int SCATTER_CHUNK = 16;           // large batches: long register-accumulated runs, few cut runs
constexpr int SCATTER_CHUNK_HUGE = 64;      // >= 2^18 occurrences (measured on 1 M: 0.47 / 0.53 / 0.57 / 0.59 / 0.60 of the HBM peak at 4 / 8 / 16 / 32 / 64)
constexpr int SCATTER_CHUNK_SMALL = 2;      // a training batch (P ~ 2 k): the kernel is a latency chain, not a bandwidth
                                            // problem - 4x more warps with 4x shorter per-warp loops

// where a finished run of one item goes: a run that lies entirely inside the warp's chunk is added to dE with a plain
// read-modify-write (nobody else touches that row in this launch); a run cut by a chunk boundary is written as a PARTIAL
// sum into the workspace (slot 0: the run began in an earlier chunk, slot 1: it began here and continues) and added up in
// chunk order by scatter_fixup_kernel - deterministic, no atomics.  Without a workspace cut runs fall back to atomicAdd.
template <int NC>
__device__ __forceinline__ void scatter_flush(const RowVec<NC>& acc, float* __restrict__ row, int d, int lane, bool whole,
                                              float* __restrict__ ws_slot) {
  if (whole) {
    row_add_store(acc, row, d, lane);
  } else if (ws_slot != nullptr) {
    row_store(acc, ws_slot, d, lane);
  } else {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      int col = (c * 32 + lane) * 4;
      if (col < d) {
        atomicAdd(row + col, acc.v[c].x); atomicAdd(row + col + 1, acc.v[c].y);
        atomicAdd(row + col + 2, acc.v[c].z); atomicAdd(row + col + 3, acc.v[c].w);
      }
    }
  }
}

template <int NC>
__global__ void __launch_bounds__(ROW_THREADS) scatter_bwd_kernel(const float* __restrict__ E, const int* __restrict__ perm,
                                                                  const int* __restrict__ uoff, const int* __restrict__ uid,
                                                                  int U, int P, int d, int mode, int chunk, DropCfg dc,
                                                                  const float* __restrict__ rnorm,
                                                                  const float* __restrict__ dX,
                                                                  const float* __restrict__ dX_first,
                                                                  float* __restrict__ dE, float* __restrict__ ws,
                                                                  int* __restrict__ owner) {
  SRK_PDL();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int nchunks = (P + chunk - 1) / chunk;
  // the table row is only needed to rebuild the normalised output for the normalise-backward: a plain embedding (SRGNN)
  // never reads E here, so the kernel moves exactly N (4 + 4d) + 2 U 4d bytes
  const bool need_x = mode != SRK_NORM_NONE;
  for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nchunks; w += warps) {
    const int j0 = w * chunk, j1 = min(P, j0 + chunk);
    int lo = 0, hi = U - 1;                    // distinct item u with uoff[u] <= j0 < uoff[u + 1]
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (uoff[mid] <= j0) lo = mid; else hi = mid - 1;
    }
    int u = lo;
    float* wsw = ws ? ws + (long long)w * 2 * d : nullptr;
    RowVec<NC> erow, acc, dy;
    int i = perm[j0];
    row_load(dy, dX + (long long)i * d, d, lane);               // software pipeline: the gradient row of occurrence j + 1 is
    if (need_x) row_load(erow, E + (long long)uid[u] * d, d, lane);   // in flight while occurrence j is being processed
    // A run that lies entirely inside this chunk starts from the CURRENT gradient row (loaded here, beside the other loads)
    // and ends with a plain store: no dependent read-modify-write at the end of the run.  Cut runs start from zero.
    bool whole = uoff[u] >= j0 && uoff[u + 1] <= j1;
    if (whole) row_load(acc, dE + (long long)uid[u] * d, d, lane); else row_zero(acc);
    for (int j = j0; j < j1; ++j) {
      while (j >= uoff[u + 1]) {
        if (whole) row_store(acc, dE + (long long)uid[u] * d, d, lane);
        else scatter_flush(acc, dE + (long long)uid[u] * d, d, lane, false, wsw);     // a run that ends here was cut at its head
        ++u;
        if (need_x) row_load(erow, E + (long long)uid[u] * d, d, lane);
        whole = uoff[u + 1] <= j1;                                // it begins inside the chunk
        if (whole) row_load(acc, dE + (long long)uid[u] * d, d, lane); else row_zero(acc);
      }
      RowVec<NC> dy_next;
      int i_next = i;
      if (j + 1 < j1) {
        i_next = perm[j + 1];
        row_load(dy_next, dX + (long long)i_next * d, d, lane);
      }
      RowVec<NC> dx;
      if (!need_x) {
        dx = dy;
      } else {
        RowVec<NC> x = erow, y;
        row_dropout(x, dc, i, d, lane);
        const float n = rnorm[i];
        y = x;
        if (mode == SRK_NORM_L2) row_scale(y, 1.f / fmaxf(n, 1e-12f));
        else if (mode == SRK_NORM_EPS) row_scale(y, 1.f / (n + 1e-12f));
        else row_scale(y, n > 0.f ? 1.f / n : 0.f);     // NISER: y2 = x / ||x|| up to rounding
        if (mode == SRK_NORM_NISER && dX_first) {
          RowVec<NC> d1;
          row_load(d1, dX_first + (long long)i * d, d, lane);
          row_normalize_bwd(x, y, n, mode, dy, &d1, dx);
        } else {
          row_normalize_bwd<NC>(x, y, n, mode, dy, nullptr, dx);
        }
      }
      row_dropout(dx, dc, i, d, lane);
      row_axpy(acc, 1.f, dx);
      if (j + 1 < j1) dy = dy_next;
      i = i_next;
    }
    const bool head_cut = uoff[u] < j0, tail_cut = uoff[u + 1] > j1;
    if (whole) row_store(acc, dE + (long long)uid[u] * d, d, lane);
    else scatter_flush(acc, dE + (long long)uid[u] * d, d, lane, false, wsw ? wsw + (head_cut ? 0 : d) : nullptr);
    // the chunk in which a cut run BEGINS owns it: the second pass starts from this record instead of searching
    if (owner && lane == 0) owner[w] = (!head_cut && tail_cut) ? u : -1;
  }
}

// second pass of the deterministic scatter-add: the chunk in which a cut run BEGINS owns it and adds up the partial sums of
// every chunk the run crosses, in chunk order, then adds the total to the item's gradient row
template <int NC>
__global__ void __launch_bounds__(ROW_THREADS) scatter_fixup_kernel(const int* __restrict__ uoff, const int* __restrict__ uid, int U,
                                                                    int P, int d, int chunk, const float* __restrict__ ws,
                                                                    const int* __restrict__ owner, float* __restrict__ dE) {
  SRK_PDL();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int nchunks = (P + chunk - 1) / chunk;
  for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nchunks; w += warps) {
    const int u = owner[w];                    // item whose run begins in chunk w and continues beyond it, or -1
    if (u < 0) continue;
    RowVec<NC> acc, t;
    row_load(acc, ws + ((long long)w * 2 + 1) * d, d, lane);
    const int end = uoff[u + 1];
    for (int w2 = w + 1; w2 < nchunks && (long long)w2 * chunk < end; ++w2) {
      row_load(t, ws + (long long)w2 * 2 * d, d, lane);
      row_axpy(acc, 1.f, t);
    }
    row_add_store(acc, dE + (long long)uid[u] * d, d, lane);
  }
}

template <int NC>
__global__ void __launch_bounds__(ROW_THREADS) renorm_rows_kernel(float* __restrict__ E, const int* __restrict__ uid, int U,
                                                                  int d, float max_norm) {
  SRK_PDL();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < U; u += warps) {
    float* p = E + (long long)(uid ? uid[u] : u) * d;
    RowVec<NC> x;
    row_load(x, p, d, lane);
    float n = sqrtf(row_dot(x, x));
    if (n > max_norm) {
      row_scale(x, (float)((double)max_norm / ((double)n + 1e-7)));
      row_store(x, p, d, lane);
    }
  }
}

template <int NC>
__global__ void __launch_bounds__(ROW_THREADS) catalog_prep_fwd_kernel(float* __restrict__ E, int V, int d, int mode,
                                                                       float max_norm, float* __restrict__ Ehat,
                                                                       float* __restrict__ enorm, float* __restrict__ Ehi,
                                                                       float* __restrict__ Elo, uint16_t* __restrict__ Bhi,
                                                                       uint16_t* __restrict__ Blo) {
  SRK_PDL();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < V; v += warps) {
    float* p = E + (long long)v * d;
    RowVec<NC> x, y;
    row_load(x, p, d, lane);
    if (max_norm > 0.f) {
      float n0 = sqrtf(row_dot(x, x));
      if (n0 > max_norm) {
        row_scale(x, (float)((double)max_norm / ((double)n0 + 1e-7)));
        row_store(x, p, d, lane);
      }
    }
    float n = row_normalize(x, y, mode);
    row_store(y, Ehat + (long long)v * d, d, lane);
    if (Ehi) {
      RowVec<NC> hi, lo;
      row_split_tf32(y, hi, lo);
      row_store(hi, Ehi + (long long)v * d, d, lane);
      row_store(lo, Elo + (long long)v * d, d, lane);
    }
    if (Bhi) row_store_bf16_split(y, Bhi + (long long)v * d, Blo + (long long)v * d, d, lane);
    if (lane == 0) enorm[v] = n;
  }
}

template <int NC>
__global__ void __launch_bounds__(ROW_THREADS) rownorm_fwd_kernel(const float* __restrict__ X, long long ldx, int R, int d,
                                                                  int mode, float* __restrict__ Y, long long ldy,
                                                                  float* __restrict__ rnorm, uint16_t* __restrict__ Bhi,
                                                                  uint16_t* __restrict__ Blo) {
  SRK_PDL();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < R; r += warps) {
    RowVec<NC> x, y;
    row_load(x, X + r * ldx, d, lane);
    float n = row_normalize(x, y, mode);
    row_store(y, Y + r * ldy, d, lane);
    if (rnorm && lane == 0) rnorm[r] = n;
    if (Bhi) row_store_bf16_split(y, Bhi + (long long)r * d, Blo + (long long)r * d, d, lane);     // dense [R, d] pair
  }
}

template <int NC>
__global__ void __launch_bounds__(ROW_THREADS) rownorm_bwd_kernel(const float* __restrict__ X, long long ldx,
                                                                  const float* __restrict__ Y, long long ldy,
                                                                  const float* __restrict__ rnorm,
                                                                  const float* __restrict__ dY, long long lddy, int R,
                                                                  int d, int mode, float* __restrict__ dX,
                                                                  long long lddx, int accumulate, int nparts,
                                                                  long long part_stride) {
  SRK_PDL();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < R; r += warps) {
    RowVec<NC> x, y, dy, dx;
    row_load(x, X + r * ldx, d, lane);
    row_load(y, Y + r * ldy, d, lane);
    row_load(dy, dY + r * lddy, d, lane);
    for (int pt = 1; pt < nparts; ++pt) {        // dY given as partial sums (flash CE backward: one per session tile)
      RowVec<NC> t;
      row_load(t, dY + pt * part_stride + r * lddy, d, lane);
      row_axpy(dy, 1.f, t);
    }
    row_normalize_bwd<NC>(x, y, rnorm[r], mode, dy, nullptr, dx);
    if (accumulate) row_add_store(dx, dX + r * lddx, d, lane);
    else row_store(dx, dX + r * lddx, d, lane);
  }
}

// SemanticExpander (msgifsr.py:32-45, reducer = mean) tail: pre = 0.5 * mean_t X[n, t, :] + 0.5 * h[n, :], then
// F.normalize (msgifsr.py:252-253).  X is the dropped gather [N, k, d], h the final GRU state.
template <int NC>
__global__ void __launch_bounds__(ROW_THREADS) expander_combine_fwd_kernel(const float* __restrict__ X, const float* __restrict__ h,
                                                                           int N, int k, int d, float* __restrict__ out,
                                                                           float* __restrict__ rnorm) {
  SRK_PDL();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < N; n += warps) {
    RowVec<NC> acc, x;
    row_zero(acc);
    for (int t = 0; t < k; ++t) {
      row_load(x, X + ((long long)n * k + t) * d, d, lane);
      row_axpy(acc, 1.f, x);
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      acc.v[c].x /= (float)k; acc.v[c].y /= (float)k; acc.v[c].z /= (float)k; acc.v[c].w /= (float)k;
    }
    row_load(x, h + (long long)n * d, d, lane);
    row_scale(acc, 0.5f);
    row_axpy(acc, 0.5f, x);
    RowVec<NC> y;
    float nn = row_normalize(acc, y, SRK_NORM_L2);
    row_store(y, out + (long long)n * d, d, lane);
    if (lane == 0) rnorm[n] = nn;
  }
}

// dpre = d F.normalize applied to dout (pre is rebuilt as out * max(rnorm, eps)); dh = 0.5 dpre; dX[n, t] = 0.5/k dpre
template <int NC>
__global__ void __launch_bounds__(ROW_THREADS) expander_combine_bwd_kernel(const float* __restrict__ out,
                                                                           const float* __restrict__ rnorm,
                                                                           const float* __restrict__ dout, int N, int k, int d,
                                                                           float* __restrict__ dh, float* __restrict__ dX) {
  SRK_PDL();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < N; n += warps) {
    RowVec<NC> y, dy, dx, pre;
    row_load(y, out + (long long)n * d, d, lane);
    row_load(dy, dout + (long long)n * d, d, lane);
    const float nn = rnorm[n];
    pre = y;
    row_scale(pre, fmaxf(nn, 1e-12f));
    row_normalize_bwd<NC>(pre, y, nn, SRK_NORM_L2, dy, nullptr, dx);
    row_scale(dx, 0.5f);
    row_store(dx, dh + (long long)n * d, d, lane);
    row_scale(dx, 1.f / (float)k);
    for (int t = 0; t < k; ++t) row_store(dx, dX + ((long long)n * k + t) * d, d, lane);
  }
}

inline int row_grid(long long rows) {
  long long g = (rows + 7) / 8;
  if (g < 1) g = 1;
  if (g > 148LL * 64) g = 148LL * 64;
  return (int)g;
}

}  // namespace

extern "C" int srk_embed_gather_fwd(const float* E, const int* iid, int P, int d, int norm_mode, const srk_dropout* drop,
                                    float* X, float* rnorm, float* x_first, void* stream) {
  SRK_TRY(srk_check_dim(d));
  if (P <= 0) return SRK_OK;
  SRK_REQUIRE(norm_mode == SRK_NORM_NONE || rnorm, "embed_gather_fwd: rnorm is required when normalising");
  DropCfg dc = make_drop(drop);
  static int use_tma = -1;              // SESSREC_GATHER_TMA=0: plain warp-per-row loads
  if (use_tma < 0) {
    const char* e = getenv("SESSREC_GATHER_TMA");
    use_tma = !(e && e[0] == '0');
  }
  // Staging pays where the gather is a latency problem (a training batch: a few thousand L2-resident rows; measured 24.6 ->
  // 10.2 us at the cfg2 shape); a gather that streams a table far larger than L2 is already at 98.6 % of the copy peak with
  // plain 128-bit loads and loses ~12 % to the shared-memory round trip, so it keeps the plain kernel.
  if (use_tma && P < (1 << 18) && (reinterpret_cast<uintptr_t>(E) & 15u) == 0) {
    // rows per group: the 8 warps' two-deep rings share GATHER_TMA_SMEM bytes (d = 96: 16 rows, d = 256: 6, d = 1024: 1),
    // but never so many that fewer than ~2 warps per SM scheduler are left with work
    int G = GATHER_TMA_SMEM / (8 * 2 * d * 4);
    if (G > 16) G = 16;
    const int gpar = P / (148 * 8 * 2);
    if (G > gpar) G = gpar < 1 ? 1 : gpar;
    if (G >= 1) {
      const size_t smem = (size_t)8 * 2 * G * d * 4;
      const int groups = (P + G - 1) / G;
      int grid = (groups + 7) / 8;
      if (grid > 148 * 2) grid = 148 * 2;
      static bool attr_set[9] = {false};  // > 48 KB of dynamic shared memory is an opt-in per kernel instantiation
      const int nci = (d + 127) / 128;
      if (!attr_set[nci]) {
        cudaError_t ae = cudaSuccess;
        SRK_DISPATCH_NC(d, (ae = cudaFuncSetAttribute(gather_tma_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, GATHER_TMA_SMEM)));
        SRK_CUDA(ae);
        attr_set[nci] = true;
      }
      SRK_DISPATCH_NC(d, (srk_launch(gather_tma_kernel<NC>, grid, ROW_THREADS, smem, (cudaStream_t)stream, E, iid, P, d, norm_mode, dc, G, X, rnorm, x_first)));
      SRK_LAUNCH_CHECK();
      return SRK_OK;
    }
  }
  SRK_DISPATCH_NC(d, (srk_launch(gather_fwd_kernel<NC>, row_grid(P), ROW_THREADS, 0, (cudaStream_t)stream, E, iid, P, d, norm_mode, dc, X, rnorm, x_first)));
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

static inline int scatter_chunk(int P) {
  static int forced = -1;                 // SESSREC_SCATTER_CHUNK=n: experiment switch
  if (forced < 0) {
    const char* e = getenv("SESSREC_SCATTER_CHUNK");
    forced = e ? atoi(e) : 0;
  }
  if (forced > 0) return forced;
  return P >= (1 << 18) ? SCATTER_CHUNK_HUGE : (P >= 65536 ? SCATTER_CHUNK : SCATTER_CHUNK_SMALL);
}

extern "C" long long srk_embed_scatter_ws_floats(int P, int d) {
  if (P <= 0) return 0;
  const int chunk = scatter_chunk(P);
  const long long nchunks = (P + chunk - 1) / chunk;
  return nchunks * 2 * d + (nchunks + 3) / 4 * 4;          // two partial rows per chunk + one owner record per chunk
}

extern "C" int srk_embed_scatter_bwd_ws(const float* E, const int* iid, const int* perm, const int* uoff, const int* uid,
                                        int U, int P, int d, int norm_mode, const srk_dropout* drop, const float* rnorm,
                                        const float* dX, const float* dX_first, float* dE, float* ws, void* stream) {
  (void)iid;
  SRK_TRY(srk_check_dim(d));
  if (U <= 0 || P <= 0) return SRK_OK;
  DropCfg dc = make_drop(drop);
  const int chunk = scatter_chunk(P);
  const int nchunks = (P + chunk - 1) / chunk;
  int* owner = ws ? reinterpret_cast<int*>(ws + (long long)nchunks * 2 * d) : nullptr;
  SRK_DISPATCH_NC(d, (srk_launch(scatter_bwd_kernel<NC>, row_grid(nchunks), ROW_THREADS, 0, (cudaStream_t)stream, E, perm, uoff, uid, U, P, d, norm_mode, chunk, dc, rnorm, dX, dX_first, dE, ws, owner)));
  SRK_LAUNCH_CHECK();
  if (ws != nullptr && nchunks > 1) {
    SRK_DISPATCH_NC(d, (srk_launch(scatter_fixup_kernel<NC>, row_grid(nchunks), ROW_THREADS, 0, (cudaStream_t)stream, uoff, uid, U, P, d, chunk, ws, owner, dE)));
    SRK_LAUNCH_CHECK();
  }
  return SRK_OK;
}

extern "C" int srk_embed_scatter_bwd(const float* E, const int* iid, const int* perm, const int* uoff, const int* uid,
                                     int U, int P, int d, int norm_mode, const srk_dropout* drop, const float* rnorm,
                                     const float* dX, const float* dX_first, float* dE, void* stream) {
  return srk_embed_scatter_bwd_ws(E, iid, perm, uoff, uid, U, P, d, norm_mode, drop, rnorm, dX, dX_first, dE, nullptr, stream);
}

extern "C" int srk_renorm_rows(float* E, const int* uid, int U, int d, float max_norm, void* stream) {
  SRK_TRY(srk_check_dim(d));
  if (U <= 0) return SRK_OK;
  SRK_DISPATCH_NC(d, (srk_launch(renorm_rows_kernel<NC>, row_grid(U), ROW_THREADS, 0, (cudaStream_t)stream, E, uid, U, d,
                                                                                                      max_norm)));
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_catalog_prep_fwd(float* E, int V, int d, int norm_mode, float max_norm, float* Ehat, float* enorm,
                                    float* Ehat_hi, float* Ehat_lo, uint16_t* Ebf_hi, uint16_t* Ebf_lo, void* stream) {
  SRK_TRY(srk_check_dim(d));
  SRK_REQUIRE(norm_mode == SRK_NORM_L2 || norm_mode == SRK_NORM_EPS, "catalog_prep: norm_mode must be L2 or EPS");
  SRK_DISPATCH_NC(d, (srk_launch(catalog_prep_fwd_kernel<NC>, row_grid(V), ROW_THREADS, 0, (cudaStream_t)stream, E, V, d, norm_mode, max_norm, Ehat, enorm, Ehat_hi, Ehat_lo, Ebf_hi, Ebf_lo)));
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_catalog_prep_bwd(const float* E, const float* Ehat, const float* enorm, const float* dEhat, int nparts,
                                    int V, int d, int norm_mode, float* dE, void* stream) {
  SRK_TRY(srk_check_dim(d));
  if (V <= 0) return SRK_OK;
  SRK_REQUIRE(nparts >= 1, "catalog_prep_bwd: nparts must be >= 1");
  SRK_DISPATCH_NC(d, (srk_launch(rownorm_bwd_kernel<NC>, row_grid(V), ROW_THREADS, 0, (cudaStream_t)stream, E, d, Ehat, d, enorm, dEhat, d, V, d, norm_mode, dE, d, 1, nparts, (long long)V * d)));
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_rownorm_fwd(const float* X, long long ldx, int R, int d, int norm_mode, float* Y, long long ldy,
                               float* rnorm, void* stream) {
  SRK_TRY(srk_check_dim(d));
  if (R <= 0) return SRK_OK;
  SRK_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0, "rownorm: row strides must be multiples of 4");
  SRK_DISPATCH_NC(d, (srk_launch(rownorm_fwd_kernel<NC>, row_grid(R), ROW_THREADS, 0, (cudaStream_t)stream, X, ldx, R, d, norm_mode,
                                 Y, ldy, rnorm, nullptr, nullptr)));
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_rownorm_split_fwd(const float* X, long long ldx, int R, int d, int norm_mode, float* Y, long long ldy,
                                     float* rnorm, uint16_t* Ybf_hi, uint16_t* Ybf_lo, void* stream) {
  SRK_TRY(srk_check_dim(d));
  if (R <= 0) return SRK_OK;
  SRK_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0, "rownorm: row strides must be multiples of 4");
  SRK_REQUIRE(Ybf_hi != nullptr && Ybf_lo != nullptr, "rownorm_split: the bf16 pair is required");
  SRK_DISPATCH_NC(d, (srk_launch(rownorm_fwd_kernel<NC>, row_grid(R), ROW_THREADS, 0, (cudaStream_t)stream, X, ldx, R, d, norm_mode,
                                 Y, ldy, rnorm, Ybf_hi, Ybf_lo)));
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_rownorm_bwd(const float* X, long long ldx, const float* Y, long long ldy, const float* rnorm,
                               const float* dY, long long lddy, int R, int d, int norm_mode, float* dX, long long lddx,
                               int accumulate, void* stream) {
  SRK_TRY(srk_check_dim(d));
  if (R <= 0) return SRK_OK;
  SRK_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0 && lddy % 4 == 0 && lddx % 4 == 0, "rownorm: strides must be multiples of 4");
  SRK_DISPATCH_NC(d, (srk_launch(rownorm_bwd_kernel<NC>, row_grid(R), ROW_THREADS, 0, (cudaStream_t)stream, X, ldx, Y, ldy, rnorm, dY, lddy, R, d, norm_mode, dX, lddx, accumulate, 1, 0)));
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_expander_combine_fwd(const float* X, const float* h, int N, int k, int d, float* out, float* rnorm,
                                        void* stream) {
  SRK_TRY(srk_check_dim(d));
  if (N <= 0) return SRK_OK;
  SRK_DISPATCH_NC(d, (srk_launch(expander_combine_fwd_kernel<NC>, row_grid(N), ROW_THREADS, 0, (cudaStream_t)stream, X, h, N, k, d, out,
                                                                                                               rnorm)));
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

extern "C" int srk_expander_combine_bwd(const float* out, const float* rnorm, const float* dout, int N, int k, int d,
                                        float* dh, float* dX, void* stream) {
  SRK_TRY(srk_check_dim(d));
  if (N <= 0) return SRK_OK;
  SRK_DISPATCH_NC(d, (srk_launch(expander_combine_bwd_kernel<NC>, row_grid(N), ROW_THREADS, 0, (cudaStream_t)stream, out, rnorm, dout, N, k, d, dh, dX)));
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}
