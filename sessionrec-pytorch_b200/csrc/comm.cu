// Collectives of the multi-GPU path (one process per GPU): a process-wide NCCL communicator owned by this library, so
// that the exchanges of a training step - the data-parallel gradient all-reduce, the per-session soft-max statistics and
// dS of the catalog-sharded head, the broadcast of the updated table rows - are enqueued from INSIDE the native step
// (csrc/step.cu, csrc/step_srgnn.cu) on the step's own streams, between its kernels, with no return to Python and inside
// the same CUDA-graph replay.  The reference is single-device; nothing here replaces reference code.
//
// NCCL is resolved at run time with dlopen (the library PyTorch bundles, libnccl.so.2, is already mapped into a process
// that imported torch; a path can be given explicitly) - libsessrec_b200.so has no link-time dependency on it, and a
// single-GPU run never touches it.  The 128-byte unique id is created on rank 0 (srk_comm_unique_id) and handed to the other
// ranks by the caller (torch.distributed broadcast in parallel.py): bootstrap only, the data path never goes through torch.
#include <dlfcn.h>

#include "common.cuh"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32 = 7 };
enum { ncclSum = 0, ncclMax = 2, ncclAvg = 4 };

struct Nccl {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};

Nccl g_nccl;
ncclComm_t g_comm = nullptr;
int g_rank = 0, g_world = 1;

int load_nccl(const char* path) {
  if (g_nccl.h) return SRK_OK;
  void* h = nullptr;
  if (path && path[0]) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);       // the copy torch already mapped
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  SRK_REQUIRE(h != nullptr, "comm: cannot load libnccl.so.2 (%s); pass its path to srk_comm_init", dlerror());
#define SRK_SYM(field, name)                                                            \
  *reinterpret_cast<void**>(&g_nccl.field) = dlsym(h, name);                            \
  SRK_REQUIRE(g_nccl.field != nullptr, "comm: libnccl.so.2 has no symbol %s", name)
  SRK_SYM(GetUniqueId, "ncclGetUniqueId");
  SRK_SYM(CommInitRank, "ncclCommInitRank");
  SRK_SYM(CommDestroy, "ncclCommDestroy");
  SRK_SYM(AllReduce, "ncclAllReduce");
  SRK_SYM(AllGather, "ncclAllGather");
  SRK_SYM(Broadcast, "ncclBroadcast");
  SRK_SYM(GroupStart, "ncclGroupStart");
  SRK_SYM(GroupEnd, "ncclGroupEnd");
  SRK_SYM(GetErrorString, "ncclGetErrorString");
  SRK_SYM(GetVersion, "ncclGetVersion");
#undef SRK_SYM
  g_nccl.h = h;
  return SRK_OK;
}

#define SRK_NCCL(expr)                                                                              \
  do {                                                                                              \
    ncclResult_t r__ = (expr);                                                                      \
    if (r__ != 0) {                                                                                 \
      srk_set_error("%s:%d NCCL error %d (%s)", __FILE__, __LINE__, r__, g_nccl.GetErrorString(r__)); \
      return SRK_ERR_CUDA;                                                                          \
    }                                                                                               \
  } while (0)

}  // namespace

/* rank 0: fills id128 (128 bytes) for srk_comm_init on every rank.  nccl_path may be NULL. */
extern "C" int srk_comm_unique_id(char* id128_host, const char* nccl_path) {
  SRK_TRY(load_nccl(nccl_path));
  ncclUniqueId id;
  SRK_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128_host, id.internal, 128);
  return SRK_OK;
}

/* Collective over all ranks: creates the communicator of this process on the CURRENT CUDA device. */
extern "C" int srk_comm_init(const char* id128_host, int rank, int world, const char* nccl_path) {
  SRK_REQUIRE(world >= 1 && rank >= 0 && rank < world, "comm: bad rank %d / world %d", rank, world);
  SRK_TRY(load_nccl(nccl_path));
  if (g_comm) {
    g_nccl.CommDestroy(g_comm);
    g_comm = nullptr;
  }
  ncclUniqueId id;
  memcpy(id.internal, id128_host, 128);
  SRK_NCCL(g_nccl.CommInitRank(&g_comm, world, id, rank));
  g_rank = rank;
  g_world = world;
  return SRK_OK;
}

extern "C" int srk_comm_destroy(void) {
  if (g_comm) {
    g_nccl.CommDestroy(g_comm);
    g_comm = nullptr;
  }
  g_rank = 0;
  g_world = 1;
  return SRK_OK;
}

extern "C" int srk_comm_world(void) { return g_comm ? g_world : 1; }
extern "C" int srk_comm_rank(void) { return g_comm ? g_rank : 0; }
extern "C" int srk_comm_nccl_version(void) {
  int v = 0;
  if (g_nccl.h && g_nccl.GetVersion) g_nccl.GetVersion(&v);
  return v;
}

/* In-place fp32 all-reduce (op 0 = sum, 1 = max, 2 = average) of buf[n] on `stream`. */
extern "C" int srk_comm_allreduce(float* buf, long long n, int op, void* stream) {
  if (n <= 0) return SRK_OK;
  // inside a native step whose backward half is being replayed from a CUDA graph the collective is a node of that graph
  // (captured once, same buffer every step): the update / skip passes of the launch layer must not enqueue it again
  if (srk_launch_mode() == SRK_LAUNCH_UPDATE || srk_launch_mode() == SRK_LAUNCH_SKIP) return SRK_OK;
  SRK_REQUIRE(g_comm != nullptr, "comm: no communicator (call srk_comm_init on every rank first)");
  SRK_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat32, op == 1 ? ncclMax : (op == 2 ? ncclAvg : ncclSum), g_comm, (cudaStream_t)stream));
  ++g_srk_launches;
  return SRK_OK;
}

/* recv[world * n] <- every rank's send[n] (rank order). */
extern "C" int srk_comm_allgather(const float* send, float* recv, long long n, void* stream) {
  if (n <= 0) return SRK_OK;
  if (srk_launch_mode() == SRK_LAUNCH_UPDATE || srk_launch_mode() == SRK_LAUNCH_SKIP) return SRK_OK;
  SRK_REQUIRE(g_comm != nullptr, "comm: no communicator (call srk_comm_init on every rank first)");
  SRK_NCCL(g_nccl.AllGather(send, recv, (size_t)n, ncclFloat32, g_comm, (cudaStream_t)stream));
  ++g_srk_launches;
  return SRK_OK;
}

/* Rows of a row-sharded [rows, d] table, in place: rank r owns rows [r * rows / world ... ) in the balanced split of
 * parallel.shard_slice (the first rows % world ranks hold one more) and broadcasts them to every other rank - ONE grouped
 * NCCL call (uneven shards allowed).  After it every replica holds every owner's rows. */
extern "C" int srk_comm_share_rows(float* table, int rows, int d, void* stream) {
  if (rows <= 0 || g_world == 1) return SRK_OK;
  if (srk_launch_mode() == SRK_LAUNCH_UPDATE || srk_launch_mode() == SRK_LAUNCH_SKIP) return SRK_OK;
  SRK_REQUIRE(g_comm != nullptr, "comm: no communicator (call srk_comm_init on every rank first)");
  const int base = rows / g_world, rem = rows % g_world;
  SRK_NCCL(g_nccl.GroupStart());
  for (int r = 0; r < g_world; ++r) {
    const long long lo = (long long)r * base + (r < rem ? r : rem);
    const long long cnt = base + (r < rem ? 1 : 0);
    if (cnt == 0) continue;
    float* p = table + lo * d;
    ncclResult_t rc = g_nccl.Broadcast(p, p, (size_t)(cnt * d), ncclFloat32, r, g_comm, (cudaStream_t)stream);
    if (rc != 0) {
      g_nccl.GroupEnd();
      srk_set_error("comm: ncclBroadcast failed (%s)", g_nccl.GetErrorString(rc));
      return SRK_ERR_CUDA;
    }
  }
  SRK_NCCL(g_nccl.GroupEnd());
  ++g_srk_launches;
  return SRK_OK;
}

// ---- catalog-sharded head: glue around the ONE exchange of per-session soft-max statistics ---------------------------------
namespace {

__global__ void __launch_bounds__(256) shard_labels_kernel(const int* __restrict__ labels, int B, int lo, int hi, int* __restrict__ out) {
  SRK_PDL();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    const int l = labels[b];
    out[b] = (l >= lo && l < hi) ? l - lo : -1;
  }
}

__global__ void __launch_bounds__(256) shard_pack_kernel(const float* __restrict__ lse_local, const float* __restrict__ nll_local,
                                                         const int* __restrict__ labels_local, const float* __restrict__ shift,
                                                         float bound, int B, float* __restrict__ pack) {
  SRK_PDL();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    const float l = lse_local[b];
    pack[b] = expf(l - (shift ? shift[b] : bound));
    pack[B + b] = labels_local[b] >= 0 ? l - nll_local[b] : 0.f;      // the label logit lives on exactly one rank
  }
}

__global__ void __launch_bounds__(256) shard_unpack_kernel(const float* __restrict__ pack, const float* __restrict__ shift, float bound,
                                                           int B, float* __restrict__ lse, float* __restrict__ nll) {
  SRK_PDL();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    const float l = (shift ? shift[b] : bound) + logf(pack[b]);
    lse[b] = l;
    nll[b] = l - pack[B + b];
  }
}

}  // namespace

/* labels_local[b] = labels[b] - lo when this rank's catalog rows [lo, hi) hold the label, else -1 */
extern "C" int srk_shard_labels(const int* labels, int B, int lo, int hi, int* labels_local, void* stream) {
  if (B <= 0) return SRK_OK;
  srk_launch(shard_labels_kernel, srk_cdiv(B, 256), 256, 0, (cudaStream_t)stream, labels, B, lo, hi, labels_local);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

/* pack[0:B] = exp(lse_local - shift), pack[B:2B] = label logit where owned else 0: the payload of the ONE sum all-reduce of
 * the sharded head.  shift: per-session (the all-reduced max of lse_local, unbounded logits) or NULL = the constant `bound`
 * (cosine heads: |logit| <= scale, so lse_local - scale is safe to exponentiate on every rank). */
extern "C" int srk_shard_lse_pack(const float* lse_local, const float* nll_local, const int* labels_local, const float* shift,
                                  float bound, int B, float* pack, void* stream) {
  if (B <= 0) return SRK_OK;
  srk_launch(shard_pack_kernel, srk_cdiv(B, 256), 256, 0, (cudaStream_t)stream, lse_local, nll_local, labels_local, shift, bound, B, pack);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

/* after the all-reduce: lse[b] = shift + log(pack[b]) (log-sum-exp over the whole catalog), nll[b] = lse[b] - pack[B + b] */
extern "C" int srk_shard_lse_unpack(const float* pack, const float* shift, float bound, int B, float* lse, float* nll, void* stream) {
  if (B <= 0) return SRK_OK;
  srk_launch(shard_unpack_kernel, srk_cdiv(B, 256), 256, 0, (cudaStream_t)stream, pack, shift, bound, B, lse, nll);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}
