// Launch layer: plain launch, capture (remember the graph node of every launch) or update (re-parameterise the node).
#include <stdlib.h>

#include "launch.cuh"

thread_local int g_srk_launch_rc = SRK_OK;
static thread_local SrkLaunchCtx* g_ctx = nullptr;

void srk_set_launch_ctx(SrkLaunchCtx* ctx) { g_ctx = ctx; }
SrkLaunchCtx* srk_get_launch_ctx() { return g_ctx; }
int srk_launch_mode() { return g_ctx ? g_ctx->mode : SRK_LAUNCH_DIRECT; }

int srk_launch_raw(const void* func, dim3 grid, dim3 block, size_t smem, cudaStream_t st, void** args,
                   const unsigned char* blob, size_t blob_bytes) {
  SrkLaunchCtx* c = g_ctx;
  if (c && c->mode == SRK_LAUNCH_SKIP) return SRK_OK;
  ++g_srk_launches;
  if (c && c->mode == SRK_LAUNCH_UPDATE) {
    if (c->failed) return SRK_OK;                       // the step will be re-run with plain launches
    if (c->cursor >= c->g->nodes.size() || c->g->nodes[c->cursor].func != func || c->cursor == c->fail_at) {
      c->failed = true;
      c->fail_reason = c->cursor >= c->g->nodes.size() ? 1 : (c->cursor == c->fail_at ? 3 : 2);
      return SRK_OK;
    }
    SrkGraphNode& nd = c->g->nodes[c->cursor];
    if (blob_bytes > 0 && nd.blob.size() == blob_bytes && memcmp(nd.blob.data(), blob, blob_bytes) == 0) {
      ++c->cursor;                                      // same grid, same arguments: the node is already right
      return SRK_OK;
    }
    cudaKernelNodeParams p;
    memset(&p, 0, sizeof(p));
    p.func = const_cast<void*>(func);
    p.gridDim = grid;
    p.blockDim = block;
    p.sharedMemBytes = (unsigned int)smem;
    p.kernelParams = args;
    const cudaError_t ue = cudaGraphExecKernelNodeSetParams(c->g->exec, c->g->nodes[c->cursor].node, &p);
    if (ue != cudaSuccess) {
      cudaGetLastError();
      c->failed = true;
      c->fail_reason = 4;
      c->fail_cuda = (int)ue;
      return SRK_OK;
    }
    nd.blob.assign(blob, blob + blob_bytes);
    ++c->updated;
    ++c->cursor;
    return SRK_OK;
  }
  // SESSREC_PDL: 0 off, 1 programmatic dependent launch for plain launches only, 2 (default) also while capturing a graph
  // (programmatic edges between the kernel nodes; cfg1 backward-half replay 0.3625 -> 0.3493 ms per step)
  static const int pdl = [] {
    const char* e = getenv("SESSREC_PDL");
    return e ? atoi(e) : 2;
  }();
  // A kernel that is scheduled early holds its SM resources while it waits for its producers: only launches whose
  // footprint is a fraction of the machine may, or they crowd out the side streams.  Measured (B200, r2x): no limit 0.350 /
  // 0.735 ms per step at cfg1 / cfg2, <= 75 776 threads (a quarter of the resident threads) and <= 4 MB of shared memory
  // 0.345 / 0.664, PDL off 0.360 / 0.690.
  static const long long pdl_max_threads = [] {
    const char* e = getenv("SESSREC_PDL_MAX_THREADS");
    return e ? atoll(e) : 75776LL;
  }();
  static const long long pdl_max_smem = [] {
    const char* e = getenv("SESSREC_PDL_MAX_SMEM_KB");
    return (e ? atoll(e) : 4096LL) * 1024;
  }();
  const long long ctas = (long long)grid.x * grid.y * grid.z;
  const bool small = ctas * block.x * block.y * block.z <= pdl_max_threads && ctas * (long long)smem <= pdl_max_smem;
  const bool capturing = c && c->mode == SRK_LAUNCH_CAPTURE;
  if (pdl >= (capturing ? 2 : 1) && small) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at;
    memset(&at, 0, sizeof(at));
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &at;
    cfg.numAttrs = 1;
    SRK_CUDA(cudaLaunchKernelExC(&cfg, func, args));
  } else {
    SRK_CUDA(cudaLaunchKernel(func, grid, block, args, smem, st));
  }
  if (capturing) {
    cudaStreamCaptureStatus status;
    unsigned long long id = 0;
    cudaGraph_t graph = nullptr;
    const cudaGraphNode_t* deps = nullptr;
    size_t ndeps = 0;
    if (cudaStreamGetCaptureInfo_v2(st, &status, &id, &graph, &deps, &ndeps) != cudaSuccess ||
        status != cudaStreamCaptureStatusActive || ndeps != 1) {
      cudaGetLastError();
      c->failed = true;                                 // not capturing after all: the caller discards the graph
      c->fail_reason = 5;
      return SRK_OK;
    }
    c->g->nodes.push_back(SrkGraphNode{deps[0], func, std::vector<unsigned char>(blob, blob + blob_bytes)});
  }
  return SRK_OK;
}

namespace {

__global__ void __launch_bounds__(256) zero_kernel(uint4* __restrict__ p, long long n16, unsigned char* __restrict__ tail, int ntail) {
  SRK_PDL();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n16) p[i] = make_uint4(0u, 0u, 0u, 0u);
  if (i < ntail) tail[i] = 0;
}
// two regions in one launch (gradient buffer + the step's zeroed scratch pool)
__global__ void __launch_bounds__(256) zero2_kernel(uint4* __restrict__ p0, long long n0, uint4* __restrict__ p1, long long n1) {
  SRK_PDL();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n0) p0[i] = make_uint4(0u, 0u, 0u, 0u);
  else if (i - n0 < n1) p1[i - n0] = make_uint4(0u, 0u, 0u, 0u);
}
__global__ void __launch_bounds__(256) copy_kernel(uint4* __restrict__ d, const uint4* __restrict__ s, long long n16) {
  SRK_PDL();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n16) d[i] = s[i];
}

__global__ void __launch_bounds__(256) zero2d_kernel(float* __restrict__ C, long long ldc, int rows, int cols) {
  SRK_PDL();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (long long)rows * cols) C[(i / cols) * ldc + i % cols] = 0.f;
}

}  // namespace

int srk_zero2d_async(float* C, long long ldc, int rows, int cols, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return SRK_OK;
  if (ldc == cols && (reinterpret_cast<uintptr_t>(C) & 15u) == 0) return srk_zero_async(C, sizeof(float) * (size_t)rows * cols, st);
  const long long n = (long long)rows * cols;
  srk_launch(zero2d_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, C, ldc, rows, cols);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

int srk_zero_async(void* p, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return SRK_OK;
  SRK_REQUIRE((reinterpret_cast<uintptr_t>(p) & 15u) == 0, "zero_async: pointer must be 16-byte aligned");
  const long long n16 = (long long)(bytes / 16);
  const int ntail = (int)(bytes % 16);
  const long long work = n16 > ntail ? n16 : ntail;
  srk_launch(zero_kernel, dim3((unsigned)((work + 255) / 256)), dim3(256), 0, st, reinterpret_cast<uint4*>(p), n16,
             reinterpret_cast<unsigned char*>(p) + n16 * 16, ntail);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

int srk_zero2_async(void* p0, size_t bytes0, void* p1, size_t bytes1, cudaStream_t st) {
  if (bytes0 % 16 != 0 || bytes1 % 16 != 0 || ((reinterpret_cast<uintptr_t>(p0) | reinterpret_cast<uintptr_t>(p1)) & 15u) != 0) {
    SRK_TRY(srk_zero_async(p0, bytes0, st));
    return srk_zero_async(p1, bytes1, st);
  }
  const long long n0 = (long long)(bytes0 / 16), n1 = (long long)(bytes1 / 16);
  if (n0 + n1 == 0) return SRK_OK;
  srk_launch(zero2_kernel, dim3((unsigned)((n0 + n1 + 255) / 256)), dim3(256), 0, st, reinterpret_cast<uint4*>(p0), n0,
             reinterpret_cast<uint4*>(p1), n1);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}

int srk_copy_async(void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return SRK_OK;
  SRK_REQUIRE(((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15u) == 0 && bytes % 16 == 0,
              "copy_async: 16-byte aligned buffers and size required");
  const long long n16 = (long long)(bytes / 16);
  srk_launch(copy_kernel, dim3((unsigned)((n16 + 255) / 256)), dim3(256), 0, st, reinterpret_cast<uint4*>(dst),
             reinterpret_cast<const uint4*>(src), n16);
  SRK_LAUNCH_CHECK();
  return SRK_OK;
}
