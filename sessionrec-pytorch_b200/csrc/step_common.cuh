// Shared scaffolding of the native training steps (csrc/step.cu: MSGIFSR order 1; csrc/step_srgnn.cu: SRGNN / NISER):
// stage timer, the prioritised side streams, the workspace bump allocator and the CUDA-graph replay driver.
#pragma once
#include <stdlib.h>

#include <functional>
#include <vector>

#include "launch.cuh"

// Debug aid (SESSREC_STEP_TIMING=1): CUDA events at the stage boundaries of the critical-path stream.  The events of
// a step are read back a few steps LATER (when they have completed anyway), so the host keeps running ahead of the GPU
// and the numbers are the steady-state in-pipeline stage times, which neither the cold-cache ncu launch list nor a
// synchronised step can give.  SESSREC_STEP_TIMING=2 synchronises after every step instead.
struct StageTimer {
  static constexpr int RING = 4, MAXEV = 24;
  struct Slot {
    cudaEvent_t ev[MAXEV];
    const char* names[MAXEV];
    int n = 0;
    bool made = false;
  };
  int mode;
  cudaStream_t st;
  Slot* cur = nullptr;
  static Slot* ring() {
    static Slot r[RING];
    return r;
  }
  static int& counter() {
    static int c = 0;
    return c;
  }
  static void print(Slot& s) {
    if (s.n < 2) return;
    float total = 0.f;
    if (cudaEventElapsedTime(&total, s.ev[0], s.ev[s.n - 1]) != cudaSuccess) return;
    fprintf(stderr, "[step timing] total %.1f us:", total * 1e3f);
    for (int i = 1; i < s.n; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, s.ev[i - 1], s.ev[i]);
      fprintf(stderr, " %s=%.1f", s.names[i], ms * 1e3f);
    }
    fprintf(stderr, "\n");
  }
  explicit StageTimer(cudaStream_t s) : st(s) {
    const char* e = getenv("SESSREC_STEP_TIMING");
    mode = (e && srk_launch_mode() == SRK_LAUNCH_DIRECT) ? atoi(e) : 0;
    if (!mode) return;
    cur = &ring()[counter() % RING];
    if (!cur->made) {
      for (int i = 0; i < MAXEV; ++i) cudaEventCreate(&cur->ev[i]);
      cur->made = true;
    } else if (mode == 1) {
      if (cudaEventQuery(cur->ev[cur->n - 1]) == cudaSuccess) print(*cur);     // the step recorded RING steps ago
    }
    cur->n = 0;
    ++counter();
    mark("start");
  }
  void mark(const char* name) {
    if (!mode || cur->n >= MAXEV) return;
    cudaEventRecord(cur->ev[cur->n], st);
    cur->names[cur->n] = name;
    ++cur->n;
  }
  void report() {
    if (mode != 2) return;
    cudaStreamSynchronize(st);
    print(*cur);
  }
};

// Side streams of the native step.  The session encoder is a tree of small, latency-bound kernels (N ~ 2 k nodes): the
// two GAT convolutions of a layer (graph / reversed graph), the weight-gradient GEMMs and the catalog backward are
// independent of each other, so they are enqueued on side streams (fork / join with events) and overlap on the 148 SMs.
// SESSREC_STREAMS=0 keeps everything on the caller's stream.
struct SideStreams {
  static constexpr int NS = 7, NE = 64;
  cudaStream_t s[NS];     // [0] critical path, [1..3] parallel encoder chains, [4] [5] weight gradients, [6] bulk catalog passes
  cudaEvent_t ev[NE];
  cudaEvent_t ev_cat;     // "catalog backward done" (recorded early, waited for late: not from the round-robin pool)
  int next = 0;
  bool ok = false;
  int init() {
    // The latency-bound encoder chain gets the highest priority: its few-CTA kernels must not queue behind the
    // thousands of CTAs of the catalog-wide passes (row normalisation backward, zero_grad) that run beside it.
    int least = 0, greatest = 0;
    SRK_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    const int mid = greatest + (least - greatest) / 2;
    const int prio[NS] = {greatest, greatest, greatest, greatest, mid, mid, least};
    for (int i = 0; i < NS; ++i) SRK_CUDA(cudaStreamCreateWithPriority(&s[i], cudaStreamNonBlocking, prio[i]));
    for (int i = 0; i < NE; ++i) SRK_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    SRK_CUDA(cudaEventCreateWithFlags(&ev_cat, cudaEventDisableTiming));
    ok = true;
    return SRK_OK;
  }
  // everything enqueued on `to` after this call waits for what is on `from` now
  int order(cudaStream_t from, cudaStream_t to) {
    const int mode = srk_launch_mode();
    if (mode == SRK_LAUNCH_UPDATE || mode == SRK_LAUNCH_SKIP) return SRK_OK;     // the graph already holds the edge
    return order_always(from, to);
  }
  int order_always(cudaStream_t from, cudaStream_t to) {
    if (from == to) return SRK_OK;
    cudaEvent_t e = ev[next];
    next = (next + 1) % NE;
    SRK_CUDA(cudaEventRecord(e, from));
    SRK_CUDA(cudaStreamWaitEvent(to, e, 0));
    return SRK_OK;
  }
};

struct Arena {
  uint8_t* base;
  size_t cap, off;
  bool ok;
  float* f(size_t n) { return reinterpret_cast<float*>(raw(n * sizeof(float))); }
  uint8_t* raw(size_t bytes) {
    size_t a = (off + 255) & ~(size_t)255;
    if (a + bytes > cap) { ok = false; off = a + bytes; return base; }
    off = a + bytes;
    return base + a;
  }
};


// one set of side streams per device, shared by every native step (defined in step.cu); nullptr when SESSREC_STREAMS=0
SideStreams* srk_side_streams();

// Runs `body(run_stream)` - a native step written as a sequence of srk_launch() calls that switches the launch context at
// its forward / backward boundary (srk_step_boundary) - on the step's own high-priority stream, ordered after / before
// `caller`.  want_graph: after two warm-up steps per `key` the backward half is captured ONCE into a CUDA graph; every
// later step only rewrites the kernel-node parameters and issues one cudaGraphLaunch (any mismatch: plain launches).
// whole: the graph starts at srk_step_begin() - the whole step is one graph - instead of at srk_step_boundary().  With inputs of
// constant shape at constant addresses (padded batches, see batch_builder.cu) an update pass then finds nearly every node
// unchanged (launch.cu compares the argument bytes) and the host cost of a step is the bookkeeping of its body plus one
// cudaGraphLaunch.
int srk_step_driver(cudaStream_t caller, unsigned long long key, int want_graph, const std::function<int(void*)>& body,
                    bool whole = false);
int srk_step_begin();
// srk_readout_bwd that also writes the TF32 hi / lo pair of du (csrc/readout_ce.cu)
int srk_readout_bwd_split(const float* F, float* u, float* v, const float* we, const int* seg, const int* last, const float* e,
                          const float* ms, const float* sr_in, const float* dsr_in, int B, int d, int with_last, float* dF,
                          float* dwe, float* duh, float* dul, void* stream);
// srk_tc_gemm with operands whose TF32 hi / lo pair already exists (csrc/umma_gemm.cu)
int srk_tc_gemm_pre(int form, int M, int N, int K, const float* A, long long lda, const float* Ahi, const float* Alo,
                    long long ldah, const float* B, long long ldb, const float* Bhi, const float* Blo, long long ldbh, float* C,
                    long long ldc, const float* bias, float alpha, int accumulate, int split_k, float* scratch, void* stream);
// srk_flash_ce_bwd without its dS memset launch when the caller zeroed dS itself (csrc/flash_ce.cu)
int srk_flash_ce_bwd_ex(int B, int V, int d, const uint16_t* Shi, const uint16_t* Slo, long long lds, const uint16_t* Ehi,
                        const uint16_t* Elo, long long lde, float scale, const int* labels, const float* lse, const float* gout,
                        float* dS, float* dEpart, int ds_zeroed, void* stream);
bool srk_step_want_whole(int phase, int padded);      // SESSREC_GRAPH_WHOLE policy
// forward / backward boundary of a step body: starts the capture (or switches to node updates) when the driver asked for it
int srk_step_boundary();
int srk_step_want_graph(int phase);       // SESSREC_GRAPH / srk_set_graph_mode policy: 0 plain, 1 replay, 2 measure and decide
