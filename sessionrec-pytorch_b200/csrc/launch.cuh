// Graph context of the launch layer (common.cuh): used by csrc/step.cu only.
#pragma once
#include <functional>
#include <vector>

#include "common.cuh"

struct SrkGraphNode {
  cudaGraphNode_t node;
  const void* func;
  std::vector<unsigned char> blob;      // launch configuration + arguments the node holds now (empty: unknown)
};

struct SrkStepGraph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  std::vector<SrkGraphNode> nodes;      // kernel nodes in launch order
  void destroy() {
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    exec = nullptr;
    graph = nullptr;
    nodes.clear();
  }
};

struct SrkLaunchCtx {
  int mode = SRK_LAUNCH_DIRECT;
  SrkStepGraph* g = nullptr;
  size_t cursor = 0;
  bool failed = false;                  // update pass: kernel sequence differs from the captured one
  // what the step switches to at its forward / backward boundary (csrc/step.cu) and on which stream a capture starts
  int mode_after_boundary = SRK_LAUNCH_DIRECT;
  cudaStream_t capture_stream = nullptr;
  bool capturing = false;
  // whole-step graph: the capture / update mode starts at srk_step_begin() (first launch of the step) instead of at the
  // forward / backward boundary
  bool whole = false;
  size_t updated = 0;                   // update pass: nodes whose parameters had to be rewritten
  size_t fail_at = (size_t)-1;          // test hook: pretend the kernel sequence differs at this node
  // why the pass failed (SESSREC_GRAPH_DEBUG): 1 more launches than nodes, 2 other kernel, 3 test hook, 4 node update
  // refused (fail_cuda = the CUDA error), 5 launch outside the capture
  int fail_reason = 0;
  int fail_cuda = 0;
  // Work the step wants enqueued with plain calls AFTER its captured / replayed part (on the same stream): the data-parallel
  // all-reduce and the optimizer behind it.  NCCL collectives inside a replayed graph were measured pathological at 8 ranks
  // (1.9 ms per step against 0.65 ms with the collective outside), so the graph ends before them.
  std::function<int()> tail;
};
SrkLaunchCtx* srk_get_launch_ctx();

// installs `ctx` for the calling thread (nullptr = plain launches)
void srk_set_launch_ctx(SrkLaunchCtx* ctx);
