// Build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o scripts/exp/graph_update_cost scripts/exp/graph_update_cost.cu
// Measured on B200 (r1s): 60 direct launches 162.7 us, 60 node updates + cudaGraphLaunch 59.6 us, cudaGraphLaunch alone 0.9 us.
// Host cost of (a) 60 direct kernel launches vs (b) 60 x cudaGraphExecKernelNodeSetParams + one cudaGraphLaunch.
#include <chrono>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
__global__ void k(float* p, int n, float a) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = p[i] * a + 1.f; }
int main() {
  float* d; cudaMalloc(&d, 1 << 20);
  cudaStream_t s; cudaStreamCreate(&s);
  const int NK = 60, IT = 200;
  for (int i = 0; i < 10; ++i) k<<<8, 256, 0, s>>>(d, 2048, 1.f);
  cudaStreamSynchronize(s);
  auto t0 = std::chrono::steady_clock::now();
  for (int it = 0; it < IT; ++it) { for (int i = 0; i < NK; ++i) k<<<8 + (it & 3), 256, 0, s>>>(d, 2048 + it, 1.f); }
  auto t1 = std::chrono::steady_clock::now();
  cudaStreamSynchronize(s);
  printf("direct: %.1f us per %d launches (host enqueue)\n", std::chrono::duration<double, std::micro>(t1 - t0).count() / IT, NK);
  cudaGraph_t g; cudaGraphExec_t ge;
  std::vector<cudaGraphNode_t> nodes;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < NK; ++i) {
    k<<<8, 256, 0, s>>>(d, 2048, 1.f);
    cudaStreamCaptureStatus st; unsigned long long id; cudaGraph_t cg; const cudaGraphNode_t* deps; size_t nd;
    cudaStreamGetCaptureInfo_v2(s, &st, &id, &cg, &deps, &nd);
    nodes.push_back(deps[0]);
  }
  cudaStreamEndCapture(s, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaGraphLaunch(ge, s); cudaStreamSynchronize(s);
  t0 = std::chrono::steady_clock::now();
  for (int it = 0; it < IT; ++it) {
    for (int i = 0; i < NK; ++i) {
      int n = 2048 + it; float a = 1.f; void* args[3] = {&d, &n, &a};
      cudaKernelNodeParams p = {}; p.func = (void*)k; p.gridDim = dim3(8 + (it & 3)); p.blockDim = dim3(256); p.kernelParams = args;
      cudaGraphExecKernelNodeSetParams(ge, nodes[i], &p);
    }
    cudaGraphLaunch(ge, s);
  }
  t1 = std::chrono::steady_clock::now();
  cudaStreamSynchronize(s);
  printf("graph update + launch: %.1f us per step (%d node updates), last error %s\n",
         std::chrono::duration<double, std::micro>(t1 - t0).count() / IT, NK, cudaGetErrorString(cudaGetLastError()));
  t0 = std::chrono::steady_clock::now();
  for (int it = 0; it < IT; ++it) cudaGraphLaunch(ge, s);
  t1 = std::chrono::steady_clock::now();
  cudaStreamSynchronize(s);
  printf("graph launch only: %.1f us per step\n", std::chrono::duration<double, std::micro>(t1 - t0).count() / IT);
  return 0;
}
