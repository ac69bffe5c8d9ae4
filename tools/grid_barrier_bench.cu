// What would a persistent "whole step in one kernel" design pay per layer boundary? Measures the latency of a grid-wide
// barrier (cooperative groups grid.sync()) with one CTA per SM (and two), against the cost of a kernel boundary inside a
// captured CUDA graph (a chain of empty kernels). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -rdc=true -o
// tools/grid_barrier_bench tools/grid_barrier_bench.cu ; run on a B200.
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void barrier_loop(int iters, unsigned long long* out) {
  cg::grid_group grid = cg::this_grid();
  unsigned long long t0 = clock64();
  for (int i = 0; i < iters; ++i) grid.sync();
  if (blockIdx.x == 0 && threadIdx.x == 0) *out = clock64() - t0;
}
__global__ void empty_kernel(int* p) { if (p && threadIdx.x == 1000) *p = 1; }

int main() {
  int dev = 0, sms = 0, coop = 0;
  cudaSetDevice(dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  printf("SMs %d, cooperative launch %d\n", sms, coop);
  unsigned long long* out;
  cudaMalloc(&out, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int per_sm = 1; per_sm <= 2; ++per_sm)
    for (int threads : {192, 256, 1024}) {
      int iters = 2000;
      void* args[] = {&iters, &out};
      dim3 grid(sms * per_sm), block(threads);
      cudaLaunchCooperativeKernel((void*)barrier_loop, grid, block, args, 0, 0);   // warm-up
      cudaDeviceSynchronize();
      cudaEventRecord(e0);
      cudaError_t err = cudaLaunchCooperativeKernel((void*)barrier_loop, grid, block, args, 0, 0);
      cudaEventRecord(e1);
      cudaDeviceSynchronize();
      if (err != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(err)); continue; }
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      unsigned long long cyc = 0;
      cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
      printf("grid.sync: %d CTAs x %4d threads: %.2f us per barrier (%.0f cycles)\n", sms * per_sm, threads, ms * 1e3 / iters,
             (double)cyc / iters);
    }
  // kernel boundary inside a CUDA graph: chain of 200 dependent empty kernels
  cudaStream_t s;
  cudaStreamCreate(&s);
  cudaGraph_t g;
  cudaGraphExec_t ge;
  for (int blocks : {1, 148, 592}) {
    cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
    for (int i = 0; i < 200; ++i) empty_kernel<<<blocks, 256, 0, s>>>(nullptr);
    cudaStreamEndCapture(s, &g);
    cudaGraphInstantiate(&ge, g, 0);
    cudaGraphLaunch(ge, s);
    cudaStreamSynchronize(s);
    cudaEventRecord(e0, s);
    for (int r = 0; r < 10; ++r) cudaGraphLaunch(ge, s);
    cudaEventRecord(e1, s);
    cudaStreamSynchronize(s);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("graph chain of empty kernels (%3d CTAs each): %.2f us per kernel boundary\n", blocks, ms * 1e3 / 2000);
    cudaGraphExecDestroy(ge);
    cudaGraphDestroy(g);
  }
  return 0;
}
