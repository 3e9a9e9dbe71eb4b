// Microbenchmark: FP64 FMA dependent-issue latency and per-SM throughput as a function of ILP and warps/SM (B200).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_latency dfma_latency.cu && ./dfma_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double *out, double a, double b, int iters, long long *cycles) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9 + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int ILP>
void run(int threads, double *out, long long *dc) {
  const int iters = 4096;
  k<ILP><<<148, threads>>>(out, 1.0000001, 1e-9, iters, dc);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<ILP><<<148, threads>>>(out, 1.0000001, 1e-9, iters, dc);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long c;
  cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  const double per_iter = (double)c / iters;
  const double dfma_per_clk_sm = (double)ILP * threads / per_iter;  // thread-DFMAs per cycle per SM
  printf("ILP %2d warps/SM %2d: %7.1f cycles per round of ILP DFMAs per warp -> %6.1f DFMA lanes/clk/SM, %.2f TFLOP/s\n", ILP,
         threads / 32, per_iter, dfma_per_clk_sm, 2.0 * ILP * threads * 148.0 * iters / (ms * 1e-3) / 1e12);
}

int main() {
  double *out;
  long long *dc;
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaMalloc(&dc, 8);
  for (int threads : {32, 128, 256, 512, 1024}) {
    run<1>(threads, out, dc);
    run<2>(threads, out, dc);
    run<4>(threads, out, dc);
    run<8>(threads, out, dc);
    run<16>(threads, out, dc);
  }
  return 0;
}
