// Microbenchmark: register-only sliding-window FIR inner loop (8 accumulators x 8 taps per group, 3 distinct register
// operands per DFMA) - is the FP64 pipe rate sustained with three register operands?  B200.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double *out, const double *in, int iters, long long *cycles) {
  double acc[8], win[8], w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] = 0.0;
    win[i] = in[threadIdx.x + 32 * i];
    w[i] = in[256 + i];
  }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (MODE == 0) acc[r] = fma(win[(u + r) % 8], w[u], acc[r]);      // 3 register operands, rotating window
        if (MODE == 1) acc[r] = fma(acc[r], w[u], win[r]);               // chain through operand a
        if (MODE == 2) acc[r] = fma(win[r], w[0], acc[r]);               // fixed operands
      }
      if (MODE == 0) win[u] = win[u] * 1.0000001;  // stands for the reload (keeps the window live / changing)
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i] + win[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>
void run(int threads, double *out, double *in, long long *dc) {
  const int iters = 2048;
  k<MODE><<<148, threads>>>(out, in, iters, dc);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  printf("mode %d warps/SM %2d: %.1f cycles per 64 DFMAs per warp -> %.1f DFMA lanes/clk/SM\n", MODE, threads / 32,
         (double)c / iters, 64.0 * threads / ((double)c / iters));
}

int main() {
  double *out, *in;
  long long *dc;
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaMalloc(&in, 4096 * 8);
  cudaMemset(in, 0, 4096 * 8);
  cudaMalloc(&dc, 8);
  for (int threads : {128, 256, 512}) {
    run<0>(threads, out, in, dc);
    run<1>(threads, out, in, dc);
    run<2>(threads, out, in, dc);
  }
  return 0;
}
