// How fast are warp-uniform constant-bank operands when the working set is tens of kilobytes?  Every warp walks its own
// window of a 48 kB __constant__ table (uniform index: LDCU / c[3][UR+imm] operands feeding DFMAs), 16 warps per SM;
// window = bytes each warp keeps re-reading.  Compared with the same DFMAs on a 512-byte window (all warps the same).
#include <cstdio>
#include <cuda_runtime.h>

__constant__ double c_tab[6144];

template <int CW>
__device__ __forceinline__ double body(const double *x, int win_doubles, int iters, int stride_doubles) {
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int base = CW * stride_doubles;  // compile-time warp role -> uniform
  for (int it = 0; it < iters; ++it) {
    for (int j = 0; j < win_doubles; j += 8) {
      const double xv = x[(j >> 3) & 31];
#pragma unroll
      for (int r = 0; r < 8; ++r) acc[r] = fma(xv, c_tab[base + j + r], acc[r]);
    }
  }
  double s = 0;
  for (int r = 0; r < 8; ++r) s += acc[r];
  return s;
}

__global__ void __launch_bounds__(512) k(const double *in, double *out, int win_doubles, int iters, int stride_doubles) {
  __shared__ double xs[32 * 33];
  for (int i = threadIdx.x; i < 32 * 33; i += blockDim.x) xs[i] = in[i % 64];
  __syncthreads();
  const double *x = xs + (threadIdx.x & 31);
  double v = 0;
  switch (threadIdx.x >> 5) {
#define C(n) case n: v = body<n>(x, win_doubles, iters, stride_doubles); break;
    C(0) C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12) C(13) C(14) C(15)
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = v;
}

int main() {
  double *in, *out, h[6144];
  for (int i = 0; i < 6144; ++i) h[i] = 1e-9 * i;
  cudaMemcpyToSymbol(c_tab, h, sizeof h);
  cudaMalloc(&in, 8 * 64);
  cudaMemset(in, 0, 8 * 64);
  cudaMalloc(&out, 8 * 148 * 512);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  struct { int win, stride; const char *name; } cfg[] = {{64, 0, "512 B, shared by all warps"}, {384, 384, "3 kB per warp, 48 kB per SM"},
                                                        {176, 352, "1.4 kB per warp (x2 blocks apart), 45 kB span"}, {384, 0, "3 kB, shared by all warps"}};
  for (auto c : cfg) {
    const int iters = 400000 / c.win;
    k<<<148, 512>>>(in, out, c.win, 10, c.stride);
    cudaEventRecord(e0);
    k<<<148, 512>>>(in, out, c.win, iters, c.stride);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double dfma = 148.0 * 512 * (double)iters * c.win;
    printf("%-48s %.1f lane-DFMA/clk/SM (%.3f ms)\n", c.name, dfma / (ms * 1e6) / 148 / 1.965, ms);
  }
  return 0;
}
