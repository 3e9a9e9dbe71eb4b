// DFMA throughput by operand kind on sm_100a: (a) three distinct vector-register operands per instruction (the shape of
// the resample consumer: acc += x * w with x, w loaded from shared memory), (b) one operand from the constant bank /
// a uniform register (the smooth FIR's shape), (c) few distinct registers.  One CTA per SM, W warps per CTA.
#include <cstdio>
#include <cuda_runtime.h>

struct P { double w[64]; };

template <int MODE>
__global__ void k(const __grid_constant__ P p, const double *in, double *out, int iters) {
  double acc[16], x[4], w[8];
  for (int i = 0; i < 16; ++i) acc[i] = in[threadIdx.x + i];
  for (int i = 0; i < 4; ++i) x[i] = in[threadIdx.x + 32 + i];
  for (int i = 0; i < 8; ++i) w[i] = in[threadIdx.x + 64 + i];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (MODE == 0) acc[(u & 1) * 8 + r] = fma(x[u], w[r], acc[(u & 1) * 8 + r]);           // R, R, R, R all distinct
        if (MODE == 1) acc[(u & 1) * 8 + r] = fma(x[u], p.w[u * 8 + r], acc[(u & 1) * 8 + r]);  // constant-bank operand
        if (MODE == 2) acc[(u & 1) * 8 + r] = fma(x[0], x[0], acc[(u & 1) * 8 + r]);            // one source register
      }
  }
  double s = 0;
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int warps) {
  double *in, *out;
  cudaMalloc(&in, 8 * 4096);
  cudaMemset(in, 0, 8 * 4096);
  cudaMalloc(&out, 8 * 148 * 1024);
  P p;
  for (int i = 0; i < 64; ++i) p.w[i] = 1e-9 * i;
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<148, 32 * warps>>>(p, in, out, 100);
  cudaEventRecord(e0);
  k<MODE><<<148, 32 * warps>>>(p, in, out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double dfma = 148.0 * warps * 32 * 32.0 * iters;  // lane-DFMAs
  printf("%-28s warps/SM %2d: %.1f lane-DFMA/ns/SM = %.1f per clk at 1.965 GHz (%.2f ms)\n", name, warps, dfma / (ms * 1e6) / 148,
         dfma / (ms * 1e6) / 148 / 1.965, ms);
  cudaFree(in);
  cudaFree(out);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<0>("R,R,R,R distinct", w);
    run<1>("R,R,c[],R", w);
    run<2>("R,Rx,Rx,R", w);
  }
  return 0;
}
