// FP64 tensor-core (DMMA, mma.sync f64) throughput on sm_100a, alone and mixed with DFMA / LDS traffic.
// Question behind it: the 3-D stencil contraction of queries that share a cell is a small GEMM
// (queries x taps) x (taps x columns); DMMA gets its operand reuse inside the tensor core instead of through
// shared-memory broadcasts, which is what bounds eval_cellrun_kernel.  Is DMMA full-rate on B200?
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double *c, const double *a, const double *b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

// MODE 0: m8n8k4, NACC independent accumulators; 1: m16n8k8; 2: m8n8k4 + one DFMA per DMMA; 3: m8n8k4 with the B operand
// re-loaded from shared memory (LDS.64) before every DMMA
template <int MODE, int NACC>
__global__ void k(const double *in, double *out, int iters) {
  __shared__ double sh[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = in[i];
  __syncthreads();
  double c[NACC][4], a[4], b[4], f[NACC];
  for (int i = 0; i < NACC; ++i) {
    for (int j = 0; j < 4; ++j) c[i][j] = in[threadIdx.x + i + j];
    f[i] = in[threadIdx.x + 7 * i];
  }
  for (int j = 0; j < 4; ++j) {
    a[j] = in[threadIdx.x + 100 + j];
    b[j] = in[threadIdx.x + 200 + j];
  }
  const int lo = (threadIdx.x & 31) * 3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (MODE == 0) dmma884(c[i][0], c[i][1], a[i & 3], b[(i >> 2) & 3]);
      if (MODE == 1) dmma1688(c[i], a, b);
      if (MODE == 2) {
        dmma884(c[i][0], c[i][1], a[i & 3], b[(i >> 2) & 3]);
        f[i] = fma(f[i], a[i & 3], b[i & 3]);
      }
      if (MODE == 3) {
        const double bb = sh[(lo + i * 33 + (it & 7)) & 1023];
        dmma884(c[i][0], c[i][1], a[i & 3], bb);
      }
    }
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int NACC>
void run(const char *name, int warps, double fma_per_mma) {
  double *in, *out;
  cudaMalloc(&in, 8 * 8192);
  cudaMemset(in, 0, 8 * 8192);
  cudaMalloc(&out, 8 * 148 * 1024);
  const int iters = 4000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE, NACC><<<148, 32 * warps>>>(in, out, 100);
  cudaEventRecord(e0);
  k<MODE, NACC><<<148, 32 * warps>>>(in, out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double mma = 148.0 * warps * (double)NACC * iters;
  printf("%-34s acc %2d warps/SM %2d: %7.2f FMA/ns/SM = %6.1f FMA/clk/SM at 1.965 GHz, %.3f MMA/clk/SM (%.2f ms) %s\n", name, NACC, warps,
         mma * fma_per_mma / (ms * 1e6) / 148, mma * fma_per_mma / (ms * 1e6) / 148 / 1.965, mma / (ms * 1e6) / 148 / 1.965, ms,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(in);
  cudaFree(out);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<0, 8>("m8n8k4", w, 256);
    run<0, 16>("m8n8k4", w, 256);
    run<1, 8>("m16n8k8", w, 1024);
    run<2, 8>("m8n8k4 + 1 DFMA each", w, 256);
    run<3, 8>("m8n8k4, B from LDS.64", w, 256);
  }
  return 0;
}
