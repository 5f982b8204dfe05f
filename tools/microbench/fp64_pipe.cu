// FP64 pipe microbenchmark for sm_100a: dependent-issue latency and per-SM throughput of DFMA (and
// DMUL/DADD/DSETP+select, MUFU.RCP64H) as a function of resident warps per SM and independent
// chains per warp.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipe fp64_pipe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int OP>
__global__ void k(double* out, int iters, long long* cyc) {
  double a[ILP];
  const double b = 1.0000001, c = 1e-9;
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = 1.0 + threadIdx.x * 1e-6 + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        if (OP == 0) a[i] = fma(a[i], b, c);
        else if (OP == 1) a[i] = a[i] * b;
        else if (OP == 2) a[i] = a[i] + c;
        else if (OP == 3) a[i] = (a[i] > b) ? a[i] * b : b;            // DSETP + selects + DMUL
        else { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a[i])); a[i] = r; }
      }
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP, int OP>
void run(const char* name, double* out, long long* dcyc) {
  const int iters = 2048;
  for (int warps : {1, 4, 8, 16, 32}) {
    k<ILP, OP><<<148, warps * 32>>>(out, iters, dcyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<ILP, OP><<<148, warps * 32>>>(out, iters, dcyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long cyc; cudaMemcpy(&cyc, dcyc, sizeof cyc, cudaMemcpyDeviceToHost);
    const double ops = (double)iters * 8 * ILP;                 // per thread
    const double perSmPerClk = ops * warps * 32 / (double)cyc;  // thread-ops per clock per SM
    printf("%-6s ILP %d warps/SM %2d : %6.2f cycles per op per warp, %6.1f lane-ops/clk/SM, %.3f ms\n", name, ILP, warps,
           (double)cyc / ops, perSmPerClk, ms);
  }
}

int main() {
  double* out; long long* dcyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(double));
  cudaMalloc(&dcyc, sizeof(long long));
  run<1, 0>("DFMA", out, dcyc);
  run<2, 0>("DFMA", out, dcyc);
  run<4, 0>("DFMA", out, dcyc);
  run<8, 0>("DFMA", out, dcyc);
  run<1, 1>("DMUL", out, dcyc);
  run<1, 2>("DADD", out, dcyc);
  run<1, 3>("DSETP", out, dcyc);
  run<4, 3>("DSETP", out, dcyc);
  run<1, 4>("RCP64H", out, dcyc);
  run<4, 4>("RCP64H", out, dcyc);
  return 0;
}
