// Static SASS comparison of min/max/clamp formulations for the FP64 solvers (no GPU needed):
//   nvcc -gencode arch=compute_100a,code=sm_100a -cubin -o /tmp/sel.cubin tools/microbench/select_variants.cu
//   cuobjdump -sass /tmp/sel.cubin      (tools/microbench/select_variants.py prints the opcode counts)
// Background (profiles/r01_g_fused_kernels_source_hotspots.txt): 17 % of the instructions of the fused
// flux+emf+update kernel are compare-selects of doubles (DSETP on the FP64 pipe + 2 FSEL).  Candidates move the
// comparison to the integer pipe where the operands allow it:
//   clamp at zero      min(x, 0) / max(x, 0): sign-bit mask, no comparison at all
//   guard              max(x, tiny) for a radicand that is >= 0 analytically: compare the HIGH words as integers
//   max of non-negative values (wave speeds): 64-bit unsigned integer ordering == floating-point ordering
#include <cuda_runtime.h>

__device__ __forceinline__ double mx(double a, double b) { return (a > b) ? a : b; }
__device__ __forceinline__ double mn(double a, double b) { return (a < b) ? a : b; }

// ---- current formulations ----------------------------------------------------------------------
extern "C" __global__ void cur_min0(const double* x, double* y) { int i = threadIdx.x; y[i] = mn(x[i], 0.0); }
extern "C" __global__ void cur_max0(const double* x, double* y) { int i = threadIdx.x; y[i] = mx(x[i], 0.0); }
extern "C" __global__ void cur_guard(const double* x, double* y) { int i = threadIdx.x; y[i] = mx(x[i], 1e-300); }
extern "C" __global__ void cur_max4(const double* x, double* y) {
  int i = threadIdx.x;
  y[i] = mx(mx(x[i], x[i + 32]), mx(x[i + 64], x[i + 96]));
}

// ---- integer-pipe candidates --------------------------------------------------------------------
__device__ __forceinline__ double min0_int(double x) {  // x < 0 ? x : +0   (-0 -> -0, harmless: only used in products)
  const int hi = __double2hiint(x), m = hi >> 31;
  return __hiloint2double(hi & m, __double2loint(x) & m);
}
__device__ __forceinline__ double max0_int(double x) {  // x > 0 ? x : +0
  const int hi = __double2hiint(x), m = ~(hi >> 31);
  return __hiloint2double(hi & m, __double2loint(x) & m);
}
__device__ __forceinline__ double guard_int(double x) {  // max(x, 2^-996): one integer compare of the high words
  const int hi = __double2hiint(x);
  const bool small = hi < 0x01b00000;                   // also true for negative x and for +0
  return __hiloint2double(small ? 0x01b00000 : hi, small ? 0 : __double2loint(x));
}
__device__ __forceinline__ double max_nonneg(double a, double b) {  // a, b >= 0: unsigned integer ordering
  const unsigned long long ua = (unsigned long long)__double_as_longlong(a), ub = (unsigned long long)__double_as_longlong(b);
  return __longlong_as_double((long long)(ua > ub ? ua : ub));
}
extern "C" __global__ void int_min0(const double* x, double* y) { int i = threadIdx.x; y[i] = min0_int(x[i]); }
extern "C" __global__ void int_max0(const double* x, double* y) { int i = threadIdx.x; y[i] = max0_int(x[i]); }
extern "C" __global__ void int_guard(const double* x, double* y) { int i = threadIdx.x; y[i] = guard_int(x[i]); }
extern "C" __global__ void int_max4(const double* x, double* y) {
  int i = threadIdx.x;
  y[i] = max_nonneg(max_nonneg(x[i], x[i + 32]), max_nonneg(x[i + 64], x[i + 96]));
}
