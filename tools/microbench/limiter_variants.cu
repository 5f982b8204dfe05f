// Static SASS comparison of half-slope limiter formulations (monotonised-central / minmod), FP64 (no GPU needed):
//   python tools/microbench/select_variants.py tools/microbench/limiter_variants.cu
// The trace kernel evaluates 36 limited slopes per cell: 45 % of its instructions
// (profiles/r01_g_fused_kernels_source_hotspots.txt).  All variants return HALF the limited slope,
//   0                                             if the one-sided differences a, b have opposite signs (or one is 0)
//   sign(a+b) * min(hst*|a|, hst*|b|, |a+b|/4)    otherwise        (hst = slope_type / 2)
#include <cuda_runtime.h>

__device__ __forceinline__ double mn(double a, double b) { return (a < b) ? a : b; }

// v0: the product's formulation (mhd_device.cuh half_slope)
__device__ __forceinline__ double half_slope_v0(double hst, double qm, double q0, double qp) {
  const double a = q0 - qm, b = qp - q0;
  const double s = a + b;
  const double m = mn(fabs(a), fabs(b)) * hst;
  const double c = fabs(s) * 0.25;
  double r = mn(m, c);
  if ((__double2hiint(a) ^ __double2hiint(b)) < 0) r = 0.0;
  return __hiloint2double((__double2hiint(r) & 0x7fffffff) | (__double2hiint(s) & 0x80000000), __double2loint(r));
}

// v1: flip everything into the frame where s >= 0 with integer XORs of the sign bit, take ONE three-way minimum
// of signed values and clamp it at zero with a sign mask: opposite signs give a negative minimum -> 0
__device__ __forceinline__ double half_slope_v1(double hst, double qm, double q0, double qp) {
  const double a = q0 - qm, b = qp - q0;
  const double s = a + b;
  const int sg = __double2hiint(s) & 0x80000000;
  const double fa = __hiloint2double(__double2hiint(a) ^ sg, __double2loint(a)) * hst;
  const double fb = __hiloint2double(__double2hiint(b) ^ sg, __double2loint(b)) * hst;
  const double fc = fabs(s) * 0.25;
  const double r = mn(mn(fa, fb), fc);
  const int hi = __double2hiint(r), m = ~(hi >> 31);   // r < 0 (or -0) -> +0
  return __hiloint2double((hi & m) | sg, __double2loint(r) & m);
}

// v2: like v1, minimum of the two one-sided terms in the integer domain is not possible (signed values), but for
// slope_type 2 (hst = 1) the multiply disappears
__device__ __forceinline__ double half_slope_v2(double qm, double q0, double qp) {
  const double a = q0 - qm, b = qp - q0;
  const double s = a + b;
  const int sg = __double2hiint(s) & 0x80000000;
  const double fa = __hiloint2double(__double2hiint(a) ^ sg, __double2loint(a));
  const double fb = __hiloint2double(__double2hiint(b) ^ sg, __double2loint(b));
  const double fc = fabs(s) * 0.25;
  const double r = mn(mn(fa, fb), fc);
  const int hi = __double2hiint(r), m = ~(hi >> 31);
  return __hiloint2double((hi & m) | sg, __double2loint(r) & m);
}

extern "C" __global__ void lim_v0(const double* x, double* y, double hst) { int i = threadIdx.x; y[i] = half_slope_v0(hst, x[i], x[i + 32], x[i + 64]); }
extern "C" __global__ void lim_v1(const double* x, double* y, double hst) { int i = threadIdx.x; y[i] = half_slope_v1(hst, x[i], x[i + 32], x[i + 64]); }
extern "C" __global__ void lim_v2(const double* x, double* y, double hst) { int i = threadIdx.x; y[i] = half_slope_v2(x[i], x[i + 32], x[i + 64]); }
