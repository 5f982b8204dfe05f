// TEST INFRASTRUCTURE ONLY: lets g++ compile the product's device-math headers (ramsesgpu_b200/csrc/
// mhd_device.cuh, hydro_device.cuh) for the HOST, so that the CPU test suite exercises the same source as
// the sm_100a kernels.  The MUFU seeds become correctly rounded values truncated to 20 bits (the Newton
// steps after them are the product's own); CUDA's integer/float reinterpretation intrinsics become memcpy.
// The product never includes this file and never defines RG_HOST_EMULATION.
#pragma once
#define RG_HOST_EMULATION 1
#include <cmath>
#include <cstdint>
#include <cstring>

static inline double rg_host_seed20(double v) {  // keep sign, exponent and the top 20 fraction bits
  uint64_t b;
  std::memcpy(&b, &v, 8);
  b &= ~((1ULL << 32) - 1);
  std::memcpy(&v, &b, 8);
  return v;
}
static inline int __double2hiint(double v) { uint64_t b; std::memcpy(&b, &v, 8); return (int)(b >> 32); }
static inline int __double2loint(double v) { uint64_t b; std::memcpy(&b, &v, 8); return (int)(b & 0xffffffffu); }
static inline double __hiloint2double(int hi, int lo) {
  uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double v;
  std::memcpy(&v, &b, 8);
  return v;
}
static inline long long __double_as_longlong(double v) { long long b; std::memcpy(&b, &v, 8); return b; }
static inline int __float_as_int(float v) { int b; std::memcpy(&b, &v, 4); return b; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
static inline float __saturatef(float x) { return x != x ? 0.0f : (x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x)); }
using std::signbit;
