// Helpers shared by the kernel translation units: scratch/state accessors, the thread-block tiling
// of one z plane (with transposed remainder blocks) and the barrier-free max reduction.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <stdexcept>
#include <string>

#include "kernels.h"
#include "mhd_device.cuh"

namespace rg {

extern unsigned long long g_launches;  // kernels launched by this library
extern int g_hydroTile;                // run-time knob "hydro_tile" (kernels_hydro3d.cu)
extern int g_hydroFused;               // run-time knob "hydro_fused" (kernels_hydro3d_fused.cu)
extern int g_hydroTma;                 // run-time knob "hydro_tma" (kernels_hydro3d_fused.cu)
extern int g_haloP2p;                  // run-time knob "halo_p2p" (run.cu)
extern int g_hydroRows;                // run-time knob "hydro_rows" (kernels_hydro3d_fused.cu)
extern int g_tileX;                    // run-time knob "tile_x" (32 | 64 | 128); tile_y = BX / tile_x

// Every launch wrapper calls this right after its <<<>>>: counts the launch(es) and turns a failed launch
// (no kernel image for the device, too much dynamic shared memory, a bad grid) into an exception instead
// of stale results behind an RG_OK (cudaGetLastError does not synchronise).
inline void launched(int n = 1) {
  g_launches += (unsigned long long)n;
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    throw std::runtime_error(std::string("CUDA kernel launch failed: ") + cudaGetErrorString(e));
}

// Per-device state (function attributes, SM count) is cached per device id: a process may hold handles on several
// devices (rg_create_distributed takes a device argument)
constexpr int MAX_DEVICES = 64;
inline int currentDevice() {
  int d = 0;
  cudaGetDevice(&d);
  return (d >= 0 && d < MAX_DEVICES) ? d : 0;
}
inline int smCount() {
  static int n[MAX_DEVICES] = {0};
  const int d = currentDevice();
  if (n[d] == 0) {
    cudaDeviceGetAttribute(&n[d], cudaDevAttrMultiProcessorCount, d);
    if (n[d] <= 0) n[d] = 148;
  }
  return n[d];
}

namespace {

// Scratch arrays are indexed with 32-bit element offsets (one IMAD + one IMAD.WIDE per load instead
// of a 64-bit add chain); the host keeps every scratch array below 2^31 elements by z-chunking.
template <typename T>
struct View {  // [comp][kk][j][i] accessor of a scratch array
  T* p;
  int plane, comp;  // isize*jsize, planes*plane
  int isize, kbase;
  __device__ __forceinline__ T& operator()(int c, int i, int j, int k) const {
    return p[c * comp + ((k - kbase) * plane + j * isize + i)];
  }
};
template <typename T, typename PT>
__host__ __device__ inline View<T> view(T* p, const PT& P, int planes, int kbase) {
  View<T> v;
  v.p = p;
  v.plane = P.isize * P.jsize;
  v.comp = v.plane * planes;
  v.isize = P.isize;
  v.kbase = kbase;
  return v;
}
template <typename T>
struct UView {  // the state array [var][k][j][i]; cells < 2^31, variables offset in 64 bits
  const T* p;
  size_t comp;
  int plane, isize;
  __device__ __forceinline__ T operator()(int v, int i, int j, int k) const {
    return __ldg(p + (size_t)v * comp + (k * plane + j * isize + i));
  }
};
template <typename T>
__host__ __device__ inline UView<T> uview(const T* p, const KParams<T>& P) {
  UView<T> v;
  v.p = p;
  v.plane = P.isize * P.jsize;
  v.comp = (size_t)v.plane * P.ksize;
  v.isize = P.isize;
  return v;
}

constexpr int BX = 128;  // threads per block (tile shapes: 32x4, 64x2, 128x1)
#define TX ((int)blockDim.x)
#define TY ((int)blockDim.y)

// Index space [i0, i0+ni) x [j0, j0+nj) of one plane -> thread blocks.  Full TX-wide tiles first;
// when the last tile would be mostly empty (ni % TX < 24, e.g. the nx+1 = 257 faces of a 256^3
// grid), its columns are handled by "remainder" blocks (blockIdx.x == ni / TX) whose threads are
// laid out transposed (r columns x BX/r rows), so that no warp runs with 1 active lane out of 32.
constexpr int REM_MAX = 24;
inline dim3 gridFor(int ni, int nj, int nk) {
  const int tx = g_tileX, ty = BX / g_tileX;
  const int nFull = ni / tx, r = ni % tx;
  if (r > 0 && r < REM_MAX) {
    const int R = BX / r;
    return dim3(nFull + 1, std::max((nj + ty - 1) / ty, (nj + R - 1) / R), nk);
  }
  return dim3((ni + tx - 1) / tx, (nj + ty - 1) / ty, nk);
}
__device__ __forceinline__ bool tileCoords(int i0, int ni, int j0, int nj, int& i, int& j) {
  const int nFull = ni / TX, r = ni - nFull * TX;
  if (r > 0 && r < REM_MAX && (int)blockIdx.x == nFull) {
    const int tid = threadIdx.y * TX + threadIdx.x, R = BX / r;
    const int jj = blockIdx.y * R + tid / r;
    i = i0 + nFull * TX + tid % r;
    j = j0 + jj;
    return tid < R * r && jj < nj;
  }
  const int ii = blockIdx.x * TX + threadIdx.x, jj = blockIdx.y * TY + threadIdx.y;
  i = i0 + ii;
  j = j0 + jj;
  return ii < ni && jj < nj;
}
inline dim3 blockShape() { return dim3(g_tileX, BX / g_tileX, 1); }


__device__ __forceinline__ void atomicMaxOrdered(unsigned long long* addr, double v) {
  atomicMax(addr, (unsigned long long)__double_as_longlong(v));
}
// warp max -> one atomicMax per warp into one of MAX_SLOTS slots (spreads the atomics over many L2
// addresses; the host or an NCCL all-reduce finishes the max).  No block barrier.
template <typename T>
__device__ __forceinline__ void reduceMaxToSlots(T v, unsigned long long* slots) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = dev::mx(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0 && v > T(0)) {
    const unsigned slot = (blockIdx.x * 29u + (threadIdx.x >> 5) + threadIdx.y * 7u + blockIdx.y * 37u + blockIdx.z * 101u) & (MAX_SLOTS - 1);
    atomicMaxOrdered(slots + slot, (double)v);
  }
}


}  // namespace
}  // namespace rg
