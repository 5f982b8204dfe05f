// History diagnostics on the device: total mass, div B, Maxwell and Reynolds stresses,
// magnetic pressure and mean field of a 3D MHD state, as two streaming reduction passes over the
// inner cells.  The reference copies the whole state to the host and reduces it there with serial
// loops (MHDRunBase.cpp:3311-3410 history_default, :3476-3620 history_mri); here only a few KB of
// partial sums leave the GPU.  The reduction order is fixed (grid-stride rows per block, shared-memory
// tree, partials summed by the host in block order), so the result does not depend on scheduling.
#include "kernel_common.cuh"
#include "kernels.h"

namespace rg {

namespace {

constexpr int HB = 256;  // threads per block

// pass 1: sums over the inner (j,k) rows of rho, u = mx/rho, v = my/rho for every column i (ghost columns
// included, like the reference's localMean): partial[block][3][isize]
template <typename T>
__global__ void __launch_bounds__(HB) k_history_columns(const __grid_constant__ KParams<T> P, const T* __restrict__ Uin,
                                                        double* __restrict__ partial) {
  const UView<T> U = uview(Uin, P);
  const int rows = P.ny * P.nz;
  for (int i = threadIdx.x; i < P.isize; i += HB) {
    double sr = 0.0, su = 0.0, sv = 0.0;
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
      const int j = P.gw + r % P.ny, k = P.gw + r / P.ny;
      const T d = U(ID, i, j, k);
      sr += (double)d;
      su += (double)(U(IU, i, j, k) / d);
      sv += (double)(U(IV, i, j, k) / d);
    }
    double* out = partial + (size_t)blockIdx.x * 3 * P.isize;
    out[i] = sr; out[P.isize + i] = su; out[2 * P.isize + i] = sv;
  }
}

// pass 2: per-block partial sums of (mass, maxwell, reynolds, 2*magp, sum Bx, sum By, sum Bz, divB); meanUV holds
// the y-z averaged velocities [2][isize] of pass 1
template <typename T>
__global__ void __launch_bounds__(HB) k_history_sums(const __grid_constant__ KParams<T> P, const T* __restrict__ Uin,
                                                     const double* __restrict__ meanUV, double* __restrict__ partial) {
  __shared__ double red[8][HB];
  const UView<T> U = uview(Uin, P);
  const int rows = P.ny * P.nz;
  double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const int j = P.gw + r % P.ny, k = P.gw + r / P.ny;
    for (int i = P.gw + threadIdx.x; i < P.isize - P.gw; i += HB) {
      const T d = U(ID, i, j, k), a = U(IA, i, j, k), b = U(IB, i, j, k), c = U(IC, i, j, k);
      const T ap = U(IA, i + 1, j, k), bp = U(IB, i, j + 1, k), cp = U(IC, i, j, k + 1);
      s[0] += (double)d;
      s[1] -= 0.25 * (double)((a + ap) * (b + bp));
      s[2] += (double)d * ((double)(U(IU, i, j, k) / d) - meanUV[i]) * ((double)(U(IV, i, j, k) / d) - meanUV[P.isize + i]);
      s[3] += 0.25 * ((double)((a + ap) * (a + ap)) + (double)((b + bp) * (b + bp)) + (double)((c + cp) * (c + cp)));
      s[4] += (double)a; s[5] += (double)b; s[6] += (double)c;
      s[7] += (double)((ap - a) / P.dx + (bp - b) / P.dy + (cp - c) / P.dz);
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) red[q][threadIdx.x] = s[q];
  __syncthreads();
  for (int w = HB / 2; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) {
#pragma unroll
      for (int q = 0; q < 8; ++q) red[q][threadIdx.x] += red[q][threadIdx.x + w];
    }
    __syncthreads();
  }
  if (threadIdx.x < 8) partial[(size_t)blockIdx.x * 8 + threadIdx.x] = red[threadIdx.x][0];
}

}  // namespace

template <typename T>
void HistoryKernels<T>::columnSums(const KParams<T>& P, const T* U, double* partial, int nBlocks, cudaStream_t s) {
  k_history_columns<T><<<nBlocks, HB, 0, s>>>(P, U, partial);
  launched();
}
template <typename T>
void HistoryKernels<T>::sums(const KParams<T>& P, const T* U, const double* meanUV, double* partial, int nBlocks,
                             cudaStream_t s) {
  k_history_sums<T><<<nBlocks, HB, 0, s>>>(P, U, meanUV, partial);
  launched();
}

template struct HistoryKernels<double>;
template struct HistoryKernels<float>;

}  // namespace rg
