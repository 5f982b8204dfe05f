// Dissipative terms of the 3D step for sm_100a: Ohmic resistivity and Navier-Stokes
// viscosity, applied to the NEW state after the Godunov update and a ghost refresh, as streaming
// stencil kernels (HBM-bound: a few reads and writes of the state per cell, no Riemann problems).
//
//   resistivity (eta > 0):  k_res_emf     U.B -> E[3] = -eta curl B at the cell edges      (scratch)
//                           k_res_ct      E -> U.B, constrained transport, in place
//                           k_res_energy  U.B -> Poynting flux eta J x B through the six faces of a cell,
//                                         U.E updated in place (reads B only, writes E only: no scratch)
//   viscosity   (nu > 0):   k_visc_flux   U -> viscous stress through the three low faces  (scratch, 12 comps)
//                           k_visc_update scratch -> U.(E, m) in place
//
// Reference: MHDRunBase.cpp:526-571 (emf), :302-345 (CT), :790-900 (energy flux),
// HydroRunBase.cpp:582-845 (viscous flux), :1504-1528 / :1675-1697 (updates); call sites
// mhd_godunov_unsplit_cpu_v3.cpp:661-693, MHDRunGodunov.cpp:3379-3419, HydroRunGodunov.cpp:2908-2927.
// The reference keeps three 8-component flux arrays and a 3-component emf array for this; here the
// resistive energy flux is recomputed by both cells of a face (identical code and inputs, so the
// update stays conservative bit for bit) and only the 12 non-zero viscous flux components are stored.
#include "kernel_common.cuh"
#include "kernels.h"

namespace rg {

namespace {

template <typename T>
struct SV {  // state array [var][k][j][i], 64-bit offsets
  const T* p;
  size_t plane, comp;
  int isize;
  __device__ __forceinline__ T operator()(int v, int i, int j, int k) const {
    return __ldg(p + (size_t)v * comp + (size_t)k * plane + (size_t)j * isize + i);
  }
};
template <typename T>
__device__ __forceinline__ SV<T> sview(const T* p, const KParams<T>& P) {
  SV<T> s;
  s.p = p;
  s.plane = (size_t)P.isize * P.jsize;
  s.comp = s.plane * P.ksize;
  s.isize = P.isize;
  return s;
}

// box [gw, size-gw] (first upper ghost index included) or inner cells [gw, size-gw) of plane k0+blockIdx.z
template <typename T>
__device__ __forceinline__ bool boxCoords(const KParams<T>& P, int extra, int k0, int& i, int& j, int& k) {
  k = k0 + blockIdx.z;
  return tileCoords(P.gw, P.nx + extra, P.gw, P.ny + extra, i, j);
}

// ---- resistivity ---------------------------------------------------------------------------------
// E component order: 0 = z, 1 = y, 2 = x (reference I_EMFZ, I_EMFY, I_EMFX)
template <typename T>
__global__ void __launch_bounds__(BX) k_res_emf(const __grid_constant__ KParams<T> P, const T* __restrict__ Uin,
                                                T* __restrict__ D) {
  int i, j, k;
  if (!boxCoords(P, 1, P.gw, i, j, k)) return;
  const SV<T> U = sview(Uin, P);
  const T a = U(IA, i, j, k), b = U(IB, i, j, k), c = U(IC, i, j, k);
  const T dbydx = (b - U(IB, i - 1, j, k)) / P.dx, dbzdx = (c - U(IC, i - 1, j, k)) / P.dx;
  const T dbxdy = (a - U(IA, i, j - 1, k)) / P.dy, dbzdy = (c - U(IC, i, j - 1, k)) / P.dy;
  const T dbxdz = (a - U(IA, i, j, k - 1)) / P.dz, dbydz = (b - U(IB, i, j, k - 1)) / P.dz;
  const size_t idx = (size_t)k * U.plane + (size_t)j * P.isize + i;
  D[idx] = -P.eta * (dbydx - dbxdy);
  D[U.comp + idx] = -P.eta * (dbxdz - dbzdx);
  D[2 * U.comp + idx] = -P.eta * (dbzdy - dbydz);
}

// constrained transport with the resistive emf: same un-guarded range as the main step, emfs one past
// the upper ghost face were never computed and read as zero (the reference's zero-initialised array)
template <typename T>
__global__ void __launch_bounds__(BX) k_res_ct(const __grid_constant__ KParams<T> P, T* __restrict__ Uio,
                                               const T* __restrict__ D, T dt) {
  int i, j, k;
  if (!boxCoords(P, 1, P.gw, i, j, k)) return;
  const int iN = P.isize - P.gw, jN = P.jsize - P.gw, kN = P.ksize - P.gw;
  const size_t plane = (size_t)P.isize * P.jsize, comp = plane * P.ksize;
  auto emf = [&](int c, int ii, int jj, int kk) -> T {
    return (ii > iN || jj > jN || kk > kN) ? T(0) : __ldg(D + (size_t)c * comp + (size_t)kk * plane + (size_t)jj * P.isize + ii);
  };
  const T dtdx = dt / P.dx, dtdy = dt / P.dy, dtdz = dt / P.dz;
  const size_t idx = (size_t)k * plane + (size_t)j * P.isize + i;
  T bx = Uio[IA * comp + idx], by = Uio[IB * comp + idx], bz = Uio[IC * comp + idx];
  const T ez = emf(0, i, j, k), ey = emf(1, i, j, k), ex = emf(2, i, j, k);
  if (k < kN) {
    bx += (emf(0, i, j + 1, k) - ez) * dtdy;
    by -= (emf(0, i + 1, j, k) - ez) * dtdx;
  }
  bx -= (emf(1, i, j, k + 1) - ey) * dtdz;
  by += (emf(2, i, j, k + 1) - ex) * dtdz;
  bz += (emf(1, i + 1, j, k) - ey) * dtdx;
  bz -= (emf(2, i, j + 1, k) - ex) * dtdy;
  Uio[IA * comp + idx] = bx; Uio[IB * comp + idx] = by; Uio[IC * comp + idx] = bz;
}

// Poynting flux -eta (J x B).n dt/d through the LOW face of cell (i,j,k) normal to DIR
template <typename T, int DIR>
__device__ __forceinline__ T res_energy_flux(const KParams<T>& P, const SV<T>& U, int i, int j, int k, T dt) {
  const T q = T(0.25), h = T(0.5);
  auto BX_ = [&](int a, int b, int c) { return U(IA, a, b, c); };
  auto BY_ = [&](int a, int b, int c) { return U(IB, a, b, c); };
  auto BZ_ = [&](int a, int b, int c) { return U(IC, a, b, c); };
  if (DIR == 0) {
    const T by = (BY_(i, j, k) + BY_(i - 1, j, k) + BY_(i, j + 1, k) + BY_(i - 1, j + 1, k)) * q;
    const T bz = (BZ_(i, j, k) + BZ_(i - 1, j, k) + BZ_(i, j, k + 1) + BZ_(i - 1, j, k + 1)) * q;
    T jy = (BX_(i, j, k) - BX_(i, j, k - 1)) / P.dz - (BZ_(i, j, k) - BZ_(i - 1, j, k)) / P.dx;
    T jp = (BX_(i, j, k + 1) - BX_(i, j, k)) / P.dz - (BZ_(i, j, k + 1) - BZ_(i - 1, j, k + 1)) / P.dx;
    jy = (jy + jp) * h;
    T jz = (BY_(i, j, k) - BY_(i - 1, j, k)) / P.dx - (BX_(i, j, k) - BX_(i, j - 1, k)) / P.dy;
    jp = (BY_(i, j + 1, k) - BY_(i - 1, j + 1, k)) / P.dx - (BX_(i, j + 1, k) - BX_(i, j, k)) / P.dy;
    jz = (jz + jp) * h;
    return -P.eta * (jy * bz - jz * by) * dt / P.dx;
  } else if (DIR == 1) {
    const T bx = (BX_(i, j, k) + BX_(i, j - 1, k) + BX_(i + 1, j, k) + BX_(i + 1, j - 1, k)) * q;
    const T bz = (BZ_(i, j, k) + BZ_(i, j - 1, k) + BZ_(i, j, k + 1) + BZ_(i, j - 1, k + 1)) * q;
    T jx = (BZ_(i, j, k) - BZ_(i, j - 1, k)) / P.dy - (BY_(i, j, k) - BY_(i, j, k - 1)) / P.dz;
    T jp = (BZ_(i, j, k + 1) - BZ_(i, j - 1, k + 1)) / P.dy - (BY_(i, j, k + 1) - BY_(i, j, k)) / P.dz;
    jx = (jx + jp) * h;
    T jz = (BY_(i, j, k) - BY_(i - 1, j, k)) / P.dx - (BX_(i, j, k) - BX_(i, j - 1, k)) / P.dy;
    jp = (BY_(i + 1, j, k) - BY_(i, j, k)) / P.dx - (BX_(i + 1, j, k) - BX_(i + 1, j - 1, k)) / P.dy;
    jz = (jz + jp) * h;
    return -P.eta * (jz * bx - jx * bz) * dt / P.dy;
  } else {
    const T bx = (BX_(i, j, k) + BX_(i, j, k - 1) + BX_(i + 1, j, k) + BX_(i + 1, j, k - 1)) * q;
    const T by = (BY_(i, j, k) + BY_(i, j, k - 1) + BY_(i, j + 1, k) + BY_(i, j + 1, k - 1)) * q;
    T jx = (BZ_(i, j, k) - BZ_(i, j - 1, k)) / P.dy - (BY_(i, j, k) - BY_(i, j, k - 1)) / P.dz;
    T jp = (BZ_(i, j + 1, k) - BZ_(i, j, k)) / P.dy - (BY_(i, j + 1, k) - BY_(i, j + 1, k - 1)) / P.dz;
    jx = (jx + jp) * h;
    T jy = (BX_(i, j, k) - BX_(i, j, k - 1)) / P.dz - (BZ_(i, j, k) - BZ_(i - 1, j, k)) / P.dx;
    jp = (BX_(i + 1, j, k) - BX_(i + 1, j, k - 1)) / P.dz - (BZ_(i + 1, j, k) - BZ_(i, j, k)) / P.dx;
    jy = (jy + jp) * h;
    return -P.eta * (jx * by - jy * bx) * dt / P.dz;
  }
}

template <typename T>
__global__ void __launch_bounds__(BX) k_res_energy(const __grid_constant__ KParams<T> P, T* __restrict__ Uio, T dt) {
  int i, j, k;
  if (!boxCoords(P, 0, P.gw, i, j, k)) return;
  const SV<T> U = sview(static_cast<const T*>(Uio), P);
  const size_t idx = (size_t)IP * U.comp + (size_t)k * U.plane + (size_t)j * P.isize + i;
  T e = Uio[idx];  // the kernel reads B and writes E only: no hazard between cells
  e += res_energy_flux<T, 0>(P, U, i, j, k, dt) - res_energy_flux<T, 0>(P, U, i + 1, j, k, dt);
  e += res_energy_flux<T, 1>(P, U, i, j, k, dt) - res_energy_flux<T, 1>(P, U, i, j + 1, k, dt);
  e += res_energy_flux<T, 2>(P, U, i, j, k, dt) - res_energy_flux<T, 2>(P, U, i, j, k + 1, dt);
  Uio[idx] = e;
}

// ---- viscosity -------------------------------------------------------------------------------------
// Viscous stress through the low face normal to N of cell (i,j,k): f[0..2] momentum, f[3] energy.
// Normal derivatives are two-point differences across the face, transverse ones the mean of the centred
// differences of the two cells (HydroRunBase.cpp:612-842).
template <typename T, int N>
__device__ __forceinline__ void visc_face(const KParams<T>& P, const SV<T>& U, int i, int j, int k, T dt, T (&f)[4]) {
  constexpr int oi = N == 0, oj = N == 1, ok = N == 2;
  const T dd[3] = {P.dx, P.dy, P.dz};
  auto vel = [&](int c, int a, int b, int g) -> T { return U(IU + c, a, b, g) / U(ID, a, b, g); };
  const T rho = T(0.5) * (U(ID, i, j, k) + U(ID, i - oi, j - oj, k - ok));
  T vR[3], vL[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) { vR[c] = vel(c, i, j, k); vL[c] = vel(c, i - oi, j - oj, k - ok); }
  T grad[3][3];  // grad[d][c] = d v_c / d x_d
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const int ti = d == 0, tj = d == 1, tk = d == 2;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (d == N) {
        grad[d][c] = (vR[c] - vL[c]) / dd[d];
      } else if ((N == 0 && ((d == 1 && c == 2) || (d == 2 && c == 1))) || (N == 1 && ((d == 0 && c == 2) || (d == 2 && c == 0))) ||
                 (N == 2 && ((d == 0 && c == 1) || (d == 1 && c == 0)))) {
        grad[d][c] = T(0);  // not part of the stress through this face
      } else {
        const T uR = vel(c, i + ti, j + tj, k + tk) + vel(c, i + ti - oi, j + tj - oj, k + tk - ok);
        const T uL = vel(c, i - ti, j - tj, k - tk) + vel(c, i - ti - oi, j - tj - oj, k - tk - ok);
        grad[d][c] = (uR - uL) / dd[d] * T(0.25);
      }
    }
  }
  const T two3rd = T(2.) / T(3.);
  constexpr int A = N == 0 ? 1 : 0, B = N == 2 ? 1 : 2;  // the two transverse directions
  T tau[3];
  tau[N] = -two3rd * P.nu * rho * (T(2) * grad[N][N] - grad[A][A] - grad[B][B]);
  tau[A] = -P.nu * rho * (grad[A > N ? A : N][A > N ? N : A] + grad[A > N ? N : A][A > N ? A : N]);
  tau[B] = -P.nu * rho * (grad[B > N ? B : N][B > N ? N : B] + grad[B > N ? N : B][B > N ? B : N]);
  const T s = dt / dd[N];
  f[0] = tau[0] * s; f[1] = tau[1] * s; f[2] = tau[2] * s;
  f[3] = T(0);
  if (!(P.cIso > T(0))) {
    const T u = T(0.5) * (vR[0] + vL[0]), v = T(0.5) * (vR[1] + vL[1]), w = T(0.5) * (vR[2] + vL[2]);
    f[3] = (u * tau[0] + v * tau[1] + w * tau[2]) * s;
  }
}

// D component order: [dir * 4 + {0: m_x, 1: m_y, 2: m_z, 3: energy}]
template <typename T>
__global__ void __launch_bounds__(BX) k_visc_flux(const __grid_constant__ KParams<T> P, const T* __restrict__ Uin,
                                                  T* __restrict__ D, T dt) {
  int i, j, k;
  if (!boxCoords(P, 1, P.gw, i, j, k)) return;
  const SV<T> U = sview(Uin, P);
  const size_t idx = (size_t)k * U.plane + (size_t)j * P.isize + i;
  T f[4];
  visc_face<T, 0>(P, U, i, j, k, dt, f);
#pragma unroll
  for (int c = 0; c < 4; ++c) D[(size_t)c * U.comp + idx] = f[c];
  visc_face<T, 1>(P, U, i, j, k, dt, f);
#pragma unroll
  for (int c = 0; c < 4; ++c) D[(size_t)(4 + c) * U.comp + idx] = f[c];
  visc_face<T, 2>(P, U, i, j, k, dt, f);
#pragma unroll
  for (int c = 0; c < 4; ++c) D[(size_t)(8 + c) * U.comp + idx] = f[c];
}

template <typename T>
__global__ void __launch_bounds__(BX) k_visc_update(const __grid_constant__ KParams<T> P, T* __restrict__ Uio,
                                                    const T* __restrict__ D) {
  int i, j, k;
  if (!boxCoords(P, 0, P.gw, i, j, k)) return;
  const size_t plane = (size_t)P.isize * P.jsize, comp = plane * P.ksize;
  const size_t idx = (size_t)k * plane + (size_t)j * P.isize + i;
  const int var[4] = {IU, IV, IW, IP};
#pragma unroll
  for (int c = 0; c < 4; ++c) {  // one direction after the other, like the reference's update
    T u = Uio[(size_t)var[c] * comp + idx];
    u += __ldg(D + (size_t)c * comp + idx) - __ldg(D + (size_t)c * comp + idx + 1);
    u += __ldg(D + (size_t)(4 + c) * comp + idx) - __ldg(D + (size_t)(4 + c) * comp + idx + P.isize);
    u += __ldg(D + (size_t)(8 + c) * comp + idx) - __ldg(D + (size_t)(8 + c) * comp + idx + plane);
    Uio[(size_t)var[c] * comp + idx] = u;
  }
}

}  // namespace

template <typename T>
void DissKernels<T>::resistEmf(const KParams<T>& P, const T* U, T* D, cudaStream_t s) {
  k_res_emf<T><<<gridFor(P.nx + 1, P.ny + 1, P.nz + 1), blockShape(), 0, s>>>(P, U, D);
  launched();
}
template <typename T>
void DissKernels<T>::ctUpdate(const KParams<T>& P, T* U, const T* D, T dt, cudaStream_t s) {
  k_res_ct<T><<<gridFor(P.nx + 1, P.ny + 1, P.nz + 1), blockShape(), 0, s>>>(P, U, D, dt);
  launched();
}
template <typename T>
void DissKernels<T>::resistEnergy(const KParams<T>& P, T* U, T dt, cudaStream_t s) {
  k_res_energy<T><<<gridFor(P.nx, P.ny, P.nz), blockShape(), 0, s>>>(P, U, dt);
  launched();
}
template <typename T>
void DissKernels<T>::viscFlux(const KParams<T>& P, const T* U, T* D, T dt, cudaStream_t s) {
  k_visc_flux<T><<<gridFor(P.nx + 1, P.ny + 1, P.nz + 1), blockShape(), 0, s>>>(P, U, D, dt);
  launched();
}
template <typename T>
void DissKernels<T>::viscUpdate(const KParams<T>& P, T* U, const T* D, cudaStream_t s) {
  k_visc_update<T><<<gridFor(P.nx, P.ny, P.nz), blockShape(), 0, s>>>(P, U, D);
  launched();
}

template struct DissKernels<double>;
template struct DissKernels<float>;

}  // namespace rg
