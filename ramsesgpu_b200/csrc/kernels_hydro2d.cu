// 2D Euler (hydro) Godunov step for sm_100a, FP32 and FP64: what the reference's own regression harness runs
// (test/test_run.sh.in:29-82 on test/makeConfigHydro.cpp:26-79, the 2D jet).
//   trace      : U (5-point) -> cons->prim on the fly, TVD slopes, half-step predictor (+ static gravity) -> W (12 comps)
//   fluxUpdate : per cell, the four face fluxes from W of the cell and its four neighbours (each face flux is
//                evaluated by both adjacent cells with identical inputs and code, hence bitwise identical and
//                conservative), conservative update, inverse dt of the new state; ghost cells keep the old values
// Reference: HydroRunGodunov.cpp:2437-2655 (godunov_unsplit_cpu_v1, TWO_D), trace.h:332-414 (trace_unsplit_2d),
// slope.h (slope_unsplit_hydro_2d), constoprim.h:43-71, riemann.h (riemann<NVAR_2D> = the 3D solvers with w = 0),
// HydroRunBase.cpp:386-399 (dt).  The state is [var][j][i] with var = rho, E, rho u, rho v; ghost width 2.
#include "hydro_device.cuh"
#include "kernel_common.cuh"
#include "kernels.h"

namespace rg {

namespace {

enum { G_R = 0, G_P, G_U, G_V, G_DX = 4, G_DY = 8 };  // slopes: (r, p, u, v) each
static_assert(G_DY + 4 == NW_HYDRO2D, "2D hydro W layout");

template <typename T>
struct View2 {  // [comp][j][i]
  T* p;
  int isize, plane;
  __device__ __forceinline__ T& operator()(int c, int i, int j) const { return p[c * plane + j * isize + i]; }
};

template <typename T>
__device__ __forceinline__ void prim2(const KParams<T>& P, const View2<const T>& U, int i, int j, T (&q)[4]) {
  T q5[5];
  dev::cons_to_prim_hydro(P, U(ID, i, j), U(IP, i, j), U(IU, i, j), U(IV, i, j), T(0), q5);
  q[ID] = q5[ID]; q[IP] = q5[IP]; q[IU] = q5[IU]; q[IV] = q5[IV];
}

// slope types 1 and 2 share one formula in 2D (slope_unsplit_hydro_2d)
template <typename T>
__device__ __forceinline__ T slope2(T st, T qm, T q0, T qp) {
  return (st == T(1) || st == T(2)) ? dev::limited_slope(st, qm, q0, qp) : T(0);
}

template <typename T>
__global__ void __launch_bounds__(BX) k_hydro2d_trace(const __grid_constant__ KParams<T> P, const T* __restrict__ Uin,
                                                      T* __restrict__ Wp, T dt) {
  int i, j;
  if (!tileCoords(1, P.isize - 2, 1, P.jsize - 2, i, j)) return;
  const View2<const T> U{Uin, P.isize, P.isize * P.jsize};
  const View2<T> W{Wp, P.isize, P.isize * P.jsize};
  T q[4], qxm[4], qxp[4], qym[4], qyp[4];
  prim2(P, U, i, j, q);
  prim2(P, U, i - 1, j, qxm); prim2(P, U, i + 1, j, qxp);
  prim2(P, U, i, j - 1, qym); prim2(P, U, i, j + 1, qyp);
  const T st = P.slope_type, h = T(0.5);
  T dx_[4], dy_[4];
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    dx_[v] = h * slope2(st, qxm[v], q[v], qxp[v]);
    dy_[v] = h * slope2(st, qym[v], q[v], qyp[v]);
  }
  const T dtdx = dt / P.dx, dtdy = dt / P.dy;
  const T r = q[ID], p = q[IP], u = q[IU], v = q[IV], g = P.gamma0;
  const T ir = dev::rcp(r);
  // half-step predictor, trace.h:367-378
  const T sr0 = (-u * dx_[ID] - dx_[IU] * r) * dtdx + (-v * dy_[ID] - dy_[IV] * r) * dtdy;
  const T su0 = (-u * dx_[IU] - dx_[IP] * ir) * dtdx + (-v * dy_[IU]) * dtdy;
  const T sv0 = (-u * dx_[IV]) * dtdx + (-v * dy_[IV] - dy_[IP] * ir) * dtdy;
  const T sp0 = (-u * dx_[IP] - dx_[IU] * g * p) * dtdx + (-v * dy_[IP] - dy_[IV] * g * p) * dtdy;
  // static gravity: half-step predictor on the traced velocities (reference HydroRunGodunov.cpp:2485-2497 adds it to
  // every face state; the face states are centre +/- slope, built by the flux kernel)
  T gpx = T(0), gpy = T(0);
  if (P.gravity) {  // uniform field, or the field of the cell (Keplerian disc)
    const int plane = P.isize * P.jsize, idx = j * P.isize + i;
    gpx = T(0.5) * dt * (P.gCell ? __ldg(P.gCell + idx) : P.gx);
    gpy = T(0.5) * dt * (P.gCell ? __ldg(P.gCell + plane + idx) : P.gy);
  }
  W(G_R, i, j) = r + sr0; W(G_P, i, j) = p + sp0; W(G_U, i, j) = u + su0 + gpx; W(G_V, i, j) = v + sv0 + gpy;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    W(G_DX + c, i, j) = dx_[c];
    W(G_DY + c, i, j) = dy_[c];
  }
}

// state at a face of cell (i,j): W centre +/- half slope along DIR, floors (trace.h:381-413), rotated so that .u is
// the velocity normal to the face
template <typename T, int DIR>
__device__ __forceinline__ dev::HState<T> face2(const KParams<T>& P, const View2<const T>& W, int i, int j, T sgn) {
  constexpr int S = (DIR == 0) ? G_DX : G_DY;
  dev::HState<T> s;
  s.r = dev::mx(P.smallr, W(G_R, i, j) + sgn * W(S + 0, i, j));
  s.p = dev::mx(P.smallp * s.r, W(G_P, i, j) + sgn * W(S + 1, i, j));
  const T u = W(G_U, i, j) + sgn * W(S + 2, i, j), v = W(G_V, i, j) + sgn * W(S + 3, i, j);
  s.u = (DIR == 0) ? u : v;
  s.v = (DIR == 0) ? v : u;
  s.w = T(0);
  return s;
}

// flux through the LOW face of cell (i,j) along DIR, in physical component order (rho, E, rho u, rho v)
template <typename T, int DIR, int RS>
__device__ __forceinline__ void low_flux2(const KParams<T>& P, const View2<const T>& W, int i, int j, T (&f)[4]) {
  const dev::HState<T> L = face2<T, DIR>(P, W, i - (DIR == 0), j - (DIR == 1), T(1));
  const dev::HState<T> R = face2<T, DIR>(P, W, i, j, T(-1));
  T fr[5];
  dev::riemann_hydro<RS>(P, L, R, fr);
  f[ID] = fr[ID]; f[IP] = fr[IP];
  f[IU] = (DIR == 0) ? fr[IU] : fr[IV];
  f[IV] = (DIR == 0) ? fr[IV] : fr[IU];
}

template <typename T, int RS>
__global__ void __launch_bounds__(BX) k_hydro2d_flux_update(const __grid_constant__ KParams<T> P, const T* __restrict__ Uold,
                                                            T* __restrict__ Unew, const T* __restrict__ Wp, T dt,
                                                            unsigned long long* __restrict__ slots) {
  int i, j;
  const bool valid = tileCoords(0, P.isize, 0, P.jsize, i, j);
  const int gw = P.gw;
  T invDt = T(0);
  if (valid) {
    const int plane = P.isize * P.jsize, idx = j * P.isize + i;
    T un[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) un[v] = __ldg(Uold + v * plane + idx);
    if (i >= gw && i < P.isize - gw && j >= gw && j < P.jsize - gw) {
      const View2<const T> W{Wp, P.isize, plane};
      const T dtdx = dt / P.dx, dtdy = dt / P.dy;
      T fxl[4], fyl[4], fxh[4], fyh[4];
      low_flux2<T, 0, RS>(P, W, i, j, fxl);
      low_flux2<T, 1, RS>(P, W, i, j, fyl);
      low_flux2<T, 0, RS>(P, W, i + 1, j, fxh);
      low_flux2<T, 1, RS>(P, W, i, j + 1, fyh);
#pragma unroll
      for (int v = 0; v < 4; ++v) {  // summation order of the reference's serial scatter (HydroRunGodunov.cpp:2590-2640)
        T s = un[v];
        s += fxl[v] * dtdx; s += fyl[v] * dtdy;
        s -= fxh[v] * dtdx; s -= fyh[v] * dtdy;
        un[v] = s;
      }
      if (P.gravity) {  // static gravity source term, reference HydroRunBase.cpp:1946-1958
        const T rs = __ldg(Uold + idx) + un[ID];
        un[IU] += T(0.5) * dt * (P.gCell ? __ldg(P.gCell + idx) : P.gx) * rs;
        un[IV] += T(0.5) * dt * (P.gCell ? __ldg(P.gCell + plane + idx) : P.gy) * rs;
      }
      T q[5];
      const T c = dev::cons_to_prim_hydro(P, un[ID], un[IP], un[IU], un[IV], T(0), q);
      invDt = (c + dev::ab(q[IU])) * P.rdx + (c + dev::ab(q[IV])) * P.rdy;
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) Unew[v * plane + idx] = un[v];
  }
  if (slots != nullptr) reduceMaxToSlots(invDt, slots);
}

template <typename T>
__global__ void __launch_bounds__(BX) k_hydro2d_invdt(const __grid_constant__ KParams<T> P, const T* __restrict__ Uin,
                                                      unsigned long long* __restrict__ slots) {
  int i, j;
  const bool valid = tileCoords(P.gw, P.nx, P.gw, P.ny, i, j);
  T invDt = T(0);
  if (valid) {
    const int plane = P.isize * P.jsize, idx = j * P.isize + i;
    T q[5];
    const T c = dev::cons_to_prim_hydro(P, __ldg(Uin + idx), __ldg(Uin + plane + idx), __ldg(Uin + 2 * plane + idx),
                                        __ldg(Uin + 3 * plane + idx), T(0), q);
    invDt = (c + dev::ab(q[IU])) * P.rdx + (c + dev::ab(q[IV])) * P.rdy;
  }
  reduceMaxToSlots(invDt, slots);
}

}  // namespace

template <typename T>
void Hydro2dKernels<T>::step(const KParams<T>& P, const T* Uold, T* Unew, T* W, T dt, unsigned long long* slots,
                             cudaStream_t s) {
  k_hydro2d_trace<T><<<gridFor(P.isize - 2, P.jsize - 2, 1), blockShape(), 0, s>>>(P, Uold, W, dt);
  launched();
  const dim3 g = gridFor(P.isize, P.jsize, 1);
  // one instantiation per Riemann solver: a single solver body in the kernel
  switch (P.riemannSolver) {
    case RS_HLLC: k_hydro2d_flux_update<T, RS_HLLC><<<g, blockShape(), 0, s>>>(P, Uold, Unew, W, dt, slots); break;
    case RS_HLL: k_hydro2d_flux_update<T, RS_HLL><<<g, blockShape(), 0, s>>>(P, Uold, Unew, W, dt, slots); break;
    case RS_APPROX: k_hydro2d_flux_update<T, RS_APPROX><<<g, blockShape(), 0, s>>>(P, Uold, Unew, W, dt, slots); break;
    default: k_hydro2d_flux_update<T, -1><<<g, blockShape(), 0, s>>>(P, Uold, Unew, W, dt, slots); break;
  }
  launched();
}

template <typename T>
void Hydro2dKernels<T>::computeInvDt(const KParams<T>& P, const T* U, unsigned long long* slots, cudaStream_t s) {
  k_hydro2d_invdt<T><<<gridFor(P.nx, P.ny, 1), blockShape(), 0, s>>>(P, U, slots);
  launched();
}

template struct Hydro2dKernels<double>;
template struct Hydro2dKernels<float>;

}  // namespace rg
