// 2D ideal-MHD Godunov step (reference implementation 1) for sm_100a:
//   prim -> trace (30-component traced state) -> {x/y fluxes, emf_z} -> update (+CT, +next dt)
// Reference: mhd_godunov_unsplit_cpu_v1.cpp:36-243, trace_mhd.h:38-339 (2D trace: pressure floor is
// smallp*rho, B_z advanced by fluxes), constoprim.h:389-420 (the z neighbour of B_z is 0 in 2D).
#include "kernel_common.cuh"
#include "kernels.h"
#include "mhd_device.cuh"

namespace rg {

namespace {

enum {
  V_R = 0, V_P, V_U, V_V, V_W, V_A, V_B, V_C,   // cell centred, advanced by dt/2
  V_AL, V_AR, V_BL, V_BR,                        // face fields, advanced by dt/2
  V_DRX, V_DPX, V_DUX, V_DVX, V_DWX, V_DBX, V_DCX,
  V_DRY, V_DPY, V_DUY, V_DVY, V_DWY, V_DAY, V_DCY,
  V_DALY, V_DARY, V_DBLX, V_DBRX
};
static_assert(V_DBRX + 1 == NW_MHD2D, "2D W layout");

template <typename T>
__global__ void __launch_bounds__(BX) k2_prim(const __grid_constant__ KParams<T> P, const T* __restrict__ Uin,
                                              T* __restrict__ Qp, T dt) {
  int i, j;
  if (!tileCoords(0, P.isize - 1, 0, P.jsize - 1, i, j)) return;
  const UView<T> U = uview(Uin, P);
  const View<T> Q = view(Qp, P, 1, 0);
  T u[8], q[8];
#pragma unroll
  for (int v = 0; v < 8; ++v) u[v] = U(v, i, j, 0);
  dev::cons_to_prim_mhd(P, u, U(IA, i + 1, j, 0), U(IB, i, j + 1, 0), T(0), dt, q);
#pragma unroll
  for (int v = 0; v < 8; ++v) Q(v, i, j, 0) = q[v];
}

template <typename T>
__global__ void __launch_bounds__(BX) k2_trace(const __grid_constant__ KParams<T> P, const T* __restrict__ Uin,
                                               const T* __restrict__ Qp, T* __restrict__ Wp, T dt) {
  const int gw = P.gw;
  int i, j;
  if (!tileCoords(gw - 1, P.isize - 2 * gw + 2, gw - 1, P.jsize - 2 * gw + 2, i, j)) return;
  const UView<T> U = uview(Uin, P);
  const View<const T> Q = view<const T>(Qp, P, 1, 0);
  const View<T> W = view(Wp, P, 1, 0);
  const T st = P.slope_type, h = T(0.5);
  const T dtdx = dt / P.dx, dtdy = dt / P.dy;
  T q[8], dx_[8], dy_[8];
#pragma unroll
  for (int v = 0; v < 8; ++v) {
    q[v] = Q(v, i, j, 0);
    dx_[v] = (st == T(0)) ? T(0) : h * dev::limited_slope(st, Q(v, i - 1, j, 0), q[v], Q(v, i + 1, j, 0));
    dy_[v] = (st == T(0)) ? T(0) : h * dev::limited_slope(st, Q(v, i, j - 1, 0), q[v], Q(v, i, j + 1, 0));
  }
  T AL = U(IA, i, j, 0), AR = U(IA, i + 1, j, 0), BL = U(IB, i, j, 0), BR = U(IB, i, j + 1, 0);
  const T dALy = h * dev::limited_slope(st, U(IA, i, j - 1, 0), AL, U(IA, i, j + 1, 0));
  const T dARy = h * dev::limited_slope(st, U(IA, i + 1, j - 1, 0), AR, U(IA, i + 1, j + 1, 0));
  const T dBLx = h * dev::limited_slope(st, U(IB, i - 1, j, 0), BL, U(IB, i + 1, j, 0));
  const T dBRx = h * dev::limited_slope(st, U(IB, i - 1, j + 1, 0), BR, U(IB, i + 1, j + 1, 0));
  auto Ez = [&](int ii, int jj) {
    const T u = T(0.25) * (Q(IU, ii - 1, jj - 1, 0) + Q(IU, ii - 1, jj, 0) + Q(IU, ii, jj - 1, 0) + Q(IU, ii, jj, 0));
    const T v = T(0.25) * (Q(IV, ii - 1, jj - 1, 0) + Q(IV, ii - 1, jj, 0) + Q(IV, ii, jj - 1, 0) + Q(IV, ii, jj, 0));
    const T A = h * (U(IA, ii, jj - 1, 0) + U(IA, ii, jj, 0));
    const T B = h * (U(IB, ii - 1, jj, 0) + U(IB, ii, jj, 0));
    return u * B - v * A;
  };
  const T ELL = Ez(i, j), ELR = Ez(i, j + 1), ERL = Ez(i + 1, j), ERR = Ez(i + 1, j + 1);
  const T r = q[ID], p = q[IP], u = q[IU], v = q[IV], w = q[IW], A = q[IA], B = q[IB], C = q[IC];
  const T drx = dx_[ID], dpx = dx_[IP], dux = dx_[IU], dvx = dx_[IV], dwx = dx_[IW], dBx = dx_[IB], dCx = dx_[IC];
  const T dry = dy_[ID], dpy = dy_[IP], duy = dy_[IU], dvy = dy_[IV], dwy = dy_[IW], dAy = dy_[IA], dCy = dy_[IC];
  const T dAx = h * (AR - AL), dBy = h * (BR - BL);
  const T ir = dev::rcp(r), g = P.gamma0;
  // trace_mhd.h:214-227
  const T sr0 = (-u * drx - dux * r) * dtdx + (-v * dry - dvy * r) * dtdy;
  const T su0 = (-u * dux - (dpx + B * dBx + C * dCx) * ir) * dtdx + (-v * duy + B * dAy * ir) * dtdy;
  const T sv0 = (-u * dvx + A * dBx * ir) * dtdx + (-v * dvy - (dpy + A * dAy + C * dCy) * ir) * dtdy;
  const T sw0 = (-u * dwx + A * dCx * ir) * dtdx + (-v * dwy + B * dCy * ir) * dtdy;
  const T sp0 = (-u * dpx - dux * g * p) * dtdx + (-v * dpy - dvy * g * p) * dtdy;
  const T sA0 = (u * dBy + B * duy - v * dAy - A * dvy) * dtdy;
  const T sB0 = (-u * dBx - B * dux + v * dAx + A * dvx) * dtdx;
  T sC0 = (w * dAx + A * dwx - u * dCx - C * dux) * dtdx + (-v * dCy - C * dvy + w * dBy + B * dwy) * dtdy;
  if (P.Omega0 > T(0)) {
    const T xPos = P.xMin + P.dx * h + (i - gw) * P.dx;
    const T shear = T(-1.5) * P.Omega0 * xPos;
    sC0 += (shear * dAx - T(1.5) * P.Omega0 * A) * dtdx;
    sC0 += shear * dBy * dtdy;
  }
  AL += (ELR - ELL) * h * dtdy;
  AR += (ERR - ERL) * h * dtdy;
  BL += -(ERL - ELL) * h * dtdx;
  BR += -(ERR - ELR) * h * dtdx;
  W(V_R, i, j, 0) = r + sr0; W(V_P, i, j, 0) = p + sp0; W(V_U, i, j, 0) = u + su0; W(V_V, i, j, 0) = v + sv0;
  W(V_W, i, j, 0) = w + sw0; W(V_A, i, j, 0) = A + sA0; W(V_B, i, j, 0) = B + sB0; W(V_C, i, j, 0) = C + sC0;
  W(V_AL, i, j, 0) = AL; W(V_AR, i, j, 0) = AR; W(V_BL, i, j, 0) = BL; W(V_BR, i, j, 0) = BR;
  W(V_DRX, i, j, 0) = drx; W(V_DPX, i, j, 0) = dpx; W(V_DUX, i, j, 0) = dux; W(V_DVX, i, j, 0) = dvx;
  W(V_DWX, i, j, 0) = dwx; W(V_DBX, i, j, 0) = dBx; W(V_DCX, i, j, 0) = dCx;
  W(V_DRY, i, j, 0) = dry; W(V_DPY, i, j, 0) = dpy; W(V_DUY, i, j, 0) = duy; W(V_DVY, i, j, 0) = dvy;
  W(V_DWY, i, j, 0) = dwy; W(V_DAY, i, j, 0) = dAy; W(V_DCY, i, j, 0) = dCy;
  W(V_DALY, i, j, 0) = dALy; W(V_DARY, i, j, 0) = dARy; W(V_DBLX, i, j, 0) = dBLx; W(V_DBRX, i, j, 0) = dBRx;
}

template <typename T, int DIR>
__device__ __forceinline__ dev::State<T> face2d(const KParams<T>& P, const View<const T>& W, int i, int j, T sgn) {
  constexpr int S = (DIR == 0) ? V_DRX : V_DRY;
  dev::State<T> s;
  s.r = dev::mx(P.smallr, W(V_R, i, j, 0) + sgn * W(S + 0, i, j, 0));
  s.p = dev::mx(P.smallp * s.r, W(V_P, i, j, 0) + sgn * W(S + 1, i, j, 0));
  const T u = W(V_U, i, j, 0) + sgn * W(S + 2, i, j, 0);
  const T v = W(V_V, i, j, 0) + sgn * W(S + 3, i, j, 0);
  s.w = W(V_W, i, j, 0) + sgn * W(S + 4, i, j, 0);
  if (DIR == 0) {
    s.u = u; s.v = v;
    s.a = (sgn > T(0)) ? W(V_AR, i, j, 0) : W(V_AL, i, j, 0);
    s.b = W(V_B, i, j, 0) + sgn * W(V_DBX, i, j, 0);
    s.c = W(V_C, i, j, 0) + sgn * W(V_DCX, i, j, 0);
  } else {
    s.u = v; s.v = u;
    s.a = (sgn > T(0)) ? W(V_BR, i, j, 0) : W(V_BL, i, j, 0);
    s.b = W(V_A, i, j, 0) + sgn * W(V_DAY, i, j, 0);
    s.c = W(V_C, i, j, 0) + sgn * W(V_DCY, i, j, 0);
  }
  return s;
}

template <typename T>
__device__ __forceinline__ dev::Corner<T> edge2d(const KParams<T>& P, const View<const T>& W, int i, int j, T s1, T s2) {
  dev::Corner<T> c;
  c.r = dev::mx(P.smallr, W(V_R, i, j, 0) + (s1 * W(V_DRX, i, j, 0) + s2 * W(V_DRY, i, j, 0)));
  c.p = dev::mx(P.smallp * c.r, W(V_P, i, j, 0) + (s1 * W(V_DPX, i, j, 0) + s2 * W(V_DPY, i, j, 0)));
  c.u = W(V_U, i, j, 0) + (s1 * W(V_DUX, i, j, 0) + s2 * W(V_DUY, i, j, 0));
  c.v = W(V_V, i, j, 0) + (s1 * W(V_DVX, i, j, 0) + s2 * W(V_DVY, i, j, 0));
  c.w = W(V_W, i, j, 0) + (s1 * W(V_DWX, i, j, 0) + s2 * W(V_DWY, i, j, 0));
  c.a = (s1 > T(0)) ? W(V_AR, i, j, 0) + s2 * W(V_DARY, i, j, 0) : W(V_AL, i, j, 0) + s2 * W(V_DALY, i, j, 0);
  c.b = (s2 > T(0)) ? W(V_BR, i, j, 0) + s1 * W(V_DBRX, i, j, 0) : W(V_BL, i, j, 0) + s1 * W(V_DBLX, i, j, 0);
  c.c = W(V_C, i, j, 0) + (s1 * W(V_DCX, i, j, 0) + s2 * W(V_DCY, i, j, 0));
  return c;
}

// F: 12 components = flux_x(ID,IP,IU,IV,IW,IC) then flux_y(same, physical order); E: emf_z
template <typename T>
__global__ void __launch_bounds__(BX, 4) k2_flux_emf(const __grid_constant__ KParams<T> P, const T* __restrict__ Wp,
                                                     T* __restrict__ Fp, T* __restrict__ Ep) {
  const int gw = P.gw;
  int i, j;
  if (!tileCoords(gw, P.nx + 1, gw, P.ny + 1, i, j)) return;
  const View<const T> W = view<const T>(Wp, P, 1, 0);
  const View<T> F = view(Fp, P, 1, 0);
  const View<T> E = view(Ep, P, 1, 0);
  T f[8];
  if (j < P.jsize - gw) {
    dev::riemann_mhd(P, face2d<T, 0>(P, W, i - 1, j, T(1)), face2d<T, 0>(P, W, i, j, T(-1)), f);
    F(0, i, j, 0) = f[ID]; F(1, i, j, 0) = f[IP]; F(2, i, j, 0) = f[IU]; F(3, i, j, 0) = f[IV];
    F(4, i, j, 0) = f[IW]; F(5, i, j, 0) = f[IC];
  }
  if (i < P.isize - gw) {
    dev::riemann_mhd(P, face2d<T, 1>(P, W, i, j - 1, T(1)), face2d<T, 1>(P, W, i, j, T(-1)), f);
    F(6, i, j, 0) = f[ID]; F(7, i, j, 0) = f[IP]; F(8, i, j, 0) = f[IV]; F(9, i, j, 0) = f[IU];
    F(10, i, j, 0) = f[IW]; F(11, i, j, 0) = f[IC];
  }
  const T xPos = P.xMin + P.dx * T(0.5) + (i - gw) * P.dx;
  E(0, i, j, 0) = dev::compute_emf(P, edge2d(P, W, i - 1, j - 1, T(1), T(1)), edge2d(P, W, i - 1, j, T(1), T(-1)),
                                   edge2d(P, W, i, j - 1, T(-1), T(1)), edge2d(P, W, i, j, T(-1), T(-1)), 2, xPos);
}

template <typename T>
__global__ void __launch_bounds__(BX) k2_update(const __grid_constant__ KParams<T> P, const T* __restrict__ Uold,
                                                T* __restrict__ Unew, const T* __restrict__ Fp, const T* __restrict__ Ep,
                                                T dt, unsigned long long* __restrict__ slots) {
  const int gw = P.gw;
  int i, j;
  const bool valid = tileCoords(0, P.isize, 0, P.jsize, i, j);
  const int iN = P.isize - gw, jN = P.jsize - gw;
  T invDt = T(0);
  if (valid) {
    const UView<T> U = uview(Uold, P);
    const size_t comp = (size_t)P.isize * P.jsize, idx = (size_t)j * P.isize + i;
    T un[8];
#pragma unroll
    for (int v = 0; v < 8; ++v) un[v] = U(v, i, j, 0);
    if (i >= gw && i <= iN && j >= gw && j <= jN) {
      const View<const T> F = view<const T>(Fp, P, 1, 0);
      const View<const T> E = view<const T>(Ep, P, 1, 0);
      const T dtdx = dt / P.dx, dtdy = dt / P.dy;
      const bool inner = i < iN && j < jN;
      if (inner) {  // hydro variables and B_z, reference scatter order: +Fx(i) +Fy(j) -Fx(i+1) -Fy(j+1)
        const int vars[6] = {ID, IP, IU, IV, IW, IC};
#pragma unroll
        for (int n = 0; n < 6; ++n) {
          T s = un[vars[n]];
          s += F(n, i, j, 0) * dtdx;
          s += F(6 + n, i, j, 0) * dtdy;
          s -= F(n, i + 1, j, 0) * dtdx;
          s -= F(6 + n, i, j + 1, 0) * dtdy;
          un[vars[n]] = s;
        }
      }
      auto emf = [&](int ii, int jj) -> T { return (ii > iN || jj > jN) ? T(0) : E(0, ii, jj, 0); };
      auto ct = [&](int ii, int jj, T& bx, T& by) {
        const T e = emf(ii, jj);
        bx += (emf(ii, jj + 1) - e) * dtdy;
        by -= (emf(ii + 1, jj) - e) * dtdx;
      };
      ct(i, j, un[IA], un[IB]);
      if (inner) {
        T bxp = U(IA, i + 1, j, 0), byp = U(IB, i, j + 1, 0), d;
        d = U(IB, i + 1, j, 0); ct(i + 1, j, bxp, d);
        d = U(IA, i, j + 1, 0); ct(i, j + 1, d, byp);
        T q[8];
        dev::cons_to_prim_mhd(P, un, bxp, byp, T(0), T(0), q);
        const T irho = dev::rcp(q[ID]);
        const T a2 = q[IA] * q[IA], b2 = q[IB] * q[IB], bb = a2 + b2 + q[IC] * q[IC];
        invDt = (dev::fast_speed(P.gamma0, q[IP], irho, bb, a2) + dev::ab(q[IU])) * P.rdx +
                (dev::fast_speed(P.gamma0, q[IP], irho, bb, b2) + dev::ab(q[IV])) * P.rdy;
      }
    }
#pragma unroll
    for (int v = 0; v < 8; ++v) Unew[v * comp + idx] = un[v];
  }
  if (slots != nullptr) reduceMaxToSlots(invDt, slots);
}

}  // namespace

template <typename T>
void Mhd2dKernels<T>::step(const KParams<T>& P, const T* Uold, T* Unew, T* Q, T* W, T* F, T* E, T dt,
                           unsigned long long* slots, cudaStream_t s) {
  k2_prim<T><<<gridFor(P.isize - 1, P.jsize - 1, 1), blockShape(), 0, s>>>(P, Uold, Q, dt);
  const int n = P.isize - 2 * P.gw + 2, m = P.jsize - 2 * P.gw + 2;
  k2_trace<T><<<gridFor(n, m, 1), blockShape(), 0, s>>>(P, Uold, Q, W, dt);
  k2_flux_emf<T><<<gridFor(P.nx + 1, P.ny + 1, 1), blockShape(), 0, s>>>(P, W, F, E);
  k2_update<T><<<gridFor(P.isize, P.jsize, 1), blockShape(), 0, s>>>(P, Uold, Unew, F, E, dt, slots);
  launched(4);
}

template struct Mhd2dKernels<double>;
template struct Mhd2dKernels<float>;

}  // namespace rg
