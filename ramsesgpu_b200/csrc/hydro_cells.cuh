// Per-cell building blocks of the 3D hydro step, generic in the way the arrays are reached (U(v,i,j,k) conservative
// state, W(c,i,j,k) traced state): instantiated by the kernels of kernels_hydro3d.cu on global-memory views and by the
// CPU test suite on host arrays (tests/host_emul).
#pragma once
#include "hydro_device.cuh"
#include "kernels.h"

namespace rg {

namespace {

enum { H_R = 0, H_P, H_U, H_V, H_W, H_DX = 5, H_DY = 10, H_DZ = 15 };  // slopes: (r, p, u, v, w) each
static_assert(H_DZ + 5 == NW_HYDRO, "hydro W layout");

// slopes + half-step predictor of one cell from its primitive state and the six neighbours' -> w[NW_HYDRO]
// (reference HydroRunGodunov.cpp:2663-2740, trace.h:544-661).  Shared by the separate trace kernel (hydro_trace_cell
// below) and by the fused hydro kernel, which keeps the primitives in registers / shared memory.
template <typename T>
__device__ __forceinline__ void hydro_trace_from_prims(const KParams<T>& P, const T (&q)[5], const T (&qxm)[5],
                                                       const T (&qxp)[5], const T (&qym)[5], const T (&qyp)[5],
                                                       const T (&qzm)[5], const T (&qzp)[5], T dt, T (&wv)[NW_HYDRO]) {
  // HALF slopes, reference slope.h:324-427: type 1 = minmod, type 2 = monotonised central, anything else = none.  Both
  // limiters are ONE branch-free formula, sign(a+b) min(hst |a|, hst |b|, |a+b|/4) clipped to zero on a sign change with
  // hst = slope_type / 2 (for type 1 the centred term never binds: |a+b|/4 >= min(|a|,|b|)/2 when the signs agree), so
  // the 15 slopes of a cell cost no slope_type compare and no branch.  Equal to the reference's two formulas up to the
  // rounding of a+b against q+ - q-.
  const T st = P.slope_type, h = T(0.5);
  const T hst = (st == T(1) || st == T(2)) ? h * st : T(0);
  T dx_[5], dy_[5], dz_[5];
#pragma unroll
  for (int v = 0; v < 5; ++v) {
    dx_[v] = dev::half_slope(hst, qxm[v], q[v], qxp[v]);
    dy_[v] = dev::half_slope(hst, qym[v], q[v], qyp[v]);
    dz_[v] = dev::half_slope(hst, qzm[v], q[v], qzp[v]);
  }
  const T dtdx = dt / P.dx, dtdy = dt / P.dy, dtdz = dt / P.dz;
  const T r = q[ID], p = q[IP], u = q[IU], v = q[IV], w = q[IW], g = P.gamma0;
  const T ir = dev::rcp(r);
  // half-step predictor, trace.h:585-600
  const T sr0 = (-u * dx_[ID] - dx_[IU] * r) * dtdx + (-v * dy_[ID] - dy_[IV] * r) * dtdy + (-w * dz_[ID] - dz_[IW] * r) * dtdz;
  const T su0 = (-u * dx_[IU] - dx_[IP] * ir) * dtdx + (-v * dy_[IU]) * dtdy + (-w * dz_[IU]) * dtdz;
  const T sv0 = (-u * dx_[IV]) * dtdx + (-v * dy_[IV] - dy_[IP] * ir) * dtdy + (-w * dz_[IV]) * dtdz;
  const T sw0 = (-u * dx_[IW]) * dtdx + (-v * dy_[IW]) * dtdy + (-w * dz_[IW] - dz_[IP] * ir) * dtdz;
  const T sp0 = (-u * dx_[IP] - dx_[IU] * g * p) * dtdx + (-v * dy_[IP] - dy_[IV] * g * p) * dtdy + (-w * dz_[IP] - dz_[IW] * g * p) * dtdz;
  T gpx = T(0), gpy = T(0), gpz = T(0);
  if (P.gravity) {  // gravity predictor on the traced velocities, reference HydroRunGodunov.cpp:2705-2734
    gpx = h * dt * P.gx; gpy = h * dt * P.gy; gpz = h * dt * P.gz;
  }
  wv[H_R] = r + sr0; wv[H_P] = p + sp0;
  wv[H_U] = u + su0 + gpx; wv[H_V] = v + sv0 + gpy; wv[H_W] = w + sw0 + gpz;
  // slope component order in W: r, p, u, v, w  (ID, IP, IU, IV, IW)
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    wv[H_DX + c] = dx_[c];
    wv[H_DY + c] = dy_[c];
    wv[H_DZ + c] = dz_[c];
  }
}

template <typename T, typename UV, typename WV>
__device__ __forceinline__ void hydro_trace_cell(const KParams<T>& P, const UV& U, const WV& W, int i, int j, int k, T dt) {
  auto prim = [&](int ii, int jj, int kk, T(&q)[5]) {
    dev::cons_to_prim_hydro(P, U(ID, ii, jj, kk), U(IP, ii, jj, kk), U(IU, ii, jj, kk), U(IV, ii, jj, kk),
                            U(IW, ii, jj, kk), q);
  };
  T q[5], qxm[5], qxp[5], qym[5], qyp[5], qzm[5], qzp[5];
  prim(i, j, k, q);
  prim(i - 1, j, k, qxm); prim(i + 1, j, k, qxp);
  prim(i, j - 1, k, qym); prim(i, j + 1, k, qyp);
  prim(i, j, k - 1, qzm); prim(i, j, k + 1, qzp);
  T wv[NW_HYDRO];
  hydro_trace_from_prims(P, q, qxm, qxp, qym, qyp, qzm, qzp, dt, wv);
#pragma unroll
  for (int c = 0; c < NW_HYDRO; ++c) W(c, i, j, k) = wv[c];
}

// state at a face of cell (i,j,k): W centre +/- half slope along DIR, floors (trace.h:603-660),
// rotated so that .u is the velocity normal to the face
template <typename T, int DIR, typename WV>
__device__ __forceinline__ dev::HState<T> hydro_face(const KParams<T>& P, const WV& W, int i, int j, int k, T sgn) {
  constexpr int S = (DIR == 0) ? H_DX : (DIR == 1) ? H_DY : H_DZ;
  dev::HState<T> s;
  s.r = dev::mx(P.smallr, W(H_R, i, j, k) + sgn * W(S + 0, i, j, k));
  s.p = dev::mx(P.smallp * s.r, W(H_P, i, j, k) + sgn * W(S + 1, i, j, k));
  const T u = W(H_U, i, j, k) + sgn * W(S + 2, i, j, k);
  const T v = W(H_V, i, j, k) + sgn * W(S + 3, i, j, k);
  const T w = W(H_W, i, j, k) + sgn * W(S + 4, i, j, k);
  if (DIR == 0) { s.u = u; s.v = v; s.w = w; }
  else if (DIR == 1) { s.u = v; s.v = u; s.w = w; }
  else { s.u = w; s.v = v; s.w = u; }
  return s;
}

// flux through the LOW face of cell (i,j,k) along DIR, in physical component order
template <typename T, int DIR, int RS, typename WV>
__device__ __forceinline__ void hydro_low_flux(const KParams<T>& P, const WV& W, int i, int j, int k, T (&f)[5]) {
  const dev::HState<T> L = hydro_face<T, DIR>(P, W, i - (DIR == 0), j - (DIR == 1), k - (DIR == 2), T(1));
  const dev::HState<T> R = hydro_face<T, DIR>(P, W, i, j, k, T(-1));
  T fr[5];
  dev::riemann_hydro<RS>(P, L, R, fr);
  f[ID] = fr[ID]; f[IP] = fr[IP];
  f[IU] = (DIR == 0) ? fr[IU] : (DIR == 1) ? fr[IV] : fr[IW];
  f[IV] = (DIR == 1) ? fr[IU] : fr[IV];
  f[IW] = (DIR == 2) ? fr[IU] : fr[IW];
}

template <typename T, int DIR>
__device__ __forceinline__ dev::HState<T> face_from_regs(const KParams<T>& P, const T (&w)[NW_HYDRO], T sgn) {
  constexpr int S = (DIR == 0) ? H_DX : (DIR == 1) ? H_DY : H_DZ;
  dev::HState<T> s;
  s.r = dev::mx(P.smallr, w[H_R] + sgn * w[S + 0]);
  s.p = dev::mx(P.smallp * s.r, w[H_P] + sgn * w[S + 1]);
  const T u = w[H_U] + sgn * w[S + 2], v = w[H_V] + sgn * w[S + 3], ww = w[H_W] + sgn * w[S + 4];
  if (DIR == 0) { s.u = u; s.v = v; s.w = ww; }
  else if (DIR == 1) { s.u = v; s.v = u; s.w = ww; }
  else { s.u = ww; s.v = v; s.w = u; }
  return s;
}

// Riemann flux of a face normal to DIR from its rotated left/right states, in physical component order
template <typename T, int DIR, int RS>
__device__ __forceinline__ void face_flux(const KParams<T>& P, const dev::HState<T>& L, const dev::HState<T>& R, T (&f)[5]) {
  T fr[5];
  dev::riemann_hydro<RS>(P, L, R, fr);
  f[ID] = fr[ID]; f[IP] = fr[IP];
  f[IU] = (DIR == 0) ? fr[IU] : (DIR == 1) ? fr[IV] : fr[IW];
  f[IV] = (DIR == 1) ? fr[IU] : fr[IV];
  f[IW] = (DIR == 2) ? fr[IU] : fr[IW];
}

}  // namespace

}  // namespace rg
