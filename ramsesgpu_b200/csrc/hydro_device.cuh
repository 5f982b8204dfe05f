// Point-wise Euler (hydro) numerics for sm_100a: cons->prim, TVD slopes, and the approx (two-shock
// Newton) / HLL / HLLC Riemann solvers of the reference (cited per function).  Same maths as the
// reference headers, reciprocal-sharing arithmetic; agreement to rounding, not bitwise.
#pragma once
#include <cuda_runtime.h>

#include "mhd_device.cuh"
#include "params.h"

namespace rg {
namespace dev {

// cons -> prim, reference constoprim.h:82-111 (constoprim_3D) + eos :24-28; returns sound speed
template <typename T>
__device__ __forceinline__ T cons_to_prim_hydro(const KParams<T>& P, T d, T e, T mx_, T my_, T mz_, T (&q)[5]) {
  const T r = mx(d, P.smallr);
  const T ir = rcp(r);
  const T u = mx_ * ir, v = my_ * ir, w = mz_ * ir;
  T p, c;
  if (P.cIso > T(0)) {
    p = r * P.cIso * P.cIso;
    c = P.cIso;
  } else {
    const T eken = T(0.5) * (u * u + v * v + w * w);
    const T eint = e * ir - eken;
    p = mx((P.gamma0 - T(1)) * r * eint, r * P.smallp);
    c = sqr_t(mx(P.gamma0 * p * ir, tiny<T>()));
  }
  q[ID] = r; q[IP] = p; q[IU] = u; q[IV] = v; q[IW] = w;
  return c;
}

// hydro slope, reference slope.h:324-427: type 1 = minmod, type 2 = monotonised central
template <typename T>
__device__ __forceinline__ T hydro_slope(T st, T qm, T q0, T qp) {
  if (st == T(1)) {
    const T dlft = q0 - qm, drgt = qp - q0;
    if (dlft * drgt <= T(0)) return T(0);
    return dlft > T(0) ? mn(dlft, drgt) : mx(dlft, drgt);
  }
  if (st == T(2)) return limited_slope(st, qm, q0, qp);
  return T(0);
}

// frame state: (r, p, u = normal velocity, v, w)
template <typename T>
struct HState {
  T r, p, u, v, w;
};

// flux of a Godunov state, reference cmpflx.h:23-49
template <typename T>
__device__ __forceinline__ void cmpflx(const KParams<T>& P, T r, T u, T v, T w, T p, T (&f)[5]) {
  f[ID] = r * u;
  f[IU] = f[ID] * u + p;
  f[IV] = f[ID] * v;
  f[IW] = f[ID] * w;
  const T entho = rcp(P.gamma0 - T(1));
  const T ekin = T(0.5) * r * (u * u + v * v + w * w);
  f[IP] = u * (p * entho + ekin + p);
}

// two-shock approximate Riemann solver with Newton iterations, reference riemann.h:31-160
template <typename T>
__device__ __forceinline__ void riemann_approx(const KParams<T>& P, const HState<T>& L, const HState<T>& Rr, T (&flux)[5]) {
  const T rl = mx(L.r, P.smallr), ul = L.u, pl = mx(L.p, rl * P.smallp);
  const T rr = mx(Rr.r, P.smallr), ur = Rr.u, pr = mx(Rr.p, rr * P.smallp);
  const T cl = P.gamma0 * pl * rl, cr = P.gamma0 * pr * rr;
  T wl = sqr_t(cl), wr = sqr_t(cr);
  const T ipl = rcp(pl), ipr = rcp(pr);
  T pold = mx(((wr * pl + wl * pr) + wl * wr * (ul - ur)) * rcp(wl + wr), T(0));
  T conv = T(1);
  for (int iter = 0; iter < P.niter_riemann && conv > T(1e-6); ++iter) {
    const T wwl = sqr_t(cl * (T(1) + P.gamma6 * (pold - pl) * ipl));
    const T wwr = sqr_t(cr * (T(1) + P.gamma6 * (pold - pr) * ipr));
    const T ql = T(2) * wwl * wwl * wwl * rcp(wwl * wwl + cl);
    const T qr = T(2) * wwr * wwr * wwr * rcp(wwr * wwr + cr);
    const T usl = ul - (pold - pl) * rcp(wwl);
    const T usr = ur + (pold - pr) * rcp(wwr);
    const T delp = mx(qr * ql * rcp(qr + ql) * (usl - usr), -pold);
    pold = pold + delp;
    conv = ab(delp * rcp(pold + P.smallpp));
  }
  const T pstar = pold;
  wl = sqr_t(cl * (T(1) + P.gamma6 * (pstar - pl) * ipl));
  wr = sqr_t(cr * (T(1) + P.gamma6 * (pstar - pr) * ipr));
  const T ustar = T(0.5) * (ul + (pl - pstar) * rcp(wl) + ur - (pr - pstar) * rcp(wr));
  const bool left = !signbit(ustar);  // sgnm = copysign(1, ustar) > 0
  const T sgnm = left ? T(1) : T(-1);
  const T ro = left ? rl : rr, uo = left ? ul : ur, po = left ? pl : pr, wo = left ? wl : wr;
  const T iro = rcp(ro);
  const T co = mx(P.smallc, sqr_t(mx(ab(P.gamma0 * po * iro), tiny<T>())));
  const T rstar = mx(ro * rcp(T(1) + ro * (po - pstar) * rcp(wo * wo)), P.smallr);
  const T cstar = mx(P.smallc, sqr_t(mx(ab(P.gamma0 * pstar * rcp(rstar)), tiny<T>())));
  T spout = co - sgnm * uo;
  T spin = cstar - sgnm * ustar;
  const T ushock = wo * iro - sgnm * uo;
  if (pstar >= po) { spin = ushock; spout = ushock; }
  const T scr = mx(spout - spin, P.smallc + ab(spout + spin));
  T frac = T(0.5) * (T(1) + (spout + spin) * rcp(scr));
  // the reference saturates through a FLOAT helper even in the double build (gpu_macros.cpp:25-30)
  frac = (frac != frac) ? T(0) : static_cast<T>(__saturatef(static_cast<float>(frac)));
  T gr = frac * rstar + (T(1) - frac) * ro;
  T gu = frac * ustar + (T(1) - frac) * uo;
  T gp = frac * pstar + (T(1) - frac) * po;
  if (spout < T(0)) { gr = ro; gu = uo; gp = po; }
  if (spin > T(0)) { gr = rstar; gu = ustar; gp = pstar; }
  const T gv = left ? L.v : Rr.v, gw_ = left ? L.w : Rr.w;
  cmpflx(P, gr, gu, gv, gw_, gp, flux);
}

// HLL, reference riemann.h:177-253
template <typename T>
__device__ __forceinline__ void riemann_hll_hydro(const KParams<T>& P, const HState<T>& L, const HState<T>& Rr, T (&flux)[5]) {
  const T entho = rcp(P.gamma0 - T(1));
  const T rl = mx(L.r, P.smallr), pl = mx(L.p, rl * P.smallp);
  const T rr = mx(Rr.r, P.smallr), pr = mx(Rr.p, rr * P.smallp);
  const T cm = sqr_t(mx(P.gamma0 * pl * rcp(rl), P.gamma0 * pr * rcp(rr)));
  const T SL = mn(mn(L.u, Rr.u) - cm, T(0)), SR = mx(mx(L.u, Rr.u) + cm, T(0));
  T uL[5], uR[5], fL[5], fR[5];
  uL[ID] = L.r; uR[ID] = Rr.r;
  uL[IP] = L.p * entho + T(0.5) * L.r * (L.u * L.u + L.v * L.v + L.w * L.w);
  uR[IP] = Rr.p * entho + T(0.5) * Rr.r * (Rr.u * Rr.u + Rr.v * Rr.v + Rr.w * Rr.w);
  uL[IU] = L.r * L.u; uR[IU] = Rr.r * Rr.u;
  uL[IV] = L.r * L.v; uR[IV] = Rr.r * Rr.v;
  uL[IW] = L.r * L.w; uR[IW] = Rr.r * Rr.w;
  fL[ID] = uL[IU]; fR[ID] = uR[IU];
  fL[IP] = L.u * (uL[IP] + L.p); fR[IP] = Rr.u * (uR[IP] + Rr.p);
  fL[IU] = L.p + uL[IU] * L.u; fR[IU] = Rr.p + uR[IU] * Rr.u;
  fL[IV] = fL[ID] * L.v; fR[IV] = fR[ID] * Rr.v;
  fL[IW] = fL[ID] * L.w; fR[IW] = fR[ID] * Rr.w;
  const T inv = rcp(SR - SL);
#pragma unroll
  for (int n = 0; n < 5; ++n) flux[n] = (SR * fL[n] - SL * fR[n] + SR * SL * (uR[n] - uL[n])) * inv;
}

// HLLC, reference riemann.h:270-371
template <typename T>
__device__ __forceinline__ void riemann_hllc(const KParams<T>& P, const HState<T>& L, const HState<T>& Rr, T (&flux)[5]) {
  const T entho = rcp(P.gamma0 - T(1));
  const T rl = mx(L.r, P.smallr), pl = mx(L.p, rl * P.smallp), ul = L.u;
  const T rr = mx(Rr.r, P.smallr), pr = mx(Rr.p, rr * P.smallp), ur = Rr.u;
  const T etotl = pl * entho + T(0.5) * rl * (ul * ul + L.v * L.v + L.w * L.w);
  const T etotr = pr * entho + T(0.5) * rr * (ur * ur + Rr.v * Rr.v + Rr.w * Rr.w);
  const T sc2 = P.smallc * P.smallc;
  const T cm = sqr_t(mx(mx(P.gamma0 * pl * rcp(rl), sc2), mx(P.gamma0 * pr * rcp(rr), sc2)));
  const T SL = mn(ul, ur) - cm, SR = mx(ul, ur) + cm;
  const T rcl = rl * (ul - SL), rcr = rr * (SR - ur);
  const T irc = rcp(rcr + rcl);
  const T ustar = (rcr * ur + rcl * ul + (pl - pr)) * irc;
  const T ptotstar = (rcr * pl + rcl * pr + rcl * rcr * (ul - ur)) * irc;
  T ro, uo, ptoto, etoto;
  if (SL > T(0)) {
    ro = rl; uo = ul; ptoto = pl; etoto = etotl;
  } else if (ustar > T(0)) {
    const T i = rcp(SL - ustar);
    ro = rl * (SL - ul) * i; uo = ustar; ptoto = ptotstar;
    etoto = ((SL - ul) * etotl - pl * ul + ptotstar * ustar) * i;
  } else if (SR > T(0)) {
    const T i = rcp(SR - ustar);
    ro = rr * (SR - ur) * i; uo = ustar; ptoto = ptotstar;
    etoto = ((SR - ur) * etotr - pr * ur + ptotstar * ustar) * i;
  } else {
    ro = rr; uo = ur; ptoto = pr; etoto = etotr;
  }
  flux[ID] = ro * uo;
  flux[IU] = ro * uo * uo + ptoto;
  flux[IP] = (etoto + ptoto) * uo;
  const bool fromLeft = flux[ID] > T(0);
  flux[IV] = flux[ID] * (fromLeft ? L.v : Rr.v);
  flux[IW] = flux[ID] * (fromLeft ? L.w : Rr.w);
}

// dispatch, reference riemann.h:388-401
// RS >= 0: solver fixed at compile time (one solver body per kernel instead of three per call site)
template <int RS = -1, typename T>
__device__ __forceinline__ void riemann_hydro(const KParams<T>& P, const HState<T>& L, const HState<T>& R, T (&flux)[5]) {
  const int rs = (RS >= 0) ? RS : P.riemannSolver;
  if (rs == RS_HLLC) riemann_hllc(P, L, R, flux);
  else if (rs == RS_APPROX) riemann_approx(P, L, R, flux);
  else if (rs == RS_HLL) riemann_hll_hydro(P, L, R, flux);
  else {
#pragma unroll
    for (int n = 0; n < 5; ++n) flux[n] = T(0);
  }
}

}  // namespace dev
}  // namespace rg
