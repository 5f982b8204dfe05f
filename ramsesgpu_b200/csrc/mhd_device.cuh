// Point-wise MHD numerics for sm_100a, written for the FP64 pipe: every quotient that shares a
// denominator is turned into one reciprocal + multiplies, |x|/sqrt(y) into x*rsqrt(y), and the
// HLLD / 2-D HLLD solvers are straight-line (select, no divergent branches).
//
// The MATHS is that of the reference's headers (cited per function); the operation order is not,
// so results agree with the reference to rounding (see tests/ for the stated tolerances), not
// bitwise.
#pragma once
#include <cuda_runtime.h>

#include "params.h"

namespace rg {
namespace dev {

// ---- scalar helpers -------------------------------------------------------------------------
// FP64 reciprocal / rsqrt / sqrt as MUFU seed + Newton steps WITHOUT the library's slow-path
// branch (denormal / inf / zero handling): every denominator on this path is a normal number
// (densities, wave-speed differences, positive radicands), and dropping the branch removes ~8
// non-FP64 instructions + a BSSY/BSYNC pair per call.  ONE cubic Newton step after the 2^-20 seed
// (error 2^-60 before rounding): results are within ~2 ulp, far inside the 1e-12 parity tolerance.
__device__ __forceinline__ double rcp(double x) {
  double r;
#ifdef RG_HOST_EMULATION  // tests/host_emul: the same Newton steps on a 20-bit seed, compiled for the host (CPU test suite)
  r = rg_host_seed20(1.0 / x);
#else
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));   // MUFU.RCP64H, ~20 bits
#endif
  double e = fma(-x, r, 1.0);
  e = fma(e, e, e);
  return fma(r, e, r);                                     // cubic step: 2^-20 -> 2^-60 + rounding
}
// FP32: the MUFU results (<= 2 ulp) without the IEEE slow-path branches of __frcp_rn / sqrtf
__device__ __forceinline__ float rcp(float x) {
#ifdef RG_HOST_EMULATION
  return 1.0f / x;
#else
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));   // one MUFU.RCP (__fdividef(1, x) adds a multiply)
  return r;
#endif
}
__device__ __forceinline__ double rsq(double x) {
  double y;
#ifdef RG_HOST_EMULATION
  y = rg_host_seed20(1.0 / sqrt(x));
#else
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));  // MUFU.RSQ64H
#endif
  const double e = fma(-x * y, y, 1.0);
  return fma(y * e, fma(e, 0.375, 0.5), y);                // cubic step
}
__device__ __forceinline__ float rsq(float x) { return rsqrtf(x); }
// sqrt for x > 0 (callers clamp radicands that can reach 0 to a tiny positive number)
__device__ __forceinline__ double sqr_t(double x) { return x * rsq(x); }
__device__ __forceinline__ float sqr_t(float x) { return x * rsqrtf(x); }
template <typename T> __device__ __forceinline__ T tiny();
template <> __device__ __forceinline__ double tiny<double>() { return 1e-300; }
template <> __device__ __forceinline__ float tiny<float>() { return 1e-37f; }
// double max/min as compare + select (3 instructions); fmax/fmin cost 7 with their NaN handling,
// and no NaN is ever an operand on this path
__device__ __forceinline__ double mx(double a, double b) { return (a > b) ? a : b; }
__device__ __forceinline__ float mx(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double mn(double a, double b) { return (a < b) ? a : b; }
__device__ __forceinline__ float mn(float a, float b) { return fminf(a, b); }
// forced select (the compiler otherwise turns long select ladders into divergent branches)
__device__ __forceinline__ double pick(bool c, double a, double b) {
#ifdef RG_HOST_EMULATION
  return c ? a : b;
#else
  double r;
  asm("{ .reg .pred p; setp.ne.s32 p, %3, 0; selp.f64 %0, %1, %2, p; }" : "=d"(r) : "d"(a), "d"(b), "r"((int)c));
  return r;
#endif
}
__device__ __forceinline__ float pick(bool c, float a, float b) { return c ? a : b; }
__device__ __forceinline__ double ab(double a) { return fabs(a); }
__device__ __forceinline__ float ab(float a) { return fabsf(a); }
template <typename T> __device__ __forceinline__ T max4(T a, T b, T c, T d) { return mx(mx(a, b), mx(c, d)); }
template <typename T> __device__ __forceinline__ T min4(T a, T b, T c, T d) { return mn(mn(a, b), mn(c, d)); }
// clamps at zero and the guard in front of a square root, on the INTEGER pipe: a sign-bit mask for the clamps (min0 keeps
// a -0.0), one integer compare of the high words for the guard (threshold 2^-996 instead of 1e-300: both only keep rsq()
// away from 0 and negatives).  No DSETP on the FP64 pipe: measured -3.2 % on the fused update kernel against the
// compare-select forms (profiles/r02_a_ab_variants.txt).  RG_FP64_CLAMP restores the compare-select forms.
#if !defined(RG_FP64_CLAMP)
__device__ __forceinline__ double min0(double x) {
  const int hi = __double2hiint(x), m = hi >> 31;
  return __hiloint2double(hi & m, __double2loint(x) & m);
}
__device__ __forceinline__ double max0(double x) {
  const int hi = __double2hiint(x), m = ~(hi >> 31);
  return __hiloint2double(hi & m, __double2loint(x) & m);
}
__device__ __forceinline__ double guard_tiny(double x) {
  const int hi = __double2hiint(x);
  const bool small = hi < 0x01b00000;  // also negative values and +0
  return __hiloint2double(small ? 0x01b00000 : hi, small ? 0 : __double2loint(x));
}
#else
__device__ __forceinline__ double min0(double x) { return mn(x, 0.0); }
__device__ __forceinline__ double max0(double x) { return mx(x, 0.0); }
__device__ __forceinline__ double guard_tiny(double x) { return mx(x, tiny<double>()); }
#endif
__device__ __forceinline__ float min0(float x) { return mn(x, 0.0f); }
__device__ __forceinline__ float max0(float x) { return mx(x, 0.0f); }
__device__ __forceinline__ float guard_tiny(float x) { return mx(x, tiny<float>()); }

// A primitive state in the frame of the interface normal: (r, p, u=normal v, v, w, a=normal B, b, c)
template <typename T>
struct State {
  T r, p, u, v, w, a, b, c;
};

// TVD slope (minmod / MC by slope_type), reference slope_mhd.h:466-474 and :640-648.
template <typename T>
__device__ __forceinline__ T limited_slope(T st, T qm, T q0, T qp) {
  T dlft = st * (q0 - qm);
  T drgt = st * (qp - q0);
  T dcen = T(0.5) * (qp - qm);
  T dsgn = (dcen >= T(0)) ? T(1) : T(-1);
  T dlim = mn(ab(dlft), ab(drgt));
  if (dlft * drgt <= T(0)) dlim = T(0);
  return dsgn * mn(dlim, ab(dcen));
}
// the same limiter returning HALF the slope, arranged for the FP64 pipe: 3 adds, 2 multiplies and
// 2 compares; the sign of dcen is copied with an integer op and "dlft*drgt <= 0" is the sign-bit
// test of the two one-sided differences (a zero difference already gives min(|.|,|.|) = 0).
// hst = st/2.  dcen is formed as (a+b)/2 instead of (qp-qm)/2: equal up to rounding.
__device__ __forceinline__ double half_slope(double hst, double qm, double q0, double qp) {
  const double a = q0 - qm, b = qp - q0;
  const double s = a + b;
#if !defined(RG_LIMITER_V0)
  // one three-way minimum of the sign-flipped terms, clamped at zero with a sign mask -- opposite signs give a negative
  // minimum (tools/microbench/limiter_variants.cu: 19 -> 16 instructions, 9 -> 6 on the FP64 pipe per slope)
  const int sg = __double2hiint(s) & 0x80000000;
  const double fa = __hiloint2double(__double2hiint(a) ^ sg, __double2loint(a)) * hst;
  const double fb = __hiloint2double(__double2hiint(b) ^ sg, __double2loint(b)) * hst;
  const double rr = mn(mn(fa, fb), ab(s) * 0.25);
  const int hi = __double2hiint(rr), m = ~(hi >> 31);
  return __hiloint2double((hi & m) | sg, __double2loint(rr) & m);
#else
  const double m = mn(ab(a), ab(b)) * hst;
  const double c = ab(s) * 0.25;
  double r = mn(m, c);
  if ((__double2hiint(a) ^ __double2hiint(b)) < 0) r = 0.0;
  return __hiloint2double((__double2hiint(r) & 0x7fffffff) | (__double2hiint(s) & 0x80000000), __double2loint(r));
#endif
}
__device__ __forceinline__ float half_slope(float hst, float qm, float q0, float qp) {
  const float a = q0 - qm, b = qp - q0;
  const float s = a + b;
  float r = fminf(fminf(fabsf(a), fabsf(b)) * hst, fabsf(s) * 0.25f);
  if ((__float_as_int(a) ^ __float_as_int(b)) < 0) r = 0.0f;
  return copysignf(r, s);
}

// cons -> prim for one cell. reference constoprim.h:137-199 (constoprim_mhd).
//   u[8] conservative (B = left faces), bn[3] = B faces of the +1 neighbours.
// FAST = compile-time promise "adiabatic (cIso == 0), non-rotating (Omega0 == 0), HLLD + 2-D HLLD":
// the launch wrappers pick the FAST instantiation of the kernels when the run parameters say so.
template <bool FAST = false, typename T>
__device__ __forceinline__ void cons_to_prim_mhd(const KParams<T>& P, const T (&u)[8], T bxp, T byp, T bzp,
                                                 T dt, T (&q)[8]) {
  T r = mx(u[ID], P.smallr);
  T ir = rcp(r);
  T vx = u[IU] * ir, vy = u[IV] * ir, vz = u[IW] * ir;
  T A = T(0.5) * (u[IA] + bxp), B = T(0.5) * (u[IB] + byp), C = T(0.5) * (u[IC] + bzp);
  T p;
  if (!FAST && P.cIso > T(0)) {
    p = r * P.cIso * P.cIso;
  } else {
    T eken = T(0.5) * (vx * vx + vy * vy + vz * vz);
    T emag = T(0.5) * (A * A + B * B + C * C);
    T eint = (u[IP] - emag) * ir - eken;
    p = mx((P.gamma0 - T(1)) * r * eint, r * P.smallp);
  }
  if (!FAST && P.Omega0 > T(0)) {  // Coriolis predictor, constoprim.h:189-195
    T dvx = T(2.0) * P.Omega0 * vy;
    T dvy = T(-0.5) * P.Omega0 * vx;
    vx += dvx * dt * T(0.5);
    vy += dvy * dt * T(0.5);
  }
  q[ID] = r; q[IP] = p; q[IU] = vx; q[IV] = vy; q[IW] = vz; q[IA] = A; q[IB] = B; q[IC] = C;
}

// fast magnetosonic speed along the normal (component `n` of B), reference mhd_utils.h:28-52.
// ir = 1/r.  The inner radicand is clamped at 0 (it is >= 0 analytically).
template <typename T>
__device__ __forceinline__ T fast_speed(T gamma, T p, T ir, T b2, T n2) {
  T c2 = gamma * p * ir;
  T d2 = T(0.5) * (b2 * ir + c2);
  return sqr_t(d2 + sqr_t(guard_tiny(d2 * d2 - c2 * n2 * ir)));
}

// square of the fast speed (one sqrt less when only a max over states is needed)
template <typename T>
__device__ __forceinline__ T fast_speed2(T gamma, T p, T ir, T b2, T n2) {
  T c2 = gamma * p * ir;
  T d2 = T(0.5) * (b2 * ir + c2);
  return d2 + sqr_t(guard_tiny(d2 * d2 - c2 * n2 * ir));
}

// 1-D physical flux + conservative vector, reference mhd_utils.h:106-156
template <typename T>
__device__ __forceinline__ void mhd_flux(const KParams<T>& P, const State<T>& s, T (&cv)[8], T (&ff)[8]) {
  T p = (P.cIso > T(0)) ? s.r * P.cIso * P.cIso : s.p;
  T entho = rcp(P.gamma0 - T(1));
  T ecin = T(0.5) * (s.u * s.u + s.v * s.v + s.w * s.w) * s.r;
  T emag = T(0.5) * (s.a * s.a + s.b * s.b + s.c * s.c);
  T etot = p * entho + ecin + emag;
  T ptot = p + emag;
  cv[ID] = s.r; cv[IP] = etot; cv[IU] = s.r * s.u; cv[IV] = s.r * s.v; cv[IW] = s.r * s.w;
  cv[IA] = s.a; cv[IB] = s.b; cv[IC] = s.c;
  ff[ID] = s.r * s.u;
  ff[IP] = (etot + ptot) * s.u - s.a * (s.a * s.u + s.b * s.v + s.c * s.w);
  ff[IU] = s.r * s.u * s.u - s.a * s.a + ptot;
  ff[IV] = s.r * s.u * s.v - s.a * s.b;
  ff[IW] = s.r * s.u * s.w - s.a * s.c;
  ff[IA] = T(0);
  ff[IB] = s.b * s.u - s.a * s.v;
  ff[IC] = s.c * s.u - s.a * s.w;
}

// HLL and LLF fluxes (reference riemann_mhd.h:41-71, :86-118); not the tuned path.
template <typename T>
__device__ void riemann_hll(const KParams<T>& P, State<T> l, State<T> r, T (&flux)[8]) {
  T bm = T(0.5) * (l.a + r.a);
  l.a = bm; r.a = bm;
  T ul[8], fl[8], ur[8], fr[8];
  mhd_flux(P, l, ul, fl);
  mhd_flux(P, r, ur, fr);
  T cl = fast_speed(P.gamma0, l.p, rcp(l.r), l.a * l.a + l.b * l.b + l.c * l.c, l.a * l.a);
  T cr = fast_speed(P.gamma0, r.p, rcp(r.r), r.a * r.a + r.b * r.b + r.c * r.c, r.a * r.a);
  T cm = mx(cl, cr);
  T sl = min0(mn(l.u, r.u) - cm);
  T sr = max0(mx(l.u, r.u) + cm);
  T inv = rcp(sr - sl);
#pragma unroll
  for (int n = 0; n < 8; ++n) flux[n] = (sr * fl[n] - sl * fr[n] + sr * sl * (ur[n] - ul[n])) * inv;
}

template <typename T>
__device__ void riemann_llf(const KParams<T>& P, State<T> l, State<T> r, T (&flux)[8], T zero_flux) {
  T bm = T(0.5) * (l.a + r.a);
  l.a = bm; r.a = bm;
  T ul[8], fl[8], ur[8], fr[8];
  mhd_flux(P, l, ul, fl);
  mhd_flux(P, r, ur, fr);
  // the reference averages the PRIMITIVE states here (riemann_mhd.h:105-106), kept as is
  const T ql[8] = {l.r, l.p, l.u, l.v, l.w, l.a, l.b, l.c};
  const T qr[8] = {r.r, r.p, r.u, r.v, r.w, r.a, r.b, r.c};
  T cl = fast_speed(P.gamma0, l.p, rcp(l.r), l.a * l.a + l.b * l.b + l.c * l.c, l.a * l.a) + ab(l.u);
  T cr = fast_speed(P.gamma0, r.p, rcp(r.r), r.a * r.a + r.b * r.b + r.c * r.c, r.a * r.a) + ab(r.u);
  T vel = mx(cl, cr);
#pragma unroll
  for (int n = 0; n < 8; ++n)
    flux[n] = (ql[n] + qr[n]) * T(0.5) * zero_flux - vel * (ur[n] - ul[n]) * T(0.5);
}

// HLLD (Miyoshi & Kusano 2005) as in reference riemann_mhd.h:139-342.  8 reciprocals, 3 square
// roots and 2 reciprocal square roots per interface instead of 29 divisions + 6 square roots.
template <bool FAST = false, typename T>
__device__ __forceinline__ void riemann_hlld(const KParams<T>& P, State<T> L, State<T> Rr, T (&flux)[8]) {
  const T entho = rcp(P.gamma0 - T(1));
  const T a = T(0.5) * (L.a + Rr.a);
  const T sgnm = (a >= T(0)) ? T(1) : T(-1);
  const T a2 = a * a;
  if (!FAST && P.cIso > T(0)) {
    L.p = L.r * P.cIso * P.cIso;
    Rr.p = Rr.r * P.cIso * P.cIso;
  }
  const T rl = L.r, pl = L.p, ul = L.u, vl = L.v, wl = L.w, bl = L.b, cl = L.c;
  const T rr = Rr.r, pr = Rr.p, ur = Rr.u, vr = Rr.v, wr = Rr.w, br = Rr.b, cr = Rr.c;

  const T emagl = T(0.5) * (a2 + bl * bl + cl * cl);
  const T etotl = pl * entho + T(0.5) * (ul * ul + vl * vl + wl * wl) * rl + emagl;
  const T ptotl = pl + emagl;
  const T vdotbl = ul * a + vl * bl + wl * cl;
  const T emagr = T(0.5) * (a2 + br * br + cr * cr);
  const T etotr = pr * entho + T(0.5) * (ur * ur + vr * vr + wr * wr) * rr + emagr;
  const T ptotr = pr + emagr;
  const T vdotbr = ur * a + vr * br + wr * cr;

  const T cmax = sqr_t(mx(fast_speed2(P.gamma0, pl, rcp(rl), T(2) * emagl, a2),
                           fast_speed2(P.gamma0, pr, rcp(rr), T(2) * emagr, a2)));
  const T sl = mn(ul, ur) - cmax;
  const T sr = mx(ul, ur) + cmax;

  const T rcl = rl * (ul - sl), rcr = rr * (sr - ur);
  const T irc = rcp(rcr + rcl);
  const T ustar = (rcr * ur + rcl * ul + (ptotl - ptotr)) * irc;
  const T ptotstar = (rcr * ptotl + rcl * ptotr + rcl * rcr * (ul - ur)) * irc;

  // left star state
  const T dsl = sl - ul, dslu = sl - ustar;
  const T idslu = rcp(dslu);
  const T rstarl = rl * dsl * idslu;
  const T estarl = rl * dsl * dslu - a2;
  const T el = rl * dsl * dsl - a2;
  // degenerate case of the reference (riemann_mhd.h:205-219): the star values stay the outer ones;
  // computed unconditionally and selected (a non-finite unselected value is harmless)
  const bool degl = a2 > T(0) && ab(estarl - a2) <= T(1e-8) * a2;
  const T iel = rcp(estarl);
  const T kl = a * (ustar - ul) * iel;
  const T vstarl = pick(degl, vl, vl - kl * bl);
  const T wstarl = pick(degl, wl, wl - kl * cl);
  const T bstarl = pick(degl, bl, bl * el * iel);
  const T cstarl = pick(degl, cl, cl * el * iel);
  const T vdotbstarl = ustar * a + vstarl * bstarl + wstarl * cstarl;
  const T etotstarl = (dsl * etotl - ptotl * ul + ptotstar * ustar + a * (vdotbl - vdotbstarl)) * idslu;
  const T irsl = rsq(rstarl);
  const T sqrl = rstarl * irsl;
  const T sal = ustar - ab(a) * irsl;

  // right star state
  const T dsr = sr - ur, dsru = sr - ustar;
  const T idsru = rcp(dsru);
  const T rstarr = rr * dsr * idsru;
  const T estarr = rr * dsr * dsru - a2;
  const T er = rr * dsr * dsr - a2;
  const bool degr = a2 > T(0) && ab(estarr - a2) <= T(1e-8) * a2;
  const T ier = rcp(estarr);
  const T kr = a * (ustar - ur) * ier;
  const T vstarr = pick(degr, vr, vr - kr * br);
  const T wstarr = pick(degr, wr, wr - kr * cr);
  const T bstarr = pick(degr, br, br * er * ier);
  const T cstarr = pick(degr, cr, cr * er * ier);
  const T vdotbstarr = ustar * a + vstarr * bstarr + wstarr * cstarr;
  const T etotstarr = (dsr * etotr - ptotr * ur + ptotstar * ustar + a * (vdotbr - vdotbstarr)) * idsru;
  const T irsr = rsq(rstarr);
  const T sqrr = rstarr * irsr;
  const T sar = ustar + ab(a) * irsr;

  // double star state
  const T isq = rcp(sqrl + sqrr);
  const T vss = (sqrl * vstarl + sqrr * vstarr + sgnm * (bstarr - bstarl)) * isq;
  const T wss = (sqrl * wstarl + sqrr * wstarr + sgnm * (cstarr - cstarl)) * isq;
  const T bss = (sqrl * bstarr + sqrr * bstarl + sgnm * sqrl * sqrr * (vstarr - vstarl)) * isq;
  const T css = (sqrl * cstarr + sqrr * cstarl + sgnm * sqrl * sqrr * (wstarr - wstarl)) * isq;
  const T vdotbss = ustar * a + vss * bss + wss * css;
  const T etotssl = etotstarl - sgnm * sqrl * (vdotbstarl - vdotbss);
  const T etotssr = etotstarr + sgnm * sqrr * (vdotbstarr - vdotbss);

  // sample at x/t = 0 (riemann_mhd.h:268-330).  The reference's ladder
  //   sl > 0: L | sal > 0: L* | ustar > 0: L** | sar > 0: R** | sr > 0: R* | else: R
  // is evaluated with explicit selects per output (32 selects; as an if/else ladder over nine
  // variables the compiler emits divergent branches with ~18 register moves per level)
  const bool c1 = sl > T(0), c2 = sal > T(0), c3 = ustar > T(0), c4 = sar > T(0), c5 = sr > T(0);
  const bool cR = !(c1 || c2 || c3 || c4 || c5);  // region R
  const bool c34 = c3 || c4;                        // (given !c1, !c2) one of the two double-star regions
  const bool c23 = c2 || c3;                        // (given !c1) left of the contact
  const T uo = pick(c1, ul, pick(cR, ur, ustar));
  const T ptoto = pick(c1, ptotl, pick(cR, ptotr, ptotstar));
  const T ro = pick(c1, rl, pick(cR, rr, pick(c23, rstarl, rstarr)));
  const T vo = pick(c1, vl, pick(cR, vr, pick(c2, vstarl, pick(c34, vss, vstarr))));
  const T wo = pick(c1, wl, pick(cR, wr, pick(c2, wstarl, pick(c34, wss, wstarr))));
  const T bo = pick(c1, bl, pick(cR, br, pick(c2, bstarl, pick(c34, bss, bstarr))));
  const T co = pick(c1, cl, pick(cR, cr, pick(c2, cstarl, pick(c34, css, cstarr))));
  const T vdotbo = pick(c1, vdotbl, pick(cR, vdotbr, pick(c2, vdotbstarl, pick(c34, vdotbss, vdotbstarr))));
  const T etoto = pick(c1, etotl, pick(cR, etotr, pick(c2, etotstarl, pick(c3, etotssl, pick(c4, etotssr, etotstarr)))));
  flux[ID] = ro * uo;
  flux[IP] = (etoto + ptoto) * uo - a * vdotbo;
  flux[IU] = ro * uo * uo - a2 + ptoto;
  flux[IV] = ro * uo * vo - a * bo;
  flux[IW] = ro * uo * wo - a * co;
  flux[IA] = T(0);
  flux[IB] = bo * uo - a * vo;
  flux[IC] = co * uo - a * wo;
}

// dispatch, reference riemann_mhd.h:354-368
template <bool FAST = false, typename T>
__device__ __forceinline__ void riemann_mhd(const KParams<T>& P, const State<T>& l, const State<T>& r, T (&flux)[8]) {
  if (FAST || P.riemannSolver == RS_HLLD) {
    riemann_hlld<FAST>(P, l, r, flux);
  } else if (P.riemannSolver == RS_HLL) {
    riemann_hll(P, l, r, flux);
  } else if (P.riemannSolver == RS_LLF) {
    riemann_llf(P, l, r, flux, T(1));
  } else {
#pragma unroll
    for (int n = 0; n < 8; ++n) flux[n] = T(0);  // reference leaves flux untouched (zero-initialised)
  }
}

// ---- 2-D magnetic Riemann solvers -----------------------------------------------------------
// The four states around an edge, in the edge frame (u,v parallel velocities; a,b parallel B;
// w,c orthogonal), index 0..3 = LL, RL, LR, RR (reference constants.h:179-184).
template <typename T>
struct Corner {
  T r, p, u, v, w, a, b, c;
};

// 2-D HLLD, reference riemann_mhd.h:615-821.  12 reciprocals + 10 sqrt + 16 rsqrt
// (reference: 82 divisions + 32 sqrt).
template <typename T>
__device__ __forceinline__ T mag_riemann2d_hlld(const KParams<T>& P, const Corner<T>& LL, const Corner<T>& RL,
                                                const Corner<T>& LR, const Corner<T>& RR) {
  const T g = P.gamma0;
  T cxLL, cyLL, cxLR, cyLR, cxRL, cyRL, cxRR, cyRR;
  T PtotLL, PtotLR, PtotRL, PtotRR;
#define RG_SPEEDS(S, cx, cy, Ptot)                                  \
  {                                                                 \
    const T ir = rcp(S.r);                                          \
    const T a2 = S.a * S.a, b2_ = S.b * S.b;                        \
    const T bb = a2 + b2_ + S.c * S.c;                              \
    const T c2 = g * S.p * ir;                                      \
    const T d2 = T(0.5) * (bb * ir + c2);                           \
    const T dd = d2 * d2, ci = c2 * ir;                             \
    cx = d2 + sqr_t(guard_tiny(dd - ci * a2));                        \
    cy = d2 + sqr_t(guard_tiny(dd - ci * b2_));                       \
    Ptot = S.p + T(0.5) * bb;                                       \
  }
  RG_SPEEDS(LL, cxLL, cyLL, PtotLL)
  RG_SPEEDS(LR, cxLR, cyLR, PtotLR)
  RG_SPEEDS(RL, cxRL, cyRL, PtotRL)
  RG_SPEEDS(RR, cxRR, cyRR, PtotRR)
#undef RG_SPEEDS
  // sqrt is monotonic: max of the four fast speeds = sqrt of the max of their squares
  const T cxm = sqr_t(max4(cxLL, cxLR, cxRL, cxRR)), cym = sqr_t(max4(cyLL, cyLR, cyRL, cyRR));
  const T SL = min4(LL.u, LR.u, RL.u, RR.u) - cxm;
  const T SR = max4(LL.u, LR.u, RL.u, RR.u) + cxm;
  const T SB = min4(LL.v, LR.v, RL.v, RR.v) - cym;
  const T ST = max4(LL.v, LR.v, RL.v, RR.v) + cym;

  const T rcLLx = LL.r * (LL.u - SL), rcRLx = RL.r * (SR - RL.u);
  const T rcLRx = LR.r * (LR.u - SL), rcRRx = RR.r * (SR - RR.u);
  const T rcLLy = LL.r * (LL.v - SB), rcLRy = LR.r * (ST - LR.v);
  const T rcRLy = RL.r * (RL.v - SB), rcRRy = RR.r * (ST - RR.v);

  const T ustar = (rcLLx * LL.u + rcLRx * LR.u + rcRLx * RL.u + rcRRx * RR.u + (PtotLL - PtotRL + PtotLR - PtotRR)) *
                  rcp(rcLLx + rcLRx + rcRLx + rcRRx);
  const T vstar = (rcLLy * LL.v + rcLRy * LR.v + rcRLy * RL.v + rcRRy * RR.v + (PtotLL - PtotLR + PtotRL - PtotRR)) *
                  rcp(rcLLy + rcLRy + rcRLy + rcRRy);

  const T iSL = rcp(SL - ustar), iSR = rcp(SR - ustar), iSB = rcp(SB - vstar), iST = rcp(ST - vstar);

  // per corner: fx/fy compression factors, star fields, star emfs, Alfven candidates
#define RG_CORNER(S, Sx, iSx, Sy, iSy, Bs, As, Ex, Ey, Es, cax, caxy, cby, cbxy) \
  T Bs, As, Ex, Ey, Es, cax, caxy, cby, cbxy;                                     \
  {                                                                               \
    const T fx = (Sx - S.u) * iSx, fy = (Sy - S.v) * iSy;                         \
    const T rsx = S.r * fx, rsy = S.r * fy, rs = rsx * fy;                        \
    Bs = S.b * fx;                                                                \
    As = S.a * fy;                                                                \
    Ex = ustar * Bs - S.v * S.a;                                                  \
    Ey = S.u * S.b - vstar * As;                                                  \
    Es = ustar * Bs - vstar * As;                                                 \
    const T irs = rsq(rs);                                                        \
    cax = ab(S.a) * rsq(rsx);                                                     \
    caxy = ab(As) * irs;                                                          \
    cby = ab(S.b) * rsq(rsy);                                                     \
    cbxy = ab(Bs) * irs;                                                          \
  }
  RG_CORNER(LL, SL, iSL, SB, iSB, BsLL, AsLL, ExLL, EyLL, EsLL, caLLx, caLL, cbLLy, cbLL)
  RG_CORNER(LR, SL, iSL, ST, iST, BsLR, AsLR, ExLR, EyLR, EsLR, caLRx, caLR, cbLRy, cbLR)
  RG_CORNER(RL, SR, iSR, SB, iSB, BsRL, AsRL, ExRL, EyRL, EsRL, caRLx, caRL, cbRLy, cbRL)
  RG_CORNER(RR, SR, iSR, ST, iST, BsRR, AsRR, ExRR, EyRR, EsRR, caRRx, caRR, cbRRy, cbRR)
#undef RG_CORNER
  const T calfL = mx(max4(caLRx, caLR, caLLx, caLL), P.smallc);
  const T calfR = mx(max4(caRRx, caRR, caRLx, caRL), P.smallc);
  const T calfB = mx(max4(cbLLy, cbLL, cbRLy, cbRL), P.smallc);
  const T calfT = mx(max4(cbLRy, cbLR, cbRRy, cbRR), P.smallc);

  const T SAL = min0(ustar - calfL), SAR = max0(ustar + calfR);
  const T SAB = min0(vstar - calfB), SAT = max0(vstar + calfT);
  const T iA = rcp(SAR - SAL), iB = rcp(SAT - SAB);

  // region selection: the reference's integer masks (riemann_mhd.h:759-787) are equivalent to this
  // if/else ladder with "positive" meaning sign bit clear (copysign semantics: -0.0 is negative).
  const bool SBp = !signbit(SB), STp = !signbit(ST), SLp = !signbit(SL), SRp = !signbit(SR);
  const T ELL = LL.u * LL.b - LL.v * LL.a, ERL = RL.u * RL.b - RL.v * RL.a;
  const T ELR = LR.u * LR.b - LR.v * LR.a, ERR = RR.u * RR.b - RR.v * RR.a;
  T E;
  if (SBp) {
    E = SLp ? ELL : (!SRp ? ERL : (SAR * ExLL - SAL * ExRL + SAR * SAL * (RL.b - LL.b)) * iA);
  } else if (!STp) {
    E = SLp ? ELR : (!SRp ? ERR : (SAR * ExLR - SAL * ExRR + SAR * SAL * (RR.b - LR.b)) * iA);
  } else if (SLp) {
    E = (SAT * EyLL - SAB * EyLR - SAT * SAB * (LR.a - LL.a)) * iB;
  } else if (!SRp) {
    E = (SAT * EyRL - SAB * EyRR - SAT * SAB * (RR.a - RL.a)) * iB;
  } else {
    const T AstarT = (SAR * AsRR - SAL * AsLR) * iA, AstarB = (SAR * AsRL - SAL * AsLL) * iA;
    const T BstarR = (SAT * BsRR - SAB * BsRL) * iB, BstarL = (SAT * BsLR - SAB * BsLL) * iB;
    E = (SAL * SAB * EsRR - SAL * SAT * EsRL - SAR * SAB * EsLR + SAR * SAT * EsLL) * iA * iB -
        SAT * SAB * iB * (AstarT - AstarB) + SAR * SAL * iA * (BstarR - BstarL);
  }
  return E;
}

// HLLA (alfven) / HLLF (fast) 2-D solvers, reference riemann_mhd.h:417-507
template <typename T>
__device__ T mag_riemann2d_hll(const KParams<T>& P, const Corner<T> (&q)[4], bool alfven) {
  T cx[4], cy[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const T ir = rcp(q[s].r);
    if (alfven) {
      cx[s] = ab(q[s].a) * sqr_t(ir);
      cy[s] = ab(q[s].b) * sqr_t(ir);
    } else {
      const T bb = q[s].a * q[s].a + q[s].b * q[s].b + q[s].c * q[s].c;
      cx[s] = fast_speed(P.gamma0, q[s].p, ir, bb, q[s].a * q[s].a);
      cy[s] = fast_speed(P.gamma0, q[s].p, ir, bb, q[s].b * q[s].b);
    }
  }
  T cxm = max4(cx[0], cx[1], cx[2], cx[3]), cym = max4(cy[0], cy[1], cy[2], cy[3]);
  if (alfven) { cxm = mx(cxm, P.smallc); cym = mx(cym, P.smallc); }
  const T SL = min0(min4(q[0].u, q[1].u, q[2].u, q[3].u) - cxm);
  const T SR = max0(max4(q[0].u, q[1].u, q[2].u, q[3].u) + cxm);
  const T SB = min0(min4(q[0].v, q[1].v, q[2].v, q[3].v) - cym);
  const T ST = max0(max4(q[0].v, q[1].v, q[2].v, q[3].v) + cym);
  T e[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) e[s] = q[s].u * q[s].b - q[s].v * q[s].a;
  const T ix = rcp(SR - SL), iy = rcp(ST - SB);
  // 0 = LL, 1 = RL, 2 = LR, 3 = RR
  return (SL * SB * e[3] - SL * ST * e[1] - SR * SB * e[2] + SR * ST * e[0]) * ix * iy -
         ST * SB * iy * (q[3].a - q[0].a) + SR * SL * ix * (q[3].b - q[0].b);
}

// LLF 2-D solver, reference riemann_mhd.h:518-609
template <typename T>
__device__ T mag_riemann2d_llf(const KParams<T>& P, const Corner<T> (&q)[4]) {
  T E = T(0);
#pragma unroll
  for (int s = 0; s < 4; ++s) E += q[s].u * q[s].b - q[s].v * q[s].a;
  E *= T(0.25);
  const Corner<T>&LL = q[0], &RL = q[1], &LR = q[2], &RR = q[3];
  State<T> l, r;
  T fx[8], fy[8];
  const T h = T(0.5);
  l = {h * (LL.r + LR.r), h * (LL.p + LR.p), h * (LL.u + LR.u), h * (LL.v + LR.v), h * (LL.w + LR.w),
       h * (LL.a + LR.a), h * (LL.b + LR.b), h * (LL.c + LR.c)};
  r = {h * (RR.r + RL.r), h * (RR.p + RL.p), h * (RR.u + RL.u), h * (RR.v + RL.v), h * (RR.w + RL.w),
       h * (RR.a + RL.a), h * (RR.b + RL.b), h * (RR.c + RL.c)};
  riemann_llf(P, l, r, fx, T(0));
  l = {h * (LL.r + RL.r), h * (LL.p + RL.p), h * (LL.v + RL.v), h * (LL.u + RL.u), h * (LL.w + RL.w),
       h * (LL.b + RL.b), h * (LL.a + RL.a), h * (LL.c + RL.c)};
  r = {h * (RR.r + LR.r), h * (RR.p + LR.p), h * (RR.v + LR.v), h * (RR.u + LR.u), h * (RR.w + LR.w),
       h * (RR.b + LR.b), h * (RR.a + LR.a), h * (RR.c + LR.c)};
  riemann_llf(P, l, r, fy, T(0));
  return E + (fx[IB] - fy[IB]);
}

// Edge emf from the four surrounding edge states, reference riemann_mhd.h:1054-1193
// (compute_emf<dir>).  Inputs are already permuted to the edge frame by the caller:
//   s[0] = RT-state of cell (-1,-1), s[1] = RB of (-1,0), s[2] = LT of (0,-1), s[3] = LB of (0,0)
// in the reference's (IRT, IRB, ILT, ILB) order, each as (r, p, u, v, w, a, b, c) edge-frame.
// emfDir: 0 = X, 1 = Y, 2 = Z (shear terms only).
template <bool FAST = false, bool HLLD_ONLY = false, typename T>
__device__ __forceinline__ T compute_emf(const KParams<T>& P, const Corner<T>& RT, const Corner<T>& RB,
                                         const Corner<T>& LT, const Corner<T>& LB, int emfDir, T xPos) {
  Corner<T> q[4];  // LL <- RT, RL <- LT, LR <- RB, RR <- LB
  q[0] = RT; q[1] = LT; q[2] = RB; q[3] = LB;
  if (!FAST && P.cIso > T(0)) {
#pragma unroll
    for (int s = 0; s < 4; ++s) q[s].p = q[s].r * P.cIso * P.cIso;
  }
  const T aT = T(0.5) * (RT.a + LT.a), aB = T(0.5) * (RB.a + LB.a);
  const T bR = T(0.5) * (RT.b + RB.b), bL = T(0.5) * (LT.b + LB.b);
  q[0].a = aT; q[1].a = aT; q[2].a = aB; q[3].a = aB;
  q[0].b = bR; q[1].b = bL; q[2].b = bR; q[3].b = bL;
  T emf = T(0);
  if (FAST || HLLD_ONLY || P.magRiemannSolver == MAG_HLLD) emf = mag_riemann2d_hlld(P, q[0], q[1], q[2], q[3]);
  else if (P.magRiemannSolver == MAG_HLLA) emf = mag_riemann2d_hll(P, q, true);
  else if (P.magRiemannSolver == MAG_HLLF) emf = mag_riemann2d_hll(P, q, false);
  else if (P.magRiemannSolver == MAG_LLF) emf = mag_riemann2d_llf(P, q);
  if (!FAST && P.Omega0 > T(0)) {  // shearing-box upwind terms, riemann_mhd.h:1171-1189
    if (emfDir == 0) {
      const T shear = T(-1.5) * P.Omega0 * xPos;
      emf += shear * (shear > T(0) ? q[0].b : q[3].b);
    }
    if (emfDir == 2) {
      const T shear = T(-1.5) * P.Omega0 * (xPos - P.dx * T(0.5));
      emf -= shear * (shear > T(0) ? q[0].a : q[3].a);
    }
  }
  return emf;
}

}  // namespace dev
}  // namespace rg
