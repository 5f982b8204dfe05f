/*
 * oracle_hydro.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * 3D Euler (hydro) unsplit Godunov step "version 1" of the reference CPU path, restated with the
 * same floating-point operation order:
 *   HydroRunGodunov.cpp:2658-2949 (godunov_unsplit_cpu_v1, THREE_D), slope.h:324-427,
 *   trace.h:544-661, riemann.h:31-401 (approx / hll / hllc), cmpflx.h:23-49,
 *   constoprim.h:24-111, HydroRunBase.cpp:386-426 (compute_dt).
 */
#include "oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef ORACLE_FLOAT
#define R(x) x##f
#define FMAX_ fmaxf
#define FMIN_ fminf
#define SQRT_ sqrtf
#define FABS_ fabsf
#define COPYSIGN_ copysignf
#else
#define R(x) x
#define FMAX_ fmax
#define FMIN_ fmin
#define SQRT_ sqrt
#define FABS_ fabs
#define COPYSIGN_ copysign
#endif
#define HALF R(0.5)
#define ZERO R(0.0)
#define ONE R(1.0)

#define AT(arr, i, j, k, v) (arr)[(size_t)(i) + (size_t)isz * ((size_t)(j) + (size_t)jsz * ((size_t)(k) + (size_t)ksz * (size_t)(v)))]

/* gpu_macros.cpp:25-30: note the FLOAT argument even in the double build */
static float saturate_cpu(float a) {
  if (isnan(a)) return 0.0f;
  return a >= 1.0f ? 1.0f : a <= 0.0f ? 0.0f : a;
}

/* constoprim.h:24-28 eos + :82-111 constoprim_3D */
static void constoprim_3d(const orc_params *P, const real_t u[5], real_t q[5], real_t *c) {
  q[ID] = FMAX_(u[ID], P->smallr);
  q[IU] = u[IU] / q[ID]; q[IV] = u[IV] / q[ID]; q[IW] = u[IW] / q[ID];
  real_t eken = 0.5f * (q[IU] * q[IU] + q[IV] * q[IV] + q[IW] * q[IW]);
  if (P->cIso > 0) {
    q[IP] = q[ID] * (P->cIso) * (P->cIso);
    *c = P->cIso;
  } else {
    real_t eint = u[IP] / q[ID] - eken;
    q[IP] = FMAX_((P->gamma0 - 1.0f) * q[ID] * eint, q[ID] * P->smallp);
    *c = SQRT_(P->gamma0 * q[IP] / q[ID]);
  }
}

/* constoprim.h:43-71 constoprim_2D */
static void constoprim_2d(const orc_params *P, const real_t u[4], real_t q[4], real_t *c) {
  q[ID] = FMAX_(u[ID], P->smallr);
  q[IU] = u[IU] / q[ID]; q[IV] = u[IV] / q[ID];
  real_t eken = 0.5f * (q[IU] * q[IU] + q[IV] * q[IV]);
  if (P->cIso > 0) {
    q[IP] = q[ID] * (P->cIso) * (P->cIso);
    *c = P->cIso;
  } else {
    real_t eint = u[IP] / q[ID] - eken;
    q[IP] = FMAX_((P->gamma0 - 1.0f) * q[ID] * eint, q[ID] * P->smallp);
    *c = SQRT_(P->gamma0 * q[IP] / q[ID]);
  }
}

/* cmpflx.h:23-49 */
static void cmpflx3(const orc_params *P, const real_t g[5], real_t f[5]) {
  f[ID] = g[ID] * g[IU];
  f[IU] = f[ID] * g[IU] + g[IP];
  f[IV] = f[ID] * g[IV];
  f[IW] = f[ID] * g[IW];
  real_t entho = (real_t)1.0f / (P->gamma0 - 1.0f);
  real_t ekin = 0.5f * g[ID] * (g[IU] * g[IU] + g[IV] * g[IV] + g[IW] * g[IW]);
  real_t etot = g[IP] * entho + ekin;
  f[IP] = g[IU] * (etot + g[IP]);
}

/* riemann.h:31-160 */
static void riemann_approx3(const orc_params *P, const real_t ql[5], const real_t qr[5], real_t flux[5]) {
  real_t g[5];
  real_t rl = FMAX_(ql[ID], P->smallr), ul = ql[IU], pl = FMAX_(ql[IP], rl * P->smallp);
  real_t rr = FMAX_(qr[ID], P->smallr), ur = qr[IU], pr = FMAX_(qr[IP], rr * P->smallp);
  real_t cl = P->gamma0 * pl * rl, cr = P->gamma0 * pr * rr;
  real_t wl = SQRT_(cl), wr = SQRT_(cr);
  real_t pstar = FMAX_(((wr * pl + wl * pr) + wl * wr * (ul - ur)) / (wl + wr), (real_t)ZERO);
  real_t pold = pstar, conv = ONE;
  for (int iter = 0; iter < P->niter_riemann && conv > 1e-6; ++iter) {
    real_t wwl = SQRT_(cl * (ONE + P->gamma6 * (pold - pl) / pl));
    real_t wwr = SQRT_(cr * (ONE + P->gamma6 * (pold - pr) / pr));
    real_t q_l = 2.0f * wwl * wwl * wwl / (wwl * wwl + cl);
    real_t q_r = 2.0f * wwr * wwr * wwr / (wwr * wwr + cr);
    real_t usl = ul - (pold - pl) / wwl;
    real_t usr = ur + (pold - pr) / wwr;
    real_t delp = FMAX_(q_r * q_l / (q_r + q_l) * (usl - usr), -pold);
    pold = pold + delp;
    conv = FABS_(delp / (pold + P->smallpp));
  }
  pstar = pold;
  wl = SQRT_(cl * (ONE + P->gamma6 * (pstar - pl) / pl));
  wr = SQRT_(cr * (ONE + P->gamma6 * (pstar - pr) / pr));
  real_t ustar = HALF * (ul + (pl - pstar) / wl + ur - (pr - pstar) / wr);
  real_t sgnm = COPYSIGN_(ONE, ustar);
  real_t ro, uo, po, wo;
  if (sgnm > ZERO) { ro = rl; uo = ul; po = pl; wo = wl; } else { ro = rr; uo = ur; po = pr; wo = wr; }
  real_t co = FMAX_(P->smallc, SQRT_(FABS_(P->gamma0 * po / ro)));
  real_t rstar = FMAX_((real_t)(ro / (ONE + ro * (po - pstar) / (wo * wo))), (real_t)(P->smallr));
  real_t cstar = FMAX_(P->smallc, SQRT_(FABS_(P->gamma0 * pstar / rstar)));
  real_t spout = co - sgnm * uo;
  real_t spin = cstar - sgnm * ustar;
  real_t ushock = wo / ro - sgnm * uo;
  if (pstar >= po) { spin = ushock; spout = ushock; }
  real_t scr = FMAX_(spout - spin, P->smallc + FABS_(spout + spin));
  real_t frac = HALF * (ONE + (spout + spin) / scr);
  if (frac != frac) frac = ZERO; else frac = saturate_cpu(frac); /* float round trip, as the reference */
  g[ID] = frac * rstar + (ONE - frac) * ro;
  g[IU] = frac * ustar + (ONE - frac) * uo;
  g[IP] = frac * pstar + (ONE - frac) * po;
  if (spout < ZERO) { g[ID] = ro; g[IU] = uo; g[IP] = po; }
  if (spin > ZERO) { g[ID] = rstar; g[IU] = ustar; g[IP] = pstar; }
  if (sgnm > ZERO) { g[IV] = ql[IV]; g[IW] = ql[IW]; } else { g[IV] = qr[IV]; g[IW] = qr[IW]; }
  cmpflx3(P, g, flux);
}

/* riemann.h:177-253 */
static void riemann_hll3(const orc_params *P, const real_t ql[5], const real_t qr[5], real_t flux[5]) {
  const real_t entho = ONE / (P->gamma0 - ONE);
  real_t rl = FMAX_(ql[ID], P->smallr), ul = ql[IU], pl = FMAX_(ql[IP], rl * P->smallp);
  real_t rr = FMAX_(qr[ID], P->smallr), ur = qr[IU], pr = FMAX_(qr[IP], rr * P->smallp);
  real_t cl = SQRT_(P->gamma0 * pl / rl), cr = SQRT_(P->gamma0 * pr / rr);
  real_t SL = FMIN_(FMIN_(ul, ur) - FMAX_(cl, cr), (real_t)ZERO);
  real_t SR = FMAX_(FMAX_(ul, ur) + FMAX_(cl, cr), (real_t)ZERO);
  real_t uL[5], uR[5], fL[5], fR[5];
  uL[ID] = ql[ID]; uR[ID] = qr[ID];
  uL[IP] = ql[IP] * entho + HALF * ql[ID] * ql[IU] * ql[IU];
  uR[IP] = qr[IP] * entho + HALF * qr[ID] * qr[IU] * qr[IU];
  uL[IP] += HALF * ql[ID] * ql[IV] * ql[IV]; uR[IP] += HALF * qr[ID] * qr[IV] * qr[IV];
  uL[IP] += HALF * ql[ID] * ql[IW] * ql[IW]; uR[IP] += HALF * qr[ID] * qr[IW] * qr[IW];
  uL[IU] = ql[ID] * ql[IU]; uR[IU] = qr[ID] * qr[IU];
  uL[IV] = ql[ID] * ql[IV]; uR[IV] = qr[ID] * qr[IV];
  uL[IW] = ql[ID] * ql[IW]; uR[IW] = qr[ID] * qr[IW];
  fL[ID] = uL[IU]; fR[ID] = uR[IU];
  fL[IP] = ql[IU] * (uL[IP] + ql[IP]); fR[IP] = qr[IU] * (uR[IP] + qr[IP]);
  fL[IU] = ql[IP] + uL[IU] * ql[IU]; fR[IU] = qr[IP] + uR[IU] * qr[IU];
  fL[IV] = fL[ID] * ql[IV]; fR[IV] = fR[ID] * qr[IV];
  fL[IW] = fL[ID] * ql[IW]; fR[IW] = fR[ID] * qr[IW];
  for (int n = 0; n < 5; ++n) flux[n] = (SR * fL[n] - SL * fR[n] + SR * SL * (uR[n] - uL[n])) / (SR - SL);
}

/* riemann.h:270-371 */
static void riemann_hllc3(const orc_params *P, const real_t ql[5], const real_t qr[5], real_t flux[5]) {
  const real_t entho = ONE / (P->gamma0 - ONE);
  real_t rl = FMAX_(ql[ID], P->smallr), pl = FMAX_(ql[IP], rl * P->smallp), ul = ql[IU];
  real_t ecinl = HALF * rl * ul * ul;
  ecinl += HALF * rl * ql[IV] * ql[IV];
  ecinl += HALF * rl * ql[IW] * ql[IW];
  real_t etotl = pl * entho + ecinl, ptotl = pl;
  real_t rr = FMAX_(qr[ID], P->smallr), pr = FMAX_(qr[IP], rr * P->smallp), ur = qr[IU];
  real_t ecinr = HALF * rr * ur * ur;
  ecinr += HALF * rr * qr[IV] * qr[IV];
  ecinr += HALF * rr * qr[IW] * qr[IW];
  real_t etotr = pr * entho + ecinr, ptotr = pr;
  real_t cfastl = SQRT_(FMAX_(P->gamma0 * pl / rl, P->smallc * P->smallc));
  real_t cfastr = SQRT_(FMAX_(P->gamma0 * pr / rr, P->smallc * P->smallc));
  real_t SL = FMIN_(ul, ur) - FMAX_(cfastl, cfastr);
  real_t SR = FMAX_(ul, ur) + FMAX_(cfastl, cfastr);
  real_t rcl = rl * (ul - SL), rcr = rr * (SR - ur);
  real_t ustar = (rcr * ur + rcl * ul + (ptotl - ptotr)) / (rcr + rcl);
  real_t ptotstar = (rcr * ptotl + rcl * ptotr + rcl * rcr * (ul - ur)) / (rcr + rcl);
  real_t rstarl = rl * (SL - ul) / (SL - ustar);
  real_t etotstarl = ((SL - ul) * etotl - ptotl * ul + ptotstar * ustar) / (SL - ustar);
  real_t rstarr = rr * (SR - ur) / (SR - ustar);
  real_t etotstarr = ((SR - ur) * etotr - ptotr * ur + ptotstar * ustar) / (SR - ustar);
  real_t ro, uo, ptoto, etoto;
  if (SL > ZERO) { ro = rl; uo = ul; ptoto = ptotl; etoto = etotl; }
  else if (ustar > ZERO) { ro = rstarl; uo = ustar; ptoto = ptotstar; etoto = etotstarl; }
  else if (SR > ZERO) { ro = rstarr; uo = ustar; ptoto = ptotstar; etoto = etotstarr; }
  else { ro = rr; uo = ur; ptoto = ptotr; etoto = etotr; }
  flux[ID] = ro * uo;
  flux[IU] = ro * uo * uo + ptoto;
  flux[IP] = (etoto + ptoto) * uo;
  flux[IV] = (flux[ID] > ZERO) ? flux[ID] * ql[IV] : flux[ID] * qr[IV];
  flux[IW] = (flux[ID] > ZERO) ? flux[ID] * ql[IW] : flux[ID] * qr[IW];
}

/* riemann.h:388-401 */
void orc_riemann_hydro(const orc_params *P, const real_t ql[5], const real_t qr[5], real_t flux[5]) {
  for (int n = 0; n < 5; ++n) flux[n] = 0;
  if (P->riemannSolver == RS_APPROX) riemann_approx3(P, ql, qr, flux);
  else if (P->riemannSolver == RS_HLL) riemann_hll3(P, ql, qr, flux);
  else if (P->riemannSolver == RS_HLLC) riemann_hllc3(P, ql, qr, flux);
}

/* slope.h:324-427 (one variable, one direction) */
static real_t slope1(const orc_params *P, real_t qm, real_t q0, real_t qp) {
  if (P->slope_type == 0) return ZERO;
  if (P->slope_type == 1) {
    real_t dlft = q0 - qm, drgt = qp - q0;
    if ((dlft * drgt) <= ZERO) return ZERO;
    else if (dlft > 0) return FMIN_(dlft, drgt);
    else return FMAX_(dlft, drgt);
  } else if (P->slope_type == 2) {
    real_t dlft = P->slope_type * (q0 - qm), drgt = P->slope_type * (qp - q0);
    real_t dcen = HALF * (qp - qm);
    real_t dsgn = (dcen >= ZERO) ? ONE : -ONE;
    real_t slop = FMIN_(FABS_(dlft), FABS_(drgt));
    real_t dlim = slop;
    if ((dlft * drgt) <= ZERO) dlim = ZERO;
    return dsgn * FMIN_(dlim, FABS_(dcen));
  }
  return ZERO; /* other slope types leave dq untouched (zero-initialised stack in practice) */
}

/* slope.h slope_unsplit_hydro_2d: slope types 1 and 2 share one formula */
static real_t slope2d(const orc_params *P, real_t qm, real_t q0, real_t qp) {
  if (P->slope_type == 0) return ZERO;
  if (P->slope_type == 1 || P->slope_type == 2) {
    real_t dlft = P->slope_type * (q0 - qm);
    real_t drgt = P->slope_type * (qp - q0);
    real_t dcen = HALF * (qp - qm);
    real_t dsgn = (dcen >= ZERO) ? ONE : -ONE;
    real_t slop = FMIN_(FABS_(dlft), FABS_(drgt));
    real_t dlim = slop;
    if ((dlft * drgt) <= ZERO) dlim = ZERO;
    return dsgn * FMIN_(dlim, FABS_(dcen));
  }
  return ZERO;
}

/* HydroRunBase.cpp:386-426 (3D and 2D branches) */
real_t orc_compute_dt_hydro(const orc_params *P, const real_t *U) {
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize, gw = P->ghostWidth;
  real_t invDt = 0;
  if (P->dim != 3) { /* HydroRunBase.cpp:386-399 */
    for (int j = gw; j < jsz - gw; ++j)
      for (int i = gw; i < isz - gw; ++i) {
        real_t u[4], q[4], c;
        for (int v = 0; v < 4; ++v) u[v] = AT(U, i, j, 0, v);
        constoprim_2d(P, u, q, &c);
        real_t vx = c + FABS_(q[IU]), vy = c + FABS_(q[IV]);
        invDt = FMAX_(invDt, vx / P->dx + vy / P->dy);
      }
    if (P->enableJet) invDt = FMAX_(invDt, (P->ujet + P->cjet) / P->dx);
    return P->cfl / invDt;
  }
  for (int k = gw; k < ksz - gw; ++k)
    for (int j = gw; j < jsz - gw; ++j)
      for (int i = gw; i < isz - gw; ++i) {
        real_t u[5], q[5], c;
        for (int v = 0; v < 5; ++v) u[v] = AT(U, i, j, k, v);
        constoprim_3d(P, u, q, &c);
        real_t vx = c + FABS_(q[IU]), vy = c + FABS_(q[IV]), vz = c + FABS_(q[IW]);
        invDt = FMAX_(invDt, vx / P->dx + vy / P->dy + vz / P->dz);
      }
  /* the inflow speed of the jet limits the step too: HydroRunBase.cpp:420-422, MHDRunBase.cpp:184-186,228-230 */
  if (P->enableJet) invDt = FMAX_(invDt, (P->ujet + P->cjet) / P->dx);
  return P->cfl / invDt;
}

void orc_dissipative_3d(const orc_params *P, real_t *Unew, real_t dt, real_t totalTime, int shear);

/* HydroRunGodunov.cpp:2658-2890 + convertToPrimitives :4210-4250 */
/* 2D: HydroRunGodunov.cpp:2446-2655 (godunov_unsplit_cpu_v1, TWO_D) + trace.h:332-414; the Riemann solvers are the
 * 3D ones with w = 0 (riemann<NVAR_2D> performs the same operations without the third velocity) */
static void hydro2d_step_v1(const orc_params *P, const real_t *Uold, real_t *Unew, real_t dt) {
  const int isz = P->isize, jsz = P->jsize, ksz = 1, gw = P->ghostWidth;
  const size_t ncell = (size_t)isz * jsz;
  const real_t dtdx = dt / P->dx, dtdy = dt / P->dy;
  real_t *Q = calloc(ncell * 4, sizeof(real_t));
  real_t *tr = calloc(ncell * 4 * 4, sizeof(real_t));
  real_t *qm_[2] = {tr, tr + ncell * 4}, *qp_[2] = {tr + ncell * 8, tr + ncell * 12};
  for (int j = 0; j < jsz; ++j)
    for (int i = 0; i < isz; ++i) {
      real_t u[4], q[4], c;
      for (int v = 0; v < 4; ++v) u[v] = AT(Uold, i, j, 0, v);
      constoprim_2d(P, u, q, &c);
      for (int v = 0; v < 4; ++v) AT(Q, i, j, 0, v) = q[v];
    }
  for (int j = 1; j < jsz - 1; ++j)
    for (int i = 1; i < isz - 1; ++i) {
      real_t q[4], d[2][4];
      for (int v = 0; v < 4; ++v) {
        q[v] = AT(Q, i, j, 0, v);
        d[0][v] = HALF * slope2d(P, AT(Q, i - 1, j, 0, v), q[v], AT(Q, i + 1, j, 0, v));
        d[1][v] = HALF * slope2d(P, AT(Q, i, j - 1, 0, v), q[v], AT(Q, i, j + 1, 0, v));
      }
      real_t r = q[ID], p = q[IP], u = q[IU], v_ = q[IV];
      const real_t drx = d[0][ID], dpx = d[0][IP], dux = d[0][IU], dvx = d[0][IV];
      const real_t dry = d[1][ID], dpy = d[1][IP], duy = d[1][IU], dvy = d[1][IV];
      const real_t gamma = P->gamma0;
      real_t sr0 = (-u * drx - dux * r) * dtdx + (-v_ * dry - dvy * r) * dtdy;
      real_t su0 = (-u * dux - dpx / r) * dtdx + (-v_ * duy) * dtdy;
      real_t sv0 = (-u * dvx) * dtdx + (-v_ * dvy - dpy / r) * dtdy;
      real_t sp0 = (-u * dpx - dux * gamma * p) * dtdx + (-v_ * dpy - dvy * gamma * p) * dtdy;
      r = r + sr0; u = u + su0; v_ = v_ + sv0; p = p + sp0;
      real_t gf[3];
      orc_gravity_cell(P, i, j, 0, gf);
      for (int dd = 0; dd < 2; ++dd)
        for (int side = 0; side < 2; ++side) { /* side 0: qp (low face), 1: qm (high face) */
          real_t *dst = side ? qm_[dd] : qp_[dd];
          real_t rr = side ? r + d[dd][ID] : r - d[dd][ID];
          real_t uu = side ? u + d[dd][IU] : u - d[dd][IU];
          real_t vv = side ? v_ + d[dd][IV] : v_ - d[dd][IV];
          real_t pp = side ? p + d[dd][IP] : p - d[dd][IP];
          rr = FMAX_(P->smallr, rr);
          pp = FMAX_(P->smallp * rr, pp);
          if (P->gravityEnabled) { uu += HALF * dt * gf[0]; vv += HALF * dt * gf[1]; }
          AT(dst, i, j, 0, ID) = rr; AT(dst, i, j, 0, IP) = pp; AT(dst, i, j, 0, IU) = uu; AT(dst, i, j, 0, IV) = vv;
        }
    }
  for (int j = gw; j < jsz - gw + 1; ++j)
    for (int i = gw; i < isz - gw + 1; ++i) {
      real_t ql[5], qr[5], fx[5], fy[5];
      ql[ID] = AT(qm_[0], i - 1, j, 0, ID); ql[IP] = AT(qm_[0], i - 1, j, 0, IP); ql[IU] = AT(qm_[0], i - 1, j, 0, IU); ql[IV] = AT(qm_[0], i - 1, j, 0, IV); ql[IW] = 0;
      qr[ID] = AT(qp_[0], i, j, 0, ID); qr[IP] = AT(qp_[0], i, j, 0, IP); qr[IU] = AT(qp_[0], i, j, 0, IU); qr[IV] = AT(qp_[0], i, j, 0, IV); qr[IW] = 0;
      orc_riemann_hydro(P, ql, qr, fx);
      ql[ID] = AT(qm_[1], i, j - 1, 0, ID); ql[IP] = AT(qm_[1], i, j - 1, 0, IP); ql[IU] = AT(qm_[1], i, j - 1, 0, IV); ql[IV] = AT(qm_[1], i, j - 1, 0, IU); ql[IW] = 0;
      qr[ID] = AT(qp_[1], i, j, 0, ID); qr[IP] = AT(qp_[1], i, j, 0, IP); qr[IU] = AT(qp_[1], i, j, 0, IV); qr[IV] = AT(qp_[1], i, j, 0, IU); qr[IW] = 0;
      orc_riemann_hydro(P, ql, qr, fy);
      const int in_i = i < isz - gw, in_j = j < jsz - gw;
      static const int sw[4] = {ID, IP, IV, IU};
      if (i > gw && in_j) for (int v = 0; v < 4; ++v) AT(Unew, i - 1, j, 0, v) -= fx[v] * dtdx;
      if (in_i && in_j) for (int v = 0; v < 4; ++v) AT(Unew, i, j, 0, v) += fx[v] * dtdx;
      if (in_i && j > gw) for (int v = 0; v < 4; ++v) AT(Unew, i, j - 1, 0, v) -= fy[sw[v]] * dtdy;
      if (in_i && in_j) for (int v = 0; v < 4; ++v) AT(Unew, i, j, 0, v) += fy[sw[v]] * dtdy;
    }
  if (P->gravityEnabled) { /* HydroRunBase.cpp:1946-1958 */
    real_t gf[3];
    for (int j = gw; j < jsz - gw; ++j)
      for (int i = gw; i < isz - gw; ++i) {
        orc_gravity_cell(P, i, j, 0, gf);
        real_t rhoOld = AT(Uold, i, j, 0, ID), rhoNew = AT(Unew, i, j, 0, ID);
        AT(Unew, i, j, 0, IU) += HALF * dt * gf[0] * (rhoOld + rhoNew);
        AT(Unew, i, j, 0, IV) += HALF * dt * gf[1] * (rhoOld + rhoNew);
      }
  }
  free(Q); free(tr);
}

void orc_hydro_step_v1(const orc_params *P, const real_t *Uold, real_t *Unew, real_t dt) {
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize, gw = P->ghostWidth;
  if (P->dim != 3) { hydro2d_step_v1(P, Uold, Unew, dt); return; }
  const size_t ncell = (size_t)isz * jsz * ksz;
  const real_t dtdx = dt / P->dx, dtdy = dt / P->dy, dtdz = dt / P->dz;
  real_t *Q = calloc(ncell * 5, sizeof(real_t));
  real_t *tr = calloc(ncell * 5 * 6, sizeof(real_t));
  real_t *qm_[3] = {tr, tr + ncell * 5, tr + ncell * 10}, *qp_[3] = {tr + ncell * 15, tr + ncell * 20, tr + ncell * 25};
  for (int k = 0; k < ksz; ++k)
    for (int j = 0; j < jsz; ++j)
      for (int i = 0; i < isz; ++i) {
        real_t u[5], q[5], c;
        for (int v = 0; v < 5; ++v) u[v] = AT(Uold, i, j, k, v);
        constoprim_3d(P, u, q, &c);
        for (int v = 0; v < 5; ++v) AT(Q, i, j, k, v) = q[v];
      }
  /* slopes + trace: loop 1..size-2 (:2663-2665), trace.h:544-661 */
  for (int k = 1; k < ksz - 1; ++k)
    for (int j = 1; j < jsz - 1; ++j)
      for (int i = 1; i < isz - 1; ++i) {
        real_t q[5], d[3][5];
        for (int v = 0; v < 5; ++v) {
          q[v] = AT(Q, i, j, k, v);
          d[0][v] = HALF * slope1(P, AT(Q, i - 1, j, k, v), q[v], AT(Q, i + 1, j, k, v));
          d[1][v] = HALF * slope1(P, AT(Q, i, j - 1, k, v), q[v], AT(Q, i, j + 1, k, v));
          d[2][v] = HALF * slope1(P, AT(Q, i, j, k - 1, v), q[v], AT(Q, i, j, k + 1, v));
        }
        real_t r = q[ID], p = q[IP], u = q[IU], v_ = q[IV], w = q[IW];
        real_t drx = d[0][ID], dpx = d[0][IP], dux = d[0][IU], dvx = d[0][IV], dwx = d[0][IW];
        real_t dry = d[1][ID], dpy = d[1][IP], duy = d[1][IU], dvy = d[1][IV], dwy = d[1][IW];
        real_t drz = d[2][ID], dpz = d[2][IP], duz = d[2][IU], dvz = d[2][IV], dwz = d[2][IW];
        const real_t gamma = P->gamma0;
        real_t sr0 = (-u * drx - dux * r) * dtdx + (-v_ * dry - dvy * r) * dtdy + (-w * drz - dwz * r) * dtdz;
        real_t su0 = (-u * dux - dpx / r) * dtdx + (-v_ * duy) * dtdy + (-w * duz) * dtdz;
        real_t sv0 = (-u * dvx) * dtdx + (-v_ * dvy - dpy / r) * dtdy + (-w * dvz) * dtdz;
        real_t sw0 = (-u * dwx) * dtdx + (-v_ * dwy) * dtdy + (-w * dwz - dpz / r) * dtdz;
        real_t sp0 = (-u * dpx - dux * gamma * p) * dtdx + (-v_ * dpy - dvy * gamma * p) * dtdy + (-w * dpz - dwz * gamma * p) * dtdz;
        r = r + sr0; u = u + su0; v_ = v_ + sv0; w = w + sw0; p = p + sp0;
        /* gravity predictor, added to qm/qp AFTER the trace (HydroRunGodunov.cpp:2705-2734) */
        real_t gf[3];
        orc_gravity_at(P, k, gf);
        const real_t gpx = P->gravityEnabled ? HALF * dt * gf[0] : 0;
        const real_t gpy = P->gravityEnabled ? HALF * dt * gf[1] : 0;
        const real_t gpz = P->gravityEnabled ? HALF * dt * gf[2] : 0;
        for (int dd = 0; dd < 3; ++dd) {
          real_t s[2] = {-ONE, ONE};
          for (int side = 0; side < 2; ++side) { /* side 0: qp (low face), 1: qm (high face) */
            real_t *dst = side ? qm_[dd] : qp_[dd];
            real_t sg = s[side];
            real_t rr = (sg > 0) ? r + d[dd][ID] : r - d[dd][ID];
            real_t uu = (sg > 0) ? u + d[dd][IU] : u - d[dd][IU];
            real_t vv = (sg > 0) ? v_ + d[dd][IV] : v_ - d[dd][IV];
            real_t ww = (sg > 0) ? w + d[dd][IW] : w - d[dd][IW];
            real_t pp = (sg > 0) ? p + d[dd][IP] : p - d[dd][IP];
            rr = FMAX_(P->smallr, rr);
            pp = FMAX_(P->smallp * rr, pp);
            AT(dst, i, j, k, ID) = rr; AT(dst, i, j, k, IP) = pp;
            if (P->gravityEnabled) { uu += gpx; vv += gpy; ww += gpz; }
            AT(dst, i, j, k, IU) = uu; AT(dst, i, j, k, IV) = vv; AT(dst, i, j, k, IW) = ww;
          }
        }
      }
  /* fluxes + update :2760-2890 */
  for (int k = gw; k < ksz - gw + 1; ++k)
    for (int j = gw; j < jsz - gw + 1; ++j)
      for (int i = gw; i < isz - gw + 1; ++i) {
        real_t ql[5], qr[5], fx[5], fy[5], fz[5];
        for (int v = 0; v < 5; ++v) { ql[v] = AT(qm_[0], i - 1, j, k, v); qr[v] = AT(qp_[0], i, j, k, v); }
        orc_riemann_hydro(P, ql, qr, fx);
        static const int swy[5] = {ID, IP, IV, IU, IW};
        for (int v = 0; v < 5; ++v) { ql[v] = AT(qm_[1], i, j - 1, k, swy[v]); qr[v] = AT(qp_[1], i, j, k, swy[v]); }
        orc_riemann_hydro(P, ql, qr, fy);
        static const int swz[5] = {ID, IP, IW, IV, IU};
        for (int v = 0; v < 5; ++v) { ql[v] = AT(qm_[2], i, j, k - 1, swz[v]); qr[v] = AT(qp_[2], i, j, k, swz[v]); }
        orc_riemann_hydro(P, ql, qr, fz);
        const int in_i = i < isz - gw, in_j = j < jsz - gw, in_k = k < ksz - gw;
        if (i > gw && in_j && in_k) for (int v = 0; v < 5; ++v) AT(Unew, i - 1, j, k, v) -= fx[v] * dtdx;
        if (in_i && in_j && in_k) for (int v = 0; v < 5; ++v) AT(Unew, i, j, k, v) += fx[v] * dtdx;
        if (in_i && j > gw && in_k) for (int v = 0; v < 5; ++v) AT(Unew, i, j - 1, k, v) -= fy[swy[v]] * dtdy;
        if (in_i && in_j && in_k) for (int v = 0; v < 5; ++v) AT(Unew, i, j, k, v) += fy[swy[v]] * dtdy;
        if (in_i && in_j && k > gw) for (int v = 0; v < 5; ++v) AT(Unew, i, j, k - 1, v) -= fz[swz[v]] * dtdz;
        if (in_i && in_j && in_k) for (int v = 0; v < 5; ++v) AT(Unew, i, j, k, v) += fz[swz[v]] * dtdz;
      }
  /* gravity source term, HydroRunGodunov.cpp:2900-2903 -> HydroRunBase.cpp:1962-1976 */
  if (P->gravityEnabled)
    for (int k = gw; k < ksz - gw; ++k)
      for (int j = gw; j < jsz - gw; ++j)
        for (int i = gw; i < isz - gw; ++i) {
          real_t rhoOld = AT(Uold, i, j, k, ID), rhoNew = AT(Unew, i, j, k, ID);
          real_t gf[3];
          orc_gravity_at(P, k, gf);
          AT(Unew, i, j, k, IU) += HALF * dt * gf[0] * (rhoOld + rhoNew);
          AT(Unew, i, j, k, IV) += HALF * dt * gf[1] * (rhoOld + rhoNew);
          AT(Unew, i, j, k, IW) += HALF * dt * gf[2] * (rhoOld + rhoNew);
        }
  free(Q); free(tr);
  orc_dissipative_3d(P, Unew, dt, 0, 0); /* viscosity, HydroRunGodunov.cpp:2908-2927 */
}
