/*
 * oracle_mhd.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Point-wise MHD numerics and the 3D MHD unsplit step, restated from the reference CPU path
 * with the SAME floating-point operation order (so that it can be pinned against
 * oracle/_ref/euler_cpu).  Compile with -ffp-contract=off.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef ORACLE_FLOAT
#define R(x) x##f
#define FMAX_ fmaxf
#define FMIN_ fminf
#define SQRT_ sqrtf
#define FABS_ fabsf
#define COPYSIGN_ copysignf
#define FMOD_ fmodf
#else
#define R(x) x
#define FMAX_ fmax
#define FMIN_ fmin
#define SQRT_ sqrt
#define FABS_ fabs
#define COPYSIGN_ copysign
#define FMOD_ fmod
#endif
#define HALF R(0.5)
#define ZERO R(0.0)
#define ONE R(1.0)
#define FOURTH R(0.25)
#ifdef ORACLE_FLOAT
#define fmodf_ fmodf
#endif

/* ------------------------------------------------------------------------------------------
 * mhd_utils.h:28-52  find_speed_fast<dir>
 * ---------------------------------------------------------------------------------------- */
static real_t fast_speed(const orc_params *P, const real_t q[8], int dir) {
  real_t d = q[ID], p = q[IP], a = q[IA], b = q[IB], c = q[IC];
  real_t b2 = a * a + b * b + c * c;
  real_t c2 = P->gamma0 * p / d;
  real_t d2 = HALF * (b2 / d + c2);
  real_t n = (dir == 0) ? a : (dir == 1) ? b : c;
  return SQRT_(d2 + SQRT_(d2 * d2 - c2 * n * n / d));
}

/* mhd_utils.h:106-156  find_mhd_flux */
static void mhd_flux(const orc_params *P, const real_t q[8], real_t cv[8], real_t ff[8]) {
  real_t p = (P->cIso > 0) ? q[ID] * P->cIso * P->cIso : q[IP];
  const real_t entho = ONE / (P->gamma0 - ONE);
  real_t d = q[ID], u = q[IU], v = q[IV], w = q[IW], a = q[IA], b = q[IB], c = q[IC];
  real_t ecin = HALF * (u * u + v * v + w * w) * d;
  real_t emag = HALF * (a * a + b * b + c * c);
  real_t etot = p * entho + ecin + emag;
  real_t ptot = p + emag;
  cv[ID] = d; cv[IP] = etot; cv[IU] = d * u; cv[IV] = d * v; cv[IW] = d * w;
  cv[IA] = a; cv[IB] = b; cv[IC] = c;
  ff[ID] = d * u;
  ff[IP] = (etot + ptot) * u - a * (a * u + b * v + c * w);
  ff[IU] = d * u * u - a * a + ptot;
  ff[IV] = d * u * v - a * b;
  ff[IW] = d * u * w - a * c;
  ff[IA] = ZERO;
  ff[IB] = b * u - a * v;
  ff[IC] = c * u - a * w;
}

/* mhd_utils.h:295-316  find_speed_info (1-D) */
static real_t info_speed_x(const orc_params *P, const real_t q[8]) {
  return fast_speed(P, q, 0) + FABS_(q[IU]);
}

/* riemann_mhd.h:41-71 */
static void riemann_hll_mhd(const orc_params *P, real_t ql[8], real_t qr[8], real_t flux[8]) {
  real_t bm = HALF * (ql[IA] + qr[IA]);
  ql[IA] = bm; qr[IA] = bm;
  real_t ul[8], fl[8], ur[8], fr[8];
  mhd_flux(P, ql, ul, fl);
  mhd_flux(P, qr, ur, fr);
  real_t cfl_ = fast_speed(P, ql, 0), cfr = fast_speed(P, qr, 0);
  real_t vl = ql[IU], vr = qr[IU];
  real_t sl = FMIN_(FMIN_(vl, vr) - FMAX_(cfl_, cfr), ZERO);
  real_t sr = FMAX_(FMAX_(vl, vr) + FMAX_(cfl_, cfr), ZERO);
  for (int n = 0; n < 8; ++n)
    flux[n] = (sr * fl[n] - sl * fr[n] + sr * sl * (ur[n] - ul[n])) / (sr - sl);
}

/* riemann_mhd.h:86-118 */
static void riemann_llf_mhd(const orc_params *P, real_t ql[8], real_t qr[8], real_t flux[8],
                            real_t zero_flux) {
  real_t bm = HALF * (ql[IA] + qr[IA]);
  ql[IA] = bm; qr[IA] = bm;
  real_t ul[8], fl[8], ur[8], fr[8];
  mhd_flux(P, ql, ul, fl);
  mhd_flux(P, qr, ur, fr);
  for (int n = 0; n < 8; ++n) flux[n] = (ql[n] + qr[n]) / 2 * zero_flux;
  real_t vel = FMAX_(info_speed_x(P, ql), info_speed_x(P, qr));
  for (int n = 0; n < 8; ++n) flux[n] -= vel * (ur[n] - ul[n]) / 2;
}

/* riemann_mhd.h:139-342  HLLD (Miyoshi & Kusano 2005) */
static void riemann_hlld_mhd(const orc_params *P, real_t ql[8], real_t qr[8], real_t flux[8]) {
  const real_t entho = ONE / (P->gamma0 - ONE);
  real_t a = HALF * (ql[IA] + qr[IA]);
  real_t sgnm = (a >= 0) ? ONE : -ONE;
  ql[IA] = a; qr[IA] = a;
  if (P->cIso > 0) { /* :155-160 */
    ql[IP] = ql[ID] * P->cIso * P->cIso;
    qr[IP] = qr[ID] * P->cIso * P->cIso;
  }
  real_t rl = ql[ID], pl = ql[IP], ul = ql[IU], vl = ql[IV], wl = ql[IW], bl = ql[IB], cl = ql[IC];
  real_t ecinl = HALF * (ul * ul + vl * vl + wl * wl) * rl;
  real_t emagl = HALF * (a * a + bl * bl + cl * cl);
  real_t etotl = pl * entho + ecinl + emagl;
  real_t ptotl = pl + emagl;
  real_t vdotbl = ul * a + vl * bl + wl * cl;

  real_t rr = qr[ID], pr = qr[IP], ur = qr[IU], vr = qr[IV], wr = qr[IW], br = qr[IB], cr = qr[IC];
  real_t ecinr = HALF * (ur * ur + vr * vr + wr * wr) * rr;
  real_t emagr = HALF * (a * a + br * br + cr * cr);
  real_t etotr = pr * entho + ecinr + emagr;
  real_t ptotr = pr + emagr;
  real_t vdotbr = ur * a + vr * br + wr * cr;

  real_t cfastl = fast_speed(P, ql, 0), cfastr = fast_speed(P, qr, 0);
  real_t sl = FMIN_(ul, ur) - FMAX_(cfastl, cfastr);
  real_t sr = FMAX_(ul, ur) + FMAX_(cfastl, cfastr);
  real_t rcl = rl * (ul - sl), rcr = rr * (sr - ur);
  real_t ustar = (rcr * ur + rcl * ul + (ptotl - ptotr)) / (rcr + rcl);
  real_t ptotstar = (rcr * ptotl + rcl * ptotr + rcl * rcr * (ul - ur)) / (rcr + rcl);

  /* left star region :202-226 */
  real_t rstarl = rl * (sl - ul) / (sl - ustar);
  real_t estar = rl * (sl - ul) * (sl - ustar) - a * a;
  real_t el = rl * (sl - ul) * (sl - ul) - a * a;
  real_t vstarl, wstarl, bstarl, cstarl;
  if (a * a > 0 && FABS_(estar / (a * a) - ONE) <= 1e-8) {
    vstarl = vl; bstarl = bl; wstarl = wl; cstarl = cl;
  } else {
    vstarl = vl - a * bl * (ustar - ul) / estar;
    bstarl = bl * el / estar;
    wstarl = wl - a * cl * (ustar - ul) / estar;
    cstarl = cl * el / estar;
  }
  real_t vdotbstarl = ustar * a + vstarl * bstarl + wstarl * cstarl;
  real_t etotstarl = ((sl - ul) * etotl - ptotl * ul + ptotstar * ustar + a * (vdotbl - vdotbstarl)) / (sl - ustar);
  real_t sqrrstarl = SQRT_(rstarl);
  real_t calfvenl = FABS_(a) / sqrrstarl;
  real_t sal = ustar - calfvenl;

  /* right star region :228-251 */
  real_t rstarr = rr * (sr - ur) / (sr - ustar);
  estar = rr * (sr - ur) * (sr - ustar) - a * a;
  real_t er = rr * (sr - ur) * (sr - ur) - a * a;
  real_t vstarr, wstarr, bstarr, cstarr;
  if (a * a > 0 && FABS_(estar / (a * a) - ONE) <= 1e-8) {
    vstarr = vr; bstarr = br; wstarr = wr; cstarr = cr;
  } else {
    vstarr = vr - a * br * (ustar - ur) / estar;
    bstarr = br * er / estar;
    wstarr = wr - a * cr * (ustar - ur) / estar;
    cstarr = cr * er / estar;
  }
  real_t vdotbstarr = ustar * a + vstarr * bstarr + wstarr * cstarr;
  real_t etotstarr = ((sr - ur) * etotr - ptotr * ur + ptotstar * ustar + a * (vdotbr - vdotbstarr)) / (sr - ustar);
  real_t sqrrstarr = SQRT_(rstarr);
  real_t calfvenr = FABS_(a) / sqrrstarr;
  real_t sar = ustar + calfvenr;

  /* double star :253-266 */
  real_t vstarstar = (sqrrstarl * vstarl + sqrrstarr * vstarr + sgnm * (bstarr - bstarl)) / (sqrrstarl + sqrrstarr);
  real_t wstarstar = (sqrrstarl * wstarl + sqrrstarr * wstarr + sgnm * (cstarr - cstarl)) / (sqrrstarl + sqrrstarr);
  real_t bstarstar = (sqrrstarl * bstarr + sqrrstarr * bstarl + sgnm * sqrrstarl * sqrrstarr * (vstarr - vstarl)) / (sqrrstarl + sqrrstarr);
  real_t cstarstar = (sqrrstarl * cstarr + sqrrstarr * cstarl + sgnm * sqrrstarl * sqrrstarr * (wstarr - wstarl)) / (sqrrstarl + sqrrstarr);
  real_t vdotbstarstar = ustar * a + vstarstar * bstarstar + wstarstar * cstarstar;
  real_t etotstarstarl = etotstarl - sgnm * sqrrstarl * (vdotbstarl - vdotbstarstar);
  real_t etotstarstarr = etotstarr + sgnm * sqrrstarr * (vdotbstarr - vdotbstarstar);

  /* sample at x/t = 0 :268-330 */
  real_t ro, uo, vo, wo, bo, co, ptoto, etoto, vdotbo;
  if (sl > 0) {
    ro = rl; uo = ul; vo = vl; wo = wl; bo = bl; co = cl; ptoto = ptotl; etoto = etotl; vdotbo = vdotbl;
  } else if (sal > 0) {
    ro = rstarl; uo = ustar; vo = vstarl; wo = wstarl; bo = bstarl; co = cstarl;
    ptoto = ptotstar; etoto = etotstarl; vdotbo = vdotbstarl;
  } else if (ustar > 0) {
    ro = rstarl; uo = ustar; vo = vstarstar; wo = wstarstar; bo = bstarstar; co = cstarstar;
    ptoto = ptotstar; etoto = etotstarstarl; vdotbo = vdotbstarstar;
  } else if (sar > 0) {
    ro = rstarr; uo = ustar; vo = vstarstar; wo = wstarstar; bo = bstarstar; co = cstarstar;
    ptoto = ptotstar; etoto = etotstarstarr; vdotbo = vdotbstarstar;
  } else if (sr > 0) {
    ro = rstarr; uo = ustar; vo = vstarr; wo = wstarr; bo = bstarr; co = cstarr;
    ptoto = ptotstar; etoto = etotstarr; vdotbo = vdotbstarr;
  } else {
    ro = rr; uo = ur; vo = vr; wo = wr; bo = br; co = cr; ptoto = ptotr; etoto = etotr; vdotbo = vdotbr;
  }
  /* :332-340 */
  flux[ID] = ro * uo;
  flux[IP] = (etoto + ptoto) * uo - a * vdotbo;
  flux[IU] = ro * uo * uo - a * a + ptoto;
  flux[IV] = ro * uo * vo - a * bo;
  flux[IW] = ro * uo * wo - a * co;
  flux[IA] = ZERO;
  flux[IB] = bo * uo - a * vo;
  flux[IC] = co * uo - a * wo;
}

/* riemann_mhd.h:354-368: dispatch; inputs are modified in place like the reference */
static void riemann_mhd_(const orc_params *P, real_t ql[8], real_t qr[8], real_t flux[8]) {
  if (P->riemannSolver == RS_HLL) riemann_hll_mhd(P, ql, qr, flux);
  else if (P->riemannSolver == RS_LLF) riemann_llf_mhd(P, ql, qr, flux, ONE);
  else if (P->riemannSolver == RS_HLLD) riemann_hlld_mhd(P, ql, qr, flux);
}
void orc_riemann_mhd(const orc_params *P, const real_t ql_[8], const real_t qr_[8], real_t flux[8]) {
  real_t ql[8], qr[8];
  memcpy(ql, ql_, sizeof ql); memcpy(qr, qr_, sizeof qr);
  for (int n = 0; n < 8; ++n) flux[n] = 0;
  riemann_mhd_(P, ql, qr, flux);
}

/* riemann_mhd.h:373-411 */
static real_t max4(real_t a0, real_t a1, real_t a2, real_t a3) {
  real_t r = a0; r = (a1 > r) ? a1 : r; r = (a2 > r) ? a2 : r; r = (a3 > r) ? a3 : r; return r;
}
static real_t min4(real_t a0, real_t a1, real_t a2, real_t a3) {
  real_t r = a0; r = (a1 < r) ? a1 : r; r = (a2 < r) ? a2 : r; r = (a3 < r) ? a3 : r; return r;
}
static real_t max5(real_t a0, real_t a1, real_t a2, real_t a3, real_t a4) {
  real_t r = max4(a0, a1, a2, a3); r = (a4 > r) ? a4 : r; return r;
}

enum { ILL = 0, IRL = 1, ILR = 2, IRR = 3 };
enum { IRT = 0, IRB = 1, ILT = 2, ILB = 3 };

/* riemann_mhd.h:417-507  HLLA (alfven = 1) / HLLF (alfven = 0) */
static real_t mag2d_hll(const orc_params *P, real_t q[4][8], const real_t e[4], int alfven) {
  real_t cx[4], cy[4];
  for (int s = 0; s < 4; ++s) {
    if (alfven) {
      cx[s] = SQRT_(q[s][IA] * q[s][IA] / q[s][ID]);
      cy[s] = SQRT_(q[s][IB] * q[s][IB] / q[s][ID]);
    } else {
      cx[s] = fast_speed(P, q[s], 0);
      cy[s] = fast_speed(P, q[s], 1);
    }
  }
  /* argument order of the reference: LL, LR, RL, RR */
  real_t cMaxx = alfven ? max5(cx[ILL], cx[ILR], cx[IRL], cx[IRR], P->smallc) : max4(cx[ILL], cx[ILR], cx[IRL], cx[IRR]);
  real_t cMaxy = alfven ? max5(cy[ILL], cy[ILR], cy[IRL], cy[IRR], P->smallc) : max4(cy[ILL], cy[ILR], cy[IRL], cy[IRR]);
  real_t SL = FMIN_(min4(q[ILL][IU], q[ILR][IU], q[IRL][IU], q[IRR][IU]) - cMaxx, ZERO);
  real_t SR = FMAX_(max4(q[ILL][IU], q[ILR][IU], q[IRL][IU], q[IRR][IU]) + cMaxx, ZERO);
  real_t SB = FMIN_(min4(q[ILL][IV], q[ILR][IV], q[IRL][IV], q[IRR][IV]) - cMaxy, ZERO);
  real_t ST = FMAX_(max4(q[ILL][IV], q[ILR][IV], q[IRL][IV], q[IRR][IV]) + cMaxy, ZERO);
  real_t ELL = e[ILL], ERL = e[IRL], ELR = e[ILR], ERR = e[IRR];
  return (SL * SB * ERR - SL * ST * ERL - SR * SB * ELR + SR * ST * ELL) / (SR - SL) / (ST - SB)
         - ST * SB / (ST - SB) * (q[IRR][IA] - q[ILL][IA])
         + SR * SL / (SR - SL) * (q[IRR][IB] - q[ILL][IB]);
}

/* riemann_mhd.h:518-609 */
static real_t mag2d_llf(const orc_params *P, real_t q[4][8], const real_t e[4]) {
  real_t E = (e[ILL] + e[IRL] + e[ILR] + e[IRR]) / 4;
  real_t ql[8], qr[8], fx[8], fy[8];
  for (int n = 0; n < 8; ++n) {
    ql[n] = (q[ILL][n] + q[ILR][n]) / 2;
    qr[n] = (q[IRR][n] + q[IRL][n]) / 2;
  }
  riemann_llf_mhd(P, ql, qr, fx, ZERO);
  static const int sw[8] = {ID, IP, IV, IU, IW, IB, IA, IC}; /* :577-600 swap u<->v, a<->b */
  for (int n = 0; n < 8; ++n) {
    ql[n] = (q[ILL][sw[n]] + q[IRL][sw[n]]) / 2;
    qr[n] = (q[IRR][sw[n]] + q[ILR][sw[n]]) / 2;
  }
  riemann_llf_mhd(P, ql, qr, fy, ZERO);
  E += (fx[IB] - fy[IB]);
  return E;
}

/* riemann_mhd.h:615-821  2-D HLLD */
static real_t mag2d_hlld(const orc_params *P, real_t q[4][8], const real_t e[4]) {
  const real_t *qLL = q[ILL], *qRL = q[IRL], *qLR = q[ILR], *qRR = q[IRR];
  real_t ELL = e[ILL], ERL = e[IRL], ELR = e[ILR], ERR = e[IRR];
  real_t rLL = qLL[ID], pLL = qLL[IP], uLL = qLL[IU], vLL = qLL[IV], aLL = qLL[IA], bLL = qLL[IB], cLL = qLL[IC];
  real_t rLR = qLR[ID], pLR = qLR[IP], uLR = qLR[IU], vLR = qLR[IV], aLR = qLR[IA], bLR = qLR[IB], cLR = qLR[IC];
  real_t rRL = qRL[ID], pRL = qRL[IP], uRL = qRL[IU], vRL = qRL[IV], aRL = qRL[IA], bRL = qRL[IB], cRL = qRL[IC];
  real_t rRR = qRR[ID], pRR = qRR[IP], uRR = qRR[IU], vRR = qRR[IV], aRR = qRR[IA], bRR = qRR[IB], cRR = qRR[IC];

  real_t cFastLLx = fast_speed(P, qLL, 0), cFastLRx = fast_speed(P, qLR, 0);
  real_t cFastRLx = fast_speed(P, qRL, 0), cFastRRx = fast_speed(P, qRR, 0);
  real_t cFastLLy = fast_speed(P, qLL, 1), cFastLRy = fast_speed(P, qLR, 1);
  real_t cFastRLy = fast_speed(P, qRL, 1), cFastRRy = fast_speed(P, qRR, 1);

  real_t SL = min4(uLL, uLR, uRL, uRR) - max4(cFastLLx, cFastLRx, cFastRLx, cFastRRx);
  real_t SR = max4(uLL, uLR, uRL, uRR) + max4(cFastLLx, cFastLRx, cFastRLx, cFastRRx);
  real_t SB = min4(vLL, vLR, vRL, vRR) - max4(cFastLLy, cFastLRy, cFastRLy, cFastRRy);
  real_t ST = max4(vLL, vLR, vRL, vRR) + max4(cFastLLy, cFastLRy, cFastRLy, cFastRRy);

  real_t PtotLL = pLL + HALF * (aLL * aLL + bLL * bLL + cLL * cLL);
  real_t PtotLR = pLR + HALF * (aLR * aLR + bLR * bLR + cLR * cLR);
  real_t PtotRL = pRL + HALF * (aRL * aRL + bRL * bRL + cRL * cRL);
  real_t PtotRR = pRR + HALF * (aRR * aRR + bRR * bRR + cRR * cRR);

  real_t rcLLx = rLL * (uLL - SL), rcRLx = rRL * (SR - uRL);
  real_t rcLRx = rLR * (uLR - SL), rcRRx = rRR * (SR - uRR);
  real_t rcLLy = rLL * (vLL - SB), rcLRy = rLR * (ST - vLR);
  real_t rcRLy = rRL * (vRL - SB), rcRRy = rRR * (ST - vRR);

  real_t ustar = (rcLLx * uLL + rcLRx * uLR + rcRLx * uRL + rcRRx * uRR + (PtotLL - PtotRL + PtotLR - PtotRR)) / (rcLLx + rcLRx + rcRLx + rcRRx);
  real_t vstar = (rcLLy * vLL + rcLRy * vLR + rcRLy * vRL + rcRRy * vRR + (PtotLL - PtotLR + PtotRL - PtotRR)) / (rcLLy + rcLRy + rcRLy + rcRRy);

  real_t rstarLLx = rLL * (SL - uLL) / (SL - ustar);
  real_t BstarLL = bLL * (SL - uLL) / (SL - ustar);
  real_t rstarLLy = rLL * (SB - vLL) / (SB - vstar);
  real_t AstarLL = aLL * (SB - vLL) / (SB - vstar);
  real_t rstarLL = rLL * (SL - uLL) / (SL - ustar) * (SB - vLL) / (SB - vstar);
  real_t EstarLLx = ustar * BstarLL - vLL * aLL;
  real_t EstarLLy = uLL * bLL - vstar * AstarLL;
  real_t EstarLL = ustar * BstarLL - vstar * AstarLL;

  real_t rstarLRx = rLR * (SL - uLR) / (SL - ustar);
  real_t BstarLR = bLR * (SL - uLR) / (SL - ustar);
  real_t rstarLRy = rLR * (ST - vLR) / (ST - vstar);
  real_t AstarLR = aLR * (ST - vLR) / (ST - vstar);
  real_t rstarLR = rLR * (SL - uLR) / (SL - ustar) * (ST - vLR) / (ST - vstar);
  real_t EstarLRx = ustar * BstarLR - vLR * aLR;
  real_t EstarLRy = uLR * bLR - vstar * AstarLR;
  real_t EstarLR = ustar * BstarLR - vstar * AstarLR;

  real_t rstarRLx = rRL * (SR - uRL) / (SR - ustar);
  real_t BstarRL = bRL * (SR - uRL) / (SR - ustar);
  real_t rstarRLy = rRL * (SB - vRL) / (SB - vstar);
  real_t AstarRL = aRL * (SB - vRL) / (SB - vstar);
  real_t rstarRL = rRL * (SR - uRL) / (SR - ustar) * (SB - vRL) / (SB - vstar);
  real_t EstarRLx = ustar * BstarRL - vRL * aRL;
  real_t EstarRLy = uRL * bRL - vstar * AstarRL;
  real_t EstarRL = ustar * BstarRL - vstar * AstarRL;

  real_t rstarRRx = rRR * (SR - uRR) / (SR - ustar);
  real_t BstarRR = bRR * (SR - uRR) / (SR - ustar);
  real_t rstarRRy = rRR * (ST - vRR) / (ST - vstar);
  real_t AstarRR = aRR * (ST - vRR) / (ST - vstar);
  real_t rstarRR = rRR * (SR - uRR) / (SR - ustar) * (ST - vRR) / (ST - vstar);
  real_t EstarRRx = ustar * BstarRR - vRR * aRR;
  real_t EstarRRy = uRR * bRR - vstar * AstarRR;
  real_t EstarRR = ustar * BstarRR - vstar * AstarRR;

  real_t calfvenL = max5(FABS_(aLR) / SQRT_(rstarLRx), FABS_(AstarLR) / SQRT_(rstarLR),
                         FABS_(aLL) / SQRT_(rstarLLx), FABS_(AstarLL) / SQRT_(rstarLL), P->smallc);
  real_t calfvenR = max5(FABS_(aRR) / SQRT_(rstarRRx), FABS_(AstarRR) / SQRT_(rstarRR),
                         FABS_(aRL) / SQRT_(rstarRLx), FABS_(AstarRL) / SQRT_(rstarRL), P->smallc);
  real_t calfvenB = max5(FABS_(bLL) / SQRT_(rstarLLy), FABS_(BstarLL) / SQRT_(rstarLL),
                         FABS_(bRL) / SQRT_(rstarRLy), FABS_(BstarRL) / SQRT_(rstarRL), P->smallc);
  real_t calfvenT = max5(FABS_(bLR) / SQRT_(rstarLRy), FABS_(BstarLR) / SQRT_(rstarLR),
                         FABS_(bRR) / SQRT_(rstarRRy), FABS_(BstarRR) / SQRT_(rstarRR), P->smallc);

  real_t SAL = FMIN_(ustar - calfvenL, ZERO);
  real_t SAR = FMAX_(ustar + calfvenR, ZERO);
  real_t SAB = FMIN_(vstar - calfvenB, ZERO);
  real_t SAT = FMAX_(vstar + calfvenT, ZERO);

  real_t AstarT = (SAR * AstarRR - SAL * AstarLR) / (SAR - SAL);
  real_t AstarB = (SAR * AstarRL - SAL * AstarLL) / (SAR - SAL);
  real_t BstarR = (SAT * BstarRR - SAB * BstarRL) / (SAT - SAB);
  real_t BstarL = (SAT * BstarLR - SAB * BstarLL) / (SAT - SAB);

  real_t E = 0, tmpE = 0;
  /* integer masks, :759-762 (copysign: -0.0 counts as negative) */
  int SB_pos = (int)(1 + COPYSIGN_(ONE, SB)) / 2, SB_neg = 1 - SB_pos;
  int ST_pos = (int)(1 + COPYSIGN_(ONE, ST)) / 2, ST_neg = 1 - ST_pos;
  int SL_pos = (int)(1 + COPYSIGN_(ONE, SL)) / 2, SL_neg = 1 - SL_pos;
  int SR_pos = (int)(1 + COPYSIGN_(ONE, SR)) / 2, SR_neg = 1 - SR_pos;

  tmpE = (SAL * SAB * EstarRR - SAL * SAT * EstarRL - SAR * SAB * EstarLR + SAR * SAT * EstarLL) / (SAR - SAL) / (SAT - SAB)
         - SAT * SAB / (SAT - SAB) * (AstarT - AstarB) + SAR * SAL / (SAR - SAL) * (BstarR - BstarL);
  E += (SB_neg * ST_pos * SL_neg * SR_pos) * tmpE;

  tmpE = (SAR * EstarLLx - SAL * EstarRLx + SAR * SAL * (bRL - bLL)) / (SAR - SAL);
  tmpE = SL_pos * ELL + SL_neg * SR_neg * ERL + SL_neg * SR_pos * tmpE;
  E += SB_pos * tmpE;

  tmpE = (SAR * EstarLRx - SAL * EstarRRx + SAR * SAL * (bRR - bLR)) / (SAR - SAL);
  tmpE = SL_pos * ELR + SL_neg * SR_neg * ERR + SL_neg * SR_pos * tmpE;
  E += (SB_neg * ST_neg) * tmpE;

  tmpE = (SAT * EstarLLy - SAB * EstarLRy - SAT * SAB * (aLR - aLL)) / (SAT - SAB);
  E += (SB_neg * ST_pos * SL_pos) * tmpE;

  tmpE = (SAT * EstarRLy - SAB * EstarRRy - SAT * SAB * (aRR - aRL)) / (SAT - SAB);
  E += (SB_neg * ST_pos * SL_neg * SR_neg) * tmpE;
  return E;
}

/* riemann_mhd.h:1054-1193  compute_emf<emfDir>; emfDir: 0 = EMFX, 1 = EMFY, 2 = EMFZ */
real_t orc_compute_emf(const orc_params *P, int emfDir, const real_t qEdge[4][8], real_t xPos) {
  const real_t *qRT = qEdge[IRT], *qLT = qEdge[ILT], *qRB = qEdge[IRB], *qLB = qEdge[ILB];
  real_t q[4][8];
  q[ILL][ID] = qRT[ID]; q[IRL][ID] = qLT[ID]; q[ILR][ID] = qRB[ID]; q[IRR][ID] = qLB[ID];
  if (P->cIso > 0) {
    for (int s = 0; s < 4; ++s) q[s][IP] = q[s][ID] * P->cIso * P->cIso;
  } else {
    q[ILL][IP] = qRT[IP]; q[IRL][IP] = qLT[IP]; q[ILR][IP] = qRB[IP]; q[IRR][IP] = qLB[IP];
  }
  int iu, iv, iw, ia, ib, ic;
  if (emfDir == 2)      { iu = IU; iv = IV; iw = IW; ia = IA; ib = IB; ic = IC; }
  else if (emfDir == 1) { iu = IW; iv = IU; iw = IV; ia = IC; ib = IA; ic = IB; }
  else                  { iu = IV; iv = IW; iw = IU; ia = IB; ib = IC; ic = IA; }
  q[ILL][IU] = qRT[iu]; q[IRL][IU] = qLT[iu]; q[ILR][IU] = qRB[iu]; q[IRR][IU] = qLB[iu];
  q[ILL][IV] = qRT[iv]; q[IRL][IV] = qLT[iv]; q[ILR][IV] = qRB[iv]; q[IRR][IV] = qLB[iv];
  q[ILL][IA] = HALF * (qRT[ia] + qLT[ia]); q[IRL][IA] = HALF * (qRT[ia] + qLT[ia]);
  q[ILR][IA] = HALF * (qRB[ia] + qLB[ia]); q[IRR][IA] = HALF * (qRB[ia] + qLB[ia]);
  q[ILL][IB] = HALF * (qRT[ib] + qRB[ib]); q[IRL][IB] = HALF * (qLT[ib] + qLB[ib]);
  q[ILR][IB] = HALF * (qRT[ib] + qRB[ib]); q[IRR][IB] = HALF * (qLT[ib] + qLB[ib]);
  q[ILL][IW] = qRT[iw]; q[IRL][IW] = qLT[iw]; q[ILR][IW] = qRB[iw]; q[IRR][IW] = qLB[iw];
  q[ILL][IC] = qRT[ic]; q[IRL][IC] = qLT[ic]; q[ILR][IC] = qRB[ic]; q[IRR][IC] = qLB[ic];

  real_t e[4];
  for (int s = 0; s < 4; ++s) e[s] = q[s][IU] * q[s][IB] - q[s][IV] * q[s][IA];

  real_t emf = 0;
  if (P->magRiemannSolver == MAG_HLLD) emf = mag2d_hlld(P, q, e);
  else if (P->magRiemannSolver == MAG_HLLA) emf = mag2d_hll(P, q, e, 1);
  else if (P->magRiemannSolver == MAG_HLLF) emf = mag2d_hll(P, q, e, 0);
  else if (P->magRiemannSolver == MAG_LLF) emf = mag2d_llf(P, q, e);

  if (P->Omega0 > 0) { /* :1171-1189 shearing-box upwind terms */
    if (emfDir == 0) {
      real_t shear = -1.5 * P->Omega0 * xPos;
      if (shear > 0) emf += shear * q[ILL][IB]; else emf += shear * q[IRR][IB];
    }
    if (emfDir == 2) {
      real_t shear = -1.5 * P->Omega0 * (xPos - P->dx / 2);
      if (shear > 0) emf -= shear * q[ILL][IA]; else emf -= shear * q[IRR][IA];
    }
  }
  return emf;
}

/* ------------------------------------------------------------------------------------------
 * slopes.  slope_mhd.h:459-500 (hydro part, slope_type 1|2) and :636-700 (face-B)
 * ---------------------------------------------------------------------------------------- */
static real_t lim_slope(real_t st, real_t qm, real_t q0, real_t qp) {
  real_t dlft = st * (q0 - qm);
  real_t drgt = st * (qp - q0);
  real_t dcen = HALF * (qp - qm);
  real_t dsgn = (dcen >= ZERO) ? ONE : -ONE;
  real_t slop = FMIN_(FABS_(dlft), FABS_(drgt));
  real_t dlim = slop;
  if ((dlft * drgt) <= ZERO) dlim = ZERO;
  return dsgn * FMIN_(dlim, FABS_(dcen));
}

#define AT(arr, i, j, k, v) (arr)[(size_t)(i) + (size_t)isz * ((size_t)(j) + (size_t)jsz * ((size_t)(k) + (size_t)ksz * (size_t)(v)))]

/* slope_mhd.h:352-409  slope_type 3: central differences scaled by one positivity-preserving factor taken over the
 * 27-cell neighbourhood.  nb(di,dj,dk) returns the neighbour value of the variable. */
static void slope27(const real_t *Q, const orc_params *P, int i, int j, int k, int v, real_t *dx_, real_t *dy_, real_t *dz_) {
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize;
  const real_t q0 = AT(Q, i, j, k, v);
  real_t vmin = ZERO, vmax = ZERO; /* the centre difference (0) takes part in both */
  for (int di = -1; di <= 1; ++di)
    for (int dj = -1; dj <= 1; ++dj)
      for (int dk = -1; dk <= 1; ++dk) {
        real_t d = AT(Q, i + di, j + dj, k + dk, v) - q0;
        vmin = FMIN_(vmin, d);
        vmax = FMAX_(vmax, d);
      }
  real_t dfx = HALF * (AT(Q, i + 1, j, k, v) - AT(Q, i - 1, j, k, v));
  real_t dfy = HALF * (AT(Q, i, j + 1, k, v) - AT(Q, i, j - 1, k, v));
  real_t dfz = HALF * (AT(Q, i, j, k + 1, v) - AT(Q, i, j, k - 1, v));
  real_t dff = HALF * (FABS_(dfx) + FABS_(dfy) + FABS_(dfz));
  real_t slop = ONE;
  if (dff > ZERO) slop = FMIN_(ONE, FMIN_(FABS_(vmin), FABS_(vmax)) / dff);
  *dx_ = slop * dfx; *dy_ = slop * dfy; *dz_ = slop * dfz;
}

/* trace_mhd.h:1853-2248  trace_unsplit_mhd_3d_simpler (dq is consumed: halved in place) */
void orc_trace_mhd_3d(const orc_params *P, const real_t q[8], const real_t dq_[3][8],
                      const real_t bfNb[6], const real_t dbf[12], const real_t E[3][2][2],
                      real_t dtdx, real_t dtdy, real_t dtdz, real_t xPos,
                      real_t qm[3][8], real_t qp[3][8], real_t qEdge[4][3][8]) {
  const real_t gamma = P->gamma0, smallR = P->smallr, smallp = P->smallp, Omega0 = P->Omega0, dx = P->dx;
  real_t ELL = E[0][0][0], ELR = E[0][0][1], ERL = E[0][1][0], ERR = E[0][1][1];
  real_t FLL = E[1][0][0], FLR = E[1][0][1], FRL = E[1][1][0], FRR = E[1][1][1];
  real_t GLL = E[2][0][0], GLR = E[2][0][1], GRL = E[2][1][0], GRR = E[2][1][1];

  real_t r = q[ID], p = q[IP], u = q[IU], v = q[IV], w = q[IW], A = q[IA], B = q[IB], C = q[IC];
  real_t AL = bfNb[0], AR = bfNb[1], BL = bfNb[2], BR = bfNb[3], CL = bfNb[4], CR = bfNb[5];

  real_t drx = HALF * dq_[0][ID], dpx = HALF * dq_[0][IP], dux = HALF * dq_[0][IU], dvx = HALF * dq_[0][IV];
  real_t dwx = HALF * dq_[0][IW], dCx = HALF * dq_[0][IC], dBx = HALF * dq_[0][IB];
  real_t dry = HALF * dq_[1][ID], dpy = HALF * dq_[1][IP], duy = HALF * dq_[1][IU], dvy = HALF * dq_[1][IV];
  real_t dwy = HALF * dq_[1][IW], dCy = HALF * dq_[1][IC], dAy = HALF * dq_[1][IA];
  real_t drz = HALF * dq_[2][ID], dpz = HALF * dq_[2][IP], duz = HALF * dq_[2][IU], dvz = HALF * dq_[2][IV];
  real_t dwz = HALF * dq_[2][IW], dAz = HALF * dq_[2][IA], dBz = HALF * dq_[2][IB];

  real_t dALy = HALF * dbf[0], dALz = HALF * dbf[1], dBLx = HALF * dbf[2], dBLz = HALF * dbf[3];
  real_t dCLx = HALF * dbf[4], dCLy = HALF * dbf[5], dARy = HALF * dbf[6], dARz = HALF * dbf[7];
  real_t dBRx = HALF * dbf[8], dBRz = HALF * dbf[9], dCRx = HALF * dbf[10], dCRy = HALF * dbf[11];

  real_t dAx = HALF * (AR - AL), dBy = HALF * (BR - BL), dCz = HALF * (CR - CL);

  /* :1985-1992 */
  real_t sr0 = (-u * drx - dux * r) * dtdx + (-v * dry - dvy * r) * dtdy + (-w * drz - dwz * r) * dtdz;
  real_t su0 = (-u * dux - (dpx + B * dBx + C * dCx) / r) * dtdx + (-v * duy + B * dAy / r) * dtdy + (-w * duz + C * dAz / r) * dtdz;
  real_t sv0 = (-u * dvx + A * dBx / r) * dtdx + (-v * dvy - (dpy + A * dAy + C * dCy) / r) * dtdy + (-w * dvz + C * dBz / r) * dtdz;
  real_t sw0 = (-u * dwx + A * dCx / r) * dtdx + (-v * dwy + B * dCy / r) * dtdy + (-w * dwz - (dpz + A * dAz + B * dBz) / r) * dtdz;
  real_t sp0 = (-u * dpx - dux * gamma * p) * dtdx + (-v * dpy - dvy * gamma * p) * dtdy + (-w * dpz - dwz * gamma * p) * dtdz;
  real_t sA0 = (u * dBy + B * duy - v * dAy - A * dvy) * dtdy + (u * dCz + C * duz - w * dAz - A * dwz) * dtdz;
  real_t sB0 = (v * dAx + A * dvx - u * dBx - B * dux) * dtdx + (v * dCz + C * dvz - w * dBz - B * dwz) * dtdz;
  real_t sC0 = (w * dAx + A * dwx - u * dCx - C * dux) * dtdx + (w * dBy + B * dwy - v * dCy - C * dvy) * dtdy;
  if (Omega0 > 0) { /* :1993-2003 */
    real_t shear = -1.5 * Omega0 * xPos;
    sr0 = sr0 - shear * dry * dtdy;
    su0 = su0 - shear * duy * dtdy;
    sv0 = sv0 - shear * dvy * dtdy;
    sw0 = sw0 - shear * dwy * dtdy;
    sp0 = sp0 - shear * dpy * dtdy;
    sA0 = sA0 - shear * dAy * dtdy;
    sB0 = sB0 + (shear * dAx - 1.5 * Omega0 * A * dx) * dtdx + shear * dBz * dtdz;
    sC0 = sC0 - shear * dCy * dtdy;
  }
  /* :2006-2011 */
  real_t sAL0 = +(GLR - GLL) * dtdy * HALF - (FLR - FLL) * dtdz * HALF;
  real_t sAR0 = +(GRR - GRL) * dtdy * HALF - (FRR - FRL) * dtdz * HALF;
  real_t sBL0 = -(GRL - GLL) * dtdx * HALF + (ELR - ELL) * dtdz * HALF;
  real_t sBR0 = -(GRR - GLR) * dtdx * HALF + (ERR - ERL) * dtdz * HALF;
  real_t sCL0 = +(FRL - FLL) * dtdx * HALF - (ERL - ELL) * dtdy * HALF;
  real_t sCR0 = +(FRR - FLR) * dtdx * HALF - (ERR - ELR) * dtdy * HALF;

  r = r + sr0; u = u + su0; v = v + sv0; w = w + sw0; p = p + sp0; A = A + sA0; B = B + sB0; C = C + sC0;
  AL = AL + sAL0; AR = AR + sAR0; BL = BL + sBL0; BR = BR + sBR0; CL = CL + sCL0; CR = CR + sCR0;

#define FLOOR_(s) do { (s)[ID] = FMAX_(smallR, (s)[ID]); (s)[IP] = FMAX_(smallp, (s)[IP]); } while (0)
#define SET_(s, r_, u_, v_, w_, p_, a_, b_, c_) do { (s)[ID] = (r_); (s)[IU] = (u_); (s)[IV] = (v_); \
    (s)[IW] = (w_); (s)[IP] = (p_); (s)[IA] = (a_); (s)[IB] = (b_); (s)[IC] = (c_); FLOOR_(s); } while (0)
  /* faces :2032-2102 */
  SET_(qp[0], r - drx, u - dux, v - dvx, w - dwx, p - dpx, AL, B - dBx, C - dCx);
  SET_(qm[0], r + drx, u + dux, v + dvx, w + dwx, p + dpx, AR, B + dBx, C + dCx);
  SET_(qp[1], r - dry, u - duy, v - dvy, w - dwy, p - dpy, A - dAy, BL, C - dCy);
  SET_(qm[1], r + dry, u + duy, v + dvy, w + dwy, p + dpy, A + dAy, BR, C + dCy);
  SET_(qp[2], r - drz, u - duz, v - dvz, w - dwz, p - dpz, A - dAz, B - dBz, CL);
  SET_(qm[2], r + drz, u + duz, v + dvz, w + dwz, p + dpz, A + dAz, B + dBz, CR);
  /* X edges :2104-2150 */
  SET_(qEdge[IRT][0], r + (+dry + drz), u + (+duy + duz), v + (+dvy + dvz), w + (+dwy + dwz), p + (+dpy + dpz), A + (+dAy + dAz), BR + (+dBRz), CR + (+dCRy));
  SET_(qEdge[IRB][0], r + (+dry - drz), u + (+duy - duz), v + (+dvy - dvz), w + (+dwy - dwz), p + (+dpy - dpz), A + (+dAy - dAz), BR + (-dBRz), CL + (+dCLy));
  SET_(qEdge[ILT][0], r + (-dry + drz), u + (-duy + duz), v + (-dvy + dvz), w + (-dwy + dwz), p + (-dpy + dpz), A + (-dAy + dAz), BL + (+dBLz), CR + (-dCRy));
  SET_(qEdge[ILB][0], r + (-dry - drz), u + (-duy - duz), v + (-dvy - dvz), w + (-dwy - dwz), p + (-dpy - dpz), A + (-dAy - dAz), BL + (-dBLz), CL + (-dCLy));
  /* Y edges :2152-2198 */
  SET_(qEdge[IRT][1], r + (+drx + drz), u + (+dux + duz), v + (+dvx + dvz), w + (+dwx + dwz), p + (+dpx + dpz), AR + (+dARz), B + (+dBx + dBz), CR + (+dCRx));
  SET_(qEdge[IRB][1], r + (+drx - drz), u + (+dux - duz), v + (+dvx - dvz), w + (+dwx - dwz), p + (+dpx - dpz), AR + (-dARz), B + (+dBx - dBz), CL + (+dCLx));
  SET_(qEdge[ILT][1], r + (-drx + drz), u + (-dux + duz), v + (-dvx + dvz), w + (-dwx + dwz), p + (-dpx + dpz), AL + (+dALz), B + (-dBx + dBz), CR + (-dCRx));
  SET_(qEdge[ILB][1], r + (-drx - drz), u + (-dux - duz), v + (-dvx - dvz), w + (-dwx - dwz), p + (-dpx - dpz), AL + (-dALz), B + (-dBx - dBz), CL + (-dCLx));
  /* Z edges :2200-2246 */
  SET_(qEdge[IRT][2], r + (+drx + dry), u + (+dux + duy), v + (+dvx + dvy), w + (+dwx + dwy), p + (+dpx + dpy), AR + (+dARy), BR + (+dBRx), C + (+dCx + dCy));
  SET_(qEdge[IRB][2], r + (+drx - dry), u + (+dux - duy), v + (+dvx - dvy), w + (+dwx - dwy), p + (+dpx - dpy), AR + (-dARy), BL + (+dBLx), C + (+dCx - dCy));
  SET_(qEdge[ILT][2], r + (-drx + dry), u + (-dux + duy), v + (-dvx + dvy), w + (-dwx + dwy), p + (-dpx + dpy), AL + (+dALy), BR + (-dBRx), C + (-dCx + dCy));
  SET_(qEdge[ILB][2], r + (-drx - dry), u + (-dux - duy), v + (-dvx - dvy), w + (-dwx - dwy), p + (-dpx - dpy), AL + (-dALy), BL + (-dBLx), C + (-dCx - dCy));
#undef SET_
#undef FLOOR_
}

/* ------------------------------------------------------------------------------------------
 * constoprim.h:137-199 constoprim_mhd + :438-463 computePrimitives_MHD_3D / :389-420 (2D)
 * ---------------------------------------------------------------------------------------- */

void orc_constoprim_mhd(const orc_params *P, const real_t u[8], const real_t bn[3], real_t q[8], real_t dt) {
  q[ID] = FMAX_(u[ID], P->smallr);
  q[IU] = u[IU] / q[ID]; q[IV] = u[IV] / q[ID]; q[IW] = u[IW] / q[ID];
  q[IA] = HALF * (u[IA] + bn[0]); q[IB] = HALF * (u[IB] + bn[1]); q[IC] = HALF * (u[IC] + bn[2]);
  real_t eken = HALF * (q[IU] * q[IU] + q[IV] * q[IV] + q[IW] * q[IW]);
  real_t emag = HALF * (q[IA] * q[IA] + q[IB] * q[IB] + q[IC] * q[IC]);
  if (P->cIso > 0) {
    q[IP] = q[ID] * (P->cIso) * (P->cIso);
  } else {
    real_t eint = (u[IP] - emag) / q[ID] - eken;
    q[IP] = FMAX_((P->gamma0 - ONE) * q[ID] * eint, q[ID] * P->smallp);
  }
  if (P->Omega0 > 0) { /* Coriolis predictor :189-195 */
    real_t dvx = 2.0 * P->Omega0 * q[IV];
    real_t dvy = -0.5 * P->Omega0 * q[IU];
    q[IU] += dvx * dt * HALF;
    q[IV] += dvy * dt * HALF;
  }
}

static void prim_at(const orc_params *P, const real_t *U, int i, int j, int k, real_t q[8], real_t dt) {
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize;
  real_t u[8], bn[3];
  for (int v = 0; v < 8; ++v) u[v] = AT(U, i, j, k, v);
  bn[0] = AT(U, i + 1, j, k, IA);
  bn[1] = AT(U, i, j + 1, k, IB);
  bn[2] = (P->dim == 3) ? AT(U, i, j, k + 1, IC) : ZERO;
  orc_constoprim_mhd(P, u, bn, q, dt);
}

/* mhd_utils.h:241-284 find_speed_info<NDIM> summed as in MHDRunBase.cpp:141-250 */
real_t orc_compute_dt_mhd(const orc_params *P, const real_t *U) {
  const int gw = P->ghostWidth;
  real_t invDt = P->smallc / FMIN_(P->dx, P->dy);
  const real_t deltaX = P->xMax - P->xMin;
  int k0 = (P->dim == 3) ? gw : 0, k1 = (P->dim == 3) ? P->ksize - gw : 1;
  for (int k = k0; k < k1; ++k)
    for (int j = gw; j < P->jsize - gw; ++j)
      for (int i = gw; i < P->isize - gw; ++i) {
        real_t q[8];
        prim_at(P, U, i, j, k, q, ZERO);
        real_t d = q[ID], p = q[IP], a = q[IA], b = q[IB], c = q[IC];
        real_t b2 = a * a + b * b + c * c;
        real_t c2 = P->gamma0 * p / d;
        real_t d2 = HALF * (b2 / d + c2);
        real_t vx = SQRT_(d2 + SQRT_(d2 * d2 - c2 * a * a / d)) + FABS_(q[IU]);
        real_t vy = SQRT_(d2 + SQRT_(d2 * d2 - c2 * b * b / d)) + FABS_(q[IV]);
        if (P->dim == 3) {
          real_t vz = SQRT_(d2 + SQRT_(d2 * d2 - c2 * c * c / d)) + FABS_(q[IW]);
          if (P->Omega0 > 0) vy += 1.5 * P->Omega0 * deltaX / 2;
          invDt = FMAX_(invDt, vx / P->dx + vy / P->dy + vz / P->dz);
        } else {
          invDt = FMAX_(invDt, vx / P->dx + vy / P->dy);
        }
      }
  /* the inflow speed of the jet limits the step too: HydroRunBase.cpp:420-422, MHDRunBase.cpp:184-186,228-230 */
  if (P->enableJet) invDt = FMAX_(invDt, (P->ujet + P->cjet) / P->dx);
  return P->cfl / invDt;
}

/* ------------------------------------------------------------------------------------------
 * 3D MHD unsplit step, implementation 3/4 on the CPU: mhd_godunov_unsplit_cpu_v3.cpp:11-715
 * (the part of godunov_unsplit_cpu after boundaries + copy; Omega0 == 0 only)
 * ---------------------------------------------------------------------------------------- */
/* test hook: when set, mhd3d_core copies its 18 trace arrays (qm[3], qp[3], qEdge[4][3]; each [var][k][j][i]) here */
static real_t *g_trace_dump = NULL;

static void mhd3d_core(const orc_params *P, const real_t *Uold, real_t *Unew, real_t dt, real_t totalTime, int rot) {
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize, gw = P->ghostWidth;
  const int nx = P->nx, ny = P->ny;
  const real_t Omega0 = P->Omega0, dx = P->dx, dy = P->dy;
  const int shearBox = rot && P->bc[0] == BC_SHEARINGBOX && P->bc[1] == BC_SHEARINGBOX && Omega0 > 0;
  /* Crank-Nicolson Coriolis coefficients, MHDRunGodunov.cpp:2039-2053 */
  real_t lambda = 0, ratio = 1, alpha1 = 1, alpha2 = 0;
  if (rot) {
    lambda = Omega0 * dt;
    lambda = FOURTH * lambda * lambda;
    ratio = (ONE - lambda) / (ONE + lambda);
    alpha1 = ONE / (ONE + lambda);
    alpha2 = Omega0 * dt / (ONE + lambda);
  }
  /* x-border density flux and emf_y strips [comp][k][j] (h_shear_flux_xmin/xmax) and their remaps */
  real_t *sfmin = calloc((size_t)2 * jsz * ksz, sizeof(real_t)), *sfmax = calloc((size_t)2 * jsz * ksz, sizeof(real_t));
  real_t *rmmin = calloc((size_t)jsz * ksz, sizeof(real_t)), *rmmax = calloc((size_t)jsz * ksz, sizeof(real_t));
#define SF(arr, j, k, c) (arr)[(size_t)(j) + (size_t)jsz * ((size_t)(k) + (size_t)ksz * (size_t)(c))]
  const size_t ncell = (size_t)isz * jsz * ksz;
  const real_t dtdx = dt / P->dx, dtdy = dt / P->dy, dtdz = dt / P->dz;

  real_t *Q = calloc(ncell * 8, sizeof(real_t));
  real_t *elec = calloc(ncell * 3, sizeof(real_t));
  real_t *dA = calloc(ncell * 3, sizeof(real_t)), *dB = calloc(ncell * 3, sizeof(real_t)), *dC = calloc(ncell * 3, sizeof(real_t));
  real_t *emf = calloc(ncell * 3, sizeof(real_t));
  /* 18 trace arrays: qm[3], qp[3], qEdge[4][3] */
  real_t *tr = calloc(ncell * 8 * 18, sizeof(real_t));
#define TR(s) (tr + (size_t)(s) * ncell * 8)
  real_t *qm_[3] = {TR(0), TR(1), TR(2)}, *qp_[3] = {TR(3), TR(4), TR(5)};
  real_t *qe_[4][3];
  for (int e = 0; e < 4; ++e) for (int d = 0; d < 3; ++d) qe_[e][d] = TR(6 + e * 3 + d);

  /* convertToPrimitives: MHDRunGodunov.cpp:538-560, loop 0..size-2 */
  for (int k = 0; k < ksz - 1; ++k)
    for (int j = 0; j < jsz - 1; ++j)
      for (int i = 0; i < isz - 1; ++i) {
        real_t q[8];
        prim_at(P, Uold, i, j, k, q, dt);
        for (int v = 0; v < 8; ++v) AT(Q, i, j, k, v) = q[v];
      }

  /* electric field: cpu_v3.cpp:36-101 */
  for (int k = 1; k < ksz - 1; ++k)
    for (int j = 1; j < jsz - 1; ++j)
      for (int i = 1; i < isz - 1; ++i) {
        real_t u, v, w, A, B, C;
        v = FOURTH * (AT(Q, i, j - 1, k - 1, IV) + AT(Q, i, j - 1, k, IV) + AT(Q, i, j, k - 1, IV) + AT(Q, i, j, k, IV));
        w = FOURTH * (AT(Q, i, j - 1, k - 1, IW) + AT(Q, i, j - 1, k, IW) + AT(Q, i, j, k - 1, IW) + AT(Q, i, j, k, IW));
        B = HALF * (AT(Uold, i, j, k - 1, IB) + AT(Uold, i, j, k, IB));
        C = HALF * (AT(Uold, i, j - 1, k, IC) + AT(Uold, i, j, k, IC));
        AT(elec, i, j, k, 0) = v * C - w * B;
        if (rot) { /* MHDRunGodunov.cpp:2474-2478 */
          real_t xPos = P->xMin + dx / 2 + (i - gw) * dx;
          real_t shear = -1.5 * Omega0 * xPos;
          AT(elec, i, j, k, 0) += shear * C;
        }
        u = FOURTH * (AT(Q, i - 1, j, k - 1, IU) + AT(Q, i - 1, j, k, IU) + AT(Q, i, j, k - 1, IU) + AT(Q, i, j, k, IU));
        w = FOURTH * (AT(Q, i - 1, j, k - 1, IW) + AT(Q, i - 1, j, k, IW) + AT(Q, i, j, k - 1, IW) + AT(Q, i, j, k, IW));
        A = HALF * (AT(Uold, i, j, k - 1, IA) + AT(Uold, i, j, k, IA));
        C = HALF * (AT(Uold, i - 1, j, k, IC) + AT(Uold, i, j, k, IC));
        AT(elec, i, j, k, 1) = w * A - u * C;
        u = FOURTH * (AT(Q, i - 1, j - 1, k, IU) + AT(Q, i - 1, j, k, IU) + AT(Q, i, j - 1, k, IU) + AT(Q, i, j, k, IU));
        v = FOURTH * (AT(Q, i - 1, j - 1, k, IV) + AT(Q, i - 1, j, k, IV) + AT(Q, i, j - 1, k, IV) + AT(Q, i, j, k, IV));
        A = HALF * (AT(Uold, i, j - 1, k, IA) + AT(Uold, i, j, k, IA));
        B = HALF * (AT(Uold, i - 1, j, k, IB) + AT(Uold, i, j, k, IB));
        AT(elec, i, j, k, 2) = u * B - v * A;
        if (rot) { /* :2517-2521 */
          real_t xPos = P->xMin + dx / 2 + (i - gw) * dx;
          real_t shear = -1.5 * Omega0 * (xPos - dx / 2);
          AT(elec, i, j, k, 2) -= shear * A;
        }
      }

  /* magnetic slopes: cpu_v3.cpp:115-163 + slope_mhd.h:636-700 */
  {
    const real_t xst = FMIN_(P->slope_type, R(2.0));
    for (int k = 1; k < ksz - 1; ++k)
      for (int j = 1; j < jsz - 1; ++j)
        for (int i = 1; i < isz - 1; ++i) {
          real_t a0 = AT(Uold, i, j, k, IA), b0 = AT(Uold, i, j, k, IB), c0 = AT(Uold, i, j, k, IC);
          AT(dA, i, j, k, 0) = ZERO;
          AT(dA, i, j, k, 1) = lim_slope(xst, AT(Uold, i, j - 1, k, IA), a0, AT(Uold, i, j + 1, k, IA));
          AT(dA, i, j, k, 2) = lim_slope(xst, AT(Uold, i, j, k - 1, IA), a0, AT(Uold, i, j, k + 1, IA));
          AT(dB, i, j, k, 0) = lim_slope(xst, AT(Uold, i - 1, j, k, IB), b0, AT(Uold, i + 1, j, k, IB));
          AT(dB, i, j, k, 1) = ZERO;
          AT(dB, i, j, k, 2) = lim_slope(xst, AT(Uold, i, j, k - 1, IB), b0, AT(Uold, i, j, k + 1, IB));
          AT(dC, i, j, k, 0) = lim_slope(xst, AT(Uold, i - 1, j, k, IC), c0, AT(Uold, i + 1, j, k, IC));
          AT(dC, i, j, k, 1) = lim_slope(xst, AT(Uold, i, j - 1, k, IC), c0, AT(Uold, i, j + 1, k, IC));
          AT(dC, i, j, k, 2) = ZERO;
        }
  }

  /* trace: cpu_v3.cpp:172-361 (slope_type 0/1/2 branch) */
  for (int k = gw - 2; k < ksz - gw + 1; ++k)
    for (int j = gw - 2; j < jsz - gw + 1; ++j)
      for (int i = gw - 2; i < isz - gw + 1; ++i) {
        real_t q[8], dq[3][8], bfNb[6], dbf[12], E[3][2][2], qm[3][8], qp[3][8], qEdge[4][3][8];
        real_t xPos = P->xMin + P->dx / 2 + (i - gw) * P->dx;
        for (int v = 0; v < 8; ++v) {
          q[v] = AT(Q, i, j, k, v);
          if (P->slope_type == 0) {
            dq[0][v] = dq[1][v] = dq[2][v] = ZERO;
          } else if (P->slope_type == 3 && !rot) { /* cpu_v3.cpp:197-210; the rotating CPU step of the reference
                                                      never fills dq for this type (MHDRunGodunov.cpp:2620-2636) */
            slope27(Q, P, i, j, k, v, &dq[0][v], &dq[1][v], &dq[2][v]);
          } else { /* slope_mhd.h:459-500 */
            dq[0][v] = lim_slope(P->slope_type, AT(Q, i - 1, j, k, v), q[v], AT(Q, i + 1, j, k, v));
            dq[1][v] = lim_slope(P->slope_type, AT(Q, i, j - 1, k, v), q[v], AT(Q, i, j + 1, k, v));
            dq[2][v] = lim_slope(P->slope_type, AT(Q, i, j, k - 1, v), q[v], AT(Q, i, j, k + 1, v));
          }
        }
        bfNb[0] = AT(Uold, i, j, k, IA); bfNb[1] = AT(Uold, i + 1, j, k, IA);
        bfNb[2] = AT(Uold, i, j, k, IB); bfNb[3] = AT(Uold, i, j + 1, k, IB);
        bfNb[4] = AT(Uold, i, j, k, IC); bfNb[5] = AT(Uold, i, j, k + 1, IC);
        dbf[0] = AT(dA, i, j, k, 1); dbf[1] = AT(dA, i, j, k, 2);
        dbf[2] = AT(dB, i, j, k, 0); dbf[3] = AT(dB, i, j, k, 2);
        dbf[4] = AT(dC, i, j, k, 0); dbf[5] = AT(dC, i, j, k, 1);
        dbf[6] = AT(dA, i + 1, j, k, 1); dbf[7] = AT(dA, i + 1, j, k, 2);
        dbf[8] = AT(dB, i, j + 1, k, 0); dbf[9] = AT(dB, i, j + 1, k, 2);
        dbf[10] = AT(dC, i, j, k + 1, 0); dbf[11] = AT(dC, i, j, k + 1, 1);
        E[0][0][0] = AT(elec, i, j, k, 0); E[0][0][1] = AT(elec, i, j, k + 1, 0);
        E[0][1][0] = AT(elec, i, j + 1, k, 0); E[0][1][1] = AT(elec, i, j + 1, k + 1, 0);
        E[1][0][0] = AT(elec, i, j, k, 1); E[1][0][1] = AT(elec, i, j, k + 1, 1);
        E[1][1][0] = AT(elec, i + 1, j, k, 1); E[1][1][1] = AT(elec, i + 1, j, k + 1, 1);
        E[2][0][0] = AT(elec, i, j, k, 2); E[2][0][1] = AT(elec, i, j + 1, k, 2);
        E[2][1][0] = AT(elec, i + 1, j, k, 2); E[2][1][1] = AT(elec, i + 1, j + 1, k, 2);
        orc_trace_mhd_3d(P, q, (const real_t(*)[8])dq, bfNb, dbf, (const real_t(*)[2][2])E, dtdx, dtdy, dtdz, xPos, qm, qp, qEdge);
        if (P->gravityEnabled) { /* gravity predictor on every traced velocity, cpu_v3.cpp:277-332 */
          real_t gf[3];
          orc_gravity_at(P, k, gf);
          const real_t g[3] = {HALF * dt * gf[0], HALF * dt * gf[1], HALF * dt * gf[2]};
          for (int d = 0; d < 3; ++d)
            for (int c = 0; c < 3; ++c) {
              qm[d][IU + c] += g[c];
              qp[d][IU + c] += g[c];
              for (int e = 0; e < 4; ++e) qEdge[e][d][IU + c] += g[c];
            }
        }
        for (int v = 0; v < 8; ++v) {
          for (int d = 0; d < 3; ++d) {
            AT(qm_[d], i, j, k, v) = qm[d][v];
            AT(qp_[d], i, j, k, v) = qp[d][v];
            for (int e = 0; e < 4; ++e) AT(qe_[e][d], i, j, k, v) = qEdge[e][d][v];
          }
        }
      }

  if (g_trace_dump) memcpy(g_trace_dump, tr, ncell * 8 * 18 * sizeof(real_t));

  /* fluxes + emf + hydro update: cpu_v3.cpp:372-583 ; rotating frame: MHDRunGodunov.cpp:2786-3100 */
  for (int k = gw; k < ksz - gw + 1; ++k)
    for (int j = gw; j < jsz - gw + 1; ++j)
      for (int i = gw; i < isz - gw + 1; ++i) {
        real_t ql[8], qr[8], fx[8] = {0}, fy[8] = {0}, fz[8] = {0};
        real_t xPos = P->xMin + dx / 2 + (i - gw) * dx;
        for (int v = 0; v < 8; ++v) { ql[v] = AT(qm_[0], i - 1, j, k, v); qr[v] = AT(qp_[0], i, j, k, v); }
        riemann_mhd_(P, ql, qr, fx);
        static const int swy[8] = {ID, IP, IV, IU, IW, IB, IA, IC};
        for (int v = 0; v < 8; ++v) { ql[v] = AT(qm_[1], i, j - 1, k, swy[v]); qr[v] = AT(qp_[1], i, j, k, swy[v]); }
        riemann_mhd_(P, ql, qr, fy);
        if (rot) { /* shear correction of the y flux, :2860-2899 (ql/qr as modified by the solver) */
          real_t shear_y = -1.5 * Omega0 * xPos;
          real_t bn_mean = HALF * (ql[IA] + qr[IA]);
          const real_t *s_ = (shear_y > 0) ? ql : qr;
          real_t eMag = HALF * (s_[IA] * s_[IA] + s_[IB] * s_[IB] + s_[IC] * s_[IC]);
          real_t eKin = HALF * (s_[IU] * s_[IU] + s_[IV] * s_[IV] + s_[IW] * s_[IW]);
          real_t eTot = eKin + eMag + s_[IP] / (P->gamma0 - ONE);
          fy[ID] = fy[ID] + shear_y * s_[ID];
          fy[IP] = fy[IP] + shear_y * (eTot + eMag - bn_mean * bn_mean);
          fy[IU] = fy[IU] + shear_y * s_[ID] * s_[IU];
          fy[IV] = fy[IV] + shear_y * s_[ID] * s_[IV];
          fy[IW] = fy[IW] + shear_y * s_[ID] * s_[IW];
        }
        static const int swz[8] = {ID, IP, IW, IV, IU, IC, IB, IA};
        for (int v = 0; v < 8; ++v) { ql[v] = AT(qm_[2], i, j, k - 1, swz[v]); qr[v] = AT(qp_[2], i, j, k, swz[v]); }
        riemann_mhd_(P, ql, qr, fz);

        const int in_j = j < jsz - gw, in_k = k < ksz - gw, in_i = i < isz - gw;
        if (rot && in_i && in_j && in_k) { /* :2966-2973 */
          real_t dsx = R(2.0) * Omega0 * dt * AT(Unew, i, j, k, IV) / (ONE + lambda);
          real_t dsy = -HALF * Omega0 * dt * AT(Unew, i, j, k, IU) / (ONE + lambda);
          AT(Unew, i, j, k, IU) = AT(Unew, i, j, k, IU) * ratio + dsx;
          AT(Unew, i, j, k, IV) = AT(Unew, i, j, k, IV) * ratio + dsy;
        }
        if (i > gw && in_j && in_k) {
          if (shearBox && i == nx + gw) SF(sfmax, j, k, 0) = fx[ID] * dtdx;
          else AT(Unew, i - 1, j, k, ID) -= fx[ID] * dtdx;
          AT(Unew, i - 1, j, k, IP) -= fx[IP] * dtdx;
          AT(Unew, i - 1, j, k, IU) -= (alpha1 * fx[IU] + alpha2 * fx[IV]) * dtdx;
          AT(Unew, i - 1, j, k, IV) -= (alpha1 * fx[IV] - 0.25 * alpha2 * fx[IU]) * dtdx;
          AT(Unew, i - 1, j, k, IW) -= fx[IW] * dtdx;
        }
        if (in_i && in_j && in_k) {
          if (shearBox && i == gw) SF(sfmin, j, k, 0) = fx[ID] * dtdx;
          else AT(Unew, i, j, k, ID) += fx[ID] * dtdx;
          AT(Unew, i, j, k, IP) += fx[IP] * dtdx;
          AT(Unew, i, j, k, IU) += (alpha1 * fx[IU] + alpha2 * fx[IV]) * dtdx;
          AT(Unew, i, j, k, IV) += (alpha1 * fx[IV] - 0.25 * alpha2 * fx[IU]) * dtdx;
          AT(Unew, i, j, k, IW) += fx[IW] * dtdx;
        }
        if (in_i && j > gw && in_k) {
          AT(Unew, i, j - 1, k, ID) -= fy[ID] * dtdy; AT(Unew, i, j - 1, k, IP) -= fy[IP] * dtdy;
          AT(Unew, i, j - 1, k, IU) -= (alpha1 * fy[IV] + alpha2 * fy[IU]) * dtdy;
          AT(Unew, i, j - 1, k, IV) -= (alpha1 * fy[IU] - 0.25 * alpha2 * fy[IV]) * dtdy;
          AT(Unew, i, j - 1, k, IW) -= fy[IW] * dtdy;
        }
        if (in_i && in_j && in_k) {
          AT(Unew, i, j, k, ID) += fy[ID] * dtdy; AT(Unew, i, j, k, IP) += fy[IP] * dtdy;
          AT(Unew, i, j, k, IU) += (alpha1 * fy[IV] + alpha2 * fy[IU]) * dtdy;
          AT(Unew, i, j, k, IV) += (alpha1 * fy[IU] - 0.25 * alpha2 * fy[IV]) * dtdy;
          AT(Unew, i, j, k, IW) += fy[IW] * dtdy;
        }
        if (in_i && in_j && k > gw) {
          AT(Unew, i, j, k - 1, ID) -= fz[ID] * dtdz; AT(Unew, i, j, k - 1, IP) -= fz[IP] * dtdz;
          AT(Unew, i, j, k - 1, IU) -= (alpha1 * fz[IW] + alpha2 * fz[IV]) * dtdz;
          AT(Unew, i, j, k - 1, IV) -= (alpha1 * fz[IV] - 0.25 * alpha2 * fz[IW]) * dtdz;
          AT(Unew, i, j, k - 1, IW) -= fz[IU] * dtdz;
        }
        if (in_i && in_j && in_k) {
          AT(Unew, i, j, k, ID) += fz[ID] * dtdz; AT(Unew, i, j, k, IP) += fz[IP] * dtdz;
          AT(Unew, i, j, k, IU) += (alpha1 * fz[IW] + alpha2 * fz[IV]) * dtdz;
          AT(Unew, i, j, k, IV) += (alpha1 * fz[IV] - 0.25 * alpha2 * fz[IW]) * dtdz;
          AT(Unew, i, j, k, IW) += fz[IU] * dtdz;
        }

        real_t qe[4][8];
        for (int v = 0; v < 8; ++v) { /* emfZ :550-557 */
          qe[IRT][v] = AT(qe_[IRT][2], i - 1, j - 1, k, v); qe[IRB][v] = AT(qe_[IRB][2], i - 1, j, k, v);
          qe[ILT][v] = AT(qe_[ILT][2], i, j - 1, k, v);     qe[ILB][v] = AT(qe_[ILB][2], i, j, k, v);
        }
        real_t emfZ = orc_compute_emf(P, 2, (const real_t(*)[8])qe, xPos);
        if (!rot || in_k) AT(emf, i, j, k, 0) = emfZ;
        for (int v = 0; v < 8; ++v) { /* emfY :561-569, RB and LT swapped */
          qe[IRT][v] = AT(qe_[IRT][1], i - 1, j, k - 1, v); qe[IRB][v] = AT(qe_[ILT][1], i, j, k - 1, v);
          qe[ILT][v] = AT(qe_[IRB][1], i - 1, j, k, v);     qe[ILB][v] = AT(qe_[ILB][1], i, j, k, v);
        }
        real_t emfY = orc_compute_emf(P, 1, (const real_t(*)[8])qe, xPos);
        if (!rot || in_j) {
          AT(emf, i, j, k, 1) = emfY;
          if (shearBox) { /* :3076-3084 */
            if (i == gw) SF(sfmin, j, k, 1) = emfY;
            if (i == nx + gw) SF(sfmax, j, k, 1) = emfY;
          }
        }
        for (int v = 0; v < 8; ++v) { /* emfX :572-579 */
          qe[IRT][v] = AT(qe_[IRT][0], i, j - 1, k - 1, v); qe[IRB][v] = AT(qe_[IRB][0], i, j - 1, k, v);
          qe[ILT][v] = AT(qe_[ILT][0], i, j, k - 1, v);     qe[ILB][v] = AT(qe_[ILB][0], i, j, k, v);
        }
        real_t emfX = orc_compute_emf(P, 0, (const real_t(*)[8])qe, xPos);
        if (!rot || in_i) AT(emf, i, j, k, 2) = emfX;
      }

  /* gravity source term on the momenta of the inner cells, cpu_v3.cpp:585-588 -> HydroRunBase.cpp:1962-1976;
   * in the rotating step it comes BEFORE the border remap of the density (MHDRunGodunov.cpp:3188-3192 vs :3203) */
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

  if (shearBox) { /* flux / emf remap and border density update, MHDRunGodunov.cpp:3203-3305 */
    real_t deltay = 1.5 * Omega0 * (dx * nx) * (totalTime + dt / 2);
    deltay = FMOD_(deltay, (dy * ny));
    int jplus = (int)(deltay / dy);
    real_t epsi = FMOD_(deltay, dy);
    for (int k = 0; k < ksz; ++k)
      for (int j = 0; j < jsz; ++j) {
        int jremap = j - jplus - 1, jremapp1 = jremap + 1;
        real_t eps = 1.0 - epsi / dy;
        if (jremap < gw) jremap += ny;
        if (jremapp1 < gw) jremapp1 += ny;
        if (j >= gw && j < jsz - gw + 1 && k >= gw && k < ksz - gw + 1) {
          rmmin[j + (size_t)jsz * k] = SF(sfmin, j, k, 0) + (1.0 - eps) * SF(sfmax, jremap, k, 0) + eps * SF(sfmax, jremapp1, k, 0);
          rmmin[j + (size_t)jsz * k] *= HALF;
        }
        AT(emf, gw, j, k, 1) += (1.0 - eps) * SF(sfmax, jremap, k, 1) + eps * SF(sfmax, jremapp1, k, 1);
        AT(emf, gw, j, k, 1) *= HALF;
        jremap = j + jplus; jremapp1 = jremap + 1;
        eps = epsi / dy;
        if (jremap > ny + gw - 1) jremap -= ny;
        if (jremapp1 > ny + gw - 1) jremapp1 -= ny;
        if (j >= gw && j < jsz - gw + 1 && k >= gw && k < ksz - gw + 1) {
          rmmax[j + (size_t)jsz * k] = SF(sfmax, j, k, 0) + (1.0 - eps) * SF(sfmin, jremap, k, 0) + eps * SF(sfmin, jremapp1, k, 0);
          rmmax[j + (size_t)jsz * k] *= HALF;
        }
        AT(emf, nx + gw, j, k, 1) += (1.0 - eps) * SF(sfmin, jremap, k, 1) + eps * SF(sfmin, jremapp1, k, 1);
        AT(emf, nx + gw, j, k, 1) *= HALF;
      }
    for (int k = gw; k < ksz - gw + 1; ++k)
      for (int j = gw; j < jsz - gw + 1; ++j) {
        AT(Unew, gw, j, k, ID) += rmmin[j + (size_t)jsz * k];
        AT(Unew, nx + gw - 1, j, k, ID) -= rmmax[j + (size_t)jsz * k];
        AT(Unew, gw, j, k, ID) = FMAX_(AT(Unew, gw, j, k, ID), P->smallr);
        AT(Unew, nx + gw - 1, j, k, ID) = FMAX_(AT(Unew, nx + gw - 1, j, k, ID), P->smallr);
      }
  }

  /* constrained transport: cpu_v3.cpp:600-630 (emf index: 0 = Z, 1 = Y, 2 = X) */
  for (int k = gw; k < ksz - gw + 1; ++k)
    for (int j = gw; j < jsz - gw + 1; ++j)
      for (int i = gw; i < isz - gw + 1; ++i) {
        if (k < ksz - gw) {
          AT(Unew, i, j, k, IA) += (AT(emf, i, j + 1, k, 0) - AT(emf, i, j, k, 0)) * dtdy;
          AT(Unew, i, j, k, IB) -= (AT(emf, i + 1, j, k, 0) - AT(emf, i, j, k, 0)) * dtdx;
        }
        AT(Unew, i, j, k, IA) -= (AT(emf, i, j, k + 1, 1) - AT(emf, i, j, k, 1)) * dtdz;
        AT(Unew, i, j, k, IB) += (AT(emf, i, j, k + 1, 2) - AT(emf, i, j, k, 2)) * dtdz;
        AT(Unew, i, j, k, IC) += (AT(emf, i + 1, j, k, 1) - AT(emf, i, j, k, 1)) * dtdx;
        AT(Unew, i, j, k, IC) -= (AT(emf, i, j + 1, k, 2) - AT(emf, i, j, k, 2)) * dtdy;
      }
#undef TR
#undef SF
  free(Q); free(elec); free(dA); free(dB); free(dC); free(emf); free(tr);
  free(sfmin); free(sfmax); free(rmmin); free(rmmax);
}

void orc_dissipative_3d(const orc_params *P, real_t *Unew, real_t dt, real_t totalTime, int shear);

/* the trace arrays of one step from Uold (ghosts filled by the caller): out[18][8][ksize][jsize][isize] in the order
 * qm_x, qm_y, qm_z, qp_x, qp_y, qp_z, then qEdge[e][d] at 6 + 3 e + d (e: RT, RB, LT, LB; d: edge direction x, y, z) */
void orc_mhd3d_trace_arrays(const orc_params *P, const real_t *Uold, real_t dt, real_t *out) {
  real_t *tmp = malloc((size_t)orc_array_len(P) * sizeof(real_t));
  memcpy(tmp, Uold, (size_t)orc_array_len(P) * sizeof(real_t));
  g_trace_dump = out;
  mhd3d_core(P, Uold, tmp, dt, ZERO, P->Omega0 > 0);
  g_trace_dump = NULL;
  free(tmp);
}

void orc_mhd3d_step_v3(const orc_params *P, const real_t *Uold, real_t *Unew, real_t dt) {
  mhd3d_core(P, Uold, Unew, dt, ZERO, 0);
  orc_dissipative_3d(P, Unew, dt, ZERO, 0); /* cpu_v3.cpp:661-693 */
}

/* shearing-box ghost remap in x, MHDRunGodunov.cpp:3539-3759 (time-dependent shift deltay) */
static void make_boundaries_shear(const orc_params *P, real_t *U, real_t dt, real_t totalTime) {
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize, gw = P->ghostWidth, nx = P->nx, ny = P->ny, nv = P->nbVar;
  const real_t dy = P->dy, st = P->slope_type;
  real_t deltay = 1.5 * P->Omega0 * (P->dx * nx) * (totalTime + dt);
  deltay = FMOD_(deltay, (dy * ny));
  int jplus = (int)(deltay / dy);
  real_t epsi = FMOD_(deltay, dy);
  /* border copies [var][k][j][g] of the first / last gw inner columns (shearBorderUtils.h:46-97) */
  const size_t nb = (size_t)gw * jsz * ksz * nv;
  real_t *bmin = calloc(nb, sizeof(real_t)), *bmax = calloc(nb, sizeof(real_t));
  real_t *smin = calloc(nb, sizeof(real_t)), *smax = calloc(nb, sizeof(real_t));
#define BD(arr, g, j, k, v) (arr)[(size_t)(g) + (size_t)gw * ((size_t)(j) + (size_t)jsz * ((size_t)(k) + (size_t)ksz * (size_t)(v)))]
  for (int v = 0; v < nv; ++v)
    for (int k = 0; k < ksz; ++k)
      for (int j = 0; j < jsz; ++j)
        for (int g = 0; g < gw; ++g) {
          BD(bmin, g, j, k, v) = AT(U, gw + g, j, k, v);
          BD(bmax, g, j, k, v) = AT(U, isz - 2 * gw + g, j, k, v);
        }
  if (st == 1 || st == 2) { /* :3583-3637 */
    for (int k = 0; k < ksz; ++k)
      for (int j = 1; j < jsz - 1; ++j)
        for (int g = 0; g < gw; ++g)
          for (int v = 0; v < nv; ++v) {
            if (v == IB) {
              BD(smin, g, j, k, IB) = BD(bmin, g, j + 1, k, IB) - BD(bmin, g, j, k, IB);
              BD(smax, g, j, k, IB) = BD(bmax, g, j + 1, k, IB) - BD(bmax, g, j, k, IB);
            } else {
              for (int side = 0; side < 2; ++side) {
                real_t *b = side ? bmax : bmin, *sl = side ? smax : smin;
                real_t dlft = st * (BD(b, g, j, k, v) - BD(b, g, j - 1, k, v));
                real_t drgt = st * (BD(b, g, j + 1, k, v) - BD(b, g, j, k, v));
                real_t dcen = HALF * (dlft + drgt) / st;
                real_t dsgn = (dcen >= ZERO) ? ONE : -ONE;
                real_t slop = FMIN_(FABS_(dlft), FABS_(drgt));
                real_t dlim = slop;
                if ((dlft * drgt) <= ZERO) dlim = ZERO;
                BD(sl, g, j, k, v) = dsgn * FMIN_(dlim, FABS_(dcen));
              }
            }
          }
  }
  for (int k = 0; k < ksz; ++k)
    for (int j = gw; j < jsz - gw; ++j) {
      int jremap = j - jplus - 1, jremapp1 = jremap + 1;
      real_t eps = 1.0 - epsi / dy;
      if (jremap < gw) jremap += ny;
      if (jremapp1 < gw) jremapp1 += ny;
      real_t lam = HALF * eps * (eps - 1.0);
      for (int v = 0; v < nv; ++v)
        for (int g = 0; g < gw; ++g) {
          if (v == IB) AT(U, g, j, k, IB) = BD(bmax, g, jremap, k, IB) + eps * BD(smax, g, jremap, k, IB);
          else AT(U, g, j, k, v) = (1.0 - eps) * BD(bmax, g, jremap, k, v) + eps * BD(bmax, g, jremapp1, k, v) +
                                   lam * (BD(smax, g, jremap, k, v) - BD(smax, g, jremapp1, k, v));
        }
      jremap = j + jplus; jremapp1 = jremap + 1;
      eps = epsi / dy;
      if (jremap > ny + gw - 1) jremap -= ny;
      if (jremapp1 > ny + gw - 1) jremapp1 -= ny;
      lam = HALF * eps * (eps - 1.0);
      for (int v = 0; v < nv; ++v)
        for (int g = 0; g < gw; ++g) {
          const real_t interp = (1.0 - eps) * BD(bmin, g, jremap, k, v) + eps * BD(bmin, g, jremapp1, k, v) +
                                lam * (BD(smin, g, jremapp1, k, v) - BD(smin, g, jremap, k, v));
          if (v < 5) AT(U, nx + gw + g, j, k, v) = interp;
          if (v == IA && g > 0) AT(U, nx + gw + g, j, k, IA) = interp; /* not the first outer ghost face */
          if (v == IB) AT(U, nx + gw + g, j, k, IB) = BD(bmin, g, jremap, k, IB) + eps * BD(smin, g, jremap, k, IB);
          if (v == IC) AT(U, nx + gw + g, j, k, IC) = interp;
        }
    }
#undef BD
  free(bmin); free(bmax); free(smin); free(smax);
}

void orc_make_boundaries(const orc_params *P, real_t *U, int idim);
void orc_make_all_boundaries(const orc_params *P, real_t *U);

/* MHDRunGodunov.cpp:3763-3793: Y, shear-X, Z, Y */
void orc_make_all_boundaries_shear(const orc_params *P, real_t *U, real_t dt, real_t totalTime) {
  orc_make_boundaries(P, U, 2);
  make_boundaries_shear(P, U, dt, totalTime);
  orc_make_boundaries(P, U, 3);
  orc_make_boundaries(P, U, 2);
}

/* MHDRunGodunov.cpp:2031-3440 godunov_unsplit_rotating_cpu (3D): no leading ghost fill, ghosts of
 * UNew are filled at the END of the step */
void orc_mhd3d_rotating_step(const orc_params *P, real_t *Uold, real_t *Unew, real_t dt, real_t totalTime) {
  memcpy(Unew, Uold, (size_t)orc_array_len(P) * sizeof(real_t));
  mhd3d_core(P, Uold, Unew, dt, totalTime, 1);
  const int shearBox = P->bc[0] == BC_SHEARINGBOX && P->bc[1] == BC_SHEARINGBOX && P->Omega0 > 0;
  orc_dissipative_3d(P, Unew, dt, totalTime, shearBox); /* MHDRunGodunov.cpp:3379-3419 */
  if (shearBox)
    orc_make_all_boundaries_shear(P, Unew, dt, totalTime);
  else
    orc_make_all_boundaries(P, Unew);
}
