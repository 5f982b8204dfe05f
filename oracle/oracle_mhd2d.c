/*
 * oracle_mhd2d.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * 2D MHD unsplit step, implementation 1 of the reference CPU path (BASELINE.json configs[0]):
 *   mhd_godunov_unsplit_cpu_v1.cpp:36-243, trace_mhd.h:38-339 (trace_unsplit_mhd_2d),
 *   slope_mhd.h:77-129 and :523-574, constoprim.h:389-420 (2D prim: B_z neighbour = 0).
 * Same floating-point operation order as the reference.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef ORACLE_FLOAT
#define R(x) x##f
#define FMAX_ fmaxf
#define FMIN_ fminf
#define FABS_ fabsf
#else
#define R(x) x
#define FMAX_ fmax
#define FMIN_ fmin
#define FABS_ fabs
#endif
#define HALF R(0.5)
#define ZERO R(0.0)
#define ONE R(1.0)

#define AT2(arr, i, j, v) (arr)[(size_t)(i) + (size_t)isz * ((size_t)(j) + (size_t)jsz * (size_t)(v))]

void orc_constoprim_mhd(const orc_params *P, const real_t u[8], const real_t bn[3], real_t q[8], real_t dt);
void orc_riemann_mhd(const orc_params *P, const real_t ql[8], const real_t qr[8], real_t flux[8]);

static real_t lim2(real_t st, real_t qm, real_t q0, real_t qp) {
  real_t dlft = st * (q0 - qm), drgt = st * (qp - q0);
  real_t dcen = HALF * (qp - qm);
  real_t dsgn = (dcen >= ZERO) ? ONE : -ONE;
  real_t slop = FMIN_(FABS_(dlft), FABS_(drgt));
  real_t dlim = slop;
  if ((dlft * drgt) <= ZERO) dlim = ZERO;
  return dsgn * FMIN_(dlim, FABS_(dcen));
}

void orc_mhd2d_step_v1(const orc_params *P, const real_t *Uold, real_t *Unew, real_t dt) {
  const int isz = P->isize, jsz = P->jsize, gw = P->ghostWidth;
  const size_t ncell = (size_t)isz * jsz;
  const real_t dtdx = dt / P->dx, dtdy = dt / P->dy;
  const real_t smallR = P->smallr, smallp = P->smallp, gamma = P->gamma0, Omega0 = P->Omega0;
  real_t *Q = calloc(ncell * 8, sizeof(real_t));
  real_t *tr = calloc(ncell * 8 * 8, sizeof(real_t)); /* qm_x qp_x qm_y qp_y RT RB LT LB */
  real_t *emf = calloc(ncell, sizeof(real_t));
#define TR(s) (tr + (size_t)(s) * ncell * 8)
  real_t *qm_x = TR(0), *qp_x = TR(1), *qm_y = TR(2), *qp_y = TR(3), *eRT = TR(4), *eRB = TR(5), *eLT = TR(6), *eLB = TR(7);

  /* prim: MHDRunGodunov.cpp:494-516 */
  for (int j = 0; j < jsz - 1; ++j)
    for (int i = 0; i < isz - 1; ++i) {
      real_t u[8], bn[3], q[8];
      for (int v = 0; v < 8; ++v) u[v] = AT2(Uold, i, j, v);
      bn[0] = AT2(Uold, i + 1, j, IA); bn[1] = AT2(Uold, i, j + 1, IB); bn[2] = ZERO;
      orc_constoprim_mhd(P, u, bn, q, dt);
      for (int v = 0; v < 8; ++v) AT2(Q, i, j, v) = q[v];
    }

  /* trace: only the cells whose states are consumed (gw-1 .. size-gw); the reference also runs it on
   * one more ring (cpu_v1.cpp:43-44) reading one element past the row there, results unused */
  for (int j = gw - 1; j <= jsz - gw; ++j)
    for (int i = gw - 1; i <= isz - gw; ++i) {
      real_t xPos = P->xMin + P->dx / 2 + (i - gw) * P->dx;
#define QN(di, dj, v) AT2(Q, i + (di), j + (dj), v)
#define BF(di, dj, c) AT2(Uold, i + (di), j + (dj), IA + (c))
      real_t Ez[2][2];
      for (int di = 0; di < 2; ++di)
        for (int dj = 0; dj < 2; ++dj) {
          real_t u = 0.25f * (QN(di - 1, dj - 1, IU) + QN(di - 1, dj, IU) + QN(di, dj - 1, IU) + QN(di, dj, IU));
          real_t v = 0.25f * (QN(di - 1, dj - 1, IV) + QN(di - 1, dj, IV) + QN(di, dj - 1, IV) + QN(di, dj, IV));
          real_t A = 0.5f * (BF(di, dj - 1, 0) + BF(di, dj, 0));
          real_t B = 0.5f * (BF(di - 1, dj, 1) + BF(di, dj, 1));
          Ez[di][dj] = u * B - v * A;
        }
      real_t ELL = Ez[0][0], ELR = Ez[0][1], ERL = Ez[1][0], ERR = Ez[1][1];
      real_t r = QN(0, 0, ID), p = QN(0, 0, IP), u = QN(0, 0, IU), v = QN(0, 0, IV), w = QN(0, 0, IW);
      real_t A = QN(0, 0, IA), B = QN(0, 0, IB), C = QN(0, 0, IC);
      real_t AL = BF(0, 0, 0), AR = BF(1, 0, 0), BL = BF(0, 0, 1), BR = BF(0, 1, 1);
      real_t dqx[8], dqy[8];
      for (int n = 0; n < 8; ++n) {
        if (P->slope_type == 0) { dqx[n] = dqy[n] = ZERO; }
        else {
          dqx[n] = lim2(P->slope_type, QN(-1, 0, n), QN(0, 0, n), QN(1, 0, n));
          dqy[n] = lim2(P->slope_type, QN(0, -1, n), QN(0, 0, n), QN(0, 1, n));
        }
      }
      real_t drx = HALF * dqx[ID], dpx = HALF * dqx[IP], dux = HALF * dqx[IU], dvx = HALF * dqx[IV];
      real_t dwx = HALF * dqx[IW], dCx = HALF * dqx[IC], dBx = HALF * dqx[IB];
      real_t dry = HALF * dqy[ID], dpy = HALF * dqy[IP], duy = HALF * dqy[IU], dvy = HALF * dqy[IV];
      real_t dwy = HALF * dqy[IW], dCy = HALF * dqy[IC], dAy = HALF * dqy[IA];
      /* face slopes: slope_unsplit_mhd_2d with the 2D (uncapped) slope_type */
      const real_t st = P->slope_type;
      real_t dALy = HALF * lim2(st, BF(0, -1, 0), BF(0, 0, 0), BF(0, 1, 0));
      real_t dBLx = HALF * lim2(st, BF(-1, 0, 1), BF(0, 0, 1), BF(1, 0, 1));
      real_t dARy = HALF * lim2(st, BF(1, -1, 0), BF(1, 0, 0), BF(1, 1, 0));
      real_t dBRx = HALF * lim2(st, BF(-1, 1, 1), BF(0, 1, 1), BF(1, 1, 1));
      real_t dAx = HALF * (AR - AL), dBy = HALF * (BR - BL);
      real_t sr0 = (-u * drx - dux * r) * dtdx + (-v * dry - dvy * r) * dtdy;
      real_t su0 = (-u * dux - dpx / r - B * dBx / r - C * dCx / r) * dtdx + (-v * duy + B * dAy / r) * dtdy;
      real_t sv0 = (-u * dvx + A * dBx / r) * dtdx + (-v * dvy - dpy / r - A * dAy / r - C * dCy / r) * dtdy;
      real_t sw0 = (-u * dwx + A * dCx / r) * dtdx + (-v * dwy + B * dCy / r) * dtdy;
      real_t sp0 = (-u * dpx - dux * gamma * p) * dtdx + (-v * dpy - dvy * gamma * p) * dtdy;
      real_t sA0 = (u * dBy + B * duy - v * dAy - A * dvy) * dtdy;
      real_t sB0 = (-u * dBx - B * dux + v * dAx + A * dvx) * dtdx;
      real_t sC0 = (w * dAx + A * dwx - u * dCx - C * dux) * dtdx + (-v * dCy - C * dvy + w * dBy + B * dwy) * dtdy;
      if (Omega0 > ZERO) {
        real_t shear = -1.5 * Omega0 * xPos;
        sC0 += (shear * dAx - 1.5 * Omega0 * A) * dtdx;
        sC0 += shear * dBy * dtdy;
      }
      real_t sAL0 = +(ELR - ELL) * HALF * dtdy, sAR0 = +(ERR - ERL) * HALF * dtdy;
      real_t sBL0 = -(ERL - ELL) * HALF * dtdx, sBR0 = -(ERR - ELR) * HALF * dtdx;
      r = r + sr0; u = u + su0; v = v + sv0; w = w + sw0; p = p + sp0; A = A + sA0; B = B + sB0; C = C + sC0;
      AL = AL + sAL0; AR = AR + sAR0; BL = BL + sBL0; BR = BR + sBR0;
#define PUT(arr, r_, u_, v_, w_, p_, a_, b_, c_) do { real_t rr_ = FMAX_(smallR, (r_)); \
      AT2(arr, i, j, ID) = rr_; AT2(arr, i, j, IP) = FMAX_(smallp * rr_, (p_)); AT2(arr, i, j, IU) = (u_); \
      AT2(arr, i, j, IV) = (v_); AT2(arr, i, j, IW) = (w_); AT2(arr, i, j, IA) = (a_); AT2(arr, i, j, IB) = (b_); \
      AT2(arr, i, j, IC) = (c_); } while (0)
      PUT(qp_x, r - drx, u - dux, v - dvx, w - dwx, p - dpx, AL, B - dBx, C - dCx);
      PUT(qm_x, r + drx, u + dux, v + dvx, w + dwx, p + dpx, AR, B + dBx, C + dCx);
      PUT(qp_y, r - dry, u - duy, v - dvy, w - dwy, p - dpy, A - dAy, BL, C - dCy);
      PUT(qm_y, r + dry, u + duy, v + dvy, w + dwy, p + dpy, A + dAy, BR, C + dCy);
      PUT(eRT, r + (+drx + dry), u + (+dux + duy), v + (+dvx + dvy), w + (+dwx + dwy), p + (+dpx + dpy), AR + (+dARy), BR + (+dBRx), C + (+dCx + dCy));
      PUT(eRB, r + (+drx - dry), u + (+dux - duy), v + (+dvx - dvy), w + (+dwx - dwy), p + (+dpx - dpy), AR + (-dARy), BL + (+dBLx), C + (+dCx - dCy));
      PUT(eLB, r + (-drx - dry), u + (-dux - duy), v + (-dvx - dvy), w + (-dwx - dwy), p + (-dpx - dpy), AL + (-dALy), BL + (-dBLx), C + (-dCx - dCy));
      PUT(eLT, r + (-drx + dry), u + (-dux + duy), v + (-dvx + dvy), w + (-dwx + dwy), p + (-dpx + dpy), AL + (+dALy), BR + (-dBRx), C + (-dCx + dCy));
#undef PUT
#undef QN
#undef BF
    }

  /* fluxes, update (no write guards in 2D), emf: cpu_v1.cpp:98-222 */
  for (int j = gw; j < jsz - gw + 1; ++j)
    for (int i = gw; i < isz - gw + 1; ++i) {
      real_t ql[8], qr[8], fx[8], fy[8];
      for (int v = 0; v < 8; ++v) { ql[v] = AT2(qm_x, i - 1, j, v); qr[v] = AT2(qp_x, i, j, v); }
      orc_riemann_mhd(P, ql, qr, fx);
      static const int sw[8] = {ID, IP, IV, IU, IW, IB, IA, IC};
      for (int v = 0; v < 8; ++v) { ql[v] = AT2(qm_y, i, j - 1, sw[v]); qr[v] = AT2(qp_y, i, j, sw[v]); }
      orc_riemann_mhd(P, ql, qr, fy);
      static const int upd[6] = {ID, IP, IU, IV, IW, IC};
      for (int n = 0; n < 6; ++n) AT2(Unew, i - 1, j, upd[n]) -= fx[upd[n]] * dtdx;
      for (int n = 0; n < 6; ++n) AT2(Unew, i, j, upd[n]) += fx[upd[n]] * dtdx;
      for (int n = 0; n < 6; ++n) AT2(Unew, i, j - 1, upd[n]) -= fy[sw[upd[n]]] * dtdy;
      for (int n = 0; n < 6; ++n) AT2(Unew, i, j, upd[n]) += fy[sw[upd[n]]] * dtdy;
      real_t qe[4][8];
      for (int v = 0; v < 8; ++v) {
        qe[0][v] = AT2(eRT, i - 1, j - 1, v); qe[1][v] = AT2(eRB, i - 1, j, v);
        qe[2][v] = AT2(eLT, i, j - 1, v);     qe[3][v] = AT2(eLB, i, j, v);
      }
      AT2(emf, i, j, 0) = orc_compute_emf(P, 2, (const real_t(*)[8])qe, ZERO);
    }
  /* CT: cpu_v1.cpp:233-240 */
  for (int j = gw; j < jsz - gw + 1; ++j)
    for (int i = gw; i < isz - gw + 1; ++i) {
      AT2(Unew, i, j, IA) += (AT2(emf, i, j + 1, 0) - AT2(emf, i, j, 0)) * dtdy;
      AT2(Unew, i, j, IB) -= (AT2(emf, i + 1, j, 0) - AT2(emf, i, j, 0)) * dtdx;
    }
#undef TR
  free(Q); free(tr); free(emf);
}
