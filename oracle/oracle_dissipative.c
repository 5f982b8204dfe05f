/*
 * oracle_dissipative.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Dissipative terms of the 3D step, restated from the reference CPU path with the same
 * floating-point operation order: Ohmic resistivity (emf -> constrained transport -> energy flux)
 * and Navier-Stokes viscosity (momentum + energy flux -> conservative update), plus the static
 * gravity source term.  The step drivers (oracle_mhd.c / oracle_hydro.c) call orc_dissipative_3d
 * where the reference does (mhd_godunov_unsplit_cpu_v3.cpp:661-693, MHDRunGodunov.cpp:3379-3419,
 * HydroRunGodunov.cpp:2908-2927).
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>

#define AT(arr, i, j, k, v) (arr)[(size_t)(i) + (size_t)isz * ((size_t)(j) + (size_t)jsz * ((size_t)(k) + (size_t)ksz * (size_t)(v)))]
#define HALF ((real_t)0.5)
#define TWO ((real_t)2.0)

void orc_make_all_boundaries_shear(const orc_params *P, real_t *U, real_t dt, real_t totalTime);

/* velocity component c (IU..IW) of a cell: momentum / density, as written at every use in
 * HydroRunBase.cpp:582-845 */
#define VEL(c, i, j, k) (AT(U, i, j, k, c) / AT(U, i, j, k, ID))

/* MHDRunBase.cpp:526-571  compute_resistivity_emf_3d: emf = -eta * curl(B) at the cell edges;
 * emf component order I_EMFZ = 0, I_EMFY = 1, I_EMFX = 2 (constants.h:191-195) */
void orc_resistivity_emf_3d(const orc_params *P, const real_t *U, real_t *emf) {
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize, gw = P->ghostWidth;
  const real_t dx = P->dx, dy = P->dy, dz = P->dz, eta = P->eta;
  for (int k = gw; k < ksz - gw + 1; ++k)
    for (int j = gw; j < jsz - gw + 1; ++j)
      for (int i = gw; i < isz - gw + 1; ++i) {
        real_t dbydx = (AT(U, i, j, k, IB) - AT(U, i - 1, j, k, IB)) / dx;
        real_t dbzdx = (AT(U, i, j, k, IC) - AT(U, i - 1, j, k, IC)) / dx;
        real_t dbxdy = (AT(U, i, j, k, IA) - AT(U, i, j - 1, k, IA)) / dy;
        real_t dbzdy = (AT(U, i, j, k, IC) - AT(U, i, j - 1, k, IC)) / dy;
        real_t dbxdz = (AT(U, i, j, k, IA) - AT(U, i, j, k - 1, IA)) / dz;
        real_t dbydz = (AT(U, i, j, k, IB) - AT(U, i, j, k - 1, IB)) / dz;
        real_t jx = dbzdy - dbydz, jy = dbxdz - dbzdx, jz = dbydx - dbxdy;
        AT(emf, i, j, k, 2) = -eta * jx;
        AT(emf, i, j, k, 1) = -eta * jy;
        AT(emf, i, j, k, 0) = -eta * jz;
      }
}

/* MHDRunBase.cpp:302-345  compute_ct_update_3d (same un-guarded range as the main step) */
void orc_ct_update_3d(const orc_params *P, real_t *U, const real_t *emf, real_t dt) {
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize, gw = P->ghostWidth;
  const real_t dtdx = dt / P->dx, dtdy = dt / P->dy, dtdz = dt / P->dz;
  for (int k = gw; k < ksz - gw + 1; ++k)
    for (int j = gw; j < jsz - gw + 1; ++j)
      for (int i = gw; i < isz - gw + 1; ++i) {
        if (k < ksz - gw) {
          AT(U, i, j, k, IA) += (AT(emf, i, j + 1, k, 0) - AT(emf, i, j, k, 0)) * dtdy;
          AT(U, i, j, k, IB) -= (AT(emf, i + 1, j, k, 0) - AT(emf, i, j, k, 0)) * dtdx;
        }
        AT(U, i, j, k, IA) -= (AT(emf, i, j, k + 1, 1) - AT(emf, i, j, k, 1)) * dtdz;
        AT(U, i, j, k, IB) += (AT(emf, i, j, k + 1, 2) - AT(emf, i, j, k, 2)) * dtdz;
        AT(U, i, j, k, IC) += (AT(emf, i + 1, j, k, 1) - AT(emf, i, j, k, 1)) * dtdx;
        AT(U, i, j, k, IC) -= (AT(emf, i, j + 1, k, 2) - AT(emf, i, j, k, 2)) * dtdy;
      }
}

/* MHDRunBase.cpp:790-900  compute_resistivity_energy_flux_3d: Poynting flux eta J x B through the
 * low faces; only the energy component is non-zero.  F is [dir][var 0..4][k][j][i]. */
void orc_resistivity_energy_flux_3d(const orc_params *P, const real_t *U, real_t *fx, real_t *fy, real_t *fz, real_t dt) {
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize, gw = P->ghostWidth;
  const real_t dx = P->dx, dy = P->dy, dz = P->dz, eta = P->eta;
#define BX(i, j, k) AT(U, i, j, k, IA)
#define BY(i, j, k) AT(U, i, j, k, IB)
#define BZ(i, j, k) AT(U, i, j, k, IC)
  for (int k = gw; k < ksz - gw + 1; ++k)
    for (int j = gw; j < jsz - gw + 1; ++j)
      for (int i = gw; i < isz - gw + 1; ++i) {
        real_t bx, by, bz, jx, jy, jz, jp;
        /* x face */
        by = (BY(i, j, k) + BY(i - 1, j, k) + BY(i, j + 1, k) + BY(i - 1, j + 1, k)) / 4;
        bz = (BZ(i, j, k) + BZ(i - 1, j, k) + BZ(i, j, k + 1) + BZ(i - 1, j, k + 1)) / 4;
        jy = (BX(i, j, k) - BX(i, j, k - 1)) / dz - (BZ(i, j, k) - BZ(i - 1, j, k)) / dx;
        jp = (BX(i, j, k + 1) - BX(i, j, k)) / dz - (BZ(i, j, k + 1) - BZ(i - 1, j, k + 1)) / dx;
        jy = (jy + jp) / 2;
        jz = (BY(i, j, k) - BY(i - 1, j, k)) / dx - (BX(i, j, k) - BX(i, j - 1, k)) / dy;
        jp = (BY(i, j + 1, k) - BY(i - 1, j + 1, k)) / dx - (BX(i, j + 1, k) - BX(i, j, k)) / dy;
        jz = (jz + jp) / 2;
        for (int v = 0; v < 5; ++v) AT(fx, i, j, k, v) = 0;
        AT(fx, i, j, k, IP) = -eta * (jy * bz - jz * by) * dt / dx;
        /* y face */
        bx = (BX(i, j, k) + BX(i, j - 1, k) + BX(i + 1, j, k) + BX(i + 1, j - 1, k)) / 4;
        bz = (BZ(i, j, k) + BZ(i, j - 1, k) + BZ(i, j, k + 1) + BZ(i, j - 1, k + 1)) / 4;
        jx = (BZ(i, j, k) - BZ(i, j - 1, k)) / dy - (BY(i, j, k) - BY(i, j, k - 1)) / dz;
        jp = (BZ(i, j, k + 1) - BZ(i, j - 1, k + 1)) / dy - (BY(i, j, k + 1) - BY(i, j, k)) / dz;
        jx = (jx + jp) / 2;
        jz = (BY(i, j, k) - BY(i - 1, j, k)) / dx - (BX(i, j, k) - BX(i, j - 1, k)) / dy;
        jp = (BY(i + 1, j, k) - BY(i, j, k)) / dx - (BX(i + 1, j, k) - BX(i + 1, j - 1, k)) / dy;
        jz = (jz + jp) / 2;
        for (int v = 0; v < 5; ++v) AT(fy, i, j, k, v) = 0;
        AT(fy, i, j, k, IP) = -eta * (jz * bx - jx * bz) * dt / dy;
        /* z face */
        bx = (BX(i, j, k) + BX(i, j, k - 1) + BX(i + 1, j, k) + BX(i + 1, j, k - 1)) / 4;
        by = (BY(i, j, k) + BY(i, j, k - 1) + BY(i, j + 1, k) + BY(i, j + 1, k - 1)) / 4;
        jx = (BZ(i, j, k) - BZ(i, j - 1, k)) / dy - (BY(i, j, k) - BY(i, j, k - 1)) / dz;
        jp = (BZ(i, j + 1, k) - BZ(i, j, k)) / dy - (BY(i, j + 1, k) - BY(i, j + 1, k - 1)) / dz;
        jx = (jx + jp) / 2;
        jy = (BX(i, j, k) - BX(i, j, k - 1)) / dz - (BZ(i, j, k) - BZ(i - 1, j, k)) / dx;
        jp = (BX(i + 1, j, k) - BX(i + 1, j, k - 1)) / dz - (BZ(i + 1, j, k) - BZ(i, j, k)) / dx;
        jy = (jy + jp) / 2;
        for (int v = 0; v < 5; ++v) AT(fz, i, j, k, v) = 0;
        AT(fz, i, j, k, IP) = -eta * (jx * by - jy * bx) * dt / dz;
      }
#undef BX
#undef BY
#undef BZ
}

/* HydroRunBase.cpp:582-845  compute_viscosity_flux (3D): viscous stress through the low faces; normal
 * derivatives are two-point, transverse ones the mean of the centred differences of both cells */
void orc_viscosity_flux_3d(const orc_params *P, const real_t *U, real_t *fx, real_t *fy, real_t *fz, real_t dt) {
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize, gw = P->ghostWidth;
  const real_t dx = P->dx, dy = P->dy, dz = P->dz, nu = P->nu, cIso = P->cIso;
  const real_t two3rd = 2. / 3.;
  const real_t dd[3] = {dx, dy, dz};
  real_t *F[3] = {fx, fy, fz};
  memset(fx, 0, sizeof(real_t) * (size_t)isz * jsz * ksz * 5);
  memset(fy, 0, sizeof(real_t) * (size_t)isz * jsz * ksz * 5);
  memset(fz, 0, sizeof(real_t) * (size_t)isz * jsz * ksz * 5);
  for (int k = gw; k < ksz - gw + 1; ++k)
    for (int j = gw; j < jsz - gw + 1; ++j)
      for (int i = gw; i < isz - gw + 1; ++i)
        for (int n = 0; n < 3; ++n) { /* face normal: x, y, z */
          /* (oi,oj,ok): unit offset along the normal; the face lies between cell - o and the cell */
          const int oi = n == 0, oj = n == 1, ok = n == 2;
          real_t rho = HALF * (AT(U, i, j, k, ID) + AT(U, i - oi, j - oj, k - ok, ID));
          real_t vel[3] = {0, 0, 0};
          if (cIso <= 0)
            for (int c = 0; c < 3; ++c) vel[c] = HALF * (VEL(IU + c, i, j, k) + VEL(IU + c, i - oi, j - oj, k - ok));
          /* grad[d][c] = d(v_c)/d(x_d) at the face */
          real_t grad[3][3];
          for (int d = 0; d < 3; ++d) {
            const int ti = d == 0, tj = d == 1, tk = d == 2;
            for (int c = 0; c < 3; ++c) {
              if (d == n) {
                real_t uR = VEL(IU + c, i, j, k), uL = VEL(IU + c, i - oi, j - oj, k - ok);
                grad[d][c] = (uR - uL) / dd[d];
              } else {
                real_t uRR = VEL(IU + c, i + ti, j + tj, k + tk);
                real_t uRL = VEL(IU + c, i + ti - oi, j + tj - oj, k + tk - ok);
                real_t uLR = VEL(IU + c, i - ti, j - tj, k - tk);
                real_t uLL = VEL(IU + c, i - ti - oi, j - tj - oj, k - tk - ok);
                real_t uR = uRR + uRL, uL = uLR + uLL;
                grad[d][c] = (uR - uL) / dd[d] / 4;
              }
            }
          }
          /* stress components through this face: normal one and the two shear ones */
          const int t1 = (n + 1) % 3, t2 = (n + 2) % 3;
          real_t tnn = -two3rd * nu * rho * (TWO * grad[n][n] - grad[n == 0 ? 1 : 0][n == 0 ? 1 : 0] - grad[n == 2 ? 1 : 2][n == 2 ? 1 : 2]);
          real_t tau[3];
          tau[n] = tnn;
          /* shear: -(nu rho)(d_a v_b + d_b v_a), written by the reference with the larger-index derivative first
           * for txy (dudy[IX] + dudx[IY]), txz (dudz[IX] + dudx[IZ]) and tyz (dudz[IY] + dudy[IZ]) */
          for (int s = 0; s < 2; ++s) {
            const int m = s == 0 ? t1 : t2;
            const int lo = n < m ? n : m, hi = n < m ? m : n;
            tau[m] = -nu * rho * (grad[hi][lo] + grad[lo][hi]);
          }
          real_t *f = F[n];
          AT(f, i, j, k, ID) = 0;
          AT(f, i, j, k, IU) = tau[0] * dt / dd[n];
          AT(f, i, j, k, IV) = tau[1] * dt / dd[n];
          AT(f, i, j, k, IW) = tau[2] * dt / dd[n];
          AT(f, i, j, k, IP) = (cIso <= 0) ? (vel[0] * tau[0] + vel[1] * tau[1] + vel[2] * tau[2]) * dt / dd[n] : 0;
        }
}

/* HydroRunBase.cpp:1504-1528 compute_hydro_update / :1675-1697 compute_hydro_update_energy:
 * U(var) += F_lo - F_hi, one direction after the other, inner cells */
void orc_hydro_update_3d(const orc_params *P, real_t *U, const real_t *fx, const real_t *fy, const real_t *fz, int energyOnly) {
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize, gw = P->ghostWidth;
  for (int v = 0; v < 5; ++v) {
    if (energyOnly && v != IP) continue;
    for (int k = gw; k < ksz - gw; ++k)
      for (int j = gw; j < jsz - gw; ++j)
        for (int i = gw; i < isz - gw; ++i) {
          AT(U, i, j, k, v) += (AT(fx, i, j, k, v) - AT(fx, i + 1, j, k, v));
          AT(U, i, j, k, v) += (AT(fy, i, j, k, v) - AT(fy, i, j + 1, k, v));
          AT(U, i, j, k, v) += (AT(fz, i, j, k, v) - AT(fz, i, j, k + 1, v));
        }
  }
}

/* the dissipative block at the end of the 3D step: ghost refresh of UNew, resistivity, viscosity.
 * mhd_godunov_unsplit_cpu_v3.cpp:661-693 ; MHDRunGodunov.cpp:3379-3419 (rotating) ;
 * HydroRunGodunov.cpp:2908-2927 (hydro: viscosity only).
 * Split in two stages for the z-slab protocol test (tests/test_slab_protocol_gloo.py): stage 0 = resistive
 * emf + constrained transport, stage 1 = resistive energy flux + viscosity; the caller fills the ghosts
 * before stage 0.  orc_dissipative_3d = ghost fill + stage 0 + stage 1. */
static int g_skip_dissipative = 0;
void orc_set_skip_dissipative(int on) { g_skip_dissipative = on; }

void orc_dissipative_stage(const orc_params *P, real_t *Unew, real_t dt, int stage) {
  const real_t nu = P->nu, eta = P->mhdEnabled ? P->eta : 0;
  if ((!(nu > 0) && !(eta > 0)) || P->dim != 3) return;
  const size_t ncell = (size_t)P->isize * P->jsize * P->ksize;
  if (stage == 0) {
    if (eta > 0) {
      real_t *emf = calloc(ncell * 3, sizeof(real_t));
      orc_resistivity_emf_3d(P, Unew, emf);
      orc_ct_update_3d(P, Unew, emf, dt);
      free(emf);
    }
    return;
  }
  real_t *fx = calloc(ncell * 5, sizeof(real_t)), *fy = calloc(ncell * 5, sizeof(real_t)), *fz = calloc(ncell * 5, sizeof(real_t));
  if (eta > 0 && P->cIso <= 0) {
    orc_resistivity_energy_flux_3d(P, Unew, fx, fy, fz, dt);
    orc_hydro_update_3d(P, Unew, fx, fy, fz, 1);
  }
  if (nu > 0) {
    orc_viscosity_flux_3d(P, Unew, fx, fy, fz, dt);
    orc_hydro_update_3d(P, Unew, fx, fy, fz, 0);
  }
  free(fx); free(fy); free(fz);
}

void orc_dissipative_3d(const orc_params *P, real_t *Unew, real_t dt, real_t totalTime, int shear) {
  const real_t nu = P->nu, eta = P->mhdEnabled ? P->eta : 0;
  if (g_skip_dissipative) return;
  if (!(nu > 0) && !(eta > 0)) return;
  if (P->dim != 3) return;
  if (shear) orc_make_all_boundaries_shear(P, Unew, dt, totalTime);
  else orc_make_all_boundaries(P, Unew);
  orc_dissipative_stage(P, Unew, dt, 0);
  orc_dissipative_stage(P, Unew, dt, 1);
}

/* ------------------------------------------------------------------------------------------
 * History diagnostics, 3D MHD: MHDRunBase.cpp:3311-3410 (history_default: mass, divB)
 * and :3476-3620 (history_mri: + Maxwell / Reynolds stresses, magnetic pressure, mean field).
 * Accumulation in double in the reference's loop order.  out[8] = mass, maxwell, reynolds, magp,
 * mean_Bx, mean_By, mean_Bz, divB  (history_default prints mass and divB only).
 * ---------------------------------------------------------------------------------------- */
void orc_history_mhd3d(const orc_params *P, const real_t *U, double out[8]) {
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize, gw = P->ghostWidth;
  const real_t dx = P->dx, dy = P->dy, dz = P->dz;
  double mass = 0.0, magp = 0.0, maxwell = 0.0, mbx = 0.0, mby = 0.0, mbz = 0.0;
#define SQR_(x) ((x) * (x))
  for (int k = gw; k < ksz - gw; ++k)
    for (int j = gw; j < jsz - gw; ++j)
      for (int i = gw; i < isz - gw; ++i) {
        mass += AT(U, i, j, k, ID);
        magp += 0.25 * SQR_(AT(U, i, j, k, IA) + AT(U, i + 1, j, k, IA));
        magp += 0.25 * SQR_(AT(U, i, j, k, IB) + AT(U, i, j + 1, k, IB));
        magp += 0.25 * SQR_(AT(U, i, j, k, IC) + AT(U, i, j, k + 1, IC));
        maxwell -= 0.25 * (AT(U, i, j, k, IA) + AT(U, i + 1, j, k, IA)) * (AT(U, i, j, k, IB) + AT(U, i, j + 1, k, IB));
        mbx += AT(U, i, j, k, IA);
        mby += AT(U, i, j, k, IB);
        mbz += AT(U, i, j, k, IC);
      }
  double dTau = dx * dy * dz / (P->xMax - P->xMin) / (P->yMax - P->yMin) / (P->zMax - P->zMin);
  magp = magp * dTau / 2.;
  mass = mass * dTau;
  maxwell = maxwell * dTau;
  mbx *= dTau; mby *= dTau; mbz *= dTau;
  /* y-z averages of rho, u, v per x column (ghost columns included), :3565-3586 */
  real_t *mean = calloc((size_t)isz * 3, sizeof(real_t));
  for (int k = gw; k < ksz - gw; ++k)
    for (int j = gw; j < jsz - gw; ++j)
      for (int i = 0; i < isz; ++i) {
        mean[i] += AT(U, i, j, k, ID);
        mean[isz + i] += AT(U, i, j, k, IU) / AT(U, i, j, k, ID);
        mean[2 * isz + i] += AT(U, i, j, k, IV) / AT(U, i, j, k, ID);
      }
  for (int i = 0; i < isz; ++i) {
    mean[i] /= (P->ny * P->nz);
    mean[isz + i] /= (P->ny * P->nz);
    mean[2 * isz + i] /= (P->ny * P->nz);
  }
  double reynolds = 0.0, divB = 0.0;
  for (int k = gw; k < ksz - gw; ++k)
    for (int j = gw; j < jsz - gw; ++j)
      for (int i = gw; i < isz - gw; ++i) {
        reynolds += AT(U, i, j, k, ID) * dTau * (AT(U, i, j, k, IU) / AT(U, i, j, k, ID) - mean[isz + i]) *
                    (AT(U, i, j, k, IV) / AT(U, i, j, k, ID) - mean[2 * isz + i]);
        divB += (AT(U, i + 1, j, k, IA) - AT(U, i, j, k, IA)) / dx + (AT(U, i, j + 1, k, IB) - AT(U, i, j, k, IB)) / dy +
                (AT(U, i, j, k + 1, IC) - AT(U, i, j, k, IC)) / dz;
      }
  free(mean);
#undef SQR_
  out[0] = mass; out[1] = maxwell; out[2] = reynolds; out[3] = magp;
  out[4] = mbx; out[5] = mby; out[6] = mbz; out[7] = divB;
}
