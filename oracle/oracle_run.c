/*
 * oracle_run.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Ghost-cell fill, dt dispatch, step dispatch and the start()/oneStepIntegration loop of the
 * reference, restated in C.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define AT(arr, i, j, k, v) (arr)[(size_t)(i) + (size_t)isz * ((size_t)(j) + (size_t)jsz * ((size_t)(k) + (size_t)ksz * (size_t)(v)))]

/* implemented in the other oracle_*.c files */
real_t orc_compute_dt_mhd(const orc_params *P, const real_t *U);
real_t orc_compute_dt_hydro(const orc_params *P, const real_t *U);
void orc_mhd3d_step_v3(const orc_params *P, const real_t *Uold, real_t *Unew, real_t dt);
void orc_mhd2d_step_v1(const orc_params *P, const real_t *Uold, real_t *Unew, real_t dt);
void orc_hydro_step_v1(const orc_params *P, const real_t *Uold, real_t *Unew, real_t dt);
void orc_mhd3d_rotating_step(const orc_params *P, real_t *Uold, real_t *Unew, real_t dt, real_t totalTime);
void orc_make_all_boundaries_shear(const orc_params *P, real_t *U, real_t dt, real_t totalTime);

/* make_boundary_base.h:1040-1332 (CPU make_boundary2<bct,loc>); one face.
 * Dirichlet = mirror and flip the NORMAL momentum only, Neumann = copy edge cell,
 * periodic = wrap.  Full transverse extent (ghost corners included). */
static void fill_face(const orc_params *P, real_t *U, int face, int bct) {
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize, ng = P->ghostWidth;
  const int nn[3] = {P->nx, P->ny, P->nz};
  const int dir = face / 2, hi = face & 1, n = nn[dir];
  const int normalVar = IU + dir;
  if (bct != BC_DIRICHLET && bct != BC_NEUMANN && bct != BC_PERIODIC) return;
  for (int v = 0; v < P->nbVar; ++v)
    for (int g = 0; g < ng; ++g) {
      int gi = hi ? n + ng + g : g, src;
      real_t sign = 1;
      if (bct == BC_DIRICHLET) {
        src = hi ? 2 * n + 2 * ng - 1 - gi : 2 * ng - 1 - gi;
        if (v == normalVar) sign = -1;
      } else if (bct == BC_NEUMANN) {
        src = hi ? n + ng - 1 : ng;
      } else {
        src = hi ? gi - n : gi + n;
      }
      if (dir == 0) {
        for (int k = 0; k < ksz; ++k) for (int j = 0; j < jsz; ++j) AT(U, gi, j, k, v) = AT(U, src, j, k, v) * sign;
      } else if (dir == 1) {
        for (int k = 0; k < ksz; ++k) for (int i = 0; i < isz; ++i) AT(U, i, gi, k, v) = AT(U, i, src, k, v) * sign;
      } else {
        for (int j = 0; j < jsz; ++j) for (int i = 0; i < isz; ++i) AT(U, i, j, gi, v) = AT(U, i, j, src, v) * sign;
      }
    }
}

/* reference h_gravity(i,j,k,:): uniform static field (HydroRunBase.cpp:6400-6408) or the vertical field of the
 * stratified shearing box, g_z = -(phi(z+dz) - phi(z-dz)) / (2 dz) with phi = Omega0^2 z^2 / 2, optionally
 * flattened above zFloor (MHDRunBase.cpp:3163-3211; phi is held in double) */
void orc_gravity_cell(const orc_params *P, int i, int j, int k, real_t g[3]) {
  if (P->gravityMode != 3) {
    orc_gravity_at(P, k, g);
    return;
  }
  /* 2D Keplerian disc: g = -grav grad(Phi), Phi = -(r^2 + eps^2)^(-1/2); x and y themselves multiply the power, not the
     offsets to the centre (HydroRunBase.cpp:6488-6499) */
  const real_t epsilon = P->kepler[0], xCenter = P->kepler[2], yCenter = P->kepler[3], grav = P->kepler[4];
  real_t xPos = P->xMin + P->dx / 2 + (i - P->ghostWidth) * P->dx;
  real_t yPos = P->yMin + P->dy / 2 + (j - P->ghostWidth) * P->dy;
  real_t r = sqrt((xPos - xCenter) * (xPos - xCenter) + (yPos - yCenter) * (yPos - yCenter));
  real_t dphi_dx = xPos * pow(r * r + epsilon * epsilon, -3.0 / 2);
  real_t dphi_dy = yPos * pow(r * r + epsilon * epsilon, -3.0 / 2);
  g[0] = -grav * dphi_dx;
  g[1] = -grav * dphi_dy;
  g[2] = 0;
}

void orc_gravity_at(const orc_params *P, int k, real_t g[3]) {
  g[0] = g[1] = g[2] = 0;
  if (P->gravityMode == 1) {
    g[0] = P->gravity_x; g[1] = P->gravity_y; g[2] = P->gravity_z;
  } else if (P->gravityMode == 2) {
    const real_t Omega0 = P->Omega0, dz = P->dz, HALF = (real_t)0.5;
    real_t zPos = P->zMin + dz / 2 + (k - P->ghostWidth) * dz;
    double phi0 = HALF * Omega0 * Omega0 * (zPos - dz) * (zPos - dz);
    double phi1 = HALF * Omega0 * Omega0 * (zPos + dz) * (zPos + dz);
    if (P->mri_smoothGravity) {
      double zFloor = P->mri_zFloor;
      if ((zPos - dz) > zFloor) phi0 = HALF * Omega0 * Omega0 * zFloor * zFloor;
      if ((zPos + dz) > zFloor) phi1 = HALF * Omega0 * Omega0 * zFloor * zFloor;
    }
    const double zero = 0.0;
    g[0] = -HALF * (zero - zero) / P->dx;
    g[1] = -HALF * (zero - zero) / P->dy;
    g[2] = -HALF * (phi1 - phi0) / dz;
  }
}

/* make_boundary_base.h:1357-1647 make_boundary2_z_stratified_cpu (ghost width 3): hydrostatic extrapolation of
 * the density, velocities copied (outflow only for w), zero horizontal field, B_z from div B = 0.  The reference's
 * ZMAX branch differences B_y with itself (dbydy = 0, :1611-1612, :1623-1624); kept. */
static void fill_z_stratified(const orc_params *P, real_t *U, int hi) {
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize;
  const real_t dx = P->dx, dy = P->dy, dz = P->dz, HALF = (real_t)0.5;
  const real_t H = P->cIso / P->Omega0;
  const real_t factor = -dz / 2.0 / H / H;
  real_t r1 = 1, r2 = 1, r3 = 1;
  if (!P->mri_bcFloor) {
    if (!hi) {
      r1 = exp(factor * (-2 * (P->zMin + HALF * dz) + dz));
      r2 = exp(factor * (-2 * (P->zMin + HALF * dz) + 3.0 * dz));
      r3 = exp(factor * (-2 * (P->zMin + HALF * dz) + 5.0 * dz));
    } else {
      r1 = exp(factor * (2 * (P->zMax - HALF * dz) + dz));
      r2 = exp(factor * (2 * (P->zMax - HALF * dz) + 3.0 * dz));
      r3 = exp(factor * (2 * (P->zMax - HALF * dz) + 5.0 * dz));
    }
  }
  /* planes: e = first inner plane next to the face, g1..g3 = ghost planes going outwards */
  const int e = hi ? ksz - 4 : 3, s = hi ? 1 : -1;
  const int g1 = e + s, g2 = e + 2 * s, g3 = e + 3 * s;
  for (int j = 0; j < jsz; ++j)
    for (int i = 0; i < isz; ++i) {
      real_t rho_e = AT(U, i, j, e, ID);
      real_t rho1 = rho_e * r1, rho2 = rho_e * r1 * r2, rho3 = rho_e * r1 * r2 * r3;
      AT(U, i, j, g1, ID) = rho1; AT(U, i, j, g2, ID) = rho2; AT(U, i, j, g3, ID) = rho3;
      for (int v = IU; v <= IV; ++v) {
        real_t m = AT(U, i, j, e, v);
        AT(U, i, j, g3, v) = m / rho_e * rho3;
        AT(U, i, j, g2, v) = m / rho_e * rho2;
        AT(U, i, j, g1, v) = m / rho_e * rho1;
      }
      real_t w = hi ? fmax(AT(U, i, j, e, IW), 0) : fmin(AT(U, i, j, e, IW), 0);
      AT(U, i, j, g1, IW) = w; AT(U, i, j, g2, IW) = w; AT(U, i, j, g3, IW) = w;
      for (int v = IA; v <= IB; ++v) { AT(U, i, j, g1, v) = 0; AT(U, i, j, g2, v) = 0; AT(U, i, j, g3, v) = 0; }
    }
  for (int j = 0; j < jsz - 1; ++j)
    for (int i = 0; i < isz - 1; ++i) {
      if (!hi) { /* lower face: B_z on the low faces of the ghost planes 2, 1, 0 from the one of plane 3 */
        real_t bz = AT(U, i, j, 3, IC), acc = bz;
        for (int k = 2; k >= 0; --k) {
          real_t dbxdx = (AT(U, i + 1, j, k, IA) - AT(U, i, j, k, IA)) / dx;
          real_t dbydy = (AT(U, i, j + 1, k, IB) - AT(U, i, j, k, IB)) / dy;
          acc = acc + dz * (dbxdx + dbydy);
          AT(U, i, j, k, IC) = acc;
        }
      } else {   /* upper face: planes ksz-2, ksz-1 from the low-face B_z of plane ksz-3 */
        real_t bz = AT(U, i, j, ksz - 3, IC), acc = bz;
        for (int k = ksz - 3; k <= ksz - 2; ++k) {
          real_t dbxdx = (AT(U, i + 1, j, k, IA) - AT(U, i, j, k, IA)) / dx;
          real_t dbydy = (AT(U, i, j, k, IB) - AT(U, i, j, k, IB)) / dy;
          acc = acc - dz * (dbxdx + dbydy);
          AT(U, i, j, k + 1, IC) = acc;
        }
      }
    }
}

/* HydroRunBase.cpp:2374-2408 make_jet: matter injected through a square patch of the LOWER ghost rows (2D: y) or
 * planes (3D: z); hydro variables only, the magnetic field of the patch keeps the boundary values */
static void make_jet(const orc_params *P, real_t *U) {
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize, gw = P->ghostWidth;
  const int a = gw + P->offsetJet, b = gw + P->offsetJet + P->ijet;
  const real_t e = P->pjet / (P->gamma0 - 1.) + 0.5 * P->djet * P->ujet * P->ujet;
  if (P->dim == 2) {
    for (int j = 0; j < gw; ++j)
      for (int i = a; i < b; ++i) {
        AT(U, i, j, 0, ID) = P->djet; AT(U, i, j, 0, IP) = e;
        AT(U, i, j, 0, IU) = 0.0f; AT(U, i, j, 0, IV) = P->djet * P->ujet;
      }
  } else {
    for (int k = 0; k < gw; ++k)
      for (int j = a; j < b; ++j)
        for (int i = a; i < b; ++i) {
          AT(U, i, j, k, ID) = P->djet; AT(U, i, j, k, IP) = e;
          AT(U, i, j, k, IU) = 0.0f; AT(U, i, j, k, IV) = 0.0f; AT(U, i, j, k, IW) = P->djet * P->ujet;
        }
  }
}

/* HydroRunBase.cpp:2280-2316 */
void orc_make_boundaries(const orc_params *P, real_t *U, int idim) {
  int d = idim - 1;
  if (d == 2 && P->dim == 2) return;
  if (d == 2 && P->bc[4] == BC_Z_STRATIFIED) fill_z_stratified(P, U, 0); else fill_face(P, U, 2 * d, P->bc[2 * d]);
  if (d == 2 && P->bc[5] == BC_Z_STRATIFIED) fill_z_stratified(P, U, 1); else fill_face(P, U, 2 * d + 1, P->bc[2 * d + 1]);
  if (P->enableJet && d == P->dim - 1) make_jet(P, U); /* after the last direction, :2290-2291, :2310-2311 */
}

/* HydroRunBase.cpp:2333-2342: X, then Y, then Z */
void orc_make_all_boundaries(const orc_params *P, real_t *U) {
  orc_make_boundaries(P, U, 1);
  orc_make_boundaries(P, U, 2);
  if (P->dim == 3) orc_make_boundaries(P, U, 3);
}

real_t orc_compute_dt(const orc_params *P, const real_t *U) {
  return P->mhdEnabled ? orc_compute_dt_mhd(P, U) : orc_compute_dt_hydro(P, U);
}

/* MHDRunGodunov.cpp:572-594 + :1447-1503 (MHD) ; HydroRunGodunov.cpp:1820-1870 (hydro) */
void orc_godunov_unsplit(const orc_params *P, real_t *Uold, real_t *Unew, real_t dt, real_t totalTime) {
  if (P->mhdEnabled && P->Omega0 > 0 && P->dim == 3) {
    orc_mhd3d_rotating_step(P, Uold, Unew, dt, totalTime);
    return;
  }
  orc_make_all_boundaries(P, Uold);
  memcpy(Unew, Uold, (size_t)orc_array_len(P) * sizeof(real_t));
  if (P->mhdEnabled) {
    if (P->dim == 3) orc_mhd3d_step_v3(P, Uold, Unew, dt);
    else orc_mhd2d_step_v1(P, Uold, Unew, dt);
  } else {
    orc_hydro_step_v1(P, Uold, Unew, dt);
  }
}

/* the step WITHOUT its leading ghost fill (the caller has filled the ghosts of Uold, e.g. slab by
 * slab with a z-halo exchange): copy + prim + step, MHDRunGodunov.cpp:1463-1497 */
void orc_step_no_boundaries(const orc_params *P, const real_t *Uold, real_t *Unew, real_t dt) {
  memcpy(Unew, Uold, (size_t)orc_array_len(P) * sizeof(real_t));
  if (P->mhdEnabled) {
    if (P->dim == 3) orc_mhd3d_step_v3(P, Uold, Unew, dt);
    else orc_mhd2d_step_v1(P, Uold, Unew, dt);
  } else {
    orc_hydro_step_v1(P, Uold, Unew, dt);
  }
}

/* start() prologue + hot loop: MHDRunGodunov.cpp:3801-3921, :4077-4089 ; the caller has
 * already run orc_init_problem.  Ghost fill of U, copy to U2, then nsteps of
 * { dt = compute_dt(U or U2 by parity); godunov_unsplit; ++nStep; t += dt }. */
int orc_run_steps(const orc_params *P, real_t *U, real_t *U2, int nsteps, real_t *t, real_t *dt_trace) {
  if (P->mhdEnabled && P->Omega0 > 0 && P->dim == 3 && P->bc[0] == BC_SHEARINGBOX)
    orc_make_all_boundaries_shear(P, U, 0, 0);
  else
    orc_make_all_boundaries(P, U);
  memcpy(U2, U, (size_t)orc_array_len(P) * sizeof(real_t));
  real_t time = t ? *t : 0;
  int nStep = 0;
  for (; nStep < nsteps; ++nStep) {
    real_t *a = (nStep % 2 == 0) ? U : U2, *b = (nStep % 2 == 0) ? U2 : U;
    real_t dt = orc_compute_dt(P, a);
    if (dt_trace) dt_trace[nStep] = dt;
    orc_godunov_unsplit(P, a, b, dt, time);
    time += dt;
  }
  if (t) *t = time;
  return nStep % 2;
}
