/*
 * oracle_init.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Initial conditions used by the five BASELINE.json configs, restated from the reference:
 *   Orszag-Tang  MHDRunBase.cpp:1378-1750   (2D, and 3D direction 0/1/2)
 *   MRI          MHDRunBase.cpp:2677-2760   (no gravity)
 *   implode      HydroRunBase.cpp:5449-5535
 *   Kelvin-Helmholtz (perturbation_rand) HydroRunBase.cpp:5857-6120
 * PRNG streams are glibc's (drand48 / rand), consumed in the reference's loop order.
 */
#include "oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define AT(arr, i, j, k, v) (arr)[(size_t)(i) + (size_t)isz * ((size_t)(j) + (size_t)jsz * ((size_t)(k) + (size_t)ksz * (size_t)(v)))]
#define SQR(x) ((x) * (x))

static void init_orszag_tang(const orc_params *P, real_t *U) {
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize, gw = P->ghostWidth;
  const double TwoPi = 4.0 * asin(1.0);
  const double B0 = 1.0 / sqrt(2.0 * TwoPi);
  const double p0 = (double)(P->gamma0 / (2.0 * TwoPi));
  const double d0 = (double)(P->gamma0 * p0);
  const double v0 = 1.0;
  const real_t dx = P->dx, dy = P->dy, dz = P->dz;
  memset(U, 0, (size_t)orc_array_len(P) * sizeof(real_t));

  if (P->dim == 2) { /* :1410-1478 */
    for (int j = 0; j < jsz; ++j) {
      double yPos = P->yMin + dy / 2 + (j - gw) * dy;
      for (int i = 0; i < isz; ++i) {
        double xPos = P->xMin + dx / 2 + (i - gw) * dx;
        AT(U, i, j, 0, ID) = (real_t)d0;
        AT(U, i, j, 0, IU) = (real_t)(-d0 * v0 * sin(yPos * TwoPi));
        AT(U, i, j, 0, IV) = (real_t)(d0 * v0 * sin(xPos * TwoPi));
        AT(U, i, j, 0, IW) = 0;
        AT(U, i, j, 0, IA) = (real_t)(-B0 * sin(yPos * TwoPi));
        AT(U, i, j, 0, IB) = (real_t)(B0 * sin(2.0 * xPos * TwoPi));
        AT(U, i, j, 0, IC) = 0;
      }
    }
    for (int j = 0; j < jsz; ++j)
      for (int i = 0; i < isz; ++i) {
        int ip = (i < isz - 1) ? i + 1 : 2 * gw, jp = (j < jsz - 1) ? j + 1 : 2 * gw;
        AT(U, i, j, 0, IP) = p0 / (P->gamma0 - 1.0) +
            0.5 * (SQR(AT(U, i, j, 0, IU)) / AT(U, i, j, 0, ID) + SQR(AT(U, i, j, 0, IV)) / AT(U, i, j, 0, ID) +
                   0.25 * SQR(AT(U, i, j, 0, IA) + AT(U, ip, j, 0, IA)) +
                   0.25 * SQR(AT(U, i, j, 0, IB) + AT(U, i, jp, 0, IB)));
      }
    return;
  }

  const double kt = P->ot_kt;
  /* the three orientations differ by a cyclic relabelling (a,b,c): vortex plane (a,b), c = normal */
  /* direction 0: (x,y,z); 1: (y,z,x); 2: (z,x,y).  MHDRunBase.cpp:1489-1750 */
  int dirn = P->ot_direction;
  if (dirn != 0) {
    /* only direction 0 is exercised by the configs; the others are not restated */
    fprintf(stderr, "oracle: Orszag-Tang direction %d not restated\n", dirn);
    return;
  }
  for (int k = 0; k < ksz; ++k) {
    double zPos = P->zMin + dz / 2 + (k - gw) * dz;
    for (int j = 0; j < jsz; ++j) {
      double yPos = P->yMin + dy / 2 + (j - gw) * dy;
      for (int i = 0; i < isz; ++i) {
        double xPos = P->xMin + dx / 2 + (i - gw) * dx;
        AT(U, i, j, k, ID) = (real_t)d0;
        AT(U, i, j, k, IU) = (real_t)(-d0 * v0 * sin(yPos * TwoPi));
        AT(U, i, j, k, IV) = (real_t)(d0 * v0 * sin(xPos * TwoPi));
        AT(U, i, j, k, IW) = 0;
        AT(U, i, j, k, IA) = (real_t)(-B0 * cos(2 * TwoPi * kt * (zPos - P->zMin) / (P->zMax - P->zMin)) * sin(yPos * TwoPi));
        AT(U, i, j, k, IB) = (real_t)(B0 * cos(2 * TwoPi * kt * (zPos - P->zMin) / (P->zMax - P->zMin)) * sin(2.0 * xPos * TwoPi));
        AT(U, i, j, k, IC) = 0;
      }
    }
  }
  /* total energy :1539-1573.  In the reference the i==isize-1 / j==jsize-1 branches use the
   * 3-index accessor h_U(i,j,IP) on the 4-D array, i.e. they write element (i,j,k=1,var 0)
   * instead of the energy (Arrays.h:95-98): reproduced here.  Those cells are ghosts. */
  for (int k = 0; k < ksz; ++k)
    for (int j = 0; j < jsz; ++j)
      for (int i = 0; i < isz; ++i) {
        int ip = (i < isz - 1) ? i + 1 : 2 * gw, jp = (j < jsz - 1) ? j + 1 : 2 * gw;
        real_t e = p0 / (P->gamma0 - 1.0) +
            0.5 * (SQR(AT(U, i, j, k, IU)) / AT(U, i, j, k, ID) + SQR(AT(U, i, j, k, IV)) / AT(U, i, j, k, ID) +
                   0.25 * SQR(AT(U, i, j, k, IA) + AT(U, ip, j, k, IA)) +
                   0.25 * SQR(AT(U, i, j, k, IB) + AT(U, i, jp, k, IB)));
        if (i < isz - 1 && j < jsz - 1) AT(U, i, j, k, IP) = e;
        else AT(U, i, j, 1, ID) = e;
      }
}

static void init_mri(const orc_params *P, real_t *U) {
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize, gw = P->ghostWidth;
  if (!P->mhdEnabled || P->dim == 2) return;
  if (P->bc[0] != BC_SHEARINGBOX || P->bc[1] != BC_SHEARINGBOX) return;
  memset(U, 0, (size_t)orc_array_len(P) * sizeof(real_t));
  const double TwoPi = 4.0 * asin(1.0);
  const double d0 = P->mri_density, beta = P->mri_beta;
  const double p0 = d0 * P->cIso * P->cIso;
  double B0;
  if (!strcmp(P->mri_type, "pyl"))
    B0 = 3.0 / 2.0 * sqrt(d0 * P->Omega0 * P->Omega0 * (P->zMax - P->zMin) * (P->zMax - P->zMin) / beta);
  else
    B0 = 2.0 * sqrt(p0 / beta);
  const double amp = P->mri_amp, d_amp = P->mri_densfluct;
  srand48(P->mri_seed);
  for (int k = 0; k < ksz; ++k)
    for (int j = 0; j < jsz; ++j)
      for (int i = 0; i < isz; ++i) {
        double xPos = P->xMin + P->dx / 2 + (i - gw) * P->dx;
        AT(U, i, j, k, ID) = d0 * (1 + d_amp * 2 * (drand48() - 0.5));
        AT(U, i, j, k, IP) = 0;
        AT(U, i, j, k, IU) = d0 * amp * (drand48() - 0.5) * sqrt(p0);
        AT(U, i, j, k, IV) = d0 * amp * (drand48() - 0.5) * sqrt(p0);
        AT(U, i, j, k, IW) = d0 * amp * (drand48() - 0.5) * sqrt(p0);
        AT(U, i, j, k, IA) = 0;
        AT(U, i, j, k, IB) = 0;
        if (!strcmp(P->mri_type, "noflux")) AT(U, i, j, k, IC) = B0 * sin(TwoPi * xPos);
        else if (!strcmp(P->mri_type, "pyl") || !strcmp(P->mri_type, "fluxZ")) AT(U, i, j, k, IC) = B0;
        else AT(U, i, j, k, IC) = 0;
      }
  if (P->gravityEnabled) { /* stratified disc: MHDRunBase.cpp:2763-2800 */
    const double zFloor = P->mri_zFloor, H = P->cIso / P->Omega0;
    for (int k = 0; k < ksz; ++k) {
      real_t zPos = P->zMin + P->dz / 2 + (k - gw) * P->dz;
      for (int j = 0; j < jsz; ++j)
        for (int i = 0; i < isz; ++i) {
          AT(U, i, j, k, ID) = d0 * fmax(exp(-(zPos * zPos) / 2.0 / (H * H)), exp(-zFloor * zFloor / 2.0));
          AT(U, i, j, k, IA) = 0; AT(U, i, j, k, IB) = 0; AT(U, i, j, k, IC) = 0;
          if (zPos < H && zPos > -H) AT(U, i, j, k, IB) = B0;
        }
    }
  }
}

static void fill_corners_gw2(const orc_params *P, real_t *U) {
  /* HydroRunBase.cpp:5490-5503 / :5526-5543: only when ghostWidth == 2 */
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize, nx = P->nx, ny = P->ny, nz = P->nz;
  if (P->ghostWidth != 2) return;
  for (int v = 0; v < P->nbVar; ++v) {
    if (P->dim == 2) {
      for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) {
          AT(U, i, j, 0, v) = AT(U, 2, 2, 0, v);
          AT(U, nx + 2 + i, j, 0, v) = AT(U, nx + 1, 2, 0, v);
          AT(U, i, ny + 2 + j, 0, v) = AT(U, 2, ny + 1, 0, v);
          AT(U, nx + 2 + i, ny + 2 + j, 0, v) = AT(U, nx + 1, ny + 1, 0, v);
        }
    } else {
      for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
          for (int k = 0; k < 2; ++k) {
            AT(U, i, j, k, v) = AT(U, 2, 2, 2, v);
            AT(U, nx + 2 + i, j, k, v) = AT(U, nx + 1, 2, 2, v);
            AT(U, i, ny + 2 + j, k, v) = AT(U, 2, ny + 1, 2, v);
            AT(U, nx + 2 + i, ny + 2 + j, k, v) = AT(U, nx + 1, ny + 1, 2, v);
            AT(U, i, j, nz + 2 + k, v) = AT(U, 2, 2, nz + 1, v);
            AT(U, nx + 2 + i, j, nz + 2 + k, v) = AT(U, nx + 1, 2, nz + 1, v);
            AT(U, i, ny + 2 + j, nz + 2 + k, v) = AT(U, 2, ny + 1, nz + 1, v);
            AT(U, nx + 2 + i, ny + 2 + j, nz + 2 + k, v) = AT(U, nx + 1, ny + 1, nz + 1, v);
          }
    }
  }
}

static void init_implode(const orc_params *P, real_t *U) {
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize, gw = P->ghostWidth;
  const int nx = P->nx, ny = P->ny, nz = P->nz;
  memset(U, 0, (size_t)orc_array_len(P) * sizeof(real_t));
  srand(P->implode_seed);
  const real_t amplitude = P->implode_amp;
  if (P->dim == 2) {
    for (int j = gw; j < jsz - gw; ++j)
      for (int i = gw; i < isz - gw; ++i) {
        if (((float)i / nx + (float)j / ny) > 0.5) {
          AT(U, i, j, 0, ID) = 1.0f + amplitude * (1.0 * rand() / RAND_MAX - 0.5);
          AT(U, i, j, 0, IP) = 1.0f / (P->gamma0 - 1.0f);
        } else {
          AT(U, i, j, 0, ID) = 0.125f + amplitude * (1.0 * rand() / RAND_MAX - 0.5);
          AT(U, i, j, 0, IP) = 0.14f / (P->gamma0 - 1.0f);
        }
        AT(U, i, j, 0, IU) = 0.0f; AT(U, i, j, 0, IV) = 0.0f;
      }
  } else {
    for (int k = gw; k < ksz - gw; ++k)
      for (int j = gw; j < jsz - gw; ++j)
        for (int i = gw; i < isz - gw; ++i) {
          if (((float)i / nx + (float)j / ny + (float)k / nz) > 0.5) {
            AT(U, i, j, k, ID) = 1.0f + amplitude * (1.0 * rand() / RAND_MAX - 0.5);
            AT(U, i, j, k, IP) = 1.0f / (P->gamma0 - 1.0f);
          } else {
            AT(U, i, j, k, ID) = 0.125f + amplitude * (1.0 * rand() / RAND_MAX - 0.5);
            AT(U, i, j, k, IP) = 0.14f / (P->gamma0 - 1.0f);
          }
          AT(U, i, j, k, IU) = 0.0f; AT(U, i, j, k, IV) = 0.0f; AT(U, i, j, k, IW) = 0.0f;
        }
  }
  fill_corners_gw2(P, U);
}

/* 3D Kelvin-Helmholtz, perturbation_rand branch only (the one the configs use):
 * HydroRunBase.cpp:6073-6120; shear layer normal to z. */
static int init_kelvin_helmholtz(const orc_params *P, real_t *U) {
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize, gw = P->ghostWidth;
  memset(U, 0, (size_t)orc_array_len(P) * sizeof(real_t));
  srand(P->kh_seed);
  if (P->dim != 3 || !P->kh_p_rand) {
    fprintf(stderr, "oracle: Kelvin-Helmholtz variant not restated (3D perturbation_rand only)\n");
    return -1;
  }
  const real_t amplitude = P->kh_amp, rho_inner = P->kh_rho_in, rho_outer = P->kh_rho_out;
  const real_t pressure = P->kh_pressure, outer_size = P->kh_outer;
  const real_t vflow_in = P->kh_vin, vflow_out = P->kh_vout;
  const real_t zSize = P->zMax - P->zMin, zCenter = (P->zMin + P->zMax) / 2;
  for (int k = gw; k < ksz - gw; ++k) {
    real_t zPos = P->zMin + P->dz / 2 + (k - gw) * P->dz;
    for (int j = gw; j < jsz - gw; ++j)
      for (int i = gw; i < isz - gw; ++i) {
        real_t rho = (fabs(zPos - zCenter) > outer_size * zSize) ? rho_outer : rho_inner;
        real_t vf = (fabs(zPos - zCenter) > outer_size * zSize) ? vflow_out : vflow_in;
        AT(U, i, j, k, ID) = rho;
        AT(U, i, j, k, IU) = rho * (vf + amplitude * (1.0 * rand() / RAND_MAX - 0.5));
        AT(U, i, j, k, IV) = rho * (0.0 + amplitude * (1.0 * rand() / RAND_MAX - 0.5));
        AT(U, i, j, k, IW) = rho * (0.0 + amplitude * (1.0 * rand() / RAND_MAX - 0.5));
        AT(U, i, j, k, IP) = pressure / (P->gamma0 - 1.0f) +
            0.5 * (SQR(AT(U, i, j, k, IU)) + SQR(AT(U, i, j, k, IV)) + SQR(AT(U, i, j, k, IW))) / AT(U, i, j, k, ID);
      }
  }
  fill_corners_gw2(P, U);
  return 0;
}

/* Rayleigh-Taylor, 3D: HydroRunBase.cpp:6262-6434 (hydro part, whole array incl. ghosts) and
 * MHDRunBase.cpp:2995-3040 (uniform seed field added to the energy).  Heavy fluid d1 above the
 * mid-plane in z, hydrostatic pressure P0 + rho g.x, single-mode or rand() velocity perturbation. */
static int init_rayleigh_taylor(const orc_params *P, real_t *U) {
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize, gw = P->ghostWidth;
  memset(U, 0, (size_t)orc_array_len(P) * sizeof(real_t));
  if (P->dim != 3 && P->mhdEnabled) {
    fprintf(stderr, "oracle: Rayleigh-Taylor is restated in 3D (hydro, MHD) and 2D hydro only\n");
    return -1;
  }
  if (P->rt_random) srand(P->rt_seed);
  const real_t P0 = 1.0f / (P->gamma0 - 1.0f);
  const real_t Lx = P->xMax - P->xMin, Ly = P->yMax - P->yMin, Lz = P->zMax - P->zMin;
  if (P->dim == 2) { /* HydroRunBase.cpp:6298-6330: heavy fluid above the mid-line in y, every cell incl. ghosts */
    for (int j = 0; j < jsz; ++j) {
      real_t y = P->yMin + P->dy / 2 + (j - gw) * P->dy;
      for (int i = 0; i < isz; ++i) {
        real_t x = P->xMin + P->dx / 2 + (i - gw) * P->dx;
        real_t d = (y > (P->yMin + P->yMax) / 2) ? P->rt_d1 : P->rt_d0;
        AT(U, i, j, 0, ID) = d;
        AT(U, i, j, 0, IP) = P0 + d * (P->gravity_x * x + P->gravity_y * y);
        AT(U, i, j, 0, IU) = 0.0f;
        if (P->rt_random)
          AT(U, i, j, 0, IV) = P->rt_amp * (rand() * 1.0 / RAND_MAX - 0.5);
        else
          AT(U, i, j, 0, IV) = P->rt_amp * (1 + cos(2 * M_PI * x / Lx)) * (1 + cos(2 * M_PI * y / Ly)) / 4;
      }
    }
    fill_corners_gw2(P, U);
    return 0;
  }
  for (int k = 0; k < ksz; ++k) {
    real_t z = P->zMin + P->dz / 2 + (k - gw) * P->dz;
    for (int j = 0; j < jsz; ++j) {
      real_t y = P->yMin + P->dy / 2 + (j - gw) * P->dy;
      for (int i = 0; i < isz; ++i) {
        real_t x = P->xMin + P->dx / 2 + (i - gw) * P->dx;
        real_t d = (z > (P->zMin + P->zMax) / 2) ? P->rt_d1 : P->rt_d0;
        AT(U, i, j, k, ID) = d;
        AT(U, i, j, k, IP) = P0 + d * (P->gravity_x * x + P->gravity_y * y + P->gravity_z * z);
        AT(U, i, j, k, IU) = 0.0f;
        AT(U, i, j, k, IV) = 0.0f;
        if (P->rt_random)
          AT(U, i, j, k, IW) = P->rt_amp * (rand() * 1.0 / RAND_MAX - 0.5);
        else
          AT(U, i, j, k, IW) = P->rt_amp * (1 + cos(2 * M_PI * x / Lx)) * (1 + cos(2 * M_PI * y / Ly)) * (1 + cos(2 * M_PI * z / Lz)) / 8;
      }
    }
  }
  fill_corners_gw2(P, U);
  if (P->mhdEnabled) {
    const real_t Bx0 = P->rt_bx, By0 = P->rt_by, Bz0 = P->rt_bz;
    for (int k = 0; k < ksz; ++k)
      for (int j = 0; j < jsz; ++j)
        for (int i = 0; i < isz; ++i) {
          AT(U, i, j, k, IA) = Bx0;
          AT(U, i, j, k, IB) = By0;
          AT(U, i, j, k, IC) = Bz0;
          AT(U, i, j, k, IP) += 0.5 * (Bx0 * Bx0 + By0 * By0 + Bz0 * Bz0);
        }
  }
  return 0;
}

/* jet: uniform medium at rest (HydroRunBase.cpp:5282-5350; MHD: + static field, MHDRunBase.cpp:1747-1800),
 * inner cells only; the jet itself enters through the boundary patch of make_jet */
static int init_jet(const orc_params *P, real_t *U) {
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize, gw = P->ghostWidth;
  memset(U, 0, (size_t)orc_array_len(P) * sizeof(real_t));
  const real_t Bx = P->jet_bx, By = P->jet_by, Bz = P->jet_bz;
  for (int k = (P->dim == 3 ? gw : 0); k < (P->dim == 3 ? ksz - gw : 1); ++k)
    for (int j = gw; j < jsz - gw; ++j)
      for (int i = gw; i < isz - gw; ++i) {
        AT(U, i, j, k, ID) = 1.0f;
        if (P->mhdEnabled) {
          AT(U, i, j, k, IP) = 1.0f / (P->gamma0 - 1.0f) + 0.5 * (P->dim == 3 ? (Bx * Bx + By * By + Bz * Bz) : (Bx * Bx + By * By));
          AT(U, i, j, k, IA) = Bx; AT(U, i, j, k, IB) = By; AT(U, i, j, k, IC) = Bz;
        } else {
          AT(U, i, j, k, IP) = 1.0f / (P->gamma0 - 1.0f);
        }
      }
  if (!P->mhdEnabled) fill_corners_gw2(P, U);
  return 0;
}

/* spherical blast wave (hydro 2D/3D), HydroRunBase.cpp:5551-5680; the ini defaults go through getFloat (float) */
static int init_blast(const orc_params *P, real_t *U, const real_t par[8]) {
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize, gw = P->ghostWidth;
  memset(U, 0, (size_t)orc_array_len(P) * sizeof(real_t));
  real_t radius = par[0];
  const real_t cx = par[1], cy = par[2], cz = par[3], dIn = par[4], dOut = par[5], pIn = par[6], pOut = par[7];
  radius *= radius;
  for (int k = (P->dim == 3 ? gw : 0); k < (P->dim == 3 ? ksz - gw : 1); ++k) {
    real_t zPos = P->zMin + P->dz / 2 + (k - gw) * P->dz;
    for (int j = gw; j < jsz - gw; ++j) {
      real_t yPos = P->yMin + P->dy / 2 + (j - gw) * P->dy;
      for (int i = gw; i < isz - gw; ++i) {
        real_t xPos = P->xMin + P->dx / 2 + (i - gw) * P->dx;
        real_t d2 = (xPos - cx) * (xPos - cx) + (yPos - cy) * (yPos - cy);
        if (P->dim == 3) d2 = (xPos - cx) * (xPos - cx) + (yPos - cy) * (yPos - cy) + (zPos - cz) * (zPos - cz);
        int in = d2 < radius;
        AT(U, i, j, k, ID) = in ? dIn : dOut;
        AT(U, i, j, k, IP) = (in ? pIn : pOut) / (P->gamma0 - 1.0f);
      }
    }
  }
  fill_corners_gw2(P, U);
  return 0;
}


/* Sod shock tube along x (hydro 2D/3D), HydroRunBase.cpp:5358-5437 */
static int init_sod(const orc_params *P, real_t *U) {
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize, gw = P->ghostWidth;
  memset(U, 0, (size_t)orc_array_len(P) * sizeof(real_t));
  for (int k = (P->dim == 3 ? gw : 0); k < (P->dim == 3 ? ksz - gw : 1); ++k)
    for (int j = gw; j < jsz - gw; ++j)
      for (int i = gw; i < isz - gw; ++i) {
        if (i < isz / 2) {
          AT(U, i, j, k, ID) = 1.0f;
          AT(U, i, j, k, IP) = 1.0f / (P->gamma0 - 1.0f);
        } else {
          AT(U, i, j, k, ID) = 0.125f;
          AT(U, i, j, k, IP) = 0.1f / (P->gamma0 - 1.0f);
        }
      }
  fill_corners_gw2(P, U);
  return 0;
}

/* Gresho vortex (a vortex tube along z in 3D), HydroRunBase.cpp:5688-5838 */
static int init_gresho_vortex(const orc_params *P, real_t *U) {
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize, gw = P->ghostWidth;
  memset(U, 0, (size_t)orc_array_len(P) * sizeof(real_t));
  const real_t cx = P->gresho[0], cy = P->gresho[1], vbx = P->gresho[2], vby = P->gresho[3], vbz = P->gresho[4];
  for (int k = (P->dim == 3 ? gw : 0); k < (P->dim == 3 ? ksz - gw : 1); ++k)
    for (int j = gw; j < jsz - gw; ++j) {
      real_t yPos = P->yMin + P->dy / 2 + (j - gw) * P->dy;
      for (int i = gw; i < isz - gw; ++i) {
        real_t xPos = P->xMin + P->dx / 2 + (i - gw) * P->dx;
        real_t r = sqrt((xPos - cx) * (xPos - cx) + (yPos - cy) * (yPos - cy));
        real_t phi = atan2(yPos - cy, xPos - cx);
        real_t Pr, v_phi;
        if (r < 0.2) {
          Pr = 5 + 12.5 * r * r;
          v_phi = 5 * r;
        } else if (r < 0.4) {
          Pr = 9 + 12.5 * r * r - 20 * r + 4 * log(5 * r);
          v_phi = 2 - 5 * r;
        } else {
          Pr = 3 + 4 * log(2);
          v_phi = 0.0f;
        }
        AT(U, i, j, k, ID) = 1.0f;
        AT(U, i, j, k, IU) = -sin(phi) * v_phi + vbx;
        AT(U, i, j, k, IV) = cos(phi) * v_phi + vby;
        if (P->dim == 3) {
          AT(U, i, j, k, IW) = vbz;
          AT(U, i, j, k, IP) = Pr / (P->gamma0 - 1.0f) +
              0.5 * (SQR(AT(U, i, j, k, IU)) + SQR(AT(U, i, j, k, IV)) + SQR(AT(U, i, j, k, IW))) / AT(U, i, j, k, ID);
        } else {
          AT(U, i, j, k, IP) = Pr / (P->gamma0 - 1.0f) +
              0.5 * (SQR(AT(U, i, j, k, IU)) + SQR(AT(U, i, j, k, IV))) / AT(U, i, j, k, ID);
        }
      }
    }
  fill_corners_gw2(P, U);
  return 0;
}

/* The 19 two-dimensional Riemann problems of Lax & Liu (SIAM J. Sci. Comput. 19, 1998), as tabulated by the reference
 * (initHydro.cpp:25-420): primitive (rho, u, v, p) of quadrants 1 (upper right), 2 (upper left), 3 (lower left),
 * 4 (lower right); single-precision literals like there. */
static const float LAX_LIU[19][4][4] = {
  {{1.0f, 0.0f, 0.0f, 1.0f}, {0.5197f, -0.7259f, 0.0f, 0.4f}, {0.1072f, -0.7259f, -1.4045f, 0.0439f}, {0.2579f, 0.0f, -1.4045f, 0.15f}},
  {{1.0f, 0.0f, 0.0f, 1.0f}, {0.5197f, -0.7259f, 0.0f, 0.4f}, {1.0f, -0.7259f, -0.7259f, 1.0f}, {0.5197f, 0.0f, -0.7259f, 0.4f}},
  {{1.5f, 0.0f, 0.0f, 1.5f}, {0.5323f, 1.206f, 0.0f, 0.3f}, {0.138f, 1.206f, 1.206f, 0.029f}, {0.5323f, 0.0f, 1.206f, 0.3f}},
  {{1.1f, 0.0f, 0.0f, 1.1f}, {0.5065f, 0.8939f, 0.0f, 0.35f}, {1.1f, 0.8939f, 0.8939f, 1.1f}, {0.5065f, 0.0f, 0.8939f, 0.35f}},
  {{1.0f, -0.75f, -0.5f, 1.0f}, {2.0f, -0.75f, 0.5f, 1.0f}, {1.0f, 0.75f, 0.5f, 1.0f}, {3.0f, 0.75f, -0.5f, 1.0f}},
  {{1.0f, 0.75f, -0.5f, 1.0f}, {2.0f, 0.75f, 0.5f, 0.5f}, {1.0f, -0.75f, 0.5f, 1.0f}, {3.0f, -0.75f, -0.5f, 1.0f}},
  {{1.0f, 0.1f, 0.1f, 1.0f}, {0.5197f, -0.6259f, 0.1f, 0.4f}, {0.8f, 0.1f, 0.1f, 0.4f}, {0.5197f, 0.1f, -0.6259f, 0.4f}},
  {{0.5197f, 0.1f, 0.1f, 0.4f}, {1.0f, -0.6259f, 0.1f, 1.0f}, {0.8f, 0.1f, 0.1f, 1.0f}, {1.0f, 0.1f, -0.6259f, 1.0f}},
  {{1.0f, 0.0f, 0.3f, 1.0f}, {2.0f, 0.0f, -0.3f, 1.0f}, {1.039f, 0.0f, -0.8133f, 0.4f}, {0.5197f, 0.0f, -0.4259f, 0.4f}},
  {{1.0f, 0.0f, 0.4297f, 1.0f}, {0.5f, 0.0f, 0.6076f, 1.0f}, {0.2281f, 0.0f, -0.6076f, 0.3333f}, {0.4562f, 0.0f, -0.4259f, 0.3333f}},
  {{1.0f, 0.1f, 0.0f, 1.0f}, {0.5313f, 0.8276f, 0.0f, 0.4f}, {0.8f, 0.1f, 0.0f, 0.4f}, {0.5313f, 0.1f, 0.7276f, 0.4f}},
  {{0.5313f, 0.0f, 0.0f, 0.4f}, {1.0f, 0.7276f, 0.0f, 1.0f}, {0.8f, 0.0f, 0.0f, 1.0f}, {1.0f, 0.0f, 0.7276f, 1.0f}},
  {{1.0f, 0.0f, -0.3f, 1.0f}, {2.0f, 0.0f, 0.3f, 1.0f}, {1.0625f, 0.0f, 0.8145f, 0.4f}, {0.5313f, 0.0f, 0.4276f, 0.4f}},
  {{2.0f, 0.0f, -0.5606f, 8.0f}, {1.0f, 0.0f, -1.2172f, 8.0f}, {0.4736f, 0.0f, 1.2172f, 2.6667f}, {0.9474f, 0.0f, 1.1606f, 2.6667f}},
  {{1.0f, 0.1f, -0.3f, 1.0f}, {0.5197f, -0.6259f, -0.3f, 0.4f}, {0.8f, 0.1f, -0.3f, 0.4f}, {0.5313f, 0.1f, 0.4276f, 0.4f}},
  {{0.5313f, 0.1f, 0.1f, 0.4f}, {1.0222f, -0.6179f, 0.1f, 1.0f}, {0.8f, 0.1f, 0.1f, 1.0f}, {1.0f, 0.1f, 0.8276f, 1.0f}},
  {{1.0f, 0.0f, -0.4f, 1.0f}, {2.0f, 0.0f, -0.3f, 1.0f}, {1.0625f, 0.0f, 0.2145f, 0.4f}, {0.5197f, 0.0f, -1.1259f, 0.4f}},
  {{1.0f, 0.0f, 1.0f, 1.0f}, {2.0f, 0.0f, -0.3f, 1.0f}, {1.0625f, 0.0f, 0.2145f, 0.4f}, {0.5197f, 0.0f, 0.2741f, 0.4f}},
  {{1.0f, 0.0f, 0.3f, 1.0f}, {2.0f, 0.0f, -0.3f, 1.0f}, {1.0625f, 0.0f, 0.2145f, 0.4f}, {0.5197f, 0.0f, -0.4259f, 0.4f}},
};

/* four-quadrant 2D Riemann problem, HydroRunBase.cpp:6798-6910 (2D runs only, like the loops there) */
static int init_riemann2d(const orc_params *P, real_t *U) {
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize, gw = P->ghostWidth;
  memset(U, 0, (size_t)orc_array_len(P) * sizeof(real_t));
  if (P->dim != 2) {
    fprintf(stderr, "oracle: riemann2d is a 2D problem\n");
    return -1;
  }
  int nb = P->riemannConfId;
  if (nb < 0) nb = 0;
  else if (nb > 18) nb = 18;
  const real_t xt = P->riemann2d[0], yt = P->riemann2d[1];
  real_t q[4][4]; /* conservative (ID, IP, IU, IV) of the four quadrants, primToCons_2D (constoprim.h:221-234) */
  for (int n = 0; n < 4; ++n) {
    const real_t rho = LAX_LIU[nb][n][0], u = LAX_LIU[nb][n][1], v = LAX_LIU[nb][n][2], p = LAX_LIU[nb][n][3];
    q[n][ID] = rho;
    q[n][IU] = u * rho;
    q[n][IV] = v * rho;
    q[n][IP] = p / (P->gamma0 - 1.0f) + rho * (u * u + v * v) * 0.5f;
  }
  for (int j = gw; j < jsz - gw; ++j) {
    real_t y = P->yMin + P->dy / 2 + (j - gw) * P->dy;
    for (int i = gw; i < isz - gw; ++i) {
      real_t x = P->xMin + P->dx / 2 + (i - gw) * P->dx;
      const int n = (x < xt) ? ((y < yt) ? 2 : 1) : ((y < yt) ? 3 : 0);
      AT(U, i, j, 0, ID) = q[n][ID];
      AT(U, i, j, 0, IP) = q[n][IP];
      AT(U, i, j, 0, IU) = q[n][IU];
      AT(U, i, j, 0, IV) = q[n][IV];
    }
  }
  fill_corners_gw2(P, U);
  return 0;
}

/* Keplerian disc around a softened point mass, 2D (HydroRunBase.cpp:6445-6531), every cell incl. ghosts */
static int init_keplerian_disk(const orc_params *P, real_t *U) {
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize, gw = P->ghostWidth;
  memset(U, 0, (size_t)orc_array_len(P) * sizeof(real_t));
  if (P->dim != 2) {
    fprintf(stderr, "oracle: Keplerian-disk is restated in 2D only\n");
    return -1;
  }
  const real_t epsilon = P->kepler[0], P0 = P->kepler[1], xCenter = P->kepler[2], yCenter = P->kepler[3];
  for (int j = 0; j < jsz; ++j) {
    real_t yPos = P->yMin + P->dy / 2 + (j - gw) * P->dy;
    for (int i = 0; i < isz; ++i) {
      real_t xPos = P->xMin + P->dx / 2 + (i - gw) * P->dx;
      real_t theta = atan2(yPos - yCenter, xPos - xCenter);
      real_t r = sqrt((xPos - xCenter) * (xPos - xCenter) + (yPos - yCenter) * (yPos - yCenter));
      real_t velocity = r * pow(r * r + epsilon * epsilon, -3.0 / 4.0);
      if (r < 0.5) AT(U, i, j, 0, ID) = 0.01 + pow(r / 0.5, 3.0);
      else if (r <= 2) AT(U, i, j, 0, ID) = 0.01 + 1;
      else if (r > 2) AT(U, i, j, 0, ID) = 0.01 + pow(1 + (r - 2) / 0.1, -3.0);
      AT(U, i, j, 0, IU) = -sin(theta) * velocity * AT(U, i, j, 0, ID);
      AT(U, i, j, 0, IV) = cos(theta) * velocity * AT(U, i, j, 0, ID);
      AT(U, i, j, 0, IP) = P0 / (P->gamma0 - (real_t)1) +
          0.5 * (AT(U, i, j, 0, IU) * AT(U, i, j, 0, IU) + AT(U, i, j, 0, IV) * AT(U, i, j, 0, IV)) / AT(U, i, j, 0, ID);
    }
  }
  fill_corners_gw2(P, U);
  return 0;
}

/* falling bubble in a hydrostatic atmosphere (2D; the 3D branch of the reference indexes its 3D array with two
 * indices, HydroRunBase.cpp:6737-6744, and is not restated), HydroRunBase.cpp:6633-6712: every cell incl. ghosts */
static int init_falling_bubble(const orc_params *P, real_t *U) {
  const int isz = P->isize, jsz = P->jsize, ksz = P->ksize, gw = P->ghostWidth;
  memset(U, 0, (size_t)orc_array_len(P) * sizeof(real_t));
  if (P->dim != 2) {
    fprintf(stderr, "oracle: falling-bubble is restated in 2D only\n");
    return -1;
  }
  const real_t P0 = 1.0f / (P->gamma0 - 1.0f), Ly = P->yMax - P->yMin;
  const real_t radius = P->bubble[0], x_c = P->bubble[1], y_c = P->bubble[2], v0 = P->bubble[4], d0 = P->bubble[5], d1 = P->bubble[6];
  for (int j = 0; j < jsz; ++j) {
    real_t y = P->yMin + P->dy / 2 + (j - gw) * P->dy;
    for (int i = 0; i < isz; ++i) {
      real_t x = P->xMin + P->dx / 2 + (i - gw) * P->dx;
      AT(U, i, j, 0, ID) = (y < P->yMin + 0.3 * Ly) ? d0 : d1;
      real_t r2 = (x - x_c) * (x - x_c) + (y - y_c) * (y - y_c);
      if (r2 < radius * radius) AT(U, i, j, 0, ID) = d0;
      AT(U, i, j, 0, IP) = P0 + AT(U, i, j, 0, ID) * (P->gravity_x * x + P->gravity_y * y);
      AT(U, i, j, 0, IU) = 0.0f;
      AT(U, i, j, 0, IV) = (r2 < radius * radius) ? v0 : (real_t)0.0f;
    }
  }
  fill_corners_gw2(P, U);
  return 0;
}

/* MHDRunBase.cpp:1286-1342 (MHD) / HydroRunBase.cpp:7023-7100 (hydro) name dispatch */
int orc_init_problem(const orc_params *P, real_t *U) {
  const char *n = P->problem;
  if (P->mhdEnabled) {
    if (!strcmp(n, "Orszag-Tang") || !strcmp(n, "OrszagTang")) { init_orszag_tang(P, U); return 0; }
    if (!strcmp(n, "MRI") || !strcmp(n, "Mri") || !strcmp(n, "mri")) { init_mri(P, U); return 0; }
    if (!strcmp(n, "Rayleigh-Taylor")) return init_rayleigh_taylor(P, U);
    if (!strcmp(n, "jet") || !strcmp(n, "Jet")) return init_jet(P, U);
  } else {
    if (!strcmp(n, "jet")) return init_jet(P, U);
    if (!strcmp(n, "blast")) return init_blast(P, U, P->blast);
    if (!strcmp(n, "Rayleigh-Taylor")) return init_rayleigh_taylor(P, U);
    if (!strcmp(n, "implode")) { init_implode(P, U); return 0; }
    if (!strcmp(n, "Kelvin-Helmholtz")) return init_kelvin_helmholtz(P, U);
    if (!strcmp(n, "sod")) return init_sod(P, U);
    if (!strcmp(n, "Gresho-vortex")) return init_gresho_vortex(P, U);
    if (!strcmp(n, "riemann2d")) return init_riemann2d(P, U);
    if (!strcmp(n, "falling-bubble")) return init_falling_bubble(P, U);
    if (!strcmp(n, "Keplerian-disk")) return init_keplerian_disk(P, U);
  }
  fprintf(stderr, "oracle: problem '%s' not restated\n", n);
  return -1;
}
