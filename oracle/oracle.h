/*
 * oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C99) of the per-timestep Godunov update path of
 * pkestene/ramsesGPU (reference CPU build "euler_cpu"), used ONLY as the checker in
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  The product
 * (ramsesgpu_b200/csrc, the CUDA path behind include/ramsesgpu_b200.h) never includes,
 * links or calls anything in this directory.
 *
 * Parity status: PINNED.  The restatement is checked bit-for-bit / to rounding against
 * the unmodified reference executable built by oracle/Makefile.ref (oracle/_ref/euler_cpu)
 * and against golden vectors generated from it (tests/golden/, script oracle/gen_golden.py).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/src/hydro unless noted).
 *
 * Precision: real_t is double unless compiled with -DORACLE_FLOAT (the reference makes the
 * same build-time choice, real_type.h:27-31).
 */
#ifndef RAMSES_ORACLE_H_
#define RAMSES_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

#ifdef ORACLE_FLOAT
typedef float real_t;
#else
typedef double real_t;
#endif

/* variable indexes, constants.h:59-71 */
enum { ID = 0, IP = 1, IU = 2, IV = 3, IW = 4, IA = 5, IB = 6, IC = 7 };
enum { NVAR_2D = 4, NVAR_3D = 5, NVAR_MHD = 8 };
/* boundary types, constants.h:209-217 */
enum { BC_UNDEFINED = 0, BC_DIRICHLET = 1, BC_NEUMANN = 2, BC_PERIODIC = 3,
       BC_SHEARINGBOX = 4, BC_COPY = 5, BC_Z_STRATIFIED = 6 };
/* Riemann solver ids, constants.h:145-160 */
enum { RS_APPROX = 0, RS_HLL = 1, RS_HLLC = 2, RS_HLLD = 3, RS_LLF = 4 };
enum { MAG_HLLD = 0, MAG_HLLF = 1, MAG_HLLA = 2, MAG_ROE = 3, MAG_LLF = 4, MAG_UPWIND = 5 };

/* ini -> parameters; HydroParameters.h:166-525 + constants.h:277-317 */
typedef struct orc_params {
  /* [run] */
  int nStepmax, nOutput;
  real_t tEnd;
  /* [mesh] */
  int nx, ny, nz, dim;          /* dim = 2 or 3 */
  int nbVar, ghostWidth;
  int isize, jsize, ksize;
  real_t xMin, xMax, yMin, yMax, zMin, zMax, dx, dy, dz;
  int bc[6];                    /* xmin,xmax,ymin,ymax,zmin,zmax */
  /* [hydro] / [MHD] */
  int mhdEnabled;
  real_t cfl, gamma0, smallr, smallc, smallp, smallpp, smalle, gamma6, cIso;
  real_t Omega0, slope_type, nu, eta;
  int niter_riemann, iorder;
  int riemannSolver, magRiemannSolver;
  int implementationVersion, unsplitVersion;
  char problem[64];
  /* problem blocks */
  int ot_direction; real_t ot_kt;                       /* [OrszagTang] */
  real_t mri_density, mri_beta, mri_amp, mri_densfluct; /* [MRI] */
  int mri_seed; char mri_type[32];
  int implode_seed; real_t implode_amp;                 /* [implode] */
  /* [kelvin-helmholtz] */
  int kh_seed, kh_p_rand, kh_p_sine, kh_p_sine_robertson;
  real_t kh_amp, kh_rho_in, kh_rho_out, kh_pressure, kh_inner, kh_outer, kh_vin, kh_vout;
  real_t kh_mode, kh_w0, kh_delta;
  /* [gravity] static field (HydroRunBase.cpp:253-260, HydroParameters.h:322-324): the reference's h_gravity
     array is uniform for the one shipped problem that fills it (Rayleigh-Taylor, :6400-6408) */
  int gravityEnabled;
  real_t gravity_x, gravity_y, gravity_z;
  /* [rayleigh-taylor] HydroRunBase.cpp:6269-6277, MHDRunBase.cpp:2999-3001 */
  int rt_random, rt_seed;
  real_t rt_amp, rt_d0, rt_d1, rt_bx, rt_by, rt_bz;
  /* [jet] inflow patch in the lower ghost rows/planes (HydroParameters.h:434-444, HydroRunBase.cpp:2374-2408) */
  int enableJet, ijet, offsetJet;
  real_t djet, ujet, pjet, cjet, jet_bx, jet_by, jet_bz;
  /* stratified shearing box (MHDRunBase.cpp:2763-2800, :3163-3211; make_boundary_base.h:1357-1647):
     gravityMode 0 = none, 1 = uniform static field (Rayleigh-Taylor), 2 = vertical field of the MRI problem */
  int gravityMode, mri_smoothGravity, mri_bcFloor;
  real_t mri_zFloor;
  real_t blast[8]; /* [blast] radius, center_x/y/z, density_in/out, pressure_in/out (HydroRunBase.cpp:5570-5577) */
  real_t gresho[5];    /* [Gresho_vortex] center_x/y, v_bulk_x/y/z (HydroRunBase.cpp:5706-5710) */
  real_t riemann2d[2]; /* [riemann2d] x, y: the transition point (HydroRunBase.cpp:6812-6813) */
  int riemannConfId;   /* [hydro] riemann_config_number (HydroRunBase.cpp:291) */
  real_t bubble[7];    /* [falling-bubble] radius, center_x/y/z, v0, d0, d1 (HydroRunBase.cpp:6658-6668) */
  real_t kepler[5];    /* [Keplerian-disk] epsilon, pressure, xCenter, yCenter; [gravity] g (HydroRunBase.cpp:6465-6469) */
} orc_params;
/* gravity field of cell plane k (reference h_gravity(i,j,k,0..2)) */
void orc_gravity_at(const orc_params *p, int k, real_t g[3]);
/* ... of cell (i, j, k): the same, or the field of the 2D Keplerian disc (gravityMode 3, HydroRunBase.cpp:6488-6499) */
void orc_gravity_cell(const orc_params *p, int i, int j, int k, real_t g[3]);

/* parse ini TEXT with the reference's inih + ConfigMap semantics (float parse!) */
int  orc_params_from_ini(const char *ini_text, orc_params *p);
long orc_array_len(const orc_params *p); /* isize*jsize*ksize*nbVar */

/* initial conditions (host), MHDRunBase.cpp:1231/1378/2677, HydroRunBase.cpp:5449/5857 */
int  orc_init_problem(const orc_params *p, real_t *U);
/* ghost fill, HydroRunBase.cpp:2322 + make_boundary_base.h:1040 */
void orc_make_all_boundaries(const orc_params *p, real_t *U);
void orc_make_boundaries(const orc_params *p, real_t *U, int idim /*1,2,3*/);
/* shearing-box variant at time totalTime + dt, MHDRunGodunov.cpp:3763-3793 */
void orc_make_all_boundaries_shear(const orc_params *p, real_t *U, real_t dt, real_t totalTime);
/* CFL time step, MHDRunBase.cpp:141-250 (MHD), HydroRunBase.cpp:314-426 (hydro) */
real_t orc_compute_dt(const orc_params *p, const real_t *U);
/* one godunov_unsplit call (boundaries of Uold, copy, prim, step):
   MHDRunGodunov.cpp:1447 -> cpu_v3 (3D) / cpu_v1 (2D); HydroRunGodunov.cpp:1820 -> cpu_v1 */
void orc_godunov_unsplit(const orc_params *p, real_t *Uold, real_t *Unew, real_t dt,
                         real_t totalTime);
void orc_step_no_boundaries(const orc_params *p, const real_t *Uold, real_t *Unew, real_t dt);
/* run n steps like start()/oneStepIntegration (MHDRunGodunov.cpp:3921,4077): returns
   final buffer index (0 -> U, 1 -> U2); dt_trace (may be NULL) receives every dt */
int orc_run_steps(const orc_params *p, real_t *U, real_t *U2, int nsteps, real_t *t,
                  real_t *dt_trace);

/* point-wise probes (known-answer tests) */
void   orc_riemann_mhd(const orc_params *p, const real_t ql[8], const real_t qr[8], real_t flux[8]);
real_t orc_compute_emf(const orc_params *p, int emfDir, const real_t qEdge[4][8], real_t xPos);
void   orc_trace_mhd_3d(const orc_params *p, const real_t q[8], const real_t dq[3][8],
                        const real_t bfNb[6], const real_t dbf[12], const real_t elec[3][2][2],
                        real_t dtdx, real_t dtdy, real_t dtdz, real_t xPos,
                        real_t qm[3][8], real_t qp[3][8], real_t qEdge[4][3][8]);
void   orc_riemann_hydro(const orc_params *p, const real_t ql[5], const real_t qr[5], real_t flux[5]);

/* history diagnostics of a 3D MHD state, MHDRunBase.cpp:3311-3410 / :3476-3620:
   out = mass, maxwell, reynolds, magp, mean_Bx, mean_By, mean_Bz, divB */
void orc_history_mhd3d(const orc_params *p, const real_t *U, double out[8]);

/* dissipative block in two stages + switch that removes it from the step drivers (slab protocol test) */
void orc_set_skip_dissipative(int on);
void orc_dissipative_stage(const orc_params *p, real_t *Unew, real_t dt, int stage);

/* test hook: the 18 trace arrays (qm[3], qp[3], qEdge[4][3]) of one 3D MHD step from a ghost-filled state */
void orc_mhd3d_trace_arrays(const orc_params *p, const real_t *Uold, real_t dt, real_t *out);

int orc_sizeof_real(void);

#ifdef __cplusplus
}
#endif
#endif
