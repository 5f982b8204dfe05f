// Run parameters: ini -> POD, with the defaults and derived quantities of the reference's
// HydroParameters constructor (src/hydro/HydroParameters.h:166-525) and GlobalConstants
// (src/hydro/constants.h:277-317).  KParams<T> is the part passed BY VALUE to every kernel
// (__grid_constant__), replacing the reference's __constant__ gParams.
#pragma once
#include <string>
#include <vector>

#include "config_map.h"

namespace rg {

// variable indexes (reference constants.h:59-71): B components are LEFT-face values
enum { ID = 0, IP = 1, IU = 2, IV = 3, IW = 4, IA = 5, IB = 6, IC = 7 };
enum { NVAR_2D = 4, NVAR_3D = 5, NVAR_MHD = 8 };
enum BoundaryType { BC_UNDEFINED = 0, BC_DIRICHLET = 1, BC_NEUMANN = 2, BC_PERIODIC = 3,
                    BC_SHEARINGBOX = 4, BC_COPY = 5, BC_Z_STRATIFIED = 6 };
enum RiemannSolver { RS_APPROX = 0, RS_HLL = 1, RS_HLLC = 2, RS_HLLD = 3, RS_LLF = 4 };
enum MagRiemannSolver { MAG_HLLD = 0, MAG_HLLF = 1, MAG_HLLA = 2, MAG_ROE = 3, MAG_LLF = 4, MAG_UPWIND = 5 };

// What kernels need. All sizes are LOCAL (this rank's z-slab), ghosts included.
template <typename T>
struct KParams {
  int nx, ny, nz;           // local inner sizes
  int isize, jsize, ksize;  // local sizes with ghosts
  int gw;
  int nvar;
  int dim;
  int kglob0;               // global k index of local k = 0 (z-slab decomposition)
  int nzGlobal;
  T xMin, yMin, zMin, xMax, yMax, zMax;
  T dx, dy, dz;
  T rdx, rdy, rdz;  // 1/dx, 1/dy, 1/dz rounded once on the host (the inverse-dt estimates multiply instead of divide)
  T gamma0, smallr, smallc, smallp, smallpp, smalle, gamma6, cIso, Omega0, slope_type, cfl;
  int niter_riemann;
  int riemannSolver, magRiemannSolver;
  // dissipative terms and static gravity: kinematic viscosity [hydro] nu, Ohmic
  // resistivity [MHD] eta, uniform field [gravity] static_field_x/y/z (reference h_gravity array,
  // uniform for the problem that fills it: HydroRunBase.cpp:6400-6408)
  T nu, eta;
  int gravity;
  T gx, gy, gz;
  // vertical field of the stratified shearing box (MRI problem with [gravity] static=yes; reference
  // init_mhd_mri_grav_field, MHDRunBase.cpp:3163-3211): g_z of every LOCAL plane, a device array of ksize reals owned
  // by the run handle (null: the uniform field above)
  const T* gzPlane;
  const T* gCell;  // 2D hydro: gravity field per cell, [2][jsize][isize] (Keplerian disc), or nullptr (uniform gx, gy)
  // jet inflow through a square patch of the lower ghost rows (2D: y) / planes (3D: z), problem "jet"
  // (reference HydroParameters.h:434-444, HydroRunBase.cpp:2374-2408)
  int jet, ijet, offsetJet;
  T djet, ujet, pjet, cjet;
};

struct RunParams {
  // [run]
  int nStepmax = 1000, nOutput = 100, nLog = 0;
  double tEnd = 0.0;
  // [mesh] (GLOBAL sizes)
  int nx = 2, ny = 2, nz = 1, dim = 2, nbVar = NVAR_2D, ghostWidth = 2;
  int bc[6] = {1, 1, 1, 1, 1, 1};
  bool mhdEnabled = false;
  std::string problem = "unknown";
  int implementationVersion = 4, unsplitVersion = 1;
  // outputs
  bool outputVtk = true, outputVtkAscii = false, outputXsm = false, ghostIncluded = false;
  std::string outputDir = "./", outputPrefix = "output";
  // [run] restart / restart_filename (reference MHDRunBase.cpp:1234-1244); [history] (MHDRunBase.cpp:3234-3283)
  bool restart = false;
  std::string restartFilename;
  bool historyEnabled = false;
  std::string historyFilename = "history.txt";
};

RunParams parseRunParams(const ConfigMap& cfg);

// g_z of the local planes of the stratified shearing box: -(phi(z+dz) - phi(z-dz)) / (2 dz) with phi = Omega0^2 z^2 / 2,
// optionally flattened above [MRI] zFloor ([MRI] smoothGravity); phi is held in double like the reference.  Returns
// false when the run has no such field (no gravity, or the uniform field of Rayleigh-Taylor).
template <typename T>
bool stratifiedGravityPlanes(const ConfigMap& cfg, const RunParams& rp, const KParams<T>& kp, std::vector<T>& gz);
// per-cell gravity field of the 2D Keplerian disc ([2][jsize][isize]); false for every other run
template <typename T>
bool keplerianGravityField(const ConfigMap& cfg, const RunParams& rp, const KParams<T>& kp, std::vector<T>& g);

// Fills the kernel parameter block in precision T (derived quantities are computed IN T, like the
// reference build for that precision).  nzLocal/kglob0 describe this rank's z-slab.
template <typename T>
KParams<T> makeKParams(const ConfigMap& cfg, const RunParams& rp, int nzLocal, int kglob0);

}  // namespace rg
