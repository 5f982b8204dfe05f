#include "params.h"

#include <cmath>

#include <algorithm>
#include <cctype>

namespace rg {

namespace {
std::string lower(std::string s) {
  std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return std::tolower(c); });
  return s;
}
}  // namespace

// reference HydroParameters.h:196-271, 352-417, 464-485; MHDRunGodunov.cpp:161-164
RunParams parseRunParams(const ConfigMap& cfg) {
  RunParams rp;
  rp.nStepmax = static_cast<int>(cfg.getInteger("run", "nstepmax", 1000));
  rp.tEnd = cfg.getFloat("run", "tend", 0.0f);
  rp.nOutput = static_cast<int>(cfg.getInteger("run", "noutput", 100));
  rp.nLog = static_cast<int>(cfg.getInteger("run", "nlog", 0));
  rp.restart = cfg.getBool("run", "restart", false);
  rp.restartFilename = cfg.getString("run", "restart_filename", "");
  rp.historyEnabled = cfg.getBool("history", "enabled", false);
  rp.historyFilename = cfg.getString("history", "filename", "history.txt");
  rp.nx = static_cast<int>(cfg.getInteger("mesh", "nx", 2));
  rp.ny = static_cast<int>(cfg.getInteger("mesh", "ny", 2));
  rp.nz = static_cast<int>(cfg.getInteger("mesh", "nz", 1));
  rp.dim = (rp.nz == 1) ? 2 : 3;
  rp.nbVar = (rp.nz == 1) ? NVAR_2D : NVAR_3D;
  rp.mhdEnabled = cfg.getBool("MHD", "enable", false);
  if (rp.mhdEnabled) rp.nbVar = NVAR_MHD;
  static const char* names[6] = {"boundary_xmin", "boundary_xmax", "boundary_ymin",
                                 "boundary_ymax", "boundary_zmin", "boundary_zmax"};
  for (int f = 0; f < 6; ++f) rp.bc[f] = static_cast<int>(cfg.getInteger("mesh", names[f], BC_DIRICHLET));
  rp.ghostWidth = static_cast<int>(cfg.getInteger("mesh", "ghostWidth", 2));
  if (rp.ghostWidth != 2 && rp.ghostWidth != 3) rp.ghostWidth = 2;
  if (rp.mhdEnabled) rp.ghostWidth = 3;
  rp.problem = cfg.getString("hydro", "problem", "unknown");
  rp.implementationVersion =
      static_cast<int>(cfg.getInteger("MHD", "implementationVersion", rp.dim == 2 ? 1 : 4));
  rp.unsplitVersion = static_cast<int>(cfg.getInteger("hydro", "unsplitVersion", 1));
  rp.outputVtk = cfg.getBool("output", "outputVtk", true);
  rp.outputVtkAscii = cfg.getBool("output", "outputVtkAscii", false);
  rp.outputXsm = cfg.getBool("output", "outputXsm", false);
  rp.ghostIncluded = cfg.getBool("output", "ghostIncluded", false);
  rp.outputDir = cfg.getString("output", "outputDir", "./");
  rp.outputPrefix = cfg.getString("output", "outputPrefix", "output");
  return rp;
}

template <typename T>
KParams<T> makeKParams(const ConfigMap& cfg, const RunParams& rp, int nzLocal, int kglob0) {
  KParams<T> k{};
  k.nx = rp.nx;
  k.ny = rp.ny;
  k.nz = (rp.dim == 2) ? 1 : nzLocal;
  k.nzGlobal = rp.nz;
  k.kglob0 = kglob0;
  k.gw = rp.ghostWidth;
  k.nvar = rp.nbVar;
  k.dim = rp.dim;
  k.isize = k.nx + 2 * k.gw;
  k.jsize = k.ny + 2 * k.gw;
  k.ksize = (rp.dim == 2) ? 1 : k.nz + 2 * k.gw;
  // geometry: float-parsed, arithmetic in T  (HydroParameters.h:238-247)
  k.xMin = cfg.getFloat("mesh", "xmin", 0.0f);
  k.xMax = cfg.getFloat("mesh", "xmax", 1.0f);
  k.yMin = cfg.getFloat("mesh", "ymin", 0.0f);
  k.yMax = cfg.getFloat("mesh", "ymax", 1.0f);
  k.zMin = cfg.getFloat("mesh", "zmin", 0.0f);
  k.zMax = cfg.getFloat("mesh", "zmax", 1.0f);
  k.dx = (k.xMax - k.xMin) / rp.nx;
  k.dy = (k.yMax - k.yMin) / rp.ny;
  k.dz = (k.zMax - k.zMin) / rp.nz;
  k.rdx = T(1) / k.dx; k.rdy = T(1) / k.dy; k.rdz = T(1) / k.dz;
  // hydro constants (HydroParameters.h:274-330)
  k.cfl = cfg.getFloat("hydro", "cfl", 0.5f);
  if (!k.cfl) k.cfl = T(0.5);
  k.cIso = cfg.getFloat("hydro", "cIso", 0.0f);
  k.gamma0 = cfg.getFloat("hydro", "gamma0", 1.4f);
  k.smallr = cfg.getFloat("hydro", "smallr", 1e-10f);
  k.smallc = cfg.getFloat("hydro", "smallc", 1e-10f);
  k.niter_riemann = static_cast<int>(cfg.getInteger("hydro", "niter_riemann", 10));
  k.smalle = T(1e-7);
  k.smallp = k.smallc * k.smallc / k.gamma0;
  if (k.cIso > 0) k.smallp = k.smallr * k.cIso * k.cIso;
  k.smallpp = k.smallr * k.smallp;
  k.gamma6 = (k.gamma0 + 1.0f) / (2.0f * k.gamma0);
  k.Omega0 = cfg.getFloat("MHD", "omega0", 0.0f);
  k.slope_type = cfg.getFloat("hydro", "slope_type", 1.0f);
  if (cfg.getInteger("hydro", "traceVersion", 1) == 0) k.slope_type = T(0);
  // Riemann solver selection (HydroParameters.h:352-417): hlld / llf only exist with MHD
  std::string rs = lower(cfg.getString("hydro", "riemannSolver", "approx"));
  k.riemannSolver = RS_APPROX;
  if (rs == "hll") k.riemannSolver = RS_HLL;
  else if (rs == "hllc") k.riemannSolver = RS_HLLC;
  else if (rp.mhdEnabled && rs == "hlld") k.riemannSolver = RS_HLLD;
  else if (rp.mhdEnabled && rs == "llf") k.riemannSolver = RS_LLF;
  k.magRiemannSolver = MAG_HLLD;
  if (rp.mhdEnabled) {
    std::string ms = lower(cfg.getString("MHD", "magRiemannSolver", "hlld"));
    if (ms == "hllf") k.magRiemannSolver = MAG_HLLF;
    else if (ms == "hlla") k.magRiemannSolver = MAG_HLLA;
    else if (ms == "roe") k.magRiemannSolver = MAG_ROE;
    else if (ms == "llf") k.magRiemannSolver = MAG_LLF;
    else if (ms == "upwind") k.magRiemannSolver = MAG_UPWIND;
  }
  // dissipative terms (HydroParameters.h nu / eta) and static gravity (HydroRunBase.cpp:253-260,
  // HydroParameters.h:322-324): only the Rayleigh-Taylor problem fills the reference's gravity array
  k.nu = cfg.getFloat("hydro", "nu", 0.0f);
  k.eta = rp.mhdEnabled ? T(cfg.getFloat("MHD", "eta", 0.0f)) : T(0);
  k.gravity = (cfg.getBool("gravity", "static", false) || cfg.getBool("gravity", "self", false)) ? 1 : 0;
  k.gx = k.gy = k.gz = T(0);
  k.gzPlane = nullptr;
  k.gCell = nullptr;
  if (rp.problem == "Keplerian-disk") k.gravity = 1;  // (HydroRunBase.cpp:258-260; the field itself: keplerianGravityField)
  if (rp.problem == "falling-bubble" && k.gravity) {  // its set-up fills the gravity array too (HydroRunBase.cpp:6701-6706)
    k.gx = cfg.getFloat("gravity", "static_field_x", 0.0f);
    k.gy = cfg.getFloat("gravity", "static_field_y", 0.0f);
    k.gz = cfg.getFloat("gravity", "static_field_z", 0.0f);
  }
  if (rp.problem == "Rayleigh-Taylor") {
    k.gravity = 1;
    k.gx = cfg.getFloat("gravity", "static_field_x", 0.0f);
    k.gy = cfg.getFloat("gravity", "static_field_y", 0.0f);
    k.gz = cfg.getFloat("gravity", "static_field_z", 0.0f);
  }
  k.jet = rp.problem == "jet" ? 1 : 0;
  k.ijet = static_cast<int>(cfg.getInteger("jet", "ijet", 0));
  k.offsetJet = static_cast<int>(cfg.getInteger("jet", "offsetJet", 0));
  k.djet = cfg.getFloat("jet", "djet", 1.0f);
  k.ujet = cfg.getFloat("jet", "ujet", 0.0f);
  k.pjet = cfg.getFloat("jet", "pjet", 0.0f);
  k.cjet = std::sqrt(k.gamma0 * k.pjet / k.djet);
  return k;
}

template <typename T>
bool stratifiedGravityPlanes(const ConfigMap& cfg, const RunParams& rp, const KParams<T>& kp, std::vector<T>& gz) {
  const bool mri = rp.problem == "MRI" || rp.problem == "Mri" || rp.problem == "mri";
  if (!kp.gravity || !rp.mhdEnabled || !mri) return false;
  const bool smooth = cfg.getBool("MRI", "smoothGravity", false);
  const double zFloor = cfg.getFloat("MRI", "zFloor", 5.0f);
  const T Omega0 = kp.Omega0, dz = kp.dz, HALF = T(0.5);
  gz.assign(kp.ksize, T(0));
  for (int k = 0; k < kp.ksize; ++k) {
    const T zPos = kp.zMin + dz / 2 + (k + kp.kglob0 - kp.gw) * dz;
    double phi0 = HALF * Omega0 * Omega0 * (zPos - dz) * (zPos - dz);
    double phi1 = HALF * Omega0 * Omega0 * (zPos + dz) * (zPos + dz);
    if (smooth) {
      if ((zPos - dz) > zFloor) phi0 = HALF * Omega0 * Omega0 * zFloor * zFloor;
      if ((zPos + dz) > zFloor) phi1 = HALF * Omega0 * Omega0 * zFloor * zFloor;
    }
    gz[k] = static_cast<T>(-HALF * (phi1 - phi0) / dz);
  }
  return true;
}
// static gravity field of the Keplerian disc, g = -grav * grad(Phi) with the softened potential Phi = -(r^2 + eps^2)^(-1/2),
// evaluated like the reference does (HydroRunBase.cpp:6488-6499: x and y themselves, not the offsets to the centre): 2D
template <typename T>
bool keplerianGravityField(const ConfigMap& cfg, const RunParams& rp, const KParams<T>& kp, std::vector<T>& g) {
  if (rp.problem != "Keplerian-disk" || rp.mhdEnabled || rp.dim != 2) return false;
  const T epsilon = cfg.getFloat("Keplerian-disk", "epsilon", 0.01f);
  const T xCenter = cfg.getFloat("Keplerian-disk", "xCenter", (float)((kp.xMax + kp.xMin) / 2.0));
  const T yCenter = cfg.getFloat("Keplerian-disk", "yCenter", (float)((kp.yMax + kp.yMin) / 2.0));
  const T grav = cfg.getFloat("gravity", "g", 1.0f);
  const size_t plane = (size_t)kp.isize * kp.jsize;
  g.assign(2 * plane, T(0));
  for (int j = 0; j < kp.jsize; ++j) {
    const T yPos = kp.yMin + kp.dy / 2 + (j - kp.gw) * kp.dy;
    for (int i = 0; i < kp.isize; ++i) {
      const T xPos = kp.xMin + kp.dx / 2 + (i - kp.gw) * kp.dx;
      const T r = std::sqrt((xPos - xCenter) * (xPos - xCenter) + (yPos - yCenter) * (yPos - yCenter));
      const T dphi_dx = xPos * std::pow(r * r + epsilon * epsilon, -3.0 / 2);
      const T dphi_dy = yPos * std::pow(r * r + epsilon * epsilon, -3.0 / 2);
      g[(size_t)j * kp.isize + i] = -grav * dphi_dx;
      g[plane + (size_t)j * kp.isize + i] = -grav * dphi_dy;
    }
  }
  return true;
}
template bool keplerianGravityField<double>(const ConfigMap&, const RunParams&, const KParams<double>&, std::vector<double>&);
template bool keplerianGravityField<float>(const ConfigMap&, const RunParams&, const KParams<float>&, std::vector<float>&);

template bool stratifiedGravityPlanes<double>(const ConfigMap&, const RunParams&, const KParams<double>&, std::vector<double>&);
template bool stratifiedGravityPlanes<float>(const ConfigMap&, const RunParams&, const KParams<float>&, std::vector<float>&);

template KParams<double> makeKParams<double>(const ConfigMap&, const RunParams&, int, int);
template KParams<float> makeKParams<float>(const ConfigMap&, const RunParams&, int, int);

}  // namespace rg
