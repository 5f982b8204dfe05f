#include "init_conditions.h"

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <thread>

namespace rg {

namespace {

// 48-bit linear congruential generator with the constants of POSIX drand48, written out so that a
// slab can jump to its position in the GLOBAL stream in O(log n).
class Rand48 {
 public:
  explicit Rand48(long seed) : x_(((uint64_t)(uint32_t)seed << 16) | 0x330Eu) {}
  double next() {
    x_ = (kA * x_ + kC) & kMask;
    return (double)x_ * (1.0 / 281474976710656.0);
  }
  void skip(uint64_t n) {
    uint64_t a = kA, c = kC, accA = 1, accC = 0;  // composition of n affine maps
    while (n) {
      if (n & 1) { accA = (accA * a) & kMask; accC = (accC * a + c) & kMask; }
      c = ((a + 1) * c) & kMask;
      a = (a * a) & kMask;
      n >>= 1;
    }
    x_ = (accA * x_ + accC) & kMask;
  }

 private:
  static constexpr uint64_t kA = 0x5DEECE66DULL, kC = 0xB, kMask = (1ULL << 48) - 1;
  uint64_t x_;
};

template <typename T>
struct Grid {
  const KParams<T>& kp;
  std::vector<T>& U;
  size_t plane, comp;
  Grid(const KParams<T>& k, std::vector<T>& u) : kp(k), U(u) {
    plane = (size_t)k.isize * k.jsize;
    comp = plane * k.ksize;
  }
  T& at(int v, int i, int j, int k) { return U[(size_t)v * comp + (size_t)k * plane + (size_t)j * kp.isize + i]; }
};

template <typename T>
inline T sqr(T x) { return x * x; }

// runs fn(kBegin, kEnd) over [0, nPlanes) on a few host threads when the slab is large (planes are independent)
template <typename F>
void parallelPlanes(int nPlanes, size_t cellsPerPlane, F&& fn) {
  const size_t cells = cellsPerPlane * (size_t)nPlanes;
  unsigned nt = std::thread::hardware_concurrency();
  nt = std::min<unsigned>(std::min<unsigned>(nt ? nt : 1u, 32u), (unsigned)nPlanes);
  if (cells < (size_t)1 << 22 || nt <= 1) { fn(0, nPlanes); return; }
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < nt; ++t) {
    const int a = (int)((long)nPlanes * t / nt), b = (int)((long)nPlanes * (t + 1) / nt);
    pool.emplace_back([&fn, a, b] { fn(a, b); });
  }
  for (std::thread& th : pool) th.join();
}

// Orszag-Tang vortex; reference MHDRunBase.cpp:1378-1573 (2D, and 3D with the vortex in x-y)
template <typename T>
bool initOrszagTang(const ConfigMap& cfg, const RunParams& rp, const KParams<T>& kp, std::vector<T>& U,
                    std::string* msg) {
  if (!rp.mhdEnabled) { if (msg) *msg = "MHD must be enabled for Orszag-Tang"; return false; }
  Grid<T> g(kp, U);
  const int gw = kp.gw;
  const double TwoPi = 4.0 * std::asin(1.0);
  const double B0 = 1.0 / std::sqrt(2.0 * TwoPi);
  const double p0 = (double)(kp.gamma0 / (2.0 * TwoPi));
  const double d0 = (double)(kp.gamma0 * p0);
  const double v0 = 1.0;
  int direction = (int)cfg.getInteger("OrszagTang", "direction", 0);
  if (direction < 0 || direction > 3) direction = 0;
  if (rp.dim == 3 && direction != 0) {
    if (msg) *msg = "Orszag-Tang: only direction=0 (vortex in the x-y plane) is implemented";
    return false;
  }
  const double kt = (rp.dim == 3) ? cfg.getFloat("OrszagTang", "kt", 0.0f) : 0.0;
  const T dx = kp.dx, dy = kp.dy, dz = kp.dz;
  // the sines depend on one index each: tabulated once (the same double expressions the per-cell loop would
  // evaluate, so the values are bit-identical), and the planes are filled by a few host threads -- a 1024^3 slab
  // set-up is then bound by host memory bandwidth instead of by 4 sin() per cell on one core
  std::vector<double> sinY(kp.jsize), sinX(kp.isize), sin2X(kp.isize);
  for (int j = 0; j < kp.jsize; ++j) {
    const double yPos = kp.yMin + dy / 2 + (j - gw) * dy;
    sinY[j] = std::sin(yPos * TwoPi);
  }
  for (int i = 0; i < kp.isize; ++i) {
    const double xPos = kp.xMin + dx / 2 + (i - gw) * dx;
    sinX[i] = std::sin(xPos * TwoPi);
    sin2X[i] = std::sin(2.0 * xPos * TwoPi);
  }
  auto fillPlanes = [&](int kBegin, int kEnd) {
    for (int k = kBegin; k < kEnd; ++k) {
      const int kg = k + kp.kglob0;  // index in the global (ghost-inclusive) array
      const double zPos = kp.zMin + dz / 2 + (kg - gw) * dz;
      const double cz = (rp.dim == 3) ? std::cos(2 * TwoPi * kt * (zPos - kp.zMin) / (kp.zMax - kp.zMin)) : 1.0;
      for (int j = 0; j < kp.jsize; ++j) {
        for (int i = 0; i < kp.isize; ++i) {
          g.at(ID, i, j, k) = static_cast<T>(d0);
          g.at(IU, i, j, k) = static_cast<T>(-d0 * v0 * sinY[j]);
          g.at(IV, i, j, k) = static_cast<T>(d0 * v0 * sinX[i]);
          g.at(IW, i, j, k) = T(0);
          if (rp.dim == 3) {
            g.at(IA, i, j, k) = static_cast<T>(-B0 * cz * sinY[j]);
            g.at(IB, i, j, k) = static_cast<T>(B0 * cz * sin2X[i]);
          } else {
            g.at(IA, i, j, k) = static_cast<T>(-B0 * sinY[j]);
            g.at(IB, i, j, k) = static_cast<T>(B0 * sin2X[i]);
          }
          g.at(IC, i, j, k) = T(0);
        }
      }
      // total energy with the cell-centred field = average of the two faces (same plane only); the last
      // row/column (ghost cells, overwritten by the first ghost fill) wraps like the reference's 2D branch
      for (int j = 0; j < kp.jsize; ++j)
        for (int i = 0; i < kp.isize; ++i) {
          const bool last = (i == kp.isize - 1) || (j == kp.jsize - 1);
          if (last && rp.dim == 3) continue;  // the reference never sets these ghost energies in 3D
          const int ip = (i < kp.isize - 1) ? i + 1 : 2 * gw, jp = (j < kp.jsize - 1) ? j + 1 : 2 * gw;
          // the wrapped neighbours of the last row/column (2D) are in rows/columns this loop nest has not
          // reached only when they lie in the SAME plane, which is already filled above
          g.at(IP, i, j, k) = p0 / (kp.gamma0 - 1.0) +
                              0.5 * (sqr(g.at(IU, i, j, k)) / g.at(ID, i, j, k) + sqr(g.at(IV, i, j, k)) / g.at(ID, i, j, k) +
                                     0.25 * sqr(g.at(IA, i, j, k) + g.at(IA, ip, j, k)) +
                                     0.25 * sqr(g.at(IB, i, j, k) + g.at(IB, i, jp, k)));
        }
    }
  };
  parallelPlanes(kp.ksize, (size_t)kp.isize * kp.jsize, fillPlanes);
  return true;
}

// MRI in the shearing box (no gravity); reference MHDRunBase.cpp:2677-2760.
// One drand48 stream over the GLOBAL array in (k,j,i) order, ghosts included, 4 draws per cell.
template <typename T>
bool initMri(const ConfigMap& cfg, const RunParams& rp, const KParams<T>& kp, std::vector<T>& U, std::string* msg) {
  if (!rp.mhdEnabled || rp.dim == 2) { if (msg) *msg = "MRI needs 3D MHD"; return false; }
  if (rp.bc[0] != BC_SHEARINGBOX || rp.bc[1] != BC_SHEARINGBOX) {
    if (msg) *msg = "MRI needs shearing-box boundaries along x";
    return false;
  }
  Grid<T> g(kp, U);
  const double TwoPi = 4.0 * std::asin(1.0);
  const double d0 = cfg.getFloat("MRI", "density", 1.0f);
  const double beta = cfg.getFloat("MRI", "beta", 400.0f);
  const double p0 = d0 * kp.cIso * kp.cIso;
  const std::string type = cfg.getString("MRI", "type", "noflux");
  const T zMax = cfg.getFloat("mesh", "zmax", 1.0f);
  double B0;
  if (type == "pyl") B0 = 3.0 / 2.0 * std::sqrt(d0 * kp.Omega0 * kp.Omega0 * (zMax - kp.zMin) * (zMax - kp.zMin) / beta);
  else B0 = 2.0 * std::sqrt(p0 / beta);
  const double amp = cfg.getFloat("MRI", "amp", 0.01f);
  const long seed = cfg.getInteger("MRI", "seed", 0);
  const double d_amp = cfg.getFloat("MRI", "density_fluctuations", 0.0f);
  Rand48 rng(seed);
  rng.skip((uint64_t)4 * kp.isize * kp.jsize * (uint64_t)kp.kglob0);
  for (int k = 0; k < kp.ksize; ++k)
    for (int j = 0; j < kp.jsize; ++j)
      for (int i = 0; i < kp.isize; ++i) {
        const double xPos = kp.xMin + kp.dx / 2 + (i - kp.gw) * kp.dx;
        g.at(ID, i, j, k) = d0 * (1 + d_amp * 2 * (rng.next() - 0.5));
        g.at(IP, i, j, k) = T(0);
        g.at(IU, i, j, k) = d0 * amp * (rng.next() - 0.5) * std::sqrt(p0);
        g.at(IV, i, j, k) = d0 * amp * (rng.next() - 0.5) * std::sqrt(p0);
        g.at(IW, i, j, k) = d0 * amp * (rng.next() - 0.5) * std::sqrt(p0);
        g.at(IA, i, j, k) = T(0);
        g.at(IB, i, j, k) = T(0);
        if (type == "noflux") g.at(IC, i, j, k) = B0 * std::sin(TwoPi * xPos);
        else if (type == "pyl" || type == "fluxZ") g.at(IC, i, j, k) = B0;
        else g.at(IC, i, j, k) = T(0);
      }
  if (kp.gravity) {  // stratified disc (reference MHDRunBase.cpp:2763-2800): Gaussian density profile with a floor,
                     // toroidal field within one scale height; the velocity perturbation above is kept
    const double zFloor = cfg.getFloat("MRI", "zFloor", 5.0f), H = kp.cIso / kp.Omega0;
    for (int k = 0; k < kp.ksize; ++k) {
      const T zPos = kp.zMin + kp.dz / 2 + (k + kp.kglob0 - kp.gw) * kp.dz;
      for (int j = 0; j < kp.jsize; ++j)
        for (int i = 0; i < kp.isize; ++i) {
          g.at(ID, i, j, k) = d0 * std::fmax(std::exp(-(zPos * zPos) / 2.0 / (H * H)), std::exp(-zFloor * zFloor / 2.0));
          g.at(IA, i, j, k) = T(0);
          g.at(IB, i, j, k) = (zPos < H && zPos > -H) ? T(B0) : T(0);
          g.at(IC, i, j, k) = T(0);
        }
    }
  }
  return true;
}

// the reference copies the first inner corner cells into the ghost corners when ghostWidth == 2
// (HydroRunBase.cpp:5490-5503, 5526-5543); only corners this slab owns are touched
template <typename T>
void fillCornersGw2(const RunParams& rp, const KParams<T>& kp, Grid<T>& g) {
  if (kp.gw != 2) return;
  const int nx = kp.nx, ny = kp.ny;
  for (int v = 0; v < kp.nvar; ++v) {
    if (rp.dim == 2) {
      for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) {
          g.at(v, i, j, 0) = g.at(v, 2, 2, 0);
          g.at(v, nx + 2 + i, j, 0) = g.at(v, nx + 1, 2, 0);
          g.at(v, i, ny + 2 + j, 0) = g.at(v, 2, ny + 1, 0);
          g.at(v, nx + 2 + i, ny + 2 + j, 0) = g.at(v, nx + 1, ny + 1, 0);
        }
    } else {
      const bool first = kp.kglob0 == 0, last = kp.kglob0 + kp.nz == kp.nzGlobal;
      const int nz = kp.nz;
      for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
          for (int k = 0; k < 2; ++k) {
            if (first) {
              g.at(v, i, j, k) = g.at(v, 2, 2, 2);
              g.at(v, nx + 2 + i, j, k) = g.at(v, nx + 1, 2, 2);
              g.at(v, i, ny + 2 + j, k) = g.at(v, 2, ny + 1, 2);
              g.at(v, nx + 2 + i, ny + 2 + j, k) = g.at(v, nx + 1, ny + 1, 2);
            }
            if (last) {
              g.at(v, i, j, nz + 2 + k) = g.at(v, 2, 2, nz + 1);
              g.at(v, nx + 2 + i, j, nz + 2 + k) = g.at(v, nx + 1, 2, nz + 1);
              g.at(v, i, ny + 2 + j, nz + 2 + k) = g.at(v, 2, ny + 1, nz + 1);
              g.at(v, nx + 2 + i, ny + 2 + j, nz + 2 + k) = g.at(v, nx + 1, ny + 1, nz + 1);
            }
          }
    }
  }
}

// implosion test; reference HydroRunBase.cpp:5449-5545.  glibc rand() stream over the GLOBAL
// inner cells in (k,j,i) order, one draw per cell.
template <typename T>
bool initImplode(const ConfigMap& cfg, const RunParams& rp, const KParams<T>& kp, std::vector<T>& U,
                 std::string* msg) {
  if (rp.mhdEnabled) { if (msg) *msg = "implode is a hydro problem"; return false; }
  Grid<T> g(kp, U);
  const int gw = kp.gw, nx = kp.nx, ny = kp.ny, nzg = kp.nzGlobal;
  std::srand((unsigned)cfg.getInteger("implode", "seed", 1));
  const T amplitude = cfg.getFloat("implode", "amplitude", 0.0f);
  if (rp.dim == 2) {
    for (int j = gw; j < kp.jsize - gw; ++j)
      for (int i = gw; i < kp.isize - gw; ++i) {
        const bool hi = ((float)i / nx + (float)j / ny) > 0.5;
        g.at(ID, i, j, 0) = (hi ? 1.0f : 0.125f) + amplitude * (1.0 * std::rand() / RAND_MAX - 0.5);
        g.at(IP, i, j, 0) = (hi ? 1.0f : 0.14f) / (kp.gamma0 - 1.0f);
        g.at(IU, i, j, 0) = 0.0f;
        g.at(IV, i, j, 0) = 0.0f;
      }
  } else {
    // amplitude == 0 (the shipped parameter files): the perturbation term is an exact zero whatever rand() returns,
    // so the stream is not consumed and the planes are filled by a few host threads (1024^3: 10^9 draws saved)
    const bool noise = amplitude != T(0);
    if (noise)
      for (long n = (long)kp.kglob0 * nx * ny; n > 0; --n) (void)std::rand();  // draws of the slabs below
    auto fillPlanes = [&](int kBegin, int kEnd) {
      for (int k = std::max(kBegin, gw); k < std::min(kEnd, kp.ksize - gw); ++k) {
        const int kg = k + kp.kglob0;
        for (int j = gw; j < kp.jsize - gw; ++j)
          for (int i = gw; i < kp.isize - gw; ++i) {
            const bool hi = ((float)i / nx + (float)j / ny + (float)kg / nzg) > 0.5;
            g.at(ID, i, j, k) = (hi ? 1.0f : 0.125f) + (noise ? amplitude * (1.0 * std::rand() / RAND_MAX - 0.5) : 0.0);
            g.at(IP, i, j, k) = (hi ? 1.0f : 0.14f) / (kp.gamma0 - 1.0f);
            g.at(IU, i, j, k) = 0.0f;
            g.at(IV, i, j, k) = 0.0f;
            g.at(IW, i, j, k) = 0.0f;
          }
      }
    };
    if (noise) fillPlanes(0, kp.ksize);
    else parallelPlanes(kp.ksize, (size_t)kp.isize * kp.jsize, fillPlanes);
  }
  fillCornersGw2(rp, kp, g);
  return true;
}

// 2D Kelvin-Helmholtz (shear layers normal to y): random, Athena single-mode, Robertson et al. single-mode or plain
// sine perturbation, in the reference's order of precedence; reference HydroRunBase.cpp:5857-5913 (parameters) and
// :5915-6071 (2D branches).
template <typename T>
bool initKelvinHelmholtz2d(const ConfigMap& cfg, const RunParams& rp, const KParams<T>& kp, std::vector<T>& U,
                           std::string* msg) {
  Grid<T> g(kp, U);
  const int gw = kp.gw;
  const char* S = "kelvin-helmholtz";
  std::srand((unsigned)cfg.getInteger(S, "seed", 1));
  const T amplitude = cfg.getFloat(S, "amplitude", 0.01f);
  const bool pRand = cfg.getBool(S, "perturbation_rand", true), pSine = cfg.getBool(S, "perturbation_sine", false);
  const bool pAthena = cfg.getBool(S, "perturbation_sine_athena", false);
  const bool pRobertson = cfg.getBool(S, "perturbation_sine_robertson", false);
  const T rhoIn = cfg.getFloat(S, "rho_inner", 2.0f), rhoOut = cfg.getFloat(S, "rho_outer", 1.0f);
  const T pressure = cfg.getFloat(S, "pressure", 2.5f);
  const T innerSize = cfg.getFloat(S, "inner_size", 0.2f), outerSize = cfg.getFloat(S, "outer_size", 0.2f);
  const T vIn = cfg.getFloat(S, "vflow_in", -0.5f), vOut = cfg.getFloat(S, "vflow_out", 0.5f);
  const T xSize = kp.xMax - kp.xMin, ySize = kp.yMax - kp.yMin, yCenter = (kp.yMin + kp.yMax) * 0.5;
  if (!pRand && !pAthena && !pRobertson && !pSine) {  // (the reference leaves a zero state)
    if (msg) *msg = "Kelvin-Helmholtz: no perturbation type selected";
    return false;
  }
  const int mode = (int)cfg.getInteger(S, "mode", 4);
  const T w0 = cfg.getFloat(S, "w0", 0.1f), deltaY = cfg.getFloat(S, "deltaY", 0.03f);
  const T y1 = kp.yMin + 0.25 * ySize, y2 = kp.yMin + 0.75 * ySize;
  for (int j = gw; j < kp.jsize - gw; ++j) {
    const T yPos = kp.yMin + kp.dy / 2 + (j - gw) * kp.dy;
    const T ramp = 1.0 / (1.0 + std::exp(2 * (yPos - y1) / deltaY)) + 1.0 / (1.0 + std::exp(2 * (y2 - yPos) / deltaY));
    for (int i = gw; i < kp.isize - gw; ++i) {
      const T xPos = kp.xMin + kp.dx / 2 + (i - gw) * kp.dx;
      T& d = g.at(ID, i, j, 0);
      T& mu = g.at(IU, i, j, 0);
      T& mv = g.at(IV, i, j, 0);
      if (pRand) {
        const bool outer = std::fabs(yPos - yCenter) > outerSize * ySize;
        const T rho = outer ? rhoOut : rhoIn, vf = outer ? vOut : vIn;
        d = rho;
        mu = rho * (vf + amplitude * (1.0 * std::rand() / RAND_MAX - 0.5));
        mv = rho * (0.0f + amplitude * (1.0 * std::rand() / RAND_MAX - 0.5));
      } else if (pAthena) {
        const T a = 0.05, sigma = 0.2, vflow = 0.5;
        d = rhoIn;
        mu = rhoIn * vflow * std::tanh(yPos / a);
        mv = rhoIn * amplitude * std::sin(2.0 * M_PI * xPos) * std::exp(-(yPos * yPos) / (sigma * sigma));
      } else if (pRobertson) {
        d = rhoIn + ramp * (rhoOut - rhoIn);
        mu = d * (vIn + ramp * (vOut - vIn));
        mv = d * w0 * std::sin(mode * M_PI * xPos);
      } else {
        const T perturbVx = 0, perturbVy = amplitude * std::sin(2.0 * M_PI * xPos / xSize);
        if (std::fabs(yPos - yCenter) > outerSize * ySize) {
          d = rhoOut;
          mu = rhoOut * vOut * (1.0 + perturbVx);
          mv = rhoOut * perturbVy;
        } else if (std::fabs(yPos - yCenter) <= innerSize * ySize) {
          d = rhoIn;
          mu = rhoIn * vIn * (1.0 + perturbVx);
          mv = rhoIn * perturbVy;
        } else {  // linear transition layer
          const T interpSize = outerSize - innerSize;
          const T rhoSlope = (rhoOut - rhoIn) / (interpSize * ySize), uSlope = (vOut - vIn) / (interpSize * ySize);
          T dY, dRho, dU;
          if (yPos > yCenter) {
            dY = yPos - (yCenter + innerSize * ySize);
            dRho = rhoSlope * dY;
            dU = uSlope * dY;
          } else {
            dY = yPos - (yCenter - innerSize * ySize);
            dRho = -rhoSlope * dY;
            dU = -uSlope * dY;
          }
          d = rhoIn + dRho;
          mu = d * (vIn + dU) * (1.0 + perturbVx);
          mv = d * perturbVy;
        }
      }
      g.at(IP, i, j, 0) = pressure / (kp.gamma0 - 1.0f) + 0.5 * (sqr(mu) + sqr(mv)) / d;
    }
  }
  fillCornersGw2(rp, kp, g);
  return true;
}

// 3D Kelvin-Helmholtz (shear layer normal to z), random or single-mode perturbation;
// reference HydroRunBase.cpp:5857-5892 (parameters) and :6073-6175 (3D branches).
template <typename T>
bool initKelvinHelmholtz(const ConfigMap& cfg, const RunParams& rp, const KParams<T>& kp, std::vector<T>& U,
                         std::string* msg) {
  if (rp.mhdEnabled) {
    if (msg) *msg = "Kelvin-Helmholtz: this is the hydro variant";
    return false;
  }
  if (rp.dim == 2) return initKelvinHelmholtz2d(cfg, rp, kp, U, msg);
  Grid<T> g(kp, U);
  const int gw = kp.gw;
  const char* S = "kelvin-helmholtz";
  std::srand((unsigned)cfg.getInteger(S, "seed", 1));
  const T amplitude = cfg.getFloat(S, "amplitude", 0.01f);
  const bool pRand = cfg.getBool(S, "perturbation_rand", true);
  const bool pSine = cfg.getBool(S, "perturbation_sine", false);
  if (!pRand && !pSine) {
    if (msg) *msg = "Kelvin-Helmholtz: only perturbation_rand / perturbation_sine are implemented";
    return false;
  }
  const T rhoIn = cfg.getFloat(S, "rho_inner", 2.0f), rhoOut = cfg.getFloat(S, "rho_outer", 1.0f);
  const T pressure = cfg.getFloat(S, "pressure", 2.5f);
  const T outerSize = cfg.getFloat(S, "outer_size", 0.2f);
  const T vIn = cfg.getFloat(S, "vflow_in", -0.5f), vOut = cfg.getFloat(S, "vflow_out", 0.5f);
  const T xSize = kp.xMax - kp.xMin, zSize = kp.zMax - kp.zMin;
  const T zCenter = (kp.zMin + kp.zMax) / 2;
  if (pRand)
    for (long n = 3L * kp.kglob0 * kp.nx * kp.ny; n > 0; --n) (void)std::rand();
  for (int k = gw; k < kp.ksize - gw; ++k) {
    const T zPos = kp.zMin + kp.dz / 2 + (k + kp.kglob0 - gw) * kp.dz;
    const bool outer = std::fabs(zPos - zCenter) > outerSize * zSize;
    const T rho = outer ? rhoOut : rhoIn, vf = outer ? vOut : vIn;
    for (int j = gw; j < kp.jsize - gw; ++j)
      for (int i = gw; i < kp.isize - gw; ++i) {
        g.at(ID, i, j, k) = rho;
        if (pRand) {
          g.at(IU, i, j, k) = rho * (vf + amplitude * (1.0 * std::rand() / RAND_MAX - 0.5));
          g.at(IV, i, j, k) = rho * (0.0 + amplitude * (1.0 * std::rand() / RAND_MAX - 0.5));
          g.at(IW, i, j, k) = rho * (0.0 + amplitude * (1.0 * std::rand() / RAND_MAX - 0.5));
        } else {
          const T xPos = kp.xMin + kp.dx / 2 + (i - gw) * kp.dx;
          g.at(IU, i, j, k) = rho * vf;
          g.at(IV, i, j, k) = rho * T(0);
          g.at(IW, i, j, k) = rho * (amplitude * std::sin(2.0 * M_PI * xPos / xSize));
        }
        g.at(IP, i, j, k) = pressure / (kp.gamma0 - 1.0f) +
                            0.5 * (sqr(g.at(IU, i, j, k)) + sqr(g.at(IV, i, j, k)) + sqr(g.at(IW, i, j, k))) / g.at(ID, i, j, k);
      }
  }
  fillCornersGw2(rp, kp, g);
  return true;
}

// Rayleigh-Taylor instability in 3D (hydro and MHD): heavy fluid d1 above the mid-plane in z,
// hydrostatic pressure P0 + rho g.x, single-mode or rand() perturbation of the vertical momentum,
// uniform seed field for MHD.  Reference HydroRunBase.cpp:6262-6434 (every cell incl. ghosts, glibc
// rand() in (k,j,i) order over the WHOLE array) and MHDRunBase.cpp:2995-3040.
template <typename T>
bool initRayleighTaylor(const ConfigMap& cfg, const RunParams& rp, const KParams<T>& kp, std::vector<T>& U,
                        std::string* msg) {
  if (rp.dim != 3 && rp.mhdEnabled) {
    if (msg) *msg = "Rayleigh-Taylor: the 2D MHD variant is not implemented (3D hydro / MHD and 2D hydro are)";
    return false;
  }
  Grid<T> g(kp, U);
  const int gw = kp.gw;
  const char* S = "rayleigh-taylor";
  const T amplitude = cfg.getFloat(S, "amplitude", 0.01f);
  const T d0 = cfg.getFloat(S, "d0", 1.0f), d1 = cfg.getFloat(S, "d1", 2.0f);
  const bool randomEnabled = cfg.getBool(S, "randomEnabled", false);
  if (rp.dim == 2) {  // heavy fluid above the mid-line in y; every cell incl. ghosts (HydroRunBase.cpp:6298-6330)
    if (randomEnabled) std::srand((unsigned)cfg.getInteger(S, "random_seed", 33));
    const T P0 = 1.0f / (kp.gamma0 - 1.0f);
    const T Lx = kp.xMax - kp.xMin, Ly = kp.yMax - kp.yMin;
    for (int j = 0; j < kp.jsize; ++j) {
      const T y = kp.yMin + kp.dy / 2 + (j - gw) * kp.dy;
      for (int i = 0; i < kp.isize; ++i) {
        const T x = kp.xMin + kp.dx / 2 + (i - gw) * kp.dx;
        const T d = (y > (kp.yMin + kp.yMax) / 2) ? d1 : d0;
        g.at(ID, i, j, 0) = d;
        g.at(IP, i, j, 0) = P0 + d * (kp.gx * x + kp.gy * y);
        if (randomEnabled)
          g.at(IV, i, j, 0) = amplitude * (std::rand() * 1.0 / RAND_MAX - 0.5);
        else
          g.at(IV, i, j, 0) = amplitude * (1 + std::cos(2 * M_PI * x / Lx)) * (1 + std::cos(2 * M_PI * y / Ly)) / 4;
      }
    }
    fillCornersGw2(rp, kp, g);
    return true;
  }
  if (randomEnabled) {
    std::srand((unsigned)cfg.getInteger(S, "random_seed", 33));
    // draws of the slabs below: the reference draws for every cell of the (global) array, ghosts included;
    // local plane k is global plane k + kglob0
    for (long n = (long)kp.kglob0 * kp.isize * kp.jsize; n > 0; --n) (void)std::rand();
  }
  const T P0 = 1.0f / (kp.gamma0 - 1.0f);
  const T Lx = kp.xMax - kp.xMin, Ly = kp.yMax - kp.yMin, Lz = kp.zMax - kp.zMin;
  for (int k = 0; k < kp.ksize; ++k) {
    const T z = kp.zMin + kp.dz / 2 + (k + kp.kglob0 - gw) * kp.dz;
    for (int j = 0; j < kp.jsize; ++j) {
      const T y = kp.yMin + kp.dy / 2 + (j - gw) * kp.dy;
      for (int i = 0; i < kp.isize; ++i) {
        const T x = kp.xMin + kp.dx / 2 + (i - gw) * kp.dx;
        const T d = (z > (kp.zMin + kp.zMax) / 2) ? d1 : d0;
        g.at(ID, i, j, k) = d;
        g.at(IP, i, j, k) = P0 + d * (kp.gx * x + kp.gy * y + kp.gz * z);
        if (randomEnabled)
          g.at(IW, i, j, k) = amplitude * (std::rand() * 1.0 / RAND_MAX - 0.5);
        else
          g.at(IW, i, j, k) = amplitude * (1 + std::cos(2 * M_PI * x / Lx)) * (1 + std::cos(2 * M_PI * y / Ly)) *
                              (1 + std::cos(2 * M_PI * z / Lz)) / 8;
      }
    }
  }
  fillCornersGw2(rp, kp, g);
  if (rp.mhdEnabled) {
    const T bx = cfg.getFloat(S, "bx", 1e-8f), by = cfg.getFloat(S, "by", 1e-8f), bz = cfg.getFloat(S, "bz", 1e-8f);
    for (int k = 0; k < kp.ksize; ++k)
      for (int j = 0; j < kp.jsize; ++j)
        for (int i = 0; i < kp.isize; ++i) {
          g.at(IA, i, j, k) = bx;
          g.at(IB, i, j, k) = by;
          g.at(IC, i, j, k) = bz;
          g.at(IP, i, j, k) += 0.5 * (bx * bx + by * by + bz * bz);
        }
  }
  return true;
}

// Brio-Wu shock tube in 2D / 3D; reference MHDRunBase.cpp:1870-2100.  Inner cells only; the interface is
// at half of the (ghosted, GLOBAL) index range, the ghosts come from the boundary conditions.
template <typename T>
bool initBrioWu(const ConfigMap& cfg, const RunParams& rp, const KParams<T>& kp, std::vector<T>& U, std::string* msg) {
  if (!rp.mhdEnabled) { if (msg) *msg = "Brio-Wu needs MHD"; return false; }
  Grid<T> g(kp, U);
  const int gw = kp.gw;
  const T B0 = cfg.getFloat("BrioWu", "B0", 1.0f), B1 = cfg.getFloat("BrioWu", "B1", 0.75f);
  const T d0 = cfg.getFloat("BrioWu", "d0", 1.0f), d1 = cfg.getFloat("BrioWu", "d1", 0.125f);
  const T p0 = 1.0, p1 = 0.1;
  int direction = (int)cfg.getInteger("BrioWu", "direction", 0);
  if (direction < 0 || direction > 4) direction = 0;
  const int isize = kp.isize, jsize = kp.jsize, ksizeG = kp.nzGlobal + 2 * gw;
  const T gm1 = kp.gamma0 - 1.0f;
  if (rp.dim == 2) {
    for (int j = gw; j < jsize - gw; ++j)
      for (int i = gw; i < isize - gw; ++i) {
        bool left;
        T a, b, e;
        if (direction == 0 || direction == 2) {  // direction 2 does not exist in 2D: the reference leaves zeros
          if (direction == 2) continue;
          left = i < isize / 2;
          a = B1; b = left ? B0 : -B0;
          e = (left ? p0 : p1) / gm1 + 0.5 * (B0 * B0 + B1 * B1);
        } else if (direction == 1) {
          left = j < jsize / 2;
          a = left ? B0 : -B0; b = B1;
          e = (left ? p0 : p1) / gm1 + 0.5 * (B0 * B0 + B1 * B1);
        } else if (direction == 3) {
          left = 1.0 * i / isize + 1.0 * j / jsize < 1;
          const T s2 = std::sqrt(T(2.));
          a = left ? -B0 / s2 + B1 / s2 : B0 / s2 + B1 / s2;
          b = left ? B0 / s2 + B1 / s2 : -B0 / s2 + B1 / s2;
          e = (left ? p0 : p1) / gm1 + 0.5 * ((-B0 + B1) * (-B0 + B1) / 2 + (B0 + B1) * (B0 + B1) / 2);
        } else {  // 4: quarter circle in the lower-left corner
          const T phi = std::atan2(1.0 * j, 1.0 * i);
          left = 1.0 * i * i / (isize * isize) + 1.0 * j * j / (jsize * jsize) < 1.0 / 4;
          a = left ? T(-B0 * std::sin(phi) + B1 * std::cos(phi)) : T(B0 * std::sin(phi) + B1 * std::cos(phi));
          b = left ? T(B0 * std::cos(phi) + B1 * std::sin(phi)) : T(-B0 * std::cos(phi) + B1 * std::sin(phi));
          e = (left ? p0 : p1) / gm1 + 0.5 * (a * a + b * b);
        }
        g.at(ID, i, j, 0) = left ? d0 : d1;
        g.at(IP, i, j, 0) = e;
        g.at(IA, i, j, 0) = a;
        g.at(IB, i, j, 0) = b;
      }
    return true;
  }
  for (int k = gw; k < kp.ksize - gw; ++k) {
    const int kg = k + kp.kglob0;
    for (int j = gw; j < jsize - gw; ++j)
      for (int i = gw; i < isize - gw; ++i) {
        bool left;
        T a, b, c, e;
        if (direction == 3) {
          left = 1.0 * i / isize + 1.0 * j / jsize + 1.0 * kg / ksizeG < 1;
          const T s3 = std::sqrt(T(3.));
          a = B1 / s3;
          b = left ? B1 / s3 + B0 * std::sqrt(T(2.0 / 3)) : B1 / s3 - B0 * std::sqrt(T(2.0 / 3));
          c = left ? B1 / s3 - 2 * B0 / std::sqrt(T(6.0)) : B1 / s3 + 2 * B0 / std::sqrt(T(6.0));
          e = (left ? p0 : p1) / gm1 + 0.5 * (a * a + b * b + c * c);
        } else if (direction == 4) {
          continue;  // not defined in 3D
        } else {
          left = direction == 0 ? i < isize / 2 : direction == 1 ? j < jsize / 2 : kg < ksizeG / 2;
          const T s = left ? B0 : -B0;
          a = direction == 0 ? B1 : s;
          b = direction == 1 ? B1 : s;
          c = direction == 2 ? B1 : s;
          e = (left ? p0 : p1) / gm1 + 0.5 * (B0 * B0 + B0 * B0 + B1 * B1);
        }
        g.at(ID, i, j, k) = left ? d0 : d1;
        g.at(IP, i, j, k) = e;
        g.at(IA, i, j, k) = a;
        g.at(IB, i, j, k) = b;
        g.at(IC, i, j, k) = c;
      }
  }
  return true;
}

// MHD rotor (Balsara & Spicer 1999, Toth 2000), 2D; reference MHDRunBase.cpp:2117-2190
template <typename T>
bool initRotor(const ConfigMap& cfg, const RunParams& rp, const KParams<T>& kp, std::vector<T>& U, std::string* msg) {
  if (!rp.mhdEnabled) { if (msg) *msg = "rotor needs MHD"; return false; }
  if (rp.dim != 2) return true;  // the reference's 3D branch is empty: zero state
  Grid<T> g(kp, U);
  const int gw = kp.gw;
  const T FourPi = (T)(8.0 * std::asin(1.0));
  const T r0 = cfg.getFloat("rotor", "r0", 0.1f), r1 = cfg.getFloat("rotor", "r1", 0.115f);
  const T u0 = cfg.getFloat("rotor", "u0", 2.0f), p0 = cfg.getFloat("rotor", "p0", 1.0f);
  const T b0 = cfg.getFloat("rotor", "b0", (float)(5.0 / std::sqrt(FourPi)));
  const T xMax = cfg.getFloat("mesh", "xmax", 1.0f), yMax = cfg.getFloat("mesh", "ymax", 1.0f);
  const T xC = (xMax + kp.xMin) / 2, yC = (yMax + kp.yMin) / 2;
  for (int j = gw; j < kp.jsize - gw; ++j) {
    const T yPos = kp.yMin + kp.dx / 2 + (j - gw) * kp.dy;  // dx/2: as written in the reference
    for (int i = gw; i < kp.isize - gw; ++i) {
      const T xPos = kp.xMin + kp.dx / 2 + (i - gw) * kp.dx;
      const T r = std::sqrt((xPos - xC) * (xPos - xC) + (yPos - yC) * (yPos - yC));
      const T f_r = (r1 - r) / (r1 - r0);
      T d, mx, my;
      if (r <= r0) { d = 10.0; mx = -u0 * (yPos - yC) / r0; my = u0 * (xPos - xC) / r0; }
      else if (r <= r1) { d = 1 + 9 * f_r; mx = -f_r * u0 * (yPos - yC) / r; my = f_r * u0 * (xPos - xC) / r; }
      else { d = 1.0; mx = 0.0; my = 0.0; }
      g.at(ID, i, j, 0) = d; g.at(IU, i, j, 0) = mx; g.at(IV, i, j, 0) = my;
      g.at(IA, i, j, 0) = b0;
      const T mz = 0.0;
      g.at(IP, i, j, 0) = p0 / (kp.gamma0 - 1.0) + (mx * mx + my * my + mz * mz) / 2 / d + (b0 * b0) / 2;
    }
  }
  return true;
}

// field-loop advection (Gardiner & Stone 2005), 2D and 3D (loop in the x-y plane); reference
// MHDRunBase.cpp:2214-2400.  B = curl A on the staggered grid; in 3D the reference adds drand48 noise to
// A_z over the WHOLE (ghosted) array in (k,j,i) order: one global stream, jump-ahead per slab.
template <typename T>
bool initFieldLoop(const ConfigMap& cfg, const RunParams& rp, const KParams<T>& kp, std::vector<T>& U, std::string* msg) {
  if (!rp.mhdEnabled) { if (msg) *msg = "field loop needs MHD"; return false; }
  Grid<T> g(kp, U);
  const int gw = kp.gw, isize = kp.isize, jsize = kp.jsize;
  const T radius = cfg.getFloat("FieldLoop", "radius", 1.0f), density_in = cfg.getFloat("FieldLoop", "density_in", 1.0f);
  const T amplitude = cfg.getFloat("FieldLoop", "amplitude", 1.0f), vflow = cfg.getFloat("FieldLoop", "vflow", 1.0f);
  const T cos_theta = 2.0 / std::sqrt(5.0);
  const T sin_theta = std::sqrt(1 - cos_theta * cos_theta);
  const T dx = kp.dx, dy = kp.dy, dz = kp.dz;
  auto xOf = [&](int i) { return kp.xMin + dx / 2 + (i - gw) * dx; };
  auto yOf = [&](int j) { return kp.yMin + dy / 2 + (j - gw) * dy; };
  if (rp.dim == 2) {
    std::vector<T> Az((size_t)isize * jsize, T(0));
    for (int j = gw; j < jsize - gw + 1; ++j)
      for (int i = gw; i < isize - gw + 1; ++i) {
        const T xPos = xOf(i), yPos = yOf(j), r = std::sqrt(xPos * xPos + yPos * yPos);
        Az[(size_t)j * isize + i] = r < radius ? amplitude * (radius - r) : T(0);
      }
    for (int j = gw; j < jsize - gw; ++j)
      for (int i = gw; i < isize - gw; ++i) {
        const T xPos = xOf(i), yPos = yOf(j);
        const T diag = std::sqrt(1.0 * (rp.nx * rp.nx + rp.ny * rp.ny + rp.nz * rp.nz));
        const T r = std::sqrt(xPos * xPos + yPos * yPos);
        const T d = r < radius ? density_in : T(1.0f);
        const T mx = d * vflow * cos_theta, my = d * vflow * sin_theta, mz = d * vflow * rp.nz / diag;
        const T a = (Az[(size_t)(j + 1) * isize + i] - Az[(size_t)j * isize + i]) / dy;
        const T b = -(Az[(size_t)j * isize + i + 1] - Az[(size_t)j * isize + i]) / dx;
        g.at(ID, i, j, 0) = d; g.at(IU, i, j, 0) = mx; g.at(IV, i, j, 0) = my; g.at(IW, i, j, 0) = mz;
        g.at(IA, i, j, 0) = a; g.at(IB, i, j, 0) = b;
        g.at(IP, i, j, 0) = 1.0f / (kp.gamma0 - 1.0f) + 0.5 * (a * a + b * b) + 0.5 * (mx * mx + my * my) / d;
      }
    return true;
  }
  // 3D: only A_z is non-zero
  const double amp = cfg.getFloat("FieldLoop", "amp", 0.01f);
  Rand48 rng(cfg.getInteger("FieldLoop", "seed", 0));
  rng.skip((uint64_t)isize * jsize * (uint64_t)kp.kglob0);
  const size_t plane = (size_t)isize * jsize;
  std::vector<T> Az(plane * kp.ksize, T(0));
  for (int k = 0; k < kp.ksize; ++k)
    for (int j = 0; j < jsize; ++j)
      for (int i = 0; i < isize; ++i) {
        const T xPos = xOf(i), yPos = yOf(j);
        T a = T(0) + amp * (rng.next() - 0.5);
        const T r = std::sqrt(xPos * xPos + yPos * yPos);
        if (r < radius) a = amplitude * (radius - r);
        Az[(size_t)k * plane + (size_t)j * isize + i] = a;
      }
  for (int k = gw; k < kp.ksize - gw; ++k)
    for (int j = gw; j < jsize - gw; ++j)
      for (int i = gw; i < isize - gw; ++i) {
        const T xPos = xOf(i), yPos = yOf(j), r = std::sqrt(xPos * xPos + yPos * yPos);
        const T d = r < radius ? density_in : T(1.0f);
        const T mx = d * vflow * cos_theta, my = d * vflow * sin_theta, mz = T(0);
        const size_t c = (size_t)k * plane + (size_t)j * isize + i;
        const T zero = T(0);
        const T a = (Az[c + isize] - Az[c]) / dy - (zero - zero) / dz;
        const T b = (zero - zero) / dz - (Az[c + 1] - Az[c]) / dx;
        const T cz = (zero - zero) / dx - (zero - zero) / dy;
        g.at(ID, i, j, k) = d; g.at(IU, i, j, k) = mx; g.at(IV, i, j, k) = my; g.at(IW, i, j, k) = mz;
        g.at(IA, i, j, k) = a; g.at(IB, i, j, k) = b; g.at(IC, i, j, k) = cz;
        if (kp.cIso > 0) g.at(IP, i, j, k) = T(0);
        else g.at(IP, i, j, k) = 1.0f / (kp.gamma0 - 1.0f) + 0.5 * (a * a + b * b + cz * cz) + 0.5 * (mx * mx + my * my + mz * mz) / d;
      }
  return true;
}

// current sheet, 2D and 3D, every cell (ghosts included); reference MHDRunBase.cpp:2424-2490
template <typename T>
bool initCurrentSheet(const ConfigMap& cfg, const RunParams& rp, const KParams<T>& kp, std::vector<T>& U, std::string* msg) {
  if (!rp.mhdEnabled) { if (msg) *msg = "current sheet needs MHD"; return false; }
  Grid<T> g(kp, U);
  const T A = cfg.getFloat("CurrentSheet", "A", 0.1f), B0 = cfg.getFloat("CurrentSheet", "B0", 1.0f);
  const T beta = cfg.getFloat("CurrentSheet", "beta", 0.1f);
  for (int k = 0; k < kp.ksize; ++k)
    for (int j = 0; j < kp.jsize; ++j)
      for (int i = 0; i < kp.isize; ++i) {
        const T xPos = kp.xMin + kp.dx / 2 + (i - kp.gw) * kp.dx, yPos = kp.yMin + kp.dy / 2 + (j - kp.gw) * kp.dy;
        g.at(ID, i, j, k) = T(1);
        g.at(IP, i, j, k) = beta;
        g.at(IU, i, j, k) = T(1) * A * std::sin(M_PI * yPos);
        g.at(IB, i, j, k) = (xPos < 0.5 || xPos > 1.5) ? B0 : -B0;
      }
  return true;
}

// shear wave in the shearing box (test of the rotating-frame terms); reference MHDRunBase.cpp:2574-2660
template <typename T>
bool initShearWave(const ConfigMap& cfg, const RunParams& rp, const KParams<T>& kp, std::vector<T>& U, std::string* msg) {
  if (!rp.mhdEnabled) { if (msg) *msg = "shear wave needs MHD"; return false; }
  if (rp.bc[0] != BC_SHEARINGBOX || rp.bc[1] != BC_SHEARINGBOX) {
    if (msg) *msg = "shear wave needs shearing-box boundaries along x";
    return false;
  }
  Grid<T> g(kp, U);
  const double TwoPi = 4.0 * std::asin(1.0), d0 = 1.0;
  const double Lx = kp.dx * rp.nx, Ly = kp.dy * rp.ny;
  const double energy = cfg.getFloat("ShearWave", "energy", 1.0f);
  const double delta_vx = (-4.0e-4) * kp.cIso, delta_vy = (1.0e-4) * kp.cIso;
  const double kx0 = -4 * TwoPi / Lx, ky0 = TwoPi / Ly;
  const double xi0 = 0.5 * kp.Omega0 / d0;
  const double delta_rho = (kx0 * delta_vy - ky0 * delta_vx) / xi0;
  for (int k = 0; k < kp.ksize; ++k)
    for (int j = 0; j < kp.jsize; ++j) {
      const double yPos = kp.yMin + kp.dy / 2 + (j - kp.gw) * kp.dy;
      for (int i = 0; i < kp.isize; ++i) {
        const double xPos = kp.xMin + kp.dx / 2 + (i - kp.gw) * kp.dx;
        const T d = d0 * (1.0 - delta_rho * std::sin(kx0 * xPos + ky0 * yPos));
        g.at(ID, i, j, k) = d;
        g.at(IP, i, j, k) = energy;
        g.at(IU, i, j, k) = d * delta_vx * std::cos(kx0 * xPos + ky0 * yPos);
        g.at(IV, i, j, k) = d * delta_vy * std::cos(kx0 * xPos + ky0 * yPos);
      }
    }
  return true;
}

// magnetised Kelvin-Helmholtz (shear layers normal to y, uniform B_x), 2D and 3D; reference
// MHDRunBase.cpp:2814-2990 with its quirks: the two perturbation switches are read with getFloat (so only a
// NUMERIC value overrides the default), `pressure` is read from a section spelt "kelvin_helmholtz", and the 2D
// branch derives both positions from j.  glibc rand() over the inner cells in (k,j,i) order, 2 (2D) or 3 (3D)
// draws per cell, consumed whether or not the random perturbation is switched on.
template <typename T>
bool initKelvinHelmholtzMhd(const ConfigMap& cfg, const RunParams& rp, const KParams<T>& kp, std::vector<T>& U,
                            std::string* msg) {
  (void)msg;
  Grid<T> g(kp, U);
  const int gw = kp.gw;
  const char* S = "kelvin-helmholtz";
  std::srand((unsigned)cfg.getInteger(S, "seed", 1));
  const T amplitude = cfg.getFloat(S, "amplitude", 0.01f);
  const bool p_sine = cfg.getFloat(S, "perturbation_sine", 0.0f) != 0.0f;
  const bool p_rand = cfg.getFloat(S, "perturbation_rand", 1.0f) != 0.0f;
  const T rho_inner = cfg.getFloat(S, "rho_inner", 2.0f), rho_outer = cfg.getFloat(S, "rho_outer", 1.0f);
  const T pressure = cfg.getFloat("kelvin_helmholtz", "pressure", 2.5f);
  const T v0 = cfg.getFloat(S, "v0", 1.0f), b0 = cfg.getFloat(S, "b0", 1.0f);
  const T xMin = kp.xMin, yMin = kp.yMin, xMax = kp.xMax, yMax = kp.yMax;
  auto pert = [&](T xPos) { return p_rand * amplitude * (1.0 * std::rand() / RAND_MAX - 0.5) + p_sine * amplitude * std::sin(2 * M_PI * xPos); };
  if (rp.dim == 3)
    for (long n = 3L * kp.kglob0 * kp.nx * kp.ny; n > 0; --n) (void)std::rand();  // draws of the slabs below
  for (int k = (rp.dim == 3 ? gw : 0); k < (rp.dim == 3 ? kp.ksize - gw : 1); ++k)
    for (int j = gw; j < kp.jsize - gw; ++j) {
      const T yPos = rp.dim == 2 ? T(yMin + (yMax - yMin) * j / kp.jsize) : T(yMin + kp.dy / 2 + (j - gw) * kp.dy);
      for (int i = gw; i < kp.isize - gw; ++i) {
        const T xPos = rp.dim == 2 ? T(xMin + (xMax - xMin) * j / kp.jsize) : T(xMin + kp.dx / 2 + (i - gw) * kp.dx);
        const bool outer = yPos < yMin + 0.25 * (yMax - yMin) || yPos > yMin + 0.75 * (yMax - yMin);
        const T rho = outer ? rho_outer : rho_inner, vs = outer ? v0 : -v0;
        const T mx = rho * (vs + pert(xPos));
        const T my = rho * (pert(xPos));
        const T mz = rp.dim == 3 ? T(rho * (pert(xPos))) : T(0);
        g.at(ID, i, j, k) = rho; g.at(IU, i, j, k) = mx; g.at(IV, i, j, k) = my; g.at(IW, i, j, k) = mz;
        g.at(IA, i, j, k) = b0;
        g.at(IP, i, j, k) = pressure / (kp.gamma0 - 1.0f) + 0.5 * (mx * mx + my * my + mz * mz) / rho + 0.5 * b0 * b0;
      }
    }
  return true;
}

// jet: uniform medium at rest (+ static field for MHD), inner cells; the jet enters through the boundary patch.
// Reference HydroRunBase.cpp:5282-5350, MHDRunBase.cpp:1747-1800.
template <typename T>
bool initJet(const ConfigMap& cfg, const RunParams& rp, const KParams<T>& kp, std::vector<T>& U, std::string* msg) {
  (void)msg;
  Grid<T> g(kp, U);
  const int gw = kp.gw;
  const T Bx = cfg.getFloat("jet", "BStatic_x", 0.0f), By = cfg.getFloat("jet", "BStatic_y", 0.0f), Bz = cfg.getFloat("jet", "BStatic_z", 0.0f);
  for (int k = (rp.dim == 3 ? gw : 0); k < (rp.dim == 3 ? kp.ksize - gw : 1); ++k)
    for (int j = gw; j < kp.jsize - gw; ++j)
      for (int i = gw; i < kp.isize - gw; ++i) {
        g.at(ID, i, j, k) = 1.0f;
        if (rp.mhdEnabled) {
          g.at(IP, i, j, k) = 1.0f / (kp.gamma0 - 1.0f) + 0.5 * (rp.dim == 3 ? (Bx * Bx + By * By + Bz * Bz) : (Bx * Bx + By * By));
          g.at(IA, i, j, k) = Bx; g.at(IB, i, j, k) = By; g.at(IC, i, j, k) = Bz;
        } else {
          g.at(IP, i, j, k) = 1.0f / (kp.gamma0 - 1.0f);
        }
      }
  if (!rp.mhdEnabled) fillCornersGw2(rp, kp, g);
  return true;
}

// spherical blast wave, 2D and 3D hydro; reference HydroRunBase.cpp:5551-5680 (inner cells, defaults passed through
// getFloat, i.e. rounded to float)
template <typename T>
bool initBlast(const ConfigMap& cfg, const RunParams& rp, const KParams<T>& kp, std::vector<T>& U, std::string* msg) {
  if (rp.mhdEnabled) { if (msg) *msg = "blast is a hydro problem"; return false; }
  Grid<T> g(kp, U);
  const int gw = kp.gw;
  T radius = cfg.getFloat("blast", "radius", (float)(0.25 * (kp.xMax - kp.xMin)));
  const T cx = cfg.getFloat("blast", "center_x", (float)((kp.xMax + kp.xMin) / 2));
  const T cy = cfg.getFloat("blast", "center_y", (float)((kp.yMax + kp.yMin) / 2));
  const T cz = cfg.getFloat("blast", "center_z", (float)((kp.zMax + kp.zMin) / 2));
  const T dIn = cfg.getFloat("blast", "density_in", 1.0f), dOut = cfg.getFloat("blast", "density_out", 1.0f);
  const T pIn = cfg.getFloat("blast", "pressure_in", 10.0f), pOut = cfg.getFloat("blast", "pressure_out", 0.1f);
  radius *= radius;
  for (int k = (rp.dim == 3 ? gw : 0); k < (rp.dim == 3 ? kp.ksize - gw : 1); ++k) {
    const T zPos = kp.zMin + kp.dz / 2 + (k + kp.kglob0 - gw) * kp.dz;
    for (int j = gw; j < kp.jsize - gw; ++j) {
      const T yPos = kp.yMin + kp.dy / 2 + (j - gw) * kp.dy;
      for (int i = gw; i < kp.isize - gw; ++i) {
        const T xPos = kp.xMin + kp.dx / 2 + (i - gw) * kp.dx;
        T d2 = (xPos - cx) * (xPos - cx) + (yPos - cy) * (yPos - cy);
        if (rp.dim == 3) d2 = (xPos - cx) * (xPos - cx) + (yPos - cy) * (yPos - cy) + (zPos - cz) * (zPos - cz);
        const bool in = d2 < radius;
        g.at(ID, i, j, k) = in ? dIn : dOut;
        g.at(IP, i, j, k) = (in ? pIn : pOut) / (kp.gamma0 - 1.0f);
      }
    }
  }
  fillCornersGw2(rp, kp, g);
  return true;
}


// Sod shock tube along x; reference HydroRunBase.cpp:5358-5437
template <typename T>
bool initSod(const ConfigMap&, const RunParams& rp, const KParams<T>& kp, std::vector<T>& U, std::string* msg) {
  if (rp.mhdEnabled) { if (msg) *msg = "sod is a hydro problem"; return false; }
  Grid<T> g(kp, U);
  const int gw = kp.gw;
  for (int k = (rp.dim == 3 ? gw : 0); k < (rp.dim == 3 ? kp.ksize - gw : 1); ++k)
    for (int j = gw; j < kp.jsize - gw; ++j)
      for (int i = gw; i < kp.isize - gw; ++i) {
        const bool left = i < kp.isize / 2;
        g.at(ID, i, j, k) = left ? 1.0f : 0.125f;
        g.at(IP, i, j, k) = (left ? 1.0f : 0.1f) / (kp.gamma0 - 1.0f);
      }
  fillCornersGw2(rp, kp, g);
  return true;
}

// Gresho vortex (a vortex tube along z in 3D); reference HydroRunBase.cpp:5688-5838.  The radial profiles are
// evaluated like there: real_t operands against double literals (libm sqrt / atan2 / log / sin / cos).
template <typename T>
bool initGreshoVortex(const ConfigMap& cfg, const RunParams& rp, const KParams<T>& kp, std::vector<T>& U, std::string* msg) {
  if (rp.mhdEnabled) { if (msg) *msg = "Gresho-vortex is a hydro problem"; return false; }
  Grid<T> g(kp, U);
  const int gw = kp.gw;
  const T cx = cfg.getFloat("Gresho_vortex", "center_x", (float)((kp.xMax + kp.xMin) / 2));
  const T cy = cfg.getFloat("Gresho_vortex", "center_y", (float)((kp.yMax + kp.yMin) / 2));
  const T vbx = cfg.getFloat("Gresho_vortex", "v_bulk_x", 0.0f), vby = cfg.getFloat("Gresho_vortex", "v_bulk_y", 0.0f);
  const T vbz = cfg.getFloat("Gresho_vortex", "v_bulk_z", 0.0f);
  for (int k = (rp.dim == 3 ? gw : 0); k < (rp.dim == 3 ? kp.ksize - gw : 1); ++k)
    for (int j = gw; j < kp.jsize - gw; ++j) {
      const T yPos = kp.yMin + kp.dy / 2 + (j - gw) * kp.dy;
      for (int i = gw; i < kp.isize - gw; ++i) {
        const T xPos = kp.xMin + kp.dx / 2 + (i - gw) * kp.dx;
        const T r = std::sqrt((xPos - cx) * (xPos - cx) + (yPos - cy) * (yPos - cy));
        const T phi = std::atan2(yPos - cy, xPos - cx);
        T P, vphi;
        if (r < 0.2) {
          P = 5 + 12.5 * r * r;
          vphi = 5 * r;
        } else if (r < 0.4) {
          P = 9 + 12.5 * r * r - 20 * r + 4 * std::log(5 * r);
          vphi = 2 - 5 * r;
        } else {
          P = 3 + 4 * std::log(2.0);
          vphi = T(0);
        }
        const T mu = -std::sin(phi) * vphi + vbx, mv = std::cos(phi) * vphi + vby;
        g.at(ID, i, j, k) = 1.0f;
        g.at(IU, i, j, k) = mu;
        g.at(IV, i, j, k) = mv;
        if (rp.dim == 3) {
          g.at(IW, i, j, k) = vbz;
          g.at(IP, i, j, k) = P / (kp.gamma0 - 1.0f) + 0.5 * (sqr(mu) + sqr(mv) + sqr(vbz)) / g.at(ID, i, j, k);
        } else {
          g.at(IP, i, j, k) = P / (kp.gamma0 - 1.0f) + 0.5 * (sqr(mu) + sqr(mv)) / g.at(ID, i, j, k);
        }
      }
    }
  fillCornersGw2(rp, kp, g);
  return true;
}

// The 19 two-dimensional Riemann problems of Lax & Liu (SIAM J. Sci. Comput. 19, 1998) as the reference tabulates them
// (initHydro.cpp:25-420): primitive (rho, u, v, p) of quadrants 1 (upper right), 2 (upper left), 3 (lower left),
// 4 (lower right), single-precision literals.
static const float kLaxLiu[19][4][4] = {
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

// four-quadrant 2D Riemann problem; reference HydroRunBase.cpp:6798-6910 with primToCons_2D (constoprim.h:221-234)
template <typename T>
bool initRiemann2d(const ConfigMap& cfg, const RunParams& rp, const KParams<T>& kp, std::vector<T>& U, std::string* msg) {
  if (rp.mhdEnabled || rp.dim != 2) { if (msg) *msg = "riemann2d is a 2D hydro problem"; return false; }
  Grid<T> g(kp, U);
  const int gw = kp.gw;
  const int nb = std::min(18, std::max(0, (int)cfg.getInteger("hydro", "riemann_config_number", 0)));
  const T xt = cfg.getFloat("riemann2d", "x", 0.5f), yt = cfg.getFloat("riemann2d", "y", 0.5f);
  T q[4][4];  // conservative (ID, IP, IU, IV) of the four quadrants
  for (int n = 0; n < 4; ++n) {
    const T rho = kLaxLiu[nb][n][0], u = kLaxLiu[nb][n][1], v = kLaxLiu[nb][n][2], p = kLaxLiu[nb][n][3];
    q[n][ID] = rho;
    q[n][IU] = u * rho;
    q[n][IV] = v * rho;
    q[n][IP] = p / (kp.gamma0 - 1.0f) + rho * (u * u + v * v) * 0.5f;
  }
  for (int j = gw; j < kp.jsize - gw; ++j) {
    const T y = kp.yMin + kp.dy / 2 + (j - gw) * kp.dy;
    for (int i = gw; i < kp.isize - gw; ++i) {
      const T x = kp.xMin + kp.dx / 2 + (i - gw) * kp.dx;
      const int n = (x < xt) ? ((y < yt) ? 2 : 1) : ((y < yt) ? 3 : 0);
      for (int v = 0; v < 4; ++v) g.at(v, i, j, 0) = q[n][v];
    }
  }
  fillCornersGw2(rp, kp, g);
  return true;
}

// inertial wave (test of the Omega0 terms): uniform state with an x velocity of delta_vx * cIso, every cell incl.
// ghosts; reference MHDRunBase.cpp:2503-2556
template <typename T>
bool initInertialWave(const ConfigMap& cfg, const RunParams& rp, const KParams<T>& kp, std::vector<T>& U, std::string* msg) {
  if (!rp.mhdEnabled) { if (msg) *msg = "InertialWave is an MHD problem"; return false; }
  Grid<T> g(kp, U);
  const T density = cfg.getFloat("InertialWave", "density", 1.0f), energy = cfg.getFloat("InertialWave", "energy", 1.0f);
  T deltaVx = cfg.getFloat("InertialWave", "delta_vx", 1.0f);
  deltaVx *= kp.cIso;
  for (int k = 0; k < kp.ksize; ++k)
    for (int j = 0; j < kp.jsize; ++j)
      for (int i = 0; i < kp.isize; ++i) {
        g.at(ID, i, j, k) = density;
        g.at(IP, i, j, k) = energy;
        g.at(IU, i, j, k) = density * deltaVx;
      }
  return true;
}

// Keplerian disc around a softened point mass, 2D; reference HydroRunBase.cpp:6445-6531 (every cell incl. ghosts; the
// gravity field it also fills is keplerianGravityField, params.cpp).  libm atan2 / sqrt / pow / sin / cos.
template <typename T>
bool initKeplerianDisk(const ConfigMap& cfg, const RunParams& rp, const KParams<T>& kp, std::vector<T>& U, std::string* msg) {
  if (rp.mhdEnabled || rp.dim != 2) { if (msg) *msg = "Keplerian-disk is built as a 2D hydro problem"; return false; }
  Grid<T> g(kp, U);
  const int gw = kp.gw;
  const T epsilon = cfg.getFloat("Keplerian-disk", "epsilon", 0.01f), P0 = cfg.getFloat("Keplerian-disk", "pressure", 1e-6f);
  const T xCenter = cfg.getFloat("Keplerian-disk", "xCenter", (float)((kp.xMax + kp.xMin) / 2.0));
  const T yCenter = cfg.getFloat("Keplerian-disk", "yCenter", (float)((kp.yMax + kp.yMin) / 2.0));
  for (int j = 0; j < kp.jsize; ++j) {
    const T yPos = kp.yMin + kp.dy / 2 + (j - gw) * kp.dy;
    for (int i = 0; i < kp.isize; ++i) {
      const T xPos = kp.xMin + kp.dx / 2 + (i - gw) * kp.dx;
      const T theta = std::atan2(yPos - yCenter, xPos - xCenter);
      const T r = std::sqrt((xPos - xCenter) * (xPos - xCenter) + (yPos - yCenter) * (yPos - yCenter));
      const T velocity = r * std::pow(r * r + epsilon * epsilon, -3.0 / 4.0);
      T d = T(0);
      if (r < 0.5) d = 0.01 + std::pow(r / 0.5, 3.0);
      else if (r <= 2) d = 0.01 + 1;
      else if (r > 2) d = 0.01 + std::pow(1 + (r - 2) / 0.1, -3.0);
      const T mu = -std::sin(theta) * velocity * d, mv = std::cos(theta) * velocity * d;
      g.at(ID, i, j, 0) = d;
      g.at(IU, i, j, 0) = mu;
      g.at(IV, i, j, 0) = mv;
      g.at(IP, i, j, 0) = P0 / (kp.gamma0 - T(1)) + 0.5 * (mu * mu + mv * mv) / d;
    }
  }
  fillCornersGw2(rp, kp, g);
  return true;
}

// falling bubble in a hydrostatic atmosphere, 2D; reference HydroRunBase.cpp:6633-6712 (every cell incl. ghosts).  The
// reference's 3D branch indexes its 3D array with two indices (:6737-6744) and is not reproduced.
template <typename T>
bool initFallingBubble(const ConfigMap& cfg, const RunParams& rp, const KParams<T>& kp, std::vector<T>& U, std::string* msg) {
  if (rp.mhdEnabled || rp.dim != 2) { if (msg) *msg = "falling-bubble is built as a 2D hydro problem"; return false; }
  Grid<T> g(kp, U);
  const int gw = kp.gw;
  const char* S = "falling-bubble";
  const T P0 = 1.0f / (kp.gamma0 - 1.0f), Ly = kp.yMax - kp.yMin;
  const T radius = cfg.getFloat(S, "radius", 0.1f);
  const T xc = cfg.getFloat(S, "center_x", (float)((kp.xMin + kp.xMax) / 2));
  const T yc = cfg.getFloat(S, "center_y", (float)(kp.yMin + 0.8 * Ly));
  const T v0 = cfg.getFloat(S, "v0", 0.0f), d0 = cfg.getFloat(S, "d0", 2.0f), d1 = cfg.getFloat(S, "d1", 1.0f);
  // the pressure uses the configured field whether or not gravity is switched on ([gravity] static)
  const T gx = cfg.getFloat("gravity", "static_field_x", 0.0f), gy = cfg.getFloat("gravity", "static_field_y", 0.0f);
  for (int j = 0; j < kp.jsize; ++j) {
    const T y = kp.yMin + kp.dy / 2 + (j - gw) * kp.dy;
    for (int i = 0; i < kp.isize; ++i) {
      const T x = kp.xMin + kp.dx / 2 + (i - gw) * kp.dx;
      const T r2 = (x - xc) * (x - xc) + (y - yc) * (y - yc);
      const bool in = r2 < radius * radius;
      const T d = in ? d0 : ((y < kp.yMin + 0.3 * Ly) ? d0 : d1);
      g.at(ID, i, j, 0) = d;
      g.at(IP, i, j, 0) = P0 + d * (gx * x + gy * y);
      g.at(IV, i, j, 0) = in ? v0 : T(0);
    }
  }
  fillCornersGw2(rp, kp, g);
  return true;
}

}  // namespace

template <typename T>
bool initProblem(const ConfigMap& cfg, const RunParams& rp, const KParams<T>& kp, const std::string& problem,
                 std::vector<T>& U, std::string* message) {
  U.assign((size_t)kp.isize * kp.jsize * kp.ksize * kp.nvar, T(0));
  if (rp.mhdEnabled) {  // reference MHDRunBase.cpp:1286-1342
    if (problem == "Orszag-Tang" || problem == "OrszagTang") return initOrszagTang(cfg, rp, kp, U, message);
    if (problem == "MRI" || problem == "Mri" || problem == "mri") return initMri(cfg, rp, kp, U, message);
    if (problem == "Rayleigh-Taylor") return initRayleighTaylor(cfg, rp, kp, U, message);
    if (problem == "Kelvin-Helmholtz") return initKelvinHelmholtzMhd(cfg, rp, kp, U, message);
    if (problem == "jet" || problem == "Jet") return initJet(cfg, rp, kp, U, message);
    if (problem == "ShearWave" || problem == "shearwave" || problem == "Shear-Wave" || problem == "shear-wave" || problem == "Shearwave")
      return initShearWave(cfg, rp, kp, U, message);
    if (problem == "Brio-Wu" || problem == "BrioWu" || problem == "brio-wu" || problem == "briowu") return initBrioWu(cfg, rp, kp, U, message);
    if (problem == "Rotor" || problem == "rotor") return initRotor(cfg, rp, kp, U, message);
    if (problem == "FieldLoop" || problem == "fieldloop" || problem == "Fieldloop" || problem == "field-loop" || problem == "Field-Loop")
      return initFieldLoop(cfg, rp, kp, U, message);
    if (problem == "InertialWave" || problem == "inertialwave" || problem == "Inertial-Wave" || problem == "inertial-wave" ||
        problem == "Inertialwave")
      return initInertialWave(cfg, rp, kp, U, message);
    if (problem == "CurrentSheet" || problem == "currentsheet" || problem == "Current-Sheet" || problem == "current-sheet" || problem == "Currentsheet")
      return initCurrentSheet(cfg, rp, kp, U, message);
  } else {  // reference HydroRunBase.cpp:7023-7100
    if (problem == "implode") return initImplode(cfg, rp, kp, U, message);
    if (problem == "Kelvin-Helmholtz") return initKelvinHelmholtz(cfg, rp, kp, U, message);
    if (problem == "Rayleigh-Taylor") return initRayleighTaylor(cfg, rp, kp, U, message);
    if (problem == "jet") return initJet(cfg, rp, kp, U, message);
    if (problem == "blast") return initBlast(cfg, rp, kp, U, message);
    if (problem == "sod") return initSod(cfg, rp, kp, U, message);
    if (problem == "Gresho-vortex") return initGreshoVortex(cfg, rp, kp, U, message);
    if (problem == "riemann2d") return initRiemann2d(cfg, rp, kp, U, message);
    if (problem == "falling-bubble") return initFallingBubble(cfg, rp, kp, U, message);
    if (problem == "Keplerian-disk") return initKeplerianDisk(cfg, rp, kp, U, message);
  }
  if (message) *message = "unknown problem name '" + problem + "' for this solver";
  return false;
}

template bool initProblem<double>(const ConfigMap&, const RunParams&, const KParams<double>&, const std::string&,
                                  std::vector<double>&, std::string*);
template bool initProblem<float>(const ConfigMap&, const RunParams&, const KParams<float>&, const std::string&,
                                 std::vector<float>&, std::string*);

}  // namespace rg
