// TEST INFRASTRUCTURE ONLY (see cuda_host_shim.h): C entry points around the product's point-wise device
// functions compiled for the host; tests/test_device_math_host.py compares them with the oracle.
#include "cuda_host_shim.h"

#include "config_map.h"
#include "params.h"
#include "mhd_device.cuh"
#include "hydro_device.cuh"

using namespace rg;

template <typename T>
static KParams<T> paramsOf(const char* ini) {
  const ConfigMap cfg = ConfigMap::fromText(ini);
  const RunParams rp = parseRunParams(cfg);
  return makeKParams<T>(cfg, rp, rp.nz, 0);
}

extern "C" {

// Godunov flux of n MHD Riemann problems, states in (ID, IP, IU, IV, IW, IA, IB, IC) order (k_probe_riemann)
void emu_riemann_mhd(const char* ini, int n, const double* ql, const double* qr, double* flux) {
  const KParams<double> P = paramsOf<double>(ini);
  for (int t = 0; t < n; ++t) {
    const double *l = ql + 8 * t, *r = qr + 8 * t;
    dev::State<double> L{l[ID], l[IP], l[IU], l[IV], l[IW], l[IA], l[IB], l[IC]};
    dev::State<double> R{r[ID], r[IP], r[IU], r[IV], r[IW], r[IA], r[IB], r[IC]};
    double f[8];
    dev::riemann_mhd(P, L, R, f);
    for (int v = 0; v < 8; ++v) flux[8 * t + v] = f[v];
  }
}

// corner emf of n problems, qEdge[t][4][8] in the reference's (IRT, IRB, ILT, ILB) physical layout (k_probe_emf)
void emu_compute_emf(const char* ini, int n, int emfDir, const double* qEdge, const double* xPos, double* emf) {
  const KParams<double> P = paramsOf<double>(ini);
  int iu, iv, iw, ia, ib, ic;
  if (emfDir == 2) { iu = IU; iv = IV; iw = IW; ia = IA; ib = IB; ic = IC; }
  else if (emfDir == 1) { iu = IW; iv = IU; iw = IV; ia = IC; ib = IA; ic = IB; }
  else { iu = IV; iv = IW; iw = IU; ia = IB; ib = IC; ic = IA; }
  for (int t = 0; t < n; ++t) {
    dev::Corner<double> c[4];
    for (int e = 0; e < 4; ++e) {
      const double* q = qEdge + (size_t)t * 32 + e * 8;
      c[e] = dev::Corner<double>{q[ID], q[IP], q[iu], q[iv], q[iw], q[ia], q[ib], q[ic]};
    }
    emf[t] = dev::compute_emf(P, c[0], c[1], c[2], c[3], emfDir, xPos ? xPos[t] : 0.0);
  }
}

// hydro Riemann flux (approx / HLL / HLLC by the ini), states (ID, IP, IU, IV, IW); FP64 and FP32
void emu_riemann_hydro(const char* ini, int n, const double* ql, const double* qr, double* flux) {
  const KParams<double> P = paramsOf<double>(ini);
  for (int t = 0; t < n; ++t) {
    const double *l = ql + 5 * t, *r = qr + 5 * t;
    dev::HState<double> L{l[ID], l[IP], l[IU], l[IV], l[IW]}, R{r[ID], r[IP], r[IU], r[IV], r[IW]};
    double f[5];
    dev::riemann_hydro(P, L, R, f);
    for (int v = 0; v < 5; ++v) flux[5 * t + v] = f[v];
  }
}
void emu_riemann_hydro_f32(const char* ini, int n, const float* ql, const float* qr, float* flux) {
  const KParams<float> P = paramsOf<float>(ini);
  for (int t = 0; t < n; ++t) {
    const float *l = ql + 5 * t, *r = qr + 5 * t;
    dev::HState<float> L{l[ID], l[IP], l[IU], l[IV], l[IW]}, R{r[ID], r[IP], r[IU], r[IV], r[IW]};
    float f[5];
    dev::riemann_hydro(P, L, R, f);
    for (int v = 0; v < 5; ++v) flux[5 * t + v] = f[v];
  }
}

// the two limiters (full slope, reference formulation; half slope, FP64-pipe formulation) and the scalar helpers
void emu_slopes(double st, int n, const double* qm, const double* q0, const double* qp, double* full, double* half) {
  for (int t = 0; t < n; ++t) {
    full[t] = dev::limited_slope(st, qm[t], q0[t], qp[t]);
    half[t] = dev::half_slope(0.5 * st, qm[t], q0[t], qp[t]);
  }
}
void emu_rcp_rsq(int n, const double* x, double* r, double* s, double* q) {
  for (int t = 0; t < n; ++t) { r[t] = dev::rcp(x[t]); s[t] = dev::rsq(x[t]); q[t] = dev::sqr_t(x[t]); }
}

// cons -> prim of n MHD cells: u[8] + the three +1 face fields
void emu_cons_to_prim_mhd(const char* ini, int n, const double* u, const double* bnext, double dt, double* q) {
  const KParams<double> P = paramsOf<double>(ini);
  for (int t = 0; t < n; ++t) {
    double uu[8], qq[8];
    for (int v = 0; v < 8; ++v) uu[v] = u[8 * t + v];
    dev::cons_to_prim_mhd(P, uu, bnext[3 * t], bnext[3 * t + 1], bnext[3 * t + 2], dt, qq);
    for (int v = 0; v < 8; ++v) q[8 * t + v] = qq[v];
  }
}

}  // extern "C"
