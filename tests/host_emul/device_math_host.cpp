// TEST INFRASTRUCTURE ONLY (see cuda_host_shim.h): C entry points around the product's point-wise device
// functions compiled for the host; tests/test_device_math_host.py compares them with the oracle.
#include "cuda_host_shim.h"

#include "config_map.h"
#include "params.h"
#include "mhd_device.cuh"
#include "hydro_device.cuh"
#include "mhd_cells.cuh"
#include "hydro_cells.cuh"

#include <vector>

using namespace rg;

template <typename T>
static KParams<T> paramsOf(const char* ini) {
  const ConfigMap cfg = ConfigMap::fromText(ini);
  const RunParams rp = parseRunParams(cfg);
  return makeKParams<T>(cfg, rp, rp.nz, 0);
}

template <typename T>
struct HostView {  // [comp][k][j][i]
  T* p;
  int isize, jsize, ksize;
  T& operator()(int c, int i, int j, int k) const { return p[(((size_t)c * ksize + k) * jsize + j) * isize + i]; }
};

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

// The trace stage of one 3D MHD step with the product's per-cell functions (mhd_cells.cuh) on host arrays:
// cons -> prim, edge electric fields, trace -> W, then the face / edge states the flux and emf stages rebuild from W,
// un-rotated to physical component order and laid out like the oracle's 18 trace arrays
// (orc_mhd3d_trace_arrays: qm[3], qp[3], qEdge[4][3], each [var][k][j][i]).  Cells outside the range the product
// traces (gw-1 .. size-gw) and states that would need W of a cell beyond it are left at 0.
void emu_mhd3d_trace_arrays(const char* ini, const double* Uin, double dt, double* out) {
  const KParams<double> P = paramsOf<double>(ini);
  const int is = P.isize, js = P.jsize, ks = P.ksize, gw = P.gw;
  const size_t ncell = (size_t)is * js * ks;
  std::vector<double> Qs(ncell * 8, 0.0), ELs(ncell * 3, 0.0), Ws(ncell * NW_MHD, 0.0);
  const HostView<const double> U{Uin, is, js, ks};
  const HostView<double> Q{Qs.data(), is, js, ks}, EL{ELs.data(), is, js, ks}, W{Ws.data(), is, js, ks};
  for (int k = 0; k < ks - 1; ++k)      // k_prim: 0 .. size-2
    for (int j = 0; j < js - 1; ++j)
      for (int i = 0; i < is - 1; ++i) {
        double u[8], q[8];
        for (int v = 0; v < 8; ++v) u[v] = U(v, i, j, k);
        dev::cons_to_prim_mhd(P, u, U(IA, i + 1, j, k), U(IB, i, j + 1, k), U(IC, i, j, k + 1), dt, q);
        for (int v = 0; v < 8; ++v) Q(v, i, j, k) = q[v];
      }
  for (int k = 1; k < ks - 1; ++k)      // k_elec: 1 .. size-2
    for (int j = 1; j < js - 1; ++j)
      for (int i = 1; i < is - 1; ++i) elec_cell<false>(P, Q, U, EL, i, j, k);
  for (int k = gw - 1; k <= ks - gw; ++k)   // k_trace: gw-1 .. size-gw
    for (int j = gw - 1; j <= js - gw; ++j)
      for (int i = gw - 1; i <= is - gw; ++i) {
        if (P.slope_type == 3.0) trace_cell<false, true>(P, Q, U, EL, W, i, j, k, dt);   // k_trace<.., S3 = true>
        else trace_cell<false>(P, Q, U, EL, W, i, j, k, dt);
      }
  auto arr = [&](int s) { return HostView<double>{out + (size_t)s * ncell * 8, is, js, ks}; };
  auto putState = [&](const HostView<double>& A, int i, int j, int k, int dir, const dev::State<double>& s) {
    A(ID, i, j, k) = s.r; A(IP, i, j, k) = s.p;
    if (dir == 0) { A(IU, i, j, k) = s.u; A(IV, i, j, k) = s.v; A(IW, i, j, k) = s.w; A(IA, i, j, k) = s.a; A(IB, i, j, k) = s.b; A(IC, i, j, k) = s.c; }
    else if (dir == 1) { A(IU, i, j, k) = s.v; A(IV, i, j, k) = s.u; A(IW, i, j, k) = s.w; A(IA, i, j, k) = s.b; A(IB, i, j, k) = s.a; A(IC, i, j, k) = s.c; }
    else { A(IU, i, j, k) = s.w; A(IV, i, j, k) = s.v; A(IW, i, j, k) = s.u; A(IA, i, j, k) = s.c; A(IB, i, j, k) = s.b; A(IC, i, j, k) = s.a; }
  };
  auto putCorner = [&](const HostView<double>& A, int i, int j, int k, int edir, const dev::Corner<double>& c) {
    A(ID, i, j, k) = c.r; A(IP, i, j, k) = c.p;
    if (edir == 2) { A(IU, i, j, k) = c.u; A(IV, i, j, k) = c.v; A(IW, i, j, k) = c.w; A(IA, i, j, k) = c.a; A(IB, i, j, k) = c.b; A(IC, i, j, k) = c.c; }
    else if (edir == 1) { A(IW, i, j, k) = c.u; A(IU, i, j, k) = c.v; A(IV, i, j, k) = c.w; A(IC, i, j, k) = c.a; A(IA, i, j, k) = c.b; A(IB, i, j, k) = c.c; }
    else { A(IV, i, j, k) = c.u; A(IW, i, j, k) = c.v; A(IU, i, j, k) = c.w; A(IB, i, j, k) = c.a; A(IC, i, j, k) = c.b; A(IA, i, j, k) = c.c; }
  };
  const int hi[3] = {is - gw, js - gw, ks - gw};  // last traced cell per direction
  for (int k = gw - 1; k <= hi[2]; ++k)
    for (int j = gw - 1; j <= hi[1]; ++j)
      for (int i = gw - 1; i <= hi[0]; ++i) {
        const int c[3] = {i, j, k};
        // face states: qp (low face) always; qm (high face) reads the low-face field of the +1 neighbour
        putState(arr(3), i, j, k, 0, face_state<double, 0>(P, W, i, j, k, -1.0));
        putState(arr(4), i, j, k, 1, face_state<double, 1>(P, W, i, j, k, -1.0));
        putState(arr(5), i, j, k, 2, face_state<double, 2>(P, W, i, j, k, -1.0));
        if (c[0] < hi[0]) putState(arr(0), i, j, k, 0, face_state<double, 0>(P, W, i, j, k, 1.0));
        if (c[1] < hi[1]) putState(arr(1), i, j, k, 1, face_state<double, 1>(P, W, i, j, k, 1.0));
        if (c[2] < hi[2]) putState(arr(2), i, j, k, 2, face_state<double, 2>(P, W, i, j, k, 1.0));
        // edge states: e = RT (+,+), RB (+,-), LT (-,+), LB (-,-) along (d1, d2) = (y,z) for x edges, (x,z) for y, (x,y) for z
        const double sg[4][2] = {{1, 1}, {1, -1}, {-1, 1}, {-1, -1}};
        for (int e = 0; e < 4; ++e) {
          const double s1 = sg[e][0], s2 = sg[e][1];
          { const bool ok = (s1 < 0 || c[1] < hi[1]) && (s2 < 0 || c[2] < hi[2]);
            if (ok) putCorner(arr(6 + 3 * e + 0), i, j, k, 0, edge_state<double, 0>(P, W, i, j, k, s1, s2)); }
          { const bool ok = (s1 < 0 || c[0] < hi[0]) && (s2 < 0 || c[2] < hi[2]);
            if (ok) putCorner(arr(6 + 3 * e + 1), i, j, k, 1, edge_state<double, 1>(P, W, i, j, k, s1, s2)); }
          { const bool ok = (s1 < 0 || c[0] < hi[0]) && (s2 < 0 || c[1] < hi[1]);
            if (ok) putCorner(arr(6 + 3 * e + 2), i, j, k, 2, edge_state<double, 2>(P, W, i, j, k, s1, s2)); }
        }
      }
}

// One whole step of the FAST configuration (adiabatic, non-rotating, HLLD + 2-D HLLD: the headline workload) from a
// ghost-filled state, with the product's per-cell functions in the order and over the index ranges of its kernels
// (k_prim, k_elec, k_trace, flux / emf tasks, update_cell with the constrained-transport update and the next
// inverse dt).  Unew must hold a copy of Uold on entry (the kernels leave the cells outside the update box alone).
// Returns the maximum inverse dt of the new state.
double emu_mhd3d_step_fast(const char* ini, const double* Uin, double dt, double* Unew) {
  const KParams<double> P = paramsOf<double>(ini);
  const int is = P.isize, js = P.jsize, ks = P.ksize, gw = P.gw;
  const int iN = is - gw, jN = js - gw, kN = ks - gw;
  const size_t ncell = (size_t)is * js * ks;
  std::vector<double> Qs(ncell * 8, 0.0), ELs(ncell * 3, 0.0), Ws(ncell * NW_MHD, 0.0), Fs(ncell * 15, 0.0), Es(ncell * 3, 0.0);
  const HostView<const double> U{Uin, is, js, ks};
  const HostView<double> Q{Qs.data(), is, js, ks}, EL{ELs.data(), is, js, ks}, W{Ws.data(), is, js, ks};
  const HostView<double> F{Fs.data(), is, js, ks}, E{Es.data(), is, js, ks};
  for (int k = 0; k < ks - 1; ++k)
    for (int j = 0; j < js - 1; ++j)
      for (int i = 0; i < is - 1; ++i) {
        double u[8], q[8];
        for (int v = 0; v < 8; ++v) u[v] = U(v, i, j, k);
        dev::cons_to_prim_mhd<true>(P, u, U(IA, i + 1, j, k), U(IB, i, j + 1, k), U(IC, i, j, k + 1), dt, q);
        for (int v = 0; v < 8; ++v) Q(v, i, j, k) = q[v];
      }
  for (int k = 1; k < ks - 1; ++k)
    for (int j = 1; j < js - 1; ++j)
      for (int i = 1; i < is - 1; ++i) elec_cell<true>(P, Q, U, EL, i, j, k);
  for (int k = gw - 1; k <= kN; ++k)
    for (int j = gw - 1; j <= jN; ++j)
      for (int i = gw - 1; i <= iN; ++i) trace_cell<true>(P, Q, U, EL, W, i, j, k, dt);
  for (int k = gw; k <= kN; ++k)
    for (int j = gw; j <= jN; ++j)
      for (int i = gw; i <= iN; ++i) {
        for (int dir = 0; dir < 3; ++dir) {  // a face is only needed where both transverse indexes are inner (k_flux)
          if (dir != 0 && i >= iN) continue;
          if (dir != 1 && j >= jN) continue;
          if (dir != 2 && k >= kN) continue;
          fused_flux_task(P, W, F, dir, i, j, k);
        }
        for (int edir = 0; edir < 3; ++edir) fused_emf_task(P, W, E, edir, i, j, k);
      }
  double invDt = 0.0;
  for (int k = gw; k <= kN; ++k)
    for (int j = gw; j <= jN; ++j)
      for (int i = gw; i <= iN; ++i) {
        const double v = update_cell<true>(P, U, Unew, F, E, i, j, k, dt);
        if (v > invDt) invDt = v;
      }
  return invDt;
}

}  // extern "C"

// The same for the GENERIC path (any Riemann / emf solver, isothermal, 27-point slopes, static gravity; Omega0 = 0):
// k_prim, k_elec, k_trace<false>, k_flux<false> x3, k_emf<false> x3, k_update<false>.
template <int DIR>
static void fluxAll(const KParams<double>& P, const HostView<double>& W, const HostView<double>& F) {
  const int gw = P.gw, iN = P.isize - gw, jN = P.jsize - gw, kN = P.ksize - gw;
  for (int k = gw; k <= kN; ++k)
    for (int j = gw; j <= jN; ++j)
      for (int i = gw; i <= iN; ++i) {
        if (DIR != 0 && i >= iN) continue;
        if (DIR != 1 && j >= jN) continue;
        if (DIR != 2 && k >= kN) continue;
        flux_cell<double, DIR, false>(P, W, F, i, j, k);
      }
}
template <int EDIR>
static void emfAll(const KParams<double>& P, const HostView<double>& W, const HostView<double>& E) {
  const int gw = P.gw, iN = P.isize - gw, jN = P.jsize - gw, kN = P.ksize - gw;
  for (int k = gw; k <= kN; ++k)
    for (int j = gw; j <= jN; ++j)
      for (int i = gw; i <= iN; ++i) emf_cell<double, EDIR, false>(P, W, E, i, j, k);
}

double emu_mhd3d_step_generic_impl(const char* ini, const double* Uin, double dt, double* Unew) {
  const KParams<double> P = paramsOf<double>(ini);
  const int is = P.isize, js = P.jsize, ks = P.ksize, gw = P.gw;
  const int iN = is - gw, jN = js - gw, kN = ks - gw;
  const size_t ncell = (size_t)is * js * ks;
  std::vector<double> Qs(ncell * 8, 0.0), ELs(ncell * 3, 0.0), Ws(ncell * NW_MHD, 0.0), Fs(ncell * 15, 0.0), Es(ncell * 3, 0.0);
  const HostView<const double> U{Uin, is, js, ks};
  const HostView<double> Q{Qs.data(), is, js, ks}, EL{ELs.data(), is, js, ks}, W{Ws.data(), is, js, ks};
  const HostView<double> F{Fs.data(), is, js, ks}, E{Es.data(), is, js, ks};
  for (int k = 0; k < ks - 1; ++k)
    for (int j = 0; j < js - 1; ++j)
      for (int i = 0; i < is - 1; ++i) {
        double u[8], q[8];
        for (int v = 0; v < 8; ++v) u[v] = U(v, i, j, k);
        dev::cons_to_prim_mhd(P, u, U(IA, i + 1, j, k), U(IB, i, j + 1, k), U(IC, i, j, k + 1), dt, q);
        for (int v = 0; v < 8; ++v) Q(v, i, j, k) = q[v];
      }
  for (int k = 1; k < ks - 1; ++k)
    for (int j = 1; j < js - 1; ++j)
      for (int i = 1; i < is - 1; ++i) elec_cell<false>(P, Q, U, EL, i, j, k);
  for (int k = gw - 1; k <= kN; ++k)
    for (int j = gw - 1; j <= jN; ++j)
      for (int i = gw - 1; i <= iN; ++i) {
        if (P.slope_type == 3.0) trace_cell<false, true>(P, Q, U, EL, W, i, j, k, dt);
        else trace_cell<false>(P, Q, U, EL, W, i, j, k, dt);
      }
  fluxAll<0>(P, W, F); fluxAll<1>(P, W, F); fluxAll<2>(P, W, F);
  emfAll<0>(P, W, E); emfAll<1>(P, W, E); emfAll<2>(P, W, E);
  double invDt = 0.0;
  for (int k = gw; k <= kN; ++k)
    for (int j = gw; j <= jN; ++j)
      for (int i = gw; i <= iN; ++i) {
        const double v = update_cell<false>(P, U, Unew, F, E, i, j, k, dt);
        if (v > invDt) invDt = v;
      }
  return invDt;
}


extern "C" double emu_mhd3d_step_generic(const char* ini, const double* Uin, double dt, double* Unew) {
  return emu_mhd3d_step_generic_impl(ini, Uin, dt, Unew);
}

// One whole 3D hydro step (FP64 or FP32) from a ghost-filled state with the product's per-cell functions
// (hydro_cells.cuh: trace -> W, face states, Riemann fluxes) and the update of k_hydro_flux_update (reference
// summation order, gravity source term, inverse dt of the new state).  Unew holds a copy of Uold on entry.
template <typename T>
static double hydroStep(const char* ini, const T* Uin, T dt, T* Unew) {
  const KParams<T> P = paramsOf<T>(ini);
  const int is = P.isize, js = P.jsize, ks = P.ksize, gw = P.gw;
  const size_t ncell = (size_t)is * js * ks;
  std::vector<T> Ws(ncell * NW_HYDRO, T(0));
  const HostView<const T> U{Uin, is, js, ks};
  const HostView<T> W{Ws.data(), is, js, ks};
  for (int k = 1; k < ks - 1; ++k)
    for (int j = 1; j < js - 1; ++j)
      for (int i = 1; i < is - 1; ++i) hydro_trace_cell(P, U, W, i, j, k, dt);
  const T dtdx = dt / P.dx, dtdy = dt / P.dy, dtdz = dt / P.dz;
  double invDt = 0.0;
  for (int k = gw; k < ks - gw; ++k)
    for (int j = gw; j < js - gw; ++j)
      for (int i = gw; i < is - gw; ++i) {
        T fxl[5], fyl[5], fzl[5], fxh[5], fyh[5], fzh[5], un[5];
        hydro_low_flux<T, 0, -1>(P, W, i, j, k, fxl); hydro_low_flux<T, 0, -1>(P, W, i + 1, j, k, fxh);
        hydro_low_flux<T, 1, -1>(P, W, i, j, k, fyl); hydro_low_flux<T, 1, -1>(P, W, i, j + 1, k, fyh);
        hydro_low_flux<T, 2, -1>(P, W, i, j, k, fzl); hydro_low_flux<T, 2, -1>(P, W, i, j, k + 1, fzh);
        for (int v = 0; v < 5; ++v) {
          T s = U(v, i, j, k);
          s += fxl[v] * dtdx; s += fyl[v] * dtdy; s += fzl[v] * dtdz;
          s -= fxh[v] * dtdx; s -= fyh[v] * dtdy; s -= fzh[v] * dtdz;
          un[v] = s;
        }
        if (P.gravity) {
          const T hdt = T(0.5) * dt, rs = U(ID, i, j, k) + un[ID];
          un[IU] += hdt * P.gx * rs; un[IV] += hdt * P.gy * rs; un[IW] += hdt * P.gz * rs;
        }
        for (int v = 0; v < 5; ++v) Unew[((size_t)v * ks + k) * js * is + (size_t)j * is + i] = un[v];
        T q[5];
        const T c = dev::cons_to_prim_hydro(P, un[ID], un[IP], un[IU], un[IV], un[IW], q);
        const double d = (c + dev::ab(q[IU])) / P.dx + (c + dev::ab(q[IV])) / P.dy + (c + dev::ab(q[IW])) / P.dz;
        if (d > invDt) invDt = d;
      }
  return invDt;
}
extern "C" double emu_hydro3d_step(const char* ini, const double* U, double dt, double* Unew) { return hydroStep<double>(ini, U, dt, Unew); }
extern "C" double emu_hydro3d_step_f32(const char* ini, const float* U, float dt, float* Unew) { return hydroStep<float>(ini, U, dt, Unew); }
