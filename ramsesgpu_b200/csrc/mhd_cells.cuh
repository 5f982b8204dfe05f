// Per-cell building blocks of the 3D MHD step, generic in the way the arrays are reached: Q(v,i,j,k) primitives,
// U(v,i,j,k) conservative state (face fields), EL(c,i,j,k) edge electric fields, W(c,i,j,k) the traced state.
// The kernels of kernels_mhd3d.cu instantiate them on global-memory views and on shared-memory tiles; the CPU test
// suite instantiates the same source on host arrays (tests/host_emul) and compares it with the oracle.
#pragma once
#include "kernels.h"
#include "mhd_device.cuh"

namespace rg {

namespace {

// W component ids
enum {
  W_R = 0, W_P, W_U, W_V, W_W, W_A, W_B, W_C,            // cell centred, advanced by dt/2
  W_AL, W_BL, W_CL,                                      // LOW-face fields, advanced by dt/2
  W_DRX, W_DPX, W_DUX, W_DVX, W_DWX, W_DBX, W_DCX,       // half slopes along x
  W_DRY, W_DPY, W_DUY, W_DVY, W_DWY, W_DAY, W_DCY,       // half slopes along y
  W_DRZ, W_DPZ, W_DUZ, W_DVZ, W_DWZ, W_DAZ, W_DBZ,       // half slopes along z
  W_DALY, W_DALZ, W_DBLX, W_DBLZ, W_DCLX, W_DCLY         // half slopes of the low-face fields
};
// The HIGH-face field of a cell and its slopes are, bit for bit, the LOW-face values of the +1
// neighbour (same face, same edge electric fields, same limiter inputs), so consumers read them
// there and W carries 38 components instead of 47.
static_assert(W_DCLY + 1 == NW_MHD, "W layout");

// vertical gravity of local plane k: the per-plane field of the stratified shearing box, else the uniform field
template <typename T>
__device__ __forceinline__ T grav_z(const KParams<T>& P, int k) {
  return (P.gzPlane != nullptr) ? P.gzPlane[k] : P.gz;
}

// ------------------------------------------------------------------------------------------------
// edge-centred electric field E = v x B at the LOW edges of a cell (reference cpu_v3.cpp:36-101)
// ------------------------------------------------------------------------------------------------
// edge electric fields of one cell (low edges); QV / UV / ELV as in trace_cell
template <bool FAST, typename T, typename QV, typename UV, typename ELV>
__device__ __forceinline__ void elec_cell(const KParams<T>& P, const QV& Q, const UV& U, const ELV& EL, int i, int j,
                                          int k) {
  const T h = T(0.5), f = T(0.25);
  const T u00 = Q(IU, i, j, k), v00 = Q(IV, i, j, k), w00 = Q(IW, i, j, k);
  const T A = U(IA, i, j, k), B = U(IB, i, j, k), C = U(IC, i, j, k);
  {  // Ex: average over (j-1..j, k-1..k)
    const T v = f * (Q(IV, i, j - 1, k - 1) + Q(IV, i, j - 1, k) + Q(IV, i, j, k - 1) + v00);
    const T w = f * (Q(IW, i, j - 1, k - 1) + Q(IW, i, j - 1, k) + Q(IW, i, j, k - 1) + w00);
    const T Bm = h * (U(IB, i, j, k - 1) + B), Cm = h * (U(IC, i, j - 1, k) + C);
    T ex = v * Cm - w * Bm;
    if (!FAST && P.Omega0 > T(0)) {  // rotating frame: advection by the background shear, MHDRunGodunov.cpp:2474-2478
      const T xPos = P.xMin + P.dx * h + (i - P.gw) * P.dx;
      ex += T(-1.5) * P.Omega0 * xPos * Cm;
    }
    EL(0, i, j, k) = ex;
  }
  {  // Ey: average over (i-1..i, k-1..k)
    const T u = f * (Q(IU, i - 1, j, k - 1) + Q(IU, i - 1, j, k) + Q(IU, i, j, k - 1) + u00);
    const T w = f * (Q(IW, i - 1, j, k - 1) + Q(IW, i - 1, j, k) + Q(IW, i, j, k - 1) + w00);
    const T Am = h * (U(IA, i, j, k - 1) + A), Cm = h * (U(IC, i - 1, j, k) + C);
    EL(1, i, j, k) = w * Am - u * Cm;
  }
  {  // Ez: average over (i-1..i, j-1..j)
    const T u = f * (Q(IU, i - 1, j - 1, k) + Q(IU, i - 1, j, k) + Q(IU, i, j - 1, k) + u00);
    const T v = f * (Q(IV, i - 1, j - 1, k) + Q(IV, i - 1, j, k) + Q(IV, i, j - 1, k) + v00);
    const T Am = h * (U(IA, i, j - 1, k) + A), Bm = h * (U(IB, i - 1, j, k) + B);
    T ez = u * Bm - v * Am;
    if (!FAST && P.Omega0 > T(0)) {  // MHDRunGodunov.cpp:2517-2521 (shear at the x face)
      const T xFace = P.xMin + P.dx * h + (i - P.gw) * P.dx - P.dx * h;
      ez -= T(-1.5) * P.Omega0 * xFace * Am;
    }
    EL(2, i, j, k) = ez;
  }
}

// ------------------------------------------------------------------------------------------------
// slopes + face-B slopes + half-step trace of one cell -> W (reference cpu_v3.cpp:36-361, slope_mhd.h, trace_mhd.h)
// ------------------------------------------------------------------------------------------------
// slope_type 3 (reference slope_mhd.h:352-409): the central differences of a variable along x, y, z are scaled by ONE
// positivity-preserving factor, min(1, min(|vmin|, |vmax|) / (|dfx| + |dfy| + |dfz|) * 2), taken over the 27-cell
// neighbourhood.  Returns that factor for variable v of cell (i,j,k).
template <typename T, typename QV>
__device__ __forceinline__ T slope27_factor(const QV& Q, int v, int i, int j, int k) {
  const T q0 = Q(v, i, j, k);
  T vmin = T(0), vmax = T(0);  // the centre difference (0) takes part in both
#pragma unroll
  for (int dk = -1; dk <= 1; ++dk)
#pragma unroll
    for (int dj = -1; dj <= 1; ++dj)
#pragma unroll
      for (int di = -1; di <= 1; ++di) {
        const T d = Q(v, i + di, j + dj, k + dk) - q0;
        vmin = dev::mn(vmin, d);
        vmax = dev::mx(vmax, d);
      }
  const T dfx = T(0.5) * (Q(v, i + 1, j, k) - Q(v, i - 1, j, k));
  const T dfy = T(0.5) * (Q(v, i, j + 1, k) - Q(v, i, j - 1, k));
  const T dfz = T(0.5) * (Q(v, i, j, k + 1) - Q(v, i, j, k - 1));
  const T dff = T(0.5) * (dev::ab(dfx) + dev::ab(dfy) + dev::ab(dfz));
  return (dff > T(0)) ? dev::mn(T(1), dev::mn(dev::ab(vmin), dev::ab(vmax)) / dff) : T(1);
}

// trace of one cell: QV(v,i,j,k) primitives, UV(v,i,j,k) conservative state (face fields), ELV(c,i,j,k)
// edge electric fields, WV(c,i,j,k) the traced state (written).  Used by k_trace (global arrays) and by
// the fused prim+elec+trace kernel (shared-memory tiles).
template <bool FAST, bool S3 = false, typename T, typename QV, typename UV, typename ELV, typename WV>
__device__ __forceinline__ void trace_cell(const KParams<T>& P, const QV& Q, const UV& U, const ELV& EL, const WV& W,
                                           int i, int j, int k, T dt) {
  const int gw = P.gw;
  const T h = T(0.5);
  const T hst = h * P.slope_type;  // slope_type 0 gives zero slopes through hst = 0
  const T dtdx = dt / P.dx, dtdy = dt / P.dy, dtdz = dt / P.dz;
  // HALF slope of primitive variable v between its -1 / +1 neighbours along one direction: TVD limiter (types 1, 2),
  // or the centred difference times the 27-point factor (S3 = slope_type 3: its own instantiation, so that the
  // kernels of the other slope types are untouched)
  auto hslope = [&](int v, T qm, T q0, T qp) -> T {
    if (S3) return h * (slope27_factor<T>(Q, v, i, j, k) * (h * (qp - qm)));
    return dev::half_slope(hst, qm, q0, qp);
  };
  // The work is arranged direction by direction (slopes of one direction -> stored -> their share of
  // the half-step source terms accumulated) so that few values are live at any time.

  // face fields: transverse HALF slopes (slope type capped at 2, slope_mhd.h:636), induction by the 12
  // edge electric fields of the cell (trace_mhd.h:2006-2011)
  const T hxst = h * dev::mn(P.slope_type, T(2));
  const T AL = U(IA, i, j, k), BL = U(IB, i, j, k), CL = U(IC, i, j, k);
  const T dAx = h * (U(IA, i + 1, j, k) - AL), dBy = h * (U(IB, i, j + 1, k) - BL), dCz = h * (U(IC, i, j, k + 1) - CL);
  W(W_DALY, i, j, k) = dev::half_slope(hxst, U(IA, i, j - 1, k), AL, U(IA, i, j + 1, k));
  W(W_DALZ, i, j, k) = dev::half_slope(hxst, U(IA, i, j, k - 1), AL, U(IA, i, j, k + 1));
  W(W_DBLX, i, j, k) = dev::half_slope(hxst, U(IB, i - 1, j, k), BL, U(IB, i + 1, j, k));
  W(W_DBLZ, i, j, k) = dev::half_slope(hxst, U(IB, i, j, k - 1), BL, U(IB, i, j, k + 1));
  W(W_DCLX, i, j, k) = dev::half_slope(hxst, U(IC, i - 1, j, k), CL, U(IC, i + 1, j, k));
  W(W_DCLY, i, j, k) = dev::half_slope(hxst, U(IC, i, j - 1, k), CL, U(IC, i, j + 1, k));
  {
    const T ELL = EL(0, i, j, k), ELR = EL(0, i, j, k + 1), ERL = EL(0, i, j + 1, k);
    const T FLL = EL(1, i, j, k), FLR = EL(1, i, j, k + 1), FRL = EL(1, i + 1, j, k);
    const T GLL = EL(2, i, j, k), GLR = EL(2, i, j + 1, k), GRL = EL(2, i + 1, j, k);
    W(W_AL, i, j, k) = AL + ((GLR - GLL) * dtdy * h - (FLR - FLL) * dtdz * h);
    W(W_BL, i, j, k) = BL + (-(GRL - GLL) * dtdx * h + (ELR - ELL) * dtdz * h);
    W(W_CL, i, j, k) = CL + ((FRL - FLL) * dtdx * h - (ERL - ELL) * dtdy * h);
  }

  // cell-centred state; half-step source terms (trace_mhd.h:1985-2011) accumulated per direction
  const T r = Q(ID, i, j, k), p = Q(IP, i, j, k), u = Q(IU, i, j, k), v = Q(IV, i, j, k), w = Q(IW, i, j, k);
  const T A = Q(IA, i, j, k), B = Q(IB, i, j, k), C = Q(IC, i, j, k);
  const T ir = dev::rcp(r);
  const T gp = P.gamma0 * p;
  T sr0, su0, sv0, sw0, sp0, sA0, sB0, sC0;
  {  // x
    const T drx = hslope(ID, Q(ID, i - 1, j, k), r, Q(ID, i + 1, j, k));
    const T dpx = hslope(IP, Q(IP, i - 1, j, k), p, Q(IP, i + 1, j, k));
    const T dux = hslope(IU, Q(IU, i - 1, j, k), u, Q(IU, i + 1, j, k));
    const T dvx = hslope(IV, Q(IV, i - 1, j, k), v, Q(IV, i + 1, j, k));
    const T dwx = hslope(IW, Q(IW, i - 1, j, k), w, Q(IW, i + 1, j, k));
    const T dBx = hslope(IB, Q(IB, i - 1, j, k), B, Q(IB, i + 1, j, k));
    const T dCx = hslope(IC, Q(IC, i - 1, j, k), C, Q(IC, i + 1, j, k));
    W(W_DRX, i, j, k) = drx; W(W_DPX, i, j, k) = dpx; W(W_DUX, i, j, k) = dux; W(W_DVX, i, j, k) = dvx;
    W(W_DWX, i, j, k) = dwx; W(W_DBX, i, j, k) = dBx; W(W_DCX, i, j, k) = dCx;
    sr0 = (-u * drx - dux * r) * dtdx;
    su0 = (-u * dux - (dpx + B * dBx + C * dCx) * ir) * dtdx;
    sv0 = (-u * dvx + A * dBx * ir) * dtdx;
    sw0 = (-u * dwx + A * dCx * ir) * dtdx;
    sp0 = (-u * dpx - dux * gp) * dtdx;
    sB0 = (v * dAx + A * dvx - u * dBx - B * dux) * dtdx;
    sC0 = (w * dAx + A * dwx - u * dCx - C * dux) * dtdx;
  }
  T shr = T(0), shu = T(0), shv = T(0), shw = T(0), shp = T(0), shA = T(0), shC = T(0);  // shearing box only
  {  // y
    const T dry = hslope(ID, Q(ID, i, j - 1, k), r, Q(ID, i, j + 1, k));
    const T dpy = hslope(IP, Q(IP, i, j - 1, k), p, Q(IP, i, j + 1, k));
    const T duy = hslope(IU, Q(IU, i, j - 1, k), u, Q(IU, i, j + 1, k));
    const T dvy = hslope(IV, Q(IV, i, j - 1, k), v, Q(IV, i, j + 1, k));
    const T dwy = hslope(IW, Q(IW, i, j - 1, k), w, Q(IW, i, j + 1, k));
    const T dAy = hslope(IA, Q(IA, i, j - 1, k), A, Q(IA, i, j + 1, k));
    const T dCy = hslope(IC, Q(IC, i, j - 1, k), C, Q(IC, i, j + 1, k));
    W(W_DRY, i, j, k) = dry; W(W_DPY, i, j, k) = dpy; W(W_DUY, i, j, k) = duy; W(W_DVY, i, j, k) = dvy;
    W(W_DWY, i, j, k) = dwy; W(W_DAY, i, j, k) = dAy; W(W_DCY, i, j, k) = dCy;
    sr0 += (-v * dry - dvy * r) * dtdy;
    su0 += (-v * duy + B * dAy * ir) * dtdy;
    sv0 += (-v * dvy - (dpy + A * dAy + C * dCy) * ir) * dtdy;
    sw0 += (-v * dwy + B * dCy * ir) * dtdy;
    sp0 += (-v * dpy - dvy * gp) * dtdy;
    sA0 = (u * dBy + B * duy - v * dAy - A * dvy) * dtdy;
    sC0 += (w * dBy + B * dwy - v * dCy - C * dvy) * dtdy;
    if (!FAST && P.Omega0 > T(0)) {  // shearing-box terms, trace_mhd.h:1993-2003 (applied below)
      const T xPos = P.xMin + P.dx * h + (i - gw) * P.dx;
      const T shear = T(-1.5) * P.Omega0 * xPos;
      shr = shear * dry * dtdy; shu = shear * duy * dtdy; shv = shear * dvy * dtdy; shw = shear * dwy * dtdy;
      shp = shear * dpy * dtdy; shA = shear * dAy * dtdy; shC = shear * dCy * dtdy;
    }
  }
  {  // z
    const T drz = hslope(ID, Q(ID, i, j, k - 1), r, Q(ID, i, j, k + 1));
    const T dpz = hslope(IP, Q(IP, i, j, k - 1), p, Q(IP, i, j, k + 1));
    const T duz = hslope(IU, Q(IU, i, j, k - 1), u, Q(IU, i, j, k + 1));
    const T dvz = hslope(IV, Q(IV, i, j, k - 1), v, Q(IV, i, j, k + 1));
    const T dwz = hslope(IW, Q(IW, i, j, k - 1), w, Q(IW, i, j, k + 1));
    const T dAz = hslope(IA, Q(IA, i, j, k - 1), A, Q(IA, i, j, k + 1));
    const T dBz = hslope(IB, Q(IB, i, j, k - 1), B, Q(IB, i, j, k + 1));
    W(W_DRZ, i, j, k) = drz; W(W_DPZ, i, j, k) = dpz; W(W_DUZ, i, j, k) = duz; W(W_DVZ, i, j, k) = dvz;
    W(W_DWZ, i, j, k) = dwz; W(W_DAZ, i, j, k) = dAz; W(W_DBZ, i, j, k) = dBz;
    sr0 += (-w * drz - dwz * r) * dtdz;
    su0 += (-w * duz + C * dAz * ir) * dtdz;
    sv0 += (-w * dvz + C * dBz * ir) * dtdz;
    sw0 += (-w * dwz - (dpz + A * dAz + B * dBz) * ir) * dtdz;
    sp0 += (-w * dpz - dwz * gp) * dtdz;
    sA0 += (u * dCz + C * duz - w * dAz - A * dwz) * dtdz;
    sB0 += (v * dCz + C * dvz - w * dBz - B * dwz) * dtdz;
    if (!FAST && P.Omega0 > T(0)) {
      const T xPos = P.xMin + P.dx * h + (i - gw) * P.dx;
      const T shear = T(-1.5) * P.Omega0 * xPos;
      sr0 -= shr; su0 -= shu; sv0 -= shv; sw0 -= shw; sp0 -= shp; sA0 -= shA;
      sB0 += (shear * dAx - T(1.5) * P.Omega0 * A * P.dx) * dtdx + shear * dBz * dtdz;
      sC0 -= shC;
    }
  }
  if (P.gravity) {  // gravity predictor on every traced velocity of the cell (reference cpu_v3.cpp:277-332):
    const T hdt = h * dt;  // face / edge states are centre +/- slopes, so it goes into the centre value
    su0 += hdt * P.gx; sv0 += hdt * P.gy; sw0 += hdt * grav_z(P, k);
  }
  W(W_R, i, j, k) = r + sr0;  W(W_P, i, j, k) = p + sp0;
  W(W_U, i, j, k) = u + su0;  W(W_V, i, j, k) = v + sv0;  W(W_W, i, j, k) = w + sw0;
  W(W_A, i, j, k) = A + sA0;  W(W_B, i, j, k) = B + sB0;  W(W_C, i, j, k) = C + sC0;
}

// ------------------------------------------------------------------------------------------------
// face state = W cell-centred value +/- half slope along the normal, floors on rho and p (trace_mhd.h:2032-2102)
// ------------------------------------------------------------------------------------------------
template <typename T, int DIR, typename WV>
__device__ __forceinline__ dev::State<T> face_state(const KParams<T>& P, const WV& W, int i, int j, int k, T sgn) {
  // sgn = +1 : state at the HIGH face of cell (qm), -1 : at the LOW face (qp)
  constexpr int S = (DIR == 0) ? W_DRX : (DIR == 1) ? W_DRY : W_DRZ;  // first slope component
  dev::State<T> s;
  s.r = dev::mx(P.smallr, W(W_R, i, j, k) + sgn * W(S + 0, i, j, k));
  s.p = dev::mx(P.smallp, W(W_P, i, j, k) + sgn * W(S + 1, i, j, k));
  const T u = W(W_U, i, j, k) + sgn * W(S + 2, i, j, k);
  const T v = W(W_V, i, j, k) + sgn * W(S + 3, i, j, k);
  const T w = W(W_W, i, j, k) + sgn * W(S + 4, i, j, k);
  if (DIR == 0) {
    s.u = u; s.v = v; s.w = w;
    s.a = W(W_AL, (sgn > T(0)) ? i + 1 : i, j, k);
    s.b = W(W_B, i, j, k) + sgn * W(W_DBX, i, j, k);
    s.c = W(W_C, i, j, k) + sgn * W(W_DCX, i, j, k);
  } else if (DIR == 1) {  // swap (u,v) and (a,b)
    s.u = v; s.v = u; s.w = w;
    s.a = W(W_BL, i, (sgn > T(0)) ? j + 1 : j, k);
    s.b = W(W_A, i, j, k) + sgn * W(W_DAY, i, j, k);
    s.c = W(W_C, i, j, k) + sgn * W(W_DCY, i, j, k);
  } else {  // swap (u,w) and (a,c)
    s.u = w; s.v = v; s.w = u;
    s.a = W(W_CL, i, j, (sgn > T(0)) ? k + 1 : k);
    s.b = W(W_B, i, j, k) + sgn * W(W_DBZ, i, j, k);
    s.c = W(W_A, i, j, k) + sgn * W(W_DAZ, i, j, k);
  }
  return s;
}

// ------------------------------------------------------------------------------------------------
// edge states for the corner emfs (trace_mhd.h:2104-2246)
// ------------------------------------------------------------------------------------------------
// edge state of cell (i,j,k) for edge direction EDIR, signs (s1, s2) along the two transverse
// directions (d1,d2) = (x,y) for Z, (x,z) for Y, (y,z) for X; returned in the EDGE frame
//   Z: u<-U v<-V w<-W a<-A b<-B c<-C ; Y: u<-W v<-U w<-V a<-C b<-A c<-B ; X: u<-V v<-W w<-U a<-B b<-C c<-A
template <typename T, int EDIR, typename WV>
__device__ __forceinline__ dev::Corner<T> edge_state(const KParams<T>& P, const WV& W, int i, int j, int k, T s1,
                                                     T s2) {
  constexpr int S1 = (EDIR == 0) ? W_DRY : W_DRX;  // slopes along d1
  constexpr int S2 = (EDIR == 2) ? W_DRY : W_DRZ;  // slopes along d2
  const T r = dev::mx(P.smallr, W(W_R, i, j, k) + (s1 * W(S1 + 0, i, j, k) + s2 * W(S2 + 0, i, j, k)));
  const T p = dev::mx(P.smallp, W(W_P, i, j, k) + (s1 * W(S1 + 1, i, j, k) + s2 * W(S2 + 1, i, j, k)));
  const T U = W(W_U, i, j, k) + (s1 * W(S1 + 2, i, j, k) + s2 * W(S2 + 2, i, j, k));
  const T V = W(W_V, i, j, k) + (s1 * W(S1 + 3, i, j, k) + s2 * W(S2 + 3, i, j, k));
  const T Wv = W(W_W, i, j, k) + (s1 * W(S1 + 4, i, j, k) + s2 * W(S2 + 4, i, j, k));
  T A, B, C;
  if (EDIR == 2) {  // (x,y): A from the x-face on side s1 with its y slope, B from the y-face on side s2 with its x slope
    { const int ia = (s1 > T(0)) ? i + 1 : i; A = W(W_AL, ia, j, k) + s2 * W(W_DALY, ia, j, k); }
    { const int jb = (s2 > T(0)) ? j + 1 : j; B = W(W_BL, i, jb, k) + s1 * W(W_DBLX, i, jb, k); }
    C = W(W_C, i, j, k) + (s1 * W(W_DCX, i, j, k) + s2 * W(W_DCY, i, j, k));
  } else if (EDIR == 1) {  // (x,z)
    { const int ia = (s1 > T(0)) ? i + 1 : i; A = W(W_AL, ia, j, k) + s2 * W(W_DALZ, ia, j, k); }
    B = W(W_B, i, j, k) + (s1 * W(W_DBX, i, j, k) + s2 * W(W_DBZ, i, j, k));
    { const int kc = (s2 > T(0)) ? k + 1 : k; C = W(W_CL, i, j, kc) + s1 * W(W_DCLX, i, j, kc); }
  } else {  // (y,z)
    A = W(W_A, i, j, k) + (s1 * W(W_DAY, i, j, k) + s2 * W(W_DAZ, i, j, k));
    { const int jb = (s1 > T(0)) ? j + 1 : j; B = W(W_BL, i, jb, k) + s2 * W(W_DBLZ, i, jb, k); }
    { const int kc = (s2 > T(0)) ? k + 1 : k; C = W(W_CL, i, j, kc) + s1 * W(W_DCLY, i, j, kc); }
  }
  dev::Corner<T> c;
  c.r = r; c.p = p;
  if (EDIR == 2) { c.u = U; c.v = V; c.w = Wv; c.a = A; c.b = B; c.c = C; }
  else if (EDIR == 1) { c.u = Wv; c.v = U; c.w = V; c.a = C; c.b = A; c.c = B; }
  else { c.u = V; c.v = Wv; c.w = U; c.a = B; c.b = C; c.c = A; }
  return c;
}

// ------------------------------------------------------------------------------------------------
// conservative update + constrained transport + inverse dt of the NEW state for one cell of the update box
// (reference cpu_v3.cpp:475-533, :600-630; dt: MHDRunBase.cpp:141-250)
// ------------------------------------------------------------------------------------------------
// One cell of the update box (gw <= i <= iN etc.).  FV(c, i, j, k) / EV(c, i, j, k) give the face
// fluxes and corner emfs (global scratch arrays or the shared-memory tile of the fused kernel).
// Returns the inverse time step of the NEW state (0 outside the inner cells).
template <bool FAST, typename T, typename UV, typename FV, typename EV>
__device__ __forceinline__ T update_cell(const KParams<T>& P, const UV& U, T* __restrict__ Unew, const FV& F,
                                         const EV& E, int i, int j, int k, T dt) {
  const int gw = P.gw;
  const int iN = P.isize - gw, jN = P.jsize - gw, kN = P.ksize - gw;  // first upper ghost index
  const size_t plane = (size_t)P.isize * P.jsize, comp = plane * P.ksize;
  const size_t idx = (size_t)k * plane + (size_t)j * P.isize + i;
  const T dtdx = dt / P.dx, dtdy = dt / P.dy, dtdz = dt / P.dz;
  const bool inner = i < iN && j < jN && k < kN;
  T invDt = T(0);
  T un[8];
#pragma unroll
  for (int v = 0; v < 8; ++v) un[v] = U(v, i, j, k);
  if (inner) {
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      // same summation order as the reference's serial scatter
      T s = un[v];
      s += F(v, i, j, k) * dtdx;
      s += F(5 + v, i, j, k) * dtdy;
      s += F(10 + v, i, j, k) * dtdz;
      s -= F(v, i + 1, j, k) * dtdx;
      s -= F(5 + v, i, j + 1, k) * dtdy;
      s -= F(10 + v, i, j, k + 1) * dtdz;
      un[v] = s;
    }
    if (P.gravity) {  // static gravity source term on the momenta, reference HydroRunBase.cpp:1962-1976
      const T hdt = T(0.5) * dt, rs = U(ID, i, j, k) + un[ID];
      un[IU] += hdt * P.gx * rs; un[IV] += hdt * P.gy * rs; un[IW] += hdt * grav_z(P, k) * rs;
    }
  }
  // emf(c, ...) with the never-computed indexes (one past the upper ghost face) read as zero,
  // exactly like the reference's zero-initialised h_emf
  auto emf = [&](int c, int ii, int jj, int kk) -> T {
    return (ii > iN || jj > jN || kk > kN) ? T(0) : E(c, ii, jj, kk);
  };
  // constrained transport of one face component (reference cpu_v3.cpp:600-630), per component so
  // that the dt estimate below touches only the emfs it needs
  auto ctx = [&](int ii, int jj, int kk, T bx) -> T {
    if (kk < kN) bx += (emf(0, ii, jj + 1, kk) - emf(0, ii, jj, kk)) * dtdy;
    return bx - (emf(1, ii, jj, kk + 1) - emf(1, ii, jj, kk)) * dtdz;
  };
  auto cty = [&](int ii, int jj, int kk, T by) -> T {
    if (kk < kN) by -= (emf(0, ii + 1, jj, kk) - emf(0, ii, jj, kk)) * dtdx;
    return by + (emf(2, ii, jj, kk + 1) - emf(2, ii, jj, kk)) * dtdz;
  };
  auto ctz = [&](int ii, int jj, int kk, T bz) -> T {
    bz += (emf(1, ii + 1, jj, kk) - emf(1, ii, jj, kk)) * dtdx;
    return bz - (emf(2, ii, jj + 1, kk) - emf(2, ii, jj, kk)) * dtdy;
  };
  un[IA] = ctx(i, j, k, un[IA]);
  un[IB] = cty(i, j, k, un[IB]);
  un[IC] = ctz(i, j, k, un[IC]);
#pragma unroll
  for (int v = 0; v < 8; ++v) Unew[v * comp + idx] = un[v];

  if (inner) {  // inverse dt of the new state: needs the new B on the three upper faces
    const T bxp = ctx(i + 1, j, k, U(IA, i + 1, j, k));
    const T byp = cty(i, j + 1, k, U(IB, i, j + 1, k));
    const T bzp = ctz(i, j, k + 1, U(IC, i, j, k + 1));
    T q[8];
    dev::cons_to_prim_mhd<FAST>(P, un, bxp, byp, bzp, T(0), q);
    const T irho = dev::rcp(q[ID]);
    const T a2 = q[IA] * q[IA], b2 = q[IB] * q[IB], c2 = q[IC] * q[IC];
    const T bb = a2 + b2 + c2;
    T vx = dev::fast_speed(P.gamma0, q[IP], irho, bb, a2) + dev::ab(q[IU]);
    T vy = dev::fast_speed(P.gamma0, q[IP], irho, bb, b2) + dev::ab(q[IV]);
    T vz = dev::fast_speed(P.gamma0, q[IP], irho, bb, c2) + dev::ab(q[IW]);
    if (!FAST && P.Omega0 > T(0)) vy += T(1.5) * P.Omega0 * (P.xMax - P.xMin) * T(0.5);
    invDt = vx * P.rdx + vy * P.rdy + vz * P.rdz;
  }
  return invDt;
}

// ------------------------------------------------------------------------------------------------
// rotating frame / shearing box (Omega0 > 0): the y shift of the opposite x border, the compact strips that hold the
// fluxes / emfs of the four border position columns, and the update of one cell
// ------------------------------------------------------------------------------------------------
template <typename T>
struct ShearShift {   // y shift of the opposite x border: deltay = 1.5 Omega0 Lx t, reference :3213-3216
  int enabled;        // shearing-box boundaries in x
  int jplus;          // whole cells
  T frac;             // epsi / dy
};

template <typename T>
__device__ __forceinline__ void remapRows(const KParams<T>& P, const ShearShift<T>& sh, int j, bool xmin, int& j0,
                                          int& j1, T& eps) {
  const int gw = P.gw, ny = P.ny;
  if (xmin) {  // inner (xmin) border looks at the xmax border shifted by -jplus-1
    j0 = j - sh.jplus - 1; j1 = j0 + 1; eps = T(1) - sh.frac;
    if (j0 < gw) j0 += ny;
    if (j1 < gw) j1 += ny;
  } else {
    j0 = j + sh.jplus; j1 = j0 + 1; eps = sh.frac;
    if (j0 > ny + gw - 1) j0 -= ny;
    if (j1 > ny + gw - 1) j1 -= ny;
  }
}


// Fluxes (c0 = 0) / emfs (c0 = 15) of the x-border position columns i = gw, gw+1, nx+gw-1, nx+gw only, laid out
// [comp][kk][j][4]: written by the fused kernel of the shearing box, read by k_update_rot_border (the cells next to an x
// border read the y-remapped values of the OPPOSITE border, which belong to other tiles)
template <typename T>
struct BorderView {
  T* p;
  int jsize, planes, kbase, gw, nx, c0;
  __device__ __forceinline__ static int slot(int i, int gw, int nx) { return (i <= gw + 1) ? i - gw : 2 + (i - (nx + gw - 1)); }
  __device__ __forceinline__ static bool holds(int i, int gw, int nx) {
    return i == gw || i == gw + 1 || i == nx + gw - 1 || i == nx + gw;
  }
  __device__ __forceinline__ T& operator()(int c, int i, int j, int k) const {
    return p[(((size_t)(c0 + c) * planes + (k - kbase)) * jsize + j) * 4 + slot(i, gw, nx)];
  }
};

// Update of one cell of the update box in the rotating frame (reference MHDRunGodunov.cpp:2938-3348): Crank-Nicolson
// Coriolis rotation of (rho u, rho v), alpha-mixed momentum fluxes, constrained transport, inverse dt of the new
// state; with shearing-box boundaries the density flux and emf_y of the two x borders are averaged with the y-remapped
// opposite border (:3237-3297), read through Fb / Eb (any row of the border columns), and the border density is floored.
// F / E give the fluxes / emfs around the cell itself.  Returns the inverse time step (0 outside the inner cells).
// BORDERS = false: a caller that never passes a cell next to a shearing x border (the fused kernel leaves those three
// columns to k_update_rot_border) compiles the remap code away -- the fused kernel has to fit the instruction cache.
// WITH_DT = false: no inverse-dt estimate of the new state (the caller runs the stand-alone reduction afterwards).
template <bool BORDERS = true, bool WITH_DT = true, typename T, typename UV, typename FV, typename EV, typename FBV, typename EBV>
__device__ __forceinline__ T update_cell_rot(const KParams<T>& P, const UV& U, T* __restrict__ Unew, const FV& F,
                                             const EV& E, const FBV& Fb, const EBV& Eb, int i, int j, int k, T dt,
                                             const ShearShift<T>& sh) {
  const int gw = P.gw;
  const int iN = P.isize - gw, jN = P.jsize - gw, kN = P.ksize - gw;
  const size_t plane = (size_t)P.isize * P.jsize, comp = plane * P.ksize;
  const size_t idx = (size_t)k * plane + (size_t)j * P.isize + i;
  const T dtdx = dt / P.dx, dtdy = dt / P.dy, dtdz = dt / P.dz;
  const bool inner = i < iN && j < jN && k < kN;
  T lambda = P.Omega0 * dt;
  lambda = T(0.25) * lambda * lambda;
  const T il = dev::rcp(T(1) + lambda);
  const T ratio = (T(1) - lambda) * il, alpha1 = il, alpha2 = P.Omega0 * dt * il;
  T invDt = T(0);
  T un[8];
#pragma unroll
  for (int v = 0; v < 8; ++v) un[v] = U(v, i, j, k);
  if (inner) {
    const T dsx = T(2) * P.Omega0 * dt * un[IV] * il, dsy = T(-0.5) * P.Omega0 * dt * un[IU] * il;
    T m[5];
    m[ID] = un[ID]; m[IP] = un[IP]; m[IW] = un[IW];
    m[IU] = un[IU] * ratio + dsx;
    m[IV] = un[IV] * ratio + dsy;
    const bool bLo = BORDERS && sh.enabled && i == gw, bHi = BORDERS && sh.enabled && i == P.nx + gw - 1;
    // flux contributions in the reference's order: +x(i) +y(j) +z(k) -x(i+1) -y(j+1) -z(k+1)
    auto add = [&](int c0, int ii, int jj, int kk, T s, T dtd, bool skipDensity) {
      const T fd = F(c0 + 0, ii, jj, kk), fp = F(c0 + 1, ii, jj, kk), fu = F(c0 + 2, ii, jj, kk),
              fv = F(c0 + 3, ii, jj, kk), fw = F(c0 + 4, ii, jj, kk);
      if (!skipDensity) m[ID] += s * fd * dtd;
      m[IP] += s * fp * dtd;
      m[IU] += s * (alpha1 * fu + alpha2 * fv) * dtd;
      m[IV] += s * (alpha1 * fv - T(0.25) * alpha2 * fu) * dtd;
      m[IW] += s * fw * dtd;
    };
    add(0, i, j, k, T(1), dtdx, bLo);
    add(5, i, j, k, T(1), dtdy, false);
    add(10, i, j, k, T(1), dtdz, false);
    add(0, i + 1, j, k, T(-1), dtdx, bHi);
    add(5, i, j + 1, k, T(-1), dtdy, false);
    add(10, i, j, k + 1, T(-1), dtdz, false);
    if (P.gravity) {  // gravity source term BEFORE the border remap of the density (MHDRunGodunov.cpp:3188-3192 vs :3203)
      const T hdt = T(0.5) * dt, rs = un[ID] + m[ID];
      m[IU] += hdt * P.gx * rs; m[IV] += hdt * P.gy * rs;
      m[IW] += hdt * grav_z(P, k) * rs;
    }
    if (BORDERS && (bLo || bHi)) {  // remapped border density flux, :3237-3297
      int j0, j1; T eps;
      remapRows(P, sh, j, bLo, j0, j1, eps);
      const int iOwn = bLo ? gw : P.nx + gw, iOpp = bLo ? P.nx + gw : gw;
      const T own = F(0, iOwn, j, k) * dtdx;
      const T rem = T(0.5) * (own + (T(1) - eps) * (Fb(0, iOpp, j0, k) * dtdx) + eps * (Fb(0, iOpp, j1, k) * dtdx));
      m[ID] = bLo ? m[ID] + rem : m[ID] - rem;
      m[ID] = dev::mx(m[ID], P.smallr);
    }
#pragma unroll
    for (int v = 0; v < 5; ++v) un[v] = m[v];
  }
  // emf_y on the two x borders is the average with the remapped opposite border, :3251-3274
  auto emfY = [&](int ii, int jj, int kk) -> T {
    if (ii > iN || jj > jN || kk > kN) return T(0);
    const T own = E(1, ii, jj, kk);
    if (BORDERS && sh.enabled && (ii == gw || ii == P.nx + gw)) {
      int j0, j1; T eps;
      remapRows(P, sh, jj, ii == gw, j0, j1, eps);
      const int iOpp = (ii == gw) ? P.nx + gw : gw;
      return T(0.5) * (own + (T(1) - eps) * Eb(1, iOpp, j0, kk) + eps * Eb(1, iOpp, j1, kk));
    }
    return own;
  };
  auto emf = [&](int c, int ii, int jj, int kk) -> T {
    return (ii > iN || jj > jN || kk > kN) ? T(0) : E(c, ii, jj, kk);
  };
  // constrained transport per face component (same operation order as the reference's per-cell sequence)
  auto ctx = [&](int ii, int jj, int kk, T bx) -> T {
    if (kk < kN) bx += (emf(0, ii, jj + 1, kk) - emf(0, ii, jj, kk)) * dtdy;
    return bx - (emfY(ii, jj, kk + 1) - emfY(ii, jj, kk)) * dtdz;
  };
  auto cty = [&](int ii, int jj, int kk, T by) -> T {
    if (kk < kN) by -= (emf(0, ii + 1, jj, kk) - emf(0, ii, jj, kk)) * dtdx;
    return by + (emf(2, ii, jj, kk + 1) - emf(2, ii, jj, kk)) * dtdz;
  };
  auto ctz = [&](int ii, int jj, int kk, T bz) -> T {
    bz += (emfY(ii + 1, jj, kk) - emfY(ii, jj, kk)) * dtdx;
    return bz - (emf(2, ii, jj + 1, kk) - emf(2, ii, jj, kk)) * dtdy;
  };
  un[IA] = ctx(i, j, k, un[IA]);
  un[IB] = cty(i, j, k, un[IB]);
  un[IC] = ctz(i, j, k, un[IC]);
#pragma unroll
  for (int v = 0; v < 8; ++v) Unew[v * comp + idx] = un[v];
  if (WITH_DT && inner) {
    const T bxp = ctx(i + 1, j, k, U(IA, i + 1, j, k));
    const T byp = cty(i, j + 1, k, U(IB, i, j + 1, k));
    const T bzp = ctz(i, j, k + 1, U(IC, i, j, k + 1));
    T q[8];
    dev::cons_to_prim_mhd(P, un, bxp, byp, bzp, T(0), q);
    const T irho = dev::rcp(q[ID]);
    const T a2 = q[IA] * q[IA], b2 = q[IB] * q[IB], c2 = q[IC] * q[IC];
    const T bb = a2 + b2 + c2;
    const T vx = dev::fast_speed(P.gamma0, q[IP], irho, bb, a2) + dev::ab(q[IU]);
    const T vy = dev::fast_speed(P.gamma0, q[IP], irho, bb, b2) + dev::ab(q[IV]) +
                 T(1.5) * P.Omega0 * (P.xMax - P.xMin) * T(0.5);
    const T vz = dev::fast_speed(P.gamma0, q[IP], irho, bb, c2) + dev::ab(q[IW]);
    invDt = vx * P.rdx + vy * P.rdy + vz * P.rdz;
  }
  return invDt;
}

// ------------------------------------------------------------------------------------------------
// one face flux / one corner emf of the FAST configuration (adiabatic, non-rotating, HLLD + 2-D HLLD) from W;
// W, F, E are any accessors: shared-memory tiles in the fused kernel, host arrays in the CPU test suite
// ------------------------------------------------------------------------------------------------
template <bool FAST = true, typename T, typename WV, typename FV>
__device__ __forceinline__ void fused_flux_task(const KParams<T>& P, const WV& W,
                                                const FV& F, int dir, int i, int j, int k) {
  dev::State<T> L, R;
  if (dir == 0) {
    L = face_state<T, 0>(P, W, i - 1, j, k, T(1));
    R = face_state<T, 0>(P, W, i, j, k, T(-1));
  } else if (dir == 1) {
    L = face_state<T, 1>(P, W, i, j - 1, k, T(1));
    R = face_state<T, 1>(P, W, i, j, k, T(-1));
  } else {
    L = face_state<T, 2>(P, W, i, j, k - 1, T(1));
    R = face_state<T, 2>(P, W, i, j, k, T(-1));
  }
  T f[8];
  dev::riemann_hlld<FAST>(P, L, R, f);
  if (!FAST && dir == 1 && P.Omega0 > T(0)) {
    // rotating frame: upwind advection of the y flux by the background shear, as in flux_cell (MHDRunGodunov.cpp:2860-2899)
    const T xPos = P.xMin + P.dx * T(0.5) + (i - P.gw) * P.dx;
    const T shear_y = T(-1.5) * P.Omega0 * xPos;
    const T bn = T(0.5) * (L.a + R.a);
    const dev::State<T>& S = (shear_y > T(0)) ? L : R;
    const T pS = (P.cIso > T(0)) ? S.r * P.cIso * P.cIso : S.p;
    const T eMag = T(0.5) * (bn * bn + S.b * S.b + S.c * S.c);
    const T eKin = T(0.5) * (S.u * S.u + S.v * S.v + S.w * S.w);
    const T eTot = eKin + eMag + pS / (P.gamma0 - T(1));
    f[ID] += shear_y * S.r;
    f[IP] += shear_y * (eTot + eMag - bn * bn);
    f[IU] += shear_y * S.r * S.u;
    f[IV] += shear_y * S.r * S.v;
    f[IW] += shear_y * S.r * S.w;
  }
  const int c0 = 5 * dir;
  F(c0 + 0, i, j, k) = f[ID];
  F(c0 + 1, i, j, k) = f[IP];
  F(c0 + 2, i, j, k) = (dir == 0) ? f[IU] : (dir == 1) ? f[IV] : f[IW];
  F(c0 + 3, i, j, k) = (dir == 1) ? f[IU] : f[IV];
  F(c0 + 4, i, j, k) = (dir == 2) ? f[IU] : f[IW];
}

template <bool FAST = true, typename T, typename WV, typename EV>
__device__ __forceinline__ void fused_emf_task(const KParams<T>& P, const WV& W,
                                               const EV& E, int edir, int i, int j, int k) {
  dev::Corner<T> RT, RB, LT, LB;
  if (edir == 2) {
    RT = edge_state<T, 2>(P, W, i - 1, j - 1, k, T(1), T(1));
    RB = edge_state<T, 2>(P, W, i - 1, j, k, T(1), T(-1));
    LT = edge_state<T, 2>(P, W, i, j - 1, k, T(-1), T(1));
    LB = edge_state<T, 2>(P, W, i, j, k, T(-1), T(-1));
  } else if (edir == 1) {
    RT = edge_state<T, 1>(P, W, i - 1, j, k - 1, T(1), T(1));
    RB = edge_state<T, 1>(P, W, i, j, k - 1, T(-1), T(1));
    LT = edge_state<T, 1>(P, W, i - 1, j, k, T(1), T(-1));
    LB = edge_state<T, 1>(P, W, i, j, k, T(-1), T(-1));
  } else {
    RT = edge_state<T, 0>(P, W, i, j - 1, k - 1, T(1), T(1));
    RB = edge_state<T, 0>(P, W, i, j - 1, k, T(1), T(-1));
    LT = edge_state<T, 0>(P, W, i, j, k - 1, T(-1), T(1));
    LB = edge_state<T, 0>(P, W, i, j, k, T(-1), T(-1));
  }
  // the fused kernels run the 2-D HLLD solver only (the launch wrappers check magRiemannSolver); xPos: shear terms
  const T xPos = FAST ? T(0) : P.xMin + P.dx * T(0.5) + (i - P.gw) * P.dx;
  E(2 - edir, i, j, k) = dev::compute_emf<FAST, true>(P, RT, RB, LT, LB, edir, xPos);
}

// ------------------------------------------------------------------------------------------------
// generic path: Riemann flux through the LOW face normal to DIR of cell (i,j,k) (any solver; rotating-frame shear
// advection of the y flux), stored in physical component order (reference cpu_v3.cpp:397-465)
// ------------------------------------------------------------------------------------------------
template <typename T, int DIR, bool FAST, typename WV, typename FV>
__device__ __forceinline__ void flux_cell(const KParams<T>& P, const WV& W, const FV& F, int i, int j, int k) {
  const int gw = P.gw;
  const int il = i - (DIR == 0), jl = j - (DIR == 1), kl = k - (DIR == 2);
  const dev::State<T> L = face_state<T, DIR>(P, W, il, jl, kl, T(1));
  const dev::State<T> R = face_state<T, DIR>(P, W, i, j, k, T(-1));
  T f[8];
  dev::riemann_mhd<FAST>(P, L, R, f);
  if (!FAST && DIR == 1 && P.Omega0 > T(0)) {
    // rotating frame: upwind advection of the y flux by the background shear
    // (MHDRunGodunov.cpp:2860-2899; the states are those the Riemann solver has seen: mean normal
    // field, isothermal pressure when cIso > 0 and the solver is HLLD)
    const T xPos = P.xMin + P.dx * T(0.5) + (i - gw) * P.dx;
    const T shear_y = T(-1.5) * P.Omega0 * xPos;
    const T bn = T(0.5) * (L.a + R.a);
    const dev::State<T>& S = (shear_y > T(0)) ? L : R;
    const T pS = (P.cIso > T(0) && P.riemannSolver == RS_HLLD) ? S.r * P.cIso * P.cIso : S.p;
    const T eMag = T(0.5) * (bn * bn + S.b * S.b + S.c * S.c);
    const T eKin = T(0.5) * (S.u * S.u + S.v * S.v + S.w * S.w);
    const T eTot = eKin + eMag + pS / (P.gamma0 - T(1));
    f[ID] += shear_y * S.r;
    f[IP] += shear_y * (eTot + eMag - bn * bn);
    f[IU] += shear_y * S.r * S.u;
    f[IV] += shear_y * S.r * S.v;
    f[IW] += shear_y * S.r * S.w;
  }
  // store in physical component order (undo the frame permutation)
  const int c0 = 5 * DIR;
  F(c0 + 0, i, j, k) = f[ID];
  F(c0 + 1, i, j, k) = f[IP];
  F(c0 + 2, i, j, k) = (DIR == 0) ? f[IU] : (DIR == 1) ? f[IV] : f[IW];
  F(c0 + 3, i, j, k) = (DIR == 1) ? f[IU] : f[IV];
  F(c0 + 4, i, j, k) = (DIR == 2) ? f[IU] : f[IW];
}

// ------------------------------------------------------------------------------------------------
// generic path: corner emf along EDIR at the low edge of cell (i,j,k) (reference cpu_v3.cpp:539-579)
// ------------------------------------------------------------------------------------------------
template <typename T, int EDIR, bool FAST, typename WV, typename EV>
__device__ __forceinline__ void emf_cell(const KParams<T>& P, const WV& W, const EV& E, int i, int j, int k) {
  const int gw = P.gw;
  const T xPos = P.xMin + P.dx * T(0.5) + (i - gw) * P.dx;
  dev::Corner<T> RT, RB, LT, LB;
  if (EDIR == 2) {  // cpu_v3.cpp:550-557 : (s1,s2) = (x,y)
    RT = edge_state<T, 2>(P, W, i - 1, j - 1, k, T(1), T(1));
    RB = edge_state<T, 2>(P, W, i - 1, j, k, T(1), T(-1));
    LT = edge_state<T, 2>(P, W, i, j - 1, k, T(-1), T(1));
    LB = edge_state<T, 2>(P, W, i, j, k, T(-1), T(-1));
  } else if (EDIR == 1) {  // cpu_v3.cpp:561-569 : (s1,s2) = (x,z); RB and LT swapped
    RT = edge_state<T, 1>(P, W, i - 1, j, k - 1, T(1), T(1));
    RB = edge_state<T, 1>(P, W, i, j, k - 1, T(-1), T(1));   // LT2(i,j,k-1)
    LT = edge_state<T, 1>(P, W, i - 1, j, k, T(1), T(-1));   // RB2(i-1,j,k)
    LB = edge_state<T, 1>(P, W, i, j, k, T(-1), T(-1));
  } else {  // cpu_v3.cpp:572-579 : (s1,s2) = (y,z)
    RT = edge_state<T, 0>(P, W, i, j - 1, k - 1, T(1), T(1));
    RB = edge_state<T, 0>(P, W, i, j - 1, k, T(1), T(-1));
    LT = edge_state<T, 0>(P, W, i, j, k - 1, T(-1), T(1));
    LB = edge_state<T, 0>(P, W, i, j, k, T(-1), T(-1));
  }
  // reference component order: I_EMFZ = 0, I_EMFY = 1, I_EMFX = 2
  E(2 - EDIR, i, j, k) = dev::compute_emf<FAST>(P, RT, RB, LT, LB, EDIR, xPos);
}

}  // namespace

}  // namespace rg
