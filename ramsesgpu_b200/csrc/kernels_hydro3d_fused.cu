// 3D Euler (hydro) Godunov step for sm_100a in ONE kernel: U -> U', no traced-state scratch in HBM.
//
// A 32 x CY thread block owns the 32 x CY column of cells around a 28 x (CY-4) column of UPDATED cells and marches along
// z.  Per plane every thread
//   * converts the conservative state of ITS cell two planes ahead (prefetched one iteration earlier) to primitives:
//     the three planes a z slope needs live in registers,
//   * gets the primitives of its x neighbours by warp shuffle and of its y neighbours from a shared-memory tile,
//   * evaluates the 15 limited slopes and the half-step predictor (hydro_trace_from_prims: the code of the separate
//     trace kernel) -- the traced state never leaves the registers,
//   * builds its six face states, takes the LEFT state of its low x face from lane-1 (shuffle), of its low y face from
//     row-1 (shared memory), of its low z face from its own previous plane (register carry), solves the three Riemann
//     problems, and takes the high x / y fluxes back from lane+1 / row+1; the update of plane p is closed by the low z
//     flux of plane p+1.
// Every face flux is solved once per tile; faces on tile seams are solved by both tiles with identical inputs and code.
// Per updated cell the kernel reads U 1.5 times (tile halo) and writes it once: 50 B in FP32 against the 220 B of the
// two-kernel path through W[20] (reference: HydroRunGodunov.cpp:2658-2890 godunov_unsplit_cpu_v1, trace.h:544-661,
// slope.h:324-427, riemann.h, and the GPU kernels godunov_unsplit.cuh:1829,3212 whose work this replaces).
#include <algorithm>

#include "hydro_cells.cuh"
#include "hydro_device.cuh"
#include "kernel_common.cuh"
#include "kernels.h"

namespace rg {

int g_hydroFused = 1;  // run-time knob "hydro_fused": the one-kernel step (default) or trace + flux/update through W
int g_hydroRows = 0;   // run-time knob "hydro_rows": rows of the thread block (0 = default: 20 in FP32, 12 in FP64)

namespace {

template <typename T, int TY_>
struct HydroFusedTile {
  static constexpr int CX = 32, CY = TY_, UX = CX - 4, UY = CY - 4, THREADS = CX * CY;
  static constexpr unsigned SMEM = (unsigned)(3 * 5 * CY * CX * sizeof(T));  // primitives | high-y face states | low-y fluxes
};

template <typename T, int RS, typename C>
__global__ void __launch_bounds__(C::THREADS, 1)
k_hydro_fused(const __grid_constant__ KParams<T> P, const T* __restrict__ Uold, T* __restrict__ Unew, int k0, int k1, int lz,
              T dt, unsigned long long* __restrict__ slots) {
  extern __shared__ unsigned char smemRawH[];
  T* sQ = reinterpret_cast<T*>(smemRawH);   // [5][CY][32] primitives of the plane being traced
  T* sY = sQ + 5 * C::CY * C::CX;           // [5][CY][32] high-y face states
  T* sF = sY + 5 * C::CY * C::CX;           // [5][CY][32] low-y fluxes
  const int tx = threadIdx.x, ty = threadIdx.y, gw = P.gw;
  const int i = gw + blockIdx.x * C::UX - 2 + tx, j = gw + blockIdx.y * C::UY - 2 + ty;
  const int za = k0 + blockIdx.z * lz, zb = min(za + lz, k1);  // updated planes [za, zb)
  if (za >= zb) return;
  const int iN = P.isize - gw, jN = P.jsize - gw;
  const bool cellOK = i < P.isize && j < P.jsize;
  const bool traceOK = tx >= 1 && tx <= C::CX - 2 && ty >= 1 && ty <= C::CY - 2 && i <= P.isize - 2 && j <= P.jsize - 2;
  const bool rowUpd = ty >= 2 && ty <= C::CY - 3 && j < jN, colUpd = tx >= 2 && tx <= C::CX - 3 && i < iN;
  const bool upd = rowUpd && colUpd;
  const bool solveX = tx >= 2 && tx <= C::CX - 2 && i <= iN && rowUpd;  // low x face of the cell is a face of an updated cell
  const bool solveY = ty >= 2 && ty <= C::CY - 2 && j <= jN && colUpd;
  const size_t plane = (size_t)P.isize * P.jsize, comp = plane * P.ksize;
  const size_t col = (size_t)j * P.isize + i;
  const T dtdx = dt / P.dx, dtdy = dt / P.dy, dtdz = dt / P.dz;
  const int sidx = ty * C::CX + tx;
  constexpr int SV = C::CY * C::CX;  // stride between variables in the shared tiles

  auto loadRaw = [&](int q, T (&u)[5]) {
    if (cellOK) {
#pragma unroll
      for (int v = 0; v < 5; ++v) u[v] = __ldg(Uold + v * comp + (size_t)q * plane + col);
    } else {
      u[ID] = T(1); u[IP] = T(1); u[IU] = T(0); u[IV] = T(0); u[IW] = T(0);
    }
  };
  auto toPrim = [&](const T (&u)[5], T (&q)[5]) { dev::cons_to_prim_hydro(P, u[ID], u[IP], u[IU], u[IV], u[IW], q); };

  T qm1[5], q0[5], qp1[5], raw[5];
  loadRaw(za - 2, raw); toPrim(raw, qm1);
  loadRaw(za - 1, raw); toPrim(raw, q0);
  loadRaw(za, raw);     toPrim(raw, qp1);
#pragma unroll
  for (int v = 0; v < 5; ++v) sQ[v * SV + sidx] = q0[v];
  __syncthreads();

  T invDt = T(0);
  dev::HState<T> zPrev{T(1), T(1), T(0), T(0), T(0)};  // high-z face state of the previous plane
  T acc[5] = {T(0), T(0), T(0), T(0), T(0)};            // update of the previous plane, all terms but the high z flux
  for (int p = za - 1; p <= zb; ++p) {
    const bool more = p + 2 <= zb + 1;  // block-uniform: plane p+2 is needed by a later iteration
    if (more) loadRaw(p + 2, raw);      // in flight during the work below
    // primitives of the x neighbours by shuffle, of the y neighbours from the shared tile
    T qxm[5], qxp[5];
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      qxm[v] = __shfl_up_sync(0xffffffffu, q0[v], 1);
      qxp[v] = __shfl_down_sync(0xffffffffu, q0[v], 1);
    }
    T w[NW_HYDRO];
    if (traceOK) {
      T qym[5], qyp[5];
#pragma unroll
      for (int v = 0; v < 5; ++v) {
        qym[v] = sQ[v * SV + sidx - C::CX];
        qyp[v] = sQ[v * SV + sidx + C::CX];
      }
      hydro_trace_from_prims(P, q0, qxm, qxp, qym, qyp, qm1, qp1, dt, w);
    } else {
#pragma unroll
      for (int c = 0; c < NW_HYDRO; ++c) w[c] = (c < 2) ? T(1) : T(0);
    }
    const bool mid = p >= za && p < zb;  // block-uniform
    // z face: left state carried over from the previous plane
    T fz[5] = {T(0), T(0), T(0), T(0), T(0)};
    if (p >= za && upd) face_flux<T, 2, RS>(P, zPrev, face_from_regs<T, 2>(P, w, T(-1)), fz);
    if (p > za && upd) {  // the low z flux of this plane closes the update of the plane below
      const size_t idx = (size_t)(p - 1) * plane + col;
      T r5[5];
#pragma unroll
      for (int v = 0; v < 5; ++v) r5[v] = acc[v] - fz[v] * dtdz;
      if (P.gravity) {  // static gravity source term, reference HydroRunBase.cpp:1962-1976
        const T hdt = T(0.5) * dt, rs = __ldg(Uold + idx) + r5[ID];
        r5[IU] += hdt * P.gx * rs; r5[IV] += hdt * P.gy * rs; r5[IW] += hdt * P.gz * rs;
      }
#pragma unroll
      for (int v = 0; v < 5; ++v) Unew[v * comp + idx] = r5[v];
      T q[5];
      const T c = dev::cons_to_prim_hydro(P, r5[ID], r5[IP], r5[IU], r5[IV], r5[IW], q);
      invDt = dev::mx(invDt, (c + dev::ab(q[IU])) * P.rdx + (c + dev::ab(q[IV])) * P.rdy + (c + dev::ab(q[IW])) * P.rdz);
    }
    zPrev = face_from_regs<T, 2>(P, w, T(1));
    T fxl[5] = {T(0), T(0), T(0), T(0), T(0)}, fxh[5], fyl[5] = {T(0), T(0), T(0), T(0), T(0)};
    T un[5] = {T(0), T(0), T(0), T(0), T(0)};
    if (mid && upd) {  // old state of the cell: requested here, consumed after the two barriers below
      const size_t idx = (size_t)p * plane + col;
#pragma unroll
      for (int v = 0; v < 5; ++v) un[v] = __ldg(Uold + v * comp + idx);
    }
    if (mid) {
      // x faces: left state from lane-1, high flux from lane+1
      const dev::HState<T> hi = face_from_regs<T, 0>(P, w, T(1));
      dev::HState<T> L;
      L.r = __shfl_up_sync(0xffffffffu, hi.r, 1); L.p = __shfl_up_sync(0xffffffffu, hi.p, 1);
      L.u = __shfl_up_sync(0xffffffffu, hi.u, 1); L.v = __shfl_up_sync(0xffffffffu, hi.v, 1);
      L.w = __shfl_up_sync(0xffffffffu, hi.w, 1);
      if (solveX) face_flux<T, 0, RS>(P, L, face_from_regs<T, 0>(P, w, T(-1)), fxl);
#pragma unroll
      for (int v = 0; v < 5; ++v) fxh[v] = __shfl_down_sync(0xffffffffu, fxl[v], 1);
      // y faces: publish the high-y face state of this cell
      const dev::HState<T> hy = face_from_regs<T, 1>(P, w, T(1));
      sY[0 * SV + sidx] = hy.r; sY[1 * SV + sidx] = hy.p; sY[2 * SV + sidx] = hy.u; sY[3 * SV + sidx] = hy.v; sY[4 * SV + sidx] = hy.w;
    }
    __syncthreads();  // A: sY complete; every read of sQ (plane p) is done
    // the primitives of plane p+1 replace those of plane p for the next iteration
#pragma unroll
    for (int v = 0; v < 5; ++v) sQ[v * SV + sidx] = qp1[v];
    if (mid) {
      if (solveY) {
        const dev::HState<T> L{sY[0 * SV + sidx - C::CX], sY[1 * SV + sidx - C::CX], sY[2 * SV + sidx - C::CX],
                               sY[3 * SV + sidx - C::CX], sY[4 * SV + sidx - C::CX]};
        face_flux<T, 1, RS>(P, L, face_from_regs<T, 1>(P, w, T(-1)), fyl);
      }
#pragma unroll
      for (int v = 0; v < 5; ++v) sF[v * SV + sidx] = fyl[v];
    }
    __syncthreads();  // B: sF and the new sQ complete; every read of sY is done
    if (mid && upd) {
#pragma unroll
      for (int v = 0; v < 5; ++v) {  // summation order of the reference's serial scatter
        T s = un[v];
        s += fxl[v] * dtdx; s += fyl[v] * dtdy; s += fz[v] * dtdz;
        s -= fxh[v] * dtdx; s -= sF[v * SV + sidx + C::CX] * dtdy;
        acc[v] = s;
      }
    }
#pragma unroll
    for (int v = 0; v < 5; ++v) { qm1[v] = q0[v]; q0[v] = qp1[v]; }
    if (more) toPrim(raw, qp1);
  }
  if (slots != nullptr) reduceMaxToSlots(invDt, slots);
}

template <typename T, typename C>
void launchHydroFused(const KParams<T>& P, const T* Uold, T* Unew, int k0, int k1, T dt, unsigned long long* slots,
                      cudaStream_t s) {
  static bool attrSetDev[MAX_DEVICES] = {false};
  bool& attrSet = attrSetDev[currentDevice()];
  if (!attrSet) {
    cudaFuncSetAttribute(k_hydro_fused<T, RS_HLLC, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    cudaFuncSetAttribute(k_hydro_fused<T, RS_HLL, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    cudaFuncSetAttribute(k_hydro_fused<T, RS_APPROX, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    cudaFuncSetAttribute(k_hydro_fused<T, -1, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    attrSet = true;
  }
  const int ntx = (P.nx + C::UX - 1) / C::UX, nty = (P.ny + C::UY - 1) / C::UY, planes = k1 - k0, nSM = smCount();
  // z ranges: whole waves of one block per SM; a range costs its planes + 3 (pipeline fill)
  int bestNz = 1;
  double bestCost = 1e300;
  for (int nz = 1; nz <= planes; ++nz) {
    const int lz = (planes + nz - 1) / nz;
    if (lz < 8 && nz > 1) break;
    const long blocks = (long)ntx * nty * ((planes + lz - 1) / lz);
    const double cost = (double)((blocks + nSM - 1) / nSM) * (lz + 3.0);
    if (cost < bestCost) { bestCost = cost; bestNz = nz; }
  }
  const int lz = (planes + bestNz - 1) / bestNz;
  const dim3 g(ntx, nty, (planes + lz - 1) / lz), b(C::CX, C::CY, 1);
  switch (P.riemannSolver) {  // one instantiation per Riemann solver: a single solver body in the kernel
    case RS_HLLC: k_hydro_fused<T, RS_HLLC, C><<<g, b, C::SMEM, s>>>(P, Uold, Unew, k0, k1, lz, dt, slots); break;
    case RS_HLL: k_hydro_fused<T, RS_HLL, C><<<g, b, C::SMEM, s>>>(P, Uold, Unew, k0, k1, lz, dt, slots); break;
    case RS_APPROX: k_hydro_fused<T, RS_APPROX, C><<<g, b, C::SMEM, s>>>(P, Uold, Unew, k0, k1, lz, dt, slots); break;
    default: k_hydro_fused<T, -1, C><<<g, b, C::SMEM, s>>>(P, Uold, Unew, k0, k1, lz, dt, slots); break;
  }
  launched();
}

}  // namespace

bool hydroFusedRequested() { return g_hydroFused != 0; }

template <typename T>
void HydroKernels<T>::fusedStep(const KParams<T>& P, const T* Uold, T* Unew, int k0, int k1, T dt, unsigned long long* slots,
                                cudaStream_t s) {
  if (k1 <= k0) return;
  // measured at 512^3 FP32 HLLC / 384^3 FP64 (profiles/r02_c_hydro_rows_ab.txt): 12 rows 7.80 / 6.24 ms, 16: 6.50 / 6.50,
  // 20: 5.95 / 7.27 (spills in FP64), 24: 6.00 / 8.65
  const int rows = g_hydroRows ? g_hydroRows : (sizeof(T) == 4 ? 20 : 12);
  if (rows == 24) launchHydroFused<T, HydroFusedTile<T, 24>>(P, Uold, Unew, k0, k1, dt, slots, s);
  else if (rows == 20) launchHydroFused<T, HydroFusedTile<T, 20>>(P, Uold, Unew, k0, k1, dt, slots, s);
  else if (rows == 16) launchHydroFused<T, HydroFusedTile<T, 16>>(P, Uold, Unew, k0, k1, dt, slots, s);
  else launchHydroFused<T, HydroFusedTile<T, 12>>(P, Uold, Unew, k0, k1, dt, slots, s);
  copyGhosts(P, Uold, Unew, k0, k1, s);
}

template void HydroKernels<double>::fusedStep(const KParams<double>&, const double*, double*, int, int, double,
                                              unsigned long long*, cudaStream_t);
template void HydroKernels<float>::fusedStep(const KParams<float>&, const float*, float*, int, int, float, unsigned long long*,
                                             cudaStream_t);

}  // namespace rg
