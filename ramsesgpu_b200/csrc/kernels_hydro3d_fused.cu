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
#include <cstring>

#include "hydro_cells.cuh"
#include "hydro_device.cuh"
#include "kernel_common.cuh"
#include "kernels.h"
#include "tma.cuh"

namespace rg {

int g_hydroFused = 1;  // run-time knob "hydro_fused": the one-kernel step (default) or trace + flux/update through W
int g_hydroTma = 1;    // run-time knob "hydro_tma": conservative tiles by TMA (default) or per-thread loads
int g_hydroRows = 0;   // run-time knob "hydro_rows": rows of the thread block (0 = default: 24 in FP32, 16 in FP64 with TMA tiles; 20 / 12 without)

namespace {

template <typename T, int TY_>
struct HydroFusedTile {
  static constexpr int CX = 32, CY = TY_, UX = CX - 4, UY = CY - 4, THREADS = CX * CY;
  static constexpr unsigned XCH = (unsigned)(3 * 5 * CY * CX * sizeof(T));   // primitives | high-y face states | low-y fluxes
  // TMA path: ring of 4 planes of the conservative tile [5][CY][32] (plane p for the update, p+1, p+2 being converted,
  // p+3 in flight) behind the exchange arrays, 128-byte aligned, + 4 mbarriers
  static constexpr unsigned U_BYTES = (unsigned)(5 * CY * CX * sizeof(T));
  static constexpr unsigned U_OFF = (XCH + 127u) / 128u * 128u;
  static constexpr unsigned SMEM = XCH, SMEM_TMA = 128u + U_OFF + 4u * U_BYTES + 64u;
};

// TMAU = true: the conservative tile of a plane (all 5 variables, 32 x CY cells with the halo) arrives by ONE
// cp.async.bulk.tensor.4d per plane (UTMALDG, mbarrier completion, issued three planes ahead by one thread) in a 4-plane
// shared-memory ring; threads read their cell from there for cons->prim and again, two planes later, as the old state of
// the update.  No per-thread global loads of U, no prefetch registers.  TMAU = false: per-thread coalesced loads (row
// pitch not a multiple of 16 bytes).
template <typename T, int RS, typename C, bool TMAU>
__global__ void __launch_bounds__(C::THREADS, 1)
k_hydro_fused(const __grid_constant__ KParams<T> P, const __grid_constant__ CUtensorMap mapU, const T* __restrict__ Uold,
              T* __restrict__ Unew, int k0, int k1, int lz, T dt, unsigned long long* __restrict__ slots) {
  extern __shared__ unsigned char smemRawH0[];
  // 128-byte alignment for the TMA destination (computed on the shared-window address: LDS/STS, not generic LD/ST)
  unsigned char* smemRawH = smemRawH0 + (TMAU ? ((128u - (tma::smemAddr(smemRawH0) & 127u)) & 127u) : 0u);
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

  // conservative tiles (TMA path): plane q lives in ring slot (q - (za-2)) & 3; out-of-range cells arrive as zeros
  // (cons->prim floors them to a finite state, which no updated cell ever reads)
  unsigned char* const uring = smemRawH + C::U_OFF;
  uint64_t* const ubar = reinterpret_cast<uint64_t*>(uring + 4u * C::U_BYTES);
  const bool leader = threadIdx.x == 0 && threadIdx.y == 0;
  auto uslot = [&](int q) -> const T* { return reinterpret_cast<const T*>(uring + (unsigned)((q - (za - 2)) & 3) * C::U_BYTES); };
  auto issue = [&](int q) {  // one thread
    const int n = q - (za - 2);
    tma::mbarExpectTx(&ubar[n & 3], C::U_BYTES);
    tma::loadTile4D(uring + (unsigned)(n & 3) * C::U_BYTES, &mapU, &ubar[n & 3], gw + blockIdx.x * C::UX - 2,
                    gw + blockIdx.y * C::UY - 2, q, 0);
  };
  auto arrived = [&](int q) {
    const int n = q - (za - 2);
    tma::mbarWait(&ubar[n & 3], (unsigned)(n >> 2) & 1u);
  };
  if (TMAU) {
    if (leader) {
      for (int n = 0; n < 4; ++n) tma::mbarInit(&ubar[n], 1);
      tma::fenceBarrierInit();
    }
    __syncthreads();
    if (leader)
      for (int q = za - 2; q <= min(za + 1, zb + 1); ++q) issue(q);
  }
  T raw[5];
  auto loadRaw = [&](int q, T (&u)[5]) {
    if (TMAU) return;  // (the tile is on its way)
    if (cellOK) {
#pragma unroll
      for (int v = 0; v < 5; ++v) u[v] = __ldg(Uold + v * comp + (size_t)q * plane + col);
    } else {
      u[ID] = T(1); u[IP] = T(1); u[IU] = T(0); u[IV] = T(0); u[IW] = T(0);
    }
  };
  auto toPrim = [&](int q, const T (&u)[5], T (&pr)[5]) {
    if (TMAU) {
      arrived(q);
      const T* t = uslot(q);
      dev::cons_to_prim_hydro(P, t[ID * SV + sidx], t[IP * SV + sidx], t[IU * SV + sidx], t[IV * SV + sidx], t[IW * SV + sidx], pr);
    } else {
      dev::cons_to_prim_hydro(P, u[ID], u[IP], u[IU], u[IV], u[IW], pr);
    }
  };

  T qm1[5], q0[5], qp1[5];
  loadRaw(za - 2, raw); toPrim(za - 2, raw, qm1);
  loadRaw(za - 1, raw); toPrim(za - 1, raw, q0);
  loadRaw(za, raw);     toPrim(za, raw, qp1);
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
    if (!TMAU && mid && upd) {  // old state of the cell: requested here, consumed after the two barriers below
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
    if (TMAU && leader && p + 3 <= zb + 1) {  // the slot of plane p-1 (last read before this barrier) takes plane p+3
      tma::fenceProxyAsync();
      issue(p + 3);
    }
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
      if (TMAU) {  // old state of the cell from the ring (the slot of plane p is refilled one iteration later)
        const T* t = uslot(p);
#pragma unroll
        for (int v = 0; v < 5; ++v) un[v] = t[v * SV + sidx];
      }
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
    if (more) toPrim(p + 2, raw, qp1);
  }
  if (slots != nullptr) reduceMaxToSlots(invDt, slots);
}

// tensor map of the state array [var][k][j][i] with a box of 32 x CY x 1 x 5 elements, cached per (array, shape)
template <typename T, typename C>
static const CUtensorMap* hydroTensorMap(const KParams<T>& P, const T* U) {
  struct Entry { const void* base; int isize, jsize, ksize, dev; CUtensorMap map; bool ok; };
  static Entry cache[8];
  static int used = 0;
  const int dev = currentDevice();
  for (int n = 0; n < used; ++n)
    if (cache[n].base == U && cache[n].isize == P.isize && cache[n].jsize == P.jsize && cache[n].ksize == P.ksize &&
        cache[n].dev == dev)
      return cache[n].ok ? &cache[n].map : nullptr;
  Entry& e = cache[used < 8 ? used++ : (used = 1, 0)];
  e.base = U; e.isize = P.isize; e.jsize = P.jsize; e.ksize = P.ksize; e.dev = dev;
  // box rows start at cell gw - 2 + 28 bx: 16-byte aligned when (gw - 2) and 28 elements are
  e.ok = ((P.gw - 2) * sizeof(T)) % 16 == 0 && (C::UX * sizeof(T)) % 16 == 0 &&
         tma::encodeTile4D(&e.map, U, (int)sizeof(T), P.isize, P.jsize, P.ksize, 5, C::CX, C::CY);
  return e.ok ? &e.map : nullptr;
}

template <typename T, typename C>
void launchHydroFused(const KParams<T>& P, const T* Uold, T* Unew, int k0, int k1, T dt, unsigned long long* slots,
                      cudaStream_t s) {
  static bool attrSetDev[MAX_DEVICES] = {false};
  bool& attrSet = attrSetDev[currentDevice()];
  if (!attrSet) {
#define RG_HATTR(RS)                                                                                                      \
  cudaFuncSetAttribute(k_hydro_fused<T, RS, C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);      \
  cudaFuncSetAttribute(k_hydro_fused<T, RS, C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_TMA)
    RG_HATTR(RS_HLLC); RG_HATTR(RS_HLL); RG_HATTR(RS_APPROX); RG_HATTR(-1);
#undef RG_HATTR
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
  const CUtensorMap* mp = g_hydroTma ? hydroTensorMap<T, C>(P, Uold) : nullptr;
  if (mp != nullptr) {
    const CUtensorMap map = *mp;
    switch (P.riemannSolver) {  // one instantiation per Riemann solver: a single solver body in the kernel
      case RS_HLLC: k_hydro_fused<T, RS_HLLC, C, true><<<g, b, C::SMEM_TMA, s>>>(P, map, Uold, Unew, k0, k1, lz, dt, slots); break;
      case RS_HLL: k_hydro_fused<T, RS_HLL, C, true><<<g, b, C::SMEM_TMA, s>>>(P, map, Uold, Unew, k0, k1, lz, dt, slots); break;
      case RS_APPROX: k_hydro_fused<T, RS_APPROX, C, true><<<g, b, C::SMEM_TMA, s>>>(P, map, Uold, Unew, k0, k1, lz, dt, slots); break;
      default: k_hydro_fused<T, -1, C, true><<<g, b, C::SMEM_TMA, s>>>(P, map, Uold, Unew, k0, k1, lz, dt, slots); break;
    }
  } else {
    CUtensorMap map;
    memset(&map, 0, sizeof map);
    switch (P.riemannSolver) {
      case RS_HLLC: k_hydro_fused<T, RS_HLLC, C, false><<<g, b, C::SMEM, s>>>(P, map, Uold, Unew, k0, k1, lz, dt, slots); break;
      case RS_HLL: k_hydro_fused<T, RS_HLL, C, false><<<g, b, C::SMEM, s>>>(P, map, Uold, Unew, k0, k1, lz, dt, slots); break;
      case RS_APPROX: k_hydro_fused<T, RS_APPROX, C, false><<<g, b, C::SMEM, s>>>(P, map, Uold, Unew, k0, k1, lz, dt, slots); break;
      default: k_hydro_fused<T, -1, C, false><<<g, b, C::SMEM, s>>>(P, map, Uold, Unew, k0, k1, lz, dt, slots); break;
    }
  }
  launched();
}

}  // namespace

bool hydroFusedRequested() { return g_hydroFused != 0; }

template <typename T>
void HydroKernels<T>::fusedStep(const KParams<T>& P, const T* Uold, T* Unew, int k0, int k1, T dt, unsigned long long* slots,
                                cudaStream_t s) {
  if (k1 <= k0) return;
  // measured at 512^3 FP32 HLLC / 384^3 FP64 (profiles/r02_p_hydro_tma_ab.txt), ms per launch by rows of the block:
  //   TMA tiles        12: 7.26 / 5.65   16: 6.13 / 5.24   20: 5.60 / 6.56   24: 5.44 / 8.74
  //   per-thread loads 12: 7.81 / 6.24   16: 6.50 / 6.50   20: 5.96 / 7.26   24: 6.01 / 8.64   (FP64: spills from 20 rows on)
  const bool tmaTiles = g_hydroTma && ((size_t)P.isize * sizeof(T)) % 16 == 0 && ((P.gw - 2) * sizeof(T)) % 16 == 0 &&
                        (reinterpret_cast<uintptr_t>(Uold) & 15) == 0;
  const int rows = g_hydroRows ? g_hydroRows : (sizeof(T) == 4 ? (tmaTiles ? 24 : 20) : (tmaTiles ? 16 : 12));
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
