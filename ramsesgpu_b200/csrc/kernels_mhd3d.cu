// 3D ideal-MHD Godunov step for sm_100a: prim -> trace -> {fluxes, emfs} -> update(+CT, +next dt).
//
// What the reference does with 18 eight-component trace arrays in DRAM (reference
// MHDRunGodunov.cpp:219-241, mhd_godunov_unsplit_cpu_v3.cpp:172-361) is done here with ONE
// 47-component "traced state" W per cell from which every face/edge state is rebuilt with adds in
// the consumer kernels; electric field and magnetic slopes are never materialised.  All arrays are
// SoA with x fastest so that a warp reads/writes 32 consecutive reals per component.
//
// Index ranges follow the reference: prim 0..size-2, trace gw-1..size-gw,
// flux/emf/update gw..size-gw (inclusive), with the reference's write guards.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "kernels.h"
#include "kernel_common.cuh"
#include "mhd_cells.cuh"
#include "mhd_device.cuh"
#include "tma.cuh"

namespace rg {

unsigned long long g_launches = 0;
int g_tileX = 32;  // run-time knob "tile_x"; tile_y = BX / tile_x
// run-time knob "rot_dt": the rotating fused kernel reduces the next dt itself (1) or leaves it to k_invdt (0, default:
// 3.61 against 3.71 ms per step on the 256 x 512 x 64 shearing-box slab -- the smaller hot loop fits the instruction cache,
// profiles/r02_d_mri_ab.txt)
int g_rotDt = 0;
bool rotDtInKernel() { return g_rotDt != 0; }
extern int g_handoffHead;
extern int g_fusedHandoff;  // run-time knob "fused_handoff": 16 x 8 hand-off tiles (1, default) or 15 x 7 self-closing tiles (0)
int g_fusedB = 1;  // run-time knob "fused_b": 1 = fused flux+emf+update when available, 0 = separate kernels
bool fusedRequested() { return g_fusedB != 0; }
extern int g_fusedA;
int g_traceQY = 12;  // run-time knob "trace_qy": 12 (one 384-thread block per SM, 168 registers) or 8 (two 256-thread blocks, 128)

namespace {

// occupancy knobs (minimum resident blocks per SM the kernel is compiled for), see setTuning()
int g_fluxMinB = 5, g_emfMinB = 4, g_traceMinB = 3, g_updateMinB = 4;

// ------------------------------------------------------------------------------------------------
// K0: conservative -> primitive (reference MHDRunGodunov.cpp:538-560 + constoprim.h:137-199)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(BX) k_prim(const __grid_constant__ KParams<T> P, const T* __restrict__ Uin,
                                             T* __restrict__ Qp, int planes, int kbase, int k0, T dt) {
  int i, j;
  const int k = k0 + blockIdx.z;
  if (!tileCoords(0, P.isize - 1, 0, P.jsize - 1, i, j)) return;
  const UView<T> U = uview(Uin, P);
  const View<T> Q = view(Qp, P, planes, kbase);
  T u[8], q[8];
#pragma unroll
  for (int v = 0; v < 8; ++v) u[v] = U(v, i, j, k);
  dev::cons_to_prim_mhd(P, u, U(IA, i + 1, j, k), U(IB, i, j + 1, k), U(IC, i, j, k + 1), dt, q);
#pragma unroll
  for (int v = 0; v < 8; ++v) Q(v, i, j, k) = q[v];
}

// ------------------------------------------------------------------------------------------------
// K0b: edge-centred electric field E = v x B at the LOW edges of every cell
//      (reference cpu_v3.cpp:36-101, kernel_mhd_compute_elec_field): 4-cell average of the
//      velocities, 2-face average of the face fields
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(BX) k_elec(const __grid_constant__ KParams<T> P, const T* __restrict__ Uin,
                                             const T* __restrict__ Qp, T* __restrict__ ELp, int planes, int kbase,
                                             int k0) {
  int i, j;
  const int k = k0 + blockIdx.z;
  if (!tileCoords(1, P.isize - 2, 1, P.jsize - 2, i, j)) return;
  const UView<T> U = uview(Uin, P);
  const View<const T> Q = view<const T>(Qp, P, planes, kbase);
  const View<T> EL = view(ELp, P, planes, kbase);
  elec_cell<false>(P, Q, U, EL, i, j, k);
}

// ------------------------------------------------------------------------------------------------
// K1: slopes + edge electric fields + face-B slopes + half-step trace -> W
//     (reference cpu_v3.cpp:36-361, slope_mhd.h:436-502/598-704, trace_mhd.h:1854-2030)
// ------------------------------------------------------------------------------------------------
template <typename T, int MINB, bool FAST, bool S3 = false>
__global__ void __launch_bounds__(BX, MINB) k_trace(const __grid_constant__ KParams<T> P, const T* __restrict__ Uin,
                                                    const T* __restrict__ Qp, const T* __restrict__ ELp,
                                                    T* __restrict__ Wp, int planes, int kbase, int k0, T dt) {
  const int gw = P.gw;
  int i, j;
  const int k = k0 + blockIdx.z;
  if (!tileCoords(gw - 1, P.isize - 2 * gw + 2, gw - 1, P.jsize - 2 * gw + 2, i, j)) return;
  const UView<T> U = uview(Uin, P);
  const View<const T> Q = view<const T>(Qp, P, planes, kbase);
  const View<const T> EL = view<const T>(ELp, P, planes, kbase);
  const View<T> W = view(Wp, P, planes, kbase);
  trace_cell<FAST, S3>(P, Q, U, EL, W, i, j, k, dt);
}

// ------------------------------------------------------------------------------------------------
// KA: cons->prim + edge electric fields + slopes + trace in ONE kernel (FAST configuration): U -> W.
//
// A block of 32 x 8 threads owns the 32 x 8 tile of PRIMITIVE cells around a 30 x 6 tile of traced
// cells and marches along z.  Every thread converts one cell of the incoming plane (its conservative
// state was prefetched into registers during the previous plane), the tile of primitives and face
// fields lives in a 4-plane shared-memory ring, the edge electric fields in a 2-plane ring, and the
// inner 30 x 6 threads evaluate the 36 limited slopes + the half-step predictor from shared memory.
// Q and the edge electric field never go to HBM: per cell the kernel reads U once (x 1.4 tile halo)
// and writes the 38 W components.
// ------------------------------------------------------------------------------------------------
template <int QY_, int RINGQ_ = 4, int RINGE_ = 2>
struct TraceTileT {
  static constexpr int QX = 32, QY = QY_, TW = QX - 2, TH = QY - 2, QCELLS = QX * QY, RING = RINGQ_, RINGE = RINGE_;
  static constexpr int THREADS = QX * QY;
  static constexpr unsigned SMEM = (unsigned)((RING * 8 + RING * 3 + RINGE * 3) * QCELLS * sizeof(double));
  static constexpr int MINB = (512 / THREADS) * SMEM <= 227u * 1024u ? 512 / THREADS : 1;
};
template <typename T, typename TraceTile>
struct QTileView {  // primitives, ring of RING planes, [plane][var][QY][QX]
  T* buf;
  int ib, jb;
  __device__ __forceinline__ T& operator()(int v, int i, int j, int k) const {
    return buf[(((unsigned)k % (unsigned)TraceTile::RING) * 8 + v) * TraceTile::QCELLS + (j - jb) * TraceTile::QX + (i - ib)];
  }
};
template <typename T, typename TraceTile>
struct BTileView {  // face fields U(IA..IC), ring of RING planes
  T* buf;
  int ib, jb;
  __device__ __forceinline__ T& operator()(int v, int i, int j, int k) const {
    return buf[(((unsigned)k % (unsigned)TraceTile::RING) * 3 + (v - IA)) * TraceTile::QCELLS + (j - jb) * TraceTile::QX + (i - ib)];
  }
};
template <typename T, typename TraceTile>
struct ETileView {  // edge electric fields, ring of RINGE planes
  T* buf;
  int ib, jb;
  __device__ __forceinline__ T& operator()(int c, int i, int j, int k) const {
    return buf[(((unsigned)k % (unsigned)TraceTile::RINGE) * 3 + c) * TraceTile::QCELLS + (j - jb) * TraceTile::QX + (i - ib)];
  }
};

// (A one-barrier-per-plane variant -- 5-plane primitive ring, prim(k+3) | elec(k+2) | trace(k) per iteration -- was built
// and measured in round 2: 1.58 ms against 1.52 ms for this two-barrier pipeline at 256^3, profiles/r02_a_ab_trace_ring.txt;
// the modulo-5 ring indexing and the larger shared-memory footprint cost more than the second barrier.)
template <typename T, typename TraceTile, bool FAST>
__global__ void __launch_bounds__(TraceTile::THREADS, TraceTile::MINB)
k_fused_trace(const __grid_constant__ KParams<T> P, const T* __restrict__ Uin, T* __restrict__ Wp, int planes,
              int kbase, int k0, int k1, int lz, T dt) {
  extern __shared__ unsigned char smemRawA[];
  T* sm = reinterpret_cast<T*>(smemRawA);
  const int gw = P.gw;
  const int ti = threadIdx.x, tj = threadIdx.y;
  const int ib = gw - 2 + blockIdx.x * TraceTile::TW, jb = gw - 2 + blockIdx.y * TraceTile::TH;  // tile origin
  const int i = ib + ti, j = jb + tj;
  const int za = k0 + blockIdx.z * lz, zb = min(za + lz, k1);  // traced planes [za, zb)
  if (za >= zb) return;
  const QTileView<T, TraceTile> Q{sm, ib, jb};
  const BTileView<T, TraceTile> B{sm + TraceTile::RING * 8 * TraceTile::QCELLS, ib, jb};
  const ETileView<T, TraceTile> EL{sm + TraceTile::RING * 11 * TraceTile::QCELLS, ib, jb};
  const UView<T> U = uview(Uin, P);
  const View<T> W = view(Wp, P, planes, kbase);

  // prim range of the reference: 0 .. size-2 (needs the +1 faces)
  const bool primOK = i <= P.isize - 2 && j <= P.jsize - 2;
  const bool elecOK = primOK && ti >= 1 && tj >= 1;
  const bool traceOK = ti >= 1 && ti <= TraceTile::TW && tj >= 1 && tj <= TraceTile::TH && i <= P.isize - gw &&
                       j <= P.jsize - gw;
  T u[8], ap = T(0), bp = T(0), cp = T(0);
#pragma unroll
  for (int v = 0; v < 8; ++v) u[v] = T(1);
  auto load = [&](int q) {  // conservative state of plane q -> registers
    if (primOK && q <= zb && q <= P.ksize - 2) {
#pragma unroll
      for (int v = 0; v < 8; ++v) u[v] = U(v, i, j, q);
      ap = U(IA, i + 1, j, q);
      bp = U(IB, i, j + 1, q);
      cp = U(IC, i, j, q + 1);
    }
  };
  auto prim = [&](int q) {  // registers -> primitives + face fields of plane q in the rings
    T qv[8];
    dev::cons_to_prim_mhd<FAST>(P, u, ap, bp, cp, dt, qv);
#pragma unroll
    for (int v = 0; v < 8; ++v) Q(v, i, j, q) = qv[v];
    B(IA, i, j, q) = u[IA];
    B(IB, i, j, q) = u[IB];
    B(IC, i, j, q) = u[IC];
  };
  load(za - 1);
  prim(za - 1);
  load(za);
  prim(za);
  load(za + 1);
  __syncthreads();
  if (elecOK) elec_cell<FAST>(P, Q, B, EL, i, j, za);
  for (int k = za; k < zb; ++k) {
    prim(k + 1);
    load(k + 2);  // prefetch: consumed by the next iteration, in flight during the trace below
    __syncthreads();
    if (elecOK) elec_cell<FAST>(P, Q, B, EL, i, j, k + 1);
    __syncthreads();
    if (traceOK) trace_cell<FAST>(P, Q, B, EL, W, i, j, k, dt);
  }
}

// ------------------------------------------------------------------------------------------------
// K2: HLLD (or HLL/LLF) fluxes at the three low faces (reference cpu_v3.cpp:397-465, trace_mhd.h:2032-2102)
//     face state = W cell-centred value +/- half slope along the normal, floors on rho and p
// ------------------------------------------------------------------------------------------------
template <typename T, int DIR, int MINB, bool FAST>
__global__ void __launch_bounds__(BX, MINB) k_flux(const __grid_constant__ KParams<T> P, const T* __restrict__ Wp,
                                             T* __restrict__ Fp, int planes, int kbase, int k0) {
  const int gw = P.gw;
  int i, j;
  const int k = k0 + blockIdx.z;
  if (!tileCoords(gw, P.nx + 1, gw, P.ny + 1, i, j)) return;
  // a face is only needed where both transverse indexes are inner
  if (DIR != 0 && i >= P.isize - gw) return;
  if (DIR != 1 && j >= P.jsize - gw) return;
  if (DIR != 2 && k >= P.ksize - gw) return;
  const View<const T> W = view<const T>(Wp, P, planes, kbase);
  const View<T> F = view(Fp, P, planes, kbase);
  flux_cell<T, DIR, FAST>(P, W, F, i, j, k);
}

// ------------------------------------------------------------------------------------------------
// K3: corner emfs with the 2-D HLLD solver (reference cpu_v3.cpp:539-579, trace_mhd.h:2104-2246,
//     riemann_mhd.h:1054-1193).  EDIR = 2: emf_z from Z-edge states of (i-1,j-1),(i-1,j),(i,j-1),(i,j)
// ------------------------------------------------------------------------------------------------
template <typename T, int EDIR, int MINB, bool FAST>
__global__ void __launch_bounds__(BX, MINB) k_emf(const __grid_constant__ KParams<T> P, const T* __restrict__ Wp,
                                            T* __restrict__ Ep, int planes, int kbase, int k0) {
  const int gw = P.gw;
  int i, j;
  const int k = k0 + blockIdx.z;
  if (!tileCoords(gw, P.nx + 1, gw, P.ny + 1, i, j)) return;
  const View<const T> W = view<const T>(Wp, P, planes, kbase);
  const View<T> E = view(Ep, P, planes, kbase);
  emf_cell<T, EDIR, FAST>(P, W, E, i, j, k);
}

// ------------------------------------------------------------------------------------------------
// K4: conservative update + constrained transport + inverse-dt reduction of the NEW state
//     (reference cpu_v3.cpp:475-533 and :600-630; dt: MHDRunBase.cpp:141-250)
// ------------------------------------------------------------------------------------------------
template <typename T, int MINB, bool FAST>
__global__ void __launch_bounds__(BX, MINB) k_update(const __grid_constant__ KParams<T> P, const T* __restrict__ Uold,
                                               T* __restrict__ Unew, const T* __restrict__ Fp,
                                               const T* __restrict__ Ep, int planes, int kbase, int k0, T dt,
                                               unsigned long long* __restrict__ dMaxInvDt) {
  const int gw = P.gw;
  int i, j;
  const int k = k0 + blockIdx.z;
  const bool valid = tileCoords(0, P.isize, 0, P.jsize, i, j);
  const int iN = P.isize - gw, jN = P.jsize - gw, kN = P.ksize - gw;  // first upper ghost index
  T invDt = T(0);
  if (valid) {
    const UView<T> U = uview(Uold, P);
    const bool inBox = i >= gw && i <= iN && j >= gw && j <= jN && k >= gw && k <= kN;
    if (!inBox) {
      const size_t plane = (size_t)P.isize * P.jsize, comp = plane * P.ksize;
      const size_t idx = (size_t)k * plane + (size_t)j * P.isize + i;
#pragma unroll
      for (int v = 0; v < 8; ++v) Unew[v * comp + idx] = U(v, i, j, k);
    } else {
      const View<const T> F = view<const T>(Fp, P, planes, kbase);
      const View<const T> E = view<const T>(Ep, P, planes, kbase);
      invDt = update_cell<FAST>(P, U, Unew, F, E, i, j, k, dt);
    }
  }
  if (dMaxInvDt != nullptr) reduceMaxToSlots(invDt, dMaxInvDt);
}

// ------------------------------------------------------------------------------------------------
// KF: fluxes + corner emfs + conservative/CT update + next dt in ONE kernel (FAST configuration).
//
// A thread block (one per SM) owns a TW x TH column of cells and marches along z.  Per plane, ONE TMA
// bulk-tensor copy brings the (TW+2) x (TH+2) tile of all 38 W components into shared memory (ring of
// three plane buffers, the copy of plane p+1 is issued when plane p starts).  The 6 x (TW+1) x (TH+1)
// face / edge Riemann problems of a plane and the update of the plane below are warp-sized TASKS
// handed out in a fixed order by a ticket counter; a task waits (spinning on shared-memory counters
// and on the TMA mbarrier of its plane) only for the earlier tasks it really depends on, so there is
// no block-wide barrier in the march and warps flow from one plane into the next.  Face fluxes and
// corner emfs live in a three-plane shared-memory ring and never go to HBM; W is read from HBM once
// per step instead of once per kernel (3 flux + 3 emf kernels), and global-load latency is off the
// critical path.
//   task order of plane p: TMA(p+1) | emf_x emf_y flux_z (need W(p-1), W(p)) | emf_z flux_x flux_y
//   (need W(p)) | update of plane p-1 (needs the z group of p and everything of p-1)
//   run-ahead gate: a task of plane p starts only when plane p-2 is complete (ring safety)
// The arithmetic is the one of k_flux / k_emf / k_update (same device functions).
// ------------------------------------------------------------------------------------------------
// HANDOFF_ = true (knob "fused_handoff" = 1; see the measurement at g_fusedHandoff): a tile owns TW x TH = 16 x 8 CELLS and solves exactly the 16 x 8 low faces / edges of its
// own cells; the faces / edges that close its last column and last row (x-face fluxes and emf_z, emf_y of column i0+16;
// y-face fluxes and emf_z, emf_x of row j0+8) are NOT solved a second time: the tile to the right / above publishes
// them plane by plane in a small HBM buffer (168 reals per tile and plane) and this tile's update reads them there.
// Tiles take their index from an atomic counter in "right to left, top to bottom" order, so that a tile only ever waits
// for tiles whose blocks are already running (no assumption on the block scheduler).  Every Riemann problem of the grid
// is solved once: 128 cells per 128 solves instead of 105.  HANDOFF_ = false (default): the 15 x 7 tile that solves its
// closing column / row itself.
template <typename T, int TW_, int TH_, int THREADS_, bool HANDOFF_ = false>
struct FusedTile {
  static constexpr int TW = TW_, TH = TH_, THREADS = THREADS_;
  static constexpr bool HANDOFF = HANDOFF_;
  static constexpr int PX = HANDOFF ? TW : TW + 1, PY = HANDOFF ? TH : TH + 1;  // faces / edges solved by the tile
  // W tile with one halo cell on the low sides.  TMA wants the first element of a box row on a 16-byte
  // boundary: the box starts at the even cell index at or below i0-1 and is PX+2 (rounded up to even)
  // reals wide
  static constexpr int WX = (PX + 2) / 2 * 2, WY = PY + 1, WCELLS = WX * WY;  // i0-1 is even (gw = 3)
  // a warp task covers two rows of 16 positions: every half warp reads 16 consecutive reals of one
  // tile row, which keeps the 64-bit shared-memory loads free of bank conflicts
  static constexpr int PXP = 16, NCH = PY / 2;
  static_assert(PX <= PXP && PY % 2 == 0, "tile shape");
  // flux / emf ring in shared memory: the tile's own PX x PY positions; a HANDOFF tile keeps one more column and row for
  // the values imported from its neighbours, so that the update indexes one array without any branch
  static constexpr int FEX = HANDOFF ? PXP + 1 : PXP, FEY = HANDOFF ? PY + 1 : PY, NPOSP = FEX * FEY;
  static constexpr int NFE = 18;                                     // flux_x[5] flux_y[5] flux_z[5] emf z,y,x
  static constexpr int NT = 1 + 7 * NCH + (HANDOFF ? 4 : 0);         // tasks per plane (+ the four hand-off helpers)
  // HANDOFF ticket order of a plane p: TMA | import, then publish, the plane-local group of plane p-1 | z group | half
  // of the plane-local group | publish, then import, the z group of plane p | rest of the plane-local group | updates of
  // plane p-1.  The solver tasks are exactly those of the self-closing tile (no global store, no extra atomic); ONE
  // helper warp per group copies the tile's first row / column from the ring to the record and releases the flag, one
  // imports the neighbours' records.  Helpers sit several solver tasks behind what they wait for and ahead of what
  // waits for them, so their L2 round trips are hidden.
  static constexpr int T_IMPORT_XY = 1, T_PUBLISH_XY = 2, T_FIRST = 3;
  static constexpr int T_PUBLISH_Z = T_FIRST + 3 * NCH + (3 * NCH) / 2, T_IMPORT_Z = T_PUBLISH_Z + 1;
  static constexpr int LZMAX = 160;                                  // planes per block (counter arrays)
  static constexpr unsigned W_BYTES = (unsigned)(NW_MHD * WCELLS * sizeof(T));
  static constexpr unsigned W_STRIDE = (W_BYTES + 127u) / 128u * 128u;
  static constexpr unsigned FE_SLOT = (unsigned)(NFE * NPOSP);       // reals per plane slot
  static constexpr unsigned FE_BYTES = (unsigned)(3 * FE_SLOT * sizeof(T));
  static constexpr unsigned NBAR = LZMAX + 2, NCNT = LZMAX + 4;
  static constexpr unsigned SMEM = 128u + 3u * W_STRIDE + FE_BYTES + NBAR * 8u + 4u * NCNT * 4u + 32u;
  // hand-off record of one tile and plane: row part [7][16] (flux_y[5], emf_z, emf_x of the tile's FIRST row), then
  // column part [7][8] (flux_x[5], emf_z, emf_y of its FIRST column)
  static constexpr int HROW = 7 * PXP, HREC = HROW + 7 * PY;
};

template <typename T, typename C>
struct WTileView {  // plane k of the W tile lives in ring buffer k % 3, laid out [comp][WY][WX]
  const unsigned char* buf;
  int ib, jb;  // cell index of tile element (0, 0)
  __device__ __forceinline__ T operator()(int c, int i, int j, int k) const {
    const T* b = reinterpret_cast<const T*>(buf + ((unsigned)k % 3u) * C::W_STRIDE);
    return b[c * C::WCELLS + (j - jb) * C::WX + (i - ib)];
  }
};
template <typename T, typename C>
struct FETileView {  // fluxes (c0 = 0) or emfs (c0 = 15) of plane k (ring slot k % 3), [comp][PY][PXP]
  T* buf;
  int i0, j0, c0;
  __device__ __forceinline__ T& operator()(int c, int i, int j, int k) const {
    return buf[((unsigned)k % 3u) * C::FE_SLOT + (c0 + c) * C::NPOSP + (j - j0) * C::FEX + (i - i0)];
  }
};

// progress flags of the hand-off in global memory: release-store by the publishing tile, acquire-poll by its consumers
__device__ __forceinline__ void publishProgress(int* flag, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(v) : "memory");
}
__device__ __forceinline__ void waitProgress(const int* flag, int need) {
  int v;
  for (unsigned spins = 0;; ++spins) {
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v >= need) break;
    __nanosleep(200);
    if (spins > (1u << 25)) __trap();  // several seconds without progress: fail loudly instead of hanging the device
  }
}

// completion counters in shared memory: release-add by the finishing warp, acquire-poll by waiters
__device__ __forceinline__ void signalCount(int* c) {
  asm volatile("red.release.cta.shared::cta.add.s32 [%0], 1;" ::"r"(tma::smemAddr(c)) : "memory");
}
__device__ __forceinline__ void waitCount(const int* c, int full) {
  int v;
  for (;;) {
    asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(tma::smemAddr(c)) : "memory");
    if (v >= full) break;
    __nanosleep(64);
  }
}

// FAST = false: the rotating frame / shearing box with the isothermal closure (BASELINE.json configs[3]): same tasks with
// the shear terms of the y flux and of the emfs, update_cell_rot for the cells of the tile, and -- with shearing-box
// boundaries -- the fluxes / emfs of the four x-border position columns copied to compact strips in HBM, from which
// k_update_rot_border updates the three cell columns that read the y-remapped OPPOSITE border (other tiles' data).
// ROTDT (FAST = false only): the inverse dt of the new state inside the update (true) or left to k_invdt (false: less
// code in the hot loop, which has to fit the instruction cache; knob "rot_dt")
template <typename T, typename C, bool FAST, bool ROTDT = true>
__global__ void __launch_bounds__(C::THREADS, 1)
k_fused_flux_emf_update(const __grid_constant__ KParams<T> P, const __grid_constant__ CUtensorMap mapW,
                        const T* __restrict__ Uold, T* __restrict__ Unew, int kbase, int ka, int kb, int lz, T dt,
                        unsigned long long* __restrict__ dMaxInvDt, const ShearShift<T> sh, T* __restrict__ strips,
                        int stripPlanes, T* __restrict__ hbuf, int* __restrict__ hsync, int ntx, int nty, int hplanes,
                        int headPlanes) {
  extern __shared__ unsigned char smemRaw[];
  // 128-byte alignment for the TMA destination, computed on the shared-window address so that the
  // compiler keeps the shared address space (LDS/STS, not generic LD/ST)
  unsigned char* sm = smemRaw + ((128u - (tma::smemAddr(smemRaw) & 127u)) & 127u);
  unsigned char* wbuf = sm;
  T* fe = reinterpret_cast<T*>(sm + 3 * C::W_STRIDE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 3 * C::W_STRIDE + C::FE_BYTES);  // bars[q - (za-1)]
  int* cntAll = reinterpret_cast<int*>(bars + C::NBAR);  // cnt*[pl + 2], pl = plane - za
  int* cntZ = cntAll + C::NCNT;
  int* cntXY = cntZ + C::NCNT;
  int* cntImp = cntXY + C::NCNT;
  int* ticket = cntImp + C::NCNT;
  int* pubSeq = ticket + 1;  // [0] z-group, [1] plane-local group: planes published so far, in order

  const int gw = P.gw;
  const int iN = P.isize - gw, jN = P.jsize - gw, kN = P.ksize - gw;
  // tile of this block: from the block index, or (HANDOFF) from an atomic counter in right-to-left, top-to-bottom order:
  // the producers of tile n are tiles n-1 (right), n-ntx (above) and n-ntx-1, whose blocks took their number earlier
  int tbx = blockIdx.x, tby = blockIdx.y, tbz = blockIdx.z, tileId = 0;
  if (C::HANDOFF) {
    if (threadIdx.x == 0) *ticket = atomicAdd(hsync, 1);
    __syncthreads();
    tileId = *ticket;
    __syncthreads();
    const int r = tileId / ntx;
    tbx = ntx - 1 - (tileId - r * ntx);
    tbz = r / nty;
    tby = nty - 1 - (r - tbz * nty);
  }
  const int i0 = gw + tbx * C::TW, j0 = gw + tby * C::TH;
  const int za = ka + tbz * lz, zb = min(za + lz, kb);  // update planes [za, zb) of this block
  if (za >= zb) return;
  const int fhi = min(zb, kN);                       // last plane whose low faces / edges are needed
  const int nPl = fhi - za + 1 + (zb > fhi ? 1 : 0);  // + the pseudo plane that only updates plane kN
  const int tid = threadIdx.x, lane = tid & 31;

  const int ib = (i0 - 1) & ~1;  // even: 16-byte aligned box rows
  const WTileView<T, C> W{wbuf, ib, j0 - 1};
  const FETileView<T, C> F{fe, i0, j0, 0}, E{fe, i0, j0, 15};
  const UView<T> U = uview(Uold, P);
  // hand-off records: mine (written), and those of the producers (read by the update)
  const bool hasRight = C::HANDOFF && tbx < ntx - 1, hasAbove = C::HANDOFF && tby < nty - 1;
  T* const recMine = C::HANDOFF ? hbuf + (size_t)tileId * hplanes * C::HREC : nullptr;
  const T* const recRight = hasRight ? hbuf + (size_t)(tileId - 1) * hplanes * C::HREC : nullptr;
  const T* const recAbove = hasAbove ? hbuf + (size_t)(tileId - ntx) * hplanes * C::HREC : nullptr;
  const T* const recCorner = (hasRight && hasAbove) ? hbuf + (size_t)(tileId - ntx - 1) * hplanes * C::HREC : nullptr;
  int* const prog = hsync + 1;  // prog[2 tile + g] = planes whose group g (0: emf_x, emf_y; 1: emf_z, flux_x, flux_y) is out

  for (int n = tid; n < (int)C::NBAR; n += C::THREADS) tma::mbarInit(&bars[n], 1);
  for (int n = tid; n < (int)C::NCNT; n += C::THREADS) {
    cntAll[n] = (n < 2) ? C::NT : 0;      // planes za-2, za-1 count as complete
    cntZ[n] = (n < 2) ? 3 * C::NCH : 0;
    cntXY[n] = (n < 2) ? 3 * C::NCH : 0;
    cntImp[n] = (n < 2) ? 2 : 0;
  }
  if (tid == 0) {
    *ticket = 0;
    pubSeq[0] = 0;
    pubSeq[1] = 0;
  }
  tma::fenceBarrierInit();
  __syncthreads();
  auto loadPlane = [&](int q) {  // one thread
    uint64_t* bar = &bars[q - (za - 1)];
    tma::mbarExpectTx(bar, C::W_BYTES);
    tma::loadTile4D(wbuf + ((unsigned)q % 3u) * C::W_STRIDE, &mapW, bar, ib, j0 - 1, q - kbase, 0);
  };
  if (tid == 0) {
    loadPlane(za - 1);
    loadPlane(za);
  }
  if (C::HANDOFF) {
    // head start of the producers: a tile begins when the tiles to its right and above have published their first
    // HEAD planes, so that from then on its import tasks find the records they need already there (tiles do equal work
    // and keep their distance; the consumer can never overtake).  One-time skew of a wave of tiles: (columns + rows of
    // the wave) x HEAD planes, about 2 % of a 130-plane march; without it every import waits about one task for the
    // producer running in lockstep and the four updates behind it wait too.
    const int head = min(headPlanes, nPl);
    if (tid == 0 && hasRight) waitProgress(&prog[2 * (tileId - 1) + 1], head);
    if (tid == 32 && hasAbove) waitProgress(&prog[2 * (tileId - ntx) + 1], head);
    __syncthreads();
  }

  for (;;) {
    int tk = 0;
    if (lane == 0) tk = atomicAdd(ticket, 1);
    tk = __shfl_sync(0xffffffffu, tk, 0);
    if (tk >= nPl * C::NT) break;
    const int pl = tk / C::NT, t = tk - pl * C::NT, p = za + pl;
    waitCount(&cntAll[pl], C::NT);  // plane p-2 complete: ring slots p % 3 of W and F/E are free
    if (t == 0) {                   // producer task: W(p+1) -> ring buffer of plane p-2
      waitCount(&cntZ[pl + 1], 3 * C::NCH);  // the z group of plane p-1 was the last reader of W(p-2)
      if (lane == 0 && p + 1 <= fhi) {
        tma::fenceProxyAsync();
        loadPlane(p + 1);
      }
      __syncwarp();
      if (lane == 0) signalCount(&cntAll[pl + 2]);
      continue;
    }
    if (C::HANDOFF && (t == C::T_PUBLISH_XY || t == C::T_PUBLISH_Z)) {
      // publish tasks: when the group's solver tasks of the plane are done, copy the tile's FIRST row and column from the
      // ring to the record of the plane and release-store the progress flag (in plane order).
      // XY: flux_y, emf_z of row 0 | flux_x, emf_z of column 0, of plane p-1;  Z: emf_x of row 0, emf_y of column 0, of plane p
      const bool isZ = t == C::T_PUBLISH_Z;
      const int q = isZ ? pl : pl - 1, g = isZ ? 0 : 1;
      if (q >= 0) {
        waitCount(isZ ? &cntZ[q + 2] : &cntXY[q + 2], 3 * C::NCH);
        const T* slot = fe + ((unsigned)(za + q) % 3u) * C::FE_SLOT;
        T* rec = recMine + (size_t)q * C::HREC;
        if (isZ) {
          if (lane < C::PXP) rec[6 * C::PXP + lane] = slot[17 * C::NPOSP + lane];
          else if (lane < C::PXP + C::PY) rec[C::HROW + 6 * C::PY + lane - C::PXP] = slot[16 * C::NPOSP + (lane - C::PXP) * C::FEX];
        } else {
          for (int n = lane; n < 6 * C::PXP; n += 32) {
            const int comp = n / C::PXP, x = n - comp * C::PXP;
            rec[n] = slot[(comp < 5 ? 5 + comp : 15) * C::NPOSP + x];
          }
          for (int n = lane; n < 6 * C::PY; n += 32) {
            const int comp = n / C::PY, y = n - comp * C::PY;
            rec[C::HROW + n] = slot[(comp < 5 ? comp : 15) * C::NPOSP + y * C::FEX];
          }
        }
        __threadfence();  // every lane's record entries are visible device-wide before the flag moves
        __syncwarp();
        if (lane == 0) {
          waitCount(&pubSeq[g], q);  // a flag counts CONSECUTIVE published planes (a block runs up to two planes ahead)
          publishProgress(&prog[2 * tileId + g], q + 1);
          signalCount(&pubSeq[g]);
        }
      }
      __syncwarp();
      if (lane == 0) signalCount(&cntAll[pl + 2]);
      continue;
    }
    if (C::HANDOFF && (t == C::T_IMPORT_XY || t == C::T_IMPORT_Z)) {
      // import tasks: copy what the update of plane p-1 misses into the extra column / row of the flux / emf ring.
      // XY: flux_y, emf_z of the closing row | flux_x, emf_z of the closing column | corner emf_z, of plane p-1;
      // Z: emf_x of the closing row, emf_y of the closing column, of plane p.
      const bool isZ = t == C::T_IMPORT_Z;
      const int need = isZ ? pl + 1 : pl, g = isZ ? 0 : 1;
      if (need >= 1) {
        // lanes 0..2 poll the flags of the tile to the right, above, and above-right at the same time
        const int prod = lane == 0 ? tileId - 1 : (lane == 1 ? tileId - ntx : tileId - ntx - 1);
        const bool poll = (lane == 0 && hasRight) || (lane == 1 && hasAbove) || (lane == 2 && hasRight && hasAbove && !isZ);
        if (poll) waitProgress(&prog[2 * prod + g], need);
        __syncwarp();
        const int ps = isZ ? p : p - 1, pls = isZ ? pl : pl - 1;  // plane (and its index) whose values are imported
        T* slot = fe + ((unsigned)(ps + 3) % 3u) * C::FE_SLOT;
        // record component -> ring component: row part flux_y[5] (5..9), emf_z (15), emf_x (17); column part flux_x[5]
        // (0..4), emf_z (15), emf_y (16)
        if (hasAbove) {
          const T* rec = recAbove + (size_t)pls * C::HREC;
          if (isZ) {
            if (lane < C::PXP) slot[17 * C::NPOSP + C::PY * C::FEX + lane] = __ldcg(rec + 6 * C::PXP + lane);
          } else {
            for (int n = lane; n < 6 * C::PXP; n += 32) {
              const int comp = n / C::PXP, x = n - comp * C::PXP;
              slot[(comp < 5 ? 5 + comp : 15) * C::NPOSP + C::PY * C::FEX + x] = __ldcg(rec + n);
            }
          }
        }
        if (hasRight) {
          const T* rec = recRight + (size_t)pls * C::HREC + C::HROW;
          if (isZ) {
            if (lane >= 16 && lane < 16 + C::PY) slot[16 * C::NPOSP + (lane - 16) * C::FEX + C::PXP] = __ldcg(rec + 6 * C::PY + lane - 16);
          } else {
            for (int n = lane; n < 6 * C::PY; n += 32) {
              const int comp = n / C::PY, y = n - comp * C::PY;
              slot[(comp < 5 ? comp : 15) * C::NPOSP + y * C::FEX + C::PXP] = __ldcg(rec + n);
            }
          }
        }
        if (!isZ && hasRight && hasAbove && lane == 0)
          slot[15 * C::NPOSP + C::PY * C::FEX + C::PXP] = __ldcg(recCorner + (size_t)pls * C::HREC + 5 * C::PXP);
      }
      __syncwarp();
      if (lane == 0) {
        signalCount(&cntImp[pl + 2]);
        signalCount(&cntAll[pl + 2]);
      }
      continue;
    }
    // solver / update task index without the helper tasks
    const int ts = !C::HANDOFF ? t - 1 : (t > C::T_IMPORT_Z ? t - C::T_FIRST - 2 : t - C::T_FIRST);
    const int kind = ts / C::NCH, chunk = ts - kind * C::NCH;
    const int pi = lane & 15, pj = chunk * 2 + (lane >> 4);
    const int i = i0 + pi, j = j0 + pj;
    const bool ok = pi < C::PX && i <= iN && j <= jN;
    if (kind < 6) {
      // kinds 0..2 = z group (emf_x, emf_y, flux_z), 3..5 = plane-local group (emf_z, flux_x, flux_y; not on
      // the closing plane of a z range).  ONE call site per solver keeps a single copy of each in the
      // instruction cache; the direction only selects which W components are gathered.
      const bool zgrp = kind < 3;
      const bool isEmf = kind == 0 || kind == 1 || kind == 3;
      const int dir = isEmf ? (kind == 3 ? 2 : kind) : (kind == 2 ? 2 : kind - 4);
      if (p <= fhi && (zgrp || p < zb)) {
        if (zgrp) tma::mbarWait(&bars[pl], 0);
        tma::mbarWait(&bars[pl + 1], 0);
        const bool strip = !FAST && sh.enabled && BorderView<T>::holds(i, gw, P.nx);
        if (isEmf) {
          if (ok) {
            fused_emf_task<FAST>(P, W, E, dir, i, j, p);
            if (strip) {
              const BorderView<T> Eb{strips, P.jsize, stripPlanes, kbase, gw, P.nx, 15};
              Eb(2 - dir, i, j, p) = E(2 - dir, i, j, p);
            }
          }
        } else {
          // a face is only needed where both transverse indexes are inner (k_flux)
          const bool need = (dir == 0 || i < iN) && (dir == 1 || j < jN) && (dir == 2 || p < kN);
          if (ok && need) {
            fused_flux_task<FAST>(P, W, F, dir, i, j, p);
            if (strip) {
              const BorderView<T> Fb{strips, P.jsize, stripPlanes, kbase, gw, P.nx, 0};
#pragma unroll
              for (int c = 0; c < 5; ++c) Fb(5 * dir + c, i, j, p) = F(5 * dir + c, i, j, p);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) {
        signalCount(zgrp ? &cntZ[pl + 2] : &cntXY[pl + 2]);
        signalCount(&cntAll[pl + 2]);
      }
    } else {  // update of plane p-1
      if (p - 1 >= za) {
        waitCount(&cntZ[pl + 2], 3 * C::NCH);
        waitCount(&cntZ[pl + 1], 3 * C::NCH);
        waitCount(&cntXY[pl + 1], 3 * C::NCH);
        if (C::HANDOFF) {  // closing column / row imported: z group of planes p-1 and p, plane-local group of plane p-1
          waitCount(&cntImp[pl + 2], 2);
          waitCount(&cntImp[pl + 1], 2);
        }
        __syncwarp();
        // legacy tile: a cell of the closing column/row belongs to this tile only when it is the ghost face (iN / jN)
        const bool mine = C::HANDOFF ? ok : ok && (i < i0 + C::TW || i == iN) && (j < j0 + C::TH || j == jN);
        T invDt = T(0);
        if (mine) {
          if (FAST) {
            invDt = update_cell<true>(P, U, Unew, F, E, i, j, p - 1, dt);
          } else if (!(sh.enabled && (i == gw || i == P.nx + gw - 1 || i == P.nx + gw))) {
            // (the three cell columns next to a shearing x border are updated by k_update_rot_border)
            invDt = update_cell_rot<false, ROTDT>(P, U, Unew, F, E, F, E, i, j, p - 1, dt, sh);
          }
        }
        if (dMaxInvDt != nullptr) reduceMaxToSlots(invDt, dMaxInvDt);
      }
      __syncwarp();
      if (lane == 0) signalCount(&cntAll[pl + 2]);
    }
  }
}

// ghost cells outside the update box keep the old values (the separate k_update does this itself).
// One thread per ghost cell of a plane: the gw lower and gw-1 upper ghost rows (full width), then the
// gw lower and gw-1 upper ghost columns of the remaining rows.
template <typename T>
__global__ void __launch_bounds__(256) k_copy_outside_box(const __grid_constant__ KParams<T> P,
                                                          const T* __restrict__ Uold, T* __restrict__ Unew, int k0) {
  const int gw = P.gw, ng = 2 * gw - 1;  // ghost rows / columns outside the box per direction
  const int nRowCells = ng * P.isize, nColCells = ng * (P.jsize - ng);
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nRowCells + nColCells) return;
  int i, j;
  if (t < nRowCells) {
    const int r = t / P.isize;
    i = t - r * P.isize;
    j = (r < gw) ? r : P.jsize - gw + 1 + (r - gw);
  } else {
    const int q = t - nRowCells, r = q / ng, c = q - r * ng;
    j = gw + r;
    i = (c < gw) ? c : P.isize - gw + 1 + (c - gw);
  }
  const int k = k0 + blockIdx.y;
  const size_t plane = (size_t)P.isize * P.jsize, comp = plane * P.ksize;
  const size_t idx = (size_t)k * plane + (size_t)j * P.isize + i;
  for (int v = 0; v < P.nvar; ++v) Unew[v * comp + idx] = Uold[v * comp + idx];
}

// ------------------------------------------------------------------------------------------------
// K4r: update in the rotating frame / shearing box (reference MHDRunGodunov.cpp:2938-3348):
//   Crank-Nicolson Coriolis rotation of (rho u, rho v), alpha-mixed momentum fluxes, density flux
//   and emf_y of the two x borders averaged with the y-remapped opposite border (gathered here
//   instead of the reference's border strips), density floor on the border columns, CT, next dt
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(BX, 3) k_update_rot(const __grid_constant__ KParams<T> P, const T* __restrict__ Uold,
                                                      T* __restrict__ Unew, const T* __restrict__ Fp,
                                                      const T* __restrict__ Ep, int planes, int kbase, int k0, T dt,
                                                      const ShearShift<T> sh, unsigned long long* __restrict__ dMaxInvDt) {
  const int gw = P.gw;
  int i, j;
  const int k = k0 + blockIdx.z;
  const bool valid = tileCoords(0, P.isize, 0, P.jsize, i, j);
  const int iN = P.isize - gw, jN = P.jsize - gw, kN = P.ksize - gw;
  T invDt = T(0);
  if (valid) {
    const UView<T> U = uview(Uold, P);
    const bool inBox = i >= gw && i <= iN && j >= gw && j <= jN && k >= gw && k <= kN;
    if (!inBox) {
      const size_t plane = (size_t)P.isize * P.jsize, comp = plane * P.ksize;
      const size_t idx = (size_t)k * plane + (size_t)j * P.isize + i;
#pragma unroll
      for (int v = 0; v < 8; ++v) Unew[v * comp + idx] = U(v, i, j, k);
    } else {
      const View<const T> F = view<const T>(Fp, P, planes, kbase);
      const View<const T> E = view<const T>(Ep, P, planes, kbase);
      invDt = update_cell_rot(P, U, Unew, F, E, F, E, i, j, k, dt, sh);
    }
  }
  if (dMaxInvDt != nullptr) reduceMaxToSlots(invDt, dMaxInvDt);
}

// The three cell columns of a shearing box whose update reads the y-remapped opposite x border (i = gw, nx+gw-1 and the
// ghost-face column nx+gw), after the fused kernel has left the fluxes / emfs of the four border position columns in the
// compact strips (BorderView): thread = (column, row j), plane = blockIdx.z
template <typename T>
__global__ void __launch_bounds__(128) k_update_rot_border(const __grid_constant__ KParams<T> P, const T* __restrict__ Uold,
                                                           T* __restrict__ Unew, const T* __restrict__ strips, int planes,
                                                           int kbase, int k0, T dt, const ShearShift<T> sh,
                                                           unsigned long long* __restrict__ dMaxInvDt) {
  const int gw = P.gw, jN = P.jsize - gw;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int nRows = jN - gw + 1;  // rows gw .. jN of the update box
  const int c = t / nRows, j = gw + t - c * nRows;
  const int k = k0 + blockIdx.z;
  T invDt = T(0);
  if (c < 3) {
    const int i = (c == 0) ? gw : P.nx + gw - 2 + c;  // gw, nx+gw-1, nx+gw
    const UView<T> U = uview(Uold, P);
    const BorderView<const T> F{strips, P.jsize, planes, kbase, gw, P.nx, 0}, E{strips, P.jsize, planes, kbase, gw, P.nx, 15};
    invDt = update_cell_rot(P, U, Unew, F, E, F, E, i, j, k, dt, sh);
  }
  if (dMaxInvDt != nullptr) reduceMaxToSlots(invDt, dMaxInvDt);
}

// ------------------------------------------------------------------------------------------------
// shearing-box ghost cells in x (reference MHDRunGodunov.cpp:3539-3759): every x-ghost cell of an
// inner row is interpolated from the opposite border's inner columns, shifted along y by
// deltay(t+dt), second order with limited y slopes (B_y: first-order difference slope)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_shear_ghosts(const __grid_constant__ KParams<T> P, T* __restrict__ U, const ShearShift<T> sh, int k0) {
  const int gw = P.gw, nx = P.nx;
  // thread = (ghost slot g in [0, 2gw), inner row j, plane k)
  const int g = threadIdx.x % (2 * gw);
  const int j = gw + blockIdx.x * (blockDim.x / (2 * gw)) + threadIdx.x / (2 * gw);
  const int k = k0 + blockIdx.y;
  if (threadIdx.x >= (blockDim.x / (2 * gw)) * 2 * gw || j >= P.jsize - gw) return;
  const bool lo = g < gw;
  const int gg = lo ? g : g - gw;
  const int iDst = lo ? gg : nx + gw + gg;
  const int iSrc = lo ? P.isize - 2 * gw + gg : gw + gg;   // opposite border's inner column
  int j0, j1; T eps;
  remapRows(P, sh, j, lo, j0, j1, eps);
  const T lam = T(0.5) * eps * (eps - T(1));
  const T st = P.slope_type;
  const size_t plane = (size_t)P.isize * P.jsize, comp = plane * P.ksize;
  auto at = [&](int v, int ii, int jj) -> T& { return U[(size_t)v * comp + (size_t)k * plane + (size_t)jj * P.isize + ii]; };
  auto slope = [&](int v, int jj) -> T {  // limited y slope of the border column, :3606-3616
    if (st != T(1) && st != T(2)) return T(0);
    const T dlft = st * (at(v, iSrc, jj) - at(v, iSrc, jj - 1)), drgt = st * (at(v, iSrc, jj + 1) - at(v, iSrc, jj));
    const T dcen = T(0.5) * (dlft + drgt) / st;
    const T dsgn = (dcen >= T(0)) ? T(1) : T(-1);
    T dlim = dev::mn(dev::ab(dlft), dev::ab(drgt));
    if (dlft * drgt <= T(0)) dlim = T(0);
    return dsgn * dev::mn(dlim, dev::ab(dcen));
  };
  for (int v = 0; v < 8; ++v) {
    if (v == IB) {
      const T sl = (st == T(1) || st == T(2)) ? at(IB, iSrc, j0 + 1) - at(IB, iSrc, j0) : T(0);
      at(IB, iDst, j) = at(IB, iSrc, j0) + eps * sl;
      continue;
    }
    if (!lo && v == IA && gg == 0) continue;  // the first outer ghost face keeps its CT value
    const T s0 = slope(v, j0), s1 = slope(v, j1);
    const T d = lo ? (s0 - s1) : (s1 - s0);
    at(v, iDst, j) = (T(1) - eps) * at(v, iSrc, j0) + eps * at(v, iSrc, j1) + lam * d;
  }
}

template <typename T>
__global__ void __launch_bounds__(BX) k_copy_planes(const __grid_constant__ KParams<T> P, const T* __restrict__ Uold,
                                                    T* __restrict__ Unew, int k0) {
  int i, j;
  const int k = k0 + blockIdx.z;
  if (!tileCoords(0, P.isize, 0, P.jsize, i, j)) return;
  const size_t plane = (size_t)P.isize * P.jsize, comp = plane * P.ksize;
  const size_t idx = (size_t)k * plane + (size_t)j * P.isize + i;
  for (int v = 0; v < P.nvar; ++v) Unew[v * comp + idx] = Uold[v * comp + idx];
}

// ------------------------------------------------------------------------------------------------
// stand-alone inverse-dt reduction (reference MHDRunBase.cpp:141-250 / cmpdt_mhd.cuh:155)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(BX) k_invdt(const __grid_constant__ KParams<T> P, const T* __restrict__ Uin,
                                              unsigned long long* __restrict__ dMaxInvDt) {
  const int gw = P.gw;
  int i, j;
  const bool valid = tileCoords(gw, P.nx, gw, P.ny, i, j);
  const int k = (P.dim == 3) ? gw + blockIdx.z : 0;
  T invDt = T(0);
  if (valid) {
    const UView<T> U = uview(Uin, P);
    T u[8], q[8];
#pragma unroll
    for (int v = 0; v < 8; ++v) u[v] = U(v, i, j, k);
    const T bzp = (P.dim == 3) ? U(IC, i, j, k + 1) : T(0);
    dev::cons_to_prim_mhd(P, u, U(IA, i + 1, j, k), U(IB, i, j + 1, k), bzp, T(0), q);
    // 2D: the reference passes magFieldNeighbors[IZ] = 0 (constoprim.h:409)
    const T irho = dev::rcp(q[ID]);
    const T a2 = q[IA] * q[IA], b2 = q[IB] * q[IB], c2 = q[IC] * q[IC];
    const T bb = a2 + b2 + c2;
    T vx = dev::fast_speed(P.gamma0, q[IP], irho, bb, a2) + dev::ab(q[IU]);
    T vy = dev::fast_speed(P.gamma0, q[IP], irho, bb, b2) + dev::ab(q[IV]);
    if (P.dim == 3) {
      T vz = dev::fast_speed(P.gamma0, q[IP], irho, bb, c2) + dev::ab(q[IW]);
      if (P.Omega0 > T(0)) vy += T(1.5) * P.Omega0 * (P.xMax - P.xMin) * T(0.5);
      invDt = vx * P.rdx + vy * P.rdy + vz * P.rdz;
    } else {
      invDt = vx * P.rdx + vy * P.rdy;
    }
  }
  reduceMaxToSlots(invDt, dMaxInvDt);
}

// ------------------------------------------------------------------------------------------------
// ghost fill (reference make_boundary_base.h:1040-1332): one launch per direction, both faces
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_boundary(const __grid_constant__ KParams<T> P, T* __restrict__ U, int dir, int bcLo, int bcHi,
                           int skipLo, int skipHi, int kLo, int kHi) {
  // thread = one (ghost slot g in [0, 2*gw), transverse a, transverse b) for all variables
  const int gw = P.gw;
  const int sizes[3] = {P.isize, P.jsize, P.ksize};
  const int nn[3] = {P.nx, P.ny, P.nz};
  const int n = nn[dir];
  const int d1 = (dir == 0) ? 1 : 0, d2 = (dir == 2) ? 1 : 2;  // transverse dirs (d1 faster)
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y * blockDim.y + threadIdx.y;
  const int g = blockIdx.z;
  if (a >= sizes[d1] || b >= sizes[d2]) return;
  const bool hi = g >= gw;
  if (hi ? skipHi : skipLo) return;
  const int bct = hi ? bcHi : bcLo;
  if (bct != BC_DIRICHLET && bct != BC_NEUMANN && bct != BC_PERIODIC) return;
  const int gi = hi ? n + g : g;  // ghost index along dir (hi: n+gw+(g-gw))
  int src;
  if (bct == BC_DIRICHLET) src = hi ? 2 * n + 2 * gw - 1 - gi : 2 * gw - 1 - gi;
  else if (bct == BC_NEUMANN) src = hi ? n + gw - 1 : gw;
  else src = hi ? gi - n : gi + n;
  int c[3], s[3];
  c[dir] = gi; c[d1] = a; c[d2] = b;
  s[dir] = src; s[d1] = a; s[d2] = b;
  if (c[2] < kLo || c[2] >= kHi) return;
  const size_t plane = (size_t)P.isize * P.jsize, comp = plane * P.ksize;
  const size_t o = (size_t)c[2] * plane + (size_t)c[1] * P.isize + c[0];
  const size_t in = (size_t)s[2] * plane + (size_t)s[1] * P.isize + s[0];
  const int normalVar = IU + dir;
  for (int v = 0; v < P.nvar; ++v) {
    T val = U[v * comp + in];
    if (bct == BC_DIRICHLET && v == normalVar) val = -val;
    U[v * comp + o] = val;
  }
}

// ------------------------------------------------------------------------------------------------
// z ghost planes of the stratified shearing box (reference make_boundary2_z_stratified, make_boundary_base.h:1357-1647,
// ghost width 3): hydrostatic extrapolation of the density (r1, r2, r3 = density ratios of successive planes, 1 with
// [MRI] floor), horizontal momenta scaled with it, vertical momentum copied when it points outwards and zero otherwise,
// zero horizontal field, B_z continued with div B = 0 (the ghost B_x, B_y are zero, so B_z is constant along z; the
// last row / column is left alone like in the reference).  One thread per (i, j) column of one face.
template <typename T>
__global__ void k_boundary_zstrat(const __grid_constant__ KParams<T> P, T* __restrict__ U, int hi, T r1, T r2, T r3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= P.isize || j >= P.jsize) return;
  const int ksz = P.ksize;
  const int e = hi ? ksz - 4 : 3, s = hi ? 1 : -1;
  const int g1 = e + s, g2 = e + 2 * s, g3 = e + 3 * s;
  const size_t plane = (size_t)P.isize * P.jsize, comp = plane * ksz;
  auto at = [&](int v, int k) -> T& { return U[(size_t)v * comp + (size_t)k * plane + (size_t)j * P.isize + i]; };
  const T rho_e = at(ID, e);
  const T rho1 = rho_e * r1, rho2 = rho_e * r1 * r2, rho3 = rho_e * r1 * r2 * r3;
  at(ID, g1) = rho1; at(ID, g2) = rho2; at(ID, g3) = rho3;
  for (int v = IU; v <= IV; ++v) {
    const T m = at(v, e);
    at(v, g3) = m / rho_e * rho3;
    at(v, g2) = m / rho_e * rho2;
    at(v, g1) = m / rho_e * rho1;
  }
  const T we = at(IW, e);
  const T w = hi ? ((we > T(0)) ? we : T(0)) : ((we < T(0)) ? we : T(0));
  at(IW, g1) = w; at(IW, g2) = w; at(IW, g3) = w;
  for (int v = IA; v <= IB; ++v) { at(v, g1) = T(0); at(v, g2) = T(0); at(v, g3) = T(0); }
  if (i < P.isize - 1 && j < P.jsize - 1) {
    if (!hi) {  // B_z on the low faces of ghost planes 2, 1, 0 from the one of plane 3
      const T bz = at(IC, 3);
      at(IC, 2) = bz; at(IC, 1) = bz; at(IC, 0) = bz;
    } else {    // planes ksize-2, ksize-1 from the low-face B_z of plane ksize-3 (set by the CT update)
      const T bz = at(IC, ksz - 3);
      at(IC, ksz - 2) = bz; at(IC, ksz - 1) = bz;
    }
  }
}

// x direction of the ghost fill with the 2 gw ghost cells of a row on CONSECUTIVE lanes: a warp touches 32 / (2 gw) rows
// with one short contiguous segment per row and face, instead of 32 rows with one element each (k_boundary's mapping,
// right for y and z where a warp runs along x).  Same copies, same values.
template <typename T>
__global__ void __launch_bounds__(256) k_boundary_x(const __grid_constant__ KParams<T> P, T* __restrict__ U, int bcLo, int bcHi,
                                                    int skipLo, int skipHi, int kLo, int kHi) {
  const int gw = P.gw, per = 2 * gw, n = P.nx;
  const int rowsPerBlock = blockDim.x / per;
  const int g = threadIdx.x % per;
  const long row = (long)blockIdx.x * rowsPerBlock + threadIdx.x / per;  // row = j + jsize * k
  if ((int)threadIdx.x >= rowsPerBlock * per || row >= (long)P.jsize * P.ksize) return;
  const int k = (int)(row / P.jsize);
  if (k < kLo || k >= kHi) return;
  const bool hi = g >= gw;
  if (hi ? skipHi : skipLo) return;
  const int bct = hi ? bcHi : bcLo;
  if (bct != BC_DIRICHLET && bct != BC_NEUMANN && bct != BC_PERIODIC) return;
  const int gi = hi ? n + g : g;
  int src;
  if (bct == BC_DIRICHLET) src = hi ? 2 * n + 2 * gw - 1 - gi : 2 * gw - 1 - gi;
  else if (bct == BC_NEUMANN) src = hi ? n + gw - 1 : gw;
  else src = hi ? gi - n : gi + n;
  const size_t comp = (size_t)P.isize * P.jsize * P.ksize;
  const size_t base = (size_t)row * P.isize;
  for (int v = 0; v < P.nvar; ++v) {
    T val = U[v * comp + base + src];
    if (bct == BC_DIRICHLET && v == IU) val = -val;
    U[v * comp + base + gi] = val;
  }
}

// ------------------------------------------------------------------------------------------------
// probes for the known-answer tests
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_probe_riemann(const __grid_constant__ KParams<T> P, int n, const T* ql, const T* qr, T* flux) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const T* l = ql + 8 * t;
  const T* r = qr + 8 * t;
  dev::State<T> L{l[ID], l[IP], l[IU], l[IV], l[IW], l[IA], l[IB], l[IC]};
  dev::State<T> R{r[ID], r[IP], r[IU], r[IV], r[IW], r[IA], r[IB], r[IC]};
  T f[8];
  dev::riemann_mhd(P, L, R, f);
  for (int v = 0; v < 8; ++v) flux[8 * t + v] = f[v];
}

template <typename T>
__global__ void k_probe_emf(const __grid_constant__ KParams<T> P, int n, int emfDir, const T* qEdge, const T* xPos,
                            T* emf) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  // qEdge[t][4][8] in the reference's (IRT, IRB, ILT, ILB) x (ID..IC) PHYSICAL layout
  int iu, iv, iw, ia, ib, ic;
  if (emfDir == 2) { iu = IU; iv = IV; iw = IW; ia = IA; ib = IB; ic = IC; }
  else if (emfDir == 1) { iu = IW; iv = IU; iw = IV; ia = IC; ib = IA; ic = IB; }
  else { iu = IV; iv = IW; iw = IU; ia = IB; ib = IC; ic = IA; }
  dev::Corner<T> c[4];
  for (int e = 0; e < 4; ++e) {
    const T* q = qEdge + (size_t)t * 32 + e * 8;
    c[e] = dev::Corner<T>{q[ID], q[IP], q[iu], q[iv], q[iw], q[ia], q[ib], q[ic]};
  }
  emf[t] = dev::compute_emf(P, c[0], c[1], c[2], c[3], emfDir, xPos ? xPos[t] : T(0));
}

}  // namespace

unsigned long long kernelLaunchCount() { return g_launches; }
bool setTuning(const char* key, int value) {
  const std::string k = key ? key : "";
  if (k == "tile_x") {
    if (value != 32 && value != 64 && value != 128) return false;
    g_tileX = value;
    return true;
  }
  if (k == "fused_b") {
    g_fusedB = value ? 1 : 0;
    return true;
  }
  if (k == "handoff_head") {
    if (value < 0 || value > 8) return false;
    g_handoffHead = value;
    return true;
  }
  if (k == "fused_handoff") {
    g_fusedHandoff = value ? 1 : 0;
    return true;
  }
  if (k == "rot_dt") {
    g_rotDt = value ? 1 : 0;
    return true;
  }
  if (k == "fused_a") {
    g_fusedA = value ? 1 : 0;
    return true;
  }
  if (k == "halo_p2p") {  // read when a multi-rank run is created
    g_haloP2p = value ? 1 : 0;
    return true;
  }
  if (k == "hydro_tma") {
    g_hydroTma = value ? 1 : 0;
    return true;
  }
  if (k == "hydro_rows") {
    if (value != 0 && value != 12 && value != 16 && value != 20 && value != 24) return false;
    g_hydroRows = value;
    return true;
  }
  if (k == "hydro_fused") {
    g_hydroFused = value ? 1 : 0;
    return true;
  }
  if (k == "hydro_tile") {
    g_hydroTile = value ? 1 : 0;
    return true;
  }
  if (k == "trace_qy") {
    if (value != 8 && value != 12 && value != 16) return false;
    g_traceQY = value;
    return true;
  }
  if (value < 2 || value > 8) return false;
  if (k == "flux_minb") g_fluxMinB = value;
  else if (k == "emf_minb") g_emfMinB = value;
  else if (k == "trace_minb") g_traceMinB = value;
  else if (k == "update_minb") g_updateMinB = value;
  else return false;
  return true;
}
void resetKernelLaunchCount() { g_launches = 0; }

// jet inflow patch: (ijet x ijet x gw) cells in 3D, (ijet x gw) in 2D (reference make_jet)
template <typename T>
__global__ void k_jet(const __grid_constant__ KParams<T> P, T* __restrict__ U) {
  const int a = P.gw + P.offsetJet;
  const int i = a + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a + P.ijet || i >= P.isize) return;
  const size_t plane = (size_t)P.isize * P.jsize, comp = plane * P.ksize;
  const T e = P.pjet / (P.gamma0 - T(1)) + T(0.5) * P.djet * P.ujet * P.ujet;
  size_t idx;
  if (P.dim == 2) {
    idx = (size_t)blockIdx.y * P.isize + i;              // j = blockIdx.y < gw
  } else {
    const int j = a + blockIdx.y, k = blockIdx.z;      // k < gw
    if (j >= P.jsize) return;
    idx = (size_t)k * plane + (size_t)j * P.isize + i;
  }
  U[ID * comp + idx] = P.djet;
  U[IP * comp + idx] = e;
  U[IU * comp + idx] = T(0);
  if (P.dim == 2) {
    U[IV * comp + idx] = P.djet * P.ujet;
  } else {
    U[IV * comp + idx] = T(0);
    U[IW * comp + idx] = P.djet * P.ujet;
  }
}

// ---- launch wrappers ---------------------------------------------------------------------------
template <typename T>
void MhdKernels<T>::fillBoundary(const KParams<T>& P, T* U, int dir, int bcLo, int bcHi, bool skipLo, bool skipHi,
                                 int kLo, int kHi, cudaStream_t s) {
  if (dir == 2 && P.dim == 2) return;
  if (dir == 0) {
    const int per = 2 * P.gw, rowsPerBlock = 256 / per;
    const long rows = (long)P.jsize * P.ksize;
    k_boundary_x<T><<<(unsigned)((rows + rowsPerBlock - 1) / rowsPerBlock), 256, 0, s>>>(P, U, bcLo, bcHi, skipLo ? 1 : 0,
                                                                                         skipHi ? 1 : 0, kLo, kHi);
    launched();
    return;
  }
  const int sizes[3] = {P.isize, P.jsize, P.ksize};
  const int d1 = (dir == 0) ? 1 : 0, d2 = (dir == 2) ? 1 : 2;
  dim3 block(32, 8, 1);
  dim3 grid((sizes[d1] + 31) / 32, (sizes[d2] + 7) / 8, 2 * P.gw);
  k_boundary<T><<<grid, block, 0, s>>>(P, U, dir, bcLo, bcHi, skipLo ? 1 : 0, skipHi ? 1 : 0, kLo, kHi);
  launched();
}

template <typename T>
void MhdKernels<T>::fillBoundaryZStratified(const KParams<T>& P, T* U, bool lo, bool hi, bool floorDensity, cudaStream_t s) {
  // density ratios of successive ghost planes: exp(-dz / (2 H^2) (+-2 z_face + (2n - 1) dz)), H = cIso / Omega0
  const T dz = P.dz, HALF = T(0.5);
  const T H = P.cIso / P.Omega0;
  const T factor = static_cast<T>(-dz / 2.0 / H / H);
  const dim3 block(32, 8, 1), grid((P.isize + 31) / 32, (P.jsize + 7) / 8, 1);
  for (int side = 0; side < 2; ++side) {
    if (side == 0 ? !lo : !hi) continue;
    T r1 = T(1), r2 = T(1), r3 = T(1);
    if (!floorDensity) {
      if (side == 0) {
        r1 = static_cast<T>(std::exp(factor * (-2 * (P.zMin + HALF * dz) + dz)));
        r2 = static_cast<T>(std::exp(factor * (-2 * (P.zMin + HALF * dz) + 3.0 * dz)));
        r3 = static_cast<T>(std::exp(factor * (-2 * (P.zMin + HALF * dz) + 5.0 * dz)));
      } else {
        r1 = static_cast<T>(std::exp(factor * (2 * (P.zMax - HALF * dz) + dz)));
        r2 = static_cast<T>(std::exp(factor * (2 * (P.zMax - HALF * dz) + 3.0 * dz)));
        r3 = static_cast<T>(std::exp(factor * (2 * (P.zMax - HALF * dz) + 5.0 * dz)));
      }
    }
    k_boundary_zstrat<T><<<grid, block, 0, s>>>(P, U, side, r1, r2, r3);
    launched();
  }
}

template <typename T>
void MhdKernels<T>::jetInflow(const KParams<T>& P, T* U, cudaStream_t s) {
  if (!P.jet || P.ijet <= 0) return;
  const dim3 grid((P.ijet + 31) / 32, P.dim == 2 ? P.gw : P.ijet, P.dim == 2 ? 1 : P.gw);
  k_jet<T><<<grid, 32, 0, s>>>(P, U);
  launched();
}

template <typename T>
void MhdKernels<T>::computeInvDt(const KParams<T>& P, const T* U, unsigned long long* d, cudaStream_t s) {
  k_invdt<T><<<gridFor(P.nx, P.ny, P.dim == 3 ? P.nz : 1), blockShape(), 0, s>>>(P, U, d);
  launched();
}

template <typename T>
void MhdKernels<T>::prim(const KParams<T>& P, const T* U, MhdScratch<T> sc, int k0, int k1, T dt, cudaStream_t s) {
  if (k1 <= k0) return;
  k_prim<T><<<gridFor(P.isize - 1, P.jsize - 1, k1 - k0), blockShape(), 0, s>>>(P, U, sc.Q, sc.planes, sc.kbase, k0, dt);
  launched();
}

// MINB dispatch: FP64 kernels are compiled for several occupancy targets (register caps), picked at
// run time by setTuning(); the FP32 flavour keeps one variant each.
#define RG_MINB_SWITCH(T, minb, LAUNCH, DFLT)              \
  if (sizeof(T) == 4) { LAUNCH(DFLT, false); }             \
  else if (!fastPath(P)) { LAUNCH(4, false); }             \
  else switch (minb) {                                     \
    case 2:                                                \
    case 3: LAUNCH(3, true); break;                        \
    case 4: LAUNCH(4, true); break;                        \
    case 5: LAUNCH(5, true); break;                        \
    case 6: LAUNCH(6, true); break;                        \
    default: LAUNCH(8, true); break;                       \
  }
// the FAST instantiations (see mhd_device.cuh) serve the headline configuration: adiabatic,
// non-rotating, HLLD fluxes + 2-D HLLD emfs; everything else runs the generic kernels
template <typename T>
static bool fastPath(const KParams<T>& P) {
  return P.cIso <= T(0) && P.Omega0 <= T(0) && P.riemannSolver == RS_HLLD && P.magRiemannSolver == MAG_HLLD &&
         P.slope_type != T(3);  // the 27-point slopes run on the separate kernels (k_trace<.., S3 = true>)
}

template <typename T>
void MhdKernels<T>::elec(const KParams<T>& P, const T* U, MhdScratch<T> sc, int k0, int k1, cudaStream_t s) {
  if (k1 <= k0) return;
  k_elec<T><<<gridFor(P.isize - 2, P.jsize - 2, k1 - k0), blockShape(), 0, s>>>(P, U, sc.Q, sc.EL, sc.planes, sc.kbase, k0);
  launched();
}

template <typename T>
void MhdKernels<T>::trace(const KParams<T>& P, const T* U, MhdScratch<T> sc, int k0, int k1, T dt, cudaStream_t s) {
  if (k1 <= k0) return;
  const int n = P.isize - 2 * P.gw + 2, m = P.jsize - 2 * P.gw + 2;  // gw-1 .. size-gw
  const dim3 g = gridFor(n, m, k1 - k0);
  if (P.slope_type == T(3)) {  // 27-point slopes: their own instantiation of the generic kernel
    k_trace<T, 4, false, true><<<g, blockShape(), 0, s>>>(P, U, sc.Q, sc.EL, sc.W, sc.planes, sc.kbase, k0, dt);
    launched();
    return;
  }
#define RG_L(M, FA) k_trace<T, M, FA><<<g, blockShape(), 0, s>>>(P, U, sc.Q, sc.EL, sc.W, sc.planes, sc.kbase, k0, dt)
  RG_MINB_SWITCH(T, g_traceMinB, RG_L, 4)
#undef RG_L
  launched();
}

template <typename T>
void MhdKernels<T>::flux(const KParams<T>& P, MhdScratch<T> sc, int k0, int k1, cudaStream_t s) {
  if (k1 <= k0) return;
  const dim3 g = gridFor(P.nx + 1, P.ny + 1, k1 - k0);
#define RG_L(M, FA)                                                                   \
  k_flux<T, 0, M, FA><<<g, blockShape(), 0, s>>>(P, sc.W, sc.F, sc.planes, sc.kbase, k0);       \
  k_flux<T, 1, M, FA><<<g, blockShape(), 0, s>>>(P, sc.W, sc.F, sc.planes, sc.kbase, k0);       \
  k_flux<T, 2, M, FA><<<g, blockShape(), 0, s>>>(P, sc.W, sc.F, sc.planes, sc.kbase, k0)
  RG_MINB_SWITCH(T, g_fluxMinB, RG_L, 6)
#undef RG_L
  launched(3);
}

template <typename T>
void MhdKernels<T>::emf(const KParams<T>& P, MhdScratch<T> sc, int k0, int k1, cudaStream_t s) {
  if (k1 <= k0) return;
  const dim3 g = gridFor(P.nx + 1, P.ny + 1, k1 - k0);
#define RG_L(M, FA)                                                                   \
  k_emf<T, 2, M, FA><<<g, blockShape(), 0, s>>>(P, sc.W, sc.E, sc.planes, sc.kbase, k0);        \
  k_emf<T, 1, M, FA><<<g, blockShape(), 0, s>>>(P, sc.W, sc.E, sc.planes, sc.kbase, k0);        \
  k_emf<T, 0, M, FA><<<g, blockShape(), 0, s>>>(P, sc.W, sc.E, sc.planes, sc.kbase, k0)
  RG_MINB_SWITCH(T, g_emfMinB, RG_L, 6)
#undef RG_L
  launched(3);
}

template <typename T>
void MhdKernels<T>::update(const KParams<T>& P, const T* Uold, T* Unew, MhdScratch<T> sc, int k0, int k1, T dt,
                           unsigned long long* d, cudaStream_t s) {
  if (k1 <= k0) return;
  const dim3 g = gridFor(P.isize, P.jsize, k1 - k0);
#define RG_L(M, FA) k_update<T, M, FA><<<g, blockShape(), 0, s>>>(P, Uold, Unew, sc.F, sc.E, sc.planes, sc.kbase, k0, dt, d)
  RG_MINB_SWITCH(T, g_updateMinB, RG_L, 8)
#undef RG_L
  launched();
}

// ---- fused path ---------------------------------------------------------------------------------
int g_fusedA = 1;  // run-time knob "fused_a": fused prim+elec+trace kernel when available
bool fusedTraceRequested() { return g_fusedA != 0; }

template <typename T>
bool MhdKernels<T>::fusedTraceAvailable(const KParams<T>& P) {
  // every FP64 3D MHD configuration: the FAST instantiation for the headline one, the generic one
  // (rotating frame, isothermal, other Riemann solvers) otherwise
  if (sizeof(T) != 8 || P.dim != 3) return false;
  if (P.slope_type == T(3)) return false;  // 27-point slopes: separate prim / elec / trace kernels
  static int okDev[MAX_DEVICES];
  static bool init = false;
  if (!init) { for (int d = 0; d < MAX_DEVICES; ++d) okDev[d] = -1; init = true; }
  int& ok = okDev[currentDevice()];
  if (ok < 0)
    ok = (cudaFuncSetAttribute(k_fused_trace<T, TraceTileT<8>, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)TraceTileT<8>::SMEM) == cudaSuccess &&
          cudaFuncSetAttribute(k_fused_trace<T, TraceTileT<16>, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)TraceTileT<16>::SMEM) == cudaSuccess &&
          cudaFuncSetAttribute(k_fused_trace<T, TraceTileT<12>, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)TraceTileT<12>::SMEM) == cudaSuccess &&
          cudaFuncSetAttribute(k_fused_trace<T, TraceTileT<12>, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)TraceTileT<12>::SMEM) == cudaSuccess)
             ? 1
             : 0;
  if (!ok) cudaGetLastError();
  return ok == 1;
}

template <typename T, typename TT, bool FAST>
static void launchFusedTrace(const KParams<T>& P, const T* U, const MhdScratch<T>& sc, int k0, int k1, T dt, int nSM,
                             cudaStream_t s) {
  const int n = P.isize - 2 * P.gw + 2, m = P.jsize - 2 * P.gw + 2;  // traced cells gw-1 .. size-gw
  const int ntx = (n + TT::TW - 1) / TT::TW, nty = (m + TT::TH - 1) / TT::TH;
  const int planes = k1 - k0, slots = TT::MINB * nSM;  // resident blocks
  int bestNz = 1;
  double bestCost = 1e300;
  for (int nz = 1; nz <= planes; ++nz) {
    const int lz = (planes + nz - 1) / nz;
    if (lz < 8 && nz > 1) break;
    const long blocks = (long)ntx * nty * ((planes + lz - 1) / lz);
    const double cost = (double)((blocks + slots - 1) / slots) * (lz + 2.5);
    if (cost < bestCost) { bestCost = cost; bestNz = nz; }
  }
  const int lz = (planes + bestNz - 1) / bestNz;
  const dim3 grid(ntx, nty, (planes + lz - 1) / lz);
  k_fused_trace<T, TT, FAST><<<grid, dim3(TT::QX, TT::QY, 1), TT::SMEM, s>>>(P, U, sc.W, sc.planes, sc.kbase, k0, k1, lz,
                                                                           dt);
}

template <typename T>
void MhdKernels<T>::fusedTrace(const KParams<T>& P, const T* U, MhdScratch<T> sc, int k0, int k1, T dt, cudaStream_t s) {
  if (k1 <= k0) return;
  const int nSM = smCount();
  if (!fastPath(P)) launchFusedTrace<T, TraceTileT<12>, false>(P, U, sc, k0, k1, dt, nSM, s);
  else if (g_traceQY == 12) launchFusedTrace<T, TraceTileT<12>, true>(P, U, sc, k0, k1, dt, nSM, s);
  else if (g_traceQY == 16) launchFusedTrace<T, TraceTileT<16>, true>(P, U, sc, k0, k1, dt, nSM, s);
  else launchFusedTrace<T, TraceTileT<8>, true>(P, U, sc, k0, k1, dt, nSM, s);
  launched();
}

// 512 threads / 124 registers: a 384-thread build (152 registers) measured 7 % slower, 640 threads spill
template <typename T>
struct FusedSel {
  typedef FusedTile<T, 16, 8, 512, true> Cfg;      // hand-off tile (knob "fused_handoff" = 1)
  typedef FusedTile<T, 15, 7, 512, false> Legacy;  // self-closing tile (default)
};
// Measured on the B200 (profiles/r02_l_handoff_vs_self_closing.txt): the hand-off kernel executes 15 % fewer instructions
// (every Riemann problem solved once) but runs at 6.8 ms against 4.7 ms at 256^3: its 3400 SASS instructions (54 KB, the
// four helper tasks on top of the solvers and the update) no longer fit the instruction cache that the 2856-instruction
// self-closing kernel just fits (stall no_instruction 1.07 against 0.16 per issue, issue 42 % against 63 %).  Results are
// bitwise identical; the self-closing tile stays the default until the solver code is small enough for both.
int g_fusedHandoff = 0;
bool fusedHandoffRequested() { return g_fusedHandoff != 0; }
int g_handoffHead = 1;  // run-time knob "handoff_head": planes of head start of a tile's producers (0 = none)

// the rotating-frame instantiation of the fused kernel: HLLD + 2-D HLLD in the rotating frame (any closure)
template <typename T>
static bool rotatingFusedPath(const KParams<T>& P) {
  return P.Omega0 > T(0) && P.riemannSolver == RS_HLLD && P.magRiemannSolver == MAG_HLLD && P.slope_type != T(3) &&
         !P.gravity;
}

template <typename T>
bool MhdKernels<T>::fusedUpdateEligible(const KParams<T>& P) {
  if (sizeof(T) != 8 || !(fastPath(P) || rotatingFusedPath(P)) || P.dim != 3) return false;
  if ((P.gw - 1) % 2 != 0) return false;                    // box rows must start on an even cell index
  return ((size_t)P.isize * sizeof(T)) % 16 == 0;           // TMA: row pitch a multiple of 16 bytes
}

// launch geometry of the fused update over `planes` update planes: tiles, z ranges (whole waves of one block per SM)
struct FusedGeom { int ntx, nty, nz, lz; };
template <typename T, typename C>
static FusedGeom fusedGeometry(const KParams<T>& P, int planes) {
  FusedGeom g;
  if (C::HANDOFF) {  // every cell of the update box, ghost-face column / row included, belongs to exactly one tile
    g.ntx = (P.nx + 1 + C::TW - 1) / C::TW;
    g.nty = (P.ny + 1 + C::TH - 1) / C::TH;
  } else {           // the ghost-face column / row is folded into the last tile when it would otherwise open a tile of its own
    g.ntx = std::max(1, (P.nx + C::TW - 1) / C::TW);
    g.nty = std::max(1, (P.ny + C::TH - 1) / C::TH);
  }
  const int nSM = smCount();
  int bestNz = 1;
  double bestCost = 1e300;
  for (int nz = 1; nz <= planes; ++nz) {
    const int lz = (planes + nz - 1) / nz;
    if (lz > C::LZMAX) continue;
    if (lz < 8 && nz > 1) break;
    const long blocks = (long)g.ntx * g.nty * ((planes + lz - 1) / lz);
    // a z range costs its planes + the pipeline fill (+ the start-up skew of the hand-off chain of a wave of tiles)
    const double cost = (double)((blocks + nSM - 1) / nSM) * (lz + (C::HANDOFF ? 4.0 : 1.5));
    if (cost < bestCost) { bestCost = cost; bestNz = nz; }
  }
  g.lz = (planes + bestNz - 1) / bestNz;
  g.nz = (planes + g.lz - 1) / g.lz;
  return g;
}

template <typename T>
void MhdKernels<T>::fusedHandoffSize(const KParams<T>& P, int planes, size_t* reals, size_t* ints) {
  typedef typename FusedSel<T>::Cfg C;
  *reals = 0;
  *ints = 0;
  if (!g_fusedHandoff || planes <= 0) return;
  const FusedGeom g = fusedGeometry<T, C>(P, planes);
  const size_t blocks = (size_t)g.ntx * g.nty * g.nz;
  *reals = blocks * (size_t)(g.lz + 2) * C::HREC;
  *ints = 2 * blocks + 1;
}

template <typename T>
void MhdKernels<T>::fusedPrepare(const KParams<T>& P, MhdScratch<T>& sc) {
  typedef typename FusedSel<T>::Cfg C;
  typedef typename FusedSel<T>::Legacy L;
  static_assert(C::WX == L::WX && C::WY == L::WY, "both tiles share the tensor map of W");
  sc.fused = 0;
  if (!fusedUpdateEligible(P) || sc.W == nullptr) return;
  CUtensorMap map;
  if (!tma::encodeTile4D(&map, sc.W, (int)sizeof(T), P.isize, P.jsize, sc.planes, NW_MHD, C::WX, C::WY)) return;
  static_assert(sizeof(CUtensorMap) <= sizeof(sc.mapW), "tensor map storage");
  memcpy(sc.mapW, &map, sizeof(map));
  static bool attrSetDev[MAX_DEVICES] = {false};
  bool& attrSet = attrSetDev[currentDevice()];
  if (!attrSet) {
    bool ok = true;
#define RG_ATTR(K, SM) ok = ok && cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SM)) == cudaSuccess
    RG_ATTR((k_fused_flux_emf_update<T, C, true>), C::SMEM);
    RG_ATTR((k_fused_flux_emf_update<T, C, false>), C::SMEM);
    RG_ATTR((k_fused_flux_emf_update<T, C, false, false>), C::SMEM);
    RG_ATTR((k_fused_flux_emf_update<T, L, true>), L::SMEM);
    RG_ATTR((k_fused_flux_emf_update<T, L, false>), L::SMEM);
    RG_ATTR((k_fused_flux_emf_update<T, L, false, false>), L::SMEM);
#undef RG_ATTR
    if (!ok) {
      cudaGetLastError();
      return;
    }
    attrSet = true;
  }
  sc.fused = 1;
}

template <typename T, typename C>
static void launchFusedUpdate(const KParams<T>& P, const T* Uold, T* Unew, const MhdScratch<T>& sc, int ka, int kb, T dt,
                              unsigned long long* d, cudaStream_t s, int shearEnabled, int jplus, T frac) {
  const FusedGeom g = fusedGeometry<T, C>(P, kb - ka);
  const dim3 grid(g.ntx, g.nty, g.nz);
  const int lz = g.lz, hplanes = g.lz + 2;
  T* hbuf = nullptr;
  int* hsync = nullptr;
  if (C::HANDOFF) {
    const size_t blocks = (size_t)g.ntx * g.nty * g.nz;
    if (sc.hbuf == nullptr || sc.hsync == nullptr || sc.hbufReals < blocks * (size_t)hplanes * C::HREC || sc.hsyncInts < 2 * blocks + 1)
      throw std::runtime_error("fused update: the hand-off buffers are too small for this launch");
    hbuf = sc.hbuf;
    hsync = sc.hsync;
    // tile counter and per-tile progress flags start at zero for every launch
    if (cudaMemsetAsync(hsync, 0, (2 * blocks + 1) * sizeof(int), s) != cudaSuccess)
      throw std::runtime_error("CUDA error: cudaMemsetAsync(hand-off flags)");
  }
  CUtensorMap map;
  memcpy(&map, sc.mapW, sizeof(map));
  const ShearShift<T> sh{shearEnabled, jplus, frac};
  if (fastPath(P)) {
    k_fused_flux_emf_update<T, C, true><<<grid, C::THREADS, C::SMEM, s>>>(P, map, Uold, Unew, sc.kbase, ka, kb, lz, dt, d, sh,
                                                                          nullptr, 0, hbuf, hsync, g.ntx, g.nty, hplanes, g_handoffHead);
    launched();
    return;
  }
  // rotating frame; with shearing-box boundaries the three border cell columns follow from the strips
  if (g_rotDt)
    k_fused_flux_emf_update<T, C, false, true><<<grid, C::THREADS, C::SMEM, s>>>(
        P, map, Uold, Unew, sc.kbase, ka, kb, lz, dt, d, sh, sc.strips, sc.planes, hbuf, hsync, g.ntx, g.nty, hplanes,
        g_handoffHead);
  else
    k_fused_flux_emf_update<T, C, false, false><<<grid, C::THREADS, C::SMEM, s>>>(
        P, map, Uold, Unew, sc.kbase, ka, kb, lz, dt, nullptr, sh, sc.strips, sc.planes, hbuf, hsync, g.ntx, g.nty, hplanes,
        g_handoffHead);
  launched();
  if (shearEnabled) {
    const int nRows = P.jsize - 2 * P.gw + 1;
    k_update_rot_border<T><<<dim3((3 * nRows + 127) / 128, 1, kb - ka), 128, 0, s>>>(P, Uold, Unew, sc.strips, sc.planes,
                                                                                       sc.kbase, ka, dt, sh, g_rotDt ? d : nullptr);
    launched();
  }
}

template <typename T>
void MhdKernels<T>::fusedFluxEmfUpdate(const KParams<T>& P, const T* Uold, T* Unew, const MhdScratch<T>& sc, int ka,
                                       int kb, T dt, unsigned long long* d, cudaStream_t s, int shearEnabled, int jplus,
                                       T frac) {
  if (kb <= ka) return;
  if (g_fusedHandoff)
    launchFusedUpdate<T, typename FusedSel<T>::Cfg>(P, Uold, Unew, sc, ka, kb, dt, d, s, shearEnabled, jplus, frac);
  else
    launchFusedUpdate<T, typename FusedSel<T>::Legacy>(P, Uold, Unew, sc, ka, kb, dt, d, s, shearEnabled, jplus, frac);
}

template <typename T>
void MhdKernels<T>::copyOutsideBox(const KParams<T>& P, const T* Uold, T* Unew, int k0, int k1, cudaStream_t s) {
  if (k1 <= k0) return;
  const int ng = 2 * P.gw - 1, cells = ng * P.isize + ng * (P.jsize - ng);
  k_copy_outside_box<T><<<dim3((cells + 255) / 256, k1 - k0, 1), 256, 0, s>>>(P, Uold, Unew, k0);
  launched();
}

template <typename T>
void MhdKernels<T>::updateRotating(const KParams<T>& P, const T* Uold, T* Unew, MhdScratch<T> sc, int k0, int k1, T dt,
                                   int shearEnabled, int jplus, T frac, unsigned long long* d, cudaStream_t s) {
  if (k1 <= k0) return;
  ShearShift<T> sh{shearEnabled, jplus, frac};
  k_update_rot<T><<<gridFor(P.isize, P.jsize, k1 - k0), blockShape(), 0, s>>>(P, Uold, Unew, sc.F, sc.E, sc.planes,
                                                                              sc.kbase, k0, dt, sh, d);
  launched();
}

template <typename T>
void MhdKernels<T>::shearGhosts(const KParams<T>& P, T* U, int jplus, T frac, int k0, int k1, cudaStream_t s) {
  if (k1 <= k0) return;
  ShearShift<T> sh{1, jplus, frac};
  const int per = 2 * P.gw, rows = 128 / per;
  dim3 grid((P.ny + rows - 1) / rows, k1 - k0, 1);
  k_shear_ghosts<T><<<grid, 128, 0, s>>>(P, U, sh, k0);
  launched();
}

template <typename T>
void MhdKernels<T>::copyPlanes(const KParams<T>& P, const T* Uold, T* Unew, int k0, int k1, cudaStream_t s) {
  if (k1 <= k0) return;
  k_copy_planes<T><<<gridFor(P.isize, P.jsize, k1 - k0), blockShape(), 0, s>>>(P, Uold, Unew, k0);
  launched();
}

template <typename T>
void MhdKernels<T>::probeRiemann(const KParams<T>& P, int n, const T* ql, const T* qr, T* flux, cudaStream_t s) {
  k_probe_riemann<T><<<(n + 127) / 128, 128, 0, s>>>(P, n, ql, qr, flux);
  launched();
}

template <typename T>
void MhdKernels<T>::probeEmf(const KParams<T>& P, int n, int emfDir, const T* qEdge, const T* xPos, T* emf,
                             cudaStream_t s) {
  k_probe_emf<T><<<(n + 127) / 128, 128, 0, s>>>(P, n, emfDir, qEdge, xPos, emf);
  launched();
}

template struct MhdKernels<double>;
template struct MhdKernels<float>;

}  // namespace rg
