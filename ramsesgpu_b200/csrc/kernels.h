// Launch wrappers of the sm_100a kernels (definitions in kernels_*.cu).  Host-callable, templated on
// the solver precision; explicit instantiations exist for double and float.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "params.h"

namespace rg {

// Number of components of the "traced state" W written by the trace kernel per cell
// ("HBM layout"): 8 cell-centred + 3 low-face B (all advanced by dt/2) + 21 half
// slopes of the cell-centred variables + 6 half slopes of the low-face fields (the high-face values
// of a cell are the low-face values of its +1 neighbour, bit for bit).
constexpr int NW_MHD = 38;

// Scratch for one z-chunk of the 3D MHD step.  All arrays are SoA [component][kk][j][i] with
// kk = k - kbase and `planes` allocated planes.
template <typename T>
struct MhdScratch {
  T* Q = nullptr;     // 8 components, primitive variables
  T* W = nullptr;     // NW_MHD components, traced state
  T* F = nullptr;     // 15 components: flux_x[5], flux_y[5], flux_z[5] at the LOW faces
  T* E = nullptr;     // 3 components: emf z, y, x at the LOW edges (reference order I_EMFZ=0..)
  T* EL = nullptr;    // 3 components: v x B at the LOW edges (x, y, z), input of the trace
  // hand-off records of the fused update (closing column / row of a tile published by its neighbours) and the tile
  // counter + per-tile progress flags; sized by fusedHandoffSize() for the largest launch of a chunk
  T* hbuf = nullptr;
  int* hsync = nullptr;
  size_t hbufReals = 0, hsyncInts = 0;
  T* strips = nullptr;  // shearing box on the fused kernels: F[15] + E[3] of the 4 x-border position columns, [18][kk][j][4]
  int planes = 0;     // allocated planes per component
  int kbase = 0;      // k of scratch plane 0 for the chunk being processed
  // TMA descriptor (CUtensorMap) of W for the fused flux+emf+update kernel; valid when fused != 0
  alignas(64) unsigned char mapW[128] = {0};
  int fused = 0;
};

template <typename T>
struct MhdKernels {
  // ghost fill of one direction (0,1,2), both faces; reference make_boundary2
  static void fillBoundary(const KParams<T>& P, T* U, int dir, int bcLo, int bcHi, bool skipLo, bool skipHi,
                           int kLo, int kHi, cudaStream_t s);
  // z ghost planes of the stratified shearing box (BC_Z_STRATIFIED, reference make_boundary_base.h:1357-1647) on the
  // faces this rank owns; floorDensity = [MRI] floor (no hydrostatic extrapolation)
  static void fillBoundaryZStratified(const KParams<T>& P, T* U, bool lo, bool hi, bool floorDensity, cudaStream_t s);
  // jet inflow patch in the lower ghost rows (2D) / planes (3D), after the ghost fill of the last direction;
  // reference make_jet, HydroRunBase.cpp:2374-2408 (hydro variables only)
  static void jetInflow(const KParams<T>& P, T* U, cudaStream_t s);
  // max over inner cells of the inverse time step -> *dMaxInvDt (ordered-uint encoding)
  static void computeInvDt(const KParams<T>& P, const T* U, unsigned long long* dMaxInvDt, cudaStream_t s);
  // 3D MHD chunk pipeline on planes [ka, kb) of the update range
  static void prim(const KParams<T>& P, const T* U, MhdScratch<T> sc, int k0, int k1, T dt, cudaStream_t s);
  static void elec(const KParams<T>& P, const T* U, MhdScratch<T> sc, int k0, int k1, cudaStream_t s);
  static void trace(const KParams<T>& P, const T* U, MhdScratch<T> sc, int k0, int k1, T dt, cudaStream_t s);
  static void flux(const KParams<T>& P, MhdScratch<T> sc, int k0, int k1, cudaStream_t s);
  static void emf(const KParams<T>& P, MhdScratch<T> sc, int k0, int k1, cudaStream_t s);
  static void update(const KParams<T>& P, const T* Uold, T* Unew, MhdScratch<T> sc, int k0, int k1, T dt,
                     unsigned long long* dMaxInvDt, cudaStream_t s);
  // fused flux + emf + update (TMA-staged W tiles, z-marching blocks) for the FAST configuration:
  // fusedPrepare() encodes the tensor map of sc.W into sc.mapW and sets sc.fused when the run
  // parameters and the array shapes qualify; fusedFluxEmfUpdate() replaces flux()+emf()+update()
  static void fusedPrepare(const KParams<T>& P, MhdScratch<T>& sc);
  // what fusedPrepare() will decide from the run parameters alone (before any scratch exists)
  static bool fusedUpdateEligible(const KParams<T>& P);
  // reals / ints of the hand-off buffers a fused update of `planes` update planes needs (0 with knob fused_handoff = 0)
  static void fusedHandoffSize(const KParams<T>& P, int planes, size_t* reals, size_t* ints);
  // rotating frame (Omega0 > 0, HLLD + 2-D HLLD): the FAST = false instantiation with update_cell_rot; with shearing-box
  // boundaries (shearEnabled; jplus, frac = y shift of the opposite border) the cells next to the x borders are updated
  // by a small second kernel from the border strips (sc.strips)
  static void fusedFluxEmfUpdate(const KParams<T>& P, const T* Uold, T* Unew, const MhdScratch<T>& sc, int ka, int kb,
                                 T dt, unsigned long long* dMaxInvDt, cudaStream_t s, int shearEnabled = 0, int jplus = 0,
                                 T frac = T(0));
  // fused cons->prim + edge electric field + trace (U -> W, shared-memory rings, z-marching blocks) for the
  // FAST configuration; replaces prim() + elec() + trace() on traced planes [k0, k1)
  static bool fusedTraceAvailable(const KParams<T>& P);
  static void fusedTrace(const KParams<T>& P, const T* U, MhdScratch<T> sc, int k0, int k1, T dt, cudaStream_t s);
  // x/y ghost cells of planes [k0,k1) keep the old values (the fused kernel only writes the update box)
  static void copyOutsideBox(const KParams<T>& P, const T* Uold, T* Unew, int k0, int k1, cudaStream_t s);
  // rotating frame / shearing box variants (Omega0 > 0): update with Coriolis + border remap, and the
  // y-shifted x ghost cells; (jplus, frac) = whole cells and fraction of dy of the border shift
  static void updateRotating(const KParams<T>& P, const T* Uold, T* Unew, MhdScratch<T> sc, int k0, int k1, T dt,
                             int shearEnabled, int jplus, T frac, unsigned long long* dMaxInvDt, cudaStream_t s);
  static void shearGhosts(const KParams<T>& P, T* U, int jplus, T frac, int k0, int k1, cudaStream_t s);  // planes [k0, k1)
  // copy planes [k0,k1) of every variable (ghost planes that the update does not touch)
  static void copyPlanes(const KParams<T>& P, const T* Uold, T* Unew, int k0, int k1, cudaStream_t s);
  // device probes for known-answer tests (n independent problems, arrays are [n][...])
  static void probeRiemann(const KParams<T>& P, int n, const T* ql, const T* qr, T* flux, cudaStream_t s);
  static void probeEmf(const KParams<T>& P, int n, int emfDir, const T* qEdge, const T* xPos, T* emf,
                       cudaStream_t s);
};

// 2D MHD traced state: 8 cell centred + 4 face B + 14 half slopes + 4 face-field half slopes
constexpr int NW_MHD2D = 30;

template <typename T>
struct Mhd2dKernels {
  // scratch: Q[8], W[NW_MHD2D], F[12], E[1], each [comp][j][i]
  static void step(const KParams<T>& P, const T* Uold, T* Unew, T* Q, T* W, T* F, T* E, T dt,
                   unsigned long long* slots, cudaStream_t s);
};

// hydro traced state: 5 cell-centred primitives advanced by dt/2 + 15 half slopes
constexpr int NW_HYDRO = 20;

template <typename T>
struct HydroKernels {
  // W is [NW_HYDRO][planes][j][i] with kk = k - kbase
  static void trace(const KParams<T>& P, const T* U, T* W, int planes, int kbase, int k0, int k1, T dt, cudaStream_t s);
  static void fluxUpdate(const KParams<T>& P, const T* Uold, T* Unew, const T* W, int planes, int kbase, int k0, int k1,
                         T dt, unsigned long long* slots, cudaStream_t s);
  static void computeInvDt(const KParams<T>& P, const T* U, unsigned long long* slots, cudaStream_t s);
  static void probeRiemann(const KParams<T>& P, int n, const T* ql, const T* qr, T* flux, cudaStream_t s);
  // the whole step of planes [k0, k1) in ONE kernel (kernels_hydro3d_fused.cu): U -> Unew, no W scratch
  static void fusedStep(const KParams<T>& P, const T* Uold, T* Unew, int k0, int k1, T dt, unsigned long long* slots,
                        cudaStream_t s);
  // x/y ghost cells of planes [k0, k1) keep the old values (the tiled kernels only write inner cells)
  static void copyGhosts(const KParams<T>& P, const T* Uold, T* Unew, int k0, int k1, cudaStream_t s);
};
// "hydro_fused" knob (default on): one-kernel hydro step
bool hydroFusedRequested();

// 2D hydro traced state: 4 cell-centred primitives advanced by dt/2 + 8 half slopes
constexpr int NW_HYDRO2D = 12;

template <typename T>
struct Hydro2dKernels {
  // one step U -> Unew of the 2D Euler solver (kernels_hydro2d.cu); W is [NW_HYDRO2D][j][i]
  static void step(const KParams<T>& P, const T* Uold, T* Unew, T* W, T dt, unsigned long long* slots, cudaStream_t s);
  static void computeInvDt(const KParams<T>& P, const T* U, unsigned long long* slots, cudaStream_t s);
};

// dissipative terms on the NEW state after the Godunov update (kernels_dissipative.cu).
// D is a scratch array of 12 components over the whole local array [c][k][j][i].
template <typename T>
struct DissKernels {
  static void resistEmf(const KParams<T>& P, const T* U, T* D, cudaStream_t s);     // D[0..2] = -eta curl B
  static void ctUpdate(const KParams<T>& P, T* U, const T* D, T dt, cudaStream_t s);  // U.B += dt curl-difference of D[0..2]
  static void resistEnergy(const KParams<T>& P, T* U, T dt, cudaStream_t s);         // U.E -= div(eta J x B) dt
  static void viscFlux(const KParams<T>& P, const T* U, T* D, T dt, cudaStream_t s);  // D[dir*4 + (mx,my,mz,E)]
  static void viscUpdate(const KParams<T>& P, T* U, const T* D, cudaStream_t s);
};

// history diagnostics of a 3D MHD state (kernels_history.cu): two reduction passes with a fixed order
template <typename T>
struct HistoryKernels {
  static void columnSums(const KParams<T>& P, const T* U, double* partial /*[nBlocks][3][isize]*/, int nBlocks, cudaStream_t s);
  static void sums(const KParams<T>& P, const T* U, const double* meanUV /*[2][isize]*/, double* partial /*[nBlocks][8]*/,
                   int nBlocks, cudaStream_t s);
};

// number of slots of the inverse-dt max reduction (power of two); every slot holds the bit pattern
// of a non-negative double, so "max" works on the integer or on the floating view alike
constexpr int MAX_SLOTS = 1024;

// ordered encoding of a non-negative floating value for atomicMax
inline double decodeMax(unsigned long long v) {
  double d;
  static_assert(sizeof(d) == sizeof(v), "size");
  __builtin_memcpy(&d, &v, sizeof d);
  return d;
}

// launch counter (bench.py's gpu_launches): every wrapper above bumps it once per kernel launch
unsigned long long kernelLaunchCount();
// occupancy knobs: "flux_minb" | "emf_minb" | "trace_minb" | "update_minb" = 2..8 resident blocks/SM
bool setTuning(const char* key, int value);
// "fused_b" knob (default on): use the fused flux+emf+update kernel when MhdScratch::fused is set
bool fusedRequested();
// "rot_dt" knob (default off): the rotating-frame fused kernel reduces the inverse dt of the new state itself
bool rotDtInKernel();
// "fused_handoff" knob (default off): 16 x 8 hand-off tiles of the fused update
bool fusedHandoffRequested();
// "fused_a" knob (default on): fused prim+elec+trace kernel
bool fusedTraceRequested();
void resetKernelLaunchCount();

}  // namespace rg
