// 3D Euler (hydro) Godunov step for sm_100a, FP32 and FP64: two kernels per z-chunk.
//   trace      : U (7-point) -> cons->prim on the fly, TVD slopes, half-step predictor -> W (20 comps)
//   fluxUpdate : per cell, the six face fluxes from W of the cell and its six neighbours (each face
//                flux is evaluated by both adjacent cells with identical inputs and code, hence
//                bitwise identical and conservative), conservative update, inverse dt of the new state
// Reference: HydroRunGodunov.cpp:2658-2890 (godunov_unsplit_cpu_v1, 3D), trace.h:544-661,
// slope.h:324-427, riemann.h, HydroRunBase.cpp:386-426.  Where the reference stores qm/qp x3
// (30 reals per cell) plus Q, this keeps 20 reals of traced state and no primitive array.
#include "hydro_cells.cuh"
#include "hydro_device.cuh"
#include "kernel_common.cuh"
#include "kernels.h"

namespace rg {

int g_hydroTile = 1;  // run-time knob "hydro_tile": register-tiled flux+update kernel (default) or the gather variant

namespace {

template <typename T>
__global__ void __launch_bounds__(BX) k_hydro_trace(const __grid_constant__ KParams<T> P, const T* __restrict__ Uin,
                                                    T* __restrict__ Wp, int planes, int kbase, int k0, T dt) {
  int i, j;
  const int k = k0 + blockIdx.z;
  if (!tileCoords(1, P.isize - 2, 1, P.jsize - 2, i, j)) return;
  const UView<T> U = uview(Uin, P);
  const View<T> W = view(Wp, P, planes, kbase);
  hydro_trace_cell(P, U, W, i, j, k, dt);
}


// flux + conservative update + next dt, z-marching: a 32 x 4 thread block owns a column of cells and
// walks a range of planes.  Every face flux is evaluated ONCE per tile: a thread solves the Riemann
// problems of its three LOW faces (the z one is carried over from the previous plane, where it was
// the high face), takes the high x flux from the next lane (warp shuffle), the high y flux from the
// next row (shared memory) and only the tile's closing faces are solved a second time by the
// neighbouring tile -- with identical inputs and code, hence bitwise identical and conservative.
// 3.3 Riemann problems per cell instead of 6.
template <typename T, int RS>
__global__ void __launch_bounds__(128) k_hydro_flux_update(const __grid_constant__ KParams<T> P, const T* __restrict__ Uold,
                                                           T* __restrict__ Unew, const T* __restrict__ Wp, int planes,
                                                           int kbase, int k0, int k1, int lzc, T dt,
                                                           unsigned long long* __restrict__ slots) {
  __shared__ T sfy[4][32][5];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int i = blockIdx.x * 32 + tx, j = blockIdx.y * 4 + ty;
  const int za = k0 + blockIdx.z * lzc, zb = min(za + lzc, k1);
  const int gw = P.gw;
  const bool valid = i < P.isize && j < P.jsize;
  const bool inI = i >= gw && i < P.isize - gw, inJ = j >= gw && j < P.jsize - gw;
  const bool innerXY = valid && inI && inJ;
  const bool faceX = valid && i >= gw && i <= P.isize - gw && inJ;  // low x face of the cell is a face of an inner cell
  const bool faceY = valid && j >= gw && j <= P.jsize - gw && inI;
  const UView<T> U = uview(Uold, P);
  const View<const T> W = view<const T>(Wp, P, planes, kbase);
  const size_t plane = (size_t)P.isize * P.jsize, comp = plane * P.ksize;
  const T dtdx = dt / P.dx, dtdy = dt / P.dy, dtdz = dt / P.dz;
  T invDt = T(0);
  T fzl[5];
  bool haveZ = false;
  for (int k = za; k < zb; ++k) {
    const bool inK = k >= gw && k < P.ksize - gw;  // block-uniform
    const size_t idx = (size_t)k * plane + (size_t)j * P.isize + i;
    T un[5];
    if (valid) {
#pragma unroll
      for (int v = 0; v < 5; ++v) un[v] = U(v, i, j, k);
    }
    if (inK) {
      T fxl[5], fyl[5], fzh[5], fxh[5], fyh[5];
#pragma unroll
      for (int v = 0; v < 5; ++v) { fxl[v] = T(0); fyl[v] = T(0); fzh[v] = T(0); }
      if (faceX) hydro_low_flux<T, 0, RS>(P, W, i, j, k, fxl);
      if (faceY) hydro_low_flux<T, 1, RS>(P, W, i, j, k, fyl);
      if (innerXY) {
        if (!haveZ) hydro_low_flux<T, 2, RS>(P, W, i, j, k, fzl);
        hydro_low_flux<T, 2, RS>(P, W, i, j, k + 1, fzh);
      }
#pragma unroll
      for (int v = 0; v < 5; ++v) {
        fxh[v] = __shfl_down_sync(0xffffffffu, fxl[v], 1);
        sfy[ty][tx][v] = fyl[v];
      }
      __syncthreads();
      if (innerXY) {
        if (tx == 31) hydro_low_flux<T, 0, RS>(P, W, i + 1, j, k, fxh);
        if (ty == 3) {
          hydro_low_flux<T, 1, RS>(P, W, i, j + 1, k, fyh);
        } else {
#pragma unroll
          for (int v = 0; v < 5; ++v) fyh[v] = sfy[ty + 1][tx][v];
        }
#pragma unroll
        for (int v = 0; v < 5; ++v) {  // summation order of the reference's serial scatter
          T s = un[v];
          s += fxl[v] * dtdx; s += fyl[v] * dtdy; s += fzl[v] * dtdz;
          s -= fxh[v] * dtdx; s -= fyh[v] * dtdy; s -= fzh[v] * dtdz;
          un[v] = s;
        }
        if (P.gravity) {  // static gravity source term, reference HydroRunBase.cpp:1962-1976
          const T hdt = T(0.5) * dt, rs = U(ID, i, j, k) + un[ID];
          un[IU] += hdt * P.gx * rs; un[IV] += hdt * P.gy * rs; un[IW] += hdt * P.gz * rs;
        }
        T q[5];
        const T c = dev::cons_to_prim_hydro(P, un[ID], un[IP], un[IU], un[IV], un[IW], q);
        invDt = dev::mx(invDt, (c + dev::ab(q[IU])) * P.rdx + (c + dev::ab(q[IV])) * P.rdy + (c + dev::ab(q[IW])) * P.rdz);
#pragma unroll
        for (int v = 0; v < 5; ++v) fzl[v] = fzh[v];  // the high z face is the low face of the next plane
      }
      haveZ = true;
      __syncthreads();  // sfy is rewritten by the next plane
    } else {
      haveZ = false;
    }
    if (valid) {
#pragma unroll
      for (int v = 0; v < 5; ++v) Unew[v * comp + idx] = un[v];
    }
  }
  if (slots != nullptr) reduceMaxToSlots(invDt, slots);
}


// ------------------------------------------------------------------------------------------------
// flux + conservative update + next dt, register-tiled ("tile" variant, default): a 32 x HR thread
// block marches along z over a (32-2) x (HR-2) column of updated cells surrounded by a one-cell halo.
// Every thread loads the 20 W components of ITS OWN cell once per plane (coalesced, each W value is
// read from HBM once per tile instead of six times through L1) and builds its six face states in
// registers.  The left state of a face comes from the neighbour: lane-1 by warp shuffle (x), row-1
// through shared memory (y), the thread's own high-z state of the previous plane by register carry
// (z).  Every face flux of the tile is solved exactly once (no serial closing-face solves); the high
// z flux of a cell is the low z flux of the next plane, so the update of plane p is finished one
// iteration later.  Faces on tile seams are solved by both tiles with identical inputs and code,
// hence bitwise identical and conservative.  3.4 Riemann problems per updated cell.
// ------------------------------------------------------------------------------------------------
constexpr int HR = 8;  // rows of the thread block

template <typename T, int RS>
__global__ void __launch_bounds__(32 * HR) k_hydro_flux_update_tile(const __grid_constant__ KParams<T> P,
                                                                    const T* __restrict__ Uold, T* __restrict__ Unew,
                                                                    const T* __restrict__ Wp, int planes, int kbase,
                                                                    int k0, int k1, int lzc, T dt,
                                                                    unsigned long long* __restrict__ slots) {
  __shared__ T sy[HR][5][32];  // high-y face states, then low-y fluxes, of the current plane
  const int tx = threadIdx.x, ty = threadIdx.y, gw = P.gw;
  const int i = gw + blockIdx.x * 30 - 1 + tx, j = gw + blockIdx.y * (HR - 2) - 1 + ty;
  const int za = k0 + blockIdx.z * lzc, zb = min(za + lzc, k1);  // updated planes [za, zb)
  const int iN = P.isize - gw, jN = P.jsize - gw;
  const bool cellOK = i <= P.isize - 2 && j <= P.jsize - 2;  // W exists (i, j >= 1 by construction)
  const bool rowUpd = ty >= 1 && ty <= HR - 2 && j < jN, colUpd = tx >= 1 && tx <= 30 && i < iN;
  const bool upd = rowUpd && colUpd;
  const bool solveX = tx >= 1 && i <= iN && rowUpd;  // low x face of the cell is a face of an updated cell of this tile
  const bool solveY = ty >= 1 && j <= jN && colUpd;
  const View<const T> W = view<const T>(Wp, P, planes, kbase);
  const size_t plane = (size_t)P.isize * P.jsize, comp = plane * P.ksize;
  const T dtdx = dt / P.dx, dtdy = dt / P.dy, dtdz = dt / P.dz;
  T invDt = T(0);
  dev::HState<T> zPrev{T(1), T(1), T(0), T(0), T(0)};  // high-z face state of the previous plane
  T acc[5] = {T(0), T(0), T(0), T(0), T(0)};            // update of the previous plane, all terms but the high z flux
  for (int p = za - 1; p <= zb; ++p) {
    T w[NW_HYDRO];
#pragma unroll
    for (int c = 0; c < NW_HYDRO; ++c) w[c] = cellOK ? W(c, i, j, p) : T(1);
    const bool mid = p >= za && p < zb;  // block-uniform
    T un[5];
    if (mid && upd) {
      const size_t idx = (size_t)p * plane + (size_t)j * P.isize + i;
#pragma unroll
      for (int v = 0; v < 5; ++v) un[v] = __ldg(Uold + v * comp + idx);
    }
    T fz[5] = {T(0), T(0), T(0), T(0), T(0)};
    if (p >= za && upd) face_flux<T, 2, RS>(P, zPrev, face_from_regs<T, 2>(P, w, T(-1)), fz);
    if (p > za && upd) {  // the low z flux of this plane closes the update of the plane below
      const size_t idx = (size_t)(p - 1) * plane + (size_t)j * P.isize + i;
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
    if (!mid) continue;  // block-uniform: first (za-1) and last (zb) planes only feed the z faces
    // x faces: left state from lane-1, high flux from lane+1
    T fxl[5] = {T(0), T(0), T(0), T(0), T(0)}, fxh[5];
    {
      const dev::HState<T> hi = face_from_regs<T, 0>(P, w, T(1));
      dev::HState<T> L;
      L.r = __shfl_up_sync(0xffffffffu, hi.r, 1); L.p = __shfl_up_sync(0xffffffffu, hi.p, 1);
      L.u = __shfl_up_sync(0xffffffffu, hi.u, 1); L.v = __shfl_up_sync(0xffffffffu, hi.v, 1);
      L.w = __shfl_up_sync(0xffffffffu, hi.w, 1);
      if (solveX) face_flux<T, 0, RS>(P, L, face_from_regs<T, 0>(P, w, T(-1)), fxl);
#pragma unroll
      for (int v = 0; v < 5; ++v) fxh[v] = __shfl_down_sync(0xffffffffu, fxl[v], 1);
    }
    // y faces: left state from row-1, high flux from row+1, both through shared memory
    T fyl[5] = {T(0), T(0), T(0), T(0), T(0)}, fyh[5];
    {
      const dev::HState<T> hi = face_from_regs<T, 1>(P, w, T(1));
      sy[ty][0][tx] = hi.r; sy[ty][1][tx] = hi.p; sy[ty][2][tx] = hi.u; sy[ty][3][tx] = hi.v; sy[ty][4][tx] = hi.w;
      __syncthreads();
      if (solveY) {
        const dev::HState<T> L{sy[ty - 1][0][tx], sy[ty - 1][1][tx], sy[ty - 1][2][tx], sy[ty - 1][3][tx], sy[ty - 1][4][tx]};
        face_flux<T, 1, RS>(P, L, face_from_regs<T, 1>(P, w, T(-1)), fyl);
      }
      __syncthreads();
#pragma unroll
      for (int v = 0; v < 5; ++v) sy[ty][v][tx] = fyl[v];
      __syncthreads();
#pragma unroll
      for (int v = 0; v < 5; ++v) fyh[v] = (ty < HR - 1) ? sy[ty + 1][v][tx] : T(0);
      __syncthreads();  // sy is rewritten by the next plane
    }
    if (upd) {
#pragma unroll
      for (int v = 0; v < 5; ++v) {  // summation order of the reference's serial scatter
        T s = un[v];
        s += fxl[v] * dtdx; s += fyl[v] * dtdy; s += fz[v] * dtdz;
        s -= fxh[v] * dtdx; s -= fyh[v] * dtdy;
        acc[v] = s;
      }
    }
  }
  if (slots != nullptr) reduceMaxToSlots(invDt, slots);
}

// x/y ghost cells of planes [k0, k1) keep the old values (the tile kernel only writes inner cells)
template <typename T>
__global__ void __launch_bounds__(256) k_hydro_copy_ghosts(const __grid_constant__ KParams<T> P, const T* __restrict__ Uold,
                                                           T* __restrict__ Unew, int k0) {
  const int gw = P.gw, ng = 2 * gw;
  const int nRowCells = ng * P.isize, nColCells = ng * (P.jsize - ng);
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nRowCells + nColCells) return;
  int i, j;
  if (t < nRowCells) {
    const int r = t / P.isize;
    i = t - r * P.isize;
    j = (r < gw) ? r : P.jsize - ng + r;
  } else {
    const int q = t - nRowCells, r = q / ng, c = q - r * ng;
    j = gw + r;
    i = (c < gw) ? c : P.isize - ng + c;
  }
  const size_t plane = (size_t)P.isize * P.jsize, comp = plane * P.ksize;
  const size_t idx = (size_t)(k0 + blockIdx.y) * plane + (size_t)j * P.isize + i;
  for (int v = 0; v < P.nvar; ++v) Unew[v * comp + idx] = Uold[v * comp + idx];
}

template <typename T>
__global__ void __launch_bounds__(BX) k_hydro_invdt(const __grid_constant__ KParams<T> P, const T* __restrict__ Uin,
                                                    unsigned long long* __restrict__ slots) {
  int i, j;
  const bool valid = tileCoords(P.gw, P.nx, P.gw, P.ny, i, j);
  const int k = P.gw + blockIdx.z;
  T invDt = T(0);
  if (valid) {
    const UView<T> U = uview(Uin, P);
    T q[5];
    const T c = dev::cons_to_prim_hydro(P, U(ID, i, j, k), U(IP, i, j, k), U(IU, i, j, k), U(IV, i, j, k), U(IW, i, j, k), q);
    invDt = (c + dev::ab(q[IU])) * P.rdx + (c + dev::ab(q[IV])) * P.rdy + (c + dev::ab(q[IW])) * P.rdz;
  }
  reduceMaxToSlots(invDt, slots);
}

template <typename T>
__global__ void k_probe_riemann_hydro(const __grid_constant__ KParams<T> P, int n, const T* ql, const T* qr, T* flux) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const T *l = ql + 5 * t, *r = qr + 5 * t;
  dev::HState<T> L{l[ID], l[IP], l[IU], l[IV], l[IW]}, R{r[ID], r[IP], r[IU], r[IV], r[IW]};
  T f[5];
  dev::riemann_hydro(P, L, R, f);
  for (int v = 0; v < 5; ++v) flux[5 * t + v] = f[v];
}

}  // namespace

template <typename T>
void HydroKernels<T>::trace(const KParams<T>& P, const T* U, T* W, int planes, int kbase, int k0, int k1, T dt, cudaStream_t s) {
  if (k1 <= k0) return;
  k_hydro_trace<T><<<gridFor(P.isize - 2, P.jsize - 2, k1 - k0), blockShape(), 0, s>>>(P, U, W, planes, kbase, k0, dt);
  launched();
}
template <typename T>
void HydroKernels<T>::fluxUpdate(const KParams<T>& P, const T* Uold, T* Unew, const T* W, int planes, int kbase, int k0,
                                 int k1, T dt, unsigned long long* slots, cudaStream_t s) {
  if (k1 <= k0) return;
  if (g_hydroTile) {  // register-tiled variant (default): z ranges of up to 64 planes
    const int lzc = std::min(k1 - k0, 64);
    const dim3 g((P.nx + 29) / 30, (P.ny + HR - 3) / (HR - 2), (k1 - k0 + lzc - 1) / lzc), b(32, HR, 1);
    switch (P.riemannSolver) {
      case RS_HLLC: k_hydro_flux_update_tile<T, RS_HLLC><<<g, b, 0, s>>>(P, Uold, Unew, W, planes, kbase, k0, k1, lzc, dt, slots); break;
      case RS_HLL: k_hydro_flux_update_tile<T, RS_HLL><<<g, b, 0, s>>>(P, Uold, Unew, W, planes, kbase, k0, k1, lzc, dt, slots); break;
      case RS_APPROX: k_hydro_flux_update_tile<T, RS_APPROX><<<g, b, 0, s>>>(P, Uold, Unew, W, planes, kbase, k0, k1, lzc, dt, slots); break;
      default: k_hydro_flux_update_tile<T, -1><<<g, b, 0, s>>>(P, Uold, Unew, W, planes, kbase, k0, k1, lzc, dt, slots); break;
    }
    launched();
    copyGhosts(P, Uold, Unew, k0, k1, s);
    return;
  }
  // z ranges of about 32 planes (one redundant z face per range), at least a few waves of blocks
  const int lzc = std::min(k1 - k0, 32);
  const dim3 g((P.isize + 31) / 32, (P.jsize + 3) / 4, (k1 - k0 + lzc - 1) / lzc), b(32, 4, 1);
  // one instantiation per Riemann solver: a single solver body in the kernel instead of three per face
  switch (P.riemannSolver) {
    case RS_HLLC: k_hydro_flux_update<T, RS_HLLC><<<g, b, 0, s>>>(P, Uold, Unew, W, planes, kbase, k0, k1, lzc, dt, slots); break;
    case RS_HLL: k_hydro_flux_update<T, RS_HLL><<<g, b, 0, s>>>(P, Uold, Unew, W, planes, kbase, k0, k1, lzc, dt, slots); break;
    case RS_APPROX: k_hydro_flux_update<T, RS_APPROX><<<g, b, 0, s>>>(P, Uold, Unew, W, planes, kbase, k0, k1, lzc, dt, slots); break;
    default: k_hydro_flux_update<T, -1><<<g, b, 0, s>>>(P, Uold, Unew, W, planes, kbase, k0, k1, lzc, dt, slots); break;
  }
  launched();
}
template <typename T>
void HydroKernels<T>::copyGhosts(const KParams<T>& P, const T* Uold, T* Unew, int k0, int k1, cudaStream_t s) {
  if (k1 <= k0) return;
  const int ng = 2 * P.gw, cells = ng * P.isize + ng * (P.jsize - ng);
  k_hydro_copy_ghosts<T><<<dim3((cells + 255) / 256, k1 - k0, 1), 256, 0, s>>>(P, Uold, Unew, k0);
  launched();
}
template <typename T>
void HydroKernels<T>::computeInvDt(const KParams<T>& P, const T* U, unsigned long long* slots, cudaStream_t s) {
  k_hydro_invdt<T><<<gridFor(P.nx, P.ny, P.nz), blockShape(), 0, s>>>(P, U, slots);
  launched();
}
template <typename T>
void HydroKernels<T>::probeRiemann(const KParams<T>& P, int n, const T* ql, const T* qr, T* flux, cudaStream_t s) {
  k_probe_riemann_hydro<T><<<(n + 127) / 128, 128, 0, s>>>(P, n, ql, qr, flux);
  launched();
}

template struct HydroKernels<double>;
template struct HydroKernels<float>;

}  // namespace rg
