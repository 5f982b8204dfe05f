/*
 * ramsesgpu_b200.h -- C ABI of the B200-native Godunov update path.
 *
 * The reference (pkestene/ramsesGPU) has no plugin/FFI layer: callers (src/euler_main.cpp:176-195,
 * src/glutGui/HydroWindow.cpp:66-70,585, src/qtGui/qtHydro2d/HydroWidget.cpp:62-72) use the C++
 * virtual surface of HydroRunBase / MHDRunBase / MHDRunGodunov / HydroRunGodunov.  Each entry point
 * below replaces one of those virtuals (cited per function); include/ramsesgpu_b200_shim.hpp wraps
 * them back into C++ classes with the reference's names so that a reference main() links unchanged.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a non-zero
 * rg_status otherwise, with a message available from rg_last_error() (thread-local).  The handle
 * owns all device memory.  Calls on one handle are not re-entrant.  There is NO CPU fallback:
 * rg_create fails with RG_ERR_NO_DEVICE when no CUDA device is present.
 *
 * State arrays exchanged with the host are the reference's HostArray layout
 * (src/hydro/Arrays.h:95-98): SoA [var][k][j][i], x fastest, ghosts included, variable order
 * ID, IP(E_tot), IU, IV, IW, IA, IB, IC (constants.h:59-71; B = LEFT-face values).
 * For a multi-GPU run (z-slab decomposition) every array is the LOCAL slab of the calling rank.
 */
#ifndef RAMSESGPU_B200_H_
#define RAMSESGPU_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rg_run_s* rg_handle;

typedef enum {
  RG_OK = 0,
  RG_ERR_INVALID = 1,    /* bad argument / handle */
  RG_ERR_NO_DEVICE = 2,  /* no CUDA device: the product path refuses to run */
  RG_ERR_CUDA = 3,
  RG_ERR_UNSUPPORTED = 4,
  RG_ERR_IO = 5,
  RG_ERR_NCCL = 6
} rg_status;

enum { RG_FLAG_FP32 = 1 }; /* solver precision is a create flag (reference: build-time real_t, real_type.h:27-31) */

typedef struct rg_layout {
  int nx, ny, nz;          /* global inner sizes */
  int isize, jsize, ksize; /* local sizes with ghosts */
  int nvar, ghost_width, dim, mhd, real_bytes;
  int nz_local, k_offset;  /* this rank's slab: inner planes and global index of the first one */
  int rank, nranks;
} rg_layout;

typedef struct rg_stats {
  unsigned long long kernel_launches; /* kernels launched by this library since load / last reset */
  double last_step_ms;                /* device time of the last rg_godunov_unsplit (CUDA events) */
  double halo_bytes_per_step;         /* bytes this rank sends to its z-neighbours per ghost fill */
  size_t device_bytes;                /* device memory owned by the handle */
  int chunk_planes;                   /* z planes per chunk of the step pipeline */
  int halo_peer_copies;               /* 1: the z halo travels by copy engines over peer-mapped memory; 0: NCCL send/recv (or one rank) */
} rg_stats;

const char* rg_last_error(void);
const char* rg_version(void);
int rg_device_count(void);

/* MHDRunGodunov(ConfigMap&) / HydroRunGodunov(ConfigMap&) (MHDRunGodunov.cpp:75, HydroRunGodunov.cpp:60;
 * the solver family is picked by [MHD] enable like src/euler_main.cpp:109).  ini_text is the
 * CONTENT of a parameter file (reference ConfigMap(buffer) ctor, ConfigMap.cpp:28). */
int rg_create(const char* ini_text, int flags, rg_handle* out);
int rg_create_from_file(const char* ini_path, int flags, rg_handle* out);
/* multi-GPU: one process per GPU, z-slab `rank` of `nranks`; nccl_unique_id = 128 bytes from
 * rg_nccl_unique_id() on rank 0, broadcast by the caller (replaces HydroMpiParameters' Cartesian
 * communicator, HydroMpiParameters.cpp:38-239, for mx = my = 1, mz = nranks).  device < 0 = current. */
int rg_create_distributed(const char* ini_text, int flags, int rank, int nranks, const void* nccl_unique_id,
                          int device, rg_handle* out);
int rg_nccl_unique_id(void* out128);
int rg_destroy(rg_handle h);

int rg_get_layout(rg_handle h, rg_layout* out);
/* derived parameters (HydroParameters.h:166-525), by lower-case name: dx, dy, dz, gamma0, cfl, smallr,
 * smallc, smallp, ciso, omega0, slope_type, tend, nstepmax, noutput, xmin ... zmax */
int rg_get_param(rg_handle h, const char* name, double* value);

/* init_simulation(problem) (MHDRunBase.cpp:1231, HydroRunBase.cpp:7023); problem NULL/"" = [hydro] problem */
int rg_init_simulation(rg_handle h, const char* problem, int* nStep);
/* The host half of init_simulation alone: initial condition of z-slab `rank` of `nranks` written into a caller
 * buffer ([var][k][j][i] of the local slab, ghosts included), no device involved (a coupler that builds its own
 * state, or a check of the problem setup on a machine without a GPU).  dst == NULL only fills *layout_out. */
int rg_initial_condition_host(const char* ini_text, int flags, int rank, int nranks, void* dst, size_t bytes,
                              rg_layout* layout_out);
/* make_all_boundaries(U) (HydroRunBase.h:422): which = 0 -> U, 1 -> U2 */
int rg_make_all_boundaries(rg_handle h, int which);
/* compute_dt(int useU) / compute_dt_mhd(int useU) (HydroRunBase.h:80, MHDRunBase.h:49); global over all slabs */
int rg_compute_dt(rg_handle h, int useU, double* dt);
/* godunov_unsplit(int nStep, real_t dt) (MHDRunGodunov.h:120): nStep even U -> U2, odd U2 -> U */
int rg_godunov_unsplit(rg_handle h, int nStep, double dt);
/* oneStepIntegration(int& nStep, real_t& t, real_t& dt) (HydroRunBase.h:433, MHDRunGodunov.cpp:4077-4089) */
int rg_one_step(rg_handle h, int* nStep, double* t, double* dt);
/* start() (MHDRunGodunov.cpp:3801): init, loop while (t < tEnd && nStep < nStepmax), outputs, perf line */
int rg_run(rg_handle h);
/* output(U, nStep) (HydroRunBase.cpp:4348): raw .vti and/or .xsm as the [output] section asks */
int rg_output(rg_handle h, int nStep);

/* history(nStep, dt) diagnostics (MHDRunBase.cpp:3311 history_default, :3476 history_mri), reduced on the device and
 * summed over all slabs; 3D MHD only.  out[8] = mass, maxwell, reynolds, magp, mean_Bx, mean_By, mean_Bz, divB.
 * rg_run writes the reference's history file when [history] enabled=yes. */
int rg_history(rg_handle h, int nStep, double* out8);

/* getData(nStep) / copyGpuToCpu(nStep) / getDataHost (HydroRunBase.h:437,456,512) */
int rg_get_data_device(rg_handle h, int which, void** device_ptr);
int rg_copy_to_host(rg_handle h, int which, void* dst, size_t bytes);
int rg_copy_from_host(rg_handle h, int which, const void* src, size_t bytes);
int rg_synchronize(rg_handle h);

/* host-buffer round trip used for end-to-end timing: H2D(host_in -> U), n_steps x rg_one_step from
 * (nStep, t) = (0, 0), D2H(result -> host_out) */
int rg_steps_from_host(rg_handle h, const void* host_in, void* host_out, size_t bytes, int n_steps,
                       double* t_out, double* dt_last);

/* n_jobs INDEPENDENT one-step jobs (ensemble members, parameter scans sharing one .ini):
 * host_out[j] = one rg_one_step of host_in[j] from (nStep, t) = (0, 0); dt_out[j] (optional) = the
 * step's dt.  Same results, bit for bit, as n_jobs calls of rg_steps_from_host(.., 1, ..), but the H2D
 * copy of job j+1, the step of job j and the D2H copy of job j-1 run concurrently on three streams
 * (two device buffer pairs in rotation; pinned host buffers are needed for the overlap).  A host
 * buffer may be reused by jobs j and j+2. */
int rg_steps_from_host_batch(rg_handle h, int n_jobs, const void* const* host_in, void* const* host_out,
                             size_t bytes, double* dt_out);

/* page-locked host memory for the host-buffer calls above (cudaHostAlloc / cudaFreeHost) */
int rg_alloc_pinned(size_t bytes, void** out);
int rg_free_pinned(void* p);

int rg_get_stats(rg_handle h, rg_stats* out);
int rg_reset_launch_count(void);
/* z planes per pipeline chunk (0 = automatic: whole slab if the scratch fits in device memory) */
int rg_set_chunk_planes(rg_handle h, int planes);
/* multi-GPU: exchange the z halo of the buffer being written as soon as the slab's boundary planes
 * are final, on a second stream, overlapped with the interior update (default on) */
int rg_set_halo_overlap(rg_handle h, int on);

/* occupancy knobs of the FP64 kernels (process-wide): key = "flux_minb" | "emf_minb" | "trace_minb" |
 * "update_minb", value = 2..8 resident blocks per SM the kernel variant is compiled for; "tile_x" = 32|64|128;
 * "fused_b" = 0|1 selects the separate flux/emf/update kernels or the fused TMA kernel (default 1) */
int rg_set_tuning(const char* key, int value);

/* device timing of a region on the library's stream (CUDA events): total and per kernel family.
 * phase order: RG_PHASE_* below.  Events are recorded around every launch between begin and end;
 * rg_profile_end synchronises and returns milliseconds and launch counts summed over the region. */
enum { RG_PHASE_BOUNDARY = 0, RG_PHASE_PRIM, RG_PHASE_TRACE, RG_PHASE_FLUX, RG_PHASE_EMF, RG_PHASE_UPDATE,
       RG_PHASE_DT, RG_PHASE_COPY, RG_PHASE_HALO, RG_PHASE_FUSED /* fused flux+emf+update */,
       RG_PHASE_DISS /* resistivity + viscosity */, RG_NPHASE };
int rg_profile_begin(rg_handle h);
int rg_profile_end(rg_handle h, double* total_ms, double* phase_ms, unsigned long long* phase_launches);

/* device-side known-answer probes (one thread per problem): riemann_mhd (riemann_mhd.h:354) and
 * compute_emf<dir> (riemann_mhd.h:1054; emf_dir 0 = X, 1 = Y, 2 = Z; q_edge = [n][4][8] in the
 * reference's IRT, IRB, ILT, ILB order).  Arrays are in the handle's precision. */
int rg_probe_riemann_mhd(rg_handle h, int n, const void* qleft, const void* qright, void* flux);
int rg_probe_compute_emf(rg_handle h, int n, int emf_dir, const void* q_edge, const void* x_pos, void* emf);

/* pure host helper: inner planes [k_offset, k_offset + nz_local) of slab `rank` */
int rg_slab_extent(int nz_global, int nranks, int rank, int* nz_local, int* k_offset);

#ifdef __cplusplus
}
#endif
#endif /* RAMSESGPU_B200_H_ */
