// Host-side run driver: owns device memory, streams and (optionally) the NCCL communicator, and
// sequences the kernels of one time step.  Plays the role of the reference's
// HydroRunBase / MHDRunBase / MHDRunGodunov / HydroRunGodunov objects (reference
// src/hydro/HydroRunBase.h:63-639, MHDRunGodunov.h) for the per-timestep update path.
#pragma once
#include <cstddef>
#include <memory>
#include <string>
#include <vector>

#include "config_map.h"
#include "params.h"

namespace rg {

struct Layout {
  int nx, ny, nz;           // global inner sizes
  int isize, jsize, ksize;  // LOCAL sizes with ghosts (this rank's slab)
  int nvar, ghostWidth, dim, mhd, realBytes;
  int nzLocal, kOffset;     // slab: local inner planes and global index of the first one
  int rank, nranks;
};

struct Stats {
  unsigned long long kernelLaunches;
  double lastStepMs;        // device time of the last godunov_unsplit (CUDA events)
  double haloBytesPerStep;  // bytes sent to z-neighbours per step by this rank
  size_t deviceBytes;       // device memory owned by the handle
  int chunkPlanes;          // z planes per chunk of the step pipeline
  int haloPeerCopies;       // 1: z halo by copy engines over peer-mapped memory, 0: NCCL send/recv
};

enum Phase { PH_BOUNDARY = 0, PH_PRIM, PH_TRACE, PH_FLUX, PH_EMF, PH_UPDATE, PH_DT, PH_COPY, PH_HALO, PH_FUSED, PH_DISS, PH_COUNT };

struct DistInit {
  int rank = 0, nranks = 1;
  const void* ncclUniqueId = nullptr;  // 128 bytes, required when nranks > 1
  int device = -1;                     // CUDA device ordinal; -1 = current
};

// precision-erased interface (one implementation per real type)
class Run {
 public:
  virtual ~Run() {}
  static std::unique_ptr<Run> create(const ConfigMap& cfg, bool fp32, const DistInit& dist);

  virtual Layout layout() const = 0;
  virtual const ConfigMap& config() const = 0;
  virtual const RunParams& runParams() const = 0;
  virtual double param(const std::string& name, bool* ok) const = 0;

  // operator surface (same names/meaning as the reference's virtuals)
  virtual int init_simulation(const std::string& problem) = 0;             // returns start step
  virtual void make_all_boundaries(int which) = 0;                         // 0 -> U, 1 -> U2
  virtual double compute_dt(int useU) = 0;
  virtual void godunov_unsplit(int nStep, double dt) = 0;
  virtual void oneStepIntegration(int& nStep, double& t, double& dt) = 0;
  virtual void start() = 0;                                                // full run loop
  virtual void output(int nStep) = 0;
  // history diagnostics of buffer nStep % 2 (reference MHDRunBase::history_default / history_mri), global over all
  // slabs: out[8] = mass, maxwell, reynolds, magp, mean_Bx, mean_By, mean_Bz, divB
  virtual void history(int nStep, double* out) = 0;

  // data access; host arrays are [var][k][j][i] of the LOCAL slab, ghosts included
  virtual void copyToHost(int which, void* dst, size_t bytes) = 0;
  virtual void copyFromHost(int which, const void* src, size_t bytes) = 0;
  virtual void* deviceData(int which) = 0;
  virtual void synchronize() = 0;

  // host-buffer step used by bench.py's e2e leg: H2D(U) -> nSteps steps -> D2H(result)
  virtual void stepsFromHost(const void* hostIn, void* hostOut, size_t bytes, int nSteps, double* tOut,
                             double* dtLast) = 0;

  // n independent one-step jobs from host buffers with copy/compute overlap (see rg_steps_from_host_batch)
  virtual void stepsFromHostBatch(int nJobs, const void* const* in, void* const* out, size_t bytes,
                                  double* dtOut) = 0;

  virtual Stats stats() const = 0;
  // device timing of a region, total and per kernel family (see rg_profile_begin/end)
  virtual void profileBegin() = 0;
  virtual void profileEnd(double* totalMs, double* phaseMs, unsigned long long* phaseLaunches) = 0;
  virtual void setChunkPlanes(int planes) = 0;
  virtual void setOverlap(bool on) = 0;  // early z-halo exchange overlapped with the interior update

  // device probes for known-answer tests
  virtual void probeRiemann(int n, const void* ql, const void* qr, void* flux) = 0;
  virtual void probeEmf(int n, int emfDir, const void* qEdge, const void* xPos, void* emf) = 0;

  double totalTime = 0.0;
  int stepCount = 0;
};

// z-slab decomposition helper (pure host logic; also used by the CPU tests): inner planes of rank r
void slabExtent(int nzGlobal, int nranks, int rank, int* nzLocal, int* kOffset);

}  // namespace rg
