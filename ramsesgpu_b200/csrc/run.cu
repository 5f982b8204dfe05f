#include "run.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <fstream>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <unistd.h>

#include "init_conditions.h"
#include "kernels.h"
#include "nccl_dyn.h"
#include "output.h"

namespace rg {

#define RG_CUDA(call)                                                                              \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e__) + " at " +    \
                               __FILE__ + ":" + std::to_string(__LINE__));                         \
  } while (0)

// ---- NCCL run-time binding ---------------------------------------------------------------------
const NcclApi* NcclApi::get(const char** err) {
  static NcclApi api;
  static bool tried = false, ok = false;
  static std::string msg;
  if (!tried) {
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // PyTorch's copy if mapped
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      msg = std::string("cannot load libnccl.so.2: ") + dlerror();
    } else {
#define RG_SYM(field, name)                                         \
  *(void**)(&api.field) = dlsym(h, name);                           \
  if (!api.field) msg += std::string(" missing symbol ") + name;
      RG_SYM(GetUniqueId, "ncclGetUniqueId")
      RG_SYM(CommInitRank, "ncclCommInitRank")
      RG_SYM(CommDestroy, "ncclCommDestroy")
      RG_SYM(Send, "ncclSend")
      RG_SYM(Recv, "ncclRecv")
      RG_SYM(GroupStart, "ncclGroupStart")
      RG_SYM(GroupEnd, "ncclGroupEnd")
      RG_SYM(AllReduce, "ncclAllReduce")
      RG_SYM(AllGather, "ncclAllGather")
      RG_SYM(GetErrorString, "ncclGetErrorString")
#undef RG_SYM
      ok = msg.empty();
    }
  }
  if (!ok) {
    if (err) *err = msg.c_str();
    return nullptr;
  }
  return &api;
}

int g_haloP2p = 1;  // run-time knob "halo_p2p": z halo by copy engines over peer-mapped state arrays (default) or NCCL send/recv

void slabExtent(int nzGlobal, int nranks, int rank, int* nzLocal, int* kOffset) {
  // contiguous slabs, the first (nz % nranks) ranks get one extra plane
  const int base = nzGlobal / nranks, rem = nzGlobal % nranks;
  *nzLocal = base + (rank < rem ? 1 : 0);
  *kOffset = rank * base + std::min(rank, rem);
}

namespace {

template <typename T>
class RunImpl final : public Run {
 public:
  RunImpl(const ConfigMap& cfg, const DistInit& dist) : cfg_(cfg), rank_(dist.rank), nranks_(dist.nranks) {
    rp_ = parseRunParams(cfg_);
    if (rp_.dim == 2 && nranks_ > 1) throw std::runtime_error("z-slab decomposition needs a 3D run");
    if (dist.device >= 0) RG_CUDA(cudaSetDevice(dist.device));
    int nzLocal = rp_.nz, kOff = 0;
    if (rp_.dim == 3) slabExtent(rp_.nz, nranks_, rank_, &nzLocal, &kOff);
    if (nranks_ > 1 && nzLocal < rp_.ghostWidth)
      throw std::runtime_error("z-slab thinner than the ghost width");
    kp_ = makeKParams<T>(cfg_, rp_, nzLocal, kOff);
    // variants of the reference that are not built must fail loudly, not run with silently different numerics
    for (int f = 0; f < 2 * rp_.dim; ++f)
      if (rp_.bc[f] == BC_COPY || (rp_.bc[f] == BC_Z_STRATIFIED && !(f >= 4 && rp_.mhdEnabled && rp_.dim == 3)))
        throw std::runtime_error("boundary type " + std::to_string(rp_.bc[f]) +
                                 " (copy; z-stratified outside the z faces of a 3D MHD run) is not available in this build");
    // slope_type 3 (27-point slopes, slope_mhd.h:352-409) exists for the non-rotating 3D MHD step, like in the
    // reference (its rotating CPU step never fills the slopes for that type, its 2D and hydro steps ignore it)
    if (kp_.slope_type == T(3) && !(rp_.mhdEnabled && rp_.dim == 3 && !(kp_.Omega0 > T(0))))
      throw std::runtime_error("slope_type 3 is available for the non-rotating 3D MHD solver only");
    cells_ = (size_t)kp_.isize * kp_.jsize * kp_.ksize;
    elems_ = cells_ * kp_.nvar;
    RG_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    // (measured on 2 GPUs, profiles/r02_n_halo_overlap_notes.txt: a highest-priority communication stream lets the NCCL
    // blocks take SMs from the persistent compute kernel mid-flight and spin there for the peer -- fused update 4.90 ->
    // 5.11 ms, step 6.70 -> 6.84 ms; two point-to-point channels instead of the default starve the exchange -- config-4
    // slab 3.97 -> 5.17 ms.  Default priority and channels it is.)
    RG_CUDA(cudaStreamCreateWithFlags(&comm_stream_, cudaStreamNonBlocking));
    RG_CUDA(cudaEventCreate(&ev0_));
    RG_CUDA(cudaEventCreate(&ev1_));
    RG_CUDA(cudaEventCreateWithFlags(&ev_sync_, cudaEventDisableTiming));
    RG_CUDA(cudaEventCreateWithFlags(&evEdge_, cudaEventDisableTiming));
    RG_CUDA(cudaEventCreateWithFlags(&evHalo_, cudaEventDisableTiming));
    RG_CUDA(cudaEventCreate(&evProf0_));
    RG_CUDA(cudaEventCreate(&evProf1_));
    for (int b = 0; b < 2; ++b) {
      RG_CUDA(cudaMalloc(&dU_[b], elems_ * sizeof(T)));
      RG_CUDA(cudaMemset(dU_[b], 0, elems_ * sizeof(T)));
    }
    deviceBytes_ += 2 * elems_ * sizeof(T);
    RG_CUDA(cudaMalloc(&dMax_, 2 * MAX_SLOTS * sizeof(unsigned long long)));
    RG_CUDA(cudaMemset(dMax_, 0, 2 * MAX_SLOTS * sizeof(unsigned long long)));
    RG_CUDA(cudaMallocHost(&hMax_, MAX_SLOTS * sizeof(unsigned long long)));
    // vertical gravity field of the stratified shearing box: one value per local plane
    std::vector<T> gz;
    if (stratifiedGravityPlanes<T>(cfg_, rp_, kp_, gz)) {
      RG_CUDA(cudaMalloc(&dGz_, gz.size() * sizeof(T)));
      RG_CUDA(cudaMemcpy(dGz_, gz.data(), gz.size() * sizeof(T), cudaMemcpyHostToDevice));
      kp_.gzPlane = dGz_;
    }
    std::vector<T> gc;
    if (keplerianGravityField<T>(cfg_, rp_, kp_, gc)) {
      RG_CUDA(cudaMalloc(&dGcell_, gc.size() * sizeof(T)));
      RG_CUDA(cudaMemcpy(dGcell_, gc.data(), gc.size() * sizeof(T), cudaMemcpyHostToDevice));
      kp_.gCell = dGcell_;
    }
    if (nranks_ > 1) {
      initComm(dist);
      initPeerHalo();
    }
  }

  ~RunImpl() override {
    cudaDeviceSynchronize();
    closePeerHalo();
    if (comm_ && nccl_) nccl_->CommDestroy(comm_);
    freeScratch();
    for (int b = 0; b < 2; ++b) cudaFree(dU_[b]);
    cudaFree(dDiss_);
    cudaFree(dHist_);
    cudaFree(dGz_);
    cudaFree(dGcell_);
    for (int b = 0; b < 2; ++b) {
      if (batchBuf_[b]) cudaFree(batchBuf_[b]);
      if (evH2D_[b]) cudaEventDestroy(evH2D_[b]);
      if (evD2H_[b]) cudaEventDestroy(evD2H_[b]);
      if (evStepDone_[b]) cudaEventDestroy(evStepDone_[b]);
    }
    if (h2dStream_) cudaStreamDestroy(h2dStream_);
    if (d2hStream_) cudaStreamDestroy(d2hStream_);
    cudaFree(dMax_);
    cudaFreeHost(hMax_);
    cudaEventDestroy(ev0_);
    cudaEventDestroy(ev1_);
    cudaEventDestroy(ev_sync_);
    cudaEventDestroy(evEdge_);
    cudaEventDestroy(evHalo_);
    cudaEventDestroy(evProf0_);
    cudaEventDestroy(evProf1_);
    recycleEvents();
    for (cudaEvent_t e : eventPool_) cudaEventDestroy(e);
    cudaStreamDestroy(stream_);
    cudaStreamDestroy(comm_stream_);
  }

  Layout layout() const override {
    Layout l{};
    l.nx = rp_.nx; l.ny = rp_.ny; l.nz = rp_.nz;
    l.isize = kp_.isize; l.jsize = kp_.jsize; l.ksize = kp_.ksize;
    l.nvar = kp_.nvar; l.ghostWidth = kp_.gw; l.dim = rp_.dim; l.mhd = rp_.mhdEnabled ? 1 : 0;
    l.realBytes = (int)sizeof(T);
    l.nzLocal = kp_.nz; l.kOffset = kp_.kglob0; l.rank = rank_; l.nranks = nranks_;
    return l;
  }
  const ConfigMap& config() const override { return cfg_; }
  const RunParams& runParams() const override { return rp_; }

  double param(const std::string& n, bool* ok) const override {
    if (ok) *ok = true;
    if (n == "dx") return kp_.dx;
    if (n == "dy") return kp_.dy;
    if (n == "dz") return kp_.dz;
    if (n == "xmin") return kp_.xMin;
    if (n == "xmax") return kp_.xMax;
    if (n == "ymin") return kp_.yMin;
    if (n == "ymax") return kp_.yMax;
    if (n == "zmin") return kp_.zMin;
    if (n == "zmax") return kp_.zMax;
    if (n == "gamma0") return kp_.gamma0;
    if (n == "cfl") return kp_.cfl;
    if (n == "smallr") return kp_.smallr;
    if (n == "smallc") return kp_.smallc;
    if (n == "smallp") return kp_.smallp;
    if (n == "ciso") return kp_.cIso;
    if (n == "omega0") return kp_.Omega0;
    if (n == "slope_type") return kp_.slope_type;
    if (n == "tend") return rp_.tEnd;
    if (n == "nstepmax") return rp_.nStepmax;
    if (n == "noutput") return rp_.nOutput;
    if (n == "riemannsolver") return kp_.riemannSolver;
    if (n == "magriemannsolver") return kp_.magRiemannSolver;
    if (ok) *ok = false;
    return 0.0;
  }

  // reference MHDRunBase::init_simulation (MHDRunBase.cpp:1231-1360): host IC -> device, U2 = U
  int init_simulation(const std::string& problemIn) override {
    const std::string problem = problemIn.empty() ? rp_.problem : problemIn;
    std::vector<T> h;
    std::string msg;
    int startStep = 0;
    double startTime = 0.0;
    if (rp_.restart) {  // reference MHDRunBase.cpp:1234-1282: reload a dump instead of the problem's initial condition
      // variants of the reference's restart that are not built fail loudly
      if (cfg_.getBool("run", "restart_upscale", false))
        throw std::runtime_error("restart_upscale (restart from a half-resolution dump) is not available in this build");
      std::string dir = rp_.outputDir;
      if (!dir.empty() && dir.back() != '/') dir += "/";
      // every rank resumes from ITS slab's dump (<prefix>_rankNNNN_<step>.vti, the naming of output()); when that
      // file does not exist the name is taken as a dump of the global grid, of which the rank reads its planes
      std::string path = dir + restartSlabName(rp_.restartFilename, rank_, nranks_);
      if (nranks_ > 1 && !std::ifstream(path.c_str()).good()) path = dir + rp_.restartFilename;
      h.assign(elems_, T(0));
      RestartMeta meta;
      bool ghosts = false, global = false;
      if (!readVti<T>(path, layout(), h.data(), &ghosts, &msg, &global)) throw std::runtime_error(msg);
      if (!readRestartMeta(path, &meta)) throw std::runtime_error("restart: no readable '" + path + ".meta' (time step / total time)");
      if (!global && (meta.nranks != nranks_ || meta.rank != rank_ || meta.kOffset != kp_.kglob0))
        throw std::runtime_error("restart: '" + path + "' was written by rank " + std::to_string(meta.rank) + " of " +
                                 std::to_string(meta.nranks) + " (first plane " + std::to_string(meta.kOffset) +
                                 "), not by this rank's slab");
      if (global && meta.nranks != 1) throw std::runtime_error("restart: '" + path + "' is not a dump of the global grid");
      // reference MHDRunBase.cpp:1358-1360 / MHDRunGodunov.cpp:3875-3877
      startStep = cfg_.getBool("run", "restart_reset_timestep", false) ? 0 : meta.nStep;
      startTime = cfg_.getBool("run", "restart_reset_totaltime", false) ? 0.0 : meta.totalTime;
      lastDt_ = meta.dt;
      // next dt of the uninterrupted run: valid only when the resumed run sees the same decomposition-independent state
      resumeDt_ = meta.dtNext;  // used for the first step after the restart (cleared by godunov_unsplit)
    } else if (!initProblem<T>(cfg_, rp_, kp_, problem, h, &msg)) {
      std::fprintf(stderr, "ramsesgpu_b200: %s\n", msg.c_str());
      lastWarning_ = msg;
    }
    RG_CUDA(cudaMemcpyAsync(dU_[0], h.data(), elems_ * sizeof(T), cudaMemcpyHostToDevice, stream_));
    RG_CUDA(cudaMemcpyAsync(dU_[1], dU_[0], elems_ * sizeof(T), cudaMemcpyDeviceToDevice, stream_));
    RG_CUDA(cudaStreamSynchronize(stream_));
    invalidate(0);
    invalidate(1);
    totalTime = startTime;
    stepCount = startStep;
    return startStep;
  }

  // reference HydroRunBase::make_all_boundaries (HydroRunBase.cpp:2322-2342): X, Y, then Z
  void make_all_boundaries(int which) override {
    // reference start(): make_all_boundaries_shear(U, 0, 0) when the shearing box is enabled
    if (shearingBox()) fillGhostsShear(which, T(0));
    else fillGhosts(which, 0, kp_.ksize);
    ghostsValid_[which] = true;
  }

  // reference MHDRunBase::compute_dt_mhd / HydroRunBase::compute_dt: cfl / max inverse dt
  double compute_dt(int useU) override {
    if (resumeDt_ > 0.0) return resumeDt_;
    const int b = useU ? 1 : 0;
    unsigned long long* slots = dMax_ + (size_t)b * MAX_SLOTS;
    if (!dtCached_[b]) {
      RG_CUDA(cudaMemsetAsync(slots, 0, MAX_SLOTS * sizeof(unsigned long long), stream_));
      phase(PH_DT, [&] {
        if (rp_.mhdEnabled) MhdKernels<T>::computeInvDt(kp_, dU_[b], slots, stream_);
        else if (rp_.dim == 3) HydroKernels<T>::computeInvDt(kp_, dU_[b], slots, stream_);
        else Hydro2dKernels<T>::computeInvDt(kp_, dU_[b], slots, stream_);
      });
      dtCached_[b] = true;
    }
    if (nranks_ > 1) {  // slots hold bit patterns of non-negative doubles: a floating max is exact
      // an NCCL halo in flight on the communication stream shares the communicator: keep the order.  (A copy-engine halo
      // does not: it keeps travelling behind this reduction and the host turn-around; the next ghost fill waits for it.)
      if (haloOnNccl_ && (haloDone_[0] || haloDone_[1])) RG_CUDA(cudaStreamWaitEvent(stream_, evHalo_, 0));
      ncclCheck(nccl_->AllReduce(slots, slots, MAX_SLOTS, NcclApi::kFloat64, NcclApi::kMax, comm_, stream_),
                "allreduce(dt)");
    }
    RG_CUDA(cudaMemcpyAsync(hMax_, slots, MAX_SLOTS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream_));
    RG_CUDA(cudaStreamSynchronize(stream_));
    unsigned long long best = 0;
    for (int i = 0; i < MAX_SLOTS; ++i) best = std::max(best, hMax_[i]);
    // seed of the running max, reference MHDRunBase.cpp:144
    // (the hydro driver starts from 0, HydroRunBase.cpp:386)
    T invDt = rp_.mhdEnabled ? kp_.smallc / std::min(kp_.dx, kp_.dy) : T(0);
    invDt = std::max(invDt, static_cast<T>(decodeMax(best)));
    // the inflow speed of the jet limits the step too (reference HydroRunBase.cpp:420-422, MHDRunBase.cpp:228-230)
    if (kp_.jet) invDt = std::max(invDt, (kp_.ujet + kp_.cjet) / kp_.dx);
    return static_cast<double>(kp_.cfl / invDt);
  }

  // reference MHDRunGodunov::godunov_unsplit (MHDRunGodunov.cpp:572-594): even step U -> U2
  void godunov_unsplit(int nStep, double dt) override {
    const int src = (nStep % 2 == 0) ? 0 : 1, dst = 1 - src;
    resumeDt_ = 0.0;
    RG_CUDA(cudaEventRecord(ev0_, stream_));
    const bool rotating = rp_.mhdEnabled && kp_.Omega0 > T(0);
    // the rotating-frame step fills the ghosts of UNew at its END (reference MHDRunGodunov.cpp:3429-3437)
    if (!rotating && !ghostsValid_[src]) make_all_boundaries(src);
    if (kp_.gravity && rp_.dim != 3 && rp_.mhdEnabled)
      throw std::runtime_error("static gravity is available for the 3D solvers and the 2D hydro solver only");
    if ((kp_.nu > T(0) || kp_.eta > T(0)) && rp_.dim != 3)
      throw std::runtime_error("viscosity / resistivity are available for the 3D solvers only");
    if (rotating && rp_.dim != 3)  // (the reference has a 2D rotating step, MHDRunGodunov.cpp:2089-2435: not built)
      throw std::runtime_error("the rotating frame ([MHD] omega0 > 0) is available for the 3D MHD solver only");
    if (rotating && rp_.dim == 3) {
      stepMhd3dRotating(src, dst, static_cast<T>(dt));
    } else if (rp_.mhdEnabled && rp_.dim == 3) {
      stepMhd3d(src, dst, static_cast<T>(dt));
    } else if (!rp_.mhdEnabled && rp_.dim == 3) {
      stepHydro3d(src, dst, static_cast<T>(dt));
    } else if (rp_.mhdEnabled && rp_.dim == 2) {
      stepMhd2d(src, dst, static_cast<T>(dt));
    } else {
      stepHydro2d(src, dst, static_cast<T>(dt));
    }
    RG_CUDA(cudaEventRecord(ev1_, stream_));
    timed_ = true;
  }

  // reference MHDRunGodunov::oneStepIntegration (MHDRunGodunov.cpp:4077-4089)
  void oneStepIntegration(int& nStep, double& t, double& dt) override {
    dt = compute_dt(nStep % 2);
    totalTime = t;  // the shearing-box remap depends on the time at the start of the step
    godunov_unsplit(nStep, dt);
    nStep++;
    // the reference accumulates time in real_t
    t = static_cast<double>(static_cast<T>(t) + static_cast<T>(dt));
    totalTime = t;
    stepCount = nStep;
    lastDt_ = dt;
  }

  // reference MHDRunGodunov::start (MHDRunGodunov.cpp:3801-4070), outputs every nOutput steps
  void start() override {
    int nStep = init_simulation("");
    make_all_boundaries(0);
    RG_CUDA(cudaMemcpyAsync(dU_[1], dU_[0], elems_ * sizeof(T), cudaMemcpyDeviceToDevice, stream_));
    ghostsValid_[1] = true;
    dtCached_[1] = false;
    double t = totalTime, dt = compute_dt(nStep % 2);  // t > 0, nStep > 0 only when restarting from a dump
    if (rank_ == 0) std::printf("Initial dt : %.12g\n", dt);
    const int firstStep = nStep;
    // history cadence of the reference, MHDRunGodunov.cpp:3915-3916 / :3975-3983 (arithmetic in real_t)
    const T dtHist = cfg_.getFloat("history", "dtHist", static_cast<float>(10 * dt));
    T tHist = static_cast<T>(t);
    const auto t0 = std::chrono::steady_clock::now();
    double ioSeconds = 0.0;
    while (t < rp_.tEnd && nStep < rp_.nStepmax) {
      if (rp_.nOutput > 0 && (nStep % rp_.nOutput) == 0 && !(rp_.restart && nStep == firstStep)) {
        const auto a = std::chrono::steady_clock::now();
        stepCount = nStep; totalTime = t; lastDt_ = dt;
        output(nStep);
        ioSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count();
        if (rank_ == 0) std::printf("step=%9d t=%.10g dt=%.12g\n", nStep, t, dt);
      }
      if (rp_.historyEnabled) {
        const T tt = static_cast<T>(t), dd = static_cast<T>(dt);
        if (tHist == T(0) || ((tt - dd <= tHist + dtHist) && (tt > tHist + dtHist))) {
          writeHistoryLine(nStep, t, dt);
          tHist += dtHist;
        }
      }
      oneStepIntegration(nStep, t, dt);
    }
    synchronize();
    stepCount = nStep; totalTime = t;
    {
      const auto a = std::chrono::steady_clock::now();
      output(nStep);
      ioSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count();
    }
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (rank_ == 0) {
      // the reference's metric line, MHDRunGodunov.cpp:4064-4068
      std::printf("total time: %5.3f sec, output time: %5.3f sec\n", wall, ioSeconds);
      std::printf("####################################\nGlobal performance                  \n%g cell updates per seconds (based on wall time)\n####################################\n",
                  1.0 * nStep * rp_.nx * rp_.ny * rp_.nz / (wall - ioSeconds));
    }
  }

  void output(int nStep) override {
    std::vector<T> h(elems_);
    copyToHost(nStep % 2, h.data(), elems_ * sizeof(T));
    writeOutputs<T>(rp_, layout(), h.data(), nStep);
    if (rp_.outputVtk && !rp_.outputVtkAscii)  // what a restart needs besides the fields (see output.h)
      writeRestartMeta(vtiPath(rp_, layout(), nStep),
                       RestartMeta{nStep, totalTime, lastDt_, compute_dt(nStep % 2), rank_, nranks_, kp_.kglob0});
  }

  // reference MHDRunBase::history_default / history_mri (MHDRunBase.cpp:3311-3410, :3476-3620), reduced on the
  // device (kernels_history.cu); sums over slabs with ncclAllReduce
  void history(int nStep, double* out) override {
    if (!(rp_.mhdEnabled && rp_.dim == 3)) throw std::runtime_error("history diagnostics are available for 3D MHD only");
    const T* U = dU_[nStep % 2];
    const int nB = 296, is = kp_.isize;
    const size_t n1 = (size_t)nB * 3 * is, nMean = (size_t)3 * is, n2 = (size_t)nB * 8;
    if (!dHist_) {
      RG_CUDA(cudaMalloc(&dHist_, (n1 + nMean + n2 + 8) * sizeof(double)));
      deviceBytes_ += (n1 + nMean + n2 + 8) * sizeof(double);
      hHist_.resize(n1 + nMean + n2 + 8);
    }
    double *dPart1 = dHist_, *dMean = dHist_ + n1, *dPart2 = dMean + nMean, *dTot = dPart2 + n2;
    auto allSum = [&](double* dBuf, double* hBuf, size_t n) {  // sum over ranks of a small host array
      if (nranks_ == 1) return;
      RG_CUDA(cudaMemcpyAsync(dBuf, hBuf, n * sizeof(double), cudaMemcpyHostToDevice, stream_));
      if (haloDone_[0] || haloDone_[1]) RG_CUDA(cudaStreamWaitEvent(stream_, evHalo_, 0));
      ncclCheck(nccl_->AllReduce(dBuf, dBuf, n, NcclApi::kFloat64, NcclApi::kSum, comm_, stream_), "allreduce(history)");
      RG_CUDA(cudaMemcpyAsync(hBuf, dBuf, n * sizeof(double), cudaMemcpyDeviceToHost, stream_));
      RG_CUDA(cudaStreamSynchronize(stream_));
    };
    // an overlapped step may still be filling the ghosts of this buffer on the communication stream
    if (haloDone_[nStep % 2]) RG_CUDA(cudaStreamWaitEvent(stream_, evHalo_, 0));
    HistoryKernels<T>::columnSums(kp_, U, dPart1, nB, stream_);
    RG_CUDA(cudaMemcpyAsync(hHist_.data(), dPart1, n1 * sizeof(double), cudaMemcpyDeviceToHost, stream_));
    RG_CUDA(cudaStreamSynchronize(stream_));
    std::vector<double> col(nMean, 0.0);
    for (int b = 0; b < nB; ++b)
      for (size_t q = 0; q < nMean; ++q) col[q] += hHist_[(size_t)b * nMean + q];
    allSum(dMean, col.data(), nMean);
    const double cells = (double)rp_.ny * (double)rp_.nz;  // GLOBAL y-z plane
    for (int i = 0; i < is; ++i) { col[i] = col[is + i] / cells; col[is + i] = col[2 * is + i] / cells; }  // mean u, mean v
    RG_CUDA(cudaMemcpyAsync(dMean, col.data(), 2 * is * sizeof(double), cudaMemcpyHostToDevice, stream_));
    HistoryKernels<T>::sums(kp_, U, dMean, dPart2, nB, stream_);
    RG_CUDA(cudaMemcpyAsync(hHist_.data(), dPart2, n2 * sizeof(double), cudaMemcpyDeviceToHost, stream_));
    RG_CUDA(cudaStreamSynchronize(stream_));
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = 0; b < nB; ++b)
      for (int q = 0; q < 8; ++q) s[q] += hHist_[(size_t)b * 8 + q];
    allSum(dTot, s, 8);
    const double dTau = (double)kp_.dx * (double)kp_.dy * (double)kp_.dz / (double)(kp_.xMax - kp_.xMin) /
                        (double)(kp_.yMax - kp_.yMin) / (double)(kp_.zMax - kp_.zMin);
    out[0] = s[0] * dTau; out[1] = s[1] * dTau; out[2] = s[2] * dTau; out[3] = s[3] * dTau / 2.;
    out[4] = s[4] * dTau; out[5] = s[5] * dTau; out[6] = s[6] * dTau; out[7] = s[7];
  }

  // one line of the reference's history file (<outputDir>/<outputPrefix>_<[history] filename>), rank 0 only
  void writeHistoryLine(int nStep, double t, double dt) {
    const bool mri = rp_.problem == "MRI" || rp_.problem == "Mri" || rp_.problem == "mri";
    const bool ot = rp_.problem == "Orszag-Tang" || rp_.problem == "OrszagTang";
    if (!(mri || ot) || !(rp_.mhdEnabled && rp_.dim == 3)) return;  // the reference's history_empty
    double h[8];
    history(nStep, h);
    if (rank_ != 0) return;
    std::string path = rp_.outputDir;
    if (!path.empty() && path.back() != '/') path += "/";
    path += rp_.outputPrefix + "_" + rp_.historyFilename;
    std::ofstream histo(path.c_str(), std::ios::out | std::ios::app | std::ios::ate);
    if (t <= 0) {
      histo << "# history\n";
      if (rp_.restart) histo << "# history : this is a restart run\n";
      histo << (mri ? "# totalTime dt mass maxwell reynolds maxwell+reynolds magp mean_Bx mean_By mean_Bz divB\n"
                    : "# totalTime dt mass divB\n");
    }
    histo << t << "\t" << dt << "\t" << h[0] << "\t";
    if (mri)
      histo << h[1] << "\t" << h[2] << "\t" << h[1] + h[2] << "\t" << h[3] << "\t" << h[4] << "\t" << h[5] << "\t"
            << h[6] << "\t";
    histo << h[7] << "\n";
  }

  void copyToHost(int which, void* dst, size_t bytes) override {
    checkBytes(bytes);
    if (haloDone_[which ? 1 : 0]) RG_CUDA(cudaStreamWaitEvent(stream_, evHalo_, 0));
    RG_CUDA(cudaMemcpyAsync(dst, dU_[which ? 1 : 0], bytes, cudaMemcpyDeviceToHost, stream_));
    RG_CUDA(cudaStreamSynchronize(stream_));
  }
  void copyFromHost(int which, const void* src, size_t bytes) override {
    checkBytes(bytes);
    RG_CUDA(cudaMemcpyAsync(dU_[which ? 1 : 0], src, bytes, cudaMemcpyHostToDevice, stream_));
    RG_CUDA(cudaStreamSynchronize(stream_));
    invalidate(which ? 1 : 0);
  }
  void* deviceData(int which) override { return dU_[which ? 1 : 0]; }
  void synchronize() override {
    RG_CUDA(cudaStreamSynchronize(stream_));
    RG_CUDA(cudaStreamSynchronize(comm_stream_));
  }

  void stepsFromHost(const void* hostIn, void* hostOut, size_t bytes, int nSteps, double* tOut, double* dtLast) override {
    checkBytes(bytes);
    RG_CUDA(cudaMemcpyAsync(dU_[0], hostIn, bytes, cudaMemcpyHostToDevice, stream_));
    invalidate(0);
    invalidate(1);
    int nStep = 0;
    double t = 0.0, dt = 0.0;
    for (int s = 0; s < nSteps; ++s) oneStepIntegration(nStep, t, dt);
    RG_CUDA(cudaMemcpyAsync(hostOut, dU_[nStep % 2], bytes, cudaMemcpyDeviceToHost, stream_));
    RG_CUDA(cudaStreamSynchronize(stream_));
    if (tOut) *tOut = t;
    if (dtLast) *dtLast = dt;
  }

  // n independent one-step jobs out[j] = step(in[j]) from HOST buffers.  The copy engines and the SMs
  // work concurrently: H2D of job j+1 (copy stream), the step of job j (compute stream) and D2H of
  // job j-1 (second copy stream) overlap, with two device buffer pairs in rotation (PCIe is full
  // duplex, so the steady state costs max(H2D, D2H, step) per job instead of their sum).  Results
  // are bitwise those of n calls of stepsFromHost(in[j], out[j], bytes, 1).
  void stepsFromHostBatch(int nJobs, const void* const* in, void* const* out, size_t bytes, double* dtOut) override {
    checkBytes(bytes);
    if (nJobs <= 0) return;
    ensureBatchResources();
    T* saved[2] = {dU_[0], dU_[1]};
    T* pairIn[2] = {saved[0], batchBuf_[0]};
    T* pairOut[2] = {saved[1], batchBuf_[1]};
    RG_CUDA(cudaStreamSynchronize(stream_));
    auto enqueueH2D = [&](int j) {
      const int s = j & 1;
      if (j >= 2) RG_CUDA(cudaStreamWaitEvent(h2dStream_, evStepDone_[s], 0));  // job j-2 no longer reads pairIn[s]
      RG_CUDA(cudaMemcpyAsync(pairIn[s], in[j], bytes, cudaMemcpyHostToDevice, h2dStream_));
      RG_CUDA(cudaEventRecord(evH2D_[s], h2dStream_));
    };
    try {
      enqueueH2D(0);
      for (int j = 0; j < nJobs; ++j) {
        const int s = j & 1;
        if (j + 1 < nJobs) enqueueH2D(j + 1);  // before the host blocks on this job's dt read-back
        RG_CUDA(cudaStreamWaitEvent(stream_, evH2D_[s], 0));
        if (j >= 2) RG_CUDA(cudaStreamWaitEvent(stream_, evD2H_[s], 0));  // job j-2 has left pairOut[s]
        dU_[0] = pairIn[s];
        dU_[1] = pairOut[s];
        invalidate(0);
        invalidate(1);
        int nStep = 0;
        double t = 0.0, dt = 0.0;
        oneStepIntegration(nStep, t, dt);
        if (haloDone_[1]) RG_CUDA(cudaStreamWaitEvent(stream_, evHalo_, 0));
        RG_CUDA(cudaEventRecord(evStepDone_[s], stream_));
        RG_CUDA(cudaStreamWaitEvent(d2hStream_, evStepDone_[s], 0));
        RG_CUDA(cudaMemcpyAsync(out[j], pairOut[s], bytes, cudaMemcpyDeviceToHost, d2hStream_));
        RG_CUDA(cudaEventRecord(evD2H_[s], d2hStream_));
        if (dtOut) dtOut[j] = dt;
      }
      RG_CUDA(cudaStreamSynchronize(d2hStream_));
      RG_CUDA(cudaStreamSynchronize(stream_));
    } catch (...) {
      cudaDeviceSynchronize();
      dU_[0] = saved[0];
      dU_[1] = saved[1];
      invalidate(0);
      invalidate(1);
      throw;
    }
    dU_[0] = saved[0];
    dU_[1] = saved[1];
    invalidate(0);
    invalidate(1);
  }

  void ensureBatchResources() {
    if (batchBuf_[0]) return;
    for (int b = 0; b < 2; ++b) {
      RG_CUDA(cudaMalloc(&batchBuf_[b], elems_ * sizeof(T)));
      RG_CUDA(cudaEventCreateWithFlags(&evH2D_[b], cudaEventDisableTiming));
      RG_CUDA(cudaEventCreateWithFlags(&evD2H_[b], cudaEventDisableTiming));
      RG_CUDA(cudaEventCreateWithFlags(&evStepDone_[b], cudaEventDisableTiming));
    }
    RG_CUDA(cudaStreamCreateWithFlags(&h2dStream_, cudaStreamNonBlocking));
    RG_CUDA(cudaStreamCreateWithFlags(&d2hStream_, cudaStreamNonBlocking));
    deviceBytes_ += 2 * elems_ * sizeof(T);
  }

  Stats stats() const override {
    Stats s{};
    s.kernelLaunches = kernelLaunchCount();
    s.lastStepMs = 0.0;
    if (timed_) {
      float ms = 0.f;
      if (cudaEventSynchronize(ev1_) == cudaSuccess && cudaEventElapsedTime(&ms, ev0_, ev1_) == cudaSuccess)
        s.lastStepMs = ms;
    }
    s.haloBytesPerStep = haloBytesPerStep_;
    s.haloPeerCopies = p2p_ ? 1 : 0;
    s.deviceBytes = deviceBytes_;
    s.chunkPlanes = chunkPlanes_;
    return s;
  }

  void setOverlap(bool on) override { overlapHalo_ = on; }
  void setChunkPlanes(int planes) override {
    userChunk_ = planes;
    freeScratch();
  }

  void probeRiemann(int n, const void* ql, const void* qr, void* flux) override {
    T *dl, *dr, *df;
    const size_t nv = rp_.mhdEnabled ? 8 : 5;  // state size: MHD 8, hydro 5
    RG_CUDA(cudaMalloc(&dl, n * nv * sizeof(T)));
    RG_CUDA(cudaMalloc(&dr, n * nv * sizeof(T)));
    RG_CUDA(cudaMalloc(&df, n * nv * sizeof(T)));
    RG_CUDA(cudaMemcpy(dl, ql, n * nv * sizeof(T), cudaMemcpyHostToDevice));
    RG_CUDA(cudaMemcpy(dr, qr, n * nv * sizeof(T), cudaMemcpyHostToDevice));
    if (rp_.mhdEnabled) MhdKernels<T>::probeRiemann(kp_, n, dl, dr, df, stream_);
    else HydroKernels<T>::probeRiemann(kp_, n, dl, dr, df, stream_);
    RG_CUDA(cudaStreamSynchronize(stream_));
    RG_CUDA(cudaMemcpy(flux, df, n * nv * sizeof(T), cudaMemcpyDeviceToHost));
    cudaFree(dl); cudaFree(dr); cudaFree(df);
  }
  void probeEmf(int n, int emfDir, const void* qEdge, const void* xPos, void* emf) override {
    T *dq, *dx = nullptr, *de;
    RG_CUDA(cudaMalloc(&dq, n * 32 * sizeof(T)));
    RG_CUDA(cudaMalloc(&de, n * sizeof(T)));
    RG_CUDA(cudaMemcpy(dq, qEdge, n * 32 * sizeof(T), cudaMemcpyHostToDevice));
    if (xPos) {
      RG_CUDA(cudaMalloc(&dx, n * sizeof(T)));
      RG_CUDA(cudaMemcpy(dx, xPos, n * sizeof(T), cudaMemcpyHostToDevice));
    }
    MhdKernels<T>::probeEmf(kp_, n, emfDir, dq, dx, de, stream_);
    RG_CUDA(cudaStreamSynchronize(stream_));
    RG_CUDA(cudaMemcpy(emf, de, n * sizeof(T), cudaMemcpyDeviceToHost));
    cudaFree(dq); cudaFree(de);
    if (dx) cudaFree(dx);
  }

 public:
  void profileBegin() override {
    recycleEvents();
    profiling_ = true;
    RG_CUDA(cudaEventRecord(evProf0_, stream_));
  }
  void profileEnd(double* totalMs, double* phaseMs, unsigned long long* phaseLaunches) override {
    RG_CUDA(cudaEventRecord(evProf1_, stream_));
    RG_CUDA(cudaStreamSynchronize(stream_));
    profiling_ = false;
    float ms = 0.f;
    RG_CUDA(cudaEventElapsedTime(&ms, evProf0_, evProf1_));
    if (totalMs) *totalMs = ms;
    double acc[PH_COUNT] = {0};
    unsigned long long cnt[PH_COUNT] = {0};
    for (const Span& sp : spans_) {
      RG_CUDA(cudaEventElapsedTime(&ms, sp.a, sp.b));
      acc[sp.phase] += ms;
      cnt[sp.phase] += sp.launches;
    }
    for (int p = 0; p < PH_COUNT; ++p) {
      if (phaseMs) phaseMs[p] = acc[p];
      if (phaseLaunches) phaseLaunches[p] = cnt[p];
    }
    recycleEvents();
  }

 private:
  struct Span { int phase; cudaEvent_t a, b; unsigned long long launches; };
  cudaEvent_t newEvent() {
    if (!eventPool_.empty()) { cudaEvent_t e = eventPool_.back(); eventPool_.pop_back(); return e; }
    cudaEvent_t e;
    RG_CUDA(cudaEventCreate(&e));
    return e;
  }
  void recycleEvents() {
    for (const Span& sp : spans_) { eventPool_.push_back(sp.a); eventPool_.push_back(sp.b); }
    spans_.clear();
  }
  // runs `fn` (kernel launches on stream_) and, when profiling, brackets it with events
  template <typename F>
  void phase(int ph, F&& fn) {
    if (!profiling_) { fn(); return; }
    Span sp{ph, newEvent(), newEvent(), kernelLaunchCount()};
    RG_CUDA(cudaEventRecord(sp.a, stream_));
    fn();
    RG_CUDA(cudaEventRecord(sp.b, stream_));
    sp.launches = kernelLaunchCount() - sp.launches;
    spans_.push_back(sp);
  }

  void checkBytes(size_t bytes) const {
    if (bytes != elems_ * sizeof(T)) throw std::runtime_error("host buffer size does not match the state array");
  }
  void invalidate(int b) {
    if (haloDone_[b]) { cudaStreamSynchronize(comm_stream_); haloDone_[b] = false; }
    ghostsValid_[b] = false;
    dtCached_[b] = false;
  }

  void ncclCheck(int rc, const char* what) {
    if (rc != 0) throw std::runtime_error(std::string("NCCL error in ") + what + ": " + nccl_->GetErrorString(rc));
  }

  void initComm(const DistInit& dist) {
    const char* err = nullptr;
    nccl_ = NcclApi::get(&err);
    if (!nccl_) throw std::runtime_error(err ? err : "NCCL unavailable");
    if (!dist.ncclUniqueId) throw std::runtime_error("nranks > 1 needs an NCCL unique id");
    NcclApi::UniqueId id;
    std::memcpy(&id, dist.ncclUniqueId, sizeof id);
    ncclCheck(nccl_->CommInitRank(&comm_, nranks_, id, rank_), "ncclCommInitRank");
  }

  // ---- ghost cells -------------------------------------------------------------------------------
  // x and y faces are always local; z faces are local for a single slab, otherwise the gw planes
  // next to each slab interface travel over NCCL (reference: copy_boundaries + transfert_boundaries
  // + make_boundary, HydroRunBaseMpi.cpp:3294-3389, without the host staging).
  void zNeighbours(bool* hasLo, bool* hasHi) const {
    const bool periodic = rp_.bc[4] == BC_PERIODIC && rp_.bc[5] == BC_PERIODIC;
    *hasLo = rank_ > 0 || periodic;
    *hasHi = rank_ < nranks_ - 1 || periodic;
  }

  // physical z faces this rank owns: Dirichlet / Neumann / periodic copies, or the hydrostatic extrapolation of the
  // stratified shearing box (BC_Z_STRATIFIED)
  void fillZFaces(T* U, bool skipLo, bool skipHi) {
    MhdKernels<T>::fillBoundary(kp_, U, 2, rp_.bc[4], rp_.bc[5], skipLo, skipHi, 0, kp_.ksize, stream_);
    const bool lo = rp_.bc[4] == BC_Z_STRATIFIED && !skipLo, hi = rp_.bc[5] == BC_Z_STRATIFIED && !skipHi;
    if (lo || hi) MhdKernels<T>::fillBoundaryZStratified(kp_, U, lo, hi, cfg_.getBool("MRI", "floor", false), stream_);
  }

  void fillGhosts(int b, int kLo, int kHi) {
    T* U = dU_[b];
    if (haloDone_[b]) {  // the z halo of this buffer was exchanged on the communication stream
      RG_CUDA(cudaStreamWaitEvent(stream_, evHalo_, 0));
    }
    phase(PH_BOUNDARY, [&] {
      MhdKernels<T>::fillBoundary(kp_, U, 0, rp_.bc[0], rp_.bc[1], false, false, kLo, kHi, stream_);
      MhdKernels<T>::fillBoundary(kp_, U, 1, rp_.bc[2], rp_.bc[3], false, false, kLo, kHi, stream_);
      if (rp_.dim == 3 && nranks_ == 1)
        fillZFaces(U, false, false);
      // jet inflow patch after the last direction (reference HydroRunBase.cpp:2290, :2310); with slabs below
      if (kp_.jet && (rp_.dim == 2 || nranks_ == 1)) MhdKernels<T>::jetInflow(kp_, U, stream_);
    });
    if (rp_.dim == 2 || nranks_ == 1) return;
    bool hasLo, hasHi;
    zNeighbours(&hasLo, &hasHi);
    // physical z faces of the outermost slabs
    if (!hasLo || !hasHi)
      phase(PH_BOUNDARY, [&] {
        fillZFaces(U, hasLo, hasHi);
        if (kp_.jet && !hasLo) MhdKernels<T>::jetInflow(kp_, U, stream_);  // the slab that owns the lower z face
      });
    if (haloDone_[b]) {
      haloDone_[b] = false;
      return;
    }
    phase(PH_HALO, [&] { exchangeZ(U, hasLo, hasHi, stream_); });
  }

  // Copy-engine halo (p2p_): the transfer needs no SM, so the step makes ONE pass over the slab like on one GPU and the
  // planes leave right after it on the communication stream, as they are: the x/y fills act plane by plane, so the
  // receiver's ghost fill (x, y over all planes, ghost planes included, after the wait) gives them the x/y ghosts the
  // sender would have.  The transfer hides behind the dt reduction and the host turn-around between two steps.
  void startLateHalo(int b) {
    bool hasLo, hasHi;
    zNeighbours(&hasLo, &hasHi);
    RG_CUDA(cudaEventRecord(evEdge_, stream_));
    RG_CUDA(cudaStreamWaitEvent(comm_stream_, evEdge_, 0));
    exchangeZ(dU_[b], hasLo, hasHi, comm_stream_);
    RG_CUDA(cudaEventRecord(evHalo_, comm_stream_));
    haloDone_[b] = true;
  }

  // NCCL halo (its send/recv blocks need SMs): early halo of the buffer being written by the current step -- as soon as
  // the gw inner planes next to each slab interface are final, fill their x/y ghosts and exchange them on the
  // communication stream while the interior chunks are still being computed on the main stream.
  void startEarlyHalo(int b) {
    T* U = dU_[b];
    const int gw = kp_.gw, kN = kp_.ksize - gw;
    bool hasLo, hasHi;
    zNeighbours(&hasLo, &hasHi);
    RG_CUDA(cudaEventRecord(evEdge_, stream_));
    RG_CUDA(cudaStreamWaitEvent(comm_stream_, evEdge_, 0));
    MhdKernels<T>::fillBoundary(kp_, U, 0, rp_.bc[0], rp_.bc[1], false, false, gw, 2 * gw, comm_stream_);
    MhdKernels<T>::fillBoundary(kp_, U, 1, rp_.bc[2], rp_.bc[3], false, false, gw, 2 * gw, comm_stream_);
    MhdKernels<T>::fillBoundary(kp_, U, 0, rp_.bc[0], rp_.bc[1], false, false, kN - gw, kN, comm_stream_);
    MhdKernels<T>::fillBoundary(kp_, U, 1, rp_.bc[2], rp_.bc[3], false, false, kN - gw, kN, comm_stream_);
    exchangeZ(U, hasLo, hasHi, comm_stream_);
    RG_CUDA(cudaEventRecord(evHalo_, comm_stream_));
    haloDone_[b] = true;
  }


  // ---- z halo by copy engines over peer-mapped memory ----------------------------------------------
  // Ranks of one node map each other's state arrays (CUDA IPC) and PUT their boundary planes straight into the
  // neighbour's ghost planes with cudaMemcpyAsync (copy engines over NVLink: no SM, so the transfer runs while the
  // persistent compute blocks own every SM -- NCCL's send/recv blocks wait for one to retire, profiles/r02_n_halo_overlap_notes.txt).
  // Cross-process ordering without the host: 64-bit sequence flags in device memory, written into the neighbour's flag
  // array by an 8-byte copy behind the data and awaited with a stream memory operation (cuStreamWaitValue64):
  //   READY(s): "my ghost planes may be overwritten by exchange s" (everything that read them is earlier in my stream)
  //   DATA(s) : "the planes of exchange s are in your ghost planes"
  // Every rank calls the exchanges in the same order, so the exchange counter s agrees.  Falls back to NCCL send/recv when
  // a rank is on another host, IPC or peer access is unavailable, or for arrays that are not the two state arrays.
  struct PeerBlock {  // what a rank publishes through the all-gather
    cudaIpcMemHandle_t state[2], flags;
    unsigned long long host;
    long long pid;
    int ksize, want, pad[10];
  };
  static_assert(sizeof(PeerBlock) == 256, "peer block layout");
  struct Peer {
    T* U[2] = {nullptr, nullptr};
    unsigned long long* flags = nullptr;
    int ksize = 0;
  };
  enum { FL_DATA_FROM_LO = 0, FL_DATA_FROM_HI = 1, FL_READY_FROM_LO = 2, FL_READY_FROM_HI = 3, FL_SLOTS = 8, FL_NSLOT = 48 };
  typedef int (*StreamMemOp)(cudaStream_t, unsigned long long, unsigned long long, unsigned int);

  static unsigned long long hostId() {
    char name[256] = {0};
    gethostname(name, sizeof name - 1);
    unsigned long long h = 1469598103934665603ull;
    for (const char* c = name; *c; ++c) h = (h ^ (unsigned char)*c) * 1099511628211ull;
    // the boot id tells two containers with the same host name apart
    if (FILE* f = std::fopen("/proc/sys/kernel/random/boot_id", "r")) {
      char b[64] = {0};
      if (std::fgets(b, sizeof b, f))
        for (const char* c = b; *c; ++c) h = (h ^ (unsigned char)*c) * 1099511628211ull;
      std::fclose(f);
    }
    return h;
  }

  void initPeerHalo() {
    int want = g_haloP2p ? 1 : 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue64", &fn, cudaEnableDefault, &q) == cudaSuccess && fn) waitValue_ = (StreamMemOp)fn;
    fn = nullptr;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue64", &fn, cudaEnableDefault, &q) == cudaSuccess && fn) writeValue_ = (StreamMemOp)fn;
    if (!waitValue_ || !writeValue_) want = 0;
    cudaGetLastError();
    RG_CUDA(cudaMalloc(&dFlags_, (FL_SLOTS + FL_NSLOT) * sizeof(unsigned long long)));
    RG_CUDA(cudaMemset(dFlags_, 0, (FL_SLOTS + FL_NSLOT) * sizeof(unsigned long long)));
    if (want) {  // stream memory operations usable on this device / driver?
      const unsigned long long a = (unsigned long long)(uintptr_t)(dFlags_ + FL_SLOTS);
      if (writeValue_(stream_, a, 1ull, 0u) != 0 || waitValue_(stream_, a, 1ull, 0u) != 0 || cudaStreamSynchronize(stream_) != cudaSuccess) {
        cudaGetLastError();
        want = 0;
      }
    }
    PeerBlock mine;
    std::memset(&mine, 0, sizeof mine);
    if (cudaIpcGetMemHandle(&mine.state[0], dU_[0]) != cudaSuccess || cudaIpcGetMemHandle(&mine.state[1], dU_[1]) != cudaSuccess ||
        cudaIpcGetMemHandle(&mine.flags, dFlags_) != cudaSuccess) {
      cudaGetLastError();
      want = 0;
    }
    mine.host = hostId();
    mine.pid = (long long)getpid();
    mine.ksize = kp_.ksize;
    mine.want = want;
    // every rank learns every block (in-place all-gather), then opens its neighbours'
    std::vector<PeerBlock> all(nranks_);
    PeerBlock* dBlocks = nullptr;
    RG_CUDA(cudaMalloc(&dBlocks, sizeof(PeerBlock) * nranks_));
    RG_CUDA(cudaMemcpyAsync(dBlocks + rank_, &mine, sizeof mine, cudaMemcpyHostToDevice, stream_));
    ncclCheck(nccl_->AllGather(dBlocks + rank_, dBlocks, sizeof(PeerBlock), NcclApi::kInt8, comm_, stream_), "allgather(peer handles)");
    RG_CUDA(cudaMemcpyAsync(all.data(), dBlocks, sizeof(PeerBlock) * nranks_, cudaMemcpyDeviceToHost, stream_));
    RG_CUDA(cudaStreamSynchronize(stream_));
    bool ok = want != 0;
    for (int r = 0; r < nranks_ && ok; ++r) ok = all[r].want != 0 && all[r].host == mine.host && (r == rank_ || all[r].pid != mine.pid);
    bool hasLo, hasHi;
    zNeighbours(&hasLo, &hasHi);
    const int lo = (rank_ + nranks_ - 1) % nranks_, hi = (rank_ + 1) % nranks_;
    auto open = [&](int r, Peer* p) {
      p->ksize = all[r].ksize;
      const unsigned fl = cudaIpcMemLazyEnablePeerAccess;
      if (cudaIpcOpenMemHandle((void**)&p->U[0], all[r].state[0], fl) != cudaSuccess ||
          cudaIpcOpenMemHandle((void**)&p->U[1], all[r].state[1], fl) != cudaSuccess ||
          cudaIpcOpenMemHandle((void**)&p->flags, all[r].flags, fl) != cudaSuccess) {
        cudaGetLastError();
        return false;
      }
      return true;
    };
    if (ok && hasLo) { ok = open(lo, &peerLo_); peerLoOpened_ = true; }
    if (ok && hasHi) {
      if (hasLo && hi == lo) peerHi_ = peerLo_;  // two ranks, periodic: one neighbour on both sides (a handle opens once)
      else { ok = open(hi, &peerHi_); peerHiOpened_ = true; }
    }
    // all or nothing: the exchange protocol of a rank must match its neighbours'
    double* dFail = reinterpret_cast<double*>(dBlocks);
    const double fail = ok ? 0.0 : 1.0;
    double failAny = 1.0;
    RG_CUDA(cudaMemcpyAsync(dFail, &fail, sizeof fail, cudaMemcpyHostToDevice, stream_));
    ncclCheck(nccl_->AllReduce(dFail, dFail, 1, NcclApi::kFloat64, NcclApi::kMax, comm_, stream_), "allreduce(peer halo)");
    RG_CUDA(cudaMemcpyAsync(&failAny, dFail, sizeof failAny, cudaMemcpyDeviceToHost, stream_));
    RG_CUDA(cudaStreamSynchronize(stream_));
    cudaFree(dBlocks);
    p2p_ = failAny == 0.0;
    p2pBase_[0] = dU_[0];
    p2pBase_[1] = dU_[1];
    if (!p2p_) {
      closePeerMappings();
      return;
    }
    int dev = 0, pitch = 0;
    RG_CUDA(cudaGetDevice(&dev));
    if (cudaDeviceGetAttribute(&pitch, cudaDevAttrMaxPitch, dev) == cudaSuccess && pitch > 0) maxPitch_ = (size_t)pitch;
    // can a stream memory operation write straight into the neighbour's flag array?  Probe slots 6 (written by the rank
    // below) and 7 (by the rank above) and read them back through the mapping; otherwise flags travel by 8-byte copies.
    // Every rank decides alike only if the answer is the same everywhere: agree through the communicator.
    bool direct = true;
    const unsigned long long magic = 0x5eedf1a9ull;
    auto probe = [&](Peer& p, int slot) {
      if (writeValue_(stream_, (unsigned long long)(uintptr_t)(p.flags + slot), magic, 0u) != 0) { direct = false; return; }
      unsigned long long v = 0;
      if (cudaStreamSynchronize(stream_) != cudaSuccess ||
          cudaMemcpy(&v, p.flags + slot, sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess || v != magic) direct = false;
    };
    if (hasHi) probe(peerHi_, 6);
    if (hasLo && direct) probe(peerLo_, 7);
    cudaGetLastError();
    const double nd = direct ? 0.0 : 1.0;
    double ndAny = 1.0;
    double* dAgree = reinterpret_cast<double*>(dFlags_ + FL_SLOTS);
    RG_CUDA(cudaMemcpyAsync(dAgree, &nd, sizeof nd, cudaMemcpyHostToDevice, stream_));
    ncclCheck(nccl_->AllReduce(dAgree, dAgree, 1, NcclApi::kFloat64, NcclApi::kMax, comm_, stream_), "allreduce(peer flags)");
    RG_CUDA(cudaMemcpyAsync(&ndAny, dAgree, sizeof ndAny, cudaMemcpyDeviceToHost, stream_));
    RG_CUDA(cudaStreamSynchronize(stream_));
    RG_CUDA(cudaMemsetAsync(dAgree, 0, sizeof(double), stream_));
    RG_CUDA(cudaStreamSynchronize(stream_));
    directFlags_ = ndAny == 0.0;
  }

  void closePeerMappings() {
    auto close = [](Peer* p) {
      for (int b = 0; b < 2; ++b) if (p->U[b]) cudaIpcCloseMemHandle(p->U[b]);
      if (p->flags) cudaIpcCloseMemHandle(p->flags);
      cudaGetLastError();
    };
    if (peerLoOpened_) close(&peerLo_);
    if (peerHiOpened_) close(&peerHi_);
    peerLo_ = Peer();
    peerHi_ = Peer();
    peerLoOpened_ = peerHiOpened_ = false;
  }

  // destruction with peer mappings is collective: no rank frees its arrays before every rank has unmapped them
  void closePeerHalo() {
    if (p2p_) {
      closePeerMappings();
      if (comm_ && nccl_ && dFlags_) {
        double* d = reinterpret_cast<double*>(dFlags_);
        if (nccl_->AllReduce(d, d, 1, NcclApi::kFloat64, NcclApi::kMax, comm_, stream_) == 0) cudaStreamSynchronize(stream_);
      }
      p2p_ = false;
    }
    cudaFree(dFlags_);
    dFlags_ = nullptr;
  }

  void memOp(StreamMemOp op, cudaStream_t st, const unsigned long long* addr, unsigned long long value, const char* what) {
    const int rc = op(st, (unsigned long long)(uintptr_t)addr, value, 0u);  // wait: CU_STREAM_WAIT_VALUE_GEQ = 0
    if (rc != 0) throw std::runtime_error(std::string("CUDA driver error ") + std::to_string(rc) + " in " + what);
  }

  // the gw boundary planes of variables [v0, v0+nv) of state array U to the z neighbours; false: not a peer-mapped array
  bool peerExchange(T* U, int v0, int nv, bool hasLo, bool hasHi, cudaStream_t st) {
    if (!p2p_) return false;
    const int b = U == p2pBase_[0] ? 0 : (U == p2pBase_[1] ? 1 : -1);
    if (b < 0) return false;
    const unsigned long long s = ++haloSeq_;
    const int gw = kp_.gw;
    const size_t plane = (size_t)kp_.isize * kp_.jsize, comp = plane * kp_.ksize, n = plane * gw;
    auto signal = [&](unsigned long long* peerFlag) {
      if (directFlags_) {  // one stream memory operation straight into the neighbour's flag
        memOp(writeValue_, st, peerFlag, s, "cuStreamWriteValue64(peer)");
        return;
      }
      unsigned long long* slot = dFlags_ + FL_SLOTS + (signalCount_++ % FL_NSLOT);
      memOp(writeValue_, st, slot, s, "cuStreamWriteValue64");
      RG_CUDA(cudaMemcpyAsync(peerFlag, slot, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
    };
    // nv rows of gw planes, one row per variable: ONE strided copy per direction when the pitches allow it
    auto put = [&](T* dst, size_t dstComp, const T* src) {
      if (dstComp * sizeof(T) <= maxPitch_ && comp * sizeof(T) <= maxPitch_) {
        RG_CUDA(cudaMemcpy2DAsync(dst, dstComp * sizeof(T), src, comp * sizeof(T), n * sizeof(T), (size_t)nv,
                                  cudaMemcpyDeviceToDevice, st));
      } else {
        for (int v = 0; v < nv; ++v)
          RG_CUDA(cudaMemcpyAsync(dst + (size_t)v * dstComp, src + (size_t)v * comp, n * sizeof(T), cudaMemcpyDeviceToDevice, st));
      }
    };
    // I am my lower neighbour's upper neighbour and the other way round
    if (hasLo) signal(peerLo_.flags + FL_READY_FROM_HI);
    if (hasHi) signal(peerHi_.flags + FL_READY_FROM_LO);
    if (hasHi) {  // my top inner planes -> the low ghost planes of the rank above
      memOp(waitValue_, st, dFlags_ + FL_READY_FROM_HI, s, "cuStreamWaitValue64");
      const size_t compHi = plane * peerHi_.ksize;
      put(peerHi_.U[b] + (size_t)v0 * compHi, compHi, U + (size_t)v0 * comp + (size_t)(kp_.ksize - 2 * gw) * plane);
      signal(peerHi_.flags + FL_DATA_FROM_LO);
    }
    if (hasLo) {  // my bottom inner planes -> the high ghost planes of the rank below
      memOp(waitValue_, st, dFlags_ + FL_READY_FROM_LO, s, "cuStreamWaitValue64");
      const size_t compLo = plane * peerLo_.ksize;
      put(peerLo_.U[b] + (size_t)v0 * compLo + (size_t)(peerLo_.ksize - gw) * plane, compLo, U + (size_t)v0 * comp + (size_t)gw * plane);
      signal(peerLo_.flags + FL_DATA_FROM_HI);
    }
    if (hasLo) memOp(waitValue_, st, dFlags_ + FL_DATA_FROM_LO, s, "cuStreamWaitValue64");
    if (hasHi) memOp(waitValue_, st, dFlags_ + FL_DATA_FROM_HI, s, "cuStreamWaitValue64");
    return true;
  }

  void exchangeZ(T* U, bool hasLo, bool hasHi, cudaStream_t st) {
    const int gw = kp_.gw, lo = (rank_ + nranks_ - 1) % nranks_, hi = (rank_ + 1) % nranks_;
    const size_t plane = (size_t)kp_.isize * kp_.jsize, comp = plane * kp_.ksize, n = plane * gw;
    haloBytesPerStep_ = (double)((hasLo ? 1 : 0) + (hasHi ? 1 : 0)) * n * kp_.nvar * sizeof(T);
    haloOnNccl_ = !peerExchange(U, 0, kp_.nvar, hasLo, hasHi, st);
    if (!haloOnNccl_) return;
    const int dtype = sizeof(T) == 8 ? NcclApi::kFloat64 : NcclApi::kFloat32;
    ncclCheck(nccl_->GroupStart(), "group start");
    for (int v = 0; v < kp_.nvar; ++v) {
      T* base = U + (size_t)v * comp;
      if (hasHi) ncclCheck(nccl_->Send(base + (size_t)(kp_.ksize - 2 * gw) * plane, n, dtype, hi, comm_, st), "send up");
      if (hasLo) ncclCheck(nccl_->Send(base + (size_t)gw * plane, n, dtype, lo, comm_, st), "send down");
      if (hasLo) ncclCheck(nccl_->Recv(base, n, dtype, lo, comm_, st), "recv from below");
      if (hasHi) ncclCheck(nccl_->Recv(base + (size_t)(kp_.ksize - gw) * plane, n, dtype, hi, comm_, st), "recv from above");
    }
    ncclCheck(nccl_->GroupEnd(), "group end");
  }

  // B of the ghost planes next to INTERIOR slab interfaces only (no periodic wrap): used between the
  // resistive CT update and the resistive energy flux, where the mono-domain run sees updated inner
  // cells across an interface but stale ghosts across the global z boundary (a property of the
  // reference's sequence, mhd_godunov_unsplit_cpu_v3.cpp:666-680; tests/test_slab_protocol_gloo.py)
  void exchangeZInteriorB(T* U, cudaStream_t st) {
    const bool hasLo = rank_ > 0, hasHi = rank_ < nranks_ - 1;
    const int gw = kp_.gw, lo = rank_ - 1, hi = rank_ + 1;
    const size_t plane = (size_t)kp_.isize * kp_.jsize, comp = plane * kp_.ksize, n = plane * gw;
    if (peerExchange(U, IA, 3, hasLo, hasHi, st)) return;
    const int dtype = sizeof(T) == 8 ? NcclApi::kFloat64 : NcclApi::kFloat32;
    ncclCheck(nccl_->GroupStart(), "group start");
    for (int v = IA; v <= IC; ++v) {
      T* base = U + (size_t)v * comp;
      if (hasHi) ncclCheck(nccl_->Send(base + (size_t)(kp_.ksize - 2 * gw) * plane, n, dtype, hi, comm_, st), "send up");
      if (hasLo) ncclCheck(nccl_->Send(base + (size_t)gw * plane, n, dtype, lo, comm_, st), "send down");
      if (hasLo) ncclCheck(nccl_->Recv(base, n, dtype, lo, comm_, st), "recv from below");
      if (hasHi) ncclCheck(nccl_->Recv(base + (size_t)(kp_.ksize - gw) * plane, n, dtype, hi, comm_, st), "recv from above");
    }
    ncclCheck(nccl_->GroupEnd(), "group end");
  }

  // ---- scratch / chunking ------------------------------------------------------------------------
  void freeScratch() {
    if (sc_.W) {  // cudaFree(nullptr) is a no-op for the arrays the hydro path does not use
      cudaFree(sc_.Q); cudaFree(sc_.W); cudaFree(sc_.F); cudaFree(sc_.E); cudaFree(sc_.EL); cudaFree(sc_.strips);
      cudaFree(sc_.hbuf); cudaFree(sc_.hsync);
      deviceBytes_ -= scratchBytes_;
    }
    sc_ = MhdScratch<T>();
    scratchBytes_ = 0;
    chunkPlanes_ = 0;
  }

  // The headline configuration runs on the two fused kernels and needs W only; the separate-kernel path (rotating
  // frame fluxes/emfs, other solvers, slope_type 3, or a knob turned off) also needs Q, F, E, EL.
  bool fusedPairUsable() {
    return fusedTraceRequested() && fusedRequested() && MhdKernels<T>::fusedTraceAvailable(kp_) &&
           MhdKernels<T>::fusedUpdateEligible(kp_);
  }
  void ensureScratchMhd3d(bool generic) {
    if (sc_.W && generic && !sc_.F) freeScratch();  // a knob asked for the separate kernels after a W-only allocation
    if (sc_.W && sc_.fused && fusedHandoffRequested() && !sc_.hbuf) freeScratch();  // ... or for the hand-off tiles
    if (sc_.W) return;
    const size_t plane = (size_t)kp_.isize * kp_.jsize;
    const size_t perPlane = plane * sizeof(T) * (generic ? (8 + NW_MHD + 15 + 3 + 3) : NW_MHD);
    const int updPlanes = kp_.ksize - 2 * kp_.gw + 1;  // gw .. ksize-gw inclusive
    int chunk = updPlanes;
    size_t freeB = 0, totalB = 0;
    RG_CUDA(cudaMemGetInfo(&freeB, &totalB));
    const size_t budget = (size_t)(0.85 * (double)freeB);
    const long fit = (long)(budget / perPlane) - 4;
    if (fit < chunk) chunk = (int)std::max<long>(fit, 1);
    // scratch arrays are indexed with 32-bit element offsets: keep the largest (W) below 2^31
    const long idxFit = (long)(2147483647LL / ((long long)NW_MHD * (long long)plane)) - 4;
    if (idxFit < chunk) chunk = (int)std::max<long>(idxFit, 1);
    if (userChunk_ > 0) chunk = std::min(std::min(userChunk_, updPlanes), chunk);
    chunkPlanes_ = chunk;
    sc_.planes = chunk + 4;
    RG_CUDA(cudaMalloc(&sc_.W, plane * sc_.planes * NW_MHD * sizeof(T)));
    if (generic) {
      RG_CUDA(cudaMalloc(&sc_.Q, plane * sc_.planes * 8 * sizeof(T)));
      RG_CUDA(cudaMalloc(&sc_.F, plane * sc_.planes * 15 * sizeof(T)));
      RG_CUDA(cudaMalloc(&sc_.E, plane * sc_.planes * 3 * sizeof(T)));
      RG_CUDA(cudaMalloc(&sc_.EL, plane * sc_.planes * 3 * sizeof(T)));
    }
    scratchBytes_ = perPlane * sc_.planes;
    if (shearingBox()) {  // fluxes / emfs of the four x-border position columns (fused kernels of the shearing box)
      const size_t stripBytes = (size_t)18 * sc_.planes * kp_.jsize * 4 * sizeof(T);
      RG_CUDA(cudaMalloc(&sc_.strips, stripBytes));
      RG_CUDA(cudaMemsetAsync(sc_.strips, 0, stripBytes, stream_));
      scratchBytes_ += stripBytes;
    }
    MhdKernels<T>::fusedPrepare(kp_, sc_);
    if (sc_.fused) {  // hand-off records of the fused update: the largest launch is a whole chunk (+ the ghost-face plane)
      size_t reals = 0, ints = 0;
      for (int n : {chunk + 1, chunk, std::max(chunk - 2 * kp_.gw, 1), kp_.gw, kp_.gw + 1}) {
        size_t r = 0, i = 0;
        MhdKernels<T>::fusedHandoffSize(kp_, n, &r, &i);
        reals = std::max(reals, r);
        ints = std::max(ints, i);
      }
      if (reals > 0) {
        RG_CUDA(cudaMalloc(&sc_.hbuf, reals * sizeof(T)));
        RG_CUDA(cudaMalloc(&sc_.hsync, ints * sizeof(int)));
        sc_.hbufReals = reals;
        sc_.hsyncInts = ints;
        scratchBytes_ += reals * sizeof(T) + ints * sizeof(int);
      }
    }
    deviceBytes_ += scratchBytes_;
  }

  // ---- 3D MHD step: reference godunov_unsplit_cpu/gpu (MHDRunGodunov.cpp:623-672, 1447-1503) ------
  void stepMhd3d(int src, int dst, T dt) {
    ensureScratchMhd3d(!fusedPairUsable());
    if (!sc_.fused && !sc_.F) {  // the tensor map / kernel attribute could not be set up: separate kernels after all
      freeScratch();
      ensureScratchMhd3d(true);
    }
    const T* Uold = dU_[src];
    T* Unew = dU_[dst];
    const int gw = kp_.gw, kN = kp_.ksize - gw;
    unsigned long long* slots = dMax_ + (size_t)dst * MAX_SLOTS;
    RG_CUDA(cudaMemsetAsync(slots, 0, MAX_SLOTS * sizeof(unsigned long long), stream_));
    // ghost planes outside the update box keep the (ghost-filled) old values, like copyTo
    phase(PH_COPY, [&] {
      MhdKernels<T>::copyPlanes(kp_, Uold, Unew, 0, gw, stream_);
      MhdKernels<T>::copyPlanes(kp_, Uold, Unew, kN + 1, kp_.ksize, stream_);
    });
    auto runChunk = [&](int ka, int kb) {  // update planes [ka, kb)
      const int fhi = std::min(kb, kN);
      MhdScratch<T> sc = sc_;
      sc.kbase = ka - 2;
      if (fusedTraceRequested() && MhdKernels<T>::fusedTraceAvailable(kp_)) {
        phase(PH_TRACE, [&] { MhdKernels<T>::fusedTrace(kp_, Uold, sc, ka - 1, fhi + 1, dt, stream_); });
      } else {
        phase(PH_PRIM, [&] { MhdKernels<T>::prim(kp_, Uold, sc, ka - 2, fhi + 2, dt, stream_); });
        phase(PH_PRIM, [&] { MhdKernels<T>::elec(kp_, Uold, sc, ka - 1, fhi + 2, stream_); });
        phase(PH_TRACE, [&] { MhdKernels<T>::trace(kp_, Uold, sc, ka - 1, fhi + 1, dt, stream_); });
      }
      if (sc.fused && fusedRequested()) {
        phase(PH_FUSED, [&] { MhdKernels<T>::fusedFluxEmfUpdate(kp_, Uold, Unew, sc, ka, kb, dt, slots, stream_); });
        phase(PH_COPY, [&] { MhdKernels<T>::copyOutsideBox(kp_, Uold, Unew, ka, kb, stream_); });
        return;
      }
      phase(PH_FLUX, [&] { MhdKernels<T>::flux(kp_, sc, ka, fhi + 1, stream_); });
      phase(PH_EMF, [&] { MhdKernels<T>::emf(kp_, sc, ka, fhi + 1, stream_); });
      phase(PH_UPDATE, [&] { MhdKernels<T>::update(kp_, Uold, Unew, sc, ka, kb, dt, slots, stream_); });
    };
    auto runRange = [&](int k0, int k1) {
      for (int ka = k0; ka < k1; ka += chunkPlanes_) runChunk(ka, std::min(ka + chunkPlanes_, k1));
    };
    // overlap needs three disjoint plane ranges: bottom gw planes, top gw planes (+ the ghost-face
    // plane kN), interior
    const bool overlap = nranks_ > 1 && overlapHalo_ && (kN - gw) >= 3 * gw && !dissipative();
    // one chunk holds the whole slab and both fused kernels run: the trace does not depend on the halo, so it runs ONCE
    // over the slab and only the update is cut into the three ranges (two launches and their pipeline fills less)
    const bool traceOnce = overlap && chunkPlanes_ >= kN + 1 - gw && sc_.fused && fusedRequested() &&
                           fusedTraceRequested() && MhdKernels<T>::fusedTraceAvailable(kp_);
    if (overlap && p2p_) {
      // copy-engine halo: it needs no SM, so nothing is gained by cutting the update into three launches -- the whole
      // slab in one pass like on one GPU, then the halo of the NEW state on the communication stream (startLateHalo)
      runRange(gw, kN + 1);
      startLateHalo(dst);
    } else if (traceOnce) {
      MhdScratch<T> sc = sc_;
      sc.kbase = gw - 2;
      phase(PH_TRACE, [&] { MhdKernels<T>::fusedTrace(kp_, Uold, sc, gw - 1, kN + 1, dt, stream_); });
      auto updateRange = [&](int ka, int kb) {
        phase(PH_FUSED, [&] { MhdKernels<T>::fusedFluxEmfUpdate(kp_, Uold, Unew, sc, ka, kb, dt, slots, stream_); });
        phase(PH_COPY, [&] { MhdKernels<T>::copyOutsideBox(kp_, Uold, Unew, ka, kb, stream_); });
      };
      updateRange(gw, 2 * gw);
      updateRange(kN - gw, kN + 1);
      startEarlyHalo(dst);
      updateRange(2 * gw, kN - gw);
    } else if (overlap) {
      runRange(gw, 2 * gw);
      runRange(kN - gw, kN + 1);
      startEarlyHalo(dst);
      runRange(2 * gw, kN - gw);
    } else {
      runRange(gw, kN + 1);
    }
    ghostsValid_[dst] = false;
    dtCached_[dst] = true;  // the update kernel reduced the inverse dt of the new state
    if (dissipative()) {    // reference mhd_godunov_unsplit_cpu_v3.cpp:661-693
      fillGhosts(dst, 0, kp_.ksize);
      dissipativeTerms(dst, dt);
    }
  }

  // ---- dissipative terms on the new state (its ghosts have just been refreshed) -------------------
  bool dissipative() const { return rp_.dim == 3 && (kp_.nu > T(0) || kp_.eta > T(0)); }
  void dissipativeTerms(int b, T dt) {
    T* U = dU_[b];
    if (!dDiss_) {
      const size_t bytes = cells_ * 12 * sizeof(T);
      RG_CUDA(cudaMalloc(&dDiss_, bytes));
      deviceBytes_ += bytes;
    }
    const bool resistiveEnergy = kp_.eta > T(0) && !(kp_.cIso > T(0));
    if (kp_.eta > T(0))
      phase(PH_DISS, [&] {
        DissKernels<T>::resistEmf(kp_, U, dDiss_, stream_);
        DissKernels<T>::ctUpdate(kp_, U, dDiss_, dt, stream_);
      });
    if (resistiveEnergy && nranks_ > 1) phase(PH_HALO, [&] { exchangeZInteriorB(U, stream_); });
    phase(PH_DISS, [&] {
      if (resistiveEnergy) DissKernels<T>::resistEnergy(kp_, U, dt, stream_);
      if (kp_.nu > T(0)) {
        DissKernels<T>::viscFlux(kp_, U, dDiss_, dt, stream_);
        DissKernels<T>::viscUpdate(kp_, U, dDiss_, stream_);
      }
    });
    ghostsValid_[b] = false;
    dtCached_[b] = false;  // the state changed after the update kernel reduced its inverse dt
  }

  bool shearingBox() const {
    return rp_.mhdEnabled && rp_.dim == 3 && rp_.bc[0] == BC_SHEARINGBOX && rp_.bc[1] == BC_SHEARINGBOX &&
           kp_.Omega0 > T(0);
  }
  // y shift of the opposite x border at time t: whole cells and fraction of dy (MHDRunGodunov.cpp:3213-3216)
  void shearShift(T t, int* jplus, T* frac) const {
    T deltay = T(1.5) * kp_.Omega0 * (kp_.dx * rp_.nx) * t;
    deltay = std::fmod(deltay, kp_.dy * rp_.ny);
    *jplus = static_cast<int>(deltay / kp_.dy);
    *frac = std::fmod(deltay, kp_.dy) / kp_.dy;
  }
  // reference make_all_boundaries_shear (MHDRunGodunov.cpp:3763-3793): Y, shear-X, Z, Y at time t + dt
  void fillGhostsShear(int b, T dt) {
    T* U = dU_[b];
    int jplus; T frac;
    shearShift(static_cast<T>(totalTime) + dt, &jplus, &frac);
    const int gw = kp_.gw, kN = kp_.ksize - gw;
    if (haloDone_[b]) {
      // the planes next to the slab interfaces were filled (Y, shear-X) and exchanged by startEarlyHaloShear() while the
      // interior was being updated: the same per-plane operations on the remaining inner planes, then the order of the
      // reference again (physical z faces, Y)
      bool hasLo, hasHi;
      zNeighbours(&hasLo, &hasHi);
      phase(PH_BOUNDARY, [&] {
        MhdKernels<T>::fillBoundary(kp_, U, 1, rp_.bc[2], rp_.bc[3], false, false, 2 * gw, kN - gw, stream_);
        MhdKernels<T>::shearGhosts(kp_, U, jplus, frac, 2 * gw, kN - gw, stream_);
      });
      RG_CUDA(cudaStreamWaitEvent(stream_, evHalo_, 0));
      haloDone_[b] = false;
      phase(PH_BOUNDARY, [&] {
        if (!hasLo || !hasHi) fillZFaces(U, hasLo, hasHi);
        MhdKernels<T>::fillBoundary(kp_, U, 1, rp_.bc[2], rp_.bc[3], false, false, 0, kp_.ksize, stream_);
      });
      return;
    }
    phase(PH_BOUNDARY, [&] {
      MhdKernels<T>::fillBoundary(kp_, U, 1, rp_.bc[2], rp_.bc[3], false, false, 0, kp_.ksize, stream_);
      MhdKernels<T>::shearGhosts(kp_, U, jplus, frac, 0, kp_.ksize, stream_);
    });
    if (nranks_ == 1) {
      phase(PH_BOUNDARY, [&] {
        fillZFaces(U, false, false);
      });
    } else {
      bool hasLo, hasHi;
      zNeighbours(&hasLo, &hasHi);
      if (!hasLo || !hasHi)
        phase(PH_BOUNDARY, [&] {
          fillZFaces(U, hasLo, hasHi);
        });
      phase(PH_HALO, [&] { exchangeZ(U, hasLo, hasHi, stream_); });
    }
    phase(PH_BOUNDARY, [&] {
      MhdKernels<T>::fillBoundary(kp_, U, 1, rp_.bc[2], rp_.bc[3], false, false, 0, kp_.ksize, stream_);
    });
  }

  // Early halo of the shearing box: the rotating step fills the ghosts of the NEW state at its end (Y, shear-X, Z, Y).
  // Y and shear-X act plane by plane, so as soon as the gw inner planes next to each slab interface are final they are
  // filled and exchanged on the communication stream while the interior is still being updated; fillGhostsShear()
  // finishes the remaining planes.  (Not with z-stratified faces: their ghost planes keep two entries of the first fill.)
  void startEarlyHaloShear(int b, T dt) {
    T* U = dU_[b];
    const int gw = kp_.gw, kN = kp_.ksize - gw;
    int jplus; T frac;
    shearShift(static_cast<T>(totalTime) + dt, &jplus, &frac);
    bool hasLo, hasHi;
    zNeighbours(&hasLo, &hasHi);
    RG_CUDA(cudaEventRecord(evEdge_, stream_));
    RG_CUDA(cudaStreamWaitEvent(comm_stream_, evEdge_, 0));
    MhdKernels<T>::fillBoundary(kp_, U, 1, rp_.bc[2], rp_.bc[3], false, false, gw, 2 * gw, comm_stream_);
    MhdKernels<T>::shearGhosts(kp_, U, jplus, frac, gw, 2 * gw, comm_stream_);
    MhdKernels<T>::fillBoundary(kp_, U, 1, rp_.bc[2], rp_.bc[3], false, false, kN - gw, kN, comm_stream_);
    MhdKernels<T>::shearGhosts(kp_, U, jplus, frac, kN - gw, kN, comm_stream_);
    exchangeZ(U, hasLo, hasHi, comm_stream_);
    RG_CUDA(cudaEventRecord(evHalo_, comm_stream_));
    haloDone_[b] = true;
  }

  // ---- 3D MHD step in the rotating frame (Omega0 > 0), with or without shearing-box boundaries:
  //      reference godunov_unsplit_rotating_cpu / _gpu (MHDRunGodunov.cpp:1511, 2031)
  void stepMhd3dRotating(int src, int dst, T dt) {
    ensureScratchMhd3d(!fusedPairUsable());
    if (!sc_.fused && !sc_.F) {
      freeScratch();
      ensureScratchMhd3d(true);
    }
    const T* Uold = dU_[src];
    T* Unew = dU_[dst];
    const int gw = kp_.gw, kN = kp_.ksize - gw;
    unsigned long long* slots = dMax_ + (size_t)dst * MAX_SLOTS;
    RG_CUDA(cudaMemsetAsync(slots, 0, MAX_SLOTS * sizeof(unsigned long long), stream_));
    const int shear = shearingBox() ? 1 : 0;
    int jplus = 0; T frac = T(0);
    if (shear) shearShift(static_cast<T>(totalTime) + dt / 2, &jplus, &frac);
    phase(PH_COPY, [&] {
      MhdKernels<T>::copyPlanes(kp_, Uold, Unew, 0, gw, stream_);
      MhdKernels<T>::copyPlanes(kp_, Uold, Unew, kN + 1, kp_.ksize, stream_);
    });
    bool fusedRotNoDt = false;
    auto runChunk = [&](int ka, int kb) {  // update planes [ka, kb)
      const int fhi = std::min(kb, kN);
      MhdScratch<T> sc = sc_;
      sc.kbase = ka - 2;
      if (fusedTraceRequested() && MhdKernels<T>::fusedTraceAvailable(kp_)) {
        phase(PH_TRACE, [&] { MhdKernels<T>::fusedTrace(kp_, Uold, sc, ka - 1, fhi + 1, dt, stream_); });
      } else {
        phase(PH_PRIM, [&] { MhdKernels<T>::prim(kp_, Uold, sc, ka - 2, fhi + 2, dt, stream_); });
        phase(PH_PRIM, [&] { MhdKernels<T>::elec(kp_, Uold, sc, ka - 1, fhi + 2, stream_); });
        phase(PH_TRACE, [&] { MhdKernels<T>::trace(kp_, Uold, sc, ka - 1, fhi + 1, dt, stream_); });
      }
      if (sc.fused && fusedRequested()) {  // fused flux + emf + update (+ the border columns of the shearing box)
        phase(PH_FUSED, [&] {
          MhdKernels<T>::fusedFluxEmfUpdate(kp_, Uold, Unew, sc, ka, kb, dt, slots, stream_, shear, jplus, frac);
        });
        phase(PH_COPY, [&] { MhdKernels<T>::copyOutsideBox(kp_, Uold, Unew, ka, kb, stream_); });
        fusedRotNoDt = !rotDtInKernel();
        return;
      }
      phase(PH_FLUX, [&] { MhdKernels<T>::flux(kp_, sc, ka, fhi + 1, stream_); });
      phase(PH_EMF, [&] { MhdKernels<T>::emf(kp_, sc, ka, fhi + 1, stream_); });
      phase(PH_UPDATE, [&] {
        MhdKernels<T>::updateRotating(kp_, Uold, Unew, sc, ka, kb, dt, shear, jplus, frac, slots, stream_);
      });
    };
    auto runRange = [&](int k0, int k1) {
      for (int ka = k0; ka < k1; ka += chunkPlanes_) runChunk(ka, std::min(ka + chunkPlanes_, k1));
    };
    // the ghosts of the new state are filled at the END of this step: with slabs, the planes next to the interfaces are
    // updated first and travel on the communication stream while the interior is updated (see startEarlyHaloShear)
    const bool stratZ = rp_.bc[4] == BC_Z_STRATIFIED || rp_.bc[5] == BC_Z_STRATIFIED;
    const bool overlap = nranks_ > 1 && overlapHalo_ && (kN - gw) >= 3 * gw && !dissipative() && !stratZ;
    const bool traceOnce = overlap && chunkPlanes_ >= kN + 1 - gw && sc_.fused && fusedRequested() &&
                           fusedTraceRequested() && MhdKernels<T>::fusedTraceAvailable(kp_);
    if (overlap && p2p_) {  // as in stepMhd3d: copy-engine halo after ONE pass over the slab
      runRange(gw, kN + 1);
      if (shear) startEarlyHaloShear(dst, dt);
      else startLateHalo(dst);
    } else if (traceOnce) {  // as in stepMhd3d: one trace launch over the slab, the update in three ranges
      MhdScratch<T> sc = sc_;
      sc.kbase = gw - 2;
      phase(PH_TRACE, [&] { MhdKernels<T>::fusedTrace(kp_, Uold, sc, gw - 1, kN + 1, dt, stream_); });
      auto updateRange = [&](int ka, int kb) {
        phase(PH_FUSED, [&] {
          MhdKernels<T>::fusedFluxEmfUpdate(kp_, Uold, Unew, sc, ka, kb, dt, slots, stream_, shear, jplus, frac);
        });
        phase(PH_COPY, [&] { MhdKernels<T>::copyOutsideBox(kp_, Uold, Unew, ka, kb, stream_); });
      };
      updateRange(gw, 2 * gw);
      updateRange(kN - gw, kN + 1);
      if (shear) startEarlyHaloShear(dst, dt);
      else startEarlyHalo(dst);
      updateRange(2 * gw, kN - gw);
      fusedRotNoDt = !rotDtInKernel();
    } else if (overlap) {
      runRange(gw, 2 * gw);
      runRange(kN - gw, kN + 1);
      if (shear) startEarlyHaloShear(dst, dt);
      else startEarlyHalo(dst);
      runRange(2 * gw, kN - gw);
    } else {
      runRange(gw, kN + 1);
    }
    dtCached_[dst] = !fusedRotNoDt;  // (knob rot_dt = 0: compute_dt runs the stand-alone reduction on the new state)
    if (dissipative()) {  // reference MHDRunGodunov.cpp:3379-3419: ghost refresh, then the dissipative terms
      if (shear) fillGhostsShear(dst, dt);
      else fillGhosts(dst, 0, kp_.ksize);
      dissipativeTerms(dst, dt);
    }
    // ghosts of the NEW state, at the end of the step
    if (shear) fillGhostsShear(dst, dt);
    else fillGhosts(dst, 0, kp_.ksize);
    ghostsValid_[dst] = true;
  }

  // ---- 2D MHD step: reference godunov_unsplit_cpu + _v1 (mhd_godunov_unsplit_cpu_v1.cpp:36-243)
  void stepMhd2d(int src, int dst, T dt) {
    if (!sc_.W) {
      const size_t plane = (size_t)kp_.isize * kp_.jsize;
      RG_CUDA(cudaMalloc(&sc_.Q, plane * 8 * sizeof(T)));
      RG_CUDA(cudaMalloc(&sc_.W, plane * NW_MHD2D * sizeof(T)));
      RG_CUDA(cudaMalloc(&sc_.F, plane * 12 * sizeof(T)));
      RG_CUDA(cudaMalloc(&sc_.E, plane * sizeof(T)));
      scratchBytes_ = plane * (8 + NW_MHD2D + 12 + 1) * sizeof(T);
      deviceBytes_ += scratchBytes_;
      sc_.planes = 1;
      chunkPlanes_ = 1;
    }
    unsigned long long* slots = dMax_ + (size_t)dst * MAX_SLOTS;
    RG_CUDA(cudaMemsetAsync(slots, 0, MAX_SLOTS * sizeof(unsigned long long), stream_));
    phase(PH_UPDATE, [&] {
      Mhd2dKernels<T>::step(kp_, dU_[src], dU_[dst], sc_.Q, sc_.W, sc_.F, sc_.E, dt, slots, stream_);
    });
    ghostsValid_[dst] = false;
    dtCached_[dst] = true;
  }

  // ---- 2D hydro step: reference HydroRunGodunov::godunov_unsplit_cpu + _v1, TWO_D branch (HydroRunGodunov.cpp:2437-2655)
  void stepHydro2d(int src, int dst, T dt) {
    if (!sc_.W) {
      const size_t plane = (size_t)kp_.isize * kp_.jsize;
      RG_CUDA(cudaMalloc(&sc_.W, plane * NW_HYDRO2D * sizeof(T)));
      scratchBytes_ = plane * NW_HYDRO2D * sizeof(T);
      deviceBytes_ += scratchBytes_;
      sc_.planes = 1;
      chunkPlanes_ = 1;
    }
    unsigned long long* slots = dMax_ + (size_t)dst * MAX_SLOTS;
    RG_CUDA(cudaMemsetAsync(slots, 0, MAX_SLOTS * sizeof(unsigned long long), stream_));
    phase(PH_UPDATE, [&] { Hydro2dKernels<T>::step(kp_, dU_[src], dU_[dst], sc_.W, dt, slots, stream_); });
    ghostsValid_[dst] = false;
    dtCached_[dst] = true;
  }

  // ---- 3D hydro step: reference HydroRunGodunov::godunov_unsplit_cpu + _v1 (HydroRunGodunov.cpp:1820, 2658)
  void ensureScratchHydro3d() {
    if (sc_.W) return;
    const size_t plane = (size_t)kp_.isize * kp_.jsize;
    const size_t perPlane = plane * sizeof(T) * NW_HYDRO;
    const int updPlanes = kp_.nz;
    int chunk = updPlanes;
    size_t freeB = 0, totalB = 0;
    RG_CUDA(cudaMemGetInfo(&freeB, &totalB));
    const long fit = (long)((size_t)(0.85 * (double)freeB) / perPlane) - 2;
    if (fit < chunk) chunk = (int)std::max<long>(fit, 1);
    const long idxFit = (long)(2147483647LL / ((long long)NW_HYDRO * (long long)plane)) - 2;
    if (idxFit < chunk) chunk = (int)std::max<long>(idxFit, 1);
    if (userChunk_ > 0) chunk = std::min(std::min(userChunk_, updPlanes), chunk);
    chunkPlanes_ = chunk;
    sc_.planes = chunk + 2;
    RG_CUDA(cudaMalloc(&sc_.W, plane * sc_.planes * NW_HYDRO * sizeof(T)));
    scratchBytes_ = perPlane * sc_.planes;
    deviceBytes_ += scratchBytes_;
  }

  void stepHydro3d(int src, int dst, T dt) {
    const bool fused = hydroFusedRequested();  // one kernel, no traced-state scratch
    if (!fused) ensureScratchHydro3d();
    else if (chunkPlanes_ == 0) chunkPlanes_ = kp_.nz;
    const T* Uold = dU_[src];
    T* Unew = dU_[dst];
    const int gw = kp_.gw, kN = kp_.ksize - gw;
    unsigned long long* slots = dMax_ + (size_t)dst * MAX_SLOTS;
    RG_CUDA(cudaMemsetAsync(slots, 0, MAX_SLOTS * sizeof(unsigned long long), stream_));
    phase(PH_COPY, [&] {
      MhdKernels<T>::copyPlanes(kp_, Uold, Unew, 0, gw, stream_);
      MhdKernels<T>::copyPlanes(kp_, Uold, Unew, kN, kp_.ksize, stream_);
    });
    auto runRange = [&](int k0, int k1) {
      if (fused) {
        phase(PH_FUSED, [&] { HydroKernels<T>::fusedStep(kp_, Uold, Unew, k0, k1, dt, slots, stream_); });
        return;
      }
      for (int ka = k0; ka < k1; ka += chunkPlanes_) {
        const int kb = std::min(ka + chunkPlanes_, k1);
        phase(PH_TRACE, [&] { HydroKernels<T>::trace(kp_, Uold, sc_.W, sc_.planes, ka - 1, ka - 1, kb + 1, dt, stream_); });
        phase(PH_UPDATE, [&] {
          HydroKernels<T>::fluxUpdate(kp_, Uold, Unew, sc_.W, sc_.planes, ka - 1, ka, kb, dt, slots, stream_);
        });
      }
    };
    const bool overlap = nranks_ > 1 && overlapHalo_ && (kN - gw) >= 3 * gw && !dissipative();
    if (overlap && p2p_) {  // copy-engine halo after ONE pass over the slab (see stepMhd3d)
      runRange(gw, kN);
      startLateHalo(dst);
    } else if (overlap) {
      runRange(gw, 2 * gw);
      runRange(kN - gw, kN);
      startEarlyHalo(dst);
      runRange(2 * gw, kN - gw);
    } else {
      runRange(gw, kN);
    }
    ghostsValid_[dst] = false;
    dtCached_[dst] = true;
    if (dissipative()) {  // viscosity, reference HydroRunGodunov.cpp:2908-2927
      fillGhosts(dst, 0, kp_.ksize);
      dissipativeTerms(dst, dt);
    }
  }

  ConfigMap cfg_;
  RunParams rp_;
  KParams<T> kp_;
  int rank_, nranks_;
  size_t cells_ = 0, elems_ = 0, deviceBytes_ = 0, scratchBytes_ = 0;
  T* dU_[2] = {nullptr, nullptr};
  // stepsFromHostBatch: second device buffer pair, copy streams, events (created on first use)
  T* batchBuf_[2] = {nullptr, nullptr};
  cudaStream_t h2dStream_ = nullptr, d2hStream_ = nullptr;
  cudaEvent_t evH2D_[2] = {nullptr, nullptr}, evD2H_[2] = {nullptr, nullptr}, evStepDone_[2] = {nullptr, nullptr};
  MhdScratch<T> sc_;
  double lastDt_ = 0.0;  // dt of the last step (restart sidecar, history)
  double resumeDt_ = 0.0;  // next dt read from a restart sidecar
  double* dHist_ = nullptr;  // partial sums of the history kernels
  std::vector<double> hHist_;
  T* dGz_ = nullptr;    // g_z per local plane (stratified shearing box)
  T* dGcell_ = nullptr; // (g_x, g_y) per cell of a 2D hydro run (Keplerian disc)
  T* dDiss_ = nullptr;  // 12-component scratch of the dissipative kernels (allocated on first use)
  int chunkPlanes_ = 0, userChunk_ = 0;
  unsigned long long* dMax_ = nullptr;
  unsigned long long* hMax_ = nullptr;
  bool ghostsValid_[2] = {false, false}, dtCached_[2] = {false, false};
  bool haloDone_[2] = {false, false};  // z halo already exchanged by startEarlyHalo()
  bool overlapHalo_ = true;
  cudaEvent_t evEdge_ = nullptr, evHalo_ = nullptr;
  cudaStream_t stream_ = nullptr, comm_stream_ = nullptr;
  cudaEvent_t ev0_ = nullptr, ev1_ = nullptr, ev_sync_ = nullptr;
  bool timed_ = false;
  bool profiling_ = false;
  cudaEvent_t evProf0_ = nullptr, evProf1_ = nullptr;
  std::vector<Span> spans_;
  std::vector<cudaEvent_t> eventPool_;
  const NcclApi* nccl_ = nullptr;
  NcclApi::comm_t comm_ = nullptr;
  double haloBytesPerStep_ = 0.0;
  // copy-engine halo (initPeerHalo)
  bool p2p_ = false, peerLoOpened_ = false, peerHiOpened_ = false;
  Peer peerLo_, peerHi_;
  T* p2pBase_[2] = {nullptr, nullptr};
  unsigned long long* dFlags_ = nullptr;  // [0..4): flags written by the neighbours, [FL_SLOTS..): local sources of flag values
  unsigned long long haloSeq_ = 0, signalCount_ = 0;
  StreamMemOp waitValue_ = nullptr, writeValue_ = nullptr;
  bool haloOnNccl_ = true;      // the last z halo went through NCCL send/recv (not through the copy engines)
  bool directFlags_ = false;    // flags written by cuStreamWriteValue64 straight into peer memory (probed at start-up)
  size_t maxPitch_ = 0;         // cudaMemcpy2D pitch limit
  std::string lastWarning_;
};

}  // namespace

std::unique_ptr<Run> Run::create(const ConfigMap& cfg, bool fp32, const DistInit& dist) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    throw std::runtime_error("ramsesgpu_b200 needs a CUDA device (sm_100a); there is no CPU fallback");
  // the library ships sm_100a SASS only: any other device would fail every launch
  int dev = dist.device;
  if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  int major = 0;
  if (dev >= count || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess || major != 10)
    throw std::runtime_error("ramsesgpu_b200 is not available on device " + std::to_string(dev) + " (compute capability major " +
                             std::to_string(major) + "): the library holds sm_100a (B200) code only");
  if (fp32) return std::unique_ptr<Run>(new RunImpl<float>(cfg, dist));
  return std::unique_ptr<Run>(new RunImpl<double>(cfg, dist));
}

}  // namespace rg
