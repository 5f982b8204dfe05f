// extern "C" boundary: every entry point of include/ramsesgpu_b200.h; no exception crosses it.
#include <cuda_runtime.h>

#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ramsesgpu_b200.h"
#include "init_conditions.h"
#include "kernels.h"
#include "nccl_dyn.h"
#include "run.h"

struct rg_run_s {
  std::unique_ptr<rg::Run> run;
};

namespace {
thread_local std::string g_error;

int fail(int code, const std::string& msg) {
  g_error = msg;
  return code;
}
int classify(const std::exception& e) {
  const std::string m = e.what();
  if (m.find("no CPU fallback") != std::string::npos) return RG_ERR_NO_DEVICE;
  if (m.find("CUDA error") != std::string::npos || m.find("CUDA kernel launch") != std::string::npos) return RG_ERR_CUDA;
  if (m.find("NCCL") != std::string::npos) return RG_ERR_NCCL;
  if (m.find("not available") != std::string::npos) return RG_ERR_UNSUPPORTED;
  return RG_ERR_INVALID;
}
}  // namespace

#define RG_TRY(h, body)                                   \
  if (!(h) || !(h)->run) return fail(RG_ERR_INVALID, "null handle"); \
  try {                                                   \
    body;                                                 \
    return RG_OK;                                         \
  } catch (const std::exception& e) {                     \
    return fail(classify(e), e.what());                   \
  }

// Host-only initial condition of z-slab `rank` of `nranks` (the host half of init_simulation: no device is touched).
template <typename T>
static int initialConditionHost(const rg::ConfigMap& cfg, int rank, int nranks, void* dst, size_t bytes, rg_layout* lay) {
  const rg::RunParams rp = rg::parseRunParams(cfg);
  int nzLocal = rp.nz, kOff = 0;
  if (rp.dim != 3 && nranks > 1) return fail(RG_ERR_INVALID, "z-slab decomposition needs a 3D run");
  if (rp.dim == 3) rg::slabExtent(rp.nz, nranks, rank, &nzLocal, &kOff);
  const rg::KParams<T> kp = rg::makeKParams<T>(cfg, rp, nzLocal, kOff);
  const size_t n = (size_t)kp.isize * kp.jsize * kp.ksize * kp.nvar;
  if (lay) {
    *lay = rg_layout{rp.nx, rp.ny, rp.nz, kp.isize, kp.jsize, kp.ksize, kp.nvar, kp.gw, rp.dim, rp.mhdEnabled ? 1 : 0,
                     (int)sizeof(T), nzLocal, kOff, rank, nranks};
  }
  if (!dst) return RG_OK;  // layout query
  if (bytes != n * sizeof(T)) return fail(RG_ERR_INVALID, "buffer size does not match the local array");
  std::vector<T> U;
  std::string msg;
  const bool ok = rg::initProblem<T>(cfg, rp, kp, rp.problem, U, &msg);
  std::memcpy(dst, U.data(), n * sizeof(T));
  return ok ? RG_OK : fail(RG_ERR_UNSUPPORTED, msg);
}

extern "C" {

const char* rg_last_error(void) { return g_error.c_str(); }
const char* rg_version(void) { return "ramsesgpu_b200 0.1 (sm_100a)"; }

int rg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

static int createImpl(const rg::ConfigMap& cfg, int flags, const rg::DistInit& dist, rg_handle* out) {
  if (!out) return fail(RG_ERR_INVALID, "null output handle");
  *out = nullptr;
  try {
    std::unique_ptr<rg_run_s> h(new rg_run_s);
    h->run = rg::Run::create(cfg, (flags & RG_FLAG_FP32) != 0, dist);
    *out = h.release();
    return RG_OK;
  } catch (const std::exception& e) {
    return fail(classify(e), e.what());
  }
}

int rg_create(const char* ini_text, int flags, rg_handle* out) {
  if (!ini_text) return fail(RG_ERR_INVALID, "null ini text");
  return createImpl(rg::ConfigMap::fromText(ini_text), flags, rg::DistInit(), out);
}

int rg_create_from_file(const char* ini_path, int flags, rg_handle* out) {
  if (!ini_path) return fail(RG_ERR_INVALID, "null path");
  bool ok = false;
  rg::ConfigMap cfg = rg::ConfigMap::fromFile(ini_path, &ok);
  if (!ok) return fail(RG_ERR_IO, std::string("cannot read parameter file ") + ini_path);
  return createImpl(cfg, flags, rg::DistInit(), out);
}

int rg_create_distributed(const char* ini_text, int flags, int rank, int nranks, const void* id, int device,
                          rg_handle* out) {
  if (!ini_text) return fail(RG_ERR_INVALID, "null ini text");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(RG_ERR_INVALID, "bad rank / nranks");
  rg::DistInit d;
  d.rank = rank; d.nranks = nranks; d.ncclUniqueId = id; d.device = device;
  return createImpl(rg::ConfigMap::fromText(ini_text), flags, d, out);
}

int rg_nccl_unique_id(void* out128) {
  if (!out128) return fail(RG_ERR_INVALID, "null buffer");
  const char* err = nullptr;
  const rg::NcclApi* api = rg::NcclApi::get(&err);
  if (!api) return fail(RG_ERR_NCCL, err ? err : "NCCL unavailable");
  rg::NcclApi::UniqueId id;
  int rc = api->GetUniqueId(&id);
  if (rc != 0) return fail(RG_ERR_NCCL, api->GetErrorString(rc));
  std::memcpy(out128, &id, sizeof id);
  return RG_OK;
}

int rg_destroy(rg_handle h) {
  if (!h) return RG_OK;
  try { delete h; } catch (...) {}
  return RG_OK;
}

int rg_get_layout(rg_handle h, rg_layout* out) {
  if (!out) return fail(RG_ERR_INVALID, "null layout");
  RG_TRY(h, {
    rg::Layout l = h->run->layout();
    out->nx = l.nx; out->ny = l.ny; out->nz = l.nz;
    out->isize = l.isize; out->jsize = l.jsize; out->ksize = l.ksize;
    out->nvar = l.nvar; out->ghost_width = l.ghostWidth; out->dim = l.dim; out->mhd = l.mhd;
    out->real_bytes = l.realBytes; out->nz_local = l.nzLocal; out->k_offset = l.kOffset;
    out->rank = l.rank; out->nranks = l.nranks;
  })
}

int rg_get_param(rg_handle h, const char* name, double* value) {
  if (!name || !value) return fail(RG_ERR_INVALID, "null argument");
  RG_TRY(h, {
    bool ok = false;
    *value = h->run->param(name, &ok);
    if (!ok) throw std::runtime_error(std::string("unknown parameter ") + name);
  })
}

int rg_init_simulation(rg_handle h, const char* problem, int* nStep) {
  RG_TRY(h, {
    int n = h->run->init_simulation(problem ? problem : "");
    if (nStep) *nStep = n;
  })
}
int rg_make_all_boundaries(rg_handle h, int which) { RG_TRY(h, h->run->make_all_boundaries(which ? 1 : 0)) }
int rg_compute_dt(rg_handle h, int useU, double* dt) {
  if (!dt) return fail(RG_ERR_INVALID, "null dt");
  RG_TRY(h, *dt = h->run->compute_dt(useU))
}
int rg_godunov_unsplit(rg_handle h, int nStep, double dt) { RG_TRY(h, h->run->godunov_unsplit(nStep, dt)) }
int rg_one_step(rg_handle h, int* nStep, double* t, double* dt) {
  if (!nStep || !t || !dt) return fail(RG_ERR_INVALID, "null argument");
  RG_TRY(h, h->run->oneStepIntegration(*nStep, *t, *dt))
}
int rg_run(rg_handle h) { RG_TRY(h, h->run->start()) }
int rg_output(rg_handle h, int nStep) { RG_TRY(h, h->run->output(nStep)) }

int rg_get_data_device(rg_handle h, int which, void** p) {
  if (!p) return fail(RG_ERR_INVALID, "null pointer");
  RG_TRY(h, *p = h->run->deviceData(which))
}
int rg_copy_to_host(rg_handle h, int which, void* dst, size_t bytes) {
  if (!dst) return fail(RG_ERR_INVALID, "null destination");
  RG_TRY(h, h->run->copyToHost(which, dst, bytes))
}
int rg_copy_from_host(rg_handle h, int which, const void* src, size_t bytes) {
  if (!src) return fail(RG_ERR_INVALID, "null source");
  RG_TRY(h, h->run->copyFromHost(which, src, bytes))
}
int rg_synchronize(rg_handle h) { RG_TRY(h, h->run->synchronize()) }

int rg_steps_from_host(rg_handle h, const void* in, void* outp, size_t bytes, int n, double* t, double* dt) {
  if (!in || !outp) return fail(RG_ERR_INVALID, "null buffer");
  RG_TRY(h, h->run->stepsFromHost(in, outp, bytes, n, t, dt))
}

int rg_steps_from_host_batch(rg_handle h, int n_jobs, const void* const* in, void* const* outp, size_t bytes,
                              double* dt_out) {
  if (n_jobs < 0 || (n_jobs > 0 && (!in || !outp))) return fail(RG_ERR_INVALID, "null buffer list");
  for (int j = 0; j < n_jobs; ++j)
    if (!in[j] || !outp[j]) return fail(RG_ERR_INVALID, "null buffer");
  RG_TRY(h, h->run->stepsFromHostBatch(n_jobs, in, outp, bytes, dt_out))
}

int rg_alloc_pinned(size_t bytes, void** out) {
  if (!out) return fail(RG_ERR_INVALID, "null output pointer");
  *out = nullptr;
  const cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocDefault);
  if (e != cudaSuccess) return fail(RG_ERR_CUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
  return RG_OK;
}
int rg_free_pinned(void* p) {
  if (p && cudaFreeHost(p) != cudaSuccess) return fail(RG_ERR_CUDA, "cudaFreeHost failed");
  return RG_OK;
}

int rg_get_stats(rg_handle h, rg_stats* out) {
  if (!out) return fail(RG_ERR_INVALID, "null stats");
  RG_TRY(h, {
    rg::Stats s = h->run->stats();
    out->kernel_launches = s.kernelLaunches;
    out->last_step_ms = s.lastStepMs;
    out->halo_bytes_per_step = s.haloBytesPerStep;
    out->device_bytes = s.deviceBytes;
    out->chunk_planes = s.chunkPlanes;
    out->halo_peer_copies = s.haloPeerCopies;
  })
}
int rg_reset_launch_count(void) { rg::resetKernelLaunchCount(); return RG_OK; }
int rg_set_chunk_planes(rg_handle h, int planes) { RG_TRY(h, h->run->setChunkPlanes(planes)) }
int rg_set_halo_overlap(rg_handle h, int on) { RG_TRY(h, h->run->setOverlap(on != 0)) }

int rg_set_tuning(const char* key, int value) {
  return rg::setTuning(key, value) ? RG_OK : fail(RG_ERR_INVALID, "unknown tuning key or value out of range");
}

int rg_profile_begin(rg_handle h) { RG_TRY(h, h->run->profileBegin()) }
int rg_history(rg_handle h, int nStep, double* out8) {
  RG_TRY(h, h->run->history(nStep, out8))
}

int rg_profile_end(rg_handle h, double* total, double* phase, unsigned long long* launches) {
  RG_TRY(h, h->run->profileEnd(total, phase, launches))
}

int rg_probe_riemann_mhd(rg_handle h, int n, const void* ql, const void* qr, void* flux) {
  if (!ql || !qr || !flux || n <= 0) return fail(RG_ERR_INVALID, "bad probe arguments");
  RG_TRY(h, h->run->probeRiemann(n, ql, qr, flux))
}
int rg_probe_compute_emf(rg_handle h, int n, int dir, const void* q, const void* x, void* emf) {
  if (!q || !emf || n <= 0 || dir < 0 || dir > 2) return fail(RG_ERR_INVALID, "bad probe arguments");
  RG_TRY(h, h->run->probeEmf(n, dir, q, x, emf))
}

int rg_initial_condition_host(const char* ini_text, int flags, int rank, int nranks, void* dst, size_t bytes,
                              rg_layout* layout_out) {
  if (!ini_text) return fail(RG_ERR_INVALID, "null ini text");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(RG_ERR_INVALID, "bad rank / nranks");
  try {
    const rg::ConfigMap cfg = rg::ConfigMap::fromText(ini_text);
    return (flags & RG_FLAG_FP32) ? initialConditionHost<float>(cfg, rank, nranks, dst, bytes, layout_out)
                                  : initialConditionHost<double>(cfg, rank, nranks, dst, bytes, layout_out);
  } catch (const std::exception& e) {
    return fail(classify(e), e.what());
  }
}

int rg_slab_extent(int nzGlobal, int nranks, int rank, int* nzLocal, int* kOffset) {
  if (!nzLocal || !kOffset || nranks < 1 || rank < 0 || rank >= nranks) return fail(RG_ERR_INVALID, "bad slab arguments");
  rg::slabExtent(nzGlobal, nranks, rank, nzLocal, kOffset);
  return RG_OK;
}

}  // extern "C"
