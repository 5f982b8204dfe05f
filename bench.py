#!/usr/bin/env python
"""bench.py -- cell-updates/s of the 3D MHD Godunov update path (BASELINE.json metric).

    python bench.py --gpus 1 --steps 20 --warmup 3             # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 4 --warmup 1   # the reference's CPU path

A "step" is one oneStepIntegration (compute_dt + godunov_unsplit) of the Orszag-Tang 3D problem
(data/orszag-tang3d.ini: HLLD fluxes, 2-D HLLD emfs, periodic box, FP64) with 256^3 cells PER GPU
(BASELINE.json configs[1]; weak scaling: the global grid is 256 x 256 x 256*N, z-slab per rank).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mcell-updates/s (FP64) on 3D MHD Godunov"
UNIT = "Mcell-updates/s"
B_ALG_CELL = 128.0        # compulsory bytes per MHD FP64 cell update: read U once, write U once
# algorithmic bytes per launch unit of each kernel family (DESIGN.md "kernels"): reals read + written once
# prim = k_prim (8r+8w) + k_elec (6r+3w); trace: Q 8 + face B 3 + E 3 read, W 38 written; flux: 15 W
# components read + 5 written per direction; emf: 22 read + 1 written per direction; update: U 8 +
# F 15 + E 3 read, U 8 written; fused (flux+emf+update in one kernel): W 38 + U 8 read, U 8 written
B_ALG_KERNEL = {"prim": (8 + 8 + 6 + 3) * 8.0, "trace": (8 + 3 + 3 + 38) * 8.0, "flux": (15 + 5) * 8.0 * 3,
                "emf": (22 + 1) * 8.0 * 3, "update": (8 + 15 + 3 + 8) * 8.0, "fused": (38 + 8 + 8) * 8.0}


def ncu_evidence(fam):
    """Per-launch DRAM traffic and pipe utilisation of a kernel family from the committed ncu capture
    (profiles/ncu_evidence.json, written from `ncu --set full` runs of this same workload)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_evidence.json"))).get(fam)
    except Exception:
        return None


def base_ini():
    z = np.load(os.path.join(ROOT, "tests", "golden", "ot3d_16_s10.npz"))
    return str(z["ini"])


def workload_ini(n, nz_total):
    from ramsesgpu_b200.io import ini_override
    return ini_override(base_ini(), {
        "mesh": {"nx": n, "ny": n, "nz": nz_total},
        "run": {"nstepmax": 1000000, "tend": 1000.0, "noutput": -1},
        "output": {"outputVtk": "no", "outputXsm": "no", "outputHdf5": "no"}})


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "MEASURED_PEAKS.json"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi samples (100 ms) of SM clock and throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.rows, self.proc, self.dev = [], None, device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 9 for i in range(4) if r[5 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# -------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU implementation (oracle/_ref/euler_cpu, built
# from the unmodified sources by oracle/Makefile.ref), one single-thread replica per host core
# (the reference's CMake never enables OpenMP and its OpenMP scatter loop is racy)
# -------------------------------------------------------------------------------------------------
def run_reference_cpu(n, steps, replicas):
    from oracle.oracle import ref_exe
    from ramsesgpu_b200.io import ini_override
    exe = ref_exe("f64")
    kind = "reference"
    ini = ini_override(base_ini(), {"mesh": {"nx": n, "ny": n, "nz": n},
                                    "run": {"nstepmax": steps, "tend": 1000.0, "noutput": 1000000},
                                    "output": {"outputVtk": "no", "outputXsm": "no", "outputHdf5": "no"}})
    t0 = time.time()
    if exe is not None:
        procs = []
        for r in range(replicas):
            wd = tempfile.mkdtemp(prefix="ramses_ref_%d_" % r)
            with open(os.path.join(wd, "run.ini"), "w") as f:
                f.write(ini)
            procs.append(subprocess.Popen([exe, "--param", "run.ini"], cwd=wd, stdout=subprocess.PIPE,
                                          stderr=subprocess.STDOUT, text=True))
        rates = []
        for p in procs:
            out = p.communicate()[0]
            m = re.search(r"([0-9.eE+-]+) cell updates per seconds", out)
            rates.append(float(m.group(1)) if m else 0.0)
        total = sum(rates)  # the reference's own printed figure, summed over the replicas
    else:  # the C restatement of the same algorithm (kind "port"), threads = processes via fork
        kind = "port"
        from concurrent.futures import ProcessPoolExecutor
        with ProcessPoolExecutor(max_workers=replicas) as ex:
            rates = list(ex.map(_oracle_rate, [(ini, steps)] * replicas))
        total = sum(rates)
    wall = time.time() - t0
    return total / 1e6, kind, wall


def _oracle_rate(args):
    ini, steps = args
    from oracle.oracle import Oracle
    o = Oracle("f64")
    p = o.params(ini)
    U = o.init_problem(p)
    t0 = time.time()
    o.run_steps(p, U, steps)
    return steps * p.nx * p.ny * p.nz / (time.time() - t0)


def workload_name(n, nz_local, nz_total):
    return "orszag-tang3d.ini 3D MHD %dx%dx%d per GPU (global nz=%d), HLLD + 2D-HLLD CT, periodic, FP64" % (n, n, nz_local, nz_total)


def reference_arm(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = args.ref_size
    nz_total = args.global_nz if args.global_nz > 0 else args.size * args.gpus
    steps = args.steps + args.warmup
    value, kind, wall = run_reference_cpu(n, steps, cores)
    sample = "%d single-thread replicas of Orszag-Tang 3D %d^3, %d steps each (reference prints nStep*cells/(wall-io))" % (cores, n, steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # the same workload as the native arm; every step is a bounded sample of it (one 64^3 box per host core)
        "config": {"workload": workload_name(args.size, nz_total // max(args.gpus, 1), nz_total),
                   "sample": "reference CPU path euler_cpu, %d^3 cells of the same problem per core, %d cores" % (n, cores)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def measure_e2e(run, args, nstep, cells_all, barrier, max_over_ranks):
    """End to end through the C ABI with HOST buffers: H2D(state) + step + D2H(state) for every step."""
    from ramsesgpu_b200 import PinnedArray
    shape = run.shape
    pins = [PinnedArray(shape) for _ in range(4)]   # page-locked host buffers (rg_alloc_pinned)
    host_in, host_out, host_in2, host_out2 = [p.array for p in pins]
    host_in[...] = run.getDataHost(nstep)
    hin, hout = host_in, host_out
    run.steps_from_host(hin, hout, 1)  # warm-up of the path
    barrier()
    e0 = time.time()
    for _ in range(args.e2e_steps):
        run.steps_from_host(hin, hout, 1)
        hin, hout = hout, hin
    barrier()
    e2e_sync_s = (time.time() - e0) / args.e2e_steps
    # the same work submitted as a batch of independent one-step jobs (rg_steps_from_host_batch): every
    # job still pays its own H2D and D2H inside the timed region, the three engines overlap
    host_in2[...] = host_in
    ins = [host_in, host_in2]
    outs = [host_out, host_out2]
    njobs = max(args.e2e_steps, 2) * 2
    run.steps_from_host_batch(ins, outs)  # warm-up (allocates the second buffer pair)
    barrier()
    e0 = time.time()
    run.steps_from_host_batch([ins[j % 2] for j in range(njobs)], [outs[j % 2] for j in range(njobs)])
    barrier()
    e2e_s = (time.time() - e0) / njobs
    e2e_s, e2e_sync_s = max_over_ranks([e2e_s, e2e_sync_s])
    e2e_value = cells_all / e2e_s / 1e6
    e2e_sync_value = cells_all / e2e_sync_s / 1e6
    return e2e_value, e2e_sync_value, njobs, pins


class Ctx:
    """Process plumbing of one bench run: rank / world, barrier, max over ranks, NCCL unique ids."""

    halo_overlap = True  # --no-halo-overlap: exchange after the whole update instead of early on the communication stream

    def __init__(self, rank, world, local_rank, torch=None, dist=None):
        self.rank, self.world, self.local, self.torch, self.dist = rank, world, local_rank, torch, dist

    def make(self, Run, ini, fp32=False):
        uid = None
        if self.world > 1:
            from ramsesgpu_b200.distcheck import broadcast_unique_id
            uid = broadcast_unique_id(self.torch, self.dist, self.rank)
        run = Run(ini, fp32=fp32, rank=self.rank, nranks=self.world, nccl_unique_id=uid, device=self.local)
        if self.world > 1 and not self.halo_overlap:
            run.set_halo_overlap(False)
        return run

    def barrier(self, run=None):
        if run is not None:
            run.synchronize()
        if self.world > 1:
            self.torch.cuda.synchronize()
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        if self.world == 1:
            return [float(v) for v in values]
        tm = self.torch.tensor(list(values), dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(tm, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in tm.tolist()]


def time_windows(ctx, run, state, steps, windows):
    """`windows` back-to-back measurements of EXACTLY `steps` steps each: barrier + synchronize on both sides,
    CUDA events on the library's stream (rg_profile_begin/end), max over ranks.  Returns the windows' ms (max over
    ranks), the per-family device times and launch count of the MEDIAN window, and the advanced state."""
    per_window, details = [], []
    for _ in range(windows):
        ctx.barrier(run)
        l0 = run.stats().kernel_launches
        run.profile_begin()
        for _ in range(steps):
            state = run.oneStepIntegration(*state)
        total_ms, phases = run.profile_end()
        ctx.barrier(run)
        per_window.append(ctx.max_over_ranks([total_ms])[0])
        details.append((phases, run.stats().kernel_launches - l0, total_ms))
    order = sorted(range(windows), key=lambda w: per_window[w])
    med = order[len(order) // 2]
    return per_window, med, details[med], state


def side_config(ctx, Run, ini, fp32, cells_global, warmup, steps, b_alg, label, peak):
    """One of BASELINE.json's other configurations at full size: device-resident throughput, same timing rules."""
    t0 = time.time()
    run = ctx.make(Run, ini, fp32=fp32)
    try:
        run.init_simulation()
        run.make_all_boundaries(0)
        state = (0, 0.0, 0.0)
        for _ in range(warmup):
            state = run.oneStepIntegration(*state)
        per_window, med, (phases, launches, _), state = time_windows(ctx, run, state, steps, 3)
        ms = per_window[med] / steps
        value = cells_global / (ms * 1e-3) / 1e6
        st = run.stats()
        out = {"workload": label, "value": value, "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": warmup,
               "windows_ms": per_window, "dtype": "f32" if fp32 else "f64", "gpu_launches": int(launches),
               "chunk_planes": st.chunk_planes, "device_gb": st.device_bytes / 1e9,
               "kernels_ms_per_step": {k: v[0] / steps for k, v in phases.items() if v[0] > 0},
               "roofline_step": {"bound": "hbm", "bytes_per_cell": b_alg, "peak": peak, "unit": "GB/s",
                                 "achieved": b_alg * cells_global / ctx.world / (ms * 1e-3) / 1e9,
                                 "frac": b_alg * cells_global / ctx.world / (ms * 1e-3) / 1e9 / peak},
               "leg_wall_s": None}
    finally:
        run.close()
    out["leg_wall_s"] = time.time() - t0
    return out


def host_mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return float(line.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


def device_total_gb():
    import ctypes as C
    try:
        rt = C.CDLL("libcudart.so.12")
        free, total = C.c_size_t(0), C.c_size_t(0)
        if rt.cudaMemGetInfo(C.byref(free), C.byref(total)) == 0:
            return total.value / 1e9
    except Exception:
        pass
    return 0.0


# -------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--size", type=int, default=256, help="cells per direction per GPU")
    ap.add_argument("--ref-size", type=int, default=64, help="grid of each CPU replica of the reference arm")
    ap.add_argument("--windows", type=int, default=5, help="back-to-back timed windows of --steps steps; the median is reported")
    ap.add_argument("--e2e-steps", type=int, default=8,
                    help="host-buffer steps of the e2e leg (the batch leg submits twice as many one-step jobs)")
    ap.add_argument("--global-nz", type=int, default=0,
                    help="headline as strong scaling: fixed global grid size x size x GLOBAL_NZ split into z slabs (default: weak, size^3 per GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the sub-lines of configs 3 and 4 (`configs`)")
    ap.add_argument("--halo", choices=["peer", "nccl"], default="peer",
                    help="z halo at N > 1: copy engines over peer-mapped state arrays (default) or NCCL send/recv (A/B)")
    ap.add_argument("--no-halo-overlap", action="store_true",
                    help="N > 1: one update launch and the halo after it, instead of boundary ranges first + early halo (A/B)")
    ap.add_argument("--no-strong", action="store_true", help="skip the fixed-global-grid strong-scaling legs (`strong`)")
    ap.add_argument("--no-parity", action="store_true", help="skip the N-GPU == 1-GPU bitwise cases (`parity_multi`, N > 1)")
    ap.add_argument("--strong-grid", default="auto", help="auto | NXxNYxNZ of the strong-scaling legs (auto: 1024^3 when one GPU can hold it)")
    ap.add_argument("--strong-steps", type=int, default=3)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank)
        return

    # torch is process plumbing only (rendezvous, barriers, max over ranks): the single-GPU run does not
    # import it at all (which also keeps `ncu python bench.py` light)
    torch = dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from ramsesgpu_b200 import HydroRunGodunov, MHDRunGodunov, _lib
    from ramsesgpu_b200 import build as native_build
    from ramsesgpu_b200.io import ini_override
    if rank == 0:
        native_build.build()
    if world > 1:
        dist.barrier()
    L = _lib.load()
    if L.rg_device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    from ramsesgpu_b200 import set_tuning
    set_tuning("halo_p2p", 0 if args.halo == "nccl" else 1)
    ctx = Ctx(rank, world, local_rank, torch, dist)
    ctx.halo_overlap = not args.no_halo_overlap
    peak, peak_src = measured_peaks()
    quiet = {"run": {"nstepmax": 1000000, "tend": 1.0e9, "noutput": -1},
             "output": {"outputVtk": "no", "outputXsm": "no", "outputHdf5": "no"}}

    def golden_ini(name, mesh, extra=None):
        z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        ov = {k: dict(v) for k, v in quiet.items()}
        ov["mesh"] = dict(mesh)
        for k, v in (extra or {}).items():
            ov.setdefault(k, {}).update(v)
        return ini_override(str(z["ini"]), ov)

    # ---- multi-GPU parity in front of the driver: slabs over NCCL == one GPU, bit for bit ----------
    parity_multi = None
    if world > 1 and not args.no_parity:
        from ramsesgpu_b200.distcheck import slabs_match_single_gpu
        cases = [
            ("orszag-tang3d kt=1 (HLLD + 2D-HLLD, periodic, fused kernels)", MHDRunGodunov,
             golden_ini("ot3d_kt1_16x20x24_s8", {"nx": 64, "ny": 64, "nz": 16 * world}), False, 5),
            ("mhd_mri_3d (shearing box, isothermal, rotating frame)", MHDRunGodunov,
             golden_ini("mri3d_16x32x16_s12", {"nx": 16, "ny": 32, "nz": 16 * world}), False, 5),
            ("implode3d_mpi_zslab (hydro, Dirichlet walls)", HydroRunGodunov,
             golden_ini("implode3d_16_s8", {"nx": 32, "ny": 24, "nz": 16 * world}), False, 5),
            ("kelvin_helmholtz_gpu_3d (hydro FP32, HLLC, rand() perturbation)", HydroRunGodunov,
             golden_ini("kh3d_16x8x16_f32_s10", {"nx": 32, "ny": 16, "nz": 16 * world}), True, 5),
        ]
        results, all_ok = [], True
        for label, Run, ini_c, fp32, nst in cases:
            ok, info = slabs_match_single_gpu(torch, dist, Run, ini_c, nst, rank, world, local_rank, fp32=fp32)
            info["case"] = label
            results.append(info)
            all_ok = all_ok and ok
        parity_multi = {"identical": all_ok, "ranks": world, "cases": results}

    n = args.size
    nz_total = args.global_nz if args.global_nz > 0 else n * world
    if nz_total % world:
        raise SystemExit("bench.py: --global-nz must be a multiple of the number of GPUs")
    nz_local = nz_total // world
    ini = workload_ini(n, nz_total)
    run = ctx.make(MHDRunGodunov, ini)

    def barrier():
        ctx.barrier(run)

    # ---- device-resident throughput -------------------------------------------------------------
    run.init_simulation()
    run.make_all_boundaries(0)
    state = (0, 0.0, 0.0)
    for _ in range(args.warmup):
        state = run.oneStepIntegration(*state)
    run.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()          # before the barrier: the other ranks do not wait for rank 0's fork inside the timed region
    barrier()
    wall0 = time.time()
    per_window, med, (phases, launches, _), state = time_windows(ctx, run, state, args.steps, max(args.windows, 1))
    wall = time.time() - wall0
    clocks = sampler.stop() if rank == 0 else None
    nstep = state[0]
    total_ms = per_window[med]
    cells_per_gpu = float(n) * n * nz_local
    cells_all = cells_per_gpu * world
    ms_per_step = total_ms / args.steps
    value = cells_all / (ms_per_step * 1e-3) / 1e6

    state_bytes = int(np.prod(run.shape)) * 8
    e2e_value = e2e_sync_value = None
    njobs = 0
    pins = []
    if args.e2e_steps > 0:
        e2e_value, e2e_sync_value, njobs, pins = measure_e2e(run, args, nstep, cells_all, barrier, ctx.max_over_ranks)
    chunk_planes = run.stats().chunk_planes
    halo_peer = bool(run.stats().halo_peer_copies)
    device_gb = run.stats().device_bytes / 1e9
    for p_ in pins:
        p_.free()
    run.close()

    line = None
    if rank == 0:
        # dominant kernel family by measured device time (CUDA events around every launch of the median window)
        fam = max(B_ALG_KERNEL, key=lambda k: phases[k][0])
        fam_ms, fam_launches = phases[fam]
        launch_ms = fam_ms / max(fam_launches, 1)
        cells_per_launch = cells_per_gpu * args.steps / max(fam_launches, 1)
        # ALGORITHMIC bytes = 128 B per cell update (read U once, write U once) x the cells one launch
        # processes, over that kernel's average launch duration
        achieved = B_ALG_CELL * cells_per_launch / (launch_ms * 1e-3) / 1e9
        step_achieved = B_ALG_CELL * cells_per_gpu / (ms_per_step * 1e-3) / 1e9
        ev = ncu_evidence(fam)
        # the committed ncu capture describes ONE launch shape (one GPU, 256^3, one launch per step): its DRAM bytes per
        # launch are only attached when this run's launches have that shape
        same_shape = bool(ev) and world == 1 and ev.get("cells_per_launch") == cells_per_launch
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if args.global_nz > 0 else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n, nz_local, nz_total),
                       "parallelism": "z-slab x%d" % world, "cache": "inputs larger than L2 (state %.2f GB per GPU)" % (state_bytes / 1e9),
                       "chunk_planes": chunk_planes, "device_gb": device_gb,
                       "halo": ("none (one GPU)" if world == 1 else
                                "copy engines over peer-mapped state arrays + stream memory operations (no SM, no NCCL kernel)"
                                if halo_peer else "NCCL send/recv") + ("" if world == 1 or ctx.halo_overlap else ", not overlapped"),
                       "timing": "%d back-to-back windows of exactly %d steps, each bracketed by barrier + synchronize, CUDA events, "
                                 "max over ranks; ms_per_step / value are the MEDIAN window" % (len(per_window), args.steps)},
            "windows_ms": per_window,
            "e2e": None if e2e_value is None else {
                "value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": state_bytes, "d2h_bytes_per_step": state_bytes,
                "call": "rg_steps_from_host_batch: %d independent one-step jobs, pinned host buffers, H2D(j+1) | step(j) | D2H(j-1) overlapped" % njobs,
                "single_call_value": e2e_sync_value,
                "single_call": "rg_steps_from_host: H2D, one step, D2H back to back (PCIe-bound: 2 x %.2f GB per step)" % (state_bytes / 1e9),
                "note": "whole-state round trip per step: bound by host memory / PCIe (2 x state bytes per job), not by the kernels"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": fam, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "bytes_per_cell": B_ALG_CELL, "cells_per_launch": cells_per_launch,
                         "traffic": ev.get("dram_bytes_per_launch") if same_shape else None, "peak_source": peak_src,
                         "launch_ms": launch_ms, "share_of_step": fam_ms / total_ms,
                         # what this design moves through HBM for that kernel (W scratch included), NOT the roofline numerator
                         "kernel_traffic_model": {"bytes_per_cell": B_ALG_KERNEL[fam],
                                                  "gbs": B_ALG_KERNEL[fam] * cells_per_launch / (launch_ms * 1e-3) / 1e9},
                         "ncu": ev if same_shape else None,
                         "note": "FP64-pipe / issue bound (see DESIGN.md 4): the compulsory-traffic fraction is small by construction of the metric"},
            "roofline_step": {"bound": "hbm", "achieved": step_achieved, "peak": peak, "unit": "GB/s",
                              "frac": step_achieved / peak, "bytes_per_cell": B_ALG_CELL,
                              "note": "whole step at 128 B/cell"},
            "kernels_ms_per_step": {k: v[0] / args.steps for k, v in phases.items()},
            "wall_ms_per_step": 1e3 * wall / (args.steps * len(per_window)),
        }
        if parity_multi is not None:
            line["parity_multi"] = parity_multi

    # ---- the other configurations of BASELINE.json at full size (sub-lines) -----------------------------------
    def guarded(fn):
        try:
            return fn()
        except Exception as e:  # a sub-line must never take the headline down
            return {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}

    configs = {}
    if not args.no_configs:
        if world == 1:  # configs[2]: one B200
            configs["config3_kh512_fp32_hllc"] = guarded(lambda: side_config(
                ctx, HydroRunGodunov, golden_ini("kh3d_16x8x16_f32_s10", {"nx": 512, "ny": 512, "nz": 512}), True,
                512.0 ** 3, 3, 5, 40.0, "kelvin_helmholtz_gpu_3d.ini 3D hydro 512^3 FP32, HLLC, periodic, one GPU", peak))
        nzc = 64 * world  # configs[3]: 256 x 512 x 256 over 4 GPUs = 64 planes per GPU; the same slab at any N
        configs["config4_mri_slab"] = guarded(lambda: side_config(
            ctx, MHDRunGodunov, golden_ini("mri3d_16x32x16_s12", {"nx": 256, "ny": 512, "nz": nzc}), False,
            256.0 * 512 * nzc, 3, 5, 128.0,
            "mhd_mri_3d.ini shearing box 256x512x%d FP64 (64 planes per GPU, z-slab x%d%s)" % (nzc, world, ", = configs[3]" if world == 4 else ""), peak))
    strong = None
    if not args.no_strong:
        if args.strong_grid == "auto":
            # 1024^3 FP64 MHD: 2 x 69.9 GB of state on ONE GPU at N = 1 (+ W chunks) and one global array on the host(s)
            big = device_total_gb() >= 170.0 and host_mem_available_gb() >= 160.0
            sg = (1024, 1024, 1024) if big else (512, 512, 1024)
        else:
            sg = tuple(int(v) for v in args.strong_grid.lower().split("x"))
        if ctx.world > 1:  # one decision for all ranks
            t = torch.tensor(list(sg), dtype=torch.int64, device="cuda")
            dist.broadcast(t, 0)
            sg = tuple(int(v) for v in t.tolist())
        if sg[2] % world == 0:
            gname = "%dx%dx%d" % sg
            cells = float(sg[0]) * sg[1] * sg[2]
            strong = {"grid": gname, "note": "fixed GLOBAL grid split into z slabs: speed-up over N = value(N) / value(1)"}
            strong["mhd_ot3d"] = guarded(lambda: side_config(
                ctx, MHDRunGodunov, workload_ini3(sg), False, cells, 2, args.strong_steps, 128.0,
                "orszag-tang3d.ini 3D MHD %s global FP64, HLLD + 2D-HLLD CT, z-slab x%d (the MHD sweep of configs[4])" % (gname, world), peak))
            strong["hydro_implode"] = guarded(lambda: side_config(
                ctx, HydroRunGodunov, golden_ini("implode3d_16_s8", {"nx": sg[0], "ny": sg[1], "nz": sg[2]}), False, cells, 2,
                args.strong_steps, 80.0,
                "implode3d_mpi_zslab.ini 3D hydro %s global FP64, approx Riemann, Dirichlet walls, z-slab x%d (configs[4] as shipped)" % (gname, world), peak))
    if rank == 0:
        if configs:
            line["configs"] = configs
        if strong is not None:
            line["strong"] = strong
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            cv, kind, cwall = run_reference_cpu(args.ref_size, 16, cores)
            line["cpu_baseline"] = {"value": cv, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": "%d single-thread replicas of the same problem at %d^3, 16 steps (%.0f s)" % (cores, args.ref_size, cwall)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if parity_multi is not None and not parity_multi["identical"]:
        sys.exit(3)


def workload_ini3(grid):
    from ramsesgpu_b200.io import ini_override
    return ini_override(base_ini(), {
        "mesh": {"nx": grid[0], "ny": grid[1], "nz": grid[2]},
        "run": {"nstepmax": 1000000, "tend": 1000.0, "noutput": -1},
        "output": {"outputVtk": "no", "outputXsm": "no", "outputHdf5": "no"}})


if __name__ == "__main__":
    main()
