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
    """nvidia-smi samples (200 ms) of SM clock and throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.rows, self.proc, self.dev = [], None, device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
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


# -------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--size", type=int, default=256, help="cells per direction per GPU")
    ap.add_argument("--ref-size", type=int, default=64, help="grid of each CPU replica of the reference arm")
    ap.add_argument("--e2e-steps", type=int, default=8,
                    help="host-buffer steps of the e2e leg (the batch leg submits twice as many one-step jobs)")
    ap.add_argument("--global-nz", type=int, default=0,
                    help="strong scaling: fixed global grid size x size x GLOBAL_NZ split into z slabs (default: weak, size^3 per GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    from ramsesgpu_b200 import MHDRunGodunov, PinnedArray, _lib
    from ramsesgpu_b200 import build as native_build
    if rank == 0:
        native_build.build()
    if world > 1:
        dist.barrier()
    L = _lib.load()
    if L.rg_device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")

    n = args.size
    nz_total = args.global_nz if args.global_nz > 0 else n * world
    if nz_total % world:
        raise SystemExit("bench.py: --global-nz must be a multiple of the number of GPUs")
    nz_local = nz_total // world
    ini = workload_ini(n, nz_total)
    uid = None
    if world > 1:
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            import ctypes as C
            raw = C.create_string_buffer(128)
            _lib.check(L.rg_nccl_unique_id(raw))
            buf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        uid = bytes(buf.cpu().numpy().tobytes())
    run = MHDRunGodunov(ini, rank=rank, nranks=world, nccl_unique_id=uid, device=local_rank)

    def barrier():
        run.synchronize()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(values):
        if world == 1:
            return [float(v) for v in values]
        tm = torch.tensor(list(values), dtype=torch.float64, device="cuda")
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        return [float(v) for v in tm.tolist()]

    # ---- device-resident throughput -------------------------------------------------------------
    run.init_simulation()
    run.make_all_boundaries(0)
    nstep, t, dt = 0, 0.0, 0.0
    for _ in range(args.warmup):
        nstep, t, dt = run.oneStepIntegration(nstep, t, dt)
    run.synchronize()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = run.stats().kernel_launches
    run.profile_begin()
    wall0 = time.time()
    for _ in range(args.steps):
        nstep, t, dt = run.oneStepIntegration(nstep, t, dt)
    total_ms, phases = run.profile_end()
    barrier()
    wall = time.time() - wall0
    clocks = sampler.stop() if rank == 0 else None
    launches = run.stats().kernel_launches - launches0
    total_ms = max_over_ranks([total_ms])[0]
    cells_per_gpu = float(n) * n * nz_local
    cells_all = cells_per_gpu * world
    ms_per_step = total_ms / args.steps
    value = cells_all / (ms_per_step * 1e-3) / 1e6

    state_bytes = int(np.prod(run.shape)) * 8
    e2e_value = e2e_sync_value = None
    njobs = 0
    pins = []
    if args.e2e_steps > 0:
        e2e_value, e2e_sync_value, njobs, pins = measure_e2e(run, args, nstep, cells_all, barrier, max_over_ranks)

    if rank == 0:
        peak, peak_src = measured_peaks()
        # dominant kernel family by measured device time
        fam = max(B_ALG_KERNEL, key=lambda k: phases[k][0])
        fam_ms, fam_launches = phases[fam]
        units = cells_per_gpu * args.steps            # cell updates processed by that family in the region
        achieved = B_ALG_KERNEL[fam] * units / (fam_ms * 1e-3) / 1e9
        step_achieved = B_ALG_CELL * cells_per_gpu / (ms_per_step * 1e-3) / 1e9
        ev = ncu_evidence(fam)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if args.global_nz > 0 else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n, nz_local, nz_total),
                       "parallelism": "z-slab x%d" % world, "cache": "inputs larger than L2 (state %.2f GB per GPU)" % (state_bytes / 1e9),
                       "chunk_planes": run.stats().chunk_planes},
            "e2e": None if e2e_value is None else {
                "value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": state_bytes, "d2h_bytes_per_step": state_bytes,
                "call": "rg_steps_from_host_batch: %d independent one-step jobs, pinned host buffers, H2D(j+1) | step(j) | D2H(j-1) overlapped" % njobs,
                "single_call_value": e2e_sync_value,
                "single_call": "rg_steps_from_host: H2D, one step, D2H back to back (PCIe-bound: 2 x %.2f GB per step)" % (state_bytes / 1e9)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": fam, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": (ev or {}).get("dram_bytes_per_launch"), "peak_source": peak_src,
                         "launch_ms": fam_ms / max(fam_launches, 1), "share_of_step": fam_ms / total_ms,
                         "ncu": ev},
            "roofline_step": {"bound": "hbm", "achieved": step_achieved, "peak": peak, "unit": "GB/s",
                              "frac": step_achieved / peak, "bytes_per_cell": B_ALG_CELL,
                              "note": "whole fused-equivalent step at 128 B/cell; FP64-pipe bound, see DESIGN.md"},
            "kernels_ms_per_step": {k: v[0] / args.steps for k, v in phases.items()},
            "wall_ms_per_step": 1e3 * wall / args.steps,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            cv, kind, cwall = run_reference_cpu(args.ref_size, 16, cores)
            line["cpu_baseline"] = {"value": cv, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": "%d single-thread replicas of the same problem at %d^3, 16 steps (%.0f s)" % (cores, args.ref_size, cwall)}
        print(json.dumps(line))
    for p_ in pins:
        p_.free()
    run.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
