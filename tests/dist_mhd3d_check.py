"""Launched by tests/test_gpu_multi.py under torchrun (one rank per GPU): runs the z-slab decomposed
3D MHD path over NCCL and checks that the gathered result is BITWISE identical to the single-GPU
run of the same global problem (every cell sees the same inputs whatever the decomposition)."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from conftest import load_golden, ot3d_ini
    from ramsesgpu_b200 import HydroRunGodunov, MHDRunGodunov, _lib
    from ramsesgpu_b200.io import ini_override
    L = _lib.load()
    nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    nz = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    periodic_z = (sys.argv[3] == "periodic") if len(sys.argv) > 3 else True
    overlap = (sys.argv[4] != "nooverlap") if len(sys.argv) > 4 else True
    problem = sys.argv[5] if len(sys.argv) > 5 else "ot3d"
    mesh = {} if periodic_z else {"boundary_zmin": 2, "boundary_zmax": 1}
    fp32 = False
    Run = MHDRunGodunov
    if problem == "ot3d":
        ini = ot3d_ini((20, 16, nz), OrszagTang={"kt": 1.0}, mesh=mesh)
    elif problem == "mri":      # BASELINE.json configs[3]: shearing box, z-slabs
        ini = ini_override(str(load_golden("mri3d_12x20x8_s40")["ini"]), {"mesh": {"nx": 12, "ny": 20, "nz": nz}})
    elif problem == "ot3d_diss":  # resistivity + viscosity: ghost refresh (z halo) of the new state inside the step
        ini = ot3d_ini((20, 16, nz), OrszagTang={"kt": 1.0}, mesh=mesh, hydro={"nu": 0.004}, MHD={"eta": 0.003})
    elif problem == "mri_diss":
        ini = ini_override(str(load_golden("mri3d_diss_12x20x8_s10")["ini"]), {"mesh": {"nx": 12, "ny": 20, "nz": nz}})
    elif problem == "rt_mhd":     # static gravity, rand() stream over every cell of the global array (ghosts included)
        ini = ini_override(str(load_golden("rt3d_mhd_visc_rand_8x10x16_s5")["ini"]), {"mesh": {"nx": 8, "ny": 10, "nz": nz}})
    elif problem == "implode":  # configs[4]: hydro, Dirichlet walls (physical z faces on the outer slabs)
        ini = ini_override(str(load_golden("implode3d_16_s8")["ini"]), {"mesh": {"nx": 16, "ny": 12, "nz": nz}})
        Run = HydroRunGodunov
    else:                       # configs[2]: FP32 Kelvin-Helmholtz, rand() perturbation stream
        ini = ini_override(str(load_golden("kh3d_16x8x16_f32_s10")["ini"]), {"mesh": {"nx": 16, "ny": 8, "nz": nz}})
        Run = HydroRunGodunov
        fp32 = True
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        raw = C.create_string_buffer(128)
        _lib.check(L.rg_nccl_unique_id(raw))
        buf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    uid = bytes(buf.cpu().numpy().tobytes())

    def run_steps(run):
        run.init_simulation()
        run.make_all_boundaries(0)
        run.setDataHost(run.getDataHost(0), 1)
        n, t, dt, dts = 0, 0.0, 0.0, []
        for _ in range(nsteps):
            n, t, dt = run.oneStepIntegration(n, t, dt)
            dts.append(dt)
        return run.getDataHost(n), dts

    with Run(ini, fp32=fp32, rank=rank, nranks=world, nccl_unique_id=uid, device=local) as run:
        run.set_halo_overlap(overlap)
        U, dts = run_steps(run)
        g, nzl, koff = run.layout.ghost_width, run.layout.nz_local, run.layout.k_offset
        halo = run.stats().halo_bytes_per_step
    inner = torch.from_numpy(np.ascontiguousarray(U[:, g:g + nzl, g:-g, g:-g])).cuda()
    sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([nzl], dtype=torch.int64, device="cuda"))
    ok = True
    if rank == 0:
        parts = [inner.cpu().numpy()]
        for r in range(1, world):
            t = torch.empty((inner.shape[0], int(sizes[r].item()), inner.shape[2], inner.shape[3]), dtype=inner.dtype, device="cuda")
            dist.recv(t, r)
            parts.append(t.cpu().numpy())
        got = np.concatenate(parts, axis=1)
        with Run(ini, fp32=fp32) as mono:
            Um, dtm = run_steps(mono)
        want = Um[:, g:-g, g:-g, g:-g]
        ok = bool(np.array_equal(got, want)) and dts == dtm
        print("dist check: problem=" + problem + " world=%d nz=%d steps=%d periodic_z=%s overlap=%s halo_bytes=%d identical=%s maxdiff=%.3e" %
              (world, nz, nsteps, periodic_z, overlap, halo, ok, float(np.abs(got - want).max())), flush=True)
    else:
        dist.send(inner, 0)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
