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
    halo = sys.argv[6] if len(sys.argv) > 6 else "peer"   # "peer": copy engines over peer-mapped arrays (default), "nccl": send/recv
    from ramsesgpu_b200 import set_tuning
    set_tuning("halo_p2p", 0 if halo == "nccl" else 1)
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
    from ramsesgpu_b200.distcheck import slabs_match_single_gpu
    ok, info = slabs_match_single_gpu(torch, dist, Run, ini, nsteps, rank, world, local, fp32=fp32, overlap=overlap)
    if rank == 0:
        print("dist check: problem=" + problem + " world=%d nz=%d steps=%d periodic_z=%s overlap=%s halo_bytes=%d peer_copies=%s identical=%s maxdiff=%.3e" %
              (world, nz, nsteps, periodic_z, overlap, info["halo_bytes_per_step"], info["halo_peer_copies"], ok, info["max_abs_diff"]), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
