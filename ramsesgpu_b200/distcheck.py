"""Multi-GPU parity check shared by bench.py (`parity_multi`) and tests/dist_mhd3d_check.py: the z-slab decomposed
run over NCCL (one rank per GPU, torch.distributed for the process plumbing) must reproduce the single-GPU run of
the same global problem BIT FOR BIT -- every cell sees the same inputs whatever the decomposition (reference halo
semantics: HydroRunBaseMpi.cpp:3294-3389; the reference's own MPI build is not decomposition independent because it
reseeds its pseudo-random streams per rank, HydroRunBaseMpi.cpp:10009-10015)."""
import ctypes as C

import numpy as np


def broadcast_unique_id(torch, dist, rank):
    """128-byte NCCL unique id of the library's own communicator, created on rank 0 and broadcast."""
    from . import _lib
    L = _lib.load()
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        raw = C.create_string_buffer(128)
        _lib.check(L.rg_nccl_unique_id(raw))
        buf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().numpy().tobytes())


def run_steps(run, nsteps):
    run.init_simulation()
    run.make_all_boundaries(0)
    run.setDataHost(run.getDataHost(0), 1)
    n, t, dt, dts = 0, 0.0, 0.0, []
    for _ in range(nsteps):
        n, t, dt = run.oneStepIntegration(n, t, dt)
        dts.append(dt)
    return run.getDataHost(n), dts


def slabs_match_single_gpu(torch, dist, Run, ini, nsteps, rank, world, local, fp32=False, overlap=True):
    """Runs `nsteps` of `ini` decomposed over `world` ranks, gathers the inner cells on rank 0, runs the same problem
    on rank 0's GPU alone and compares (np.array_equal on the state, == on every dt).  Collective: every rank calls it.
    Returns (identical, info dict) on every rank."""
    uid = broadcast_unique_id(torch, dist, rank)
    with Run(ini, fp32=fp32, rank=rank, nranks=world, nccl_unique_id=uid, device=local) as run:
        run.set_halo_overlap(overlap)
        U, dts = run_steps(run, nsteps)
        g, nzl = run.layout.ghost_width, run.layout.nz_local
        halo = run.stats().halo_bytes_per_step
        peer = bool(run.stats().halo_peer_copies)
        lay = run.layout
        grid = (lay.nx, lay.ny, lay.nz)
    inner = torch.from_numpy(np.ascontiguousarray(U[:, g:g + nzl, g:-g, g:-g])).cuda()
    sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([nzl], dtype=torch.int64, device="cuda"))
    ok, maxdiff = True, 0.0
    if rank == 0:
        parts = [inner.cpu().numpy()]
        for r in range(1, world):
            t = torch.empty((inner.shape[0], int(sizes[r].item()), inner.shape[2], inner.shape[3]), dtype=inner.dtype, device="cuda")
            dist.recv(t, r)
            parts.append(t.cpu().numpy())
        got = np.concatenate(parts, axis=1)
        with Run(ini, fp32=fp32) as mono:
            Um, dtm = run_steps(mono, nsteps)
        want = Um[:, g:-g, g:-g, g:-g]
        ok = bool(np.array_equal(got, want)) and dts == dtm
        maxdiff = float(np.abs(got - want).max())
    else:
        dist.send(inner, 0)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    ok = bool(flag.item() == 1)
    return ok, {"grid": "%dx%dx%d" % grid, "steps": nsteps, "ranks": world, "identical": ok, "max_abs_diff": maxdiff,
                "halo_bytes_per_step": halo, "overlap": bool(overlap), "halo_peer_copies": peer}
