"""CPU, world_size 2 over gloo: the z-slab decomposition protocol of the multi-GPU path
(slab extents from rg_slab_extent, local x/y ghost fill, z-halo exchange of gw planes with periodic
wrap, dt = global max of the inverse dt) reproduces the mono-domain result BIT FOR BIT when every
rank advances its slab with the oracle.  This is the host logic of run.cu::fillGhosts/exchangeZ/
compute_dt, exercised without a GPU; one case follows the order of the copy-engine halo (startLateHalo)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ot3d_ini
from ramsesgpu_b200.io import ini_override

NSTEPS = 3
N = (12, 10, 16)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


DISS = {"hydro": {"nu": 0.004}, "MHD": {"eta": 0.003}}


def _worker(rank, world, port, out_dir, diss=False, refresh=True, late=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import Oracle
    from ramsesgpu_b200 import slab_extent
    o = Oracle("f64")
    ini = ot3d_ini(N, OrszagTang={"kt": 1.0}, **(DISS if diss else {}))
    o.set_skip_dissipative(diss)                 # the dissipative block is driven by the slab protocol below
    pg = o.params(ini)
    Ug = o.init_problem(pg)                      # every rank builds the GLOBAL initial state ...
    gw = pg.ghostWidth
    nzl, koff = slab_extent(pg.nz, world, rank)
    # ... and keeps its slab (ghost planes included)
    U = np.ascontiguousarray(Ug[:, koff:koff + nzl + 2 * gw])
    # local parameters: same dz, nz = slab thickness (zmax chosen so that (zmax-zmin)/nz == dz exactly)
    pl = o.params(ini_override(ini, {"mesh": {"nz": nzl, "zmin": 0.0, "zmax": nzl / pg.nz}}))
    assert pl.dz == pg.dz and pl.ksize == nzl + 2 * gw
    up, down = (rank + 1) % world, (rank - 1) % world

    def fill_ghosts(A):
        o.make_boundaries(pl, A, 1)
        o.make_boundaries(pl, A, 2)
        exchange_z(A)

    def exchange_z(A):
        top = torch.from_numpy(np.ascontiguousarray(A[:, nzl:nzl + gw]))      # my top inner planes
        bot = torch.from_numpy(np.ascontiguousarray(A[:, gw:2 * gw]))        # my bottom inner planes
        from_below, from_above = torch.empty_like(top), torch.empty_like(bot)
        reqs = [dist.isend(top, up, tag=1), dist.isend(bot, down, tag=2),
                dist.irecv(from_below, down, tag=1), dist.irecv(from_above, up, tag=2)]
        for r in reqs:
            r.wait()
        A[:, :gw] = from_below.numpy()
        A[:, nzl + gw:] = from_above.numpy()

    def refresh_interior_B(A):
        """B of the ghost planes next to INTERIOR slab interfaces, after the resistive CT update: in the
        mono-domain run those planes are inner cells and carry the updated field, while the ghosts of
        the global (here periodic) boundary keep their pre-CT values (run.cu::exchangeZInterior)."""
        reqs, bufs = [], {}
        if rank < world - 1:
            top = torch.from_numpy(np.ascontiguousarray(A[5:8, nzl:nzl + gw]))
            bufs["above"] = torch.empty_like(top)
            reqs += [dist.isend(top, rank + 1, tag=3), dist.irecv(bufs["above"], rank + 1, tag=4)]
        if rank > 0:
            bot = torch.from_numpy(np.ascontiguousarray(A[5:8, gw:2 * gw]))
            bufs["below"] = torch.empty_like(bot)
            reqs += [dist.isend(bot, rank - 1, tag=4), dist.irecv(bufs["below"], rank - 1, tag=3)]
        for r in reqs:
            r.wait()
        if "above" in bufs:
            A[5:8, nzl + gw:] = bufs["above"].numpy()
        if "below" in bufs:
            A[5:8, :gw] = bufs["below"].numpy()

    fill_ghosts(U)
    U2 = U.copy()
    a, b = U, U2
    for _ in range(NSTEPS):
        mn = torch.tensor([o.compute_dt(pl, a)], dtype=torch.float64)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)   # reference: allReduce(MIN) of dt, HydroRunBaseMpi.cpp:700
        dt = float(mn.item())
        if late:   # the z-ghost planes arrived raw at the end of the previous step: the x/y fill over ALL planes completes them
            o.make_boundaries(pl, a, 1)
            o.make_boundaries(pl, a, 2)
        else:
            fill_ghosts(a)
        o.step_no_boundaries(pl, a, b, dt)
        if late:   # run.cu::startLateHalo (copy-engine halo): the boundary planes of the NEW state leave as they are
            exchange_z(b)
        if diss:                                   # run.cu::stepMhd3d: ghost refresh, then the dissipative terms
            fill_ghosts(b)
            o.dissipative_stage(pl, b, dt, 0)
            if refresh:
                refresh_interior_B(b)
            o.dissipative_stage(pl, b, dt, 1)
        a, b = b, a
    np.save(os.path.join(out_dir, "slab%d.npy" % rank), a[:, gw:gw + nzl])
    dist.destroy_process_group()


def _run(tmp_path, oracle64, diss, refresh=True, late=False):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path), diss, refresh, late), nprocs=world, join=True)
    ini = ot3d_ini(N, OrszagTang={"kt": 1.0}, **(DISS if diss else {}))
    p = oracle64.params(ini)
    Uf, _, _ = oracle64.run_steps(p, oracle64.init_problem(p), NSTEPS)
    gw = p.ghostWidth
    got = np.concatenate([np.load(tmp_path / ("slab%d.npy" % r)) for r in range(world)], axis=1)
    return got[:, :, gw:-gw, gw:-gw], Uf[:, gw:-gw, gw:-gw, gw:-gw]


def test_two_slabs_equal_mono_domain(tmp_path, oracle64):
    got, want = _run(tmp_path, oracle64, diss=False)
    assert np.array_equal(got, want)


def test_two_slabs_raw_planes_sent_after_the_step_equal_mono_domain(tmp_path, oracle64):
    """the order of the copy-engine halo (run.cu::startLateHalo): the boundary planes of the new state are exchanged
    right after the step WITHOUT their x/y ghosts; the receiver's x/y fill over all planes, ghost planes included, gives
    them the same x/y ghosts the sender's fill would have (the fills act plane by plane)"""
    got, want = _run(tmp_path, oracle64, diss=False, late=True)
    assert np.array_equal(got, want)


def test_two_slabs_with_dissipative_terms_equal_mono_domain(tmp_path, oracle64):
    """resistivity + viscosity: a second ghost refresh inside the step, and the field of the interior
    interfaces refreshed once more between the resistive CT update and the resistive energy flux"""
    got, want = _run(tmp_path, oracle64, diss=True)
    assert np.array_equal(got, want)


def test_interior_field_refresh_is_needed(tmp_path, oracle64):
    """without that refresh the energy next to the slab interface differs from the mono-domain run"""
    got, want = _run(tmp_path, oracle64, diss=True, refresh=False)
    assert not np.array_equal(got[1], want[1])
