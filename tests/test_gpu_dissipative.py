"""GPU: resistivity, viscosity and static gravity of the CUDA path, through the C ABI,
against golden vectors from the unmodified reference executable and against the C oracle
(whole arrays, ghosts included; chunked pipeline)."""
import numpy as np
import pytest

from conftest import TOL_F64, load_golden
from ramsesgpu_b200.io import ini_override

pytestmark = pytest.mark.gpu

TOL_F32 = 2e-5   # reference float build against our float kernels (see test_gpu_hydro3d.py)

MHD_CASES = ["ot3d_diss_16x12x20_s6", "ot3d_eta_walls_16_s4", "mri3d_diss_12x20x8_s10", "rt3d_mhd_10x8x24_s8",
             "rt3d_mhd_visc_rand_8x10x16_s5"]
HYDRO_CASES = [("implode3d_visc_16_s6", TOL_F64), ("kh3d_visc_16x8x16_f32_s6", TOL_F32), ("rt3d_hydro_10x8x24_s8", TOL_F64)]


def run_gpu(ini, nsteps, mhd, fp32=False, chunk=0):
    from ramsesgpu_b200 import HydroRunGodunov, MHDRunGodunov
    with (MHDRunGodunov(ini) if mhd else HydroRunGodunov(ini, fp32=fp32)) as run:
        if chunk:
            run.set_chunk_planes(chunk)
        run.init_simulation()
        U0 = run.getDataHost(0)
        run.make_all_boundaries(0)
        run.setDataHost(run.getDataHost(0), 1)
        n, t, dt, dts = 0, 0.0, 0.0, []
        for _ in range(nsteps):
            n, t, dt = run.oneStepIntegration(n, t, dt)
            dts.append(dt)
        return run.getDataHost(n), t, np.array(dts), run.layout.ghost_width, U0, run.stats().kernel_launches


def check(ref, got, names, tol, tag):
    """L2-relative error per variable; vector components against the norm of their vector field
    (components that stay ~0 by symmetry have no meaningful norm of their own at round-off)."""
    ref, got = ref.astype(np.float64), got.astype(np.float64)
    mom = np.sqrt(sum(float(np.sum(ref[v] ** 2)) for v in (2, 3, 4)))
    mag = np.sqrt(sum(float(np.sum(ref[v] ** 2)) for v in (5, 6, 7))) if len(ref) > 5 else 1.0
    for v, vname in enumerate(names):
        norm = np.sqrt(np.sum(ref[v] ** 2)) if v < 2 else (mom if v < 5 else mag)
        err = np.sqrt(np.sum((ref[v] - got[v]) ** 2)) / max(norm, 1e-300)
        assert err < tol, (tag, vname, err)


@pytest.mark.parametrize("name", MHD_CASES)
def test_mhd_golden_reference_run(native, name):
    g = load_golden(name)
    U, t, dts, gw, U0, _ = run_gpu(str(g["ini"]), int(g["steps"]), mhd=True)
    assert np.array_equal(U0[:, gw:-gw, gw:-gw, gw:-gw], g["initial"])   # initial condition: bitwise
    check(g["final"], U[:, gw:-gw, gw:-gw, gw:-gw], g["names"], TOL_F64, name)
    assert abs(t - g["total_time"]) < 1e-10 * g["total_time"]
    assert abs(dts[-1] - g["dt_last"]) < 1e-10 * g["dt_last"]


@pytest.mark.parametrize("name,tol", HYDRO_CASES)
def test_hydro_golden_reference_run(native, name, tol):
    g = load_golden(name)
    fp32 = str(g["precision"]) == "f32"
    U, t, dts, gw, U0, _ = run_gpu(str(g["ini"]), int(g["steps"]), mhd=False, fp32=fp32)
    assert np.array_equal(U0[:, gw:-gw, gw:-gw, gw:-gw], g["initial"])
    check(g["final"], U[:, gw:-gw, gw:-gw, gw:-gw], g["names"], tol, name)
    assert abs(dts[0] - g["dt0"]) < 2e-6 * g["dt0"]


@pytest.mark.parametrize("name,over", [
    # resistivity only, viscosity only, both; chunked pipeline (the dissipative kernels run on the whole slab)
    ("ot3d_diss_16x12x20_s6", {"hydro": {"nu": 0.0}}),
    ("ot3d_diss_16x12x20_s6", {"MHD": {"eta": 0.0}}),
    ("ot3d_diss_16x12x20_s6", {"mesh": {"nx": 18, "ny": 14, "nz": 12}}),
    ("rt3d_mhd_visc_rand_8x10x16_s5", {"mesh": {"nx": 12, "ny": 8, "nz": 20}}),
    ("mri3d_diss_12x20x8_s10", {"mesh": {"nx": 10, "ny": 16, "nz": 12}}),
])
def test_mhd_full_array_vs_oracle(native, oracle64, name, over):
    g = load_golden(name)
    ini = ini_override(str(g["ini"]), over)
    p = oracle64.params(ini)
    nsteps = 7
    Ug, tg, dtg, gw, U0, _ = run_gpu(ini, nsteps, mhd=True, chunk=5)
    Uo0 = oracle64.init_problem(p)
    if name.startswith("rt3d"):
        assert np.array_equal(U0, Uo0)          # Rayleigh-Taylor initialises every cell, ghosts included
    Uo, to, dto = oracle64.run_steps(p, Uo0, nsteps)
    names = ["d", "e", "mx", "my", "mz", "bx", "by", "bz"]
    rot = p.Omega0 > 0
    sl = (slice(None),) * 4 if rot else (slice(None), slice(gw, -gw), slice(gw, -gw), slice(gw, -gw))
    check(Uo[sl], Ug[sl], names, TOL_F64, name)   # rotating step ends with its ghost fill: whole arrays
    assert np.allclose(dtg, dto, rtol=1e-12)


def test_hydro_viscous_gravity_vs_oracle(native, oracle64):
    """viscosity + gravity together on the hydro path, HLLC, chunked"""
    g = load_golden("rt3d_hydro_10x8x24_s8")
    ini = ini_override(str(g["ini"]), {"mesh": {"nx": 12, "ny": 10, "nz": 18}, "hydro": {"nu": 0.003, "riemannSolver": "hllc"},
                                      "gravity": {"static_field_y": 0.07}})
    p = oracle64.params(ini)
    nsteps = 8
    Ug, tg, dtg, gw, U0, launches = run_gpu(ini, nsteps, mhd=False, chunk=7)
    Uo0 = oracle64.init_problem(p)
    assert np.array_equal(U0, Uo0)
    Uo, to, dto = oracle64.run_steps(p, Uo0, nsteps)
    inner = (slice(None), slice(gw, -gw), slice(gw, -gw), slice(gw, -gw))
    check(Uo[inner], Ug[inner], ["d", "e", "mx", "my", "mz"], TOL_F64, "rt hydro visc")
    assert np.allclose(dtg, dto, rtol=1e-12)
    assert launches > 0


def test_resistive_divb_and_energy_budget(native):
    """size-independent properties at a larger size: the resistive CT update keeps div B at round-off and,
    in the periodic box, resistivity + viscosity conserve mass and momentum."""
    from ramsesgpu_b200 import MHDRunGodunov
    g = load_golden("ot3d_diss_16x12x20_s6")
    ini = ini_override(str(g["ini"]), {"mesh": {"nx": 48, "ny": 40, "nz": 56}})
    with MHDRunGodunov(ini) as run:
        run.init_simulation()
        run.make_all_boundaries(0)
        run.setDataHost(run.getDataHost(0), 1)
        gw = run.layout.ghost_width
        U0 = run.getDataHost(0)
        n, t, dt = 0, 0.0, 0.0
        for _ in range(12):
            n, t, dt = run.oneStepIntegration(n, t, dt)
        run.make_all_boundaries(n % 2)
        U = run.getDataHost(n)
        dx, dy, dz = (run.param(k) for k in ("dx", "dy", "dz"))
    s = (slice(gw, -gw),) * 3
    def sh(a, ax):
        return np.roll(a, -1, axis=ax)[s]
    divb = (sh(U[5], 2) - U[5][s]) / dx + (sh(U[6], 1) - U[6][s]) / dy + (sh(U[7], 0) - U[7][s]) / dz
    bscale = np.abs(U[5:8]).max() / min(dx, dy, dz)
    assert np.abs(divb).max() < 1e-12 * bscale
    for v in (0, 2, 3, 4):   # mass and momenta: conserved to round-off by the flux-form updates
        a, b = U0[v][s].sum(), U[v][s].sum()
        scale = np.abs(U0[v][s]).sum() if v == 0 else sum(np.abs(U0[c][s]).sum() for c in (2, 3, 4))
        assert abs(a - b) < 1e-12 * scale, (v, a, b)
    # Total energy is NOT conserved to round-off by the reference's scheme: the resistive energy flux is
    # evaluated after the resistive CT update with ghost-cell B that was refreshed BEFORE it
    # (mhd_godunov_unsplit_cpu_v3.cpp:666-680), so the flux through the periodic seam differs between its
    # two images.  The CUDA path reproduces that (golden tests above); here only bound the drift.
    a, b = U0[1][s].sum(), U[1][s].sum()
    assert abs(a - b) < 1e-4 * abs(a), (a, b)
    assert np.isfinite(U).all()
