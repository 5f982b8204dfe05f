"""GPU: rotating frame + shearing box (BASELINE.json configs[3], MRI) against golden vectors from the
unmodified reference executable and against the C oracle."""
import numpy as np
import pytest

from conftest import TOL_F64, load_golden
from ramsesgpu_b200.io import ini_override, l2_relative

pytestmark = pytest.mark.gpu


def run_gpu(ini, nsteps, chunk=0):
    from ramsesgpu_b200 import MHDRunGodunov
    with MHDRunGodunov(ini) as run:
        if chunk:
            run.set_chunk_planes(chunk)
        run.init_simulation()
        run.make_all_boundaries(0)           # shearing-box variant at t = 0, like the reference's start()
        run.setDataHost(run.getDataHost(0), 1)
        n, t, dt, dts = 0, 0.0, 0.0, []
        for _ in range(nsteps):
            n, t, dt = run.oneStepIntegration(n, t, dt)
            dts.append(dt)
        return run.getDataHost(n), t, np.array(dts), run.layout.ghost_width


def check(ref, got, names, tol):
    mom = np.sqrt(sum(float(np.sum(ref[v] ** 2)) for v in (2, 3, 4)))
    mag = np.sqrt(sum(float(np.sum(ref[v] ** 2)) for v in (5, 6, 7)))
    for v, vname in enumerate(names):
        # vector components are measured against the norm of their vector field (B_x and B_y start at
        # exactly zero in the MRI problem and stay tiny compared with B_z over a few steps)
        norm = np.sqrt(np.sum(ref[v] ** 2)) if v < 2 else (mom if v < 5 else mag)
        err = np.sqrt(np.sum((ref[v] - got[v]) ** 2)) / norm
        assert err < tol, (vname, err)


@pytest.mark.parametrize("name", ["mri3d_16x32x16_s12", "mri3d_12x20x8_s40"])
def test_golden_reference_run(native, name):
    g = load_golden(name)
    U, t, dts, gw = run_gpu(str(g["ini"]), int(g["steps"]))
    check(g["final"], U[:, gw:-gw, gw:-gw, gw:-gw], g["names"], TOL_F64)
    assert abs(t - g["total_time"]) < 1e-10 * g["total_time"]
    assert abs(dts[-1] - g["dt_last"]) < 1e-10 * g["dt_last"]


def test_full_array_with_ghosts_vs_oracle(native, oracle64):
    """ghost cells included: the shearing-box remap of the x ghosts (y shift growing with time) and
    the end-of-step boundary order Y, shear-X, Z, Y"""
    g = load_golden("mri3d_12x20x8_s40")
    ini = ini_override(str(g["ini"]), {"mesh": {"nx": 10, "ny": 16, "nz": 12}})
    p = oracle64.params(ini)
    nsteps = 25
    Ug, tg, dtg, gw = run_gpu(ini, nsteps)
    Uo, to, dto = oracle64.run_steps(p, oracle64.init_problem(p), nsteps)
    names = ["d", "e", "mx", "my", "mz", "bx", "by", "bz"]
    check(Uo, Ug, names, TOL_F64)          # whole arrays, ghosts included
    assert np.allclose(dtg, dto, rtol=1e-12)


def test_chunked_pipeline_is_identical(native):
    g = load_golden("mri3d_16x32x16_s12")
    ref, _, _, _ = run_gpu(str(g["ini"]), 5)
    got, _, _, _ = run_gpu(str(g["ini"]), 5, chunk=4)
    assert np.array_equal(ref, got)


@pytest.mark.parametrize("mesh,chunk", [
    ({"nx": 16, "ny": 32, "nz": 16}, 0),                                             # one tile column: both x borders in the same tile
    ({"nx": 40, "ny": 36, "nz": 20}, 0),                                             # 3 x 6 tiles, partial last tiles
    ({"nx": 30, "ny": 14, "nz": 24}, 9),                                             # exact tile multiples, z chunks
    ("ot3d", 0),                                                                     # rotating frame without the shearing border
])
def test_fused_rotating_kernel_equals_separate_kernels(native, mesh, chunk):
    """The rotating-frame instantiation of the fused flux + emf + update kernel (shear terms in the y flux and the emfs,
    update_cell_rot, border strips + k_update_rot_border for the three cell columns that read the y-remapped opposite
    border) against the separate k_flux / k_emf / k_update_rot kernels: the same device functions on the same inputs;
    the compiler contracts a few multiply-adds differently, so agreement is to the last bits."""
    from ramsesgpu_b200 import set_tuning
    if mesh == "ot3d":   # adiabatic Orszag-Tang in a rotating frame, periodic box: update_cell_rot on every cell of every tile
        from conftest import ot3d_ini
        ini = ot3d_ini((34, 20, 12), OrszagTang={"kt": 1.0}, MHD={"omega0": 0.4})
    else:
        g = load_golden("mri3d_16x32x16_s12")
        ini = ini_override(str(g["ini"]), {"mesh": mesh})
    try:
        set_tuning("fused_b", 0)
        ref, tr, dtr, gw = run_gpu(ini, 8, chunk=chunk)
        set_tuning("fused_b", 1)
        got, tg, dtg, _ = run_gpu(ini, 8, chunk=chunk)
    finally:
        set_tuning("fused_b", 1)
    assert np.allclose(dtr, dtg, rtol=1e-13, atol=0)
    names = ["d", "e", "mx", "my", "mz", "bx", "by", "bz"]
    check(ref, got, names, 1e-13)            # whole arrays, ghosts included


def test_many_tiles_shearing_box_vs_oracle(native, oracle64):
    """several tile columns / rows of the fused kernels with both shearing borders, against the oracle (6 steps)"""
    g = load_golden("mri3d_16x32x16_s12")
    ini = ini_override(str(g["ini"]), {"mesh": {"nx": 48, "ny": 40, "nz": 12}})
    p = oracle64.params(ini)
    nsteps = 6
    Ug, tg, dtg, gw = run_gpu(ini, nsteps)
    Uo, to, dto = oracle64.run_steps(p, oracle64.init_problem(p), nsteps)
    check(Uo, Ug, ["d", "e", "mx", "my", "mz", "bx", "by", "bz"], TOL_F64)
    assert np.allclose(dtg, dto, rtol=1e-12)


def test_rotating_fused_kernel_dt_in_kernel_or_separate(native):
    """knob rot_dt: the rotating-frame fused kernel reduces the inverse dt of the new state itself, or leaves it to the
    stand-alone reduction (k_invdt) on the ghost-filled new state: same cells, same formula, same bits"""
    from ramsesgpu_b200 import set_tuning
    g = load_golden("mri3d_16x32x16_s12")
    ini = ini_override(str(g["ini"]), {"mesh": {"nx": 40, "ny": 36, "nz": 20}})
    try:
        set_tuning("rot_dt", 1)
        a, ta, dta, _ = run_gpu(ini, 6)
        set_tuning("rot_dt", 0)
        b, tb, dtb, _ = run_gpu(ini, 6)
    finally:
        set_tuning("rot_dt", 0)
    assert np.array_equal(dta, dtb) and np.array_equal(a, b)


def test_stratified_shearing_box_golden(native):
    """mhd_mri_3d_stratified.ini: vertical gravity g_z(z) (predictor in the trace, source term before the border remap
    of the density), BC_Z_STRATIFIED ghost planes (hydrostatic density extrapolation, div-B-free B_z), stratified initial
    condition -- against the final state of the unmodified reference executable."""
    g = load_golden("mri3d_strat_8x12x24_s10")
    U, t, dts, gw = run_gpu(str(g["ini"]), int(g["steps"]))
    check(g["final"], U[:, gw:-gw, gw:-gw, gw:-gw], g["names"], TOL_F64)
    assert abs(t - g["total_time"]) < 1e-10 * g["total_time"]
    assert abs(dts[-1] - g["dt_last"]) < 1e-10 * g["dt_last"]


def test_stratified_shearing_box_with_ghosts_vs_oracle(native, oracle64):
    """whole arrays, ghost planes of the stratified z boundary included, several tiles, 12 steps"""
    g = load_golden("mri3d_strat_8x12x24_s10")
    ini = ini_override(str(g["ini"]), {"mesh": {"nx": 20, "ny": 16, "nz": 28}})
    p = oracle64.params(ini)
    nsteps = 12
    Ug, tg, dtg, gw = run_gpu(ini, nsteps)
    Uo, to, dto = oracle64.run_steps(p, oracle64.init_problem(p), nsteps)
    check(Uo, Ug, ["d", "e", "mx", "my", "mz", "bx", "by", "bz"], TOL_F64)
    assert np.allclose(dtg, dto, rtol=1e-12)
