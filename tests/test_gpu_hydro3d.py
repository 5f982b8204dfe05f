"""GPU: parity of the CUDA 3D hydro (Euler) path against golden vectors from the unmodified reference
executable (FP64 and FP32 builds) and against the C oracle."""
import numpy as np
import pytest

from conftest import TOL_F64, load_golden
from ramsesgpu_b200.io import ini_override, l2_relative

pytestmark = pytest.mark.gpu

# FP32: the reference's float build against our float kernels; both round every operation to 24 bits
# but in a different order (FMA contraction, reciprocals), so agreement is a few float ulps per step.
TOL_F32 = 2e-5

CASES = [("implode3d_16_s8", TOL_F64), ("implode3d_hll_20x12x16_s5", TOL_F64), ("kh3d_16x8x16_f64_s10", TOL_F64),
         ("kh3d_16x8x16_f32_s10", TOL_F32)]


def run_gpu(ini, nsteps, fp32=False, U0=None, chunk=0):
    from ramsesgpu_b200 import HydroRunGodunov
    with HydroRunGodunov(ini, fp32=fp32) as run:
        if chunk:
            run.set_chunk_planes(chunk)
        run.init_simulation()
        if U0 is not None:
            run.setDataHost(U0, 0)
        run.make_all_boundaries(0)
        run.setDataHost(run.getDataHost(0), 1)
        n, t, dt, dts = 0, 0.0, 0.0, []
        for _ in range(nsteps):
            n, t, dt = run.oneStepIntegration(n, t, dt)
            dts.append(dt)
        return run.getDataHost(n), np.array(dts), run.layout.ghost_width


@pytest.mark.parametrize("name,tol", CASES)
def test_golden_reference_run(native, name, tol):
    g = load_golden(name)
    fp32 = str(g["precision"]) == "f32"
    U, dts, gw = run_gpu(str(g["ini"]), int(g["steps"]), fp32=fp32)
    assert U.dtype == (np.float32 if fp32 else np.float64)
    inner = U[:, gw:-gw, gw:-gw, gw:-gw]
    # momentum components are measured against the norm of the whole momentum field: a component
    # that only carries the small seeded perturbation (KH "my") or stays ~0 by symmetry (implode)
    # has no meaningful norm of its own at float / double round-off
    mom_norm = np.sqrt(sum(float(np.sum(g["final"][v].astype(np.float64) ** 2)) for v in (2, 3, 4)))
    for v, vname in enumerate(g["names"]):
        ref, got = g["final"][v].astype(np.float64), inner[v].astype(np.float64)
        norm = np.sqrt(np.sum(ref ** 2)) if v < 2 else max(mom_norm, 1e-300)
        err = np.sqrt(np.sum((ref - got) ** 2)) / norm
        assert err < tol, (name, vname, err)
    assert abs(dts[0] - g["dt0"]) < 2e-6 * g["dt0"]


def test_initial_conditions_bitwise(native, oracle64, oracle32):
    from ramsesgpu_b200 import HydroRunGodunov
    for name in ("implode3d_16_s8", "kh3d_16x8x16_f32_s10", "kh3d_16x8x16_f64_s10"):
        g = load_golden(name)
        fp32 = str(g["precision"]) == "f32"
        with HydroRunGodunov(str(g["ini"]), fp32=fp32) as run:
            run.init_simulation()
            U = run.getDataHost(0)
            gw = run.layout.ghost_width
        # glibc rand() stream consumed in the reference's order: bit-identical initial state
        assert np.array_equal(U[:, gw:-gw, gw:-gw, gw:-gw], g["initial"]), name


@pytest.mark.parametrize("solver,slope", [("hllc", 2.0), ("approx", 2.0), ("hll", 1.0), ("approx", 1.0)])
def test_random_state_vs_oracle(native, oracle64, solver, slope):
    g = load_golden("kh3d_16x8x16_f64_s10")
    ini = ini_override(str(g["ini"]), {"mesh": {"nx": 14, "ny": 10, "nz": 12}, "hydro": {"riemannSolver": solver, "slope_type": slope}})
    p = oracle64.params(ini)
    rng = np.random.default_rng(11)
    nz, ny, nx = p.ksize, p.jsize, p.isize
    z, y, x = np.meshgrid(np.arange(nz) / p.nz, np.arange(ny) / p.ny, np.arange(nx) / p.nx, indexing="ij")
    def field(a):
        ph = rng.uniform(0, 6.28, 3)
        return a * np.sin(2 * np.pi * x + ph[0]) * np.cos(2 * np.pi * y + ph[1]) * np.sin(4 * np.pi * z + ph[2])
    U0 = np.zeros((5, nz, ny, nx))
    U0[0] = 1.0 + field(0.3)
    for v in (2, 3, 4):
        U0[v] = U0[0] * field(0.6)
    U0[1] = (1.0 + field(0.3)) / (p.gamma0 - 1) + 0.5 * (U0[2] ** 2 + U0[3] ** 2 + U0[4] ** 2) / U0[0]
    Ug, dtg, gw = run_gpu(ini, 4, U0=U0)
    Uo, _, dto = oracle64.run_steps(p, U0.copy(), 4)
    for v in range(5):
        err = l2_relative(Uo[v, gw:-gw, gw:-gw, gw:-gw], Ug[v, gw:-gw, gw:-gw, gw:-gw])
        assert err < TOL_F64, (v, err)
    assert np.allclose(dtg, dto, rtol=1e-12)


def test_chunked_pipeline_is_identical(native):
    """z chunks of the two-kernel path (trace + flux/update through W; the fused one-kernel step has no scratch to chunk)"""
    from ramsesgpu_b200 import set_tuning
    g = load_golden("implode3d_16_s8")
    try:
        set_tuning("hydro_fused", 0)
        ref, _, _ = run_gpu(str(g["ini"]), 4)
        for chunk in (1, 3, 5):
            got, _, _ = run_gpu(str(g["ini"]), 4, chunk=chunk)
            assert np.array_equal(ref, got), chunk
    finally:
        set_tuning("hydro_fused", 1)


def test_conservation_periodic_fp32_full_size(native):
    """128^3 FP32 Kelvin-Helmholtz, 10 steps: mass, momentum and energy sums are conserved to float
    round-off in the periodic box (size-independent property, no oracle needed)."""
    g = load_golden("kh3d_16x8x16_f32_s10")
    ini = ini_override(str(g["ini"]), {"mesh": {"nx": 128, "ny": 128, "nz": 128}})
    from ramsesgpu_b200 import HydroRunGodunov
    with HydroRunGodunov(ini, fp32=True) as run:
        run.init_simulation()
        run.make_all_boundaries(0)
        run.setDataHost(run.getDataHost(0), 1)
        gw = run.layout.ghost_width
        U0 = run.getDataHost(0).astype(np.float64)
        n, t, dt = 0, 0.0, 0.0
        for _ in range(10):
            n, t, dt = run.oneStepIntegration(n, t, dt)
        U = run.getDataHost(n).astype(np.float64)
    for v in range(5):
        a, b = U0[v, gw:-gw, gw:-gw, gw:-gw], U[v, gw:-gw, gw:-gw, gw:-gw]
        assert abs(a.sum() - b.sum()) / (np.abs(a).sum() + 1.0) < 5e-6, v
    assert np.isfinite(U).all() and U[0].min() > 0


def test_tile_and_gather_kernels_agree(native):
    """The register-tiled flux+update kernel and the gather variant of the two-kernel path solve every face with the same
    code on the same inputs and sum in the same order; the compiler may contract a multiply-add differently in the two
    kernels, so agreement is to the last bits (FP32 and FP64, odd sizes so that tile seams and partial tiles are exercised)."""
    from ramsesgpu_b200 import set_tuning
    for name, mesh in (("kh3d_16x8x16_f32_s10", {"nx": 67, "ny": 19, "nz": 23}),
                       ("implode3d_16_s8", {"nx": 35, "ny": 31, "nz": 70})):
        g = load_golden(name)
        fp32 = str(g["precision"]) == "f32"
        ini = ini_override(str(g["ini"]), {"mesh": mesh})
        try:
            set_tuning("hydro_fused", 0)
            set_tuning("hydro_tile", 0)
            Ua, dta, gw = run_gpu(ini, 6, fp32=fp32)
            set_tuning("hydro_tile", 1)
            Ub, dtb, _ = run_gpu(ini, 6, fp32=fp32)
        finally:
            set_tuning("hydro_tile", 1)
            set_tuning("hydro_fused", 1)
        inner = (slice(None), slice(gw, -gw), slice(gw, -gw), slice(gw, -gw))
        eps = 2e-6 if fp32 else 1e-14
        a, b = Ua[inner].astype(np.float64), Ub[inner].astype(np.float64)
        scale = np.abs(a).max(axis=(1, 2, 3), keepdims=True)
        scale[2:5] = scale[2:5].max()
        assert (np.abs(a - b) <= 20 * eps * scale).all(), name
        assert np.allclose(dta, dtb, rtol=10 * eps, atol=0), name


@pytest.mark.parametrize("name,mesh,over", [
    ("kh3d_16x8x16_f32_s10", {"nx": 67, "ny": 19, "nz": 23}, {}),            # partial tiles in x and y
    ("kh3d_16x8x16_f64_s10", {"nx": 56, "ny": 16, "nz": 40}, {}),            # exact multiples of the 28 x 8 FP64 tile
    ("implode3d_16_s8", {"nx": 35, "ny": 31, "nz": 70}, {}),                 # walls, approx solver, several z ranges
    ("implode3d_hll_20x12x16_s5", {"nx": 29, "ny": 13, "nz": 9}, {}),        # HLL, minmod
    ("rt3d_hydro_10x8x24_s8", {"nx": 30, "ny": 12, "nz": 24}, {}),           # static gravity (predictor + source term)
])
def test_fused_step_equals_two_kernel_path(native, name, mesh, over):
    """The one-kernel hydro step (kernels_hydro3d_fused.cu: primitives in registers / shared memory, no W scratch)
    against trace + flux/update through W: the same per-cell functions on the same inputs.  The compiler may contract
    multiply-adds differently in the two kernels, so agreement is to the last bits (a few ulp), not bitwise."""
    from ramsesgpu_b200 import set_tuning
    g = load_golden(name)
    fp32 = str(g["precision"]) == "f32"
    ini = ini_override(str(g["ini"]), dict({"mesh": mesh}, **over))
    try:
        set_tuning("hydro_fused", 0)
        Ua, dta, gw = run_gpu(ini, 6, fp32=fp32)
        set_tuning("hydro_fused", 1)
        Ub, dtb, _ = run_gpu(ini, 6, fp32=fp32)
    finally:
        set_tuning("hydro_fused", 1)
    eps = 2e-6 if fp32 else 1e-14
    assert np.allclose(dta, dtb, rtol=10 * eps, atol=0)
    scale = np.abs(Ua.astype(np.float64)).max(axis=(1, 2, 3), keepdims=True)
    scale[2:5] = scale[2:5].max()   # momentum components against the largest of them
    assert (np.abs(Ua.astype(np.float64) - Ub.astype(np.float64)) <= 20 * eps * scale).all()      # ghosts included
    print(name, "fused vs two-kernel: bitwise" if np.array_equal(Ua, Ub) else "fused vs two-kernel: max diff %.2e" %
          float(np.abs(Ua.astype(np.float64) - Ub.astype(np.float64)).max()))


def test_fused_step_full_size_fp32_vs_oracle_slab(native, oracle32):
    """configs[2] geometry at a size the oracle finishes in seconds: 120 x 100 x 48 FP32 Kelvin-Helmholtz (HLLC, rand()
    perturbation), 3 steps, 5 x 9 tiles x several z ranges of the fused kernel against the FP32 oracle restatement."""
    g = load_golden("kh3d_16x8x16_f32_s10")
    ini = ini_override(str(g["ini"]), {"mesh": {"nx": 120, "ny": 100, "nz": 48}})
    p = oracle32.params(ini)
    Ug, dtg, gw = run_gpu(ini, 3, fp32=True)
    Uo, _, dto = oracle32.run_steps(p, oracle32.init_problem(p), 3)
    mom = np.sqrt(sum(float(np.sum(Uo[v, gw:-gw, gw:-gw, gw:-gw].astype(np.float64) ** 2)) for v in (2, 3, 4)))
    for v in range(5):
        ref, got = Uo[v, gw:-gw, gw:-gw, gw:-gw].astype(np.float64), Ug[v, gw:-gw, gw:-gw, gw:-gw].astype(np.float64)
        norm = np.sqrt(np.sum(ref ** 2)) if v < 2 else mom
        assert np.sqrt(np.sum((ref - got) ** 2)) / norm < TOL_F32, v
    assert np.allclose(dtg, dto, rtol=1e-5)


@pytest.mark.parametrize("name,mesh", [
    ("kh3d_16x8x16_f32_s10", {"nx": 68, "ny": 20, "nz": 24}),          # FP32: row pitch 72 floats = 288 bytes (16-byte multiple)
    ("kh3d_16x8x16_f64_s10", {"nx": 56, "ny": 16, "nz": 40}),
    ("implode3d_16_s8", {"nx": 36, "ny": 30, "nz": 70}),               # walls, several z ranges
])
def test_fused_step_tma_tiles_equal_per_thread_loads(native, name, mesh):
    """The conservative tiles of the fused hydro kernel arrive by TMA (one cp.async.bulk.tensor per plane into a 4-plane
    shared-memory ring, knob hydro_tma = 1, default) or by per-thread loads (hydro_tma = 0, also the fallback for row
    pitches that are not a multiple of 16 bytes): same values, same arithmetic, BITWISE equal results for the same block
    shape (the default shape differs between the two, and two block shapes agree to round-off only: the compiler contracts
    multiply-adds per instantiation)."""
    from ramsesgpu_b200 import set_tuning
    g = load_golden(name)
    fp32 = str(g["precision"]) == "f32"
    ini = ini_override(str(g["ini"]), {"mesh": mesh})
    try:
        set_tuning("hydro_rows", 20 if fp32 else 12)
        set_tuning("hydro_tma", 0)
        Ua, dta, gw = run_gpu(ini, 6, fp32=fp32)
        set_tuning("hydro_tma", 1)
        Ub, dtb, _ = run_gpu(ini, 6, fp32=fp32)
    finally:
        set_tuning("hydro_tma", 1)
        set_tuning("hydro_rows", 0)
    inner = (slice(None), slice(gw, -gw), slice(gw, -gw), slice(gw, -gw))
    assert np.array_equal(Ua[inner], Ub[inner]) and np.array_equal(dta, dtb)
