"""CPU: the product's point-wise device functions (ramsesgpu_b200/csrc/mhd_device.cuh, hydro_device.cuh) compiled
for the HOST (tests/host_emul: CUDA intrinsics shimmed, the MUFU seeds emulated as 20-bit values) against the
oracle on random states.  This is the same SOURCE as the sm_100a kernels run: every solver (HLLD / HLL / LLF,
2-D HLLD / HLLA / HLLF / LLF with and without the shearing-box terms, hydro approx / HLL / HLLC in FP64 and
FP32), the two limiter formulations, the branch-free reciprocal / rsqrt / sqrt, cons -> prim.  The GPU tests
check the same functions on the device (tests/test_gpu_mhd3d.py::test_device_probes_vs_oracle)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden, ot3d_ini
from ramsesgpu_b200.io import ini_override

HERE = os.path.join(ROOT, "tests", "host_emul")
D = C.POINTER(C.c_double)
F = C.POINTER(C.c_float)


# the product as built, and the experimental formulations kept behind macros for the next A/B on the GPU
# (product: clamps / sqrt guards on the integer pipe, sign-flipped three-way limiter; "fp64forms": the compare-select
# formulations they replaced in round 2, kept behind RG_FP64_CLAMP / RG_LIMITER_V0)
FLAVOURS = {"product": [], "fp64forms": ["-DRG_FP64_CLAMP", "-DRG_LIMITER_V0"]}


@pytest.fixture(scope="module", params=list(FLAVOURS))
def emu(request):
    out = os.path.join(HERE, "_build")
    os.makedirs(out, exist_ok=True)
    lib = os.path.join(out, "libdevice_math_host_%s.so" % request.param)
    csrc = os.path.join(ROOT, "ramsesgpu_b200", "csrc")
    srcs = [os.path.join(HERE, "device_math_host.cpp"), os.path.join(csrc, "config_map.cpp"), os.path.join(csrc, "params.cpp")]
    deps = srcs + [os.path.join(HERE, "cuda_host_shim.h"), os.path.join(csrc, "mhd_device.cuh"), os.path.join(csrc, "hydro_device.cuh")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        # -ffp-contract=fast: let the host compiler fuse multiply-adds like nvcc does
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", "-fPIC", "-shared", "-ffp-contract=fast", "-I", csrc, "-I", HERE,
                               "-I", cuda_inc] + FLAVOURS[request.param] + srcs + ["-o", lib])
    L = C.CDLL(lib)
    L.emu_riemann_mhd.argtypes = [C.c_char_p, C.c_int, D, D, D]
    L.emu_compute_emf.argtypes = [C.c_char_p, C.c_int, C.c_int, D, D, D]
    L.emu_riemann_hydro.argtypes = [C.c_char_p, C.c_int, D, D, D]
    L.emu_riemann_hydro_f32.argtypes = [C.c_char_p, C.c_int, F, F, F]
    L.emu_slopes.argtypes = [C.c_double, C.c_int, D, D, D, D, D]
    L.emu_rcp_rsq.argtypes = [C.c_int, D, D, D, D]
    L.emu_cons_to_prim_mhd.argtypes = [C.c_char_p, C.c_int, D, D, C.c_double, D]
    return L


def p64(a):
    return a.ctypes.data_as(D)


def states(rng, m):
    q = np.empty((m, 8))
    q[:, 0] = rng.uniform(0.5, 2.0, m); q[:, 1] = rng.uniform(0.3, 2.0, m)
    q[:, 2:5] = rng.uniform(-1, 1, (m, 3)); q[:, 5:8] = rng.uniform(-1, 1, (m, 3))
    return q


@pytest.mark.parametrize("solver", ["hlld", "hll", "llf"])
def test_mhd_riemann_solvers(emu, oracle64, solver):
    ini = ot3d_ini((8, 8, 8), hydro={"riemannSolver": solver})
    p = oracle64.params(ini)
    rng = np.random.default_rng(11)
    n = 2048
    ql, qr = states(rng, n), states(rng, n)
    qr[: n // 8] = ql[: n // 8] * (1 + 1e-9)            # nearly equal states: the degenerate branches of HLLD
    ql[n // 8: n // 4, 5] = qr[n // 8: n // 4, 5] = 0.0   # vanishing normal field
    f = np.zeros((n, 8))
    emu.emu_riemann_mhd(ini.encode(), n, p64(ql), p64(qr), p64(f))
    fo = np.array([oracle64.riemann_mhd(p, ql[i], qr[i]) for i in range(n)])
    assert np.allclose(f, fo, rtol=1e-11, atol=1e-12), np.abs(f - fo).max()


@pytest.mark.parametrize("mag,omega", [("hlld", 0.0), ("hlla", 0.0), ("hllf", 0.0), ("llf", 0.0), ("hlld", 0.7)])
def test_corner_emf_solvers(emu, oracle64, mag, omega):
    over = {"MHD": {"magRiemannSolver": mag}}
    if omega:
        over["MHD"]["omega0"] = omega
    ini = ot3d_ini((8, 8, 8), **over)
    p = oracle64.params(ini)
    rng = np.random.default_rng(13)
    n = 1024
    qe = states(rng, 4 * n).reshape(n, 4, 8).copy()
    xpos = rng.uniform(-0.5, 0.5, n)
    for d in range(3):
        e = np.zeros(n)
        emu.emu_compute_emf(ini.encode(), n, d, p64(qe), p64(xpos), p64(e))
        eo = np.array([oracle64.compute_emf(p, d, qe[i], xpos[i]) for i in range(n)])
        assert np.allclose(e, eo, rtol=1e-10, atol=1e-11), (d, np.abs(e - eo).max())


@pytest.mark.parametrize("solver", ["approx", "hll", "hllc"])
def test_hydro_riemann_solvers(emu, oracle64, oracle32, solver):
    base = str(load_golden("implode3d_16_s8")["ini"])
    ini = ini_override(base, {"hydro": {"riemannSolver": solver}})
    rng = np.random.default_rng(17)
    n = 2048
    def hs(m):
        q = np.empty((m, 5))
        q[:, 0] = rng.uniform(0.1, 2.0, m); q[:, 1] = rng.uniform(0.1, 2.0, m); q[:, 2:5] = rng.uniform(-1.5, 1.5, (m, 3))
        return q
    ql, qr = hs(n), hs(n)
    p = oracle64.params(ini)
    f = np.zeros((n, 5))
    emu.emu_riemann_hydro(ini.encode(), n, p64(ql), p64(qr), p64(f))
    fo = np.array([oracle64.riemann_hydro(p, ql[i], qr[i]) for i in range(n)])
    assert np.allclose(f, fo, rtol=1e-10, atol=1e-12), np.abs(f - fo).max()
    # FP32 (BASELINE.json configs[2] runs in float): a few float ulps, like the GPU test tolerance
    p32 = oracle32.params(ini)
    l32, r32 = ql.astype(np.float32), qr.astype(np.float32)
    f32 = np.zeros((n, 5), np.float32)
    emu.emu_riemann_hydro_f32(ini.encode(), n, l32.ctypes.data_as(F), r32.ctypes.data_as(F), f32.ctypes.data_as(F))
    fo32 = np.array([oracle32.riemann_hydro(p32, l32[i], r32[i]) for i in range(n)])
    scale = np.abs(fo32).max(axis=1, keepdims=True) + 1.0
    assert (np.abs(f32 - fo32) / scale).max() < 2e-5


def test_limiters(emu):
    rng = np.random.default_rng(19)
    n = 20000
    qm, q0, qp = rng.normal(size=n), rng.normal(size=n), rng.normal(size=n)
    q0[:200] = qm[:200]                      # zero one-sided difference
    qp[200:400] = q0[200:400]
    qm[400:500] = q0[400:500] = qp[400:500]  # flat
    for st in (1.0, 2.0):
        dlft, drgt, dcen = st * (q0 - qm), st * (qp - q0), 0.5 * (qp - qm)
        ref = np.where(dlft * drgt <= 0, 0.0, np.where(dcen >= 0, 1.0, -1.0) * np.minimum(np.minimum(np.abs(dlft), np.abs(drgt)), np.abs(dcen)))
        full, half = np.zeros(n), np.zeros(n)
        emu.emu_slopes(st, n, p64(qm), p64(q0), p64(qp), p64(full), p64(half))
        assert np.array_equal(full, ref)                                   # reference formulation: bitwise
        assert np.allclose(2 * half, ref, rtol=4e-16, atol=1e-300)         # FP64-pipe formulation: dcen as (a+b)/2


def test_reciprocal_and_rsqrt_two_ulp(emu):
    rng = np.random.default_rng(23)
    x = np.concatenate([10.0 ** rng.uniform(-200, 200, 20000), rng.uniform(0.5, 2.0, 20000), [1e-300, 1.0, 4.0]])
    n = len(x)
    r, s, q = np.zeros(n), np.zeros(n), np.zeros(n)
    emu.emu_rcp_rsq(n, p64(x), p64(r), p64(s), p64(q))
    ulp = np.finfo(np.float64).eps
    assert (np.abs(r * x - 1.0)).max() < 3 * ulp
    assert (np.abs(s * s * x - 1.0)).max() < 6 * ulp
    assert (np.abs(q / np.sqrt(x) - 1.0)).max() < 4 * ulp


@pytest.mark.parametrize("over", [{}, {"hydro": {"cIso": 0.3}}, {"MHD": {"omega0": 0.5}}])
def test_cons_to_prim_mhd(emu, oracle64, over):
    ini = ot3d_ini((8, 8, 8), **over)
    p = oracle64.params(ini)
    rng = np.random.default_rng(29)
    n = 4096
    u = np.empty((n, 8))
    u[:, 0] = rng.uniform(0.2, 2.0, n); u[:, 2:5] = rng.uniform(-1, 1, (n, 3)) * u[:, :1]; u[:, 5:8] = rng.uniform(-1, 1, (n, 3))
    bn = u[:, 5:8] + rng.uniform(-0.1, 0.1, (n, 3))
    u[:, 1] = rng.uniform(0.5, 3.0, n) + 0.5 * (u[:, 2:5] ** 2).sum(1) / u[:, 0] + 0.5 * (0.25 * (u[:, 5:8] + bn) ** 2).sum(1)
    u[:16, 0] = 1e-12                                                   # density floor
    dt = 0.01
    q = np.zeros((n, 8))
    emu.emu_cons_to_prim_mhd(ini.encode(), n, p64(u), p64(np.ascontiguousarray(bn)), dt, p64(q))
    r = np.maximum(u[:, 0], p.smallr)
    v = u[:, 2:5] / r[:, None]
    B = 0.5 * (u[:, 5:8] + bn)
    if p.cIso > 0:
        pr = r * p.cIso * p.cIso
    else:
        eint = (u[:, 1] - 0.5 * (B ** 2).sum(1)) / r - 0.5 * (v ** 2).sum(1)
        pr = np.maximum((p.gamma0 - 1.0) * r * eint, r * p.smallp)
    if p.Omega0 > 0:   # Coriolis predictor, constoprim.h:189-195
        v = v.copy()
        dvx, dvy = 2.0 * p.Omega0 * v[:, 1], -0.5 * p.Omega0 * v[:, 0]
        v[:, 0] += dvx * dt * 0.5; v[:, 1] += dvy * dt * 0.5
    want = np.column_stack([r, pr, v, B])
    assert np.allclose(q, want, rtol=1e-12, atol=1e-13), np.abs(q - want).max()


def smooth_state(p, seed):
    """smooth, genuinely 3D periodic state with all 8 variables active (ghosts included)"""
    rng = np.random.default_rng(seed)
    nz, ny, nx, g = p.ksize, p.jsize, p.isize, p.ghostWidth
    z, y, x = np.meshgrid((np.arange(nz) - g) / p.nz, (np.arange(ny) - g) / p.ny, (np.arange(nx) - g) / p.nx, indexing="ij")
    def field(amp):
        f = np.zeros_like(x)
        for _ in range(3):
            kx, ky, kz = rng.integers(1, 3, size=3)
            ph = rng.uniform(0, 2 * np.pi, size=3)
            f += amp * np.sin(2 * np.pi * kx * x + ph[0]) * np.cos(2 * np.pi * ky * y + ph[1]) * np.sin(2 * np.pi * kz * z + ph[2])
        return f
    U = np.zeros((8, nz, ny, nx))
    U[0] = 1.0 + field(0.1)
    for v in (2, 3, 4):
        U[v] = U[0] * field(0.3)
    for v in (5, 6, 7):
        U[v] = field(0.4)
    U[1] = 2.0 + field(0.2) + 0.5 * (U[2] ** 2 + U[3] ** 2 + U[4] ** 2) / U[0] + 0.5 * (U[5] ** 2 + U[6] ** 2 + U[7] ** 2)
    return U


@pytest.mark.parametrize("over,name", [
    ({}, "adiabatic"),
    ({"hydro": {"slope_type": 1.0}}, "minmod"),
    ({"hydro": {"slope_type": 3.0}}, "27-point slopes"),
    ({"hydro": {"cIso": 0.4}}, "isothermal"),
    ({"MHD": {"omega0": 0.3}, "hydro": {"cIso": 0.4}}, "rotating frame, isothermal"),
    ({"hydro": {"problem": "Rayleigh-Taylor"}, "gravity": {"static_field_x": 0.3, "static_field_z": -0.7}}, "static gravity"),
])
def test_trace_stage_of_the_product_vs_oracle(emu, oracle64, over, name):
    """The product's per-cell trace functions (mhd_cells.cuh: cons->prim, edge electric field, limited slopes, half-step
    predictor -> W) and the face / edge states its flux and emf stages rebuild from W, run on host arrays, against the
    oracle's 18 trace arrays (the reference's qm, qp, qEdge) of the same state: the W representation (38 components, high
    faces read from the +1 neighbour, gravity predictor folded into the centre value) carries what the reference stores
    in 144 reals per cell."""
    emu.emu_mhd3d_trace_arrays.argtypes = [C.c_char_p, D, C.c_double, D]
    ini = ot3d_ini((10, 8, 9), **over)
    p = oracle64.params(ini)
    U = smooth_state(p, 5)
    dt = 0.4 * oracle64.compute_dt(p, U)
    want = oracle64.mhd3d_trace_arrays(p, U, dt)
    got = np.zeros_like(want)
    emu.emu_mhd3d_trace_arrays(ini.encode(), p64(U), dt, p64(got))
    gw = p.ghostWidth
    lo, hi = gw - 1, (p.isize - gw, p.jsize - gw, p.ksize - gw)
    checked = 0
    for s in range(18):
        # range the product traces; states on a HIGH face / edge need the +1 neighbour: one cell less on that side
        if s < 3:
            plus = {s}
        elif s < 6:
            plus = set()
        else:
            e, d = divmod(s - 6, 3)
            d1, d2 = [(1, 2), (0, 2), (0, 1)][d]
            plus = ({d1} if e in (0, 1) else set()) | ({d2} if e in (0, 2) else set())
        sl = tuple(slice(lo, hi[ax] + (0 if ax in plus else 1)) for ax in (2, 1, 0))   # k, j, i
        a, b = got[s][(slice(None),) + sl], want[s][(slice(None),) + sl]
        scale = np.abs(b).max(axis=(1, 2, 3), keepdims=True) + 1e-3
        err = (np.abs(a - b) / scale).max()
        assert err < 1e-13, (name, s, err)
        checked += a.size
    assert checked > 18 * 8 * 200


@pytest.mark.parametrize("n,kt,seed", [((10, 8, 9), 1.0, 3), ((14, 6, 7), 0.0, 4)])
def test_whole_step_of_the_headline_configuration_vs_oracle(emu, oracle64, n, kt, seed):
    """One whole step of the FAST configuration (adiabatic HLLD + 2-D HLLD, the bench workload) assembled on the host from
    the product's per-cell functions (mhd_cells.cuh: the same source the fused TMA kernels and the separate kernels
    instantiate) over the kernels' index ranges, against the oracle's step: conservative update, constrained transport
    (incl. the first upper ghost faces) and the dt of the new state."""
    emu.emu_mhd3d_step_fast.argtypes = [C.c_char_p, D, C.c_double, D]
    emu.emu_mhd3d_step_fast.restype = C.c_double
    ini = ot3d_ini(n, OrszagTang={"kt": kt})
    p = oracle64.params(ini)
    U = smooth_state(p, seed)
    oracle64.make_all_boundaries(p, U)
    dt = oracle64.compute_dt(p, U)
    want = np.zeros_like(U)
    oracle64.step_no_boundaries(p, U, want, dt)
    got = U.copy()
    inv_dt = emu.emu_mhd3d_step_fast(ini.encode(), p64(U), dt, p64(got))
    gw = p.ghostWidth
    box = (slice(None), slice(gw, p.ksize - gw + 1), slice(gw, p.jsize - gw + 1), slice(gw, p.isize - gw + 1))
    inner = (slice(None), slice(gw, -gw), slice(gw, -gw), slice(gw, -gw))
    for v in range(8):
        scale = np.abs(want[v][inner[1:]]).max() + 1e-3
        sl = inner if v < 5 else box          # face fields: the update box includes the first upper ghost faces
        assert (np.abs(got[v][sl[1:]] - want[v][sl[1:]]) / scale).max() < 1e-13, v
    # cells outside the update box are untouched, like the reference's copy of UOld
    mask = np.ones(U.shape[1:], bool); mask[box[1:]] = False
    assert np.array_equal(got[:, mask], U[:, mask])
    # next dt from the inverse dt reduced inside the update (seed of the running max: MHDRunBase.cpp:144)
    dt_next = p.cfl / max(inv_dt, p.smallc / min(p.dx, p.dy))
    assert abs(dt_next - oracle64.compute_dt(p, want)) < 1e-13 * dt_next


@pytest.mark.parametrize("over,name", [
    ({"hydro": {"riemannSolver": "hll"}, "MHD": {"magRiemannSolver": "hlla"}}, "HLL + HLLA"),
    ({"hydro": {"riemannSolver": "llf"}, "MHD": {"magRiemannSolver": "llf"}}, "LLF"),
    ({"MHD": {"magRiemannSolver": "hllf"}}, "HLLD + HLLF"),
    ({"hydro": {"cIso": 0.4}}, "isothermal"),
    ({"hydro": {"slope_type": 3.0}}, "27-point slopes"),
    ({"hydro": {"slope_type": 1.0, "problem": "Rayleigh-Taylor"}, "gravity": {"static_field_y": 0.2, "static_field_z": -0.6}}, "minmod + gravity"),
    ({"mesh": {"boundary_xmin": 2, "boundary_xmax": 2, "boundary_ymin": 1, "boundary_ymax": 1, "boundary_zmin": 2, "boundary_zmax": 1}}, "walls"),
])
def test_whole_step_of_the_generic_path_vs_oracle(emu, oracle64, over, name):
    """the separate-kernel path (k_prim, k_elec, k_trace, k_flux, k_emf, k_update instantiated with FAST = false) assembled
    on the host from the same per-cell functions, for the solver / equation-of-state / slope / gravity / boundary variants"""
    emu.emu_mhd3d_step_generic.argtypes = [C.c_char_p, D, C.c_double, D]
    emu.emu_mhd3d_step_generic.restype = C.c_double
    ini = ot3d_ini((9, 10, 8), OrszagTang={"kt": 1.0}, **over)
    p = oracle64.params(ini)
    U = smooth_state(p, 8)
    oracle64.make_all_boundaries(p, U)
    dt = oracle64.compute_dt(p, U)
    want = np.zeros_like(U)
    oracle64.step_no_boundaries(p, U, want, dt)
    got = U.copy()
    inv_dt = emu.emu_mhd3d_step_generic(ini.encode(), p64(U), dt, p64(got))
    gw = p.ghostWidth
    box = (slice(gw, p.ksize - gw + 1), slice(gw, p.jsize - gw + 1), slice(gw, p.isize - gw + 1))
    inner = (slice(gw, -gw), slice(gw, -gw), slice(gw, -gw))
    for v in range(8):
        scale = np.abs(want[v][inner]).max() + 1e-3
        sl = inner if v < 5 else box
        assert (np.abs(got[v][sl] - want[v][sl]) / scale).max() < 1e-13, (name, v)
    dt_next = p.cfl / max(inv_dt, p.smallc / min(p.dx, p.dy))
    assert abs(dt_next - oracle64.compute_dt(p, want)) < 1e-13 * dt_next


@pytest.mark.parametrize("solver,over,name", [
    ("approx", {}, "two-shock solver"), ("hll", {"hydro": {"slope_type": 1.0}}, "HLL minmod"), ("hllc", {}, "HLLC"),
    ("hllc", {"hydro": {"problem": "Rayleigh-Taylor"}, "gravity": {"static_field_x": 0.1, "static_field_z": -0.4}}, "HLLC + gravity"),
])
def test_whole_hydro_step_vs_oracle(emu, oracle64, oracle32, solver, over, name):
    """3D hydro (BASELINE.json configs[2] and [4]): one step assembled on the host from hydro_cells.cuh in FP64 and FP32"""
    emu.emu_hydro3d_step.argtypes = [C.c_char_p, D, C.c_double, D]; emu.emu_hydro3d_step.restype = C.c_double
    emu.emu_hydro3d_step_f32.argtypes = [C.c_char_p, F, C.c_float, F]; emu.emu_hydro3d_step_f32.restype = C.c_double
    ov = {"mesh": {"nx": 9, "ny": 7, "nz": 8, "boundary_xmin": 3, "boundary_xmax": 3, "boundary_ymin": 3, "boundary_ymax": 3,
                   "boundary_zmin": 3, "boundary_zmax": 3}, "hydro": {"riemannSolver": solver}}
    for k, v in over.items():
        ov.setdefault(k, {}).update(v)
    ini = ini_override(str(load_golden("implode3d_16_s8")["ini"]), ov)
    for orc, dtype, fn, tol in ((oracle64, np.float64, emu.emu_hydro3d_step, 1e-13), (oracle32, np.float32, emu.emu_hydro3d_step_f32, 2e-6)):
        p = orc.params(ini)
        g = p.ghostWidth
        U8 = smooth_state(type("P", (), dict(ksize=p.ksize, jsize=p.jsize, isize=p.isize, ghostWidth=g, nx=p.nx, ny=p.ny, nz=p.nz)), 9)
        U = np.ascontiguousarray(U8[:5]).astype(dtype)
        U[1] = (2.0 + 0.5 * (U8[2] ** 2 + U8[3] ** 2 + U8[4] ** 2) / U8[0]).astype(dtype)
        orc.make_all_boundaries(p, U)
        dt = orc.compute_dt(p, U)
        want = np.zeros_like(U)
        orc.step_no_boundaries(p, U, want, dt)
        got = U.copy()
        ptr = (lambda a: a.ctypes.data_as(D if dtype == np.float64 else F))
        inv_dt = fn(ini.encode(), ptr(U), dt, ptr(got))
        inner = (slice(None), slice(g, -g), slice(g, -g), slice(g, -g))
        scale = np.abs(want[inner]).max(axis=(1, 2, 3), keepdims=True) + 1e-3
        assert (np.abs(got[inner].astype(np.float64) - want[inner]) / scale).max() < tol, (name, dtype)
        assert abs(p.cfl / inv_dt - orc.compute_dt(p, want)) < 10 * tol * dt


def test_rough_random_states_stress(emu, oracle64):
    """discontinuous random states (jumps of an order of magnitude, regions without field, vanishing normal field, random
    grid shapes, random fractions of the CFL step) through the host-assembled step of the headline and generic paths:
    the product's per-cell functions stay within a few ulp of the oracle (measured 4e-16)"""
    for fn in (emu.emu_mhd3d_step_fast, emu.emu_mhd3d_step_generic):
        fn.argtypes = [C.c_char_p, D, C.c_double, D]
        fn.restype = C.c_double
    rng = np.random.default_rng(1)
    for trial in range(12):
        n = tuple(int(x) for x in rng.integers(5, 11, 3))
        kind = trial % 4
        over, fn = {}, emu.emu_mhd3d_step_fast
        if kind == 1:
            over, fn = {"hydro": {"riemannSolver": "hll"}, "MHD": {"magRiemannSolver": "hllf"}}, emu.emu_mhd3d_step_generic
        elif kind == 2:
            over, fn = {"hydro": {"cIso": 0.7}}, emu.emu_mhd3d_step_generic
        elif kind == 3:
            over, fn = {"hydro": {"slope_type": 3.0}}, emu.emu_mhd3d_step_generic
        ini = ot3d_ini(n, **over)
        p = oracle64.params(ini)
        shape = (8, p.ksize, p.jsize, p.isize)
        U = np.zeros(shape)
        U[0] = rng.uniform(0.1, 3.0, shape[1:])
        U[2:5] = U[0] * rng.uniform(-2, 2, (3,) + shape[1:])
        U[5:8] = rng.uniform(-1.5, 1.5, (3,) + shape[1:]) * (0.0 if trial % 7 == 0 else 1.0)
        if trial % 5 == 0:
            U[5] = 0.0
        U[1] = rng.uniform(0.05, 3.0, shape[1:]) + 0.5 * (U[2:5] ** 2).sum(0) / U[0] + 0.65 * (U[5:8] ** 2).sum(0)
        oracle64.make_all_boundaries(p, U)
        dt = oracle64.compute_dt(p, U) * rng.uniform(0.3, 1.0)
        want = np.zeros_like(U)
        oracle64.step_no_boundaries(p, U, want, dt)
        got = U.copy()
        fn(ini.encode(), p64(U), dt, p64(got))
        g = p.ghostWidth
        inner = (slice(None), slice(g, -g), slice(g, -g), slice(g, -g))
        scale = np.abs(want[inner]).max(axis=(1, 2, 3), keepdims=True) + 1e-3
        assert np.isfinite(want).all()
        assert (np.abs(got[inner] - want[inner]) / scale).max() < 1e-13, (trial, n, kind)
