"""GPU: parity of the CUDA 2D MHD path (BASELINE.json configs[0]) against golden vectors from the
unmodified reference executable, the C oracle, and the reference's 256^2 x 100-step known answers
recorded in SURVEY.md 8(c)."""
import numpy as np
import pytest

from conftest import TOL_F64, load_golden
from ramsesgpu_b200.io import ini_override, l2_relative

pytestmark = pytest.mark.gpu


def run_gpu(ini, nsteps, U0=None):
    from ramsesgpu_b200 import MHDRunGodunov
    with MHDRunGodunov(ini) as run:
        run.init_simulation()
        if U0 is not None:
            run.setDataHost(U0, 0)
        run.make_all_boundaries(0)
        run.setDataHost(run.getDataHost(0), 1)
        n, t, dt, dts = 0, 0.0, 0.0, []
        for _ in range(nsteps):
            n, t, dt = run.oneStepIntegration(n, t, dt)
            dts.append(dt)
        return run.getDataHost(n), t, np.array(dts), run.layout.ghost_width


@pytest.mark.parametrize("name", ["ot2d_32_s12", "ot2d_40x24_hll_s6"])
def test_golden_reference_run(native, name):
    g = load_golden(name)
    U, t, dts, gw = run_gpu(str(g["ini"]), int(g["steps"]))
    inner = U[:, 0, gw:-gw, gw:-gw]
    for v, vname in enumerate(g["names"]):
        ref = g["final"][v]
        if np.abs(ref).max() == 0.0:          # mz and bz stay exactly zero in the 2D vortex
            assert np.abs(inner[v]).max() == 0.0
            continue
        assert l2_relative(ref, inner[v]) < TOL_F64, (name, vname)
    assert abs(t - g["total_time"]) < 1e-10 * g["total_time"]
    assert abs(dts[-1] - g["dt_last"]) < 1e-10 * g["dt_last"]


def test_random_state_vs_oracle(native, oracle64):
    g = load_golden("ot2d_32_s12")
    ini = ini_override(str(g["ini"]), {"mesh": {"nx": 28, "ny": 20}})
    p = oracle64.params(ini)
    rng = np.random.default_rng(5)
    ny, nx = p.jsize, p.isize
    y, x = np.meshgrid((np.arange(ny) - 3) / p.ny, (np.arange(nx) - 3) / p.nx, indexing="ij")
    def field(a):
        ph = rng.uniform(0, 6.28, 2)
        return a * np.sin(2 * np.pi * x + ph[0]) * np.cos(4 * np.pi * y + ph[1])
    U0 = np.zeros((8, 1, ny, nx))
    U0[0, 0] = 1.0 + field(0.2)
    for v in (2, 3, 4):
        U0[v, 0] = U0[0, 0] * field(0.4)
    for v in (5, 6, 7):
        U0[v, 0] = 0.2 + field(0.2)
    U0[1, 0] = (1.0 + field(0.2)) / (p.gamma0 - 1) + 0.5 * (U0[2, 0] ** 2 + U0[3, 0] ** 2 + U0[4, 0] ** 2) / U0[0, 0] + \
        0.75 * (U0[5, 0] ** 2 + U0[6, 0] ** 2 + U0[7, 0] ** 2)
    Ug, tg, dtg, gw = run_gpu(ini, 4, U0=U0)
    Uo, to, dto = oracle64.run_steps(p, U0.copy(), 4)
    for v in range(8):
        err = l2_relative(Uo[v, 0, gw:-gw, gw:-gw], Ug[v, 0, gw:-gw, gw:-gw])
        assert err < TOL_F64, (v, err)
    assert np.allclose(dtg, dto, rtol=1e-12)


def test_reference_known_answers_256_100_steps(native):
    """BASELINE.json configs[0]: data/orszag-tang.ini at 256x256, 100 steps, values printed by the
    reference executable: initial dt, total time, last dt, density statistics."""
    g = load_golden("ot2d_32_s12")
    ini = ini_override(str(g["ini"]), {"mesh": {"nx": 256, "ny": 256}})
    U, t, dts, gw = run_gpu(ini, 100)
    rho = U[0, 0, gw:-gw, gw:-gw]
    assert abs(dts[0] - 0.000368686397) < 2e-12
    assert abs(t - 0.036300667052) < 1e-11
    assert abs(dts[-1] - 0.000357719884475) < 1e-14
    assert abs(rho.min() - 0.20594064367113102) < 1e-12
    assert abs(rho.max() - 0.25523530898485131) < 1e-12
    assert abs(rho.mean() - 0.22087173089000339) < 1e-13
    assert abs(np.sqrt((rho ** 2).sum()) - 56.590774442039233) < 1e-10
