"""GPU: error behaviour of a live handle (status code + message, the handle stays usable): wrong buffer sizes,
variants that are not built, and that a failed call does not poison the next one."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_golden
from ramsesgpu_b200.io import ini_override

pytestmark = pytest.mark.gpu


def test_wrong_buffer_size_is_refused_and_the_handle_survives(native):
    from ramsesgpu_b200 import MHDRunGodunov, _lib
    g = load_golden("ot3d_16_s10")
    with MHDRunGodunov(str(g["ini"])) as run:
        run.init_simulation()
        U = run.getDataHost(0)
        short = np.zeros(U.size - 1)
        rc = native.rg_copy_from_host(run._h, 0, short.ctypes.data, short.nbytes)
        assert rc == _lib.RG_ERR_INVALID and "size" in native.rg_last_error().decode()
        rc = native.rg_copy_to_host(run._h, 0, short.ctypes.data, short.nbytes)
        assert rc == _lib.RG_ERR_INVALID
        # the handle is still good: the golden run goes through
        n, t, dt = 0, 0.0, 0.0
        for _ in range(int(g["steps"])):
            n, t, dt = run.oneStepIntegration(n, t, dt)
        gw = run.layout.ghost_width
        got = run.getDataHost(n)[:, gw:-gw, gw:-gw, gw:-gw]
        assert np.max(np.abs(got[0] - g["final"][0])) < 1e-12


@pytest.mark.parametrize("override,needle", [
    ({"mesh": {"boundary_xmin": 5, "boundary_xmax": 5}}, "boundary type"),                 # BC_COPY
    ({"hydro": {"slope_type": 3.0}, "MHD": {"omega0": 0.5}}, "slope_type 3"),              # 27-point slopes, rotating frame
])
def test_variants_that_are_not_built_fail_at_create(native, override, needle):
    from ramsesgpu_b200 import MHDRunGodunov, _lib
    ini = ini_override(str(load_golden("ot3d_16_s10")["ini"]), override)
    with pytest.raises(_lib.RgError, match=needle):
        MHDRunGodunov(ini)


def test_dissipative_terms_in_2d_are_refused_at_the_step(native):
    from ramsesgpu_b200 import MHDRunGodunov, _lib
    ini = ini_override(str(load_golden("ot2d_32_s12")["ini"]), {"MHD": {"eta": 0.01}})
    with MHDRunGodunov(ini) as run:
        run.init_simulation()
        with pytest.raises(_lib.RgError, match="3D solvers only"):
            run.oneStepIntegration(0, 0.0, 0.0)
