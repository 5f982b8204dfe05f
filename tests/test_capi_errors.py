"""CPU: error behaviour of the C ABI on the host side (status code + rg_last_error(), never an abort): bad arguments,
unknown problems, variants that are not built, tuning keys.  The reference prints a message and carries on with a zero
state for an unknown problem name (MHDRunBase.cpp:1338-1341); the host set-up call reports it."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_golden
from ramsesgpu_b200.io import ini_override


def last_error(L):
    return L.rg_last_error().decode(errors="replace")


def test_null_handle_and_null_outputs_are_refused(native):
    from ramsesgpu_b200 import _lib
    L = native
    st = _lib.RgStats()
    assert L.rg_get_stats(None, C.byref(st)) == _lib.RG_ERR_INVALID and "null handle" in last_error(L)
    assert L.rg_one_step(None, None, None, None) == _lib.RG_ERR_INVALID
    assert L.rg_destroy(None) in (_lib.RG_OK, _lib.RG_ERR_INVALID)   # destroying nothing is harmless


def test_unknown_problem_is_reported_with_its_name(native):
    from ramsesgpu_b200 import _lib, initial_condition_host
    ini = ini_override(str(load_golden("implode3d_16_s8")["ini"]), {"hydro": {"problem": "no-such-problem"}})
    with pytest.raises(_lib.RgError) as e:
        initial_condition_host(ini)
    assert e.value.code == _lib.RG_ERR_UNSUPPORTED and "no-such-problem" in str(e.value)


@pytest.mark.parametrize("problem,mesh,mhd,needle", [
    ("riemann2d", {"nx": 8, "ny": 8, "nz": 8}, False, "2D"),            # the reference's set-up is written for 2D only
    ("falling-bubble", {"nx": 8, "ny": 8, "nz": 8}, False, "2D"),        # (its 3D branch mis-indexes the array)
    ("Rayleigh-Taylor", {"nx": 8, "ny": 8}, True, "2D MHD"),             # gravity in the 2D MHD solver is not built
    ("sod", {"nx": 8, "ny": 8, "nz": 8}, True, "for this solver"),       # hydro set-ups are not offered to an MHD run
])
def test_variants_that_are_not_built_fail_loudly(native, problem, mesh, mhd, needle):
    from ramsesgpu_b200 import _lib, initial_condition_host
    base = "ot3d_16_s10" if mhd else "implode3d_16_s8"
    m = {"nx": 8, "ny": 8, "nz": 1}
    m.update(mesh)
    ini = ini_override(str(load_golden(base)["ini"]), {"hydro": {"problem": problem}, "mesh": m})
    with pytest.raises(_lib.RgError) as e:
        initial_condition_host(ini)
    assert e.value.code == _lib.RG_ERR_UNSUPPORTED and needle in str(e.value), str(e.value)


def test_host_buffer_of_the_wrong_size_is_refused(native):
    from ramsesgpu_b200 import _lib
    L = native
    ini = str(load_golden("implode3d_16_s8")["ini"]).encode()
    lay = _lib.RgLayout()
    assert L.rg_initial_condition_host(ini, 0, 0, 1, None, 0, C.byref(lay)) == _lib.RG_OK   # layout query
    n = lay.nvar * lay.ksize * lay.jsize * lay.isize
    U = np.zeros(n - 1)
    rc = L.rg_initial_condition_host(ini, 0, 0, 1, U.ctypes.data_as(C.c_void_p), U.nbytes, C.byref(lay))
    assert rc == _lib.RG_ERR_INVALID and "size" in last_error(L)


def test_slab_arguments_are_checked(native):
    from ramsesgpu_b200 import _lib
    L = native
    a, b = C.c_int(0), C.c_int(0)
    assert L.rg_slab_extent(16, 0, 0, C.byref(a), C.byref(b)) == _lib.RG_ERR_INVALID
    assert L.rg_slab_extent(16, 4, 4, C.byref(a), C.byref(b)) == _lib.RG_ERR_INVALID
    ini = str(load_golden("ot2d_32_s12")["ini"]).encode()
    lay = _lib.RgLayout()
    # z slabs need a 3D run (the minimum slab thickness is checked where it matters, by rg_create_distributed)
    assert L.rg_initial_condition_host(ini, 0, 1, 2, None, 0, C.byref(lay)) == _lib.RG_ERR_INVALID and "3D" in last_error(L)


def test_tuning_keys_and_ranges(native):
    from ramsesgpu_b200 import _lib, set_tuning
    for key, good, bad in (("hydro_rows", 16, 13), ("tile_x", 64, 48), ("handoff_head", 2, 99)):
        set_tuning(key, good)
        with pytest.raises(_lib.RgError) as e:
            set_tuning(key, bad)
        assert e.value.code == _lib.RG_ERR_INVALID
    with pytest.raises(_lib.RgError):
        set_tuning("no_such_knob", 1)
    for key, default in (("hydro_rows", 0), ("tile_x", 32), ("handoff_head", 2), ("halo_p2p", 1), ("hydro_tma", 1)):
        set_tuning(key, default)
