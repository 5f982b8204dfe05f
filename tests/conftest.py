import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# FP64 parity tolerance of the CUDA path against the reference / oracle: the L2-relative error of
# every variable (test/computeL2relatif.py.in:43-50) must stay below this.  BASELINE.json's
# north_star states 1e-12 on Orszag-Tang after 100 steps.
TOL_F64 = 1e-12


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a CUDA device: the gpu-marked tests are skipped (the product has no CPU
    fallback, rg_create would fail with RG_ERR_NO_DEVICE); on a GPU box nothing is skipped."""
    if not any("gpu" in item.keywords for item in items):
        return
    try:
        from ramsesgpu_b200 import build as b
        b.build()
        from ramsesgpu_b200 import _lib
        ndev = _lib.load().rg_device_count()
    except Exception:
        return  # a missing library must fail loudly in the tests themselves
    if ndev > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (rg_device_count() == 0): GPU parity tests run on the B200 box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: (z[k].item() if z[k].shape == () else z[k]) for k in z.files}


@pytest.fixture(scope="session")
def oracle64():
    from oracle.oracle import Oracle
    return Oracle("f64")


@pytest.fixture(scope="session")
def oracle32():
    from oracle.oracle import Oracle
    return Oracle("f32")


@pytest.fixture(scope="session")
def native():
    """Builds (if needed) and loads the native library; GPU tests call through it."""
    from ramsesgpu_b200 import build as b
    b.build()
    from ramsesgpu_b200 import _lib
    return _lib.load()


def ot3d_ini(n=(16, 16, 16), **sections):
    """Orszag-Tang 3D parameter text (the reference's data/orszag-tang3d.ini, inlined so that the
    GPU box does not need /root/reference)."""
    base = load_golden("ot3d_16_s10")["ini"]
    from ramsesgpu_b200.io import ini_override
    ov = {"mesh": {"nx": n[0], "ny": n[1], "nz": n[2]}}
    for k, v in sections.items():
        ov.setdefault(k, {}).update(v)
    return ini_override(str(base), ov)
