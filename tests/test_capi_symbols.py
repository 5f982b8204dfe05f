"""CPU: the C-ABI library builds for sm_100a, loads, exports every symbol the header declares, and
refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ramsesgpu_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(native):
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(native, s), "missing export %s" % s
    from ramsesgpu_b200 import _lib
    assert sorted(_lib.SIGNATURES) == syms  # the ctypes table covers exactly the header


def test_no_device_means_error_not_fallback(native):
    import torch
    if torch.cuda.is_available():
        pytest.skip("box has a GPU")
    from ramsesgpu_b200 import MHDRunGodunov, _lib
    with pytest.raises(_lib.RgError) as e:
        MHDRunGodunov("[mesh]\nnx=8\nny=8\nnz=8\n[MHD]\nenable=true\n")
    assert e.value.code == _lib.RG_ERR_NO_DEVICE


def test_slab_extent_covers_domain(native):
    from ramsesgpu_b200 import slab_extent
    for nz, n in ((256, 4), (1024, 8), (10, 3), (7, 7)):
        ext = [slab_extent(nz, n, r) for r in range(n)]
        assert ext[0][1] == 0
        for (a, o), (b, p) in zip(ext, ext[1:]):
            assert o + a == p
        assert ext[-1][1] + ext[-1][0] == nz
        assert max(e[0] for e in ext) - min(e[0] for e in ext) <= 1


def test_sass_is_sm100(native):
    """the shipped .so carries sm_100a SASS (cuobjdump) -- skipped when the CUDA toolkit is absent"""
    import shutil
    import subprocess
    from ramsesgpu_b200 import _lib
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("no cuobjdump")
    out = subprocess.run([exe, "-lelf", _lib.LIB_PATH], stdout=subprocess.PIPE).stdout.decode()
    assert "sm_100a" in out
