"""GPU: checkpoint / resume from the .vti + .meta dump and the history diagnostics reduced on
the device, through the `ramsesgpu_b200_main` executable and the C ABI."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden
from ramsesgpu_b200.io import ini_override

pytestmark = pytest.mark.gpu
MAIN = os.path.join(ROOT, "ramsesgpu_b200", "lib", "ramsesgpu_b200_main")


def run_main(wd, ini):
    p = wd / "run.ini"
    p.write_text(ini)
    r = subprocess.run([MAIN, "--param", str(p)], cwd=str(wd), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
    assert r.returncode == 0, r.stdout.decode()[-2000:]
    return r.stdout.decode()


@pytest.mark.parametrize("name,ghosts,extra", [
    ("ot3d_kt1_16x20x24_s8", "no", {}),                                  # periodic box: inner cells are enough
    ("mri3d_12x20x8_s40", "yes", {}),                                    # shearing box: dump with ghosts
    ("ot3d_diss_16x12x20_s6", "no", {}),                                 # resistivity + viscosity
    ("rt3d_mhd_10x8x24_s8", "yes", {}),                                  # walls + gravity
])
def test_restart_continues_bit_for_bit(native, tmp_path, name, ghosts, extra):
    g = load_golden(name)
    base = ini_override(str(g["ini"]), {"run": {"nstepmax": 8, "noutput": 4, "tend": 1.0e9},
                                        "output": {"ghostIncluded": ghosts, "outputXsm": "no"}})
    prefix = re.search(r"outputPrefix=(\S+)", base).group(1)
    full, resumed = tmp_path / "full", tmp_path / "resumed"
    full.mkdir(); resumed.mkdir()
    run_main(full, base)
    for ext in (".vti", ".vti.meta"):
        (resumed / ("%s_%07d%s" % (prefix, 4, ext))).write_bytes((full / ("%s_%07d%s" % (prefix, 4, ext))).read_bytes())
    meta = (resumed / ("%s_%07d.vti.meta" % (prefix, 4))).read_text()
    assert "nStep 4" in meta and "totalTime 0x" in meta
    out = run_main(resumed, ini_override(base, {"run": {"restart": "yes", "restart_filename": "%s_%07d.vti" % (prefix, 4)}}))
    a = (full / ("%s_%07d.vti" % (prefix, 8))).read_bytes()
    b = (resumed / ("%s_%07d.vti" % (prefix, 8))).read_bytes()
    assert a == b, out[-800:]
    assert (full / ("%s_%07d.vti.meta" % (prefix, 8))).read_text() == (resumed / ("%s_%07d.vti.meta" % (prefix, 8))).read_text()
    assert not (resumed / ("%s_%07d.vti" % (prefix, 0))).exists()


@pytest.mark.parametrize("name,nsteps", [("mri3d_history_12x20x8_s10", 6), ("ot3d_history_16_s8", 5)])
def test_history_device_reduction_vs_oracle(native, oracle64, name, nsteps):
    """same state, two reductions: the device kernels (fixed-order tree) against the oracle's serial loops"""
    from ramsesgpu_b200 import MHDRunGodunov
    g = load_golden(name)
    ini = str(g["ini"])
    with MHDRunGodunov(ini) as run:
        run.init_simulation()
        run.make_all_boundaries(0)
        run.setDataHost(run.getDataHost(0), 1)
        n, t, dt = 0, 0.0, 0.0
        for _ in range(nsteps):
            n, t, dt = run.oneStepIntegration(n, t, dt)
        h = run.history(n)
        U = run.getDataHost(n)
        dx = run.param("dx")
    p = oracle64.params(ini)
    ho = oracle64.history_mhd3d(p, U)
    brms = np.sqrt(2 * ho["magp"])
    vel = np.abs(U[2:4] / U[0]).max()
    scale = {"mass": ho["mass"], "magp": ho["magp"], "maxwell": ho["magp"], "reynolds": ho["mass"] * vel * vel,
             "mean_Bx": brms, "mean_By": brms, "mean_Bz": brms, "divB": brms / dx * U[0].size}
    for k in scale:
        assert abs(h[k] - ho[k]) <= 1e-12 * scale[k], (k, h[k], ho[k], scale[k])


@pytest.mark.parametrize("name", ["mri3d_history_12x20x8_s10", "ot3d_history_16_s8"])
def test_main_writes_reference_history_file(native, tmp_path, name):
    """the history file written by start() against the one written by the unmodified reference (6 digits)"""
    g = load_golden(name)
    ini = str(g["ini"])
    run_main(tmp_path, ini)
    prefix = re.search(r"outputPrefix=(\S+)", ini).group(1)
    lines = (tmp_path / (prefix + "_history.txt")).read_text().splitlines()
    header = [l for l in lines if l.startswith("# totalTime")][0][2:].split()
    assert header == [str(c) for c in g["columns"]]
    table = np.array([[float(x) for x in l.split()] for l in lines if l and not l.startswith("#")])
    ref = g["table"]
    assert table.shape == ref.shape
    col = {c: i for i, c in enumerate(header)}
    for c in ("totalTime", "dt", "mass") + (("magp",) if "magp" in col else ()):
        assert np.allclose(table[:, col[c]], ref[:, col[c]], rtol=6e-6, atol=0), c
    if "magp" in col:
        brms = np.sqrt(2 * ref[:, col["magp"]]).max()
        for c, s in (("maxwell", ref[:, col["magp"]].max()), ("reynolds", 1e-9), ("mean_Bx", brms), ("mean_By", brms), ("mean_Bz", brms)):
            assert np.allclose(table[:, col[c]], ref[:, col[c]], rtol=6e-6, atol=1e-9 * s), c
