"""CPU: the oracle's history diagnostics against the history files written by the unmodified
reference executable (tests/golden/*_history_*.npz, oracle/gen_golden.py).  The reference prints with
6 significant digits; sums of round-off (div B, the MRI stresses of a still laminar flow, the mean
horizontal field) are compared on their natural scale."""
import numpy as np
import pytest

from conftest import load_golden


def history_rows(orc, ini, nsteps):
    """start() loop of the reference (MHDRunGodunov.cpp:3915-3986): history of the state at the top of the
    loop when tHist == 0 or a history time lies in the last step."""
    p = orc.params(ini)
    import re
    dt_hist = np.float32(re.search(r"dtHist\s*=\s*(\S+)", ini).group(1)).astype(np.float64)
    U = orc.init_problem(p)
    shear = p.Omega0 > 0
    if shear:
        orc.make_all_boundaries_shear(p, U, 0.0, 0.0)
    else:
        orc.make_all_boundaries(p, U)
    U2 = U.copy()
    t, n, rows = 0.0, 0, []
    dt = orc.compute_dt(p, U)
    t_hist = 0.0
    while n < nsteps:
        a, b = (U, U2) if n % 2 == 0 else (U2, U)
        if t_hist == 0 or (t - dt <= t_hist + dt_hist and t > t_hist + dt_hist):
            h = orc.history_mhd3d(p, a)
            rows.append((t, dt, h))
            t_hist += dt_hist
        dt = orc.compute_dt(p, a)
        orc.godunov_unsplit(p, a, b, dt, t)
        t += dt
        n += 1
    return rows


@pytest.mark.parametrize("name", ["mri3d_history_12x20x8_s10", "ot3d_history_16_s8"])
def test_history_matches_reference_file(oracle64, name):
    g = load_golden(name)
    orc = oracle64
    rows = history_rows(orc, str(g["ini"]), int(g["steps"]))
    cols = [str(c) for c in g["columns"]]
    table = g["table"]
    assert len(rows) == len(table), (len(rows), len(table))
    for (t, dt, h), ref in zip(rows, table):
        r = dict(zip(cols, ref))
        assert abs(t - r["totalTime"]) <= 6e-6 * max(abs(r["totalTime"]), 1e-300) or t == r["totalTime"]
        assert abs(dt - r["dt"]) <= 6e-6 * r["dt"]
        assert abs(h["mass"] - r["mass"]) <= 6e-6 * r["mass"]
        bscale = np.sqrt(2 * max(h["magp"], 1e-300))   # rms field
        assert abs(h["divB"] - r["divB"]) <= 1e-12 * max(bscale, 1.0) * 1e3 + 6e-6 * abs(r["divB"])
        if "magp" in r:
            assert abs(h["magp"] - r["magp"]) <= 6e-6 * r["magp"]
            for k in ("maxwell", "reynolds", "mean_Bx", "mean_By", "mean_Bz"):
                # laminar start: these are sums that cancel to round-off of their scale
                scale = {"maxwell": r["magp"], "reynolds": 1e-9, "mean_Bx": bscale, "mean_By": bscale, "mean_Bz": bscale}[k]
                assert abs(h[k] - r[k]) <= 6e-6 * abs(r[k]) + 1e-9 * scale, (k, h[k], r[k])
