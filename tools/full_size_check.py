"""Runs the BASELINE.json configurations at their full sizes for a few steps on ONE GPU (GPU box) and
prints device throughput + sanity checks (finite state, conserved mass where the box is periodic).
   python tools/full_size_check.py [case ...]   cases: kh512f32 mri256 mri256slab implode512 implode1024 ot1024 ot512"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ramsesgpu_b200 import HydroRunGodunov, MHDRunGodunov  # noqa: E402
from ramsesgpu_b200.io import ini_override  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def ini_of(name):
    return str(np.load(os.path.join(G, name + ".npz"))["ini"])


QUIET = {"run": {"nstepmax": 1000000, "tend": 1e9, "noutput": -1},
         "output": {"outputVtk": "no", "outputXsm": "no", "outputHdf5": "no"}}
CASES = {
    "kh512f32": (HydroRunGodunov, "kh3d_16x8x16_f32_s10", {"mesh": {"nx": 512, "ny": 512, "nz": 512}}, True),
    "mri256": (MHDRunGodunov, "mri3d_16x32x16_s12", {"mesh": {"nx": 256, "ny": 512, "nz": 256}}, False),
    "mri256slab": (MHDRunGodunov, "mri3d_16x32x16_s12", {"mesh": {"nx": 256, "ny": 512, "nz": 64}}, False),
    "implode512": (HydroRunGodunov, "implode3d_16_s8", {"mesh": {"nx": 512, "ny": 512, "nz": 512}}, False),
    "implode1024": (HydroRunGodunov, "implode3d_16_s8", {"mesh": {"nx": 1024, "ny": 1024, "nz": 1024}}, False),
    "ot1024": (MHDRunGodunov, "ot3d_16_s10", {"mesh": {"nx": 1024, "ny": 1024, "nz": 1024}}, False),
    "ot512": (MHDRunGodunov, "ot3d_16_s10", {"mesh": {"nx": 512, "ny": 512, "nz": 512}}, False),
}
for case in (sys.argv[1:] or ["kh512f32", "mri256", "ot512"]):
    cls, gold, over, f32 = CASES[case]
    ov = dict(QUIET)
    ov.update(over)
    ini = ini_override(ini_of(gold), ov)
    t0 = time.time()
    kw = {"fp32": True} if f32 else {}
    with cls(ini, **kw) as run:
        run.init_simulation()
        run.make_all_boundaries(0)
        s = (0, 0.0, 0.0)
        for _ in range(2):
            s = run.oneStepIntegration(*s)
        run.synchronize()
        steps = 4
        run.profile_begin()
        for _ in range(steps):
            s = run.oneStepIntegration(*s)
        tot, ph = run.profile_end()
        lay = run.layout
        cells = lay.nx * lay.ny * lay.nz
        st = run.stats()
        print("%-12s %dx%dx%d %s: %.1f Mcell-updates/s, %.2f ms/step, dt=%.6g, chunk_planes=%d, device %.1f GB, setup %.0f s | %s"
              % (case, lay.nx, lay.ny, lay.nz, "f32" if f32 else "f64", cells * steps / (tot * 1e-3) / 1e6, tot / steps, s[2],
                 st.chunk_planes, st.device_bytes / 1e9, time.time() - t0,
                 " ".join("%s %.2f" % (k, v[0] / steps) for k, v in ph.items() if v[0] > 0)), flush=True)
        assert np.isfinite(s[2]) and s[2] > 0
