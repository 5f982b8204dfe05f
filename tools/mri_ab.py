"""A/B timing of the rotating-frame step on the config-4 slab (mhd_mri_3d.ini at 256 x 512 x 64, GPU box): separate
flux / emf / update kernels against the fused kernel, with the next dt reduced in the kernel or by k_invdt.
   python tools/mri_ab.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from ramsesgpu_b200 import MHDRunGodunov, set_tuning  # noqa: E402
from ramsesgpu_b200.io import ini_override  # noqa: E402

ini = ini_override(str(np.load(os.path.join(ROOT, "tests", "golden", "mri3d_16x32x16_s12.npz"))["ini"]),
                   {"mesh": {"nx": 256, "ny": 512, "nz": 64}, "run": {"nstepmax": 1000000, "tend": 1e9, "noutput": -1},
                    "output": {"outputVtk": "no", "outputXsm": "no", "outputHdf5": "no"}})
cells = 256 * 512 * 64
for rep in range(2):
    for fused, rotdt in ((0, 1), (1, 1), (1, 0)):
        set_tuning("fused_b", fused)
        set_tuning("rot_dt", rotdt)
        with MHDRunGodunov(ini) as run:
            run.init_simulation()
            run.make_all_boundaries(0)
            s = (0, 0.0, 0.0)
            for _ in range(3):
                s = run.oneStepIntegration(*s)
            run.profile_begin()
            for _ in range(10):
                s = run.oneStepIntegration(*s)
            tot, ph = run.profile_end()
        print("fused_b=%d rot_dt=%d total %.3f ms/step %.0f Mcell/s |" % (fused, rotdt, tot / 10, cells * 10 / tot / 1e3),
              " ".join("%s %.3f" % (k, v[0] / 10) for k, v in ph.items() if v[0] > 0), flush=True)
set_tuning("fused_b", 1)
set_tuning("rot_dt", 0)
