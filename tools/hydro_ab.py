"""A/B timing of the hydro step variants at 512^3 FP32 (BASELINE.json configs[2]) on the GPU box: the fused one-kernel
step (block rows, conservative tiles by TMA or per-thread loads) against trace + flux/update through W; `f64` as argument: 384^3 FP64.
   python tools/hydro_ab.py [f64]"""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from ramsesgpu_b200 import HydroRunGodunov, set_tuning
from ramsesgpu_b200.io import ini_override
G = "tests/golden/"
ini = ini_override(str(np.load(G + "kh3d_16x8x16_f32_s10.npz")["ini"]), {"mesh": {"nx": 512, "ny": 512, "nz": 512},
      "run": {"nstepmax": 1000000, "tend": 1e9, "noutput": -1}, "output": {"outputVtk": "no", "outputXsm": "no", "outputHdf5": "no"}})
F64 = "f64" in sys.argv[1:]
N = 384 if F64 else 512
if F64:
    ini = ini_override(ini, {"mesh": {"nx": N, "ny": N, "nz": N}})
with HydroRunGodunov(ini, fp32=not F64) as run:
    run.init_simulation(); run.make_all_boundaries(0)
    s = (0, 0.0, 0.0)
    for _ in range(3): s = run.oneStepIntegration(*s)
    for rep in range(2):
        for fused, tile, rows, utma in ((0, 1, 0, 0), (1, 1, 12, 0), (1, 1, 12, 1), (1, 1, 16, 0), (1, 1, 16, 1), (1, 1, 20, 0),
                                        (1, 1, 20, 1), (1, 1, 24, 0), (1, 1, 24, 1)):
            set_tuning("hydro_fused", fused)
            set_tuning("hydro_tma", utma)
            set_tuning("hydro_tile", tile)
            set_tuning("hydro_rows", rows)
            for _ in range(2): s = run.oneStepIntegration(*s)
            run.profile_begin()
            for _ in range(5): s = run.oneStepIntegration(*s)
            tot, ph = run.profile_end()
            print("rows=%2d tma=%d hydro_fused=%d hydro_tile=%d total %.3f ms/step %.0f Mcell/s |" % (rows, utma, fused, tile, tot / 5, N**3 * 5 / tot / 1e3), " ".join("%s %.3f" % (k, v[0] / 5) for k, v in ph.items() if v[0] > 0), flush=True)
