"""Writes the parameter files of the compute-sanitizer pass (tools/sanitize.sh) from the committed golden fixtures:
    python tools/sanitize_cases.py <directory>
One small case per kernel family, file output off, a few steps; names end in _f32 where the executable takes --fp32."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ramsesgpu_b200.io import ini_override  # noqa: E402

# golden fixture -> (steps, mesh override): sizes that open several tiles / remainder tiles of the fused kernels
CASES = {
    "ot3d_16_s10": (3, {"nx": 40, "ny": 20, "nz": 24}),              # fused trace + fused flux/emf/update, 3 x 3 tiles
    "ot3d_slope3_16x12x20_s6": (2, None),                            # separate kernels (27-point slopes)
    "ot3d_16_neumann_hll_s4": (2, None),                             # generic Riemann pair, walls
    "ot3d_diss_16x12x20_s6": (2, None),                              # resistivity + viscosity
    "mri3d_16x32x16_s12": (3, {"nx": 24, "ny": 40, "nz": 16}),       # rotating fused kernel, border strips, y remap
    "mri3d_strat_8x12x24_s10": (3, None),                            # stratified box: per-plane gravity, z ghost kernel
    "kh3d_16x8x16_f32_s10": (3, {"nx": 68, "ny": 40, "nz": 24}),     # one-kernel hydro step, FP32, TMA tiles
    "kh3d_16x8x16_f64_s10": (3, {"nx": 56, "ny": 30, "nz": 24}),     # ... FP64
    "implode3d_16_s8": (3, {"nx": 36, "ny": 30, "nz": 40}),          # ... walls
    "implode3d_visc_16_s6": (2, None),                               # two-kernel hydro path + viscosity
    "ot2d_32_s12": (3, None),                                        # 2D MHD
    "jet2d_hydro_24x32_s10": (3, None),                              # 2D hydro
    "jet3d_mhd_15x15x20_s8": (2, None),                              # odd sizes (no TMA: row pitch), jet inflow
    "kepler2d_32_s10": (3, None),                                    # 2D hydro with a per-cell gravity field
    "rt2d_hydro_16x48_s10": (3, None),                               # 2D hydro, uniform gravity, walls in y
    "inertialwave3d_12x16x8_s12": (3, {"nx": 24, "ny": 24, "nz": 16}),  # rotating frame, periodic box (no border strips)
}


def main(out):
    os.makedirs(out, exist_ok=True)
    for name, (steps, mesh) in CASES.items():
        g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"), allow_pickle=True)
        ov = {"run": {"nstepmax": steps, "noutput": -1}, "output": {"outputVtk": "no", "outputDir": out}}
        if mesh:
            ov["mesh"] = mesh
        suffix = "_f32" if "precision" in g.files and str(g["precision"]) == "f32" else ""
        with open(os.path.join(out, name + suffix + ".ini"), "w") as fh:
            fh.write(ini_override(str(g["ini"]), ov))
    print("wrote %d parameter files to %s" % (len(CASES), out))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/tmp/rg_sanitize")
