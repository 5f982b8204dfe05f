"""Generates tests/golden/*.npz from the UNMODIFIED reference executable (oracle/_ref/euler_cpu,
built by oracle/Makefile.ref).  Run here (where /root/reference exists); the small fixtures are
committed and travel to the GPU box.

    python oracle/gen_golden.py
"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle import run_reference  # noqa: E402
from ramsesgpu_b200.io import ini_override, read_vti  # noqa: E402

REF_DATA = os.environ.get("RAMSES_REFERENCE", "/root/reference") + "/data"
OUT = os.path.join(ROOT, "tests", "golden")
VTK_ON = {"outputVtk": "yes", "outputVtkAscii": "no", "outputHdf5": "no", "outputXsm": "no", "outputPng": "no",
          "ghostIncluded": "no", "outputDir": "./"}


def case(name, ini_file, overrides, steps, precision="f64", final_only=False):
    text = open(os.path.join(REF_DATA, ini_file)).read()
    ov = {k: dict(v) for k, v in overrides.items()}
    ov.setdefault("run", {}).update({"nstepmax": steps, "noutput": steps, "tend": 1000.0})
    ov.setdefault("output", {}).update(VTK_ON)
    text = ini_override(text, ov)
    stdout, wd = run_reference(text, precision=precision)
    prefix = re.search(r"outputPrefix=(\S+)", text).group(1)
    fields = read_vti(os.path.join(wd, "%s_%07d.vti" % (prefix, steps)))
    init = read_vti(os.path.join(wd, "%s_%07d.vti" % (prefix, 0)))
    def grab(pattern):
        m = re.search(pattern, stdout)
        return float(m.group(1)) if m else float("nan")
    dt0 = grab(r"Initial dt :\s*(\S+)")
    if dt0 != dt0:  # the hydro driver only prints dt in its step lines
        dt0 = grab(r"step=\s*0 t=\s*\S+ dt=\s*(\S+)")
    ttot = grab(r"DEBUG : totalTime\s*(\S+)")
    dtl = grab(r"DEBUG : dt\s*(\S+)")
    names = list(fields.keys())
    # final_only (the long 64^3 run): the initial state is pinned by the small fixtures of the same problem
    initial = np.zeros((0,)) if final_only else np.stack([init[n] for n in names])
    final = np.stack([fields[n] for n in names])
    extra = {}
    if final_only and final.ndim == 4 and all(np.array_equal(final[:, k], final[:, 0]) for k in range(final.shape[1])):
        # the kt = 0 Orszag-Tang problem is invariant along z and the reference keeps it so bit for bit: one plane stored
        final, extra = final[:, :1].copy(), {"z_invariant": np.array(True)}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), ini=np.array(text), steps=steps, names=np.array(names),
                        final=final, initial=initial, **extra,
                        dt0=dt0, total_time=ttot, dt_last=dtl, precision=np.array(precision))
    print(name, "steps", steps, "dt0", dt0, "t", ttot, "last dt", dtl, "vars", names)


def history_case(name, ini_file, overrides, steps, dt_hist):
    """History diagnostics (MHDRunBase.cpp:3311 / :3476) written by the reference into <prefix>_history.txt:
    the table is frozen together with the ini (values are printed with 6 significant digits)."""
    text = open(os.path.join(REF_DATA, ini_file)).read()
    ov = {k: dict(v) for k, v in overrides.items()}
    ov.setdefault("run", {}).update({"nstepmax": steps, "noutput": steps, "tend": 1.0e6})
    ov.setdefault("output", {}).update(VTK_ON)
    ov.setdefault("history", {}).update({"enabled": "yes", "dtHist": dt_hist, "filename": "history.txt"})
    text = ini_override(text, ov)
    stdout, wd = run_reference(text)
    prefix = re.search(r"outputPrefix=(\S+)", text).group(1)
    lines = open(os.path.join(wd, prefix + "_history.txt")).read().splitlines()
    header = [l for l in lines if l.startswith("# totalTime")][0][2:].split()
    table = np.array([[float(x) for x in l.split()] for l in lines if l and not l.startswith("#")])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), ini=np.array(text), steps=steps, columns=np.array(header), table=table)
    print(name, header, table.shape)
    print(table)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    cases = {
        # BASELINE.json configs[1] at a parity size
        "ot3d_16_s10": ("orszag-tang3d.ini", {"mesh": {"nx": 16, "ny": 16, "nz": 16}}, 10, "f64"),
        "ot3d_24x16x20_s6": ("orszag-tang3d.ini", {"mesh": {"nx": 24, "ny": 16, "nz": 20}}, 6, "f64"),
        # transverse wave number kt=1 makes the problem genuinely three-dimensional
        "ot3d_kt1_16x20x24_s8": ("orszag-tang3d.ini", {"mesh": {"nx": 16, "ny": 20, "nz": 24}, "OrszagTang": {"kt": 1.0}}, 8, "f64"),
        # non-periodic boundaries + HLL / HLLA solvers on the same problem (edge cases)
        "ot3d_16_neumann_hll_s4": ("orszag-tang3d.ini", {
            "mesh": {"nx": 16, "ny": 16, "nz": 16, "boundary_xmin": 2, "boundary_xmax": 2, "boundary_ymin": 1,
                     "boundary_ymax": 1, "boundary_zmin": 2, "boundary_zmax": 1},
            "hydro": {"riemannSolver": "hll"}, "MHD": {"magRiemannSolver": "hlla"}}, 4, "f64"),
        # BASELINE.json configs[0]: 2D MHD Orszag-Tang (implementation 1) at a parity size
        "ot2d_32_s12": ("orszag-tang.ini", {"mesh": {"nx": 32, "ny": 32}}, 12, "f64"),
        "ot2d_40x24_hll_s6": ("orszag-tang.ini", {"mesh": {"nx": 40, "ny": 24}, "hydro": {"riemannSolver": "hll", "slope_type": 1.0}, "MHD": {"magRiemannSolver": "hllf"}}, 6, "f64"),
        # BASELINE.json configs[3]: MRI in the shearing box (rotating frame, isothermal), parity size
        "mri3d_16x32x16_s12": ("mhd_mri_3d.ini", {"mesh": {"nx": 16, "ny": 32, "nz": 16}}, 12, "f64"),
        "mri3d_12x20x8_s40": ("mhd_mri_3d.ini", {"mesh": {"nx": 12, "ny": 20, "nz": 8}}, 40, "f64"),
        # hydro: BASELINE.json configs[4] (implode, approx Riemann solver, Dirichlet walls) and configs[2]
        # (Kelvin-Helmholtz, HLLC, periodic; FP32 like the config, and FP64)
        "implode3d_16_s8": ("implode3d_mpi_zslab.ini", {"mesh": {"nx": 16, "ny": 16, "nz": 16}}, 8, "f64"),
        "implode3d_hll_20x12x16_s5": ("implode3d_mpi_zslab.ini", {"mesh": {"nx": 20, "ny": 12, "nz": 16}, "hydro": {"riemannSolver": "hll", "slope_type": 1.0}}, 5, "f64"),
        "kh3d_16x8x16_f32_s10": ("kelvin_helmholtz_gpu_3d.ini", {"mesh": {"nx": 16, "ny": 8, "nz": 16}}, 10, "f32"),
        "kh3d_16x8x16_f64_s10": ("kelvin_helmholtz_gpu_3d.ini", {"mesh": {"nx": 16, "ny": 8, "nz": 16}}, 10, "f64"),
        # SURVEY 8(f).2 -- dissipative terms: Ohmic resistivity + viscosity on the adiabatic 3D MHD step,
        # on the isothermal shearing box (no resistive energy flux) and viscosity on the hydro step
        "ot3d_diss_16x12x20_s6": ("orszag-tang3d.ini", {"mesh": {"nx": 16, "ny": 12, "nz": 20}, "OrszagTang": {"kt": 1.0},
                                                        "hydro": {"nu": 0.004}, "MHD": {"eta": 0.003}}, 6, "f64"),
        "ot3d_eta_walls_16_s4": ("orszag-tang3d.ini", {
            "mesh": {"nx": 16, "ny": 16, "nz": 16, "boundary_xmin": 2, "boundary_xmax": 2, "boundary_ymin": 1,
                     "boundary_ymax": 1, "boundary_zmin": 2, "boundary_zmax": 1}, "MHD": {"eta": 0.005}}, 4, "f64"),
        "mri3d_diss_12x20x8_s10": ("mhd_mri_3d.ini", {"mesh": {"nx": 12, "ny": 20, "nz": 8},
                                                      "hydro": {"nu": 2e-6}, "MHD": {"eta": 1e-6}}, 10, "f64"),
        "implode3d_visc_16_s6": ("implode3d_mpi_zslab.ini", {"mesh": {"nx": 16, "ny": 16, "nz": 16}, "hydro": {"nu": 0.002}}, 6, "f64"),
        "kh3d_visc_16x8x16_f32_s6": ("kelvin_helmholtz_gpu_3d.ini", {"mesh": {"nx": 16, "ny": 8, "nz": 16}, "hydro": {"nu": 0.001}}, 6, "f32"),
        # SURVEY 8(f).2 -- static gravity: Rayleigh-Taylor, hydro (approx solver) and MHD (HLLD), z walls
        "rt3d_hydro_10x8x24_s8": ("rayleigh_taylor_gpu_3d.ini", {"mesh": {"nx": 10, "ny": 8, "nz": 24}}, 8, "f64"),
        "rt3d_mhd_10x8x24_s8": ("rayleigh_taylor_gpu_3d_mhd.ini", {"mesh": {"nx": 10, "ny": 8, "nz": 24}}, 8, "f64"),
        "rt3d_mhd_visc_rand_8x10x16_s5": ("rayleigh_taylor_gpu_3d_mhd.ini", {
            "mesh": {"nx": 8, "ny": 10, "nz": 16}, "hydro": {"nu": 0.001}, "MHD": {"eta": 0.002},
            "gravity": {"static_field_x": 0.05, "static_field_z": -0.3},
            "rayleigh-taylor": {"randomEnabled": "yes", "random_seed": 7, "bx": 0.05, "bz": 0.02}}, 5, "f64"),
        # SURVEY 8(f).4 -- further MHD test problems of the reference on the same step kernels
        "briowu2d_32x24_s8": ("mhd_BrioWu.ini", {"mesh": {"nx": 32, "ny": 24}}, 8, "f64"),
        "briowu2d_diag_24_s6": ("mhd_BrioWu.ini", {"mesh": {"nx": 24, "ny": 24}, "BrioWu": {"direction": 3}}, 6, "f64"),
        "briowu3d_z_10x8x16_s5": ("mhd_BrioWu.ini", {"mesh": {"nx": 10, "ny": 8, "nz": 16}, "BrioWu": {"direction": 2}}, 5, "f64"),
        "briowu3d_xyz_12_s5": ("mhd_BrioWu.ini", {"mesh": {"nx": 12, "ny": 12, "nz": 12}, "BrioWu": {"direction": 3}}, 5, "f64"),
        # (the reference run of this problem blows up after ~5 steps at any resolution, incl. the shipped 128^2: 3 steps)
        "rotor2d_48_s1": ("mhd_rotor.ini", {"mesh": {"nx": 48, "ny": 48}, "MHD": {"implementationVersion": 1}}, 1, "f64"),
        "fieldloop2d_32x20_s8": ("mhd_fieldloop2d.ini", {"mesh": {"nx": 32, "ny": 20}}, 8, "f64"),
        "fieldloop3d_16x12x10_s6": ("mhd_fieldloop3d.ini", {"mesh": {"nx": 16, "ny": 12, "nz": 10}}, 6, "f64"),
        "currentsheet2d_24_s8": ("mhd_currentSheet_2d.ini", {"mesh": {"nx": 24, "ny": 24}}, 8, "f64"),
        "currentsheet3d_16x16x8_s5": ("mhd_currentSheet_3d.ini", {"mesh": {"nx": 16, "ny": 16, "nz": 8}}, 5, "f64"),
        # jet inflow through the lower z (3D) / y (2D) ghost cells: boundary patch + dt limit
        "jet3d_hydro_14x14x20_s8": ("jet3d_gpu.ini", {"mesh": {"nx": 14, "ny": 14, "nz": 20}, "jet": {"ijet": 4, "offsetJet": 5}}, 8, "f64"),
        "jet3d_mhd_15x15x20_s8": ("mhd_jet3d.ini", {"mesh": {"nx": 15, "ny": 15, "nz": 20}, "MHD": {"implementationVersion": 4},
                                                     "jet": {"BStatic_z": 0.5, "BStatic_x": 0.1}}, 8, "f64"),
        "jet2d_mhd_24x32_s10": ("mhd_jet2d.ini", {"mesh": {"nx": 24, "ny": 32}, "MHD": {"implementationVersion": 1}, "jet": {"ijet": 4, "offsetJet": 10}}, 10, "f64"),
        # stratified shearing box: vertical gravity, z-stratified boundaries (slope_type 2: the reference's CPU
        # rotating step leaves the slopes uninitialised for slope_type 3, MHDRunGodunov.cpp:2631-2636)
        "mri3d_strat_8x12x24_s10": ("mhd_mri_3d_stratified.ini", {"mesh": {"nx": 8, "ny": 12, "nz": 24}, "hydro": {"slope_type": 2.0}}, 10, "f64"),
        "ot3d_slope3_16x12x20_s6": ("orszag-tang3d.ini", {"mesh": {"nx": 16, "ny": 12, "nz": 20}, "OrszagTang": {"kt": 1.0}, "hydro": {"slope_type": 3.0}}, 6, "f64"),
        # 2D hydro (oracle ahead of the CUDA path): implosion (approx solver, minmod), jet, blast (HLLC, MC slopes)
        "implode2d_32_s10": ("implode2d.ini", {"mesh": {"nx": 32, "ny": 32}}, 10, "f64"),
        "jet2d_hydro_24x32_s10": ("jet2d_cpu.ini", {"mesh": {"nx": 24, "ny": 32}, "jet": {"ijet": 4, "offsetJet": 10}}, 10, "f64"),
        "blast2d_hllc_32_s8": ("blast2d.ini", {"mesh": {"nx": 32, "ny": 32}}, 8, "f64"),
        "blast3d_hllc_16x12x20_s8": ("blast2d.ini", {"mesh": {"nx": 16, "ny": 12, "nz": 20}, "blast": {"center_z": 0.4}}, 8, "f64"),
        "khmhd2d_24x32_s8": ("mhd_kelvin_helmholtz_2d.ini", {"mesh": {"nx": 24, "ny": 32}}, 8, "f64"),
        "khmhd3d_12x16x8_s5": ("mhd_kelvin_helmholtz_2d.ini", {"mesh": {"nx": 12, "ny": 16, "nz": 8}, "MHD": {"implementationVersion": 4}}, 5, "f64"),
        "shearwave3d_16x12x8_s10": ("mhd_shearWave_3d.ini", {"mesh": {"nx": 16, "ny": 12, "nz": 8}}, 10, "f64"),
        # SURVEY 8(f).4 -- further hydro test problems of the reference: Sod tube (Dirichlet walls), Gresho vortex
        # (periodic, libm sin / cos / atan2 / log in the initial state), Lax-Liu 2D Riemann configurations (Neumann)
        "sod2d_32x24_s8": ("hydro_sod2d.ini", {"mesh": {"nx": 32, "ny": 24}}, 8, "f64"),
        "sod3d_16x12x10_s6": ("hydro_sod2d.ini", {"mesh": {"nx": 16, "ny": 12, "nz": 10}}, 6, "f64"),
        "gresho2d_32_s8": ("Gresho_vortex2d.ini", {"mesh": {"nx": 32, "ny": 32}, "hydro": {"unsplitVersion": 1}}, 8, "f64"),
        "gresho3d_16x16x8_s5": ("Gresho_vortex2d.ini", {"mesh": {"nx": 16, "ny": 16, "nz": 8}, "hydro": {"unsplitVersion": 1},
                                                        "Gresho_vortex": {"v_bulk_z": 0.25}}, 5, "f64"),
        "riemann2d_c2_32_s8": ("riemann2d.ini", {"mesh": {"nx": 32, "ny": 32}}, 8, "f64"),
        # 2D Kelvin-Helmholtz (kelvin_helmholtz_cpu_2d.ini / _gpu_2d.ini): the four perturbation types of the 2D branch
        "kh2d_rand_32_s8": ("kelvin_helmholtz_cpu_2d.ini", {"mesh": {"nx": 32, "ny": 32}}, 8, "f64"),
        "kh2d_robertson_32x40_s8": ("kelvin_helmholtz_gpu_2d.ini", {"mesh": {"nx": 32, "ny": 40}, "hydro": {"unsplitVersion": 1}}, 8, "f64"),
        "kh2d_athena_40x32_s6": ("kelvin_helmholtz_cpu_2d.ini", {"mesh": {"nx": 40, "ny": 32, "ymin": -0.5, "ymax": 0.5},
                                 "kelvin-helmholtz": {"perturbation_rand": "no", "perturbation_sine_athena": "yes"}}, 6, "f64"),
        "kh2d_sine_32x48_s6": ("kelvin_helmholtz_cpu_2d.ini", {"mesh": {"nx": 32, "ny": 48}, "kelvin-helmholtz": {
            "perturbation_rand": "no", "perturbation_sine": "yes", "inner_size": 0.1, "outer_size": 0.3}}, 6, "f64"),
        # 2D hydro with static gravity: Rayleigh-Taylor (rayleigh_taylor_gpu_2d.ini), single mode and rand()
        "rt2d_hydro_16x48_s10": ("rayleigh_taylor_gpu_2d.ini", {"mesh": {"nx": 16, "ny": 48}}, 10, "f64"),
        "rt2d_hydro_rand_24x40_s8": ("rayleigh_taylor_gpu_2d.ini", {"mesh": {"nx": 24, "ny": 40}, "gravity": {"static_field_x": 0.02},
                                     "rayleigh-taylor": {"randomEnabled": "yes", "random_seed": 5}, "hydro": {"riemannSolver": "hllc"}}, 8, "f64"),
        # inertial wave in the rotating frame (periodic box, isothermal), 3D: mhd_inertialWave_2d.ini with nz > 1
        "inertialwave3d_12x16x8_s12": ("mhd_inertialWave_2d.ini", {"mesh": {"nx": 12, "ny": 16, "nz": 8},
                                       "MHD": {"implementationVersion": 4, "omega0": 0.3}, "hydro": {"cIso": 0.05}}, 12, "f64"),
        # 2D hydro with a gravity FIELD: Keplerian disc around a softened point mass (Keplerian_disk2d.ini)
        "kepler2d_32_s10": ("Keplerian_disk2d.ini", {"mesh": {"nx": 32, "ny": 32}}, 10, "f64"),
        "bubble2d_24x32_s10": ("falling_bubble_gpu_2d.ini", {"mesh": {"nx": 24, "ny": 32}, "falling-bubble": {"center_y": 0.7}}, 10, "f64"),
        "riemann2d_c5_40x24_s6": ("riemann2d.ini", {"mesh": {"nx": 40, "ny": 24}, "hydro": {"riemann_config_number": 5},
                                                    "riemann2d": {"x": 0.5, "y": 0.45}}, 6, "f64"),
    }
    for name, (ini, ov, steps, prec) in cases.items():
        if only and name not in only:
            continue
        case(name, ini, ov, steps, prec)
    # the north star's parity statement at the survey's parity size: Orszag-Tang 3D, 64^3, 100 steps (about 100 s of
    # the reference on one core; only generated when asked for by name: `python oracle/gen_golden.py ot3d_64_s100`)
    long_cases = {
        "ot3d_64_s100": ("orszag-tang3d.ini", {"mesh": {"nx": 64, "ny": 64, "nz": 64}}, 100, "f64"),
    }
    for name, (ini, ov, steps, prec) in long_cases.items():
        if name in only:
            case(name, ini, ov, steps, prec, final_only=True)
    hist = {
        # SURVEY 8(f).3 -- history files of the reference: MRI stresses, and mass / div B of Orszag-Tang
        "mri3d_history_12x20x8_s10": ("mhd_mri_3d.ini", {"mesh": {"nx": 12, "ny": 20, "nz": 8}}, 10, 24.0),
        "ot3d_history_16_s8": ("orszag-tang3d.ini", {"mesh": {"nx": 16, "ny": 16, "nz": 16}, "OrszagTang": {"kt": 1.0}}, 8, 0.0045),
    }
    for name, (ini, ov, steps, dth) in hist.items():
        if only and name not in only:
            continue
        history_case(name, ini, ov, steps, dth)
