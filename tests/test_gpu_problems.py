"""GPU: the further test problems of the reference (Brio-Wu shock tube, field-loop advection,
current sheet, magnetised Kelvin-Helmholtz, shear wave in the shearing box, jets; hydro: blast, Sod tube, Gresho vortex,
Lax-Liu 2D Riemann problems; 2D and 3D, outflow, wall and periodic boundaries) on the same CUDA step kernels, through the C ABI,
against golden vectors from the unmodified reference executable."""
import numpy as np
import pytest

from conftest import TOL_F64, load_golden

pytestmark = pytest.mark.gpu

CASES = ["briowu2d_32x24_s8", "briowu2d_diag_24_s6", "briowu3d_z_10x8x16_s5", "briowu3d_xyz_12_s5",
         "fieldloop2d_32x20_s8", "fieldloop3d_16x12x10_s6", "currentsheet2d_24_s8", "currentsheet3d_16x16x8_s5",
         "khmhd2d_24x32_s8", "khmhd3d_12x16x8_s5", "shearwave3d_16x12x8_s10",
         # jet inflow boundary patch + dt limit (hydro 3D, MHD 3D, MHD 2D)
         "jet3d_hydro_14x14x20_s8", "jet3d_mhd_15x15x20_s8", "jet2d_mhd_24x32_s10",
         # inertial wave: rotating frame in a periodic box (no shearing-box borders), isothermal
         "inertialwave3d_12x16x8_s12",
         # further hydro problems (2D kernels and the fused 3D hydro kernel): spherical blast, Sod tube (Dirichlet walls),
         # Gresho vortex (periodic), Lax-Liu 2D Riemann configurations 3 and 6 (Neumann)
         "blast3d_hllc_16x12x20_s8", "sod2d_32x24_s8", "sod3d_16x12x10_s6", "gresho2d_32_s8", "gresho3d_16x16x8_s5",
         "riemann2d_c2_32_s8", "riemann2d_c5_40x24_s6",
         # 2D Kelvin-Helmholtz, the four perturbation types of the reference's 2D branch
         "kh2d_rand_32_s8", "kh2d_robertson_32x40_s8", "kh2d_athena_40x32_s6", "kh2d_sine_32x48_s6",
         # 2D hydro with static gravity: Rayleigh-Taylor, single mode and rand() perturbation
         "rt2d_hydro_16x48_s10", "rt2d_hydro_rand_24x40_s8", "bubble2d_24x32_s10",
         # ... and a gravity FIELD (one vector per cell): Keplerian disc around a softened point mass
         "kepler2d_32_s10"]


@pytest.mark.parametrize("name", CASES)
def test_golden_reference_run(native, name):
    from ramsesgpu_b200 import HydroRunGodunov, MHDRunGodunov
    g = load_golden(name)
    Run = MHDRunGodunov if len(g["names"]) == 8 else HydroRunGodunov
    with Run(str(g["ini"])) as run:
        run.init_simulation()
        run.make_all_boundaries(0)
        run.setDataHost(run.getDataHost(0), 1)
        n, t, dt = 0, 0.0, 0.0
        for _ in range(int(g["steps"])):
            n, t, dt = run.oneStepIntegration(n, t, dt)
        U = run.getDataHost(n)
        gw, dim = run.layout.ghost_width, run.layout.dim
        isothermal = run.param("ciso") > 0
    got = U[:, 0, gw:-gw, gw:-gw] if dim == 2 else U[:, gw:-gw, gw:-gw, gw:-gw]
    ref = g["final"]
    mom = np.sqrt(sum(float(np.sum(ref[v] ** 2)) for v in (2, 3, 4) if v < min(len(ref), 5) and (len(ref) != 4 or v < 4)))
    mag = np.sqrt(sum(float(np.sum(ref[v] ** 2)) for v in (5, 6, 7))) if len(ref) == 8 else 1.0
    for v, vname in enumerate(g["names"]):
        # vector components against the norm of their vector field (components that stay ~0 by symmetry)
        norm = np.sqrt(np.sum(ref[v] ** 2)) if v < 2 else (mom if v < 5 else mag)
        err = np.sqrt(np.sum((ref[v] - got[v]) ** 2)) / max(norm, 1e-300)
        # isothermal runs (shear wave): the energy is carried along by the fluxes but never read (p = rho cIso^2);
        # it starts at 0 and is a sum of cancelling flux differences, measured 2e-12 against the reference
        tol = 1e-10 if (isothermal and v == 1) else TOL_F64
        if isothermal and v == 1 and norm < 1e-12 * np.sqrt(np.sum(ref[0] ** 2)):
            # (inertial wave: the unused energy stays at the round-off residue of cancelling fluxes, 1e-18 of the density:
            # it has no digits to compare; bounded instead)
            assert np.sqrt(np.sum(got[v] ** 2)) < 1e-12 * np.sqrt(np.sum(ref[0] ** 2)), (name, vname)
            continue
        assert err < tol, (name, vname, err)
    if g["total_time"] == g["total_time"]:   # the hydro driver of the reference does not print these
        assert abs(t - g["total_time"]) < 1e-10 * g["total_time"]
        assert abs(dt - g["dt_last"]) < 1e-10 * g["dt_last"]


def test_rotating_frame_in_2d_is_refused_not_ignored(native):
    """[MHD] omega0 > 0 on a 2D grid (mhd_inertialWave_2d.ini as shipped): the reference has a 2D rotating step that is not
    built here; the step must refuse instead of silently running the non-rotating 2D solver"""
    from ramsesgpu_b200 import MHDRunGodunov
    from ramsesgpu_b200._lib import RgError
    from ramsesgpu_b200.io import ini_override
    ini = ini_override(str(load_golden("inertialwave3d_12x16x8_s12")["ini"]), {"mesh": {"nz": 1}})
    with MHDRunGodunov(ini) as run:
        run.init_simulation()
        assert run.layout.dim == 2
        with pytest.raises(RgError, match="rotating frame"):
            run.oneStepIntegration(0, 0.0, 0.0)
