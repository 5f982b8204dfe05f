"""ctypes binding of the C oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.  The product (ramsesgpu_b200) never does.
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _params_struct(real):
    class OrcParams(C.Structure):
        _fields_ = [
            ("nStepmax", C.c_int), ("nOutput", C.c_int), ("tEnd", real),
            ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("dim", C.c_int),
            ("nbVar", C.c_int), ("ghostWidth", C.c_int),
            ("isize", C.c_int), ("jsize", C.c_int), ("ksize", C.c_int),
            ("xMin", real), ("xMax", real), ("yMin", real), ("yMax", real), ("zMin", real), ("zMax", real),
            ("dx", real), ("dy", real), ("dz", real),
            ("bc", C.c_int * 6),
            ("mhdEnabled", C.c_int),
            ("cfl", real), ("gamma0", real), ("smallr", real), ("smallc", real), ("smallp", real),
            ("smallpp", real), ("smalle", real), ("gamma6", real), ("cIso", real),
            ("Omega0", real), ("slope_type", real), ("nu", real), ("eta", real),
            ("niter_riemann", C.c_int), ("iorder", C.c_int),
            ("riemannSolver", C.c_int), ("magRiemannSolver", C.c_int),
            ("implementationVersion", C.c_int), ("unsplitVersion", C.c_int),
            ("problem", C.c_char * 64),
            ("ot_direction", C.c_int), ("ot_kt", real),
            ("mri_density", real), ("mri_beta", real), ("mri_amp", real), ("mri_densfluct", real),
            ("mri_seed", C.c_int), ("mri_type", C.c_char * 32),
            ("implode_seed", C.c_int), ("implode_amp", real),
            ("kh_seed", C.c_int), ("kh_p_rand", C.c_int), ("kh_p_sine", C.c_int), ("kh_p_sine_robertson", C.c_int),
            ("kh_amp", real), ("kh_rho_in", real), ("kh_rho_out", real), ("kh_pressure", real),
            ("kh_inner", real), ("kh_outer", real), ("kh_vin", real), ("kh_vout", real),
            ("kh_mode", real), ("kh_w0", real), ("kh_delta", real),
            ("gravityEnabled", C.c_int), ("gravity_x", real), ("gravity_y", real), ("gravity_z", real),
            ("rt_random", C.c_int), ("rt_seed", C.c_int),
            ("rt_amp", real), ("rt_d0", real), ("rt_d1", real), ("rt_bx", real), ("rt_by", real), ("rt_bz", real),
            ("enableJet", C.c_int), ("ijet", C.c_int), ("offsetJet", C.c_int),
            ("djet", real), ("ujet", real), ("pjet", real), ("cjet", real), ("jet_bx", real), ("jet_by", real), ("jet_bz", real),
            ("gravityMode", C.c_int), ("mri_smoothGravity", C.c_int), ("mri_bcFloor", C.c_int), ("mri_zFloor", real),
            ("blast", real * 8),
            ("gresho", real * 5),
            ("riemann2d", real * 2),
            ("riemannConfId", C.c_int),
            ("bubble", real * 7),
            ("kepler", real * 5),
        ]
    return OrcParams


class Oracle:
    """One precision flavour of the C restatement (f64 default, f32 for the FP32 hydro config)."""

    def __init__(self, precision="f64"):
        from . import build as _b  # noqa: local import keeps module import cheap
        _b.build_restatement()
        self.real = C.c_double if precision == "f64" else C.c_float
        self.dtype = np.float64 if precision == "f64" else np.float32
        self.lib = C.CDLL(os.path.join(HERE, "liboracle_%s.so" % precision))
        assert self.lib.orc_sizeof_real() == C.sizeof(self.real)
        self.Params = _params_struct(self.real)
        L, P, R = self.lib, C.POINTER(self.Params), self.real
        RP = C.POINTER(R)
        L.orc_params_from_ini.argtypes = [C.c_char_p, P]; L.orc_params_from_ini.restype = C.c_int
        L.orc_array_len.argtypes = [P]; L.orc_array_len.restype = C.c_long
        L.orc_init_problem.argtypes = [P, RP]; L.orc_init_problem.restype = C.c_int
        L.orc_make_all_boundaries.argtypes = [P, RP]; L.orc_make_all_boundaries.restype = None
        L.orc_compute_dt.argtypes = [P, RP]; L.orc_compute_dt.restype = R
        L.orc_godunov_unsplit.argtypes = [P, RP, RP, R, R]; L.orc_godunov_unsplit.restype = None
        L.orc_step_no_boundaries.argtypes = [P, RP, RP, R]; L.orc_step_no_boundaries.restype = None
        L.orc_make_boundaries.argtypes = [P, RP, C.c_int]; L.orc_make_boundaries.restype = None
        L.orc_run_steps.argtypes = [P, RP, RP, C.c_int, RP, RP]; L.orc_run_steps.restype = C.c_int
        L.orc_riemann_mhd.argtypes = [P, RP, RP, RP]; L.orc_riemann_mhd.restype = None
        L.orc_compute_emf.argtypes = [P, C.c_int, RP, R]; L.orc_compute_emf.restype = R
        L.orc_trace_mhd_3d.argtypes = [P, RP, RP, RP, RP, RP, R, R, R, R, RP, RP, RP]
        L.orc_trace_mhd_3d.restype = None
        L.orc_riemann_hydro.argtypes = [P, RP, RP, RP]; L.orc_riemann_hydro.restype = None
        L.orc_make_all_boundaries_shear.argtypes = [P, RP, R, R]; L.orc_make_all_boundaries_shear.restype = None
        L.orc_set_skip_dissipative.argtypes = [C.c_int]; L.orc_set_skip_dissipative.restype = None
        L.orc_dissipative_stage.argtypes = [P, RP, R, C.c_int]; L.orc_dissipative_stage.restype = None
        L.orc_mhd3d_trace_arrays.argtypes = [P, RP, R, RP]; L.orc_mhd3d_trace_arrays.restype = None
        L.orc_history_mhd3d.argtypes = [P, RP, C.POINTER(C.c_double)]; L.orc_history_mhd3d.restype = None

    # -- helpers ---------------------------------------------------------------------------
    def _p(self, a):
        assert a.dtype == self.dtype and a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(C.POINTER(self.real))

    def params(self, ini_text):
        p = self.Params()
        self.lib.orc_params_from_ini(ini_text.encode(), C.byref(p))
        return p

    def shape(self, p):
        return (p.nbVar, p.ksize, p.jsize, p.isize)

    def alloc(self, p):
        return np.zeros(self.shape(p), dtype=self.dtype)

    def init_problem(self, p):
        U = self.alloc(p)
        rc = self.lib.orc_init_problem(C.byref(p), self._p(U))
        if rc != 0:
            raise RuntimeError("oracle: problem %r not available" % p.problem.decode())
        return U

    def make_all_boundaries(self, p, U):
        self.lib.orc_make_all_boundaries(C.byref(p), self._p(U))

    def make_all_boundaries_shear(self, p, U, dt=0.0, t=0.0):
        """shearing-box ghost fill at time t + dt (MHDRunGodunov.cpp:3763-3793)"""
        self.lib.orc_make_all_boundaries_shear(C.byref(p), self._p(U), dt, t)

    def make_boundaries(self, p, U, idim):
        """idim = 1, 2, 3 (XDIR, YDIR, ZDIR)"""
        self.lib.orc_make_boundaries(C.byref(p), self._p(U), idim)

    def step_no_boundaries(self, p, Uold, Unew, dt):
        self.lib.orc_step_no_boundaries(C.byref(p), self._p(Uold), self._p(Unew), dt)

    def compute_dt(self, p, U):
        return float(self.lib.orc_compute_dt(C.byref(p), self._p(U)))

    def godunov_unsplit(self, p, Uold, Unew, dt, t=0.0):
        self.lib.orc_godunov_unsplit(C.byref(p), self._p(Uold), self._p(Unew), dt, t)

    def run_steps(self, p, U, nsteps):
        """start()-like loop.  Returns (final array, total time, dt trace)."""
        U2 = np.zeros_like(U)
        t = self.real(0)
        dts = np.zeros(max(nsteps, 1), dtype=self.dtype)
        which = self.lib.orc_run_steps(C.byref(p), self._p(U), self._p(U2), nsteps, C.byref(t), self._p(dts))
        return (U2 if which else U), float(t.value), dts[:nsteps]

    def set_skip_dissipative(self, on):
        self.lib.orc_set_skip_dissipative(1 if on else 0)

    def dissipative_stage(self, p, U, dt, stage):
        self.lib.orc_dissipative_stage(C.byref(p), self._p(U), dt, stage)

    def mhd3d_trace_arrays(self, p, U, dt):
        """qm[3], qp[3], qEdge[4][3] of one step: array [18, 8, k, j, i] (see orc_mhd3d_trace_arrays)"""
        out = np.zeros((18, 8, p.ksize, p.jsize, p.isize), dtype=self.dtype)
        self.lib.orc_mhd3d_trace_arrays(C.byref(p), self._p(U), dt, self._p(out))
        return out

    HISTORY_NAMES = ("mass", "maxwell", "reynolds", "magp", "mean_Bx", "mean_By", "mean_Bz", "divB")

    def history_mhd3d(self, p, U):
        out = (C.c_double * 8)()
        self.lib.orc_history_mhd3d(C.byref(p), self._p(U), out)
        return dict(zip(self.HISTORY_NAMES, [float(v) for v in out]))

    def riemann_mhd(self, p, ql, qr):
        ql = np.ascontiguousarray(ql, self.dtype); qr = np.ascontiguousarray(qr, self.dtype)
        f = np.zeros(8, self.dtype)
        self.lib.orc_riemann_mhd(C.byref(p), self._p(ql), self._p(qr), self._p(f))
        return f

    def riemann_hydro(self, p, ql, qr):
        ql = np.ascontiguousarray(ql, self.dtype); qr = np.ascontiguousarray(qr, self.dtype)
        f = np.zeros(5, self.dtype)
        self.lib.orc_riemann_hydro(C.byref(p), self._p(ql), self._p(qr), self._p(f))
        return f

    def compute_emf(self, p, emf_dir, qedge, xpos=0.0):
        qe = np.ascontiguousarray(qedge, self.dtype).reshape(4, 8)
        return float(self.lib.orc_compute_emf(C.byref(p), emf_dir, self._p(qe), xpos))

    def trace_mhd_3d(self, p, q, dq, bfNb, dbf, elec, dtdx, dtdy, dtdz, xpos=0.0):
        a = lambda x, s: np.ascontiguousarray(x, self.dtype).reshape(s)
        q, dq, bfNb, dbf, elec = a(q, 8), a(dq, (3, 8)), a(bfNb, 6), a(dbf, 12), a(elec, (3, 2, 2))
        qm = np.zeros((3, 8), self.dtype); qp = np.zeros((3, 8), self.dtype); qe = np.zeros((4, 3, 8), self.dtype)
        self.lib.orc_trace_mhd_3d(C.byref(p), self._p(q), self._p(dq), self._p(bfNb), self._p(dbf), self._p(elec),
                                  dtdx, dtdy, dtdz, xpos, self._p(qm), self._p(qp), self._p(qe))
        return qm, qp, qe


# ---------------------------------------------------------------------------------------------
# the unmodified reference executable (oracle/_ref/euler_cpu), when it has been built
# ---------------------------------------------------------------------------------------------
def ref_exe(precision="f64"):
    path = os.path.join(HERE, "_ref", "euler_cpu" if precision == "f64" else "euler_cpu_f32")
    return path if os.path.exists(path) else None


def run_reference(ini_text, workdir=None, precision="f64", timeout=None):
    """Runs oracle/_ref/euler_cpu --param <ini> in workdir; returns (stdout, workdir)."""
    exe = ref_exe(precision)
    if exe is None:
        raise FileNotFoundError("oracle/_ref not built (python oracle/build.py)")
    workdir = workdir or tempfile.mkdtemp(prefix="ramses_ref_")
    ini = os.path.join(workdir, "run.ini")
    with open(ini, "w") as f:
        f.write(ini_text)
    out = subprocess.run([exe, "--param", ini], cwd=workdir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         timeout=timeout, check=True).stdout.decode()
    return out, workdir
