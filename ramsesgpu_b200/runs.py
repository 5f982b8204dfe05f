"""Python mirror of the reference's run classes (HydroRunBase / MHDRunBase / HydroRunGodunov /
MHDRunGodunov; reference src/hydro/HydroRunBase.h:63-639, MHDRunGodunov.h) over the C ABI.
Method names and argument meaning are the reference's; every call runs the CUDA path."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import RgLayout, RgStats, check


class HydroRunBase:
    """`configMap` is the TEXT of a parameter file (or a path with from_file=True)."""

    def __init__(self, configMap, fp32=False, from_file=False, rank=0, nranks=1, nccl_unique_id=None, device=-1):
        self._L = _lib.load()
        self._h = C.c_void_p()
        flags = _lib.RG_FLAG_FP32 if fp32 else 0
        if nranks > 1:
            if from_file:
                configMap = open(configMap).read()
            buf = C.create_string_buffer(bytes(nccl_unique_id), 128)
            check(self._L.rg_create_distributed(configMap.encode(), flags, rank, nranks, buf, device, C.byref(self._h)))
        elif from_file:
            check(self._L.rg_create_from_file(configMap.encode(), flags, C.byref(self._h)))
        else:
            check(self._L.rg_create(configMap.encode(), flags, C.byref(self._h)))
        self.layout = RgLayout()
        check(self._L.rg_get_layout(self._h, C.byref(self.layout)))
        self.dtype = np.float32 if self.layout.real_bytes == 4 else np.float64
        self.shape = (self.layout.nvar, self.layout.ksize, self.layout.jsize, self.layout.isize)

    # -- lifetime --------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.rg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- the reference's operator surface ---------------------------------------------------------
    def init_simulation(self, problem=""):
        n = C.c_int(0)
        check(self._L.rg_init_simulation(self._h, problem.encode(), C.byref(n)))
        return n.value

    def make_all_boundaries(self, which=0):
        check(self._L.rg_make_all_boundaries(self._h, which))

    def compute_dt(self, useU=0):
        dt = C.c_double(0)
        check(self._L.rg_compute_dt(self._h, useU, C.byref(dt)))
        return dt.value

    def godunov_unsplit(self, nStep, dt):
        check(self._L.rg_godunov_unsplit(self._h, nStep, dt))

    def oneStepIntegration(self, nStep, t, dt=0.0):
        """Returns (nStep, t, dt) after the step (the reference takes them by reference)."""
        n, tt, d = C.c_int(nStep), C.c_double(t), C.c_double(dt)
        check(self._L.rg_one_step(self._h, C.byref(n), C.byref(tt), C.byref(d)))
        return n.value, tt.value, d.value

    def start(self):
        check(self._L.rg_run(self._h))

    def output(self, nStep):
        check(self._L.rg_output(self._h, nStep))

    HISTORY_NAMES = ("mass", "maxwell", "reynolds", "magp", "mean_Bx", "mean_By", "mean_Bz", "divB")

    def history(self, nStep):
        """History diagnostics of the state of step nStep (reference MHDRunBase::history_default / history_mri),
        reduced on the device, global over all slabs."""
        out = (C.c_double * 8)()
        check(self._L.rg_history(self._h, nStep, out))
        return dict(zip(self.HISTORY_NAMES, [float(v) for v in out]))

    # -- data ------------------------------------------------------------------------------------
    def getData(self, nStep=0):
        """Device pointer (int) of the buffer holding step nStep."""
        p = C.c_void_p()
        check(self._L.rg_get_data_device(self._h, nStep % 2, C.byref(p)))
        return p.value

    def getDataHost(self, nStep=0):
        """copyGpuToCpu(nStep) + getDataHost(nStep): ndarray [var, k, j, i] of the local slab."""
        out = np.empty(self.shape, dtype=self.dtype)
        check(self._L.rg_copy_to_host(self._h, nStep % 2, out.ctypes.data, out.nbytes))
        return out

    def setDataHost(self, array, which=0):
        a = np.ascontiguousarray(array, dtype=self.dtype)
        assert a.shape == self.shape, (a.shape, self.shape)
        check(self._L.rg_copy_from_host(self._h, which, a.ctypes.data, a.nbytes))

    def steps_from_host(self, host_in, host_out, nsteps):
        """H2D(host_in) -> nsteps steps -> D2H(host_out); returns (t, last dt)."""
        t, dt = C.c_double(0), C.c_double(0)
        check(self._L.rg_steps_from_host(self._h, host_in.ctypes.data, host_out.ctypes.data, host_in.nbytes,
                                         nsteps, C.byref(t), C.byref(dt)))
        return t.value, dt.value

    def steps_from_host_batch(self, host_in, host_out):
        """Independent one-step jobs host_out[j] = step(host_in[j]) with copy/compute overlap
        (rg_steps_from_host_batch); returns the dt of every job."""
        n = len(host_in)
        assert len(host_out) == n
        ins = (C.c_void_p * n)(*[a.ctypes.data for a in host_in])
        outs = (C.c_void_p * n)(*[a.ctypes.data for a in host_out])
        dts = (C.c_double * n)()
        check(self._L.rg_steps_from_host_batch(self._h, n, ins, outs, host_in[0].nbytes, dts))
        return list(dts)

    def synchronize(self):
        check(self._L.rg_synchronize(self._h))

    def param(self, name):
        v = C.c_double(0)
        check(self._L.rg_get_param(self._h, name.encode(), C.byref(v)))
        return v.value

    def stats(self):
        s = RgStats()
        check(self._L.rg_get_stats(self._h, C.byref(s)))
        return s

    def profile_begin(self):
        check(self._L.rg_profile_begin(self._h))

    def profile_end(self):
        """Returns (total_ms, {phase: (ms, launches)}) for the region since profile_begin()."""
        n = len(_lib.PHASES)
        tot, ms, cnt = C.c_double(0), (C.c_double * n)(), (C.c_ulonglong * n)()
        check(self._L.rg_profile_end(self._h, C.byref(tot), ms, cnt))
        return tot.value, {name: (ms[i], cnt[i]) for i, name in enumerate(_lib.PHASES)}

    def set_halo_overlap(self, on):
        check(self._L.rg_set_halo_overlap(self._h, 1 if on else 0))

    def set_chunk_planes(self, planes):
        check(self._L.rg_set_chunk_planes(self._h, planes))

    def inner(self, U):
        g = self.layout.ghost_width
        if self.layout.dim == 2:
            return U[:, 0, g:-g, g:-g]
        return U[:, g:-g, g:-g, g:-g]

    # -- device probes ---------------------------------------------------------------------------
    def probe_riemann_mhd(self, ql, qr):
        """riemann_mhd (MHD handle, 8-component states) or riemann<NVAR_3D> (hydro handle, 5)."""
        nv = 8 if self.layout.mhd else 5
        ql = np.ascontiguousarray(ql, self.dtype).reshape(-1, nv)
        qr = np.ascontiguousarray(qr, self.dtype).reshape(-1, nv)
        f = np.empty_like(ql)
        check(self._L.rg_probe_riemann_mhd(self._h, ql.shape[0], ql.ctypes.data, qr.ctypes.data, f.ctypes.data))
        return f

    def probe_compute_emf(self, emf_dir, q_edge, x_pos=None):
        q = np.ascontiguousarray(q_edge, self.dtype).reshape(-1, 4, 8)
        e = np.empty(q.shape[0], self.dtype)
        xp = None if x_pos is None else np.ascontiguousarray(x_pos, self.dtype)
        check(self._L.rg_probe_compute_emf(self._h, q.shape[0], emf_dir, q.ctypes.data,
                                           None if xp is None else xp.ctypes.data, e.ctypes.data))
        return e


class MHDRunBase(HydroRunBase):
    def compute_dt_mhd(self, useU=0):
        return self.compute_dt(useU)


class HydroRunGodunov(HydroRunBase):
    def probe_riemann(self, ql, qr):
        return self.probe_riemann_mhd(ql, qr)


class MHDRunGodunov(MHDRunBase):
    pass


class PinnedArray:
    """numpy view of page-locked host memory (rg_alloc_pinned); free() or garbage collection releases it."""

    def __init__(self, shape, dtype=np.float64):
        self._L = _lib.load()
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        check(self._L.rg_alloc_pinned(n, C.byref(p)))
        self._p = p
        self.array = np.frombuffer((C.c_char * n).from_address(p.value), dtype=dtype).reshape(shape)

    def free(self):
        if self._p is not None:
            self.array = None
            self._L.rg_free_pinned(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def set_tuning(key, value):
    check(_lib.load().rg_set_tuning(key.encode(), int(value)))


def reset_launch_count():
    _lib.load().rg_reset_launch_count()


def initial_condition_host(ini_text, fp32=False, rank=0, nranks=1):
    """Initial condition of a z-slab computed on the HOST only (rg_initial_condition_host): ndarray [var, k, j, i]
    with ghosts, and the layout.  Works without a GPU."""
    import numpy as np
    L = _lib.load()
    lay = _lib.RgLayout()
    flags = 1 if fp32 else 0
    check(L.rg_initial_condition_host(ini_text.encode(), flags, rank, nranks, None, 0, C.byref(lay)))
    U = np.zeros((lay.nvar, lay.ksize, lay.jsize, lay.isize), dtype=np.float32 if fp32 else np.float64)
    check(L.rg_initial_condition_host(ini_text.encode(), flags, rank, nranks, U.ctypes.data_as(C.c_void_p), U.nbytes, C.byref(lay)))
    return U, lay


def slab_extent(nz_global, nranks, rank):
    a, b = C.c_int(0), C.c_int(0)
    check(_lib.load().rg_slab_extent(nz_global, nranks, rank, C.byref(a), C.byref(b)))
    return a.value, b.value
