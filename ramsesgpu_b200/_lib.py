"""ctypes binding of include/ramsesgpu_b200.h.  The native library is REQUIRED: importing the
compute API without it raises (there is no Python/CPU fallback)."""
import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
# RG_LIB_PATH: an experimental flavour built with RG_VARIANT (see build.py), for A/B timing only
LIB_PATH = os.environ.get("RG_LIB_PATH") or os.path.join(PKG, "lib", "libramsesgpu_b200.so")


class RgLayout(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "nx", "ny", "nz", "isize", "jsize", "ksize", "nvar", "ghost_width", "dim", "mhd", "real_bytes",
        "nz_local", "k_offset", "rank", "nranks")]


class RgStats(C.Structure):
    _fields_ = [("kernel_launches", C.c_ulonglong), ("last_step_ms", C.c_double),
                ("halo_bytes_per_step", C.c_double), ("device_bytes", C.c_size_t), ("chunk_planes", C.c_int),
                ("halo_peer_copies", C.c_int)]


RG_FLAG_FP32 = 1
PHASES = ["boundary", "prim", "trace", "flux", "emf", "update", "dt", "copy", "halo", "fused", "diss"]
RG_OK, RG_ERR_INVALID, RG_ERR_NO_DEVICE, RG_ERR_CUDA, RG_ERR_UNSUPPORTED, RG_ERR_IO, RG_ERR_NCCL = range(7)  # include/ramsesgpu_b200.h

# name -> (restype, argtypes); also the list the symbol-export test checks against the header
H = C.c_void_p
SIGNATURES = {
    "rg_last_error": (C.c_char_p, []),
    "rg_version": (C.c_char_p, []),
    "rg_device_count": (C.c_int, []),
    "rg_create": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(H)]),
    "rg_create_from_file": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(H)]),
    "rg_create_distributed": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(H)]),
    "rg_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "rg_destroy": (C.c_int, [H]),
    "rg_get_layout": (C.c_int, [H, C.POINTER(RgLayout)]),
    "rg_get_param": (C.c_int, [H, C.c_char_p, C.POINTER(C.c_double)]),
    "rg_init_simulation": (C.c_int, [H, C.c_char_p, C.POINTER(C.c_int)]),
    "rg_make_all_boundaries": (C.c_int, [H, C.c_int]),
    "rg_compute_dt": (C.c_int, [H, C.c_int, C.POINTER(C.c_double)]),
    "rg_godunov_unsplit": (C.c_int, [H, C.c_int, C.c_double]),
    "rg_one_step": (C.c_int, [H, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "rg_run": (C.c_int, [H]),
    "rg_output": (C.c_int, [H, C.c_int]),
    "rg_get_data_device": (C.c_int, [H, C.c_int, C.POINTER(C.c_void_p)]),
    "rg_copy_to_host": (C.c_int, [H, C.c_int, C.c_void_p, C.c_size_t]),
    "rg_copy_from_host": (C.c_int, [H, C.c_int, C.c_void_p, C.c_size_t]),
    "rg_synchronize": (C.c_int, [H]),
    "rg_steps_from_host": (C.c_int, [H, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "rg_alloc_pinned": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "rg_free_pinned": (C.c_int, [C.c_void_p]),
    "rg_steps_from_host_batch": (C.c_int, [H, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_size_t, C.POINTER(C.c_double)]),
    "rg_get_stats": (C.c_int, [H, C.POINTER(RgStats)]),
    "rg_reset_launch_count": (C.c_int, []),
    "rg_set_chunk_planes": (C.c_int, [H, C.c_int]),
    "rg_set_halo_overlap": (C.c_int, [H, C.c_int]),
    "rg_set_tuning": (C.c_int, [C.c_char_p, C.c_int]),
    "rg_history": (C.c_int, [H, C.c_int, C.POINTER(C.c_double)]),
    "rg_initial_condition_host": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(RgLayout)]),
    "rg_profile_begin": (C.c_int, [H]),
    "rg_profile_end": (C.c_int, [H, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_ulonglong)]),
    "rg_probe_riemann_mhd": (C.c_int, [H, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rg_probe_compute_emf": (C.c_int, [H, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rg_slab_extent": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
}

_lib = None


def load():
    """Loads the native library (building it is the job of __graft_entry__.build / ramsesgpu_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "ramsesgpu_b200: native library %s is missing -- run `python -m ramsesgpu_b200.build`. "
                "There is no Python or CPU fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class RgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ramsesgpu_b200 error %d: %s" % (code, msg))
        self.code = code


def check(rc):
    if rc != 0:
        raise RgError(rc, load().rg_last_error().decode(errors="replace"))
