"""ramsesgpu_b200 -- B200-native (sm_100a) per-timestep Godunov update path behind the operator
surface of pkestene/ramsesGPU.  See DESIGN.md / INTEGRATION.md."""
from .io import ini_override, l2_relative, read_vti, read_xsm  # noqa: F401


def __getattr__(name):
    # the compute API needs the native library; keep `import ramsesgpu_b200` itself light
    if name in ("HydroRunBase", "MHDRunBase", "HydroRunGodunov", "MHDRunGodunov", "reset_launch_count", "slab_extent", "set_tuning", "PinnedArray", "initial_condition_host"):
        from . import runs
        return getattr(runs, name)
    raise AttributeError(name)
