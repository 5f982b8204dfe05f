"""A few steps of the bench workload for profiling under ncu (GPU box).
   python tools/prof_step.py [n=256] [steps=3] [key=value tuning knobs ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ramsesgpu_b200 import MHDRunGodunov, set_tuning  # noqa: E402

args = [a for a in sys.argv[1:] if "=" not in a]
n = int(args[0]) if args else 256
steps = int(args[1]) if len(args) > 1 else 3
for a in sys.argv[1:]:
    if "=" in a:
        k, v = a.split("=")
        set_tuning(k, int(v))
run = MHDRunGodunov(bench.workload_ini(n, n))
run.init_simulation()
run.make_all_boundaries(0)
s = (0, 0.0, 0.0)
for _ in range(steps):
    s = run.oneStepIntegration(*s)
