"""A/B timing of tuning knobs on the bench workload (GPU box):
   python tools/ab.py [n=256] knob=v1,v2,... [knob2=...]   -> per-phase device times for every combination"""
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ramsesgpu_b200 import MHDRunGodunov, set_tuning  # noqa: E402

args = [a for a in sys.argv[1:] if "=" not in a]
n = int(args[0]) if args else 256
knobs = [(a.split("=")[0], [int(v) for v in a.split("=")[1].split(",")]) for a in sys.argv[1:] if "=" in a]
run = MHDRunGodunov(bench.workload_ini(n, n))
run.init_simulation()
run.make_all_boundaries(0)
s = (0, 0.0, 0.0)
for _ in range(40):
    s = run.oneStepIntegration(*s)
for rep in range(2):
    for combo in itertools.product(*[v for _, v in knobs]):
        for (k, _), v in zip(knobs, combo):
            set_tuning(k, v)
        for _ in range(2):
            s = run.oneStepIntegration(*s)
        run.profile_begin()
        for _ in range(5):
            s = run.oneStepIntegration(*s)
        tot, ph = run.profile_end()
        print(" ".join("%s=%d" % (k, v) for (k, _), v in zip(knobs, combo)), "total %.3f |" % (tot / 5),
              " ".join("%s %.3f" % (k, v[0] / 5) for k, v in ph.items() if v[0] > 0), flush=True)
