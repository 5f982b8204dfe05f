"""Sweeps the occupancy knobs of the FP64 kernels on the bench workload and prints per-phase device
times (ms per step).  Usage (GPU box): python tools/tune.py [size] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ramsesgpu_b200 import MHDRunGodunov, set_tuning  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
run = MHDRunGodunov(bench.workload_ini(n, n))
run.init_simulation()
run.make_all_boundaries(0)
state = {"nstep": 0, "t": 0.0, "dt": 0.0}


def measure(label):
    s = state
    for _ in range(2):
        s["nstep"], s["t"], s["dt"] = run.oneStepIntegration(s["nstep"], s["t"], s["dt"])
    run.profile_begin()
    for _ in range(steps):
        s["nstep"], s["t"], s["dt"] = run.oneStepIntegration(s["nstep"], s["t"], s["dt"])
    tot, ph = run.profile_end()
    print("%-28s total %7.3f | " % (label, tot / steps) + " ".join("%s %.3f" % (k, v[0] / steps) for k, v in ph.items() if v[0] > 0), flush=True)
    return {k: v[0] / steps for k, v in ph.items()}


# bring the GPU to its steady-state clocks first (the first second after idle runs at boost clocks)
for _ in range(60):
    state["nstep"], state["t"], state["dt"] = run.oneStepIntegration(state["nstep"], state["t"], state["dt"])
measure("steady-state warm-up")
best = {}
res = {}
for v in (128, 64, 32, 128, 64, 32):
    set_tuning("tile_x", v)
    res[v] = min(res.get(v, 1e9), sum(measure("tile_x=%d" % v).values()))
best["tile_x"] = min(res, key=res.get)
set_tuning("tile_x", best["tile_x"])
for key, phase in (("emf_minb", "emf"), ("flux_minb", "flux"), ("trace_minb", "trace"), ("update_minb", "update")):
    res = {}
    for v in (3, 4, 5, 6, 6, 5, 4, 3):
        set_tuning(key, v)
        res[v] = min(res.get(v, 1e9), measure("%s=%d" % (key, v))[phase])
    b = min(res, key=res.get)
    best[key] = b
    set_tuning(key, b)
print("best:", best)
measure("best")
