"""Probe: wall time per step of bench1 at the reference size with and without per-launch event profiling."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pfhubbenchmarks_b200 import engine, _lib
ctx = engine.Context("bench1", (100, 100), (200.0, 200.0))
opts = _lib.default_opts()
for prof in (0, 1, 0):
    w = ctx.initial_condition(); w0 = torch.empty_like(w)
    ctx.set_option("profiling", prof); ctx.set_option("reset_launch_count", 0)
    torch.cuda.synchronize(); t0 = time.perf_counter(); its = 0; kr = 0
    for k in range(8):
        w0.copy_(w); ctx.dt = min(0.1 * 2 ** k, 1.6)
        rep = ctx.newton_solve(w, w0, opts=opts); its += rep["niters"]; kr += rep["krylov_its"]
        ctx.energy(w)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("profiling %d: %.2f ms/step, newton %d krylov %d, launches %d, %.2f us/launch" % (prof, dt / 8 * 1e3, its, kr, ctx.launch_count(), dt / ctx.launch_count() * 1e6))
