"""Per-family device times of the nl.py kernels (csrc/pfh_nl.cu) on a synthetic grid, tiled vs gather mapping:
    python tests/probe/nl_kernels.py [n=4096]
Prints one JSON line per mapping: ms per launch and algorithmic GB/s (residual 16, J.v 24, sweep 40 B/vertex)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from pfhubbenchmarks_b200 import engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for tiled in (1, 0):
    ctx = engine.Context("nl", (n, n), (1.0, 1.0))
    ctx.set_option("nl_tiled", tiled)
    N = ctx.n_owned
    g = torch.Generator(device="cuda").manual_seed(1)
    u = 0.9 + 0.3 * torch.rand(N, dtype=torch.float64, device="cuda", generator=g)
    x = torch.randn(N, dtype=torch.float64, device="cuda", generator=g)
    out = ctx.new_vector()
    ctx.jac_setup(u)
    ctx.precond_setup(pc_sweeps_shifted=12)
    for rep in range(2):                      # first pass = warm-up
        ctx.set_option("profiling", 1)
        ctx.timers(reset=True)
        for _ in range(10):
            ctx.residual(u, u, out=out)
            ctx.jac_apply(x, out=out)
            ctx.precond_apply(x, out=out)
        t = ctx.timers(reset=True)
    ctx.close()
    ns = min(n, 1024)                           # the solve: Krylov counts grow like 1/h, 4096^2 does not finish in 5000
    ctx = engine.Context("nl", (ns, ns), (1.0, 1.0))
    ctx.set_option("nl_tiled", tiled)
    w = torch.full((ctx.n_owned,), 0.9, dtype=torch.float64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    rep = ctx.newton_solve(w, w, absolute_tolerance=1e-6, maximum_iterations=10, krylov_maximum_iterations=5000, pc_sweeps_shifted=32)
    e1.record()
    torch.cuda.synchronize()
    line = {"grid": n, "tiled": tiled, "vertices": N}
    for k in ("residual", "jac_apply", "cheb_sweep"):
        ms = t[k]["ms"] / max(1, t[k]["launches"])
        line[k] = {"ms": round(ms, 4), "GB/s": round(t[k]["bytes"] / t[k]["launches"] / ms / 1e6, 1), "launches": t[k]["launches"]}
    line["solve"] = {"grid": ns, "ms": round(e0.elapsed_time(e1), 1), "newton": rep["niters"], "krylov": rep["krylov_its"], "converged": rep["converged"]}
    print(json.dumps(line), flush=True)
    ctx.close()
