"""One process, b13d 256^3: a few steps with the fp64 and with the fp32 quadrature block of the Krylov J.v; prints the
Newton/Krylov counts, the jac_apply family time and the relative difference of the solutions."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from pfhubbenchmarks_b200 import engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
res = {}
for mode in (0, 1):
    ctx = engine.Context("b13d", (n - 1,) * 3, (2.0 * (n - 1),) * 3)
    ctx.set_option("jac_quadrature_fp32", mode)
    w = ctx.initial_condition()
    its = [0, 0]
    for k, dt in enumerate((0.1, 0.2, 0.4, 0.8, 1.6)):
        if k == 2:
            ctx.set_option("profiling", 1)
            ctx.timers(reset=True)
        w0 = w.clone()
        ctx.dt = dt
        rep = ctx.newton_solve(w, w0)
        assert rep["converged"], rep
        its[0] += rep["niters"]; its[1] += rep["krylov_its"]
    t = ctx.timers()
    res[mode] = (w.clone(), its, t)
    print("fp32q" if mode else "fp64q", "newton/krylov", its, {k: (round(v["ms"], 2), v["launches"]) for k, v in t.items() if k in ("jac_apply", "cheb_sweep", "multidot")}, flush=True)
    ctx.close()
d = (res[0][0] - res[1][0])[: (n ** 3)]
print("rel diff of c after 5 steps: %.2e" % float(d.norm() / res[0][0][: n ** 3].norm()))
