"""Diagnostic (not a test): bench2 (six fields) on growing crossed meshes with the drivers' settings; prints the solver report per step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pfhubbenchmarks_b200 import _lib, engine
for n in [int(a) for a in sys.argv[1:]] or [512, 1024, 2048]:
    ctx = engine.Context("bench2", (n, n), (2.0 * n, 2.0 * n))
    w = ctx.initial_condition()
    w0 = w.clone()
    opts = _lib.default_opts(line_search=1, krylov_maximum_iterations=1000)
    for k, dt in enumerate([0.01, 0.02, 0.04, 0.08, 0.08, 0.08]):
        w0.copy_(w)
        ctx.dt = dt
        rep = ctx.newton_solve(w, w0, opts=opts)
        print(n, k, dt, {key: rep[key] for key in ("niters", "converged", "reason", "krylov_its", "fnorm0", "fnorm", "last_krylov_relres")}, flush=True)
        if not rep["converged"]:
            break
    ctx.close()
