"""Profiling aid (not a test): launches every kernel family of the b13d step once or twice on a 512^3 grid, for
`ncu --set full` (one GPU).  usage: ncu_kernels.py [3d|2d] [grid]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pfhubbenchmarks_b200 import _lib, engine
which = sys.argv[1] if len(sys.argv) > 1 else "3d"
if which == "3d":
    nv = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    n = nv - 1
    ctx = engine.Context("b13d", (n, n, n), (2.0 * n, 2.0 * n, 2.0 * n), krylov_basis=30)
else:
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    ctx = engine.Context("bench1", (n, n), (2.0 * n, 2.0 * n), krylov_basis=30)
w = ctx.initial_condition()
w0 = w.clone()
ctx.dt = 1.6
# one Newton iteration with three Krylov iterations: residual, J.v, the whole preconditioner, multi-dot / multi-axpy of
# the fp32-stored basis, the vector kernels; then the energy reduction
opts = _lib.default_opts(maximum_iterations=1, krylov_maximum_iterations=3, pc_sweeps_mass=3, pc_sweeps_shifted=4)
rep = ctx.newton_solve(w, w0, opts=opts)
print(rep)
print(ctx.energy(w))
ctx.close()
