"""Tuning aid: Chebyshev-sweep bandwidth of the TMA scalar kernel variants at 512^3 (not a test)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pfhubbenchmarks_b200 import engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 511
nz = int(sys.argv[2]) if len(sys.argv) > 2 else n
ctx = engine.Context("b13d", (n, n, nz), (2.0 * n, 2.0 * n, 2.0 * nz), krylov_basis=2)
U = ctx.initial_condition()
X = torch.randn(2 * ctx.n_owned, dtype=torch.float64, device="cuda")
ctx.dt = 1.6
ctx.jac_setup(U)
ctx.precond_setup(pc_variant=1, pc_sweeps_mass=3, pc_sweeps_shifted=5)
z = ctx.new_vector()
ref = None
for v in (0, 1, 2, 0):
    ctx.set_option("scalar_variant", v)
    ctx.precond_apply(X, out=z)          # warm-up / attribute setup
    ctx.set_option("profiling", 1); ctx.timers(reset=True)
    for _ in range(3):
        ctx.precond_apply(X, out=z)
    t = ctx.timers(reset=True); ctx.set_option("profiling", 0)
    zz = z[:1000].cpu().numpy()
    if ref is None: ref = zz
    cs, la = t["cheb_sweep"], t["lin_apply"]
    print("variant %d: cheb_sweep %.1f GB/s (%.3f ms/launch)  lin_apply %.1f GB/s  maxdiff %.1e" % (
        v, cs["bytes"] / cs["ms"] / 1e6, cs["ms"] / cs["launches"], la["bytes"] / la["ms"] / 1e6, float(np.abs(zz - ref).max())), flush=True)
