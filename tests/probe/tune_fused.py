"""Tuning aid: time of one preconditioner application (512^3) with per-sweep kernels vs the fused two-sweep kernels."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pfhubbenchmarks_b200 import engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 511
ctx = engine.Context("b13d", (n, n, n), (2.0 * n,) * 3, krylov_basis=2)
U = ctx.initial_condition()
X = torch.randn(2 * ctx.n_owned, dtype=torch.float64, device="cuda")
ctx.dt = 1.6
ctx.jac_setup(U)
ctx.precond_setup(pc_variant=1, pc_sweeps_mass=3, pc_sweeps_shifted=5)
z = ctx.new_vector()
ref = None
for fused in (0, 1, 0, 1):
    ctx.set_option("use_fused", fused)
    ctx.precond_apply(X, out=z)
    ctx.set_option("profiling", 1); ctx.timers(reset=True)
    for _ in range(3):
        ctx.precond_apply(X, out=z)
    t = ctx.timers(reset=True); ctx.set_option("profiling", 0)
    zz = z.cpu().numpy()
    if ref is None: ref = zz
    tot = sum(v["ms"] for v in t.values()) / 3
    cs, cf = t["cheb_sweep"], t["cheb_first"]
    print("fused %d: precond apply %.2f ms | cheb kernels %.2f ms (%d launches, %.3f ms each) cheb_first %.2f ms | maxdiff %.1e" % (
        fused, tot, cs["ms"] / 3, cs["launches"] // 3, cs["ms"] / max(cs["launches"], 1), cf["ms"] / 3, float(np.abs(zz - ref).max() / np.abs(ref).max())), flush=True)
