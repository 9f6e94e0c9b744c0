"""Tuning aid (not a test): per-launch times of the staged 3-D kernels at 512^3 for every tile / ring / chunk variant.
usage: tune_kernels.py [n_vertices_per_edge]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pfhubbenchmarks_b200 import engine
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = nv - 1
ctx = engine.Context("b13d", (n, n, n), (2.0 * n, 2.0 * n, 2.0 * n), krylov_basis=2)
U = ctx.initial_condition()
X = torch.randn(2 * ctx.n_owned, dtype=torch.float64, device="cuda")
ctx.dt = 1.6
ctx.jac_setup(U)
z = ctx.new_vector()
ref = None


def fam(t, k):
    v = t[k]
    return "%s %.3f ms %.0f GB/s" % (k, v["ms"] / max(1, v["launches"]), v["bytes"] / max(v["ms"], 1e-9) / 1e6)


for loader in (1,):
    ctx.set_option("use_tma_loader", loader)
    for v in (0, 3):
        for zc in (32,):
            ctx.set_option("scalar_variant", v)
            ctx.set_option("scalar_zchunk", zc)
            ctx.precond_setup(pc_variant=1, pc_sweeps_mass=3, pc_sweeps_shifted=4)
            ctx.precond_apply(X, out=z)
            ctx.set_option("profiling", 1); ctx.timers(reset=True)
            for _ in range(3):
                ctx.precond_apply(X, out=z)
            t = ctx.timers(reset=True); ctx.set_option("profiling", 0)
            zz = z[:4096].cpu().numpy()
            if ref is None:
                ref = zz
            print("loader %d scalar variant %d zchunk %d: %s | %s | %s | maxdiff %.1e" % (
                loader, v, zc, fam(t, "cheb_sweep"), fam(t, "cheb_first"), fam(t, "lin_apply"), float(np.abs(zz - ref).max())), flush=True)
            if loader == 0 and zc == 32:
                break
ctx.set_option("use_tma_loader", 1)
ctx.set_option("scalar_variant", 0); ctx.set_option("scalar_zchunk", 0)
yref = None
y = ctx.new_vector()
for loader in (1,):
    ctx.set_option("use_tma_loader", loader)
    for jv in (0, 1):
        for zc in (32, 64, 128):
            ctx.set_option("jac_variant", jv); ctx.set_option("jac_zchunk", zc)
            ctx.jac_apply(X, out=y)
            ctx.residual(U, X, out=z)
            ctx.set_option("profiling", 1); ctx.timers(reset=True)
            for _ in range(3):
                ctx.jac_apply(X, out=y)
                ctx.residual(U, X, out=z)
            t = ctx.timers(reset=True); ctx.set_option("profiling", 0)
            yy = y[:4096].cpu().numpy()
            if yref is None:
                yref = yy
            print("loader %d jac variant %d zchunk %d: %s | %s | maxdiff %.1e" % (loader, jv, zc, fam(t, "jac_apply"), fam(t, "residual"),
                                                                                 float(np.abs(yy - yref).max())), flush=True)
E = ctx.energy(U)
ctx.set_option("profiling", 1); ctx.timers(reset=True)
for _ in range(3):
    ctx.energy(U)
t = ctx.timers(reset=True)
print(fam(t, "energy"))
