"""Debug helper: exercise one TMA kernel family in isolation. usage: dbg_tma.py scalar|jac [nx ny nz]"""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pfhubbenchmarks_b200 import engine

which = sys.argv[1]
ncell = tuple(int(a) for a in sys.argv[2:5]) if len(sys.argv) >= 5 else (15, 16, 17)
length = tuple(2.0 * n for n in ncell)
rng = np.random.default_rng(0)
ctx = engine.Context("b13d", ncell, length)
n = ctx.n_owned
U = ctx.initial_condition()
X = torch.from_numpy(rng.standard_normal(2 * n)).cuda()
ctx.dt = 0.4
res = {}
for tma in (0, 1):
    ctx.set_option("use_tma", tma)
    ctx.jac_setup(U)
    if which == "scalar":
        ctx.precond_setup(pc_variant=1, pc_sweeps_mass=3, pc_sweeps_shifted=5)
        res[tma] = ctx.precond_apply(X).cpu().numpy()
    else:
        res[tma] = ctx.jac_apply(X).cpu().numpy()
    print(which, "tma", tma, "ok", float(np.abs(res[tma]).max()), flush=True)
print(which, "max rel diff", float(np.abs(res[1] - res[0]).max() / np.abs(res[0]).max()))
