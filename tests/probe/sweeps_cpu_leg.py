import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import emul
from helpers import discretization
name, mk, pid, ncell, length, params = "b13d", 3, 1, (20, 18, 16), (40.0, 36.0, 32.0), [0.3, 0.7, 5.0, 2.0, 5.0]
D = discretization(name, ncell, length); u00 = D.prob.ic(D.mesh)
def run(nM, nA, variant=1):
    u = u00; tot = 0; nit = 0
    for dt in (0.1, 0.2, 0.4, 0.8, 1.6, 1.6, 1.6, 1.6):
        w, rep = emul.newton_step(mk, pid, ncell, length, params, dt, u, u, basis_precision=2, pc_sweeps_mass=nM, pc_sweeps_shifted=nA, pc_variant=variant)
        assert rep["converged"] == 1
        if dt >= 0.8: tot += rep["krylov_its"]; nit += rep["niters"]
        u = w
    return int(nit), int(tot)
for cfg in [(3,5),(2,4),(3,4),(2,5),(2,3),(3,6),(4,5),(4,6),(3,7),(5,5)]:
    nit, its = run(*cfg)
    nM, nA = cfg
    cost = 0.78*((nM-2)+2*(nA-2)) + 0.86*3 + 1.47 + 4.5 + 5.8
    print(cfg, "newton", nit, "krylov", its, "model ms/it %.1f  total %.0f" % (cost, cost*its))
