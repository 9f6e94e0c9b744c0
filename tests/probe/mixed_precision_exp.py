"""Mixed-precision experiments on the CPU solver leg (tests/probe/README.md).  Build:
  g++ -O2 -fopenmp -std=c++17 -shared -fPIC -Wl,-Bsymbolic -I/usr/local/cuda/include -Itests/emul -o /tmp/libexp.so \
      tests/probe/emul_solver_mixed_precision_exp.cpp pfhubbenchmarks_b200/csrc/pfh_tables.cpp
Run:  [PFH_EXP_F32=1] [PFH_EXP_F32ARITH=1] [PFH_EXP_JAC_F32=1] python tests/probe/mixed_precision_exp.py
"""
import sys, os, ctypes, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import emul
from helpers import discretization
emul._SOLVER_LIB = "/tmp/libexp.so"; emul.build_solver = lambda user_header=None: "/tmp/libexp.so"
name, mk, pid, ncell, length, params = "b13d", 3, 1, (14, 12, 10), (28.0, 24.0, 20.0), [0.3, 0.7, 5.0, 2.0, 5.0]
D = discretization(name, ncell, length); u = D.prob.ic(D.mesh)
tot = 0
for dt in (0.1, 0.2, 0.4, 0.8, 1.6, 1.6):
    w, rep = emul.newton_step(mk, pid, ncell, length, params, dt, u, u, basis_precision=2)
    print(dt, int(rep["niters"]), int(rep["krylov_its"]), int(rep["converged"]), "%.2e" % rep["fnorm"])
    tot += rep["krylov_its"]; u = w
print("total krylov", tot, "f32" if os.environ.get("PFH_EXP_F32") else "f64")
