"""Probe: how far does the bench2 driver loop get (dt growth) with the reference's solver settings?"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pfhubbenchmarks_b200 import api as df
from pfhubbenchmarks_b200.drivers import common
name = sys.argv[1] if len(sys.argv) > 1 else "bench2"
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
L = {"bench2": 200.0, "bench3": 960.0, "bench6": 100.0}[name]; N = {"bench2": 100, "bench3": 350, "bench6": 100}[name]
mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(L, L), N, N, "crossed")
dt = df.Constant(1e-2)
problem = df.NonlinearVariationalProblem(name, mesh, dt)
solver = df.NonlinearVariationalSolver(problem)
nl = solver.parameters["snes_solver"]; nl["absolute_tolerance"] = 1e-6; nl["maximum_iterations"] = 10
nl["krylov_solver"]["maximum_iterations"] = 5000 if name == "bench6" else 1000
ic = {"bench2": df.InitialConditionsBench2(0.5, 0.05, 0.1, 1.5), "bench3": df.InitialConditionsBench3(-0.3, 8.0, 1.0, 1.0, -1.0),
      "bench6": df.InitialConditionsBench6(0.5, 0.04)}[name]
t0 = time.time()
rows, spent = common.time_loop(problem, solver, ic, 1e-2, 1e-4, 1e9, "/tmp/%s_probe.csv" % name, "time,a,b", max_steps=nsteps, quiet=True)
rows = np.array(rows)
print(name, "steps", len(rows), "t_end %.3f" % rows[-1, 0], "max dt %.3f" % np.diff(np.concatenate([[0], rows[:, 0]])).max(), "wall %.1f s" % spent)
print("times:", np.round(rows[::max(1, len(rows) // 12), 0], 3))
print("energy first/last:", rows[0, 1], rows[-1, 1], "col2 drift", abs(rows[-1, 2] / rows[0, 2] - 1))
