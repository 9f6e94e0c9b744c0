"""Probe: the complete bench1 run (end_time = 1e3) through the driver loop with the reference's solver settings."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pfhubbenchmarks_b200 import api as df
from pfhubbenchmarks_b200.drivers import common
mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(200.0, 200.0), 100, 100, "crossed")
dt = df.Constant(1e-1)
problem = df.NonlinearVariationalProblem("bench1", mesh, dt)
solver = df.NonlinearVariationalSolver(problem)
nl = solver.parameters["snes_solver"]; nl["absolute_tolerance"] = 1e-6; nl["maximum_iterations"] = 10
nl["krylov_solver"]["maximum_iterations"] = 1000
rows, spent = common.time_loop(problem, solver, df.InitialConditionsBench1(0.5, 0.05), 1e-1, 1e-2, 1e3, "/tmp/bench1_full.csv",
                               "time,total_free_energy,total_solute", quiet=True)
rows = np.array(rows)
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "golden", "csv.npz"))["bench1"]
print("steps %d (reference %d)  t_end %.1f (reference %.1f)  wall %.1f s" % (len(rows), len(g), rows[-1, 0], g[-1, 0], spent))
print("max dt %.1f (reference %.1f)" % (np.diff(np.concatenate([[0], rows[:, 0]])).max(), np.diff(np.concatenate([[0], g[:, 0]])).max()))
print("final energy %.4f (reference %.4f)  solute drift %.2e  energy monotone: %s" % (rows[-1, 1], g[-1, 1], abs(rows[-1, 2] / rows[0, 2] - 1), bool(np.all(np.diff(rows[:, 1]) < 0))))
# energy at the reference's times by interpolation in log-time
Ei = np.interp(np.log(g[:, 0]), np.log(rows[:, 0]), rows[:, 1])
print("max relative energy difference vs reference curve (interpolated in log t): %.3e" % np.max(np.abs(Ei / g[:, 1] - 1)))
prefix = 0
while prefix < min(len(rows), len(g)) and abs(rows[prefix, 0] - g[prefix, 0]) < 1e-9: prefix += 1
print("rows matching time for time:", prefix)
