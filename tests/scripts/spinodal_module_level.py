# A module-level driver in the style of the reference's scripts (import dolfin as df / from pfbase import *, globals,
# an inline adaptive time loop, HDF5 + CSV output); written for the tests, small mesh, three steps.
# Run: python -m pfhubbenchmarks_b200.run_script tests/scripts/spinodal_module_level.py
import os
import sys

import dolfin as df
import numpy as np
import matplotlib.pyplot as plt  # noqa: F401  (imported by the reference drivers, never used while stepping)

from pfbase import *  # noqa: F401,F403

outdir = sys.argv[1] if len(sys.argv) > 1 else "results/spinodal_small"
os.makedirs(outdir, exist_ok=True)
df.parameters["form_compiler"]["cpp_optimize"] = True
df.parameters["form_compiler"]["quadrature_degree"] = 3

Lx = Ly = 48.0
mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(Lx, Ly), 24, 24, 'crossed')
dt = df.Constant(1e-1)
c_alpha, c_beta, rho_s, kappa, M = df.Constant(0.3), df.Constant(0.7), df.Constant(5.0), df.Constant(2.0), df.Constant(5.0)

P = df.FunctionSpace(mesh, 'P', 1)
W = df.FunctionSpace(mesh, df.MixedElement([P.ufl_element(), P.ufl_element()]))
w, w0 = df.Function(W), df.Function(W)
dw, w_ = df.TrialFunction(W), df.TestFunction(W)
w_ic = InitialConditionsBench1(0.5, 0.05, degree=2)
w0.interpolate(w_ic)

c, mu = df.split(w)
c = variable(c)
f_chem = rho_s * (c - c_alpha)**2 * (c_beta - c)**2
dfdc = df.diff(f_chem, c)
F = cahn_hilliard_weak_form(w[0], w[1], w_[0], w_[1], w0[0], dt, M, kappa, dfdc)
J = df.derivative(F, w, dw)

df.set_log_level(df.LogLevel.WARNING)
problem = df.NonlinearVariationalProblem(F, w, [], J)
solver = df.NonlinearVariationalSolver(problem)
solver.parameters['nonlinear_solver'] = 'snes'
nlparams = solver.parameters['snes_solver']
nlparams['report'] = True
nlparams['error_on_nonconvergence'] = False
nlparams['absolute_tolerance'] = 1e-6
nlparams['maximum_iterations'] = 10
nlparams['line_search'] = 'basic'
nlparams['linear_solver'] = 'gmres'
nlparams['preconditioner'] = 'sor'
nlparams['krylov_solver']['maximum_iterations'] = 1000

outfile = HDF5File(MPI.comm_world, os.path.join(outdir, "conc.h5"), "w")
outfile.write(mesh, "mesh")


def total_solute(c):
    return df.assemble(c * dx)


def total_free_energy(f_chem, kappa, c):
    return df.assemble(f_chem*dx + kappa/2.0*inner(grad(c), grad(c))*dx)


t = df.Constant(0.0)
w.interpolate(w_ic)
w0.interpolate(w_ic)
benchmark_output = []
dt_min = 1e-2
dt.assign(1e-1)
for iteration_count in range(3):
    tprev = float(t)
    w0.assign(w)
    t.assign(tprev + float(dt))
    niters, converged = solver.solve()
    while not converged:
        dt.assign(max(0.5*float(dt), dt_min))
        t.assign(tprev + float(dt))
        w.assign(w0)
        niters, converged = solver.solve()
    if niters < 5:
        dt.assign(2*float(dt))
    else:
        dt.assign(max(0.5*float(dt), dt_min))
    c, _ = w.split()                      # from here on c is the component of w, as in the reference's analysis block
    outfile.write(c, "c", float(t))
    benchmark_output.append([float(t), total_free_energy(f_chem, kappa, c), total_solute(c)])

if df.MPI.rank(mesh.mpi_comm()) == 0:
    np.savetxt(os.path.join(outdir, 'stats.csv'), np.array(benchmark_output), fmt='%1.10f',
               header="time,total_free_energy,total_solute", delimiter=',', comments='')
