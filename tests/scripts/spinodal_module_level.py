# A module-level driver in the manner of the reference's scripts: `import dolfin as df`, `from pfbase import *`, everything
# at global scope, an inline adaptive time loop, an HDF5 series plus a CSV at the end.  Written for the tests (small mesh,
# three accepted steps).  Run:  python -m pfhubbenchmarks_b200.run_script tests/scripts/spinodal_module_level.py [outdir]
import os
import sys

import dolfin as df
import numpy as np
import matplotlib.pyplot as plt  # noqa: F401  -- the reference drivers import it; nothing is plotted while stepping

from pfbase import *  # noqa: F401,F403

target = sys.argv[1] if len(sys.argv) > 1 else "results/spinodal_small"
os.makedirs(target, exist_ok=True)
df.parameters["form_compiler"]["quadrature_degree"] = 3
df.parameters["form_compiler"]["cpp_optimize"] = True

side, cells = 48.0, 24
domain = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(side, side), cells, cells, 'crossed')
step = df.Constant(1e-1)
well_lo, well_hi, barrier, grad_coef, mobility = [df.Constant(v) for v in (0.3, 0.7, 5.0, 2.0, 5.0)]

scalar = df.FunctionSpace(domain, 'P', 1).ufl_element()
space = df.FunctionSpace(domain, df.MixedElement([scalar, scalar]))
state, previous = df.Function(space), df.Function(space)
trial, test = df.TrialFunction(space), df.TestFunction(space)
start = InitialConditionsBench1(0.5, 0.05, degree=2)
previous.interpolate(start)                     # before the problem exists, as the reference does

conc, pot = df.split(state)
conc = variable(conc)
bulk = barrier * (conc - well_lo)**2 * (well_hi - conc)**2
form = cahn_hilliard_weak_form(state[0], state[1], test[0], test[1], previous[0], step, mobility, grad_coef, df.diff(bulk, conc))
jac = df.derivative(form, state, trial)

df.set_log_level(df.LogLevel.WARNING)
pde = df.NonlinearVariationalProblem(form, state, [], jac)
newton = df.NonlinearVariationalSolver(pde)
newton.parameters['nonlinear_solver'] = 'snes'
for key, val in dict(report=True, error_on_nonconvergence=False, absolute_tolerance=1e-6, maximum_iterations=10,
                     line_search='basic', linear_solver='gmres', preconditioner='sor').items():
    newton.parameters['snes_solver'][key] = val
newton.parameters['snes_solver']['krylov_solver']['maximum_iterations'] = 1000

series = HDF5File(MPI.comm_world, os.path.join(target, "conc.h5"), "w")
series.write(domain, "mesh")


def solute(field):
    return df.assemble(field * dx)


def free_energy(density, coef, field):
    return df.assemble(density*dx + coef/2.0*inner(grad(field), grad(field))*dx)


clock = df.Constant(0.0)
state.interpolate(start)
previous.interpolate(start)
rows, smallest = [], 1e-2
step.assign(1e-1)
while len(rows) < 3:
    before = float(clock)
    previous.assign(state)
    clock.assign(before + float(step))
    its, ok = newton.solve()
    while not ok:                          # halve the step and retry from the previous state
        step.assign(max(0.5*float(step), smallest))
        clock.assign(before + float(step))
        state.assign(previous)
        its, ok = newton.solve()
    step.assign(2*float(step) if its < 5 else max(0.5*float(step), smallest))
    conc, _ = state.split()               # from here on `conc` is the component of the solution, as in the reference's analysis block
    series.write(conc, "c", float(clock))
    rows.append([float(clock), free_energy(bulk, grad_coef, conc), solute(conc)])

if df.MPI.rank(domain.mpi_comm()) == 0:
    np.savetxt(os.path.join(target, 'stats.csv'), np.array(rows), fmt='%1.10f', header="time,total_free_energy,total_solute",
               delimiter=',', comments='')
