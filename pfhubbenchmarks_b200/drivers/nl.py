"""The reference's nonlinear-solver script dolfin/nl.py on the B200 backend.

    python -m pfhubbenchmarks_b200.drivers.nl [--nx 60 --ny 75] [--atol 1e-6] [--backend b200|fenics]

Steady quasi-linear diffusion  -div(|u| grad u) + u = 5  on the unit square, P1 on the right-diagonal UnitSquareMesh
(nl.py:17-20), u = 0.9 on x = 0 and x = 1 (nl.py:31-35), solved from u = 0.9 twice with identical settings: once with the
'snes' tree, once with the 'newton' tree (nl.py:44-96), printing `(converged, niters)` of each as the reference does.
"""
import argparse
import os

import numpy as np

from .. import dolfin as df
from ..dolfin import inner, grad, dx, sqrt
from .common import run_fenics


def parse_args(argv=None):
    ap = argparse.ArgumentParser(description="nl.py of vpuri3/PFHubBenchmarks on the B200 backend")
    ap.add_argument("--backend", default="b200", choices=["b200", "fenics"])
    ap.add_argument("--nx", type=int, default=60)      # nl.py:17
    ap.add_argument("--ny", type=int, default=75)      # nl.py:18
    ap.add_argument("--atol", type=float, default=1e-6)
    ap.add_argument("--reference-dir", default=os.environ.get("PFHUB_REFERENCE", "/root/reference"))
    ap.add_argument("--quiet", action="store_true")
    return ap.parse_args(argv)


def configure(solver, tree, atol):
    """The block nl.py repeats for both solvers (nl.py:52-64, 79-91)."""
    solver.parameters['nonlinear_solver'] = tree
    nlparams = solver.parameters[tree + '_solver']
    if tree == 'snes':
        nlparams['line_search'] = 'basic'
    nlparams['report'] = True
    nlparams['error_on_nonconvergence'] = False
    nlparams['absolute_tolerance'] = atol
    nlparams['maximum_iterations'] = 10
    nlparams['linear_solver'] = 'gmres'
    nlparams['preconditioner'] = 'sor'       # mapped to Jacobi-Chebyshev sweeps on the SPD part of J (csrc/pfh_solver.cpp)
    nlparams['krylov_solver']['maximum_iterations'] = 1000
    nlparams['krylov_solver']['error_on_nonconvergence'] = False
    return nlparams


def main(argv=None):
    args = parse_args(argv)
    if args.backend == "fenics":
        return run_fenics(args, "nl.py")
    df.parameters["form_compiler"]["quadrature_degree"] = 3                                  # nl.py:12
    mesh = df.UnitSquareMesh(args.nx, args.ny)                                                 # nl.py:20
    V = df.FunctionSpace(mesh, "Lagrange", 1)
    u, v, du = df.Function(V), df.TestFunction(V), df.TrialFunction(V)
    u_bc = df.Function(V)
    u_bc.interpolate(df.Constant(0.9))
    bc = df.DirichletBC(V, u_bc, lambda x: np.logical_or(np.isclose(x[0], 0.0), np.isclose(x[0], 1.0)))   # nl.py:34-35
    F = 5.0 * v * dx - sqrt(u * u) * inner(grad(u), grad(v)) * dx - inner(u, v) * dx                      # nl.py:37
    J = df.derivative(F, u, du)
    problem = df.NonlinearVariationalProblem(F, u, [bc], J)
    out = {}
    for tree, label in (("snes", "SNES"), ("newton", "Newton")):
        u.vector()[:] = 0.9                                                                    # nl.py:48, 75
        solver = df.NonlinearVariationalSolver(problem)
        configure(solver, tree, args.atol)
        niters, converged = solver.solve()
        out[tree] = (niters, converged, dict(solver.last_report))
        if df.MPI.rank(mesh.mpi_comm()) == 0 and not args.quiet:
            print(label, "converged =", converged, ", niters =", niters)
    out["u"] = u.to_numpy()[0]
    return out


if __name__ == "__main__":
    main()
