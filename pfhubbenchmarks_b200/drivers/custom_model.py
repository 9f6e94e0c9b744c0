"""A model that is NOT one of the reference's four systems, written with the same vocabulary: Cahn-Hilliard with a
Flory-Huggins (logarithmic) free energy coupled to one Allen-Cahn order parameter.

    python -m pfhubbenchmarks_b200.drivers.custom_model [--end-time 5.0] [--max-steps N]

`df.NonlinearVariationalProblem` does not recognise the weak form, so it lowers it (codegen.py), emits the pointwise
functions f', f'' (sympy derivatives, common subexpressions eliminated) as a CUDA functor and builds a copy of
libpfh_b200.so for it with nvcc (about 15 s, cached under pfhubbenchmarks_b200/_jit/) -- the step FFC performs for every
form in the reference (bench1.py:76).  The initial condition is a user-written `UserExpression` (evaluated in Python at
the vertices, as the reference does, pfbase.py:177-193).
"""
import os

import numpy as np

from .. import dolfin as df
from ..pfbase import allen_cahn_weak_form, cahn_hilliard_weak_form, dx, grad, inner
from .common import mixed_p1, parse_args, snes_setup, time_loop

CHI, MOBILITY, KAPPA, L_ETA, KAPPA_ETA = 2.5, 2.0, 1.5, 3.0, 0.7


class InitialConditions(df.UserExpression):
    """c: the bench1 perturbation pattern around 0.5 (inside the logarithms' domain); eta: a smooth bump; mu = 0."""

    def eval(self, values, x):
        values[0] = 0.5 + 0.05 * (np.cos(0.105 * x[0]) * np.cos(0.11 * x[1]) + (np.cos(0.13 * x[0]) * np.cos(0.087 * x[1])) ** 2
                                  + np.cos(0.025 * x[0] - 0.15 * x[1]) * np.cos(0.07 * x[0] - 0.02 * x[1]))
        values[1] = 0.0
        values[2] = 0.5 + 0.4 * np.cos(0.09 * x[0]) * np.cos(0.07 * x[1])

    def value_shape(self):
        return (3,)


def free_energy(c, eta, ln=df.ln):
    return c * ln(c) + (1 - c) * ln(1 - c) + CHI * c * (1 - c) + 0.5 * eta ** 2 * (1 - eta) ** 2 + 0.2 * c * eta ** 2


def setup(mesh, dt0=0.05, compile_only=False, **placement):
    """(problem, solver, total_free_energy, total_solute) for the model on `mesh`; compile_only: just build the
    generated library (no GPU needed) and return its path."""
    dt = df.Constant(dt0)
    M, kappa, L, kappa_eta = (df.Constant(v) for v in (MOBILITY, KAPPA, L_ETA, KAPPA_ETA))
    W, w, w0, w_, dw = mixed_p1(mesh, 3)
    c, mu, eta = df.split(w)
    f_chem = free_energy(c, eta)
    F = cahn_hilliard_weak_form(c, mu, w_[0], w_[1], w0[0], dt, M, kappa, df.diff(f_chem, c)) \
        + allen_cahn_weak_form(eta, w_[2], w0[2], dt, L, kappa_eta, df.diff(f_chem, eta), df.Constant(0))
    E = f_chem * dx + kappa / 2.0 * inner(grad(c), grad(c)) * dx + kappa_eta / 2.0 * inner(grad(eta), grad(eta)) * dx
    S = c * dx
    if compile_only:
        return df.compile_only(F, w, [], df.derivative(F, w, dw), functionals=(E, S))
    problem = df.NonlinearVariationalProblem(F, w, [], df.derivative(F, w, dw), functionals=(E, S), **placement)
    solver = df.NonlinearVariationalSolver(problem)
    return problem, solver, (lambda: df.assemble(E)), (lambda: df.assemble(S))


def main():
    args = parse_args("custom_model", 5.0)
    mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(200.0, 200.0), 100, 100, "crossed")
    problem, solver, total_free_energy, total_solute = setup(mesh)
    snes_setup(solver, args.atol, "basic", 1000)
    return time_loop(problem, solver, InitialConditions(degree=2), 5e-2, 1e-4, args.end_time,
                     os.path.join(args.outdir, "custom_model_out.csv"), "time,total_free_energy,total_solute", args.max_steps,
                     args.quiet, functionals=(total_free_energy, total_solute))


if __name__ == "__main__":
    main()
