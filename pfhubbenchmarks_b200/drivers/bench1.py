"""PFHub benchmark 1, spinodal decomposition on a 2-D square -- the model of dolfin/bench1.py.

    python -m pfhubbenchmarks_b200.drivers.bench1 [--backend b200|fenics] [--end-time 1e3]

The free energy and the weak form are written as in the reference (UFL-style expressions, pfbase builders); the
problem object recognises the form and runs it on the compiled Cahn-Hilliard operators.
"""
import os

from .. import dolfin as df
from ..pfbase import (InitialConditionsBench1, allen_cahn_RHS_IBP, allen_cahn_weak_form, cahn_hilliard_weak_form,  # noqa: F401
                      diffusion_weak_form, dx, grad, h, inner, poisson_weak_form)
from .common import mixed_p1, parse_args, run_fenics, snes_setup, time_loop


def main():
    args = parse_args("bench1", 1e3)
    if args.backend == "fenics":
        return run_fenics(args, "bench1.py")
    df.parameters["form_compiler"]["quadrature_degree"] = 3
    mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(200.0, 200.0), 100, 100, "crossed")      # bench1.py:21-23
    dt = df.Constant(1e-1)
    c_alpha, c_beta, rho_s, kappa, M = (df.Constant(v) for v in (0.3, 0.7, 5.0, 2.0, 5.0))        # bench1.py:29-36
    W, w, w0, w_, dw = mixed_p1(mesh, 2)
    w_ic = InitialConditionsBench1(0.5, 0.05, degree=2)                                            # bench1.py:48-53
    c, mu = df.split(w)
    f_chem = rho_s * (c - c_alpha) ** 2 * (c_beta - c) ** 2
    F = cahn_hilliard_weak_form(c, mu, w_[0], w_[1], w0[0], dt, M, kappa, df.diff(f_chem, c))
    problem = df.NonlinearVariationalProblem(F, w, [], df.derivative(F, w, dw))
    solver = df.NonlinearVariationalSolver(problem)
    snes_setup(solver, args.atol, "basic", 1000)

    def total_free_energy():
        return df.assemble(f_chem * dx + kappa / 2.0 * inner(grad(c), grad(c)) * dx)

    def total_solute():
        return df.assemble(c * dx)

    # time integration + CSV (bench1.py:128-217): dt0 = 1e-1, dt_min = 1e-2
    return time_loop(problem, solver, w_ic, 1e-1, 1e-2, args.end_time, os.path.join(args.outdir, "bench1", "stats.csv"),
                     "time,total_free_energy,total_solute", args.max_steps, args.quiet,
                     save_dir=os.path.join(args.outdir, "bench1") if args.save_solution else None,
                     functionals=(total_free_energy, total_solute))


if __name__ == "__main__":
    main()
