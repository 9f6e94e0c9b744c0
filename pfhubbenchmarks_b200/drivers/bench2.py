"""PFHub benchmark 2, Ostwald ripening (c, mu, eta1..4) -- the model of dolfin/bench2.py.

    python -m pfhubbenchmarks_b200.drivers.bench2 [--backend b200|fenics] [--end-time 100.0]
"""
import itertools
import os

from .. import dolfin as df
from ..pfbase import (InitialConditionsBench2, allen_cahn_RHS_IBP, allen_cahn_weak_form, cahn_hilliard_weak_form,  # noqa: F401
                      diffusion_weak_form, dx, grad, h, inner, poisson_weak_form)
from .common import mixed_p1, parse_args, run_fenics, snes_setup, time_loop

NUM_ETA = 4


def main():
    args = parse_args("bench2", 100.0)
    if args.backend == "fenics":
        return run_fenics(args, "bench2.py")
    df.parameters["form_compiler"]["quadrature_degree"] = 3
    mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(200.0, 200.0), 100, 100, "crossed")      # bench2.py:21-23
    dt = df.Constant(1e-1)
    # bench2.py:34-42 (rho = sqrt 2 enters as rho^2)
    c_alpha, c_beta, rho, kappa_c, kappa_eta, M, ww, alpha, L = (
        df.Constant(v) for v in (0.3, 0.7, 2.0 ** 0.5, 3.0, 3.0, 5.0, 1.0, 5.0, 5.0))
    W, w, w0, w_, dw = mixed_p1(mesh, 2 + NUM_ETA)
    w_ic = InitialConditionsBench2(0.5, 0.05, 0.1, 1.5, degree=2)                                  # bench2.py:61-67
    c, mu, *eta = df.split(w)
    interp = sum(h(e) for e in eta)
    wells = sum(e ** 2 * (1 - e) ** 2 for e in eta) + alpha * sum(a ** 2 * b ** 2 for a, b in itertools.combinations(eta, 2))
    f_chem = rho ** 2 * (c - c_alpha) ** 2 * (1 - interp) + rho ** 2 * (c - c_beta) ** 2 * interp + ww * wells
    F = cahn_hilliard_weak_form(c, mu, w_[0], w_[1], w0[0], dt, M, kappa_c, df.diff(f_chem, c))
    for i, e in enumerate(eta):
        F = F + allen_cahn_weak_form(e, w_[2 + i], w0[2 + i], dt, L, kappa_eta, df.diff(f_chem, e), df.Constant(0))
    problem = df.NonlinearVariationalProblem(F, w, [], df.derivative(F, w, dw))
    solver = df.NonlinearVariationalSolver(problem)
    snes_setup(solver, args.atol, "cp", 1000)         # 'cp' as the reference (bench2.py:140)

    def total_free_energy():                           # bench2.py:187-194: one assemble per gradient term
        E = df.assemble(f_chem * dx + kappa_c / 2.0 * inner(grad(c), grad(c)) * dx)
        return E + sum(df.assemble(kappa_eta / 2.0 * inner(grad(e), grad(e)) * dx) for e in eta)

    def total_solute():
        return df.assemble(c * dx)

    # dt0 = 1e-2, dt_min = 1e-4 (bench2.py time-integration block)
    return time_loop(problem, solver, w_ic, 1e-2, 1e-4, args.end_time, os.path.join(args.outdir, "bench2_out.csv"),
                     "time,total_free_energy,total_solute", args.max_steps, args.quiet,
                     save_dir=os.path.join(args.outdir, "bench2") if args.save_solution else None,
                     functionals=(total_free_energy, total_solute))


if __name__ == "__main__":
    main()
