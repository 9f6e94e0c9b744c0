"""PFHub benchmark 6, Cahn-Hilliard with electrostatics (c, mu, phi) -- the model of dolfin/bench6.py.

    python -m pfhubbenchmarks_b200.drivers.bench6 [--backend b200|fenics] [--end-time 3.0]
"""
import os

from .. import dolfin as df
from ..pfbase import (InitialConditionsBench6, allen_cahn_RHS_IBP, allen_cahn_weak_form, cahn_hilliard_weak_form,  # noqa: F401
                      diffusion_weak_form, dx, grad, h, inner, poisson_weak_form)
from .common import mixed_p1, parse_args, run_fenics, snes_setup, time_loop


def main():
    args = parse_args("bench6", 3.0)
    if args.backend == "fenics":
        return run_fenics(args, "bench6.py")
    df.parameters["form_compiler"]["quadrature_degree"] = 3
    Lx = Ly = 100.0
    mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(Lx, Ly), 100, 100, "crossed")            # bench6.py:21-23
    dt = df.Constant(1e-1)
    c_alpha, c_beta, kappa, rho, M, k, epsilon = (df.Constant(v) for v in (0.3, 0.7, 2.0, 5.0, 5.0, 0.09, 90.0))   # bench6.py:33-39
    W, w, w0, w_, dw = mixed_p1(mesh, 3)
    w_ic = InitialConditionsBench6(0.5, 0.04, degree=2)                                            # bench6.py:52-57
    c, mu, phi = df.split(w)
    f_chem = rho * (c - c_alpha) ** 2 * (c_beta - c) ** 2
    f_elec = k * c * phi / 2.0
    F = cahn_hilliard_weak_form(c, mu, w_[0], w_[1], w0[0], dt, M, kappa, df.diff(f_chem, c) + k * phi)
    F = F + poisson_weak_form(phi, w_[2], -k * c / epsilon, df.Constant(1.0))
    # phi = 0 on x = 0, phi = sin(y/7) on x = Lx; no-flux top and bottom (bench6.py:77-90)
    W_phi = W.split()[2]
    bcs = [df.DirichletBC(W_phi, df.Constant(0.0), lambda x, on_boundary: on_boundary and df.near(x[0], 0.0, 1e-12)),
           df.DirichletBC(W_phi, df.Expression("sin(x[1]/7)", degree=2), lambda x, on_boundary: on_boundary and df.near(x[0], Lx, 1e-12))]
    problem = df.NonlinearVariationalProblem(F, w, bcs, df.derivative(F, w, dw))
    solver = df.NonlinearVariationalSolver(problem)
    snes_setup(solver, args.atol, "cp", 5000)          # 'cp' as the reference (bench6.py:117)

    def total_free_energy():                           # bench6.py:158-165
        return df.assemble((f_chem + f_elec + kappa / 2.0 * inner(grad(c), grad(c))) * dx)

    def total_solute():
        return df.assemble(c * dx)

    # dt0 = 1e-2, dt_min = 1e-4 (bench6.py time-integration block)
    return time_loop(problem, solver, w_ic, 1e-2, 1e-4, args.end_time, os.path.join(args.outdir, "bench6_out.csv"),
                     "time,total_free_energy,total_solute", args.max_steps, args.quiet,
                     save_dir=os.path.join(args.outdir, "bench6") if args.save_solution else None,
                     functionals=(total_free_energy, total_solute))


if __name__ == "__main__":
    main()
