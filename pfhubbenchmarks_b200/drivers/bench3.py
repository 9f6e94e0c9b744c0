"""PFHub benchmark 3, dendritic growth (U, phi), isotropic as coded -- the model of dolfin/bench3.py.

    python -m pfhubbenchmarks_b200.drivers.bench3 [--backend b200|fenics] [--end-time 100.0]
"""
import os

from .. import dolfin as df
from ..pfbase import (InitialConditionsBench3, allen_cahn_RHS_IBP, allen_cahn_weak_form, cahn_hilliard_weak_form,  # noqa: F401
                      diffusion_weak_form, dx, grad, h, inner, poisson_weak_form)
from .common import mixed_p1, parse_args, run_fenics, snes_setup, time_loop


def main():
    args = parse_args("bench3", 100.0)
    if args.backend == "fenics":
        return run_fenics(args, "bench3.py")
    df.parameters["form_compiler"]["quadrature_degree"] = 3
    Lx = Ly = 960.0
    mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(Lx, Ly), 350, 350, "crossed")            # bench3.py:21-23
    dt = df.Constant(1e-1)
    W0, tau0, D, Delta = (df.Constant(v) for v in (1.0, 1.0, 10.0, -0.3))                          # bench3.py:32-38
    W, w, w0, w_, dw = mixed_p1(mesh, 2)
    w_ic = InitialConditionsBench3(float(Delta), 8.0, 1.0, 1.0, -1.0, degree=2)                    # bench3.py:53-60
    U, phi = df.split(w)
    lam = D * tau0 / (0.6267 * W0 ** 2)
    f_chem = -phi ** 2 / 2 + phi ** 4 / 4 + lam * U * phi * (1 - 2 * phi ** 2 / 3 + phi ** 4 / 5)
    a = 1.0                                                # the anisotropy is switched off in the reference (bench3.py:82-83)
    Wa, tau = W0 * a, tau0 * a ** 2
    dfdp = (phi - lam * U * (1 - phi ** 2)) * (1 - phi ** 2)
    zero = df.Constant(0.0)
    # dU/dt = D lapl U + 1/2 dphi/dt, the latter substituted by the Allen-Cahn right-hand side (bench3.py:94-100)
    F_U = diffusion_weak_form(U, w_[0], w0[0], dt, df.Constant(1.0), D, df.Constant((0.0, 0.0)), zero)
    F_U = F_U - 0.5 * allen_cahn_RHS_IBP(phi, w_[0], 1 / tau, Wa ** 2, -dfdp, zero) * dx
    F_phi = allen_cahn_weak_form(phi, w_[1], w0[1], dt, 1 / tau, Wa ** 2, -dfdp, zero)
    F = F_U + F_phi
    problem = df.NonlinearVariationalProblem(F, w, [], df.derivative(F, w, dw))
    solver = df.NonlinearVariationalSolver(problem)
    snes_setup(solver, args.atol, "basic", 1000)

    def total_free_energy():                               # bench3.py:162-167
        return df.assemble((f_chem + 0.5 * Wa ** 2 * inner(grad(phi), grad(phi))) * dx)

    def solid_fraction():                                  # bench3.py:169-170
        return df.assemble(0.5 * (phi + 1.0) * dx) / (Lx * Ly)

    # dt0 = 1e-2, dt_min = 1e-4 (bench3.py time-integration block)
    return time_loop(problem, solver, w_ic, 1e-2, 1e-4, args.end_time, os.path.join(args.outdir, "bench3_out.csv"),
                     "time,total_free_energy,solid_fraction", args.max_steps, args.quiet, abort_on_negative_energy=True,
                     save_dir=os.path.join(args.outdir, "bench3") if args.save_solution else None, save_fields=(0, 1),
                     functionals=(total_free_energy, solid_fraction))


if __name__ == "__main__":
    main()
