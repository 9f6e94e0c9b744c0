"""PFHub benchmark 1, spinodal decomposition on a 2-D square -- mirrors dolfin/bench1.py.

    python -m pfhubbenchmarks_b200.drivers.bench1 [--backend b200|fenics] [--end-time 1e3]
"""
import os

from .. import api as df
from .common import parse_args, run_fenics, time_loop


def main():
    args = parse_args("bench1", 1e3)
    if args.backend == "fenics":
        return run_fenics(args, "bench1.py")
    # Create mesh (bench1.py:21-23)
    Lx = Ly = 200.0
    Nx = Ny = 100
    mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(Lx, Ly), Nx, Ny, "crossed")
    # Model setup (bench1.py:29-36): c_alpha, c_beta, rho_s, kappa, M
    dt = df.Constant(1e-1)
    c_alpha, c_beta, rho_s, kappa, M = 0.3, 0.7, 5.0, 2.0, 5.0
    # Initial conditions (bench1.py:48-53)
    epsilon, c0 = 0.05, 0.5
    w_ic = df.InitialConditionsBench1(c0, epsilon, degree=2)
    # F = cahn_hilliard_weak_form(...), J = derivative(F, w, dw), bcs = []   (bench1.py:59-69)
    problem = df.NonlinearVariationalProblem("bench1", mesh, dt, params=[c_alpha, c_beta, rho_s, kappa, M])
    solver = df.NonlinearVariationalSolver(problem)
    # Nonlinear solver setup (bench1.py:82-110)
    solver.parameters["nonlinear_solver"] = "snes"
    nlparams = solver.parameters["snes_solver"]
    nlparams["report"] = True
    nlparams["error_on_nonconvergence"] = False
    nlparams["absolute_tolerance"] = args.atol
    nlparams["maximum_iterations"] = 10
    nlparams["line_search"] = "basic"
    nlparams["linear_solver"] = "gmres"
    nlparams["preconditioner"] = "sor"          # mapped to the Schur/Chebyshev preconditioner
    nlparams["krylov_solver"]["maximum_iterations"] = 1000
    # time integration + CSV (bench1.py:128-217): dt0 = 1e-1, dt_min = 1e-2
    return time_loop(problem, solver, w_ic, 1e-1, 1e-2, args.end_time, os.path.join(args.outdir, "bench1", "stats.csv"),
                     "time,total_free_energy,total_solute", args.max_steps, args.quiet,
                     save_dir=os.path.join(args.outdir, os.path.basename(__file__)[:-3]) if args.save_solution else None)


if __name__ == "__main__":
    main()
