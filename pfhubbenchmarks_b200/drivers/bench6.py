"""PFHub benchmark 6, Cahn-Hilliard with electrostatics (c, mu, phi) -- mirrors dolfin/bench6.py.

    python -m pfhubbenchmarks_b200.drivers.bench6 [--backend b200|fenics] [--end-time 3.0]
"""
import os

from .. import api as df
from .common import parse_args, run_fenics, time_loop


def main():
    args = parse_args("bench6", 3.0)
    if args.backend == "fenics":
        return run_fenics(args, "bench6.py")
    mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(100.0, 100.0), 100, 100, "crossed")     # bench6.py:21-23
    dt = df.Constant(1e-1)
    w_ic = df.InitialConditionsBench6(0.5, 0.04, degree=2)
    problem = df.NonlinearVariationalProblem("bench6", mesh, dt)     # coefficients: engine.PROBLEMS["bench6"] (bench6.py:33-39)
    solver = df.NonlinearVariationalSolver(problem)
    solver.parameters["nonlinear_solver"] = "snes"
    nlparams = solver.parameters["snes_solver"]
    nlparams["report"] = True
    nlparams["error_on_nonconvergence"] = False
    nlparams["absolute_tolerance"] = args.atol
    nlparams["maximum_iterations"] = 10
    nlparams["line_search"] = "cp"        # as the reference (bench2.py:140 / bench6.py:117)
    nlparams["linear_solver"] = "gmres"
    nlparams["preconditioner"] = "sor"
    nlparams["krylov_solver"]["maximum_iterations"] = 5000
    # dt0 = 1e-2, dt_min = 1e-4 (bench6.py time-integration block)
    return time_loop(problem, solver, w_ic, 1e-2, 1e-4, args.end_time, os.path.join(args.outdir, "bench6_out.csv"),
                     "time,total_free_energy,total_solute", args.max_steps, args.quiet,
                     save_dir=os.path.join(args.outdir, os.path.basename(__file__)[:-3]) if args.save_solution else None)


if __name__ == "__main__":
    main()
