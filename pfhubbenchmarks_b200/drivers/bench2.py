"""PFHub benchmark 2, Ostwald ripening (c, mu, eta1..4) -- mirrors dolfin/bench2.py.

    python -m pfhubbenchmarks_b200.drivers.bench2 [--backend b200|fenics] [--end-time 100.0]
"""
import os

from .. import api as df
from .common import parse_args, run_fenics, time_loop


def main():
    args = parse_args("bench2", 100.0)
    if args.backend == "fenics":
        return run_fenics(args, "bench2.py")
    mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(200.0, 200.0), 100, 100, "crossed")     # bench2.py:21-23
    dt = df.Constant(1e-1)
    w_ic = df.InitialConditionsBench2(0.5, 0.05, 0.1, 1.5, degree=2)
    problem = df.NonlinearVariationalProblem("bench2", mesh, dt)     # coefficients: engine.PROBLEMS["bench2"] (bench2.py:34-42)
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
    nlparams["krylov_solver"]["maximum_iterations"] = 1000
    # dt0 = 1e-2, dt_min = 1e-4 (bench2.py time-integration block)
    return time_loop(problem, solver, w_ic, 1e-2, 1e-4, args.end_time, os.path.join(args.outdir, "bench2_out.csv"),
                     "time,total_free_energy,total_solute", args.max_steps, args.quiet,
                     save_dir=os.path.join(args.outdir, os.path.basename(__file__)[:-3]) if args.save_solution else None)


if __name__ == "__main__":
    main()
