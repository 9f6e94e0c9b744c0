"""PFHub benchmark 3, dendritic growth (U, phi), isotropic as coded -- mirrors dolfin/bench3.py.

    python -m pfhubbenchmarks_b200.drivers.bench3 [--backend b200|fenics] [--end-time 100.0]
"""
import os

from .. import api as df
from .common import parse_args, run_fenics, time_loop


def main():
    args = parse_args("bench3", 100.0)
    if args.backend == "fenics":
        return run_fenics(args, "bench3.py")
    mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(960.0, 960.0), 350, 350, "crossed")     # bench3.py:21-23
    dt = df.Constant(1e-1)
    w_ic = df.InitialConditionsBench3(-0.3, 8.0, 1.0, 1.0, -1.0, degree=2)
    problem = df.NonlinearVariationalProblem("bench3", mesh, dt)     # coefficients: engine.PROBLEMS["bench3"] (bench3.py:32-38)
    solver = df.NonlinearVariationalSolver(problem)
    solver.parameters["nonlinear_solver"] = "snes"
    nlparams = solver.parameters["snes_solver"]
    nlparams["report"] = True
    nlparams["error_on_nonconvergence"] = False
    nlparams["absolute_tolerance"] = args.atol
    nlparams["maximum_iterations"] = 10
    nlparams["line_search"] = "basic"        # 'basic' in the reference; only the full Newton step is implemented
    nlparams["linear_solver"] = "gmres"
    nlparams["preconditioner"] = "sor"
    nlparams["krylov_solver"]["maximum_iterations"] = 1000
    # dt0 = 1e-2, dt_min = 1e-4 (bench3.py time-integration block)
    return time_loop(problem, solver, w_ic, 1e-2, 1e-4, args.end_time, os.path.join(args.outdir, "bench3_out.csv"),
                     "time,total_free_energy,solid_fraction", args.max_steps, args.quiet, abort_on_negative_energy=True,
                     save_dir=os.path.join(args.outdir, 'bench3') if args.save_solution else None, save_fields=(0, 1))


if __name__ == "__main__":
    main()
