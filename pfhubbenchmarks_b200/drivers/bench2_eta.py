"""The reduced Ostwald-ripening models of dolfin/bench2_eta1.py / bench2_eta2.py: benchmark 2's free energy with one or
two order parameters instead of four (fields c, mu, eta1[, eta2]).

    python -m pfhubbenchmarks_b200.drivers.bench2_eta [--num-eta 1|2] [--end-time 20.0] [--max-steps N]

These weak forms are not among the four shipped systems, so the problem constructor generates their functor and builds
their copy of the library (codegen.py); the initial data are the host-evaluated ICB2_jank1 / ICB2_jank2.
"""
import itertools
import os
import sys

from .. import dolfin as df
from ..pfbase import ICB2_jank1, ICB2_jank2, allen_cahn_weak_form, cahn_hilliard_weak_form, dx, grad, h, inner
from .common import mixed_p1, parse_args, snes_setup, time_loop


def setup(mesh, num_eta, compile_only=False):
    dt = df.Constant(1e-1)
    # bench2.py:34-42
    c_alpha, c_beta, rho, kappa_c, kappa_eta, M, ww, alpha, L = (
        df.Constant(v) for v in (0.3, 0.7, 2.0 ** 0.5, 3.0, 3.0, 5.0, 1.0, 5.0, 5.0))
    W, w, w0, w_, dw = mixed_p1(mesh, 2 + num_eta)
    c, mu, *eta = df.split(w)
    interp = sum(h(e) for e in eta)
    wells = sum(e ** 2 * (1 - e) ** 2 for e in eta) + alpha * sum(a ** 2 * b ** 2 for a, b in itertools.combinations(eta, 2))
    f_chem = rho ** 2 * (c - c_alpha) ** 2 * (1 - interp) + rho ** 2 * (c - c_beta) ** 2 * interp + ww * wells
    F = cahn_hilliard_weak_form(c, mu, w_[0], w_[1], w0[0], dt, M, kappa_c, df.diff(f_chem, c))
    for i, e in enumerate(eta):
        F = F + allen_cahn_weak_form(e, w_[2 + i], w0[2 + i], dt, L, kappa_eta, df.diff(f_chem, e), df.Constant(0))
    E = f_chem * dx + kappa_c / 2.0 * inner(grad(c), grad(c)) * dx
    for e in eta:
        E = E + kappa_eta / 2.0 * inner(grad(e), grad(e)) * dx
    S = c * dx
    J = df.derivative(F, w, dw)
    if compile_only:
        return df.compile_only(F, w, [], J, functionals=(E, S))
    problem = df.NonlinearVariationalProblem(F, w, [], J, functionals=(E, S))
    return problem, df.NonlinearVariationalSolver(problem), (lambda: df.assemble(E)), (lambda: df.assemble(S))


def main():
    num_eta = 1
    if "--num-eta" in sys.argv:
        i = sys.argv.index("--num-eta")
        num_eta = int(sys.argv[i + 1])
        del sys.argv[i:i + 2]
    assert num_eta in (1, 2), "bench2_eta1.py / bench2_eta2.py: one or two order parameters (four is drivers.bench2)"
    args = parse_args("bench2_eta%d" % num_eta, 20.0)
    mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(200.0, 200.0), 100, 100, "crossed")
    problem, solver, total_free_energy, total_solute = setup(mesh, num_eta)
    snes_setup(solver, args.atol, "bt", 1000, linear_solver="bicgstab")      # bench2_eta1.py:107,113
    w_ic = (ICB2_jank1 if num_eta == 1 else ICB2_jank2)(0.5, 0.05, 0.1, 1.5, degree=2)
    return time_loop(problem, solver, w_ic, 1e-2, 1e-4, args.end_time, os.path.join(args.outdir, "bench2_eta%d_out.csv" % num_eta),
                     "time,total_free_energy,total_solute", args.max_steps, args.quiet, functionals=(total_free_energy, total_solute))


if __name__ == "__main__":
    main()
