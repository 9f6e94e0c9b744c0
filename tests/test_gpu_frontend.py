"""GPU tests of the UFL-style front end: a problem built from the driver's own weak form runs on the same compiled
operators as the name-selected one (bit-identical), df.assemble evaluates the driver's functionals through the C ABI,
and a user-defined UserExpression interpolates like the built-in initial-condition kernel."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bench1(n=100, length=200.0):
    import pfhubbenchmarks_b200.dolfin as df
    from pfhubbenchmarks_b200.pfbase import cahn_hilliard_weak_form
    mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(length, length), n, n, "crossed")
    pe = df.FunctionSpace(mesh, "P", 1).ufl_element()
    W = df.FunctionSpace(mesh, df.MixedElement([pe, pe]))
    w, w0, w_, dw = df.Function(W), df.Function(W), df.TestFunction(W), df.TrialFunction(W)
    dt = df.Constant(1e-1)
    c_alpha, c_beta, rho_s, kappa, M = (df.Constant(v) for v in (0.3, 0.7, 5.0, 2.0, 5.0))
    c, mu = df.split(w)
    f_chem = rho_s * (c - c_alpha) ** 2 * (c_beta - c) ** 2
    F = cahn_hilliard_weak_form(c, mu, w_[0], w_[1], w0[0], dt, M, kappa, df.diff(f_chem, c))
    return df, mesh, w, w0, dt, f_chem, kappa, F, dw


def test_form_built_problem_equals_named_problem():
    import torch
    from pfhubbenchmarks_b200 import api
    from pfhubbenchmarks_b200.pfbase import InitialConditionsBench1, dx, grad, inner
    df, mesh, w, w0, dt, f_chem, kappa, F, dw = _bench1()
    ic = InitialConditionsBench1(0.5, 0.05, degree=2)
    w0.interpolate(ic)                                     # before the problem exists, as bench1.py:56 does
    problem = df.NonlinearVariationalProblem(F, w, [], df.derivative(F, w, dw))
    assert problem.benchmark == "bench1" and problem.params == [0.3, 0.7, 5.0, 2.0, 5.0]
    solver = df.NonlinearVariationalSolver(problem)
    solver.parameters["snes_solver"]["absolute_tolerance"] = 1e-6
    w.interpolate(ic)
    ref = api.NonlinearVariationalProblem("bench1", api.RectangleMesh(api.Point(0.0, 0.0), api.Point(200.0, 200.0), 100, 100, "crossed"),
                                          api.Constant(1e-1))
    rsolver = api.NonlinearVariationalSolver(ref)
    rsolver.parameters["snes_solver"]["absolute_tolerance"] = 1e-6
    ref.w.interpolate(ic)
    ref.w0.interpolate(ic)
    assert torch.equal(w0.tensor, ref.w0.tensor) and torch.equal(w.tensor, ref.w.tensor)
    for step in range(3):
        w0.assign(w)
        ref.w0.assign(ref.w)
        a, b = solver.solve(), rsolver.solve()
        assert a == b and a[1]
        assert torch.equal(w.tensor, ref.w.tensor)
    c = w[0]
    E = df.assemble(f_chem * dx + kappa / 2.0 * inner(grad(c), grad(c)) * dx)
    S = df.assemble(c * dx)
    assert (E, S) == ref.ctx.energy(ref.w.tensor)
    # a gradient piece on its own, and the same functional with another coefficient: u^T K u through pfh_bilinear_dev
    G = df.assemble(inner(grad(c), grad(c)) * dx)
    E3 = df.assemble(f_chem * dx + 3.0 * inner(grad(c), grad(c)) * dx)
    assert G > 0 and abs(E3 - (E + 2.0 * G)) < 1e-10 * abs(E3)
    problem.ctx.close()
    ref.ctx.close()


def test_bilinear_matches_oracle():
    """pfh_bilinear_dev: u_a^T (am M + ak K) u_b against the oracle's assembled matrices."""
    from oracle import fem, mesh as omesh
    from pfhubbenchmarks_b200 import engine
    ctx = engine.Context("bench2", (12, 9), (6.0, 5.4))
    w = ctx.initial_condition()
    sp = fem.P1Space(omesh.crossed2d(12, 9, 6.0, 5.4))
    u = w.cpu().numpy().reshape(6, -1)
    for a, b, am, ak in [(0, 0, 0.0, 1.0), (2, 3, 1.0, 0.0), (1, 4, 0.3, 1.7)]:
        want = u[a] @ (am * (sp.M @ u[b]) + ak * (sp.K @ u[b]))
        got = ctx.bilinear(w, a, b, am, ak)
        assert abs(got - want) <= 1e-12 * max(1.0, abs(want)), (a, b, got, want)
    with pytest.raises(RuntimeError):
        ctx.bilinear(w, 0, 6, 1.0, 0.0)
    ctx.close()
    ctx3 = engine.Context("b13d", (8, 6, 5), (4.0, 3.0, 2.5))
    w3 = ctx3.initial_condition()
    sp = fem.P1Space(omesh.kuhn3d(8, 6, 5, 4.0, 3.0, 2.5))
    u = w3.cpu().numpy().reshape(2, -1)
    want = u[0] @ (0.5 * (sp.M @ u[0]) + 2.0 * (sp.K @ u[0]))
    assert abs(ctx3.bilinear(w3, 0, 0, 0.5, 2.0) - want) <= 1e-12 * abs(want)
    ctx3.close()


def test_bench2_energy_in_pieces():
    """bench2.py:187-194 assembles the energy as one bulk + c-gradient term plus one term per order parameter: the
    sum of the pieces equals the fused energy kernel's value."""
    import itertools
    import pfhubbenchmarks_b200.dolfin as df
    from pfhubbenchmarks_b200.pfbase import (InitialConditionsBench2, allen_cahn_weak_form, cahn_hilliard_weak_form, dx, grad, h,
                                              inner)
    mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(200.0, 200.0), 100, 100, "crossed")
    pe = df.FunctionSpace(mesh, "P", 1).ufl_element()
    W = df.FunctionSpace(mesh, df.MixedElement([pe] * 6))
    w, w0, w_, dw = df.Function(W), df.Function(W), df.TestFunction(W), df.TrialFunction(W)
    dt = df.Constant(1e-2)
    c_alpha, c_beta, rho, kc, ke, M, ww, alpha, L = (df.Constant(v) for v in (0.3, 0.7, 2.0 ** 0.5, 3.0, 3.0, 5.0, 1.0, 5.0, 5.0))
    c, mu, *eta = df.split(w)
    H = sum(h(e) for e in eta)
    wells = sum(e ** 2 * (1 - e) ** 2 for e in eta) + alpha * sum(a ** 2 * b ** 2 for a, b in itertools.combinations(eta, 2))
    f = rho ** 2 * (c - c_alpha) ** 2 * (1 - H) + rho ** 2 * (c - c_beta) ** 2 * H + ww * wells
    F = cahn_hilliard_weak_form(c, mu, w_[0], w_[1], w0[0], dt, M, kc, df.diff(f, c))
    for i, e in enumerate(eta):
        F = F + allen_cahn_weak_form(e, w_[2 + i], w0[2 + i], dt, L, ke, df.diff(f, e), df.Constant(0))
    problem = df.NonlinearVariationalProblem(F, w, [], df.derivative(F, w, dw))
    assert problem.benchmark == "bench2"
    ic = InitialConditionsBench2(0.5, 0.05, 0.1, 1.5, degree=2)
    w.interpolate(ic)
    w0.interpolate(ic)
    solver = df.NonlinearVariationalSolver(problem)
    solver.parameters["snes_solver"]["line_search"] = "cp"
    w0.assign(w)
    assert solver.solve()[1]
    E = df.assemble(f * dx + kc / 2.0 * inner(grad(c), grad(c)) * dx)
    pieces = [df.assemble(ke / 2.0 * inner(grad(e), grad(e)) * dx) for e in eta]
    total = problem.ctx.energy(w.tensor)[0]
    assert all(p > 0 for p in pieces)
    assert abs(E + sum(pieces) - total) < 1e-11 * abs(total), (E, pieces, total)
    problem.ctx.close()


def test_user_expression_interpolates_like_the_kernel():
    """A UserExpression written as the reference's InitialConditionsBench1.eval (pfbase.py:177-193) is evaluated at the
    vertices on the host and uploaded: same vector as the device kernel (to libm round-off)."""
    df, mesh, w, w0, dt, f_chem, kappa, F, dw = _bench1(n=24, length=48.0)
    from pfhubbenchmarks_b200.pfbase import InitialConditionsBench1

    class Mine(df.UserExpression):
        def eval(self, values, x):
            c0, eps = 0.5, 0.05
            values[0] = c0 + eps * (np.cos(0.105 * x[0]) * np.cos(0.11 * x[1]) + (np.cos(0.13 * x[0]) * np.cos(0.087 * x[1])) ** 2
                                    + np.cos(0.025 * x[0] - 0.15 * x[1]) * np.cos(0.07 * x[0] - 0.02 * x[1]))
            values[1] = 0.0

        def value_shape(self):
            return (2,)

    problem = df.NonlinearVariationalProblem(F, w, [], df.derivative(F, w, dw))
    w.interpolate(Mine(degree=2))
    w0.interpolate(InitialConditionsBench1(0.5, 0.05, degree=2))
    assert float((w.tensor - w0.tensor).abs().max()) < 1e-14
    problem.ctx.close()


def test_peval_and_sample():
    """pfbase.peval / sample (pfbase.py:50-94) through pfh_eval_points_dev: vertex values are reproduced, a lattice
    sample equals the oracle's cell search, outside points are +inf (peval) or raise (f(x))."""
    from oracle import fem, mesh as omesh
    from pfhubbenchmarks_b200.pfbase import InitialConditionsBench1, peval, sample
    df, mesh, w, w0, dt, f_chem, kappa, F, dw = _bench1(n=24, length=48.0)
    problem = df.NonlinearVariationalProblem(F, w, [], df.derivative(F, w, dw))
    w.interpolate(InitialConditionsBench1(0.5, 0.05, degree=2))
    c = w.sub(0)
    u = w.to_numpy()[0]
    m = omesh.crossed2d(24, 24, 48.0, 48.0)
    assert np.max(np.abs(c.at(m.coords) - u)) < 1e-14                       # all vertices (corners and centres)
    xs, ys, us = sample(c, 31, 17, 48.0, 48.0)
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    want = fem.eval_points(m, u, np.stack([X.ravel(), Y.ravel()], 1)).reshape(31, 17)
    assert us.shape == (31, 17) and np.max(np.abs(us - want)) < 1e-13
    assert abs(c(np.array([10.3, 7.9])) - fem.eval_points(m, u, [[10.3, 7.9]])[0]) < 1e-14
    assert np.isinf(peval(c, np.array([50.0, 1.0]))[0])
    # `c, _ = w.split(); cfile << (c, t)` with a File opened without a mesh, as the reference does (bench2.py:180,269)
    import tempfile
    from oracle import golden_io
    with tempfile.TemporaryDirectory() as tmp:
        cfile = df.File(tmp + "/conc.pvd", "compressed")
        c_split, _ = w.split()
        cfile << (c_split, 0.5)
        (t0, fn), = golden_io.read_pvd(tmp + "/conc.pvd")
        assert t0 == 0.5 and np.array_equal(golden_io.read_vtu_scalar(tmp + "/" + fn), u)
    with pytest.raises(RuntimeError):
        c(np.array([-1.0, 1.0]))
    problem.ctx.close()
    # 3-D: the Kuhn tets
    from pfhubbenchmarks_b200 import engine
    ctx = engine.Context("b13d", (9, 7, 6), (18.0, 14.0, 12.0))
    w3 = ctx.initial_condition()
    m3 = omesh.kuhn3d(9, 7, 6, 18.0, 14.0, 12.0)
    pts = np.random.default_rng(3).random((400, 3)) * [18.0, 14.0, 12.0]
    got = ctx.eval_points(w3, 0, pts)
    assert np.max(np.abs(got - fem.eval_points(m3, w3.cpu().numpy()[:m3.nv], pts))) < 1e-13
    ctx.close()


def test_module_level_script_end_to_end(tmp_path):
    """`python -m pfhubbenchmarks_b200.run_script tests/scripts/spinodal_module_level.py`: a reference-style program
    (import dolfin as df / from pfbase import *, inline time loop, HDF5File, split() components inside assemble) runs
    unchanged; its CSV equals the name-selected API stepping the same mesh with the same dt sequence."""
    import os
    import subprocess
    import sys
    from oracle import golden_io
    from pfhubbenchmarks_b200 import engine
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pfhubbenchmarks_b200.run_script", os.path.join(root, "tests", "scripts", "spinodal_module_level.py"),
                        str(tmp_path)], cwd=root, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = np.loadtxt(str(tmp_path / "stats.csv"), delimiter=",", skiprows=1)
    assert rows.shape == (3, 3) and np.allclose(rows[:, 0], [0.1, 0.3, 0.7])
    ctx = engine.Context("bench1", (24, 24), (48.0, 48.0))
    w = ctx.initial_condition()
    for k, h in enumerate((0.1, 0.2, 0.4)):
        w0 = w.clone()
        ctx.dt = h
        assert ctx.newton_solve(w, w0)["converged"]
        E, S = ctx.energy(w)
        assert abs(rows[k, 1] - E) <= 1e-9 * abs(E) and abs(rows[k, 2] - S) <= 1e-9 * abs(S)
    entries = golden_io.read_pvd(str(tmp_path / "conc.pvd"))              # the HDF5File stand-in
    assert [t for t, _ in entries] == [0.1, pytest.approx(0.3), pytest.approx(0.7)]
    last = golden_io.read_vtu_scalar(str(tmp_path / entries[-1][1]))
    assert np.array_equal(last, w.cpu().numpy()[:ctx.n_owned])
    ctx.close()
