"""CPU tests of the UFL-style front end (pfhubbenchmarks_b200/forms.py, dolfin.py, pfbase.py): the weak forms the
reference drivers write (dolfin/bench1.py:59-69, bench2.py:66-116, bench3.py:67-104, bench6.py:60-92) are lowered,
recognised as the compiled operator family they are, and give back the coefficients the driver put in."""
import importlib
import sys

import numpy as np
import pytest
import sympy as sy

import pfhubbenchmarks_b200.dolfin as df
from pfhubbenchmarks_b200 import engine, forms
from pfhubbenchmarks_b200.pfbase import (allen_cahn_weak_form, cahn_hilliard_weak_form, dx, grad, inner,
                                          poisson_weak_form)


class _Stop(Exception):
    pass


def _lower(F, w, bcs=()):
    prev = next(s.owner for t in F.terms if t[0] == "src" for s in sy.sympify(t[2]).free_symbols
                if isinstance(s, forms.FieldSymbol) and s.owner is not w)
    return forms.compile_form(F, w.W.nfields, w.symbols, prev.symbols, bcs)


@pytest.mark.parametrize("name", ["bench1", "bench2", "bench3", "bench6", "b13d"])
def test_driver_forms_are_recognised(name, monkeypatch):
    """Run each driver's main() up to the problem constructor: the form it built maps to its operator family with
    exactly the coefficients of the reference script (engine.PROBLEMS holds the same numbers for the name-selected API)."""
    got = {}

    def fake(F, w, bcs=None, J=None, **kw):
        assert J == ("derivative", F)
        got["r"] = _lower(F, w, bcs or ())
        raise _Stop

    monkeypatch.setattr(df, "NonlinearVariationalProblem", fake)
    monkeypatch.setattr(sys, "argv", [name])
    mod = importlib.import_module("pfhubbenchmarks_b200.drivers." + name)
    with pytest.raises(_Stop):
        mod.main()
    family, params, dt = got["r"]
    assert family == ("bench1" if name == "b13d" else name)
    want = engine.PROBLEMS[name][2]
    assert len(params) == len(want)
    for a, b in zip(params, want):
        assert (a is None and b is None) or abs(a - b) <= 2e-16 * abs(b), (params, want)
    assert float(dt) == 0.1


def _ch(n=2):
    mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(10.0, 10.0), 4, 4, "crossed")
    pe = df.FunctionSpace(mesh, "P", 1).ufl_element()
    W = df.FunctionSpace(mesh, df.MixedElement([pe] * n))
    return W, df.Function(W), df.Function(W), df.TestFunction(W)


def test_coefficients_come_from_the_form():
    W, w, w0, w_ = _ch()
    dt, M, kappa = df.Constant(0.25), df.Constant(1.5), df.Constant(0.75)
    c, mu = df.split(w)
    f = 3.0 * (c - 0.2) ** 2 * (0.9 - c) ** 2
    F = cahn_hilliard_weak_form(c, mu, w_[0], w_[1], w0[0], dt, M, kappa, df.diff(f, c))
    family, params, dtc = _lower(F, w)
    assert family == "bench1" and dtc is dt
    assert params == pytest.approx([0.2, 0.9, 3.0, 0.75, 1.5], rel=1e-13)
    # split(w)[i] and w[i] are the same symbols (bench1.py:59 vs :68)
    assert c is w[0] and mu is w[1]


def test_unknown_free_energy_is_refused():
    W, w, w0, w_ = _ch()
    dt, one = df.Constant(0.1), df.Constant(1.0)
    c, mu = df.split(w)
    F = cahn_hilliard_weak_form(c, mu, w_[0], w_[1], w0[0], dt, one, one, df.diff(c ** 6 - c ** 2, c))
    with pytest.raises(NotImplementedError, match="compiled operator families"):
        _lower(F, w)
    # an Allen-Cahn equation alone is none of the shipped systems either
    W1, v, v0, v_ = _ch(1)
    e = v[0]
    with pytest.raises(NotImplementedError):
        _lower(allen_cahn_weak_form(e, v_[0], v0[0], dt, one, one, df.diff(e ** 2 * (1 - e) ** 2, e), df.Constant(0)), v)


def test_form_algebra_matches_the_weak_forms():
    """The lowered rows of the Cahn-Hilliard form: (c - c0)/dt and mu - f'(c) pointwise, M and -kappa stiffness couplings."""
    W, w, w0, w_ = _ch()
    dt, M, kappa = df.Constant(0.1), df.Constant(5.0), df.Constant(2.0)
    c, mu = df.split(w)
    dfdc = df.diff((c - 0.3) ** 2 * (0.7 - c) ** 2, c)
    F = cahn_hilliard_weak_form(c, mu, w_[0], w_[1], w0[0], dt, M, kappa, dfdc)
    L = forms.Lowered(F, 2, w.symbols, w0.symbols, dt.sym)
    assert sy.simplify(L.src[0] - (c - w0[0]) / dt.sym) == 0
    assert forms._close(L.src[1] - (mu - dfdc), 1e-14)
    assert L.stiff[0][1] == M.sym and L.stiff[1][0] == -kappa.sym
    assert L.stiff[0][0] == 0 and L.stiff[1][1] == 0
    # Poisson row: -int grad v . grad u - int v f
    Wp, u, u0, u_ = _ch(3)
    P = poisson_weak_form(u[2], u_[2], -0.5 * u[0], df.Constant(1.0))
    Lp = forms.Lowered(P, 3, u.symbols, u0.symbols, dt.sym)
    assert float(Lp.values(Lp.stiff[2][2])) == -1.0 and forms._is_zero(Lp.src[2] - 0.5 * u[0])


def test_dirichlet_boundaries():
    W, w, w0, w_ = _ch(3)
    sub = W.split()[2]
    left = df.DirichletBC(sub, df.Constant(0.0), lambda x, on: on and df.near(x[0], 0.0, 1e-12))
    right = df.DirichletBC(sub, df.Expression("sin(x[1]/7)", degree=2), lambda x, on: on and df.near(x[0], 10.0, 1e-12))
    assert (left.side, left.field, left.value_expr) == ("left", 2, 0)
    assert right.side == "right" and str(right.value_expr) == "sin(x1/7)"
    whole = df.DirichletBC(sub, df.Constant(0.0), lambda x, on: on)     # several edges in one condition (nl.py:34 has two):
    assert set(whole.sides) == {"left", "right", "bottom", "top"}       # accepted here, refused by the form matchers
    with pytest.raises(NotImplementedError):
        df.DirichletBC(sub, df.Constant(0.0), lambda x, on: True)        # interior points too
    with pytest.raises(NotImplementedError):
        df.DirichletBC(W, df.Constant(0.0), lambda x, on: on and df.near(x[0], 0.0))


def test_functional_plans():
    """df.assemble plans: the fused energy kernel gives the benchmark's two integrals; pieces with other gradient
    coefficients are corrected by u^T K u terms (bench2.py:187-194 assembles the energy in five pieces)."""
    W, w, w0, w_ = _ch(6)
    p = engine.PROBLEMS["bench2"][2]
    c, mu, *eta = df.split(w)
    H = sum(e ** 3 * (6 * e ** 2 - 15 * e + 10) for e in eta)
    Wd = sum(e ** 2 * (1 - e) ** 2 for e in eta) + p[7] * sum(eta[i] ** 2 * eta[j] ** 2 for i in range(4) for j in range(i + 1, 4))
    f = p[2] * (c - p[0]) ** 2 * (1 - H) + p[2] * (c - p[1]) ** 2 * H + p[6] * Wd
    kc, ke = df.Constant(p[3]), df.Constant(p[4])
    which, scale, corr = forms.plan_functional(f * dx + kc / 2.0 * inner(grad(c), grad(c)) * dx, "bench2", p, w.symbols)
    assert (which, scale) == (0, 1.0) and corr == [(k, pytest.approx(-1.5)) for k in (2, 3, 4, 5)]
    assert forms.plan_functional(ke / 2.0 * inner(grad(eta[1]), grad(eta[1])) * dx, "bench2", p, w.symbols) == (None, 1.0, [(3, 1.5)])
    assert forms.plan_functional(c * dx, "bench2", p, w.symbols) == (1, 1.0, [])
    with pytest.raises(NotImplementedError):
        forms.plan_functional(c ** 2 * dx, "bench2", p, w.symbols)
    with pytest.raises(NotImplementedError):
        forms.plan_functional(inner(grad(c), grad(mu)) * dx, "bench2", p, w.symbols)


def test_quadrature_degree_is_checked():
    W, w, w0, w_ = _ch()
    dt, one = df.Constant(0.1), df.Constant(1.0)
    c, mu = df.split(w)
    F = cahn_hilliard_weak_form(c, mu, w_[0], w_[1], w0[0], dt, one, one, df.diff((c - 0.3) ** 2 * (0.7 - c) ** 2, c))
    df.parameters["form_compiler"]["quadrature_degree"] = 2
    try:
        with pytest.raises(NotImplementedError, match="degree-3"):
            df.NonlinearVariationalProblem(F, w, [], df.derivative(F, w, df.TrialFunction(W)))
    finally:
        df.parameters["form_compiler"]["quadrature_degree"] = 3
    with pytest.raises(NotImplementedError, match="derivative"):
        df.NonlinearVariationalProblem(F, w, [], J="something else")


def test_function_storage_is_deferred():
    W, w, w0, w_ = _ch()
    with pytest.raises(RuntimeError, match="no storage"):
        w.tensor
    with pytest.raises(NotImplementedError):
        df.Constant((1.0, 0.0))
    assert isinstance(df.Constant((0.0, 0.0)), forms.VectorZero)
    with pytest.raises(FileNotFoundError):
        df.HDF5File(None, "no_such_series.h5", "r")      # mode "r" reads the stand-in's own series (tests/test_io.py)
    with pytest.raises(NotImplementedError):
        df.HDF5File(None, "x.h5", "a")
    with pytest.warns(UserWarning, match="HDF5 is not available"):
        h5 = df.HDF5File(None, "results/conc.h5", "w")          # stands in with a VTU series next to the requested file
    assert h5._pvd.path == "results/conc.pvd"


def test_problem_constructor_binds_functions(monkeypatch):
    """The whole df.NonlinearVariationalProblem constructor with the GPU context stubbed out: benchmark selection (2-D
    -> bench1, BoxMesh -> b13d), coefficients handed to the context, w / w0 bound, an interpolate() issued before the
    problem existed replayed on the new storage."""
    from pfhubbenchmarks_b200.pfbase import InitialConditionsBench1
    calls = []

    class StubContext:
        def __init__(self, problem, ncell, length, **kw):
            calls.append(("ctx", problem, tuple(ncell), kw["params"]))
            self.nf, self.params, self.problem = 2, kw["params"], problem

        def new_vector(self):
            return object()

        def initial_condition(self, args, out=None):
            calls.append(("ic", list(args)))

    monkeypatch.setattr(engine, "Context", StubContext)
    for mesh, want in [(df.RectangleMesh(df.Point(0.0, 0.0), df.Point(8.0, 8.0), 4, 4, "crossed"), "bench1"),
                       (df.BoxMesh(df.Point(0.0, 0.0, 0.0), df.Point(8.0, 8.0, 8.0), 4, 4, 4), "b13d")]:
        del calls[:]
        pe = df.FunctionSpace(mesh, "P", 1).ufl_element()
        W = df.FunctionSpace(mesh, df.MixedElement([pe, pe]))
        w, w0, w_, dw = df.Function(W), df.Function(W), df.TestFunction(W), df.TrialFunction(W)
        dt = df.Constant(0.1)
        c, mu = df.split(w)
        w0.interpolate(InitialConditionsBench1(0.5, 0.05, degree=2))
        F = cahn_hilliard_weak_form(c, mu, w_[0], w_[1], w0[0], dt, df.Constant(5.0), df.Constant(2.0),
                                    df.diff(5.0 * (c - 0.3) ** 2 * (0.7 - c) ** 2, c))
        p = df.NonlinearVariationalProblem(F, w, [], df.derivative(F, w, dw))
        assert p.benchmark == want and p.model is None and p.w is w and p.w0 is w0 and p.dt is dt
        assert calls == [("ctx", want, mesh.ncell, [0.3, 0.7, 5.0, 2.0, 5.0]), ("ic", [0.5, 0.05])]


def test_namespace_covers_the_reference_drivers():
    """Every name the reference drivers take from `import dolfin as df` / `from pfbase import *` exists here
    (bench1.py, bench2.py, bench3.py, bench6.py, b13d.py and the weak-form section of pfbase.py)."""
    import pfhubbenchmarks_b200.pfbase as pb
    for name in ["parameters", "RectangleMesh", "BoxMesh", "Point", "Constant", "FunctionSpace", "MixedElement", "Function",
                 "TrialFunction", "TestFunction", "split", "variable", "diff", "derivative", "set_log_level", "LogLevel",
                 "NonlinearVariationalProblem", "NonlinearVariationalSolver", "assemble", "File", "MPI", "DOLFIN_EPS", "grad",
                 "inner", "atan", "near", "Expression", "DirichletBC", "UserExpression", "SubDomain", "warning", "HDF5File"]:
        assert hasattr(df, name), name
    for name in ["InitialConditionsBench1", "InitialConditionsBench2", "InitialConditionsBench3", "InitialConditionsBench6",
                 "cahn_hilliard_weak_form", "allen_cahn_RHS_IBP", "allen_cahn_weak_form", "poisson_weak_form",
                 "diffusion_weak_form_RHS", "diffusion_weak_form", "euler_bwd_weak_form", "peval", "sample", "dx", "inner", "grad",
                 "variable", "MPI", "HDF5File", "near", "UserExpression"]:
        assert hasattr(pb, name), name
    assert df.MPI.comm_world.rank == 0 and df.MPI.comm_world.size == 1 and df.MPI.rank(None) == 0
    assert df.LogLevel.WARNING < df.LogLevel.ERROR
    df.parameters["form_compiler"]["cpp_optimize_flags"] = "-O3 -ffast-math"      # accepted and ignored (bench1.py:12-13)
    del df.parameters["form_compiler"]["cpp_optimize_flags"]


def test_module_level_script_runs_through_the_aliases(monkeypatch, tmp_path):
    """run_script: a reference-style module-level program (`import dolfin as df`, `from pfbase import *`, HDF5File,
    `c, _ = w.split()` re-used inside assemble, inline adaptive loop) executes unchanged.  The GPU context is replaced by
    a stand-in holding CPU tensors, so that everything around the C ABI calls runs here (the GPU suite runs it for real)."""
    import os
    import torch
    from pfhubbenchmarks_b200 import io, run_script
    log = []

    class StubContext:
        nranks = 1

        def __init__(self, problem, ncell, length, **kw):
            log.append(("ctx", problem, tuple(ncell), kw["params"]))
            self.nf, self.params, self.dt, self.problem = 2, kw["params"], None, problem
            self.n_owned = (ncell[0] + 1) * (ncell[1] + 1) + ncell[0] * ncell[1]

        def new_vector(self):
            return torch.zeros(self.nf * self.n_owned, dtype=torch.float64)

        def fields(self, w):
            return [w[i * self.n_owned:(i + 1) * self.n_owned] for i in range(self.nf)]

        def initial_condition(self, args, out=None):
            out[:self.n_owned] = args[0]
            log.append(("ic", list(args)))

        def newton_solve(self, w, w0, opts=None):
            w[:self.n_owned] += self.dt                      # something that depends on dt, to see the loop's dt sequence
            log.append(("solve", self.dt, opts.line_search, opts.absolute_tolerance))
            return dict(niters=2, converged=1, reason=2)

        def energy(self, w):
            return float(w[0]) * 10.0, float(w[0])

    monkeypatch.setattr(engine, "Context", StubContext)
    saved = {k: sys.modules.get(k) for k in ("dolfin", "pfbase", "matplotlib", "matplotlib.pyplot")}
    try:
        with pytest.warns(UserWarning, match="HDF5 is not available"):
            run_script.run(os.path.join(os.path.dirname(__file__), "scripts", "spinodal_module_level.py"), [str(tmp_path)])
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    assert log[0] == ("ctx", "bench1", (24, 24), [0.3, 0.7, 5.0, 2.0, 5.0])
    solves = [e for e in log if e[0] == "solve"]
    assert [e[1] for e in solves] == pytest.approx([0.1, 0.2, 0.4]) and all(e[2] == 0 and e[3] == 1e-6 for e in solves)
    rows = np.loadtxt(str(tmp_path / "stats.csv"), delimiter=",", skiprows=1)
    assert np.allclose(rows[:, 0], [0.1, 0.3, 0.7]) and np.allclose(rows[:, 2], [0.6, 0.8, 1.2]) and np.allclose(rows[:, 1], 10 * rows[:, 2])
    times, vals = io.read_series(str(tmp_path / "conc.pvd"))                   # what the HDF5File stand-in wrote
    assert np.allclose(times, [0.1, 0.3, 0.7]) and vals.shape == (3, 25 * 25 + 24 * 24) and np.allclose(vals[:, 0], [0.6, 0.8, 1.2])


def test_reassigned_coefficient_is_refused(monkeypatch):
    """Coefficients are read from the form when the problem is created: re-assigning one later (DOLFIN would pick it
    up) must not be silently ignored."""
    class StubContext:
        def __init__(self, problem, ncell, length, **kw):
            self.nf, self.params, self.dt = 2, kw["params"], None

        def new_vector(self):
            return object()

        def newton_solve(self, w, w0, opts=None):
            return dict(niters=2, converged=1, reason=2)

    monkeypatch.setattr(engine, "Context", StubContext)
    W, w, w0, w_ = _ch()
    dt, M, kappa = df.Constant(0.1), df.Constant(5.0), df.Constant(2.0)
    c, mu = df.split(w)
    F = cahn_hilliard_weak_form(c, mu, w_[0], w_[1], w0[0], dt, M, kappa, df.diff(5.0 * (c - 0.3) ** 2 * (0.7 - c) ** 2, c))
    problem = df.NonlinearVariationalProblem(F, w, [], df.derivative(F, w, df.TrialFunction(W)))
    solver = df.NonlinearVariationalSolver(problem)
    assert solver.solve() == (2, True)
    dt.assign(0.2)                       # the time step may change
    assert solver.solve() == (2, True) and problem.ctx.dt == 0.2
    kappa.assign(3.0)                    # anything else may not
    with pytest.raises(RuntimeError, match="only dt may change"):
        solver.solve()


def test_vertex_coordinates_follow_the_dolfin_order():
    """engine.Context.vertex_coordinates (used to evaluate user-written UserExpressions on the host) lists vertices in the
    library's order: the oracle's crossed / Kuhn meshes, and the owned planes only for a z-slab."""
    import types
    from oracle import mesh as omesh
    ctx = types.SimpleNamespace(dim=2, ncell=(7, 5), length=(14.0, 10.0))
    assert np.allclose(engine.Context.vertex_coordinates(ctx), omesh.crossed2d(7, 5, 14.0, 10.0).coords, atol=1e-14)
    full = omesh.kuhn3d(4, 3, 5, 8.0, 6.0, 10.0).coords
    ctx = types.SimpleNamespace(dim=3, ncell=(4, 3, 5), length=(8.0, 6.0, 10.0), z_range=(0, 6))
    assert np.allclose(engine.Context.vertex_coordinates(ctx), full, atol=1e-14)
    ctx.z_range = (2, 4)
    assert np.allclose(engine.Context.vertex_coordinates(ctx), full[2 * 20:4 * 20], atol=1e-14)


class _TensorStubContext:
    """Stand-in for engine.Context holding CPU tensors: lets the Python side of a driver (model setup, recognition, time
    loop, df.assemble plans, CSV / VTU output) run without a GPU.  Numbers are fake; the calls are recorded."""
    nranks = 1
    calls = []

    def __init__(self, problem, ncell, length, **kw):
        import torch
        self._torch = torch
        self.problem = problem
        self.nf = engine.PROBLEMS[problem][1]
        self.field_names = engine.PROBLEMS[problem][4]
        self.params, self.dt = list(kw["params"]), None
        if problem == "bench3" and self.params[4] is None:        # as engine.Context: the area the solid fraction divides by
            self.params[4] = float(length[0] * length[1])
        n = 1
        for c in ncell:
            n *= c + 1
        self.n_owned = n + (ncell[0] * ncell[1] if len(ncell) == 2 else 0)
        type(self).calls.append(("ctx", problem, tuple(ncell)))

    def new_vector(self):
        return self._torch.zeros(self.nf * self.n_owned, dtype=self._torch.float64)

    def fields(self, w):
        return [w[i * self.n_owned:(i + 1) * self.n_owned] for i in range(self.nf)]

    def initial_condition(self, args, out=None):
        out[:] = 1.0

    def newton_solve(self, w, w0, opts=None):
        w -= 0.01 * self.dt
        type(self).calls.append(("solve", self.dt, opts.line_search))
        return dict(niters=3, converged=1, reason=2)

    def energy(self, w):
        return 100.0 * float(w[0]), 7.0

    def bilinear(self, w, a, b, am, ak):
        type(self).calls.append(("bilinear", a, b, am, ak))
        return 0.5


@pytest.mark.parametrize("name,csv,ls,nbil", [("bench1", "bench1/stats.csv", 0, 0), ("bench2", "bench2_out.csv", 1, 8), ("bench3", "bench3_out.csv", 0, 0),
                                              ("bench6", "bench6_out.csv", 1, 0), ("b13d", "bench1_out.csv", 0, 0)])
def test_driver_mains_run_on_a_stub_context(monkeypatch, tmp_path, name, csv, ls, nbil):
    """Every shipped driver's main() end to end on the stand-in context: the model is recognised, the solver gets the
    driver's line search, the driver's own df.assemble functionals are planned (bench2's energy in five pieces: the four
    order-parameter pieces go through pfh_bilinear_dev once as corrections and once on their own), the CSV is written."""
    _TensorStubContext.calls = []
    monkeypatch.setattr(engine, "Context", _TensorStubContext)
    argv = [name, "--max-steps", "2", "--quiet", "--outdir", str(tmp_path)] + (["--n", "6"] if name == "b13d" else [])
    monkeypatch.setattr(sys, "argv", argv)
    rows, spent = importlib.import_module("pfhubbenchmarks_b200.drivers." + name).main()
    calls = _TensorStubContext.calls
    assert calls[0][:2] == ("ctx", name)
    solves = [c for c in calls if c[0] == "solve"]
    assert len(solves) == 2 and all(c[2] == ls for c in solves)
    assert len([c for c in calls if c[0] == "bilinear"]) == nbil * 2
    back = np.loadtxt(str(tmp_path / csv), delimiter=",", skiprows=1)
    assert back.shape == (2, 3) and np.allclose(back, np.array(rows), atol=1e-9)
    if name == "bench3":
        assert np.allclose(back[:, 2], 7.0)                  # solid fraction: assemble(...) / (Lx Ly) undoes the area factor
