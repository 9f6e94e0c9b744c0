"""CPU tests of the nl.py path (dolfin/nl.py of the reference: steady quasi-linear diffusion on the right-diagonal
UnitSquareMesh, SNES then Newton).

The reference records no output of that script (no expected values), so the oracle (oracle/nl.py) is
"parity unpinned" and is checked here on the discrete problem itself -- Jacobian against differences of the residual,
second-order convergence to a manufactured solution -- and the product is then compared with it on the CPU legs: the
__host__ __device__ per-vertex / per-square operators of csrc/pfh_nl.cuh (tests/emul) and the unchanged host solver
csrc/pfh_solver.cpp over loop launchers.  The GPU launch layer is covered by tests/test_gpu_nl.py.
"""
import importlib
import os
import sys

import numpy as np
import pytest
import sympy as sy

import emul
import pfhubbenchmarks_b200.dolfin as df
from oracle.nl import NlProblem
from pfhubbenchmarks_b200 import _lib, engine, forms

PRM = [5.0, 1.0, 1.0, 0.9]
REF_NL = "/root/reference/dolfin/nl.py"


# ---------------------------------------------------------------------------------------------------------------------
# oracle
# ---------------------------------------------------------------------------------------------------------------------
def test_oracle_jacobian_is_the_derivative_of_the_residual():
    P = NlProblem(9, 7, 1.3, 0.8)
    rng = np.random.default_rng(3)
    u = 0.9 + 0.4 * rng.random(P.nv)
    x = rng.standard_normal(P.nv)
    eps = 1e-6
    fd = (P.residual(u + eps * x) - P.residual(u - eps * x)) / (2 * eps)
    assert np.abs(P.jacobian(u) @ x - fd).max() <= 1e-8 * np.abs(fd).max()
    # sign(u) where u changes sign: compare away from the kink only (one-sided differences at a kink are not the derivative)
    u = -0.9 - 0.4 * rng.random(P.nv)
    fd = (P.residual(u + eps * x) - P.residual(u - eps * x)) / (2 * eps)
    assert np.abs(P.jacobian(u) @ x - fd).max() <= 1e-8 * np.abs(fd).max()


def test_oracle_converges_at_second_order_to_a_manufactured_solution():
    xs, ys = sy.symbols("x y", real=True)
    ue = 0.9 + sy.sin(sy.pi * xs) * (1 + sy.cos(sy.pi * ys) / 2)          # 0.9 on x = 0, 1; no flux through y = 0, 1; positive
    f = -(sy.diff(ue * sy.diff(ue, xs), xs) + sy.diff(ue * sy.diff(ue, ys), ys)) + ue - 5     # so that -div(|u| grad u) + u = 5 + f
    ue_f, f_f = sy.lambdify((xs, ys), ue, "numpy"), sy.lambdify((xs, ys), f, "numpy")
    errs = []
    for n in (8, 16, 32):
        P = NlProblem(n, n + n // 4)
        xq = P.V.Phi @ P.mesh.coords
        P.load = f_f(xq[:, 0], xq[:, 1])
        u, its, ok = P.solve(atol=1e-11, max_it=20)
        assert ok
        e = u - ue_f(P.mesh.coords[:, 0], P.mesh.coords[:, 1])
        errs.append(float(np.sqrt(e @ (P.V.M @ e))))
    assert errs[0] / errs[1] > 3.5 and errs[1] / errs[2] > 3.7, errs


def test_reference_configuration_newton_history():
    """nl.py's own configuration: 60 x 75, start 0.9, atol 1e-6, at most 10 iterations (nl.py:17-18,48,58-59)."""
    P = NlProblem()
    hist = []
    u, its, ok = P.solve(report=hist)
    assert ok and its == 4 and hist[0] == pytest.approx(0.0604053611, rel=1e-8) and hist[-1] < 1e-6
    assert u.min() == pytest.approx(0.9) and 1.3 < u.max() < 1.35


# ---------------------------------------------------------------------------------------------------------------------
# product operators on the CPU (csrc/pfh_nl.cuh) vs oracle
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nx,ny,Lx,Ly", [(60, 75, 1.0, 1.0), (7, 5, 1.3, 0.7), (1, 1, 1.0, 1.0), (2, 1, 1.0, 2.0), (33, 9, 1.0, 1.0)])
@pytest.mark.parametrize("square", [False, True])
def test_operators_match_the_oracle(nx, ny, Lx, Ly, square):
    P = NlProblem(nx, ny, Lx, Ly)
    rng = np.random.default_rng(nx * 100 + ny)
    u = 0.9 + 0.5 * rng.standard_normal(P.nv)            # changes sign: |u| and sign(u) both exercised
    x = rng.standard_normal(P.nv)
    R, _ = emul.run_nl((nx, ny), (Lx, Ly), PRM, 0, u, square=square)
    Jx, _ = emul.run_nl((nx, ny), (Lx, Ly), PRM, 1, u, x, square=square)
    Ax, d = emul.run_nl((nx, ny), (Lx, Ly), PRM, 2, u, x, square=square)
    A = P.sym_operator(u)
    scale = max(1.0, np.abs(P.jacobian(u) @ x).max())
    assert np.abs(R - P.residual(u)).max() <= 1e-13 * scale
    assert np.abs(Jx - P.jacobian(u) @ x).max() <= 1e-13 * scale
    assert np.abs(Ax - A @ x).max() <= 1e-13 * scale
    assert np.abs(d - A.diagonal()).max() <= 1e-13 * max(1.0, A.diagonal().max())


def test_other_parameters():
    P = NlProblem(6, 4, 2.0, 1.0, src=3.0, kc=0.5, reac=2.0, bc=1.2)
    rng = np.random.default_rng(0)
    u = 1.2 + 0.3 * rng.random(P.nv)
    R, _ = emul.run_nl((6, 4), (2.0, 1.0), [3.0, 0.5, 2.0, 1.2], 0, u)
    assert np.abs(R - P.residual(u)).max() <= 1e-14
    # one-signed triangles take the closed forms of tri_op (no quadrature loop): all positive above, all negative here
    x = rng.standard_normal(P.nv)
    for sq in (False, True):
        Jx, _ = emul.run_nl((6, 4), (2.0, 1.0), [3.0, 0.5, 2.0, 1.2], 1, -u, x, square=sq)
        Ax, d = emul.run_nl((6, 4), (2.0, 1.0), [3.0, 0.5, 2.0, 1.2], 2, -u, x, square=sq)
        assert np.abs(Jx - P.jacobian(-u) @ x).max() <= 1e-13 * np.abs(Jx).max()
        A = P.sym_operator(-u)
        assert np.abs(Ax - A @ x).max() <= 1e-13 * np.abs(Ax).max() and np.abs(d - A.diagonal()).max() <= 1e-13 * d.max()


# ---------------------------------------------------------------------------------------------------------------------
# product solver on the CPU (csrc/pfh_solver.cpp over loop launchers) vs oracle
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def converged():
    return NlProblem().solve(atol=1e-12, max_it=30)[0]


@pytest.mark.parametrize("kw,tol", [(dict(absolute_tolerance=1e-6), 1e-10),                       # nl.py's settings
                                    (dict(absolute_tolerance=1e-11, krylov_rtol=1e-9), 1e-14),
                                    (dict(absolute_tolerance=1e-6, linear_solver=1), 1e-10),        # bicgstab
                                    (dict(absolute_tolerance=1e-6, pc_sweeps_shifted=4), 1e-10),
                                    (dict(absolute_tolerance=1e-6, line_search=2), 1e-10)])
def test_newton_krylov_solve_matches_the_oracle(converged, kw, tol):
    w0 = np.full((1, converged.size), 0.9)
    w, rep = emul.newton_step(_lib.MESH_RIGHT2D, _lib.PROBLEM_NLDIFF, (60, 75), (1.0, 1.0), PRM, 0.1, w0, w0,
                              maximum_iterations=10, **kw)
    assert rep["converged"] == 1 and rep["niters"] == 4          # the oracle's Newton count (exact linear solves)
    assert rep["fnorm0"] == pytest.approx(0.0604053611, rel=1e-8)
    assert np.linalg.norm(w[0] - converged) <= tol * np.linalg.norm(converged)
    assert rep["krylov_its"] <= (200 if kw.get("pc_sweeps_shifted") else 160)


def test_nonconvergence_is_reported_not_raised():
    w0 = np.full((1, 61 * 76), 0.9)
    w, rep = emul.newton_step(_lib.MESH_RIGHT2D, _lib.PROBLEM_NLDIFF, (60, 75), (1.0, 1.0), PRM, 0.1, w0, w0,
                              absolute_tolerance=1e-6, maximum_iterations=2)
    assert rep["converged"] == 0 and rep["reason"] == -5 and rep["niters"] == 2


# ---------------------------------------------------------------------------------------------------------------------
# front end
# ---------------------------------------------------------------------------------------------------------------------
def _nl_form(nx=60, ny=75, value=0.9):
    mesh = df.UnitSquareMesh(nx, ny)
    V = df.FunctionSpace(mesh, "Lagrange", 1)
    u, v, du = df.Function(V), df.TestFunction(V), df.TrialFunction(V)
    u_bc = df.Function(V)
    u_bc.interpolate(df.Constant(value))
    bc = df.DirichletBC(V, u_bc, lambda x: np.logical_or(np.isclose(x[0], 0.0), np.isclose(x[0], 1.0)))
    F = 5.0 * v * df.dx - df.sqrt(u * u) * df.inner(df.grad(u), df.grad(v)) * df.dx - df.inner(u, v) * df.dx
    return mesh, u, du, F, bc


def test_form_is_recognised():
    mesh, u, du, F, bc = _nl_form()
    assert mesh.kind == "right2d" and bc.sides == ("left", "right") and float(bc.value_expr) == 0.9
    assert forms.compile_steady(F, 1, u.symbols, [bc]) == ("nl", PRM)
    name, _, params, dt, prev, model = df.NonlinearVariationalProblem._analyse(F, u, [bc], df.derivative(F, u, du), None)
    assert (name, params, dt, prev, model) == ("nl", PRM, None, None, None)


def test_other_steady_forms_are_refused():
    mesh, u, du, F, bc = _nl_form()
    v = df.TestFunction(u.W)
    with pytest.raises(NotImplementedError):
        forms.compile_steady(F + u * u * v * df.dx, 1, u.symbols, [bc])               # quadratic reaction
    with pytest.raises(NotImplementedError):
        forms.compile_steady(5.0 * v * df.dx - u * df.inner(df.grad(u), df.grad(v)) * df.dx, 1, u.symbols, [bc])   # u, not |u|
    with pytest.raises(NotImplementedError, match="Dirichlet"):
        forms.compile_steady(F, 1, u.symbols, [])
    one_side = df.DirichletBC(u.W, df.Constant(0.9), lambda x, on_boundary: on_boundary and df.near(x[0], 0.0, 1e-12))
    with pytest.raises(NotImplementedError, match="Dirichlet"):
        forms.compile_steady(F, 1, u.symbols, [one_side])
    crossed = df.UnitSquareMesh(4, 4, "crossed")
    assert crossed.kind == "crossed2d"                                                  # nl.py:19, overwritten at :20


def test_right_diagonal_vertex_order():
    import types
    from oracle import mesh as omesh
    ctx = types.SimpleNamespace(dim=2, ncell=(7, 5), length=(1.0, 1.0), mesh_kind=_lib.MESH_RIGHT2D)
    assert np.allclose(engine.Context.vertex_coordinates(ctx), omesh.right2d(7, 5).coords, atol=1e-15)


class _EmulContext:
    """engine.Context for the nl problem with CPU tensors and the product's solver compiled for the CPU (tests/emul):
    lets the driver / the reference's own script run end to end here, numerically."""
    nranks = 1

    def __init__(self, problem, ncell, length, params=None, **kw):
        import torch
        assert problem == "nl"
        self._torch, self.problem, self.nf, self.params = torch, problem, 1, list(params)
        self.ncell, self.length, self.n_owned = tuple(ncell), tuple(length), (ncell[0] + 1) * (ncell[1] + 1)

    def new_vector(self):
        return self._torch.zeros(self.n_owned, dtype=self._torch.float64)

    def newton_solve(self, w, w0, opts=None):
        kw = {k: getattr(opts, k) for k, _ in opts._fields_}
        out, rep = emul.newton_step(_lib.MESH_RIGHT2D, _lib.PROBLEM_NLDIFF, self.ncell, self.length, self.params, 0.1,
                                    w.numpy()[None, :], w.numpy()[None, :], **kw)
        w[:] = self._torch.from_numpy(out[0])
        rep = {k: (int(v) if k != "fnorm0" and k != "fnorm" else float(v)) for k, v in rep.items()}
        return rep


def test_driver_runs_both_solvers(monkeypatch, converged, capsys):
    monkeypatch.setattr(engine, "Context", _EmulContext)
    out = importlib.import_module("pfhubbenchmarks_b200.drivers.nl").main([])
    assert out["snes"][:2] == (4, True) and out["newton"][:2] == (4, True)
    assert np.linalg.norm(out["u"] - converged) <= 1e-10 * np.linalg.norm(converged)
    text = capsys.readouterr().out
    assert "SNES converged = True , niters = 4" in text and "Newton converged = True , niters = 4" in text


@pytest.mark.skipif(not os.path.exists(REF_NL), reason="reference checkout not present")
def test_reference_script_runs_unmodified(monkeypatch, capsys):
    """The reference's own nl.py, text untouched, through `from dolfin import *` -> pfhubbenchmarks_b200.dolfin."""
    from pfhubbenchmarks_b200 import run_script
    monkeypatch.setattr(engine, "Context", _EmulContext)
    saved = {k: sys.modules.get(k) for k in ("dolfin", "pfbase")}
    try:
        ns = run_script.run(REF_NL)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    text = capsys.readouterr().out
    assert "SNES converged = True , niters = 4" in text and "Newton converged = True , niters = 4" in text
    assert ns["niters"] == 4 and ns["converged"] is True
