"""sympy -> functor code generation (pfhubbenchmarks_b200/codegen.py), CPU side: a model that is none of the shipped
systems (Cahn-Hilliard with a Flory-Huggins free energy coupled to an Allen-Cahn field) is written with the pfbase
builders, lowered, emitted as `struct ProbUser`, compiled for the CPU by tests/emul together with the product's
per-vertex operator code, and compared with the oracle's assembled operators for the same expressions."""
import os

import numpy as np
import pytest
import sympy as sy

import emul
import pfhubbenchmarks_b200.dolfin as df
from helpers import rel_max
from oracle import mesh as omesh, problems, stepper
from pfhubbenchmarks_b200 import codegen
from pfhubbenchmarks_b200.drivers import custom_model as cm
from pfhubbenchmarks_b200.drivers.common import mixed_p1
from pfhubbenchmarks_b200.pfbase import allen_cahn_weak_form, cahn_hilliard_weak_form, dx, grad, inner

HERE = os.path.dirname(os.path.abspath(__file__))
PARAMS = dict(M=cm.MOBILITY, kappa=cm.KAPPA, L=cm.L_ETA, ke=cm.KAPPA_ETA, chi=cm.CHI)
free_energy = cm.free_energy


def build_model(mesh):
    """The weak form of drivers/custom_model.py, statement for statement, stopping before the problem constructor."""
    W, w, w0, w_, dw = mixed_p1(mesh, 3)
    dt = df.Constant(0.05)
    M, kappa, L, ke = (df.Constant(PARAMS[k]) for k in ("M", "kappa", "L", "ke"))
    c, mu, eta = df.split(w)
    f = free_energy(c, eta)
    F = cahn_hilliard_weak_form(c, mu, w_[0], w_[1], w0[0], dt, M, kappa, df.diff(f, c)) \
        + allen_cahn_weak_form(eta, w_[2], w0[2], dt, L, ke, df.diff(f, eta), df.Constant(0))
    E = f * dx + kappa / 2.0 * inner(grad(c), grad(c)) * dx + ke / 2.0 * inner(grad(eta), grad(eta)) * dx
    return dict(W=W, w=w, w0=w0, dw=dw, dt=dt, F=F, E=E, S=c * dx, f=f, c=c, eta=eta, kappa=kappa, ke=ke)


class OracleModel(problems.Problem):
    """The same model in the oracle's terms (assembled operators, lambdified sympy)."""
    nf = 3
    fields = ("c", "mu", "eta")

    def __init__(self):
        c, mu, eta = sy.symbols("c mu eta")
        self.syms = (c, mu, eta)
        f = free_energy(c, eta, sy.log)
        self.g_exprs = [sy.Integer(0), -sy.diff(f, c), PARAMS["L"] * sy.diff(f, eta)]
        self.e_expr, self.m_expr = f, c
        self.grad_e = [(0, PARAMS["kappa"] / 2.0), (2, PARAMS["ke"] / 2.0)]
        self._compile()

    def lin(self, dt):
        Mc = np.zeros((3, 3)); Kc = np.zeros((3, 3)); Mp = np.zeros(3)
        Mc[0, 0] = 1.0 / dt; Mp[0] = 1.0 / dt; Kc[0, 1] = PARAMS["M"]
        Mc[1, 1] = 1.0; Kc[1, 0] = -PARAMS["kappa"]
        Mc[2, 2] = 1.0 / dt; Mp[2] = 1.0 / dt; Kc[2, 2] = PARAMS["L"] * PARAMS["ke"]
        return Mc, Kc, Mp


@pytest.fixture(scope="module")
def model():
    mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(14.0, 10.0), 7, 5, "crossed")
    parts = build_model(mesh)
    L = codegen.lower(parts["F"], 3, parts["w"].symbols, parts["w0"].symbols)
    m = codegen.UserModel(L, parts["E"], parts["S"], ("c", "mu", "eta"))
    hdr = os.path.join(HERE, "emul", "user_%s.cuh" % m.tag)
    if not os.path.exists(hdr) or open(hdr).read() != m.source:
        open(hdr, "w").write(m.source)
    return m, hdr


def test_lowering_splits_linear_and_nonlinear_parts(model):
    m, _ = model
    assert m.inputs == [0, 2] and m.outputs == [1, 2]
    dt = m.dt_sym
    assert sy.simplify(m.Mc[0][0] - 1 / dt) == 0 and sy.simplify(m.Mp[0] - 1 / dt) == 0 and float(m.Kc[0][1]) == 2.0
    # the -2 chi c term of f' is linear with a constant coefficient: it lands in the mass stencil, not in g
    assert float(m.Mc[1][0]) == pytest.approx(2 * PARAMS["chi"]) and float(m.Mc[1][1]) == 1.0 and float(m.Kc[1][0]) == -1.5
    assert sy.simplify(m.Mc[2][2] - (PARAMS["L"] * 1.0 + 1 / dt)) == 0 and float(m.Kc[2][2]) == pytest.approx(PARAMS["L"] * PARAMS["ke"])
    assert m.layout == dict(has_ch=True, Mob=2.0, kappa=1.5, spd=[2], jacobi_only=False)
    assert m.grad == {0: 0.75, 2: 0.35}
    assert "struct ProbUser" in m.source and "log(" in m.source and "pow(" not in m.source


@pytest.mark.parametrize("mk,ncell,length", [(2, (7, 5), (14.0, 10.0)), (3, (4, 3, 3), (8.0, 6.0, 6.0))])
def test_generated_functor_matches_oracle(model, mk, ncell, length):
    m, hdr = model
    rng = np.random.default_rng(2)
    om = omesh.crossed2d(*ncell, *length) if mk == 2 else omesh.kuhn3d(*ncell, *length)
    D = stepper.Discretization(om, OracleModel())
    U = np.stack([0.5 + 0.2 * rng.random(om.nv), rng.standard_normal(om.nv), 0.5 + 0.3 * rng.standard_normal(om.nv)])
    U0 = np.stack([0.5 + 0.2 * rng.random(om.nv), rng.standard_normal(om.nv), 0.5 + 0.3 * rng.standard_normal(om.nv)])
    X = rng.standard_normal((3, om.nv))
    dt = 0.05
    kw = dict(nf=3, user_header=hdr)
    R = emul.run(mk, 100, ncell, length, [], dt, 0, U=U, U0=U0, **kw)
    assert rel_max(R, D.residual(U, U0, dt)) < 1e-12
    JX = emul.run(mk, 100, ncell, length, [], dt, 1, U=U, X=X, **kw)
    assert rel_max(JX, (D.jacobian(U, dt) @ X.ravel()).reshape(3, -1)) < 1e-12
    E, S = emul.run(mk, 100, ncell, length, [], dt, 2, U=U, **kw)
    assert abs(E / D.energy(U) - 1) < 1e-12 and abs(S / D.mass(U) - 1) < 1e-13


def test_unsupported_forms_are_refused():
    mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(4.0, 4.0), 2, 2, "crossed")
    pe = df.FunctionSpace(mesh, "P", 1).ufl_element()
    W = df.FunctionSpace(mesh, df.MixedElement([pe] * 2))
    w, w0, w_ = df.Function(W), df.Function(W), df.TestFunction(W)
    dt, one = df.Constant(0.1), df.Constant(1.0)
    c, mu = df.split(w)
    # a mobility that depends on the field is not a constant-coefficient stiffness term
    F = cahn_hilliard_weak_form(c, mu, w_[0], w_[1], w0[0], dt, c * (1 - c), one, df.diff(c ** 6, c))
    with pytest.raises(NotImplementedError, match="diffusion"):
        codegen.UserModel(codegen.lower(F, 2, w.symbols, w0.symbols))
    # a load that depends on the previous step (semi-implicit splitting) does not fit the discrete form either
    F = cahn_hilliard_weak_form(c, mu, w_[0], w_[1], w0[0], dt, one, one, df.diff(c ** 6, c) - w0[0] ** 3)
    with pytest.raises(NotImplementedError):
        codegen.UserModel(codegen.lower(F, 2, w.symbols, w0.symbols))


class _Generic(problems.Problem):
    """An oracle problem from explicit pieces (syms, g, e, m, grad_e, lin matrices)."""

    def __init__(self, syms, g_exprs, e_expr, m_expr, grad_e, lin):
        self.nf, self.syms, self.g_exprs, self.e_expr, self.m_expr, self.grad_e, self._lin = len(syms), syms, g_exprs, e_expr, m_expr, grad_e, lin
        self.fields = tuple(str(s) for s in syms)
        self._compile()

    def lin(self, dt):
        return self._lin(dt)


def _emit(parts_F, w, w0, E, S, names):
    m = codegen.UserModel(codegen.lower(parts_F, len(names), w.symbols, w0.symbols), E, S, names)
    hdr = os.path.join(HERE, "emul", "user_%s.cuh" % m.tag)
    if not os.path.exists(hdr) or open(hdr).read() != m.source:
        open(hdr, "w").write(m.source)
    return m, hdr


def _compare(m, hdr, D, nf, mk, ncell, length, U, U0, X, dt):
    kw = dict(nf=nf, user_header=hdr)
    assert rel_max(emul.run(mk, 100, ncell, length, [], dt, 0, U=U, U0=U0, **kw), D.residual(U, U0, dt)) < 1e-12
    assert rel_max(emul.run(mk, 100, ncell, length, [], dt, 1, U=U, X=X, **kw), (D.jacobian(U, dt) @ X.ravel()).reshape(nf, -1)) < 1e-12
    E, S = emul.run(mk, 100, ncell, length, [], dt, 2, U=U, **kw)
    assert abs(E / D.energy(U) - 1) < 1e-12 and abs(S / D.mass(U) - 1) < 1e-12


def test_two_coupled_parabolic_fields():
    """A thermal phase-field pair with a different coupling polynomial than bench3 and an exponential source:
    off-diagonal stiffness (the latent-heat term), both fields nonlinear inputs, SPD/SPD preconditioner layout."""
    from pfhubbenchmarks_b200.pfbase import allen_cahn_RHS_IBP, diffusion_weak_form
    ncell, length = (6, 5), (12.0, 10.0)
    mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(*length), *ncell, "crossed")
    W, w, w0, w_, dw = mixed_p1(mesh, 2)
    dt, D_, tau, W2 = df.Constant(0.02), df.Constant(4.0), df.Constant(0.8), df.Constant(1.3)
    U, p = df.split(w)
    dfdp = (p - 0.7 * U * (1 - p ** 2) * df.exp(-U ** 2)) * (1 - p ** 2) * (1 + 0.1 * p)
    zero = df.Constant(0.0)
    F = diffusion_weak_form(U, w_[0], w0[0], dt, df.Constant(1.0), D_, df.Constant((0.0, 0.0)), zero) \
        - 0.5 * allen_cahn_RHS_IBP(p, w_[0], 1 / tau, W2, -dfdp, zero) * dx \
        + allen_cahn_weak_form(p, w_[1], w0[1], dt, 1 / tau, W2, -dfdp, zero)
    E = (0.25 * p ** 4 - 0.5 * p ** 2 + 0.5 * W2 * inner(grad(p), grad(p))) * dx
    m, hdr = _emit(F, w, w0, E, 0.5 * (p + 1.0) * dx, ("U", "phi"))
    assert m.inputs == [0, 1] and m.outputs == [0, 1] and m.layout["spd"] == [0, 1] and not m.layout["has_ch"]
    sU, sp = sy.symbols("U phi")
    g = (sp - 0.7 * sU * (1 - sp ** 2) * sy.exp(-sU ** 2)) * (1 - sp ** 2) * (1 + 0.1 * sp)

    def lin(h):
        return (np.array([[1 / h, 0.0], [0.0, 1 / h]]), np.array([[4.0, 0.5 * 1.3 / 0.8], [0.0, 1.3 / 0.8]]), np.array([1 / h, 1 / h]))

    om = omesh.crossed2d(*ncell, *length)
    D = stepper.Discretization(om, _Generic((sU, sp), [-(0.5 / 0.8) * g, -(1 / 0.8) * g], 0.25 * sp ** 4 - 0.5 * sp ** 2,
                                            0.5 * (sp + 1.0), [(1, 0.65)], lin))
    rng = np.random.default_rng(4)
    U_, U0_, X_ = (rng.standard_normal((2, om.nv)) * 0.6 for _ in range(3))
    _compare(m, hdr, D, 2, 2, ncell, length, U_, U0_, X_, 0.02)


def test_single_allen_cahn_field():
    """One field only (nf = 1): the smallest model the generator and the operator code accept."""
    ncell, length = (5, 4), (5.0, 4.0)
    mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(*length), *ncell, "crossed")
    W, w, w0, w_, dw = mixed_p1(mesh, 1)
    dt, L, k = df.Constant(0.1), df.Constant(2.0), df.Constant(0.5)
    e = df.split(w)[0]
    f = e ** 2 * (1 - e) ** 2 + 0.3 * df.sin(e)
    F = allen_cahn_weak_form(e, w_[0], w0[0], dt, L, k, df.diff(f, e), df.Constant(0))
    m, hdr = _emit(F, w, w0, f * dx + k / 2.0 * inner(grad(e), grad(e)) * dx, e * dx, ("eta",))
    assert m.nf == 1 and m.inputs == [0] and m.outputs == [0] and m.layout["spd"] == [0]
    s = sy.Symbol("eta")
    fs = s ** 2 * (1 - s) ** 2 + 0.3 * sy.sin(s)
    om = omesh.crossed2d(*ncell, *length)
    D = stepper.Discretization(om, _Generic((s,), [2.0 * sy.diff(fs, s)], fs, s, [(0, 0.25)],
                                            lambda h: (np.array([[1 / h]]), np.array([[1.0]]), np.array([1 / h]))))
    rng = np.random.default_rng(6)
    U_, U0_, X_ = (rng.standard_normal((1, om.nv)) * 0.5 for _ in range(3))
    _compare(m, hdr, D, 1, 2, ncell, length, U_, U0_, X_, 0.1)


@pytest.mark.parametrize("num_eta", [1, 2])
def test_reduced_ostwald_models(num_eta):
    """dolfin/bench2_eta1.py / bench2_eta2.py (benchmark 2's free energy with one / two order parameters) are not
    shipped systems: drivers/bench2_eta.py runs them as generated models.  The emitted functor against the oracle."""
    import itertools
    from pfhubbenchmarks_b200 import dolfin as shim
    from pfhubbenchmarks_b200.drivers import bench2_eta
    ncell, length = (6, 5), (12.0, 10.0)
    mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(*length), *ncell, "crossed")
    got = {}
    orig = shim.compile_only

    def capture(F, w, bcs=None, J=None, functionals=None):
        got["model"] = shim._analyse(F, w, bcs, J, functionals)[5]
        return None

    shim.compile_only = capture
    try:
        bench2_eta.setup(mesh, num_eta, compile_only=True)
    finally:
        shim.compile_only = orig
    m = got["model"]
    assert m is not None and m.nf == 2 + num_eta and m.layout["has_ch"] and m.layout["spd"] == list(range(2, 2 + num_eta))
    hdr = os.path.join(HERE, "emul", "user_%s.cuh" % m.tag)
    if not os.path.exists(hdr) or open(hdr).read() != m.source:
        open(hdr, "w").write(m.source)
    syms = sy.symbols("c mu eta1 eta2")[:2 + num_eta]
    c, es = syms[0], syms[2:]
    H = sum(e ** 3 * (6 * e ** 2 - 15 * e + 10) for e in es)
    Wd = sum(e ** 2 * (1 - e) ** 2 for e in es) + 5.0 * sum(a ** 2 * b ** 2 for a, b in itertools.combinations(es, 2))
    f = 2.0 * (c - 0.3) ** 2 * (1 - H) + 2.0 * (c - 0.7) ** 2 * H + Wd
    nf = 2 + num_eta

    def lin(h):
        Mc = np.zeros((nf, nf)); Kc = np.zeros((nf, nf)); Mp = np.zeros(nf)
        Mc[0, 0] = 1 / h; Mp[0] = 1 / h; Kc[0, 1] = 5.0; Mc[1, 1] = 1.0; Kc[1, 0] = -3.0
        for i in range(2, nf):
            Mc[i, i] = 1 / h; Mp[i] = 1 / h; Kc[i, i] = 15.0
        return Mc, Kc, Mp

    om = omesh.crossed2d(*ncell, *length)
    D = stepper.Discretization(om, _Generic(syms, [sy.Integer(0), -sy.diff(f, c)] + [5.0 * sy.diff(f, e) for e in es], f, c,
                                            [(0, 1.5)] + [(2 + i, 1.5) for i in range(num_eta)], lin))
    rng = np.random.default_rng(8)
    U_, U0_, X_ = (0.5 + 0.3 * rng.standard_normal((nf, om.nv)) for _ in range(3))
    _compare(m, hdr, D, nf, 2, ncell, length, U_, U0_, X_, 0.01)


def test_generated_model_newton_step_on_the_cpu(model):
    """The whole Newton-Krylov step of the generated Flory-Huggins model through the product's host solver (CPU leg,
    tests/emul/emul_solver.cpp, with the model's own preconditioner layout from the emitted pc_layout) against the
    oracle's Newton/LU step -- the CPU counterpart of tests/test_gpu_codegen.py."""
    m, hdr = model
    ncell, length = (10, 8), (20.0, 16.0)
    om = omesh.crossed2d(*ncell, *length)
    D = stepper.Discretization(om, OracleModel())
    ic = cm.InitialConditions(degree=2)
    u0 = np.zeros((om.nv, 3))
    for i in range(om.nv):
        ic.eval(u0[i], om.coords[i])
    u0 = np.ascontiguousarray(u0.T)
    w, rep = emul.newton_step(2, 100, ncell, length, [], 0.05, u0, u0, basis_precision=1, user_header=hdr, absolute_tolerance=1e-10,
                              krylov_rtol=1e-10, maximum_iterations=30, krylov_maximum_iterations=5000)
    assert rep["converged"] == 1, rep
    u, n, ok = D.solve_step(u0, u0, 0.05, atol=1e-11)
    assert ok and max(np.linalg.norm(w[f] - u[f]) / np.linalg.norm(u[f]) for f in (0, 2)) < 1e-8
    w2, rep2 = emul.newton_step(2, 100, ncell, length, [], 0.05, u0, u0, user_header=hdr)        # default tolerances, fp32 basis
    assert rep2["converged"] == 1 and np.linalg.norm(w2[0] - u[0]) / np.linalg.norm(u[0]) < 1e-6
