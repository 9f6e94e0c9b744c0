"""A small UFL-like layer: just enough symbolic machinery to write the reference's weak forms and free energies the way
the reference drivers do, and to map them onto the compiled CUDA operator families.

The reference builds its problems with UFL (`dolfin/bench1.py:39-69`): `split(w)`, `variable`, `diff`, the weak-form
builders of `dolfin/pfbase.py:361-460` (`inner(grad(c_), M*grad(mu))*dx` ...), `derivative(F, w, dw)`, and energy
functionals such as `f_chem*dx + kappa/2*inner(grad(c), grad(c))*dx` (`bench1.py:124-125`).  Here the same expressions
are sympy expressions; a `Form` is a sum of three kinds of terms

    ("src",   test a, expr)          int v_a expr(u, u0, dt) dx        (pointwise, quadrature)
    ("stiff", test a, field b, k)    int k grad v_a . grad u_b dx
    ("func",  expr) / ("gradfunc", a, b, k)   functionals without a test function (energies, solute)

`compile_form` then recognises which of the compiled operator families (`PFH_PROBLEM_*`, csrc/pfh_problems.cuh) the
form is, extracts its coefficients, and checks symbolically that the driver's free energy IS the one the kernels
implement.  Anything else raises NotImplementedError (a sympy -> CUDA functor generator is the planned widening).
"""
import itertools

import sympy as sy

_counter = itertools.count()


# ------------------------------------------------------------------------------------------------------------------
# symbols
# ------------------------------------------------------------------------------------------------------------------
class FieldSymbol(sy.Symbol):
    """Component `index` of a mixed Function (`w[0]`, `split(w)[1]`)."""
    def __new__(cls, name, owner=None, index=0):
        obj = sy.Symbol.__new__(cls, name, real=True)
        obj.owner, obj.index = owner, index
        return obj


class ConstSymbol(sy.Symbol):
    """The symbol behind a `Constant`: its value may change between solves (dt)."""
    def __new__(cls, name, const=None):
        obj = sy.Symbol.__new__(cls, name, real=True)
        obj.const = const
        return obj


class Test:
    """Component of a TestFunction, possibly scaled (`L * u_`)."""
    def __init__(self, index, coef=1):
        self.index, self.coef = index, sy.sympify(coef)

    def __mul__(self, other):
        if isinstance(other, (Grad, Measure)):
            return NotImplemented
        if isinstance(other, Form):
            return NotImplemented
        return Tested(self.index, self.coef * _expr(other))

    __rmul__ = __mul__


class Tested:
    """v_a * expr, waiting for `* dx`."""
    def __init__(self, index, expr):
        self.index, self.expr = index, sy.sympify(expr)

    def __mul__(self, other):
        if isinstance(other, Measure):
            return Form([("src", self.index, self.expr)])
        return Tested(self.index, self.expr * _expr(other))

    __rmul__ = __mul__

    def __neg__(self):
        return Tested(self.index, -self.expr)

    def __add__(self, other):
        if isinstance(other, Tested) and other.index == self.index:
            return Tested(self.index, self.expr + other.expr)
        if isinstance(other, GradPair):
            return Mixed([self, other])
        return NotImplemented

    def __truediv__(self, other):
        return Tested(self.index, self.expr / _expr(other))


class Grad:
    """grad of a field symbol or of a (scaled) test function, times a scalar coefficient."""
    def __init__(self, of, coef=1):
        self.of, self.coef = of, sy.sympify(coef)

    def __mul__(self, other):
        return Grad(self.of, self.coef * _expr(other))

    __rmul__ = __mul__

    def __getitem__(self, i):          # g = grad(phi); g[1] / g[0]  (bench3.py:80-81, unused there)
        return sy.Symbol("d%d_%s" % (i, self.of), real=True)

    def __add__(self, other):          # D*grad(u) + f1 with f1 = Constant((0, 0))  (pfbase.py:442)
        if isinstance(other, VectorZero):
            return self
        return NotImplemented

    __radd__ = __add__


class VectorZero:
    pass


class GradPair:
    """inner(grad a, grad b) * coef."""
    def __init__(self, a, b, coef):
        self.a, self.b, self.coef = a, b, sy.sympify(coef)

    def __mul__(self, other):
        if isinstance(other, Measure):
            if isinstance(self.a, Test) or isinstance(self.b, Test):
                t, f = (self.a, self.b) if isinstance(self.a, Test) else (self.b, self.a)
                if isinstance(f, Test):
                    raise NotImplementedError("bilinear forms in two test functions")
                return Form([("stiff", t.index, f, sy.expand(self.coef * t.coef))])
            return Form([("gradfunc", self.a, self.b, self.coef)])
        return GradPair(self.a, self.b, self.coef * _expr(other))

    __rmul__ = __mul__

    def __neg__(self):
        return GradPair(self.a, self.b, -self.coef)

    def __truediv__(self, other):
        return GradPair(self.a, self.b, self.coef / _expr(other))

    def __add__(self, other):
        if isinstance(other, (Tested, GradPair)):
            return Mixed([self, other])
        if isinstance(other, sy.Expr):
            return Mixed([self, other])
        return NotImplemented

    __radd__ = __add__


class Mixed:
    """A sum of integrands of different kinds, e.g. `eta_*dfdeta + kappa*inner(grad(eta_), grad(eta))` (pfbase.py:392)."""
    def __init__(self, parts):
        self.parts = list(parts)

    def __mul__(self, other):
        if isinstance(other, Measure):
            out = Form([])
            for p in self.parts:
                out = out + (p * other if not isinstance(p, sy.Expr) else Form([("func", p)]))
            return out
        return Mixed([p * other for p in self.parts])

    __rmul__ = __mul__

    def __neg__(self):
        return Mixed([-p for p in self.parts])

    def __add__(self, other):
        if isinstance(other, Mixed):
            return Mixed(self.parts + other.parts)
        return Mixed(self.parts + [other])

    __radd__ = __add__

    def __iadd__(self, other):
        return self.__add__(other)


class Measure:
    def __rmul__(self, other):
        if isinstance(other, (Tested, GradPair, Mixed)):
            return other * self
        return Form([("func", _expr(other))])


dx = Measure()


class Form:
    def __init__(self, terms):
        self.terms = list(terms)

    def __add__(self, other):
        if other == 0:
            return self
        return Form(self.terms + other.terms)

    __radd__ = __add__

    def __neg__(self):
        return self * -1

    def __sub__(self, other):
        return self + (-other)

    def __mul__(self, s):
        s = _expr(s)
        out = []
        for t in self.terms:
            if t[0] == "src":
                out.append(("src", t[1], t[2] * s))
            elif t[0] == "stiff":
                out.append(("stiff", t[1], t[2], t[3] * s))
            elif t[0] == "func":
                out.append(("func", t[1] * s))
            else:
                out.append(("gradfunc", t[1], t[2], t[3] * s))
        return Form(out)

    __rmul__ = __mul__


def _expr(x):
    if hasattr(x, "sym"):          # dolfin.Constant, dolfin.SubFunction (a component from `c, _ = w.split()`)
        return x.sym
    return sy.sympify(x)


def grad(u):
    if hasattr(u, "sym") and isinstance(u.sym, FieldSymbol):   # grad(c) with c from `c, _ = w.split()` (bench1.py:188-193)
        u = u.sym
    if isinstance(u, Test):
        return Grad(u)
    if isinstance(u, Tested):                       # grad(L * u_) with constant L  (pfbase.py:442)
        return Grad(Test(u.index, u.expr))
    return Grad(u)


def inner(a, b):
    if isinstance(a, Grad) and isinstance(b, Grad):
        return GradPair(a.of, b.of, a.coef * b.coef)
    if isinstance(a, Grad) or isinstance(b, Grad):
        raise NotImplementedError("inner() of a gradient and a scalar")
    return a * b          # inner(u, v) of two scalars (nl.py:37)


def variable(x):
    return x


def diff(f, x):
    return sy.diff(_expr(f), x)


# ------------------------------------------------------------------------------------------------------------------
# recognising the compiled operator families
# ------------------------------------------------------------------------------------------------------------------
class Lowered:
    """A weak form in the common discrete form of csrc/pfh_problems.cuh: per test row the pointwise source, split into
    the parts linear in the current / previous fields (mass terms) and the rest, plus the stiffness couplings."""
    def __init__(self, form, nfields, fields, prev, dt):
        self.nf, self.fields, self.prev, self.dt = nfields, fields, prev, dt
        self.src = [sy.Integer(0)] * nfields
        self.stiff = [[sy.Integer(0)] * nfields for _ in range(nfields)]
        for t in form.terms:
            if t[0] == "src":
                self.src[t[1]] += t[2]
            elif t[0] == "stiff":
                if not isinstance(t[2], FieldSymbol):
                    raise NotImplementedError("gradient of a non-field expression")
                self.stiff[t[1]][t[2].index] += t[3]
            else:
                raise NotImplementedError("functional term inside a residual form")
        self.src = [sy.expand(s) for s in self.src]

    def values(self, expr):
        """Substitute the current value of every Constant except dt."""
        expr = sy.sympify(expr)
        subs = {s: s.const.value for s in expr.free_symbols if isinstance(s, ConstSymbol) and s is not self.dt}
        return sy.expand(expr.subs(subs)) if subs else expr


def _num(x):
    return float(sy.N(x))


def _snap(x):
    """Coefficients come back out of polynomial algebra with a few ulp of round-off (0.2999999999999997 for the
    driver's 0.3): return the 13-digit decimal when x is within 1e-14 of it, else x unchanged."""
    if x is None:
        return None
    r = float("%.13g" % x)
    return r if abs(r - x) <= 1e-14 * abs(x) else x


def _is_zero(e):
    e = sy.expand(e)
    if e.is_number:
        return abs(_num(e)) < 1e-14          # sympy does not equate Float(0.0) or Float(-1.0) with the Integers
    return sy.simplify(e) == 0 or _close(e, 1e-14)


def _close(e, tol=1e-12):
    """Numerically-zero test of a polynomial expression (coefficients carry float round-off)."""
    e = sy.expand(e)
    if e == 0:
        return True
    try:
        coeffs = sy.Poly(e, *sorted(e.free_symbols, key=str)).coeffs() if e.free_symbols else [e]
        return all(abs(_num(c)) < tol for c in coeffs)
    except (sy.PolynomialError, TypeError):
        return False


def lower(F, nfields, fields, prev):
    """Lowered form of a residual; dt is the Constant in the coefficient of the previous-step fields, (u - u0)/dt."""
    dt = None
    for t in F.terms:
        if t[0] != "src":
            continue
        for p in prev:
            for s in sy.diff(sy.sympify(t[2]), p).free_symbols:
                if isinstance(s, ConstSymbol):
                    dt = s
    if dt is None:
        raise NotImplementedError("no time-derivative term (u - u0)/dt found")
    return Lowered(F, nfields, fields, prev, dt)


def compile_form(F, nfields, fields, prev, bcs=()):
    """Map a residual form onto (benchmark name, params) of engine.PROBLEMS; NotImplementedError if it is none of them
    (the caller then generates a functor for it, codegen.py)."""
    L = lower(F, nfields, fields, prev)
    dt = L.dt
    for matcher in (_match_ch, _match_bench6, _match_bench3, _match_bench2):
        try:
            out = matcher(L, bcs)
        except (sy.PolynomialError, TypeError, ValueError, ZeroDivisionError):
            out = None          # not a polynomial of the expected shape: not this family
        if out is not None:
            return out[0], [_snap(v) for v in out[1]], dt.const
    raise NotImplementedError(
        "this weak form is not one of the compiled operator families (bench1/b13d, bench2, bench3, bench6)")


def compile_steady(F, nfields, fields, bcs=()):
    """A residual form without a time derivative.  The one steady system the reference holds is dolfin/nl.py:37,
        F = int src v - int kc sqrt(u u) grad u . grad v - int reac u v,   u = g on x = 0 and x = Lx (nl.py:31-35);
    returns ("nl", [src, kc, reac, g]) for it (PFH_PROBLEM_NLDIFF on the right-diagonal mesh), NotImplementedError for
    anything else."""
    no = NotImplementedError("steady form that is not dolfin/nl.py's  src*v - kc*sqrt(u*u)*inner(grad(u), grad(v)) - reac*u*v")
    if nfields != 1:
        raise no
    u = fields[0]
    src, stiff = sy.Integer(0), sy.Integer(0)
    for t in F.terms:
        if t[0] == "src" and t[1] == 0:
            src += t[2]
        elif t[0] == "stiff" and t[1] == 0 and t[2] is u:
            stiff += t[3]
        else:
            raise no

    def values(e):
        e = sy.sympify(e)
        return e.subs({s: s.const.value for s in e.free_symbols if isinstance(s, ConstSymbol)})

    src, stiff = sy.expand(values(src)), values(stiff)
    try:
        p = sy.Poly(src, u)
    except sy.PolynomialError:
        raise no
    if p.degree() > 1 or (src.free_symbols - {u}):
        raise no
    reac, s0 = -_num(p.coeff_monomial(u)), _num(p.coeff_monomial(1))
    kc = sy.simplify(-stiff / sy.Abs(u))
    if kc.free_symbols or _num(kc) <= 0.0 or reac < 0.0:
        raise no
    sides = sorted(s for bc in bcs for s in bc.sides)
    vals = {float(bc.value_expr) for bc in bcs if not sy.sympify(bc.value_expr).free_symbols}
    if sides != ["left", "right"] or len(vals) != 1 or any(bc.field != 0 for bc in bcs):
        raise NotImplementedError("nl.py-type problem needs one constant Dirichlet value on x = 0 and x = Lx (nl.py:31-35)")
    return "nl", [_snap(s0), _snap(_num(kc)), _snap(reac), _snap(vals.pop())]


def _time_row(L, a):
    """True if the pointwise part of row a is (u_a - u0_a)/dt plus terms free of dt and of the previous step."""
    rest = sy.expand(L.src[a] - (L.fields[a] - L.prev[a]) / L.dt)
    return not rest.has(L.dt) and not any(rest.has(p) for p in L.prev)


def _quartic(fprime, c):
    """(rho, ca, cb) with fprime = d/dc rho (c-ca)^2 (cb-c)^2, else None."""
    p = sy.Poly(sy.expand(fprime), c)
    if p.degree() != 3:
        return None
    rho = _num(p.LC()) / 4.0
    roots = sorted(_num(sy.re(r)) for r in sy.nroots(p, n=30))
    ca, cb = roots[0], roots[2]
    tmpl = sy.diff(rho * (c - ca) ** 2 * (cb - c) ** 2, c)
    return (rho, ca, cb) if _close(fprime - tmpl, 1e-9) else None


def _ch_pair(L, ic, imu):
    """Cahn-Hilliard pair (pfbase.py:361-383): returns (Mob, kappa, dfdc expr) or None."""
    c, mu = L.fields[ic], L.fields[imu]
    if not _time_row(L, ic):
        return None
    if not _is_zero(L.src[ic] - (c - L.prev[ic]) / L.dt):
        return None
    Mob = L.values(L.stiff[ic][imu])
    if any(not _is_zero(L.stiff[ic][b]) for b in range(L.nf) if b != imu) or _is_zero(Mob):
        return None
    kappa = -L.values(L.stiff[imu][ic])
    if any(not _is_zero(L.stiff[imu][b]) for b in range(L.nf) if b != ic):
        return None
    smu = L.values(L.src[imu])
    if not _is_zero(sy.diff(smu, mu) - 1):
        return None
    dfdc = sy.expand(mu - smu)
    return _num(Mob), _num(kappa), dfdc


def _match_ch(L, bcs):
    if L.nf != 2 or bcs:
        return None
    got = _ch_pair(L, 0, 1)
    if got is None:
        return None
    Mob, kappa, dfdc = got
    if dfdc.free_symbols - {L.fields[0]}:
        return None
    q = _quartic(dfdc, L.fields[0])
    if q is None:
        return None
    rho, ca, cb = q
    return "bench1", [ca, cb, rho, kappa, Mob]


def _match_bench6(L, bcs):
    if L.nf != 3:
        return None
    got = _ch_pair(L, 0, 1)
    if got is None:
        return None
    Mob, kappa, dfdc = got
    c, phi = L.fields[0], L.fields[2]
    k = _num(sy.diff(dfdc, phi))
    q = _quartic(sy.expand(dfdc - k * phi), c)
    if q is None:
        return None
    rho, ca, cb = q
    # Poisson row (pfbase.py:410-421): -int grad v . grad phi - int v f, f = -k c / eps
    if not _is_zero(L.values(L.stiff[2][2]) + 1) or any(not _is_zero(L.stiff[2][b]) for b in (0, 1)):
        return None
    sphi = L.values(L.src[2])
    if sphi.free_symbols - {c}:
        return None
    k_over_eps = _num(sy.diff(sphi, c))
    eps = k / k_over_eps
    # Dirichlet data: phi = 0 on x = 0, sin(y/7) on x = Lx  (bench6.py:77-90)
    sides = sorted((bc.side, str(bc.value_expr)) for bc in bcs)
    if len(sides) != 2 or any(bc.field != 2 for bc in bcs) or sides[0] != ("left", "0") or sides[1][0] != "right":
        raise NotImplementedError("bench6-type form with boundary conditions other than the reference's")
    xs, ys = sy.symbols("x0 x1", real=True)
    if not _is_zero(sy.sympify(bcs[0].value_expr if bcs[0].side == "right" else bcs[1].value_expr) - sy.sin(ys / 7)):
        raise NotImplementedError("Dirichlet value on x = Lx must be sin(x[1]/7) (bench6.py:84)")
    return "bench6", [ca, cb, rho, kappa, Mob, k, eps]


def _match_bench3(L, bcs):
    if L.nf != 2 or bcs:
        return None
    U, p = L.fields
    if not (_time_row(L, 0) and _time_row(L, 1)):
        return None
    D = _num(L.values(L.stiff[0][0]))
    sp = L.values(sy.expand(L.src[1] - (p - L.prev[1]) / L.dt))       # -(1/tau) dfdp
    inv_tau = -_num(sy.diff(sp, p).subs({p: 0, U: 0}))
    if inv_tau <= 0:
        return None
    tau = 1.0 / inv_tau
    W2 = _num(L.values(L.stiff[1][1])) * tau
    lam = _num(sy.diff(sp, U).subs({p: 0, U: 0})) * tau
    dfdp = (p - lam * U * (1 - p ** 2)) * (1 - p ** 2)
    if not _close(sp + dfdp / tau, 1e-9):
        return None
    sU = L.values(sy.expand(L.src[0] - (U - L.prev[0]) / L.dt))
    if not _close(sU + 0.5 * dfdp / tau, 1e-9):
        return None
    if abs(_num(L.values(L.stiff[0][1])) - 0.5 * W2 / tau) > 1e-12 or not _is_zero(L.stiff[1][0]):
        return None
    return "bench3", [W2 ** 0.5, tau, D, lam, None]


def _match_bench2(L, bcs):
    if L.nf != 6 or bcs:
        return None
    got = _ch_pair(L, 0, 1)
    if got is None:
        return None
    Mob, kc, dfdc = got
    c = L.fields[0]
    etas = L.fields[2:]
    zero = {e: 0 for e in etas}
    # dfdc = 2 rho^2 (c - ca)(1 - H) + 2 rho^2 (c - cb) H
    a0 = sy.Poly(sy.expand(dfdc.subs(zero)), c)
    if a0.degree() != 1:
        return None
    rho2 = _num(a0.LC()) / 2.0
    ca = -_num(a0.TC()) / (2.0 * rho2)
    one = dict(zero)
    one[etas[0]] = 1
    a1 = sy.Poly(sy.expand(dfdc.subs(one)), c)
    cb = -_num(a1.TC()) / (2.0 * rho2)
    # Allen-Cahn rows (pfbase.py:396-408): v (eta - eta0)/dt + L (v dfdeta + kappa grad v . grad eta)
    Ls, kes, dfde = [], [], []
    for i, e in enumerate(etas):
        a = 2 + i
        if not _time_row(L, a):
            return None
        s = L.values(sy.expand(L.src[a] - (e - L.prev[a]) / L.dt))      # L dfdeta_i
        dfde.append(s)
        kes.append(_num(L.values(L.stiff[a][a])))
        if any(not _is_zero(L.stiff[a][b]) for b in range(6) if b != a):
            return None
    # w and alpha from the double well: at c = ca, only eta1: L (rho^2 (ca-cb)^2 h'(e) + w (2e(1-e)^2 - 2e^2(1-e)))
    e1, e2 = etas[0], etas[1]
    only1 = {c: ca, etas[1]: 0, etas[2]: 0, etas[3]: 0}
    lw2 = _num(sy.diff(dfde[0].subs(only1), e1).subs({e1: 0}))            # = 2 L w
    cross = _num(sy.diff(sy.diff(dfde[0].subs({c: ca, etas[2]: 0, etas[3]: 0}), e1), e2, 2).subs({e1: 0, e2: 0}))   # = 4 L w alpha
    # L from the c-dependence: d/dc dfde1 = L * 2 rho^2 ((c-cb) - (c-ca)) h'(e1)
    hp = 30 * 0.5 ** 2 * (1 - 0.5) ** 2
    Lc = _num(sy.diff(dfde[0], c).subs({etas[1]: 0, etas[2]: 0, etas[3]: 0, e1: 0.5, c: 0.0})) / (2 * rho2 * (ca - cb) * hp)
    w = lw2 / (2.0 * Lc)
    alpha = cross / (4.0 * Lc * w)
    ke = kes[0] / Lc
    H = sum(e ** 3 * (6 * e ** 2 - 15 * e + 10) for e in etas)
    Wd = sum(e ** 2 * (1 - e) ** 2 for e in etas) + alpha * sum(etas[i] ** 2 * etas[j] ** 2 for i in range(4) for j in range(i + 1, 4))
    f = rho2 * (c - ca) ** 2 * (1 - H) + rho2 * (c - cb) ** 2 * H + w * Wd
    if not _close(dfdc - sy.diff(f, c), 1e-8):
        return None
    for i, e in enumerate(etas):
        if not _close(dfde[i] - Lc * sy.diff(f, e), 1e-8) or abs(kes[i] - Lc * ke) > 1e-12:
            return None
    return "bench2", [ca, cb, rho2, kc, ke, Mob, w, alpha, Lc]


# ------------------------------------------------------------------------------------------------------------------
# functionals: the fused energy kernel returns (total free energy, second diagnostic); anything else is a correction
# ------------------------------------------------------------------------------------------------------------------
def plan_functional(form, name, params, fields, templates=None):
    """How to evaluate df.assemble(form): (which, scale, corrections) meaning
    scale * energy()[which] + sum_k coef_k * u_k^T K u_k  (which None: only the gradient pieces).
    `templates` = (energy density, {field: gradient coefficient}, second integrand, scale) of a generated model."""
    pt = sy.Integer(0)
    grads = {}
    for t in form.terms:
        if t[0] == "func":
            pt += t[1]
        elif t[0] == "gradfunc":
            if t[1] is not t[2] or not isinstance(t[1], FieldSymbol):
                raise NotImplementedError("gradient functionals other than |grad u_k|^2")
            grads[t[1].index] = grads.get(t[1].index, 0) + t[3]
        else:
            raise NotImplementedError("test function inside a functional")

    def values(e):
        e = sy.sympify(e)
        return sy.expand(e.subs({s: s.const.value for s in e.free_symbols if isinstance(s, ConstSymbol)}))

    pt = values(pt)
    grads = {k: _num(values(v)) for k, v in grads.items()}
    e_t, g_t, m_t, m_scale = templates if templates is not None else _functional_templates(name, params, fields)
    if _is_zero(pt):
        which, scale, base = None, 1.0, {}
    elif _close(pt - m_t, 1e-9):
        which, scale, base = 1, m_scale, {}
    elif _close(pt - e_t, 1e-9):
        which, scale, base = 0, 1.0, g_t
    else:
        raise NotImplementedError("the integrand is neither the benchmark's free-energy density nor its second diagnostic")
    corrections = []
    for k in sorted(set(grads) | set(base)):
        d = grads.get(k, 0.0) - base.get(k, 0.0)
        if abs(d) > 1e-14 * max(1.0, abs(base.get(k, 0.0))):
            corrections.append((k, d))
    return which, scale, corrections


def _functional_templates(name, p, F):
    """(free-energy density, {field: gradient coefficient}, second integrand, factor energy()[1] -> its integral)."""
    if name in ("bench1", "b13d"):
        c = F[0]
        return p[2] * (c - p[0]) ** 2 * (p[1] - c) ** 2, {0: p[3] / 2.0}, c, 1.0
    if name == "bench6":
        c, phi = F[0], F[2]
        return p[2] * (c - p[0]) ** 2 * (p[1] - c) ** 2 + p[5] * c * phi / 2.0, {0: p[3] / 2.0}, c, 1.0
    if name == "bench3":
        U, ph = F
        lam = p[3]
        e = -ph ** 2 / 2 + ph ** 4 / 4 + lam * U * ph * (1 - sy.Rational(2, 3) * ph ** 2 + sy.Rational(1, 5) * ph ** 4)
        return e, {1: 0.5 * p[0] ** 2}, 0.5 * (ph + 1.0), p[4]        # energy()[1] is the solid FRACTION (bench3.py:169-170)
    if name == "bench2":
        c, etas = F[0], F[2:]
        H = sum(e ** 3 * (6 * e ** 2 - 15 * e + 10) for e in etas)
        Wd = sum(e ** 2 * (1 - e) ** 2 for e in etas) + p[7] * sum(etas[i] ** 2 * etas[j] ** 2 for i in range(4) for j in range(i + 1, 4))
        f = p[2] * (c - p[0]) ** 2 * (1 - H) + p[2] * (c - p[1]) ** 2 * H + p[6] * Wd
        g = {0: p[3] / 2.0}
        g.update({2 + i: p[4] / 2.0 for i in range(4)})
        return f, g, c, 1.0
    raise KeyError(name)
