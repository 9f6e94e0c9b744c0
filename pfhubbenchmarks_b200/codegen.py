"""sympy -> CUDA functor: a weak form that is none of the shipped systems gets its own copy of libpfh_b200.so.

`forms.compile_form` recognises the reference's four systems and runs them on the precompiled, hand-tuned functors.
Any other form written with the same builders (pfbase.py: Cahn-Hilliard / Allen-Cahn / diffusion / Poisson /
backward-Euler pieces, any smooth free energy) is lowered here to the common discrete form

    R_a = sum_b (Mc[a][b] M + Kc[a][b] K) u_b - Mp[a] M u0_a + N[g_a](u)

and emitted as `struct ProbUser` with the interface of csrc/pfh_problems.cuh: the constant-coefficient linear part goes
to `lin()` (stencils), the rest of the pointwise source to `g` / `dg` (quadrature; derivatives by sympy, common
subexpressions eliminated), the energy and second functional to `e` / `m` / `grad_*`, and the block structure the
preconditioner needs to `pc_layout`.  build.build_user compiles it (nvcc, sm_100a) into pfhubbenchmarks_b200/_jit/.
The reference does the same job with FFC at `df.NonlinearVariationalProblem` time (UFL -> C++ -> JIT, bench1.py:76).

Constants other than dt are baked in at generation time.  Dirichlet conditions are not generated (no-flux only).
"""
import hashlib

import sympy as sy
from sympy.printing.c import C99CodePrinter

from . import forms


class _Printer(C99CodePrinter):
    """Small integer powers as products (pow() in fp64 device code is a slow library call)."""

    def _print_Pow(self, expr):
        b, e = expr.as_base_exp()
        if e.is_Integer and 2 <= int(e) <= 6:
            s = self._print(b)
            s = s if b.is_Atom else "(" + s + ")"
            return "(" + "*".join([s] * int(e)) + ")"
        if e.is_Integer and -6 <= int(e) <= -1:
            return "(1.0/(" + self._print(b ** (-e)) + "))"
        return super()._print_Pow(expr)

    def _print_Symbol(self, expr):
        return expr.name


_P = _Printer()


def _body(exprs, targets, indent="    "):
    """C statements computing targets[i] = exprs[i] with common subexpressions in locals."""
    repl, red = sy.cse([sy.sympify(e) for e in exprs], symbols=sy.numbered_symbols("t_"), optimizations="basic")
    lines = ["%sconst double %s = %s;" % (indent, _P.doprint(s), _P.doprint(e)) for s, e in repl]
    lines += ["%s%s = %s;" % (indent, t, _P.doprint(e)) for t, e in zip(targets, red)]
    return "\n".join(lines)


def _table(ctype, vals):
    vals = list(vals) or [0]
    return "constexpr %s t[] = {%s}; return t[i];" % (ctype, ", ".join(str(v) for v in vals))


class UserModel:
    """The lowered model: coefficient expressions, pointwise functions and the generated header text."""

    def __init__(self, L, energy=None, second=None, field_names=None):
        nf, U, U0, dt = L.nf, list(L.fields), list(L.prev), L.dt
        self.nf, self.dt_sym = nf, dt
        self.field_names = tuple(field_names or ("u%d" % i for i in range(nf)))
        self.dt_const = dt.const
        Mc = [[sy.Integer(0)] * nf for _ in range(nf)]
        Mp = [sy.Integer(0)] * nf
        Kc = [[L.values(L.stiff[a][b]) for b in range(nf)] for a in range(nf)]
        G = []
        for a in range(nf):
            rest = L.values(L.src[a])
            # previous-step terms: only -Mp_a u0_a with a field-free coefficient fits the discrete form
            for b in range(nf):
                cb = sy.diff(rest, U0[b])
                if cb == 0:
                    continue
                if b != a or cb.free_symbols - {dt}:
                    raise NotImplementedError("previous-step terms other than a constant multiple of the row's own field")
                Mp[a] = -cb
                rest = sy.expand(rest - cb * U0[b])
            # constant-coefficient linear terms go to the mass stencil, everything else is a pointwise load
            load = sy.Integer(0)
            for term in sy.Add.make_args(sy.expand(rest)):
                coef, dep = term.as_independent(*U)
                if dep in U:
                    Mc[a][U.index(dep)] += coef
                else:
                    load += term
            rest = load
            if rest.has(dt) or any(rest.has(p) for p in U0):
                raise NotImplementedError("time-step dependent nonlinear terms")
            if any(sy.sympify(Kc[a][b]).free_symbols - {dt} for b in range(nf)):
                raise NotImplementedError("field-dependent diffusion coefficients")
            G.append(rest)
        self.Mc, self.Kc, self.Mp, self.G = Mc, Kc, Mp, G
        self.inputs = [b for b in range(nf) if any(g.has(U[b]) for g in G)] or [0]
        self.outputs = [a for a in range(nf) if G[a] != 0] or [1 if nf > 1 else 0]
        self.U = U
        # functionals: energy density + gradient coefficients, second integrand
        self.e_expr, self.grad, self.m_expr = sy.Integer(0), {}, U[0]
        if energy is not None:
            pt = sy.Integer(0)
            for t in energy.terms:
                if t[0] == "func":
                    pt += t[1]
                elif t[0] == "gradfunc" and t[1] is t[2]:
                    self.grad[t[1].index] = self.grad.get(t[1].index, 0) + float(L.values(t[3]))
                else:
                    raise NotImplementedError("energy functional term %r" % (t[0],))
            self.e_expr = L.values(pt)
        if second is not None:
            if any(t[0] != "func" for t in second.terms):
                raise NotImplementedError("second functional must be a pointwise integrand")
            self.m_expr = L.values(sum(t[1] for t in second.terms))
        self.layout = self._pc_layout()
        self.source = self._emit()
        self.tag = hashlib.sha256(self.source.encode()).hexdigest()[:16]

    # ---- preconditioner block structure ----
    def _num(self, e, dtv=0.1):
        return float(sy.sympify(e).subs(self.dt_sym, dtv))

    def _pc_layout(self):
        nf, Mc, Kc, Mp = self.nf, self.Mc, self.Kc, self.Mp
        lay = dict(has_ch=False, Mob=0.0, kappa=0.0, spd=[], jacobi_only=False)
        rest = list(range(nf))
        if nf >= 2:
            dt = self.dt_sym
            z = lambda e: forms._is_zero(sy.sympify(e))
            ch = (z(Mc[0][0] - 1 / dt) and z(Mp[0] - 1 / dt) and z(Kc[0][0]) and z(Kc[1][1]) and z(Mc[1][1] - 1)
                  and z(Mp[1]) and self._num(Kc[0][1]) > 0 and self._num(Kc[1][0]) < 0)
            if ch:
                lay.update(has_ch=True, Mob=self._num(Kc[0][1]), kappa=-self._num(Kc[1][0]))
                rest = rest[2:]
        for a in rest:
            if Mp[a] != 0 and self._num(Mc[a][a]) > 0 and self._num(Kc[a][a]) >= 0 and len(lay["spd"]) < 4:
                lay["spd"].append(a)
            else:
                lay["jacobi_only"] = True
        return lay

    # ---- code ----
    def _emit(self):
        nf, U = self.nf, self.U
        dt = self.dt_sym
        insym = {U[b]: sy.Symbol("u[%d]" % i) for i, b in enumerate(self.inputs)}
        fldsym = {U[b]: sy.Symbol("u[%d]" % b) for b in range(nf)}
        xs = [sy.Symbol("x[%d]" % i) for i in range(len(self.inputs))]
        g = [self.G[a].subs(insym) for a in self.outputs]
        dg = [sum(sy.diff(self.G[a], U[b]).subs(insym) * xs[i] for i, b in enumerate(self.inputs)) for a in self.outputs]
        lin = []
        for a in range(nf):
            for b in range(nf):
                if self.Mc[a][b] != 0:
                    lin.append("    L.Mc[%d][%d] = %s;" % (a, b, _P.doprint(sy.sympify(self.Mc[a][b]).subs(dt, sy.Symbol("dt")))))
                if sy.sympify(self.Kc[a][b]) != 0:
                    lin.append("    L.Kc[%d][%d] = %s;" % (a, b, _P.doprint(sy.sympify(self.Kc[a][b]).subs(dt, sy.Symbol("dt")))))
            if self.Mp[a] != 0:
                lin.append("    L.Mp[%d] = %s;" % (a, _P.doprint(sy.sympify(self.Mp[a]).subs(dt, sy.Symbol("dt")))))
        gf = sorted(self.grad)
        lay = self.layout
        pc = ["    P.has_ch = %s; P.Mob = %r; P.kappa = %r; P.jacobi_only = %s; P.n_spd = %d;" % (
            "true" if lay["has_ch"] else "false", lay["Mob"], lay["kappa"], "true" if lay["jacobi_only"] else "false", len(lay["spd"]))]
        for k, a in enumerate(lay["spd"]):
            pc.append("    P.spd_field[%d] = %d; P.spd_am[%d] = lc.Mc[%d][%d]; P.spd_ak[%d] = lc.Kc[%d][%d]; P.spd_sign[%d] = 1.0;" % (
                k, a, k, a, a, k, a, a, k))
        return """// generated by pfhubbenchmarks_b200/codegen.py -- one user model, do not edit
// fields: %(names)s
struct ProbUser {
  static constexpr int NF = %(nf)d, NIN = %(nin)d, NOUT = %(nout)d;
  static constexpr bool QUARTIC = false;
  static constexpr bool HAS_BC = false;
  PFH_HD static constexpr int in(int i) { %(in_t)s }
  PFH_HD static constexpr int out(int i) { %(out_t)s }
  explicit ProbUser(const ProblemParams&) {}
  PFH_HD void g(const double* u, double* o) const {
%(g)s
  }
  PFH_HD void dg(const double* u, const double* x, double* o) const {
%(dg)s
  }
  PFH_HD double e(const double* u) const {
    double r;
%(e)s
    return r;
  }
  PFH_HD double m(const double* u) const {
    double r;
%(m)s
    return r;
  }
  PFH_HD static constexpr int n_grad() { return %(ngrad)d; }
  PFH_HD int grad_field(int i) const { %(gf_t)s }
  PFH_HD double grad_coef(int i) const { %(gc_t)s }
  PFH_HD bool dirichlet(int, int, int, int, double, double&) const { return false; }
  void lin(double dt, LinCoefs& L) const {
    (void)dt;
%(lin)s
  }
  PFH_HD void ic(const double*, double, double, double* u) const {   // initial data come from the host (UserExpression)
    for (int a = 0; a < NF; ++a) u[a] = 0.0;
  }
  static void pc_layout(const LinCoefs& lc, PfhPcLayout& P) {
    (void)lc;
%(pc)s
  }
};
""" % dict(names=", ".join(self.field_names), nf=nf, nin=len(self.inputs), nout=len(self.outputs),
           in_t=_table("int", self.inputs), out_t=_table("int", self.outputs),
           g=_body(g, ["o[%d]" % i for i in range(len(g))]), dg=_body(dg, ["o[%d]" % i for i in range(len(dg))]),
           e=_body([self.e_expr.subs(fldsym)], ["r"]), m=_body([sy.sympify(self.m_expr).subs(fldsym)], ["r"]),
           ngrad=max(1, len(gf)), gf_t=_table("int", gf or [0]), gc_t=_table("double", [repr(float(self.grad[k])) for k in gf] or ["0.0"]),
           lin="\n".join(lin), pc="\n".join(pc))


lower = forms.lower
