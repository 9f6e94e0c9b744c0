"""`import pfhubbenchmarks_b200.dolfin as df`: the slice of the legacy-FEniCS namespace the reference drivers touch.

A reference driver (dolfin/bench1.py, bench2.py, bench3.py, bench6.py, b13d.py) is a module-level program: it builds a
mesh, a mixed P1 space, Functions, a free energy and a weak form in UFL, hands `F, w, bcs, J` to
`NonlinearVariationalProblem`, configures the SNES parameter tree and time-steps with `solver.solve()` and
`df.assemble(...)` diagnostics.  With this module (and `pfhubbenchmarks_b200.pfbase` for `from pfbase import *`) the
same program text runs on libpfh_b200.so: expressions are sympy (forms.py), the weak form is recognised as one of the
compiled operator families and its coefficients are extracted from it, and every numerical operation -- interpolation of
the initial condition, Newton-Krylov solve, assembled functionals -- is a call through the C ABI on device vectors.

Nothing here computes on the CPU except evaluating a *user-defined* `UserExpression.eval` at the mesh vertices (the
reference does that in Python too, pfbase.py:177-193); the shipped InitialConditionsBench* classes run as kernels.
"""
import warnings

import numpy as np
import sympy as sy

from . import engine, forms
from .api import (DOLFIN_EPS, MPI, BoxMesh, Mesh, NonlinearVariationalSolver as _Solver, Point,  # noqa: F401
                  RectangleMesh, UnitSquareMesh, _IC)
from .forms import diff, dx, grad, inner, variable  # noqa: F401
from .io import File  # noqa: F401

# ---- UFL math on expressions (bench3.py:81 uses atan; the rest for user free energies) ----
atan, sin, cos, exp, sqrt, ln, tanh = sy.atan, sy.sin, sy.cos, sy.exp, sy.sqrt, sy.log, sy.tanh


class _Tree(dict):
    def __missing__(self, k):
        self[k] = _Tree()
        return self[k]


parameters = _Tree()      # df.parameters["form_compiler"]["quadrature_degree"] = 3  (bench1.py:12-14)


class LogLevel:
    DEBUG, TRACE, PROGRESS, INFO, WARNING, ERROR, CRITICAL = 10, 13, 16, 20, 30, 40, 50


_log_level = [LogLevel.INFO]


def set_log_level(level):
    _log_level[0] = level


def warning(msg):
    warnings.warn(msg)


def near(a, b, tol=DOLFIN_EPS):
    return abs(a - b) < tol


class SubDomain:
    def __init__(self, *a, **k):
        pass


class _SymArith:
    """Arithmetic of objects that stand for a sympy symbol (`.sym`) inside expressions."""

    def _op(self, other, fn):
        if isinstance(other, (forms.Grad, forms.GradPair, forms.Test, forms.Tested, forms.Mixed, forms.Form, forms.Measure)):
            return NotImplemented
        return fn(self.sym, forms._expr(other))

    def __add__(self, o): return self._op(o, lambda a, b: a + b)
    def __radd__(self, o): return self._op(o, lambda a, b: b + a)
    def __sub__(self, o): return self._op(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._op(o, lambda a, b: b - a)
    def __mul__(self, o): return self._op(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._op(o, lambda a, b: b * a)
    def __truediv__(self, o): return self._op(o, lambda a, b: a / b)
    def __rtruediv__(self, o): return self._op(o, lambda a, b: b / a)
    def __pow__(self, o): return self._op(o, lambda a, b: a ** b)
    def __neg__(self): return -self.sym


class Constant(_SymArith):
    """df.Constant: a symbol in forms, a float for the time loop (`float(dt)`, `dt.assign(x)`, bench1.py:29,171,180-183)."""

    def __new__(cls, value):
        if isinstance(value, (tuple, list)):
            if any(float(v) != 0.0 for v in value):
                raise NotImplementedError("vector Constants other than zero (pfbase.py:442 only adds f1 = (0, 0))")
            return forms.VectorZero()
        return object.__new__(cls)

    def __init__(self, value):
        self.value = float(value)
        self.sym = forms.ConstSymbol("k%d" % next(forms._counter), const=self)

    def assign(self, value):
        self.value = float(value)

    def __float__(self):
        return self.value

    def __repr__(self):
        return "Constant(%r)" % self.value


class _Element:
    pass


def MixedElement(elems):
    return list(elems)


class FunctionSpace:
    """df.FunctionSpace(mesh, 'P', 1) and df.FunctionSpace(mesh, MixedElement([...]))  (bench1.py:39-41)."""

    def __init__(self, mesh, family, degree=None, constrained_domain=None, _index=None):
        if constrained_domain is not None:
            raise NotImplementedError("periodic constrained_domain (commented out in every reference driver)")
        self.mesh, self.index = mesh, _index
        self.mixed = isinstance(family, list)          # a MixedElement space is indexed (w_[0]) even with one component
        if isinstance(family, list):
            self.nfields = len(family)
        else:
            if family not in ("P", "CG", "Lagrange") or degree != 1:
                raise NotImplementedError("only P1 Lagrange spaces (all the reference drivers use)")
            self.nfields = 1

    def ufl_element(self):
        return _Element()

    def split(self):                                   # _, _, Wphi = W.split()  (bench6.py:86)
        return [FunctionSpace(self.mesh, "P", 1, _index=i) for i in range(self.nfields)]

    def sub(self, i):
        return self.split()[i]


class UserExpression:
    """Base of user-defined initial conditions: subclasses implement eval(values, x) and value_shape (pfbase.py:177-193)."""

    def __init__(self, *a, **k):
        pass


class Expression:
    """df.Expression("sin(x[1]/7)", degree=2)  (bench6.py:84): a C-syntax scalar expression of x[0], x[1], x[2]."""

    def __init__(self, code, degree=None, **kw):
        if isinstance(code, (tuple, list)):
            raise NotImplementedError("vector-valued Expression")
        x0, x1, x2 = sy.symbols("x0 x1 x2", real=True)
        self.code = code
        self.expr = sy.sympify(code.replace("x[0]", "x0").replace("x[1]", "x1").replace("x[2]", "x2"),
                               locals={"x0": x0, "x1": x1, "x2": x2})


class DirichletBC:
    """df.DirichletBC(Wphi, value, boundary_fn)  (bench6.py:87-88; nl.py:35).  The boundary is identified by probing
    `boundary_fn` at the four edge midpoints of the rectangle: `boundary_fn(x, on_boundary)` as DOLFIN calls it, or a
    one-argument locator over an array of points (nl.py:34 writes `lambda x: np.logical_or(np.isclose(x[0], 0.0), ...)`).
    bench6 gives one edge per condition, nl.py the pair x = 0 and x = 1 in one."""

    def __init__(self, space, value, where):
        if space.index is None and space.nfields != 1:
            raise NotImplementedError("DirichletBC on the whole mixed space")
        self.field = space.index or 0
        if isinstance(value, Function):                # u_bc = Function(V); u_bc.interpolate(Constant(0.9))  (nl.py:31-32)
            if not isinstance(value._pending, Constant):
                raise NotImplementedError("Dirichlet data from a Function other than an interpolated Constant")
            value = value._pending
        self.value_expr = value.expr if isinstance(value, Expression) else sy.Float(float(value)) if float(value) != 0 else sy.Integer(0)
        lx, ly = space.mesh.length[:2]
        probes = {"left": (0.0, 0.5 * ly), "right": (lx, 0.5 * ly), "bottom": (0.5 * lx, 0.0), "top": (0.5 * lx, ly)}

        def inside(p, on_boundary):
            try:
                return bool(where(np.array(p), on_boundary))
            except TypeError:                          # one-argument locator
                return bool(where(np.array(p)))

        hit = [k for k, p in probes.items() if inside(p, True)]
        if inside((0.5 * lx, 0.5 * ly), False) or not hit:
            raise NotImplementedError("Dirichlet boundary must consist of edges of the rectangle")
        self.sides = tuple(hit)
        self.side = hit[0] if len(hit) == 1 else "+".join(hit)


class _FunctionIndex:
    """w_[i] of a TestFunction / dw[i] of a TrialFunction."""

    def __init__(self, space, test):
        self.space, self.test = space, test

    def __getitem__(self, i):
        if not 0 <= i < self.space.nfields:
            raise IndexError(i)
        return forms.Test(i) if self.test else None


def TestFunction(W):
    if not W.mixed:                                    # v = TestFunction(V) on a scalar space is used as is (nl.py:28,37)
        return forms.Test(0)
    return _FunctionIndex(W, True)


def TrialFunction(W):
    return _FunctionIndex(W, False)


class Function(_SymArith):
    """df.Function(W): symbols in forms; an fp64 device vector [field0 | field1 | ...] once a problem owns it.
    interpolate()/assign() before the problem exists are replayed when the storage is created.  On a scalar space the
    Function itself stands for its one symbol in expressions (`sqrt(u * u)`, `inner(u, v)`, nl.py:37)."""

    @property
    def sym(self):
        if self.W.nfields != 1:
            raise AttributeError("a Function on a mixed space enters forms through split(w) or w[i]")
        return self.symbols[0]

    def __init__(self, W):
        self.W = W
        k = next(forms._counter)
        self.symbols = [forms.FieldSymbol("w%d_%d" % (k, i), owner=self, index=i) for i in range(W.nfields)]
        self._tensor, self._pending, self._problem = None, None, None
        self._host = None          # vertex values of a Function read back from a stored series (HDF5File.read), host numpy

    def __getitem__(self, i):
        return self.symbols[i]

    @property
    def tensor(self):
        if self._tensor is None:
            raise RuntimeError("Function has no storage yet: create the NonlinearVariationalProblem first")
        return self._tensor

    def _bind(self, problem):
        self._problem = problem
        self._tensor = problem.ctx.new_vector()
        pend, self._pending = self._pending, None
        if pend is not None:
            self.interpolate(pend) if not isinstance(pend, Function) else self.assign(pend)

    def interpolate(self, ic):                         # w.interpolate(w_ic)  (bench1.py:134-135)
        if self._tensor is None:
            self._pending = ic
            return
        ctx = self._problem.ctx
        if isinstance(ic, Constant):                   # u.interpolate(Constant(0.9))
            self._tensor.fill_(float(ic))
        elif isinstance(ic, _IC) and ctx.problem in ic.kernel_problems:
            ctx.initial_condition(ic.args, out=self._tensor)      # the problem's own expression: evaluated by its kernel
        elif isinstance(ic, _IC):
            # another benchmark's expression, or a generated model (whose kernel has no initial condition): the closed
            # form on the host at the vertex coordinates, as DOLFIN's interpolate evaluates the UserExpression
            import torch
            if ctx.nranks != 1:
                raise NotImplementedError("host-evaluated initial condition on a slab-decomposed mesh")
            if ic.nvalues != ctx.nf:
                raise ValueError("%s has %d components, the function space %d" % (type(ic).__name__, ic.nvalues, ctx.nf))
            vals = ic.values(ctx.vertex_coordinates())
            self._tensor.copy_(torch.from_numpy(np.ascontiguousarray(vals).reshape(-1)).to(self._tensor.device))
        elif isinstance(ic, UserExpression):
            self._tensor.copy_(_eval_user_expression(ic, ctx))
        else:
            raise TypeError("cannot interpolate %r" % (ic,))

    def assign(self, other):                           # w0.assign(w), w.assign(w0)  (bench1.py:158,173)
        if self._tensor is None or other._tensor is None:
            self._pending = other
            return
        self._tensor.copy_(other._tensor)

    def split(self):                                   # c, _ = w.split()  (bench1.py:188): components to write / evaluate
        return [SubFunction(self, i) for i in range(self.W.nfields)]

    def sub(self, i):                                  # w.sub(0)
        return SubFunction(self, i)

    def vector(self):
        return self.tensor

    def to_numpy(self):
        return self.tensor.cpu().numpy().reshape(self.W.nfields, -1)


class SubFunction(_SymArith):
    """One component of a mixed Function (`c, _ = w.split()`, `w.sub(0)`).  In expressions it is the same symbol as
    `df.split(w)[i]` -- the reference assembles `f_chem*dx + kappa/2*inner(grad(c), grad(c))*dx` and `c*dx` with the
    split() component (bench1.py:188-194); `f(x)` evaluates its P1 interpolant at a point (what pfbase.peval calls,
    pfbase.py:50-62), `f.at(points)` many points in one kernel launch; `file << (c, t)` writes it."""

    def __init__(self, owner, index):
        self.owner, self.index = owner, index

    @property
    def sym(self):
        return self.owner.symbols[self.index]

    def at(self, points):
        return self.owner._problem.ctx.eval_points(self.owner.tensor, self.index, points)

    def __call__(self, x):
        v = float(self.at(np.asarray(x, dtype=float)[None, :])[0])
        if not np.isfinite(v):
            raise RuntimeError("point %r is outside the mesh" % (tuple(x),))     # dolfin raises the same way
        return v

    def value_shape(self):
        return ()

    @property
    def tensor(self):                                  # view of the component inside the owner's device vector
        return self.owner._problem.ctx.fields(self.owner.tensor)[self.index]

    def cpu(self):
        return self.tensor.cpu()

    def function_space(self):
        return self.owner.W.split()[self.index]


def split(w):
    """c, mu = df.split(w)  (bench1.py:59): the same symbols as w[0], w[1]."""
    return tuple(w.symbols)


def _eval_user_expression(ic, ctx):
    """Vertex values of a user-defined expression, in the library's vertex order (single rank)."""
    import torch
    if ctx.nranks != 1:
        raise NotImplementedError("user-defined UserExpression on a slab-decomposed mesh")
    xy = ctx.vertex_coordinates()
    vals = np.zeros((xy.shape[0], ctx.nf))
    for i in range(xy.shape[0]):
        ic.eval(vals[i], xy[i])
    return torch.from_numpy(np.ascontiguousarray(vals.T).reshape(-1)).to(ctx.new_vector().device)


def derivative(F, w, dw):
    """J = df.derivative(F, w, dw)  (bench1.py:69): the compiled operators carry their exact Gateaux derivative
    (csrc/pfh_problems.cuh dg), so J is only a marker that the Jacobian of F is wanted."""
    return ("derivative", F)


class NonlinearVariationalProblem:
    """df.NonlinearVariationalProblem(F, w, bcs, J)  (bench1.py:76).  F is lowered and matched to a compiled operator
    family (forms.compile_form); the physical coefficients are read out of the form itself."""

    def __init__(self, F, w, bcs=None, J=None, device=0, rank=0, nranks=1, nccl_id=None, krylov_basis=30, functionals=None):
        """`functionals=(energy_form, second_form)`: only for forms that are none of the shipped systems -- the two
        integrals the generated model's fused reduction kernel evaluates (df.assemble of exactly these is then one
        launch); the shipped systems have theirs built in."""
        problem, name, params, dt, prev, self.model = self._analyse(F, w, bcs, J, functionals)
        mesh = w.W.mesh
        if (name == "nl") != (mesh.kind == "right2d"):
            raise NotImplementedError("the right-diagonal UnitSquareMesh carries the steady problem of nl.py and nothing else does"
                                      " (the benchmark drivers use RectangleMesh(..., 'crossed') / BoxMesh)")
        if name == "bench1" and mesh.kind == "kuhn3d":
            name = problem = "b13d"
        self.benchmark, self.mesh, self.dt, self.params = name, mesh, dt, params
        self._frozen = [(s.const, s.const.value) for t in F.terms for e in t[1:] if isinstance(e, sy.Basic)
                        for s in e.free_symbols if isinstance(s, forms.ConstSymbol) and s.const is not dt]
        self.ctx = engine.Context(problem, mesh.ncell, mesh.length, params=params, device=device, rank=rank, nranks=nranks,
                                  nccl_id=nccl_id, krylov_basis=krylov_basis)
        self.w, self.w0 = w, (prev if prev is not None else w)     # steady: the library ignores w0
        w._bind(self)
        if prev is not None:
            prev._bind(self)
        self._functionals = {}


def _analyse(F, w, bcs, J, functionals):
    """Everything the problem constructor does before it touches the GPU: checks, recognition of the shipped systems or
    code generation for anything else.  Returns (problem for engine.Context, name, params, dt Constant, w0, model)."""
    if parameters["form_compiler"].get("quadrature_degree", 3) != 3:
        raise NotImplementedError("the compiled operators integrate with the degree-3 rules the reference sets (bench1.py:14)")
    if J is not None and not (isinstance(J, tuple) and J[0] == "derivative" and J[1] is F):
        raise NotImplementedError("J must be df.derivative(F, w, dw)")
    prev = None
    for t in F.terms:
        if t[0] == "src":
            for s in sy.sympify(t[2]).free_symbols:
                if isinstance(s, forms.FieldSymbol) and s.owner is not w:
                    prev = s.owner
    if prev is None:                                   # no time derivative: the steady problem of nl.py, or nothing we know
        name, params = forms.compile_steady(F, w.W.nfields, w.symbols, bcs or ())
        return name, name, params, None, None, None
    try:
        name, params, dt = forms.compile_form(F, w.W.nfields, w.symbols, prev.symbols, bcs or ())
        return name, name, params, dt, prev, None
    except NotImplementedError:
        # not one of the shipped systems: generate a functor for it and build its own copy of the library
        # (what FFC does for every form at this point in the reference, bench1.py:76)
        from . import codegen
        if bcs:
            raise NotImplementedError("Dirichlet conditions on a generated model")
        L = codegen.lower(F, w.W.nfields, w.symbols, prev.symbols)
        e_form, m_form = functionals if functionals is not None else (None, None)
        model = codegen.UserModel(L, e_form, m_form)
        return model, "user_" + model.tag, [], model.dt_const, prev, model


NonlinearVariationalProblem._analyse = staticmethod(_analyse)


def compile_only(F, w, bcs=None, J=None, functionals=None):
    """Build (or find in the cache) the library a form needs, without creating a GPU context: None for the shipped
    systems, the path of the generated copy otherwise.  Lets a GPU-less build host pre-compile a model."""
    from . import build
    model = _analyse(F, w, bcs, J, functionals)[5]
    return None if model is None else build.build_user(model.source, model.tag)


class NonlinearVariationalSolver(_Solver):
    def solve(self):
        """As api.NonlinearVariationalSolver.solve; additionally refuses to run with a stale model: the coefficients
        were read out of the form (or baked into generated code) when the problem was created, so a Constant other than
        dt that has been re-assigned since would silently be ignored, where DOLFIN would honour it."""
        for const, value in self.problem._frozen:
            if const.value != value:
                raise RuntimeError("%r was %r when the problem was created; only dt may change between solves -- "
                                   "create a new NonlinearVariationalProblem" % (const, value))
        return super().solve()


def assemble(form):
    """df.assemble of a functional (bench1.py:121-125, bench2.py:184-194, bench3.py:162-170, bench6.py:155-165).
    The pointwise part must be the benchmark's free-energy density or its second diagnostic (both come out of the
    fused energy kernel); gradient pieces with other coefficients are corrected with pfh_bilinear_dev."""
    owner = None
    for t in form.terms:
        syms = sy.sympify(t[1]).free_symbols if t[0] == "func" else {t[1], t[2]}
        for s in syms:
            if isinstance(s, forms.FieldSymbol):
                owner = s.owner
    if owner is None or owner._problem is None:
        raise RuntimeError("assemble: the functional does not involve the Function of a NonlinearVariationalProblem")
    pr = owner._problem
    key = repr(form.terms)
    plan = pr._functionals.get(key)
    if plan is None:
        tmpl = None
        if pr.model is not None:       # generated model: its own energy density / second integrand
            sub = dict(zip(pr.model.U, owner.symbols))
            tmpl = (sy.sympify(pr.model.e_expr).subs(sub), dict(pr.model.grad), sy.sympify(pr.model.m_expr).subs(sub), 1.0)
        plan = forms.plan_functional(form, pr.benchmark, pr.ctx.params, owner.symbols, tmpl)
        pr._functionals[key] = plan
    which, scale, corrections = plan
    val = scale * pr.ctx.energy(owner.tensor)[which] if which is not None else 0.0
    for field, coef in corrections:
        val += coef * pr.ctx.bilinear(owner.tensor, field, field, 0.0, 1.0)
    return val


class HDF5File:
    """`outfile = HDF5File(MPI.comm_world, filename + ".h5", "w")`; `outfile.write(mesh, "mesh")`;
    `outfile.write(c, "c", float(t))`  (bench1.py:116-118,191).  HDF5 is not available in this image: the same time
    series is written as <filename>.pvd + one VTU per step (what b13d.py / bench2.py write), with a warning.
    Mode "r" reads that series back the way the reference's post-processing reads its container
    (`hdf.read(mesh, "mesh", True)`, `hdf.read(c, f"c/vector_{i}")`, process_bench1.py:8-27, b1_read.py:9-30)."""

    def __init__(self, comm, path, mode="w"):
        base = path[:-3] if path.endswith(".h5") else path
        self._mode = mode
        if mode == "r":
            from . import io
            self._entries = io.read_pvd(base + ".pvd")
            if not self._entries:
                raise RuntimeError("%s.pvd holds no time steps" % base)
            self._geom = None
            return
        if mode != "w":
            raise NotImplementedError("HDF5File mode %r" % (mode,))
        self._pvd = File(base + ".pvd", "compressed")
        warnings.warn("HDF5 is not available: writing %s (VTU series) instead of %s" % (self._pvd.path, path))

    def read(self, obj, name, *args):
        from . import io
        if self._mode != "r":
            raise RuntimeError("HDF5File opened for writing")
        if isinstance(obj, Mesh):                      # hdf.read(mesh, "mesh", True)
            pts, cells, _ = io.read_vtu(self._entries[0][1])
            obj.kind, obj.ncell, obj.length = io.identify_mesh(pts, cells)
            return
        field, _, vec = name.partition("/vector_")      # "c/vector_12": step 12 of field c
        if not vec.isdigit() or int(vec) >= len(self._entries):
            raise RuntimeError("no dataset %r in the series (%d steps)" % (name, len(self._entries)))
        data = io.read_vtu(self._entries[int(vec)][1])[2]
        if field not in data and len(data) == 1:
            field = next(iter(data))
        obj._host = np.array(data[field], dtype=np.float64)

    def attributes(self, name):                        # hdf.attributes("c")["count"]: number of stored steps
        return {"count": len(self._entries)}

    def write(self, obj, name, t=None):
        if isinstance(obj, Mesh):                      # outfile.write(mesh, "mesh"): every VTU carries the mesh
            self._pvd.mesh = obj
            return
        self._pvd.name = name
        self._pvd << (obj, 0.0 if t is None else t)

    def close(self):
        pass
