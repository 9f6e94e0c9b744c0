"""The reference drivers' Python surface on top of libpfh_b200.so.

The reference scripts (dolfin/bench{1,2,3,6}.py, b13d.py) are module-level programs that build
`mesh, W, w, w0, dt, F, J, bcs`, then

    problem = df.NonlinearVariationalProblem(F, w, bcs, J)          # bench1.py:76
    solver  = df.NonlinearVariationalSolver(problem)                # bench1.py:77
    solver.parameters['nonlinear_solver'] = 'snes'                  # bench1.py:82
    nlparams = solver.parameters['snes_solver']; nlparams[...] = ...   # bench1.py:83-110
    niters, converged = solver.solve()                              # bench1.py:162

and loop with the adaptive-dt rule.  This module keeps those object names, parameter keys and call
semantics (solve() mutates w in place from its current value, reads w0 and the current dt, returns
(niters, converged) and never raises on non-convergence).  What it does NOT re-implement is UFL:
the weak form is selected by benchmark name (the four forms the reference ships, pfbase.py:361-460)
and its coefficients are passed as keyword arguments named as in the reference scripts.
"""
import warnings

import numpy as np

from . import _lib, engine

DOLFIN_EPS = 3.0e-16


class Point:
    def __init__(self, *xyz):
        self.x = tuple(float(v) for v in xyz)


class Constant:
    """df.Constant: float(dt), dt.assign(x)  (bench1.py:29,171,180-183)"""

    def __init__(self, value):
        self._v = float(value)

    def assign(self, value):
        self._v = float(value)

    def __float__(self):
        return self._v

    def __repr__(self):
        return "Constant(%r)" % self._v


class Mesh:
    """A structured DOLFIN mesh by kind and size; `Mesh()` is the empty mesh a stored series is read into
    (`mesh = Mesh(); hdf.read(mesh, "mesh", True)`, process_bench1.py:14-15)."""

    def __init__(self, kind=None, ncell=(), length=()):
        self.kind, self.ncell, self.length = kind, tuple(int(n) for n in ncell), tuple(float(v) for v in length)

    def mpi_comm(self):
        return None


def RectangleMesh(p0, p1, nx, ny, diagonal="crossed"):
    """df.RectangleMesh(Point, Point, Nx, Ny, 'crossed')  (bench1.py:21-23). Only the crossed split the drivers use."""
    if diagonal != "crossed":
        raise NotImplementedError("only the 'crossed' diagonal of the reference drivers is supported")
    if any(abs(v) > 0 for v in p0.x):
        raise NotImplementedError("meshes start at the origin in all reference drivers")
    return Mesh("crossed2d", (nx, ny), p1.x[:2])


def UnitSquareMesh(nx, ny, diagonal="right"):
    """df.UnitSquareMesh(Nx, Ny[, 'crossed'])  (nl.py:19-20): DOLFIN's default split cuts every square by the diagonal
    from its lower-left to its upper-right vertex ('right')."""
    if diagonal == "crossed":
        return Mesh("crossed2d", (nx, ny), (1.0, 1.0))
    if diagonal != "right":
        raise NotImplementedError("UnitSquareMesh diagonal %r (the reference uses the default 'right' and 'crossed')" % (diagonal,))
    return Mesh("right2d", (nx, ny), (1.0, 1.0))


def BoxMesh(p0, p1, nx, ny, nz):
    """df.BoxMesh(Point, Point, Nx, Ny, Nz)  (b13d.py:24-26)."""
    if any(abs(v) > 0 for v in p0.x):
        raise NotImplementedError("meshes start at the origin in all reference drivers")
    return Mesh("kuhn3d", (nx, ny, nz), p1.x[:3])


class _CommWorld:
    """MPI.comm_world of the drivers (bench1.py:117, pfbase.py:73): rank / size of the torch.distributed job."""

    @property
    def rank(self):
        return MPI.rank()

    @property
    def size(self):
        try:
            import torch.distributed as dist
            return dist.get_world_size() if dist.is_initialized() else 1
        except Exception:
            return 1


class MPI:
    """df.MPI.rank(mesh.mpi_comm()) guards (bench1.py:150): rank of this process in the torch.distributed job."""
    comm_world = _CommWorld()

    @staticmethod
    def rank(_comm=None):
        try:
            import torch.distributed as dist
            return dist.get_rank() if dist.is_initialized() else 0
        except Exception:
            return 0


class Function:
    """df.Function on the mixed space: a fp64 device vector [field0 | field1 | ...] owned by torch."""

    def __init__(self, problem):
        self._p = problem
        self.tensor = problem.ctx.new_vector()

    def assign(self, other):                      # w0.assign(w), w.assign(w0)  (bench1.py:158,173)
        self.tensor.copy_(other.tensor)

    def interpolate(self, ic):                    # w.interpolate(w_ic)  (bench1.py:134-135)
        ctx = self._p.ctx
        if ctx.problem in ic.kernel_problems:
            ctx.initial_condition(ic.args, out=self.tensor)
            return
        # an expression of another benchmark: its closed form on the host, never this problem's kernel formula
        import torch
        if ic.nvalues != ctx.nf or ctx.nranks != 1:
            raise ValueError("%s (%d components) does not fit this problem (%d fields)" % (type(ic).__name__, ic.nvalues, ctx.nf))
        vals = ic.values(ctx.vertex_coordinates())
        self.tensor.copy_(torch.from_numpy(np.ascontiguousarray(vals).reshape(-1)).to(self.tensor.device))

    def split(self):                              # c, _ = w.split()  (bench1.py:188)
        return self._p.ctx.fields(self.tensor)

    def vector(self):
        return self.tensor

    def to_numpy(self):
        return self.tensor.cpu().numpy().reshape(self._p.ctx.nf, -1)


class _IC:
    """The reference's initial-condition expressions (dolfin/pfbase.py:177-339).  On a shipped problem of the matching
    benchmark they are evaluated by the library's kernel (`kernel_problems`); anywhere else -- a generated model, or an
    expression of another benchmark -- `values(xy)` evaluates the same closed form on the host, vertex by vertex as
    DOLFIN's interpolate does."""
    kernel_problems = ()
    nvalues = 0

    def __init__(self, *args, **kwargs):
        self.args = [float(a) for a in args]

    def value_shape(self):
        return (self.nvalues,)

    def values(self, xy):
        raise NotImplementedError

    def eval(self, values, x):                       # UserExpression.eval signature, one point
        values[:self.nvalues] = self.values(np.asarray(x, dtype=float)[None, :])[:, 0]


def _spinodal_c(x, y, c0, eps, kx=0.105):
    return c0 + eps * (np.cos(kx * x) * np.cos(0.11 * y) + (np.cos(0.13 * x) * np.cos(0.087 * y)) ** 2
                       + np.cos(0.025 * x - 0.15 * y) * np.cos(0.07 * x - 0.02 * y))


class InitialConditionsBench1(_IC):   # pfbase.py:177-193
    kernel_problems, nvalues = ("bench1", "b13d"), 2

    def values(self, xy):
        c0, eps = self.args[:2]
        return np.stack([_spinodal_c(xy[:, 0], xy[:, 1], c0, eps), np.zeros(len(xy))])


class InitialConditionsBench2(_IC):   # pfbase.py:262-291
    kernel_problems, nvalues = ("bench2",), 6

    def values(self, xy):
        c0, eps, eps_eta, psi = self.args[:4]
        x, y = xy[:, 0], xy[:, 1]
        out = [_spinodal_c(x, y, c0, eps), np.zeros(len(xy))]
        for i in range(4):
            ii = i + 1.0
            # the squared term uses i, not ii (pfbase.py:285-286); the reference's committed fields confirm it
            sq = (np.cos((0.046 + 0.001 * i) * x - (0.0405 + 0.001 * i) * y) * np.cos((0.031 + 0.001 * i) * x - (0.004 + 0.001 * i) * y)) ** 2
            out.append(eps_eta * (np.cos(0.01 * ii * x - 4.0) * np.cos((0.007 + 0.01 * ii) * y)
                                  + np.cos((0.11 + 0.01 * ii) * x) * np.cos((0.11 + 0.01 * ii) * y) + psi * sq) ** 2)
        return np.stack(out)


class InitialConditionsBench3(_IC):   # pfbase.py:293-320
    kernel_problems, nvalues = ("bench3",), 2

    def values(self, xy):
        delta, r0, wd, vin, vout = self.args[:5]
        r = np.hypot(xy[:, 0], xy[:, 1])
        ramp = vout + 0.5 * (vin - vout) * (1.0 + np.cos(np.pi * (r - r0 + 0.5 * wd) / wd))
        phi = np.where(r < r0 - 0.5 * wd, vin, np.where(r > r0 + 0.5 * wd, vout, ramp))
        return np.stack([np.full(len(xy), delta), phi])


class InitialConditionsBench6(_IC):   # pfbase.py:322-339
    kernel_problems, nvalues = ("bench6",), 3

    def values(self, xy):
        c0, c1 = self.args[:2]
        return np.stack([_spinodal_c(xy[:, 0], xy[:, 1], c0, c1, kx=0.2), np.zeros(len(xy)), np.zeros(len(xy))])


class NonlinearVariationalProblem:
    """Stands in for `df.NonlinearVariationalProblem(F, w, bcs, J)`.  F and J are implied by `benchmark`
    ('bench1', 'b13d', 'bench2', 'bench3', 'bench6') and its coefficients; dt is the mutable Constant the
    time loop re-assigns.  `device`, `rank`, `nranks`, `nccl_id` place the problem (z-slabs for 3-D)."""

    def __init__(self, benchmark, mesh, dt, params=None, device=0, rank=0, nranks=1, nccl_id=None, krylov_basis=30):
        self.benchmark, self.mesh, self.dt = benchmark, mesh, dt
        self.ctx = engine.Context(benchmark, mesh.ncell, mesh.length, params=params, device=device, rank=rank,
                                  nranks=nranks, nccl_id=nccl_id, krylov_basis=krylov_basis)
        self.w, self.w0 = Function(self), Function(self)


class _Params(dict):
    """Parameter tree with the reference's keys; unknown keys warn instead of raising (dolfin would raise)."""

    def __init__(self, known, **kw):
        super().__init__(**kw)
        self._known = set(known) | set(kw)

    def __setitem__(self, k, v):
        if k not in self._known:
            warnings.warn("unknown solver parameter %r ignored" % (k,))
        super().__setitem__(k, v)


_LINE_SEARCH = {"basic": 0, "cp": 1, "bt": 2}


class NonlinearVariationalSolver:
    def __init__(self, problem):
        self.problem = problem
        krylov = _Params(["maximum_iterations", "error_on_nonconvergence", "monitor_convergence", "relative_tolerance",
                          "absolute_tolerance", "restart"],
                         maximum_iterations=1000, relative_tolerance=1e-5, absolute_tolerance=1e-50, restart=30,
                         error_on_nonconvergence=False, monitor_convergence=False)
        snes = _Params(["report", "error_on_nonconvergence", "absolute_tolerance", "relative_tolerance",
                        "solution_tolerance", "maximum_iterations", "line_search", "linear_solver", "preconditioner",
                        "krylov_solver", "method"],
                       report=True, error_on_nonconvergence=False, absolute_tolerance=1e-10, relative_tolerance=1e-9,
                       solution_tolerance=1e-16, maximum_iterations=50, line_search="basic", linear_solver="gmres",
                       preconditioner="schur_chebyshev", krylov_solver=krylov)
        self.parameters = _Params(["nonlinear_solver", "snes_solver", "newton_solver"], nonlinear_solver="snes",
                                  snes_solver=snes, newton_solver=snes)
        self.last_report = None

    def _opts(self):
        # nl.py:66-96 configures and runs both trees; here they are one object (DOLFIN's defaults for them coincide)
        p = self.parameters["newton_solver" if self.parameters["nonlinear_solver"] == "newton" else "snes_solver"]
        k = p["krylov_solver"]
        ls = p["line_search"]
        if ls not in _LINE_SEARCH:
            warnings.warn("line_search %r is not implemented; using 'basic' (full Newton step)" % (ls,))
        lin = {"gmres": 0, "default": 0, "bicgstab": 1}.get(p["linear_solver"])      # bench1.py:98, bench2_eta1.py:113
        if lin is None:
            warnings.warn("linear_solver %r is not implemented; using gmres" % (p["linear_solver"],))
            lin = 0
        pc = {"schur_chebyshev": 0, "sor": 0, "default": 0, "jacobi": 1, "none": 2}.get(p["preconditioner"])
        if pc is None:
            warnings.warn("preconditioner %r is not implemented; using the Schur/Chebyshev preconditioner" % (p["preconditioner"],))
            pc = 0
        return _lib.default_opts(
            absolute_tolerance=float(p["absolute_tolerance"]), relative_tolerance=float(p["relative_tolerance"]),
            solution_tolerance=float(p["solution_tolerance"]), maximum_iterations=int(p["maximum_iterations"]),
            krylov_rtol=float(k["relative_tolerance"]), krylov_atol=float(k["absolute_tolerance"]),
            krylov_maximum_iterations=int(k["maximum_iterations"]),
            krylov_restart=min(int(k["restart"]), 30), preconditioner=pc, line_search=_LINE_SEARCH.get(ls, 0), verbose=0,
            linear_solver=lin)

    def solve(self):
        """niters, converged = solver.solve()  (bench1.py:162). In place on problem.w; never raises on non-convergence
        unless error_on_nonconvergence is set (the drivers clear it, bench1.py:86)."""
        pr = self.problem
        if pr.dt is not None:                      # steady problems (nl.py) have no time step
            pr.ctx.dt = float(pr.dt)
        rep = pr.ctx.newton_solve(pr.w.tensor, pr.w0.tensor, opts=self._opts())
        self.last_report = rep
        if not rep["converged"] and self.parameters["snes_solver"]["error_on_nonconvergence"]:
            raise RuntimeError("Newton solver did not converge (reason %d)" % rep["reason"])
        return rep["niters"], bool(rep["converged"])


def total_free_energy(problem, w=None):
    """df.assemble(f_chem*dx + kappa/2*inner(grad(c), grad(c))*dx) and its bench2/3/6 variants
    (bench1.py:124-125, bench2.py:187-194, bench3.py:162-167, bench6.py:158-165)."""
    return problem.ctx.energy((w or problem.w).tensor)[0]


def total_solute(problem, w=None):
    """df.assemble(c*dx) (bench1.py:121-122); bench3: solid_fraction = assemble(0.5*(phi+1)*dx)/(Lx*Ly) (bench3.py:169-170)."""
    return problem.ctx.energy((w or problem.w).tensor)[1]


def diagnostics(problem, w=None):
    """Both reductions from one fused kernel launch."""
    return problem.ctx.energy((w or problem.w).tensor)


def write_csv(path, rows, header):
    """np.savetxt(..., fmt='%1.10f', header=..., delimiter=',', comments='')  (bench1.py:210-217): byte-compatible CSV."""
    np.savetxt(path, np.array(rows), fmt="%1.10f", header=header, delimiter=",", comments="")
