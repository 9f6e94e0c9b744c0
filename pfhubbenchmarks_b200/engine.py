"""Thin Python handle on a pfh_ctx: one structured mesh + one PFHub problem on one GPU (or one
z-slab of a 3-D mesh per rank).  All arithmetic happens in libpfh_b200.so; torch only owns the
fp64 device buffers that are passed in by pointer.
"""
import ctypes as C

import numpy as np

from . import _lib

PROBLEMS = {
    # name: (problem id, n_fields, default params, default initial-condition params, field names)
    # bench1.py:29-36 / b13d.py:  c_alpha, c_beta, rho_s, kappa, M ; IC c0, epsilon (bench1.py:48-53)
    "bench1": (_lib.PROBLEM_CH, 2, [0.3, 0.7, 5.0, 2.0, 5.0], [0.5, 0.05], ("c", "mu")),
    # bench2.py:34-42: c_alpha, c_beta, rho^2, kappa_c, kappa_eta, M, w, alpha, L ; IC c0, eps, eps_eta, psi (:61-67)
    "bench2": (_lib.PROBLEM_BENCH2, 6, [0.3, 0.7, 2.0, 3.0, 3.0, 5.0, 1.0, 5.0, 5.0], [0.5, 0.05, 0.1, 1.5],
               ("c", "mu", "eta1", "eta2", "eta3", "eta4")),
    # bench3.py:32-38,67: W0, tau0, D, lambda = D tau0 / (0.6267 W0^2), Lx*Ly ; IC Delta, r, w, vin, vout (:53-60)
    "bench3": (_lib.PROBLEM_BENCH3, 2, [1.0, 1.0, 10.0, 10.0 * 1.0 / (0.6267 * 1.0 ** 2), None],
               [-0.3, 8.0, 1.0, 1.0, -1.0], ("U", "phi")),
    # bench6.py:33-39: c_alpha, c_beta, rho, kappa, M, k, epsilon ; IC c0, c1 (:52-57)
    "bench6": (_lib.PROBLEM_BENCH6, 3, [0.3, 0.7, 5.0, 2.0, 5.0, 0.09, 90.0], [0.5, 0.04], ("c", "mu", "phi")),
}
PROBLEMS["b13d"] = PROBLEMS["bench1"]
# nl.py:37,31-35: src, kc, reac of F = int src v - int kc |u| grad u . grad v - int reac u v, Dirichlet value on x = 0, Lx.
# Steady, on the right-diagonal mesh (nl.py:20); no initial-condition expression (nl.py:48 fills the vector with 0.9).
PROBLEMS["nl"] = (_lib.PROBLEM_NLDIFF, 1, [5.0, 1.0, 1.0, 0.9], [0.0], ("u",))


def slab_range(nplanes, rank, nranks):
    """Contiguous z-slab [z_begin, z_end) of `nplanes` vertex planes owned by `rank` (balanced)."""
    base, extra = divmod(nplanes, nranks)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def _ptr(t):
    """Device pointer of a torch tensor.  The library enqueues on its own stream, so work torch has
    queued on its current stream (e.g. the H2D copy that filled `t`) is drained first."""
    import torch
    assert isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous(), \
        "expected a contiguous fp64 CUDA tensor"
    torch.cuda.current_stream(t.device).synchronize()
    return C.c_void_p(t.data_ptr())


class Context:
    def __init__(self, problem, ncell, length, params=None, device=0, rank=0, nranks=1, nccl_id=None,
                 krylov_basis=30, stream=None):
        if isinstance(problem, str):
            self.lib = _lib.load()
            pid, nf, dparams, dic, names = PROBLEMS[problem]
        else:      # codegen.UserModel: its own copy of the library, built (or found in the cache) now
            from . import build
            self.lib = _lib.load_user(build.build_user(problem.source, problem.tag))
            pid, nf, dparams, dic, names = _lib.PROBLEM_USER, problem.nf, [], [0.0], problem.field_names
            problem = "user_" + problem.tag
        self.problem, self.nf, self.field_names, self.default_ic = problem, nf, names, list(dic)
        dim = len(ncell)
        assert dim in (2, 3) and len(length) == dim
        params = list(dparams if params is None else params)
        if pid == _lib.PROBLEM_BENCH3 and params[4] is None:
            params[4] = float(length[0] * length[1])
        cfg = _lib.Config()
        cfg.mesh_kind = _lib.MESH_KUHN3D if dim == 3 else (_lib.MESH_RIGHT2D if pid == _lib.PROBLEM_NLDIFF else _lib.MESH_CROSSED2D)
        self.mesh_kind = cfg.mesh_kind
        cfg.problem = pid
        for d in range(3):
            cfg.ncell[d] = int(ncell[d]) if d < dim else 1
            cfg.length[d] = float(length[d]) if d < dim else 1.0
        cfg.n_params = len(params)
        for i, v in enumerate(params):
            cfg.params[i] = float(v)
        cfg.device, cfg.rank, cfg.nranks = int(device), int(rank), int(nranks)
        if dim == 3:
            cfg.z_begin, cfg.z_end = slab_range(ncell[2] + 1, rank, nranks)
        self.z_range = (cfg.z_begin, cfg.z_end)
        self._idbuf = None
        if nranks > 1:
            assert nccl_id is not None and len(nccl_id) == _lib.NCCL_ID_BYTES
            self._idbuf = C.create_string_buffer(bytes(nccl_id), _lib.NCCL_ID_BYTES)
            cfg.nccl_unique_id = C.cast(self._idbuf, C.c_void_p)
        cfg.stream = stream
        cfg.krylov_basis = int(krylov_basis)
        self.device = int(device)
        self.ncell, self.length, self.dim, self.params = tuple(ncell), tuple(length), dim, params
        h = C.c_void_p()
        self._ck(self.lib.pfh_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.n_owned = int(self.lib.pfh_n_owned(h))
        self.n_global = int(self.lib.pfh_n_global(h))
        self.rank, self.nranks = rank, nranks

    def _ck(self, rc):
        _lib.check(rc, self.lib)

    # ---- lifetime ----
    def close(self):
        if getattr(self, "_h", None):
            self.lib.pfh_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- buffers ----
    def new_vector(self):
        import torch
        return torch.zeros(self.nf * self.n_owned, dtype=torch.float64, device="cuda:%d" % self.device)

    def fields(self, w):
        """Per-field views of a state vector (w.split() in the drivers, bench1.py:188)."""
        return [w[i * self.n_owned:(i + 1) * self.n_owned] for i in range(self.nf)]

    def vertex_coordinates(self):
        """(n_owned, dim) coordinates in the library's vertex order: 2-D corners x-fastest then cell centres
        (DOLFIN 'crossed' numbering); 3-D x-fastest over this rank's z-planes."""
        if self.dim == 2:
            (nx, ny), (lx, ly) = self.ncell, self.length
            x, y = np.meshgrid(np.linspace(0.0, lx, nx + 1), np.linspace(0.0, ly, ny + 1))
            if getattr(self, "mesh_kind", _lib.MESH_CROSSED2D) == _lib.MESH_RIGHT2D:     # grid vertices only (UnitSquareMesh, nl.py:20)
                return np.stack([x.ravel(), y.ravel()], 1)
            xc, yc = np.meshgrid((np.arange(nx) + 0.5) * (lx / nx), (np.arange(ny) + 0.5) * (ly / ny))
            return np.concatenate([np.stack([x.ravel(), y.ravel()], 1), np.stack([xc.ravel(), yc.ravel()], 1)])
        (nx, ny, nz), (lx, ly, lz) = self.ncell, self.length
        zs = np.linspace(0.0, lz, nz + 1)[self.z_range[0]:self.z_range[1]]
        z, y, x = np.meshgrid(zs, np.linspace(0.0, ly, ny + 1), np.linspace(0.0, lx, nx + 1), indexing="ij")
        return np.stack([x.ravel(), y.ravel(), z.ravel()], 1)

    # ---- scalar state ----
    @property
    def dt(self):
        return float(self.lib.pfh_get_dt(self._h))

    @dt.setter
    def dt(self, v):
        self._ck(self.lib.pfh_set_dt(self._h, float(v)))

    # ---- device entry points ----
    def initial_condition(self, ic_params=None, out=None):
        q = np.ascontiguousarray(self.default_ic if ic_params is None else ic_params, dtype=np.float64)
        w = self.new_vector() if out is None else out
        self._ck(self.lib.pfh_initial_condition_dev(self._h, q.ctypes.data_as(C.POINTER(C.c_double)), len(q), _ptr(w)))
        return w

    def residual(self, w, w0, out=None):
        F = self.new_vector() if out is None else out
        self._ck(self.lib.pfh_residual_dev(self._h, _ptr(w), _ptr(w0), _ptr(F)))
        return F

    def jac_setup(self, w):
        self._jac_point = w       # keep the tensor alive: the ctx only stores the pointer
        self._ck(self.lib.pfh_jac_setup_dev(self._h, _ptr(w)))

    def jac_apply(self, x, out=None):
        y = self.new_vector() if out is None else out
        self._ck(self.lib.pfh_jac_apply_dev(self._h, _ptr(x), _ptr(y)))
        return y

    def precond_setup(self, **opts):
        o = _lib.default_opts(**opts)
        self._ck(self.lib.pfh_precond_setup(self._h, C.byref(o)))

    def precond_apply(self, r, out=None):
        z = self.new_vector() if out is None else out
        self._ck(self.lib.pfh_precond_apply_dev(self._h, _ptr(r), _ptr(z)))
        return z

    def newton_solve(self, w, w0, opts=None, **kw):
        """In place on w (initial guess -> solution). Returns the report dict."""
        o = opts if opts is not None else _lib.default_opts(**kw)
        rep = _lib.SolveReport()
        self._ck(self.lib.pfh_newton_solve_dev(self._h, _ptr(w), _ptr(w0), C.byref(o), C.byref(rep)))
        return rep.as_dict()

    def energy(self, w):
        out = (C.c_double * 4)()
        self._ck(self.lib.pfh_energy_dev(self._h, _ptr(w), out))
        return float(out[0]), float(out[1])

    def eval_points(self, w, field, points):
        """P1 interpolant of one field at `points` ((npts, dim) array): f(x) of pfbase.py:50-94; +inf outside the mesh."""
        import torch
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, self.dim)
        dev = torch.device("cuda", self.device)
        d_pts = torch.from_numpy(pts).to(dev)
        out = torch.empty(len(pts), dtype=torch.float64, device=dev)
        if len(pts):
            self._ck(self.lib.pfh_eval_points_dev(self._h, _ptr(w), int(field), _ptr(d_pts.reshape(-1)), len(pts), _ptr(out)))
            self.synchronize()               # the library ran on its own stream
        return out.cpu().numpy()

    def bilinear(self, w, field_a, field_b, am, ak):
        """u_a^T (am M + ak K) u_b: one piece of a df.assemble functional (bench2.py:190-192)."""
        out = C.c_double()
        self._ck(self.lib.pfh_bilinear_dev(self._h, _ptr(w), int(field_a), int(field_b), float(am), float(ak), C.byref(out)))
        return float(out.value)

    # ---- host-buffer entry points (numpy fp64, C-contiguous) ----
    @staticmethod
    def _np(a):
        assert isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous
        return C.c_void_p(a.ctypes.data)

    def step_host(self, w, w0, dt, opts=None, **kw):
        o = opts if opts is not None else _lib.default_opts(**kw)
        rep = _lib.SolveReport()
        diag = (C.c_double * 4)()
        self._ck(self.lib.pfh_step_host(self._h, self._np(w), self._np(w0), float(dt), C.byref(o), C.byref(rep), diag))
        d = rep.as_dict()
        d["energy"], d["mass"] = float(diag[0]), float(diag[1])
        return d

    def residual_host(self, w, w0):
        F = np.empty_like(w)
        self._ck(self.lib.pfh_residual_host(self._h, self._np(w), self._np(w0), self._np(F)))
        return F

    def jac_apply_host(self, w, x):
        y = np.empty_like(w)
        self._ck(self.lib.pfh_jac_apply_host(self._h, self._np(w), self._np(x), self._np(y)))
        return y

    def energy_host(self, w):
        out = (C.c_double * 4)()
        self._ck(self.lib.pfh_energy_host(self._h, self._np(w), out))
        return float(out[0]), float(out[1])

    # ---- instrumentation ----
    def set_option(self, key, value):
        self._ck(self.lib.pfh_set_option(self._h, key.encode(), float(value)))

    def launch_count(self):
        return int(self.lib.pfh_launch_count(self._h))

    def timers(self, reset=False):
        ms = (C.c_double * _lib.N_TIMERS)()
        n = (C.c_longlong * _lib.N_TIMERS)()
        nbytes = (C.c_double * _lib.N_TIMERS)()
        self._ck(self.lib.pfh_timers(self._h, ms, n, nbytes, 1 if reset else 0))
        return {self.lib.pfh_timer_name(i).decode(): {"ms": float(ms[i]), "launches": int(n[i]), "bytes": float(nbytes[i])}
                for i in range(_lib.N_TIMERS)}

    def synchronize(self):
        self._ck(self.lib.pfh_synchronize(self._h))


def nccl_unique_id():
    buf = C.create_string_buffer(_lib.NCCL_ID_BYTES)
    _lib.check(_lib.load().pfh_nccl_unique_id(buf))
    return bytes(buf.raw)
