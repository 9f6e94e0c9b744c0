"""Test-only CPU build of the product's per-vertex operator code (see emul.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_LIB = os.path.join(_HERE, "libpfh_emul.so")
_SRC = [os.path.join(_HERE, "emul.cpp"), os.path.join(_ROOT, "pfhubbenchmarks_b200", "csrc", "pfh_tables.cpp")]
_DEPS = _SRC + [os.path.join(_ROOT, "pfhubbenchmarks_b200", "csrc", f)
                for f in ("pfh_vertex.cuh", "pfh_elem.cuh", "pfh_problems.cuh", "pfh_internal.h", "pfh_sample.cuh", "pfh_nl.cuh", "pfh_op2_tiled.cuh", "pfh_mg2.cuh", "pfh_tables.cpp")]


def build(user_header=None):
    """user_header: path of a generated model (codegen.UserModel.source written to a file): a separate library that
    also dispatches PFH_PROBLEM_USER to it."""
    lib, extra, deps = _LIB, [], _DEPS
    if user_header is not None:
        tag = os.path.basename(user_header).rsplit(".", 1)[0]
        lib, extra, deps = os.path.join(_HERE, "libpfh_emul_%s.so" % tag), ['-DPFH_USER_PROBLEM_HEADER="%s"' % user_header], _DEPS + [user_header]
    if os.path.exists(lib) and all(os.path.getmtime(lib) >= os.path.getmtime(s) for s in deps):
        return lib
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    # -Bsymbolic: the product library may be loaded RTLD_GLOBAL in the same process and exports functions of the same names
    subprocess.check_call(["g++", "-O2", "-fopenmp", "-std=c++17", "-shared", "-fPIC", "-Wl,-Bsymbolic", "-I" + cuda_inc, "-o", lib] + extra + _SRC)
    return lib


_P = ctypes.POINTER(ctypes.c_double)


def _p(a):
    return a.ctypes.data_as(_P) if a is not None else None


def run(mesh_kind, problem, ncell, length, params, dt, mode, U=None, X=None, U0=None, icp=None, nf=2,
        z0=0, nzl=-1, glo=None, ghi=None, user_header=None):
    """mode: 0 residual, 1 jac_apply, 2 energy -> [E, M], 3 scalar op -> [Av, diag], 4 initial condition,
    5 nonlinear Jacobian block via the low-flop cube routine (3-D quartic problems), 6 its residual counterpart,
    7 the single-precision variant of 5, 8 / 9 residual / jac_apply through the phases of the square-centric 2-D kernel
    (csrc/pfh_op2_tiled.cuh; returns None when the problem's tile does not fit and the product keeps the gather kernel)."""
    lib = ctypes.CDLL(build(user_header))
    nc = (ctypes.c_int * 3)(*(list(ncell) + [1] * (3 - len(ncell))))
    ln = (ctypes.c_double * 3)(*(list(length) + [1.0] * (3 - len(length))))
    pr = np.ascontiguousarray(params if len(params) else [0.0], dtype=np.float64)
    if mesh_kind == 3:
        nzv = ncell[2] + 1
        n = (ncell[0] + 1) * (ncell[1] + 1) * (nzv if nzl < 0 else nzl)
    else:
        n = (ncell[0] + 1) * (ncell[1] + 1) + ncell[0] * ncell[1]
    arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (U, X, U0, glo, ghi)]
    out = np.zeros({0: nf * n, 1: nf * n, 2: 2, 3: 2 * n, 4: nf * n, 5: n, 6: n, 7: n, 8: nf * n, 9: nf * n}[mode])
    icp = np.ascontiguousarray(icp if icp is not None else [0.0], dtype=np.float64)
    rc = lib.emul_run(mesh_kind, problem, nc, ln, _p(pr), len(pr), ctypes.c_double(dt), mode, z0, nzl,
                      _p(arrs[0]), _p(arrs[1]), _p(arrs[2]), _p(arrs[3]), _p(arrs[4]), _p(out), _p(icp))
    if mode in (8, 9) and rc == -3:
        return None
    assert rc == 0, rc
    if mode in (0, 1, 4, 8, 9):
        return out.reshape(nf, n)
    if mode == 3:
        return out.reshape(2, n)
    if mode in (5, 6, 7):
        return out
    return out


def run_nl(ncell, length, params, mode, U, X=None, square=False):
    """dolfin/nl.py's operators (csrc/pfh_nl.cuh) on the CPU: mode 0 residual F(U), 1 J(U) X, 2 (A X, diag A) with A the SPD
    part of -J.  square=True: square-centric scatter (the tiled kernel's order of work).  Returns (values, diagonal)."""
    lib = ctypes.CDLL(build())
    nc = (ctypes.c_int * 3)(*(list(ncell) + [1]))
    ln = (ctypes.c_double * 3)(*(list(length) + [1.0]))
    pr = np.ascontiguousarray(params, dtype=np.float64)
    U = np.ascontiguousarray(U, dtype=np.float64)
    X = None if X is None else np.ascontiguousarray(X, dtype=np.float64)
    n = (ncell[0] + 1) * (ncell[1] + 1)
    out = np.zeros(2 * n)
    lib.emul_nl.argtypes = [ctypes.c_void_p, ctypes.c_void_p, _P, ctypes.c_int, ctypes.c_int, _P, _P, _P]
    rc = lib.emul_nl(nc, ln, _p(pr), mode, int(square), _p(U), _p(X), _p(out))
    assert rc == 0, rc
    return out[:n], out[n:]


def mg_transfer(ncell, length, mode, v):
    """csrc/pfh_mg2.cuh on the CPU: mode 0 restriction P^T v (v on the crossed mesh `ncell`, result on ncell / 2), mode 1
    prolongation P v (v on ncell / 2).  Also returns the number of coarse levels pfh_mg_build_levels finds for `ncell`."""
    lib = ctypes.CDLL(build())
    nc = (ctypes.c_int * 3)(*(list(ncell) + [1]))
    ln = (ctypes.c_double * 3)(*(list(length) + [1.0]))
    nf = (ncell[0] + 1) * (ncell[1] + 1) + ncell[0] * ncell[1]
    ncs = (ncell[0] // 2 + 1) * (ncell[1] // 2 + 1) + (ncell[0] // 2) * (ncell[1] // 2)
    v = np.ascontiguousarray(v, dtype=np.float64)
    out = np.zeros(ncs if mode == 0 else nf)
    lib.emul_mg_transfer.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, _P, _P]
    nlev = lib.emul_mg_transfer(nc, ln, mode, _p(v), _p(out))
    assert nlev > 0, nlev
    return out, nlev


def sample(mesh_kind, ncell, length, u, pts, z0=0, nzl=-1, glo=None, ghi=None):
    """csrc/pfh_sample.cuh on the CPU: values at pts ((npts, dim)); +inf outside the mesh, NaN where a slab does not own the cube."""
    lib = ctypes.CDLL(build())
    nc = (ctypes.c_int * 3)(*(list(ncell) + [1] * (3 - len(ncell))))
    ln = (ctypes.c_double * 3)(*(list(length) + [1.0] * (3 - len(length))))
    arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (u, glo, ghi)]
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    out = np.zeros(len(pts))
    lib.emul_sample.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _P, _P, _P, _P,
                                ctypes.c_longlong, _P]
    rc = lib.emul_sample(mesh_kind, nc, ln, z0, nzl, _p(arrs[0]), _p(arrs[1]), _p(arrs[2]), _p(pts), len(pts), _p(out))
    assert rc == 0, rc
    return out


_SOLVER_LIB = os.path.join(_HERE, "libpfh_emul_solver.so")
_SOLVER_SRC = [os.path.join(_HERE, "emul_solver.cpp"), os.path.join(_ROOT, "pfhubbenchmarks_b200", "csrc", "pfh_tables.cpp")]


def build_solver(user_header=None):
    """CPU build of the product's host solver (csrc/pfh_solver.cpp) over loop versions of the kernel launchers;
    user_header: additionally for one generated model (codegen.py), in a library of its own."""
    lib, extra = _SOLVER_LIB, []
    deps = _DEPS + _SOLVER_SRC + [os.path.join(_ROOT, "pfhubbenchmarks_b200", "csrc", "pfh_solver.cpp")]
    if user_header is not None:
        tag = os.path.basename(user_header).rsplit(".", 1)[0]
        lib, extra, deps = os.path.join(_HERE, "libpfh_emul_solver_%s.so" % tag), ['-DPFH_USER_PROBLEM_HEADER="%s"' % user_header], deps + [user_header]
    if os.path.exists(lib) and all(os.path.getmtime(lib) >= os.path.getmtime(s) for s in deps):
        return lib
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.check_call(["g++", "-O2", "-fopenmp", "-std=c++17", "-shared", "-fPIC", "-Wl,-Bsymbolic", "-I" + cuda_inc, "-o", lib] + extra + _SOLVER_SRC)
    return lib


def mg_solve(ncell, length, cycles, rhs):
    """x ~ K_D^-1 rhs on the crossed mesh by `cycles` V-cycles of the product's multigrid (mg_solve in csrc/pfh_solver.cpp),
    K_D = stiffness matrix with identity rows on the corner vertices of x = 0 and x = Lx."""
    lib = ctypes.CDLL(build_solver())
    nc = (ctypes.c_int * 3)(*(list(ncell) + [1]))
    ln = (ctypes.c_double * 3)(*(list(length) + [1.0]))
    rhs = np.ascontiguousarray(rhs, dtype=np.float64)
    x = np.zeros_like(rhs)
    lib.emul_mg_solve.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, _P, _P]
    rc = lib.emul_mg_solve(nc, ln, int(cycles), _p(rhs), _p(x))
    assert rc > 0, rc
    return x


def newton_step(mesh_kind, problem, ncell, length, params, dt, w, w0, basis_precision=0, user_header=None, mg_cycles=1, **opts):
    """One backward-Euler step with the product's Newton-Krylov solver on the CPU.  Returns (w_new (nf, n), report dict);
    `opts` are pfh_solver_opts fields (pfhubbenchmarks_b200._lib.default_opts)."""
    from pfhubbenchmarks_b200 import _lib
    lib = ctypes.CDLL(build_solver(user_header))
    nc = (ctypes.c_int * 3)(*(list(ncell) + [1] * (3 - len(ncell))))
    ln = (ctypes.c_double * 3)(*(list(length) + [1.0] * (3 - len(length))))
    pr = np.ascontiguousarray(params if len(params) else [0.0], dtype=np.float64)
    o = _lib.default_opts(**opts)
    w = np.ascontiguousarray(w, dtype=np.float64).copy()
    w0 = np.ascontiguousarray(w0, dtype=np.float64)
    rep = np.zeros(8)
    lib.emul_newton_step.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, _P, ctypes.c_int, ctypes.c_double,
                                     ctypes.c_void_p, ctypes.c_int, _P, _P, _P, ctypes.c_int]
    rc = lib.emul_newton_step(mesh_kind, problem, nc, ln, _p(pr), len(params), float(dt), ctypes.byref(o), int(basis_precision),
                              _p(w.reshape(-1)), _p(w0.reshape(-1)), _p(rep), int(mg_cycles))
    assert rc == 0, rc
    keys = ("niters", "converged", "reason", "fnorm0", "fnorm", "krylov_its", "residual_evals", "pc_applies")
    return w, dict(zip(keys, rep))
