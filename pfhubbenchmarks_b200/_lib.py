"""ctypes binding of libpfh_b200.so (the C ABI declared in include/pfh_b200.h).

There is no CPU fallback: if the CUDA library is missing or a call fails, this raises.
torch is used only to own device memory (tensors -> data_ptr) by the callers of this module.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpfh_b200.so")

MESH_RIGHT2D, MESH_CROSSED2D, MESH_KUHN3D = 1, 2, 3
PROBLEM_CH, PROBLEM_BENCH2, PROBLEM_BENCH3, PROBLEM_BENCH6, PROBLEM_NLDIFF, PROBLEM_USER = 1, 2, 3, 6, 7, 100
MAX_PARAMS, NCCL_ID_BYTES, N_TIMERS = 16, 128, 16


class Config(C.Structure):
    _fields_ = [("mesh_kind", C.c_int), ("problem", C.c_int), ("ncell", C.c_int * 3), ("length", C.c_double * 3),
                ("n_params", C.c_int), ("params", C.c_double * MAX_PARAMS), ("device", C.c_int),
                ("rank", C.c_int), ("nranks", C.c_int), ("z_begin", C.c_int), ("z_end", C.c_int),
                ("nccl_unique_id", C.c_void_p), ("stream", C.c_void_p), ("krylov_basis", C.c_int)]


class SolverOpts(C.Structure):
    _fields_ = [("absolute_tolerance", C.c_double), ("relative_tolerance", C.c_double),
                ("solution_tolerance", C.c_double), ("maximum_iterations", C.c_int), ("line_search", C.c_int),
                ("krylov_rtol", C.c_double), ("krylov_atol", C.c_double), ("krylov_maximum_iterations", C.c_int),
                ("krylov_restart", C.c_int), ("preconditioner", C.c_int), ("pc_variant", C.c_int),
                ("pc_sweeps_mass", C.c_int), ("pc_sweeps_shifted", C.c_int), ("verbose", C.c_int),
                ("linear_solver", C.c_int)]


class SolveReport(C.Structure):
    _fields_ = [("niters", C.c_int), ("converged", C.c_int), ("reason", C.c_int), ("fnorm0", C.c_double),
                ("fnorm", C.c_double), ("krylov_its", C.c_int), ("residual_evals", C.c_int),
                ("jac_applies", C.c_int), ("pc_applies", C.c_int), ("last_krylov_relres", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


ABI_VERSION = 2      # PFH_ABI_VERSION of include/pfh_b200.h (2: pfh_solver_opts.linear_solver)
_P = C.c_void_p
# name -> (restype, argtypes); every symbol include/pfh_b200.h declares
SIGNATURES = {
    "pfh_default_solver_opts": (None, [C.POINTER(SolverOpts)]),
    "pfh_abi_version": (C.c_int, []),
    "pfh_last_error": (C.c_char_p, []),
    "pfh_nccl_unique_id": (C.c_int, [_P]),
    "pfh_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "pfh_destroy": (C.c_int, [_P]),
    "pfh_n_fields": (C.c_int, [_P]),
    "pfh_n_owned": (C.c_longlong, [_P]),
    "pfh_n_global": (C.c_longlong, [_P]),
    "pfh_set_dt": (C.c_int, [_P, C.c_double]),
    "pfh_get_dt": (C.c_double, [_P]),
    "pfh_initial_condition_dev": (C.c_int, [_P, C.POINTER(C.c_double), C.c_int, _P]),
    "pfh_residual_dev": (C.c_int, [_P, _P, _P, _P]),
    "pfh_jac_setup_dev": (C.c_int, [_P, _P]),
    "pfh_jac_apply_dev": (C.c_int, [_P, _P, _P]),
    "pfh_precond_setup": (C.c_int, [_P, C.POINTER(SolverOpts)]),
    "pfh_precond_apply_dev": (C.c_int, [_P, _P, _P]),
    "pfh_newton_solve_dev": (C.c_int, [_P, _P, _P, C.POINTER(SolverOpts), C.POINTER(SolveReport)]),
    "pfh_energy_dev": (C.c_int, [_P, _P, C.POINTER(C.c_double)]),
    "pfh_eval_points_dev": (C.c_int, [_P, _P, C.c_int, _P, C.c_longlong, _P]),
    "pfh_bilinear_dev": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double)]),
    "pfh_step_host": (C.c_int, [_P, _P, _P, C.c_double, C.POINTER(SolverOpts), C.POINTER(SolveReport),
                                C.POINTER(C.c_double)]),
    "pfh_residual_host": (C.c_int, [_P, _P, _P, _P]),
    "pfh_jac_apply_host": (C.c_int, [_P, _P, _P, _P]),
    "pfh_energy_host": (C.c_int, [_P, _P, C.POINTER(C.c_double)]),
    "pfh_launch_count": (C.c_longlong, [_P]),
    "pfh_set_option": (C.c_int, [_P, C.c_char_p, C.c_double]),
    "pfh_timers": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.POINTER(C.c_double), C.c_int]),
    "pfh_timer_name": (C.c_char_p, [C.c_int]),
    "pfh_stream": (_P, [_P]),
    "pfh_synchronize": (C.c_int, [_P]),
}

_lib = None


class PfhError(RuntimeError):
    pass


def load():
    """Load the CUDA library; raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PfhError(
            "libpfh_b200.so is not built (%s). Run `python -m pfhubbenchmarks_b200.build`; "
            "there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    if lib.pfh_abi_version() != ABI_VERSION:
        raise PfhError("ABI version mismatch")
    _lib = lib
    return lib


_user_libs = {}


def load_user(path):
    """Load the copy of the library built for one generated model (build.build_user); same ABI, same checks."""
    if path in _user_libs:
        return _user_libs[path]
    if not os.path.exists(path):
        raise PfhError("generated library %s is missing; there is no CPU fallback" % path)
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.pfh_abi_version() != ABI_VERSION:
        raise PfhError("ABI version mismatch")
    _user_libs[path] = lib
    return lib


def check(rc, lib=None):
    if rc != 0:
        raise PfhError((lib or load()).pfh_last_error().decode())


def default_opts(**kw):
    o = SolverOpts()
    load().pfh_default_solver_opts(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise KeyError(k)
        setattr(o, k, v)
    return o
