"""Shared helpers for the tests: oracle discretisations per benchmark and error norms."""
import functools

import numpy as np

from oracle import krylov, mesh, problems, stepper


def rel_l2(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b)))


def rel_max(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(np.asarray(b)).max())


# (oracle problem factory, engine problem name) per benchmark
def make_problem(name, lengths=None):
    if name in ("bench1", "b13d"):
        return problems.CahnHilliard()
    if name == "bench2":
        return problems.Bench2()
    if name == "bench3":
        return problems.Bench3(Lx=lengths[0], Ly=lengths[1])
    if name == "bench6":
        return problems.Bench6()
    raise KeyError(name)


def make_mesh(ncell, length):
    if len(ncell) == 2:
        return mesh.crossed2d(ncell[0], ncell[1], length[0], length[1])
    return mesh.kuhn3d(ncell[0], ncell[1], ncell[2], length[0], length[1], length[2])


@functools.lru_cache(maxsize=8)
def discretization(name, ncell, length):
    m = make_mesh(ncell, length)
    return stepper.Discretization(m, make_problem(name, length))


REFERENCE_MESH = {  # the meshes the reference drivers build
    "bench1": ((100, 100), (200.0, 200.0)),      # dolfin/bench1.py:21-23
    "bench2": ((100, 100), (200.0, 200.0)),      # dolfin/bench2.py:21-23
    "bench3": ((350, 350), (960.0, 960.0)),      # dolfin/bench3.py:21-23
    "bench6": ((100, 100), (100.0, 100.0)),      # dolfin/bench6.py:21-23
    "b13d": ((50, 50, 50), (100.0, 100.0, 100.0)),   # dolfin/b13d.py:24-26
}


def oracle_step_3d(D, u, u0, dt, rtol=1e-11, atol=1e-10):
    return D.solve_step(u, u0, dt, atol=atol, linear=krylov.make_linear(rtol=rtol, n_m=4, n_a=8))


def lu_once():
    """linear= callback for Discretization.solve_step that factors the FIRST Jacobian of a step and solves the later
    Newton systems by GMRES on the exact Jacobian preconditioned with that factorisation (a handful of iterations to
    1e-13): the same Newton iterates to solver tolerance, one sparse LU per step instead of one per Newton iteration."""
    import scipy.sparse.linalg as spla
    cache = {}

    def linear(disc, J, rhs, u, dt):
        Jc = J.tocsc()
        if "lu" not in cache:
            cache["lu"] = spla.splu(Jc)
            return cache["lu"].solve(rhs)
        M = spla.LinearOperator(Jc.shape, matvec=cache["lu"].solve)
        x, info = spla.gmres(Jc, rhs, M=M, rtol=1e-13, atol=0.0, restart=30, maxiter=10)
        assert info == 0, info
        return x
    return linear
