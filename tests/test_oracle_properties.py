"""Size-independent properties that pin the oracle's discretisation where the reference commits no output (3-D): the
degree-3 quadrature rules integrate cubics exactly, the P1 operators are the Galerkin matrices of the Laplacian and of
the identity (exact on linears, second-order convergent on a manufactured solution), the Kuhn split fills the box with
positively sized tets of equal volume, and a backward-Euler Cahn-Hilliard step conserves mass and dissipates energy.
(2-D is additionally pinned against the reference's files, tests/test_oracle_golden.py.)"""
import itertools

import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import fem, mesh as omesh, problems, stepper


def _meshes():
    return [omesh.crossed2d(5, 4, 2.5, 3.0), omesh.kuhn3d(4, 3, 5, 2.0, 1.5, 3.0)]


@pytest.mark.parametrize("m", _meshes(), ids=["crossed2d", "kuhn3d"])
def test_mesh_partitions_the_domain(m):
    V = fem.P1Space(m)
    vol = float(np.prod(m.lengths))
    assert np.all(V.vol > 0) and abs(V.vol.sum() - vol) < 1e-12 * vol
    assert np.allclose(V.vol, V.vol[0])                              # congruent-volume simplices (4 per square, 6 per cube)
    ncell = int(np.prod(m.shape)) * (4 if m.dim == 2 else 6)
    assert m.cells.shape == (ncell, m.dim + 1) and len({tuple(sorted(c)) for c in m.cells.tolist()}) == ncell


@pytest.mark.parametrize("m", _meshes(), ids=["crossed2d", "kuhn3d"])
def test_quadrature_is_exact_for_cubics(m):
    """The rule behind `quadrature_degree = 3` (bench1.py:14): every monomial of total degree <= 3 is integrated exactly
    over the box; a quartic is not (the rule is not accidentally of higher degree)."""
    V = fem.P1Space(m)
    X, C = m.coords, m.cells
    lam = fem._bary(fem.TRI6_PTS if m.dim == 2 else fem.TET5_PTS)                  # (nq, d+1)
    xq = np.einsum("qa,cad->cqd", lam, X[C]).reshape(-1, m.dim)                     # physical quadrature points
    L = m.lengths
    for powers in itertools.product(range(4), repeat=m.dim):
        if sum(powers) > 3:
            continue
        exact = np.prod([L[d] ** (p + 1) / (p + 1) for d, p in enumerate(powers)])
        got = V.integrate(np.prod([xq[:, d] ** p for d, p in enumerate(powers)], axis=0))
        assert abs(got - exact) < 1e-12 * max(1.0, abs(exact)), powers
    quartic = V.integrate(xq[:, 0] ** 4)
    assert abs(quartic - L[0] ** 5 / 5 * np.prod(L[1:])) > 1e-6


@pytest.mark.parametrize("m", _meshes(), ids=["crossed2d", "kuhn3d"])
def test_p1_operators_are_galerkin_matrices(m):
    V = fem.P1Space(m)
    X = m.coords
    one = np.ones(m.nv)
    lin = 0.3 + X @ np.arange(1, m.dim + 1)
    vol = float(np.prod(m.lengths))
    assert abs(one @ (V.M @ one) - vol) < 1e-12 * vol                               # int 1 = |domain|
    assert abs(one @ (V.M @ lin) - V.integrate(V.at_quad(lin[None])[0])) < 1e-11    # consistent mass = quadrature on P1 x P1
    assert np.max(np.abs(V.K @ one)) < 1e-12                                        # constants are in the kernel of K
    grad2 = float(np.sum(np.arange(1, m.dim + 1) ** 2))
    assert abs(lin @ (V.K @ lin) - grad2 * vol) < 1e-10 * grad2 * vol               # int |grad u|^2 exact for linear u
    interior = np.all((X > 1e-12) & (X < np.asarray(m.lengths) - 1e-12), axis=1)
    assert np.max(np.abs((V.K @ lin)[interior])) < 1e-11                            # discrete Laplacian of a linear function
    assert abs((V.M - V.M.T)).max() < 1e-15 and abs((V.K - V.K.T)).max() < 1e-13


@pytest.mark.parametrize("dim", [2, 3])
def test_second_order_convergence_on_a_manufactured_solution(dim):
    """(K + M) u = M f with u = prod cos(pi x_d / L_d) (natural no-flux boundary): the P1 error in L2 drops by ~4 when h
    is halved -- a consistency check of the stiffness matrix on the crossed / Kuhn splits that needs no reference data."""
    errs = []
    for n in ((8, 16) if dim == 2 else (6, 12)):
        m = omesh.crossed2d(n, n, 1.0, 1.0) if dim == 2 else omesh.kuhn3d(n, n, n, 1.0, 1.0, 1.0)
        V = fem.P1Space(m)
        u_exact = np.prod(np.cos(np.pi * m.coords), axis=1)
        f = (1.0 + dim * np.pi ** 2) * u_exact
        u = spla.spsolve((V.K + V.M).tocsc(), V.M @ f)
        e = u - u_exact
        errs.append(np.sqrt(e @ (V.M @ e)))
    assert 3.3 < errs[0] / errs[1] < 4.7, errs


def test_3d_cahn_hilliard_step_conserves_mass_and_dissipates_energy():
    m = omesh.kuhn3d(6, 5, 4, 12.0, 10.0, 8.0)
    P = problems.CahnHilliard()
    D = stepper.Discretization(m, P)
    u = P.ic(m)
    rng = np.random.default_rng(0)
    u[0] += 0.02 * rng.standard_normal(m.nv)                         # z-dependence (the reference IC has none)
    E0, M0 = D.energy(u), D.mass(u)
    for dt in (0.1, 0.2, 0.4):
        u, n, ok = D.solve_step(u, u, dt, atol=1e-11)
        assert ok
        E1, M1 = D.energy(u), D.mass(u)
        assert E1 < E0 and abs(M1 - M0) < 1e-11 * abs(M0)
        E0 = E1


def test_symbol_bound_of_the_shifted_operator_matches_the_assembled_spectrum():
    """The Chebyshev interval of the product's shifted solves takes its upper end from the symbol of D^-1 (M + alpha K) on
    the interior rows of the Kuhn mesh (csrc/pfh_solver.cpp symbol_bounds3; oracle.krylov.symbol_lmax is its CPU statement).
    Against the assembled operator: the symbol's maximum is the largest eigenvalue to a fraction of a per cent (boundary
    rows), so 1.015 x symbol is a safe upper bound, where the generic bound (d+2)/2 = 2.5 is 30 % too wide."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    from oracle import krylov
    m = omesh.kuhn3d(9, 8, 7, 18.0, 16.0, 14.0)
    D = stepper.Discretization(m, problems.CahnHilliard())
    for alpha in (1.0, 4.0, 8.0):
        A = (D.V.M + alpha * D.V.K).tocsr()
        S = sp.diags(1.0 / np.sqrt(A.diagonal()))
        lmax = spl.eigsh(S @ A @ S, k=1, which="LA", return_eigenvectors=False)[0]
        sym = krylov.symbol_lmax(A, m)
        assert abs(sym / lmax - 1.0) < 5e-3 and 1.015 * sym > lmax and lmax < 2.0 < 2.5, (alpha, sym, lmax)
