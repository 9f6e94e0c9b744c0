"""The product's per-vertex operator code (csrc/pfh_vertex.cuh etc.), compiled for the CPU by
tests/emul, against the oracle's assembled operators.  CPU only; exercises stencil tables,
quadrature, hand-written f'/f'', Dirichlet rows and slab ghost planes."""
import numpy as np
import pytest

import emul
from helpers import discretization, rel_max
from oracle import mesh as omesh, problems, stepper

CASES = [
    ("bench1", 2, 1, (7, 5), (14.0, 10.0), [0.3, 0.7, 5.0, 2.0, 5.0], [0.5, 0.05]),
    ("bench1", 2, 1, (6, 4), (9.0, 10.0), [0.3, 0.7, 5.0, 2.0, 5.0], [0.5, 0.05]),       # hx != hy
    ("b13d", 3, 1, (5, 4, 3), (10.0, 8.0, 6.0), [0.3, 0.7, 5.0, 2.0, 5.0], [0.5, 0.05]),
    ("b13d", 3, 1, (3, 4, 5), (4.0, 8.0, 7.0), [0.3, 0.7, 5.0, 2.0, 5.0], [0.5, 0.05]),   # anisotropic cells
    ("bench2", 2, 2, (6, 5), (12.0, 10.0), [0.3, 0.7, 2.0, 3.0, 3.0, 5.0, 1.0, 5.0, 5.0], [0.5, 0.05, 0.1, 1.5]),
    ("bench3", 2, 3, (6, 5), (12.0, 10.0), [1.0, 1.0, 10.0, 10.0 / 0.6267, 120.0], [-0.3, 8.0, 1.0, 1.0, -1.0]),
    ("bench6", 2, 6, (6, 5), (12.0, 10.0), [0.3, 0.7, 5.0, 2.0, 5.0, 0.09, 90.0], [0.5, 0.04]),
]


@pytest.mark.parametrize("name,mk,pid,ncell,length,params,icp", CASES)
def test_operators_match_oracle(name, mk, pid, ncell, length, params, icp):
    rng = np.random.default_rng(1)
    D = discretization(name, ncell, length)
    P, m, nf = D.prob, D.mesh, D.nf
    U = P.ic(m) + 0.1 * rng.standard_normal((nf, m.nv))
    U0 = P.ic(m) + 0.1 * rng.standard_normal((nf, m.nv))
    X = rng.standard_normal((nf, m.nv))
    dt = 0.37
    args = (mk, pid, list(ncell), list(length), params, dt)
    assert rel_max(emul.run(*args, 0, U=U, U0=U0, nf=nf), D.residual(U, U0, dt)) < 1e-13
    assert rel_max(emul.run(*args, 1, U=U, X=X, nf=nf), (D.jacobian(U, dt) @ X.ravel()).reshape(nf, -1)) < 1e-13
    EM = emul.run(*args, 2, U=U, nf=nf)
    assert abs(EM[0] / D.energy(U) - 1) < 1e-13 and abs(EM[1] / D.mass(U) - 1) < 1e-13
    assert rel_max(emul.run(*args, 4, icp=icp, nf=nf), P.ic(m)) < 1e-14
    A = 1.3 * D.V.M + 0.7 * D.V.K
    got = emul.run(*args, 3, U=U, icp=[1.3, 0.7], nf=nf)
    assert rel_max(got[0], A @ U[0]) < 1e-13 and rel_max(got[1], A.diagonal()) < 1e-14


def test_jacobian_is_derivative_of_residual():
    """J.v against a centred finite difference of R (SURVEY section 7 gate: <= 1e-6)."""
    rng = np.random.default_rng(2)
    for (name, mk, pid, ncell, length, params, _) in CASES:
        D = discretization(name, ncell, length)
        nf, nv = D.nf, D.nv
        U = D.prob.ic(D.mesh) + 0.05 * rng.standard_normal((nf, nv))
        X = rng.standard_normal((nf, nv))
        args = (mk, pid, list(ncell), list(length), params, 0.2)
        eps = 1e-6
        fd = (emul.run(*args, 0, U=U + eps * X, U0=U, nf=nf) - emul.run(*args, 0, U=U - eps * X, U0=U, nf=nf)) / (2 * eps)
        assert rel_max(emul.run(*args, 1, U=U, X=X, nf=nf), fd) < 1e-7


def test_slab_with_ghost_planes_equals_global():
    """z-slab ownership + one ghost plane per side reproduces the single-domain operator."""
    rng = np.random.default_rng(3)
    ncell, length = (4, 3, 7), (8.0, 6.0, 14.0)
    params = [0.3, 0.7, 5.0, 2.0, 5.0]
    D = discretization("b13d", ncell, length)
    nf, nv = 2, D.nv
    nxy = 5 * 4
    U = D.prob.ic(D.mesh) + 0.1 * rng.standard_normal((nf, nv))
    U0 = D.prob.ic(D.mesh) + 0.1 * rng.standard_normal((nf, nv))
    X = rng.standard_normal((nf, nv))
    Rg = D.residual(U, U0, 0.3)
    Yg = (D.jacobian(U, 0.3) @ X.ravel()).reshape(nf, -1)
    Etot = 0.0
    for (z0, z1) in [(0, 3), (3, 5), (5, 8)]:
        sl = slice(z0 * nxy, z1 * nxy)

        def ghosts(zp):
            g = np.zeros((3, nf, nxy))
            if 0 <= zp < 8:
                for i, A in enumerate((U, X, U0)):
                    g[i] = A[:, zp * nxy:(zp + 1) * nxy]
            return g
        kw = dict(z0=z0, nzl=z1 - z0, glo=ghosts(z0 - 1), ghi=ghosts(z1), nf=nf)
        R = emul.run(3, 1, list(ncell), list(length), params, 0.3, 0, U=U[:, sl], U0=U0[:, sl], **kw)
        Y = emul.run(3, 1, list(ncell), list(length), params, 0.3, 1, U=U[:, sl], X=X[:, sl], **kw)
        assert np.abs(R - Rg[:, sl]).max() < 1e-12 * np.abs(Rg).max()
        assert np.abs(Y - Yg[:, sl]).max() < 1e-12 * np.abs(Yg).max()
        Etot += emul.run(3, 1, list(ncell), list(length), params, 0.3, 2, U=U[:, sl], **kw)[0]
    assert abs(Etot / D.energy(U) - 1) < 1e-13


def test_low_flop_cube_jacobian_routine():
    """elem::cube_jac_quartic (folded-constant form used by the TMA J.v kernel) == int phi_i g'(c_h) x_h of the oracle."""
    rng = np.random.default_rng(4)
    ncell, length = (4, 5, 3), (8.0, 7.5, 6.0)
    D = discretization("b13d", ncell, length)
    U = D.prob.ic(D.mesh) + 0.1 * rng.standard_normal((2, D.nv))
    X = rng.standard_normal((2, D.nv))
    got = emul.run(3, 1, list(ncell), list(length), [0.3, 0.7, 5.0, 2.0, 5.0], 0.1, 5, U=U, X=X, nf=2)
    dg = D.prob.dg(D.V.at_quad(U))
    ref = D.V.weighted_mass(dg[1, 0]) @ X[0]
    assert rel_max(got, ref) < 1e-13
    # the opt-in single-precision variant (Krylov J.v only): fp32 accuracy, no worse
    got32 = emul.run(3, 1, list(ncell), list(length), [0.3, 0.7, 5.0, 2.0, 5.0], 0.1, 7, U=U, X=X, nf=2)
    assert 1e-9 < rel_max(got32, ref) < 5e-6


def test_low_flop_cube_residual_routine():
    """elem::cube_res_quartic == int phi_i g(c_h), g = -f'(c), of the oracle."""
    rng = np.random.default_rng(5)
    ncell, length = (4, 5, 3), (8.0, 7.5, 6.0)
    D = discretization("b13d", ncell, length)
    U = D.prob.ic(D.mesh) + 0.1 * rng.standard_normal((2, D.nv))
    got = emul.run(3, 1, list(ncell), list(length), [0.3, 0.7, 5.0, 2.0, 5.0], 0.1, 6, U=U, nf=2)
    ref = D.V.load(D.prob.g(D.V.at_quad(U)))[1]
    assert rel_max(got, ref) < 1e-13


def _probe_points(rng, lengths, n):
    """Random interior points plus the awkward ones: vertices, cell centres, points on faces/diagonals, the far corner."""
    L = np.asarray(lengths, dtype=float)
    pts = [rng.random((n, len(L))) * L, np.zeros((1, len(L))), L[None, :], 0.5 * L[None, :]]
    return np.concatenate(pts)


def test_point_evaluation_matches_cell_search():
    """csrc/pfh_sample.cuh (cell index + ordering of local coordinates) against the oracle's brute-force search over
    all simplices of the same mesh: f(x) of the reference's peval/sample (pfbase.py:50-94)."""
    from oracle import fem
    rng = np.random.default_rng(5)
    for ncell, length in [((7, 5), (14.0, 10.0)), ((4, 6), (3.0, 9.0))]:
        m = omesh.crossed2d(*ncell, *length)
        u = rng.standard_normal(m.nv)
        pts = _probe_points(rng, length, 300)
        hx, hy = length[0] / ncell[0], length[1] / ncell[1]
        pts = np.concatenate([pts, [[2 * hx, 3 * hy], [2.5 * hx, 1.5 * hy], [2.25 * hx, 1.25 * hy], [2.75 * hx, 1.25 * hy], [hx, 0.3 * hy]]])
        got = emul.sample(2, ncell, length, u, pts)
        want = fem.eval_points(m, u, pts)
        assert np.all(np.isfinite(want)) and np.max(np.abs(got - want)) < 1e-12
        # a linear function is reproduced exactly; outside points give +inf
        lin = 0.3 + 1.7 * m.coords[:, 0] - 0.6 * m.coords[:, 1]
        assert np.max(np.abs(emul.sample(2, ncell, length, lin, pts) - (0.3 + 1.7 * pts[:, 0] - 0.6 * pts[:, 1]))) < 1e-12
        out = emul.sample(2, ncell, length, u, [[-0.1, 1.0], [1.0, length[1] + 0.1]])
        assert np.all(np.isinf(out)) and np.all(np.isinf(fem.eval_points(m, u, [[-0.1, 1.0], [1.0, length[1] + 0.1]])))
    for ncell, length in [((5, 4, 3), (10.0, 8.0, 6.0)), ((3, 4, 5), (4.0, 8.0, 7.0))]:
        m = omesh.kuhn3d(*ncell, *length)
        u = rng.standard_normal(m.nv)
        pts = _probe_points(rng, length, 300)
        h = np.asarray(length) / np.asarray(ncell)
        pts = np.concatenate([pts, [h * [1, 2, 1], h * [1.5, 1.5, 1.5], h * [1.25, 1.25, 1.7], h * [2.0, 1.3, 1.3]]])
        got = emul.sample(3, ncell, length, u, pts)
        want = fem.eval_points(m, u, pts)
        assert np.all(np.isfinite(want)) and np.max(np.abs(got - want)) < 1e-12
        # slabs: every point is owned by exactly one slab (the one holding its cube's lower plane) with the same value
        nzv, nxy = ncell[2] + 1, (ncell[0] + 1) * (ncell[1] + 1)
        U = u.reshape(nzv, nxy)
        owners = np.zeros(len(pts), dtype=int)
        for z0, z1 in [(0, 2), (2, 3), (3, nzv)]:
            glo = U[z0 - 1] if z0 > 0 else None
            ghi = U[z1] if z1 < nzv else None
            part = emul.sample(3, ncell, length, U[z0:z1].ravel(), pts, z0=z0, nzl=z1 - z0, glo=glo, ghi=ghi)
            mine = ~np.isnan(part)
            owners += mine
            assert not mine.any() or np.max(np.abs(part[mine] - want[mine])) < 1e-12     # the top plane alone owns no cube
        assert np.all(owners == 1)


# sizes around the 31 x 7 vertex tile of the square-centric kernel: inside one tile, exactly one tile, one vertex more,
# 2 x 2 tiles exactly, ragged and anisotropic
TILED_CASES = [("bench1", 1, (7, 5), (14.0, 10.0)), ("bench1", 1, (30, 6), (60.0, 12.0)), ("bench1", 1, (31, 7), (62.0, 21.0)),
               ("bench1", 1, (61, 13), (61.0, 26.0)), ("bench1", 1, (70, 9), (140.0, 18.0)), ("bench1", 1, (1, 1), (2.0, 2.0)),
               ("bench3", 3, (33, 9), (66.0, 18.0)), ("bench6", 6, (33, 16), (33.0, 16.0)), ("bench2", 2, (6, 5), (12.0, 10.0))]


@pytest.mark.parametrize("name,pid,ncell,length", TILED_CASES)
def test_square_centric_2d_kernel_phases(name, pid, ncell, length):
    """k2_op_tiled (csrc/pfh_op2_tiled.cuh): stage / square / vertex phases run block by block on the CPU against the
    oracle: every vertex and every cell centre is written exactly once (the output starts as NaN) with the assembled
    operator's value; Dirichlet rows of bench6 included.  bench2's six-field tile does not fit: the product falls back."""
    rng = np.random.default_rng(4)
    D = discretization(name, ncell, length)
    P, m, nf = D.prob, D.mesh, D.nf
    params = {c[0]: c[5] for c in CASES}[name]
    if name == "bench3":
        params = params[:4] + [length[0] * length[1]]
    U = P.ic(m) + 0.1 * rng.standard_normal((nf, m.nv))
    U0 = P.ic(m) + 0.1 * rng.standard_normal((nf, m.nv))
    X = rng.standard_normal((nf, m.nv))
    dt = 0.37
    args = (2, pid, list(ncell), list(length), params, dt)
    R = emul.run(*args, 8, U=U, U0=U0, nf=nf)
    Jx = emul.run(*args, 9, U=U, X=X, nf=nf)
    if name == "bench2":
        assert R is None and Jx is None
        return
    assert np.isfinite(R).all() and np.isfinite(Jx).all()
    assert rel_max(R, D.residual(U, U0, dt)) < 1e-13
    assert rel_max(Jx, (D.jacobian(U, dt) @ X.ravel()).reshape(nf, -1)) < 1e-13
    # and against the vertex-centric code the gather kernel runs: same arithmetic, sums in the same order
    assert np.abs(R - emul.run(*args, 0, U=U, U0=U0, nf=nf)).max() <= 1e-14 * np.abs(R).max()
    assert np.abs(Jx - emul.run(*args, 1, U=U, X=X, nf=nf)).max() <= 1e-14 * np.abs(Jx).max()
