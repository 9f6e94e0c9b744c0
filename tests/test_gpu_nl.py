"""GPU tests of the nl.py path (dolfin/nl.py: steady quasi-linear diffusion on the right-diagonal UnitSquareMesh), through
the C ABI: the square-centric tiled kernels and the vertex-centric gather kernels of csrc/pfh_nl.cu against the oracle
(oracle/nl.py), the Newton-Krylov solve with the script's settings, the driver, and full-size properties."""
import numpy as np
import pytest

from oracle.nl import NlProblem

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _ctx(ncell, length=(1.0, 1.0), params=None, tiled=True):
    from pfhubbenchmarks_b200 import engine
    ctx = engine.Context("nl", ncell, length, params=params)
    ctx.set_option("nl_tiled", 1.0 if tiled else 0.0)
    return ctx


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).ravel()).cuda()


# sizes around the tiles (31 x 7 vertices per CTA of the tiled kernel, 32 x 8 of the gather kernel): one square, one tile
# exactly, one vertex more than a tile, 2 x 2 tiles exactly, ragged in both directions
CASES = [((60, 75), (1.0, 1.0)), ((1, 1), (1.0, 1.0)), ((30, 6), (1.0, 1.0)), ((31, 7), (2.0, 0.5)), ((32, 8), (1.0, 1.0)),
         ((61, 13), (1.0, 1.0)), ((62, 14), (1.0, 1.0)), ((100, 37), (1.3, 0.7)), ((7, 130), (1.0, 3.0))]


@pytest.mark.parametrize("ncell,length", CASES)
@pytest.mark.parametrize("tiled", [True, False])
def test_operators_match_oracle(ncell, length, tiled):
    """residual, J.v and the preconditioner's sweep operator <= 1e-13 (bit-level: fp64 sums in a different order)."""
    P = NlProblem(ncell[0], ncell[1], length[0], length[1])
    ctx = _ctx(ncell, length, tiled=tiled)
    assert ctx.n_owned == P.nv and ctx.nf == 1
    rng = np.random.default_rng(ncell[0] * 1000 + ncell[1])
    u = 0.9 + 0.5 * rng.standard_normal(P.nv)          # changes sign: |u| and sign(u)
    x = rng.standard_normal(P.nv)
    du, dx = dev(u), dev(x)
    Jx_o = P.jacobian(u) @ x
    scale = max(1.0, np.abs(Jx_o).max())
    R = ctx.residual(du, du).cpu().numpy()
    assert np.abs(R - P.residual(u)).max() <= 1e-13 * scale
    ctx.jac_setup(du)
    Jx = ctx.jac_apply(dx).cpu().numpy()
    assert np.abs(Jx - Jx_o).max() <= 1e-13 * scale
    # point Jacobi: z = -D^-1 r on the free vertices, z = r on the Dirichlet vertices (one sweep of the sweep kernel)
    ctx.precond_setup(preconditioner=1)
    z = ctx.precond_apply(dx).cpu().numpy()
    A = P.sym_operator(u)
    want = -x / A.diagonal()
    want[P.bc_idx] = x[P.bc_idx]
    assert np.abs(z - want).max() <= 1e-13 * np.abs(want).max()
    # host-buffer entry points == device path
    assert np.array_equal(ctx.residual_host(u, u), R)
    assert np.array_equal(ctx.jac_apply_host(u, x), Jx)


def test_chebyshev_preconditioner_matches_its_cpu_statement():
    from oracle.krylov import cheb_coeffs
    P = NlProblem(40, 23)
    rng = np.random.default_rng(5)
    u = 0.9 + 0.3 * rng.random(P.nv)
    r = rng.standard_normal(P.nv)
    A = P.sym_operator(u)
    dinv = 1.0 / A.diagonal()
    for n in (12, 5):
        zp, z = np.zeros(P.nv), np.zeros(P.nv)
        for a, b in cheb_coeffs(2.0 / (5.0 * n), 2.0, n):
            zn = z + a * (z - zp) + b * dinv * (-r - A @ z)
            zn[P.bc_idx] = r[P.bc_idx]
            zp, z = z, zn
        for tiled in (True, False):
            ctx = _ctx((40, 23), tiled=tiled)
            ctx.jac_setup(dev(u))
            ctx.precond_setup(pc_sweeps_shifted=n)
            got = ctx.precond_apply(dev(r)).cpu().numpy()
            assert np.abs(got - z).max() <= 1e-12 * np.abs(z).max()


@pytest.mark.parametrize("tiled", [True, False])
def test_reference_configuration_solve(tiled):
    """nl.py as written: 60 x 75, start 0.9, atol 1e-6, max 10 iterations, gmres (nl.py:17-18,48,50-64): the oracle's Newton
    count and its converged solution."""
    P = NlProblem()
    want = P.solve(atol=1e-12, max_it=30)[0]
    ctx = _ctx((60, 75), tiled=tiled)
    w = torch.full((P.nv,), 0.9, dtype=torch.float64, device="cuda")
    rep = ctx.newton_solve(w, w, absolute_tolerance=1e-6, maximum_iterations=10)
    assert rep["converged"] == 1 and rep["niters"] == 4 and rep["krylov_its"] <= 160
    assert rep["fnorm0"] == pytest.approx(0.0604053611, rel=1e-8)
    got = w.cpu().numpy()
    assert np.linalg.norm(got - want) <= 1e-10 * np.linalg.norm(want)
    # tight tolerances: the discrete solution itself
    w.fill_(0.9)
    rep = ctx.newton_solve(w, w, absolute_tolerance=1e-11, maximum_iterations=10, krylov_rtol=1e-9)
    assert rep["converged"] == 1
    assert np.linalg.norm(w.cpu().numpy() - want) <= 1e-13 * np.linalg.norm(want)
    # the reference-facing host call: same iterates
    wh = np.full(P.nv, 0.9)
    rep_h = ctx.step_host(wh, wh.copy(), 0.1, absolute_tolerance=1e-11, maximum_iterations=10, krylov_rtol=1e-9)
    assert rep_h["converged"] == 1 and np.array_equal(wh, w.cpu().numpy())


def test_nonconvergence_and_unsupported_calls():
    ctx = _ctx((60, 75))
    w = torch.full((ctx.n_owned,), 0.9, dtype=torch.float64, device="cuda")
    rep = ctx.newton_solve(w, w, absolute_tolerance=1e-6, maximum_iterations=2)
    assert rep["converged"] == 0 and rep["reason"] == -5 and rep["niters"] == 2      # reported, never raised (nl.py:57)
    with pytest.raises(RuntimeError, match="energy"):
        ctx.energy(w)
    from pfhubbenchmarks_b200 import engine
    with pytest.raises(RuntimeError, match="right-diagonal"):
        engine.Context("nl", (8, 8, 8), (1.0, 1.0, 1.0))


def test_driver_on_the_gpu(capsys):
    """drivers/nl.py = the reference script on the front end: SNES then Newton, both 4 iterations, and the launch counter
    shows the library's own kernels did the work."""
    from pfhubbenchmarks_b200.drivers import nl
    out = nl.main([])
    want = NlProblem().solve(atol=1e-12, max_it=30)[0]
    assert out["snes"][:2] == (4, True) and out["newton"][:2] == (4, True)
    assert np.linalg.norm(out["u"] - want) <= 1e-10 * np.linalg.norm(want)
    text = capsys.readouterr().out
    assert "SNES converged = True , niters = 4" in text and "Newton converged = True , niters = 4" in text


def test_full_size_properties():
    """2048^2 (the synthetic size of the other 2-D problems): tiled == gather to round-off, J.x = directional derivative
    of F; a Newton solve from the script's start at 1024^2."""
    n = 2048
    ctx_t, ctx_g = _ctx((n, n), tiled=True), _ctx((n, n), tiled=False)
    N = ctx_t.n_owned
    g = torch.Generator(device="cuda").manual_seed(11)
    u = 0.9 + 0.3 * torch.rand(N, dtype=torch.float64, device="cuda", generator=g)
    x = torch.randn(N, dtype=torch.float64, device="cuda", generator=g)
    Rt, Rg = ctx_t.residual(u, u), ctx_g.residual(u, u)
    assert float((Rt - Rg).abs().max()) <= 1e-13 * float(Rg.abs().max())
    ctx_t.jac_setup(u)
    ctx_g.jac_setup(u)
    Jt, Jg = ctx_t.jac_apply(x), ctx_g.jac_apply(x)
    assert float((Jt - Jg).abs().max()) <= 1e-13 * float(Jg.abs().max())
    eps = 1e-6
    fd = (ctx_t.residual(u + eps * x, u) - ctx_t.residual(u - eps * x, u)) / (2 * eps)
    assert float((Jt - fd).abs().max()) <= 1e-6 * float(fd.abs().max())
    assert ctx_t.launch_count() > 0
    # 1024^2 solve: Krylov counts grow like 1/h with a fixed polynomial degree (CPU leg: 52 / 96 iterations at 128^2 / 256^2
    # with 12 sweeps, 44 / 79 at 256^2 / 512^2 with 32), Newton's count does not
    ctx = _ctx((1024, 1024))
    w = torch.full((ctx.n_owned,), 0.9, dtype=torch.float64, device="cuda")
    rep = ctx.newton_solve(w, w, absolute_tolerance=1e-6, maximum_iterations=10, krylov_maximum_iterations=5000, pc_sweeps_shifted=32)
    assert rep["converged"] == 1 and rep["niters"] <= 5 and rep["krylov_its"] <= 2000   # measured: 3 / 817 (GMRES(30) restarts)
    R = ctx.residual(w, w)
    assert float(R.norm()) < 1e-6
