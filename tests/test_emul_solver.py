"""The product's host solver (csrc/pfh_solver.cpp: Newton, 'basic' / 'cp' line searches, GMRES with the fp64 and the
fp32-stored basis, Schur/Chebyshev preconditioner with its block layouts, Dirichlet handling) compiled for the CPU by
tests/emul over loop versions of the kernel launchers, against the oracle's Newton/LU step.  No GPU: this is the leg that
lets solver-side changes be checked where there is none (the launch layer is what the GPU tests cover)."""
import numpy as np
import pytest

import emul
from helpers import discretization
from oracle import krylov

CASES = [
    # name, mesh kind, problem id, cells, lengths, params, dt, line search
    ("bench1", 2, 1, (12, 10), (24.0, 20.0), [0.3, 0.7, 5.0, 2.0, 5.0], 0.1, 0),
    ("bench2", 2, 2, (10, 8), (20.0, 16.0), [0.3, 0.7, 2.0, 3.0, 3.0, 5.0, 1.0, 5.0, 5.0], 0.01, 1),
    ("bench3", 2, 3, (10, 8), (27.0, 22.0), [1.0, 1.0, 10.0, 10.0 / 0.6267, 27.0 * 22.0], 0.01, 0),
    ("bench6", 2, 6, (10, 8), (10.0, 8.0), [0.3, 0.7, 5.0, 2.0, 5.0, 0.09, 90.0], 0.01, 1),
    ("b13d", 3, 1, (6, 5, 4), (12.0, 10.0, 8.0), [0.3, 0.7, 5.0, 2.0, 5.0], 0.1, 0),
]


@pytest.mark.parametrize("name,mk,pid,ncell,length,params,dt,ls", CASES)
def test_newton_krylov_step_matches_oracle(name, mk, pid, ncell, length, params, dt, ls):
    D = discretization(name, ncell, length)
    u0 = D.prob.ic(D.mesh)
    w, rep = emul.newton_step(mk, pid, ncell, length, params, dt, u0, u0, basis_precision=1, absolute_tolerance=1e-10, krylov_rtol=1e-10,
                              maximum_iterations=30, krylov_maximum_iterations=5000, line_search=ls)
    assert rep["converged"] == 1 and rep["fnorm"] < 1e-10 * max(1.0, rep["fnorm0"]) + 1e-10, rep
    u, n, ok = D.solve_step(u0, u0, dt, atol=1e-11, linear="lu" if mk == 2 else krylov.make_linear(rtol=1e-12))
    assert ok
    for f in range(D.nf):
        scale = max(np.linalg.norm(u[f]), 1e-3 * np.sqrt(u[f].size))
        assert np.linalg.norm(w[f] - u[f]) / scale < 2e-8, (name, f, np.linalg.norm(w[f] - u[f]) / scale)


@pytest.mark.parametrize("name,mk,pid,ncell,length,params,dt,ls", CASES)
def test_bicgstab_step_matches_oracle(name, mk, pid, ncell, length, params, dt, ls):
    """linear_solver = 1: right-preconditioned BiCGStab (the reduced Ostwald drivers' choice, bench2_eta1.py:113) lands on
    the same Newton iterate as the oracle's direct solve; an iteration is two operator applies."""
    D = discretization(name, ncell, length)
    u0 = D.prob.ic(D.mesh)
    w, rep = emul.newton_step(mk, pid, ncell, length, params, dt, u0, u0, absolute_tolerance=1e-10, krylov_rtol=1e-10,
                              maximum_iterations=30, krylov_maximum_iterations=5000, line_search=ls, linear_solver=1)
    assert rep["converged"] == 1 and rep["fnorm"] < 1e-10 * max(1.0, rep["fnorm0"]) + 1e-10, rep
    assert rep["pc_applies"] <= 2 * rep["krylov_its"]
    u, n, ok = D.solve_step(u0, u0, dt, atol=1e-11, linear="lu" if mk == 2 else krylov.make_linear(rtol=1e-12))
    assert ok
    for f in range(D.nf):
        scale = max(np.linalg.norm(u[f]), 1e-3 * np.sqrt(u[f].size))
        assert np.linalg.norm(w[f] - u[f]) / scale < 2e-8, (name, f, np.linalg.norm(w[f] - u[f]) / scale)


@pytest.mark.parametrize("name,mk,pid,ncell,length,params,dt,ls", CASES[:1] + CASES[4:])
def test_compressed_basis_is_equivalent(name, mk, pid, ncell, length, params, dt, ls):
    """Default tolerances (atol 1e-6, GMRES rtol 1e-5): the fp32-stored basis needs the same number of Krylov iterations
    as the fp64 one (+-1) and lands on the same Newton iterate to the nonlinear tolerance."""
    D = discretization(name, ncell, length)
    u0 = D.prob.ic(D.mesh)
    w64, r64 = emul.newton_step(mk, pid, ncell, length, params, dt, u0, u0, basis_precision=1)
    w32, r32 = emul.newton_step(mk, pid, ncell, length, params, dt, u0, u0, basis_precision=2)
    wau, rau = emul.newton_step(mk, pid, ncell, length, params, dt, u0, u0, basis_precision=0)
    assert r64["converged"] == r32["converged"] == 1 and r64["niters"] == r32["niters"]
    assert abs(r64["krylov_its"] - r32["krylov_its"]) <= 1, (r64, r32)
    assert np.linalg.norm(w64[0] - w32[0]) / np.linalg.norm(w64[0]) < 1e-7
    assert np.array_equal(wau, w32) and rau["krylov_its"] == r32["krylov_its"]          # auto = fp32 at rtol 1e-5


def test_nonconvergence_is_reported_not_raised():
    """maximum_iterations hit => converged = 0, reason < 0 (the drivers' dt-halving retry depends on it, bench1.py:164-177)."""
    D = discretization("bench1", (8, 6), (16.0, 12.0))
    u0 = D.prob.ic(D.mesh)
    w, rep = emul.newton_step(2, 1, (8, 6), (16.0, 12.0), [0.3, 0.7, 5.0, 2.0, 5.0], 0.1, u0, u0, absolute_tolerance=1e-14,
                              relative_tolerance=0.0, maximum_iterations=1)
    assert rep["converged"] == 0 and rep["reason"] == -5 and rep["niters"] == 1
    # a preconditioner choice the Schur layout does not cover falls back to point Jacobi and still converges
    w, rep = emul.newton_step(2, 1, (8, 6), (16.0, 12.0), [0.3, 0.7, 5.0, 2.0, 5.0], 0.1, u0, u0, preconditioner=1,
                              krylov_maximum_iterations=4000)
    assert rep["converged"] == 1


@pytest.mark.parametrize("basis", [1, 2])
@pytest.mark.parametrize("opts", [dict(krylov_restart=5), dict(krylov_restart=3), dict(pc_variant=0), dict(pc_variant=0, krylov_restart=4),
                                  dict(preconditioner=2, krylov_restart=30, krylov_maximum_iterations=20000)])
def test_restarts_and_preconditioner_variants(basis, opts):
    """Short restart cycles (the restart path recomputes the true residual and re-seeds the basis), the full
    block-triangular variant of the Schur preconditioner and no preconditioner at all reach the same Newton solution,
    with the fp64 and with the fp32-stored basis."""
    ncell, length, params = (8, 6), (16.0, 12.0), [0.3, 0.7, 5.0, 2.0, 5.0]
    D = discretization("bench1", ncell, length)
    u0 = D.prob.ic(D.mesh)
    ref, r0 = emul.newton_step(2, 1, ncell, length, params, 0.1, u0, u0, basis_precision=1, absolute_tolerance=1e-9, krylov_rtol=1e-6)
    kw = dict(absolute_tolerance=1e-9, krylov_rtol=1e-6, maximum_iterations=30)
    kw.update(opts)
    w, rep = emul.newton_step(2, 1, ncell, length, params, 0.1, u0, u0, basis_precision=basis, **kw)
    assert rep["converged"] == 1, rep
    assert np.linalg.norm(w[0] - ref[0]) / np.linalg.norm(ref[0]) < 1e-7, (opts, rep)


def test_stagnating_linear_solve_stops_newton():
    """GMRES(2) stagnates on the block-triangularly preconditioned system of the second Newton iteration: the Krylov
    iteration cap is hit, Newton stops with reason -3 (SNES_DIVERGED_LINEAR_SOLVE) and reports converged = 0 -- identically
    for both basis storages."""
    ncell, length, params = (8, 6), (16.0, 12.0), [0.3, 0.7, 5.0, 2.0, 5.0]
    D = discretization("bench1", ncell, length)
    u0 = D.prob.ic(D.mesh)
    reps = [emul.newton_step(2, 1, ncell, length, params, 0.1, u0, u0, basis_precision=b, absolute_tolerance=1e-9, krylov_rtol=1e-6,
                             krylov_restart=2, krylov_maximum_iterations=300)[1] for b in (1, 2)]
    for rep in reps:
        assert rep["converged"] == 0 and rep["reason"] == -3 and rep["niters"] == 1
    assert reps[0]["krylov_its"] == reps[1]["krylov_its"]


def test_backtracking_line_search():
    """'bt' (PETSc's default backtracking): identical iterates to 'basic' when the full Newton step already decreases
    the residual, and a converged step where the full step ('basic') and the secant step ('cp') both diverge
    (benchmark 2 from its initial condition at dt = 0.64)."""
    ncell, length, params = (10, 8), (20.0, 16.0), [0.3, 0.7, 5.0, 2.0, 5.0]
    D = discretization("bench1", ncell, length)
    u0 = D.prob.ic(D.mesh)
    wb, rb = emul.newton_step(2, 1, ncell, length, params, 0.1, u0, u0, basis_precision=1, line_search=0)
    wt, rt = emul.newton_step(2, 1, ncell, length, params, 0.1, u0, u0, basis_precision=1, line_search=2)
    assert np.array_equal(wb, wt) and rb["niters"] == rt["niters"] and rb["krylov_its"] == rt["krylov_its"]
    p2 = [0.3, 0.7, 2.0, 3.0, 3.0, 5.0, 1.0, 5.0, 5.0]
    D2 = discretization("bench2", (12, 10), (24.0, 20.0))
    v0 = D2.prob.ic(D2.mesh)
    out = {ls: emul.newton_step(2, 2, (12, 10), (24.0, 20.0), p2, 0.64, v0, v0, basis_precision=1, line_search=ls, maximum_iterations=25)
           for ls in (0, 1, 2)}
    assert out[0][1]["converged"] == 0 and out[1][1]["converged"] == 0
    w, rep = out[2]
    assert rep["converged"] == 1 and rep["fnorm"] < 1e-6
    u, n, ok = D2.solve_step(w, v0, 0.64, atol=1e-11)            # the oracle's Newton/LU started from that solution stays there
    assert ok and max(np.linalg.norm(w[f] - u[f]) / max(np.linalg.norm(u[f]), 1.0) for f in range(6)) < 1e-6


# ---------------------------------------------------------------------------------------------------------------------
# geometric multigrid for the Poisson block of bench6 (csrc/pfh_mg2.cuh, mg_vcycle in csrc/pfh_solver.cpp)
# ---------------------------------------------------------------------------------------------------------------------
def _prolongation(nxc, nyc):
    """P1 interpolation from the crossed mesh (nxc, nyc) to (2 nxc, 2 nyc), assembled explicitly from the oracle's meshes:
    every fine node is evaluated in the coarse P1 space (oracle.fem.eval_points is a cell search: no use of the structure)."""
    import scipy.sparse as sp
    from oracle import mesh as omesh
    from oracle.fem import eval_points
    mc, mf = omesh.crossed2d(nxc, nyc, 2.0 * nxc, 3.0 * nyc), omesh.crossed2d(2 * nxc, 2 * nyc, 2.0 * nxc, 3.0 * nyc)
    cols = []
    for j in range(mc.nv):
        e = np.zeros(mc.nv)
        e[j] = 1.0
        cols.append(eval_points(mc, e, mf.coords))
    return sp.csr_matrix(np.array(cols).T), mc, mf


@pytest.mark.parametrize("nxc,nyc", [(2, 2), (3, 2), (4, 5)])
def test_multigrid_transfers_are_p1_interpolation_and_its_transpose(nxc, nyc):
    P, mc, mf = _prolongation(nxc, nyc)
    rng = np.random.default_rng(nxc * 10 + nyc)
    xc, tf = rng.standard_normal(mc.nv), rng.standard_normal(mf.nv)
    got_p, _ = emul.mg_transfer((2 * nxc, 2 * nyc), (2.0 * nxc, 3.0 * nyc), 1, xc)
    got_r, _ = emul.mg_transfer((2 * nxc, 2 * nyc), (2.0 * nxc, 3.0 * nyc), 0, tf)
    assert np.abs(got_p - P @ xc).max() < 1e-14
    assert np.abs(got_r - P.T @ tf).max() < 1e-14
    # nested spaces: the re-discretised coarse stiffness matrix is the Galerkin operator
    from oracle.fem import P1Space
    Kc, Kf = P1Space(mc).K, P1Space(mf).K
    assert abs(P.T @ Kf @ P - Kc).max() < 1e-13


def test_multigrid_level_count():
    assert emul.mg_transfer((100, 100), (100.0, 100.0), 1, np.zeros(51 * 51 + 50 * 50))[1] == 2      # 100 -> 50 -> 25
    assert emul.mg_transfer((64, 32), (64.0, 32.0), 1, np.zeros(33 * 17 + 32 * 16))[1] == 4          # .. -> 4 x 2


@pytest.mark.parametrize("ncell,length", [((16, 12), (16.0, 12.0)), ((40, 40), (40.0, 40.0))])
def test_bench6_step_with_the_multigrid_poisson_block(ncell, length):
    """bench6 with V-cycles in the preconditioner's Poisson block lands on the oracle's Newton iterate like the Chebyshev
    solve it replaces (mg_cycles = 0), with no more Krylov iterations."""
    name, mk, pid, _, _, params, dt, ls = CASES[3]
    D = discretization(name, ncell, length)
    u0 = D.prob.ic(D.mesh)
    kw = dict(basis_precision=1, absolute_tolerance=1e-10, krylov_rtol=1e-10, maximum_iterations=30, krylov_maximum_iterations=5000,
              line_search=ls)
    u, n, ok = D.solve_step(u0, u0, dt, atol=1e-11, linear="lu")
    assert ok
    its = {}
    for cycles in (0, 1, 2):
        w, rep = emul.newton_step(mk, pid, ncell, length, params, dt, u0, u0, mg_cycles=cycles, **kw)
        assert rep["converged"] == 1, (cycles, rep)
        for f in range(D.nf):
            scale = max(np.linalg.norm(u[f]), 1e-3 * np.sqrt(u[f].size))
            assert np.linalg.norm(w[f] - u[f]) / scale < 2e-8, (cycles, f)
        its[cycles] = rep["krylov_its"]
    assert its[1] <= 1.1 * its[0] and its[2] <= 1.1 * its[0], its      # measured 69 / 68 / 69 and 65 / 68 / 65


@pytest.mark.parametrize("ncell", [(64, 64), (100, 100), (96, 40)])
def test_multigrid_contraction_is_mesh_independent(ncell):
    """Residual reduction of the product's V(4,4) cycle on K_D x = b (Dirichlet on x = 0, Lx as the phi row of bench6,
    bench6.py:77-90): <= 0.06 per cycle at every size (oracle prototype: 0.02-0.05), where the 240 Chebyshev sweeps it
    replaces reach 0.14 at 100^2."""
    import scipy.sparse as sp
    from oracle import mesh as omesh
    from oracle.fem import P1Space
    nx, ny = ncell
    m = omesh.crossed2d(nx, ny, float(nx), float(ny))
    K = P1Space(m).K
    ix = np.arange((nx + 1) * (ny + 1)) % (nx + 1)
    D = np.where((ix == 0) | (ix == nx))[0]
    keep = np.ones(m.nv)
    keep[D] = 0.0
    A = (sp.diags(keep) @ K + sp.diags(1.0 - keep)).tocsr()
    rng = np.random.default_rng(1)
    xt = rng.standard_normal(m.nv)
    b = A @ xt
    r0 = np.linalg.norm(b)
    r1 = np.linalg.norm(b - A @ emul.mg_solve(ncell, (float(nx), float(ny)), 1, b))
    r3 = np.linalg.norm(b - A @ emul.mg_solve(ncell, (float(nx), float(ny)), 3, b))
    assert r1 / r0 < 0.06 and (r3 / r0) ** (1.0 / 3.0) < 0.08, (r1 / r0, (r3 / r0) ** (1.0 / 3.0))
    x1 = emul.mg_solve(ncell, (float(nx), float(ny)), 1, b)
    assert np.array_equal(x1[D], b[D])                       # Dirichlet values are imposed exactly
