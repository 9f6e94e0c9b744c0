"""GPU parity tests proper: the CUDA path (through the C ABI) against the oracle, the committed
golden vectors of the reference, and size-independent properties.  Run on the B200 box:
    python -m pytest tests -m gpu -x -q
"""
import numpy as np
import pytest

from helpers import REFERENCE_MESH, discretization, oracle_step_3d, rel_l2, rel_max

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _engine():
    from pfhubbenchmarks_b200 import engine
    return engine


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).ravel()).cuda()


def to_np(t, nf):
    return t.cpu().numpy().reshape(nf, -1)


OP_CASES = [
    ("bench1", (7, 5), (14.0, 10.0)),
    ("bench1", (33, 17), (50.0, 34.0)),
    ("bench1", (100, 100), (200.0, 200.0)),
    ("b13d", (5, 4, 3), (10.0, 8.0, 6.0)),
    ("b13d", (70, 9, 21), (140.0, 18.0, 42.0)),      # ragged: not multiples of the tile sizes
    ("b13d", (16, 16, 16), (32.0, 32.0, 32.0)),
    # even vertex counts in x take the TMA-staged kernels (pfh_stencil3_tma.cu); ragged vs their 64x8 / 32x8 tiles
    ("b13d", (15, 16, 17), (30.0, 32.0, 34.0)),
    ("b13d", (71, 9, 40), (142.0, 18.0, 80.0)),
    ("b13d", (129, 21, 11), (258.0, 42.0, 22.0)),
    ("bench2", (12, 9), (24.0, 18.0)),
    ("bench3", (12, 9), (24.0, 18.0)),
    ("bench6", (12, 9), (24.0, 18.0)),
]


@pytest.mark.parametrize("name,ncell,length", OP_CASES)
def test_operators_match_oracle(name, ncell, length):
    """residual, J.v, energy/solute, initial condition: <= 1e-13 relative (SURVEY section 7 gate)."""
    eng = _engine()
    rng = np.random.default_rng(7)
    D = discretization(name, ncell, length)
    nf, nv = D.nf, D.nv
    ctx = eng.Context(name, ncell, length)
    assert ctx.n_owned == nv and ctx.nf == nf
    ic = to_np(ctx.initial_condition(), nf)
    ic_o = D.prob.ic(D.mesh)
    assert rel_max(ic, ic_o) < 1e-14
    U = ic_o + 0.1 * rng.standard_normal((nf, nv))
    U0 = ic_o + 0.1 * rng.standard_normal((nf, nv))
    X = rng.standard_normal((nf, nv))
    ctx.dt = 0.37
    R = to_np(ctx.residual(to_dev(U), to_dev(U0)), nf)
    assert rel_max(R, D.residual(U, U0, 0.37)) < 1e-13
    Ud = to_dev(U)
    ctx.jac_setup(Ud)
    Y = to_np(ctx.jac_apply(to_dev(X)), nf)
    assert rel_max(Y, (D.jacobian(U, 0.37) @ X.ravel()).reshape(nf, -1)) < 1e-13
    E, M = ctx.energy(Ud)
    assert abs(E / D.energy(U) - 1) < 1e-13 and abs(M / D.mass(U) - 1) < 1e-13
    # host-buffer entry points give the same numbers
    assert np.array_equal(ctx.residual_host(U.ravel().copy(), U0.ravel().copy()).reshape(nf, -1), R)
    assert np.array_equal(ctx.jac_apply_host(U.ravel().copy(), X.ravel().copy()).reshape(nf, -1), Y)
    ctx.close()


@pytest.mark.parametrize("name,ncell,length", [("bench1", (30, 6), (60.0, 12.0)), ("bench1", (31, 7), (62.0, 21.0)),
                                               ("bench1", (61, 13), (61.0, 26.0)), ("bench1", (70, 9), (140.0, 18.0)),
                                               ("bench1", (1, 1), (2.0, 2.0)), ("bench1", (257, 130), (514.0, 260.0)),
                                               ("bench3", (33, 9), (66.0, 18.0)), ("bench6", (33, 16), (33.0, 16.0)),
                                               ("bench2", (12, 9), (24.0, 18.0))])
def test_square_centric_2d_operators(name, ncell, length):
    """k2_op_tiled (every crossed square evaluated once per launch, csrc/pfh_op2_tiled.cuh; sizes around its 31 x 7 vertex
    tile) against the oracle and against the vertex-centric gather kernel ("op2_tiled" 0).  bench2's tile does not fit the
    static shared memory: both settings run the gather kernel there."""
    eng = _engine()
    rng = np.random.default_rng(9)
    D = discretization(name, ncell, length)
    nf, nv = D.nf, D.nv
    ic_o = D.prob.ic(D.mesh)
    U = ic_o + 0.1 * rng.standard_normal((nf, nv))
    U0 = ic_o + 0.1 * rng.standard_normal((nf, nv))
    X = rng.standard_normal((nf, nv))
    got = {}
    for tiled in (1, 0):
        ctx = eng.Context(name, ncell, length)
        ctx.set_option("op2_tiled", tiled)
        ctx.dt = 0.37
        R = to_np(ctx.residual(to_dev(U), to_dev(U0)), nf)
        Ud = to_dev(U)
        ctx.jac_setup(Ud)
        Y = to_np(ctx.jac_apply(to_dev(X)), nf)
        assert rel_max(R, D.residual(U, U0, 0.37)) < 1e-13
        assert rel_max(Y, (D.jacobian(U, 0.37) @ X.ravel()).reshape(nf, -1)) < 1e-13
        got[tiled] = (R, Y)
        ctx.close()
    assert rel_max(got[1][0], got[0][0]) < 1e-14 and rel_max(got[1][1], got[0][1]) < 1e-14


@pytest.mark.parametrize("name,ncell,length", [("bench1", (20, 14), (40.0, 28.0)), ("b13d", (9, 8, 7), (18.0, 16.0, 14.0))])
def test_preconditioner_matches_cpu_statement(name, ncell, length):
    """P^-1 r of the CUDA Schur/Chebyshev preconditioner == oracle/krylov.SchurChebPC (same sweeps)."""
    from oracle import krylov
    eng = _engine()
    rng = np.random.default_rng(5)
    D = discretization(name, ncell, length)
    U = D.prob.ic(D.mesh)
    r = rng.standard_normal(2 * D.nv)
    for variant, vname in [(0, "full"), (1, "upper")]:
        ctx = eng.Context(name, ncell, length)
        ctx.dt = 0.4
        Ud = to_dev(U)
        ctx.jac_setup(Ud)
        ctx.precond_setup(pc_variant=variant, pc_sweeps_mass=3, pc_sweeps_shifted=5)
        z = ctx.precond_apply(to_dev(r)).cpu().numpy()
        # 3-D: the product takes the upper end of the shifted solves' Chebyshev interval from the operator's symbol
        pc = krylov.SchurChebPC(D, U, 0.4, n_m=3, n_a=5, variant=vname, symbol_lmax=(name == "b13d"))
        assert rel_max(z, pc(r)) < 1e-11, vname
        ctx.close()


def test_bench1_golden_step_kats(golden):
    """FEniCS snapshot k -> CUDA step -> FEniCS snapshot k+1 (results/bench1/conc*.vtu)."""
    eng = _engine()
    ncell, length = REFERENCE_MESH["bench1"]
    snaps = golden["bench1_fields"]
    ctx = eng.Context("bench1", ncell, length)
    nv = ctx.n_owned
    w0 = ctx.initial_condition()
    w = w0.clone()
    ctx.dt = 0.1
    rep = ctx.newton_solve(w, w0, absolute_tolerance=1e-10, krylov_rtol=1e-8, maximum_iterations=20)
    assert rep["converged"], rep
    assert rel_l2(w[:nv].cpu().numpy(), snaps["c"][0]) < 1e-10
    for k in range(5):
        dt = float(snaps["times"][k + 1] - snaps["times"][k])
        u0 = np.zeros((2, nv)); u0[0] = snaps["c"][k]
        w0 = to_dev(u0); w = w0.clone()
        ctx.dt = dt
        rep = ctx.newton_solve(w, w0, absolute_tolerance=1e-10, krylov_rtol=1e-8, maximum_iterations=20)
        assert rep["converged"], rep
        assert rel_l2(w[:nv].cpu().numpy(), snaps["c"][k + 1]) < 2e-9, k
    ctx.close()


def test_bench1_trajectory_vs_reference_csv_and_oracle(golden):
    """North-star bar: replaying the reference's dt list, free energy within 1e-6 of results/bench1_out.csv
    and fields within 1e-8 relative L2 of the oracle trajectory (first rows; the 100-step run is in bench/)."""
    eng = _engine()
    ncell, length = REFERENCE_MESH["bench1"]
    D = discretization("bench1", ncell, length)
    csv = golden["csv"]["bench1"]
    ctx = eng.Context("bench1", ncell, length)
    nv = ctx.n_owned
    w = ctx.initial_condition()
    u = D.prob.ic(D.mesh)
    t = 0.0
    for k in range(12):
        dt = float(csv[k, 0] - t); t = float(csv[k, 0])
        w0 = w.clone()
        ctx.dt = dt
        rep = ctx.newton_solve(w, w0, absolute_tolerance=1e-10, krylov_rtol=1e-8, maximum_iterations=25)
        assert rep["converged"], (k, rep)
        u, n, ok = D.solve_step(u, u, dt)
        E, M = ctx.energy(w)
        assert abs(E / csv[k, 1] - 1) < 1e-6 and abs(M / csv[k, 2] - 1) < 1e-9, (k, E, csv[k, 1])
        assert rel_l2(w[:nv].cpu().numpy(), u[0]) < 1e-8, k
    ctx.close()


def test_reference_solver_settings_semantics():
    """(niters, converged) as the drivers use it: atol 1e-6 / 10 its converges from the IC at dt=0.1;
    max_it=1 at a large dt returns converged=False without raising (bench1.py:86,164-177)."""
    eng = _engine()
    ncell, length = REFERENCE_MESH["bench1"]
    ctx = eng.Context("bench1", ncell, length)
    w0 = ctx.initial_condition(); w = w0.clone()
    ctx.dt = 0.1
    rep = ctx.newton_solve(w, w0)
    assert rep["converged"] and 1 <= rep["niters"] <= 5 and rep["fnorm"] < 1e-6
    w = w0.clone()
    ctx.dt = 3.2
    rep = ctx.newton_solve(w, w0, maximum_iterations=1)
    assert not rep["converged"] and rep["reason"] < 0
    ctx.close()


def test_b13d_step_matches_oracle():
    """3-D Kuhn mesh (no reference data exists: parity unpinned, oracle self-consistency only)."""
    eng = _engine()
    _b13d_steps(eng, (16, 16, 16), (32.0, 32.0, 32.0))      # odd vertex count: gather kernels
    _b13d_steps(eng, (15, 17, 16), (30.0, 34.0, 32.0))      # even vertex count: TMA-staged kernels


def _b13d_steps(eng, ncell, length):
    D = discretization("b13d", ncell, length)
    ctx = eng.Context("b13d", ncell, length)
    nv = ctx.n_owned
    u = D.prob.ic(D.mesh)
    w = ctx.initial_condition()
    for dt in (0.1, 0.2, 0.4):
        w0 = w.clone()
        ctx.dt = dt
        rep = ctx.newton_solve(w, w0, absolute_tolerance=1e-10, krylov_rtol=1e-8, maximum_iterations=20)
        assert rep["converged"], rep
        u, n, ok = oracle_step_3d(D, u, u, dt)
        assert ok
        got = w.cpu().numpy().reshape(2, -1)
        assert rel_l2(got[0], u[0]) < 1e-9 and rel_l2(got[1], u[1]) < 1e-7
        E, M = ctx.energy(w)
        assert abs(E / D.energy(u) - 1) < 1e-9 and abs(M / D.mass(u) - 1) < 1e-12
    ctx.close()


def test_tma_kernels_match_gather_kernels():
    """Fast path (staged scalar kernel with the z-march register reuse, cube-centric J.v) against the simple gather
    kernels: TMA loader (even vertex count in x), cp.async loader (odd count, e.g. the reference's 51^3 mesh) and the
    cp.async loader forced onto the even mesh; every tile-shape variant of the scalar kernel."""
    eng = _engine()
    rng = np.random.default_rng(11)
    for ncell, length in [((127, 70, 45), (254.0, 140.0, 90.0)), ((50, 50, 50), (100.0, 100.0, 100.0)), ((70, 9, 5), (140.0, 18.0, 10.0))]:
        ctx = eng.Context("b13d", ncell, length)
        n = ctx.n_owned
        U = ctx.initial_condition()
        U += 0.05 * torch.from_numpy(rng.standard_normal(2 * n)).cuda()
        X = torch.from_numpy(rng.standard_normal(2 * n)).cuda()
        ctx.dt = 0.8
        out = {}
        for key, (tma, loader, variant, jvar) in {"gather": (0, 1, 0, 0), "tma": (1, 1, 0, 0), "cpasync": (1, 0, 0, 0), "v1": (1, 1, 1, 1),
                                                  "v2": (1, 1, 2, 1), "v3": (1, 0, 3, 0), "v4": (1, 1, 4, 0), "v5": (1, 1, 5, 1)}.items():
            ctx.set_option("use_tma", tma)
            ctx.set_option("use_tma_loader", loader)
            ctx.set_option("scalar_variant", variant)
            ctx.set_option("jac_variant", jvar)
            ctx.jac_setup(U)
            y = ctx.jac_apply(X)
            R = ctx.residual(U, X)
            res = []
            for pcv in (0, 1):
                ctx.precond_setup(pc_variant=pcv, pc_sweeps_mass=3, pc_sweeps_shifted=6)
                res.append(ctx.precond_apply(X).cpu().numpy())
            out[key] = (y.cpu().numpy(), R.cpu().numpy(), res[0], res[1])
        for key in out:
            if key == "gather":
                continue
            assert rel_max(out[key][0], out["gather"][0]) < 1e-13, (ncell, key)
            assert rel_max(out[key][1], out["gather"][1]) < 1e-13, (ncell, key)
            assert rel_max(out[key][2], out["gather"][2]) < 1e-12, (ncell, key)
            assert rel_max(out[key][3], out["gather"][3]) < 1e-12, (ncell, key)
        ctx.close()


def test_backtracking_line_search_on_the_gpu():
    """'bt' (PETSc's default backtracking; csrc/pfh_solver.cpp) through the C ABI on the device: the same iterates as 'basic'
    when the full Newton step is accepted, and a converged step of benchmark 2 from its initial condition at dt = 0.64,
    where the full step ('basic') and the secant step ('cp') both diverge -- the CPU solver leg's test, on the real kernels."""
    eng = _engine()
    ctx = eng.Context("bench1", (10, 8), (20.0, 16.0))
    w0 = ctx.initial_condition()
    ctx.dt = 0.1
    out = {}
    for ls in (0, 2):
        w = w0.clone()
        rep = ctx.newton_solve(w, w0, line_search=ls)
        out[ls] = (w.cpu().numpy(), rep)
    assert np.array_equal(out[0][0], out[2][0]) and out[0][1]["niters"] == out[2][1]["niters"]
    assert out[0][1]["krylov_its"] == out[2][1]["krylov_its"]
    ctx.close()
    ncell, length = (12, 10), (24.0, 20.0)
    D2 = discretization("bench2", ncell, length)
    ctx = eng.Context("bench2", ncell, length)
    v0 = ctx.initial_condition()
    ctx.dt = 0.64
    reps = {}
    for ls in (0, 1, 2):
        w = v0.clone()
        reps[ls] = (ctx.newton_solve(w, v0, line_search=ls, maximum_iterations=25), w.cpu().numpy().reshape(6, -1))
    assert not reps[0][0]["converged"] and not reps[1][0]["converged"]
    rep, w = reps[2]
    assert rep["converged"] and rep["fnorm"] < 1e-6 and rep["residual_evals"] > rep["niters"] + 1   # it did backtrack
    u0 = v0.cpu().numpy().reshape(6, -1)
    u, n, ok = D2.solve_step(w, u0, 0.64, atol=1e-11)            # the oracle's Newton/LU started from that solution stays there
    assert ok and max(np.linalg.norm(w[f] - u[f]) / max(np.linalg.norm(u[f]), 1.0) for f in range(6)) < 1e-6
    ctx.close()


def test_step_host_equals_device_path():
    eng = _engine()
    ncell, length = (40, 30), (80.0, 60.0)
    ctx = eng.Context("bench1", ncell, length)
    w0 = ctx.initial_condition()
    w = w0.clone()
    ctx.dt = 0.2
    rep = ctx.newton_solve(w, w0, absolute_tolerance=1e-9)
    E, M = ctx.energy(w)
    h0 = w0.cpu().numpy().copy()
    h = h0.copy()
    rep2 = ctx.step_host(h, h0, 0.2, absolute_tolerance=1e-9)
    assert rep2["converged"] and rep2["niters"] == rep["niters"]
    assert np.array_equal(h, w.cpu().numpy())
    assert rep2["energy"] == E and rep2["mass"] == M
    ctx.close()


def test_large_grid_properties():
    """Full-size properties: solute conserved to round-off, free energy decreases, on 192^3 and 2048^2."""
    eng = _engine()
    for name, ncell, length in [("b13d", (191, 191, 191), (382.0, 382.0, 382.0)), ("bench1", (2048, 2048), (4096.0, 4096.0))]:
        ctx = eng.Context(name, ncell, length, krylov_basis=20)
        w = ctx.initial_condition()
        E0, M0 = ctx.energy(w)
        for dt in (0.1, 0.2):
            w0 = w.clone()
            ctx.dt = dt
            rep = ctx.newton_solve(w, w0, krylov_restart=20)
            assert rep["converged"], rep
            E, M = ctx.energy(w)
            assert E < E0 and abs(M / M0 - 1) < 1e-12
            E0 = E
        ctx.close()


def test_bench1_100_steps_fields_and_energy_vs_oracle():
    """North-star bar at its stated horizon: 100 backward-Euler steps of bench1 (reference mesh, dt list of SURVEY
    section 8(d): 0.1, 0.2, 0.4, 0.8, 1.6, 1.6, ...), CUDA vs the exactly converged oracle trajectory:
    concentration within 1e-8 relative L2 and free energy within 1e-6 relative at EVERY step."""
    eng = _engine()
    ncell, length = REFERENCE_MESH["bench1"]
    D = discretization("bench1", ncell, length)
    ctx = eng.Context("bench1", ncell, length)
    nv = ctx.n_owned
    w = ctx.initial_condition()
    u = D.prob.ic(D.mesh)
    worst_c = worst_e = 0.0
    for k in range(100):
        dt = min(0.1 * 2 ** k, 1.6) if k < 6 else 1.6
        w0 = w.clone()
        ctx.dt = dt
        rep = ctx.newton_solve(w, w0, absolute_tolerance=1e-10, krylov_rtol=1e-8, maximum_iterations=30)
        assert rep["converged"], (k, rep)
        u, n, ok = D.solve_step(u, u, dt, atol=1e-10)
        assert ok, k
        E, M = ctx.energy(w)
        worst_c = max(worst_c, rel_l2(w[:nv].cpu().numpy(), u[0]))
        worst_e = max(worst_e, abs(E / D.energy(u) - 1))
        assert abs(M / D.mass(u) - 1) < 1e-11
    print("100 steps: worst rel L2(c) %.2e, worst rel energy %.2e" % (worst_c, worst_e))
    assert worst_c < 1e-8 and worst_e < 1e-6
    ctx.close()


def test_c_abi_error_behaviour():
    """Argument errors come back as return codes + pfh_last_error (never a crash); non-convergence is not an error."""
    import ctypes as C
    from pfhubbenchmarks_b200 import _lib
    eng = _engine()
    lib = _lib.load()
    with pytest.raises(_lib.PfhError, match="3-D meshes support"):
        eng.Context("bench2", (8, 8, 8), (16.0, 16.0, 16.0))            # only Cahn-Hilliard exists in 3-D (b13d)
    with pytest.raises(_lib.PfhError, match="number of problem parameters"):
        eng.Context("bench1", (8, 8), (16.0, 16.0), params=[0.3, 0.7])
    cfg = _lib.Config(mesh_kind=7, problem=1)
    h = C.c_void_p()
    assert lib.pfh_create(C.byref(cfg), C.byref(h)) < 0 and b"mesh kind" in lib.pfh_last_error()
    ctx = eng.Context("bench1", (16, 16), (32.0, 32.0), krylov_basis=10)
    w0 = ctx.initial_condition(); w = w0.clone()
    with pytest.raises(_lib.PfhError, match="krylov_restart"):
        ctx.newton_solve(w, w0, krylov_restart=20)
    with pytest.raises(_lib.PfhError, match="dt must be positive"):
        ctx.dt = -1.0
    ctx.dt = 0.1
    rep = ctx.newton_solve(w, w0, krylov_restart=10, krylov_maximum_iterations=2, maximum_iterations=3)
    assert rep["converged"] == 0 and rep["reason"] < 0        # reported, not raised (bench1.py:86)
    ctx.close()


def test_snes_callback_boundary_drives_the_same_operators():
    """N3: an external Newton-GMRES that only uses the SNES-style callbacks (F, J, mult, pc_apply) reaches the same
    step solution as pfh_newton_solve and as the FEniCS snapshot."""
    from pfhubbenchmarks_b200 import pde_problems
    eng = _engine()
    ncell, length = REFERENCE_MESH["bench1"]
    ctx = eng.Context("bench1", ncell, length)
    nv = ctx.n_owned
    w0 = ctx.initial_condition()
    ctx.dt = 0.1
    x = w0.clone()
    prob = pde_problems.SnesPDEProblem(ctx, w0)
    its, ok = pde_problems.newton_krylov(prob, x, atol=1e-10)
    assert ok and its <= 6
    w = w0.clone()
    rep = ctx.newton_solve(w, w0, absolute_tolerance=1e-10, krylov_rtol=1e-8, maximum_iterations=20)
    assert rep["converged"]
    assert rel_l2(x.cpu().numpy(), w.cpu().numpy()) < 1e-10
    ctx.close()


def test_compressed_krylov_basis():
    """The GMRES basis is stored in fp32 at the default linear tolerance (fp64 arithmetic): the
    Newton solve converges to the same solution as with the fp64 basis (both to the nonlinear tolerance), with
    essentially the same Krylov iteration count; a tight linear tolerance selects the fp64 basis by itself."""
    from pfhubbenchmarks_b200 import engine
    out = {}
    for mode in (1, 2, 0):
        ctx = engine.Context("b13d", (48, 40, 36), (96.0, 80.0, 72.0))
        ctx.set_option("basis_precision", mode)
        w0 = ctx.initial_condition()
        w = w0.clone()
        its = 0
        for dt in (0.1, 0.2, 0.4):
            ctx.dt = dt
            w0.copy_(w)
            rep = ctx.newton_solve(w, w0, absolute_tolerance=1e-9)
            assert rep["converged"], rep
            its += rep["krylov_its"]
        out[mode] = (w.cpu().numpy().reshape(2, -1), its, ctx.energy(w))
        ctx.close()
    c64, i64, e64 = out[1]
    c32, i32, e32 = out[2]
    assert rel_l2(c32[0], c64[0]) < 1e-9 and abs(e32[0] / e64[0] - 1) < 1e-10
    assert abs(i32 - i64) <= max(2, i64 // 20), (i32, i64)
    assert np.array_equal(out[0][0], c32) and out[0][1] == i32          # auto = fp32 at krylov_rtol 1e-5
    # tight linear tolerance: auto falls back to the fp64 basis
    res = {}
    for mode in (1, 0):
        ctx = engine.Context("b13d", (24, 20, 18), (48.0, 40.0, 36.0))
        ctx.set_option("basis_precision", mode)
        w0 = ctx.initial_condition()
        w = w0.clone()
        ctx.dt = 0.1
        rep = ctx.newton_solve(w, w0, absolute_tolerance=1e-10, krylov_rtol=1e-9, maximum_iterations=25)
        assert rep["converged"]
        res[mode] = w.cpu().numpy()
        ctx.close()
    assert np.array_equal(res[0], res[1])
