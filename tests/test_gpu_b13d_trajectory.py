"""North-star horizon on the configuration that is benchmarked: 100 backward-Euler steps of b13d (dolfin/b13d.py) with
bench.py's EXACT solver options, against the oracle's tightly converged trajectory (tests/golden/b13d_*.npz, generated
by oracle/make_golden_b13d.py).  51^3 vertices is the reference mesh (b13d.py:24-26; odd vertex count: cp.async loader of
the staged kernels), 52^3 takes the TMA loader.  The reference commits no 3-D output: this pins product to oracle.

Two runs per mesh:
  * bench options (Newton atol 1e-6 as the drivers set it, GMRES rtol 1e-5, fp32-stored Krylov basis, Schur/Chebyshev
    3/4 upper): the north star's bars, relative L2 of the concentration field <= 1e-8 and free energy <= 1e-6 relative,
    over all 100 steps (measured on a B200: 1.5e-10 / 5.5e-10 on the field, 3e-10 on the energy).
  * tight options (atol 1e-10, GMRES rtol 1e-9, fp64 basis): fields <= 1e-9, energy <= 1e-9 (measured 5e-11 / 2e-11).
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def dt_schedule(n):
    return [min(0.1 * 2 ** k, 1.6) for k in range(n)]


def _run(nv, opts_kw, basis):
    from pfhubbenchmarks_b200 import _lib, engine
    g = np.load(os.path.join(GOLDEN, "b13d_%d.npz" % nv))
    n, L, nsteps = int(g["ncell"]), float(g["length"]), int(g["nsteps"])
    assert n + 1 == nv
    ctx = engine.Context("b13d", (n, n, n), (L, L, L))
    ctx.set_option("basis_precision", basis)
    opts = _lib.default_opts(**opts_kw)
    w = ctx.initial_condition()
    w0 = w.clone()
    worst_e, worst_c, errs, newton, krylov = 0.0, 0.0, {}, 0, 0
    for k, dt in enumerate(dt_schedule(nsteps)):
        w0.copy_(w)
        ctx.dt = dt
        rep = ctx.newton_solve(w, w0, opts=opts)
        assert rep["converged"], (k, rep)
        newton += rep["niters"]; krylov += rep["krylov_its"]
        E, C = ctx.energy(w)
        worst_e = max(worst_e, abs(E / g["energy"][k] - 1.0))
        worst_c = max(worst_c, abs(C / g["solute"][k] - 1.0))
        if k + 1 in (10, nsteps):
            ref = g["c_step10"] if k + 1 == 10 else g["c_final"]
            c = w[:ctx.n_owned].cpu().numpy()
            errs[k + 1] = float(np.linalg.norm(c - ref) / np.linalg.norm(ref))
    launches = ctx.launch_count()
    ctx.close()
    return worst_e, worst_c, errs, newton, krylov, launches


@pytest.mark.parametrize("nv", [51, 52])
def test_b13d_100_steps_bench_options(nv):
    n = nv - 1
    # exactly what bench.py passes (run_b200 / bench_case): restart 30, upper-triangular Schur variant, 3/4 sweeps
    kw = dict(krylov_restart=30, pc_variant=1, pc_sweeps_mass=3, pc_sweeps_shifted=4, krylov_maximum_iterations=max(1000, int(n ** 3 / 10)))
    e, c, errs, newton, krylov, launches = _run(nv, kw, 0)
    print("b13d %d^3, bench options, 100 steps: worst energy error %.2e, solute %.2e, rel L2(c) step 10 %.2e, step 100 %.2e; "
          "%d Newton / %d Krylov iterations, %d launches" % (nv, e, c, errs[10], errs[100], newton, krylov, launches))
    assert e < 1e-6 and c < 1e-11                     # the north-star bar on the free energy; solute conserved
    assert errs[10] < 1e-8 and errs[100] < 1e-8       # the north-star bar on the fields


@pytest.mark.parametrize("nv", [51, 52])
def test_b13d_100_steps_tight_options(nv):
    kw = dict(krylov_restart=30, pc_variant=1, pc_sweeps_mass=3, pc_sweeps_shifted=4, krylov_maximum_iterations=5000,
              absolute_tolerance=1e-10, krylov_rtol=1e-9, maximum_iterations=25)
    e, c, errs, newton, krylov, launches = _run(nv, kw, 1)
    print("b13d %d^3, tight options, 100 steps: worst energy error %.2e, solute %.2e, rel L2(c) step 10 %.2e, step 100 %.2e; "
          "%d Newton / %d Krylov iterations" % (nv, e, c, errs[10], errs[100], newton, krylov))
    assert e < 1e-9 and c < 1e-11
    assert errs[10] < 1e-9 and errs[100] < 1e-9
