"""GPU test of a generated model (codegen.py): the weak form of drivers/custom_model.py is not one of the shipped
systems, so the problem constructor emits a functor, builds its own copy of the library and runs on it; the steps are
compared with the oracle's assembled-operator Newton/LU solve of the same model."""
import numpy as np
import pytest

from oracle import mesh as omesh, stepper
from test_codegen import OracleModel

pytestmark = pytest.mark.gpu


def test_generated_model_steps_match_oracle():
    import pfhubbenchmarks_b200.dolfin as df
    from pfhubbenchmarks_b200 import engine
    from pfhubbenchmarks_b200.drivers import custom_model as cm
    from pfhubbenchmarks_b200.pfbase import grad, inner, dx
    ncell, length = (24, 20), (48.0, 40.0)
    mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(*length), *ncell, "crossed")
    problem, solver, total_free_energy, total_solute = cm.setup(mesh)
    assert problem.model is not None and problem.benchmark.startswith("user_") and problem.ctx.nf == 3
    w, w0, dt = problem.w, problem.w0, problem.dt
    w.interpolate(cm.InitialConditions(degree=2))
    nl = solver.parameters["snes_solver"]
    nl["absolute_tolerance"], nl["maximum_iterations"] = 1e-10, 25
    nl["krylov_solver"]["relative_tolerance"] = 1e-9
    nl["krylov_solver"]["maximum_iterations"] = 5000
    om = omesh.crossed2d(*ncell, *length)
    D = stepper.Discretization(om, OracleModel())
    u = w.to_numpy().copy()
    X = om.coords
    vals = np.zeros((om.nv, 3))
    ic = cm.InitialConditions(degree=2)
    for i in range(om.nv):
        ic.eval(vals[i], X[i])
    assert np.array_equal(u, vals.T)                       # host-evaluated UserExpression, library vertex order
    for step, h in enumerate((0.05, 0.1, 0.2)):
        dt.assign(h)
        w0.assign(w)
        niters, converged = solver.solve()
        assert converged, solver.last_report
        u, n, ok = D.solve_step(u, u, h, atol=1e-10)
        got = w.to_numpy()
        assert ok
        for f in range(3):
            err = np.linalg.norm(got[f] - u[f]) / np.linalg.norm(u[f])
            assert err < 1e-8, (step, f, err)
        E, S = total_free_energy(), total_solute()
        assert abs(E / D.energy(u) - 1) < 1e-9 and abs(S / D.mass(u) - 1) < 1e-12, (E, D.energy(u), S, D.mass(u))
    # a piece of the energy on its own goes through pfh_bilinear_dev of the generated library
    eta = w[2]
    piece = df.assemble(cm.KAPPA_ETA / 2.0 * inner(grad(eta), grad(eta)) * dx)
    assert abs(piece - cm.KAPPA_ETA / 2.0 * float(u[2] @ (D.V.K @ u[2]))) < 1e-9 * abs(piece)
    # the default tolerances (compressed Krylov basis, reaction-shifted Allen-Cahn block) converge as well
    dt.assign(0.2)
    w0.assign(w)
    nl["absolute_tolerance"], nl["maximum_iterations"] = 1e-6, 10
    nl["krylov_solver"]["relative_tolerance"] = 1e-5
    niters, converged = solver.solve()
    assert converged and niters <= 6, solver.last_report
    # the stock library is untouched by the generated copy living in the same process
    ctx = engine.Context("bench1", (16, 12), (32.0, 24.0))
    v = ctx.initial_condition()
    ctx.dt = 0.1
    assert ctx.newton_solve(v.clone(), v)["converged"]
    ctx.close()
    problem.ctx.close()


def test_shipped_initial_condition_on_a_generated_model():
    """A reference _IC class handed to a generated model is interpolated on the host from its closed form (the generated
    kernel has no initial condition of its own); on a shipped problem of ANOTHER benchmark it is evaluated the same way
    instead of silently running that problem's kernel formula; a component-count mismatch raises."""
    import pfhubbenchmarks_b200.dolfin as df
    from pfhubbenchmarks_b200 import api
    from pfhubbenchmarks_b200.drivers import custom_model as cm
    from pfhubbenchmarks_b200.pfbase import InitialConditionsBench1, InitialConditionsBench6
    ncell, length = (12, 10), (24.0, 20.0)
    mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(*length), *ncell, "crossed")
    problem, solver, _, _ = cm.setup(mesh)                                  # three fields, generated
    ic6 = InitialConditionsBench6(0.5, 0.04, degree=2)
    problem.w.interpolate(ic6)
    xy = omesh.crossed2d(*ncell, *length).coords
    got = problem.w.to_numpy()
    assert np.array_equal(got, ic6.values(xy)) and np.abs(got[0] - 0.5).max() > 0.01
    with pytest.raises(ValueError):
        problem.w.interpolate(InitialConditionsBench1(0.5, 0.05, degree=2))  # two components into three fields
    problem.ctx.close()
    # thin API: benchmark 1's expression on a bench1 problem runs the kernel; benchmark 6's does not fit two fields
    p1 = api.NonlinearVariationalProblem("bench1", api.RectangleMesh(api.Point(0.0, 0.0), api.Point(*length), *ncell, "crossed"), api.Constant(0.1))
    ic1 = InitialConditionsBench1(0.5, 0.05)
    p1.w.interpolate(ic1)
    assert np.allclose(p1.w.to_numpy(), ic1.values(xy), rtol=0, atol=1e-14)
    with pytest.raises(ValueError):
        p1.w.interpolate(ic6)
    p1.ctx.close()


def test_reduced_ostwald_driver(tmp_path):
    """`python -m pfhubbenchmarks_b200.drivers.bench2_eta --num-eta 1`: the reference's bench2_eta1 model end to end as a
    generated model with the host-evaluated ICB2_jank1 initial data: converges, solute conserved, energy decays."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pfhubbenchmarks_b200.drivers.bench2_eta", "--num-eta", "1", "--max-steps", "4", "--quiet",
                        "--outdir", str(tmp_path)], cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = np.loadtxt(str(tmp_path / "bench2_eta1_out.csv"), delimiter=",", skiprows=1)
    assert rows.shape == (4, 3) and np.all(np.diff(rows[:, 0]) > 0)
    assert np.all(np.diff(rows[:, 1]) < 0) and np.allclose(rows[:, 2], rows[0, 2], rtol=1e-9)
