"""GPU parity for the other configs of BASELINE.json (bench2, bench3, bench6) and for the driver loop:
CUDA Newton steps against the reference's committed FEniCS snapshots / CSV rows (tests/golden)."""
import os

import numpy as np
import pytest

from helpers import REFERENCE_MESH, rel_l2

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

TIGHT = dict(absolute_tolerance=1e-10, krylov_rtol=1e-9, maximum_iterations=25, krylov_maximum_iterations=5000)


def _ctx(name):
    from pfhubbenchmarks_b200 import engine
    ncell, length = REFERENCE_MESH[name]
    return engine.Context(name, ncell, length)


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).ravel()).cuda()


def test_bench2_golden_step_kats(golden):
    """IC -> snapshot 0 (pins the eta-IC i/ii quirk, pfbase.py:285-286) and snapshot 0 -> 1, six fields."""
    g, csv = golden["bench2_fields"], golden["csv"]["bench2"]
    ctx = _ctx("bench2")
    nv = ctx.n_owned
    rows = [0, 2, 3, 4, 5]
    w0 = ctx.initial_condition(); w = w0.clone()
    ctx.dt = 0.01
    rep = ctx.newton_solve(w, w0, **TIGHT)
    assert rep["converged"], rep
    got = w.cpu().numpy().reshape(6, nv)
    for j, r in enumerate(rows):
        assert rel_l2(got[r], g["fields"][0][j]) < 1e-9, (r, rep)
    E, M = ctx.energy(w)
    assert abs(E / csv[0, 1] - 1) < 1e-9 and abs(M / csv[0, 2] - 1) < 1e-10
    s0 = np.zeros((6, nv)); s0[rows] = g["fields"][0]
    w0 = to_dev(s0); w = w0.clone()
    ctx.dt = float(g["times"][1] - g["times"][0])
    rep = ctx.newton_solve(w, w0, **TIGHT)
    assert rep["converged"], rep
    got = w.cpu().numpy().reshape(6, nv)
    for j, r in enumerate(rows):
        assert rel_l2(got[r], g["fields"][1][j]) < 1e-9, (r, rep)
    ctx.close()


def test_bench6_golden_step_kats(golden):
    """IC -> snapshot 0 and snapshot 1 -> 2; phi is limited by the reference's own Krylov tolerance."""
    g, csv = golden["bench6_fields"], golden["csv"]["bench6"]
    ctx = _ctx("bench6")
    nv = ctx.n_owned
    w0 = ctx.initial_condition(); w = w0.clone()
    ctx.dt = 0.01
    rep = ctx.newton_solve(w, w0, **TIGHT)
    assert rep["converged"], rep
    got = w.cpu().numpy().reshape(3, nv)
    assert rel_l2(got[0], g["c"][0]) < 1e-9 and rel_l2(got[2], g["phi"][0]) < 1e-6, rep
    E, M = ctx.energy(w)
    assert abs(M / csv[0, 2] - 1) < 1e-10 and abs(E / csv[0, 1] - 1) < 1e-6
    a = np.zeros((3, nv)); a[0] = g["c"][1]; a[2] = g["phi"][1]
    w0 = to_dev(a); w = w0.clone()
    ctx.dt = float(g["times"][2] - g["times"][1])
    rep = ctx.newton_solve(w, w0, **TIGHT)
    assert rep["converged"], rep
    got = w.cpu().numpy().reshape(3, nv)
    assert rel_l2(got[0], g["c"][2]) < 1e-9 and rel_l2(got[2], g["phi"][2]) < 1e-6, rep
    ctx.close()


@pytest.mark.parametrize("n", [64, 100, 512])
def test_bench6_multigrid_poisson_block(n):
    """The preconditioner's Poisson block by one geometric-multigrid V-cycle (default) against the 240 Chebyshev sweeps it
    replaces ("mg_cycles" 0): same Newton iterate, no more Krylov iterations -- and fewer where 240 sweeps stop resolving the
    smooth modes (512^2).  The CPU leg (tests/test_emul_solver.py) pins the cycle itself: transfers = P1 interpolation and its
    transpose, contraction <= 0.06 per cycle; here the device launch layer (level meshes, tables, work vectors) is what is new."""
    from pfhubbenchmarks_b200 import engine
    out = {}
    for cycles in (1, 0, 2):
        ctx = engine.Context("bench6", (n, n), (float(n), float(n)))
        ctx.set_option("mg_cycles", cycles)
        w0 = ctx.initial_condition(); w = w0.clone()
        ctx.dt = 0.01
        rep = ctx.newton_solve(w, w0, absolute_tolerance=1e-9, maximum_iterations=10, krylov_rtol=1e-8, line_search=1,
                               krylov_maximum_iterations=5000)
        assert rep["converged"], (cycles, rep)
        out[cycles] = (w.cpu().numpy().reshape(3, -1), rep)
        ctx.close()
    for cycles in (1, 2):
        for f in range(3):
            assert rel_l2(out[cycles][0][f], out[0][0][f]) < 1e-6 or np.abs(out[0][0][f]).max() < 1e-12, (cycles, f)
        assert out[cycles][1]["niters"] == out[0][1]["niters"]
        assert out[cycles][1]["krylov_its"] <= 1.1 * out[0][1]["krylov_its"], (n, cycles, out[cycles][1], out[0][1])


def test_bench3_first_csv_rows(golden):
    """bench3 has no committed fields; the CSV pins energy and solid fraction of the first steps."""
    csv = golden["csv"]["bench3"]
    ctx = _ctx("bench3")
    w = ctx.initial_condition()
    t = 0.0
    for k in range(3):
        dt = float(csv[k, 0] - t); t = float(csv[k, 0])
        w0 = w.clone()
        ctx.dt = dt
        rep = ctx.newton_solve(w, w0, **TIGHT)
        assert rep["converged"], rep
        E, S = ctx.energy(w)
        assert abs(E / csv[k, 1] - 1) < 1e-9, (k, E, csv[k, 1])
        assert abs(S - csv[k, 2]) < 5e-10, (k, S, csv[k, 2])
    ctx.close()


def test_bench1_driver_reproduces_reference_csv(golden, tmp_path):
    """The driver loop (adaptive dt, retry, CSV writer) with the reference's own solver settings: rows whose
    time matches the committed results/bench1_out.csv agree in free energy (1e-6) and solute (1e-9)."""
    from pfhubbenchmarks_b200 import api as df
    from pfhubbenchmarks_b200.drivers import common
    csv = golden["csv"]["bench1"]
    mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(200.0, 200.0), 100, 100, "crossed")
    dt = df.Constant(1e-1)
    problem = df.NonlinearVariationalProblem("bench1", mesh, dt)
    solver = df.NonlinearVariationalSolver(problem)
    nl = solver.parameters["snes_solver"]
    nl["absolute_tolerance"] = 1e-6
    nl["maximum_iterations"] = 10
    nl["line_search"] = "basic"
    nl["linear_solver"] = "gmres"
    nl["preconditioner"] = "sor"
    nl["krylov_solver"]["maximum_iterations"] = 1000
    out = str(tmp_path / "stats.csv")
    rows, spent = common.time_loop(problem, solver, df.InitialConditionsBench1(0.5, 0.05, degree=2), 1e-1, 1e-2, 1e3, out,
                                   "time,total_free_energy,total_solute", max_steps=14, quiet=True)
    rows = np.array(rows)
    assert open(out).readline().strip() == "time,total_free_energy,total_solute"
    back = np.loadtxt(out, delimiter=",", skiprows=1)
    assert back.shape == rows.shape and np.allclose(back, rows, rtol=0, atol=1e-9)
    # the dt sequence depends on the nonlinear solver's iteration counts, so only the common
    # prefix of accepted times is comparable row by row
    prefix = 0
    while prefix < len(rows) and abs(rows[prefix, 0] - csv[prefix, 0]) < 1e-9:
        prefix += 1
    assert prefix >= 4, (prefix, rows[:, 0], csv[:14, 0])
    for k in range(prefix):
        assert abs(rows[k, 1] / csv[k, 1] - 1) < 1e-6 and abs(rows[k, 2] / csv[k, 2] - 1) < 1e-9, (k, rows[k], csv[k])
    print("driver rows matching the reference CSV time for time:", prefix, "of", len(rows))
    assert np.all(np.diff(rows[:, 1]) < 0)                     # energy decays
    problem.ctx.close()


def test_driver_cli_writes_csv_and_fields(tmp_path):
    """`python -m pfhubbenchmarks_b200.drivers.bench1 --max-steps 3 --save-solution`: CSV + PVD/VTU come out, and the
    dumped concentration of step 1 equals the FEniCS snapshot results/bench1/conc000000.vtu to solver tolerance."""
    import subprocess
    import sys
    from oracle import golden_io
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pfhubbenchmarks_b200.drivers.bench1", "--max-steps", "3", "--save-solution",
                        "--outdir", str(tmp_path), "--quiet", "--atol", "1e-10"], cwd=root, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = np.loadtxt(str(tmp_path / "bench1" / "stats.csv"), delimiter=",", skiprows=1)
    assert rows.shape == (3, 3) and np.allclose(rows[:, 0], [0.1, 0.3, 0.7])
    entries = golden_io.read_pvd(str(tmp_path / "bench1" / "conc.pvd"))
    assert len(entries) == 3
    c1 = golden_io.read_vtu_scalar(str(tmp_path / "bench1" / entries[0][1]))
    g = np.load(os.path.join(root, "tests", "golden", "bench1_fields.npz"))
    assert rel_l2(c1, g["c"][0]) < 1e-9


@pytest.mark.parametrize("name,csvfile,extra", [("bench2", "bench2_out.csv", []), ("bench3", "bench3_out.csv", []),
                                                ("bench6", "bench6_out.csv", []), ("b13d", "bench1_out.csv", ["--n", "16"]),
                                                ("b63d", "bench6_out.csv", [])])     # b63d.py = bench6.py in the reference
def test_other_driver_clis(golden, tmp_path, name, csvfile, extra):
    """Each driver module end to end (UFL-style model setup -> recognised form -> time loop -> CSV): the first accepted
    steps reproduce the committed reference CSV rows (b13d has no reference output: energy decays, solute conserved)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pfhubbenchmarks_b200.drivers." + name, "--max-steps", "3", "--outdir", str(tmp_path),
                        "--quiet"] + extra, cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = np.loadtxt(str(tmp_path / csvfile), delimiter=",", skiprows=1)
    assert rows.shape == (3, 3)
    if name == "b13d":
        assert np.all(np.diff(rows[:, 1]) < 0) and np.allclose(rows[:, 2], rows[0, 2], rtol=1e-9)
        return
    csv = golden["csv"]["bench6" if name == "b63d" else name]
    for k in range(3):
        assert abs(rows[k, 0] - csv[k, 0]) < 1e-9, (rows[:, 0], csv[:3, 0])
        assert abs(rows[k, 1] - csv[k, 1]) <= 2e-6 * abs(csv[k, 1]) and abs(rows[k, 2] - csv[k, 2]) <= 1e-8 * abs(csv[k, 2]), (rows[k], csv[k])


@pytest.mark.parametrize("name,ic,ncsv", [("bench2", (0.5, 0.05, 0.1, 1.5), 40), ("bench3", (-0.3, 8.0, 1.0, 1.0, -1.0), 46),
                                          ("bench6", (0.5, 0.04), 14)])
def test_other_drivers_follow_reference_csv(golden, tmp_path, name, ic, ncsv):
    """Driver loop with the reference's solver settings (atol 1e-6, 10 its, dt0 = 1e-2, niters < 5 doubling): the first
    rows reproduce results/<name>_out.csv (same accepted times; energy to 1e-6, second column to 1e-9 / 1e-6 abs)."""
    from pfhubbenchmarks_b200 import api as df
    from pfhubbenchmarks_b200.drivers import common
    csv = golden["csv"][name]
    ncell, length = REFERENCE_MESH[name]
    mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(*length), ncell[0], ncell[1], "crossed")
    dt = df.Constant(1e-2)
    problem = df.NonlinearVariationalProblem(name, mesh, dt)
    solver = df.NonlinearVariationalSolver(problem)
    nl = solver.parameters["snes_solver"]
    nl["absolute_tolerance"] = 1e-6
    nl["maximum_iterations"] = 10
    nl["krylov_solver"]["maximum_iterations"] = 5000 if name == "bench6" else 1000
    nl["line_search"] = "basic" if name == "bench3" else "cp"        # bench3.py:124 / bench2.py:140, bench6.py:117
    ic_cls = {"bench2": df.InitialConditionsBench2, "bench3": df.InitialConditionsBench3, "bench6": df.InitialConditionsBench6}[name]
    hdr = "time,total_free_energy,solid_fraction" if name == "bench3" else "time,total_free_energy,total_solute"
    rows, _ = common.time_loop(problem, solver, ic_cls(*ic, degree=2), 1e-2, 1e-4, 100.0, str(tmp_path / (name + "_out.csv")), hdr,
                               max_steps=ncsv, quiet=True)
    rows = np.array(rows)
    prefix = 0
    while prefix < len(rows) and abs(rows[prefix, 0] - csv[prefix, 0]) < 1e-9:
        prefix += 1
    print("%s: %d of %d driver rows match the reference CSV time for time" % (name, prefix, len(rows)))
    # bench3's whole committed run (46 rows, t = 0.01 .. 103) and the first 40 rows of bench2 (incl. the reference's own
    # rejected dt = 0.16 at t = 0.15) are reproduced time for time; bench6's dt sequence depends on the nonlinear
    # solver's iteration counts after the first rows
    assert prefix >= {"bench3": 46, "bench2": 30}.get(name, 3), (rows[:, 0], csv[:ncsv, 0])
    etol = 2e-6 if name == "bench6" else 1e-6          # bench6's committed CSV is itself only converged to ~1e-6
    for k in range(prefix):
        assert abs(rows[k, 1] / csv[k, 1] - 1) < etol, (k, rows[k], csv[k])
        if name == "bench3":
            assert abs(rows[k, 2] - csv[k, 2]) < 1e-8
        else:
            assert abs(rows[k, 2] / csv[k, 2] - 1) < 1e-8
    problem.ctx.close()


def test_bench1_full_length_run_statistics(golden, tmp_path):
    """The complete bench1 run (end_time = 1e3, dolfin/bench1.py:138) through the driver loop.  Long spinodal runs are
    compared on domain-averaged statistics (north star): the run completes, the free energy decays monotonically, the
    solute is conserved to round-off, and the energy-vs-time curve stays within 2 % of results/bench1_out.csv."""
    from pfhubbenchmarks_b200 import api as df
    from pfhubbenchmarks_b200.drivers import common
    g = golden["csv"]["bench1"]
    mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(200.0, 200.0), 100, 100, "crossed")
    dt = df.Constant(1e-1)
    problem = df.NonlinearVariationalProblem("bench1", mesh, dt)
    solver = df.NonlinearVariationalSolver(problem)
    nl = solver.parameters["snes_solver"]
    nl["absolute_tolerance"] = 1e-6
    nl["maximum_iterations"] = 10
    nl["krylov_solver"]["maximum_iterations"] = 1000
    rows, spent = common.time_loop(problem, solver, df.InitialConditionsBench1(0.5, 0.05, degree=2), 1e-1, 1e-2, 1e3,
                                   str(tmp_path / "stats.csv"), "time,total_free_energy,total_solute", quiet=True)
    rows = np.array(rows)
    assert rows[-1, 0] >= 1e3 and len(rows) < 200
    assert np.all(np.diff(rows[:, 1]) < 0)
    assert abs(rows[-1, 2] / rows[0, 2] - 1) < 1e-12
    Ei = np.interp(np.log(g[:, 0]), np.log(rows[:, 0]), rows[:, 1])
    sel = g[:, 0] <= rows[-1, 0]
    assert np.max(np.abs(Ei[sel] / g[sel, 1] - 1)) < 2e-2
    print("bench1 to t = %.0f: %d steps in %.1f s; energy curve within %.2e of the reference" % (
        rows[-1, 0], len(rows), spent, np.max(np.abs(Ei[sel] / g[sel, 1] - 1))))
    problem.ctx.close()
