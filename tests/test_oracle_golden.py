"""Pins the oracle (oracle/) to the reference's own committed outputs (tests/golden/*.npz, made
by oracle/make_golden.py from /root/reference/results).  CPU only."""
import numpy as np
import pytest

from helpers import REFERENCE_MESH, discretization, lu_once, rel_l2
from oracle import mesh as omesh


def test_crossed_mesh_matches_golden_vtu(golden):
    g = golden["bench1_mesh"]
    m = omesh.crossed2d(100, 100, 200.0, 200.0)
    assert np.array_equal(m.cells, g["cells"])                     # DOLFIN cell/vertex order
    assert np.abs(m.coords - g["points"]).max() < 1e-6


def test_bench1_energy_and_solute_of_golden_snapshots(golden):
    D = discretization("bench1", *REFERENCE_MESH["bench1"])
    csv, snaps = golden["csv"]["bench1"], golden["bench1_fields"]
    tol = [1e-11, 1e-11, 5e-11, 5e-11, 1e-9, 1e-8]   # VTUs and CSV come from two runs
    for k in range(6):
        u = np.zeros((2, D.nv)); u[0] = snaps["c"][k]
        assert abs(snaps["times"][k] - csv[k, 0]) < 1e-9
        assert abs(D.energy(u) / csv[k, 1] - 1) < tol[k]
        assert abs(D.mass(u) / csv[k, 2] - 1) < 1e-11


def test_bench1_step_kats(golden):
    """snapshot k -> k+1 with dt from the PVD times: the oracle reproduces FEniCS' fields."""
    D = discretization("bench1", *REFERENCE_MESH["bench1"])
    snaps = golden["bench1_fields"]
    u0 = D.prob.ic(D.mesh)
    u, n, ok = D.solve_step(u0, u0, 0.1)
    assert ok and rel_l2(u[0], snaps["c"][0]) < 1e-10
    for k in range(5):
        dt = snaps["times"][k + 1] - snaps["times"][k]
        u0 = np.zeros((2, D.nv)); u0[0] = snaps["c"][k]
        u, n, ok = D.solve_step(u0, u0, dt)
        assert ok and rel_l2(u[0], snaps["c"][k + 1]) < 2e-9, (k, rel_l2(u[0], snaps["c"][k + 1]))


def test_bench1_trajectory_energy_first_rows(golden):
    """Replay the CSV's dt list from the IC: free energy within the north-star 1e-6 (we get ~1e-9)."""
    D = discretization("bench1", *REFERENCE_MESH["bench1"])
    csv = golden["csv"]["bench1"]
    u = D.prob.ic(D.mesh)
    t = 0.0
    for k in range(8):
        dt = csv[k, 0] - t
        u, n, ok = D.solve_step(u, u, dt)
        t = csv[k, 0]
        assert ok
        assert abs(D.energy(u) / csv[k, 1] - 1) < 2e-8, k
        assert abs(D.mass(u) / csv[k, 2] - 1) < 1e-10


def test_bench2_ic_quirk_and_step_kat(golden):
    D = discretization("bench2", *REFERENCE_MESH["bench2"])
    g = golden["bench2_fields"]
    csv = golden["csv"]["bench2"]
    rows = [0, 2, 3, 4, 5]
    s0 = np.zeros((6, D.nv)); s0[rows] = g["fields"][0]
    s1 = np.zeros((6, D.nv)); s1[rows] = g["fields"][1]
    assert abs(D.energy(s0) / csv[0, 1] - 1) < 1e-10 and abs(D.mass(s0) / csv[0, 2] - 1) < 1e-10
    u0 = D.prob.ic(D.mesh)
    u, n, ok = D.solve_step(u0, u0, 0.01, linear=lu_once())            # IC -> snapshot 0 pins the i/ii quirk (pfbase.py:285-286)
    assert ok
    for r in rows:
        assert rel_l2(u[r], s0[r]) < 1e-9
    u, n, ok = D.solve_step(s0, s0, g["times"][1] - g["times"][0], linear=lu_once())
    assert ok
    for r in rows:
        assert rel_l2(u[r], s1[r]) < 1e-11


def test_bench6_step_kat(golden):
    D = discretization("bench6", *REFERENCE_MESH["bench6"])
    g = golden["bench6_fields"]
    csv = golden["csv"]["bench6"]
    u0 = D.prob.ic(D.mesh)
    u, n, ok = D.solve_step(u0, u0, 0.01)
    assert ok and rel_l2(u[0], g["c"][0]) < 1e-9 and rel_l2(u[2], g["phi"][0]) < 1e-6
    assert abs(D.mass(u) / csv[0, 2] - 1) < 1e-10 and abs(D.energy(u) / csv[0, 1] - 1) < 1e-6
    a = np.zeros((3, D.nv)); a[0] = g["c"][1]; a[2] = g["phi"][1]
    u, n, ok = D.solve_step(a, a, g["times"][2] - g["times"][1])
    assert ok and rel_l2(u[0], g["c"][2]) < 1e-9 and rel_l2(u[2], g["phi"][2]) < 1e-6


@pytest.mark.slow
def test_bench3_first_csv_row(golden):
    D = discretization("bench3", *REFERENCE_MESH["bench3"])
    csv = golden["csv"]["bench3"]
    u0 = D.prob.ic(D.mesh)
    u, n, ok = D.solve_step(u0, u0, csv[0, 0], linear=lu_once())     # one 491k x 491k sparse LU instead of one per Newton iteration
    assert ok
    assert abs(D.energy(u) / csv[0, 1] - 1) < 1e-11
    assert abs(D.mass(u) - csv[0, 2]) < 1e-10


def test_b13d_trajectory_fixtures_are_consistent():
    """tests/golden/b13d_{51,52}.npz (oracle/make_golden_b13d.py; the GPU suite compares the product with them): 100 steps
    of the section-8(d) dt list, solute conserved to round-off, free energy decaying monotonically, fields of the right size
    and within the double-well range."""
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    for nv in (51, 52):
        g = np.load(os.path.join(here, "golden", "b13d_%d.npz" % nv))
        n = int(g["ncell"])
        assert n + 1 == nv and int(g["nsteps"]) == 100 and float(g["length"]) == 2.0 * n
        dts = np.diff(np.concatenate([[0.0], g["time"]]))
        assert np.allclose(dts[:5], [0.1, 0.2, 0.4, 0.8, 1.6]) and np.allclose(dts[5:], 1.6)
        assert np.all(np.diff(g["energy"]) < 0.0)
        assert np.max(np.abs(g["solute"] / g["solute"][0] - 1.0)) < 1e-13
        for key in ("c_step10", "c_final"):
            assert g[key].shape == (nv ** 3,) and 0.2 < g[key].min() < g[key].max() < 0.8
