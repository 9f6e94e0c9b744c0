"""Field output (VTU/PVD, "next" row N2): round trip through the golden readers and mesh identity with DOLFIN's files."""
import os

import numpy as np
import pytest

from oracle import golden_io, mesh as omesh
from pfhubbenchmarks_b200 import api, io


def test_meshes_match_dolfin_layout(golden):
    pts, cells = io.crossed_mesh(100, 100, 200.0, 200.0)
    g = golden["bench1_mesh"]
    assert np.array_equal(cells, g["cells"]) and np.abs(pts[:, :2] - g["points"]).max() < 1e-6
    m = omesh.kuhn3d(3, 4, 2, 3.0, 8.0, 1.0)
    pts, cells = io.box_mesh(3, 4, 2, 3.0, 8.0, 1.0)
    assert np.array_equal(cells, m.cells) and np.allclose(pts, m.coords)


@pytest.mark.parametrize("mode", ["ascii", "compressed"])
def test_pvd_vtu_round_trip(tmp_path, mode):
    rng = np.random.default_rng(0)
    mesh = api.RectangleMesh(api.Point(0.0, 0.0), api.Point(8.0, 6.0), 4, 3, "crossed")
    f = io.File(str(tmp_path / "conc.pvd"), mode, mesh=mesh, name="f_13-0")
    vals = [rng.standard_normal(5 * 4 + 4 * 3) for _ in range(3)]
    for k, v in enumerate(vals):
        f << (v, 0.1 * (k + 1))
    entries = golden_io.read_pvd(str(tmp_path / "conc.pvd"))
    assert [e[1] for e in entries] == ["conc000000.vtu", "conc000001.vtu", "conc000002.vtu"]
    assert np.allclose([e[0] for e in entries], [0.1, 0.2, 0.3], rtol=0, atol=1e-15)
    pts, cells = io.mesh_of(mesh)
    for (t, fn), v in zip(entries, vals):
        d = golden_io.read_vtu(str(tmp_path / fn))
        assert np.array_equal(d["cells"], cells) and np.allclose(d["points"], pts, rtol=0, atol=1e-15)
        got = d["data"]["f_13-0"]
        assert np.array_equal(got, v)          # 17 significant digits / raw bytes: exact round trip


def test_large_binary_array_uses_several_zlib_blocks(tmp_path):
    pts, cells = io.box_mesh(12, 11, 10, 12.0, 11.0, 10.0)
    v = np.random.default_rng(1).standard_normal(pts.shape[0])
    io.write_vtu(str(tmp_path / "a.vtu"), pts, cells, v, name="c", binary=True)
    d = golden_io.read_vtu(str(tmp_path / "a.vtu"))
    assert np.array_equal(d["data"]["c"], v) and np.array_equal(d["cells"], cells) and np.array_equal(d["points"], pts)


@pytest.mark.parametrize("mode", ["ascii", "compressed"])
def test_product_readers_round_trip(tmp_path, mode):
    """io.read_vtu / read_pvd / read_series (the PVD/VTU counterpart of process_bench1.py's HDF5 reading) against the
    writer and against the oracle's independent reader; large enough for several zlib blocks in compressed mode."""
    from oracle import golden_io
    mesh = api.RectangleMesh(api.Point(0.0, 0.0), api.Point(40.0, 30.0), 40, 30, "crossed")
    n = 41 * 31 + 40 * 30
    rng = np.random.default_rng(0)
    f = io.File(str(tmp_path / "out" / "conc.pvd"), mode, mesh=mesh, name="c")
    fields = [rng.standard_normal(n) for _ in range(3)]
    for k, v in enumerate(fields):
        f << (v, 0.1 * (k + 1))
    times, vals = io.read_series(str(tmp_path / "out" / "conc.pvd"))
    assert np.allclose(times, [0.1, 0.2, 0.3]) and vals.shape == (3, n)
    assert all(np.array_equal(vals[k], fields[k]) for k in range(3))
    pts, cells, data = io.read_vtu(io.read_pvd(str(tmp_path / "out" / "conc.pvd"))[1][1])
    ref = golden_io.read_vtu(io.read_pvd(str(tmp_path / "out" / "conc.pvd"))[1][1])
    assert np.array_equal(pts, ref["points"]) and np.array_equal(cells, ref["cells"]) and list(data) == ["c"]
    gp, gc = io.crossed_mesh(40, 30, 40.0, 30.0)
    assert np.array_equal(pts[:, :2], gp[:, :2]) and np.array_equal(cells, gc)


@pytest.mark.skipif(not os.path.exists("/root/reference/results/bench1/conc.pvd"), reason="reference checkout not present on this box")
def test_product_reader_on_the_committed_reference_series(golden):
    """The reference's own output files (zlib/base64 VTU for bench1, ASCII for bench2) through the product reader:
    identical to the golden arrays that were extracted with the oracle's reader."""
    t, v = io.read_series("/root/reference/results/bench1/conc.pvd")
    assert np.allclose(t, [0.1, 0.3, 0.7, 1.5, 3.1, 4.7]) and np.array_equal(v, golden["bench1_fields"]["c"])
    t2, v2 = io.read_series("/root/reference/results/bench2/conc.pvd")
    g2 = golden["bench2_fields"]                      # a subset of the 29 snapshots: index, times, fields[k, (c, eta1..4), :]
    assert v2.shape == (29, 20201)
    for k, idx in enumerate(g2["index"]):
        assert t2[int(idx)] == g2["times"][k] and np.array_equal(v2[int(idx)], g2["fields"][k, 0])


@pytest.mark.parametrize("kind,ncell,length", [("crossed2d", (6, 4), (12.0, 8.0)), ("crossed2d", (1, 3), (2.0, 6.0)),
                                               ("right2d", (5, 7), (1.0, 1.0)), ("kuhn3d", (3, 4, 2), (6.0, 8.0, 4.0))])
def test_stored_series_is_read_back_like_the_reference_post_processing(tmp_path, kind, ncell, length):
    """process_bench1.py:8-43 / b1_read.py:9-30 on the container the HDF5File stand-in writes: open "r", read the mesh into
    an empty Mesh(), read every `c/vector_{i}` into a Function of a P1 space on it, re-emit the series with `File << c`."""
    import pfhubbenchmarks_b200.dolfin as df
    from pfhubbenchmarks_b200 import io
    from pfhubbenchmarks_b200.api import Mesh
    mesh = Mesh(kind, ncell, length)
    pts, cells = io.mesh_of(mesh)
    assert io.identify_mesh(pts, cells) == (kind, ncell, length)
    rng = np.random.default_rng(0)
    series = rng.standard_normal((3, len(pts)))
    with pytest.warns(UserWarning, match="HDF5 is not available"):
        out = df.HDF5File(df.MPI.comm_world, str(tmp_path / "conc.h5"), "w")
    out.write(mesh, "mesh")
    for i, v in enumerate(series):
        out.write(v, "c", 0.1 * i)
    out.close()
    # the reader's side, in the reference's words
    hdf = df.HDF5File(df.MPI.comm_world, str(tmp_path / "conc.h5"), "r")
    back = df.Mesh()
    hdf.read(back, "mesh", True)
    assert (back.kind, back.ncell, back.length) == (kind, ncell, length)
    V = df.FunctionSpace(back, "P", 1)
    cs = []
    for i in range(hdf.attributes("c")["count"]):
        c = df.Function(V)
        hdf.read(c, "c/vector_%d" % i)
        cs.append(c)
    assert len(cs) == 3 and all(np.array_equal(c._host, v) for c, v in zip(cs, series))
    with pytest.raises(RuntimeError):
        hdf.read(df.Function(V), "c/vector_3")
    pvd = df.File(str(tmp_path / "again" / "c.pvd"), "compressed")
    pvd << back
    for c in cs:
        pvd << c
    times, vals = io.read_series(str(tmp_path / "again" / "c.pvd"))
    assert np.array_equal(times, [0.0, 1.0, 2.0]) and np.array_equal(vals, series)
    p2, c2, _ = io.read_vtu(str(tmp_path / "again" / "c000000.vtu"))
    assert np.allclose(p2, pts) and np.array_equal(c2, cells)
