"""Field output (VTU/PVD, "next" row N2): round trip through the golden readers and mesh identity with DOLFIN's files."""
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
