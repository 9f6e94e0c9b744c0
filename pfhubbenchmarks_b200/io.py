"""Field output in the formats the reference drivers write ("next" row N2 of SURVEY.md section 8f).

* `File(path + ".pvd") << (c, t)`  (dolfin/bench2.py:178-182,269-274; b13d.py:104-106,179): a PVD collection
  plus one VTU per step, `format="ascii"` with 17 significant digits (as committed under results/bench2,
  results/bench6) or `format="binary"` with vtkZLibDataCompressor (as committed under results/bench1).
* Points / connectivity follow DOLFIN's structured meshes: crossed rectangle (corners x-fastest, then cell
  centres; 4 triangles per square) and BoxMesh (6 Kuhn tetrahedra per cube).
Host-side only; fields arrive as numpy arrays (one value per mesh vertex, DOLFIN vertex order).
"""
import base64
import os
import struct
import zlib

import numpy as np


def crossed_mesh(nx, ny, Lx, Ly):
    """(points (n,3), cells (m,3)) of df.RectangleMesh(Point(0,0), Point(Lx,Ly), nx, ny, 'crossed')."""
    hx, hy = Lx / nx, Ly / ny
    jx, jy = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="xy")
    corners = np.stack([jx.ravel() * hx, jy.ravel() * hy, np.zeros(jx.size)], axis=1)
    cx, cy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    centres = np.stack([(cx.ravel() + 0.5) * hx, (cy.ravel() + 0.5) * hy, np.zeros(cx.size)], axis=1)
    v0 = (cy * (nx + 1) + cx).ravel()
    vm = (nx + 1) * (ny + 1) + (cy * nx + cx).ravel()
    tri = np.stack([v0, v0 + 1, vm, v0, v0 + nx + 1, vm, v0 + 1, v0 + nx + 2, vm, v0 + nx + 1, v0 + nx + 2, vm], axis=1)
    return np.concatenate([corners, centres]), tri.reshape(-1, 3)


_KUHN = ((0, 1, 3, 7), (0, 1, 7, 5), (0, 5, 7, 4), (0, 3, 2, 7), (0, 6, 4, 7), (0, 2, 6, 7))


def box_mesh(nx, ny, nz, Lx, Ly, Lz):
    """(points, cells (m,4)) of df.BoxMesh(Point(0,0,0), Point(Lx,Ly,Lz), nx, ny, nz)."""
    kz, ky, kx = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    pts = np.stack([kx.ravel() * (Lx / nx), ky.ravel() * (Ly / ny), kz.ravel() * (Lz / nz)], axis=1)
    cz, cy, cx = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    sy, sz = nx + 1, (nx + 1) * (ny + 1)
    v0 = (cz * sz + cy * sy + cx).ravel()
    loc = [v0 + (l & 1) + ((l >> 1) & 1) * sy + (l >> 2) * sz for l in range(8)]
    cells = np.stack([np.stack([loc[a] for a in t], axis=1) for t in _KUHN], axis=1).reshape(-1, 4)
    return pts, cells


def right_mesh(nx, ny, Lx, Ly):
    """(points, cells (m,3)) of df.UnitSquareMesh-style right-diagonal meshes (dolfin/nl.py:20)."""
    hx, hy = Lx / nx, Ly / ny
    jx, jy = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="xy")
    pts = np.stack([jx.ravel() * hx, jy.ravel() * hy, np.zeros(jx.size)], axis=1)
    cx, cy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    v0 = (cy * (nx + 1) + cx).ravel()
    tri = np.stack([v0, v0 + 1, v0 + nx + 2, v0, v0 + nx + 1, v0 + nx + 2], axis=1)
    return pts, tri.reshape(-1, 3)


def mesh_of(mesh):
    """Points / cells of a pfhubbenchmarks_b200.api.Mesh."""
    if mesh.kind == "crossed2d":
        return crossed_mesh(mesh.ncell[0], mesh.ncell[1], mesh.length[0], mesh.length[1])
    if mesh.kind == "right2d":
        return right_mesh(mesh.ncell[0], mesh.ncell[1], mesh.length[0], mesh.length[1])
    return box_mesh(*mesh.ncell, *mesh.length)


def identify_mesh(points, cells):
    """(kind, ncell, length) of a stored structured mesh -- the inverse of mesh_of, from the vertex coordinates alone
    (corners are stored x-fastest): what `hdf.read(mesh, "mesh", True)` recovers (process_bench1.py:14-15).  Every candidate
    size is checked by regenerating the mesh and comparing points and connectivity."""
    from .api import Mesh
    pts = np.asarray(points, dtype=float)
    cells = np.asarray(cells)
    n = len(pts)
    nxv = int(np.argmax(pts[:, 1] > 0)) if np.any(pts[:, 1] > 0) else n           # first vertex of the second row
    cands = []
    if cells.shape[1] == 4 and np.any(pts[:, 2] > 0):
        nxy = int(np.argmax(pts[:, 2] > 0))
        if nxv >= 2 and nxy % nxv == 0 and n % nxy == 0:
            cands.append(("kuhn3d", (nxv - 1, nxy // nxv - 1, n // nxy - 1), tuple(float(v) for v in pts.max(axis=0))))
    elif cells.shape[1] == 3 and nxv >= 2:
        length = tuple(float(v) for v in pts[:, :2].max(axis=0))
        for nyv in range(2, n // nxv + 1):                                          # corners only, or corners + centres
            if nxv * nyv == n:
                cands.append(("right2d", (nxv - 1, nyv - 1), length))
            if nxv * nyv + (nxv - 1) * (nyv - 1) == n:
                cands.append(("crossed2d", (nxv - 1, nyv - 1), length))
    for kind, ncell, length in cands:
        gp, gc = mesh_of(Mesh(kind, ncell, length))
        if gp.shape == pts.shape and gc.shape == cells.shape and np.allclose(gp, pts, atol=1e-9 * max(length)) and np.array_equal(gc, cells):
            return kind, ncell, length
    raise ValueError("the stored mesh is not one of the structured DOLFIN meshes (crossed / right-diagonal rectangle, box)")


def _b64_zlib(arr, block=32768):
    raw = np.ascontiguousarray(arr).tobytes()
    blocks = [raw[i:i + block] for i in range(0, len(raw), block)] or [b""]
    comp = [zlib.compress(b) for b in blocks]
    header = struct.pack("<%dI" % (3 + len(comp)), len(blocks), block, len(blocks[-1]) if len(blocks[-1]) != block else block,
                         *[len(cb) for cb in comp])
    if len(blocks[-1]) == block:   # VTK convention: last-block size 0 means "full block"
        header = struct.pack("<%dI" % (3 + len(comp)), len(blocks), block, 0, *[len(cb) for cb in comp])
    return base64.b64encode(header).decode() + base64.b64encode(b"".join(comp)).decode()


def _fmt(arr, binary):
    if binary:
        return _b64_zlib(arr)
    if arr.dtype.kind == "f":
        return "  ".join(" ".join("%.17g" % v for v in row) for row in np.atleast_2d(arr)) + "  "
    return "  ".join(" ".join(str(int(v)) for v in row) for row in np.atleast_2d(arr)) + "  "


def write_vtu(path, points, cells, values, name="f", binary=False):
    """One scalar per vertex on a simplicial mesh, DOLFIN VTU layout (Points, connectivity, offsets, types, PointData)."""
    points = np.asarray(points, dtype=np.float64)
    cells = np.asarray(cells)
    values = np.asarray(values, dtype=np.float64).ravel()
    assert values.size == points.shape[0]
    nper = cells.shape[1]
    vtk_type = 5 if nper == 3 else 10
    fmt = 'format="binary"' if binary else 'format="ascii"'
    comp = ' byte_order="LittleEndian" compressor="vtkZLibDataCompressor"' if binary else ""
    offsets = (np.arange(cells.shape[0], dtype=np.uint32) + 1) * nper
    types = np.full(cells.shape[0], vtk_type, dtype=np.uint8)
    with open(path, "w") as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid"  version="0.1" %s>\n<UnstructuredGrid>\n' % comp)
        f.write('<Piece  NumberOfPoints="%d" NumberOfCells="%d">\n<Points>\n' % (points.shape[0], cells.shape[0]))
        f.write('<DataArray  type="Float64"  NumberOfComponents="3"  %s>%s</DataArray>\n</Points>\n<Cells>\n' % (fmt, _fmt(points, binary)))
        f.write('<DataArray  type="UInt32"  Name="connectivity"  %s>%s</DataArray>\n' % (fmt, _fmt(cells.astype(np.uint32), binary)))
        f.write('<DataArray  type="UInt32"  Name="offsets"  %s>%s</DataArray>\n' % (fmt, _fmt(offsets[None, :], binary)))
        f.write('<DataArray  type="UInt8"  Name="types"  %s>%s</DataArray>\n</Cells>\n' % (fmt, _fmt(types[None, :], binary)))
        f.write('<PointData  Scalars="%s"> \n<DataArray  type="Float64"  Name="%s"  %s>%s</DataArray>\n</PointData> \n' % (
            name, name, fmt, _fmt(values[None, :] if binary else values[:, None], binary)))
        f.write("</Piece>\n</UnstructuredGrid>\n</VTKFile>")


class File:
    """`cfile = File(filename + ".pvd", "compressed")` ; `cfile << (c, t)`  (dolfin/b13d.py:106,179)."""

    def __init__(self, path, mode="ascii", mesh=None, name="f"):
        assert path.endswith(".pvd")
        self.path, self.binary, self.mesh, self.name = path, mode == "compressed", mesh, name
        self.entries = []
        self._geom = None

    def write(self, values, t):
        if self._geom is None:
            self._geom = mesh_of(self.mesh)
        stem = os.path.basename(self.path)[:-4]
        os.makedirs(os.path.dirname(self.path) or ".", exist_ok=True)      # DOLFIN creates results/<name>/ on first write
        fn = "%s%06d.vtu" % (stem, len(self.entries))
        write_vtu(os.path.join(os.path.dirname(self.path) or ".", fn), self._geom[0], self._geom[1], values, self.name, self.binary)
        self.entries.append((float(t), fn))
        with open(self.path, "w") as f:
            f.write('<?xml version="1.0"?>\n<VTKFile type="Collection" version="0.1">\n  <Collection>\n')
            for (tt, name) in self.entries:
                f.write('    <DataSet timestep="%s" part="0" file="%s" />\n' % (repr(tt), name))
            f.write("  </Collection>\n</VTKFile>\n")

    def __lshift__(self, pair):
        """`file << (c, t)` with c a component of a Function (`c, _ = w.split()`, bench2.py:269-274), a device tensor or
        a numpy array; the mesh is taken from the Function when the File was opened without one, as in the reference."""
        if hasattr(pair, "kind") and hasattr(pair, "ncell"):       # `file << mesh` (process_bench1.py:39): every VTU carries it
            self.mesh = pair
            return self
        values, t = pair if isinstance(pair, tuple) else (pair, len(self.entries))   # `file << c` numbers the steps itself
        if hasattr(values, "_host") and values._host is not None:  # a Function read back from a stored series
            if self.mesh is None:
                self.mesh = values.W.mesh
            values = values._host
        if hasattr(values, "owner"):                      # dolfin.SubFunction
            if self.mesh is None:
                self.mesh = values.owner.W.mesh
            values = values.tensor
        if hasattr(values, "detach"):                     # torch tensor
            values = values.detach().cpu().numpy()
        self.write(values, float(t))
        return self


# ---------------------------------------------------------------------------------------------------------------------
# readers: what process_bench1.py / b1_read.py do with the HDF5 time series (`c/vector_{i}`), on the PVD/VTU series
# ---------------------------------------------------------------------------------------------------------------------
_NP_TYPES = {"Float64": "<f8", "Float32": "<f4", "UInt32": "<u4", "Int32": "<i4", "UInt8": "u1", "Int64": "<i8", "UInt64": "<u8"}


def _inline_zlib(text, dtype):
    """VTK inline compressed DataArray: base64 of the uint32 header (nblocks, block size, last block size, compressed
    sizes...) directly followed by base64 of the concatenated zlib blocks."""
    text = "".join(text.split())
    nblocks = struct.unpack("<I", base64.b64decode(text[:8])[:4])[0]
    hlen = 4 * (3 + nblocks)
    hchars = 4 * ((hlen + 2) // 3)
    sizes = struct.unpack("<%dI" % (3 + nblocks), base64.b64decode(text[:hchars])[:hlen])[3:]
    blob, out, pos = base64.b64decode(text[hchars:]), [], 0
    for n in sizes:
        out.append(zlib.decompress(blob[pos:pos + n]))
        pos += n
    return np.frombuffer(b"".join(out), dtype=dtype)


def read_vtu(path):
    """(points (n, 3), cells (m, k), {name: values (n,)}) of a VTU file as DOLFIN (or write_vtu) writes it:
    `format="ascii"` or `format="binary"` with vtkZLibDataCompressor."""
    import xml.etree.ElementTree as ET
    root = ET.parse(path).getroot()
    piece = root.find("./UnstructuredGrid/Piece")
    npts, ncells = int(piece.get("NumberOfPoints")), int(piece.get("NumberOfCells"))

    def array(node):
        dtype = np.dtype(_NP_TYPES[node.get("type")])
        if node.get("format") == "binary":
            if root.get("compressor") != "vtkZLibDataCompressor":
                raise ValueError("only zlib-compressed inline binary arrays are supported")
            return _inline_zlib(node.text, dtype)
        return np.array(node.text.split(), dtype=np.float64).astype(dtype)

    points = array(piece.find("./Points/DataArray")).reshape(npts, 3).astype(np.float64)
    conn = next(a for a in piece.findall("./Cells/DataArray") if a.get("Name") == "connectivity")
    cells = array(conn).reshape(ncells, -1).astype(np.int64)
    data = {a.get("Name"): array(a).astype(np.float64) for a in piece.findall("./PointData/DataArray")}
    return points, cells, data


def read_pvd(path):
    """[(time, vtu path), ...] of a PVD collection (paths resolved relative to the PVD file)."""
    import xml.etree.ElementTree as ET
    base = os.path.dirname(os.path.abspath(path))
    return [(float(d.get("timestep")), os.path.join(base, d.get("file"))) for d in ET.parse(path).getroot().iter("DataSet")]


def read_series(path):
    """times (nt,), values (nt, n_vertices) of the single scalar field of a PVD series: the arrays the reference's
    post-processing builds from `c/vector_{i}` (process_bench1.py:8-34)."""
    entries = read_pvd(path)
    vals = []
    for _, fn in entries:
        data = read_vtu(fn)[2]
        if len(data) != 1:
            raise ValueError("expected one scalar field per VTU, found %s" % sorted(data))
        vals.append(next(iter(data.values())))
    return np.array([t for t, _ in entries]), np.array(vals)
