"""DOLFIN structured simplicial meshes, restated (test infrastructure).

* ``crossed2d``  = ``df.RectangleMesh(P0, P1, nx, ny, 'crossed')`` (``dolfin/bench1.py:21-23``):
  vertices = grid corners (x fastest) followed by cell centres; four triangles per
  square ``(v0,v1,vm) (v0,v2,vm) (v1,v3,vm) (v2,v3,vm)``.  Verified against the
  connectivity stored in ``results/bench1/conc000000.vtu`` (tests/test_oracle_golden.py).
* ``kuhn3d``     = ``df.BoxMesh(P0, P1, nx, ny, nz)`` (``dolfin/b13d.py:24-26``): vertices x
  fastest, then y, then z; six tetrahedra per cube, all sharing the diagonal v0-v7
  (DOLFIN ``BoxMesh::build_tet``; third-party, restated from its published source).
* ``right2d``    = ``df.UnitSquareMesh``-style right-diagonal split (``dolfin/nl.py:20``).
"""
import numpy as np


class Mesh:
    def __init__(self, kind, coords, cells, shape, lengths):
        self.kind = kind
        self.coords = coords          # (nv, dim)
        self.cells = cells            # (nc, dim+1)
        self.shape = shape            # number of cells per direction
        self.lengths = lengths
        self.dim = coords.shape[1]
        self.nv = coords.shape[0]


def crossed2d(nx, ny, Lx, Ly):
    hx, hy = Lx / nx, Ly / ny
    ix, iy = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="xy")
    corners = np.stack([ix.ravel() * hx, iy.ravel() * hy], axis=1)
    cx, cy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    centres = np.stack([(cx.ravel() + 0.5) * hx, (cy.ravel() + 0.5) * hy], axis=1)
    coords = np.concatenate([corners, centres])
    ncorner = (nx + 1) * (ny + 1)
    v0 = (cy * (nx + 1) + cx).ravel()
    v1, v2 = v0 + 1, v0 + nx + 1
    v3 = v2 + 1
    vm = ncorner + (cy * nx + cx).ravel()
    cells = np.stack([np.stack([v0, v1, vm], 1), np.stack([v0, v2, vm], 1),
                      np.stack([v1, v3, vm], 1), np.stack([v2, v3, vm], 1)], axis=1).reshape(-1, 3)
    return Mesh("crossed2d", coords, cells, (nx, ny), (Lx, Ly))


def right2d(nx, ny, Lx=1.0, Ly=1.0):
    hx, hy = Lx / nx, Ly / ny
    ix, iy = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="xy")
    coords = np.stack([ix.ravel() * hx, iy.ravel() * hy], axis=1)
    cx, cy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    v0 = (cy * (nx + 1) + cx).ravel()
    v1, v2 = v0 + 1, v0 + nx + 1
    v3 = v2 + 1
    cells = np.stack([np.stack([v0, v1, v3], 1), np.stack([v0, v2, v3], 1)], axis=1).reshape(-1, 3)
    return Mesh("right2d", coords, cells, (nx, ny), (Lx, Ly))


KUHN_TETS = ((0, 1, 3, 7), (0, 1, 7, 5), (0, 5, 7, 4), (0, 3, 2, 7), (0, 6, 4, 7), (0, 2, 6, 7))


def kuhn3d(nx, ny, nz, Lx, Ly, Lz):
    hx, hy, hz = Lx / nx, Ly / ny, Lz / nz
    iz, iy, ix = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    coords = np.stack([ix.ravel() * hx, iy.ravel() * hy, iz.ravel() * hz], axis=1)
    cz, cy, cx = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    sx, sy = 1, nx + 1
    sz = (nx + 1) * (ny + 1)
    v0 = (cz * sz + cy * sy + cx).ravel()
    v = [v0, v0 + sx, v0 + sy, v0 + sx + sy, v0 + sz, v0 + sx + sz, v0 + sy + sz, v0 + sx + sy + sz]
    cells = np.stack([np.stack([v[a], v[b], v[c], v[d]], 1) for (a, b, c, d) in KUHN_TETS],
                     axis=1).reshape(-1, 4)
    return Mesh("kuhn3d", coords, cells, (nx, ny, nz), (Lx, Ly, Lz))
