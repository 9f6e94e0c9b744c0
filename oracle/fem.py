"""P1 finite-element operators with the reference's quadrature (test infrastructure).

Restates what FFC/uflacs + DOLFIN assemble for the forms in ``dolfin/pfbase.py:361-460``
under ``parameters["form_compiler"]["quadrature_degree"] = 3`` (``dolfin/bench1.py:14-16``):
FIAT "default" degree-3 rules -- 6-point Strang-Fix on triangles, 5-point
Zienkiewicz-Taylor (negative centroid weight) on tetrahedra.  Mass/stiffness terms are
integrated exactly by these rules; f', f'' and f_chem terms are not, which is
parity-critical.
"""
import numpy as np
import scipy.sparse as sp

_SF = (0.659027622374092, 0.231933368553031, 0.109039009072877)
# points in reference coordinates (xi, eta); barycentric = (1-xi-eta, xi, eta)
TRI6_PTS = np.array([[_SF[0], _SF[1]], [_SF[0], _SF[2]], [_SF[1], _SF[0]],
                     [_SF[1], _SF[2]], [_SF[2], _SF[0]], [_SF[2], _SF[1]]])
TRI6_WTS = np.full(6, 1.0 / 12.0)          # sums to reference area 1/2
TET5_PTS = np.array([[0.25, 0.25, 0.25], [0.5, 1 / 6, 1 / 6], [1 / 6, 0.5, 1 / 6],
                     [1 / 6, 1 / 6, 0.5], [1 / 6, 1 / 6, 1 / 6]])
TET5_WTS = np.array([-0.8, 0.45, 0.45, 0.45, 0.45]) / 6.0   # sums to reference volume 1/6


def _bary(pts):
    return np.concatenate([1.0 - pts.sum(1, keepdims=True), pts], axis=1)   # (nq, dim+1)


class P1Space:
    """Assembled scalar P1 operators + quadrature interpolation on a simplicial mesh."""

    def __init__(self, mesh):
        self.mesh = mesh
        X, C = mesh.coords, mesh.cells
        d = mesh.dim
        nc = C.shape[0]
        P = X[C]                                   # (nc, d+1, d)
        Jm = np.transpose(P[:, 1:, :] - P[:, :1, :], (0, 2, 1))      # columns = edge vectors
        detJ = np.abs(np.linalg.det(Jm))
        Jinv = np.linalg.inv(Jm)
        # reference gradients of P1 basis: rows of G (d+1, d)
        G = np.concatenate([-np.ones((1, d)), np.eye(d)], axis=0)
        grads = np.einsum("ak,ckj->caj", G, Jinv)                    # (nc, d+1, d)
        vol = detJ / (2.0 if d == 2 else 6.0)
        self.vol = vol
        self.detJ = detJ
        # element stiffness and exact consistent mass
        Ke = np.einsum("cad,cbd->cab", grads, grads) * vol[:, None, None]
        Mref = (np.ones((d + 1, d + 1)) + np.eye(d + 1)) / ((d + 1) * (d + 2))
        Me = Mref[None] * vol[:, None, None]
        rows = np.repeat(C, d + 1, axis=1).ravel()
        cols = np.tile(C, (1, d + 1)).ravel()
        nv = mesh.nv
        self.K = sp.csr_matrix((Ke.ravel(), (rows, cols)), shape=(nv, nv))
        self.M = sp.csr_matrix((Me.ravel(), (rows, cols)), shape=(nv, nv))
        self.K.sum_duplicates(); self.M.sum_duplicates()
        # quadrature
        pts, wts = (TRI6_PTS, TRI6_WTS) if d == 2 else (TET5_PTS, TET5_WTS)
        lam = _bary(pts)                           # (nq, d+1)
        nq = lam.shape[0]
        qrow = np.repeat(np.arange(nc * nq), d + 1)
        qcol = np.repeat(C, nq, axis=0).ravel()
        qval = np.tile(lam.ravel(), nc)
        self.Phi = sp.csr_matrix((qval, (qrow, qcol)), shape=(nc * nq, nv))
        self.PhiT = self.Phi.T.tocsr()
        self.qw = (detJ[:, None] * wts[None, :]).ravel()             # |det J| w_q
        self.nq = nq

    def at_quad(self, u):
        """u: (nf, nv) -> (nf, nQ)"""
        return (self.Phi @ u.T).T

    def load(self, gq):
        """int phi_i g  for g given at quadrature points: (nf, nQ) -> (nf, nv)"""
        return (self.PhiT @ (gq * self.qw).T).T

    def weighted_mass(self, gq):
        """int phi_i phi_j g (quadrature), g at quadrature points (nQ,) -> sparse (nv, nv)"""
        return (self.PhiT @ sp.diags(gq * self.qw) @ self.Phi).tocsr()

    def integrate(self, eq):
        return float(np.dot(eq, self.qw))


def eval_points(mesh, u, pts, tol=1e-10):
    """f(x) of the reference's peval/sample (dolfin/pfbase.py:50-94): value of the P1 function with vertex values `u`
    at each point, by searching ALL cells for one whose barycentric coordinates are >= -tol (no use of the mesh
    structure; O(ncells * npts), small cases only).  Points in no cell give +inf, as peval's fill value."""
    X, C = mesh.coords, mesh.cells
    d = mesh.dim
    P = X[C]                                                       # (nc, d+1, d)
    T = np.transpose(P[:, 1:, :] - P[:, :1, :], (0, 2, 1))         # columns = edge vectors
    Tinv = np.linalg.inv(T)
    out = np.full(len(pts), np.inf)
    for k, x in enumerate(np.asarray(pts, dtype=float)):
        lam = np.einsum("cij,cj->ci", Tinv, x[None, :] - P[:, 0, :])          # lambda_1..d
        lam0 = 1.0 - lam.sum(axis=1)
        ok = np.where((lam.min(axis=1) >= -tol) & (lam0 >= -tol))[0]
        if ok.size:
            c = ok[0]
            out[k] = lam0[c] * u[C[c, 0]] + lam[c] @ u[C[c, 1:]]
    return out
