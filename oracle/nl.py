"""Steady quasi-linear diffusion demo of the reference, restated (TEST INFRASTRUCTURE, NOT PRODUCT).

``dolfin/nl.py`` builds, on ``UnitSquareMesh(60, 75)`` (right-diagonal split; the second assignment at ``nl.py:19-20``
wins) with P1 elements,

    F(u; v) = int 5 v dx - int sqrt(u u) grad u . grad v dx - int u v dx                  (nl.py:37)
    J       = derivative(F, u, du)                                                        (nl.py:38)
    u = 0.9 on x = 0 and x = 1 (nl.py:31-35), start u = 0.9 (nl.py:48,75),

and solves it twice with the same settings (SNES newtonls 'basic', then DOLFIN's Newton solver; atol 1e-6, at most 10
iterations, gmres + sor; nl.py:50-96), printing ``(niters, converged)``.  Under ``quadrature_degree = 3`` (nl.py:12)
every term is integrated with the 6-point Strang-Fix rule: on a triangle T with constant grad u

    int_T |u| grad u . grad phi_i = (grad u . grad phi_i) sum_q w_q |u_q|
    d/du_j of it                  = (grad u . grad phi_i) sum_q w_q sign(u_q) phi_j(q) + (grad phi_j . grad phi_i) sum_q w_q |u_q|

(UFL differentiates sqrt(u u) to u / sqrt(u u) = sign u).  Dirichlet rows as DOLFIN applies them to a Newton system:
F_i = u_i - g_i, J_ii = 1.

**Parity unpinned**: the reference records no output of this script (no expected values, no asserts: SURVEY.md section 4),
and FEniCS is not installable here.  What the tests check instead is the discrete problem itself (``tests/test_nl.py``:
J against finite differences of F, second-order convergence to a manufactured solution, mesh order) and product == oracle.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .fem import P1Space
from .mesh import right2d


class NlProblem:
    """F_i(u) = src int phi_i - kc int |u| grad u . grad phi_i - reac int u phi_i, u = bc on x = 0 and x = Lx."""

    def __init__(self, nx=60, ny=75, Lx=1.0, Ly=1.0, src=5.0, kc=1.0, reac=1.0, bc=0.9, load=None):
        self.mesh = right2d(nx, ny, Lx, Ly)
        self.V = P1Space(self.mesh)
        self.src, self.kc, self.reac, self.bc = src, kc, reac, bc
        X, C = self.mesh.coords, self.mesh.cells
        self.nv = self.mesh.nv
        ix = np.arange(self.nv) % (nx + 1)
        self.bc_idx = np.where((ix == 0) | (ix == nx))[0]             # np.isclose(x[0], 0) | np.isclose(x[0], 1)  (nl.py:34)
        self.bc_val = np.full(self.bc_idx.size, bc)                    # may be overwritten (manufactured solutions)
        # constant P1 gradients per cell
        P = X[C]
        Jm = np.transpose(P[:, 1:, :] - P[:, :1, :], (0, 2, 1))
        Jinv = np.linalg.inv(Jm)
        G = np.concatenate([-np.ones((1, 2)), np.eye(2)], axis=0)
        self.grads = np.einsum("ak,ckj->caj", G, Jinv)                # (nc, 3, 2)
        self.load = load                                               # optional extra source at quadrature points (nQ,)
        self.one = np.ones(self.nv)

    # ---- pieces at quadrature points ----
    def _cell_q(self, uq):
        return uq.reshape(self.mesh.cells.shape[0], self.V.nq)

    def _kbar(self, u):
        """sum_q w_q kc |u_q| per cell."""
        V = self.V
        uq = V.Phi @ u
        return (self.kc * np.abs(uq) * V.qw).reshape(-1, V.nq).sum(1)

    def residual(self, u):
        V, C = self.V, self.mesh.cells
        uq = V.Phi @ u
        sq = self.src - self.reac * uq
        if self.load is not None:
            sq = sq + self.load
        R = V.PhiT @ (sq * V.qw)
        gu = np.einsum("ca,caj->cj", u[C], self.grads)                # grad u per cell
        flux = np.einsum("cj,caj->ca", gu, self.grads) * self._kbar(u)[:, None]
        np.subtract.at(R, C.ravel(), flux.ravel())
        R[self.bc_idx] = u[self.bc_idx] - self.bc_val
        return R

    def jacobian(self, u):
        V, C = self.V, self.mesh.cells
        nc = C.shape[0]
        uq = V.Phi @ u
        J = -self.reac * V.M
        kbar = self._kbar(u)
        Ke = -np.einsum("cad,cbd->cab", self.grads, self.grads) * kbar[:, None, None]
        # convective part: -(grad u . grad phi_i) sum_q w_q kc sign(u_q) phi_j(q)
        lam = np.concatenate([1.0 - np.array(_TRI_PTS).sum(1, keepdims=True), np.array(_TRI_PTS)], axis=1)   # (nq, 3)
        sg = (self.kc * np.sign(uq) * V.qw).reshape(nc, V.nq)
        sj = sg @ lam                                                  # (nc, 3): sum_q w_q kc sign(u_q) phi_j(q)
        gu = np.einsum("ca,caj->cj", u[C], self.grads)
        gi = np.einsum("cj,caj->ca", gu, self.grads)                  # grad u . grad phi_i
        Ce = -gi[:, :, None] * sj[:, None, :]
        rows = np.repeat(C, 3, axis=1).ravel()
        cols = np.tile(C, (1, 3)).ravel()
        J = J + sp.csr_matrix(((Ke + Ce).ravel(), (rows, cols)), shape=(self.nv, self.nv))
        keep = np.ones(self.nv)
        keep[self.bc_idx] = 0.0
        J = sp.diags(keep) @ J + sp.diags(1.0 - keep)
        return J.tocsc()

    def sym_operator(self, u):
        """A = K_{kc |u|} + reac M with identity Dirichlet rows: the SPD operator the product's Chebyshev preconditioner
        sweeps on (csrc/pfh_nl.cuh, MODE 2); -A is J without the convective term."""
        V, C = self.V, self.mesh.cells
        Ke = np.einsum("cad,cbd->cab", self.grads, self.grads) * self._kbar(u)[:, None, None]
        rows = np.repeat(C, 3, axis=1).ravel()
        cols = np.tile(C, (1, 3)).ravel()
        A = self.reac * V.M + sp.csr_matrix((Ke.ravel(), (rows, cols)), shape=(self.nv, self.nv))
        keep = np.ones(self.nv)
        keep[self.bc_idx] = 0.0
        return (sp.diags(keep) @ A + sp.diags(1.0 - keep)).tocsr()

    # ---- niters, converged = solver.solve()  (nl.py:66, 93) ----
    def solve(self, u=None, atol=1e-6, rtol=1e-9, max_it=10, report=None):
        u = np.full(self.nv, self.bc) if u is None else np.array(u, dtype=float)
        f0 = None
        for it in range(max_it + 1):
            R = self.residual(u)
            fn = float(np.linalg.norm(R))
            if f0 is None:
                f0 = fn
            if report is not None:
                report.append(fn)
            if fn < atol or (it > 0 and fn < rtol * f0):
                return u, it, True
            if it == max_it:
                break
            u = u + spla.splu(self.jacobian(u)).solve(-R)
        return u, max_it, False


from .fem import TRI6_PTS as _TRI_PTS  # noqa: E402  (reference coordinates of the 6 Strang-Fix points)
