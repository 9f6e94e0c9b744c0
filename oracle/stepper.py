"""Backward-Euler step = Newton on the assembled residual/Jacobian (test infrastructure).

Restates ``niters, converged = solver.solve()`` (``dolfin/bench1.py:162``) plus the two
``df.assemble`` reductions that follow it (``dolfin/bench1.py:121-125,193-194``).  Oracle
policy: iterate Newton to the exactly converged discrete solution
(||F|| ~ 1e-11), using sparse LU in 2-D and a tight preconditioned Krylov solve in 3-D.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .fem import P1Space


class Discretization:
    def __init__(self, mesh, problem):
        self.mesh, self.prob = mesh, problem
        self.V = P1Space(mesh)
        self.nf, self.nv = problem.nf, mesh.nv
        self.bcs = problem.dirichlet(mesh)

    # ---- residual / Jacobian (fields as (nf, nv) arrays) ----
    def residual(self, u, u0, dt):
        V, P = self.V, self.prob
        Mc, Kc, Mp = P.lin(dt)
        Mu = (V.M @ u.T).T
        Ku = (V.K @ u.T).T
        R = Mc @ Mu + Kc @ Ku - Mp[:, None] * (V.M @ u0.T).T
        R += V.load(P.g(V.at_quad(u)))
        for (a, idx, val) in self.bcs:
            R[a, idx] = u[a, idx] - val
        return R

    def jacobian(self, u, dt):
        V, P = self.V, self.prob
        Mc, Kc, _ = P.lin(dt)
        dg = P.dg(V.at_quad(u))
        blocks = [[None] * self.nf for _ in range(self.nf)]
        for a in range(self.nf):
            for b in range(self.nf):
                B = Mc[a, b] * V.M + Kc[a, b] * V.K
                if np.any(dg[a, b] != 0.0):
                    B = B + V.weighted_mass(dg[a, b])
                blocks[a][b] = B.tocsr()
        for (a, idx, _) in self.bcs:
            keep = np.ones(self.nv); keep[idx] = 0.0
            D = sp.diags(keep)
            for b in range(self.nf):
                blocks[a][b] = D @ blocks[a][b]
            blocks[a][a] = blocks[a][a] + sp.diags(1.0 - keep)
        return sp.bmat(blocks, format="csc")

    def energy(self, u):
        V, P = self.V, self.prob
        E = V.integrate(P.e(V.at_quad(u)))
        for (a, coef) in P.grad_e:
            E += coef * float(u[a] @ (V.K @ u[a]))
        return E

    def mass(self, u):
        return self.V.integrate(self.prob.m(self.V.at_quad(u)))

    # ---- Newton ----
    def solve_step(self, u, u0, dt, atol=1e-11, rtol=0.0, max_it=50, linear="lu", report=None):
        """In-place semantics of the reference: u is the initial guess and the result.
        Returns (u, niters, converged)."""
        u = u.copy()
        for (a, idx, val) in self.bcs:
            u[a, idx] = val
        f0 = None
        for it in range(max_it + 1):
            R = self.residual(u, u0, dt)
            fn = float(np.linalg.norm(R))
            if f0 is None:
                f0 = fn
            if report is not None:
                report.append(fn)
            if fn < atol or fn < rtol * f0:
                return u, it, True
            if it == max_it:
                break
            J = self.jacobian(u, dt)
            if linear == "lu":
                du = spla.splu(J).solve(-R.ravel())
            else:
                du = linear(self, J, -R.ravel(), u, dt)
            u = u + du.reshape(self.nf, self.nv)
        return u, max_it, False
