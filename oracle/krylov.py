"""Preconditioned GMRES for the Cahn-Hilliard Newton system (test infrastructure).

Used (i) as the practical 3-D oracle solver -- sparse LU does not finish at 50^3
 -- and (ii) as the CPU statement of the algorithm the CUDA
solver implements: right-preconditioned restarted GMRES with a Schur-complement block
preconditioner whose inner solves are Jacobi-Chebyshev sweeps on the SPD operators
M and A = M + alpha K, alpha = sqrt(dt Mob kappa).  The
reference itself uses PETSc GMRES(30)+SOR (``dolfin/bench1.py:98-99``); any
convergent linear solver yields the same Newton iterates to solver tolerance.
"""
import numpy as np
import scipy.sparse as sp


def cheb_coeffs(lmin, lmax, nsweeps):
    """(a_k, b_k) of z_{k+1} = z_k + a_k (z_k - z_{k-1}) + b_k D^-1 (r - A z_k), z_0 = 0, z_{-1} = 0."""
    theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    sigma = theta / delta
    rho = 1.0 / sigma
    out = [(0.0, 1.0 / theta)]
    for _ in range(1, nsweeps):
        rho_new = 1.0 / (2.0 * sigma - rho)
        out.append((rho_new * rho, 2.0 * rho_new / delta))
        rho = rho_new
    return out


def cheb_solve(A, dinv, r, lmin, lmax, nsweeps):
    z_prev = np.zeros_like(r)
    z = np.zeros_like(r)
    for (a, b) in cheb_coeffs(lmin, lmax, nsweeps):
        z_new = z + a * (z - z_prev) + b * dinv * (r - A @ z)
        z_prev, z = z, z_new
    return z


def jacobi_bounds(A, M=None):
    """Cheap bounds of spec(D^-1 A) for A = M + alpha K on P1 simplices:
    lmax <= (d+2)/2 (Wathen bound for the mass part; the stiffness part is <= 2),
    lmin ~ min_i rowsum(A)_i / A_ii (the constant mode), floored by the mass bound 1/2 * M_ii/A_ii."""
    d = A.diagonal()
    rs = np.asarray(A.sum(axis=1)).ravel()
    return float(np.min(rs / d))


def symbol_lmax(A, mesh, ng=24):
    """Largest value of the symbol of D^-1 A on the interior rows of a structured 3-D mesh (the spectrum's upper end up to
    boundary effects of a fraction of a per cent): the row of a vertex in the middle of the mesh, Fourier-transformed on a
    grid of (ng+1) x (2 ng+1)^2 wave vectors.  CPU statement of symbol_bounds3 in csrc/pfh_solver.cpp."""
    nx, ny, nz = mesh.shape
    h = np.array(mesh.lengths) / np.array(mesh.shape)
    i = (nz // 2) * (nx + 1) * (ny + 1) + (ny // 2) * (nx + 1) + nx // 2
    row = A.getrow(i)
    off = np.rint((mesh.coords[row.indices] - mesh.coords[i]) / h)          # integer offsets of the stencil
    th = np.pi * np.arange(0, ng + 1) / ng
    th2 = np.pi * np.arange(-ng, ng + 1) / ng
    tx, ty, tz = np.meshgrid(th, th2, th2, indexing="ij")
    sym = np.zeros_like(tx)
    for v, (dx, dy, dz) in zip(row.data, off):
        sym += v * np.cos(dx * tx + dy * ty + dz * tz)
    return float(sym.max() / A[i, i])


class SchurChebPC:
    """P^-1 for J = [[M/dt, Mob K], [-(M_f'' + kappa K), M]]  (SURVEY.md Appendix B.1)."""

    def __init__(self, disc, u, dt, n_m=4, n_a=8, variant="full", lmax=None, symbol_lmax=False):
        V, p = disc.V, disc.prob.p
        self.V, self.dt, self.Mob, self.kappa = V, dt, p["M"], p.get("kappa", p.get("kappa_c"))
        self.alpha = np.sqrt(dt * self.Mob * self.kappa)
        self.A = (V.M + self.alpha * V.K).tocsr()
        self.dinvM = 1.0 / V.M.diagonal()
        self.dinvA = 1.0 / self.A.diagonal()
        dim = disc.mesh.dim
        self.lmax = lmax if lmax is not None else 0.5 * (dim + 2)
        self.lminM = 0.5
        self.lminA = min(0.5, jacobi_bounds(self.A))
        # interval of the shifted solves: generic upper end (d+2)/2, or the product's default in 3-D (pfh_solver.cpp):
        # 1.015 x the maximum of the operator's symbol, which needs one sweep less for the same iteration count
        self.lmaxA = self.lmax
        if symbol_lmax and dim == 3 and min(disc.mesh.shape) >= 4:
            self.lmaxA = 1.015 * globals()["symbol_lmax"](self.A, disc.mesh)
        dg = disc.prob.dg(V.at_quad(u))
        self.C = (self.kappa * V.K - V.weighted_mass(dg[1, 0])).tocsr()   # kappa K + M_f''  (g_mu = -f')
        self.n_m, self.n_a, self.variant = n_m, n_a, variant
        self.nv = disc.nv
        self.napply = 0

    def Minv(self, r):
        return cheb_solve(self.V.M, self.dinvM, r, self.lminM, self.lmax, self.n_m)

    def Ainv(self, r):
        return cheb_solve(self.A, self.dinvA, r, self.lminA, self.lmaxA, self.n_a)

    def __call__(self, r):
        self.napply += 1
        nv, dt = self.nv, self.dt
        r1, r2 = r[:nv], r[nv:]
        if self.variant == "full":
            t = self.Minv(r2)
            q = dt * r1 - dt * self.Mob * (self.V.K @ t)
            dc = self.Ainv(self.V.M @ self.Ainv(q))
            dmu = t + self.Minv(self.C @ dc)
        elif self.variant == "upper":      # [[S, B],[0, D]]^-1
            t = self.Minv(r2)
            q = dt * r1 - dt * self.Mob * (self.V.K @ t)
            dc = self.Ainv(self.V.M @ self.Ainv(q))
            dmu = t
        elif self.variant == "lower":      # [[S, 0],[C, D]]^-1
            dc = self.Ainv(self.V.M @ self.Ainv(dt * r1))
            dmu = self.Minv(r2 + self.C @ dc)
        return np.concatenate([dc, dmu])


def gmres(Aop, b, pc, rtol=1e-5, atol=0.0, restart=30, maxit=1000):
    """Right-preconditioned GMRES(restart), classical Gram-Schmidt, Givens on the Hessenberg.
    Returns (x, its, relres)."""
    n = b.size
    x = np.zeros(n)
    bnorm = np.linalg.norm(b)
    if bnorm == 0.0:
        return x, 0, 0.0
    tol = max(rtol * bnorm, atol)
    its = 0
    r = b.copy()
    beta = bnorm
    while its < maxit:
        V = np.zeros((restart + 1, n))
        H = np.zeros((restart + 1, restart))
        cs = np.zeros(restart); sn = np.zeros(restart)
        g = np.zeros(restart + 1); g[0] = beta
        V[0] = r / beta
        j_done = 0
        for j in range(restart):
            w = Aop(pc(V[j]))
            h = V[:j + 1] @ w                 # classical Gram-Schmidt: one multi-dot
            w = w - V[:j + 1].T @ h           # one multi-axpy
            hn = np.linalg.norm(w)
            H[:j + 1, j] = h; H[j + 1, j] = hn
            for i in range(j):
                t = cs[i] * H[i, j] + sn[i] * H[i + 1, j]
                H[i + 1, j] = -sn[i] * H[i, j] + cs[i] * H[i + 1, j]
                H[i, j] = t
            den = np.hypot(H[j, j], H[j + 1, j])
            cs[j], sn[j] = H[j, j] / den, H[j + 1, j] / den
            H[j, j] = den; H[j + 1, j] = 0.0
            g[j + 1] = -sn[j] * g[j]; g[j] = cs[j] * g[j]
            its += 1; j_done = j + 1
            if hn > 0:
                V[j + 1] = w / hn
            if abs(g[j + 1]) <= tol or its >= maxit:
                break
        y = np.linalg.solve(np.triu(H[:j_done, :j_done]), g[:j_done])
        x = x + pc(V[:j_done].T @ y)
        r = b - Aop(x)
        beta = np.linalg.norm(r)
        if beta <= tol:
            break
    return x, its, beta / bnorm


def make_linear(rtol=1e-10, restart=30, maxit=2000, stats=None, **pc_kw):
    """linear= callback for Discretization.solve_step (Cahn-Hilliard problems)."""
    def linear(disc, J, rhs, u, dt):
        pc = SchurChebPC(disc, u, dt, **pc_kw)
        Jc = J.tocsr()
        x, its, rr = gmres(lambda v: Jc @ v, rhs, pc, rtol=rtol, restart=restart, maxit=maxit)
        if stats is not None:
            stats.append((its, rr))
        return x
    return linear
