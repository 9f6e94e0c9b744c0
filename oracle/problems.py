"""The PFHub problems as the reference drivers define them (test infrastructure).

Every problem is brought to the common discrete form

    R_a(u) = sum_b (Mc[a,b] M + Kc[a,b] K) u_b  -  Mp[a] M u0_a  +  N[g_a](u)

with M the consistent P1 mass matrix, K the stiffness matrix and
N[g]_i = int phi_i g(u_h) evaluated with the degree-3 quadrature rule.  The free
energies are written symbolically and differentiated with sympy, mirroring the
reference's ``df.variable`` / ``df.diff`` (e.g. ``dolfin/bench1.py:59-65``).
"""
import numpy as np
import sympy as sy


def _lambdify(syms, exprs):
    fs = [sy.lambdify(syms, e, "numpy") for e in exprs]

    def f(u):
        return np.stack([np.broadcast_to(np.asarray(fi(*u), dtype=float), u[0].shape) for fi in fs])
    return f


class Problem:
    """Base: subclasses set nf, fields, syms, g_exprs (list of nf sympy exprs), e_expr, m_expr."""
    dirichlet_fields = ()

    def _compile(self):
        s = self.syms
        self.g = _lambdify(s, self.g_exprs)
        dg = [[sy.diff(ga, sb) for sb in s] for ga in self.g_exprs]
        flat = _lambdify(s, [x for row in dg for x in row])
        self.dg = lambda u: flat(u).reshape(self.nf, self.nf, -1)
        self.e = lambda u: _lambdify(s, [self.e_expr])(u)[0]
        self.m = lambda u: _lambdify(s, [self.m_expr])(u)[0]

    def dirichlet(self, mesh):
        return []


def _ic_c(x, y, c0, eps, kx=0.105):
    return c0 + eps * (np.cos(kx * x) * np.cos(0.11 * y) + (np.cos(0.13 * x) * np.cos(0.087 * y)) ** 2
                       + np.cos(0.025 * x - 0.15 * y) * np.cos(0.07 * x - 0.02 * y))


class CahnHilliard(Problem):
    """bench1 / b13d: ``dolfin/bench1.py:29-69`` with ``pfbase.cahn_hilliard_weak_form`` (``pfbase.py:361-383``)."""
    name = "bench1"
    nf = 2
    fields = ("c", "mu")
    csv_header = "time,total_free_energy,total_solute"

    def __init__(self, c_alpha=0.3, c_beta=0.7, rho_s=5.0, kappa=2.0, M=5.0, c0=0.5, epsilon=0.05):
        self.p = dict(c_alpha=c_alpha, c_beta=c_beta, rho_s=rho_s, kappa=kappa, M=M)
        self.c0, self.epsilon = c0, epsilon
        c, mu = sy.symbols("c mu")
        self.syms = (c, mu)
        f = rho_s * (c - c_alpha) ** 2 * (c_beta - c) ** 2
        self.g_exprs = [sy.Integer(0), -sy.diff(f, c)]
        self.e_expr = f
        self.m_expr = c
        self.grad_e = [(0, kappa / 2.0)]
        self._compile()

    def lin(self, dt):
        p = self.p
        Mc = np.array([[1.0 / dt, 0.0], [0.0, 1.0]])
        Kc = np.array([[0.0, p["M"]], [-p["kappa"], 0.0]])
        Mp = np.array([1.0 / dt, 0.0])
        return Mc, Kc, Mp

    def ic(self, mesh):
        X = mesh.coords
        u = np.zeros((2, mesh.nv))
        u[0] = _ic_c(X[:, 0], X[:, 1], self.c0, self.epsilon)     # pfbase.py:187-190 (z ignored in 3-D)
        return u


class Bench2(Problem):
    """Ostwald ripening: ``dolfin/bench2.py:29-116``; IC ``pfbase.py:262-291`` (incl. the i/ii quirk)."""
    name = "bench2"
    nf = 6
    fields = ("c", "mu", "eta1", "eta2", "eta3", "eta4")
    csv_header = "time,total_free_energy,total_solute"

    def __init__(self):
        ca, cb, rho2, kc, ke, Mob, ww, alpha, L = 0.3, 0.7, 2.0, 3.0, 3.0, 5.0, 1.0, 5.0, 5.0
        self.p = dict(c_alpha=ca, c_beta=cb, rho2=rho2, kappa_c=kc, kappa_eta=ke, M=Mob, w=ww, alpha=alpha, L=L)
        c, mu, e1, e2, e3, e4 = sy.symbols("c mu eta1 eta2 eta3 eta4")
        self.syms = (c, mu, e1, e2, e3, e4)
        es = (e1, e2, e3, e4)
        H = sum(e ** 3 * (6 * e ** 2 - 15 * e + 10) for e in es)
        Wd = sum(e ** 2 * (1 - e) ** 2 for e in es) + alpha * sum(
            es[i] ** 2 * es[j] ** 2 for i in range(4) for j in range(i + 1, 4))
        f = rho2 * (c - ca) ** 2 * (1 - H) + rho2 * (c - cb) ** 2 * H + ww * Wd
        self.g_exprs = [sy.Integer(0), -sy.diff(f, c)] + [L * sy.diff(f, e) for e in es]
        self.e_expr = f
        self.m_expr = c
        self.grad_e = [(0, kc / 2.0)] + [(2 + i, ke / 2.0) for i in range(4)]
        self._compile()

    def lin(self, dt):
        p = self.p
        Mc = np.zeros((6, 6)); Kc = np.zeros((6, 6)); Mp = np.zeros(6)
        Mc[0, 0] = 1.0 / dt; Mp[0] = 1.0 / dt; Kc[0, 1] = p["M"]
        Mc[1, 1] = 1.0; Kc[1, 0] = -p["kappa_c"]
        for i in range(2, 6):
            Mc[i, i] = 1.0 / dt; Mp[i] = 1.0 / dt; Kc[i, i] = p["L"] * p["kappa_eta"]
        return Mc, Kc, Mp

    def ic(self, mesh, c0=0.5, epsilon=0.05, epsilon_eta=0.1, psi=1.5):
        x, y = mesh.coords[:, 0], mesh.coords[:, 1]
        u = np.zeros((6, mesh.nv))
        u[0] = _ic_c(x, y, c0, epsilon)
        for i in range(4):
            ii = i + 1.0
            u[2 + i] = epsilon_eta * (
                np.cos((0.01 * ii) * x - 4.0) * np.cos((0.007 + 0.01 * ii) * y)
                + np.cos((0.11 + 0.01 * ii) * x) * np.cos((0.11 + 0.01 * ii) * y)
                + psi * (np.cos((0.046 + 0.001 * i) * x - (0.0405 + 0.001 * i) * y)
                         * np.cos((0.031 + 0.001 * i) * x - (0.004 + 0.001 * i) * y)) ** 2) ** 2
        return u


class Bench3(Problem):
    """Dendrite (isotropic as coded): ``dolfin/bench3.py:32-98``; IC ``pfbase.py:293-320``."""
    name = "bench3"
    nf = 2
    fields = ("U", "phi")
    csv_header = "time,total_free_energy,solid_fraction"

    def __init__(self, Lx=960.0, Ly=960.0):
        W0, tau0, D, Delta = 1.0, 1.0, 10.0, -0.3
        lam = D * tau0 / (0.6267 * W0 ** 2)
        self.p = dict(W=W0, tau=tau0, D=D, Delta=Delta, lam=lam)
        U, phi = sy.symbols("U phi")
        self.syms = (U, phi)
        dfdp = (phi - lam * U * (1 - phi ** 2)) * (1 - phi ** 2)
        self.g_exprs = [-(0.5 / tau0) * dfdp, -(1.0 / tau0) * dfdp]
        self.e_expr = (-sy.Rational(1, 2) * phi ** 2 + sy.Rational(1, 4) * phi ** 4
                       + lam * U * phi * (1 - sy.Rational(2, 3) * phi ** 2 + sy.Rational(1, 5) * phi ** 4))
        self.m_expr = 0.5 * (phi + 1.0) / (Lx * Ly)
        self.grad_e = [(1, 0.5 * W0 ** 2)]
        self._compile()

    def lin(self, dt):
        p = self.p
        Mc = np.diag([1.0 / dt, 1.0 / dt]); Mp = np.array([1.0 / dt, 1.0 / dt])
        Kc = np.array([[p["D"], 0.5 * p["W"] ** 2 / p["tau"]], [0.0, p["W"] ** 2 / p["tau"]]])
        return Mc, Kc, Mp

    def ic(self, mesh, r0=8.0, w=1.0, vin=1.0, vout=-1.0):
        x, y = mesh.coords[:, 0], mesh.coords[:, 1]
        r = np.sqrt(x ** 2 + y ** 2)
        ramp = vout + 0.5 * (vin - vout) * (1.0 + np.cos(np.pi * (r - r0 + 0.5 * w) / w))
        phi = np.where(r < r0 - 0.5 * w, vin, np.where(r > r0 + 0.5 * w, vout, ramp))
        return np.stack([np.full(mesh.nv, self.p["Delta"]), phi])


class Bench6(Problem):
    """CH + electrostatics: ``dolfin/bench6.py:33-93``; IC ``pfbase.py:322-339``; Dirichlet ``bench6.py:77-90``."""
    name = "bench6"
    nf = 3
    fields = ("c", "mu", "phi")
    csv_header = "time,total_free_energy,total_solute"
    dirichlet_fields = (2,)

    def __init__(self):
        ca, cb, kappa, rho, Mob, k, eps = 0.3, 0.7, 2.0, 5.0, 5.0, 0.09, 90.0
        self.p = dict(c_alpha=ca, c_beta=cb, kappa=kappa, rho=rho, M=Mob, k=k, epsilon=eps)
        c, mu, phi = sy.symbols("c mu phi")
        self.syms = (c, mu, phi)
        f = rho * (c - ca) ** 2 * (cb - c) ** 2
        self.g_exprs = [sy.Integer(0), -sy.diff(f, c), sy.Integer(0)]
        self.e_expr = f + k * c * phi / 2.0
        self.m_expr = c
        self.grad_e = [(0, kappa / 2.0)]
        self._compile()

    def lin(self, dt):
        p = self.p
        Mc = np.zeros((3, 3)); Kc = np.zeros((3, 3)); Mp = np.zeros(3)
        Mc[0, 0] = 1.0 / dt; Mp[0] = 1.0 / dt; Kc[0, 1] = p["M"]
        Mc[1, 1] = 1.0; Kc[1, 0] = -p["kappa"]; Mc[1, 2] = -p["k"]
        Kc[2, 2] = -1.0; Mc[2, 0] = p["k"] / p["epsilon"]
        return Mc, Kc, Mp

    def ic(self, mesh, c0=0.5, c1=0.04):
        x, y = mesh.coords[:, 0], mesh.coords[:, 1]
        u = np.zeros((3, mesh.nv))
        u[0] = _ic_c(x, y, c0, c1, kx=0.2)
        return u

    def dirichlet(self, mesh):
        x, y = mesh.coords[:, 0], mesh.coords[:, 1]
        Lx = mesh.lengths[0]
        left = np.where(np.abs(x) < 1e-12)[0]
        right = np.where(np.abs(x - Lx) < 1e-12)[0]
        idx = np.concatenate([left, right])
        val = np.concatenate([np.zeros(left.size), np.sin(y[right] / 7.0)])
        return [(2, idx, val)]
