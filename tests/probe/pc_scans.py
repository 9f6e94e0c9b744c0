"""Solver-side scans quoted in DESIGN.md section 5 / profiles/README.md (round 2), on the oracle (CPU, minutes each):

    python tests/probe/pc_scans.py interval      # Chebyshev interval of the shifted solves x sweep count
    python tests/probe/pc_scans.py lump          # lumped mass in the middle of the Schur approximation
    python tests/probe/pc_scans.py uniform       # one uniform Jacobi scale instead of the true diagonal
    python tests/probe/pc_scans.py schur         # exact inner solves: alpha-form vs mean-f'' form vs exact S

b13d on a 20 x 18 x 16-cell grid, dt list 0.1, 0.2, 0.4, 0.8, 1.6 x 4, reference tolerances; counted: Newton and Krylov
iterations of the last five steps (what bench.py times).  The CPU leg reproduces the 512^3 counts of the GPU to 1-3 %.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import krylov, mesh, problems, stepper  # noqa: E402

N = (20, 18, 16)
M = mesh.kuhn3d(*N, 2.0 * N[0], 2.0 * N[1], 2.0 * N[2])
P = problems.CahnHilliard()
D = stepper.Discretization(M, P)


def trajectory(make_pc):
    stats = []

    def linear(disc, J, rhs, u, dt):
        pc = make_pc(disc, u, dt)
        Jc = J.tocsr()
        x, its, rr = krylov.gmres(lambda v: Jc @ v, rhs, pc, rtol=1e-5)
        stats.append(its)
        return x
    u = P.ic(M)
    newton = 0
    for k, dt in enumerate([0.1, 0.2, 0.4, 0.8, 1.6, 1.6, 1.6, 1.6]):
        u, nit, ok = D.solve_step(u, u, dt, atol=1e-6, rtol=1e-9, max_it=10, linear=linear)
        if k >= 3:
            newton += nit
        if k == 2:
            stats.clear()
    return newton, sum(stats)


def symbol_extremes(A):
    hi = krylov.symbol_lmax(A, M)
    d = A.diagonal()
    return hi, float(np.min(np.asarray(A.sum(axis=1)).ravel() / d))


def scan_interval():
    for n_a in (5, 4, 3):
        for flo in (0.5, 0.76, 1.0, 1.5):
            for fhi in (1.015, 1.15, 1.31):
                def make(disc, u, dt, n_a=n_a, flo=flo, fhi=fhi):
                    pc = krylov.SchurChebPC(disc, u, dt, n_m=3, n_a=n_a, variant="upper")
                    hi, _ = symbol_extremes(pc.A)
                    # lower end: flo x (symbol minimum ~ Rayleigh estimate / 0.76); upper end: fhi x symbol maximum
                    pc.lminA, pc.lmaxA = flo * pc.lminA / 0.76, fhi * hi
                    return pc
                print("sweeps 3/%d  lmin x%.2f  lmax x%.3f  (newton, krylov) =" % (n_a, flo, fhi), trajectory(make), flush=True)


def scan_lump():
    class PCL(krylov.SchurChebPC):
        def __call__(self, r):
            nv, dt = self.nv, self.dt
            ml = np.asarray(self.V.M.sum(axis=1)).ravel()
            t = self.Minv(r[nv:])
            q = dt * r[:nv] - dt * self.Mob * (self.V.K @ t)
            return np.concatenate([self.Ainv(ml * self.Ainv(q)), t])
    print("consistent mass:", trajectory(lambda d, u, dt: krylov.SchurChebPC(d, u, dt, n_m=3, n_a=4, variant="upper", symbol_lmax=True)))
    print("lumped mass in the middle:", trajectory(lambda d, u, dt: PCL(d, u, dt, n_m=3, n_a=4, variant="upper", symbol_lmax=True)))


def scan_uniform():
    def make(disc, u, dt):
        pc = krylov.SchurChebPC(disc, u, dt, n_m=3, n_a=4, variant="upper", symbol_lmax=True)
        pc.dinvM = np.full_like(pc.dinvM, pc.dinvM.min())
        pc.dinvA = np.full_like(pc.dinvA, pc.dinvA.min())
        return pc
    print("uniform Jacobi scale:", trajectory(make))


def scan_schur():
    V = D.V

    class Exact:
        def __init__(self, disc, u, dt, mode):
            p = disc.prob.p
            self.dt, self.Mob, self.kappa, self.nv = dt, p["M"], p["kappa"], disc.nv
            Mm, K = V.M.tocsc(), V.K.tocsc()
            self.K = K
            alpha = np.sqrt(dt * self.Mob * self.kappa)
            Mf = -V.weighted_mass(disc.prob.dg(V.at_quad(u))[1, 0])     # M_f''
            self.Mlu = spl.splu(Mm)
            nv = self.nv
            if mode == "alpha":
                lu = spl.splu((Mm + alpha * K).tocsc())
                self.S = lambda q: lu.solve(Mm @ lu.solve(q))
            else:
                C = self.kappa * K + (Mf if mode == "exact" else float(Mf.sum() / Mm.sum()) * Mm)
                lu = spl.splu(sp.bmat([[Mm, dt * self.Mob * K], [-C, Mm]]).tocsc())
                self.S = lambda q: lu.solve(np.concatenate([q, np.zeros(nv)]))[:nv]

        def __call__(self, r):
            nv, dt = self.nv, self.dt
            t = self.Mlu.solve(r[nv:])
            return np.concatenate([self.S(dt * r[:nv] - dt * self.Mob * (self.K @ t)), t])
    for mode in ("alpha", "mean", "exact"):
        print("exact inner solves, %s:" % mode, trajectory(lambda d, u, dt, mode=mode: Exact(d, u, dt, mode)), flush=True)


if __name__ == "__main__":
    {"interval": scan_interval, "lump": scan_lump, "uniform": scan_uniform, "schur": scan_schur}[sys.argv[1] if len(sys.argv) > 1 else "lump"]()
