"""SNES-callback style boundary ("next" row N3): the explicit residual / Jacobian callbacks of
dolfinx/pfbase/pde_problems.py:113-185 (`SnesPDEProblem.F(snes, x, b)`, `.J(snes, x, J, P)`, `.vector()`),
matrix-free on top of the C ABI, so that an external nonlinear solver (PETSc SNES through a shell Mat / shell PC,
or the small torch Newton-GMRES below) can drive the GPU operators:

    F(snes, x, b)      -> pfh_residual_dev          b <- R(x; w0, dt)
    J(snes, x, J, P)   -> pfh_jac_setup_dev + pfh_precond_setup      (refresh the linearisation point)
    mult(x, y)         -> pfh_jac_apply_dev         y <- J(x_lin) x          (shell MatMult)
    pc_apply(r, z)     -> pfh_precond_apply_dev     z <- P^-1 r              (shell PCApply)
"""
import torch


class SnesPDEProblem:
    def __init__(self, ctx, w0, **pc_opts):
        self.ctx, self.w0, self.pc_opts = ctx, w0, pc_opts

    def vector(self):                       # pde_problems.py:179-181
        return self.ctx.new_vector()

    def F(self, snes, x, b):                # pde_problems.py:138-162
        self.ctx.residual(x, self.w0, out=b)

    def J(self, snes, x, J=None, P=None):   # pde_problems.py:164-177
        self.ctx.jac_setup(x)
        self.ctx.precond_setup(**self.pc_opts)

    def mult(self, x, y):
        self.ctx.jac_apply(x, out=y)

    def pc_apply(self, r, z):
        self.ctx.precond_apply(r, out=z)


def newton_krylov(problem, x, atol=1e-10, max_it=20, ksp_rtol=1e-8, restart=30, ksp_max_it=500):
    """A deliberately plain external solver (torch vector algebra, right-preconditioned GMRES) that uses only the
    callbacks above: what a PETSc SNES with a shell Mat/PC would do.  Returns (niters, converged)."""
    b = problem.vector()
    for it in range(max_it + 1):
        problem.F(None, x, b)
        if float(torch.linalg.vector_norm(b)) < atol:
            return it, True
        if it == max_it:
            break
        problem.J(None, x)
        dx = _gmres(problem, -b, ksp_rtol, restart, ksp_max_it)
        x += dx
    return max_it, False


def _gmres(problem, rhs, rtol, restart, max_it):
    n = rhs.numel()
    x = torch.zeros_like(rhs)
    bnorm = float(torch.linalg.vector_norm(rhs))
    r = rhs.clone()
    z, w = problem.vector(), problem.vector()
    its = 0
    while its < max_it:
        beta = float(torch.linalg.vector_norm(r))
        if beta <= rtol * bnorm:
            break
        V = torch.zeros(restart + 1, n, dtype=rhs.dtype, device=rhs.device)
        H = torch.zeros(restart + 1, restart, dtype=torch.float64)
        V[0] = r / beta
        k = 0
        for j in range(restart):
            problem.pc_apply(V[j].contiguous(), z)
            problem.mult(z, w)
            h = V[:j + 1] @ w
            wv = w - V[:j + 1].T @ h
            hn = float(torch.linalg.vector_norm(wv))
            H[:j + 1, j] = h.cpu()
            H[j + 1, j] = hn
            V[j + 1] = wv / hn
            its += 1
            k = j + 1
            e1 = torch.zeros(k + 1, dtype=torch.float64)
            e1[0] = beta
            y = torch.linalg.lstsq(H[:k + 1, :k], e1.unsqueeze(1)).solution.squeeze(1)
            res = float(torch.linalg.vector_norm(H[:k + 1, :k] @ y - e1))
            if res <= rtol * bnorm or its >= max_it:
                break
        upd = (V[:k].T @ y.to(rhs.device)).contiguous()
        problem.pc_apply(upd, z)
        x += z
        problem.mult(x, w)
        r = rhs - w
    return x
