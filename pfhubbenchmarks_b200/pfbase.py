"""`from pfhubbenchmarks_b200.pfbase import *`: the names a reference driver gets from `from pfbase import *`
(dolfin/pfbase.py): the initial conditions, the weak-form builders, and -- because the reference module itself does
`from dolfin import *` -- the UFL vocabulary (`dx`, `inner`, `grad`, `variable`, `MPI`, `HDF5File`, ...).

The builders return forms.Form objects; the strong forms they discretise (backward Euler, integrated by parts,
natural no-flux boundaries) are

    cahn_hilliard_weak_form   dc/dt = div(M grad mu),  mu = f'(c) - kappa lapl c          (pfbase.py:361-383)
    allen_cahn_weak_form      deta/dt = -L (f'(eta) - kappa lapl eta) + f                 (pfbase.py:396-408)
    allen_cahn_RHS_IBP        the right-hand side of the above, un-integrated             (pfbase.py:385-394)
    poisson_weak_form         div(M grad u) = f                                           (pfbase.py:410-421)
    diffusion_weak_form       du/dt = L div(D grad u + f1) + f2                           (pfbase.py:435-449)
    euler_bwd_weak_form       du/dt = f                                                   (pfbase.py:451-462)
"""
import numpy as np

from .api import (InitialConditionsBench1, InitialConditionsBench2, InitialConditionsBench3,  # noqa: F401
                  InitialConditionsBench6)
from .dolfin import *  # noqa: F401,F403
from .dolfin import MPI, HDF5File, UserExpression, dx, grad, inner  # noqa: F401


def _spinodal(x, c0, epsilon):
    return c0 + epsilon * (np.cos(0.105 * x[0]) * np.cos(0.11 * x[1]) + (np.cos(0.13 * x[0]) * np.cos(0.087 * x[1])) ** 2
                           + np.cos(0.025 * x[0] - 0.15 * x[1]) * np.cos(0.07 * x[0] - 0.02 * x[1]))


class _ReducedBench2(UserExpression):
    """Initial data of the reduced Ostwald-ripening drivers (bench2_eta1.py / bench2_eta2.py; pfbase.py:195-260): c as in
    benchmark 1, mu = 0, and the first N_ETA order parameters of the PFHub formula.  Evaluated on the host; a weak form
    with fewer than four order parameters is not one of the shipped systems, so it runs as a generated model."""
    N_ETA = 1

    def __init__(self, c0, epsilon, epsilon_eta, psi, **kwargs):
        super().__init__(**kwargs)
        self.c0, self.epsilon, self.epsilon_eta, self.psi = c0, epsilon, epsilon_eta, psi

    def eval(self, values, x):
        values[0] = _spinodal(x, self.c0, self.epsilon)
        values[1] = 0.0
        for i in range(self.N_ETA):
            k = i + 1.0
            wave = (np.cos(0.01 * k * x[0] - 4.0) * np.cos((0.007 + 0.01 * k) * x[1])
                    + np.cos((0.11 + 0.01 * k) * x[0]) * np.cos((0.11 + 0.01 * k) * x[1]))
            beat = np.cos((0.046 + 0.001 * k) * x[0] - (0.0405 + 0.001 * k) * x[1]) * np.cos((0.031 + 0.001 * k) * x[0] - (0.004 + 0.001 * k) * x[1])
            values[2 + i] = self.epsilon_eta * (wave + self.psi * beat ** 2) ** 2

    def value_shape(self):
        return (2 + self.N_ETA,)


class ICB2_jank1(_ReducedBench2):
    N_ETA = 1


class ICB2_jank2(_ReducedBench2):
    N_ETA = 2


class LangevinNoise(UserExpression):
    """Uniform noise of amplitude `amp` on three components (pfbase.py:341-356; unused by the drivers).  `seed` makes it
    reproducible, which the reference's version is not."""

    def __init__(self, amp, seed=None, **kwargs):
        super().__init__(**kwargs)
        self.amp, self._rng = amp, np.random.default_rng(seed)

    def eval(self, values, x):
        values[:3] = self.amp * self._rng.uniform(-0.5, 0.5, 3)

    def value_shape(self):
        return (3,)


def peval(f, x):
    """Parallel-safe f(x) (pfbase.py:50-62): the value, or +inf where the point is outside the mesh.  `f` is a scalar
    component, `w.sub(i)`.  On a slab-decomposed mesh the library already returns the value on every rank."""
    return np.array([f.at(np.asarray(x, dtype=float)[None, :])[0]])


def sample(u, Nx, Ny, Lx, Ly):
    """xs, ys, us[Nx, Ny] on a regular lattice of the rectangle (pfbase.py:68-94) -- all Nx*Ny points in one launch."""
    xs = np.linspace(0, Lx, Nx)
    ys = np.linspace(0, Ly, Ny)
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    us = u.at(np.stack([X.ravel(), Y.ravel()], axis=1)).reshape(Nx, Ny)
    return xs, ys, us


def _time_derivative(u, u_, u0, dt):
    return 1 / dt * u_ * (u - u0) * dx


def cahn_hilliard_weak_form(c, mu, c_, mu_, c0, dt, M, kappa, dfdc):
    transport = _time_derivative(c, c_, c0, dt) + inner(grad(c_), M * grad(mu)) * dx
    potential = mu_ * mu * dx - (mu_ * dfdc * dx + kappa * inner(grad(mu_), grad(c)) * dx)
    return transport + potential


def allen_cahn_RHS_IBP(eta, eta_, L, kappa, dfdeta, f):
    return -L * (eta_ * dfdeta + kappa * inner(grad(eta_), grad(eta))) + eta_ * f


def allen_cahn_weak_form(eta, eta_, eta0, dt, L, kappa, dfdeta, f):
    return _time_derivative(eta, eta_, eta0, dt) - allen_cahn_RHS_IBP(eta, eta_, L, kappa, dfdeta, f) * dx


def poisson_weak_form(u, u_, f, M):
    return -inner(grad(u_), M * grad(u)) * dx - u_ * f * dx


def diffusion_weak_form_RHS(u, u_, L, D, f1, f2):
    return -inner(grad(L * u_), D * grad(u) + f1) * dx + u_ * f2 * dx


def diffusion_weak_form(u, u_, u0, dt, L, D, f1, f2):
    return _time_derivative(u, u_, u0, dt) - diffusion_weak_form_RHS(u, u_, L, D, f1, f2)


def euler_bwd_weak_form(u, u_, u0, dt, f):
    return _time_derivative(u, u_, u0, dt) - u_ * f * dx


def h(u):
    """Interpolation polynomial u^3 (6u^2 - 15u + 10) of benchmark 2 (bench2.py:83-87)."""
    return u ** 3 * (6 * u ** 2 - 15 * u + 10)


def dh(u):
    """h'(u) = 30 u^2 (u - 1)^2, written as the reference writes it (pfbase.py:470-472): non-zero only in the interface."""
    return 3 * u ** 2 * (6 * u ** 2 - 15 * u + 10) + u ** 3 * (12 * u - 15)
