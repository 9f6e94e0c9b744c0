"""TEST INFRASTRUCTURE.  Golden trajectory of the 3-D benchmark (dolfin/b13d.py) from the oracle, tightly converged.

The reference commits no 3-D output (3-D parity is unpinned), so the fixture pins the PRODUCT to
the ORACLE over the north-star horizon: 100 backward-Euler steps of b13d on the reference mesh (BoxMesh 50^3 cells on
[0,100]^3, b13d.py:24-26 -> 51^3 vertices, odd: cp.async loader) and on a 51^3-cell mesh (52^3 vertices, even: TMA
loader), dt list of SURVEY.md section 8(d) (0.1, 0.2, 0.4, 0.8, 1.6, 1.6, ...).  Newton to |F| < 1e-10 with GMRES to 1e-10.

    python -m oracle.make_golden_b13d 50      # -> tests/golden/b13d_51.npz  (about an hour on one core)
    python -m oracle.make_golden_b13d 51      # -> tests/golden/b13d_52.npz

Stored: time, free energy and solute after every step; the concentration field after steps 10 and 100.
"""
import os
import sys
import time

import numpy as np

from oracle import krylov, mesh, problems, stepper


def dt_schedule(n):
    return [min(0.1 * 2 ** k, 1.6) for k in range(n)]


def main(ncell, nsteps=100, out_dir=None):
    L = 2.0 * ncell                                    # h = 2 as in b13d.py (100 / 50)
    m = mesh.kuhn3d(ncell, ncell, ncell, L, L, L)
    P = problems.CahnHilliard()
    D = stepper.Discretization(m, P)
    u = P.ic(m)
    t, times, E, C, snaps = 0.0, [], [], [], {}
    t0 = time.time()
    for k, dt in enumerate(dt_schedule(nsteps)):
        u, nit, ok = D.solve_step(u, u, dt, atol=1e-10, rtol=1e-14, max_it=25,
                                  linear=krylov.make_linear(rtol=1e-10, n_m=4, n_a=7, variant="upper", symbol_lmax=True))
        assert ok, (k, dt)
        t += dt
        times.append(t); E.append(D.energy(u)); C.append(D.mass(u))
        if k + 1 in (10, nsteps):
            snaps[k + 1] = u[0].copy()
        print("step %3d dt %.2f newton %d  E %.10f  C %.10f  (%.0f s)" % (k + 1, dt, nit, E[-1], C[-1], time.time() - t0), flush=True)
    out_dir = out_dir or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    path = os.path.join(out_dir, "b13d_%d.npz" % (ncell + 1))
    np.savez_compressed(path, ncell=ncell, length=L, time=np.array(times), energy=np.array(E), solute=np.array(C),
                        c_step10=snaps[10], c_final=snaps[nsteps], nsteps=nsteps)
    print("wrote", path)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 50, int(sys.argv[2]) if len(sys.argv) > 2 else 100)
