"""3-D Cahn-Hilliard on a box (DOLFIN BoxMesh, 6 tets per cube) -- mirrors dolfin/b13d.py.

    python -m pfhubbenchmarks_b200.drivers.b13d [--backend b200|fenics] [--end-time 50] [--n 50]
    torchrun --nproc-per-node N -m pfhubbenchmarks_b200.drivers.b13d      (z-slab sharded over N GPUs)
"""
import os
import sys

from .. import dolfin as df
from ..pfbase import InitialConditionsBench1, cahn_hilliard_weak_form, dx, grad, inner
from .common import mixed_p1, parse_args, run_fenics, snes_setup, time_loop


def main():
    n_cells = 50
    if "--n" in sys.argv:
        i = sys.argv.index("--n")
        n_cells = int(sys.argv[i + 1])
        del sys.argv[i:i + 2]
    args = parse_args("b13d", 50.0)
    if args.backend == "fenics":
        return run_fenics(args, "b13d.py")
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    nccl_id = None
    if world > 1:
        import torch
        import torch.distributed as dist
        from ..engine import nccl_unique_id
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        nccl_id = buf.cpu().numpy().tobytes()
    # mesh (b13d.py:24-26): 100^3 box, 50^3 cells (h = 2)
    Nx = Ny = Nz = n_cells
    Lx = Ly = Lz = 2.0 * n_cells
    mesh = df.BoxMesh(df.Point(0.0, 0.0, 0.0), df.Point(Lx, Ly, Lz), Nx, Ny, Nz)
    dt = df.Constant(1e-1)
    c_alpha, c_beta, rho_s, kappa, M = (df.Constant(v) for v in (0.3, 0.7, 5.0, 2.0, 5.0))        # b13d.py:35-39
    W, w, w0, w_, dw = mixed_p1(mesh, 2)
    w_ic = InitialConditionsBench1(0.5, 0.05, degree=2)
    c, mu = df.split(w)
    f_chem = rho_s * (c - c_alpha) ** 2 * (c_beta - c) ** 2
    F = cahn_hilliard_weak_form(c, mu, w_[0], w_[1], w0[0], dt, M, kappa, df.diff(f_chem, c))
    # device / rank / nranks / nccl_id place the problem: one z-slab of the box per GPU
    problem = df.NonlinearVariationalProblem(F, w, [], df.derivative(F, w, dw), device=local, rank=rank, nranks=world,
                                             nccl_id=nccl_id)
    solver = df.NonlinearVariationalSolver(problem)
    snes_setup(solver, args.atol, "basic", int(Nx * Ny * Nz / 10))      # b13d.py:97

    def total_free_energy():
        return df.assemble(f_chem * dx + kappa / 2.0 * inner(grad(c), grad(c)) * dx)

    def total_solute():
        return df.assemble(c * dx)

    # b13d.py writes results/bench1_out.csv (b13d.py:200)
    return time_loop(problem, solver, w_ic, 1e-1, 1e-2, args.end_time, os.path.join(args.outdir, "bench1_out.csv"),
                     "time,total_free_energy,total_solute", args.max_steps, args.quiet,
                     save_dir=os.path.join(args.outdir, "b13d") if args.save_solution else None,
                     functionals=(total_free_energy, total_solute))


if __name__ == "__main__":
    main()
