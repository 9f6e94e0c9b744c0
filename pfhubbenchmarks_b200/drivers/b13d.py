"""3-D Cahn-Hilliard on a box (DOLFIN BoxMesh, 6 tets per cube) -- mirrors dolfin/b13d.py.

    python -m pfhubbenchmarks_b200.drivers.b13d [--backend b200|fenics] [--end-time 50] [--n 50]
    torchrun --nproc-per-node N -m pfhubbenchmarks_b200.drivers.b13d      (z-slab sharded over N GPUs)
"""
import os
import sys

from .. import api as df
from .common import parse_args, run_fenics, time_loop


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
    c_alpha, c_beta, rho_s, kappa, M = 0.3, 0.7, 5.0, 2.0, 5.0
    w_ic = df.InitialConditionsBench1(0.5, 0.05, degree=2)
    problem = df.NonlinearVariationalProblem("b13d", mesh, dt, params=[c_alpha, c_beta, rho_s, kappa, M], device=local,
                                             rank=rank, nranks=world, nccl_id=nccl_id)
    solver = df.NonlinearVariationalSolver(problem)
    solver.parameters["nonlinear_solver"] = "snes"
    nlparams = solver.parameters["snes_solver"]
    nlparams["line_search"] = "basic"
    nlparams["report"] = True
    nlparams["error_on_nonconvergence"] = False
    nlparams["absolute_tolerance"] = args.atol
    nlparams["maximum_iterations"] = 10
    nlparams["linear_solver"] = "gmres"
    nlparams["preconditioner"] = "sor"
    nlparams["krylov_solver"]["maximum_iterations"] = int(Nx * Ny * Nz / 10)      # b13d.py:97
    nlparams["krylov_solver"]["error_on_nonconvergence"] = False
    # b13d.py writes results/bench1_out.csv (b13d.py:200)
    return time_loop(problem, solver, w_ic, 1e-1, 1e-2, args.end_time, os.path.join(args.outdir, "bench1_out.csv"),
                     "time,total_free_energy,total_solute", args.max_steps, args.quiet,
                     save_dir=os.path.join(args.outdir, os.path.basename(__file__)[:-3]) if args.save_solution else None)


if __name__ == "__main__":
    main()
