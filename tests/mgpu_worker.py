"""Multi-GPU worker (launched by torchrun): z-slab sharded b13d steps must equal the single-GPU run.
Every rank steps its slab; rank 0 also steps the whole grid on its own GPU and compares."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pfhubbenchmarks_b200 import _lib, engine  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    grid = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    buf = torch.zeros(_lib.NCCL_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(engine.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    n = grid - 1
    ncell, length = (n, n + 3, n), (2.0 * n, 2.0 * (n + 3), 2.0 * n)
    ctx = engine.Context("b13d", ncell, length, device=local, rank=rank, nranks=world, nccl_id=buf.cpu().numpy().tobytes())
    # argv[2] = 0: one ghost-plane exchange per sweep instead of the wide-ghost preconditioner (both must equal 1 GPU)
    wide = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    ctx.set_option("use_wide", wide)
    nxy = (ncell[0] + 1) * (ncell[1] + 1)
    assert ctx.n_owned == nxy * (ctx.z_range[1] - ctx.z_range[0])
    w = ctx.initial_condition()
    # perturb with a z-dependent term so that slabs differ (the reference IC ignores z)
    z = torch.arange(ctx.z_range[0], ctx.z_range[1], device="cuda", dtype=torch.float64).repeat_interleave(nxy)
    w[:ctx.n_owned] += 0.01 * torch.cos(0.3 * z)
    hist = []
    for dt in (0.1, 0.2, 0.4):
        w0 = w.clone()
        ctx.dt = dt
        rep = ctx.newton_solve(w, w0, absolute_tolerance=1e-10, krylov_rtol=1e-9, maximum_iterations=20)
        assert rep["converged"], rep
        hist.append((rep["niters"], rep["krylov_its"]) + ctx.energy(w))
    # gather slabs on rank 0
    parts = [None] * world
    dist.all_gather_object(parts, (ctx.z_range, w.cpu().numpy()))
    # point evaluation and an assembled bilinear functional across slabs (every rank gets the global answer)
    pts = np.random.default_rng(11).random((257, 3)) * np.asarray(length)
    vals = ctx.eval_points(w, 0, pts)
    quad = ctx.bilinear(w, 0, 1, 0.5, 2.0)
    # one more step at the drivers' default tolerances: the compressed (fp32-stored) Krylov basis with its allreduced norms
    w0 = w.clone()
    ctx.dt = 0.4
    rep_c = ctx.newton_solve(w, w0)
    assert rep_c["converged"], rep_c
    e_c = ctx.energy(w)
    ctx.close()
    ok = True
    if rank == 0:
        one = engine.Context("b13d", ncell, length, device=local)
        v = one.initial_condition()
        zz = torch.arange(0, ncell[2] + 1, device="cuda", dtype=torch.float64).repeat_interleave(nxy)
        v[:one.n_owned] += 0.01 * torch.cos(0.3 * zz)
        ref_hist = []
        for dt in (0.1, 0.2, 0.4):
            v0 = v.clone()
            one.dt = dt
            rep = one.newton_solve(v, v0, absolute_tolerance=1e-10, krylov_rtol=1e-9, maximum_iterations=20)
            assert rep["converged"], rep
            ref_hist.append((rep["niters"], rep["krylov_its"]) + one.energy(v))
        ref = v.cpu().numpy().reshape(2, -1)
        got = np.zeros_like(ref)
        for (z0, z1), arr in parts:
            got[:, z0 * nxy:z1 * nxy] = arr.reshape(2, -1)
        err = [np.linalg.norm(got[f] - ref[f]) / np.linalg.norm(ref[f]) for f in range(2)]
        de = abs(hist[-1][2] / ref_hist[-1][2] - 1)
        dm = abs(hist[-1][3] / ref_hist[-1][3] - 1)
        print("mgpu world=%d grid=%d wide=%d  rel L2 (c, mu) = %.2e %.2e  dE %.1e dM %.1e  its %s vs %s" % (
            world, grid, wide, err[0], err[1], de, dm, [h[:2] for h in hist], [h[:2] for h in ref_hist]), flush=True)
        # N-GPU == 1-GPU to 1e-12 (the two runs differ only in the summation order of the
        # allreduced dot products).  Newton counts are identical.  Krylov counts are NOT comparable at this tolerance: a
        # linear solve needs about 30 iterations, GMRES(30) restarts at 30, and a solve that crosses the restart (round-off
        # decides on which side it lands) pays up to twenty extra iterations for the lost Krylov space -- in either run
        # (measured: 46 vs 67 and 96 vs 87 in the same trajectory).  The default-tolerance step below, whose solves stay
        # under the restart length, must agree to +-2.
        ok = err[0] < 1e-12 and err[1] < 1e-10 and de < 1e-12 and dm < 1e-13
        ok = ok and all(a[0] == b[0] and 0.5 * b[1] <= a[1] <= 2.0 * b[1] for a, b in zip(hist, ref_hist))
        dv = float(np.max(np.abs(vals - one.eval_points(v, 0, pts))))
        q1 = one.bilinear(v, 0, 1, 0.5, 2.0)
        v0 = v.clone()
        one.dt = 0.4
        rep_1 = one.newton_solve(v, v0)
        e_1 = one.energy(v)
        print("mgpu eval_points max diff %.1e  bilinear rel diff %.1e  default-tolerance step: its %s vs %s, dE %.1e" % (
            dv, abs(quad / q1 - 1), (rep_c["niters"], rep_c["krylov_its"]), (rep_1["niters"], rep_1["krylov_its"]),
            abs(e_c[0] / e_1[0] - 1)), flush=True)
        ok = ok and dv < 1e-9 and abs(quad / q1 - 1) < 1e-9 and rep_1["converged"] and abs(e_c[0] / e_1[0] - 1) < 1e-8 \
            and abs(rep_c["krylov_its"] - rep_1["krylov_its"]) <= 2
        one.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
