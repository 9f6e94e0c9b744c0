"""world_size-2 (and 3) CPU test of the sharded path's host logic over the gloo backend: slab ownership,
ghost-plane exchange protocol (pfhubbenchmarks_b200/slab.py, mirrored in csrc/pfh_comm.cpp) and the
per-rank operator evaluation with ghosts (product per-vertex code compiled for the CPU, tests/emul)
reproduce the single-domain residual, Jacobian-vector product and energy/solute reductions."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
NCELL, LENGTH, PARAMS, DT = (5, 4, 8), (10.0, 8.0, 16.0), [0.3, 0.7, 5.0, 2.0, 5.0], 0.3


def _fields():
    rng = np.random.default_rng(12)
    nv = (NCELL[0] + 1) * (NCELL[1] + 1) * (NCELL[2] + 1)
    return 0.5 + 0.1 * rng.standard_normal((2, nv)), 0.5 + 0.1 * rng.standard_normal((2, nv)), rng.standard_normal((2, nv))


def _worker(rank, world, port, out_dir):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import emul
    from pfhubbenchmarks_b200 import slab
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    nxy = (NCELL[0] + 1) * (NCELL[1] + 1)
    nz = NCELL[2] + 1
    U, U0, X = _fields()
    z0, z1 = slab.slab_range(nz, rank, world)
    loc = [torch.from_numpy(np.ascontiguousarray(slab.scatter_global(A, nxy, rank, world))).reshape(2, z1 - z0, nxy)
           for A in (U, X, U0)]
    ghosts = [slab.exchange_ghost_planes(t, rank, world, dist, tag_base=10 * i) for i, t in enumerate(loc)]
    glo = np.stack([g[0].numpy() for g in ghosts])      # [3 vectors][nf][nxy], the layout tests/emul expects
    ghi = np.stack([g[1].numpy() for g in ghosts])
    kw = dict(z0=z0, nzl=z1 - z0, glo=glo, ghi=ghi, nf=2)
    flat = [t.reshape(2, -1).numpy() for t in loc]
    R = emul.run(3, 1, list(NCELL), list(LENGTH), PARAMS, DT, 0, U=flat[0], U0=flat[2], **kw)
    Y = emul.run(3, 1, list(NCELL), list(LENGTH), PARAMS, DT, 1, U=flat[0], X=flat[1], **kw)
    EM = torch.from_numpy(emul.run(3, 1, list(NCELL), list(LENGTH), PARAMS, DT, 2, U=flat[0], **kw).copy())
    dist.all_reduce(EM)                                  # ncclAllReduce(sum) of the partial reductions on the device path
    Rg = slab.gather_global(torch.from_numpy(R), nxy, nz, rank, world, dist)
    Yg = slab.gather_global(torch.from_numpy(Y), nxy, nz, rank, world, dist)
    if rank == 0:
        np.savez(os.path.join(out_dir, "out.npz"), R=Rg.numpy(), Y=Yg.numpy(), EM=EM.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_operators_equal_single_domain(world, tmp_path):
    import emul
    emul.build()
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "out.npz"))
    U, U0, X = _fields()
    R = emul.run(3, 1, list(NCELL), list(LENGTH), PARAMS, DT, 0, U=U, U0=U0, nf=2)
    Y = emul.run(3, 1, list(NCELL), list(LENGTH), PARAMS, DT, 1, U=U, X=X, nf=2)
    EM = emul.run(3, 1, list(NCELL), list(LENGTH), PARAMS, DT, 2, U=U, nf=2)
    assert np.abs(got["R"] - R).max() <= 1e-13 * np.abs(R).max()
    assert np.abs(got["Y"] - Y).max() <= 1e-13 * np.abs(Y).max()
    assert np.allclose(got["EM"], EM, rtol=1e-13, atol=0)
