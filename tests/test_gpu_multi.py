"""N-GPU (z-slab + NCCL halo/allreduce) vs 1-GPU equality. Needs >= 2 GPUs on the box; skipped otherwise."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("world,grid,wide", [(2, 40, 1), (2, 40, 0), (2, 41, 1), (4, 40, 1), (8, 44, 1), (8, 66, 0)])
def test_slab_sharded_steps_equal_single_gpu(world, grid, wide):
    """grid 40 / 44 / 66: even vertex count (TMA loader); 41: odd (cp.async loader).  wide = 1: wide-ghost preconditioner,
    0: one exchange per sweep.  44 planes on 8 ranks: slabs of 5-6 planes, the thinnest the wide path accepts."""
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29610 + world), os.path.join(HERE, "mgpu_worker.py"), str(grid), str(wide)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
