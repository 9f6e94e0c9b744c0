"""Host-side z-slab bookkeeping for the sharded 3-D cases (b13d on N GPUs).

The device path exchanges ghost planes with grouped ncclSend/ncclRecv inside libpfh_b200.so
(csrc/pfh_comm.cpp: comm_halo).  This module states the same ownership and exchange protocol
with torch.distributed on host tensors; it is what drivers use to scatter an initial field to the
ranks / gather slabs for output, and what the gloo tests run on CPU.

Ownership: the (ncell_z + 1) vertex planes are split into contiguous, balanced slabs; every plane is
owned by exactly one rank; the P1 Kuhn stencil reaches planes z-1 and z+1 only, so one ghost plane
per side suffices and the cube layer between two owners is evaluated by both (no partial-sum
exchange) -- SURVEY.md section 8(e).
"""
from .engine import slab_range  # noqa: F401  (re-exported)


def exchange_ghost_planes(owned, rank, world, dist, tag_base=0):
    """owned: CPU/GPU tensor [nfields, nz_local, nxy]. Returns (ghost_lo, ghost_hi), each [nfields, nxy]
    (zeros on a physical boundary): rank r sends its first plane to r-1 and its last plane to r+1."""
    import torch
    lo = torch.zeros_like(owned[:, 0])
    hi = torch.zeros_like(owned[:, 0])
    reqs = []
    first = owned[:, 0].contiguous()
    last = owned[:, -1].contiguous()
    if rank > 0:
        reqs.append(dist.isend(first, rank - 1, tag=tag_base + 1))
        reqs.append(dist.irecv(lo, rank - 1, tag=tag_base + 2))
    if rank < world - 1:
        reqs.append(dist.isend(last, rank + 1, tag=tag_base + 2))
        reqs.append(dist.irecv(hi, rank + 1, tag=tag_base + 1))
    for r in reqs:
        r.wait()
    return lo, hi


def scatter_global(field_global, nxy, rank, world):
    """Slice this rank's slab out of a global [nfields, nz*nxy] array (numpy or torch)."""
    nz = field_global.shape[-1] // nxy
    z0, z1 = slab_range(nz, rank, world)
    return field_global[..., z0 * nxy:z1 * nxy]


def gather_global(local, nxy, nz, rank, world, dist):
    """All ranks contribute [nfields, nz_local*nxy] CPU tensors; returns the global [nfields, nz*nxy] on every rank."""
    import torch
    parts = [None] * world
    dist.all_gather_object(parts, local)
    return torch.cat(parts, dim=-1)
