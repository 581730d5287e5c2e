"""Z-slab partition of one lattice across ranks (the reference's own CPU split is
FlatRenderer.evalGrid's k-slabs, glrender/flatrenderer.go:120-122).

Rank g of G meshes cells cz in [cuts[g], cuts[g+1]) and evaluates corner planes [cuts[g], cuts[g+1]] (one shared plane
between neighbours). There is NO collective on the data path: every rank produces its own triangle buffer; buffers
concatenated in rank order reproduce the single-device output (FlatRenderer cell order: z slowest). torch.distributed is
used only to gather results / counts at the end when the caller wants them in one place.
"""
import numpy as np


def slab_cuts(nz, world, align=4):
    """Cell-layer cut points [0, ..., nz] (gsdf_slab_cuts of the C ABI, the cuts gsdf_multi_begin uses): near-equal slabs;
    interior cuts are aligned down to the 4-layer prune blocks while slabs are at least two blocks thick, thinner slabs keep
    the exact split. More slabs than layers gives empty slabs at the end (cut == next cut)."""
    import ctypes as C
    from ._lib import lib, check
    if nz <= 0 or world <= 0:
        raise ValueError("nz and world must be positive")
    cuts = (C.c_int32 * (world + 1))()
    check(lib.gsdf_slab_cuts(int(nz), int(world), cuts))
    return [int(c) for c in cuts]


def rebalance_cuts(cuts, costs):
    """gsdf_slab_rebalance: new cuts that equalise the per-slab cost (evaluations executed, or milliseconds) each slab of
    `cuts` reported. Equal layers are not equal work on a pruned lattice."""
    import ctypes as C
    from ._lib import lib, check
    n = len(cuts) - 1
    cin = (C.c_int32 * (n + 1))(*[int(c) for c in cuts])
    cost = (C.c_double * n)(*[float(c) for c in costs])
    out = (C.c_int32 * (n + 1))()
    check(lib.gsdf_slab_rebalance(int(cuts[-1]), n, cin, cost, out))
    return [int(c) for c in out]


def rank_slab(nz, rank, world, align=4):
    cuts = slab_cuts(nz, world, align)
    return cuts[rank], cuts[rank + 1]


def gather_triangles(local_tris, group=None, dst=0):
    """Concatenate per-rank triangle arrays (n_i,3,3) on `dst` in rank order (= FlatRenderer cell order). Returns the
    concatenated array on dst, None elsewhere. Uses torch.distributed (gloo or nccl); object gather keeps it backend
    neutral -- this is result collection, not a data-path collective."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    parts = [None] * world if rank == dst else None
    dist.gather_object(np.ascontiguousarray(local_tris, dtype=np.float32), parts, dst=dst, group=group)
    if rank != dst:
        return None
    return np.concatenate([p.reshape(-1, 3, 3) for p in parts]) if parts else np.zeros((0, 3, 3), np.float32)


def total_count(n_local, group=None):
    """Sum of per-rank counts (triangles, evaluations) via all_reduce on a CPU tensor (gloo) or CUDA tensor (nccl)."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([int(n_local)], dtype=torch.int64, device=dev)
    dist.all_reduce(t, group=group)
    return int(t.item())


def octant_part(levels, part, nparts):
    """Dual-contouring partition (gsdf_dc_begin_part): the part's key range [k0, k1) in the octree's BFS cube order and
    the half-open box (lo[3], hi[3]) of cube indices it evaluates (own octants + one-cube border on each side). Host arithmetic only."""
    import ctypes as C
    from ._lib import lib, check
    keys = (C.c_uint32 * 2)()
    box = (C.c_int32 * 6)()
    check(lib.gsdf_dc_part_region(int(levels), int(part), int(nparts), keys, box))
    return (int(keys[0]), int(keys[1])), (tuple(box[:3]), tuple(box[3:]))
