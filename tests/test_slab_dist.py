"""N>1 host logic on CPU: world_size-2 gloo processes partition one lattice by Z-slab, each rank meshes its slab (the
CPU oracle stands in for the device mesher -- no GPU here), and rank 0 gathers. The concatenation must equal the
single-rank result bit for bit, and no collective is needed before the final gather."""
import os
import socket

import numpy as np
import pytest

from gsdf_b200 import slab


def test_slab_cuts_cover_and_align():
    for nz in (1, 3, 4, 10, 84, 85, 1407):
        for world in (1, 2, 3, 4, 8):
            cuts = slab.slab_cuts(nz, world)
            assert cuts[0] == 0 and cuts[-1] == nz and len(cuts) == world + 1
            assert all(a <= b for a, b in zip(cuts[:-1], cuts[1:]))
            if nz // world >= 8:   # slabs at least two prune blocks thick: interior cuts on block boundaries
                assert all(c % 4 == 0 for c in cuts[1:-1])
            assert sum(b - a for a, b in zip(cuts[:-1], cuts[1:])) == nz
    assert slab.slab_cuts(84, 8) == [0, 8, 20, 28, 40, 52, 60, 72, 84]
    assert slab.rank_slab(84, 1, 2) == (40, 84)
    with pytest.raises(ValueError):
        slab.slab_cuts(0, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch.distributed as dist
    from gsdf_b200 import gsdf
    from gsdf_b200 import slab as S
    from oracle import oracle as O
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        bld = gsdf.Builder()
        s = gsdf.scene(bld, "npt-flange")
        res = np.float32(s.Diagonal() / np.float32(90))
        tree = O.Tree.from_shader(s)
        lat = O.flat_lattice(*s.Bounds(), res)
        grid, _ = O.flat_eval_grid(tree, lat)
        mask, _ = O.octree_prune_mask(tree, lat)
        cz0, cz1 = S.rank_slab(lat.n[2], rank, world)
        mine, _ = O.flat_march(lat, grid, blockmask=mask, cz_range=(cz0, cz1))
        total = S.total_count(len(mine))
        allt = S.gather_triangles(mine, dst=0)
        # bench.py's set-up step: every rank contributes the cost of its slab (here: its triangle count), all ranks derive the
        # same re-balanced cuts from the gathered costs, and the re-cut slabs still concatenate to the whole mesh
        import torch
        cost = torch.zeros(world, dtype=torch.float64)
        cost[rank] = len(mine)
        dist.all_reduce(cost)
        cuts = S.slab_cuts(lat.n[2], world)
        new = S.rebalance_cuts(cuts, [float(v) for v in cost.tolist()])
        mine2, _ = O.flat_march(lat, grid, blockmask=mask, cz_range=(new[rank], new[rank + 1]))
        allt2 = S.gather_triangles(mine2, dst=0)
        counts = torch.zeros(world, dtype=torch.float64)
        counts[rank] = len(mine2)
        dist.all_reduce(counts)
        if rank == 0:
            whole, _ = O.flat_march(lat, grid, blockmask=mask)
            ok = total == len(whole) and allt.shape == whole.shape and np.array_equal(allt.view(np.uint32), whole.view(np.uint32))
            ok = ok and allt2.shape == whole.shape and np.array_equal(allt2.view(np.uint32), whole.view(np.uint32))
            ok = ok and new[0] == 0 and new[-1] == lat.n[2] and float(counts.max()) <= float(cost.max())   # no worse balanced than before
            with open(out_path, "w") as f:
                f.write("ok %d %d" % (total, len(whole)) if ok else "mismatch %d %d" % (total, len(whole)))
        else:
            assert allt is None
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_slab_gather(tmp_path, oracle):
    import torch.multiprocessing as mp
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    txt = open(out).read()
    assert txt.startswith("ok"), txt
    assert int(txt.split()[1]) > 1000


def _bfs_unkey(key, bits):
    """Octree BFS cube order -> cube index: child order of i3.Cube.Octree() = Bourke corner order, top level first."""
    x = y = z = 0
    for b in range(bits - 1, -1, -1):
        c = (key >> (3 * b)) & 7
        zb, r = c >> 2, c & 3
        yb = r >> 1
        xb = 3 - r if yb else r
        x |= xb << b
        y |= yb << b
        z |= zb << b
    return x, y, z


def test_dual_contour_octant_parts_cover_and_border():
    """gsdf_dc_part_region (host arithmetic of gsdf_dc_begin_part): the parts' key ranges tile the BFS order; every cube a
    part owns, its -1 neighbours (quad corners) and their +1 neighbours (QEF edge data: at most the cube's own +1) lie
    inside the part's box."""
    for levels in (3, 4, 6):
        bits = levels - 1
        N = 1 << bits
        for nparts in (1, 2, 4, 8):
            ranges = []
            for part in range(nparts):
                (k0, k1), (lo, hi) = slab.octant_part(levels, part, nparts)
                ranges.append((k0, k1))
                assert all(0 <= lo[a] < hi[a] <= N for a in range(3))
                step = max(1, (k1 - k0) // 4096)
                for key in list(range(k0, k1, step)) + [k1 - 1]:
                    c = _bfs_unkey(key, bits)
                    for a in range(3):
                        assert max(c[a] - 1, 0) >= lo[a]
                        assert min(c[a] + 1, N - 1) < hi[a]
                if nparts == 8:  # one octant plus its border, not the whole grid
                    assert all(hi[a] - lo[a] <= N // 2 + 2 for a in range(3))
            assert ranges[0][0] == 0 and ranges[-1][1] == N ** 3
            assert all(a[1] == b[0] for a, b in zip(ranges[:-1], ranges[1:]))
    import gsdf_b200
    with pytest.raises(gsdf_b200.GsdfError):
        slab.octant_part(5, 0, 3)
    with pytest.raises(gsdf_b200.GsdfError):
        slab.octant_part(5, 4, 4)


def _dc_worker(rank, world, port, out_path):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch.distributed as dist
    from gsdf_b200 import gsdf
    from gsdf_b200 import slab as S
    from oracle import oracle as O
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        bld = gsdf.Builder()
        s = bld.Union(bld.NewSphere(0.4), bld.Translate(bld.NewSphere(0.3), 0, 0, 0.45), bld.Translate(bld.NewSphere(0.2), 0, 0, 0.8))
        res = np.float32(3.0 / 64)
        tree = O.Tree.from_shader(s)
        whole, st = O.dual_contour(tree, *s.Bounds(), res, O.DC_LSQ_CHISELED)
        # the oracle meshes globally; a rank keeps the triangles of the cubes it owns. Two triangles per quad, quads in cube
        # (= BFS key) order: ownership is decided by the quad's owning cube, recovered from the triangle's first vertex run.
        (k0, k1), _ = S.octant_part(st["levels"], rank, world)
        keys = O.dual_contour_quad_keys(tree, *s.Bounds(), res, O.DC_LSQ_CHISELED)
        mine = whole[np.repeat((keys >= k0) & (keys < k1), 2)]
        total = S.total_count(len(mine))
        allt = S.gather_triangles(mine, dst=0)
        if rank == 0:
            ok = total == len(whole) and np.array_equal(allt.view(np.uint32), whole.view(np.uint32))
            with open(out_path, "w") as f:
                f.write("ok %d" % total if ok else "mismatch %d %d" % (total, len(whole)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_dual_contour_parts(tmp_path, oracle):
    """The dual-contour mesh split by owned key range over 2 gloo ranks and gathered in rank order equals the whole mesh
    (the oracle stands in for the device renderer; quads are emitted in BFS key order, so ranges concatenate)."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "dc.txt")
    mp.spawn(_dc_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    txt = open(out).read()
    assert txt.startswith("ok"), txt
    assert int(txt.split()[1]) > 1000
