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
        if rank == 0:
            whole, _ = O.flat_march(lat, grid, blockmask=mask)
            ok = total == len(whole) and allt.shape == whole.shape and np.array_equal(allt.view(np.uint32), whole.view(np.uint32))
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
