"""BASELINE configs at full size on one GPU (or one Z-slab per rank under torchrun): timings + size-independent checks."""
import hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gsdf_b200
from gsdf_b200 import gsdf, gleval, glrender, slab

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lrank = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lrank); gsdf_b200.set_device(lrank)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
b = gsdf.Builder()
out = []
SCENES = [("npt-flange", 400), ("bolt", 800), ("knurled-cylinder", 1600)]
if len(sys.argv) > 1:
    SCENES = [x for x in SCENES if x[0] in sys.argv[1:]]
for name, resdiv in SCENES:
    s = gsdf.scene(b, name)
    sdf = gleval.NewCUDASDF3(s)
    res = np.float32(s.Diagonal() / np.float32(resdiv))
    lat = glrender.lattice_from_bounds(*s.Bounds(), res)
    cz = slab.rank_slab(lat.n[2], rank, world)
    t0 = time.perf_counter()
    R = glrender.Octree(sdf, res, cz_range=cz, stage_timing=True)
    t1 = time.perf_counter()
    for _ in range(3):
        R.Rerun()
    tm = R.Timings()
    nt, ev = R.NumTriangles(), R.Evaluations()
    t2 = time.perf_counter()
    tris = R.AllTriangles()
    t3 = time.perf_counter()
    sha = hashlib.sha256(tris.tobytes()).hexdigest()[:16]
    # every vertex lies on a lattice edge inside this rank's slab
    org = np.array(list(lat.origin), np.float64)
    rel = (tris.reshape(-1, 3)[:: max(1, len(tris) // 200000)].astype(np.float64) - org) / float(lat.res)
    ok = bool(((np.abs(rel - np.round(rel)) < 2e-3).sum(axis=1) >= 2).all() and rel[:, 2].min() >= cz[0] - 1e-3 and rel[:, 2].max() <= cz[1] + 1e-3)
    nx, ny, nz = lat.n
    dense = (nx + 1) * (ny + 1) * (cz[1] - cz[0] + 1)
    rec = dict(scene=name, resdiv=resdiv, lattice=[nx + 1, ny + 1, nz + 1], slab=list(cz), rank=rank, dense_corners=dense, evals=ev, tris=nt,
               first_ms=(t1 - t0) * 1e3, stage_ms=tm, d2h_ms=(t3 - t2) * 1e3, dense_equiv_Gevals_s=dense / tm["total_ms"] / 1e6,
               Mtris_s=nt / tm["total_ms"] / 1e3, vertices_on_lattice=ok, tri_sha=sha, pruned_unit_cubes=R.TotalPruned())
    if world > 1:
        tot = slab.total_count(nt)
        rec["tris_all_ranks"] = tot
    print(json.dumps(rec), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/full_size_w%d_r%d.jsonl" % (world, rank), "a") as f:
        f.write(json.dumps(rec) + "\n")
    R.Close(); sdf.Close(); del tris
    torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
