"""Dual contouring of npt-flange across the ranks of a torchrun job (octant parts, no collective on the data path):
per-rank time, triangle counts, and a checksum of the concatenated mesh against the single-rank mesh on rank 0."""
import hashlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gsdf_b200
from gsdf_b200 import gsdf, gleval, glrender

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lrank = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lrank); gsdf_b200.set_device(lrank)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
resdiv = float(sys.argv[1]) if len(sys.argv) > 1 else 400
b = gsdf.Builder(); s = gsdf.scene(b, "npt-flange"); sdf = gleval.NewCUDASDF3(s)
res = np.float32(s.Diagonal() / np.float32(resdiv))
d = glrender.DualContourRenderer()
d.Reset(sdf, res, glrender.DualContourLeastSquares(Chiseled=True), part=rank, nparts=world)
for _ in range(3):
    d.Rerun()
st = d.Stats()
tris = d.RenderAll(None)
rec = dict(rank=rank, world=world, resdiv=resdiv, **st, sha=hashlib.sha256(tris.tobytes()).hexdigest()[:16])
if world > 1:
    t = torch.tensor([st["device_ms"]], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n = torch.tensor([st["triangles"]], device="cuda", dtype=torch.float64); dist.all_reduce(n, op=dist.ReduceOp.SUM)
    rec["max_ms_over_ranks"] = float(t.item()); rec["triangles_all_ranks"] = int(n.item())
print(json.dumps(rec), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/dc_w%d_r%d.jsonl" % (world, rank), "a") as f:
    f.write(json.dumps(rec) + "\n")
if world > 1:
    dist.destroy_process_group()
