#!/usr/bin/env python3
"""bench.py -- the hot path on BASELINE.json's headline config: npt-flange at resdiv 400, tree -> pruned lattice
evaluation -> marching cubes -> triangles (README.md:116,130 of the reference).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step is one full render of the workload. Own arm (CUDA):
  value   dense-equivalent SDF evaluations per second: lattice corners of the config (6,711,685; what the
          reference's FlatRenderer evaluates, README.md:130) divided by the device time of the whole step
          (prune + evaluate + classify + scan + emit), program and buffers resident in HBM, L2 flushed between steps.
  e2e     the same metric through the reference-facing API with HOST buffers: upload the flattened tree, render,
          read every triangle back into (pinned) host memory; wall clock per step.
  N > 1   one process per GPU (torchrun). Default: every rank renders the full workload ("weak", independent
          renders, no collective on the data path); the Z-slab partition of ONE lattice across the ranks
          (north_star's layout, strong scaling) is timed as well and reported under "zslab".
Reference arm (--impl reference): the CPU oracle's FlatRenderer restatement (the reference is Go and cannot run here)
on all host threads, same config / metric / unit.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sdf_evals_per_sec"
UNIT = "evals/s"
RESDIV = 400
SCENE = "npt-flange"
KERNELS_PER_STEP = 7  # centres (+prune bits), compact, fine eval, mc-count (TMA), look-back scan, mc-emit, finish (publish counters + re-arm)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def build_scene():
    from gsdf_b200 import gsdf
    bld = gsdf.Builder()
    s = gsdf.scene(bld, SCENE)
    res = np.float32(s.Diagonal() / np.float32(RESDIV))
    return bld, s, res


def cpu_render(O, tree, lat, threads, reps):
    """The reference's CPU render step (gsdfaux.go:159-226): FlatRenderer grid evaluation on `threads` workers in
    4096-point batches + the serial marching-cubes sweep. Returns (seconds per render, evals, triangles)."""
    best = []
    ev = ntri = 0
    for _ in range(reps):
        t0 = time.perf_counter()
        grid, ev = O.flat_eval_grid(tree, lat, nthreads=threads, batch=4096)
        tris, _ = O.flat_march(lat, grid)
        best.append(time.perf_counter() - t0)
        ntri = len(tris)
    return best, ev, ntri


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    bld, s, res = build_scene()
    tree = O.Tree.from_shader(s)
    lat = O.flat_lattice(*s.Bounds(), res)
    cores = os.cpu_count() or 1
    threads = max(1, cores - 1)  # gsdfaux.go:161-164: max(1, GOMAXPROCS-1)
    cpu_render(O, tree, lat, threads, args.warmup)
    times, ev, ntri = cpu_render(O, tree, lat, threads, args.steps)
    sec = sum(times) / len(times)
    value = ev / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s resdiv %d: FlatRenderer dense lattice %dx%dx%d corners + serial marching cubes (CPU restatement of the reference; Go cannot run here)"
                   % (SCENE, RESDIV, lat.n[0] + 1, lat.n[1] + 1, lat.n[2] + 1), "evals_per_step": ev, "triangles_per_step": ntri},
        "triangles_per_sec": ntri / sec,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "full workload per step: %d lattice evaluations + marching cubes -> %d triangles, %d steps" % (ev, ntri, args.steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "readme_reference": {"evals_per_sec": 6711686 / 0.313, "hardware": "i5-12400F, 11 goroutines (README.md:128-130)"},
    }
    emit_json(line)


def bind_to_gpu_numa(local_rank):
    """Pin this rank's host threads to the CPUs local to its GPU (sysfs local_cpulist of the GPU's PCI function) before any
    pinned buffer is allocated, so first-touch places the triangle staging buffers on the GPU's NUMA node. With 8 ranks
    each moving 15 MB per step to the host, remote-node placement halves the end-to-end rate. No-op without NUMA info."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        cpus = set()
        for part in open(path).read().strip().split(","):
            if not part:
                continue
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus and len(cpus) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, cpus)
            return "%d cpus of %s" % (len(cpus), path)
    except Exception:
        pass
    return None


def run_cuda(args, rank, local_rank, world):
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's banner / debug lines must not share stdout with the JSON line
    import torch
    import gsdf_b200
    from gsdf_b200 import gsdf, gleval, glrender, _lib

    torch.cuda.set_device(local_rank)
    gsdf_b200.set_device(local_rank)
    numa = bind_to_gpu_numa(local_rank) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    bld, s, res = build_scene()
    sdf = gleval.NewCUDASDF3(s)
    R = glrender.NewOctreeRenderer(sdf, res, 1 << 15)
    nx, ny, nz = R.lat.n
    lattice_evals = (nx + 1) * (ny + 1) * (nz + 1)
    ntri = R.NumTriangles()
    evals_exec = R.Evaluations()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def l2_flush():
        flush.fill_(1)
        torch.cuda.synchronize()

    # ---------------- per-stage device times (roofline of the dominant kernel): same workload, launched eagerly with
    # CUDA events between the stages. The headline loop below replays the render as one CUDA graph, inside which
    # single kernels cannot be bracketed by events.
    Rt = glrender.Octree(sdf, res, stage_timing=True)
    stages = []
    for i in range(max(args.warmup, 3) + min(args.steps, 50)):
        l2_flush(); Rt.Rerun()
        if i >= max(args.warmup, 3):
            stages.append(Rt.Timings())
    assert Rt.NumTriangles() == ntri
    Rt.Close()

    # ---------------- device-resident steps (value)
    for _ in range(max(args.warmup, 3)):
        l2_flush(); R.Rerun()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    wall0 = time.perf_counter()
    step_ms = []
    for _ in range(args.steps):
        l2_flush()
        R.Rerun()
        step_ms.append(R.Timings()["total_ms"])
    barrier()
    wall = time.perf_counter() - wall0
    dev_ms = allmax(sum(step_ms))
    units = allsum(lattice_evals * args.steps)
    value = units / (dev_ms * 1e-3)
    tri_rate = allsum(ntri * args.steps) / (dev_ms * 1e-3)

    if args.device_only:
        if rank == 0:
            emit_json({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "ms_per_step": dev_ms / args.steps,
                       "note": "--device-only: profiling aid, not a bench line", "stage_ms": {k: sum(st[k] for st in stages) / len(stages) for k in stages[0]}})
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---------------- end to end through the public API with host buffers
    import ctypes as C
    host_tris = torch.empty((ntri + 8, 3, 3), dtype=torch.float32).pin_memory()
    host_np = host_tris.numpy()
    flat = bld.flatten(s)
    blob, aux = flat["blob"], np.ascontiguousarray(flat["aux"])
    h2d = len(blob) - 32 + aux.nbytes
    d2h = ntri * 36 + 32

    auxp = aux.ctypes.data_as(C.POINTER(C.c_float))

    def e2e_step():
        # upload the flattened tree (host -> device), render, read every triangle back to host memory: one
        # synchronous round trip per step, nothing carried over between steps
        _lib.check(_lib.lib.gsdf_program_update(sdf._h, blob, len(blob), auxp, aux.size))
        R.Rerun()
        got = 0
        while got < ntri:
            n = _lib.lib.gsdf_mesh_read(R._h, C.c_void_p(host_np[got:].ctypes.data), ntri + 8 - got)
            if n <= 0:
                break
            got += n
        return got

    def timed(step):
        for _ in range(max(args.warmup, 3)):
            l2_flush(); step()
        barrier()
        times = []
        for _ in range(args.steps):
            l2_flush()
            t0 = time.perf_counter()
            got = step()
            times.append(time.perf_counter() - t0)
            assert got == ntri
        barrier()
        return allmax(sum(times))

    single_sec = timed(e2e_step)
    ref_tris = host_np[:ntri].copy()

    # The same round trip through the Z-slab pipeline of the public API (glrender.SlabPipeline): the lattice is meshed as
    # three Z-slabs on this GPU and each slab's triangles are copied to the host under the next slab's kernels. In steady
    # state the copy sizes are predicted from the previous render and verified afterwards, so the step has no host
    # synchronisation between slabs. This is the headline e2e path.
    E2E_SLABS = 3
    pipe = glrender.SlabPipeline(sdf, res, nslabs=E2E_SLABS)
    host_np[:] = 0

    def e2e_pipe_step():
        _lib.check(_lib.lib.gsdf_program_update(sdf._h, blob, len(blob), auxp, aux.size))
        return pipe.RenderToHost(host_np)

    pipe_sec = timed(e2e_pipe_step)
    assert np.array_equal(host_np[:ntri].view(np.uint32), ref_tris.view(np.uint32)), "slab pipeline and single renderer disagree"
    pipe.Close()
    # Both are calls a user can make; the headline is the faster one on this box (with several ranks behind one PCIe
    # uplink the pipelined copies contend and the plain round trip can win). Both are always reported.
    e2e_sec = min(pipe_sec, single_sec)
    e2e_is_pipe = pipe_sec <= single_sec
    e2e_value = units / e2e_sec

    # Throughput form of the same loop (extra, not the headline): two renderers / two pinned buffers, the D2H copy of
    # step i (gsdf_mesh_read_async) overlaps the kernels of step i+1. Every step still uploads its tree and delivers
    # all its triangles to host memory inside the timed region.
    R2 = [R, glrender.NewOctreeRenderer(sdf, res, 1 << 15)]
    host2 = [host_np, torch.empty((ntri + 8, 3, 3), dtype=torch.float32).pin_memory().numpy()]

    def overlapped(k):
        for i in range(k):
            j = i & 1
            _lib.check(_lib.lib.gsdf_program_update(sdf._h, blob, len(blob), auxp, aux.size))
            R2[j].Rerun()   # waits for this renderer's previous copy before touching its triangle buffer
            n = _lib.lib.gsdf_mesh_read_async(R2[j]._h, C.c_void_p(host2[j].ctypes.data), ntri + 8)
            assert n == ntri
        for r in R2:
            _lib.check(_lib.lib.gsdf_mesh_wait(r._h))

    overlapped(4)
    barrier()
    t0 = time.perf_counter()
    overlapped(args.steps)
    ov_sec = allmax(time.perf_counter() - t0)
    barrier()
    assert np.array_equal(host2[1][:ntri].view(np.uint32), ref_tris.view(np.uint32)) and np.array_equal(host2[0][:ntri].view(np.uint32), ref_tris.view(np.uint32))
    clocks = sampler.stop() if rank == 0 else None
    e2e_value = units / e2e_sec

    # ---------------- Z-slab partition of ONE lattice across the ranks (north_star's layout; strong scaling)
    zslab = None
    if world > 1:
        from gsdf_b200 import slab
        cuts = slab.slab_cuts(nz, world)  # interior cuts aligned to the 4-cell prune blocks
        Rz = glrender.NewOctreeRenderer(sdf, res, 1 << 15, cz_range=(cuts[rank], cuts[rank + 1]))
        for _ in range(3):
            l2_flush(); Rz.Rerun()
        barrier()
        zs = []
        for _ in range(args.steps):
            l2_flush(); Rz.Rerun(); zs.append(Rz.Timings()["total_ms"])
        barrier()
        zms = allmax(sum(zs))
        ztri = allsum(Rz.NumTriangles())
        zslab = {"scaling": "strong", "value": lattice_evals * args.steps / (zms * 1e-3), "unit": UNIT, "ms_per_step": zms / args.steps,
                 "triangles_total": int(ztri), "cuts": cuts}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel, from the live CUDA-event stage times
    mean = {k: sum(st[k] for st in stages) / len(stages) for k in stages[0]}
    peak, peak_src = measured_peaks()
    fine_evals = evals_exec - ((nx + 3) // 4) * ((ny + 3) // 4) * ((nz + 3) // 4)
    cand = {
        "k_eval<GenGrid> (fine lattice evaluation)": (4.0 * fine_evals, mean["eval_ms"]),
        "k_mc_emit (marching-cubes emit)": (4.0 * fine_evals + 36.0 * ntri, mean["emit_ms"]),
    }
    kname = max(cand, key=lambda k: cand[k][1])
    kbytes, kms = cand[kname]
    achieved = kbytes / (kms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(kname.split(" ")[0])
        except Exception:
            traffic = None

    # What actually bounds the dominant kernel: warp-instruction issue. Warp instructions per launch come from the ncu
    # capture of this same command (profiles/traffic.json, smsp__inst_executed.sum); the duration is the live one.
    issue = None
    try:
        winst = json.load(open(tpath)).get(kname.split(" ")[0] + ".warp_inst")
        if winst and clocks and clocks.get("sm_mhz"):
            props = torch.cuda.get_device_properties(local_rank)
            peak_issue = props.multi_processor_count * 4 * clocks["sm_mhz"] * 1e6  # 4 schedulers/SM, 1 warp-instr/clk each
            ach = winst / (kms * 1e-3)
            issue = {"bound": "warp-instruction issue (FP32/ALU/XU pipes)", "achieved": ach / 1e9, "peak": peak_issue / 1e9, "unit": "G warp-instr/s",
                     "frac": ach / peak_issue, "warp_instructions_per_launch": winst,
                     "thread_instructions_per_eval": winst * 32 / max(fine_evals, 1),
                     "source": "smsp__inst_executed.sum from profiles/ (ncu --set full) / live CUDA-event kernel time; peak = SMs x 4 x SM clock"}
    except Exception:
        issue = None

    # ---------------- CPU baseline beside it (rank 0, N == 1 only): bounded sample of the same workload
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build()
        tree = O.Tree.from_shader(s)
        lat = O.flat_lattice(*s.Bounds(), res)
        threads = max(1, (os.cpu_count() or 1) - 1)
        # bounded sample: full renders of the same workload until ~12 s of CPU work have accumulated
        times, ev, nt = cpu_render(O, tree, lat, threads, 3)
        reps = max(5, min(400, int(12.0 / max(sum(times) / len(times), 1e-3))))
        times, ev, nt = cpu_render(O, tree, lat, threads, reps)
        assert nt == ntri, "CPU oracle and CUDA path disagree on the triangle count"
        sec = sum(times) / len(times)
        cpu = {"value": ev / sec, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d full renders of the workload (%d evaluations + marching cubes each), %.1f s of CPU work" % (reps, ev, sum(times)),
               "triangles_per_sec": nt / sec, "ms_per_render": sec * 1e3}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s resdiv %d: octree level-3 prune + marching cubes on a %dx%dx%d-corner lattice%s" %
                   (SCENE, RESDIV, nx + 1, ny + 1, nz + 1, "" if world == 1 else "; one full render per GPU per step"),
                   "renderer": "Octree (prune)", "evals_per_step_dense_equivalent": lattice_evals, "evals_executed_per_step": evals_exec,
                   "triangles_per_step": ntri, "l2": "flushed between steps (256 MiB write); working set 27 MB < 126 MB L2",
                   "timing": "CUDA events on the launching stream around each step (one CUDA-graph replay of 7 kernel nodes chained by programmatic dependent launch), summed over the timed steps, max over ranks"},
        "triangles_per_sec": tri_rate,
        "evals_executed_per_sec": allsum(evals_exec * args.steps) / (dev_ms * 1e-3) if world == 1 else None,
        "stage_ms": mean,
        "stage_ms_note": "eager launches with events between stages (separate loop of the same workload); the timed steps replay one CUDA graph",
        "wall_ms_per_step_incl_flush": wall * 1e3 / args.steps,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_sec * 1e3 / args.steps, "triangles_per_sec": ntri * world * args.steps / e2e_sec,
                "path": "slab_pipeline" if e2e_is_pipe else "single_renderer",
                "slab_pipeline": {"value": units / pipe_sec, "unit": UNIT, "ms_per_step": pipe_sec * 1e3 / args.steps,
                                  "path": "per step: gsdf_program_update(upload flattened tree) -> glrender.SlabPipeline(%d Z-slabs).RenderToHost: gsdf_mesh_rerun_begin + gsdf_mesh_read_prefix_async per slab (copy of slab i under the kernels of slab i+1, copy sizes predicted from the previous step and verified), all triangles in pinned host memory when the step returns" % E2E_SLABS},
                "single_renderer": {"value": units / single_sec, "unit": UNIT, "ms_per_step": single_sec * 1e3 / args.steps,
                                    "path": "per step: gsdf_program_update -> gsdf_mesh_rerun -> gsdf_mesh_read, one renderer, fully synchronous"},
                "overlapped": {"value": units / ov_sec, "unit": UNIT, "ms_per_step": ov_sec * 1e3 / args.steps,
                               "note": "same per-step work, D2H of step i overlapped with the kernels of step i+1 (gsdf_mesh_read_async, two renderers)"}},
        "gpu_launches": KERNELS_PER_STEP * args.steps,
        "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "algorithmic_bytes_per_launch": kbytes, "avg_launch_ms": kms, "peak_source": peak_src,
                     "note": "deep CSG trees are FP32-issue bound, not HBM bound (DESIGN.md); see profiles/ for issue-slot utilisation"},
        "roofline_issue": issue,
        "host_affinity": numa,
        "cpu_baseline": cpu,
        "clocks": clocks,
    }
    if zslab:
        line["zslab"] = zslab
    emit_json(line)
    if dist is not None:
        dist.destroy_process_group()


_JSON_OUT = None


def emit_json(line):
    """The contract is ONE JSON line on stdout. Libraries (NCCL's version banner when NCCL_DEBUG is set, torch warnings)
    also write to fd 1, so main() points fd 1 at stderr for the whole run and the line goes to the saved descriptor."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_OUT is not None:
        os.write(_JSON_OUT, data)
    else:
        sys.stdout.write(data.decode()); sys.stdout.flush()


def main():
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--device-only", action="store_true",
                    help="profiling aid: run only the device-resident loops (stage-timed + graph replay) and print their summary, no e2e legs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if args.steps > 10:
            args.steps = 10
        run_reference(args, rank, world)
    else:
        run_cuda(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
