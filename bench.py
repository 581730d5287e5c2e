#!/usr/bin/env python3
"""bench.py -- the hot path on BASELINE.json's headline config: npt-flange at resdiv 400, tree -> coarse-to-fine pruned
lattice evaluation -> marching cubes -> triangles (README.md:116,130 of the reference).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step is ONE render of ONE lattice (281x281x85 corners, 423,852 triangles).
  value   dense-equivalent SDF evaluations per second: the lattice corners of the config (6,711,685 -- what the reference's
          CPU path evaluates, README.md:130) divided by the device time of the step, program and buffers resident in HBM,
          L2 flushed between steps. The evaluations actually executed (the prune skips most corners) are reported beside it
          as evals_executed_per_sec; the ratio against the CPU arm is a ratio of times to the same triangles.
  e2e     the same metric through the reference-facing C ABI with HOST buffers: upload the flattened tree, render, all
          triangles in one pinned host buffer in FlatRenderer order; wall clock per step (gsdf_multi_update +
          gsdf_multi_render: Z-slabs pipelined against their own read-back, counts read, nothing predicted).
  N > 1   STRONG scaling: the one lattice is split into N Z-slabs (north_star's partition), one process per GPU under
          torchrun, no collective on the data path; value = lattice corners / max-over-ranks device time. e2e at N > 1 is
          the in-process multi-device driver (gsdf_multi_*: one host thread per GPU) run by rank 0 over all N GPUs while the
          other ranks idle at a barrier -- the call a Go program makes.
Reference arm (--impl reference): the CPU oracle's FlatRenderer restatement (the reference is Go and cannot run here) on
all host threads, same config / metric / unit; it maps only the CUDA-free host library.
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

METRIC = "dense_equiv_sdf_evals_per_sec"
UNIT = "evals/s"
RESDIV = 400
SCENE = "npt-flange"


def workload_config(nx, ny, nz, evals, ntri):
    """The same dict in both arms."""
    return {"workload": "%s resdiv %d: %dx%dx%d-corner lattice -> marching-cubes triangles (README.md:116,130)" % (SCENE, RESDIV, nx + 1, ny + 1, nz + 1),
            "evals_dense_equivalent_per_step": int(evals), "triangles_per_step": int(ntri)}


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


def build_scene(name=SCENE, resdiv=RESDIV):
    from gsdf_b200 import gsdf
    bld = gsdf.Builder()
    s = gsdf.scene(bld, name)
    res = np.float32(s.Diagonal() / np.float32(resdiv))
    return bld, s, res


def cpu_render(O, tree, lat, threads, reps):
    """The reference's CPU render step (gsdfaux.go:159-226): FlatRenderer grid evaluation on `threads` workers in
    4096-point batches + the serial marching-cubes sweep. Returns (seconds per render list, evals, triangles)."""
    times = []
    ev = ntri = 0
    for _ in range(reps):
        t0 = time.perf_counter()
        grid, ev = O.flat_eval_grid(tree, lat, nthreads=threads, batch=4096)
        tris, _ = O.flat_march(lat, grid)
        times.append(time.perf_counter() - t0)
        ntri = len(tris)
    return times, ev, ntri


def run_reference(args, rank, world):
    if rank != 0:
        return
    os.environ["GSDF_HOST_ONLY"] = "1"   # scene builders only: the CUDA library is never mapped in this arm
    from oracle import oracle as O
    O.build()
    bld, s, res = build_scene()
    tree = O.Tree.from_shader(s)
    lat = O.flat_lattice(*s.Bounds(), res)
    cores = os.cpu_count() or 1
    threads = max(1, cores - 1)  # gsdfaux.go:161-164: max(1, GOMAXPROCS-1)
    cpu_render(O, tree, lat, threads, max(args.warmup, 1))
    times, ev, ntri = cpu_render(O, tree, lat, threads, args.steps)
    sec = sum(times) / len(times)
    value = ev / sec
    nx, ny, nz = lat.n
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(nx, ny, nz, ev, ntri),
        "arm": {"renderer": "FlatRenderer (dense lattice on %d threads, 4096-point batches) + serial marching cubes: CPU restatement of the reference's path (Go cannot run here)" % threads},
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
    pinned buffer is allocated, so first-touch places the triangle staging buffers on the GPU's NUMA node. No-op without
    NUMA information."""
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
    import ctypes as C
    import torch
    import gsdf_b200
    from gsdf_b200 import gleval, glrender, slab, _lib

    torch.cuda.set_device(local_rank)
    gsdf_b200.set_device(local_rank)
    numa = bind_to_gpu_numa(local_rank) if world > 1 and args.numa else None
    dist = None
    cpu_group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")   # host-side waits: an NCCL barrier would spin on the idle ranks' GPUs

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def host_barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier(group=cpu_group)

    def allreduce(x, op):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=op)
        return float(t.item())

    def allmax(x):
        return allreduce(x, dist.ReduceOp.MAX) if dist is not None else float(x)

    def allsum(x):
        return allreduce(x, dist.ReduceOp.SUM) if dist is not None else float(x)

    warm = max(args.warmup, 3)
    bld, s, res = build_scene()
    sdf = gleval.NewCUDASDF3(s)
    # what constructing the GPU evaluator does in the reference (it compiles the tree's shader, gleval/gpu.go:35-54): kernels
    # specialised for this tree's instruction stream, compiled by NVRTC (a few seconds, outside every timed region). Where
    # NVRTC is missing the interpreter kernels run (same results); the line says which.
    specialised = (not args.no_specialize) and sdf.Specialize()
    lat = glrender.lattice_from_bounds(*s.Bounds(), res)
    nx, ny, nz = lat.n
    lattice_evals = (nx + 1) * (ny + 1) * (nz + 1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def l2_flush():
        flush.fill_(1)
        torch.cuda.synchronize()

    # ---------------- device-resident steps (value): this rank's Z-slab of the ONE lattice (the whole lattice at N = 1)
    cuts = slab.slab_cuts(nz, world)
    R = glrender.NewOctreeRenderer(sdf, res, 1 << 15, cz_range=(cuts[rank], cuts[rank + 1]))
    plan = R.Plan()
    if world > 1 and not args.no_rebalance:
        # equal layers are not equal work (the flange's plate faces sit in a few layers): two rounds of re-cutting by the
        # evaluations each slab executed, from counts every rank reads off its own renderer (an all_gather of N integers
        # at set-up time -- not on the data path)
        for _ in range(2):
            ev = torch.zeros(world, dtype=torch.float64, device="cuda")
            ev[rank] = R.Evaluations()
            dist.all_reduce(ev)
            new = slab.rebalance_cuts(cuts, [float(v) for v in ev.tolist()])
            if new == cuts:
                break
            cuts = new
            R.Close()
            R = glrender.NewOctreeRenderer(sdf, res, 1 << 15, cz_range=(cuts[rank], cuts[rank + 1]))
    ntri_local = R.NumTriangles()
    evals_local = R.Evaluations()
    ntri = int(allsum(ntri_local))
    evals_exec = int(allsum(evals_local))
    for _ in range(warm):
        l2_flush(); R.Rerun()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    wall0 = time.perf_counter()
    step_ms, stages = [], []
    for _ in range(args.steps):
        l2_flush()
        R.Rerun()
        t = R.Timings()
        step_ms.append(t["total_ms"])
        stages.append(t)
    barrier()
    wall = time.perf_counter() - wall0
    dev_ms = allmax(sum(step_ms))
    value = lattice_evals * args.steps / (dev_ms * 1e-3)
    mean = {k: sum(st[k] for st in stages) / len(stages) for k in stages[0]}
    # sustained leg (an extra, not the headline): back-to-back renders without the L2 flush, long enough (~0.4 s of kernels) for the
    # clocks and the power state to settle -- what a caller that meshes in a loop sees
    sustained = None
    if world == 1:
        n_s, tot = 2000, 0.0
        t0 = time.perf_counter()
        for _ in range(n_s):
            R.Rerun()
            tot += R.Timings()["total_ms"]
        sustained = {"steps": n_s, "ms_per_step": tot / n_s, "wall_ms_per_step": (time.perf_counter() - t0) * 1e3 / n_s,
                     "note": "back-to-back graph replays, no L2 flush (working set 27 MB stays in the 126 MB L2), each followed by the host's read of the counters"}
    kernels_per_step = len(plan) + 4  # one launch per centre level, work lists, lattice evaluation, count pass, emit pass (which scans and ends the render)

    if args.device_only:
        if rank == 0:
            emit_json({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "ms_per_step": dev_ms / args.steps,
                       "note": "--device-only: profiling aid, not a bench line", "stage_ms": mean})
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---------------- end to end through the C ABI with host buffers: rank 0 drives all N GPUs from one process
    flat = bld.flatten(s)
    blob, aux = flat["blob"], np.ascontiguousarray(flat["aux"])
    h2d = (len(blob) - 32 + aux.nbytes) * world
    d2h = ntri * 36
    e2e = None
    e2e_launches = 0
    host_barrier()   # ranks > 0 now sleep in a socket wait: their GPUs are idle while rank 0 drives all of them
    if rank == 0:
        host = glrender.pinned_empty((ntri + 8, 3, 3))
        ref_tris = None
        table = {}
        # the e2e steps run on every GPU of the box from this process: flush the L2 of each of them between steps
        flushes = [flush] + [torch.empty(256 << 20, dtype=torch.uint8, device="cuda:%d" % i) for i in range(1, world)]

        def l2_flush_all():
            for i, f in enumerate(flushes):
                f.fill_(1)
            for i in range(world):
                torch.cuda.synchronize(i)

        for spd in ((1, 2, 3, 4) if world == 1 else (1, 2)):
            M = glrender.MultiRenderer(s, res, devices=list(range(world)), slabs_per_device=spd)
            assert M.NumTriangles() == ntri
            if specialised:
                M.Specialize()
            if spd * world > 1 and not args.no_rebalance:
                M.Rebalance(2)

            def step():
                M.UpdateBlob(blob, aux)          # host -> device: the flattened tree, to every device
                return M.RenderToHost(host)      # render + device -> host: every triangle, slab order, one buffer

            for _ in range(warm):
                l2_flush_all(); step()
            times = []
            for _ in range(args.steps):
                l2_flush_all()
                host[:8] = 0
                t0 = time.perf_counter()
                got = step()
                times.append(time.perf_counter() - t0)
                assert got == ntri
            if ref_tris is None:
                ref_tris = glrender.Octree(sdf, res).AllTriangles()
            assert np.array_equal(host[:ntri].view(np.uint32), ref_tris.view(np.uint32)), "multi-slab output differs from the single renderer"
            nsl = len(M.Slabs()[1])
            tl = M.Timeline()   # host-clock microseconds of the LAST timed step: slabs enqueued, per slab (count seen, copy enqueued), delivered
            table["slabs_per_device_%d" % spd] = {"ms_per_step": sum(times) / len(times) * 1e3, "slabs": nsl, "device_ms": M.DeviceMs(), "cuts": M.Slabs()[0],
                                                  "last_step_timeline_us": {"enqueued": round(tl["enqueued"], 1), "slabs": [[round(a, 1), round(c, 1)] for a, c in tl["slabs"]],
                                                                            "delivered": round(tl["delivered"], 1)}}
            e2e_launches = max(e2e_launches, nsl * kernels_per_step)
            M.Close()
        best = min(table, key=lambda k: table[k]["ms_per_step"])
        sec = table[best]["ms_per_step"] * 1e-3
        e2e = {"value": lattice_evals / sec, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": sec * 1e3,
               "triangles_per_sec": ntri / sec, "chosen": best, "by_slabs_per_device": table,
               "path": "per step: gsdf_multi_update (flattened tree to every device) -> gsdf_multi_render into ONE pinned host buffer: every slab on its own "
                       "stream, counts read from the device (no prediction), slab i's read-back under the kernels of the later slabs; "
                       "slabs_per_device_1 at N=1 is the fully synchronous single-renderer round trip"}
    host_barrier()

    # ---------------- gleval.SDF3.Evaluate on host slices (the contract north_star names first), rank 0, N = 1 only
    evaluate = None
    if rank == 0 and world == 1:
        evaluate = bench_evaluate(sdf, s, glrender, torch)

    # ---------------- knurled-cylinder resdiv 1600 (BASELINE config 4) by Z-slab: strong scaling where slabs pay
    knurled = None
    if not args.no_knurled:
        knurled = bench_knurled(rank, world, barrier, allmax, allsum, l2_flush, gleval, glrender, slab, specialised)

    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel, from the stage times of the timed steps themselves (the kernels
    # stamp %globaltimer inside the graph replay)
    peak, peak_src = measured_peaks()
    # executed = prune-cube centres + 4 per listed lattice quad; with the one-level plan on the whole lattice every level-3
    # centre is evaluated, otherwise the centres (a few per cent) are left in
    centres = ((nx + 3) // 4) * ((ny + 3) // 4) * ((nz + 3) // 4) if len(plan) == 1 and world == 1 else 0
    fine_evals = max(evals_exec - centres, 0)
    kname = ("k_jit_grid2 (fine lattice evaluation, kernel specialised for the tree at run time, two corners per thread)" if specialised
             else "k_eval<GenGrid> (fine lattice evaluation, interpreter)")
    kbytes, kms = 4.0 * fine_evals / world, mean["eval_ms"]
    achieved = kbytes / (kms * 1e-3) / 1e9
    traffic = winst = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            tk = "k_jit_grid2" if specialised else "k_eval<GenGrid>"
            traffic, winst = tj.get(tk), tj.get(tk + ".warp_inst")
        except Exception:
            pass
    issue = None
    if winst and clocks and clocks.get("sm_mhz") and world == 1:
        props = torch.cuda.get_device_properties(local_rank)
        peak_issue = props.multi_processor_count * 4 * clocks["sm_mhz"] * 1e6  # 4 schedulers/SM, 1 warp-instr/clk each
        ach = winst / (kms * 1e-3)
        issue = {"bound": "warp-instruction issue (FP32/ALU/XU pipes)", "achieved": ach / 1e9, "peak": peak_issue / 1e9, "unit": "G warp-instr/s",
                 "frac": ach / peak_issue, "warp_instructions_per_launch": winst, "thread_instructions_per_eval": winst * 32 / max(fine_evals, 1),
                 "source": "smsp__inst_executed.sum of the ncu capture under profiles/ / the live in-graph kernel time; peak = SMs x 4 x SM clock"}

    # ---------------- CPU baseline beside it (rank 0, N == 1 only): bounded sample of the same workload
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build()
        tree = O.Tree.from_shader(s)
        olat = O.flat_lattice(*s.Bounds(), res)
        threads = max(1, (os.cpu_count() or 1) - 1)
        times, ev, nt = cpu_render(O, tree, olat, threads, 3)
        reps = max(5, min(400, int(12.0 / max(sum(times) / len(times), 1e-3))))
        times, ev, nt = cpu_render(O, tree, olat, threads, reps)
        assert nt == ntri, "CPU oracle and CUDA path disagree on the triangle count"
        sec = sum(times) / len(times)
        cpu = {"value": ev / sec, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d full renders of the workload (%d evaluations + marching cubes each), %.1f s of CPU work" % (reps, ev, sum(times)),
               "triangles_per_sec": nt / sec, "ms_per_render": sec * 1e3}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(nx, ny, nz, lattice_evals, ntri),
        "arm": {"renderer": "Octree: coarse-to-fine prune %s + marching cubes" % plan, "evals_executed_per_step": evals_exec,
                "kernels": ("lattice evaluation and prune-centre pass specialised for the tree at run time (gsdf_program_specialize: NVRTC, compiled before "
                            "the timed region, as the reference compiles its shader at construction)" if specialised else "interpreter kernels"),
                "partition": "one lattice, %d Z-slab(s), cuts %s (one shared corner plane, no collective)" % (world, cuts),
                "l2": "flushed between steps (256 MiB write); working set 27 MB < 126 MB L2",
                "timing": "CUDA events on the launching stream around each step (one CUDA-graph replay of %d kernel nodes chained by programmatic "
                          "dependent launch), summed over the timed steps, max over ranks" % kernels_per_step},
        "triangles_per_sec": ntri * args.steps / (dev_ms * 1e-3),
        "evals_executed_per_sec": evals_exec * args.steps / (dev_ms * 1e-3),
        "stage_ms": mean,
        "sustained": sustained,
        "stage_ms_note": "rank 0, from %globaltimer stamps the kernels of the timed graph replays write themselves (same loop as ms_per_step)",
        "wall_ms_per_step_incl_flush": wall * 1e3 / args.steps,
        "e2e": e2e,
        "gpu_launches": kernels_per_step * args.steps,
        "gpu_launches_e2e_per_step": e2e_launches,
        "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "algorithmic_bytes_per_launch": kbytes, "avg_launch_ms": kms, "peak_source": peak_src,
                     "note": "deep CSG trees are FP32-issue bound, not HBM bound (DESIGN.md); roofline_issue is the view that binds"},
        "roofline_issue": issue,
        "evaluate_e2e": evaluate,
        "knurled1600_zslab": knurled,
        "host_affinity": numa,
        "cpu_baseline": cpu,
        "clocks": clocks,
    }
    emit_json(line)
    if dist is not None:
        dist.destroy_process_group()


def bench_evaluate(sdf, s, glrender, torch):
    """gleval.SDF3.Evaluate(pos, dist) on HOST slices through gsdf_eval3 (pipelined chunks): the reference's batch size
    (32768 points, the evaluation buffer of gsdfaux.go:150) and a 2^24-point batch, pinned and pageable. PCIe roofline:
    12 B/eval travel host->device and 4 B/eval back on a full-duplex link, so the H2D direction bounds it."""
    rng = np.random.default_rng(0)
    mn, mx = s.Bounds()
    out = {}
    big = 1 << 24
    ppos = glrender.pinned_empty((big, 3))
    ppos[:] = (mn + rng.random((big, 3), dtype=np.float32) * (mx - mn)).astype(np.float32)
    pdist = glrender.pinned_empty((big,))
    # live H2D bandwidth of this box (pinned, 192 MiB)
    dbuf = torch.empty(big * 3, dtype=torch.float32, device="cuda")
    hsrc = torch.from_numpy(ppos.reshape(-1))
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        dbuf.copy_(hsrc, non_blocking=True); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    h2d_gbs = big * 12 / best / 1e9
    del dbuf
    for label, n, reps in (("batch_32768", 32768, 200), ("batch_2p24", big, 5)):
        for kind in ("pinned", "pageable"):
            if kind == "pinned":
                pos, dst = ppos[:n], pdist[:n]
            else:
                pos, dst = np.array(ppos[:n]), np.empty(n, np.float32)
                dst[:] = 0   # touch the pages once: first-touch faults are the allocator's cost, not the transfer's
            for _ in range(2):
                sdf.Evaluate(pos, dst)
            t0 = time.perf_counter()
            for _ in range(reps):
                sdf.Evaluate(pos, dst)
            sec = (time.perf_counter() - t0) / reps
            out["%s_%s" % (label, kind)] = {"evals_per_sec": n / sec, "ms_per_call": sec * 1e3, "h2d_GBps": n * 12 / sec / 1e9,
                                           "pcie_roofline_frac": (n * 12 / sec / 1e9) / h2d_gbs}
    out["h2d_peak_GBps_measured"] = h2d_gbs
    out["bytes_per_eval"] = {"h2d": 12, "d2h": 4}
    out["path"] = ("gsdf_eval3 on host memory: chunks of 256 Ki points rotate through three streams (copy in, k_eval_stream, copy out); pageable "
                   "batches of 8 chunks or more are dealt to up to four host threads, each with its own three streams and pinned staging")
    return out


def bench_knurled(rank, world, barrier, allmax, allsum, l2_flush, gleval, glrender, slab, specialise):
    """BASELINE config 4: knurled-cylinder at resdiv 1600 (564x564x1408 corners = 447.9 M), one lattice split into `world`
    Z-slabs. Device time per render, max over ranks."""
    bld, s, res = build_scene("knurled-cylinder", 1600)
    sdf = gleval.NewCUDASDF3(s)
    special = bool(specialise and sdf.Specialize())
    lat = glrender.lattice_from_bounds(*s.Bounds(), res)
    nz = lat.n[2]
    cuts = slab.slab_cuts(nz, world)
    dense = (lat.n[0] + 1) * (lat.n[1] + 1) * (nz + 1)

    def timed(R):
        for _ in range(2):
            l2_flush(); R.Rerun()
        barrier()
        ms = []
        for _ in range(5):
            l2_flush(); R.Rerun(); ms.append(R.Timings()["total_ms"])
        barrier()
        tot = allmax(sum(ms)) / len(ms)
        return {"ms_per_render": tot, "dense_equiv_evals_per_sec": dense / (tot * 1e-3), "evals_executed": int(allsum(R.Evaluations())),
                "triangles": int(allsum(R.NumTriangles())), "plan": R.Plan()}

    R = glrender.NewOctreeRenderer(sdf, res, 1 << 15, cz_range=(cuts[rank], cuts[rank + 1]))
    out = {"scaling": "strong", "lattice_corners": dense, "cuts": cuts, "kernels": "specialised at run time" if special else "interpreter"}
    out.update(timed(R))
    plan = R.Plan()
    R.Close()
    # the same render with the 2-cell level behind the default plan (explicit plans may end with it, include/gsdf_b200.h): the
    # corners of dropped 2-cell cubes are not evaluated; same triangles on this part (asserted)
    F = glrender.Octree(sdf, res, cz_range=(cuts[rank], cuts[rank + 1]), prune=plan + [(2, 1.25)])
    fine = timed(F)
    F.Close()
    assert fine["triangles"] == out["triangles"], "the 2-cell prune level changed the triangle count"
    out["with_2cell_level"] = fine
    return out


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
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-knurled", action="store_true")
    ap.add_argument("--no-specialize", action="store_true", help="keep the interpreter kernels (A/B)")
    ap.add_argument("--no-rebalance", action="store_true", help="keep the equal-layer Z-slab cuts (A/B)")
    ap.add_argument("--numa", action="store_true", help="pin each rank to the CPUs local to its GPU")
    ap.add_argument("--device-only", action="store_true",
                    help="profiling aid: run only the device-resident loop and print its summary, no e2e legs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_cuda(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
