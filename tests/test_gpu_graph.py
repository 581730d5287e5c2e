"""Steady-state reruns replay the render as one CUDA graph: results must be identical to the eager launches, the graph
must follow buffer regrowth / program updates, and stage timings must only be claimed by the eager mode."""
import numpy as np
import pytest

from gsdf_b200 import gsdf, gleval, glrender

pytestmark = pytest.mark.gpu


def test_graph_replay_equals_eager(bld):
    s = gsdf.scene(bld, "npt-flange")
    sdf = gleval.NewCUDASDF3(s)
    res = np.float32(s.Diagonal() / np.float32(120))
    eager = glrender.Octree(sdf, res, stage_timing=True)
    graph = glrender.Octree(sdf, res)
    want = eager.AllTriangles()
    for i in range(4):  # run 0 is eager (allocations), 1 captures, 2.. replay
        graph.Rerun()
        eager.Rerun()
        assert graph.NumTriangles() == len(want)
        assert np.array_equal(graph.AllTriangles().view(np.uint32), want.view(np.uint32)), i
        assert (graph.Evaluations(), graph.TotalPruned()) == (eager.Evaluations(), eager.TotalPruned())
    tg, te = graph.Timings(), eager.Timings()
    # no events inside a graph: its stage times come from the kernels' own %globaltimer stamps and add up to the total
    assert tg["total_ms"] > 0 and tg["eval_ms"] > 0 and tg["emit_ms"] > 0
    assert 0.5 * tg["total_ms"] < tg["prune_ms"] + tg["eval_ms"] + tg["classify_ms"] + tg["emit_ms"] <= 1.05 * tg["total_ms"] + 0.01
    assert te["eval_ms"] > 0 and te["emit_ms"] > 0 and abs(sum(te[k] for k in ("prune_ms", "eval_ms", "classify_ms", "emit_ms")) - te["total_ms"]) < 0.02


def test_graph_follows_program_updates_and_rebinding(bld):
    s1 = gsdf.scene(bld, "npt-flange")
    s2 = bld.Scale(gsdf.scene(bld, "bolt"), 3.0)
    sdf = gleval.NewCUDASDF3(s1)
    res = np.float32(s1.Diagonal() / np.float32(100))
    R = glrender.Octree(sdf, res)                        # replays a graph from its third run on
    E = glrender.Octree(sdf, res, stage_timing=True)     # same lattice (s1's bounds), always eager
    for _ in range(3):
        R.Rerun()
    a = R.AllTriangles()
    assert np.array_equal(a.view(np.uint32), E.AllTriangles().view(np.uint32))
    # another tree uploaded into the same program handle: the program's size changes -> the graph is re-captured
    sdf.Update(s2)
    for _ in range(3):
        R.Rerun()
    E.Rerun()
    b = R.AllTriangles()
    assert len(b) > 0 and len(b) != len(a)
    assert np.array_equal(b.view(np.uint32), E.AllTriangles().view(np.uint32))
    # rebinding the renderer to another program handle, and back
    sdf1 = gleval.NewCUDASDF3(s1)
    R.Rebind(sdf1)
    for _ in range(2):
        R.Rerun()
    assert np.array_equal(R.AllTriangles().view(np.uint32), a.view(np.uint32))
    R.Rebind(sdf)
    R.Rerun()
    assert np.array_equal(R.AllTriangles().view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("scene", ["sphere", "npt-flange", "poly2d"])
def test_streaming_evaluate_tiles_and_tails(oracle, bld, scene):
    """k_eval_stream (bulk-async double-buffered position tiles): full tiles, the partial last tile, sizes below one
    tile (generic kernel), many tiles per CTA, and unaligned device pointers (generic kernel) all agree with the oracle."""
    import torch
    s = {"sphere": lambda: bld.NewSphere(1.0), "npt-flange": lambda: gsdf.scene(bld, "npt-flange"),
         "poly2d": lambda: bld.NewPolygon(np.array([[0, 0], [2, 0], [2.5, 1.5], [1, 1], [0.2, 2]], np.float32))}[scene]()
    d = 2 if s.is2d else 3
    sdf = gleval.NewCUDASDF2(s) if d == 2 else gleval.NewCUDASDF3(s)
    t = oracle.Tree.from_shader(s)
    mn, mx = s.Bounds()
    rng = np.random.default_rng(11)
    nmax = 2048 * 700 + 1234  # more tiles than resident CTAs (296): every CTA loops over >= 2 tiles
    pos_all = (mn - 0.1 * (mx - mn) + rng.random((nmax, d), dtype=np.float32) * 1.2 * (mx - mn)).astype(np.float32)
    want_all = t.eval2(pos_all) if d == 2 else t.eval3(pos_all)
    dpos_all = torch.from_numpy(pos_all).cuda()
    for n in (5, 2047, 2048, 2049, 4096, 6151, 2048 * 300 + 1, nmax):
        dpos = dpos_all[:n].contiguous()
        out = torch.full((n,), 7.0, dtype=torch.float32, device="cuda")
        sdf.Evaluate(dpos, out)
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        assert np.array_equal(got.view(np.uint32), want_all[:n].view(np.uint32)), (scene, n, int((got.view(np.uint32) != want_all[:n].view(np.uint32)).sum()))
    # a misaligned view (offset by one point: 12 or 8 bytes) must take the generic path and still be right
    n = 5000
    dpos = dpos_all[1:n + 1]
    out = torch.empty(n + 1, dtype=torch.float32, device="cuda")[1:]
    sdf.Evaluate(dpos, out)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want_all[1:n + 1].view(np.uint32))
    # host buffers go through the same kernel
    host = np.empty(6151, np.float32)
    sdf.Evaluate(pos_all[:6151], host)
    assert np.array_equal(host.view(np.uint32), want_all[:6151].view(np.uint32))


def test_slab_pipeline_speculative_reads(bld):
    """SlabPipeline.RenderToHost in steady state copies each slab's triangles with the PREVIOUS render's count, without a
    host round trip per slab; the result must equal the single renderer's, and a tree change between two calls (counts
    no longer match) must be detected and repaired."""
    s1 = gsdf.scene(bld, "npt-flange")
    s2 = bld.Difference(s1, bld.Translate(bld.NewSphere(9.0), 20.0, 0.0, -9.0))  # same bounds, different surface
    sdf = gleval.NewCUDASDF3(s1)
    res = np.float32(s1.Diagonal() / np.float32(150))
    single = glrender.Octree(sdf, res)
    want1 = single.AllTriangles()
    for ns in (2, 3, 5):
        P = glrender.SlabPipeline(sdf, res, nslabs=ns)
        host = np.zeros((len(want1) + 4096, 3, 3), np.float32)
        for it in range(4):  # call 0: synchronous; 1..: speculative
            n = P.RenderToHost(host)
            assert n == len(want1) and np.array_equal(host[:n].view(np.uint32), want1.view(np.uint32)), (ns, it)
        sdf.Update(s2)
        single.Rerun()
        want2 = single.AllTriangles()
        assert len(want2) != len(want1)
        for it in range(3):  # first call after the change mis-speculates and re-reads
            n = P.RenderToHost(host)
            assert n == len(want2) and np.array_equal(host[:n].view(np.uint32), want2.view(np.uint32)), (ns, it)
        sdf.Update(s1)
        single.Rerun()
        n = P.RenderToHost(host)
        assert n == len(want1) and np.array_equal(host[:n].view(np.uint32), want1.view(np.uint32))
        # a too small destination is an error, not a silent truncation
        with pytest.raises(Exception):
            P.RenderToHost(np.zeros((100, 3, 3), np.float32))
        # accessors refuse to run while a render is in flight
        from gsdf_b200._lib import lib, check
        import gsdf_b200
        check(lib.gsdf_mesh_rerun_begin(P.parts[0]._h))
        with pytest.raises(gsdf_b200.GsdfError, match="in flight"):
            P.parts[0].NumTriangles()
        check(lib.gsdf_mesh_rerun_end(P.parts[0]._h))
        assert P.parts[0].NumTriangles() > 0
        P.Close()


def test_empty_and_tiny_meshes(oracle, bld):
    """Edge cases: a lattice the surface never enters (zero triangles through every path), and lattices smaller than one
    prune block / one TMA tile."""
    import io
    import gsdf_b200
    # a sphere whose bounds were overloaded to a box far outside it: every distance is positive, no triangle anywhere
    far = bld.OverloadShader3DBounds(bld.NewSphere(1.0), [10, 10, 10], [12, 12.5, 11])
    sdf = gleval.NewCUDASDF3(far)
    t = oracle.Tree.from_shader(far)
    for cls in (glrender.Octree, glrender.FlatRenderer):
        R = cls(sdf, np.float32(0.1))
        for _ in range(3):
            R.Rerun()
        assert R.NumTriangles() == 0 and len(R.AllTriangles()) == 0
        with pytest.raises(glrender.EOF):
            R.ReadTriangles(np.empty((16, 3, 3), np.float32))
        with pytest.raises(gsdf_b200.GsdfError):
            R.STLBytes()  # WriteBinarySTL refuses an empty model (stl.go:16-18)
    P = glrender.SlabPipeline(sdf, np.float32(0.1), nslabs=3)
    host = np.zeros((64, 3, 3), np.float32)
    for _ in range(3):
        assert P.RenderToHost(host) == 0
    d = glrender.DualContourRenderer()
    d.Reset(sdf, np.float32(0.1), glrender.DualContourLeastSquares())
    assert d.Stats()["triangles"] == 0 and len(d.RenderAll(None)) == 0
    d.Rerun()
    assert d.Stats()["cubes"] == 0
    # tiny lattices: 1..3 cells per axis (one partial prune block, a partial TMA tile), reruns through the graph
    s = bld.NewSphere(1.0)
    sdf = gleval.NewCUDASDF3(s)
    mn, mx = s.Bounds()
    for res in (0.7, 1.1, 1.9):
        lat = oracle.flat_lattice(mn, mx, np.float32(res))
        grid, _ = oracle.flat_eval_grid(t := oracle.Tree.from_shader(s), lat)
        want, _ = oracle.flat_march(lat, grid)
        for cls in (glrender.FlatRenderer, glrender.Octree):
            R = cls(sdf, np.float32(res))
            for _ in range(3):
                R.Rerun()
            got = R.AllTriangles()
            if cls is glrender.FlatRenderer:
                assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), res
            else:
                mask, _, _ = oracle.octree_prune_plan(t, lat, R.Plan())
                wp, _ = oracle.flat_march(lat, grid, blockmask=mask)
                assert np.array_equal(got.view(np.uint32), wp.view(np.uint32)), res


def test_programmatic_dependent_launch_is_bit_identical():
    """The kernels of a render are chained by programmatic dependent launch (generators.cuh pdl_trigger / pdl_wait; GSDF_PDL=0
    switches it off). The switch is read once per process, so the check runs in children: graph replays, eager renders
    and the 3-slab pipeline, with and without the programmatic edges, against stage-timed renders, which always use plain
    launches (scripts/check_pdl.py)."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for pdl in ("1", "0"):
        env = dict(os.environ, GSDF_PDL=pdl)
        r = subprocess.run([sys.executable, os.path.join(root, "scripts", "check_pdl.py")], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "PDL CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("knobs", [{"GSDF_BLK_GRID": "1"}, {"GSDF_BLK_GRID": "5"}, {"GSDF_MC": "v1", "GSDF_COUNT_GRID": "3"},
                                   {"GSDF_MC": "v1", "GSDF_COUNT_GRID": "37"}, {"GSDF_MC": "tile5"}, {"GSDF_EVAL_P": "1"}, {"GSDF_RXY": "0"},
                                   {"GSDF_SCAN_FUSED": "0"}, {"GSDF_BLK_GRID": "3", "GSDF_CHILD_NO_CASES": "1"}, {"GSDF_TILESUM": "1"},
                                   {"GSDF_PRUNE_FINE": "1"}, {"GSDF_EVAL_CTA": "384"}, {"GSDF_EVAL_CTA": "64"},
                                   {"GSDF_CHILD_SPECIALIZE": "1"}, {"GSDF_EVAL_P": "4", "GSDF_CHILD_SPECIALIZE": "1"},
                                   {"GSDF_HALF_QUADS": "1", "GSDF_CHILD_SPECIALIZE": "1"}],
                         ids=lambda k: "+".join("%s=%s" % kv for kv in k.items()))
def test_marching_cubes_kernel_variants_match_oracle(knobs):
    """Every marching-cubes kernel family, in a child process with its A/B knob set (the knobs are read once per process):
    the kept-block kernels (default) with the grid capped to 1 and 5 CTAs, so that every warp walks tens of blocks through both
    stencil buffers and mbarrier phases; the one-layer-tile pair with a capped grid (tile loop, stencil reuse, kept and pruned
    tiles in any order); the 4-layer-tile pair; one-corner-per-thread lattice evaluation; programs without radius-reuse flags;
    the stand-alone segment scan (the default runs it inside the emit pass); the block kernels without the parity-mode case store; scan-tile sums
    accumulated by the count pass; the 2-cell prune level in the default plan; interpreter CTAs of 384 and of 64 threads; the run-time compiled kernels with
    two (their default) and with four corners per thread, and with half-quad work lists forced on (large lattices only otherwise).
    Cases and triangles are compared with the oracle, eager, graph capture and graph replay (tests/count_pipeline_child.py)."""
    import os, subprocess, sys
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, **knobs)
    r = subprocess.run([sys.executable, os.path.join(here, "count_pipeline_child.py")], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "COUNT PIPELINE OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
