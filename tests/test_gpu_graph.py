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
    assert tg["total_ms"] > 0 and tg["eval_ms"] == 0 and tg["emit_ms"] == 0     # no events inside a graph
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
