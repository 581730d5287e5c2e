"""The multi-device mesher (gsdf_multi_*), concurrent use of one program from several streams / threads / devices, and the
pipelined host Evaluate. GPU tests run with however many devices the box has (1 is enough for every test but the
explicitly multi-device ones, which then skip); the partition arithmetic is also checked without a GPU."""
import ctypes as C
import os
import threading

import numpy as np
import pytest

import gsdf_b200
from gsdf_b200 import gsdf, gleval, glrender, slab, _lib
import shapes


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


# ---------------------------------------------------------------------------------------------- host arithmetic (no GPU)
def test_slab_cuts_cover_every_layer_once():
    for nz in (1, 3, 4, 5, 17, 84, 85, 526, 1407):
        for n in (1, 2, 3, 4, 8, 16, 24):
            cuts = slab.slab_cuts(nz, n)
            assert cuts[0] == 0 and cuts[-1] == nz and len(cuts) == n + 1
            assert all(b >= a for a, b in zip(cuts[:-1], cuts[1:]))
            if nz // n >= 8:   # thick slabs: interior cuts on 4-layer block boundaries, nobody empty
                assert all(c % 4 == 0 for c in cuts[1:-1]) and all(b > a for a, b in zip(cuts[:-1], cuts[1:]))
            if n <= nz and nz // n < 8:
                assert all(b > a for a, b in zip(cuts[:-1], cuts[1:])), (nz, n, cuts)


def test_rebalanced_cuts_equalise_the_cost():
    """gsdf_slab_rebalance: cuts move towards the expensive slabs, every slab keeps a layer, totals are preserved."""
    cuts = slab.slab_cuts(84, 8)
    cost = [100, 10, 10, 10, 10, 10, 10, 50]
    new = slab.rebalance_cuts(cuts, cost)
    assert new[0] == 0 and new[-1] == 84 and all(b > a for a, b in zip(new[:-1], new[1:]))
    dens = np.repeat(np.array(cost, float) / np.diff(cuts), np.diff(cuts))      # cost per layer under the old cuts
    per = [dens[a:b].sum() for a, b in zip(new[:-1], new[1:])]
    assert max(per) < 0.5 * max(cost) and max(per) / (sum(cost) / 8) < 1.6   # far better than 100 vs a mean of 26
    assert slab.rebalance_cuts([0, 42, 84], [1, 3]) == [0, 56, 84]
    assert slab.rebalance_cuts([0, 1, 2, 3], [0, 0, 5]) == [0, 1, 2, 3]           # nobody loses its last layer
    assert slab.rebalance_cuts(cuts, [0] * 8) == cuts                             # no information: unchanged
    with pytest.raises(gsdf_b200.GsdfError):
        slab.rebalance_cuts([0, 5, 5, 9], [1, 1, 1])


def test_default_prune_plan_arithmetic():
    lat = _lib.Lattice()
    plan = _lib.PrunePlan()
    for n, want in (((280, 280, 84), [3]), ((563, 563, 1407), [5, 3]), ((2200, 2200, 4400), [7, 5, 3])):
        lat.n[0], lat.n[1], lat.n[2] = n
        lat.res = 1.0
        _lib.check(_lib.lib.gsdf_prune_plan_default(C.byref(lat), _lib.MESH_PRUNE, C.byref(plan)))
        assert [l for l, _ in plan.levels()] == want
        assert all(abs(m - 1.25) < 1e-6 for _, m in plan.levels())
        _lib.check(_lib.lib.gsdf_prune_plan_default(C.byref(lat), _lib.MESH_PRUNE | _lib.MESH_PRUNE_LITERAL, C.byref(plan)))
        assert plan.levels()[-1] == (3, 1.0)


def test_multi_begin_rejects_bad_arguments(bld):
    f = bld.flatten(bld.NewSphere(1.0))
    aux = np.zeros(4, np.float32)
    lat = glrender.lattice_from_bounds((-1, -1, -1), (1, 1, 1), 0.1)
    h = C.c_void_p()
    devs = (C.c_int32 * 1)(0)
    args = (f["blob"], len(f["blob"]), aux.ctypes.data_as(C.POINTER(C.c_float)), 0, C.byref(lat), _lib.MESH_PRUNE, C.byref(h))
    assert _lib.lib.gsdf_multi_begin(0, devs, 1, *args) == _lib.EINVAL
    assert _lib.lib.gsdf_multi_begin(1, devs, 0, *args) == _lib.EINVAL
    assert _lib.lib.gsdf_multi_begin(1, None, 1, *args) == _lib.EINVAL
    bad = (C.c_int32 * 1)(99)
    rc = _lib.lib.gsdf_multi_begin(1, bad, 1, *args)
    assert rc in (_lib.EINVAL, _lib.ECUDA)   # without a device: ECUDA (no CPU fallback); with one: device out of range


# ---------------------------------------------------------------------------------------------- multi-device mesher
@pytest.mark.gpu
@pytest.mark.parametrize("scene,resdiv", [("npt-flange", 200), ("knurled-cylinder", 150)])
def test_multi_renderer_equals_single_renderer(bld, scene, resdiv):
    """Any number of Z-slabs over any number of devices concatenates to the single renderer's output bit for bit; pinned and
    pageable destinations; ReadTriangles streaming; STL; program update."""
    s = gsdf.scene(bld, scene)
    res = np.float32(s.Diagonal() / np.float32(resdiv))
    sdf = gleval.NewCUDASDF3(s)
    single = glrender.Octree(sdf, res)
    want = single.AllTriangles()
    wstl = single.STLBytes()
    ndev = gsdf_b200.device_count()
    combos = [([0], 1), ([0], 3), ([0], 7)]
    if ndev >= 2:
        combos += [([0, 1], 1), ([0, 1], 3), (list(range(ndev)), 2)]
    for devs, spd in combos:
        M = glrender.MultiRenderer(s, res, devices=devs, slabs_per_device=spd)
        assert M.NumTriangles() == len(want)
        cuts, sdev = M.Slabs()
        assert cuts[0] == 0 and cuts[-1] == single.lat.n[2] and sdev == [devs[j % len(devs)] for j in range(len(sdev))]
        pinned = glrender.pinned_empty((len(want) + 8, 3, 3))
        pageable = np.zeros((len(want) + 8, 3, 3), np.float32)
        for dst in (pinned, pageable, pinned):
            dst[:] = 0
            n = M.RenderToHost(dst)
            assert n == len(want)
            assert np.array_equal(bits(dst[:n]), bits(want)), (devs, spd)
        assert M.Evaluations() >= single.Evaluations() and M.TotalPruned() > 0 and M.DeviceMs() > 0
        assert np.array_equal(bits(M.AllTriangles()), bits(want))
        assert np.array_equal(bits(glrender.RenderAll(M)), bits(want)) or True   # stream position is at EOF after AllTriangles
        _lib.check(_lib.lib.gsdf_multi_rewind(M._h))
        assert np.array_equal(bits(glrender.RenderAll(M)), bits(want))
        assert M.STLBytes() == wstl
        small = np.zeros((10, 3, 3), np.float32)
        with pytest.raises(gsdf_b200.GsdfError) as e:
            M.RenderToHost(small)
        assert e.value.code == _lib.ESHORT
        if spd > 1 or len(devs) > 1:       # re-cut by executed evaluations: another partition, the same triangles
            before = M.Slabs()[0]
            M.Rebalance(2)
            after = M.Slabs()[0]
            assert after[0] == 0 and after[-1] == before[-1] and len(after) == len(before)
            pinned[:] = 0
            assert M.RenderToHost(pinned) == len(want) and np.array_equal(bits(pinned[:len(want)]), bits(want)), (devs, spd, after)
        M.Close()


@pytest.mark.gpu
def test_multi_renderer_update_uploads_on_the_next_render(oracle, bld):
    a = bld.NewSphere(1.0)
    b = bld.Difference(bld.NewBox(1.6, 1.6, 1.6, 0.1), bld.NewCylinder(0.4, 3, 0))   # fits inside a's bounds
    res = np.float32(0.05)
    devs = list(range(min(2, gsdf_b200.device_count())))
    M = glrender.MultiRenderer(a, res, devices=devs, slabs_per_device=2)
    n_a = M.NumTriangles()
    M.Update(b)
    dst = np.zeros((4 * n_a + 64, 3, 3), np.float32)
    n = M.RenderToHost(dst)
    lat = oracle.flat_lattice(*a.Bounds(), res)
    tb = oracle.Tree.from_shader(b)
    grid, _ = oracle.flat_eval_grid(tb, lat)
    mask, _, _ = oracle.octree_prune_plan(tb, lat, [(3, 1.25)])
    wt, _ = oracle.flat_march(lat, grid, blockmask=mask)
    assert n == len(wt) != n_a and np.array_equal(bits(dst[:n]), bits(wt))
    with pytest.raises(gsdf_b200.GsdfError):
        M.Update(bld.NewCircle(1.0))
    M.Close()


@pytest.mark.gpu
def test_two_devices_in_one_process_from_two_threads(oracle, bld):
    """Per-device function attributes, occupancy and SM counts (a program needing more than 48 KB of dynamic shared memory
    on a second device used to skip the opt-in); the default device is per thread."""
    ndev = gsdf_b200.device_count()
    if ndev < 2:
        pytest.skip("needs two devices")
    s = gsdf.scene(bld, "knurled-cylinder")   # 3 distance + 2 position stack slots: > 48 KB of dynamic shared memory
    f = bld.flatten(s)
    aux = np.ascontiguousarray(f["aux"], np.float32)
    t = oracle.Tree.from_shader(s)
    pos = shapes.sample_points(s, dense=[24, 24, 24])
    want = t.eval3(pos)
    errs = []

    def worker(dev):
        try:
            _lib.check(_lib.lib.gsdf_set_device(dev))
            h = C.c_void_p()
            _lib.check(_lib.lib.gsdf_program_create(f["blob"], len(f["blob"]), aux.ctypes.data_as(C.POINTER(C.c_float)), aux.size, C.byref(h)))
            out = np.empty(len(pos), np.float32)
            for _ in range(3):
                _lib.check(_lib.lib.gsdf_eval3(h, C.c_void_p(pos.ctypes.data), C.c_void_p(out.ctypes.data), len(pos)))
                assert np.array_equal(bits(out), bits(want))
            _lib.lib.gsdf_program_destroy(h)
        except Exception as e:  # noqa: BLE001
            errs.append((dev, repr(e)))

    th = [threading.Thread(target=worker, args=(d,)) for d in range(2)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errs, errs


@pytest.mark.gpu
def test_one_program_on_several_streams_at_once(oracle, bld):
    """Every interpreter launch has its own work-tile scheduler: a pending render, device-tensor Evaluate calls on two
    torch streams (the k_eval fallback: misaligned views) and a host Evaluate in flight together all complete and agree
    with the oracle; gsdf_program_update waits for all of them."""
    torch = pytest.importorskip("torch")
    s = gsdf.scene(bld, "npt-flange")
    sdf = gleval.NewCUDASDF3(s)
    t = oracle.Tree.from_shader(s)
    res = np.float32(s.Diagonal() / np.float32(150))
    R = glrender.Octree(sdf, res)
    want_tris = R.AllTriangles()
    rng = np.random.default_rng(3)
    mn, mx = s.Bounds()
    n = 200001
    pos = (mn + rng.random((n, 3), dtype=np.float32) * (mx - mn)).astype(np.float32)
    want = t.eval3(pos)
    dpos = torch.from_numpy(np.concatenate([np.zeros((1, 3), np.float32), pos])).cuda()
    views = [dpos[1:], dpos[1:].clone()]            # the first one is misaligned: k_eval<GenPoints3> instead of the stream kernel
    outs = [torch.empty(n, dtype=torch.float32, device="cuda") for _ in views]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    evals0 = sdf.Evaluations()
    for rep in range(5):
        _lib.check(_lib.lib.gsdf_mesh_rerun_begin(R._h))
        for v, o, st in zip(views, outs, streams):
            with torch.cuda.stream(st):
                sdf.Evaluate(v, o)
        host_out = np.empty(n, np.float32)
        sdf.Evaluate(pos, host_out)
        _lib.check(_lib.lib.gsdf_mesh_rerun_end(R._h))
        torch.cuda.synchronize()
        assert np.array_equal(bits(host_out), bits(want))
        for o in outs:
            assert np.array_equal(bits(o.cpu().numpy()), bits(want))
        assert np.array_equal(bits(R.AllTriangles()), bits(want_tris))
    assert sdf.Evaluations() - evals0 == 5 * (3 * n + R.Evaluations())   # device, host and renderer evaluations all counted
    # update while launches are in flight on foreign streams: must wait for them, then everything sees the new tree
    for v, o, st in zip(views, outs, streams):
        with torch.cuda.stream(st):
            sdf.Evaluate(v, o)
    sdf.Update(bld.NewSphere(0.7))
    torch.cuda.synchronize()
    for o in outs:
        assert np.array_equal(bits(o.cpu().numpy()), bits(want))
    other = torch.empty(n, dtype=torch.float32, device="cuda:%d" % 0)
    sdf.Evaluate(views[1], other)
    torch.cuda.synchronize()
    assert np.array_equal(bits(other.cpu().numpy()), bits(oracle.Tree.from_shader(bld.NewSphere(0.7)).eval3(pos)))


@pytest.mark.gpu
def test_device_pointer_from_another_gpu_is_rejected(bld):
    torch = pytest.importorskip("torch")
    if gsdf_b200.device_count() < 2:
        pytest.skip("needs two devices")
    sdf = gleval.NewCUDASDF3(bld.NewSphere(1.0))      # lives on device 0
    pos = torch.zeros((64, 3), dtype=torch.float32, device="cuda:1")
    out = torch.zeros(64, dtype=torch.float32, device="cuda:1")
    with pytest.raises(gsdf_b200.GsdfError) as e:
        sdf.Evaluate(pos, out)
    assert e.value.code == _lib.EINVAL


# ---------------------------------------------------------------------------------------------- pipelined host Evaluate
@pytest.mark.gpu
@pytest.mark.parametrize("chunk,lanes", [(0, 0), (2048, 0), (6144, 0), (2048, 3), (4096, 4)])
def test_host_evaluate_is_chunked_and_exact(oracle, bld, monkeypatch, chunk, lanes):
    """gsdf_eval3 / gsdf_eval2 on host slices: three rotating chunks in flight per lane, up to four lanes (host threads) for
    pageable memory; ragged sizes, pinned and pageable buffers, sizes below / at / above the chunk size all equal the oracle."""
    if chunk:
        monkeypatch.setenv("GSDF_EVAL_CHUNK", str(chunk))
    if lanes:
        monkeypatch.setenv("GSDF_EVAL_LANES", str(lanes))
    import importlib
    rng = np.random.default_rng(11)
    s3 = gsdf.scene(bld, "bolt")
    s2 = bld.Union2D(bld.NewCircle(0.5), bld.Translate2D(bld.NewRectangle(1.0, 0.4), 0.8, 0.3))
    for s, dim in ((s3, 3), (s2, 2)):
        sdf = (gleval.NewCUDASDF3 if dim == 3 else gleval.NewCUDASDF2)(s)
        t = oracle.Tree.from_shader(s)
        mn, mx = s.Bounds()
        for n in (1, 5, 2047, 2048, 2049, 4096, 12288, 12289, 50001, 300007):
            pos = (mn + rng.random((n, dim), dtype=np.float32) * (mx - mn)).astype(np.float32)
            want = (t.eval3 if dim == 3 else t.eval2)(pos)
            out = np.empty(n, np.float32)
            sdf.Evaluate(pos, out)
            assert np.array_equal(bits(out), bits(want)), (dim, n, "pageable")
            ppos = glrender.pinned_empty((n, dim)); ppos[:] = pos
            pout = glrender.pinned_empty((n,)); pout[:] = 0
            sdf.Evaluate(ppos, pout)
            assert np.array_equal(bits(pout), bits(want)), (dim, n, "pinned")
            sdf.Evaluate(ppos, out)
            assert np.array_equal(bits(out), bits(want)), (dim, n, "pinned in, pageable out")


@pytest.mark.gpu
def test_device_driven_read_back_matches_the_single_renderer():
    """GSDF_MULTI_COPYK=1: the multi-slab driver's read-back as kernels enqueued up front that read the slabs' triangle
    counts themselves (mesher.cu k_copy_out) instead of copies the host enqueues once it has seen a count. The switch is read
    once per process, so the check runs in a child: 1-7 slabs, first and steady-state renders, re-cut slabs, a pageable
    destination (which keeps the classic path) -- the destination always equals the single renderer's triangles bit for bit
    and nothing is written behind them (tests/multi_copyk_child.py)."""
    import os, subprocess, sys
    here = os.path.dirname(os.path.abspath(__file__))
    for mode in ("1", "0"):
        env = dict(os.environ, GSDF_MULTI_COPYK=mode)
        r = subprocess.run([sys.executable, os.path.join(here, "multi_copyk_child.py")], env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0 and "MULTI COPYK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
