"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes), against the CPU oracle on the same
inputs. Distances, cube-case indices, triangles and STL bytes must be BIT-IDENTICAL: both sides evaluate the same
float32 operation sequences with individually rounded operations (-fmad=false vs -ffp-contract=off).
north_star's tolerance for distances is 1e-5 relative to cpu_evaluators.go; bit equality with the oracle is stricter.
"""
import ctypes as C
import hashlib
import io
import os
import struct

import numpy as np
import pytest

import shapes
import gsdf_b200
from gsdf_b200 import gsdf, gleval, glrender, _lib

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REL_TOL = 1e-5  # north_star: "within 1e-5 relative float tolerance"


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def gpu_eval(s, pos):
    sdf = gleval.NewCUDASDF2(s) if s.is2d else gleval.NewCUDASDF3(s)
    out = np.empty(len(pos), np.float32)
    sdf.Evaluate(np.ascontiguousarray(pos, np.float32), out)
    return out, sdf


def check_field(name, s, oracle, pos):
    t = oracle.Tree.from_shader(s)
    want = t.eval2(pos) if s.is2d else t.eval3(pos)
    got, sdf = gpu_eval(s, pos)
    rel = np.abs(got.astype(np.float64) - want) / np.maximum(1.0, np.abs(want))
    assert rel.max() <= REL_TOL, (name, float(rel.max()))
    nbad = int((bits(got) != bits(want)).sum())
    assert nbad == 0, "%s: %d of %d distances differ in their bits (max rel %.3g)" % (name, nbad, len(pos), rel.max())
    assert sdf.Evaluations() == len(pos)


# ---------------------------------------------------------------------------------------------- Evaluate parity
@pytest.mark.parametrize("corpus", ["primitives3d", "binops3d", "unary3d", "threads3d", "scenes3d", "guards3d", "dag3d",
                                    "primitives2d", "binops2d", "unary2d", "threads2d"])
def test_evaluate_matches_oracle_on_reference_lattices(oracle, bld, corpus):
    """testShader3D/testShader2D (gsdf_test.go:429-525): sample the AppendGrid lattice of Bounds()."""
    for name, s in getattr(shapes, corpus)(bld):
        check_field(name, s, oracle, shapes.sample_points(s))


def test_nan_propagation_of_min_max_matches_go_on_gpu(oracle, bld):
    """math32.Min / Max propagate NaN (gsdf.go:141-143 on Go's math semantics); the kernels use min.NaN.f32 / max.NaN.f32.
    Far-field positions (ellipse2D turns NaN at |p| ~ 1e6 in the Go formula too) and NaN coordinates through every node
    type and 60 random trees: NaN exactly where the oracle has NaN, bit-equal elsewhere."""
    from test_host_interp import far_and_nan_points
    todo = [(n, sh) for c in ("all3d", "all2d") for n, sh in getattr(shapes, c)(bld)]
    todo += [(n, sh) for dim in (3, 2) for n, sh in shapes.random_trees(bld, 5, 30, dim)]
    nans = 0
    for name, s in todo:
        pos = far_and_nan_points(s)
        t = oracle.Tree.from_shader(s)
        want = t.eval2(pos) if s.is2d else t.eval3(pos)
        got, _ = gpu_eval(s, pos)
        diff = (bits(got) != bits(want)) & ~(np.isnan(got) & np.isnan(want))
        assert not diff.any(), (name, int(diff.sum()))
        nans += int(np.isnan(want).sum())
    assert nans > 100   # the inputs do exercise the NaN paths


def test_array_folds_start_from_the_reference_seed_on_gpu(oracle, bld):
    """largenum (1e20) / math.MaxFloat32 as the first operand of the array folds (cpu_evaluators.go:364,932,1172): GSDF_OP_MIN_CONST."""
    from test_host_interp import array_far_cases
    for name, s, pos in array_far_cases(bld):
        t = oracle.Tree.from_shader(s)
        want = t.eval2(pos) if s.is2d else t.eval3(pos)
        got, _ = gpu_eval(s, pos)
        diff = (bits(got) != bits(want)) & ~(np.isnan(got) & np.isnan(want))
        assert not diff.any(), (name, got, want)


@pytest.mark.parametrize("dim", [3, 2])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_trees_bit_identical(oracle, bld, seed, dim):
    """Seeded random compositions of every constructor (testRandomUnary3D/2D taken further, gsdf_test.go:233-283):
    nesting exercises the flattener's stack allocation, position liveness and guard placement on trees nobody wrote
    by hand. Points past the end of the reference lattice are included through the margin."""
    for name, s in shapes.random_trees(bld, seed, 40, dim):
        check_field(name, s, oracle, shapes.sample_points(s))


def test_box_guards_are_sound_for_overlapping_operands(oracle, bld):
    """Unions / differences of OVERLAPPING bounded 2-D shapes (tests/shapes.py::overlap2d): a box guard must not fire for
    a tile that has points inside the operand's box. The point sets put whole 2048-point tiles where the first version of
    the guard skipped a live operand (found on the CPU model, tests/test_progsim.py)."""
    for name, s in shapes.overlap2d(bld):
        check_field(name, s, oracle, shapes.overlap2d_points(name, s))


def test_evaluate_golden_fixtures(bld):
    g = np.load(os.path.join(GOLD, "distances.npz"))
    for name, s in shapes.all3d(bld) + shapes.all2d(bld):
        got, _ = gpu_eval(s, g[name + ".pos"])
        assert np.array_equal(bits(got), bits(g[name + ".dist"])), name


def test_sphere_64cubed_plumbing(oracle, bld):
    """BASELINE config 0: single sphere, dense 64^3 lattice over Bounds()."""
    s = bld.NewSphere(1)
    pos = shapes.sample_points(s, margin=0, dense=[64, 64, 64])
    assert len(pos) == 262144
    check_field("sphere64", s, oracle, pos)
    got, _ = gpu_eval(s, pos)
    assert np.abs(got - (np.linalg.norm(pos.astype(np.float64), axis=1) - 1)).max() < 1e-6


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 31, 32, 33, 255, 256, 257, 1023, 4097])
def test_evaluate_ragged_sizes(oracle, bld, n):
    s = gsdf.scene(bld, "npt-flange")
    rng = np.random.default_rng(n)
    mn, mx = s.Bounds()
    pos = (mn + rng.random((n, 3), dtype=np.float32) * (mx - mn)).astype(np.float32)
    check_field("flange_n%d" % n, s, oracle, pos)


def test_evaluate_unaligned_views(oracle, bld):
    """Host slices that are not 16-byte aligned must still work (Go passes &pos[k])."""
    s = bld.NewBox(1, 0.6, 0.8, 0.1)
    rng = np.random.default_rng(5)
    base = rng.uniform(-1, 1, (1001, 3)).astype(np.float32)
    pos = base[1:]            # +12 bytes
    dist_base = np.empty(1003, np.float32)
    dist = dist_base[3:]      # +12 bytes
    sdf = gleval.NewCUDASDF3(s)
    sdf.Evaluate(pos, dist)
    want = oracle.Tree.from_shader(s).eval3(pos)
    assert np.array_equal(bits(dist), bits(want))


def test_evaluate_errors(bld):
    """gleval/cpu.go:93-100: errMismatchBufferLength / errEmptyBuffers; wrong dimension is rejected."""
    sdf = gleval.NewCUDASDF3(bld.NewSphere(1))
    with pytest.raises(gsdf_b200.GsdfError) as e:
        sdf.Evaluate(np.zeros((4, 3), np.float32), np.zeros(3, np.float32))
    assert e.value.code == _lib.ELEN
    with pytest.raises(gsdf_b200.GsdfError) as e:
        sdf.Evaluate(np.zeros((0, 3), np.float32), np.zeros(0, np.float32))
    assert e.value.code == _lib.EEMPTY
    assert sdf.Evaluations() == 0  # evals only advance on success (cpu.go:116)
    with pytest.raises(gsdf_b200.GsdfError):
        gleval.NewCUDASDF2(bld.NewSphere(1))
    with pytest.raises(gsdf_b200.GsdfError):
        gleval.NewCUDASDF3(bld.NewCircle(1))
    rc = _lib.lib.gsdf_eval3(None, None, None, 0)
    assert rc == _lib.EINVAL and "NULL" in _lib.last_error()


def test_program_create_rejects_malformed_blobs(bld):
    f = bld.flatten(bld.NewSphere(1))
    h = C.c_void_p()
    blob = bytearray(f["blob"])
    blob[0] ^= 0xff  # magic
    assert _lib.lib.gsdf_program_create(bytes(blob), len(blob), None, 0, C.byref(h)) == _lib.EPROGRAM
    blob = bytearray(f["blob"])
    blob[32] = 250   # unknown opcode
    assert _lib.lib.gsdf_program_create(bytes(blob), len(blob), None, 0, C.byref(h)) == _lib.EPROGRAM
    blob = bytearray(f["blob"])[:-16]  # END chopped off
    assert _lib.lib.gsdf_program_create(bytes(blob), len(blob), None, 0, C.byref(h)) == _lib.EPROGRAM


def test_slab_guards_do_not_change_a_single_bit(oracle, bld, monkeypatch):
    """Dense x-fastest lattices, so whole CTA tiles lie outside the guarded slab and really take the skip: the guarded
    program, the unguarded program (GSDF_NO_GUARDS=1 at flatten time) and the oracle agree bit for bit."""
    for name, s in shapes.guards3d(bld) + [("npt-flange", gsdf.scene(bld, "npt-flange")), ("bolt", gsdf.scene(bld, "bolt"))]:
        pos = shapes.sample_points(s, margin=0.3, dense=(96, 40, 48))
        guarded, _ = gpu_eval(s, pos)
        with monkeypatch.context() as mp:
            mp.setenv("GSDF_NO_GUARDS", "1")
            plain, _ = gpu_eval(s, pos)
        assert np.array_equal(bits(guarded), bits(plain)), name
        t = oracle.Tree.from_shader(s)
        assert np.array_equal(bits(guarded), bits(t.eval3(pos))), name


def test_program_create_rejects_bad_guard_targets(bld):
    s = shapes.guards3d(bld)[0][1]
    f = bld.flatten(s)
    words = np.frombuffer(f["blob"], np.uint32, offset=32).reshape(-1, 4)
    at = [i for i in range(len(words)) if (int(words[i, 1]) & 0xff) == 1 and (int(words[i, 1]) >> 8) > i][0]
    h = C.c_void_p()
    aux = np.ascontiguousarray(f["aux"], np.float32)
    auxp = aux.ctypes.data_as(C.POINTER(C.c_float))
    for bad in (at, 1 << 20, at + 1):  # not forward, out of range, the very next chunk
        blob = bytearray(f["blob"])
        struct.pack_into("<I", blob, 32 + 16 * at + 4, 1 | (bad << 8))
        assert _lib.lib.gsdf_program_create(bytes(blob), len(blob), auxp, aux.size, C.byref(h)) == _lib.EPROGRAM, bad
    blob = bytearray(f["blob"])
    struct.pack_into("<I", blob, 32 + 16 * at + 4, 9 | (int(words[at, 1]) >> 8 << 8))  # unknown guard kind
    assert _lib.lib.gsdf_program_create(bytes(blob), len(blob), auxp, aux.size, C.byref(h)) == _lib.EPROGRAM


def test_evaluate_device_pointers(oracle, bld):
    """gsdf_eval3_device on buffers already resident in HBM (torch only provides the device memory)."""
    import torch
    s = gsdf.scene(bld, "bolt")
    pos = shapes.sample_points(s, dense=[40, 40, 40])
    dpos = torch.from_numpy(pos).cuda()
    ddist = torch.empty(len(pos), dtype=torch.float32, device="cuda")
    sdf = gleval.NewCUDASDF3(s)
    sdf.Evaluate(dpos, ddist)
    torch.cuda.synchronize()
    want = oracle.Tree.from_shader(s).eval3(pos)
    assert np.array_equal(bits(ddist.cpu().numpy()), bits(want))


def test_large_polygon_side_buffer_from_global(oracle, bld):
    """A polygon too large to stage in shared memory is read from global memory; results are unchanged."""
    rng = np.random.default_rng(9)
    ang = np.sort(rng.uniform(0, 2 * np.pi, 6000))
    r = 1 + 0.05 * rng.standard_normal(6000)
    verts = np.stack([r * np.cos(ang), r * np.sin(ang)], 1).astype(np.float32)
    s = bld.NewPolygon(verts)
    pos = shapes.sample_points(s, dense=[48, 48])
    check_field("bigpoly", s, oracle, pos)


# ---------------------------------------------------------------------------------------------- dense lattice
def test_grid_eval_matches_flatrenderer_lattice(oracle, bld):
    """FlatRenderer.evalGrid (flatrenderer.go:103-182): same positions, same order, same bits; k-slabs concatenate."""
    s = gsdf.scene(bld, "npt-flange")
    sdf = gleval.NewCUDASDF3(s)
    res = np.float32(s.Diagonal() / np.float32(90))
    lat = glrender.lattice_from_bounds(*s.Bounds(), res)
    olat = oracle.flat_lattice(*s.Bounds(), res)
    assert list(lat.n) == list(olat.n) and list(lat.origin) == list(olat.origin) and lat.res == olat.res
    want, ev = oracle.flat_eval_grid(oracle.Tree.from_shader(s), olat, nthreads=4)
    nx, ny, nz = lat.n
    got = np.empty((nz + 1, ny + 1, nx + 1), np.float32)
    _lib.check(_lib.lib.gsdf_grid_eval(sdf._h, C.byref(lat), 0, nz + 1, C.c_void_p(got.ctypes.data)))
    assert np.array_equal(bits(got), bits(want))
    assert sdf.Evaluations() == ev
    # split like evalGrid's goroutines: g*(nz+1)/G .. (g+1)*(nz+1)/G
    G = 3
    parts = []
    for g in range(G):
        k0, k1 = g * (nz + 1) // G, (g + 1) * (nz + 1) // G
        part = np.empty((k1 - k0, ny + 1, nx + 1), np.float32)
        _lib.check(_lib.lib.gsdf_grid_eval(sdf._h, C.byref(lat), k0, k1, C.c_void_p(part.ctypes.data)))
        parts.append(part)
    assert np.array_equal(bits(np.concatenate(parts)), bits(want))


# ---------------------------------------------------------------------------------------------- mesher
def oracle_mesh(oracle, s, res, plan):
    """plan: [] = FlatRenderer, else the renderer's prune plan [(level, margin), ...] (glrender.Octree.Plan())."""
    t = oracle.Tree.from_shader(s)
    lat = oracle.flat_lattice(*s.Bounds(), res)
    grid, ev = oracle.flat_eval_grid(t, lat, nthreads=os.cpu_count() or 1)
    mask, centres = None, 0
    if plan:
        mask, _, centres = oracle.octree_prune_plan(t, lat, plan)
    tris, cases = oracle.flat_march(lat, grid, want_cases=True, blockmask=mask)
    return lat, grid, mask, tris, cases, centres


PLANS = {"flat": False, "default": None, "literal": "literal", "level-3": [(3, 1.25)], "two-level": [(5, 1.25), (3, 1.25)],
         "three-level-literal": [(6, 1.0), (4, 1.0), (3, 1.0)], "fine": [(3, 1.25), (2, 1.25)], "three-level-fine": [(5, 1.25), (3, 1.25), (2, 1.5)]}


def test_sphere_41072_on_gpu(oracle, bld):
    """TestSphereMarchingTriangles (glrender_test.go:83-102) through the GPU Octree renderer + STL round trip."""
    sdf = gleval.NewCUDASDF3(bld.NewSphere(1.0))
    r = glrender.NewOctreeRenderer(sdf, np.float32(1.0 / 33), (1 << 12) + 1)
    tris = glrender.RenderAll(r)
    assert len(tris) == 41072
    buf = io.BytesIO()
    n = glrender.WriteBinarySTL(buf, tris)
    assert n == buf.getbuffer().nbytes == 84 + 50 * 41072
    buf.seek(0)
    back = glrender.ReadBinarySTL(buf)
    assert np.array_equal(bits(back), bits(tris))        # glrender_test.go:149-153
    assert r.TotalPruned() > 0 and r.Evaluations() > 0
    f = glrender.NewFlatRenderer(sdf, np.float32(1.0 / 33), 4096, 1)
    assert np.array_equal(bits(glrender.RenderAll(f)), bits(tris))


@pytest.mark.parametrize("prune", list(PLANS))
@pytest.mark.parametrize("scene,resdiv", [("sphere", 70), ("npt-flange", 150), ("bolt", 160), ("knurled-cylinder", 170)])
def test_mesh_bit_identical_to_oracle(oracle, bld, scene, resdiv, prune):
    """Every renderer mode against the oracle's restatement: the dense sweep, the default coarse-to-fine prune (level 3 with
    margin 1.25), the reference's literal rule at level 3, and explicit multi-level plans (parent / child bookkeeping)."""
    s = bld.NewSphere(1.0) if scene == "sphere" else gsdf.scene(bld, scene)
    res = np.float32(s.Diagonal() / np.float32(resdiv))
    sdf = gleval.NewCUDASDF3(s)
    if PLANS[prune] is False:
        R = glrender.FlatRenderer(sdf, res, keep_cases=True, keep_grid=True)
    else:
        R = glrender.Octree(sdf, res, keep_cases=True, keep_grid=True, prune=PLANS[prune])
    prune = PLANS[prune] is not False
    if prune:
        assert R.Plan()[-1][0] in (2, 3)   # explicit plans may end with the 2-cell level
    lat, grid, mask, wt, wc, centres = oracle_mesh(oracle, s, res, R.Plan())
    assert list(R.lat.n) == list(lat.n)
    cases = R.Cases()
    assert int((cases != wc).sum()) == 0                       # cube-case indices bit-identical
    tris = R.AllTriangles()
    assert len(tris) == len(wt) == R.NumTriangles()
    assert np.array_equal(bits(tris), bits(wt))                # same triangles, same order, same bits
    assert R.STLBytes() == oracle.stl_write(wt)
    if not prune:
        assert np.array_equal(bits(R.Grid()), bits(grid))
        assert R.Evaluations() == grid.size and R.TotalPruned() == 0
    else:
        kept = int(mask.sum())
        assert R.TotalPruned() == (mask.size - kept) * (1 << (R.Plan()[-1][0] - 1)) ** 3   # Cube.DecomposesTo(1) of the finest cubes
        g = R.Grid()                                           # evaluated corners agree; pruned ones hold the fill value
        ev = bits(g) != np.uint32(0x7f7f7f7f)
        assert np.array_equal(bits(g)[ev], bits(grid)[ev]) and ev.sum() < grid.size
        nq = (lat.n[0] + 1 + 3) // 4 * 4                       # evaluations = cube centres of every level + listed lattice quads
        assert (R.Evaluations() - centres) % 4 == 0 and centres < R.Evaluations() <= centres + nq * (lat.n[1] + 1) * (lat.n[2] + 1)
    R.Rerun()                                                  # Reset/re-render reuses buffers and reproduces the result
    assert np.array_equal(bits(R.AllTriangles()), bits(wt))


@pytest.mark.parametrize("prune", ["flat", "default", "literal", "two-level"])
def test_random_trees_mesh_bit_identical(oracle, bld, prune):
    """The mesher on seeded random trees (tests/shapes.py): lattice, cube-case indices, triangles and their order are
    the oracle's for shapes with thin shells, arrays and non-Lipschitz fields, with and without the octree prune."""
    for name, s in shapes.random_trees(bld, 11, 10, 3, depth=3):
        res = np.float32(s.Diagonal() / np.float32(60))
        sdf = gleval.NewCUDASDF3(s)
        R = glrender.FlatRenderer(sdf, res, keep_cases=True) if PLANS[prune] is False else glrender.Octree(sdf, res, keep_cases=True, prune=PLANS[prune])
        lat, grid, mask, wt, wc, _ = oracle_mesh(oracle, s, res, R.Plan())
        assert list(R.lat.n) == list(lat.n), name
        assert int((R.Cases() != wc).sum()) == 0, name
        tris = R.AllTriangles()
        assert len(tris) == len(wt) == R.NumTriangles(), name
        assert np.array_equal(bits(tris), bits(wt)), name
        R.Close()


def test_flange_resdiv400_readme_counts(bld):
    """README.md:116,130: 423,852 triangles from both renderers; FlatRenderer evaluates 6,711,685 lattice corners."""
    s = gsdf.scene(bld, "npt-flange")
    sdf = gleval.NewCUDASDF3(s)
    res = np.float32(s.Diagonal() / np.float32(400))
    f = glrender.NewFlatRenderer(sdf, res)
    assert list(f.lat.n) == [280, 280, 84]
    assert f.NumTriangles() == 423852 and f.Evaluations() == 6711685
    o = glrender.NewOctreeRenderer(sdf, res, 32768)
    assert o.NumTriangles() == 423852 and o.Evaluations() < f.Evaluations()
    g = np.load(os.path.join(GOLD, "meshes.npz"))
    assert np.array_equal(bits(f.AllTriangles()), bits(o.AllTriangles()))


def test_showerhead_resdiv350_readme_counts(oracle, bld):
    """README.md:152,165: fibonacci-showerhead at resdiv 350 -> 309,872 triangles from both renderers, 1,512,024 lattice
    corners; a 675-instruction program (131-operand union). Bit-identical to the oracle."""
    s = gsdf.scene(bld, "fibonacci-showerhead")
    sdf = gleval.NewCUDASDF3(s)
    res = np.float32(s.Diagonal() / np.float32(350))
    f = glrender.NewFlatRenderer(sdf, res)
    assert f.NumTriangles() == 309872 and f.Evaluations() == 1512024
    o = glrender.NewOctreeRenderer(sdf, res, 32768)
    assert o.NumTriangles() == 309872  # README.md:152: the reference's octree run; default plan (level 3, margin 1.25)
    assert o.Evaluations() < f.Evaluations()
    t = oracle.Tree.from_shader(s)
    lat = oracle.flat_lattice(*s.Bounds(), res)
    grid, _ = oracle.flat_eval_grid(t, lat, nthreads=os.cpu_count() or 1)
    want, _ = oracle.flat_march(lat, grid)
    assert np.array_equal(bits(f.AllTriangles()), bits(want))
    assert np.array_equal(bits(o.AllTriangles()), bits(want))   # the pruned render equals the dense sweep bit for bit
    # the reference's rule applied literally to EVERY level-3 cube drops 23 triangles on this field (smooth blends and
    # knurls are not 1-Lipschitz); the reference's scheduler never gets to most of those cubes
    lit = glrender.NewOctreeRenderer(sdf, res, 32768, prune="literal")
    assert lit.NumTriangles() == 309849
    mask, _ = oracle.octree_prune_mask(t, lat)
    wantp, _ = oracle.flat_march(lat, grid, blockmask=mask)
    assert np.array_equal(bits(lit.AllTriangles()), bits(wantp))


def test_golden_mesh_fixtures(bld):
    g = np.load(os.path.join(GOLD, "meshes.npz"))
    for name in ["sphere", "npt-flange", "bolt", "knurled-cylinder"]:
        s = bld.NewSphere(1.0) if name == "sphere" else gsdf.scene(bld, name)
        sdf = gleval.NewCUDASDF3(s)
        R = glrender.FlatRenderer(sdf, np.float32(g[name + ".res"]), keep_cases=True)
        assert list(R.lat.n) == list(g[name + ".n"])
        tris = R.AllTriangles()
        assert len(tris) == int(g[name + ".ntri"])
        assert hashlib.sha256(tris.tobytes()).digest() == g[name + ".tri_sha"].tobytes()
        assert hashlib.sha256(R.Cases().tobytes()).digest() == g[name + ".case_sha"].tobytes()
        assert hashlib.sha256(R.STLBytes()).digest() == g[name + ".stl_sha"].tobytes()
        P = glrender.Octree(sdf, np.float32(g[name + ".res"]), prune="literal")
        assert P.NumTriangles() == int(g[name + ".ntri_pruned"])
        assert glrender.Octree(sdf, np.float32(g[name + ".res"])).NumTriangles() == int(g[name + ".ntri"])  # default plan == dense sweep


def test_read_triangles_streaming_contract(bld):
    """Renderer.ReadTriangles: any len(dst) >= 5 works and resumes; < 5 is io.ErrShortBuffer; then io.EOF."""
    sdf = gleval.NewCUDASDF3(bld.NewSphere(1.0))
    r = glrender.NewOctreeRenderer(sdf, np.float32(0.11), 64)
    allt = r.AllTriangles()
    r.Rerun()
    with pytest.raises(glrender.ErrShortBuffer):
        r.ReadTriangles(np.empty((4, 3, 3), np.float32))
    got = []
    for size in [5, 7, 64, 5, 1000, 33] * 1000:
        buf = np.empty((size, 3, 3), np.float32)
        try:
            n = r.ReadTriangles(buf)
        except glrender.EOF:
            break
        assert 0 < n <= size
        got.append(buf[:n].copy())
    assert np.array_equal(bits(np.concatenate(got)), bits(allt))
    with pytest.raises(glrender.EOF):
        r.ReadTriangles(np.empty((5, 3, 3), np.float32))


def test_renderer_errors(bld):
    sdf = gleval.NewCUDASDF3(bld.NewSphere(1.0))
    with pytest.raises(gsdf_b200.GsdfError):
        glrender.NewOctreeRenderer(sdf, 0.1, 63)        # octreerenderer.go:46
    with pytest.raises(gsdf_b200.GsdfError):
        glrender.NewOctreeRenderer(sdf, -1.0, 64)       # :73
    with pytest.raises(gsdf_b200.GsdfError):
        glrender.NewFlatRenderer(sdf, 0.1, 4)           # flatrenderer.go:41
    with pytest.raises(gsdf_b200.GsdfError) as e:
        glrender.WriteBinarySTL(io.BytesIO(), np.zeros((0, 3, 3), np.float32))  # stl.go:16
    assert e.value.code == _lib.EEMPTY
    # a renderer whose surface produces no triangles still terminates with EOF
    far = gleval.NewCUDASDF3(bld.Offset(bld.NewSphere(1.0), 10.0))
    r = glrender.NewFlatRenderer(far, np.float32(0.5))
    assert r.NumTriangles() == 0 and len(glrender.RenderAll(r)) == 0


def test_octree_awkward_resolutions_gpu(oracle, bld):
    """TestOctree (glrender_test.go:104-124)."""
    s = bld.NewSphere(1.0)
    sdf = gleval.NewCUDASDF3(s)
    r = glrender.NewOctreeRenderer(sdf, np.float32(1 / 32), 1 << 12)
    for res in [1 / 4, 1 / 8, 1 / 37, 1 / 4.000001, 1 / 13, 1 / 3.5]:
        r.Reset(sdf, np.float32(res))
        _, _, _, wt, _, _ = oracle_mesh(oracle, s, np.float32(res), r.Plan())
        tris = glrender.RenderAll(r)
        assert len(tris) > 0 and np.array_equal(bits(tris), bits(wt))


def test_z_slabs_concatenate_to_the_whole(bld):
    """Multi-GPU partition: cells split by Z-slab, one shared corner plane, per-slab triangle buffers
    concatenated in slab order reproduce the single-device output bit for bit, for aligned and unaligned cuts."""
    s = gsdf.scene(bld, "npt-flange")
    sdf = gleval.NewCUDASDF3(s)
    res = np.float32(s.Diagonal() / np.float32(200))
    whole = glrender.Octree(sdf, res, keep_cases=True)
    nz = whole.lat.n[2]
    wt, wc = whole.AllTriangles(), whole.Cases()
    for cuts in ([0, 8, 20, nz], [0, 5, 6, 23, nz], [0, 1, nz]):
        parts, cparts, ev = [], [], 0
        for a, b in zip(cuts[:-1], cuts[1:]):
            r = glrender.Octree(sdf, res, cz_range=(a, b), keep_cases=True)
            parts.append(r.AllTriangles())
            cparts.append(r.Cases())
            ev += r.Evaluations()
        assert np.array_equal(bits(np.concatenate(parts)), bits(wt)), cuts
        assert np.array_equal(np.concatenate(cparts), wc), cuts
    # the same with a plan that ends with the 2-cell level (levels 3 + 2 in one launch, child masks in the marching-cubes
    # kernels): slabs whose cuts split 4-cell blocks and 2-cell cubes still concatenate to the whole, which equals the default
    fine = [(3, 1.25), (2, 1.25)]
    wf = glrender.Octree(sdf, res, prune=fine)
    assert np.array_equal(bits(wf.AllTriangles()), bits(wt)) and wf.Evaluations() < whole.Evaluations()
    for cuts in ([0, 8, 20, nz], [0, 5, 6, 23, nz], [0, 1, nz]):
        parts = [glrender.Octree(sdf, res, cz_range=(a, b), prune=fine).AllTriangles() for a, b in zip(cuts[:-1], cuts[1:])]
        assert np.array_equal(bits(np.concatenate(parts)), bits(wt)), cuts


def test_image_eval_matches_oracle(oracle, bld):
    """ImageRendererSDF2 evaluation (image.go:76-105) on a polygon-heavy 2D tree (config 5's evaluator path)."""
    T = gsdf.threads
    glyphs = [bld.Translate2D(bld.NewPolygon(shapes.nagon(n, 0.4)), 1.1 * i, 0) for i, n in enumerate([3, 5, 8, 13, 21])]
    hole = bld.Translate2D(bld.NewCircle(0.15), 2.2, 0)
    s = bld.Difference2D(bld.Union2D(*glyphs), hole)
    sdf = gleval.NewCUDASDF2(s)
    for w, h in [(640, 123), (257, 64), (1024, 32)]:
        got = glrender.ImageEvaluateSDF2(sdf, w, h)
        want = oracle.Tree.from_shader(s).image_eval2(*s.Bounds(), w, h)
        assert np.array_equal(bits(got), bits(want)), (w, h)


# ---------------------------------------------------------------------------------------------- full-size properties
def test_flange_full_size_properties(bld):
    """BASELINE config 1 at full size, size-independent properties: prune is lossless, every vertex lies on a cell
    edge of the lattice, STL pack round-trips, re-running is idempotent."""
    s = gsdf.scene(bld, "npt-flange")
    sdf = gleval.NewCUDASDF3(s)
    res = np.float32(s.Diagonal() / np.float32(400))
    o = glrender.Octree(sdf, res)
    tris = o.AllTriangles()
    assert len(tris) == 423852
    lat = o.lat
    org = np.array(list(lat.origin), np.float64)
    rel = (tris.reshape(-1, 3).astype(np.float64) - org) / float(lat.res)
    frac = np.abs(rel - np.round(rel))
    on_lattice = (frac < 1e-3).sum(axis=1)
    assert (on_lattice >= 2).all()        # an edge vertex has at least two lattice-aligned coordinates
    assert rel.min() >= -1e-3 and (rel.max(axis=0) <= np.array(list(lat.n)) + 1e-3).all()
    stl = o.STLBytes()
    back = glrender.ReadBinarySTL(io.BytesIO(stl))
    assert np.array_equal(bits(back), bits(tris))
    h1 = hashlib.sha256(tris.tobytes()).digest()
    o.Rerun()
    assert hashlib.sha256(o.AllTriangles().tobytes()).digest() == h1


def test_knurled_large_prune_is_a_subsequence_of_flat(bld):
    """A larger lattice than the oracle can finish quickly (knurled @ resdiv 600, 32 M corners). The level-3 prune rule
    (octreerenderer.go:180-191) is only lossless for 1-Lipschitz fields; twist + smooth-k make this field slightly
    steeper, so -- exactly like the reference's Octree -- the pruned renderer may drop a few cells the dense one keeps.
    Properties that must hold: every pruned triangle is a dense triangle, in the same order; the loss is tiny; and
    pruning skips most evaluations."""
    s = gsdf.scene(bld, "knurled-cylinder")
    sdf = gleval.NewCUDASDF3(s)
    res = np.float32(s.Diagonal() / np.float32(600))
    f = glrender.FlatRenderer(sdf, res)
    o = glrender.Octree(sdf, res)
    nf, no = f.NumTriangles(), o.NumTriangles()
    assert 100000 < no <= nf and (nf - no) < 1e-3 * nf
    ft = np.ascontiguousarray(f.AllTriangles()).view(np.uint8).reshape(nf, 36)
    ot = np.ascontiguousarray(o.AllTriangles()).view(np.uint8).reshape(no, 36)
    fkeys = [x.tobytes() for x in ft]
    okeys = [x.tobytes() for x in ot]
    it = iter(fkeys)
    assert all(k in it for k in okeys)          # subsequence test: consumes `it` in order
    assert o.Evaluations() < 0.5 * f.Evaluations()


def test_program_update_reuses_the_handle(oracle, bld):
    """gsdf_program_update: an edited tree is re-uploaded into the same evaluator; a bound renderer sees it on its next
    run; the dimension cannot change."""
    a = bld.NewSphere(1.0)
    b = bld.Difference(bld.NewBox(1.6, 1.6, 1.6, 0.1), bld.NewCylinder(0.4, 3, 0))   # fits inside a's bounds
    sdf = gleval.NewCUDASDF3(a)
    res = np.float32(0.07)
    r = glrender.NewOctreeRenderer(sdf, res, 64)
    n_a = r.NumTriangles()
    pos = shapes.sample_points(a, dense=[20, 20, 20])
    out = np.empty(len(pos), np.float32)
    sdf.Evaluate(pos, out)
    assert np.array_equal(bits(out), bits(oracle.Tree.from_shader(a).eval3(pos)))
    sdf.Update(b)
    sdf.Evaluate(pos, out)
    assert np.array_equal(bits(out), bits(oracle.Tree.from_shader(b).eval3(pos)))
    r.Rerun()     # same lattice (sphere bounds), new tree
    lat = oracle.flat_lattice(*a.Bounds(), res)
    tb = oracle.Tree.from_shader(b)
    grid, _ = oracle.flat_eval_grid(tb, lat)
    mask, _, _ = oracle.octree_prune_plan(tb, lat, r.Plan())
    wt, _ = oracle.flat_march(lat, grid, blockmask=mask)
    assert r.NumTriangles() == len(wt) != n_a
    assert np.array_equal(bits(r.AllTriangles()), bits(wt))
    with pytest.raises(gsdf_b200.GsdfError):
        sdf.Update(bld.NewCircle(1.0))


def test_slab_pipeline_equals_single_renderer(bld):
    """glrender.SlabPipeline: slabs rendered one after the other with asynchronous reads reproduce the single renderer."""
    s = gsdf.scene(bld, "npt-flange")
    sdf = gleval.NewCUDASDF3(s)
    res = np.float32(s.Diagonal() / np.float32(220))
    whole = glrender.Octree(sdf, res).AllTriangles()
    for nslabs in (1, 2, 3, 5):
        pipe = glrender.SlabPipeline(sdf, res, nslabs=nslabs)
        dst = np.zeros((len(whole) + 8, 3, 3), np.float32)
        for _ in range(2):
            n = pipe.RenderToHost(dst)
            assert n == len(whole) == pipe.NumTriangles()
            assert np.array_equal(bits(dst[:n]), bits(whole)), nslabs
        pipe.Close()


def test_bounds_overload_and_driver(oracle, bld, capsys):
    """OverloadShader3DBounds changes the lattice (Bounds) but not the field; gsdfaux.RenderShader3D logs the
    reference's lines and writes a valid STL with one write."""
    from gsdf_b200 import gsdfaux
    s = bld.NewSphere(1.0)
    w = bld.OverloadShader3DBounds(s, (-1.5, -1.5, -1.5), (1.5, 1.5, 1.5))
    check_field("overload", w, oracle, shapes.sample_points(w))
    buf = io.BytesIO()
    tris = gsdfaux.RenderShader3D(w, gsdfaux.RenderConfig(STLOutput=buf, Resolution=0.09, UseGPU=True))
    out = capsys.readouterr().out
    assert "evaluated SDF" in out and "percent evaluations omitted in octree pruning step" in out and "render done" in out
    lat, grid, mask, wt, _, _ = oracle_mesh(oracle, w, np.float32(0.09), [(3, 1.25)])
    assert np.array_equal(bits(tris), bits(wt))
    buf.seek(0)
    assert np.array_equal(bits(glrender.ReadBinarySTL(buf)), bits(wt))
    with pytest.raises(gsdf_b200.GsdfError) as e:
        glrender.NewOctreeRenderer(gleval.NewCUDASDF3(s), 5.0, 64)     # makeICube: resolution not fine enough
    assert e.value.code == _lib.ERES
