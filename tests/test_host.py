"""Host layer (no GPU needed): Builder validation and Bounds() mirror the reference's rules, the flattener emits a
well-formed program, and the C ABI library exports every symbol its headers declare."""
import ctypes as C
import os
import re
import struct

import numpy as np
import pytest

import gsdf_b200
from gsdf_b200 import gsdf, _lib
import shapes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """include/gsdf_b200.h is exported by libgsdfb200.so, include/gsdf_host.h by libgsdfhost.so, and the ctypes binding
    covers exactly the declared set. The host library has no CUDA dependency and calls nothing of the device library."""
    declared = set()
    for h, path in (("gsdf_b200.h", _lib.LIB_PATH), ("gsdf_host.h", _lib.HOST_LIB_PATH)):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names = set(re.findall(r"\b(gsdfh?_[a-z0-9_]+)\s*\(", src))
        assert len(names) >= 30
        lib = C.CDLL(path)
        missing = [s for s in sorted(names) if not hasattr(lib, s)]
        assert not missing, (h, missing)
        declared |= names
    bound = set(_lib.lib._gsdf_signatures)
    assert declared == bound, (declared ^ bound)
    import subprocess
    needed = subprocess.run(["readelf", "-d", _lib.HOST_LIB_PATH], capture_output=True, text=True).stdout
    assert "cudart" not in needed and "libcuda" not in needed and "gsdfb200" not in needed
    undefined = subprocess.run(["nm", "-D", "--undefined-only", _lib.HOST_LIB_PATH], capture_output=True, text=True).stdout
    assert "gsdf_" not in undefined and "cuda" not in undefined.lower()


def test_no_cpu_fallback_without_a_device(bld):
    """Without a CUDA device compute entry points fail with GSDF_ECUDA; nothing silently runs on the CPU."""
    if gsdf_b200.device_count() > 0:
        pytest.skip("a CUDA device is present")
    from gsdf_b200 import gleval
    with pytest.raises(gsdf_b200.GsdfError) as e:
        gleval.NewCUDASDF3(bld.NewSphere(1))
    assert e.value.code == _lib.ECUDA
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_shape_errors_match_reference_rules():
    b = gsdf.Builder()
    for bad in (lambda: b.NewSphere(0), lambda: b.NewSphere(-1),                       # primitives.go:29
                lambda: b.NewBox(1, 1, 1, 0.6), lambda: b.NewBox(0, 1, 1, 0),          # :66-71
                lambda: b.NewCylinder(1, 1, 0.5), lambda: b.NewCylinder(1, 0, 0),      # :108-115
                lambda: b.NewTorus(1, 0.6), lambda: b.NewBoxFrame(1, 1, 1, 1.5),       # :217, :260
                lambda: b.NewHexagonalPrism(0, 1),                                     # :158
                lambda: b.Symmetry(b.NewSphere(1), False, False, False),               # operations.go:286
                lambda: b.CircularArray(b.NewSphere(1), 5, 4),                         # :771
                lambda: b.CircularArray(b.NewSphere(1), 1, 1),                         # :768
                lambda: b.Array(b.NewSphere(1), 1, 1, 1, 0, 1, 1),                     # :489
                lambda: b.Twist(b.NewSphere(1), 0),                                    # :839
                lambda: b.Rotate(b.NewSphere(1), 1, (0, 0, 0)),                        # :395
                lambda: b.NewPolygon([[0, 0], [1, 1]]),                                # primitives2d.go:477
                lambda: b.NewPolygon([[0, 0], [1, 1], [1, 1], [0, 1]]),                # :483
                lambda: b.NewCircle(0), lambda: b.NewRectangle(1, -1),
                lambda: b.Annulus(b.NewCircle(1), 0),                                  # operations2d.go:608
                lambda: b.Extrude(b.NewCircle(1), -1),                                 # :108
                lambda: b.Union(b.NewSphere(1)),                                       # operations.go:36
                lambda: b.Difference(b.NewSphere(1), b.NewCircle(1)),                  # 2D passed as 3D
                ):
        with pytest.raises(gsdf.ShapeError):
            bad()
    # FlagNoDimensionPanic behaviour: errors accumulate (gsdf.go:100-106)
    b2 = gsdf.Builder(panic_on_error=False)
    b2.NewSphere(-1)
    b2.NewCylinder(1, 1, 2)
    assert "sphere" in b2.Err() and "cylinder" in b2.Err()
    b2.ClearErrors()
    assert b2.Err() == ""


def test_degenerate_line_becomes_circle(bld):
    """primitives2d.go:23-28"""
    s = bld.NewLine2D(1, 1, 1, 1, 0.5)
    mn, mx = s.Bounds()
    assert np.allclose(mn, [-0.25, -0.25]) and np.allclose(mx, [0.25, 0.25])


def test_bounds_mirror_reference(bld):
    s = bld.NewSphere(2)
    assert np.array_equal(np.concatenate(s.Bounds()), [-2, -2, -2, 2, 2, 2])
    c = bld.NewCylinder(1, 3, 0.1)
    assert np.array_equal(np.concatenate(c.Bounds()), [-1, -1, -1.5, 1, 1, 1.5])
    h = bld.NewHexagonalPrism(1, 2)  # primitives.go:169-176: z extent is +-h (not h/2)
    assert np.allclose(np.concatenate(h.Bounds()), [-1 / 0.8660254, -1, -2, 1 / 0.8660254, 1, 2])
    t = bld.Translate(s, 1, 2, 3)
    assert np.array_equal(np.concatenate(t.Bounds()), [-1, 0, 1, 3, 4, 5])
    d = bld.Difference(s, t)
    assert np.array_equal(np.concatenate(d.Bounds()), np.concatenate(s.Bounds()))          # operations.go:128
    i = bld.Intersection(s, t)
    assert np.array_equal(np.concatenate(i.Bounds()), [-1, 0, 1, 2, 2, 2])                  # :171
    sc = bld.Scale(t, 2)
    assert np.array_equal(np.concatenate(sc.Bounds()), [-2, 0, 2, 6, 8, 10])                # :257 (about the origin)
    sy = bld.Symmetry(t, True, False, False)
    assert np.array_equal(np.concatenate(sy.Bounds()), [-3, 0, 1, 3, 4, 5])                 # :297
    of = bld.Offset(s, -0.5)
    assert np.array_equal(np.concatenate(of.Bounds()), [-2.5, -2.5, -2.5, 2.5, 2.5, 2.5])   # :455
    ar = bld.Array(s, 5, 5, 5, 2, 3, 1)
    assert np.array_equal(np.concatenate(ar.Bounds()), [-2, -2, -2, 12, 17, 7])             # :504
    el = bld.Elongate(s, 2, 0, 4)
    assert np.array_equal(np.concatenate(el.Bounds()), [-3, -2, -4, 3, 2, 4])               # :688
    ex = bld.Extrude(bld.NewRectangle(2, 4), 6)
    assert np.array_equal(np.concatenate(ex.Bounds()), [-1, -2, -3, 1, 2, 3])               # operations2d.go:119
    rv = bld.Revolve(bld.Translate2D(bld.NewRectangle(2, 4), 3, 0), 1)
    assert np.array_equal(np.concatenate(rv.Bounds()), [-3, -2, -3, 3, 2, 3])               # :168
    tw = bld.Twist(bld.NewBox(2, 2, 4, 0), 0.3)
    r = np.float32(np.hypot(1, 1))
    assert np.allclose(np.concatenate(tw.Bounds()), [-r, -r, -2, r, r, 2])                  # operations.go:850
    u = bld.Union(s, t, bld.Translate(s, -5, 0, 0))
    assert np.array_equal(np.concatenate(u.Bounds()), [-7, -2, -2, 3, 4, 5])
    o2 = bld.Offset2D(bld.NewRectangle(2, 2), 0.5)                                          # operations2d.go:421-429: f>0 keeps bb
    assert np.array_equal(np.concatenate(o2.Bounds()), [-1, -1, 1, 1])
    o3 = bld.Offset2D(bld.NewRectangle(2, 2), -0.5)
    assert np.array_equal(np.concatenate(o3.Bounds()), [-1.5, -1.5, 1.5, 1.5])


def test_union_absorbs_nested_unions(bld):
    """operations.go:44-48"""
    a, b, c = bld.NewSphere(1), bld.NewSphere(2), bld.NewSphere(3)
    u = bld.Union(bld.Union(a, b), c)
    nb, ch, aux = bld.tree_table()
    node = _lib.TreeNode.from_buffer_copy(nb[u.id * 96:(u.id + 1) * 96])
    assert node.nchild == 3 and list(ch[node.child_off:node.child_off + 3]) == [a.id, b.id, c.id]


def test_scene_trees(bld):
    fl = gsdf.scene(bld, "npt-flange")
    f = bld.flatten(fl)
    assert f["dim"] == 3 and f["ninstr"] >= 12 and f["pstack"] >= 1 and f["dstack"] >= 2
    assert abs(fl.Diagonal() - 86.71794) < 1e-3
    kn = gsdf.scene(bld, "knurled-cylinder")
    assert np.array_equal(np.concatenate(kn.Bounds()), [-10, -10, -25, 10, 10, 25])
    assert abs(kn.Diagonal() / np.float32(1600) - 0.035903517) < 1e-8
    bo = gsdf.scene(bld, "bolt")
    assert bld.flatten(bo)["ninstr"] > 15
    # NPT 1/2 internal ISO profile: 7 base vertices, apex smoothed with 5 facets -> 12 vertices (iso.go:59-70)
    prof = gsdf.threads.Thread(bld, gsdf.threads.NPT(0.5))
    nb, ch, aux = bld.tree_table()
    node = _lib.TreeNode.from_buffer_copy(nb[prof.id * 96:(prof.id + 1) * 96])
    assert node.aux_cnt == 24
    ext = gsdf.threads.Thread(bld, gsdf.threads.ISO(3, 0.5, True))                          # 8 base, 2 smoothed -> 18
    nb, ch, aux = bld.tree_table()
    node = _lib.TreeNode.from_buffer_copy(nb[ext.id * 96:(ext.id + 1) * 96])
    assert node.aux_cnt == 36


def test_flattened_programs_are_well_formed(bld):
    """Every corpus shape flattens; the stream is a chain of length-prefixed instructions ending in END."""
    for name, s in shapes.all3d(bld) + shapes.all2d(bld):
        f = bld.flatten(s)
        blob = f["blob"]
        magic, ver, nchunks, dim, dstack, pstack, ninstr, _ = struct.unpack_from("<8I", blob, 0)
        assert magic == 0x46445347 and ver == 1 and dim == (2 if s.is2d else 3), name
        assert len(blob) == 32 + 16 * nchunks and f["aux"].size % 4 == 0, name
        words = np.frombuffer(blob, np.uint32, offset=32).reshape(-1, 4)
        pc = n = depth_p = 0
        last = None
        while pc < nchunks:
            op, ln = int(words[pc, 0]) & 0xff, (int(words[pc, 0]) >> 8) & 0xff
            assert ln >= 1, name
            last = op
            pc += ln
            n += 1
        assert pc == nchunks and last == 0 and n == ninstr, name
        assert 1 <= dstack <= 16 and pstack <= 8, name


def test_random_trees_flatten_within_stack_limits(oracle, bld):
    """The seeded random trees of the GPU fuzz test (tests/shapes.py): every one flattens to a well-formed program, the
    oracle evaluates it to finite values, and the generator is reproducible (same seed, same programs)."""
    import shapes
    for dim in (3, 2):
        a = shapes.random_trees(bld, 1, 25, dim)
        b = shapes.random_trees(bld, 1, 25, dim)
        for (name, s), (_, s2) in zip(a, b):
            f, f2 = bld.flatten(s), bld.flatten(s2)
            assert f["blob"] == f2["blob"] and np.array_equal(f["aux"], f2["aux"]), name
            assert 1 <= f["dstack"] <= 16 and f["pstack"] <= 8, name
            t = oracle.Tree.from_shader(s)
            pos = shapes.sample_points(s)
            d = t.eval3(pos) if dim == 3 else t.eval2(pos)
            assert np.isfinite(d).all(), name


def test_shared_subtrees_flatten_like_duplicated_ones(oracle, bld):
    """TestTransformDuplicateBug (gsdf_test.go:90-133): a node reachable through several parents (a DAG) must behave like
    separate identical nodes -- same program bytes, same oracle field."""
    import shapes
    a, b = shapes.geb(bld, True), shapes.geb(bld, False)
    fa, fb = bld.flatten(a), bld.flatten(b)
    assert fa["blob"] == fb["blob"] and np.array_equal(fa["aux"], fb["aux"])
    assert np.array_equal(a.Bounds()[0], b.Bounds()[0]) and np.array_equal(a.Bounds()[1], b.Bounds()[1])
    pos = shapes.sample_points(a)
    da, db = oracle.Tree.from_shader(a).eval3(pos), oracle.Tree.from_shader(b).eval3(pos)
    assert np.array_equal(da.view(np.uint32), db.view(np.uint32)) and (da < 0).any() and (da > 0).any()
    for name, s in shapes.dag3d(bld):
        assert np.isfinite(oracle.Tree.from_shader(s).eval3(shapes.sample_points(s))).all(), name


def test_program_create_checks_the_stack_discipline_before_any_device_work(bld):
    """gsdf_program_create validates the blob before it touches a device (so this runs without a GPU): the kernels size
    their shared-memory stacks from the header, and a program that pushes deeper than its header declares, pops an empty
    stack or leaves values behind is refused. Every program of the flattener passes (ECUDA here = accepted, no device)."""
    import shapes
    create = _lib.lib.gsdf_program_create

    def rc_of(blob, aux):
        aux = np.ascontiguousarray(aux, np.float32)
        h = C.c_void_p()
        rc = create(bytes(blob), len(blob), aux.ctypes.data_as(C.POINTER(C.c_float)) if aux.size else None, aux.size, C.byref(h))
        if rc == 0:
            _lib.lib.gsdf_program_destroy(h)
        return rc
    accepted = (_lib.OK, _lib.ECUDA)
    for name, s in shapes.all3d(bld) + shapes.all2d(bld) + shapes.dag3d(bld) + shapes.overlap2d(bld) + shapes.random_trees(bld, 9, 60, 3, depth=5, rich=True):
        f = bld.flatten(s)
        assert rc_of(f["blob"], f["aux"]) in accepted, name
    f = bld.flatten(gsdf.scene(bld, "npt-flange"))     # dstack 2, pstack 1
    hdr = list(struct.unpack_from("<8I", f["blob"], 0))
    assert (hdr[4], hdr[5]) == (2, 1)
    for field, value in ((4, 1), (5, 0)):               # header declares fewer slots than the stream uses
        h2 = list(hdr)
        h2[field] = value
        assert rc_of(struct.pack("<8I", *h2) + f["blob"][32:], f["aux"]) == _lib.EPROGRAM
        assert "stack slots" in _lib.last_error()
    words = np.frombuffer(f["blob"], np.uint32, offset=32).reshape(-1, 4).copy()
    headers, pc = [], 0                                                        # (chunk index, opcode) of every instruction
    while True:
        op, ln = int(words[pc, 0]) & 0xff, (int(words[pc, 0]) >> 8) & 0xff
        headers.append((pc, op))
        if op == 0:
            break
        pc += ln
    PUSH_POS, POP_POS, DIFF, OFFSET, MIN = 34, 35, 22, 27, 20
    pop = [i for i, op in headers if op == POP_POS][0]
    bad = words.copy()
    bad[pop, 0] = (bad[pop, 0] & ~np.uint32(0xff)) | np.uint32(PUSH_POS)      # a POP_POS turned into a second PUSH_POS
    assert rc_of(f["blob"][:32] + bad.tobytes(), f["aux"]) == _lib.EPROGRAM
    bad = words.copy()
    diff = [i for i, op in headers if op == DIFF][-1]                          # the last DIFF becomes OFFSET: a value is left behind
    bad[diff, 0] = (bad[diff, 0] & ~np.uint32(0xff)) | np.uint32(OFFSET)
    assert rc_of(f["blob"][:32] + bad.tobytes(), f["aux"]) == _lib.EPROGRAM
    # a guard whose skipped region does not end right in front of its combiner: retarget the flange's screw guard (DIFF kind,
    # lands on the POP_POS in front of the DIFF) one instruction further, behind the DIFF
    SCREW_ENTER = 50
    scr = [i for i, op in headers if op == SCREW_ENTER][0]
    kind, target = int(words[scr, 1]) & 0xff, int(words[scr, 1]) >> 8
    assert kind == 1 and dict(headers)[target] == POP_POS
    after = [i for i, op in headers if i > target and op == DIFF][0]
    nxt = headers[[i for i, _ in headers].index(after) + 1][0]
    bad = words.copy()
    bad[scr, 1] = np.uint32(kind | (nxt << 8))
    assert rc_of(f["blob"][:32] + bad.tobytes(), f["aux"]) == _lib.EPROGRAM and "guard" in _lib.last_error()
    bad = words.copy()
    bad[scr, 1] = np.uint32(2 | (target << 8))                                 # MIN kind in front of a DIFF
    assert rc_of(f["blob"][:32] + bad.tobytes(), f["aux"]) == _lib.EPROGRAM and "combiner" in _lib.last_error()
    sphere = bld.flatten(bld.NewSphere(1))
    w = np.frombuffer(sphere["blob"], np.uint32, offset=32).reshape(-1, 4).copy()
    w[0, 0] = (w[0, 0] & ~np.uint32(0xff)) | np.uint32(MIN)                   # MIN on an empty stack
    assert rc_of(sphere["blob"][:32] + w.tobytes(), sphere["aux"]) == _lib.EPROGRAM and "underflow" in _lib.last_error()


def test_position_liveness(bld):
    """A transform whose position nobody reads again must not save it: scale(translate(sphere)) needs no P slots,
    union(translate(a), b) needs one."""
    s = bld.NewSphere(1)
    assert bld.flatten(bld.Scale(bld.Translate(s, 1, 0, 0), 2))["pstack"] == 0
    assert bld.flatten(bld.Union(bld.Translate(s, 1, 0, 0), s))["pstack"] == 1
    assert bld.flatten(bld.Union(s, bld.Translate(s, 1, 0, 0)))["pstack"] == 0
    assert bld.flatten(bld.CircularArray(s, 3, 5))["pstack"] == 1   # CIRC_ENTER parks p0 on the stack


def test_unknown_constructor_kinds_fail_loudly(bld):
    import ctypes
    from gsdf_b200._lib import lib
    f = (ctypes.c_float * 2)(1.0, 2.0)
    assert lib.gsdfh_node(bld._h, 999, f, 2, None, 0, None, 0, None, 0) < 0
    assert "unsupported constructor kind" in bld.Err()
    bld.ClearErrors()


def test_bounds_overload_wrapper(oracle, bld):
    """glbuild.OverloadShader3DBounds (glbuild.go:1080-1102): Bounds() is replaced, Evaluate forwards, and the flattener
    sees through the wrapper (same program as the wrapped shader)."""
    s = bld.NewSphere(1.0)
    w = bld.OverloadShader3DBounds(s, (-2, -2, -2), (2, 3, 4))
    assert np.array_equal(np.concatenate(w.Bounds()), [-2, -2, -2, 2, 3, 4])
    assert bld.flatten(w)["blob"] == bld.flatten(s)["blob"]
    pts = np.random.default_rng(0).uniform(-2, 2, (64, 3)).astype(np.float32)
    assert np.array_equal(oracle.Tree.from_shader(w).eval3(pts), oracle.Tree.from_shader(s).eval3(pts))
    c = bld.NewCircle(1.0)
    w2 = bld.OverloadShader2DBounds(c, (-3, -3), (3, 3))
    assert np.array_equal(np.concatenate(w2.Bounds()), [-3, -3, 3, 3]) and w2.is2d
    assert bld.flatten(bld.Extrude(w2, 1))["ninstr"] == bld.flatten(bld.Extrude(c, 1))["ninstr"]


def test_octree_levels_formula(oracle):
    """makeICube (octreerenderer.go:222-235): sphere r=1, res=1/33 -> 8 levels; flange@400 -> 10; too coarse -> error."""
    from gsdf_b200 import glrender
    assert glrender.octree_levels((-1, -1, -1), (1, 1, 1), 1 / 33) == 8 == oracle.octree_levels((-1, -1, -1), (1, 1, 1), 1 / 33)
    assert glrender.octree_levels((-30, -30, -12.5), (30, 30, 5.3886), 0.21679485) == 10
    with pytest.raises(gsdf_b200.GsdfError) as e:
        glrender.octree_levels((-1, -1, -1), (1, 1, 1), 3.0)
    assert e.value.code == _lib.ERES
    with pytest.raises(gsdf_b200.GsdfError):
        glrender.octree_levels((-1, -1, -1), (1, 1, 1), 0.0)


def test_read_binary_stl_validation(oracle, bld):
    """ReadBinarySTL / stlTriangle.validate (stl.go:129-225)."""
    import io
    from gsdf_b200 import glrender
    s = bld.NewSphere(1.0)
    t = oracle.Tree.from_shader(s)
    lat = oracle.flat_lattice(*s.Bounds(), np.float32(0.3))
    grid, _ = oracle.flat_eval_grid(t, lat)
    tris, _ = oracle.flat_march(lat, grid)
    data = bytearray(oracle.stl_write(tris))
    back = glrender.ReadBinarySTL(io.BytesIO(bytes(data)))
    assert np.array_equal(back.view(np.uint32), tris.view(np.uint32))
    with pytest.raises(gsdf_b200.GsdfError):
        glrender.ReadBinarySTL(io.BytesIO(bytes(data[:50])))            # truncated header
    with pytest.raises(gsdf_b200.GsdfError):
        glrender.ReadBinarySTL(io.BytesIO(bytes(data[:84 + 49])))       # truncated record
    bad = bytearray(data)
    bad[80:84] = struct.pack("<I", 0)
    with pytest.raises(gsdf_b200.GsdfError):
        glrender.ReadBinarySTL(io.BytesIO(bytes(bad)))                  # zero triangles (stl.go:183)
    bad = bytearray(data)
    bad[84 + 12:84 + 16] = struct.pack("<f", float("nan"))
    with pytest.raises(gsdf_b200.GsdfError):
        glrender.ReadBinarySTL(io.BytesIO(bytes(bad)))                  # NaN vertex
    bad = bytearray(data)
    bad[84 + 24:84 + 48] = bad[84 + 12:84 + 24] * 2                     # all three vertices equal -> degenerate
    with pytest.raises(gsdf_b200.GsdfError):
        glrender.ReadBinarySTL(io.BytesIO(bytes(bad)))


def _instructions(blob):
    nchunks = struct.unpack_from("<8I", blob, 0)[2]
    words = np.frombuffer(blob, np.uint32, offset=32).reshape(-1, 4)
    pc, out = 0, []
    while pc < nchunks:
        op, ln = int(words[pc, 0]) & 0xff, (int(words[pc, 0]) >> 8) & 0xff
        out.append((pc, op, ln, int(words[pc, 1])))
        pc += ln
    return out


def _opcodes():
    import re
    src = open(os.path.join(os.path.dirname(__file__), "..", "include", "gsdf_program.h")).read()
    body = src[src.index("enum gsdf_opcode {"):src.index("GSDF_OP__COUNT")]
    return {n[len("GSDF_OP_"):]: i for i, n in enumerate(re.findall(r"^\s*(GSDF_OP_[A-Z0-9_]+)", body, re.M))}


def test_slab_guards_are_planted_where_they_are_sound(bld, monkeypatch):
    """include/gsdf_program.h "slab guards": a screw/extrude that is the LATER operand of difference / union / smooth
    union (through position-only wrappers) carries a guard whose target is the instruction right behind its own exit
    op; first operands, intersections and distance-changing wrappers carry none."""
    OP = _opcodes()
    enter = {OP["EXTRUDE_ENTER"]: OP["EXTRUDE_EXIT"], OP["SCREW_ENTER"]: OP["MAX_BELOW"]}
    want = {"diff_box_extrude": [1], "union_sphere_extrude": [2, 2], "smoothunion_cyl_extrude": [3],
            "smoothunion_extrude_first": [0], "diff_cyl_rotated_screw": [1], "union_screw_symmetry": [2],
            "nested_guards": [1], "guard_under_scale": [1], "union_two_extrudes": [0, 2]}
    for name, s in shapes.guards3d(bld):
        ins = _instructions(bld.flatten(s)["blob"])
        kinds = []
        for i, (pc, op, ln, w1) in enumerate(ins):
            if op not in enter:
                continue
            kinds.append(w1 & 0xff)
            if w1 & 0xff:
                depth, j = 0, i
                while True:  # the matching exit op of this node
                    if ins[j][1] in enter: depth += 1
                    if ins[j][1] in enter.values():
                        depth -= 1
                        if depth == 0: break
                    j += 1
                assert w1 >> 8 == ins[j + 1][0], name
        assert kinds == want[name], (name, kinds)
    # the flange's nut thread and the bolt's screw are the benchmark scenes' guarded nodes
    for scene, kind in (("npt-flange", 1), ("bolt", 2)):
        ins = _instructions(bld.flatten(gsdf.scene(bld, scene))["blob"])
        assert [w1 & 0xff for pc, op, ln, w1 in ins if op == OP["SCREW_ENTER"]] == [kind], scene
    monkeypatch.setenv("GSDF_NO_GUARDS", "1")
    for name, s in shapes.guards3d(bld):
        assert all((w1 & 0xff) == 0 for pc, op, ln, w1 in _instructions(bld.flatten(s)["blob"]) if op in enter), name


def test_multiply_shift_division_constants():
    """generators.cuh fastdiv_init / fastdiv (work-item decode of the lattice, prune-centre and kept-block kernels): for every
    divisor d the kernels can meet and every n < 2^31, __umulhi(n, mul) >> shr == n // d. Python restatement of the two
    functions (the GPU parity tests run the real ones on every lattice size of the suite)."""
    def init(d):
        if d <= 1:
            return 0, 0
        lg = 0
        while (1 << lg) < d:
            lg += 1
        p = 31 + lg
        return ((1 << p) + d - 1) // d, p - 32

    rng = np.random.default_rng(5)
    divisors = list(range(1, 3000)) + [int(x) for x in rng.integers(3000, 1 << 30, 400)] + [(1 << k) + s for k in range(2, 31) for s in (-1, 0, 1)]
    edge = np.array([0, 1, 2, (1 << 31) - 1, (1 << 31) - 2, 1 << 30], dtype=np.uint64)
    for d in divisors:
        mul, shr = init(d)
        assert mul < (1 << 32)
        n = np.concatenate([edge, rng.integers(0, 1 << 31, 64).astype(np.uint64),
                            (np.arange(1, 40, dtype=np.uint64) * np.uint64(d)) % np.uint64(1 << 31),
                            ((np.arange(1, 40, dtype=np.uint64) * np.uint64(d)) - np.uint64(1)) % np.uint64(1 << 31)])
        got = ((n * np.uint64(mul)) >> np.uint64(32)) >> np.uint64(shr) if mul else n
        assert np.array_equal(got, n // np.uint64(d)), d
