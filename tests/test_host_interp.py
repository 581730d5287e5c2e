"""The interpreter source the GPU runs (interp.cuh), compiled for the host (tests/hostinterp.py), against the oracle: the
same bit-for-bit bar as the GPU parity tests, on the CPU. This is what validated the box-guard correction of DESIGN.md
section 2 on the real source before any GPU saw it (the previous revision of interp.cuh fails these tests on the
overlapping-operand shapes and on two of the seeded random trees)."""
import struct

import numpy as np
import pytest

import fontfix
import hostinterp
import progsim
import shapes
from gsdf_b200 import gsdf


def check(oracle, bld, name, s, pos=None, variant=None):
    pos = shapes.sample_points(s) if pos is None else pos
    t = oracle.Tree.from_shader(s)
    want = t.eval2(pos) if s.is2d else t.eval3(pos)
    got = hostinterp.run(bld.flatten(s), pos, variant)
    diff = (got.view(np.uint32) != want.view(np.uint32)) & ~(np.isnan(got) & np.isnan(want))
    assert not diff.any(), "%s: %d of %d distances differ from the oracle" % (name, int(diff.sum()), len(pos))


@pytest.mark.parametrize("corpus", ["all3d", "all2d", "dag3d"])
def test_device_interpreter_source_matches_oracle(oracle, bld, corpus):
    for name, s in getattr(shapes, corpus)(bld):
        check(oracle, bld, name, s)


def test_device_interpreter_source_on_random_trees(oracle, bld):
    for dim in (3, 2):
        for seed in (1, 2, 3):
            for name, s in shapes.random_trees(bld, seed, 40, dim):
                check(oracle, bld, name, s)
        for name, s in shapes.random_trees(bld, 77, 60, dim, depth=5, rich=True):
            check(oracle, bld, name, s)


def test_device_interpreter_source_guards(oracle, bld, monkeypatch):
    """Guards vote over the 4 points of one machine here (the finest tiling the source allows): slab guards on the
    BASELINE scenes, box guards on overlapping operands and on a text line, each equal to the oracle; unguarded too."""
    for name, s in shapes.overlap2d(bld):
        check(oracle, bld, name, s, shapes.overlap2d_points(name, s)[:8192])
    text = fontfix.text_scene(bld, "A8b")
    mn, mx = text.Bounds()
    pos = shapes.append_grid(mn - 0.05, mx + 0.05, [160, 60])
    check(oracle, bld, "text", text, pos)
    flange = gsdf.scene(bld, "npt-flange")
    dense = shapes.sample_points(flange, dense=[48, 48, 24])
    check(oracle, bld, "flange", flange, dense)
    monkeypatch.setenv("GSDF_NO_GUARDS", "1")
    check(oracle, bld, "text/no guards", text, pos)
    check(oracle, bld, "flange/no guards", flange, dense)


def test_radius_reuse_device_code_and_the_build_without_it(oracle, bld, monkeypatch):
    """Radius reuse (Machine::radius and its four consumers) is part of the default build: flagged programs equal the oracle
    bit for bit; so do unflagged programs (GSDF_RXY=0) on a -DGSDF_NO_RXY build of the same source."""
    flagged = 0
    for name, s in shapes.all3d(bld) + shapes.all2d(bld) + shapes.dag3d(bld) + shapes.random_trees(bld, 5, 60, 3, depth=5, rich=True):
        words = np.frombuffer(bld.flatten(s)["blob"], np.uint32, offset=32).reshape(-1, 4)
        pc, hit = 0, False
        while True:   # instruction headers only: CYLINDER / TORUS / CIRCLE2D keep the flags in w1, SCREW_ENTER in w2
            op, ln = int(words[pc, 0]) & 0xff, (int(words[pc, 0]) >> 8) & 0xff
            if op in (progsim.OP["CYLINDER"], progsim.OP["TORUS"], progsim.OP["CIRCLE2D"]):
                hit |= bool(int(words[pc, 1]) & 0x300)
            if op == progsim.OP["SCREW_ENTER"]:
                hit |= bool(int(words[pc, 2]) & 0x300)
            if op == 0:
                break
            pc += ln
        flagged += hit
        check(oracle, bld, name, s)
    assert flagged >= 5
    monkeypatch.setenv("GSDF_RXY", "0")
    for name, s in shapes.threads3d(bld) + shapes.scenes3d(bld):
        check(oracle, bld, name, s, variant="GSDF_NO_RXY")


def test_math32_source_matches_oracle(oracle):
    """math32.cuh (the kernels' elementary functions, host branches) against the oracle's independent restatement of
    chewxy/math32, input by input, special values included: bit-equal everywhere, NaN propagation of Min/Max included
    (min.NaN.f32 / max.NaN.f32 on the device); the one corner left is Min(NaN, -Inf) / Max(NaN, +Inf), where Go tests the
    infinity first."""
    import ctypes as C
    L = C.CDLL(hostinterp.build())
    OL = oracle.lib()
    rng = np.random.default_rng(0)
    special = [0.0, -0.0, 1, -1, 0.5, -0.5, 1e-30, -1e-30, np.inf, -np.inf, np.nan, 1e30, -1e30, 0.66, 2.4142137, 0.41421357, 3.1415927, 1.5707964]

    def inputs(lo, hi, n=4000):
        return np.concatenate([rng.uniform(lo, hi, n), rng.normal(0, 1, n // 4) * (hi - lo) / 8, special]).astype(np.float32)

    def differs(a, b):
        return (a.view(np.uint32) != b.view(np.uint32)) & ~(np.isnan(a) & np.isnan(b))

    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    unary = [(0, "go_sin", (-50, 50)), (1, "go_cos", (-50, 50)), (2, "go_tan", (-10, 10)), (3, "go_atan", (-100, 100)), (4, "go_acos", (-1.2, 1.2)),
             (5, "go_cbrt", (-1000, 1000)), (6, "go_log", (1e-6, 1000)), (7, "go_exp", (-90, 90)), (8, "go_sin", (-50, 50)), (9, "go_cos", (-50, 50)),
             (10, "go_sqrt", (0, 1e6))]
    for fn, ofn, (lo, hi) in unary:
        x = inputs(lo, hi)
        out = np.empty_like(x)
        L.host_m32_unary(fn, vp(x), vp(out), C.c_size_t(len(x)))
        want = np.array([getattr(OL, ofn)(C.c_float(float(v))) for v in x], np.float32)
        assert not differs(out, want).any(), (ofn, fn)
    for fn, ofn in ((0, "go_hypot"), (1, "go_atan2"), (2, "go_min"), (3, "go_max")):
        x, y = inputs(-100, 100), rng.permutation(inputs(-100, 100))
        out = np.empty_like(x)
        L.host_m32_binary(fn, vp(x), vp(y), vp(out), C.c_size_t(len(x)))
        want = np.array([getattr(OL, ofn)(C.c_float(float(a)), C.c_float(float(b))) for a, b in zip(x, y)], np.float32)
        d = differs(out, want)
        if fn >= 2:   # Go: Min(NaN, -Inf) = -Inf, Max(NaN, +Inf) = +Inf; min.NaN / max.NaN give NaN
            d &= ~((np.isnan(x) | np.isnan(y)) & (np.isinf(x) | np.isinf(y)))
        assert not d.any(), ofn


@pytest.mark.parametrize("scene,resdiv,stride", [("npt-flange", 400, 1), ("bolt", 800, 32), ("knurled-cylinder", 1600, 128)])
def test_full_size_baseline_lattices(oracle, bld, scene, resdiv, stride):
    """BASELINE configs 1-3 at their full resolution: every corner of the npt-flange@400 lattice (6,711,685 points), every
    32nd / 128th corner plane of bolt@800 and knurled-cylinder@1600 -- the interpreter source equals the oracle on each."""
    s = gsdf.scene(bld, scene)
    res = np.float32(s.Diagonal() / np.float32(resdiv))
    lat = oracle.flat_lattice(*s.Bounds(), res)
    nx, ny, nz = lat.n
    ax = [(np.float32(lat.origin[a]) + np.arange(n + 1, dtype=np.float32) * res).astype(np.float32) for a, n in enumerate((nx, ny, nz))]
    Z, Y, X = np.meshgrid(ax[2][::stride], ax[1], ax[0], indexing="ij")
    pos = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1).astype(np.float32)
    if scene == "npt-flange":
        assert len(pos) == 6711685   # README.md:130 (+1 probe)
    check(oracle, bld, scene, s, pos)


def far_and_nan_points(s, n=3000, seed=5):
    """Positions far outside Bounds() (up to 1e7 times the size: where ellipse2D and friends produce NaN in the Go formulas
    too) plus points with a NaN coordinate. No infinities: Min(NaN, -Inf) is the one documented corner (math32.cuh)."""
    rng = np.random.default_rng(seed)
    mn, mx = s.Bounds()
    dim = len(mn)
    c, h = (mn + mx) / 2, (mx - mn) / 2 + 1e-3
    mag = 10.0 ** rng.uniform(0, 7, (n, 1))
    pos = (c + rng.uniform(-1, 1, (n, dim)) * h * mag).astype(np.float32)
    pos[::97, rng.integers(0, dim)] = np.nan
    return pos


def test_nan_propagation_of_min_max_matches_go(oracle, bld):
    """math32.Min / Max propagate NaN (Go math semantics); so do the device's min.NaN.f32 / max.NaN.f32. Far-field and NaN
    inputs through every node type: where the oracle yields NaN the interpreter source must, and the other way round."""
    for corpus in ("all3d", "all2d"):
        for name, s in getattr(shapes, corpus)(bld):
            check(oracle, bld, name + "/far", s, far_and_nan_points(s))
    for dim in (3, 2):
        for name, s in shapes.random_trees(bld, 5, 30, dim):
            check(oracle, bld, name + "/far", s, far_and_nan_points(s))


def array_far_cases(bld):
    """Array / Array2D / TranslateMulti2D evaluated so far away that the child's distance exceeds the reference's fold seeds
    (largenum = 1e20, gsdf.go:21; math.MaxFloat32 in translateMulti2D): the result is the seed, not the distance."""
    f = np.float32
    a3 = bld.Array(bld.NewSphere(0.4), 1.0, 1.2, 0.9, 3, 2, 2)
    a2 = bld.Array2D(bld.NewCircle(0.3), 1.0, 0.8, 3, 2)
    tm = bld.TranslateMulti2D(bld.NewCircle(0.3), np.array([[0, 0], [1, 0.5], [2, -0.25]], f))
    big3 = np.array([[1e22, 0, 0], [0, -3e24, 1e21], [5e19, 5e19, 5e19], [1e19, 0, 0], [0.3, 0.2, 0.1]], f)
    big2 = np.array([[1e22, 0], [0, -3e24], [9e19, 9e19], [1e19, 0], [0.3, 0.2], [3e38, 3e38]], f)
    return [("array/far", a3, big3), ("array2d/far", a2, big2), ("translatemulti2d/far", tm, big2)]


def test_array_folds_start_from_the_reference_seed(oracle, bld):
    for name, s, pos in array_far_cases(bld):
        check(oracle, bld, name, s, pos)
        t = oracle.Tree.from_shader(s)
        want = t.eval2(pos) if s.is2d else t.eval3(pos)
        if "translatemulti" not in name:
            assert want[0] == np.float32(1e20)   # the seed shows


def device_image(flat):
    """The program as gsdf_program_create uploads it (gsdf_program_device_image): chunks + side buffer with the operand
    tables the library appends (Sincos tables of circular arrays), repacked as a flatten() dict for hostinterp.run."""
    import ctypes as C
    from gsdf_b200 import _lib
    blob = flat["blob"]
    aux = np.ascontiguousarray(flat["aux"], np.float32)
    auxp = aux.ctypes.data_as(C.POINTER(C.c_float))
    need = _lib.lib.gsdf_program_device_image(blob, len(blob), auxp, aux.size, None, 0)
    assert need > 0, _lib.last_error()
    img = (C.c_uint8 * need)()
    assert _lib.lib.gsdf_program_device_image(blob, len(blob), auxp, aux.size, img, need) == need
    nchunks = struct.unpack_from("<8I", blob, 0)[2]
    raw = bytes(img)
    return {"blob": bytes(blob[:32]) + raw[:16 * nchunks], "aux": np.frombuffer(raw, np.float32, offset=16 * nchunks).copy()}, nchunks


def test_circular_array_tables_of_the_device_image(oracle, bld):
    """The library replaces the two Sincos per point of a circular array by loads from a table it appends to the side buffer
    when a program is uploaded (capi.cu augment_program, interp.cuh CIRC_ENTER). The augmented image -- exactly what the device
    gets -- run through the interpreter source equals the oracle bit for bit: the circular arrays of the corpora (3-D and 2-D,
    partial and full circles), the knurled cylinder on lattice planes, and far / NaN positions (which leave the table)."""
    seen = 0
    cases = [(n, s) for n, s in shapes.all3d(bld) + shapes.all2d(bld) if "circarray" in n]
    cases.append(("knurled-cylinder", gsdf.scene(bld, "knurled-cylinder")))
    for name, s in cases:
        flat = bld.flatten(s)
        img, nchunks = device_image(flat)
        words = np.frombuffer(img["blob"], np.uint32, offset=32)
        ntab = 0
        pc = 0
        while pc < nchunks:
            op, ln = int(words[4 * pc]) & 0xff, (int(words[4 * pc]) >> 8) & 0xff
            if op == progsim.OPS.index("CIRC_ENTER"):
                assert words[4 * (pc + 1) + 3] != 0, name
                ntab += 1
            pc += ln if ln else 1
        assert ntab >= 1 and len(img["aux"]) > len(flat["aux"]), name
        seen += ntab
        t = oracle.Tree.from_shader(s)
        for pos in (shapes.sample_points(s), far_and_nan_points(s)):
            want = t.eval2(pos) if s.is2d else t.eval3(pos)
            for f in (img, flat):  # with the tables and (the blob as flattened) without
                got = hostinterp.run(f, pos)
                diff = (got.view(np.uint32) != want.view(np.uint32)) & ~(np.isnan(got) & np.isnan(want))
                assert not diff.any(), "%s: %d of %d distances differ from the oracle" % (name, int(diff.sum()), len(pos))
    assert seen >= 4
    # seeded random compositions that contain a circular array somewhere in the tree
    circ, nrand = progsim.OPS.index("CIRC_ENTER"), 0
    for dim in (3, 2):
        for name, s in shapes.random_trees(bld, 11, 120, dim) + shapes.random_trees(bld, 78, 60, dim, depth=5, rich=True):
            flat = bld.flatten(s)
            words = np.frombuffer(flat["blob"], np.uint32, offset=32)
            nchunks = struct.unpack_from("<8I", flat["blob"], 0)[2]
            pc, has = 0, False
            while pc < nchunks:
                op, ln = int(words[4 * pc]) & 0xff, (int(words[4 * pc]) >> 8) & 0xff
                has |= op == circ
                pc += ln if ln else 1
            if not has:
                continue
            nrand += 1
            img, _ = device_image(flat)
            pos = shapes.sample_points(s)
            t = oracle.Tree.from_shader(s)
            want = t.eval2(pos) if s.is2d else t.eval3(pos)
            got = hostinterp.run(img, pos)
            diff = (got.view(np.uint32) != want.view(np.uint32)) & ~(np.isnan(got) & np.isnan(want))
            assert not diff.any(), "%s: %d of %d distances differ from the oracle" % (name, int(diff.sum()), len(pos))
    assert nrand >= 5, nrand


def test_device_image_owns_the_table_word(bld):
    """The fourth operand word of CIRC_ENTER belongs to the library: a blob that carries something there (a flattener bug, a
    hostile blob) gets it overwritten -- by the table position, or by 0 when no table is built -- so the kernel can never be
    sent to read outside the side buffer."""
    s = [s for n, s in shapes.all3d(bld) if n == "circarray"][0]
    flat = bld.flatten(s)
    blob = bytearray(flat["blob"])
    words = np.frombuffer(bytes(blob), np.uint32, offset=32)
    nchunks = struct.unpack_from("<8I", blob, 0)[2]
    circ, pc, at = progsim.OPS.index("CIRC_ENTER"), 0, None
    while pc < nchunks:
        op, ln = int(words[4 * pc]) & 0xff, (int(words[4 * pc]) >> 8) & 0xff
        if op == circ:
            at = pc + 1
        pc += ln if ln else 1
    assert at is not None
    struct.pack_into("<I", blob, 32 + 16 * at + 12, 0x7fffffff)
    good, _ = device_image(flat)
    bad, _ = device_image({"blob": bytes(blob), "aux": flat["aux"]})
    assert good["blob"] == bad["blob"] and np.array_equal(good["aux"], bad["aux"])
    # a fractional instance count builds no table: the word is cleared
    struct.pack_into("<f", blob, 32 + 16 * at + 4, 6.5)
    img, _ = device_image({"blob": bytes(blob), "aux": flat["aux"]})
    assert struct.unpack_from("<I", img["blob"], 32 + 16 * at + 12)[0] == 0 and len(img["aux"]) == len(np.atleast_1d(flat["aux"]))
