"""The host-side flattener checked without a GPU: the flattened program, executed by the CPU model of the device
interpreter (tests/progsim.py), must reproduce the oracle's evaluation of the TREE bit for bit -- operand order, derived
float32 constants, distance / position stack slots, position liveness, slab-guard and box-guard jump targets. On the GPU
box the same programs run through the CUDA interpreter against the same oracle (tests/test_gpu_parity.py)."""
import numpy as np
import pytest

import fontfix
import progsim
import shapes
from gsdf_b200 import gsdf


@pytest.fixture(scope="module")
def M(oracle):
    return progsim.Math(oracle)


def sim_vs_oracle(oracle, bld, M, name, s, pos=None, tile=progsim.TILE, stats=None):
    f = bld.flatten(s)
    P = progsim.Program(f["blob"], f["aux"])
    assert (P.dim, P.dstack, P.pstack) == (2 if s.is2d else 3, f["dstack"], f["pstack"]), name
    if not P.supported():
        return None
    pos = shapes.sample_points(s) if pos is None else pos
    t = oracle.Tree.from_shader(s)
    want = t.eval2(pos) if s.is2d else t.eval3(pos)
    got = progsim.run(P, pos, M, tile=tile, stats=stats)
    nbad = int((got.view(np.uint32) != want.view(np.uint32)).sum())
    assert nbad == 0, "%s: %d of %d distances differ from the oracle" % (name, nbad, len(pos))
    return got


@pytest.mark.parametrize("corpus", ["primitives3d", "binops3d", "unary3d", "threads3d", "scenes3d", "guards3d", "dag3d",
                                    "primitives2d", "binops2d", "unary2d", "threads2d"])
def test_flattened_programs_reproduce_the_oracle(oracle, bld, M, corpus):
    ran = 0
    for name, s in getattr(shapes, corpus)(bld):
        ran += sim_vs_oracle(oracle, bld, M, name, s) is not None
    assert ran >= 2


@pytest.mark.parametrize("dim", [3, 2])
def test_random_trees_reproduce_the_oracle(oracle, bld, M, dim):
    """The seeded random compositions of the GPU fuzz test, through the CPU model."""
    for seed in (1, 2, 3):
        for name, s in shapes.random_trees(bld, seed, 40, dim):
            sim_vs_oracle(oracle, bld, M, name, s)


def test_guards_fire_and_change_nothing(oracle, bld, M, monkeypatch):
    """Slab guards (include/gsdf_program.h): on the flange the screw subtree is skipped by whole tiles, and the guarded
    and unguarded programs give the same bits; with small tiles more guards fire, still the same bits."""
    s = gsdf.scene(bld, "npt-flange")
    pos = shapes.sample_points(s)
    st = {}
    a = sim_vs_oracle(oracle, bld, M, "flange", s, pos, stats=st)
    assert st["guards"] > 0 and st["fired"] > 0
    st_small = {}
    b = sim_vs_oracle(oracle, bld, M, "flange/small tiles", s, pos, tile=64, stats=st_small)
    assert st_small["fired"] > st["fired"]
    guarded_blob = bld.flatten(s)["blob"]
    monkeypatch.setenv("GSDF_NO_GUARDS", "1")
    assert bld.flatten(s)["blob"] != guarded_blob
    st_off = {}
    c = sim_vs_oracle(oracle, bld, M, "flange/no guards", s, pos, stats=st_off)
    assert not st_off.get("guards") and not st_off.get("fired")
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and np.array_equal(a.view(np.uint32), c.view(np.uint32))


def test_box_guards_of_the_text_scene(oracle, bld, M, monkeypatch):
    """Box guards (CULL_UB2D + BBOX_GUARD2D, what forge/textsdf scenes get): image-like tiles skip whole glyphs and glyph
    holes; every pixel still equals the oracle's brute-force union, with the guards switched off as well."""
    s = fontfix.text_scene(bld, "Ab1")
    mn, mx = s.Bounds()
    w, h = 96, 40
    xs = (mn[0] + (np.arange(w, dtype=np.float32) + np.float32(0.5)) * np.float32((mx[0] - mn[0]) / w)).astype(np.float32)
    ys = (mn[1] + (np.arange(h, dtype=np.float32) + np.float32(0.5)) * np.float32((mx[1] - mn[1]) / h)).astype(np.float32)
    # 16 x 8-pixel tiles, tile after tile (the device groups image work items into 128 x 16-pixel tiles)
    pos = np.array([[xs[tx * 16 + i], ys[ty * 8 + j]] for ty in range(h // 8) for tx in range(w // 16) for j in range(8) for i in range(16)], np.float32)
    ops = progsim.Program(bld.flatten(s)["blob"], bld.flatten(s)["aux"]).ops()
    assert progsim.OP["CULL_UB2D"] in ops and progsim.OP["BBOX_GUARD2D"] in ops
    st = {}
    a = sim_vs_oracle(oracle, bld, M, "text", s, pos, tile=128, stats=st)
    assert st["fired"] > 0
    monkeypatch.setenv("GSDF_NO_GUARDS", "1")
    ops = progsim.Program(bld.flatten(s)["blob"], bld.flatten(s)["aux"]).ops()
    assert progsim.OP["CULL_UB2D"] not in ops and progsim.OP["BBOX_GUARD2D"] not in ops
    b = sim_vs_oracle(oracle, bld, M, "text/no guards", s, pos, tile=128)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and (a < 0).any() and (a > 0).any()


def test_box_guards_never_vote_with_points_inside_the_box(oracle, bld, M):
    """A bounding box bounds an operand's value from below only OUTSIDE the box. The box guard as first shipped let points
    inside the box vote for the skip (w = -margin there), which is harmless for text (glyphs do not overlap, holes lie
    inside their glyph) but wrong for overlapping operands; the random-tree fuzz on this model found it. The overlap
    corpus passes with the corrected predicate at every tile size and fails with the legacy one."""
    for name, s in shapes.overlap2d(bld):
        pos = shapes.overlap2d_points(name, s)
        for tile in (64, progsim.TILE):
            sim_vs_oracle(oracle, bld, M, name, s, pos[:4096] if tile == 64 else pos, tile=tile)
    caught = 0
    progsim.LEGACY_BOX_GUARD = True
    try:
        for name, s in shapes.overlap2d(bld):
            pos = shapes.overlap2d_points(name, s)
            f = bld.flatten(s)
            got = progsim.run(progsim.Program(f["blob"], f["aux"]), pos, M)
            t = oracle.Tree.from_shader(s)
            want = t.eval2(pos) if s.is2d else t.eval3(pos)
            caught += int((got.view(np.uint32) != want.view(np.uint32)).any())
    finally:
        progsim.LEGACY_BOX_GUARD = False
    assert caught >= 2   # at the device's own tile size


def test_model_covers_every_opcode(oracle, bld, M):
    """Every opcode of include/gsdf_program.h is modelled, the EXT interpreter's ellipse2D / quadbezier2d included
    (their acos / cbrt / exp-log cube roots come from the oracle's math32 restatement)."""
    import re, os
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "gsdf_program.h")).read()
    enum = hdr[hdr.index("enum gsdf_opcode"):]
    names = [n for n in re.findall(r"\bGSDF_OP_([A-Z0-9_]+)\b\s*(?:=\s*0\s*)?,", enum[:enum.index("};")]) if n != "_COUNT"]
    assert names == progsim.OPS
    for name, s in shapes.primitives2d(bld):
        if "ellipse" in name or "bezier" in name:
            assert sim_vs_oracle(oracle, bld, M, name, s, shapes.sample_points(s, dense=[40, 40])) is not None


def test_radius_reuse_programs_are_bit_identical_and_gated(oracle, bld, M, monkeypatch):
    """Radius reuse (include/gsdf_program.h; the flattener's default, GSDF_RXY=0 switches it off): consumers of
    Hypot(p.x, p.y) are marked READ when the flattener proves the one-slot cache holds the radius of bit-identical x, y.
    The model checks that claim directly at every read and the result against the oracle."""
    import gsdf_b200
    from gsdf_b200 import gleval, _lib
    flange = gsdf.scene(bld, "npt-flange")
    monkeypatch.setenv("GSDF_RXY", "0")
    plain = bld.flatten(flange)["blob"]
    monkeypatch.delenv("GSDF_RXY")
    f = bld.flatten(flange)
    assert f["blob"] != plain and len(f["blob"]) == len(plain)
    P = progsim.Program(f["blob"], f["aux"])
    flags = []
    pc = 0
    while True:
        op, ln = int(P.u[pc, 0]) & 0xff, (int(P.u[pc, 0]) >> 8) & 0xff
        if progsim.OPS[op] in ("CYLINDER", "SCREW_ENTER"):
            flags.append((progsim.OPS[op], int(P.u[pc, 2 if progsim.OPS[op] == "SCREW_ENTER" else 1]) & 0x300))
        if op == 0:
            break
        pc += ln
    # pipe cylinder computes and stores; the screw, the plate cylinder (translated along z only) and the bore read
    assert flags == [("CYLINDER", progsim.RXY_WRITE), ("SCREW_ENTER", progsim.RXY_READ), ("CYLINDER", progsim.RXY_READ), ("CYLINDER", progsim.RXY_READ)]
    reads = 0
    shapes_ = shapes.all3d(bld) + shapes.all2d(bld) + shapes.dag3d(bld) + shapes.random_trees(bld, 1, 40, 3) + shapes.random_trees(bld, 1, 40, 2)
    for name, s in shapes_:
        st = {}
        sim_vs_oracle(oracle, bld, M, name, s, stats=st)
        reads += st.get("rxy_reads", 0)
    assert reads >= 5
    assert "+rxy" in gsdf_b200.version()   # the default build carries the radius slot
