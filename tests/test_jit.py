"""Run-time specialised kernels (gsdf_program_specialize, gsdf_b200/csrc/jit.cu): the reference compiles a GLSL shader per
tree (gleval/gpu.go:35-54); here the tree's instruction stream is compiled into straight-line code around the interpreter's own
opcode bodies. The compile step runs without a GPU (NVRTC); on the GPU the specialised renders must equal the oracle -- and the
interpreter -- bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

from gsdf_b200 import _lib, gsdf

import shapes


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def jit_compile(bld, s):
    f = bld.flatten(s)
    aux = np.ascontiguousarray(f["aux"], dtype=np.float32)
    return int(_lib.lib.gsdf_jit_compile(f["blob"], len(f["blob"]), aux.ctypes.data_as(C.POINTER(C.c_float)), aux.size))


def nvrtc_present():
    try:
        C.CDLL("libnvrtc.so.12")
        return True
    except OSError:
        try:
            C.CDLL("libnvrtc.so")
            return True
        except OSError:
            return False


def test_compile_step_without_a_device(bld):
    """The generated source compiles to CUBINs for sm_100a here, without a GPU (the headers NVRTC needs are embedded in the
    library); a tree with a slab guard (forward goto) and a polygon (side buffer) is among them. Too long a program and a 2-D
    program are refused with GSDF_EUNSUPPORTED, not an error."""
    if not nvrtc_present():
        pytest.skip("libnvrtc is not installed")
    T = gsdf.threads
    small = bld.Union(bld.NewSphere(1.0), bld.Translate(bld.NewBox(1, 0.5, 0.8, 0.1), 0.5, 0, 0))
    guarded = bld.Difference(bld.NewCylinder(2.0, 3.0, 0.1), T.Screw(bld, 0.43, T.NPT(0.5)))
    for s in (small, guarded):
        n = jit_compile(bld, s)
        assert n > 10000, _lib.last_error()
    long_tree = bld.Union(*[bld.Translate(bld.NewSphere(0.1), 0.3 * i, 0, 0) for i in range(120)])
    assert jit_compile(bld, long_tree) == _lib.EUNSUPPORTED and "too long" in _lib.last_error()
    assert jit_compile(bld, bld.NewCircle(1.0)) < 0   # 2-D programs: the image path keeps the interpreter


@pytest.mark.gpu
@pytest.mark.parametrize("scene,resdiv", [("sphere", 70), ("npt-flange", 150), ("bolt", 160)])
def test_specialised_render_is_bit_identical(oracle, bld, scene, resdiv):
    """Octree and Flat renders through the specialised kernels (lattice evaluation with four corners and with one corner per
    thread, prune-centre pass): lattice values, cube cases and triangles equal the oracle bit for bit, before and after the
    steady-state graph capture."""
    from gsdf_b200 import gleval, glrender
    from test_gpu_parity import oracle_mesh
    s = bld.NewSphere(1.0) if scene == "sphere" else gsdf.scene(bld, scene)
    res = np.float32(s.Diagonal() / np.float32(resdiv))
    sdf = gleval.NewCUDASDF3(s)
    if not sdf.Specialize():
        pytest.skip("run-time compilation is not available on this box: " + _lib.last_error())
    assert sdf.Specialized()
    for cls in (glrender.Octree, glrender.FlatRenderer):
        R = cls(sdf, res, keep_cases=True, keep_grid=True)
        lat, grid, mask, wt, wc, _ = oracle_mesh(oracle, s, res, R.Plan())
        assert int((R.Cases() != wc).sum()) == 0
        assert np.array_equal(bits(R.AllTriangles()), bits(wt))
        g = R.Grid()
        ev = bits(g) != np.uint32(0x7f7f7f7f)
        assert np.array_equal(bits(g)[ev], bits(grid)[ev])
        R.Close()
    # steady state (graph replays) and a thin slab, whose lattice evaluation switches to one corner per thread
    nz = glrender.lattice_from_bounds(*s.Bounds(), res).n[2]
    for cz in (None, (nz // 2, min(nz, nz // 2 + 4))):
        R = glrender.Octree(sdf, res, cz_range=cz)
        lat, grid, mask, wt, _, _ = oracle_mesh(oracle, s, res, R.Plan())
        if cz is not None:
            wt, _ = oracle.flat_march(lat, grid, blockmask=mask, cz_range=cz)
        for _ in range(4):
            R.Rerun()
            assert np.array_equal(bits(R.AllTriangles()), bits(wt))
        R.Close()


@pytest.mark.gpu
def test_specialisation_follows_the_structure_not_the_parameters(oracle, bld):
    """gsdf_program_update with new parameters of the same tree keeps the compiled kernels (operands are read from the
    program); an update that changes the structure drops them, and the interpreter takes over until Specialize is called again.
    Results equal the oracle in every state."""
    from gsdf_b200 import gleval, glrender
    from test_gpu_parity import oracle_mesh

    def tree(r, k):
        return bld.SmoothUnion(k, bld.NewSphere(r), bld.Translate(bld.NewBox(1.0, 0.6, 0.8, 0.05), 0.7, 0.1, 0))

    s = tree(1.0, 0.2)
    sdf = gleval.NewCUDASDF3(s)
    if not sdf.Specialize():
        pytest.skip("run-time compilation is not available on this box")
    res = np.float32(s.Diagonal() / np.float32(60))
    for shader, want_special in ((s, True), (tree(0.8, 0.3), True), (bld.Union(bld.NewSphere(1.0), bld.NewTorus(1.2, 0.3)), False)):
        if shader is not s:
            sdf.Update(shader)
        assert sdf.Specialized() == want_special
        R = glrender.Octree(sdf, np.float32(shader.Diagonal() / np.float32(60)))
        _, _, _, wt, _, _ = oracle_mesh(oracle, shader, np.float32(shader.Diagonal() / np.float32(60)), R.Plan())
        assert np.array_equal(bits(R.AllTriangles()), bits(wt))
        R.Close()
    assert sdf.Specialize() and sdf.Specialized()   # the new structure compiles as well


@pytest.mark.gpu
def test_specialised_random_trees_and_multi_slab(oracle, bld):
    """Seeded random trees (guards, arrays, smooth blends, 2-D sections) through the specialised kernels against the oracle, and
    the multi-slab driver with gsdf_multi_specialize."""
    from gsdf_b200 import gleval, glrender
    from test_gpu_parity import oracle_mesh
    done = 0
    for name, s in shapes.random_trees(bld, 4242, 6, dim=3, depth=3):
        sdf = gleval.NewCUDASDF3(s)
        if not sdf.Specialize():
            if _lib.last_error().startswith("program of"):
                continue   # too long: stays with the interpreter
            pytest.skip("run-time compilation is not available on this box")
        res = np.float32(s.Diagonal() / np.float32(40))
        R = glrender.Octree(sdf, res, prune="literal")
        _, _, _, wt, _, _ = oracle_mesh(oracle, s, res, R.Plan())
        assert np.array_equal(bits(R.AllTriangles()), bits(wt)), name
        R.Close()
        done += 1
    assert done >= 3
    s = gsdf.scene(bld, "npt-flange")
    res = np.float32(s.Diagonal() / np.float32(120))
    M = glrender.MultiRenderer(s, res, devices=[0], slabs_per_device=3)
    assert M.Specialize()
    ref = glrender.Octree(gleval.NewCUDASDF3(s), res).AllTriangles()
    host = glrender.pinned_empty((len(ref) + 8, 3, 3))
    for _ in range(3):
        n = M.RenderToHost(host)
        assert n == len(ref) and np.array_equal(bits(host[:n]), bits(ref))
    M.Close()
