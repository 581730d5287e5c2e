"""ctypes binding of libgsdfb200.so (include/gsdf_b200.h: CUDA kernels + C ABI) and libgsdfhost.so (include/gsdf_host.h:
the C++ host layer -- builder, forge/threads, textsdf, flattener; no CUDA dependency).

The libraries are the product: there is no Python or CPU fallback. If a shared object is missing the import fails
loudly and tells the caller how to build it. `lib` resolves a symbol from whichever library exports it.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GSDF_B200_LIB") or os.path.join(_HERE, "libgsdfb200.so")  # env override: kernel A/B experiments
HOST_LIB_PATH = os.path.join(_HERE, "libgsdfhost.so")


class GsdfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("gsdf_b200 error %d: %s" % (code, msg))
        self.code = code


# gsdf_status (include/gsdf_b200.h)
OK, EINVAL, ELEN, EEMPTY, ECUDA, ENOMEM, EPROGRAM, ESHORT, ERES, EUNSUPPORTED = 0, -1, -2, -3, -4, -5, -6, -7, -8, -9
MESH_PRUNE, MESH_KEEP_CASES, MESH_KEEP_GRID, MESH_STAGE_TIMING, MESH_PRUNE_LITERAL = 1, 2, 4, 8, 16
PRUNE_MARGIN_DEFAULT, PRUNE_MAX_LEVELS = 1.25, 4
DC_NAIVE, DC_LEAST_SQUARES, DC_LEAST_SQUARES_CHISELED = 0, 1, 2


class Lattice(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("res", C.c_float), ("n", C.c_int32 * 3)]


class PrunePlan(C.Structure):  # gsdf_prune_plan
    _fields_ = [("nlevels", C.c_int32), ("level", C.c_int32 * 4), ("margin", C.c_float * 4)]

    @classmethod
    def make(cls, levels, margins):
        p = cls()
        p.nlevels = len(levels)
        for i, (l, m) in enumerate(zip(levels, margins)):
            p.level[i], p.margin[i] = int(l), float(m)
        return p

    def levels(self):
        return [(int(self.level[i]), float(self.margin[i])) for i in range(self.nlevels)]


class ColorConv(C.Structure):  # gsdf_colorconv
    _fields_ = [("kind", C.c_int32), ("p", C.c_float * 7), ("c0", C.c_uint32), ("c1", C.c_uint32)]


class TreeNode(C.Structure):  # gsdf_tree_node, 96 bytes
    _fields_ = [("kind", C.c_int32), ("nchild", C.c_int32), ("child_off", C.c_int32), ("aux_off", C.c_int32),
                ("aux_cnt", C.c_int32), ("iparam", C.c_int32 * 3), ("fparam", C.c_float * 16)]


class _Libs:
    """Symbol lookup over the two shared objects (device library first)."""

    def __init__(self, libs):
        self._libs = libs

    def __getattr__(self, name):
        for l in self._libs:
            try:
                fn = getattr(l, name)
            except AttributeError:
                continue
            setattr(self, name, fn)
            return fn
        raise AttributeError(name)


def _load(host_only=False):
    paths = [HOST_LIB_PATH] if host_only else [LIB_PATH, HOST_LIB_PATH]
    for p in paths:
        if not os.path.exists(p):
            raise ImportError(
                "gsdf_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C gsdf_b200/csrc` (needs nvcc, sm_100a). There is no CPU fallback." % p)
    lib = _Libs([C.CDLL(p) for p in paths])
    vp, f32p, i32p, u8p, u64p = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)
    sig = {
        # include/gsdf_b200.h
        "gsdf_version": (C.c_char_p, []),
        "gsdf_last_error": (C.c_char_p, []),
        "gsdf_device_count": (C.c_int, []),
        "gsdf_set_device": (C.c_int, [C.c_int]),
        "gsdf_host_alloc": (vp, [C.c_size_t]),
        "gsdf_host_free": (None, [vp]),
        "gsdf_program_create": (C.c_int, [vp, C.c_size_t, f32p, C.c_size_t, C.POINTER(vp)]),
        "gsdf_program_create_on": (C.c_int, [C.c_int, vp, C.c_size_t, f32p, C.c_size_t, C.POINTER(vp)]),
        "gsdf_program_update": (C.c_int, [vp, vp, C.c_size_t, f32p, C.c_size_t]),
        "gsdf_program_specialize": (C.c_int, [vp]),
        "gsdf_program_is_specialized": (C.c_int, [vp]),
        "gsdf_jit_compile": (C.c_int64, [vp, C.c_size_t, f32p, C.c_size_t]),
        "gsdf_program_device_image": (C.c_int64, [vp, C.c_size_t, f32p, C.c_size_t, vp, C.c_size_t]),
        "gsdf_program_destroy": (None, [vp]),
        "gsdf_program_evaluations": (C.c_uint64, [vp]),
        "gsdf_eval3": (C.c_int, [vp, vp, vp, C.c_size_t]),
        "gsdf_eval2": (C.c_int, [vp, vp, vp, C.c_size_t]),
        "gsdf_eval3_device": (C.c_int, [vp, vp, vp, C.c_size_t, vp]),
        "gsdf_eval2_device": (C.c_int, [vp, vp, vp, C.c_size_t, vp]),
        "gsdf_lattice_from_bounds": (C.c_int, [f32p, f32p, C.c_float, C.POINTER(Lattice)]),
        "gsdf_octree_levels": (C.c_int, [f32p, f32p, C.c_float]),
        "gsdf_grid_eval": (C.c_int, [vp, C.POINTER(Lattice), C.c_int, C.c_int, vp]),
        "gsdf_grid_eval_device": (C.c_int, [vp, C.POINTER(Lattice), C.c_int, C.c_int, vp, vp]),
        "gsdf_mesh_begin": (C.c_int, [vp, C.POINTER(Lattice), C.c_int, C.c_int, C.c_uint, C.POINTER(vp)]),
        "gsdf_mesh_begin_plan": (C.c_int, [vp, C.POINTER(Lattice), C.c_int, C.c_int, C.c_uint, C.POINTER(PrunePlan), C.POINTER(vp)]),
        "gsdf_prune_plan_default": (C.c_int, [C.POINTER(Lattice), C.c_uint, C.POINTER(PrunePlan)]),
        "gsdf_multi_begin": (C.c_int, [C.c_int, i32p, C.c_int, vp, C.c_size_t, f32p, C.c_size_t, C.POINTER(Lattice), C.c_uint, C.POINTER(vp)]),
        "gsdf_multi_update": (C.c_int, [vp, vp, C.c_size_t, f32p, C.c_size_t]),
        "gsdf_multi_render": (C.c_int64, [vp, vp, C.c_size_t]),
        "gsdf_multi_read": (C.c_int64, [vp, vp, C.c_size_t]),
        "gsdf_multi_rewind": (C.c_int, [vp]),
        "gsdf_multi_timeline": (C.c_int, [vp, C.POINTER(C.c_double), C.c_int]),
        "gsdf_multi_specialize": (C.c_int, [vp]),
        "gsdf_multi_stats": (C.c_int, [vp, u64p, u64p, u64p, f32p]),
        "gsdf_multi_slabs": (C.c_int, [vp, i32p, i32p, C.c_int]),
        "gsdf_multi_stl": (C.c_int64, [vp, vp, C.c_size_t]),
        "gsdf_multi_destroy": (None, [vp]),
        "gsdf_slab_cuts": (C.c_int, [C.c_int, C.c_int, i32p]),
        "gsdf_slab_rebalance": (C.c_int, [C.c_int, C.c_int, i32p, C.POINTER(C.c_double), i32p]),
        "gsdf_multi_rebalance": (C.c_int, [vp, C.c_int]),
        "gsdf_mesh_rerun": (C.c_int, [vp]),
        "gsdf_mesh_rerun_begin": (C.c_int, [vp]),
        "gsdf_mesh_rerun_end": (C.c_int, [vp]),
        "gsdf_mesh_read_prefix_async": (C.c_int64, [vp, vp, C.c_size_t]),
        "gsdf_mesh_set_program": (C.c_int, [vp, vp]),
        "gsdf_mesh_read": (C.c_int64, [vp, vp, C.c_size_t]),
        "gsdf_mesh_read_async": (C.c_int64, [vp, vp, C.c_size_t]),
        "gsdf_mesh_wait": (C.c_int, [vp]),
        "gsdf_mesh_device_triangles": (C.c_int, [vp, C.POINTER(vp), u64p]),
        "gsdf_mesh_stats": (C.c_int, [vp, u64p, u64p, u64p]),
        "gsdf_mesh_cases": (C.c_int, [vp, vp, C.c_size_t]),
        "gsdf_mesh_grid": (C.c_int, [vp, vp, C.c_size_t]),
        "gsdf_mesh_timings": (C.c_int, [vp, f32p]),
        "gsdf_mesh_destroy": (None, [vp]),
        "gsdf_stl_pack": (C.c_int64, [vp, C.c_size_t, vp, C.c_size_t]),
        "gsdf_mesh_stl": (C.c_int64, [vp, vp, C.c_size_t]),
        "gsdf_image_eval2": (C.c_int, [vp, f32p, f32p, C.c_int, C.c_int, vp]),
        "gsdf_colorconv_inigo_quilez": (C.c_int, [C.c_float, C.POINTER(ColorConv)]),
        "gsdf_colorconv_linear_gradient": (C.c_int, [C.c_float, C.c_uint32, C.c_uint32, C.POINTER(ColorConv)]),
        "gsdf_image_render2": (C.c_int, [vp, f32p, f32p, C.c_int, C.c_int, C.POINTER(ColorConv), vp]),
        "gsdf_image_eval2_device": (C.c_int, [vp, f32p, f32p, C.c_int, C.c_int, vp, vp]),
        "gsdf_image_render2_device": (C.c_int, [vp, f32p, f32p, C.c_int, C.c_int, C.POINTER(ColorConv), vp, vp]),
        "gsdf_dc_levels": (C.c_int, [f32p, f32p, C.c_float, f32p]),
        "gsdf_dc_begin": (C.c_int, [vp, f32p, f32p, C.c_float, C.c_int, C.POINTER(vp)]),
        "gsdf_dc_begin_part": (C.c_int, [vp, f32p, f32p, C.c_float, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
        "gsdf_dc_part_region": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint32), i32p]),
        "gsdf_dc_rerun": (C.c_int, [vp]),
        "gsdf_dc_read": (C.c_int64, [vp, vp, C.c_size_t]),
        "gsdf_dc_device_triangles": (C.c_int, [vp, C.POINTER(vp), u64p]),
        "gsdf_dc_stats": (C.c_int, [vp, u64p]),
        "gsdf_dc_destroy": (None, [vp]),
        # include/gsdf_host.h
        "gsdfh_builder_new": (vp, []),
        "gsdfh_builder_free": (None, [vp]),
        "gsdfh_builder_err": (C.c_char_p, [vp]),
        "gsdfh_builder_clear_errors": (None, [vp]),
        "gsdfh_node": (C.c_int32, [vp, C.c_int32, f32p, C.c_int, i32p, C.c_int, i32p, C.c_int, f32p, C.c_int]),
        "gsdfh_is2d": (C.c_int, [vp, C.c_int32]),
        "gsdfh_bounds3": (C.c_int, [vp, C.c_int32, f32p]),
        "gsdfh_bounds2": (C.c_int, [vp, C.c_int32, f32p]),
        "gsdfh_thread_profile": (C.c_int32, [vp, C.c_int, C.c_float, C.c_float, C.c_int]),
        "gsdfh_screw": (C.c_int32, [vp, C.c_float, C.c_int, C.c_float, C.c_float, C.c_int]),
        "gsdfh_nut": (C.c_int32, [vp, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_float]),
        "gsdfh_bolt": (C.c_int32, [vp, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]),
        "gsdfh_hexhead": (C.c_int32, [vp, C.c_float, C.c_float, C.c_int, C.c_int]),
        "gsdfh_scene": (C.c_int32, [vp, C.c_char_p, C.c_float]),
        "gsdfh_font_new": (vp, []),
        "gsdfh_font_free": (None, [vp]),
        "gsdfh_font_err": (C.c_char_p, [vp]),
        "gsdfh_font_configure": (C.c_int, [vp, C.c_float]),
        "gsdfh_font_load_ttf": (C.c_int, [vp, vp, C.c_size_t]),
        "gsdfh_font_textline": (C.c_int32, [vp, vp, C.c_char_p]),
        "gsdfh_font_glyph": (C.c_int32, [vp, vp, C.c_uint32]),
        "gsdfh_font_kern": (C.c_float, [vp, C.c_uint32, C.c_uint32]),
        "gsdfh_font_advance_width": (C.c_float, [vp, C.c_uint32]),
        "gsdfh_font_scaleout": (C.c_float, [vp]),
        "gsdfh_font_glyph_index": (C.c_int32, [vp, C.c_uint32]),
        "gsdfh_font_glyph_segments": (C.c_int32, [vp, C.c_int32, i32p, C.c_int32]),
        "gsdfh_font_info": (C.c_int, [vp, i32p]),
        "gsdfh_tree": (C.c_int, [vp, C.POINTER(C.POINTER(TreeNode)), i32p, C.POINTER(i32p), i32p, C.POINTER(f32p), i32p]),
        "gsdfh_flatten": (vp, [vp, C.c_int32]),
        "gsdfh_flat_blob": (vp, [vp, C.POINTER(C.c_size_t)]),
        "gsdfh_flat_aux": (f32p, [vp, C.POINTER(C.c_size_t)]),
        "gsdfh_flat_info": (None, [vp, i32p]),
        "gsdfh_flat_free": (None, [vp]),
    }
    for name, (res, args) in sig.items():
        if host_only and not name.startswith("gsdfh_"):
            continue
        fn = getattr(lib, name)  # AttributeError here = header/library drift; tests check every symbol
        fn.restype = res
        fn.argtypes = args
    lib._gsdf_signatures = sig
    return lib


# GSDF_HOST_ONLY=1 (bench.py --impl reference): only the host layer is mapped -- scene builders for the CPU oracle, no CUDA
HOST_ONLY = os.environ.get("GSDF_HOST_ONLY") == "1"
lib = _load(host_only=HOST_ONLY)


def last_error():
    if HOST_ONLY:
        return "host-only mode"
    return lib.gsdf_last_error().decode("utf-8", "replace")


def check(rc):
    if rc < 0:
        raise GsdfError(rc, last_error())
    return rc
