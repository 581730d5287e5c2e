"""2-D image path (BASELINE config 5): colour conversions (gsdfaux/color.go, glrender/image.go:50-61) and the text
scene. CPU tests pin the oracle's conversions on hand-computed values; GPU tests compare the fused CUDA image kernel with
the oracle bit for bit / byte for byte."""
import os

import numpy as np
import pytest

import fontfix
import gsdf_b200
from gsdf_b200 import gsdf, gleval, glrender, gsdfaux, _lib


def rgba(c):
    return (c & 255, (c >> 8) & 255, (c >> 16) & 255, (c >> 24) & 255)


# ---------------------------------------------------------------------------------------------- oracle (CPU)
def test_default_conversion_known_answers(oracle):
    cc = oracle.colorconv(0)
    got = [rgba(int(c)) for c in oracle.color_of(cc, [1.0, -1.0, 0.0, np.nan, np.inf, -np.inf])]
    # image.go:50-61: NaN/Inf red, f > 0 white, default (incl. 0) black
    assert got == [(255, 255, 255, 255), (0, 0, 0, 255), (0, 0, 0, 255), (255, 0, 0, 255), (255, 0, 0, 255), (255, 0, 0, 255)]


def test_bw_linear_known_answers(oracle):
    cc = oracle.colorconv_linear_gradient(0.5, gsdfaux.Black, gsdfaux.White)  # color.go:52-54 -> blackAndWhiteLinearSmooth
    assert cc.kind == 1
    got = [rgba(int(c)) for c in oracle.color_of(cc, [-0.25, -1.0, 0.25, 1.0, 0.0, 0.125])]
    # blend = d/0.5 + 0.5: <=0 black, >=1 white, else gray uint8(blend*255): 0 -> 127, 0.125 -> uint8(0.75*255)=191
    assert got == [(0, 0, 0, 255), (0, 0, 0, 255), (255, 255, 255, 255), (255, 255, 255, 255), (127, 127, 127, 255), (191, 191, 191, 255)]
    hard = oracle.colorconv_linear_gradient(0.0, gsdfaux.Black, gsdfaux.White)  # color.go:78-80 / 97-102
    assert [rgba(int(c)) for c in oracle.color_of(hard, [-1e-9, 0.0])] == [(0, 0, 0, 255), (255, 255, 255, 255)]


def test_hsv_gradient_known_answers(oracle):
    red, blue = gsdfaux.RGBA(255, 0, 0), gsdfaux.RGBA(0, 0, 255)
    cc = oracle.colorconv_linear_gradient(2.0, red, blue)
    assert cc.kind == 3
    assert list(cc.p)[:6] == [0.0, 1.0, 1.0, np.float32(2.0 / 3), 1.0, 1.0]  # rgbToHSV (color.go:192-217)
    got = [rgba(int(c)) for c in oracle.color_of(cc, [-1.0, 1.0, 0.0])]
    assert got[0] == rgba(red) and got[1] == rgba(blue)  # blend <= 0 -> c0, >= 1 -> c1 (color.go:60-64)
    # d = 0: blend 0.5; h1-h0 = 2/3 > 0.5 so h0 += 1 (color.go:114-116): h = 0.5*1 + 0.5*(2/3) = 5/6 -> magenta (x = c)
    assert got[2] == (255, 0, 255, 255)


def test_inigo_quilez_properties(oracle):
    cc = oracle.colorconv_inigo_quilez(2.0)
    assert cc.p[0] == 0.5
    d = np.linspace(-3, 3, 601).astype(np.float32)
    col = oracle.color_of(cc, d)
    r, g, b = col & 255, (col >> 8) & 255, (col >> 16) & 255
    assert rgba(int(oracle.color_of(cc, [0.0])[0])) == (255, 255, 255, 255)  # white line on the surface (1-smoothstep = 1)
    far_out, far_in = d > 1.0, d < -1.0
    assert (r[far_out] >= b[far_out]).all() and (b[far_in] >= r[far_in]).all()  # orange outside, blue inside
    assert rgba(int(oracle.color_of(cc, [np.nan])[0])) == (255, 0, 0, 255)       # color.go:24-26


def test_png_encoder_round_trips(tmp_path):
    from PIL import Image
    img = (np.arange(5 * 7 * 4) % 251).astype(np.uint8).reshape(5, 7, 4)
    p = tmp_path / "x.png"
    p.write_bytes(gsdfaux._png_bytes(img))
    back = np.asarray(Image.open(p).convert("RGBA"))
    assert np.array_equal(back, img)


# ---------------------------------------------------------------------------------------------- CUDA (GPU)
def _convs(oracle, s):
    mn, mx = s.Bounds()
    edge = np.float32(mx[1] - mn[1]) / np.float32(1000)  # examples/image-text/text.go:66-68
    diag3 = gsdf._hypot32(mx[0] - mn[0], mx[1] - mn[1]) / np.float32(3)
    red, blue = gsdfaux.RGBA(200, 30, 10), gsdfaux.RGBA(20, 90, 250)
    return [
        ("default", None, None),
        ("bw-linear", gsdfaux.ColorConversionLinearGradient(edge, gsdfaux.Black, gsdfaux.White), oracle.colorconv_linear_gradient(edge, gsdfaux.Black, gsdfaux.White)),
        ("bw-hard", gsdfaux.ColorConversionLinearGradient(0, gsdfaux.Black, gsdfaux.White), oracle.colorconv_linear_gradient(0, gsdfaux.Black, gsdfaux.White)),
        ("inigo-quilez", gsdfaux.ColorConversionInigoQuilez(diag3), oracle.colorconv_inigo_quilez(diag3)),
        ("hsv", gsdfaux.ColorConversionLinearGradient(0.2, red, blue), oracle.colorconv_linear_gradient(0.2, red, blue)),
    ]


@pytest.mark.gpu
def test_host_colorconv_builders_match_the_oracle(oracle, bld):
    s = bld.NewCircle(1.0)
    for name, mine, theirs in _convs(oracle, s):
        if mine is None:
            continue
        assert mine.kind == theirs.kind and mine.c0 == theirs.c0 and mine.c1 == theirs.c1, name
        assert np.array_equal(np.array(mine.p, np.float32).view(np.uint32), np.array(theirs.p, np.float32).view(np.uint32)), name


@pytest.mark.gpu
@pytest.mark.parametrize("w,h", [(384, 96), (203, 51)])
def test_text_image_distances_bit_identical(oracle, bld, w, h):
    """Config 5 (forge/textsdf TextLine("Abc123~"), tolerance 0.001) at a reduced image size, through image.go's
    positions: distances bit-identical with the oracle (also for a width that is not a multiple of 4)."""
    s = fontfix.text_scene(bld)
    sdf = gleval.NewCUDASDF2(s)
    got = glrender.ImageEvaluateSDF2(sdf, w, h)
    mn, mx = s.Bounds()
    want = oracle.Tree.from_shader(s).image_eval2(mn, mx, w, h)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    if (w, h) == (96, 24):
        z = np.load(os.path.join(os.path.dirname(__file__), "golden", "text_image.npz"))
        assert np.array_equal(got.view(np.uint32), z["dist"].view(np.uint32))


@pytest.mark.gpu
def test_text_image_golden_fixture(bld):
    s = fontfix.text_scene(bld)
    got = glrender.ImageEvaluateSDF2(gleval.NewCUDASDF2(s), 96, 24)
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "text_image.npz"))
    assert np.array_equal(got.view(np.uint32), z["dist"].view(np.uint32))


@pytest.mark.gpu
def test_fused_colour_rendering_is_byte_identical(oracle, bld):
    """ImageRendererSDF2.Render with each conversion gsdfaux offers: RGBA bytes equal the oracle's, for the text scene
    and for a shape with smooth distance variation (exercises every gray level / hue)."""
    shapes2 = [("text", fontfix.text_scene(bld)), ("annulus", bld.Annulus(bld.NewCircle(1.0), 0.3))]
    for sname, s in shapes2:
        sdf = gleval.NewCUDASDF2(s)
        mn, mx = s.Bounds()
        t = oracle.Tree.from_shader(s)
        for w, h in [(256, 64), (131, 37)]:
            for name, mine, theirs in _convs(oracle, s):
                img = np.zeros((h, w, 4), np.uint8)
                glrender.NewImageRendererSDF2(4096, mine).Render(sdf, img)
                want = oracle.image_render2(t, mn, mx, w, h, theirs)
                assert np.array_equal(img, want), (sname, name, w, h, int((img != want).any(axis=2).sum()))
                assert (img[..., 3] == 255).all()


@pytest.mark.gpu
def test_render_png_file(oracle, bld, tmp_path):
    """gsdfaux.RenderPNGFile (gsdfaux.go:267-296): width from aspect ratio, default Inigo Quilez conversion."""
    from PIL import Image
    s = fontfix.text_scene(bld, "Abp8")  # forge/textsdf/glyph_test.go:14
    sdf = gleval.NewCUDASDF2(s)
    p = tmp_path / "shape.png"
    img = gsdfaux.RenderPNGFile(str(p), sdf, 128, None)
    mn, mx = s.Bounds()
    assert img.shape[0] == 128 and img.shape[1] == int(128.0 / float(np.float32(mx[1] - mn[1])) * float(np.float32(mx[0] - mn[0])))
    assert np.array_equal(np.asarray(Image.open(p).convert("RGBA")), img)
    diag3 = gsdf._hypot32(mx[0] - mn[0], mx[1] - mn[1]) / np.float32(3)
    want = oracle.image_render2(oracle.Tree.from_shader(s), mn, mx, img.shape[1], 128, oracle.colorconv_inigo_quilez(diag3))
    assert np.array_equal(img, want)


@pytest.mark.gpu
def test_image_renderer_errors(bld):
    sdf = gleval.NewCUDASDF2(bld.NewCircle(1.0))
    with pytest.raises(gsdf_b200.GsdfError, match="too small evaluation buffer size"):
        glrender.NewImageRendererSDF2(100)  # image.go:46-48
    with pytest.raises(gsdf_b200.GsdfError, match="at least of length of image rows"):
        glrender.NewImageRendererSDF2(4096).Render(sdf, np.zeros((2, 5000, 4), np.uint8))  # image.go:80-82
    sdf3 = gleval.NewCUDASDF3(bld.NewSphere(1.0))
    bad = _lib.ColorConv()
    bad.kind = 9
    with pytest.raises(gsdf_b200.GsdfError):
        glrender.NewImageRendererSDF2(4096, bad).Render(sdf, np.zeros((4, 4, 4), np.uint8))
    with pytest.raises(gsdf_b200.GsdfError):
        glrender.NewImageRendererSDF2(4096).Render(sdf3, np.zeros((4, 4, 4), np.uint8))


@pytest.mark.gpu
def test_box_guards_do_not_change_a_single_bit(oracle, bld, monkeypatch):
    """Unions of bounded 2-D shapes run with box guards (include/gsdf_program.h): tiles skip operands whose bounding box is
    farther than the running minimum. Images (2-D tiles) and point lists, guarded and unguarded, equal the oracle."""
    rng = np.random.default_rng(5)

    def blob(n, r, cx, cy, seed):
        g = np.random.default_rng(seed)
        ang = np.sort(g.uniform(0, 2 * np.pi, n))
        rad = r * (1 + 0.3 * np.sin(3 * ang + seed))
        return np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], 1).astype(np.float32)

    shapes2 = {
        "text": fontfix.text_scene(bld),
        "text-small-tol": fontfix.text_scene(bld, "p8~A", tol=0.01),
        "blobs": bld.Union2D(*[bld.NewPolygon(blob(9 + k, 0.4, 1.3 * (k % 4), 1.1 * (k // 4), k)) for k in range(10)]),
        "mixed": bld.Union2D(bld.NewCircle(0.5), bld.Translate2D(bld.NewRectangle(1.0, 0.4), 2.0, 0.3),
                             bld.Translate2D(bld.Difference2D(bld.NewCircle(0.6), bld.NewCircle(0.3)), 4.0, -0.2),
                             bld.Translate2D(bld.NewHexagon(0.4), 1.0, 1.5),   # not box-bounded: never guarded, still correct
                             bld.Translate2D(bld.NewPolygon(blob(12, 0.5, 0, 0, 3)), 3.0, 1.4)),
    }
    for name, s in shapes2.items():
        t = oracle.Tree.from_shader(s)
        mn, mx = s.Bounds()
        for guards in (True, False):
            if guards:
                monkeypatch.delenv("GSDF_NO_GUARDS", raising=False)
            else:
                monkeypatch.setenv("GSDF_NO_GUARDS", "1")
            sdf = gleval.NewCUDASDF2(s)
            for w, h in [(640, 200), (131, 37), (64, 1024)]:
                got = glrender.ImageEvaluateSDF2(sdf, w, h)
                want = t.image_eval2(mn, mx, w, h)
                assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (name, guards, w, h, int((got.view(np.uint32) != want.view(np.uint32)).sum()))
            # point lists: spatially sorted (guards fire) and shuffled (they rarely do)
            pos = (mn - 0.5 * (mx - mn) + rng.random((30000, 2), dtype=np.float32) * 2.0 * (mx - mn)).astype(np.float32)
            for arr in (pos[np.lexsort((pos[:, 0], pos[:, 1]))], pos):
                arr = np.ascontiguousarray(arr)
                out = np.empty(len(arr), np.float32)
                sdf.Evaluate(arr, out)
                assert np.array_equal(out.view(np.uint32), t.eval2(arr).view(np.uint32)), (name, guards)
