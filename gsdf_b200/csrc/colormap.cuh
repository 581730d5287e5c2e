// colormap.cuh -- the float32 -> colour conversions of the reference's 2-D image path, as device code.
//
// ImageRendererSDF2 calls a Go closure per pixel (glrender/image.go:112-116); the closures gsdfaux offers
// (gsdfaux/color.go) and the default of NewImageRendererSDF2 (image.go:50-61) are restated here so that the kernel that
// evaluated the distance also writes the RGBA8 pixel. Operation order follows the Go code; every operation rounds to
// float32 individually (-fmad=false). ms1.SmoothStep / ms1.Interp / ms3.InterpElem come from the un-vendored
// github.com/soypat/geometry module and are restated from their documented meaning (GLSL smoothstep / mix).
#pragma once
#include <stdint.h>

#include "../../include/gsdf_b200.h"
#include "math32.cuh"

namespace gsdfk {

struct ColorConv {
    int32_t kind;
    float p[7];
    uint32_t c0, c1;
};

M32_HD uint32_t pack_rgba(uint32_t r, uint32_t g, uint32_t b, uint32_t a) { return (r & 255u) | (g & 255u) << 8 | (b & 255u) << 16 | (a & 255u) << 24; }
// Go's uint8(f) / uint32(f) on amd64 for in-range values: truncate toward zero (CVTTSS2SL), then keep the low bits.
M32_HD uint32_t trunc_u(float f) { return (uint32_t)(int32_t)f; }

// gsdfaux/color.go:165-188
M32_HD void hsv_to_rgb(float h, float s, float v, float &r, float &g, float &b) {
    const float c = s * v;
    const float x = c * (1.f - fabsf(fmodf(h * 6.f, 2.f) - 1.f));
    const float m = v - c;
    r = g = b = 0.f;
    if (h >= 0.f && h <= (float)(1.0 / 6)) { r = c; g = x; b = 0.f; }
    else if (h > (float)(1.0 / 6) && h <= (float)(2.0 / 6)) { r = x; g = c; b = 0.f; }
    else if (h > (float)(2.0 / 6) && h <= (float)(3.0 / 6)) { r = 0.f; g = c; b = x; }
    else if (h > (float)(3.0 / 6) && h <= (float)(4.0 / 6)) { r = 0.f; g = x; b = c; }
    else if (h > (float)(4.0 / 6) && h <= (float)(5.0 / 6)) { r = x; g = 0.f; b = c; }
    else if (h > (float)(5.0 / 6) && h <= 1.0f) { r = c; g = 0.f; b = x; }
    r += m; g += m; b += m;
}

M32_HD uint32_t color_of(const ColorConv &cc, float d) {
    const uint32_t black = 0xff000000u, white = 0xffffffffu, red = 0xff0000ffu;
    switch (cc.kind) {
    case GSDF_CONV_BW_LINEAR: {  // color.go:77-102
        const float edge = cc.p[0];
        if (edge == 0.f) return d < 0.f ? black : white;
        float blend = d / edge + 0.5f;
        if (blend <= 0.f) return black;
        if (blend >= 1.f) return white;
        blend = m32::clampf(blend, 0.f, 1.f);
        const uint32_t y = trunc_u(blend * 255.f);
        return pack_rgba(y, y, y, 255u);
    }
    case GSDF_CONV_INIGO_QUILEZ: {  // color.go:21-47
        if (d != d) return red;
        d *= cc.p[0];
        float cx, cy, cz;
        if (d > 0.f) { cx = 0.9f; cy = 0.6f; cz = 0.3f; } else { cx = 0.65f; cy = 0.85f; cz = 1.0f; }
        float f = 1.f - m32::exp32(-6.f * fabsf(d));
        cx *= f; cy *= f; cz *= f;
        f = 0.8f + 0.2f * m32::cos(150.f * d);
        cx *= f; cy *= f; cz *= f;
        float t = m32::clampf((fabsf(d) - 0.f) / (0.01f - 0.f), 0.f, 1.f);  // ms1.SmoothStep(0, 0.01, |d|)
        t = t * t * (3.f - 2.f * t);
        const float mx = 1.f - t;
        cx = cx * (1.f - mx) + 1.f * mx;  // ms3.InterpElem(c, one, max)
        cy = cy * (1.f - mx) + 1.f * mx;
        cz = cz * (1.f - mx) + 1.f * mx;
        return pack_rgba(trunc_u(cx * 255.f), trunc_u(cy * 255.f), trunc_u(cz * 255.f), 255u);
    }
    case GSDF_CONV_HSV_GRADIENT: {  // color.go:57-72
        const float blend = d / cc.p[6] + 0.5f;
        if (blend <= 0.f) return cc.c0;
        if (blend >= 1.f) return cc.c1;
        float h0 = cc.p[0], h1 = cc.p[3];
        if (h1 - h0 > 0.5f) h0 += 1.0f;           // interpHSV, color.go:113-124
        else if (h1 - h0 < -0.5f) h1 += 1.0f;
        const float h = h0 * (1.f - blend) + h1 * blend;
        const float s = cc.p[1] * (1.f - blend) + cc.p[4] * blend;
        const float v = cc.p[2] * (1.f - blend) + cc.p[5] * blend;
        float r, g, b;
        hsv_to_rgb(h, s, v, r, g, b);
        // rgbToC (color.go:156-160) then the byte split of color.go:71
        return pack_rgba(trunc_u(m32::clampf(r, 0.f, 1.f) * 255.f), trunc_u(m32::clampf(g, 0.f, 1.f) * 255.f), trunc_u(m32::clampf(b, 0.f, 1.f) * 255.f), 255u);
    }
    default:  // image.go:50-61
        if (d != d || isinf(d)) return red;
        return d > 0.f ? white : black;
    }
}

}  // namespace gsdfk
