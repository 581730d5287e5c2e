// flatten.cpp -- tree table -> packed node program (include/gsdf_program.h).
//
// Walks the tree in ForEachChild / ForEach2DChild order (glbuild/glbuild.go:63-89) and emits a postfix stream.
// Position liveness is resolved here so the kernels never save a position nobody reads again:
//   emit(node, restore): after the emitted code runs, D holds one more value; p equals its value on entry iff
//   `restore` was requested (the caller is about to evaluate a later sibling at the same p).
// All per-node constants are derived in float32 exactly as the node's Evaluate method derives them per call.
#include "flatten.h"

#include <cstdlib>
#include <cstring>

#include "../math32.cuh"

namespace gsdfhost {

namespace {

constexpr double kSqrt3 = 1.7320508075688772935274463415058723669428052538103806280558069794;

inline uint32_t fbits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

// Slab guard handed down from a combiner to a later child (include/gsdf_program.h "slab guards").
struct Guard {
    uint32_t kind = GSDF_GUARD_NONE;
    float k = 0;
};

struct Emitter {
    const Builder &b;
    Program &out;
    int d = 0, p = 0, dmax = 0, pmax = 0;
    bool guards = true;
    std::string err;

    Emitter(const Builder &bb, Program &o) : b(bb), out(o) {
        const char *e = std::getenv("GSDF_NO_GUARDS");
        guards = !(e && *e && *e != '0');
    }

    // Nodes whose Evaluate forwards the child's distance unchanged (they only move p).
    static bool transparent(int kind) {
        return kind == GSDF_N_TRANSLATE || kind == GSDF_N_TRANSFORM || kind == GSDF_N_SYMMETRY || kind == GSDF_N_TWIST ||
               kind == GSDF_N_BOUNDS3;
    }
    // True when `id` is a screw/extrude reached through transparent nodes only: its value is >= |z|-h/2 in its own
    // frame, which the ENTER op has in hand before any of the expensive 2-D work.
    bool guardable(NodeId id) const {
        while (b.valid(id) && transparent(b.node(id).kind) && b.node(id).nchild == 1) id = b.child(b.node(id), 0);
        return b.valid(id) && (b.node(id).kind == GSDF_N_SCREW || b.node(id).kind == GSDF_N_EXTRUDE);
    }
    Guard guardFor(NodeId child, uint32_t kind, float k = 0) const {
        Guard g;
        if (guards && guardable(child)) { g.kind = kind; g.k = k; }
        return g;
    }
    // ---- box guards (include/gsdf_program.h): 2-D operands whose value is >= dist(p, Bounds()) outside their box
    bool boxBounded(NodeId id) const {
        if (!b.valid(id)) return false;
        const gsdf_tree_node &n = b.node(id);
        switch (n.kind) {
        case GSDF_N_POLY2D: case GSDF_N_CIRCLE2D: case GSDF_N_RECT2D: return true;
        case GSDF_N_TRANSLATE2D: return boxBounded(b.child(n, 0));
        case GSDF_N_DIFF2D: return boxBounded(b.child(n, 0));  // max(a, -b) >= a
        case GSDF_N_UNION2D:
            for (int k = 0; k < n.nchild; k++) if (!boxBounded(b.child(n, k))) return false;
            return n.nchild > 0;
        default: return false;
        }
    }
    // Points ON the outline of the operand's zero set, in the operand's own frame: the operand's value at p is <= |p - v|.
    void anchorsOf(NodeId id, std::vector<Vec2> &out) const {
        if (!b.valid(id)) return;
        const gsdf_tree_node &n = b.node(id);
        const float *f = n.fparam;
        switch (n.kind) {
        case GSDF_N_POLY2D: {
            const float *v = &b.aux()[n.aux_off];
            for (int i = 0; i < n.aux_cnt / 2; i++) out.push_back({v[2 * i], v[2 * i + 1]});
            return;
        }
        case GSDF_N_CIRCLE2D: out.insert(out.end(), {{f[0], 0}, {-f[0], 0}, {0, f[0]}, {0, -f[0]}}); return;
        case GSDF_N_RECT2D: out.insert(out.end(), {{0.5f * f[0], 0.5f * f[1]}, {-0.5f * f[0], 0.5f * f[1]}, {0.5f * f[0], -0.5f * f[1]}, {-0.5f * f[0], -0.5f * f[1]}}); return;
        case GSDF_N_TRANSLATE2D: {
            size_t first = out.size();
            anchorsOf(b.child(n, 0), out);
            for (size_t i = first; i < out.size(); i++) { out[i].x += f[0]; out[i].y += f[1]; }
            return;
        }
        case GSDF_N_UNION2D:  // min(a, b) <= each operand
            for (int k = 0; k < n.nchild; k++) anchorsOf(b.child(n, k), out);
            return;
        case GSDF_N_DIFF2D: {  // max(a, -b) <= |p - v| for v on a's outline and outside b: keep a's anchors clear of b's box
            // ... which is only sound when b's Bounds() really encloses b: an OverloadShader2DBounds wrapper or a primitive with
            // a loose / cosmetic box could leave an anchor INSIDE the hole, where max(a, -b) > 0 and the upper bound would be
            // too small. Same whitelist as the guards themselves; otherwise this operand contributes no anchors.
            if (!boxBounded(b.child(n, 1))) return;
            std::vector<Vec2> a;
            anchorsOf(b.child(n, 0), a);
            const Box2 bb = b.Bounds2(b.child(n, 1));
            for (Vec2 v : a) if (v.x < bb.min.x || v.x > bb.max.x || v.y < bb.min.y || v.y > bb.max.y) out.push_back(v);
            return;
        }
        default: return;
        }
    }
    static float boxScale(const Box2 &bb) {
        return 1e-5f * std::max(std::max(std::fabs(bb.min.x), std::fabs(bb.max.x)), std::max(std::fabs(bb.min.y), std::fabs(bb.max.y)));
    }
    // Emits BBOX_GUARD2D for the operand `child` (box in the current frame); returns the header word to patch, or 0.
    size_t boxGuard(NodeId child, uint32_t kind, float margin) {
        const Box2 bb = b.Bounds2(child);
        size_t hw = out.chunks.size();
        header(GSDF_OP_BBOX_GUARD2D, 2, kind, 0, fbits(margin));
        chunk(bb.min.x, bb.min.y, bb.max.x, bb.max.y);
        return hw;
    }
    void patchBoxGuard(size_t hw, uint32_t kind) { out.chunks[hw + 1] = kind | ((uint32_t)(out.chunks.size() / 4) << 8); }

    // The skip lands right behind the guarded node's own exit op (the restoring POP_POS ops of its wrappers follow).
    void patchGuard(size_t headerWord, const Guard &g) {
        if (g.kind != GSDF_GUARD_NONE) out.chunks[headerWord + 1] = g.kind | ((uint32_t)(out.chunks.size() / 4) << 8);
    }

    void header(uint32_t op, uint32_t nchunks, uint32_t w1 = 0, uint32_t w2 = 0, uint32_t w3 = 0) {
        out.chunks.insert(out.chunks.end(), {op | (nchunks << 8), w1, w2, w3});
        out.ninstr++;
    }
    void chunk(float a, float b_ = 0, float c = 0, float dd = 0) {
        out.chunks.insert(out.chunks.end(), {fbits(a), fbits(b_), fbits(c), fbits(dd)});
    }
    void op0(uint32_t op) { header(op, 1); }
    void opf(uint32_t op, float f2, float f3 = 0) { header(op, 1, 0, fbits(f2), fbits(f3)); }
    void pushD() { if (++d > dmax) dmax = d; }
    void popD() { --d; }
    void pushP() { op0(GSDF_OP_PUSH_POS); if (++p > pmax) pmax = p; }
    void popP() { op0(GSDF_OP_POP_POS); --p; }

    // Emits `enter`, the child, `exit` for a unary position transform.
    template <class Enter, class Exit>
    bool unaryPos(const gsdf_tree_node &n, bool restore, bool child2d, Enter enter, Exit exit, Guard g = Guard{}) {
        if (restore) pushP();
        enter();
        if (!emit(b.child(n, 0), false, child2d, g)) return false;
        exit();
        if (restore) popP();
        return true;
    }
    bool binary(const gsdf_tree_node &n, bool restore, bool is2d, uint32_t op, float k, bool hasK) {
        if (n.nchild != 2) { err = "binary operation needs 2 children"; return false; }
        if (!emit(b.child(n, 0), true, is2d)) return false;
        Guard g;
        if (!is2d && op == GSDF_OP_DIFF) g = guardFor(b.child(n, 1), GSDF_GUARD_DIFF);
        if (!is2d && op == GSDF_OP_SMOOTH_UNION && k > 0) g = guardFor(b.child(n, 1), GSDF_GUARD_SMOOTH_UNION, k);
        // box guard for the subtrahend of a 2-D difference: tiles that stay outside the minuend never evaluate it
        size_t bhw = 0;
        const bool bguard = guards && is2d && op == GSDF_OP_DIFF && boxBounded(b.child(n, 1));
        if (bguard) bhw = boxGuard(b.child(n, 1), GSDF_GUARD_DIFF, boxScale(b.Bounds2(b.child(n, 1))));
        if (!emit(b.child(n, 1), restore, is2d, g)) return false;
        if (bguard) patchBoxGuard(bhw, GSDF_GUARD_DIFF);  // lands on the DIFF op, which keeps `a`
        if (hasK) opf(op, k); else op0(op);
        popD();
        return true;
    }

    bool emit(NodeId id, bool restore, bool expect2d, Guard g = Guard{}) {
        if (!b.valid(id)) { err = "invalid node id in tree"; return false; }
        const gsdf_tree_node &n = b.node(id);
        if (expect2d != b.is2D(id)) { err = "2D/3D node kind mismatch in tree"; return false; }
        const float *f = n.fparam;
        const float *aux = n.aux_cnt ? &b.aux()[n.aux_off] : nullptr;
        switch (n.kind) {
        // ---------------- 3D primitives
        case GSDF_N_SPHERE: opf(GSDF_OP_SPHERE, f[0]); pushD(); return true;
        case GSDF_N_BOX:  // d := Scale(0.5, dims)  cpu_evaluators.go:29
            header(GSDF_OP_BOX, 2); chunk(0.5f * f[0], 0.5f * f[1], 0.5f * f[2], f[3]); pushD(); return true;
        case GSDF_N_BOXFRAME: {  // args primitives.go:292-297
            float e = f[3];
            header(GSDF_OP_BOXFRAME, 2);
            chunk(0.5f * f[0] + (-2 * e), 0.5f * f[1] + (-2 * e), 0.5f * f[2] + (-2 * e), e);
            pushD();
            return true;
        }
        case GSDF_N_TORUS: opf(GSDF_OP_TORUS, f[1], f[0]); pushD(); return true;  // (rGreater, rLesser)
        case GSDF_N_CYLINDER: {  // args primitives.go:147-149
            float round = f[2], h = (f[1] - 2 * round) / 2;
            header(GSDF_OP_CYLINDER, 2, round != 0 ? 1u : 0u); chunk(f[0], h, round); pushD();
            return true;
        }
        case GSDF_N_HEX: {  // clm := k3*h1  cpu_evaluators.go:94
            header(GSDF_OP_HEX, 2); chunk(f[0], f[1], 0.57735f * f[0]); pushD();
            return true;
        }
        // ---------------- 3D boolean / smooth
        case GSDF_N_UNION:
        case GSDF_N_UNION2D: {
            bool is2d = n.kind == GSDF_N_UNION2D;
            if (n.nchild < 2) { err = "OpUnion must have at least 2 elements"; return false; }  // operations.go:110-114
            // min is order-independent, so guardable children go last where the running min can guard them
            std::vector<NodeId> order;
            for (int k = 0; k < n.nchild; k++) if (is2d || !guards || !guardable(b.child(n, k))) order.push_back(b.child(n, k));
            for (int k = 0; k < n.nchild; k++) if (!(is2d || !guards || !guardable(b.child(n, k)))) order.push_back(b.child(n, k));
            // 2-D union of bounded shapes (text): upper bound from anchor points, then a box guard per operand
            if (is2d && guards && n.nchild >= 3) {
                std::vector<Vec2> anchors;
                int nbounded = 0;
                for (int k = 0; k < n.nchild; k++) {
                    std::vector<Vec2> a;
                    anchorsOf(order[k], a);
                    const size_t want = 8;  // a few per operand, evenly spaced along its outline
                    for (size_t i = 0; i < std::min(want, a.size()); i++) anchors.push_back(a[i * a.size() / std::min(want, a.size())]);
                    nbounded += boxBounded(order[k]) ? 1 : 0;
                }
                if (!anchors.empty() && nbounded >= 2) {
                    if (anchors.size() % 2) anchors.push_back(anchors.back());
                    const float margin = boxScale(b.Bounds2(id));
                    while (out.aux.size() % 4) out.aux.push_back(0.f);
                    const uint32_t off = (uint32_t)out.aux.size();
                    for (Vec2 v : anchors) { out.aux.push_back(v.x); out.aux.push_back(v.y); }
                    header(GSDF_OP_CULL_UB2D, 1, off, (uint32_t)anchors.size(), fbits(margin));
                    pushD();
                    for (int k = 0; k < n.nchild; k++) {
                        const bool last = k == n.nchild - 1;
                        const bool gd = boxBounded(order[k]);
                        size_t hw = gd ? boxGuard(order[k], GSDF_GUARD_MIN, margin) : 0;
                        if (!emit(order[k], last ? restore : true, true)) return false;
                        if (gd) patchBoxGuard(hw, GSDF_GUARD_MIN);  // lands on this operand's MIN, which keeps the running minimum
                        op0(GSDF_OP_MIN); popD();
                    }
                    return true;
                }
            }
            for (int k = 0; k < n.nchild; k++) {
                bool last = k == n.nchild - 1;
                Guard cg = (k > 0 && !is2d) ? guardFor(order[k], GSDF_GUARD_MIN) : Guard{};
                if (!emit(order[k], last ? restore : true, is2d, cg)) return false;
                if (k > 0) { op0(GSDF_OP_MIN); popD(); }
            }
            return true;
        }
        case GSDF_N_INTERSECT: return binary(n, restore, false, GSDF_OP_MAX, 0, false);
        case GSDF_N_DIFF: return binary(n, restore, false, GSDF_OP_DIFF, 0, false);
        case GSDF_N_XOR: return binary(n, restore, false, GSDF_OP_XOR, 0, false);
        case GSDF_N_SMOOTH_UNION: return binary(n, restore, false, GSDF_OP_SMOOTH_UNION, f[0], true);
        case GSDF_N_SMOOTH_DIFF: return binary(n, restore, false, GSDF_OP_SMOOTH_DIFF, f[0], true);
        case GSDF_N_SMOOTH_INTERSECT: return binary(n, restore, false, GSDF_OP_SMOOTH_INTERSECT, f[0], true);
        case GSDF_N_INTERSECT2D: return binary(n, restore, true, GSDF_OP_MAX, 0, false);
        case GSDF_N_DIFF2D: return binary(n, restore, true, GSDF_OP_DIFF, 0, false);
        case GSDF_N_XOR2D: return binary(n, restore, true, GSDF_OP_XOR, 0, false);
        // ---------------- 3D unary
        case GSDF_N_SCALE:
        case GSDF_N_SCALE2D: {  // factorInv := 1. / s.scale   cpu_evaluators.go:300, :1216
            float inv = 1.f / f[0], fac = f[0];
            return unaryPos(n, restore, n.kind == GSDF_N_SCALE2D, [&] { opf(GSDF_OP_SCALE_POS, inv); },
                            [&] { opf(GSDF_OP_MULDIST, fac); });
        }
        case GSDF_N_SHELL: {  // cpu_evaluators.go:435-449
            float th = f[0], inv = 1 / th;
            return unaryPos(n, restore, false, [&] { opf(GSDF_OP_SCALE_POS, inv); }, [&] { opf(GSDF_OP_SHELL_EXIT, th); });
        }
        case GSDF_N_SYMMETRY:
        case GSDF_N_SYMMETRY2D: {
            uint32_t mask = (uint32_t)n.iparam[0];
            return unaryPos(n, restore, n.kind == GSDF_N_SYMMETRY2D, [&] { header(GSDF_OP_SYMMETRY, 1, mask); }, [] {}, g);
        }
        case GSDF_N_TRANSFORM:
            return unaryPos(n, restore, false,
                            [&] {
                                header(GSDF_OP_TRANSFORM, 4);
                                chunk(f[0], f[1], f[2], f[3]); chunk(f[4], f[5], f[6], f[7]); chunk(f[8], f[9], f[10], f[11]);
                            },
                            [] {}, g);
        case GSDF_N_TRANSLATE:
            return unaryPos(n, restore, false, [&] { header(GSDF_OP_TRANSLATE, 2); chunk(f[0], f[1], f[2]); }, [] {}, g);
        case GSDF_N_TRANSLATE2D:
            return unaryPos(n, restore, true, [&] { header(GSDF_OP_TRANSLATE, 2); chunk(f[0], f[1], 0.f); }, [] {});
        case GSDF_N_ROTATE2D:
            return unaryPos(n, restore, true, [&] { header(GSDF_OP_ROTATE2D, 2); chunk(f[0], f[1], f[2], f[3]); }, [] {});
        case GSDF_N_OFFSET:
        case GSDF_N_OFFSET2D:
            if (!emit(b.child(n, 0), restore, n.kind == GSDF_N_OFFSET2D)) return false;
            opf(GSDF_OP_OFFSET, f[0]);
            return true;
        case GSDF_N_BOUNDS3:  // glbuild wrappers only change Bounds(); the flattener sees through them
            return emit(b.child(n, 0), restore, false, g);
        case GSDF_N_BOUNDS2:
            return emit(b.child(n, 0), restore, true);
        case GSDF_N_ANNULUS2D:
            if (!emit(b.child(n, 0), restore, true)) return false;
            opf(GSDF_OP_ANNULUS, f[0]);
            return true;
        case GSDF_N_TWIST:
            return unaryPos(n, restore, false, [&] { opf(GSDF_OP_TWIST, f[0]); }, [] {}, g);
        case GSDF_N_ELONGATE: {  // h := Scale(0.5, e.h)  cpu_evaluators.go:412
            bool ok = unaryPos(n, restore, false,
                               [&] { header(GSDF_OP_ELONGATE, 2); chunk(0.5f * f[0], 0.5f * f[1], 0.5f * f[2]); pushD(); },
                               [&] { op0(GSDF_OP_ADD_BELOW); popD(); });
            return ok;
        }
        case GSDF_N_ELONGATE2D:
            return unaryPos(n, restore, true, [&] { opf(GSDF_OP_ELONGATE2D, 0.5f * f[0], 0.5f * f[1]); pushD(); },
                            [&] { op0(GSDF_OP_ADD_BELOW); popD(); });
        case GSDF_N_ARRAY:
        case GSDF_N_ARRAY2D: {  // 8 (4) child evaluations, min-reduced, cpu_evaluators.go:363-396, :931-960
            bool is2d = n.kind == GSDF_N_ARRAY2D;
            int nvar = is2d ? 4 : 8;
            pushP();
            for (int v = 0; v < nvar; v++) {
                if (v > 0) op0(GSDF_OP_PEEK_POS);
                if (is2d) {
                    header(GSDF_OP_ARRAY2D_VAR, 2, (uint32_t)v);
                    chunk(f[0], f[1], (float)n.iparam[0] + -1, (float)n.iparam[1] + -1);
                } else {
                    header(GSDF_OP_ARRAY_VAR, 3, (uint32_t)v);
                    chunk(f[0], f[1], f[2]);
                    chunk((float)n.iparam[0] + -1, (float)n.iparam[1] + -1, (float)n.iparam[2] + -1);
                }
                if (!emit(b.child(n, 0), false, is2d)) return false;
                if (v > 0) { op0(GSDF_OP_MIN); popD(); }
            }
            opf(GSDF_OP_MIN_CONST, 1e20f);  // the reference's fold starts from largenum (gsdf.go:21, cpu_evaluators.go:364,932)
            popP();  // restores p (harmless when not needed)
            return true;
        }
        case GSDF_N_CIRCARRAY:
        case GSDF_N_CIRCARRAY2D: {  // cpu_evaluators.go:1056-1090
            bool is2d = n.kind == GSDF_N_CIRCARRAY2D;
            float ncirc = (float)n.iparam[1];
            float angle = m32::kTwoPiF / ncirc;
            float ninsm1 = (float)(n.iparam[0] - 1);
            if (restore) pushP();
            header(GSDF_OP_CIRC_ENTER, 2); chunk(angle, ncirc, ninsm1);
            if (++p > pmax) pmax = p;  // CIRC_ENTER pushes p0
            if (!emit(b.child(n, 0), false, is2d)) return false;  // evaluated at p1 first (:1082)
            popP();                                               // p = p0
            if (!emit(b.child(n, 0), false, is2d)) return false;
            op0(GSDF_OP_MIN); popD();
            if (restore) popP();
            return true;
        }
        // ---------------- 2D -> 3D
        case GSDF_N_EXTRUDE: {  // h := e.h / 2  cpu_evaluators.go:524
            size_t hw = out.chunks.size();
            header(GSDF_OP_EXTRUDE_ENTER, 1, 0, fbits(f[0] / 2), fbits(g.k)); pushD();
            if (!emit(b.child(n, 0), restore, true)) return false;
            op0(GSDF_OP_EXTRUDE_EXIT); popD();
            patchGuard(hw, g);
            return true;
        }
        case GSDF_N_REVOLVE:
            return unaryPos(n, restore, true, [&] { opf(GSDF_OP_REVOLVE, f[0]); }, [] {});
        case GSDF_N_SCREW: {  // threads.go:151-155: atanTaper := math.Tan(taper)
            float tanTaper = m32::tan(f[3]);
            size_t hw = 0;
            return unaryPos(n, restore, true,
                            [&] {
                                hw = out.chunks.size();
                                header(GSDF_OP_SCREW_ENTER, 2, 0, 0, fbits(g.k)); chunk(f[0], f[1], f[2], tanTaper); pushD();
                            },
                            [&] { op0(GSDF_OP_MAX_BELOW); popD(); patchGuard(hw, g); });
        }
        // ---------------- 2D primitives
        case GSDF_N_CIRCLE2D: opf(GSDF_OP_CIRCLE2D, f[0]); pushD(); return true;
        case GSDF_N_RECT2D: opf(GSDF_OP_RECT2D, 0.5f * f[0], 0.5f * f[1]); pushD(); return true;
        case GSDF_N_LINE2D: {  // cpu_evaluators.go:552-555
            float bax = f[3] - f[1], bay = f[4] - f[2];
            header(GSDF_OP_LINE2D, 3); chunk(f[1], f[2], bax, bay); chunk(bax * bax + bay * bay, f[0] / 2);
            pushD();
            return true;
        }
        case GSDF_N_LINES2D: {
            uint32_t off = (uint32_t)out.aux.size();
            out.aux.insert(out.aux.end(), aux, aux + n.aux_cnt);
            header(GSDF_OP_LINES2D, 1, off, (uint32_t)(n.aux_cnt / 4), fbits(f[0] / 2));
            pushD();
            return true;
        }
        case GSDF_N_ARC2D: {  // cpu_evaluators.go:565-569
            float s, c;
            m32::sincos(f[1] / 2, s, c);
            header(GSDF_OP_ARC2D, 3); chunk(f[0], f[2] / 2, s, c); chunk(f[0] * s, f[0] * c);
            pushD();
            return true;
        }
        case GSDF_N_EQTRI2D: {  // cpu_evaluators.go:670-671
            float r = f[0] / (float)kSqrt3;
            opf(GSDF_OP_EQTRI2D, r, r / (float)kSqrt3); pushD();
            return true;
        }
        case GSDF_N_HEX2D: opf(GSDF_OP_HEX2D, f[0], 0.577350269f * f[0]); pushD(); return true;
        case GSDF_N_OCT2D: opf(GSDF_OP_OCT2D, f[0], 0.4142135623f * f[0]); pushD(); return true;
        case GSDF_N_DIAMOND2D: {
            float bx = 0.5f * f[0], by = 0.5f * f[1];
            header(GSDF_OP_DIAMOND2D, 2); chunk(bx, by, bx * bx + by * by); pushD();
            return true;
        }
        case GSDF_N_ROUNDX2D: opf(GSDF_OP_ROUNDX2D, f[0], f[1]); pushD(); return true;
        case GSDF_N_POLY2D: {  // cpu_evaluators.go:793-818: per-edge records, edge iv runs v1=verts[iv], v2=verts[iv-1]
            int nv = n.aux_cnt / 2;
            if (nv < 3) { err = "polygon needs at least 3 vertices"; return false; }
            while (out.aux.size() % 4) out.aux.push_back(0.f);  // 16-byte align the records
            uint32_t off = (uint32_t)out.aux.size();
            int jv = nv - 1;
            for (int iv = 0; iv < nv; iv++) {
                float v1x = aux[2 * iv], v1y = aux[2 * iv + 1], v2x = aux[2 * jv], v2y = aux[2 * jv + 1];
                float ex = v2x - v1x, ey = v2y - v1y;
                float rec[GSDF_POLY_EDGE_FLOATS] = {v1x, v1y, ex, ey, ex * ex + ey * ey, v2y, 0.f, 0.f};
                out.aux.insert(out.aux.end(), rec, rec + GSDF_POLY_EDGE_FLOATS);
                jv = iv;
            }
            header(GSDF_OP_POLY2D, 1, off, (uint32_t)nv);
            pushD();
            return true;
        }
        case GSDF_N_TRANSLATEMULTI2D: {  // cpu_evaluators.go:1167-1182: min over displaced copies
            int nd = n.aux_cnt / 2;
            pushP();
            for (int k = 0; k < nd; k++) {
                if (k > 0) op0(GSDF_OP_PEEK_POS);
                header(GSDF_OP_TRANSLATE, 2); chunk(aux[2 * k], aux[2 * k + 1], 0.f);
                if (!emit(b.child(n, 0), false, true)) return false;
                if (k > 0) { op0(GSDF_OP_MIN); popD(); }
            }
            opf(GSDF_OP_MIN_CONST, 3.40282346638528859811704183484516925440e+38f);  // math.MaxFloat32, cpu_evaluators.go:1172
            popP();
            return true;
        }
        case GSDF_N_ELLIPSE2D: opf(GSDF_OP_ELLIPSE2D, f[0], f[1]); pushD(); return true;
        case GSDF_N_BEZIERQ2D: {  // per-shape constants of cpu_evaluators.go:583-593
            float Ax = f[0], Ay = f[1], Bx = f[2], By = f[3], Cx = f[4], Cy = f[5];
            float ax = Bx - Ax, ay = By - Ay;
            float a2 = ax * ax + ay * ay;
            float bx = Ax + (Cx - 2 * Bx), by = Ay + (Cy - 2 * By);
            float cx = 2 * ax, cy = 2 * ay;
            float kk = 1.f / (bx * bx + by * by);
            float kx = kk * (ax * bx + ay * by);
            header(GSDF_OP_BEZIERQ2D, 4, 0, fbits(f[6] / 2), 0);
            chunk(Ax, Ay, ax, ay); chunk(bx, by, cx, cy); chunk(kk, kx, kx * kx, a2);
            pushD();
            return true;
        }
        }
        err = "unknown node kind";
        return false;
    }
};

// Radius reuse (include/gsdf_program.h, "Radius reuse"; GSDF_RXY=0 disables): a post-pass over the finished
// straight-line stream. `ver` names the current (x, y) of the machine symbolically: ops that can change x or y give it a
// fresh name, the position stack restores earlier names, and a one-slot cache remembers for which name the radius
// Hypot(x, y) was stored. Stores happen only outside every region a guard can skip, so the simulation is exact.
void planRadiusReuse(Program &out) {
    std::vector<uint32_t> &c = out.chunks;
    const uint32_t nch = (uint32_t)(c.size() / 4);
    std::vector<int> delta(nch + 1, 0);      // +1 where a skippable region starts, -1 at its jump target
    std::vector<uint8_t> isTarget(nch + 1, 0);
    for (uint32_t pc = 0; pc < nch;) {
        const uint32_t op = c[4 * pc] & 0xff, len = (c[4 * pc] >> 8) & 0xff;
        if ((op == GSDF_OP_EXTRUDE_ENTER || op == GSDF_OP_SCREW_ENTER) && (c[4 * pc + 1] & 0xff)) {
            const uint32_t t = c[4 * pc + 1] >> 8;   // a firing slab guard skips the ENTER op itself
            delta[pc]++; delta[t]--; isTarget[t] = 1;
        }
        if (op == GSDF_OP_BBOX_GUARD2D) {
            const uint32_t t = c[4 * pc + 1] >> 8;   // the guard op always runs; the operand behind it may not
            delta[pc + len]++; delta[t]--; isTarget[t] = 1;
        }
        if (op == GSDF_OP_END || len == 0) break;
        pc += len;
    }
    uint32_t ver = 1, next = 2, slot = 0;
    long writer = -1;          // word index of the flag word of the op that filled the slot
    bool writerRead = false;
    std::vector<uint32_t> pstk;
    int depth = 0;
    auto fresh = [&] { ver = next++; };
    auto retire = [&] { if (writer >= 0 && !writerRead) c[(size_t)writer] &= ~GSDF_RXY_WRITE; };  // nobody read it: plain op
    for (uint32_t pc = 0; pc < nch;) {
        const uint32_t op = c[4 * pc] & 0xff, len = (c[4 * pc] >> 8) & 0xff;
        depth += delta[pc];
        if (isTarget[pc]) fresh();  // a skipped region may or may not have moved p: what follows must not assume either
        switch (op) {
        case GSDF_OP_CYLINDER: case GSDF_OP_TORUS: case GSDF_OP_CIRCLE2D: case GSDF_OP_SCREW_ENTER: {
            const size_t fw = 4 * (size_t)pc + (op == GSDF_OP_SCREW_ENTER ? 2 : 1);
            if (slot == ver) { c[fw] |= GSDF_RXY_READ; writerRead = true; }
            else if (depth == 0) { retire(); c[fw] |= GSDF_RXY_WRITE; slot = ver; writer = (long)fw; writerRead = false; }
            if (op == GSDF_OP_SCREW_ENTER) fresh();  // p = (sawtooth, radius)
        } break;
        case GSDF_OP_PUSH_POS: pstk.push_back(ver); break;
        case GSDF_OP_POP_POS: if (!pstk.empty()) { ver = pstk.back(); pstk.pop_back(); } else fresh(); break;
        case GSDF_OP_PEEK_POS: if (!pstk.empty()) ver = pstk.back(); else fresh(); break;
        case GSDF_OP_CIRC_ENTER: pstk.push_back(next++); fresh(); break;  // parks the rotated p0, continues at p1
        case GSDF_OP_TRANSLATE: if (c[4 * (pc + 1)] != 0u || c[4 * (pc + 1) + 1] != 0u) fresh(); break;  // x - (+0) == x bit for bit
        case GSDF_OP_SYMMETRY: if (c[4 * pc + 1] & 3u) fresh(); break;
        case GSDF_OP_SCALE_POS: case GSDF_OP_TRANSFORM: case GSDF_OP_ROTATE2D: case GSDF_OP_TWIST: case GSDF_OP_ELONGATE:
        case GSDF_OP_ELONGATE2D: case GSDF_OP_ARRAY_VAR: case GSDF_OP_ARRAY2D_VAR: case GSDF_OP_REVOLVE:
            fresh(); break;
        default: break;
        }
        if (op == GSDF_OP_END || len == 0) break;
        pc += len;
    }
    retire();
}

}  // namespace

bool Flatten(const Builder &b, NodeId root, Program &out, std::string &err) {
    out = Program{};
    if (!b.valid(root)) { err = "invalid root node"; return false; }
    out.dim = b.is2D(root) ? 2 : 3;
    Emitter e(b, out);
    if (!e.emit(root, false, out.dim == 2)) { err = e.err; return false; }
    e.header(GSDF_OP_END, 1);
    if (e.d != 1) { err = "internal: distance stack imbalance"; return false; }
    {   // radius reuse (gsdf_program.h): on by default, GSDF_RXY=0 switches the post-pass off (A/B, -DGSDF_NO_RXY libraries)
        const char *rx = std::getenv("GSDF_RXY");
        if (!(rx && *rx == '0')) planRadiusReuse(out);
    }
    out.dstack = e.dmax > 1 ? e.dmax - 1 : 1;  // top is cached in a register; slot 0 also absorbs the first push
    out.pstack = e.pmax;
    while (out.aux.size() % 4) out.aux.push_back(0.f);
    return true;
}

std::vector<uint8_t> Program::blob() const {
    gsdf_program_header h;
    std::memset(&h, 0, sizeof h);
    h.magic = GSDF_PROGRAM_MAGIC; h.version = GSDF_PROGRAM_VERSION;
    h.nchunks = (uint32_t)(chunks.size() / 4); h.dim = (uint32_t)dim;
    h.dstack = (uint32_t)dstack; h.pstack = (uint32_t)pstack; h.ninstr = (uint32_t)ninstr;
    std::vector<uint8_t> r(sizeof h + chunks.size() * 4);
    std::memcpy(r.data(), &h, sizeof h);
    std::memcpy(r.data() + sizeof h, chunks.data(), chunks.size() * 4);
    return r;
}

}  // namespace gsdfhost
