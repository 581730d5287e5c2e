"""A CPU model of the node-program interpreter (gsdf_b200/csrc/interp.cuh) for the `-m "not gpu"` tests.

It executes the FLATTENED program (include/gsdf_program.h: the bytes gsdf_program_create receives) on numpy float32
arrays, opcode by opcode, with the stack discipline, the operand layouts and the CTA-uniform guard jumps of the device
interpreter. Every float32 operation is an individually rounded numpy operation in the order the device code performs
it; the transcendental functions are the oracle's restated math32 routines, called per element. The oracle evaluates
the TREE, so   progsim(flatten(tree)) == oracle(tree)   bit for bit is a check of the host-side flattener (operand
order, derived constants, stack slots, position liveness, guard targets) that needs no GPU. It is test
infrastructure: nothing in the product path imports it. Every opcode is modelled.
"""
import ctypes as C
import struct

import numpy as np

F = np.float32
TILE = 1536  # points per CTA tile on the device (384 threads x 4 points): the granularity of the guards' "all points" vote

OPS = """END SPHERE BOX BOXFRAME TORUS CYLINDER HEX CIRCLE2D RECT2D LINE2D LINES2D ARC2D EQTRI2D HEX2D OCT2D DIAMOND2D ROUNDX2D
POLY2D ELLIPSE2D BEZIERQ2D MIN MAX DIFF XOR SMOOTH_UNION SMOOTH_DIFF SMOOTH_INTERSECT OFFSET ANNULUS MULDIST SHELL_EXIT ADD_BELOW
EXTRUDE_EXIT MAX_BELOW PUSH_POS POP_POS PEEK_POS TRANSLATE SCALE_POS SYMMETRY TRANSFORM ROTATE2D TWIST ELONGATE ELONGATE2D
ARRAY_VAR ARRAY2D_VAR CIRC_ENTER EXTRUDE_ENTER REVOLVE SCREW_ENTER CULL_UB2D BBOX_GUARD2D MIN_CONST""".split()
OP = {name: i for i, name in enumerate(OPS)}
GUARD_DIFF, GUARD_MIN, GUARD_SMOOTH_UNION = 1, 2, 3
RXY_READ, RXY_WRITE = 0x100, 0x200  # experimental radius reuse (include/gsdf_program.h)
# True reproduces the box-guard predicate as first shipped (points INSIDE the operand's box could vote for the skip): kept so
# that a test can show the overlapping-operand shapes of tests/shapes.py::overlap2d catch exactly that defect.
LEGACY_BOX_GUARD = False
UNSUPPORTED = set()  # every opcode is modelled (ellipse2D / quadbezier2d: the EXT interpreter's primitives, per element)


class Math:
    """Element-wise wrappers over the oracle's restated math32 functions."""

    def __init__(self, oracle):
        self.L = oracle.lib()

    def _1(self, fn, x):
        x = np.asarray(x, F)
        return np.array([fn(C.c_float(float(v))) for v in x.ravel()], dtype=F).reshape(x.shape)

    def _2(self, fn, x, y):
        x, y = np.broadcast_arrays(np.asarray(x, F), np.asarray(y, F))
        return np.array([fn(C.c_float(float(a)), C.c_float(float(b))) for a, b in zip(x.ravel(), y.ravel())], dtype=F).reshape(x.shape)

    def acos(self, x): return self._1(self.L.go_acos, x)
    def cbrt(self, x): return self._1(self.L.go_cbrt, x)

    def pow_frac(self, x, y):
        """math32.Pow for the quadratic Bezier's cube roots: 0 -> 0, 1 -> 1, else Exp(y * Log(x)) (oracle go_pow_frac)."""
        x = np.asarray(x, F)
        with np.errstate(divide="ignore", invalid="ignore"):
            r = self._1(self.L.go_exp, F(y) * self._1(self.L.go_log, np.where(x > 0, x, F(1))))
        return np.where(x == 0, F(0), np.where(x == 1, F(1), r)).astype(F)

    def sin(self, x): return self._1(self.L.go_sin, x)
    def cos(self, x): return self._1(self.L.go_cos, x)
    def atan2(self, y, x): return self._2(self.L.go_atan2, y, x)

    @staticmethod
    def hypot(p, q):
        """math32.Hypot: p*sqrt(1+(q/p)^2) with p >= q (oracle go_hypot, device hypot32), vectorised."""
        p, q = np.abs(np.asarray(p, F)), np.abs(np.asarray(q, F))
        p, q = np.broadcast_arrays(p, q)
        hi, lo = np.maximum(p, q), np.minimum(p, q)
        with np.errstate(divide="ignore", invalid="ignore"):
            r = lo / hi
            out = hi * np.sqrt(F(1) + r * r)
        return np.where(hi == 0, F(0), out).astype(F)

    def norm2(self, x, y): return self.hypot(x, y)
    def norm3(self, x, y, z): return self.hypot(x, self.hypot(y, z))


def clampf(v, lo, hi):
    return np.where(v < lo, F(lo), np.where(v > hi, F(hi), v)).astype(F)


def signf(x):
    return np.where(x == 0, F(0), np.where(x > 0, F(1), F(-1))).astype(F)


def roundf(x):
    """Half away from zero, exact in float32."""
    r = np.trunc(x)
    return (r + np.where(np.abs(x - r) >= F(0.5), np.copysign(F(1), x), F(0))).astype(F)


def guard_dead(kind, w, a, k):
    if kind == GUARD_DIFF:
        return -w < a
    if kind == GUARD_MIN:
        return w > a
    return ((w - a) >= F(k)) & (a != 0)


class Program:
    def __init__(self, blob, aux):
        magic, ver, nchunks, dim, dstack, pstack, ninstr, _ = struct.unpack_from("<8I", blob, 0)
        assert magic == 0x46445347 and ver == 1 and len(blob) == 32 + 16 * nchunks
        self.dim, self.dstack, self.pstack = dim, dstack, pstack
        self.u = np.frombuffer(blob, np.uint32, offset=32).reshape(-1, 4)
        self.f = np.frombuffer(blob, np.float32, offset=32).reshape(-1, 4)
        self.aux4 = np.concatenate([np.asarray(aux, F).ravel(), np.zeros((-len(np.asarray(aux).ravel())) % 4, F)]).reshape(-1, 4)

    def ops(self):
        pc, out = 0, []
        while True:
            op, ln = int(self.u[pc, 0]) & 0xff, (int(self.u[pc, 0]) >> 8) & 0xff
            out.append(op)
            if op == 0:
                return out
            pc += ln

    def supported(self):
        return not (set(self.ops()) & UNSUPPORTED)


def run(prog, pos, M, tile=TILE, stats=None):
    """Distances of the program at pos ((n,3) or (n,2) float32), tile by tile like the device's CTAs."""
    pos = np.ascontiguousarray(pos, F)
    out = np.empty(len(pos), F)
    for a in range(0, len(pos), tile):
        out[a:a + tile] = _run_tile(prog, pos[a:a + tile], M, stats)
    return out


def _run_tile(P, pos, M, stats):
    n = len(pos)
    px, py = pos[:, 0].copy(), pos[:, 1].copy()
    pz = pos[:, 2].copy() if pos.shape[1] == 3 else np.zeros(n, F)
    top = np.zeros(n, F)
    dstk, pstk = [], []
    skip = False
    rxy = [None]  # the one-slot radius cache

    def radius(flags):
        """Hypot(px, py), from the cache when the program says it holds the radius of these very x, y."""
        if flags & RXY_READ:
            assert rxy[0] is not None, "radius cache read before any write"
            # the flattener's claim, checked directly: the cached radius is Hypot of bit-identical x, y
            assert np.array_equal(rxy[0].view(np.uint32), M.hypot(px, py).view(np.uint32)), "stale radius cache"
            if stats is not None: stats["rxy_reads"] = stats.get("rxy_reads", 0) + 1
            return rxy[0]
        r = M.hypot(px, py)
        if flags & RXY_WRITE:
            rxy[0] = r
            if stats is not None: stats["rxy_writes"] = stats.get("rxy_writes", 0) + 1
        return r

    max_d = max_p = 0
    pc = 0
    u, f = P.u, P.f
    while True:
        h = u[pc]
        op, ln = int(h[0]) & 0xff, (int(h[0]) >> 8) & 0xff
        w1, w2 = int(h[1]), int(h[2])
        f2, f3 = f[pc, 2], f[pc, 3]
        c1 = f[pc + 1] if ln > 1 else None
        c2 = f[pc + 2] if ln > 2 else None
        c3 = f[pc + 3] if ln > 3 else None

        def pushD():
            nonlocal max_d
            dstk.append(top)
            max_d = max(max_d, len(dstk))

        name = OPS[op]
        if name == "END":
            break
        elif name == "SPHERE":
            pushD(); top = M.norm3(px, py, pz) - f2
        elif name == "BOX":
            pushD()
            qx, qy, qz = (np.abs(px) - c1[0]) + c1[3], (np.abs(py) - c1[1]) + c1[3], (np.abs(pz) - c1[2]) + c1[3]
            top = M.norm3(np.maximum(qx, 0), np.maximum(qy, 0), np.maximum(qz, 0)) + np.minimum(np.maximum(qx, np.maximum(qy, qz)), 0) - c1[3]
        elif name == "BOXFRAME":
            pushD()
            e = c1[3]
            x, y, z = np.abs(px) - c1[0], np.abs(py) - c1[1], np.abs(pz) - c1[2]
            qx, qy, qz = np.abs(x + e) + (-e), np.abs(y + e) + (-e), np.abs(z + e) + (-e)
            n1 = M.norm3(np.maximum(x, 0), np.maximum(qy, 0), np.maximum(qz, 0)) + np.minimum(F(0), np.maximum(x, np.maximum(qy, qz)))
            n2 = M.norm3(np.maximum(qx, 0), np.maximum(y, 0), np.maximum(qz, 0)) + np.minimum(F(0), np.maximum(qx, np.maximum(y, qz)))
            n3 = M.norm3(np.maximum(qx, 0), np.maximum(qy, 0), np.maximum(z, 0)) + np.minimum(F(0), np.maximum(qx, np.maximum(qy, z)))
            top = np.minimum(n1, np.minimum(n2, n3))
        elif name == "TORUS":
            pushD(); top = M.norm2(radius(w1) - f2, pz) - f3
        elif name == "CYLINDER":
            pushD()
            if w1 & 1 == 0:
                dx, dy = radius(w1) - c1[0], np.abs(pz) - c1[1]
                top = np.minimum(F(0), np.maximum(dx, dy)) + M.hypot(np.maximum(F(0), dx), np.maximum(F(0), dy))
            else:
                dx, dy = radius(w1) - c1[0] + c1[2], np.abs(pz) - c1[1]
                top = np.minimum(np.maximum(dx, dy), F(0)) + M.hypot(np.maximum(dx, 0), np.maximum(dy, 0)) - c1[2]
        elif name == "HEX":
            pushD()
            k1, twok1 = F(-0.8660254037844386), F(2 * -0.8660254037844386)
            x, y, z = np.abs(px), np.abs(py), np.abs(pz)
            pm = np.minimum(k1 * x + F(0.5) * y, F(0))
            x = x - twok1 * pm
            y = y - F(1.0) * pm
            d1 = M.hypot(x - clampf(x, -c1[2], c1[2]), y - c1[0]) * signf(y - c1[0])
            d2 = z - c1[1]
            top = np.minimum(np.maximum(d1, d2), F(0)) + M.hypot(np.maximum(d1, 0), np.maximum(d2, 0))
        elif name == "CIRCLE2D":
            pushD(); top = radius(w1) - f2
        elif name == "RECT2D":
            pushD()
            dx, dy = np.abs(px) - f2, np.abs(py) - f3
            top = M.norm2(np.maximum(dx, 0), np.maximum(dy, 0)) + np.minimum(F(0), np.maximum(dx, dy))
        elif name == "LINE2D":
            pushD()
            pax, pay = px - c1[0], py - c1[1]
            hh = clampf((pax * c1[2] + pay * c1[3]) / c2[0], 0, 1)
            top = M.norm2(pax - hh * c1[2], pay - hh * c1[3]) - c2[1]
        elif name == "LINES2D":
            seg = P.aux4[w1 >> 2:]
            d = np.full(n, F(1e23))
            for s in range(w2):
                ax, ay, bx, by = seg[s]
                bax, bay = bx - ax, by - ay
                dotba = bax * bax + bay * bay
                pax, pay = px - ax, py - ay
                hh = clampf((pax * bax + pay * bay) / dotba, 0, 1)
                ex, ey = pax - hh * bax, pay - hh * bay
                d = np.minimum(d, ex * ex + ey * ey)
            pushD(); top = np.sqrt(d) - f3
        elif name == "ARC2D":
            pushD()
            x, y = np.abs(px), py
            top = np.where(c1[3] * x > c1[2] * y, M.norm2(x - c2[0], y - c2[1]) - c1[1], np.abs(M.norm2(x, y) - c1[0]) - c1[1])
        elif name == "EQTRI2D":
            pushD()
            k = F(1.7320508075688772)
            x, y = np.abs(px) - f2, py + f3
            fold = x + k * y > 0
            nx, ny = x - k * y, -k * x - y
            x, y = np.where(fold, F(0.5) * nx, x), np.where(fold, F(0.5) * ny, y)
            x = x - clampf(x, F(-2.0) * f2, 0)
            top = -M.norm2(x, y) * signf(y)
        elif name in ("HEX2D", "OCT2D"):
            pushD()
            x, y = np.abs(px), np.abs(py)
            if name == "HEX2D":
                kx, ky = F(-0.8660254037844386), F(0.5)
                mm = F(2) * np.minimum(kx * x + ky * y, F(0))
                x, y = x - mm * kx, y - mm * ky
            else:
                kx, ky = F(-0.9238795325), F(0.3826834323)
                mm = F(2) * np.minimum(kx * x + ky * y, F(0))
                x, y = x - mm * kx, y - mm * ky
                mm = F(2) * np.minimum(-kx * x + ky * y, F(0))
                x, y = x - mm * -kx, y - mm * ky
            x, y = x - clampf(x, -f3, f3), y - f2
            top = signf(y) * M.norm2(x, y)
        elif name == "DIAMOND2D":
            pushD()
            x, y = np.abs(px), np.abs(py)
            ux, uy = c1[0] - F(2) * x, c1[1] - F(2) * y
            hh = clampf((ux * c1[0] - uy * c1[1]) / c1[2], -1, 1)
            d = M.norm2(x - (F(0.5) * c1[0]) * (F(1) - hh), y - (F(0.5) * c1[1]) * (F(1) + hh))
            top = d * signf(x * c1[1] + y * c1[0] - c1[0] * c1[1])
        elif name == "ROUNDX2D":
            pushD()
            x, y = np.abs(px), np.abs(py)
            sub = F(0.5) * np.minimum(x + y, f2)
            top = M.norm2(x - sub, y - sub) - f3
        elif name == "ELLIPSE2D":  # cpu_evaluators.go:750-791, interp.cuh EXT
            pushD()
            sq3 = F(1.7320508075688772)
            x, y = np.abs(px), np.abs(py)
            swap = x > y
            a = np.where(swap, f3, f2).astype(F)
            b = np.where(swap, f2, f3).astype(F)
            x, y = np.where(swap, y, x).astype(F), np.where(swap, x, y).astype(F)
            with np.errstate(all="ignore"):
                l = b * b - a * a
                mm = a * x / l; m2 = mm * mm
                nn = b * y / l; n2 = nn * nn
                c = (m2 + n2 - F(1)) / F(3)
                c3_ = c * c * c
                q = c3_ + F(2) * m2 * n2
                d = c3_ + m2 * n2
                g = mm + mm * n2
                # d < 0 branch
                hh = M.acos(q / c3_) / F(3)
                sh, ch = M.sin(hh), M.cos(hh)
                t = sq3 * sh
                rx = np.sqrt(-c * (ch + t + F(2)) + m2)
                ry = np.sqrt(-c * (ch - t + F(2)) + m2)
                co1 = (ry + signf(l) * rx + np.abs(g) / (rx * ry) - mm) / F(2)
                # d >= 0 branch
                h2 = F(2) * mm * nn * np.sqrt(d)
                s_ = signf(q + h2) * M.cbrt(np.abs(q + h2))
                u_ = signf(q - h2) * M.cbrt(np.abs(q - h2))
                rx2 = -s_ - u_ - F(4) * c + F(2) * m2
                ry2 = sq3 * (s_ - u_)
                rm = M.hypot(rx2, ry2)
                co2 = (ry2 / np.sqrt(rm - rx2) + F(2) * g / rm - mm) / F(2)
                co = np.where(d < 0, co1, co2).astype(F)
                ex, ey = a * co, b * np.sqrt(F(1) - co * co)
                top = M.norm2(ex - x, ey - y) * signf(y - ey)
        elif name == "BEZIERQ2D":  # cpu_evaluators.go:581-659, interp.cuh EXT
            pushD()
            sq3 = F(1.7320508075688772)
            third = F(1. / 3)
            with np.errstate(all="ignore"):
                dx, dy = c1[0] - px, c1[1] - py
                ky = c3[0] * (F(2) * c3[3] + (dx * c2[0] + dy * c2[1])) / F(3)
                kz = c3[0] * (dx * c1[2] + dy * c1[3])
                g = ky - c3[2]
                q = c3[1] * (F(2) * c3[2] - F(3) * ky) + kz
                g3 = g * g * g
                q2 = q * q
                hh = q2 + F(4) * g3
                # hh >= 0
                sh = np.sqrt(hh)
                xx, xy = F(0.5) * (sh + -q), F(0.5) * (-sh + -q)
                k = (F(1.0) - g3 / q2) * g3 / q
                small = np.abs(g) < F(0.001)
                xx = np.where(small, k, xx).astype(F); xy = np.where(small, -k - q, xy).astype(F)
                ux = signf(xx) * M.pow_frac(np.abs(xx), third)
                uy = signf(xy) * M.pow_frac(np.abs(xy), third)
                t = ux + uy
                t = t - (t * (t * t + F(3.0) * g) + q) / (F(3.0) * t * t + F(3.0) * g)
                t = clampf(t - c3[1], 0, 1)
                wx, wy = dx + t * (c2[2] + t * c2[0]), dy + t * (c2[3] + t * c2[1])
                res1 = wx * wx + wy * wy
                # hh < 0
                z = np.sqrt(-g)
                xm = np.sqrt(F(0.5) + F(0.5) * (q / (F(2) * g * z)))
                mm = xm * (xm * (xm * (xm * F(-0.008972) + F(0.039071)) - F(0.107074)) + F(0.576975)) + F(0.5)
                nn = np.sqrt(F(1) - mm * mm)
                nn = nn * sq3
                tx = clampf((mm + mm) * z - c3[1], 0, 1)
                ty = clampf((-nn - mm) * z - c3[1], 0, 1)
                qxx, qxy = dx + tx * (c2[2] + tx * c2[0]), dy + tx * (c2[3] + tx * c2[1])
                qyx, qyy = dx + ty * (c2[2] + ty * c2[0]), dy + ty * (c2[3] + ty * c2[1])
                ddx, ddy = qxx * qxx + qxy * qxy, qyx * qyx + qyy * qyy
                res2 = np.where(ddx < ddy, ddx, ddy)
                res = np.where(hh >= 0, res1, res2).astype(F)
                top = np.sqrt(res) - f2
        elif name == "POLY2D":
            rec = P.aux4[w1 >> 2:]
            ax, ay = px - rec[0][0], py - rec[0][1]
            d = ax * ax + ay * ay
            neg = np.zeros(n, bool)
            for iv in range(w2):
                ra, rb = rec[2 * iv], rec[2 * iv + 1]
                wx, wy = px - ra[0], py - ra[1]
                c = clampf((wx * ra[2] + wy * ra[3]) / rb[0], 0, 1)
                bx, by = wx - c * ra[2], wy - c * ra[3]
                d = np.minimum(d, bx * bx + by * by)
                b1, b2, b3 = py >= ra[1], py < rb[1], ra[2] * wy > ra[3] * wx
                neg ^= (b1 & b2 & b3) | (~b1 & ~b2 & ~b3)
            pushD(); top = np.where(neg, F(-1), F(1)) * np.sqrt(d)
        elif name in ("MIN", "DIFF", "SMOOTH_UNION") and skip:
            skip = False  # a guard fired: the combiner keeps `a`, which is still on top
        elif name == "MIN":
            top = np.minimum(dstk.pop(), top)
        elif name == "MAX":
            top = np.maximum(dstk.pop(), top)
        elif name == "DIFF":
            top = np.maximum(dstk.pop(), -top)
        elif name == "XOR":
            a = dstk.pop(); top = np.maximum(np.minimum(a, top), -np.maximum(a, top))
        elif name == "SMOOTH_UNION":
            a, b = dstk.pop(), top
            hh = clampf(F(0.5) + F(0.5) * (b - a) / f2, 0, 1)
            top = (b * (F(1) - hh) + a * hh) - f2 * hh * (F(1) - hh)
        elif name == "SMOOTH_DIFF":
            a, b = dstk.pop(), top
            hh = clampf(F(0.5) - F(0.5) * (b + a) / f2, 0, 1)
            top = (a * (F(1) - hh) + (-b) * hh) + f2 * hh * (F(1) - hh)
        elif name == "SMOOTH_INTERSECT":
            a, b = dstk.pop(), top
            hh = clampf(F(0.5) - F(0.5) * (b - a) / f2, 0, 1)
            top = (b * (F(1) - hh) + a * hh) + f2 * hh * (F(1) - hh)
        elif name == "OFFSET":
            top = top + f2
        elif name == "ANNULUS":
            top = np.abs(top) - f2
        elif name == "MIN_CONST":
            top = np.minimum(F(f2), top)
        elif name == "MULDIST":
            top = top * f2
        elif name == "SHELL_EXIT":
            top = f2 * (np.abs(top) - f2)
        elif name == "ADD_BELOW":
            top = top + dstk.pop()
        elif name == "EXTRUDE_EXIT":
            wy, d = dstk.pop(), top
            top = np.minimum(F(0), np.maximum(d, wy)) + M.hypot(np.maximum(d, 0), np.maximum(wy, 0))
        elif name == "MAX_BELOW":
            top = np.maximum(top, dstk.pop())
        elif name == "PUSH_POS":
            pstk.append((px, py, pz)); max_p = max(max_p, len(pstk))
        elif name == "POP_POS":
            px, py, pz = pstk.pop()
        elif name == "PEEK_POS":
            px, py, pz = pstk[-1]
        elif name == "TRANSLATE":
            px, py, pz = px - c1[0], py - c1[1], pz - c1[2]
        elif name == "SCALE_POS":
            px, py, pz = f2 * px, f2 * py, f2 * pz
        elif name == "SYMMETRY":
            if w1 & 1: px = np.abs(px)
            if w1 & 2: py = np.abs(py)
            if w1 & 4: pz = np.abs(pz)
        elif name == "TRANSFORM":
            x, y, z = px, py, pz
            px = c1[0] * x + c1[1] * y + c1[2] * z + c1[3]
            py = c2[0] * x + c2[1] * y + c2[2] * z + c2[3]
            pz = c3[0] * x + c3[1] * y + c3[2] * z + c3[3]
        elif name == "ROTATE2D":
            x, y = px, py
            px, py = c1[0] * x + c1[1] * y, c1[2] * x + c1[3] * y
        elif name == "TWIST":
            ang = f2 * pz
            s, c = M.sin(ang), M.cos(ang)
            x, y = px, py
            px, py = c * x - s * y, s * x + c * y
        elif name == "ELONGATE":
            pushD()
            qx, qy, qz = np.abs(px) - c1[0], np.abs(py) - c1[1], np.abs(pz) - c1[2]
            top = np.minimum(np.maximum(qx, np.maximum(qy, qz)), F(0))
            px, py, pz = np.maximum(qx, F(0)), np.maximum(qy, F(0)), np.maximum(qz, F(0))
        elif name == "ELONGATE2D":
            pushD()
            qx, qy = np.abs(px) - f2, np.abs(py) - f3
            top = np.minimum(np.maximum(qx, qy), F(0))
            px, py = np.maximum(qx, F(0)), np.maximum(qy, F(0))
        elif name == "ARRAY_VAR":
            fi, fj, fk = F(w1 & 1), F((w1 >> 1) & 1), F((w1 >> 2) & 1)
            x, y, z = px, py, pz
            idx, idy, idz = roundf(x / c1[0]), roundf(y / c1[1]), roundf(z / c1[2])
            ox, oy, oz = signf(x - c1[0] * idx), signf(y - c1[1] * idy), signf(z - c1[2] * idz)
            rx, ry, rz = clampf(idx + fi * ox, 0, c2[0]), clampf(idy + fj * oy, 0, c2[1]), clampf(idz + fk * oz, 0, c2[2])
            px, py, pz = x - c1[0] * rx, y - c1[1] * ry, z - c1[2] * rz
        elif name == "ARRAY2D_VAR":
            fi, fj = F(w1 & 1), F((w1 >> 1) & 1)
            x, y = px, py
            idx, idy = roundf(x / c1[0]), roundf(y / c1[1])
            ox, oy = signf(x - c1[0] * idx), signf(y - c1[1] * idy)
            rx, ry = clampf(idx + fi * ox, 0, c1[2]), clampf(idy + fj * oy, 0, c1[3])
            px, py = x - c1[0] * rx, y - c1[1] * ry
        elif name == "CIRC_ENTER":
            x, y = px, py
            idf = np.floor(M.atan2(y, x) / c1[0])
            idf = np.where(idf < 0, idf + c1[1], idf).astype(F)
            wrap = idf >= c1[2]
            i0 = np.where(wrap, c1[2], idf).astype(F)
            i1 = np.where(wrap, F(0), idf + F(1)).astype(F)
            s0, c0, s1, cc1 = M.sin(c1[0] * i0), M.cos(c1[0] * i0), M.sin(c1[0] * i1), M.cos(c1[0] * i1)
            pstk.append((c0 * x + s0 * y, -s0 * x + c0 * y, pz)); max_p = max(max_p, len(pstk))
            px, py = cc1 * x + s1 * y, -s1 * x + cc1 * y
        elif name == "CULL_UB2D":
            an = P.aux4[w1 >> 2:]
            best = np.full(n, F(3.0e38))
            for q in range(w2 >> 1):
                vx, vy, vz, vw = an[q]
                ax, ay, bx, by = px - vx, py - vy, px - vz, py - vw
                best = np.minimum(best, np.minimum(ax * ax + ay * ay, bx * bx + by * by))
            pushD(); top = np.sqrt(best) * F(1.0001) + f3
        elif name == "BBOX_GUARD2D":
            dx = np.maximum(np.maximum(c1[0] - px, px - c1[2]), F(0))
            dy = np.maximum(np.maximum(c1[1] - py, py - c1[3]), F(0))
            w = np.sqrt(dx * dx + dy * dy) * F(0.9999) - f3
            outside = (dx > 0) | (dy > 0) | LEGACY_BOX_GUARD  # the box bounds the operand from below only OUTSIDE the box
            if (outside & guard_dead(w1 & 0xff, w, top, 0)).all():
                if stats is not None: stats["fired"] = stats.get("fired", 0) + 1
                skip = True; pc = w1 >> 8
                continue
        elif name == "EXTRUDE_ENTER":
            if w1 & 0xff:
                if stats is not None: stats["guards"] = stats.get("guards", 0) + 1
                if guard_dead(w1 & 0xff, np.abs(pz) - f2, top, f3).all():
                    if stats is not None: stats["fired"] = stats.get("fired", 0) + 1
                    skip = True; pc = w1 >> 8
                    continue
            pushD(); top = np.abs(pz) - f2
        elif name == "REVOLVE":
            px = M.hypot(px, pz) - f2
        elif name == "SCREW_ENTER":
            if w1 & 0xff:
                if stats is not None: stats["guards"] = stats.get("guards", 0) + 1
                if guard_dead(w1 & 0xff, np.abs(pz) - c1[2], top, f3).all():
                    if stats is not None: stats["fired"] = stats.get("fired", 0) + 1
                    skip = True; pc = w1 >> 8
                    continue
            pushD()
            x, y, z = px, py, pz
            yy = radius(w2)
            yy = yy + z * c1[3]
            theta = M.atan2(y, x)
            zz = z + c1[1] * theta / F(2 * np.pi)
            sx = zz + c1[0] / F(2)
            t = sx / c1[0]
            px = c1[0] * (t - np.floor(t)) - c1[0] / F(2)
            py = yy
            top = np.abs(z) - c1[2]
        else:
            raise NotImplementedError(name)
        top = np.asarray(top, F)
        px, py, pz = np.asarray(px, F), np.asarray(py, F), np.asarray(pz, F)
        pc += ln
    # the header's stack sizes are what the device allocates: slot 0 of the distance stack absorbs the first push
    assert max(max_d - 1, 1) <= P.dstack and max_p <= P.pstack, (max_d, P.dstack, max_p, P.pstack)
    assert len(dstk) == 1 and not pstk and not skip, (len(dstk), len(pstk), skip)
    return top


def main(argv):
    """python tests/progsim.py program.blob program.aux points.npy [out.npy] -- run a flattened program (the two buffers a
    host-side flattener hands to gsdf_program_create, e.g. written by the Go flattener of integration/go/) on the CPU
    model and print / save the distances. Lets a maintainer check a flattener against the C++ one without a GPU."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle as O
    O.build()
    blob = open(argv[1], "rb").read()
    aux = np.fromfile(argv[2], dtype=np.float32)
    pos = np.load(argv[3]).astype(np.float32)
    P = Program(blob, aux)
    assert P.supported(), "the model does not implement ellipse2D / quadbezier2d"
    assert pos.ndim == 2 and pos.shape[1] == P.dim, "points must be (n, %d)" % P.dim
    d = run(P, pos, Math(O))
    if len(argv) > 4:
        np.save(argv[4], d)
    else:
        np.set_printoptions(precision=9, suppress=False)
        print(d)
    return 0


if __name__ == "__main__":
    import sys
    sys.exit(main(sys.argv))
