package gsdf

// CUDA program emitters of the node types: what gsdf_b200/csrc/host/flatten.cpp emits for the same node, case by case.
// Every derived constant is computed in float32 exactly as the node's CPU Evaluate derives it (cpu_evaluators.go), so the
// kernels stay bit-comparable with the CPU path. These methods live in package gsdf because the node structs and their
// fields are unexported; the Program type and the tree walk are glbuild/cuda_program.go.

import (
	"errors"

	math "github.com/chewxy/math32"
	"github.com/soypat/geometry/ms2"
	"github.com/soypat/gsdf/glbuild"
)

func fbits(f float32) uint32 { return math.Float32bits(f) }

// ---------------------------------------------------------------------------------------------- 3D primitives

func (s *sphere) AppendProgram(p *glbuild.Program, _ bool) error { p.Opf(glbuild.OpSphere, s.r, 0); p.PushD(); return nil }

func (s *box) AppendProgram(p *glbuild.Program, _ bool) error { // d := Scale(0.5, dims), cpu_evaluators.go:29
	p.Header(glbuild.OpBox, 2, 0, 0, 0)
	p.Chunk(0.5*s.dims.X, 0.5*s.dims.Y, 0.5*s.dims.Z, s.round)
	p.PushD()
	return nil
}

func (s *boxframe) AppendProgram(p *glbuild.Program, _ bool) error { // args, primitives.go:292-297
	e := s.e
	p.Header(glbuild.OpBoxFrame, 2, 0, 0, 0)
	p.Chunk(0.5*s.dims.X+(-2*e), 0.5*s.dims.Y+(-2*e), 0.5*s.dims.Z+(-2*e), e)
	p.PushD()
	return nil
}

func (s *torus) AppendProgram(p *glbuild.Program, _ bool) error { p.Opf(glbuild.OpTorus, s.rGreater, s.rLesser); p.PushD(); return nil }

func (s *cylinder) AppendProgram(p *glbuild.Program, _ bool) error {
	r, h, round := s.args() // primitives.go:147-149
	flag := uint32(0)
	if round != 0 {
		flag = 1
	}
	p.Header(glbuild.OpCylinder, 2, flag, 0, 0)
	p.Chunk(r, h, round, 0)
	p.PushD()
	return nil
}

func (s *hex) AppendProgram(p *glbuild.Program, _ bool) error { // clm := k3*h1, cpu_evaluators.go:94
	p.Header(glbuild.OpHex, 2, 0, 0, 0)
	p.Chunk(s.side, s.h, 0.57735*s.side, 0)
	p.PushD()
	return nil
}

// ---------------------------------------------------------------------------------------------- booleans

func (u *OpUnion) AppendProgram(p *glbuild.Program, restore bool) error {
	if len(u.joined) < 2 {
		return errors.New("OpUnion must have at least 2 elements") // operations.go:110-114
	}
	// min is order-independent (math32.Min of finite values), so guardable operands go last, where the running minimum
	// can guard them (flatten.cpp, GSDF_N_UNION)
	guarded := func(c glbuild.Shader3D) bool { return p.GuardFor(c, glbuild.GuardMin, 0).Kind != glbuild.GuardNone }
	order := make([]glbuild.Shader3D, 0, len(u.joined))
	for _, c := range u.joined {
		if !guarded(c) {
			order = append(order, c)
		}
	}
	for _, c := range u.joined {
		if guarded(c) {
			order = append(order, c)
		}
	}
	for k := range order {
		last := k == len(order)-1
		if k > 0 {
			p.SetGuard(p.GuardFor(order[k], glbuild.GuardMin, 0))
		}
		if err := glbuild.Emit(p, order[k], restore || !last); err != nil {
			return err
		}
		if k > 0 {
			p.Op0(glbuild.OpMin)
			p.PopD()
		}
	}
	return nil
}

// ---- 2-D box guards (include/gsdf_program.h, "box guards"; flatten.cpp boxBounded / anchorsOf / boxScale) ----

// boxBounded: shapes whose Bounds() truly encloses them and whose value outside that box is >= the distance to it: exact
// primitives, translations, unions of such, and differences of such (max(a, -b) >= a). Bounds overrides are NOT in the list:
// glbuild.Unwrap is not applied here on purpose, an OverloadShader2DBounds box is cosmetic.
func boxBounded(s glbuild.Shader2D) bool {
	switch n := s.(type) {
	case *poly2D, *circle2D, *rect2D:
		return true
	case *translate2D:
		return boxBounded(n.s)
	case *diff2D:
		return boxBounded(n.s1)
	case *OpUnion2D:
		for _, c := range n.joined {
			if !boxBounded(c) {
				return false
			}
		}
		return len(n.joined) > 0
	}
	return false
}

// anchorsOf appends points ON the outline of s in the current frame: the value of s at p is <= |p - v| for each of them.
func anchorsOf(s glbuild.Shader2D, out []ms2.Vec) []ms2.Vec {
	switch n := s.(type) {
	case *poly2D:
		return append(out, n.vert...)
	case *circle2D:
		return append(out, ms2.Vec{X: n.r}, ms2.Vec{X: -n.r}, ms2.Vec{Y: n.r}, ms2.Vec{Y: -n.r})
	case *rect2D:
		hx, hy := 0.5*n.d.X, 0.5*n.d.Y
		return append(out, ms2.Vec{X: hx, Y: hy}, ms2.Vec{X: -hx, Y: hy}, ms2.Vec{X: hx, Y: -hy}, ms2.Vec{X: -hx, Y: -hy})
	case *translate2D:
		first := len(out)
		out = anchorsOf(n.s, out)
		for i := first; i < len(out); i++ {
			out[i].X += n.p.X
			out[i].Y += n.p.Y
		}
		return out
	case *OpUnion2D: // min(a, b) <= each operand
		for _, c := range n.joined {
			out = anchorsOf(c, out)
		}
		return out
	case *diff2D: // max(a, -b) <= |p - v| for v on a's outline and outside b: only when b's box really encloses b
		if !boxBounded(n.s2) {
			return out
		}
		bb := n.s2.Bounds()
		for _, v := range anchorsOf(n.s1, nil) {
			if v.X < bb.Min.X || v.X > bb.Max.X || v.Y < bb.Min.Y || v.Y > bb.Max.Y {
				out = append(out, v)
			}
		}
		return out
	}
	return out
}

func boxScale(bb ms2.Box) float32 {
	return 1e-5 * math.Max(math.Max(math.Abs(bb.Min.X), math.Abs(bb.Max.X)), math.Max(math.Abs(bb.Min.Y), math.Abs(bb.Max.Y)))
}

func (u *OpUnion2D) AppendProgram(p *glbuild.Program, restore bool) error {
	if len(u.joined) < 2 {
		return errors.New("OpUnion2D must have at least 2 elements")
	}
	// Union of bounded shapes (what forge/textsdf builds): an upper bound U of the union from a few anchor points per
	// operand, pushed as an extra operand of the min fold, then a box guard in front of every bounded operand
	// (flatten.cpp, GSDF_N_UNION2D). min(U, d1..dn) == min(d1..dn) because U >= the operand that owns the nearest anchor.
	if glbuild.BoxGuardsOn() && len(u.joined) >= 3 {
		var anchors []ms2.Vec
		nbounded := 0
		for _, c := range u.joined {
			a := anchorsOf(c, nil)
			want := 8
			if len(a) < want {
				want = len(a)
			}
			for i := 0; i < want; i++ {
				anchors = append(anchors, a[i*len(a)/want])
			}
			if boxBounded(c) {
				nbounded++
			}
		}
		if len(anchors) > 0 && nbounded >= 2 {
			if len(anchors)%2 != 0 {
				anchors = append(anchors, anchors[len(anchors)-1])
			}
			margin := boxScale(u.Bounds())
			off := p.AlignAux(4)
			for _, v := range anchors {
				p.Aux = append(p.Aux, v.X, v.Y)
			}
			p.Header(glbuild.OpCullUB2D, 1, off, uint32(len(anchors)), fbits(margin))
			p.PushD()
			for k, c := range u.joined {
				last := k == len(u.joined)-1
				gd := boxBounded(c)
				hw := 0
				if gd {
					bb := c.Bounds()
					hw = p.BoxGuard(glbuild.GuardMin, margin, bb.Min.X, bb.Min.Y, bb.Max.X, bb.Max.Y)
				}
				if err := glbuild.Emit(p, c, restore || !last); err != nil {
					return err
				}
				if gd {
					p.PatchBoxGuard(hw, glbuild.GuardMin) // lands on this operand's MIN, which keeps the running minimum
				}
				p.Op0(glbuild.OpMin)
				p.PopD()
			}
			return nil
		}
	}
	for k := range u.joined {
		last := k == len(u.joined)-1
		if err := glbuild.Emit(p, u.joined[k], restore || !last); err != nil {
			return err
		}
		if k > 0 {
			p.Op0(glbuild.OpMin)
			p.PopD()
		}
	}
	return nil
}

func (u *diff) AppendProgram(p *glbuild.Program, r bool) error {
	return p.BinaryGuarded(u.s1, u.s2, r, glbuild.GuardDiff, 0, func() { p.Op0(glbuild.OpDiff) })
}
func (u *intersect) AppendProgram(p *glbuild.Program, r bool) error   { return p.Binary(u.s1, u.s2, r, func() { p.Op0(glbuild.OpMax) }) }
func (u *xor) AppendProgram(p *glbuild.Program, r bool) error         { return p.Binary(u.s1, u.s2, r, func() { p.Op0(glbuild.OpXor) }) }
// diff2D: the subtrahend gets a box guard -- tiles that stay outside the hole never evaluate it (flatten.cpp, binary()).
func (u *diff2D) AppendProgram(p *glbuild.Program, r bool) error {
	if err := glbuild.Emit(p, u.s1, true); err != nil {
		return err
	}
	bguard := glbuild.BoxGuardsOn() && boxBounded(u.s2)
	hw := 0
	if bguard {
		bb := u.s2.Bounds()
		hw = p.BoxGuard(glbuild.GuardDiff, boxScale(bb), bb.Min.X, bb.Min.Y, bb.Max.X, bb.Max.Y)
	}
	if err := glbuild.Emit(p, u.s2, r); err != nil {
		return err
	}
	if bguard {
		p.PatchBoxGuard(hw, glbuild.GuardDiff) // lands on the DIFF op, which keeps `a`
	}
	p.Op0(glbuild.OpDiff)
	p.PopD()
	return nil
}
func (u *intersect2D) AppendProgram(p *glbuild.Program, r bool) error { return p.Binary(u.s1, u.s2, r, func() { p.Op0(glbuild.OpMax) }) }
func (u *xor2D) AppendProgram(p *glbuild.Program, r bool) error       { return p.Binary(u.s1, u.s2, r, func() { p.Op0(glbuild.OpXor) }) }
func (u *smoothUnion) AppendProgram(p *glbuild.Program, r bool) error {
	kind := glbuild.GuardNone
	if u.k > 0 {
		kind = glbuild.GuardSmoothUnion
	}
	return p.BinaryGuarded(u.s1, u.s2, r, kind, u.k, func() { p.Opf(glbuild.OpSmoothUnion, u.k, 0) })
}
func (u *smoothDiff) AppendProgram(p *glbuild.Program, r bool) error {
	return p.Binary(u.s1, u.s2, r, func() { p.Opf(glbuild.OpSmoothDiff, u.k, 0) })
}
func (u *smoothIntersect) AppendProgram(p *glbuild.Program, r bool) error {
	return p.Binary(u.s1, u.s2, r, func() { p.Opf(glbuild.OpSmoothIntersect, u.k, 0) })
}

// ---------------------------------------------------------------------------------------------- 3D unary

func (u *scale) AppendProgram(p *glbuild.Program, r bool) error { // factorInv := 1. / s.scale, cpu_evaluators.go:300
	return p.Unary(u.s, r, func() { p.Opf(glbuild.OpScalePos, 1./u.scale, 0) }, func() { p.Opf(glbuild.OpMulDist, u.scale, 0) })
}
func (u *scale2D) AppendProgram(p *glbuild.Program, r bool) error { // cpu_evaluators.go:1216
	return p.Unary(u.s, r, func() { p.Opf(glbuild.OpScalePos, 1./u.scale, 0) }, func() { p.Opf(glbuild.OpMulDist, u.scale, 0) })
}
func (u *shell) AppendProgram(p *glbuild.Program, r bool) error { // cpu_evaluators.go:435-449
	return p.Unary(u.s, r, func() { p.Opf(glbuild.OpScalePos, 1/u.thick, 0) }, func() { p.Opf(glbuild.OpShellExit, u.thick, 0) })
}
func (u *symmetry) AppendProgram(p *glbuild.Program, r bool) error {
	return p.Unary(u.s, r, func() { p.Header(glbuild.OpSymmetry, 1, uint32(u.xyz), 0, 0) }, nil)
}
func (u *symmetry2D) AppendProgram(p *glbuild.Program, r bool) error {
	return p.Unary(u.s, r, func() { p.Header(glbuild.OpSymmetry, 1, uint32(u.xy), 0, 0) }, nil)
}
func (u *transform) AppendProgram(p *glbuild.Program, r bool) error { // rows of tInv; MulPosition uses w = 1 (cpu_evaluators.go:497)
	m := u.tInv.Array() // row-major 4x4
	return p.Unary(u.s, r, func() {
		p.Header(glbuild.OpTransform, 4, 0, 0, 0)
		p.Chunk(m[0], m[1], m[2], m[3])
		p.Chunk(m[4], m[5], m[6], m[7])
		p.Chunk(m[8], m[9], m[10], m[11])
	}, nil)
}
func (u *translate) AppendProgram(p *glbuild.Program, r bool) error {
	return p.Unary(u.s, r, func() { p.Header(glbuild.OpTranslate, 2, 0, 0, 0); p.Chunk(u.p.X, u.p.Y, u.p.Z, 0) }, nil)
}
func (u *translate2D) AppendProgram(p *glbuild.Program, r bool) error {
	return p.Unary(u.s, r, func() { p.Header(glbuild.OpTranslate, 2, 0, 0, 0); p.Chunk(u.p.X, u.p.Y, 0, 0) }, nil)
}
func (u *rotation2D) AppendProgram(p *glbuild.Program, r bool) error { // tInv applied to p, cpu_evaluators.go:1186
	m := u.tInv.Array()
	return p.Unary(u.s, r, func() { p.Header(glbuild.OpRotate2D, 2, 0, 0, 0); p.Chunk(m[0], m[1], m[2], m[3]) }, nil)
}
func (u *offset) AppendProgram(p *glbuild.Program, r bool) error {
	if err := glbuild.Emit(p, u.s, r); err != nil {
		return err
	}
	p.Opf(glbuild.OpOffset, u.off, 0)
	return nil
}
func (u *offset2D) AppendProgram(p *glbuild.Program, r bool) error {
	if err := glbuild.Emit(p, u.s, r); err != nil {
		return err
	}
	p.Opf(glbuild.OpOffset, u.f, 0)
	return nil
}
func (u *annulus2D) AppendProgram(p *glbuild.Program, r bool) error {
	if err := glbuild.Emit(p, u.s, r); err != nil {
		return err
	}
	p.Opf(glbuild.OpAnnulus, u.r, 0)
	return nil
}
func (u *twist) AppendProgram(p *glbuild.Program, r bool) error {
	return p.Unary(u.s, r, func() { p.Opf(glbuild.OpTwist, u.k, 0) }, nil)
}
func (u *elongate) AppendProgram(p *glbuild.Program, r bool) error { // h := Scale(0.5, e.h), cpu_evaluators.go:412
	return p.Unary(u.s, r, func() {
		p.Header(glbuild.OpElongate, 2, 0, 0, 0)
		p.Chunk(0.5*u.h.X, 0.5*u.h.Y, 0.5*u.h.Z, 0)
		p.PushD()
	}, func() { p.Op0(glbuild.OpAddBelow); p.PopD() })
}
func (u *elongate2D) AppendProgram(p *glbuild.Program, r bool) error {
	return p.Unary(u.s, r, func() { p.Opf(glbuild.OpElongate2D, 0.5*u.h.X, 0.5*u.h.Y); p.PushD() }, func() { p.Op0(glbuild.OpAddBelow); p.PopD() })
}

// array / array2D: 8 (4) child evaluations, min-reduced (cpu_evaluators.go:363-396, 931-960).
func (u *array) AppendProgram(p *glbuild.Program, _ bool) error {
	p.PushP()
	for v := 0; v < 8; v++ {
		if v > 0 {
			p.Op0(glbuild.OpPeekPos)
		}
		p.Header(glbuild.OpArrayVar, 3, uint32(v), 0, 0)
		p.Chunk(u.d.X, u.d.Y, u.d.Z, 0)
		p.Chunk(float32(u.nx)+-1, float32(u.ny)+-1, float32(u.nz)+-1, 0)
		if err := glbuild.Emit(p, u.s, false); err != nil {
			return err
		}
		if v > 0 {
			p.Op0(glbuild.OpMin)
			p.PopD()
		}
	}
	p.Opf(glbuild.OpMinConst, largenum, 0) // the fold starts from largenum (cpu_evaluators.go:364)
	p.PopP()
	return nil
}
func (u *array2D) AppendProgram(p *glbuild.Program, _ bool) error {
	p.PushP()
	for v := 0; v < 4; v++ {
		if v > 0 {
			p.Op0(glbuild.OpPeekPos)
		}
		p.Header(glbuild.OpArray2DVar, 2, uint32(v), 0, 0)
		p.Chunk(u.d.X, u.d.Y, float32(u.nx)+-1, float32(u.ny)+-1)
		if err := glbuild.Emit(p, u.s, false); err != nil {
			return err
		}
		if v > 0 {
			p.Op0(glbuild.OpMin)
			p.PopD()
		}
	}
	p.Opf(glbuild.OpMinConst, largenum, 0) // cpu_evaluators.go:932
	p.PopP()
	return nil
}

// circarray / circarray2D: the child is evaluated at the two neighbouring sector positions (cpu_evaluators.go:1056-1090).
func circ(p *glbuild.Program, child glbuild.Shader, nInst, circleDiv int, restore bool) error {
	ncirc := float32(circleDiv)
	angle := 2 * math.Pi / ncirc
	if restore {
		p.PushP()
	}
	p.Header(glbuild.OpCircEnter, 2, 0, 0, 0)
	p.Chunk(angle, ncirc, float32(nInst-1), 0)
	p.P++ // CIRC_ENTER parks p0 on the position stack
	if p.P > p.Pmax {
		p.Pmax = p.P
	}
	if err := glbuild.Emit(p, child, false); err != nil { // evaluated at p1 first
		return err
	}
	p.PopP() // p = p0
	if err := glbuild.Emit(p, child, false); err != nil {
		return err
	}
	p.Op0(glbuild.OpMin)
	p.PopD()
	if restore {
		p.PopP()
	}
	return nil
}
func (u *circarray) AppendProgram(p *glbuild.Program, r bool) error   { return circ(p, u.s, u.nInst, u.circleDiv, r) }
func (u *circarray2D) AppendProgram(p *glbuild.Program, r bool) error { return circ(p, u.s, u.nInst, u.circleDiv, r) }

// ---------------------------------------------------------------------------------------------- 2D -> 3D

func (u *extrusion) AppendProgram(p *glbuild.Program, r bool) error { // h := e.h / 2, cpu_evaluators.go:524
	g := p.TakeGuard()
	hw := len(p.Chunks)
	p.Header(glbuild.OpExtrudeEnter, 1, 0, fbits(u.h/2), fbits(g.K))
	p.PushD()
	if err := glbuild.Emit(p, u.s, r); err != nil {
		return err
	}
	p.Op0(glbuild.OpExtrudeExit)
	p.PopD()
	p.PatchGuard(hw, g)
	return nil
}

// Slab guards (glbuild/cuda_program.go): the extrusion evaluates them; these nodes only move p and pass them on.
func (u *extrusion) SlabBounded()                            {}
func (u *translate) DistTransparentChild() glbuild.Shader   { return u.s }
func (u *transform) DistTransparentChild() glbuild.Shader   { return u.s }
func (u *symmetry) DistTransparentChild() glbuild.Shader    { return u.s }
func (u *twist) DistTransparentChild() glbuild.Shader       { return u.s }
func (u *revolution) AppendProgram(p *glbuild.Program, r bool) error {
	return p.Unary(u.s2d, r, func() { p.Opf(glbuild.OpRevolve, u.off, 0) }, nil)
}

// ---------------------------------------------------------------------------------------------- 2D primitives

func (c *circle2D) AppendProgram(p *glbuild.Program, _ bool) error { p.Opf(glbuild.OpCircle2D, c.r, 0); p.PushD(); return nil }
func (c *rect2D) AppendProgram(p *glbuild.Program, _ bool) error {
	p.Opf(glbuild.OpRect2D, 0.5*c.d.X, 0.5*c.d.Y)
	p.PushD()
	return nil
}
func (c *line2D) AppendProgram(p *glbuild.Program, _ bool) error { // cpu_evaluators.go:552-555
	ba := ms2.Sub(c.b, c.a)
	p.Header(glbuild.OpLine2D, 3, 0, 0, 0)
	p.Chunk(c.a.X, c.a.Y, ba.X, ba.Y)
	p.Chunk(ba.X*ba.X+ba.Y*ba.Y, c.width/2, 0, 0)
	p.PushD()
	return nil
}
func (c *lines2D) AppendProgram(p *glbuild.Program, _ bool) error {
	off := uint32(len(p.Aux))
	for _, seg := range c.points {
		p.Aux = append(p.Aux, seg[0].X, seg[0].Y, seg[1].X, seg[1].Y)
	}
	p.Header(glbuild.OpLines2D, 1, off, uint32(len(c.points)), fbits(c.width/2))
	p.PushD()
	return nil
}
func (c *arc2D) AppendProgram(p *glbuild.Program, _ bool) error { // cpu_evaluators.go:565-569
	s, cs := math.Sincos(c.angle / 2)
	p.Header(glbuild.OpArc2D, 3, 0, 0, 0)
	p.Chunk(c.radius, c.thick/2, s, cs)
	p.Chunk(c.radius*s, c.radius*cs, 0, 0)
	p.PushD()
	return nil
}
func (c *equilateralTri2d) AppendProgram(p *glbuild.Program, _ bool) error { // cpu_evaluators.go:670-671
	r := c.hTri / sqrt3
	p.Opf(glbuild.OpEqTri2D, r, r/sqrt3)
	p.PushD()
	return nil
}
func (c *hex2D) AppendProgram(p *glbuild.Program, _ bool) error { p.Opf(glbuild.OpHex2D, c.side, 0.577350269*c.side); p.PushD(); return nil }
func (c *oct2D) AppendProgram(p *glbuild.Program, _ bool) error { p.Opf(glbuild.OpOct2D, c.c, 0.4142135623*c.c); p.PushD(); return nil }
func (c *diamond) AppendProgram(p *glbuild.Program, _ bool) error {
	bx, by := 0.5*c.d.X, 0.5*c.d.Y
	p.Header(glbuild.OpDiamond2D, 2, 0, 0, 0)
	p.Chunk(bx, by, bx*bx+by*by, 0)
	p.PushD()
	return nil
}
func (c *x2d) AppendProgram(p *glbuild.Program, _ bool) error     { p.Opf(glbuild.OpRoundX2D, c.dim, c.thick); p.PushD(); return nil }
func (c *ellipse2D) AppendProgram(p *glbuild.Program, _ bool) error { p.Opf(glbuild.OpEllipse2D, c.a, c.b); p.PushD(); return nil }

// poly2D: 8 floats per edge, edge i runs v1 = verts[i], v2 = verts[i-1] (cpu_evaluators.go:793-818).
func (c *poly2D) AppendProgram(p *glbuild.Program, _ bool) error {
	if len(c.vert) < 3 {
		return errors.New("polygon needs at least 3 vertices")
	}
	off := p.AlignAux(4)
	j := len(c.vert) - 1
	for i, v1 := range c.vert {
		v2 := c.vert[j]
		e := ms2.Sub(v2, v1)
		p.Aux = append(p.Aux, v1.X, v1.Y, e.X, e.Y, e.X*e.X+e.Y*e.Y, v2.Y, 0, 0)
		j = i
	}
	p.Header(glbuild.OpPoly2D, 1, off, uint32(len(c.vert)), 0)
	p.PushD()
	return nil
}

// quadbezier2d: per-shape constants of cpu_evaluators.go:583-593.
func (c *quadbezier2d) AppendProgram(p *glbuild.Program, _ bool) error {
	A, B, C := c.a, c.b, c.c
	a := ms2.Sub(B, A)
	a2 := a.X*a.X + a.Y*a.Y
	bx, by := A.X+(C.X-2*B.X), A.Y+(C.Y-2*B.Y)
	cx, cy := 2*a.X, 2*a.Y
	kk := 1. / (bx*bx + by*by)
	kx := kk * (a.X*bx + a.Y*by)
	p.Header(glbuild.OpBezierQ2D, 4, 0, fbits(c.thick/2), 0)
	p.Chunk(A.X, A.Y, a.X, a.Y)
	p.Chunk(bx, by, cx, cy)
	p.Chunk(kk, kx, kx*kx, a2)
	p.PushD()
	return nil
}

// translateMulti2D: min over displaced copies (cpu_evaluators.go:1167-1182).
func (u *translateMulti2D) AppendProgram(p *glbuild.Program, _ bool) error {
	p.PushP()
	for k, d := range u.displacements {
		if k > 0 {
			p.Op0(glbuild.OpPeekPos)
		}
		p.Header(glbuild.OpTranslate, 2, 0, 0, 0)
		p.Chunk(d.X, d.Y, 0, 0)
		if err := glbuild.Emit(p, u.s, false); err != nil {
			return err
		}
		if k > 0 {
			p.Op0(glbuild.OpMin)
			p.PopD()
		}
	}
	p.Opf(glbuild.OpMinConst, math.MaxFloat32, 0) // cpu_evaluators.go:1172
	p.PopP()
	return nil
}
