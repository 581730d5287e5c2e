package gsdf

// CUDA node-program flattener: glbuild.Shader3D / Shader2D tree -> the packed program of include/gsdf_program.h that
// libgsdfb200.so interprets on the GPU. Mirrors gsdf_b200/csrc/host/flatten.cpp case by case; every derived constant
// is computed in float32 exactly as the node's CPU Evaluate derives it (cpu_evaluators.go), so the kernels stay
// bit-comparable with the CPU path.
//
// Lives in package gsdf because the node structs and their fields are unexported.

import (
	"errors"
	"fmt"
	"unsafe"

	math "github.com/chewxy/math32"
	"github.com/soypat/geometry/ms2"
	"github.com/soypat/gsdf/glbuild"
)

// Opcodes: enum gsdf_opcode of include/gsdf_program.h, in declaration order.
const (
	opEnd uint32 = iota
	opSphere
	opBox
	opBoxFrame
	opTorus
	opCylinder
	opHex
	opCircle2D
	opRect2D
	opLine2D
	opLines2D
	opArc2D
	opEqTri2D
	opHex2D
	opOct2D
	opDiamond2D
	opRoundX2D
	opPoly2D
	opEllipse2D
	opBezierQ2D
	opMin
	opMax
	opDiff
	opXor
	opSmoothUnion
	opSmoothDiff
	opSmoothIntersect
	opOffset
	opAnnulus
	opMulDist
	opShellExit
	opAddBelow
	opExtrudeExit
	opMaxBelow
	opPushPos
	opPopPos
	opPeekPos
	opTranslate
	opScalePos
	opSymmetry
	opTransform
	opRotate2D
	opTwist
	opElongate
	opElongate2D
	opArrayVar
	opArray2DVar
	opCircEnter
	opExtrudeEnter
	opRevolve
	opScrewEnter
)

const (
	programMagic   = 0x46445347 // "GSDF"
	programVersion = 1
	polyEdgeFloats = 8
)

// Program accumulates the instruction chunks (4 x uint32 each) and the float side buffer.
type Program struct {
	Chunks       []uint32
	Aux          []float32
	Dim          int
	d, p         int // current distance / position stack depth
	dmax, pmax   int
	ninstr       int
}

// ProgramEmitter is implemented by every node type of gsdf and forge/threads.
// restore: a later sibling still needs the current position, so the node must leave p as it found it.
type ProgramEmitter interface {
	AppendProgram(p *Program, restore bool) error
}

func fbits(f float32) uint32 { return math.Float32bits(f) }

func (p *Program) Header(op, nchunks, w1, w2, w3 uint32) {
	p.Chunks = append(p.Chunks, op|nchunks<<8, w1, w2, w3)
	p.ninstr++
}
func (p *Program) Chunk(a, b, c, d float32) { p.Chunks = append(p.Chunks, fbits(a), fbits(b), fbits(c), fbits(d)) }
func (p *Program) Op0(op uint32)            { p.Header(op, 1, 0, 0, 0) }
func (p *Program) Opf(op uint32, f2, f3 float32) { p.Header(op, 1, 0, fbits(f2), fbits(f3)) }
func (p *Program) PushD() {
	p.d++
	if p.d > p.dmax {
		p.dmax = p.d
	}
}
func (p *Program) PopD() { p.d-- }
func (p *Program) PushP() {
	p.Op0(opPushPos)
	p.p++
	if p.p > p.pmax {
		p.pmax = p.p
	}
}
func (p *Program) PopP() { p.Op0(opPopPos); p.p-- }
func (p *Program) alignAux(n int) uint32 {
	for len(p.Aux)%n != 0 {
		p.Aux = append(p.Aux, 0)
	}
	return uint32(len(p.Aux))
}

// Emit sees through the glbuild wrappers, then dispatches to the node's emitter.
func Emit(p *Program, s glbuild.Shader, restore bool) error {
	for {
		u := glbuild.Unwrap(s)
		if u == nil {
			break
		}
		s = u
	}
	e, ok := s.(ProgramEmitter)
	if !ok {
		return fmt.Errorf("%T has no CUDA program emitter", s)
	}
	return e.AppendProgram(p, restore)
}

// unary emits enter, the child, exit for a position transform.
func unary(p *Program, child glbuild.Shader, restore bool, enter, exit func()) error {
	if restore {
		p.PushP()
	}
	enter()
	if err := Emit(p, child, false); err != nil {
		return err
	}
	if exit != nil {
		exit()
	}
	if restore {
		p.PopP()
	}
	return nil
}

func binary(p *Program, a, b glbuild.Shader, restore bool, emitOp func()) error {
	if err := Emit(p, a, true); err != nil { // the first operand must leave p intact for the second
		return err
	}
	if err := Emit(p, b, restore); err != nil {
		return err
	}
	emitOp()
	p.PopD()
	return nil
}

// Flatten3 / Flatten2 replace Programmer.WriteComputeSDF3/2 (glbuild/glbuild.go:175,218) for the CUDA backend:
// blob = gsdf_program_header + chunks, aux = side buffer, ready for gsdf_program_create.
func Flatten3(root glbuild.Shader3D) (blob []byte, aux []float32, err error) { return flatten(root, 3) }
func Flatten2(root glbuild.Shader2D) (blob []byte, aux []float32, err error) { return flatten(root, 2) }

func flatten(root glbuild.Shader, dim int) ([]byte, []float32, error) {
	var p Program
	p.Dim = dim
	if err := Emit(&p, root, false); err != nil {
		return nil, nil, err
	}
	p.Header(opEnd, 1, 0, 0, 0)
	if p.d != 1 {
		return nil, nil, errors.New("internal: distance stack imbalance")
	}
	dstack := 1 // the top is cached in a register; slot 0 also absorbs the first push
	if p.dmax > 1 {
		dstack = p.dmax - 1
	}
	p.alignAux(4)
	hdr := [8]uint32{programMagic, programVersion, uint32(len(p.Chunks) / 4), uint32(dim), uint32(dstack), uint32(p.pmax), uint32(p.ninstr), 0}
	words := append(hdr[:], p.Chunks...)
	blob := unsafe.Slice((*byte)(unsafe.Pointer(&words[0])), 4*len(words)) // little-endian hosts, like the C side
	return append([]byte(nil), blob...), p.Aux, nil
}

// ---------------------------------------------------------------------------------------------- 3D primitives

func (s *sphere) AppendProgram(p *Program, _ bool) error { p.Opf(opSphere, s.r, 0); p.PushD(); return nil }

func (s *box) AppendProgram(p *Program, _ bool) error { // d := Scale(0.5, dims), cpu_evaluators.go:29
	p.Header(opBox, 2, 0, 0, 0)
	p.Chunk(0.5*s.dims.X, 0.5*s.dims.Y, 0.5*s.dims.Z, s.round)
	p.PushD()
	return nil
}

func (s *boxframe) AppendProgram(p *Program, _ bool) error { // args, primitives.go:292-297
	e := s.e
	p.Header(opBoxFrame, 2, 0, 0, 0)
	p.Chunk(0.5*s.dims.X+(-2*e), 0.5*s.dims.Y+(-2*e), 0.5*s.dims.Z+(-2*e), e)
	p.PushD()
	return nil
}

func (s *torus) AppendProgram(p *Program, _ bool) error { p.Opf(opTorus, s.rGreater, s.rLesser); p.PushD(); return nil }

func (s *cylinder) AppendProgram(p *Program, _ bool) error {
	r, h, round := s.args() // primitives.go:147-149
	flag := uint32(0)
	if round != 0 {
		flag = 1
	}
	p.Header(opCylinder, 2, flag, 0, 0)
	p.Chunk(r, h, round, 0)
	p.PushD()
	return nil
}

func (s *hex) AppendProgram(p *Program, _ bool) error { // clm := k3*h1, cpu_evaluators.go:94
	p.Header(opHex, 2, 0, 0, 0)
	p.Chunk(s.side, s.h, 0.57735*s.side, 0)
	p.PushD()
	return nil
}

// ---------------------------------------------------------------------------------------------- booleans

func (u *OpUnion) AppendProgram(p *Program, restore bool) error {
	if len(u.joined) < 2 {
		return errors.New("OpUnion must have at least 2 elements") // operations.go:110-114
	}
	for k := range u.joined {
		last := k == len(u.joined)-1
		if err := Emit(p, u.joined[k], restore || !last); err != nil {
			return err
		}
		if k > 0 {
			p.Op0(opMin)
			p.PopD()
		}
	}
	return nil
}

func (u *OpUnion2D) AppendProgram(p *Program, restore bool) error {
	if len(u.joined) < 2 {
		return errors.New("OpUnion2D must have at least 2 elements")
	}
	for k := range u.joined {
		last := k == len(u.joined)-1
		if err := Emit(p, u.joined[k], restore || !last); err != nil {
			return err
		}
		if k > 0 {
			p.Op0(opMin)
			p.PopD()
		}
	}
	return nil
}

func (u *diff) AppendProgram(p *Program, r bool) error        { return binary(p, u.s1, u.s2, r, func() { p.Op0(opDiff) }) }
func (u *intersect) AppendProgram(p *Program, r bool) error   { return binary(p, u.s1, u.s2, r, func() { p.Op0(opMax) }) }
func (u *xor) AppendProgram(p *Program, r bool) error         { return binary(p, u.s1, u.s2, r, func() { p.Op0(opXor) }) }
func (u *diff2D) AppendProgram(p *Program, r bool) error      { return binary(p, u.s1, u.s2, r, func() { p.Op0(opDiff) }) }
func (u *intersect2D) AppendProgram(p *Program, r bool) error { return binary(p, u.s1, u.s2, r, func() { p.Op0(opMax) }) }
func (u *xor2D) AppendProgram(p *Program, r bool) error       { return binary(p, u.s1, u.s2, r, func() { p.Op0(opXor) }) }
func (u *smoothUnion) AppendProgram(p *Program, r bool) error {
	return binary(p, u.s1, u.s2, r, func() { p.Opf(opSmoothUnion, u.k, 0) })
}
func (u *smoothDiff) AppendProgram(p *Program, r bool) error {
	return binary(p, u.s1, u.s2, r, func() { p.Opf(opSmoothDiff, u.k, 0) })
}
func (u *smoothIntersect) AppendProgram(p *Program, r bool) error {
	return binary(p, u.s1, u.s2, r, func() { p.Opf(opSmoothIntersect, u.k, 0) })
}

// ---------------------------------------------------------------------------------------------- 3D unary

func (u *scale) AppendProgram(p *Program, r bool) error { // factorInv := 1. / s.scale, cpu_evaluators.go:300
	return unary(p, u.s, r, func() { p.Opf(opScalePos, 1./u.scale, 0) }, func() { p.Opf(opMulDist, u.scale, 0) })
}
func (u *scale2D) AppendProgram(p *Program, r bool) error { // cpu_evaluators.go:1216
	return unary(p, u.s, r, func() { p.Opf(opScalePos, 1./u.scale, 0) }, func() { p.Opf(opMulDist, u.scale, 0) })
}
func (u *shell) AppendProgram(p *Program, r bool) error { // cpu_evaluators.go:435-449
	return unary(p, u.s, r, func() { p.Opf(opScalePos, 1/u.thick, 0) }, func() { p.Opf(opShellExit, u.thick, 0) })
}
func (u *symmetry) AppendProgram(p *Program, r bool) error {
	return unary(p, u.s, r, func() { p.Header(opSymmetry, 1, uint32(u.xyz), 0, 0) }, nil)
}
func (u *symmetry2D) AppendProgram(p *Program, r bool) error {
	return unary(p, u.s, r, func() { p.Header(opSymmetry, 1, uint32(u.xy), 0, 0) }, nil)
}
func (u *transform) AppendProgram(p *Program, r bool) error { // rows of tInv; MulPosition uses w = 1 (cpu_evaluators.go:497)
	m := u.tInv.Array() // row-major 4x4
	return unary(p, u.s, r, func() {
		p.Header(opTransform, 4, 0, 0, 0)
		p.Chunk(m[0], m[1], m[2], m[3])
		p.Chunk(m[4], m[5], m[6], m[7])
		p.Chunk(m[8], m[9], m[10], m[11])
	}, nil)
}
func (u *translate) AppendProgram(p *Program, r bool) error {
	return unary(p, u.s, r, func() { p.Header(opTranslate, 2, 0, 0, 0); p.Chunk(u.p.X, u.p.Y, u.p.Z, 0) }, nil)
}
func (u *translate2D) AppendProgram(p *Program, r bool) error {
	return unary(p, u.s, r, func() { p.Header(opTranslate, 2, 0, 0, 0); p.Chunk(u.p.X, u.p.Y, 0, 0) }, nil)
}
func (u *rotation2D) AppendProgram(p *Program, r bool) error { // tInv applied to p, cpu_evaluators.go:1186
	m := u.tInv.Array()
	return unary(p, u.s, r, func() { p.Header(opRotate2D, 2, 0, 0, 0); p.Chunk(m[0], m[1], m[2], m[3]) }, nil)
}
func (u *offset) AppendProgram(p *Program, r bool) error {
	if err := Emit(p, u.s, r); err != nil {
		return err
	}
	p.Opf(opOffset, u.off, 0)
	return nil
}
func (u *offset2D) AppendProgram(p *Program, r bool) error {
	if err := Emit(p, u.s, r); err != nil {
		return err
	}
	p.Opf(opOffset, u.f, 0)
	return nil
}
func (u *annulus2D) AppendProgram(p *Program, r bool) error {
	if err := Emit(p, u.s, r); err != nil {
		return err
	}
	p.Opf(opAnnulus, u.r, 0)
	return nil
}
func (u *twist) AppendProgram(p *Program, r bool) error {
	return unary(p, u.s, r, func() { p.Opf(opTwist, u.k, 0) }, nil)
}
func (u *elongate) AppendProgram(p *Program, r bool) error { // h := Scale(0.5, e.h), cpu_evaluators.go:412
	return unary(p, u.s, r, func() {
		p.Header(opElongate, 2, 0, 0, 0)
		p.Chunk(0.5*u.h.X, 0.5*u.h.Y, 0.5*u.h.Z, 0)
		p.PushD()
	}, func() { p.Op0(opAddBelow); p.PopD() })
}
func (u *elongate2D) AppendProgram(p *Program, r bool) error {
	return unary(p, u.s, r, func() { p.Opf(opElongate2D, 0.5*u.h.X, 0.5*u.h.Y); p.PushD() }, func() { p.Op0(opAddBelow); p.PopD() })
}

// array / array2D: 8 (4) child evaluations, min-reduced (cpu_evaluators.go:363-396, 931-960).
func (u *array) AppendProgram(p *Program, _ bool) error {
	p.PushP()
	for v := 0; v < 8; v++ {
		if v > 0 {
			p.Op0(opPeekPos)
		}
		p.Header(opArrayVar, 3, uint32(v), 0, 0)
		p.Chunk(u.d.X, u.d.Y, u.d.Z, 0)
		p.Chunk(float32(u.nx)+-1, float32(u.ny)+-1, float32(u.nz)+-1, 0)
		if err := Emit(p, u.s, false); err != nil {
			return err
		}
		if v > 0 {
			p.Op0(opMin)
			p.PopD()
		}
	}
	p.PopP()
	return nil
}
func (u *array2D) AppendProgram(p *Program, _ bool) error {
	p.PushP()
	for v := 0; v < 4; v++ {
		if v > 0 {
			p.Op0(opPeekPos)
		}
		p.Header(opArray2DVar, 2, uint32(v), 0, 0)
		p.Chunk(u.d.X, u.d.Y, float32(u.nx)+-1, float32(u.ny)+-1)
		if err := Emit(p, u.s, false); err != nil {
			return err
		}
		if v > 0 {
			p.Op0(opMin)
			p.PopD()
		}
	}
	p.PopP()
	return nil
}

// circarray / circarray2D: the child is evaluated at the two neighbouring sector positions (cpu_evaluators.go:1056-1090).
func circ(p *Program, child glbuild.Shader, nInst, circleDiv int, restore bool) error {
	ncirc := float32(circleDiv)
	angle := 2 * math.Pi / ncirc
	if restore {
		p.PushP()
	}
	p.Header(opCircEnter, 2, 0, 0, 0)
	p.Chunk(angle, ncirc, float32(nInst-1), 0)
	p.p++ // CIRC_ENTER parks p0 on the position stack
	if p.p > p.pmax {
		p.pmax = p.p
	}
	if err := Emit(p, child, false); err != nil { // evaluated at p1 first
		return err
	}
	p.PopP() // p = p0
	if err := Emit(p, child, false); err != nil {
		return err
	}
	p.Op0(opMin)
	p.PopD()
	if restore {
		p.PopP()
	}
	return nil
}
func (u *circarray) AppendProgram(p *Program, r bool) error   { return circ(p, u.s, u.nInst, u.circleDiv, r) }
func (u *circarray2D) AppendProgram(p *Program, r bool) error { return circ(p, u.s, u.nInst, u.circleDiv, r) }

// ---------------------------------------------------------------------------------------------- 2D -> 3D

func (u *extrusion) AppendProgram(p *Program, r bool) error { // h := e.h / 2, cpu_evaluators.go:524
	p.Header(opExtrudeEnter, 1, 0, fbits(u.h/2), 0)
	p.PushD()
	if err := Emit(p, u.s, r); err != nil {
		return err
	}
	p.Op0(opExtrudeExit)
	p.PopD()
	return nil
}
func (u *revolution) AppendProgram(p *Program, r bool) error {
	return unary(p, u.s2d, r, func() { p.Opf(opRevolve, u.off, 0) }, nil)
}

// ---------------------------------------------------------------------------------------------- 2D primitives

func (c *circle2D) AppendProgram(p *Program, _ bool) error { p.Opf(opCircle2D, c.r, 0); p.PushD(); return nil }
func (c *rect2D) AppendProgram(p *Program, _ bool) error {
	p.Opf(opRect2D, 0.5*c.d.X, 0.5*c.d.Y)
	p.PushD()
	return nil
}
func (c *line2D) AppendProgram(p *Program, _ bool) error { // cpu_evaluators.go:552-555
	ba := ms2.Sub(c.b, c.a)
	p.Header(opLine2D, 3, 0, 0, 0)
	p.Chunk(c.a.X, c.a.Y, ba.X, ba.Y)
	p.Chunk(ba.X*ba.X+ba.Y*ba.Y, c.width/2, 0, 0)
	p.PushD()
	return nil
}
func (c *lines2D) AppendProgram(p *Program, _ bool) error {
	off := uint32(len(p.Aux))
	for _, seg := range c.points {
		p.Aux = append(p.Aux, seg[0].X, seg[0].Y, seg[1].X, seg[1].Y)
	}
	p.Header(opLines2D, 1, off, uint32(len(c.points)), fbits(c.width/2))
	p.PushD()
	return nil
}
func (c *arc2D) AppendProgram(p *Program, _ bool) error { // cpu_evaluators.go:565-569
	s, cs := math.Sincos(c.angle / 2)
	p.Header(opArc2D, 3, 0, 0, 0)
	p.Chunk(c.radius, c.thick/2, s, cs)
	p.Chunk(c.radius*s, c.radius*cs, 0, 0)
	p.PushD()
	return nil
}
func (c *equilateralTri2d) AppendProgram(p *Program, _ bool) error { // cpu_evaluators.go:670-671
	r := c.hTri / sqrt3
	p.Opf(opEqTri2D, r, r/sqrt3)
	p.PushD()
	return nil
}
func (c *hex2D) AppendProgram(p *Program, _ bool) error { p.Opf(opHex2D, c.side, 0.577350269*c.side); p.PushD(); return nil }
func (c *oct2D) AppendProgram(p *Program, _ bool) error { p.Opf(opOct2D, c.c, 0.4142135623*c.c); p.PushD(); return nil }
func (c *diamond) AppendProgram(p *Program, _ bool) error {
	bx, by := 0.5*c.d.X, 0.5*c.d.Y
	p.Header(opDiamond2D, 2, 0, 0, 0)
	p.Chunk(bx, by, bx*bx+by*by, 0)
	p.PushD()
	return nil
}
func (c *x2d) AppendProgram(p *Program, _ bool) error     { p.Opf(opRoundX2D, c.dim, c.thick); p.PushD(); return nil }
func (c *ellipse2D) AppendProgram(p *Program, _ bool) error { p.Opf(opEllipse2D, c.a, c.b); p.PushD(); return nil }

// poly2D: 8 floats per edge, edge i runs v1 = verts[i], v2 = verts[i-1] (cpu_evaluators.go:793-818).
func (c *poly2D) AppendProgram(p *Program, _ bool) error {
	if len(c.vert) < 3 {
		return errors.New("polygon needs at least 3 vertices")
	}
	off := p.alignAux(4)
	j := len(c.vert) - 1
	for i, v1 := range c.vert {
		v2 := c.vert[j]
		e := ms2.Sub(v2, v1)
		p.Aux = append(p.Aux, v1.X, v1.Y, e.X, e.Y, e.X*e.X+e.Y*e.Y, v2.Y, 0, 0)
		j = i
	}
	p.Header(opPoly2D, 1, off, uint32(len(c.vert)), 0)
	p.PushD()
	return nil
}

// quadbezier2d: per-shape constants of cpu_evaluators.go:583-593.
func (c *quadbezier2d) AppendProgram(p *Program, _ bool) error {
	A, B, C := c.a, c.b, c.c
	a := ms2.Sub(B, A)
	a2 := a.X*a.X + a.Y*a.Y
	bx, by := A.X+(C.X-2*B.X), A.Y+(C.Y-2*B.Y)
	cx, cy := 2*a.X, 2*a.Y
	kk := 1. / (bx*bx + by*by)
	kx := kk * (a.X*bx + a.Y*by)
	p.Header(opBezierQ2D, 4, 0, fbits(c.thick/2), 0)
	p.Chunk(A.X, A.Y, a.X, a.Y)
	p.Chunk(bx, by, cx, cy)
	p.Chunk(kk, kx, kx*kx, a2)
	p.PushD()
	return nil
}

// translateMulti2D: min over displaced copies (cpu_evaluators.go:1167-1182).
func (u *translateMulti2D) AppendProgram(p *Program, _ bool) error {
	p.PushP()
	for k, d := range u.displacements {
		if k > 0 {
			p.Op0(opPeekPos)
		}
		p.Header(opTranslate, 2, 0, 0, 0)
		p.Chunk(d.X, d.Y, 0, 0)
		if err := Emit(p, u.s, false); err != nil {
			return err
		}
		if k > 0 {
			p.Op0(opMin)
			p.PopD()
		}
	}
	p.PopP()
	return nil
}
