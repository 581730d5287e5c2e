package glbuild

// CUDA node program: the packed instruction stream of include/gsdf_program.h that libgsdfb200.so interprets on the GPU,
// and the tree walk that produces it. The Program type and the emitter interface live in glbuild (like Programmer, the
// GLSL code generator they stand beside, glbuild.go:175) so that gleval can flatten a tree without importing package
// gsdf, which itself imports gleval; the per-node emitters live next to the node types (gsdf/cuda_flatten.go,
// forge/threads/cuda_flatten.go).

import (
	"errors"
	"fmt"
	"os"
	"unsafe"

	math "github.com/chewxy/math32"
)

// Opcodes: enum gsdf_opcode of include/gsdf_program.h, in declaration order.
const (
	OpEnd uint32 = iota
	OpSphere
	OpBox
	OpBoxFrame
	OpTorus
	OpCylinder
	OpHex
	OpCircle2D
	OpRect2D
	OpLine2D
	OpLines2D
	OpArc2D
	OpEqTri2D
	OpHex2D
	OpOct2D
	OpDiamond2D
	OpRoundX2D
	OpPoly2D
	OpEllipse2D
	OpBezierQ2D
	OpMin
	OpMax
	OpDiff
	OpXor
	OpSmoothUnion
	OpSmoothDiff
	OpSmoothIntersect
	OpOffset
	OpAnnulus
	OpMulDist
	OpShellExit
	OpAddBelow
	OpExtrudeExit
	OpMaxBelow
	OpPushPos
	OpPopPos
	OpPeekPos
	OpTranslate
	OpScalePos
	OpSymmetry
	OpTransform
	OpRotate2D
	OpTwist
	OpElongate
	OpElongate2D
	OpArrayVar
	OpArray2DVar
	OpCircEnter
	OpExtrudeEnter
	OpRevolve
	OpScrewEnter
	OpCullUB2D    // box guards: upper bound of a 2-D union from anchor points on the operands' outlines
	OpBBoxGuard2D // box guards: skip the operand that follows when the tile is farther from its box than the running value
	OpMinConst    // top = min(w2, top): accumulator seed of the array folds (largenum, math.MaxFloat32)
)

const (
	programMagic   = 0x46445347 // "GSDF"
	programVersion = 1
	polyEdgeFloats = 8
)

// Slab guards (include/gsdf_program.h, "slab guards"; gsdf_b200/csrc/host/flatten.cpp is the executable specification).
// An extrusion or screw returns a value >= w = |z| - h/2 in its own frame, and its ENTER op has w in hand before any of the
// 2-D work below it. When such a node is the LATER operand of a difference, union or smooth union -- directly or through
// nodes that only move p -- the combiner hands it a guard: if the combiner's result cannot depend on any value >= w for
// every point of the CTA's tile, the interpreter jumps behind the node's exit op. Results are bit-identical either way.
const (
	GuardNone uint32 = iota
	GuardDiff        // a = top:  -w < a             =>  max(a, -s) == a
	GuardMin         // a = top:   w > a             =>  min(a, s)  == a
	GuardSmoothUnion // a = top:   w - a >= k, a != 0 =>  the blend returns a
)

// Guard is what a combiner hands down to a later operand.
type Guard struct {
	Kind uint32
	K    float32 // smooth-union radius
}

// SlabBoundedNode is implemented by the node types whose value is bounded below by |z| - h/2 of their own frame and whose
// ENTER op evaluates the guard: gsdf.extrusion and threads.screw.
type SlabBoundedNode interface{ SlabBounded() }

// DistTransparentNode is implemented by the node types whose Evaluate forwards the child's distance unchanged and only
// moves p (translate, transform, symmetry, twist): a guard passes through them to their child.
type DistTransparentNode interface{ DistTransparentChild() Shader }

// guardsOff mirrors the C++ flattener's A/B switch.
var guardsOff = func() bool { e := os.Getenv("GSDF_NO_GUARDS"); return e != "" && e[0] != '0' }()

func unwrapAll(s Shader) Shader {
	for {
		u := Unwrap(s)
		if u == nil {
			return s
		}
		s = u
	}
}

// Guardable reports whether s is an extrusion / screw reached through distance-transparent nodes (and wrappers) only.
func Guardable(s Shader) bool {
	for s != nil {
		s = unwrapAll(s)
		if _, ok := s.(SlabBoundedNode); ok {
			return true
		}
		t, ok := s.(DistTransparentNode)
		if !ok {
			return false
		}
		s = t.DistTransparentChild()
	}
	return false
}

// Program accumulates the instruction chunks (4 x uint32 each) and the float side buffer.
type Program struct {
	Chunks       []uint32
	Aux          []float32
	Dim          int
	D, P         int // current distance / position stack depth
	Dmax, Pmax   int
	ninstr       int
	pending      Guard // set by a combiner right before it emits a guardable operand, taken by that operand's ENTER op
}

// GuardFor returns the guard a combiner of the given kind may hand to `child` (none if the child is not guardable).
func (p *Program) GuardFor(child Shader, kind uint32, k float32) Guard {
	if p.Dim == 3 && !guardsOff && Guardable(child) {
		return Guard{Kind: kind, K: k}
	}
	return Guard{}
}

// SetGuard arms the guard for the operand emitted next. Guardable(child) guarantees that the chain of emitters below
// consists of distance-transparent nodes, which leave it alone, and ends in a SlabBoundedNode, which takes it.
func (p *Program) SetGuard(g Guard) { p.pending = g }

// TakeGuard is called by extrusion / screw at the start of their emitter.
func (p *Program) TakeGuard() Guard { g := p.pending; p.pending = Guard{}; return g }

// PatchGuard stores kind | target<<8 into word 1 of the ENTER header at chunk word index hw; the target is the chunk
// right behind the node's own exit op (the restoring POP_POS ops of its wrappers follow and still run).
func (p *Program) PatchGuard(hw int, g Guard) {
	if g.Kind != GuardNone {
		p.Chunks[hw+1] = g.Kind | uint32(len(p.Chunks)/4)<<8
	}
}

// ProgramEmitter is implemented by every node type of gsdf and forge/threads (their cuda_flatten.go files).
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
	p.D++
	if p.D > p.Dmax {
		p.Dmax = p.D
	}
}
func (p *Program) PopD() { p.D-- }
func (p *Program) PushP() {
	p.Op0(OpPushPos)
	p.P++
	if p.P > p.Pmax {
		p.Pmax = p.P
	}
}
func (p *Program) PopP() { p.Op0(OpPopPos); p.P-- }
// BoxGuard emits BBOX_GUARD2D for an operand whose bounding box (in the current frame) is [minx,miny]-[maxx,maxy] and returns
// the chunk word index of its header for PatchBoxGuard (flatten.cpp: boxGuard / patchBoxGuard).
func (p *Program) BoxGuard(kind uint32, margin, minx, miny, maxx, maxy float32) int {
	hw := len(p.Chunks)
	p.Header(OpBBoxGuard2D, 2, kind, 0, fbits(margin))
	p.Chunk(minx, miny, maxx, maxy)
	return hw
}

// PatchBoxGuard makes the guard jump to the chunk that comes next: the operand's combiner, which then keeps the running value.
func (p *Program) PatchBoxGuard(hw int, kind uint32) { p.Chunks[hw+1] = kind | uint32(len(p.Chunks)/4)<<8 }

// BoxGuardsOn mirrors the C++ flattener: box guards are emitted unless GSDF_NO_GUARDS is set.
func BoxGuardsOn() bool { return !guardsOff }

func (p *Program) AlignAux(n int) uint32 {
	for len(p.Aux)%n != 0 {
		p.Aux = append(p.Aux, 0)
	}
	return uint32(len(p.Aux))
}

// Emit sees through the glbuild wrappers, then dispatches to the node's emitter.
func Emit(p *Program, s Shader, restore bool) error {
	for {
		u := Unwrap(s)
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
func (p *Program) Unary(child Shader, restore bool, enter, exit func()) error {
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

func (p *Program) Binary(a, b Shader, restore bool, emitOp func()) error {
	return p.BinaryGuarded(a, b, restore, GuardNone, 0, emitOp)
}

// BinaryGuarded is Binary for the combiners that can guard their second operand (difference, smooth union).
func (p *Program) BinaryGuarded(a, b Shader, restore bool, kind uint32, k float32, emitOp func()) error {
	if err := Emit(p, a, true); err != nil { // the first operand must leave p intact for the second
		return err
	}
	if kind != GuardNone {
		p.SetGuard(p.GuardFor(b, kind, k))
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
func Flatten3(root Shader3D) (blob []byte, aux []float32, err error) { return flatten(root, 3) }
func Flatten2(root Shader2D) (blob []byte, aux []float32, err error) { return flatten(root, 2) }

func flatten(root Shader, dim int) ([]byte, []float32, error) {
	var p Program
	p.Dim = dim
	if err := Emit(&p, root, false); err != nil {
		return nil, nil, err
	}
	p.Header(OpEnd, 1, 0, 0, 0)
	if p.D != 1 {
		return nil, nil, errors.New("internal: distance stack imbalance")
	}
	dstack := 1 // the top is cached in a register; slot 0 also absorbs the first push
	if p.Dmax > 1 {
		dstack = p.Dmax - 1
	}
	p.AlignAux(4)
	hdr := [8]uint32{programMagic, programVersion, uint32(len(p.Chunks) / 4), uint32(dim), uint32(dstack), uint32(p.Pmax), uint32(p.ninstr), 0}
	words := append(hdr[:], p.Chunks...)
	blob := unsafe.Slice((*byte)(unsafe.Pointer(&words[0])), 4*len(words)) // little-endian hosts, like the C side
	return append([]byte(nil), blob...), p.Aux, nil
}

