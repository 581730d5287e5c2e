package threads

import (
	math "github.com/chewxy/math32"
	"github.com/soypat/gsdf/glbuild"
)

// AppendProgram emits the screw node for the CUDA interpreter (include/gsdf_program.h, GSDF_OP_SCREW_ENTER):
// push(|z| - L/2); p = (sawTooth(z + lead*atan2(y,x)/2pi, pitch), hypot(x,y) + z*tan(taper)); child; top = max(top, below)
// exactly as screw.Evaluate does (threads.go:141-181; the CPU path uses Tan(taper), :155).
func (s *screw) AppendProgram(p *glbuild.Program, restore bool) error {
	g := p.TakeGuard() // slab guard handed down by a difference / union / smooth union (glbuild/cuda_program.go)
	if restore {
		p.PushP()
	}
	hw := len(p.Chunks)
	p.Header(glbuild.OpScrewEnter, 2, 0, 0, math.Float32bits(g.K))
	p.Chunk(s.pitch, s.lead, s.lengthDiv2, math.Tan(s.taper))
	p.PushD()
	if err := glbuild.Emit(p, s.thread, false); err != nil {
		return err
	}
	p.Op0(glbuild.OpMaxBelow)
	p.PopD()
	p.PatchGuard(hw, g) // the skip lands right behind MAX_BELOW; the restoring POP_POS below still runs
	if restore {
		p.PopP()
	}
	return nil
}

// SlabBounded marks the screw as a node whose value is >= |z| - L/2 in its own frame (threads.go:176-180).
func (s *screw) SlabBounded() {}
