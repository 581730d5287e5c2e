package threads

import (
	math "github.com/chewxy/math32"
	"github.com/soypat/gsdf/glbuild"
)

// AppendProgram emits the screw node for the CUDA interpreter (include/gsdf_program.h, GSDF_OP_SCREW_ENTER):
// push(|z| - L/2); p = (sawTooth(z + lead*atan2(y,x)/2pi, pitch), hypot(x,y) + z*tan(taper)); child; top = max(top, below)
// exactly as screw.Evaluate does (threads.go:141-181; the CPU path uses Tan(taper), :155).
func (s *screw) AppendProgram(p *glbuild.Program, restore bool) error {
	if restore {
		p.PushP()
	}
	p.Header(glbuild.OpScrewEnter, 2, 0, 0, 0)
	p.Chunk(s.pitch, s.lead, s.lengthDiv2, math.Tan(s.taper))
	p.PushD()
	if err := glbuild.Emit(p, s.thread, false); err != nil {
		return err
	}
	p.Op0(glbuild.OpMaxBelow)
	p.PopD()
	if restore {
		p.PopP()
	}
	return nil
}
