//go:build cgo && cuda

package gleval

/*
#cgo LDFLAGS: -lgsdfb200
#include "gsdf_b200.h"
*/
import "C"

import (
	"errors"
	"unsafe"

	"github.com/soypat/geometry/ms2"
	"github.com/soypat/geometry/ms3"
	"github.com/soypat/gsdf/glbuild"
)

func cudaErr() error { return errors.New(C.GoString(C.gsdf_last_error())) }

// InitCUDA selects the device later handles are created on (one process or OS thread per GPU). It takes the place of
// Init1x1GLFW (gpu.go:21); no OS-thread pinning is needed.
func InitCUDA(device int) error {
	if C.gsdf_set_device(C.int(device)) != 0 {
		return cudaErr()
	}
	return nil
}

// HostAlloc returns n float32 in page-locked host memory (gsdf_host_alloc): Evaluate and the renderers' read-backs move data
// from / into such slices by DMA directly instead of through the library's staging buffers. Free with HostFree.
func HostAlloc(n int) []float32 {
	p := C.gsdf_host_alloc(C.size_t(4 * n))
	if p == nil {
		return nil
	}
	return unsafe.Slice((*float32)(p), n)
}
func HostFree(s []float32) {
	if len(s) > 0 {
		C.gsdf_host_free(unsafe.Pointer(&s[0]))
	}
}

// createProgramOn uploads to an explicitly named device: goroutines migrate between OS threads, so the per-thread default
// of gsdf_set_device is not something Go code should lean on when it drives several GPUs.
func createProgramOn(device int, blob []byte, aux []float32) (*C.gsdf_program, error) {
	var h *C.gsdf_program
	var auxp *C.float
	if len(aux) > 0 {
		auxp = (*C.float)(unsafe.Pointer(&aux[0]))
	}
	if rc := C.gsdf_program_create_on(C.int(device), unsafe.Pointer(&blob[0]), C.size_t(len(blob)), auxp, C.size_t(len(aux)), &h); rc != 0 {
		return nil, cudaErr()
	}
	return h, nil
}

// NewCUDASDF3On is NewCUDASDF3 on the given device.
func NewCUDASDF3On(device int, root glbuild.Shader3D) (*SDF3CUDA, error) {
	blob, aux, err := glbuild.Flatten3(root)
	if err != nil {
		return nil, err
	}
	h, err := createProgramOn(device, blob, aux)
	if err != nil {
		return nil, err
	}
	return &SDF3CUDA{h: h, bb: root.Bounds()}, nil
}

func createProgram(blob []byte, aux []float32) (*C.gsdf_program, error) {
	var h *C.gsdf_program
	var auxp *C.float
	if len(aux) > 0 {
		auxp = (*C.float)(unsafe.Pointer(&aux[0]))
	}
	if rc := C.gsdf_program_create(unsafe.Pointer(&blob[0]), C.size_t(len(blob)), auxp, C.size_t(len(aux)), &h); rc != 0 {
		return nil, cudaErr()
	}
	return h, nil
}

// SDF3CUDA implements SDF3 (gleval.go:15-24) on the CUDA backend, next to SDF3Compute (gpu.go:56).
type SDF3CUDA struct {
	h     *C.gsdf_program
	bb    ms3.Box
	evals uint64
}

// NewCUDASDF3 mirrors NewComputeGPUSDF3(source, bb, cfg) (gpu.go:35): the tree is flattened once and uploaded once.
func NewCUDASDF3(root glbuild.Shader3D) (*SDF3CUDA, error) {
	blob, aux, err := glbuild.Flatten3(root)
	if err != nil {
		return nil, err
	}
	h, err := createProgram(blob, aux)
	if err != nil {
		return nil, err
	}
	return &SDF3CUDA{h: h, bb: root.Bounds()}, nil
}

// Evaluate implements SDF3. userData is ignored, like the GL evaluator does (gpu.go:82).
func (s *SDF3CUDA) Evaluate(pos []ms3.Vec, dist []float32, userData any) error {
	if len(pos) != len(dist) {
		return errMismatchBufferLength // gleval.go:48: checked before the cgo call so the sentinel is exact
	} else if len(dist) == 0 {
		return errEmptyBuffers // gleval.go:47
	}
	// []ms3.Vec is pointer-free (3 x float32): passing &pos[0] is legal under the cgo rules, as gpu_cgo.go:166 does.
	if rc := C.gsdf_eval3(s.h, (*C.float)(unsafe.Pointer(&pos[0])), (*C.float)(unsafe.Pointer(&dist[0])), C.size_t(len(pos))); rc != 0 {
		return cudaErr()
	}
	s.evals += uint64(len(pos))
	return nil
}
func (s *SDF3CUDA) Bounds() ms3.Box     { return s.bb }
func (s *SDF3CUDA) Evaluations() uint64 { return s.evals } // asserted unchecked by gsdfaux.go:219
func (s *SDF3CUDA) Close()              { C.gsdf_program_destroy(s.h); s.h = nil }

// Handle exposes the program to package glrender (the mesher and the dual-contour renderer run on the device).
func (s *SDF3CUDA) Handle() unsafe.Pointer { return unsafe.Pointer(s.h) }

// Update re-flattens root into this evaluator's device buffers (an edited tree costs one small upload; the GL path
// recompiles its shader instead, gpu.go:35-54).
func (s *SDF3CUDA) Update(root glbuild.Shader3D) error {
	blob, aux, err := glbuild.Flatten3(root)
	if err != nil {
		return err
	}
	var auxp *C.float
	if len(aux) > 0 {
		auxp = (*C.float)(unsafe.Pointer(&aux[0]))
	}
	if rc := C.gsdf_program_update(s.h, unsafe.Pointer(&blob[0]), C.size_t(len(blob)), auxp, C.size_t(len(aux))); rc != 0 {
		return cudaErr()
	}
	s.bb = root.Bounds()
	return nil
}

// Specialize compiles kernels for this tree's instruction stream (NVRTC at run time) -- the step the GL path performs when
// it compiles the tree's compute shader (gpu.go:35-54). Renderers built on the evaluator then use them for the lattice
// evaluation and the prune-centre passes: bit-identical distances, about a quarter less evaluation time. It reports false
// (and no error) where run-time compilation is unavailable: the interpreter kernels keep running.
func (s *SDF3CUDA) Specialize() (bool, error) {
	switch rc := C.gsdf_program_specialize(s.h); rc {
	case 0:
		return true, nil
	case C.GSDF_EUNSUPPORTED:
		return false, nil
	default:
		return false, cudaErr()
	}
}

// Specialized reports whether the run-time compiled kernels are in use for the tree's current structure (an Update that
// changes the structure drops them; parameters alone do not).
func (s *SDF3CUDA) Specialized() bool { return C.gsdf_program_is_specialized(s.h) != 0 }

// SDF2CUDA implements SDF2 (gleval.go:28-37).
type SDF2CUDA struct {
	h     *C.gsdf_program
	bb    ms2.Box
	evals uint64
}

func NewCUDASDF2(root glbuild.Shader2D) (*SDF2CUDA, error) {
	blob, aux, err := glbuild.Flatten2(root)
	if err != nil {
		return nil, err
	}
	h, err := createProgram(blob, aux)
	if err != nil {
		return nil, err
	}
	return &SDF2CUDA{h: h, bb: root.Bounds()}, nil
}

func (s *SDF2CUDA) Evaluate(pos []ms2.Vec, dist []float32, userData any) error {
	if len(pos) != len(dist) {
		return errMismatchBufferLength
	} else if len(dist) == 0 {
		return errEmptyBuffers
	}
	if rc := C.gsdf_eval2(s.h, (*C.float)(unsafe.Pointer(&pos[0])), (*C.float)(unsafe.Pointer(&dist[0])), C.size_t(len(pos))); rc != 0 {
		return cudaErr()
	}
	s.evals += uint64(len(pos))
	return nil
}
func (s *SDF2CUDA) Bounds() ms2.Box        { return s.bb }
func (s *SDF2CUDA) Evaluations() uint64    { return s.evals }
func (s *SDF2CUDA) Close()                 { C.gsdf_program_destroy(s.h); s.h = nil }
func (s *SDF2CUDA) Handle() unsafe.Pointer { return unsafe.Pointer(s.h) }
