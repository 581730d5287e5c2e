//go:build cgo && cuda

package glrender

/*
#cgo LDFLAGS: -lgsdfb200
#include "gsdf_b200.h"
*/
import "C"

import (
	"errors"
	"image"
	"io"
	"unsafe"

	"github.com/soypat/geometry/ms3"
	"github.com/soypat/gsdf/glbuild"
	"github.com/soypat/gsdf/gleval"
)

func cudaErr() error { return errors.New(C.GoString(C.gsdf_last_error())) }

// MesherCUDA implements Renderer (glrender.go:11-13): octree-pruned or dense marching cubes on the device.
type MesherCUDA struct{ h *C.gsdf_mesher }

// NewCUDARenderer mirrors NewOctreeRenderer(s, res, evalBufferSize) (octreerenderer.go:45); prune=false gives
// FlatRenderer semantics (flatrenderer.go:37). cz0, cz1 select a Z-slab of cell layers (0, 0 = the whole lattice): one
// MesherCUDA per GPU / per pipeline stage, triangles appended in slab order. The slab is meshed inside this call.
func NewCUDARenderer(s *gleval.SDF3CUDA, res float32, prune bool, cz0, cz1 int) (*MesherCUDA, error) {
	bb := s.Bounds()
	var lat C.gsdf_lattice
	if rc := C.gsdf_lattice_from_bounds((*C.float)(unsafe.Pointer(&bb.Min)), (*C.float)(unsafe.Pointer(&bb.Max)), C.float(res), &lat); rc != 0 {
		return nil, cudaErr() // "resolution not fine enough for marching cubes" (flatrenderer.go:53)
	}
	if cz0 == 0 && cz1 == 0 {
		cz1 = int(lat.n[2])
	}
	flags := C.uint(0)
	if prune {
		flags = C.GSDF_MESH_PRUNE
	}
	var h *C.gsdf_mesher
	if rc := C.gsdf_mesh_begin((*C.gsdf_program)(s.Handle()), &lat, C.int(cz0), C.int(cz1), flags, &h); rc != 0 {
		return nil, cudaErr()
	}
	return &MesherCUDA{h: h}, nil
}

// PruneLevel is one level of a coarse-to-fine prune plan (include/gsdf_b200.h, gsdf_prune_plan): cubes of 2^(Level-1) cells
// are kept iff |d(centre)| < Margin * size * sqrt3/2. Margin 1 is octreePrunea's literal rule (octreerenderer.go:240-284).
type PruneLevel struct {
	Level  int
	Margin float32
}

// NewCUDARendererPlan is NewCUDARenderer with an explicit prune plan (coarse to fine, ending with level 3 or with levels 3, 2:
// the 2-cell level decides which corners inside kept 4-cell blocks are evaluated at all). The default
// plan of NewCUDARenderer is level 3 with margin 1.25 (plus coarse levels on large lattices): it reproduces the dense sweep
// on every example scene, README.md:152's 309,872 included; []PruneLevel{{3, 1}} is the literal rule at every level-3 cube.
func NewCUDARendererPlan(s *gleval.SDF3CUDA, res float32, plan []PruneLevel, cz0, cz1 int) (*MesherCUDA, error) {
	bb := s.Bounds()
	var lat C.gsdf_lattice
	if rc := C.gsdf_lattice_from_bounds((*C.float)(unsafe.Pointer(&bb.Min)), (*C.float)(unsafe.Pointer(&bb.Max)), C.float(res), &lat); rc != 0 {
		return nil, cudaErr()
	}
	if cz0 == 0 && cz1 == 0 {
		cz1 = int(lat.n[2])
	}
	if len(plan) == 0 || len(plan) > C.GSDF_PRUNE_MAX_LEVELS {
		return nil, errors.New("prune plan needs 1 to 4 levels")
	}
	var cp C.gsdf_prune_plan
	cp.nlevels = C.int32_t(len(plan))
	for i, l := range plan {
		cp.level[i] = C.int32_t(l.Level)
		cp.margin[i] = C.float(l.Margin)
	}
	var h *C.gsdf_mesher
	if rc := C.gsdf_mesh_begin_plan((*C.gsdf_program)(s.Handle()), &lat, C.int(cz0), C.int(cz1), C.GSDF_MESH_PRUNE, &cp, &h); rc != 0 {
		return nil, cudaErr()
	}
	return &MesherCUDA{h: h}, nil
}

// MultiMesherCUDA implements Renderer over several GPUs from ONE process (gsdf_multi_*): the device analogue of
// FlatRenderer.evalGrid's split of the corner planes over goroutines (flatrenderer.go:103-141). Z-slabs are dealt round-robin
// to the devices; every device has its own copy of the program, one library-owned host thread and pinned staging; the
// slabs' triangles are concatenated in slab order, which is FlatRenderer cell order. No goroutine or OS-thread pinning is
// needed on the Go side: the devices are named explicitly and the handle may be used from any goroutine (one at a time).
type MultiMesherCUDA struct{ h *C.gsdf_multimesher }

// NewCUDAMultiRenderer flattens root once and meshes its lattice on `devices` with slabsPerDevice Z-slabs each (1 device
// with 3 slabs pipelines the read-back of one GPU). The lattice is rendered once inside this call.
func NewCUDAMultiRenderer(root glbuild.Shader3D, res float32, devices []int, slabsPerDevice int, prune bool) (*MultiMesherCUDA, error) {
	blob, aux, err := glbuild.Flatten3(root)
	if err != nil {
		return nil, err
	}
	bb := root.Bounds()
	var lat C.gsdf_lattice
	if rc := C.gsdf_lattice_from_bounds((*C.float)(unsafe.Pointer(&bb.Min)), (*C.float)(unsafe.Pointer(&bb.Max)), C.float(res), &lat); rc != 0 {
		return nil, cudaErr()
	}
	devs := make([]C.int, len(devices))
	for i, d := range devices {
		devs[i] = C.int(d)
	}
	if len(devs) == 0 {
		return nil, errors.New("no devices")
	}
	flags := C.uint(0)
	if prune {
		flags = C.GSDF_MESH_PRUNE
	}
	var auxp *C.float
	if len(aux) > 0 {
		auxp = (*C.float)(unsafe.Pointer(&aux[0]))
	}
	var h *C.gsdf_multimesher
	if rc := C.gsdf_multi_begin(C.int(len(devs)), &devs[0], C.int(slabsPerDevice), unsafe.Pointer(&blob[0]), C.size_t(len(blob)), auxp, C.size_t(len(aux)), &lat, flags, &h); rc != 0 {
		return nil, cudaErr()
	}
	return &MultiMesherCUDA{h: h}, nil
}

// Update hands a re-flattened tree to every device; it is uploaded at the start of the next Render.
func (m *MultiMesherCUDA) Update(root glbuild.Shader3D) error {
	blob, aux, err := glbuild.Flatten3(root)
	if err != nil {
		return err
	}
	var auxp *C.float
	if len(aux) > 0 {
		auxp = (*C.float)(unsafe.Pointer(&aux[0]))
	}
	if C.gsdf_multi_update(m.h, unsafe.Pointer(&blob[0]), C.size_t(len(blob)), auxp, C.size_t(len(aux))) != 0 {
		return cudaErr()
	}
	return nil
}

// Specialize compiles kernels for the tree's instruction stream once and uses them on every device (gleval.SDF3CUDA.Specialize).
func (m *MultiMesherCUDA) Specialize() (bool, error) {
	switch rc := C.gsdf_multi_specialize(m.h); rc {
	case 0:
		return true, nil
	case C.GSDF_EUNSUPPORTED:
		return false, nil
	default:
		return false, cudaErr()
	}
}

// Rebalance re-cuts the slabs by the evaluations each executed last (equal layers are not equal work on a pruned lattice).
func (m *MultiMesherCUDA) Rebalance(rounds int) error {
	if C.gsdf_multi_rebalance(m.h, C.int(rounds)) < 0 {
		return cudaErr()
	}
	return nil
}

// RenderAll re-renders the lattice and returns every triangle in FlatRenderer order. dst is reused when it is large enough;
// a slice over memory from gleval.HostAlloc is filled by DMA directly, any other slice through the library's pinned staging.
func (m *MultiMesherCUDA) RenderAll(dst []ms3.Triangle) ([]ms3.Triangle, error) {
	var t C.uint64_t
	C.gsdf_multi_stats(m.h, nil, nil, &t, nil)
	if cap(dst) < int(t)+8 {
		dst = make([]ms3.Triangle, int(t)+int(t)/8+8)
	}
	dst = dst[:cap(dst)]
	n := C.gsdf_multi_render(m.h, (*C.float)(unsafe.Pointer(&dst[0])), C.size_t(len(dst)))
	if n == C.GSDF_ESHORT { // the tree changed and produced more: the count is known now
		C.gsdf_multi_stats(m.h, nil, nil, &t, nil)
		dst = make([]ms3.Triangle, int(t)+int(t)/8+8)
		n = C.gsdf_multi_render(m.h, (*C.float)(unsafe.Pointer(&dst[0])), C.size_t(len(dst)))
	}
	if n < 0 {
		return nil, cudaErr()
	}
	return dst[:n], nil
}

// ReadTriangles implements Renderer on the last render.
func (m *MultiMesherCUDA) ReadTriangles(dst []ms3.Triangle, userData any) (int, error) {
	if len(dst) < 5 {
		return 0, io.ErrShortBuffer
	}
	n := C.gsdf_multi_read(m.h, (*C.float)(unsafe.Pointer(&dst[0])), C.size_t(len(dst)))
	switch {
	case n < 0:
		return 0, cudaErr()
	case n == 0:
		return 0, io.EOF
	}
	return int(n), nil
}

func (m *MultiMesherCUDA) Evaluations() uint64 {
	var e C.uint64_t
	C.gsdf_multi_stats(m.h, &e, nil, nil, nil)
	return uint64(e)
}
func (m *MultiMesherCUDA) TotalPruned() uint64 {
	var p C.uint64_t
	C.gsdf_multi_stats(m.h, nil, &p, nil, nil)
	return uint64(p)
}

// WriteBinarySTL packs every slab's records on its device and does ONE Write.
func (m *MultiMesherCUDA) WriteBinarySTL(w io.Writer) (int, error) {
	var t C.uint64_t
	C.gsdf_multi_stats(m.h, nil, nil, &t, nil)
	buf := make([]byte, 84+50*int(t))
	if n := C.gsdf_multi_stl(m.h, unsafe.Pointer(&buf[0]), C.size_t(len(buf))); n < 0 {
		return 0, cudaErr()
	}
	return w.Write(buf)
}
func (m *MultiMesherCUDA) Close() { C.gsdf_multi_destroy(m.h); m.h = nil }

// ReadTriangles implements Renderer. ms3.Triangle = [3]ms3.Vec = 9 float32.
func (m *MesherCUDA) ReadTriangles(dst []ms3.Triangle, userData any) (int, error) {
	if len(dst) < 5 {
		return 0, io.ErrShortBuffer // octreerenderer.go:132, flatrenderer.go:187
	}
	n := C.gsdf_mesh_read(m.h, (*C.float)(unsafe.Pointer(&dst[0])), C.size_t(len(dst)))
	switch {
	case n < 0:
		return 0, cudaErr()
	case n == 0:
		return 0, io.EOF // octreerenderer.go:156, flatrenderer.go:206
	}
	return int(n), nil
}

func (m *MesherCUDA) stats() (evals, pruned, tris uint64) {
	var e, p, t C.uint64_t
	C.gsdf_mesh_stats(m.h, &e, &p, &t)
	return uint64(e), uint64(p), uint64(t)
}
func (m *MesherCUDA) Evaluations() uint64 { e, _, _ := m.stats(); return e }
func (m *MesherCUDA) TotalPruned() uint64 { _, p, _ := m.stats(); return p } // gsdfaux.go:220-221
func (m *MesherCUDA) Rerun() error {
	if C.gsdf_mesh_rerun(m.h) != 0 {
		return cudaErr()
	}
	return nil
}
func (m *MesherCUDA) Close() { C.gsdf_mesh_destroy(m.h); m.h = nil }

// WriteBinarySTL packs the records on the device and does ONE Write (stl.go:53 does one per triangle).
func (m *MesherCUDA) WriteBinarySTL(w io.Writer) (int, error) {
	_, _, t := m.stats()
	buf := make([]byte, 84+50*int(t))
	if n := C.gsdf_mesh_stl(m.h, unsafe.Pointer(&buf[0]), C.size_t(len(buf))); n < 0 {
		return 0, cudaErr() // "empty triangle slice" (stl.go:16)
	}
	return w.Write(buf)
}

// CUDAColorConv is the data form of the conversions gsdfaux offers as closures (gsdfaux/color.go:21-102).
type CUDAColorConv struct{ c C.gsdf_colorconv }

func ColorConvInigoQuilez(characteristicDistance float32) (cc CUDAColorConv) {
	C.gsdf_colorconv_inigo_quilez(C.float(characteristicDistance), &cc.c)
	return cc
}
func ColorConvLinearGradient(gradientLength float32, c0, c1 [4]uint8) (cc CUDAColorConv) {
	pack := func(c [4]uint8) C.uint32_t { return C.uint32_t(c[0]) | C.uint32_t(c[1])<<8 | C.uint32_t(c[2])<<16 | C.uint32_t(c[3])<<24 }
	C.gsdf_colorconv_linear_gradient(C.float(gradientLength), pack(c0), pack(c1), &cc.c)
	return cc
}

// RenderRGBA is ImageRendererSDF2.Render (image.go:76-118) with the conversion fused into the evaluation kernel.
// conv == nil selects NewImageRendererSDF2(nil)'s black / white / red scheme (image.go:50-61).
func RenderRGBA(s *gleval.SDF2CUDA, img *image.RGBA, conv *CUDAColorConv) error {
	bb := s.Bounds()
	r := img.Bounds()
	if img.Stride != 4*r.Dx() {
		return errors.New("RenderRGBA needs a tightly packed image.RGBA")
	}
	var cp *C.gsdf_colorconv
	if conv != nil {
		cp = &conv.c
	}
	if rc := C.gsdf_image_render2((*C.gsdf_program)(s.Handle()), (*C.float)(unsafe.Pointer(&bb.Min)), (*C.float)(unsafe.Pointer(&bb.Max)),
		C.int(r.Dx()), C.int(r.Dy()), cp, (*C.uint8_t)(unsafe.Pointer(&img.Pix[0]))); rc != 0 {
		return cudaErr()
	}
	return nil
}

// DualContourCUDA is DualContourRenderer (dual_contour.go:12-218) on the device.
type DualContourCUDA struct{ h *C.gsdf_dualcontour }

// Reset mirrors DualContourRenderer.Reset(sdf, res, vertexPlacer, userData): vertexPlacer *DualContourLeastSquares maps to
// GSDF_DC_LEAST_SQUARES(_CHISELED). part / nparts (1, 2, 4, 8) split the octree by top-level octants across GPUs.
func (dcr *DualContourCUDA) Reset(sdf *gleval.SDF3CUDA, res float32, vertexPlacer *DualContourLeastSquares, part, nparts int) error {
	if vertexPlacer == nil {
		return errors.New("nil DualContourer argument to Reset") // dual_contour.go:28-30
	}
	placer := C.int(C.GSDF_DC_LEAST_SQUARES)
	if vertexPlacer.Chiseled {
		placer = C.GSDF_DC_LEAST_SQUARES_CHISELED
	}
	if dcr.h != nil {
		C.gsdf_dc_destroy(dcr.h)
		dcr.h = nil
	}
	bb := sdf.Bounds()
	if rc := C.gsdf_dc_begin_part((*C.gsdf_program)(sdf.Handle()), (*C.float)(unsafe.Pointer(&bb.Min)), (*C.float)(unsafe.Pointer(&bb.Max)),
		C.float(res), placer, C.int(part), C.int(nparts), &dcr.h); rc != 0 {
		return cudaErr() // same messages as makeICube (octreerenderer.go:223-233)
	}
	return nil
}

// RenderAll appends the mesh to dst like DualContourRenderer.RenderAll (dual_contour.go:73-218).
func (dcr *DualContourCUDA) RenderAll(dst []ms3.Triangle, userData any) ([]ms3.Triangle, error) {
	var st [6]C.uint64_t
	C.gsdf_dc_stats(dcr.h, &st[0])
	n := int(st[3])
	out := make([]ms3.Triangle, n)
	if n > 0 {
		if got := C.gsdf_dc_read(dcr.h, (*C.float)(unsafe.Pointer(&out[0])), C.size_t(n)); got < 0 {
			return dst, cudaErr()
		}
	}
	return append(dst, out...), nil
}
