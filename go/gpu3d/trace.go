package gpu3d

/*
#include "m3d.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"unsafe"

	"github.com/unixpickle/model3d/render3d"
)

// Partition selects the rows and the absolute sample range one GPU renders (m3d_partition):
// path tracers shard by sample index, RayCaster by row band.  The zero value is the whole
// frame starting at sample 0.
type Partition struct {
	RowBegin, RowEnd int
	SampleBegin      int64
}

func (p *Partition) c() *C.m3d_partition {
	if p == nil {
		return nil
	}
	return &C.m3d_partition{row_begin: C.int32_t(p.RowBegin), row_end: C.int32_t(p.RowEnd),
		sample_begin: C.int64_t(p.SampleBegin)}
}

// RecursiveRayTracer mirrors render3d.RecursiveRayTracer (render3d/raytrace.go:14-95): the
// exported fields have the same names, types and meaning.  Render runs the wavefront path
// tracer of libm3dgpu (m3d_render_path).
type RecursiveRayTracer struct {
	Camera *render3d.Camera
	Lights []*render3d.PointLight

	FocusPoints     []render3d.FocusPoint
	FocusPointProbs []float64

	MaxDepth   int
	NumSamples int

	MinSamples           int
	MaxStddev            float64
	OversaturatedStddevs float64
	Convergence          func(mean, stddev render3d.Color) bool

	Cutoff    float64
	Antialias float64
	Epsilon   float64
	LogFunc   func(frac float64, sampleRate float64)

	// Seed keys the Philox streams (the reference seeds math/rand from the global source).
	Seed uint64
}

func (r *RecursiveRayTracer) params(scene *Scene) (C.m3d_path_params, error) {
	var p C.m3d_path_params
	if r.NumSamples == 0 {
		return p, errors.New("must set NumSamples to non-zero for rayRenderer") // ray_renderer.go:26-28
	}
	if len(r.FocusPoints) != len(r.FocusPointProbs) {
		return p, errors.New("FocusPoints and FocusPointProbs must match in length") // raytrace.go:186-188
	}
	if len(r.FocusPoints) > C.M3D_MAX_FOCUS_POINTS {
		return p, fmt.Errorf("gpu3d: at most %d focus points are supported", int(C.M3D_MAX_FOCUS_POINTS))
	}
	if r.Convergence != nil {
		return p, errors.New("gpu3d: Convergence callbacks cannot run on the GPU path")
	}
	p.max_depth = C.int32_t(r.MaxDepth)
	p.num_samples = C.int32_t(r.NumSamples)
	p.min_samples = C.int32_t(r.MinSamples)
	p.max_stddev = C.double(r.MaxStddev)
	p.oversaturated_stddevs = C.double(r.OversaturatedStddevs)
	p.cutoff = C.double(r.Cutoff)
	p.antialias = C.double(r.Antialias)
	p.epsilon = C.double(r.Epsilon)
	p.seed = C.uint64_t(r.Seed)
	p.num_focus_points = C.int32_t(len(r.FocusPoints))
	for i, fp := range r.FocusPoints {
		f := &p.focus[i]
		var filter func(render3d.Material) bool
		switch fp := fp.(type) {
		case *render3d.PhongFocusPoint:
			f.kind = C.M3D_FOCUS_PHONG
			f.target = cvec(fp.Target)
			f.alpha = C.double(fp.Alpha)
			filter = fp.MaterialFilter
		case *render3d.SphereFocusPoint:
			f.kind = C.M3D_FOCUS_SPHERE
			f.target = cvec(fp.Center)
			f.radius = C.double(fp.Radius)
			filter = fp.MaterialFilter
		default:
			return p, fmt.Errorf("gpu3d: focus point type %T is not supported on the GPU path", fp)
		}
		// MaterialFilter closures cannot cross the C ABI: evaluate once per scene material
		var mask uint64
		for m, idx := range scene.materials {
			if idx < 64 && (filter == nil || filter(m)) {
				mask |= 1 << uint(idx)
			}
		}
		f.material_mask = C.uint64_t(mask)
		f.prob = C.double(r.FocusPointProbs[i])
	}
	return p, nil
}

func clights(ls []*render3d.PointLight) []C.m3d_point_light {
	out := make([]C.m3d_point_light, len(ls)+1)
	for i, l := range ls {
		out[i].origin = cvec(l.Origin)
		out[i].color = cvec(l.Color)
		if l.QuadDropoff {
			out[i].quad_dropoff = 1
		}
	}
	return out
}

// RenderSums returns the per-pixel colour SUMS of sampleCount samples of this partition
// (3 float32 per pixel, idx = x + y*W): the quantity that adds up across GPUs.
func (r *RecursiveRayTracer) RenderSums(width, height int, scene *Scene, part *Partition,
	sampleCount int) ([]float32, error) {
	p, err := r.params(scene)
	if err != nil {
		return nil, err
	}
	cam := ccamera(r.Camera)
	lights := clights(r.Lights)
	sums := make([]float32, 3*width*height)
	err = status(C.m3d_render_path(scene.h, &cam, &lights[0], C.int32_t(len(r.Lights)), &p,
		C.int32_t(width), C.int32_t(height), part.c(), C.int32_t(sampleCount),
		(*C.float)(unsafe.Pointer(&sums[0])), nil, nil))
	return sums, err
}

// Render renders the scene like (*render3d.RecursiveRayTracer).Render (raytrace.go:98-100).
func (r *RecursiveRayTracer) Render(img *render3d.Image, scene *Scene) error {
	sums, err := r.RenderSums(img.Width, img.Height, scene, nil, r.NumSamples)
	if err != nil {
		return err
	}
	inv := 1 / float64(r.NumSamples) // colorSum.Scale(1/numSamples), ray_renderer.go:150
	for i := range img.Data {
		img.Data[i] = render3d.Color{X: float64(sums[3*i]) * inv, Y: float64(sums[3*i+1]) * inv,
			Z: float64(sums[3*i+2]) * inv}
	}
	if r.LogFunc != nil {
		r.LogFunc(1, float64(r.NumSamples))
	}
	return nil
}

// AreaLightRef names a scene object that BidirPathTracer samples as an emitter
// (render3d.AreaLight, light.go:104-314): the object must be a sphere or a MeshObject that
// is part of the scene, with the emission colour of its material.
type AreaLightRef struct {
	Object   int // index of the leaf object in scene order
	Emission render3d.Color
}

// BidirPathTracer mirrors render3d.BidirPathTracer (render3d/bidir.go:14-63).
type BidirPathTracer struct {
	Camera *render3d.Camera
	Light  []AreaLightRef

	MaxDepth      int
	MaxLightDepth int
	MinDepth      int

	RouletteDelta  float64
	PowerHeuristic float64

	NumSamples           int
	MinSamples           int
	MaxStddev            float64
	OversaturatedStddevs float64

	Cutoff    float64
	Antialias float64
	Epsilon   float64
	LogFunc   func(frac float64, sampleRate float64)
	Seed      uint64
}

// RenderSums: see RecursiveRayTracer.RenderSums.
func (b *BidirPathTracer) RenderSums(width, height int, scene *Scene, part *Partition,
	sampleCount int) ([]float32, error) {
	if b.NumSamples == 0 {
		return nil, errors.New("must set NumSamples to non-zero for rayRenderer")
	}
	if len(b.Light) == 0 {
		return nil, errors.New("gpu3d: BidirPathTracer needs at least one area light")
	}
	var p C.m3d_bidir_params
	p.max_depth = C.int32_t(b.MaxDepth)
	p.max_light_depth = C.int32_t(b.MaxLightDepth)
	p.min_depth = C.int32_t(b.MinDepth)
	p.num_samples = C.int32_t(b.NumSamples)
	p.roulette_delta = C.double(b.RouletteDelta)
	p.power_heuristic = C.double(b.PowerHeuristic)
	p.cutoff = C.double(b.Cutoff)
	p.antialias = C.double(b.Antialias)
	p.epsilon = C.double(b.Epsilon)
	p.seed = C.uint64_t(b.Seed)
	lights := make([]C.m3d_area_light, len(b.Light))
	for i, l := range b.Light {
		lights[i].object = C.int32_t(l.Object)
		lights[i].emission = cvec(l.Emission)
	}
	cam := ccamera(b.Camera)
	sums := make([]float32, 3*width*height)
	err := status(C.m3d_render_bidir(scene.h, &cam, &lights[0], C.int32_t(len(lights)), &p,
		C.int32_t(width), C.int32_t(height), part.c(), C.int32_t(sampleCount),
		(*C.float)(unsafe.Pointer(&sums[0])), nil, nil))
	return sums, err
}

// Render renders the scene like (*render3d.BidirPathTracer).Render (bidir.go:66-68).
func (b *BidirPathTracer) Render(img *render3d.Image, scene *Scene) error {
	sums, err := b.RenderSums(img.Width, img.Height, scene, nil, b.NumSamples)
	if err != nil {
		return err
	}
	inv := 1 / float64(b.NumSamples)
	for i := range img.Data {
		img.Data[i] = render3d.Color{X: float64(sums[3*i]) * inv, Y: float64(sums[3*i+1]) * inv,
			Z: float64(sums[3*i+2]) * inv}
	}
	if b.LogFunc != nil {
		b.LogFunc(1, float64(b.NumSamples))
	}
	return nil
}
