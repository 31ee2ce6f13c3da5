package gpu3d

/*
#include "m3d.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"math"
	"time"

	"github.com/unixpickle/model3d/render3d"
)

// Partition selects the rows and the absolute sample range one call renders (m3d_partition):
// path tracers shard by sample index, RayCaster by row band.  The zero value is the whole
// frame starting at sample 0.  On a multi-device Context the library splits the partition
// further over its GPUs by itself.
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

// sums holds the per-pixel colour sums (and optionally sums of squares) of a render in pinned
// host memory.
type sums struct {
	sum, sq *HostFloats
}

func newSums(n int, variance bool) (*sums, error) {
	s := &sums{}
	var err error
	if s.sum, err = NewHostFloats(n); err != nil {
		return nil, err
	}
	if variance {
		if s.sq, err = NewHostFloats(n); err != nil {
			s.sum.Free()
			return nil, err
		}
	}
	return s, nil
}

func (s *sums) free() {
	s.sum.Free()
	if s.sq != nil {
		s.sq.Free()
	}
}

func (s *sums) sqPtr() *C.float {
	if s.sq == nil {
		return nil
	}
	return s.sq.c()
}

// renderChunks drives one frame as a sequence of library calls so that LogFunc fires between
// them like the reference's per-pixel progress reports (ray_renderer.go:40-55): fixed-spp renders
// are cut by sample index, adaptive renders (per-pixel early stop) by row band.  Each call ADDS
// into the same sums.  one(part, sampleCount) runs a call; it returns the samples it took.
func renderChunks(width, height, numSamples int, adaptive bool, logFunc func(frac, sampleRate float64),
	one func(part *Partition, sampleCount int) (int64, error)) error {
	chunks := 1
	if logFunc != nil {
		// about 2^28 samples per call: 0.2 s of C3 on one B200
		chunks = int(math.Ceil(float64(width) * float64(height) * float64(numSamples) / float64(1<<28)))
		if chunks < 1 {
			chunks = 1
		}
	}
	start := time.Now()
	var taken int64
	for i := 0; i < chunks; i++ {
		part := &Partition{}
		count := numSamples
		if adaptive {
			if chunks > height {
				chunks = height
			}
			part.RowBegin, part.RowEnd = i*height/chunks, (i+1)*height/chunks
			if part.RowBegin == part.RowEnd {
				continue
			}
		} else {
			if chunks > numSamples {
				chunks = numSamples
			}
			b, e := i*numSamples/chunks, (i+1)*numSamples/chunks
			part.SampleBegin, count = int64(b), e-b
			if count == 0 {
				continue
			}
		}
		n, err := one(part, count)
		if err != nil {
			return err
		}
		taken += n
		if logFunc != nil {
			// (fraction of the frame done, samples per pixel-second like ray_renderer.go:50-52)
			logFunc(float64(i+1)/float64(chunks), float64(taken)/time.Since(start).Seconds())
		}
	}
	return nil
}

func fillImage(img *render3d.Image, sum []float32, numSamples int) {
	inv := 1 / float64(numSamples) // colorSum.Scale(1/numSamples), ray_renderer.go:150
	for i := range img.Data {
		img.Data[i] = render3d.Color{X: float64(sum[3*i]) * inv, Y: float64(sum[3*i+1]) * inv,
			Z: float64(sum[3*i+2]) * inv}
	}
}

// fillVariance writes the per-pixel sample variance with Bessel's correction, clamped at zero
// (rayRenderer.estimateVariance, ray_renderer.go:90-110).
func fillVariance(img *render3d.Image, sum, sq []float32, numSamples int) {
	n := float64(numSamples)
	v := func(s, q float32) float64 {
		mean := float64(s) / n
		return math.Max(0, (float64(q)/n-mean*mean)*(n/(n-1)))
	}
	for i := range img.Data {
		img.Data[i] = render3d.Color{X: v(sum[3*i], sq[3*i]), Y: v(sum[3*i+1], sq[3*i+1]), Z: v(sum[3*i+2], sq[3*i+2])}
	}
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
	// Context the object is compiled on when it is not a *Scene (nil: DefaultContext, all GPUs).
	Context *Context
}

func (r *RecursiveRayTracer) params(scene *Scene, numSamples int) (C.m3d_path_params, error) {
	var p C.m3d_path_params
	if numSamples == 0 {
		return p, errors.New("must set NumSamples to non-zero for rayRenderer") // ray_renderer.go:26-28
	}
	if len(r.FocusPoints) != len(r.FocusPointProbs) {
		return p, errors.New("FocusPoints and FocusPointProbs must match in length") // raytrace.go:186-188
	}
	if len(r.FocusPoints) > C.M3D_MAX_FOCUS_POINTS {
		return p, &Error{Code: int(C.M3D_ERR_UNSUPPORTED),
			Msg: fmt.Sprintf("at most %d focus points are supported", int(C.M3D_MAX_FOCUS_POINTS))}
	}
	if r.Convergence != nil {
		return p, &Error{Code: int(C.M3D_ERR_UNSUPPORTED),
			Msg: "Convergence callbacks cannot run on the GPU path (MinSamples / MaxStddev / OversaturatedStddevs can)"}
	}
	p.max_depth = C.int32_t(r.MaxDepth)
	p.num_samples = C.int32_t(numSamples)
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
			return p, &Error{Code: int(C.M3D_ERR_UNSUPPORTED),
				Msg: fmt.Sprintf("focus point type %T is not supported on the GPU path", fp)}
		}
		// MaterialFilter closures cannot cross the C ABI: evaluate once per scene material
		var mask uint64
		for idx, m := range scene.matList {
			if idx < 64 && (filter == nil || filter(m)) {
				mask |= 1 << uint(idx)
			}
		}
		f.material_mask = C.uint64_t(mask)
		f.prob = C.double(r.FocusPointProbs[i])
	}
	return p, nil
}

func (r *RecursiveRayTracer) adaptive() bool { return r.MinSamples != 0 && r.MaxStddev != 0 }

// renderSums adds the sums of numSamples samples per pixel into out.  fixed: no early stop
// (RenderVariance / sample shards).
func (r *RecursiveRayTracer) renderSums(width, height int, scene *Scene, numSamples int, fixed bool,
	antialias float64, out *sums) error {
	p, err := r.params(scene, numSamples)
	if err != nil {
		return err
	}
	p.antialias = C.double(antialias)
	if fixed {
		p.min_samples = 0
	}
	cam := ccamera(r.Camera)
	lights := clights(r.Lights)
	adaptive := !fixed && r.adaptive()
	return renderChunks(width, height, numSamples, adaptive, r.LogFunc, func(part *Partition, count int) (int64, error) {
		// the host-buffer call overwrites its output: render each chunk into a scratch pair and add
		chunk, err := newSums(3*width*height, out.sq != nil)
		if err != nil {
			return 0, err
		}
		defer chunk.free()
		var st C.m3d_stats
		err = call(func() C.int32_t {
			return C.m3d_render_path(scene.h, &cam, &lights[0], C.int32_t(len(r.Lights)), &p, C.int32_t(width),
				C.int32_t(height), part.c(), C.int32_t(count), chunk.sum.c(), chunk.sqPtr(), &st)
		})
		if err != nil {
			return 0, err
		}
		for i, v := range chunk.sum.S {
			out.sum.S[i] += v
		}
		if out.sq != nil {
			for i, v := range chunk.sq.S {
				out.sq.S[i] += v
			}
		}
		return int64(st.samples), nil
	})
}

// RenderSums returns the per-pixel colour SUMS of sampleCount samples of this partition
// (3 float32 per pixel, idx = x + y*W): the quantity that adds up across processes.
func (r *RecursiveRayTracer) RenderSums(width, height int, obj render3d.Object, part *Partition,
	sampleCount int) ([]float32, error) {
	scene, release, err := sceneFor(r.Context, obj)
	if err != nil {
		return nil, err
	}
	defer release()
	p, err := r.params(scene, sampleCount)
	if err != nil {
		return nil, err
	}
	if part != nil && part.SampleBegin != 0 {
		p.min_samples = 0 // a sample shard cannot stop early
	}
	cam := ccamera(r.Camera)
	lights := clights(r.Lights)
	out := make([]float32, 3*width*height)
	err = call(func() C.int32_t {
		return C.m3d_render_path(scene.h, &cam, &lights[0], C.int32_t(len(r.Lights)), &p, C.int32_t(width),
			C.int32_t(height), part.c(), C.int32_t(sampleCount), fptr(out), nil, nil)
	})
	return out, err
}

// Render renders obj like (*render3d.RecursiveRayTracer).Render (raytrace.go:98-100).
func (r *RecursiveRayTracer) Render(img *render3d.Image, obj render3d.Object) error {
	scene, release, err := sceneFor(r.Context, obj)
	if err != nil {
		return err
	}
	defer release()
	out, err := newSums(3*img.Width*img.Height, false)
	if err != nil {
		return err
	}
	defer out.free()
	if err := r.renderSums(img.Width, img.Height, scene, r.NumSamples, false, r.Antialias, out); err != nil {
		return err
	}
	// adaptive renders return mean * NumSamples per pixel, so the same division applies
	fillImage(img, out.sum.S, r.NumSamples)
	return nil
}

// RenderVariance mirrors rayRenderer.RenderVariance (raytrace.go:104-110, ray_renderer.go:59-67):
// img receives the per-pixel variance of numSamples samples.
func (r *RecursiveRayTracer) RenderVariance(img *render3d.Image, obj render3d.Object, numSamples int) error {
	return r.renderVariance(img, obj, numSamples, r.Antialias)
}

func (r *RecursiveRayTracer) renderVariance(img *render3d.Image, obj render3d.Object, numSamples int,
	antialias float64) error {
	if numSamples < 2 {
		return errors.New("need to take at least two samples") // ray_renderer.go:92-94
	}
	scene, release, err := sceneFor(r.Context, obj)
	if err != nil {
		return err
	}
	defer release()
	out, err := newSums(3*img.Width*img.Height, true)
	if err != nil {
		return err
	}
	defer out.free()
	if err := r.renderSums(img.Width, img.Height, scene, numSamples, true, antialias, out); err != nil {
		return err
	}
	fillVariance(img, out.sum.S, out.sq.S, numSamples)
	return nil
}

// RayVariance mirrors rayRenderer.RayVariance (raytrace.go:112-119, ray_renderer.go:69-88): the
// mean per-channel variance of single samples over a width x height frame without antialiasing.
func (r *RecursiveRayTracer) RayVariance(obj render3d.Object, width, height, samples int) (float64, error) {
	img := render3d.NewImage(width, height)
	if err := r.renderVariance(img, obj, samples, 0); err != nil {
		return 0, err
	}
	var total float64
	for _, c := range img.Data {
		total += c.X + c.Y + c.Z
	}
	return total / float64(3*width*height), nil
}

// BidirPathTracer mirrors render3d.BidirPathTracer (render3d/bidir.go:14-63): the exported
// fields have the same names, types and meaning.  Light takes *render3d.SphereAreaLight,
// gpu3d.MeshAreaLight and gpu3d.JoinedAreaLight (the reference's mesh / joined lights hide their
// geometry); every light must also be an object of the rendered scene, as in the reference.
type BidirPathTracer struct {
	Camera *render3d.Camera
	Light  render3d.AreaLight

	MaxDepth      int
	MaxLightDepth int
	MinDepth      int

	NumSamples           int
	MinSamples           int
	MaxStddev            float64
	OversaturatedStddevs float64
	Convergence          func(mean, stddev render3d.Color) bool

	RouletteDelta  float64
	PowerHeuristic float64

	Cutoff    float64
	Antialias float64
	Epsilon   float64
	LogFunc   func(frac float64, sampleRate float64)

	Seed    uint64
	Context *Context
}

// areaLights resolves b.Light into scene object indices and emissions.
func areaLights(scene *Scene, light render3d.AreaLight) ([]C.m3d_area_light, error) {
	var out []C.m3d_area_light
	var walk func(l render3d.AreaLight) error
	walk = func(l render3d.AreaLight) error {
		var emission render3d.Color
		switch l := l.(type) {
		case *JoinedAreaLight:
			for _, part := range l.Lights {
				if err := walk(part); err != nil {
					return err
				}
			}
			return nil
		case *MeshAreaLight:
			emission = l.Emission
		case *render3d.SphereAreaLight:
			co, ok := l.Object.(*render3d.ColliderObject)
			if !ok {
				return &Error{Code: int(C.M3D_ERR_UNSUPPORTED), Msg: "SphereAreaLight without a ColliderObject"}
			}
			emission = co.Material.Emission()
		default:
			return &Error{Code: int(C.M3D_ERR_UNSUPPORTED), Msg: fmt.Sprintf(
				"area light type %T is not supported on the GPU path (use gpu3d.NewMeshAreaLight / JoinAreaLights)", l)}
		}
		idx := scene.leafIndex(l)
		if idx < 0 {
			return &Error{Code: int(C.M3D_ERR_INVALID_ARG), Msg: "the area light is not an object of the rendered scene"}
		}
		out = append(out, C.m3d_area_light{object: C.int32_t(idx), emission: cvec(emission)})
		return nil
	}
	if light == nil {
		return nil, errors.New("gpu3d: BidirPathTracer needs an area light")
	}
	if err := walk(light); err != nil {
		return nil, err
	}
	if len(out) == 0 {
		return nil, errors.New("gpu3d: BidirPathTracer needs at least one area light")
	}
	return out, nil
}

func (b *BidirPathTracer) params(numSamples int) (C.m3d_bidir_params, error) {
	var p C.m3d_bidir_params
	if numSamples == 0 {
		return p, errors.New("must set NumSamples to non-zero for rayRenderer")
	}
	if b.Convergence != nil {
		return p, &Error{Code: int(C.M3D_ERR_UNSUPPORTED), Msg: "Convergence callbacks cannot run on the GPU path"}
	}
	p.max_depth = C.int32_t(b.MaxDepth)
	p.max_light_depth = C.int32_t(b.MaxLightDepth)
	p.min_depth = C.int32_t(b.MinDepth)
	p.num_samples = C.int32_t(numSamples)
	p.roulette_delta = C.double(b.RouletteDelta)
	p.power_heuristic = C.double(b.PowerHeuristic)
	p.cutoff = C.double(b.Cutoff)
	p.antialias = C.double(b.Antialias)
	p.epsilon = C.double(b.Epsilon)
	p.seed = C.uint64_t(b.Seed)
	// adaptive stop (bidir.go:45-52), as in m3d_path_params
	p.min_samples = C.int32_t(b.MinSamples)
	p.max_stddev = C.double(b.MaxStddev)
	p.oversaturated_stddevs = C.double(b.OversaturatedStddevs)
	return p, nil
}

func (b *BidirPathTracer) renderSums(width, height int, scene *Scene, numSamples int, fixed bool,
	antialias float64, out *sums) error {
	p, err := b.params(numSamples)
	if err != nil {
		return err
	}
	p.antialias = C.double(antialias)
	if fixed {
		p.min_samples = 0
	}
	lights, err := areaLights(scene, b.Light)
	if err != nil {
		return err
	}
	cam := ccamera(b.Camera)
	adaptive := !fixed && b.MinSamples != 0 && b.MaxStddev != 0
	return renderChunks(width, height, numSamples, adaptive, b.LogFunc, func(part *Partition, count int) (int64, error) {
		chunk, err := newSums(3*width*height, out.sq != nil)
		if err != nil {
			return 0, err
		}
		defer chunk.free()
		var st C.m3d_stats
		err = call(func() C.int32_t {
			return C.m3d_render_bidir(scene.h, &cam, &lights[0], C.int32_t(len(lights)), &p, C.int32_t(width),
				C.int32_t(height), part.c(), C.int32_t(count), chunk.sum.c(), chunk.sqPtr(), &st)
		})
		if err != nil {
			return 0, err
		}
		for i, v := range chunk.sum.S {
			out.sum.S[i] += v
		}
		if out.sq != nil {
			for i, v := range chunk.sq.S {
				out.sq.S[i] += v
			}
		}
		return int64(st.samples), nil
	})
}

// Render renders obj like (*render3d.BidirPathTracer).Render (bidir.go:66-68).
func (b *BidirPathTracer) Render(img *render3d.Image, obj render3d.Object) error {
	scene, release, err := sceneFor(b.Context, obj)
	if err != nil {
		return err
	}
	defer release()
	out, err := newSums(3*img.Width*img.Height, false)
	if err != nil {
		return err
	}
	defer out.free()
	if err := b.renderSums(img.Width, img.Height, scene, b.NumSamples, false, b.Antialias, out); err != nil {
		return err
	}
	fillImage(img, out.sum.S, b.NumSamples)
	return nil
}

// RenderVariance mirrors (*render3d.BidirPathTracer).RenderVariance (bidir.go:70-76).
func (b *BidirPathTracer) RenderVariance(img *render3d.Image, obj render3d.Object, numSamples int) error {
	return b.renderVariance(img, obj, numSamples, b.Antialias)
}

func (b *BidirPathTracer) renderVariance(img *render3d.Image, obj render3d.Object, numSamples int,
	antialias float64) error {
	if numSamples < 2 {
		return errors.New("need to take at least two samples")
	}
	scene, release, err := sceneFor(b.Context, obj)
	if err != nil {
		return err
	}
	defer release()
	out, err := newSums(3*img.Width*img.Height, true)
	if err != nil {
		return err
	}
	defer out.free()
	if err := b.renderSums(img.Width, img.Height, scene, numSamples, true, antialias, out); err != nil {
		return err
	}
	fillVariance(img, out.sum.S, out.sq.S, numSamples)
	return nil
}

// RayVariance mirrors (*render3d.BidirPathTracer).RayVariance (bidir.go:78-84).
func (b *BidirPathTracer) RayVariance(obj render3d.Object, width, height, samples int) (float64, error) {
	img := render3d.NewImage(width, height)
	if err := b.renderVariance(img, obj, samples, 0); err != nil {
		return 0, err
	}
	var total float64
	for _, c := range img.Data {
		total += c.X + c.Y + c.Z
	}
	return total / float64(3*width*height), nil
}
