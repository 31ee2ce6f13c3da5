// Package gpu3d is the cgo binding of libm3dgpu (include/m3d.h): GPU-backed drop-ins for
// model3d.Collider, render3d.Object and the render3d renderers.  It keeps the reference's
// signatures; the only additions are batch entry points and error returns.  There is no CPU
// fallback: unsupported object / material types return an error.
//
// NOTE: the build container of this repository has no Go toolchain, so this package is shipped
// as source and has not been compiled there.  The same call sequence is exercised from C by
// tests/c_abi/cgo_sequence.c (run by tests/test_c_abi.py on the GPU box).  Build with
//
//	CGO_CFLAGS="-I${REPO}/include" CGO_LDFLAGS="-L${REPO}/model3d_b200 -lm3dgpu" go build ./go/gpu3d
package gpu3d

/*
#cgo LDFLAGS: -lm3dgpu
#include <stdlib.h>
#include "m3d.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"sync"
	"unsafe"

	"github.com/unixpickle/model3d/model3d"
	"github.com/unixpickle/model3d/render3d"
)

// Error is a failed libm3dgpu call: the m3d_status code and the library's message.
type Error struct {
	Code int
	Msg  string
}

func (e *Error) Error() string { return fmt.Sprintf("m3dgpu (status %d): %s", e.Code, e.Msg) }

// Unsupported reports whether the error means "this object / material / feature is outside the
// GPU path" (M3D_ERR_UNSUPPORTED); callers that want the reference's CPU renderer for such scenes
// branch on it themselves -- this package never falls back silently.
func (e *Error) Unsupported() bool { return e.Code == int(C.M3D_ERR_UNSUPPORTED) }

// call runs one library call and, if it failed, reads its message.  m3d_last_error() is
// thread-local, and a goroutine may move to another OS thread between two cgo calls, so the pair
// runs with the goroutine locked to its thread.
func call(f func() C.int32_t) error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	rc := f()
	if rc == C.M3D_OK {
		return nil
	}
	return &Error{Code: int(rc), Msg: C.GoString(C.m3d_last_error())}
}

// Context owns one CUDA device, or several devices of one node (NewMultiContext): meshes and
// scenes built on a multi-device context are replicated, ray batches are sliced and renders are
// sharded over the devices inside the library (one host thread per device; the per-pixel sums
// meet in the first device's accumulator through the flush kernels' NVLink reductions).
// Calls on one Context are serialised by the library; handles may be shared between goroutines.
type Context struct{ h *C.m3d_ctx }

// NewContext opens a device (-1 = current device).
func NewContext(device int) (*Context, error) {
	c := &Context{}
	if err := call(func() C.int32_t { return C.m3d_ctx_create(C.int32_t(device), &c.h) }); err != nil {
		return nil, err
	}
	runtime.SetFinalizer(c, (*Context).Close)
	return c, nil
}

// NewMultiContext opens the given devices of this node as one context (nil: every visible device).
// It replaces the reference's goroutine scheduler (render3d/concurrency.go:17-43) across GPUs.
func NewMultiContext(devices []int) (*Context, error) {
	c := &Context{}
	var ptr *C.int32_t
	ids := make([]C.int32_t, len(devices))
	for i, d := range devices {
		ids[i] = C.int32_t(d)
	}
	if len(ids) > 0 {
		ptr = &ids[0]
	}
	if err := call(func() C.int32_t { return C.m3d_ctx_create_multi(ptr, C.int32_t(len(ids)), &c.h) }); err != nil {
		return nil, err
	}
	runtime.SetFinalizer(c, (*Context).Close)
	return c, nil
}

var (
	defaultOnce sync.Once
	defaultCtx  *Context
	defaultErr  error
)

// DefaultContext is the process-wide context over all visible GPUs that Render uses for objects
// that were not built on an explicit Context.
func DefaultContext() (*Context, error) {
	defaultOnce.Do(func() { defaultCtx, defaultErr = NewMultiContext(nil) })
	return defaultCtx, defaultErr
}

// NumDevices is the number of GPUs behind the context.
func (c *Context) NumDevices() int { return int(C.m3d_ctx_num_devices(c.h)) }

// Close releases the device context.
func (c *Context) Close() {
	if c.h != nil {
		C.m3d_ctx_destroy(c.h)
		c.h = nil
	}
}

// Trim releases the scratch buffers the renderers keep between calls (path state of up to
// 48-64 GB for the largest batches); later calls allocate again on demand.
func (c *Context) Trim() error {
	return call(func() C.int32_t { return C.m3d_ctx_trim(c.h) })
}

// ---- pinned host memory ------------------------------------------------------------------------

// HostFloats is a []float32 in page-locked memory from m3d_host_alloc: the host-buffer calls
// reach the PCIe rate only from such buffers (pageable Go slices are staged by the driver at about
// a sixth of the rate).  The memory is outside the Go heap; call Free when done.
type HostFloats struct {
	S   []float32
	ptr unsafe.Pointer
}

// NewHostFloats allocates n float32 of pinned host memory.
func NewHostFloats(n int) (*HostFloats, error) {
	h := &HostFloats{}
	if n == 0 {
		return h, nil
	}
	if err := call(func() C.int32_t { return C.m3d_host_alloc(C.int64_t(4*n), &h.ptr) }); err != nil {
		return nil, err
	}
	h.S = unsafe.Slice((*float32)(h.ptr), n)
	runtime.SetFinalizer(h, (*HostFloats).Free)
	return h, nil
}

// Free releases the buffer; S must not be used afterwards.
func (h *HostFloats) Free() {
	if h.ptr != nil {
		C.m3d_host_free(h.ptr)
		h.ptr, h.S = nil, nil
	}
}

func (h *HostFloats) c() *C.float { return (*C.float)(h.ptr) }

// HostInts is the int32 counterpart of HostFloats.
type HostInts struct {
	S   []int32
	ptr unsafe.Pointer
}

// NewHostInts allocates n int32 of pinned host memory.
func NewHostInts(n int) (*HostInts, error) {
	h := &HostInts{}
	if n == 0 {
		return h, nil
	}
	if err := call(func() C.int32_t { return C.m3d_host_alloc(C.int64_t(4*n), &h.ptr) }); err != nil {
		return nil, err
	}
	h.S = unsafe.Slice((*int32)(h.ptr), n)
	runtime.SetFinalizer(h, (*HostInts).Free)
	return h, nil
}

// Free releases the buffer; S must not be used afterwards.
func (h *HostInts) Free() {
	if h.ptr != nil {
		C.m3d_host_free(h.ptr)
		h.ptr, h.S = nil, nil
	}
}

func (h *HostInts) c() *C.int32_t { return (*C.int32_t)(h.ptr) }

// ---- mesh collider -----------------------------------------------------------------------------

// MeshCollider implements model3d.Collider on the GPU for a triangle mesh.
// Triangle ids are indices into Triangles (the reference identifies triangles by
// pointer, model3d/collisions.go:39-46).
type MeshCollider struct {
	h         *C.m3d_mesh
	ctx       *Context
	Triangles []*model3d.Triangle
	// VertexNormals is non-nil for colliders made by MeshToInterpNormalCollider: one normal per
	// triangle corner, in Triangles order.
	VertexNormals [][3]model3d.Coord3D
	min, max      model3d.Coord3D
}

// BVH builders (m3d_mesh_create build_flags).
const (
	BuildHostSAH        uint32 = C.M3D_MESH_BUILD_HOST_SAH        // best tree, ~0.75 s per million triangles
	BuildDeviceLBVH     uint32 = C.M3D_MESH_BUILD_DEVICE_LBVH     // device binary tree, host collapse
	BuildDeviceCollapse uint32 = C.M3D_MESH_BUILD_DEVICE_COLLAPSE // whole build on the device, ~46 ms per million
)

// MeshToCollider replaces model3d.MeshToCollider (collisions.go:138-142).
func MeshToCollider(ctx *Context, m *model3d.Mesh) (*MeshCollider, error) {
	return MeshToColliderBuild(ctx, m, BuildHostSAH)
}

// MeshToInterpNormalCollider replaces model3d.MeshToInterpNormalCollider (collisions.go:147-162):
// collision normals are interpolated from the mesh's vertex normals (Mesh.VertexNormals,
// mesh_ops.go:146-169, computed by the reference's own code here) instead of flat per triangle.
func MeshToInterpNormalCollider(ctx *Context, m *model3d.Mesh) (*MeshCollider, error) {
	return meshToCollider(ctx, m, BuildHostSAH, true)
}

// MeshToColliderBuild is MeshToCollider with an explicit BVH builder.
func MeshToColliderBuild(ctx *Context, m *model3d.Mesh, buildFlags uint32) (*MeshCollider, error) {
	return meshToCollider(ctx, m, buildFlags, false)
}

func flatTriangles(tris []*model3d.Triangle) []float32 {
	flat := make([]float32, 0, len(tris)*9)
	for _, t := range tris {
		for _, p := range t {
			flat = append(flat, float32(p.X), float32(p.Y), float32(p.Z))
		}
	}
	return flat
}

func flatNormals(ns [][3]model3d.Coord3D) []float32 {
	flat := make([]float32, 0, len(ns)*9)
	for _, n3 := range ns {
		for _, n := range n3 {
			flat = append(flat, float32(n.X), float32(n.Y), float32(n.Z))
		}
	}
	return flat
}

func meshToCollider(ctx *Context, m *model3d.Mesh, buildFlags uint32, interpNormals bool) (*MeshCollider, error) {
	tris := m.TriangleSlice()
	res := &MeshCollider{Triangles: tris, ctx: ctx}
	if interpNormals {
		vn := m.VertexNormals()
		res.VertexNormals = make([][3]model3d.Coord3D, len(tris))
		for i, t := range tris {
			for j, p := range t {
				res.VertexNormals[i][j] = vn.Value(p)
			}
		}
	}
	flat := flatTriangles(tris)
	normals := flatNormals(res.VertexNormals)
	err := call(func() C.int32_t {
		return C.m3d_mesh_create(ctx.h, fptr(flat), C.int64_t(len(tris)), fptr(normals), C.uint32_t(buildFlags), &res.h)
	})
	runtime.KeepAlive(flat)
	runtime.KeepAlive(normals)
	if err != nil {
		return nil, err
	}
	var mn, mx [3]C.double
	C.m3d_mesh_bounds(res.h, &mn[0], &mx[0])
	res.min = model3d.XYZ(float64(mn[0]), float64(mn[1]), float64(mn[2]))
	res.max = model3d.XYZ(float64(mx[0]), float64(mx[1]), float64(mx[2]))
	runtime.SetFinalizer(res, (*MeshCollider).Close)
	return res, nil
}

// Close frees the device BVH.
func (m *MeshCollider) Close() {
	if m.h != nil {
		C.m3d_mesh_destroy(m.h)
		m.h = nil
	}
}

func (m *MeshCollider) Min() model3d.Coord3D { return m.min }
func (m *MeshCollider) Max() model3d.Coord3D { return m.max }

// stagingPool hands out pinned staging buffers for the per-ray arrays of the batch calls, so that
// a caller who passes ordinary Go slices of model3d.Ray still gets page-locked transfers.
type staging struct {
	org, dir, t, normal, bary *HostFloats
	prim                      *HostInts
	n                         int
}

var stagingPool sync.Pool

func getStaging(n int) (*staging, error) {
	if s, ok := stagingPool.Get().(*staging); ok && s.n >= n {
		return s, nil
	}
	s := &staging{n: n}
	var err error
	alloc := func(k int) *HostFloats {
		if err != nil {
			return nil
		}
		var h *HostFloats
		h, err = NewHostFloats(k)
		return h
	}
	s.org, s.dir, s.t, s.normal, s.bary = alloc(3*n), alloc(3*n), alloc(n), alloc(3*n), alloc(3*n)
	if err == nil {
		s.prim, err = NewHostInts(n)
	}
	if err != nil {
		return nil, err
	}
	return s, nil
}

// FirstRayCollisions is the batched form of Collider.FirstRayCollision
// (collisions.go:275-290): out[i], hit[i] describe rays[i].
func (m *MeshCollider) FirstRayCollisions(rays []model3d.Ray, out []model3d.RayCollision, hit []bool) error {
	n := len(rays)
	if n == 0 {
		return nil
	}
	if len(out) < n || len(hit) < n {
		return &Error{Code: int(C.M3D_ERR_INVALID_ARG), Msg: "out / hit are shorter than rays"}
	}
	s, err := getStaging(n)
	if err != nil {
		return err
	}
	defer stagingPool.Put(s)
	org, dir := s.org.S, s.dir.S
	for i, r := range rays {
		org[3*i], org[3*i+1], org[3*i+2] = float32(r.Origin.X), float32(r.Origin.Y), float32(r.Origin.Z)
		dir[3*i], dir[3*i+1], dir[3*i+2] = float32(r.Direction.X), float32(r.Direction.Y), float32(r.Direction.Z)
	}
	err = call(func() C.int32_t {
		return C.m3d_mesh_first_ray_collisions(m.h, s.org.c(), s.dir.c(), C.int64_t(n), s.t.c(), s.prim.c(),
			s.normal.c(), s.bary.c(), 0, nil)
	})
	if err != nil {
		return err
	}
	t, prim, normal, bary := s.t.S, s.prim.S, s.normal.S, s.bary.S
	for i := 0; i < n; i++ {
		hit[i] = prim[i] >= 0
		if !hit[i] {
			out[i] = model3d.RayCollision{}
			continue
		}
		out[i] = model3d.RayCollision{
			Scale:  float64(t[i]),
			Normal: model3d.XYZ(float64(normal[3*i]), float64(normal[3*i+1]), float64(normal[3*i+2])),
			Extra: &model3d.TriangleCollision{
				Triangle:    m.Triangles[prim[i]],
				Barycentric: [3]float64{float64(bary[3*i]), float64(bary[3*i+1]), float64(bary[3*i+2])},
			},
		}
	}
	return nil
}

// FirstHitIDs is the lean batch query for callers that only need the hit distance and the
// triangle id: rays share one origin when origins has length 1 (camera batches: only the
// directions cross the link) and only 8 bytes per ray come back.  dirs and the results live in
// pinned buffers the caller keeps across calls (NewHostFloats / NewHostInts).
func (m *MeshCollider) FirstHitIDs(origins, dirs *HostFloats, t *HostFloats, prim *HostInts) error {
	n := len(dirs.S) / 3
	if n == 0 {
		return nil
	}
	var flags C.uint32_t
	if len(origins.S) == 3 && n > 1 {
		flags |= C.M3D_TRACE_SHARED_ORIGIN
	} else if len(origins.S) != 3*n {
		return &Error{Code: int(C.M3D_ERR_INVALID_ARG), Msg: "origins must hold one point or one per ray"}
	}
	if len(t.S) < n || len(prim.S) < n {
		return &Error{Code: int(C.M3D_ERR_INVALID_ARG), Msg: "t / prim are shorter than the ray batch"}
	}
	return call(func() C.int32_t {
		return C.m3d_mesh_first_ray_collisions(m.h, origins.c(), dirs.c(), C.int64_t(n), t.c(), prim.c(), nil, nil,
			flags, nil)
	})
}

// FirstRayCollision implements model3d.Collider with a batch of one (correct, slow).
func (m *MeshCollider) FirstRayCollision(r *model3d.Ray) (model3d.RayCollision, bool) {
	out := make([]model3d.RayCollision, 1)
	hit := make([]bool, 1)
	if err := m.FirstRayCollisions([]model3d.Ray{*r}, out, hit); err != nil {
		panic(err)
	}
	return out[0], hit[0]
}

// RayCollisions, SphereCollision, Contains and the SDF queries live in queries.go.

func cvec(v model3d.Coord3D) [3]C.double {
	return [3]C.double{C.double(v.X), C.double(v.Y), C.double(v.Z)}
}

func ccamera(c *render3d.Camera) C.m3d_camera {
	return C.m3d_camera{origin: cvec(c.Origin), screen_x: cvec(c.ScreenX), screen_y: cvec(c.ScreenY),
		field_of_view: C.double(c.FieldOfView)}
}
