// Package gpu3d is the cgo binding of libm3dgpu (include/m3d.h): GPU-backed drop-ins for
// model3d.Collider and the render3d renderers.  It keeps the reference's signatures; the
// only additions are batch entry points and error returns.  There is no CPU fallback:
// unsupported object / material types return an error.
//
// NOTE: the build container of this repository has no Go toolchain, so this package is
// shipped as source and has not been compiled there.  Build with
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
	"errors"
	"runtime"
	"unsafe"

	"github.com/unixpickle/model3d/model3d"
	"github.com/unixpickle/model3d/render3d"
)

func status(rc C.int32_t) error {
	if rc == C.M3D_OK {
		return nil
	}
	return errors.New("m3dgpu: " + C.GoString(C.m3d_last_error()))
}

// Context owns one CUDA device.
type Context struct{ h *C.m3d_ctx }

// NewContext opens a device (-1 = current device).
func NewContext(device int) (*Context, error) {
	c := &Context{}
	if err := status(C.m3d_ctx_create(C.int32_t(device), &c.h)); err != nil {
		return nil, err
	}
	runtime.SetFinalizer(c, (*Context).Close)
	return c, nil
}

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
	return status(C.m3d_ctx_trim(c.h))
}

// MeshCollider implements model3d.Collider on the GPU for a triangle mesh.
// Triangle ids are indices into Triangles (the reference identifies triangles by
// pointer, model3d/collisions.go:39-46).
type MeshCollider struct {
	h         *C.m3d_mesh
	Triangles []*model3d.Triangle
	min, max  model3d.Coord3D
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

func meshToCollider(ctx *Context, m *model3d.Mesh, buildFlags uint32, interpNormals bool) (*MeshCollider, error) {
	tris := m.TriangleSlice()
	flat := make([]float32, 0, len(tris)*9)
	for _, t := range tris {
		for _, p := range t {
			flat = append(flat, float32(p.X), float32(p.Y), float32(p.Z))
		}
	}
	var normals []float32
	if interpNormals {
		vn := m.VertexNormals()
		normals = make([]float32, 0, len(tris)*9)
		for _, t := range tris {
			for _, p := range t {
				n := vn.Value(p)
				normals = append(normals, float32(n.X), float32(n.Y), float32(n.Z))
			}
		}
	}
	res := &MeshCollider{Triangles: tris}
	var ptr, nptr *C.float
	if len(flat) > 0 {
		ptr = (*C.float)(unsafe.Pointer(&flat[0]))
	}
	if len(normals) > 0 {
		nptr = (*C.float)(unsafe.Pointer(&normals[0]))
	}
	if err := status(C.m3d_mesh_create(ctx.h, ptr, C.int64_t(len(tris)), nptr, C.uint32_t(buildFlags), &res.h)); err != nil {
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

// FirstRayCollisions is the batched form of Collider.FirstRayCollision
// (collisions.go:275-290): out[i], hit[i] describe rays[i].
func (m *MeshCollider) FirstRayCollisions(rays []model3d.Ray, out []model3d.RayCollision, hit []bool) error {
	n := len(rays)
	if n == 0 {
		return nil
	}
	org := make([]float32, 3*n)
	dir := make([]float32, 3*n)
	for i, r := range rays {
		org[3*i], org[3*i+1], org[3*i+2] = float32(r.Origin.X), float32(r.Origin.Y), float32(r.Origin.Z)
		dir[3*i], dir[3*i+1], dir[3*i+2] = float32(r.Direction.X), float32(r.Direction.Y), float32(r.Direction.Z)
	}
	t := make([]float32, n)
	prim := make([]int32, n)
	normal := make([]float32, 3*n)
	bary := make([]float32, 3*n)
	err := status(C.m3d_mesh_first_ray_collisions(m.h,
		(*C.float)(unsafe.Pointer(&org[0])), (*C.float)(unsafe.Pointer(&dir[0])), C.int64_t(n),
		(*C.float)(unsafe.Pointer(&t[0])), (*C.int32_t)(unsafe.Pointer(&prim[0])),
		(*C.float)(unsafe.Pointer(&normal[0])), (*C.float)(unsafe.Pointer(&bary[0])), 0, nil))
	if err != nil {
		return err
	}
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
