package gpu3d

/*
#include <stdlib.h>
#include "m3d.h"
*/
import "C"

import (
	"errors"
	"unsafe"

	"github.com/unixpickle/model3d/model3d"
)

func flatCoords(cs []model3d.Coord3D) []float32 {
	flat := make([]float32, 0, len(cs)*3)
	for _, c := range cs {
		flat = append(flat, float32(c.X), float32(c.Y), float32(c.Z))
	}
	return flat
}

func fptr(f []float32) *C.float {
	if len(f) == 0 {
		return nil
	}
	return (*C.float)(unsafe.Pointer(&f[0]))
}

// RayCollisionCounts is the batched Collider.RayCollisions(r, nil)
// (model3d/collisions.go:263-273): the number of triangles each ray crosses.
func (m *MeshCollider) RayCollisionCounts(rays []model3d.Ray) ([]int, error) {
	org := make([]float32, 0, len(rays)*3)
	dir := make([]float32, 0, len(rays)*3)
	for _, r := range rays {
		org = append(org, float32(r.Origin.X), float32(r.Origin.Y), float32(r.Origin.Z))
		dir = append(dir, float32(r.Direction.X), float32(r.Direction.Y), float32(r.Direction.Z))
	}
	counts := make([]int32, len(rays))
	var cp *C.int32_t
	if len(counts) > 0 {
		cp = (*C.int32_t)(unsafe.Pointer(&counts[0]))
	}
	if err := call(func() C.int32_t { return C.m3d_mesh_ray_collision_counts(m.h, fptr(org), fptr(dir), C.int64_t(len(rays)), cp, nil) }); err != nil {
		return nil, err
	}
	res := make([]int, len(rays))
	for i, c := range counts {
		res[i] = int(c)
	}
	return res, nil
}

// AllRayCollisions is the batched Collider.RayCollisions(r, f) with the collisions delivered
// (model3d/collisions.go:263-273, primitives.go:189-196): res[i] holds the collisions of rays[i]
// in order of Scale; Extra is a *model3d.TriangleCollision into the collider's own triangles.
func (m *MeshCollider) AllRayCollisions(rays []model3d.Ray) ([][]model3d.RayCollision, error) {
	n := len(rays)
	org := make([]float32, 0, n*3)
	dir := make([]float32, 0, n*3)
	for _, r := range rays {
		org = append(org, float32(r.Origin.X), float32(r.Origin.Y), float32(r.Origin.Z))
		dir = append(dir, float32(r.Direction.X), float32(r.Direction.Y), float32(r.Direction.Z))
	}
	offsets := make([]int64, n+1)
	op := (*C.int64_t)(unsafe.Pointer(&offsets[0]))
	// first call sizes the outputs (capacity 0), second call fills them
	if err := call(func() C.int32_t { return C.m3d_mesh_ray_collisions(m.h, fptr(org), fptr(dir), C.int64_t(n), 0, op,
		nil, nil, nil, nil, nil) }); err != nil {
		return nil, err
	}
	total := int(offsets[n])
	res := make([][]model3d.RayCollision, n)
	if total == 0 {
		return res, nil
	}
	ts := make([]float32, total)
	prim := make([]int32, total)
	normal := make([]float32, total*3)
	bary := make([]float32, total*3)
	if err := call(func() C.int32_t { return C.m3d_mesh_ray_collisions(m.h, fptr(org), fptr(dir), C.int64_t(n), C.int64_t(total), op,
		fptr(ts), (*C.int32_t)(unsafe.Pointer(&prim[0])), fptr(normal), fptr(bary), nil) }); err != nil {
		return nil, err
	}
	for i := 0; i < n; i++ {
		for k := offsets[i]; k < offsets[i+1]; k++ {
			res[i] = append(res[i], model3d.RayCollision{
				Scale:  float64(ts[k]),
				Normal: model3d.XYZ(float64(normal[3*k]), float64(normal[3*k+1]), float64(normal[3*k+2])),
				Extra: &model3d.TriangleCollision{
					Triangle:    m.Triangles[prim[k]],
					Barycentric: [3]float64{float64(bary[3*k]), float64(bary[3*k+1]), float64(bary[3*k+2])},
				},
			})
		}
	}
	return res, nil
}

// RayCollisions implements model3d.Collider (model3d/collisions.go:263-273): f, if not nil, is
// called on the calling goroutine for every triangle the ray crosses (batch of one).
func (m *MeshCollider) RayCollisions(r *model3d.Ray, f func(model3d.RayCollision)) int {
	if f == nil {
		c, err := m.RayCollisionCounts([]model3d.Ray{*r})
		if err != nil {
			panic(err)
		}
		return c[0]
	}
	res, err := m.AllRayCollisions([]model3d.Ray{*r})
	if err != nil {
		panic(err)
	}
	for _, c := range res[0] {
		f(c)
	}
	return len(res[0])
}

// SphereCollisions is the batched Collider.SphereCollision (model3d/collisions.go:292-303).
func (m *MeshCollider) SphereCollisions(centers []model3d.Coord3D, radii []float64) ([]bool, error) {
	if len(centers) != len(radii) {
		return nil, errors.New("gpu3d: centers and radii differ in length")
	}
	rad := make([]float32, len(radii))
	for i, r := range radii {
		rad[i] = float32(r)
	}
	out := make([]uint8, len(centers))
	var op *C.uint8_t
	if len(out) > 0 {
		op = (*C.uint8_t)(unsafe.Pointer(&out[0]))
	}
	flat := flatCoords(centers)
	if err := call(func() C.int32_t { return C.m3d_mesh_sphere_collisions(m.h, fptr(flat), fptr(rad), C.int64_t(len(centers)), op, nil) }); err != nil {
		return nil, err
	}
	res := make([]bool, len(out))
	for i, v := range out {
		res[i] = v != 0
	}
	return res, nil
}

// SphereCollision implements model3d.Collider with a batch of one.
func (m *MeshCollider) SphereCollision(c model3d.Coord3D, r float64) bool {
	res, err := m.SphereCollisions([]model3d.Coord3D{c}, []float64{r})
	if err != nil {
		panic(err)
	}
	return res[0]
}

// Contains is the batched model3d.ColliderContains(m, p, margin) (collisions.go:113-134).
func (m *MeshCollider) Contains(points []model3d.Coord3D, margin float64) ([]bool, error) {
	out := make([]uint8, len(points))
	var op *C.uint8_t
	if len(out) > 0 {
		op = (*C.uint8_t)(unsafe.Pointer(&out[0]))
	}
	flat := flatCoords(points)
	if err := call(func() C.int32_t { return C.m3d_mesh_contains(m.h, fptr(flat), C.int64_t(len(points)), C.double(margin), op, nil) }); err != nil {
		return nil, err
	}
	res := make([]bool, len(out))
	for i, v := range out {
		res[i] = v != 0
	}
	return res, nil
}

// MeshSDF implements model3d.FaceSDF (model3d/sdf.go:44-53) on the collider's device hierarchy;
// it replaces model3d.MeshToSDF (sdf.go:186-240).
type MeshSDF struct{ *MeshCollider }

// MeshToSDF wraps a GPU collider; the SDF and the collider share one device BVH.
func MeshToSDF(m *MeshCollider) (*MeshSDF, error) {
	if len(m.Triangles) == 0 {
		return nil, errors.New("gpu3d: cannot create empty SDF")
	}
	return &MeshSDF{m}, nil
}

// FaceSDFs is the batched FaceSDF: nearest face, nearest point and signed distance per point.
func (s *MeshSDF) FaceSDFs(points []model3d.Coord3D) ([]*model3d.Triangle, []model3d.Coord3D, []float64, error) {
	n := len(points)
	sdf := make([]float32, n)
	cp := make([]float32, 3*n)
	face := make([]int32, n)
	var fp *C.int32_t
	if n > 0 {
		fp = (*C.int32_t)(unsafe.Pointer(&face[0]))
	}
	flat := flatCoords(points)
	if err := call(func() C.int32_t { return C.m3d_mesh_sdf(s.h, fptr(flat), C.int64_t(n), fptr(sdf), fptr(cp), fp, nil, nil) }); err != nil {
		return nil, nil, nil, err
	}
	tris := make([]*model3d.Triangle, n)
	pts := make([]model3d.Coord3D, n)
	dists := make([]float64, n)
	for i := 0; i < n; i++ {
		tris[i] = s.Triangles[face[i]]
		pts[i] = model3d.XYZ(float64(cp[3*i]), float64(cp[3*i+1]), float64(cp[3*i+2]))
		dists[i] = float64(sdf[i])
	}
	return tris, pts, dists, nil
}

// FaceSDF implements model3d.FaceSDF with a batch of one.
func (s *MeshSDF) FaceSDF(c model3d.Coord3D) (*model3d.Triangle, model3d.Coord3D, float64) {
	t, p, d, err := s.FaceSDFs([]model3d.Coord3D{c})
	if err != nil {
		panic(err)
	}
	return t[0], p[0], d[0]
}

// SDF implements model3d.SDF.
func (s *MeshSDF) SDF(c model3d.Coord3D) float64 {
	_, _, d := s.FaceSDF(c)
	return d
}

// PointSDF implements model3d.PointSDF.
func (s *MeshSDF) PointSDF(c model3d.Coord3D) (model3d.Coord3D, float64) {
	_, p, d := s.FaceSDF(c)
	return p, d
}

// NormalSDF implements model3d.NormalSDF.
func (s *MeshSDF) NormalSDF(c model3d.Coord3D) (model3d.Coord3D, float64) {
	t, _, d := s.FaceSDF(c)
	return t.Normal(), d
}
