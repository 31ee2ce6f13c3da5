package gpu3d

/*
#include "m3d.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"unsafe"

	"github.com/unixpickle/model3d/model3d"
	"github.com/unixpickle/model3d/render3d"
)

// Scene is a device-resident render3d.Object tree.  It is built by a type switch over the
// supported object / collider / material types; anything else is an error (no fallback).
type Scene struct {
	h         *C.m3d_scene
	materials map[render3d.Material]int32
}

// MeshObject tells NewScene which triangles a mesh collider was built from (the reference's
// joined colliders do not expose them).
type MeshObject struct {
	Mesh     *model3d.Mesh
	Material render3d.Material
}

func (m *MeshObject) Min() model3d.Coord3D { return m.Mesh.Min() }
func (m *MeshObject) Max() model3d.Coord3D { return m.Mesh.Max() }
func (m *MeshObject) Cast(r *model3d.Ray) (model3d.RayCollision, render3d.Material, bool) {
	panic("gpu3d.MeshObject is a scene description; render it with a gpu3d renderer")
}

func (s *Scene) material(b *C.m3d_scene_builder, m render3d.Material) (int32, error) {
	if idx, ok := s.materials[m]; ok {
		return idx, nil
	}
	var d C.m3d_material_desc
	set := func(dst *[3]C.double, c render3d.Color) { *dst = cvec(c) }
	switch m := m.(type) {
	case *render3d.LambertMaterial:
		d.kind = C.M3D_MAT_LAMBERT
		set(&d.diffuse, m.DiffuseColor)
		set(&d.ambient, m.AmbientColor)
		set(&d.emission, m.EmissionColor)
	case *render3d.PhongMaterial:
		d.kind = C.M3D_MAT_PHONG
		d.alpha = C.double(m.Alpha)
		set(&d.specular, m.SpecularColor)
		set(&d.diffuse, m.DiffuseColor)
		set(&d.emission, m.EmissionColor)
		set(&d.ambient, m.AmbientColor)
		if m.NoFluxCorrection {
			d.flags |= C.M3D_MAT_NO_FLUX_CORRECTION
		}
	case *render3d.RefractMaterial:
		d.kind = C.M3D_MAT_REFRACT
		d.index_of_refraction = C.double(m.IndexOfRefraction)
		set(&d.refract, m.RefractColor)
		set(&d.specular, m.SpecularColor)
	case *render3d.JoinedMaterial:
		if len(m.Materials) > C.M3D_MAX_SUBMATERIALS || len(m.Probs) != len(m.Materials) {
			return 0, fmt.Errorf("gpu3d: JoinedMaterial with %d parts is not supported", len(m.Materials))
		}
		d.kind = C.M3D_MAT_JOINED
		d.num_sub = C.int32_t(len(m.Materials))
		for i, sub := range m.Materials {
			idx, err := s.material(b, sub)
			if err != nil {
				return 0, err
			}
			d.sub[i] = C.int32_t(idx)
			d.sub_prob[i] = C.double(m.Probs[i])
		}
	default:
		return 0, fmt.Errorf("gpu3d: material type %T is not supported on the GPU path", m)
	}
	var idx C.int32_t
	if err := status(C.m3d_scene_add_material(b, &d, &idx)); err != nil {
		return 0, err
	}
	s.materials[m] = int32(idx)
	return int32(idx), nil
}

func (s *Scene) add(b *C.m3d_scene_builder, obj render3d.Object) error {
	switch obj := obj.(type) {
	case render3d.JoinedObject:
		for _, o := range obj {
			if err := s.add(b, o); err != nil {
				return err
			}
		}
		return nil
	case *MeshObject:
		mat, err := s.material(b, obj.Material)
		if err != nil {
			return err
		}
		tris := obj.Mesh.TriangleSlice()
		flat := make([]float32, 0, 9*len(tris))
		for _, t := range tris {
			for _, p := range t {
				flat = append(flat, float32(p.X), float32(p.Y), float32(p.Z))
			}
		}
		return status(C.m3d_scene_add_mesh(b, (*C.float)(unsafe.Pointer(&flat[0])), C.int64_t(len(tris)),
			nil, C.int32_t(mat), 0, nil, nil))
	case *render3d.ColliderObject:
		mat, err := s.material(b, obj.Material)
		if err != nil {
			return err
		}
		switch c := obj.Collider.(type) {
		case *model3d.Sphere:
			ctr := cvec(c.Center)
			return status(C.m3d_scene_add_sphere(b, &ctr[0], C.double(c.Radius), C.int32_t(mat), 0, nil, nil))
		case *model3d.Rect:
			mn, mx := cvec(c.MinVal), cvec(c.MaxVal)
			return status(C.m3d_scene_add_rect(b, &mn[0], &mx[0], C.int32_t(mat), 0, nil, nil))
		case *model3d.Cylinder:
			p1, p2 := cvec(c.P1), cvec(c.P2)
			return status(C.m3d_scene_add_cylinder(b, &p1[0], &p2[0], C.double(c.Radius), C.int32_t(mat), 0, nil, nil))
		default:
			return fmt.Errorf("gpu3d: collider type %T is not supported (wrap meshes in gpu3d.MeshObject)", c)
		}
	default:
		return fmt.Errorf("gpu3d: object type %T is not supported on the GPU path", obj)
	}
}

// NewScene uploads obj.  Supported: JoinedObject of ColliderObject{Sphere,Rect,Cylinder},
// MeshObject, with Lambert / Phong / Refract / Joined materials.
func NewScene(ctx *Context, obj render3d.Object) (*Scene, error) {
	var b *C.m3d_scene_builder
	if err := status(C.m3d_scene_builder_create(ctx.h, &b)); err != nil {
		return nil, err
	}
	defer C.m3d_scene_builder_destroy(b)
	s := &Scene{materials: map[render3d.Material]int32{}}
	if err := s.add(b, obj); err != nil {
		return nil, err
	}
	if err := status(C.m3d_scene_build(b, 0, &s.h)); err != nil {
		return nil, err
	}
	runtime.SetFinalizer(s, (*Scene).Close)
	return s, nil
}

// Close frees the device scene.
func (s *Scene) Close() {
	if s.h != nil {
		C.m3d_scene_destroy(s.h)
		s.h = nil
	}
}

// RayCaster mirrors render3d.RayCaster (render3d/raycast.go:9-39).
type RayCaster struct {
	Camera *render3d.Camera
	Lights []*render3d.PointLight
}

// Render renders the scene into img like (*render3d.RayCaster).Render; pixels whose ray
// misses keep their previous value.
func (r *RayCaster) Render(img *render3d.Image, scene *Scene) error {
	cam := ccamera(r.Camera)
	lights := make([]C.m3d_point_light, len(r.Lights)+1)
	for i, l := range r.Lights {
		lights[i].origin = cvec(l.Origin)
		lights[i].color = cvec(l.Color)
		if l.QuadDropoff {
			lights[i].quad_dropoff = 1
		}
	}
	rgb := make([]float32, 3*len(img.Data))
	for i, c := range img.Data {
		rgb[3*i], rgb[3*i+1], rgb[3*i+2] = float32(c.X), float32(c.Y), float32(c.Z)
	}
	err := status(C.m3d_render_raycast(scene.h, &cam, &lights[0], C.int32_t(len(r.Lights)),
		C.int32_t(img.Width), C.int32_t(img.Height), nil, (*C.float)(unsafe.Pointer(&rgb[0])), nil))
	if err != nil {
		return err
	}
	for i := range img.Data {
		img.Data[i] = render3d.Color{X: float64(rgb[3*i]), Y: float64(rgb[3*i+1]), Z: float64(rgb[3*i+2])}
	}
	return nil
}
