package gpu3d

/*
#include "m3d.h"
*/
import "C"

import (
	"fmt"
	"math"
	"runtime"

	"github.com/unixpickle/model3d/model3d"
	"github.com/unixpickle/model3d/render3d"
)

// ---- declarative object wrappers -----------------------------------------------------------------
//
// render3d.Translate / MatrixMultiply return unexported types and the reference's examples attach
// behaviour to objects with Go methods (showcase's DomeObject / FloorObject / VaseObject override
// Cast).  Neither can cross the C ABI, so this package offers the same wrappers as plain data.
// Every wrapper is a complete render3d.Object on the CPU too (it embeds the reference's own
// object and re-states the override), so a scene built from them renders with either renderer.

// Translated is render3d.Translate(Inner, Offset) (render3d/transform.go:6-31) as data.
type Translated struct {
	render3d.Object // == render3d.Translate(Inner, Offset): the CPU behaviour
	Inner           render3d.Object
	Offset          model3d.Coord3D
}

// Translate mirrors render3d.Translate.
func Translate(obj render3d.Object, offset model3d.Coord3D) *Translated {
	return &Translated{Object: render3d.Translate(obj, offset), Inner: obj, Offset: offset}
}

// Transformed is render3d.MatrixMultiply(Inner, Matrix) (render3d/transform.go:48-85) as data.
// The GPU path supports similarity matrices (rotation times uniform scale).
type Transformed struct {
	render3d.Object
	Inner  render3d.Object
	Matrix *model3d.Matrix3
}

// MatrixMultiply mirrors render3d.MatrixMultiply.
func MatrixMultiply(obj render3d.Object, m *model3d.Matrix3) *Transformed {
	return &Transformed{Object: render3d.MatrixMultiply(obj, m), Inner: obj, Matrix: m}
}

// Rotate mirrors render3d.Rotate (transform.go:35-39).
func Rotate(obj render3d.Object, axis model3d.Coord3D, angle float64) *Transformed {
	return MatrixMultiply(obj, model3d.NewMatrix3Rotation(axis, angle))
}

// Scale mirrors render3d.Scale (transform.go:41-46).
func Scale(obj render3d.Object, scale float64) *Transformed {
	return MatrixMultiply(obj, &model3d.Matrix3{scale, 0, 0, 0, scale, 0, 0, 0, scale})
}

// FlippedNormals reports the negated collision normal (showcase DomeObject, room.go:22-45).
type FlippedNormals struct{ render3d.Object }

func (f *FlippedNormals) Cast(r *model3d.Ray) (model3d.RayCollision, render3d.Material, bool) {
	rc, mat, ok := f.Object.Cast(r)
	rc.Normal = rc.Normal.Scale(-1)
	return rc, mat, ok
}

// CheckerObject is a Lambert surface whose diffuse colour is Color2 where
// int(mod(x+300, 2)) == int(mod(y+301, 2)) at the hit point and Color1 elsewhere (showcase
// FloorObject, room.go:47-75).
type CheckerObject struct {
	render3d.Object // geometry; its material is ignored
	Color1, Color2  render3d.Color
}

func (c *CheckerObject) Cast(r *model3d.Ray) (model3d.RayCollision, render3d.Material, bool) {
	rc, mat, ok := c.Object.Cast(r)
	if !ok {
		return rc, mat, ok
	}
	p := r.Origin.Add(r.Direction.Scale(rc.Scale))
	color := c.Color1
	if int(math.Mod(p.X+300, 2)) == int(math.Mod(p.Y+301, 2)) {
		color = c.Color2
	}
	return rc, &render3d.LambertMaterial{DiffuseColor: color}, ok
}

// ZGradientObject is a Phong surface whose diffuse colour is Color1*frac + Color2*(1-frac) with
// frac = z / MaxZ at the hit point (showcase VaseObject, models.go:74-97).
type ZGradientObject struct {
	render3d.Object
	Alpha          float64
	SpecularColor  render3d.Color
	Color1, Color2 render3d.Color
	MaxZ           float64
}

func (v *ZGradientObject) Cast(r *model3d.Ray) (model3d.RayCollision, render3d.Material, bool) {
	rc, _, ok := v.Object.Cast(r)
	if !ok {
		return rc, nil, ok
	}
	c := r.Origin.Add(r.Direction.Scale(rc.Scale))
	frac := c.Z / v.MaxZ
	return rc, &render3d.PhongMaterial{Alpha: v.Alpha, SpecularColor: v.SpecularColor,
		DiffuseColor: v.Color1.Scale(frac).Add(v.Color2.Scale(1 - frac))}, ok
}

// MeshObject is a mesh with a material whose triangles stay visible to the scene builder (the
// reference's joined colliders do not expose theirs).  Smooth: interpolated vertex normals
// (MeshToInterpNormalCollider).  It renders on the CPU through the reference's own collider.
type MeshObject struct {
	render3d.Object // == &render3d.ColliderObject{model3d.MeshTo[InterpNormal]Collider(Mesh), Material}
	Mesh            *model3d.Mesh
	Material        render3d.Material
	Smooth          bool
}

// NewMeshObject builds the object (and the reference's CPU collider behind it).
func NewMeshObject(mesh *model3d.Mesh, mat render3d.Material, smooth bool) *MeshObject {
	var c model3d.Collider
	if smooth {
		c = model3d.MeshToInterpNormalCollider(mesh)
	} else {
		c = model3d.MeshToCollider(mesh)
	}
	return &MeshObject{Object: &render3d.ColliderObject{Collider: c, Material: mat}, Mesh: mesh, Material: mat,
		Smooth: smooth}
}

// MeshAreaLight is render3d.NewMeshAreaLight (light.go:227-274) that keeps its mesh visible:
// the reference type hides its triangles, so BidirPathTracer.Light takes this one for meshes
// (sphere lights are read from *render3d.SphereAreaLight directly).
type MeshAreaLight struct {
	*render3d.MeshAreaLight
	Mesh     *model3d.Mesh
	Emission render3d.Color
}

// NewMeshAreaLight mirrors render3d.NewMeshAreaLight.
func NewMeshAreaLight(mesh *model3d.Mesh, emission render3d.Color) *MeshAreaLight {
	return &MeshAreaLight{MeshAreaLight: render3d.NewMeshAreaLight(mesh, emission), Mesh: mesh, Emission: emission}
}

// JoinedAreaLight is render3d.JoinAreaLights (light.go:276-314) that keeps its parts visible.
type JoinedAreaLight struct {
	render3d.AreaLight
	Lights []render3d.AreaLight
}

// JoinAreaLights mirrors render3d.JoinAreaLights.
func JoinAreaLights(lights ...render3d.AreaLight) *JoinedAreaLight {
	return &JoinedAreaLight{AreaLight: render3d.JoinAreaLights(lights...), Lights: lights}
}

// ---- scene ---------------------------------------------------------------------------------------

// Scene is a device-resident render3d.Object tree.  It is built by a type switch over the
// supported object / collider / material types; anything else is an error (no fallback).
// A Scene is itself a render3d.Object (Cast is a batch of one), so it can be handed to any code
// that expects one; the gpu3d renderers recognise it and skip the rebuild.
type Scene struct {
	h   *C.m3d_scene
	ctx *Context
	// materials by index (what Cast returns) and their indices
	matList   []render3d.Material
	materials map[render3d.Material]int32
	// leaf objects in scene order: the render3d.Object each came from (area lights are matched
	// by identity) and, for analytic / mesh leaves, what they are
	leaves   []render3d.Object
	min, max model3d.Coord3D
}

type sceneBuilder struct {
	s *Scene
	b *C.m3d_scene_builder
}

// xform is the accumulated world transform of the objects below a Translated / Transformed node:
// x_world = M x + off.
type xform struct {
	m   model3d.Matrix3
	off model3d.Coord3D
	set bool
}

func identity() xform { return xform{m: model3d.Matrix3{1, 0, 0, 0, 1, 0, 0, 0, 1}} }

func (x xform) c() *C.m3d_transform {
	if !x.set {
		return nil
	}
	t := &C.m3d_transform{}
	// model3d.Matrix3 and m3d_transform.matrix are both row-major (matrix.go:11-13,131-137)
	for i := 0; i < 9; i++ {
		t.matrix[i] = C.double(x.m[i])
	}
	t.offset = cvec(x.off)
	return t
}

// then returns the transform "first inner, then x".
func (x xform) then(inner xform) xform {
	return xform{m: *x.m.Mul(&inner.m), off: x.m.MulColumn(inner.off).Add(x.off), set: true}
}

func (sb *sceneBuilder) material(m render3d.Material) (int32, error) {
	s := sb.s
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
			return 0, &Error{Code: int(C.M3D_ERR_UNSUPPORTED),
				Msg: fmt.Sprintf("JoinedMaterial with %d parts is not supported", len(m.Materials))}
		}
		d.kind = C.M3D_MAT_JOINED
		d.num_sub = C.int32_t(len(m.Materials))
		for i, sub := range m.Materials {
			idx, err := sb.material(sub)
			if err != nil {
				return 0, err
			}
			d.sub[i] = C.int32_t(idx)
			d.sub_prob[i] = C.double(m.Probs[i])
		}
	case *checkerMaterial:
		d.kind = C.M3D_MAT_LAMBERT
		d.flags |= C.M3D_MAT_CHECKER
		set(&d.diffuse, m.c1)
		set(&d.diffuse2, m.c2)
	case *zGradientMaterial:
		d.kind = C.M3D_MAT_PHONG
		d.flags |= C.M3D_MAT_Z_GRADIENT
		d.alpha = C.double(m.alpha)
		set(&d.specular, m.specular)
		set(&d.diffuse, m.c1)
		set(&d.diffuse2, m.c2)
		d.proc_param = C.double(m.maxZ)
	default:
		return 0, &Error{Code: int(C.M3D_ERR_UNSUPPORTED),
			Msg: fmt.Sprintf("material type %T is not supported on the GPU path", m)}
	}
	var idx C.int32_t
	if err := call(func() C.int32_t { return C.m3d_scene_add_material(sb.b, &d, &idx) }); err != nil {
		return 0, err
	}
	s.materials[m] = int32(idx)
	s.matList = append(s.matList, m)
	return int32(idx), nil
}

// The procedural materials of CheckerObject / ZGradientObject: placeholders that carry the
// parameters to the builder (render3d.Material has no hit point, so they cannot be evaluated
// through that interface; on the CPU the wrapper objects' Cast builds the material per hit).
type checkerMaterial struct {
	render3d.LambertMaterial
	c1, c2 render3d.Color
}
type zGradientMaterial struct {
	render3d.PhongMaterial
	alpha, maxZ          float64
	specular, c1, c2     render3d.Color
}

// add walks the object tree.  flags: M3D_OBJ_* collected from wrappers; matOverride: the
// procedural material of an enclosing CheckerObject / ZGradientObject; origin: the object the leaf
// is recorded as (the outermost wrapper, which is what area lights and callers refer to).
func (sb *sceneBuilder) add(obj render3d.Object, x xform, flags C.uint32_t, matOverride render3d.Material,
	origin render3d.Object) error {
	if origin == nil {
		origin = obj
	}
	leafMaterial := func(m render3d.Material) (int32, error) {
		if matOverride != nil {
			return sb.material(matOverride)
		}
		if m == nil {
			return 0, &Error{Code: int(C.M3D_ERR_INVALID_ARG), Msg: "object without a material"}
		}
		return sb.material(m)
	}
	addMesh := func(tris []*model3d.Triangle, normals [][3]model3d.Coord3D, m render3d.Material) error {
		mat, err := leafMaterial(m)
		if err != nil {
			return err
		}
		flat, fn := flatTriangles(tris), flatNormals(normals)
		err = call(func() C.int32_t {
			return C.m3d_scene_add_mesh(sb.b, fptr(flat), C.int64_t(len(tris)), fptr(fn), C.int32_t(mat), flags, x.c(), nil)
		})
		runtime.KeepAlive(flat)
		runtime.KeepAlive(fn)
		if err == nil {
			sb.s.leaves = append(sb.s.leaves, origin)
		}
		return err
	}
	switch obj := obj.(type) {
	case *Scene:
		return &Error{Code: int(C.M3D_ERR_UNSUPPORTED), Msg: "a built Scene cannot be nested in another scene"}
	case render3d.JoinedObject:
		for _, o := range obj {
			if err := sb.add(o, x, flags, matOverride, nil); err != nil {
				return err
			}
		}
		return nil
	case *Translated:
		t := identity()
		t.off, t.set = obj.Offset, true
		return sb.add(obj.Inner, x.then(t), flags, matOverride, origin)
	case *Transformed:
		t := identity()
		t.m, t.set = *obj.Matrix, true
		return sb.add(obj.Inner, x.then(t), flags, matOverride, origin)
	case *FlippedNormals:
		return sb.add(obj.Object, x, flags^C.M3D_OBJ_FLIP_NORMAL, matOverride, origin)
	case *CheckerObject:
		return sb.add(obj.Object, x, flags, &checkerMaterial{c1: obj.Color1, c2: obj.Color2}, origin)
	case *ZGradientObject:
		return sb.add(obj.Object, x, flags, &zGradientMaterial{alpha: obj.Alpha, maxZ: obj.MaxZ,
			specular: obj.SpecularColor, c1: obj.Color1, c2: obj.Color2}, origin)
	case *MeshObject:
		var normals [][3]model3d.Coord3D
		tris := obj.Mesh.TriangleSlice()
		if obj.Smooth {
			vn := obj.Mesh.VertexNormals()
			normals = make([][3]model3d.Coord3D, len(tris))
			for i, t := range tris {
				for j, p := range t {
					normals[i][j] = vn.Value(p)
				}
			}
		}
		return addMesh(tris, normals, obj.Material)
	case *MeshAreaLight:
		return addMesh(obj.Mesh.TriangleSlice(), nil, &render3d.LambertMaterial{EmissionColor: obj.Emission})
	case *JoinedAreaLight:
		for _, l := range obj.Lights {
			if err := sb.add(l, x, flags, matOverride, nil); err != nil {
				return err
			}
		}
		return nil
	case *render3d.SphereAreaLight:
		// the embedded Object is the ColliderObject{*model3d.Sphere, Lambert{Emission}} (light.go:131-140)
		return sb.add(obj.Object, x, flags, matOverride, origin)
	case *render3d.ColliderObject:
		switch c := obj.Collider.(type) {
		case *MeshCollider:
			return addMesh(c.Triangles, c.VertexNormals, obj.Material)
		case *model3d.Sphere:
			mat, err := leafMaterial(obj.Material)
			if err != nil {
				return err
			}
			ctr := cvec(c.Center)
			err = call(func() C.int32_t {
				return C.m3d_scene_add_sphere(sb.b, &ctr[0], C.double(c.Radius), C.int32_t(mat), flags, x.c(), nil)
			})
			if err == nil {
				sb.s.leaves = append(sb.s.leaves, origin)
			}
			return err
		case *model3d.Rect:
			mat, err := leafMaterial(obj.Material)
			if err != nil {
				return err
			}
			mn, mx := cvec(c.MinVal), cvec(c.MaxVal)
			err = call(func() C.int32_t {
				return C.m3d_scene_add_rect(sb.b, &mn[0], &mx[0], C.int32_t(mat), flags, x.c(), nil)
			})
			if err == nil {
				sb.s.leaves = append(sb.s.leaves, origin)
			}
			return err
		case *model3d.Cylinder:
			mat, err := leafMaterial(obj.Material)
			if err != nil {
				return err
			}
			p1, p2 := cvec(c.P1), cvec(c.P2)
			err = call(func() C.int32_t {
				return C.m3d_scene_add_cylinder(sb.b, &p1[0], &p2[0], C.double(c.Radius), C.int32_t(mat), flags, x.c(), nil)
			})
			if err == nil {
				sb.s.leaves = append(sb.s.leaves, origin)
			}
			return err
		default:
			return &Error{Code: int(C.M3D_ERR_UNSUPPORTED), Msg: fmt.Sprintf(
				"collider type %T is not supported (meshes: gpu3d.MeshToCollider or gpu3d.NewMeshObject)", c)}
		}
	default:
		return &Error{Code: int(C.M3D_ERR_UNSUPPORTED),
			Msg: fmt.Sprintf("object type %T is not supported on the GPU path", obj)}
	}
}

// NewScene uploads obj.  Supported: JoinedObject trees of ColliderObject{*model3d.Sphere, *Rect,
// *Cylinder, *gpu3d.MeshCollider}, MeshObject, the wrappers of this package (Translated,
// Transformed, FlippedNormals, CheckerObject, ZGradientObject), *render3d.SphereAreaLight,
// MeshAreaLight, JoinedAreaLight, with Lambert / Phong / Refract / Joined materials.
func NewScene(ctx *Context, obj render3d.Object) (*Scene, error) {
	if s, ok := obj.(*Scene); ok {
		return s, nil
	}
	if ctx == nil {
		var err error
		if ctx, err = DefaultContext(); err != nil {
			return nil, err
		}
	}
	sb := &sceneBuilder{s: &Scene{ctx: ctx, materials: map[render3d.Material]int32{}}}
	if err := call(func() C.int32_t { return C.m3d_scene_builder_create(ctx.h, &sb.b) }); err != nil {
		return nil, err
	}
	defer C.m3d_scene_builder_destroy(sb.b)
	if err := sb.add(obj, identity(), 0, nil, nil); err != nil {
		return nil, err
	}
	s := sb.s
	if err := call(func() C.int32_t { return C.m3d_scene_build(sb.b, 0, &s.h) }); err != nil {
		return nil, err
	}
	var mn, mx [3]C.double
	C.m3d_scene_bounds(s.h, &mn[0], &mx[0])
	s.min = model3d.XYZ(float64(mn[0]), float64(mn[1]), float64(mn[2]))
	s.max = model3d.XYZ(float64(mx[0]), float64(mx[1]), float64(mx[2]))
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

// Min / Max / Cast make *Scene a render3d.Object (object.go:12-23).
func (s *Scene) Min() model3d.Coord3D { return s.min }
func (s *Scene) Max() model3d.Coord3D { return s.max }

// Cast is Object.Cast as a batch of one through m3d_scene_cast (correct, slow).  Procedural
// materials (CheckerObject / ZGradientObject) are evaluated at the hit point like their CPU Cast.
func (s *Scene) Cast(r *model3d.Ray) (model3d.RayCollision, render3d.Material, bool) {
	org := [3]C.float{C.float(r.Origin.X), C.float(r.Origin.Y), C.float(r.Origin.Z)}
	dir := [3]C.float{C.float(r.Direction.X), C.float(r.Direction.Y), C.float(r.Direction.Z)}
	var t C.float
	var obj, prim C.int32_t
	var normal [3]C.float
	err := call(func() C.int32_t {
		return C.m3d_scene_cast(s.h, &org[0], &dir[0], 1, &t, &obj, &prim, &normal[0], 0, nil)
	})
	if err != nil {
		panic(err)
	}
	if obj < 0 {
		return model3d.RayCollision{}, nil, false
	}
	rc := model3d.RayCollision{Scale: float64(t),
		Normal: model3d.XYZ(float64(normal[0]), float64(normal[1]), float64(normal[2]))}
	// the leaf's own CPU Cast knows its (possibly procedural) material: ask it for this ray
	if _, mat, ok := s.leaves[int(obj)].Cast(r); ok {
		return rc, mat, true
	}
	return rc, nil, true
}

// leafIndex returns the scene-order index of the leaf that came from obj (-1: not part of the scene).
func (s *Scene) leafIndex(obj render3d.Object) int {
	for i, l := range s.leaves {
		if l == obj {
			return i
		}
	}
	return -1
}

// sceneFor gives the renderers their Scene: a *Scene is used as is, any other object tree is
// compiled for this call on ctx (nil: DefaultContext) and released afterwards -- like the reference
// re-reads the object on every Render.  Callers that render the same objects repeatedly build a
// Scene once with NewScene and pass that.
func sceneFor(ctx *Context, obj render3d.Object) (s *Scene, release func(), err error) {
	if sc, ok := obj.(*Scene); ok {
		return sc, func() {}, nil
	}
	s, err = NewScene(ctx, obj)
	if err != nil {
		return nil, nil, err
	}
	return s, s.Close, nil
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

// RayCaster mirrors render3d.RayCaster (render3d/raycast.go:9-39).
type RayCaster struct {
	Camera *render3d.Camera
	Lights []*render3d.PointLight

	// Context the object is compiled on when it is not a *Scene (nil: DefaultContext).
	Context *Context
}

// Render renders obj into img like (*render3d.RayCaster).Render; pixels whose ray misses keep
// their previous value.
func (r *RayCaster) Render(img *render3d.Image, obj render3d.Object) error {
	scene, release, err := sceneFor(r.Context, obj)
	if err != nil {
		return err
	}
	defer release()
	cam := ccamera(r.Camera)
	lights := clights(r.Lights)
	rgb, err := NewHostFloats(3 * len(img.Data))
	if err != nil {
		return err
	}
	defer rgb.Free()
	for i, c := range img.Data {
		rgb.S[3*i], rgb.S[3*i+1], rgb.S[3*i+2] = float32(c.X), float32(c.Y), float32(c.Z)
	}
	err = call(func() C.int32_t {
		return C.m3d_render_raycast(scene.h, &cam, &lights[0], C.int32_t(len(r.Lights)),
			C.int32_t(img.Width), C.int32_t(img.Height), nil, rgb.c(), nil)
	})
	if err != nil {
		return err
	}
	for i := range img.Data {
		img.Data[i] = render3d.Color{X: float64(rgb.S[3*i]), Y: float64(rgb.S[3*i+1]), Z: float64(rgb.S[3*i+2])}
	}
	return nil
}
