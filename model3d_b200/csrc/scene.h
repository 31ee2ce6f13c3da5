// Device-side scene description: replaces a render3d.Object tree (JoinedObject of
// ColliderObjects, render3d/object.go:26-153) and its materials (render3d/material.go).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace m3d {

// SHAPE_INSTANCE: a mesh collider (its own wide BVH, object space, one copy of the triangles)
// under a similarity transform -- render3d.Translate / MatrixMultiply of a shared collider
// (render3d/transform.go:6-85, examples/renderings/golf_balls/main.go:25-39)
enum ShapeKind : int32_t { SHAPE_SPHERE = 1, SHAPE_RECT = 2, SHAPE_CYLINDER = 3, SHAPE_INSTANCE = 4 };

// Analytic collider (model3d/shapes.go Sphere :12, Rect :142, Cylinder :543), float64 like
// the reference: there are only a handful per scene and they are evaluated in the coherent
// finish pass.
struct DeviceShape {
  int32_t kind;
  int32_t object;  // scene object index
  double p0[3];    // sphere centre | rect min | cylinder P1 | instance: world bounds min
  double p1[3];    //               | rect max | cylinder P2 | instance: world bounds max
  double radius;
  int32_t instance;  // SHAPE_INSTANCE: index into DeviceScene::instances
  int32_t _pad;
};

// One instance of a mesh collider: x_world = fwd * x_object + off (similarity: fwd = s * R).
struct DeviceInstance {
  float inv[9];   // fwd^-1, row-major: world direction / offset point -> object space
  float fwd[9];   // for normals: n_world = normalize(fwd * n_object) (transform.go:81-83)
  float off[3];
  float size;     // largest extent of the world bounds (self-intersection floor)
  DeviceBVH blas; // the mesh's own hierarchy and triangles (shared by all of its instances)
};

struct DeviceMaterial {
  int32_t kind;
  uint32_t flags;
  float diffuse[3], specular[3], emission[3], ambient[3], refract[3], diffuse2[3];
  float alpha, ior, proc_param;
  int32_t num_sub;
  int32_t sub[4];
  float sub_prob[4];
};

struct DeviceObject {
  int32_t material;
  uint32_t flags;  // M3D_OBJ_FLIP_NORMAL
};

struct DeviceScene {
  DeviceBVH bvh;  // all mesh objects merged, world space; triangle records carry object ids
  const DeviceShape *shapes = nullptr;
  int32_t num_shapes = 0;
  int32_t spheres_only = 0;  // every analytic shape is a sphere (kernels without the rect / cylinder code)
  const DeviceObject *objects = nullptr;
  int32_t num_objects = 0;
  const DeviceMaterial *materials = nullptr;
  int32_t num_materials = 0;
  // Object-level hierarchy (render3d.BVHToObject, object.go:172-185; FilteredObject :155-167):
  // when a scene has many analytic shapes / instances, the finish pass walks this wide BVH over
  // their bounds (leaf "triangles" are proxies whose prim id is the shape index) instead of
  // testing every shape for every ray.  nodes == nullptr: linear scan (a handful of shapes).
  DeviceBVH shape_bvh;
  const DeviceInstance *instances = nullptr;
  int32_t num_instances = 0;
};

// Scene trace = BVH traversal (trace_first_hit_kernel) + finish pass that also intersects
// the analytic shapes and applies JoinedObject's closest-wins rule (object.go:141-153).
// skip_ids: optional per-ray id of the surface the ray starts on (triangle index in leaf
// order, or -2-shape_index), ignored by the query (self-intersection guard for secondary
// rays; the reference uses a 1e-8 origin offset, raytrace.go:217-229).
struct SceneTraceLaunch {
  TraceLaunch t;
  const int32_t *skip_ids = nullptr;  // per ray: surface the ray starts on (see above) or -1
  int32_t *surf_ids = nullptr;        // optional out: surface that was hit (same encoding)
};
// BVH traversal + finish pass
void launch_trace_scene(const DeviceScene &scene, const SceneTraceLaunch &p, cudaStream_t stream);
// finish pass alone (after launch_trace_bvh_only)
void launch_finish_scene_hits(const DeviceScene &scene, const SceneTraceLaunch &p, cudaStream_t stream);

// Camera rays (render3d/camera.go:74-113; RayCaster passes W-1, H-1: raycast.go:16-18).
struct DeviceCamera {
  double origin[3];
  double x[3], y[3], z[3];  // scaled axes of Camera.axes()
  double cx, cy;
};
// rows [row_begin,row_end) of a W x H frame -> rays (idx = x + (y-row_begin)*W)
void launch_raygen_camera(const DeviceCamera &cam, int W, int row_begin, int row_end, float4 *org_tmin,
                          float4 *dir_tmax, cudaStream_t stream);

struct DevicePointLight {
  float origin[3];
  float color[3];
  int32_t quad_dropoff;
};
// RayCaster.Render body (raycast.go:25-37): writes rgb (3 floats / pixel) where the ray hit.
void launch_shade_raycast(const DeviceScene &scene, const DeviceCamera &cam, const DevicePointLight *lights,
                          int num_lights, const float4 *org_tmin, const float4 *dir_tmax,
                          const float4 *hit0, const float4 *hit1, int64_t n, float *rgb,
                          cudaStream_t stream);

// Image.Downsample (image.go:100-120): factor x factor box filter of a W x H x 3 image
void launch_downsample_image(const float *src, int W, int H, int factor, float *dst, cudaStream_t stream);

// colorSum/numSamples and optional sRGB-8 (ray_renderer.go:150, image.go:125-145)
void launch_finalize_image(const float *sum, int64_t num_values, float inv_samples, float *mean,
                           uint8_t *srgb8, cudaStream_t stream);

}  // namespace m3d
