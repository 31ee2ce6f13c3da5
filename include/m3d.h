/*
 * m3d.h -- C ABI of libm3dgpu: the B200 (sm_100a) implementation of model3d's
 * ray-tracing hot path.
 *
 * This is the drop-in boundary.  The reference (github.com/unixpickle/model3d, Go)
 * has no FFI; its seam is a set of Go interfaces.  Every entry point below names
 * the reference interface it replaces (file:line relative to the reference tree).
 * A Go maintainer binds these with cgo (see INTEGRATION.md and go/gpu3d/).
 *
 * Conventions
 *   - every function returns an int32 status (M3D_OK == 0); m3d_last_error() gives
 *     a thread-local, NUL-terminated message owned by the library;
 *   - host buffers are borrowed for the duration of the call only;
 *   - handles are opaque, immutable after build, freed with *_destroy;
 *   - bulk arrays (vertices, rays, hits, pixels) are float32 / int32;
 *     scalar parameters (camera, materials, lights) are float64 like the Go fields;
 *   - "triangle id" == index of the triangle in the caller's array
 *     (the reference identifies triangles by pointer, model3d/collisions.go:39-46);
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     fails with M3D_ERR_CUDA;
 *   - thread safety: every call on a context (or on a mesh / scene built from it) takes the
 *     context's lock, so handles may be shared between threads like the reference's Colliders
 *     and Objects ("safe for concurrency", model3d/collisions.go:51); calls on ONE context run one
 *     after another -- use one context per thread (or a multi-device context) for parallelism.
 *     m3d_last_error() is per thread: read it on the thread whose call failed.
 */
#ifndef M3D_H_
#define M3D_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define M3D_ABI_VERSION 5

typedef enum {
  M3D_OK = 0,
  M3D_ERR_INVALID_ARG = 1,
  M3D_ERR_UNSUPPORTED = 2, /* object / material / feature outside the GPU path */
  M3D_ERR_CUDA = 3,
  M3D_ERR_NCCL = 4,
  M3D_ERR_OOM = 5
} m3d_status;

typedef struct m3d_ctx m3d_ctx;     /* one CUDA device + stream pool            */
typedef struct m3d_mesh m3d_mesh;   /* triangle mesh + wide BVH, device resident */
typedef struct m3d_scene m3d_scene; /* objects + materials, device resident      */
typedef struct m3d_scene_builder m3d_scene_builder;

/* ---- library / context ------------------------------------------------- */

int32_t m3d_abi_version(void);
const char *m3d_last_error(void);

/* device < 0 -> current CUDA device. */
int32_t m3d_ctx_create(int32_t device, m3d_ctx **out);
/* A context over n devices of this node (devices[0] is the primary; NULL/n<=0 -> all visible
 * devices).  Replaces the reference's goroutine scheduler (render3d/concurrency.go:17-43,
 * ray_renderer.go:25-56) across GPUs: meshes and scenes built on it are replicated on every
 * device; m3d_mesh_first_ray_collisions splits the ray batch into contiguous slices,
 * m3d_render_raycast renders row bands, m3d_render_path / m3d_render_bidir shard the samples of
 * every pixel by index (adaptive renders: row bands), one host thread per device, and every
 * device's flush kernel adds its per-pixel sums straight into the primary's accumulator over
 * NVLink peer mappings (system-scope red.add) -- no collective afterwards.  The image does not
 * depend on the number of devices (Philox streams are keyed by pixel and sample index) beyond
 * float32 summation order.  Needs peer access between the primary and every other device,
 * else M3D_ERR_UNSUPPORTED. */
int32_t m3d_ctx_create_multi(const int32_t *devices, int32_t n, m3d_ctx **out);
int32_t m3d_ctx_num_devices(const m3d_ctx *ctx);
void m3d_ctx_destroy(m3d_ctx *ctx);
int32_t m3d_ctx_device(const m3d_ctx *ctx);
int32_t m3d_ctx_synchronize(m3d_ctx *ctx);
/* Releases the context's scratch buffers (ray batches, path state: the renderers size their
 * batches for up to 48-64 GB of a 180 GB device and keep the allocation for the next call).
 * Waits for the context's streams first; later calls allocate again on demand. */
int32_t m3d_ctx_trim(m3d_ctx *ctx);

/* ---- statistics ---------------------------------------------------------- */

/* Filled by trace / render calls when a non-NULL pointer is passed.
 * nodes_visited / tris_tested are only counted when M3D_TRACE_COUNTERS is set
 * (a separate kernel instantiation; never set in timed runs). */
typedef struct {
  int64_t rays;          /* rays traced (all kinds)                      */
  int64_t hits;          /* rays that hit something                      */
  int64_t nodes_visited; /* wide-BVH nodes fetched                       */
  int64_t tris_tested;   /* ray/triangle tests executed                  */
  double kernel_ms;      /* device time of the compute kernels           */
  double h2d_ms, d2h_ms; /* host<->device copy time (host-buffer calls)  */
  int64_t h2d_bytes, d2h_bytes;
  int64_t launches;      /* kernels of this library launched by the call */
  int64_t samples;       /* renderers: samples taken (adaptive mode stops per pixel) */
} m3d_stats;

/* ---- mesh collider ----------------------------------------------------------
 * Replaces model3d.MeshToCollider / GroupTriangles / GroupedTrianglesToCollider
 * (model3d/collisions.go:138-179, model3d/bvh.go:118-156) and
 * MeshToInterpNormalCollider (collisions.go:147-162) when vnormals != NULL.
 */
#define M3D_MESH_BUILD_HOST_SAH 0u   /* host binned-SAH build, collapsed to 8-wide */
#define M3D_MESH_BUILD_DEVICE_LBVH 1u /* device Morton / radix sort / Karras / refit binary
                                         tree, then the same 8-wide collapse on the host */
#define M3D_MESH_BUILD_DEVICE_COLLAPSE 2u /* whole build on the device: LBVH, cost-optimal
                                             8-wide collapse and node emission            */

typedef struct {
  int64_t num_triangles;
  int64_t num_nodes;     /* 8-wide compressed nodes                   */
  int64_t node_bytes;    /* bytes of one node record (80)             */
  int64_t tri_bytes;     /* bytes of one triangle record (48)         */
  int64_t device_bytes;  /* total HBM footprint                       */
  int32_t max_depth;
  double build_ms;
  double sah_cost;
} m3d_mesh_info;

/* tris: n*9 floats (v0 v1 v2 per triangle, caller order == triangle id).
 * vnormals: NULL or n*9 floats (per-corner normals, smooth shading). */
int32_t m3d_mesh_create(m3d_ctx *ctx, const float *tris, int64_t n,
                        const float *vnormals, uint32_t build_flags,
                        m3d_mesh **out);
void m3d_mesh_destroy(m3d_mesh *mesh);
int32_t m3d_mesh_get_info(const m3d_mesh *mesh, m3d_mesh_info *info);
/* Collider.Min()/Max()  (model3d/collisions.go:255-261) */
int32_t m3d_mesh_bounds(const m3d_mesh *mesh, double min_out[3], double max_out[3]);

#define M3D_TRACE_COUNTERS 1u   /* count nodes/triangles (slower; not for timing) */
#define M3D_TRACE_NO_REFINE 2u  /* skip the float64 final-hit refinement          */
#define M3D_TRACE_SHARED_ORIGIN 4u /* host-buffer call: `org` holds ONE origin (3 floats) shared by
                                      all rays (camera batches): halves the host->device bytes */

/* Batched Collider.FirstRayCollision (model3d/collisions.go:275-290 +
 * model3d/primitives.go:181-249).  Host buffers.
 *   org, dir : n*3 floats; dir is NOT normalised, t is in units of |dir|
 *   t        : n floats   (RayCollision.Scale; undefined on miss)
 *   prim     : n int32    (triangle id; -1 == miss)
 *   normal   : n*3 floats or NULL (RayCollision.Normal: flat, never flipped,
 *              or interpolated when the mesh has vnormals)
 *   bary     : n*3 floats or NULL (TriangleCollision.Barycentric)
 * Outputs the caller does not need may be NULL and are then neither unpacked nor copied back
 * (t + prim only: 8 bytes per ray device->host instead of 32).  The call runs a chunked copy /
 * compute pipeline; it reaches the PCIe rate only when the arrays are page-locked
 * (m3d_host_alloc / m3d_host_register below).  On a multi-device context the batch is split into
 * contiguous slices, one per GPU.
 */
int32_t m3d_mesh_first_ray_collisions(m3d_mesh *mesh, const float *org,
                                      const float *dir, int64_t n, float *t,
                                      int32_t *prim, float *normal, float *bary,
                                      uint32_t flags, m3d_stats *stats);

/* Batched Collider.RayCollisions(r, nil) (model3d/collisions.go:263-273, primitives.go:189-196):
 * counts[i] = number of triangles the forward half-line of ray i crosses
 * (m3d_mesh_ray_collisions below also delivers them). */
int32_t m3d_mesh_ray_collision_counts(m3d_mesh *mesh, const float *org, const float *dir,
                                      int64_t n, int32_t *counts, m3d_stats *stats);

/* Batched Collider.RayCollisions(r, f) with the collisions delivered (model3d/collisions.go:263-273,
 * primitives.go:189-196; the reference calls f once per crossed triangle).  Two-call idiom:
 *   offsets : n+1 int64, always written: exclusive prefix sum of the per-ray counts,
 *             offsets[n] == total number of collisions
 *   capacity: number of collisions the output arrays can hold.  If offsets[n] <= capacity the
 *             collisions of ray i are written at indices [offsets[i], offsets[i+1]) in order of
 *             increasing t (the reference's order is its BVH's traversal order, i.e. unspecified);
 *             otherwise only offsets is written (call with capacity 0 to size the arrays).
 *   t       : capacity floats   (RayCollision.Scale)
 *   prim    : capacity int32    (triangle id of TriangleCollision.Triangle)
 *   normal  : capacity*3 floats or NULL (RayCollision.Normal; interpolated with vnormals)
 *   bary    : capacity*3 floats or NULL (TriangleCollision.Barycentric)
 * Each collision is re-evaluated in float64 like the first-hit query. */
int32_t m3d_mesh_ray_collisions(m3d_mesh *mesh, const float *org, const float *dir, int64_t n,
                                int64_t capacity, int64_t *offsets, float *t, int32_t *prim,
                                float *normal, float *bary, m3d_stats *stats);

/* Batched model3d.ColliderContains(c, p, margin) (collisions.go:113-134) and with it
 * ColliderSolid.Contains (model3d/solid.go:256-300): odd number of triangles along the
 * reference's fixed probe direction from points[i]; margin > 0 additionally requires that no
 * triangle is closer than margin, margin < 0 also accepts outside points closer than -margin
 * (both through the nearest-triangle query below, == Collider.SphereCollision). */
int32_t m3d_mesh_contains(m3d_mesh *mesh, const float *points, int64_t n, double margin,
                          uint8_t *inside, m3d_stats *stats);

/* Batched MeshToSDF(mesh).FaceSDF / PointSDF / NormalSDF / SDF (model3d/sdf.go:186-311,
 * Triangle.Closest primitives.go:153-175).  Any output may be NULL.
 *   sdf     : n floats, distance to the nearest triangle, positive inside
 *             (ColliderSolid.Contains, solid.go:292-300), negative outside
 *   closest : n*3 floats, nearest point on the surface
 *   face    : n int32, triangle id of the nearest face (ties: any of the equidistant faces)
 *   normal  : n*3 floats, flat normal of that face (meshSDF.NormalSDF)
 * An empty mesh is M3D_ERR_INVALID_ARG (the reference panics, sdf.go:198-200). */
int32_t m3d_mesh_sdf(m3d_mesh *mesh, const float *points, int64_t n, float *sdf, float *closest,
                     int32_t *face, float *normal, m3d_stats *stats);

/* Batched Collider.SphereCollision(centers[i], radii[i]) (model3d/collisions.go:292-303,
 * primitives.go:253-279): collides[i] = 1 iff some triangle is closer than radii[i]. */
int32_t m3d_mesh_sphere_collisions(m3d_mesh *mesh, const float *centers, const float *radii,
                                   int64_t n, uint8_t *collides, m3d_stats *stats);

/* Same query on device-resident SoA buffers (what the renderers use internally
 * and what bench.py times with inputs already in HBM).
 *   d_org_tmin : n float4  (ox, oy, oz, tmin)
 *   d_dir_tmax : n float4  (dx, dy, dz, tmax)
 *   d_hit0     : n float4  (t, bary1, bary2, bits(prim))   prim == -1 on miss
 *   d_hit1     : n float4  (nx, ny, nz, bits(object))
 * stream: a cudaStream_t cast to void* (NULL == the context's stream). */
int32_t m3d_mesh_first_ray_collisions_device(m3d_mesh *mesh, const void *d_org_tmin,
                                             const void *d_dir_tmax, int64_t n,
                                             void *d_hit0, void *d_hit1,
                                             uint32_t flags, void *stream,
                                             m3d_stats *stats);

/* ---- scene -------------------------------------------------------------------
 * Replaces render3d.Object trees built from JoinedObject / ColliderObject /
 * Translate / MatrixMultiply (render3d/object.go:12-153, transform.go:6-85) over
 * model3d.Sphere / Rect / Cylinder / mesh colliders (model3d/shapes.go:35-93,
 * 177-247, 601-705) and render3d materials (render3d/material.go:119-479,554-631).
 * Anything else is M3D_ERR_UNSUPPORTED at build time.
 */
typedef enum {
  M3D_MAT_LAMBERT = 0, /* material.go:119-167 */
  M3D_MAT_PHONG = 1,   /* material.go:172-269 */
  M3D_MAT_REFRACT = 2, /* material.go:343-479 */
  M3D_MAT_JOINED = 3   /* material.go:554-631 */
} m3d_material_kind;

#define M3D_MAT_NO_FLUX_CORRECTION 1u /* PhongMaterial.NoFluxCorrection */
#define M3D_MAT_CHECKER 2u   /* showcase FloorObject: diffuse = checker(p.x,p.y) ? c0 : c1 */
#define M3D_MAT_Z_GRADIENT 4u /* showcase VaseObject: diffuse = lerp over z/max_z */
#define M3D_MAX_SUBMATERIALS 4

typedef struct {
  int32_t kind;
  uint32_t flags;
  double diffuse[3];
  double specular[3];
  double emission[3];
  double ambient[3];
  double refract[3];
  double alpha;               /* Phong exponent */
  double index_of_refraction; /* RefractMaterial.IndexOfRefraction */
  /* procedural variants: second colour + scalar (checker alt colour, gradient max_z) */
  double diffuse2[3];
  double proc_param;
  /* JoinedMaterial */
  int32_t num_sub;
  int32_t sub[M3D_MAX_SUBMATERIALS];
  double sub_prob[M3D_MAX_SUBMATERIALS];
} m3d_material_desc;

#define M3D_OBJ_FLIP_NORMAL 1u /* showcase DomeObject: reported normal negated */

/* Optional rigid/affine wrapper: x_world = matrix * x_object + offset
 * (render3d.Translate / MatrixMultiply, transform.go:6-85).  Row-major like
 * the wrapper passes it; NULL == identity. */
typedef struct {
  double matrix[9];
  double offset[3];
} m3d_transform;

int32_t m3d_scene_builder_create(m3d_ctx *ctx, m3d_scene_builder **out);
void m3d_scene_builder_destroy(m3d_scene_builder *b);
/* each add_* returns the new index in *index_out (may be NULL) */
int32_t m3d_scene_add_material(m3d_scene_builder *b, const m3d_material_desc *mat,
                               int32_t *index_out);
int32_t m3d_scene_add_mesh(m3d_scene_builder *b, const float *tris, int64_t n,
                           const float *vnormals, int32_t material, uint32_t flags,
                           const m3d_transform *xf, int32_t *index_out);
int32_t m3d_scene_add_sphere(m3d_scene_builder *b, const double center[3], double radius,
                             int32_t material, uint32_t flags, const m3d_transform *xf,
                             int32_t *index_out);
int32_t m3d_scene_add_rect(m3d_scene_builder *b, const double min[3], const double max[3],
                           int32_t material, uint32_t flags, const m3d_transform *xf,
                           int32_t *index_out);
int32_t m3d_scene_add_cylinder(m3d_scene_builder *b, const double p1[3], const double p2[3],
                               double radius, int32_t material, uint32_t flags,
                               const m3d_transform *xf, int32_t *index_out);
/* An instance of a device-resident mesh collider under a similarity transform: the scene keeps
 * ONE copy of the triangles and of the mesh's hierarchy however many instances refer to it
 * (render3d.Translate / MatrixMultiply of a shared collider, transform.go:6-85;
 * examples/renderings/golf_balls/main.go:25-39).  `mesh` must come from m3d_mesh_create on the
 * builder's context and outlive the scene.  Rays are taken to object space at the instance's
 * bounds and walk the mesh's own BVH there; prim is the triangle id inside the mesh. */
int32_t m3d_scene_add_instance(m3d_scene_builder *b, m3d_mesh *mesh, int32_t material, uint32_t flags,
                               const m3d_transform *xf, int32_t *index_out);
/* Scenes with more than a handful of analytic shapes / instances get an object-level hierarchy
 * over their bounds (render3d.BVHToObject, object.go:172-185), walked per ray instead of testing
 * every shape: thousands of spheres cost O(log n) per ray. */
int32_t m3d_scene_build(m3d_scene_builder *b, uint32_t build_flags, m3d_scene **out);
void m3d_scene_destroy(m3d_scene *scene);
int32_t m3d_scene_bounds(const m3d_scene *scene, double min_out[3], double max_out[3]);
/* The scene's own (merged, world-space) triangle hierarchy: num_triangles counts the triangles the
 * scene stores itself -- instanced meshes are not among them, they stay in their m3d_mesh. */
int32_t m3d_scene_get_info(const m3d_scene *scene, m3d_mesh_info *info);

/* Batched Object.Cast (render3d/object.go:141-153): like
 * m3d_mesh_first_ray_collisions plus obj (object index, -1 miss); prim is the
 * triangle id inside that object's mesh (or 0 for analytic shapes). */
int32_t m3d_scene_cast(m3d_scene *scene, const float *org, const float *dir, int64_t n,
                       float *t, int32_t *obj, int32_t *prim, float *normal,
                       uint32_t flags, m3d_stats *stats);

/* ---- renderers ------------------------------------------------------------- */

/* render3d.Camera (render3d/camera.go:19-41) */
typedef struct {
  double origin[3];
  double screen_x[3];
  double screen_y[3];
  double field_of_view;
} m3d_camera;

/* render3d.PointLight (render3d/light.go:57-66) */
typedef struct {
  double origin[3];
  double color[3];
  int32_t quad_dropoff;
  int32_t _pad;
} m3d_point_light;

/* Work partition for multi-GPU: this call renders rows [row_begin,row_end) of the
 * frame and samples [sample_begin, sample_begin+sample_count) of every pixel. */
#define M3D_PART_ATOMIC 1u /* several GPUs flush into ONE accumulator at the same time: the
                              per-pixel sums are added with system-scope red.add (the output
                              pointer may be another GPU's memory, mapped by peer access inside a
                              multi-device context or by m3d_ipc_open across processes) */
typedef struct {
  int32_t row_begin, row_end; /* 0,0 == whole frame */
  int64_t sample_begin;       /* first Philox sample index of this shard */
  uint32_t flags;             /* M3D_PART_* */
  uint32_t _pad;
} m3d_partition;

/* (*RayCaster).Render (render3d/raycast.go:15-39).
 * rgb: W*H*3 floats, linear RGB, row-major idx = x + y*W (render3d/image.go:33-47);
 * pixels whose ray misses are left untouched, exactly like the reference. */
int32_t m3d_render_raycast(m3d_scene *scene, const m3d_camera *cam,
                           const m3d_point_light *lights, int32_t num_lights,
                           int32_t width, int32_t height, const m3d_partition *part,
                           float *rgb, m3d_stats *stats);
/* Device-buffer variant: d_rgb is a device pointer to W*H*3 floats. */
int32_t m3d_render_raycast_device(m3d_scene *scene, const m3d_camera *cam,
                                  const m3d_point_light *lights, int32_t num_lights,
                                  int32_t width, int32_t height, const m3d_partition *part,
                                  void *d_rgb, void *stream, m3d_stats *stats);

/* Many RayCaster frames of one scene in ONE call: view v has its own camera and its own lights
 * (lights[light_begin[v] .. light_begin[v+1])); every frame starts black, is rendered at
 * width x height and, with downsample > 1, box-filtered on the device like Image.Downsample
 * (render3d/image.go:100-120); rgb receives num_views frames of (width/downsample) x
 * (height/downsample) x 3 floats, view-major, in one device-to-host copy.  What
 * render3d.SaveRandomGrid / SaveRotatingGIF do view by view (helpers.go:133-236): rows*cols or
 * `frames` renderings of one object at 2x supersampling.  The BVH is built once, the launch chains of
 * all views queue up behind each other without host round trips, and on a multi-device context the
 * views are spread over the GPUs. */
int32_t m3d_render_raycast_views(m3d_scene *scene, const m3d_camera *cams, int32_t num_views,
                                 const m3d_point_light *lights, const int32_t *light_begin, int32_t width,
                                 int32_t height, int32_t downsample, float *rgb, m3d_stats *stats);

typedef enum {
  M3D_FOCUS_PHONG = 0, /* render3d.PhongFocusPoint  (focus_point.go:30-71)  */
  M3D_FOCUS_SPHERE = 1 /* render3d.SphereFocusPoint (focus_point.go:73-153) */
} m3d_focus_kind;

#define M3D_MAX_FOCUS_POINTS 4
/* MaterialFilter closures cannot cross the ABI: the wrapper pre-evaluates the
 * filter once per material and passes the result as a bit mask
 * (bit i set == focus applies to material i; all-ones when the filter is nil). */
typedef struct {
  int32_t kind;
  int32_t _pad;
  double target[3]; /* Target / Center */
  double alpha;     /* PhongFocusPoint.Alpha */
  double radius;    /* SphereFocusPoint.Radius */
  uint64_t material_mask;
  double prob;      /* FocusPointProbs[i] */
} m3d_focus_point;

/* RecursiveRayTracer fields (render3d/raytrace.go:14-95). */
typedef struct {
  int32_t max_depth;
  int32_t num_samples;
  int32_t min_samples;           /* adaptive stop (ray_renderer.go:128-148): with max_stddev != 0
                                    every pixel stops at the reference's sample; the call must
                                    then cover all samples (no sample sharding, rows may be
                                    banded) and rgb_sum holds mean * num_samples */
  int32_t num_focus_points;
  double max_stddev;
  double oversaturated_stddevs;
  double cutoff;
  double antialias;
  double epsilon;                /* 0 -> DefaultEpsilon (raytrace.go:10) */
  m3d_focus_point focus[M3D_MAX_FOCUS_POINTS];
  uint64_t seed;                 /* Philox key */
} m3d_path_params;

/* (*RecursiveRayTracer).Render (render3d/raytrace.go:98-100, ray_renderer.go:25-56).
 * Outputs are per-pixel SUMS over the samples of this partition so that shards
 * add up: rgb_sum (W*H*3), rgb_sumsq (W*H*3 or NULL).  The caller divides by the
 * total sample count (ray_renderer.go:150) after reducing across GPUs. */
int32_t m3d_render_path(m3d_scene *scene, const m3d_camera *cam,
                        const m3d_point_light *lights, int32_t num_lights,
                        const m3d_path_params *params, int32_t width, int32_t height,
                        const m3d_partition *part, int32_t sample_count,
                        float *rgb_sum, float *rgb_sumsq, m3d_stats *stats);
int32_t m3d_render_path_device(m3d_scene *scene, const m3d_camera *cam,
                               const m3d_point_light *lights, int32_t num_lights,
                               const m3d_path_params *params, int32_t width, int32_t height,
                               const m3d_partition *part, int32_t sample_count,
                               void *d_rgb_sum, void *d_rgb_sumsq, void *stream,
                               m3d_stats *stats);

/* Area lights for BidirPathTracer.Light (render3d/light.go:104-314): the objects
 * of the scene that are also sampled as emitters (mesh or sphere objects). */
typedef struct {
  int32_t object;      /* scene object index */
  int32_t _pad;
  double emission[3];
} m3d_area_light;

/* BidirPathTracer fields (render3d/bidir.go:14-63). */
typedef struct {
  int32_t max_depth;
  int32_t max_light_depth;
  int32_t min_depth;
  int32_t num_samples;
  double roulette_delta;
  double power_heuristic;
  double cutoff;
  double antialias;
  double epsilon;
  uint64_t seed;
  /* adaptive stop, as in m3d_path_params (bidir.go:45-52) */
  int32_t min_samples;
  int32_t _pad;
  double max_stddev;
  double oversaturated_stddevs;
} m3d_bidir_params;

/* (*BidirPathTracer).Render (render3d/bidir.go:66-68,101-159). Sums, as above. */
int32_t m3d_render_bidir(m3d_scene *scene, const m3d_camera *cam,
                         const m3d_area_light *lights, int32_t num_lights,
                         const m3d_bidir_params *params, int32_t width, int32_t height,
                         const m3d_partition *part, int32_t sample_count,
                         float *rgb_sum, float *rgb_sumsq, m3d_stats *stats);
int32_t m3d_render_bidir_device(m3d_scene *scene, const m3d_camera *cam,
                                const m3d_area_light *lights, int32_t num_lights,
                                const m3d_bidir_params *params, int32_t width, int32_t height,
                                const m3d_partition *part, int32_t sample_count,
                                void *d_rgb_sum, void *d_rgb_sumsq, void *stream,
                                m3d_stats *stats);

/* colorSum/numSamples + optional 8-bit sRGB (ray_renderer.go:150, image.go:125-145,
 * light.go:41-47).  d_sum: W*H*3 floats; writes mean into d_mean (may alias d_sum)
 * and, if d_srgb8 != NULL, W*H*3 bytes. */
int32_t m3d_finalize_image_device(m3d_ctx *ctx, const void *d_sum, int64_t num_pixels,
                                  double inv_samples, void *d_mean, void *d_srgb8,
                                  void *stream);

/* ---- memory the caller shares with the library --------------------------------
 * Pinned host memory: host-buffer calls reach the PCIe rate only from page-locked buffers
 * (pageable memory is staged by the driver: about half the rate, and copies stop overlapping).
 * A Go / C caller allocates its ray and hit arrays here (the cgo binding wraps the pointer in a
 * slice with unsafe.Slice) or registers arrays it already owns. */
int32_t m3d_host_alloc(int64_t bytes, void **out);
int32_t m3d_host_free(void *ptr);
int32_t m3d_host_register(void *ptr, int64_t bytes);
int32_t m3d_host_unregister(void *ptr);

/* Device buffers owned by the library (zero-filled) and their export to the other processes of a
 * one-process-per-GPU job: rank 0 allocates the frame accumulator and exports it, every other
 * rank opens it and passes the mapped pointer as d_rgb_sum with M3D_PART_ATOMIC, so that its
 * flush kernel reduces into rank 0's memory over NVLink.  handle: M3D_IPC_HANDLE_BYTES bytes. */
#define M3D_IPC_HANDLE_BYTES 64
int32_t m3d_device_alloc(m3d_ctx *ctx, int64_t bytes, void **d_ptr);
int32_t m3d_device_free(m3d_ctx *ctx, void *d_ptr);
int32_t m3d_ipc_export(m3d_ctx *ctx, void *d_ptr, uint8_t *handle);
int32_t m3d_ipc_open(m3d_ctx *ctx, const uint8_t *handle, void **d_ptr);
int32_t m3d_ipc_close(m3d_ctx *ctx, void *d_ptr);

/* ---- diagnostics ---------------------------------------------------------------
 * On-chip bandwidth microbenchmarks for the roofline of the traversal kernel (the BVH
 * is L2 resident, node / triangle fetches are bounded by L2 -> SM bandwidth, not HBM).
 *   mode 0: every SM streams a working set of working_set_bytes (choose it < L2, > the L1s) with
 *           coalesced 16-byte loads that bypass L1;
 *   mode 1: every lane reads whole pseudo-random 80-byte records (the wide-node access pattern:
 *           32 different nodes per warp instruction).
 * Best of `repeats` timed launches, GB/s.  Not on any product path. */
int32_t m3d_measure_l2_bandwidth(m3d_ctx *ctx, int32_t mode, int64_t working_set_bytes,
                                 int32_t repeats, double *gb_per_s);

#ifdef __cplusplus
}
#endif
#endif /* M3D_H_ */
