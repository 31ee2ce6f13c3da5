// Launch wrappers for the sm_100a kernels (definitions in *.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace m3d {

struct DeviceBVH {
  const uint4 *nodes = nullptr;    // M3D_NODE_QUADS x uint4 per WideNode (node_layout.h)
  const float4 *tris = nullptr;    // 3 x float4 per TriRecord
  const float4 *vnormals = nullptr; // optional, 3 x float4 per triangle (leaf order)
  int64_t num_nodes = 0, num_tris = 0;
  float bmin[3] = {0, 0, 0}, bmax[3] = {0, 0, 0};  // bounds of all vertices
};

// counters: device pointer to 2 x uint64 (nodes, tris) or nullptr
struct TraceLaunch {
  const float4 *org_tmin;
  const float4 *dir_tmax;
  int64_t n;
  float4 *hit0;  // (t, b1, b2, bits(prim))
  float4 *hit1;  // (nx, ny, nz, bits(object))
  bool refine;   // float64 re-evaluation of the winning hit
  unsigned long long *counters;     // optional device 2 x u64 (nodes, tris)
  unsigned long long *ray_counter;  // device u64 work counter of the persistent kernel
  const int32_t *skip_tris = nullptr;  // optional per ray: leaf-order triangle index to ignore
  const int *n_ptr = nullptr;          // optional device count (<= n): wavefront queues whose
                                       // length is only known on the device
  uint32_t one_bits = 0x3f800000u;     // bits of 1.0f, kept opaque to ptxas (see unit_plus_byte)
  // Optional scratch of the bounds cull (trace_cull_wanted): n + 1 ints of device memory.  When given
  // (plain mesh batches only: no skip ids, no device-side count), rays that miss the bounds of all
  // vertices are retired by a streaming pre-pass and the traversal walks the list of survivors.
  int *cull_scratch = nullptr;
  const int *ray_list = nullptr;       // set by launch_trace_bvh_only itself
};
// Whether a plain mesh batch of n rays should be given cull_scratch
bool trace_cull_wanted(int64_t n, int64_t num_nodes);

// BVH traversal only: hit0 = (t_f32, 0, 0, bits(leaf-order triangle index | -1))
void launch_trace_bvh_only(const DeviceBVH &bvh, const TraceLaunch &p, cudaStream_t stream);
// traversal + finish pass (float64 refine) for a plain mesh
void launch_trace_first_hit(const DeviceBVH &bvh, const TraceLaunch &p, cudaStream_t stream);

// n*3 float arrays -> float4 SoA with constant tmin/tmax
// (shared_origin: optional HOST pointer to 3 floats, the origin of every ray; org3 is then ignored)
void launch_pack_rays(const float *org3, const float *dir3, int64_t n, float tmin, float tmax,
                      float4 *org_tmin, float4 *dir_tmax, cudaStream_t stream,
                      const float *shared_origin = nullptr);
// hit SoA -> the separate arrays of the host ABI (any output may be null)
void launch_unpack_hits(const float4 *hit0, const float4 *hit1, int64_t n, float *t, int32_t *prim,
                        int32_t *obj, float *normal3, float *bary3, cudaStream_t stream);

// all-hits counts per ray (counts) and / or their parity (inside); dir3 == nullptr: the fixed
// ColliderContains direction (collisions.go:119-134)
void launch_count_hits(const DeviceBVH &bvh, const float *org3, const float *dir3, int64_t n, int32_t *counts,
                       uint8_t *inside, cudaStream_t stream);

// all hits delivered: hit k of ray i goes to index offsets[i] + k of t / prim (caller triangle id) /
// normal3 / bary3 (the last two optional), ordered by t; offsets = exclusive prefix sum of the counts
void launch_collect_hits(const DeviceBVH &bvh, const float *org3, const float *dir3, int64_t n,
                         const int64_t *offsets, float *t, int32_t *prim, float *normal3, float *bary3,
                         cudaStream_t stream);

// nearest-triangle queries (sdf_kernels.cu): meshSDF.FaceSDF per point (any output may be null) ...
// (counters: optional device 3 x u64: nodes fetched, float32 screens, float64 evaluations)
void launch_mesh_sdf(const DeviceBVH &bvh, const float *pts3, int64_t n, float *sdf, float *closest3,
                     int32_t *face, float *normal3, unsigned long long *counters, cudaStream_t stream);
// ... Collider.SphereCollision per (centre, radius); radii == nullptr: one radius for all ...
void launch_sphere_collisions(const DeviceBVH &bvh, const float *centers3, const float *radii, float radius,
                              int64_t n, uint8_t *out, cudaStream_t stream);
// ... and ColliderContains' margin rule from the crossing parity and the sphere test
void launch_contains_margin(const uint8_t *parity, const uint8_t *near, int64_t n, bool margin_negative,
                            uint8_t *inside, cudaStream_t stream);
// deepest wide-BVH the nearest-triangle traversal stack covers (7 entries per level + 1)
int sdf_stack_capacity();

int device_sm_count();

}  // namespace m3d
