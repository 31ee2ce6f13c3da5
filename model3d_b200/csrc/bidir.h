// Bidirectional path tracer: device-side parameter blocks and launch wrappers.
// Replaces BidirPathTracer.rayColor and helpers (render3d/bidir.go:101-576) and the area
// lights it samples (render3d/light.go:104-314).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "path.h"
#include "scene.h"

namespace m3d {

constexpr int kBidirMaxDepth = 16;  // per sub-path; joined paths have at most 32 vertices

// 1: the connection kernel also walks every joined path like the reference's `densities` and reports
// weights that differ from the table-based ones (tuning / verification builds:
// scripts/build_variant.sh chk "-DM3D_CONNECT_CHECK=1" "bidir_kernels.cu api_bidir.cu")
#ifndef M3D_CONNECT_CHECK
#define M3D_CONNECT_CHECK 0
#endif

struct DeviceBidirParams {
  int32_t max_depth, max_light_depth, min_depth;
  float cutoff, antialias;
  float eps;  // BidirPathTracer.Epsilon above float32 resolution, else 0 (see DevicePathParams.eps)
  double roulette_delta, power_heuristic;
  uint64_t seed;
  int32_t num_lights;
  double total_light;  // AreaLight.TotalEmission of the joined light (light.go:313-314)
};

// One emitter of the joined area light (light.go:104-161, 227-274).
struct DeviceAreaLight {
  int32_t kind;       // 0 mesh, SHAPE_SPHERE
  int32_t object;
  int32_t surf;       // sphere: -2 - shape index (self-intersection guard id)
  int32_t tri_begin, tri_count;  // range in the light-triangle table (mesh lights)
  float emission[3];
  float center[3], radius;
  double cumu_total;  // cumulative TotalEmission up to and including this light
  double total_area;
};

// One triangle of a mesh light: vertices, flat normal, cumulative area inside its light,
// and its leaf-order index in the scene BVH (self-intersection guard id).
struct DeviceLightTri {
  float v[9];
  float n[3];
  double cumu_area;
  int32_t leaf_index;
  int32_t pad;
};

// SoA path-vertex storage: field f of vertex `depth` of slot s lives at
// verts[(f * D + depth) * cap + s]; 7 float4 fields per vertex (see bidir_kernels.cu).
// Sub-path vertices: seven float4 fields per vertex, field-major (consecutive slots of one field and depth
// are adjacent).  M3D_BIDIR_VERT_AOS = 1 (build option) keeps the fields of one vertex together in a
// 128-byte record instead: measured SLOWER on C5 (1024^2 x 128 spp: eye 127 -> 140 ms, light 106 -> 126,
// prefix 61 -> 73, connect 228 -> 248): most paths are alive at the shallow depths where the time goes, so
// the queue-ordered lanes of the shading kernels hold nearly consecutive slots and the field-major stores
// are coalesced, while the record layout strides every lane by 128 bytes.
#ifndef M3D_BIDIR_VERT_AOS
#define M3D_BIDIR_VERT_AOS 0
#endif
constexpr int kBidirVertexFields = M3D_BIDIR_VERT_AOS ? 8 : 7;

struct BidirBuffers {
  int64_t cap;
  int32_t De, Dl;  // vertex capacity of the eye / light sub-paths
  float4 *ev, *lv;
  int32_t *ne, *nl;
  float4 *org[2], *dir[2];
  int32_t *skip[2], *queue[2];
  float4 *raw;
  float4 *ender_full, *ender_roul;  // PathEnder state: (fullMask, currentRoulette), rouletteMask
  float4 *accum;
  // running products of allPathCombinations per sub-path prefix, 4 doubles each:
  // eyepre[(i-1)*cap + slot] = eyeDensity, eyeBSDF.xyz; lightpre[(j-1)*cap + slot] likewise
  double *eyepre, *lightpre;
  // M3D_CONNECT_CHECK builds only: compact per-vertex records of the path walk, depth-major: eye
  // vertices at [0, De), light vertices at [De, De+Dl)
  float4 *misA;
  double2 *misB;
  float *misC;
  // per-sample tables of the MIS sums (bidir_kernels.cu, "MIS weights in O(1) per connection"),
  // entry-major: mistab[entry * cap + slot], entries enumerated by MisTab
  double *mistab;
  // connection (visibility) rays of all (i, j) pairs of the batch, compacted
  float4 *corg, *cdir, *craw, *cpay;
  int32_t *cskip;
  // connection work items slot | i << 20 | j << 25 (cap <= 2^20, depths <= 16), grouped by
  // (i, j) class: class c = (i-1) * (max_light_depth+1) + j owns work[c*cap, c*cap + class_counts[c])
  uint32_t *work;
  int *class_counts;
  // chunked layout (bidir_kernels.cu, M3D_CONNECT_CHUNKED): the items of samples [256 b, 256 b + 256) are
  // work[256 b * classes, ... + chunk_counts[b]), ordered by class
  int *chunk_counts;
  int *counts;     // [0],[1] queue lengths, [2] connection rays
  unsigned long long *ray_total;
};

// Entry numbering of BidirBuffers::mistab for sub-path capacities De (eye) and Dl (light).
struct MisTab {
  int De, Dl;
  __host__ __device__ int tri(int x) const { return x * (x + 1) / 2; }
  __host__ __device__ int lt(int t) const { return t - 1; }                                  // t = 1..Dl
  __host__ __device__ int slp(int j) const { return Dl + j - 1; }                            // j = 1..Dl
  __host__ __device__ int h(int j, int t_lo) const { return 2 * Dl + tri(j - 3) + t_lo - 1; }  // j = 3..Dl, t_lo = 1..j-2
  __host__ __device__ int eye0() const { return 2 * Dl + tri(Dl > 2 ? Dl - 2 : 0); }
  __host__ __device__ int ge(int e) const { return eye0() + e; }                             // e = 1..De-1
  __host__ __device__ int rsp(int e) const { return eye0() + De + e; }                       // e = 0..De-2
  __host__ __device__ int k(int i, int m_lo) const { return eye0() + 2 * De + tri(i - 3) + m_lo; }  // i = 3..De+1, m_lo = 0..i-3
  __host__ __device__ int entries() const { return eye0() + 2 * De + tri(De > 1 ? De - 1 : 0); }
};

void launch_bidir_eye_raygen(const DeviceCamera &cam, const DeviceBidirParams &bp, const PathBatch &b,
                             const BidirBuffers &buf, cudaStream_t stream);
void launch_bidir_eye_shade(const DeviceScene &sc, const DeviceBidirParams &bp, const PathBatch &b,
                            const BidirBuffers &buf, int cur, int depth, cudaStream_t stream);
void launch_bidir_light_raygen(const DeviceScene &sc, const DeviceBidirParams &bp, const DeviceAreaLight *lights,
                               const DeviceLightTri *tris, const PathBatch &b, const BidirBuffers &buf,
                               cudaStream_t stream);
void launch_bidir_light_shade(const DeviceScene &sc, const DeviceBidirParams &bp, const PathBatch &b,
                              const BidirBuffers &buf, int cur, int depth, cudaStream_t stream);
// running products + compact MIS records of both sub-paths (one thread per sample)
void launch_bidir_prefix(const DeviceBidirParams &bp, const PathBatch &b, const BidirBuffers &buf,
                         cudaStream_t stream);
// every (eye prefix, light prefix) connection of the batch; emits compacted visibility rays
void launch_bidir_connect(const DeviceScene &sc, const DeviceBidirParams &bp, const PathBatch &b,
                          const BidirBuffers &buf, cudaStream_t stream);
void launch_bidir_connect_resolve(const DeviceScene &sc, const BidirBuffers &buf, cudaStream_t stream);

}  // namespace m3d
