// Wavefront path tracer: device-side parameter blocks and launch wrappers.
// Replaces rayRenderer.Render / estimateColor (render3d/ray_renderer.go:25-151) and
// RecursiveRayTracer.recurse (render3d/raytrace.go:138-229), which the reference runs as one
// goroutine-scheduled recursion per pixel (render3d/concurrency.go:17-43).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "scene.h"

namespace m3d {

struct DeviceFocus {  // render3d/focus_point.go:30-44, 73-85
  int32_t kind;       // M3D_FOCUS_PHONG / M3D_FOCUS_SPHERE
  float target[3];
  float alpha, radius, prob;
  uint64_t mask;      // bit i: MaterialFilter(material i) (pre-evaluated by the wrapper)
};

struct DevicePathParams {
  int32_t max_depth;
  int32_t num_focus;
  int32_t num_lights;
  float cutoff, antialias;
  // RecursiveRayTracer.Epsilon (raytrace.go:217-229: bounce / shadow origins move eps along the
  // ray).  The default 1e-8 is below float32 resolution and is replaced by the exact surface skip
  // ids; a larger user value is honoured ON TOP of them as the rays' tmin (in units of |dir|), so
  // that coplanar / duplicated surfaces within eps of the start point are stepped over like in
  // the reference.  0: skip ids only.
  float eps;
  uint64_t seed;
  DeviceFocus focus[4];
};

// One batch = nP consecutive pixels (frame-linear index pix0 .. pix0+nP) x S samples
// (absolute sample indices sample0 .. sample0+S).  Path slot = s * nP + p.
struct PathBatch {
  int32_t W;
  int32_t pix0, nP, S;
  uint32_t sample0;
  // optional explicit pixel list (frame-linear indices, device memory) instead of the range
  // pix0 .. pix0+nP: the still-unconverged pixels of an adaptive render
  const int32_t *pixels = nullptr;
};
#if defined(__CUDACC__)
__device__ __forceinline__ int batch_pixel(const PathBatch &b, int p) { return b.pixels ? b.pixels[p] : b.pix0 + p; }
#endif

// Structure-of-arrays path state, capacity `cap` slots.  Rays / raw hits / skip ids / queue
// entries are stored in QUEUE order (compacted, ping-pong [2]); throughput and accumulated
// colour are stored by SLOT and stay in place for the whole batch.
struct PathBuffers {
  int64_t cap;
  float4 *org[2], *dir[2];
  int32_t *skip[2], *queue[2];
  float4 *raw;
  float4 *thr, *accum;
  // hit records of the current bounce, in queue order: (point, surface id), (normal, material index),
  // (direction to the viewer, slot); and the per-material-kind work lists of queue positions
  float4 *hrA, *hrB, *hrC;
  int32_t *klist[4];
  // shadow rays towards point lights (num_lights per queue entry, fixed positions)
  float4 *sorg, *sdir, *sraw, *spay;
  int32_t *sskip;
  int *counts;                    // [0],[1]: queue lengths (ping-pong); [2]: shadow rays;
                                  // [4..7]: work-list lengths per material kind
  unsigned long long *ray_total;  // rays cast (statistics)
};

void launch_path_raygen(const DeviceCamera &cam, const DevicePathParams &pp, const PathBatch &b,
                        const PathBuffers &buf, cudaStream_t stream);
// bounce stage 1: resolve hits of queue `cur`, emission / ambient, shadow rays, hit records and
// per-material-kind work lists
void launch_path_resolve(const DeviceScene &sc, const DevicePathParams &pp, const DevicePointLight *lights,
                         const PathBatch &b, const PathBuffers &buf, int cur, int depth, cudaStream_t stream);
// bounce stage 2 for one material kind (m3d_material_kind): sample, weight, compact into queue 1-cur
void launch_path_sample(int kind, const DeviceScene &sc, const DevicePathParams &pp, const PathBatch &b,
                        const PathBuffers &buf, int cur, int depth, cudaStream_t stream);
void launch_path_shadow_resolve(const DeviceScene &sc, const DevicePathParams &pp, const PathBuffers &buf,
                                int cur, cudaStream_t stream);
// per pixel: add the S per-sample colours (and their squares) into the frame accumulators
void launch_path_flush(const PathBatch &b, const float4 *accum, float *rgb_sum, float *rgb_sumsq,
                       cudaStream_t stream);
// Shared-accumulator mode (M3D_PART_ATOMIC): several GPUs flush into one frame accumulator.
// A pixel range that takes several batches keeps its partial sums in `carry` (3 floats per batch
// position, local memory); the last batch adds carry + its own sums to the shared accumulator with
// system-scope red.add (over NVLink when the accumulator lives on another GPU).
void launch_path_flush_carry(const PathBatch &b, const float4 *accum, float *carry, float *carry_sq,
                             cudaStream_t stream);
void launch_path_flush_red(const PathBatch &b, const float4 *accum, float *rgb_sum, float *rgb_sumsq,
                           float *carry, float *carry_sq, cudaStream_t stream);

// The flush sequence of the non-adaptive batch loops (api_path.cu, api_bidir.cu): chooses between
// the three kernels above.  `atomic`: M3D_PART_ATOMIC; carry buffers hold 3 * nP floats (zeroed by
// the caller at the start of every pixel range that takes more than one batch) or are null.
struct FlushPlan {
  bool atomic = false;
  float *carry = nullptr, *carry_sq = nullptr;
};
inline void flush_batch(const FlushPlan &f, const PathBatch &b, const float4 *accum, float *rgb_sum, float *rgb_sumsq,
                        bool first_of_range, bool last_of_range, cudaStream_t stream) {
  if (!f.atomic) {
    launch_path_flush(b, accum, rgb_sum, rgb_sumsq, stream);
  } else if (last_of_range) {
    launch_path_flush_red(b, accum, rgb_sum, rgb_sumsq, first_of_range ? nullptr : f.carry,
                          first_of_range || !rgb_sumsq ? nullptr : f.carry_sq, stream);
  } else {
    launch_path_flush_carry(b, accum, f.carry, rgb_sumsq ? f.carry_sq : nullptr, stream);
  }
}

}  // namespace m3d
