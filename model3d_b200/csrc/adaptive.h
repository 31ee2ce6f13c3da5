// Adaptive per-pixel sampling: rayRenderer.estimateColor's early stop
// (render3d/ray_renderer.go:112-173) for the wavefront renderers.
//
// The reference samples one pixel at a time and tests convergence after every sample.  Here
// all still-active pixels advance in lock step through rounds of S samples; the flush kernel
// of a round walks a pixel's S new sample colours IN ORDER and applies the reference's test
// after each one (same statistics, same count-1 quirk), so the stopping sample of a pixel is
// exactly the reference's for the same sample sequence; samples of the round after the stop
// are discarded.  Round sizes grow geometrically, bounding the discarded work.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>

#include "api_common.h"
#include "path.h"

namespace m3d {

struct AdaptiveParams {
  int32_t num_samples, min_samples;
  double max_stddev, oversaturated_stddevs;
};

// Per-pixel state, indexed by (frame pixel index - pix_begin).
struct AdaptiveState {
  int32_t pix_begin;
  double *csum;      // 3 per pixel: colorSum
  double *csq;       // 3 per pixel: colorSqSum
  int32_t *divisor;  // 0 while active; else the reference's final numSamples (loop index)
};

void launch_adaptive_flush(const PathBatch &b, const float4 *accum, const AdaptiveState &st,
                           const AdaptiveParams &ap, cudaStream_t stream);
// pixels_in == nullptr: the range pix0 .. pix0+n.  Appends the still-active pixels to pixels_out.
void launch_adaptive_compact(const int32_t *pixels_in, int32_t pix0, int32_t n, const AdaptiveState &st,
                             int32_t *pixels_out, int *count_out, cudaStream_t stream);
// rgb_sum[pix] += mean * num_samples (so that the caller's division by NumSamples yields the
// reference's colorSum/numSamples); samples_total += sum of divisors... (statistics)
void launch_adaptive_finalize(int32_t npix, const AdaptiveState &st, const AdaptiveParams &ap, float *rgb_sum,
                              unsigned long long *samples_total, cudaStream_t stream);

// Runs the rounds.  run_batch traces one batch and leaves one colour per slot in `accum`.
// Scratch for state / pixel lists comes from ctx->scratch[8].  *samples_out: samples taken.
int32_t run_adaptive(m3d_ctx *ctx, cudaStream_t s, int32_t width, int32_t pix_begin, int32_t npix, int64_t cap,
                     const AdaptiveParams &ap, const float4 *accum, float *d_rgb_sum,
                     const std::function<int32_t(const PathBatch &)> &run_batch, int64_t *samples_out);

}  // namespace m3d
