// C ABI: (*RecursiveRayTracer).Render (include/m3d.h, m3d_render_path*).
// Replaces render3d/raytrace.go:98-229 + render3d/ray_renderer.go:25-151 with a wavefront
// pipeline: raygen -> [trace -> shade/compact (-> shadow trace -> shadow resolve)] x depth
// -> flush, batch by batch, all enqueued on one stream with device-side queue lengths
// (no host round trip between bounces).
#include <chrono>
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <cstring>

#include "adaptive.h"
#include "api_common.h"
#include "path.h"
#include "scene.h"
#include "scene_host.h"

using namespace m3d;

namespace m3d {
const DeviceScene &scene_device(const m3d_scene *s);
m3d_ctx *scene_ctx(const m3d_scene *s);
DeviceCamera device_camera(const m3d_camera &c, int W, int H);
const std::vector<m3d_material_desc> &scene_materials(const m3d_scene *s);
}  // namespace m3d

namespace {

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// Bytes carve_path_buffers reserves for `cap` slots (every array padded to 256 bytes)
size_t path_buffer_bytes(int64_t cap, int num_lights) {
  const size_t c = (size_t)cap, nl = (size_t)num_lights;
  size_t total = 0;
  for (int i = 0; i < 2; i++) total += 2 * align256(c * 16) + 2 * align256(c * 4);  // org, dir, skip, queue
  total += 6 * align256(c * 16) + 4 * align256(c * 4);  // raw, thr, accum, hit records; work lists
  total += 4 * align256(c * nl * 16) + align256(c * nl * 4) + align256(64);  // shadow rays; counters
  return total;
}

// Carves the path-state arrays out of one scratch allocation.
int32_t carve_path_buffers(m3d_ctx *ctx, int64_t cap, int num_lights, PathBuffers &b) {
  const size_t c = (size_t)cap, nl = (size_t)num_lights;
  size_t total = 0;
  auto take = [&](size_t bytes) {
    const size_t off = total;
    total += align256(bytes);
    return off;
  };
  size_t o_org[2], o_dir[2], o_skip[2], o_queue[2];
  for (int i = 0; i < 2; i++) {
    o_org[i] = take(c * 16);
    o_dir[i] = take(c * 16);
    o_skip[i] = take(c * 4);
    o_queue[i] = take(c * 4);
  }
  const size_t o_raw = take(c * 16), o_thr = take(c * 16), o_acc = take(c * 16);
  const size_t o_hra = take(c * 16), o_hrb = take(c * 16), o_hrc = take(c * 16);
  size_t o_kl[4];
  for (int i = 0; i < 4; i++) o_kl[i] = take(c * 4);
  const size_t o_sorg = take(c * nl * 16), o_sdir = take(c * nl * 16), o_sraw = take(c * nl * 16),
               o_spay = take(c * nl * 16), o_sskip = take(c * nl * 4);
  const size_t o_counts = take(64);
  if (total != path_buffer_bytes(cap, num_lights)) return fail(M3D_ERR_CUDA, "path buffer size bookkeeping is out of step");
  M3D_CUDA(ctx->scratch[4].reserve(total));
  char *p = ctx->scratch[4].as<char>();
  b.cap = cap;
  for (int i = 0; i < 2; i++) {
    b.org[i] = (float4 *)(p + o_org[i]);
    b.dir[i] = (float4 *)(p + o_dir[i]);
    b.skip[i] = (int32_t *)(p + o_skip[i]);
    b.queue[i] = (int32_t *)(p + o_queue[i]);
  }
  b.raw = (float4 *)(p + o_raw);
  b.thr = (float4 *)(p + o_thr);
  b.accum = (float4 *)(p + o_acc);
  b.hrA = (float4 *)(p + o_hra);
  b.hrB = (float4 *)(p + o_hrb);
  b.hrC = (float4 *)(p + o_hrc);
  for (int i = 0; i < 4; i++) b.klist[i] = (int32_t *)(p + o_kl[i]);
  b.sorg = (float4 *)(p + o_sorg);
  b.sdir = (float4 *)(p + o_sdir);
  b.sraw = (float4 *)(p + o_sraw);
  b.spay = (float4 *)(p + o_spay);
  b.sskip = (int32_t *)(p + o_sskip);
  b.counts = (int *)(p + o_counts);
  b.ray_total = (unsigned long long *)(p + o_counts + 32);
  return M3D_OK;
}

}  // namespace

static int32_t render_path_one_device(m3d_scene *scene, const m3d_camera *cam, const m3d_point_light *lights,
                                      int32_t num_lights, const m3d_path_params *params, int32_t width,
                                      int32_t height, const m3d_partition *part, int32_t sample_count,
                                      void *d_rgb_sum, void *d_rgb_sumsq, void *stream, m3d_stats *stats);

namespace m3d {
// Splits a render over the devices of a multi-device scene (shared by the path tracer and the
// bidirectional tracer): fixed-spp renders by sample index -- every device takes a contiguous
// range of the absolute sample indices of every pixel, so the load is balanced whatever the
// scene looks like and the image does not depend on the device count -- adaptive renders
// (per-pixel early stop) by row band.  Every shard flushes into the one accumulator with
// M3D_PART_ATOMIC.
int32_t shard_render(m3d_scene *scene, bool adaptive, int32_t height, const m3d_partition *part,
                     int32_t sample_count, cudaStream_t stream, m3d_stats *stats,
                     const std::function<int32_t(m3d_scene *, const m3d_partition &, int32_t, cudaStream_t,
                                                 m3d_stats *)> &one) {
  const int g = 1 + (int)scene->replicas.size();
  int r0 = 0, r1 = height;
  int64_t s_begin = 0;
  if (part) {
    if (!(part->row_begin == 0 && part->row_end == 0)) {
      r0 = part->row_begin;
      r1 = part->row_end;
      if (r0 < 0 || r1 > height || r0 > r1)
        return fail(M3D_ERR_INVALID_ARG, "bad row partition [%d,%d) of %d rows", r0, r1, height);
    }
    s_begin = part->sample_begin;
  }
  return render_sharded(scene, stream, stats, [&](int i, m3d_scene *si, cudaStream_t s, m3d_stats *st) -> int32_t {
    m3d_partition pi{};
    pi.flags = (part ? part->flags : 0u) | M3D_PART_ATOMIC;
    int64_t b, e;
    int32_t count = sample_count;
    if (adaptive) {
      split_range(r1 - r0, g, i, &b, &e);
      if (b == e) return M3D_OK;
      pi.row_begin = r0 + (int32_t)b;
      pi.row_end = r0 + (int32_t)e;
      pi.sample_begin = s_begin;
    } else {
      split_range(sample_count, g, i, &b, &e);
      if (b == e) return M3D_OK;
      pi.row_begin = r0;
      pi.row_end = r1;
      if (r0 == 0 && r1 == 0) return M3D_OK;  // empty frame
      pi.sample_begin = s_begin + b;
      count = (int32_t)(e - b);
    }
    return one(si, pi, count, s, st);
  });
}
}  // namespace m3d

extern "C" {

int32_t m3d_render_path_device(m3d_scene *scene, const m3d_camera *cam, const m3d_point_light *lights,
                               int32_t num_lights, const m3d_path_params *params, int32_t width,
                               int32_t height, const m3d_partition *part, int32_t sample_count,
                               void *d_rgb_sum, void *d_rgb_sumsq, void *stream, m3d_stats *stats) {
  if (!scene || !cam || !params || width <= 0 || height <= 0 || !d_rgb_sum || num_lights < 0 ||
      (num_lights > 0 && !lights) || sample_count < 0)
    return fail(M3D_ERR_INVALID_ARG, "m3d_render_path: bad arguments");
  M3D_LOCK(scene->ctx);
  if (!scene->replicas.empty() && sample_count > 0) {
    const bool adaptive = params->min_samples != 0 && params->max_stddev != 0;
    return shard_render(scene, adaptive, height, part, sample_count, (cudaStream_t)stream, stats,
                        [&](m3d_scene *si, const m3d_partition &pi, int32_t count, cudaStream_t s, m3d_stats *st) {
                          return render_path_one_device(si, cam, lights, num_lights, params, width, height, &pi,
                                                        count, d_rgb_sum, d_rgb_sumsq, s, st);
                        });
  }
  return render_path_one_device(scene, cam, lights, num_lights, params, width, height, part, sample_count,
                                d_rgb_sum, d_rgb_sumsq, stream, stats);
}

}  // extern "C"

static int32_t render_path_one_device(m3d_scene *scene, const m3d_camera *cam, const m3d_point_light *lights,
                                      int32_t num_lights, const m3d_path_params *params, int32_t width,
                                      int32_t height, const m3d_partition *part, int32_t sample_count,
                                      void *d_rgb_sum, void *d_rgb_sumsq, void *stream, m3d_stats *stats) {
  if (params->num_focus_points < 0 || params->num_focus_points > M3D_MAX_FOCUS_POINTS)
    return fail(M3D_ERR_UNSUPPORTED, "at most %d focus points are supported", M3D_MAX_FOCUS_POINTS);
  // rayRenderer.HasConvergenceCheck (ray_renderer.go:153-155) without the Convergence callback
  const bool adaptive = params->min_samples != 0 && params->max_stddev != 0;
  if (params->max_depth < 0 || params->max_depth > 1000) return fail(M3D_ERR_INVALID_ARG, "bad max_depth");
  if ((int64_t)width * height > (int64_t)0x7fffffff / 4) return fail(M3D_ERR_INVALID_ARG, "frame too large");
  m3d_ctx *ctx = scene_ctx(scene);
  const DeviceScene &sc = scene_device(scene);
  M3D_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  int row_begin = 0, row_end = height;
  int64_t sample_begin = 0;
  if (part) {
    if (!(part->row_begin == 0 && part->row_end == 0)) {
      row_begin = part->row_begin;
      row_end = part->row_end;
      if (row_begin < 0 || row_end > height || row_begin > row_end)
        return fail(M3D_ERR_INVALID_ARG, "bad row partition [%d,%d) of %d rows", row_begin, row_end, height);
    }
    sample_begin = part->sample_begin;
    if (sample_begin < 0 || sample_begin + sample_count > (int64_t)0xffffffffll)
      return fail(M3D_ERR_INVALID_ARG, "sample range out of bounds");
  }
  if (stats) std::memset(stats, 0, sizeof(*stats));
  static const bool dbg = getenv("M3D_DEBUG_T") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_entry = now();
  double t_carved = 0, t_enq = 0;
  const int64_t npix = (int64_t)width * (row_end - row_begin);
  if (npix == 0 || sample_count == 0) return M3D_OK;
  if (adaptive && (sample_begin != 0 || sample_count != params->num_samples || d_rgb_sumsq))
    return fail(M3D_ERR_INVALID_ARG,
                "adaptive sampling (MinSamples/MaxStddev) stops per pixel: it cannot be sharded by sample index "
                "(sample_begin must be 0 and sample_count == num_samples; shard by rows) and has no sumsq output");

  DevicePathParams pp;
  std::memset(&pp, 0, sizeof(pp));
  pp.max_depth = params->max_depth;
  pp.num_focus = params->num_focus_points;
  pp.num_lights = num_lights;
  pp.cutoff = (float)params->cutoff;
  pp.antialias = (float)params->antialias;
  pp.eps = params->epsilon > 1e-7 ? (float)params->epsilon : 0.f;  // DefaultEpsilon 1e-8 -> surface skip ids
  pp.seed = params->seed;
  for (int i = 0; i < pp.num_focus; i++) {
    const m3d_focus_point &f = params->focus[i];
    if (f.kind != M3D_FOCUS_PHONG && f.kind != M3D_FOCUS_SPHERE)
      return fail(M3D_ERR_UNSUPPORTED, "focus point kind %d is not supported on the GPU path", f.kind);
    pp.focus[i].kind = f.kind;
    for (int k = 0; k < 3; k++) pp.focus[i].target[k] = (float)f.target[k];
    pp.focus[i].alpha = (float)f.alpha;
    pp.focus[i].radius = (float)f.radius;
    pp.focus[i].prob = (float)f.prob;
    pp.focus[i].mask = f.material_mask;
  }
  if (pp.cutoff > 1.f) return M3D_OK;  // recurse() returns black at depth 0 (raytrace.go:139-142)

  // batch geometry: nP pixels x S samples <= cap slots
  static const int batch_log2 = [] {  // M3D_PATH_BATCH_LOG2: path slots per batch (tuning runs)
    const char *e = getenv("M3D_PATH_BATCH_LOG2");
    const int v = e ? atoi(e) : 26;
    return (v < 10 || v > 28) ? 26 : v;
  }();
  // <= 48 GB of the 180 GB for the path state (192 B per slot + 68 B per point light), and queue
  // positions / shadow-ray counts stay below 2^31
  const int64_t per_slot = 192 + 68 * (int64_t)num_lights;
  // ... and at most half of what is free on the device right now (beyond what this context already
  // holds).  cudaMemGetInfo is only asked when the scratch allocation would have to grow: it takes
  // 10-40 ms when several devices / processes share peer mappings (measured: 2 x B200, r2).
  const int64_t total = npix * sample_count;
  const int64_t want_slots = std::min<int64_t>(
      total, std::min<int64_t>((int64_t)1 << batch_log2, ((int64_t)1 << 30) / std::max<int64_t>(1, num_lights)));
  int64_t budget = (int64_t)48 << 30;
  // (the bound has to be the carve's own byte count: an estimate above it asked on every call, and
  // with CUDA IPC / peer mappings in the process the query costs 4-70 ms -- measured at 2 GPUs under
  // torchrun: 78 ms of kernels per C3 frame, 79-161 ms between the events around the call)
  if (ctx->scratch[4].bytes < path_buffer_bytes(want_slots, num_lights)) {
    size_t free_b = 0, total_b = 0;
    M3D_CUDA(cudaMemGetInfo(&free_b, &total_b));
    budget = std::min<int64_t>(budget, (int64_t)((free_b + ctx->scratch[4].bytes) / 2));
  }
  const int64_t kMaxSlots = std::min<int64_t>(std::min<int64_t>((int64_t)1 << batch_log2, budget / per_slot),
                                              ((int64_t)1 << 30) / std::max<int64_t>(1, num_lights));
  const int64_t cap = std::min(total, std::max<int64_t>(kMaxSlots, 1));
  const int64_t nP_max = std::min(npix, cap);
  PathBuffers buf;
  if (int32_t rc = carve_path_buffers(ctx, cap, num_lights, buf)) return rc;
  DevicePointLight *d_lights = nullptr;
  std::vector<DevicePointLight> hl(num_lights);
  if (num_lights) {
    M3D_CUDA(ctx->scratch[5].reserve(hl.size() * sizeof(DevicePointLight)));
    d_lights = ctx->scratch[5].as<DevicePointLight>();
    for (int i = 0; i < num_lights; i++) {
      for (int k = 0; k < 3; k++) {
        hl[i].origin[k] = (float)lights[i].origin[k];
        hl[i].color[k] = (float)lights[i].color[k];
      }
      hl[i].quad_dropoff = lights[i].quad_dropoff;
    }
    M3D_CUDA(cudaMemcpyAsync(d_lights, hl.data(), hl.size() * sizeof(DevicePointLight), cudaMemcpyHostToDevice, s));
  }
  M3D_CUDA(cudaMemsetAsync(buf.ray_total, 0, sizeof(unsigned long long), s));
  t_carved = now();
  const DeviceCamera dc = device_camera(*cam, width, height);

  // material kinds that occur in the scene: one sampling kernel per kind and bounce
  // (kinds of the objects' own materials; parts of a JoinedMaterial are sampled by the joined kernel)
  unsigned kinds_present = 0;
  for (int32_t mi : scene->object_material) kinds_present |= 1u << (unsigned)scene->host_materials[(size_t)mi].kind;
  GpuTimer tm;
  tm.start(s);
  StageTimer stages(s);  // M3D_STAGE_TIMING=1: device time per kernel type, printed to stderr
  static const char *const kStageNames[] = {"raygen", "trace", "resolve", "sample", "shadow", "flush"};
  int64_t launches = 0;
  // traces one batch: raygen -> [trace -> shade (-> shadow trace -> resolve)] x depth; leaves one
  // colour per slot in buf.accum
  auto run_batch = [&](const PathBatch &b) -> int32_t {
    const int64_t n = (int64_t)b.nP * b.S;
    stages.mark(0);
    launch_path_raygen(dc, pp, b, buf, s);
    launches++;
    int cur = 0;
    for (int depth = 0; depth <= pp.max_depth; depth++) {
      TraceLaunch t;
      t.org_tmin = buf.org[cur];
      t.dir_tmax = buf.dir[cur];
      t.n = n;
      t.n_ptr = buf.counts + cur;
      t.hit0 = buf.raw;
      t.hit1 = nullptr;
      t.refine = false;
      t.counters = nullptr;
      t.skip_tris = buf.skip[cur];
      t.ray_counter = next_work_counter(ctx);
      if (!t.ray_counter) return fail(M3D_ERR_OOM, "work counter allocation failed");
      stages.mark(1);
      launch_trace_bvh_only(sc.bvh, t, s);
      stages.mark(2);
      launch_path_resolve(sc, pp, d_lights, b, buf, cur, depth, s);
      stages.mark(3);
      launches += 2;
      if (depth < pp.max_depth)
        for (int k = 0; k < 4; k++)
          if (kinds_present & (1u << k)) {
            launch_path_sample(k, sc, pp, b, buf, cur, depth, s);
            launches++;
          }
      stages.mark(4);
      if (num_lights > 0) {
        TraceLaunch ts;
        ts.org_tmin = buf.sorg;
        ts.dir_tmax = buf.sdir;
        ts.n = n * num_lights;
        ts.n_ptr = buf.counts + 2;
        ts.hit0 = buf.sraw;
        ts.hit1 = nullptr;
        ts.refine = false;
        ts.counters = nullptr;
        ts.skip_tris = buf.sskip;
        ts.ray_counter = next_work_counter(ctx);
        if (!ts.ray_counter) return fail(M3D_ERR_OOM, "work counter allocation failed");
        launch_trace_bvh_only(sc.bvh, ts, s);
        launch_path_shadow_resolve(sc, pp, buf, cur, s);
        launches += 2;
      }
      // the consumed queue becomes the next output queue; the work lists start empty again
      M3D_CUDA(cudaMemsetAsync(buf.counts + cur, 0, sizeof(int), s));
      M3D_CUDA(cudaMemsetAsync(buf.counts + 4, 0, 4 * sizeof(int), s));
      cur ^= 1;
    }
    stages.mark(5);
    return M3D_OK;
  };
  int64_t samples_taken = npix * sample_count;
  if (adaptive) {
    AdaptiveParams ap;
    ap.num_samples = params->num_samples;
    ap.min_samples = params->min_samples;
    ap.max_stddev = params->max_stddev;
    ap.oversaturated_stddevs = params->oversaturated_stddevs;
    if (int32_t rc = run_adaptive(ctx, s, width, (int32_t)((int64_t)row_begin * width), (int32_t)npix, cap, ap,
                                  buf.accum, (float *)d_rgb_sum, run_batch, &samples_taken))
      return rc;
  } else {
    // M3D_PART_ATOMIC: other GPUs flush into the same accumulator at the same time.  A pixel range
    // that takes several batches keeps its partial sums in local memory and only its last batch
    // adds to the shared accumulator (with red.add, over NVLink when it lives on another GPU).
    FlushPlan fp;
    fp.atomic = part && (part->flags & M3D_PART_ATOMIC);
    if (fp.atomic && sample_count > std::max<int64_t>(1, cap / nP_max)) {
      M3D_CUDA(ctx->scratch[11].reserve((size_t)nP_max * 6 * sizeof(float)));
      fp.carry = ctx->scratch[11].as<float>();
      fp.carry_sq = fp.carry + (size_t)nP_max * 3;
    }
    for (int64_t p0 = 0; p0 < npix; p0 += nP_max) {
      const int64_t nP = std::min(nP_max, npix - p0);
      const int64_t S_max = std::max<int64_t>(1, cap / nP);
      if (fp.carry && sample_count > S_max)
        M3D_CUDA(cudaMemsetAsync(fp.carry, 0, (size_t)nP_max * 6 * sizeof(float), s));
      for (int64_t s0 = 0; s0 < sample_count; s0 += S_max) {
        PathBatch b;
        b.W = width;
        b.pix0 = (int32_t)((int64_t)row_begin * width + p0);
        b.nP = (int32_t)nP;
        b.S = (int32_t)std::min<int64_t>(S_max, sample_count - s0);
        b.sample0 = (uint32_t)(sample_begin + s0);
        if (int32_t rc = run_batch(b)) return rc;
        flush_batch(fp, b, buf.accum, (float *)d_rgb_sum, (float *)d_rgb_sumsq, s0 == 0,
                    s0 + S_max >= sample_count, s);
        launches++;
      }
    }
  }
  stages.mark(6);
  tm.stop(s);
  t_enq = now();
  // host-side tables (lights) must outlive the async copies; stats need the counters
  unsigned long long rays = 0;
  M3D_CUDA(cudaMemcpyAsync(&rays, buf.ray_total, sizeof(rays), cudaMemcpyDeviceToHost, s));
  M3D_CUDA(cudaStreamSynchronize(s));
  if (dbg)
    fprintf(stderr, "[m3d] render_path: setup %.3f ms, enqueue %.3f ms, wait %.3f ms (cap %lld)\n", t_carved - t_entry,
            t_enq - t_carved, now() - t_enq, (long long)cap);
  M3D_CUDA(cudaGetLastError());
  stages.report("m3d_render_path", kStageNames, 6);
  if (stats) {
    stats->rays = (int64_t)rays;
    stats->kernel_ms = tm.ms();
    stats->launches = launches;
    stats->samples = samples_taken;
  }
  return M3D_OK;
}

extern "C" {

int32_t m3d_render_path(m3d_scene *scene, const m3d_camera *cam, const m3d_point_light *lights,
                        int32_t num_lights, const m3d_path_params *params, int32_t width, int32_t height,
                        const m3d_partition *part, int32_t sample_count, float *rgb_sum, float *rgb_sumsq,
                        m3d_stats *stats) {
  if (!scene || !rgb_sum || width <= 0 || height <= 0)
    return fail(M3D_ERR_INVALID_ARG, "m3d_render_path: bad arguments");
  m3d_ctx *ctx = scene_ctx(scene);
  M3D_LOCK(ctx);
  M3D_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)width * height * 3 * sizeof(float);
  M3D_CUDA(ctx->scratch[3].reserve(bytes * 2));
  float *d_sum = ctx->scratch[3].as<float>();
  float *d_sq = rgb_sumsq ? d_sum + (size_t)width * height * 3 : nullptr;
  M3D_CUDA(cudaMemsetAsync(d_sum, 0, bytes * (rgb_sumsq ? 2 : 1), ctx->stream));
  int32_t rc = m3d_render_path_device(scene, cam, lights, num_lights, params, width, height, part, sample_count,
                                      d_sum, d_sq, ctx->stream, stats);
  if (rc != M3D_OK) return rc;
  M3D_CUDA(cudaMemcpyAsync(rgb_sum, d_sum, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  if (rgb_sumsq) M3D_CUDA(cudaMemcpyAsync(rgb_sumsq, d_sq, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  M3D_CUDA(cudaStreamSynchronize(ctx->stream));
  if (stats) stats->d2h_bytes = (int64_t)(bytes * (rgb_sumsq ? 2 : 1));
  return M3D_OK;
}

}  // extern "C"
