// C ABI: (*BidirPathTracer).Render (include/m3d.h, m3d_render_bidir*).
// Replaces render3d/bidir.go:66-576 with a wavefront pipeline per batch of samples:
//   eye raygen -> [trace -> eye shade] x MaxDepth
//   light raygen -> [trace -> light shade] x (MaxLightDepth-1)
//   for every eye prefix length: connect (MIS in float64) -> trace visibility rays -> resolve
//   flush
// all on one stream with device-side queue lengths.
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <cstring>

#include "adaptive.h"
#include "api_common.h"
#include "bidir.h"
#include "scene_host.h"

using namespace m3d;

namespace m3d {
DeviceCamera device_camera(const m3d_camera &c, int W, int H);
}

namespace {

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

int32_t carve_bidir_buffers(m3d_ctx *ctx, int64_t cap, int De, int Dl, BidirBuffers &b) {
  const size_t c = (size_t)cap;
  size_t total = 0;
  auto take = [&](size_t bytes) {
    const size_t off = total;
    total += align256(bytes);
    return off;
  };
  const size_t o_ev = take(c * 16 * kBidirVertexFields * De), o_lv = take(c * 16 * kBidirVertexFields * Dl);
  const size_t o_ne = take(c * 4), o_nl = take(c * 4);
  size_t o_org[2], o_dir[2], o_skip[2], o_queue[2];
  for (int i = 0; i < 2; i++) {
    o_org[i] = take(c * 16);
    o_dir[i] = take(c * 16);
    o_skip[i] = take(c * 4);
    o_queue[i] = take(c * 4);
  }
  const size_t o_raw = take(c * 16), o_ef = take(c * 16), o_er = take(c * 16), o_acc = take(c * 16);
  const size_t o_ep = take(c * 32 * De), o_lp = take(c * 32 * Dl);
  const size_t mis_c = M3D_CONNECT_CHECK ? c : 0;  // the path-walk records exist in verification builds only
  const size_t o_ma = take(mis_c * 16 * (De + Dl)), o_mb = take(mis_c * 16 * (De + Dl)), o_mc = take(mis_c * 4 * (De + Dl));
  const size_t cc = c * (size_t)De * (size_t)Dl;
  const size_t o_corg = take(cc * 16), o_cdir = take(cc * 16), o_craw = take(cc * 16), o_cpay = take(cc * 16),
               o_cskip = take(cc * 4);
  const size_t o_work = take(c * (size_t)De * (size_t)(Dl + 1) * 4);
  const size_t o_counts = take(64);
  const size_t o_class = take((size_t)De * (size_t)(Dl + 1) * 4);
  const size_t o_chunk = take((c / 256 + 1) * 4);
  const MisTab mt{De, Dl};
  const size_t o_mt = take(c * 8 * (size_t)mt.entries());
  M3D_CUDA(ctx->scratch[6].reserve(total));
  char *p = ctx->scratch[6].as<char>();
  b.cap = cap;
  b.De = De;
  b.Dl = Dl;
  b.ev = (float4 *)(p + o_ev);
  b.lv = (float4 *)(p + o_lv);
  b.ne = (int32_t *)(p + o_ne);
  b.nl = (int32_t *)(p + o_nl);
  for (int i = 0; i < 2; i++) {
    b.org[i] = (float4 *)(p + o_org[i]);
    b.dir[i] = (float4 *)(p + o_dir[i]);
    b.skip[i] = (int32_t *)(p + o_skip[i]);
    b.queue[i] = (int32_t *)(p + o_queue[i]);
  }
  b.raw = (float4 *)(p + o_raw);
  b.ender_full = (float4 *)(p + o_ef);
  b.ender_roul = (float4 *)(p + o_er);
  b.accum = (float4 *)(p + o_acc);
  b.eyepre = (double *)(p + o_ep);
  b.lightpre = (double *)(p + o_lp);
  b.misA = (float4 *)(p + o_ma);
  b.misB = (double2 *)(p + o_mb);
  b.misC = (float *)(p + o_mc);
  b.corg = (float4 *)(p + o_corg);
  b.cdir = (float4 *)(p + o_cdir);
  b.craw = (float4 *)(p + o_craw);
  b.cpay = (float4 *)(p + o_cpay);
  b.cskip = (int32_t *)(p + o_cskip);
  b.work = (uint32_t *)(p + o_work);
  b.counts = (int *)(p + o_counts);
  b.class_counts = (int *)(p + o_class);
  b.chunk_counts = (int *)(p + o_chunk);
  b.mistab = (double *)(p + o_mt);
  b.ray_total = (unsigned long long *)(p + o_counts + 32);
  return M3D_OK;
}

}  // namespace

static int32_t render_bidir_one_device(m3d_scene *scene, const m3d_camera *cam, const m3d_area_light *lights,
                                       int32_t num_lights, const m3d_bidir_params *params, int32_t width,
                                       int32_t height, const m3d_partition *part, int32_t sample_count,
                                       void *d_rgb_sum, void *d_rgb_sumsq, void *stream, m3d_stats *stats);

namespace m3d {
int32_t shard_render(m3d_scene *scene, bool adaptive, int32_t height, const m3d_partition *part,
                     int32_t sample_count, cudaStream_t stream, m3d_stats *stats,
                     const std::function<int32_t(m3d_scene *, const m3d_partition &, int32_t, cudaStream_t,
                                                 m3d_stats *)> &one);  // api_path.cu
}

extern "C" {

int32_t m3d_render_bidir_device(m3d_scene *scene, const m3d_camera *cam, const m3d_area_light *lights,
                                int32_t num_lights, const m3d_bidir_params *params, int32_t width,
                                int32_t height, const m3d_partition *part, int32_t sample_count,
                                void *d_rgb_sum, void *d_rgb_sumsq, void *stream, m3d_stats *stats) {
  if (!scene || !cam || !params || !lights || num_lights <= 0 || width <= 0 || height <= 0 || !d_rgb_sum ||
      sample_count < 0)
    return fail(M3D_ERR_INVALID_ARG, "m3d_render_bidir: bad arguments");
  M3D_LOCK(scene->ctx);
  if (!scene->replicas.empty() && sample_count > 0) {
    const bool adaptive = params->min_samples != 0 && params->max_stddev != 0;
    return shard_render(scene, adaptive, height, part, sample_count, (cudaStream_t)stream, stats,
                        [&](m3d_scene *si, const m3d_partition &pi, int32_t count, cudaStream_t s, m3d_stats *st) {
                          return render_bidir_one_device(si, cam, lights, num_lights, params, width, height, &pi,
                                                         count, d_rgb_sum, d_rgb_sumsq, s, st);
                        });
  }
  return render_bidir_one_device(scene, cam, lights, num_lights, params, width, height, part, sample_count,
                                 d_rgb_sum, d_rgb_sumsq, stream, stats);
}

}  // extern "C"

static int32_t render_bidir_one_device(m3d_scene *scene, const m3d_camera *cam, const m3d_area_light *lights,
                                       int32_t num_lights, const m3d_bidir_params *params, int32_t width,
                                       int32_t height, const m3d_partition *part, int32_t sample_count,
                                       void *d_rgb_sum, void *d_rgb_sumsq, void *stream, m3d_stats *stats) {
  const int max_depth = params->max_depth;
  const int max_ld = params->max_light_depth != 0 ? params->max_light_depth : max_depth;  // bidir.go:257-262
  if (max_depth < 1 || max_ld < 1) return fail(M3D_ERR_INVALID_ARG, "MaxDepth and MaxLightDepth must be >= 1");
  if (max_depth > kBidirMaxDepth || max_ld > kBidirMaxDepth)
    return fail(M3D_ERR_UNSUPPORTED, "BidirPathTracer depths above %d are not supported on the GPU path",
                kBidirMaxDepth);
  if ((int64_t)width * height > (int64_t)0x7fffffff / 4) return fail(M3D_ERR_INVALID_ARG, "frame too large");
  m3d_ctx *ctx = scene->ctx;
  const DeviceScene &sc = scene->dev;
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
  const int64_t npix = (int64_t)width * (row_end - row_begin);
  if (npix == 0 || sample_count == 0) return M3D_OK;
  // BidirPathTracer.MinSamples / MaxStddev / OversaturatedStddevs (bidir.go:45-52) arrive in the
  // trailing fields of m3d_bidir_params
  const bool adaptive = params->min_samples != 0 && params->max_stddev != 0;
  if (adaptive && (sample_begin != 0 || sample_count != params->num_samples || d_rgb_sumsq))
    return fail(M3D_ERR_INVALID_ARG,
                "adaptive sampling (MinSamples/MaxStddev) stops per pixel: it cannot be sharded by sample index "
                "(sample_begin must be 0 and sample_count == num_samples; shard by rows) and has no sumsq output");

  // ---- area-light tables (light.go:131-161, 237-252, 283-301) ------------------------------
  std::vector<DeviceAreaLight> hl((size_t)num_lights);
  std::vector<DeviceLightTri> ht;
  double total_light = 0;
  for (int i = 0; i < num_lights; i++) {
    const m3d_area_light &a = lights[i];
    if (a.object < 0 || a.object >= (int32_t)scene->object_kind.size())
      return fail(M3D_ERR_INVALID_ARG, "area light %d refers to object %d which does not exist", i, a.object);
    DeviceAreaLight &L = hl[(size_t)i];
    std::memset(&L, 0, sizeof(L));
    L.object = a.object;
    const double esum = a.emission[0] + a.emission[1] + a.emission[2];
    for (int k = 0; k < 3; k++) L.emission[k] = (float)a.emission[k];
    const int kind = scene->object_kind[(size_t)a.object];
    double total = 0;
    if (kind == SHAPE_SPHERE) {
      int si = -1;
      for (size_t q = 0; q < scene->host_shapes.size(); q++)
        if (scene->host_shapes[q].object == a.object) si = (int)q;
      const DeviceShape &sh = scene->host_shapes[(size_t)si];
      L.kind = SHAPE_SPHERE;
      L.surf = -2 - si;
      for (int k = 0; k < 3; k++) L.center[k] = (float)sh.p0[k];
      L.radius = (float)sh.radius;
      total = esum * 4 * M_PI * sh.radius * sh.radius;  // light.go:159-161
    } else if (kind == 0) {
      L.kind = 0;
      L.tri_begin = (int32_t)ht.size();
      const int64_t t0 = scene->object_tri_begin[(size_t)a.object], tn = scene->object_tri_count[(size_t)a.object];
      if (tn == 0) return fail(M3D_ERR_INVALID_ARG, "area light %d is an empty mesh", i);
      L.tri_count = (int32_t)tn;
      double area_sum = 0;
      for (int64_t t = 0; t < tn; t++) {
        const float *v = scene->merged_tris.data() + (size_t)(t0 + t) * 9;
        DeviceLightTri T;
        std::memset(&T, 0, sizeof(T));
        for (int k = 0; k < 9; k++) T.v[k] = v[k];
        const double e1[3] = {(double)v[3] - v[0], (double)v[4] - v[1], (double)v[5] - v[2]};
        const double e2[3] = {(double)v[6] - v[0], (double)v[7] - v[1], (double)v[8] - v[2]};
        const double c[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2],
                             e1[0] * e2[1] - e1[1] * e2[0]};
        const double cn = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
        area_sum += cn / 2;  // Triangle.Area (primitives.go:21-23)
        T.cumu_area = area_sum;
        for (int k = 0; k < 3; k++) T.n[k] = (float)(c[k] * (1.0 / cn));
        T.leaf_index = scene->leaf_of_merged[(size_t)(t0 + t)];
        ht.push_back(T);
      }
      L.total_area = area_sum;
      total = area_sum * esum;  // light.go:272-274
    } else {
      return fail(M3D_ERR_UNSUPPORTED, "area lights must be mesh or sphere objects (light %d is object kind %d)", i,
                  kind);
    }
    total_light += total;
    L.cumu_total = total_light;
  }
  if (!(total_light > 0)) return fail(M3D_ERR_INVALID_ARG, "the area lights emit nothing");

  DeviceBidirParams bp;
  std::memset(&bp, 0, sizeof(bp));
  bp.max_depth = max_depth;
  bp.max_light_depth = max_ld;
  bp.min_depth = params->min_depth;
  bp.cutoff = (float)params->cutoff;
  bp.antialias = (float)params->antialias;
  bp.eps = params->epsilon > 1e-7 ? (float)params->epsilon : 0.f;  // DefaultEpsilon 1e-8 -> surface skip ids
  bp.roulette_delta = params->roulette_delta;
  bp.power_heuristic = params->power_heuristic;
  bp.seed = params->seed;
  bp.num_lights = num_lights;
  bp.total_light = total_light;

  const size_t lights_bytes = align256(hl.size() * sizeof(DeviceAreaLight));
  M3D_CUDA(ctx->scratch[7].reserve(lights_bytes + std::max<size_t>(1, ht.size()) * sizeof(DeviceLightTri)));
  DeviceAreaLight *d_lights = ctx->scratch[7].as<DeviceAreaLight>();
  DeviceLightTri *d_tris = (DeviceLightTri *)(ctx->scratch[7].as<char>() + lights_bytes);
  M3D_CUDA(cudaMemcpyAsync(d_lights, hl.data(), hl.size() * sizeof(DeviceAreaLight), cudaMemcpyHostToDevice, s));
  if (!ht.empty())
    M3D_CUDA(cudaMemcpyAsync(d_tris, ht.data(), ht.size() * sizeof(DeviceLightTri), cudaMemcpyHostToDevice, s));

  // batch geometry: per-slot footprint is dominated by the stored path vertices
  const size_t per_slot = (size_t)(16 * kBidirVertexFields + 32 + (M3D_CONNECT_CHECK ? 36 : 0)) * (max_depth + max_ld) +
                          (size_t)max_depth * (max_ld + 1) * 72 + (size_t)MisTab{max_depth, max_ld}.entries() * 8 + 256;
#ifndef M3D_BIDIR_BUDGET_GB
#define M3D_BIDIR_BUDGET_GB 64  // device memory for one batch's path vertices / work lists (of 180 GB)
#endif
  static const int batch_log2 = [] {  // M3D_BIDIR_BATCH_LOG2: samples per batch (tuning runs)
    const char *e = getenv("M3D_BIDIR_BATCH_LOG2");
    const int v = e ? atoi(e) : 22;
    return (v < 10 || v > 22) ? 22 : v;
  }();
  // at most half of what is free on the device right now (beyond what this context already holds);
  // cudaMemGetInfo only when the batch does not fit what the last call of the same depths carved
  // (it costs 10-40 ms when several devices / processes share peer mappings)
  const int64_t total = npix * sample_count;
  const int64_t depth_key = (int64_t)max_depth * 1024 + max_ld;
  int64_t cap = std::min<int64_t>((int64_t)1 << batch_log2, total);
  if (!(ctx->bidir_carved_key == depth_key && cap <= ctx->bidir_carved_cap && ctx->scratch[6].bytes > 0)) {
    size_t free_b = 0, total_b = 0;
    M3D_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const size_t budget = std::min<size_t>((size_t)M3D_BIDIR_BUDGET_GB << 30, (free_b + ctx->scratch[6].bytes) / 2);
    cap = std::min<int64_t>(cap, (int64_t)(budget / per_slot));
  }
  cap = std::max<int64_t>(1, cap);
  const int64_t nP_max = std::min(npix, cap);
  BidirBuffers buf;
  if (int32_t rc = carve_bidir_buffers(ctx, cap, max_depth, max_ld, buf)) return rc;
  if (!(ctx->bidir_carved_key == depth_key && cap <= ctx->bidir_carved_cap)) {
    ctx->bidir_carved_key = depth_key;
    ctx->bidir_carved_cap = cap;
  }
  M3D_CUDA(cudaMemsetAsync(buf.ray_total, 0, sizeof(unsigned long long), s));
  const DeviceCamera dc = device_camera(*cam, width, height);

  auto trace = [&](const float4 *org, const float4 *dir, const int32_t *skip, float4 *raw, int64_t n_max,
                   const int *n_ptr) -> int32_t {
    TraceLaunch t;
    t.org_tmin = org;
    t.dir_tmax = dir;
    t.n = n_max;
    t.n_ptr = n_ptr;
    t.hit0 = raw;
    t.hit1 = nullptr;
    t.refine = false;
    t.counters = nullptr;
    t.skip_tris = skip;
    t.ray_counter = next_work_counter(ctx);
    if (!t.ray_counter) return fail(M3D_ERR_OOM, "work counter allocation failed");
    launch_trace_bvh_only(sc.bvh, t, s);
    return M3D_OK;
  };

  int64_t launches = 0;
  GpuTimer tm;
  tm.start(s);
  StageTimer stages(s);
  static const char *const kStageNames[] = {"eye", "light", "prefix", "connect", "visibility", "resolve", "flush"};
  // traces one batch of samples; leaves one colour per slot in buf.accum
  auto run_batch = [&](const PathBatch &b) -> int32_t {
    const int64_t n = (int64_t)b.nP * b.S;
    // eye sub-paths
    stages.mark(0);
    launch_bidir_eye_raygen(dc, bp, b, buf, s);
    launches++;
    int cur = 0;
    for (int depth = 0; depth < max_depth; depth++) {
      if (int32_t rc = trace(buf.org[cur], buf.dir[cur], buf.skip[cur], buf.raw, n, buf.counts + cur)) return rc;
      launch_bidir_eye_shade(sc, bp, b, buf, cur, depth, s);
      M3D_CUDA(cudaMemsetAsync(buf.counts + cur, 0, sizeof(int), s));
      cur ^= 1;
      launches += 2;
    }
    // light sub-paths
    stages.mark(1);
    launch_bidir_light_raygen(sc, bp, d_lights, d_tris, b, buf, s);
    launches++;
    cur = 0;
    for (int depth = 0; depth + 1 < max_ld; depth++) {
      if (int32_t rc = trace(buf.org[cur], buf.dir[cur], buf.skip[cur], buf.raw, n, buf.counts + cur)) return rc;
      launch_bidir_light_shade(sc, bp, b, buf, cur, depth, s);
      M3D_CUDA(cudaMemsetAsync(buf.counts + cur, 0, sizeof(int), s));
      cur ^= 1;
      launches += 2;
    }
    // connections: all (eye prefix, light prefix) pairs at once
    stages.mark(2);
    launch_bidir_prefix(bp, b, buf, s);
    stages.mark(3);
    launch_bidir_connect(sc, bp, b, buf, s);
    stages.mark(4);
    if (int32_t rc = trace(buf.corg, buf.cdir, buf.cskip, buf.craw, n * max_depth * max_ld, buf.counts + 2))
      return rc;
    stages.mark(5);
    launch_bidir_connect_resolve(sc, buf, s);
    stages.mark(6);
    launches += 4;
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
  stages.mark(7);
  tm.stop(s);
  unsigned long long rays = 0;
  M3D_CUDA(cudaMemcpyAsync(&rays, buf.ray_total, sizeof(rays), cudaMemcpyDeviceToHost, s));
  M3D_CUDA(cudaStreamSynchronize(s));  // also keeps the host light tables alive for the copies
  M3D_CUDA(cudaGetLastError());
  stages.report("m3d_render_bidir", kStageNames, 7);
  if (stats) {
    stats->rays = (int64_t)rays;
    stats->kernel_ms = tm.ms();
    stats->launches = launches;
    stats->samples = samples_taken;
  }
  return M3D_OK;
}

extern "C" {

int32_t m3d_render_bidir(m3d_scene *scene, const m3d_camera *cam, const m3d_area_light *lights, int32_t num_lights,
                         const m3d_bidir_params *params, int32_t width, int32_t height,
                         const m3d_partition *part, int32_t sample_count, float *rgb_sum, float *rgb_sumsq,
                         m3d_stats *stats) {
  if (!scene || !rgb_sum || width <= 0 || height <= 0)
    return fail(M3D_ERR_INVALID_ARG, "m3d_render_bidir: bad arguments");
  m3d_ctx *ctx = scene->ctx;
  M3D_LOCK(ctx);
  M3D_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)width * height * 3 * sizeof(float);
  M3D_CUDA(ctx->scratch[3].reserve(bytes * 2));
  float *d_sum = ctx->scratch[3].as<float>();
  float *d_sq = rgb_sumsq ? d_sum + (size_t)width * height * 3 : nullptr;
  M3D_CUDA(cudaMemsetAsync(d_sum, 0, bytes * (rgb_sumsq ? 2 : 1), ctx->stream));
  int32_t rc = m3d_render_bidir_device(scene, cam, lights, num_lights, params, width, height, part, sample_count,
                                       d_sum, d_sq, ctx->stream, stats);
  if (rc != M3D_OK) return rc;
  M3D_CUDA(cudaMemcpyAsync(rgb_sum, d_sum, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  if (rgb_sumsq) M3D_CUDA(cudaMemcpyAsync(rgb_sumsq, d_sq, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  M3D_CUDA(cudaStreamSynchronize(ctx->stream));
  if (stats) stats->d2h_bytes = (int64_t)(bytes * (rgb_sumsq ? 2 : 1));
  return M3D_OK;
}

}  // extern "C"
