// C ABI: library, context and mesh-collider entry points (include/m3d.h).
// Replaces model3d.MeshToCollider + Collider.FirstRayCollision
// (model3d/collisions.go:138-142, 275-290) with a device-resident wide BVH and a
// batched query.  No CPU fallback: every compute call needs a CUDA device.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "api_common.h"
#include "trace_core.cuh"

namespace m3d {

std::string &last_error_ref() {
  static thread_local std::string s;
  return s;
}

int32_t fail(int32_t code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}

static bool ensure_counters(m3d_ctx *ctx) {
  if (ctx->counters.p) return true;
  const size_t bytes = (4 + 4 * (size_t)m3d_ctx::kWorkSlots) * sizeof(unsigned long long);
  if (ctx->counters.reserve(bytes) != cudaSuccess) return false;
  return cudaMemset(ctx->counters.p, 0, bytes) == cudaSuccess;
}

unsigned long long *next_work_counter(m3d_ctx *ctx) {
  if (!ensure_counters(ctx)) return nullptr;
  const unsigned slot = ctx->work_slot++ % m3d_ctx::kWorkSlots;
  return ctx->counters.as<unsigned long long>() + 4 + 4 * slot;
}

unsigned long long *stats_counters(m3d_ctx *ctx, cudaStream_t s) {
  if (!ensure_counters(ctx)) return nullptr;
  if (cudaMemsetAsync(ctx->counters.p, 0, 4 * sizeof(unsigned long long), s) != cudaSuccess) return nullptr;
  return ctx->counters.as<unsigned long long>();
}

int32_t upload_bvh(m3d_ctx *ctx, const WideBVH &bvh, const float *vnormals_by_prim, DevBuf &nodes,
                   DevBuf &tris, DevBuf &vnormals, DeviceBVH &out) {
  const size_t nb = bvh.nodes.size() * sizeof(WideNode);
  const size_t tb = std::max<size_t>(bvh.tris.size(), 1) * sizeof(TriRecord);
  M3D_CUDA(nodes.reserve(nb));
  M3D_CUDA(tris.reserve(tb));
  M3D_CUDA(cudaMemcpyAsync(nodes.p, bvh.nodes.data(), nb, cudaMemcpyHostToDevice, ctx->stream));
  if (!bvh.tris.empty())
    M3D_CUDA(cudaMemcpyAsync(tris.p, bvh.tris.data(), bvh.tris.size() * sizeof(TriRecord),
                             cudaMemcpyHostToDevice, ctx->stream));
  out.nodes = nodes.as<const uint4>();
  out.tris = tris.as<const float4>();
  out.num_nodes = (int64_t)bvh.nodes.size();
  out.num_tris = (int64_t)bvh.tris.size();
  out.vnormals = nullptr;
  for (int k = 0; k < 3; k++) {
    out.bmin[k] = bvh.bounds_min[k];
    out.bmax[k] = bvh.bounds_max[k];
  }
  std::vector<float> vn;
  if (vnormals_by_prim && !bvh.tris.empty()) {
    // leaf order, padded to float4 per corner
    vn.resize(bvh.tris.size() * 12);
    for (size_t i = 0; i < bvh.tris.size(); i++) {
      const float *src = vnormals_by_prim + (size_t)bvh.tris[i].prim * 9;
      for (int k = 0; k < 3; k++) {
        vn[i * 12 + k * 4 + 0] = src[k * 3 + 0];
        vn[i * 12 + k * 4 + 1] = src[k * 3 + 1];
        vn[i * 12 + k * 4 + 2] = src[k * 3 + 2];
        vn[i * 12 + k * 4 + 3] = 0.f;
      }
    }
    M3D_CUDA(vnormals.reserve(vn.size() * sizeof(float)));
    M3D_CUDA(cudaMemcpyAsync(vnormals.p, vn.data(), vn.size() * sizeof(float), cudaMemcpyHostToDevice,
                             ctx->stream));
    out.vnormals = vnormals.as<const float4>();
  }
  M3D_CUDA(cudaStreamSynchronize(ctx->stream));
  return M3D_OK;
}

int32_t build_bvh_with_flags(m3d_ctx *ctx, const BuildInput &in, uint32_t build_flags, WideBVH &out) {
  if ((build_flags & M3D_MESH_BUILD_DEVICE_COLLAPSE) && in.n > 0) {
    out.nodes.clear();
    out.tris.clear();
    return lbvh_build_wide(ctx, in, bvh_cost_prim(), out);
  }
  if ((build_flags & M3D_MESH_BUILD_DEVICE_LBVH) && in.n > 0) {
    std::vector<BinaryNode> bn;
    std::vector<int32_t> order;
    int32_t root = 0;
    double device_ms = 0;
    if (int32_t rc = lbvh_build_binary(ctx, in.tris, in.n, bn, order, &root, &device_ms)) return rc;
    build_wide_bvh_from_binary(in, bn.data(), (int64_t)bn.size(), root, order.data(), out);
    out.build_ms += device_ms;
    return M3D_OK;
  }
  build_wide_bvh(in, out);
  return M3D_OK;
}

}  // namespace m3d

using namespace m3d;

extern "C" {

int32_t m3d_abi_version(void) { return M3D_ABI_VERSION; }

const char *m3d_last_error(void) { return last_error_ref().c_str(); }

int32_t m3d_ctx_create(int32_t device, m3d_ctx **out) {
  if (!out) return fail(M3D_ERR_INVALID_ARG, "m3d_ctx_create: out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(M3D_ERR_CUDA, "no CUDA device available (%s); libm3dgpu has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0) M3D_CUDA(cudaGetDevice(&device));
  if (device >= count) return fail(M3D_ERR_INVALID_ARG, "device %d out of range (%d devices)", device, count);
  M3D_CUDA(cudaSetDevice(device));
  auto *ctx = new m3d_ctx();
  ctx->device = device;
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return fail(M3D_ERR_CUDA, "cudaStreamCreate failed");
  }
  *out = ctx;
  return M3D_OK;
}

void m3d_ctx_destroy(m3d_ctx *ctx) {
  if (!ctx) return;
  for (m3d_ctx *mc : ctx->members) m3d_ctx_destroy(mc);
  ctx->members.clear();
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
  if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
  delete ctx;
}

int32_t m3d_ctx_device(const m3d_ctx *ctx) { return ctx ? ctx->device : -1; }

int32_t m3d_ctx_synchronize(m3d_ctx *ctx) {
  if (!ctx) return fail(M3D_ERR_INVALID_ARG, "ctx is NULL");
  M3D_LOCK(ctx);
  for (m3d_ctx *mc : ctx->members)
    if (int32_t rc = m3d_ctx_synchronize(mc)) return rc;
  M3D_CUDA(cudaSetDevice(ctx->device));
  M3D_CUDA(cudaStreamSynchronize(ctx->stream));
  M3D_CUDA(cudaStreamSynchronize(ctx->copy_in));
  M3D_CUDA(cudaStreamSynchronize(ctx->copy_out));
  return M3D_OK;
}

int32_t m3d_ctx_trim(m3d_ctx *ctx) {
  if (!ctx) return fail(M3D_ERR_INVALID_ARG, "ctx is NULL");
  M3D_LOCK(ctx);
  int32_t rc = m3d_ctx_synchronize(ctx);
  if (rc != M3D_OK) return rc;
  for (m3d_ctx *mc : ctx->members)
    if (int32_t rc2 = m3d_ctx_trim(mc)) return rc2;
  M3D_CUDA(cudaSetDevice(ctx->device));
  for (auto &b : ctx->scratch) b.release();
  return M3D_OK;
}

int32_t m3d_mesh_create(m3d_ctx *ctx, const float *tris, int64_t n, const float *vnormals,
                        uint32_t build_flags, m3d_mesh **out) {
  if (!ctx || !out || n < 0 || (n > 0 && !tris))
    return fail(M3D_ERR_INVALID_ARG, "m3d_mesh_create: bad arguments");
  if (n > (int64_t)0x7fffff00) return fail(M3D_ERR_INVALID_ARG, "too many triangles (%lld)", (long long)n);
  *out = nullptr;
  M3D_LOCK(ctx);
  M3D_CUDA(cudaSetDevice(ctx->device));
  for (int64_t i = 0; i < n * 9; i++)
    if (!(tris[i] == tris[i]) || tris[i] > 3e38f || tris[i] < -3e38f)
      return fail(M3D_ERR_INVALID_ARG, "non-finite vertex coordinate at float %lld", (long long)i);
  WideBVH bvh;
  BuildInput in;
  in.tris = tris;
  in.n = n;
  auto *m = new m3d_mesh();
  m->ctx = ctx;
  int64_t num_nodes = 0;
  if ((build_flags & M3D_MESH_BUILD_DEVICE_COLLAPSE) && n > 0) {
    // whole build on the device, arrays stay there (no round trip through the host)
    ResidentBVH res;
    res.nodes = &m->nodes;
    res.tris = &m->tris;
    res.vnormals = &m->vnormals;
    res.vnormals_by_prim = vnormals;
    if (int32_t rc = lbvh_build_wide(ctx, in, bvh_cost_prim(), bvh, &res)) {
      delete m;
      return rc;
    }
    m->bvh.nodes = m->nodes.as<const uint4>();
    m->bvh.tris = m->tris.as<const float4>();
    m->bvh.vnormals = vnormals ? m->vnormals.as<const float4>() : nullptr;
    m->bvh.num_nodes = res.num_nodes;
    m->bvh.num_tris = res.num_tris;
    for (int k = 0; k < 3; k++) {
      m->bvh.bmin[k] = bvh.bounds_min[k];
      m->bvh.bmax[k] = bvh.bounds_max[k];
    }
    num_nodes = res.num_nodes;
  } else {
    if (int32_t rc = build_bvh_with_flags(ctx, in, build_flags, bvh)) {
      delete m;
      return rc;
    }
    int32_t rc = upload_bvh(ctx, bvh, vnormals, m->nodes, m->tris, m->vnormals, m->bvh);
    if (rc != M3D_OK) {
      delete m;
      return rc;
    }
    num_nodes = (int64_t)bvh.nodes.size();
  }
  m->info.num_triangles = n;
  m->info.num_nodes = num_nodes;
  m->info.node_bytes = sizeof(WideNode);
  m->info.tri_bytes = sizeof(TriRecord);
  m->info.device_bytes = (int64_t)(m->nodes.bytes + m->tris.bytes + m->vnormals.bytes);
  m->info.max_depth = bvh.max_depth;
  m->info.build_ms = bvh.build_ms;
  m->info.sah_cost = bvh.sah_cost;
  for (int k = 0; k < 3; k++) {
    m->bmin[k] = n ? bvh.bounds_min[k] : 0.0;  // nullCollider bounds are zero (collisions.go:360-366)
    m->bmax[k] = n ? bvh.bounds_max[k] : 0.0;
  }
  // The traversal kernels keep 64 stack entries per ray (10 in shared memory, the rest in local
  // memory, trace_kernels.cu) and the scalar walkers M3D_STACK_SIZE; a stack entry is one node
  // group, at most one per level, so a deeper hierarchy could drop pushes silently: refuse it.
  if (bvh.max_depth > M3D_MAX_BVH_DEPTH) {
    const int d = bvh.max_depth;
    delete m;
    return fail(M3D_ERR_UNSUPPORTED, "BVH depth %d exceeds the traversal stack (%d levels)", d, M3D_MAX_BVH_DEPTH);
  }
  if (int32_t rc2 = replicate_mesh(m)) {
    m3d_mesh_destroy(m);
    return rc2;
  }
  *out = m;
  return M3D_OK;
}

void m3d_mesh_destroy(m3d_mesh *mesh) {
  if (!mesh) return;
  for (m3d_mesh *r : mesh->replicas) m3d_mesh_destroy(r);
  mesh->replicas.clear();
  std::lock_guard<std::recursive_mutex> lock(mesh->ctx->mu);
  cudaSetDevice(mesh->ctx->device);
  delete mesh;
}

int32_t m3d_mesh_get_info(const m3d_mesh *mesh, m3d_mesh_info *info) {
  if (!mesh || !info) return fail(M3D_ERR_INVALID_ARG, "m3d_mesh_get_info: NULL argument");
  *info = mesh->info;
  return M3D_OK;
}

int32_t m3d_mesh_bounds(const m3d_mesh *mesh, double min_out[3], double max_out[3]) {
  if (!mesh || !min_out || !max_out) return fail(M3D_ERR_INVALID_ARG, "m3d_mesh_bounds: NULL argument");
  for (int k = 0; k < 3; k++) {
    min_out[k] = mesh->bmin[k];
    max_out[k] = mesh->bmax[k];
  }
  return M3D_OK;
}

int32_t m3d_mesh_first_ray_collisions_device(m3d_mesh *mesh, const void *d_org_tmin,
                                             const void *d_dir_tmax, int64_t n, void *d_hit0,
                                             void *d_hit1, uint32_t flags, void *stream,
                                             m3d_stats *stats) {
  if (!mesh || n < 0 || (n > 0 && (!d_org_tmin || !d_dir_tmax || !d_hit0 || !d_hit1)))
    return fail(M3D_ERR_INVALID_ARG, "m3d_mesh_first_ray_collisions_device: bad arguments");
  if (n > (int64_t)0x7ff00000)
    return fail(M3D_ERR_INVALID_ARG, "batch too large (%lld rays); split it", (long long)n);
  m3d_ctx *ctx = mesh->ctx;
  M3D_LOCK(ctx);
  M3D_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  TraceLaunch p;
  p.org_tmin = (const float4 *)d_org_tmin;
  p.dir_tmax = (const float4 *)d_dir_tmax;
  p.n = n;
  p.hit0 = (float4 *)d_hit0;
  p.hit1 = (float4 *)d_hit1;
  p.refine = !(flags & M3D_TRACE_NO_REFINE);
  p.counters = nullptr;
  p.ray_counter = next_work_counter(ctx);
  if (!p.ray_counter) return fail(M3D_ERR_OOM, "work counter allocation failed");
  if (trace_cull_wanted(n, mesh->bvh.num_nodes)) {  // survivor list of the bounds cull
    M3D_CUDA(ctx->scratch[13].reserve((size_t)(n + 1) * sizeof(int)));
    p.cull_scratch = ctx->scratch[13].as<int>();
  }
  if (flags & M3D_TRACE_COUNTERS) {
    p.counters = stats_counters(ctx, s);
    if (!p.counters) return fail(M3D_ERR_CUDA, "counter allocation failed");
  }
  if (stats) {
    GpuTimer tm;
    tm.start(s);
    launch_trace_first_hit(mesh->bvh, p, s);
    tm.stop(s);
    M3D_CUDA(cudaGetLastError());
    std::memset(stats, 0, sizeof(*stats));
    stats->kernel_ms = tm.ms();
    stats->rays = n;
    stats->launches = n > 0 ? 2 : 0;
    if (p.counters) {
      unsigned long long c[2];
      M3D_CUDA(cudaMemcpyAsync(c, p.counters, sizeof(c), cudaMemcpyDeviceToHost, s));
      M3D_CUDA(cudaStreamSynchronize(s));
      stats->nodes_visited = (int64_t)c[0];
      stats->tris_tested = (int64_t)c[1];
    }
  } else {
    launch_trace_first_hit(mesh->bvh, p, s);
    M3D_CUDA(cudaGetLastError());
  }
  return M3D_OK;
}

static constexpr int kMaxPipeBuf = 8;

static int32_t first_ray_collisions_one_device(m3d_mesh *mesh, const float *org, const float *dir, int64_t n,
                                               float *t, int32_t *prim, float *normal, float *bary,
                                               uint32_t flags, m3d_stats *stats);

static void add_stats(m3d_stats *sum, const m3d_stats &s) {
  sum->rays += s.rays;
  sum->hits += s.hits;
  sum->nodes_visited += s.nodes_visited;
  sum->tris_tested += s.tris_tested;
  sum->kernel_ms = std::max(sum->kernel_ms, s.kernel_ms);  // the devices run side by side
  sum->h2d_bytes += s.h2d_bytes;
  sum->d2h_bytes += s.d2h_bytes;
  sum->launches += s.launches;
  sum->samples += s.samples;
}

int32_t m3d_mesh_first_ray_collisions(m3d_mesh *mesh, const float *org, const float *dir, int64_t n,
                                      float *t, int32_t *prim, float *normal, float *bary,
                                      uint32_t flags, m3d_stats *stats) {
  if (!mesh || n < 0 || (n > 0 && (!org || !dir)))
    return fail(M3D_ERR_INVALID_ARG, "m3d_mesh_first_ray_collisions: bad arguments");
  M3D_LOCK(mesh->ctx);
  const int g = 1 + (int)mesh->replicas.size();
  // multi-device context: contiguous slices of the ray arrays, one host thread and one copy /
  // compute pipeline per GPU, no exchange between them
  if (g > 1 && n >= (int64_t)g * 4096) {
    std::vector<m3d_stats> st((size_t)g);
    int32_t rc = parallel_members(g, [&](int i) -> int32_t {
      int64_t b, e;
      split_range(n, g, i, &b, &e);
      m3d_mesh *mi = i == 0 ? mesh : mesh->replicas[(size_t)i - 1];
      std::lock_guard<std::recursive_mutex> lock(mi->ctx->mu);
      return first_ray_collisions_one_device(mi, (flags & M3D_TRACE_SHARED_ORIGIN) ? org : org + 3 * b,
                                             dir + 3 * b, e - b, t ? t + b : nullptr,
                                             prim ? prim + b : nullptr, normal ? normal + 3 * b : nullptr,
                                             bary ? bary + 3 * b : nullptr, flags, stats ? &st[(size_t)i] : nullptr);
    });
    if (stats) {
      std::memset(stats, 0, sizeof(*stats));
      for (const m3d_stats &x : st) add_stats(stats, x);
    }
    return rc;
  }
  return first_ray_collisions_one_device(mesh, org, dir, n, t, prim, normal, bary, flags, stats);
}

static int32_t first_ray_collisions_one_device(m3d_mesh *mesh, const float *org, const float *dir, int64_t n,
                                               float *t, int32_t *prim, float *normal, float *bary,
                                               uint32_t flags, m3d_stats *stats) {
  m3d_ctx *ctx = mesh->ctx;
  M3D_CUDA(cudaSetDevice(ctx->device));
  if (stats) std::memset(stats, 0, sizeof(*stats));
  if (n == 0) return M3D_OK;

  // Chunked 3-stage pipeline: H2D (copy_in) -> pack + trace + unpack (stream) -> D2H (copy_out).
  // Three stages need three buffers in flight to overlap fully (a fourth absorbs jitter); small
  // chunks keep the exposed pipeline fill (first H2D) and drain (last kernels + D2H) short.
  // M3D_PIPE_CHUNK_LOG2 / M3D_PIPE_NBUF override both for tuning runs.
  static const int64_t kChunk = [] {
    const char *e = getenv("M3D_PIPE_CHUNK_LOG2");
    int lg = e ? atoi(e) : 20;
    if (lg < 14 || lg > 24) lg = 20;
    return (int64_t)1 << lg;
  }();
  static const int nbuf = [] {
    const char *f = getenv("M3D_PIPE_NBUF");
    int nb = f ? atoi(f) : 4;
    return (nb < 2 || nb > kMaxPipeBuf) ? 4 : nb;
  }();
  // per buffer: org3, dir3, org4, dir4, hit0, hit1, out_t, out_prim, out_normal, out_bary
  const size_t per = (size_t)kChunk;
  const size_t sz_in3 = per * 3 * sizeof(float), sz_f4 = per * sizeof(float4);
  const size_t sz_out = per * (sizeof(float) + sizeof(int32_t) + 6 * sizeof(float));
  const size_t per_buf = 2 * sz_in3 + 4 * sz_f4 + sz_out;
  const int64_t chunk = n < kChunk ? n : kChunk;
  (void)chunk;
  M3D_CUDA(ctx->scratch[0].reserve(per_buf * nbuf));
  cudaEvent_t ev_in[kMaxPipeBuf], ev_k[kMaxPipeBuf], ev_out[kMaxPipeBuf];
  for (int b = 0; b < nbuf; b++) {
    cudaEventCreateWithFlags(&ev_in[b], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ev_k[b], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ev_out[b], cudaEventDisableTiming);
  }
  unsigned long long *counters = nullptr;
  if (flags & M3D_TRACE_COUNTERS) {
    counters = stats_counters(ctx, ctx->stream);
    if (!counters) return fail(M3D_ERR_CUDA, "counter allocation failed");
  }
  cudaEvent_t k0 = nullptr, k1 = nullptr;
  double kernel_ms = 0;
  if (stats) {
    cudaEventCreate(&k0);
    cudaEventCreate(&k1);
  }
  int32_t rc = M3D_OK;
  int64_t launches = 0;
  int it = 0;
  for (int64_t base = 0; base < n && rc == M3D_OK; base += kChunk, it++) {
    const int b = it % nbuf;
    const int64_t m = std::min<int64_t>(kChunk, n - base);
    char *buf = ctx->scratch[0].as<char>() + per_buf * b;
    float *d_org3 = (float *)buf;
    float *d_dir3 = (float *)(buf + sz_in3);
    float4 *d_org4 = (float4 *)(buf + 2 * sz_in3);
    float4 *d_dir4 = d_org4 + per;
    float4 *d_hit0 = d_dir4 + per;
    float4 *d_hit1 = d_hit0 + per;
    float *d_t = (float *)(d_hit1 + per);
    int32_t *d_prim = (int32_t *)(d_t + per);
    float *d_normal = (float *)(d_prim + per);
    float *d_bary = d_normal + 3 * per;
    // the previous user of this buffer must have finished its D2H
    if (it >= nbuf) cudaStreamWaitEvent(ctx->copy_in, ev_out[b], 0);
    const bool shared_org = (flags & M3D_TRACE_SHARED_ORIGIN) != 0;
    if (!shared_org)
      cudaMemcpyAsync(d_org3, org + base * 3, m * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->copy_in);
    cudaMemcpyAsync(d_dir3, dir + base * 3, m * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->copy_in);
    cudaEventRecord(ev_in[b], ctx->copy_in);
    cudaStreamWaitEvent(ctx->stream, ev_in[b], 0);
    if (stats) cudaEventRecord(k0, ctx->stream);
    launch_pack_rays(d_org3, d_dir3, m, 0.f, __builtin_inff(), d_org4, d_dir4, ctx->stream, shared_org ? org : nullptr);
    TraceLaunch p;
    p.org_tmin = d_org4;
    p.dir_tmax = d_dir4;
    p.n = m;
    p.hit0 = d_hit0;
    p.hit1 = d_hit1;
    p.refine = !(flags & M3D_TRACE_NO_REFINE);
    p.counters = counters;
    p.ray_counter = next_work_counter(ctx);
    if (!p.ray_counter) {
      rc = fail(M3D_ERR_OOM, "work counter allocation failed");
      break;
    }
    if (trace_cull_wanted(m, mesh->bvh.num_nodes)) {  // (the chunks' kernels run one after the other on ctx->stream)
      if (ctx->scratch[13].reserve((size_t)(chunk + 1) * sizeof(int)) != cudaSuccess) {
        rc = fail(M3D_ERR_OOM, "cull scratch allocation failed");
        break;
      }
      p.cull_scratch = ctx->scratch[13].as<int>();
    }
    launch_trace_first_hit(mesh->bvh, p, ctx->stream);
    launch_unpack_hits(d_hit0, d_hit1, m, t ? d_t : nullptr, prim ? d_prim : nullptr, nullptr,
                       normal ? d_normal : nullptr, bary ? d_bary : nullptr, ctx->stream);
    launches += 4;
    if (stats) {
      cudaEventRecord(k1, ctx->stream);
    }
    cudaEventRecord(ev_k[b], ctx->stream);
    cudaStreamWaitEvent(ctx->copy_out, ev_k[b], 0);
    if (t) cudaMemcpyAsync(t + base, d_t, m * sizeof(float), cudaMemcpyDeviceToHost, ctx->copy_out);
    if (prim) cudaMemcpyAsync(prim + base, d_prim, m * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->copy_out);
    if (normal)
      cudaMemcpyAsync(normal + base * 3, d_normal, m * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->copy_out);
    if (bary)
      cudaMemcpyAsync(bary + base * 3, d_bary, m * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->copy_out);
    cudaEventRecord(ev_out[b], ctx->copy_out);
    if (stats) {
      cudaEventSynchronize(k1);
      float f = 0;
      cudaEventElapsedTime(&f, k0, k1);
      kernel_ms += f;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = fail(M3D_ERR_CUDA, "first_ray_collisions pipeline: %s", cudaGetErrorString(e));
  }
  cudaError_t e = cudaStreamSynchronize(ctx->copy_out);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess && rc == M3D_OK) rc = fail(M3D_ERR_CUDA, "first_ray_collisions: %s", cudaGetErrorString(e));
  for (int b = 0; b < nbuf; b++) {
    cudaEventDestroy(ev_in[b]);
    cudaEventDestroy(ev_k[b]);
    cudaEventDestroy(ev_out[b]);
  }
  if (stats) {
    cudaEventDestroy(k0);
    cudaEventDestroy(k1);
    stats->rays = n;
    stats->kernel_ms = kernel_ms;
    stats->launches = launches;
    stats->h2d_bytes = n * ((flags & M3D_TRACE_SHARED_ORIGIN) ? 3 : 6) * (int64_t)sizeof(float);
    stats->d2h_bytes = n * (int64_t)((t ? 4 : 0) + (prim ? 4 : 0) + (normal ? 12 : 0) + (bary ? 12 : 0));
    if (counters && rc == M3D_OK) {
      unsigned long long c[2] = {0, 0};
      cudaMemcpy(c, counters, sizeof(c), cudaMemcpyDeviceToHost);
      stats->nodes_visited = (int64_t)c[0];
      stats->tris_tested = (int64_t)c[1];
    }
    if (prim && rc == M3D_OK) {
      int64_t h = 0;
      for (int64_t i = 0; i < n; i++) h += prim[i] >= 0;
      stats->hits = h;
    }
  }
  return rc;
}

int32_t m3d_mesh_ray_collision_counts(m3d_mesh *mesh, const float *org, const float *dir, int64_t n,
                                      int32_t *counts, m3d_stats *stats) {
  if (!mesh || n < 0 || (n > 0 && (!org || !dir || !counts)))
    return fail(M3D_ERR_INVALID_ARG, "m3d_mesh_ray_collision_counts: bad arguments");
  if (n > (int64_t)0x7ff00000) return fail(M3D_ERR_INVALID_ARG, "batch too large; split it");
  m3d_ctx *ctx = mesh->ctx;
  M3D_LOCK(ctx);
  M3D_CUDA(cudaSetDevice(ctx->device));
  if (stats) std::memset(stats, 0, sizeof(*stats));
  if (n == 0) return M3D_OK;
  cudaStream_t s = ctx->stream;
  const size_t per = (size_t)n;
  M3D_CUDA(ctx->scratch[9].reserve(per * (6 * sizeof(float) + sizeof(int32_t))));
  float *d_org = ctx->scratch[9].as<float>();
  float *d_dir = d_org + 3 * per;
  int32_t *d_counts = (int32_t *)(d_dir + 3 * per);
  M3D_CUDA(cudaMemcpyAsync(d_org, org, per * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
  M3D_CUDA(cudaMemcpyAsync(d_dir, dir, per * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
  GpuTimer tm;
  tm.start(s);
  launch_count_hits(mesh->bvh, d_org, d_dir, n, d_counts, nullptr, s);
  tm.stop(s);
  M3D_CUDA(cudaMemcpyAsync(counts, d_counts, per * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  M3D_CUDA(cudaStreamSynchronize(s));
  M3D_CUDA(cudaGetLastError());
  if (stats) {
    stats->rays = n;
    stats->kernel_ms = tm.ms();
    stats->launches = 1;
    stats->h2d_bytes = n * 24;
    stats->d2h_bytes = n * 4;
  }
  return M3D_OK;
}

int32_t m3d_mesh_ray_collisions(m3d_mesh *mesh, const float *org, const float *dir, int64_t n,
                                int64_t capacity, int64_t *offsets, float *t, int32_t *prim, float *normal,
                                float *bary, m3d_stats *stats) {
  if (!mesh || n < 0 || capacity < 0 || !offsets || (n > 0 && (!org || !dir)) ||
      (capacity > 0 && (!t || !prim)))
    return fail(M3D_ERR_INVALID_ARG, "m3d_mesh_ray_collisions: bad arguments");
  if (n > (int64_t)0x7ff00000) return fail(M3D_ERR_INVALID_ARG, "batch too large; split it");
  m3d_ctx *ctx = mesh->ctx;
  M3D_LOCK(ctx);
  M3D_CUDA(cudaSetDevice(ctx->device));
  if (stats) std::memset(stats, 0, sizeof(*stats));
  offsets[0] = 0;
  if (n == 0) return M3D_OK;
  cudaStream_t s = ctx->stream;
  const size_t per = (size_t)n;
  // rays + counts, then (n + 1) offsets; 8-byte aligned
  const size_t ray_bytes = per * (6 * sizeof(float) + sizeof(int32_t));
  const size_t off_at = (ray_bytes + 7) & ~(size_t)7;
  M3D_CUDA(ctx->scratch[9].reserve(off_at + (per + 1) * sizeof(int64_t)));
  float *d_org = ctx->scratch[9].as<float>();
  float *d_dir = d_org + 3 * per;
  int32_t *d_counts = (int32_t *)(d_dir + 3 * per);
  int64_t *d_off = (int64_t *)(ctx->scratch[9].as<char>() + off_at);
  M3D_CUDA(cudaMemcpyAsync(d_org, org, per * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
  M3D_CUDA(cudaMemcpyAsync(d_dir, dir, per * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
  GpuTimer tm, tm2;
  tm.start(s);
  launch_count_hits(mesh->bvh, d_org, d_dir, n, d_counts, nullptr, s);
  tm.stop(s);
  std::vector<int32_t> counts(per);
  M3D_CUDA(cudaMemcpyAsync(counts.data(), d_counts, per * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  M3D_CUDA(cudaStreamSynchronize(s));
  int64_t total = 0;
  for (size_t i = 0; i < per; i++) {
    offsets[i] = total;
    total += counts[i];
  }
  offsets[per] = total;
  double ms = tm.ms();
  int launches = 1;
  int64_t d2h = n * 4;
  if (total > 0 && total <= capacity) {
    const size_t tot = (size_t)total;
    const size_t out_bytes = tot * (sizeof(float) + sizeof(int32_t) + (normal ? 12 : 0) + (bary ? 12 : 0));
    M3D_CUDA(ctx->scratch[10].reserve(out_bytes));
    float *d_t = ctx->scratch[10].as<float>();
    int32_t *d_prim = (int32_t *)(d_t + tot);
    float *d_normal = normal ? (float *)(d_prim + tot) : nullptr;
    float *d_bary = bary ? (float *)(d_prim + tot) + (normal ? 3 * tot : 0) : nullptr;
    M3D_CUDA(cudaMemcpyAsync(d_off, offsets, (per + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    tm2.start(s);
    launch_collect_hits(mesh->bvh, d_org, d_dir, n, d_off, d_t, d_prim, d_normal, d_bary, s);
    tm2.stop(s);
    M3D_CUDA(cudaMemcpyAsync(t, d_t, tot * sizeof(float), cudaMemcpyDeviceToHost, s));
    M3D_CUDA(cudaMemcpyAsync(prim, d_prim, tot * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (normal) M3D_CUDA(cudaMemcpyAsync(normal, d_normal, tot * 12, cudaMemcpyDeviceToHost, s));
    if (bary) M3D_CUDA(cudaMemcpyAsync(bary, d_bary, tot * 12, cudaMemcpyDeviceToHost, s));
    M3D_CUDA(cudaStreamSynchronize(s));
    ms += tm2.ms();
    launches++;
    d2h += (int64_t)out_bytes;
  }
  M3D_CUDA(cudaGetLastError());
  if (stats) {
    stats->rays = n;
    stats->hits = total;
    stats->kernel_ms = ms;
    stats->launches = launches;
    stats->h2d_bytes = n * 24 + (launches > 1 ? (n + 1) * 8 : 0);
    stats->d2h_bytes = d2h;
  }
  return M3D_OK;
}

static int32_t check_sdf_depth(const m3d_mesh *mesh, const char *who) {
  if (7 * (int64_t)mesh->info.max_depth + 1 > sdf_stack_capacity())
    return fail(M3D_ERR_UNSUPPORTED, "%s: BVH depth %d exceeds the nearest-triangle traversal stack", who,
                mesh->info.max_depth);
  return M3D_OK;
}

int32_t m3d_mesh_contains(m3d_mesh *mesh, const float *points, int64_t n, double margin, uint8_t *inside,
                          m3d_stats *stats) {
  if (!mesh || n < 0 || (n > 0 && (!points || !inside)) || margin != margin)
    return fail(M3D_ERR_INVALID_ARG, "m3d_mesh_contains: bad arguments");
  if (n > (int64_t)0x7ff00000) return fail(M3D_ERR_INVALID_ARG, "batch too large; split it");
  if (margin != 0) {
    int32_t rc = check_sdf_depth(mesh, "m3d_mesh_contains");
    if (rc != M3D_OK) return rc;
  }
  m3d_ctx *ctx = mesh->ctx;
  M3D_LOCK(ctx);
  M3D_CUDA(cudaSetDevice(ctx->device));
  if (stats) std::memset(stats, 0, sizeof(*stats));
  if (n == 0) return M3D_OK;
  cudaStream_t s = ctx->stream;
  const size_t per = (size_t)n;
  M3D_CUDA(ctx->scratch[9].reserve(per * (3 * sizeof(float) + 3) + 64));
  float *d_pts = ctx->scratch[9].as<float>();
  uint8_t *d_inside = (uint8_t *)(d_pts + 3 * per);
  uint8_t *d_parity = d_inside + per, *d_near = d_parity + per;
  M3D_CUDA(cudaMemcpyAsync(d_pts, points, per * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
  GpuTimer tm;
  tm.start(s);
  int launches = 1;
  if (margin == 0) {
    launch_count_hits(mesh->bvh, d_pts, nullptr, n, nullptr, d_inside, s);
  } else {
    // collisions.go:127-133: the parity decides which way the sphere test is read
    launch_count_hits(mesh->bvh, d_pts, nullptr, n, nullptr, d_parity, s);
    launch_sphere_collisions(mesh->bvh, d_pts, nullptr, (float)std::fabs(margin), n, d_near, s);
    launch_contains_margin(d_parity, d_near, n, margin < 0, d_inside, s);
    launches = 3;
  }
  tm.stop(s);
  M3D_CUDA(cudaMemcpyAsync(inside, d_inside, per, cudaMemcpyDeviceToHost, s));
  M3D_CUDA(cudaStreamSynchronize(s));
  M3D_CUDA(cudaGetLastError());
  if (stats) {
    stats->rays = n;
    stats->kernel_ms = tm.ms();
    stats->launches = launches;
    stats->h2d_bytes = n * 12;
    stats->d2h_bytes = n;
  }
  return M3D_OK;
}

int32_t m3d_mesh_sdf(m3d_mesh *mesh, const float *points, int64_t n, float *sdf, float *closest, int32_t *face,
                     float *normal, m3d_stats *stats) {
  if (!mesh || n < 0 || (n > 0 && !points))
    return fail(M3D_ERR_INVALID_ARG, "m3d_mesh_sdf: bad arguments");
  if (mesh->info.num_triangles == 0)
    return fail(M3D_ERR_INVALID_ARG, "m3d_mesh_sdf: cannot create empty SDF");  // sdf.go:198-200 panics
  if (n > (int64_t)0x7ff00000) return fail(M3D_ERR_INVALID_ARG, "batch too large; split it");
  int32_t rc = check_sdf_depth(mesh, "m3d_mesh_sdf");
  if (rc != M3D_OK) return rc;
  m3d_ctx *ctx = mesh->ctx;
  M3D_LOCK(ctx);
  M3D_CUDA(cudaSetDevice(ctx->device));
  if (stats) std::memset(stats, 0, sizeof(*stats));
  if (n == 0) return M3D_OK;
  cudaStream_t s = ctx->stream;
  const size_t per = (size_t)n;
  // points | sdf | closest | normal | face
  M3D_CUDA(ctx->scratch[9].reserve(per * (3 + 1 + 3 + 3 + 1) * sizeof(float) + 64));
  float *d_pts = ctx->scratch[9].as<float>();
  float *d_sdf = d_pts + 3 * per, *d_cp = d_sdf + per, *d_nrm = d_cp + 3 * per;
  int32_t *d_face = (int32_t *)(d_nrm + 3 * per);
  M3D_CUDA(cudaMemcpyAsync(d_pts, points, per * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
  GpuTimer tm;
  tm.start(s);
  unsigned long long *counters = stats ? stats_counters(ctx, s) : nullptr;  // work counters ride along with stats
  launch_mesh_sdf(mesh->bvh, d_pts, n, sdf ? d_sdf : nullptr, closest ? d_cp : nullptr, face ? d_face : nullptr,
                  normal ? d_nrm : nullptr, counters, s);
  tm.stop(s);
  int64_t out_bytes = 0;
  if (sdf) {
    M3D_CUDA(cudaMemcpyAsync(sdf, d_sdf, per * 4, cudaMemcpyDeviceToHost, s));
    out_bytes += n * 4;
  }
  if (closest) {
    M3D_CUDA(cudaMemcpyAsync(closest, d_cp, per * 12, cudaMemcpyDeviceToHost, s));
    out_bytes += n * 12;
  }
  if (normal) {
    M3D_CUDA(cudaMemcpyAsync(normal, d_nrm, per * 12, cudaMemcpyDeviceToHost, s));
    out_bytes += n * 12;
  }
  if (face) {
    M3D_CUDA(cudaMemcpyAsync(face, d_face, per * 4, cudaMemcpyDeviceToHost, s));
    out_bytes += n * 4;
  }
  M3D_CUDA(cudaStreamSynchronize(s));
  M3D_CUDA(cudaGetLastError());
  if (stats) {
    stats->rays = n;
    stats->kernel_ms = tm.ms();
    stats->launches = 1;
    stats->h2d_bytes = n * 12;
    stats->d2h_bytes = out_bytes;
    if (counters) {
      unsigned long long c[3] = {0, 0, 0};
      cudaMemcpy(c, counters, sizeof(c), cudaMemcpyDeviceToHost);
      stats->nodes_visited = (int64_t)c[0];
      stats->tris_tested = (int64_t)c[1];
      stats->hits = (int64_t)c[2];  // float64 Triangle.Closest evaluations
    }
  }
  return M3D_OK;
}

int32_t m3d_mesh_sphere_collisions(m3d_mesh *mesh, const float *centers, const float *radii, int64_t n,
                                   uint8_t *collides, m3d_stats *stats) {
  if (!mesh || n < 0 || (n > 0 && (!centers || !radii || !collides)))
    return fail(M3D_ERR_INVALID_ARG, "m3d_mesh_sphere_collisions: bad arguments");
  if (n > (int64_t)0x7ff00000) return fail(M3D_ERR_INVALID_ARG, "batch too large; split it");
  int32_t rc = check_sdf_depth(mesh, "m3d_mesh_sphere_collisions");
  if (rc != M3D_OK) return rc;
  m3d_ctx *ctx = mesh->ctx;
  M3D_LOCK(ctx);
  M3D_CUDA(cudaSetDevice(ctx->device));
  if (stats) std::memset(stats, 0, sizeof(*stats));
  if (n == 0) return M3D_OK;
  cudaStream_t s = ctx->stream;
  const size_t per = (size_t)n;
  M3D_CUDA(ctx->scratch[9].reserve(per * (4 * sizeof(float) + 1) + 64));
  float *d_pts = ctx->scratch[9].as<float>();
  float *d_rad = d_pts + 3 * per;
  uint8_t *d_out = (uint8_t *)(d_rad + per);
  M3D_CUDA(cudaMemcpyAsync(d_pts, centers, per * 12, cudaMemcpyHostToDevice, s));
  M3D_CUDA(cudaMemcpyAsync(d_rad, radii, per * 4, cudaMemcpyHostToDevice, s));
  GpuTimer tm;
  tm.start(s);
  launch_sphere_collisions(mesh->bvh, d_pts, d_rad, 0.f, n, d_out, s);
  tm.stop(s);
  M3D_CUDA(cudaMemcpyAsync(collides, d_out, per, cudaMemcpyDeviceToHost, s));
  M3D_CUDA(cudaStreamSynchronize(s));
  M3D_CUDA(cudaGetLastError());
  if (stats) {
    stats->rays = n;
    stats->kernel_ms = tm.ms();
    stats->launches = 1;
    stats->h2d_bytes = n * 16;
    stats->d2h_bytes = n;
  }
  return M3D_OK;
}

}  // extern "C"