// C ABI: multi-device contexts, memory shared with the caller (pinned host buffers, device
// buffers exported over CUDA IPC) -- include/m3d.h.
//
// Replaces the reference's scheduler across GPUs: render3d spreads the pixels of a frame over
// NumCPU goroutines (render3d/concurrency.go:17-43, ray_renderer.go:25-56); here one host thread
// drives each GPU of the node, the frame is partitioned by sample index (path tracers), row band
// (RayCaster, adaptive renders) or ray slice (first-hit batches), and the per-pixel sums meet in
// the primary GPU's accumulator through the flush kernel's red.add over NVLink peer mappings
// (path_kernels.cu) instead of a collective.
#include <algorithm>
#include <cstring>
#include <thread>

#include "api_common.h"
#include "scene_host.h"

namespace m3d {

int32_t parallel_members(int n, const std::function<int32_t(int)> &fn) {
  if (n <= 1) return fn(0);
  std::vector<int32_t> rc((size_t)n, M3D_OK);
  std::vector<std::string> msg((size_t)n);
  std::vector<std::thread> th;
  th.reserve((size_t)n - 1);
  for (int i = 1; i < n; i++)
    th.emplace_back([&, i] {
      rc[(size_t)i] = fn(i);
      if (rc[(size_t)i] != M3D_OK) msg[(size_t)i] = last_error_ref();
    });
  rc[0] = fn(0);
  if (rc[0] != M3D_OK) msg[0] = last_error_ref();
  for (auto &t : th) t.join();
  for (int i = 0; i < n; i++)
    if (rc[(size_t)i] != M3D_OK) {
      last_error_ref() = "device " + std::to_string(i) + " of the group: " + msg[(size_t)i];
      return rc[(size_t)i];
    }
  return M3D_OK;
}

int32_t replicate_buffer(m3d_ctx *dst_ctx, DevBuf &dst, const DevBuf &src, int src_device) {
  if (!src.p || !src.bytes) return M3D_OK;
  M3D_CUDA(cudaSetDevice(dst_ctx->device));
  M3D_CUDA(dst.reserve(src.bytes));
  M3D_CUDA(cudaMemcpyPeer(dst.p, dst_ctx->device, src.p, src_device, src.bytes));
  return M3D_OK;
}

int32_t replicate_mesh(m3d_mesh *mesh) {
  m3d_ctx *root = mesh->ctx;
  for (m3d_ctx *mc : root->members) {
    auto *r = new m3d_mesh();
    mesh->replicas.push_back(r);
    r->ctx = mc;
    r->info = mesh->info;
    std::memcpy(r->bmin, mesh->bmin, sizeof(r->bmin));
    std::memcpy(r->bmax, mesh->bmax, sizeof(r->bmax));
    if (int32_t rc = replicate_buffer(mc, r->nodes, mesh->nodes, root->device)) return rc;
    if (int32_t rc = replicate_buffer(mc, r->tris, mesh->tris, root->device)) return rc;
    if (int32_t rc = replicate_buffer(mc, r->vnormals, mesh->vnormals, root->device)) return rc;
    r->bvh = mesh->bvh;
    r->bvh.nodes = r->nodes.as<const uint4>();
    r->bvh.tris = r->tris.as<const float4>();
    r->bvh.vnormals = mesh->bvh.vnormals ? r->vnormals.as<const float4>() : nullptr;
  }
  cudaSetDevice(root->device);
  return M3D_OK;
}

int32_t replicate_scene(m3d_scene *scene) {
  m3d_ctx *root = scene->ctx;
  for (m3d_ctx *mc : root->members) {
    auto *r = new m3d_scene();
    scene->replicas.push_back(r);
    r->ctx = mc;
    r->host_shapes = scene->host_shapes;
    r->host_materials = scene->host_materials;
    r->object_material = scene->object_material;
    r->object_kind = scene->object_kind;
    r->object_tri_begin = scene->object_tri_begin;
    r->object_tri_count = scene->object_tri_count;
    r->merged_tris = scene->merged_tris;
    r->leaf_of_merged = scene->leaf_of_merged;
    std::memcpy(r->bmin, scene->bmin, sizeof(r->bmin));
    std::memcpy(r->bmax, scene->bmax, sizeof(r->bmax));
    r->info = scene->info;
    if (int32_t rc = replicate_buffer(mc, r->nodes, scene->nodes, root->device)) return rc;
    if (int32_t rc = replicate_buffer(mc, r->tris, scene->tris, root->device)) return rc;
    if (int32_t rc = replicate_buffer(mc, r->vnormals, scene->vnormals, root->device)) return rc;
    if (int32_t rc = replicate_buffer(mc, r->shapes, scene->shapes, root->device)) return rc;
    if (int32_t rc = replicate_buffer(mc, r->objects, scene->objects, root->device)) return rc;
    if (int32_t rc = replicate_buffer(mc, r->materials, scene->materials, root->device)) return rc;
    r->dev = scene->dev;
    r->dev.bvh.nodes = r->nodes.as<const uint4>();
    r->dev.bvh.tris = r->tris.as<const float4>();
    r->dev.bvh.vnormals = scene->dev.bvh.vnormals ? r->vnormals.as<const float4>() : nullptr;
    r->dev.shapes = scene->dev.shapes ? r->shapes.as<const DeviceShape>() : nullptr;
    r->dev.objects = r->objects.as<const DeviceObject>();
    r->dev.materials = scene->dev.materials ? r->materials.as<const DeviceMaterial>() : nullptr;
    if (int32_t rc = replicate_buffer(mc, r->shape_nodes, scene->shape_nodes, root->device)) return rc;
    if (int32_t rc = replicate_buffer(mc, r->shape_tris, scene->shape_tris, root->device)) return rc;
    if (scene->dev.shape_bvh.nodes) {
      r->dev.shape_bvh.nodes = r->shape_nodes.as<const uint4>();
      r->dev.shape_bvh.tris = r->shape_tris.as<const float4>();
    }
    // instances refer to this device's replica of their mesh
    r->host_instances = scene->host_instances;
    r->instance_meshes = scene->instance_meshes;
    const size_t member = scene->replicas.size() - 1;  // index of mc in root->members
    for (size_t k = 0; k < r->host_instances.size(); k++) {
      m3d_mesh *mesh = scene->instance_meshes[k];
      if (mesh->replicas.size() <= member)
        return fail(M3D_ERR_INVALID_ARG, "an instanced mesh has no replica on device %d", mc->device);
      r->host_instances[k].blas = mesh->replicas[member]->bvh;
    }
    if (!r->host_instances.empty()) {
      M3D_CUDA(cudaSetDevice(mc->device));
      M3D_CUDA(r->instances.reserve(r->host_instances.size() * sizeof(DeviceInstance)));
      M3D_CUDA(cudaMemcpy(r->instances.p, r->host_instances.data(),
                          r->host_instances.size() * sizeof(DeviceInstance), cudaMemcpyHostToDevice));
      r->dev.instances = r->instances.as<const DeviceInstance>();
    }
  }
  cudaSetDevice(root->device);
  return M3D_OK;
}

int32_t render_sharded(m3d_scene *scene, cudaStream_t stream, m3d_stats *stats,
                       const std::function<int32_t(int, m3d_scene *, cudaStream_t, m3d_stats *)> &fn) {
  m3d_ctx *root = scene->ctx;
  const int g = 1 + (int)scene->replicas.size();
  cudaStream_t s0 = stream ? stream : root->stream;
  M3D_CUDA(cudaSetDevice(root->device));
  cudaEvent_t ready = nullptr;
  M3D_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
  cudaError_t e = cudaEventRecord(ready, s0);
  if (e != cudaSuccess) {
    cudaEventDestroy(ready);
    return fail(M3D_ERR_CUDA, "cudaEventRecord: %s", cudaGetErrorString(e));
  }
  std::vector<m3d_stats> st((size_t)g);
  for (auto &x : st) std::memset(&x, 0, sizeof(x));
  const int32_t rc = parallel_members(g, [&](int i) -> int32_t {
    m3d_scene *si = i == 0 ? scene : scene->replicas[(size_t)i - 1];
    std::lock_guard<std::recursive_mutex> lock(si->ctx->mu);
    M3D_CUDA(cudaSetDevice(si->ctx->device));
    cudaStream_t s = i == 0 ? s0 : si->ctx->stream;
    if (i > 0) M3D_CUDA(cudaStreamWaitEvent(s, ready, 0));
    return fn(i, si, s, &st[(size_t)i]);
  });
  cudaSetDevice(root->device);
  cudaEventDestroy(ready);
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    for (const m3d_stats &x : st) {
      stats->rays += x.rays;
      stats->hits += x.hits;
      stats->kernel_ms = std::max(stats->kernel_ms, x.kernel_ms);  // the devices run side by side
      stats->launches += x.launches;
      stats->samples += x.samples;
    }
  }
  return rc;
}

}  // namespace m3d

using namespace m3d;

extern "C" {

int32_t m3d_ctx_create_multi(const int32_t *devices, int32_t n, m3d_ctx **out) {
  if (!out) return fail(M3D_ERR_INVALID_ARG, "m3d_ctx_create_multi: out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(M3D_ERR_CUDA, "no CUDA device available (%s); libm3dgpu has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  std::vector<int32_t> devs;
  if (!devices || n <= 0) {
    for (int i = 0; i < count; i++) devs.push_back(i);
  } else {
    devs.assign(devices, devices + n);
  }
  for (size_t i = 0; i < devs.size(); i++) {
    if (devs[i] < 0 || devs[i] >= count)
      return fail(M3D_ERR_INVALID_ARG, "device %d out of range (%d devices)", devs[i], count);
    for (size_t j = 0; j < i; j++)
      if (devs[j] == devs[i]) return fail(M3D_ERR_INVALID_ARG, "device %d listed twice", devs[i]);
  }
  m3d_ctx *root = nullptr;
  if (int32_t rc = m3d_ctx_create(devs[0], &root)) return rc;
  for (size_t i = 1; i < devs.size(); i++) {
    int can01 = 0, can10 = 0;
    cudaDeviceCanAccessPeer(&can01, devs[0], devs[i]);
    cudaDeviceCanAccessPeer(&can10, devs[i], devs[0]);
    if (!can01 || !can10) {
      m3d_ctx_destroy(root);
      return fail(M3D_ERR_UNSUPPORTED, "devices %d and %d cannot access each other's memory (no NVLink / P2P)",
                  devs[0], devs[i]);
    }
    m3d_ctx *mc = nullptr;
    if (int32_t rc = m3d_ctx_create(devs[i], &mc)) {
      m3d_ctx_destroy(root);
      return rc;
    }
    root->members.push_back(mc);
    // member -> primary: the flush kernels add into the primary's accumulator; primary -> member:
    // peer copies of the replicas.  "Already enabled" is fine (another context of this process).
    cudaSetDevice(devs[i]);
    cudaError_t pe = cudaDeviceEnablePeerAccess(devs[0], 0);
    if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) {
      m3d_ctx_destroy(root);
      return fail(M3D_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", devs[i], devs[0], cudaGetErrorString(pe));
    }
    cudaSetDevice(devs[0]);
    pe = cudaDeviceEnablePeerAccess(devs[i], 0);
    if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) {
      m3d_ctx_destroy(root);
      return fail(M3D_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", devs[0], devs[i], cudaGetErrorString(pe));
    }
    cudaGetLastError();  // clear the sticky "already enabled" status
  }
  cudaSetDevice(devs[0]);
  *out = root;
  return M3D_OK;
}

int32_t m3d_ctx_num_devices(const m3d_ctx *ctx) { return ctx ? group_size(ctx) : 0; }

// ---- pinned host memory -------------------------------------------------------------------

int32_t m3d_host_alloc(int64_t bytes, void **out) {
  if (!out || bytes <= 0) return fail(M3D_ERR_INVALID_ARG, "m3d_host_alloc: bad arguments");
  *out = nullptr;
  // portable: page-locked for every device of the process (multi-device contexts copy from all of them)
  M3D_CUDA(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocPortable));
  return M3D_OK;
}

int32_t m3d_host_free(void *ptr) {
  if (!ptr) return M3D_OK;
  M3D_CUDA(cudaFreeHost(ptr));
  return M3D_OK;
}

int32_t m3d_host_register(void *ptr, int64_t bytes) {
  if (!ptr || bytes <= 0) return fail(M3D_ERR_INVALID_ARG, "m3d_host_register: bad arguments");
  M3D_CUDA(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
  return M3D_OK;
}

int32_t m3d_host_unregister(void *ptr) {
  if (!ptr) return M3D_OK;
  M3D_CUDA(cudaHostUnregister(ptr));
  return M3D_OK;
}

// ---- device buffers shared between the processes of a one-process-per-GPU job ----------------

int32_t m3d_device_alloc(m3d_ctx *ctx, int64_t bytes, void **d_ptr) {
  if (!ctx || !d_ptr || bytes <= 0) return fail(M3D_ERR_INVALID_ARG, "m3d_device_alloc: bad arguments");
  *d_ptr = nullptr;
  M3D_LOCK(ctx);
  M3D_CUDA(cudaSetDevice(ctx->device));
  // a plain cudaMalloc (not a pool / caching allocator block): its base address is what
  // cudaIpcGetMemHandle exports
  M3D_CUDA(cudaMalloc(d_ptr, (size_t)bytes));
  M3D_CUDA(cudaMemset(*d_ptr, 0, (size_t)bytes));
  return M3D_OK;
}

int32_t m3d_device_free(m3d_ctx *ctx, void *d_ptr) {
  if (!ctx) return fail(M3D_ERR_INVALID_ARG, "m3d_device_free: ctx is NULL");
  if (!d_ptr) return M3D_OK;
  M3D_LOCK(ctx);
  M3D_CUDA(cudaSetDevice(ctx->device));
  M3D_CUDA(cudaFree(d_ptr));
  return M3D_OK;
}

static_assert(sizeof(cudaIpcMemHandle_t) == M3D_IPC_HANDLE_BYTES, "IPC handle size");

int32_t m3d_ipc_export(m3d_ctx *ctx, void *d_ptr, uint8_t *handle) {
  if (!ctx || !d_ptr || !handle) return fail(M3D_ERR_INVALID_ARG, "m3d_ipc_export: bad arguments");
  M3D_LOCK(ctx);
  M3D_CUDA(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  M3D_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
  std::memcpy(handle, &h, sizeof(h));
  return M3D_OK;
}

int32_t m3d_ipc_open(m3d_ctx *ctx, const uint8_t *handle, void **d_ptr) {
  if (!ctx || !d_ptr || !handle) return fail(M3D_ERR_INVALID_ARG, "m3d_ipc_open: bad arguments");
  *d_ptr = nullptr;
  M3D_LOCK(ctx);
  M3D_CUDA(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  // maps the exporting GPU's allocation into this process; enables peer access to that GPU
  M3D_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return M3D_OK;
}

int32_t m3d_ipc_close(m3d_ctx *ctx, void *d_ptr) {
  if (!ctx) return fail(M3D_ERR_INVALID_ARG, "m3d_ipc_close: ctx is NULL");
  if (!d_ptr) return M3D_OK;
  M3D_LOCK(ctx);
  M3D_CUDA(cudaSetDevice(ctx->device));
  M3D_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return M3D_OK;
}

}  // extern "C"
