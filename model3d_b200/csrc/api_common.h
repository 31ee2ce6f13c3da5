// Shared internals of the C ABI implementation (include/m3d.h).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/m3d.h"
#include "kernels.h"
#include "wide_bvh.h"

namespace m3d {

std::string &last_error_ref();
int32_t fail(int32_t code, const char *fmt, ...);

#define M3D_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return ::m3d::fail(_e == cudaErrorMemoryAllocation ? M3D_ERR_OOM : M3D_ERR_CUDA,      \
                         "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                         __LINE__);                                                         \
  } while (0)

// RAII device buffer
struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  cudaError_t reserve(size_t n) {
    if (n <= bytes) return cudaSuccess;
    release();
    cudaError_t e = cudaMalloc(&p, n);
    if (e == cudaSuccess) bytes = n;
    return e;
  }
  template <class T>
  T *as() const {
    return reinterpret_cast<T *>(p);
  }
};

struct GpuTimer {
  cudaEvent_t a = nullptr, b = nullptr;
  GpuTimer() {
    cudaEventCreate(&a);
    cudaEventCreate(&b);
  }
  ~GpuTimer() {
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
  }
  void start(cudaStream_t s) { cudaEventRecord(a, s); }
  void stop(cudaStream_t s) { cudaEventRecord(b, s); }
  double ms() {
    float f = 0;
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&f, a, b);
    return f;
  }
};

// Per-stage device times of a multi-launch render, for tuning runs: with M3D_STAGE_TIMING set in the
// environment, mark(k) records an event where stage k begins (and the previous one ends) and report()
// prints the summed times per stage to stderr after the stream has been synchronised.  Off: no events.
struct StageTimer {
  bool on = false;
  cudaStream_t stream = nullptr;
  std::vector<std::pair<int, cudaEvent_t>> marks;
  explicit StageTimer(cudaStream_t s) : stream(s) { on = getenv("M3D_STAGE_TIMING") != nullptr; }
  ~StageTimer() {
    for (auto &m : marks) cudaEventDestroy(m.second);
  }
  void mark(int stage) {
    if (!on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, stream);
    marks.emplace_back(stage, e);
  }
  void report(const char *what, const char *const *names, int num_stages) {
    if (!on || marks.size() < 2) return;
    std::vector<double> ms((size_t)num_stages + 1, 0.0);
    for (size_t i = 0; i + 1 < marks.size(); i++) {
      float f = 0;
      cudaEventElapsedTime(&f, marks[i].second, marks[i + 1].second);
      const int k = marks[i].first;
      ms[(size_t)(k >= 0 && k < num_stages ? k : num_stages)] += f;
    }
    fprintf(stderr, "[m3d stage timing] %s:", what);
    for (int k = 0; k < num_stages; k++) fprintf(stderr, " %s %.2f ms", names[k], ms[(size_t)k]);
    fprintf(stderr, "\n");
  }
};

}  // namespace m3d

struct m3d_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;       // compute
  cudaStream_t copy_in = nullptr;      // H2D
  cudaStream_t copy_out = nullptr;     // D2H
  // scratch reused by host-buffer calls (grown on demand, never shrunk)
  m3d::DevBuf scratch[14];
  m3d::DevBuf counters;       // [0..3] node/triangle statistics, then kWorkSlots work counters
  unsigned work_slot = 0;
  std::vector<char> host_stage;  // small per-call tables staged for async uploads (light lists)
  int64_t bidir_carved_key = -1, bidir_carved_cap = 0;  // depths / batch size scratch[6] was last carved for
  static constexpr int kWorkSlots = 64;
  // Every entry point that touches scratch / streams / counters holds this lock for the whole
  // call: the reference's Collider and Object methods are "safe for concurrency"
  // (model3d/collisions.go:51) and its consumers call them from many goroutines at once, so calls
  // on one context serialise here (recursive: host-buffer calls nest the device-buffer ones).
  std::recursive_mutex mu;
  // Multi-device context (m3d_ctx_create_multi): this context is device 0 of the group and owns
  // the contexts of the other devices.  Meshes and scenes built on it carry one replica per member.
  std::vector<m3d_ctx *> members;
};
#define M3D_LOCK(ctx) std::lock_guard<std::recursive_mutex> m3d_lock_((ctx)->mu)

namespace m3d {
// Device u64 work counter for one persistent-kernel launch (rotating pool so that launches
// in flight on different streams never share one).  nullptr on allocation failure.
unsigned long long *next_work_counter(m3d_ctx *ctx);
// Device 2 x u64 statistics counters, zeroed on `s`.
unsigned long long *stats_counters(m3d_ctx *ctx, cudaStream_t s);
}  // namespace m3d

struct m3d_mesh {
  m3d_ctx *ctx = nullptr;
  std::vector<m3d_mesh *> replicas;  // multi-device context: copies on ctx->members[i] (owned)
  m3d::DevBuf nodes, tris, vnormals;
  m3d::DeviceBVH bvh;
  m3d_mesh_info info{};
  double bmin[3] = {0, 0, 0}, bmax[3] = {0, 0, 0};
};

namespace m3d {
// ---- multi-device contexts (api_multi.cu) ----------------------------------------------------
inline int group_size(const m3d_ctx *c) { return 1 + (int)c->members.size(); }
inline m3d_ctx *group_member(m3d_ctx *c, int i) { return i == 0 ? c : c->members[(size_t)i - 1]; }
// Runs fn(i) for i in [0, n): i == 0 on the calling thread, the others on one host thread each
// (the reference schedules its pixels over NumCPU goroutines, render3d/concurrency.go:17-43; here a
// thread drives one GPU).  Returns the first failing status and makes its message this thread's
// m3d_last_error().
int32_t parallel_members(int n, const std::function<int32_t(int)> &fn);
// Copies a device buffer of `src_device` into `dst` on dst_ctx's device (peer copy over NVLink).
int32_t replicate_buffer(m3d_ctx *dst_ctx, DevBuf &dst, const DevBuf &src, int src_device);
// Gives a mesh built on a multi-device context its replicas (no-op for single-device contexts).
int32_t replicate_mesh(m3d_mesh *mesh);
// [begin, end) of part i when n items are split into `parts` nearly equal contiguous ranges
inline void split_range(int64_t n, int parts, int i, int64_t *begin, int64_t *end) {
  const int64_t base = n / parts, extra = n % parts;
  *begin = base * i + (i < extra ? i : extra);
  *end = *begin + base + (i < extra ? 1 : 0);
}

// Device LBVH (lbvh.cu): binary hierarchy from Morton codes, downloaded for the wide collapse.
int32_t lbvh_build_binary(m3d_ctx *ctx, const float *tris, int64_t n, std::vector<BinaryNode> &nodes,
                          std::vector<int32_t> &order, int32_t *root_out, double *device_ms);
// Where a device build leaves its arrays when they are to stay on the device (lbvh_build_wide).
struct ResidentBVH {
  DevBuf *nodes = nullptr, *tris = nullptr, *vnormals = nullptr;
  const float *vnormals_by_prim = nullptr;  // optional host n*9 per-corner normals, caller order
  int64_t num_nodes = 0, num_tris = 0;
};
// Full device build (lbvh.cu): LBVH + cost-optimal 8-wide collapse + node emission on the device.
int32_t lbvh_build_wide(m3d_ctx *ctx, const BuildInput &in, double cost_prim_value, WideBVH &out,
                        ResidentBVH *resident = nullptr);
// Builds the compressed wide BVH with the builder build_flags selects.
int32_t build_bvh_with_flags(m3d_ctx *ctx, const BuildInput &in, uint32_t build_flags, WideBVH &out);
// uploads a built BVH (and optional per-corner normals in caller order, remapped to leaf order)
int32_t upload_bvh(m3d_ctx *ctx, const WideBVH &bvh, const float *vnormals_by_prim, DevBuf &nodes,
                   DevBuf &tris, DevBuf &vnormals, DeviceBVH &out);
}  // namespace m3d
