// Shared internals of the C ABI implementation (include/m3d.h).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
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

}  // namespace m3d

struct m3d_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;       // compute
  cudaStream_t copy_in = nullptr;      // H2D
  cudaStream_t copy_out = nullptr;     // D2H
  // scratch reused by host-buffer calls (grown on demand, never shrunk)
  m3d::DevBuf scratch[12];
  m3d::DevBuf counters;       // [0..3] node/triangle statistics, then kWorkSlots work counters
  unsigned work_slot = 0;
  static constexpr int kWorkSlots = 64;
};

namespace m3d {
// Device u64 work counter for one persistent-kernel launch (rotating pool so that launches
// in flight on different streams never share one).  nullptr on allocation failure.
unsigned long long *next_work_counter(m3d_ctx *ctx);
// Device 2 x u64 statistics counters, zeroed on `s`.
unsigned long long *stats_counters(m3d_ctx *ctx, cudaStream_t s);
}  // namespace m3d

struct m3d_mesh {
  m3d_ctx *ctx = nullptr;
  m3d::DevBuf nodes, tris, vnormals;
  m3d::DeviceBVH bvh;
  m3d_mesh_info info{};
  double bmin[3] = {0, 0, 0}, bmax[3] = {0, 0, 0};
};

namespace m3d {
// Device LBVH (lbvh.cu): binary hierarchy from Morton codes, downloaded for the wide collapse.
int32_t lbvh_build_binary(m3d_ctx *ctx, const float *tris, int64_t n, std::vector<BinaryNode> &nodes,
                          std::vector<int32_t> &order, int32_t *root_out, double *device_ms);
// Full device build (lbvh.cu): LBVH + cost-optimal 8-wide collapse + node emission on the device.
int32_t lbvh_build_wide(m3d_ctx *ctx, const BuildInput &in, double cost_prim_value, WideBVH &out);
// Builds the compressed wide BVH with the builder build_flags selects.
int32_t build_bvh_with_flags(m3d_ctx *ctx, const BuildInput &in, uint32_t build_flags, WideBVH &out);
// uploads a built BVH (and optional per-corner normals in caller order, remapped to leaf order)
int32_t upload_bvh(m3d_ctx *ctx, const WideBVH &bvh, const float *vnormals_by_prim, DevBuf &nodes,
                   DevBuf &tris, DevBuf &vnormals, DeviceBVH &out);
}  // namespace m3d
