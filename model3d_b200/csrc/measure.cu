// Diagnostics: on-chip bandwidth microbenchmarks that give the roofline of the traversal kernel
// its denominator ("vs L2 peak using the full B_ray").  The BVH of every BASELINE
// config is L2 resident, so node / triangle fetches are bounded by what L2 delivers to the SMs:
//   mode 0  sequential: every SM streams a working set that fits L2 (but not the L1s) with
//           coalesced 16-byte ld.global.cg loads -> the L2 -> SM bandwidth peak;
//   mode 1  node gather: every lane reads the five 16-byte quads of a pseudo-random 80-byte record
//           (the wide-node access pattern of trace_first_hit_kernel: 32 different nodes per warp
//           instruction) -> what L2 delivers for divergent sector-granular reads.
// Not on the product path: called by bench.py / scripts only.
#include <algorithm>
#include <cstring>

#include "api_common.h"

namespace m3d {
namespace {

__global__ void __launch_bounds__(256)
l2_stream_kernel(const uint4 *__restrict__ p, size_t n_quads, int iters, unsigned long long *sink) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (int it = 0; it < iters; it++) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n_quads; i += 4 * stride) {
      const uint4 a = __ldcg(p + i), b = __ldcg(p + i + stride), c = __ldcg(p + i + 2 * stride),
                  d = __ldcg(p + i + 3 * stride);
      acc.x ^= a.x ^ b.x ^ c.x ^ d.x;
      acc.y ^= a.y ^ b.y ^ c.y ^ d.y;
      acc.z ^= a.z ^ b.z ^ c.z ^ d.z;
      acc.w ^= a.w ^ b.w ^ c.w ^ d.w;
    }
    for (; i < n_quads; i += stride) {
      const uint4 a = __ldcg(p + i);
      acc.x ^= a.x;
      acc.y ^= a.y;
      acc.z ^= a.z;
      acc.w ^= a.w;
    }
  }
  if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x9e3779b9u) atomicAdd(sink, 1ull);  // keeps the loads alive
}

__global__ void __launch_bounds__(256)
l2_gather_kernel(const uint4 *__restrict__ p, uint32_t n_records, int per_thread, unsigned long long *sink) {
  uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (int k = 0; k < per_thread; k += 2) {
    // two independent records in flight per trip (the traversal has one node + one triangle)
    x = x * 1664525u + 1013904223u;
    const uint32_t r0 = (uint32_t)(((unsigned long long)x * n_records) >> 32);
    x = x * 1664525u + 1013904223u;
    const uint32_t r1 = (uint32_t)(((unsigned long long)x * n_records) >> 32);
    const uint4 *a = p + (size_t)r0 * 5, *b = p + (size_t)r1 * 5;
#pragma unroll
    for (int q = 0; q < 5; q++) {
      const uint4 u = __ldcg(a + q), v = __ldcg(b + q);
      acc.x ^= u.x ^ v.x;
      acc.y ^= u.y ^ v.y;
      acc.z ^= u.z ^ v.z;
      acc.w ^= u.w ^ v.w;
    }
  }
  if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x9e3779b9u) atomicAdd(sink, 1ull);
}

// 256-bit loads (sm_100: ld.global.v8.u32): records of REC_BYTES at a REC_BYTES stride, read as
// V8 x 32-byte loads + V4 x 16-byte loads per lane (e.g. 96-byte stride: 2 + 1 = 80 useful bytes).
__device__ __forceinline__ void ldcg256(const void *p, uint4 &a, uint4 &b) {
  asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p));
}
template <int REC_BYTES, int V8, int V4>
__global__ void __launch_bounds__(256)
l2_gather_wide_kernel(const char *__restrict__ p, uint32_t n_records, int per_thread, unsigned long long *sink) {
  uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (int k = 0; k < per_thread; k += 2) {
    x = x * 1664525u + 1013904223u;
    const uint32_t r0 = (uint32_t)(((unsigned long long)x * n_records) >> 32);
    x = x * 1664525u + 1013904223u;
    const uint32_t r1 = (uint32_t)(((unsigned long long)x * n_records) >> 32);
    const char *a = p + (size_t)r0 * REC_BYTES, *b = p + (size_t)r1 * REC_BYTES;
#pragma unroll
    for (int q = 0; q < V8; q++) {
      uint4 u0, u1, v0, v1;
      ldcg256(a + 32 * q, u0, u1);
      ldcg256(b + 32 * q, v0, v1);
      acc.x ^= u0.x ^ v0.x ^ u1.x ^ v1.x;
      acc.y ^= u0.y ^ v0.y ^ u1.y ^ v1.y;
      acc.z ^= u0.z ^ v0.z ^ u1.z ^ v1.z;
      acc.w ^= u0.w ^ v0.w ^ u1.w ^ v1.w;
    }
#pragma unroll
    for (int q = 0; q < V4; q++) {
      const uint4 u = __ldcg((const uint4 *)(a + 32 * V8) + q), v = __ldcg((const uint4 *)(b + 32 * V8) + q);
      acc.x ^= u.x ^ v.x;
      acc.y ^= u.y ^ v.y;
      acc.z ^= u.z ^ v.z;
      acc.w ^= u.w ^ v.w;
    }
  }
  if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x9e3779b9u) atomicAdd(sink, 1ull);
}

}  // namespace
}  // namespace m3d

using namespace m3d;

extern "C" int32_t m3d_measure_l2_bandwidth(m3d_ctx *ctx, int32_t mode, int64_t working_set_bytes,
                                            int32_t repeats, double *gb_per_s) {
  if (!ctx || !gb_per_s || working_set_bytes < 4096 || repeats <= 0 || mode < 0 || mode > 5)
    return fail(M3D_ERR_INVALID_ARG, "m3d_measure_l2_bandwidth: bad arguments");
  M3D_LOCK(ctx);
  M3D_CUDA(cudaSetDevice(ctx->device));
  DevBuf buf, sink;
  // modes 2..5: 256-bit loads -- 2: 96-byte records read as 2 x 32 + 16 bytes (80 useful),
  // 3: 64-byte records (2 x 32), 4: 128-byte records (4 x 32), 5: 96-byte records read whole (3 x 32)
  const size_t rec = mode <= 1 ? 80 : (mode == 3 ? 64 : (mode == 4 ? 128 : 96));
  const size_t useful = mode == 2 ? 80 : rec;
  const size_t n_rec = (size_t)working_set_bytes / rec;
  const size_t bytes = n_rec * rec;
  M3D_CUDA(buf.reserve(bytes));
  M3D_CUDA(sink.reserve(8));
  M3D_CUDA(cudaMemsetAsync(buf.p, 0x5a, bytes, ctx->stream));
  M3D_CUDA(cudaMemsetAsync(sink.p, 0, 8, ctx->stream));
  const int grid = ctx->sm_count * 8;
  const int per_thread = 64;
  double moved = 0;
  auto launch = [&](int iters) {
    if (mode == 0) {
      l2_stream_kernel<<<grid, 256, 0, ctx->stream>>>(buf.as<uint4>(), bytes / 16, iters,
                                                      sink.as<unsigned long long>());
      moved = (double)bytes * iters;
    } else {
      unsigned long long *sk = sink.as<unsigned long long>();
      const int pt = per_thread * iters;
      if (mode == 1) l2_gather_kernel<<<grid, 256, 0, ctx->stream>>>(buf.as<uint4>(), (uint32_t)n_rec, pt, sk);
      if (mode == 2) l2_gather_wide_kernel<96, 2, 1><<<grid, 256, 0, ctx->stream>>>(buf.as<char>(), (uint32_t)n_rec, pt, sk);
      if (mode == 3) l2_gather_wide_kernel<64, 2, 0><<<grid, 256, 0, ctx->stream>>>(buf.as<char>(), (uint32_t)n_rec, pt, sk);
      if (mode == 4) l2_gather_wide_kernel<128, 4, 0><<<grid, 256, 0, ctx->stream>>>(buf.as<char>(), (uint32_t)n_rec, pt, sk);
      if (mode == 5) l2_gather_wide_kernel<96, 3, 0><<<grid, 256, 0, ctx->stream>>>(buf.as<char>(), (uint32_t)n_rec, pt, sk);
      moved = (double)grid * 256 * per_thread * iters * (double)useful;
    }
  };
  launch(2);  // warm the L2
  GpuTimer tm;
  double best = 0;
  for (int r = 0; r < repeats; r++) {
    tm.start(ctx->stream);
    launch(8);
    tm.stop(ctx->stream);
    const double ms = tm.ms();
    M3D_CUDA(cudaGetLastError());
    if (ms > 0) best = std::max(best, moved / (ms * 1e-3) / 1e9);
  }
  *gb_per_s = best;
  return M3D_OK;
}
