// sm_100a kernels for batched first-hit queries:
//   trace_first_hit_kernel  == JoinedCollider.FirstRayCollision per ray
//                              (model3d/collisions.go:275-290), see trace_core.cuh
//   pack_rays / unpack_hits == marshalling between the host ABI's n*3 arrays
//                              (Ray{Origin,Direction}, collisions.go:12-15;
//                              RayCollision / TriangleCollision, collisions.go:19-46)
//                              and the device float4 SoA layout.
#include "kernels.h"
#include "trace_core.cuh"

namespace m3d {

namespace {

constexpr int kTraceBlock = 128;

template <bool COUNT>
__global__ void __launch_bounds__(kTraceBlock)
trace_first_hit_kernel(DeviceBVH bvh, TraceLaunch p) {
  const int64_t i = (int64_t)blockIdx.x * kTraceBlock + threadIdx.x;
  TraceCounters cnt;
  cnt.nodes = 0;
  cnt.tris = 0;
  if (i < p.n) {
    const float4 o = __ldg(p.org_tmin + i);
    const float4 d = __ldg(p.dir_tmax + i);
    RayF ray;
    ray.ox = o.x;
    ray.oy = o.y;
    ray.oz = o.z;
    ray.tmin = o.w;
    ray.dx = d.x;
    ray.dy = d.y;
    ray.dz = d.z;
    ray.tmax = d.w;
    HitF h;
    trace_bvh<COUNT, false>(bvh.nodes, bvh.tris, ray, -1, h, &cnt);

    float4 h0 = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
    float4 h1 = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
    if (h.tri >= 0) {
      const float4 *tri = bvh.tris + (size_t)h.tri * 3;
      const int prim = __float_as_int(__ldg(&tri[0].w));
      const int obj = __float_as_int(__ldg(&tri[1].w));
      if (p.refine) {
        const HitD r = refine_hit_f64(tri, o.x, o.y, o.z, d.x, d.y, d.z);
        const double t = r.t >= 0.0 ? r.t : (double)h.t;
        double nx = r.nx, ny = r.ny, nz = r.nz;
        if (bvh.vnormals) {
          // InterpNormalTriangle.InterpNormal (primitives.go:508-516)
          const float4 *vn = bvh.vnormals + (size_t)h.tri * 3;
          const float4 a = __ldg(vn), b = __ldg(vn + 1), c = __ldg(vn + 2);
          nx = r.b0 * a.x + r.b1 * b.x + r.b2 * c.x;
          ny = r.b0 * a.y + r.b1 * b.y + r.b2 * c.y;
          nz = r.b0 * a.z + r.b1 * b.z + r.b2 * c.z;
          const double s = 1.0 / sqrt(nx * nx + ny * ny + nz * nz);
          nx *= s;
          ny *= s;
          nz *= s;
        }
        h0 = make_float4((float)t, (float)r.b1, (float)r.b2, __int_as_float(prim));
        h1 = make_float4((float)nx, (float)ny, (float)nz, __int_as_float(obj));
      } else {
        const float4 q0 = __ldg(tri), q1 = __ldg(tri + 1), q2 = __ldg(tri + 2);
        const float e1x = q1.x - q0.x, e1y = q1.y - q0.y, e1z = q1.z - q0.z;
        const float e2x = q2.x - q0.x, e2y = q2.y - q0.y, e2z = q2.z - q0.z;
        float nx = e1y * e2z - e1z * e2y, ny = e1z * e2x - e1x * e2z, nz = e1x * e2y - e1y * e2x;
        if (bvh.vnormals) {
          const float4 *vn = bvh.vnormals + (size_t)h.tri * 3;
          const float4 a = __ldg(vn), b = __ldg(vn + 1), c = __ldg(vn + 2);
          const float b0 = 1.f - (h.b1 + h.b2);
          nx = b0 * a.x + h.b1 * b.x + h.b2 * c.x;
          ny = b0 * a.y + h.b1 * b.y + h.b2 * c.y;
          nz = b0 * a.z + h.b1 * b.z + h.b2 * c.z;
        }
        const float s = rsqrtf(nx * nx + ny * ny + nz * nz);
        h0 = make_float4(h.t, h.b1, h.b2, __int_as_float(prim));
        h1 = make_float4(nx * s, ny * s, nz * s, __int_as_float(obj));
      }
    }
    p.hit0[i] = h0;
    p.hit1[i] = h1;
  }
  if (COUNT) {
    unsigned long long n = cnt.nodes, t = cnt.tris;
    for (int off = 16; off > 0; off >>= 1) {
      n += __shfl_down_sync(0xffffffffu, n, off);
      t += __shfl_down_sync(0xffffffffu, t, off);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(p.counters, n);
      atomicAdd(p.counters + 1, t);
    }
  }
}

__global__ void pack_rays_kernel(const float *__restrict__ org3, const float *__restrict__ dir3,
                                 int64_t n, float tmin, float tmax, float4 *__restrict__ org_tmin,
                                 float4 *__restrict__ dir_tmax) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  org_tmin[i] = make_float4(org3[3 * i], org3[3 * i + 1], org3[3 * i + 2], tmin);
  dir_tmax[i] = make_float4(dir3[3 * i], dir3[3 * i + 1], dir3[3 * i + 2], tmax);
}

__global__ void unpack_hits_kernel(const float4 *__restrict__ hit0, const float4 *__restrict__ hit1,
                                   int64_t n, float *__restrict__ t, int32_t *__restrict__ prim,
                                   int32_t *__restrict__ obj, float *__restrict__ normal3,
                                   float *__restrict__ bary3) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 a = hit0[i], b = hit1[i];
  if (t) t[i] = a.x;
  if (prim) prim[i] = __float_as_int(a.w);
  if (obj) obj[i] = __float_as_int(b.w);
  if (normal3) {
    normal3[3 * i] = b.x;
    normal3[3 * i + 1] = b.y;
    normal3[3 * i + 2] = b.z;
  }
  if (bary3) {
    const bool hit = __float_as_int(a.w) >= 0;
    bary3[3 * i] = hit ? 1.f - (a.y + a.z) : 0.f;
    bary3[3 * i + 1] = a.y;
    bary3[3 * i + 2] = a.z;
  }
}

}  // namespace

void launch_trace_first_hit(const DeviceBVH &bvh, const TraceLaunch &p, cudaStream_t stream) {
  if (p.n <= 0) return;
  const unsigned blocks = (unsigned)((p.n + kTraceBlock - 1) / kTraceBlock);
  if (p.counters)
    trace_first_hit_kernel<true><<<blocks, kTraceBlock, 0, stream>>>(bvh, p);
  else
    trace_first_hit_kernel<false><<<blocks, kTraceBlock, 0, stream>>>(bvh, p);
}

void launch_pack_rays(const float *org3, const float *dir3, int64_t n, float tmin, float tmax,
                      float4 *org_tmin, float4 *dir_tmax, cudaStream_t stream) {
  if (n <= 0) return;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  pack_rays_kernel<<<blocks, 256, 0, stream>>>(org3, dir3, n, tmin, tmax, org_tmin, dir_tmax);
}

void launch_unpack_hits(const float4 *hit0, const float4 *hit1, int64_t n, float *t, int32_t *prim,
                        int32_t *obj, float *normal3, float *bary3, cudaStream_t stream) {
  if (n <= 0) return;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  unpack_hits_kernel<<<blocks, 256, 0, stream>>>(hit0, hit1, n, t, prim, obj, normal3, bary3);
}

int device_sm_count() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

}  // namespace m3d
