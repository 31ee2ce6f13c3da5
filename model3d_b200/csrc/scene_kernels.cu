// Scene-level kernels:
//   finish_scene_hits_kernel  closest-of over the merged triangle BVH hit and the analytic
//                             shapes == JoinedObject.Cast (render3d/object.go:141-153) with
//                             Sphere/Rect/Cylinder.FirstRayCollision (model3d/shapes.go:35-93,
//                             177-247,601-705,816-856) and the float64 re-evaluation of the
//                             winning triangle (model3d/primitives.go:27-33,207-249,508-516)
//   raygen_camera_kernel      Camera.Caster (render3d/camera.go:74-82,100-113)
//   shade_raycast_kernel      RayCaster.Render body (render3d/raycast.go:25-37)
//   finalize_image_kernel     colorSum/numSamples, sRGB-8 (ray_renderer.go:150, image.go:125-145,
//                             light.go:41-47)
#include "materials.cuh"
#include "scene.h"
#include "scene_hit.cuh"
#include "trace_core.cuh"

namespace m3d {

namespace {

// SHAPES = false: the scene is one triangle BVH (MeshCollider batches, the C2 headline): misses
// need no ray, and the kernel stays at half the registers of the general one.
template <int SHAPES>  // 0: mesh only, 1: few shapes (linear), 2: object-level BVH over the shapes
__global__ void __launch_bounds__(256, SHAPES ? 1 : 4)
finish_scene_hits_kernel(DeviceScene sc, SceneTraceLaunch sp) {
  const TraceLaunch &p = sp.t;
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= p.n) return;
  const float4 raw = __ldcs(p.hit0 + i);
  const bool need_ray = __float_as_int(raw.w) >= 0 || (SHAPES && sc.num_shapes > 0);
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f), d = make_float4(0.f, 0.f, 1.f, 0.f);
  if (need_ray) {
    o = __ldcs(p.org_tmin + i);
    d = __ldcs(p.dir_tmax + i);
  }
  const int skip = (SHAPES && sp.skip_ids) ? sp.skip_ids[i] : -1;
  const SceneHit h = resolve_scene_hit<SHAPES>(sc, o, d, raw, skip, p.refine);
  __stcs(p.hit0 + i, make_float4(h.t, h.b1, h.b2, __int_as_float(h.prim)));
  __stcs(p.hit1 + i, make_float4(h.nx, h.ny, h.nz, __int_as_float(h.obj)));
  if (sp.surf_ids) sp.surf_ids[i] = h.surf;
}

// Mesh-only finish pass with K rays per thread.  One ray per thread keeps only ~16 KB (raw hits),
// then ~24 KB (rays + triangles of the 30 % hits) in flight per SM, short of what HBM latency x
// bandwidth needs (~35 KB per SM): the K raw hits are loaded together, the rays and triangle
// records of the hits are pulled towards L1 with non-blocking prefetches, and only then does the
// thread run the float64 re-evaluation ray by ray.
#ifndef M3D_FINISH_MINB
#define M3D_FINISH_MINB 4
#endif
template <int K>
__global__ void __launch_bounds__(256, M3D_FINISH_MINB)
finish_mesh_hits_kernel(DeviceScene sc, SceneTraceLaunch sp) {
  const TraceLaunch &p = sp.t;
  const int64_t base = (int64_t)blockIdx.x * (256 * K) + threadIdx.x;
  float raw_t[K];
  int raw_tri[K];  // the raw record is (t, 0, 0, triangle)
#pragma unroll
  for (int k = 0; k < K; k++) {
    const int64_t i = base + 256 * k;
    const float4 r = i < p.n ? __ldcs(p.hit0 + i) : make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
    raw_t[k] = r.x;
    raw_tri[k] = __float_as_int(r.w);
  }
#pragma unroll
  for (int k = 0; k < K; k++) {
    const int tri_idx = raw_tri[k];
    if (tri_idx >= 0) {
      const int64_t i = base + 256 * k;
      const float4 *tri = sc.bvh.tris + (size_t)tri_idx * 3;
      asm volatile("prefetch.global.L1 [%0];" ::"l"(p.org_tmin + i));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(p.dir_tmax + i));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(tri));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(tri + 2));
    }
  }
#pragma unroll 1
  for (int k = 0; k < K; k++) {
    const int64_t i = base + 256 * k;
    if (i >= p.n) break;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f), d = make_float4(0.f, 0.f, 1.f, 0.f);
    float rt = raw_t[0];
    int rtri = raw_tri[0];
#pragma unroll
    for (int q = 1; q < K; q++) {  // no dynamic register indexing
      rt = k == q ? raw_t[q] : rt;
      rtri = k == q ? raw_tri[q] : rtri;
    }
    const float4 r = make_float4(rt, 0.f, 0.f, __int_as_float(rtri));
    if (rtri >= 0) {
      o = __ldcs(p.org_tmin + i);
      d = __ldcs(p.dir_tmax + i);
    }
    const SceneHit h = resolve_scene_hit<false>(sc, o, d, r, -1, p.refine);
    __stcs(p.hit0 + i, make_float4(h.t, h.b1, h.b2, __int_as_float(h.prim)));
    __stcs(p.hit1 + i, make_float4(h.nx, h.ny, h.nz, __int_as_float(h.obj)));
    if (sp.surf_ids) sp.surf_ids[i] = h.surf;
  }
}

__global__ void raygen_camera_kernel(DeviceCamera cam, int W, int row_begin, int row_end,
                                     float4 *__restrict__ org_tmin, float4 *__restrict__ dir_tmax) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = (int64_t)W * (row_end - row_begin);
  if (i >= n) return;
  const int x = (int)(i % W), y = row_begin + (int)(i / W);
  // float64 like the reference (camera.go:77-81), rounded once to the float32 ray
  const double fx = ((double)x - cam.cx) / cam.cx, fy = ((double)y - cam.cy) / cam.cy;
  org_tmin[i] = make_float4((float)cam.origin[0], (float)cam.origin[1], (float)cam.origin[2], 0.f);
  dir_tmax[i] = make_float4((float)(cam.x[0] * fx + cam.y[0] * fy + cam.z[0]),
                            (float)(cam.x[1] * fx + cam.y[1] * fy + cam.z[1]),
                            (float)(cam.x[2] * fx + cam.y[2] * fy + cam.z[2]), INFINITY);
}

__global__ void shade_raycast_kernel(DeviceScene sc, const DevicePointLight *__restrict__ lights,
                                     int num_lights, const float4 *__restrict__ org_tmin,
                                     const float4 *__restrict__ dir_tmax, const float4 *__restrict__ hit0,
                                     const float4 *__restrict__ hit1, int64_t n, float *__restrict__ rgb) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 h1 = hit1[i];
  const int obj = __float_as_int(h1.w);
  if (obj < 0) return;  // miss: the pixel keeps its previous value (raycast.go:26-28)
  const float4 h0 = hit0[i], o4 = org_tmin[i], d4 = dir_tmax[i];
  const V3f org = v3f(o4.x, o4.y, o4.z), dir = v3f(d4.x, d4.y, d4.z), nrm = v3f(h1.x, h1.y, h1.z);
  const V3f point = org + dir * h0.x;
  const MatAt m = material_at(sc, obj, point);
  V3f color = mat_ambient(sc, m) + mat_emission(sc, m);
  const V3f to_eye = normalize(org - point);
  for (int l = 0; l < num_lights; l++) {
    const DevicePointLight lt = lights[l];
    const V3f lo = v3f(lt.origin);
    const V3f brdf = mat_bsdf(sc, m, nrm, normalize(point - lo), to_eye);
    color = color + shade_collision(lt, nrm, lo - point) * brdf;
  }
  rgb[3 * i] = color.x;
  rgb[3 * i + 1] = color.y;
  rgb[3 * i + 2] = color.z;
}

__device__ __forceinline__ float gamma_compress(float u) {  // light.go:41-47
  return u <= 0.0031308f ? 12.92f * u : 1.055f * powf(u, 1.f / 2.4f) - 0.055f;
}

__global__ void finalize_image_kernel(const float *__restrict__ sum, int64_t n, float inv_samples,
                                      float *__restrict__ mean, uint8_t *__restrict__ srgb8) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = sum[i] * inv_samples;
  if (mean) mean[i] = v;
  if (srgb8) {
    const float c = fminf(1.f, fmaxf(0.f, v));  // ClampColor light.go:22-24
    srgb8[i] = (uint8_t)(gamma_compress(c) * (256.0f - 0.001f));
  }
}

// Image.Downsample (image.go:100-120): box filter, sum in float64 like the reference, * 1/f^2
__global__ void downsample_image_kernel(const float *__restrict__ src, int W, int H, int f, float *__restrict__ dst) {
  const int ow = W / f, oh = H / f;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ow * oh) return;
  const int ox = i % ow, oy = i / ow;
  double s0 = 0, s1 = 0, s2 = 0;
  for (int k = 0; k < f; k++)
    for (int l = 0; l < f; l++) {
      const float *p = src + ((size_t)(oy * f + k) * W + (ox * f + l)) * 3;
      s0 += p[0];
      s1 += p[1];
      s2 += p[2];
    }
  const double w = 1.0 / (double)(f * f);
  dst[(size_t)i * 3] = (float)(s0 * w);
  dst[(size_t)i * 3 + 1] = (float)(s1 * w);
  dst[(size_t)i * 3 + 2] = (float)(s2 * w);
}

}  // namespace

void launch_downsample_image(const float *src, int W, int H, int factor, float *dst, cudaStream_t stream) {
  const int n = (W / factor) * (H / factor);
  if (n <= 0) return;
  downsample_image_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(src, W, H, factor, dst);
}

void launch_finish_scene_hits(const DeviceScene &scene, const SceneTraceLaunch &p, cudaStream_t stream) {
  if (p.t.n <= 0) return;
  const unsigned blocks = (unsigned)((p.t.n + 255) / 256);
#ifndef M3D_FINISH_K
#define M3D_FINISH_K 2
#endif
  if (scene.num_shapes > 0 && scene.shape_bvh.nodes)
    finish_scene_hits_kernel<2><<<blocks, 256, 0, stream>>>(scene, p);
  else if (scene.num_shapes > 0)
    finish_scene_hits_kernel<1><<<blocks, 256, 0, stream>>>(scene, p);
  else if (M3D_FINISH_K > 1 && p.t.n >= (int64_t)256 * M3D_FINISH_K * 148)
    finish_mesh_hits_kernel<M3D_FINISH_K><<<(unsigned)((p.t.n + 256 * M3D_FINISH_K - 1) / (256 * M3D_FINISH_K)), 256, 0, stream>>>(scene, p);
  else
    finish_scene_hits_kernel<0><<<blocks, 256, 0, stream>>>(scene, p);
}

void launch_raygen_camera(const DeviceCamera &cam, int W, int row_begin, int row_end, float4 *org_tmin,
                          float4 *dir_tmax, cudaStream_t stream) {
  const int64_t n = (int64_t)W * (row_end - row_begin);
  if (n <= 0) return;
  raygen_camera_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(cam, W, row_begin, row_end, org_tmin,
                                                                         dir_tmax);
}

void launch_shade_raycast(const DeviceScene &scene, const DeviceCamera &cam, const DevicePointLight *lights,
                          int num_lights, const float4 *org_tmin, const float4 *dir_tmax,
                          const float4 *hit0, const float4 *hit1, int64_t n, float *rgb,
                          cudaStream_t stream) {
  (void)cam;
  if (n <= 0) return;
  shade_raycast_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(scene, lights, num_lights, org_tmin,
                                                                         dir_tmax, hit0, hit1, n, rgb);
}

void launch_finalize_image(const float *sum, int64_t num_values, float inv_samples, float *mean,
                           uint8_t *srgb8, cudaStream_t stream) {
  if (num_values <= 0) return;
  finalize_image_kernel<<<(unsigned)((num_values + 255) / 256), 256, 0, stream>>>(sum, num_values,
                                                                                  inv_samples, mean, srgb8);
}

}  // namespace m3d
