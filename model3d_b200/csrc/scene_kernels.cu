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
#include "trace_core.cuh"

namespace m3d {

namespace {

struct D3 {
  double x, y, z;
};
__device__ __forceinline__ D3 d3(double x, double y, double z) {
  D3 r;
  r.x = x;
  r.y = y;
  r.z = z;
  return r;
}
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return d3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return d3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ D3 operator*(D3 a, double s) { return d3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ double ddot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ double dnorm(D3 a) { return sqrt(ddot(a, a)); }
__device__ __forceinline__ D3 dnormalize(D3 a) { return a * (1.0 / dnorm(a)); }
__device__ __forceinline__ double comp(D3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// Every analytic test takes t_floor: hits with t < t_floor are ignored (0 for primary rays;
// a small positive value when the ray starts on this very shape, standing in for the
// reference's 1e-8 origin offset which float32 origins cannot express).

// shapes.go:52-93 (first root >= 0; discriminant <= 0 misses; outward normal)
__device__ bool sphere_hit(const DeviceShape &s, D3 o, D3 d, double t_floor, double &t, D3 &n) {
  const D3 c = d3(s.p0[0], s.p0[1], s.p0[2]);
  const D3 oc = o - c;
  const double a = ddot(d, d), b = 2 * ddot(d, oc), cc = ddot(oc, oc) - s.radius * s.radius;
  const double disc = b * b - 4 * a * cc;
  if (disc <= 0) return false;
  const double sq = sqrt(disc);
  double t1 = (-b + sq) / (2 * a), t2 = (-b - sq) / (2 * a);
  if (t1 > t2) {
    const double tmp = t1;
    t1 = t2;
    t2 = tmp;
  }
  double tt;
  if (t1 >= t_floor)
    tt = t1;
  else if (t2 >= t_floor)
    tt = t2;
  else
    return false;
  t = tt;
  n = dnormalize((o + d * tt) - c);
  return true;
}

// bvh.go:322-351
__device__ void ray_bounds(D3 o, D3 d, D3 mn, D3 mx, double &min_frac, double &max_frac) {
  min_frac = -INFINITY;
  max_frac = INFINITY;
  for (int axis = 0; axis < 3; axis++) {
    const double origin = comp(o, axis), rate = comp(d, axis);
    if (rate == 0) {
      if (origin < comp(mn, axis) || origin > comp(mx, axis)) {
        min_frac = 0;
        max_frac = -1;
        return;
      }
      continue;
    }
    double t1 = (comp(mn, axis) - origin) / rate, t2 = (comp(mx, axis) - origin) / rate;
    if (t1 > t2) {
      const double tmp = t1;
      t1 = t2;
      t2 = tmp;
    }
    if (t2 < 0) {
      min_frac = 0;
      max_frac = -1;
      return;
    }
    if (t1 > min_frac) min_frac = t1;
    if (t2 < max_frac) max_frac = t2;
  }
}

// shapes.go:177-196, 221-247
__device__ bool rect_hit(const DeviceShape &s, D3 o, D3 d, double t_floor, double &t, D3 &n) {
  const D3 mn = d3(s.p0[0], s.p0[1], s.p0[2]), mx = d3(s.p1[0], s.p1[1], s.p1[2]);
  double tmin, tmax;
  ray_bounds(o, d, mn, mx, tmin, tmax);
  if (tmax < tmin || tmax < t_floor) return false;
  double tt = tmin;
  if (tt < t_floor) tt = tmax;
  t = tt;
  const D3 c = o + d * tt;
  int axis = 0;
  double sign = 0, min_dist = INFINITY;
  for (int i = 0; i < 3; i++) {
    double dd = fabs(comp(c, i) - comp(mn, i));
    if (dd < min_dist) {
      min_dist = dd;
      sign = -1;
      axis = i;
    }
    dd = fabs(comp(c, i) - comp(mx, i));
    if (dd < min_dist) {
      min_dist = dd;
      sign = 1;
      axis = i;
    }
  }
  n = d3(axis == 0 ? sign : 0.0, axis == 1 ? sign : 0.0, axis == 2 ? sign : 0.0);
  return true;
}

// shapes.go:816-856
__device__ bool circle_hit(D3 normal, D3 center, double radius, D3 o, D3 d, double t_floor, double &t) {
  const double ddn = ddot(d, normal);
  if (fabs(ddn) < 1e-8 * dnorm(d) * dnorm(normal)) return false;
  const double tt = (ddot(normal, center) - ddot(o, normal)) / ddn;
  if (tt < t_floor) return false;
  const D3 p = o + d * tt;
  if (dnorm(p - center) > radius) return false;
  t = tt;
  return true;
}

// shapes.go:601-705: minimum over side roots and the two caps (first strictly smaller wins)
__device__ bool cylinder_hit(const DeviceShape &s, D3 o_in, D3 d, double t_floor, double &t, D3 &n) {
  const D3 p1 = d3(s.p0[0], s.p0[1], s.p0[2]), p2 = d3(s.p1[0], s.p1[1], s.p1[2]);
  bool ok = false;
  const D3 v = dnormalize(p2 - p1);
  const D3 o = o_in - p1;
  const D3 v1 = v * ddot(o, v) - o;
  const D3 v2 = v * ddot(d, v) - d;
  const double a = ddot(v2, v2), b = 2 * ddot(v1, v2), cv = ddot(v1, v1) - s.radius * s.radius;
  const double disc = b * b - 4 * a * cv;
  if (disc > 0) {
    const double sq = sqrt(disc);
    const double max_scale = dnorm(p2 - p1);
    for (int k = 0; k < 2; k++) {
      const double sign = k == 0 ? -1.0 : 1.0;
      const double tt = (-b + sign * sq) / (2 * a);
      if (tt < t_floor) continue;
      const D3 p = o + d * tt;
      const double frac = ddot(v, p);
      if (frac >= 0 && frac < max_scale && (!ok || tt < t)) {
        t = tt;
        n = dnormalize(p - v * frac);
        ok = true;
      }
    }
  }
  for (int i = 0; i < 2; i++) {
    const D3 tip = i == 0 ? p1 : p2;
    const D3 nn = i == 0 ? v * -1.0 : v;
    double tt;
    if (circle_hit(nn, tip, s.radius, o_in, d, t_floor, tt) && (!ok || tt < t)) {
      t = tt;
      n = nn;
      ok = true;
    }
  }
  return ok;
}

__global__ void __launch_bounds__(256)
finish_scene_hits_kernel(DeviceScene sc, SceneTraceLaunch sp) {
  const TraceLaunch &p = sp.t;
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= p.n) return;
  const float4 raw = p.hit0[i];
  const int tri_idx = __float_as_int(raw.w);
  float4 h0 = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
  float4 h1 = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
  int surf = -1;  // leaf-order triangle index, or -2-shape, or -1
  const bool need_ray = tri_idx >= 0 || sc.num_shapes > 0;
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f), d = make_float4(0.f, 0.f, 1.f, 0.f);
  if (need_ray) {
    o = __ldcs(p.org_tmin + i);
    d = __ldcs(p.dir_tmax + i);
  }
  double best_t = INFINITY;
  int best_obj = 0x7fffffff;
  if (tri_idx >= 0) {
    const float4 *tri = sc.bvh.tris + (size_t)tri_idx * 3;
    const int prim = __float_as_int(__ldg(&tri[0].w));
    const int obj = __float_as_int(__ldg(&tri[1].w));
    surf = tri_idx;
    if (p.refine) {
      const HitD r = refine_hit_f64(tri, o.x, o.y, o.z, d.x, d.y, d.z);
      const double t = r.t >= 0.0 ? r.t : (double)raw.x;
      double nx = r.nx, ny = r.ny, nz = r.nz;
      if (sc.bvh.vnormals) {
        // InterpNormalTriangle.InterpNormal (primitives.go:508-516)
        const float4 *vn = sc.bvh.vnormals + (size_t)tri_idx * 3;
        const float4 a = __ldg(vn), b = __ldg(vn + 1), c = __ldg(vn + 2);
        nx = r.b0 * a.x + r.b1 * b.x + r.b2 * c.x;
        ny = r.b0 * a.y + r.b1 * b.y + r.b2 * c.y;
        nz = r.b0 * a.z + r.b1 * b.z + r.b2 * c.z;
        const double s = 1.0 / sqrt(nx * nx + ny * ny + nz * nz);
        nx *= s;
        ny *= s;
        nz *= s;
      }
      best_t = t;
      h0 = make_float4((float)t, (float)r.b1, (float)r.b2, __int_as_float(prim));
      h1 = make_float4((float)nx, (float)ny, (float)nz, __int_as_float(obj));
    } else {
      const float4 q0 = __ldg(tri), q1 = __ldg(tri + 1), q2 = __ldg(tri + 2);
      const float e1x = q1.x - q0.x, e1y = q1.y - q0.y, e1z = q1.z - q0.z;
      const float e2x = q2.x - q0.x, e2y = q2.y - q0.y, e2z = q2.z - q0.z;
      float nx = e1y * e2z - e1z * e2y, ny = e1z * e2x - e1x * e2z, nz = e1x * e2y - e1y * e2x;
      // float32 barycentrics (Moeller-Trumbore, as in the traversal)
      const float c1x = d.y * e2z - d.z * e2y, c1y = d.z * e2x - d.x * e2z, c1z = d.x * e2y - d.y * e2x;
      const float inv = 1.0f / (c1x * e1x + c1y * e1y + c1z * e1z);
      const float px = o.x - q0.x, py = o.y - q0.y, pz = o.z - q0.z;
      const float fb1 = inv * (px * c1x + py * c1y + pz * c1z);
      const float fb2 =
          inv * (d.x * (py * e1z - pz * e1y) + d.y * (pz * e1x - px * e1z) + d.z * (px * e1y - py * e1x));
      if (sc.bvh.vnormals) {
        const float4 *vn = sc.bvh.vnormals + (size_t)tri_idx * 3;
        const float4 a = __ldg(vn), b = __ldg(vn + 1), c = __ldg(vn + 2);
        const float b0 = 1.f - (fb1 + fb2);
        nx = b0 * a.x + fb1 * b.x + fb2 * c.x;
        ny = b0 * a.y + fb1 * b.y + fb2 * c.y;
        nz = b0 * a.z + fb1 * b.z + fb2 * c.z;
      }
      const float s = rsqrtf(nx * nx + ny * ny + nz * nz);
      best_t = raw.x;
      h0 = make_float4(raw.x, fb1, fb2, __int_as_float(prim));
      h1 = make_float4(nx * s, ny * s, nz * s, __int_as_float(obj));
    }
    best_obj = obj;
  }
  if (sc.num_shapes > 0) {
    const D3 od = d3(o.x, o.y, o.z), dd = d3(d.x, d.y, d.z);
    const int skip = sp.skip_ids ? sp.skip_ids[i] : -1;
    const double t_hi = (double)d.w;  // ray tmax (shadow / visibility rays)
    const double inv_len = 1.0 / dnorm(dd);
    for (int s = 0; s < sc.num_shapes; s++) {
      const DeviceShape &sh = sc.shapes[s];
      double t_floor = (double)o.w;
      if (skip == -2 - s) {
        double size = sh.radius;
        if (sh.kind == SHAPE_RECT)
          size = fmax(fmax(sh.p1[0] - sh.p0[0], sh.p1[1] - sh.p0[1]), sh.p1[2] - sh.p0[2]);
        t_floor = fmax(t_floor, 1e-4 * size * inv_len);
      }
      double t;
      D3 n;
      bool ok = false;
      if (sh.kind == SHAPE_SPHERE) ok = sphere_hit(sh, od, dd, t_floor, t, n);
      else if (sh.kind == SHAPE_RECT) ok = rect_hit(sh, od, dd, t_floor, t, n);
      else if (sh.kind == SHAPE_CYLINDER) ok = cylinder_hit(sh, od, dd, t_floor, t, n);
      if (!ok || t > t_hi) continue;
      // JoinedObject.Cast: strict '<' in object order, the first object wins ties
      if (t < best_t || (t == best_t && sh.object < best_obj)) {
        best_t = t;
        best_obj = sh.object;
        surf = -2 - s;
        h0 = make_float4((float)t, 0.f, 0.f, __int_as_float(0));
        h1 = make_float4((float)n.x, (float)n.y, (float)n.z, __int_as_float(sh.object));
      }
    }
  }
  if (sc.objects && best_obj != 0x7fffffff && (sc.objects[best_obj].flags & M3D_OBJ_FLIP_NORMAL)) {
    h1.x = -h1.x;  // showcase DomeObject (room.go:40-44)
    h1.y = -h1.y;
    h1.z = -h1.z;
  }
  __stcs(p.hit0 + i, h0);
  __stcs(p.hit1 + i, h1);
  if (sp.surf_ids) sp.surf_ids[i] = surf;
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

}  // namespace

void launch_finish_scene_hits(const DeviceScene &scene, const SceneTraceLaunch &p, cudaStream_t stream) {
  if (p.t.n <= 0) return;
  finish_scene_hits_kernel<<<(unsigned)((p.t.n + 255) / 256), 256, 0, stream>>>(scene, p);
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
