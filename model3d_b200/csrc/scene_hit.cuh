// Device functions shared by the scene-level kernels: analytic shapes in float64
// (model3d/shapes.go:35-93,177-247,601-705,816-856), the float64 re-evaluation of the
// winning triangle (model3d/primitives.go:27-33,207-249,508-516) and JoinedObject's
// closest-of rule (render3d/object.go:141-153).
#pragma once
#include "scene.h"
#include "trace_core.cuh"

namespace m3d {

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
static __device__ __noinline__ bool sphere_hit(const DeviceShape &s, D3 o, D3 d, double t_floor, double &t, D3 &n) {
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
static __device__ __noinline__ void ray_bounds(D3 o, D3 d, D3 mn, D3 mx, double &min_frac, double &max_frac) {
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
static __device__ __noinline__ bool rect_hit(const DeviceShape &s, D3 o, D3 d, double t_floor, double &t, D3 &n) {
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
static __device__ __noinline__ bool circle_hit(D3 normal, D3 center, double radius, D3 o, D3 d, double t_floor, double &t) {
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
static __device__ __noinline__ bool cylinder_hit(const DeviceShape &s, D3 o_in, D3 d, double t_floor, double &t, D3 &n) {
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

// The float64 tests behind a float32 interface, out of line: the Monte-Carlo renderers (statistical
// parity) keep no float64 state across their shape loop, which is what had their resolve kernels at the
// 128-register limit.
static __device__ __noinline__ bool shape_hit_from_f32(const DeviceShape &sh, float4 o, float4 d, float t_floor,
                                                       float &t_out, float &nx, float &ny, float &nz) {
  const D3 od = d3(o.x, o.y, o.z), dd = d3(d.x, d.y, d.z);
  double t;
  D3 n;
  bool ok = false;
  if (sh.kind == SHAPE_SPHERE) ok = sphere_hit(sh, od, dd, (double)t_floor, t, n);
  else if (sh.kind == SHAPE_RECT) ok = rect_hit(sh, od, dd, (double)t_floor, t, n);
  else if (sh.kind == SHAPE_CYLINDER) ok = cylinder_hit(sh, od, dd, (double)t_floor, t, n);
  if (!ok) return false;
  t_out = (float)t;
  nx = (float)n.x;
  ny = (float)n.y;
  nz = (float)n.z;
  return true;
}

// Conservative float32 pre-test: true only if the ray's supporting line provably misses the
// shape's bounding sphere (sphere: the shape itself).  Margin 1e-4 relative covers float32
// rounding of the discriminant with two orders of magnitude to spare.
__device__ __forceinline__ bool shape_certainly_missed(const DeviceShape &sh, float4 o, float4 d) {
  float cx, cy, cz, r;
  if (sh.kind == SHAPE_SPHERE) {
    cx = (float)sh.p0[0];
    cy = (float)sh.p0[1];
    cz = (float)sh.p0[2];
    r = (float)sh.radius;
  } else {
    // rect: centre and half diagonal; cylinder: segment midpoint, half length + radius
    cx = 0.5f * (float)(sh.p0[0] + sh.p1[0]);
    cy = 0.5f * (float)(sh.p0[1] + sh.p1[1]);
    cz = 0.5f * (float)(sh.p0[2] + sh.p1[2]);
    const float hx = 0.5f * (float)(sh.p1[0] - sh.p0[0]), hy = 0.5f * (float)(sh.p1[1] - sh.p0[1]),
                hz = 0.5f * (float)(sh.p1[2] - sh.p0[2]);
    r = sqrtf(hx * hx + hy * hy + hz * hz) + (sh.kind == SHAPE_CYLINDER ? (float)sh.radius : 0.f);
  }
  r *= 1.001f;
  const float ox = o.x - cx, oy = o.y - cy, oz = o.z - cz;
  const float a = d.x * d.x + d.y * d.y + d.z * d.z;
  const float b = ox * d.x + oy * d.y + oz * d.z;
  const float oo = ox * ox + oy * oy + oz * oz;
  const float c = oo - r * r;
  // discriminant / 4 of a t^2 + 2 b t + c; every term is bounded by a * oo
  const float disc = b * b - a * c;
  const float margin = 1e-4f * (a * oo + a * r * r);
  if (disc < -margin) return true;
  // both roots behind the origin: outside the sphere and pointing away
  if (c > margin && b > 0.f && b * b > 1e-4f * a * oo) return true;
  return false;
}

// Sphere.FirstRayCollision (shapes.go:35-93) in float32 for the Monte-Carlo renderers (their
// parity contract is statistical).  The discriminant is formed as a*(r^2 - |oc - (b/a) d|^2),
// which does not cancel for distant origins.  Returns 0 miss, 1 hit, 2 undecided (grazing ray or
// a root next to t_floor): the caller falls back to the float64 test.
__device__ __forceinline__ int sphere_hit_f32(const DeviceShape &sh, float4 o, float4 d, float t_floor, float &t,
                                              float &nx, float &ny, float &nz) {
  const float cx = (float)sh.p0[0], cy = (float)sh.p0[1], cz = (float)sh.p0[2], r = (float)sh.radius;
  const float ox = o.x - cx, oy = o.y - cy, oz = o.z - cz;
  const float a = d.x * d.x + d.y * d.y + d.z * d.z;
  const float b = ox * d.x + oy * d.y + oz * d.z;
  const float ba = b / a;
  const float lx = ox - ba * d.x, ly = oy - ba * d.y, lz = oz - ba * d.z;
  const float r2 = r * r;
  const float q = r2 - (lx * lx + ly * ly + lz * lz);  // discriminant / (4 a)
  if (fabsf(q) < 1e-4f * r2) return 2;
  if (q < 0.f) return 0;
  const float sq = sqrtf(q / a);
  const float t1 = -ba - sq, t2 = -ba + sq;
  const float slack = 1e-5f * (fabsf(ba) + sq) + 1e-5f * t_floor;
  float tt;
  if (fabsf(t1 - t_floor) < slack || fabsf(t2 - t_floor) < slack) return 2;
  if (t1 >= t_floor)
    tt = t1;
  else if (t2 >= t_floor)
    tt = t2;
  else
    return 0;
  t = tt;
  const float px = ox + d.x * tt, py = oy + d.y * tt, pz = oz + d.z * tt;
  const float inv = rsqrtf(px * px + py * py + pz * pz);
  nx = px * inv;
  ny = py * inv;
  nz = pz * inv;
  return 1;
}

// Rect.FirstRayCollision (shapes.go:177-247) in float32 for the Monte-Carlo renderers: the slab test of
// bvh.go:322-351 with the face taken from the binding axis.  Returns 0 miss, 1 hit, 2 undecided (a
// decision within the float32 error of its threshold: grazing an edge, a root next to t_floor, the
// origin on a face plane of a parallel ray): the caller falls back to the float64 test.
__device__ __forceinline__ int rect_hit_f32(const DeviceShape &sh, float4 o, float4 d, float t_floor, float &t,
                                            float &nx, float &ny, float &nz) {
  const float oo[3] = {o.x, o.y, o.z}, dv[3] = {d.x, d.y, d.z};
  float tmin = -INFINITY, tmax = INFINITY, emin = 0.f, emax = 0.f;
  int amin = 0, amax = 0;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const float mn = (float)sh.p0[a], mx = (float)sh.p1[a];
    const float perr = 2e-7f * (fabsf(mn) + fabsf(mx) + fabsf(oo[a]));  // rounding of mn - o, mx - o and of mn, mx
    if (dv[a] == 0.f) {
      if (fabsf(oo[a] - mn) <= perr || fabsf(oo[a] - mx) <= perr) return 2;
      if (oo[a] < mn || oo[a] > mx) return 0;
      continue;
    }
    const float inv = 1.0f / dv[a];
    float t1 = (mn - oo[a]) * inv, t2 = (mx - oo[a]) * inv;
    if (t1 > t2) {
      const float tmp = t1;
      t1 = t2;
      t2 = tmp;
    }
    const float e = perr * fabsf(inv) + 1e-6f * (fabsf(t1) + fabsf(t2));
    if (t2 < -e) return 0;  // bvh.go:341-343
    if (t2 <= e) return 2;
    if (t1 > tmin) {
      tmin = t1;
      emin = e;
      amin = a;
    }
    if (t2 < tmax) {
      tmax = t2;
      emax = e;
      amax = a;
    }
  }
  if (tmax < tmin - (emin + emax) || tmax < t_floor - emax) return 0;
  if (tmax <= tmin + (emin + emax) || tmax <= t_floor + emax) return 2;
  if (fabsf(tmin - t_floor) <= emin) return 2;
  const bool enter = tmin >= t_floor;
  t = enter ? tmin : tmax;
  const int axis = enter ? amin : amax;
  // the face: entering through the low side when moving up the axis, leaving through the high side
  const float d_axis = axis == 0 ? d.x : (axis == 1 ? d.y : d.z);
  const float sign = (d_axis > 0.f) == enter ? -1.f : 1.f;
  // a second face within the error of the hit point (edge / corner): the float64 rule picks by distance
#pragma unroll
  for (int a = 0; a < 3; a++) {
    if (a == axis) continue;
    const float c = oo[a] + dv[a] * t;
    const float mn = (float)sh.p0[a], mx = (float)sh.p1[a];
    const float tol = 4e-6f * (fabsf(mn) + fabsf(mx) + fabsf(oo[a]) + fabsf(dv[a] * t));
    if (fabsf(c - mn) <= tol || fabsf(c - mx) <= tol) return 2;
  }
  nx = axis == 0 ? sign : 0.f;
  ny = axis == 1 ? sign : 0.f;
  nz = axis == 2 ? sign : 0.f;
  return 1;
}

// Result of resolving one scene ray: the closest of the BVH's raw triangle hit and the
// analytic shapes.  surf: leaf-order triangle index, or -2-shape index, or -1 (miss).
struct SceneHit {
  float t;
  float b1, b2;
  float nx, ny, nz;
  int32_t prim, obj, surf;
};

// o = (origin, tmin), d = (direction, tmax); raw = trace_first_hit_kernel's output
// (t_f32, -, -, bits(leaf-order triangle | -1)); skip = surface the ray starts on.
// SHAPES = 0 compiles the analytic-shape part out (mesh colliders: a third of the registers);
// 1: every shape is tested (a handful of shapes); 2: scenes with an object-level hierarchy
// (DeviceScene::shape_bvh) walk it instead; 3: like 1 for scenes whose shapes are all spheres
// (DeviceScene::spheres_only; Monte-Carlo renderers only: no rect / cylinder code in the kernel).
template <int SHAPES = 1>
__device__ inline SceneHit resolve_scene_hit(const DeviceScene &sc, float4 o, float4 d, float4 raw, int skip,
                                             bool refine) {
  SceneHit h;
  h.t = 0.f;
  h.b1 = h.b2 = 0.f;
  h.nx = h.ny = h.nz = 0.f;
  h.prim = -1;
  h.obj = -1;
  h.surf = -1;
  const int tri_idx = __float_as_int(raw.w);
  double best_t = INFINITY;
  int best_obj = 0x7fffffff;
  if (tri_idx >= 0) {
    const float4 *tri = sc.bvh.tris + (size_t)tri_idx * 3;
    const int prim = __float_as_int(__ldg(&tri[0].w));
    const int obj = __float_as_int(__ldg(&tri[1].w));
    h.surf = tri_idx;
    h.prim = prim;
    h.obj = obj;
    if (refine) {
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
      h.t = (float)t;
      h.b1 = (float)r.b1;
      h.b2 = (float)r.b2;
      h.nx = (float)nx;
      h.ny = (float)ny;
      h.nz = (float)nz;
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
      h.t = raw.x;
      h.b1 = fb1;
      h.b2 = fb2;
      h.nx = nx * s;
      h.ny = ny * s;
      h.nz = nz * s;
    }
    best_obj = obj;
  }
  if ((SHAPES == 1 || SHAPES == 3) && !refine && sc.num_shapes > 0) {
    // Monte-Carlo renderers, a handful of shapes: float32 state only (spheres by their float32 test, the
    // float64 tests out of line behind the bounding-sphere rejection)
    float best_tf = tri_idx >= 0 ? raw.x : INFINITY;
    const float inv_len_f = rsqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
    for (int s = 0; s < sc.num_shapes; s++) {
      const DeviceShape &sh = sc.shapes[s];
      float t_floor = o.w;
      if (skip == -2 - s) {
        float size = (float)sh.radius;
        if (SHAPES != 3 && sh.kind == SHAPE_RECT)
          size = fmaxf(fmaxf((float)(sh.p1[0] - sh.p0[0]), (float)(sh.p1[1] - sh.p0[1])), (float)(sh.p1[2] - sh.p0[2]));
        t_floor = fmaxf(t_floor, 1e-4f * size * inv_len_f);
      }
      float tf, fx, fy, fz;
      int fast = 2;
      if (sh.kind == SHAPE_SPHERE) {
        fast = sphere_hit_f32(sh, o, d, t_floor, tf, fx, fy, fz);
        if (fast == 0) continue;
      } else if (SHAPES == 3) {
        continue;  // not reached: spheres only
      } else if (sh.kind == SHAPE_RECT) {
        fast = rect_hit_f32(sh, o, d, t_floor, tf, fx, fy, fz);
        if (fast == 0) continue;
      } else if (shape_certainly_missed(sh, o, d)) {
        continue;
      }
      if (fast == 2) {
        if (SHAPES == 3) {  // a grazing sphere: the float64 sphere test alone
          double td;
          D3 nd;
          if (!sphere_hit(sh, d3(o.x, o.y, o.z), d3(d.x, d.y, d.z), (double)t_floor, td, nd)) continue;
          tf = (float)td;
          fx = (float)nd.x;
          fy = (float)nd.y;
          fz = (float)nd.z;
        } else if (!shape_hit_from_f32(sh, o, d, t_floor, tf, fx, fy, fz)) {
          continue;
        }
      }
      if (tf > d.w) continue;  // ray tmax (shadow / visibility rays)
      // JoinedObject.Cast: strict '<' in object order, the first object wins ties
      if (tf < best_tf || (tf == best_tf && sh.object < best_obj)) {
        best_tf = tf;
        best_obj = sh.object;
        h.surf = -2 - s;
        h.t = tf;
        h.b1 = h.b2 = 0.f;
        h.prim = 0;
        h.obj = sh.object;
        h.nx = fx;
        h.ny = fy;
        h.nz = fz;
      }
    }
  } else if (SHAPES && sc.num_shapes > 0) {
    const D3 od = d3(o.x, o.y, o.z), dd = d3(d.x, d.y, d.z);
    const double t_hi = (double)d.w;  // ray tmax (shadow / visibility rays)
    const double inv_len = 1.0 / dnorm(dd);
    // one analytic shape / instance against the ray; keeps the closest hit in h / best_t / best_obj
    auto test_shape = [&](int s) {
      const DeviceShape &sh = sc.shapes[s];
      double t_floor = (double)o.w;
      if (skip == -2 - s) {
        double size = sh.radius;
        if (sh.kind == SHAPE_RECT || sh.kind == SHAPE_INSTANCE)
          size = fmax(fmax(sh.p1[0] - sh.p0[0], sh.p1[1] - sh.p0[1]), sh.p1[2] - sh.p0[2]);
        t_floor = fmax(t_floor, 1e-4 * size * inv_len);
      }
      // (instances only exist in scenes with an object-level hierarchy, SHAPES == 2: the kernels of the
      // few-shapes scenes stay free of the nested traversal's stack and registers)
      if (SHAPES == 2 && sh.kind == SHAPE_INSTANCE) {
        // render3d.Translate / MatrixMultiply of a shared collider (transform.go:26-31,76-85): the ray
        // goes to object space (direction NOT re-normalised, so t is the same parameter), the mesh's own
        // hierarchy is walked by this thread, the normal comes back through the forward matrix
        if (shape_certainly_missed(sh, o, d)) return;  // bounding sphere of the world bounds
        const DeviceInstance &in = sc.instances[sh.instance];
        const float px = o.x - in.off[0], py = o.y - in.off[1], pz = o.z - in.off[2];
        RayF r;
        r.ox = in.inv[0] * px + in.inv[1] * py + in.inv[2] * pz;
        r.oy = in.inv[3] * px + in.inv[4] * py + in.inv[5] * pz;
        r.oz = in.inv[6] * px + in.inv[7] * py + in.inv[8] * pz;
        r.dx = in.inv[0] * d.x + in.inv[1] * d.y + in.inv[2] * d.z;
        r.dy = in.inv[3] * d.x + in.inv[4] * d.y + in.inv[5] * d.z;
        r.dz = in.inv[6] * d.x + in.inv[7] * d.y + in.inv[8] * d.z;
        r.tmin = (float)t_floor;
        r.tmax = (float)fmin(fmin(best_t, t_hi) * 1.000001, 3.0e38);
        HitF hf;
        trace_bvh<false, false>(in.blas.nodes, in.blas.tris, in.blas.bmin, in.blas.bmax, r, -1, hf, nullptr);
        if (hf.tri < 0) return;
        const float4 *tri = in.blas.tris + (size_t)hf.tri * 3;
        double t = hf.t, b1 = hf.b1, b2 = hf.b2, nx, ny, nz;
        if (refine) {
          const HitD rr = refine_hit_f64(tri, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz);
          if (rr.t >= 0.0) t = rr.t;
          b1 = rr.b1;
          b2 = rr.b2;
          nx = rr.nx;
          ny = rr.ny;
          nz = rr.nz;
        } else {
          const float4 q0 = __ldg(tri), q1 = __ldg(tri + 1), q2 = __ldg(tri + 2);
          const float e1x = q1.x - q0.x, e1y = q1.y - q0.y, e1z = q1.z - q0.z;
          const float e2x = q2.x - q0.x, e2y = q2.y - q0.y, e2z = q2.z - q0.z;
          nx = e1y * e2z - e1z * e2y;
          ny = e1z * e2x - e1x * e2z;
          nz = e1x * e2y - e1y * e2x;
        }
        if (in.blas.vnormals) {  // InterpNormalTriangle.InterpNormal (primitives.go:508-516)
          const float4 *vn = in.blas.vnormals + (size_t)hf.tri * 3;
          const float4 a = __ldg(vn), b = __ldg(vn + 1), c = __ldg(vn + 2);
          const double b0 = 1.0 - (b1 + b2);
          nx = b0 * a.x + b1 * b.x + b2 * c.x;
          ny = b0 * a.y + b1 * b.y + b2 * c.y;
          nz = b0 * a.z + b1 * b.z + b2 * c.z;
        }
        if (t > t_hi || t < t_floor) return;
        if (t < best_t || (t == best_t && sh.object < best_obj)) {
          const double wx = in.fwd[0] * nx + in.fwd[1] * ny + in.fwd[2] * nz;
          const double wy = in.fwd[3] * nx + in.fwd[4] * ny + in.fwd[5] * nz;
          const double wz = in.fwd[6] * nx + in.fwd[7] * ny + in.fwd[8] * nz;
          const double inv_n = 1.0 / sqrt(wx * wx + wy * wy + wz * wz);
          best_t = t;
          best_obj = sh.object;
          h.surf = -2 - s;
          h.t = (float)t;
          h.b1 = (float)b1;
          h.b2 = (float)b2;
          h.prim = __float_as_int(__ldg(&tri[0].w));
          h.obj = sh.object;
          h.nx = (float)(wx * inv_n);
          h.ny = (float)(wy * inv_n);
          h.nz = (float)(wz * inv_n);
        }
        return;
      }
      // cheap float32 rejection with a generous error margin: most rays miss most shapes, and
      // the float64 tests (sqrt / divisions) are an order of magnitude more instructions
      // (Monte-Carlo spheres go straight to their float32 test, which all lanes run together)
      if ((refine || sh.kind != SHAPE_SPHERE) && shape_certainly_missed(sh, o, d)) return;
      double t;
      D3 n;
      bool ok = false;
      int fast = 2;
      if (!refine && sh.kind == SHAPE_SPHERE) {
        float tf, fx, fy, fz;
        fast = sphere_hit_f32(sh, o, d, (float)t_floor, tf, fx, fy, fz);
        if (fast == 0) return;
        if (fast == 1) {
          ok = true;
          t = (double)tf;
          n = d3(fx, fy, fz);
        }
      }
      if (fast != 2) {
      } else if (sh.kind == SHAPE_SPHERE) ok = sphere_hit(sh, od, dd, t_floor, t, n);
      else if (sh.kind == SHAPE_RECT) ok = rect_hit(sh, od, dd, t_floor, t, n);
      else if (sh.kind == SHAPE_CYLINDER) ok = cylinder_hit(sh, od, dd, t_floor, t, n);
      if (!ok || t > t_hi) return;
      // JoinedObject.Cast: strict '<' in object order, the first object wins ties
      if (t < best_t || (t == best_t && sh.object < best_obj)) {
        best_t = t;
        best_obj = sh.object;
        h.surf = -2 - s;
        h.t = (float)t;
        h.b1 = h.b2 = 0.f;
        h.prim = 0;
        h.obj = sh.object;
        h.nx = (float)n.x;
        h.ny = (float)n.y;
        h.nz = (float)n.z;
      }
    };
    if (SHAPES == 2 && sc.shape_bvh.nodes) {
      // BVHToObject (object.go:172-185): walk the hierarchy over the shapes' bounds, near to far,
      // pruned by the closest hit so far (the triangle hit of the mesh BVH included)
      RayF r;
      r.ox = o.x; r.oy = o.y; r.oz = o.z; r.tmin = o.w;
      r.dx = d.x; r.dy = d.y; r.dz = d.z;
      r.tmax = (float)fmin(fmin(best_t, t_hi) * 1.000001, 3.0e38);
      const float4 *proxies = sc.shape_bvh.tris;
      walk_bvh_leaves(sc.shape_bvh.nodes, sc.shape_bvh.bmin, sc.shape_bvh.bmax, r, [&](int32_t leaf, float tmax) {
        test_shape(__float_as_int(__ldg(&proxies[(size_t)leaf * 3].w)));
        const float nt = (float)fmin(fmin(best_t, t_hi) * 1.000001, 3.0e38);
        return nt < tmax ? nt : tmax;
      });
    } else {
      for (int s = 0; s < sc.num_shapes; s++) test_shape(s);
    }
  }
  if (sc.objects && h.obj >= 0 && (sc.objects[h.obj].flags & M3D_OBJ_FLIP_NORMAL)) {
    h.nx = -h.nx;  // showcase DomeObject (room.go:40-44)
    h.ny = -h.ny;
    h.nz = -h.nz;
  }
  return h;
}

}  // namespace m3d
