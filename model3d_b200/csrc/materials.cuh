// Device restatement (float32) of render3d's materials for shading kernels:
//   LambertMaterial  render3d/material.go:119-167
//   PhongMaterial    render3d/material.go:172-269 (+ sampleAroundDirection :274-335)
//   RefractMaterial  render3d/material.go:343-479
//   JoinedMaterial   render3d/material.go:554-631
//   showcase procedural variants (checker floor room.go:61-75, z-gradient vase models.go:79-97)
//
// Delta lobes.  RefractMaterial encodes Dirac lobes as windows of half-width 1e-8 with
// magnitude 2/1e-8 (material.go:401-423,441-462); 1-1e-8 is not representable in float32.
// Here a lobe is carried symbolically: BSDF = finite + delta_bsdf * (2/eps) and density =
// finite + delta_density * (2/eps) for the ONE direction that the delta sampler itself
// produced (identified by a tag, never by comparing directions); for every other direction
// the delta parts are zero, exactly as in the reference up to probability-zero events.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/m3d.h"
#include "scene.h"

namespace m3d {

struct V3f {
  float x, y, z;
};
__device__ __forceinline__ V3f v3f(float x, float y, float z) {
  V3f r;
  r.x = x;
  r.y = y;
  r.z = z;
  return r;
}
__device__ __forceinline__ V3f v3f(const float *p) { return v3f(p[0], p[1], p[2]); }
__device__ __forceinline__ V3f operator+(V3f a, V3f b) { return v3f(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3f operator-(V3f a, V3f b) { return v3f(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3f operator*(V3f a, float s) { return v3f(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3f operator*(V3f a, V3f b) { return v3f(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float dot(V3f a, V3f b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3f cross(V3f a, V3f b) {
  return v3f(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float norm(V3f a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ V3f normalize(V3f a) { return a * (1.0f / norm(a)); }
__device__ __forceinline__ bool is_zero(V3f a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f; }
__device__ __forceinline__ float sum3(V3f a) { return a.x + a.y + a.z; }

// coords.go:431-434 (n is unit already on every call site here)
__device__ __forceinline__ V3f reflect_about(V3f n, V3f c1) { return (c1 + n * (-2.f * dot(n, c1))) * -1.f; }

// coords.go:388-421
__device__ __forceinline__ void ortho_basis(V3f c, V3f &b1o, V3f &b2o) {
  const float ax = fabsf(c.x), ay = fabsf(c.y), az = fabsf(c.z);
  V3f b1 = v3f(0.f, 0.f, 0.f);
  if (ax > ay && ax > az) {
    b1.x = c.y / ax;
    b1.y = -c.x / ax;
  } else {
    const float m = ay > az ? ay : az;
    b1.y = c.z / m;
    b1.z = -c.y / m;
  }
  const V3f b2 = v3f(b1.y * c.z - b1.z * c.y, b1.z * c.x - b1.x * c.z, b1.x * c.y - b1.y * c.x);
  b1o = normalize(b1);
  b2o = normalize(b2);
}

constexpr float kCosEps = 1e-8f;  // cosineEpsilon material.go:10

// Resolved material at a hit: index + procedural diffuse colour.
struct MatAt {
  int32_t index;
  V3f diffuse;  // of the top-level material (procedural variants resolved)
};

__device__ __forceinline__ MatAt material_at(const DeviceScene &sc, int32_t object, V3f point) {
  MatAt m;
  m.index = sc.objects[object].material;
  const DeviceMaterial &d = sc.materials[m.index];
  m.diffuse = v3f(d.diffuse);
  if (d.flags & M3D_MAT_CHECKER) {
    // showcase FloorObject (room.go:66-70)
    const bool same = (int)fmodf(point.x + 300.f, 2.f) == (int)fmodf(point.y + 301.f, 2.f);
    m.diffuse = same ? v3f(d.diffuse2) : v3f(d.diffuse);
  } else if (d.flags & M3D_MAT_Z_GRADIENT) {
    // showcase VaseObject (models.go:86-90)
    const float frac = point.z / d.proc_param;
    m.diffuse = v3f(d.diffuse) * frac + v3f(d.diffuse2) * (1.f - frac);
  }
  return m;
}

__device__ __forceinline__ float maximum_cosine(float c1, float c2) {
  return fmaxf(fmaxf(fabsf(c1), fabsf(c2)), kCosEps);
}

// Finite part of the BSDF of a non-joined material (delta lobes excluded, see header).
__device__ __forceinline__ V3f simple_bsdf(const DeviceMaterial &d, V3f diffuse, V3f n, V3f src, V3f dst) {
  if (d.kind == M3D_MAT_LAMBERT) {  // material.go:125-134
    if (dot(dst, n) < 0.f || dot(src, n) > 0.f) return v3f(0.f, 0.f, 0.f);
    return diffuse * 4.f;
  }
  if (d.kind == M3D_MAT_PHONG) {  // material.go:187-216
    const float dest_dot = dot(dst, n), source_dot = -dot(src, n);
    if (dest_dot < 0.f || source_dot < 0.f) return v3f(0.f, 0.f, 0.f);
    V3f color = v3f(0.f, 0.f, 0.f);
    if (!is_zero(diffuse)) color = diffuse * 4.f;
    const V3f reflection = reflect_about(n, src) * -1.f;
    const float ref_dot = dot(reflection, dst);
    if (ref_dot < 0.f) return color;
    float intensity = powf(ref_dot, d.alpha) * (1.f + d.alpha);
    if (!(d.flags & M3D_MAT_NO_FLUX_CORRECTION)) intensity /= maximum_cosine(source_dot, dest_dot);
    return color + v3f(d.specular) * (2.f * intensity);
  }
  return v3f(0.f, 0.f, 0.f);  // RefractMaterial: delta lobes only
}

__device__ __forceinline__ V3f mat_bsdf(const DeviceScene &sc, const MatAt &m, V3f n, V3f src, V3f dst) {
  const DeviceMaterial &d = sc.materials[m.index];
  if (d.kind != M3D_MAT_JOINED) return simple_bsdf(d, m.diffuse, n, src, dst);
  V3f r = v3f(0.f, 0.f, 0.f);
  for (int i = 0; i < d.num_sub; i++) {
    const DeviceMaterial &s = sc.materials[d.sub[i]];
    r = r + simple_bsdf(s, v3f(s.diffuse), n, src, dst);
  }
  return r;
}

__device__ __forceinline__ V3f mat_emission(const DeviceScene &sc, const MatAt &m) {
  const DeviceMaterial &d = sc.materials[m.index];
  if (d.kind == M3D_MAT_REFRACT) return v3f(0.f, 0.f, 0.f);
  if (d.kind != M3D_MAT_JOINED) return v3f(d.emission);
  V3f r = v3f(0.f, 0.f, 0.f);
  for (int i = 0; i < d.num_sub; i++) {
    const DeviceMaterial &s = sc.materials[d.sub[i]];
    if (s.kind != M3D_MAT_REFRACT) r = r + v3f(s.emission);
  }
  return r;
}
__device__ __forceinline__ V3f mat_ambient(const DeviceScene &sc, const MatAt &m) {
  const DeviceMaterial &d = sc.materials[m.index];
  if (d.kind == M3D_MAT_REFRACT) return v3f(0.f, 0.f, 0.f);
  if (d.kind != M3D_MAT_JOINED) return v3f(d.ambient);
  V3f r = v3f(0.f, 0.f, 0.f);
  for (int i = 0; i < d.num_sub; i++) {
    const DeviceMaterial &s = sc.materials[d.sub[i]];
    if (s.kind != M3D_MAT_REFRACT) r = r + v3f(s.ambient);
  }
  return r;
}

// PointLight.ShadeCollision (light.go:71-92)
__device__ __forceinline__ V3f shade_collision(const DevicePointLight &l, V3f n, V3f point_to_light) {
  const float dist = norm(point_to_light);
  V3f color = v3f(l.color);
  if (l.quad_dropoff) color = color * (1.f / (dist * dist));
  const float density = 0.25f * fmaxf(0.f, dot(n, point_to_light * (1.f / dist)));
  return color * density;
}

}  // namespace m3d
