// ORACLE -- TEST INFRASTRUCTURE ONLY (see vec.hpp header).
//
// float64 CPU restatement of render3d's scene model:
//   Object / ColliderObject / JoinedObject   render3d/object.go:12-46,118-153
//   Translate / MatrixMultiply               render3d/transform.go:6-85
//   Camera / NewCameraAt / Caster / axes     render3d/camera.go:48-113
//   PointLight.ShadeCollision                render3d/light.go:57-92
//   Materials (Lambert/Phong/Refract/Joined) render3d/material.go:97-479,554-631
//   showcase procedural objects (Dome/Floor/Vase)
//                       examples/renderings/showcase/room.go:36-75, models.go:79-97
//   RayCaster.Render                         render3d/raycast.go:15-39
//   sRGB + 8-bit                             render3d/light.go:41-55, image.go:125-145
// Descriptors are the shared PODs of include/m3d.h.
#pragma once
#include <cstring>
#include <functional>
#include <random>
#include <thread>

#include "../include/m3d.h"
#include "collide.hpp"

namespace orc {

inline V3 v3(const double *p) { return {p[0], p[1], p[2]}; }

// Per-thread RNG standing in for Go's math/rand (*rand.Rand).  The reference's
// generator is an ALFG seeded from an unseeded global source (concurrency.go:
// 33-35), so bit-level RNG parity is neither possible nor required.
struct Rng {
  std::mt19937_64 eng;
  explicit Rng(uint64_t seed) : eng(seed) {}
  double f64() { return (eng() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
  int intn(int n) { return int(eng() % (uint64_t)n); }
  double normal() {
    std::normal_distribution<double> d(0.0, 1.0);
    return d(eng);
  }
};

// ---- materials ---------------------------------------------------------------
constexpr double kCosineEpsilon = 1e-8;  // material.go:10

struct Material {
  m3d_material_desc d;
};

struct MaterialTable {
  std::vector<m3d_material_desc> mats;
};

// A material instance at a hit: index + (for procedural variants) resolved colour.
struct MatRef {
  const MaterialTable *tab = nullptr;
  int32_t index = -1;
  V3 diffuse_override;
  bool has_override = false;
};

inline V3 mat_diffuse(const MatRef &m, const m3d_material_desc &d) {
  return m.has_override ? m.diffuse_override : v3(d.diffuse);
}

// forward decls (Joined recursion)
inline V3 mat_bsdf(const MatRef &m, int32_t idx, V3 normal, V3 source, V3 dest);
inline V3 mat_sample_source(const MatRef &m, int32_t idx, Rng &g, V3 normal, V3 dest);
inline double mat_source_density(const MatRef &m, int32_t idx, V3 normal, V3 source, V3 dest);
inline V3 mat_sample_dest(const MatRef &m, int32_t idx, Rng &g, V3 normal, V3 source);
inline double mat_dest_density(const MatRef &m, int32_t idx, V3 normal, V3 source, V3 dest);

// material.go:153-159
inline double lambert_density(V3 normal, V3 source) {
  double nd = -dot(normal, source);
  if (nd < 0) return 0;
  return 4 * nd;
}
// material.go:136-151
inline V3 lambert_sample(Rng &g, V3 normal) {
  double u = g.f64();
  double cos_lat = std::sqrt(u), sin_lat = std::sqrt(1 - u);
  double lon = g.f64() * 2 * M_PI;
  V3 xa, za;
  ortho_basis(normal, xa, za);
  V3 lon_point = add(scale(xa, std::cos(lon)), scale(za, std::sin(lon)));
  return add(scale(normal, -cos_lat), scale(lon_point, sin_lat));
}
// material.go:274-323
inline V3 sample_around_direction(Rng &g, double alpha, V3 direction) {
  V3 xa, za;
  ortho_basis(direction, xa, za);
  double u = g.f64(), v = g.f64();
  double lon = 2 * M_PI * u;
  double cos_lat = std::pow(v, 1 / (alpha + 1));
  double sin_lat = std::sqrt(1 - cos_lat * cos_lat);
  V3 lon_point = add(scale(xa, std::cos(lon)), scale(za, std::sin(lon)));
  return add(scale(direction, cos_lat), scale(lon_point, sin_lat));
}
// material.go:328-335
inline double density_around_direction(double alpha, V3 direction, V3 sample) {
  double d = dot(direction, sample);
  if (d < 0) return 0;
  double v = std::pow(d, alpha + 1);
  return 2 * (alpha + 1) / std::pow(v, 1 / (alpha + 1) - 1);
}
// material.go:218-223
inline double maximum_cosine(double c1, double c2) {
  double r = std::fmax(std::fabs(c1), std::fabs(c2));
  return std::fmax(r, kCosineEpsilon);
}

// material.go:360-378
inline V3 refract_dir(double ior, V3 normal, V3 source) {
  V3 sine_part = project_out(source, normal);
  double sine_scale = ior;
  V3 cosine_part = normal;
  if (dot(normal, source) < 0) {
    sine_scale = 1 / sine_scale;
    cosine_part = scale(cosine_part, -1);
  }
  sine_part = scale(sine_part, sine_scale);
  double sine_norm = norm(sine_part);
  if (std::fabs(sine_norm) > 1) return scale(reflect(normal, source), -1);
  cosine_part = scale(cosine_part, std::sqrt(1 - sine_norm * sine_norm));
  return add(sine_part, cosine_part);
}
inline V3 refract_inverse(double ior, V3 normal, V3 dest) {
  return scale(refract_dir(ior, normal, scale(dest, -1)), -1);
}
// material.go:384-389
inline double reflect_amount(double ior, V3 normal, V3 source) {
  double x = (ior - 1) / (ior + 1);
  double r0 = x * x;
  return r0 * (1 - r0) * std::pow(1 - std::fabs(dot(normal, source)), 5);
}
// material.go:401-414
inline double refract_bsdf(double ior, V3 normal, V3 source, V3 dest) {
  V3 refracted = refract_dir(ior, normal, source);
  if (dot(dest, refracted) < 1 - kCosineEpsilon) return 0;
  double s = 1 / std::fmax(kCosineEpsilon, std::fabs(dot(dest, normal)));
  return s * 2 / kCosineEpsilon;
}
// material.go:416-423
inline double reflect_bsdf(V3 normal, V3 source, V3 dest) {
  V3 reflected = scale(reflect(normal, source), -1);
  if (dot(dest, reflected) < 1 - kCosineEpsilon) return 0;
  double s = 1 / maximum_cosine(dot(dest, normal), dot(source, normal));
  return s * 2 / kCosineEpsilon;
}

inline bool is_zero(const double *c) { return c[0] == 0 && c[1] == 0 && c[2] == 0; }

inline V3 mat_bsdf(const MatRef &m, int32_t idx, V3 normal, V3 source, V3 dest) {
  const m3d_material_desc &d = m.tab->mats[idx];
  switch (d.kind) {
    case M3D_MAT_LAMBERT: {  // material.go:125-134
      if (dot(dest, normal) < 0 || dot(source, normal) > 0) return V3();
      return scale(mat_diffuse(m, d), 4);
    }
    case M3D_MAT_PHONG: {  // material.go:187-216
      double dest_dot = dot(dest, normal), source_dot = -dot(source, normal);
      if (dest_dot < 0 || source_dot < 0) return V3();
      V3 color;
      V3 diff = mat_diffuse(m, d);
      if (diff != V3()) color = scale(diff, 4);
      V3 reflection = scale(reflect(normal, source), -1);
      double ref_dot = dot(reflection, dest);
      if (ref_dot < 0) return color;
      double intensity = std::pow(ref_dot, d.alpha);
      intensity *= (1 + d.alpha);
      if (!(d.flags & M3D_MAT_NO_FLUX_CORRECTION)) intensity /= maximum_cosine(source_dot, dest_dot);
      return add(color, scale(v3(d.specular), 2 * intensity));
    }
    case M3D_MAT_REFRACT: {  // material.go:391-399
      if (is_zero(d.specular))
        return scale(v3(d.refract), refract_bsdf(d.index_of_refraction, normal, source, dest));
      double ra = reflect_amount(d.index_of_refraction, normal, source);
      double refr = (1 - ra) * refract_bsdf(d.index_of_refraction, normal, source, dest);
      double refl = ra * reflect_bsdf(normal, source, dest);
      return add(scale(v3(d.refract), refr), scale(v3(d.specular), refl));
    }
    case M3D_MAT_JOINED: {  // material.go:565-571
      V3 res;
      for (int i = 0; i < d.num_sub; i++) res = add(res, mat_bsdf(m, d.sub[i], normal, source, dest));
      return res;
    }
  }
  return V3();
}

inline V3 mat_sample_source(const MatRef &m, int32_t idx, Rng &g, V3 normal, V3 dest) {
  const m3d_material_desc &d = m.tab->mats[idx];
  switch (d.kind) {
    case M3D_MAT_LAMBERT:
      return lambert_sample(g, normal);
    case M3D_MAT_PHONG: {  // material.go:230-236, 251-256
      V3 diff = mat_diffuse(m, d);
      if (diff == V3() || g.intn(2) == 0) {
        V3 reflection = scale(reflect(normal, dest), -1);
        return sample_around_direction(g, d.alpha, reflection);
      }
      return lambert_sample(g, normal);
    }
    case M3D_MAT_REFRACT: {  // material.go:425-439
      if (is_zero(d.specular)) return refract_inverse(d.index_of_refraction, normal, dest);
      double refl = reflect_amount(d.index_of_refraction, normal, dest);
      if (g.f64() > refl) return refract_inverse(d.index_of_refraction, normal, dest);
      return scale(reflect(normal, dest), -1);
    }
    case M3D_MAT_JOINED: {  // material.go:573-586
      double p = g.f64();
      for (int i = 0; i < d.num_sub; i++) {
        p -= d.sub_prob[i];
        if (p < 0 || i == d.num_sub - 1) return mat_sample_source(m, d.sub[i], g, normal, dest);
      }
    }
  }
  return V3();
}

inline double mat_source_density(const MatRef &m, int32_t idx, V3 normal, V3 source, V3 dest) {
  const m3d_material_desc &d = m.tab->mats[idx];
  switch (d.kind) {
    case M3D_MAT_LAMBERT:
      return lambert_density(normal, source);
    case M3D_MAT_PHONG: {  // material.go:240-247, 258-261
      V3 reflection = scale(reflect(normal, dest), -1);
      double pw = density_around_direction(d.alpha, reflection, source);
      if (mat_diffuse(m, d) == V3()) return pw;
      return (pw + lambert_density(normal, source)) / 2;
    }
    case M3D_MAT_REFRACT: {  // material.go:441-462
      double ior = d.index_of_refraction;
      if (is_zero(d.specular)) {
        V3 refracted = refract_inverse(ior, normal, dest);
        if (dot(source, refracted) < 1 - kCosineEpsilon) return 0;
        return 2 / kCosineEpsilon;
      }
      double refl = reflect_amount(ior, normal, dest);
      V3 reflected = scale(reflect(normal, dest), -1);
      V3 refracted = refract_inverse(ior, normal, dest);
      double density = 0;
      if (dot(source, refracted) >= 1 - kCosineEpsilon) density += 1 - refl;
      if (dot(source, reflected) >= 1 - kCosineEpsilon) density += refl;
      return density * 2 / kCosineEpsilon;
    }
    case M3D_MAT_JOINED: {  // material.go:588-594
      double dens = 0;
      for (int i = 0; i < d.num_sub; i++)
        dens += d.sub_prob[i] * mat_source_density(m, d.sub[i], normal, source, dest);
      return dens;
    }
  }
  return 0;
}

// material.go:97-106, 464-467, 596-608
inline V3 mat_sample_dest(const MatRef &m, int32_t idx, Rng &g, V3 normal, V3 source) {
  const m3d_material_desc &d = m.tab->mats[idx];
  if (d.kind == M3D_MAT_REFRACT) return mat_sample_source(m, idx, g, scale(normal, -1), source);
  if (d.kind == M3D_MAT_JOINED) {
    double p = g.f64();
    for (int i = 0; i < d.num_sub; i++) {
      p -= d.sub_prob[i];
      if (p < 0 || i == d.num_sub - 1) return mat_sample_dest(m, d.sub[i], g, normal, source);
    }
  }
  return scale(mat_sample_source(m, idx, g, normal, scale(source, -1)), -1);
}
// material.go:108-116, 469-471, 610-616
inline double mat_dest_density(const MatRef &m, int32_t idx, V3 normal, V3 source, V3 dest) {
  const m3d_material_desc &d = m.tab->mats[idx];
  if (d.kind == M3D_MAT_REFRACT) return mat_source_density(m, idx, scale(normal, -1), dest, source);
  if (d.kind == M3D_MAT_JOINED) {
    double dens = 0;
    for (int i = 0; i < d.num_sub; i++)
      dens += d.sub_prob[i] * mat_dest_density(m, d.sub[i], normal, source, dest);
    return dens;
  }
  return mat_source_density(m, idx, normal, scale(dest, -1), scale(source, -1));
}

inline V3 mat_emission(const MatRef &m, int32_t idx) {
  const m3d_material_desc &d = m.tab->mats[idx];
  if (d.kind == M3D_MAT_REFRACT) return V3();
  if (d.kind == M3D_MAT_JOINED) {
    V3 r;
    for (int i = 0; i < d.num_sub; i++) r = add(r, mat_emission(m, d.sub[i]));
    return r;
  }
  return v3(d.emission);
}
inline V3 mat_ambient(const MatRef &m, int32_t idx) {
  const m3d_material_desc &d = m.tab->mats[idx];
  if (d.kind == M3D_MAT_REFRACT) return V3();
  if (d.kind == M3D_MAT_JOINED) {
    V3 r;
    for (int i = 0; i < d.num_sub; i++) r = add(r, mat_ambient(m, d.sub[i]));
    return r;
  }
  return v3(d.ambient);
}

// ---- objects ---------------------------------------------------------------------
enum ObjKind { OBJ_MESH = 0, OBJ_SPHERE = 1, OBJ_RECT = 2, OBJ_CYLINDER = 3 };

struct Object {
  int kind = OBJ_MESH;
  int32_t material = 0;
  uint32_t flags = 0;
  std::shared_ptr<MeshCollider> mesh;
  Sphere sphere;
  Rect rect;
  Cylinder cyl;
  bool has_xf = false;
  M3 matrix, inv;
  V3 offset;
};

struct Scene {
  std::vector<Object> objects;
  MaterialTable mats;

  // ColliderObject.Cast (object.go:43-46) under the optional transform wrappers
  // (transform.go:26-31, 76-85: ray mapped by the inverse, normal by the forward
  // matrix and re-normalised).
  bool cast_object(const Object &o, const Ray &r_in, Hit &h, Counters *cnt) const {
    Ray r = r_in;
    if (o.has_xf) {
      r.origin = mul_column(o.inv, sub(r_in.origin, o.offset));
      r.direction = mul_column(o.inv, r_in.direction);
    }
    bool ok = false;
    switch (o.kind) {
      case OBJ_MESH:
        ok = o.mesh->first_ray_collision(r, h, cnt);
        break;
      case OBJ_SPHERE:
        ok = sphere_first_hit(o.sphere, r, h);
        break;
      case OBJ_RECT:
        ok = rect_first_hit(o.rect, r, h);
        break;
      case OBJ_CYLINDER:
        ok = cylinder_first_hit(o.cyl, r, h);
        break;
    }
    if (!ok) return false;
    if (o.has_xf) h.normal = normalize(mul_column(o.matrix, h.normal));
    if (o.flags & M3D_OBJ_FLIP_NORMAL) h.normal = scale(h.normal, -1);  // showcase room.go:40-44
    return true;
  }

  // JoinedObject.Cast (object.go:141-153): linear scan, strict '<', first wins ties.
  bool cast(const Ray &r, Hit &out, int32_t &obj, Counters *cnt = nullptr) const {
    bool found = false;
    for (size_t i = 0; i < objects.size(); i++) {
      Hit h;
      if (cast_object(objects[i], r, h, cnt) && (!found || h.scale < out.scale)) {
        out = h;
        obj = (int32_t)i;
        found = true;
      }
    }
    return found;
  }

  // Material at a hit, resolving the showcase procedural variants
  // (room.go:61-75 FloorObject checker, models.go:79-97 VaseObject gradient).
  MatRef material_at(int32_t obj, V3 point) const {
    MatRef m;
    m.tab = &mats;
    m.index = objects[obj].material;
    const m3d_material_desc &d = mats.mats[m.index];
    if (d.flags & M3D_MAT_CHECKER) {
      bool same = int(std::fmod(point.x + 300, 2)) == int(std::fmod(point.y + 301, 2));
      m.diffuse_override = same ? v3(d.diffuse2) : v3(d.diffuse);
      m.has_override = true;
    } else if (d.flags & M3D_MAT_Z_GRADIENT) {
      double frac = point.z / d.proc_param;
      m.diffuse_override = add(scale(v3(d.diffuse), frac), scale(v3(d.diffuse2), 1 - frac));
      m.has_override = true;
    }
    return m;
  }
};

// ---- camera ------------------------------------------------------------------------
struct Caster {
  V3 x, y, z;
  double cx, cy;
  // camera.go:74-82
  V3 operator()(double ix, double iy) const {
    return add(add(scale(x, (ix - cx) / cx), scale(y, (iy - cy) / cy)), z);
  }
};
// camera.go:100-113 + 74-76
inline Caster make_caster(const m3d_camera &c, double w, double h) {
  double plane = 1 / std::tan(c.field_of_view / 2);
  V3 x = v3(c.screen_x), y = v3(c.screen_y);
  V3 z = normalize(cross(x, y));
  if (w > h)
    y = scale(y, h / w);
  else
    x = scale(x, w / h);
  z = scale(z, plane);
  return {x, y, z, w / 2, h / 2};
}
// camera.go:48-66
inline m3d_camera camera_at(V3 source, V3 dest, double fov) {
  if (fov == 0) fov = M_PI / 2;
  V3 za = normalize(sub(dest, source));
  V3 xa{za.y, -za.x, 0};
  if (norm(xa) < 1e-5) xa = project_out(V3(1, 0, 0), za);
  xa = normalize(xa);
  V3 ya = cross(za, xa);
  m3d_camera c;
  std::memset(&c, 0, sizeof(c));
  for (int i = 0; i < 3; i++) {
    c.origin[i] = source[i];
    c.screen_x[i] = xa[i];
    c.screen_y[i] = ya[i];
  }
  c.field_of_view = fov;
  return c;
}

// light.go:71-92
inline V3 shade_collision(const m3d_point_light &l, V3 normal, V3 point_to_light) {
  double d = norm(point_to_light);
  V3 color = v3(l.color);
  if (l.quad_dropoff) color = scale(color, 1 / (d * d));
  double density = 0.25 * std::fmax(0.0, dot(normal, scale(point_to_light, 1 / d)));
  return scale(color, density);
}

// light.go:41-47
inline double gamma_compress(double u) {
  if (u <= 0.0031308) return 12.92 * u;
  return 1.055 * std::pow(u, 1 / 2.4) - 0.055;
}
// light.go:49-55
inline double gamma_expand(double u) {
  if (u <= 0.04045) return u / 12.92;
  return std::pow((u + 0.055) / 1.055, 2.4);
}
// image.go:125-145
inline uint8_t to_srgb8(double c) {
  c = std::fmin(1.0, std::fmax(0.0, c));
  return (uint8_t)(gamma_compress(c) * (256.0 - 0.001));
}

// mapCoordinates (concurrency.go:17-43) stand-in: rows striped over threads.
inline void parallel_rows(int height, int nthreads, const std::function<void(int, int)> &f) {
  if (nthreads <= 1) {
    for (int y = 0; y < height; y++) f(y, 0);
    return;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++)
    th.emplace_back([&, t] {
      for (int y = t; y < height; y += nthreads) f(y, t);
    });
  for (auto &x : th) x.join();
}

// raycast.go:15-39.  img: W*H*3 doubles, left untouched where the ray misses.
// Optional per-pixel outputs for parity checks: t, obj, prim.
inline void render_raycast(const Scene &sc, const m3d_camera &cam, const m3d_point_light *lights,
                           int nl, int W, int H, double *img, double *t_out, int32_t *obj_out,
                           int32_t *prim_out, int nthreads) {
  Caster caster = make_caster(cam, double(W) - 1, double(H) - 1);
  parallel_rows(H, nthreads, [&](int y, int) {
    for (int x = 0; x < W; x++) {
      int idx = x + y * W;
      Ray ray{v3(cam.origin), caster(double(x), double(y))};
      Hit h;
      int32_t obj = -1;
      bool ok = sc.cast(ray, h, obj);
      if (obj_out) obj_out[idx] = ok ? obj : -1;
      if (prim_out) prim_out[idx] = ok ? h.prim : -1;
      if (t_out) t_out[idx] = ok ? h.scale : 0;
      if (!ok) continue;
      V3 point = add(ray.origin, scale(ray.direction, h.scale));
      MatRef m = sc.material_at(obj, point);
      V3 color = add(mat_ambient(m, m.index), mat_emission(m, m.index));
      for (int li = 0; li < nl; li++) {
        const m3d_point_light &l = lights[li];
        V3 lo = v3(l.origin);
        V3 brdf = mat_bsdf(m, m.index, h.normal, normalize(sub(point, lo)),
                           normalize(sub(ray.origin, point)));
        V3 p2l = sub(lo, point);
        color = add(color, mul(shade_collision(l, h.normal, p2l), brdf));
      }
      img[idx * 3 + 0] = color.x;
      img[idx * 3 + 1] = color.y;
      img[idx * 3 + 2] = color.z;
    }
  });
}

}  // namespace orc
