// ORACLE -- TEST INFRASTRUCTURE ONLY (see vec.hpp header).
// C entry points (ctypes) over the float64 restatement.  Built by oracle/Makefile
// into oracle/liboracle.so.  Never linked into or called by libm3dgpu.
#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>

#include "collide.hpp"
#include "mcubes.hpp"
#include "meshgen.hpp"
#include "render.hpp"
#include "scene.hpp"
#include "sdf.hpp"

using namespace orc;

namespace {
template <class F>
void parallel_for(int64_t n, int nthreads, F f) {
  if (nthreads <= 1 || n < 2) {
    f(0, n, 0);
    return;
  }
  std::vector<std::thread> th;
  // dynamic chunks: ray cost is very uneven
  std::atomic<int64_t> next{0};
  const int64_t chunk = 4096;
  for (int t = 0; t < nthreads; t++)
    th.emplace_back([&, t] {
      for (;;) {
        int64_t b = next.fetch_add(chunk);
        if (b >= n) break;
        f(b, std::min(n, b + chunk), t);
      }
    });
  for (auto &x : th) x.join();
}

void tris_from_f32(const float *v, int64_t n, const float *vn, std::vector<Triangle> &out) {
  out.resize(n);
  for (int64_t i = 0; i < n; i++)
    for (int k = 0; k < 3; k++) {
      out[i].p[k] = V3(v[i * 9 + k * 3], v[i * 9 + k * 3 + 1], v[i * 9 + k * 3 + 2]);
      if (vn) out[i].vn[k] = V3(vn[i * 9 + k * 3], vn[i * 9 + k * 3 + 1], vn[i * 9 + k * 3 + 2]);
    }
}
void tris_to_f64(const std::vector<Triangle> &m, double *out) {
  for (size_t i = 0; i < m.size(); i++)
    for (int k = 0; k < 3; k++) {
      out[i * 9 + k * 3 + 0] = m[i].p[k].x;
      out[i * 9 + k * 3 + 1] = m[i].p[k].y;
      out[i * 9 + k * 3 + 2] = m[i].p[k].z;
    }
}
}  // namespace

extern "C" {

int orc_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

// ---- mesh generators (float64 out, n*9) --------------------------------------------
int64_t orc_mesh_icosphere_count(int n) { return 20LL * n * n; }
void orc_mesh_icosphere(double cx, double cy, double cz, double radius, int n, double *out) {
  tris_to_f64(mesh_icosphere(V3(cx, cy, cz), radius, n), out);
}
void orc_mesh_rect(const double mn[3], const double mx[3], double *out /*12*9*/) {
  tris_to_f64(mesh_rect(v3(mn), v3(mx)), out);
}
// marching-cubes sphere (C1): two calls, count then fill (the mesh is cached between them)
static std::vector<Triangle> g_mc_cache;
int64_t orc_mesh_mc_sphere_build(double cx, double cy, double cz, double radius, double delta, int iters) {
  try {
    g_mc_cache = marching_cubes_sphere(V3(cx, cy, cz), radius, delta, iters);
  } catch (const std::exception &e) {
    g_mc_cache.clear();
    return -1;
  }
  return (int64_t)g_mc_cache.size();
}
void orc_mesh_mc_sphere_fetch(double *out) {
  tris_to_f64(g_mc_cache, out);
  g_mc_cache.clear();
  g_mc_cache.shrink_to_fit();
}
// the 256-entry case table, flattened: counts[256], corners[256][5][6] (0-padded)
void orc_mc_table(int32_t *counts, uint8_t *corners) {
  const auto &t = mc_lookup_table();
  for (int i = 0; i < 256; i++) {
    counts[i] = (int32_t)t[i].size();
    for (size_t k = 0; k < 5; k++)
      for (int c = 0; c < 6; c++) corners[(i * 5 + k) * 6 + c] = k < t[i].size() ? t[i][k][c] : 0;
  }
}
int64_t orc_mesh_polar_count(int stops) { return 2LL * stops * (stops - 1); }
void orc_mesh_polar(double ra, double rb, int stops, double *out) {
  tris_to_f64(mesh_polar(ra, rb, stops), out);
}

// ---- mesh collider --------------------------------------------------------------------
void *orc_collider_create(const float *tris, int64_t n, const float *vnormals) {
  auto *m = new MeshCollider();
  tris_from_f32(tris, n, vnormals, m->tris);
  m->interp = vnormals != nullptr;
  m->build(false);
  return m;
}
void *orc_collider_create_f64(const double *tris, int64_t n) {
  auto *m = new MeshCollider();
  m->tris.resize(n);
  for (int64_t i = 0; i < n; i++)
    for (int k = 0; k < 3; k++) m->tris[i].p[k] = V3(tris[i * 9 + k * 3], tris[i * 9 + k * 3 + 1], tris[i * 9 + k * 3 + 2]);
  m->build(false);
  return m;
}
void orc_collider_destroy(void *h) { delete (MeshCollider *)h; }
int64_t orc_collider_num_nodes(void *h) { return (int64_t)((MeshCollider *)h)->nodes.size(); }
void orc_collider_bounds(void *h, double mn[3], double mx[3]) {
  auto *m = (MeshCollider *)h;
  V3 a = m->mn(), b = m->mx();
  for (int i = 0; i < 3; i++) {
    mn[i] = a[i];
    mx[i] = b[i];
  }
}
void orc_collider_order(void *h, int32_t *out) {
  auto *m = (MeshCollider *)h;
  std::memcpy(out, m->order.data(), m->order.size() * sizeof(int32_t));
}

// Batched FirstRayCollision.  org/dir: n*3 float64.  Outputs may be NULL.
// counters (optional, 2 x int64): total nodes visited, triangles tested.
void orc_collider_first_hits(void *h, const double *org, const double *dir, int64_t n, double *t,
                             int32_t *prim, double *normal, double *bary, int64_t *counters,
                             int nthreads) {
  auto *m = (MeshCollider *)h;
  std::atomic<int64_t> cn{0}, ct{0};
  parallel_for(n, nthreads, [&](int64_t b, int64_t e, int) {
    Counters c;
    for (int64_t i = b; i < e; i++) {
      Ray r{V3(org[i * 3], org[i * 3 + 1], org[i * 3 + 2]), V3(dir[i * 3], dir[i * 3 + 1], dir[i * 3 + 2])};
      Hit hit;
      bool ok = m->first_ray_collision(r, hit, counters ? &c : nullptr);
      if (prim) prim[i] = ok ? hit.prim : -1;
      if (t) t[i] = ok ? hit.scale : 0;
      if (normal)
        for (int k = 0; k < 3; k++) normal[i * 3 + k] = ok ? hit.normal[k] : 0;
      if (bary)
        for (int k = 0; k < 3; k++) bary[i * 3 + k] = ok ? hit.bary[k] : 0;
    }
    cn += c.nodes;
    ct += c.tris;
  });
  if (counters) {
    counters[0] = cn;
    counters[1] = ct;
  }
}
// float32 in (widened), same outputs; used by bench.py's CPU baseline on the same inputs.
void orc_collider_first_hits_f32(void *h, const float *org, const float *dir, int64_t n, double *t,
                                 int32_t *prim, double *normal, double *bary, int64_t *counters,
                                 int nthreads) {
  auto *m = (MeshCollider *)h;
  std::atomic<int64_t> cn{0}, ct{0};
  parallel_for(n, nthreads, [&](int64_t b, int64_t e, int) {
    Counters c;
    for (int64_t i = b; i < e; i++) {
      Ray r{V3(org[i * 3], org[i * 3 + 1], org[i * 3 + 2]), V3(dir[i * 3], dir[i * 3 + 1], dir[i * 3 + 2])};
      Hit hit;
      bool ok = m->first_ray_collision(r, hit, counters ? &c : nullptr);
      if (prim) prim[i] = ok ? hit.prim : -1;
      if (t) t[i] = ok ? hit.scale : 0;
      if (normal)
        for (int k = 0; k < 3; k++) normal[i * 3 + k] = ok ? hit.normal[k] : 0;
      if (bary)
        for (int k = 0; k < 3; k++) bary[i * 3 + k] = ok ? hit.bary[k] : 0;
    }
    cn += c.nodes;
    ct += c.tris;
  });
  if (counters) {
    counters[0] = cn;
    counters[1] = ct;
  }
}

// RayCollisions (all hits) for one ray via the BVH (brute == 0) or by testing every
// triangle (brute == 1).  Returns the count; writes up to cap (t, prim) pairs sorted by t.
int64_t orc_collider_all_hits(void *h, const double org[3], const double dir[3], int brute,
                              double *t_out, int32_t *prim_out, int64_t cap) {
  auto *m = (MeshCollider *)h;
  Ray r{v3(org), v3(dir)};
  std::vector<Hit> hits;
  if (brute)
    m->brute_collisions(r, hits);
  else
    m->ray_collisions(r, hits);
  std::stable_sort(hits.begin(), hits.end(), [](const Hit &a, const Hit &b) { return a.scale < b.scale; });
  for (int64_t i = 0; i < (int64_t)hits.size() && i < cap; i++) {
    t_out[i] = hits[i].scale;
    prim_out[i] = hits[i].prim;
  }
  return (int64_t)hits.size();
}

// Batched Collider.RayCollisions(r, nil) counts (collisions.go:263-273, primitives.go:189-196).
void orc_collider_hit_counts(void *h, const float *org, const float *dir, int64_t n, int32_t *counts,
                             int nthreads) {
  auto *m = (MeshCollider *)h;
  parallel_for(n, nthreads, [&](int64_t b, int64_t e, int) {
    std::vector<Hit> hits;
    for (int64_t i = b; i < e; i++) {
      Ray r{V3(org[3 * i], org[3 * i + 1], org[3 * i + 2]), V3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2])};
      hits.clear();
      m->ray_collisions(r, hits);
      counts[i] = (int32_t)hits.size();
    }
  });
}

// Batched Collider.RayCollisions(r, f) with the collisions delivered (collisions.go:263-273): the
// hits of ray i go to rows [offsets[i], offsets[i+1]) ordered by scale (stable: ties keep the
// reference traversal order); offsets must come from orc_collider_hit_counts.
void orc_collider_all_hits_batch(void *h, const float *org, const float *dir, int64_t n, const int64_t *offsets,
                                 double *t, int32_t *prim, double *normal, double *bary, int nthreads) {
  auto *m = (MeshCollider *)h;
  parallel_for(n, nthreads, [&](int64_t b, int64_t e, int) {
    std::vector<Hit> hits;
    for (int64_t i = b; i < e; i++) {
      Ray r{V3(org[3 * i], org[3 * i + 1], org[3 * i + 2]), V3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2])};
      hits.clear();
      m->ray_collisions(r, hits);
      std::stable_sort(hits.begin(), hits.end(), [](const Hit &a, const Hit &b) { return a.scale < b.scale; });
      const int64_t cap = offsets[i + 1] - offsets[i];
      for (int64_t k = 0; k < (int64_t)hits.size() && k < cap; k++) {
        const int64_t o = offsets[i] + k;
        t[o] = hits[k].scale;
        prim[o] = hits[k].prim;
        if (normal) {
          normal[3 * o] = hits[k].normal.x;
          normal[3 * o + 1] = hits[k].normal.y;
          normal[3 * o + 2] = hits[k].normal.z;
        }
        if (bary) {
          bary[3 * o] = hits[k].bary[0];
          bary[3 * o + 1] = hits[k].bary[1];
          bary[3 * o + 2] = hits[k].bary[2];
        }
      }
    }
  });
}

// ColliderContains with margin 0 (collisions.go:119-134): odd number of collisions along the
// reference's fixed direction.
void orc_collider_contains(void *h, const float *pts, int64_t n, uint8_t *inside, int nthreads) {
  auto *m = (MeshCollider *)h;
  parallel_for(n, nthreads, [&](int64_t b, int64_t e, int) {
    std::vector<Hit> hits;
    for (int64_t i = b; i < e; i++) {
      Ray r{V3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]),
            V3(0.5224892708603626, 0.10494477243214506, 0.43558938446126527)};
      hits.clear();
      m->ray_collisions(r, hits);
      inside[i] = (uint8_t)(hits.size() % 2);
    }
  });
}

// ---- nearest-triangle queries (sdf.hpp) ---------------------------------------------------
// MeshToSDF(...).FaceSDF (sdf.go:186-240): signed distance (positive inside), nearest point
// (n*3 or NULL), face id (n or NULL).
void orc_collider_sdf(void *h, const float *pts, int64_t n, double *sdf, double *point, int32_t *face,
                      int nthreads) {
  auto *m = (MeshCollider *)h;
  MeshDistFunc mdf(*m);
  parallel_for(n, nthreads, [&](int64_t b, int64_t e, int) {
    for (int64_t i = b; i < e; i++) {
      V3 p;
      int32_t f;
      sdf[i] = mdf.face_sdf(V3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), p, f);
      if (point) point[3 * i] = p.x, point[3 * i + 1] = p.y, point[3 * i + 2] = p.z;
      if (face) face[i] = f;
    }
  });
}
// Collider.SphereCollision (collisions.go:292-303, primitives.go:253-279), batched.
void orc_collider_sphere_collisions(void *h, const float *centers, const double *radii, int64_t n,
                                    uint8_t *out, int nthreads) {
  auto *m = (MeshCollider *)h;
  parallel_for(n, nthreads, [&](int64_t b, int64_t e, int) {
    for (int64_t i = b; i < e; i++)
      out[i] = sphere_collision(*m, V3(centers[3 * i], centers[3 * i + 1], centers[3 * i + 2]), radii[i]);
  });
}
// ColliderContains(c, p, margin) (collisions.go:119-134); solid != 0: ColliderSolid.Contains
// (solid.go:292-300, bounds check first, margin = inset).
void orc_collider_contains_margin(void *h, const float *pts, int64_t n, double margin, int solid,
                                  uint8_t *inside, int nthreads) {
  auto *m = (MeshCollider *)h;
  parallel_for(n, nthreads, [&](int64_t b, int64_t e, int) {
    for (int64_t i = b; i < e; i++) {
      V3 p(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
      bool in = true;
      if (solid) {
        V3 mn = m->mn(), mx = m->mx();
        if (solid == 2) {  // NewColliderSolidInset bounds (solid.go:265-270)
          V3 iv(margin, margin, margin);
          mn = add(mn, iv);
          mx = vmax(mn, sub(m->mx(), iv));
        }
        in = !m->empty && vmin(p, mn) == mn && vmax(p, mx) == mx;
      }
      inside[i] = in && collider_contains(*m, p, margin);
    }
  });
}
void orc_triangle_closest(const double tri[9], const double p[3], double out[3]) {
  Triangle tr = mk_tri(v3(tri), v3(tri + 3), v3(tri + 6));
  V3 c = tri_closest(tr, v3(p));
  out[0] = c.x, out[1] = c.y, out[2] = c.z;
}
int orc_triangle_sphere_collision(const double tri[9], const double c[3], double r) {
  Triangle tr = mk_tri(v3(tri), v3(tri + 3), v3(tri + 6));
  return tri_sphere_collision(tr, v3(c), r);
}

// Single-triangle test exposed for edge-case tests (primitives.go:181-249).
int orc_triangle_first_hit(const double tri[9], const double org[3], const double dir[3], double *t,
                           double normal[3], double bary[3]) {
  Triangle tr = mk_tri(v3(tri), v3(tri + 3), v3(tri + 6));
  Hit h;
  if (!tri_first_hit(tr, false, Ray{v3(org), v3(dir)}, h)) return 0;
  *t = h.scale;
  for (int i = 0; i < 3; i++) {
    normal[i] = h.normal[i];
    bary[i] = h.bary[i];
  }
  return 1;
}

// ---- analytic shapes (kind: 1 sphere, 2 rect, 3 cylinder; params as in m3d.h) ---------
int orc_shape_first_hit(int kind, const double p0[3], const double p1[3], double radius,
                        const double org[3], const double dir[3], double *t, double normal[3]) {
  Ray r{v3(org), v3(dir)};
  Hit h;
  bool ok = false;
  if (kind == OBJ_SPHERE) ok = sphere_first_hit(Sphere{v3(p0), radius}, r, h);
  if (kind == OBJ_RECT) ok = rect_first_hit(Rect{v3(p0), v3(p1)}, r, h);
  if (kind == OBJ_CYLINDER) ok = cylinder_first_hit(Cylinder{v3(p0), v3(p1), radius}, r, h);
  if (!ok) return 0;
  *t = h.scale;
  for (int i = 0; i < 3; i++) normal[i] = h.normal[i];
  return 1;
}

// ---- scene ------------------------------------------------------------------------------
void *orc_scene_create() { return new Scene(); }
void orc_scene_destroy(void *s) { delete (Scene *)s; }
int32_t orc_scene_add_material(void *s, const m3d_material_desc *d) {
  auto *sc = (Scene *)s;
  sc->mats.mats.push_back(*d);
  return (int32_t)sc->mats.mats.size() - 1;
}
static void set_xf(Object &o, const m3d_transform *xf) {
  if (!xf) return;
  o.has_xf = true;
  std::memcpy(o.matrix.m, xf->matrix, sizeof(double) * 9);
  o.inv = inverse(o.matrix);
  o.offset = v3(xf->offset);
}
int32_t orc_scene_add_mesh(void *s, const float *tris, int64_t n, const float *vn, int32_t material,
                           uint32_t flags, const m3d_transform *xf) {
  auto *sc = (Scene *)s;
  Object o;
  o.kind = OBJ_MESH;
  o.material = material;
  o.flags = flags;
  o.mesh = std::make_shared<MeshCollider>();
  tris_from_f32(tris, n, vn, o.mesh->tris);
  o.mesh->interp = vn != nullptr;
  o.mesh->build(false);
  set_xf(o, xf);
  sc->objects.push_back(std::move(o));
  return (int32_t)sc->objects.size() - 1;
}
int32_t orc_scene_add_sphere(void *s, const double c[3], double radius, int32_t material,
                             uint32_t flags, const m3d_transform *xf) {
  auto *sc = (Scene *)s;
  Object o;
  o.kind = OBJ_SPHERE;
  o.material = material;
  o.flags = flags;
  o.sphere = Sphere{v3(c), radius};
  set_xf(o, xf);
  sc->objects.push_back(std::move(o));
  return (int32_t)sc->objects.size() - 1;
}
int32_t orc_scene_add_rect(void *s, const double mn[3], const double mx[3], int32_t material,
                           uint32_t flags, const m3d_transform *xf) {
  auto *sc = (Scene *)s;
  Object o;
  o.kind = OBJ_RECT;
  o.material = material;
  o.flags = flags;
  o.rect = Rect{v3(mn), v3(mx)};
  set_xf(o, xf);
  sc->objects.push_back(std::move(o));
  return (int32_t)sc->objects.size() - 1;
}
int32_t orc_scene_add_cylinder(void *s, const double p1[3], const double p2[3], double radius,
                               int32_t material, uint32_t flags, const m3d_transform *xf) {
  auto *sc = (Scene *)s;
  Object o;
  o.kind = OBJ_CYLINDER;
  o.material = material;
  o.flags = flags;
  o.cyl = Cylinder{v3(p1), v3(p2), radius};
  set_xf(o, xf);
  sc->objects.push_back(std::move(o));
  return (int32_t)sc->objects.size() - 1;
}

// Batched Object.Cast over a scene.
void orc_scene_cast(void *s, const double *org, const double *dir, int64_t n, double *t,
                    int32_t *obj, int32_t *prim, double *normal, int nthreads) {
  auto *sc = (Scene *)s;
  parallel_for(n, nthreads, [&](int64_t b, int64_t e, int) {
    for (int64_t i = b; i < e; i++) {
      Ray r{V3(org[i * 3], org[i * 3 + 1], org[i * 3 + 2]), V3(dir[i * 3], dir[i * 3 + 1], dir[i * 3 + 2])};
      Hit hit;
      int32_t o = -1;
      bool ok = sc->cast(r, hit, o);
      if (obj) obj[i] = ok ? o : -1;
      if (prim) prim[i] = ok ? hit.prim : -1;
      if (t) t[i] = ok ? hit.scale : 0;
      if (normal)
        for (int k = 0; k < 3; k++) normal[i * 3 + k] = ok ? hit.normal[k] : 0;
    }
  });
}

void orc_camera_at(const double src[3], const double dst[3], double fov, m3d_camera *out) {
  *out = camera_at(v3(src), v3(dst), fov);
}
void orc_camera_rays(const m3d_camera *cam, int W, int H, double *dir_out /*W*H*3*/) {
  Caster c = make_caster(*cam, double(W) - 1, double(H) - 1);
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      V3 d = c(x, y);
      for (int k = 0; k < 3; k++) dir_out[(x + y * W) * 3 + k] = d[k];
    }
}

void orc_render_raycast(void *s, const m3d_camera *cam, const m3d_point_light *lights, int nl, int W,
                        int H, double *img, double *t_out, int32_t *obj_out, int32_t *prim_out,
                        int nthreads) {
  render_raycast(*(Scene *)s, *cam, lights, nl, W, H, img, t_out, obj_out, prim_out, nthreads);
}

void orc_srgb8(const double *rgb, int64_t n, uint8_t *out) {
  for (int64_t i = 0; i < n; i++) out[i] = to_srgb8(rgb[i]);
}
double orc_gamma_expand(double u) { return gamma_expand(u); }

// ---- path tracers (render.hpp) ---------------------------------------------------------
void orc_render_path(void *s, const m3d_camera *cam, const m3d_point_light *lights, int nl,
                     const m3d_path_params *p, int W, int H, double *mean, double *var_of_mean,
                     int64_t *rays_cast, int nthreads) {
  render_path(*(Scene *)s, *cam, lights, nl, *p, W, H, mean, var_of_mean, rays_cast, nthreads);
}
void orc_render_bidir(void *s, const m3d_camera *cam, const m3d_area_light *lights, int nl,
                      const m3d_bidir_params *p, int W, int H, double *mean, double *var_of_mean,
                      int64_t *rays_cast, int nthreads) {
  render_bidir(*(Scene *)s, *cam, lights, nl, *p, W, H, mean, var_of_mean, rays_cast, nthreads);
}

// Material sampling hooks for the restated material tests (material_test.go).
void orc_material_eval(void *s, int32_t mat, const double normal[3], const double source[3],
                       const double dest[3], double bsdf[3], double *source_density,
                       double *dest_density) {
  auto *sc = (Scene *)s;
  MatRef m;
  m.tab = &sc->mats;
  m.index = mat;
  V3 b = mat_bsdf(m, mat, v3(normal), v3(source), v3(dest));
  for (int i = 0; i < 3; i++) bsdf[i] = b[i];
  *source_density = mat_source_density(m, mat, v3(normal), v3(source), v3(dest));
  *dest_density = mat_dest_density(m, mat, v3(normal), v3(source), v3(dest));
}
void orc_material_sample_source(void *s, int32_t mat, uint64_t seed, const double normal[3],
                                const double dest[3], int64_t n, double *out /*n*3*/) {
  auto *sc = (Scene *)s;
  MatRef m;
  m.tab = &sc->mats;
  m.index = mat;
  Rng g(seed);
  for (int64_t i = 0; i < n; i++) {
    V3 v = mat_sample_source(m, mat, g, v3(normal), v3(dest));
    for (int k = 0; k < 3; k++) out[i * 3 + k] = v[k];
  }
}
// sampleAroundUniform / densityAroundUniform (focus_point.go:155-177): n directions in the cone
// around `direction` and their densities (for the restated TestSampleAroundUniform)
void orc_sample_around_uniform(uint64_t seed, double min_cos, const double direction[3], int64_t n,
                               double *out /*n*3*/, double *density /*n*/) {
  Rng g(seed);
  for (int64_t i = 0; i < n; i++) {
    V3 v = sample_around_uniform(g, min_cos, v3(direction));
    for (int k = 0; k < 3; k++) out[i * 3 + k] = v[k];
    density[i] = density_around_uniform(min_cos, v3(direction), v);
  }
}
void orc_material_sample_dest(void *s, int32_t mat, uint64_t seed, const double normal[3],
                              const double source[3], int64_t n, double *out /*n*3*/) {
  auto *sc = (Scene *)s;
  MatRef m;
  m.tab = &sc->mats;
  m.index = mat;
  Rng g(seed);
  for (int64_t i = 0; i < n; i++) {
    V3 v = mat_sample_dest(m, mat, g, v3(normal), v3(source));
    for (int k = 0; k < 3; k++) out[i * 3 + k] = v[k];
  }
}

// Batched material evaluation (restated material_test.go integrals run over many directions).
void orc_material_eval_batch(void *s, int32_t mat, const double normal[3], const double *sources,
                             const double *dests, int64_t n, double *bsdf /*n*3*/,
                             double *source_density /*n*/, double *dest_density /*n*/) {
  auto *sc = (Scene *)s;
  MatRef m;
  m.tab = &sc->mats;
  m.index = mat;
  for (int64_t i = 0; i < n; i++) {
    V3 src = v3(sources + 3 * i), dst = v3(dests + 3 * i);
    V3 b = mat_bsdf(m, mat, v3(normal), src, dst);
    for (int k = 0; k < 3; k++) bsdf[i * 3 + k] = b[k];
    source_density[i] = mat_source_density(m, mat, v3(normal), src, dst);
    dest_density[i] = mat_dest_density(m, mat, v3(normal), src, dst);
  }
}

}  // extern "C"
