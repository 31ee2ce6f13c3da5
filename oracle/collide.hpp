// ORACLE -- TEST INFRASTRUCTURE ONLY (see vec.hpp header for the parity status: pinned by the
// renderings the reference commits -- statistically by cornell_box / showcase, pixel for pixel by
// the deterministic smooth_shading RayCaster image -- and by the reference's own property tests
// restated in tests/test_oracle_collide.py; per-ray outputs as numbers otherwise "parity
// unpinned" -- no Go toolchain).
//
// float64 CPU restatement of the reference's collision path:
//   Ray / RayCollision               model3d/collisions.go:12-46
//   rayCollisionWithBounds           model3d/bvh.go:322-351
//   GroupBounders & helpers          model3d/bvh.go:131-292
//   GroupedTrianglesToCollider       model3d/collisions.go:169-179
//   NewJoinedCollider / FirstRayCollision / RayCollisions
//                                    model3d/collisions.go:225-308
//   Triangle.rayCollision / Normal   model3d/primitives.go:27-33,181-249
//   InterpNormalTriangle             model3d/primitives.go:499-537
//   Sphere / Rect / Cylinder         model3d/shapes.go:35-93,177-247,601-705,816-856
//
// Determinism: the reference's triangle order is random per run (Go map
// iteration, mesh.go:757-763) and sort.Slice is unstable; here the caller's
// order is the triangle id and sorts are stable.
#pragma once
#include <algorithm>
#include <cstdint>
#include <limits>
#include <memory>
#include <vector>

#include "vec.hpp"

namespace orc {

struct Ray {
  V3 origin, direction;
};

struct Hit {
  double scale = 0;
  V3 normal;
  int32_t prim = -1;  // triangle id (caller order) or 0 for analytic shapes
  double bary[3] = {0, 0, 0};
};

struct Counters {
  int64_t nodes = 0, tris = 0;
};

// ---- bvh.go:322-351 --------------------------------------------------------
inline void ray_bounds(const Ray &r, const V3 &mn, const V3 &mx, double &min_frac, double &max_frac) {
  min_frac = -std::numeric_limits<double>::infinity();
  max_frac = std::numeric_limits<double>::infinity();
  for (int axis = 0; axis < 3; axis++) {
    double origin = r.origin[axis];
    double rate = r.direction[axis];
    if (rate == 0) {
      if (origin < mn[axis] || origin > mx[axis]) {
        min_frac = 0;
        max_frac = -1;
        return;
      }
      continue;
    }
    double t1 = (mn[axis] - origin) / rate;
    double t2 = (mx[axis] - origin) / rate;
    if (t1 > t2) std::swap(t1, t2);
    if (t2 < 0) {
      min_frac = 0;
      max_frac = -1;
      return;
    }
    if (t1 > min_frac) min_frac = t1;
    if (t2 < max_frac) max_frac = t2;
  }
}

// ---- primitives.go -------------------------------------------------------------
struct Triangle {
  V3 p[3];
  V3 vn[3];  // vertex normals (InterpNormalTriangle), unused when !interp
  V3 mn() const { return vmin(vmin(p[0], p[1]), p[2]); }
  V3 mx() const { return vmax(vmax(p[0], p[1]), p[2]); }
  // primitives.go:27-33
  V3 normal() const { return normalize(cross(sub(p[1], p[0]), sub(p[2], p[0]))); }
  double area() const { return norm(cross(sub(p[1], p[0]), sub(p[2], p[0]))) / 2; }
};

// primitives.go:207-249.  Returns false for "nil"; scale may be negative.
inline bool tri_ray_collision(const Triangle &t, const Ray &r, double &scale, double bary[3]) {
  double d = dot(t.normal(), normalize(r.direction));
  if (d < 1e-8 && d > -1e-8) return false;
  V3 v1 = sub(t.p[1], t.p[0]);
  V3 v2 = sub(t.p[2], t.p[0]);
  V3 cross1 = cross(r.direction, v2);
  double det = dot(cross1, v1);
  if (det == 0) return false;
  double inv_det = 1 / det;
  V3 o = sub(r.origin, t.p[0]);
  double bary1 = inv_det * dot(o, cross1);
  if (bary1 < 0 || bary1 > 1) return false;
  V3 cross2 = cross(o, v1);
  double bary2 = inv_det * dot(r.direction, cross2);
  if (bary2 < 0 || bary1 + bary2 > 1) return false;
  bary[0] = 1 - (bary1 + bary2);
  bary[1] = bary1;
  bary[2] = bary2;
  scale = inv_det * dot(v2, cross2);
  return true;
}

// primitives.go:181-188 and :520-531 (interp)
inline bool tri_first_hit(const Triangle &t, bool interp, const Ray &r, Hit &h) {
  double sc, bary[3];
  if (!tri_ray_collision(t, r, sc, bary) || !(sc >= 0)) return false;
  h.scale = sc;
  for (int i = 0; i < 3; i++) h.bary[i] = bary[i];
  if (interp) {
    V3 n;
    for (int j = 0; j < 3; j++) n = add(n, scale(t.vn[j], bary[j]));
    h.normal = normalize(n);
  } else {
    h.normal = t.normal();
  }
  return true;
}

// ---- bvh.go:131-292 GroupBounders ---------------------------------------------
struct Flagged {
  int32_t id;
  V3 mn, mx, md;
  bool flag = false;
};

inline double bounds_area(const V3 &mn, const V3 &mx) {
  V3 d = sub(mx, mn);
  return 2 * (d.x * (d.y + d.z) + d.y * d.z);
}

inline double multiple_bounds_area(Flagged *const *bs, size_t n) {
  V3 mn = bs[0]->mn, mx = bs[0]->mx;
  for (size_t i = 1; i < n; i++) {
    mn = vmin(mn, bs[i]->mn);
    mx = vmax(mx, bs[i]->mx);
  }
  return bounds_area(mn, mx);
}

struct Sorted3 {
  std::vector<Flagged *> a[3];
};

// bvh.go:158-197
inline void split_bounders(const Sorted3 &s, int axis, size_t mid, Sorted3 &lo, Sorted3 &hi) {
  size_t n = s.a[0].size();
  for (size_t i = 0; i < n; i++) s.a[axis][i]->flag = i < mid;
  for (int ax = 0; ax < 3; ax++) {
    lo.a[ax].clear();
    hi.a[ax].clear();
    lo.a[ax].reserve(mid);
    hi.a[ax].reserve(n - mid);
    if (ax == axis) {
      lo.a[ax].assign(s.a[ax].begin(), s.a[ax].begin() + mid);
      hi.a[ax].assign(s.a[ax].begin() + mid, s.a[ax].end());
    } else {
      for (Flagged *b : s.a[ax]) (b->flag ? lo : hi).a[ax].push_back(b);
    }
  }
}

// bvh.go:199-217
inline int best_split_axis(const Sorted3 &s) {
  size_t n = s.a[0].size(), mid = n / 2;
  int axis = 0;
  double best = 0;
  for (int i = 0; i < 3; i++) {
    double a = multiple_bounds_area(s.a[i].data(), mid) +
               multiple_bounds_area(s.a[i].data() + mid, n - mid);
    if (i == 0 || a < best) {
      best = a;
      axis = i;
    }
  }
  return axis;
}

// bvh.go:135-156
inline void group_bounders_rec(Sorted3 &s, int32_t *out) {
  size_t n = s.a[0].size();
  if (n == 0) return;
  if (n <= 2) {
    for (size_t i = 0; i < n; i++) out[i] = s.a[0][i]->id;
    return;
  }
  size_t mid = n / 2;
  int axis = best_split_axis(s);
  Sorted3 lo, hi;
  split_bounders(s, axis, mid, lo, hi);
  for (int ax = 0; ax < 3; ax++) std::vector<Flagged *>().swap(s.a[ax]);  // free early
  group_bounders_rec(lo, out);
  group_bounders_rec(hi, out + mid);
}

// GroupTriangles (bvh.go:118-133, 219-255): returns the grouped order as ids.
inline std::vector<int32_t> group_triangles(const std::vector<Triangle> &tris) {
  size_t n = tris.size();
  std::vector<Flagged> flagged(n);
  for (size_t i = 0; i < n; i++) {
    flagged[i].id = (int32_t)i;
    flagged[i].mn = tris[i].mn();
    flagged[i].mx = tris[i].mx();
    flagged[i].md = mid(flagged[i].mn, flagged[i].mx);
  }
  Sorted3 s;
  for (int ax = 0; ax < 3; ax++) {
    s.a[ax].resize(n);
    for (size_t i = 0; i < n; i++) s.a[ax][i] = &flagged[i];
    std::stable_sort(s.a[ax].begin(), s.a[ax].end(),
                     [ax](const Flagged *x, const Flagged *y) { return x->md[ax] < y->md[ax]; });
  }
  std::vector<int32_t> out(n);
  group_bounders_rec(s, out.data());
  return out;
}

// ---- collisions.go:217-308 JoinedCollider ---------------------------------------
struct MeshCollider {
  struct Node {
    V3 mn, mx;
    // child >= 0: node index; child < 0: ~child is a position in `order`
    std::vector<int32_t> children;
  };
  std::vector<Triangle> tris;  // caller order (triangle id)
  std::vector<int32_t> order;  // grouped order
  std::vector<Node> nodes;
  int32_t root = -1;           // node index, or ~pos when the mesh has one triangle
  bool empty = true;
  bool interp = false;

  V3 child_min(int32_t c) const { return c >= 0 ? nodes[c].mn : tris[order[~c]].mn(); }
  V3 child_max(int32_t c) const { return c >= 0 ? nodes[c].mx : tris[order[~c]].mx(); }

  // collisions.go:169-179 + 225-253
  int32_t build_rec(size_t lo, size_t hi) {
    if (hi - lo == 1) return ~(int32_t)lo;
    size_t midi = lo + (hi - lo) / 2;
    int32_t c1 = build_rec(lo, midi);
    int32_t c2 = build_rec(midi, hi);
    Node nd;
    nd.mn = vmin(child_min(c1), child_min(c2));
    nd.mx = vmax(child_max(c1), child_max(c2));
    for (int32_t c : {c1, c2}) {
      if (c >= 0 && nodes[c].mn == nd.mn && nodes[c].mx == nd.mx) {
        // flatten joined colliders with identical bounds (collisions.go:235-250)
        nd.children.insert(nd.children.end(), nodes[c].children.begin(), nodes[c].children.end());
      } else {
        nd.children.push_back(c);
      }
    }
    nodes.push_back(std::move(nd));
    return (int32_t)nodes.size() - 1;
  }

  void build(bool grouped_already) {
    empty = tris.empty();
    if (empty) return;
    if (grouped_already) {
      order.resize(tris.size());
      for (size_t i = 0; i < tris.size(); i++) order[i] = (int32_t)i;
    } else {
      order = group_triangles(tris);
    }
    nodes.reserve(tris.size());
    root = build_rec(0, tris.size());
  }

  V3 mn() const { return empty ? V3() : child_min(root); }
  V3 mx() const { return empty ? V3() : child_max(root); }

  // collisions.go:275-290, 305-308
  bool first_rec(int32_t c, const Ray &r, Hit &out, Counters *cnt) const {
    if (c < 0) {
      if (cnt) cnt->tris++;
      int32_t id = order[~c];
      if (!tri_first_hit(tris[id], interp, r, out)) return false;
      out.prim = id;
      return true;
    }
    const Node &nd = nodes[c];
    if (cnt) cnt->nodes++;
    double a, b;
    ray_bounds(r, nd.mn, nd.mx, a, b);
    if (!(b >= a && b >= 0)) return false;
    bool any = false;
    Hit h;
    for (int32_t ch : nd.children) {
      if (first_rec(ch, r, h, cnt)) {
        if (h.scale < out.scale || !any) {
          out = h;
          any = true;
        }
      }
    }
    return any;
  }
  bool first_ray_collision(const Ray &r, Hit &out, Counters *cnt = nullptr) const {
    if (empty) return false;
    return first_rec(root, r, out, cnt);
  }

  // collisions.go:263-273 (RayCollisions): appends every hit
  void all_rec(int32_t c, const Ray &r, std::vector<Hit> &out) const {
    if (c < 0) {
      Hit h;
      int32_t id = order[~c];
      if (tri_first_hit(tris[id], interp, r, h)) {
        h.prim = id;
        out.push_back(h);
      }
      return;
    }
    const Node &nd = nodes[c];
    double a, b;
    ray_bounds(r, nd.mn, nd.mx, a, b);
    if (!(b >= a && b >= 0)) return;
    for (int32_t ch : nd.children) all_rec(ch, r, out);
  }
  void ray_collisions(const Ray &r, std::vector<Hit> &out) const {
    if (!empty) all_rec(root, r, out);
  }
  // brute force over all triangles (the comparison arm of TestMeshRayCollisions)
  void brute_collisions(const Ray &r, std::vector<Hit> &out) const {
    for (size_t i = 0; i < tris.size(); i++) {
      Hit h;
      if (tri_first_hit(tris[i], interp, r, h)) {
        h.prim = (int32_t)i;
        out.push_back(h);
      }
    }
  }
};

// ---- shapes.go -----------------------------------------------------------------
struct Sphere {
  V3 center;
  double radius = 1;
};
// shapes.go:35-93: first root with t >= 0, discriminant <= 0 misses.
inline bool sphere_first_hit(const Sphere &s, const Ray &r, Hit &h) {
  V3 o = sub(r.origin, s.center);
  V3 d = r.direction;
  double a = dot(d, d);
  double b = 2 * dot(d, o);
  double c = dot(o, o) - s.radius * s.radius;
  double disc = b * b - 4 * a * c;
  if (disc <= 0) return false;
  double sq = std::sqrt(disc);
  double t1 = (-b + sq) / (2 * a);
  double t2 = (-b - sq) / (2 * a);
  if (t1 > t2) std::swap(t1, t2);
  for (double t : {t1, t2}) {
    if (t < 0) continue;
    V3 point = add(r.origin, scale(r.direction, t));
    h.scale = t;
    h.normal = normalize(sub(point, s.center));
    h.prim = 0;
    return true;
  }
  return false;
}

struct Rect {
  V3 mn, mx;
};
// shapes.go:221-247
inline V3 rect_normal_at(const Rect &r, V3 c) {
  int axis = 0;
  double sign = 0;
  double min_dist = std::numeric_limits<double>::infinity();
  for (int i = 0; i < 3; i++) {
    double d = std::fabs(c[i] - r.mn[i]);
    if (d < min_dist) {
      min_dist = d;
      sign = -1;
      axis = i;
    }
    d = std::fabs(c[i] - r.mx[i]);
    if (d < min_dist) {
      min_dist = d;
      sign = 1;
      axis = i;
    }
  }
  V3 res;
  res.at(axis) = sign;
  return res;
}
// shapes.go:177-196
inline bool rect_first_hit(const Rect &rc, const Ray &r, Hit &h) {
  double tmin, tmax;
  ray_bounds(r, rc.mn, rc.mx, tmin, tmax);
  if (tmax < tmin || tmax < 0) return false;
  double t = tmin;
  if (t < 0) t = tmax;
  h.scale = t;
  h.normal = rect_normal_at(rc, add(r.origin, scale(r.direction, t)));
  h.prim = 0;
  return true;
}

struct Cylinder {
  V3 p1, p2;
  double radius = 1;
};
// shapes.go:832-856
inline bool cast_plane(V3 normal, double bias, const Ray &r, double &t) {
  double ddot = dot(r.direction, normal);
  if (std::fabs(ddot) < 1e-8 * norm(r.direction) * norm(normal)) return false;
  t = (bias - dot(r.origin, normal)) / ddot;
  return !(t < 0);
}
// shapes.go:816-828
inline bool cast_circle(V3 normal, V3 center, double radius, const Ray &r, double &t) {
  double bias = dot(normal, center);
  if (!cast_plane(normal, bias, r, t)) return false;
  V3 p = add(r.origin, scale(r.direction, t));
  return !(dist(p, center) > radius);
}
// shapes.go:601-705: min over all reported collisions (first strictly-smaller wins)
inline bool cylinder_first_hit(const Cylinder &c, const Ray &r, Hit &h) {
  bool ok = false;
  auto report = [&](double t, V3 n) {
    if (!ok || t < h.scale) {
      h.scale = t;
      h.normal = n;
      h.prim = 0;
      ok = true;
    }
  };
  V3 v = normalize(sub(c.p2, c.p1));
  V3 o = sub(r.origin, c.p1);
  V3 d = r.direction;
  V3 v1 = sub(scale(v, dot(o, v)), o);
  V3 v2 = sub(scale(v, dot(d, v)), d);
  double a = dot(v2, v2);
  double b = 2 * dot(v1, v2);
  double cv = dot(v1, v1) - c.radius * c.radius;
  double disc = b * b - 4 * a * cv;
  if (disc > 0) {
    double sq = std::sqrt(disc);
    double max_scale = norm(sub(c.p2, c.p1));
    for (double sign : {-1.0, 1.0}) {
      double t = (-b + sign * sq) / (2 * a);
      if (t < 0) continue;
      V3 p = add(o, scale(d, t));
      double frac = dot(v, p);
      if (frac >= 0 && frac < max_scale) report(t, normalize(sub(p, scale(v, frac))));
    }
  }
  for (int i = 0; i < 2; i++) {
    V3 tip = i == 0 ? c.p1 : c.p2;
    V3 n = i == 0 ? scale(v, -1) : v;
    double t;
    if (cast_circle(n, tip, c.radius, r, t)) report(t, n);
  }
  return ok;
}

}  // namespace orc
