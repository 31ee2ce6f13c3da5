// ORACLE -- TEST INFRASTRUCTURE ONLY (see vec.hpp header).  "parity unpinned" against
// reference outputs (no Go toolchain); pinned by restated reference properties
// (tests/test_oracle_sdf.py).
//
// float64 CPU restatement of the reference's nearest-triangle queries on a mesh:
//   Triangle.Closest                 model3d/primitives.go:153-175
//   NewSegment / Segment.Closest     model3d/primitives.go:547-554, 579-593
//   Triangle.SphereCollision         model3d/primitives.go:253-279
//   JoinedCollider.SphereCollision   model3d/collisions.go:292-303
//   sphereTouchesBounds / pointToBoundsDistSquared  model3d/bvh.go:302-320
//   ColliderContains (with margin)   model3d/collisions.go:119-134
//   InBounds                         model3d/bounder.go:23-27
//   ColliderSolid.Contains           model3d/solid.go:292-300
//   MeshToSDF / GroupedTrianglesToSDF / meshSDF.{SDF,PointSDF,NormalSDF,FaceSDF}
//                                    model3d/sdf.go:186-240
//   newMeshDistFunc / meshDistFunc.Dist  model3d/sdf.go:242-311
#pragma once
#include <limits>
#include <vector>

#include "collide.hpp"

namespace orc {

// primitives.go:547-554 + 579-593
inline V3 segment_closest(V3 p1, V3 p2, V3 c) {
  V3 s0 = p1, s1 = p2;
  if (!(p1.x < p2.x || (p1.x == p2.x && p1.y < p2.y) || (p1.x == p2.x && p1.y == p2.y && p1.z < p2.z))) {
    s0 = p2;
    s1 = p1;
  }
  V3 v1 = sub(s1, s0);
  double nrm = norm(v1);
  V3 v = scale(v1, 1 / nrm);
  V3 v2 = sub(c, s0);
  double mag = dot(v, v2);
  if (mag > nrm) return s1;
  if (mag < 0) return s0;
  return add(scale(v, mag), s0);
}

// primitives.go:153-175
inline V3 tri_closest(const Triangle &t, V3 c) {
  V3 v1 = sub(t.p[1], t.p[0]);
  V3 v2 = sub(t.p[2], t.p[0]);
  V3 n = t.normal();
  M3 mat;  // NewMatrix3Columns(v1, v2, normal) (matrix.go:16-22, row-major storage)
  mat.m[0] = v1.x, mat.m[1] = v2.x, mat.m[2] = n.x;
  mat.m[3] = v1.y, mat.m[4] = v2.y, mat.m[5] = n.y;
  mat.m[6] = v1.z, mat.m[7] = v2.z, mat.m[8] = n.z;
  V3 comp = mul_column(inverse(mat), sub(c, t.p[0]));
  if (comp.x >= 0 && comp.y >= 0 && comp.x + comp.y <= 1)
    return add(add(t.p[0], scale(v1, comp.x)), scale(v2, comp.y));
  double best = std::numeric_limits<double>::infinity();
  V3 best_p;
  for (int i = 0; i < 3; i++) {
    V3 c1 = segment_closest(t.p[i], t.p[(i + 1) % 3], c);
    double d = dist(c1, c);
    if (d < best) {
      best = d;
      best_p = c1;
    }
  }
  return best_p;
}

// bvh.go:306-320
inline double point_bounds_dist2(V3 c, V3 mn, V3 mx) {
  double d2 = 0;
  for (int a = 0; a < 3; a++) {
    double lo = mn[a], hi = mx[a], v = c[a];
    if (v < lo)
      d2 += (lo - v) * (lo - v);
    else if (v > hi)
      d2 += (hi - v) * (hi - v);
  }
  return d2;
}

// primitives.go:253-279
inline bool tri_sphere_collision(const Triangle &t, V3 c, double r) {
  for (int i = 0; i < 3; i++)
    if (dist(t.p[i], c) < r) return true;
  for (int i = 0; i < 3; i++) {
    V3 p1 = t.p[i], p2 = t.p[(i + 1) % 3];
    V3 v = sub(p2, p1);
    double frac = (dot(c, v) - dot(p1, v)) / dot(v, v);
    V3 closest = add(p1, scale(v, frac));
    if (frac >= 0 && frac <= 1 && dist(closest, c) < r) return true;
  }
  Ray ray{c, t.normal()};
  double sc, bary[3];
  return tri_ray_collision(t, ray, sc, bary) && std::fabs(sc) < r;
}

// collisions.go:292-303 over the MeshCollider's joined tree
inline bool sphere_collision_rec(const MeshCollider &m, int32_t c, V3 center, double r) {
  if (c < 0) return tri_sphere_collision(m.tris[m.order[~c]], center, r);
  const MeshCollider::Node &nd = m.nodes[c];
  if (!(point_bounds_dist2(center, nd.mn, nd.mx) <= r * r)) return false;
  for (int32_t ch : nd.children)
    if (sphere_collision_rec(m, ch, center, r)) return true;
  return false;
}
inline bool sphere_collision(const MeshCollider &m, V3 center, double r) {
  return !m.empty && sphere_collision_rec(m, m.root, center, r);
}

// collisions.go:119-134
inline bool collider_contains(const MeshCollider &m, V3 p, double margin) {
  Ray r{p, V3(0.5224892708603626, 0.10494477243214506, 0.43558938446126527)};
  std::vector<Hit> hits;
  m.ray_collisions(r, hits);
  if (hits.size() % 2 == 0) {
    if (margin < 0) return sphere_collision(m, p, -margin);
    return false;
  }
  return margin <= 0 || !sphere_collision(m, p, margin);
}

// solid.go:292-300 for NewColliderSolid(c) (inset 0, radius 0)
inline bool collider_solid_contains(const MeshCollider &m, V3 p) {
  if (m.empty) return false;
  V3 mn = m.mn(), mx = m.mx();
  if (!(vmin(p, mn) == mn && vmax(p, mx) == mx)) return false;  // bounder.go:23-27
  return collider_contains(m, p, 0);
}

// sdf.go:242-311: binary tree over the grouped triangle list (halves at len/2).
struct MeshDistFunc {
  struct Node {
    V3 mn, mx;
    int32_t tri = -1;  // leaf: triangle id
    int32_t child[2] = {-1, -1};
  };
  const MeshCollider *mesh = nullptr;
  std::vector<Node> nodes;
  int32_t root = -1;

  int32_t build_rec(size_t lo, size_t hi) {
    Node nd;
    if (hi - lo == 1) {
      nd.tri = mesh->order[lo];
      nd.mn = mesh->tris[nd.tri].mn();
      nd.mx = mesh->tris[nd.tri].mx();
    } else {
      size_t midi = lo + (hi - lo) / 2;
      nd.child[0] = build_rec(lo, midi);
      nd.child[1] = build_rec(midi, hi);
      nd.mn = vmin(nodes[nd.child[0]].mn, nodes[nd.child[1]].mn);
      nd.mx = vmax(nodes[nd.child[0]].mx, nodes[nd.child[1]].mx);
    }
    nodes.push_back(nd);
    return (int32_t)nodes.size() - 1;
  }
  explicit MeshDistFunc(const MeshCollider &m) : mesh(&m) {
    nodes.reserve(2 * m.tris.size());
    if (!m.empty) root = build_rec(0, m.tris.size());
  }

  void dist_rec(int32_t n, V3 c, double &cur, V3 &cur_point, int32_t &cur_face) const {
    const Node &nd = nodes[n];
    if (nd.tri >= 0) {
      V3 cp = tri_closest(mesh->tris[nd.tri], c);
      double d = dist(cp, c);
      if (d < cur) {
        cur = d;
        cur_point = cp;
        cur_face = nd.tri;
      }
      return;
    }
    double bd[2] = {point_bounds_dist2(c, nodes[nd.child[0]].mn, nodes[nd.child[0]].mx),
                    point_bounds_dist2(c, nodes[nd.child[1]].mn, nodes[nd.child[1]].mx)};
    int32_t it[2] = {nd.child[0], nd.child[1]};
    if (bd[0] > bd[1]) {
      std::swap(it[0], it[1]);
      std::swap(bd[0], bd[1]);
    }
    for (int i = 0; i < 2; i++) {
      if (bd[i] > cur * cur) continue;
      dist_rec(it[i], c, cur, cur_point, cur_face);
    }
  }

  // meshSDF.FaceSDF (sdf.go:229-240): signed distance (positive inside), nearest point, face
  double face_sdf(V3 c, V3 &point, int32_t &face) const {
    double d = std::numeric_limits<double>::infinity();
    point = V3();
    face = -1;
    dist_rec(root, c, d, point, face);
    if (!collider_solid_contains(*mesh, c)) d = -d;
    return d;
  }
};

}  // namespace orc
