// ORACLE -- TEST INFRASTRUCTURE ONLY (see vec.hpp header).
//
// Restatement of the reference mesh generators used to synthesise the benchmark
// and test meshes, with insertion order as the (deterministic) triangle order:
//   GeoCoord.Coord3D       model3d/coords.go:32-38
//   NewMeshPolar           model3d/mesh.go:62-118
//   NewMeshIcosahedron     model3d/mesh.go:304-330
//   NewMeshIcosphere       model3d/mesh.go:124-128
//   SubdivideEdges         model3d/subdivision.go:88-139
//   NewMeshRect / AddQuad  model3d/mesh.go:132-165, 389-397
#pragma once
#include <cmath>
#include <vector>

#include "collide.hpp"

namespace orc {

inline V3 geo(double lat, double lon) {
  return {std::sin(lon) * std::cos(lat), std::sin(lat), std::cos(lon) * std::cos(lat)};
}

inline Triangle mk_tri(V3 a, V3 b, V3 c) {
  Triangle t;
  t.p[0] = a;
  t.p[1] = b;
  t.p[2] = c;
  return t;
}

// radius(g) = ra + rb*cos(lon): covers the reference tests' unit sphere (1,0) and
// the "0.5 + 0.1*cos(lon)" shape of TestMeshRayCollisions (collisions_test.go:24-26).
inline std::vector<Triangle> mesh_polar(double ra, double rb, int stops) {
  auto radius = [&](double lat, double lon) {
    (void)lat;
    return ra + rb * std::cos(lon);
  };
  std::vector<Triangle> res;
  double lon_step = M_PI * 2 / stops, lat_step = M_PI / stops;
  auto lat_f = [&](int i) { return -M_PI / 2 + i * lat_step; };
  auto lon_f = [&](int i) { return i == stops ? -M_PI : -M_PI + i * lon_step; };
  for (int lo = 0; lo < stops; lo++) {
    for (int la = 0; la < stops; la++) {
      double lon = lon_f(lo), lat = lat_f(la), lon_n = lon_f(lo + 1), lat_n = lat_f(la + 1);
      double g[4][2] = {{lat, lon}, {lat, lon_n}, {lat_n, lon_n}, {lat_n, lon}};
      V3 p[4];
      for (int i = 0; i < 4; i++) p[i] = scale(geo(g[i][0], g[i][1]), radius(g[i][0], g[i][1]));
      if (la == 0)
        p[0] = V3(0, -radius(lat, 0), 0);
      else if (la == stops - 1)
        p[2] = V3(0, radius(lat, 0), 0);
      if (la != 0) res.push_back(mk_tri(p[0], p[1], p[2]));
      if (la != stops - 1) res.push_back(mk_tri(p[0], p[2], p[3]));
    }
  }
  return res;
}

inline std::vector<Triangle> mesh_icosahedron() {
  std::vector<Triangle> m;
  double mid_lat = std::atan(0.5);
  auto top_c = [&](int i) { return geo(-mid_lat, M_PI * 2 * double(i % 5) / 5.0); };
  auto bot_c = [&](int i) { return geo(mid_lat, M_PI * 2 * (1.0 / 10.0 + double(i % 5) / 5.0)); };
  V3 top = geo(-M_PI / 2, 0), bottom = geo(M_PI / 2, 0);
  for (int i = 0; i < 5; i++) {
    m.push_back(mk_tri(top, top_c(i + 1), top_c(i)));
    m.push_back(mk_tri(bottom, bot_c(i), bot_c(i + 1)));
    m.push_back(mk_tri(top_c(i), top_c(i + 1), bot_c(i)));
    m.push_back(mk_tri(bot_c(i + 1), bot_c(i), top_c(i + 1)));
  }
  return m;
}

// primitives.go:547-554 canonical segment ordering
inline bool seg_first_is(V3 p1, V3 p2) {
  return p1.x < p2.x || (p1.x == p2.x && p1.y < p2.y) || (p1.x == p2.x && p1.y == p2.y && p1.z < p2.z);
}

// subdivision.go:117-139
inline void divide_segment(V3 c1, V3 c2, V3 *result, int len) {
  if (len == 1) {
    result[0] = c1;
    return;
  }
  if (!seg_first_is(c1, c2)) {
    // NewSegment would have put c2 first (also when c1 == c2: the else branch
    // of NewSegment yields {p2, p1}, whose [0] equals c1 bit-for-bit, so the
    // reference does not swap in that case).
    if (!(c1 == c2)) {
      divide_segment(c2, c1, result, len);
      for (int i = 0; i < len / 2; i++) std::swap(result[i], result[len - (i + 1)]);
      return;
    }
  }
  result[0] = c1;
  result[len - 1] = c2;
  for (int i = 1; i + 1 < len; i++) {
    double t = double(i) / double(len - 1);
    result[i] = add(scale(c1, 1 - t), scale(c2, t));
  }
}

// subdivision.go:88-115
inline std::vector<Triangle> subdivide_edges(const std::vector<Triangle> &m, int n) {
  std::vector<V3> side1(n + 1), side2(n + 1), wide(n + 1), narrow(n);
  std::vector<Triangle> res;
  res.reserve(m.size() * (size_t)n * n);
  for (const Triangle &t : m) {
    divide_segment(t.p[0], t.p[1], side1.data(), n + 1);
    divide_segment(t.p[0], t.p[2], side2.data(), n + 1);
    for (int i = 0; i < n; i++) {
      int wl = i + 2, nl = i + 1;
      divide_segment(side1[i], side2[i], narrow.data(), nl);
      divide_segment(side1[i + 1], side2[i + 1], wide.data(), wl);
      for (int k = 0; k < nl; k++) {
        res.push_back(mk_tri(narrow[k], wide[k], wide[k + 1]));
        if (k > 0) res.push_back(mk_tri(narrow[k], narrow[k - 1], wide[k]));
      }
    }
  }
  return res;
}

// mesh.go:124-128: SubdivideEdges(icosahedron, n) -> Normalize -> Scale -> Translate
inline std::vector<Triangle> mesh_icosphere(V3 center, double radius, int n) {
  std::vector<Triangle> m = subdivide_edges(mesh_icosahedron(), n);
  for (Triangle &t : m)
    for (int i = 0; i < 3; i++) {
      V3 c = normalize(t.p[i]);
      c = mul(V3(radius, radius, radius), c);  // Mesh.Scale: XYZ(s,s,s).Mul
      t.p[i] = add(center, c);                 // Mesh.Translate: v.Add
    }
  return m;
}

// mesh.go:132-165
inline std::vector<Triangle> mesh_rect(V3 mn, V3 mx) {
  std::vector<Triangle> m;
  auto point = [&](int x, int y, int z) {
    V3 r = mn;
    if (x == 1) r.x = mx.x;
    if (y == 1) r.y = mx.y;
    if (z == 1) r.z = mx.z;
    return r;
  };
  auto quad = [&](V3 p1, V3 p2, V3 p3, V3 p4) {
    m.push_back(mk_tri(p1, p2, p4));
    m.push_back(mk_tri(p2, p3, p4));
  };
  quad(mn, point(1, 0, 0), point(1, 0, 1), point(0, 0, 1));
  quad(mx, point(1, 1, 0), point(0, 1, 0), point(0, 1, 1));
  quad(mn, point(0, 0, 1), point(0, 1, 1), point(0, 1, 0));
  quad(mx, point(1, 0, 1), point(1, 0, 0), point(1, 1, 0));
  quad(mn, point(0, 1, 0), point(1, 1, 0), point(1, 0, 0));
  quad(mx, point(0, 1, 1), point(0, 0, 1), point(1, 0, 1));
  return m;
}

}  // namespace orc
