// ORACLE -- TEST INFRASTRUCTURE ONLY (see vec.hpp header).
//
// Restatement of the reference's marching cubes for the C1 config ("model3d.Sphere ->
// MarchingCubesSearch(0.01, 8)").  Follows, in float64:
//   MarchingCubes            model3d/mc.go:14-37   (scan z, y, x; table[bits] triangles)
//   MarchingCubesSearch      model3d/mc.go:45-51
//   mcSearch / mcSearchPoint model3d/mc.go:182-260 (bisection of every vertex along its edge)
//   mcCornerCoordinates      model3d/mc.go:289-301
//   allMcRotations/Compose   model3d/mc.go:312-361 (closure of the z/x generators, sorted)
//   ApplyTriangle/ApplyIntersections  mc.go:369-388
//   mcTriangle.Triangle      model3d/mc.go:400-406 (edge midpoints, Coord3D.Mid = (a+b)*0.5)
//   mcLookupTable            model3d/mc.go:431-454 (first rotation in sorted order wins)
//   baseTriangleTable        model3d/mc.go:460-586 (the 23 base cases, data)
//   newSquareSpacer          model3d/mc.go:594-612 (x = min-delta; x <= max+delta; x += delta)
//   LookupEdgePoint          model3d/mc.go:647-660
//   solidCache.GetSquare     model3d/mc.go:697-710 (bit = x + 2*y (+4 for the top layer))
//   Sphere.Contains          model3d/shapes.go:29-31 (Dist(center) <= radius)
//
// Determinism: the reference stores triangles in a Go map (random iteration order); here the
// triangle order is the scan order (z outer, then y, x, then the table's triangle order).
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <functional>
#include <map>
#include <set>
#include <stdexcept>
#include <vector>

#include "meshgen.hpp"

namespace orc {

using McRot = std::array<uint8_t, 8>;
using McTri = std::array<uint8_t, 6>;

struct McBase {
  std::vector<int> inside;
  std::vector<McTri> tris;
};

// mc.go:460-586 (data table).
inline const std::vector<McBase> &mc_base_table() {
  static const std::vector<McBase> t = {
      {{}, {}},
      {{0}, {{0, 1, 0, 2, 0, 4}}},
      {{0, 1}, {{0, 4, 1, 5, 0, 2}, {1, 5, 1, 3, 0, 2}}},
      {{0, 5}, {{0, 1, 0, 2, 0, 4}, {5, 7, 1, 5, 4, 5}}},
      {{0, 7}, {{0, 1, 0, 2, 0, 4}, {6, 7, 3, 7, 5, 7}}},
      {{1, 2, 3}, {{0, 1, 1, 5, 0, 2}, {0, 2, 1, 5, 2, 6}, {2, 6, 1, 5, 3, 7}}},
      {{0, 1, 7}, {{0, 4, 1, 5, 0, 2}, {1, 5, 1, 3, 0, 2}, {6, 7, 3, 7, 5, 7}}},
      {{1, 4, 7}, {{4, 6, 4, 5, 0, 4}, {1, 5, 1, 3, 0, 1}, {6, 7, 3, 7, 5, 7}}},
      {{0, 1, 2, 3}, {{0, 4, 1, 5, 3, 7}, {0, 4, 3, 7, 2, 6}}},
      {{0, 2, 3, 6}, {{0, 1, 4, 6, 0, 4}, {0, 1, 6, 7, 4, 6}, {0, 1, 1, 3, 6, 7}, {1, 3, 3, 7, 6, 7}}},
      {{1, 2, 5, 6}, {{0, 2, 2, 3, 6, 7}, {0, 2, 6, 7, 4, 6}, {0, 1, 4, 5, 5, 7}, {5, 7, 1, 3, 0, 1}}},
      {{0, 2, 3, 7}, {{0, 4, 0, 1, 2, 6}, {0, 1, 5, 7, 2, 6}, {2, 6, 5, 7, 6, 7}, {0, 1, 1, 3, 5, 7}}},
      {{1, 2, 3, 4}, {{0, 1, 1, 5, 0, 2}, {0, 2, 1, 5, 2, 6}, {2, 6, 1, 5, 3, 7}, {4, 5, 0, 4, 4, 6}}},
      {{1, 2, 4, 7}, {{0, 1, 1, 5, 1, 3}, {0, 2, 2, 3, 2, 6}, {4, 5, 0, 4, 4, 6}, {5, 7, 6, 7, 3, 7}}},
      {{1, 2, 3, 6}, {{0, 2, 0, 1, 4, 6}, {0, 1, 3, 7, 4, 6}, {0, 1, 1, 5, 3, 7}, {4, 6, 3, 7, 6, 7}}},
      {{0, 2, 3, 5, 6},
       {{0, 1, 4, 6, 0, 4}, {0, 1, 6, 7, 4, 6}, {0, 1, 1, 3, 6, 7}, {1, 3, 3, 7, 6, 7}, {5, 7, 1, 5, 4, 5}}},
      {{2, 3, 4, 5, 6},
       {{5, 7, 1, 5, 0, 4}, {0, 4, 6, 7, 5, 7}, {0, 2, 6, 7, 0, 4}, {0, 2, 3, 7, 6, 7}, {0, 2, 1, 3, 3, 7}}},
      {{0, 4, 5, 6, 7}, {{1, 5, 0, 1, 0, 2}, {0, 2, 2, 6, 1, 5}, {1, 5, 2, 6, 3, 7}}},
      {{1, 2, 3, 4, 5, 6}, {{0, 2, 0, 1, 0, 4}, {3, 7, 6, 7, 5, 7}}},
      {{1, 2, 3, 4, 6, 7}, {{0, 2, 4, 5, 0, 4}, {0, 2, 5, 7, 4, 5}, {0, 2, 1, 5, 5, 7}, {0, 1, 1, 5, 0, 2}}},
      {{2, 3, 4, 5, 6, 7}, {{1, 5, 0, 4, 0, 2}, {1, 3, 1, 5, 0, 2}}},
      {{1, 2, 3, 4, 5, 6, 7}, {{0, 2, 0, 1, 0, 4}}},
      {{0, 1, 2, 3, 4, 5, 6, 7}, {}},
  };
  return t;
}

// mc.go:312-361
inline std::vector<McRot> mc_rotations() {
  auto compose = [](const McRot &m, const McRot &m1) {
    McRot r;
    for (int i = 0; i < 8; i++) r[i] = m[m1[i]];
    return r;
  };
  const McRot zr = {2, 0, 3, 1, 6, 4, 7, 5}, xr = {2, 3, 6, 7, 0, 1, 4, 5};
  std::vector<McRot> queue = {McRot{0, 1, 2, 3, 4, 5, 6, 7}};
  std::set<McRot> seen = {queue[0]};
  for (size_t head = 0; head < queue.size(); head++) {
    McRot next = queue[head];
    for (const McRot &op : {zr, xr}) {
      McRot r = compose(op, next);
      if (seen.insert(r).second) queue.push_back(r);
    }
  }
  return std::vector<McRot>(seen.begin(), seen.end());  // std::set order == the lexicographic sort
}

// mc.go:431-454.  The reference walks a Go map of base cases (random order); the result does
// not depend on that order as long as no two base cases share a rotation class, which is
// asserted here.
inline const std::array<std::vector<McTri>, 256> &mc_lookup_table() {
  static std::array<std::vector<McTri>, 256> table;
  static bool ready = false;
  if (ready) return table;
  std::vector<McRot> rots = mc_rotations();
  if (rots.size() != 24) throw std::runtime_error("expected 24 cube rotations");
  std::array<int, 256> owner;
  owner.fill(-1);
  const auto &base = mc_base_table();
  for (size_t b = 0; b < base.size(); b++) {
    unsigned bits = 0;
    for (int c : base[b].inside) bits |= 1u << c;
    for (const McRot &rot : rots) {
      unsigned nb = 0;
      for (int c = 0; c < 8; c++)
        if (bits & (1u << c)) nb |= 1u << rot[c];
      if (owner[nb] >= 0) {
        if (owner[nb] != (int)b) throw std::runtime_error("two base cases in one rotation class");
        continue;
      }
      owner[nb] = (int)b;
      for (const McTri &t : base[b].tris) {
        McTri r;
        for (int i = 0; i < 6; i++) r[i] = rot[t[i]];
        table[nb].push_back(r);
      }
    }
  }
  for (int i = 0; i < 256; i++)
    if (owner[i] < 0) throw std::runtime_error("marching cubes table incomplete");
  ready = true;
  return table;
}

struct McSpacer {
  std::vector<double> xs, ys, zs;
};

// mc.go:594-612
inline McSpacer mc_spacer(V3 mn, V3 mx, double delta) {
  McSpacer s;
  for (double x = mn.x - delta; x <= mx.x + delta; x += delta) s.xs.push_back(x);
  for (double y = mn.y - delta; y <= mx.y + delta; y += delta) s.ys.push_back(y);
  for (double z = mn.z - delta; z <= mx.z + delta; z += delta) s.zs.push_back(z);
  return s;
}

using SolidFn = std::function<bool(V3)>;

// mc.go:14-37
inline std::vector<Triangle> marching_cubes(const SolidFn &contains, V3 mn, V3 mx, double delta) {
  const auto &table = mc_lookup_table();
  McSpacer sp = mc_spacer(mn, mx, delta);
  size_t nx = sp.xs.size(), ny = sp.ys.size(), nz = sp.zs.size();
  std::vector<uint8_t> bottom(nx * ny), top(nx * ny);
  auto fetch = [&](size_t z, std::vector<uint8_t> &dst) {
    bool on_edge = z == 0 || z == nz - 1;
    for (size_t i = 0; i < ny; i++)
      for (size_t j = 0; j < nx; j++) {
        bool b = contains(V3(sp.xs[j], sp.ys[i], sp.zs[z]));
        dst[j + i * nx] = b;
        if (b && (on_edge || i == 0 || j == 0 || i == ny - 1 || j == nx - 1))
          throw std::runtime_error("solid is true outside of bounds");
      }
  };
  auto square = [&](const std::vector<uint8_t> &c, size_t x, size_t y) {
    unsigned r = 0, mask = 1;
    for (size_t y1 = y; y1 < y + 2; y1++)
      for (size_t x1 = x; x1 < x + 2; x1++) {
        if (c[x1 + y1 * nx]) r |= mask;
        mask <<= 1;
      }
    return r;
  };
  auto mid = [](V3 a, V3 b) { return scale(add(a, b), 0.5); };  // coords.go Mid
  std::vector<Triangle> mesh;
  fetch(0, bottom);
  for (size_t z = 1; z < nz; z++) {
    fetch(z, top);
    for (size_t y = 0; y + 1 < ny; y++)
      for (size_t x = 0; x + 1 < nx; x++) {
        unsigned bits = square(bottom, x, y) | (square(top, x, y) << 4);
        const auto &tris = table[bits];
        if (tris.empty()) continue;
        V3 lo(sp.xs[x], sp.ys[y], sp.zs[z - 1]), hi(sp.xs[x + 1], sp.ys[y + 1], sp.zs[z]);
        V3 corners[8] = {lo, V3(hi.x, lo.y, lo.z), V3(lo.x, hi.y, lo.z), V3(hi.x, hi.y, lo.z),
                         V3(lo.x, lo.y, hi.z), V3(hi.x, lo.y, hi.z), V3(lo.x, hi.y, hi.z), hi};
        for (const McTri &t : tris)
          mesh.push_back(mk_tri(mid(corners[t[0]], corners[t[1]]), mid(corners[t[2]], corners[t[3]]),
                                mid(corners[t[4]], corners[t[5]])));
      }
    bottom.swap(top);
  }
  return mesh;
}

// mc.go:232-260 + 647-660.  The search result is a function of the vertex alone, so moving
// every triangle corner independently equals the reference's per-unique-vertex update.
inline V3 mc_search_point(const SolidFn &contains, int iters, const McSpacer &sp, V3 c) {
  double arr[3] = {c.x, c.y, c.z};
  const double origin[3] = {sp.xs[0], sp.ys[0], sp.zs[0]};
  const std::vector<double> *vals[3] = {&sp.xs, &sp.ys, &sp.zs};
  double delta = sp.xs[1] - sp.xs[0];
  int axis = -1;
  double fp = 0, tp = 0;
  for (int i = 0; i < 3; i++) {
    double modulus = std::fabs(std::fmod(arr[i] - origin[i], delta));
    if (modulus > delta / 4 && modulus < 3 * delta / 4) {
      int idx = (int)((arr[i] - origin[i]) / delta);
      axis = i;
      fp = (*vals[i])[idx];
      tp = (*vals[i])[idx + 1];
      break;
    }
  }
  if (axis < 0) throw std::runtime_error("vertex not on edge");
  double probe[3] = {arr[0], arr[1], arr[2]};
  probe[axis] = tp;
  if (!contains(V3(probe[0], probe[1], probe[2]))) std::swap(fp, tp);
  for (int i = 0; i < iters; i++) {
    double m = (fp + tp) / 2;
    arr[axis] = m;
    if (contains(V3(arr[0], arr[1], arr[2])))
      tp = m;
    else
      fp = m;
  }
  arr[axis] = (fp + tp) / 2;
  return V3(arr[0], arr[1], arr[2]);
}

// mc.go:45-51
inline std::vector<Triangle> marching_cubes_search(const SolidFn &contains, V3 mn, V3 mx, double delta,
                                                   int iters) {
  std::vector<Triangle> mesh = marching_cubes(contains, mn, mx, delta);
  if (iters == 0) return mesh;
  McSpacer sp = mc_spacer(mn, mx, delta);
  for (Triangle &t : mesh)
    for (int k = 0; k < 3; k++) t.p[k] = mc_search_point(contains, iters, sp, t.p[k]);
  return mesh;
}

// Sphere as a Solid: shapes.go:17-31 (Min/Max = center -+ radius; Contains = Dist <= r).
inline std::vector<Triangle> marching_cubes_sphere(V3 center, double radius, double delta, int iters) {
  SolidFn f = [=](V3 p) { return norm(sub(p, center)) <= radius; };
  V3 r(radius, radius, radius);
  return marching_cubes_search(f, sub(center, r), add(center, r), delta, iters);
}

}  // namespace orc
