// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// float64 CPU restatement of the vector arithmetic of the reference
// (model3d/coords.go:195-434).  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this code.
//
// Parity status: the reference is Go and no Go toolchain exists in the build container, so
// this restatement cannot be run against the reference binary.  It is pinned two ways:
//  (1) against OUTPUTS OF THE REFERENCE ITSELF: the renderings the Go code produced and the
//      reference commits (examples/renderings/cornell_box/output.png and output_hd.png, copied to
//      tests/golden/): the oracle's path tracer reproduces them statistically -- z-scores of mean
//      0.04 and rms 1.05 over the 200x200 image, image means equal to 0.04 %
//      (tests/test_reference_golden.py).  That scene exercises mesh BVH first hits, analytic
//      spheres, every material kind, the focus point and the recursive tracer.  Likewise
//      examples/renderings/showcase/output.png (BASELINE config 4; block means within 1.2 % median
//      with the vase, whose mesh is missing, masked) and -- deterministically --
//      examples/renderings/smooth_shading/rendering.png: the RayCaster image of two icospheres
//      (flat and interpolated normals, Phong material, point light, 4x downsampling), which the
//      oracle reproduces with identical silhouettes and 99.994 % of the 82,696 lit pixels equal
//      at 8 bits.  That image fixes, per pixel, which triangle is hit first and its normal;
//  (2) by restating the reference's own property tests (tests/test_oracle_*.py,
//      tests/test_marching_cubes.py); the reference ships no golden vectors for this path
//      (SURVEY.md section 8c).
// Deterministic per-ray outputs as numbers (triangle ids, t) have no reference output to compare
// with beyond what the deterministic image in (1) implies: for those the oracle remains "parity
// unpinned" beyond (1) and (2).
#pragma once
#include <cmath>
#include <cstdint>

namespace orc {

struct V3 {
  double x = 0, y = 0, z = 0;
  V3() = default;
  V3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
  double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  double &at(int i) { return i == 0 ? x : (i == 1 ? y : z); }
  bool operator==(const V3 &o) const { return x == o.x && y == o.y && z == o.z; }
  bool operator!=(const V3 &o) const { return !(*this == o); }
};

inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 scale(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 mul(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// coords.go Cross
inline V3 cross(V3 a, V3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }
inline double dist(V3 a, V3 b) { return norm(sub(a, b)); }
// coords.go:379-381: Normalize = Scale(1/Norm) (not a per-component division)
inline V3 normalize(V3 a) { return scale(a, 1.0 / norm(a)); }
inline V3 vmin(V3 a, V3 b) { return {std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)}; }
inline V3 vmax(V3 a, V3 b) { return {std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)}; }
inline V3 mid(V3 a, V3 b) { return scale(add(a, b), 0.5); }
inline double sum(V3 a) { return a.x + a.y + a.z; }
inline double maxcoord(V3 a) { return std::fmax(a.x, std::fmax(a.y, a.z)); }

// coords.go:424-427
inline V3 project_out(V3 c, V3 c1) {
  V3 n = normalize(c1);
  return sub(c, scale(n, dot(n, c)));
}

// coords.go:431-434: reflect c1 around c
inline V3 reflect(V3 c, V3 c1) {
  V3 n = normalize(c);
  return scale(add(c1, scale(n, -2 * dot(n, c1))), -1);
}

// coords.go:388-421
inline void ortho_basis(V3 c, V3 &b1o, V3 &b2o) {
  double ax = std::fabs(c.x), ay = std::fabs(c.y), az = std::fabs(c.z);
  V3 b1;
  if (ax > ay && ax > az) {
    b1.x = c.y / ax;
    b1.y = -c.x / ax;
  } else {
    b1.y = c.z;
    b1.z = -c.y;
    if (ay > az) {
      b1.y /= ay;
      b1.z /= ay;
    } else {
      b1.y /= az;
      b1.z /= az;
    }
  }
  V3 b2{b1.y * c.z - b1.z * c.y, b1.z * c.x - b1.x * c.z, b1.x * c.y - b1.y * c.x};
  b1o = normalize(b1);
  b2o = normalize(b2);
}

// model3d/matrix.go:11-12 Matrix3 is row-major; MulColumn matrix.go:131-137.
struct M3 {
  double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
};
inline V3 mul_column(const M3 &a, V3 c) {
  return {a.m[0] * c.x + a.m[1] * c.y + a.m[2] * c.z,
          a.m[3] * c.x + a.m[4] * c.y + a.m[5] * c.z,
          a.m[6] * c.x + a.m[7] * c.y + a.m[8] * c.z};
}
// matrix.go:57-59, 83-90 (adjugate / det)
inline double det(const M3 &a) {
  const double *m = a.m;
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
         m[2] * (m[3] * m[7] - m[4] * m[6]);
}
inline M3 inverse(const M3 &a) {
  const double *m = a.m;
  double d = 1.0 / det(a);
  M3 r;
  double adj[9] = {m[4] * m[8] - m[5] * m[7], m[2] * m[7] - m[1] * m[8], m[1] * m[5] - m[2] * m[4],
                   m[5] * m[6] - m[3] * m[8], m[0] * m[8] - m[2] * m[6], m[2] * m[3] - m[0] * m[5],
                   m[3] * m[7] - m[4] * m[6], m[1] * m[6] - m[0] * m[7], m[0] * m[4] - m[1] * m[3]};
  for (int i = 0; i < 9; i++) r.m[i] = adj[i] * d;
  return r;
}

}  // namespace orc
