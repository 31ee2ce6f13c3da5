// Per-ray traversal of the compressed 8-wide BVH and the ray/triangle tests.
//
// Replaces, per ray, the reference's recursive walk
//   JoinedCollider.FirstRayCollision   model3d/collisions.go:275-290
//   rayCollisionWithBounds             model3d/bvh.go:322-351
//   Triangle.rayCollision              model3d/primitives.go:207-249
// Differences that are allowed by the parity contract (ties / grazing only):
//   * children are visited near-to-far with tmax pruning (the reference visits every
//     child whose box the ray enters, unordered);
//   * the float32 triangle test is the reference's Moeller-Trumbore with an explicit
//     rounding-error band; rays inside the band of an edge are decided in float64 with
//     the reference's own arithmetic, so no crack can open that the reference does not
//     have, and triangle ids match except for exact ties;
//   * the winning hit is re-evaluated in float64 with the reference's own
//     Moeller-Trumbore arithmetic (refine_hit_f64), so Scale / Barycentric / Normal
//     match the float64 oracle to float rounding.
//
// Everything here is __host__ __device__ so that tests can run the identical code on
// the CPU against the same flattened BVH (tests only; the product launches kernels).
#pragma once
#include <stdint.h>

#include "node_layout.h"

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define M3D_HD __host__ __device__ __forceinline__
#else
#include <vector_types.h>
#include <cmath>
#include <cstring>
#define M3D_HD inline
#endif

namespace m3d {

// ---- small portability shims ---------------------------------------------------------
M3D_HD float f_from_bits(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}
M3D_HD uint32_t bits_from_f(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
#endif
}
M3D_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}
// index of the highest set bit (x != 0)
M3D_HD int bfind32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return 31 - __clz((int)x);
#else
  return 31 - __builtin_clz(x);
#endif
}
// each byte -> 0xff if its msb is set else 0x00
M3D_HD uint32_t sign_extend_s8x4(uint32_t x) {
#if defined(__CUDA_ARCH__)
  uint32_t r;
  asm("prmt.b32 %0, %1, 0x0, 0x0000ba98;" : "=r"(r) : "r"(x));
  return r;
#else
  uint32_t r = 0;
  for (int i = 0; i < 4; i++)
    if (x & (0x80u << (8 * i))) r |= 0xffu << (8 * i);
  return r;
#endif
}
// non-contracted arithmetic: the watertightness argument needs a*b - c*d to be the
// exact negation of c*d - a*b, which an FMA contraction would break.
M3D_HD float mul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  volatile float r = a * b;
  return r;
#endif
}
M3D_HD float sub_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  volatile float r = a - b;
  return r;
#endif
}
M3D_HD float add_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b;
  return r;
#endif
}

// 1-ulp reciprocal (one MUFU.RCP on the device instead of the IEEE division sequence);
// only used for slab-test slopes and 1/|d|^2, both of which are consumed with slack.
M3D_HD float rcp_fast(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / x;
#endif
}

struct F3 {
  float x, y, z;
};
M3D_HD F3 mk3(float x, float y, float z) {
  F3 r;
  r.x = x;
  r.y = y;
  r.z = z;
  return r;
}

// exact-negation-symmetric cross and dot
M3D_HD F3 cross_sym(F3 a, F3 b) {
  return mk3(sub_rn(mul_rn(a.y, b.z), mul_rn(a.z, b.y)), sub_rn(mul_rn(a.z, b.x), mul_rn(a.x, b.z)),
             sub_rn(mul_rn(a.x, b.y), mul_rn(a.y, b.x)));
}
M3D_HD float dot_sym(F3 a, F3 b) {
  return add_rn(add_rn(mul_rn(a.x, b.x), mul_rn(a.y, b.y)), mul_rn(a.z, b.z));
}

// ---- ray ----------------------------------------------------------------------------------
struct RayF {
  float ox, oy, oz, tmin;
  float dx, dy, dz, tmax;
};

struct HitF {
  float t;       // in units of |d|
  float b1, b2;  // barycentric weights of v1, v2
  int32_t tri;   // index into the leaf-ordered triangle array, -1 = miss
};

struct TraceCounters {
  uint32_t nodes, tris;
};

#define M3D_STACK_SIZE 32
// deepest wide BVH the traversal stacks cover (one pending node group per level); builds beyond it
// are refused at m3d_mesh_create / m3d_scene_build instead of dropping pushes silently
#define M3D_MAX_BVH_DEPTH (M3D_STACK_SIZE - 1)

M3D_HD float max3f(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
#else
  return fmaxf(fmaxf(a, b), c);
#endif
}
M3D_HD float min3f(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
#else
  return fminf(fminf(a, b), c);
#endif
}
// byte j of q placed in mantissa bits 8..15 of 1.0f: value == 1 + q_j * 2^-15 exactly.
// One PRMT on the device instead of a byte extract + integer->float conversion.
// `one` is the bit pattern of 1.0f.  The traversal kernel passes it as a kernel parameter so that
// ptxas cannot fold it: PRMT takes only one non-register operand, and with both the selector and
// 0x3f800000 known it keeps re-materialising the four selectors in registers (one IMAD.U32 per
// two PRMTs in the node test); with `one` opaque the selector is the immediate.
M3D_HD float unit_plus_byte(uint32_t q, int j, uint32_t one) {
#if defined(__CUDA_ARCH__)
  uint32_t r;
  switch (j) {
    case 0: asm("prmt.b32 %0, %1, %2, 0x7604;" : "=r"(r) : "r"(q), "r"(one)); break;
    case 1: asm("prmt.b32 %0, %1, %2, 0x7614;" : "=r"(r) : "r"(q), "r"(one)); break;
    case 2: asm("prmt.b32 %0, %1, %2, 0x7624;" : "=r"(r) : "r"(q), "r"(one)); break;
    default: asm("prmt.b32 %0, %1, %2, 0x7634;" : "=r"(r) : "r"(q), "r"(one)); break;
  }
  return __uint_as_float(r);
#else
  return f_from_bits(one | (((q >> (8 * j)) & 0xffu) << 8));
#endif
}

#ifndef M3D_FFMA2
#define M3D_FFMA2 0
#endif
#if defined(__CUDACC__)
// (a0, a1) * s + (b0, b1) as one packed float32 FMA (sm_100: fma.rn.f32x2 -> FFMA2)
__device__ __forceinline__ void fma2(float a0, float a1, float s, float b0, float b1, float &r0, float &r1) {
  unsigned long long a, ss, b, r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(ss) : "f"(s));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(ss), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(r));
}
#endif

// M3D_L2_HINT (tuning builds): bit 0 = node loads, bit 1 = triangle loads carry an L2 evict_last
// cache policy (createpolicy folds into the load's descriptor: no register cost), so that the ray /
// hit streams passing through L2 evict each other instead of the hierarchy.
#ifndef M3D_L2_HINT
#define M3D_L2_HINT 0
#endif
#if defined(__CUDACC__)
__device__ __forceinline__ uint4 ldg_keep_u4(const uint4 *p) {
  uint64_t pol;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  uint4 v;
  asm("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
      : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
      : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float4 ldg_keep_f4(const float4 *p) {
  const uint4 v = ldg_keep_u4(reinterpret_cast<const uint4 *>(p));
  return make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
}
#endif

// The reference's Moeller-Trumbore in float64 (primitives.go:207-249) on the float32
// inputs widened exactly: the arbiter for rays that pass so close to a triangle edge (or
// are so nearly parallel to it) that the float32 test cannot decide.  Kept out of line: it
// runs for well under 1 % of the triangle tests and must not cost registers on the hot path.
#if defined(__CUDACC__)
static __host__ __device__ __noinline__
#else
static inline
#endif
bool tri_decide_f64(const float4 *__restrict__ tri, float oxf, float oyf, float ozf, float dxf,
                    float dyf, float dzf, float tmin, float tmax, float &t_out, float &b1, float &b2) {
  const float4 q0 = tri[0], q1 = tri[1], q2 = tri[2];
  const double dx = dxf, dy = dyf, dz = dzf;
  const double v1x = (double)q1.x - q0.x, v1y = (double)q1.y - q0.y, v1z = (double)q1.z - q0.z;
  const double v2x = (double)q2.x - q0.x, v2y = (double)q2.y - q0.y, v2z = (double)q2.z - q0.z;
  {
    // primitives.go:208: |normalize(v1 x v2) . normalize(d)| < 1e-8 -> parallel, no hit
    const double nx = v1y * v2z - v1z * v2y, ny = v1z * v2x - v1x * v2z, nz = v1x * v2y - v1y * v2x;
    const double nn = 1.0 / sqrt(nx * nx + ny * ny + nz * nz), dn = 1.0 / sqrt(dx * dx + dy * dy + dz * dz);
    const double c = (nx * nn) * (dx * dn) + (ny * nn) * (dy * dn) + (nz * nn) * (dz * dn);
    if (!(c >= 1e-8 || c <= -1e-8)) return false;
  }
  const double c1x = dy * v2z - dz * v2y, c1y = dz * v2x - dx * v2z, c1z = dx * v2y - dy * v2x;
  const double det = c1x * v1x + c1y * v1y + c1z * v1z;
  if (det == 0.0) return false;
  const double inv = 1.0 / det;
  const double px = (double)oxf - q0.x, py = (double)oyf - q0.y, pz = (double)ozf - q0.z;
  const double bary1 = inv * (px * c1x + py * c1y + pz * c1z);
  if (bary1 < 0 || bary1 > 1) return false;
  const double c2x = py * v1z - pz * v1y, c2y = pz * v1x - px * v1z, c2z = px * v1y - py * v1x;
  const double bary2 = inv * (dx * c2x + dy * c2y + dz * c2z);
  if (bary2 < 0 || bary1 + bary2 > 1) return false;
  const double t = inv * (v2x * c2x + v2y * c2y + v2z * c2z);
  if (!(t >= (double)tmin && t <= (double)tmax)) return false;
  t_out = (float)t;
  b1 = (float)bary1;
  b2 = (float)bary2;
  return true;
}

// Per-ray constants of the traversal.
struct RayPre {
  float ox, oy, oz;
  F3 d;
  float idx, idy, idz;  // reciprocal direction (zero components replaced by +-2^-64)
  float tmin;
  float err;            // float32 error scale of the triangle test: 3e-6 * |d| * Dmax
  uint32_t octinv4;     // (7 - octant) replicated in four bytes; bit 2/1/0 clear <=> dx/dy/dz < 0
};

// (7 - octant) of a direction replicated in four bytes
M3D_HD uint32_t ray_octinv4(float dx, float dy, float dz) {
  // by SIGN BIT, like the copysignf() of precompute_ray: a -0.0 component gets a negative
  // reciprocal there, so it must also count as negative here (near / far planes swap with it)
  const uint32_t neg = ((bits_from_f(dx) >> 31) << 2) | ((bits_from_f(dy) >> 31) << 1) | (bits_from_f(dz) >> 31);
  return (neg ^ 7u) * 0x01010101u;
}

// scene_min/max: bounds of all triangle vertices (for the error bound: no vertex is
// farther than Dmax from the origin in the infinity norm).
M3D_HD RayPre precompute_ray(const RayF &ray, const float *scene_min, const float *scene_max) {
  RayPre rp;
  const float ooeps = 5.421010862e-20f;  // 2^-64 (bvh.go:328-333 handles rate == 0 exactly)
  const float dx = fabsf(ray.dx) > ooeps ? ray.dx : copysignf(ooeps, ray.dx);
  const float dy = fabsf(ray.dy) > ooeps ? ray.dy : copysignf(ooeps, ray.dy);
  const float dz = fabsf(ray.dz) > ooeps ? ray.dz : copysignf(ooeps, ray.dz);
  rp.ox = ray.ox;
  rp.oy = ray.oy;
  rp.oz = ray.oz;
  rp.d = mk3(ray.dx, ray.dy, ray.dz);
  rp.idx = rcp_fast(dx);
  rp.idy = rcp_fast(dy);
  rp.idz = rcp_fast(dz);
  rp.tmin = ray.tmin;
  const float dmax = max3f(fmaxf(fabsf(ray.ox - scene_min[0]), fabsf(ray.ox - scene_max[0])),
                           fmaxf(fabsf(ray.oy - scene_min[1]), fabsf(ray.oy - scene_max[1])),
                           fmaxf(fabsf(ray.oz - scene_min[2]), fabsf(ray.oz - scene_max[2])));
  rp.err = 3e-6f * dmax * sqrtf(ray.dx * ray.dx + ray.dy * ray.dy + ray.dz * ray.dz);
  rp.octinv4 = ray_octinv4(ray.dx, ray.dy, ray.dz);
  return rp;
}

// Whether a ray certainly misses the bounds [bmin, bmax] of all vertices (the optional bounds cull in
// front of large mesh batches, cull_rays_kernel).  Conservative: the box is widened by 4e-6 of the origin's
// largest distance to it (> 30 x the float32 error of the slab arithmetic below), the accept rule is the
// traversal's (bvh.go:322-351: near <= far within [tmin, tmax]; zero direction components as +-2^-64 like
// precompute_ray), and a comparison with a NaN keeps the ray.
M3D_HD bool ray_misses_bounds(float ox, float oy, float oz, float tmin, float dx, float dy, float dz, float tmax,
                              const float *bmin, const float *bmax) {
  const float ooeps = 5.421010862e-20f;
  const float ex = fabsf(dx) > ooeps ? dx : copysignf(ooeps, dx);
  const float ey = fabsf(dy) > ooeps ? dy : copysignf(ooeps, dy);
  const float ez = fabsf(dz) > ooeps ? dz : copysignf(ooeps, dz);
  const float ix = 1.0f / ex, iy = 1.0f / ey, iz = 1.0f / ez;
  const float dmax = max3f(fmaxf(fabsf(ox - bmin[0]), fabsf(ox - bmax[0])), fmaxf(fabsf(oy - bmin[1]), fabsf(oy - bmax[1])),
                           fmaxf(fabsf(oz - bmin[2]), fabsf(oz - bmax[2])));
  const float m = 4e-6f * dmax + 1e-30f;
  const float ax = (bmin[0] - m - ox) * ix, bx = (bmax[0] + m - ox) * ix;
  const float ay = (bmin[1] - m - oy) * iy, by = (bmax[1] + m - oy) * iy;
  const float az = (bmin[2] - m - oz) * iz, bz = (bmax[2] + m - oz) * iz;
  const float t_near = fmaxf(max3f(fminf(ax, bx), fminf(ay, by), fminf(az, bz)), tmin);
  const float t_far = fminf(min3f(fmaxf(ax, bx), fmaxf(ay, by), fmaxf(az, bz)), tmax);
  return t_near > t_far;
}

// Ray/triangle test of the traversal: the reference's Moeller-Trumbore (primitives.go:
// 207-249) in float32.  Its barycentrics carry an absolute error of a few ulp of
// |o - v0| * |d| * |edge| / det, far larger than 1 ulp for small, distant triangles, so a
// ray whose barycentric (or det) lies within that bound of the accept/reject boundary is
// re-decided by tri_decide_f64: triangle ids then agree with the float64 reference except
// for exact ties.  tri[2].w carries the triangle's longest edge (infinity norm) for the
// bound.  Accepts t in [tmin, tmax], barycentrics inclusive, no back-face culling
// (primitives.go:183,232,238).
M3D_HD bool intersect_tri_loaded(const float4 *__restrict__ tri, const float4 q0, const float4 q1, const float4 q2,
                                 const RayPre &rp, float tmax, float &t_out, float &b1, float &b2);

M3D_HD bool intersect_tri(const float4 *__restrict__ tri, const RayPre &rp, float tmax, float &t_out,
                          float &b1, float &b2) {
#if defined(__CUDA_ARCH__) && (M3D_L2_HINT & 2)
  const float4 q0 = ldg_keep_f4(tri), q1 = ldg_keep_f4(tri + 1), q2 = ldg_keep_f4(tri + 2);
#elif defined(__CUDA_ARCH__)
  const float4 q0 = __ldg(tri), q1 = __ldg(tri + 1), q2 = __ldg(tri + 2);
#else
  const float4 q0 = tri[0], q1 = tri[1], q2 = tri[2];
#endif
  return intersect_tri_loaded(tri, q0, q1, q2, rp, tmax, t_out, b1, b2);
}

// Same test on a triangle record that is already in registers (the traversal kernel issues the
// three loads one node visit ahead of the test to hide their latency).
M3D_HD bool intersect_tri_loaded(const float4 *__restrict__ tri, const float4 q0, const float4 q1, const float4 q2,
                                 const RayPre &rp, float tmax, float &t_out, float &b1, float &b2) {
  const float e1x = q1.x - q0.x, e1y = q1.y - q0.y, e1z = q1.z - q0.z;
  const float e2x = q2.x - q0.x, e2y = q2.y - q0.y, e2z = q2.z - q0.z;
  const float c1x = rp.d.y * e2z - rp.d.z * e2y, c1y = rp.d.z * e2x - rp.d.x * e2z,
              c1z = rp.d.x * e2y - rp.d.y * e2x;  // d x e2
  const float det = c1x * e1x + c1y * e1y + c1z * e1z;
  const float px = rp.ox - q0.x, py = rp.oy - q0.y, pz = rp.oz - q0.z;
  const float n1 = px * c1x + py * c1y + pz * c1z;  // bary1 * det
  const float c2x = py * e1z - pz * e1y, c2y = pz * e1x - px * e1z, c2z = px * e1y - py * e1x;  // o x e1
  const float n2 = rp.d.x * c2x + rp.d.y * c2y + rp.d.z * c2z;  // bary2 * det
  const float n0 = det - (n1 + n2);                             // bary0 * det
  const float band = rp.err * q2.w;  // error bound of n0, n1, n2 and det
  const float adet = fabsf(det);
  // clearly outside (beyond the band) with respect to some edge: no hit
  const float sgn = det < 0.f ? -1.f : 1.f;
  const float m = min3f(n0 * sgn, n1 * sgn, n2 * sgn);
  if (m < -band) return false;
  if (m <= band || adet <= band)
    return tri_decide_f64(tri, rp.ox, rp.oy, rp.oz, rp.d.x, rp.d.y, rp.d.z, rp.tmin, tmax, t_out, b1, b2);
  const float inv = rcp_fast(det);
  const float t = (e2x * c2x + e2y * c2y + e2z * c2z) * inv;
  if (!(t >= rp.tmin && t <= tmax)) return false;
  t_out = t;
  b1 = n1 * inv;
  b2 = n2 * inv;
  return true;
}

// Fetch one wide node (five 16-byte loads) and slab-test its eight quantised child boxes.
// Outputs the node group (x = child base, y = internal hit bits 24..31 | imask) and the
// triangle group (x = triangle base, y = leaf hit bits 0..23).
//
// Quantised plane q on axis a sits at origin_a + q * 2^e_a.  With v = 1 + q * 2^-15
// (unit_plus_byte) the ray parameter of the plane is  v * S + b  where
// S = 2^(e_a+15) / d_a and b = (origin_a - o_a) / d_a - S: one PRMT and one FMA per plane.
// b absorbs a rounding error of at most ulp(S)/2, so near planes are moved back and far
// planes forward by |S| * 2^-22 (0.8 % of a grid step) to stay conservative.
// lut (optional, 256 entries 12 bytes apart in shared memory): lut[m] = ((m >> 5) & 7) << (m & 31), the hit bits of a
// child with meta byte m.  The shift-and-mask sequence it replaces is ~3 ALU-pipe instructions per
// child, and the ALU pipe (PRMT / FMNMX / LOP3 / SHF, half rate) is the busiest unit of the kernel.
template <bool USE_LUT = false>
M3D_HD void intersect_node(const uint4 *__restrict__ nodes, uint32_t node_index, const RayPre &rp,
                           float tmax, uint2 &ngroup, uint2 &tgroup, uint32_t one = 0x3f800000u,
                           uint32_t lut_saddr = 0u /* shared-window address of the table */) {
  const uint4 *np = nodes + (size_t)node_index * M3D_NODE_QUADS;
#if defined(__CUDA_ARCH__) && M3D_NODE_QUADS == 6
  // 32-byte aligned nodes: two 256-bit loads + one 128-bit load, one sector each
  uint4 n0, n1, n2, n3;
  asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(n0.x), "=r"(n0.y), "=r"(n0.z), "=r"(n0.w), "=r"(n1.x), "=r"(n1.y), "=r"(n1.z), "=r"(n1.w)
      : "l"(np));
  asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(n2.x), "=r"(n2.y), "=r"(n2.z), "=r"(n2.w), "=r"(n3.x), "=r"(n3.y), "=r"(n3.z), "=r"(n3.w)
      : "l"(np + 2));
  const uint4 n4 = __ldg(np + 4);
#elif defined(__CUDA_ARCH__) && (M3D_L2_HINT & 1)
  const uint4 n0 = ldg_keep_u4(np), n1 = ldg_keep_u4(np + 1), n2 = ldg_keep_u4(np + 2), n3 = ldg_keep_u4(np + 3),
              n4 = ldg_keep_u4(np + 4);
#elif defined(__CUDA_ARCH__)
  const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
#else
  const uint4 n0 = np[0], n1 = np[1], n2 = np[2], n3 = np[3], n4 = np[4];
#endif
  const uint32_t e = n0.w;
  const float Sx = f_from_bits((e & 0xffu) << 23) * (32768.0f * rp.idx);
  const float Sy = f_from_bits(((e >> 8) & 0xffu) << 23) * (32768.0f * rp.idy);
  const float Sz = f_from_bits(((e >> 16) & 0xffu) << 23) * (32768.0f * rp.idz);
  const float bx = (f_from_bits(n0.x) - rp.ox) * rp.idx - Sx;
  const float by = (f_from_bits(n0.y) - rp.oy) * rp.idy - Sy;
  const float bz = (f_from_bits(n0.z) - rp.oz) * rp.idz - Sz;
  const float kSlack = 2.384185791e-7f;  // 2^-22
  const float gx = fabsf(Sx) * kSlack, gy = fabsf(Sy) * kSlack, gz = fabsf(Sz) * kSlack;
  const float bnx = bx - gx, bfx = bx + gx;
  const float bny = by - gy, bfy = by + gy;
  const float bnz = bz - gz, bfz = bz + gz;
  const float tmax_w = tmax * 1.0000005f + 1e-30f;  // few-ulp widening of the far bound
  uint32_t hitmask = 0;
#pragma unroll
  for (int half = 0; half < 2; half++) {
    const uint32_t meta4 = half == 0 ? n1.z : n1.w;
    const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
    const uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
    const uint32_t lut_index4 = meta4 ^ (rp.octinv4 & inner_mask4);
    const uint32_t bit_index4 = lut_index4 & 0x1f1f1f1fu;
    const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
    const uint32_t qlox = half == 0 ? n2.x : n2.y, qloy = half == 0 ? n2.z : n2.w;
    const uint32_t qloz = half == 0 ? n3.x : n3.y, qhix = half == 0 ? n3.z : n3.w;
    const uint32_t qhiy = half == 0 ? n4.x : n4.y, qhiz = half == 0 ? n4.z : n4.w;
    const uint32_t xn = (rp.octinv4 & 4u) ? qlox : qhix, xf = (rp.octinv4 & 4u) ? qhix : qlox;
    const uint32_t yn = (rp.octinv4 & 2u) ? qloy : qhiy, yf = (rp.octinv4 & 2u) ? qhiy : qloy;
    const uint32_t zn = (rp.octinv4 & 1u) ? qloz : qhiz, zf = (rp.octinv4 & 1u) ? qhiz : qloz;
#pragma unroll
    for (int j = 0; j < 4; j++) {
#if defined(__CUDA_ARCH__) && M3D_FFMA2
      // near and far plane of one axis in one packed FMA (FFMA2: two float32 FMAs, one issue slot;
      // the kernel is bound by issue slots, not by the FMA pipe)
      float t0x, t1x, t0y, t1y, t0z, t1z;
      fma2(unit_plus_byte(xn, j, one), unit_plus_byte(xf, j, one), Sx, bnx, bfx, t0x, t1x);
      fma2(unit_plus_byte(yn, j, one), unit_plus_byte(yf, j, one), Sy, bny, bfy, t0y, t1y);
      fma2(unit_plus_byte(zn, j, one), unit_plus_byte(zf, j, one), Sz, bnz, bfz, t0z, t1z);
#else
      const float t0x = fmaf(unit_plus_byte(xn, j, one), Sx, bnx);
      const float t0y = fmaf(unit_plus_byte(yn, j, one), Sy, bny);
      const float t0z = fmaf(unit_plus_byte(zn, j, one), Sz, bnz);
      const float t1x = fmaf(unit_plus_byte(xf, j, one), Sx, bfx);
      const float t1y = fmaf(unit_plus_byte(yf, j, one), Sy, bfy);
      const float t1z = fmaf(unit_plus_byte(zf, j, one), Sz, bfz);
#endif
      const float cmin = fmaxf(max3f(t0x, t0y, t0z), rp.tmin);
      const float cmax = fminf(min3f(t1x, t1y, t1z), tmax_w);
      if (USE_LUT) {
#if defined(__CUDA_ARCH__)
        // PRMT (byte j) + IMAD (address, FMA pipe) + LDS, issued for all lanes: no dependence on the
        // slab test, so the eight loads are in flight while the min / max chains run
        const uint32_t m = __byte_perm(lut_index4, 0u, 0x4440u + (uint32_t)j);
        uint32_t bits;
        uint32_t addr;
        // stride 12 bytes, not 4: a power-of-two scale becomes LEA (ALU pipe), this one IMAD (FMA pipe)
        asm("mad.lo.u32 %0, %1, 12, %2;" : "=r"(addr) : "r"(m), "r"(lut_saddr));
        asm("ld.shared.u32 %0, [%1];" : "=r"(bits) : "r"(addr));
        if (cmin <= cmax) hitmask |= bits;
#endif
      } else if (cmin <= cmax) {
        const uint32_t cb = (child_bits4 >> (8 * j)) & 0xffu;
        const uint32_t bi = (bit_index4 >> (8 * j)) & 0xffu;
        hitmask |= cb << bi;
      }
    }
  }
  ngroup.x = n1.x;
  ngroup.y = (hitmask & 0xff000000u) | (e >> 24);
  tgroup.x = n1.y;
  tgroup.y = hitmask & 0x00ffffffu;
}

// Pop the nearest pending child of a node group: returns its node index and clears its
// hit bit (the caller pushes the remainder if bits 24..31 are still non-zero).
M3D_HD uint32_t take_nearest_child(uint2 &ngroup, uint32_t octinv4) {
  const uint32_t hits_imask = ngroup.y;
  const int bit = bfind32(hits_imask);
  ngroup.y &= ~(1u << bit);
  const uint32_t slot = (uint32_t)(bit - 24) ^ (octinv4 & 7u);
  const uint32_t rel = (uint32_t)popc32(hits_imask & ~(0xffffffffu << slot) & 0xffu);
  return ngroup.x + rel;
}

// Scalar traversal of one ray (used by the CPU-side tests through tests/emul and by the
// shading kernels for single secondary rays).  nodes: 5 x uint4 per node.  tris: 3 x
// float4 per triangle.  skip_tri: triangle index to ignore (self-intersection guard for
// secondary rays), -1 none.  ANY_HIT: return at the first accepted hit.
template <bool COUNT, bool ANY_HIT>
M3D_HD void trace_bvh(const uint4 *__restrict__ nodes, const float4 *__restrict__ tris,
                      const float *scene_min, const float *scene_max, const RayF &ray,
                      int32_t skip_tri, HitF &hit, TraceCounters *cnt) {
  const RayPre rp = precompute_ray(ray, scene_min, scene_max);
  float tmax = ray.tmax;
  hit.tri = -1;
  hit.t = tmax;
  hit.b1 = hit.b2 = 0.f;

  uint2 stack[M3D_STACK_SIZE];
  int sp = 0;
  uint2 ngroup, tgroup;
  uint32_t node_index = 0;  // root
  for (;;) {
    if (COUNT) cnt->nodes++;
    intersect_node(nodes, node_index, rp, tmax, ngroup, tgroup);
    while (tgroup.y) {
      const int bit = bfind32(tgroup.y);
      tgroup.y &= ~(1u << bit);
      const int32_t ti = (int32_t)(tgroup.x + (uint32_t)bit);
      if (ti == skip_tri) continue;
      if (COUNT) cnt->tris++;
      float t, b1, b2;
      if (intersect_tri(tris + (size_t)ti * 3, rp, tmax, t, b1, b2)) {
        tmax = t;
        hit.t = t;
        hit.b1 = b1;
        hit.b2 = b2;
        hit.tri = ti;
        if (ANY_HIT) return;
      }
    }
    if ((ngroup.y & 0xff000000u) == 0) {
      if (sp == 0) break;
      ngroup = stack[--sp];
    }
    node_index = take_nearest_child(ngroup, rp.octinv4);
    if ((ngroup.y & 0xff000000u) && sp < M3D_STACK_SIZE) stack[sp++] = ngroup;
  }
}

// Generic first-hit walk over a wide BVH whose leaves are not triangles (the object-level
// hierarchy of a scene: analytic shapes and mesh instances).  leaf(index, tmax) tests the primitive
// in leaf slot `index` and returns the (possibly shortened) far bound; children are visited near to
// far and pruned with it like in trace_bvh.
template <class LeafFn>
M3D_HD void walk_bvh_leaves(const uint4 *__restrict__ nodes, const float *scene_min, const float *scene_max,
                            const RayF &ray, LeafFn &&leaf) {
  const RayPre rp = precompute_ray(ray, scene_min, scene_max);
  float tmax = ray.tmax;
  uint2 stack[M3D_STACK_SIZE];
  int sp = 0;
  uint2 ngroup, tgroup;
  uint32_t node_index = 0;  // root
  for (;;) {
    intersect_node(nodes, node_index, rp, tmax, ngroup, tgroup);
    while (tgroup.y) {
      const int bit = bfind32(tgroup.y);
      tgroup.y &= ~(1u << bit);
      tmax = leaf((int32_t)(tgroup.x + (uint32_t)bit), tmax);
    }
    if ((ngroup.y & 0xff000000u) == 0) {
      if (sp == 0) break;
      ngroup = stack[--sp];
    }
    node_index = take_nearest_child(ngroup, rp.octinv4);
    if ((ngroup.y & 0xff000000u) && sp < M3D_STACK_SIZE) stack[sp++] = ngroup;
  }
}

// All-hits traversal: the number of triangles the ray's forward half-line (t >= tmin) crosses ==
// Collider.RayCollisions(r, nil) (model3d/collisions.go:263-273, primitives.go:189-196).  No
// tmax pruning and no ordering: every child whose box the ray enters is visited, like the
// reference.  Each triangle is decided by the same float32 test + float64 arbitration as the
// first-hit query, so counts equal the float64 reference's except for exact ties.
M3D_HD int count_bvh_hits(const uint4 *__restrict__ nodes, const float4 *__restrict__ tris,
                          const float *scene_min, const float *scene_max, const RayF &ray) {
  const RayPre rp = precompute_ray(ray, scene_min, scene_max);
  const float tmax = ray.tmax;
  int count = 0;
  uint2 stack[M3D_STACK_SIZE];
  int sp = 0;
  uint2 ngroup, tgroup;
  uint32_t node_index = 0;
  for (;;) {
    intersect_node(nodes, node_index, rp, tmax, ngroup, tgroup);
    while (tgroup.y) {
      const int bit = bfind32(tgroup.y);
      tgroup.y &= ~(1u << bit);
      const int32_t ti = (int32_t)(tgroup.x + (uint32_t)bit);
      float t, b1, b2;
      if (intersect_tri(tris + (size_t)ti * 3, rp, tmax, t, b1, b2)) count++;
    }
    if ((ngroup.y & 0xff000000u) == 0) {
      if (sp == 0) break;
      ngroup = stack[--sp];
    }
    node_index = take_nearest_child(ngroup, rp.octinv4);
    if ((ngroup.y & 0xff000000u) && sp < M3D_STACK_SIZE) stack[sp++] = ngroup;
  }
  return count;
}

// All-hits traversal that also delivers the hits (Collider.RayCollisions(r, f) with f != nil,
// model3d/collisions.go:263-273): same walk and same per-triangle decision as count_bvh_hits,
// so the two agree hit for hit; the first `cap` hits are written as (float32 t, leaf-order
// triangle index).  Returns the number of hits found (may exceed cap).
M3D_HD int collect_bvh_hits(const uint4 *__restrict__ nodes, const float4 *__restrict__ tris,
                            const float *scene_min, const float *scene_max, const RayF &ray, int cap,
                            float *t_out, int32_t *tri_out) {
  const RayPre rp = precompute_ray(ray, scene_min, scene_max);
  const float tmax = ray.tmax;
  int count = 0;
  uint2 stack[M3D_STACK_SIZE];
  int sp = 0;
  uint2 ngroup, tgroup;
  uint32_t node_index = 0;
  for (;;) {
    intersect_node(nodes, node_index, rp, tmax, ngroup, tgroup);
    while (tgroup.y) {
      const int bit = bfind32(tgroup.y);
      tgroup.y &= ~(1u << bit);
      const int32_t ti = (int32_t)(tgroup.x + (uint32_t)bit);
      float t, b1, b2;
      if (intersect_tri(tris + (size_t)ti * 3, rp, tmax, t, b1, b2)) {
        if (count < cap) {
          t_out[count] = t;
          tri_out[count] = ti;
        }
        count++;
      }
    }
    if ((ngroup.y & 0xff000000u) == 0) {
      if (sp == 0) break;
      ngroup = stack[--sp];
    }
    node_index = take_nearest_child(ngroup, rp.octinv4);
    if ((ngroup.y & 0xff000000u) && sp < M3D_STACK_SIZE) stack[sp++] = ngroup;
  }
  return count;
}

// ---- float64 re-evaluation of the winning triangle ---------------------------------------
// The reference's Moeller-Trumbore (primitives.go:207-249) and flat normal
// (primitives.go:27-33) in double, on the float32 inputs widened exactly.
struct HitD {
  double t, b0, b1, b2;
  double nx, ny, nz;
  bool inside;  // float64 test agrees that the ray is inside the triangle and t >= 0
};

M3D_HD HitD refine_hit_f64(const float4 *__restrict__ tri, float oxf, float oyf, float ozf, float dxf,
                           float dyf, float dzf) {
  const float4 q0 = tri[0], q1 = tri[1], q2 = tri[2];
  const double ox = oxf, oy = oyf, oz = ozf, dx = dxf, dy = dyf, dz = dzf;
  const double v1x = (double)q1.x - q0.x, v1y = (double)q1.y - q0.y, v1z = (double)q1.z - q0.z;
  const double v2x = (double)q2.x - q0.x, v2y = (double)q2.y - q0.y, v2z = (double)q2.z - q0.z;
  // cross1 = d x v2
  const double c1x = dy * v2z - dz * v2y, c1y = dz * v2x - dx * v2z, c1z = dx * v2y - dy * v2x;
  const double det = c1x * v1x + c1y * v1y + c1z * v1z;
  const double inv = 1.0 / det;
  const double px = ox - q0.x, py = oy - q0.y, pz = oz - q0.z;
  const double bary1 = inv * (px * c1x + py * c1y + pz * c1z);
  // cross2 = o x v1
  const double c2x = py * v1z - pz * v1y, c2y = pz * v1x - px * v1z, c2z = px * v1y - py * v1x;
  const double bary2 = inv * (dx * c2x + dy * c2y + dz * c2z);
  HitD h;
  h.t = inv * (v2x * c2x + v2y * c2y + v2z * c2z);
  h.b1 = bary1;
  h.b2 = bary2;
  h.b0 = 1.0 - (bary1 + bary2);
  h.inside = !(bary1 < 0 || bary1 > 1 || bary2 < 0 || bary1 + bary2 > 1) && h.t >= 0 && det != 0;
  // normal = normalize(v1 x v2) as Scale(1/Norm) (coords.go:379-381)
  double nx = v1y * v2z - v1z * v2y, ny = v1z * v2x - v1x * v2z, nz = v1x * v2y - v1y * v2x;
  const double s = 1.0 / sqrt(nx * nx + ny * ny + nz * nz);
  h.nx = nx * s;
  h.ny = ny * s;
  h.nz = nz * s;
  return h;
}

}  // namespace m3d
