// Per-ray traversal of the compressed 8-wide BVH and the ray/triangle tests.
//
// Replaces, per ray, the reference's recursive walk
//   JoinedCollider.FirstRayCollision   model3d/collisions.go:275-290
//   rayCollisionWithBounds             model3d/bvh.go:322-351
//   Triangle.rayCollision              model3d/primitives.go:207-249
// Differences that are allowed by the parity contract (ties / grazing only):
//   * children are visited near-to-far with tmax pruning (the reference visits every
//     child whose box the ray enters, unordered);
//   * the float32 triangle test is an edge-function (scalar triple product) test whose
//     shared-edge values are exact negations of each other, i.e. watertight;
//   * the winning hit is re-evaluated in float64 with the reference's own
//     Moeller-Trumbore arithmetic (refine_hit_f64), so Scale / Barycentric / Normal
//     match the float64 oracle to float rounding.
//
// Everything here is __host__ __device__ so that tests can run the identical code on
// the CPU against the same flattened BVH (tests only; the product launches kernels).
#pragma once
#include <stdint.h>

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

// float32 edge-function test.  Accepts t in [tmin, tmax], bary inclusive, no culling
// (primitives.go:183,232,238 accept scale >= 0 and inclusive barycentrics).
M3D_HD bool intersect_tri_f32(const float4 *__restrict__ tri, float ox, float oy, float oz, F3 d,
                              float inv_dd, float tmin, float tmax, float &t_out, float &b1,
                              float &b2) {
  float4 q0 = tri[0], q1 = tri[1], q2 = tri[2];
  F3 A = mk3(q0.x - ox, q0.y - oy, q0.z - oz);
  F3 B = mk3(q1.x - ox, q1.y - oy, q1.z - oz);
  F3 C = mk3(q2.x - ox, q2.y - oy, q2.z - oz);
  float U = dot_sym(d, cross_sym(B, C));  // weight of v0
  float V = dot_sym(d, cross_sym(C, A));  // weight of v1
  float W = dot_sym(d, cross_sym(A, B));  // weight of v2
  if ((U < 0.f || V < 0.f || W < 0.f) && (U > 0.f || V > 0.f || W > 0.f)) return false;
  float det = U + V + W;
  if (det == 0.f) return false;
  // hit point relative to the origin is (U*A + V*B + W*C)/det; project on d
  float T = U * dot_sym(A, d) + V * dot_sym(B, d) + W * dot_sym(C, d);
  float rdet = 1.0f / det;
  float t = T * rdet * inv_dd;
  if (!(t >= tmin && t <= tmax)) return false;
  t_out = t;
  b1 = V * rdet;
  b2 = W * rdet;
  return true;
}

// Traverse.  nodes: 5 x uint4 per node.  tris: 3 x float4 per triangle.
// skip_tri: triangle index to ignore (self-intersection guard for secondary rays), -1 none.
// ANY_HIT: return at the first accepted hit (shadow / visibility rays).
template <bool COUNT, bool ANY_HIT>
M3D_HD void trace_bvh(const uint4 *__restrict__ nodes, const float4 *__restrict__ tris,
                      const RayF &ray, int32_t skip_tri, HitF &hit, TraceCounters *cnt) {
  const float ooeps = 5.421010862e-20f;  // 2^-64: avoid 1/0 (bvh.go:328-333 handles rate==0 exactly)
  float dx = fabsf(ray.dx) > ooeps ? ray.dx : copysignf(ooeps, ray.dx);
  float dy = fabsf(ray.dy) > ooeps ? ray.dy : copysignf(ooeps, ray.dy);
  float dz = fabsf(ray.dz) > ooeps ? ray.dz : copysignf(ooeps, ray.dz);
  const float idx = 1.0f / dx, idy = 1.0f / dy, idz = 1.0f / dz;
  const F3 d = mk3(ray.dx, ray.dy, ray.dz);
  const float inv_dd = 1.0f / (ray.dx * ray.dx + ray.dy * ray.dy + ray.dz * ray.dz);
  const uint32_t octinv = ((ray.dx < 0.f ? 0u : 4u) | (ray.dy < 0.f ? 0u : 2u) | (ray.dz < 0.f ? 0u : 1u));
  const uint32_t octinv4 = octinv * 0x01010101u;
  const float tmin = ray.tmin;
  float tmax = ray.tmax;

  hit.tri = -1;
  hit.t = tmax;
  hit.b1 = hit.b2 = 0.f;

  uint2 stack[M3D_STACK_SIZE];
  int sp = 0;

  // node group: x = child base index, y = (hit bits 24..31) | imask (bits 0..7)
  // tri group:  x = triangle base index, y = hit bits 0..23
  uint2 ngroup, tgroup;
  ngroup.x = 0;
  ngroup.y = 0x80000000u;  // root: pretend slot (7 ^ octinv) of a virtual parent, see below
  tgroup.x = 0;
  tgroup.y = 0;
  bool root_pending = true;

  for (;;) {
    if (ngroup.y & 0xff000000u) {
      uint32_t node_index;
      if (root_pending) {
        root_pending = false;
        node_index = 0;
        ngroup.y = 0;
      } else {
        const uint32_t hits_imask = ngroup.y;
        const int bit = bfind32(hits_imask);
        ngroup.y &= ~(1u << bit);
        if (ngroup.y & 0xff000000u) {
          if (sp < M3D_STACK_SIZE) stack[sp++] = ngroup;
        }
        const uint32_t slot = (uint32_t)(bit - 24) ^ (octinv & 7u);
        const uint32_t rel = (uint32_t)popc32(hits_imask & ~(0xffffffffu << slot) & 0xffu);
        node_index = ngroup.x + rel;
      }
      if (COUNT) cnt->nodes++;

      const uint4 *np = nodes + (size_t)node_index * 5;
      const uint4 n0 = np[0], n1 = np[1], n2 = np[2], n3 = np[3], n4 = np[4];
      const uint32_t e = n0.w;
      const float sx = f_from_bits((e & 0xffu) << 23) * idx;
      const float sy = f_from_bits(((e >> 8) & 0xffu) << 23) * idy;
      const float sz = f_from_bits(((e >> 16) & 0xffu) << 23) * idz;
      const float bx = (f_from_bits(n0.x) - ray.ox) * idx;
      const float by = (f_from_bits(n0.y) - ray.oy) * idy;
      const float bz = (f_from_bits(n0.z) - ray.oz) * idz;
      uint32_t hitmask = 0;
#pragma unroll
      for (int half = 0; half < 2; half++) {
        const uint32_t meta4 = half == 0 ? n1.z : n1.w;
        const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
        const uint32_t bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1f1f1f1fu;
        const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
        const uint32_t qlox = half == 0 ? n2.x : n2.y, qloy = half == 0 ? n2.z : n2.w;
        const uint32_t qloz = half == 0 ? n3.x : n3.y, qhix = half == 0 ? n3.z : n3.w;
        const uint32_t qhiy = half == 0 ? n4.x : n4.y, qhiz = half == 0 ? n4.z : n4.w;
        const uint32_t xn = ray.dx < 0.f ? qhix : qlox, xf = ray.dx < 0.f ? qlox : qhix;
        const uint32_t yn = ray.dy < 0.f ? qhiy : qloy, yf = ray.dy < 0.f ? qloy : qhiy;
        const uint32_t zn = ray.dz < 0.f ? qhiz : qloz, zf = ray.dz < 0.f ? qloz : qhiz;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int sh = 8 * j;
          const float t0x = (float)((xn >> sh) & 0xffu) * sx + bx;
          const float t0y = (float)((yn >> sh) & 0xffu) * sy + by;
          const float t0z = (float)((zn >> sh) & 0xffu) * sz + bz;
          const float t1x = (float)((xf >> sh) & 0xffu) * sx + bx;
          const float t1y = (float)((yf >> sh) & 0xffu) * sy + by;
          const float t1z = (float)((zf >> sh) & 0xffu) * sz + bz;
          const float cmin = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tmin));
          const float cmax = fminf(fminf(t1x, t1y), fminf(t1z, tmax));
          // widen by a few ulp: the quantised slab arithmetic is not exactly conservative
          if (cmin <= cmax * 1.0000005f + 1e-30f) {
            const uint32_t cb = (child_bits4 >> sh) & 0xffu;
            const uint32_t bi = (bit_index4 >> sh) & 0xffu;
            hitmask |= cb << bi;
          }
        }
      }
      ngroup.x = n1.x;
      ngroup.y = (hitmask & 0xff000000u) | (e >> 24);
      tgroup.x = n1.y;
      tgroup.y = hitmask & 0x00ffffffu;
    }
    // (ngroup always carries pending internal hits here: it is either the root, a
    // freshly decoded node, or a popped stack entry)

    while (tgroup.y) {
      const int bit = bfind32(tgroup.y);
      tgroup.y &= ~(1u << bit);
      const int32_t ti = (int32_t)(tgroup.x + (uint32_t)bit);
      if (ti == skip_tri) continue;
      if (COUNT) cnt->tris++;
      float t, b1, b2;
      if (intersect_tri_f32(tris + (size_t)ti * 3, ray.ox, ray.oy, ray.oz, d, inv_dd, tmin, tmax, t,
                            b1, b2)) {
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
  }
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
