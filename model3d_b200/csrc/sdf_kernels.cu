// sm_100a kernels for nearest-triangle queries on the compressed wide BVH:
//   meshDistFunc.Dist / meshSDF.{SDF,PointSDF,NormalSDF,FaceSDF}   model3d/sdf.go:186-311
//   Triangle.Closest                                                model3d/primitives.go:153-175
//   JoinedCollider.SphereCollision / Triangle.SphereCollision       model3d/collisions.go:292-303,
//                                                                   model3d/primitives.go:253-279
//   ColliderSolid.Contains (sign of the SDF)                        model3d/solid.go:292-300
//
// One thread per query point.  The traversal visits children near-to-far by the distance from
// the point to the (outward-rounded, hence conservative) quantised child boxes and prunes with
// the best distance found so far (the reference prunes a binary tree the same way,
// sdf.go:289-310).  Triangles are screened with a float32 closest-point test (Ericson's
// region walk), one triangle per trip for all lanes that have one pending; the four triangles
// with the smallest screened distances are kept, and those within the float32 error band of the
// best are evaluated at the end -- by all lanes together -- in float64 with the reference's own
// arithmetic (matrix inverse for the interior case, NewSegment-ordered segment projections for
// the edges).  The result equals the reference's float64 minimum except (a) which face wins a
// tie on a shared edge / vertex and (b) when more than four distinct triangles lie within the
// float32 band (~1e-6 of the scene extent) of the minimum, where it may exceed it by that band.
#include <cfloat>

#include "kernels.h"
#include "trace_core.cuh"

namespace m3d {

namespace {

constexpr int kSdfBlock = 128;
constexpr int kSdfStack = 128;

struct D3 {
  double x, y, z;
};
__device__ __forceinline__ D3 d3(double x, double y, double z) {
  D3 r;
  r.x = x, r.y = y, r.z = z;
  return r;
}
__device__ __forceinline__ D3 dsub(D3 a, D3 b) { return d3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ D3 dadd(D3 a, D3 b) { return d3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ D3 dscale(D3 a, double s) { return d3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ double ddot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ D3 dcross(D3 a, D3 b) {
  return d3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ double dnorm(D3 a) { return sqrt(ddot(a, a)); }

// Segment.Closest on NewSegment(p1, p2) (primitives.go:547-554, 579-593)
__device__ __forceinline__ D3 segment_closest_f64(D3 p1, D3 p2, D3 c) {
  D3 s0 = p1, s1 = p2;
  if (!(p1.x < p2.x || (p1.x == p2.x && p1.y < p2.y) || (p1.x == p2.x && p1.y == p2.y && p1.z < p2.z))) {
    s0 = p2;
    s1 = p1;
  }
  const D3 v1 = dsub(s1, s0);
  const double nrm = dnorm(v1);
  const D3 v = dscale(v1, 1.0 / nrm);
  const double mag = ddot(v, dsub(c, s0));
  if (mag > nrm) return s1;
  if (mag < 0) return s0;
  return dadd(dscale(v, mag), s0);
}

// Triangle.Closest (primitives.go:153-175) on the float32 inputs widened exactly.  Out of line:
// it runs for the few candidates per query that survive the float32 screen.
__device__ __noinline__ double tri_closest_f64(const float4 *__restrict__ tri, float pxf, float pyf, float pzf,
                                               double *cp_out) {
  const float4 q0 = tri[0], q1 = tri[1], q2 = tri[2];
  const D3 c = d3(pxf, pyf, pzf);
  const D3 t0 = d3(q0.x, q0.y, q0.z), t1 = d3(q1.x, q1.y, q1.z), t2 = d3(q2.x, q2.y, q2.z);
  const D3 v1 = dsub(t1, t0), v2 = dsub(t2, t0);
  D3 n = dcross(v1, v2);
  n = dscale(n, 1.0 / dnorm(n));  // Normalize = Scale(1/Norm) (coords.go:379-381)
  // NewMatrix3Columns(v1, v2, n), row-major (matrix.go:16-22); InvertInPlace (matrix.go:57-90)
  const double m0 = v1.x, m1 = v2.x, m2 = n.x, m3 = v1.y, m4 = v2.y, m5 = n.y, m6 = v1.z, m7 = v2.z, m8 = n.z;
  const double det = m0 * (m4 * m8 - m5 * m7) - m1 * (m3 * m8 - m5 * m6) + m2 * (m3 * m7 - m4 * m6);
  const double id = 1.0 / det;
  const double i0 = (m4 * m8 - m5 * m7) * id, i1 = (m2 * m7 - m1 * m8) * id, i2 = (m1 * m5 - m2 * m4) * id;
  const double i3 = (m5 * m6 - m3 * m8) * id, i4 = (m0 * m8 - m2 * m6) * id, i5 = (m2 * m3 - m0 * m5) * id;
  const D3 r = dsub(c, t0);
  const double cx = i0 * r.x + i1 * r.y + i2 * r.z;
  const double cy = i3 * r.x + i4 * r.y + i5 * r.z;
  D3 best;
  if (cx >= 0 && cy >= 0 && cx + cy <= 1) {
    best = dadd(dadd(t0, dscale(v1, cx)), dscale(v2, cy));
  } else {
    double bd = INFINITY;
    best = d3(0, 0, 0);
    const D3 p[3] = {t0, t1, t2};
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const D3 c1 = segment_closest_f64(p[i], p[(i + 1) % 3], c);
      const double d = dnorm(dsub(c1, c));
      if (d < bd) {
        bd = d;
        best = c1;
      }
    }
  }
  cp_out[0] = best.x;
  cp_out[1] = best.y;
  cp_out[2] = best.z;
  return dnorm(dsub(best, c));
}

// float32 squared distance from p to the triangle (Ericson, "Real-Time Collision Detection"
// 5.1.5: Voronoi-region walk); screening only.
__device__ __forceinline__ float tri_dist2_f32(const float4 q0, const float4 q1, const float4 q2, float px,
                                               float py, float pz) {
  const float abx = q1.x - q0.x, aby = q1.y - q0.y, abz = q1.z - q0.z;
  const float acx = q2.x - q0.x, acy = q2.y - q0.y, acz = q2.z - q0.z;
  const float apx = px - q0.x, apy = py - q0.y, apz = pz - q0.z;
  const float d1 = abx * apx + aby * apy + abz * apz;
  const float d2 = acx * apx + acy * apy + acz * apz;
  float cx, cy, cz;  // closest - a
  if (d1 <= 0.f && d2 <= 0.f) {
    cx = cy = cz = 0.f;
  } else {
    const float bpx = px - q1.x, bpy = py - q1.y, bpz = pz - q1.z;
    const float d3_ = abx * bpx + aby * bpy + abz * bpz;
    const float d4 = acx * bpx + acy * bpy + acz * bpz;
    const float cpx = px - q2.x, cpy = py - q2.y, cpz = pz - q2.z;
    const float d5 = abx * cpx + aby * cpy + abz * cpz;
    const float d6 = acx * cpx + acy * cpy + acz * cpz;
    const float vc = d1 * d4 - d3_ * d2;
    const float vb = d5 * d2 - d1 * d6;
    const float va = d3_ * d6 - d5 * d4;
    if (d3_ >= 0.f && d4 <= d3_) {
      cx = abx, cy = aby, cz = abz;
    } else if (d6 >= 0.f && d5 <= d6) {
      cx = acx, cy = acy, cz = acz;
    } else if (vc <= 0.f && d1 >= 0.f && d3_ <= 0.f) {
      const float v = d1 / (d1 - d3_);
      cx = v * abx, cy = v * aby, cz = v * abz;
    } else if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) {
      const float w = d2 / (d2 - d6);
      cx = w * acx, cy = w * acy, cz = w * acz;
    } else if (va <= 0.f && (d4 - d3_) >= 0.f && (d5 - d6) >= 0.f) {
      const float w = (d4 - d3_) / ((d4 - d3_) + (d5 - d6));
      cx = abx + w * (acx - abx), cy = aby + w * (acy - aby), cz = abz + w * (acz - abz);
    } else {
      const float den = 1.f / (va + vb + vc);
      const float v = vb * den, w = vc * den;
      cx = abx * v + acx * w, cy = aby * v + acy * w, cz = abz * v + acz * w;
    }
  }
  const float ex = apx - cx, ey = apy - cy, ez = apz - cz;
  return ex * ex + ey * ey + ez * ez;
}

struct NearestResult {
  double dist;   // +inf: nothing within the initial bound
  double cp[3];
  int tri;       // leaf-order triangle index, -1 none
  unsigned nodes, screened, exact;  // work counters: nodes fetched, float32 screens, float64 evaluations
};

constexpr int kCand = 4;  // float32 near-ties kept for the float64 decision

// Nearest triangle within `bound` (exclusive, like Triangle.SphereCollision's `< r`; +inf for
// the SDF).  The traversal runs entirely in float32: it keeps the kCand triangles with the
// smallest screened distances (everything within the float32 error band of the best), and
// only those are evaluated in float64 at the end, by all lanes of the warp together.  ANY
// (SphereCollision): stop as soon as a triangle is closer than the bound by more than the band.
template <bool ANY>
__device__ __forceinline__ void nearest_triangle(const DeviceBVH &bvh, float px, float py, float pz, double bound,
                                                 NearestResult &res, uint2 *stack) {
  res.dist = bound;
  res.tri = -1;
  res.cp[0] = res.cp[1] = res.cp[2] = 0.0;
  res.nodes = res.screened = res.exact = 0u;
  if (bvh.num_tris <= 0) return;
  // absolute float32 error scale of a distance: a few ulp of the largest coordinate difference
  const float big = max3f(fmaxf(fabsf(px - bvh.bmin[0]), fabsf(px - bvh.bmax[0])),
                          fmaxf(fabsf(py - bvh.bmin[1]), fabsf(py - bvh.bmax[1])),
                          fmaxf(fabsf(pz - bvh.bmin[2]), fabsf(pz - bvh.bmax[2])));
  const float band = 2e-6f * big;
  const float fbound = isinf(bound) ? INFINITY : __double2float_ru(bound) * 1.000001f;
  // candidates, sorted by screened distance; cd_[0] is the float32 best
  float cd_[kCand];
  int ci_[kCand];
#pragma unroll
  for (int k = 0; k < kCand; k++) {
    cd_[k] = INFINITY;
    ci_[k] = -1;
  }
  float prune = fbound + band;  // nothing farther than this can win (or tie within the band)
  bool certain = false;         // ANY: a triangle is closer than the bound beyond doubt
  int sp = 0;
  stack[sp++] = make_uint2(0u, 0u);
  while (sp > 0 && !(ANY && certain)) {
    const uint2 e = stack[--sp];
    if (__uint_as_float(e.y) > prune) continue;
    res.nodes++;
    const uint4 *np = bvh.nodes + (size_t)e.x * M3D_NODE_QUADS;
    const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
    const float sx = __uint_as_float((n0.w & 0xffu) << 23);
    const float sy = __uint_as_float(((n0.w >> 8) & 0xffu) << 23);
    const float sz = __uint_as_float(((n0.w >> 16) & 0xffu) << 23);
    const float rx = px - __uint_as_float(n0.x), ry = py - __uint_as_float(n0.y), rz = pz - __uint_as_float(n0.z);
    const uint32_t imask = n0.w >> 24;
    float cdist[8];
    uint32_t inner = 0;  // slots of internal children that survive pruning
    uint32_t tmask = 0;  // leaf triangles to screen: bit = offset from tri_base
#pragma unroll
    for (int s = 0; s < 8; s++) {
      const uint32_t meta = ((s < 4 ? n1.z : n1.w) >> (8 * (s & 3))) & 0xffu;
      const int sh = 8 * (s & 3);
      const float lox = (float)(((s < 4 ? n2.x : n2.y) >> sh) & 0xffu) * sx;
      const float loy = (float)(((s < 4 ? n2.z : n2.w) >> sh) & 0xffu) * sy;
      const float loz = (float)(((s < 4 ? n3.x : n3.y) >> sh) & 0xffu) * sz;
      const float hix = (float)(((s < 4 ? n3.z : n3.w) >> sh) & 0xffu) * sx;
      const float hiy = (float)(((s < 4 ? n4.x : n4.y) >> sh) & 0xffu) * sy;
      const float hiz = (float)(((s < 4 ? n4.z : n4.w) >> sh) & 0xffu) * sz;
      // distance from the point (relative to the node origin) to the child box
      const float ex = fmaxf(fmaxf(lox - rx, rx - hix), 0.f);
      const float ey = fmaxf(fmaxf(loy - ry, ry - hiy), 0.f);
      const float ez = fmaxf(fmaxf(loz - rz, rz - hiz), 0.f);
      const float d = sqrtf(ex * ex + ey * ey + ez * ez) * 0.999999f - band;  // lower bound
      cdist[s] = d;
      const bool keep = meta != 0u && d <= prune;
      if (keep && (imask & (1u << s))) inner |= 1u << s;
      // leaf: unary triangle count in bits 5..7, offset from tri_base in bits 0..4
      if (keep && !(imask & (1u << s))) tmask |= (meta >> 5) << (meta & 31u);
    }
    // one triangle per trip for every lane that has one pending (lock step, like the ray kernel)
    while (tmask) {
      const int bit = __ffs((int)tmask) - 1;
      tmask &= tmask - 1u;
      const int ti = (int)(n1.y + (uint32_t)bit);
      const float4 *tp = bvh.tris + (size_t)ti * 3;
      const float4 q0 = __ldg(tp), q1 = __ldg(tp + 1), q2 = __ldg(tp + 2);
      const float df = sqrtf(tri_dist2_f32(q0, q1, q2, px, py, pz));
      res.screened++;
      if (!(df <= prune)) continue;  // (NaN distances of degenerate triangles drop out here)
      if (ANY && df < fbound - 2.f * band) {
        certain = true;
        ci_[0] = ti;
        break;
      }
      // sorted insertion into the candidate list
      float d_in = df;
      int i_in = ti;
#pragma unroll
      for (int k = 0; k < kCand; k++) {
        if (d_in < cd_[k]) {
          const float td = cd_[k];
          const int tt = ci_[k];
          cd_[k] = d_in;
          ci_[k] = i_in;
          d_in = td;
          i_in = tt;
        }
      }
      prune = fminf(prune, cd_[0] + 2.f * band);
    }
    if (ANY && certain) break;
    // push the surviving internal children far-to-near so that the nearest is popped first
    while (inner) {
      int far_s = 0;
      float far_d = -INFINITY;
#pragma unroll
      for (int s = 0; s < 8; s++)
        if (((inner >> s) & 1u) && cdist[s] > far_d) {
          far_d = cdist[s];
          far_s = s;
        }
      inner &= ~(1u << far_s);
      if (far_d > prune) continue;
      const uint32_t child = n1.x + (uint32_t)__popc(imask & ((1u << far_s) - 1u));
      if (sp < kSdfStack) stack[sp++] = make_uint2(child, __float_as_uint(fmaxf(far_d, 0.f)));
    }
  }
  if (ANY && certain) {
    res.tri = ci_[0];
    res.dist = 0.0;
    return;
  }
  // float64 decision among the float32 near-ties, with the reference's Triangle.Closest
#pragma unroll
  for (int k = 0; k < kCand; k++) {
    if (ci_[k] < 0 || cd_[k] > cd_[0] + 2.f * band) continue;
    double cp[3];
    const double dd = tri_closest_f64(bvh.tris + (size_t)ci_[k] * 3, px, py, pz, cp);
    res.exact++;
    if (dd < res.dist) {
      res.dist = dd;
      res.tri = ci_[k];
      res.cp[0] = cp[0], res.cp[1] = cp[1], res.cp[2] = cp[2];
    }
  }
}

// meshSDF.FaceSDF (sdf.go:229-240): sdf > 0 inside (ColliderSolid.Contains, solid.go:292-300:
// inside the bounds and an odd number of crossings along the fixed direction).
__global__ void __launch_bounds__(kSdfBlock)
mesh_sdf_kernel(DeviceBVH bvh, const float *__restrict__ pts3, int64_t n, float *__restrict__ sdf,
                float *__restrict__ closest3, int32_t *__restrict__ face, float *__restrict__ normal3,
                unsigned long long *__restrict__ counters) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint2 stack[kSdfStack];
  const float px = pts3[3 * i], py = pts3[3 * i + 1], pz = pts3[3 * i + 2];
  NearestResult r;
  nearest_triangle<false>(bvh, px, py, pz, (double)INFINITY, r, stack);
  if (counters) {
    atomicAdd(counters, (unsigned long long)r.nodes);
    atomicAdd(counters + 1, (unsigned long long)r.screened);
    atomicAdd(counters + 2, (unsigned long long)r.exact);
  }
  bool inside = false;
  if (bvh.num_tris > 0 && px >= bvh.bmin[0] && px <= bvh.bmax[0] && py >= bvh.bmin[1] && py <= bvh.bmax[1] &&
      pz >= bvh.bmin[2] && pz <= bvh.bmax[2]) {
    RayF ray;
    ray.ox = px, ray.oy = py, ray.oz = pz;
    ray.dx = 0.5224892708603626f;  // collisions.go:124 (float32 like every direction on this path)
    ray.dy = 0.10494477243214506f;
    ray.dz = 0.43558938446126527f;
    ray.tmin = 0.f;
    ray.tmax = INFINITY;
    inside = (count_bvh_hits(bvh.nodes, bvh.tris, bvh.bmin, bvh.bmax, ray) & 1) != 0;
  }
  if (sdf) sdf[i] = (float)(inside ? r.dist : -r.dist);
  if (closest3) {
    closest3[3 * i] = (float)r.cp[0];
    closest3[3 * i + 1] = (float)r.cp[1];
    closest3[3 * i + 2] = (float)r.cp[2];
  }
  int prim = -1;
  if (r.tri >= 0) prim = __float_as_int(__ldg(&bvh.tris[(size_t)r.tri * 3].w));
  if (face) face[i] = prim;
  if (normal3) {
    // meshSDF.NormalSDF (sdf.go:224-227): the face's flat normal (primitives.go:27-33)
    double nx = 0, ny = 0, nz = 0;
    if (r.tri >= 0) {
      const float4 q0 = bvh.tris[(size_t)r.tri * 3], q1 = bvh.tris[(size_t)r.tri * 3 + 1],
                   q2 = bvh.tris[(size_t)r.tri * 3 + 2];
      const D3 v1 = d3((double)q1.x - q0.x, (double)q1.y - q0.y, (double)q1.z - q0.z);
      const D3 v2 = d3((double)q2.x - q0.x, (double)q2.y - q0.y, (double)q2.z - q0.z);
      D3 nn = dcross(v1, v2);
      nn = dscale(nn, 1.0 / dnorm(nn));
      nx = nn.x, ny = nn.y, nz = nn.z;
    }
    normal3[3 * i] = (float)nx;
    normal3[3 * i + 1] = (float)ny;
    normal3[3 * i + 2] = (float)nz;
  }
}

// Collider.SphereCollision (collisions.go:292-303): some triangle closer than r.
__global__ void __launch_bounds__(kSdfBlock)
sphere_collision_kernel(DeviceBVH bvh, const float *__restrict__ centers3, const float *__restrict__ radii,
                        float radius, int64_t n, uint8_t *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint2 stack[kSdfStack];
  const double r = radii ? (double)radii[i] : (double)radius;
  NearestResult res;
  nearest_triangle<true>(bvh, centers3[3 * i], centers3[3 * i + 1], centers3[3 * i + 2], r, res, stack);
  out[i] = (uint8_t)(res.tri >= 0);
}

// ColliderContains(c, p, margin) for margin != 0 (collisions.go:119-134), given the parity.
__global__ void contains_margin_kernel(const uint8_t *__restrict__ parity, const uint8_t *__restrict__ near,
                                       int64_t n, int margin_negative, uint8_t *__restrict__ inside) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool odd = parity[i] != 0, touches = near[i] != 0;
  inside[i] = (uint8_t)(odd ? (margin_negative ? 1 : !touches) : (margin_negative ? touches : 0));
}

}  // namespace

int sdf_stack_capacity() { return kSdfStack; }

void launch_mesh_sdf(const DeviceBVH &bvh, const float *pts3, int64_t n, float *sdf, float *closest3,
                     int32_t *face, float *normal3, unsigned long long *counters, cudaStream_t stream) {
  if (n <= 0) return;
  const unsigned blocks = (unsigned)((n + kSdfBlock - 1) / kSdfBlock);
  mesh_sdf_kernel<<<blocks, kSdfBlock, 0, stream>>>(bvh, pts3, n, sdf, closest3, face, normal3, counters);
}

void launch_sphere_collisions(const DeviceBVH &bvh, const float *centers3, const float *radii, float radius,
                              int64_t n, uint8_t *out, cudaStream_t stream) {
  if (n <= 0) return;
  const unsigned blocks = (unsigned)((n + kSdfBlock - 1) / kSdfBlock);
  sphere_collision_kernel<<<blocks, kSdfBlock, 0, stream>>>(bvh, centers3, radii, radius, n, out);
}

void launch_contains_margin(const uint8_t *parity, const uint8_t *near, int64_t n, bool margin_negative,
                            uint8_t *inside, cudaStream_t stream) {
  if (n <= 0) return;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  contains_margin_kernel<<<blocks, 256, 0, stream>>>(parity, near, n, margin_negative ? 1 : 0, inside);
}

}  // namespace m3d
