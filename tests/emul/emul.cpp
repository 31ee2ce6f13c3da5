// TEST INFRASTRUCTURE ONLY.  Compiles the product's __host__ __device__ traversal core
// (model3d_b200/csrc/trace_core.cuh) and host BVH builder with g++ so that the kernel
// logic can be checked against the oracle on machines without a GPU.  Never shipped,
// never called by the product.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../model3d_b200/csrc/trace_core.cuh"
#include "../../model3d_b200/csrc/wide_bvh.h"

using namespace m3d;

struct Emul {
  WideBVH bvh;
};

extern "C" {
void *emul_build(const float *tris, int64_t n) {
  auto *e = new Emul();
  BuildInput in;
  in.tris = tris;
  in.n = n;
  build_wide_bvh(in, e->bvh, 8);
  return e;
}
void emul_destroy(void *h) { delete (Emul *)h; }
// the bounds-cull predicate of cull_rays_kernel (trace_core.cuh: ray_misses_bounds) per ray
void emul_ray_misses_bounds(const float *org3, const float *dir3, int64_t n, float tmin, float tmax, const float *bmin,
                            const float *bmax, uint8_t *out) {
  for (int64_t i = 0; i < n; i++)
    out[i] = m3d::ray_misses_bounds(org3[3 * i], org3[3 * i + 1], org3[3 * i + 2], tmin, dir3[3 * i], dir3[3 * i + 1],
                                    dir3[3 * i + 2], tmax, bmin, bmax)
                 ? 1
                 : 0;
}
void emul_info(void *h, int64_t *out /*nodes, tris, depth*/, double *sah) {
  auto *e = (Emul *)h;
  out[0] = (int64_t)e->bvh.nodes.size();
  out[1] = (int64_t)e->bvh.tris.size();
  out[2] = e->bvh.max_depth;
  *sah = e->bvh.sah_cost;
}
void emul_trace(void *h, const float *org, const float *dir, int64_t n, float *t, int32_t *prim,
                float *normal, float *bary, int64_t *counters, int refine) {
  auto *e = (Emul *)h;
  const uint4 *nodes = (const uint4 *)e->bvh.nodes.data();
  const float4 *tris = (const float4 *)e->bvh.tris.data();
  int64_t cn = 0, ct = 0;
  for (int64_t i = 0; i < n; i++) {
    RayF r;
    r.ox = org[3 * i]; r.oy = org[3 * i + 1]; r.oz = org[3 * i + 2]; r.tmin = 0.f;
    r.dx = dir[3 * i]; r.dy = dir[3 * i + 1]; r.dz = dir[3 * i + 2]; r.tmax = __builtin_inff();
    HitF hit;
    TraceCounters c{0, 0};
    trace_bvh<true, false>(nodes, tris, e->bvh.bounds_min, e->bvh.bounds_max, r, -1, hit, &c);
    cn += c.nodes; ct += c.tris;
    prim[i] = -1; t[i] = 0;
    if (hit.tri >= 0) {
      const float4 *tri = tris + (size_t)hit.tri * 3;
      int32_t p; memcpy(&p, &tri[0].w, 4);
      prim[i] = p;
      if (refine) {
        HitD d = refine_hit_f64(tri, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz);
        t[i] = (float)(d.t >= 0 ? d.t : hit.t);
        normal[3 * i] = (float)d.nx; normal[3 * i + 1] = (float)d.ny; normal[3 * i + 2] = (float)d.nz;
        bary[3 * i] = (float)d.b0; bary[3 * i + 1] = (float)d.b1; bary[3 * i + 2] = (float)d.b2;
      } else {
        t[i] = hit.t;
        bary[3 * i] = 1.f - hit.b1 - hit.b2; bary[3 * i + 1] = hit.b1; bary[3 * i + 2] = hit.b2;
      }
    }
  }
  if (counters) { counters[0] = cn; counters[1] = ct; }
}
// collect_bvh_hits + count_bvh_hits per ray: counts[n], and the first cap hits of every ray as
// (t, caller triangle id) rows at i * cap
void emul_collect(void *h, const float *org, const float *dir, int64_t n, int cap, int32_t *counts,
                  int32_t *counts2, float *t, int32_t *prim) {
  auto *e = (Emul *)h;
  const uint4 *nodes = (const uint4 *)e->bvh.nodes.data();
  const float4 *tris = (const float4 *)e->bvh.tris.data();
  std::vector<int32_t> ids(cap);
  for (int64_t i = 0; i < n; i++) {
    RayF r;
    r.ox = org[3 * i]; r.oy = org[3 * i + 1]; r.oz = org[3 * i + 2]; r.tmin = 0.f;
    r.dx = dir[3 * i]; r.dy = dir[3 * i + 1]; r.dz = dir[3 * i + 2]; r.tmax = __builtin_inff();
    counts[i] = collect_bvh_hits(nodes, tris, e->bvh.bounds_min, e->bvh.bounds_max, r, cap, t + i * cap, ids.data());
    counts2[i] = count_bvh_hits(nodes, tris, e->bvh.bounds_min, e->bvh.bounds_max, r);
    for (int k = 0; k < counts[i] && k < cap; k++) {
      int32_t p; memcpy(&p, &tris[(size_t)ids[k] * 3].w, 4);
      prim[i * cap + k] = p;
    }
  }
}
}
