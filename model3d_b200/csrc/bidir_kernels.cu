// Bidirectional path tracer kernels (sm_100a), a wavefront restatement of
// render3d/bidir.go:
//   bidir_eye_raygen / bidir_eye_shade      sampleEyePath        bidir.go:161-189
//   bidir_light_raygen / bidir_light_shade  sampleLightPath      bidir.go:191-234
//                                           (+ SampleLight light.go:142-157,254-270,303-311)
//   bidir_connect                           allPathCombinations  bidir.go:476-530, combinePaths
//                                           :532-572, densities :421-471, the MIS weighting and
//                                           Russian roulette of rayColor :101-159
//   bidir_connect_resolve                   the visibility test  bidir.go:144-152
// Path vertices (bptPathVertex bidir.go:311-346) live in HBM as SoA float4 fields.
//
// Numerics: traversal, hit points and BSDF values are float32.  Densities, path throughputs
// and MIS weights are products over up to 2*depth vertices and include the reference's
// Dirac-lobe magnitude 2/cosineEpsilon = 2e8 per specular vertex (material.go:401-462), which
// overflows float32 after 5 vertices: they are carried as (finite, delta-coefficient) pairs per
// vertex and combined in float64, exactly like the reference's float64 arithmetic.
#include <cstdio>

#include "bidir.h"
#include "materials.cuh"
#include "scene_hit.cuh"

#ifndef M3D_BSHADE_MINB
#define M3D_BSHADE_MINB 8
#endif
#ifndef M3D_BCONNECT_MINB
#define M3D_BCONNECT_MINB 8
#endif
// Work list of the connection stage grouped by CHUNK (the 256 consecutive samples of one
// bidir_prefix block) and, inside a chunk, by (i, j) class.  0: one list per class over the whole batch.
#ifndef M3D_CONNECT_CHUNKED
#define M3D_CONNECT_CHUNKED 1
#endif
#ifndef M3D_CONNECT_PARTS
#define M3D_CONNECT_PARTS 32  // connection blocks per chunk (each strides over the chunk's items)
#endif

namespace m3d {

namespace {

constexpr int kBlock = 128;
constexpr double kDeltaMag = 2.0e8;  // 2 / cosineEpsilon
constexpr double kFourPi = 12.566370614359172;

struct BVert {
  V3f point, normal, source, dest;
  V3f bsdf_fin, bsdf_del, emission;
  float sd_fin, sd_del, dd_fin, dd_del;
  float roulette;
  int32_t obj;   // < 0: the emitter vertex of a light path (no material, bidir.go:338-342)
  int32_t surf;  // surface id for the self-intersection guard
};

__device__ __forceinline__ size_t vidx(int field, int D, int64_t cap, int depth, int slot) {
#if M3D_BIDIR_VERT_AOS
  return ((size_t)slot * D + depth) * kBidirVertexFields + field;
#else
  return ((size_t)field * D + depth) * (size_t)cap + slot;
#endif
}

__device__ __forceinline__ void store_vertex(float4 *__restrict__ verts, int D, int64_t cap, int depth, int slot,
                                             const BVert &v) {
  verts[vidx(0, D, cap, depth, slot)] = make_float4(v.point.x, v.point.y, v.point.z, v.roulette);
  verts[vidx(1, D, cap, depth, slot)] = make_float4(v.normal.x, v.normal.y, v.normal.z, __int_as_float(v.obj));
  verts[vidx(2, D, cap, depth, slot)] = make_float4(v.source.x, v.source.y, v.source.z, v.sd_fin);
  verts[vidx(3, D, cap, depth, slot)] = make_float4(v.dest.x, v.dest.y, v.dest.z, v.dd_fin);
  verts[vidx(4, D, cap, depth, slot)] = make_float4(v.bsdf_fin.x, v.bsdf_fin.y, v.bsdf_fin.z, v.sd_del);
  verts[vidx(5, D, cap, depth, slot)] = make_float4(v.bsdf_del.x, v.bsdf_del.y, v.bsdf_del.z, v.dd_del);
  verts[vidx(6, D, cap, depth, slot)] = make_float4(v.emission.x, v.emission.y, v.emission.z, __int_as_float(v.surf));
}

__device__ __forceinline__ BVert load_vertex(const float4 *__restrict__ verts, int D, int64_t cap, int depth,
                                             int slot) {
  const float4 f0 = verts[vidx(0, D, cap, depth, slot)], f1 = verts[vidx(1, D, cap, depth, slot)],
               f2 = verts[vidx(2, D, cap, depth, slot)], f3 = verts[vidx(3, D, cap, depth, slot)],
               f4 = verts[vidx(4, D, cap, depth, slot)], f5 = verts[vidx(5, D, cap, depth, slot)],
               f6 = verts[vidx(6, D, cap, depth, slot)];
  BVert v;
  v.point = v3f(f0.x, f0.y, f0.z);
  v.roulette = f0.w;
  v.normal = v3f(f1.x, f1.y, f1.z);
  v.obj = __float_as_int(f1.w);
  v.source = v3f(f2.x, f2.y, f2.z);
  v.sd_fin = f2.w;
  v.dest = v3f(f3.x, f3.y, f3.z);
  v.dd_fin = f3.w;
  v.bsdf_fin = v3f(f4.x, f4.y, f4.z);
  v.sd_del = f4.w;
  v.bsdf_del = v3f(f5.x, f5.y, f5.z);
  v.dd_del = f5.w;
  v.emission = v3f(f6.x, f6.y, f6.z);
  v.surf = __float_as_int(f6.w);
  return v;
}

// bptPathVertex.EvalMaterial (bidir.go:338-346); tag: Dirac lobe linking source and dest
__device__ __forceinline__ void eval_vertex(const DeviceScene &sc, BVert &v, int tag) {
  v.sd_fin = v.sd_del = v.dd_fin = v.dd_del = 0.f;
  v.bsdf_fin = v.bsdf_del = v3f(0.f, 0.f, 0.f);
  if (v.obj < 0) {
    v.dd_fin = 4.f * fmaxf(0.f, dot(v.dest, v.normal));
    return;
  }
  const MatAt m = material_at(sc, v.obj, v.point);
  const Density sd = mat_source_density(sc, m, v.normal, v.source, v.dest, tag);
  const Density dd = mat_dest_density(sc, m, v.normal, v.source, v.dest, tag);
  v.sd_fin = sd.fin;
  v.sd_del = sd.del;
  v.dd_fin = dd.fin;
  v.dd_del = dd.del;
  v.bsdf_fin = mat_bsdf(sc, m, v.normal, v.source, v.dest);
  v.bsdf_del = mat_bsdf_delta(sc, m, v.normal, v.source, v.dest, tag);
}

// One out-of-line copy for callers that evaluate several vertices (the connection kernel's two junction
// vertices): the material code is ~2,000 instructions per inlined copy.
#ifndef M3D_CONNECT_EVAL_NOINLINE
// 0: inlined twice (4,488 instructions), 1: out of line by reference (round 2: 202 -> 162 Msamples/s),
// 2: out of line by value (3,304 instructions; connection stage 227.6 -> 223.8 ms per 128-spp C5 frame)
#define M3D_CONNECT_EVAL_NOINLINE 2
#endif
#if M3D_CONNECT_EVAL_NOINLINE == 2
// values in, values out (a BVert & would put the whole vertex into local memory: measured 202 -> 162
// Msamples/s in round 2); the scene is reached through the address of the __grid_constant__ parameter
struct VertexEval {
  float sd_fin, sd_del, dd_fin, dd_del;
  V3f bsdf_fin, bsdf_del;
};
static __device__ __noinline__ VertexEval eval_vertex_values(const DeviceScene *scp, int obj, V3f point, V3f normal,
                                                          V3f source, V3f dest) {
  BVert v;
  v.obj = obj;
  v.point = point;
  v.normal = normal;
  v.source = source;
  v.dest = dest;
  eval_vertex(*scp, v, 0);
  VertexEval r;
  r.sd_fin = v.sd_fin;
  r.sd_del = v.sd_del;
  r.dd_fin = v.dd_fin;
  r.dd_del = v.dd_del;
  r.bsdf_fin = v.bsdf_fin;
  r.bsdf_del = v.bsdf_del;
  return r;
}
__device__ __forceinline__ void eval_vertex_call(const DeviceScene &sc, BVert &v) {
  const VertexEval r = eval_vertex_values(&sc, v.obj, v.point, v.normal, v.source, v.dest);
  v.sd_fin = r.sd_fin;
  v.sd_del = r.sd_del;
  v.dd_fin = r.dd_fin;
  v.dd_del = r.dd_del;
  v.bsdf_fin = r.bsdf_fin;
  v.bsdf_del = r.bsdf_del;
}
#elif M3D_CONNECT_EVAL_NOINLINE
static __device__ __noinline__ void eval_vertex_call(const DeviceScene &sc, BVert &v) { eval_vertex(sc, v, 0); }
#else
__device__ __forceinline__ void eval_vertex_call(const DeviceScene &sc, BVert &v) { eval_vertex(sc, v, 0); }
#endif

__device__ __forceinline__ float source_dot(const BVert &v) { return fabsf(dot(v.normal, v.source)); }
__device__ __forceinline__ float dest_dot(const BVert &v) { return fabsf(dot(v.normal, v.dest)); }
__device__ __forceinline__ double full_sd(const BVert &v) { return (double)v.sd_fin + (double)v.sd_del * kDeltaMag; }
__device__ __forceinline__ double full_dd(const BVert &v) { return (double)v.dd_fin + (double)v.dd_del * kDeltaMag; }

// pathEnder.End (bidir.go:282-309).  full = (fullMask, currentRoulette), roul = rouletteMask.
__device__ __forceinline__ bool ender_end(float4 &full, float4 &roul, Rng &g, int i, V3f mask, int min_length,
                                          float cutoff) {
  full.x *= mask.x;
  full.y *= mask.y;
  full.z *= mask.z;
  const float mean = (full.x + full.y + full.z) * (1.f / 3.f);
  if (!(mean == mean) || mean == INFINITY) return true;  // degenerate density: nothing can be carried on
  if (mean < cutoff) {
    const float keep = mean / cutoff;
    if (g.f32() > keep) return true;
    full.w *= 1.f / keep;
  }
  if (min_length != 0 && i + 1 >= min_length) {
    roul.x *= mask.x;
    roul.y *= mask.y;
    roul.z *= mask.z;
    const float mv = fmaxf(fmaxf(roul.x, roul.y), roul.z);
    if (mv < 1.f) {
      roul = make_float4(1.f, 1.f, 1.f, 0.f);
      const float keep = mv;
      if (g.f32() > keep) return true;
      full.w *= 1.f / keep;
    }
  }
  return false;
}

// BSDF * cos / density of the sampled direction; Dirac lobes cancel analytically.
__device__ __forceinline__ V3f sampled_mask(V3f bsdf_fin, V3f bsdf_del, float d_fin, float d_del, float cosv) {
  if (d_del > 0.f) return bsdf_del * (cosv / d_del);
  return bsdf_fin * (cosv / d_fin);  // d_fin == 0 -> inf / NaN, caught by ender_end
}

__device__ __forceinline__ int warp_aggregated_alloc(int *counter) {
  const unsigned m = __activemask();
  const unsigned lane = threadIdx.x & 31u;
  const int leader = __ffs(m) - 1;
  int base = 0;
  if ((int)lane == leader) base = atomicAdd(counter, __popc(m));
  base = __shfl_sync(m, base, leader);
  return base + __popc(m & ((1u << lane) - 1u));
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bidir_eye_raygen_kernel(DeviceCamera cam, DeviceBidirParams bp, PathBatch b, BidirBuffers buf) {
  const int64_t n = (int64_t)b.nP * b.S;
  const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot == 0) {
    buf.counts[0] = (int)n;
    buf.counts[1] = 0;
    buf.counts[2] = 0;
    buf.counts[4] = buf.counts[5] = 0;  // M3D_CONNECT_CHECK: mismatching / checked MIS weights
  }
  if (slot < (int64_t)buf.De * (buf.Dl + 1)) buf.class_counts[slot] = 0;
  if (slot >= n) return;
  const int p = (int)(slot % b.nP), s = (int)(slot / b.nP);
  const int pix = batch_pixel(b, p);
  const int x = pix % b.W, y = pix / b.W;
  double fx = (double)x, fy = (double)y;
  if (bp.antialias != 0.f) {  // ray_renderer.go:118-124
    Rng g;
    g.init(bp.seed, (uint32_t)pix, b.sample0 + (uint32_t)s, 0xA11A5u);
    fx += (double)(bp.antialias * (g.f32() - 0.5f));
    fy += (double)(bp.antialias * (g.f32() - 0.5f));
  }
  fx = (fx - cam.cx) / cam.cx;
  fy = (fy - cam.cy) / cam.cy;
  buf.org[0][slot] = make_float4((float)cam.origin[0], (float)cam.origin[1], (float)cam.origin[2], 0.f);
  buf.dir[0][slot] = make_float4((float)(cam.x[0] * fx + cam.y[0] * fy + cam.z[0]),
                                 (float)(cam.x[1] * fx + cam.y[1] * fy + cam.z[1]),
                                 (float)(cam.x[2] * fx + cam.y[2] * fy + cam.z[2]), INFINITY);
  buf.skip[0][slot] = -1;
  buf.queue[0][slot] = (int32_t)slot;
  buf.ne[slot] = 0;
  buf.nl[slot] = 0;
  buf.ender_full[slot] = make_float4(1.f, 1.f, 1.f, 1.f);
  buf.ender_roul[slot] = make_float4(1.f, 1.f, 1.f, 0.f);
  buf.accum[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// One step of sampleEyePath (EYE) or sampleLightPath's loop (!EYE): resolve the hit, build and
// store the vertex, sample the continuation, apply pathEnder, compact survivors.
template <bool EYE, int SB>  // SB: see path_resolve_kernel
__global__ void __launch_bounds__(kBlock, M3D_BSHADE_MINB)
bidir_shade_kernel(DeviceScene sc, DeviceBidirParams bp, PathBatch b, BidirBuffers buf, int cur, int depth) {
  const int n = buf.counts[cur];
  const unsigned lane = threadIdx.x & 31u;
  const int warps_total = (gridDim.x * kBlock) >> 5;
  const int warp_id = (blockIdx.x * kBlock + threadIdx.x) >> 5;
  const int nxt = cur ^ 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(buf.ray_total, (unsigned long long)n);
  const int max_vertices = EYE ? bp.max_depth : bp.max_light_depth;
  const int vdepth = EYE ? depth : depth + 1;  // index of the vertex this step creates
  float4 *__restrict__ verts = EYE ? buf.ev : buf.lv;
  const int D = EYE ? buf.De : buf.Dl;

  for (int base = warp_id * 32; base < n; base += warps_total * 32) {
    const int q = base + (int)lane;
    bool alive = false;
    int slot = 0;
    BVert v;
    V3f next_dir = v3f(0.f, 0.f, 1.f);
    v.point = v3f(0.f, 0.f, 0.f);
    v.surf = -1;
    if (q < n) {
      slot = buf.queue[cur][q];
      const float4 o = __ldcs(buf.org[cur] + q), d = __ldcs(buf.dir[cur] + q), raw = __ldcs(buf.raw + q);
      const SceneHit h = resolve_scene_hit<SB>(sc, o, d, raw, buf.skip[cur][q], true);
      if (h.obj >= 0) {
        const V3f org = v3f(o.x, o.y, o.z), dir = v3f(d.x, d.y, d.z);
        v.point = org + dir * h.t;
        v.normal = v3f(h.nx, h.ny, h.nz);
        v.obj = h.obj;
        v.surf = h.surf;
        const MatAt m = material_at(sc, h.obj, v.point);
        v.emission = mat_emission(sc, m);
        Rng g;
        g.init(bp.seed, (uint32_t)batch_pixel(b, slot % b.nP), b.sample0 + (uint32_t)(slot / b.nP),
               (EYE ? 0x100u : 0x300u) + (uint32_t)depth);
        int tag = 0;
        if (EYE) {
          v.dest = normalize(dir) * -1.f;                                // bidir.go:171
          v.source = mat_sample_source(sc, m, g, v.normal, v.dest, tag);  // bidir.go:172
          next_dir = v.source * -1.f;
        } else {
          v.source = dir;                                                // bidir.go:213
          v.dest = mat_sample_dest(sc, m, g, v.normal, v.source, tag);    // bidir.go:214
          next_dir = v.dest;
        }
        float4 full = buf.ender_full[slot], roul = buf.ender_roul[slot];
        v.roulette = full.w;
        eval_vertex(sc, v, tag);
        store_vertex(verts, D, buf.cap, vdepth, slot, v);
        (EYE ? buf.ne : buf.nl)[slot] = vdepth + 1;
        const V3f mask = EYE ? sampled_mask(v.bsdf_fin, v.bsdf_del, v.sd_fin, v.sd_del, source_dot(v))
                             : sampled_mask(v.bsdf_fin, v.bsdf_del, v.dd_fin, v.dd_del, dest_dot(v));
        const bool ended = ender_end(full, roul, g, depth, mask, bp.min_depth, bp.cutoff);
        if (!ended && vdepth + 1 < max_vertices) {
          alive = true;
          buf.ender_full[slot] = full;
          buf.ender_roul[slot] = roul;
        }
      }
    }
    const unsigned live = __ballot_sync(0xffffffffu, alive);
    if (live) {
      int pos0 = 0;
      if (lane == (unsigned)(__ffs(live) - 1)) pos0 = atomicAdd(buf.counts + nxt, __popc(live));
      pos0 = __shfl_sync(0xffffffffu, pos0, __ffs(live) - 1);
      if (alive) {
        const int pos = pos0 + __popc(live & ((1u << lane) - 1u));
        // bounceRay (bidir.go:243-255): unit direction, origin eps along it
        buf.org[nxt][pos] = make_float4(v.point.x, v.point.y, v.point.z, bp.eps);
        buf.dir[nxt][pos] = make_float4(next_dir.x, next_dir.y, next_dir.z, INFINITY);
        buf.skip[nxt][pos] = v.surf;
        buf.queue[nxt][pos] = slot;
      }
    }
  }
}
// (Measured and rejected, round 2: splitting this kernel like a bounce of the path tracer -- a resolve
// kernel writing 48-byte hit records and per-kind work lists, then bidir_sample_kernel<EYE, KIND> per
// material kind, 1,400-3,300 instructions each instead of 6,450 -- removes the instruction-fetch stalls
// (43 % of this kernel's stall samples) but not its time: eye + light stages 305 ms vs 290 ms per
// 1024^2 x 128 spp frame.  The stage is bound by the scattered 112-byte vertex stores and the pathEnder
// gathers, and the split adds the hit-record round trip and four launches per step.)

// Box-Muller pair from two uniforms
__device__ __forceinline__ void gauss2(Rng &g, float &a, float &b) {
  const float u1 = fmaxf(g.f32(), 5.9604644775390625e-8f), u2 = g.f32();
  const float r = sqrtf(-2.f * __logf(u1));
  float s, c;
  sincos_2pi(u2, &s, &c);
  a = r * c;
  b = r * s;
}

// sampleLightPath up to its first ray (bidir.go:191-208)
__global__ void __launch_bounds__(256)
bidir_light_raygen_kernel(DeviceScene sc, DeviceBidirParams bp, const DeviceAreaLight *__restrict__ lights,
                          const DeviceLightTri *__restrict__ tris, PathBatch b, BidirBuffers buf) {
  const int64_t n = (int64_t)b.nP * b.S;
  const int64_t slot64 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot64 == 0) {
    buf.counts[0] = bp.max_light_depth > 1 ? (int)n : 0;
    buf.counts[1] = 0;
  }
  if (slot64 >= n) return;
  const int slot = (int)slot64;
  Rng g;
  g.init(bp.seed, (uint32_t)batch_pixel(b, slot % b.nP), b.sample0 + (uint32_t)(slot / b.nP), 0x200u);
  // joinedAreaLight.SampleLight (light.go:303-311): light chosen in proportion to TotalEmission
  int li = 0;
  if (bp.num_lights > 1) {
    const double x = (double)g.f32() * bp.total_light;
    int lo = 0, hi = bp.num_lights;  // sort.SearchFloat64s: smallest index with cumu >= x
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (lights[mid].cumu_total < x) lo = mid + 1; else hi = mid;
    }
    li = lo == bp.num_lights ? lo - 1 : lo;
  }
  const DeviceAreaLight L = lights[li];
  BVert v;
  v.obj = -1;
  v.emission = v3f(L.emission);
  v.roulette = 1.f;
  if (L.kind == SHAPE_SPHERE) {  // SphereAreaLight.SampleLight (light.go:142-157)
    V3f nrm;
    for (;;) {
      float a, bb, c, dmy;
      gauss2(g, a, bb);
      gauss2(g, c, dmy);
      nrm = v3f(a, bb, c);
      const float len = norm(nrm);
      if (len > 0.01f && len < 100.f) {
        nrm = nrm * (1.f / len);
        break;
      }
    }
    v.normal = nrm;
    v.point = v3f(L.center) + nrm * L.radius;
    v.surf = L.surf;
  } else {  // MeshAreaLight.SampleLight (light.go:254-270)
    const double x = (double)g.f32() * L.total_area;
    int lo = 0, hi = L.tri_count;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (tris[L.tri_begin + mid].cumu_area < x) lo = mid + 1; else hi = mid;
    }
    const int ti = L.tri_begin + (lo == L.tri_count ? lo - 1 : lo);
    const DeviceLightTri T = tris[ti];
    const float r1 = sqrtf(g.f32()), r2 = g.f32();
    const float w0 = 1.f - r1, w1 = r1 * (1.f - r2), w2 = r1 * r2;
    v.point = v3f(T.v[0] * w0 + T.v[3] * w1 + T.v[6] * w2, T.v[1] * w0 + T.v[4] * w1 + T.v[7] * w2,
                  T.v[2] * w0 + T.v[5] * w1 + T.v[8] * w2);
    v.normal = v3f(T.n);
    v.surf = T.leaf_index;
  }
  v.source = v.normal * -1.f;
  v.dest = lambert_sample(g, v.normal) * -1.f;  // sampleAngularDest (bidir.go:574-576)
  eval_vertex(sc, v, 0);
  store_vertex(buf.lv, buf.Dl, buf.cap, 0, slot, v);
  buf.nl[slot] = 1;
  buf.org[0][slot] = make_float4(v.point.x, v.point.y, v.point.z, bp.eps);
  buf.dir[0][slot] = make_float4(v.dest.x, v.dest.y, v.dest.z, INFINITY);
  buf.skip[0][slot] = v.surf;
  buf.queue[0][slot] = slot;
  buf.ender_full[slot] = make_float4(1.f, 1.f, 1.f, 1.f);
  buf.ender_roul[slot] = make_float4(1.f, 1.f, 1.f, 0.f);
}

// ---- MIS weights in O(1) per connection ------------------------------------------------------
// The weight of a connection is the sum, over every way t of splitting the JOINED path x_0 .. x_{n-1}
// (light end first) into t light-sampled and n - t eye-sampled vertices, of that strategy's density
// (`densities`, bidir.go:421-471) under the power / balance heuristic (bidir.go:118-131).  The joined
// path of item (i, j) is light[0..j-2], JL, JE, eye[i-2..0], where only the two junction vertices JL, JE
// are re-evaluated (combinePaths, bidir.go:532-572); every other vertex keeps the densities of its own
// sub-path.  With T_t = (density of sampling x_0..x_{t-1} from the light) * area(x_{t-1}, x_t) /
// destDot(x_{t-1}) and acc_t = product of sourceDensity(x_k), k > t, the strategy densities are
// acc_t * T_t (t >= 1) and acc_0, and they factor into per-sub-path partial products times junction terms:
//   t <= j-2 : LT[t] * sd(L_{t+1}) .. sd(L_{j-2}) * sd(JL) * sd(JE) * EP[i-1]
//   t == j-1 : LT[j-1] * sd(JE) * EP[i-1]
//   t == j   : the connection's own density
//   t == j+1 : ld1 * GE[i-1] * EP[i-2],   ld1 = LD[j] * dd(JL) * sourceDot(JE) / destDot(JL)
//   t >= j+2 : ld1 * dd(JE) * RSP[i-2] * RS[i-3] .. RS[m+1] * GE[m+1] * EP[m],   m = n-1-t
//   t == 0   : SLP[j] * sd(JL) * sd(JE) * EP[i-1]
// (LD = lightpre density, EP = eyepre density, LT[t] = LD[t] * area(L_{t-1}, L_t) / destDot(L_{t-1}),
// GE[e] = area(E_e, E_{e-1}) / destDot(E_e), RSP[e] = sourceDot(E_e) / destDot(E_{e+1}), RS[e] =
// dd(E_{e+1}) * RSP[e], SLP[j] = sd(L_1) .. sd(L_{j-2}).)  The heuristic's power is multiplicative, so the
// sums over t <= j-2 and over t >= j+2 are per-sample tables
//   H[j][t_lo] = sum_{t=t_lo}^{j-2} g(LT[t] * sd(L_{t+1}) .. sd(L_{j-2}))
//   K[i][m_lo] = sum_{m=m_lo}^{i-3} g(RS[i-3] .. RS[m+1] * GE[m+1] * EP[m])
// (g(x) = x^ph, x for the balance heuristic) with one-step recurrences in j and i; the lower limits carry
// the reference's depth limits (a strategy needs n - t <= MaxDepth eye vertices and t <= MaxLightDepth
// light vertices).  bidir_prefix_kernel fills the tables (O(depth^2) per sample), a connection reads seven
// entries instead of walking its joined path twice (O(depth) dependent 32-byte gathers per item, 40 % of
// the connection kernel's stall samples, and a 33-entry float64 array in local memory).
// M3D_CONNECT_CHECK=1 (bidir.h) also keeps the path walk and reports items whose two weights differ.
// PHK selects the heuristic at compile time: 0 balance (PowerHeuristic 0: plain sum), 2 the squared
// special case, 1 any other exponent.  pow() is ~100 instructions per call site and the connection kernel
// has a dozen of them: compiled in unconditionally they were a third of its 7,100 instructions, with 28 %
// of the stall samples waiting for instruction fetch.
static __device__ __noinline__ double pow_general(double x, double ph) { return pow(x, ph); }
template <int PHK>
__device__ __forceinline__ double mis_pow(double x, double ph) {
  return PHK == 0 ? x : (PHK == 2 ? x * x : pow_general(x, ph));
}
__host__ __device__ inline int heuristic_kind(double ph) { return ph == 0.0 ? 0 : (ph == 2.0 ? 2 : 1); }
__device__ __forceinline__ double area_between(V3f a, V3f b) {
  const double dx = (double)a.x - b.x, dy = (double)a.y - b.y, dz = (double)a.z - b.z;
  return kFourPi * (dx * dx + dy * dy + dz * dz);
}

#if M3D_CONNECT_CHECK
// compact per-vertex records of the path walk (check build only)
struct Mis {
  double sd, dd;
  float sdot, ddot;
  float px, py, pz;
};
__device__ __forceinline__ Mis mis_of(const BVert &v) {
  Mis m;
  m.sd = full_sd(v);
  m.dd = full_dd(v);
  m.sdot = source_dot(v);
  m.ddot = dest_dot(v);
  m.px = v.point.x;
  m.py = v.point.y;
  m.pz = v.point.z;
  return m;
}
__device__ __forceinline__ void store_mis(const BidirBuffers &buf, int depth, int slot, const Mis &m) {
  const size_t o = (size_t)depth * buf.cap + slot;
  buf.misA[o] = make_float4(m.px, m.py, m.pz, m.sdot);
  buf.misB[o] = make_double2(m.sd, m.dd);
  buf.misC[o] = m.ddot;
}
__device__ __forceinline__ Mis load_mis(const BidirBuffers &buf, int depth, int slot) {
  const size_t o = (size_t)depth * buf.cap + slot;
  const float4 a = buf.misA[o];
  const double2 b = buf.misB[o];
  Mis m;
  m.px = a.x;
  m.py = a.y;
  m.pz = a.z;
  m.sdot = a.w;
  m.sd = b.x;
  m.dd = b.y;
  m.ddot = buf.misC[o];
  return m;
}
__device__ __forceinline__ double out_area(const Mis &a, const Mis &b) {
  const double dx = (double)a.px - b.px, dy = (double)a.py - b.py, dz = (double)a.pz - b.pz;
  return kFourPi * (dx * dx + dy * dy + dz * dz);
}
#endif

struct D3c {
  double x, y, z;
};

// The running products of allPathCombinations (bidir.go:482-483, 493-507, 527-528), which the
// reference carries through its two nested loops, precomputed per sub-path so that every
// (eye prefix i, light prefix j) pair becomes an independent work item:
//   eyepre[i-1]   = (eyeDensity, eyeBSDF.xyz) as they stand when the outer loop reaches i
//   lightpre[j-1] = (density / eyeDensity, lightBSDF.xyz) as they stand when the inner loop reaches j
// One thread per sample; also fills the per-sample MIS tables (see above).
#ifndef M3D_BPREFIX_MINB
#define M3D_BPREFIX_MINB 3  // 80 registers, 3 blocks of 256: C5 prefix stage 68 -> 61 ms per 128-spp frame
#endif
template <int PHK>
__global__ void __launch_bounds__(256, M3D_BPREFIX_MINB)
bidir_prefix_kernel(DeviceBidirParams bp, PathBatch b, BidirBuffers buf) {
  __shared__ int s_cnt[kBidirMaxDepth * (kBidirMaxDepth + 1)], s_base[kBidirMaxDepth * (kBidirMaxDepth + 1)];
  // the thread's current row of the K table, then of the H table: indexed by a loop counter, so as a
  // local array it lived in local memory (14 M local loads per launch of a 128-spp C5 frame)
  __shared__ double s_row[kBidirMaxDepth][256];
  const int n_slots = b.nP * b.S;
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = bp.max_light_depth + 1;
  const int n_classes = bp.max_depth * row;
  for (int c = threadIdx.x; c < n_classes; c += blockDim.x) s_cnt[c] = 0;
  int ne = 0, nl = 0;
  if (slot < n_slots) {
    ne = buf.ne[slot];
    nl = buf.nl[slot];
    const MisTab mt{buf.De, buf.Dl};
    const double ph = bp.power_heuristic;
    auto tab = [&](int entry) -> double & { return buf.mistab[(size_t)entry * buf.cap + slot]; };
    double eye_density = 1.0;
    D3c eye_bsdf = {1.0, 1.0, 1.0};
    double ep_prev = 1.0;             // EP[i-2] while vertex E_{i-1} is handled
    // s_row[m][thread] = K[i+1][m], carried from row to row
    BVert pv;
    for (int i = 1; i <= ne; i++) {
      const BVert v = load_vertex(buf.ev, buf.De, buf.cap, i - 1, slot);
      double *st = buf.eyepre + ((size_t)(i - 1) * buf.cap + slot) * 4;
      st[0] = eye_density;
      st[1] = eye_bsdf.x;
      st[2] = eye_bsdf.y;
      st[3] = eye_bsdf.z;
#if M3D_CONNECT_CHECK
      store_mis(buf, i - 1, slot, mis_of(v));
#endif
      if (i >= 2) {
        // v = E_e, pv = E_{e-1}, e = i-1
        const double ddv = (double)dest_dot(v);
        const double ge = area_between(v.point, pv.point) / ddv;
        const double rsp = (double)source_dot(pv) / ddv;
        tab(mt.ge(i - 1)) = ge;
        tab(mt.rsp(i - 2)) = rsp;
        // K[i+1][m] = g(RS[i-2]) K[i][m] + g(GE[i-1] EP[i-2]),  m = 0 .. i-2
        const double grs = mis_pow<PHK>(full_dd(v) * rsp, ph), gnew = mis_pow<PHK>(ge * ep_prev, ph);
        for (int m = 0; m <= i - 2; m++) {
          const double kv = (m <= i - 3 ? grs * s_row[m][threadIdx.x] : 0.0) + gnew;
          s_row[m][threadIdx.x] = kv;
          tab(mt.k(i + 1, m)) = kv;
        }
      }
      const double sdv = (double)source_dot(v);
      ep_prev = eye_density;
      eye_density *= full_sd(v);
      eye_bsdf.x *= ((double)v.bsdf_fin.x + (double)v.bsdf_del.x * kDeltaMag) * sdv;
      eye_bsdf.y *= ((double)v.bsdf_fin.y + (double)v.bsdf_del.y * kDeltaMag) * sdv;
      eye_bsdf.z *= ((double)v.bsdf_fin.z + (double)v.bsdf_del.z * kDeltaMag) * sdv;
      pv = v;
    }
    double density = 0.0;
    D3c light_bsdf = {0.0, 0.0, 0.0};
    double lt_prev = 0.0, slp = 1.0;  // LT[j-2], SLP[j] while vertex L_{j-1} is handled
    // s_row[t-1][thread] = H[j][t] from here on (the K rows are complete)
    BVert prev;
    for (int j = 1; j <= nl; j++) {
      const BVert lj = load_vertex(buf.lv, buf.Dl, buf.cap, j - 1, slot);
      if (j == 1) {
        density = ((double)lj.emission.x + lj.emission.y + lj.emission.z) / bp.total_light;
        light_bsdf.x = lj.emission.x;
        light_bsdf.y = lj.emission.y;
        light_bsdf.z = lj.emission.z;
      } else {
        // LT[j-1] from LD[j-1] (density before this step); prev = L_{j-2}
        const double lt = density * area_between(prev.point, lj.point) / (double)dest_dot(prev);
        tab(mt.lt(j - 1)) = lt;
        if (j >= 3) {
          const double sdp = full_sd(prev);
          slp *= sdp;
          // H[j][t_lo] = g(sd(L_{j-2})) H[j-1][t_lo] + g(LT[j-2]),  t_lo = 1 .. j-2
          const double gs = mis_pow<PHK>(sdp, ph), gl = mis_pow<PHK>(lt_prev, ph);
          for (int t = 1; t <= j - 2; t++) {
            const double hv = (t <= j - 3 ? gs * s_row[t - 1][threadIdx.x] : 0.0) + gl;
            s_row[t - 1][threadIdx.x] = hv;
            tab(mt.h(j, t)) = hv;
          }
        }
        lt_prev = lt;
        density *= full_dd(prev);
        density *= (double)source_dot(lj) / (double)dest_dot(prev);
        if (j > 2) {
          light_bsdf.x *= (double)prev.bsdf_fin.x + (double)prev.bsdf_del.x * kDeltaMag;
          light_bsdf.y *= (double)prev.bsdf_fin.y + (double)prev.bsdf_del.y * kDeltaMag;
          light_bsdf.z *= (double)prev.bsdf_fin.z + (double)prev.bsdf_del.z * kDeltaMag;
        }
        const double sdj = (double)source_dot(lj);
        light_bsdf.x *= sdj;
        light_bsdf.y *= sdj;
        light_bsdf.z *= sdj;
      }
      tab(mt.slp(j)) = slp;
      double *st = buf.lightpre + ((size_t)(j - 1) * buf.cap + slot) * 4;
      st[0] = density;
      st[1] = light_bsdf.x;
      st[2] = light_bsdf.y;
      st[3] = light_bsdf.z;
#if M3D_CONNECT_CHECK
      store_mis(buf, buf.De + j - 1, slot, mis_of(lj));
#endif
      prev = lj;
    }
  }
  // Work list of the connection stage: one item per (i, j) pair of this sample, packed
  // slot | i << 22 | j << 27 (slots < 2^22, depths <= 16) and grouped by (i, j) class, so that the threads of a warp of the
  // connection kernel walk joined paths of the same length (uniform loops) and read their
  // vertices from neighbouring slots.  Block-level counting sort: count per class in shared memory,
  // place the classes one after the other, scatter.
  __syncthreads();
  for (int i = 1; i <= ne; i++)
    for (int j = 0; j <= nl; j++) atomicAdd(&s_cnt[(i - 1) * row + j], 1);
  __syncthreads();
#if M3D_CONNECT_CHUNKED
  // The block's items stay together (segment blockIdx.x of the work array, classes back to back):
  // the connection blocks of one chunk run next to each other in time, so the ~0.9 MB of path
  // vertices, running products and MIS records of these 256 samples are fetched from HBM once and
  // then served by L2 to all ~40 items per sample (one list per class over the whole batch swept
  // the batch's 10+ GB of path state through L2 once per class).
  if (threadIdx.x < 32) {
    const int per = (n_classes + 31) / 32;
    const int c0 = (int)threadIdx.x * per, c1 = min(c0 + per, n_classes);
    int sum = 0;
    for (int c = c0; c < c1; c++) sum += s_cnt[c];
    int incl = sum;
    for (int off = 1; off < 32; off <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, off);
      if ((int)threadIdx.x >= off) incl += v;
    }
    int run = incl - sum;
    for (int c = c0; c < c1; c++) {
      s_base[c] = run;
      run += s_cnt[c];
      s_cnt[c] = 0;
    }
    if (threadIdx.x == 31) buf.chunk_counts[blockIdx.x] = incl;
  }
  __syncthreads();
  uint32_t *seg = buf.work + (size_t)blockIdx.x * blockDim.x * n_classes;
  for (int i = 1; i <= ne; i++)
    for (int j = 0; j <= nl; j++) {
      const int c = (i - 1) * row + j;
      const int r = atomicAdd(&s_cnt[c], 1);
      seg[s_base[c] + r] = (uint32_t)slot | ((uint32_t)i << 22) | ((uint32_t)j << 27);
    }
#else
  for (int c = threadIdx.x; c < n_classes; c += blockDim.x) {
    s_base[c] = s_cnt[c] > 0 ? atomicAdd(buf.class_counts + c, s_cnt[c]) : 0;
    s_cnt[c] = 0;
  }
  __syncthreads();
  for (int i = 1; i <= ne; i++)
    for (int j = 0; j <= nl; j++) {
      const int c = (i - 1) * row + j;
      const int r = atomicAdd(&s_cnt[c], 1);
      buf.work[(size_t)c * buf.cap + s_base[c] + r] = (uint32_t)slot | ((uint32_t)i << 22) | ((uint32_t)j << 27);
    }
#endif
}

// allPathCombinations (bidir.go:476-530) + rayColor's callback (bidir.go:113-158) for one
// (eye prefix length i, light prefix length j, sample) work item.
template <int PHK>
__device__ __forceinline__ void connect_item(const DeviceScene &sc, const DeviceBidirParams &bp, const PathBatch &b,
                                             const BidirBuffers &buf, uint32_t item) {
  const int slot = (int)(item & 0x3fffffu);
  const int i = (int)((item >> 22) & 31u);
  const int j = (int)(item >> 27);

  const double *est = buf.eyepre + ((size_t)(i - 1) * buf.cap + slot) * 4;
  const double eye_density = est[0];
  const D3c eye_bsdf = {est[1], est[2], est[3]};
  const BVert ev = load_vertex(buf.ev, buf.De, buf.cap, i - 1, slot);
  const int max_ld = bp.max_light_depth;  // already defaulted to max_depth by the host
  const double ph = bp.power_heuristic;
  const int n = i + j;

  const MisTab mt{buf.De, buf.Dl};
  auto tab = [&](int entry) -> double { return buf.mistab[(size_t)entry * buf.cap + slot]; };
  // f(x) = (x / density^((ph-1)/ph))^ph, the heuristic's term of a strategy of density x, normalised like
  // the reference's so that the sums stay inside float64; f(x y) = f(x) g(y)
  auto heur_scale = [&](double density) -> double {
    return PHK == 0 ? 1.0 : (PHK == 2 ? rsqrt(density) : pow_general(density, -(ph - 1.0) / ph));
  };
#if M3D_CONNECT_CHECK
  // the path walk the tables replace: joined path light[0..j-2], JL, JE, eye[i-2..0] for j >= 1,
  // eye[i-1..0] for j == 0; pass 1 from the light end (term[t]), pass 2 from the eye end (acc)
  Mis JL, JE;
  auto at = [&](int k) -> Mis {
    if (j == 0) return load_mis(buf, i - 1 - k, slot);
    if (k < j - 1) return load_mis(buf, buf.De + k, slot);
    if (k == j - 1) return JL;
    if (k == j) return JE;
    return load_mis(buf, i - 1 - (k - j), slot);
  };
  auto walk_weight = [&](double density, double emis0_sum) -> double {
    const double s = heur_scale(density);
    double weight = 0.0;
    auto f = [&](double d) { weight += mis_pow<PHK>(d * s, ph); };
    double term[2 * kBidirMaxDepth + 1];
    if (n > 1) {
      double ld = emis0_sum / bp.total_light;
      Mis m0 = at(0), m1 = at(1);
      term[1] = n - 1 <= bp.max_depth ? ld * out_area(m0, m1) / (double)m0.ddot : 0.0;
      int t = 2;
      for (int k = 0; k + 2 < n; k++, t++) {
        if (k + 1 >= max_ld) break;
        const Mis m2 = at(k + 2);
        ld *= m0.dd;
        ld *= (double)m1.sdot / (double)m0.ddot;
        term[t] = n - (k + 2) <= bp.max_depth ? ld * out_area(m1, m2) / (double)m1.ddot : 0.0;
        m0 = m1;
        m1 = m2;
      }
      double acc = 1.0;
      for (int k = n - 1; k >= 1; k--) {
        if (k < t && term[k] != 0.0) f(acc * term[k]);
        acc *= at(k).sd;
      }
      if (n <= bp.max_depth) f(acc);
    } else if (n <= bp.max_depth) {
      f(1.0);
    }
    return weight;
  };
  auto check_weight = [&](double w, double w_walk) {
    atomicAdd(buf.counts + 5, 1);
    const bool both_bad = !(w > 0.0 && w < INFINITY) && !(w_walk > 0.0 && w_walk < INFINITY);
    if (both_bad || fabs(w - w_walk) <= 1e-9 * fabs(w_walk)) return;
    if (atomicAdd(buf.counts + 4, 1) < 16)
      printf("MIS weight mismatch: slot %d i %d j %d tables %.17g walk %.17g\n", slot, i, j, w, w_walk);
  };
#endif

  if (j == 0) {
    // the eye path itself reached an emitter (bidir.go:486-491)
    if (is_zero(ev.emission)) return;
    const double r = (double)ev.roulette;
    const D3c cur = {ev.emission.x * eye_bsdf.x * r, ev.emission.y * eye_bsdf.y * r, ev.emission.z * eye_bsdf.z * r};
    if (cur.x + cur.y + cur.z < 1e-8) return;
    const double es = (double)ev.emission.x + ev.emission.y + ev.emission.z;
    // pure eye path E_{i-1} .. E_0: strategy 0 has density EP[i-1]; strategies t = 1 .. min(i-1,
    // MaxLightDepth) are f(es / total) K[i+1][m_lo] (the t >= j+2 sum with nothing re-evaluated)
    const double s = heur_scale(eye_density);
    double w = mis_pow<PHK>(eye_density * s, ph);  // n = i <= MaxDepth always
    if (i >= 2) {
      const int m_lo = max(0, i - 1 - max_ld);
      if (m_lo <= i - 2) w += mis_pow<PHK>(es / bp.total_light * s, ph) * tab(mt.k(i + 1, m_lo));
    }
#if M3D_CONNECT_CHECK
    check_weight(w, walk_weight(eye_density, es));
#endif
    if (!(w > 0.0 && w < INFINITY)) return;
    float *a = reinterpret_cast<float *>(buf.accum + slot);
    atomicAdd(a, (float)(cur.x / w));
    atomicAdd(a + 1, (float)(cur.y / w));
    atomicAdd(a + 2, (float)(cur.z / w));
    return;
  }
  // j >= 1 (bidir.go:493-525)
  const double *lst = buf.lightpre + ((size_t)(j - 1) * buf.cap + slot) * 4;
  const double density = eye_density * lst[0];
  const D3c light_bsdf = {lst[1], lst[2], lst[3]};
  const BVert lj = load_vertex(buf.lv, buf.Dl, buf.cap, j - 1, slot);
#if M3D_CONNECT_CHECK
  const float4 l0e = buf.lv[vidx(6, buf.Dl, buf.cap, 0, slot)];
  const double l0_sum = (double)l0e.x + l0e.y + l0e.z;
#endif
  const V3f diff = lj.point - ev.point;
  const float dist2 = dot(diff, diff);
  const float dist = sqrtf(dist2);
  if (!(dist > 0.f)) return;
  // combinePaths (bidir.go:544-566): both junction vertices re-evaluated for the new edge
  BVert jl = lj;
  jl.dest = (ev.point - lj.point) * (1.f / dist);
  eval_vertex_call(sc, jl);
  BVert je = ev;
  je.source = jl.dest;
  eval_vertex_call(sc, je);
  const float dd = dest_dot(jl);
  const float sd = source_dot(je);
  if (!(dd > 0.f && sd > 0.f)) return;
  const double cur_density = density * (kFourPi * (double)dist2) / (double)dd;
  const double scale = (double)sd * (double)lj.roulette * (double)ev.roulette;
  D3c inten = {eye_bsdf.x * light_bsdf.x * scale * (double)je.bsdf_fin.x,
               eye_bsdf.y * light_bsdf.y * scale * (double)je.bsdf_fin.y,
               eye_bsdf.z * light_bsdf.z * scale * (double)je.bsdf_fin.z};
  if (j > 1) {
    inten.x *= (double)jl.bsdf_fin.x;
    inten.y *= (double)jl.bsdf_fin.y;
    inten.z *= (double)jl.bsdf_fin.z;
  }
  if (inten.x + inten.y + inten.z < 1e-8) return;
  // the strategies of the joined path (see "MIS weights in O(1) per connection")
  const double jl_sd = full_sd(jl), jl_dd = full_dd(jl), je_sd = full_sd(je), je_dd = full_dd(je);
  const double s = heur_scale(cur_density);
  // t == j, this connection's own strategy (the edge length in float64 like every other edge of the sums)
  double w = mis_pow<PHK>(lst[0] * area_between(jl.point, je.point) / (double)dd * eye_density * s, ph);
  const double tail = je_sd * eye_density;                  // sd(JE) * EP[i-1]
  const int t_lo = max(1, n - bp.max_depth);
  if (j >= 2) {
    const double tail2 = jl_sd * tail;
    if (n <= bp.max_depth) w += mis_pow<PHK>(tab(mt.slp(j)) * tail2 * s, ph);            // t == 0
    if (j >= 3 && t_lo <= j - 2) w += mis_pow<PHK>(tail2 * s, ph) * tab(mt.h(j, t_lo));  // t <= j-2
    if (t_lo <= j - 1) w += mis_pow<PHK>(tab(mt.lt(j - 1)) * tail * s, ph);              // t == j-1
  } else if (n <= bp.max_depth) {
    w += mis_pow<PHK>(tail * s, ph);  // t == 0 with j == 1: x_0 = JL is not eye-sampled
  }
  if (i >= 2 && j + 1 <= max_ld) {
    const double ld1 = lst[0] * jl_dd * (double)sd / (double)dd;  // sd = sourceDot(JE), dd = destDot(JL)
    w += mis_pow<PHK>(ld1 * tab(mt.ge(i - 1)) * buf.eyepre[((size_t)(i - 2) * buf.cap + slot) * 4] * s, ph);  // t == j+1
    if (i >= 3 && j + 2 <= max_ld) {
      const int m_lo = max(0, n - 1 - max_ld);
      if (m_lo <= i - 3) w += mis_pow<PHK>(ld1 * je_dd * tab(mt.rsp(i - 2)) * s, ph) * tab(mt.k(i, m_lo));  // t >= j+2
    }
  }
#if M3D_CONNECT_CHECK
  JL = mis_of(jl);
  JE = mis_of(je);
  check_weight(w, walk_weight(cur_density, l0_sum));
#endif
  if (!(w > 0.0 && w < INFINITY)) return;
  D3c color = {inten.x / w, inten.y / w, inten.z / w};
  const double brightness = fmax(fmax(color.x, color.y), color.z);
  if (bp.roulette_delta > 0.0 && brightness < bp.roulette_delta) {  // bidir.go:133-142
    const double keep = brightness / bp.roulette_delta;
    Rng g;
    g.init(bp.seed, (uint32_t)batch_pixel(b, slot % b.nP), b.sample0 + (uint32_t)(slot / b.nP),
           0x1000u + (uint32_t)(i * 64 + j));
    if ((double)g.f32() > keep) return;
    color.x /= keep;
    color.y /= keep;
    color.z /= keep;
  }
  if (!(color.x > 0.0 || color.y > 0.0 || color.z > 0.0)) return;
  // visibility ray eye vertex -> light vertex (bidir.go:144-152): blocked iff something lies
  // strictly between the end points; both end surfaces are excluded (start: skip id, end:
  // parameter interval) since float32 cannot express the reference's 1e-8 offsets
  const int pos = warp_aggregated_alloc(buf.counts + 2);
  const V3f dirn = diff * (1.f / dist);
  // with a user Epsilon: the ray starts eps after the eye vertex and is blocked iff
  // Scale < dist - 2 eps from there (bidir.go:144-152), i.e. t in [eps, dist - eps)
  buf.corg[pos] = make_float4(ev.point.x, ev.point.y, ev.point.z, bp.eps);
  buf.cdir[pos] = make_float4(dirn.x, dirn.y, dirn.z, fminf(dist * (1.f - 2e-4f), dist - bp.eps));
  buf.cskip[pos] = ev.surf;
  buf.cpay[pos] = make_float4((float)color.x, (float)color.y, (float)color.z, __int_as_float(slot));
}

// One thread per work item; threads of a warp share (i, j) except where two classes meet.
template <int PHK>
__global__ void __launch_bounds__(kBlock, M3D_BCONNECT_MINB)
bidir_connect_kernel(const __grid_constant__ DeviceScene sc, DeviceBidirParams bp, PathBatch b, BidirBuffers buf) {
#if M3D_CONNECT_CHUNKED
  // M3D_CONNECT_PARTS consecutive blocks share a chunk and stride over its items
  const int chunk = (int)(blockIdx.x / M3D_CONNECT_PARTS), part = (int)(blockIdx.x % M3D_CONNECT_PARTS);
  const int count = buf.chunk_counts[chunk];
  const uint32_t *seg = buf.work + (size_t)chunk * 256 * (bp.max_depth * (bp.max_light_depth + 1));
  for (int k = part * kBlock + (int)threadIdx.x; k < count; k += M3D_CONNECT_PARTS * kBlock)
    connect_item<PHK>(sc, bp, b, buf, seg[k]);
#else
  // grid: x over the items of a class, y = class; most blocks of a sparsely filled class exit here
  const int cls = (int)blockIdx.y, rank = (int)(blockIdx.x * kBlock + threadIdx.x);
  if (rank >= buf.class_counts[cls]) return;
  connect_item<PHK>(sc, bp, b, buf, buf.work[(size_t)cls * buf.cap + rank]);
#endif
}

template <int SB>
__global__ void __launch_bounds__(256)
bidir_connect_resolve_kernel(DeviceScene sc, BidirBuffers buf) {
  const int n = buf.counts[2];
  const int stride = gridDim.x * blockDim.x;
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(buf.ray_total, (unsigned long long)n);
#if M3D_CONNECT_CHECK
  if (blockIdx.x == 0 && threadIdx.x == 0)
    printf("MIS check: %d of %d connection weights differ from the path walk\n", buf.counts[4], buf.counts[5]);
#endif
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
    const float4 pay = buf.cpay[q];
    const SceneHit h = resolve_scene_hit<SB>(sc, buf.corg[q], buf.cdir[q], buf.craw[q], buf.cskip[q], false);
    if (h.obj >= 0) continue;  // blocked
    float *a = reinterpret_cast<float *>(buf.accum + __float_as_int(pay.w));
    atomicAdd(a, pay.x);
    atomicAdd(a + 1, pay.y);
    atomicAdd(a + 2, pay.z);
  }
}

template <class K>
int persistent_grid(K kernel, int block, int64_t n) {
  int x = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&x, kernel, block, 0);
  int64_t grid = (int64_t)device_sm_count() * (x > 0 ? x : 1);
  const int64_t want = (n + block - 1) / block;
  if (grid > want) grid = want;
  return (int)(grid < 1 ? 1 : grid);
}

}  // namespace

void launch_bidir_eye_raygen(const DeviceCamera &cam, const DeviceBidirParams &bp, const PathBatch &b,
                             const BidirBuffers &buf, cudaStream_t stream) {
  const int64_t n = (int64_t)b.nP * b.S;
  if (n <= 0) return;
  bidir_eye_raygen_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(cam, bp, b, buf);
}

void launch_bidir_eye_shade(const DeviceScene &sc, const DeviceBidirParams &bp, const PathBatch &b,
                            const BidirBuffers &buf, int cur, int depth, cudaStream_t stream) {
  if (sc.shape_bvh.nodes) {
    const int grid = persistent_grid(bidir_shade_kernel<true, 2>, kBlock, (int64_t)b.nP * b.S);
    bidir_shade_kernel<true, 2><<<grid, kBlock, 0, stream>>>(sc, bp, b, buf, cur, depth);
  } else {
    const int grid = persistent_grid(bidir_shade_kernel<true, 1>, kBlock, (int64_t)b.nP * b.S);
    bidir_shade_kernel<true, 1><<<grid, kBlock, 0, stream>>>(sc, bp, b, buf, cur, depth);
  }
}

void launch_bidir_light_raygen(const DeviceScene &sc, const DeviceBidirParams &bp, const DeviceAreaLight *lights,
                               const DeviceLightTri *tris, const PathBatch &b, const BidirBuffers &buf,
                               cudaStream_t stream) {
  const int64_t n = (int64_t)b.nP * b.S;
  if (n <= 0) return;
  bidir_light_raygen_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(sc, bp, lights, tris, b, buf);
}

void launch_bidir_light_shade(const DeviceScene &sc, const DeviceBidirParams &bp, const PathBatch &b,
                              const BidirBuffers &buf, int cur, int depth, cudaStream_t stream) {
  if (sc.shape_bvh.nodes) {
    const int grid = persistent_grid(bidir_shade_kernel<false, 2>, kBlock, (int64_t)b.nP * b.S);
    bidir_shade_kernel<false, 2><<<grid, kBlock, 0, stream>>>(sc, bp, b, buf, cur, depth);
  } else {
    const int grid = persistent_grid(bidir_shade_kernel<false, 1>, kBlock, (int64_t)b.nP * b.S);
    bidir_shade_kernel<false, 1><<<grid, kBlock, 0, stream>>>(sc, bp, b, buf, cur, depth);
  }
}

void launch_bidir_prefix(const DeviceBidirParams &bp, const PathBatch &b, const BidirBuffers &buf,
                         cudaStream_t stream) {
  const int64_t n = (int64_t)b.nP * b.S;
  if (n <= 0) return;
  const unsigned grid = (unsigned)((n + 255) / 256);
  switch (heuristic_kind(bp.power_heuristic)) {
    case 0: bidir_prefix_kernel<0><<<grid, 256, 0, stream>>>(bp, b, buf); break;
    case 2: bidir_prefix_kernel<2><<<grid, 256, 0, stream>>>(bp, b, buf); break;
    default: bidir_prefix_kernel<1><<<grid, 256, 0, stream>>>(bp, b, buf); break;
  }
}

void launch_bidir_connect(const DeviceScene &sc, const DeviceBidirParams &bp, const PathBatch &b,
                          const BidirBuffers &buf, cudaStream_t stream) {
  const int64_t n = (int64_t)b.nP * b.S;
  if (n <= 0) return;
#if M3D_CONNECT_CHUNKED
  const unsigned grid = (unsigned)((n + 255) / 256) * M3D_CONNECT_PARTS;  // chunks of bidir_prefix_kernel's 256 samples
#else
  const dim3 grid((unsigned)((n + kBlock - 1) / kBlock), (unsigned)(bp.max_depth * (bp.max_light_depth + 1)));
#endif
  switch (heuristic_kind(bp.power_heuristic)) {
    case 0: bidir_connect_kernel<0><<<grid, kBlock, 0, stream>>>(sc, bp, b, buf); break;
    case 2: bidir_connect_kernel<2><<<grid, kBlock, 0, stream>>>(sc, bp, b, buf); break;
    default: bidir_connect_kernel<1><<<grid, kBlock, 0, stream>>>(sc, bp, b, buf); break;
  }
}

void launch_bidir_connect_resolve(const DeviceScene &sc, const BidirBuffers &buf, cudaStream_t stream) {
  const int grid = device_sm_count() * 8;
  if (sc.shape_bvh.nodes)
    bidir_connect_resolve_kernel<2><<<grid, 256, 0, stream>>>(sc, buf);
  else
    bidir_connect_resolve_kernel<1><<<grid, 256, 0, stream>>>(sc, buf);
}

}  // namespace m3d
