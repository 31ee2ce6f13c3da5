// Wavefront path tracer kernels (sm_100a):
//   path_raygen_kernel          rayRenderer.Render's per-sample camera ray with antialias
//                               jitter (render3d/ray_renderer.go:29-31,118-124, camera.go:74-82)
//   path_shade_kernel           one level of RecursiveRayTracer.recurse (raytrace.go:138-181):
//                               JoinedObject.Cast closest-of + float64 hit refinement, emission /
//                               ambient, point-light shadow rays, sampleNextSource /
//                               sourceDensity with focus points (raytrace.go:183-215,
//                               focus_point.go:44-177), throughput update, cutoff, and
//                               ballot/popc compaction of the surviving paths into the next queue
//   path_shadow_resolve_kernel  shadow test `hit && Scale < 1` (raytrace.go:157-163)
//   path_flush_kernel           colorSum (+ squares) per pixel (ray_renderer.go:125-127,150)
// Traversal between the stages is trace_first_hit_kernel (trace_kernels.cu).
#include <algorithm>

// SFU-only square roots / normalisations / sincos in the sampling helpers (materials.cuh) are a gain for
// the one large shading kernel of the bidirectional tracer and a loss here: measured on C3, sample stage
// of a 1024^2 x 64 spp frame, 11.2-11.8 ms with the IEEE forms against 13.5-15.4 ms with the SFU forms
// (same rays per sample; profiles/r2c_mc_math.log), C4 6.6 against 7.3 ms.
#ifndef M3D_FAST_SQRT
#define M3D_FAST_SQRT 0
#endif
#ifndef M3D_FAST_NORMALIZE
#define M3D_FAST_NORMALIZE 0
#endif
#ifndef M3D_FAST_SINCOS
#define M3D_FAST_SINCOS 0
#endif
#include "materials.cuh"
#include "path.h"
#include "scene_hit.cuh"

namespace m3d {

namespace {

constexpr int kShadeBlock = 128;

// Per-thread asynchronous copies global -> shared (LDGSTS): the streamed inputs of the bounce kernels
// are fetched two tiles ahead without holding registers.  Every thread reads back only what it copied
// itself, so cp.async.wait_group orders the data and no block barrier is needed.
__device__ __forceinline__ void cp_async_16(void *smem, const void *gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_4(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
#ifndef M3D_RESOLVE_STAGES
#define M3D_RESOLVE_STAGES 3  // tiles of streamed inputs in flight per block (1: plain loads)
#endif

__device__ __forceinline__ bool focus_applies(const DeviceFocus &f, int material) {
  return material < 64 ? ((f.mask >> material) & 1ull) != 0ull : false;
}

// focus_point.go:155-163
__device__ __forceinline__ V3f sample_around_uniform(Rng &g, float min_cos, V3f direction) {
  const float cos_lat = 1.f - g.f32() * (1.f - min_cos);
  const float sin_lat = sqrt_fast(fmaxf(0.f, 1.f - cos_lat * cos_lat));
  float sl, cl;
  sincos_2pi(g.f32(), &sl, &cl);
  const Basis b = ortho_basis(direction);
  return direction * cos_lat + (b.x * cl + b.z * sl) * sin_lat;
}

// Whether focus point f overrides the material's sampler at `point`, and its direction info.
// PhongFocusPoint (focus_point.go:46-64): falls back when Target == point or the filter
// rejects; SphereFocusPoint (focus_point.go:87-153): when inside the sphere or rejected.
__device__ __forceinline__ bool focus_active(const DeviceFocus &f, int material, V3f point, V3f &dir,
                                             float &min_cos) {
  if (!focus_applies(f, material)) return false;
  const V3f diff = point - v3f(f.target);
  const float d2 = dot(diff, diff);
  if (d2 == 0.f) return false;  // Target == point (inside any sphere as well)
  const float inv_d = rsqrtf(d2);  // SFU reciprocal square root: Monte-Carlo sampling, see sqrt_fast
  dir = diff * inv_d;
  if (f.kind == M3D_FOCUS_PHONG) {
    min_cos = 0.f;
    return true;
  }
  if (d2 < f.radius * f.radius) return false;
  const float ratio = f.radius * inv_d;
  min_cos = sqrt_fast(fmaxf(0.f, 1.f - ratio * ratio));
  return true;
}

__device__ __forceinline__ float focus_density(const DeviceFocus &f, V3f dir, float min_cos, V3f source) {
  if (f.kind == M3D_FOCUS_PHONG) return density_around_direction(f.alpha, dir, source);
  // focus_point.go:172-177
  return dot(dir, source) < min_cos ? 0.f : 2.f / (1.f - min_cos);
}

__global__ void __launch_bounds__(256)
path_raygen_kernel(DeviceCamera cam, DevicePathParams pp, PathBatch b, PathBuffers buf) {
  const int64_t n = (int64_t)b.nP * b.S;
  const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot == 0) {
    buf.counts[0] = (int)n;
    buf.counts[1] = 0;
    buf.counts[2] = 0;
    buf.counts[4] = buf.counts[5] = buf.counts[6] = buf.counts[7] = 0;
  }
  if (slot >= n) return;
  const int p = (int)(slot % b.nP), s = (int)(slot / b.nP);
  const int pix = batch_pixel(b, p);
  const int x = pix % b.W, y = pix / b.W;
  double fx = (double)x, fy = (double)y;
  if (pp.antialias != 0.f) {  // ray_renderer.go:118-124
    Rng g;
    g.init(pp.seed, (uint32_t)pix, b.sample0 + (uint32_t)s, 0xA11A5u);
    fx += (double)(pp.antialias * (g.f32() - 0.5f));
    fy += (double)(pp.antialias * (g.f32() - 0.5f));
  }
  fx = (fx - cam.cx) / cam.cx;
  fy = (fy - cam.cy) / cam.cy;
  buf.org[0][slot] = make_float4((float)cam.origin[0], (float)cam.origin[1], (float)cam.origin[2], 0.f);
  buf.dir[0][slot] = make_float4((float)(cam.x[0] * fx + cam.y[0] * fy + cam.z[0]),
                                 (float)(cam.x[1] * fx + cam.y[1] * fy + cam.z[1]),
                                 (float)(cam.x[2] * fx + cam.y[2] * fy + cam.z[2]), INFINITY);
  buf.skip[0][slot] = -1;
  buf.queue[0][slot] = (int32_t)slot;
  buf.thr[slot] = make_float4(1.f, 1.f, 1.f, 0.f);
  buf.accum[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// Stage 1 of a bounce: resolve the hits of queue `cur` (closest of BVH hit and analytic shapes),
// add emission / ambient, emit the shadow rays, and -- unless the path ends here -- write a hit
// record and append the queue position to the work list of the hit material's KIND.  The
// sampling stage then runs one kernel per material kind over its own list, so that warps are
// coherent in material code and every kernel's instruction footprint fits the instruction
// cache (the fused shade kernel of the first version stalled on instruction fetch 2/3 of the
// time, profiles/ncu_path_r1.md).
// LIGHTS = false (no point lights: C3 / C4 / C5) compiles the shadow-ray and BSDF evaluation out.
#ifndef M3D_RESOLVE_MINB
#define M3D_RESOLVE_MINB 4
#endif
#ifndef M3D_SAMPLE_MINB
#define M3D_SAMPLE_MINB 8
#endif
// SB: 1 = the scene's few analytic shapes are tested one by one, 2 = scenes with an object-level
// hierarchy (many shapes / mesh instances) walk it (resolve_scene_hit)
#ifndef M3D_RESOLVE_SPHERES_MINB
#define M3D_RESOLVE_SPHERES_MINB 4
#endif
template <bool LIGHTS, int SB>  // SB = 3: scenes whose analytic shapes are all spheres (no point lights)
__global__ void __launch_bounds__(kShadeBlock, SB == 3 ? M3D_RESOLVE_SPHERES_MINB : M3D_RESOLVE_MINB)
path_resolve_kernel(DeviceScene sc, DevicePathParams pp, const DevicePointLight *__restrict__ lights, PathBatch b,
                    PathBuffers buf, int cur, int depth) {
  const int n = buf.counts[cur];
  const unsigned lane = threadIdx.x & 31u;
  const float4 *__restrict__ org_in = buf.org[cur];
  const float4 *__restrict__ dir_in = buf.dir[cur];
  const int32_t *__restrict__ skip_in = buf.skip[cur];
  const int32_t *__restrict__ queue_in = buf.queue[cur];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    atomicAdd(buf.ray_total, (unsigned long long)n);
    buf.counts[2] = n * pp.num_lights;
  }

  // List appends of the spheres-only instantiation are aggregated per block (one global atomic per block
  // and kind instead of one per warp: the per-kind counters are single addresses that every warp of the
  // grid adds to; measured on C3: resolve stage 14.6 -> 13.5 ms per 1024^2 x 64 spp frame).  The general
  // instantiation keeps the per-warp form: its rare float64 shape tests make the block barriers cost more
  // than the atomics (C4: 12.1 -> 12.9 ms with them).
  constexpr bool kBlockAgg = SB == 3;
  __shared__ int s_wcnt[2][4][kShadeBlock / 32];
  __shared__ int s_kbase[2][4];
  const int wib = (int)(threadIdx.x >> 5);
  int parity = 0;
  constexpr int kStages = M3D_RESOLVE_STAGES;
  __shared__ float4 s_in[kStages > 1 ? kStages : 1][3][kShadeBlock];  // origin, direction, raw hit
  __shared__ int32_t s_ids[kStages > 1 ? kStages : 1][2][kShadeBlock];  // slot, skip id
  const int stride = (int)gridDim.x * kShadeBlock;
  auto prefetch_tile = [&](int tile_base, int st) {
    const int qq = tile_base + (int)threadIdx.x;
    if (qq < n) {
      cp_async_16(&s_in[st][0][threadIdx.x], org_in + qq);
      cp_async_16(&s_in[st][1][threadIdx.x], dir_in + qq);
      cp_async_16(&s_in[st][2][threadIdx.x], buf.raw + qq);
      cp_async_4(&s_ids[st][0][threadIdx.x], queue_in + qq);
      cp_async_4(&s_ids[st][1][threadIdx.x], skip_in + qq);
    }
    cp_async_commit();  // (possibly empty: the group count per trip stays fixed)
  };
  int st_use = 0, st_fill = kStages - 1;
  if (kStages > 1) {
#pragma unroll
    for (int k = 0; k < kStages - 1; k++) {
      // (int64: the look-ahead may pass INT_MAX for the largest batches)
      const int64_t tb = (int64_t)blockIdx.x * kShadeBlock + (int64_t)k * stride;
      prefetch_tile(tb < n ? (int)tb : n, k);
    }
  }
  for (int bbase = (int)blockIdx.x * kShadeBlock; bbase < n; bbase += stride, parity ^= 1) {
    const int q = bbase + (int)threadIdx.x;
    int kind = -1;  // material kind whose sampler continues this path, -1: the path ends
    if (kStages > 1) {
      const int64_t tb = (int64_t)bbase + (int64_t)(kStages - 1) * stride;
      prefetch_tile(tb < n ? (int)tb : n, st_fill);
      cp_async_wait<kStages - 1>();  // this trip's tile has landed
    }
    if (q < n) {
      int slot, skip_id;
      float4 o, d, raw;
      if (kStages > 1) {
        slot = s_ids[st_use][0][threadIdx.x];
        skip_id = s_ids[st_use][1][threadIdx.x];
        o = s_in[st_use][0][threadIdx.x];
        d = s_in[st_use][1][threadIdx.x];
        raw = s_in[st_use][2][threadIdx.x];
      } else {
        slot = queue_in[q];
        skip_id = skip_in[q];
        o = __ldcs(org_in + q);
        d = __ldcs(dir_in + q);
        raw = __ldcs(buf.raw + q);
      }
      // float32 hit evaluation: Monte-Carlo parity is statistical, the float64 refinement of the
      // first-hit API (1e-5 on t and normals) is not needed here; shapes stay float64
      const SceneHit h = resolve_scene_hit<SB>(sc, o, d, raw, skip_id, false);
      if (LIGHTS && pp.num_lights > 0 && h.obj < 0) {
        // no shadow rays for a miss: give the slots an empty parameter interval
        for (int l = 0; l < pp.num_lights; l++) {
          const size_t si = (size_t)q * pp.num_lights + l;
          buf.sorg[si] = make_float4(0.f, 0.f, 0.f, 1.f);
          buf.sdir[si] = make_float4(0.f, 0.f, 1.f, -1.f);
          buf.sskip[si] = -1;
          buf.spay[si] = make_float4(0.f, 0.f, 0.f, __int_as_float(slot));
        }
      }
      if (h.obj >= 0) {
        const V3f org = v3f(o.x, o.y, o.z), dir = v3f(d.x, d.y, d.z), nrm = v3f(h.nx, h.ny, h.nz);
        const V3f point = org + dir * h.t;
        const MatAt m = material_at(sc, h.obj, point);
        const V3f dest = dir * -rsqrtf(dot(dir, dir));  // (one SFU op; measured faster here than the IEEE form)
        // raytrace.go:150-155
        V3f color = mat_emission(sc, m);
        if (depth == 0) color = color + mat_ambient(sc, m);
        // the throughput is a random 16-byte gather: only paths that add light (or cast shadow
        // rays) need it here
        V3f tv = v3f(0.f, 0.f, 0.f);
        if ((LIGHTS && pp.num_lights > 0) || !is_zero(color)) {
          const float4 thr = buf.thr[slot];
          tv = v3f(thr.x, thr.y, thr.z);
        }
        if (!is_zero(color)) {
          float4 a = buf.accum[slot];
          a.x += tv.x * color.x;
          a.y += tv.y * color.y;
          a.z += tv.z * color.z;
          buf.accum[slot] = a;
        }
        // raytrace.go:156-168: one shadow ray per point light, un-normalised direction,
        // occluded iff a hit has Scale < 1
        for (int l = 0; LIGHTS && l < pp.num_lights; l++) {
          const DevicePointLight lt = lights[l];
          const V3f lo = v3f(lt.origin);
          const V3f light_dir = lo - point;
          const V3f brdf = mat_bsdf(sc, m, nrm, normalize(point - lo), dest);
          const V3f c = shade_collision(lt, nrm, light_dir) * brdf * tv;
          const size_t si = (size_t)q * pp.num_lights + l;
          const bool useful = !is_zero(c);
          // bounceRay(point, lightDirection): the origin moves eps along the unit direction (raytrace.go:217-229)
          const float s_tmin = pp.eps > 0.f ? pp.eps * rsqrtf(dot(light_dir, light_dir)) : 0.f;
          buf.sorg[si] = make_float4(point.x, point.y, point.z, useful ? s_tmin : 1.f);
          buf.sdir[si] = make_float4(light_dir.x, light_dir.y, light_dir.z, useful ? 1.f : -1.f);
          buf.sskip[si] = h.surf;
          buf.spay[si] = make_float4(c.x, c.y, c.z, __int_as_float(slot));
        }
        if (depth < pp.max_depth) {
          kind = sc.materials[m.index].kind;
          buf.hrA[q] = make_float4(point.x, point.y, point.z, __int_as_float(h.surf));
          buf.hrB[q] = make_float4(nrm.x, nrm.y, nrm.z, __int_as_float(m.index));  // the material, not the object:
                                                                                 // one gather less in the sampler
          buf.hrC[q] = make_float4(dest.x, dest.y, dest.z, __int_as_float(slot));
        }
      }
    }
    if (kBlockAgg) {
      // append to the per-kind work lists: one global atomic per block and kind
      unsigned mine = 0u;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const unsigned m = __ballot_sync(0xffffffffu, kind == k);
        if (lane == 0) s_wcnt[parity][k][wib] = __popc(m);
        if (kind == k) mine = m;
      }
      __syncthreads();
      if (threadIdx.x < 4) {
        int tot = 0;
#pragma unroll
        for (int w = 0; w < kShadeBlock / 32; w++) tot += s_wcnt[parity][threadIdx.x][w];
        s_kbase[parity][threadIdx.x] = tot ? atomicAdd(buf.counts + 4 + (int)threadIdx.x, tot) : 0;
      }
      __syncthreads();
      if (kind >= 0) {
        int off = s_kbase[parity][kind];
        for (int w = 0; w < wib; w++) off += s_wcnt[parity][kind][w];
        buf.klist[kind][off + __popc(mine & ((1u << lane) - 1u))] = q;
      }
      // (the next trip writes the other half of s_wcnt / s_kbase; its first barrier orders it after these reads)
    } else {
      // one atomic per warp and kind present in the warp
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const unsigned m = __ballot_sync(0xffffffffu, kind == k);
        if (m) {
          int pos0 = 0;
          if (lane == (unsigned)(__ffs(m) - 1)) pos0 = atomicAdd(buf.counts + 4 + k, __popc(m));
          pos0 = __shfl_sync(0xffffffffu, pos0, __ffs(m) - 1);
          if (kind == k) buf.klist[k][pos0 + __popc(m & ((1u << lane) - 1u))] = q;
        }
      }
    }
    st_use = st_use + 1 == kStages ? 0 : st_use + 1;
    st_fill = st_fill + 1 == kStages ? 0 : st_fill + 1;
  }
}

// Material sampling specialised by kind: the direction, its density (finite, Dirac
// coefficient) and BSDF * |cos| / density.
template <int KIND>
__device__ __forceinline__ V3f kind_sample_source(const DeviceScene &sc, const MatAt &m, Rng &g, V3f nrm, V3f dest,
                                                  int &tag) {
  if (KIND == M3D_MAT_LAMBERT) {
    tag = 0;
    return lambert_sample(g, nrm);
  }
  if (KIND == M3D_MAT_JOINED) return mat_sample_source(sc, m, g, nrm, dest, tag);
  return simple_sample_source(sc.materials[m.index], m.diffuse, g, nrm, dest, tag);
}
template <int KIND>
__device__ __forceinline__ Density kind_source_density(const DeviceScene &sc, const MatAt &m, V3f nrm, V3f source,
                                                       V3f dest, int tag) {
  if (KIND == M3D_MAT_LAMBERT) {
    Density r;
    r.fin = lambert_density(nrm, source);
    r.del = 0.f;
    return r;
  }
  if (KIND == M3D_MAT_JOINED) return mat_source_density(sc, m, nrm, source, dest, tag);
  return simple_source_density(sc.materials[m.index], m.diffuse, nrm, source, dest, tag & 3);
}
template <int KIND>
__device__ __forceinline__ V3f kind_bsdf(const DeviceScene &sc, const MatAt &m, V3f nrm, V3f source, V3f dest) {
  if (KIND == M3D_MAT_LAMBERT) {  // material.go:125-134
    if (dot(dest, nrm) < 0.f || dot(source, nrm) > 0.f) return v3f(0.f, 0.f, 0.f);
    return m.diffuse * 4.f;
  }
  if (KIND == M3D_MAT_REFRACT) return v3f(0.f, 0.f, 0.f);  // Dirac lobes only
  if (KIND == M3D_MAT_JOINED) return mat_bsdf(sc, m, nrm, source, dest);
  return simple_bsdf(sc.materials[m.index], m.diffuse, nrm, source, dest);
}
template <int KIND>
__device__ __forceinline__ V3f kind_bsdf_delta(const DeviceScene &sc, const MatAt &m, V3f nrm, V3f source, V3f dest,
                                               int tag) {
  if (KIND == M3D_MAT_LAMBERT || KIND == M3D_MAT_PHONG) return v3f(0.f, 0.f, 0.f);
  return mat_bsdf_delta(sc, m, nrm, source, dest, tag);
}

// Stage 2 of a bounce, one launch per material kind present in the scene: sampleNextSource /
// sourceDensity with focus points (raytrace.go:183-215), throughput update and cutoff
// (raytrace.go:139-142, 170-180), ballot/popc compaction of the survivors into the next queue.
template <int KIND>
__global__ void __launch_bounds__(kShadeBlock, M3D_SAMPLE_MINB)
path_sample_kernel(DeviceScene sc, DevicePathParams pp, PathBatch b, PathBuffers buf, int cur, int depth) {
  const int n = buf.counts[4 + KIND];
  const unsigned lane = threadIdx.x & 31u;
  const int nxt = cur ^ 1;
  const int32_t *__restrict__ list = buf.klist[KIND];

  for (int bbase = (int)blockIdx.x * kShadeBlock; bbase < n; bbase += (int)gridDim.x * kShadeBlock) {
    const int li = bbase + (int)threadIdx.x;
    bool alive = false;
    int slot = 0, surf = -1;
    V3f point = v3f(0.f, 0.f, 0.f), next_dir = v3f(0.f, 0.f, 1.f);
    float4 thr = make_float4(0.f, 0.f, 0.f, 0.f);
    if (li < n) {
      const int q = list[li];
      const float4 ra = buf.hrA[q], rb = buf.hrB[q], rc = buf.hrC[q];
      point = v3f(ra.x, ra.y, ra.z);
      surf = __float_as_int(ra.w);
      const V3f nrm = v3f(rb.x, rb.y, rb.z), dest = v3f(rc.x, rc.y, rc.z);
      slot = __float_as_int(rc.w);
      const MatAt m = material_at_index(sc, __float_as_int(rb.w), point);
      thr = buf.thr[slot];
      const V3f tv = v3f(thr.x, thr.y, thr.z);
      Rng g;
      g.init(pp.seed, (uint32_t)batch_pixel(b, slot % b.nP), b.sample0 + (uint32_t)(slot / b.nP), (uint32_t)depth);
      // sampleNextSource (raytrace.go:183-199)
      V3f source;
      int tag = 0;
      int chosen = -1;
      if (pp.num_focus > 0) {
        float u = g.f32();
        for (int i = 0; i < pp.num_focus; i++) {
          u -= pp.focus[i].prob;
          if (u < 0.f) {
            chosen = i;
            break;
          }
        }
      }
      bool from_focus = false;
      if (chosen >= 0) {
        V3f fdir;
        float min_cos;
        if (focus_active(pp.focus[chosen], m.index, point, fdir, min_cos)) {
          from_focus = true;
          source = pp.focus[chosen].kind == M3D_FOCUS_PHONG ? sample_around_direction(g, pp.focus[chosen].alpha, fdir)
                                                            : sample_around_uniform(g, min_cos, fdir);
        }
      }
      if (!from_focus) source = kind_sample_source<KIND>(sc, m, g, nrm, dest, tag);
      // sourceDensity (raytrace.go:201-215): mixture of focus densities and the material's
      const Density md = kind_source_density<KIND>(sc, m, nrm, source, dest, tag);
      float dens_fin = md.fin, dens_del = md.del;
      if (pp.num_focus > 0) {
        float mat_prob = 1.f, fin = 0.f;
        for (int i = 0; i < pp.num_focus; i++) {
          V3f fdir;
          float min_cos;
          if (focus_active(pp.focus[i], m.index, point, fdir, min_cos)) {
            fin += pp.focus[i].prob * focus_density(pp.focus[i], fdir, min_cos, source);
            mat_prob -= pp.focus[i].prob;
          }
        }
        dens_fin = fin + mat_prob * md.fin;
        dens_del = mat_prob * md.del;
      }
      // weight = |cos| / density; mask = BSDF * weight (raytrace.go:170-176).  A direction drawn
      // from a Dirac lobe carries bsdf and density proportional to 2/cosineEpsilon: their ratio
      // is taken analytically (the finite parts are 1e-8 relative).
      const float cosv = fabsf(dot(source, nrm));
      V3f mask;
      if (dens_del > 0.f)
        mask = kind_bsdf_delta<KIND>(sc, m, nrm, source, dest, tag) * (cosv / dens_del);
      else
        mask = kind_bsdf<KIND>(sc, m, nrm, source, dest) * (dens_fin > 0.f ? cosv / dens_fin : 0.f);
      const V3f nt = tv * mask;
      const float mean = (nt.x + nt.y + nt.z) * (1.f / 3.f);
      // recurse() entry test (raytrace.go:139-142); a zero / non-finite throughput can never
      // contribute again
      if (mean >= pp.cutoff && mean > 0.f && mean < INFINITY) {
        alive = true;
        thr = make_float4(nt.x, nt.y, nt.z, 0.f);
        next_dir = source * -1.f;
      }
    }
    // compaction: surviving lanes take consecutive positions of the next queue
    const unsigned live = __ballot_sync(0xffffffffu, alive);
    // one atomic per warp (aggregating the warps of a block behind two barriers measured slower here:
    // C3 sample stage 16.1 -> 16.5 ms, C4 9.05 -> 9.37 ms)
    if (live) {
      int pos0 = 0;
      if (lane == (unsigned)(__ffs(live) - 1)) pos0 = atomicAdd(buf.counts + nxt, __popc(live));
      pos0 = __shfl_sync(0xffffffffu, pos0, __ffs(live) - 1);
      if (alive) {
        const int pos = pos0 + __popc(live & ((1u << lane) - 1u));
        const float b_tmin = pp.eps > 0.f ? pp.eps * rsqrtf(dot(next_dir, next_dir)) : 0.f;
        buf.org[nxt][pos] = make_float4(point.x, point.y, point.z, b_tmin);
        buf.dir[nxt][pos] = make_float4(next_dir.x, next_dir.y, next_dir.z, INFINITY);
        buf.skip[nxt][pos] = surf;
        buf.queue[nxt][pos] = slot;
        buf.thr[slot] = thr;
      }
    }
  }
}

template <int SB>
__global__ void __launch_bounds__(256)
path_shadow_resolve_kernel(DeviceScene sc, DevicePathParams pp, PathBuffers buf, int cur) {
  const int n = buf.counts[cur];
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q == 0) atomicAdd(buf.ray_total, (unsigned long long)n * pp.num_lights);
  if (q >= n) return;
  float3 add = make_float3(0.f, 0.f, 0.f);
  int slot = 0;
  for (int l = 0; l < pp.num_lights; l++) {
    const size_t si = (size_t)q * pp.num_lights + l;
    const float4 pay = buf.spay[si];
    slot = __float_as_int(pay.w);
    const float4 d = buf.sdir[si];
    if (d.w < 0.f) continue;
    const SceneHit h = resolve_scene_hit<SB>(sc, buf.sorg[si], d, buf.sraw[si], buf.sskip[si], false);
    if (h.obj >= 0) continue;  // occluded (tmax == 1 bounds the query)
    add.x += pay.x;
    add.y += pay.y;
    add.z += pay.z;
  }
  if (add.x != 0.f || add.y != 0.f || add.z != 0.f) {
    float4 a = buf.accum[slot];
    a.x += add.x;
    a.y += add.y;
    a.z += add.z;
    buf.accum[slot] = a;
  }
}

// colorSum (+ squares) per pixel over the S samples of a batch (ray_renderer.go:125-127,150).
//   FLUSH_ADD     dst[pixel] += sum                 (one GPU owns dst)
//   FLUSH_CARRY   carry[p]   += sum                 (p = position in the batch: the local partial
//                                                    sums of a pixel range that takes several batches)
//   FLUSH_RED     dst[pixel] (+)= carry[p] + sum    with red.add: dst is the frame accumulator that
//                 several GPUs flush into at once, usually rank 0's memory mapped over NVLink
//                 (peer access / CUDA IPC), so the cross-GPU reduce of the per-pixel sums
//                  happens inside this kernel, tile by tile, instead of in a
//                 collective afterwards.  System-scope reductions: the adds of different GPUs meet
//                 in the owning GPU's L2.
enum { FLUSH_ADD = 0, FLUSH_CARRY = 1, FLUSH_RED = 2 };

__device__ __forceinline__ void red_add_sys(float *addr, float v) {
  asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add_sys_v4(float *addr, float a, float b, float c, float d) {
  asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(256)
path_flush_kernel(PathBatch b, const float4 *__restrict__ accum, float *__restrict__ rgb_sum,
                  float *__restrict__ rgb_sumsq, float *__restrict__ carry, float *__restrict__ carry_sq) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= b.nP) return;
  float sx = 0.f, sy = 0.f, sz = 0.f, qx = 0.f, qy = 0.f, qz = 0.f;
  for (int s = 0; s < b.S; s++) {
    const float4 a = __ldcs(accum + (size_t)s * b.nP + p);
    sx += a.x;
    sy += a.y;
    sz += a.z;
    qx += a.x * a.x;
    qy += a.y * a.y;
    qz += a.z * a.z;
  }
  if (MODE == FLUSH_CARRY) {
    const size_t c = (size_t)p * 3;
    carry[c] += sx;
    carry[c + 1] += sy;
    carry[c + 2] += sz;
    if (carry_sq) {
      carry_sq[c] += qx;
      carry_sq[c + 1] += qy;
      carry_sq[c + 2] += qz;
    }
    return;
  }
  const size_t o = (size_t)batch_pixel(b, p) * 3;
  if (MODE == FLUSH_ADD) {
    rgb_sum[o] += sx;
    rgb_sum[o + 1] += sy;
    rgb_sum[o + 2] += sz;
    if (rgb_sumsq) {
      rgb_sumsq[o] += qx;
      rgb_sumsq[o + 1] += qy;
      rgb_sumsq[o + 2] += qz;
    }
  } else {
    if (carry) {
      const size_t c = (size_t)p * 3;
      sx += carry[c];
      sy += carry[c + 1];
      sz += carry[c + 2];
      if (carry_sq) {
        qx += carry_sq[c];
        qy += carry_sq[c + 1];
        qz += carry_sq[c + 2];
      }
    }
    red_add_sys(rgb_sum + o, sx);
    red_add_sys(rgb_sum + o + 1, sy);
    red_add_sys(rgb_sum + o + 2, sz);
    if (rgb_sumsq) {
      red_add_sys(rgb_sumsq + o, qx);
      red_add_sys(rgb_sumsq + o + 1, qy);
      red_add_sys(rgb_sumsq + o + 2, qz);
    }
  }
}

// FLUSH_RED for a contiguous pixel range whose first pixel is a multiple of four: one thread sums
// four consecutive pixels and pushes their 12 floats as three 16-byte vector reductions (a warp
// covers 1.5 KB of the remote accumulator per instruction triple instead of 384 scattered bytes).
__global__ void __launch_bounds__(256)
path_flush_red4_kernel(PathBatch b, const float4 *__restrict__ accum, float *__restrict__ rgb_sum,
                       float *__restrict__ rgb_sumsq, const float *__restrict__ carry,
                       const float *__restrict__ carry_sq) {
  const int p4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (p4 >= b.nP) return;
  float su[12], sq[12];
#pragma unroll
  for (int k = 0; k < 12; k++) su[k] = sq[k] = 0.f;
  const int np = b.nP - p4 < 4 ? b.nP - p4 : 4;
  for (int s = 0; s < b.S; s++) {
    const float4 *row = accum + (size_t)s * b.nP + p4;
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (j < np) {
        const float4 a = __ldcs(row + j);
        su[3 * j] += a.x;
        su[3 * j + 1] += a.y;
        su[3 * j + 2] += a.z;
        sq[3 * j] += a.x * a.x;
        sq[3 * j + 1] += a.y * a.y;
        sq[3 * j + 2] += a.z * a.z;
      }
  }
  if (carry) {
    for (int k = 0; k < 3 * np; k++) {
      su[k] += carry[(size_t)p4 * 3 + k];
      if (carry_sq) sq[k] += carry_sq[(size_t)p4 * 3 + k];
    }
  }
  const size_t o = (size_t)(b.pix0 + p4) * 3;  // multiple of 12 floats: 16-byte aligned
  if (np == 4) {
#pragma unroll
    for (int v = 0; v < 3; v++) red_add_sys_v4(rgb_sum + o + 4 * v, su[4 * v], su[4 * v + 1], su[4 * v + 2], su[4 * v + 3]);
    if (rgb_sumsq) {
#pragma unroll
      for (int v = 0; v < 3; v++)
        red_add_sys_v4(rgb_sumsq + o + 4 * v, sq[4 * v], sq[4 * v + 1], sq[4 * v + 2], sq[4 * v + 3]);
    }
  } else {
    for (int k = 0; k < 3 * np; k++) {
      red_add_sys(rgb_sum + o + k, su[k]);
      if (rgb_sumsq) red_add_sys(rgb_sumsq + o + k, sq[k]);
    }
  }
}

}  // namespace

void launch_path_raygen(const DeviceCamera &cam, const DevicePathParams &pp, const PathBatch &b,
                        const PathBuffers &buf, cudaStream_t stream) {
  const int64_t n = (int64_t)b.nP * b.S;
  if (n <= 0) return;
  path_raygen_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(cam, pp, b, buf);
}

template <class K>
static int shade_grid(K kernel, int64_t n) {
  int x = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&x, kernel, kShadeBlock, 0);
  int64_t grid = (int64_t)device_sm_count() * (x > 0 ? x : 1);
  const int64_t want = (n + kShadeBlock - 1) / kShadeBlock;
  if (grid > want) grid = want;
  return (int)(grid < 1 ? 1 : grid);
}

void launch_path_resolve(const DeviceScene &sc, const DevicePathParams &pp, const DevicePointLight *lights,
                         const PathBatch &b, const PathBuffers &buf, int cur, int depth, cudaStream_t stream) {
  // resident-grid sizes, computed once (thread-safe static initialisation: renders may run on
  // several host threads, one per device)
  static const int grid_full[5] = {shade_grid(path_resolve_kernel<false, 1>, (int64_t)1 << 40),
                                   shade_grid(path_resolve_kernel<true, 1>, (int64_t)1 << 40),
                                   shade_grid(path_resolve_kernel<false, 2>, (int64_t)1 << 40),
                                   shade_grid(path_resolve_kernel<true, 2>, (int64_t)1 << 40),
                                   shade_grid(path_resolve_kernel<false, 3>, (int64_t)1 << 40)};
  const int64_t n = (int64_t)b.nP * b.S;
  int li = (pp.num_lights > 0 ? 1 : 0) + (sc.shape_bvh.nodes ? 2 : 0);
  if (li == 0 && sc.spheres_only) li = 4;
  const int64_t want = (n + kShadeBlock - 1) / kShadeBlock;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(grid_full[li], want));
  switch (li) {
    case 0: path_resolve_kernel<false, 1><<<grid, kShadeBlock, 0, stream>>>(sc, pp, lights, b, buf, cur, depth); break;
    case 1: path_resolve_kernel<true, 1><<<grid, kShadeBlock, 0, stream>>>(sc, pp, lights, b, buf, cur, depth); break;
    case 2: path_resolve_kernel<false, 2><<<grid, kShadeBlock, 0, stream>>>(sc, pp, lights, b, buf, cur, depth); break;
    case 4: path_resolve_kernel<false, 3><<<grid, kShadeBlock, 0, stream>>>(sc, pp, lights, b, buf, cur, depth); break;
    default: path_resolve_kernel<true, 2><<<grid, kShadeBlock, 0, stream>>>(sc, pp, lights, b, buf, cur, depth); break;
  }
}

void launch_path_sample(int kind, const DeviceScene &sc, const DevicePathParams &pp, const PathBatch &b,
                        const PathBuffers &buf, int cur, int depth, cudaStream_t stream) {
  const int64_t n = (int64_t)b.nP * b.S;
  static const int grids[4] = {
      shade_grid(path_sample_kernel<0>, (int64_t)1 << 40), shade_grid(path_sample_kernel<1>, (int64_t)1 << 40),
      shade_grid(path_sample_kernel<2>, (int64_t)1 << 40), shade_grid(path_sample_kernel<3>, (int64_t)1 << 40)};
  const int64_t want = (n + kShadeBlock - 1) / kShadeBlock;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(grids[kind], want));
  switch (kind) {
    case 0: path_sample_kernel<0><<<grid, kShadeBlock, 0, stream>>>(sc, pp, b, buf, cur, depth); break;
    case 1: path_sample_kernel<1><<<grid, kShadeBlock, 0, stream>>>(sc, pp, b, buf, cur, depth); break;
    case 2: path_sample_kernel<2><<<grid, kShadeBlock, 0, stream>>>(sc, pp, b, buf, cur, depth); break;
    default: path_sample_kernel<3><<<grid, kShadeBlock, 0, stream>>>(sc, pp, b, buf, cur, depth); break;
  }
}

void launch_path_shadow_resolve(const DeviceScene &sc, const DevicePathParams &pp, const PathBuffers &buf,
                                int cur, cudaStream_t stream) {
  if (buf.cap <= 0) return;
  if (sc.shape_bvh.nodes)
    path_shadow_resolve_kernel<2><<<(unsigned)((buf.cap + 255) / 256), 256, 0, stream>>>(sc, pp, buf, cur);
  else
    path_shadow_resolve_kernel<1><<<(unsigned)((buf.cap + 255) / 256), 256, 0, stream>>>(sc, pp, buf, cur);
}

void launch_path_flush(const PathBatch &b, const float4 *accum, float *rgb_sum, float *rgb_sumsq,
                       cudaStream_t stream) {
  if (b.nP <= 0) return;
  path_flush_kernel<FLUSH_ADD><<<(unsigned)((b.nP + 255) / 256), 256, 0, stream>>>(b, accum, rgb_sum, rgb_sumsq,
                                                                                    nullptr, nullptr);
}

void launch_path_flush_carry(const PathBatch &b, const float4 *accum, float *carry, float *carry_sq,
                             cudaStream_t stream) {
  if (b.nP <= 0) return;
  path_flush_kernel<FLUSH_CARRY><<<(unsigned)((b.nP + 255) / 256), 256, 0, stream>>>(b, accum, nullptr, nullptr,
                                                                                      carry, carry_sq);
}

void launch_path_flush_red(const PathBatch &b, const float4 *accum, float *rgb_sum, float *rgb_sumsq,
                           float *carry, float *carry_sq, cudaStream_t stream) {
  if (b.nP <= 0) return;
  const bool aligned = !b.pixels && (b.pix0 & 3) == 0 && ((uintptr_t)rgb_sum & 15) == 0 &&
                       ((uintptr_t)rgb_sumsq & 15) == 0;
  if (aligned) {
    const int threads = (b.nP + 3) / 4;
    path_flush_red4_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(b, accum, rgb_sum, rgb_sumsq,
                                                                                   carry, carry_sq);
  } else {
    path_flush_kernel<FLUSH_RED><<<(unsigned)((b.nP + 255) / 256), 256, 0, stream>>>(b, accum, rgb_sum, rgb_sumsq,
                                                                                      carry, carry_sq);
  }
}

}  // namespace m3d
