// Device restatement (float32) of render3d's materials for shading kernels:
//   LambertMaterial  render3d/material.go:119-167
//   PhongMaterial    render3d/material.go:172-269 (+ sampleAroundDirection :274-335)
//   RefractMaterial  render3d/material.go:343-479
//   JoinedMaterial   render3d/material.go:554-631
//   showcase procedural variants (checker floor room.go:61-75, z-gradient vase models.go:79-97)
//
// Delta lobes.  RefractMaterial encodes Dirac lobes as windows of half-width 1e-8 with
// magnitude 2/1e-8 (material.go:401-423,441-462); 1-1e-8 is not representable in float32.
// Here a lobe is carried symbolically: BSDF = finite + delta_bsdf * (2/eps) and density =
// finite + delta_density * (2/eps) for the ONE direction that the delta sampler itself
// produced (identified by a tag, never by comparing directions); for every other direction
// the delta parts are zero, exactly as in the reference up to probability-zero events.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/m3d.h"
#include "scene.h"

namespace m3d {

struct V3f {
  float x, y, z;
};
__device__ __forceinline__ V3f v3f(float x, float y, float z) {
  V3f r;
  r.x = x;
  r.y = y;
  r.z = z;
  return r;
}
__device__ __forceinline__ V3f v3f(const float *p) { return v3f(p[0], p[1], p[2]); }
__device__ __forceinline__ V3f operator+(V3f a, V3f b) { return v3f(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3f operator-(V3f a, V3f b) { return v3f(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3f operator*(V3f a, float s) { return v3f(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3f operator*(V3f a, V3f b) { return v3f(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float dot(V3f a, V3f b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3f cross(V3f a, V3f b) {
  return v3f(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float norm(V3f a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ V3f normalize(V3f a) { return a * (1.0f / norm(a)); }
__device__ __forceinline__ bool is_zero(V3f a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f; }
__device__ __forceinline__ float sum3(V3f a) { return a.x + a.y + a.z; }

// coords.go:431-434 (n is unit already on every call site here)
__device__ __forceinline__ V3f reflect_about(V3f n, V3f c1) { return (c1 + n * (-2.f * dot(n, c1))) * -1.f; }

// Reciprocal square root / square root through the SFU alone (MUFU.RSQ / MUFU.SQRT, ~2 ulp): for the
// Monte-Carlo sampling formulas only, whose results feed an estimator with 1e-2 .. 1e-3 noise; the
// deterministic shading paths keep normalize() / sqrtf().
#ifndef M3D_FAST_SQRT
#define M3D_FAST_SQRT 1
#endif
#ifndef M3D_FAST_NORMALIZE
#define M3D_FAST_NORMALIZE 1
#endif
#ifndef M3D_FAST_SINCOS
#define M3D_FAST_SINCOS 1
#endif
__device__ __forceinline__ float sqrt_fast(float x) {
#if M3D_FAST_SQRT
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return sqrtf(x);
#endif
}
__device__ __forceinline__ V3f normalize_fast(V3f a) {
#if M3D_FAST_NORMALIZE
  return a * rsqrtf(dot(a, a));
#else
  return a * (1.0f / sqrtf(dot(a, a)));
#endif
}

// coords.go:388-421.  The reference divides the first basis vector by the largest component before
// normalising it; the direction is the same without the division.  b1 is orthogonal to c by
// construction, so b1 x c only needs the (near-unit) length of c divided out.
// (Out-of-line helpers take and return VALUES: a reference parameter would force the caller's object
// into local memory for the call, and with it every other access to that object.)
struct Basis {
  V3f x, z;
};
static __device__ __noinline__ Basis ortho_basis(V3f c) {
  const float ax = fabsf(c.x), ay = fabsf(c.y), az = fabsf(c.z);
  V3f b1 = v3f(0.f, 0.f, 0.f);
  if (ax > ay && ax > az) {
    b1.x = c.y;
    b1.y = -c.x;
  } else {
    b1.y = c.z;
    b1.z = -c.y;
  }
  b1 = normalize_fast(b1);
  const V3f b2 = v3f(b1.y * c.z - b1.z * c.y, b1.z * c.x - b1.x * c.z, b1.x * c.y - b1.y * c.x);
  Basis r;
  r.x = b1;
  r.z = normalize_fast(b2);
  return r;
}

// x^y for x in [0, 1], y >= 0 through the SFU (lg2.approx / ex2.approx): a dozen instructions
// instead of powf's ~150, relative error ~ y * 2^-22 (1e-4 at alpha = 400) -- far below the
// Monte-Carlo noise these sampling formulas feed, and the same function is used by the sampler
// and by its density so the estimator stays consistent.
__device__ __forceinline__ float pow_unit(float x, float y) {
  if (y == 0.f) return 1.f;  // pow(0, 0) == 1 like math.Pow
  return exp2f(y * __log2f(x));
}
// sin and cos of 2*pi*u without libm's large-argument slow path
// (u in [0, 1): the angle is folded to [-pi, pi) where MUFU.SIN / MUFU.COS are accurate to ~5e-7
// absolute; sin(2 pi u) = -sin(2 pi (u - 1/2)), likewise for the cosine)
__device__ __forceinline__ void sincos_2pi(float u, float *s, float *c) {
#if M3D_FAST_SINCOS
  const float a = 6.283185307179586f * (u - 0.5f);
  *s = -__sinf(a);
  *c = -__cosf(a);
#else
  sincospif(2.f * u, s, c);
#endif
}

constexpr float kCosEps = 1e-8f;  // cosineEpsilon material.go:10

// Resolved material at a hit: index + procedural diffuse colour.
struct MatAt {
  int32_t index;
  V3f diffuse;  // of the top-level material (procedural variants resolved)
};

// material `index` at `point` (procedural variants resolved)
__device__ __forceinline__ MatAt material_at_index(const DeviceScene &sc, int32_t index, V3f point);
__device__ __forceinline__ MatAt material_at(const DeviceScene &sc, int32_t object, V3f point) {
  return material_at_index(sc, sc.objects[object].material, point);
}
__device__ __forceinline__ MatAt material_at_index(const DeviceScene &sc, int32_t index, V3f point) {
  MatAt m;
  m.index = index;
  const DeviceMaterial &d = sc.materials[m.index];
  m.diffuse = v3f(d.diffuse);
  if (d.flags & M3D_MAT_CHECKER) {
    // showcase FloorObject (room.go:66-70)
    const bool same = (int)fmodf(point.x + 300.f, 2.f) == (int)fmodf(point.y + 301.f, 2.f);
    m.diffuse = same ? v3f(d.diffuse2) : v3f(d.diffuse);
  } else if (d.flags & M3D_MAT_Z_GRADIENT) {
    // showcase VaseObject (models.go:86-90)
    const float frac = point.z / d.proc_param;
    m.diffuse = v3f(d.diffuse) * frac + v3f(d.diffuse2) * (1.f - frac);
  }
  return m;
}

__device__ __forceinline__ float maximum_cosine(float c1, float c2) {
  return fmaxf(fmaxf(fabsf(c1), fabsf(c2)), kCosEps);
}

// Finite part of the BSDF of a non-joined material (delta lobes excluded, see header).
static __device__ __noinline__ V3f simple_bsdf(const DeviceMaterial &d, V3f diffuse, V3f n, V3f src, V3f dst) {
  if (d.kind == M3D_MAT_LAMBERT) {  // material.go:125-134
    if (dot(dst, n) < 0.f || dot(src, n) > 0.f) return v3f(0.f, 0.f, 0.f);
    return diffuse * 4.f;
  }
  if (d.kind == M3D_MAT_PHONG) {  // material.go:187-216
    const float dest_dot = dot(dst, n), source_dot = -dot(src, n);
    if (dest_dot < 0.f || source_dot < 0.f) return v3f(0.f, 0.f, 0.f);
    V3f color = v3f(0.f, 0.f, 0.f);
    if (!is_zero(diffuse)) color = diffuse * 4.f;
    const V3f reflection = reflect_about(n, src) * -1.f;
    const float ref_dot = dot(reflection, dst);
    if (ref_dot < 0.f) return color;
    float intensity = pow_unit(ref_dot, d.alpha) * (1.f + d.alpha);
    if (!(d.flags & M3D_MAT_NO_FLUX_CORRECTION)) intensity /= maximum_cosine(source_dot, dest_dot);
    return color + v3f(d.specular) * (2.f * intensity);
  }
  return v3f(0.f, 0.f, 0.f);  // RefractMaterial: delta lobes only
}

__device__ __forceinline__ V3f mat_bsdf(const DeviceScene &sc, const MatAt &m, V3f n, V3f src, V3f dst) {
  const DeviceMaterial &d = sc.materials[m.index];
  if (d.kind != M3D_MAT_JOINED) return simple_bsdf(d, m.diffuse, n, src, dst);
  V3f r = v3f(0.f, 0.f, 0.f);
  for (int i = 0; i < d.num_sub; i++) {
    const DeviceMaterial &s = sc.materials[d.sub[i]];
    r = r + simple_bsdf(s, v3f(s.diffuse), n, src, dst);
  }
  return r;
}

__device__ __forceinline__ V3f mat_emission(const DeviceScene &sc, const MatAt &m) {
  const DeviceMaterial &d = sc.materials[m.index];
  if (d.kind == M3D_MAT_REFRACT) return v3f(0.f, 0.f, 0.f);
  if (d.kind != M3D_MAT_JOINED) return v3f(d.emission);
  V3f r = v3f(0.f, 0.f, 0.f);
  for (int i = 0; i < d.num_sub; i++) {
    const DeviceMaterial &s = sc.materials[d.sub[i]];
    if (s.kind != M3D_MAT_REFRACT) r = r + v3f(s.emission);
  }
  return r;
}
__device__ __forceinline__ V3f mat_ambient(const DeviceScene &sc, const MatAt &m) {
  const DeviceMaterial &d = sc.materials[m.index];
  if (d.kind == M3D_MAT_REFRACT) return v3f(0.f, 0.f, 0.f);
  if (d.kind != M3D_MAT_JOINED) return v3f(d.ambient);
  V3f r = v3f(0.f, 0.f, 0.f);
  for (int i = 0; i < d.num_sub; i++) {
    const DeviceMaterial &s = sc.materials[d.sub[i]];
    if (s.kind != M3D_MAT_REFRACT) r = r + v3f(s.ambient);
  }
  return r;
}

// PointLight.ShadeCollision (light.go:71-92)
__device__ __forceinline__ V3f shade_collision(const DevicePointLight &l, V3f n, V3f point_to_light) {
  const float dist = norm(point_to_light);
  V3f color = v3f(l.color);
  if (l.quad_dropoff) color = color * (1.f / (dist * dist));
  const float density = 0.25f * fmaxf(0.f, dot(n, point_to_light * (1.f / dist)));
  return color * density;
}

// ---- random numbers: Philox4x32-10, counter based ------------------------------------------
// The reference draws from Go's math/rand per goroutine (concurrency.go:33-35), so parity is
// statistical.  Here every (pixel, sample) owns a Philox stream keyed by the seed: the image
// is independent of how samples are partitioned over GPUs or batches.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// One Philox block, out of line (ten rounds, ~80 instructions per call site otherwise); values in, values out, so
// that the generator's state stays in registers.
static __device__ __noinline__ uint4 philox_block(uint32_t pixel, uint32_t sample, uint32_t block, uint32_t domain,
                                                  uint32_t k0, uint32_t k1) {
  return philox4x32_10(make_uint4(pixel, sample, block, domain), make_uint2(k0, k1));
}

struct Rng {
  uint2 key;
  uint32_t pixel, sample, block, domain;
  uint4 buf;
  int have;
  __device__ __forceinline__ void init(uint64_t seed, uint32_t pixel_, uint32_t sample_, uint32_t domain_) {
    key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    pixel = pixel_;
    sample = sample_;
    domain = domain_;
    block = 0;
    have = 0;
  }
  __device__ __forceinline__ void refill() {
    buf = philox_block(pixel, sample, block++, domain, key.x, key.y);
    have = 4;
  }
  __device__ __forceinline__ uint32_t bits() {
    if (have == 0) refill();
    const uint32_t r = have == 4 ? buf.x : (have == 3 ? buf.y : (have == 2 ? buf.z : buf.w));
    have--;
    return r;
  }
  // uniform in [0, 1) with 24 bits
  __device__ __forceinline__ float f32() { return (float)(bits() >> 8) * 5.9604644775390625e-8f; }
};

// ---- sampling (material.go) ---------------------------------------------------------------
constexpr float kTwoPi = 6.283185307179586f;


// Tag of the direction a sampler produced: which Dirac lobe(s) of which sub-material.
constexpr int kLobeRefract = 1, kLobeReflect = 2;  // bits 0..1; sub-material index in bits 2..3

// material.go:136-151 (u, v: the two uniforms the reference draws)
static __device__ __noinline__ V3f lambert_direction(float u, float v, V3f normal) {
  const float cos_lat = sqrt_fast(u), sin_lat = sqrt_fast(1.f - u);
  float sl, cl;
  sincos_2pi(v, &sl, &cl);
  const Basis b = ortho_basis(normal);
  const V3f lon_point = b.x * cl + b.z * sl;
  return normal * -cos_lat + lon_point * sin_lat;
}
__device__ __forceinline__ V3f lambert_sample(Rng &g, V3f normal) {
  const float u = g.f32(), v = g.f32();
  return lambert_direction(u, v, normal);
}
// material.go:153-159
__device__ __forceinline__ float lambert_density(V3f normal, V3f source) {
  const float nd = -dot(normal, source);
  return nd < 0.f ? 0.f : 4.f * nd;
}
// material.go:274-323
static __device__ __noinline__ V3f direction_around(float u, float v, float alpha, V3f direction) {
  const Basis b = ortho_basis(direction);
  float sl, cl;
  sincos_2pi(u, &sl, &cl);
  const float cos_lat = pow_unit(v, 1.f / (alpha + 1.f));
  const float sin_lat = sqrt_fast(fmaxf(0.f, 1.f - cos_lat * cos_lat));
  const V3f lon_point = b.x * cl + b.z * sl;
  return direction * cos_lat + lon_point * sin_lat;
}
__device__ __forceinline__ V3f sample_around_direction(Rng &g, float alpha, V3f direction) {
  const float u = g.f32(), v = g.f32();
  return direction_around(u, v, alpha, direction);
}
// material.go:328-335: 2(a+1) / v^(1/(a+1) - 1) with v = d^(a+1), i.e. 2(a+1) d^a
__device__ __forceinline__ float density_around_direction(float alpha, V3f direction, V3f sample) {
  const float d = dot(direction, sample);
  if (d < 0.f) return 0.f;
  return 2.f * (alpha + 1.f) * pow_unit(d, alpha);
}

// material.go:360-378; tir: total internal reflection (the result is the mirror direction)
__device__ __forceinline__ V3f refract_dir(float ior, V3f normal, V3f source, bool &tir) {
  V3f sine_part = source - normal * dot(normal, source);  // ProjectOut (normal is unit)
  float sine_scale = ior;
  V3f cosine_part = normal;
  if (dot(normal, source) < 0.f) {
    sine_scale = 1.f / sine_scale;
    cosine_part = cosine_part * -1.f;
  }
  sine_part = sine_part * sine_scale;
  const float sine_norm = norm(sine_part);
  tir = sine_norm > 1.f;
  if (tir) return reflect_about(normal, source) * -1.f;
  return sine_part + cosine_part * sqrtf(1.f - sine_norm * sine_norm);
}
__device__ __forceinline__ V3f refract_inverse(float ior, V3f normal, V3f dest, bool &tir) {
  return refract_dir(ior, normal, dest * -1.f, tir) * -1.f;
}
// material.go:384-389
__device__ __forceinline__ float reflect_amount(float ior, V3f normal, V3f source) {
  const float x = (ior - 1.f) / (ior + 1.f);
  const float r0 = x * x;
  const float c = 1.f - fabsf(dot(normal, source));
  return r0 * (1.f - r0) * (c * c * c * c * c);
}

// SampleSource of a non-joined material; lobe: Dirac lobes the direction belongs to.
// coin / u / v: the random numbers of the branch taken (Phong: coin, then the sampler's two uniforms;
// Lambert: u, v; Refract: u), drawn by the inlined wrapper below.
struct SampledDir {
  V3f dir;
  int lobe;
};
static __device__ __noinline__ SampledDir simple_source_direction(const DeviceMaterial &d, V3f diffuse, uint32_t coin,
                                                               float u, float v, V3f normal, V3f dest) {
  SampledDir r;
  r.lobe = 0;
  if (d.kind == M3D_MAT_LAMBERT) {
    r.dir = lambert_direction(u, v, normal);
    return r;
  }
  if (d.kind == M3D_MAT_PHONG) {  // material.go:230-236,251-256
    if (is_zero(diffuse) || (coin & 1u) == 0u) {
      const V3f reflection = reflect_about(normal, dest) * -1.f;
      r.dir = direction_around(u, v, d.alpha, reflection);
      return r;
    }
    r.dir = lambert_direction(u, v, normal);
    return r;
  }
  // RefractMaterial material.go:425-439
  bool tir;
  const V3f refracted = refract_inverse(d.ior, normal, dest, tir);
  if (is_zero(v3f(d.specular))) {
    r.lobe = kLobeRefract;
    r.dir = refracted;
    return r;
  }
  const float refl = reflect_amount(d.ior, normal, dest);
  if (u > refl) {
    r.lobe = tir ? (kLobeRefract | kLobeReflect) : kLobeRefract;
    r.dir = refracted;
    return r;
  }
  r.lobe = tir ? (kLobeRefract | kLobeReflect) : kLobeReflect;
  r.dir = reflect_about(normal, dest) * -1.f;
  return r;
}
__device__ __forceinline__ V3f simple_sample_source(const DeviceMaterial &d, V3f diffuse, Rng &g, V3f normal,
                                                    V3f dest, int &lobe) {
  // the numbers are drawn whether or not the branch uses them: streams are per (pixel, sample, bounce)
  // and a Philox block holds four, so nothing is saved by drawing lazily
  const uint32_t coin = d.kind == M3D_MAT_PHONG && !is_zero(diffuse) ? g.bits() : 0u;
  const float u = g.f32();
  const float v = d.kind == M3D_MAT_REFRACT ? 0.f : g.f32();
  const SampledDir r = simple_source_direction(d, diffuse, coin, u, v, normal, dest);
  lobe = r.lobe;
  return r.dir;
}

struct Density {
  float fin;  // finite part
  float del;  // coefficient of 2/cosineEpsilon
};

// SourceDensity of a non-joined material for a direction tagged `lobe`
// (material.go:153-159, 240-247, 441-462).
static __device__ __noinline__ Density simple_source_density(const DeviceMaterial &d, V3f diffuse, V3f normal,
                                                         V3f source, V3f dest, int lobe) {
  Density r;
  r.fin = 0.f;
  r.del = 0.f;
  if (d.kind == M3D_MAT_LAMBERT) {
    r.fin = lambert_density(normal, source);
  } else if (d.kind == M3D_MAT_PHONG) {
    const V3f reflection = reflect_about(normal, dest) * -1.f;
    const float pw = density_around_direction(d.alpha, reflection, source);
    r.fin = is_zero(diffuse) ? pw : 0.5f * (pw + lambert_density(normal, source));
  } else if (d.kind == M3D_MAT_REFRACT) {
    if (is_zero(v3f(d.specular))) {
      r.del = (lobe & kLobeRefract) ? 1.f : 0.f;
    } else {
      const float refl = reflect_amount(d.ior, normal, dest);
      r.del = ((lobe & kLobeRefract) ? 1.f - refl : 0.f) + ((lobe & kLobeReflect) ? refl : 0.f);
    }
  }
  return r;
}

// Dirac part of the BSDF of a non-joined material, as a coefficient of 2/cosineEpsilon
// (material.go:391-423); the finite part is simple_bsdf().
static __device__ __noinline__ V3f simple_bsdf_delta(const DeviceMaterial &d, V3f normal, V3f source, V3f dest,
                                                 int lobe) {
  if (d.kind != M3D_MAT_REFRACT || lobe == 0) return v3f(0.f, 0.f, 0.f);
  const float s_refr = 1.f / fmaxf(kCosEps, fabsf(dot(dest, normal)));
  if (is_zero(v3f(d.specular))) return (lobe & kLobeRefract) ? v3f(d.refract) * s_refr : v3f(0.f, 0.f, 0.f);
  const float ra = reflect_amount(d.ior, normal, source);
  V3f r = v3f(0.f, 0.f, 0.f);
  if (lobe & kLobeRefract) r = r + v3f(d.refract) * ((1.f - ra) * s_refr);
  if (lobe & kLobeReflect) r = r + v3f(d.specular) * (ra / maximum_cosine(dot(dest, normal), dot(source, normal)));
  return r;
}

// Material.SampleSource incl. JoinedMaterial (material.go:573-586); tag = lobe | sub << 2.
__device__ __forceinline__ V3f mat_sample_source(const DeviceScene &sc, const MatAt &m, Rng &g, V3f normal,
                                                 V3f dest, int &tag) {
  const DeviceMaterial &d = sc.materials[m.index];
  if (d.kind != M3D_MAT_JOINED) return simple_sample_source(d, m.diffuse, g, normal, dest, tag);
  float p = g.f32();
  int pick = d.num_sub - 1;
  for (int i = 0; i < d.num_sub; i++) {
    p -= d.sub_prob[i];
    if (p < 0.f) {
      pick = i;
      break;
    }
  }
  const DeviceMaterial &s = sc.materials[d.sub[pick]];
  int lobe;
  const V3f r = simple_sample_source(s, v3f(s.diffuse), g, normal, dest, lobe);
  tag = lobe | (pick << 2);
  return r;
}

__device__ __forceinline__ Density mat_source_density(const DeviceScene &sc, const MatAt &m, V3f normal,
                                                      V3f source, V3f dest, int tag) {
  const DeviceMaterial &d = sc.materials[m.index];
  if (d.kind != M3D_MAT_JOINED) return simple_source_density(d, m.diffuse, normal, source, dest, tag & 3);
  Density r;
  r.fin = 0.f;
  r.del = 0.f;
  for (int i = 0; i < d.num_sub; i++) {  // material.go:588-594
    const DeviceMaterial &s = sc.materials[d.sub[i]];
    const Density x = simple_source_density(s, v3f(s.diffuse), normal, source, dest, (tag >> 2) == i ? (tag & 3) : 0);
    r.fin += d.sub_prob[i] * x.fin;
    r.del += d.sub_prob[i] * x.del;
  }
  return r;
}

__device__ __forceinline__ V3f mat_bsdf_delta(const DeviceScene &sc, const MatAt &m, V3f normal, V3f source,
                                              V3f dest, int tag) {
  const DeviceMaterial &d = sc.materials[m.index];
  if (d.kind != M3D_MAT_JOINED) return simple_bsdf_delta(d, normal, source, dest, tag & 3);
  const int pick = tag >> 2;
  if ((tag & 3) == 0 || pick >= d.num_sub) return v3f(0.f, 0.f, 0.f);
  return simple_bsdf_delta(sc.materials[d.sub[pick]], normal, source, dest, tag & 3);
}

// ---- destination-side sampling (light paths of the bidirectional tracer) --------------------
// Generic rule (material.go:97-116): SampleDest(n, src) = -SampleSource(n, -src) and
// DestDensity(n, src, dst) = SourceDensity(n, -dst, -src); RefractMaterial flips the normal
// instead (material.go:464-471); JoinedMaterial mixes its parts (material.go:596-616).
__device__ __forceinline__ V3f simple_sample_dest(const DeviceMaterial &d, V3f diffuse, Rng &g, V3f normal,
                                                  V3f source, int &lobe) {
  if (d.kind == M3D_MAT_REFRACT) return simple_sample_source(d, diffuse, g, normal * -1.f, source, lobe);
  return simple_sample_source(d, diffuse, g, normal, source * -1.f, lobe) * -1.f;
}
__device__ __forceinline__ Density simple_dest_density(const DeviceMaterial &d, V3f diffuse, V3f normal, V3f source,
                                                       V3f dest, int lobe) {
  if (d.kind == M3D_MAT_REFRACT) return simple_source_density(d, diffuse, normal * -1.f, dest, source, lobe);
  return simple_source_density(d, diffuse, normal, dest * -1.f, source * -1.f, 0);
}
__device__ __forceinline__ V3f mat_sample_dest(const DeviceScene &sc, const MatAt &m, Rng &g, V3f normal, V3f source,
                                               int &tag) {
  const DeviceMaterial &d = sc.materials[m.index];
  if (d.kind != M3D_MAT_JOINED) return simple_sample_dest(d, m.diffuse, g, normal, source, tag);
  float p = g.f32();
  int pick = d.num_sub - 1;
  for (int i = 0; i < d.num_sub; i++) {
    p -= d.sub_prob[i];
    if (p < 0.f) {
      pick = i;
      break;
    }
  }
  const DeviceMaterial &s = sc.materials[d.sub[pick]];
  int lobe;
  const V3f r = simple_sample_dest(s, v3f(s.diffuse), g, normal, source, lobe);
  tag = lobe | (pick << 2);
  return r;
}
__device__ __forceinline__ Density mat_dest_density(const DeviceScene &sc, const MatAt &m, V3f normal, V3f source,
                                                    V3f dest, int tag) {
  const DeviceMaterial &d = sc.materials[m.index];
  if (d.kind != M3D_MAT_JOINED) return simple_dest_density(d, m.diffuse, normal, source, dest, tag & 3);
  Density r;
  r.fin = 0.f;
  r.del = 0.f;
  for (int i = 0; i < d.num_sub; i++) {
    const DeviceMaterial &s = sc.materials[d.sub[i]];
    const Density x = simple_dest_density(s, v3f(s.diffuse), normal, source, dest, (tag >> 2) == i ? (tag & 3) : 0);
    r.fin += d.sub_prob[i] * x.fin;
    r.del += d.sub_prob[i] * x.del;
  }
  return r;
}

}  // namespace m3d
