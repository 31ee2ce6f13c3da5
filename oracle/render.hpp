// ORACLE -- TEST INFRASTRUCTURE ONLY (see vec.hpp header).
//
// float64 CPU restatement of render3d's Monte-Carlo renderers:
//   rayRenderer.estimateColor / Converged      render3d/ray_renderer.go:112-173
//   RecursiveRayTracer.recurse & helpers       render3d/raytrace.go:138-229
//   PhongFocusPoint / SphereFocusPoint         render3d/focus_point.go:30-177
//   SphereAreaLight / MeshAreaLight / joined   render3d/light.go:131-161,227-314
//   BidirPathTracer                            render3d/bidir.go:101-576
// Random streams are NOT those of Go's math/rand; parity is statistical.
#pragma once
#include <atomic>

#include "scene.hpp"

namespace orc {

inline Ray bounce_ray(V3 point, V3 dir, double eps) {
  if (eps == 0) eps = 1e-8;  // raytrace.go:10,218-221
  return Ray{add(point, scale(normalize(dir), eps)), dir};
}

// ---- focus points (focus_point.go) ---------------------------------------------------
inline bool focus_applies(const m3d_focus_point &f, const MatRef &m) {
  return m.index < 64 ? ((f.material_mask >> m.index) & 1) != 0 : false;
}
// focus_point.go:155-163
inline V3 sample_around_uniform(Rng &g, double min_cos, V3 direction) {
  double lat = std::acos(1 - g.f64() * (1 - min_cos));
  double lon = g.f64() * 2 * M_PI;
  V3 xa, za;
  ortho_basis(direction, xa, za);
  V3 lon_point = add(scale(xa, std::cos(lon)), scale(za, std::sin(lon)));
  return add(scale(direction, std::cos(lat)), scale(lon_point, std::sin(lat)));
}
// focus_point.go:172-177
inline double density_around_uniform(double min_cos, V3 direction, V3 sample) {
  if (dot(direction, sample) < min_cos) return 0;
  return 2 / (1 - min_cos);
}
// focus_point.go:112-153
inline void sphere_focus_info(const m3d_focus_point &f, V3 point, double &min_cos, V3 &dir) {
  V3 direction = sub(point, v3(f.target));
  double d = norm(direction);
  if (d < f.radius) {
    min_cos = 0;
    dir = scale(direction, 1 / d);
    return;
  }
  double ratio = f.radius / d;
  min_cos = std::sqrt(1 - ratio * ratio);
  dir = scale(direction, 1 / d);
}
inline V3 focus_sample(const m3d_focus_point &f, Rng &g, const MatRef &m, V3 point, V3 normal, V3 dest) {
  if (f.kind == M3D_FOCUS_PHONG) {  // focus_point.go:46-53
    if (v3(f.target) == point || !focus_applies(f, m)) return mat_sample_source(m, m.index, g, normal, dest);
    V3 direction = normalize(sub(point, v3(f.target)));
    return sample_around_direction(g, f.alpha, direction);
  }
  // focus_point.go:87-94
  if (dist(v3(f.target), point) < f.radius || !focus_applies(f, m))
    return mat_sample_source(m, m.index, g, normal, dest);
  double min_cos;
  V3 dir;
  sphere_focus_info(f, point, min_cos, dir);
  return sample_around_uniform(g, min_cos, dir);
}
inline double focus_density(const m3d_focus_point &f, const MatRef &m, V3 point, V3 normal, V3 source, V3 dest) {
  if (f.kind == M3D_FOCUS_PHONG) {  // focus_point.go:57-64
    if (v3(f.target) == point || !focus_applies(f, m)) return mat_source_density(m, m.index, normal, source, dest);
    V3 direction = normalize(sub(point, v3(f.target)));
    return density_around_direction(f.alpha, direction, source);
  }
  // focus_point.go:98-105
  if (dist(v3(f.target), point) < f.radius || !focus_applies(f, m))
    return mat_source_density(m, m.index, normal, source, dest);
  double min_cos;
  V3 dir;
  sphere_focus_info(f, point, min_cos, dir);
  return density_around_uniform(min_cos, dir, source);
}

// ---- RecursiveRayTracer (raytrace.go:138-229) -----------------------------------------
struct PathTracer {
  const Scene &sc;
  const m3d_point_light *lights;
  int nl;
  const m3d_path_params &p;
  int64_t casts = 0;

  V3 sample_next_source(Rng &g, V3 point, V3 normal, V3 dest, const MatRef &m) {
    if (p.num_focus_points == 0) return mat_sample_source(m, m.index, g, normal, dest);
    double u = g.f64();
    for (int i = 0; i < p.num_focus_points; i++) {
      u -= p.focus[i].prob;
      if (u < 0) return focus_sample(p.focus[i], g, m, point, normal, dest);
    }
    return mat_sample_source(m, m.index, g, normal, dest);
  }
  double source_density(V3 point, V3 normal, V3 source, V3 dest, const MatRef &m) {
    if (p.num_focus_points == 0) return mat_source_density(m, m.index, normal, source, dest);
    double mat_prob = 1.0, prob = 0;
    for (int i = 0; i < p.num_focus_points; i++) {
      prob += p.focus[i].prob * focus_density(p.focus[i], m, point, normal, source, dest);
      mat_prob -= p.focus[i].prob;
    }
    return prob + mat_prob * mat_source_density(m, m.index, normal, source, dest);
  }

  V3 recurse(Rng &g, const Ray &ray, int depth, V3 scl) {
    if (sum(scl) / 3 < p.cutoff) return V3();
    Hit h;
    int32_t obj;
    casts++;
    if (!sc.cast(ray, h, obj)) return V3();
    V3 point = add(ray.origin, scale(ray.direction, h.scale));
    MatRef m = sc.material_at(obj, point);
    V3 dest = scale(normalize(ray.direction), -1);
    V3 color = mat_emission(m, m.index);
    if (depth == 0) color = add(color, mat_ambient(m, m.index));
    for (int li = 0; li < nl; li++) {
      const m3d_point_light &l = lights[li];
      V3 light_dir = sub(v3(l.origin), point);
      Ray shadow = bounce_ray(point, light_dir, p.epsilon);
      Hit sh;
      int32_t so;
      casts++;
      if (sc.cast(shadow, sh, so) && sh.scale < 1) continue;
      V3 brdf = mat_bsdf(m, m.index, h.normal, normalize(sub(point, v3(l.origin))), dest);
      color = add(color, mul(shade_collision(l, h.normal, light_dir), brdf));
    }
    if (depth >= p.max_depth) return color;
    V3 next_source = sample_next_source(g, point, h.normal, dest, m);
    double weight = 1 / source_density(point, h.normal, next_source, dest, m);
    weight *= std::fabs(dot(next_source, h.normal));
    V3 reflect_weight = mat_bsdf(m, m.index, h.normal, next_source, dest);
    Ray next_ray = bounce_ray(point, scale(next_source, -1), p.epsilon);
    V3 next_mask = scale(reflect_weight, weight);
    V3 next_scale = mul(scl, next_mask);
    V3 next_color = recurse(g, next_ray, depth + 1, next_scale);
    return add(color, mul(next_color, next_mask));
  }
};

// ray_renderer.go:157-173
inline bool converged(double max_stddev, double oversat, V3 mean, V3 stddev) {
  for (int i = 0; i < 3; i++) {
    if (stddev[i] < max_stddev) continue;
    if (oversat != 0 && mean[i] - oversat * stddev[i] > 1) continue;
    return false;
  }
  return true;
}

// ray_renderer.go:112-151 generalised over the per-sample colour function.
// Outputs: mean (what the reference writes to img.Data), and optionally the
// estimated variance of that mean (sample variance / n) for 3-sigma parity tests.
template <class ColorFn>
inline void estimate_pixels(const m3d_camera &cam, int W, int H, int num_samples, int min_samples,
                            double max_stddev, double oversat, double antialias, uint64_t seed,
                            double *mean_out, double *var_of_mean, int nthreads, ColorFn color_fn) {
  Caster caster = make_caster(cam, double(W) - 1, double(H) - 1);
  bool has_conv = min_samples != 0 && max_stddev != 0;
  parallel_rows(H, nthreads, [&](int y, int tid) {
    Rng g(seed * 0x9E3779B97F4A7C15ull + (uint64_t)y * 1315423911ull + 17);
    (void)tid;
    for (int x = 0; x < W; x++) {
      Ray ray{v3(cam.origin), caster(double(x), double(y))};
      V3 csum, csq;
      int n = 0;
      for (n = 0; n < num_samples; n++) {
        if (antialias != 0) {
          double dx = antialias * (g.f64() - 0.5);
          double dy = antialias * (g.f64() - 0.5);
          ray.direction = caster(double(x) + dx, double(y) + dy);
        }
        V3 c = color_fn(g, ray, tid);
        csum = add(csum, c);
        csq = add(csq, mul(c, c));
        if (!has_conv) continue;
        if (n < min_samples || n < 2) continue;
        // reference quirk: statistics use the loop index (count-1), ray_renderer.go:134-146
        V3 mean = scale(csum, 1 / double(n));
        V3 var = vmax(sub(scale(csq, 1 / double(n)), mul(mean, mean)), V3());
        V3 sd = scale(V3(std::sqrt(var.x), std::sqrt(var.y), std::sqrt(var.z)),
                      std::sqrt(double(n)) / double(n - 1));
        if (converged(max_stddev, oversat, mean, sd)) break;
      }
      int idx = x + y * W;
      V3 mean = scale(csum, 1 / double(n));
      for (int k = 0; k < 3; k++) mean_out[idx * 3 + k] = mean[k];
      if (var_of_mean) {
        double nn = double(n);
        for (int k = 0; k < 3; k++) {
          double v = (csq[k] / nn - mean[k] * mean[k]) * nn / (nn - 1);
          var_of_mean[idx * 3 + k] = std::fmax(v, 0.0) / nn;
        }
      }
    }
  });
}

inline void render_path(const Scene &sc, const m3d_camera &cam, const m3d_point_light *lights, int nl,
                        const m3d_path_params &p, int W, int H, double *mean, double *var_of_mean,
                        int64_t *rays_cast, int nthreads) {
  std::vector<PathTracer> pts;
  for (int i = 0; i < std::max(1, nthreads); i++) pts.push_back(PathTracer{sc, lights, nl, p});
  estimate_pixels(cam, W, H, p.num_samples, p.min_samples, p.max_stddev, p.oversaturated_stddevs,
                  p.antialias, p.seed, mean, var_of_mean, nthreads,
                  [&](Rng &g, const Ray &ray, int tid) { return pts[tid].recurse(g, ray, 0, V3(1, 1, 1)); });
  if (rays_cast) {
    *rays_cast = 0;
    for (auto &t : pts) *rays_cast += t.casts;
  }
}

// ---- area lights (light.go) ---------------------------------------------------------------
struct AreaLights {
  struct L {
    int32_t object;
    V3 emission;
    double total;  // TotalEmission
    std::vector<double> cumu_areas;
    double total_area = 0;
  };
  std::vector<L> lights;
  std::vector<double> cumu_totals;
  double total_light = 0;

  void init(const Scene &sc, const m3d_area_light *ls, int n) {
    for (int i = 0; i < n; i++) {
      L l;
      l.object = ls[i].object;
      l.emission = v3(ls[i].emission);
      const Object &o = sc.objects[l.object];
      if (o.kind == OBJ_SPHERE) {
        l.total = sum(l.emission) * 4 * M_PI * o.sphere.radius * o.sphere.radius;  // light.go:159-161
      } else {
        for (const Triangle &t : o.mesh->tris) {  // light.go:246-251
          l.total_area += t.area();
          l.cumu_areas.push_back(l.total_area);
        }
        l.total = l.total_area * sum(l.emission);  // light.go:272-274
      }
      total_light += l.total;
      cumu_totals.push_back(total_light);
      lights.push_back(std::move(l));
    }
  }
  // sort.SearchFloat64s: smallest i with a[i] >= x
  static size_t search(const std::vector<double> &a, double x) {
    return std::lower_bound(a.begin(), a.end(), x) - a.begin();
  }
  void sample(const Scene &sc, Rng &g, V3 &point, V3 &normal, V3 &emission) const {
    size_t li = 0;
    if (lights.size() > 1) {  // light.go:303-311
      li = search(cumu_totals, g.f64() * total_light);
      if (li == cumu_totals.size()) li--;
    }
    const L &l = lights[li];
    const Object &o = sc.objects[l.object];
    emission = l.emission;
    if (o.kind == OBJ_SPHERE) {  // light.go:142-157
      for (;;) {
        normal = V3(g.normal(), g.normal(), g.normal());
        double n = norm(normal);
        if (n > 0.01 && n < 100.0) {
          normal = scale(normal, 1 / n);
          break;
        }
      }
      point = add(o.sphere.center, scale(normal, o.sphere.radius));
      return;
    }
    // light.go:254-270
    size_t ti = search(l.cumu_areas, g.f64() * l.total_area);
    if (ti == l.cumu_areas.size()) ti--;
    const Triangle &t = o.mesh->tris[ti];
    double r1 = std::sqrt(g.f64());
    double r2 = g.f64();
    V3 res = scale(t.p[0], 1 - r1);
    res = add(res, scale(t.p[1], r1 * (1 - r2)));
    res = add(res, scale(t.p[2], r1 * r2));
    point = res;
    normal = t.normal();
  }
};

// ---- BidirPathTracer (bidir.go) --------------------------------------------------------------
struct BVert {
  V3 point, normal, source, dest;
  V3 bsdf, emission;
  MatRef mat;
  bool has_mat = false;
  double source_density = 0, dest_density = 0;
  double roulette_scale = 0;
  double accumulator = 0;
  // bidir.go:338-346
  void eval_material() {
    if (!has_mat) {
      dest_density = 4 * std::fmax(0.0, dot(dest, normal));
      return;
    }
    source_density = mat_source_density(mat, mat.index, normal, source, dest);
    dest_density = mat_dest_density(mat, mat.index, normal, source, dest);
    bsdf = mat_bsdf(mat, mat.index, normal, source, dest);
  }
  double source_dot() const { return std::fabs(dot(normal, source)); }
  double dest_dot() const { return std::fabs(dot(normal, dest)); }
};

// bidir.go:264-309
struct PathEnder {
  int min_length;
  double cutoff;
  double current_roulette = 1.0;
  V3 full_mask{1, 1, 1}, roulette_mask{1, 1, 1};
  bool end(Rng &g, int i, V3 mask) {
    full_mask = mul(full_mask, mask);
    double mean = sum(full_mask) / 3;
    if (mean < cutoff) {
      double keep = mean / cutoff;
      if (g.f64() > keep) return true;
      current_roulette *= 1 / keep;
    }
    if (min_length != 0 && i + 1 >= min_length) {
      roulette_mask = mul(roulette_mask, mask);
      double mv = std::fmax(std::fmax(roulette_mask.x, roulette_mask.y), roulette_mask.z);
      if (mv < 1) {
        roulette_mask = V3(1, 1, 1);
        double keep = mv;
        if (g.f64() > keep) return true;
        current_roulette *= 1 / keep;
      }
    }
    return false;
  }
};

struct Bidir {
  const Scene &sc;
  const AreaLights &al;
  const m3d_bidir_params &p;
  int64_t casts = 0;
  std::vector<BVert> eye, light;
  std::vector<const BVert *> joined;
  BVert extra[2];

  int max_light_depth() const { return p.max_light_depth != 0 ? p.max_light_depth : p.max_depth; }

  // bidir.go:161-189
  void sample_eye_path(Rng &g, Ray ray) {
    eye.clear();
    PathEnder pe{p.min_depth, p.cutoff};
    for (int i = 0; i < p.max_depth; i++) {
      Hit h;
      int32_t obj;
      casts++;
      if (!sc.cast(ray, h, obj)) break;
      V3 point = add(ray.origin, scale(ray.direction, h.scale));
      MatRef m = sc.material_at(obj, point);
      V3 dest = normalize(scale(ray.direction, -1));
      V3 next_source = mat_sample_source(m, m.index, g, h.normal, dest);
      BVert v;
      v.point = point;
      v.normal = h.normal;
      v.source = next_source;
      v.dest = dest;
      v.emission = mat_emission(m, m.index);
      v.mat = m;
      v.has_mat = true;
      v.roulette_scale = pe.current_roulette;
      v.eval_material();
      eye.push_back(v);
      ray = bounce_ray(point, scale(next_source, -1), p.epsilon);
      if (pe.end(g, i, scale(v.bsdf, v.source_dot() / v.source_density))) break;
    }
  }

  // bidir.go:191-234
  void sample_light_path(Rng &g) {
    V3 origin, normal, emission;
    al.sample(sc, g, origin, normal, emission);
    V3 dest = scale(lambert_sample(g, normal), -1);  // sampleAngularDest bidir.go:574-576
    light.clear();
    BVert v0;
    v0.point = origin;
    v0.normal = normal;
    v0.source = scale(normal, -1);
    v0.dest = dest;
    v0.emission = emission;
    v0.roulette_scale = 1.0;
    v0.eval_material();
    light.push_back(v0);
    Ray ray = bounce_ray(origin, dest, p.epsilon);
    PathEnder pe{p.min_depth, p.cutoff};
    for (int i = 0; i < max_light_depth() - 1; i++) {
      Hit h;
      int32_t obj;
      casts++;
      if (!sc.cast(ray, h, obj)) break;
      V3 point = add(ray.origin, scale(ray.direction, h.scale));
      MatRef m = sc.material_at(obj, point);
      V3 source = ray.direction;
      V3 next_dest = mat_sample_dest(m, m.index, g, h.normal, source);
      BVert v;
      v.point = point;
      v.normal = h.normal;
      v.source = source;
      v.dest = next_dest;
      v.emission = mat_emission(m, m.index);
      v.mat = m;
      v.has_mat = true;
      v.roulette_scale = pe.current_roulette;
      v.eval_material();
      light.push_back(v);
      ray = bounce_ray(point, next_dest, p.epsilon);
      if (pe.end(g, i, scale(v.bsdf, v.dest_dot() / v.dest_density))) break;
    }
  }

  // bidir.go:532-572
  void combine_paths(int n_eye, int n_light) {
    joined.clear();
    if (n_light == 0) {
      joined.push_back(&eye[n_eye - 1]);
    } else {
      for (int i = 0; i < n_light - 1; i++) joined.push_back(&light[i]);
      const BVert &pl = light[n_light - 1];
      V3 dest = normalize(sub(eye[n_eye - 1].point, pl.point));
      BVert &v = extra[0];
      v = BVert();
      v.point = pl.point;
      v.normal = pl.normal;
      v.source = pl.source;
      v.dest = dest;
      v.emission = pl.emission;
      v.mat = pl.mat;
      v.has_mat = pl.has_mat;
      v.eval_material();
      joined.push_back(&v);
      const BVert &pe = eye[n_eye - 1];
      BVert &v1 = extra[1];
      v1 = BVert();
      v1.point = pe.point;
      v1.normal = pe.normal;
      v1.source = v.dest;
      v1.dest = pe.dest;
      v1.emission = pe.emission;
      v1.mat = pe.mat;
      v1.has_mat = pe.has_mat;
      v1.eval_material();
      joined.push_back(&v1);
    }
    for (int i = n_eye - 2; i >= 0; i--) joined.push_back(&eye[i]);
  }

  // bidir.go:421-471 over `joined`
  template <class F>
  void densities(double total_light, int max_depth, int max_ld, F f) {
    if (max_ld == 0) max_ld = max_depth;
    int n = (int)joined.size();
    std::vector<double> acc(n, 0.0);  // Accumulator lives on the vertex in the reference
    double sdp = 1.0;
    for (int i = n - 1; i > 0; i--) {
      acc[i] = sdp;
      sdp *= joined[i]->source_density;
    }
    if (n <= max_depth) f(sdp);
    auto out_area = [&](int i1, int i2) {
      V3 d = sub(joined[i1]->point, joined[i2]->point);
      return 4 * M_PI * dot(d, d);
    };
    if (n > 1) {
      double light_density = sum(joined[0]->emission) / total_light;
      if (n - 1 <= max_depth) f(light_density * acc[1] * out_area(0, 1) / joined[0]->dest_dot());
      for (int i = 0; i + 2 < n; i++) {
        if (i + 1 >= max_ld) break;
        light_density *= joined[i]->dest_density;
        light_density *= joined[i + 1]->source_dot() / joined[i]->dest_dot();
        if (n - (i + 2) <= max_depth)
          f(acc[i + 2] * light_density * out_area(i + 1, i + 2) / joined[i + 1]->dest_dot());
      }
    }
  }

  // bidir.go:101-159 + 476-530
  V3 ray_color(Rng &g, const Ray &ray) {
    sample_eye_path(g, ray);
    sample_light_path(g);
    V3 total;
    double total_emission = al.total_light;

    auto contribute = [&](double density, V3 intensity, V3 p1, V3 p2) {
      if (sum(intensity) < 1e-8) return;
      double weight = 0;
      if (p.power_heuristic == 0) {
        densities(total_emission, p.max_depth, p.max_light_depth, [&](double d) { weight += d; });
      } else {
        double s = std::pow(density, -(p.power_heuristic - 1) / p.power_heuristic);
        densities(total_emission, p.max_depth, p.max_light_depth,
                  [&](double d) { weight += std::pow(d * s, p.power_heuristic); });
      }
      V3 color = scale(intensity, 1.0 / weight);
      if (p1 != p2) {
        double brightness = maxcoord(color);
        if (p.roulette_delta > 0 && brightness < p.roulette_delta) {
          double keep = brightness / p.roulette_delta;
          if (g.f64() > keep) return;
          color = scale(color, 1 / keep);
        }
        Ray vr = bounce_ray(p1, normalize(sub(p2, p1)), p.epsilon);
        double eps = p.epsilon == 0 ? 1e-8 : p.epsilon;
        double max_dist = dist(p2, p1) - 2 * eps;
        Hit h;
        int32_t obj;
        casts++;
        if (sc.cast(vr, h, obj) && h.scale < max_dist) return;
      }
      total = add(total, color);
    };

    double eye_density = 1.0;
    V3 eye_bsdf(1, 1, 1);
    int ne = (int)eye.size(), nlp = (int)light.size();
    for (int i = 1; i <= ne; i++) {
      if (eye[i - 1].emission != V3()) {
        V3 cur = mul(eye[i - 1].emission, eye_bsdf);
        cur = scale(cur, eye[i - 1].roulette_scale);
        combine_paths(i, 0);
        contribute(eye_density, cur, V3(), V3());
      }
      double density = eye_density * sum(light[0].emission) / total_emission;
      V3 light_bsdf = light[0].emission;
      for (int j = 1; j <= nlp; j++) {
        V3 diff = sub(light[j - 1].point, eye[i - 1].point);
        double out_area = 4 * M_PI * dot(diff, diff);
        if (j > 1) {
          density *= light[j - 2].dest_density;
          density *= light[j - 1].source_dot() / light[j - 2].dest_dot();
          if (j > 2) light_bsdf = mul(light_bsdf, light[j - 2].bsdf);
          light_bsdf = scale(light_bsdf, light[j - 1].source_dot());
        }
        combine_paths(i, j);
        double dd = joined[j - 1]->dest_dot();
        if (dd > 0) {
          double sd = joined[j]->source_dot();
          if (sd > 0) {
            double cur_density = density * out_area / dd;
            V3 intensity = scale(mul(eye_bsdf, light_bsdf), sd);
            intensity = mul(intensity, joined[j]->bsdf);
            intensity = scale(intensity, light[j - 1].roulette_scale * eye[i - 1].roulette_scale);
            if (j > 1) intensity = mul(intensity, joined[j - 1]->bsdf);
            contribute(cur_density, intensity, eye[i - 1].point, light[j - 1].point);
          }
        }
      }
      eye_density *= eye[i - 1].source_density;
      eye_bsdf = scale(mul(eye_bsdf, eye[i - 1].bsdf), eye[i - 1].source_dot());
    }
    return total;
  }
};

inline void render_bidir(const Scene &sc, const m3d_camera &cam, const m3d_area_light *lights, int nl,
                         const m3d_bidir_params &p, int W, int H, double *mean, double *var_of_mean,
                         int64_t *rays_cast, int nthreads) {
  AreaLights al;
  al.init(sc, lights, nl);
  std::vector<Bidir> bs;
  for (int i = 0; i < std::max(1, nthreads); i++) bs.push_back(Bidir{sc, al, p});
  estimate_pixels(cam, W, H, p.num_samples, p.min_samples, p.max_stddev, p.oversaturated_stddevs, p.antialias,
                  p.seed, mean, var_of_mean, nthreads,
                  [&](Rng &g, const Ray &ray, int tid) { return bs[tid].ray_color(g, ray); });
  if (rays_cast) {
    *rays_cast = 0;
    for (auto &t : bs) *rays_cast += t.casts;
  }
}

}  // namespace orc
