/*
 * cgo_sequence.c -- the call sequence of the cgo binding (go/gpu3d), made from plain C.
 *
 * The build container has no Go toolchain, so go/gpu3d cannot be compiled there.  This program
 * includes the same header (include/m3d.h) with a real C compiler, links libm3dgpu.so and makes
 * exactly the calls the binding makes, in its order, with its argument conventions (float32
 * bulk arrays, float64 scalars, NULL for optional outputs, pinned buffers from m3d_host_alloc):
 *
 *   NewMultiContext / NewContext     m3d_ctx_create_multi | m3d_ctx_create, m3d_ctx_num_devices
 *   NewScene                         m3d_scene_builder_create, m3d_scene_add_material ...,
 *                                    m3d_scene_add_{mesh,sphere,rect,cylinder} ... (flags, transforms,
 *                                    vertex normals), m3d_scene_build, m3d_scene_builder_destroy,
 *                                    m3d_scene_bounds
 *   (*Scene).Cast                    m3d_scene_cast (batch of one)
 *   (*RayCaster).Render              m3d_render_raycast
 *   (*RecursiveRayTracer).Render / RenderVariance   m3d_render_path (chunks by sample index, LogFunc
 *                                    between them; sumsq for the variance)
 *   (*BidirPathTracer).Render        m3d_render_bidir (adaptive fields passed through)
 *   error path                       a failing call followed by m3d_last_error on the same thread
 *   Close                            m3d_scene_destroy, m3d_host_free, m3d_ctx_destroy
 *
 * The scene and the renderer settings come from a file written by tests/test_c_abi.py (the C3
 * cornell box and the C4 showcase scene of BASELINE.json); the sums are written back for the
 * test to compare with the Python binding's render of the same scene and seed.
 *
 *   cgo_sequence <scene.bin> <out.bin> <num_devices> <chunks>
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "m3d.h"

#define CHECK(expr)                                                                   \
  do {                                                                                \
    int32_t rc_ = (expr);                                                             \
    if (rc_ != M3D_OK) {                                                              \
      fprintf(stderr, "%s failed (%d): %s\n", #expr, (int)rc_, m3d_last_error());     \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

static int rd(FILE *f, void *dst, size_t n) { return fread(dst, 1, n, f) == n ? 0 : -1; }
#define RD(f, v)                                        \
  do {                                                  \
    if (rd(f, &(v), sizeof(v))) {                       \
      fprintf(stderr, "short read at %s\n", #v);        \
      return 1;                                         \
    }                                                   \
  } while (0)

static void log_func(double frac, double sample_rate) { fprintf(stderr, "LogFunc(%.3f, %.3e)\n", frac, sample_rate); }

int main(int argc, char **argv) {
  if (argc < 5) {
    fprintf(stderr, "usage: %s scene.bin out.bin num_devices chunks\n", argv[0]);
    return 2;
  }
  FILE *f = fopen(argv[1], "rb");
  if (!f) {
    perror(argv[1]);
    return 2;
  }
  const int num_devices = atoi(argv[3]);
  const int chunks = atoi(argv[4]) > 0 ? atoi(argv[4]) : 1;
  if (m3d_abi_version() != M3D_ABI_VERSION) {
    fprintf(stderr, "ABI mismatch: header %d, library %d\n", M3D_ABI_VERSION, (int)m3d_abi_version());
    return 1;
  }

  /* ---- NewContext / NewMultiContext ---- */
  m3d_ctx *ctx = NULL;
  if (num_devices > 1) {
    int32_t ids[16];
    for (int i = 0; i < num_devices && i < 16; i++) ids[i] = i;
    CHECK(m3d_ctx_create_multi(ids, num_devices, &ctx));
  } else {
    CHECK(m3d_ctx_create(0, &ctx));
  }
  printf("devices %d\n", (int)m3d_ctx_num_devices(ctx));

  /* ---- NewScene ---- */
  m3d_scene_builder *b = NULL;
  CHECK(m3d_scene_builder_create(ctx, &b));
  int32_t num_objects = 0, num_materials = 0;
  for (;;) {
    int32_t tag;
    RD(f, tag);
    if (tag == 0) break;
    if (tag == 1) {
      m3d_material_desc d;
      RD(f, d);
      int32_t idx = -1;
      CHECK(m3d_scene_add_material(b, &d, &idx));
      if (idx != num_materials++) {
        fprintf(stderr, "material index %d, expected %d\n", (int)idx, (int)num_materials - 1);
        return 1;
      }
      continue;
    }
    int32_t material, has_xf;
    uint32_t flags;
    m3d_transform xf;
    RD(f, material);
    RD(f, flags);
    RD(f, has_xf);
    if (has_xf) RD(f, xf);
    const m3d_transform *xp = has_xf ? &xf : NULL;
    if (tag == 2) { /* *model3d.Sphere */
      double c[3], r;
      RD(f, c);
      RD(f, r);
      CHECK(m3d_scene_add_sphere(b, c, r, material, flags, xp, NULL));
    } else if (tag == 3) { /* *model3d.Rect */
      double mn[3], mx[3];
      RD(f, mn);
      RD(f, mx);
      CHECK(m3d_scene_add_rect(b, mn, mx, material, flags, xp, NULL));
    } else if (tag == 4) { /* *model3d.Cylinder */
      double p1[3], p2[3], r;
      RD(f, p1);
      RD(f, p2);
      RD(f, r);
      CHECK(m3d_scene_add_cylinder(b, p1, p2, r, material, flags, xp, NULL));
    } else if (tag == 5) { /* *gpu3d.MeshCollider / MeshObject: flat float32 triangles */
      int64_t n;
      int32_t has_vn;
      RD(f, n);
      RD(f, has_vn);
      float *tris = (float *)malloc((size_t)n * 9 * sizeof(float) + 4);
      float *vn = has_vn ? (float *)malloc((size_t)n * 9 * sizeof(float) + 4) : NULL;
      if (rd(f, tris, (size_t)n * 9 * sizeof(float)) || (vn && rd(f, vn, (size_t)n * 9 * sizeof(float)))) {
        fprintf(stderr, "short read in mesh\n");
        return 1;
      }
      CHECK(m3d_scene_add_mesh(b, tris, n, vn, material, flags, xp, NULL));
      free(tris); /* borrowed for the duration of the call only */
      free(vn);
    } else {
      fprintf(stderr, "bad tag %d\n", (int)tag);
      return 1;
    }
    num_objects++;
  }
  /* the error path of `call`: a failing call, then m3d_last_error on the same thread */
  {
    double c[3] = {0, 0, 0};
    int32_t rc = m3d_scene_add_sphere(b, c, 1.0, 12345, 0, NULL, NULL);
    const char *msg = m3d_last_error();
    if (rc != M3D_ERR_INVALID_ARG || !msg || !strstr(msg, "material index")) {
      fprintf(stderr, "error path: rc %d msg '%s'\n", (int)rc, msg ? msg : "(null)");
      return 1;
    }
  }
  m3d_scene *scene = NULL;
  CHECK(m3d_scene_build(b, 0, &scene));
  m3d_scene_builder_destroy(b);
  double mn[3], mx[3];
  CHECK(m3d_scene_bounds(scene, mn, mx));
  printf("objects %d materials %d bounds %.6g %.6g %.6g .. %.6g %.6g %.6g\n", (int)num_objects, (int)num_materials,
         mn[0], mn[1], mn[2], mx[0], mx[1], mx[2]);

  /* ---- renderer settings ---- */
  int32_t mode, W, H;
  m3d_camera cam;
  RD(f, mode);
  RD(f, cam);
  RD(f, W);
  RD(f, H);
  const size_t npix = (size_t)W * H;

  /* (*Scene).Cast: the centre pixel's ray as a batch of one */
  {
    float org[3] = {(float)cam.origin[0], (float)cam.origin[1], (float)cam.origin[2]};
    double z[3] = {cam.screen_x[1] * cam.screen_y[2] - cam.screen_x[2] * cam.screen_y[1],
                   cam.screen_x[2] * cam.screen_y[0] - cam.screen_x[0] * cam.screen_y[2],
                   cam.screen_x[0] * cam.screen_y[1] - cam.screen_x[1] * cam.screen_y[0]};
    float dir[3] = {(float)z[0], (float)z[1], (float)z[2]};
    float t = 0, normal[3];
    int32_t obj = -1, prim = -1;
    CHECK(m3d_scene_cast(scene, org, dir, 1, &t, &obj, &prim, normal, 0, NULL));
    printf("cast obj %d prim %d t %.6g\n", (int)obj, (int)prim, (double)t);
  }

  /* pinned result buffers: NewHostFloats */
  float *sum = NULL, *sq = NULL, *chunk_sum = NULL, *chunk_sq = NULL;
  CHECK(m3d_host_alloc((int64_t)(npix * 3 * sizeof(float)), (void **)&sum));
  memset(sum, 0, npix * 3 * sizeof(float));
  m3d_stats st;
  memset(&st, 0, sizeof(st));
  int64_t samples_taken = 0, rays = 0;
  int32_t want_variance = 0;

  if (mode == 0) { /* (*RayCaster).Render */
    int32_t nl;
    RD(f, nl);
    m3d_point_light *lights = (m3d_point_light *)calloc((size_t)nl + 1, sizeof(m3d_point_light));
    if (nl && rd(f, lights, (size_t)nl * sizeof(m3d_point_light))) return 1;
    CHECK(m3d_render_raycast(scene, &cam, lights, nl, W, H, NULL, sum, &st));
    rays = st.rays;
    free(lights);
  } else if (mode == 1) { /* (*RecursiveRayTracer).Render / RenderVariance */
    m3d_path_params p;
    int32_t nl, num_samples;
    RD(f, p);
    RD(f, nl);
    m3d_point_light *lights = (m3d_point_light *)calloc((size_t)nl + 1, sizeof(m3d_point_light));
    if (nl && rd(f, lights, (size_t)nl * sizeof(m3d_point_light))) return 1;
    RD(f, num_samples);
    RD(f, want_variance);
    CHECK(m3d_host_alloc((int64_t)(npix * 3 * sizeof(float)), (void **)&chunk_sum));
    if (want_variance) {
      CHECK(m3d_host_alloc((int64_t)(npix * 3 * sizeof(float)), (void **)&sq));
      CHECK(m3d_host_alloc((int64_t)(npix * 3 * sizeof(float)), (void **)&chunk_sq));
      memset(sq, 0, npix * 3 * sizeof(float));
    }
    /* renderChunks: sample-index chunks, each call ADDS into the frame sums, LogFunc in between */
    for (int i = 0; i < chunks; i++) {
      m3d_partition part;
      memset(&part, 0, sizeof(part));
      const int b0 = (int)((int64_t)i * num_samples / chunks), b1 = (int)((int64_t)(i + 1) * num_samples / chunks);
      if (b1 == b0) continue;
      part.sample_begin = b0;
      CHECK(m3d_render_path(scene, &cam, lights, nl, &p, W, H, &part, b1 - b0, chunk_sum, chunk_sq, &st));
      for (size_t k = 0; k < npix * 3; k++) sum[k] += chunk_sum[k];
      if (sq)
        for (size_t k = 0; k < npix * 3; k++) sq[k] += chunk_sq[k];
      samples_taken += st.samples;
      rays += st.rays;
      log_func((double)(i + 1) / chunks, (double)samples_taken);
    }
    free(lights);
  } else if (mode == 2) { /* (*BidirPathTracer).Render */
    m3d_bidir_params p;
    int32_t nl, num_samples;
    RD(f, p);
    RD(f, nl);
    m3d_area_light *lights = (m3d_area_light *)calloc((size_t)nl + 1, sizeof(m3d_area_light));
    if (rd(f, lights, (size_t)nl * sizeof(m3d_area_light))) return 1;
    RD(f, num_samples);
    RD(f, want_variance);
    CHECK(m3d_host_alloc((int64_t)(npix * 3 * sizeof(float)), (void **)&chunk_sum));
    for (int i = 0; i < chunks; i++) {
      m3d_partition part;
      memset(&part, 0, sizeof(part));
      const int b0 = (int)((int64_t)i * num_samples / chunks), b1 = (int)((int64_t)(i + 1) * num_samples / chunks);
      if (b1 == b0) continue;
      part.sample_begin = b0;
      CHECK(m3d_render_bidir(scene, &cam, lights, nl, &p, W, H, &part, b1 - b0, chunk_sum, NULL, &st));
      for (size_t k = 0; k < npix * 3; k++) sum[k] += chunk_sum[k];
      samples_taken += st.samples;
      rays += st.rays;
      log_func((double)(i + 1) / chunks, (double)samples_taken);
    }
    free(lights);
  } else {
    fprintf(stderr, "bad mode %d\n", (int)mode);
    return 1;
  }
  fclose(f);
  printf("samples %lld rays %lld\n", (long long)samples_taken, (long long)rays);

  FILE *o = fopen(argv[2], "wb");
  if (!o) {
    perror(argv[2]);
    return 2;
  }
  fwrite(&W, sizeof(W), 1, o);
  fwrite(&H, sizeof(H), 1, o);
  fwrite(&want_variance, sizeof(want_variance), 1, o);
  fwrite(sum, sizeof(float), npix * 3, o);
  if (sq) fwrite(sq, sizeof(float), npix * 3, o);
  fclose(o);

  /* ---- Close ---- */
  CHECK(m3d_host_free(sum));
  CHECK(m3d_host_free(sq));
  CHECK(m3d_host_free(chunk_sum));
  CHECK(m3d_host_free(chunk_sq));
  m3d_scene_destroy(scene);
  CHECK(m3d_ctx_trim(ctx));
  m3d_ctx_destroy(ctx);
  printf("ok\n");
  return 0;
}
