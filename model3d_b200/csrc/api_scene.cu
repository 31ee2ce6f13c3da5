// C ABI: scene construction, Object.Cast batches and RayCaster (include/m3d.h).
// Replaces render3d.JoinedObject / ColliderObject / Translate / MatrixMultiply
// (render3d/object.go:26-153, transform.go:6-85) and (*RayCaster).Render
// (render3d/raycast.go:15-39).
#include <cmath>
#include <cstring>
#include <memory>

#include "api_common.h"
#include "scene.h"
#include "scene_host.h"
#include "trace_core.cuh"

using namespace m3d;

#define M3D_HAVE_SCENE 1
#define M3D_HAVE_RAYCAST 1

struct BuilderObject {
  int kind = 0;  // 0 mesh, else ShapeKind
  int32_t material = 0;
  uint32_t flags = 0;
  std::vector<float> tris;      // world space
  std::vector<float> vnormals;  // optional, world space
  DeviceShape shape{};
  double bmin[3], bmax[3];
  m3d_mesh *mesh = nullptr;  // SHAPE_INSTANCE
  m3d::DeviceInstance inst{};
};

struct m3d_scene_builder {
  m3d_ctx *ctx = nullptr;
  std::vector<m3d_material_desc> materials;
  std::vector<BuilderObject> objects;
};

namespace {

// more analytic shapes / instances than this: the scene gets an object-level BVH (a linear scan
// is faster for the handful of shapes of the BASELINE scenes)
constexpr size_t kShapeBvhThreshold = 12;

// x_world = M x + offset; only similarity transforms (M^T M = s^2 I) keep spheres spheres and
// make the reference's "normal = normalize(M n)" (transform.go:81-83) the true normal.
bool similarity_scale(const m3d_transform *xf, double &scale) {
  if (!xf) {
    scale = 1;
    return true;
  }
  const double *m = xf->matrix;
  double c[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) c[i][j] = m[0 + i] * m[0 + j] + m[3 + i] * m[3 + j] + m[6 + i] * m[6 + j];
  const double s2 = c[0][0];
  if (!(s2 > 0)) return false;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      const double want = i == j ? s2 : 0.0;
      if (std::fabs(c[i][j] - want) > 1e-9 * s2) return false;
    }
  scale = std::sqrt(s2);
  return true;
}

void apply_xf(const m3d_transform *xf, const double in[3], double out[3]) {
  if (!xf) {
    out[0] = in[0];
    out[1] = in[1];
    out[2] = in[2];
    return;
  }
  const double *m = xf->matrix;  // row-major (model3d/matrix.go:11-12,131-137)
  for (int r = 0; r < 3; r++)
    out[r] = m[3 * r] * in[0] + m[3 * r + 1] * in[1] + m[3 * r + 2] * in[2] + xf->offset[r];
}

bool is_identity_rotation(const m3d_transform *xf) {
  if (!xf) return true;
  const double id[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < 9; i++)
    if (xf->matrix[i] != id[i]) return false;
  return true;
}

int32_t check_material_index(m3d_scene_builder *b, int32_t material) {
  if (material < 0 || material >= (int32_t)b->materials.size())
    return fail(M3D_ERR_INVALID_ARG, "material index %d out of range (%zu materials)", material,
                b->materials.size());
  return M3D_OK;
}

// Cylinder.Min/Max (model3d/shapes.go:543-584)
void cylinder_bounds(const double p1[3], const double p2[3], double r, double mn[3], double mx[3]) {
  double axis[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  const double an = std::sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
  for (int k = 0; k < 3; k++) {
    double e[3] = {0, 0, 0};
    e[k] = 1;
    const double dn = (axis[0] * e[0] + axis[1] * e[1] + axis[2] * e[2]) / an;
    double proj[3];
    for (int j = 0; j < 3; j++) proj[j] = e[j] - axis[j] / an * dn;
    const double pn = std::sqrt(proj[0] * proj[0] + proj[1] * proj[1] + proj[2] * proj[2]);
    const double bound = std::fabs(proj[k] / (pn + 1e-8)) + 1e-8;
    mn[k] = std::fmin(p1[k], p2[k]) - bound * r;
    mx[k] = std::fmax(p1[k], p2[k]) + bound * r;
  }
}

DeviceCamera make_device_camera(const m3d_camera &c, int W, int H) {
  // Camera.axes with (W-1, H-1) as RayCaster / rayRenderer pass them (camera.go:100-113,
  // raycast.go:16-18, ray_renderer.go:29-31)
  const double w = (double)W - 1, h = (double)H - 1;
  const double plane = 1.0 / std::tan(c.field_of_view / 2);
  double x[3], y[3], z[3];
  for (int i = 0; i < 3; i++) {
    x[i] = c.screen_x[i];
    y[i] = c.screen_y[i];
  }
  z[0] = x[1] * y[2] - x[2] * y[1];
  z[1] = x[2] * y[0] - x[0] * y[2];
  z[2] = x[0] * y[1] - x[1] * y[0];
  const double zn = 1.0 / std::sqrt(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]);
  for (int i = 0; i < 3; i++) z[i] = z[i] * zn * plane;
  if (w > h) {
    for (int i = 0; i < 3; i++) y[i] *= h / w;
  } else {
    for (int i = 0; i < 3; i++) x[i] *= w / h;
  }
  DeviceCamera d;
  for (int i = 0; i < 3; i++) {
    d.origin[i] = c.origin[i];
    d.x[i] = x[i];
    d.y[i] = y[i];
    d.z[i] = z[i];
  }
  d.cx = w / 2;
  d.cy = h / 2;
  return d;
}

}  // namespace

namespace m3d {
const DeviceScene &scene_device(const m3d_scene *s) { return s->dev; }
m3d_ctx *scene_ctx(const m3d_scene *s) { return s->ctx; }
DeviceCamera device_camera(const m3d_camera &c, int W, int H) { return make_device_camera(c, W, H); }
const std::vector<m3d_material_desc> &scene_materials(const m3d_scene *s) { return s->host_materials; }
}  // namespace m3d

extern "C" {

int32_t m3d_scene_builder_create(m3d_ctx *ctx, m3d_scene_builder **out) {
  if (!ctx || !out) return fail(M3D_ERR_INVALID_ARG, "m3d_scene_builder_create: NULL argument");
  *out = new m3d_scene_builder();
  (*out)->ctx = ctx;
  return M3D_OK;
}

void m3d_scene_builder_destroy(m3d_scene_builder *b) { delete b; }

int32_t m3d_scene_add_material(m3d_scene_builder *b, const m3d_material_desc *mat, int32_t *index_out) {
  if (!b || !mat) return fail(M3D_ERR_INVALID_ARG, "m3d_scene_add_material: NULL argument");
  if (mat->kind < M3D_MAT_LAMBERT || mat->kind > M3D_MAT_JOINED)
    return fail(M3D_ERR_UNSUPPORTED, "material kind %d is not supported on the GPU path", mat->kind);
  if (mat->kind == M3D_MAT_JOINED) {
    if (mat->num_sub < 1 || mat->num_sub > M3D_MAX_SUBMATERIALS)
      return fail(M3D_ERR_UNSUPPORTED, "JoinedMaterial with %d parts (max %d)", mat->num_sub, M3D_MAX_SUBMATERIALS);
    for (int i = 0; i < mat->num_sub; i++) {
      if (mat->sub[i] < 0 || mat->sub[i] >= (int32_t)b->materials.size())
        return fail(M3D_ERR_INVALID_ARG, "JoinedMaterial part %d refers to material %d which does not exist yet", i, mat->sub[i]);
      if (b->materials[mat->sub[i]].kind == M3D_MAT_JOINED)
        return fail(M3D_ERR_UNSUPPORTED, "nested JoinedMaterial is not supported on the GPU path");
    }
  }
  b->materials.push_back(*mat);
  if (index_out) *index_out = (int32_t)b->materials.size() - 1;
  return M3D_OK;
}

int32_t m3d_scene_add_mesh(m3d_scene_builder *b, const float *tris, int64_t n, const float *vnormals,
                           int32_t material, uint32_t flags, const m3d_transform *xf, int32_t *index_out) {
  if (!b || n < 0 || (n > 0 && !tris)) return fail(M3D_ERR_INVALID_ARG, "m3d_scene_add_mesh: bad arguments");
  if (int32_t rc = check_material_index(b, material)) return rc;
  double scale;
  if (!similarity_scale(xf, scale))
    return fail(M3D_ERR_UNSUPPORTED, "mesh transform must be a similarity (rotation * uniform scale)");
  BuilderObject o;
  o.kind = 0;
  o.material = material;
  o.flags = flags;
  o.tris.resize((size_t)n * 9);
  for (int k = 0; k < 3; k++) {
    o.bmin[k] = INFINITY;
    o.bmax[k] = -INFINITY;
  }
  for (int64_t v = 0; v < n * 3; v++) {
    double in[3] = {tris[v * 3], tris[v * 3 + 1], tris[v * 3 + 2]}, out[3];
    apply_xf(xf, in, out);
    for (int k = 0; k < 3; k++) {
      const float f = (float)out[k];
      if (!(f == f) || std::fabs(f) > 3e38f) return fail(M3D_ERR_INVALID_ARG, "non-finite vertex coordinate");
      o.tris[v * 3 + k] = f;
      o.bmin[k] = std::fmin(o.bmin[k], (double)f);
      o.bmax[k] = std::fmax(o.bmax[k], (double)f);
    }
  }
  if (n == 0)
    for (int k = 0; k < 3; k++) o.bmin[k] = o.bmax[k] = 0;
  if (vnormals) {
    o.vnormals.resize((size_t)n * 9);
    for (int64_t v = 0; v < n * 3; v++) {
      double in[3] = {vnormals[v * 3], vnormals[v * 3 + 1], vnormals[v * 3 + 2]}, out[3] = {in[0], in[1], in[2]};
      if (xf) {
        const double *m = xf->matrix;
        for (int r = 0; r < 3; r++) out[r] = (m[3 * r] * in[0] + m[3 * r + 1] * in[1] + m[3 * r + 2] * in[2]) / scale;
      }
      for (int k = 0; k < 3; k++) o.vnormals[v * 3 + k] = (float)out[k];
    }
  }
  b->objects.push_back(std::move(o));
  if (index_out) *index_out = (int32_t)b->objects.size() - 1;
  return M3D_OK;
}

int32_t m3d_scene_add_sphere(m3d_scene_builder *b, const double center[3], double radius, int32_t material,
                             uint32_t flags, const m3d_transform *xf, int32_t *index_out) {
  if (!b || !center || !(radius > 0)) return fail(M3D_ERR_INVALID_ARG, "m3d_scene_add_sphere: bad arguments");
  if (int32_t rc = check_material_index(b, material)) return rc;
  double scale;
  if (!similarity_scale(xf, scale))
    return fail(M3D_ERR_UNSUPPORTED, "sphere transform must be a similarity (rotation * uniform scale)");
  BuilderObject o;
  o.kind = SHAPE_SPHERE;
  o.material = material;
  o.flags = flags;
  o.shape.kind = SHAPE_SPHERE;
  apply_xf(xf, center, o.shape.p0);
  o.shape.radius = radius * scale;
  for (int k = 0; k < 3; k++) {
    o.shape.p1[k] = o.shape.p0[k];
    o.bmin[k] = o.shape.p0[k] - o.shape.radius;
    o.bmax[k] = o.shape.p0[k] + o.shape.radius;
  }
  b->objects.push_back(std::move(o));
  if (index_out) *index_out = (int32_t)b->objects.size() - 1;
  return M3D_OK;
}

int32_t m3d_scene_add_rect(m3d_scene_builder *b, const double mn[3], const double mx[3], int32_t material,
                           uint32_t flags, const m3d_transform *xf, int32_t *index_out) {
  if (!b || !mn || !mx) return fail(M3D_ERR_INVALID_ARG, "m3d_scene_add_rect: bad arguments");
  if (int32_t rc = check_material_index(b, material)) return rc;
  if (!is_identity_rotation(xf))
    return fail(M3D_ERR_UNSUPPORTED, "model3d.Rect supports translation only on the GPU path (it is axis-aligned)");
  BuilderObject o;
  o.kind = SHAPE_RECT;
  o.material = material;
  o.flags = flags;
  o.shape.kind = SHAPE_RECT;
  apply_xf(xf, mn, o.shape.p0);
  apply_xf(xf, mx, o.shape.p1);
  o.shape.radius = 0;
  for (int k = 0; k < 3; k++) {
    o.bmin[k] = o.shape.p0[k];
    o.bmax[k] = o.shape.p1[k];
  }
  b->objects.push_back(std::move(o));
  if (index_out) *index_out = (int32_t)b->objects.size() - 1;
  return M3D_OK;
}

int32_t m3d_scene_add_cylinder(m3d_scene_builder *b, const double p1[3], const double p2[3], double radius,
                               int32_t material, uint32_t flags, const m3d_transform *xf,
                               int32_t *index_out) {
  if (!b || !p1 || !p2 || !(radius > 0)) return fail(M3D_ERR_INVALID_ARG, "m3d_scene_add_cylinder: bad arguments");
  if (int32_t rc = check_material_index(b, material)) return rc;
  double scale;
  if (!similarity_scale(xf, scale))
    return fail(M3D_ERR_UNSUPPORTED, "cylinder transform must be a similarity (rotation * uniform scale)");
  BuilderObject o;
  o.kind = SHAPE_CYLINDER;
  o.material = material;
  o.flags = flags;
  o.shape.kind = SHAPE_CYLINDER;
  apply_xf(xf, p1, o.shape.p0);
  apply_xf(xf, p2, o.shape.p1);
  o.shape.radius = radius * scale;
  cylinder_bounds(o.shape.p0, o.shape.p1, o.shape.radius, o.bmin, o.bmax);
  b->objects.push_back(std::move(o));
  if (index_out) *index_out = (int32_t)b->objects.size() - 1;
  return M3D_OK;
}

int32_t m3d_scene_add_instance(m3d_scene_builder *b, m3d_mesh *mesh, int32_t material, uint32_t flags,
                               const m3d_transform *xf, int32_t *index_out) {
  if (!b || !mesh) return fail(M3D_ERR_INVALID_ARG, "m3d_scene_add_instance: bad arguments");
  if (mesh->ctx != b->ctx)
    return fail(M3D_ERR_INVALID_ARG, "the instanced mesh was built on another context than the scene");
  if (mesh->info.num_triangles == 0) return fail(M3D_ERR_INVALID_ARG, "cannot instance an empty mesh");
  if (int32_t rc = check_material_index(b, material)) return rc;
  double scale;
  if (!similarity_scale(xf, scale))
    return fail(M3D_ERR_UNSUPPORTED, "instance transform must be a similarity (rotation * uniform scale)");
  BuilderObject o;
  o.kind = SHAPE_INSTANCE;
  o.material = material;
  o.flags = flags;
  o.mesh = mesh;
  o.shape.kind = SHAPE_INSTANCE;
  o.shape.radius = 0;
  // world bounds: the eight corners of the mesh's bounds through the transform
  for (int k = 0; k < 3; k++) {
    o.bmin[k] = INFINITY;
    o.bmax[k] = -INFINITY;
  }
  for (int c = 0; c < 8; c++) {
    const double in[3] = {(c & 1) ? mesh->bmax[0] : mesh->bmin[0], (c & 2) ? mesh->bmax[1] : mesh->bmin[1],
                          (c & 4) ? mesh->bmax[2] : mesh->bmin[2]};
    double out[3];
    apply_xf(xf, in, out);
    for (int k = 0; k < 3; k++) {
      o.bmin[k] = std::fmin(o.bmin[k], out[k]);
      o.bmax[k] = std::fmax(o.bmax[k], out[k]);
    }
  }
  double size = 0;
  for (int k = 0; k < 3; k++) {
    // float32 rays are tested against these bounds: widen them by a few ulp of their magnitude
    const double pad = 1e-6 * std::fmax(std::fabs(o.bmin[k]), std::fabs(o.bmax[k])) + 1e-30;
    o.bmin[k] -= pad;
    o.bmax[k] += pad;
    o.shape.p0[k] = o.bmin[k];
    o.shape.p1[k] = o.bmax[k];
    size = std::fmax(size, o.bmax[k] - o.bmin[k]);
  }
  const double id[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  const double *m = xf ? xf->matrix : id;
  // inverse of a similarity: M^-1 = M^T / s^2
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      o.inst.fwd[3 * r + c] = (float)m[3 * r + c];
      o.inst.inv[3 * r + c] = (float)(m[3 * c + r] / (scale * scale));
    }
  for (int k = 0; k < 3; k++) o.inst.off[k] = xf ? (float)xf->offset[k] : 0.f;
  o.inst.size = (float)size;
  b->objects.push_back(std::move(o));
  if (index_out) *index_out = (int32_t)b->objects.size() - 1;
  return M3D_OK;
}

int32_t m3d_scene_build(m3d_scene_builder *b, uint32_t build_flags, m3d_scene **out) {
  if (!b || !out) return fail(M3D_ERR_INVALID_ARG, "m3d_scene_build: NULL argument");
  *out = nullptr;
  if (b->objects.empty()) return fail(M3D_ERR_INVALID_ARG, "scene has no objects");
  if (b->objects.size() > 0x7fffff) return fail(M3D_ERR_INVALID_ARG, "too many objects");
  m3d_ctx *ctx = b->ctx;
  M3D_LOCK(ctx);
  M3D_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<m3d_scene> sc(new m3d_scene());
  sc->ctx = ctx;
  sc->host_materials = b->materials;

  // merge all mesh objects into one triangle soup with (object id, prim id in object)
  std::vector<int32_t> prim_ids, obj_ids;
  std::vector<float> vnormals;
  bool any_vn = false, all_vn = true;
  for (auto &o : b->objects)
    if (o.kind == 0 && !o.tris.empty()) {
      any_vn |= !o.vnormals.empty();
      all_vn &= !o.vnormals.empty();
    }
  // A scene that mixes MeshToCollider and MeshToInterpNormalCollider objects (smooth_shading/main.go)
  // keeps one per-corner normal array for the merged BVH: flat meshes get their face normal at all
  // three corners, which the interpolation (primitives.go:508-516) returns unchanged.
  (void)all_vn;
  std::vector<DeviceObject> dobjs;
  for (size_t oi = 0; oi < b->objects.size(); oi++) {
    const BuilderObject &o = b->objects[oi];
    DeviceObject dobj;
    dobj.material = o.material;
    dobj.flags = o.flags;
    dobjs.push_back(dobj);
    sc->object_material.push_back(o.material);
    sc->object_kind.push_back(o.kind);
    sc->object_tri_begin.push_back((int64_t)sc->merged_tris.size() / 9);
    sc->object_tri_count.push_back(o.kind == 0 ? (int64_t)o.tris.size() / 9 : 0);
    for (int k = 0; k < 3; k++) {
      sc->bmin[k] = oi == 0 ? o.bmin[k] : std::fmin(sc->bmin[k], o.bmin[k]);
      sc->bmax[k] = oi == 0 ? o.bmax[k] : std::fmax(sc->bmax[k], o.bmax[k]);
    }
    if (o.kind == 0) {
      const int64_t n = (int64_t)o.tris.size() / 9;
      sc->merged_tris.insert(sc->merged_tris.end(), o.tris.begin(), o.tris.end());
      if (any_vn && !o.vnormals.empty()) {
        vnormals.insert(vnormals.end(), o.vnormals.begin(), o.vnormals.end());
      } else if (any_vn) {
        for (int64_t i = 0; i < n; i++) {
          const float *t = o.tris.data() + 9 * i;
          const double ax = (double)t[3] - t[0], ay = (double)t[4] - t[1], az = (double)t[5] - t[2];
          const double bx = (double)t[6] - t[0], by = (double)t[7] - t[1], bz = (double)t[8] - t[2];
          double nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
          const double len = std::sqrt(nx * nx + ny * ny + nz * nz);
          const double inv = len > 0.0 ? 1.0 / len : 0.0;  // degenerate triangles are never hit
          for (int k = 0; k < 3; k++) {
            vnormals.push_back((float)(nx * inv));
            vnormals.push_back((float)(ny * inv));
            vnormals.push_back((float)(nz * inv));
          }
        }
      }
      for (int64_t i = 0; i < n; i++) {
        prim_ids.push_back((int32_t)i);
        obj_ids.push_back((int32_t)oi);
      }
    } else {
      DeviceShape sh = o.shape;
      sh.object = (int32_t)oi;
      sh.instance = -1;
      if (o.kind == SHAPE_INSTANCE) {
        sh.instance = (int32_t)sc->host_instances.size();
        DeviceInstance in = o.inst;
        in.blas = o.mesh->bvh;
        sc->host_instances.push_back(in);
        sc->instance_meshes.push_back(o.mesh);
      }
      sc->host_shapes.push_back(sh);
    }
  }
  const int64_t ntri = (int64_t)prim_ids.size();
  if (ntri > (int64_t)0x7fffff00) return fail(M3D_ERR_INVALID_ARG, "too many triangles");

  WideBVH bvh;
  BuildInput in;
  in.tris = sc->merged_tris.data();
  in.n = ntri;
  in.prim_ids = prim_ids.data();
  in.obj_ids = obj_ids.data();
  if (int32_t rc = build_bvh_with_flags(ctx, in, build_flags, bvh)) return rc;
  if (bvh.max_depth > M3D_MAX_BVH_DEPTH)  // see m3d_mesh_create: deeper trees would overflow the traversal stacks
    return fail(M3D_ERR_UNSUPPORTED, "BVH depth %d exceeds the traversal stack (%d levels)", bvh.max_depth,
                M3D_MAX_BVH_DEPTH);
  sc->leaf_of_merged.assign((size_t)ntri, -1);
  for (size_t li = 0; li < bvh.tris.size(); li++)
    sc->leaf_of_merged[(size_t)(sc->object_tri_begin[bvh.tris[li].object] + bvh.tris[li].prim)] = (int32_t)li;
  // upload_bvh remaps vnormals through TriRecord.prim, which is per-object here: give it a
  // table indexed by merged position instead
  std::vector<float> vn_by_leaf;
  if (any_vn && ntri > 0) {
    // map (object, prim) -> merged index
    WideBVH tmp = bvh;  // copy records, patch prim to merged index for the remap only
    for (auto &tr : tmp.tris) tr.prim = (int32_t)(sc->object_tri_begin[tr.object] + tr.prim);
    int32_t rc = upload_bvh(ctx, tmp, vnormals.data(), sc->nodes, sc->tris, sc->vnormals, sc->dev.bvh);
    if (rc != M3D_OK) return rc;
    // re-upload the true records (prim = index inside the object)
    M3D_CUDA(cudaMemcpy(sc->tris.p, bvh.tris.data(), bvh.tris.size() * sizeof(TriRecord), cudaMemcpyHostToDevice));
  } else {
    int32_t rc = upload_bvh(ctx, bvh, nullptr, sc->nodes, sc->tris, sc->vnormals, sc->dev.bvh);
    if (rc != M3D_OK) return rc;
  }

  // shapes, objects, materials
  if (!sc->host_shapes.empty()) {
    M3D_CUDA(sc->shapes.reserve(sc->host_shapes.size() * sizeof(DeviceShape)));
    M3D_CUDA(cudaMemcpy(sc->shapes.p, sc->host_shapes.data(), sc->host_shapes.size() * sizeof(DeviceShape),
                        cudaMemcpyHostToDevice));
  }
  if (!sc->host_instances.empty()) {
    M3D_CUDA(sc->instances.reserve(sc->host_instances.size() * sizeof(DeviceInstance)));
    M3D_CUDA(cudaMemcpy(sc->instances.p, sc->host_instances.data(),
                        sc->host_instances.size() * sizeof(DeviceInstance), cudaMemcpyHostToDevice));
    sc->dev.instances = sc->instances.as<const DeviceInstance>();
    sc->dev.num_instances = (int32_t)sc->host_instances.size();
  }
  // Object-level hierarchy (BVHToObject, object.go:172-185) over the bounds of the analytic shapes
  // and instances once there are more than a handful: every shape becomes a proxy "triangle" whose
  // three vertices span its bounds (the builder only looks at bounds) and whose prim id is the
  // shape index; the finish pass walks the wide BVH and tests the shapes in the leaves it reaches.
  if (sc->host_shapes.size() > kShapeBvhThreshold || !sc->host_instances.empty()) {
    std::vector<float> proxy(sc->host_shapes.size() * 9);
    std::vector<int32_t> ids(sc->host_shapes.size()), zeros(sc->host_shapes.size(), 0);
    size_t si = 0;
    for (size_t oi = 0; oi < b->objects.size(); oi++) {
      const BuilderObject &o = b->objects[oi];
      if (o.kind == 0) continue;
      float lo[3], hi[3];
      for (int k = 0; k < 3; k++) {  // outward-rounded float32 bounds
        lo[k] = std::nextafterf((float)o.bmin[k], -INFINITY);
        hi[k] = std::nextafterf((float)o.bmax[k], INFINITY);
      }
      const float v[9] = {lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], lo[0], hi[1], lo[2]};
      std::memcpy(proxy.data() + 9 * si, v, sizeof(v));
      ids[si] = (int32_t)si;
      si++;
    }
    WideBVH sbvh;
    BuildInput sin;
    sin.tris = proxy.data();
    sin.n = (int64_t)sc->host_shapes.size();
    sin.prim_ids = ids.data();
    sin.obj_ids = zeros.data();
    build_wide_bvh(sin, sbvh);
    if (sbvh.max_depth > M3D_MAX_BVH_DEPTH)
      return fail(M3D_ERR_UNSUPPORTED, "object-level BVH depth %d exceeds the traversal stack", sbvh.max_depth);
    DevBuf unused;
    if (int32_t rc = upload_bvh(ctx, sbvh, nullptr, sc->shape_nodes, sc->shape_tris, unused, sc->dev.shape_bvh))
      return rc;
  }
  M3D_CUDA(sc->objects.reserve(dobjs.size() * sizeof(DeviceObject)));
  M3D_CUDA(cudaMemcpy(sc->objects.p, dobjs.data(), dobjs.size() * sizeof(DeviceObject), cudaMemcpyHostToDevice));
  std::vector<DeviceMaterial> dmats(b->materials.size());
  for (size_t i = 0; i < b->materials.size(); i++) {
    const m3d_material_desc &m = b->materials[i];
    DeviceMaterial &d = dmats[i];
    std::memset(&d, 0, sizeof(d));
    d.kind = m.kind;
    d.flags = m.flags;
    for (int k = 0; k < 3; k++) {
      d.diffuse[k] = (float)m.diffuse[k];
      d.specular[k] = (float)m.specular[k];
      d.emission[k] = (float)m.emission[k];
      d.ambient[k] = (float)m.ambient[k];
      d.refract[k] = (float)m.refract[k];
      d.diffuse2[k] = (float)m.diffuse2[k];
    }
    d.alpha = (float)m.alpha;
    d.ior = (float)m.index_of_refraction;
    d.proc_param = (float)m.proc_param;
    d.num_sub = m.kind == M3D_MAT_JOINED ? m.num_sub : 0;
    for (int k = 0; k < M3D_MAX_SUBMATERIALS; k++) {
      d.sub[k] = m.sub[k];
      d.sub_prob[k] = (float)m.sub_prob[k];
    }
  }
  if (!dmats.empty()) {
    M3D_CUDA(sc->materials.reserve(dmats.size() * sizeof(DeviceMaterial)));
    M3D_CUDA(cudaMemcpy(sc->materials.p, dmats.data(), dmats.size() * sizeof(DeviceMaterial), cudaMemcpyHostToDevice));
  }
  sc->dev.shapes = sc->shapes.as<const DeviceShape>();
  sc->dev.num_shapes = (int32_t)sc->host_shapes.size();
  sc->dev.spheres_only = 1;
  for (const DeviceShape &sh : sc->host_shapes)
    if (sh.kind != SHAPE_SPHERE) sc->dev.spheres_only = 0;
  sc->dev.objects = sc->objects.as<const DeviceObject>();
  sc->dev.num_objects = (int32_t)dobjs.size();
  sc->dev.materials = sc->materials.as<const DeviceMaterial>();
  sc->dev.num_materials = (int32_t)dmats.size();
  sc->info.num_triangles = ntri;
  sc->info.num_nodes = (int64_t)bvh.nodes.size();
  sc->info.node_bytes = sizeof(WideNode);
  sc->info.tri_bytes = sizeof(TriRecord);
  sc->info.device_bytes = (int64_t)(sc->nodes.bytes + sc->tris.bytes + sc->vnormals.bytes);
  sc->info.max_depth = bvh.max_depth;
  sc->info.build_ms = bvh.build_ms;
  sc->info.sah_cost = bvh.sah_cost;
  m3d_scene *raw = sc.release();
  if (int32_t rc = replicate_scene(raw)) {
    m3d_scene_destroy(raw);
    return rc;
  }
  *out = raw;
  return M3D_OK;
}

void m3d_scene_destroy(m3d_scene *scene) {
  if (!scene) return;
  for (m3d_scene *r : scene->replicas) m3d_scene_destroy(r);
  scene->replicas.clear();
  std::lock_guard<std::recursive_mutex> lock(scene->ctx->mu);
  cudaSetDevice(scene->ctx->device);
  delete scene;
}

int32_t m3d_scene_bounds(const m3d_scene *scene, double min_out[3], double max_out[3]) {
  if (!scene || !min_out || !max_out) return fail(M3D_ERR_INVALID_ARG, "m3d_scene_bounds: NULL argument");
  for (int k = 0; k < 3; k++) {
    min_out[k] = scene->bmin[k];
    max_out[k] = scene->bmax[k];
  }
  return M3D_OK;
}

int32_t m3d_scene_get_info(const m3d_scene *scene, m3d_mesh_info *info) {
  if (!scene || !info) return fail(M3D_ERR_INVALID_ARG, "m3d_scene_get_info: NULL argument");
  *info = scene->info;
  info->device_bytes += (int64_t)(scene->shape_nodes.bytes + scene->shape_tris.bytes + scene->instances.bytes +
                                  scene->shapes.bytes);
  return M3D_OK;
}

int32_t m3d_scene_cast(m3d_scene *scene, const float *org, const float *dir, int64_t n, float *t,
                       int32_t *obj, int32_t *prim, float *normal, uint32_t flags, m3d_stats *stats) {
  if (!scene || n < 0 || (n > 0 && (!org || !dir))) return fail(M3D_ERR_INVALID_ARG, "m3d_scene_cast: bad arguments");
  if (n > (int64_t)0x7ff00000) return fail(M3D_ERR_INVALID_ARG, "batch too large; split it");
  m3d_ctx *ctx = scene->ctx;
  M3D_LOCK(ctx);
  M3D_CUDA(cudaSetDevice(ctx->device));
  if (stats) std::memset(stats, 0, sizeof(*stats));
  if (n == 0) return M3D_OK;
  cudaStream_t s = ctx->stream;
  // array stride: a multiple of four rays, so that the float4 arrays behind the n*3 float arrays
  // stay 16-byte aligned for any n (a batch of one is what Object.Cast maps to)
  const size_t per = ((size_t)n + 3) & ~(size_t)3;
  M3D_CUDA(ctx->scratch[1].reserve(per * (6 * sizeof(float) + 4 * sizeof(float4) + 6 * sizeof(float))));
  char *buf = ctx->scratch[1].as<char>();
  float *d_org3 = (float *)buf;
  float *d_dir3 = d_org3 + 3 * per;
  float4 *d_org4 = (float4 *)(d_dir3 + 3 * per);
  float4 *d_dir4 = d_org4 + per, *d_hit0 = d_dir4 + per, *d_hit1 = d_hit0 + per;
  float *d_t = (float *)(d_hit1 + per);
  int32_t *d_prim = (int32_t *)(d_t + per), *d_obj = d_prim + per;
  float *d_normal = (float *)(d_obj + per);
  M3D_CUDA(cudaMemcpyAsync(d_org3, org, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
  M3D_CUDA(cudaMemcpyAsync(d_dir3, dir, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
  launch_pack_rays(d_org3, d_dir3, n, 0.f, INFINITY, d_org4, d_dir4, s);
  SceneTraceLaunch p;
  p.t.org_tmin = d_org4;
  p.t.dir_tmax = d_dir4;
  p.t.n = n;
  p.t.hit0 = d_hit0;
  p.t.hit1 = d_hit1;
  p.t.refine = !(flags & M3D_TRACE_NO_REFINE);
  p.t.counters = (flags & M3D_TRACE_COUNTERS) ? stats_counters(ctx, s) : nullptr;
  p.t.ray_counter = next_work_counter(ctx);
  if (!p.t.ray_counter) return fail(M3D_ERR_OOM, "work counter allocation failed");
  GpuTimer tm;
  tm.start(s);
  launch_trace_scene(scene->dev, p, s);
  tm.stop(s);
  launch_unpack_hits(d_hit0, d_hit1, n, t ? d_t : nullptr, prim ? d_prim : nullptr, obj ? d_obj : nullptr,
                     normal ? d_normal : nullptr, nullptr, s);
  if (t) M3D_CUDA(cudaMemcpyAsync(t, d_t, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (prim) M3D_CUDA(cudaMemcpyAsync(prim, d_prim, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  if (obj) M3D_CUDA(cudaMemcpyAsync(obj, d_obj, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  if (normal) M3D_CUDA(cudaMemcpyAsync(normal, d_normal, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
  M3D_CUDA(cudaStreamSynchronize(s));
  M3D_CUDA(cudaGetLastError());
  if (stats) {
    stats->rays = n;
    stats->kernel_ms = tm.ms();
    stats->launches = 4;
    if (p.t.counters) {
      unsigned long long c[2];
      M3D_CUDA(cudaMemcpy(c, p.t.counters, sizeof(c), cudaMemcpyDeviceToHost));
      stats->nodes_visited = (int64_t)c[0];
      stats->tris_tested = (int64_t)c[1];
    }
  }
  return M3D_OK;
}

// sync = false: the caller synchronises the stream itself before it returns (the host-buffer call
// waits once, after its device-to-host copy, instead of twice per frame)
static int32_t raycast_one_device(m3d_scene *scene, const m3d_camera *cam, const m3d_point_light *lights,
                                  int32_t num_lights, int32_t width, int32_t height,
                                  const m3d_partition *part, void *d_rgb, void *stream, m3d_stats *stats,
                                  bool sync = true);

int32_t m3d_render_raycast_device(m3d_scene *scene, const m3d_camera *cam, const m3d_point_light *lights,
                                  int32_t num_lights, int32_t width, int32_t height,
                                  const m3d_partition *part, void *d_rgb, void *stream, m3d_stats *stats) {
  if (!scene || !cam || width <= 0 || height <= 0 || !d_rgb || num_lights < 0 || (num_lights > 0 && !lights))
    return fail(M3D_ERR_INVALID_ARG, "m3d_render_raycast: bad arguments");
  M3D_LOCK(scene->ctx);
  // Multi-device context: row bands, every device writes its band straight into the primary's
  // image over the NVLink peer mapping (disjoint rows: plain stores).  Small frames stay on the
  // primary: a 512 x 512 frame is 0.2 ms of GPU work, less than waking the other devices costs.
  const int g = 1 + (int)scene->replicas.size();
  if (g > 1 && (int64_t)width * height >= ((int64_t)1 << 21)) {
    int r0 = 0, r1 = height;
    if (part && !(part->row_begin == 0 && part->row_end == 0)) {
      r0 = part->row_begin;
      r1 = part->row_end;
      if (r0 < 0 || r1 > height || r0 > r1)
        return fail(M3D_ERR_INVALID_ARG, "bad row partition [%d,%d) of %d rows", r0, r1, height);
    }
    return render_sharded(scene, (cudaStream_t)stream, stats, [&](int i, m3d_scene *si, cudaStream_t s, m3d_stats *st) {
      int64_t b, e;
      split_range(r1 - r0, g, i, &b, &e);
      if (b == e) return (int32_t)M3D_OK;
      m3d_partition pi{};
      pi.row_begin = r0 + (int32_t)b;
      pi.row_end = r0 + (int32_t)e;
      return raycast_one_device(si, cam, lights, num_lights, width, height, &pi, d_rgb, s, st);
    });
  }
  return raycast_one_device(scene, cam, lights, num_lights, width, height, part, d_rgb, stream, stats);
}

static int32_t raycast_one_device(m3d_scene *scene, const m3d_camera *cam, const m3d_point_light *lights,
                                  int32_t num_lights, int32_t width, int32_t height,
                                  const m3d_partition *part, void *d_rgb, void *stream, m3d_stats *stats,
                                  bool sync) {
  m3d_ctx *ctx = scene->ctx;
  M3D_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  int row_begin = 0, row_end = height;
  if (part && !(part->row_begin == 0 && part->row_end == 0)) {
    row_begin = part->row_begin;
    row_end = part->row_end;
    if (row_begin < 0 || row_end > height || row_begin > row_end)
      return fail(M3D_ERR_INVALID_ARG, "bad row partition [%d,%d) of %d rows", row_begin, row_end, height);
  }
  const int64_t n = (int64_t)width * (row_end - row_begin);
  if (stats) std::memset(stats, 0, sizeof(*stats));
  if (n == 0) return M3D_OK;
  M3D_CUDA(ctx->scratch[2].reserve((size_t)n * 4 * sizeof(float4) + (size_t)(num_lights + 1) * sizeof(DevicePointLight)));
  float4 *d_org4 = ctx->scratch[2].as<float4>();
  float4 *d_dir4 = d_org4 + n, *d_hit0 = d_dir4 + n, *d_hit1 = d_hit0 + n;
  DevicePointLight *d_lights = (DevicePointLight *)(d_hit1 + n);
  // the light table is staged in context-owned host memory so that it outlives the async copy
  // without a synchronisation of its own (calls on a context are serialised by its lock, and every
  // call waits for the stream before it returns)
  std::vector<char> &stage = ctx->host_stage;
  stage.resize(std::max<size_t>(stage.size(), (size_t)(num_lights + 1) * sizeof(DevicePointLight)));
  DevicePointLight *hl = reinterpret_cast<DevicePointLight *>(stage.data());
  for (int i = 0; i < num_lights; i++) {
    for (int k = 0; k < 3; k++) {
      hl[i].origin[k] = (float)lights[i].origin[k];
      hl[i].color[k] = (float)lights[i].color[k];
    }
    hl[i].quad_dropoff = lights[i].quad_dropoff;
  }
  if (num_lights)
    M3D_CUDA(cudaMemcpyAsync(d_lights, hl, (size_t)num_lights * sizeof(DevicePointLight), cudaMemcpyHostToDevice, s));
  const DeviceCamera dc = make_device_camera(*cam, width, height);
  GpuTimer tm;
  tm.start(s);
  launch_raygen_camera(dc, width, row_begin, row_end, d_org4, d_dir4, s);
  SceneTraceLaunch p;
  p.t.org_tmin = d_org4;
  p.t.dir_tmax = d_dir4;
  p.t.n = n;
  p.t.hit0 = d_hit0;
  p.t.hit1 = d_hit1;
  p.t.refine = true;
  p.t.counters = nullptr;
  p.t.ray_counter = next_work_counter(ctx);
  if (!p.t.ray_counter) return fail(M3D_ERR_OOM, "work counter allocation failed");
  launch_trace_scene(scene->dev, p, s);
  launch_shade_raycast(scene->dev, dc, d_lights, num_lights, d_org4, d_dir4, d_hit0, d_hit1, n,
                       (float *)d_rgb + (size_t)row_begin * width * 3, s);
  tm.stop(s);
  if (sync || stats) M3D_CUDA(cudaStreamSynchronize(s));
  M3D_CUDA(cudaGetLastError());
  if (stats) {
    stats->rays = n;
    stats->kernel_ms = tm.ms();
    stats->launches = 4;
  }
  return M3D_OK;
}

int32_t m3d_render_raycast(m3d_scene *scene, const m3d_camera *cam, const m3d_point_light *lights,
                           int32_t num_lights, int32_t width, int32_t height, const m3d_partition *part,
                           float *rgb, m3d_stats *stats) {
  if (!scene || !rgb || width <= 0 || height <= 0) return fail(M3D_ERR_INVALID_ARG, "m3d_render_raycast: bad arguments");
  m3d_ctx *ctx = scene->ctx;
  M3D_LOCK(ctx);
  M3D_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)width * height * 3 * sizeof(float);
  M3D_CUDA(ctx->scratch[3].reserve(bytes));
  // pixels whose ray misses keep their previous value (raycast.go:26-28): start from the caller's image
  M3D_CUDA(cudaMemcpyAsync(ctx->scratch[3].p, rgb, bytes, cudaMemcpyHostToDevice, ctx->stream));
  // one device, or the multi-device dispatch for large frames (which waits for its members)
  const bool multi = !scene->replicas.empty() && (int64_t)width * height >= ((int64_t)1 << 21);
  int32_t rc = multi ? m3d_render_raycast_device(scene, cam, lights, num_lights, width, height, part,
                                                 ctx->scratch[3].p, ctx->stream, stats)
                     : raycast_one_device(scene, cam, lights, num_lights, width, height, part, ctx->scratch[3].p,
                                          ctx->stream, stats, /*sync=*/false);
  if (rc != M3D_OK) return rc;
  M3D_CUDA(cudaMemcpyAsync(rgb, ctx->scratch[3].p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  M3D_CUDA(cudaStreamSynchronize(ctx->stream));
  if (stats) {
    stats->h2d_bytes = (int64_t)bytes;
    stats->d2h_bytes = (int64_t)bytes;
  }
  return M3D_OK;
}

// Renders views [v0, v1) of a view batch on one device into d_out (view-major, downsampled frames).
static int32_t raycast_views_one_device(m3d_scene *scene, const m3d_camera *cams, int v0, int v1,
                                        const m3d_point_light *lights, const int32_t *light_begin, int32_t width,
                                        int32_t height, int32_t factor, float *rgb_host, m3d_stats *stats) {
  m3d_ctx *ctx = scene->ctx;
  M3D_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  if (stats) std::memset(stats, 0, sizeof(*stats));
  const int nv = v1 - v0;
  if (nv <= 0) return M3D_OK;
  const int64_t n = (int64_t)width * height;
  const int ow = width / factor, oh = height / factor;
  const size_t out_floats = (size_t)ow * oh * 3;
  const int nl_total = light_begin[v1] - light_begin[v0];
  // scratch: rays + hits of one frame, the full-resolution frame, all downsampled frames, all lights
  const size_t b_rays = (size_t)n * 4 * sizeof(float4), b_frame = (size_t)n * 3 * sizeof(float),
               b_out = out_floats * nv * sizeof(float), b_lights = (size_t)(nl_total + 1) * sizeof(DevicePointLight);
  M3D_CUDA(ctx->scratch[12].reserve(b_rays + b_frame + b_out + b_lights + 1024));
  char *base = ctx->scratch[12].as<char>();
  float4 *d_org4 = (float4 *)base, *d_dir4 = d_org4 + n, *d_hit0 = d_dir4 + n, *d_hit1 = d_hit0 + n;
  float *d_frame = (float *)(base + b_rays);
  float *d_out = (float *)(base + b_rays + b_frame);
  DevicePointLight *d_lights = (DevicePointLight *)(base + b_rays + b_frame + ((b_out + 255) & ~(size_t)255));
  std::vector<char> &stage = ctx->host_stage;
  stage.resize(std::max<size_t>(stage.size(), b_lights));
  DevicePointLight *hl = reinterpret_cast<DevicePointLight *>(stage.data());
  for (int i = 0; i < nl_total; i++) {
    const m3d_point_light &L = lights[light_begin[v0] + i];
    for (int k = 0; k < 3; k++) {
      hl[i].origin[k] = (float)L.origin[k];
      hl[i].color[k] = (float)L.color[k];
    }
    hl[i].quad_dropoff = L.quad_dropoff;
  }
  if (nl_total)
    M3D_CUDA(cudaMemcpyAsync(d_lights, hl, (size_t)nl_total * sizeof(DevicePointLight), cudaMemcpyHostToDevice, s));
  GpuTimer tm;
  tm.start(s);
  int64_t launches = 0;
  for (int v = v0; v < v1; v++) {
    const DeviceCamera dc = make_device_camera(cams[v], width, height);
    // a fresh image is black where rays miss (NewImage, image.go:24-31; raycast.go:26-28 leaves those pixels)
    M3D_CUDA(cudaMemsetAsync(d_frame, 0, b_frame, s));
    launch_raygen_camera(dc, width, 0, height, d_org4, d_dir4, s);
    SceneTraceLaunch p;
    p.t.org_tmin = d_org4;
    p.t.dir_tmax = d_dir4;
    p.t.n = n;
    p.t.hit0 = d_hit0;
    p.t.hit1 = d_hit1;
    p.t.refine = true;
    p.t.counters = nullptr;
    p.t.ray_counter = next_work_counter(ctx);
    if (!p.t.ray_counter) return fail(M3D_ERR_OOM, "work counter allocation failed");
    launch_trace_scene(scene->dev, p, s);
    launch_shade_raycast(scene->dev, dc, d_lights + (light_begin[v] - light_begin[v0]),
                         light_begin[v + 1] - light_begin[v], d_org4, d_dir4, d_hit0, d_hit1, n, d_frame, s);
    float *dst = d_out + out_floats * (size_t)(v - v0);
    if (factor > 1)
      launch_downsample_image(d_frame, width, height, factor, dst, s);
    else
      M3D_CUDA(cudaMemcpyAsync(dst, d_frame, b_frame, cudaMemcpyDeviceToDevice, s));
    launches += 5;
  }
  tm.stop(s);
  M3D_CUDA(cudaMemcpyAsync(rgb_host + out_floats * (size_t)v0, d_out, b_out, cudaMemcpyDeviceToHost, s));
  M3D_CUDA(cudaStreamSynchronize(s));
  M3D_CUDA(cudaGetLastError());
  if (stats) {
    stats->rays = n * nv;
    stats->kernel_ms = tm.ms();
    stats->launches = launches;
    stats->d2h_bytes = (int64_t)b_out;
  }
  return M3D_OK;
}

int32_t m3d_render_raycast_views(m3d_scene *scene, const m3d_camera *cams, int32_t num_views,
                                 const m3d_point_light *lights, const int32_t *light_begin, int32_t width,
                                 int32_t height, int32_t downsample, float *rgb, m3d_stats *stats) {
  if (!scene || !cams || num_views < 0 || !light_begin || width <= 0 || height <= 0 || downsample < 1 || !rgb)
    return fail(M3D_ERR_INVALID_ARG, "m3d_render_raycast_views: bad arguments");
  if (width % downsample || height % downsample)
    return fail(M3D_ERR_INVALID_ARG, "image size %d x %d cannot be divided evenly by factor %d", width, height,
                downsample);  // image.go:101-104
  for (int v = 0; v < num_views; v++)
    if (light_begin[v + 1] < light_begin[v] || (light_begin[v + 1] > light_begin[v] && !lights))
      return fail(M3D_ERR_INVALID_ARG, "m3d_render_raycast_views: bad light ranges");
  M3D_LOCK(scene->ctx);
  if (stats) std::memset(stats, 0, sizeof(*stats));
  if (num_views == 0) return M3D_OK;
  const int g = 1 + (int)scene->replicas.size();
  if (g > 1 && num_views > 1) {
    // the views are independent frames: spread them over the devices (no exchange at all)
    std::vector<m3d_stats> st((size_t)g);
    const int32_t rc = parallel_members(g, [&](int i) -> int32_t {
      int64_t b, e;
      split_range(num_views, g, i, &b, &e);
      m3d_scene *si = i == 0 ? scene : scene->replicas[(size_t)i - 1];
      std::lock_guard<std::recursive_mutex> lock(si->ctx->mu);
      return raycast_views_one_device(si, cams, (int)b, (int)e, lights, light_begin, width, height, downsample, rgb,
                                      &st[(size_t)i]);
    });
    if (stats)
      for (const m3d_stats &x : st) {
        stats->rays += x.rays;
        stats->kernel_ms = std::max(stats->kernel_ms, x.kernel_ms);
        stats->launches += x.launches;
        stats->d2h_bytes += x.d2h_bytes;
      }
    return rc;
  }
  return raycast_views_one_device(scene, cams, 0, num_views, lights, light_begin, width, height, downsample, rgb, stats);
}

int32_t m3d_finalize_image_device(m3d_ctx *ctx, const void *d_sum, int64_t num_pixels, double inv_samples,
                                  void *d_mean, void *d_srgb8, void *stream) {
  if (!ctx || !d_sum || num_pixels < 0) return fail(M3D_ERR_INVALID_ARG, "m3d_finalize_image_device: bad arguments");
  M3D_LOCK(ctx);
  M3D_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  launch_finalize_image((const float *)d_sum, num_pixels * 3, (float)inv_samples, (float *)d_mean,
                        (uint8_t *)d_srgb8, s);
  M3D_CUDA(cudaGetLastError());
  return M3D_OK;
}

}  // extern "C"
