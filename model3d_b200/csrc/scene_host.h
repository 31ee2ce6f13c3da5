// Host-side record behind the opaque m3d_scene handle (include/m3d.h): device buffers plus
// the host copies the renderers need at call time (area-light tables, material masks).
#pragma once
#include <vector>

#include "api_common.h"
#include "scene.h"

struct m3d_scene {
  m3d_ctx *ctx = nullptr;
  std::vector<m3d_scene *> replicas;  // multi-device context: copies on ctx->members[i] (owned)
  m3d::DevBuf nodes, tris, vnormals, shapes, objects, materials;
  m3d::DevBuf shape_nodes, shape_tris, instances;  // object-level hierarchy, mesh instances
  std::vector<m3d::DeviceInstance> host_instances;
  std::vector<m3d_mesh *> instance_meshes;    // per instance: the mesh it refers to (borrowed)
  m3d::DeviceScene dev;
  std::vector<m3d::DeviceShape> host_shapes;
  std::vector<m3d_material_desc> host_materials;
  std::vector<int32_t> object_material;
  std::vector<int32_t> object_kind;           // 0 mesh / ShapeKind
  std::vector<int64_t> object_tri_begin;      // for mesh objects: range in the merged input
  std::vector<int64_t> object_tri_count;
  std::vector<float> merged_tris;             // world-space triangles of all mesh objects
  std::vector<int32_t> leaf_of_merged;        // merged triangle index -> leaf-order index in the BVH
  double bmin[3] = {0, 0, 0}, bmax[3] = {0, 0, 0};
  m3d_mesh_info info{};
};

namespace m3d {
// Gives a scene built on a multi-device context its replicas (api_multi.cu).
int32_t replicate_scene(m3d_scene *scene);
// Runs one render call per device of the scene's multi-device context, side by side:
// fn(i, replica i, stream, stats) on one host thread each.  The primary renders on `stream` (or
// its context's stream), the others on their own; they first wait (on the device) for everything
// the caller enqueued on the primary's stream, e.g. the clearing of the shared accumulator.
// Every fn must return with its stream idle (the m3d_render_*_device bodies do), so when this
// returns every device's contribution has landed in the primary's memory.  stats (optional): summed.
int32_t render_sharded(m3d_scene *scene, cudaStream_t stream, m3d_stats *stats,
                       const std::function<int32_t(int, m3d_scene *, cudaStream_t, m3d_stats *)> &fn);
}  // namespace m3d
