// Host-side record behind the opaque m3d_scene handle (include/m3d.h): device buffers plus
// the host copies the renderers need at call time (area-light tables, material masks).
#pragma once
#include <vector>

#include "api_common.h"
#include "scene.h"

struct m3d_scene {
  m3d_ctx *ctx = nullptr;
  m3d::DevBuf nodes, tris, vnormals, shapes, objects, materials;
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
