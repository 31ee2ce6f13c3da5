// Compressed 8-wide BVH: the HBM-resident acceleration structure that replaces the
// reference's pointer tree of JoinedCollider nodes (model3d/collisions.go:217-253,
// built by model3d/bvh.go:131-156 + collisions.go:169-179).
//
// Any BVH over the same triangle set yields the same first hit (except ties), so the
// hierarchy itself is free to differ from the reference's median split: here a
// binned-SAH binary tree is collapsed into 8-wide nodes whose child boxes are
// quantised to 8 bits on a per-node power-of-two grid (layout after Ylitie, Karras,
// Laine, "Efficient Incoherent Ray Traversal on GPUs Through Compressed Wide BVHs",
// HPG 2017).  80 bytes per node at a stride of M3D_NODE_BYTES (node_layout.h).
#pragma once
#include <cstdint>
#include <vector>

#include "node_layout.h"

namespace m3d {

struct alignas(M3D_NODE_BYTES == 96 ? 32 : 16) WideNode {
  float origin[3];     // node AABB min: quantisation origin
  uint8_t exp[3];      // per-axis biased float exponent of the grid step (2^(e-127))
  uint8_t imask;       // bit s set <=> slot s holds an internal child
  uint32_t child_base; // index of this node's first internal child (children contiguous)
  uint32_t tri_base;   // index of this node's first leaf triangle
  uint8_t meta[8];     // 0 empty | internal: 0b001sssss (sssss = 24+slot) | leaf: unary count<<5 | tri offset
  uint8_t qlo[3][8];   // quantised child box mins  [axis][slot]
  uint8_t qhi[3][8];   // quantised child box maxes [axis][slot]
#if M3D_NODE_BYTES == 96
  uint8_t pad[16];     // keeps every node on a 32-byte sector boundary (node_layout.h)
#endif
};
static_assert(sizeof(WideNode) == M3D_NODE_BYTES, "WideNode stride");

// One triangle record: three float4.  The w lanes carry ids so that the winning hit
// needs no second lookup: v0.w = bits(prim id in the caller's array),
// v1.w = bits(object id), v2.w = longest edge length (float, infinity norm; the scale
// of the float32 error bound of the triangle test).
struct alignas(16) TriRecord {
  float v0[3];
  int32_t prim;
  float v1[3];
  int32_t object;
  float v2[3];
  int32_t pad;  // float bits: longest edge
};
static_assert(sizeof(TriRecord) == 48, "TriRecord must be 48 bytes");

struct BuildInput {
  const float *tris = nullptr;      // n*9
  int64_t n = 0;
  const int32_t *prim_ids = nullptr; // optional n (default i)
  const int32_t *obj_ids = nullptr;  // optional n (default 0)
};

struct WideBVH {
  std::vector<WideNode> nodes;  // nodes[0] is the root
  std::vector<TriRecord> tris;  // leaf order
  float bounds_min[3] = {0, 0, 0}, bounds_max[3] = {0, 0, 0};
  int max_depth = 0;
  double sah_cost = 0;
  double build_ms = 0;
};

// Binary BVH node as produced by the device LBVH builder (lbvh.cu): leaf iff left < 0; a leaf
// owns order[first .. first+count), an internal node's range is the union of its children's.
struct BinaryNode {
  float mn[3], mx[3];
  int32_t left, right;
  int32_t first, count;
};

// Collapse a binary BVH (any builder) into the compressed 8-wide layout: cost-optimal collapse
// -> octant slot assignment -> quantisation -> leaf-ordered triangle records.
void build_wide_bvh_from_binary(const BuildInput &in, const BinaryNode *nodes, int64_t num_nodes, int32_t root,
                                const int32_t *order, WideBVH &out);

// Relative cost of one triangle test in the collapse (0.3; M3D_BVH_CPRIM overrides it for tuning).
double bvh_cost_prim();

// Host build: binned SAH binary tree -> cost-optimal 8-wide collapse -> octant slot
// assignment -> quantisation.  n == 0 yields a single empty node.
void build_wide_bvh(const BuildInput &in, WideBVH &out, int num_threads = 0);

}  // namespace m3d
