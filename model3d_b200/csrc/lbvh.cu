// Device LBVH build (M3D_MESH_BUILD_DEVICE_LBVH): the binary hierarchy over the triangles is
// built on the GPU from Morton codes, replacing the reference's recursive median split
// (model3d/bvh.go:131-255, O(N log N) with three sorts per level on one core):
//   lbvh_boxes_kernel    per-triangle AABB + scene bounds (float atomics on ordered ints)
//   lbvh_morton_kernel   63-bit Morton code of the AABB centre (21 bits per axis)
//   radix sort           cub::DeviceRadixSort (library plumbing)
//   lbvh_karras_kernel   binary radix tree, one thread per internal node
//                        (Karras, "Maximizing Parallelism in the Construction of BVHs,
//                        Octrees, and k-d Trees", HPG 2012)
//   lbvh_refit_kernel    bottom-up AABB refit, the second child to arrive carries on
// The binary tree is then collapsed into the compressed 8-wide layout by the same code path as
// the host SAH build (wide_bvh.cpp: build_wide_bvh_from_binary).  First hits do not depend on
// the hierarchy, so parity with the oracle is unaffected; tree quality (nodes visited per ray)
// is lower than the SAH build's.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <chrono>
#include <cstring>

#include "api_common.h"
#include "wide_bvh.h"

namespace m3d {

namespace {

__device__ __forceinline__ int float_to_ordered(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// boxes: 6 floats per triangle (mn, mx); bounds: 6 ordered ints (min xyz, max xyz)
__global__ void lbvh_boxes_kernel(const float *__restrict__ tris, int n, float *__restrict__ boxes,
                                  int *__restrict__ bounds) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  if (i < n) {
    const float *v = tris + (size_t)i * 9;
    for (int k = 0; k < 3; k++) {
      const float a = v[k], b = v[3 + k], c = v[6 + k];
      mn[k] = fminf(fminf(a, b), c);
      mx[k] = fmaxf(fmaxf(a, b), c);
      boxes[(size_t)i * 6 + k] = mn[k];
      boxes[(size_t)i * 6 + 3 + k] = mx[k];
    }
  }
  // warp reduce, one atomic per warp and component
  for (int k = 0; k < 3; k++) {
    float a = mn[k], b = mx[k];
    for (int off = 16; off > 0; off >>= 1) {
      a = fminf(a, __shfl_xor_sync(0xffffffffu, a, off));
      b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, off));
    }
    if ((threadIdx.x & 31) == 0 && a <= b) {
      atomicMin(bounds + k, float_to_ordered(a));
      atomicMax(bounds + 3 + k, float_to_ordered(b));
    }
  }
}

__device__ __forceinline__ unsigned long long spread21(unsigned long long x) {
  x &= 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}

__global__ void lbvh_morton_kernel(const float *__restrict__ boxes, int n, const int *__restrict__ bounds,
                                   unsigned long long *__restrict__ keys, int *__restrict__ ids) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long code = 0;
  for (int k = 0; k < 3; k++) {
    const float lo = ordered_to_float(bounds[k]), hi = ordered_to_float(bounds[3 + k]);
    const float c = 0.5f * (boxes[(size_t)i * 6 + k] + boxes[(size_t)i * 6 + 3 + k]);
    const float ext = hi - lo;
    float u = ext > 0.f ? (c - lo) / ext : 0.f;
    u = fminf(fmaxf(u, 0.f), 1.f);
    const unsigned long long q = (unsigned long long)fminf(u * 2097152.f, 2097151.f);
    code |= spread21(q) << (2 - k);
  }
  keys[i] = code;
  ids[i] = i;
}

// length of the common prefix of the (key, position) pairs at sorted positions i and j; -1 out of range
__device__ __forceinline__ int delta(const unsigned long long *__restrict__ keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  const unsigned long long a = keys[i], b = keys[j];
  if (a != b) return __clzll((long long)(a ^ b));
  return 64 + __clz(i ^ j);
}

// nodes: internal i in [0, n-1) at index i, leaf k at index n-1+k
__global__ void lbvh_karras_kernel(const unsigned long long *__restrict__ keys, int n, BinaryNode *__restrict__ nodes,
                                   int *__restrict__ parent) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  const int d = delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
  const int dmin = delta(keys, n, i, i - d);
  int lmax = 2;
  while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1)
    if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = delta(keys, n, i, j);
  int s = 0;
  for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
    if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    if (t == 1) break;
  }
  const int gamma = i + s * d + min(d, 0);
  const int lo = min(i, j), hi = max(i, j);
  const int left = lo == gamma ? n - 1 + gamma : gamma;
  const int right = hi == gamma + 1 ? n - 1 + gamma + 1 : gamma + 1;
  BinaryNode nd;
  nd.left = left;
  nd.right = right;
  nd.first = lo;
  nd.count = hi - lo + 1;
  for (int k = 0; k < 3; k++) {
    nd.mn[k] = INFINITY;
    nd.mx[k] = -INFINITY;
  }
  nodes[i] = nd;
  parent[left] = i;
  parent[right] = i;
  if (i == 0) parent[0] = -1;
}

__global__ void lbvh_refit_kernel(const float *__restrict__ boxes, const int *__restrict__ ids, int n,
                                  BinaryNode *__restrict__ nodes, const int *__restrict__ parent,
                                  int *__restrict__ visits) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int t = ids[k];
  BinaryNode leaf;
  for (int a = 0; a < 3; a++) {
    leaf.mn[a] = boxes[(size_t)t * 6 + a];
    leaf.mx[a] = boxes[(size_t)t * 6 + 3 + a];
  }
  leaf.left = leaf.right = -1;
  leaf.first = k;
  leaf.count = 1;
  const int self = n - 1 + k;
  nodes[self] = leaf;
  if (n == 1) return;
  int cur = parent[self];
  while (cur >= 0) {
    __threadfence();
    if (atomicAdd(visits + cur, 1) == 0) return;  // the sibling subtree is not finished yet
    // the sibling's box was written by another SM: read around the (non-coherent) L1
    const int l = nodes[cur].left, r = nodes[cur].right;
    for (int c = 0; c < 3; c++) {
      nodes[cur].mn[c] = fminf(__ldcg(&nodes[l].mn[c]), __ldcg(&nodes[r].mn[c]));
      nodes[cur].mx[c] = fmaxf(__ldcg(&nodes[l].mx[c]), __ldcg(&nodes[r].mx[c]));
    }
    cur = parent[cur];
  }
}

// ---- device collapse (M3D_MESH_BUILD_DEVICE_COLLAPSE) ---------------------------------------
// The cost-optimal collapse of wide_bvh.cpp (Collapse::run / children_of / finish_wide_bvh) on the
// device: a bottom-up pass computes, for every binary node and i = 1..7, the cheapest forest of at
// most i subtree roots (Ylitie et al. 2017, section 3.1), then the wide nodes are emitted level
// by level, one thread per wide node.

constexpr float kDevCostNode = 1.0f;
constexpr int kDevMaxLeaf = 3;

__device__ __forceinline__ float box_area(const BinaryNode &b) {
  const float dx = b.mx[0] - b.mn[0], dy = b.mx[1] - b.mn[1], dz = b.mx[2] - b.mn[2];
  return 2.f * (dx * dy + dy * dz + dz * dx);
}

// cost / dec: [node * 7 + (i - 1)]; children's tables are read around the non-coherent L1
__device__ void collapse_dp_node(const BinaryNode &nd, int node, float cprim, float *__restrict__ cost,
                                 uint8_t *__restrict__ dec) {
  const float area = box_area(nd);
  float *cn = cost + (size_t)node * 7;
  uint8_t *dn = dec + (size_t)node * 7;
  if (nd.left < 0) {
    for (int i = 0; i < 7; i++) {
      cn[i] = area * cprim;
      dn[i] = 0;
    }
    return;
  }
  float cl[7], cr[7];
  for (int i = 0; i < 7; i++) {
    cl[i] = __ldcg(cost + (size_t)nd.left * 7 + i);
    cr[i] = __ldcg(cost + (size_t)nd.right * 7 + i);
  }
  float dist[9];
  uint8_t distk[9];
  for (int j = 2; j <= 8; j++) {
    float best = INFINITY;
    uint8_t bk = 1;
    for (int k = 1; k < j; k++) {
      if (k > 7 || j - k > 7) continue;
      const float c = cl[k - 1] + cr[j - k - 1];
      if (c < best) {
        best = c;
        bk = (uint8_t)k;
      }
    }
    dist[j] = best;
    distk[j] = bk;
  }
  const float c_leaf = nd.count <= kDevMaxLeaf ? area * (float)nd.count * cprim : INFINITY;
  const float c_int = dist[8] + area * kDevCostNode;
  if (c_leaf <= c_int) {
    cn[0] = c_leaf;
    dn[0] = 0;
  } else {
    cn[0] = c_int;
    dn[0] = distk[8];
  }
  for (int i = 2; i <= 7; i++) {
    if (dist[i] < cn[i - 2]) {
      cn[i - 1] = dist[i];
      dn[i - 1] = distk[i];
    } else {
      cn[i - 1] = cn[i - 2];
      dn[i - 1] = 0;
    }
  }
}

// bottom-up like the refit (which has completed: boxes are final): the second child to arrive at a
// node computes its table
__global__ void lbvh_collapse_dp_kernel(int n, const BinaryNode *__restrict__ nodes, const int *__restrict__ parent,
                                        int *__restrict__ visits, float cprim, float *__restrict__ cost,
                                        uint8_t *__restrict__ dec) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int self = n - 1 + k;
  collapse_dp_node(nodes[self], self, cprim, cost, dec);
  if (n == 1) return;
  int cur = parent[self];
  while (cur >= 0) {
    __threadfence();
    if (atomicAdd(visits + cur, 1) == 0) return;
    collapse_dp_node(nodes[cur], cur, cprim, cost, dec);
    cur = parent[cur];
  }
}

__device__ __forceinline__ bool is_leaf_root(const BinaryNode *__restrict__ nodes, const uint8_t *__restrict__ dec, int c) {
  return nodes[c].left < 0 || dec[(size_t)c * 7] == 0;
}

// roots of the forest when subtree `node` may use up to `slots` slots (Collapse::collect)
__device__ int collect_children(const BinaryNode *__restrict__ nodes, const uint8_t *__restrict__ dec, int node,
                                int slots, int *out, int cnt) {
  int st_n[16], st_i[16], sp = 0;
  st_n[sp] = node;
  st_i[sp++] = slots;
  while (sp > 0) {
    const int n0 = st_n[--sp];
    int i = st_i[sp];
    const BinaryNode &nd = nodes[n0];
    uint8_t k = 0;
    while (nd.left >= 0 && i > 1 && (k = dec[(size_t)n0 * 7 + (i - 1)]) == 0) i--;
    if (nd.left < 0 || i == 1) {
      out[cnt++] = n0;
      continue;
    }
    // right first so that the left subtree is expanded first (the host's order)
    st_n[sp] = nd.right;
    st_i[sp++] = i - k;
    st_n[sp] = nd.left;
    st_i[sp++] = k;
  }
  return cnt;
}

__device__ __forceinline__ uint8_t dev_exp_byte(double extent) {
  if (!(extent > 0)) return 1;
  int e = (int)ceil(log2(extent / 255.0));
  while (ldexp(255.0, e) < extent) e++;
  while (e > -126 && ldexp(255.0, e - 1) >= extent) e--;
  int byte = e + 127;
  if (byte < 1) byte = 1;
  if (byte > 254) byte = 254;
  return (uint8_t)byte;
}

// One thread per wide node of the current level [begin, end).  b2_of_wide[w] is the binary node a
// wide node stands for; internal children get consecutive wide indices from *node_counter and
// their binary nodes are recorded for the next level.
__global__ void lbvh_emit_level_kernel(int begin, int end, const BinaryNode *__restrict__ nodes,
                                       const uint8_t *__restrict__ dec, const int *__restrict__ sorted_ids,
                                       const float *__restrict__ tris, const int32_t *__restrict__ prim_ids,
                                       const int32_t *__restrict__ obj_ids, int *__restrict__ b2_of_wide,
                                       WideNode *__restrict__ wide, TriRecord *__restrict__ wtris,
                                       unsigned int *__restrict__ node_counter, unsigned int *__restrict__ tri_counter) {
  const int w = begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= end) return;
  const int bn = b2_of_wide[w];
  const BinaryNode nb = nodes[bn];
  int kids[8];
  int k = 0;
  if (nb.left < 0 || dec[(size_t)bn * 7] == 0) {
    kids[k++] = bn;  // a <= 3-triangle subtree as the root
  } else {
    const int kl = dec[(size_t)bn * 7];
    k = collect_children(nodes, dec, nb.left, kl, kids, k);
    k = collect_children(nodes, dec, nb.right, 8 - kl, kids, k);
  }
  // octant slot assignment: greedy minimum of the signed distance along the slot's diagonal
  int slot_of[8];
  {
    float cost[8][8];
    for (int c = 0; c < k; c++) {
      const BinaryNode &cb = nodes[kids[c]];
      float d[3];
      for (int a = 0; a < 3; a++) d[a] = 0.5f * (cb.mn[a] + cb.mx[a]) - 0.5f * (nb.mn[a] + nb.mx[a]);
      for (int s = 0; s < 8; s++)
        cost[c][s] = ((s & 4) ? -d[0] : d[0]) + ((s & 2) ? -d[1] : d[1]) + ((s & 1) ? -d[2] : d[2]);
    }
    unsigned cu = 0, su = 0;
    for (int it = 0; it < k; it++) {
      float best = INFINITY;
      int bc = 0, bs = 0;
      for (int c = 0; c < k; c++)
        if (!((cu >> c) & 1u))
          for (int s = 0; s < 8; s++)
            if (!((su >> s) & 1u) && cost[c][s] < best) {
              best = cost[c][s];
              bc = c;
              bs = s;
            }
      if (best == INFINITY) {  // NaN boxes: any free pair
        for (int c = 0; c < k; c++) if (!((cu >> c) & 1u)) bc = c;
        for (int s = 0; s < 8; s++) if (!((su >> s) & 1u)) bs = s;
      }
      cu |= 1u << bc;
      su |= 1u << bs;
      slot_of[bc] = bs;
    }
  }
  int child_in_slot[8];
  for (int s = 0; s < 8; s++) child_in_slot[s] = -1;
  for (int c = 0; c < k; c++) child_in_slot[slot_of[c]] = c;
  unsigned n_internal = 0, n_tris = 0;
  for (int c = 0; c < k; c++) {
    if (is_leaf_root(nodes, dec, kids[c]))
      n_tris += (unsigned)nodes[kids[c]].count;
    else
      n_internal++;
  }
  const unsigned child_base = n_internal ? atomicAdd(node_counter, n_internal) : 0u;
  const unsigned tri_base = n_tris ? atomicAdd(tri_counter, n_tris) : 0u;

  WideNode nd;
  memset(&nd, 0, sizeof(nd));
  double step[3];
  for (int a = 0; a < 3; a++) {
    nd.origin[a] = nb.mn[a];
    nd.exp[a] = dev_exp_byte((double)nb.mx[a] - (double)nb.mn[a]);
    step[a] = ldexp(1.0, (int)nd.exp[a] - 127);
  }
  nd.child_base = child_base;
  nd.tri_base = tri_base;
  unsigned rank = 0, tri_off = 0;
  for (int s = 0; s < 8; s++) {
    const int c = child_in_slot[s];
    if (c < 0) {
      for (int a = 0; a < 3; a++) {
        nd.qlo[a][s] = 255;
        nd.qhi[a][s] = 0;
      }
      continue;
    }
    const int cn = kids[c];
    const BinaryNode &cnode = nodes[cn];
    for (int a = 0; a < 3; a++) {
      const double lo = ((double)cnode.mn[a] - (double)nb.mn[a]) / step[a];
      const double hi = ((double)cnode.mx[a] - (double)nb.mn[a]) / step[a];
      nd.qlo[a][s] = (uint8_t)fmin(255.0, fmax(0.0, floor(lo)));
      nd.qhi[a][s] = (uint8_t)fmin(255.0, fmax(0.0, ceil(hi)));
    }
    if (is_leaf_root(nodes, dec, cn)) {
      const int cnt = cnode.count;
      const unsigned unary = cnt == 1 ? 1u : (cnt == 2 ? 3u : 7u);
      nd.meta[s] = (uint8_t)((unary << 5) | tri_off);
      for (int i = 0; i < cnt; i++) {
        const int t = sorted_ids[cnode.first + i];
        const float *v = tris + (size_t)t * 9;
        TriRecord tr;
        for (int q = 0; q < 3; q++) {
          tr.v0[q] = v[q];
          tr.v1[q] = v[3 + q];
          tr.v2[q] = v[6 + q];
        }
        tr.prim = prim_ids ? prim_ids[t] : t;
        tr.object = obj_ids ? obj_ids[t] : 0;
        float em = 0.f;
        for (int q = 0; q < 3; q++) {
          em = fmaxf(em, fabsf(v[3 + q] - v[q]));
          em = fmaxf(em, fabsf(v[6 + q] - v[q]));
          em = fmaxf(em, fabsf(v[6 + q] - v[3 + q]));
        }
        em *= 1.0000002f;
        tr.pad = __float_as_int(em);
        wtris[tri_base + tri_off + i] = tr;
      }
      tri_off += (unsigned)cnt;
    } else {
      nd.imask |= (uint8_t)(1u << s);
      nd.meta[s] = (uint8_t)((1u << 5) | (24 + s));
      b2_of_wide[child_base + rank] = cn;
      rank++;
    }
  }
  wide[w] = nd;
}

}  // namespace

// tris: host n*9 floats.  Fills nodes (2n-1 entries, root = 0 for n >= 2; n == 1: the single
// leaf) and order (sorted position -> triangle index).
int32_t lbvh_build_binary(m3d_ctx *ctx, const float *tris, int64_t n64, std::vector<BinaryNode> &nodes,
                          std::vector<int32_t> &order, int32_t *root_out, double *device_ms) {
  const int n = (int)n64;
  cudaStream_t s = ctx->stream;
  const size_t b_tris = (size_t)n * 9 * 4, b_boxes = (size_t)n * 6 * 4, b_keys = (size_t)n * 8, b_ids = (size_t)n * 4;
  const size_t n_nodes = (size_t)2 * n - 1;
  size_t temp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                  (int *)nullptr, (int *)nullptr, n, 0, 63, s);
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += al(bytes);
    return o;
  };
  const size_t o_tris = take(b_tris), o_boxes = take(b_boxes), o_k0 = take(b_keys), o_k1 = take(b_keys),
               o_i0 = take(b_ids), o_i1 = take(b_ids), o_nodes = take(n_nodes * sizeof(BinaryNode)),
               o_parent = take(n_nodes * 4), o_visits = take((size_t)n * 4), o_bounds = take(64), o_temp = take(temp_bytes);
  M3D_CUDA(ctx->scratch[10].reserve(off));
  char *p = ctx->scratch[10].as<char>();
  float *d_tris = (float *)(p + o_tris), *d_boxes = (float *)(p + o_boxes);
  unsigned long long *d_k0 = (unsigned long long *)(p + o_k0), *d_k1 = (unsigned long long *)(p + o_k1);
  int *d_i0 = (int *)(p + o_i0), *d_i1 = (int *)(p + o_i1);
  BinaryNode *d_nodes = (BinaryNode *)(p + o_nodes);
  int *d_parent = (int *)(p + o_parent), *d_visits = (int *)(p + o_visits), *d_bounds = (int *)(p + o_bounds);
  M3D_CUDA(cudaMemcpyAsync(d_tris, tris, b_tris, cudaMemcpyHostToDevice, s));
  const int h_bounds[6] = {0x7f800000, 0x7f800000, 0x7f800000, (int)0x807fffff, (int)0x807fffff, (int)0x807fffff};
  // ordered encodings of +inf (min slots) and -inf (max slots)
  M3D_CUDA(cudaMemcpyAsync(d_bounds, h_bounds, sizeof(h_bounds), cudaMemcpyHostToDevice, s));
  M3D_CUDA(cudaMemsetAsync(d_visits, 0, (size_t)n * 4, s));
  GpuTimer tm;
  tm.start(s);
  const unsigned blocks = (unsigned)((n + 255) / 256);
  lbvh_boxes_kernel<<<blocks, 256, 0, s>>>(d_tris, n, d_boxes, d_bounds);
  lbvh_morton_kernel<<<blocks, 256, 0, s>>>(d_boxes, n, d_bounds, d_k0, d_i0);
  cub::DeviceRadixSort::SortPairs(p + o_temp, temp_bytes, d_k0, d_k1, d_i0, d_i1, n, 0, 63, s);
  if (n > 1) lbvh_karras_kernel<<<blocks, 256, 0, s>>>(d_k1, n, d_nodes, d_parent);
  lbvh_refit_kernel<<<blocks, 256, 0, s>>>(d_boxes, d_i1, n, d_nodes, d_parent, d_visits);
  tm.stop(s);
  nodes.resize(n_nodes);
  order.resize((size_t)n);
  M3D_CUDA(cudaMemcpyAsync(nodes.data(), d_nodes, n_nodes * sizeof(BinaryNode), cudaMemcpyDeviceToHost, s));
  M3D_CUDA(cudaMemcpyAsync(order.data(), d_i1, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
  M3D_CUDA(cudaStreamSynchronize(s));
  M3D_CUDA(cudaGetLastError());
  *root_out = 0;
  if (device_ms) *device_ms = tm.ms();
  return M3D_OK;
}

}  // namespace m3d

namespace m3d {

// Full device build: Morton / radix sort / Karras / refit, then the cost-optimal 8-wide collapse
// and the emission of quantised nodes and leaf-ordered triangle records on the device.  The
// finished arrays are downloaded into `out` (the shared upload path re-sends them; keeping them
// resident is a later optimisation).  Node and triangle ranges are handed out with atomics, so
// the memory order of the nodes of one level may differ from run to run; slots, child order and
// therefore traversal order and hits do not.
// n*9 per-corner normals in caller (prim id) order -> 3 x float4 per triangle in leaf order
// (what upload_bvh does on the host for the other builders)
__global__ void lbvh_gather_vnormals_kernel(const TriRecord *__restrict__ wtris, const float *__restrict__ vn_by_prim,
                                            int n, float4 *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float *src = vn_by_prim + (size_t)wtris[i].prim * 9;
  for (int k = 0; k < 3; k++) out[(size_t)i * 3 + k] = make_float4(src[3 * k], src[3 * k + 1], src[3 * k + 2], 0.f);
}

// resident != nullptr: the finished node / triangle arrays (and the leaf-ordered vertex normals) stay on
// the device -- device-to-device copies into the mesh's own buffers -- and `out` only carries the
// counts, bounds, depth and cost; the 56 MB round trip through the host that the shared upload path
// costs (22 of 46 ms per million triangles) is gone.
int32_t lbvh_build_wide(m3d_ctx *ctx, const BuildInput &in, double cost_prim_value, WideBVH &out,
                        ResidentBVH *resident) {
  const int n = (int)in.n;
  cudaStream_t s = ctx->stream;
  const size_t b_tris = (size_t)n * 9 * 4, b_boxes = (size_t)n * 6 * 4, b_keys = (size_t)n * 8, b_ids = (size_t)n * 4;
  const size_t n_nodes = (size_t)2 * n - 1;
  size_t temp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                  (int *)nullptr, (int *)nullptr, n, 0, 63, s);
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += al(bytes);
    return o;
  };
  const size_t o_tris = take(b_tris), o_boxes = take(b_boxes), o_k0 = take(b_keys), o_k1 = take(b_keys),
               o_i0 = take(b_ids), o_i1 = take(b_ids), o_nodes = take(n_nodes * sizeof(BinaryNode)),
               o_parent = take(n_nodes * 4), o_visits = take((size_t)n * 4), o_bounds = take(64),
               o_temp = take(temp_bytes), o_cost = take(n_nodes * 7 * 4), o_dec = take(n_nodes * 7),
               o_prim = take(in.prim_ids ? b_ids : 0), o_obj = take(in.obj_ids ? b_ids : 0),
               o_b2 = take((size_t)(n + 1) * 4), o_wide = take((size_t)(n + 1) * sizeof(WideNode)),
               o_wtris = take((size_t)n * sizeof(TriRecord)), o_cnt = take(64);
  M3D_CUDA(ctx->scratch[10].reserve(off));
  char *p = ctx->scratch[10].as<char>();
  float *d_tris = (float *)(p + o_tris), *d_boxes = (float *)(p + o_boxes);
  unsigned long long *d_k0 = (unsigned long long *)(p + o_k0), *d_k1 = (unsigned long long *)(p + o_k1);
  int *d_i0 = (int *)(p + o_i0), *d_i1 = (int *)(p + o_i1);
  BinaryNode *d_nodes = (BinaryNode *)(p + o_nodes);
  int *d_parent = (int *)(p + o_parent), *d_visits = (int *)(p + o_visits), *d_bounds = (int *)(p + o_bounds);
  float *d_cost = (float *)(p + o_cost);
  uint8_t *d_dec = (uint8_t *)(p + o_dec);
  int32_t *d_prim = in.prim_ids ? (int32_t *)(p + o_prim) : nullptr;
  int32_t *d_obj = in.obj_ids ? (int32_t *)(p + o_obj) : nullptr;
  int *d_b2 = (int *)(p + o_b2);
  WideNode *d_wide = (WideNode *)(p + o_wide);
  TriRecord *d_wtris = (TriRecord *)(p + o_wtris);
  unsigned int *d_cnt = (unsigned int *)(p + o_cnt);  // [0] wide nodes, [1] triangles
  auto t0 = std::chrono::steady_clock::now();
  M3D_CUDA(cudaMemcpyAsync(d_tris, in.tris, b_tris, cudaMemcpyHostToDevice, s));
  if (d_prim) M3D_CUDA(cudaMemcpyAsync(d_prim, in.prim_ids, b_ids, cudaMemcpyHostToDevice, s));
  if (d_obj) M3D_CUDA(cudaMemcpyAsync(d_obj, in.obj_ids, b_ids, cudaMemcpyHostToDevice, s));
  const int h_bounds[6] = {0x7f800000, 0x7f800000, 0x7f800000, (int)0x807fffff, (int)0x807fffff, (int)0x807fffff};
  M3D_CUDA(cudaMemcpyAsync(d_bounds, h_bounds, sizeof(h_bounds), cudaMemcpyHostToDevice, s));
  M3D_CUDA(cudaMemsetAsync(d_visits, 0, (size_t)n * 4, s));
  const unsigned h_cnt[2] = {1u, 0u};  // wide node 0 is the root
  M3D_CUDA(cudaMemcpyAsync(d_cnt, h_cnt, sizeof(h_cnt), cudaMemcpyHostToDevice, s));
  const int root = 0;  // Karras: internal node 0 (n >= 2); n == 1: the single leaf has index n-1 == 0
  M3D_CUDA(cudaMemcpyAsync(d_b2, &root, 4, cudaMemcpyHostToDevice, s));
  const unsigned blocks = (unsigned)((n + 255) / 256);
  lbvh_boxes_kernel<<<blocks, 256, 0, s>>>(d_tris, n, d_boxes, d_bounds);
  lbvh_morton_kernel<<<blocks, 256, 0, s>>>(d_boxes, n, d_bounds, d_k0, d_i0);
  cub::DeviceRadixSort::SortPairs(p + o_temp, temp_bytes, d_k0, d_k1, d_i0, d_i1, n, 0, 63, s);
  if (n > 1) lbvh_karras_kernel<<<blocks, 256, 0, s>>>(d_k1, n, d_nodes, d_parent);
  lbvh_refit_kernel<<<blocks, 256, 0, s>>>(d_boxes, d_i1, n, d_nodes, d_parent, d_visits);
  M3D_CUDA(cudaMemsetAsync(d_visits, 0, (size_t)n * 4, s));
  lbvh_collapse_dp_kernel<<<blocks, 256, 0, s>>>(n, d_nodes, d_parent, d_visits, (float)cost_prim_value, d_cost, d_dec);
  // level-synchronous emission: the wide nodes allocated by one level are the work of the next
  int begin = 0, end = 1, depth = 0;
  while (begin < end) {
    depth++;
    lbvh_emit_level_kernel<<<(unsigned)((end - begin + 127) / 128), 128, 0, s>>>(
        begin, end, d_nodes, d_dec, d_i1, d_tris, d_prim, d_obj, d_b2, d_wide, d_wtris, d_cnt, d_cnt + 1);
    unsigned cnt[2] = {0, 0};
    M3D_CUDA(cudaMemcpyAsync(cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, s));
    M3D_CUDA(cudaStreamSynchronize(s));
    if ((int64_t)cnt[0] > (int64_t)n + 1 || (int64_t)cnt[1] > (int64_t)n)
      return fail(M3D_ERR_CUDA, "device collapse overran its buffers (%u nodes, %u triangles)", cnt[0], cnt[1]);
    begin = end;
    end = (int)cnt[0];
    if (depth > 64) return fail(M3D_ERR_CUDA, "device collapse did not terminate");
  }
  M3D_CUDA(cudaGetLastError());
  const int num_wide = end;
  BinaryNode root_node;
  float root_cost = 0;
  if (resident) {
    M3D_CUDA(resident->nodes->reserve((size_t)num_wide * sizeof(WideNode)));
    M3D_CUDA(resident->tris->reserve((size_t)n * sizeof(TriRecord)));
    M3D_CUDA(cudaMemcpyAsync(resident->nodes->p, d_wide, (size_t)num_wide * sizeof(WideNode), cudaMemcpyDeviceToDevice, s));
    M3D_CUDA(cudaMemcpyAsync(resident->tris->p, d_wtris, (size_t)n * sizeof(TriRecord), cudaMemcpyDeviceToDevice, s));
    if (resident->vnormals_by_prim) {
      // the build scratch is free again: stage the caller-order normals in it
      float *d_vn = (float *)(p + o_tris);  // n*9 floats, the size of the triangle upload
      M3D_CUDA(resident->vnormals->reserve((size_t)n * 3 * sizeof(float4)));
      M3D_CUDA(cudaMemcpyAsync(d_vn, resident->vnormals_by_prim, b_tris, cudaMemcpyHostToDevice, s));
      lbvh_gather_vnormals_kernel<<<blocks, 256, 0, s>>>(d_wtris, d_vn, n, resident->vnormals->as<float4>());
    }
    resident->num_nodes = num_wide;
    resident->num_tris = n;
  } else {
    out.nodes.resize((size_t)num_wide);
    out.tris.resize((size_t)n);
    M3D_CUDA(cudaMemcpyAsync(out.nodes.data(), d_wide, (size_t)num_wide * sizeof(WideNode), cudaMemcpyDeviceToHost, s));
    M3D_CUDA(cudaMemcpyAsync(out.tris.data(), d_wtris, (size_t)n * sizeof(TriRecord), cudaMemcpyDeviceToHost, s));
  }
  M3D_CUDA(cudaMemcpyAsync(&root_node, d_nodes, sizeof(BinaryNode), cudaMemcpyDeviceToHost, s));
  M3D_CUDA(cudaMemcpyAsync(&root_cost, d_cost, 4, cudaMemcpyDeviceToHost, s));
  M3D_CUDA(cudaStreamSynchronize(s));
  for (int k = 0; k < 3; k++) {
    out.bounds_min[k] = root_node.mn[k];
    out.bounds_max[k] = root_node.mx[k];
  }
  const double dx = (double)root_node.mx[0] - root_node.mn[0], dy = (double)root_node.mx[1] - root_node.mn[1],
               dz = (double)root_node.mx[2] - root_node.mn[2];
  out.sah_cost = root_cost / std::max(1e-300, 2.0 * (dx * dy + dy * dz + dz * dx));
  out.max_depth = depth;
  out.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return M3D_OK;
}

}  // namespace m3d
