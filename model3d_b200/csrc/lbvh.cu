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
