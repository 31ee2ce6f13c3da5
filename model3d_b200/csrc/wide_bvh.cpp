// Host builder for the compressed 8-wide BVH (see wide_bvh.h).
// Replaces GroupTriangles + GroupedTrianglesToCollider of the reference
// (model3d/bvh.go:118-156, model3d/collisions.go:169-179); the hierarchy differs by
// design (SAH instead of median split), the set of triangles and their ids do not.
#include "wide_bvh.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <future>
#include <limits>
#include <thread>

namespace m3d {
namespace {

struct Box {
  float mn[3], mx[3];
  void reset() {
    for (int k = 0; k < 3; k++) {
      mn[k] = std::numeric_limits<float>::infinity();
      mx[k] = -std::numeric_limits<float>::infinity();
    }
  }
  void grow(const Box &b) {
    for (int k = 0; k < 3; k++) {
      mn[k] = std::min(mn[k], b.mn[k]);
      mx[k] = std::max(mx[k], b.mx[k]);
    }
  }
  void grow(const float *p) {
    for (int k = 0; k < 3; k++) {
      mn[k] = std::min(mn[k], p[k]);
      mx[k] = std::max(mx[k], p[k]);
    }
  }
  double area() const {
    double dx = (double)mx[0] - mn[0], dy = (double)mx[1] - mn[1], dz = (double)mx[2] - mn[2];
    if (dx < 0) return 0;
    return 2 * (dx * (dy + dz) + dy * dz);
  }
};

// binary-tree node shared with the device LBVH builder (wide_bvh.h BinaryNode has this layout)
struct B2Node {
  Box box;
  int32_t left = -1, right = -1;  // children (internal)
  int32_t first = 0, count = 0;   // range in idx; leaf iff left < 0
};
static_assert(sizeof(B2Node) == sizeof(BinaryNode), "B2Node must match BinaryNode");

struct Builder2 {
  const float *tris;
  int64_t n;
  std::vector<Box> tbox;
  std::vector<float> cent;  // n*3
  std::vector<int32_t> idx;
  std::vector<B2Node> nodes;
  std::atomic<int32_t> next_node{0};
  int par_threshold = 1 << 30;

  static constexpr int kBins = 16;

  int32_t alloc() { return next_node.fetch_add(1); }

  void build_range(int32_t node_id, int32_t first, int32_t count, int depth) {
    B2Node &nd = nodes[node_id];
    nd.first = first;
    nd.count = count;
    Box b;
    b.reset();
    Box cb;
    cb.reset();
    for (int32_t i = first; i < first + count; i++) {
      b.grow(tbox[idx[i]]);
      cb.grow(&cent[3 * (size_t)idx[i]]);
    }
    nd.box = b;
    if (count == 1) return;

    // binned SAH over the three axes
    int best_axis = -1, best_split = -1;
    double best_cost = std::numeric_limits<double>::infinity();
    for (int ax = 0; ax < 3; ax++) {
      float lo = cb.mn[ax], hi = cb.mx[ax];
      if (!(hi > lo)) continue;
      float scale = (float)kBins / (hi - lo);
      Box bb[kBins];
      int cnt[kBins];
      for (int i = 0; i < kBins; i++) {
        bb[i].reset();
        cnt[i] = 0;
      }
      for (int32_t i = first; i < first + count; i++) {
        int32_t t = idx[i];
        int bi = std::min(kBins - 1, std::max(0, (int)((cent[3 * (size_t)t + ax] - lo) * scale)));
        bb[bi].grow(tbox[t]);
        cnt[bi]++;
      }
      double right_area[kBins];
      int right_cnt[kBins];
      Box acc;
      acc.reset();
      int c = 0;
      for (int i = kBins - 1; i > 0; i--) {
        acc.grow(bb[i]);
        c += cnt[i];
        right_area[i] = acc.area();
        right_cnt[i] = c;
      }
      acc.reset();
      c = 0;
      for (int i = 0; i < kBins - 1; i++) {
        acc.grow(bb[i]);
        c += cnt[i];
        if (c == 0 || right_cnt[i + 1] == 0) continue;
        double cost = acc.area() * c + right_area[i + 1] * right_cnt[i + 1];
        if (cost < best_cost) {
          best_cost = cost;
          best_axis = ax;
          best_split = i;
        }
      }
    }
    int32_t mid;
    if (best_axis < 0) {
      mid = first + count / 2;  // all centroids coincide: split the index range
    } else {
      float lo = cb.mn[best_axis], hi = cb.mx[best_axis];
      float scale = (float)kBins / (hi - lo);
      auto it = std::partition(idx.begin() + first, idx.begin() + first + count, [&](int32_t t) {
        int bi = std::min(kBins - 1, std::max(0, (int)((cent[3 * (size_t)t + best_axis] - lo) * scale)));
        return bi <= best_split;
      });
      mid = (int32_t)(it - idx.begin());
      if (mid == first || mid == first + count) mid = first + count / 2;
    }
    int32_t l = alloc(), r = alloc();
    nodes[node_id].left = l;
    nodes[node_id].right = r;
    int32_t lc = mid - first, rc = first + count - mid;
    if (count >= par_threshold && depth < 6) {
      auto fut = std::async(std::launch::async, [=] { build_range(l, first, lc, depth + 1); });
      build_range(r, mid, rc, depth + 1);
      fut.get();
    } else {
      build_range(l, first, lc, depth + 1);
      build_range(r, mid, rc, depth + 1);
    }
  }
};

constexpr double kCostNode = 1.0;
// relative cost of one triangle test; M3D_BVH_CPRIM overrides it for tuning runs
static double cost_prim() {
  static const double v = [] {
    const char *e = getenv("M3D_BVH_CPRIM");
    const double x = e ? atof(e) : 0.3;
    return x > 0 ? x : 0.3;
  }();
  return v;
}
#define kCostPrim (cost_prim())
constexpr int kMaxLeaf = 3;
constexpr double kInf = std::numeric_limits<double>::infinity();

struct Collapse {
  const Builder2 &b2;
  std::vector<double> cost;    // [node*7 + (i-1)]
  std::vector<uint8_t> dec;    // [node*7 + (i-1)]: i==1: 0 leaf / k8 (1..7) internal; i>1: 0 reduce / k
  explicit Collapse(const Builder2 &b) : b2(b) {}

  void run(int32_t root) {
    size_t nn = (size_t)b2.next_node.load();
    cost.assign(nn * 7, kInf);
    dec.assign(nn * 7, 0);
    // iterative post-order
    std::vector<std::pair<int32_t, int>> st;
    st.push_back({root, 0});
    while (!st.empty()) {
      auto [n, phase] = st.back();
      const B2Node &nd = b2.nodes[n];
      if (nd.left < 0) {
        double c = nd.box.area() * kCostPrim;
        for (int i = 0; i < 7; i++) cost[(size_t)n * 7 + i] = c;
        st.pop_back();
        continue;
      }
      if (phase == 0) {
        st.back().second = 1;
        st.push_back({nd.left, 0});
        st.push_back({nd.right, 0});
        continue;
      }
      st.pop_back();
      const double *cl = &cost[(size_t)nd.left * 7];
      const double *cr = &cost[(size_t)nd.right * 7];
      double *cn = &cost[(size_t)n * 7];
      uint8_t *dn = &dec[(size_t)n * 7];
      double area = nd.box.area();
      // distribute over j = 2..8
      double dist[9];
      uint8_t distk[9];
      for (int j = 2; j <= 8; j++) {
        double best = kInf;
        uint8_t bk = 1;
        for (int k = 1; k < j; k++) {
          if (k > 7 || j - k > 7) continue;
          double c = cl[k - 1] + cr[j - k - 1];
          if (c < best) {
            best = c;
            bk = (uint8_t)k;
          }
        }
        dist[j] = best;
        distk[j] = bk;
      }
      double c_leaf = nd.count <= kMaxLeaf ? area * nd.count * kCostPrim : kInf;
      double c_int = dist[8] + area * kCostNode;
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
  }

  // roots of the forest when subtree n may use up to i slots
  void collect(int32_t n, int i, std::vector<int32_t> &out) const {
    const B2Node &nd = b2.nodes[n];
    if (nd.left < 0 || i == 1) {
      out.push_back(n);
      return;
    }
    uint8_t k = dec[(size_t)n * 7 + (i - 1)];
    if (k == 0) {
      collect(n, i - 1, out);
    } else {
      collect(nd.left, k, out);
      collect(nd.right, i - k, out);
    }
  }
  bool is_leaf_root(int32_t n) const {
    const B2Node &nd = b2.nodes[n];
    return nd.left < 0 || dec[(size_t)n * 7] == 0;
  }
  void children_of(int32_t n, std::vector<int32_t> &out) const {
    const B2Node &nd = b2.nodes[n];
    out.clear();
    if (nd.left < 0) {
      out.push_back(n);
      return;
    }
    uint8_t k = dec[(size_t)n * 7];
    if (k == 0) {
      // a <=3-triangle subtree chosen as a leaf but needed as an internal node (root only)
      out.push_back(n);
      return;
    }
    collect(nd.left, k, out);
    collect(nd.right, 8 - k, out);
  }
};

inline uint8_t exp_byte_for(double extent) {
  if (!(extent > 0)) return 1;
  int e = (int)std::ceil(std::log2(extent / 255.0));
  // guard against log2 rounding: need 255 * 2^e >= extent
  while (std::ldexp(255.0, e) < extent) e++;
  while (e > -126 && std::ldexp(255.0, e - 1) >= extent) e--;
  int byte = e + 127;
  if (byte < 1) byte = 1;
  if (byte > 254) byte = 254;
  return (uint8_t)byte;
}

void finish_wide_bvh(const BuildInput &in, Builder2 &b2, int32_t root2, WideBVH &out);

}  // namespace

void build_wide_bvh(const BuildInput &in, WideBVH &out, int num_threads) {
  auto t0 = std::chrono::steady_clock::now();
  out.nodes.clear();
  out.tris.clear();
  out.max_depth = 0;
  out.sah_cost = 0;
  if (in.n <= 0) {
    WideNode root;
    std::memset(&root, 0, sizeof(root));
    root.exp[0] = root.exp[1] = root.exp[2] = 1;
    for (int a = 0; a < 3; a++)
      for (int s = 0; s < 8; s++) {
        root.qlo[a][s] = 255;
        root.qhi[a][s] = 0;
      }
    out.nodes.push_back(root);
    return;
  }

  Builder2 b2;
  b2.tris = in.tris;
  b2.n = in.n;
  b2.tbox.resize(in.n);
  b2.cent.resize(3 * (size_t)in.n);
  b2.idx.resize(in.n);
  for (int64_t i = 0; i < in.n; i++) {
    Box b;
    b.reset();
    for (int v = 0; v < 3; v++) b.grow(in.tris + i * 9 + v * 3);
    b2.tbox[i] = b;
    for (int k = 0; k < 3; k++) b2.cent[3 * i + k] = 0.5f * (b.mn[k] + b.mx[k]);
    b2.idx[i] = (int32_t)i;
  }
  b2.nodes.resize(2 * (size_t)in.n);
  if (num_threads <= 0) num_threads = (int)std::thread::hardware_concurrency();
  b2.par_threshold = num_threads > 1 ? 32768 : (1 << 30);
  int32_t root2 = b2.alloc();
  b2.build_range(root2, 0, (int32_t)in.n, 0);

  finish_wide_bvh(in, b2, root2, out);
  out.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

namespace {
// binary tree -> cost-optimal 8-wide collapse -> octant slots -> quantised nodes + leaf-ordered
// triangle records
void finish_wide_bvh(const BuildInput &in, Builder2 &b2, int32_t root2, WideBVH &out) {
  Collapse col(b2);
  col.run(root2);
  out.sah_cost = col.cost[(size_t)root2 * 7] / std::max(1e-300, b2.nodes[root2].box.area());

  for (int k = 0; k < 3; k++) {
    out.bounds_min[k] = b2.nodes[root2].box.mn[k];
    out.bounds_max[k] = b2.nodes[root2].box.mx[k];
  }

  // breadth-first emission so that a node's internal children are contiguous
  struct Work {
    int32_t b2node;
    uint32_t wide;
    int depth;
  };
  std::vector<Work> queue;
  out.nodes.reserve((size_t)in.n / 2 + 8);
  out.tris.reserve(in.n);
  out.nodes.emplace_back();
  queue.push_back({root2, 0u, 1});
  std::vector<int32_t> kids;
  for (size_t qi = 0; qi < queue.size(); qi++) {
    Work w = queue[qi];
    out.max_depth = std::max(out.max_depth, w.depth);
    col.children_of(w.b2node, kids);
    int k = (int)kids.size();
    const Box &nb = b2.nodes[w.b2node].box;

    // octant slot assignment (greedy minimum of signed distance along the slot's diagonal)
    int slot_of[8];
    {
      double cost[8][8];
      for (int c = 0; c < k; c++) {
        const Box &cb = b2.nodes[kids[c]].box;
        double d[3];
        for (int a = 0; a < 3; a++)
          d[a] = 0.5 * ((double)cb.mn[a] + cb.mx[a]) - 0.5 * ((double)nb.mn[a] + nb.mx[a]);
        for (int s = 0; s < 8; s++)
          cost[c][s] = ((s & 4) ? -d[0] : d[0]) + ((s & 2) ? -d[1] : d[1]) + ((s & 1) ? -d[2] : d[2]);
      }
      bool cu[8] = {false}, su[8] = {false};
      for (int it = 0; it < k; it++) {
        double best = kInf;
        int bc = -1, bs = -1;
        for (int c = 0; c < k; c++)
          if (!cu[c])
            for (int s = 0; s < 8; s++)
              if (!su[s] && cost[c][s] < best) {
                best = cost[c][s];
                bc = c;
                bs = s;
              }
        cu[bc] = true;
        su[bs] = true;
        slot_of[bc] = bs;
      }
    }
    int child_in_slot[8];
    for (int s = 0; s < 8; s++) child_in_slot[s] = -1;
    for (int c = 0; c < k; c++) child_in_slot[slot_of[c]] = c;

    WideNode nd;
    std::memset(&nd, 0, sizeof(nd));
    double step[3];
    for (int a = 0; a < 3; a++) {
      nd.origin[a] = nb.mn[a];
      nd.exp[a] = exp_byte_for((double)nb.mx[a] - (double)nb.mn[a]);
      step[a] = std::ldexp(1.0, (int)nd.exp[a] - 127);
    }
    nd.child_base = (uint32_t)out.nodes.size();
    nd.tri_base = (uint32_t)out.tris.size();
    uint32_t n_internal = 0, tri_off = 0;
    for (int s = 0; s < 8; s++) {
      int c = child_in_slot[s];
      if (c < 0) {
        for (int a = 0; a < 3; a++) {
          nd.qlo[a][s] = 255;
          nd.qhi[a][s] = 0;
        }
        continue;
      }
      int32_t cn = kids[c];
      const B2Node &cnode = b2.nodes[cn];
      for (int a = 0; a < 3; a++) {
        double lo = ((double)cnode.box.mn[a] - (double)nb.mn[a]) / step[a];
        double hi = ((double)cnode.box.mx[a] - (double)nb.mn[a]) / step[a];
        nd.qlo[a][s] = (uint8_t)std::min(255.0, std::max(0.0, std::floor(lo)));
        nd.qhi[a][s] = (uint8_t)std::min(255.0, std::max(0.0, std::ceil(hi)));
      }
      if (col.is_leaf_root(cn)) {
        int cnt = cnode.count;
        static const uint8_t unary[4] = {0, 1, 3, 7};
        nd.meta[s] = (uint8_t)((unary[cnt] << 5) | tri_off);
        for (int i = 0; i < cnt; i++) {
          int32_t t = b2.idx[cnode.first + i];
          TriRecord tr;
          const float *v = in.tris + (size_t)t * 9;
          std::memcpy(tr.v0, v, 12);
          std::memcpy(tr.v1, v + 3, 12);
          std::memcpy(tr.v2, v + 6, 12);
          tr.prim = in.prim_ids ? in.prim_ids[t] : t;
          tr.object = in.obj_ids ? in.obj_ids[t] : 0;
          {
            // longest edge (infinity norm), rounded up: scale of the float32 error bound
            float em = 0.f;
            for (int k = 0; k < 3; k++) {
              em = std::max(em, std::fabs(v[3 + k] - v[k]));
              em = std::max(em, std::fabs(v[6 + k] - v[k]));
              em = std::max(em, std::fabs(v[6 + k] - v[3 + k]));
            }
            em *= 1.0000002f;
            std::memcpy(&tr.pad, &em, 4);
          }
          out.tris.push_back(tr);
        }
        tri_off += cnt;
      } else {
        nd.imask |= (uint8_t)(1u << s);
        nd.meta[s] = (uint8_t)((1u << 5) | (24 + s));
        n_internal++;
      }
    }
    // reserve the contiguous block of internal children, in slot order
    uint32_t base = nd.child_base;
    out.nodes.resize(out.nodes.size() + n_internal);
    uint32_t rank = 0;
    for (int s = 0; s < 8; s++)
      if (nd.imask & (1u << s)) {
        queue.push_back({kids[child_in_slot[s]], base + rank, w.depth + 1});
        rank++;
      }
    out.nodes[w.wide] = nd;
  }
}
}  // namespace

double bvh_cost_prim() { return cost_prim(); }

void build_wide_bvh_from_binary(const BuildInput &in, const BinaryNode *nodes, int64_t num_nodes, int32_t root,
                                const int32_t *order, WideBVH &out) {
  auto t0 = std::chrono::steady_clock::now();
  out.nodes.clear();
  out.tris.clear();
  out.max_depth = 0;
  out.sah_cost = 0;
  Builder2 b2;
  b2.tris = in.tris;
  b2.n = in.n;
  b2.idx.assign(order, order + in.n);
  b2.nodes.resize((size_t)num_nodes);
  std::memcpy((void *)b2.nodes.data(), nodes, (size_t)num_nodes * sizeof(BinaryNode));
  b2.next_node.store((int32_t)num_nodes);
  finish_wide_bvh(in, b2, root, out);
  out.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace m3d
