// sm_100a kernels for batched first-hit queries:
//   trace_first_hit_kernel  == JoinedCollider.FirstRayCollision per ray
//                              (model3d/collisions.go:275-290), see trace_core.cuh
//   pack_rays / unpack_hits == marshalling between the host ABI's n*3 arrays
//                              (Ray{Origin,Direction}, collisions.go:12-15;
//                              RayCollision / TriangleCollision, collisions.go:19-46)
//                              and the device float4 SoA layout.
#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "scene.h"
#include "trace_core.cuh"

namespace m3d {

namespace {

#ifndef M3D_TRACE_BLOCK
#define M3D_TRACE_BLOCK 128
#endif
#ifndef M3D_RAY_BATCH
#define M3D_RAY_BATCH (32 * 6)
#endif
constexpr int kTraceBlock = M3D_TRACE_BLOCK;
constexpr int kWarpsPerBlock = kTraceBlock / 32;
#ifndef M3D_SMEM_STACK
#define M3D_SMEM_STACK 10
#endif
#ifndef M3D_OPAQUE_ONE
#define M3D_OPAQUE_ONE 1  // bits of 1.0f from the kernel parameters: PRMT selectors become immediates
#endif
#ifndef M3D_RELOAD_SKIP
#define M3D_RELOAD_SKIP 1
#endif
#ifndef M3D_TINY_MINB
#define M3D_TINY_MINB 6  // resident blocks per SM of the tiny-scene instantiation (two triangle rounds)
#endif
#ifndef M3D_TINY_TRI_ROUNDS
// triangle tests per trip of the tiny-scene instantiation; measured on C3 (trace stage of a 1024^2 x 64 spp
// frame): 2 rounds 16.7 ms, 3 17.3, 4 14.9 (the only count whose instantiation stays nearly spill-free),
// 5 16.1, 6 16.4, 8 17.1
#define M3D_TINY_TRI_ROUNDS 4
#endif
#ifndef M3D_CULL_DEFAULT
#define M3D_CULL_DEFAULT 0  // bounds cull in front of large plain mesh batches (see cull_rays_kernel)
#endif
#ifndef M3D_MASK_LUT
#define M3D_MASK_LUT 1  // child hit bits from a shared-memory table (ALU-pipe relief): C2 2.405 -> 2.347 ms
#endif
constexpr int kSmemStack = M3D_SMEM_STACK;  // stack entries per thread kept in shared memory
constexpr int kLocalStack = 64 - kSmemStack;  // overflow entries (local memory; untouched for sane trees)
constexpr int kRayBatch = M3D_RAY_BATCH;  // rays a warp claims per global atomic
constexpr int kTinySceneNodes = 32;       // at most this many wide nodes: several triangle rounds per trip
#ifndef M3D_PREFETCH_NEXT_NODE
#define M3D_PREFETCH_NEXT_NODE 0  // measured on B200: 2.58 -> 3.98 ms per 2^24 rays (L1 prefetches throttle the LSU)
#endif
constexpr bool kPrefetchNextNode = M3D_PREFETCH_NEXT_NODE != 0;
#ifndef M3D_PREFETCH_QUEUED_TRI
#define M3D_PREFETCH_QUEUED_TRI 0
#endif
constexpr bool kPrefetchQueuedTri = M3D_PREFETCH_QUEUED_TRI != 0;

// Persistent-warp traversal with dynamic ray fetch.
//
// Incoherent rays finish after very different numbers of node visits (C2: 70 % of the rays
// miss after one or two nodes, the rest visit ten or more), so a one-thread-per-ray launch
// runs its warps mostly empty (ncu, round 1: 3.98 of 32 lanes active per instruction).
// Here every lane that finishes its ray immediately claims the next one from the warp's
// batch (refilled from a global counter with one atomic per kRayBatch rays), and the loop
// body is "one node visit, then that node's triangles" for all lanes together.
// The trace kernel only writes the raw float32 hit (t, b1, b2, triangle index);
// finish_hits_kernel re-evaluates hits in float64 in a separate, fully coherent pass.
// HAS_SKIP = false (plain mesh batches: no per-ray surface to ignore) drops the skip-id load and
// compare and frees a register of the 80.
// HAS_LIST: the rays to trace are p.ray_list[0 .. n) (survivors of the bounds cull) instead of 0 .. n.
template <bool COUNT, int MIN_BLOCKS, int TRI_ROUNDS, bool HAS_SKIP, bool HAS_LIST = false>
__global__ void __launch_bounds__(kTraceBlock, (MIN_BLOCKS * 128) / kTraceBlock)
trace_first_hit_kernel(DeviceBVH bvh, TraceLaunch p, unsigned int *__restrict__ ray_counter) {
  __shared__ uint2 s_stack[kSmemStack][kTraceBlock];
  __shared__ float4 s_stage[3][kTraceBlock];  // per warp: 32 prepared rays (origin|tmin, dir|tmax, 1/dir|err)
  __shared__ int s_stage_idx[HAS_LIST ? kTraceBlock : 1];  // ... and their ray indices (HAS_LIST)
#if M3D_MASK_LUT
  __shared__ uint32_t s_lut[256 * 3];  // hit bits by child meta byte, one entry per 12 bytes (see intersect_node)
  for (int m = (int)threadIdx.x; m < 256; m += kTraceBlock) s_lut[3 * m] = (((uint32_t)m >> 5) & 7u) << (m & 31);
  __syncthreads();
  const uint32_t lut_saddr = (uint32_t)__cvta_generic_to_shared(s_lut);
#endif
  uint2 l_stack[kLocalStack];
  const unsigned lane = threadIdx.x & 31u;
  const uint4 *__restrict__ nodes = bvh.nodes;
  const float4 *__restrict__ tris = bvh.tris;
  const int n = p.n_ptr ? min(__ldg(p.n_ptr), (int)p.n) : (int)p.n;
  // Rays a warp claims per global atomic: kRayBatch for big launches; small launches (pipeline
  // chunks, late bounces of the wavefront renderers) use smaller batches so that every resident
  // warp gets work instead of a few warps walking 192 rays each.
  int ray_batch = kRayBatch;
  {
    const int warps = (int)gridDim.x * kWarpsPerBlock;
    const int fair = (n / (2 * warps) + 31) & ~31;
    ray_batch = fair < 32 ? 32 : (fair < kRayBatch ? fair : kRayBatch);
  }

  int batch_next = 0, batch_end = 0;  // warp-uniform; batch_end < 0: the global counter ran past n
  int stage_base = 0, stage_cnt = 0, stage_pos = 0;  // warp-uniform: staged rays [stage_pos, stage_cnt)
  int ray_idx = -1;                   // < 0: the lane has no ray
  RayPre rp;
  float tmax = 0.f;
  int hit_tri = -1;
  int skip_tri = -1;
  uint2 ngroup = make_uint2(0u, 0u);
  uint2 tq = make_uint2(0u, 0u), tq2 = make_uint2(0u, 0u);  // pending leaf-triangle groups
  int sp = 0;
  TraceCounters cnt;
  cnt.nodes = 0;
  cnt.tris = 0;

  for (;;) {
    // ---- refill idle lanes ----------------------------------------------------------
    // Rays are prepared 32 at a time by the whole warp (coalesced loads, precompute_ray at full
    // lane utilisation) into a shared-memory stage; a lane that finishes its ray only pops the
    // next staged entry.  Per-lane refills ran the same ~140 instructions on ~5 of 32 lanes in
    // almost every trip of the loop (ncu source view: 27 % of all issued instructions).
    const unsigned need = __ballot_sync(0xffffffffu, ray_idx < 0);
    if (need) {
      if (stage_pos >= stage_cnt && batch_end >= 0) {
        if (batch_next >= batch_end) {
          unsigned base = 0;
          if (lane == 0) base = atomicAdd(ray_counter, (unsigned)ray_batch);
          base = __shfl_sync(0xffffffffu, base, 0);
          if (base >= (unsigned)n) {
            batch_end = -1;
          } else {
            batch_next = (int)base;
            batch_end = (int)base + ray_batch < n ? (int)base + ray_batch : n;
            // pull the batch's rays towards the SM now; they are staged 32 at a time later
            if (HAS_LIST) {
              for (int r = batch_next + (int)lane; r < batch_end; r += 32) {
                const int i = __ldg(p.ray_list + r);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.org_tmin + i));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.dir_tmax + i));
              }
            } else {
              for (int r = batch_next + 4 * (int)lane; r < batch_end; r += 128) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.org_tmin + r));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.dir_tmax + r));
              }
            }
          }
        }
        if (batch_end >= 0) {
          const int cnt = batch_end - batch_next < 32 ? batch_end - batch_next : 32;
          __syncwarp();  // every lane has consumed its entry of the previous stage
          if ((int)lane < cnt) {
            const int idx = HAS_LIST ? __ldg(p.ray_list + batch_next + (int)lane) : batch_next + (int)lane;
            if (HAS_LIST) s_stage_idx[threadIdx.x] = idx;
            const float4 o = __ldcs(p.org_tmin + idx);  // streaming: keep the BVH in L2
            const float4 d = __ldcs(p.dir_tmax + idx);
            RayF ray;
            ray.ox = o.x; ray.oy = o.y; ray.oz = o.z; ray.tmin = o.w;
            ray.dx = d.x; ray.dy = d.y; ray.dz = d.z; ray.tmax = d.w;
            const RayPre r = precompute_ray(ray, bvh.bmin, bvh.bmax);
            s_stage[0][threadIdx.x] = o;
            s_stage[1][threadIdx.x] = d;
            s_stage[2][threadIdx.x] = make_float4(r.idx, r.idy, r.idz, r.err);
          }
          __syncwarp();
          stage_base = batch_next;
          stage_cnt = cnt;
          stage_pos = 0;
          batch_next += cnt;
        }
      }
      const int avail = stage_cnt - stage_pos;
      if (ray_idx < 0) {
        const int rank = __popc(need & ((1u << lane) - 1u));
        if (rank < avail) {
          const int e = (int)(threadIdx.x & ~31u) + stage_pos + rank;
          const float4 o = s_stage[0][e], d = s_stage[1][e], c = s_stage[2][e];
          const int idx = HAS_LIST ? s_stage_idx[e] : stage_base + stage_pos + rank;
          rp.ox = o.x; rp.oy = o.y; rp.oz = o.z; rp.tmin = o.w;
          rp.d = mk3(d.x, d.y, d.z);
          rp.idx = c.x; rp.idy = c.y; rp.idz = c.z; rp.err = c.w;
          rp.octinv4 = ray_octinv4(d.x, d.y, d.z);
          tmax = d.w;
          hit_tri = -1;
#if !M3D_RELOAD_SKIP
          if (HAS_SKIP) skip_tri = __ldg(p.skip_tris + idx);
#endif
          // virtual parent whose only child is the root: child base 0, slot (7 ^ octinv)
          // of an all-internal imask so that take_nearest_child() yields node 0
          ngroup.x = 0u;
          ngroup.y = 0x80000000u;
          tq.y = 0u;
          tq2.y = 0u;
          sp = 0;
          ray_idx = idx;
        }
      }
      const int want = __popc(need);
      stage_pos += want < avail ? want : avail;
      if (batch_end < 0 && stage_pos >= stage_cnt && __ballot_sync(0xffffffffu, ray_idx >= 0) == 0u) break;
    }

    if (ray_idx >= 0) {
      // ---- phase A: one node visit (skipped only while both triangle slots are full) ---
      bool have_node = (ngroup.y & 0xff000000u) != 0u;
      if (!have_node && sp > 0 && tq2.y == 0u) {
        --sp;
        ngroup = sp < kSmemStack ? s_stack[sp][threadIdx.x] : l_stack[sp - kSmemStack];
        have_node = true;
      }
      if (have_node && tq2.y == 0u) {
        const uint32_t node_index = take_nearest_child(ngroup, rp.octinv4);
        if (ngroup.y & 0xff000000u) {
          if (sp < kSmemStack)
            s_stack[sp][threadIdx.x] = ngroup;
          else if (sp < kSmemStack + kLocalStack)
            l_stack[sp - kSmemStack] = ngroup;
          sp = sp < kSmemStack + kLocalStack ? sp + 1 : sp;
        }
        if (COUNT) cnt.nodes++;
        uint2 tnew;
#if M3D_MASK_LUT
        intersect_node<true>(nodes, node_index, rp, tmax, ngroup, tnew, M3D_OPAQUE_ONE ? p.one_bits : 0x3f800000u, lut_saddr);
#else
        intersect_node(nodes, node_index, rp, tmax, ngroup, tnew, M3D_OPAQUE_ONE ? p.one_bits : 0x3f800000u);
#endif
        if (tq.y == 0u) {
          tq = tnew;
        } else {
          tq2 = tnew;
          if (kPrefetchQueuedTri && tnew.y) {
            // this leaf waits behind another one for at least a trip: pull its first triangle in
            const float4 *tp = tris + (size_t)(tnew.x + (uint32_t)bfind32(tnew.y)) * 3;
            asm volatile("prefetch.global.L1 [%0];" ::"l"(tp));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(tp + 2));
          }
        }
        if (kPrefetchNextNode && (ngroup.y & 0xff000000u)) {
          // the child the next trip descends into: L2 -> L1 while the triangle phase runs
          uint2 g = ngroup;
          const uint4 *nn = nodes + (size_t)take_nearest_child(g, rp.octinv4) * M3D_NODE_QUADS;
          asm volatile("prefetch.global.L1 [%0];" ::"l"(nn));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(nn + 4));
        }
      }
      // ---- phase B: one triangle test for every lane that has one pending --------------
      // (a per-lane "while" here ran at 2.2 of 32 lanes; one test per trip of the common
      // loop keeps the lanes of a warp in lock step, and the second slot lets the lane
      // keep traversing while a multi-triangle leaf drains)
      // TRI_ROUNDS = 2 (tiny scenes: a handful of nodes, several large triangles per ray) repeats
      // the phase so that the node phase is not paid once per triangle
#pragma unroll
      for (int round = 0; round < TRI_ROUNDS; round++) {
        if (tq.y == 0u) {
          tq = tq2;
          tq2.y = 0u;
        }
        if (tq.y) {
          const int bit = bfind32(tq.y);
          tq.y &= ~(1u << bit);
          const int ti = (int)(tq.x + (uint32_t)bit);
          if (COUNT) cnt.tris++;
          float t, b1, b2;
#if M3D_RELOAD_SKIP
          // the skip id is only needed here: re-read it (4 bytes, L1 / L2 hit) instead of holding a
          // register across the node phase, where the kernel sits exactly at its 80-register budget
          if (HAS_SKIP) skip_tri = __ldg(p.skip_tris + ray_idx);
#endif
          if ((!HAS_SKIP || ti != skip_tri) && intersect_tri(tris + (size_t)ti * 3, rp, tmax, t, b1, b2)) {
            tmax = t;
            hit_tri = ti;
          }
        }
      }
      // ---- ray finished? (checked here so that the lane is refilled before the next A) --
      if ((ngroup.y & 0xff000000u) == 0u && sp == 0 && tq.y == 0u && tq2.y == 0u) {
        // raw hit: float32 t and the triangle's index; finish_hits_kernel does the rest
        __stcs(p.hit0 + ray_idx, make_float4(tmax, 0.f, 0.f, __int_as_float(hit_tri)));
        ray_idx = -1;
      }
    }
  }

  if (COUNT) {
    unsigned long long nn = cnt.nodes, tt = cnt.tris;
    for (int off = 16; off > 0; off >>= 1) {
      nn += __shfl_down_sync(0xffffffffu, nn, off);
      tt += __shfl_down_sync(0xffffffffu, tt, off);
    }
    if (lane == 0) {
      atomicAdd(p.counters, nn);
      atomicAdd(p.counters + 1, tt);
    }
  }
}

// Bounds cull in front of large plain mesh batches: one streaming pass retires the rays that miss the
// bounds of all vertices (C2's ray mix: 54 %) with their raw miss record and appends the others to the
// list the traversal walks.  In the traversal such a ray costs a staged entry, a lane for one trip of
// the lock-step loop (the root's eight child boxes) and a scattered 16-byte store; here it costs 32
// coalesced bytes in and 16 out.  The test is ray_misses_bounds (trace_core.cuh; conservative, covered on
// the CPU by tests/test_emul_traversal.py).
__global__ void __launch_bounds__(256)
cull_rays_kernel(DeviceBVH bvh, const float4 *__restrict__ org_tmin, const float4 *__restrict__ dir_tmax, int n,
                 float4 *__restrict__ hit0, int *__restrict__ list, int *__restrict__ count) {
  constexpr int kPer = 4;
  __shared__ int s_cnt[kPer * 8];
  __shared__ int s_base;
  const int lane = (int)(threadIdx.x & 31u), warp = (int)(threadIdx.x >> 5);
  const int base = (int)blockIdx.x * (256 * kPer) + (int)threadIdx.x;
  unsigned balls[kPer];
  bool keep[kPer];
#pragma unroll
  for (int k = 0; k < kPer; k++) {
    const int i = base + 256 * k;
    keep[k] = false;
    if (i < n) {
      const float4 o = __ldg(org_tmin + i), d = __ldg(dir_tmax + i);
      keep[k] = !ray_misses_bounds(o.x, o.y, o.z, o.w, d.x, d.y, d.z, d.w, bvh.bmin, bvh.bmax);
      if (!keep[k]) __stcs(hit0 + i, make_float4(d.w, 0.f, 0.f, __int_as_float(-1)));
    }
    balls[k] = __ballot_sync(0xffffffffu, keep[k]);
    if (lane == 0) s_cnt[k * 8 + warp] = __popc(balls[k]);
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int v = threadIdx.x < kPer * 8 ? s_cnt[threadIdx.x] : 0;
    int incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += u;
    }
    if (threadIdx.x < kPer * 8) s_cnt[threadIdx.x] = incl - v;
    if (threadIdx.x == 31) s_base = incl ? atomicAdd(count, incl) : 0;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kPer; k++)
    if (keep[k]) list[s_base + s_cnt[k * 8 + warp] + __popc(balls[k] & ((1u << lane) - 1u))] = base + 256 * k;
}

__global__ void pack_rays_kernel(const float *__restrict__ org3, const float *__restrict__ dir3,
                                 int64_t n, float tmin, float tmax, float4 *__restrict__ org_tmin,
                                 float4 *__restrict__ dir_tmax, float sx, float sy, float sz) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // org3 == nullptr: every ray starts at (sx, sy, sz) (camera batches: M3D_TRACE_SHARED_ORIGIN)
  org_tmin[i] = org3 ? make_float4(org3[3 * i], org3[3 * i + 1], org3[3 * i + 2], tmin) : make_float4(sx, sy, sz, tmin);
  dir_tmax[i] = make_float4(dir3[3 * i], dir3[3 * i + 1], dir3[3 * i + 2], tmax);
}

__global__ void unpack_hits_kernel(const float4 *__restrict__ hit0, const float4 *__restrict__ hit1,
                                   int64_t n, float *__restrict__ t, int32_t *__restrict__ prim,
                                   int32_t *__restrict__ obj, float *__restrict__ normal3,
                                   float *__restrict__ bary3) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 a = hit0[i], b = hit1[i];
  if (t) t[i] = a.x;
  if (prim) prim[i] = __float_as_int(a.w);
  if (obj) obj[i] = __float_as_int(b.w);
  if (normal3) {
    normal3[3 * i] = b.x;
    normal3[3 * i + 1] = b.y;
    normal3[3 * i + 2] = b.z;
  }
  if (bary3) {
    const bool hit = __float_as_int(a.w) >= 0;
    bary3[3 * i] = hit ? 1.f - (a.y + a.z) : 0.f;
    bary3[3 * i + 1] = a.y;
    bary3[3 * i + 2] = a.z;
  }
}

// Collider.RayCollisions counts / ColliderContains parity (collisions.go:119-134, 263-273):
// one thread per ray, all-hits traversal (count_bvh_hits).  FIXED_DIR: the rays are the
// containment probes of ColliderContains, origin = query point, the reference's fixed direction.
template <bool FIXED_DIR>
__global__ void __launch_bounds__(128)
count_hits_kernel(DeviceBVH bvh, const float *__restrict__ org3, const float *__restrict__ dir3, int64_t n,
                  int32_t *__restrict__ counts, uint8_t *__restrict__ inside) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  RayF ray;
  ray.ox = org3[3 * i];
  ray.oy = org3[3 * i + 1];
  ray.oz = org3[3 * i + 2];
  if (FIXED_DIR) {
    ray.dx = 0.5224892708603626f;  // collisions.go:126 (rounded to float32 like every direction)
    ray.dy = 0.10494477243214506f;
    ray.dz = 0.43558938446126527f;
  } else {
    ray.dx = dir3[3 * i];
    ray.dy = dir3[3 * i + 1];
    ray.dz = dir3[3 * i + 2];
  }
  ray.tmin = 0.f;
  ray.tmax = INFINITY;
  const int c = bvh.num_tris > 0 ? count_bvh_hits(bvh.nodes, bvh.tris, bvh.bmin, bvh.bmax, ray) : 0;
  if (counts) counts[i] = c;
  if (inside) inside[i] = (uint8_t)(c & 1);
}


// Collider.RayCollisions(r, f) with the collisions delivered (collisions.go:263-273,
// primitives.go:189-196): one thread per ray re-walks the hierarchy exactly like the counting
// pass, writes its hits into the ray's segment [offsets[i], offsets[i+1]) of the outputs, orders
// the segment by t (the reference's callback order is its own BVH's traversal order) and
// re-evaluates every hit in float64 with the reference's arithmetic like the first-hit finish pass.
__global__ void __launch_bounds__(128)
collect_hits_kernel(DeviceBVH bvh, const float *__restrict__ org3, const float *__restrict__ dir3, int64_t n,
                    const int64_t *__restrict__ offsets, float *__restrict__ t_out, int32_t *__restrict__ prim_out,
                    float *__restrict__ normal3, float *__restrict__ bary3) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t begin = offsets[i];
  const int cap = (int)(offsets[i + 1] - begin);
  if (cap <= 0 || bvh.num_tris <= 0) return;
  RayF ray;
  ray.ox = org3[3 * i];
  ray.oy = org3[3 * i + 1];
  ray.oz = org3[3 * i + 2];
  ray.dx = dir3[3 * i];
  ray.dy = dir3[3 * i + 1];
  ray.dz = dir3[3 * i + 2];
  ray.tmin = 0.f;
  ray.tmax = INFINITY;
  float *ts = t_out + begin;
  int32_t *ids = prim_out + begin;  // leaf-order indices until the last loop
  int found = collect_bvh_hits(bvh.nodes, bvh.tris, bvh.bmin, bvh.bmax, ray, cap, ts, ids);
  if (found > cap) found = cap;
  for (int a = 1; a < found; a++) {  // insertion sort by (t, leaf index): segments are short
    const float ta = ts[a];
    const int32_t ia = ids[a];
    int b = a - 1;
    while (b >= 0 && (ts[b] > ta || (ts[b] == ta && ids[b] > ia))) {
      ts[b + 1] = ts[b];
      ids[b + 1] = ids[b];
      b--;
    }
    ts[b + 1] = ta;
    ids[b + 1] = ia;
  }
  for (int k = 0; k < cap; k++) {
    const int64_t o = begin + k;
    if (k >= found) {  // cannot happen for offsets made from the counting pass; keep the output defined
      ts[k] = 0.f;
      ids[k] = -1;
      continue;
    }
    const int32_t ti = ids[k];
    const float4 *tri = bvh.tris + (size_t)ti * 3;
    const HitD r = refine_hit_f64(tri, ray.ox, ray.oy, ray.oz, ray.dx, ray.dy, ray.dz);
    if (r.t >= 0.0) ts[k] = (float)r.t;
    ids[k] = __float_as_int(__ldg(&tri[0].w));
    if (normal3) {
      double nx = r.nx, ny = r.ny, nz = r.nz;
      if (bvh.vnormals) {  // InterpNormalTriangle.InterpNormal (primitives.go:508-516)
        const float4 *vn = bvh.vnormals + (size_t)ti * 3;
        const float4 a = __ldg(vn), b = __ldg(vn + 1), c = __ldg(vn + 2);
        nx = r.b0 * a.x + r.b1 * b.x + r.b2 * c.x;
        ny = r.b0 * a.y + r.b1 * b.y + r.b2 * c.y;
        nz = r.b0 * a.z + r.b1 * b.z + r.b2 * c.z;
        const double s = 1.0 / sqrt(nx * nx + ny * ny + nz * nz);
        nx *= s;
        ny *= s;
        nz *= s;
      }
      normal3[3 * o] = (float)nx;
      normal3[3 * o + 1] = (float)ny;
      normal3[3 * o + 2] = (float)nz;
    }
    if (bary3) {
      bary3[3 * o] = (float)r.b0;
      bary3[3 * o + 1] = (float)r.b1;
      bary3[3 * o + 2] = (float)r.b2;
    }
  }
}

}  // namespace

void launch_collect_hits(const DeviceBVH &bvh, const float *org3, const float *dir3, int64_t n,
                         const int64_t *offsets, float *t, int32_t *prim, float *normal3, float *bary3,
                         cudaStream_t stream) {
  if (n <= 0) return;
  collect_hits_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(bvh, org3, dir3, n, offsets, t, prim,
                                                                        normal3, bary3);
}

void launch_count_hits(const DeviceBVH &bvh, const float *org3, const float *dir3, int64_t n, int32_t *counts,
                       uint8_t *inside, cudaStream_t stream) {
  if (n <= 0) return;
  const unsigned blocks = (unsigned)((n + 127) / 128);
  if (dir3)
    count_hits_kernel<false><<<blocks, 128, 0, stream>>>(bvh, org3, dir3, n, counts, inside);
  else
    count_hits_kernel<true><<<blocks, 128, 0, stream>>>(bvh, org3, nullptr, n, counts, inside);
}

namespace {
}  // namespace

template <bool COUNT, int MIN_BLOCKS, int TRI_ROUNDS, bool HAS_SKIP, bool HAS_LIST = false>
static void launch_trace_instance(const DeviceBVH &bvh, const TraceLaunch &p, cudaStream_t stream) {
  // persistent grid: as many blocks as stay resident
  static const int blocks_per_sm = [] {
    int b = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &b, trace_first_hit_kernel<COUNT, MIN_BLOCKS, TRI_ROUNDS, HAS_SKIP, HAS_LIST>, kTraceBlock, 0);
    return b > 0 ? b : 1;
  }();
  long long want = (p.n + kTraceBlock - 1) / kTraceBlock;
  long long grid = (long long)device_sm_count() * blocks_per_sm;
  if (grid > want) grid = want;
  unsigned int *rc32 = reinterpret_cast<unsigned int *>(p.ray_counter);
  trace_first_hit_kernel<COUNT, MIN_BLOCKS, TRI_ROUNDS, HAS_SKIP, HAS_LIST>
      <<<(unsigned)grid, kTraceBlock, 0, stream>>>(bvh, p, rc32);
}

// M3D_CULL=0|1 overrides the default (tuning / A-B runs)
bool trace_cull_wanted(int64_t n, int64_t num_nodes) {
  static const int mode = [] {
    const char *e = getenv("M3D_CULL");
    return e ? atoi(e) : M3D_CULL_DEFAULT;
  }();
  return mode != 0 && n >= ((int64_t)1 << 20) && n < ((int64_t)1 << 31) - 4096 && num_nodes > kTinySceneNodes;
}

template <bool COUNT, int MIN_BLOCKS, int TRI_ROUNDS = 1>
static void launch_trace_variant(const DeviceBVH &bvh, const TraceLaunch &p, cudaStream_t stream) {
  if (p.skip_tris)
    launch_trace_instance<COUNT, MIN_BLOCKS, TRI_ROUNDS, true>(bvh, p, stream);
  else
    launch_trace_instance<COUNT, MIN_BLOCKS, TRI_ROUNDS, false>(bvh, p, stream);
}

void launch_trace_bvh_only(const DeviceBVH &bvh, const TraceLaunch &p_in, cudaStream_t stream) {
  if (p_in.n <= 0) return;
  // register budget of the traversal kernel: 6 resident blocks/SM (80 registers, no spills)
  // measured best on B200; M3D_TRACE_MINB=7|8 selects the tighter variants for tuning runs
  static const int minb = [] {
    const char *e = getenv("M3D_TRACE_MINB");
    const int v = e ? atoi(e) : 6;
    return (v < 5 || v > 8) ? 6 : v;
  }();
  static const int tri_rounds_env = [] {
    const char *e = getenv("M3D_TRACE_TRI_ROUNDS");
    return e ? atoi(e) : 0;
  }();
  TraceLaunch p = p_in;
  if (p.cull_scratch && !p.skip_tris && !p.n_ptr && bvh.num_nodes > kTinySceneNodes) {
    int *count = p.cull_scratch, *list = p.cull_scratch + 1;
    cudaMemsetAsync(count, 0, sizeof(int), stream);
    cull_rays_kernel<<<(unsigned)((p.n + 1023) / 1024), 256, 0, stream>>>(bvh, p.org_tmin, p.dir_tmax, (int)p.n, p.hit0,
                                                                         list, count);
    p.ray_list = list;
    p.n_ptr = count;
    cudaMemsetAsync(p.ray_counter, 0, sizeof(unsigned long long), stream);
    if (p.counters)
      launch_trace_instance<true, 6, 1, false, true>(bvh, p, stream);
    else
      launch_trace_instance<false, 6, 1, false, true>(bvh, p, stream);
    return;
  }
  // tiny hierarchies (cornell_box: 72 triangles in 4 nodes) spend their time in the triangle phase
  const int tri_rounds = tri_rounds_env > 0 ? tri_rounds_env : (bvh.num_nodes <= kTinySceneNodes ? M3D_TINY_TRI_ROUNDS : 1);
  cudaMemsetAsync(p.ray_counter, 0, sizeof(unsigned long long), stream);
  // M3D_L2_PERSIST_MB (tuning runs): the traversal kernel runs with an access-policy window that
  // marks the node array (M3D_L2_WINDOW=tris: the triangle records) as persisting in L2
  // (read per launch so that one tuning process can sweep the settings)
  const char *l2_env = getenv("M3D_L2_PERSIST_MB");
  const int l2_mb = l2_env ? atoi(l2_env) : 0;
  const bool l2_window = l2_mb > 0 && bvh.num_nodes > kTinySceneNodes;
  if (l2_window) {
    const char *we = getenv("M3D_L2_WINDOW");
    const bool on_tris = we && we[0] == 't';
    const char *re = getenv("M3D_L2_HITRATIO");
    const float ratio = re ? (float)atof(re) : 1.0f;
    static int applied_mb = -1;
    static size_t max_window = 0;
    if (applied_mb != l2_mb) {
      int dev = 0, max_persist = 0, max_win = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
      cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, dev);
      size_t want = (size_t)l2_mb << 20;
      if (want > (size_t)max_persist) want = (size_t)max_persist;
      const cudaError_t e = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
      fprintf(stderr, "m3d: L2 persisting set-aside %zu MB (device max %d MB, max window %d MB): %s\n", want >> 20,
              max_persist >> 20, max_win >> 20, cudaGetErrorString(e));
      max_window = (size_t)max_win;
      applied_mb = l2_mb;
    }
    cudaStreamAttrValue v = {};
    size_t bytes = on_tris ? (size_t)bvh.num_tris * 48 : (size_t)bvh.num_nodes * M3D_NODE_BYTES;
    v.accessPolicyWindow.base_ptr = on_tris ? (void *)bvh.tris : (void *)bvh.nodes;
    v.accessPolicyWindow.num_bytes = bytes < max_window ? bytes : max_window;
    v.accessPolicyWindow.hitRatio = ratio;
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &v);
  }
  struct WindowReset {
    cudaStream_t s;
    bool on;
    ~WindowReset() {
      if (!on) return;
      cudaStreamAttrValue v = {};
      v.accessPolicyWindow.num_bytes = 0;
      cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &v);
    }
  } window_reset{stream, l2_window};
  if (p.counters) {
    launch_trace_variant<true, 6>(bvh, p, stream);
  } else if (tri_rounds >= 2) {
    launch_trace_variant<false, M3D_TINY_MINB, M3D_TINY_TRI_ROUNDS>(bvh, p, stream);
  } else if (minb == 5) {
    launch_trace_variant<false, 5>(bvh, p, stream);
  } else if (minb == 8) {
    launch_trace_variant<false, 8>(bvh, p, stream);
  } else if (minb == 7) {
    launch_trace_variant<false, 7>(bvh, p, stream);
  } else {
    launch_trace_variant<false, 6>(bvh, p, stream);
  }
}

void launch_trace_scene(const DeviceScene &scene, const SceneTraceLaunch &p, cudaStream_t stream) {
  if (p.t.n <= 0) return;
  TraceLaunch t = p.t;
  t.skip_tris = p.skip_ids;  // negative ids (shapes / none) never match a triangle index
  launch_trace_bvh_only(scene.bvh, t, stream);
  launch_finish_scene_hits(scene, p, stream);
}

void launch_trace_first_hit(const DeviceBVH &bvh, const TraceLaunch &p, cudaStream_t stream) {
  DeviceScene sc;
  sc.bvh = bvh;
  SceneTraceLaunch sp;
  sp.t = p;
  launch_trace_scene(sc, sp, stream);
}

void launch_pack_rays(const float *org3, const float *dir3, int64_t n, float tmin, float tmax,
                      float4 *org_tmin, float4 *dir_tmax, cudaStream_t stream, const float *shared_origin) {
  if (n <= 0) return;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  const float sx = shared_origin ? shared_origin[0] : 0.f, sy = shared_origin ? shared_origin[1] : 0.f,
              sz = shared_origin ? shared_origin[2] : 0.f;
  pack_rays_kernel<<<blocks, 256, 0, stream>>>(shared_origin ? nullptr : org3, dir3, n, tmin, tmax, org_tmin,
                                               dir_tmax, sx, sy, sz);
}

void launch_unpack_hits(const float4 *hit0, const float4 *hit1, int64_t n, float *t, int32_t *prim,
                        int32_t *obj, float *normal3, float *bary3, cudaStream_t stream) {
  if (n <= 0) return;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  unpack_hits_kernel<<<blocks, 256, 0, stream>>>(hit0, hit1, n, t, prim, obj, normal3, bary3);
}

int device_sm_count() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

}  // namespace m3d
