// cluster.cu -- GPU clusterer for sm_100a: label-equivalence connected components with union-find in global
// atomics, small-cluster removal, per-cluster bounding box / centre and median-by-speed velocity.
//
// Replaces (scene_flow_clusterer/src/clusterer_nodelet.cpp):
//   calculateInitialClusterMap + comparePoints + LookupTable     :56-83, :186-219, include/lookup_table.h:10-33
//   integrateConnectedClusters                                   :253-267
//   removeSmallClusters                                          :354-393
//   clusterMap2IndicesCluster + cluster2MovingObject             :97-117, :147-184
//   publishMovingObjects (id assignment)                         :324-343
//
// What the reference's raster scan computes, restated as a graph problem:
//   * edge(a,b): b = a + (du,dv), du,dv in [-N,0], (du,dv) != (0,0), both dynamic, !(|z_a - z_b| > depth_diff)
//   * clusters  = connected components with at least one edge (isolated dynamic pixels keep NOT_BELONGED_)
//   * a component's provisional root label is created when the raster scan first visits one of its pixels that has
//     an accepted up-left neighbour, and labels are handed out in raster order of that event; removeSmallClusters
//     renumbers survivors densely in increasing root-label order.  Hence
//         final id = rank of  key(C) = min{ raster index of a in C : a has an accepted up-left neighbour }
//     among the surviving components.  That is what cl_stats/cl_rank reproduce, so ids (and the order of the
//     MovingObjectArray) match the reference exactly, not only up to permutation.
//
// Pipeline (all kernels batched over frames, all per-pixel work gated by the 1-byte dynamic mask):
//   cl_init      parent[i] = i at dynamic pixels
//   cl_union     for every accepted up-left edge: lock-free union (atomicMin on the larger root)
//   cl_flatten   parent[i] = root ; count[root] += 1 (warp-aggregated)
//   cl_slots     roots with count >= cluster_size claim a dense survivor slot
//   cl_stats     per-slot min/max of x,y,z, reference-order key (warp-shuffle reduced, then atomics), |v| plane
//   cl_rank      per frame: final ids by key rank, member-list offsets, work queue for cl_objects
//   cl_scatter   survivors' members -> contiguous per-cluster lists (speed bits, pixel index)
//   cl_objects   one CTA per cluster: 4-pass 8-bit radix select of the (n/2)-th largest |v|, emit MovingObject
//   cl_labels    optional dense label image (cluster_map_)

#include <math_constants.h>

#include "common.cuh"

namespace mod {

namespace {

constexpr int kBlock = 256;
constexpr unsigned kFull = 0xffffffffu;

// ---- helpers ---------------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ int uf_find(const int* parent, int i) {
  int p;
  while ((p = __ldcg(parent + i)) != i) i = p;
  return i;
}

// Lock-free union: the larger root is pointed at the smaller one with atomicMin; on interference retry from the
// value that was found there.  Parent links only ever decrease, so stale reads are harmless.
__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  for (;;) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) {
      const int t = a;
      a = b;
      b = t;
    }
    const int old = atomicMin(parent + a, b);
    if (old == a) return;
    a = old;
  }
}

// ---- cl_init -----------------------------------------------------------------------------------------------------
__global__ void cl_init_kernel(const ClLaunch L) {
  const size_t P = static_cast<size_t>(L.w) * L.h;
  const size_t total = P * L.batch;
  const int lane = threadIdx.x & 31;
  for (size_t g0 = blockIdx.x * static_cast<size_t>(blockDim.x); g0 < total;
       g0 += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t g = g0 + threadIdx.x;
    const bool dyn = (g < total) && (L.mask[g] != 0);
    const unsigned ballot = __ballot_sync(kFull, dyn);
    if (!dyn) continue;
    const size_t frame = g / P;
    const int i = static_cast<int>(g - frame * P);
    L.parent[g] = i;
    L.count[g] = 0;
    L.flags[g] = 0;
    // diagnostics: dynamic pixel count per frame, one atomic per (warp, frame)
    const unsigned same = __match_any_sync(ballot, static_cast<int>(frame));
    if (lane == __ffs(same) - 1) atomicAdd(L.n_dynamic + frame, __popc(same));
  }
}

// ---- cl_union ---------------------------------------------------------------------------------------------------
// comparePoints (clusterer_nodelet.cpp:186-219) for all (N+1)^2-1 up-left neighbours of every dynamic pixel.
__global__ void cl_union_kernel(const ClLaunch L) {
  const size_t P = static_cast<size_t>(L.w) * L.h;
  const size_t total = P * L.batch;
  for (size_t g = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; g < total;
       g += static_cast<size_t>(gridDim.x) * blockDim.x) {
    if (L.mask[g] == 0) continue;
    const size_t frame = g / P;
    const int i = static_cast<int>(g - frame * P);
    const int v = i / L.w;
    const int u = i - v * L.w;
    const uint8_t* mask = L.mask + frame * P;
    const float* pts = L.points + frame * L.points_frame_stride * 8;
    int* parent = L.parent + frame * P;
    const float z_here = __ldg(pts + static_cast<size_t>(i) * 8 + 2);
    bool any = false;
    const int v0 = max(v - L.n, 0), u0 = max(u - L.n, 0);  // isInRange
    for (int vv = v0; vv <= v; vv++) {
      for (int uu = u0; uu <= u; uu++) {
        const int j = vv * L.w + uu;
        if (j == i) continue;
        if (mask[j] == 0) continue;  // isDynamic
        const float z_there = __ldg(pts + static_cast<size_t>(j) * 8 + 2);
        if (fabsf(__fsub_rn(z_here, z_there)) > L.depth_thr) continue;  // depthDiff > depth_diff_th_ (NaN passes)
        any = true;
        uf_union(parent, i, j);
      }
    }
    if (any) L.flags[g] = 1;
  }
}

// ---- cl_flatten -------------------------------------------------------------------------------------------------
// integrateConnectedClusters (:253-267) + the histogram of removeSmallClusters (:359-367)
__global__ void cl_flatten_kernel(const ClLaunch L) {
  const size_t P = static_cast<size_t>(L.w) * L.h;
  const size_t total = P * L.batch;
  const int lane = threadIdx.x & 31;
  for (size_t g0 = blockIdx.x * static_cast<size_t>(blockDim.x); g0 < total;
       g0 += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t g = g0 + threadIdx.x;
    const bool dyn = (g < total) && (L.mask[g] != 0);
    const unsigned ballot = __ballot_sync(kFull, dyn);
    if (!dyn) continue;
    const size_t frame = g / P;
    const int i = static_cast<int>(g - frame * P);
    int* parent = L.parent + frame * P;
    const int root = uf_find(parent, i);
    parent[i] = root;
    // warp-aggregated size count: lanes with the same (frame, root) elect one leader
    const long long keyv = (static_cast<long long>(frame) << 32) | static_cast<unsigned>(root);
    const unsigned same = __match_any_sync(ballot, keyv);
    if (lane == __ffs(same) - 1) atomicAdd(L.count + frame * P + root, __popc(same));
  }
}

// ---- cl_slots ---------------------------------------------------------------------------------------------------
// removeSmallClusters (:369-383): components below cluster_size are dropped; survivors get a dense slot.
__global__ void cl_slots_kernel(const ClLaunch L) {
  const size_t P = static_cast<size_t>(L.w) * L.h;
  const size_t total = P * L.batch;
  for (size_t g = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; g < total;
       g += static_cast<size_t>(gridDim.x) * blockDim.x) {
    if (L.mask[g] == 0) continue;
    const size_t frame = g / P;
    const int i = static_cast<int>(g - frame * P);
    if (L.parent[g] != i) continue;  // not a root
    const int cnt = L.count[g];
    int s = -1;
    if (cnt >= L.min_size) {
      s = atomicAdd(L.n_slots + frame, 1);
      if (s < L.max_clusters) {
        ClusterStat& st = L.stats[frame * L.max_clusters + s];
        st.root = i;
        st.count = cnt;
        st.key = 0x7fffffff;
        st.rank = -1;
        st.mn[0] = st.mn[1] = st.mn[2] = 0xffffffffu;
        st.mx[0] = st.mx[1] = st.mx[2] = 0u;
        st.sel_prefix = 0;
        st.sel_k = 0;
        st.median_idx = 0x7fffffff;
      } else {
        s = -1;  // overflow: reported through frame_stats.status by cl_rank
      }
    }
    L.slot[g] = s;
  }
}

// ---- cl_stats ---------------------------------------------------------------------------------------------------
// pcl::getMinMax3D over the members (cluster2MovingObject :151-152) and the reference-order key, reduced with warp
// shuffles per (frame, slot) group before touching the per-cluster atomics.  Also stores |v| for the select.
__global__ void cl_stats_kernel(const ClLaunch L) {
  const size_t P = static_cast<size_t>(L.w) * L.h;
  const size_t total = P * L.batch;
  const int lane = threadIdx.x & 31;
  for (size_t g0 = blockIdx.x * static_cast<size_t>(blockDim.x); g0 < total;
       g0 += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t g = g0 + threadIdx.x;
    int s = -1;
    size_t frame = 0;
    int i = 0;
    if (g < total && L.mask[g] != 0) {
      frame = g / P;
      i = static_cast<int>(g - frame * P);
      const int root = L.parent[g];
      s = L.slot[frame * P + root];
    }
    const unsigned member = __ballot_sync(kFull, s >= 0);
    if (member == 0) continue;
    uint32_t ex_mn[3], ex_mx[3];
    int key = 0x7fffffff;
    long long group = -1 - lane;  // distinct for non-members so they never match each other
    if (s >= 0) {
      const float* rec = L.points + (frame * L.points_frame_stride + i) * 8;
      const float4 xyz = __ldg(reinterpret_cast<const float4*>(rec));
      const float4 vel = __ldg(reinterpret_cast<const float4*>(rec + 4));
      const float c[3] = {xyz.x, xyz.y, xyz.z};
#pragma unroll
      for (int k = 0; k < 3; k++) {
        // NaN coordinates are skipped (the reference's running min/max is order dependent there; unpinned)
        const bool nan = isnan(c[k]);
        ex_mn[k] = nan ? 0xffffffffu : float_to_ordered(c[k]);
        ex_mx[k] = nan ? 0u : float_to_ordered(c[k]);
      }
      if (L.flags[g] & 1) key = i;
      L.speed[g] = __fsqrt_rn(
          __fadd_rn(__fadd_rn(__fmul_rn(vel.x, vel.x), __fmul_rn(vel.y, vel.y)), __fmul_rn(vel.z, vel.z)));
      group = (static_cast<long long>(frame) << 32) | static_cast<unsigned>(s);
    } else {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        ex_mn[k] = 0xffffffffu;
        ex_mx[k] = 0u;
      }
    }
    // Peel groups one at a time: usually one or two clusters per warp.
    unsigned todo = member;
    while (todo) {
      const int leader = __ffs(todo) - 1;
      const long long lg = __shfl_sync(kFull, group, leader);
      const bool in = (group == lg);
      uint32_t mn[3], mx[3];
      int kk = in ? key : 0x7fffffff;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        mn[k] = in ? ex_mn[k] : 0xffffffffu;
        mx[k] = in ? ex_mx[k] : 0u;
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
          mn[k] = min(mn[k], __shfl_xor_sync(kFull, mn[k], off));
          mx[k] = max(mx[k], __shfl_xor_sync(kFull, mx[k], off));
        }
        kk = min(kk, __shfl_xor_sync(kFull, kk, off));
      }
      if (lane == leader) {
        ClusterStat& st = L.stats[frame * L.max_clusters + s];
#pragma unroll
        for (int k = 0; k < 3; k++) {
          atomicMin(&st.mn[k], mn[k]);
          atomicMax(&st.mx[k], mx[k]);
        }
        if (kk != 0x7fffffff) atomicMin(&st.key, kk);
      }
      todo &= ~__ballot_sync(kFull, in);
    }
  }
}

// ---- cl_rank ----------------------------------------------------------------------------------------------------
// One CTA per frame.  Final id = rank of the key (removeSmallClusters' dense renumbering keeps root-label order);
// member-list offsets are an exclusive scan of the counts in slot order; clusters are appended to the work queue.
__global__ void __launch_bounds__(kBlock)
    cl_rank_kernel(const ClLaunch L, int* __restrict__ work_count, int2* __restrict__ work_items,
                   int* __restrict__ member_offset) {
  extern __shared__ int s_keys[];  // [n]
  __shared__ int s_part[kBlock];
  __shared__ int s_base;
  const int frame = blockIdx.x;
  int n = L.n_slots[frame];
  const bool overflow = n > L.max_clusters;
  if (overflow) n = L.max_clusters;
  ClusterStat* st = L.stats + static_cast<size_t>(frame) * L.max_clusters;
  for (int a = threadIdx.x; a < n; a += kBlock) s_keys[a] = st[a].key;
  __syncthreads();
  for (int a = threadIdx.x; a < n; a += kBlock) {
    const int ka = s_keys[a];
    int r = 0;
    for (int b = 0; b < n; b++) r += (s_keys[b] < ka) ? 1 : 0;
    st[a].rank = r;
  }
  // exclusive scan of the member counts in slot order: per-thread chunk sums, then a 256-entry scan
  const int chunk = (n + kBlock - 1) / kBlock;
  const int lo = min(n, static_cast<int>(threadIdx.x) * chunk), hi = min(n, lo + chunk);
  int local = 0;
  for (int a = lo; a < hi; a++) local += st[a].count;
  s_part[threadIdx.x] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int t = 0; t < kBlock; t++) {
      const int v = s_part[t];
      s_part[t] = acc;
      acc += v;
    }
    s_base = atomicAdd(work_count, n);
    mod_frame_stats fs;
    fs.n_dynamic = L.n_dynamic[frame];
    fs.n_clusters = n;
    fs.n_objects = n;  // every survivor is accepted: its members all satisfy |v| >= dynamic_speed (:176)
    fs.status = overflow ? MOD_ERR_CLUSTER_SLOTS : (n > L.max_objs ? MOD_ERR_CAPACITY : MOD_OK);
    L.frame_stats[frame] = fs;
  }
  __syncthreads();
  int acc = s_part[threadIdx.x];
  for (int a = lo; a < hi; a++) {
    member_offset[static_cast<size_t>(frame) * L.max_clusters + a] = acc;
    acc += st[a].count;
  }
  for (int a = threadIdx.x; a < n; a += kBlock) work_items[s_base + a] = make_int2(frame, a);
}

// ---- cl_scatter -------------------------------------------------------------------------------------------------
// clusterMap2IndicesCluster (:97-117): contiguous member list per surviving cluster (order inside a list is
// irrelevant to every consumer: min/max, k-th largest speed).
__global__ void cl_scatter_kernel(const ClLaunch L, const int* __restrict__ member_offset, int* __restrict__ cursor,
                                  uint32_t* __restrict__ member_speed, int* __restrict__ member_pixel) {
  const size_t P = static_cast<size_t>(L.w) * L.h;
  const size_t total = P * L.batch;
  const int lane = threadIdx.x & 31;
  for (size_t g0 = blockIdx.x * static_cast<size_t>(blockDim.x); g0 < total;
       g0 += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t g = g0 + threadIdx.x;
    int s = -1;
    size_t frame = 0;
    int i = 0;
    if (g < total && L.mask[g] != 0) {
      frame = g / P;
      i = static_cast<int>(g - frame * P);
      s = L.slot[frame * P + L.parent[g]];
    }
    const unsigned member = __ballot_sync(kFull, s >= 0);
    if (member == 0) continue;
    const long long group = (s >= 0) ? ((static_cast<long long>(frame) << 32) | static_cast<unsigned>(s)) : (-1 - lane);
    const unsigned same = __match_any_sync(kFull, group);
    if (s >= 0) {
      const int leader = __ffs(same) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(cursor + frame * L.max_clusters + s, __popc(same));
      base = __shfl_sync(same, base, leader);
      const int pos = member_offset[frame * L.max_clusters + s] + base + __popc(same & ((1u << lane) - 1));
      member_speed[frame * P + pos] = __float_as_uint(L.speed[g]);  // speeds are >= 0: bit pattern is monotone
      member_pixel[frame * P + pos] = i;
    }
  }
}

// ---- cl_objects -------------------------------------------------------------------------------------------------
// cluster2MovingObject (:147-184) for one cluster per CTA (persistent CTAs pull clusters from the work queue):
// bbox = max - min, centre = (min + max) / 2 (float), velocity = member with the (n/2)-th largest |v|.
// The k-th largest is found by a 4-pass MSB-first radix select over the cluster's member list with a 256-bin
// shared-memory histogram; among members tied at that speed the smallest raster index is taken (the reference's
// std::sort leaves ties unspecified).
__global__ void __launch_bounds__(kBlock)
    cl_objects_kernel(const ClLaunch L, const int* __restrict__ work_count, const int2* __restrict__ work_items,
                      int* __restrict__ work_cursor, const int* __restrict__ member_offset,
                      const uint32_t* __restrict__ member_speed, const int* __restrict__ member_pixel) {
  __shared__ unsigned hist[256];
  __shared__ int s_item;
  __shared__ uint32_t s_prefix;
  __shared__ uint32_t s_k;
  __shared__ int s_best;
  const size_t P = static_cast<size_t>(L.w) * L.h;
  const int n_work = *work_count;
  for (;;) {
    if (threadIdx.x == 0) s_item = atomicAdd(work_cursor, 1);
    __syncthreads();
    const int item = s_item;
    if (item >= n_work) return;
    const int2 w = work_items[item];
    const int frame = w.x, slot = w.y;
    ClusterStat& st = L.stats[static_cast<size_t>(frame) * L.max_clusters + slot];
    const int n = st.count;
    const uint32_t* speeds = member_speed + frame * P + member_offset[static_cast<size_t>(frame) * L.max_clusters + slot];
    const int* pixels = member_pixel + frame * P + member_offset[static_cast<size_t>(frame) * L.max_clusters + slot];
    if (threadIdx.x == 0) {
      s_prefix = 0;
      s_k = static_cast<uint32_t>(n / 2);
      s_best = 0x7fffffff;
    }
    for (int pass = 0; pass < 4; pass++) {
      const int shift = 24 - 8 * pass;
      hist[threadIdx.x] = 0;  // kBlock == 256
      __syncthreads();
      const uint32_t prefix = s_prefix;
      const uint32_t himask = (pass == 0) ? 0u : (0xffffffffu << (shift + 8));
      for (int j = threadIdx.x; j < n; j += kBlock) {
        const uint32_t b = speeds[j];
        if ((b & himask) == prefix) atomicAdd(&hist[(b >> shift) & 0xff], 1u);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        uint32_t k = s_k;
        int bin = 255;
        for (; bin > 0; bin--) {  // descending: k counts from the largest
          const uint32_t c = hist[bin];
          if (k < c) break;
          k -= c;
        }
        s_k = k;
        s_prefix = prefix | (static_cast<uint32_t>(bin) << shift);
      }
      __syncthreads();
    }
    const uint32_t target = s_prefix;
    int best = 0x7fffffff;
    for (int j = threadIdx.x; j < n; j += kBlock)
      if (speeds[j] == target) best = min(best, pixels[j]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) best = min(best, __shfl_xor_sync(kFull, best, off));
    if ((threadIdx.x & 31) == 0 && best != 0x7fffffff) atomicMin(&s_best, best);
    __syncthreads();
    if (threadIdx.x == 0) {
      const int id = st.rank;
      st.median_idx = s_best;
      st.sel_prefix = target;
      if (id >= 0 && id < L.max_objs) {
        mod_object o;
        o.id = id;
        const float* rec = L.points + (static_cast<size_t>(frame) * L.points_frame_stride + s_best) * 8;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const float mn = ordered_to_float(st.mn[k]);
          const float mx = ordered_to_float(st.mx[k]);
          o.bbox[k] = __fsub_rn(mx, mn);
          o.center[k] = __fdiv_rn(__fadd_rn(mn, mx), 2.0f);
          o.velocity[k] = rec[4 + k];
        }
        o.n_points = n;
        L.objs[static_cast<size_t>(frame) * L.max_objs + id] = o;
      }
    }
    __syncthreads();
  }
}

// ---- cl_labels --------------------------------------------------------------------------------------------------
// Final cluster_map_ (after removeSmallClusters' relabel, :386-392), dense, -1 = NOT_BELONGED_.
__global__ void cl_labels_kernel(const ClLaunch L) {
  const size_t P = static_cast<size_t>(L.w) * L.h;
  const size_t total = P * L.batch;
  for (size_t g = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; g < total;
       g += static_cast<size_t>(gridDim.x) * blockDim.x) {
    int label = -1;
    if (L.mask[g] != 0) {
      const size_t frame = g / P;
      const int s = L.slot[frame * P + L.parent[g]];
      if (s >= 0) label = L.stats[frame * L.max_clusters + s].rank;
    }
    L.labels[g] = label;
  }
}

__global__ void cl_reset_kernel(int* a, size_t na, int* b, size_t nb, int* c, size_t nc) {
  const size_t t = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (t < na) a[t] = 0;
  if (t < nb) b[t] = 0;
  if (t < nc) c[t] = 0;
}

}  // namespace

cudaError_t launch_cluster(cudaStream_t s, const ClLaunch& L, const ClExtra& X, int* n_launches) {
  const size_t P = static_cast<size_t>(L.w) * L.h;
  const size_t total = P * L.batch;
  size_t blocks = (total + kBlock - 1) / kBlock;
  const size_t cap = 148 * 16;  // grid-stride: a multiple of the SM count, 16 CTAs of 256 threads per SM
  if (blocks > cap) blocks = cap;
  const unsigned nb = static_cast<unsigned>(blocks);

  const size_t n_fs = static_cast<size_t>(L.batch);
  const size_t n_cs = static_cast<size_t>(L.batch) * L.max_clusters;
  {
    // zero n_slots / n_dynamic (adjacent, 2*batch), cursor (batch*max_clusters), work_count+work_cursor (2)
    const size_t m = n_cs > 2 * n_fs ? n_cs : 2 * n_fs;
    cl_reset_kernel<<<static_cast<unsigned>((m + 255) / 256), 256, 0, s>>>(L.n_slots, 2 * n_fs, X.cursor, n_cs,
                                                                          X.work_count, 2);
  }
  cl_init_kernel<<<nb, kBlock, 0, s>>>(L);
  cl_union_kernel<<<nb, kBlock, 0, s>>>(L);
  cl_flatten_kernel<<<nb, kBlock, 0, s>>>(L);
  cl_slots_kernel<<<nb, kBlock, 0, s>>>(L);
  cl_stats_kernel<<<nb, kBlock, 0, s>>>(L);
  cl_rank_kernel<<<L.batch, kBlock, static_cast<size_t>(L.max_clusters) * sizeof(int), s>>>(L, X.work_count, X.work_items,
                                                                                          X.member_offset);
  cl_scatter_kernel<<<nb, kBlock, 0, s>>>(L, X.member_offset, X.cursor, X.member_speed, X.member_pixel);
  cl_objects_kernel<<<148 * 4, kBlock, 0, s>>>(L, X.work_count, X.work_items, X.work_cursor, X.member_offset,
                                               X.member_speed, X.member_pixel);
  if (L.labels != nullptr) cl_labels_kernel<<<nb, kBlock, 0, s>>>(L);
  if (n_launches) *n_launches = 9 + (L.labels != nullptr ? 1 : 0);
  return cudaGetLastError();
}

}  // namespace mod
