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
//     among the surviving components.  cl_stats/cl_rank reproduce that, so ids (and the order of the
//     MovingObjectArray) match the reference exactly, not only up to permutation.
//
// Pipeline (batched over frames).  The 1-byte dynamic mask is read ONCE (16 bytes per thread); everything after
// works on the compacted list of dynamic pixels, so cost scales with the number of dynamic pixels:
//   cl_tile      one CTA per 32x32 tile: mask+z tile with an N-pixel top/left halo staged in shared memory, tile-local
//                union-find in shared memory over the edges that are not implied by a shorter accepted 2-path (see
//                window_edges), then parent[i] = tile-local root, count[i] = 0, dynamic pixels appended to a list
//   cl_merge     pixels within N of their tile's top/left border: the same window logic on global memory, lock-free
//                union (atomicMin + path halving) of the edges that cross a tile border
//   cl_flatten   parent[i] = root ; count[root] += 1 (warp-aggregated)
//   cl_slots     roots with count >= cluster_size claim a dense survivor slot (stored over count[root])
//   cl_stats     per-slot min/max of x,y,z and reference-order key: warp-shuffle reduction per cluster, then atomics
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
constexpr int kListGrid = 148 * 8;  // grid-stride kernels over the dynamic-pixel list: a multiple of the SM count

// ---- helpers ---------------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// find with path halving.  Links only ever point to a smaller index inside the same (eventual) component, so the
// plain stores race benignly with concurrent unions (a link is only ever replaced by a link to an ancestor).
__device__ __forceinline__ int uf_find(int* parent, int i) {
  for (;;) {
    const int p = __ldcg(parent + i);
    if (p == i) return i;
    const int gp = __ldcg(parent + p);
    if (gp == p) return p;
    __stcg(parent + i, gp);
    i = gp;
  }
}

// Lock-free union: the larger root is pointed at the smaller one with atomicMin; on interference retry from the
// value that was found there.
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

__device__ __forceinline__ bool depth_gate(float za, float zb, float thr) {
  return !(fabsf(__fsub_rn(za, zb)) > thr);  // comparePoints: "if (depthDiff > depth_diff_th_) return;" -- NaN passes
}

// ---- window logic shared by cl_tile and cl_merge ------------------------------------------------------------------
// comparePoints (clusterer_nodelet.cpp:186-219) over the (N+1)^2-1 up-left neighbours of one dynamic pixel `a`.
//
// The window (dynamic flag + z) is loaded once through `probe(wx, wy, z)` (window coordinates: the neighbour at
// offset (wx-N, wy-N); a itself sits at (N,N)).  An accepted edge (a,b) is NOT unioned when a dynamic pixel c exists with
// c = b+(1,0) or c = b+(0,1), c != a, c inside a's window, edge(a,c) accepted and edge(c,b) accepted: (c,b) is a
// unit-length up-left edge of c and unit edges are never skipped, while (a,c) is one step shorter than (a,b), so by
// induction on the Manhattan length the endpoints of every accepted edge still end up in one set -- no matter which
// kernel (tile-local or cross-tile merge) performs the surviving unions.  On a smooth blob this leaves two unions
// per pixel (left, up) instead of (N+1)^2-1.
//
// The evaluation is branch-free SIMD work; the data-dependent union loops run afterwards over the returned edge set,
// so a warp never serialises over (N+1)^2-1 inlined union sites.
template <int W>
struct EdgeSet {  // one bit per window cell k = wy*(N+1)+wx; W = 1 for N <= 4 (25 bits), 4 for N <= 10 (121 bits)
  uint32_t w[W];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int q = 0; q < W; q++) w[q] = 0u;
  }
  __device__ __forceinline__ void set(int k) {
    if (W == 1)
      w[0] |= 1u << k;
    else
      w[k >> 5] |= 1u << (k & 31);
  }
  __device__ __forceinline__ bool empty() const {
    uint32_t o = 0u;
#pragma unroll
    for (int q = 0; q < W; q++) o |= w[q];
    return o == 0u;
  }
  __device__ __forceinline__ int pop() {  // lowest set bit, removed
#pragma unroll
    for (int q = 0; q < W; q++) {
      if (w[q]) {
        const int b = __ffs(w[q]) - 1;
        w[q] &= w[q] - 1;
        return q * 32 + b;
      }
    }
    return -1;
  }
};
template <int NC>
using EdgeSetFor = EdgeSet<(NC > 0 ? ((NC + 1) * (NC + 1) + 31) / 32 : 4)>;

// Returns whether a has any accepted up-left neighbour; `edges` receives the edges that must be unioned.
template <int NC, typename Probe>
__device__ __forceinline__ bool window_edges(int n_runtime, float z_here, float depth_thr, Probe probe,
                                             EdgeSetFor<NC>& edges) {
  constexpr int kMaxN = 10;
  constexpr int S = (NC > 0 ? NC : kMaxN) + 1;
  const int N = NC > 0 ? NC : n_runtime;
  const int stride = NC > 0 ? S : N + 1;  // bit numbering uses the real window pitch
  float zw[S * S];
  bool acc[S * S];
  bool any = false;
  edges.clear();
#pragma unroll(NC > 0 ? S : 1)
  for (int wy = 0; wy < S; wy++) {
#pragma unroll(NC > 0 ? S : 1)
    for (int wx = 0; wx < S; wx++) {
      const int k = wy * S + wx;
      acc[k] = false;
      zw[k] = 0.0f;
      if (wy > N || wx > N || (wy == N && wx == N)) continue;
      float z;
      if (!probe(wx, wy, z)) continue;  // isInRange && isDynamic
      zw[k] = z;
      acc[k] = depth_gate(z_here, z, depth_thr);
      any = any || acc[k];
    }
  }
  if (!any) return false;
#pragma unroll(NC > 0 ? S : 1)
  for (int wy = 0; wy < S; wy++) {
#pragma unroll(NC > 0 ? S : 1)
    for (int wx = 0; wx < S; wx++) {
      const int k = wy * S + wx;
      if (wy > N || wx > N || !acc[k]) continue;
      bool implied = false;
      if (wx < N && !(wy == N && wx + 1 == N)) implied = acc[k + 1] && depth_gate(zw[k + 1], zw[k], depth_thr);
      if (!implied && wy < N && !(wy + 1 == N && wx == N))
        implied = acc[k + S] && depth_gate(zw[k + S], zw[k], depth_thr);
      if (!implied) edges.set(wy * stride + wx);
    }
  }
  return true;
}

// ---- cl_tile ------------------------------------------------------------------------------------------------------
constexpr int kTileW = 32, kTileH = 32;

__device__ __forceinline__ int smem_find(volatile int* lab, int i) {
  for (;;) {
    const int p = lab[i];
    if (p == i) return i;
    const int gp = lab[p];
    if (gp == p) return p;
    lab[i] = gp;
    i = gp;
  }
}
__device__ __forceinline__ void smem_union(int* lab, int a, int b) {
  for (;;) {
    a = smem_find(lab, a);
    b = smem_find(lab, b);
    if (a == b) return;
    if (a < b) {
      const int t = a;
      a = b;
      b = t;
    }
    const int old = atomicMin(lab + a, b);
    if (old == a) return;
    a = old;
  }
}

template <int NC>
__global__ void __launch_bounds__(kBlock)
    cl_tile_kernel(const ClLaunch L, uint32_t* __restrict__ list, int* __restrict__ list_count,
                   uint32_t* __restrict__ border_list, int* __restrict__ border_count) {
  constexpr int kMaxN = 10;
  constexpr int HN = NC > 0 ? NC : kMaxN;  // halo capacity
  constexpr int SW = kTileW + HN;          // shared-memory row pitch
  const int N = NC > 0 ? NC : L.n;
  const int stride = N + 1;
  __shared__ uint8_t s_m[(kTileH + HN) * SW];
  __shared__ float s_z[(kTileH + HN) * SW];
  __shared__ int s_lab[kTileW * kTileH];
  __shared__ uint16_t s_dl[kTileW * kTileH];
  __shared__ uint16_t s_bl[kTileW * kTileH];
  __shared__ int s_cnt, s_bcnt, s_base, s_bbase;
  const int frame = blockIdx.z;
  const int tx0 = blockIdx.x * kTileW, ty0 = blockIdx.y * kTileH;
  const size_t P = static_cast<size_t>(L.w) * L.h;
  const uint8_t* mask = L.mask + frame * P;
  const float* pts = L.points + frame * L.points_frame_stride * 8;
  const int lane = threadIdx.x & 31;

  // A: mask of the tile plus its N-pixel top/left halo; shared (sy,sx) <-> pixel (ty0-N+sy, tx0-N+sx)
  if (threadIdx.x == 0) {
    s_cnt = 0;
    s_bcnt = 0;
  }
  const int span_w = kTileW + N, span_h = kTileH + N;
  bool any_core = false;
  for (int k = threadIdx.x; k < span_w * span_h; k += kBlock) {
    const int sy = k / span_w, sx = k - sy * span_w;
    const int y = ty0 - N + sy, x = tx0 - N + sx;
    uint8_t m = 0;
    if (x >= 0 && y >= 0 && x < L.w && y < L.h) m = mask[static_cast<size_t>(y) * L.w + x];
    s_m[sy * SW + sx] = m;
    any_core = any_core || (m != 0 && sy >= N && sx >= N);
  }
  if (!__syncthreads_or(any_core)) return;

  // B: z of the dynamic pixels (strided gather out of the 32-byte records), label init, tile-local dynamic list
  for (int k = threadIdx.x; k < span_w * span_h; k += kBlock) {
    const int sy = k / span_w, sx = k - sy * span_w;
    if (s_m[sy * SW + sx])
      s_z[sy * SW + sx] = __ldg(pts + (static_cast<size_t>(ty0 - N + sy) * L.w + (tx0 - N + sx)) * 8 + 2);
  }
  for (int c = threadIdx.x; c < kTileW * kTileH; c += kBlock) {  // a warp covers one tile row
    const int ly = c >> 5, lx = c & 31;
    const bool dyn = s_m[(ly + N) * SW + lx + N] != 0;
    s_lab[c] = dyn ? c : -1;
    const unsigned ballot = __ballot_sync(kFull, dyn);
    int base = 0;
    if (lane == 0 && ballot) base = atomicAdd(&s_cnt, __popc(ballot));
    base = __shfl_sync(kFull, base, 0);
    if (dyn) s_dl[base + __popc(ballot & ((1u << lane) - 1))] = static_cast<uint16_t>(c);
  }
  __syncthreads();
  const int nd = s_cnt;

  // C: edge sets (SIMD), then tile-local unions in shared memory; pixels with an edge leaving the tile are noted
  for (int e0 = 0; e0 < nd; e0 += kBlock) {
    const int e = e0 + threadIdx.x;
    EdgeSetFor<NC> edges;
    edges.clear();
    int c = 0, lx = 0, ly = 0;
    bool crosses = false;
    if (e < nd) {
      c = s_dl[e];
      ly = c >> 5;
      lx = c & 31;
      const float z_here = s_z[(ly + N) * SW + lx + N];
      auto probe = [&](int wx, int wy, float& z) {
        const int si = (ly + wy) * SW + lx + wx;
        if (!s_m[si]) return false;
        z = s_z[si];
        return true;
      };
      const bool any = window_edges<NC>(N, z_here, L.depth_thr, probe, edges);
      L.flags[frame * P + static_cast<size_t>(ty0 + ly) * L.w + tx0 + lx] = any ? 1 : 0;
    }
    while (__any_sync(kFull, !edges.empty())) {  // warp-uniform loop: lanes stay converged between unions
      if (!edges.empty()) {
        const int k = edges.pop();
        const int wy = k / stride, wx = k - wy * stride;
        const int nx = lx - N + wx, ny = ly - N + wy;
        if (nx >= 0 && ny >= 0)
          smem_union(s_lab, c, ny * kTileW + nx);
        else
          crosses = true;  // the up-left endpoint lies in another tile: cl_merge
      }
      __syncwarp();
    }
    const unsigned ballot = __ballot_sync(kFull, crosses);
    if (ballot) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&s_bcnt, __popc(ballot));
      base = __shfl_sync(kFull, base, 0);
      if (crosses) s_bl[base + __popc(ballot & ((1u << lane) - 1))] = static_cast<uint16_t>(c);
    }
  }
  __syncthreads();

  // D: flatten inside the tile and publish: parent = tile-local root (as a frame raster index), list entries
  const int nb = s_bcnt;
  if (threadIdx.x == 0) {
    s_base = atomicAdd(list_count, nd);
    s_bbase = nb ? atomicAdd(border_count, nb) : 0;
    atomicAdd(L.n_dynamic + frame, nd);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < nd; e += kBlock) {
    const int c = s_dl[e];
    const int r = smem_find(s_lab, c);
    const size_t i = static_cast<size_t>(ty0 + (c >> 5)) * L.w + tx0 + (c & 31);
    L.parent[frame * P + i] = (ty0 + (r >> 5)) * L.w + tx0 + (r & 31);
    L.count[frame * P + i] = 0;
    list[s_base + e] = static_cast<uint32_t>(frame * P + i);
  }
  for (int e = threadIdx.x; e < nb; e += kBlock) {
    const int c = s_bl[e];
    border_list[s_bbase + e] = static_cast<uint32_t>(frame * P + static_cast<size_t>(ty0 + (c >> 5)) * L.w + tx0 + (c & 31));
  }
}

// ---- cl_merge -----------------------------------------------------------------------------------------------------
// Edges whose up-left endpoint lies in another tile, for the pixels cl_tile put on the border list.  Same window logic,
// on global memory; lock-free union with atomicMin + path halving.
template <int NC>
__global__ void __launch_bounds__(kBlock)
    cl_merge_kernel(const ClLaunch L, const uint32_t* __restrict__ border_list, const int* __restrict__ border_count) {
  const int n = *border_count;
  const int N = NC > 0 ? NC : L.n;
  const int stride = N + 1;
  const uint32_t P = static_cast<uint32_t>(L.w) * L.h;
  for (int t0 = blockIdx.x * kBlock; t0 < n; t0 += gridDim.x * kBlock) {
    const int t = t0 + threadIdx.x;
    EdgeSetFor<NC> edges;
    edges.clear();
    int i = 0, u = 0, v = 0;
    int* parent = L.parent;
    if (t < n) {
      const uint32_t g = border_list[t];
      const uint32_t frame = g / P;
      i = static_cast<int>(g - frame * P);
      v = i / L.w;
      u = i - v * L.w;
      const uint8_t* mask = L.mask + static_cast<size_t>(frame) * P;
      const float* pts = L.points + static_cast<size_t>(frame) * L.points_frame_stride * 8;
      parent = L.parent + static_cast<size_t>(frame) * P;
      const float z_here = __ldg(pts + static_cast<size_t>(i) * 8 + 2);
      auto probe = [&](int wx, int wy, float& z) {
        const int uu = u - N + wx, vv = v - N + wy;
        if (uu < 0 || vv < 0) return false;  // isInRange (uu <= u < w and vv <= v < h always hold)
        const int j = vv * L.w + uu;
        if (mask[j] == 0) return false;
        z = __ldg(pts + static_cast<size_t>(j) * 8 + 2);
        return true;
      };
      window_edges<NC>(N, z_here, L.depth_thr, probe, edges);
    }
    const int lx = u & (kTileW - 1), ly = v & (kTileH - 1);
    while (__any_sync(kFull, !edges.empty())) {
      if (!edges.empty()) {
        const int k = edges.pop();
        const int wy = k / stride, wx = k - wy * stride;
        const int dx = wx - N, dy = wy - N;
        if (lx + dx < 0 || ly + dy < 0) uf_union(parent, i, (v + dy) * L.w + (u + dx));  // other tile
      }
      __syncwarp();
    }
  }
}

// ---- cl_flatten -------------------------------------------------------------------------------------------------
// integrateConnectedClusters (:253-267) + the histogram of removeSmallClusters (:359-367)
__global__ void __launch_bounds__(kBlock)
    cl_flatten_kernel(const ClLaunch L, const uint32_t* __restrict__ list, const int* __restrict__ list_count) {
  const int n = *list_count;
  const uint32_t P = static_cast<uint32_t>(L.w) * L.h;
  const int lane = threadIdx.x & 31;
  for (int t0 = blockIdx.x * kBlock; t0 < n; t0 += gridDim.x * kBlock) {
    const int t = t0 + threadIdx.x;
    const bool active = t < n;
    const unsigned ballot = __ballot_sync(kFull, active);
    if (!active) continue;
    const uint32_t g = list[t];
    const uint32_t frame = g / P;
    int* parent = L.parent + static_cast<size_t>(frame) * P;
    const int root = uf_find(parent, static_cast<int>(g - frame * P));
    L.parent[g] = root;
    // warp-aggregated size count: lanes with the same root cell elect one leader
    const uint32_t cell = frame * P + static_cast<uint32_t>(root);
    const unsigned same = __match_any_sync(ballot, cell);
    if (lane == __ffs(same) - 1) atomicAdd(L.count + cell, __popc(same));
  }
}

// ---- cl_slots ---------------------------------------------------------------------------------------------------
// removeSmallClusters (:369-383): components below cluster_size are dropped; survivors get a dense slot, which is
// stored over count[root] (the size moves into the slot's ClusterStat).
__global__ void __launch_bounds__(kBlock)
    cl_slots_kernel(const ClLaunch L, const uint32_t* __restrict__ list, const int* __restrict__ list_count) {
  const int n = *list_count;
  const uint32_t P = static_cast<uint32_t>(L.w) * L.h;
  for (int t = blockIdx.x * kBlock + threadIdx.x; t < n; t += gridDim.x * kBlock) {
    const uint32_t g = list[t];
    const uint32_t frame = g / P;
    const int i = static_cast<int>(g - frame * P);
    if (L.parent[g] != i) continue;  // not a root
    const int cnt = L.count[g];
    int s = -1;
    if (cnt >= L.min_size) {
      s = atomicAdd(L.n_slots + frame, 1);
      if (s < L.max_clusters) {
        ClusterStat& st = L.stats[static_cast<size_t>(frame) * L.max_clusters + s];
        st.root = i;
        st.count = cnt;
        st.key = 0x7fffffff;
        st.rank = -1;
        st.mn[0] = st.mn[1] = st.mn[2] = 0xffffffffu;
        st.mx[0] = st.mx[1] = st.mx[2] = 0u;
        st.median_idx = 0x7fffffff;
      } else {
        s = -1;  // overflow: reported through frame_stats.status by cl_rank
      }
    }
    L.count[g] = s;
  }
}

// ---- cl_stats ---------------------------------------------------------------------------------------------------
// pcl::getMinMax3D over the members (cluster2MovingObject :151-152) and the reference-order key, reduced with warp
// shuffles per cluster before touching the per-cluster atomics.
__global__ void __launch_bounds__(kBlock)
    cl_stats_kernel(const ClLaunch L, const uint32_t* __restrict__ list, const int* __restrict__ list_count) {
  const int n = *list_count;
  const uint32_t P = static_cast<uint32_t>(L.w) * L.h;
  const int lane = threadIdx.x & 31;
  for (int t0 = blockIdx.x * kBlock; t0 < n; t0 += gridDim.x * kBlock) {
    const int t = t0 + threadIdx.x;
    int s = -1;
    uint32_t frame = 0;
    int i = 0;
    uint32_t g = 0;
    if (t < n) {
      g = list[t];
      frame = g / P;
      i = static_cast<int>(g - frame * P);
      s = L.count[frame * P + static_cast<uint32_t>(L.parent[g])];  // slot of the root (cl_slots)
    }
    const unsigned member = __ballot_sync(kFull, s >= 0);
    if (member == 0) continue;
    uint32_t ex_mn[3], ex_mx[3];
    int key = 0x7fffffff;
    long long group = -1 - lane;  // distinct for non-members so they never match each other
#pragma unroll
    for (int k = 0; k < 3; k++) {
      ex_mn[k] = 0xffffffffu;
      ex_mx[k] = 0u;
    }
    if (s >= 0) {
      const float* rec = L.points + (static_cast<size_t>(frame) * L.points_frame_stride + i) * 8;
      const float4 xyz = __ldg(reinterpret_cast<const float4*>(rec));
      const float c[3] = {xyz.x, xyz.y, xyz.z};
#pragma unroll
      for (int k = 0; k < 3; k++) {
        // NaN coordinates are skipped (the reference's running min/max is order dependent there; unpinned)
        if (!isnan(c[k])) ex_mn[k] = ex_mx[k] = float_to_ordered(c[k]);
      }
      if (L.flags[g] & 1) key = i;
      group = (static_cast<long long>(frame) << 32) | static_cast<unsigned>(s);
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
        ClusterStat& st = L.stats[static_cast<size_t>(frame) * L.max_clusters + s];
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
    s_base = n > 0 ? atomicAdd(work_count, n) : 0;
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
// irrelevant to every consumer: min/max, k-th largest speed).  |v| is evaluated here, as Vector3f::norm() with the
// oracle's summation order.
__global__ void __launch_bounds__(kBlock)
    cl_scatter_kernel(const ClLaunch L, const uint32_t* __restrict__ list, const int* __restrict__ list_count,
                      const int* __restrict__ member_offset, int* __restrict__ cursor,
                      uint32_t* __restrict__ member_speed, int* __restrict__ member_pixel) {
  const int n = *list_count;
  const uint32_t P = static_cast<uint32_t>(L.w) * L.h;
  const int lane = threadIdx.x & 31;
  for (int t0 = blockIdx.x * kBlock; t0 < n; t0 += gridDim.x * kBlock) {
    const int t = t0 + threadIdx.x;
    int s = -1;
    uint32_t frame = 0;
    int i = 0;
    if (t < n) {
      const uint32_t g = list[t];
      frame = g / P;
      i = static_cast<int>(g - frame * P);
      s = L.count[frame * P + static_cast<uint32_t>(L.parent[g])];
    }
    const unsigned member = __ballot_sync(kFull, s >= 0);
    if (member == 0) continue;
    const long long group = (s >= 0) ? ((static_cast<long long>(frame) << 32) | static_cast<unsigned>(s)) : (-1 - lane);
    const unsigned same = __match_any_sync(kFull, group);
    if (s >= 0) {
      const int leader = __ffs(same) - 1;
      const size_t cs = static_cast<size_t>(frame) * L.max_clusters + s;
      int base = 0;
      if (lane == leader) base = atomicAdd(cursor + cs, __popc(same));
      base = __shfl_sync(same, base, leader);
      const size_t pos = static_cast<size_t>(frame) * P + member_offset[cs] + base + __popc(same & ((1u << lane) - 1));
      const float4 vel = __ldg(reinterpret_cast<const float4*>(
          L.points + (static_cast<size_t>(frame) * L.points_frame_stride + i) * 8 + 4));
      const float speed = __fsqrt_rn(
          __fadd_rn(__fadd_rn(__fmul_rn(vel.x, vel.x), __fmul_rn(vel.y, vel.y)), __fmul_rn(vel.z, vel.z)));
      member_speed[pos] = __float_as_uint(speed);  // speeds are >= 0: the bit pattern is monotone
      member_pixel[pos] = i;
    }
  }
}

// ---- cl_objects -------------------------------------------------------------------------------------------------
// cluster2MovingObject (:147-184) for one cluster per CTA (persistent CTAs pull clusters from the work queue):
// bbox = max - min, centre = (min + max) / 2 (float), velocity = member with the (n/2)-th largest |v|.
// The k-th largest is found by a 4-pass MSB-first radix select over the cluster's member list with a 256-bin
// shared-memory histogram (warp-aggregated increments); lists up to kSmemMembers are staged in shared memory once.
// Among members tied at the selected speed the smallest raster index is taken (the reference's std::sort leaves
// ties unspecified).
constexpr int kObjBlock = 512;
constexpr int kSmemMembers = 10240;  // 40 KB

__global__ void __launch_bounds__(kObjBlock)
    cl_objects_kernel(const ClLaunch L, const int* __restrict__ work_count, const int2* __restrict__ work_items,
                      int* __restrict__ work_cursor, const int* __restrict__ member_offset,
                      const uint32_t* __restrict__ member_speed, const int* __restrict__ member_pixel) {
  __shared__ uint32_t s_speed[kSmemMembers];
  __shared__ unsigned hist[256];
  __shared__ int s_item;
  __shared__ uint32_t s_prefix;
  __shared__ uint32_t s_k;
  __shared__ int s_best;
  const size_t P = static_cast<size_t>(L.w) * L.h;
  const int n_work = *work_count;
  const int lane = threadIdx.x & 31;
  for (;;) {
    if (threadIdx.x == 0) s_item = atomicAdd(work_cursor, 1);
    __syncthreads();
    const int item = s_item;
    if (item >= n_work) return;
    const int2 w = work_items[item];
    const int frame = w.x, slot = w.y;
    ClusterStat& st = L.stats[static_cast<size_t>(frame) * L.max_clusters + slot];
    const int n = st.count;
    const size_t off = frame * P + member_offset[static_cast<size_t>(frame) * L.max_clusters + slot];
    const uint32_t* gspeeds = member_speed + off;
    const int* pixels = member_pixel + off;
    const bool staged = n <= kSmemMembers;
    if (staged)
      for (int j = threadIdx.x; j < n; j += kObjBlock) s_speed[j] = gspeeds[j];
    const uint32_t* speeds = staged ? s_speed : gspeeds;
    if (threadIdx.x == 0) {
      s_prefix = 0;
      s_k = static_cast<uint32_t>(n / 2);  // cluster.at(cluster.size() / 2) after the descending sort
      s_best = 0x7fffffff;
    }
    for (int pass = 0; pass < 4; pass++) {
      const int shift = 24 - 8 * pass;
      if (threadIdx.x < 256) hist[threadIdx.x] = 0;
      __syncthreads();
      const uint32_t prefix = s_prefix;
      const uint32_t himask = (pass == 0) ? 0u : (0xffffffffu << (shift + 8));
      for (int j0 = 0; j0 < n; j0 += kObjBlock) {
        const int j = j0 + threadIdx.x;
        int bin = -1;
        if (j < n) {
          const uint32_t b = speeds[j];
          if ((b & himask) == prefix) bin = static_cast<int>((b >> shift) & 0xff);
        }
        const unsigned same = __match_any_sync(kFull, bin);
        if (bin >= 0 && lane == __ffs(same) - 1) atomicAdd(&hist[bin], static_cast<unsigned>(__popc(same)));
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
    for (int j = threadIdx.x; j < n; j += kObjBlock)
      if (speeds[j] == target) best = min(best, pixels[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(kFull, best, o));
    if (lane == 0 && best != 0x7fffffff) atomicMin(&s_best, best);
    __syncthreads();
    if (threadIdx.x == 0) {
      const int id = st.rank;
      st.median_idx = s_best;
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
__global__ void __launch_bounds__(kBlock) cl_labels_kernel(const ClLaunch L) {
  const size_t P = static_cast<size_t>(L.w) * L.h;
  const size_t total = P * L.batch;
  for (size_t g = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; g < total;
       g += static_cast<size_t>(gridDim.x) * blockDim.x) {
    int label = -1;
    if (L.mask[g] != 0) {
      const size_t frame = g / P;
      const int s = L.count[frame * P + L.parent[g]];
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
  const size_t n_fs = static_cast<size_t>(L.batch);
  const size_t n_cs = static_cast<size_t>(L.batch) * L.max_clusters;
  {
    // zero n_slots / n_dynamic (adjacent, 2*batch), cursor (batch*max_clusters), and the four counters
    // work_count, work_cursor, list_count, border_count
    const size_t m = n_cs > 2 * n_fs ? n_cs : 2 * n_fs;
    cl_reset_kernel<<<static_cast<unsigned>((m + 255) / 256), 256, 0, s>>>(L.n_slots, 2 * n_fs, X.cursor, n_cs,
                                                                          X.work_count, 4);
  }
  int* list_count = X.work_count + 2;
  int* border_count = X.work_count + 3;
  uint32_t* border_list = reinterpret_cast<uint32_t*>(X.member_pixel);  // free until cl_scatter
  const dim3 tgrid((L.w + kTileW - 1) / kTileW, (L.h + kTileH - 1) / kTileH, L.batch);
  switch (L.n) {
    case 1: cl_tile_kernel<1><<<tgrid, kBlock, 0, s>>>(L, X.list, list_count, border_list, border_count); break;
    case 2: cl_tile_kernel<2><<<tgrid, kBlock, 0, s>>>(L, X.list, list_count, border_list, border_count); break;
    case 3: cl_tile_kernel<3><<<tgrid, kBlock, 0, s>>>(L, X.list, list_count, border_list, border_count); break;
    case 4: cl_tile_kernel<4><<<tgrid, kBlock, 0, s>>>(L, X.list, list_count, border_list, border_count); break;
    default: cl_tile_kernel<0><<<tgrid, kBlock, 0, s>>>(L, X.list, list_count, border_list, border_count); break;
  }
  switch (L.n) {
    case 1: cl_merge_kernel<1><<<kListGrid, kBlock, 0, s>>>(L, border_list, border_count); break;
    case 2: cl_merge_kernel<2><<<kListGrid, kBlock, 0, s>>>(L, border_list, border_count); break;
    case 3: cl_merge_kernel<3><<<kListGrid, kBlock, 0, s>>>(L, border_list, border_count); break;
    case 4: cl_merge_kernel<4><<<kListGrid, kBlock, 0, s>>>(L, border_list, border_count); break;
    default: cl_merge_kernel<0><<<kListGrid, kBlock, 0, s>>>(L, border_list, border_count); break;
  }
  cl_flatten_kernel<<<kListGrid, kBlock, 0, s>>>(L, X.list, list_count);
  cl_slots_kernel<<<kListGrid, kBlock, 0, s>>>(L, X.list, list_count);
  cl_stats_kernel<<<kListGrid, kBlock, 0, s>>>(L, X.list, list_count);
  cl_rank_kernel<<<L.batch, kBlock, static_cast<size_t>(L.max_clusters) * sizeof(int), s>>>(L, X.work_count, X.work_items,
                                                                                          X.member_offset);
  cl_scatter_kernel<<<kListGrid, kBlock, 0, s>>>(L, X.list, list_count, X.member_offset, X.cursor, X.member_speed,
                                                 X.member_pixel);
  cl_objects_kernel<<<148 * 2, kObjBlock, 0, s>>>(L, X.work_count, X.work_items, X.work_cursor, X.member_offset,
                                                  X.member_speed, X.member_pixel);
  if (L.labels != nullptr) {
    size_t lb = (total + kBlock - 1) / kBlock;
    if (lb > 148 * 16) lb = 148 * 16;
    cl_labels_kernel<<<static_cast<unsigned>(lb), kBlock, 0, s>>>(L);
  }
  if (n_launches) *n_launches = 9 + (L.labels != nullptr ? 1 : 0);
  return cudaGetLastError();
}

}  // namespace mod
