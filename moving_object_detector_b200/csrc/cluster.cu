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
// Pipeline (batched over frames); only cl_tile, cl_scatter (and the optional cl_labels) touch every dynamic pixel:
//   cl_tile      one CTA per 32x32 tile.  Mask words + z of the tile and its N-pixel top/left halo are staged in shared
//                memory and every dynamic pixel evaluates its window once.  N in 1..4 (cl_tile_fast_kernel): pixels
//                link to one accepted neighbour (a forest, no atomics), the forest is flattened and the remaining
//                accepted edges are settled by label compares, with unions only where labels differ.  Other N
//                (cl_tile_kernel; only neighbor_distance 0 still takes it): implied-edge elimination, then per-edge unions
//                in a shared-memory union-find.  N in 5..10 (cl_tile_wide_kernel, cluster_wide.cuh): the halo pixels are
//                nodes of the tile's union-find, labels are compared before depth gates are evaluated, and every halo
//                pixel that joined a component yields one (halo pixel, tile root) pair for cl_merge_pairs.
//                The tile then publishes parent[pixel] = tile root, per-tile-root partial statistics (size,
//                reference-order key, min/max of x,y,z; warp-shuffle reduced, then atomics), and three compact lists:
//                dynamic pixels (with |v|), pixels with an edge leaving the tile, tile roots.
//   cl_merge     border list: lock-free union (atomicMin + path halving) of the edges that cross a tile border,
//                de-duplicated per warp to one union per distinct (tile root, tile root) pair.
//   cl_roots     tile-root list: parent[r] = final root R; partial statistics of r are merged into R
//   cl_slots     final roots with size >= cluster_size claim a dense survivor slot (removeSmallClusters)
//   cl_rank      per frame: final ids by key rank, member-list offsets, work queue for cl_objects
//   cl_scatter   survivors' members -> contiguous per-cluster lists (speed bits, pixel index)
//   cl_objects   one CTA per cluster: 3-pass (11/11/10 bit) radix select of the (n/2)-th largest |v|, MovingObject
//   cl_labels    optional dense label image (cluster_map_)

#include <cooperative_groups.h>
#include <math_constants.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace mod {

namespace {

constexpr int kBlock = 256;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kListGrid = 148 * 8;  // grid-stride kernels over the dynamic-pixel list: a multiple of the SM count

// ---- helpers ---------------------------------------------------------------------------------------------------

// Programmatic dependent launch (sm_90+): the clusterer is a chain of nine short kernels, each consuming what the
// previous one wrote.  Every kernel of the chain starts with pdl_enter(): it lets the NEXT kernel's CTAs be scheduled as
// soon as SM resources free up (their launch latency and prologue overlap this kernel's tail) and then waits until the
// PREVIOUS kernel has completed and its writes are visible -- so nothing is read early, and completion stays
// transitive along the chain.  Kernels launched without the attribute (launch_chain below with CL_PDL = 0, or <<<>>>)
// see both instructions as no-ops.
#ifndef CL_PDL
#define CL_PDL 1
#endif
__device__ __forceinline__ void pdl_enter() {
#if CL_PDL
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

// Callers may drop the return value as they do with <<<>>>: a failed launch is also what cudaGetLastError() reports at
// the end of launch_cluster.
template <typename... KArgs, typename... Args>
cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL_PDL ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

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

// Read-only find (no path compression).  cl_roots must use this one: a path-halving store issued from a stale read
// may land AFTER another thread has stored that node's final root and silently replace it by an intermediate
// ancestor, which breaks the "pixel -> tile root -> final root" two-hop invariant the later kernels rely on.
__device__ __forceinline__ int uf_find_readonly(const int* parent, int i) {
  for (;;) {
    const int p = __ldcg(parent + i);
    if (p == i) return i;
    i = p;
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
// Read-only variant for the flatten pass (D1): there every thread stores the ROOT of its own pixel, and a
// path-halving store from a stale read could overwrite that root with an intermediate ancestor afterwards.
__device__ __forceinline__ int smem_find_readonly(const volatile int* lab, int i) {
  for (;;) {
    const int p = lab[i];
    if (p == i) return i;
    i = p;
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

// Appends `flag` lanes' `value` to a shared-memory list (order inside a warp preserved, warps unordered); returns the
// position written (or -1).
__device__ __forceinline__ int smem_append(bool flag, uint16_t value, uint16_t* list, int* counter) {
  const unsigned ballot = __ballot_sync(kFull, flag);
  if (ballot == 0) return -1;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == 0) base = atomicAdd(counter, __popc(ballot));
  base = __shfl_sync(kFull, base, 0);
  const int pos = base + __popc(ballot & ((1u << lane) - 1));
  if (flag) list[pos] = value;
  return flag ? pos : -1;
}

// Flag bits of a staged pixel
constexpr uint8_t kDyn = 1;   // dynamic
constexpr uint8_t kLeft = 2;  // edge to the left neighbour accepted  (both dynamic, depth gate passes)
constexpr uint8_t kUp = 4;    // edge to the upper neighbour accepted

// Generic tile kernel (runtime N; since round 2 only neighbor_distance 0 reaches it): window_edges with the implied-edge
// elimination, then per-edge unions in a shared-memory union-find.  neighbor_distance 1..4 (the default is 4) takes
// cl_tile_fast_kernel, 5..10 cl_tile_wide_kernel.
__global__ void __launch_bounds__(kBlock)
    cl_tile_kernel(const ClLaunch L, uint32_t* __restrict__ list, uint32_t* __restrict__ list_speed,
                   uint32_t* __restrict__ border_list, uint32_t* __restrict__ root_list,
                   int* __restrict__ counters /* list, border, root */) {
  pdl_enter();
  constexpr int NC = 0;
  constexpr int kMaxN = 10;
  constexpr int HN = kMaxN;        // halo capacity
  constexpr int SW = kTileW + HN;  // shared-memory row pitch
  const int N = L.n;
  const int stride = N + 1;
  __shared__ __align__(16) uint8_t s_m[(kTileH + HN) * SW];
  __shared__ float s_z[(kTileH + HN) * SW];
  __shared__ __align__(16) int s_lab[kTileW * kTileH];
  __shared__ uint8_t s_up[kTileW * kTileH];  // pixel has an accepted up-left neighbour
  __shared__ uint16_t s_dl[kTileW * kTileH];
  __shared__ uint16_t s_bl[kTileW * kTileH];
  __shared__ uint16_t s_rl[kTileW * kTileH];
  __shared__ int s_cnt, s_bcnt, s_rcnt, s_base, s_bbase, s_rbase;
  const int frame = blockIdx.z;
  const int tx0 = blockIdx.x * kTileW, ty0 = blockIdx.y * kTileH;
  const size_t P = static_cast<size_t>(L.w) * L.h;
  const size_t fb = frame * P;
  const uint8_t* mask = L.mask + fb;
  const float* pts = L.points + frame * L.points_frame_stride * 8;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) s_cnt = s_bcnt = s_rcnt = 0;

  // A: the 32x32 core, one 4-pixel mask word per thread
  const int row = threadIdx.x >> 3, wq = threadIdx.x & 7;
  const int y = ty0 + row, x0 = tx0 + wq * 4;
  uint32_t word = 0;
  if (y < L.h && x0 < L.w) {
    const uint8_t* src = mask + static_cast<size_t>(y) * L.w + x0;
    if (x0 + 3 < L.w && (reinterpret_cast<uintptr_t>(src) & 3u) == 0) {
      word = __ldg(reinterpret_cast<const uint32_t*>(src));
    } else {
      for (int j = 0; j < 4 && x0 + j < L.w; j++) word |= static_cast<uint32_t>(src[j]) << (8 * j);
    }
  }
  if (!__syncthreads_or(word != 0)) return;  // no dynamic pixel in this tile

  const int c0 = row * kTileW + wq * 4;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const bool dyn = ((word >> (8 * j)) & 0xffu) != 0;
    const int si = (row + N) * SW + N + wq * 4 + j;
    s_m[si] = dyn ? kDyn : 0;
    if (dyn) s_z[si] = __ldg(pts + (static_cast<size_t>(y) * L.w + x0 + j) * 8 + 2);
  }
  // halo: N rows above (full width 32+N) and N columns to the left of the core
  const int top = N * (kTileW + N);
  const int n_halo = top + kTileH * N;
  for (int k = threadIdx.x; k < n_halo; k += kBlock) {
    int sy, sx;
    if (k < top) {
      sy = k / (kTileW + N);
      sx = k - sy * (kTileW + N);
    } else {
      const int k2 = k - top;
      sy = N + k2 / N;
      sx = k2 - (k2 / N) * N;
    }
    const int yy = ty0 - N + sy, xx = tx0 - N + sx;
    uint8_t m = 0;
    if (xx >= 0 && yy >= 0 && xx < L.w && yy < L.h) m = mask[static_cast<size_t>(yy) * L.w + xx] ? kDyn : 0;
    s_m[sy * SW + sx] = m;
    if (m) s_z[sy * SW + sx] = __ldg(pts + (static_cast<size_t>(yy) * L.w + xx) * 8 + 2);
  }
  __syncthreads();

  // B: initial labels and the tile-local list of dynamic pixels
  {
    int lab4[4];
#pragma unroll
    for (int j = 0; j < 4; j++) lab4[j] = ((word >> (8 * j)) & 0xffu) ? c0 + j : -1;
    *reinterpret_cast<int4*>(&s_lab[c0]) = make_int4(lab4[0], lab4[1], lab4[2], lab4[3]);
    // one warp appends one row at a time: 32 consecutive list entries are then (mostly) 32 consecutive pixels of a
    // row, which keeps the shared-memory accesses of the later phases free of bank conflicts
    for (int it = 0; it < kTileH / (kBlock / 32); it++) {
      const int r = it * (kBlock / 32) + (threadIdx.x >> 5);
      smem_append((s_m[(r + N) * SW + N + lane] & kDyn) != 0, static_cast<uint16_t>(r * kTileW + lane), s_dl, &s_cnt);
    }
  }
  __syncthreads();
  const int nd = s_cnt;

  // C: edge sets (SIMD), then tile-local unions in shared memory; edges leaving the tile go to cl_merge
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
      const int sa = (ly + N) * SW + lx + N;
      const float z_here = s_z[sa];
      auto probe = [&](int wx, int wy, float& z) {
        const int si = (ly + wy) * SW + lx + wx;
        if (!s_m[si]) return false;
        z = s_z[si];
        return true;
      };
      s_up[c] = window_edges<NC>(N, z_here, L.depth_thr, probe, edges) ? 1 : 0;
    }
    while (__any_sync(kFull, !edges.empty())) {  // warp-uniform loop: lanes stay converged between unions
      if (!edges.empty()) {
        const int k = edges.pop();
        const int wy = k / stride, wx = k - wy * stride;
        const int nx = lx - N + wx, ny = ly - N + wy;
        if (nx >= 0 && ny >= 0)
          smem_union(s_lab, c, ny * kTileW + nx);
        else
          crosses = true;  // the up-left endpoint lies in another tile: cl_merge re-evaluates this pixel's window
      }
      __syncwarp();
    }
    smem_append(crosses, static_cast<uint16_t>(c), s_bl, &s_bcnt);
  }
  __syncthreads();

  // D1: flatten the tile-local forest; roots initialise their partial record
  for (int e0 = 0; e0 < nd; e0 += kBlock) {
    const int e = e0 + threadIdx.x;
    bool is_root = false;
    int c = 0;
    if (e < nd) {
      c = s_dl[e];
      const int r = smem_find_readonly(s_lab, c);  // no halving stores here: see smem_find_readonly
      s_lab[c] = r;
      is_root = (r == c);
      if (is_root) {
        const size_t gi = fb + static_cast<size_t>(ty0 + (c >> 5)) * L.w + tx0 + (c & 31);
        L.count[gi] = 0;
        TileStat init;
        init.key = 0x7fffffff;
        init.mn[0] = init.mn[1] = init.mn[2] = 0xffffffffu;
        init.mx[0] = init.mx[1] = init.mx[2] = 0u;
        init.pad = 0;
        L.tstat[gi] = init;
      }
    }
    smem_append(is_root, static_cast<uint16_t>(c), s_rl, &s_rcnt);
  }
  __syncthreads();
  const int nb = s_bcnt, nr = s_rcnt;
  if (threadIdx.x == 0) {
    s_base = atomicAdd(counters + 0, nd);
    s_bbase = nb ? atomicAdd(counters + 1, nb) : 0;
    s_rbase = atomicAdd(counters + 2, nr);
    atomicAdd(L.n_dynamic + frame, nd);
  }
  __syncthreads();

  // D2: publish parent = tile root, list entries, and the per-root partial statistics (pcl::getMinMax3D operands,
  // size, reference-order key), reduced per tile root inside each warp (redux.sync) before the atomics
  for (int e0 = 0; e0 < nd; e0 += kBlock) {
    const int e = e0 + threadIdx.x;
    int r = -1 - lane;  // distinct for idle lanes
    uint32_t ex_mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, ex_mx[3] = {0u, 0u, 0u};
    int key = 0x7fffffff;
    if (e < nd) {
      const int c = s_dl[e];
      r = s_lab[c];
      const int i = (ty0 + (c >> 5)) * L.w + tx0 + (c & 31);
      L.parent[fb + i] = (ty0 + (r >> 5)) * L.w + tx0 + (r & 31);
      list[s_base + e] = static_cast<uint32_t>(fb + i);
      const float4 xyz = __ldg(reinterpret_cast<const float4*>(pts + static_cast<size_t>(i) * 8));
      // |v| (Vector3f::norm, the oracle's summation order) rides along with the list entry: the record's second half
      // sits in the same 32-byte sector, and cl_scatter then reads 4 bytes per member instead of a sector
      const float4 vel = __ldg(reinterpret_cast<const float4*>(pts + static_cast<size_t>(i) * 8 + 4));
      list_speed[s_base + e] = __float_as_uint(__fsqrt_rn(
          __fadd_rn(__fadd_rn(__fmul_rn(vel.x, vel.x), __fmul_rn(vel.y, vel.y)), __fmul_rn(vel.z, vel.z))));
      const float cc[3] = {xyz.x, xyz.y, xyz.z};
#pragma unroll
      for (int k = 0; k < 3; k++)  // NaN coordinates are skipped (order dependent in the reference; unpinned)
        if (!isnan(cc[k])) ex_mn[k] = ex_mx[k] = float_to_ordered(cc[k]);
      if (s_up[c]) key = i;
    }
    unsigned todo = __ballot_sync(kFull, r >= 0);
    while (todo) {  // peel one tile root at a time; usually one or two per warp
      const int leader = __ffs(todo) - 1;
      const int lr = __shfl_sync(kFull, r, leader);
      const bool in = (r == lr);
      const unsigned grp = __ballot_sync(kFull, in);
      if (in) {
        uint32_t mn[3], mx[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
          mn[k] = __reduce_min_sync(grp, ex_mn[k]);
          mx[k] = __reduce_max_sync(grp, ex_mx[k]);
        }
        const int kk = __reduce_min_sync(grp, key);
        if (lane == leader) {
          const size_t gr = fb + static_cast<size_t>(ty0 + (lr >> 5)) * L.w + tx0 + (lr & 31);
          atomicAdd(L.count + gr, __popc(grp));
          TileStat& st = L.tstat[gr];
#pragma unroll
          for (int k = 0; k < 3; k++) {
            if (mn[k] != 0xffffffffu) atomicMin(&st.mn[k], mn[k]);
            if (mx[k] != 0u) atomicMax(&st.mx[k], mx[k]);
          }
          if (kk != 0x7fffffff) atomicMin(&st.key, kk);
        }
      }
      todo &= ~grp;
    }
  }
  for (int e = threadIdx.x; e < nb; e += kBlock) {
    const int c = s_bl[e];
    border_list[s_bbase + e] = static_cast<uint32_t>(fb + static_cast<size_t>(ty0 + (c >> 5)) * L.w + tx0 + (c & 31));
  }
  for (int e = threadIdx.x; e < nr; e += kBlock) {
    const int c = s_rl[e];
    root_list[s_rbase + e] = static_cast<uint32_t>(fb + static_cast<size_t>(ty0 + (c >> 5)) * L.w + tx0 + (c & 31));
  }
}

// ---- cl_tile, fast path (N in 1..4) --------------------------------------------------------------------------------
// Same contract as cl_tile_kernel (same outputs, same lists), different tile-local connectivity.  With noisy stereo
// depth roughly every second depth gate between neighbours fails (measured on the 720p benchmark sequence: 48 % of the
// unit edges accepted, 11 of 24 window cells accepted per dynamic pixel), so the window graph is a dense random graph
// on which per-edge union-find operations are almost all redundant.  This kernel therefore
//   C1  evaluates the window once per pixel (non-dynamic cells are staged as z = +inf, so a cell costs one shared load,
//       one subtract and one compare) and LINKS the pixel to its raster-smallest accepted in-tile neighbour: a parent
//       pointer to an earlier pixel, no atomics, no finds -- a spanning forest with a handful of trees per tile;
//   F   flattens the forest (read-only finds);
//   C2a compares the flattened label of the first remaining accepted in-tile neighbour with the pixel's own and
//       unions the pair if they differ, de-duplicated per warp: the trees interleave, so this merges them;
//   F   flattens again;
//   C2b compares the labels of ALL remaining accepted in-tile neighbours (one shared load per edge) and unions
//       whatever still differs (rare).
// Every accepted in-tile edge is either the link, found equal-labelled (endpoints in one set for certain: labels
// only ever move towards the root), or unioned -- so the partition equals the one the reference's scan produces.
// Pixels within N of the tile's top / left edge additionally run the implied-edge elimination of window_edges on their
// cross-tile edges (they are listed separately so that interior warps do not pay for it).
// Eight resident CTAs per SM (32 registers per thread, ~80 bytes of spill; 22 KB of shared memory at N = 4): the kernel
// is a chain of short phases separated by barriers, and the barrier stall was its largest one at six CTAs per SM
// (ncu, r01): 418 -> 390 us per 63 frames of 1280x720.
#ifndef CL_TILE_MIN_BLOCKS
#define CL_TILE_MIN_BLOCKS 8
#endif
#ifndef CL_TILE_CMP_ROWS
#define CL_TILE_CMP_ROWS 4  // pixel rows per pass of the label-compare stencil (4: 0.558 ms, 2: 0.565, 1: 0.564, list walk 0.568)
#endif
template <int NC>
__global__ void __launch_bounds__(kBlock, CL_TILE_MIN_BLOCKS)
    cl_tile_fast_kernel(const ClLaunch L, uint32_t* __restrict__ list, uint32_t* __restrict__ list_speed,
                        uint32_t* __restrict__ border_list, uint32_t* __restrict__ border_mask,
                        uint32_t* __restrict__ root_list, int* __restrict__ counters /* list, border, root */) {
  pdl_enter();
  static_assert(NC >= 1 && NC <= 4, "fast path covers neighbor_distance 1..4");
  constexpr int S = NC + 1;
  constexpr int SW = kTileW + NC;  // shared-memory row pitch
  constexpr int kA = NC * S + NC, kL = kA - 1, kU = kA - S;
  constexpr int kPixels = kTileW * kTileH;
  // A core pixel is identified by c = ly * SW + lx: the staged arrays' pitch, so that the window cell (wx, wy) of c is
  // staged at c + wy * SW + wx, and so that neither 32 consecutive pixels of a row nor the NC leftmost pixels of
  // consecutive rows (the two orders the lists below come in) collide on a shared-memory bank.
  constexpr int kIds = kTileH * SW;
  constexpr int kCore = NC * SW + NC;  // staged index of core pixel c is c + kCore
  __shared__ __align__(16) uint8_t s_m[(kTileH + NC) * SW];
  __shared__ float s_z[(kTileH + NC) * SW];
  __shared__ int s_lab[kIds];
  __shared__ uint8_t s_up[kIds];      // pixel has an accepted up-left neighbour
  __shared__ uint32_t s_E[kIds];      // accepted in-tile window cells minus the link
  __shared__ uint16_t s_dl[kPixels];  // interior pixels from the front, edge pixels from the back
  constexpr int kEdgePixels = kPixels - (kTileW - NC) * (kTileH - NC);  // pixels within NC of the top / left tile edge
  __shared__ uint16_t s_bl[kEdgePixels];
  __shared__ uint32_t s_bx[kEdgePixels];
  __shared__ uint16_t s_rl[kPixels];
  __shared__ int s_ni, s_ne, s_bcnt, s_rcnt, s_base, s_bbase, s_rbase, s_first;
  const int frame = blockIdx.z;
  const int tx0 = blockIdx.x * kTileW, ty0 = blockIdx.y * kTileH;
  const size_t P = static_cast<size_t>(L.w) * L.h;
  const size_t fb = frame * P;
  const uint8_t* mask = L.mask + fb;
  const float* pts = L.points + frame * L.points_frame_stride * 8;
  const int lane = threadIdx.x & 31;
  const float thr = L.depth_thr;
  const bool thr_ok = thr < CUDART_INF_F;  // false for NaN too: then the +inf staging cannot stand in for "not dynamic"

  if (threadIdx.x == 0) s_ni = s_ne = s_bcnt = s_rcnt = 0;

  // A: the 32x32 core, one 4-pixel mask word per thread
  const int row = threadIdx.x >> 3, wq = threadIdx.x & 7;
  const int y = ty0 + row, x0 = tx0 + wq * 4;
  uint32_t word = 0;
  if (y < L.h && x0 < L.w) {
    const uint8_t* src = mask + static_cast<size_t>(y) * L.w + x0;
    if (x0 + 3 < L.w && (reinterpret_cast<uintptr_t>(src) & 3u) == 0) {
      word = __ldg(reinterpret_cast<const uint32_t*>(src));
    } else {
      for (int j = 0; j < 4 && x0 + j < L.w; j++) word |= static_cast<uint32_t>(src[j]) << (8 * j);
    }
  }
  if (!__syncthreads_or(word != 0)) return;  // no dynamic pixel in this tile

#pragma unroll
  for (int j = 0; j < 4; j++) {
    const bool dyn = ((word >> (8 * j)) & 0xffu) != 0;
    const int si = (row + NC) * SW + NC + wq * 4 + j;
    s_m[si] = dyn ? kDyn : 0;
    s_z[si] = dyn ? __ldg(pts + (static_cast<size_t>(y) * L.w + x0 + j) * 8 + 2) : CUDART_INF_F;
  }
  // halo: NC rows above (full width 32+NC) and NC columns to the left of the core
  constexpr int top = NC * (kTileW + NC);
  constexpr int n_halo = top + kTileH * NC;
  for (int k = threadIdx.x; k < n_halo; k += kBlock) {
    int sy, sx;
    if (k < top) {
      sy = k / (kTileW + NC);
      sx = k - sy * (kTileW + NC);
    } else {
      const int k2 = k - top;
      sy = NC + k2 / NC;
      sx = k2 - (k2 / NC) * NC;
    }
    const int yy = ty0 - NC + sy, xx = tx0 - NC + sx;
    uint8_t m = 0;
    if (xx >= 0 && yy >= 0 && xx < L.w && yy < L.h) m = mask[static_cast<size_t>(yy) * L.w + xx] ? kDyn : 0;
    s_m[sy * SW + sx] = m;
    s_z[sy * SW + sx] = m ? __ldg(pts + (static_cast<size_t>(yy) * L.w + xx) * 8 + 2) : CUDART_INF_F;
  }
  __syncthreads();

  // A2: left / up edge bits of the staged dynamic pixels in the halo band and the first NC core rows / columns:
  // only the edge pixels' elimination reads them (flags of window cells, which lie at most NC inside the core)
  {
    constexpr int span = kTileW + NC;
    constexpr int band = 2 * NC;  // staged rows / columns 0 .. 2*NC-1 hold every cell an edge pixel's window can touch
    constexpr int n_band = band * span + (span - band) * band;
    uint8_t add[(n_band + kBlock - 1) / kBlock];
#pragma unroll
    for (int it = 0; it < (n_band + kBlock - 1) / kBlock; it++) {
      const int k = it * kBlock + threadIdx.x;
      add[it] = 0;
      if (k < n_band) {
        int sy, sx;
        if (k < band * span) {
          sy = k / span;
          sx = k - sy * span;
        } else {
          const int k2 = k - band * span;
          sy = band + k2 / band;
          sx = k2 - (k2 / band) * band;
        }
        const int si = sy * SW + sx;
        if (s_m[si]) {
          const float z = s_z[si];
          if (sx > 0 && s_m[si - 1] && depth_gate(z, s_z[si - 1], thr)) add[it] |= kLeft;
          if (sy > 0 && s_m[si - SW] && depth_gate(z, s_z[si - SW], thr)) add[it] |= kUp;
        }
      }
    }
    __syncthreads();  // all dynamic bits read before any flag byte changes
#pragma unroll
    for (int it = 0; it < (n_band + kBlock - 1) / kBlock; it++) {
      const int k = it * kBlock + threadIdx.x;
      if (k < n_band && add[it]) {
        int sy, sx;
        if (k < band * span) {
          sy = k / span;
          sx = k - sy * span;
        } else {
          const int k2 = k - band * span;
          sy = band + k2 / band;
          sx = k2 - (k2 / band) * band;
        }
        s_m[sy * SW + sx] |= add[it];
      }
    }
  }

  __syncthreads();  // flags complete
  // B: the tile-local lists of dynamic pixels: interior pixels, and pixels within NC of the top / left tile edge.
  // One warp appends one row at a time, so 32 consecutive list entries are (mostly) 32 consecutive pixels of a row.
  {
    const int warp = threadIdx.x >> 5;
#pragma unroll
    for (int it = 0; it < kTileH / (kBlock / 32); it++) {
      const int r = it * (kBlock / 32) + warp;
      const bool dyn = (s_m[(r + NC) * SW + NC + lane] & kDyn) != 0;
      const bool edge = r < NC || lane < NC;
      const uint16_t c = static_cast<uint16_t>(r * SW + lane);
      smem_append(dyn && !edge, c, s_dl, &s_ni);
      const unsigned ballot = __ballot_sync(kFull, dyn && edge);
      if (ballot) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_ne, __popc(ballot));
        base = __shfl_sync(kFull, base, 0);
        if (dyn && edge) s_dl[kPixels - 1 - (base + __popc(ballot & ((1u << lane) - 1)))] = c;
      }
    }
  }
  __syncthreads();
  const int ni = s_ni, ne = s_ne, nd = ni + ne;
  {
    // move the edge entries right behind the interior ones (source and destination may overlap: read, barrier, write)
    uint16_t moved[(kPixels - (kTileW - NC) * (kTileH - NC) + kBlock - 1) / kBlock];
#pragma unroll
    for (int it = 0; it < static_cast<int>(sizeof(moved) / sizeof(moved[0])); it++) {
      const int j = it * kBlock + threadIdx.x;
      moved[it] = j < ne ? s_dl[kPixels - 1 - j] : 0;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < static_cast<int>(sizeof(moved) / sizeof(moved[0])); it++) {
      const int j = it * kBlock + threadIdx.x;
      if (j < ne) s_dl[ni + j] = moved[it];
    }
    __syncthreads();
  }
  auto pixel_at = [&](int e) -> int { return s_dl[e]; };

  // window evaluation shared by both C1 loops: accepted cells of pixel c (window coordinates, bit k = wy*S+wx)
  auto accepted_cells = [&](int c, float z_here) -> uint32_t {
    uint32_t acc = 0;
    if (thr_ok && fabsf(z_here) < CUDART_INF_F) {
#pragma unroll
      for (int wy = 0; wy < S; wy++) {
#pragma unroll
        for (int wx = 0; wx < S; wx++) {
          if (wy == NC && wx == NC) continue;
          // a non-dynamic cell holds +inf: |finite - inf| = inf > thr rejects it
          if (!(fabsf(__fsub_rn(z_here, s_z[c + wy * SW + wx])) > thr)) acc |= 1u << (wy * S + wx);
        }
      }
    } else {  // z_here not finite (foreign clouds) or no usable threshold: test the dynamic flag explicitly
#pragma unroll
      for (int wy = 0; wy < S; wy++) {
#pragma unroll
        for (int wx = 0; wx < S; wx++) {
          if (wy == NC && wx == NC) continue;
          const int si = c + wy * SW + wx;
          if ((s_m[si] & kDyn) && depth_gate(z_here, s_z[si], thr)) acc |= 1u << (wy * S + wx);
        }
      }
    }
    return acc;
  };
  auto cell_offset = [](int k) -> int { return (k / S - NC) * SW + (k % S - NC); };

  // C1, interior pixels: every window cell lies inside the tile
  for (int e = threadIdx.x; e < ni; e += kBlock) {
    const int c = s_dl[e];
    const uint32_t acc = accepted_cells(c, s_z[c + kCore]);
    s_up[c] = acc != 0;
    s_lab[c] = acc ? c + cell_offset(__ffs(acc) - 1) : c;  // lowest bit = raster-smallest neighbour
    s_E[c] = acc & (acc - 1);
  }
  // C1, edge pixels: cross-tile edges go to cl_merge after the implied-edge elimination
  for (int e0 = 0; e0 < ne; e0 += kBlock) {
    const int e = e0 + threadIdx.x;
    uint32_t cross = 0;
    int c = 0;
    if (e < ne) {
      c = pixel_at(ni + e);
      const int ly = c / SW, lx = c - ly * SW;
      const int sa = c + kCore;
      const uint32_t acc = accepted_cells(c, s_z[sa]);
      uint32_t lm = 0, um = 0;
#pragma unroll
      for (int wy = 0; wy < S; wy++) {
#pragma unroll
        for (int wx = 0; wx < S; wx++) {
          if (wy == NC && wx == NC) continue;
          const uint32_t f = s_m[c + wy * SW + wx];
          lm |= ((f >> 1) & 1u) << (wy * S + wx);
          um |= ((f >> 2) & 1u) << (wy * S + wx);
        }
      }
      // cells whose right neighbour is still inside the window (wx < N)
      uint32_t col_ok = 0;
#pragma unroll
      for (int wy = 0; wy < S; wy++) col_ok |= ((1u << NC) - 1u) << (wy * S);
      uint32_t E = acc & ~(((acc & lm) >> 1) & col_ok) & ~((acc & um) >> S);
      // the vertical unit edge is dropped when the 4-cycle p, p-(1,0), up(p)-(1,0), up(p) is closed by accepted unit
      // edges: its endpoints are then connected through the up-edge one pixel to the left (induction on x)
      const uint32_t own = s_m[sa];
      if ((own & kLeft) && ((um >> kL) & 1u) && ((lm >> kU) & 1u)) E &= ~(1u << kU);
      // a cell (wx,wy) leaves the tile iff lx-N+wx < 0 or ly-N+wy < 0
      uint32_t inside = 0;
#pragma unroll
      for (int wy = 0; wy < S; wy++) {
        if (ly - NC + wy < 0) continue;
        const int first = max(NC - lx, 0);  // wx >= first stays inside
        if (first < S) inside |= ((((1u << S) - 1u) >> first) << first) << (wy * S);
      }
      cross = E & ~inside;
      const uint32_t local = acc & inside;
      s_up[c] = acc != 0;
      s_lab[c] = local ? c + cell_offset(__ffs(local) - 1) : c;
      s_E[c] = local & (local - 1);
    }
    const int pos = smem_append(cross != 0, static_cast<uint16_t>(c), s_bl, &s_bcnt);
    if (pos >= 0) s_bx[pos] = cross;
  }
  __syncthreads();

  // labels of the accepted in-tile neighbours that differ from the pixel's own flattened label
  auto differing = [&](int c, int my, uint32_t A) -> uint32_t {
    uint32_t diff = 0;
#pragma unroll
    for (int k = 0; k < S * S - 1; k++)
      if ((A >> k) & 1u) diff |= (s_lab[c + cell_offset(k)] != my) ? (1u << k) : 0u;
    return diff;
  };

  for (int e = threadIdx.x; e < nd; e += kBlock) {  // F
    const int c = pixel_at(e);
    s_lab[c] = smem_find_readonly(s_lab, c);
  }
  __syncthreads();
  for (int e0 = 0; e0 < nd; e0 += kBlock) {  // C2a: one union per pixel, one lane per distinct (root, root) pair
    const int e = e0 + threadIdx.x;
    int my = -1, other = -1;
    if (e < nd) {
      const int c = pixel_at(e);
      my = s_lab[c];
      // only the first remaining accepted neighbour is looked at: the trees of the link forest interleave, so most
      // pixels see another tree there, and any pair this round misses is settled by C2b
      const uint32_t A = s_E[c];
      if (A) {
        const int lb = s_lab[c + cell_offset(__ffs(A) - 1)];
        if (lb != my) other = lb;
      }
    }
    const unsigned has = __ballot_sync(kFull, other >= 0);
    if (other >= 0) {
      const int lo = min(my, other), hi = max(my, other);
      const unsigned same = __match_any_sync(has, (lo << 11) | hi);
      if (lane == __ffs(same) - 1) smem_union(s_lab, lo, hi);
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < nd; e += kBlock) {  // F
    const int c = pixel_at(e);
    const int r = smem_find_readonly(s_lab, c);
    s_lab[c] = r;
    if (e == 0) s_first = r;
  }
  __syncthreads();
#ifndef CL_TILE_NO_UNIFORM_SKIP
  // One component in the tile (the usual case inside a blob): every accepted in-tile edge joins equal labels, so
  // the compare round has nothing to find.
  bool one_label = true;
  {
    const int first = s_first;
    for (int e = threadIdx.x; e < nd; e += kBlock) one_label = one_label && (s_lab[pixel_at(e)] == first);
  }
  const bool skip_compare = __syncthreads_and(one_label);
#else
  const bool skip_compare = false;
#endif
  bool unioned = false;
#ifdef CL_TILE_LIST_COMPARE
  for (int e0 = 0; e0 < (skip_compare ? 0 : nd); e0 += kBlock) {  // C2b: whatever still differs
    const int e = e0 + threadIdx.x;
    int c = 0;
    uint32_t diff = 0;
    if (e < nd) {
      c = pixel_at(e);
      diff = differing(c, s_lab[c], s_E[c]);
    }
    unioned = unioned || diff != 0;
    while (__any_sync(kFull, diff != 0)) {
      if (diff) {
        const int k = __ffs(diff) - 1;
        diff &= diff - 1;
        smem_union(s_lab, c, c + cell_offset(k));
      }
      __syncwarp();
    }
  }
#else
  // C2b: whatever still differs.  This pass is a stencil over the label image, not a list walk: a warp takes the tile's
  // rows kCmpRows at a time with one COLUMN per lane, holds the label rows it needs in registers, and obtains the
  // label of the neighbour at (dx, dy) as row -dy shifted by dx lanes (one warp shuffle per label row and dx, shared by
  // the kCmpRows pixel rows), so an accepted edge costs a compare instead of a shared-memory load behind a bit test.
  // (ncu, r01: the list form of this pass was 17 % of the kernel's instructions.)
  if (!skip_compare) {
    constexpr int kWarpRows = kTileH / (kBlock / 32);  // 4 rows of the tile per warp
    constexpr int kCmpRows = CL_TILE_CMP_ROWS;
    static_assert(kWarpRows % kCmpRows == 0, "rows per pass must divide the warp's rows");
    const int warp = threadIdx.x >> 5;
#pragma unroll 1
    for (int pass = 0; pass < kWarpRows / kCmpRows; pass++) {
      const int r0 = warp * kWarpRows + pass * kCmpRows;  // first pixel row of this pass
      int lr[kCmpRows + NC];                              // labels of rows r0 - NC .. r0 + kCmpRows - 1, this column
#pragma unroll
      for (int j = 0; j < kCmpRows + NC; j++) {
        const int rr = r0 - NC + j;
        lr[j] = rr >= 0 ? s_lab[rr * SW + lane] : -1;  // non-dynamic pixels hold stale values: never selected below
      }
      uint32_t diff[kCmpRows];
#pragma unroll
      for (int q = 0; q < kCmpRows; q++) diff[q] = 0u;
#pragma unroll
      for (int j = 0; j < kCmpRows + NC; j++) {
#pragma unroll
        for (int dx = 0; dx <= NC; dx++) {
          // lanes < dx read a wrapped-around column: those cells lie outside the tile and are not in s_E
          const int v = dx == 0 ? lr[j] : __shfl_sync(kFull, lr[j], (lane - dx) & 31);
#pragma unroll
          for (int q = 0; q < kCmpRows; q++) {
            const int dy = q + NC - j;
            if (dy < 0 || dy > NC || (dx == 0 && dy == 0)) continue;
            if (v != lr[q + NC]) diff[q] |= 1u << ((NC - dy) * S + (NC - dx));
          }
        }
      }
#pragma unroll
      for (int q = 0; q < kCmpRows; q++) {
        const int c = (r0 + q) * SW + lane;
        uint32_t d = (s_m[c + kCore] & kDyn) ? (diff[q] & s_E[c]) : 0u;
        unioned = unioned || d != 0;
        while (__any_sync(kFull, d != 0)) {
          if (d) {
            const int k = __ffs(d) - 1;
            d &= d - 1;
            smem_union(s_lab, c, c + cell_offset(k));
          }
          __syncwarp();
        }
      }
    }
  }
#endif
  const bool reflatten = __syncthreads_or(unioned);  // labels are flat unless C2b still had to union

  // D1: flatten the tile-local forest; roots initialise their partial record
  for (int e0 = 0; e0 < nd; e0 += kBlock) {
    const int e = e0 + threadIdx.x;
    bool is_root = false;
    int c = 0;
    if (e < nd) {
      c = pixel_at(e);
      int r = s_lab[c];
      if (reflatten) {
        r = smem_find_readonly(s_lab, c);  // no halving stores here: see smem_find_readonly
        s_lab[c] = r;
      }
      is_root = (r == c);
      if (is_root) {
        const size_t gi = fb + static_cast<size_t>(ty0 + c / SW) * L.w + tx0 + c % SW;
        L.count[gi] = 0;
        TileStat init;
        init.key = 0x7fffffff;
        init.mn[0] = init.mn[1] = init.mn[2] = 0xffffffffu;
        init.mx[0] = init.mx[1] = init.mx[2] = 0u;
        init.pad = 0;
        L.tstat[gi] = init;
      }
    }
    smem_append(is_root, static_cast<uint16_t>(c), s_rl, &s_rcnt);
  }
  __syncthreads();
  const int nb = s_bcnt, nr = s_rcnt;
  if (threadIdx.x == 0) {
    s_base = atomicAdd(counters + 0, nd);
    s_bbase = nb ? atomicAdd(counters + 1, nb) : 0;
    s_rbase = atomicAdd(counters + 2, nr);
    atomicAdd(L.n_dynamic + frame, nd);
  }
  __syncthreads();

  // D2: publish parent = tile root, list entries, and the per-root partial statistics (pcl::getMinMax3D operands,
  // size, reference-order key), reduced per tile root inside each warp (redux.sync) before the atomics
  for (int e0 = 0; e0 < nd; e0 += kBlock) {
    const int e = e0 + threadIdx.x;
    int r = -1 - lane;  // distinct for idle lanes
    uint32_t ex_mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, ex_mx[3] = {0u, 0u, 0u};
    int key = 0x7fffffff;
    if (e < nd) {
      const int c = pixel_at(e);
      r = s_lab[c];
      const int i = (ty0 + c / SW) * L.w + tx0 + c % SW;
      L.parent[fb + i] = (ty0 + r / SW) * L.w + tx0 + r % SW;
      list[s_base + e] = static_cast<uint32_t>(fb + i);
      const float4 xyz = __ldg(reinterpret_cast<const float4*>(pts + static_cast<size_t>(i) * 8));
      // |v| (Vector3f::norm, the oracle's summation order) rides along with the list entry: the record's second half
      // sits in the same 32-byte sector, and cl_scatter then reads 4 bytes per member instead of a sector
      const float4 vel = __ldg(reinterpret_cast<const float4*>(pts + static_cast<size_t>(i) * 8 + 4));
      list_speed[s_base + e] = __float_as_uint(__fsqrt_rn(
          __fadd_rn(__fadd_rn(__fmul_rn(vel.x, vel.x), __fmul_rn(vel.y, vel.y)), __fmul_rn(vel.z, vel.z))));
      const float cc[3] = {xyz.x, xyz.y, xyz.z};
#pragma unroll
      for (int k = 0; k < 3; k++)  // NaN coordinates are skipped (order dependent in the reference; unpinned)
        if (!isnan(cc[k])) ex_mn[k] = ex_mx[k] = float_to_ordered(cc[k]);
      if (s_up[c]) key = i;
    }
    unsigned todo = __ballot_sync(kFull, r >= 0);
    while (todo) {  // peel one tile root at a time; usually one or two per warp
      const int leader = __ffs(todo) - 1;
      const int lr = __shfl_sync(kFull, r, leader);
      const bool in = (r == lr);
      const unsigned grp = __ballot_sync(kFull, in);
      if (in) {
        uint32_t mn[3], mx[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
          mn[k] = __reduce_min_sync(grp, ex_mn[k]);
          mx[k] = __reduce_max_sync(grp, ex_mx[k]);
        }
        const int kk = __reduce_min_sync(grp, key);
        if (lane == leader) {
          const size_t gr = fb + static_cast<size_t>(ty0 + lr / SW) * L.w + tx0 + lr % SW;
          atomicAdd(L.count + gr, __popc(grp));
          TileStat& st = L.tstat[gr];
#pragma unroll
          for (int k = 0; k < 3; k++) {
            if (mn[k] != 0xffffffffu) atomicMin(&st.mn[k], mn[k]);
            if (mx[k] != 0u) atomicMax(&st.mx[k], mx[k]);
          }
          if (kk != 0x7fffffff) atomicMin(&st.key, kk);
        }
      }
      todo &= ~grp;
    }
  }
  for (int e = threadIdx.x; e < nb; e += kBlock) {
    const int c = s_bl[e];
    border_list[s_bbase + e] = static_cast<uint32_t>(fb + static_cast<size_t>(ty0 + c / SW) * L.w + tx0 + c % SW);
    border_mask[s_bbase + e] = s_bx[e];
  }
  for (int e = threadIdx.x; e < nr; e += kBlock) {
    const int c = s_rl[e];
    root_list[s_rbase + e] = static_cast<uint32_t>(fb + static_cast<size_t>(ty0 + c / SW) * L.w + tx0 + c % SW);
  }
}

// ---- cl_merge -----------------------------------------------------------------------------------------------------
// Edges whose up-left endpoint lies in another tile, for the pixels cl_tile put on the border list; lock-free union
// with atomicMin + path halving.  Fast path: cl_tile already decided which window cells to union (border_mask).
// Generic path: the window logic is re-evaluated on global memory.
template <int NC>
__global__ void __launch_bounds__(kBlock)
    cl_merge_kernel(const ClLaunch L, const uint32_t* __restrict__ border_list,
                    const uint32_t* __restrict__ border_mask, const int* __restrict__ border_count) {
  pdl_enter();
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
      parent = L.parent + static_cast<size_t>(frame) * P;
      if constexpr (NC > 0) {
        edges.w[0] = border_mask[t];
      } else {
        const uint8_t* mask = L.mask + static_cast<size_t>(frame) * P;
        const float* pts = L.points + static_cast<size_t>(frame) * L.points_frame_stride * 8;
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
    }
    const int lx = u & (kTileW - 1), ly = v & (kTileH - 1);
    // Every pixel's parent entry holds its tile root (cl_tile) or, after path halving, an ancestor of it; all border
    // pixels along one tile edge share a handful of (root, root) pairs and consecutive list entries come from the
    // same tile, so the pairs are de-duplicated inside the warp first: one lane per distinct pair performs the
    // union.  Without this, thousands of finds per blob converge on the same parent entry (one L2 slice): measured
    // 98 -> 60 us per 63 720p frames.
    int ra = -1;
    if (t < n) ra = __ldcg(parent + i);
    const uint32_t fbase = static_cast<uint32_t>(parent - L.parent);
    for (;;) {
      const bool has = !edges.empty();
      if (__ballot_sync(kFull, has) == 0) break;
      bool cross = false;
      int rb = -1;
      if (has) {
        const int k = edges.pop();
        const int wy = k / stride, wx = k - wy * stride;
        const int dx = wx - N, dy = wy - N;
        cross = NC > 0 || lx + dx < 0 || ly + dy < 0;  // the up-left endpoint lies in another tile
        if (cross) rb = __ldcg(parent + (v + dy) * L.w + (u + dx));
      }
      const unsigned crossing = __ballot_sync(kFull, cross);
      if (cross) {
        const unsigned long long key =
            (static_cast<unsigned long long>(fbase + static_cast<uint32_t>(ra)) << 32) | static_cast<uint32_t>(rb);
        const unsigned same = __match_any_sync(crossing, key);
        if ((threadIdx.x & 31) == __ffs(same) - 1 && ra != rb) uf_union(parent, ra, rb);
      }
      __syncwarp();
    }
  }
}

// ---- cl_roots -----------------------------------------------------------------------------------------------------
// integrateConnectedClusters (:253-267) on the tile roots: parent[r] = final root R, and r's partial record is merged
// into R's.  After this kernel every dynamic pixel reaches its final root in two hops (pixel -> tile root -> R).
__global__ void __launch_bounds__(kBlock)
    cl_roots_kernel(const ClLaunch L, const uint32_t* __restrict__ root_list, const int* __restrict__ root_count) {
  pdl_enter();
  const int n = *root_count;
  const uint32_t P = static_cast<uint32_t>(L.w) * L.h;
  for (int t0 = blockIdx.x * kBlock; t0 < n; t0 += gridDim.x * kBlock) {
    const int t = t0 + threadIdx.x;
    if (t < n) {
      const uint32_t g = root_list[t];
      const uint32_t frame = g / P;
      const int r = static_cast<int>(g - frame * P);
      int* parent = L.parent + static_cast<size_t>(frame) * P;
      const int R = uf_find_readonly(parent, r);  // every store of this kernel writes a final root: no races
      if (R != r) {
        __stcg(parent + r, R);
        const size_t gR = static_cast<size_t>(frame) * P + R;
        const TileStat src = L.tstat[g];
        TileStat& dst = L.tstat[gR];
        atomicAdd(L.count + gR, L.count[g]);
#pragma unroll
        for (int k = 0; k < 3; k++) {
          if (src.mn[k] != 0xffffffffu) atomicMin(&dst.mn[k], src.mn[k]);
          if (src.mx[k] != 0u) atomicMax(&dst.mx[k], src.mx[k]);
        }
        if (src.key != 0x7fffffff) atomicMin(&dst.key, src.key);
      }
    }
    __syncwarp();
  }
}

// ---- cl_slots -----------------------------------------------------------------------------------------------------
// removeSmallClusters (:369-383): components below cluster_size are dropped; survivors get a dense slot, which is
// stored over count[R] (the size moves into the slot's ClusterStat together with the merged statistics).
__global__ void __launch_bounds__(kBlock)
    cl_slots_kernel(const ClLaunch L, const uint32_t* __restrict__ root_list, const int* __restrict__ root_count) {
  pdl_enter();
  const int n = *root_count;
  const uint32_t P = static_cast<uint32_t>(L.w) * L.h;
  for (int t = blockIdx.x * kBlock + threadIdx.x; t < n; t += gridDim.x * kBlock) {
    const uint32_t g = root_list[t];
    const uint32_t frame = g / P;
    const int r = static_cast<int>(g - frame * P);
    if (L.parent[g] != r) continue;  // merged into another tile's root
    const int cnt = L.count[g];
    int s = -1;
    if (cnt >= L.min_size) {
      s = atomicAdd(L.n_slots + frame, 1);
      if (s < L.max_clusters) {
        const TileStat ts = L.tstat[g];
        ClusterStat& st = L.stats[static_cast<size_t>(frame) * L.max_clusters + s];
        st.root = r;
        st.count = cnt;
        st.key = ts.key;
        st.rank = -1;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          st.mn[k] = ts.mn[k];
          st.mx[k] = ts.mx[k];
        }
        st.median_idx = 0x7fffffff;
      } else {
        s = -1;  // overflow: reported through frame_stats.status by cl_rank
      }
    }
    L.count[g] = s;
  }
}

// ---- cl_rank ----------------------------------------------------------------------------------------------------
// One CTA per frame.  Final id = rank of the key (removeSmallClusters' dense renumbering keeps root-label order);
// member-list offsets are an exclusive scan of the counts in slot order; clusters are appended to the work queue.
__global__ void __launch_bounds__(kBlock)
    cl_rank_kernel(const ClLaunch L, int* __restrict__ work_count, int2* __restrict__ work_items,
                   int* __restrict__ member_offset) {
  pdl_enter();
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
    cl_scatter_kernel(const ClLaunch L, const uint32_t* __restrict__ list, const uint32_t* __restrict__ list_speed,
                      const int* __restrict__ list_count,
                      const int* __restrict__ member_offset, int* __restrict__ cursor,
                      uint32_t* __restrict__ member_speed, int* __restrict__ member_pixel) {
  pdl_enter();
  const int n = *list_count;
  const uint32_t P = static_cast<uint32_t>(L.w) * L.h;
  const int lane = threadIdx.x & 31;
  // The look-up of one entry is a chain of four dependent loads (list -> tile root -> final root -> slot), so the
  // kernel is bound by memory latency, not bandwidth: every thread keeps kScatterUnroll chains in flight.
#ifndef CL_SCATTER_UNROLL
#define CL_SCATTER_UNROLL 4
#endif
  constexpr int kU = CL_SCATTER_UNROLL;
  for (int t0 = blockIdx.x * (kBlock * kU); t0 < n; t0 += gridDim.x * (kBlock * kU)) {
    int s[kU], i[kU];
    uint32_t frame[kU], g[kU];
#pragma unroll
    for (int u = 0; u < kU; u++) {
      const int t = t0 + u * kBlock + threadIdx.x;
      g[u] = t < n ? list[t] : 0xffffffffu;
    }
    uint32_t r1[kU];
#pragma unroll
    for (int u = 0; u < kU; u++) {
      frame[u] = g[u] != 0xffffffffu ? g[u] / P : 0;
      i[u] = static_cast<int>(g[u] - frame[u] * P);
      r1[u] = g[u] != 0xffffffffu ? static_cast<uint32_t>(L.parent[g[u]]) : 0;
    }
#pragma unroll
    for (int u = 0; u < kU; u++)
      if (g[u] != 0xffffffffu) r1[u] = static_cast<uint32_t>(L.parent[frame[u] * P + r1[u]]);
#pragma unroll
    for (int u = 0; u < kU; u++) s[u] = g[u] != 0xffffffffu ? L.count[frame[u] * P + r1[u]] : -1;
#pragma unroll
    for (int u = 0; u < kU; u++) {
      const unsigned member = __ballot_sync(kFull, s[u] >= 0);
      if (member == 0) continue;
      const long long group =
          (s[u] >= 0) ? ((static_cast<long long>(frame[u]) << 32) | static_cast<unsigned>(s[u])) : (-1 - lane);
      const unsigned same = __match_any_sync(kFull, group);
      if (s[u] >= 0) {
        const int leader = __ffs(same) - 1;
        const size_t cs = static_cast<size_t>(frame[u]) * L.max_clusters + s[u];
        int base = 0;
        if (lane == leader) base = atomicAdd(cursor + cs, __popc(same));
        base = __shfl_sync(same, base, leader);
        const size_t pos =
            static_cast<size_t>(frame[u]) * P + member_offset[cs] + base + __popc(same & ((1u << lane) - 1));
        member_speed[pos] = list_speed[t0 + u * kBlock + threadIdx.x];  // |v| from cl_tile; >= 0: bit pattern monotone
        member_pixel[pos] = i[u];
      }
    }
  }
}

// ---- cl_objects -------------------------------------------------------------------------------------------------
// cluster2MovingObject (:147-184) for one cluster per CTA (persistent CTAs pull clusters from the work queue):
// bbox = max - min, centre = (min + max) / 2 (float), velocity = member with the (n/2)-th largest |v|.
// The k-th largest is found by a 3-pass MSB-first radix select (11/11/10 bits) over the cluster's member list with a
// 2048-bin shared-memory histogram (warp-aggregated increments) and a parallel suffix scan; lists up to kSmemMembers
// are staged in shared memory once, longer ones stream from global memory with four loads in flight per thread.
// Among members tied at the selected speed the smallest raster index is taken (the reference's std::sort leaves
// ties unspecified).
constexpr int kObjBlock = 1024;
constexpr int kSmemMembers = 9216;  // 36 KB
constexpr int kBins = 2048;

// Which of the two cl_objects kernels runs is decided on the device, identically by every CTA of both (the other one
// returns at once): cooperating CTAs pay ~8 cluster barriers per object, which only a long member list amortises, and
// a cluster of kObjCluster CTAs leaves 2 * 148 / kObjCluster objects in flight.  Measured (profiles/r01_clusterer.md):
// 3840x2160 frames (objects of ~250 k members) 0.34 -> 0.21 ms per frame, 63 frames of 1280x720 (230 objects of ~30 k)
// 0.57 -> 0.73 ms if forced.
constexpr int kObjCluster = 8;
constexpr int kObjClusterMinMembers = 96 * 1024;      // the largest object must be at least this long ...
constexpr int kObjClusterMaxObjects = 2 * (2 * 148 / kObjCluster);  // ... and the objects fit in two rounds
// Frames below this size (1280x720 and smaller: an object of 96 k members would be a tenth of the frame) never launch
// the cluster variant, so the common configurations pay nothing for the choice.
constexpr size_t kObjClusterMinPixels = 3u << 19;  // 1.5 Mpx
__device__ __forceinline__ bool objects_use_clusters(const ClLaunch& L, int n_work, const int2* __restrict__ work_items) {
  __shared__ int s_max_members;
  if (threadIdx.x == 0) s_max_members = 0;
  __syncthreads();
  if (n_work <= kObjClusterMaxObjects) {
    int mx = 0;
    for (int i = threadIdx.x; i < n_work; i += blockDim.x) {
      const int2 w = work_items[i];
      mx = max(mx, L.stats[static_cast<size_t>(w.x) * L.max_clusters + w.y].count);
    }
    if (mx > 0) atomicMax(&s_max_members, mx);
  }
  __syncthreads();
  return s_max_members >= kObjClusterMinMembers;
}

// ---- member-list passes shared by the two cl_objects kernels ---------------------------------------------------------
// A cluster's member list is walked three (radix passes) plus one (tie-break) times.  Lists that do not fit the shared
// memory staging stream from global memory (L2: cl_scatter has just written them), and a pass over a 30 k list by 1024
// threads is only a handful of iterations, so its time is load latency, not bandwidth: the streaming form reads the
// list as 16-byte vectors, two in flight per thread (eight members), after a scalar head that aligns the pointer.
#ifndef CL_OBJECTS_MATCH_FIRST_ONLY
#define CL_OBJECTS_MATCH_FIRST_ONLY 1
#endif
template <bool kTryUniform>
__device__ __forceinline__ void obj_count(unsigned* hist, uint32_t v, bool ok, uint32_t himask, uint32_t prefix, int shift,
                                          uint32_t binmask, int lane) {
  int bin = -1;
  if (ok && (v & himask) == prefix) bin = static_cast<int>((v >> shift) & binmask);
  // members of one object move alike: in the first pass nearly every warp is uniform (one atomic per warp);
  // later passes spread over the bins and plain shared-memory atomics are cheaper than a match
  if constexpr (kTryUniform) {
    int uniform;
    __match_all_sync(kFull, bin, &uniform);
    if (uniform) {
      if (bin >= 0 && lane == 0) atomicAdd(&hist[bin], 32u);
      return;
    }
  }
  if (bin >= 0) atomicAdd(&hist[bin], 1u);
}

// head / vector split of a global list of n 32-bit entries: [0, head) scalar, nvec uint4 from head, then a scalar tail
struct ObjSplit {
  int head, nvec, tail0;
};
__device__ __forceinline__ ObjSplit obj_split(const void* p, int n) {
  ObjSplit sp;
  sp.head = min(n, static_cast<int>((4u - ((reinterpret_cast<uintptr_t>(p) >> 2) & 3u)) & 3u));
  sp.nvec = (n - sp.head) >> 2;
  sp.tail0 = sp.head + 4 * sp.nvec;
  return sp;
}

// One histogram pass over the list (every thread of the CTA calls it: the match in obj_count is warp-wide).
template <int kThreads, bool kTryUniform = true>
__device__ __forceinline__ void obj_hist_pass(unsigned* hist, const uint32_t* speeds, int n, bool in_global,
                                              uint32_t himask, uint32_t prefix, int shift, uint32_t binmask) {
  const int tid = threadIdx.x, lane = tid & 31;
  if (!in_global) {  // staged in shared memory
    for (int j0 = 0; j0 < n; j0 += 4 * kThreads) {
      uint32_t v[4];
      bool ok[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int j = j0 + q * kThreads + tid;
        ok[q] = j < n;
        v[q] = ok[q] ? speeds[j] : 0u;
      }
#pragma unroll
      for (int q = 0; q < 4; q++) obj_count<kTryUniform>(hist, v[q], ok[q], himask, prefix, shift, binmask, lane);
    }
    return;
  }
  const ObjSplit sp = obj_split(speeds, n);
  {  // the at most three entries before and after the aligned body, one thread each
    int j = -1;
    if (tid < sp.head) j = tid;
    else if (tid - sp.head < n - sp.tail0) j = sp.tail0 + (tid - sp.head);
    obj_count<kTryUniform>(hist, j >= 0 ? __ldcg(speeds + j) : 0u, j >= 0, himask, prefix, shift, binmask, lane);
  }
  const uint4* vec = reinterpret_cast<const uint4*>(speeds + sp.head);
  for (int i0 = 0; i0 < sp.nvec; i0 += 2 * kThreads) {
    const int ia = i0 + tid, ib = ia + kThreads;
    const bool oka = ia < sp.nvec, okb = ib < sp.nvec;
    uint4 a = make_uint4(0u, 0u, 0u, 0u), b = a;
    if (oka) a = __ldcg(vec + ia);
    if (okb) b = __ldcg(vec + ib);
    obj_count<kTryUniform>(hist, a.x, oka, himask, prefix, shift, binmask, lane);
    obj_count<kTryUniform>(hist, a.y, oka, himask, prefix, shift, binmask, lane);
    obj_count<kTryUniform>(hist, a.z, oka, himask, prefix, shift, binmask, lane);
    obj_count<kTryUniform>(hist, a.w, oka, himask, prefix, shift, binmask, lane);
    obj_count<kTryUniform>(hist, b.x, okb, himask, prefix, shift, binmask, lane);
    obj_count<kTryUniform>(hist, b.y, okb, himask, prefix, shift, binmask, lane);
    obj_count<kTryUniform>(hist, b.z, okb, himask, prefix, shift, binmask, lane);
    obj_count<kTryUniform>(hist, b.w, okb, himask, prefix, shift, binmask, lane);
  }
}

// Smallest pixel index among the members whose speed equals `target` (this thread's share; 0x7fffffff if none).
template <int kThreads>
__device__ __forceinline__ int obj_tie_break(const uint32_t* speeds, const int* pixels, int n, bool in_global,
                                             uint32_t target) {
  const int tid = threadIdx.x;
  int best = 0x7fffffff;
  if (!in_global) {
    for (int j = tid; j < n; j += kThreads)
      if (speeds[j] == target) best = min(best, pixels[j]);
    return best;
  }
  const ObjSplit sp = obj_split(speeds, n);
  {
    int j = -1;
    if (tid < sp.head) j = tid;
    else if (tid - sp.head < n - sp.tail0) j = sp.tail0 + (tid - sp.head);
    if (j >= 0 && __ldcg(speeds + j) == target) best = min(best, pixels[j]);
  }
  const uint4* vec = reinterpret_cast<const uint4*>(speeds + sp.head);
  for (int i0 = 0; i0 < sp.nvec; i0 += 2 * kThreads) {
    const int ia = i0 + tid, ib = ia + kThreads;
    uint4 a = make_uint4(~target, ~target, ~target, ~target), b = a;
    if (ia < sp.nvec) a = __ldcg(vec + ia);
    if (ib < sp.nvec) b = __ldcg(vec + ib);
    const int ja = sp.head + 4 * ia, jb = sp.head + 4 * ib;
    if (a.x == target) best = min(best, pixels[ja]);
    if (a.y == target) best = min(best, pixels[ja + 1]);
    if (a.z == target) best = min(best, pixels[ja + 2]);
    if (a.w == target) best = min(best, pixels[ja + 3]);
    if (b.x == target) best = min(best, pixels[jb]);
    if (b.y == target) best = min(best, pixels[jb + 1]);
    if (b.z == target) best = min(best, pixels[jb + 2]);
    if (b.w == target) best = min(best, pixels[jb + 3]);
  }
  return best;
}

// Copies the members of a global list whose speed matches `prefix` in its `himask` bits -- the candidates the first two
// radix passes left -- into shared memory (speed bits and pixel index, any order).  The caller knows their number (the
// selected bin's count) and only calls this when they fit.
template <int kThreads>
__device__ __forceinline__ void obj_collect(const uint32_t* speeds, const int* pixels, int n, uint32_t himask,
                                            uint32_t prefix, uint32_t* c_speed, int* c_pixel, int* counter) {
  const int tid = threadIdx.x;
  auto take = [&](uint32_t v, int j) {
    if ((v & himask) == prefix) {
      const int pos = atomicAdd(counter, 1);
      c_speed[pos] = v;
      c_pixel[pos] = pixels[j];
    }
  };
  const ObjSplit sp = obj_split(speeds, n);
  {
    int j = -1;
    if (tid < sp.head) j = tid;
    else if (tid - sp.head < n - sp.tail0) j = sp.tail0 + (tid - sp.head);
    if (j >= 0) take(__ldcg(speeds + j), j);
  }
  const uint4* vec = reinterpret_cast<const uint4*>(speeds + sp.head);
  for (int i0 = 0; i0 < sp.nvec; i0 += 2 * kThreads) {
    const int ia = i0 + tid, ib = ia + kThreads;
    uint4 a = make_uint4(~prefix, ~prefix, ~prefix, ~prefix), b = a;  // ~prefix never matches (himask != 0)
    if (ia < sp.nvec) a = __ldcg(vec + ia);
    if (ib < sp.nvec) b = __ldcg(vec + ib);
    const int ja = sp.head + 4 * ia, jb = sp.head + 4 * ib;
    take(a.x, ja);
    take(a.y, ja + 1);
    take(a.z, ja + 2);
    take(a.w, ja + 3);
    take(b.x, jb);
    take(b.y, jb + 1);
    take(b.z, jb + 2);
    take(b.w, jb + 3);
  }
}

template <bool kClustersPossible>  // false: the instantiation small frames run never looks at the choice
__global__ void __launch_bounds__(kObjBlock, 2)
    cl_objects_kernel(const ClLaunch L, const int* __restrict__ work_count, const int2* __restrict__ work_items,
                      int* __restrict__ work_cursor, const int* __restrict__ member_offset,
                      const uint32_t* __restrict__ member_speed, const int* __restrict__ member_pixel) {
  pdl_enter();
  __shared__ uint32_t s_speed[kSmemMembers];
  __shared__ unsigned hist[kBins];
  __shared__ unsigned s_wsum[kObjBlock / 32];
  __shared__ int s_item;
  __shared__ uint32_t s_prefix;
  __shared__ uint32_t s_k;
  __shared__ int s_best;
  __shared__ unsigned s_sel;  // members in the bin the last pass selected
  __shared__ int s_cn;
  const size_t P = static_cast<size_t>(L.w) * L.h;
  const int n_work = *work_count;
  if constexpr (kClustersPossible) {
    if (objects_use_clusters(L, n_work, work_items)) return;  // cl_objects_cluster_kernel takes this batch
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (;;) {
    if (threadIdx.x == 0) s_item = atomicAdd(work_cursor, 1);
    __syncthreads();
    const int item = s_item;
    if (item >= n_work) return;
    const int2 w = work_items[item];
    const int frame = w.x, slot = w.y;
    ClusterStat& st = L.stats[static_cast<size_t>(frame) * L.max_clusters + slot];
    const size_t off = frame * P + member_offset[static_cast<size_t>(frame) * L.max_clusters + slot];
    const uint32_t* gspeeds = member_speed + off;
    const int* gpixels = member_pixel + off;
    const int n_all = st.count;
    bool staged = n_all <= kSmemMembers;
    if (staged)
      for (int j = threadIdx.x; j < n_all; j += kObjBlock) s_speed[j] = gspeeds[j];
    const uint32_t* speeds = staged ? s_speed : gspeeds;
    const int* pixels = gpixels;
    int n = n_all;  // members the passes walk: the whole list, or the candidates collected after the second pass
    if (threadIdx.x == 0) {
      s_prefix = 0;
      s_k = static_cast<uint32_t>(n / 2);  // cluster.at(cluster.size() / 2) after the descending sort
      s_best = 0x7fffffff;
      s_cn = 0;
    }
    int decided = 0;  // number of high bits already fixed
    for (int pass = 0; pass < 3; pass++) {
      const int bits = pass < 2 ? 11 : 10;
      const int shift = 32 - decided - bits;
      for (int b = threadIdx.x; b < kBins; b += kObjBlock) hist[b] = 0;
      __syncthreads();
      const uint32_t prefix = s_prefix;
      const uint32_t himask = decided == 0 ? 0u : (0xffffffffu << (32 - decided));
      const uint32_t binmask = (1u << bits) - 1u;
#ifndef CL_OBJECTS_COLLECT
#define CL_OBJECTS_COLLECT 1
#endif
      if (CL_OBJECTS_COLLECT && pass == 2 && !staged && s_sel <= kSmemMembers / 2) {
        // A list that streams from global memory: 22 bits are decided and the selected bin's members (a handful unless
        // thousands tie) are all that the last pass and the tie-break still look at.  One more walk over the list
        // copies them into the (unused) staging buffer; the last pass and the tie-break then run on shared memory
        // instead of walking the list twice more.
        uint32_t* c_speed = s_speed;
        int* c_pixel = reinterpret_cast<int*>(s_speed + kSmemMembers / 2);
        obj_collect<kObjBlock>(gspeeds, gpixels, n_all, himask, prefix, c_speed, c_pixel, &s_cn);
        __syncthreads();
        n = s_cn;
        speeds = c_speed;
        pixels = c_pixel;
        staged = true;
      }
      if (CL_OBJECTS_MATCH_FIRST_ONLY && pass > 0)
        obj_hist_pass<kObjBlock, false>(hist, speeds, n, !staged, himask, prefix, shift, binmask);
      else
        obj_hist_pass<kObjBlock, true>(hist, speeds, n, !staged, himask, prefix, shift, binmask);
      __syncthreads();
      // parallel suffix scan: thread t owns bins hi = kBins-1-2t and hi-1 (descending order, k counts from the top)
      const int hi = kBins - 1 - 2 * static_cast<int>(threadIdx.x);
      const unsigned h0 = hist[hi], h1 = hist[hi - 1];
      unsigned incl = h0 + h1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) s_wsum[warp] = incl;
      __syncthreads();
      if (warp == 0) {
        unsigned ws = s_wsum[lane];
        unsigned wi = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned t = __shfl_up_sync(kFull, wi, o);
          if (lane >= o) wi += t;
        }
        s_wsum[lane] = wi - ws;  // exclusive
      }
      __syncthreads();
      const unsigned before = s_wsum[warp] + incl - (h0 + h1);  // members in bins above `hi`
      const uint32_t k = s_k;
      __syncthreads();  // every thread has read s_k before it is overwritten
      if (k >= before && k < before + h0) {
        s_k = k - before;
        s_prefix = prefix | (static_cast<uint32_t>(hi) << shift);
        s_sel = h0;
      } else if (k >= before + h0 && k < before + h0 + h1) {
        s_k = k - before - h0;
        s_prefix = prefix | (static_cast<uint32_t>(hi - 1) << shift);
        s_sel = h1;
      }
      decided += bits;
      __syncthreads();
    }
    const uint32_t target = s_prefix;
    int best = obj_tie_break<kObjBlock>(speeds, pixels, n, !staged, target);
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
        o.n_points = n_all;
        L.objs[static_cast<size_t>(frame) * L.max_objs + id] = o;
      }
    }
    __syncthreads();
  }
}

// ---- cl_objects on thread-block clusters ----------------------------------------------------------------------------
// Same computation, one cluster of kC CTAs per moving-object cluster: the member list is cut into kC contiguous slices,
// every CTA histograms its slice into its own shared memory, and rank 0 reads the other CTAs' histograms through
// distributed shared memory for the suffix scan; the selection state (prefix, k, tie-break minimum) lives in rank 0's
// shared memory and is read / updated by the others through DSMEM, with the hardware cluster barrier between the
// phases.  A frame's few largest objects no longer serialise on one CTA each (the kernel's time was the longest list).
namespace cg = cooperative_groups;

template <int kC>
__global__ void __launch_bounds__(kObjBlock, 2)
    cl_objects_cluster_kernel(const ClLaunch L, const int* __restrict__ work_count,
                              const int2* __restrict__ work_items, int* __restrict__ work_cursor,
                              const int* __restrict__ member_offset, const uint32_t* __restrict__ member_speed,
                              const int* __restrict__ member_pixel) {
  pdl_enter();
  __shared__ uint32_t s_speed[kSmemMembers];
  __shared__ unsigned hist[kBins];
  __shared__ unsigned s_wsum[kObjBlock / 32];
  __shared__ int s_item;
  __shared__ uint32_t s_prefix;
  __shared__ uint32_t s_k;
  __shared__ int s_best;
  __shared__ unsigned s_sel;  // rank 0: members (of the whole object) in the bin the last pass selected
  __shared__ int s_cn;        // this CTA's collected candidates
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned crank = cluster.block_rank();
  const int* item0 = cluster.map_shared_rank(&s_item, 0);
  const uint32_t* prefix0 = cluster.map_shared_rank(&s_prefix, 0);
  const unsigned* sel0 = cluster.map_shared_rank(&s_sel, 0);
  int* best0 = cluster.map_shared_rank(&s_best, 0);
  const size_t P = static_cast<size_t>(L.w) * L.h;
  const int n_work = *work_count;
  if (!objects_use_clusters(L, n_work, work_items)) return;  // cl_objects_kernel takes this batch (no DSMEM touched yet)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (;;) {
    if (crank == 0 && threadIdx.x == 0) {
      const int it = atomicAdd(work_cursor, 1);
      s_item = it;
      s_prefix = 0;
      s_best = 0x7fffffff;
      if (it < n_work) {
        const int2 w0 = work_items[it];
        s_k = static_cast<uint32_t>(L.stats[static_cast<size_t>(w0.x) * L.max_clusters + w0.y].count / 2);
      }
    }
    cluster.sync();
    const int item = *item0;
    if (item >= n_work) {  // the same for every CTA of the cluster
      cluster.sync();      // rank 0's shared memory must outlive the other CTAs' read of it
      return;
    }
    const int2 w = work_items[item];
    const int frame = w.x, slot = w.y;
    ClusterStat& st = L.stats[static_cast<size_t>(frame) * L.max_clusters + slot];
    const int n = st.count;
    const size_t off = frame * P + member_offset[static_cast<size_t>(frame) * L.max_clusters + slot];
    // this CTA's slice of the member list
    const int len = (n + kC - 1) / kC;
    const int begin = min(n, static_cast<int>(crank) * len), end = min(n, begin + len);
    const int m_all = end - begin;
    int m = m_all;  // members the passes walk: the slice, or its candidates collected after the second pass
    const uint32_t* gspeeds = member_speed + off + begin;
    const int* gpixels = member_pixel + off + begin;
    const int* pixels = gpixels;
    bool staged = m <= kSmemMembers;
    if (staged)
      for (int j = threadIdx.x; j < m; j += kObjBlock) s_speed[j] = gspeeds[j];
    const uint32_t* speeds = staged ? s_speed : gspeeds;
    if (threadIdx.x == 0) s_cn = 0;  // ordered before its use by the barriers of the first two passes
    int decided = 0;  // number of high bits already fixed
    for (int pass = 0; pass < 3; pass++) {
      const int bits = pass < 2 ? 11 : 10;
      const int shift = 32 - decided - bits;
      for (int b = threadIdx.x; b < kBins; b += kObjBlock) hist[b] = 0;
      __syncthreads();
      const uint32_t prefix = *prefix0;
      const uint32_t himask = decided == 0 ? 0u : (0xffffffffu << (32 - decided));
      const uint32_t binmask = (1u << bits) - 1u;
      if (CL_OBJECTS_COLLECT && pass == 2 && !staged && *sel0 <= kSmemMembers / 2) {
        // as in cl_objects_kernel: this CTA's candidates of the selected bin move to shared memory, the last pass and
        // the tie-break run there (the whole object's bin count bounds every slice's share)
        uint32_t* c_speed = s_speed;
        int* c_pixel = reinterpret_cast<int*>(s_speed + kSmemMembers / 2);
        obj_collect<kObjBlock>(gspeeds, gpixels, m_all, himask, prefix, c_speed, c_pixel, &s_cn);
        __syncthreads();
        m = s_cn;
        speeds = c_speed;
        pixels = c_pixel;
        staged = true;
      }
      if (CL_OBJECTS_MATCH_FIRST_ONLY && pass > 0)
        obj_hist_pass<kObjBlock, false>(hist, speeds, m, !staged, himask, prefix, shift, binmask);
      else
        obj_hist_pass<kObjBlock, true>(hist, speeds, m, !staged, himask, prefix, shift, binmask);
      cluster.sync();  // every CTA's histogram is complete
      if (crank == 0) {
        // parallel suffix scan over the summed histograms: thread t owns bins hi = kBins-1-2t and hi-1
        const int hi = kBins - 1 - 2 * static_cast<int>(threadIdx.x);
        unsigned h0 = hist[hi], h1 = hist[hi - 1];
#pragma unroll
        for (int r = 1; r < kC; r++) {
          const unsigned* rh = cluster.map_shared_rank(hist, r);
          h0 += rh[hi];
          h1 += rh[hi - 1];
        }
        unsigned incl = h0 + h1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned t = __shfl_up_sync(kFull, incl, o);
          if (lane >= o) incl += t;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
          unsigned ws = s_wsum[lane];
          unsigned wi = ws;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(kFull, wi, o);
            if (lane >= o) wi += t;
          }
          s_wsum[lane] = wi - ws;  // exclusive
        }
        __syncthreads();
        const unsigned before = s_wsum[warp] + incl - (h0 + h1);  // members in bins above `hi`
        const uint32_t k = s_k;
        __syncthreads();  // every thread has read s_k before it is overwritten
        if (k >= before && k < before + h0) {
          s_k = k - before;
          s_prefix = prefix | (static_cast<uint32_t>(hi) << shift);
          s_sel = h0;
        } else if (k >= before + h0 && k < before + h0 + h1) {
          s_k = k - before - h0;
          s_prefix = prefix | (static_cast<uint32_t>(hi - 1) << shift);
          s_sel = h1;
        }
      }
      decided += bits;
      cluster.sync();  // the new prefix is visible; rank 0 has finished reading the other histograms
    }
    const uint32_t target = *prefix0;
    int best = obj_tie_break<kObjBlock>(speeds, pixels, m, !staged, target);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(kFull, best, o));
    if (lane == 0 && best != 0x7fffffff) atomicMin(best0, best);
    cluster.sync();
    if (crank == 0 && threadIdx.x == 0) {
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
    // the other CTAs touch rank 0's control words again only after the cluster barrier at the top of the loop
  }
}

// ---- cl_labels --------------------------------------------------------------------------------------------------
// Final cluster_map_ (after removeSmallClusters' relabel, :386-392), dense, -1 = NOT_BELONGED_.
__global__ void __launch_bounds__(kBlock) cl_labels_kernel(const ClLaunch L) {
  pdl_enter();
  const size_t P = static_cast<size_t>(L.w) * L.h;
  const size_t total = P * L.batch;
  for (size_t g = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; g < total;
       g += static_cast<size_t>(gridDim.x) * blockDim.x) {
    int label = -1;
    if (L.mask[g] != 0) {
      const size_t frame = g / P;
      const size_t fb = frame * P;
      const int s = L.count[fb + L.parent[fb + L.parent[g]]];  // pixel -> tile root -> final root -> slot
      if (s >= 0) label = L.stats[frame * L.max_clusters + s].rank;
    }
    L.labels[g] = label;
  }
}

// ---- cl_image ---------------------------------------------------------------------------------------------------
// ~clusters_image (ClustererNodelet::publishClustersImage, clusterer_nodelet.cpp:293-322): BGR8, black for
// NOT_BELONGED_, otherwise entry (label * 255 / n_clusters) of OpenCV's HSV colour map (ColorSet::resize,
// color_set.cpp:16-26: input_map[i] = i * 255 / set_size, cv::applyColorMap(COLORMAP_HSV)).
__constant__ uint32_t kHsvColormap[256] = {
#include "hsv_colormap.inc"
};

__global__ void __launch_bounds__(kBlock)
    cl_image_kernel(const int32_t* __restrict__ labels, const mod_frame_stats* __restrict__ stats, int fixed_n, size_t P,
                    int batch, uint8_t* __restrict__ out_bgr) {
  const size_t total = P * static_cast<size_t>(batch);
  for (size_t g = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; g < total;
       g += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int label = labels[g];
    uint32_t c = 0;  // (b, g, r) = (0, 0, 0)
    if (label >= 0) {
      const int n = stats ? stats[g / P].n_clusters : fixed_n;  // color_set_.resize(number_of_clusters_)
      if (n > 0 && label < n) c = kHsvColormap[(static_cast<long long>(label) * 255 / n) & 255];
    }
    out_bgr[3 * g + 0] = static_cast<uint8_t>(c >> 16);
    out_bgr[3 * g + 1] = static_cast<uint8_t>(c >> 8);
    out_bgr[3 * g + 2] = static_cast<uint8_t>(c);
  }
}

__global__ void cl_reset_kernel(int* a, size_t na, int* b, size_t nb, int* c, size_t nc) {
  pdl_enter();
  const size_t t = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (t < na) a[t] = 0;
  if (t < nb) b[t] = 0;
  if (t < nc) c[t] = 0;
}

#include "cluster_scan.cuh"
#include "cluster_wide.cuh"

// Developer switch for same-run A/B measurements: MOD_B200_CL_TILE=scan selects the raster-scan kernel (cl_scan_kernel,
// first version: 0.637 vs 0.559 ms of clusterer time on the benchmark), MOD_B200_CL_TILE=wide the halo-node kernel of
// neighbor_distance 5..10 (cl_tile_wide_kernel) instead of the list-driven tile kernel for neighbor_distance 1..4.
int tile_kernel_choice() {
  static const int choice = [] {
    const char* e = getenv("MOD_B200_CL_TILE");
    if (e != nullptr && strcmp(e, "scan") == 0) return 1;
    if (e != nullptr && strcmp(e, "wide") == 0) return 2;
    return 0;
  }();
  return choice;
}

template <int NC>
void launch_tile_merge(cudaStream_t s, const ClLaunch& L, const ClExtra& X, uint32_t* border_list, uint32_t* root_list,
                       int* counters) {
  const dim3 tgrid((L.w + kTileW - 1) / kTileW, (L.h + kTileH - 1) / kTileH, L.batch);
  if constexpr (NC > 4) {
    launch_chain(cl_tile_wide_kernel<NC>, tgrid, dim3(kBlock), 0, s, L, X.list, X.list_speed, border_list, X.border_mask,
                 root_list, counters);
    launch_chain(cl_merge_pairs_kernel, dim3(kListGrid), dim3(kBlock), 0, s, L, border_list, X.border_mask, counters + 1);
    return;
  } else if constexpr (NC > 0) {
    const int choice = tile_kernel_choice();
    if (choice == 2) {
      cl_tile_wide_kernel<NC><<<tgrid, kBlock, 0, s>>>(L, X.list, X.list_speed, border_list, X.border_mask, root_list,
                                                       counters);
      cl_merge_pairs_kernel<<<kListGrid, kBlock, 0, s>>>(L, border_list, X.border_mask, counters + 1);
      return;
    }
    if (choice == 0) {
      launch_chain(cl_tile_fast_kernel<NC>, tgrid, dim3(kBlock), 0, s, L, X.list, X.list_speed, border_list, X.border_mask,
                   root_list, counters);
    } else {
      const int tiles_x = (L.w + (32 - NC) - 1) / (32 - NC);
      const dim3 sgrid((tiles_x + CL_SCAN_WARPS - 1) / CL_SCAN_WARPS, (L.h + CL_SCAN_ROWS - 1) / CL_SCAN_ROWS, L.batch);
      cl_scan_kernel<NC><<<sgrid, CL_SCAN_WARPS * 32, 0, s>>>(L, X.list, X.list_speed, border_list, X.border_mask,
                                                             root_list, counters);
    }
  } else {
    cl_tile_kernel<<<tgrid, kBlock, 0, s>>>(L, X.list, X.list_speed, border_list, root_list, counters);
  }
  launch_chain(cl_merge_kernel<NC>, dim3(kListGrid), dim3(kBlock), 0, s, L, border_list, X.border_mask, counters + 1);
}

}  // namespace

void host_hsv_colormap(uint8_t* bgr768) {
  static const uint32_t table[256] = {
#include "hsv_colormap.inc"
  };
  for (int i = 0; i < 256; i++) {
    bgr768[3 * i + 0] = static_cast<uint8_t>(table[i] >> 16);
    bgr768[3 * i + 1] = static_cast<uint8_t>(table[i] >> 8);
    bgr768[3 * i + 2] = static_cast<uint8_t>(table[i]);
  }
}

cudaError_t launch_clusters_image(cudaStream_t s, const int32_t* labels, const mod_frame_stats* stats, int fixed_n,
                                  int w, int h, int batch, uint8_t* out_bgr) {
  const size_t P = static_cast<size_t>(w) * h;
  size_t blocks = (P * batch + kBlock - 1) / kBlock;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cl_image_kernel<<<static_cast<unsigned>(blocks), kBlock, 0, s>>>(labels, stats, fixed_n, P, batch, out_bgr);
  return cudaGetLastError();
}

cudaError_t launch_cluster(cudaStream_t s, const ClLaunch& L, const ClExtra& X, int* n_launches) {
  const size_t P = static_cast<size_t>(L.w) * L.h;
  const size_t total = P * L.batch;
  const size_t n_fs = static_cast<size_t>(L.batch);
  const size_t n_cs = static_cast<size_t>(L.batch) * L.max_clusters;
  {
    // zero n_slots / n_dynamic (adjacent, 2*batch), cursor (batch*max_clusters), and the five counters
    // work_count, work_cursor, list_count, border_count, root_count
    const size_t m = n_cs > 2 * n_fs ? n_cs : 2 * n_fs;
    // an ordinary launch: the chain starts here (extending it backwards -- this kernel dependent on the scene-flow
    // kernel, which then releases its dependents early -- measured +1 us on the scene-flow kernel and nothing here)
    cl_reset_kernel<<<static_cast<unsigned>((m + 255) / 256), 256, 0, s>>>(L.n_slots, 2 * n_fs, X.cursor, n_cs,
                                                                          X.work_count, 5);
  }
  int* counters = X.work_count + 2;  // list_count, border_count, root_count
  // both lists are consumed before cl_scatter fills the member arrays they borrow
  uint32_t* border_list = reinterpret_cast<uint32_t*>(X.member_pixel);
  uint32_t* root_list = X.member_speed;
  switch (L.n) {
    case 1: launch_tile_merge<1>(s, L, X, border_list, root_list, counters); break;
    case 2: launch_tile_merge<2>(s, L, X, border_list, root_list, counters); break;
    case 3: launch_tile_merge<3>(s, L, X, border_list, root_list, counters); break;
    case 4: launch_tile_merge<4>(s, L, X, border_list, root_list, counters); break;
    case 5: launch_tile_merge<5>(s, L, X, border_list, root_list, counters); break;
    case 6: launch_tile_merge<6>(s, L, X, border_list, root_list, counters); break;
    case 7: launch_tile_merge<7>(s, L, X, border_list, root_list, counters); break;
    case 8: launch_tile_merge<8>(s, L, X, border_list, root_list, counters); break;
    case 9: launch_tile_merge<9>(s, L, X, border_list, root_list, counters); break;
    case 10: launch_tile_merge<10>(s, L, X, border_list, root_list, counters); break;
    default: launch_tile_merge<0>(s, L, X, border_list, root_list, counters); break;
  }
  launch_chain(cl_roots_kernel, dim3(kListGrid), dim3(kBlock), 0, s, L, root_list, counters + 2);
  launch_chain(cl_slots_kernel, dim3(kListGrid), dim3(kBlock), 0, s, L, root_list, counters + 2);
  const size_t rank_smem = static_cast<size_t>(L.max_clusters) * sizeof(int);
  if (rank_smem > 48 * 1024) {  // contexts created for more than 12288 clusters per frame: opt in to the larger table
    const cudaError_t e = cudaFuncSetAttribute(cl_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               static_cast<int>(rank_smem));
    if (e != cudaSuccess) return e;
  }
  launch_chain(cl_rank_kernel, dim3(L.batch), dim3(kBlock), rank_smem, s, L, X.work_count, X.work_items, X.member_offset);
  launch_chain(cl_scatter_kernel, dim3(kListGrid), dim3(kBlock), 0, s, L, X.list, X.list_speed, counters, X.member_offset,
               X.cursor, X.member_speed, X.member_pixel);
  const bool clusters_possible = static_cast<size_t>(L.w) * L.h >= kObjClusterMinPixels;
  if (!clusters_possible) {
    launch_chain(cl_objects_kernel<false>, dim3(2 * 148), dim3(kObjBlock), 0, s, L, X.work_count, X.work_items,
                 X.work_cursor, X.member_offset, X.member_speed, X.member_pixel);
  } else {  // plus the thread-block-cluster variant; exactly one of the two processes the batch
    launch_chain(cl_objects_kernel<true>, dim3(2 * 148), dim3(kObjBlock), 0, s, L, X.work_count, X.work_items,
                 X.work_cursor, X.member_offset, X.member_speed, X.member_pixel);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * 148 / kObjCluster * kObjCluster);
    cfg.blockDim = dim3(kObjBlock);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kObjCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, cl_objects_cluster_kernel<kObjCluster>, L,
                                             static_cast<const int*>(X.work_count), static_cast<const int2*>(X.work_items),
                                             X.work_cursor, static_cast<const int*>(X.member_offset),
                                             static_cast<const uint32_t*>(X.member_speed),
                                             static_cast<const int*>(X.member_pixel));
    if (e != cudaSuccess) return e;
  }
  if (L.labels != nullptr) {
    size_t lb = (total + kBlock - 1) / kBlock;
    if (lb > 148 * 16) lb = 148 * 16;
    cl_labels_kernel<<<static_cast<unsigned>(lb), kBlock, 0, s>>>(L);
  }
  if (n_launches) *n_launches = 8 + (clusters_possible ? 1 : 0) + (L.labels != nullptr ? 1 : 0);
  return cudaGetLastError();
}

}  // namespace mod
