// cluster_wide.cuh -- cl_tile_wide / cl_merge_pairs: the tile-local part of the clusterer for neighbor_distance 5..10.
// Included by cluster.cu inside its anonymous namespace (uses its helpers).
//
// Same contract towards cl_roots / cl_slots / cl_rank / cl_scatter / cl_objects as the other tile kernels (parent =
// tile root, per-tile-root partial statistics, list of dynamic pixels with |v|, list of tile roots); replaces
// calculateInitialClusterMap + comparePoints + LookupTable (clusterer_nodelet.cpp:56-83, :186-219,
// include/lookup_table.h:10-33) for windows of up to 11 x 11 cells.
//
// A window of (N+1)^2 - 1 = 120 cells per pixel does not fit the bit-set formulation of cl_tile_fast_kernel (one word
// of accepted cells per pixel, implied-edge elimination on that word), and at N = 10 more than half of a 32 x 32 tile's
// pixels have window cells in the neighbouring tiles.  The generic kernel this replaces for N >= 5 kept the window in
// local-memory arrays and handed those pixels to cl_merge, which re-evaluated their windows on global memory
// (ncu, r02, 8 stress frames of 2208x1242 at N = 10: cl_tile 3.83 ms / 2.3 G warp instructions, cl_merge 0.80 ms).  Here:
//
//   * THE HALO PIXELS ARE NODES of the tile's union-find.  Labels live in the staged index space (32+N)^2 and every
//     edge of a core pixel -- in-tile or cross-tile -- is settled in shared memory.  A component's representative towards the later kernels is its smallest CORE pixel (rep); every
//     halo pixel h that ended up in a component yields ONE pair (h, rep) for cl_merge_pairs, which unions the two tile
//     roots in global memory (de-duplicated per warp).  No window is ever re-evaluated on global memory and no
//     implied-edge elimination is needed: the number of global unions is bounded by the halo size, not by the edges.
//     Every shared-memory union corresponds to an accepted edge of a core pixel, every accepted edge (a, b) is seen
//     by the tile that holds a in its core, so the global partition equals the reference's.
//   * HALO PIXELS START AS LEAVES.  Links (C1) only go to core cells -- z is staged twice, s_zc with +inf in the halo
//     -- so the link forest's roots are core pixels; then every halo pixel looks for its nearest accepted core pixel
//     (H) and takes that pixel's root as its label.  A halo pixel is nobody's parent at that point, so this needs no
//     atomics, and on smooth depth the whole tile is one flat tree before the first compare.  (First version: links
//     into the halo made the up to 740 halo pixels roots that had to be unioned with one another through core pixels,
//     smallest index wins -- chains hundreds of links deep; ncu on the N = 10 stress frames: 38 % of the instructions
//     in finds and unions.)  Edges from a core pixel to a halo pixel of another tree are found by C2 like any other.
//   * LAZY DEPTH GATES.  C1 evaluates the window once, branch-free (one shared load, one subtract, one compare per
//     cell; non-dynamic cells are staged as z = +inf) and keeps only the two raster-smallest accepted cells: the LINK
//     (parent pointer to an earlier pixel: a forest, no atomics) and one more for the first merge round (C2a, one union
//     per pixel against that neighbour's flattened label, de-duplicated per warp).  After the second flatten, C2
//     compares LABELS cell by cell (shared load + compare) and evaluates the depth gate only for window rows in which
//     some lane still sees a foreign label; whatever is accepted there and still differs is unioned.  An edge whose
//     endpoints carry equal labels needs no gate: equal labels mean one set already.
//
// Correctness of C2 in one pass: labels only ever name a member of the pixel's own set, so "labels equal" implies
// "same set"; every accepted edge of every listed pixel is either the link, seen with equal labels, or unioned.

#ifndef CL_WIDE_MIN_BLOCKS
#define CL_WIDE_MIN_BLOCKS 6
#endif

template <int NC>
__global__ void __launch_bounds__(kBlock, CL_WIDE_MIN_BLOCKS)
    cl_tile_wide_kernel(const ClLaunch L, uint32_t* __restrict__ list, uint32_t* __restrict__ list_speed,
                        uint32_t* __restrict__ pair_halo, uint32_t* __restrict__ pair_rep,
                        uint32_t* __restrict__ root_list, int* __restrict__ counters /* list, pairs, root */) {
  pdl_enter();
  static_assert(NC >= 1 && NC <= kClMaxNeighborDistance, "window of up to 11 x 11 cells");
  constexpr int S = NC + 1;
  constexpr int SW = kTileW + NC;  // staged row pitch
  constexpr int SH = kTileH + NC;
  constexpr int kCells = SH * SW;
  constexpr int kCore = NC * SW + NC;  // staged index of the tile's first core pixel; window cell (wx, wy) of the
                                       // pixel staged at p sits at p - kCore + wy * SW + wx
  constexpr int kHalo = NC * SW + kTileH * NC;
  constexpr int kNone = 0x7fffffff;
  __shared__ float s_z[kCells];        // z of every staged dynamic pixel, +inf elsewhere
  __shared__ float s_zc[kCells];       // the same with +inf in the halo: what C1 links against
  __shared__ int s_lab[kCells];        // -1: not dynamic; otherwise a staged index of the same set
  __shared__ int s_rep[kCells];        // C1..C2a: second accepted cell of a core pixel; later: min core pixel per root
  __shared__ uint8_t s_up[kCells];     // core pixel has an accepted up-left neighbour
  __shared__ uint16_t s_dl[kTileW * kTileH];
  __shared__ uint16_t s_rl[kTileW * kTileH];
  __shared__ uint16_t s_hl[kHalo];
  __shared__ int s_nd, s_hcnt, s_rcnt, s_base, s_hbase, s_rbase;
  const int frame = blockIdx.z;
  const int tx0 = blockIdx.x * kTileW, ty0 = blockIdx.y * kTileH;
  const size_t P = static_cast<size_t>(L.w) * L.h;
  const size_t fb = frame * P;
  const uint8_t* mask = L.mask + fb;
  const float* pts = L.points + frame * L.points_frame_stride * 8;
  const int lane = threadIdx.x & 31;
  const float thr = L.depth_thr;
  const bool thr_ok = thr < CUDART_INF_F;  // false for NaN too: then the +inf staging cannot stand in for "not dynamic"

  if (threadIdx.x == 0) s_nd = s_hcnt = s_rcnt = 0;

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
    s_lab[si] = dyn ? si : -1;
    const float z = dyn ? __ldg(pts + (static_cast<size_t>(y) * L.w + x0 + j) * 8 + 2) : CUDART_INF_F;
    s_z[si] = z;
    s_zc[si] = z;
  }
  // halo: NC rows above (full width 32+NC) and NC columns to the left of the core
  constexpr int top = NC * SW;
  for (int k = threadIdx.x; k < kHalo; k += kBlock) {
    int sy, sx;
    if (k < top) {
      sy = k / SW;
      sx = k - sy * SW;
    } else {
      const int k2 = k - top;
      sy = NC + k2 / NC;
      sx = k2 - (k2 / NC) * NC;
    }
    const int yy = ty0 - NC + sy, xx = tx0 - NC + sx;
    bool m = false;
    if (xx >= 0 && yy >= 0 && xx < L.w && yy < L.h) m = mask[static_cast<size_t>(yy) * L.w + xx] != 0;
    const int si = sy * SW + sx;
    s_lab[si] = m ? si : -1;
    s_z[si] = m ? __ldg(pts + (static_cast<size_t>(yy) * L.w + xx) * 8 + 2) : CUDART_INF_F;
    s_zc[si] = CUDART_INF_F;
  }
  __syncthreads();

  // B: the tile-local list of dynamic core pixels, a row per warp (32 consecutive entries are mostly 32 consecutive
  // pixels of a row: conflict-free shared-memory accesses in every later phase)
  {
    const int warp = threadIdx.x >> 5;
#pragma unroll
    for (int it = 0; it < kTileH / (kBlock / 32); it++) {
      const int r = it * (kBlock / 32) + warp;
      const int p = (r + NC) * SW + NC + lane;
      smem_append(s_lab[p] >= 0, static_cast<uint16_t>(p), s_dl, &s_nd);
    }
  }
  __syncthreads();
  const int nd = s_nd;

  // depth gate of the pixel staged at p (z_here) against the cell staged at si
  auto gate_fast = [&](float z_here, int si) -> bool { return !(fabsf(__fsub_rn(z_here, s_z[si])) > thr); };
  auto gate_flag = [&](float z_here, int si) -> bool { return s_lab[si] >= 0 && depth_gate(z_here, s_z[si], thr); };

  // halo cell k (0 .. kHalo-1) -> staged index
  auto halo_cell = [&](int k) -> int {
    int sy, sx;
    if (k < top) {
      sy = k / SW;
      sx = k - sy * SW;
    } else {
      const int k2 = k - top;
      sy = NC + k2 / NC;
      sx = k2 - (k2 / NC) * NC;
    }
    return sy * SW + sx;
  };

  // C1: link to the raster-smallest accepted CORE neighbour; remember the next one for C2a
  for (int e = threadIdx.x; e < nd; e += kBlock) {
    const int p = s_dl[e];
    const float z_here = s_z[p];
    const int base = p - kCore;
    const int lx = p % SW - NC, ly = p / SW - NC;
    int first = -1, second = -1;
    const bool fast = thr_ok && fabsf(z_here) < CUDART_INF_F;
    if (fast) {
      // a non-dynamic or halo cell holds +inf: |finite - inf| = inf > thr rejects it.  Reverse raster order: the last hit
      // is the raster-smallest one.
#pragma unroll
      for (int wy = NC; wy >= 0; wy--) {
#pragma unroll
        for (int wx = NC; wx >= 0; wx--) {
          if (wy == NC && wx == NC) continue;
          const int si = base + wy * SW + wx;
          if (!(fabsf(__fsub_rn(z_here, s_zc[si])) > thr)) {
            second = first;
            first = si;
          }
        }
      }
    } else {  // z_here not finite (foreign clouds) or no usable threshold: test the dynamic flag explicitly
#pragma unroll 1
      for (int k = S * S - 2; k >= 0; k--) {
        const int wy = k / S, wx = k % S;
        if (lx - NC + wx < 0 || ly - NC + wy < 0) continue;  // halo cell
        const int si = base + wy * SW + wx;
        if (gate_flag(z_here, si)) {
          second = first;
          first = si;
        }
      }
    }
    bool up = first >= 0;
    if (!up && (lx < NC || ly < NC)) {  // no accepted core cell: the key flag still needs the halo cells (rare)
      const int hy = max(NC - ly, 0), hx = max(NC - lx, 0);  // window rows < hy are halo rows, columns < hx halo columns
#pragma unroll 1
      for (int wy = 0; wy < S && !up; wy++) {
        const int n = wy < hy ? S : hx;  // hx <= NC: the pixel's own cell (NC, NC) is never reached
#pragma unroll 1
        for (int wx = 0; wx < n; wx++) up = up || gate_flag(z_here, base + wy * SW + wx);
      }
    }
    s_up[p] = up;
    s_rep[p] = second;
    s_lab[p] = first >= 0 ? first : p;  // only p's own entry changes here and nobody reads a label's VALUE in C1
  }
  __syncthreads();

  for (int e = threadIdx.x; e < nd; e += kBlock) {  // F
    const int p = s_dl[e];
    s_lab[p] = smem_find_readonly(s_lab, p);
  }
  __syncthreads();
  // H: every dynamic halo pixel takes the label (a root: the labels are flat) of its nearest accepted core pixel.  The
  // halo pixel is still a singleton nobody points to, so a plain store does it.  A halo pixel that touches several
  // trees joins one here; C2 finds the others.
  for (int k = threadIdx.x; k < kHalo; k += kBlock) {
    const int h = halo_cell(k);
    if (s_lab[h] < 0) continue;
    const float zh = s_z[h];
    const int sy = h / SW, sx = h - sy * SW;
    const int cy0 = max(sy, NC), cx0 = max(sx, NC);
    const int cy1 = min(sy + NC, SH - 1), cx1 = min(sx + NC, SW - 1);
    int found = -1;
#pragma unroll 1
    for (int cy = cy0; cy <= cy1 && found < 0; cy++) {
#pragma unroll 1
      for (int cx = cx0; cx <= cx1; cx++) {
        const int c = cy * SW + cx;
        if (s_lab[c] >= 0 && depth_gate(s_z[c], zh, thr)) {
          found = c;
          break;
        }
      }
    }
    if (found >= 0) s_lab[h] = s_lab[found];
  }
  __syncthreads();
  for (int e0 = 0; e0 < nd; e0 += kBlock) {  // C2a: one union per pixel, one lane per distinct (root, root) pair
    const int e = e0 + threadIdx.x;
    int my = -1, other = -1;
    if (e < nd) {
      const int p = s_dl[e];
      my = s_lab[p];
      const int sec = s_rep[p];
      if (sec >= 0) {
        const int lb = s_lab[sec];
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
    const int p = s_dl[e];
    s_lab[p] = smem_find_readonly(s_lab, p);
  }
  for (int k = threadIdx.x; k < kHalo; k += kBlock) {  // the halo leaves follow their trees
    const int h = halo_cell(k);
    if (s_lab[h] >= 0) s_lab[h] = smem_find_readonly(s_lab, h);
  }
  for (int k = threadIdx.x; k < kCells; k += kBlock) s_rep[k] = kNone;  // the second-cell slots are free now
  __syncthreads();

  // One component in the tile (the usual case inside a blob and on smooth depth): every core pixel and every halo pixel
  // that joined carries the same label, so every accepted edge joins equal labels and C2 has nothing to find.  A halo
  // pixel that did NOT join (label = itself: halo pixels are leaves or singletons, never parents) has no accepted edge
  // into this tile's core -- H scanned all of its candidates -- and is ignored.
  bool one_label = true;
  {
    const int first = s_lab[s_dl[nd >> 1]];
    for (int e = threadIdx.x; e < nd; e += kBlock) one_label = one_label && (s_lab[s_dl[e]] == first);
    for (int k = threadIdx.x; k < kHalo; k += kBlock) {
      const int h = halo_cell(k);
      const int l = s_lab[h];
      one_label = one_label && (l < 0 || l == h || l == first);
    }
  }
  const bool skip_compare = __syncthreads_and(one_label);

  // C2: compare labels over the whole window; depth gates only where a foreign label shows up
  bool unioned = false;
  for (int e0 = 0; e0 < (skip_compare ? 0 : nd); e0 += kBlock) {
    const int e = e0 + threadIdx.x;
    int p = kCore, my = 0;
    float z_here = 0.0f;
    bool fast = true;
    if (e < nd) {
      p = s_dl[e];
      my = s_lab[p];
      z_here = s_z[p];
      fast = thr_ok && fabsf(z_here) < CUDART_INF_F;
    }
    const int base = p - kCore;
#pragma unroll 1
    for (int wy = 0; wy <= NC; wy++) {
      const int rb = base + wy * SW;
      uint32_t m = 0u;
#pragma unroll
      for (int wx = 0; wx <= NC; wx++) {
        // foreign label: l >= 0 and l != my.  my >= 0, so (l ^ my) is 0 for an equal label, negative for l = -1 (not
        // dynamic) and positive exactly for a foreign one
        if ((s_lab[rb + wx] ^ my) > 0) m |= 1u << wx;
      }
      if (wy == NC) m &= ~(1u << NC);  // the pixel itself
      if (e >= nd) m = 0u;
      if (!__any_sync(kFull, m != 0u)) continue;
      if (fast) {
#pragma unroll
        for (int wx = 0; wx <= NC; wx++)
          if (!gate_fast(z_here, rb + wx)) m &= ~(1u << wx);
      } else {
#pragma unroll 1
        for (int wx = 0; wx <= NC; wx++)
          if (((m >> wx) & 1u) && !gate_flag(z_here, rb + wx)) m &= ~(1u << wx);
      }
      unioned = unioned || m != 0u;
      while (__any_sync(kFull, m != 0u)) {
        if (m) {
          const int wx = __ffs(m) - 1;
          m &= m - 1u;
          smem_union(s_lab, p, rb + wx);
        }
        __syncwarp();
      }
    }
  }
  const bool reflatten = __syncthreads_or(unioned);  // labels are flat unless C2 still had to union

  // F + representatives: the smallest core pixel of every component
  for (int e = threadIdx.x; e < nd; e += kBlock) {
    const int p = s_dl[e];
    int r = s_lab[p];
    if (reflatten) {
      r = smem_find_readonly(s_lab, p);  // no halving stores here: see smem_find_readonly
      s_lab[p] = r;
    }
    atomicMin(&s_rep[r], p);
  }
  __syncthreads();

  // D1: representatives initialise their partial record; halo pixels that joined a component are listed
  for (int e0 = 0; e0 < nd; e0 += kBlock) {
    const int e = e0 + threadIdx.x;
    bool is_rep = false;
    int p = 0;
    if (e < nd) {
      p = s_dl[e];
      is_rep = s_rep[s_lab[p]] == p;
      if (is_rep) {
        const size_t gi = fb + static_cast<size_t>(ty0 + p / SW - NC) * L.w + tx0 + p % SW - NC;
        L.count[gi] = 0;
        TileStat init;
        init.key = 0x7fffffff;
        init.mn[0] = init.mn[1] = init.mn[2] = 0xffffffffu;
        init.mx[0] = init.mx[1] = init.mx[2] = 0u;
        init.pad = 0;
        L.tstat[gi] = init;
      }
    }
    smem_append(is_rep, static_cast<uint16_t>(p), s_rl, &s_rcnt);
  }
  for (int k0 = 0; k0 < kHalo; k0 += kBlock) {
    const int k = k0 + threadIdx.x;
    bool joined = false;
    int si = 0;
    if (k < kHalo) {
      si = halo_cell(k);
      // a halo pixel's set holds a core pixel iff some core pixel linked to it or was unioned with it
      if (s_lab[si] >= 0) joined = s_rep[smem_find_readonly(s_lab, si)] != kNone;
    }
    smem_append(joined, static_cast<uint16_t>(si), s_hl, &s_hcnt);
  }
  __syncthreads();
  const int nh = s_hcnt, nr = s_rcnt;
  if (threadIdx.x == 0) {
    s_base = atomicAdd(counters + 0, nd);
    s_hbase = nh ? atomicAdd(counters + 1, nh) : 0;
    s_rbase = atomicAdd(counters + 2, nr);
    atomicAdd(L.n_dynamic + frame, nd);
  }
  __syncthreads();

  // D2: publish parent = representative, list entries, and the per-representative partial statistics
  // (pcl::getMinMax3D operands, size, reference-order key), reduced per representative inside each warp before the
  // atomics
  for (int e0 = 0; e0 < nd; e0 += kBlock) {
    const int e = e0 + threadIdx.x;
    int r = -1 - lane;  // distinct for idle lanes
    uint32_t ex_mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, ex_mx[3] = {0u, 0u, 0u};
    int key = 0x7fffffff;
    if (e < nd) {
      const int p = s_dl[e];
      r = s_rep[s_lab[p]];
      const int i = (ty0 + p / SW - NC) * L.w + tx0 + p % SW - NC;
      L.parent[fb + i] = (ty0 + r / SW - NC) * L.w + tx0 + r % SW - NC;
      list[s_base + e] = static_cast<uint32_t>(fb + i);
      const float4 xyz = __ldg(reinterpret_cast<const float4*>(pts + static_cast<size_t>(i) * 8));
      const float4 vel = __ldg(reinterpret_cast<const float4*>(pts + static_cast<size_t>(i) * 8 + 4));
      list_speed[s_base + e] = __float_as_uint(__fsqrt_rn(
          __fadd_rn(__fadd_rn(__fmul_rn(vel.x, vel.x), __fmul_rn(vel.y, vel.y)), __fmul_rn(vel.z, vel.z))));
      const float cc[3] = {xyz.x, xyz.y, xyz.z};
#pragma unroll
      for (int k = 0; k < 3; k++)  // NaN coordinates are skipped (order dependent in the reference; unpinned)
        if (!isnan(cc[k])) ex_mn[k] = ex_mx[k] = float_to_ordered(cc[k]);
      if (s_up[p]) key = i;
    }
    unsigned todo = __ballot_sync(kFull, r >= 0);
    while (todo) {  // peel one representative at a time; usually one or two per warp
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
          const size_t gr = fb + static_cast<size_t>(ty0 + lr / SW - NC) * L.w + tx0 + lr % SW - NC;
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
  for (int e = threadIdx.x; e < nh; e += kBlock) {
    const int h = s_hl[e];
    const int r = s_rep[smem_find_readonly(s_lab, h)];
    pair_halo[s_hbase + e] = static_cast<uint32_t>(fb + static_cast<size_t>(ty0 + h / SW - NC) * L.w + tx0 + h % SW - NC);
    pair_rep[s_hbase + e] = static_cast<uint32_t>(fb + static_cast<size_t>(ty0 + r / SW - NC) * L.w + tx0 + r % SW - NC);
  }
  for (int e = threadIdx.x; e < nr; e += kBlock) {
    const int p = s_rl[e];
    root_list[s_rbase + e] = static_cast<uint32_t>(fb + static_cast<size_t>(ty0 + p / SW - NC) * L.w + tx0 + p % SW - NC);
  }
}

// cl_merge for cl_tile_wide: one (halo pixel, representative) pair per entry.  The halo pixel's parent entry holds its
// own tile's root (or, after path halving, an ancestor of it); consecutive entries come from one tile and mostly name
// the same two roots, so the pairs are de-duplicated inside the warp before the lock-free union.
__global__ void __launch_bounds__(kBlock)
    cl_merge_pairs_kernel(const ClLaunch L, const uint32_t* __restrict__ pair_halo, const uint32_t* __restrict__ pair_rep,
                          const int* __restrict__ pair_count) {
  pdl_enter();
  const int n = *pair_count;
  const uint32_t P = static_cast<uint32_t>(L.w) * L.h;
  for (int t0 = blockIdx.x * kBlock; t0 < n; t0 += gridDim.x * kBlock) {
    const int t = t0 + threadIdx.x;
    const bool on = t < n;
    int ra = 0, rb = 0;
    int* parent = L.parent;
    if (on) {
      const uint32_t a = pair_halo[t], b = pair_rep[t];
      const uint32_t frame = a / P;
      parent = L.parent + static_cast<size_t>(frame) * P;
      ra = __ldcg(parent + (a - frame * P));
      rb = static_cast<int>(b - frame * P);
    }
    const unsigned active = __ballot_sync(kFull, on);
    if (on) {
      const unsigned long long key =
          (static_cast<unsigned long long>(static_cast<uint32_t>(parent - L.parent) + static_cast<uint32_t>(ra)) << 32) |
          static_cast<uint32_t>(rb);
      const unsigned same = __match_any_sync(active, key);
      if ((threadIdx.x & 31) == __ffs(same) - 1 && ra != rb) uf_union(parent, ra, rb);
    }
    __syncwarp();
  }
}
