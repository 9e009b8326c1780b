// cluster_fast_halo.cuh -- cl_tile_halo: cl_tile_fast_kernel (neighbor_distance 1..4) with the halo pixels as LEAVES of
// the tile's forest.  Included by cluster.cu inside its anonymous namespace (uses its helpers).
//
// Same tile-local algorithm as cl_tile_fast_kernel (window bit sets, link forest, one merge round, label-compare stencil;
// see there), different treatment of the edges that leave the tile.  cl_tile_fast_kernel ran the implied-edge
// elimination of window_edges on them (a band of left / up flags, a second C1 loop) and handed the surviving window
// cells to cl_merge, which looked the neighbours' tile roots up in global memory.  Here, as in cl_tile_wide_kernel:
//   * labels live in the STAGED index space (tile + halo); links still only go to core cells, so roots are core pixels;
//   * after the first flatten every dynamic halo pixel joins the tree of its nearest accepted core pixel as a leaf (H);
//   * an edge pixel's accepted halo cells are kept as one more bit set (s_bx) and confirmed in the compare round by one
//     shared load per cell -- unions only where the labels still differ;
//   * every joined halo pixel yields ONE (halo pixel, tile root) pair for cl_merge_pairs.
// Every accepted edge (a, b) is seen by the tile that holds a in its core, and every shared-memory union or leaf store
// corresponds to an accepted edge of a core pixel, so the global partition is the reference's.

template <int NC>
__global__ void __launch_bounds__(kBlock, CL_TILE_MIN_BLOCKS)
    cl_tile_halo_kernel(const ClLaunch L, uint32_t* __restrict__ list, uint32_t* __restrict__ list_speed,
                        uint32_t* __restrict__ pair_halo, uint32_t* __restrict__ pair_rep,
                        uint32_t* __restrict__ root_list, int* __restrict__ counters /* list, pairs, root */) {
  static_assert(NC >= 1 && NC <= 4, "fast path covers neighbor_distance 1..4");
  constexpr int S = NC + 1;
  constexpr int SW = kTileW + NC;  // shared-memory row pitch
  constexpr int kPixels = kTileW * kTileH;
  // A core pixel is identified by c = ly * SW + lx: the staged arrays' pitch, so that the window cell (wx, wy) of c is
  // staged at c + wy * SW + wx, and so that neither 32 consecutive pixels of a row nor the NC leftmost pixels of
  // consecutive rows (the two orders the lists below come in) collide on a shared-memory bank.
  constexpr int kIds = kTileH * SW;
  constexpr int kCore = NC * SW + NC;  // staged index of core pixel c is c + kCore
  __shared__ __align__(16) uint8_t s_m[(kTileH + NC) * SW];
  __shared__ float s_z[(kTileH + NC) * SW];
  constexpr int kHalo = NC * SW + kTileH * NC;
  __shared__ int s_lab[(kTileH + NC) * SW];  // STAGED index space (core pixel c sits at c + kCore): halo pixels are nodes
  __shared__ uint8_t s_up[kIds];      // pixel has an accepted up-left neighbour
  __shared__ uint32_t s_E[kIds];      // accepted in-tile window cells minus the link
  __shared__ uint16_t s_dl[kPixels];  // interior pixels from the front, edge pixels from the back
  constexpr int kEdgePixels = kPixels - (kTileW - NC) * (kTileH - NC);  // pixels within NC of the top / left tile edge
  __shared__ uint32_t s_bx[kEdgePixels];  // accepted HALO cells of the e-th edge pixel
  __shared__ uint16_t s_hl[kHalo];        // halo pixels that joined a tree of the tile
  __shared__ uint16_t s_rl[kPixels];
  __shared__ int s_ni, s_ne, s_hcnt, s_rcnt, s_base, s_hbase, s_rbase, s_first;
  const int frame = blockIdx.z;
  const int tx0 = blockIdx.x * kTileW, ty0 = blockIdx.y * kTileH;
  const size_t P = static_cast<size_t>(L.w) * L.h;
  const size_t fb = frame * P;
  const uint8_t* mask = L.mask + fb;
  const float* pts = L.points + frame * L.points_frame_stride * 8;
  const int lane = threadIdx.x & 31;
  const float thr = L.depth_thr;
  const bool thr_ok = thr < CUDART_INF_F;  // false for NaN too: then the +inf staging cannot stand in for "not dynamic"

  if (threadIdx.x == 0) s_ni = s_ne = s_hcnt = s_rcnt = 0;

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
    s_lab[sy * SW + sx] = sy * SW + sx;  // a halo pixel starts as a singleton
  }
  __syncthreads();

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
    s_lab[c + kCore] = c + kCore + (acc ? cell_offset(__ffs(acc) - 1) : 0);  // lowest bit = raster-smallest neighbour
    s_E[c] = acc & (acc - 1);
  }
  // C1, edge pixels: links only go to core cells; the accepted halo cells are kept for the compare round
  for (int e = threadIdx.x; e < ne; e += kBlock) {
    const int c = pixel_at(ni + e);
    const int ly = c / SW, lx = c - ly * SW;
    const uint32_t acc = accepted_cells(c, s_z[c + kCore]);
    // a cell (wx,wy) lies in the halo iff lx-N+wx < 0 or ly-N+wy < 0
    uint32_t inside = 0;
#pragma unroll
    for (int wy = 0; wy < S; wy++) {
      if (ly - NC + wy < 0) continue;
      const int first = max(NC - lx, 0);  // wx >= first stays inside
      if (first < S) inside |= ((((1u << S) - 1u) >> first) << first) << (wy * S);
    }
    const uint32_t local = acc & inside;
    s_up[c] = acc != 0;
    s_lab[c + kCore] = c + kCore + (local ? cell_offset(__ffs(local) - 1) : 0);
    s_E[c] = local & (local - 1);
    s_bx[e] = acc & ~inside;
  }
  __syncthreads();

  for (int e = threadIdx.x; e < nd; e += kBlock) {  // F
    const int p = pixel_at(e) + kCore;
    s_lab[p] = smem_find_readonly(s_lab, p);
  }
  __syncthreads();
  // H: every dynamic halo pixel becomes a LEAF of the tree of its nearest accepted core pixel (the labels are flat, so
  // that pixel's label is a root; nobody points to the halo pixel yet, so a plain store does it) and is listed for the
  // (halo pixel, tile root) pairs cl_merge_pairs unions in global memory.  The scan is exhaustive: a halo pixel that
  // stays a singleton has no accepted edge into this tile's core.  A halo pixel that touches several trees joins one
  // here; the edge pixels' halo cells (s_bx) settle the others in the compare round.
  for (int k0 = 0; k0 < kHalo; k0 += kBlock) {
    const int k = k0 + threadIdx.x;
    int h = 0;
    bool joined = false;
    if (k < kHalo) {
      int sy, sx;
      if (k < top) {
        sy = k / (kTileW + NC);
        sx = k - sy * (kTileW + NC);
      } else {
        const int k2 = k - top;
        sy = NC + k2 / NC;
        sx = k2 - (k2 / NC) * NC;
      }
      h = sy * SW + sx;
      if (s_m[h] & kDyn) {
        const float zh = s_z[h];
        const int cy1 = min(sy + NC, kTileH + NC - 1), cx0 = max(sx, NC), cx1 = min(sx + NC, SW - 1);
#pragma unroll 1
        for (int cy = max(sy, NC); cy <= cy1 && !joined; cy++) {
#pragma unroll 1
          for (int cx = cx0; cx <= cx1; cx++) {
            const int p = cy * SW + cx;
            if ((s_m[p] & kDyn) && depth_gate(s_z[p], zh, thr)) {
              s_lab[h] = s_lab[p];
              joined = true;
              break;
            }
          }
        }
      }
    }
    smem_append(joined, static_cast<uint16_t>(h), s_hl, &s_hcnt);
  }
  __syncthreads();
  const int nh = s_hcnt;
  for (int e0 = 0; e0 < nd; e0 += kBlock) {  // C2a: one union per pixel, one lane per distinct (root, root) pair
    const int e = e0 + threadIdx.x;
    int my = -1, other = -1;
    if (e < nd) {
      const int c = pixel_at(e);
      my = s_lab[c + kCore];
      // only the first remaining accepted neighbour is looked at: the trees of the link forest interleave, so most
      // pixels see another tree there, and any pair this round misses is settled by C2b
      const uint32_t A = s_E[c];
      if (A) {
        const int lb = s_lab[c + kCore + cell_offset(__ffs(A) - 1)];
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
    const int p = pixel_at(e) + kCore;
    const int r = smem_find_readonly(s_lab, p);
    s_lab[p] = r;
    if (e == (nd >> 1)) s_first = r;  // a pixel from the middle of the list: entry 0 is often an isolated speck
  }
  for (int e = threadIdx.x; e < nh; e += kBlock) {  // the halo leaves follow their trees
    const int h = s_hl[e];
    s_lab[h] = smem_find_readonly(s_lab, h);
  }
  __syncthreads();
#ifndef CL_TILE_NO_UNIFORM_SKIP
  // One component in the tile (the usual case inside a blob): every accepted in-tile edge joins equal labels, so
  // the compare round has nothing to find.
  bool one_label = true;
  {
    const int first = s_first;
    for (int e = threadIdx.x; e < nd; e += kBlock) one_label = one_label && (s_lab[pixel_at(e) + kCore] == first);
  }
  const bool skip_compare = __syncthreads_and(one_label);
#else
  const bool skip_compare = false;
#endif
  // Most tiles that fail the one-label test hold one big tree plus isolated specks.  When every joined halo pixel
  // carries the big tree's label, an edge pixel of that tree has nothing to confirm among its halo cells.
  bool halo_all_first = false;
  if (!skip_compare) {
    bool ok = true;
    const int first = s_first;
    for (int e = threadIdx.x; e < nh; e += kBlock) ok = ok && (s_lab[s_hl[e]] == first);
    halo_all_first = __syncthreads_and(ok);
  }
  bool unioned = false;
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
        lr[j] = rr >= 0 ? s_lab[rr * SW + lane + kCore] : -1;  // non-dynamic pixels hold stale values: never selected below
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
            smem_union(s_lab, c + kCore, c + kCore + cell_offset(k));
          }
          __syncwarp();
        }
      }
    }
    // the edge pixels' accepted halo cells: one shared load per cell confirms equal labels
    for (int e0 = 0; e0 < ne; e0 += kBlock) {
      const int e = e0 + threadIdx.x;
      int p = 0, my = 0;
      uint32_t d = 0u;
      if (e < ne) {
        p = pixel_at(ni + e) + kCore;
        my = s_lab[p];
        d = (halo_all_first && my == s_first) ? 0u : s_bx[e];
      }
      while (__any_sync(kFull, d != 0u)) {
        if (d) {
          const int k = __ffs(d) - 1;
          d &= d - 1;
          const int h = p + cell_offset(k);
          if (s_lab[h] != my) {
            smem_union(s_lab, p, h);
            unioned = true;
          }
        }
        __syncwarp();
      }
    }
  }
  const bool reflatten = __syncthreads_or(unioned);  // labels are flat unless C2b still had to union

  // D1: flatten the tile-local forest; roots initialise their partial record
  for (int e0 = 0; e0 < nd; e0 += kBlock) {
    const int e = e0 + threadIdx.x;
    bool is_root = false;
    int c = 0;
    if (e < nd) {
      c = pixel_at(e);
      int r = s_lab[c + kCore] - kCore;  // the root of a core pixel is a core pixel: halo pixels are only ever leaves
      if (reflatten) {
        r = smem_find_readonly(s_lab, c + kCore) - kCore;  // no halving stores here: see smem_find_readonly
        s_lab[c + kCore] = r + kCore;
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
  const int nr = s_rcnt;
  if (threadIdx.x == 0) {
    s_base = atomicAdd(counters + 0, nd);
    s_hbase = nh ? atomicAdd(counters + 1, nh) : 0;
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
      r = s_lab[c + kCore] - kCore;
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
  for (int e = threadIdx.x; e < nh; e += kBlock) {
    const int h = s_hl[e];
    const int r = smem_find_readonly(s_lab, h);  // a core pixel: this tile's root of the component h joined
    pair_halo[s_hbase + e] = static_cast<uint32_t>(fb + static_cast<size_t>(ty0 + h / SW - NC) * L.w + tx0 + h % SW - NC);
    pair_rep[s_hbase + e] = static_cast<uint32_t>(fb + static_cast<size_t>(ty0 + r / SW - NC) * L.w + tx0 + r % SW - NC);
  }
  for (int e = threadIdx.x; e < nr; e += kBlock) {
    const int c = s_rl[e];
    root_list[s_rbase + e] = static_cast<uint32_t>(fb + static_cast<size_t>(ty0 + c / SW) * L.w + tx0 + c % SW);
  }
}

