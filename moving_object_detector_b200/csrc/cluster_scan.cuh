// cluster_scan.cuh -- cl_scan: the tile-local part of the clusterer for neighbor_distance 1..4 as a warp-sequential
// raster scan.  Included by cluster.cu inside its anonymous namespace (uses its helpers).
//
// Same contract as cl_tile_kernel (same outputs: parent = tile root, per-tile-root partial statistics, the three
// compact lists), different machine mapping.  The list-driven tile kernel it replaces spent its time in ~14 CTA-wide
// phases separated by barriers (ncu, r02: 325 M warp instructions per 63 frames of 1280x720, flat profile: lists,
// finds, flattens and label compares of a link forest with several trees per tile).  Here ONE WARP owns a tile of
// (32 - N) x TH pixels and walks it top to bottom, one pixel row per step, lane = column:
//
//   * lanes 0..N-1 carry the N halo columns to the left of the tile, so the neighbour at (dx, dy) of every core pixel
//     is "row -dy, lane - dx": z of the current and the N previous rows sits in registers, a window cell costs one
//     warp shuffle (shared by the row), one subtract and one compare -- no shared memory, no lists, no CTA barrier;
//   * the dynamic flags are first packed to one bit per pixel (32 rows of the tile in the 32 lanes: that load is also
//     the empty-tile test), and everything that is a property of a window cell -- neighbour is dynamic, neighbour's
//     unit edges (for the implied-edge elimination of the cross-tile edges), neighbour's label differs from the
//     tile's dominant label -- is a per-lane bit mask in window layout that SLIDES down one window row per step
//     (shift + insert the new row's field): three instructions per mask and row;
//   * labels are the reference's own algorithm (raster scan + label equivalences, clusterer_nodelet.cpp:56-83,
//     186-219), which on a blob interior degenerates to "every accepted neighbour already carries the tile's
//     dominant root label L0": a row takes the FAST PATH when every dynamic pixel either has an accepted in-tile
//     neighbour in the rows above and all its accepted in-tile neighbours carry L0 (it gets L0: one shared store,
//     no loads), or has no accepted in-tile neighbour at all and is nobody's accepted neighbour in its own row (a new
//     singleton label).  Any other row takes the SLOW PATH: link to the first accepted neighbour above, compare the
//     labels of the other accepted in-tile neighbours, union what differs (shared-memory union-find), re-flatten the
//     rows still inside the window, re-derive L0 and the label mask.  The rows inside the window are flat (labels
//     are roots) at every row boundary, so "labels equal" <=> "same set so far";
//   * the publish pass (second walk over the rows) resolves every pixel's final tile root, writes parent / list /
//     speed and accumulates the per-root statistics per lane across rows, reduced once per distinct root and tile.
//
// Correctness argument: every accepted in-tile edge (a, b), b up-left of a, is examined when a's row is processed --
// either both labels are known equal (fast path: L0; slow path: link or equal roots) or the pair is unioned; so the
// tile-local partition equals the connected components of the accepted in-tile edges, as in cl_tile_kernel.  Cross-
// tile edges are listed for cl_merge after the implied-edge elimination of window_edges (same rule, same flags).

#ifndef CL_SCAN_WARPS
#define CL_SCAN_WARPS 4  // tiles (warps) per CTA, side by side
#endif
#ifndef CL_SCAN_ROWS
#define CL_SCAN_ROWS 32  // pixel rows per tile (<= 32: one row-bit word per lane)
#endif
#ifndef CL_SCAN_MIN_BLOCKS
#define CL_SCAN_MIN_BLOCKS 8
#endif

// Dynamic flags of columns xs .. xs+31 of image row y as one word (bit b = column xs + b), zero outside the image.
// The row's bytes are fetched as aligned 32-bit words and funnel-shifted into place (any width, any byte offset).
__device__ __forceinline__ uint32_t scan_row_bits(const uint8_t* __restrict__ maskf, int w, int h, int y, int xs) {
  const int c0 = max(xs, 0), c1 = min(xs + 32, w);
  if (y < 0 || y >= h || c1 <= c0) return 0u;
  const int n = c1 - c0;
  const uint8_t* p = maskf + static_cast<size_t>(y) * w + c0;
  const uint32_t sh = (reinterpret_cast<uintptr_t>(p) & 3u) * 8u;
  const uint32_t* al = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(p) & ~static_cast<uintptr_t>(3));
  const int n_words = (n + static_cast<int>(sh >> 3) + 3) >> 2;  // aligned words that hold the n bytes (<= 9)
  uint32_t v[9];
#pragma unroll
  for (int q = 0; q < 9; q++) v[q] = q < n_words ? __ldg(al + q) : 0u;
  uint32_t out = 0u;
#pragma unroll
  for (int q = 0; q < 8; q++) {
    const uint32_t four = __funnelshift_r(v[q], v[q + 1], sh);
    // bit 7 of every non-zero byte, then the four flags gathered into one nibble by a multiplication
    const uint32_t nz = (four | ((four & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u;
    out |= (((nz >> 7) * 0x10204080u) >> 28) << (4 * q);
  }
  if (n < 32) out &= (1u << n) - 1u;
  return out << (c0 - xs);
}

__device__ __forceinline__ bool smem_union_counted(int* lab, int a, int b) {  // true: a root was linked (one set fewer)
  for (;;) {
    a = smem_find(lab, a);
    b = smem_find(lab, b);
    if (a == b) return false;
    if (a < b) {
      const int t = a;
      a = b;
      b = t;
    }
    const int old = atomicMin(lab + a, b);
    if (old == a) return true;
    a = old;
  }
}

template <int NC>
__global__ void __launch_bounds__(CL_SCAN_WARPS * 32, CL_SCAN_MIN_BLOCKS)
    cl_scan_kernel(const ClLaunch L, uint32_t* __restrict__ list, uint32_t* __restrict__ list_speed,
                   uint32_t* __restrict__ border_list, uint32_t* __restrict__ border_mask,
                   uint32_t* __restrict__ root_list, int* __restrict__ counters /* list, border, root */) {
  pdl_enter();
  static_assert(NC >= 1 && NC <= 4, "one-word window masks cover neighbor_distance 1..4");
  constexpr int S = NC + 1;
  constexpr int CW = 32 - NC;        // core columns of a tile
  constexpr int TH = CL_SCAN_ROWS;   // core rows of a tile
  static_assert(TH >= NC && TH <= 32, "tile rows");
  constexpr int SR = TH + NC;        // staged rows (core + top halo)
  constexpr int kA = NC * S + NC, kL = kA - 1, kU = kA - S;
  constexpr uint32_t kField = (1u << S) - 1u;
  constexpr uint32_t kCells = (1u << kA) - 1u;          // every window cell but the pixel itself
  constexpr uint32_t kUpCells = (1u << (NC * S)) - 1u;  // cells of the rows above
  constexpr int kTop = NC * S;                          // bit position of the window's own row
  constexpr int kMaxBorder = TH * CW - (TH - NC) * (CW - NC);  // pixels within NC of the tile's top / left edge
  // label of staged pixel p = j * 32 + lane (staged row j, lane): a pixel of the same tile, eventually the tile root
  __shared__ int s_lab_all[CL_SCAN_WARPS][SR * 32];
  __shared__ uint32_t s_bmask_all[CL_SCAN_WARPS][kMaxBorder];
  __shared__ uint16_t s_bpix_all[CL_SCAN_WARPS][kMaxBorder];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x0 = (blockIdx.x * CL_SCAN_WARPS + warp) * CW;
  if (x0 >= L.w) return;  // no barriers in this kernel: warps are independent
  const int y0 = blockIdx.y * TH;
  const int xs = x0 - NC, ys = y0 - NC;
  const int frame = blockIdx.z;
  const size_t P = static_cast<size_t>(L.w) * L.h;
  const size_t fb = frame * P;
  const uint8_t* maskf = L.mask + fb;
  const float* pts = L.points + frame * L.points_frame_stride * 8;
  const float thr = L.depth_thr;
  const uint32_t lt_mask = (1u << lane) - 1u;
  constexpr uint32_t kCoreLanes = 0xffffffffu << NC;

  // ---- dynamic bits: core row `lane` in rbc (this is also the empty-tile test), halo row `lane` < NC in rbh ---------
  const uint32_t rbc = lane < TH ? scan_row_bits(maskf, L.w, L.h, y0 + lane, xs) : 0u;
  const int nd = __reduce_add_sync(kFull, __popc(rbc & kCoreLanes));
  if (nd == 0) return;
  const uint32_t rbh = lane < NC ? scan_row_bits(maskf, L.w, L.h, ys + lane, xs) : 0u;
  auto staged_bits = [&](int j) -> uint32_t {  // j is warp-uniform
    return j < NC ? __shfl_sync(kFull, rbh, j) : __shfl_sync(kFull, rbc, j - NC);
  };
  // window field of a row's bit word for this lane: bit wx = column x - NC + wx, zero-filled left of the staged columns
  auto field = [&](uint32_t b) -> uint32_t { return __funnelshift_r(b << NC, b >> (32 - NC), lane) & kField; };

  int* s_lab = s_lab_all[warp];
  uint32_t* s_bmask = s_bmask_all[warp];
  uint16_t* s_bpix = s_bpix_all[warp];

  // per-lane constants: window cells that lie inside the tile's columns; cells with a right neighbour in the window
  uint32_t col_inside = 0u, col_ok = 0u;
  {
    const int first = max(2 * NC - lane, 0);  // wx >= first stays inside (lane - dx >= NC)
    const uint32_t rowbits = first < S ? ((kField >> first) << first) : 0u;
#pragma unroll
    for (int wy = 0; wy < S; wy++) {
      col_inside |= rowbits << (wy * S);
      col_ok |= ((1u << NC) - 1u) << (wy * S);
    }
  }
  auto cell_offset = [](int k) -> int { return (k / S - NC) * 32 + (k % S - NC); };

  // ---- pass 1: the scan ---------------------------------------------------------------------------------------------
  const ptrdiff_t zstride = static_cast<ptrdiff_t>(L.w) * 8;
  const float* zp = pts + (static_cast<ptrdiff_t>(ys) * L.w + xs + lane) * 8 + 2;  // dereferenced only where a bit is set
  auto load_z = [&](int j) -> float {
    float z = 0.0f;
    if (j < SR && ((staged_bits(j) >> lane) & 1u)) z = __ldg(zp + j * zstride);
    return z;
  };
  float zq0 = load_z(0), zq1 = load_z(1);
  float zr[S];  // zr[d]: z of staged row j - d
#pragma unroll
  for (int d = 0; d < S; d++) zr[d] = 0.0f;
  uint32_t dm = 0u, lm = 0u, um = 0u, nlm = 0u;  // sliding window masks: dynamic, left / up unit edge, label != L0
  uint32_t upbits = 0u;  // lane r: core row r's pixels that have an accepted up-left neighbour (reference label order key)
  int L0 = -1;           // the tile's dominant root label (warp-uniform), -1: none yet
  int nbord = 0, nnew = 0, nun = 0;

#pragma unroll 1
  for (int j = 0; j < SR; j++) {
#pragma unroll
    for (int d = S - 1; d > 0; d--) zr[d] = zr[d - 1];
    const float zc = zq0;
    zr[0] = zc;
    zq0 = zq1;
    zq1 = load_z(j + 2);
    const uint32_t bits = staged_bits(j);
    dm = (dm >> S) | (field(bits) << kTop);
    const bool d = (dm >> kA) & 1u;
    uint32_t acc = 0u;
    if (j < NC || (bits & kCoreLanes) == 0u) {  // halo row, or no dynamic core pixel: only unit edges are looked at
      const float zl = __shfl_up_sync(kFull, zc, 1);
      if (d && ((dm >> kL) & 1u) && depth_gate(zc, zl, thr)) acc |= 1u << kL;
      if (d && ((dm >> kU) & 1u) && depth_gate(zc, zr[1], thr)) acc |= 1u << kU;
      lm = (lm >> S) | (field(__ballot_sync(kFull, (acc >> kL) & 1u)) << kTop);
      um = (um >> S) | (field(__ballot_sync(kFull, (acc >> kU) & 1u)) << kTop);
      nlm >>= S;
      continue;
    }
    // accepted window cells: bit k = wy * S + wx <-> neighbour at (dx, dy) = (NC - wx, NC - wy)
#pragma unroll
    for (int dy = 0; dy <= NC; dy++) {
#pragma unroll
      for (int dx = 0; dx <= NC; dx++) {
        if (dx == 0 && dy == 0) continue;
        const float v = dx == 0 ? zr[dy] : __shfl_sync(kFull, zr[dy], (lane - dx) & 31);
        if (depth_gate(zc, v, thr)) acc |= 1u << ((NC - dy) * S + (NC - dx));
      }
    }
    acc = d ? (acc & dm & kCells) : 0u;  // both endpoints dynamic (lanes < dx read a wrapped column: dm is zero there)
    lm = (lm >> S) | (field(__ballot_sync(kFull, (acc >> kL) & 1u)) << kTop);
    um = (um >> S) | (field(__ballot_sync(kFull, (acc >> kU) & 1u)) << kTop);

    const int jc = j - NC;  // core row
    const bool act = d && lane >= NC;
    const uint32_t inside = jc < NC ? (col_inside & ~((1u << ((NC - jc) * S)) - 1u)) : col_inside;
    const uint32_t local = act ? (acc & inside) : 0u;
    const uint32_t upb = __ballot_sync(kFull, act && acc != 0u);
    if (lane == jc) upbits = upb;

    // cross-tile edges after the implied-edge elimination (window_edges' rule on bit masks, see cl_tile_fast_kernel)
    if (__any_sync(kFull, act && (acc & ~inside) != 0u)) {
      uint32_t keep = acc & ~(((acc & lm) >> 1) & col_ok) & ~((acc & um) >> S);
      // the vertical unit edge is dropped when the 4-cycle p, p-(1,0), up(p)-(1,0), up(p) is closed by accepted unit
      // edges: its endpoints are then connected through the up-edge one pixel to the left (induction on x)
      if (((acc >> kL) & 1u) && ((um >> kL) & 1u) && ((lm >> kU) & 1u)) keep &= ~(1u << kU);
      const uint32_t cross = act ? (keep & ~inside) : 0u;
      const uint32_t bb = __ballot_sync(kFull, cross != 0u);
      if (cross != 0u) {
        const int pos = nbord + __popc(bb & lt_mask);
        s_bpix[pos] = static_cast<uint16_t>(j * 32 + lane);
        s_bmask[pos] = cross;
      }
      nbord += __popc(bb);
    }

    // labels
    const int p = j * 32 + lane;
    const uint32_t newb = __ballot_sync(kFull, act && local == 0u);
    nlm = (nlm >> S) | (field(newb) << kTop);
    const bool bad = act && ((local != 0u && (local & kUpCells) == 0u) || (local & nlm) != 0u);
    if (L0 >= 0 && !__any_sync(kFull, bad)) {
      if (act) s_lab[p] = local ? L0 : p;
      nnew += __popc(newb);
      continue;
    }
    // slow path.  A: link to the first accepted neighbour in the rows above (those rows are flat: its label is a root)
    int mine = p;
    {
      const uint32_t lnk = local & kUpCells;
      if (lnk) mine = s_lab[p + cell_offset(__ffs(lnk) - 1)];
      if (act) s_lab[p] = mine;
      nnew += __popc(__ballot_sync(kFull, act && lnk == 0u));
    }
    __syncwarp();
    // B: the other accepted in-tile neighbours whose label differs
    uint32_t diff = 0u;
    if (local) {
#pragma unroll
      for (int k = 0; k < kA; k++)
        if (s_lab[p + cell_offset(k)] != mine) diff |= 1u << k;
      diff &= local;
    }
    // C: unions
    bool merged = false;
    while (__any_sync(kFull, diff != 0u)) {
      if (diff) {
        const int k = __ffs(diff) - 1;
        diff &= diff - 1u;
        if (smem_union_counted(s_lab, p, p + cell_offset(k))) {
          nun++;
          merged = true;
        }
      }
      __syncwarp();
    }
    // D: the rows still inside the window become flat again
    if (__any_sync(kFull, merged)) {
#pragma unroll
      for (int dd = 0; dd <= NC; dd++) {
        const int jj = j - dd;
        if (jj < NC) break;
        const int pp = jj * 32 + lane;
        if ((staged_bits(jj) & kCoreLanes & (1u << lane)) != 0u) {
          const int r = smem_find_readonly(s_lab, pp);
          s_lab[pp] = r;
        }
      }
      __syncwarp();
      if (L0 >= 0) L0 = smem_find_readonly(s_lab, L0);
    }
    // E: the label mask relative to L0; a new L0 when no pixel of the window rows carries the old one
#pragma unroll 1
    for (int attempt = 0; attempt < 2; attempt++) {
      uint32_t present = 0u;
      nlm = 0u;
#pragma unroll
      for (int dd = NC; dd >= 0; dd--) {
        const int jj = j - dd;
        uint32_t nlb = 0u;
        if (jj >= NC) {
          const uint32_t bd = staged_bits(jj) & kCoreLanes;
          const int lab = ((bd >> lane) & 1u) ? s_lab[jj * 32 + lane] : L0;
          nlb = __ballot_sync(kFull, lab != L0);
          present |= bd & ~nlb;
        }
        nlm |= field(nlb) << ((NC - dd) * S);
      }
      if (L0 >= 0 && present != 0u) break;
      const uint32_t cur = __ballot_sync(kFull, act);
      if (cur == 0u || attempt == 1) break;
      L0 = __shfl_sync(kFull, s_lab[act ? p : 0], __ffs(cur) - 1);
    }
  }

  // ---- reservations: one atomic per list and tile --------------------------------------------------------------------
  const int nr = nnew - __reduce_add_sync(kFull, nun);
  int base = 0;
  if (lane == 0) base = atomicAdd(counters + 0, nd);
  if (lane == 1 && nbord) base = atomicAdd(counters + 1, nbord);
  if (lane == 2) base = atomicAdd(counters + 2, nr);
  if (lane == 3) atomicAdd(L.n_dynamic + frame, nd);
  int pos_l = __shfl_sync(kFull, base, 0);
  const int pos_b = __shfl_sync(kFull, base, 1);
  int pos_r = __shfl_sync(kFull, base, 2);

  auto global_index = [&](int sp) -> int {  // staged pixel -> frame-local raster index
    return (ys + (sp >> 5)) * L.w + xs + (sp & 31);
  };
  for (int e = lane; e < nbord; e += 32) {
    border_list[pos_b + e] = static_cast<uint32_t>(fb + global_index(s_bpix[e]));
    border_mask[pos_b + e] = s_bmask[e];
  }

  // ---- pass 2: publish -------------------------------------------------------------------------------------------------
  // per-lane accumulator of the partial statistics of one tile root (the rows of a column mostly share it); the warp
  // flushes (redux per distinct root, then atomics) when a lane meets a second root, and once at the end
  int a_root = -1, a_cnt = 0, a_key = 0x7fffffff;
  uint32_t a_mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, a_mx[3] = {0u, 0u, 0u};
  auto flush = [&]() {
    unsigned todo = __ballot_sync(kFull, a_cnt > 0);
    while (todo) {  // peel one tile root at a time; usually one per tile
      const int leader = __ffs(todo) - 1;
      const int lroot = __shfl_sync(kFull, a_root, leader);
      const bool in = a_cnt > 0 && a_root == lroot;
      const unsigned grp = __ballot_sync(kFull, in);
      if (in) {
        uint32_t mn[3], mx[3];
#pragma unroll
        for (int q = 0; q < 3; q++) {
          mn[q] = __reduce_min_sync(grp, a_mn[q]);
          mx[q] = __reduce_max_sync(grp, a_mx[q]);
        }
        const int kk = __reduce_min_sync(grp, a_key);
        const int cnt = __reduce_add_sync(grp, a_cnt);
        if (lane == leader) {
          const size_t gr = fb + global_index(lroot);
          atomicAdd(L.count + gr, cnt);
          TileStat& st = L.tstat[gr];
#pragma unroll
          for (int q = 0; q < 3; q++) {
            if (mn[q] != 0xffffffffu) atomicMin(&st.mn[q], mn[q]);
            if (mx[q] != 0u) atomicMax(&st.mx[q], mx[q]);
          }
          if (kk != 0x7fffffff) atomicMin(&st.key, kk);
        }
      }
      todo &= ~grp;
    }
    a_cnt = 0;
    a_root = -1;
    a_key = 0x7fffffff;
#pragma unroll
    for (int q = 0; q < 3; q++) {
      a_mn[q] = 0xffffffffu;
      a_mx[q] = 0u;
    }
  };
#pragma unroll 1
  for (int jc = 0; jc < TH; jc++) {
    const uint32_t bits = __shfl_sync(kFull, rbc, jc) & kCoreLanes;
    if (bits == 0u) continue;
    const uint32_t upb = __shfl_sync(kFull, upbits, jc);
    const bool d = (bits >> lane) & 1u;
    const int p = (jc + NC) * 32 + lane;
    int root = -1;
    if (d) {  // parents lie earlier in raster order: rows above are already flat, a chain inside the row is short
      root = s_lab[p];
      for (;;) {
        const int up = s_lab[root];
        if (up == root) break;
        root = up;
      }
      s_lab[p] = root;
    }
    const int i = (y0 + jc) * L.w + xs + lane;  // frame-local raster index
    const uint32_t rball = __ballot_sync(kFull, d && root == p);
    if (rball) {  // tile roots initialise their partial record before any lane adds to it
      if (d && root == p) {
        L.count[fb + i] = 0;
        TileStat init;
        init.key = 0x7fffffff;
        init.mn[0] = init.mn[1] = init.mn[2] = 0xffffffffu;
        init.mx[0] = init.mx[1] = init.mx[2] = 0u;
        init.pad = 0;
        L.tstat[fb + i] = init;
        root_list[pos_r + __popc(rball & lt_mask)] = static_cast<uint32_t>(fb + i);
      }
      pos_r += __popc(rball);
      __syncwarp();
    }
    if (__any_sync(kFull, d && a_cnt > 0 && a_root != root)) flush();
    if (d) {
      L.parent[fb + i] = global_index(root);
      const int e = pos_l + __popc(bits & lt_mask);
      list[e] = static_cast<uint32_t>(fb + i);
      const float4 xyz = __ldg(reinterpret_cast<const float4*>(pts + static_cast<size_t>(i) * 8));
      // |v| (Vector3f::norm, the oracle's summation order) rides along with the list entry (cl_scatter reads it)
      const float4 vel = __ldg(reinterpret_cast<const float4*>(pts + static_cast<size_t>(i) * 8 + 4));
      list_speed[e] = __float_as_uint(__fsqrt_rn(
          __fadd_rn(__fadd_rn(__fmul_rn(vel.x, vel.x), __fmul_rn(vel.y, vel.y)), __fmul_rn(vel.z, vel.z))));
      a_root = root;
      a_cnt++;
      const float cc[3] = {xyz.x, xyz.y, xyz.z};
#pragma unroll
      for (int q = 0; q < 3; q++) {  // NaN coordinates are skipped (order dependent in the reference; unpinned)
        if (!isnan(cc[q])) {
          const uint32_t o = float_to_ordered(cc[q]);
          a_mn[q] = min(a_mn[q], o);
          a_mx[q] = max(a_mx[q], o);
        }
      }
      if ((upb >> lane) & 1u) a_key = min(a_key, i);
    }
    pos_l += __popc(bits);
  }
  flush();
}
