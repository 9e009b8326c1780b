// cluster_tile_dense.cuh -- cl_pack + cl_tile_dense: the tile-local part of the clusterer for neighbor_distance 1..4,
// register-resident.  Included by cluster.cu inside its anonymous namespace (uses its helpers).
//
// Same contract as cl_tile_fast_kernel (same outputs, same lists, same partition of the tile's dynamic pixels into
// connected components, see cluster.cu), different machine mapping.  cl_tile_fast walked per-tile LISTS of dynamic
// pixels through shared memory in a chain of ~14 CTA-wide phases and was issue- and barrier-bound (ncu, r01: 303 M warp
// instructions per 63 frames of 1280x720, barrier stall 3.2 per issue).  Non-empty tiles are dense (68 % of their pixels
// are dynamic on the benchmark sequence), so this kernel drops the lists:
//
//   * the dynamic map is first packed to ONE BIT per pixel (cl_pack_kernel); a CTA owns a macro-tile of 128 x 32 pixels
//     = four 32 x 32 tiles side by side, whose bits are 32 rows x 16 bytes: every warp fetches them with one 128-bit load
//     per lane and an empty macro-tile retires after that load and a vote, without touching shared memory or a barrier;
//   * inside a tile a warp owns eight rows and a lane owns one COLUMN of them: z of the strip and its N-row / N-column
//     halo lives in registers, a window cell (dx, dy) is the value of row -dy shifted by dx lanes (one warp shuffle per
//     staged row and dx, reused by up to N+1 pixel rows), and the accepted-cell masks, the link to the raster-smallest
//     accepted neighbour and the later label comparisons are all straight-line SIMD code with every lane busy;
//   * labels sit in shared memory (one int per pixel); the forest of links is flattened by pointer jumping, trees are
//     merged by one union per differing (root, root) pair, and only tiles that still show a differing edge take the
//     second round -- three CTA barriers for a tile that is one component, which is the usual case inside a blob;
//   * per-tile-root statistics are combined per lane over the strip's rows first and reduced once per warp (redux),
//     instead of once per 32 list entries.
//
// Correctness argument (unchanged): every accepted in-tile edge is either the link, or found joining equal labels, or
// unioned; every accepted edge that leaves the tile is handed to cl_merge unless the implied-edge rule of window_edges
// proves its endpoints connected through a shorter accepted edge and a unit edge.

// ---- cl_pack ------------------------------------------------------------------------------------------------------
// One thread per 32-pixel word of the bit plane (grid: x over the words of a frame, y over frames).  The word's 32 mask
// bytes are fetched as nine aligned 32-bit words and funnel-shifted into place (rows of a frame whose width is not a
// multiple of four start at any byte offset).  Bits of pixels beyond the image width and the pitch padding are zero.
__global__ void __launch_bounds__(kBlock)
    cl_pack_kernel(const uint8_t* __restrict__ mask, uint32_t* __restrict__ bits, int w, int h, int pitch) {
  const uint32_t words = static_cast<uint32_t>(h) * pitch;  // per frame
  const uint32_t g = blockIdx.x * kBlock + threadIdx.x;
  if (g >= words) return;
  const uint32_t y = g / pitch, wx = g - y * pitch;
  const int x = static_cast<int>(wx) * 32;
  uint32_t out = 0;
  if (x < w) {
    const uint8_t* p = mask + static_cast<size_t>(blockIdx.y) * w * h + y * static_cast<uint32_t>(w) + x;
    const int n = min(32, w - x);  // pixels of this word inside the row
    const uint32_t sh = (reinterpret_cast<uintptr_t>(p) & 3u) * 8u;
    const uint32_t* al = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(p) & ~static_cast<uintptr_t>(3));
    const int n_words = (n + static_cast<int>(sh / 8u) + 3) / 4;  // aligned words that hold the n bytes
    uint32_t v[9];
#pragma unroll
    for (int q = 0; q < 9; q++) v[q] = q < n_words ? __ldg(al + q) : 0u;
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const uint32_t four = __funnelshift_r(v[q], v[q + 1], sh);  // bytes 4q .. 4q+3 of the word's 32 pixels
      // bit 7 of every non-zero byte, then the four flags gathered into one nibble by a multiplication
      const uint32_t nz = (four | ((four & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u;
      out |= (((nz >> 7) * 0x10204080u) >> 28) << (4 * q);
    }
    if (n < 32) out &= (1u << n) - 1u;
  }
  bits[static_cast<size_t>(blockIdx.y) * words + g] = out;
}

// ---- cl_tile_dense --------------------------------------------------------------------------------------------------
#ifndef CL_DENSE_MIN_BLOCKS
#define CL_DENSE_MIN_BLOCKS 5
#endif
constexpr int kDenseThreads = 128;  // four warps, eight tile rows each
constexpr int kDenseRows = 8;       // rows of a tile one warp owns

template <int NC>
__global__ void __launch_bounds__(kDenseThreads, CL_DENSE_MIN_BLOCKS)
    cl_tile_dense_kernel(const ClLaunch L, const uint32_t* __restrict__ bits, int pitch, uint32_t* __restrict__ list,
                         uint32_t* __restrict__ list_speed, uint32_t* __restrict__ border_list,
                         uint32_t* __restrict__ border_mask, uint32_t* __restrict__ root_list,
                         int* __restrict__ counters /* list, border, root */) {
  static_assert(NC >= 1 && NC <= 4, "register-resident path covers neighbor_distance 1..4");
  constexpr int S = NC + 1;
  constexpr int R = kDenseRows, SR = R + NC;  // core rows and staged rows (core + top halo) of a warp's strip
  constexpr int kA = NC * S + NC, kL = kA - 1, kU = kA - S;
  constexpr uint32_t kField = (1u << S) - 1u;
  constexpr int kPixels = kTileW * kTileH;
  __shared__ int s_lab[kPixels];  // label of core pixel c = ty * 32 + tx: a pixel of the same tile, eventually the root
  __shared__ uint32_t s_aux[kPixels];  // per pixel, private to its thread: differing-edge mask (D), then cross-edge mask (F)
  __shared__ int s_cnt[3][4];     // per warp: dynamic pixels, border-list entries, tile roots
  __shared__ int s_base[3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int frame = blockIdx.z;
  const int my0 = blockIdx.y * kTileH;
  const size_t P = static_cast<size_t>(L.w) * L.h;
  const size_t fb = frame * P;
  const float* pts = L.points + frame * L.points_frame_stride * 8;
  const uint32_t* fbits = bits + static_cast<size_t>(frame) * L.h * pitch + blockIdx.x * 4;
  const float thr = L.depth_thr;
  const uint32_t lt_mask = (1u << lane) - 1u;

  // the macro-tile's 32 rows of dynamic bits, row my0 + lane in lane `lane` (every warp loads the same 512 bytes)
  uint4 wrow = make_uint4(0u, 0u, 0u, 0u);
  if (my0 + lane < L.h) wrow = __ldg(reinterpret_cast<const uint4*>(fbits + static_cast<size_t>(my0 + lane) * pitch));
  if (!__any_sync(kFull, (wrow.x | wrow.y | wrow.z | wrow.w) != 0u)) return;  // no dynamic pixel: all four warps agree

  // this warp's staged rows (strip + NC rows above it), row ys - NC + lane in lane `lane` < SR, with the word to the left
  const int ys = my0 + R * warp;  // first core row of the strip
  uint4 srow = make_uint4(0u, 0u, 0u, 0u);
  uint32_t sleft = 0u;
  {
    const int y = ys - NC + lane;
    if (lane < SR && y >= 0 && y < L.h) {
      const uint32_t* rowp = fbits + static_cast<size_t>(y) * pitch;
      srow = __ldg(reinterpret_cast<const uint4*>(rowp));
      if (blockIdx.x > 0) sleft = __ldg(rowp - 1);
    }
  }

  bool first_tile = true;
#pragma unroll 1
  for (int k = 0; k < 4; k++) {
    const uint32_t wk = k == 0 ? wrow.x : (k == 1 ? wrow.y : (k == 2 ? wrow.z : wrow.w));
    if (!__any_sync(kFull, wk != 0u)) continue;  // empty tile (CTA-uniform: every warp holds the same words)
    if (!first_tile) __syncthreads();            // the previous tile's readers of s_lab / s_cnt are done
    first_tile = false;
    const int tx0 = blockIdx.x * 128 + k * kTileW;
    const uint32_t sk = k == 0 ? srow.x : (k == 1 ? srow.y : (k == 2 ? srow.z : srow.w));
    const uint32_t sl = k == 0 ? sleft : (k == 1 ? srow.x : (k == 2 ? srow.y : srow.z));

    // ---- A: dynamic bits of the staged rows; z of the strip (lane = column) and of the NC halo columns --------------
    // Bc[j]: bits of the tile's 32 columns in staged row j, Bl[j]: of the 32 columns to its left.  Lane l >= 32 - NC
    // additionally holds halo column tx0 - 32 + l (bit l of Bl), so that "column x - dx" is lane (x - dx) & 31 of the
    // core value for x >= dx and of the halo value for x < dx.
    uint32_t fld[SR];  // per lane: dynamic flags of columns x - NC .. x of staged row j (bit NC - dx)
    float zc[SR], zh[SR];
    int nd_warp = 0;   // dynamic pixels of the strip (warp-uniform)
    const bool halo_lane = lane >= 32 - NC;
#pragma unroll
    for (int j = 0; j < SR; j++) {
      const uint32_t bc = __shfl_sync(kFull, sk, j), bl = __shfl_sync(kFull, sl, j);
      if (j >= NC) nd_warp += __popc(bc);
      const uint32_t lo = (bc << NC) | (bl >> (32 - NC)), hi = bc >> (32 - NC);
      fld[j] = __funnelshift_r(lo, hi, lane) & kField;
      const int y = ys - NC + j;
      const size_t rowi = static_cast<size_t>(y) * L.w + tx0;  // only dereferenced where a bit is set (y, x in range)
      zc[j] = ((bc >> lane) & 1u) ? __ldg(pts + (rowi + lane) * 8 + 2) : 0.0f;
      zh[j] = (halo_lane && ((bl >> lane) & 1u)) ? __ldg(pts + (rowi + lane - 32) * 8 + 2) : 0.0f;
    }

    // ---- B: accepted window cells of the strip's 8 x 32 pixels --------------------------------------------------------
    // bit k = wy * S + wx of acc[r]: the neighbour at (dx, dy) = (NC - wx, NC - wy) is dynamic and passes the depth gate
    uint32_t acc[R];
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] = 0u;
    uint32_t top_hl = 0u, top_vu = 0u;  // unit edges of the top-halo rows' core pixels: bit j = staged row j (< NC)
#pragma unroll
    for (int j = 0; j < SR; j++) {
#pragma unroll
      for (int dx = 0; dx <= NC; dx++) {
        float v = zc[j];
        if (dx > 0) {
          const float a = __shfl_sync(kFull, zc[j], (lane - dx) & 31);
          const float b = __shfl_sync(kFull, zh[j], (lane - dx) & 31);
          v = lane >= dx ? a : b;
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
          const int dy = r + NC - j;
          if (dy < 0 || dy > NC || (dx == 0 && dy == 0)) continue;
          if (depth_gate(zc[r + NC], v, thr)) acc[r] |= 1u << ((NC - dy) * S + (NC - dx));
        }
        if (dx == 1 && j < NC && depth_gate(zc[j], v, thr)) top_hl |= 1u << j;
      }
      if (j >= 1 && j < NC && depth_gate(zc[j], zc[j - 1], thr)) top_vu |= 1u << j;
    }
    // both endpoints dynamic: window flags of the neighbours, own flag of the pixel
    bool dyn[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      uint32_t dm = 0u;
#pragma unroll
      for (int wy = 0; wy <= NC; wy++) dm |= fld[r + wy] << (wy * S);  // window row wy is staged row r + NC - dy = r + wy
      dyn[r] = (dm >> kA) & 1u;
      acc[r] = dyn[r] ? (acc[r] & dm & ~(1u << kA)) : 0u;
    }

    // ---- C: link every pixel to its raster-smallest accepted in-tile neighbour ------------------------------------------
    // a cell (wx, wy) lies inside the tile iff tx - (NC - wx) >= 0 and ty - (NC - wy) >= 0
    uint32_t col_inside = 0u;
    {
      const int first = max(NC - lane, 0);  // wx >= first stays inside
      const uint32_t rowbits = first < S ? ((kField >> first) << first) : 0u;
#pragma unroll
      for (int wy = 0; wy < S; wy++) col_inside |= rowbits << (wy * S);
    }
    auto cell_offset = [](int c) -> int { return (c / S - NC) * kTileW + (c % S - NC); };
    uint32_t E[R];      // accepted in-tile cells minus the link
    uint32_t cross[R];  // accepted cells in other tiles
    int up_any = 0;     // bit r: pixel r has an accepted up-left neighbour at all (reference label order key)
    volatile int* vlab = s_lab;
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int ty = R * warp + r;
      uint32_t inside = col_inside;
      if (ty < NC) inside &= ~((1u << ((NC - ty) * S)) - 1u);  // window rows wy < NC - ty lie above the tile
      const uint32_t loc = acc[r] & inside;
      cross[r] = acc[r] & ~inside;
      up_any |= (acc[r] != 0u ? 1 : 0) << r;
      const int c = ty * kTileW + lane;
      if (dyn[r]) s_lab[c] = loc ? c + cell_offset(__ffs(loc) - 1) : c;  // lowest bit = raster-smallest neighbour
      E[r] = loc & (loc - 1u);
    }
    __syncthreads();  // (1) links of all four strips are in place

    // pointer jumping: every pass at least halves the distance to the root; roots do not change while it runs, so a
    // label that points at a root is final even though other warps are still jumping
    auto flatten = [&]() {
      bool again;
      do {
        again = false;
#pragma unroll
        for (int r = 0; r < R; r++) {
          if (dyn[r]) {
            const int c = (R * warp + r) * kTileW + lane;
            const int p = vlab[c];
            const int g = vlab[p];
            if (g != p) {
              vlab[c] = g;
              again = true;
            }
          }
        }
      } while (__any_sync(kFull, again));
    };
    flatten();
    __syncthreads();  // (2) labels are roots

    // ---- D: every other accepted in-tile edge must join equal labels ------------------------------------------------------
    // labels of the staged rows in the lane = column layout (rows above the strip come from the strips above; rows above
    // the tile are never looked at: their cells are not in E), neighbours by shuffle as in B
    int my[R];
    uint32_t diff[R];
    {
      int lr[SR];
#pragma unroll
      for (int j = 0; j < SR; j++) {
        const int ty = R * warp - NC + j;
        lr[j] = ty >= 0 ? s_lab[ty * kTileW + lane] : -1;  // non-dynamic pixels hold stale values; never selected by E
      }
#pragma unroll
      for (int r = 0; r < R; r++) {
        my[r] = lr[r + NC];
        diff[r] = 0u;
      }
#pragma unroll
      for (int j = 0; j < SR; j++) {
#pragma unroll
        for (int dx = 0; dx <= NC; dx++) {
          const int v = dx == 0 ? lr[j] : __shfl_sync(kFull, lr[j], (lane - dx) & 31);
#pragma unroll
          for (int r = 0; r < R; r++) {
            const int dy = r + NC - j;
            if (dy < 0 || dy > NC || (dx == 0 && dy == 0)) continue;
            if (v != my[r]) diff[r] |= 1u << ((NC - dy) * S + (NC - dx));
          }
        }
      }
    }
    bool any_diff = false;
#pragma unroll
    for (int r = 0; r < R; r++) {
      diff[r] &= E[r];
      any_diff = any_diff || diff[r] != 0u;
    }
    if (__syncthreads_or(any_diff)) {  // (3) the tile holds more than one tree so far
      // The rest of D is rare enough to run as rolled loops; the per-row masks move to shared memory for that (each
      // entry is written and read by its own thread only).
#pragma unroll
      for (int r = 0; r < R; r++) s_aux[(R * warp + r) * kTileW + lane] = diff[r];
      // one union per pixel (its first differing neighbour), one lane per distinct (root, root) pair of the warp
#pragma unroll 1
      for (int r = 0; r < R; r++) {
        const int c = (R * warp + r) * kTileW + lane;
        const uint32_t d = s_aux[c];
        int mine = -1, other = -1;
        if (d) {
          mine = vlab[c];
          other = vlab[c + cell_offset(__ffs(d) - 1)];
        }
        const unsigned has = __ballot_sync(kFull, other >= 0);
        if (other >= 0) {
          const int lo = min(mine, other), hi = max(mine, other);
          const unsigned same = __match_any_sync(has, (lo << 10) | hi);
          if (lane == __ffs(same) - 1) smem_union(s_lab, lo, hi);
        }
      }
      __syncthreads();  // (4)
      flatten();
      __syncthreads();  // (5)
      // whatever still differs among the edges that differed before (equal labels stay equal) is unioned edge by edge
      bool unioned = false;
#pragma unroll 1
      for (int r = 0; r < R; r++) {
        const int c = (R * warp + r) * kTileW + lane;
        uint32_t d = s_aux[c];
        const int mine = d ? vlab[c] : -1;
        while (__any_sync(kFull, d != 0u)) {
          if (d) {
            const int cell = __ffs(d) - 1;
            d &= d - 1u;
            const int nb = c + cell_offset(cell);
            if (vlab[nb] != mine) {
              smem_union(s_lab, c, nb);
              unioned = true;
            }
          }
          __syncwarp();
        }
      }
      if (__syncthreads_or(unioned)) {  // (6)
        flatten();
        __syncthreads();  // (7)
      }
#pragma unroll
      for (int r = 0; r < R; r++) my[r] = s_lab[(R * warp + r) * kTileW + lane];
    }

    // ---- E: cross-tile edges after the implied-edge elimination (window_edges' rule, see cl_tile_fast) -------------------
    uint32_t xmask[R];
    bool any_cross = false;
#pragma unroll
    for (int r = 0; r < R; r++) {
      xmask[r] = cross[r];
      any_cross = any_cross || cross[r] != 0u;
    }
    if (__any_sync(kFull, any_cross)) {
      // unit-edge flags of every staged pixel as bit rows: HL = edge to the left neighbour accepted, VU = to the upper
      // one; for core pixels of core rows they are bits kL / kU of acc, halo rows and halo columns are evaluated here
      uint32_t lm[R], um[R];
#pragma unroll
      for (int r = 0; r < R; r++) lm[r] = um[r] = 0u;
#pragma unroll
      for (int j = 0; j < SR; j++) {
        bool hl_c, vu_c;
        if (j >= NC) {
          hl_c = (acc[j - NC] >> kL) & 1u;
          vu_c = (acc[j - NC] >> kU) & 1u;
        } else {
          const bool own = (fld[j] >> NC) & 1u;
          hl_c = own && ((fld[j] >> (NC - 1)) & 1u) && ((top_hl >> j) & 1u);
          vu_c = j >= 1 && own && ((fld[j - 1] >> NC) & 1u) && ((top_vu >> j) & 1u);
        }
        // halo columns (lanes >= 32 - NC): dynamic bits of Bl are bits of the lanes' own fld only for x < NC, so they
        // are recovered from the z staging: a halo pixel was loaded iff it is dynamic -- keep explicit flags instead
        const uint32_t bl = __shfl_sync(kFull, sl, j);
        const bool h_own = halo_lane && ((bl >> lane) & 1u);
        const float zl = __shfl_sync(kFull, zh[j], (lane - 1) & 31);
        const bool hl_h = h_own && lane > 32 - NC && ((bl >> (lane - 1)) & 1u) && depth_gate(zh[j], zl, thr);
        bool vu_h = false;
        if (j >= 1) {
          const uint32_t bl_up = __shfl_sync(kFull, sl, j - 1);
          vu_h = h_own && ((bl_up >> lane) & 1u) && depth_gate(zh[j], zh[j - 1], thr);
        }
        const uint32_t hlc = __ballot_sync(kFull, hl_c), hlh = __ballot_sync(kFull, hl_h);
        const uint32_t vuc = __ballot_sync(kFull, vu_c), vuh = __ballot_sync(kFull, vu_h);
        const uint32_t hl_f = __funnelshift_r((hlc << NC) | (hlh >> (32 - NC)), hlc >> (32 - NC), lane) & kField;
        const uint32_t vu_f = __funnelshift_r((vuc << NC) | (vuh >> (32 - NC)), vuc >> (32 - NC), lane) & kField;
#pragma unroll
        for (int r = 0; r < R; r++) {
          const int wy = j - r;  // staged row j is window row wy of pixel row r
          if (wy < 0 || wy > NC) continue;
          lm[r] |= hl_f << (wy * S);
          um[r] |= vu_f << (wy * S);
        }
      }
      uint32_t col_ok = 0u;  // cells whose right neighbour is still inside the window (wx < NC)
#pragma unroll
      for (int wy = 0; wy < S; wy++) col_ok |= ((1u << NC) - 1u) << (wy * S);
#pragma unroll
      for (int r = 0; r < R; r++) {
        const uint32_t a = acc[r];
        uint32_t keep = a & ~(((a & lm[r]) >> 1) & col_ok) & ~((a & um[r]) >> S);
        // the vertical unit edge is dropped when the 4-cycle p, p-(1,0), up(p)-(1,0), up(p) is closed by accepted unit
        // edges: its endpoints are then connected through the up-edge one pixel to the left (induction on x)
        if (((a >> kL) & 1u) && ((um[r] >> kL) & 1u) && ((lm[r] >> kU) & 1u)) keep &= ~(1u << kU);
        xmask[r] = cross[r] & keep;
      }
    }

    // ---- F: publish -------------------------------------------------------------------------------------------------------
    // counts per warp -> one reservation per list and tile; roots initialise their partial record
    int nb_warp = 0, nr_warp = 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int ty = R * warp + r;
      const int c = ty * kTileW + lane;
      const bool is_root = dyn[r] && my[r] == c;
      nb_warp += __popc(__ballot_sync(kFull, xmask[r] != 0u));
      nr_warp += __popc(__ballot_sync(kFull, is_root));
      s_aux[c] = xmask[r];
      if (is_root) {
        const size_t gi = fb + static_cast<size_t>(my0 + ty) * L.w + tx0 + lane;
        L.count[gi] = 0;
        TileStat init;
        init.key = 0x7fffffff;
        init.mn[0] = init.mn[1] = init.mn[2] = 0xffffffffu;
        init.mx[0] = init.mx[1] = init.mx[2] = 0u;
        init.pad = 0;
        L.tstat[gi] = init;
      }
    }
    if (lane == 0) {
      s_cnt[0][warp] = nd_warp;
      s_cnt[1][warp] = nb_warp;
      s_cnt[2][warp] = nr_warp;
    }
    __syncthreads();  // (8) counts and root records visible
    if (threadIdx.x < 3) {
      const int t = threadIdx.x;
      const int n = s_cnt[t][0] + s_cnt[t][1] + s_cnt[t][2] + s_cnt[t][3];
      s_base[t] = n ? atomicAdd(counters + t, n) : 0;
      if (t == 0) atomicAdd(L.n_dynamic + frame, n);
    }
    __syncthreads();  // (9)
    int pos_l = s_base[0], pos_b = s_base[1], pos_r = s_base[2];
    for (int w2 = 0; w2 < warp; w2++) {
      pos_l += s_cnt[0][w2];
      pos_b += s_cnt[1][w2];
      pos_r += s_cnt[2][w2];
    }

    // per-lane accumulator of the partial statistics of one tile root (the rows of a column mostly share it); a warp
    // flushes (redux per distinct root, then atomics) when a lane meets a second root, and once at the end
    int a_root = -1, a_cnt = 0, a_key = 0x7fffffff;
    uint32_t a_mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, a_mx[3] = {0u, 0u, 0u};
    auto flush = [&]() {
      unsigned todo = __ballot_sync(kFull, a_cnt > 0);
      while (todo) {  // peel one tile root at a time; usually one or two per warp
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
            const size_t gr = fb + static_cast<size_t>(my0 + lroot / kTileW) * L.w + tx0 + lroot % kTileW;
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
    for (int r = 0; r < R; r++) {  // rolled: per-row values come from the bit words and shared memory, not register arrays
      const int ty = R * warp + r;
      const int c = ty * kTileW + lane;
      const int i = (my0 + ty) * L.w + tx0 + lane;  // frame-local raster index
      const uint32_t dball = __shfl_sync(kFull, sk, r + NC);
      const bool d = (dball >> lane) & 1u;
      const int root = d ? s_lab[c] : -1;
      const uint32_t xm = s_aux[c];
      const uint32_t bball = __ballot_sync(kFull, xm != 0u);
      const uint32_t rball = __ballot_sync(kFull, d && root == c);
      if (__any_sync(kFull, d && a_cnt > 0 && a_root != root)) flush();
      if (d) {
        L.parent[fb + i] = (my0 + root / kTileW) * L.w + tx0 + root % kTileW;
        const int e = pos_l + __popc(dball & lt_mask);
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
        if ((up_any >> r) & 1) a_key = min(a_key, i);
      }
      pos_l += __popc(dball);
      if (xm != 0u) {
        const int e = pos_b + __popc(bball & lt_mask);
        border_list[e] = static_cast<uint32_t>(fb + i);
        border_mask[e] = xm;
      }
      pos_b += __popc(bball);
      if (d && root == c) root_list[pos_r + __popc(rball & lt_mask)] = static_cast<uint32_t>(fb + i);
      pos_r += __popc(rball);
    }
    flush();
  }
}
