// psph_mix.cuh — pair search and pair sums of the 3-D SPH stages when every row carries its OWN smoothing length.
//
// That is the regime of every time step after the first adaptive_sml call (main.cc:319-322: step 50 of a default run):
// h_i = eta (m_i / rho*_i)^(1/3) with eta = 1.5, so the kernel support of a pair, cutoff * (h_i + h_j) / 2 for
// mom_engr_update / smooth_density and cutoff * h_i for smooth_velocity, reaches 4.5-5 subcells instead of 3 and a row has
// ~790 partners instead of ~250.  The block tiles of psph_tile.cuh (one standard h, reach 3, the whole neighbourhood of a
// block in shared memory) do not apply: the neighbourhood of a 128-row block is ~3900 rows of 120 bytes.  Here
//   block    : as before, the rows of one bucket with subcell x-index in {2q, 2q+1} (one contiguous row range);
//   tile     : the subcell columns gx = 8 cx + 2q - 5 .. + 6 (12), gy = 8 cy - 5 .. + 12 (18) over the three z-buckets
//              (levels -5 .. 12), cut into MNSUB = 3 SUB-TILES of 4 x-columns each (~1300 rows);
//   k_mix_scan   : one CTA per bucket finds the pairs of both gate families (symmetric-h gates of mom / dens, h_i gates of
//                  the velocity filter) with the FP32 pre-filter + FP64 band gates of psph_tile.cuh, and writes per-row
//                  lists of 16-bit entries  slot-in-sub-tile | family tags,  ordered by sub-tile, MSTRIDE entries apart;
//   k_mix_gather : per stage, one CTA per block walks the three sub-tiles in turn: stages the sub-tile's records in shared
//                  memory, every warp sums its rows' entries of that sub-tile (lanes share a row's entries as in
//                  k_tile_gather), reduces, and adds the result to the row's accumulators in shared memory; after the
//                  last sub-tile the rows are finished (S::mix_finish).  Staging is ~6 % of a block's time.
// mom_engr_update's pass forms the phase-density sums of smooth_density on the side and lists the partners inside the
// energy-smoothing kernel (V.near), exactly as the uniform-h tile gather does; smooth_density is then k_dens_near<false>.
// Rows that do not fit (reach beyond 5 subcells, a sub-tile over MCAP rows, a row on a subcell face, list or in_support
// limits that can bite — every row of the step in which adaptive_sml itself runs —, more than MSTRIDE partners) get no
// list and are summed by the fused grouped gather (k_gather_grp LM 2), like everywhere else.
#pragma once
#include "psph_scan2.cuh"

constexpr int MR = 5;                              // reach in subcells the layout is made for
constexpr int MTX = 2 + 2 * MR, MTY = SUBDIV + 2 * MR;   // block-tile columns: 12 x 18
constexpr int MSX = 4;                             // tile x-columns per sub-tile
constexpr int MNSUB = MTX / MSX;                   // 3 sub-tiles
constexpr int MSEG = MSX * MTY * 3;                // (column, z-bucket) row segments of a sub-tile: 216
constexpr int MCAP = 1536;                         // rows of a sub-tile held in shared memory
constexpr int MCX = SUBDIV + 2 * MR, MCY = MCX;    // columns of the bucket's tile (search): 18 x 18
constexpr int MCSEG = MCX * MCY * 3;               // 972 segments
constexpr int MCCAP = 6656;                        // rows of a bucket tile held in shared memory (FP32 records)
constexpr int MLV = SUBDIV + 2 * MR;               // z levels of a column: local z = -5 .. 12
constexpr int MCZL = MLV + 2;                      // ... + end of column (+ pad)
constexpr int MQ = 256;                            // queue entries per row of a group before they go to the global list
constexpr int MSTRIDE = 1024;                      // list entries per row (fixed stride: no allocator, appends need no offsets)
constexpr int MBR = 256;                           // rows of a block the gather holds accumulators for
constexpr unsigned MTAG_SYM = 0x8000u;             // entry passes the symmetric-h gates (mom_engr_update, smooth_density)
constexpr unsigned MTAG_HI = 0x4000u;              // entry passes the h_i gates (smooth_velocity)
constexpr unsigned MSLOT = 0x07ffu;                // slot within the sub-tile
static_assert(MCAP <= (int) MSLOT + 1 && MTX % MSX == 0, "sub-tile layout");

// z-range of segment k (bucket below / own / above) inside its bucket
__device__ __forceinline__ int mix_zlo(int k) { return k == 0 ? SUBDIV - MR : 0; }
__device__ __forceinline__ int mix_zhi(int k) { return k == 2 ? MR - 1 : SUBDIV - 1; }

// (beg, len, base) of the column (gx, gy) in bucket layer bz, z-range of segment k
__device__ __forceinline__ void mix_segment(const GridView &G, int gx, int gy, int bz, int k, int &beg, int &len, int &base)
{
  beg = 0; len = 0; base = -1;
  if (gx < 0 || gy < 0 || bz < 0 || gx >= G.ncell[0] * SUBDIV || gy >= G.ncell[1] * SUBDIV || bz >= G.ncell[2]) return;
  const int cell = ((gx / SUBDIV) * G.ncell[1] + gy / SUBDIV) * G.ncell[2] + bz;
  base = __ldg(&G.cell2rank[cell]) * SUBS + (gx % SUBDIV) * (SUBDIV * SUBDIV) + (gy % SUBDIV) * SUBDIV;
  beg = __ldg(&G.sub_start[base + mix_zlo(k)]);
  len = __ldg(&G.sub_start[base + mix_zhi(k) + 1]) - beg;
}

// =========================================================================================
// pair search
// =========================================================================================
struct MixCellTable
{
  float4 geo[MCCAP];                           // slot -> FP32 record {x - o, y - o, z - o, h}
  int beg[MCSEG], base[MCSEG], off[MCSEG + 4]; // segment tables; off[MCSEG] = tile rows
  unsigned short colz[MCX * MCY][MCZL];        // column -> slot of its first row with local z >= level - MR; [MLV] = end
  int wsum[32];
  int blk_ok[4];                               // block q: every sub-tile within MCAP rows, block within MBR rows
  double hloc;                                 // largest smoothing length in the 27 buckets around this one: a partner's h at most
};
struct MixWarp
{
  GrpHdr hdr;
  unsigned short cand[TCAND];                  // candidate slots (bucket tile)
  unsigned short q[GRP][MQ];                   // entries on their way to the global lists
};

// slot of the bucket tile -> row << 1 | own-bucket flag (rare path: FP64 band gates)
__device__ __noinline__ int mix_slot_row(const MixCellTable &T, int sl)
{
  int lo = 0, hi = MCSEG;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (T.off[mid] <= sl) lo = mid; else hi = mid; }
  const int tx = lo / (MCY * 3), ty = (lo / 3) % MCY, k = lo % 3;
  const int own = (k == 1 && tx >= MR && tx < MR + SUBDIV && ty >= MR && ty < MR + SUBDIV) ? 1 : 0;
  return ((T.beg[lo] + (sl - T.off[lo])) << 1) | own;
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
k_mix_scan(DevView V, GridView G, KParams K, StepState *ss, unsigned char *__restrict__ blk_any)
{
  static_assert(WARPS * 64 >= MCSEG, "two segments per thread");
  extern __shared__ __align__(16) unsigned char s_raw[];
  MixCellTable &T = *reinterpret_cast<MixCellTable *>(s_raw);
  MixWarp *scr = reinterpret_cast<MixWarp *>(s_raw + sizeof(MixCellTable)) + (threadIdx.x >> 5);
  __shared__ int s_any[4];
  if (ss->pair_mode != 4) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  unsigned s_cand, s_q, s_hdr;
  asm volatile("mov.u32 %0, %1;" : "=r"(s_cand) : "r"((unsigned) __cvta_generic_to_shared(&scr->cand[0])));
  asm volatile("mov.u32 %0, %1;" : "=r"(s_q) : "r"((unsigned) __cvta_generic_to_shared(&scr->q[0][0])));
  asm volatile("mov.u32 %0, %1;" : "=r"(s_hdr) : "r"((unsigned) __cvta_generic_to_shared(&scr->hdr.hdr_i[0])));
  if (lane == 0) {
    scr->hdr.kc[0] = K.cutoff; scr->hdr.kc[1] = K.cutoff2; scr->hdr.kc[2] = K.neigh_radius;
    scr->hdr.geo = V.geo; scr->hdr.hlist = V.t.d[D_HLIST];
  }
  __syncwarp();
  const int list_rows = ss->mix_rows;
  // FP32 decision band (see psph_gather3.cuh): +-(2^-19 bound + 2^-21 xmax) around a cutoff face
  const float delta = __uint_as_float(ss->xmax_bits) * 0x1p-21f;
  const float cs = (float) (0.5 * K.cutoff), cv = (float) K.cutoff;
  const float cs_lo = cs * (1.0f - 0x1p-19f), cs_hi = cs * (1.0f + 0x1p-19f);
  const float cv_lo = cv * (1.0f - 0x1p-19f), cv_hi = cv * (1.0f + 0x1p-19f);
  unsigned short *__restrict__ lists = V.mlist;
  const int ncell = (int) G.ncells;

  for (int r = blockIdx.x; r < ncell; r += gridDim.x) {
    const int rb0 = __ldg(&G.sub_start[r * SUBS]), rb1 = __ldg(&G.sub_start[r * SUBS + SUBS]);
    if (rb0 == rb1) { if (threadIdx.x < 4) blk_any[r * 4 + threadIdx.x] = 0; continue; }
    const int cell0 = __ldg(&G.rank2cell[r]);
    const int cz = cell0 % G.ncell[2], cy = (cell0 / G.ncell[2]) % G.ncell[1], cx = cell0 / (G.ncell[2] * G.ncell[1]);
    if (threadIdx.x < 4) s_any[threadIdx.x] = 0;
    // ---- the bucket's tile: segment table (exclusive scan of the segment lengths, two segments per thread)
    {
      const int s0 = 2 * threadIdx.x;
      int beg[2] = { 0, 0 }, len[2] = { 0, 0 }, base[2] = { -1, -1 };
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const int s = s0 + u;
        if (s < MCSEG) {
          const int tx = s / (MCY * 3), ty = (s / 3) % MCY, k = s % 3;
          mix_segment(G, cx * SUBDIV - MR + tx, cy * SUBDIV - MR + ty, cz - 1 + k, k, beg[u], len[u], base[u]);
        }
      }
      int inc = len[0] + len[1];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += v; }
      if (lane == 31) T.wsum[warp] = inc;
      __syncthreads();
      int pre = 0;
      for (int k = 0; k < warp; k++) pre += T.wsum[k];
      const int o0 = pre + inc - len[0] - len[1];
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const int s = s0 + u;
        if (s < MCSEG) { T.beg[s] = beg[u]; T.base[s] = base[u]; T.off[s] = o0 + (u ? len[0] : 0); }
      }
      if (s0 + 1 == MCSEG - 1) T.off[MCSEG] = pre + inc;
      else if (s0 == MCSEG - 1) T.off[MCSEG] = pre + inc;
      __syncthreads();
    }
    const int ctotal = T.off[MCSEG];
    const bool cell_ok = ctotal <= MCCAP;
    if (warp == 1) {
      // every partner of a row of this bucket sits in one of the 27 buckets around it (neighs.cc:95-140): their largest h
      double hl = 0.0;
      if (lane < 27) {
        const int ni = cx + lane / 9 - 1, nj = cy + (lane / 3) % 3 - 1, nk = cz + lane % 3 - 1;
        if (ni >= 0 && nj >= 0 && nk >= 0 && ni < G.ncell[0] && nj < G.ncell[1] && nk < G.ncell[2])
          hl = __longlong_as_double((long long) G.cell_hmax[(ni * G.ncell[1] + nj) * G.ncell[2] + nk]);
      }
      for (int o = 16; o; o >>= 1) hl = fmax(hl, __shfl_xor_sync(FULL, hl, o));
      if (lane == 0) T.hloc = hl;
    }
    if (threadIdx.x < 4) {
      // block q: tile x-slabs 2q .. 2q + 11, sub-tile s = slabs 2q + 4s .. + 3; rows of the block itself
      const int q = threadIdx.x;
      bool ok = cell_ok;
      for (int s = 0; s < MNSUB; s++)
        ok = ok && T.off[(2 * q + MSX * s + MSX) * MCY * 3] - T.off[(2 * q + MSX * s) * MCY * 3] <= MCAP;
      const int b0 = __ldg(&G.sub_start[r * SUBS + q * 128]), b1 = __ldg(&G.sub_start[r * SUBS + q * 128 + 128]);
      T.blk_ok[q] = ok && (b1 - b0) <= MBR;
    }
    if (cell_ok) {
      for (int idx = threadIdx.x; idx < MCX * MCY * (MLV + 1); idx += WARPS * 32) {
        const int col = idx / (MLV + 1), lvl = idx - col * (MLV + 1);        // local z = lvl - MR
        int slot;
        if (lvl >= MLV) slot = T.off[col * 3 + 3];                            // end of the column
        else {
          const int k = lvl < MR ? 0 : (lvl < MR + SUBDIV ? 1 : 2), s = col * 3 + k;
          const int sz = k == 0 ? SUBDIV - MR + lvl : (k == 1 ? lvl - MR : lvl - MR - SUBDIV);  // subcell z inside that bucket
          if (T.base[s] < 0) slot = T.off[s];
          else slot = T.off[s] + (__ldg(&G.sub_start[T.base[s] + sz]) - T.beg[s]);
        }
        T.colz[col][lvl] = (unsigned short) slot;
      }
      // FP32 records of every tile row: one thread per segment run
      for (int s = threadIdx.x; s < MCSEG; s += WARPS * 32) {
        const int beg = T.beg[s], o = T.off[s], len = T.off[s + 1] - o;
        for (int j = 0; j < len; j++) T.geo[o + j] = V.geo32[beg + j];
      }
    }
    __syncthreads();
    const int per = (rb1 - rb0 + WARPS - 1) / WARPS;
    int e = rb0 + warp * per;
    const int eend = min(e + per, rb1);
    while (e < eend) {
      // ---- form the group: leading run of rows that are summed (F_NEED) and sit in the same subcell column
      int row = -1, fl = 0;
      unsigned long long pk = 0;
      if (lane < GRP && e + lane < eend) { row = e + lane; fl = V.flags[row]; pk = V.gsub[row]; }
      const bool need = row >= 0 && (fl & F_NEED);
      const unsigned nm = __ballot_sync(FULL, need);
      if (!(nm & 1u)) { e += nm ? __ffs(nm) - 1 : GRP; continue; }     // skip rows that are not summed
      const unsigned long long pk0 = __shfl_sync(FULL, pk, 0);
      const bool same = need && ((pk ^ pk0) & 0xffffffffffull) == 0;
      const int m = __ffs(~__ballot_sync(FULL, same)) - 1;             // 1..GRP
      const bool mine = lane < m;
      const int lx = (int) (pk0 & 0xfffff) - cx * SUBDIV, ly = (int) ((pk0 >> 20) & 0xfffff) - cy * SUBDIV;
      const int q = lx >> 1;                                           // the block of the group
      float4 p32 = make_float4(0.f, 0.f, 0.f, 0.f);
      int sz = 0, myself = -1, rsub = 0;
      bool bad = false;
      if (mine) {
        sz = (int) ((pk >> 40) & 0xfffff) - cz * SUBDIV;
        sts32(s_hdr + 4u * lane, row);
        const double hi = V.t.d[D_H][row], rmax = K.cutoff * 0.5 * (hi + T.hloc);   // both gate families reach at most this far
        // the neighbour-list / in_support limits (neighs.cc:84-92,130; sph_header.h:49-60) can bite: not a list row
        const bool nl = rmax > fmin(K.cutoff2 * hi, K.neigh_radius * V.t.d[D_HLIST][row]);
        // ... nor is one whose partners can sit beyond MR subcells, or that lies on a subcell face
        rsub = (int) ceil(rmax * G.inv_subw - 1e-9);          // subcells a partner can be away (rows on a subcell face are out)
        bad = nl || rsub > MR || (pk >> 60) != 0 || row >= list_rows;
        if (cell_ok) {
          const int s = ((lx + MR) * MCY + ly + MR) * 3 + 1;
          myself = T.off[s] + (row - T.beg[s]);
          p32 = T.geo[myself];
        }
      }
      const bool fits = cell_ok && T.blk_ok[q] && !__any_sync(FULL, bad);
      if (!fits) {
        if (mine) { V.nbr_cnt[row] = -1; V.nbr_off[row] = -1; }
        if (lane == 0) atomicAdd(&ss->n_uncached, m);
        e += m;
        continue;
      }
      const int zlo = __reduce_min_sync(FULL, mine ? sz : 0x7fffffff), zhi = __reduce_max_sync(FULL, mine ? sz : -1);
      // the candidate box is as wide as THIS group's smoothing lengths need (h = 1.2 spacings: 4 subcells, not MR = 5:
      // 9 x 9 instead of 11 x 11 columns)
      const int R = min(max(__reduce_max_sync(FULL, mine ? rsub : 0), 1), MR);
      unsigned ovf = 0;
      int ig[GRP], selfslot[GRP], cnt[GRP], c0s[GRP], c1s[GRP];
      unsigned qa[GRP];
      float fx[GRP], fy[GRP], fz[GRP], blo[GRP], bhi[GRP], vlo[GRP], vhi[GRP];
#pragma unroll
      for (int g = 0; g < GRP; g++) {
        ig[g] = __shfl_sync(FULL, row, g);
        selfslot[g] = __shfl_sync(FULL, myself, g);
        fx[g] = __shfl_sync(FULL, p32.x, g); fy[g] = __shfl_sync(FULL, p32.y, g); fz[g] = __shfl_sync(FULL, p32.z, g);
        const float hf = __shfl_sync(FULL, p32.w, g);
        blo[g] = hf * cs_lo - delta; bhi[g] = hf * cs_hi + delta;
        vlo[g] = hf * cv_lo - delta; vhi[g] = hf * cv_hi + delta;
        qa[g] = s_q + 2u * MQ * g;
        cnt[g] = 0; c0s[g] = 0; c1s[g] = 0;
      }
      // entries of queue g -> the row's global list (appended; the list is MSTRIDE entries apart)
      auto flush = [&](int g, unsigned qend) {
        const unsigned qb = s_q + 2u * MQ * g;
        const int n = (int) (qend - qb) >> 1;
        __syncwarp();
        if (cnt[g] + n > MSTRIDE) ovf |= 1u << g;
        else {
          unsigned short *__restrict__ dst = lists + (size_t) ig[g] * MSTRIDE + cnt[g];
          for (int k = lane; k < n; k += 32) dst[k] = (unsigned short) lds16(qb + 2u * k);
          cnt[g] += n;
        }
        __syncwarp();
      };
      // the group's box in bucket-tile columns (the row's own column is (lx + MR, ly + MR), its level sz + MR):
      // x [lx + MR - R, lx + MR + R], y likewise; z levels [zlo + MR - R, zhi + MR + R]
      const int lv0 = zlo + MR - R, lv1 = zhi + MR + R + 1;
      for (int s = 0; s < MNSUB; s++) {
        const int xs0 = 2 * q + MSX * s;
        const int xa = max(lx + MR - R, xs0), xb = min(lx + MR + R, xs0 + MSX - 1);
        const int subbase = T.off[xs0 * MCY * 3];                      // first slot of the sub-tile
        const int ncy = 2 * R + 1, ncol = max(xb - xa + 1, 0) * ncy;
        int col_base = -32, c0 = 0, total = 0, slot = 0, len = 0;

        auto test = [&](unsigned sl, const float4 &p) {
          const float hs_lo = p.w * cs_lo, hs_hi = p.w * cs_hi;
          unsigned okb = 0, okv = 0, bandb = 0;
#pragma unroll
          for (int g = 0; g < GRP; g++) {
            if (g < m) {
              const float m3 = fmaxf(fmaxf(fabsf(fx[g] - p.x), fabsf(fy[g] - p.y)), fabsf(fz[g] - p.z));
              const bool notself = (int) sl != selfslot[g];
              const bool os = m3 <= bhi[g] + hs_hi && notself, ov = m3 <= vhi[g] && notself;
              if (os) okb |= 1u << g;
              if (ov) okv |= 1u << g;
              if ((os && m3 > blo[g] + hs_lo) || (ov && m3 > vlo[g])) bandb |= 1u << g;
            }
          }
          if (__any_sync(FULL, bandb != 0)) {   // within the FP32 error band of a cutoff face: the reference's own tests
            if (bandb) {
              const int cj = mix_slot_row(T, (int) sl);
              okv = (okv & ~bandb) | gate_exact_grp(&scr->hdr, bandb, bandb, cj, 0);
              okb = (okb & ~bandb) | gate_exact_grp(&scr->hdr, bandb, bandb, cj, 1);
            }
          }
#pragma unroll
          for (int g = 0; g < GRP; g++) {
            if (g < m) {
              const unsigned sb = (okb >> g) & 1u, vb = (okv >> g) & 1u;
              const bool ok = (sb | vb) != 0;
              const unsigned bal = __ballot_sync(FULL, ok);
              if (ok) sts16(qa[g] + 2u * __popc(bal & lt), (sl - (unsigned) subbase) | (sb ? MTAG_SYM : 0u) | (vb ? MTAG_HI : 0u));
              qa[g] += 2u * __popc(bal);
            }
          }
        };

        for (;;) {
          if (c0 >= total) {
            // refill: the slot ranges of the next 32 columns of the box, expanded into the candidate list
            if (total) { total = 0; c0 = 0; __syncwarp(); }
            for (;;) {
              if (__any_sync(FULL, len > 0)) {
                const int quota = (TCAND - total) / 32;
                if (quota == 0) break;
                const int take = min(len, quota);
                int off = take;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, off, o); if (lane >= o) off += v; }
                const int sum = __shfl_sync(FULL, off, 31);
                off += total - take;
                for (int k = 0; k < take; k++) sts16(s_cand + 2u * (off + k), (unsigned) (slot + k));
                slot += take; len -= take; total += sum;
                continue;
              }
              col_base += 32;
              if (col_base >= ncol) break;
              const int c = col_base + lane;
              len = 0;
              if (c < ncol) {
                const int ix = c / ncy, iy = c - ix * ncy, col = (xa + ix) * MCY + ly + MR - R + iy;
                slot = T.colz[col][lv0];
                len = (int) T.colz[col][lv1] - slot;
              }
            }
            __syncwarp();
            if (total == 0) break;
          }
#pragma unroll
          for (int g = 0; g < GRP; g++)
            if (g < m && qa[g] > s_q + 2u * (MQ * g + MQ - 64)) { flush(g, qa[g]); qa[g] = s_q + 2u * MQ * g; }
          // two candidates per lane
          const int ca = c0 + lane, cb = ca + 32;
          const bool la = ca < total, lb = cb < total;
          const unsigned sa = la ? lds16(s_cand + 2u * ca) : 0u, sb = lb ? lds16(s_cand + 2u * cb) : 0u;
          float4 pa = T.geo[sa], pb = T.geo[sb];
          if (!la) pa.x = INFINITY;
          if (!lb) pb.x = INFINITY;
          const bool two = c0 + 32 < total;
          c0 += 64;
          test(sa, pa);
          if (two) test(sb, pb);
        }
        // end of the sub-tile: everything queued goes out, the running counts are the segment boundaries
#pragma unroll
        for (int g = 0; g < GRP; g++) {
          if (g < m) {
            flush(g, qa[g]); qa[g] = s_q + 2u * MQ * g;
            if (s == 0) c0s[g] = cnt[g];
            if (s == 1) c1s[g] = cnt[g];
          }
        }
      }
      // ---- publish
#pragma unroll
      for (int g = 0; g < GRP; g++) {
        if (g < m && lane == 0) {
          const bool ok = !((ovf >> g) & 1u);
          V.nbr_off[ig[g]] = ok ? 0 : -1;
          V.nbr_cnt[ig[g]] = ok ? cnt[g] : -1;
          V.mseg[ig[g]] = c0s[g] | (c1s[g] << 16);
          if (ok) s_any[q] = 1; else atomicAdd(&ss->n_uncached, 1);
        }
      }
      e += m;
      __syncwarp();
    }
    __syncthreads();   // the tables are rebuilt for the next bucket
    if (threadIdx.x < 4) blk_any[r * 4 + threadIdx.x] = (unsigned char) s_any[threadIdx.x];
  }
}

template <int WARPS>
int launch_mix_scan(psph_ctx *ctx, int kid)
{
  constexpr size_t smem = sizeof(MixCellTable) + sizeof(MixWarp) * WARPS;
  static_assert(smem + 1024 <= 227 * 1024, "bucket tile does not fit in shared memory");
  PSPH_CUDA(cudaFuncSetAttribute(k_mix_scan<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  KernelTimer kt(ctx, kid);
  const int grid = (int) std::min<long long>(ctx->ncells, ctx->sms);
  k_mix_scan<WARPS><<<grid, WARPS * 32, smem, ctx->stream>>>(ctx->cur, psph_grid_view(ctx), ctx->kp, ctx->d_ss, ctx->d_blk_any);
  PSPH_LAUNCH_CHECK();
  return 0;
}

// =========================================================================================
// pair sums off the mixed-h lists
// =========================================================================================
struct MixTable
{
  int beg[MSEG];                               // first row of the segment
  int off[MSEG + 1];                           // first slot of the segment; off[MSEG] = rows of the sub-tile
  int pad[3];
  int wsum[32];
};
static_assert(sizeof(MixTable) % 16 == 0, "MixTable padding");

// every thread of the CTA calls this: table of sub-tile s of block (bucket (cx, cy, cz), x-pair q); blockDim >= MSEG
__device__ __forceinline__ void mix_build(const GridView &G, int cx, int cy, int cz, int q, int s, MixTable &tab)
{
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  int beg = 0, len = 0, base = -1;
  if (t < MSEG) {
    const int tx = MSX * s + t / (MTY * 3), ty = (t / 3) % MTY, k = t % 3;
    mix_segment(G, cx * SUBDIV + 2 * q - MR + tx, cy * SUBDIV - MR + ty, cz - 1 + k, k, beg, len, base);
  }
  int inc = len;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += v; }
  if (lane == 31) tab.wsum[w] = inc;
  __syncthreads();
  int pre = 0;
  for (int k = 0; k < w; k++) pre += tab.wsum[k];
  if (t < MSEG) { tab.beg[t] = beg; tab.off[t] = pre + inc - len; }
  if (t == MSEG - 1) tab.off[MSEG] = pre + inc;
  __syncthreads();
}

// Stage interface on top of psph_gather.cuh / psph_stages.cu (S::I, S::init, S::R):
//   kNCm          doubles per sub-tile row in shared memory
//   kNAm          sums per row (4, or 6 when the stage forms the density sums on the side)
//   kTagMix       the family bit an entry must carry
//   mix_stage(V, row, c, t)                          write sub-tile row t
//   mix_pair(K, st, a[], c, sl, gx, gy, gz, near&)   one listed pair: adds to a[0 .. kNAm); near = partner inside the
//                                                    energy-smoothing kernel (kNAm == 6 only)
//   mix_finish(V, K, ss, i, st, tot[])               the row's totals (own / self term added here) -> outputs
// PSPH_MIX_DYN 1: the warps of a CTA take the block's rows from a shared cursor (one per sub-tile) instead of a fixed share of
// WARPS-th each: rows differ in list length by sub-tile, and a fixed split leaves warps waiting at the sub-tile barriers.  A
// row's arithmetic does not depend on which warp sums it, so results are bit-identical either way.
#ifndef PSPH_MIX_DYN
#define PSPH_MIX_DYN 0
#endif
template <class S, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_mix_gather(DevView V, GridView G, KParams K, StepState *ss, const unsigned char *__restrict__ blk_any)
{
  static_assert(WARPS * 32 >= MSEG, "mix_build needs one thread per segment");
  constexpr int NA = S::kNAm;
  extern __shared__ __align__(16) unsigned char s_raw[];
  MixTable &tab = *reinterpret_cast<MixTable *>(s_raw);
  double *comp = reinterpret_cast<double *>(s_raw + sizeof(MixTable));                     // [kNCm][MCAP]
  double *acc = comp + (size_t) S::kNCm * MCAP;                                            // [MBR][NA]
  int *nnear = reinterpret_cast<int *>(acc + MBR * NA);                                    // [MBR]
  // slot -> row of the staged sub-tile: kept next to the records when the stage lists near partners (NA == 6), else it
  // only lives while the sub-tile is staged (in the first column's storage)
  int *tile_row = NA == 6 ? nnear + MBR : reinterpret_cast<int *>(comp);                   // [MCAP]
  __shared__ int s_next;                                                                  // PSPH_MIX_DYN: next row of the sub-tile
  if (ss->pair_mode != 4) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned short *__restrict__ lists = V.mlist;
  const int nblk = (int) (G.ncells * 4);
  for (int b = blockIdx.x; b < nblk; b += gridDim.x) {
    if (!blk_any[b]) continue;
    const int r = b >> 2, q = b & 3;
    const int rb0 = __ldg(&G.sub_start[r * SUBS + q * 128]), rb1 = __ldg(&G.sub_start[r * SUBS + q * 128 + 128]);
    const int nrow = rb1 - rb0;                                     // <= MBR (k_mix_scan lists no row of a larger block)
    const int cell0 = __ldg(&G.rank2cell[r]);
    const int cz = cell0 % G.ncell[2], cy = (cell0 / G.ncell[2]) % G.ncell[1], cx = cell0 / (G.ncell[2] * G.ncell[1]);
    __syncthreads();                                                // the previous block's finish has read acc
    for (int t = threadIdx.x; t < nrow * NA; t += WARPS * 32) acc[t] = 0.0;
    for (int t = threadIdx.x; t < nrow; t += WARPS * 32) nnear[t] = 0;
    const int per = (nrow + WARPS - 1) / WARPS;
    const int w0 = rb0 + warp * per, w1 = min(w0 + per, rb1);       // this warp's rows
    for (int s = 0; s < MNSUB; s++) {
      mix_build(G, cx, cy, cz, q, s, tab);                          // (two barriers inside: the previous sub-tile is done with)
      if (PSPH_MIX_DYN && threadIdx.x == 0) s_next = 0;             // (read after the staging barrier below)
      const int total = tab.off[MSEG];
      if (threadIdx.x < MSEG) {
        const int sg = threadIdx.x, beg = tab.beg[sg], o = tab.off[sg], len = tab.off[sg + 1] - o;
        for (int k = 0; k < len; k++) tile_row[o + k] = beg + k;
      }
      __syncthreads();
      constexpr int NPT = (MCAP + WARPS * 32 - 1) / (WARPS * 32);   // sub-tile rows per thread
      int myrow[NPT];
#pragma unroll
      for (int u = 0; u < NPT; u++) { const int t = threadIdx.x + u * WARPS * 32; myrow[u] = t < total ? tile_row[t] : -1; }
      if (NA != 6) __syncthreads();                                 // the map shares its storage with the first column
#pragma unroll
      for (int u = 0; u < NPT; u++) {
        const int t = threadIdx.x + u * WARPS * 32;
        if (myrow[u] >= 0) S::mix_stage(V, myrow[u], comp, t);
      }
      __syncthreads();
      // ---- this warp's rows: the entries of sub-tile s of every listed row
#if PSPH_MIX_DYN
      for (;;) {
        int i = 0;
        if (lane == 0) i = atomicAdd(&s_next, 1);
        i = __shfl_sync(FULL, i, 0);
        if (i >= nrow) break;
        i += rb0;
#else
      for (int i = w0; i < w1; i++) {
#endif
        int n = -1, sg = 0;
        if (lane == 0 && (V.flags[i] & F_NEED)) { n = V.nbr_cnt[i]; sg = V.mseg[i]; }
        n = __shfl_sync(FULL, n, 0); sg = __shfl_sync(FULL, sg, 0);
        if (n < 0) continue;
        const int ka = s == 0 ? 0 : (s == 1 ? (sg & 0xffff) : (sg >> 16) & 0xffff);
        const int kb = s == 0 ? (sg & 0xffff) : (s == 1 ? (sg >> 16) & 0xffff : n);
        if (ka >= kb) continue;
        typename S::I st;
        S::init(V, K, ss, i, st);
        const double4 gi = V.geo[i];
        double a[NA];
#pragma unroll
        for (int k = 0; k < NA; k++) a[k] = 0.0;
        const unsigned short *__restrict__ lst = lists + (size_t) i * MSTRIDE;
        int nn = NA == 6 ? nnear[i - rb0] : 0;
        unsigned en = ka + lane < kb ? __ldg(lst + ka + lane) : 0u;
        for (int k0 = ka; k0 < kb; k0 += 32) {
          const unsigned e = en;
          en = k0 + 32 + lane < kb ? __ldg(lst + k0 + 32 + lane) : 0u;
          bool near = false;
          const int sl = (int) (e & MSLOT);
          if (e & S::kTagMix) S::mix_pair(K, st, a, comp, sl, gi.x, gi.y, gi.z, near);
          if constexpr (NA == 6) {
            const unsigned fb = __ballot_sync(FULL, near);
            if (fb) {                                                 // a partner inside the energy-smoothing kernel
              const int pos = nn + __popc(fb & ((1u << lane) - 1u));
              if (near && pos < NEAR_MAX) V.near[NEARW * i + 1 + pos] = tile_row[sl];
              nn += __popc(fb);
            }
          }
        }
        // lane partial sums -> row accumulators in shared memory (this warp owns the row in every sub-tile)
        warp_sum4(a[0], a[1], a[2], a[3], lane);
        if constexpr (NA == 6) { double z0 = 0.0, z1 = 0.0; warp_sum4(a[4], a[5], z0, z1, lane); }
        if (lane == 0) {
          double *ar = acc + (i - rb0) * NA;
#pragma unroll
          for (int k = 0; k < NA; k++) ar[k] += a[k];
          if (NA == 6) nnear[i - rb0] = nn;
        }
      }
    }
#if PSPH_MIX_DYN
    __syncthreads();                                                // every row of every sub-tile is summed
    for (int i = rb0 + (int) threadIdx.x; i < rb1; i += WARPS * 32) {
#else
    __syncwarp();
    // ---- finish this warp's rows (its own accumulators: no barrier needed)
    for (int i = w0 + lane; i < w1; i += 32) {
#endif
      if (!(V.flags[i] & F_NEED) || V.nbr_cnt[i] < 0) continue;
      typename S::I st;
      S::init(V, K, ss, i, st);
      if constexpr (NA == 6) {
        const int nn = nnear[i - rb0];
        V.near[NEARW * i] = nn <= NEAR_MAX ? nn : NEAR_OVF;
        if (nn > NEAR_MAX) atomicAdd(&ss->n_near_ovf, 1);
      }
      S::mix_finish(V, K, ss, i, st, acc + (i - rb0) * NA);
    }
  }
}

template <class S, int WARPS, int MINB>
int launch_mix_gather(psph_ctx *ctx, int kid)
{
  constexpr size_t smem = sizeof(MixTable) + sizeof(double) * ((size_t) S::kNCm * MCAP + MBR * S::kNAm) + sizeof(int) * MBR +
                          (S::kNAm == 6 ? sizeof(int) * MCAP : 0);
  static_assert((smem + 1024) * MINB <= 227 * 1024, "sub-tile does not fit in shared memory");
  PSPH_CUDA(cudaFuncSetAttribute(k_mix_gather<S, WARPS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  KernelTimer kt(ctx, kid);
  const long long nblk = ctx->ncells * 4;
  const int grid = (int) std::min<long long>(nblk, (long long) ctx->sms * MINB);
  k_mix_gather<S, WARPS, MINB><<<grid, WARPS * 32, smem, ctx->stream>>>(ctx->cur, psph_grid_view(ctx), ctx->kp, ctx->d_ss, ctx->d_blk_any);
  PSPH_LAUNCH_CHECK();
  return 0;
}
