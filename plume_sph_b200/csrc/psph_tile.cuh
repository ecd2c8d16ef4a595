// psph_tile.cuh — block-tiled pair search and pair sums of the 3-D SPH stages.
//
// The pair-list gather of psph_gather3.cuh is bound by the L1 gather rate: every record of a partner row is a
// scattered 32-byte access (~1 sector per cycle and SM).  Here the partners' records are staged ONCE per block
// of rows in shared memory and the pair bodies read them from there:
//   block : the rows of one bucket (cell) with subcell x-index in {2q, 2q+1}, q = 0..3 — one contiguous row range
//           of the sorted table (~128 rows on the unit lattice);
//   tile  : every row that can be a partner of a row of the block when the reach is 3 subcells: the subcell
//           columns gx = 8 cx + 2q - 3 .. + 4 (8), gy = 8 cy - 3 .. + 10 (14), and along z the bucket below
//           (subcells 5..7), the bucket itself (0..7) and the bucket above (0..2): 8 x 14 x 3 row segments of
//           the sorted table (~1570 rows on the unit lattice).  A row's position in that enumeration is its SLOT.
//   k_tile_scan   : one CTA per block finds the pairs (the grouped search of psph_gather3.cuh: FP32 pre-filter, the
//                   reference's gates in FP64 inside the error band) and writes per-row lists of 16-bit SLOTS;
//   k_tile_gather : per stage, one CTA per block loads the tile's records into shared memory (struct of arrays
//                   of doubles, coalesced) and sums every listed row of the block off its slot list; the row's own
//                   state comes from its slot of the tile as well (S::tile_init).
// Rows that do not qualify (tile larger than the shared-memory budget, reach other than 3 subcells, a row on a
// subcell face, smoothing lengths other than the table's standard one in the candidate box, queue overflow, no
// list space) are marked and summed by the fused grouped gather (k_gather_grp LM 2), exactly as for the row lists.
#pragma once
#include "psph_gather3.cuh"

#ifndef PSPH_WU
#define PSPH_WU 4              // rounds of 32 list entries per batch of the weight-cache readers (3: 3 % slower at 216^3)
#endif

constexpr int TX = 8, TY = 14;                 // tile columns in x (2 + 2*3) and y (8 + 2*3)
constexpr int TSEG = TX * TY * 3;              // (column, z-bucket) row segments of a tile
constexpr int TCAP = 1920;                     // tile rows that fit in shared memory (13 double columns: 195 KiB)
constexpr int TQCAP = 384;                     // queue slots per row of a group (the scan cannot drain a queue)
constexpr int TCAND = 512;                     // candidates expanded per batch
constexpr int TITEMS = 49 * 3;                 // (column, z-bucket) items of a group's candidate box

struct TileTable
{
  int beg[TSEG];                               // first row of the segment
  int base[TSEG];                              // sub_start index of subcell z = 0 of the segment's column, -1: outside
  int off[TSEG + 1];                           // first slot of the segment; off[TSEG] = tile rows
  int pad[3];                                  // wsum 16-byte aligned (its vector loads must not cover off[TSEG], which
                                               // another thread writes meanwhile: racecheck) and sizeof % 16 == 0
  int wsum[32];
};
static_assert(sizeof(TileTable) % 16 == 0 && offsetof(TileTable, wsum) % 16 == 0, "TileTable padding");

// (beg, len, base) of segment s of the tile of block (cell (cx, cy, cz), sx-pair q)
__device__ __forceinline__ void tile_segment(const GridView &G, int cx, int cy, int cz, int q, int s, int &beg, int &len, int &base)
{
  const int tx = s / (TY * 3), ty = (s / 3) % TY, k = s % 3;
  const int gx = cx * SUBDIV + 2 * q - 3 + tx, gy = cy * SUBDIV - 3 + ty, bz = cz - 1 + k;
  beg = 0; len = 0; base = -1;
  if (gx < 0 || gy < 0 || bz < 0 || gx >= G.ncell[0] * SUBDIV || gy >= G.ncell[1] * SUBDIV || bz >= G.ncell[2]) return;
  const int cell = ((gx / SUBDIV) * G.ncell[1] + gy / SUBDIV) * G.ncell[2] + bz;
  base = __ldg(&G.cell2rank[cell]) * SUBS + (gx % SUBDIV) * (SUBDIV * SUBDIV) + (gy % SUBDIV) * SUBDIV;
  const int zl = k == 0 ? SUBDIV - 3 : 0, zh = k == 2 ? 2 : SUBDIV - 1;
  beg = __ldg(&G.sub_start[base + zl]);
  len = __ldg(&G.sub_start[base + zh + 1]) - beg;
}

// every thread of the CTA calls this; fills the table (exclusive scan of the segment lengths); blockDim >= TSEG
__device__ __forceinline__ void tile_build(const GridView &G, int cx, int cy, int cz, int q, TileTable &tab)
{
  const int s = threadIdx.x, lane = s & 31, w = s >> 5;
  int beg = 0, len = 0, base = -1;
  if (s < TSEG) tile_segment(G, cx, cy, cz, q, s, beg, len, base);
  int inc = len;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += v; }
  if (lane == 31) tab.wsum[w] = inc;
  __syncthreads();
  int pre = 0;
  for (int k = 0; k < w; k++) pre += tab.wsum[k];
  if (s < TSEG) { tab.beg[s] = beg; tab.base[s] = base; tab.off[s] = pre + inc - len; }
  if (s == TSEG - 1) tab.off[TSEG] = pre + inc;
  __syncthreads();
}

__device__ __forceinline__ void sts16(unsigned addr, unsigned v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short) v) : "memory"); }
__device__ __forceinline__ unsigned lds16(unsigned addr)
{
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
  return v;
}

// =========================================================================================
// pair search, one CTA per cell: writes nbr (16-bit block slots, 32-entry = 64-byte units), nbr_off, nbr_cnt
// (-1: no list) and blk_any.  The CELL tile (columns gx = 8 cx - 3 .. + 10, gy likewise, the three z-buckets) is
// enumerated in the same order as the block tiles (x-slab, y, z-bucket, row), so that the tile of block q is the
// contiguous slot range of the x-slabs 2q .. 2q+7 and   block slot = cell slot - first slot of x-slab 2q.
// Per cell: segment table, per-column slot of every z level, slot -> row map and the FP32 records of all tile
// rows in shared memory; per group of rows the candidates of a column are then ONE slot range read off the
// z-level table, and the records come from shared memory.
// =========================================================================================
constexpr int CX = 14, CY = 14;                // columns of the cell tile: 8 + 2*3 in x and in y
constexpr int CSEG = CX * CY * 3;              // (column, z-bucket) segments
constexpr int CCAP = 3584;                     // rows of a cell tile held in shared memory
constexpr int CZL = 16;                        // z levels per column: local z = -3 .. 11 (15 used)

struct CellTable
{
  float4 geo[CCAP];                            // slot -> FP32 record {x - o, y - o, z - o, h}
  int rowmap[CCAP];                            // slot -> row << 1 | own-bucket flag
  int beg[CSEG], base[CSEG], off[CSEG + 4];    // as TileTable; off[CSEG] = tile rows
  unsigned short colz[CX * CY][CZL];           // column -> slot of its first row with local z >= level - 3
  int wsum[32];
};
struct CellWarp                                // per-warp scratch
{
  GrpHdr hdr;
  unsigned short cand[TCAND];                  // candidate slots
  unsigned short q[GRP][TQCAP];                // block slots that passed the gates, per row of the group
};

template <class S, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
k_tile_scan(DevView V, GridView G, KParams K, StepState *ss, unsigned char *__restrict__ blk_any)
{
  static_assert(WARPS * 32 >= CSEG, "one thread per segment");
  extern __shared__ __align__(16) unsigned char s_raw[];
  CellTable &T = *reinterpret_cast<CellTable *>(s_raw);
  CellWarp *scr = reinterpret_cast<CellWarp *>(s_raw + sizeof(CellTable)) + (threadIdx.x >> 5);
  __shared__ int s_any[4];
  if (ss->pair_mode != 2) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  unsigned s_cand, s_q, s_hdr, s_geo;
  asm volatile("mov.u32 %0, %1;" : "=r"(s_cand) : "r"((unsigned) __cvta_generic_to_shared(&scr->cand[0])));
  asm volatile("mov.u32 %0, %1;" : "=r"(s_q) : "r"((unsigned) __cvta_generic_to_shared(&scr->q[0][0])));
  asm volatile("mov.u32 %0, %1;" : "=r"(s_hdr) : "r"((unsigned) __cvta_generic_to_shared(&scr->hdr.hdr_i[0])));
  asm volatile("mov.u32 %0, %1;" : "=r"(s_geo) : "r"((unsigned) __cvta_generic_to_shared(&T.geo[0])));
  if (lane == 0) {
    scr->hdr.kc[0] = K.cutoff; scr->hdr.kc[1] = K.cutoff2; scr->hdr.kc[2] = K.neigh_radius;
    scr->hdr.geo = V.geo; scr->hdr.hlist = V.t.d[D_HLIST];
  }
  __syncwarp();
  const double hmax = ss->hmax;
  const float delta = __uint_as_float(ss->xmax_bits) * 0x1p-21f;
  const float hstd_f = (float) hmax;
  const float ch = (float) (S::kSymH ? 0.5 * K.cutoff : K.cutoff);
  const float c_lo = ch * (1.0f - 0x1p-19f), c_hi = ch * (1.0f + 0x1p-19f);
  // the tiles are laid out for a reach of 3 subcells of rows with the standard smoothing length
  const bool reach_ok = (int) ceil(S::reach_h(K, hmax, hmax) * G.inv_subw - 1e-9) == 3;
  unsigned short *__restrict__ nbr16 = reinterpret_cast<unsigned short *>(V.nbr);
  unsigned wcur = 0, wend = 0;                 // this warp's block of list space (64-byte units)
  const int ncell = (int) G.ncells;

  for (int r = blockIdx.x; r < ncell; r += gridDim.x) {
    const int rb0 = __ldg(&G.sub_start[r * SUBS]), rb1 = __ldg(&G.sub_start[r * SUBS + SUBS]);
    if (rb0 == rb1) { if (threadIdx.x < 4) blk_any[r * 4 + threadIdx.x] = 0; continue; }
    const int cell0 = __ldg(&G.rank2cell[r]);
    const int cz = cell0 % G.ncell[2], cy = (cell0 / G.ncell[2]) % G.ncell[1], cx = cell0 / (G.ncell[2] * G.ncell[1]);
    if (threadIdx.x < 4) s_any[threadIdx.x] = 0;
    // ---- the cell tile: segment table (exclusive scan of the segment lengths)
    {
      const int s = threadIdx.x;
      int beg = 0, len = 0, base = -1;
      if (s < CSEG) {
        const int tx = s / (CY * 3), ty = (s / 3) % CY, k = s % 3;
        const int gx = cx * SUBDIV - 3 + tx, gy = cy * SUBDIV - 3 + ty, bz = cz - 1 + k;
        if (gx >= 0 && gy >= 0 && bz >= 0 && gx < G.ncell[0] * SUBDIV && gy < G.ncell[1] * SUBDIV && bz < G.ncell[2]) {
          const int cc = ((gx / SUBDIV) * G.ncell[1] + gy / SUBDIV) * G.ncell[2] + bz;
          base = __ldg(&G.cell2rank[cc]) * SUBS + (gx % SUBDIV) * (SUBDIV * SUBDIV) + (gy % SUBDIV) * SUBDIV;
          const int zl = k == 0 ? SUBDIV - 3 : 0, zh = k == 2 ? 2 : SUBDIV - 1;
          beg = __ldg(&G.sub_start[base + zl]);
          len = __ldg(&G.sub_start[base + zh + 1]) - beg;
        }
      }
      int inc = len;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += v; }
      if (lane == 31) T.wsum[warp] = inc;
      __syncthreads();
      int pre = 0;
      for (int k = 0; k < warp; k++) pre += T.wsum[k];
      if (s < CSEG) { T.beg[s] = beg; T.base[s] = base; T.off[s] = pre + inc - len; }
      if (s == CSEG - 1) T.off[CSEG] = pre + inc;
      __syncthreads();
    }
    const int ctotal = T.off[CSEG];
    const bool cell_ok = reach_ok && ctotal <= CCAP;
    if (cell_ok) {
      // slot -> row map (with the own-bucket flag), per-column slot of every z level
      if (threadIdx.x < CSEG) {
        const int s = threadIdx.x, tx = s / (CY * 3), ty = (s / 3) % CY, k = s % 3;
        const int own = (k == 1 && tx >= 3 && tx < 3 + SUBDIV && ty >= 3 && ty < 3 + SUBDIV) ? 1 : 0;
        const int beg = T.beg[s], o = T.off[s], len = T.off[s + 1] - o;
        for (int j = 0; j < len; j++) T.rowmap[o + j] = ((beg + j) << 1) | own;
      }
      for (int idx = threadIdx.x; idx < CX * CY * CZL; idx += WARPS * 32) {
        const int col = idx / CZL, lvl = idx - col * CZL;            // local z = lvl - 3
        const int k = lvl < 3 ? 0 : (lvl < 3 + SUBDIV ? 1 : 2), s = col * 3 + k;
        const int sz = k == 0 ? SUBDIV - 3 + lvl : (k == 1 ? lvl - 3 : lvl - 3 - SUBDIV); // subcell z inside that bucket
        int slot;
        if (lvl >= 3 + SUBDIV + 3) slot = T.off[col * 3 + 3];        // past the last level: end of the column
        else if (T.base[s] < 0) slot = T.off[s];
        else slot = T.off[s] + (__ldg(&G.sub_start[T.base[s] + sz]) - T.beg[s]);
        T.colz[col][lvl] = (unsigned short) slot;
      }
      __syncthreads();
      for (int t = threadIdx.x; t < ctotal; t += WARPS * 32) T.geo[t] = V.geo32[T.rowmap[t] >> 1];
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
      if (!(nm & 1u)) { e += nm ? __ffs(nm) - 1 : GRP; continue; } // skip rows that are not summed
      const unsigned long long pk0 = __shfl_sync(FULL, pk, 0);
      const bool same = need && ((pk ^ pk0) & 0xffffffffffull) == 0;
      const int m = __ffs(~__ballot_sync(FULL, same)) - 1; // 1..GRP
      const bool mine = lane < m;
      const int lx = (int) (pk0 & 0xfffff) - cx * SUBDIV, ly = (int) ((pk0 >> 20) & 0xfffff) - cy * SUBDIV;
      const int q = lx >> 1;                                         // the block of the group
      const int qbase = T.off[2 * q * CY * 3];                       // first slot of the block's tile
      const bool tile_ok = cell_ok && T.off[(2 * q + TX) * CY * 3] - qbase <= TCAP;
      float4 p32 = make_float4(0.f, 0.f, 0.f, 0.f);
      int sz = 0, myself = -1;
      if (mine) {
        sz = (int) ((pk >> 40) & 0xfffff) - cz * SUBDIV; sts32(s_hdr + 4u * lane, row);
        const int s = ((lx + 3) * CY + ly + 3) * 3 + 1;
        myself = T.off[s] + (row - T.beg[s]);
        p32 = cell_ok ? T.geo[myself] : V.geo32[row];
      }
      // rows on a subcell face (reach widened by one) or with another smoothing length do not fit the tile layout
      const bool fits = tile_ok && !__any_sync(FULL, mine && ((pk >> 60) != 0 || p32.w != hstd_f));
      if (!fits) {
        if (mine) { V.nbr_cnt[row] = -1; V.nbr_off[row] = -1; }
        if (lane == 0) atomicAdd(&ss->n_uncached, m);
        e += m;
        continue;
      }
      const int zlo = __reduce_min_sync(FULL, mine ? sz : 0x7fffffff), zhi = __reduce_max_sync(FULL, mine ? sz : -1);
      bool nonstd = false;
      unsigned ovf = 0;
      int ig[GRP], selfslot[GRP];
      unsigned qa[GRP];
      float fx[GRP], fy[GRP], fz[GRP], blo[GRP], bhi[GRP];
#pragma unroll
      for (int g = 0; g < GRP; g++) {
        ig[g] = __shfl_sync(FULL, row, g);
        selfslot[g] = __shfl_sync(FULL, myself, g);
        fx[g] = __shfl_sync(FULL, p32.x, g); fy[g] = __shfl_sync(FULL, p32.y, g); fz[g] = __shfl_sync(FULL, p32.z, g);
        const float hf = __shfl_sync(FULL, p32.w, g);
        blo[g] = hf * c_lo - delta; bhi[g] = hf * c_hi + delta;
        qa[g] = s_q + 2u * TQCAP * g;
      }
      // the group's box: cell-tile columns (lx + ix, ly + iy), ix, iy = 0..6; z levels [zlo, zhi + 7) (local z + 3)
      const int lv0 = zlo, lv1 = zhi + 7;
      int col_base = -32, c0 = 0, total = 0, slot = 0, len = 0;

      auto test = [&](unsigned sl, const float4 &p) {
        const float hlo = S::kSymH ? p.w * c_lo : 0.f, hhi = S::kSymH ? p.w * c_hi : 0.f;
        nonstd |= p.w != hstd_f;
        unsigned okb = 0, bandb = 0;
#pragma unroll
        for (int g = 0; g < GRP; g++) {
          if (g < m) {
            const float m3 = fmaxf(fmaxf(fabsf(fx[g] - p.x), fabsf(fy[g] - p.y)), fabsf(fz[g] - p.z));
            bool ok = m3 <= bhi[g] + hhi;
            if (S::kSkipSelf) ok = ok && (int) sl != selfslot[g];
            if (ok) okb |= 1u << g;
            if (ok && m3 > blo[g] + hlo) bandb |= 1u << g;
          }
        }
        if (__any_sync(FULL, bandb != 0)) { // within the FP32 error band of a cutoff face: the reference's own test
          if (bandb) okb = gate_exact_grp(&scr->hdr, okb, bandb, T.rowmap[sl], S::kSymH ? 1 : 0);
        }
#pragma unroll
        for (int g = 0; g < GRP; g++) {
          if (g < m) {
            const bool ok = (okb >> g) & 1u;
            const unsigned bal = __ballot_sync(FULL, ok);
            if (ok) sts16(qa[g] + 2u * __popc(bal & lt), sl - (unsigned) qbase);
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
            if (col_base >= 49) break;
            const int c = col_base + lane;
            len = 0;
            if (c < 49) {
              const int ix = c / 7, iy = c - ix * 7, col = (lx + ix) * CY + ly + iy;
              slot = T.colz[col][lv0];
              len = (int) T.colz[col][lv1] - slot;
            }
          }
          __syncwarp();
          if (total == 0) break;
        }
#pragma unroll
        for (int g = 0; g < GRP; g++)
          if (qa[g] > s_q + 2u * (TQCAP * g + TQCAP - 64)) { ovf |= 1u << g; qa[g] = s_q + 2u * TQCAP * g; bhi[g] = -INFINITY; }
        // two candidates per lane
        const int ca = c0 + lane, cb = ca + 32;
        const bool la = ca < total, lb = cb < total;
        const unsigned sa = la ? lds16(s_cand + 2u * ca) : 0u, sb = lb ? lds16(s_cand + 2u * cb) : 0u;
        float4 pa = T.geo[sa], pb = T.geo[sb];
        if (!la) { pa.x = INFINITY; pa.w = hstd_f; }
        if (!lb) { pb.x = INFINITY; pb.w = hstd_f; }
        const bool two = c0 + 32 < total;
        c0 += 64;
        test(sa, pa);
        if (two) test(sb, pb);
      }
      // ---- publish the lists
      const bool guni = !__any_sync(FULL, nonstd);
      __syncwarp();
#pragma unroll
      for (int g = 0; g < GRP; g++) {
        if (g < m) {
          const int n = (int) (qa[g] - (s_q + 2u * TQCAP * g)) >> 1;
          int off = -1;
          if (guni && !((ovf >> g) & 1u)) {
            const unsigned units = (unsigned) (n + 31) >> 5;
            if (wcur + units > wend) { // next block of list space for this warp
              unsigned o = 0;
              if (lane == 0) o = atomicAdd(&ss->nbr_cursor, NBR_BLOCK);
              o = __shfl_sync(FULL, o, 0);
              const unsigned cap = ss->nbr_cap;
              wcur = min(o, cap); wend = min(o + NBR_BLOCK, cap);
              if (o + NBR_BLOCK < o) wcur = wend = cap;
            }
            if (wcur + units <= wend) { off = (int) wcur; wcur += units; }
          }
          if (lane == 0) {
            V.nbr_off[ig[g]] = off;
            V.nbr_cnt[ig[g]] = off >= 0 ? n : -1;
            if (off < 0) atomicAdd(&ss->n_uncached, 1);
          }
          if (off >= 0) {
            if (lane == 0) s_any[q] = 1;
            unsigned *__restrict__ dst = reinterpret_cast<unsigned *>(nbr16 + (size_t) (unsigned) off * 32u);
            const unsigned qb = s_q + 2u * TQCAP * g;
            for (int k = lane; 2 * k < n; k += 32) dst[k] = (unsigned) lds32(qb + 4u * k); // two slots per word
          }
        }
      }
      e += m;
      __syncwarp();
    }
    __syncthreads();   // the tables are rebuilt for the next cell
    if (threadIdx.x < 4) blk_any[r * 4 + threadIdx.x] = (unsigned char) s_any[threadIdx.x];
  }
}

// tile rows as plain double columns, x, y, z first (smooth_density, smooth_velocity): staging and position read-back
template <class S>
__device__ __forceinline__ void tile_stage_soa(const DevView &V, int row, double *c, int t)
{
  const double4 g = ld_rec(V.geo, row);
  c[0 * TCAP + t] = g.x; c[1 * TCAP + t] = g.y; c[2 * TCAP + t] = g.z;
  S::tile_load(V, row, c, t);
}
__device__ __forceinline__ void tile_xyz_soa(const double *c, int t, double &x, double &y, double &z)
{ x = c[0 * TCAP + t]; y = c[1 * TCAP + t]; z = c[2 * TCAP + t]; }

// =========================================================================================
// pair sums off the slot lists; S::kNC double columns per tile row in shared memory: x, y, z + the stage's own
// (S::tile_load writes them, S::tile_rec reads them back as the stage record)
// =========================================================================================
// WM — the pair-weight cache (V.wts, one double per list entry, parallel to the slot lists): between the three stages
// of a step neither positions nor smoothing lengths change and every listed pair carries the table's standard h, so
// the kernel weight W of a pair is the same number in all three.  0: no cache, every stage computes its weights;
// 1: the stage computes W and stores it (mom_engr_update, the first stage of a step); 2: the stage reads W instead of
// computing it (smooth_density, smooth_velocity) — 8 bytes of HBM per pair instead of an exp() on the FP64 pipe.
// W > 0, so its sign bit is free: the writer sets it on the pairs that lie inside smooth_density's second, narrower
// kernel (energy smoothing) — a handful per row — and that stage runs the second body on exactly those.
template <class S, int WARPS, int MINB, int WM>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_tile_gather(DevView V, GridView G, KParams K, StepState *ss, const unsigned char *__restrict__ blk_any)
{
  static_assert(WARPS * 32 >= TSEG, "tile_build needs one thread per segment");
  static_assert(!(S::kDefer && WM != 0), "the weight cache runs on the in-place loop");
  extern __shared__ __align__(16) unsigned char s_raw[];
  TileTable &tab = *reinterpret_cast<TileTable *>(s_raw);
  double *comp = reinterpret_cast<double *>(s_raw + sizeof(TileTable));       // [kNC][TCAP]
  // per-warp queue of deferred pairs (S::kDefer), 64 slots
  constexpr int NCOL = (WM == 2 && !S::kCachedNeedsX) ? S::kNCc : S::kNC;
  const unsigned s_dq = (unsigned) __cvta_generic_to_shared(comp + NCOL * TCAP) + 128u * (threadIdx.x >> 5);
  int *tile_row = reinterpret_cast<int *>(comp);                              // [TCAP] slot -> row, while the tile is staged
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned short *__restrict__ nbr16 = reinterpret_cast<const unsigned short *>(V.nbr);
  if (ss->pair_mode != 2) return;                                   // the step runs on another pair path (k_pair_mode)
  if constexpr (WM == 2 && S::kSkipUnlessFull)
    if (!ss->dens_full) return;                                     // k_dens_near has served every listed row
  const int nblk = (int) (G.ncells * 4);
  // every listed row carries the table's standard smoothing length: 1/h, the kernel norms and the cutoff bounds are
  // constants of the launch, not of the row
  typename S::I st;
  S::tile_uni(K, ss->hmax, st);
  [[maybe_unused]] double wself = 0.0;
  if constexpr (WM == 2) wself = S::self_weight(st);
  for (int b = blockIdx.x; b < nblk; b += gridDim.x) {
    if (!blk_any[b]) continue;
    const int r = b >> 2, q = b & 3;
    const int rb0 = __ldg(&G.sub_start[r * SUBS + q * 128]), rb1 = __ldg(&G.sub_start[r * SUBS + q * 128 + 128]);
    const int cell0 = __ldg(&G.rank2cell[r]);
    const int cz = cell0 % G.ncell[2], cy = (cell0 / G.ncell[2]) % G.ncell[1], cx = cell0 / (G.ncell[2] * G.ncell[1]);
    tile_build(G, cx, cy, cz, q, tab);
    // ---- stage the tile: slot -> row map first (one thread per segment), then one thread per tile row so that
    // all record loads of the tile are in flight together
    const int total = tab.off[TSEG];
    if (threadIdx.x < TSEG) {
      const int s = threadIdx.x, beg = tab.beg[s], o = tab.off[s], len = tab.off[s + 1] - o;
      for (int k = 0; k < len; k++) tile_row[o + k] = beg + k;
    }
    __syncthreads();
    constexpr int NPT = (TCAP + WARPS * 32 - 1) / (WARPS * 32);  // tile rows per thread
    int myrow[NPT];
#pragma unroll
    for (int u = 0; u < NPT; u++) { const int t = threadIdx.x + u * WARPS * 32; myrow[u] = t < total ? tile_row[t] : -1; }
    __syncthreads();                                              // the map shares its storage with the first column
#pragma unroll
    for (int u = 0; u < NPT; u++) {
      const int t = threadIdx.x + u * WARPS * 32, row = myrow[u];
      if (row >= 0) {
        if constexpr (WM == 2 && !S::kCachedNeedsX) S::tile_load_c(V, row, comp, t);
        else S::tile_stage(V, row, comp, t);
      }
    }
    __syncthreads();
    // ---- the rows of the block, a contiguous share per warp (at most 32: the row headers are read lane-parallel)
    const int per = (rb1 - rb0 + WARPS - 1) / WARPS;
    for (int eb = rb0 + warp * per, e1 = min(eb + per, rb1); eb < e1; eb += 32) {
      int hn = -1, hoff = 0, hself = -1;
      if (eb + lane < e1) {
        const int i = eb + lane;
        hn = (V.flags[i] & F_NEED) ? V.nbr_cnt[i] : -1;
        hoff = V.nbr_off[i];
        // the row's own slot: its state is read from the tile, and stages that count the self pair append it
        const unsigned long long pk = V.gsub[i];
        const int tx = (int) (pk & 0xfffff) - cx * SUBDIV - 2 * q + 3, ty = (int) ((pk >> 20) & 0xfffff) - cy * SUBDIV + 3;
        const int s = (tx * TY + ty) * 3 + 1;
        hself = tab.off[s] + (i - tab.beg[s]);
      }
      const int nrows = min(32, e1 - eb);
      for (int rr = 0; rr < nrows; rr++) {
        const int n = __shfl_sync(FULL, hn, rr);
        if (n < 0) continue;
        const int i = eb + rr, self = __shfl_sync(FULL, hself, rr);
        const unsigned loff = (unsigned) __shfl_sync(FULL, hoff, rr);
        const unsigned short *__restrict__ lst = nbr16 + (size_t) loff * 32u + lane;
        if constexpr (WM == 2 && !S::kCachedNeedsX) S::tile_init_c(V, K, ss, comp, self, i, st);
        else S::tile_init(V, K, ss, comp, self, i, st);
        typename S::A acc;
        S::begin(K, st, acc, 0);
        const int ntot = n + (S::kSkipSelf ? 0 : 1);
        if constexpr (WM == 2) {
          // weights from the cache: WU rounds of 32 entries per batch; the slot and weight loads of the next batch
          // are issued before the pair bodies of the current one (the loop is bound by the rate at which the 8-byte
          // weights stream in from HBM, not by arithmetic)
          constexpr int WU = PSPH_WU;
          const double *__restrict__ wl = V.wts + (size_t) loff * 32u + lane;
          int sl[WU];
          double w[WU];
          auto fetch = [&](int q0) {
#pragma unroll
            for (int u = 0; u < WU; u++) {
              const int k = q0 + 32 * u + lane;
              sl[u] = k < n ? (int) __ldg(lst + (k - lane)) : (k < ntot ? self : -1);
              w[u] = k < n ? __ldcs(wl + (k - lane)) : wself;
            }
          };
          fetch(0);
          for (int q0 = 0; q0 < ntot; q0 += 32 * WU) {
            int sc[WU];
            double wc[WU];
#pragma unroll
            for (int u = 0; u < WU; u++) { sc[u] = sl[u]; wc[u] = w[u]; }
            if (q0 + 32 * WU < ntot) fetch(q0 + 32 * WU);
#pragma unroll
            for (int u = 0; u < WU; u++)
              if (sc[u] >= 0) S::pair_c(K, st, acc, wc[u], comp, sc[u], self);
          }
          S::end(V, K, ss, i, st, acc, 0, lane);
          continue;
        }
        double gix, giy, giz;
        S::tile_xyz(comp, self, gix, giy, giz);
        auto entry = [&](int k) { return k < n ? (int) __ldg(lst + (k - lane)) : (k < ntot ? self : -1); };
        int jn = entry(lane);
        if constexpr (S::kDefer) {
          // second-kernel pairs are rare (a few per row): queue their slots, run them 32 at a time
          int dn = 0;
          auto inner = [&](int cnt) {
            if (lane < cnt) {
              const int sl = (int) lds16(s_dq + 2u * lane);
              double px, py, pz;
              const typename S::R rec = S::tile_recp(comp, sl, px, py, pz);
              S::template pair_inner<true>(K, st, acc, gix - px, giy - py, giz - pz, st.h, rec);
            }
          };
          for (int q0 = 0; q0 < ntot; q0 += 32) {
            const int sl = jn;
            jn = entry(q0 + 32 + lane);
            bool in = false;
            if (sl >= 0) {
              double px, py, pz;
              const typename S::R rec = S::tile_recp(comp, sl, px, py, pz);
              in = S::template pair_main<true>(K, st, acc, gix - px, giy - py, giz - pz, st.h, rec);
            }
            const unsigned bal = __ballot_sync(FULL, in);
            if (in) sts16(s_dq + 2u * (dn + __popc(bal & ((1u << lane) - 1u))), (unsigned) sl);
            dn += __popc(bal);
            if (dn >= 32) {
              __syncwarp();
              inner(32);
              const unsigned v = lane + 32 < dn ? lds16(s_dq + 2u * (lane + 32)) : 0u;
              __syncwarp();
              if (lane + 32 < dn) sts16(s_dq + 2u * lane, v);
              dn -= 32;
            }
          }
          __syncwarp();
          inner(dn);
          __syncwarp();
        } else {
          // WM == 1 also prepares smooth_density (k_dens_near): the phase-density sums over the list (d1, d2) and the
          // rows of the partners inside the energy-smoothing kernel (V.near)
          [[maybe_unused]] double d1 = 0.0, d2 = 0.0;
          [[maybe_unused]] int nnear = 0;
          for (int q0 = 0; q0 < ntot; q0 += 32) {
            const int sl = jn;
            jn = entry(q0 + 32 + lane);                              // (two rounds ahead measured 2 % slower)
            [[maybe_unused]] bool fl = false;
            if (sl >= 0) {
              double px, py, pz;
              const typename S::R rec = S::tile_recp(comp, sl, px, py, pz);
              const double dx = gix - px, dy = giy - py, dz = giz - pz;
              if constexpr (WM == 1) {
                const double w = S::template pair_w<true>(K, st, acc, dx, dy, dz, st.h, rec);
                fl = S::cache_flag(K, st, dx, dy, dz);
                // entry q0 + lane of the row's list; the sign carries one bit for the reading stages (S::cache_flag)
                __stcs(V.wts + (size_t) loff * 32u + (unsigned) (q0 + lane), fl ? -w : w);
                const double md = rec.r2.w, wm = w * fabs(md);       // density_massfrac.cc:159-161
                d2 += md < 0.0 ? wm : 0.0;
                d1 += md < 0.0 ? 0.0 : wm;
              } else
                S::template pair_r<true>(K, st, acc, dx, dy, dz, st.h, rec);
            }
            if constexpr (WM == 1) {
              const unsigned fb = __ballot_sync(FULL, fl);
              if (fb) {                                               // rare: a partner within 0.75 h of the row
                const int pos = nnear + __popc(fb & ((1u << lane) - 1u));
                if (fl && pos < NEAR_MAX) {
                  int lo = 0, hi = TSEG;                              // the segment of the slot -> its row
                  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (tab.off[mid] <= sl) lo = mid; else hi = mid; }
                  V.near[NEARW * i + 1 + pos] = tab.beg[lo] + (sl - tab.off[lo]);
                }
                nnear += __popc(fb);
              }
            }
          }
          if constexpr (WM == 1) {
            double z0 = 0.0, z1 = 0.0;
            warp_sum4(d1, d2, z0, z1, lane);
            if (lane == 0) {
              V.scr[4][i] = d1; V.scr[5][i] = d2;
              // a near list that overflowed flags ITS row: smooth_density sums that row by itself (k_gather<DensStage, OVF>)
              V.near[NEARW * i] = nnear <= NEAR_MAX ? nnear : NEAR_OVF;
              if (nnear > NEAR_MAX) atomicAdd(&ss->n_near_ovf, 1);
            }
          }
        }
        S::end(V, K, ss, i, st, acc, 0, lane);
      }
    }
    __syncthreads();   // the tile is overwritten for the next block
  }
}

template <class S, int WARPS>
int launch_tile_scan(psph_ctx *ctx, int kid)
{
  constexpr size_t smem = sizeof(CellTable) + sizeof(CellWarp) * WARPS;
  static_assert(smem <= 227 * 1024, "cell tile does not fit in shared memory");
  PSPH_CUDA(cudaFuncSetAttribute(k_tile_scan<S, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  KernelTimer kt(ctx, kid);
  const int grid = (int) std::min<long long>(ctx->ncells, ctx->sms);
  k_tile_scan<S, WARPS><<<grid, WARPS * 32, smem, ctx->stream>>>(ctx->cur, psph_grid_view(ctx), ctx->kp, ctx->d_ss, ctx->d_blk_any);
  PSPH_LAUNCH_CHECK();
  return 0;
}
template <class S, int WARPS, int MINB, int WM>
int launch_tile_gather_wm(psph_ctx *ctx, int kid)
{
  constexpr int nc = (WM == 2 && !S::kCachedNeedsX) ? S::kNCc : S::kNC;
  constexpr size_t smem = sizeof(TileTable) + sizeof(double) * nc * TCAP + (S::kDefer ? 128 * WARPS : 0);
  static_assert((smem + 1024) * MINB <= 227 * 1024, "tiles do not fit in shared memory");
  PSPH_CUDA(cudaFuncSetAttribute(k_tile_gather<S, WARPS, MINB, WM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  KernelTimer kt(ctx, kid);
  const long long nblk = ctx->ncells * 4;
  const int grid = (int) std::min<long long>(nblk, (long long) ctx->sms * MINB);
  k_tile_gather<S, WARPS, MINB, WM><<<grid, WARPS * 32, smem, ctx->stream>>>(ctx->cur, psph_grid_view(ctx), ctx->kp, ctx->d_ss, ctx->d_blk_any);
  PSPH_LAUNCH_CHECK();
  return 0;
}
// with the weight cache allocated (psph_nbr_prepare) the stage runs in its cache role S::kWCache (1 writes, 2 reads)
template <class S, int WARPS, int MINB>
int launch_tile_gather(psph_ctx *ctx, int kid)
{
  if (ctx->cur.wts) return launch_tile_gather_wm<S, WARPS, MINB, S::kWCache>(ctx, kid);
  return launch_tile_gather_wm<S, WARPS, MINB, 0>(ctx, kid);
}
