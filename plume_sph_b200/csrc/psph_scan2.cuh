// psph_scan2.cuh — pair search of the block-tiled path, one LANE per row.
//
// k_tile_scan (psph_tile.cuh) gives every lane one CANDIDATE and tests it against the up-to-4 rows of a group; each
// accepted (candidate, row) pair then costs a ballot, a popc and a predicated store into the row's queue — about
// 17 instructions per test, 1050 warp instructions per row on the unit lattice.  Here the roles are swapped:
//   * a warp takes a CHUNK of 32 consecutive rows of the bucket, one row per lane;
//   * the candidates of the chunk are the tile rows of the union of the lanes' boxes (all subcell columns within 3
//     subcells of a lane's column, the z levels within 3 of the chunk's z range); their slots are expanded ONCE into a
//     per-warp table in shared memory;
//   * every lane tests every candidate of the table against its own row: the candidate's FP32 record is a BROADCAST
//     load from the bucket's tile in shared memory, the test is 3 FADD, 2 FMNMX and 2 FSETP on chunk constants (every
//     row and every candidate carry the table's standard smoothing length, checked during the expansion), and the
//     result is ONE bit or-ed into a register: 12 instructions per candidate for 32 rows, no cross-lane operation;
//   * the bit masks (32 candidates per word) go to shared memory; afterwards the warp turns each row's mask into
//     its slot list (lanes take words, popc + prefix sum, expand the set bits through the slot table into a
//     staging buffer, coalesced copy to the global list).
// Candidates within the FP32 error band of a cutoff face are decided by the reference's own gates in FP64, exactly
// as in k_tile_scan (gate_exact_pair).  (Measured and dropped: every lane expanding an equal share of a row's set
// bits, found by a search in a prefix table, plus clearing the self bit once per round instead of testing it per
// candidate — 1.2 ms slower per step at 216^3 than the word-per-lane expansion below.)  Lists, nbr_off / nbr_cnt, blk_any and n_uncached have the same meaning;
// the order of a list is the candidate order (x column, y column, z) — runs of consecutive slots as before.
#pragma once
#include "psph_tile.cuh"

constexpr int S2_NW = 36;                      // mask words per row: up to 1152 candidates per chunk
constexpr int S2_CAP = S2_NW * 32;

struct Scan2Warp
{
  unsigned mask[S2_NW][33];                    // [word][lane] (+1: conflict-free both by word and by lane)
  unsigned short cslot[S2_CAP];                // candidate -> slot of the bucket tile
  unsigned short stage[TQCAP];                 // one row's list on its way out (block slots)
};
static_assert(sizeof(Scan2Warp) % 16 == 0, "Scan2Warp alignment");

__device__ __forceinline__ float4 lds_f4(unsigned addr)
{
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// The reference's gates in FP64 for row i against row j (see gate_exact_grp): kernel cutoff, list gate, in_support
__device__ __noinline__ bool gate_exact_pair(const double4 *__restrict__ geo, const double *__restrict__ hlist, double cutoff,
                                             double cutoff2, double nradius, int i, int j, int own, int sym)
{
  const double4 gi = geo[i], gj = geo[j];
  const double dx = gi.x - gj.x, dy = gi.y - gj.y, dz = gi.z - gj.z;
  const double h = sym ? 0.5 * (gi.w + gj.w) : gi.w;
  if (fabs(dx / h) > cutoff || fabs(dy / h) > cutoff || fabs(dz / h) > cutoff) return false;
  const double supp = cutoff2 * gi.w;
  const double lim = own ? supp : fmin(supp, nradius * hlist[i]);
  return !(fabs(dx) > lim || fabs(dy) > lim || fabs(dz) > lim);
}

template <class S, int WARPS, int WORK>
__global__ void __launch_bounds__(WARPS * 32, 1)
k_tile_scan2(DevView V, GridView G, KParams K, StepState *ss, unsigned char *__restrict__ blk_any)
{
  static_assert(WARPS * 32 >= CSEG, "one thread per segment");
  static_assert(WORK <= WARPS, "worker warps");
  extern __shared__ __align__(16) unsigned char s_raw[];
  CellTable &T = *reinterpret_cast<CellTable *>(s_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Scan2Warp &W = reinterpret_cast<Scan2Warp *>(s_raw + sizeof(CellTable))[warp < WORK ? warp : 0];
  __shared__ int s_any[4];
  if (ss->pair_mode != 2) return;              // the step runs on another pair path (k_pair_mode)
  unsigned s_geo, s_cslot;                     // 32-bit shared-window addresses (one base register each)
  asm volatile("mov.u32 %0, %1;" : "=r"(s_geo) : "r"((unsigned) __cvta_generic_to_shared(&T.geo[0])));
  asm volatile("mov.u32 %0, %1;" : "=r"(s_cslot) : "r"((unsigned) __cvta_generic_to_shared(&W.cslot[0])));
  const double hmax = ss->hmax;
  const float delta = __uint_as_float(ss->xmax_bits) * 0x1p-21f;
  const float hstd_f = (float) hmax;
  const float ch = (float) (S::kSymH ? 0.5 * K.cutoff : K.cutoff);
  const float c_lo = ch * (1.0f - 0x1p-19f), c_hi = ch * (1.0f + 0x1p-19f);
  // every listed row and every candidate carry the standard smoothing length: the FP32 bounds are constants
  const float BLO = (hstd_f * c_lo - delta) + (S::kSymH ? hstd_f * c_lo : 0.f);
  const float BHI = (hstd_f * c_hi + delta) + (S::kSymH ? hstd_f * c_hi : 0.f);
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
    // ---- the bucket's tile (as in k_tile_scan): segment table, slot -> row map, z-level slots, FP32 records
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
      if (threadIdx.x < CSEG) {
        const int s = threadIdx.x, tx = s / (CY * 3), ty = (s / 3) % CY, k = s % 3;
        const int own = (k == 1 && tx >= 3 && tx < 3 + SUBDIV && ty >= 3 && ty < 3 + SUBDIV) ? 1 : 0;
        const int beg = T.beg[s], o = T.off[s], len = T.off[s + 1] - o;
        for (int j = 0; j < len; j++) T.rowmap[o + j] = ((beg + j) << 1) | own;
      }
      for (int idx = threadIdx.x; idx < CX * CY * CZL; idx += WARPS * 32) {
        const int col = idx / CZL, lvl = idx - col * CZL;            // local z = lvl - 3
        const int k = lvl < 3 ? 0 : (lvl < 3 + SUBDIV ? 1 : 2), s = col * 3 + k;
        const int sz = k == 0 ? SUBDIV - 3 + lvl : (k == 1 ? lvl - 3 : lvl - 3 - SUBDIV);
        int slot;
        if (lvl >= 3 + SUBDIV + 3) slot = T.off[col * 3 + 3];
        else if (T.base[s] < 0) slot = T.off[s];
        else slot = T.off[s] + (__ldg(&G.sub_start[T.base[s] + sz]) - T.beg[s]);
        T.colz[col][lvl] = (unsigned short) slot;
      }
      __syncthreads();
      for (int t = threadIdx.x; t < ctotal; t += WARPS * 32) T.geo[t] = V.geo32[T.rowmap[t] >> 1];
    }
    __syncthreads();
    // ---- chunks of 32 consecutive rows, one per worker warp at a time
    const int nchunk = (rb1 - rb0 + 31) >> 5;
    for (int kc = warp; warp < WORK && kc < nchunk; kc += WORK) {
      const int row = rb0 + 32 * kc + lane < rb1 ? rb0 + 32 * kc + lane : -1;
      int fl = 0;
      unsigned long long pk = 0;
      if (row >= 0) { fl = V.flags[row]; pk = V.gsub[row]; }
      const bool need = row >= 0 && (fl & F_NEED);
      const unsigned needm = __ballot_sync(FULL, need);
      if (!needm) continue;
      const int lx = (int) (pk & 0xfffff) - cx * SUBDIV, ly = (int) ((pk >> 20) & 0xfffff) - cy * SUBDIV,
                sz = (int) ((pk >> 40) & 0xfffff) - cz * SUBDIV;
      const int q = lx >> 1;                                         // the block of the row
      int qbase = 0, myself = -1;
      bool fits = false;
      if (need && cell_ok) {
        qbase = T.off[2 * q * CY * 3];                               // first slot of the block's tile
        const bool tile_ok = T.off[(2 * q + TX) * CY * 3] - qbase <= TCAP;
        const int s = ((lx + 3) * CY + ly + 3) * 3 + 1;
        myself = T.off[s] + (row - T.beg[s]);
        // rows on a subcell face (reach widened by one) or with another smoothing length do not fit the tile layout
        fits = tile_ok && (pk >> 60) == 0 && T.geo[myself].w == hstd_f;
      }
      const unsigned fitm = __ballot_sync(FULL, fits);
      int ncand = 0;
      bool nonstd = false;
      if (fitm) {
        // ---- the chunk's candidate box: tile columns [x0, x1] x [y0, y1], z levels [lv0, lv1); slots -> W.cslot
        const int x0 = __reduce_min_sync(FULL, fits ? lx : 0x7fffffff), x1 = __reduce_max_sync(FULL, fits ? lx : -1) + 6;
        const int y0 = __reduce_min_sync(FULL, fits ? ly : 0x7fffffff), y1 = __reduce_max_sync(FULL, fits ? ly : -1) + 6;
        const int lv0 = __reduce_min_sync(FULL, fits ? sz : 0x7fffffff), lv1 = __reduce_max_sync(FULL, fits ? sz : -1) + 7;
        const int ny = y1 - y0 + 1, ncol = (x1 - x0 + 1) * ny;
        for (int cb = 0; cb < ncol && ncand <= S2_CAP; cb += 32) {
          const int c = cb + lane;
          int slot = 0, len = 0;
          if (c < ncol) {
            const int col = (x0 + c / ny) * CY + y0 + c % ny;
            slot = T.colz[col][lv0];
            len = (int) T.colz[col][lv1] - slot;
          }
          int inc = len;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += v; }
          const int sum = __shfl_sync(FULL, inc, 31);
          if (ncand + sum <= S2_CAP) {
            const int o = ncand + inc - len;
            for (int k = 0; k < len; k++) {
              W.cslot[o + k] = (unsigned short) (slot + k);
              nonstd |= T.geo[slot + k].w != hstd_f;
            }
          }
          ncand += sum;
        }
        nonstd = __any_sync(FULL, nonstd);
        __syncwarp();
      }
      const bool chunk_ok = fitm && ncand <= S2_CAP && !nonstd;
      int cnt = 0;
      if (chunk_ok) {
        const float4 me = fits ? T.geo[myself] : make_float4(INFINITY, 0.f, 0.f, 0.f);   // rows without a list match nothing
        const int selfslot = S::kSkipSelf ? myself : -1;
        for (int c = 0; c < ncand; c += 32) {
          unsigned acc = 0, band = 0;
          // branch-free: one predicate per candidate, or-ed into the masks as a constant bit (unrolled rounds)
          auto test = [&](int b) {
            const unsigned sl = lds16(s_cslot + 2u * (unsigned) (c + b));
            const float4 p = lds_f4(s_geo + 16u * sl);
            const float m3 = fmaxf(fmaxf(fabsf(me.x - p.x), fabsf(me.y - p.y)), fabsf(me.z - p.z));
            const unsigned ok = (m3 <= BHI) & ((int) sl != selfslot);
            acc |= ok << b;
            band |= (ok & (m3 > BLO)) << b;
          };
          if (c + 32 <= ncand) {
#pragma unroll
            for (int b = 0; b < 32; b++) test(b);
          } else
            for (int b = 0; b < ncand - c; b++) test(b);
          if (__any_sync(FULL, band != 0)) {   // within the FP32 error band of a cutoff face: the reference's own test
            while (band) {
              const int b = __ffs(band) - 1;
              band &= band - 1;
              const int rj = T.rowmap[W.cslot[c + b]];
              if (!gate_exact_pair(V.geo, V.t.d[D_HLIST], K.cutoff, K.cutoff2, K.neigh_radius, row, rj >> 1, rj & 1, S::kSymH ? 1 : 0))
                acc &= ~(1u << b);
            }
          }
          W.mask[c >> 5][lane] = acc;
          cnt += __popc(acc);
        }
      }
      // ---- list space: units of 32 entries, this warp's rows back to back
      const bool listed = chunk_ok && fits && cnt <= TQCAP;
      const int units = listed ? (cnt + 31) >> 5 : 0;
      int inc = units;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += v; }
      const unsigned total = (unsigned) __shfl_sync(FULL, inc, 31);
      bool space = true;
      if (total) {
        if (wcur + total > wend) { // next block of list space for this warp
          unsigned o = 0;
          if (lane == 0) o = atomicAdd(&ss->nbr_cursor, NBR_BLOCK);
          o = __shfl_sync(FULL, o, 0);
          const unsigned cap = ss->nbr_cap;
          wcur = min(o, cap); wend = min(o + NBR_BLOCK, cap);
          if (o + NBR_BLOCK < o) wcur = wend = cap;
        }
        space = wcur + total <= wend;
      }
      const int off = (listed && space) ? (int) (wcur + (unsigned) (inc - units)) : -1;
      if (space) wcur += total;
      if (need) {
        V.nbr_off[row] = off;
        V.nbr_cnt[row] = off >= 0 ? cnt : -1;
        if (off >= 0) s_any[q] = 1;
      }
      const unsigned miss = __ballot_sync(FULL, need && off < 0);
      if (lane == 0 && miss) atomicAdd(&ss->n_uncached, __popc(miss));
      // ---- masks -> slot lists, row by row: lanes take words, expand their set bits, coalesced copy
      const unsigned havem = __ballot_sync(FULL, off >= 0);
      if (!havem) continue;
      const int nw = (ncand + 31) >> 5;
      __syncwarp();
      for (unsigned rm = havem; rm; rm &= rm - 1) {
        const int rr = __ffs(rm) - 1;
        const int n_r = __shfl_sync(FULL, cnt, rr), off_r = __shfl_sync(FULL, off, rr), qb_r = __shfl_sync(FULL, qbase, rr);
        unsigned m0 = lane < nw ? W.mask[lane][rr] : 0u, m1 = lane + 32 < nw ? W.mask[lane + 32][rr] : 0u;
        const int c0 = __popc(m0), c1 = __popc(m1);
        int i0 = c0, i1 = c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v0 = __shfl_up_sync(FULL, i0, o), v1 = __shfl_up_sync(FULL, i1, o);
          if (lane >= o) { i0 += v0; i1 += v1; }
        }
        int p0 = i0 - c0, p1 = __shfl_sync(FULL, i0, 31) + i1 - c1;
        while (m0) {
          const int b = __ffs(m0) - 1;
          m0 &= m0 - 1;
          W.stage[p0++] = (unsigned short) ((int) W.cslot[32 * lane + b] - qb_r);
        }
        while (m1) {
          const int b = __ffs(m1) - 1;
          m1 &= m1 - 1;
          W.stage[p1++] = (unsigned short) ((int) W.cslot[32 * (lane + 32) + b] - qb_r);
        }
        __syncwarp();
        unsigned *__restrict__ dst = reinterpret_cast<unsigned *>(nbr16 + (size_t) (unsigned) off_r * 32u);
        const unsigned *src = reinterpret_cast<const unsigned *>(W.stage);
        for (int k = lane; 2 * k < n_r; k += 32) dst[k] = src[k];   // two slots per word
        __syncwarp();
      }
    }
    __syncthreads();   // the tables are rebuilt for the next bucket
    if (threadIdx.x < 4) blk_any[r * 4 + threadIdx.x] = (unsigned char) s_any[threadIdx.x];
  }
}

template <class S, int WARPS, int WORK>
int launch_tile_scan2(psph_ctx *ctx, int kid)
{
  constexpr size_t smem = sizeof(CellTable) + sizeof(Scan2Warp) * WORK;
  static_assert(smem + 1024 <= 227 * 1024, "bucket tile + chunk tables do not fit in shared memory");
  PSPH_CUDA(cudaFuncSetAttribute(k_tile_scan2<S, WARPS, WORK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  KernelTimer kt(ctx, kid);
  const int grid = (int) std::min<long long>(ctx->ncells, ctx->sms);
  k_tile_scan2<S, WARPS, WORK><<<grid, WARPS * 32, smem, ctx->stream>>>(ctx->cur, psph_grid_view(ctx), ctx->kp, ctx->d_ss, ctx->d_blk_any);
  PSPH_LAUNCH_CHECK();
  return 0;
}
