// psph_gather3.cuh — the grouped pair gather of the 3-D SPH summation stages (mom_engr_update,
// smooth_density, smooth_velocity).
//
// The per-particle gather of psph_gather.cuh spends 60-70 % of its instructions on finding the
// pairs (candidate enumeration + gates), not on the pair bodies.  Here ONE WARP takes a GROUP of
// up to GRP rows that sit in the same (x, y) subcell column of one bucket (consecutive entries of
// the sorted row list), and
//   enumerate : the candidate rows of the group's common box (x, y: column +- reach, z: the
//               group's z-range +- reach) are expanded ONCE into a dense list in shared memory;
//   scan      : every lane takes one candidate, loads its 16-byte FP32 record {x-o, y-o, z-o, h}
//               once and tests it against each row of the group in FP32.  The FP32 test only
//               DECIDES pairs that are clear of the cutoff face by more than its own error bound;
//               candidates inside the band (and every candidate of a row whose list / in_support
//               limits can bite) go through the reference's gates in FP64, bit for bit
//               (neighs.cc:84-92,130; sph_header.h:49-60; sph_lib.cc:216-218).  Rows that pass
//               are appended (ballot + popc) to the queue of that row of the group;
//   pair      : then, row by row, all 32 lanes take one queued pair each and run the stage's pair
//               body in FP64 (the same bodies as the per-particle gather), accumulators in
//               registers, warp-shuffle reduction, one lane writes.
// A queue that fills up before the scan ends is drained early: full rounds are run and the
// per-lane partial sums are parked in a global spill row until the scan of the group resumes.
// The set of contributing pairs is exactly the reference's.
#pragma once
#include "psph_gather.cuh"

constexpr int GRP = 4;          // rows per group
constexpr int GQCAP = 320;      // queue slots per row of the group (multiple of 32, >= 64 + 32)
constexpr int GQCAP_SCAN = 384; // ... of the scan-only kernel (LM 3), which cannot drain a queue
constexpr int GQCAP_TAG = 896;  // ... of the general scan (LM 4): rows after adaptive_sml have ~730 partners
constexpr int GCAND = 512;      // candidate rows expanded per batch
constexpr int RPWG = 32;        // consecutive list entries per warp and chunk
constexpr unsigned NBR_BLOCK = 512; // pair-list units a warp takes from the global cursor at a time
constexpr int GNA = 4;          // accumulators per lane (every stage's A is four doubles)

struct GrpHdr
{
  // what the FP64 gate needs (rare path): kept here so that the call carries no kernel parameters
  double kc[3];                 // cutoff, cutoff2, neigh_radius
  const double4 *geo;
  const double *hlist;
  int hdr_i[GRP];               // the rows of the group
  int hdr_n[GRP];               // queued pairs per row
};
template <int QC>
struct GrpScratch : GrpHdr
{
  int cand[GCAND];              // candidate row << 1 | own-bucket flag
  int q[GRP][QC];               // rows that passed the gates, per row of the group
};

__device__ __forceinline__ float4 ld_rec32(const float4 *__restrict__ a, int j)
{
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a + j));
  return v;
}

// The reference's gates in FP64 for the candidate cj of this lane against the rows of the group flagged in
// bandb: kernel cutoff |dx_k / h| <= cutoff (sph_lib.cc:216-218), list gate (own bucket, else
// |dx_k| <= 4.25 h_i(list), neighs.cc:84-92,130) and in_support(dx, CUTOFF2 h_i) (sph_header.h:49-60).
// Returns okb with the bits of the rows that fail cleared.  Rare path, kept out of line.
__device__ __noinline__ unsigned gate_exact_grp(const GrpHdr *scr, unsigned okb, unsigned bandb, int cj, int sym)
{
  const double cutoff = scr->kc[0], cutoff2 = scr->kc[1], nradius = scr->kc[2];
  const double4 gj = scr->geo[cj >> 1];
  for (int g = 0; g < GRP; g++) {
    if (!((bandb >> g) & 1u)) continue;
    const int i = scr->hdr_i[g];
    const double4 gi = scr->geo[i];
    const double dx = gi.x - gj.x, dy = gi.y - gj.y, dz = gi.z - gj.z;
    const double h = sym ? 0.5 * (gi.w + gj.w) : gi.w;
    bool ok = !(fabs(dx / h) > cutoff || fabs(dy / h) > cutoff || fabs(dz / h) > cutoff);
    if (ok) {
      const double supp = cutoff2 * gi.w;
      const double lim = (cj & 1) ? supp : fmin(supp, nradius * scr->hlist[i]);
      ok = !(fabs(dx) > lim || fabs(dy) > lim || fabs(dz) > lim);
    }
    if (!ok) okb &= ~(1u << g);
  }
  return okb;
}

// Stage interface: as psph_gather.cuh, plus
//   kWarpsGrp            warps per CTA of the grouped kernel (one CTA per SM)
//   double reach_h(K,h,hmax)  half-width of the candidate box of a row with smoothing length h
//   struct R / load / pair_r  the stage records of row j, loaded apart from the pair body
// PF: 1 = the records of the next pair round are loaded one round ahead (registers).
// LM (pair lists, see k_gather_list): 0 = none; 3 = scan only: when the scan of a group ends, the queue of every
// row that qualifies (queue did not overflow, standard smoothing length in the whole candidate box, list space
// left) is written to the global pair list, the other rows are marked and NO pairs are summed here; 1 = the same,
// but rows that do not qualify are summed in place; 2 = rows that have a list are skipped (the fused fallback);
// 4 = scan only, ANY smoothing lengths: an entry is kept when the pair passes the symmetric-h gates (mom/dens)
// or the h_i gates (velocity filter) and carries one tag bit for each (bit 30 / bit 31 of the row index).
template <class S, int WARPS, int PF, int LM, int QC>
__global__ void __launch_bounds__(WARPS * 32, 1)
k_gather_grp(DevView V, GridView G, KParams K, StepState *ss, double *__restrict__ spill)
{
  extern __shared__ __align__(16) unsigned char s_raw[];
  // LM 2 (rows the list paths left over): they come in clusters — a plume's hot core, rows on subcell faces — so a warp takes
  // 4 consecutive list entries per chunk instead of 32 and the cluster spreads over 8 times as many warps
  constexpr int RPW_ = LM == 2 ? GRP : RPWG;
  constexpr int CHUNK = WARPS * RPW_;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  GrpScratch<QC> *scr = reinterpret_cast<GrpScratch<QC> *>(s_raw) + warp;
  // window addresses of the warp's scratch; the volatile move keeps the compiler from re-deriving them
  // from the thread index at every access
  unsigned s_cand, s_q, s_hdr;
  asm volatile("mov.u32 %0, %1;" : "=r"(s_cand) : "r"((unsigned) __cvta_generic_to_shared(&scr->cand[0])));
  asm volatile("mov.u32 %0, %1;" : "=r"(s_q) : "r"((unsigned) __cvta_generic_to_shared(&scr->q[0][0])));
  asm volatile("mov.u32 %0, %1;" : "=r"(s_hdr) : "r"((unsigned) __cvta_generic_to_shared(&scr->hdr_i[0])));
  if (lane == 0) {
    scr->kc[0] = K.cutoff; scr->kc[1] = K.cutoff2; scr->kc[2] = K.neigh_radius;
    scr->geo = V.geo; scr->hlist = V.t.d[D_HLIST];
  }
  __syncwarp();
  if (LM == 2 && ss->n_uncached == 0) return;
  const int cnt = ss->n_need;
  const int32_t *__restrict__ rows = V.need_idx;
  const double hmax = ss->hmax;
  // FP32 decision band: coordinates are rounded to FP32 relative to the mesh origin (error <= 2^-24 xmax
  // each), the limit is formed in FP32 (a few ulp) -> band of +-(2^-19 b + 2^-21 xmax) around the face
  const float delta = __uint_as_float(ss->xmax_bits) * 0x1p-21f;
  const float hstd_f = (float) hmax;
  const float ch = (float) (S::kSymH ? 0.5 * K.cutoff : K.cutoff);
  const float c_lo = ch * (1.0f - 0x1p-19f), c_hi = ch * (1.0f + 0x1p-19f);
  unsigned wcur = 0, wend = 0;      // LM 1: this warp's block of pair-list space (32-entry units)
  double *__restrict__ my_spill = spill + (size_t) (blockIdx.x * WARPS + warp) * (GRP * GNA * 32) + lane;
  const int nsubx = G.ncell[0] * SUBDIV - 1, nsuby = G.ncell[1] * SUBDIV - 1, nsubz = G.ncell[2] * SUBDIV - 1;

  for (int chunk = blockIdx.x; chunk * CHUNK < cnt; chunk += gridDim.x) {
    int e = chunk * CHUNK + warp * RPW_;
    const int eend = min(e + RPW_, cnt);
    while (e < eend) {
      // ---- form the group: the leading run of list entries in the same subcell column of one bucket
      int row = -1, cell = -1;
      unsigned long long pk = 0;
      if (lane < GRP && e + lane < eend) {
        row = __ldg(&rows[e + lane]);
        pk = V.gsub[row];
        cell = V.t.i[I_CELL][row];
      }
      bool cached = false;
      if (LM == 2) {
        cached = row >= 0 && V.nbr_cnt[row] >= 0;
        const unsigned cm = __ballot_sync(FULL, cached);
        if (cm & 1u) { e += __ffs(~cm) - 1; continue; } // leading rows with a cached list: done elsewhere
      }
      const unsigned long long pk0 = __shfl_sync(FULL, pk, 0);
      const int cell0 = __shfl_sync(FULL, cell, 0);
      const bool same = row >= 0 && !cached && cell == cell0 && ((pk ^ pk0) & 0xffffffffffull) == 0;
      const int m = __ffs(~__ballot_sync(FULL, same)) - 1; // 1..GRP
      const bool mine = lane < m;
      float4 p32 = make_float4(0.f, 0.f, 0.f, 0.f);
      int kR = 0, sz = 0;
      bool nl = false;
      if (mine) {
        p32 = V.geo32[row];
        const double hi = V.t.d[D_H][row];
        const double supp = K.cutoff2 * hi, rlist = K.neigh_radius * V.t.d[D_HLIST][row];
        kR = (int) ceil(S::reach_h(K, hi, hmax) * G.inv_subw - 1e-9) + (int) (pk >> 60);
        sz = (int) ((pk >> 40) & 0xfffff);
        // the list / in_support limits can only bite when the largest possible cutoff box exceeds them
        nl = K.cutoff * (S::kSymH ? 0.5 * (hi + hmax) : hi) > fmin(supp, rlist);
        sts32(s_hdr + 4u * lane, row);
      }
      const int kRg = __reduce_max_sync(FULL, kR);
      const int zlo = __reduce_min_sync(FULL, mine ? sz : 0x7fffffff), zhi = __reduce_max_sync(FULL, mine ? sz : -1);
      const unsigned nlmask = __ballot_sync(FULL, nl);
      const unsigned stdmask = (LM == 1 || LM == 3) ? __ballot_sync(FULL, mine && p32.w == hstd_f) : 0u;
      bool nonstd = false;            // LM 1: a candidate of the box has another smoothing length
      bool guni = false;
      unsigned ovf = 0;               // LM 3: rows whose queue overflowed
      int ig[GRP];
      unsigned qa[GRP];               // window address of the next free slot of every queue
      float fx[GRP], fy[GRP], fz[GRP], blo[GRP], bhi[GRP];
      float vlo[GRP], vhi[GRP];       // LM 4: the FP32 band of the h_i gate |dx_k| <= cutoff h_i
#pragma unroll
      for (int g = 0; g < GRP; g++) {
        ig[g] = __shfl_sync(FULL, row, g);
        fx[g] = __shfl_sync(FULL, p32.x, g); fy[g] = __shfl_sync(FULL, p32.y, g); fz[g] = __shfl_sync(FULL, p32.z, g);
        const float hf = __shfl_sync(FULL, p32.w, g);
        // a row whose list / in_support limits can bite sends every candidate through the FP64 gates
        blo[g] = ((nlmask >> g) & 1u) ? -INFINITY : hf * c_lo - delta;
        bhi[g] = hf * c_hi + delta;
        if (LM == 4) {
          const float cv = (float) K.cutoff;
          vlo[g] = ((nlmask >> g) & 1u) ? -INFINITY : hf * (cv * (1.0f - 0x1p-19f)) - delta;
          vhi[g] = hf * (cv * (1.0f + 0x1p-19f)) + delta;
        }
        qa[g] = s_q + 4u * QC * g;
      }
      // ---- candidate enumeration state (the group's common box, in global subcell coordinates)
      const int sx = (int) (pk0 & 0xfffff), sy = (int) ((pk0 >> 20) & 0xfffff);
      const int gx0 = max(sx - kRg, 0), gy0 = max(sy - kRg, 0), gz0 = max(zlo - kRg, 0);
      const int gx1 = min(sx + kRg, nsubx), gy1 = min(sy + kRg, nsuby), gz1 = min(zhi + kRg, nsubz);
      const int ncoly = gy1 - gy0 + 1, ncol = (gx1 - gx0 + 1) * ncoly;
      const float inv_ncoly = 1.0f / (float) ncoly;
      const int cz0 = gz0 / SUBDIV, cz1 = gz1 / SUBDIV;
      int col_base = -32, cz = cz1, c0 = 0, total = 0, beg = 0, len = 0, cellxy = 0, subxy = 0;
      bool have = false, own = false;
      unsigned drained = 0;

      // one candidate of the dense list (p.x = +inf in lanes without one) against every row of the group
      auto test = [&](int cj, const float4 &p) {
        const int j = cj >> 1;
        const float hlo = S::kSymH ? p.w * c_lo : 0.f, hhi = S::kSymH ? p.w * c_hi : 0.f;
        if (LM == 1 || LM == 3) nonstd |= p.w != hstd_f;
        unsigned okb = 0, bandb = 0, okv = 0;
#pragma unroll
        for (int g = 0; g < GRP; g++) {
          if (g < m) {
            const float m3 = fmaxf(fmaxf(fabsf(fx[g] - p.x), fabsf(fy[g] - p.y)), fabsf(fz[g] - p.z));
            bool ok = m3 <= bhi[g] + hhi;
            if (S::kSkipSelf) ok = ok && j != ig[g];
            if (ok) okb |= 1u << g;
            if (ok && m3 > blo[g] + hlo) bandb |= 1u << g;
            if (LM == 4) { // the second gate family (velocity filter): |dx_k| <= cutoff h_i
              const bool v = m3 <= vhi[g] && j != ig[g];
              if (v) okv |= 1u << g;
              if (v && m3 > vlo[g]) bandb |= 1u << g;
            }
          }
        }
        if (__any_sync(FULL, bandb != 0)) { // within the FP32 error band of a cutoff face: the reference's own test
          if (bandb) {
            if (LM == 4) { // both families exactly for the flagged rows
              okv = (okv & ~bandb) | gate_exact_grp(scr, bandb, bandb, cj, 0);
              okb = (okb & ~bandb) | gate_exact_grp(scr, bandb, bandb, cj, 1);
            } else okb = gate_exact_grp(scr, okb, bandb, cj, S::kSymH ? 1 : 0);
          }
        }
#pragma unroll
        for (int g = 0; g < GRP; g++) {
          if (g < m) {
            const bool ok = ((okb | okv) >> g) & 1u;
            const unsigned b = __ballot_sync(FULL, ok);
            if (ok) sts32(qa[g] + 4u * __popc(b & lt), LM == 4 ? (int) ((unsigned) j | (((okb >> g) & 1u) << 30) | (((okv >> g) & 1u) << 31)) : j);
            qa[g] += 4u * __popc(b);
          }
        }
      };

      auto fetch = [&](int c, int &ja, int &jb, float4 &pa, float4 &pb) {
        const int ca = c + lane, cb = ca + 32;
        const bool la = ca < total, lb = cb < total;
        ja = la ? lds32(s_cand + 4u * ca) : 0; jb = lb ? lds32(s_cand + 4u * cb) : 0;
        pa = ld_rec32(V.geo32, ja >> 1); pb = ld_rec32(V.geo32, jb >> 1);
        if (!la) { pa.x = INFINITY; if (LM == 1 || LM == 3) pa.w = hstd_f; }
        if (!lb) { pb.x = INFINITY; if (LM == 1 || LM == 3) pb.w = hstd_f; }
      };

      for (bool done = false; !done;) {
        // ---- scan until the candidates are exhausted or a queue may overflow in the next iteration
        int gsel = -1;
        for (;;) {
          if (c0 >= total) {
            // refill: expand row ranges into the candidate list until it is full or nothing is left
            if (total) { total = 0; c0 = 0; __syncwarp(); }
            for (;;) {
              if (__any_sync(FULL, len > 0)) {
                const int quota = (GCAND - total) / 32;
                if (quota == 0) break;
                const int take = min(len, quota);
                int off = take;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, off, o); if (lane >= o) off += v; }
                const int sum = __shfl_sync(FULL, off, 31);
                off += total - take;
                const int tag = own ? 1 : 0;
                for (int k = 0; k < take; k++) sts32(s_cand + 4u * (off + k), ((beg + k) << 1) | tag);
                beg += take; len -= take; total += sum;
                continue;
              }
              if (cz < cz1) { // the z-run of every column of this round inside the next bucket along z
                cz++;
                beg = 0; len = 0; own = false;
                if (have) {
                  const int cc = cellxy + cz;
                  const int zl = (cz == cz0) ? (gz0 % SUBDIV) : 0, zh = (cz == cz1) ? (gz1 % SUBDIV) : SUBDIV - 1;
                  const int t = __ldg(&G.cell2rank[cc]) * SUBS + subxy;
                  beg = __ldg(&G.sub_start[t + zl]);
                  len = __ldg(&G.sub_start[t + zh + 1]) - beg;
                  own = cc == cell0;
                }
                continue;
              }
              col_base += 32; // next round of 32 (x, y) subcell columns
              if (col_base >= ncol) break;
              const int col = col_base + lane;
              have = col < ncol;
              if (have) {
                const int ix = (int) (((float) col + 0.5f) * inv_ncoly), iy = col - ix * ncoly;
                const int gx = gx0 + ix, gy = gy0 + iy;
                cellxy = ((gx / SUBDIV) * G.ncell[1] + (gy / SUBDIV)) * G.ncell[2];
                subxy = (gx % SUBDIV) * (SUBDIV * SUBDIV) + (gy % SUBDIV) * SUBDIV;
              }
              cz = cz0 - 1;
            }
            __syncwarp();
            if (total == 0) break;
          }
#pragma unroll
          for (int g = 0; g < GRP; g++)
            if (qa[g] > s_q + 4u * (QC * g + QC - 64)) gsel = g;
          if (gsel >= 0) break;
          // two candidates per lane, both loads in flight
          int ja, jb;
          float4 pa, pb;
          fetch(c0, ja, jb, pa, pb);
          const bool two = c0 + 32 < total;
          c0 += 64;
          test(ja, pa);
          if (two) test(jb, pb);
        }
        if ((LM == 3 || LM == 4) && gsel >= 0) { // no pair phase here: the row's list is dropped, the fused kernel takes the row
#pragma unroll
          for (int g = 0; g < GRP; g++)
            if (g == gsel) { ovf |= 1u << g; qa[g] = s_q + 4u * QC * g; bhi[g] = -INFINITY; }
          continue;
        }
        done = gsel < 0;
        if (lane == 0) {
#pragma unroll
          for (int g = 0; g < GRP; g++) sts32(s_hdr + 16u + 4u * g, (int) (qa[g] - (s_q + 4u * QC * g)) >> 2);
        }
        __syncwarp();

        // ---- pair phase: every row of the group (scan finished) or the one whose queue is full
        const int g0 = done ? 0 : gsel, g1 = done ? m : gsel + 1;
        for (int g = g0; g < g1; g++) {
          const int i = lds32(s_hdr + 4u * g), n = lds32(s_hdr + 16u + 4u * g);
          if ((LM == 1 || LM == 3 || LM == 4) && done) {
            // publish the row's pair list and leave its pairs to k_gather_list when the row qualifies: queue
            // complete, its own h the standard one and so that of every candidate of the group's box
            // (geo32.w == (float) hmax  <=>  h == hmax, see dev_store_geo32), list space left
            if (g == g0) guni = LM == 4 || !__any_sync(FULL, nonstd);
            int off = -1;
            if (guni && !(((drained | ovf) >> g) & 1u) && (LM == 4 || ((stdmask >> g) & 1u))) {
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
              V.nbr_off[i] = off;
              V.nbr_cnt[i] = off >= 0 ? n : -1;
              if (off < 0) atomicAdd(&ss->n_uncached, 1);
            }
            if (off >= 0) {
              int32_t *__restrict__ dst = V.nbr + (size_t) (unsigned) off * 32u;
              const unsigned qb = s_q + 4u * QC * g;
              for (int k = lane; k < n; k += 32) dst[k] = lds32(qb + 4u * k);
              continue;
            }
            if (LM == 3 || LM == 4) continue;
          }
          if (LM == 3 || LM == 4) continue; // (never reached: the scan-only kernels have no pair phase)
          typename S::I st;
          S::init(V, K, ss, i, st);
          const double4 gi = V.geo[i];
          typename S::A acc;
          S::begin(K, st, acc, 0);
          double *ap = reinterpret_cast<double *>(&acc);
          if ((drained >> g) & 1u) {
#pragma unroll
            for (int k = 0; k < GNA; k++) ap[k] = my_spill[(g * GNA + k) * 32];
          }
          const unsigned qg = s_q + 4u * QC * g + 4u * lane;
          const int nfull = done ? n : (n & ~31);
          const int nrun = nfull - lane;                    // entries left for this lane's slot
          // software pipeline: the row index of the next round is read one round ahead; PF: its records are
          // loaded one round ahead as well (registers), so that a round never waits for its own loads
          int jn = 0 < nrun ? lds32(qg) : -1;
          if (PF == 1) {
            double4 gn = ld_rec(V.geo, jn < 0 ? i : jn);
            typename S::R rn = S::load(V, jn < 0 ? i : jn);
            for (int q0 = 0; q0 < nfull; q0 += 32) {
              const int j = jn;
              const double4 gj = gn;
              const typename S::R rj = rn;
              jn = q0 + 32 < nrun ? lds32(qg + 4u * (q0 + 32)) : -1;
              if (q0 + 32 < nfull) { gn = ld_rec(V.geo, jn < 0 ? i : jn); rn = S::load(V, jn < 0 ? i : jn); }
              if (j >= 0) S::template pair_r<false>(K, st, acc, gi.x - gj.x, gi.y - gj.y, gi.z - gj.z, gj.w, rj);
            }
          } else {
            for (int q0 = 0; q0 < nfull; q0 += 32) {
              const int j = jn;
              jn = q0 + 32 < nrun ? lds32(qg + 4u * (q0 + 32)) : -1;
              if (j >= 0) {
                const double4 gj = ld_rec(V.geo, j);
                S::pair(V, K, st, acc, gi.x - gj.x, gi.y - gj.y, gi.z - gj.z, gj.w, j);
              }
            }
          }
          if (done) {
            S::end(V, K, ss, i, st, acc, 0, lane);
          } else { // park the partial sums, keep the incomplete round at the head of the queue
#pragma unroll
            for (int k = 0; k < GNA; k++) my_spill[(g * GNA + k) * 32] = ap[k];
            drained |= 1u << g;
            const int rem = n - nfull;
            int v = 0;
            if (lane < rem) v = lds32(qg + 4u * nfull);
            __syncwarp();
            if (lane < rem) sts32(qg, v);
            __syncwarp();
#pragma unroll
            for (int gg = 0; gg < GRP; gg++) if (gg == g) qa[gg] = s_q + 4u * (QC * gg + rem);
          }
        }
      }
      e += m;
      __syncwarp();
    }
  }
}

template <class S, int WARPS, int PF, int LM>
int launch_gather_grp_w(psph_ctx *ctx, int kid)
{
  static_assert(sizeof(typename S::A) == GNA * sizeof(double), "grouped gather: four accumulators per stage");
  constexpr int QC = LM == 4 ? GQCAP_TAG : (LM == 3 ? GQCAP_SCAN : GQCAP);
  constexpr int threads = WARPS * 32;
  constexpr size_t smem = sizeof(GrpScratch<QC>) * WARPS;
  // per device, and a process may drive several devices (one context each): set it at every launch
  PSPH_CUDA(cudaFuncSetAttribute(k_gather_grp<S, WARPS, PF, LM, QC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  const size_t spill_rows = (size_t) ctx->sms * 32 * GRP * GNA * 32; // up to 32 warps per CTA, one CTA per SM
  if (!ctx->d_gspill) PSPH_CUDA(cudaMalloc(&ctx->d_gspill, spill_rows * sizeof(double)));
  KernelTimer kt(ctx, kid);
  constexpr int RPW_ = LM == 2 ? GRP : RPWG;
  const int nchunks = (ctx->n + WARPS * RPW_ - 1) / (WARPS * RPW_);
  const int grid = min(max(nchunks, 1), ctx->sms);
  k_gather_grp<S, WARPS, PF, LM, QC><<<grid, threads, smem, ctx->stream>>>(ctx->cur, psph_grid_view(ctx), ctx->kp, ctx->d_ss, ctx->d_gspill);
  PSPH_LAUNCH_CHECK();
  return 0;
}
// lm: 0 plain, 3 find the pairs and write the lists only, 2 only rows without a list.  Warps per CTA and record prefetch:
// the stage's defaults, or PSPH_GWARPS / PSPH_GPF (tuning switches; list modes always run the defaults)
template <class S>
int launch_gather_grp(psph_ctx *ctx, int kid, int lm)
{
  if (lm == 4) return launch_gather_grp_w<S, 12, 0, 4>(ctx, kid);
  if (lm == 3) {
    if (ctx->gather_warps == 16) return launch_gather_grp_w<S, 16, 0, 3>(ctx, kid);
    if (ctx->gather_warps == 24) return launch_gather_grp_w<S, 24, 0, 3>(ctx, kid);
    return launch_gather_grp_w<S, 20, 0, 3>(ctx, kid);
  }
  if (lm == 2) return launch_gather_grp_w<S, S::kWarpsGrp, S::kPfGrp, 2>(ctx, kid);
  const int w = (ctx->gather_warps ? ctx->gather_warps : S::kWarpsGrp) + 100 * (ctx->gather_pf >= 0 ? ctx->gather_pf : S::kPfGrp);
  switch (w) {
  case 16: return launch_gather_grp_w<S, 16, 0, 0>(ctx, kid);
  case 20: return launch_gather_grp_w<S, 20, 0, 0>(ctx, kid);
  case 24: return launch_gather_grp_w<S, 24, 0, 0>(ctx, kid);
  case 116: return launch_gather_grp_w<S, 16, 1, 0>(ctx, kid);
  case 120: return launch_gather_grp_w<S, 20, 1, 0>(ctx, kid);
  case 124: return launch_gather_grp_w<S, 24, 1, 0>(ctx, kid);
  default: ctx->err = "PSPH_GWARPS / PSPH_GPF: no such grouped-gather variant"; return -2;
  }
}

// =========================================================================================
// Pair-list gather.  Between mom_engr_update, smooth_density and smooth_velocity of one time step neither
// positions nor smoothing lengths nor the row order change, so the three stages sum over the SAME pairs
// whenever every partner j of row i has h_j == h_i (then 0.5 (h_i + h_j) == h_i bit for bit and the
// symmetric-h gates of mom/dens equal the h_i gates of the velocity filter; the self pair, which only
// mom_engr_update skips, is not in the list and is added by the stages that count it).  The first stage
// writes those rows' queues to a global list (LM 1 above); the later stages run one warp per row straight
// off the list - no enumeration, no gates, no shared memory - and rows without a list fall back to the
// grouped gather (LM 2).  The list lives for one time step and only when memory allows.
// =========================================================================================
constexpr int RPWL = 16;       // consecutive list entries per warp and chunk

// TAG: general lists (scan LM 4): entries carry the gate-family bits, partners have their own smoothing lengths
template <class S, int WARPS, int PF, bool TAG>
__global__ void __launch_bounds__(WARPS * 32, 1) k_gather_list(DevView V, KParams K, StepState *ss)
{
  constexpr int CHUNK = WARPS * RPWL;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cnt = ss->n_need;
  const int32_t *__restrict__ rows = V.need_idx;
  for (int chunk = blockIdx.x; chunk * CHUNK < cnt; chunk += gridDim.x) {
    const int e0 = chunk * CHUNK + warp * RPWL;
    for (int r = 0; r < RPWL; r++) {
      if (e0 + r >= cnt) break;
      const int i = __ldg(&rows[e0 + r]);
      const int n = V.nbr_cnt[i];
      if (n < 0) continue;                       // no cached list: the grouped gather takes the row
      const int32_t *__restrict__ lst = V.nbr + (size_t) (unsigned) V.nbr_off[i] * 32u + lane;
      typename S::I st;
      S::init(V, K, ss, i, st);
      if (!TAG) S::init_uni(K, st);              // every partner of a listed row has the row's own h
      const double4 gi = V.geo[i];
      typename S::A acc;
      S::begin(K, st, acc, 0);
      // entry k of the row: list[k] for k < n, then (stages that count it) the self pair, which the list omits
      const int ntot = n + (S::kSkipSelf ? 0 : 1);
      // TAG: -1 unless the entry carries this stage's family bit (30: symmetric-h gates, 31: h_i gates)
      auto entry = [&](int k) {
        if (k >= n) return k < ntot ? i : -1;
        const int raw = __ldg(lst + (k - lane));
        if (!TAG) return raw;
        return (((unsigned) raw >> (S::kSymH ? 30 : 31)) & 1u) ? (raw & 0x3fffffff) : -1;
      };
      int jn = entry(lane);
      if (PF == 1) {
        double4 gn = ld_rec(V.geo, jn < 0 ? i : jn);
        typename S::R rn = S::load(V, jn < 0 ? i : jn);
        for (int q0 = 0; q0 < ntot; q0 += 32) {
          const int j = jn;
          const double4 gj = gn;
          const typename S::R rj = rn;
          jn = entry(q0 + 32 + lane);
          if (q0 + 32 < ntot) { gn = ld_rec(V.geo, jn < 0 ? i : jn); rn = S::load(V, jn < 0 ? i : jn); }
          if (j >= 0) S::template pair_r<!TAG>(K, st, acc, gi.x - gj.x, gi.y - gj.y, gi.z - gj.z, gj.w, rj);
        }
      } else {
        for (int q0 = 0; q0 < ntot; q0 += 32) {
          const int j = jn;
          jn = entry(q0 + 32 + lane);
          if (j >= 0) {
            const double4 gj = ld_rec(V.geo, j);
            S::template pair_r<!TAG>(K, st, acc, gi.x - gj.x, gi.y - gj.y, gi.z - gj.z, gj.w, S::load(V, j));
          }
        }
      }
      S::end(V, K, ss, i, st, acc, 0, lane);
    }
  }
}

template <class S, int WARPS, int PF, bool TAG = false>
int launch_gather_list_w(psph_ctx *ctx, int kid)
{
  KernelTimer kt(ctx, kid);
  const int nchunks = (ctx->n + WARPS * RPWL - 1) / (WARPS * RPWL);
  const int grid = min(max(nchunks, 1), ctx->sms);
  k_gather_list<S, WARPS, PF, TAG><<<grid, WARPS * 32, 0, ctx->stream>>>(ctx->cur, ctx->kp, ctx->d_ss);
  PSPH_LAUNCH_CHECK();
  return 0;
}
template <class S>
int launch_gather_list(psph_ctx *ctx, int kid)
{
  if (ctx->nbr_kind == 3) return launch_gather_list_w<S, S::kWarpsList, 0, true>(ctx, kid);
  const int w = (ctx->list_warps ? ctx->list_warps : S::kWarpsList) + 100 * (ctx->gather_pf >= 0 ? ctx->gather_pf : S::kPfList);
  switch (w) {
  case 16: return launch_gather_list_w<S, 16, 0>(ctx, kid);
  case 20: return launch_gather_list_w<S, 20, 0>(ctx, kid);
  case 24: return launch_gather_list_w<S, 24, 0>(ctx, kid);
  case 32: return launch_gather_list_w<S, 32, 0>(ctx, kid);
  case 116: return launch_gather_list_w<S, 16, 1>(ctx, kid);
  case 120: return launch_gather_list_w<S, 20, 1>(ctx, kid);
  case 124: return launch_gather_list_w<S, 24, 1>(ctx, kid);
  case 132: return launch_gather_list_w<S, 32, 1>(ctx, kid);
  default: ctx->err = "PSPH_LWARPS / PSPH_GPF: no such list-gather variant"; return -2;
  }
}
