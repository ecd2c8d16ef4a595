// psph_gather.cuh — the pair gather that every SPH summation stage runs on.
//
// One WARP per particle i.  The candidates of i are the rows of the subcells (h_air cubes, 512
// per bucket) that overlap the box x_i +- reach; for every (x, y) subcell column and every
// bucket the column crosses, the z-run of subcells is ONE contiguous row range ("segment") of
// the sorted table.  A warp works in two alternating phases:
//   scan : eight segments at a time, four lanes per segment, two rows per lane (both loads in
//          flight together): every lane loads the 32-byte {x,y,z,h} record of a candidate row with
//          a single 256-bit load and applies the reference's own gates, bit for bit:
//            list gate   : j in i's bucket, or |dx_k| <= 4.25*h_i(list) for all k   (neighs.cc:84-92,130)
//            stage gates : in_support(dx, CUTOFF2*h_i), kernel cutoff |dx_k/h| <= 3
//                                                            (sph_header.h:49-60, sph_lib.cc:216-218)
//          (rows the stage's contr_* predicate excludes carry zeros in the stage record instead)
//          candidates that pass are appended (ballot + popc) to a per-warp queue in shared memory;
//   pair : whenever the queue holds >= 32 entries, all 32 lanes take one entry each and run the
//          stage's pair body — the FP64-heavy part runs with every lane busy.
// The set of pairs that contribute is exactly the reference's; pairs that cannot contribute
// (zero kernel weight) are never queued.  Lane partial sums are combined with warp shuffles.
// CTAs are persistent (one resident wave) and walk chunks of 128 consecutive entries of the list
// of rows the stage updates; rows are sorted by subcell, so the eight warps of a CTA work on one
// compact block of space at a time and share its candidate rows through L1.  No CTA-wide barrier.
#pragma once
#include "psph_internal.cuh"

constexpr int GW = 8;          // warps per CTA
constexpr int GT = GW * 32;    // threads per CTA
constexpr int RPW = 16;        // consecutive list entries per warp and chunk
constexpr int CHUNK = GW * RPW;// entries per CTA chunk (spatially compact: its candidates share L1)
constexpr int QCAP = 128;      // queue slots per warp (power of two, >= 31 + 64)
constexpr int SEGCAP = 128;    // segments resolved per batch
constexpr int LPS = 4;         // lanes per segment in the scan phase (two rows per lane and iteration)
constexpr unsigned FULL = 0xffffffffu;

struct WarpScratch
{
  int seg_beg[SEGCAP];         // first row | own-bucket flag in bit 31
  int seg_end[SEGCAP];
  int qj[QCAP];
  double qdx[QCAP], qdy[QCAP], qdz[QCAP], qh[QCAP];
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v)
{
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// one 32-byte record with a single 256-bit load (LDG.E.256 on sm_100a), read-only path
__device__ __forceinline__ double4 ld_rec(const double4 *__restrict__ a, int j)
{
  double4 v;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(a + j));
  return v;
}

// fabs(dx / h) > cutoff, decided exactly as the reference's division would (sph_lib.cc:216-218)
// but without dividing unless the pair sits within a few ulp of the cutoff face
__device__ __forceinline__ bool beyond_cutoff(double dx, double h, double cutoff)
{
  double a = fabs(dx), b = cutoff * h;
  if (a > b * (1.0 + 0x1p-48)) return true;
  if (a < b * (1.0 - 0x1p-48)) return false;
  return fabs(dx / h) > cutoff;
}
template <int DIM>
__device__ __forceinline__ bool outside_kernel(double dx, double dy, double dz, double h, double cutoff)
{
  if (beyond_cutoff(dx, h, cutoff)) return true;
  if (DIM == 3) return beyond_cutoff(dy, h, cutoff) || beyond_cutoff(dz, h, cutoff);
  return false;
}
template <int DIM>
__device__ __forceinline__ bool dev_in_support(double dx, double dy, double dz, double supp)
{
  if (fabs(dx) > supp) return false;
  if (DIM == 3 && (fabs(dy) > supp || fabs(dz) > supp)) return false;
  return true;
}
// SPH kernel (sph_lib.cc:207-230) for a pair already known to be inside the cutoff:
// w = (pi^-1/2 / h)^DIM * exp(-|dx|^2/h^2)
template <int DIM>
__device__ __forceinline__ double dev_weight_in(double r2, double invh)
{
  const double t = 0.564189583547756 * invh;
  const double norm = DIM == 3 ? t * t * t : t;
  return norm * exp(-(r2 * invh * invh));
}

// Stage interface (static members):
//   kSymH         pair smoothing length is 0.5*(h_i+h_j) (true) or S::I::hk (false)
//   kSkipSelf     j == i is not a pair (mom_engr_update.cc:201)
//   kRows         RowFlags bit of the rows the stage updates (selects the row list it walks)
//   struct I      per-row state, identical in every lane; members h, supp (= CUTOFF2*h_i), hk
//   struct A      per-lane accumulators
//   void  init(V,K,ss,i,I&)             load row i
//   double reach(K,I,hmax)              half-width of the candidate box
//   void  begin(K,I&,A&,pass)           zero the accumulators
//   void  pair(V,K,I,A&,dx,dy,dz,hj,j)  one pair inside the kernel support
//   bool  end(V,K,ss,i,I&,A&,pass,lane) warp-reduce + write; true = run another pass
// Rows that must not contribute to a stage's sums (contr_dens / contr_vel_smooth false) carry
// zeros in the stage record, so the scan phase needs no flag word.
template <int DIM, class S>
__global__ void __launch_bounds__(GT) k_gather(DevView V, GridView G, KParams K, StepState *ss)
{
  __shared__ WarpScratch s_scr[GW];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  WarpScratch &W = s_scr[warp];
  const int n = ss->n;
  const int cnt = S::kRows == F_NEED ? ss->n_need : ss->n_grad;
  const int32_t *__restrict__ rows = S::kRows == F_NEED ? V.need_idx : V.grad_idx;
  const double hmax = ss->hmax;
  for (int chunk = blockIdx.x; chunk * CHUNK < cnt; chunk += gridDim.x) {
    const int e0 = chunk * CHUNK + warp * RPW;
    for (int r = 0; r < RPW; r++) {
      if (e0 + r >= cnt) break;
      const int i = __ldg(&rows[e0 + r]);
      typename S::I st;
      S::init(V, K, ss, i, st);
      const double4 gi = V.geo[i];
      const double rlist = K.neigh_radius * V.t.d[D_HLIST][i];
      typename S::A acc;
      int pass = 0;
      bool again = true;
      while (again) {
        S::begin(K, st, acc, pass);
        const double lim_own = st.supp, lim_other = fmin(st.supp, rlist);
        int qn = 0, qhead = 0; // queue state, identical in every lane

        // ---- scan phase, one candidate per lane: list gate (neighs.cc:88-92 own bucket, :130 others),
        // in_support (sph_header.h:49-60) and kernel cutoff (sph_lib.cc:216-218) on t = max_k |dx_k|
        auto consider = [&](bool live, bool own, int j, const double4 &gj) {
          const double dx = gi.x - gj.x, dy = gi.y - gj.y, dz = gi.z - gj.z;
          const double t = DIM == 3 ? fmax(fmax(fabs(dx), fabs(dy)), fabs(dz)) : fabs(dx);
          bool ok = live && !(t > (own ? lim_own : lim_other)) && !(S::kSkipSelf && j == i);
          if (ok) {
            const double h = S::kSymH ? 0.5 * (st.h + gj.w) : st.hk;
            const double b = K.cutoff * h;
            if (t > b * (1.0 - 0x1p-48)) ok = !(t > b * (1.0 + 0x1p-48)) && !(fabs(t / h) > K.cutoff);
          }
          const unsigned m = __ballot_sync(FULL, ok);
          if (ok) {
            const int s = (qhead + qn + __popc(m & lt)) & (QCAP - 1);
            W.qj[s] = j; W.qdx[s] = dx; W.qdy[s] = dy; W.qdz[s] = dz; W.qh[s] = gj.w;
          }
          qn += __popc(m);
        };

        // ---- candidate enumeration state
        int j0 = 0;                                  // 1-D: next row
        int gx0 = 0, gy0 = 0, gz0 = 0, gz1 = 0, ncolx = 0, ncoly = 0, ncol = 0, cz0 = 0, cz1 = 0, cpr = 32;
        float inv_ncoly = 1.0f;
        int col_base = 0, ns = 0, s0 = 0, j = 0, e = 0;
        bool own = false;
        int cell_i = 0;
        if (DIM == 3) {
          cell_i = V.t.i[I_CELL][i];
          const unsigned long long pk = V.gsub[i];
          const int sx = (int) (pk & 0xfffff), sy = (int) ((pk >> 20) & 0xfffff), sz = (int) ((pk >> 40) & 0xfffff);
          const int kR = (int) ceil(S::reach(K, st, hmax) * G.inv_subw - 1e-9) + (int) (pk >> 60);
          gx0 = max(sx - kR, 0); gy0 = max(sy - kR, 0); gz0 = max(sz - kR, 0);
          const int gx1 = min(sx + kR, G.ncell[0] * SUBDIV - 1), gy1 = min(sy + kR, G.ncell[1] * SUBDIV - 1);
          gz1 = min(sz + kR, G.ncell[2] * SUBDIV - 1);
          ncolx = gx1 - gx0 + 1; ncoly = gy1 - gy0 + 1; ncol = ncolx * ncoly;
          inv_ncoly = 1.0f / (float) ncoly;
          cz0 = gz0 / SUBDIV; cz1 = gz1 / SUBDIV;
          cpr = min(32, SEGCAP / (cz1 - cz0 + 1));
        }
        // one scan iteration (up to 64 candidates); false when no candidate is left
        auto scan_next = [&]() -> bool {
          if (DIM == 1) {
            // the 1-D reference has no search: every particle lists every particle (add_air.cc:470-479)
            if (j0 >= n) return false;
            const int jj = min(j0 + lane, n - 1);
            consider(j0 + lane < n, true, jj, ld_rec(V.geo, jj));
            j0 += 32;
            return true;
          }
          for (;;) {
            if (__any_sync(FULL, j < e)) {
              const bool l0 = j < e, l1 = j + LPS < e;
              const double4 g0 = ld_rec(V.geo, l0 ? j : i), g1 = ld_rec(V.geo, l1 ? j + LPS : i);
              consider(l0, own, j, g0);
              if (__any_sync(FULL, l1)) consider(l1, own, j + LPS, g1);
              j += 2 * LPS;
              return true;
            }
            if (s0 < ns) { // next group of 32/LPS segments
              const int sidx = s0 + lane / LPS;
              int b = 0;
              e = 0;
              if (sidx < ns) { b = W.seg_beg[sidx]; e = W.seg_end[sidx]; }
              own = b < 0; b &= 0x7fffffff;
              j = b + lane % LPS;
              s0 += 32 / LPS;
              continue;
            }
            if (col_base >= ncol) return false;
            // resolve the z-runs of the next cpr (x, y) subcell columns to row ranges; keep the non-empty ones
            __syncwarp();
            const int col = col_base + lane;
            const bool have = lane < cpr && col < ncol;
            int cellxy = 0, subxy = 0;
            if (have) {
              const int ix = (int) (((float) col + 0.5f) * inv_ncoly), iy = col - ix * ncoly;
              const int gx = gx0 + ix, gy = gy0 + iy;
              cellxy = ((gx / SUBDIV) * G.ncell[1] + (gy / SUBDIV)) * G.ncell[2];
              subxy = (gx % SUBDIV) * (SUBDIV * SUBDIV) + (gy % SUBDIV) * SUBDIV;
            }
            ns = 0;
            for (int cz = cz0; cz <= cz1; cz++) {
              int beg = 0, end = 0, o = 0;
              if (have) {
                const int cell = cellxy + cz;
                const int zlo = (cz == cz0) ? (gz0 % SUBDIV) : 0, zhi = (cz == cz1) ? (gz1 % SUBDIV) : SUBDIV - 1;
                const int t = __ldg(&G.cell2rank[cell]) * SUBS + subxy;
                beg = __ldg(&G.sub_start[t + zlo]);
                end = __ldg(&G.sub_start[t + zhi + 1]);
                o = cell == cell_i;
              }
              const bool ne = end > beg;
              const unsigned m = __ballot_sync(FULL, ne);
              if (ne) {
                const int p = ns + __popc(m & lt);
                W.seg_beg[p] = beg | (o << 31);
                W.seg_end[p] = end;
              }
              ns += __popc(m);
            }
            __syncwarp();
            col_base += cpr;
            s0 = 0; j = 0; e = 0;
          }
        };

        // ---- alternate: pair rounds while the queue holds a full warp's worth, else scan on
        for (bool flush = false;;) {
          while (qn >= 32 || (flush && qn > 0)) {
            __syncwarp();
            const int c = min(32, qn);
            if (lane < c) {
              const int s = (qhead + lane) & (QCAP - 1);
              S::pair(V, K, st, acc, W.qdx[s], W.qdy[s], W.qdz[s], W.qh[s], W.qj[s]);
            }
            qhead = (qhead + c) & (QCAP - 1);
            qn -= c;
            __syncwarp();
          }
          if (flush) break;
          if (!scan_next()) flush = true;
        }
        again = S::end(V, K, ss, i, st, acc, pass, lane);
        pass++;
      }
    }
  }
}

template <int DIM, class S>
int launch_gather_d(psph_ctx *ctx, int kid)
{
  static int per_sm = 0; // resident CTAs per SM of this instantiation: the grid is one persistent wave
  if (per_sm == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gather<DIM, S>, GT, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  }
  KernelTimer kt(ctx, kid);
  int nchunks = (ctx->n + CHUNK - 1) / CHUNK;
  int grid = min(max(nchunks, 1), ctx->sms * per_sm);
  k_gather<DIM, S><<<grid, GT, 0, ctx->stream>>>(ctx->cur, psph_grid_view(ctx), ctx->kp, ctx->d_ss);
  PSPH_LAUNCH_CHECK();
  return 0;
}
template <template <int> class S>
int launch_gather(psph_ctx *ctx, int kid)
{
  if (ctx->cfg.dim == 3) return launch_gather_d<3, S<3> >(ctx, kid);
  return launch_gather_d<1, S<1> >(ctx, kid);
}
