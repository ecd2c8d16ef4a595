// psph_gather.cuh — the pair gather that every SPH summation stage runs on.
//
// One WARP per particle i.  The candidates of i are the rows of the subcells (h_air cubes, 512
// per bucket) that overlap the box x_i +- reach; for every (x, y) subcell column and every
// bucket the column crosses, the z-run of subcells is ONE contiguous row range ("segment") of
// the sorted table.  A warp works in two alternating phases:
//   scan : the row ranges are expanded (warp prefix sum) into a dense candidate list in shared
//          memory; every lane takes two candidates at a time (both loads in flight together),
//          loads the 32-byte {x,y,z,h} record of the row with a single 256-bit load and applies
//          the reference's own gates, bit for bit:
//            list gate   : j in i's bucket, or |dx_k| <= 4.25*h_i(list) for all k   (neighs.cc:84-92,130)
//            stage gates : in_support(dx, CUTOFF2*h_i), kernel cutoff |dx_k/h| <= 3
//                                                            (sph_header.h:49-60, sph_lib.cc:216-218)
//          (rows the stage's contr_* predicate excludes carry zeros in the stage record instead)
//          candidates that pass are appended (ballot + popc) to a per-warp queue in shared memory;
//   pair : whenever the queue holds >= 32 entries, all 32 lanes take one entry each and run the
//          stage's pair body — the FP64-heavy part runs with every lane busy.
// The set of pairs that contribute is exactly the reference's; pairs that cannot contribute
// (zero kernel weight) are never queued.  Lane partial sums are combined with warp shuffles.
// CTAs are persistent, ONE per SM with 16-24 warps, and walk chunks of 256-384 consecutive entries
// of the list of rows the stage updates; rows are sorted by subcell, so all warps of an SM work on one
// compact block of space at a time and share its candidate rows through L1.  No CTA-wide barrier.
#pragma once
#include "psph_internal.cuh"

constexpr int RPW = 16;        // consecutive list entries per warp and chunk; a CTA chunk is S::kWarps * RPW entries
constexpr int QCAP = 128;      // queue slots per warp (power of two, >= 31 + 64)
constexpr int CANDCAP = 768;   // candidate rows expanded per batch
constexpr unsigned FULL = 0xffffffffu;

struct WarpScratch
{
  int cand[CANDCAP];           // candidate row | own-bucket flag in bit 31
  int qj[QCAP];                // rows that passed the gates, waiting for a pair round
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
// four sums at once: two exchange steps that halve the number of live values, then a three-step tree
// (9 shuffles of doubles instead of 20); every lane returns all four totals
__device__ __forceinline__ void warp_sum4(double &a, double &b, double &c, double &d, int lane)
{
  const bool hi16 = lane & 16, hi8 = lane & 8;
  double k0 = hi16 ? b : a, t0 = hi16 ? a : b;      // lanes < 16 keep a, lanes >= 16 keep b
  double k1 = hi16 ? d : c, t1 = hi16 ? c : d;
  k0 += __shfl_xor_sync(FULL, t0, 16);
  k1 += __shfl_xor_sync(FULL, t1, 16);
  double k = hi8 ? k1 : k0, t = hi8 ? k0 : k1;      // bit 3 clear: a/b, set: c/d
  k += __shfl_xor_sync(FULL, t, 8);
  k += __shfl_xor_sync(FULL, k, 4);
  k += __shfl_xor_sync(FULL, k, 2);
  k += __shfl_xor_sync(FULL, k, 1);
  a = __shfl_sync(FULL, k, 0); b = __shfl_sync(FULL, k, 16); c = __shfl_sync(FULL, k, 8); d = __shfl_sync(FULL, k, 24);
}
__device__ __forceinline__ int warp_sum_i(int v)
{
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// shared memory through 32-bit window addresses (one base register per warp instead of a generic pointer
// that the compiler re-derives at every access)
__device__ __forceinline__ void sts32(unsigned addr, int v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ int lds32(unsigned addr)
{
  int v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
// |x| of a double as an ordered 64-bit pattern (non-negative doubles compare like unsigned integers)
__device__ __forceinline__ unsigned long long abs_bits(double x)
{
  return ((unsigned long long) (unsigned) (__double2hiint(x) & 0x7fffffff) << 32) | (unsigned) __double2loint(x);
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
// ---- branch-free FP64 helpers for the pair bodies ------------------------------------------
// The pair loops are bound by dependent-instruction latency, not by the FP64 pipe: libdevice's
// exp() and '/' each carry a slow-path branch that splits the pair body into basic blocks and
// stops the scheduler from interleaving the independent chains.  These versions are straight
// line code, valid for the argument ranges that occur here (normal, finite operands), and
// accurate to <= 2 ulp — inside the FP64 tolerance the parity tests state.

// 1/x for normal finite x: hardware seed (MUFU.RCP64H: relative error e0 <= 1e-6 measured) + ONE third-order step
// r (1 + e + e^2), e = 1 - x r: the truncation error e0^3 = 1e-18 is below half an ulp by two orders of magnitude, so the
// result is the correctly rounded quotient except where 1/x lies within 1e-18 of a rounding boundary.  Three dependent FMAs
// where two Newton steps take four.  Measured over 2^24 operands in [2^-20, 2^20] (profiles/rcp_check.cu ->
// profiles/r02f_rcp_check.json): "halley" column.
__device__ __forceinline__ double dev_rcp(double x)
{
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-x, r, 1.0);
  return fma(r, fma(e, e, e), r);
}
// polynomial coefficients as constant-bank operands (no per-use 64-bit immediate moves).  The same 14 values also travel in
// KParams::cexp: kernel parameters live in constant bank 0, which FP64 instructions read as a direct operand, whereas a
// user __constant__ array (bank 3) costs one LDC per coefficient and pair round — 10 of ~150 instructions in the
// mom_engr_update body.  The hot pair bodies use the K.cexp overloads below.
__constant__ double c_exp[14] = PSPH_EXP_COEF;
// exp(x) for -700 < x <= 0 (the kernel exponent is >= -27): k = rint(x*log2 e), r = x - k ln 2,
// degree-11 polynomial in Estrin form, scaled by 2^k through the exponent field
__device__ __forceinline__ double dev_exp_neg(double x, const double *__restrict__ ce)
{
  const double kd0 = fma(x, ce[0], ce[1]);
  const int k = __double2loint(kd0);
  const double kd = kd0 - ce[1];
  double r = fma(kd, ce[2], x);   // ln 2 split hi + lo
  r = fma(kd, ce[3], r);
  const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
  const double p01 = r + 1.0;
  const double p23 = fma(r, ce[4], ce[5]);
  const double p45 = fma(r, ce[6], ce[7]);
  const double p67 = fma(r, ce[8], ce[9]);
  const double p89 = fma(r, ce[10], ce[11]);
  const double pab = fma(r, ce[12], ce[13]);
  const double q0 = fma(r2, p23, p01), q1 = fma(r2, p67, p45), q2 = fma(r2, pab, p89);
  const double p = fma(r8, q2, fma(r4, q1, q0));
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}
__device__ __forceinline__ double dev_exp_neg(double x) { return dev_exp_neg(x, c_exp); }
// SPH kernel (sph_lib.cc:207-230) for a pair already known to be inside the cutoff:
// w = (pi^-1/2 / h)^DIM * exp(-|dx|^2/h^2)
template <int DIM>
__device__ __forceinline__ double dev_weight_in(double r2, double invh)
{
  const double t = 0.564189583547756 * invh;
  const double norm = DIM == 3 ? t * t * t : t;
  return norm * dev_exp_neg(-(r2 * invh * invh));
}

// the same with the row constants hoisted: ninvh2 = -1/h^2, norm = (pi^-1/2 / h)^DIM
template <int DIM>
__device__ __forceinline__ double dev_weight_norm(double invh)
{
  const double t = 0.564189583547756 * invh;
  return DIM == 3 ? t * t * t : t;
}
__device__ __forceinline__ double dev_weight_pre(double r2, double ninvh2, double norm) { return norm * dev_exp_neg(r2 * ninvh2); }
// the same two with the coefficients taken from the kernel parameters (K.cexp)
template <int DIM>
__device__ __forceinline__ double dev_weight_in(double r2, double invh, const double *__restrict__ ce)
{
  const double t = 0.564189583547756 * invh;
  const double norm = DIM == 3 ? t * t * t : t;
  return norm * dev_exp_neg(-(r2 * invh * invh), ce);
}
__device__ __forceinline__ double dev_weight_pre(double r2, double ninvh2, double norm, const double *__restrict__ ce)
{ return norm * dev_exp_neg(r2 * ninvh2, ce); }

// Stage interface (static members):
//   kSymH         pair smoothing length is 0.5*(h_i+h_j) (true) or S::I::hk (false)
//   kSkipSelf     j == i is not a pair (mom_engr_update.cc:201)
//   kRows         RowFlags bit of the rows the stage updates (selects the row list it walks)
//   kWarps        warps per CTA; ONE CTA per SM, so that every warp of an SM works in the same compact block
//                 of space and the block's candidate rows stay in that SM's L1
//   kCountSupport count (A::nreal) the listed rows inside in_support(), whatever their weight
//   struct I      per-row state, identical in every lane; members h, supp (= CUTOFF2*h_i), hk
//   struct A      per-lane accumulators
//   void  init(V,K,ss,i,I&)             load row i
//   double reach(K,I,hmax)              half-width of the candidate box
//   void  begin(K,I&,A&,pass)           zero the accumulators
//   void  pair(V,K,I,A&,dx,dy,dz,hj,j)  one pair inside the kernel support
//   bool  end(V,K,ss,i,I&,A&,pass,lane) warp-reduce + write; true = run another pass
// Rows that must not contribute to a stage's sums (contr_dens / contr_vel_smooth false) carry
// zeros in the stage record, so the scan phase needs no flag word.
// OVF: only the rows whose near list overflowed in this step's mom_engr_update (smooth_density: the rows k_dens_near left)
template <int DIM, class S, bool OVF = false>
__global__ void __launch_bounds__(S::kWarps * 32, 1) k_gather(DevView V, GridView G, KParams K, StepState *ss)
{
  if (OVF && ss->n_near_ovf == 0) return;
  extern __shared__ __align__(16) unsigned char s_raw[];
  WarpScratch *s_scr = reinterpret_cast<WarpScratch *>(s_raw);
  constexpr int CHUNK = S::kWarps * RPW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const unsigned s_cand = (unsigned) __cvta_generic_to_shared(&s_scr[warp].cand[0]);
  const unsigned s_q = (unsigned) __cvta_generic_to_shared(&s_scr[warp].qj[0]);
  const int n = ss->n;
  const int cnt = S::kRows == F_NEED ? ss->n_need : ss->n_grad;
  const int32_t *__restrict__ rows = S::kRows == F_NEED ? V.need_idx : V.grad_idx;
  const double hmax = ss->hmax;
  for (int chunk = blockIdx.x; chunk * CHUNK < cnt; chunk += gridDim.x) {
    const int e0 = chunk * CHUNK + warp * RPW;
    for (int r = 0; r < RPW; r++) {
      if (e0 + r >= cnt) break;
      const int i = __ldg(&rows[e0 + r]);
      if (OVF && (V.nbr_cnt[i] < 0 || V.near[NEARW * i] != NEAR_OVF)) continue;
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
        // the list / in_support limits can only bite when the largest possible cutoff box exceeds them
        const bool need_lim = S::kCountSupport || K.cutoff * (S::kSymH ? 0.5 * (st.h + hmax) : st.hk) > lim_other;
        int qn = 0, qhead = 0; // queue state, identical in every lane

        // ---- scan phase, one candidate per lane: list gate (neighs.cc:88-92 own bucket, :130 others),
        // in_support (sph_header.h:49-60) and kernel cutoff |dx_k / h| <= 3 (sph_lib.cc:216-218).  The
        // cutoff is decided without dividing unless a component sits within 2^-48 of the face.
        auto consider = [&](bool live, int cj, const double4 &gj) {
          const int j = cj & 0x7fffffff;
          // |dx_k| compared as ordered bit patterns: keeps the three-way tests off the FP64 pipe
          const double dx = gi.x - gj.x, dy = DIM == 3 ? gi.y - gj.y : 0.0, dz = DIM == 3 ? gi.z - gj.z : 0.0;
          const unsigned long long ax = abs_bits(dx), ay = abs_bits(dy), az = abs_bits(dz);
          if constexpr (S::kCountSupport) { // rows of the list inside in_support(), contributing or not
            const double lim = cj < 0 ? lim_own : lim_other;
            if (live && j != i && !(fabs(dx) > lim || fabs(dy) > lim || fabs(dz) > lim)) acc.nreal++;
          }
          const double h = S::kSymH ? 0.5 * (st.h + gj.w) : st.hk;
          const double b = K.cutoff * h;
          const unsigned long long blo = (unsigned long long) __double_as_longlong(b * (1.0 - 0x1p-48));
          bool ok = live && ax <= blo && ay <= blo && az <= blo;
          if (live && !ok) { // outside, or within rounding of the cutoff face
            const unsigned long long bhi = (unsigned long long) __double_as_longlong(b * (1.0 + 0x1p-48));
            if (ax <= bhi && ay <= bhi && az <= bhi)
              ok = !(fabs(dx / h) > K.cutoff || fabs(dy / h) > K.cutoff || fabs(dz / h) > K.cutoff);
          }
          if (need_lim && ok) {
            const double lim = cj < 0 ? lim_own : lim_other;
            ok = !(fabs(dx) > lim || fabs(dy) > lim || fabs(dz) > lim);
          }
          if (S::kSkipSelf && j == i) ok = false;
          const unsigned m = __ballot_sync(FULL, ok);
          if (ok) sts32(s_q + 4u * ((qhead + qn + __popc(m & lt)) & (QCAP - 1)), j);
          qn += __popc(m);
        };

        // ---- candidate enumeration state
        int j0 = 0;                                  // 1-D: next row
        int gx0 = 0, gy0 = 0, gz0 = 0, gz1 = 0, ncoly = 1, ncol = 0, cz0 = 0, cz1 = 0;
        float inv_ncoly = 1.0f;
        int col_base = -32, cz = 0, c0 = 0, total = 0, beg = 0, len = 0, cellxy = 0, subxy = 0;
        bool have = false, own = false;
        int cell_i = 0;
        if (DIM == 3) {
          cell_i = V.t.i[I_CELL][i];
          const unsigned long long pk = V.gsub[i];
          const int sx = (int) (pk & 0xfffff), sy = (int) ((pk >> 20) & 0xfffff), sz = (int) ((pk >> 40) & 0xfffff);
          const int kR = (int) ceil(S::reach(K, st, hmax) * G.inv_subw - 1e-9) + (int) (pk >> 60);
          gx0 = max(sx - kR, 0); gy0 = max(sy - kR, 0); gz0 = max(sz - kR, 0);
          const int gx1 = min(sx + kR, G.ncell[0] * SUBDIV - 1), gy1 = min(sy + kR, G.ncell[1] * SUBDIV - 1);
          gz1 = min(sz + kR, G.ncell[2] * SUBDIV - 1);
          ncoly = gy1 - gy0 + 1; ncol = (gx1 - gx0 + 1) * ncoly;
          inv_ncoly = 1.0f / (float) ncoly;
          cz0 = gz0 / SUBDIV; cz1 = gz1 / SUBDIV;
          cz = cz1;
        }
        // one scan iteration (up to 64 candidates); false when no candidate is left
        auto scan_next = [&]() -> bool {
          if (DIM == 1) {
            // the 1-D reference has no search: every particle lists every particle (add_air.cc:470-479)
            if (j0 >= n) return false;
            const int jj = min(j0 + lane, n - 1);
            consider(j0 + lane < n, jj | 0x80000000, ld_rec(V.geo, jj));
            j0 += 32;
            return true;
          }
          for (;;) {
            if (c0 < total) { // two candidates per lane, both loads in flight
              const int ca = c0 + lane, cb = ca + 32;
              const bool la = ca < total, lb = cb < total;
              const int ja = la ? lds32(s_cand + 4u * ca) : i, jb = lb ? lds32(s_cand + 4u * cb) : i;
              const double4 ga = ld_rec(V.geo, ja & 0x7fffffff), gb = ld_rec(V.geo, jb & 0x7fffffff);
              consider(la, ja, ga);
              if (c0 + 32 < total) consider(lb, jb, gb);
              c0 += 64;
              return true;
            }
            if (total) { total = 0; c0 = 0; __syncwarp(); } // list consumed
            // refill: expand row ranges into the candidate list until it is full or nothing is left
            for (;;) {
              if (__any_sync(FULL, len > 0)) {
                const int quota = (CANDCAP - total) / 32;
                if (quota == 0) break;
                const int take = min(len, quota);
                int off = take;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, off, o); if (lane >= o) off += v; }
                const int sum = __shfl_sync(FULL, off, 31);
                off += total - take;
                const int tag = own ? 0x80000000 : 0;
                for (int k = 0; k < take; k++) sts32(s_cand + 4u * (off + k), (beg + k) | tag);
                beg += take; len -= take; total += sum;
                continue;
              }
              if (cz < cz1) { // the z-run of every column of this round inside the next bucket along z
                cz++;
                beg = 0; len = 0; own = false;
                if (have) {
                  const int cell = cellxy + cz;
                  const int zlo = (cz == cz0) ? (gz0 % SUBDIV) : 0, zhi = (cz == cz1) ? (gz1 % SUBDIV) : SUBDIV - 1;
                  const int t = __ldg(&G.cell2rank[cell]) * SUBS + subxy;
                  beg = __ldg(&G.sub_start[t + zlo]);
                  len = __ldg(&G.sub_start[t + zhi + 1]) - beg;
                  own = cell == cell_i;
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
            if (total == 0) return false;
            __syncwarp();
          }
        };

        // ---- alternate: pair rounds while the queue holds a full warp's worth, else scan on
        for (bool flush = false;;) {
          while (qn >= 32 || (flush && qn > 0)) {
            __syncwarp();
            const int c = min(32, qn);
            if (lane < c) {
              const int j = lds32(s_q + 4u * ((qhead + lane) & (QCAP - 1)));
              const double4 gj = ld_rec(V.geo, j);
              S::pair(V, K, st, acc, gi.x - gj.x, gi.y - gj.y, gi.z - gj.z, gj.w, j);
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

template <int DIM, class S, bool OVF = false>
int launch_gather_d(psph_ctx *ctx, int kid)
{
  constexpr int threads = S::kWarps * 32;
  constexpr size_t smem = sizeof(WarpScratch) * S::kWarps;
  // the attribute is per device, and a process may drive several devices (one context each): set it at every launch
  PSPH_CUDA(cudaFuncSetAttribute(k_gather<DIM, S, OVF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  static int per_sm = 0; // resident CTAs per SM of this instantiation: the grid is one persistent wave
  if (per_sm == 0 &&
      (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gather<DIM, S, OVF>, threads, smem) != cudaSuccess || per_sm < 1))
    per_sm = 1;
  KernelTimer kt(ctx, kid);
  int nchunks = (ctx->n + S::kWarps * RPW - 1) / (S::kWarps * RPW);
  int grid = min(max(nchunks, 1), ctx->sms * per_sm);
  k_gather<DIM, S, OVF><<<grid, threads, smem, ctx->stream>>>(ctx->cur, psph_grid_view(ctx), ctx->kp, ctx->d_ss);
  PSPH_LAUNCH_CHECK();
  return 0;
}
template <template <int> class S>
int launch_gather(psph_ctx *ctx, int kid)
{
  if (ctx->cfg.dim == 3) return launch_gather_d<3, S<3> >(ctx, kid);
  return launch_gather_d<1, S<1> >(ctx, kid);
}
