// psph_stages.cu — the per-time-step stages of the SPH (artificial-viscosity) path:
// timestep, adaptive_sml, mom_engr_update, smooth_density, smooth_velocity, update_pos and the
// fused step driver.  Every stage cites the reference code it replaces; the C ABI is
// include/psph.h.  Pair sums run on the warp gather of psph_gather.cuh.
#include "psph_mix.cuh"
#include <cstring>
#include <algorithm>

// =========================================================================================
// timestep  (src/plume/timestep.cc:95-135 3-D, :140-183 1-D): dt = CFL * min_i h_i / c_i over
// need_neigh particles.  TimeProps::incrtime (properties.h:71-82) clips dt on the host; the
// clip arrives here as dt_max.
// =========================================================================================
__global__ void k_timestep_reset(StepState *ss) { ss->min_hc_bits = (unsigned long long) __double_as_longlong(1.0E+10); }

__global__ void k_timestep(Tab V, StepState *ss)
{
  const int n = ss->n;
  double m = 1.0E+10;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x)
    if (dev_need_neigh(V.i[I_BC][p], V.i[I_GUEST][p])) {
      double t = V.d[D_H][p] / V.d[D_C][p];
      if (t < m) m = t;
    }
  for (int o = 16; o; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
  // positive doubles order like their bit patterns
  if ((threadIdx.x & 31) == 0 && m < 1.0E+10) atomicMin(&ss->min_hc_bits, (unsigned long long) __double_as_longlong(m));
}

__global__ void k_timestep_finish(KParams K, StepState *ss, double *dt_cfl_out)
{
  double dt = K.cfl * __longlong_as_double((long long) ss->min_hc_bits);
  if (dt_cfl_out) *dt_cfl_out = dt;
  ss->dt = fmin(dt, ss->dt_max);
}

// =========================================================================================
// stage records (32-byte rows read by the pair loops with 128-bit loads)
// =========================================================================================
__global__ void k_hmax_scan(Tab V, StepState *ss)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  double h = p < ss->n ? V.d[D_H][p] : 0.0;
  for (int o = 16; o; o >>= 1) h = fmax(h, __shfl_xor_sync(0xffffffffu, h, o));
  if ((threadIdx.x & 31) == 0 && h > 0.0) atomicMax(&ss->hmax_bits, (unsigned long long) __double_as_longlong(h));
}
__global__ void k_refresh_geo(DevView V, KParams K, StepState *ss, unsigned long long *__restrict__ cell_hmax)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  bool std_h = false;
  if (p < ss->n) {
    const double h = V.t.d[D_H][p];
    const int cell = V.t.i[I_CELL][p];
    const unsigned long long hb = (unsigned long long) __double_as_longlong(h);
    if (cell >= 0 && hb > cell_hmax[cell]) atomicMax(&cell_hmax[cell], hb);
    const double x = V.t.d[D_X][p], y = V.t.d[D_Y][p], z = V.t.d[D_Z][p];
    V.geo[p] = make_double4(x, y, z, h);
    dev_store_geo32(V.geo32, K, p, x, y, z, h, ss->hmax); // positions are those of the list build: xmax_bits still bounds them
    std_h = (V.flags[p] & F_NEED) && h == ss->hmax;
  }
  const unsigned b = __ballot_sync(0xffffffffu, std_h);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(&ss->n_std, __popc(b));
}
__global__ void k_hmax_reset(StepState *ss) { ss->hmax_bits = 0ull; ss->n_std = 0; }
// The pair path of the step (StepState::pair_mode), decided where the data is: block tiles need one standard smoothing
// length in a row's whole neighbourhood, so they serve the table while most summed rows carry it (every step before the
// first adaptive_sml call); after that every row has its own h and the sub-tiled mixed-h gather takes over.
// forced: 2 / 4 = that path whatever the table looks like, 0 = decide, -1 = neither (legacy list modes, host-driven)
__global__ void k_pair_mode(StepState *ss, int forced, int dens_near)
{
  int mode = forced > 0 ? forced : (forced == 0 ? (2LL * ss->n_std >= ss->n_need ? 2 : 4) : 0);
  if (mode == 4 && !dens_near) mode = 0;       // (the mixed-h gather has no density pass of its own: PSPH_DENSNEAR=0 -> fused gather)
  ss->pair_mode = mode;
}
// a step on neither tile path: no row has a list, the fused grouped gather sums them all
__global__ void k_no_lists(DevView V, StepState *ss)
{
  if (ss->pair_mode == 2 || ss->pair_mode == 4) return;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p == 0) ss->n_uncached = ss->n_need;
  if (p < ss->n) V.nbr_cnt[p] = -1;
}
__global__ void k_hmax_finish(StepState *ss) { ss->hmax = __longlong_as_double((long long) ss->hmax_bits); }

// mom_engr_update: rec0 = {m, p*V*V, V, rho}  rec1 = {c, Cp, T, -}  rec2 = {u_x, u_y, u_z, -}
// with V = 1/rho and u = smoothed velocity when HAVE_TURBULENCE_LANS != 0 (mom_engr_update.cc:170-176)
__global__ void k_pack_mom(DevView V, KParams K, StepState *ss)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ss->n) return;
  double rho = V.t.d[D_RHO][p], Vj = 1.0 / rho;
  V.rec[0][p] = make_double4(V.t.d[D_MASS][p], V.t.d[D_P][p] * Vj * Vj, Vj, rho);
  V.rec[1][p] = make_double4(V.t.d[D_C][p], V.t.d[D_CP][p], V.t.d[D_T][p], 0.0);
  const int u = K.lans ? D_SVX : D_VX;
  // .w: the mass the row carries in smooth_density's sums (0: contr_dens() false or phase outside {1,2},
  // density_massfrac.cc:135,141), negative for phase 2 — the tile gather forms those sums on the side (psph_tile.cuh)
  const int f = V.flags[p];
  double md = (f & (F_CDENS | F_PHOK)) == (F_CDENS | F_PHOK) ? V.t.d[D_MASS][p] : 0.0;
  if (f & F_PH2) md = -md;
  V.rec[2][p] = make_double4(V.t.d[u][p], V.t.d[u + 1][p], V.t.d[u + 2][p], md);
}
// smooth_density: rec0 = {m, m/rho, m*e/rho, phase_num == 2}
__global__ void k_pack_dens(DevView V, StepState *ss)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ss->n) return;
  double m = V.t.d[D_MASS][p], mv = m / V.t.d[D_RHO][p];
  const int f = V.flags[p];
  // contr_dens() false or phase outside {1,2} (density_massfrac.cc:135,141): the row adds exact zeros
  if ((f & (F_CDENS | F_PHOK)) != (F_CDENS | F_PHOK)) { m = 0.0; mv = 0.0; }
  V.rec[0][p] = make_double4(m, mv, mv * V.t.d[D_E][p], (f & F_PH2) ? 1.0 : 0.0);
}
// smooth_velocity: rec0 = {m/rho, v_x, v_y, v_z}
__global__ void k_pack_vel(DevView V, StepState *ss)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ss->n) return;
  // contr_vel_smooth() false (smooth_velocity.cc:100): zero weight
  const double mv = (V.flags[p] & F_CVEL) ? V.t.d[D_MASS][p] / V.t.d[D_RHO][p] : 0.0;
  V.rec[0][p] = make_double4(mv, V.t.d[D_VX][p], V.t.d[D_VY][p], V.t.d[D_VZ][p]);
}

// adaptive_sml: scr[7] = m_j (0 when contr_dens() is false, neighs.cc:556)
__global__ void k_pack_sml(DevView V, StepState *ss)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ss->n) return;
  V.scr[7][p] = (V.flags[p] & F_CDENS) ? V.t.d[D_MASS][p] : 0.0;
}

// the largest |t| with fabs(t / h) <= cutoff as the reference's division rounds it (sph_lib.cc:216-218): with one h
// per row the support test of a kernel is three compares per pair instead of three divisions
__device__ __forceinline__ double dev_cutoff_bound(double cutoff, double h)
{
  double b = cutoff * h;
  for (int k = 0; k < 4 && b / h > cutoff; k++) b = __longlong_as_double(__double_as_longlong(b) - 1);
  for (int k = 0; k < 4 && __longlong_as_double(__double_as_longlong(b) + 1) / h <= cutoff; k++)
    b = __longlong_as_double(__double_as_longlong(b) + 1);
  return b;
}

// =========================================================================================
// smooth_density  (src/plume/density_massfrac.cc:25-359; DENSITY_UPDATE_SPH 11, DENSITY_UPDATE_SML 11)
// =========================================================================================
template <int DIM>
struct DensStage
{
  static constexpr bool kSymH = true, kSkipSelf = false, kCountSupport = false;
  static constexpr int kRows = F_NEED, kWarps = 24, kWarpsGrp = 24, kPfGrp = 0, kWarpsScan = 24, kWarpsList = 24, kPfList = 0, kNRec = 1;
  struct I { double h, supp, hk, invh, ninvh2, norm, ninvh2e, norme, tin; };  // from invh on: uniform-h rows only (init_uni)
  struct A { double rho1, rho2, wn, en; };
  static __device__ __forceinline__ double reach_h(const KParams &K, double h, double hmax)
  { return fmin(K.cutoff * 0.5 * (h + hmax), K.cutoff2 * h); }
  static __device__ __forceinline__ void init(const DevView &V, const KParams &K, StepState *, int i, I &s)
  {
    s.h = V.t.d[D_H][i];
    s.supp = K.cutoff2 * s.h; // :93
  }
  static __device__ __forceinline__ double reach(const KParams &K, const I &s, double hmax)
  { return fmin(K.cutoff * 0.5 * (s.h + hmax), s.supp); }
  static __device__ __forceinline__ void begin(const KParams &, I &, A &a, int) { a.rho1 = a.rho2 = a.wn = a.en = 0.0; }
  struct R { double4 r0; };                                          // the stage record of row j
  static __device__ __forceinline__ R load(const DevView &V, int j) { R r; r.r0 = ld_rec(V.rec[0], j); return r; }
  static __device__ __forceinline__ void zero(R &r) { r.r0 = make_double4(0., 0., 0., 0.); }  // a row that adds exact zeros
  // block-tiled gather (psph_tile.cuh): columns 3.. of a tile row in shared memory, and the record read back
  // tile columns 3..5: m (negative: phase 2), m/rho, m e/rho.  kDefer (the tile gather queues the pairs that can lie
  // inside the energy-smoothing kernel and runs pair_inner on them 32 at a time) measured 6 % slower than running
  // them in place: off
  static constexpr int kNC = 6, kWarpsTile = 16, kMinbTile = 2;
  static constexpr bool kDefer = false;
  // pair-weight cache (psph_tile.cuh): this stage reads the weights mom_engr_update stored; positions stay in the tile
  // for the few pairs inside the energy-smoothing kernel
  static constexpr int kWCache = 2, kNCc = kNC;
  static constexpr bool kCachedNeedsX = true;
  static constexpr bool kSkipUnlessFull = true;                      // k_dens_near serves the listed rows unless ss->dens_full
  static constexpr int kNAm = 0;                                     // no pass of its own in the mixed-h gather (psph_mix.cuh)
  static __device__ __forceinline__ void tile_load(const DevView &V, int row, double *c, int t)
  { const double4 r = ld_rec(V.rec[0], row); c[3 * TCAP + t] = r.w != 0.0 ? -r.x : r.x; c[4 * TCAP + t] = r.y; c[5 * TCAP + t] = r.z; }
  static __device__ __forceinline__ R tile_rec(const double *c, int t)
  { R r; const double m = c[3 * TCAP + t]; r.r0 = make_double4(fabs(m), c[4 * TCAP + t], c[5 * TCAP + t], m < 0.0 ? 1.0 : 0.0); return r; }
  static __device__ __forceinline__ void tile_stage(const DevView &V, int row, double *c, int t) { tile_stage_soa<DensStage<DIM> >(V, row, c, t); }
  static __device__ __forceinline__ void tile_xyz(const double *c, int t, double &x, double &y, double &z) { tile_xyz_soa(c, t, x, y, z); }
  static __device__ __forceinline__ R tile_recp(const double *c, int t, double &x, double &y, double &z)
  { tile_xyz_soa(c, t, x, y, z); return tile_rec(c, t); }
  // the state of row i from its own slot of the tile; a listed row carries the table's standard smoothing length
  static __device__ __forceinline__ void tile_init(const DevView &, const KParams &, StepState *, const double *, int, int, I &) {}
  // the row constants of a tile launch: every listed row carries the table's standard smoothing length
  static __device__ __forceinline__ void tile_uni(const KParams &K, double hstd, I &s)
  { s.h = hstd; s.supp = K.cutoff2 * s.h; init_uni(K, s); }
  // rows whose partners all carry the row's own smoothing length (pair lists): 1/h and the kernel norms once per row
  static __device__ __forceinline__ void init_uni(const KParams &K, I &s)
  {
    s.invh = dev_rcp(s.h); s.ninvh2 = -(s.invh * s.invh); s.norm = dev_weight_norm<DIM>(s.invh);
    const double ie = K.e_ratio * s.invh;
    s.ninvh2e = -(ie * ie); s.norme = dev_weight_norm<DIM>(ie);
    s.tin = dev_cutoff_bound(K.cutoff, s.h);
  }
  static __device__ __forceinline__ double self_weight(const I &s) { return -dev_weight_pre(0.0, s.ninvh2, s.norm); }
  // one listed pair with its weight from the cache (:159-161).  A negative sign marks the pairs inside the
  // energy-smoothing kernel (MomStage::cache_flag; the row's own entry always is): only those run pair_inner.
  // (Queueing them to run 32 at a time measured slower, like kDefer.)
  static __device__ __forceinline__ void pair_c(const KParams &K, const I &s, A &a, double w, const double *c, int sl, int self)
  {
    const double m = c[3 * TCAP + sl], wm = fabs(w) * fabs(m);
    a.rho2 += m < 0.0 ? wm : 0.0;
    a.rho1 += m < 0.0 ? 0.0 : wm;
    if (w < 0.0) pair_ci(K, s, a, c, sl, self);
  }
  static __device__ __forceinline__ void pair_ci(const KParams &K, const I &s, A &a, const double *c, int sl, int self)
  {
    const double dx = c[0 * TCAP + self] - c[0 * TCAP + sl], dy = c[1 * TCAP + self] - c[1 * TCAP + sl],
                 dz = c[2 * TCAP + self] - c[2 * TCAP + sl];
    pair_inner<true>(K, s, a, dx, dy, dz, s.h, tile_rec(c, sl));
  }
  static __device__ __forceinline__ void pair(const DevView &V, const KParams &K, const I &s, A &a, double dx, double dy,
                                              double dz, double hj, int j)
  { pair_r<false>(K, s, a, dx, dy, dz, hj, load(V, j)); }
  // the density sums of one pair; returns whether the pair can lie inside the (narrower) energy-smoothing kernel:
  // all |q dx_k / h| <= cutoff implies r2 <= DIM (cutoff h / q)^2, which decides most pairs with one compare
  template <bool UNI>
  static __device__ __forceinline__ bool pair_main(const KParams &K, const I &s, A &a, double dx, double dy, double dz, double hj,
                                                   const R &rec)
  {
    const double h = UNI ? s.h : 0.5 * (s.h + hj);                  // :155
    const double r2 = dx * dx + dy * dy + dz * dz;
    const double4 rj = rec.r0;
    const double wm = (UNI ? dev_weight_pre(r2, s.ninvh2, s.norm) : dev_weight_in<DIM>(r2, dev_rcp(h))) * rj.x;
    const bool ph2 = rj.w != 0.0;                                   // phase of j
    a.rho2 += ph2 ? wm : 0.0;
    a.rho1 += ph2 ? 0.0 : wm;
    if (K.energy_smooth != 1) return false;
    const double b = K.cutoff * h * (1.0 + 0x1p-40);
    return r2 * (K.e_ratio * K.e_ratio) <= DIM * b * b;
  }
  // the energy-smoothing sums (:163-168, kernel of width h/E_SMOOTH_RATIO) of a pair that passed pair_main's test
  template <bool UNI>
  static __device__ __forceinline__ void pair_inner(const KParams &K, const I &s, A &a, double dx, double dy, double dz, double hj,
                                                    const R &rec)
  {
    const double h = UNI ? s.h : 0.5 * (s.h + hj), q = K.e_ratio;
    if (UNI) {
      if (fabs(q * dx) > s.tin || (DIM == 3 && (fabs(q * dy) > s.tin || fabs(q * dz) > s.tin))) return;
    } else if (outside_kernel<DIM>(q * dx, q * dy, q * dz, h, K.cutoff)) return;
    const double r2 = dx * dx + dy * dy + dz * dz;
    const double we = UNI ? dev_weight_pre(r2, s.ninvh2e, s.norme) : dev_weight_in<DIM>(r2, q * dev_rcp(h));
    a.wn += we * rec.r0.y;
    a.en += we * rec.r0.z;
  }
  template <bool UNI>
  static __device__ __forceinline__ void pair_r(const KParams &K, const I &s, A &a, double dx, double dy, double dz, double hj,
                                                const R &rec)
  {
    if (pair_main<UNI>(K, s, a, dx, dy, dz, hj, rec)) pair_inner<UNI>(K, s, a, dx, dy, dz, hj, rec);
  }
  static __device__ __forceinline__ bool end(const DevView &V, const KParams &K, StepState *ss, int i, I &, A &a, int, int lane)
  {
    double r1 = a.rho1, r2 = a.rho2, wn = a.wn, en = a.en;
    warp_sum4(r1, r2, wn, en, lane);
    if (lane == 0) {
      const double rho = r1 + r2;                                   // :268-273
      V.scr[0][i] = rho;
      V.scr[1][i] = r2 / rho;                                       // :285
      int bad = !(rho > 0);
      if (K.energy_smooth == 1) { V.scr[2][i] = en / wn; bad += !(wn > 0); } // :230
      if (bad) atomicAdd(&ss->bad, bad);
    }
    return false;
  }
};

// smooth_density for the rows whose sums mom_engr_update's tile gather prepared (psph_tile.cuh, WM == 1): the phase
// densities depend on positions, masses and phases only — none of which mom_engr_update changes — so that gather
// forms them while it has the pair weights in registers (V.scr[4], V.scr[5]: the sums over the row's list), and lists
// the few partners inside the energy-smoothing kernel (V.near).  What is left for this stage is the row's own
// term and the energy sums (:163-168) over those partners with the energies mom_engr_update committed: one
// thread per row, no pass over the pair lists at all.
// UNI: every listed row carries the table's standard smoothing length (block tiles, pair_mode 2); otherwise the rows
// have their own (sub-tiled mixed-h gather, pair_mode 4) and a pair's h is (h_i + h_j) / 2.  A row whose near list
// overflowed (count NEAR_OVF) is left to k_gather<DensStage, OVF>.
template <bool UNI>
__global__ void k_dens_near(DevView V, KParams K, StepState *ss)
{
  if (ss->pair_mode != (UNI ? 2 : 4)) return;
  if (ss->dens_full) return;                                        // the tile gather sums these rows off their lists
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ss->n || !(V.flags[p] & F_NEED) || V.nbr_cnt[p] < 0) return;
  const int n = V.near[NEARW * p];
  if (n == NEAR_OVF) return;
  typedef DensStage<3> S;
  S::I st;
  if (UNI) S::tile_uni(K, ss->hmax, st);
  else S::init(V, K, ss, p, st);
  S::A acc;
  S::begin(K, st, acc, 0);
  const double4 gi = V.geo[p];
  S::R ri = S::load(V, p);
  double r1 = V.scr[4][p], r2 = V.scr[5][p];
  // the row itself (:159-161)
  const double wm = (UNI ? dev_weight_pre(0.0, st.ninvh2, st.norm) : dev_weight_in<3>(0.0, dev_rcp(st.h))) * ri.r0.x;
  if (ri.r0.w != 0.0) r2 += wm; else r1 += wm;
  S::pair_inner<UNI>(K, st, acc, 0.0, 0.0, 0.0, st.h, ri);
  for (int k = 0; k < n; k++) {
    const int j = V.near[NEARW * p + 1 + k];
    const double4 gj = V.geo[j];
    S::pair_inner<UNI>(K, st, acc, gi.x - gj.x, gi.y - gj.y, gi.z - gj.z, gj.w, S::load(V, j));
  }
  const double rho = r1 + r2;                                       // :268-273
  V.scr[0][p] = rho;
  V.scr[1][p] = r2 / rho;                                           // :285
  int bad = !(rho > 0);
  if (K.energy_smooth == 1) { V.scr[2][p] = acc.en / acc.wn; bad += !(acc.wn > 0); } // :230
  if (bad) atomicAdd(&ss->bad, bad);
}

// second pass of smooth_density (:337-352): commit, flag involved, equation of state
__global__ void k_commit_dens(DevView V, KParams K, StepState *ss)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ss->n || !(V.flags[p] & F_NEED)) return;
  const double ks = V.scr[1][p], rho = V.scr[0][p];
  V.t.d[D_KS][p] = ks;
  if (ks >= K.msfrc_thresh && V.t.i[I_INVOLVED][p] != 2) V.t.i[I_INVOLVED][p] = 2;
  V.t.d[D_RHO][p] = rho;
  double e = V.t.d[D_E][p];
  if (K.energy_smooth == 1) { e = V.scr[2][p]; V.t.d[D_E][p] = e; }
  EosOut o = dev_eos(K, rho, ks, e, V.t.d[D_C][p], V.t.d[D_CP][p], V.t.d[D_GAMMA][p]);
  V.t.d[D_P][p] = o.p; V.t.d[D_GAMMA][p] = o.gamma; V.t.d[D_C][p] = o.c; V.t.d[D_T][p] = o.T; V.t.d[D_CP][p] = o.Cp;
  V.t.d[D_RPH0][p] = o.rph0; V.t.d[D_RPH1][p] = o.rph1;
}

// =========================================================================================
// mom_engr_update, SPH variant  (src/plume/mom_engr_update.cc:30-439; ME_UPDATE_SML 2,
// MOMENTUM_DISCRETIZE 0; art_vis sph_lib.cc:402-424; SPH-epsilon terms :1331-1439; compute_F :257-281)
// =========================================================================================
template <int DIM>
struct MomStage
{
  static constexpr bool kSymH = true, kSkipSelf = true, kCountSupport = false;
  static constexpr int kRows = F_NEED, kWarps = 16, kWarpsGrp = 16, kPfGrp = 0, kWarpsScan = 24, kWarpsList = 16, kPfList = 0, kNRec = 3;
  struct I { int i; double h, supp, hk, rho, hc, pvsq, c, T, Cp, u[3], v0[3], e0, invh, ninvh2, norm, tin_e; };  // from invh on: init_uni / tile_uni
  struct A { double rv[3], re; };
  static __device__ __forceinline__ double reach_h(const KParams &K, double h, double hmax)
  { return fmin(K.cutoff * 0.5 * (h + hmax), K.cutoff2 * h); }
  static __device__ __forceinline__ void init(const DevView &V, const KParams &K, StepState *ss, int i, I &s)
  {
    s.i = i;
    s.h = V.t.d[D_H][i];
    s.supp = K.cutoff2 * s.h;
    s.rho = V.t.d[D_RHO][i];
    const double Vi = 1.0 / s.rho;
    s.pvsq = V.t.d[D_P][i] * Vi * Vi;
    s.hc = K.heat_c * Vi;
    s.c = V.t.d[D_C][i]; s.T = V.t.d[D_T][i]; s.Cp = V.t.d[D_CP][i];
    const int u = K.lans ? D_SVX : D_VX;
    s.u[0] = V.t.d[u][i]; s.u[1] = V.t.d[u + 1][i]; s.u[2] = V.t.d[u + 2][i];
    load_old(V, i, s);
  }
  // what end() updates, read before the pair loop so that the row does not wait for it afterwards
  static __device__ __forceinline__ void load_old(const DevView &V, int i, I &s)
  { s.v0[0] = V.t.d[D_VX][i]; s.v0[1] = DIM == 3 ? V.t.d[D_VY][i] : 0.0; s.v0[2] = DIM == 3 ? V.t.d[D_VZ][i] : 0.0; s.e0 = V.t.d[D_E][i]; }
  static __device__ __forceinline__ double reach(const KParams &K, const I &s, double hmax)
  { return fmin(K.cutoff * 0.5 * (s.h + hmax), s.supp); }
  static __device__ __forceinline__ void begin(const KParams &, I &, A &a, int) { a.rv[0] = a.rv[1] = a.rv[2] = a.re = 0.0; }
  struct R { double4 r0, r1, r2; };                                  // the stage records of row j
  static __device__ __forceinline__ R load(const DevView &V, int j)
  { R r; r.r0 = ld_rec(V.rec[0], j); r.r1 = ld_rec(V.rec[1], j); r.r2 = ld_rec(V.rec[2], j); return r; }
  static __device__ __forceinline__ void zero(R &r) { r.r0.x = 0.; }                          // m_j = 0: the pair adds exact zeros
  // The tile of this stage is seven arrays of 16-byte elements, {x, y} {z, m} {p V^2, V} {rho, c} {Cp, T} {u_x, u_y}
  // {u_z, m_dens}: the slots of a list come in runs of consecutive rows (the z-run of a subcell column), so the 8 lanes
  // of a quarter warp read 8 consecutive elements — 128-bit loads at the conflict-free rate, where 64-bit loads of
  // plain columns measured 2.3 wavefronts over the minimum (profiles/r01o_ncu_wcache_n96.txt)
  static constexpr int kNC = 14, kWarpsTile = 16, kMinbTile = 1;
  static constexpr bool kDefer = false;
  static constexpr int kWCache = 1, kNCc = kNC;                      // first stage of a step: stores the pair weights
  static constexpr bool kCachedNeedsX = true, kSkipUnlessFull = false;
  static __device__ __forceinline__ double2 &tp(double *c, int p, int t) { return reinterpret_cast<double2 *>(c)[p * TCAP + t]; }
  static __device__ __forceinline__ double2 tp(const double *c, int p, int t) { return reinterpret_cast<const double2 *>(c)[p * TCAP + t]; }
  static __device__ __forceinline__ void tile_stage(const DevView &V, int row, double *c, int t)
  {
    const double4 g = ld_rec(V.geo, row), a = ld_rec(V.rec[0], row), b = ld_rec(V.rec[1], row), u = ld_rec(V.rec[2], row);
    tp(c, 0, t) = make_double2(g.x, g.y); tp(c, 1, t) = make_double2(g.z, a.x); tp(c, 2, t) = make_double2(a.y, a.z);
    tp(c, 3, t) = make_double2(a.w, b.x); tp(c, 4, t) = make_double2(b.y, b.z); tp(c, 5, t) = make_double2(u.x, u.y);
    tp(c, 6, t) = make_double2(u.z, u.w);
  }
  static __device__ __forceinline__ void tile_xyz(const double *c, int t, double &x, double &y, double &z)
  { const double2 p0 = tp(c, 0, t); x = p0.x; y = p0.y; z = tp(c, 1, t).x; }
  static __device__ __forceinline__ R tile_recp(const double *c, int t, double &x, double &y, double &z)
  {
    const double2 p0 = tp(c, 0, t), p1 = tp(c, 1, t), p2 = tp(c, 2, t), p3 = tp(c, 3, t), p4 = tp(c, 4, t), p5 = tp(c, 5, t),
                  p6 = tp(c, 6, t);
    x = p0.x; y = p0.y; z = p1.x;
    R r;
    r.r0 = make_double4(p1.y, p2.x, p2.y, p3.x);
    r.r1 = make_double4(p3.y, p4.x, p4.y, 0.);
    r.r2 = make_double4(p5.x, p5.y, p6.x, p6.y);
    return r;
  }
  // the state of row i from its own slot of the tile (k_pack_mom forms V and p V^2 exactly as init() does)
  static __device__ __forceinline__ void tile_init(const DevView &V, const KParams &K, StepState *ss, const double *c, int t, int i, I &s)
  {
    load_old(V, i, s);
    s.i = i;
    const double2 p2 = tp(c, 2, t), p3 = tp(c, 3, t), p4 = tp(c, 4, t), p5 = tp(c, 5, t);
    s.pvsq = p2.x; s.hc = K.heat_c * p2.y; s.rho = p3.x; s.c = p3.y; s.Cp = p4.x; s.T = p4.y;
    s.u[0] = p5.x; s.u[1] = p5.y; s.u[2] = tp(c, 6, t).x;
  }
  static __device__ __forceinline__ void init_uni(const KParams &, I &s)
  { s.invh = dev_rcp(s.h); s.ninvh2 = -(s.invh * s.invh); s.norm = dev_weight_norm<DIM>(s.invh); }
  static __device__ __forceinline__ void tile_uni(const KParams &K, double hstd, I &s)
  { s.h = hstd; s.supp = K.cutoff2 * s.h; init_uni(K, s); s.tin_e = dev_cutoff_bound(K.cutoff, s.h); }
  // the bit the weight cache carries for smooth_density: the pair lies inside the support of the energy-smoothing
  // kernel W(q dx / h, h / q) (density_massfrac.cc:163-168; the test of DensStage::pair_inner on uniform-h rows)
  static __device__ __forceinline__ bool cache_flag(const KParams &K, const I &s, double dx, double dy, double dz)
  {
    const double q = K.e_ratio;
    return K.energy_smooth == 1 && !(fabs(q * dx) > s.tin_e || fabs(q * dy) > s.tin_e || fabs(q * dz) > s.tin_e);
  }
  static __device__ __forceinline__ void pair(const DevView &V, const KParams &K, const I &s, A &a, double dx, double dy,
                                              double dz, double hj, int j)
  { pair_r<false>(K, s, a, dx, dy, dz, hj, load(V, j)); }
  template <bool UNI>
  static __device__ __forceinline__ void pair_r(const KParams &K, const I &s, A &a, double dx, double dy, double dz, double hj,
                                                const R &rec)
  { (void) pair_w<UNI>(K, s, a, dx, dy, dz, hj, rec); }
  // the pair body; returns the kernel weight W(dx / h, h) of the pair (what the weight cache keeps)
  template <bool UNI>
  static __device__ __forceinline__ double pair_w(const KParams &K, const I &s, A &a, double dx, double dy, double dz, double hj,
                                                  const R &rec)
  {
    const double h = UNI ? s.h : 0.5 * (s.h + hj), h2 = h * h;
    const double r2 = dx * dx + dy * dy + dz * dz;
    const double invh = UNI ? s.invh : dev_rcp(h);
    const double4 r0 = rec.r0, r1 = rec.r1, uj = rec.r2;
    const double w = UNI ? dev_weight_pre(r2, s.ninvh2, s.norm, K.cexp) : dev_weight_in<DIM>(r2, invh, K.cexp);
    const double g = -2.0 * invh * invh * w;                        // dW/dx_k = g * dx_k
    const double vx = s.u[0] - uj.x, vy = s.u[1] - uj.y, vz = s.u[2] - uj.z;
    const double vr = dx * vx + dy * vy + dz * vz, vv = vx * vx + vy * vy + vz * vz;
    const double cab = 0.5 * (s.c + r1.x);
    // art_vis (sph_lib.cc:402-424): miu = h vr / (r2 + ata h^2); (beta miu - alf cab) miu / rho_ab if vr < 0.
    // One reciprocal serves both quotients; selected, not branched.
    const double dA = fma(K.ata, h2, r2), dB = s.rho + r0.w, rAB = dev_rcp(dA * dB);
    const double miu = h * vr * dB * rAB;
    double vis = (K.beta * miu - K.alf * cab) * miu * (2.0 * dA * rAB);
    vis = vr < 0 ? vis : 0.0;
    if (K.lans) vis -= K.epsilon_half * vv * r0.z;                  // SPH_epsilon_mom
    const double mpv = r0.x * (r0.y + s.pvsq + vis);
    double heat = 0.0;
    if (K.lans && DIM == 3) {                                       // :307-314
      const double avr = fabs(vx * dx) + fabs(vy * dy) + fabs(vz * dz);
      const bool on = avr != 0. && r2 > 0.;
      // k_ij = 10 eps Cp_ab r2 vv / (Pr avr) / (h cab)   (sph_lib.cc:1399-1439); F_ij = (gradW . dx) / r2 = g;
      // heat = m_j V_j V_i k_ij (T_i - T_j) F_ij, the constants and V_i gathered in the row's s.hc = (5 eps / Pr) V_i
      const double num = s.hc * (s.Cp + r1.y) * (r2 * vv) * (r0.x * r0.z) * ((s.T - r1.z) * g);
      heat = on ? num * dev_rcp(on ? avr * h * cab : 1.0) : 0.0;
    }
    const double mg = mpv * g;
    a.rv[0] -= mg * dx;
    if (DIM == 3) { a.rv[1] -= mg * dy; a.rv[2] -= mg * dz; }
    a.re += 0.5 * mg * vr + heat;
    return w;
  }
  // explicit Euler update of the row from its summed right-hand sides (:340-354)
  static __device__ __forceinline__ void finish(const DevView &V, const KParams &K, StepState *ss, int i, const I &s, double r0,
                                                double r1, double r2, double re)
  {
    const double dt = ss->dt, grav = -K.g;
    double un[3] = { s.v0[0] + dt * r0, s.v0[1] + dt * r1, s.v0[2] + dt * r2 };
    un[DIM - 1] += dt * grav;                                       // :350
    V.scr[0][i] = un[0]; V.scr[1][i] = un[1]; V.scr[2][i] = un[2];
    V.scr[3][i] = s.e0 + dt * (re + s.u[DIM - 1] * grav);           // :354
    if (!(s.rho > 0)) atomicAdd(&ss->bad, 1);
  }
  static __device__ __forceinline__ bool end(const DevView &V, const KParams &K, StepState *ss, int i, I &s, A &a, int, int lane)
  {
    double r0 = a.rv[0], r1 = DIM == 3 ? a.rv[1] : 0.0, r2 = DIM == 3 ? a.rv[2] : 0.0, re = a.re;
    warp_sum4(r0, r1, r2, re, lane);
    if (lane == 0) finish(V, K, ss, i, s, r0, r1, r2, re);
    return false;
  }
  // ---- sub-tiled mixed-h gather (psph_mix.cuh): seven arrays of 16-byte elements as above, then the partner's own h
  static constexpr int kNCm = 15, kNAm = 6;
  static constexpr unsigned kTagMix = MTAG_SYM;
  static __device__ __forceinline__ double2 &tm(double *c, int p, int t) { return reinterpret_cast<double2 *>(c)[p * MCAP + t]; }
  static __device__ __forceinline__ double2 tm(const double *c, int p, int t) { return reinterpret_cast<const double2 *>(c)[p * MCAP + t]; }
  static __device__ __forceinline__ void mix_stage(const DevView &V, int row, double *c, int t)
  {
    const double4 g = ld_rec(V.geo, row), a = ld_rec(V.rec[0], row), b = ld_rec(V.rec[1], row), u = ld_rec(V.rec[2], row);
    tm(c, 0, t) = make_double2(g.x, g.y); tm(c, 1, t) = make_double2(g.z, a.x); tm(c, 2, t) = make_double2(a.y, a.z);
    tm(c, 3, t) = make_double2(a.w, b.x); tm(c, 4, t) = make_double2(b.y, b.z); tm(c, 5, t) = make_double2(u.x, u.y);
    tm(c, 6, t) = make_double2(u.z, u.w);
    c[14 * MCAP + t] = g.w;
  }
  // one listed pair: right-hand sides into a[0..3], the phase-density sums of smooth_density into a[4], a[5]
  // (density_massfrac.cc:159-161; same pairs, same symmetric-h weight); near: the pair lies inside the support of the
  // energy-smoothing kernel W(q dx / h, h / q) (:163-168), decided exactly as DensStage::pair_inner<false> does
  static __device__ __forceinline__ void mix_pair(const KParams &K, const I &s, double *a, const double *c, int sl, double gx, double gy,
                                                  double gz, bool &near)
  {
    const double2 p0 = tm(c, 0, sl), p1 = tm(c, 1, sl), p2 = tm(c, 2, sl), p3 = tm(c, 3, sl), p4 = tm(c, 4, sl), p5 = tm(c, 5, sl),
                  p6 = tm(c, 6, sl);
    const double hj = c[14 * MCAP + sl];
    const double dx = gx - p0.x, dy = gy - p0.y, dz = gz - p1.x;
    R r;
    r.r0 = make_double4(p1.y, p2.x, p2.y, p3.x);
    r.r1 = make_double4(p3.y, p4.x, p4.y, 0.);
    r.r2 = make_double4(p5.x, p5.y, p6.x, p6.y);
    const double w = pair_w<false>(K, s, *reinterpret_cast<A *>(a), dx, dy, dz, hj, r);
    const double md = p6.y, wm = w * fabs(md);
    a[5] += md < 0.0 ? wm : 0.0;
    a[4] += md < 0.0 ? 0.0 : wm;
    if (K.energy_smooth == 1) {
      const double h = 0.5 * (s.h + hj), q = K.e_ratio;
      // all |q dx_k / h| <= cutoff implies q^2 r2 <= DIM (cutoff h)^2: one compare clears most pairs (K.near_c carries
      // the constants and a 2^-39 margin)
      if ((dx * dx + dy * dy + dz * dz) * K.near_c <= h * h) near = !outside_kernel<DIM>(q * dx, q * dy, q * dz, h, K.cutoff);
    }
  }
  static __device__ __forceinline__ void mix_finish(const DevView &V, const KParams &K, StepState *ss, int i, const I &s, const double *t)
  {
    finish(V, K, ss, i, s, t[0], t[1], t[2], t[3]);
    V.scr[4][i] = t[4]; V.scr[5][i] = t[5];                         // k_dens_near<false> adds the row's own term
  }
};

// second pass of mom_engr_update (:412-418): commit + equation of state
__global__ void k_commit_mom(DevView V, KParams K, StepState *ss)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ss->n || !(V.flags[p] & F_NEED)) return;
  V.t.d[D_VX][p] = V.scr[0][p];
  if (K.dim == 3) { V.t.d[D_VY][p] = V.scr[1][p]; V.t.d[D_VZ][p] = V.scr[2][p]; }
  const double e = V.scr[3][p];
  V.t.d[D_E][p] = e;
  EosOut o = dev_eos(K, V.t.d[D_RHO][p], V.t.d[D_KS][p], e, V.t.d[D_C][p], V.t.d[D_CP][p], V.t.d[D_GAMMA][p]);
  V.t.d[D_P][p] = o.p; V.t.d[D_GAMMA][p] = o.gamma; V.t.d[D_C][p] = o.c; V.t.d[D_T][p] = o.T; V.t.d[D_CP][p] = o.Cp;
  V.t.d[D_RPH0][p] = o.rph0; V.t.d[D_RPH1][p] = o.rph1;
}

// =========================================================================================
// smooth_velocity  (src/plume/smooth_velocity.cc:25-165, HAVE_TURBULENCE_LANS == 1)
// =========================================================================================
template <int DIM>
struct VelStage
{
  static constexpr bool kSymH = false, kSkipSelf = false, kCountSupport = false;
  static constexpr int kRows = F_NEED, kWarps = 24, kWarpsGrp = 20, kPfGrp = 1, kWarpsScan = 24, kWarpsList = 24, kPfList = 1, kNRec = 1;
  struct I { double h, invh, supp, hk, v[3], ninvh2, norm; };
  struct A { double dv[3], wn; };
  static __device__ __forceinline__ double reach_h(const KParams &K, double h, double) { return K.cutoff * h; }
  static __device__ __forceinline__ void init(const DevView &V, const KParams &K, StepState *, int i, I &s)
  {
    s.h = V.t.d[D_H][i]; s.hk = s.h; s.invh = 1.0 / s.h; s.supp = K.cutoff2 * s.h;
    s.ninvh2 = -(s.invh * s.invh); s.norm = dev_weight_norm<DIM>(s.invh);
    s.v[0] = V.t.d[D_VX][i]; s.v[1] = V.t.d[D_VY][i]; s.v[2] = V.t.d[D_VZ][i];
  }
  static __device__ __forceinline__ double reach(const KParams &K, const I &s, double) { return K.cutoff * s.h; }
  static __device__ __forceinline__ void begin(const KParams &, I &, A &a, int) { a.dv[0] = a.dv[1] = a.dv[2] = a.wn = 0.0; }
  struct R { double4 r0; };                                          // the stage record of row j
  static __device__ __forceinline__ R load(const DevView &V, int j) { R r; r.r0 = ld_rec(V.rec[0], j); return r; }
  static __device__ __forceinline__ void zero(R &r) { r.r0 = make_double4(0., 0., 0., 0.); }  // a row that adds exact zeros
  static __device__ __forceinline__ void pair(const DevView &V, const KParams &K, const I &s, A &a, double dx, double dy,
                                              double dz, double hj, int j)
  { pair_r<false>(K, s, a, dx, dy, dz, hj, load(V, j)); }
  static constexpr int kNC = 7, kWarpsTile = 16, kMinbTile = 2;
  static constexpr bool kDefer = false;
  // pair-weight cache (psph_tile.cuh): with W read from the cache the stage needs no positions at all — the tile is
  // the four columns m/rho, v of the stage record
  static constexpr int kWCache = 2, kNCc = 4;
  static constexpr bool kCachedNeedsX = false, kSkipUnlessFull = false;
  // ... as two arrays of 16-byte elements {m/rho, vx}, {vy, vz}: the slots of a list come in runs of consecutive
  // rows (the z-run of a subcell column), so the 8 lanes of a quarter warp read 8 consecutive 16-byte elements —
  // conflict-free 128-bit loads where 64-bit loads of plain columns measured 2.5 wavefronts over the minimum
  static __device__ __forceinline__ void tile_load_c(const DevView &V, int row, double *c, int t)
  {
    const double4 r = ld_rec(V.rec[0], row);
    reinterpret_cast<double2 *>(c)[t] = make_double2(r.x, r.y);
    reinterpret_cast<double2 *>(c)[TCAP + t] = make_double2(r.z, r.w);
  }
  static __device__ __forceinline__ void tile_init_c(const DevView &, const KParams &, StepState *, const double *c, int t, int, I &s)
  {
    const double2 a = reinterpret_cast<const double2 *>(c)[t], b = reinterpret_cast<const double2 *>(c)[TCAP + t];
    s.v[0] = a.y; s.v[1] = b.x; s.v[2] = b.y;
  }
  static __device__ __forceinline__ void tile_uni(const KParams &K, double hstd, I &s)
  {
    s.h = hstd; s.hk = s.h; s.invh = 1.0 / s.h; s.supp = K.cutoff2 * s.h;
    s.ninvh2 = -(s.invh * s.invh); s.norm = dev_weight_norm<DIM>(s.invh);
  }
  static __device__ __forceinline__ double self_weight(const I &s) { return dev_weight_pre(0.0, s.ninvh2, s.norm); }
  static __device__ __forceinline__ void pair_c(const KParams &, const I &s, A &a, double w, const double *c, int sl, int)
  {
    const double2 p = reinterpret_cast<const double2 *>(c)[sl], q = reinterpret_cast<const double2 *>(c)[TCAP + sl];
    const double mvw = fabs(w) * p.x;                              // the sign is smooth_density's flag
    a.dv[0] += mvw * (p.y - s.v[0]);
    if (DIM == 3) { a.dv[1] += mvw * (q.x - s.v[1]); a.dv[2] += mvw * (q.y - s.v[2]); }
    a.wn += mvw;
  }
  static __device__ __forceinline__ void tile_load(const DevView &V, int row, double *c, int t)
  { const double4 r = ld_rec(V.rec[0], row); c[3 * TCAP + t] = r.x; c[4 * TCAP + t] = r.y; c[5 * TCAP + t] = r.z; c[6 * TCAP + t] = r.w; }
  static __device__ __forceinline__ R tile_rec(const double *c, int t)
  { R r; r.r0 = make_double4(c[3 * TCAP + t], c[4 * TCAP + t], c[5 * TCAP + t], c[6 * TCAP + t]); return r; }
  static __device__ __forceinline__ void tile_stage(const DevView &V, int row, double *c, int t) { tile_stage_soa<VelStage<DIM> >(V, row, c, t); }
  static __device__ __forceinline__ void tile_xyz(const double *c, int t, double &x, double &y, double &z) { tile_xyz_soa(c, t, x, y, z); }
  static __device__ __forceinline__ R tile_recp(const double *c, int t, double &x, double &y, double &z)
  { tile_xyz_soa(c, t, x, y, z); return tile_rec(c, t); }
  static __device__ __forceinline__ void tile_init(const DevView &, const KParams &, StepState *, const double *c, int t, int, I &s)
  { s.v[0] = c[4 * TCAP + t]; s.v[1] = c[5 * TCAP + t]; s.v[2] = c[6 * TCAP + t]; }
  static __device__ __forceinline__ void init_uni(const KParams &, I &) {}
  template <bool UNI>
  static __device__ __forceinline__ void pair_r(const KParams &K, const I &s, A &a, double dx, double dy, double dz, double,
                                                const R &rec)
  {
    const double4 rj = rec.r0;
    const double mvw = dev_weight_pre(dx * dx + dy * dy + dz * dz, s.ninvh2, s.norm, K.cexp) * rj.x;
    a.dv[0] += mvw * (rj.y - s.v[0]);
    if (DIM == 3) { a.dv[1] += mvw * (rj.z - s.v[1]); a.dv[2] += mvw * (rj.w - s.v[2]); }
    a.wn += mvw;
  }
  static __device__ __forceinline__ void finish(const DevView &V, StepState *ss, int i, const I &s, double d0, double d1, double d2, double wn)
  {
    if (!(wn > 0)) atomicAdd(&ss->bad, 1);
    const double rw = 1.0 / wn;
    V.t.d[D_SVX][i] = s.v[0] + d0 * rw;
    if (DIM == 3) { V.t.d[D_SVY][i] = s.v[1] + d1 * rw; V.t.d[D_SVZ][i] = s.v[2] + d2 * rw; }
  }
  static __device__ __forceinline__ bool end(const DevView &V, const KParams &, StepState *ss, int i, I &s, A &a, int, int lane)
  {
    double d0 = a.dv[0], d1 = DIM == 3 ? a.dv[1] : 0.0, d2 = DIM == 3 ? a.dv[2] : 0.0, wn = a.wn;
    warp_sum4(d0, d1, d2, wn, lane);
    if (lane == 0) finish(V, ss, i, s, d0, d1, d2, wn);
    return false;
  }
  // ---- sub-tiled mixed-h gather (psph_mix.cuh): {x, y} {z, m/rho} {vx, vy} as 16-byte elements, then vz; the weight
  // uses the ROW's smoothing length (smooth_velocity.cc:104-118), entries carry the h_i gate tag
  static constexpr int kNCm = 7, kNAm = 4;
  static constexpr unsigned kTagMix = MTAG_HI;
  static __device__ __forceinline__ void mix_stage(const DevView &V, int row, double *c, int t)
  {
    const double4 g = ld_rec(V.geo, row), r = ld_rec(V.rec[0], row);
    double2 *c2 = reinterpret_cast<double2 *>(c);
    c2[t] = make_double2(g.x, g.y); c2[MCAP + t] = make_double2(g.z, r.x); c2[2 * MCAP + t] = make_double2(r.y, r.z);
    c[6 * MCAP + t] = r.w;
  }
  static __device__ __forceinline__ void mix_pair(const KParams &K, const I &s, double *a, const double *c, int sl, double gx, double gy,
                                                  double gz, bool &)
  {
    const double2 *c2 = reinterpret_cast<const double2 *>(c);
    const double2 p0 = c2[sl], p1 = c2[MCAP + sl], p2 = c2[2 * MCAP + sl];
    R r;
    r.r0 = make_double4(p1.y, p2.x, p2.y, c[6 * MCAP + sl]);
    pair_r<false>(K, s, *reinterpret_cast<A *>(a), gx - p0.x, gy - p0.y, gz - p1.x, 0.0, r);
  }
  static __device__ __forceinline__ void mix_finish(const DevView &V, const KParams &, StepState *ss, int i, const I &s, const double *t)
  {
    // the row itself (the lists leave it out): weight W(0, h_i), zero velocity difference
    const double mvw = dev_weight_pre(0.0, s.ninvh2, s.norm) * ld_rec(V.rec[0], i).x;
    finish(V, ss, i, s, t[0], t[1], t[2], t[3] + mvw);
  }
};

// =========================================================================================
// adaptive_sml  (src/plume/neighs.cc:479-693, ADAPTIVE_SML == 1): up to num_loop_P fixed-point
// sweeps h <- eta * (m / rho*)^(1/D), rho* = sum_j m_j W(dx / 2h, 2h).  Reads only x_j, m_j of
// the neighbours, so the in-place update of h_i is order independent.
// =========================================================================================
template <int DIM>
struct SmlStage
{
  static constexpr bool kSymH = false, kSkipSelf = false, kCountSupport = false;
  static constexpr int kRows = F_NEED, kWarps = 24;
  struct I { double h, hs, hk, invhs, supp, m; };
  struct A { double rho; };
  static __device__ __forceinline__ void init(const DevView &V, const KParams &K, StepState *, int i, I &s)
  {
    s.h = V.t.d[D_H][i]; s.m = V.t.d[D_MASS][i];
  }
  static __device__ __forceinline__ double reach(const KParams &K, const I &s, double) { return K.cutoff * s.hs; }
  static __device__ __forceinline__ void begin(const KParams &K, I &s, A &a, int)
  {
    s.hs = s.h * K.c_smooth; s.hk = s.hs; s.invhs = 1.0 / s.hs; s.supp = K.cutoff2 * s.hs;
    a.rho = 0.0;
  }
  static __device__ __forceinline__ void pair(const DevView &V, const KParams &K, const I &s, A &a, double dx, double dy,
                                              double dz, double, int j)
  {
    a.rho += dev_weight_in<DIM>(dx * dx + dy * dy + dz * dz, s.invhs) * __ldg(&V.scr[7][j]);
  }
  static __device__ __forceinline__ bool end(const DevView &V, const KParams &K, StepState *, int i, I &s, A &a, int pass, int lane)
  {
    const double rho = warp_sum(a.rho);
    const double hn = DIM == 3 ? K.eta_smooth * cbrt(s.m / rho) : K.eta_smooth * (s.m / rho);
    const double err = fabs((hn - s.h) / s.h);
    s.h = hn;
    const bool again = !(err < K.thresh) && pass + 1 < K.num_loop;
    if (!again && lane == 0) V.t.d[D_H][i] = hn;
    return again;
  }
};

// =========================================================================================
// update_pos  (src/plume/update_pos.cc:35-168)
// =========================================================================================
__global__ void k_update_pos(Tab V, KParams K, StepState *ss)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ss->n) return;
  const double dt = ss->dt;
  const int bc = V.i[I_BC][p];
  if (dev_is_real(bc)) {
    const int u = K.lans ? D_SVX : D_VX;
    for (int k = 0; k < K.dim; k++) V.d[D_X + k][p] = V.d[D_X + k][p] + dt * V.d[u + k][p];
  } else if (bc == 0 && ss->erupting) {
    for (int k = 0; k < K.dim; k++) V.d[D_X + k][p] = V.d[D_X + k][p] + dt * V.d[D_VX + k][p];
    if (K.dim == 3 && V.d[D_Z][p] >= K.erupt_top) {                 // :134-142
      V.i[I_BC][p] = 100;
      V.d[D_H][p] = K.smlen_real;
      V.i[I_INVOLVED][p] = 2;
    }
  }
}

// =========================================================================================
// host side of the stages
// =========================================================================================
#define GRID(n) (unsigned) (((n) + 255) / 256), 256, 0, ctx->stream

// 3-D SPH stages run on the grouped gather (psph_gather3.cuh); 1-D, and PSPH_GATHER=row, on the per-row one
// tile gather of one stage: the stage's default shape, or PSPH_TWARPS = warps * 10 + CTAs per SM (tuning switch)
template <class S>
static int launch_tile_gather_any(psph_ctx *ctx, int kid)
{
  switch (ctx->tile_shape) {
  case 121: return launch_tile_gather<S, 12, 1>(ctx, kid);
  case 161: return launch_tile_gather<S, 16, 1>(ctx, kid);
  case 201: return launch_tile_gather<S, 20, 1>(ctx, kid);
  case 241: return launch_tile_gather<S, 24, 1>(ctx, kid);
  case 122: if constexpr (S::kNC <= 7) return launch_tile_gather<S, 12, 2>(ctx, kid); else break;
  case 162: if constexpr (S::kNC <= 7) return launch_tile_gather<S, 16, 2>(ctx, kid); else break;
  default: break;
  }
  return launch_tile_gather<S, S::kWarpsTile, S::kMinbTile>(ctx, kid);
}

// The first summation stage of a step (mom_engr_update, `writer`) finds the pairs: the grouped gather publishes the
// pair lists of the rows that qualify (and sums the others in place), then every stage runs the rows that have a
// list on k_gather_list; the later stages send the rest through the grouped gather again.
static bool tile_modes(const psph_ctx *ctx) { return ctx->nbr_mode == 5 || ctx->nbr_mode == 2 || ctx->nbr_mode == 4; }

template <template <int> class S>
static int launch_stage_gather(psph_ctx *ctx, int kid, bool writer)
{
  if (ctx->cfg.dim != 3 || ctx->gather_mode != 0) return launch_gather<S>(ctx, kid);
  if (writer) {
    ctx->nbr_valid = false;
    if (tile_modes(ctx) && psph_nbr_prepare(ctx) == 1) {
      // Slot lists + records staged in shared memory.  Which of the two layouts runs is the device's choice
      // (StepState::pair_mode, k_pair_mode): the kernels of the other one return at once, so no host round trip
      {
        KernelTimer kt(ctx, KN_PAIR_SCAN);
        if (ctx->scan_mode == 2) { // one lane per row (psph_scan2.cuh); PSPH_SCAN=1: one lane per candidate (psph_tile.cuh)
          if (launch_tile_scan2<S<3>, 20, 16>(ctx, -1)) return -1;
        } else if ((ctx->gather_warps == 24 ? launch_tile_scan<S<3>, 24>(ctx, -1) : launch_tile_scan<S<3>, 20>(ctx, -1))) return -1;
        if (launch_mix_scan<16>(ctx, -1)) return -1;
        k_no_lists<<<GRID(ctx->n)>>>(ctx->cur, ctx->d_ss);
        PSPH_LAUNCH_CHECK();
      }
      ctx->nbr_valid = true; ctx->nbr_kind = 2;
      KernelTimer kt(ctx, kid);
      if (launch_tile_gather_any<S<3> >(ctx, -1)) return -1;
      if constexpr (S<3>::kNAm == 6) {                               // (the writer is mom_engr_update)
        if (ctx->mix_warps == 24 ? launch_mix_gather<S<3>, 24, 1>(ctx, -1)
            : ctx->mix_warps == 20 ? launch_mix_gather<S<3>, 20, 1>(ctx, -1) : launch_mix_gather<S<3>, 16, 1>(ctx, -1)) return -1;
      }
      return launch_gather_grp<S<3> >(ctx, -1, 2);
    }
    if (ctx->nbr_mode == 3 && psph_nbr_prepare(ctx) == 1) { // general row lists: any smoothing lengths, tagged entries
      if (launch_gather_grp<S<3> >(ctx, KN_PAIR_SCAN, 4)) return -1;
      ctx->nbr_valid = true; ctx->nbr_kind = 3;
      KernelTimer kt(ctx, kid);
      if (launch_gather_list<S<3> >(ctx, -1)) return -1;
      return launch_gather_grp<S<3> >(ctx, -1, 2);
    }
    if (ctx->nbr_mode == 1 && psph_nbr_prepare(ctx) == 1) {
      if (launch_gather_grp<S<3> >(ctx, KN_PAIR_SCAN, 3)) return -1; // timed apart: the pair search
      ctx->nbr_valid = true; ctx->nbr_kind = 1;
      KernelTimer kt(ctx, kid); // both launches are this stage's gather
      if (launch_gather_list<S<3> >(ctx, -1)) return -1;
      return launch_gather_grp<S<3> >(ctx, -1, 2);
    }
    return launch_gather_grp<S<3> >(ctx, kid, 0);
  }
  if (ctx->nbr_valid) {
    KernelTimer kt(ctx, kid); // every launch is this stage's gather
    if (ctx->nbr_kind == 2) {
      if constexpr (S<3>::kSkipUnlessFull) { // smooth_density: the sums mom_engr_update's pass prepared + the near lists
        if (ctx->n > 0) {
          if (ctx->cur.wts) { k_dens_near<true><<<GRID(ctx->n)>>>(ctx->cur, ctx->kp, ctx->d_ss); PSPH_LAUNCH_CHECK(); }
          k_dens_near<false><<<GRID(ctx->n)>>>(ctx->cur, ctx->kp, ctx->d_ss);
          PSPH_LAUNCH_CHECK();
          if (launch_gather_d<3, S<3>, true>(ctx, -1)) return -1;     // rows whose near list overflowed
        }
        if (launch_tile_gather_any<S<3> >(ctx, -1)) return -1;
      } else {
        if (launch_tile_gather_any<S<3> >(ctx, -1)) return -1;
        if constexpr (S<3>::kNAm == 4) { if (launch_mix_gather<S<3>, 16, 2>(ctx, -1)) return -1; }
      }
    } else if (launch_gather_list<S<3> >(ctx, -1)) return -1;
    return launch_gather_grp<S<3> >(ctx, -1, 2);
  }
  return launch_gather_grp<S<3> >(ctx, kid, 0);
}

// room for the pair lists of this step: 0 none (cache off for the step), 1 ready; < 0 error
__global__ void k_nbr_reset(StepState *ss, unsigned cap, int dens_full, int mix_rows)
{ ss->nbr_cursor = 0u; ss->nbr_cap = cap; ss->n_uncached = 0; ss->dens_full = dens_full; ss->n_near_ovf = 0; ss->mix_rows = mix_rows; }
int psph_nbr_prepare(psph_ctx *ctx)
{
  // 320 pairs per row cover the unit lattice (~230-300 pairs): 10 units of 128 bytes (row indices) or 5 (16-bit
  // tile slots); rows that do not fit get no list.  Allocated once per table size: when memory capped the buffer
  // below `want`, asking again every step would only cost a free + malloc of tens of GB.
  const bool tiles = tile_modes(ctx);
  size_t want = (size_t) ctx->n * (tiles ? 5 : (ctx->nbr_mode == 3 ? 26 : 10)); // mode 3: ~730 partners
  if (want > 0x7ffffff0ull) want = 0x7ffffff0ull;
  if (ctx->nbr_units < want && want > ctx->nbr_want_tried) {
    ctx->nbr_want_tried = want + want / 8;
    size_t free_b = 0, total_b = 0;
    PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->cur.nbr); ctx->cur.nbr = nullptr; ctx->nbr_units = 0;
    cudaFree(ctx->cur.wts); ctx->cur.wts = nullptr;
    PSPH_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const size_t keep = (size_t) 2 << 30;                       // leave 2 GiB for everything else
    size_t units = free_b > keep ? (free_b - keep) / 128 : 0;
    units = std::min(units, want + want / 8);
    if (units < (size_t) ctx->n) return 0;                       // not worth it
    if (cudaMalloc(&ctx->cur.nbr, units * 128) != cudaSuccess) { cudaGetLastError(); ctx->cur.nbr = nullptr; return 0; }
    ctx->nbr_units = units;
    // pair-weight cache of the block-tiled mode: one double per 16-bit slot (4 x the list bytes), only when the full
    // list fits and the memory is there — HBM is idle during the pair sums, the FP64 pipe is not
    if (tiles && ctx->wcache && units >= want) {
      PSPH_CUDA(cudaMemGetInfo(&free_b, &total_b));
      if (free_b > keep + units * 512 && cudaMalloc(&ctx->cur.wts, units * 512) != cudaSuccess) { cudaGetLastError(); ctx->cur.wts = nullptr; }
    }
  }
  if (!tiles && ctx->cur.wts) { cudaFree(ctx->cur.wts); ctx->cur.wts = nullptr; }
  // mixed-h lists (psph_mix.cuh): MSTRIDE 16-bit entries per row at a fixed stride + one int per row; allocated once the
  // table is known to carry per-row smoothing lengths (adaptive_sml ran, or a host read-back of the step state said so)
  if (tiles && !ctx->mix_wanted && ctx->h_unknown) {
    // rows that came in from outside may carry per-row smoothing lengths already: one look at the path the device
    // picked for this list (only after an upload; a running loop learns it from adaptive_sml)
    PSPH_CUDA(cudaMemcpyAsync(&ctx->h_ss->pair_mode, &ctx->d_ss->pair_mode, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_ss->pair_mode == 4) ctx->mix_wanted = true;
  }
  ctx->h_unknown = false;
  if (tiles && ctx->mix_wanted && ctx->mix_rows < (size_t) ctx->n) {
    PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->cur.mlist); cudaFree(ctx->cur.mseg); ctx->cur.mlist = nullptr; ctx->cur.mseg = nullptr; ctx->mix_rows = 0;
    size_t free_b = 0, total_b = 0;
    PSPH_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const size_t keep = (size_t) 2 << 30, per_row = (size_t) MSTRIDE * 2 + 4;
    size_t rows = std::max(ctx->cap, (size_t) ctx->n);
    if (free_b < keep + rows * per_row) rows = free_b > keep ? (free_b - keep) / per_row : 0;  // the rows beyond get no list
    if (rows > 0 && cudaMalloc(&ctx->cur.mlist, rows * MSTRIDE * 2) == cudaSuccess && cudaMalloc(&ctx->cur.mseg, rows * 4) == cudaSuccess)
      ctx->mix_rows = rows;
    else { cudaGetLastError(); cudaFree(ctx->cur.mlist); ctx->cur.mlist = nullptr; }
  }
  // units of 32 entries: 128 bytes of row indices, or 64 bytes of 16-bit tile slots (block-tiled mode)
  const size_t cap = tiles ? ctx->nbr_units * 2 : ctx->nbr_units;  // (the writer that follows uses nbr_mode)
  k_nbr_reset<<<1, 1, 0, ctx->stream>>>(ctx->d_ss, (unsigned) std::min<size_t>(cap, 0xfffffff0ull), ctx->dens_near ? 0 : 1,
                                        (int) std::min<size_t>(ctx->mix_rows, 0x7fffffff));
  PSPH_LAUNCH_CHECK();
  return 1;
}

static int ensure_list(psph_ctx *ctx)
{
  if (ctx->list_valid) return 0;
  return psph_build_cell_list(ctx);
}

int psph_pair_mode(psph_ctx *ctx)
{
  const int forced = ctx->nbr_mode == 5 ? 0 : (ctx->nbr_mode == 2 || ctx->nbr_mode == 4 ? ctx->nbr_mode : -1);
  k_pair_mode<<<1, 1, 0, ctx->stream>>>(ctx->d_ss, forced, ctx->dens_near);
  PSPH_LAUNCH_CHECK();
  return 0;
}

int psph_refresh_geo(psph_ctx *ctx)
{
  k_hmax_reset<<<1, 1, 0, ctx->stream>>>(ctx->d_ss);
  PSPH_LAUNCH_CHECK();
  if (ctx->n > 0) { k_hmax_scan<<<GRID(ctx->n)>>>(ctx->cur.t, ctx->d_ss); PSPH_LAUNCH_CHECK(); }
  k_hmax_finish<<<1, 1, 0, ctx->stream>>>(ctx->d_ss);
  PSPH_LAUNCH_CHECK();
  PSPH_CUDA(cudaMemsetAsync(ctx->d_cell_hmax, 0, ctx->ncells * sizeof(unsigned long long), ctx->stream));
  if (ctx->n > 0) { k_refresh_geo<<<GRID(ctx->n)>>>(ctx->cur, ctx->kp, ctx->d_ss, ctx->d_cell_hmax); PSPH_LAUNCH_CHECK(); }
  if (psph_pair_mode(ctx)) return -1;
  ctx->nbr_valid = false; // smoothing lengths changed: the pair lists of this step are stale
  return 0;
}

extern "C" int psph_timestep(psph_ctx *ctx, double *dt)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  StageTimer tm(ctx, ST_TIMESTEP);
  k_timestep_reset<<<1, 1, 0, ctx->stream>>>(ctx->d_ss);
  PSPH_LAUNCH_CHECK();
  if (ctx->n > 0) {
    int grid = std::min((ctx->n + 255) / 256, ctx->sms * 8);
    k_timestep<<<grid, 256, 0, ctx->stream>>>(ctx->cur.t, ctx->d_ss);
    PSPH_LAUNCH_CHECK();
  }
  if (ctx->cfg.nranks > 1 && psph_comm_min_dt(ctx)) return -1;
  // the unclipped CFL dt goes through the pinned mirror's dt field
  k_timestep_finish<<<1, 1, 0, ctx->stream>>>(ctx->kp, ctx->d_ss, nullptr);
  PSPH_LAUNCH_CHECK();
  if (dt) {
    PSPH_CUDA(cudaMemcpyAsync(ctx->h_ss, ctx->d_ss, sizeof(StepState), cudaMemcpyDeviceToHost, ctx->stream));
    PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
    *dt = ctx->kp.cfl * __builtin_bit_cast(double, ctx->h_ss->min_hc_bits);
  }
  return 0;
}

extern "C" int psph_set_dt(psph_ctx *ctx, double dt, int step, int erupting)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  ctx->step = step;
  struct { double dt; } v = { dt };
  PSPH_CUDA(cudaMemcpyAsync(&ctx->d_ss->dt, &v.dt, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  int iv[2] = { erupting, step };
  PSPH_CUDA(cudaMemcpyAsync(&ctx->d_ss->erupting, iv, 2 * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int psph_search_neighs(psph_ctx *ctx)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  StageTimer tm(ctx, ST_SEARCH);
  if (ctx->cfg.dim == 1 && ctx->list_valid) return psph_refresh_geo(ctx); // 1-D: no mesh, no search (main.cc:676-751)
  return psph_build_cell_list(ctx);
}

extern "C" int psph_adaptive_sml(psph_ctx *ctx)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  StageTimer tm(ctx, ST_SML);
  if (ensure_list(ctx)) return -1;
  if (ctx->n == 0) { psph_comm_mark(ctx, DIRTY_H); return 0; }
  k_pack_sml<<<GRID(ctx->n)>>>(ctx->cur, ctx->d_ss);
  PSPH_LAUNCH_CHECK();
  if (launch_gather<SmlStage>(ctx, KN_GATHER_SML)) return -1;
  ctx->mix_wanted = true;                                   // every row has its own smoothing length from here on
  psph_comm_mark(ctx, DIRTY_H);
  return psph_refresh_geo(ctx);
}

// second pass of mom_engr_update: commit the scratch columns + equation of state (both variants)
int psph_commit_mom(psph_ctx *ctx)
{
  k_commit_mom<<<GRID(ctx->n)>>>(ctx->cur, ctx->kp, ctx->d_ss);
  PSPH_LAUNCH_CHECK();
  psph_comm_mark(ctx, DIRTY_MOM);
  return 0;
}

extern "C" int psph_mom_engr_update(psph_ctx *ctx)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  if (ctx->cfg.use_gsph) return psph_gsph_mom_engr(ctx);
  StageTimer tm(ctx, ST_MOM);
  if (ensure_list(ctx)) return -1;
  // a rank that holds no rows right now still takes part in the exchange that follows (the dirty mask decides which
  // collective every rank enters: it must not depend on the local row count)
  if (ctx->n == 0) { psph_comm_mark(ctx, DIRTY_MOM); return 0; }
  k_pack_mom<<<GRID(ctx->n)>>>(ctx->cur, ctx->kp, ctx->d_ss);
  PSPH_LAUNCH_CHECK();
  if (launch_stage_gather<MomStage>(ctx, KN_GATHER_MOM, true)) return -1;
  return psph_commit_mom(ctx);
}

extern "C" int psph_smooth_density(psph_ctx *ctx)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  StageTimer tm(ctx, ST_DENS);
  if (ensure_list(ctx)) return -1;
  if (ctx->n == 0) { psph_comm_mark(ctx, DIRTY_DENS); return 0; }
  k_pack_dens<<<GRID(ctx->n)>>>(ctx->cur, ctx->d_ss);
  PSPH_LAUNCH_CHECK();
  if (launch_stage_gather<DensStage>(ctx, KN_GATHER_DENS, false)) return -1;
  k_commit_dens<<<GRID(ctx->n)>>>(ctx->cur, ctx->kp, ctx->d_ss);
  PSPH_LAUNCH_CHECK();
  psph_comm_mark(ctx, DIRTY_DENS);
  return 0;
}

extern "C" int psph_smooth_velocity(psph_ctx *ctx)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  StageTimer tm(ctx, ST_VEL);
  if (ensure_list(ctx)) return -1;
  psph_comm_mark(ctx, DIRTY_VEL);
  if (ctx->n == 0) return 0;
  k_pack_vel<<<GRID(ctx->n)>>>(ctx->cur, ctx->d_ss);
  PSPH_LAUNCH_CHECK();
  return launch_stage_gather<VelStage>(ctx, KN_GATHER_VEL, false);
}

extern "C" int psph_update_pos(psph_ctx *ctx)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  StageTimer tm(ctx, ST_POS);
  psph_comm_mark(ctx, DIRTY_LAYOUT);
  if (ctx->n == 0) return 0;
  k_update_pos<<<GRID(ctx->n)>>>(ctx->cur.t, ctx->kp, ctx->d_ss);
  PSPH_LAUNCH_CHECK();
  return 0;
}

extern "C" int psph_rebin(psph_ctx *ctx, int *lost)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  {
    StageTimer tm(ctx, ST_REBIN);
    if (psph_rebin_rows(ctx)) return -1;
  }
  if (lost) {
    PSPH_CUDA(cudaMemcpyAsync(ctx->h_ss, ctx->d_ss, sizeof(StepState), cudaMemcpyDeviceToHost, ctx->stream));
    PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
    *lost = ctx->h_ss->lost;
  }
  return 0;
}

// one step in the reference's order (main.cc:217-408 3-D; :676-751 1-D), no host round trip
static int step_async(psph_ctx *ctx)
{
  const psph_config &c = ctx->cfg;
  const bool multi = c.nranks > 1;
  ctx->step++;
  if (multi && psph_exchange_ghosts(ctx)) return -1;               // pending migration / ghost rebuild
  if (psph_timestep(ctx, nullptr)) return -1;
  if (psph_search_neighs(ctx)) return -1;
  if (c.adaptive_sml && (c.dim == 1 || ctx->step % c.sml_update_int == 0)) {
    if (psph_adaptive_sml(ctx)) return -1;
    if (multi && psph_exchange_ghosts(ctx)) return -1;             // main.cc:326
  }
  if (c.use_gsph) {
    if (psph_calc_gradients(ctx)) return -1;
    if (multi && psph_exchange_ghosts(ctx)) return -1;             // main.cc:347
  }
  if (psph_mom_engr_update(ctx)) return -1;
  if (multi && psph_exchange_ghosts(ctx)) return -1;               // main.cc:374
  if (psph_smooth_density(ctx)) return -1;
  if (multi && psph_exchange_ghosts(ctx)) return -1;               // main.cc:387
  if (c.have_turbulence_lans) {
    if (psph_smooth_velocity(ctx)) return -1;
    if (multi && psph_exchange_ghosts(ctx)) return -1;             // main.cc:395
  }
  if (psph_update_pos(ctx)) return -1;
  if (psph_rebin(ctx, nullptr)) return -1;
  if (multi && psph_exchange_ghosts(ctx)) return -1;               // main.cc:430
  return 0;
}

static int set_step_inputs(psph_ctx *ctx, double dt_max, int erupting)
{
  double v = dt_max;
  PSPH_CUDA(cudaMemcpyAsync(&ctx->d_ss->dt_max, &v, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  int e = erupting;
  PSPH_CUDA(cudaMemcpyAsync(&ctx->d_ss->erupting, &e, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  PSPH_CUDA(cudaStreamSynchronize(ctx->stream)); // v, e are stack variables
  return 0;
}

extern "C" int psph_errors(psph_ctx *ctx, int reset)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  PSPH_CUDA(cudaMemcpyAsync(ctx->h_ss, ctx->d_ss, sizeof(StepState), cudaMemcpyDeviceToHost, ctx->stream));
  PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
  int bad = ctx->h_ss->bad;
  ctx->n = ctx->list_valid ? ctx->h_ss->n : ctx->n;
  // the block tiles assume a reach of 3 subcells and a bounded row count per cell: when most rows fell back to the
  // fused gather in the last step, the row lists (no such assumptions) serve this table better
  // ... and when the row lists do not apply either (smoothing lengths differ from row to row, e.g. after
  // adaptive_sml), the search that writes them is wasted work: the fused grouped gather alone serves every table
  // (the default, nbr_mode 5, needs none of this: the device picks block tiles or the sub-tiled mixed-h gather per step)
  if (ctx->nbr_mode > 0 && ctx->nbr_mode < 4 && ctx->nbr_valid && ctx->h_ss->n_need > 0 && 2LL * ctx->h_ss->n_uncached > ctx->h_ss->n_need)
    ctx->nbr_mode = ctx->nbr_mode == 2 ? 1 : 0; // tiles -> uniform-h row lists -> fused (the general row lists, mode 3,
                                                // measured no faster than the fused gather: opt-in only)
  if (ctx->h_ss->pair_mode == 4) ctx->mix_wanted = true;
  if (reset && bad) {
    PSPH_CUDA(cudaMemsetAsync(&ctx->d_ss->bad, 0, sizeof(int), ctx->stream));
    PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return bad;
}

// pair path of the last step as the device saw it (diagnostics of bench.py / the tests); synchronises
extern "C" int psph_pair_stats(psph_ctx *ctx, int32_t out[4])
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  PSPH_CUDA(cudaMemcpyAsync(ctx->h_ss, ctx->d_ss, sizeof(StepState), cudaMemcpyDeviceToHost, ctx->stream));
  PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
  out[0] = ctx->h_ss->pair_mode; out[1] = ctx->h_ss->n_need; out[2] = ctx->h_ss->n_uncached; out[3] = ctx->h_ss->n_near_ovf;
  return 0;
}

// rows the gathers update (need_neigh and the list rule), exact after a list build; synchronises
extern "C" int64_t psph_num_updated(psph_ctx *ctx)
{
  if (cudaSetDevice(ctx->cfg.device) != cudaSuccess) return -1;
  ctx->n_dirty = true;
  if (psph_sync_n(ctx)) return -1;
  return ctx->h_ss->n_need;
}

extern "C" int psph_step(psph_ctx *ctx, double dt_max, int erupting, double *dt_used, int *lost)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  if (set_step_inputs(ctx, dt_max > 0 ? dt_max : 1.0e300, erupting)) return -1;
  if (step_async(ctx)) return -1;
  int bad = psph_errors(ctx, 1);
  if (bad < 0) return bad;
  if (dt_used) *dt_used = ctx->h_ss->dt;
  if (lost) *lost = ctx->h_ss->lost;
  return bad;
}

extern "C" int psph_run(psph_ctx *ctx, int nsteps, int erupting)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  if (set_step_inputs(ctx, 1.0e300, erupting)) return -1;
  for (int s = 0; s < nsteps; s++)
    if (step_async(ctx)) return -1;
  return 0;
}

extern "C" int psph_set_timing(psph_ctx *ctx, int on)
{
  ctx->timing = on;
  memset(ctx->stage_ms, 0, sizeof ctx->stage_ms);
  memset(ctx->kernel_ms, 0, sizeof ctx->kernel_ms);
  memset(ctx->kernel_calls, 0, sizeof ctx->kernel_calls);
  return 0;
}
extern "C" double psph_kernel_ms(psph_ctx *ctx, int kernel, int reset, int64_t *calls)
{
  if (kernel < 0 || kernel >= KN_COUNT) return -1.0;
  double v = ctx->kernel_ms[kernel];
  if (calls) *calls = ctx->kernel_calls[kernel];
  if (reset) { ctx->kernel_ms[kernel] = 0; ctx->kernel_calls[kernel] = 0; }
  return v;
}
extern "C" double psph_stage_ms(psph_ctx *ctx, int stage, int reset)
{
  if (stage < 0 || stage >= ST_COUNT) return -1.0;
  double v = ctx->stage_ms[stage];
  if (reset) ctx->stage_ms[stage] = 0;
  return v;
}
