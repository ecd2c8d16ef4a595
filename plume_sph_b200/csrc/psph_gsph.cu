// psph_gsph.cu — Godunov-SPH variant: calc_gradients and the Riemann-solver momentum/energy update.
//
// Replaces  calc_gradients            src/plume/calculate_gradient.cc:22-209
//           mom_engr_update (GSPH)    src/plume/mom_engr_update.cc:443-1145
//           Riemann_Solver            src/plume/sph_lib.cc:2247-2600 (HLLC :1753-1908, Roe :1694-1749)
// for RP_MASS_WEIGHTED 0, SWITCH_OFF_AV_FOR_EXPAN 0, SHEAR_VEL_APP 2 (Roe-averaged shear velocity), HLL_WAVE_SPEED_EVA 0,
// ideal-gas sound speeds, and both option sets of the Riemann states (psph_config.gsph_reconstruct):
//   0  the 1-D defaults (options.h:141-262): GSPH_SPECIFIC_VOL_APP 0, MINI_DERIVATIVE_COND 4 — piecewise-constant states, the
//      gradients are computed and stored like the reference does but do not enter them (sph_lib.cc:2487-2493);
//   3  the 3-D defaults (options.h:544-663): states reconstructed from the gradients, GSPH_SPECIFIC_VOL_APP 3 (cubic interface
//      position), MINI_DERIVATIVE_COND 10, GSPH_MODIFIED_MONOTONICITY 1 — normally with the Roe solver.
// Both stages run on the warp gather of psph_gather.cuh.
#include "psph_gather.cuh"

// ---- stage records -----------------------------------------------------------------------
// gradients: rec0 = {m/rho, rho, p, -}  rec1 = {v_x, v_y, v_z, -}   (state velocity, calculate_gradient.cc:143)
__global__ void k_pack_grad(DevView V, StepState *ss)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ss->n) return;
  const double rho = V.t.d[D_RHO][p];
  V.rec[0][p] = make_double4(V.t.d[D_MASS][p] / rho, rho, V.t.d[D_P][p], 0.0);
  V.rec[1][p] = make_double4(V.t.d[D_VX][p], V.t.d[D_VY][p], V.t.d[D_VZ][p], 0.0);
}
// momentum/energy: rec0 = {m, rho, p, c}  rec1 = {gamma, Cp, T, -}  rec2 = {u_x, u_y, u_z, -}
__global__ void k_pack_gsph(DevView V, KParams K, StepState *ss)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ss->n) return;
  V.rec[0][p] = make_double4(V.t.d[D_MASS][p], V.t.d[D_RHO][p], V.t.d[D_P][p], V.t.d[D_C][p]);
  V.rec[1][p] = make_double4(V.t.d[D_GAMMA][p], V.t.d[D_CP][p], V.t.d[D_T][p], 0.0);
  const int u = K.lans ? D_SVX : D_VX;
  V.rec[2][p] = make_double4(V.t.d[u][p], V.t.d[u + 1][p], V.t.d[u + 2][p], 0.0);
}

// =========================================================================================
// calc_gradients: normalised SPH gradients of rho, v, p through the moment matrix
//   A[a][b] = - sum_j dx_a (m_j/rho_j) dW_b,   F[a][q] = - sum_j (q_i - q_j) (m_j/rho_j) dW_a,   grad q = A^-1 F
// with h = h_i; zero when fewer than 21 listed neighbours lie inside CUTOFF2*h_i (:162).
// =========================================================================================
template <int DIM>
struct GradStage
{
  static constexpr bool kSymH = false, kSkipSelf = true, kCountSupport = true;
  static constexpr int kRows = F_GRADROW, kWarps = 16;
  static constexpr int NQ = DIM + 2, NA = DIM == 3 ? 6 : 1;
  struct I { int i; double h, supp, hk, invh, rho, p, v[3]; };
  struct A { int nreal; double aa[NA]; double f[DIM][NQ]; };
  static __device__ __forceinline__ void init(const DevView &V, const KParams &K, StepState *, int i, I &s)
  {
    s.i = i;
    s.h = V.t.d[D_H][i]; s.hk = s.h; s.invh = 1.0 / s.h; s.supp = K.cutoff2 * s.h;
    s.rho = V.t.d[D_RHO][i]; s.p = V.t.d[D_P][i];
    s.v[0] = V.t.d[D_VX][i]; s.v[1] = V.t.d[D_VY][i]; s.v[2] = V.t.d[D_VZ][i];
  }
  // the whole in_support box: nreal counts every listed row inside it, weight or not
  static __device__ __forceinline__ double reach(const KParams &K, const I &s, double) { return s.supp; }
  static __device__ __forceinline__ void begin(const KParams &, I &, A &a, int)
  {
    a.nreal = 0;
    for (int k = 0; k < NA; k++) a.aa[k] = 0.0;
    for (int k = 0; k < DIM; k++) for (int q = 0; q < NQ; q++) a.f[k][q] = 0.0;
  }
  static __device__ __forceinline__ void pair(const DevView &V, const KParams &, const I &s, A &a, double dx, double dy,
                                              double dz, double, int j)
  {
    const double4 r0 = ld_rec(V.rec[0], j), vj = ld_rec(V.rec[1], j);
    const double r2 = dx * dx + dy * dy + dz * dz;
    // dW/dx_k = -2 (dx_k / h^2) W ; times m_j / rho_j
    const double gm = -2.0 * s.invh * s.invh * dev_weight_in<DIM>(r2, s.invh) * r0.x;
    const double d[3] = { dx, dy, dz };
    if (DIM == 3) {
      a.aa[0] -= dx * dx * gm; a.aa[1] -= dx * dy * gm; a.aa[2] -= dx * dz * gm;
      a.aa[3] -= dy * dy * gm; a.aa[4] -= dy * dz * gm; a.aa[5] -= dz * dz * gm;
    } else a.aa[0] -= dx * dx * gm;
    const double dq[5] = { s.rho - r0.y, s.v[0] - vj.x, s.v[1] - vj.y, s.v[2] - vj.z, s.p - r0.z };
#pragma unroll
    for (int k = 0; k < DIM; k++) {
      const double w = d[k] * gm;
      a.f[k][0] -= dq[0] * w;
#pragma unroll
      for (int b = 0; b < DIM; b++) a.f[k][b + 1] -= dq[1 + b] * w;
      a.f[k][NQ - 1] -= dq[4] * w;
    }
  }
  static __device__ __forceinline__ bool end(const DevView &V, const KParams &, StepState *, int i, I &, A &a, int, int lane)
  {
    const int nreal = warp_sum_i(a.nreal);
    double aa[NA], f[DIM][NQ];
    for (int k = 0; k < NA; k++) aa[k] = warp_sum(a.aa[k]);
    for (int k = 0; k < DIM; k++) for (int q = 0; q < NQ; q++) f[k][q] = warp_sum(a.f[k][q]);
    if (lane == 0) {
      double g[DIM][NQ];
      for (int k = 0; k < DIM; k++) for (int q = 0; q < NQ; q++) g[k][q] = 0.0;
      if (nreal > 20) {
        if (DIM == 1) {
          const double inv = -1 / aa[0]; // inv1 returns -1/A (sph_lib.cc:128-134): kept, it is the reference's result
          for (int q = 0; q < NQ; q++) g[0][q] = inv * f[0][q];
        } else { // inv3 (sph_lib.cc:86-110) of the symmetric moment matrix, then A^-1 F
          const double A0 = aa[0], A1 = aa[1], A2 = aa[2], A4 = aa[3], A5 = aa[4], A8 = aa[5];
          const double det = A0 * (A4 * A8 - A5 * A5) - A1 * (A1 * A8 - A5 * A2) + A2 * (A1 * A5 - A4 * A2);
          const double t = 1.0 / det;
          const double iv[3][3] = { { (A4 * A8 - A5 * A5) * t, (A2 * A5 - A1 * A8) * t, (A1 * A5 - A2 * A4) * t },
                                    { (A2 * A5 - A1 * A8) * t, (A0 * A8 - A2 * A2) * t, (A1 * A2 - A0 * A5) * t },
                                    { (A1 * A5 - A2 * A4) * t, (A1 * A2 - A0 * A5) * t, (A0 * A4 - A1 * A1) * t } };
          for (int k = 0; k < 3; k++)
            for (int q = 0; q < NQ; q++) g[k][q] = iv[k][0] * f[0][q] + iv[k][1] * f[1][q] + iv[k][2] * f[2][q];
        }
      }
      // column layout: d_rho[3], d_u[3], d_v[3], d_w[3], d_p[3]
      for (int k = 0; k < DIM; k++) {
        V.t.d[D_GRAD0 + k][i] = g[k][0];
        for (int b = 0; b < DIM; b++) V.t.d[D_GRAD0 + 3 * (b + 1) + k][i] = g[k][b + 1];
        V.t.d[D_GRAD0 + 12 + k][i] = g[k][NQ - 1];
      }
    }
    return false;
  }
};

// =========================================================================================
// Riemann problem between i ("right") and j ("left") along e = (x_i - x_j)/|x_i - x_j|
// =========================================================================================
// interface position, cubic in the specific volume (Compute_sij_star, sph_lib.cc:1667-1690)
__device__ __forceinline__ double dev_sij_star3(double hij, double Vi, double Vj, double Dri, double Drj, double dist)
{
  const double DVi = -Vi * Vi * Dri, DVj = -Vj * Vj * Drj;
  const double dist_sq = dist * dist, dist_cub = dist_sq * dist;
  const double A = -2.0 * (Vi - Vj) / dist_cub + (DVi + DVj) / dist_sq;
  const double B = 0.5 * (DVi - DVj) / dist;
  const double C = 1.5 * (Vi - Vj) / dist - 0.25 * (DVi + DVj);
  const double D = 0.5 * (Vi + Vj) - 0.125 * (DVi - DVj) * dist;
  const double hsq = hij * hij, h4 = hsq * hsq, h6 = h4 * hsq;
  const double Vijsq = 0.234375 * h6 * A * A + 0.1875 * h4 * (2.0 * A * C + B * B) + 0.25 * hsq * (2.0 * B * D + C * C) + D * D;
  return (0.46875 * h6 * A * B + 0.375 * h4 * (A * D + B * C) + 0.5 * hsq * C * D) / Vijsq;
}

// Gi / Gj: the 15 gradient columns of i / j (d_rho, d_u, d_v, d_w, d_p), read only when K.gsph_recon == 3: the Riemann
// states are then reconstructed from them (sph_lib.cc:2300-2486: GSPH_SPECIFIC_VOL_APP 3, MINI_DERIVATIVE_COND 10,
// GSPH_MODIFIED_MONOTONICITY 1, C_SHOCK 3 — the 3-D defaults, options.h:544-663)
template <int DIM>
__device__ __forceinline__ void dev_riemann(const KParams &K, double rhoi, double rhoj, const double *vi, const double *vj,
                                            double pi, double pj, const double *dx, double dist, double gi, double gj,
                                            double hij, double CSi, double CSj, double dt_half, const double *Gi, const double *Gj,
                                            double &p_star, double *v_star)
{
  double e[3] = { 0, 0, 0 }, ui = 0.0, uj = 0.0;
  for (int k = 0; k < DIM; k++) e[k] = dx[k] / dist;                 // Compute_eij, sph_lib.cc:1624-1631
  for (int k = 0; k < DIM; k++) { ui += vi[k] * e[k]; uj += vj[k] * e[k]; }
  double dr = rhoi, ur = ui, pr = pi, dl = rhoj, ul = uj, pl = pj;   // MINI_DERIVATIVE_COND 4
  if (K.gsph_recon == 3) {
    double drhoi = 0.0, dpi = 0.0, drhoj = 0.0, dpj = 0.0, dvi[3] = { 0, 0, 0 }, dvj[3] = { 0, 0, 0 };
    for (int a = 0; a < DIM; a++) {
      drhoi += Gi[a] * e[a];
      drhoj += Gj[a] * e[a];
      dpi += Gi[12 + a] * e[a];
      dpj += Gj[12 + a] * e[a];
      for (int k = 0; k < DIM; k++) { dvi[k] += Gi[3 * (k + 1) + a] * e[a]; dvj[k] += Gj[3 * (k + 1) + a] * e[a]; }
    }
    double ddui = 0.0, dduj = 0.0;
    for (int k = 0; k < DIM; k++) { ddui += dvi[k] * e[k]; dduj += dvj[k] * e[k]; }
    if (ddui * dduj < 0) { ddui = 0.0; dduj = 0.0; }                // monotonicity (:2357-2363)
    const double csmin = fmin(CSi, CSj);
    if (3.0 * fabs(uj - ui) > csmin) { drhoi = dpi = ddui = drhoj = dpj = dduj = 0.0; }   // (:2364-2376)
    const double Si = 0.5 * dist, Sj = -Si;
    const double sij = dev_sij_star3(hij, 1 / rhoi, 1 / rhoj, drhoi, drhoj, dist);
    const double delta_i = sij + CSi * dt_half - Si, delta_j = sij - CSj * dt_half - Sj;
    dr = rhoi + drhoi * delta_i; ur = ui + ddui * delta_i; pr = pi + dpi * delta_i;
    dl = rhoj + drhoj * delta_j; ul = uj + dduj * delta_j; pl = pj + dpj * delta_j;
    if (dr < 0) dr = rhoi;                                          // MINI_DERIVATIVE_COND 10 (:2467-2486)
    if (dl < 0) dl = rhoj;
    if (pr < 0) pr = pi;
    if (pl < 0) pl = pj;
    if (3.0 * fabs(ul - ur) > csmin) { ul = uj; ur = ui; }
  }
  const double rdl = sqrt(dl), rdr = sqrt(dr), den = 1.0 / (rdl + rdr);
  const double ulr = (ul * rdl + ur * rdr) * den;
  const double plr = (pl * rdl + pr * rdr) * den, gammalr = (gj * rdl + gi * rdr) * den, dlr = (dl * rdl + dr * rdr) * den;
  double clr = sqrt(gammalr * plr / dlr), u_star;
  if (K.riemann_solver == 0) {                                     // Roe, Lagrangian sound speed (sph_lib.cc:1694-1749)
    clr = dlr * clr;
    p_star = plr - 0.5 * clr * (ur - ul);
    u_star = ulr - 0.5 * (pr - pl) / clr;
  } else {                                                         // HLLC (sph_lib.cc:1753-1908)
    const double vl = ul - ulr, vr = ur - ulr;
    const double cl = sqrt(gj * pl / dl), cr = sqrt(gi * pr / dr);
    const double Sl = fmin(vl - cl, -clr), Sr = fmax(vr + cr, clr);
    const double Sm = (dr * vr * (Sr - vr) - dl * vl * (Sl - vl) + pl - pr) / (dr * (Sr - vr) - dl * (Sl - vl));
    const double p_hat = dl * (vl - Sl) * (vl - Sm) + pl;
    if (Sl > 0) { p_star = pl; u_star = ul; }
    else if (Sm > 0) {
      const double Ml = dl * ul, El = dl * (pl / (dl * (gj - 1)) + 0.5 * ul * ul);
      p_star = Sm / (Sl - Sm) * ((Sl - vl) * Ml + (p_hat - pl)) + p_hat;
      u_star = (Sm / (Sl - Sm) * ((Sl - vl) * El + p_hat * Sm - pl * vl) + (Sm + ulr) * p_hat) / p_star;
    } else if (Sr > 0) {
      const double Mr = dr * ur, Er = dr * (pr / (dr * (gi - 1)) + 0.5 * ur * ur);
      p_star = Sm / (Sr - Sm) * ((Sr - vr) * Mr + (p_hat - pr)) + p_hat;
      u_star = (Sm / (Sr - Sm) * ((Sr - vr) * Er + p_hat * Sm - pr * vr) + (Sm + ulr) * p_hat) / p_star;
    } else { p_star = pr; u_star = ur; }
  }
  for (int k = 0; k < DIM; k++) {                                   // SHEAR_VEL_APP 2
    const double vlr = (vj[k] * rdl + vi[k] * rdr) * den;
    v_star[k] = e[k] * u_star + vlr - ulr * e[k];
  }
}

// =========================================================================================
// mom_engr_update, GSPH: pass 0 accumulates rhs_v, pass 1 the energy with
// x*_dot = v_i + dt/2 (rhs_v + g)  (:983-986); the Riemann problem is solved in both passes, as the
// reference does (:772, :994).
// =========================================================================================
template <int DIM>
struct GsphMomStage
{
  static constexpr bool kSymH = true, kSkipSelf = true, kCountSupport = false;
  static constexpr int kRows = F_NEED, kWarps = DIM == 3 ? 8 : 16;   // 3-D: both gradient sets live in registers
  struct I { int i, pass; double h, supp, hk, rho, Vi, p, c, gamma, T, Cp, u[3], rhs[3], xdot[3], dt_half, G[15]; };
  struct A { double rv[3], re; };
  static __device__ __forceinline__ void init(const DevView &V, const KParams &K, StepState *ss, int i, I &s)
  {
    s.i = i;
    s.dt_half = 0.5 * ss->dt;
    if (K.gsph_recon == 3)
      for (int k = 0; k < 15; k++) s.G[k] = V.t.d[D_GRAD0 + k][i];
    s.h = V.t.d[D_H][i]; s.supp = K.cutoff2 * s.h;
    s.rho = V.t.d[D_RHO][i]; s.Vi = 1.0 / s.rho; s.p = V.t.d[D_P][i];
    s.c = V.t.d[D_C][i]; s.gamma = V.t.d[D_GAMMA][i]; s.T = V.t.d[D_T][i]; s.Cp = V.t.d[D_CP][i];
    const int u = K.lans ? D_SVX : D_VX;
    s.u[0] = V.t.d[u][i]; s.u[1] = V.t.d[u + 1][i]; s.u[2] = V.t.d[u + 2][i];
    s.rhs[0] = s.rhs[1] = s.rhs[2] = 0.0;
  }
  static __device__ __forceinline__ double reach(const KParams &K, const I &s, double hmax)
  { return fmin(K.cutoff * 0.5 * (s.h + hmax), s.supp); }
  static __device__ __forceinline__ void begin(const KParams &, I &s, A &a, int pass)
  { s.pass = pass; a.rv[0] = a.rv[1] = a.rv[2] = a.re = 0.0; }
  static __device__ __forceinline__ void pair(const DevView &V, const KParams &K, const I &s, A &a, double dx, double dy,
                                              double dz, double hj, int j)
  {
    const double r2 = dx * dx + dy * dy + dz * dz;
    if (!(r2 > 0.0)) return;                                        // dW = 0: max_dw > 0 fails (:735)
    const double h = 0.5 * (s.h + hj), invh = 1.0 / h;
    const double4 r0 = ld_rec(V.rec[0], j), r1 = ld_rec(V.rec[1], j), uj4 = ld_rec(V.rec[2], j);
    const double g = -2.0 * invh * invh * dev_weight_in<DIM>(r2, invh); // dW/dx_k = g dx_k
    const double Vj = 1.0 / r0.y;
    const double d[3] = { dx, dy, dz }, uj[3] = { uj4.x, uj4.y, uj4.z };
    double p_star, v_star[3];
    double Gj[15];
    if (K.gsph_recon == 3)
      for (int k = 0; k < 15; k++) Gj[k] = __ldg(&V.t.d[D_GRAD0 + k][j]);
    dev_riemann<DIM>(K, s.rho, r0.y, s.u, uj, s.p, r0.z, d, sqrt(r2), s.gamma, r1.x, h, s.c, r0.w, s.dt_half, s.G, Gj, p_star, v_star);
    double vv = 0.0;
    for (int k = 0; k < DIM; k++) vv += (s.u[k] - uj[k]) * (s.u[k] - uj[k]);
    const double turb = K.lans ? K.epsilon_half * vv * Vj : 0.0;    // SPH_epsilon_mom
    const double F = r0.x * (p_star * (s.Vi * s.Vi + Vj * Vj) - turb) * g; // m_j (p*(Vi^2+Vj^2) - turb) dW_k / dx_k
    if (s.pass == 0) {
      for (int k = 0; k < DIM; k++) a.rv[k] -= F * d[k];
    } else {
      double de = 0.0;
      for (int k = 0; k < DIM; k++) de -= (v_star[k] - s.xdot[k]) * F * d[k];
      double heat = 0.0;
      if (K.lans && DIM == 3) {
        const double vx = s.u[0] - uj[0], vy = s.u[1] - uj[1], vz = s.u[2] - uj[2];
        const double avr = fabs(vx * dx) + fabs(vy * dy) + fabs(vz * dz);
        if (avr != 0.) {
          const double cab = 0.5 * (s.c + r0.w);
          const double kij = 10.0 * K.epsilon_half * (0.5 * (r1.y + s.Cp)) * r2 * vv / (K.prandtl * avr) / (h * cab);
          heat = r0.x * Vj * s.Vi * kij * (s.T - r1.z) * g;
        }
      }
      a.re += de + heat;
    }
  }
  static __device__ __forceinline__ bool end(const DevView &V, const KParams &K, StepState *ss, int i, I &s, A &a, int pass, int lane)
  {
    const double dt = ss->dt, grav = -K.g;
    if (pass == 0) {
      for (int k = 0; k < DIM; k++) {
        s.rhs[k] = warp_sum(a.rv[k]);
        s.xdot[k] = s.u[k] + 0.5 * dt * (s.rhs[k] + (k == DIM - 1 ? grav : 0.0));
      }
      return true;
    }
    const double re = warp_sum(a.re);
    if (lane == 0) {
      double un[3] = { 0, 0, 0 };
      for (int k = 0; k < DIM; k++) un[k] = V.t.d[D_VX + k][i] + dt * s.rhs[k];
      un[DIM - 1] += dt * grav;
      V.scr[0][i] = un[0]; V.scr[1][i] = un[1]; V.scr[2][i] = un[2];
      V.scr[3][i] = V.t.d[D_E][i] + dt * (re + s.u[DIM - 1] * grav);
      if (!(s.rho > 0)) atomicAdd(&ss->bad, 1);
    }
    return false;
  }
};

// =========================================================================================
#define GRID(n) (unsigned) (((n) + 255) / 256), 256, 0, ctx->stream

int psph_gsph_gradients(psph_ctx *ctx)
{
  if (!ctx->has_grad) { ctx->err = "psph_calc_gradients: context was not created with use_gsph"; return -2; }
  if (!ctx->list_valid && psph_build_cell_list(ctx)) return -1;
  if (ctx->n == 0) { psph_comm_mark(ctx, DIRTY_GRAD); return 0; }   // an empty slab still enters the exchange that follows
  k_pack_grad<<<GRID(ctx->n)>>>(ctx->cur, ctx->d_ss);
  PSPH_LAUNCH_CHECK();
  if (launch_gather<GradStage>(ctx, KN_GATHER_GRAD)) return -1;
  psph_comm_mark(ctx, DIRTY_GRAD);
  return 0;
}

extern "C" int psph_calc_gradients(psph_ctx *ctx)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  StageTimer tm(ctx, ST_GRAD);
  return psph_gsph_gradients(ctx);
}

int psph_gsph_mom_engr(psph_ctx *ctx)
{
  StageTimer tm(ctx, ST_MOM);
  if (!ctx->list_valid && psph_build_cell_list(ctx)) return -1;
  if (ctx->n == 0) { psph_comm_mark(ctx, DIRTY_MOM); return 0; }
  k_pack_gsph<<<GRID(ctx->n)>>>(ctx->cur, ctx->kp, ctx->d_ss);
  PSPH_LAUNCH_CHECK();
  if (launch_gather<GsphMomStage>(ctx, KN_GATHER_GSPH)) return -1;
  return psph_commit_mom(ctx);
}
