/* psph_oracle.c — CPU restatement of the Plume-SPH per-timestep hot path.
 * TEST INFRASTRUCTURE ONLY — see psph_oracle.h for the rules and the parity pins.
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (no FMA contraction, so that sums match
 * the -O2 x86-64 build of the reference bit for bit). */
#include "psph_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define D3(a, i, k) ((a)[3 * (i) + (k)])

static int need_neigh(const psph_oracle_state *st, int64_t i)
{ /* particle.h:555-558 */
  return (st->bc[i] == 100 || st->bc[i] == 99) && !st->guest[i];
}
static int is_real(const psph_oracle_state *st, int64_t i)
{ /* particle.h:612-615 */
  return st->bc[i] == 100 || st->bc[i] == 99;
}
static int contr_dens(const psph_config *c, const psph_oracle_state *st, int64_t j)
{ /* particle.h:570-580 */
  int b = st->bc[j];
  if (c->pghost_contribute_des == 1) return b == 100 || b == 99 || b == 1;
  if (c->pghost_contribute_des == 2) return b == 100 || b == 99 || b == 1 || b == 0;
  return b == 100 || b == 99;
}
static int contr_vel_smooth(const psph_oracle_state *st, int64_t j)
{ /* particle.h:587-590 */
  int b = st->bc[j];
  return b == 100 || b == 99 || b == 2;
}

static int in_support(int dim, const double *dx, double h)
{ /* sph_header.h:49-60 */
  for (int i = 0; i < dim; i++)
    if (fabs(dx[i]) > h) return 0;
  return 1;
}

double psph_oracle_weight(int dim, const double *s, double h, double cutoff)
{ /* sph_lib.cc:207-230 */
  const double t1 = 0.564189583547756;
  for (int i = 0; i < dim; i++)
    if (fabs(s[i]) > cutoff) return 0;
  double norm = 1, expt = 0;
  for (int i = 0; i < dim; i++) {
    norm *= t1 / h;
    expt += s[i] * s[i];
  }
  return norm * exp(-expt);
}

double psph_oracle_d_weight(int dim, const double *s, double h, double cutoff, int dir)
{ /* sph_lib.cc:240-253 */
  for (int i = 0; i < dim; i++)
    if (fabs(s[i]) > cutoff) return 0;
  double tmp = -2 * s[dir] / h;
  return tmp * psph_oracle_weight(dim, s, h, cutoff);
}

void psph_oracle_eos(const psph_config *c, psph_oracle_state *st, int64_t i)
{ /* Particle::update_second_var, particle.cc:528-604 */
  double desm = st->rho[i];
  double ks = st->ks[i];
  double na = 1 - ks;
  double engr = st->e[i];
  double ng = c->ng0 * ks;
  double ns = ks - ng;
  double Cvm = ns * c->Cvs + ng * c->Cvg + na * c->Cva;
  if (c->fluid_compressibility == 0) {
    double Rm = ng * c->Rg + na * c->Ra;
    double Rmg = Rm / (ng + na);
    st->gamma[i] = 1 + Rm / Cvm;
    double desmg = desm * (na + ng);
    st->p[i] = Rmg * desmg * engr / Cvm;
    st->c[i] = pow(st->gamma[i] * st->p[i] / desmg, 0.5);
  } else {
    double gama = 7.0;
    double B = c->rhoa0 * st->c[i] * st->c[i] / gama; /* previous sound speed */
    st->p[i] = B * (pow(desm / c->rhoa0, gama) - 1) + c->pa0;
    st->c[i] = pow(gama * (st->p[i] + B - c->pa0) / desm, 0.5);
  }
  st->rho_ph[2 * i + 0] = st->rho[i] * na;
  st->rho_ph[2 * i + 1] = st->rho[i] * ks;
  st->T[i] = engr / Cvm;
  if (c->have_turbulence_lans != 0) st->Cp[i] = Cvm;
}

/* ---------------- background mesh ---------------- */
static int64_t ncells_of(const psph_config *c) { return (int64_t) c->ncell[0] * c->ncell[1] * c->ncell[2]; }
static int64_t cell_lin(const psph_config *c, int i, int j, int k)
{ return ((int64_t) i * c->ncell[1] + j) * c->ncell[2] + k; } /* x-major, z-minor as preprocess.cc:169-181 */
static double cell_min(const psph_config *c, int d, int i) { return (c->mindom[d] + i * c->del) / c->len_scale; }
static double cell_max(const psph_config *c, int d, int i) { return (c->mindom[d] + (i + 1) * c->del) / c->len_scale; }

static int cell_index_1(const psph_config *c, int d, double x)
{ /* membership is [mincrd, maxcrd) against the stored bounds, update_pos.cc:257-271 */
  int i = (int) floor((x * c->len_scale - c->mindom[d]) / c->del);
  if (i < 0) i = 0;
  if (i > c->ncell[d] - 1) i = c->ncell[d] - 1;
  while (i > 0 && x < cell_min(c, d, i)) i--;
  while (i < c->ncell[d] - 1 && x >= cell_max(c, d, i)) i++;
  return i;
}

int psph_oracle_bin(const psph_config *c, psph_oracle_state *st)
{
  int64_t nc = ncells_of(c);
  int32_t *cell = (int32_t *) malloc(sizeof(int32_t) * (st->n ? st->n : 1));
  memset(st->cell_off, 0, sizeof(int64_t) * (nc + 1));
  for (int64_t p = 0; p < st->n; p++) {
    int idx[3] = { 0, 0, 0 };
    if (c->dim == 3)
      for (int d = 0; d < 3; d++) idx[d] = cell_index_1(c, d, D3(st->x, p, d));
    cell[p] = (int32_t) cell_lin(c, idx[0], idx[1], idx[2]);
    st->cell_off[cell[p] + 1]++;
  }
  for (int64_t q = 0; q < nc; q++) st->cell_off[q + 1] += st->cell_off[q];
  int64_t *fill = (int64_t *) malloc(sizeof(int64_t) * nc);
  memcpy(fill, st->cell_off, sizeof(int64_t) * nc);
  for (int64_t p = 0; p < st->n; p++) st->cell_rows[fill[cell[p]]++] = (int32_t) p;
  free(fill);
  free(cell);
  return 0;
}

/* ---------------- neighbour search ---------------- */
void psph_oracle_free_lists(psph_oracle_state *st)
{
  free(st->nb_off); st->nb_off = NULL;
  free(st->nb_idx); st->nb_idx = NULL;
}

int psph_oracle_search(const psph_config *c, psph_oracle_state *st)
{ /* search_neighs_consth, neighs.cc:37-211 */
  int64_t nc = ncells_of(c);
  psph_oracle_free_lists(st);
  /* lists are stored per row; rows that get no list keep an empty one */
  int64_t *cnt = (int64_t *) calloc(st->n + 1, sizeof(int64_t));
  for (int pass = 0; pass < 2; pass++) {
    if (pass == 1) {
      st->nb_off = (int64_t *) malloc(sizeof(int64_t) * (st->n + 1));
      st->nb_off[0] = 0;
      for (int64_t p = 0; p < st->n; p++) st->nb_off[p + 1] = st->nb_off[p] + cnt[p];
      st->nb_idx = (int32_t *) malloc(sizeof(int32_t) * (st->nb_off[st->n] ? st->nb_off[st->n] : 1));
    }
    for (int ci = 0; ci < c->ncell[0]; ci++)
      for (int cj = 0; cj < c->ncell[1]; cj++)
        for (int ck = 0; ck < c->ncell[2]; ck++) {
          int64_t cl = cell_lin(c, ci, cj, ck);
          int64_t b0 = st->cell_off[cl], b1 = st->cell_off[cl + 1];
          if (b1 == b0) continue; /* :65 plist empty */
          int has_ghost = 0; /* :69 */
          for (int64_t q = b0; q < b1; q++) {
            int b = st->bc[st->cell_rows[q]];
            if (b == 0 || b == 1 || b == 2) has_ghost = 1;
          }
          for (int64_t q = b0; q < b1; q++) {
            int64_t i = st->cell_rows[q];
            if (st->guest[i]) continue; /* guest buckets are skipped (:65); single-rank: none */
            if (c->dim == 1) { /* no search in 1-D: every need_neigh particle lists all particles in
                                * creation order, add_air.cc:470-479 */
              if (!need_neigh(st, i)) continue;
            } else if (!c->use_gsph && has_ghost && st->bc[i] != 100) continue; /* :76-79 */
            double hi = c->neigh_radius * st->h[i]; /* :84 */
            int64_t w = pass ? st->nb_off[i] : 0, m = 0;
            /* all particles of the own bucket, no distance test (:88-92) */
            for (int64_t r = b0; r < b1; r++) {
              if (pass) st->nb_idx[w + m] = st->cell_rows[r];
              m++;
            }
            /* 26 neighbouring buckets in table order x-major/z-minor (preprocess.cc:288-314) */
            for (int di = -1; di <= 1; di++)
              for (int dj = -1; dj <= 1; dj++)
                for (int dk = -1; dk <= 1; dk++) {
                  if (!di && !dj && !dk) continue;
                  int ni = ci + di, nj = cj + dj, nk = ck + dk;
                  if (ni < 0 || nj < 0 || nk < 0 || ni >= c->ncell[0] || nj >= c->ncell[1] || nk >= c->ncell[2])
                    continue;
                  int64_t nl = cell_lin(c, ni, nj, nk);
                  for (int64_t r = st->cell_off[nl]; r < st->cell_off[nl + 1]; r++) {
                    int64_t j = st->cell_rows[r];
                    double ds[3];
                    for (int k = 0; k < c->dim; k++) ds[k] = D3(st->x, i, k) - D3(st->x, j, k);
                    if (in_support(c->dim, ds, hi)) { /* :130 */
                      if (pass) st->nb_idx[w + m] = (int32_t) j;
                      m++;
                    }
                  }
                }
            if (!pass) cnt[i] = m;
          }
        }
  }
  free(cnt);
  return 0;
}

/* ---------------- timestep ---------------- */
double psph_oracle_timestep(const psph_config *c, const psph_oracle_state *st)
{ /* timestep.cc:95-135 (3-D), :140-183 (1-D) */
  double dt = 1.0E+10;
  for (int64_t i = 0; i < st->n; i++)
    if (need_neigh(st, i)) {
      double temp = st->h[i] / st->c[i];
      if (temp < dt) dt = temp;
    }
  return c->cfl * dt;
}

/* ---------------- adaptive smoothing length ---------------- */
int psph_oracle_adaptive_sml(const psph_config *c, psph_oracle_state *st)
{ /* neighs.cc:479-693, ADAPTIVE_SML==1 branch */
  for (int64_t i = 0; i < st->n; i++) {
    if (!need_neigh(st, i)) continue;
    for (int loop = 0; loop < c->num_loop; loop++) {
      double H_old = st->h[i];
      double H_star = H_old * c->c_smooth;
      double supp = c->cutoff2 * H_star;
      double rho_star = 0.;
      for (int64_t q = st->nb_off[i]; q < st->nb_off[i + 1]; q++) {
        int64_t j = st->nb_idx[q];
        if (!contr_dens(c, st, j)) continue;
        double dx[3], si[3];
        for (int k = 0; k < c->dim; k++) {
          dx[k] = D3(st->x, i, k) - D3(st->x, j, k);
          si[k] = dx[k] / H_star;
        }
        if (in_support(c->dim, dx, supp))
          rho_star += psph_oracle_weight(c->dim, si, H_star, c->cutoff) * st->mass[j];
      }
      double H_new;
      if (c->dim == 3) H_new = c->eta_smooth * cbrt(st->mass[i] / rho_star);
      else H_new = c->eta_smooth * (st->mass[i] / rho_star);
      st->h[i] = H_new; /* put_smlen: RATIO_SML_DX == 1 (constant.h:190) */
      double err = fabs((H_new - H_old) / H_old);
      if (err < c->thresh) break;
    }
  }
  return 0;
}

/* ---------------- smooth_density ---------------- */
int psph_oracle_smooth_density(const psph_config *c, psph_oracle_state *st)
{ /* density_massfrac.cc:25-359 with DENSITY_UPDATE_SPH==11, DENSITY_UPDATE_SML==11 */
  int bad = 0;
  double *nrho = (double *) malloc(sizeof(double) * st->n);
  double *nks = (double *) malloc(sizeof(double) * st->n);
  double *ne = (double *) malloc(sizeof(double) * st->n);
  for (int64_t i = 0; i < st->n; i++) {
    if (!need_neigh(st, i)) continue;
    double supp = c->cutoff2 * st->h[i]; /* :93 */
    double tmprho[2] = { 0., 0. };
    double wnorm_e = 0.0, engr = 0.0;
    for (int64_t q = st->nb_off[i]; q < st->nb_off[i + 1]; q++) {
      int64_t j = st->nb_idx[q];
      if (!contr_dens(c, st, j)) continue; /* :135 */
      double mj = st->mass[j], rhoj = st->rho[j];
      int ph = st->phase[j];
      if (ph < 1 || ph > 2) continue;
      double ds[3], s[3], s_e[3];
      for (int k = 0; k < c->dim; k++) ds[k] = D3(st->x, i, k) - D3(st->x, j, k);
      if (!in_support(c->dim, ds, supp)) continue; /* :151 */
      double hi = 0.5 * (st->h[i] + st->h[j]); /* :155 */
      for (int k = 0; k < c->dim; k++) s[k] = ds[k] / hi;
      if (c->have_energy_smooth == 1) { /* :163-168 */
        for (int k = 0; k < c->dim; k++) s_e[k] = c->e_smooth_ratio * ds[k] / hi;
        double wm_e = psph_oracle_weight(c->dim, s_e, hi / c->e_smooth_ratio, c->cutoff) * mj;
        wnorm_e += wm_e / rhoj;
        engr += wm_e * st->e[j] / rhoj;
      }
      double wght = psph_oracle_weight(c->dim, s, hi, c->cutoff);
      double wm = wght * mj;
      tmprho[ph - 1] += wm;
    }
    if (c->have_energy_smooth == 1) {
      if (!(wnorm_e > 0)) bad++;
      engr = engr / wnorm_e; /* :230 */
      ne[i] = engr;
    }
    double density = 0.0, phaserho[2];
    for (int ph = 0; ph < 2; ph++) { /* :268-273 */
      phaserho[ph] = tmprho[ph];
      density += phaserho[ph];
    }
    double mssfrc = phaserho[1] / density; /* :285 */
    if (!(density > 0)) bad++;
    nrho[i] = density;
    nks[i] = mssfrc;
  }
  /* second pass, :337-352 */
  for (int64_t i = 0; i < st->n; i++)
    if (need_neigh(st, i)) {
      st->ks[i] = nks[i];
      if (st->ks[i] >= c->msfrc_thresh && st->involved[i] != 2) st->involved[i] = 2;
      st->rho[i] = nrho[i];
      if (c->have_energy_smooth == 1) st->e[i] = ne[i];
      psph_oracle_eos(c, st, i);
    }
  free(nrho); free(nks); free(ne);
  return bad;
}

/* ---------------- mom_engr_update (SPH) ---------------- */
static double art_vis(const psph_config *c, double rhoab, double sndspdab, const double *rab, const double *vab,
                      double rsqab, double h)
{ /* sph_lib.cc:402-424 */
  double vrab = 0., vis = 0.;
  for (int k = 0; k < c->dim; k++) vrab += rab[k] * vab[k];
  double miuab = h * vrab / (rsqab + c->ata * h * h);
  if (vrab > 0) vis = 0.;
  else if (vrab < 0) vis = (c->beta * miuab - c->alf * sndspdab) * miuab / rhoab;
  return vis;
}

static double sph_epsilon_heat_conductivity(const psph_config *c, double Cp_ab, const double *ds, const double *vab,
                                            double hab, double cab)
{ /* sph_lib.cc:1399-1439 */
  double dotvv = 0., dotvr = 0., dotrr = 0., kab;
  for (int i = 0; i < c->dim; i++) dotvr += fabs(vab[i] * ds[i]);
  if (dotvr == 0.) return 0.;
  for (int i = 0; i < c->dim; i++) dotrr += ds[i] * ds[i];
  for (int i = 0; i < c->dim; i++) dotvv += vab[i] * vab[i];
  if (c->dim == 3) kab = 10.0 * c->epsilon_half * Cp_ab * dotrr * dotvv / (c->prandtl_num * dotvr);
  else kab = 0.0;
  return kab / (hab * cab);
}

static double compute_F(int dim, const double *dwdx, const double *dx)
{ /* sph_lib.cc:257-281 */
  double r_sq = 0., W_r = 0.;
  for (int i = 0; i < dim; i++) W_r += dwdx[i] * dx[i];
  if (W_r == 0) return 0.0;
  for (int i = 0; i < dim; i++) r_sq += dx[i] * dx[i];
  return W_r / r_sq;
}

static int mom_engr_sph(const psph_config *c, psph_oracle_state *st, double dt)
{ /* mom_engr_update.cc:30-439, ME_UPDATE_SML==2, MOMENTUM_DISCRETIZE==0 */
  int bad = 0;
  int dim = c->dim, lans = c->have_turbulence_lans != 0;
  double *nv = (double *) malloc(sizeof(double) * 3 * st->n);
  double *ne = (double *) malloc(sizeof(double) * st->n);
  double gravity_last = -c->g;
  for (int64_t i = 0; i < st->n; i++) {
    if (!need_neigh(st, i)) continue;
    double hi = st->h[i], supp = c->cutoff2 * hi;
    double pressi = st->p[i], tempi = st->T[i], Cp_i = st->Cp[i];
    if (!(st->rho[i] > 0)) bad++;
    double Vi = 1.0 / st->rho[i];
    double pvsqi = pressi * Vi * Vi;
    double veli[3], rhs_v[3] = { 0, 0, 0 }, rhs_e = 0;
    for (int k = 0; k < dim; k++) veli[k] = lans ? D3(st->sv, i, k) : D3(st->v, i, k);
    for (int64_t q = st->nb_off[i]; q < st->nb_off[i + 1]; q++) {
      int64_t j = st->nb_idx[q];
      if (j == i) continue; /* :201 */
      double dist_sq = 0, dx[3], si[3], s_heati[3];
      double h = 0.5 * (hi + st->h[j]);
      for (int k = 0; k < dim; k++) {
        dx[k] = D3(st->x, i, k) - D3(st->x, j, k);
        si[k] = dx[k] / h;
        dist_sq += dx[k] * dx[k];
      }
      if (!in_support(dim, dx, supp)) continue;
      double mj = st->mass[j], pressj = st->p[j];
      if (!(st->rho[j] > 0)) bad++;
      double Vj = 1.0 / st->rho[j];
      double velij[3];
      for (int k = 0; k < dim; k++) velij[k] = veli[k] - (lans ? D3(st->sv, j, k) : D3(st->v, j, k));
      double rhoab = 0.5 * (st->rho[i] + st->rho[j]);
      double sndspdab = 0.5 * (st->c[i] + st->c[j]);
      double vis = art_vis(c, rhoab, sndspdab, dx, velij, dist_sq, h);
      if (lans) { /* SPH_epsilon_mom, sph_lib.cc:1331-1338 */
        double dotv = 0.;
        for (int k = 0; k < dim; k++) dotv += velij[k] * velij[k];
        vis -= c->epsilon_half * dotv * Vj;
      }
      double mpvsqij = mj * (pressj * Vj * Vj + pvsqi + vis);
      double dwdxi[3], dwdx_heati[3];
      for (int k = 0; k < dim; k++) s_heati[k] = si[k] * c->heat_trans_scale_ratio;
      for (int k = 0; k < dim; k++) {
        dwdxi[k] = psph_oracle_d_weight(dim, si, h, c->cutoff, k);
        dwdx_heati[k] = psph_oracle_d_weight(dim, s_heati, h / c->heat_trans_scale_ratio, c->cutoff, k);
      }
      double heat_tran = 0.;
      if (lans) { /* :307-314 */
        double Cp_ij = 0.5 * (st->Cp[j] + Cp_i);
        double hab = 0.5 * (hi + st->h[j]);
        double kij = sph_epsilon_heat_conductivity(c, Cp_ij, dx, velij, hab, sndspdab);
        double Fij = compute_F(dim, dwdx_heati, dx);
        heat_tran = mj * Vj * Vi * kij * (tempi - st->T[j]) * Fij;
      }
      for (int k = 0; k < dim; k++) rhs_v[k] -= mpvsqij * dwdxi[k];
      double deltae = 0.;
      for (int k = 0; k < dim; k++) deltae += 0.5 * (mpvsqij * dwdxi[k]) * velij[k];
      rhs_e += (deltae + heat_tran);
    }
    double unew[3];
    for (int k = 0; k < dim; k++) unew[k] = D3(st->v, i, k) + dt * rhs_v[k];
    unew[dim - 1] += dt * gravity_last; /* :350 */
    for (int k = 0; k < dim; k++) D3(nv, i, k) = unew[k];
    ne[i] = st->e[i] + dt * (rhs_e + veli[dim - 1] * gravity_last); /* :354 */
  }
  for (int64_t i = 0; i < st->n; i++)
    if (need_neigh(st, i)) { /* :412-418 */
      for (int k = 0; k < dim; k++) D3(st->v, i, k) = D3(nv, i, k);
      st->e[i] = ne[i];
      psph_oracle_eos(c, st, i);
    }
  free(nv); free(ne);
  return bad;
}

/* ---------------- GSPH: Riemann solvers, mom_engr_update, calc_gradients ---------------- */
/* Option sets restated: RP_MASS_WEIGHTED 0, SWITCH_OFF_AV_FOR_EXPAN 0, SHEAR_VEL_APP 2, HLL_WAVE_SPEED_EVA 0,
 * FLUID_COMPRESSIBILITY 0, and for the Riemann states (psph_config.gsph_reconstruct)
 *   0: GSPH_SPECIFIC_VOL_APP 0, MINI_DERIVATIVE_COND 4 — the 1-D defaults (options.h:141-262): piecewise constant,
 *      the gradients do not enter the states (sph_lib.cc:2487-2493);
 *   3: GSPH_SPECIFIC_VOL_APP 3, MINI_DERIVATIVE_COND 10, GSPH_MODIFIED_MONOTONICITY 1 — the 3-D defaults
 *      (options.h:544-663), normally with the Roe solver (RIEMANN_SOLVER 0). */
static void hllc_rp_solver(double dl, double dr, double pl, double pr, double ul, double ur, double gj, double gi,
                           double *p_star, double *u_star)
{ /* sph_lib.cc:1753-1908, USE_GSPH==1 branch */
  double rdl = sqrt(dl), rdr = sqrt(dr);
  double denominator = 1.0 / (rdl + rdr);
  double ulr = (ul * rdl + ur * rdr) * denominator;
  double plr = (pl * rdl + pr * rdr) * denominator;
  double gammalr = (gj * rdl + gi * rdr) * denominator;
  double dlr = (dl * rdl + dr * rdr) * denominator;
  double clr = sqrt(gammalr * plr / dlr);
  double vl = ul - ulr, vr = ur - ulr;
  double cl = sqrt(gj * pl / dl), cr = sqrt(gi * pr / dr);
  double Sl = fmin(vl - cl, -clr);
  double Sr = fmax(vr + cr, clr);
  double Sm = (dr * vr * (Sr - vr) - dl * vl * (Sl - vl) + pl - pr) / (dr * (Sr - vr) - dl * (Sl - vl));
  double p_hat = dl * (vl - Sl) * (vl - Sm) + pl;
  double Ml = dl * ul, Mr = dr * ur;
  double el = pl / (dl * (gj - 1)), er = pr / (dr * (gi - 1));
  double El = dl * (el + 0.5 * ul * ul), Er = dr * (er + 0.5 * ur * ur);
  if (Sl > 0) { *p_star = pl; *u_star = ul; }
  else if ((Sl <= 0) && (Sm > 0)) {
    *p_star = Sm / (Sl - Sm) * ((Sl - vl) * Ml + (p_hat - pl)) + p_hat;
    *u_star = (Sm / (Sl - Sm) * ((Sl - vl) * El + p_hat * Sm - pl * vl) + (Sm + ulr) * p_hat) / (*p_star);
  } else if ((Sm <= 0) && (Sr > 0)) {
    *p_star = Sm / (Sr - Sm) * ((Sr - vr) * Mr + (p_hat - pr)) + p_hat;
    *u_star = (Sm / (Sr - Sm) * ((Sr - vr) * Er + p_hat * Sm - pr * vr) + (Sm + ulr) * p_hat) / (*p_star);
  } else if (Sr <= 0) { *p_star = pr; *u_star = ur; }
  else { *p_star = NAN; *u_star = NAN; }
}

static void roe_rp_solver(double dl, double dr, double pl, double pr, double ul, double ur, double gj, double gi,
                          double *p_star, double *u_star)
{ /* sph_lib.cc:1694-1749 */
  double rdl = sqrt(dl), rdr = sqrt(dr);
  double denominator = 1.0 / (rdl + rdr);
  double plr = (pl * rdl + pr * rdr) * denominator;
  double gammalr = (gj * rdl + gi * rdr) * denominator;
  double dlr = (dl * rdl + dr * rdr) * denominator;
  double clr = sqrt(gammalr * plr / dlr);
  clr = dlr * clr;
  double ulr = (ul * rdl + ur * rdr) * denominator;
  *p_star = plr - 0.5 * clr * (ur - ul);
  *u_star = ulr - 0.5 * (pr - pl) / clr;
}

/* interface position, cubic in the specific volume (Compute_sij_star, sph_lib.cc:1667-1690) */
static double compute_sij_star3(double hij, double Vi, double Vj, double Dri, double Drj, double dist)
{
  double DVi = -Vi * Vi * Dri;
  double DVj = -Vj * Vj * Drj;
  double dist_sq = dist * dist;
  double dist_cub = dist_sq * dist;
  double A = -2.0 * (Vi - Vj) / dist_cub + (DVi + DVj) / dist_sq;
  double B = 0.5 * (DVi - DVj) / dist;
  double C = 1.5 * (Vi - Vj) / dist - 0.25 * (DVi + DVj);
  double D = 0.5 * (Vi + Vj) - 0.125 * (DVi - DVj) * dist;
  double hsq = hij * hij;
  double h4 = hsq * hsq;
  double h6 = h4 * hsq;
  double Vijsq = 0.234375 * h6 * A * A + 0.1875 * h4 * (2.0 * A * C + B * B) + 0.25 * hsq * (2.0 * B * D + C * C) + D * D;
  return (0.46875 * h6 * A * B + 0.375 * h4 * (A * D + B * C) + 0.5 * hsq * C * D) / Vijsq;
}

/* gi/gj: the 15 gradient columns of i / j (d_rho, d_u, d_v, d_w, d_p; calc_gradients), used when
 * c->gsph_reconstruct == 3: GSPH_SPECIFIC_VOL_APP 3, MINI_DERIVATIVE_COND 10, GSPH_MODIFIED_MONOTONICITY 1,
 * C_SHOCK = 3 (constant.h:203) — the 3-D defaults (options.h:544-663) */
static void riemann_solver(const psph_config *c, double rhoi, double rhoj, const double *vi, const double *vj, double pi,
                           double pj, double hi, double hj, const double *xi, const double *xj, double CSi, double CSj,
                           double gi, double gj, double dt_half, const double *Gi, const double *Gj, double *p_star,
                           double *v_star)
{ /* sph_lib.cc:2247-2600 */
  int dim = c->dim;
  double hij = 0.5 * (hi + hj);
  double dx[3], e[3], dist = 0.0;
  for (int i = 0; i < dim; i++) dx[i] = xi[i] - xj[i];
  for (int i = 0; i < dim; i++) dist += dx[i] * dx[i];
  dist = sqrt(dist);
  for (int i = 0; i < dim; i++) e[i] = dx[i] / dist; /* Compute_eij, :1624-1631 */
  double ui = 0.0, uj = 0.0;
  for (int i = 0; i < dim; i++) { ui += vi[i] * e[i]; uj += vj[i] * e[i]; }
  double dr = rhoi, ur = ui, pr = pi, dl = rhoj, ul = uj, pl = pj; /* MINI_DERIVATIVE_COND == 4 */
  if (c->gsph_reconstruct == 3) {
    /* gradients projected on e (:2300-2355); the velocity gradient twice: (grad u_k . e) e_k */
    double drhoi = 0.0, dpi = 0.0, drhoj = 0.0, dpj = 0.0, dvi[3] = { 0, 0, 0 }, dvj[3] = { 0, 0, 0 };
    for (int i = 0; i < dim; i++) {
      drhoi += Gi[i] * e[i];
      drhoj += Gj[i] * e[i];
      dpi += Gi[12 + i] * e[i];
      dpj += Gj[12 + i] * e[i];
      for (int k = 0; k < dim; k++) {
        dvi[k] += Gi[3 * (k + 1) + i] * e[i];
        dvj[k] += Gj[3 * (k + 1) + i] * e[i];
      }
    }
    double ddui, dduj;
    if (dim == 3) { ddui = dvi[0] * e[0] + dvi[1] * e[1] + dvi[2] * e[2]; dduj = dvj[0] * e[0] + dvj[1] * e[1] + dvj[2] * e[2]; }
    else { ddui = dvi[0] * e[0]; dduj = dvj[0] * e[0]; }
    if (ddui * dduj < 0) { ddui = 0.0; dduj = 0.0; }                 /* monotonicity (:2357-2363) */
    if (3.0 * fabs(uj - ui) > fmin(CSi, CSj)) {                      /* GSPH_MODIFIED_MONOTONICITY 1 (:2364-2376) */
      drhoi = 0.0; dpi = 0.0; ddui = 0.0;
      drhoj = 0.0; dpj = 0.0; dduj = 0.0;
    }
    double Si = 0.5 * dist, Sj = -Si;
    double sij_star = compute_sij_star3(hij, 1 / rhoi, 1 / rhoj, drhoi, drhoj, dist);
    double delta_i = sij_star + CSi * dt_half - Si;
    double delta_j = sij_star - CSj * dt_half - Sj;
    dr = rhoi + drhoi * delta_i;
    ur = ui + ddui * delta_i;
    pr = pi + dpi * delta_i;
    dl = rhoj + drhoj * delta_j;
    ul = uj + dduj * delta_j;
    pl = pj + dpj * delta_j;
    if (dr < 0) dr = rhoi;                                           /* MINI_DERIVATIVE_COND 10 (:2467-2486) */
    if (dl < 0) dl = rhoj;
    if (pr < 0) pr = pi;
    if (pl < 0) pl = pj;
    if (3.0 * fabs(ul - ur) > fmin(CSi, CSj)) { ul = uj; ur = ui; }
  }
  double u_star;
  if (c->riemann_solver == 0) roe_rp_solver(dl, dr, pl, pr, ul, ur, gj, gi, p_star, &u_star);
  else hllc_rp_solver(dl, dr, pl, pr, ul, ur, gj, gi, p_star, &u_star);
  /* SHEAR_VEL_APP == 2 */
  double rdl = sqrt(dl), rdr = sqrt(dr);
  double denominator = 1.0 / (rdl + rdr);
  double ulr = (ul * rdl + ur * rdr) * denominator;
  for (int i = 0; i < dim; i++) {
    double vlr = (vj[i] * rdl + vi[i] * rdr) * denominator;
    v_star[i] = e[i] * u_star + vlr - ulr * e[i];
  }
}

static int mom_engr_gsph(const psph_config *c, psph_oracle_state *st, double dt)
{ /* mom_engr_update.cc:443-1145, ME_UPDATE_SML==2 */
  int bad = 0;
  int dim = c->dim, lans = c->have_turbulence_lans != 0;
  double *nv = (double *) malloc(sizeof(double) * 3 * st->n);
  double *ne = (double *) malloc(sizeof(double) * st->n);
  double gravity[3] = { 0, 0, 0 };
  gravity[dim - 1] = -c->g;
  for (int64_t i = 0; i < st->n; i++) {
    if (!need_neigh(st, i)) continue;
    double xi[3], hi = st->h[i], supp = c->cutoff2 * hi;
    for (int k = 0; k < dim; k++) xi[k] = D3(st->x, i, k);
    double pressi = st->p[i], tempi = st->T[i], Cp_i = st->Cp[i], sndspdi = st->c[i], gammai = st->gamma[i];
    if (!(st->rho[i] > 0)) bad++;
    double Vi = 1.0 / st->rho[i];
    double veli[3], rhs_v[3] = { 0, 0, 0 }, rhs_e = 0;
    for (int k = 0; k < dim; k++) veli[k] = lans ? D3(st->sv, i, k) : D3(st->v, i, k);
    for (int pass = 0; pass < 2; pass++) {
      for (int64_t q = st->nb_off[i]; q < st->nb_off[i + 1]; q++) {
        int64_t j = st->nb_idx[q];
        if (j == i) continue;
        double dx[3], xj[3], si[3], dwdxi[3], vsqdwi[3], vsqdwj[3], velj[3], velij[3];
        for (int k = 0; k < dim; k++) { xj[k] = D3(st->x, j, k); dx[k] = xi[k] - xj[k]; }
        if (!in_support(dim, dx, supp)) continue;
        if (!(st->rho[j] > 0)) bad++;
        double Vj = 1.0 / st->rho[j];
        double mj = st->mass[j], pressj = st->p[j], hj = st->h[j], sndspdj = st->c[j], gammaj = st->gamma[j];
        double hij = 0.5 * (hi + hj);
        for (int k = 0; k < dim; k++) si[k] = dx[k] / hij;
        for (int k = 0; k < dim; k++) dwdxi[k] = psph_oracle_d_weight(dim, si, hij, c->cutoff, k);
        double max_dw = 0.0;
        for (int k = 0; k < dim; k++) {
          vsqdwi[k] = Vi * Vi * dwdxi[k];
          vsqdwj[k] = Vj * Vj * dwdxi[k];
          if (fabs(vsqdwi[k]) > max_dw) max_dw = fabs(vsqdwi[k]);
        }
        if (!(max_dw > 0)) continue;
        for (int k = 0; k < dim; k++) velj[k] = lans ? D3(st->sv, j, k) : D3(st->v, j, k);
        double p_star, v_star[3];
        riemann_solver(c, st->rho[i], st->rho[j], veli, velj, pressi, pressj, hi, hj, xi, xj, sndspdi, sndspdj, gammai, gammaj,
                       0.5 * dt, st->grad + 15 * i, st->grad + 15 * j, &p_star, v_star);
        double turb_stress = 0.0;
        for (int k = 0; k < dim; k++) velij[k] = veli[k] - velj[k];
        if (lans) { /* SPH_epsilon_mom, sph_lib.cc:1331-1338 */
          double dotv = 0.;
          for (int k = 0; k < dim; k++) dotv += velij[k] * velij[k];
          turb_stress = c->epsilon_half * dotv * Vj;
        }
        if (pass == 0) {
          for (int k = 0; k < dim; k++) rhs_v[k] -= mj * (p_star * (vsqdwi[k] + vsqdwj[k]) - turb_stress * dwdxi[k]);
        } else {
          double x_dot_star[3];
          for (int k = 0; k < dim; k++) x_dot_star[k] = veli[k] + 0.5 * dt * (rhs_v[k] + gravity[k]);
          double heat_tran = 0.;
          if (lans) {
            double sndspdab = 0.5 * (sndspdi + sndspdj);
            double Cp_ij = 0.5 * (st->Cp[j] + Cp_i);
            double hab = 0.5 * (hi + hj);
            double kij = sph_epsilon_heat_conductivity(c, Cp_ij, dx, velij, hab, sndspdab);
            double s_heat[3], dwdx_heat[3];
            for (int k = 0; k < dim; k++) s_heat[k] = si[k] * c->heat_trans_scale_ratio;
            for (int k = 0; k < dim; k++) {
              double a = psph_oracle_d_weight(dim, s_heat, hij / c->heat_trans_scale_ratio, c->cutoff, k);
              dwdx_heat[k] = 0.5 * (a + a);
            }
            double Fij = compute_F(dim, dwdx_heat, dx);
            heat_tran = mj * Vj * Vi * kij * (tempi - st->T[j]) * Fij;
          }
          double deltae = 0.;
          for (int k = 0; k < dim; k++)
            deltae -= mj * (v_star[k] - x_dot_star[k]) * (p_star * (vsqdwi[k] + vsqdwj[k]) - turb_stress * dwdxi[k]);
          rhs_e += (deltae + heat_tran);
        }
      }
    }
    double unew[3];
    for (int k = 0; k < dim; k++) unew[k] = D3(st->v, i, k) + dt * rhs_v[k];
    unew[dim - 1] += dt * gravity[dim - 1];
    for (int k = 0; k < dim; k++) D3(nv, i, k) = unew[k];
    ne[i] = st->e[i] + dt * (rhs_e + veli[dim - 1] * gravity[dim - 1]);
  }
  for (int64_t i = 0; i < st->n; i++)
    if (need_neigh(st, i)) {
      for (int k = 0; k < dim; k++) D3(st->v, i, k) = D3(nv, i, k);
      st->e[i] = ne[i];
      psph_oracle_eos(c, st, i);
    }
  free(nv); free(ne);
  return bad;
}

int psph_oracle_mom_engr_update(const psph_config *c, psph_oracle_state *st, double dt)
{
  if (c->use_gsph) return mom_engr_gsph(c, st, dt);
  return mom_engr_sph(c, st, dt);
}

int psph_oracle_calc_gradients(const psph_config *c, psph_oracle_state *st)
{ /* calculate_gradient.cc:22-209 (NORM_DERIVATIVE 1, DENSITY_UPDATE_SML 11: h = h_i) */
  int dim = c->dim, neq = dim + 2;
  for (int64_t i = 0; i < st->n; i++) {
    if (st->guest[i]) continue;
    double hi = st->h[i], supp = c->cutoff2 * hi;
    double AA[3][3] = { { 0 } }, FF[3][5] = { { 0 } }, ff[3][5] = { { 0 } };
    int nreal = 0;
    if (st->nb_off)
      for (int64_t q = st->nb_off[i]; q < st->nb_off[i + 1]; q++) {
        int64_t j = st->nb_idx[q];
        if (j == i) continue;
        double dx[3], s[3], dwm_rho[3];
        for (int k = 0; k < dim; k++) dx[k] = D3(st->x, i, k) - D3(st->x, j, k);
        if (!in_support(dim, dx, supp)) continue;
        nreal++;
        double mj = st->mass[j], rhoj = st->rho[j];
        for (int k = 0; k < dim; k++) s[k] = dx[k] / hi;
        for (int k = 0; k < dim; k++) dwm_rho[k] = psph_oracle_d_weight(dim, s, hi, c->cutoff, k) * mj / rhoj;
        for (int a = 0; a < dim; a++)
          for (int b = 0; b < dim; b++) AA[a][b] -= dx[a] * dwm_rho[b];
        for (int a = 0; a < dim; a++) {
          FF[a][0] -= ((st->rho[i] - st->rho[j]) * dwm_rho[a]);
          for (int b = 0; b < dim; b++) FF[a][b + 1] -= ((D3(st->v, i, b) - D3(st->v, j, b)) * dwm_rho[a]);
          FF[a][neq - 1] -= ((st->p[i] - st->p[j]) * dwm_rho[a]);
        }
      }
    if (nreal > 20) { /* linsolve, sph_lib.cc:158-181 */
      double Ainv[9];
      if (dim == 1) Ainv[0] = -1 / AA[0][0]; /* inv1 (sic), sph_lib.cc:128-134 */
      else { /* inv3, sph_lib.cc:86-110 */
        double A[9];
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) A[3 * a + b] = AA[a][b];
        double t4 = A[2] * A[3], t6 = A[2] * A[6], t8 = A[1] * A[3], t10 = A[1] * A[6], t12 = A[0] * A[4], t14 = A[0] * A[7];
        double t17 = 1.0 / (t4 * A[7] - t6 * A[4] - t8 * A[8] + t10 * A[5] + t12 * A[8] - t14 * A[5]);
        Ainv[0] = (A[4] * A[8] - A[7] * A[5]) * t17;
        Ainv[1] = -(A[3] * A[8] - A[6] * A[5]) * t17;
        Ainv[2] = (A[3] * A[7] - A[6] * A[4]) * t17;
        Ainv[3] = -(-A[2] * A[7] + A[1] * A[8]) * t17;
        Ainv[4] = (-t6 + A[0] * A[8]) * t17;
        Ainv[5] = -(-t10 + t14) * t17;
        Ainv[6] = (-A[2] * A[4] + A[1] * A[5]) * t17;
        Ainv[7] = -(-t4 + A[0] * A[5]) * t17;
        Ainv[8] = (-t8 + t12) * t17;
      }
      for (int a = 0; a < dim; a++)
        for (int q = 0; q < neq; q++) {
          ff[a][q] = 0;
          for (int k = 0; k < dim; k++) ff[a][q] += Ainv[a * dim + k] * FF[k][q];
        }
    }
    /* gradU[q*DIM + a] = ff[a][q]; rows of the 15-column layout: d_rho, d_u, d_v, d_w, d_p */
    double *g = st->grad + 15 * i;
    for (int a = 0; a < dim; a++) {
      g[0 + a] = ff[a][0];
      for (int b = 0; b < dim; b++) g[3 * (b + 1) + a] = ff[a][b + 1];
      g[12 + a] = ff[a][neq - 1];
    }
  }
  return 0;
}

/* ---------------- smooth_velocity ---------------- */
int psph_oracle_smooth_velocity(const psph_config *c, psph_oracle_state *st)
{ /* smooth_velocity.cc:25-165, HAVE_TURBULENCE_LANS==1 */
  int bad = 0, dim = c->dim;
  for (int64_t i = 0; i < st->n; i++) {
    if (!need_neigh(st, i)) continue;
    double hi = st->h[i], supp = c->cutoff2 * hi;
    double temp_dv[3] = { 0, 0, 0 }, wnorm = 0.;
    for (int64_t q = st->nb_off[i]; q < st->nb_off[i + 1]; q++) {
      int64_t j = st->nb_idx[q];
      if (!contr_vel_smooth(st, j)) continue;
      double ds[3], s[3];
      for (int k = 0; k < dim; k++) ds[k] = D3(st->x, i, k) - D3(st->x, j, k);
      if (!in_support(dim, ds, supp)) continue;
      for (int k = 0; k < dim; k++) s[k] = ds[k] / hi;
      double wght = psph_oracle_weight(dim, s, hi, c->cutoff);
      double mvw = wght * st->mass[j] / st->rho[j];
      for (int k = 0; k < dim; k++) temp_dv[k] += mvw * (D3(st->v, j, k) - D3(st->v, i, k));
      wnorm += mvw;
    }
    if (!(wnorm > 0)) bad++;
    for (int k = 0; k < dim; k++) temp_dv[k] /= wnorm;
    for (int k = 0; k < dim; k++) D3(st->sv, i, k) = D3(st->v, i, k) + temp_dv[k];
  }
  return bad;
}

/* ---------------- apply_bcond ---------------- */
static int contr_image(const psph_oracle_state *st, int64_t j)
{ /* particle.h:562-565: real and pressure-ghost particles contribute to a wall image */
  return st->bc[j] == 100 || st->bc[j] == 99 || st->bc[j] == 1;
}

int psph_oracle_apply_bcond(const psph_config *c, psph_oracle_state *st, int64_t nimg, const int32_t *img_row,
                            const int64_t *img_cell, const double *img_coord)
{ /* bcond.cc:24-249, BC_FOR_KX == 1 (options.h:269-271, :671-673).  For every wall-ghost image: Shepard
   * interpolation of (rho, v, e) at the reflection point over the contr_image() particles of the image's bucket
   * (:93-126) and of its 26 neighbours in table order (:129-188; ks only there, sic), reflection of the velocity
   * about the wall normal (:209), the ghost takes the values and its secondary variables are recomputed (:219-231) */
  const int dim = c->dim;
  for (int64_t q = 0; q < nimg; q++) {
    const double *refc = img_coord + 3 * q;
    const int64_t g = img_row[q];
    double normal[3] = { 0, 0, 0 }, d = 0.;
    for (int k = 0; k < dim; k++) { normal[k] = refc[k] - D3(st->x, g, k); d += normal[k] * normal[k]; } /* :66-74 */
    for (int k = 0; k < dim; k++) normal[k] /= sqrt(d);
    double uvec[5] = { 0, 0, 0, 0, 0 }, wnorm = 0., ks = 0.;
    const int64_t cl = img_cell[q];
    const int ck = (int) (cl % c->ncell[2]), cj = (int) ((cl / c->ncell[2]) % c->ncell[1]),
              ci = (int) (cl / ((int64_t) c->ncell[2] * c->ncell[1]));
    for (int pass = 0; pass < 28; pass++) { /* pass 0: the image's bucket; 1..27: neighbours, x-major / z-minor */
      int64_t nl = cl;
      if (pass) {
        const int t = pass - 1, di = t / 9 - 1, dj = (t / 3) % 3 - 1, dk = t % 3 - 1;
        if (!di && !dj && !dk) continue; /* neigh_proc == -2 for the bucket itself, preprocess.cc:303-304 */
        const int ni = ci + di, nj = cj + dj, nk = ck + dk;
        if (ni < 0 || nj < 0 || nk < 0 || ni >= c->ncell[0] || nj >= c->ncell[1] || nk >= c->ncell[2]) continue;
        nl = cell_lin(c, ni, nj, nk);
      }
      for (int64_t r = st->cell_off[nl]; r < st->cell_off[nl + 1]; r++) {
        const int64_t j = st->cell_rows[r];
        if (!contr_image(st, j)) continue;
        const double h = st->h[j];
        double dx[3] = { 0, 0, 0 }, s[3] = { 0, 0, 0 };
        for (int k = 0; k < dim; k++) dx[k] = refc[k] - D3(st->x, j, k);
        if (!in_support(dim, dx, c->cutoff2 * h)) continue; /* :104, :156 */
        for (int k = 0; k < dim; k++) s[k] = dx[k] / h;
        const double w = psph_oracle_weight(dim, s, h, c->cutoff);
        const double mj = st->mass[j], rj = st->rho[j];
        uvec[0] += mj * w;
        for (int k = 0; k < dim; k++) uvec[1 + k] += mj * w * D3(st->v, j, k) / rj;
        uvec[4] += mj * w * st->e[j] / rj;
        if (pass) ks += mj * w * st->ks[j] / rj; /* :172-174: neighbour buckets only */
        wnorm += mj * w / rj;
      }
    }
    if (wnorm > 0.) { /* :191-198 */
      for (int k = 0; k < 5; k++) uvec[k] /= wnorm;
      ks = ks / wnorm;
    } else {
      uvec[0] = 1.;
      for (int k = 1; k < 5; k++) uvec[k] = 0.;
      ks = 0.0;
    }
    double vdotn = 0., velrot[3] = { 0, 0, 0 }; /* reflect, sph_lib.cc:75-83 */
    for (int k = 0; k < dim; k++) vdotn += uvec[1 + k] * normal[k];
    for (int k = 0; k < dim; k++) velrot[k] = uvec[1 + k] - 2 * vdotn * normal[k];
    st->rho[g] = uvec[0];
    for (int k = 0; k < dim; k++) D3(st->v, g, k) = velrot[k];
    st->e[g] = uvec[4];
    st->ks[g] = ks;
    if (c->fluid_compressibility == 1) st->c[g] = 1482.; /* :223-225 */
    psph_oracle_eos(c, st, g);
  }
  return 0;
}

/* ---------------- update_pos ---------------- */
int psph_oracle_update_pos(const psph_config *c, psph_oracle_state *st, double dt, int erupting)
{ /* update_pos.cc:35-168 */
  int dim = c->dim, lans = c->have_turbulence_lans != 0;
  for (int64_t i = 0; i < st->n; i++) {
    if (st->removed && st->removed[i]) continue;
    if (is_real(st, i)) {
      for (int k = 0; k < dim; k++)
        D3(st->x, i, k) = D3(st->x, i, k) + dt * (lans ? D3(st->sv, i, k) : D3(st->v, i, k));
    } else if (st->bc[i] == 0 && erupting) {
      for (int k = 0; k < dim; k++) D3(st->x, i, k) = D3(st->x, i, k) + dt * D3(st->v, i, k);
      if (dim == 3 && D3(st->x, i, 2) >= c->erupt_top) { /* :134-142 */
        st->bc[i] = 100;
        st->h[i] = c->smlen_real;
        st->involved[i] = 2;
      }
    }
  }
  return 0;
}

/* ---------------- update_particles_in_bucket ---------------- */
static int determine_escape(const psph_config *c, const psph_cell_attr *a, const double *pos)
{ /* Bucket::determine_escape, bucket.cc:434-452 */
  for (int i = 0; i < c->dim; i++) {
    if (a->bucket_index[i * 2] == -1) {
      if ((pos[i] - a->bnd[i * 2]) < -c->tiny) return 1;
    } else if (a->bucket_index[i * 2 + 1] == 1) {
      if ((pos[i] - a->bnd[i * 2 + 1]) > c->tiny) return 1;
    }
  }
  return 0;
}

typedef struct { int32_t *v; int64_t n, cap; } ivec;
static void ivec_push(ivec *a, int32_t x)
{
  if (a->n == a->cap) { a->cap = a->cap ? 2 * a->cap : 8; a->v = (int32_t *) realloc(a->v, sizeof(int32_t) * a->cap); }
  a->v[a->n++] = x;
}

int psph_oracle_rebin(const psph_config *c, psph_oracle_state *st, int *lost)
{ /* update_particles_in_bucket, update_pos.cc:172-462, single rank (no guests).
   * new plist of a bucket = [arrivals from buckets visited earlier] ++ [its ghosts, old order]
   * ++ [its staying real/erupt particles, old order] ++ [arrivals from buckets visited later]
   * (add_particle appends to new_plist, bucket.h:642-645; put_new_plist appends too, :305-308). */
  if (c->dim != 3) return 0;
  int64_t nc = ncells_of(c);
  ivec *np = (ivec *) calloc(nc, sizeof(ivec));
  int nlost = 0;
  for (int64_t oi = 0; oi < nc; oi++) {
    int64_t cl = st->cell_order ? st->cell_order[oi] : oi;
    const psph_cell_attr *a = st->cell_attr ? &st->cell_attr[cl] : NULL;
    if (a && (a->type == PSPH_CELL_BRIEF || !a->active)) continue;
    int ck = (int) (cl % c->ncell[2]), cj = (int) ((cl / c->ncell[2]) % c->ncell[1]);
    int ci = (int) (cl / ((int64_t) c->ncell[2] * c->ncell[1]));
    int idx[3] = { ci, cj, ck };
    ivec ghosts = { 0, 0, 0 }, realp = { 0, 0, 0 };
    for (int64_t q = st->cell_off[cl]; q < st->cell_off[cl + 1]; q++) {
      int32_t p = st->cell_rows[q];
      if (!(is_real(st, p) || st->bc[p] == 0)) { ivec_push(&ghosts, p); continue; }
      double pos[3];
      for (int k = 0; k < 3; k++) pos[k] = D3(st->x, p, k);
      if (is_real(st, p) && a && a->type == PSPH_CELL_MIXED && determine_escape(c, a, pos)) {
        st->removed[p] = 1; nlost++; continue;
      }
      int left = 0, nidx[3];
      for (int k = 0; k < 3; k++) { /* :257-271 */
        nidx[k] = idx[k];
        if (pos[k] < cell_min(c, k, idx[k])) { nidx[k] = idx[k] - 1; left = 1; }
        else if (pos[k] >= cell_max(c, k, idx[k])) { nidx[k] = idx[k] + 1; left = 1; }
      }
      if (!left) { ivec_push(&realp, p); continue; }
      int inside = 1;
      for (int k = 0; k < 3; k++) if (nidx[k] < 0 || nidx[k] >= c->ncell[k]) inside = 0;
      if (!inside) { st->removed[p] = 1; nlost++; continue; }
      int64_t nl = cell_lin(c, nidx[0], nidx[1], nidx[2]);
      const psph_cell_attr *na = st->cell_attr ? &st->cell_attr[nl] : NULL;
      if (na && !na->has_involved) { st->removed[p] = 1; nlost++; continue; } /* :300-304 */
      ivec_push(&np[nl], p);
    }
    for (int64_t q = 0; q < ghosts.n; q++) ivec_push(&np[cl], ghosts.v[q]);
    for (int64_t q = 0; q < realp.n; q++) ivec_push(&np[cl], realp.v[q]);
    free(ghosts.v); free(realp.v);
  }
  /* second loop (:405-447): PRESS_BC buckets and inactive buckets keep their old plist */
  int64_t *noff = (int64_t *) malloc(sizeof(int64_t) * (nc + 1));
  int32_t *nrows = (int32_t *) malloc(sizeof(int32_t) * (st->n ? st->n : 1));
  noff[0] = 0;
  for (int64_t cl = 0; cl < nc; cl++) {
    const psph_cell_attr *a = st->cell_attr ? &st->cell_attr[cl] : NULL;
    int keep_old = a && (a->type == PSPH_CELL_BRIEF || !a->active || a->type == PSPH_CELL_PRESS_BC);
    int64_t w = noff[cl];
    if (keep_old) {
      for (int64_t q = st->cell_off[cl]; q < st->cell_off[cl + 1]; q++)
        if (!st->removed[st->cell_rows[q]]) nrows[w++] = st->cell_rows[q];
    } else
      for (int64_t q = 0; q < np[cl].n; q++) nrows[w++] = np[cl].v[q];
    noff[cl + 1] = w;
    free(np[cl].v);
  }
  memcpy(st->cell_off, noff, sizeof(int64_t) * (nc + 1));
  memcpy(st->cell_rows, nrows, sizeof(int32_t) * noff[nc]);
  free(noff); free(nrows); free(np);
  if (lost) *lost = nlost;
  return 0;
}

/* ---------------- THashTable iteration order ---------------- */
typedef struct { int64_t slot; uint32_t k[3]; int64_t row; } ord_t;
static int ord_cmp(const void *a, const void *b)
{
  const ord_t *x = (const ord_t *) a, *y = (const ord_t *) b;
  if (x->slot != y->slot) return x->slot < y->slot ? -1 : 1;
  for (int i = 0; i < 3; i++)
    if (x->k[i] != y->k[i]) return x->k[i] < y->k[i] ? -1 : 1;
  return 0;
}
void psph_oracle_table_order(const psph_config *c, const uint32_t *key, int64_t n, int64_t *perm)
{ /* thashtab.cc:204-222 (hash), :86-134 (sorted chains), :246-262 (iteration) */
  ord_t *o = (ord_t *) malloc(sizeof(ord_t) * (n ? n : 1));
  double umax = (double) 0xffffffffu;
  int S02 = c->ptable_size / 10, S01 = c->ptable_size - S02;
  for (int64_t i = 0; i < n; i++) {
    if (c->dim == 1) o[i].slot = key[3 * i];
    else o[i].slot = (int) ((double) key[3 * i] / umax * S01) + (int) ((double) key[3 * i + 1] / umax * S02);
    for (int k = 0; k < 3; k++) o[i].k[k] = key[3 * i + k];
    o[i].row = i;
  }
  qsort(o, n, sizeof(ord_t), ord_cmp);
  for (int64_t i = 0; i < n; i++) perm[i] = o[i].row;
  free(o);
}

/* ---------------- add_new_erupt ---------------- */
static double erupt_vel(const psph_erupt *er, double rsq)
{ /* vel_prof, sph_lib.cc:1223-1243 */
  if (er->vel_prof == 0) return er->Vv0;
  double umax = er->Vv0 * 2;
  return umax * (1 - (rsq / (er->rv * er->rv)));
}

static int erupt_column(const psph_config *c, const psph_erupt *er, const double *boxes, int nbox, double t_prev, double dt,
                        uint32_t add_step, double x, double y, psph_particle *out, int cap, int n_out)
{ /* the body of the four y-loops, setup_erupt.cc:458-490 (and its three copies) + the bucket filter :609-650 */
  double dist = x * x + y * y;
  if (!(dist <= er->rv * er->rv)) return n_out;
  double vel = erupt_vel(er, dist);
  double t_each = er->sml / vel;
  double t_add = fmod(t_prev, t_each) + dt;
  int n = (int) floor(t_add / t_each);
  t_add = t_add - n * t_each;
  for (int i = 0; i < n; i++) {
    double crd[3] = { x, y, er->bot + (t_add + i * t_each) * vel };
    int inside = 0; /* in_bucket (sph_lib.cc:1057-1065, both faces inclusive) of an erupt bucket */
    for (int b = 0; b < nbox && !inside; b++) {
      const double *mn = boxes + 6 * b, *mx = mn + 3;
      inside = crd[0] <= mx[0] && crd[0] >= mn[0] && crd[1] <= mx[1] && crd[1] >= mn[1] && crd[2] <= mx[2] && crd[2] >= mn[2];
    }
    if (!inside) continue;
    if (n_out >= cap) return -1;
    psph_particle *p = out + n_out++;
    memset(p, 0, sizeof *p);
    double normc[3];
    for (int k = 0; k < 3; k++) normc[k] = (crd[k] - er->kmin[k]) / (er->kmax[k] - er->kmin[k]);
    psph_oracle_hsfc3d(normc, p->key); /* THSFC3d, hilbert.c:329-359 */
    p->key[2] = add_step;
    /* Particle ctor for erupted material, particle.cc:210-316 */
    p->bc_type = 0; p->phase_num = 2; p->involved = 0; p->guest = 0; p->myprocess = c->rank;
    p->mass = er->mass;
    p->smlen = er->ratio_sml_dx * er->sml;
    for (int k = 0; k < 3; k++) p->coords[k] = crd[k];
    p->vel[2] = vel;
    p->smoothed_e = er->ev0;
    p->engr = er->ev0;
    p->mass_frac = 1.;
    p->rho = er->rhov;
    { /* update_second_var, particle.cc:528-604, with ks = 1; pressure, sound speed and gamma are then overwritten */
      double ks = 1., na = 1 - ks, ng = c->ng0 * ks, ns = ks - ng;
      double Cvm = ns * c->Cvs + ng * c->Cvg + na * c->Cva;
      p->phase_density[0] = p->rho * na;
      p->phase_density[1] = p->rho * ks;
      p->temperature = p->engr / Cvm;
      p->specific_heat_p = c->have_turbulence_lans != 0 ? Cvm : 0.;
    }
    p->pressure = er->pv0;
    p->sound_speed = er->sndspd;
    p->gamma = er->gamma_v;
  }
  return n_out;
}

int psph_oracle_add_new_erupt(const psph_config *c, const psph_erupt *er, const double *boxes, int nbox, double time,
                              double dt, uint32_t add_step, psph_particle *out, int cap)
{ /* setup_erupt.cc:390-670.  The duct's rough range is [-rv, rv] about the ORIGIN (:446-449) and the distance to the
   * axis is that of the origin too (:463-465), whatever Pos_v_P is; coordinates are built by repeated addition */
  const double sml = er->sml, sml2 = 0.5 * sml, t_prev = time - dt; /* :454 */
  int n = 0;
  for (int half = 0; half < 2 && n >= 0; half++) {
    double x = half == 0 ? er->pos_v[0] + sml2 : er->pos_v[0] - sml2;
    while (n >= 0 && (half == 0 ? x <= er->rv : x >= -er->rv)) {
      double y = er->pos_v[1] + sml2;
      while (n >= 0 && y <= er->rv) { n = erupt_column(c, er, boxes, nbox, t_prev, dt, add_step, x, y, out, cap, n); y += sml; }
      y = er->pos_v[1] - sml2;
      while (n >= 0 && y >= -er->rv) { n = erupt_column(c, er, boxes, nbox, t_prev, dt, add_step, x, y, out, cap, n); y -= sml; }
      if (half == 0) x += sml; else x -= sml;
    }
  }
  return n;
}

/* ---------------- domain growth: scan_outside_layer, adapt_domain, add_pressure_ghost ---------------- */
/* the 26 neighbours of a bucket that exist (neigh_proc > -1: inside the mesh) and are not brief */
static int grow_neigh(const psph_config *c, const psph_cell_attr *attr, int64_t cl, int q, int64_t *nl)
{
  int ck = (int) (cl % c->ncell[2]), cj = (int) ((cl / c->ncell[2]) % c->ncell[1]);
  int ci = (int) (cl / ((int64_t) c->ncell[2] * c->ncell[1]));
  int ni = ci + q / 9 - 1, nj = cj + (q / 3) % 3 - 1, nk = ck + q % 3 - 1;
  if (q == 13) return 0;
  if (ni < 0 || nj < 0 || nk < 0 || ni >= c->ncell[0] || nj >= c->ncell[1] || nk >= c->ncell[2]) return 0;
  *nl = cell_lin(c, ni, nj, nk);
  return attr[*nl].type != PSPH_CELL_BRIEF;
}

int psph_oracle_scan_outside_layer(const psph_config *c, const psph_oracle_state *st, psph_cell_attr *attr)
{ /* scan_outside_layer.cc:44-121 */
  int adapt = 0;
  int64_t nc = ncells_of(c);
  for (int64_t cl = 0; cl < nc; cl++) {
    psph_cell_attr *a = &attr[cl];
    if (a->type == PSPH_CELL_BRIEF || a->has_involved != 1) continue;
    int found = 0;
    for (int64_t q = st->cell_off[cl]; q < st->cell_off[cl + 1] && !found; q++) found = st->involved[st->cell_rows[q]] == 2;
    if (!found) continue;
    a->has_involved = 3; /* HAS_BOTH */
    for (int q = 0; q < 27; q++) {
      int64_t nl;
      if (!grow_neigh(c, attr, cl, q, &nl)) continue;
      if (st->cell_off[nl + 1] > st->cell_off[nl] && !attr[nl].has_involved && attr[nl].type != PSPH_CELL_UNDERGROUND) { adapt++; break; }
    }
  }
  return adapt;
}

int psph_oracle_adapt_domain(const psph_config *c, psph_oracle_state *st, psph_cell_attr *attr)
{ /* adapt_domain.cc:41-194 */
  int turned = 0;
  int64_t nc = ncells_of(c);
  for (int64_t cl = 0; cl < nc; cl++) {
    psph_cell_attr *a = &attr[cl];
    if (a->type == PSPH_CELL_BRIEF) continue;
    if (!(!a->has_involved && st->cell_off[cl + 1] > st->cell_off[cl] && a->type != PSPH_CELL_UNDERGROUND && a->type != PSPH_CELL_PRESS_BC))
      continue;
    int hit = 0;
    for (int q = 0; q < 27 && !hit; q++) {
      int64_t nl;
      if (grow_neigh(c, attr, cl, q, &nl) && (attr[nl].has_involved & 2)) hit = 1;
    }
    if (!hit) continue;
    a->has_involved = (a->has_involved & 1) | 1; /* set_has_involved(false); set_has_potential_involved(true) */
    a->flags |= PSPH_CELLF_REAL;
    int mode = a->type == PSPH_CELL_MIXED ? (a->bucket_index[4] == -1 ? 1 : 2) : 3;
    if (mode != 2) a->flags &= ~PSPH_CELLF_PRESS;
    for (int64_t q = st->cell_off[cl]; q < st->cell_off[cl + 1]; q++) {
      int32_t p = st->cell_rows[q];
      int turn;
      if (mode == 1) turn = st->bc[p] == 1;                            /* on-ground MIXED: its pressure ghosts (:113-118) */
      else if (mode == 2) {                                            /* MIXED on a side or on top: rows inside the domain */
        double pos[3];
        for (int k = 0; k < 3; k++) pos[k] = D3(st->x, p, k);
        turn = !determine_escape(c, a, pos);
      } else turn = 1;                                                 /* OVERGROUND: every row (:163-170) */
      if (turn) { st->bc[p] = 100; st->involved[p] = 1; turned++; }
    }
  }
  return turned;
}

/* Meteo::interpolate (meteo.cc:64-147): the sample itself, the end values outside the range, else linear */
static double meteo_interp(const double *z, const double *v, int n, double z0)
{
  for (int i = 0; i < n; i++) if (z[i] == z0) return v[i];
  if (z0 < z[0]) return v[0];
  if (z0 > z[n - 1]) return v[n - 1];
  int lo = 0;
  while (z[lo] < z0) lo++;            /* lower_bound: first z >= z0 (== upper_bound here: z0 is not a sample) */
  int li = lo - 1, ui = lo;
  return v[li] + (z0 - z[li]) / (z[ui] - z[li]) * (v[ui] - v[li]);
}

/* one new air particle of a growth step: Particle ctor for initial air (particle.cc:116-190) + initial_air
 * (sph_lib.cc:831-872); bc = 1 pressure ghost, 2 wall ghost */
static void grow_new_air(const psph_config *c, const psph_growth *g, const double *mz, const double *mrho, const double *mp,
                         uint32_t add_step, const double *pcrd, int bc, psph_particle *p)
{
  memset(p, 0, sizeof *p);
  double normc[3];
  for (int d = 0; d < 3; d++) normc[d] = (pcrd[d] - g->kmin[d]) / (g->kmax[d] - g->kmin[d]);
  psph_oracle_hsfc3d(normc, p->key);
  p->key[2] = add_step;
  p->bc_type = bc; p->phase_num = 1; p->involved = 0; p->guest = 0; p->myprocess = c->rank;
  p->smlen = g->ratio_sml_dx * g->smlen;
  for (int d = 0; d < 3; d++) p->coords[d] = pcrd[d];
  p->mass_frac = 0.;
  double sml2 = 0.5 * p->smlen, range[6];
  for (int d = 0; d < 3; d++) { range[2 * d] = pcrd[d] - sml2; range[2 * d + 1] = pcrd[d] + sml2; }
  double prss, dens, erg, mss;
  if (g->atmosphere_type == 4) {                               /* air_prop_meteo_based, :737-766 */
    prss = meteo_interp(mz, mp, g->nmeteo, pcrd[2] * 0.001) * 100;
    dens = meteo_interp(mz, mrho, g->nmeteo, pcrd[2] * 0.001);
    erg = prss / (dens * (g->gamma_air - 1));
    double d1 = meteo_interp(mz, mrho, g->nmeteo, range[4] * 0.001), d2 = meteo_interp(mz, mrho, g->nmeteo, range[5] * 0.001);
    mss = (range[1] - range[0]) * (range[3] - range[2]) * (range[5] - range[4]) * (0.1666667 * d1 + 0.6666666 * dens + 0.1666667 * d2);
  } else {                                                     /* air_prop_uniform, :682-712 (ideal gas) */
    prss = c->pa0;
    dens = prss / (c->Ra * g->Ta0);
    erg = prss / (dens * (g->gamma_air - 1));
    mss = (range[1] - range[0]) * (range[3] - range[2]) * (range[5] - range[4]) * dens;
  }
  p->rho = dens; p->engr = erg; p->mass = mss;
  { /* update_second_var (particle.cc:528-604) with ks = 0, ideal-gas branch; the ctor's gamma and sound speed are
     * overwritten there */
    double ks = 0., na = 1 - ks, ng = c->ng0 * ks, ns = ks - ng;
    double Cvm = ns * c->Cvs + ng * c->Cvg + na * c->Cva;
    double Rm = ng * c->Rg + na * c->Ra, Rmg = Rm / (ng + na);
    p->gamma = 1 + Rm / Cvm;
    double desmg = dens * (na + ng);
    p->pressure = Rmg * desmg * erg / Cvm;
    p->sound_speed = pow(p->gamma * p->pressure / desmg, 0.5);
    p->phase_density[0] = dens * na;
    p->phase_density[1] = dens * ks;
    p->temperature = erg / Cvm;
    p->specific_heat_p = c->have_turbulence_lans != 0 ? Cvm : 0.;
  }
}

int psph_oracle_add_pressure_ghost(const psph_config *c, const psph_oracle_state *st, psph_cell_attr *attr, const psph_growth *g,
                                   const double *mz, const double *mrho, const double *mp, const double *cell_mincrd, uint32_t add_step,
                                   psph_particle *out, int cap)
{ /* add_pressure_ghost.cc:36-316.  cell_mincrd [ncells][3]: Bucket::mincrd as the driver stores it (the lattice formula
   * when NULL; the reference's stored bounds differ from it by a few float ulps in x: the new particles sit on them) */
  int n_out = 0;
  int64_t nc = ncells_of(c);
  const double dx = g->smlen, dx2 = 0.5 * dx;
  (void) st;
  for (int64_t cl = 0; cl < nc; cl++) {
    psph_cell_attr *a = &attr[cl];
    if (a->type == PSPH_CELL_BRIEF) continue;
    if (!(!a->has_involved && !(a->flags & PSPH_CELLF_PRESS) && PSPH_CELLF_ERUPT_FLAG(a->flags) == 0)) continue;
    int hit = 0;
    for (int q = 0; q < 27 && !hit; q++) {
      int64_t nl;
      if (grow_neigh(c, attr, cl, q, &nl) && attr[nl].has_involved) hit = 1;
    }
    if (!hit) continue;
    a->active = 1;
    if (!(a->type == PSPH_CELL_MIXED || a->type == PSPH_CELL_OVERGROUND || a->type == PSPH_CELL_PRESS_BC)) continue;
    int ck = (int) (cl % c->ncell[2]), cj = (int) ((cl / c->ncell[2]) % c->ncell[1]);
    int ci = (int) (cl / ((int64_t) c->ncell[2] * c->ncell[1]));
    int idx[3] = { ci, cj, ck }, N[3];
    double mincrd[3];
    for (int k = 0; k < 3; k++) {
      mincrd[k] = cell_mincrd ? cell_mincrd[3 * cl + k] : cell_min(c, k, idx[k]);
      N[k] = (int) round(c->del / c->len_scale / dx);               /* (maxcrd - mincrd) / dx: one bucket edge */
    }
    const int ground = a->type == PSPH_CELL_MIXED && a->bucket_index[4] == -1;
    for (int i = 0; i < N[0]; i++)
      for (int j = 0; j < N[1]; j++)
        for (int k = 0; k < N[2]; k++) {
          double pcrd[3] = { mincrd[0] + dx2 + i * dx, mincrd[1] + dx2 + j * dx, mincrd[2] + dx2 + k * dx };
          if (ground && !(pcrd[2] > g->Ll[2])) continue;             /* :139 */
          if (n_out >= cap) return -1;
          grow_new_air(c, g, mz, mrho, mp, add_step, pcrd, 1, out + n_out++);
          a->flags |= PSPH_CELLF_PRESS;
        }
  }
  return n_out;
}

int psph_oracle_add_wall_ghost(const psph_config *c, psph_cell_attr *attr, const int32_t *cell_count, const psph_growth *g,
                               const double *mz, const double *mrho, const double *mp, const double *cell_mincrd, uint32_t add_step,
                               psph_particle *out, int cap)
{ /* add_wall_ghost.cc:44-293: every non-empty on-ground MIXED bucket without wall ghosts (and outside the vent) gets the
   * lattice points below the ground, and the bucket under it — marked active — a full lattice; all bc_type 2 */
  int n_out = 0;
  int64_t nc = ncells_of(c);
  const double dx = g->smlen, dx2 = 0.5 * dx;
  const int N = (int) round(c->del / c->len_scale / dx);
  for (int64_t cl = 0; cl < nc; cl++) {
    psph_cell_attr *a = &attr[cl];
    if (a->type == PSPH_CELL_BRIEF) continue;
    if (!(a->bucket_index[4] == -1 && a->type == PSPH_CELL_MIXED && cell_count[cl] > 0 && !(a->flags & PSPH_CELLF_WALL) &&
          PSPH_CELLF_ERUPT_FLAG(a->flags) == 0))
      continue;
    int ck = (int) (cl % c->ncell[2]), cj = (int) ((cl / c->ncell[2]) % c->ncell[1]);
    int ci = (int) (cl / ((int64_t) c->ncell[2] * c->ncell[1]));
    for (int pass = 0; pass < 2; pass++) {                           /* the bucket itself, then the one below (:219-283) */
      if (pass == 1 && ck == 0) return -2;                           /* assert(Down_buck) */
      int64_t tl = pass == 0 ? cl : cell_lin(c, ci, cj, ck - 1);
      psph_cell_attr *ta = &attr[tl];
      if (pass == 1) { if (ta->type == PSPH_CELL_BRIEF) return -2; ta->active = 1; }
      int idx[3] = { ci, cj, pass == 0 ? ck : ck - 1 };
      double mincrd[3];
      for (int k = 0; k < 3; k++) mincrd[k] = cell_mincrd ? cell_mincrd[3 * tl + k] : cell_min(c, k, idx[k]);
      for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++)
          for (int k = 0; k < N; k++) {
            double pcrd[3] = { mincrd[0] + dx2 + i * dx, mincrd[1] + dx2 + j * dx, mincrd[2] + dx2 + k * dx };
            if (pass == 0 && !(pcrd[2] < g->Ll[2])) continue;        /* :163: below the ground only */
            if (n_out >= cap) return -1;
            grow_new_air(c, g, mz, mrho, mp, add_step, pcrd, 2, out + n_out++);
            ta->flags |= PSPH_CELLF_WALL; ta->active = 1;            /* add_wall_ghost_particle, bucket.h:656-661 */
          }
    }
  }
  return n_out;
}

/* ---------------- Hilbert key ---------------- */
void psph_oracle_hsfc3d(const double coord[3], uint32_t key[2])
{ /* HSFC3d + hsfc3d, hilbert.c:156-271,289-303 (Zoltan's 3-D Hilbert curve), restated as a
   * 63-bit index walk: 21 levels, per level the octant is Gray-decoded under the current axis
   * permutation/reflection, then the frame is rotated by a per-octant rule.  The reference
   * stores level i, bit n at bit offset 3i+n of a 2-word key, i.e. the 63-bit index shifted
   * right by one (its last bit falls off the end of the 2-word key). */
  static const unsigned char gray_inv[8] = { 0, 1, 3, 2, 7, 6, 4, 5 };
  /* new axis slot a takes old slot src[c][a], reflected if flip[c][a] */
  static const unsigned char src[8][3] = { {2,1,0}, {0,2,1}, {0,1,2}, {2,0,1}, {2,0,1}, {0,1,2}, {0,2,1}, {2,1,0} };
  static const unsigned char flip[8][3] = { {0,0,0}, {0,0,0}, {0,0,0}, {1,1,0}, {0,1,1}, {0,0,0}, {0,1,1}, {1,0,1} };
  const uint32_t imax = 0xffffffffu;
  uint32_t cc[3];
  for (int d = 0; d < 3; d++) {
    double v = (coord[d] <= 0) ? 0 : imax * coord[d];
    cc[d] = (uint32_t) ((imax < v) ? imax : v);
  }
  unsigned char ax[3] = { 0, 2, 4 }; /* (dimension << 1) | reflect */
  uint64_t H = 0;
  for (int i = 1; i <= 21; i++) {
    unsigned s = 32 - i, g = 0;
    for (int a = 0; a < 3; a++) g |= (((cc[ax[a] >> 1] >> s) ^ ax[a]) & 1u) << a;
    unsigned oct = gray_inv[g];
    H = (H << 3) | oct;
    unsigned char old[3] = { ax[0], ax[1], ax[2] };
    for (int a = 0; a < 3; a++) ax[a] = old[src[oct][a]] ^ flip[oct][a];
  }
  uint64_t K = H >> 1;
  key[0] = (uint32_t) (K >> 32);
  key[1] = (uint32_t) (K & 0xffffffffu);
}
