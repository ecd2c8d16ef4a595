/* psph_oracle.h — CPU restatement of the Plume-SPH per-timestep hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.  The product (plume_sph_b200/,
 * include/psph.h) never links, imports or falls back to it.
 *
 * Parity status: PINNED against the reference itself.  The reference has no tests or
 * golden vectors of its own, so the pins are outputs of the unmodified
 * reference compiled here (oracle/_ref, recipe oracle/ref_build/build_ref.py): per-stage
 * particle-table dumps and neighbour lists, committed as tests/golden/ *.npz together with
 * the generating script tests/golden/make_golden.py.  tests/test_oracle_vs_ref.py holds
 * the comparisons.
 *
 * Plain scalar C, one function per reference stage, each citing the file:line it restates.
 * Sums run in the reference's neighbour-list order so that results agree with the compiled
 * reference to the last bit wherever libm does (same glibc here).
 */
#ifndef PSPH_ORACLE_H
#define PSPH_ORACLE_H
#include <stdint.h>
#include "../include/psph.h" /* psph_config only (the option record); no product code */

#ifdef __cplusplus
extern "C" {
#endif

/* Particle table, struct of arrays, rows in THashTable iteration order on entry
 * (thashtab.cc:246-262).  Vector fields are [n][3] (unused components 0 in 1-D). */
typedef struct psph_oracle_state
{
  int64_t n;
  uint32_t *key;  /* [n][3] */
  int32_t *bc, *phase, *involved, *guest;
  double *mass, *h, *ks, *p, *gamma, *c, *T, *Cp;
  double *rho_ph; /* [n][2] */
  double *x, *v, *sv; /* [n][3] */
  double *rho, *e, *se;
  double *grad;   /* [n][15] */
  /* background mesh occupancy: plist of every cell, in the reference's plist order */
  int64_t ncells;
  int64_t *cell_off; /* [ncells+1] */
  int32_t *cell_rows;/* [n] */
  int32_t *cell_order; /* [ncells] iteration order of HTIterator (hashtab.cc:206-240); NULL = 0..ncells-1 */
  const psph_cell_attr *cell_attr; /* [ncells] or NULL */
  /* neighbour lists of the last search (malloc'ed by the oracle) */
  int64_t *nb_off;   /* [n+1] */
  int32_t *nb_idx;
  /* scratch: removed flag written by rebin ([n], 1 = deleted) */
  int32_t *removed;
} psph_oracle_state;

/* initial binning of rows into cells by position (plist = ascending row order) */
int psph_oracle_bin(const psph_config *cfg, psph_oracle_state *st);
/* search_neighs_consth, neighs.cc:37-211 */
int psph_oracle_search(const psph_config *cfg, psph_oracle_state *st);
void psph_oracle_free_lists(psph_oracle_state *st);
/* timestep, timestep.cc:95-135 / :140-183 (unclipped CFL dt) */
double psph_oracle_timestep(const psph_config *cfg, const psph_oracle_state *st);
/* adaptive_sml, neighs.cc:479-693 (ADAPTIVE_SML==1) */
int psph_oracle_adaptive_sml(const psph_config *cfg, psph_oracle_state *st);
/* calc_gradients, calculate_gradient.cc:22-209 */
int psph_oracle_calc_gradients(const psph_config *cfg, psph_oracle_state *st);
/* mom_engr_update, mom_engr_update.cc:30-439 (SPH) / :443-1145 (GSPH) */
int psph_oracle_mom_engr_update(const psph_config *cfg, psph_oracle_state *st, double dt);
/* smooth_density, density_massfrac.cc:25-359 */
int psph_oracle_smooth_density(const psph_config *cfg, psph_oracle_state *st);
/* smooth_velocity, smooth_velocity.cc:25-165 */
int psph_oracle_smooth_velocity(const psph_config *cfg, psph_oracle_state *st);
/* update_pos, update_pos.cc:35-168 */
int psph_oracle_update_pos(const psph_config *cfg, psph_oracle_state *st, double dt, int erupting);
/* update_particles_in_bucket, update_pos.cc:172-462: rewrites cell_off/cell_rows, flags
 * removed rows in st->removed (rows are NOT compacted; removed rows leave every plist) */
int psph_oracle_rebin(const psph_config *cfg, psph_oracle_state *st, int *lost);
/* apply_bcond, bcond.cc:24-249: img_row = row of the wall ghost, img_cell = linear index of the image's bucket,
 * img_coord [nimg][3] = reflection points */
int psph_oracle_apply_bcond(const psph_config *cfg, psph_oracle_state *st, int64_t nimg, const int32_t *img_row,
                            const int64_t *img_cell, const double *img_coord);
/* Particle::update_second_var, particle.cc:528-604, for one row */
void psph_oracle_eos(const psph_config *cfg, psph_oracle_state *st, int64_t i);
/* THashTable iteration order (thashtab.cc:204-222 hash, :86-134 sorted chains):
 * writes the permutation that sorts rows into reference order */
void psph_oracle_table_order(const psph_config *cfg, const uint32_t *key, int64_t n, int64_t *perm);
/* kernel functions, exposed for unit tests: weight sph_lib.cc:210-230, d_weight :240-253 */
double psph_oracle_weight(int dim, const double *s, double h, double cutoff);
double psph_oracle_d_weight(int dim, const double *s, double h, double cutoff, int dir);
/* add_new_erupt, setup_erupt.cc:390-670: writes the particles the call adds to out[] (generation order: x+ / x- halves,
 * y+ / y- runs of every x column, bottom first) and returns their number (-1: more than cap).  boxes: [nbox][6] mincrd,
 * maxcrd of the erupt buckets.  time = TimeProps::time after incrtime. */
int psph_oracle_add_new_erupt(const psph_config *cfg, const psph_erupt *er, const double *boxes, int nbox, double time,
                              double dt, uint32_t add_step, psph_particle *out, int cap);
/* Domain growth.  attr [ncells] is read AND written (has_involved, active, flags); the plists come from
 * st->cell_off / cell_rows.  Single rank: no guest buckets.
 * scan_outside_layer.cc:44-121 -> adapt;  adapt_domain.cc:41-194 -> rows turned real (st->bc, st->involved change);
 * add_pressure_ghost.cc:36-316 + initial_air (sph_lib.cc:831-872): writes the new particles to out[] in bucket order
 * (cells ascending, lattice i, j, k) and returns their number (-1: more than cap). */
int psph_oracle_scan_outside_layer(const psph_config *cfg, const psph_oracle_state *st, psph_cell_attr *attr);
int psph_oracle_adapt_domain(const psph_config *cfg, psph_oracle_state *st, psph_cell_attr *attr);
int psph_oracle_add_pressure_ghost(const psph_config *cfg, const psph_oracle_state *st, psph_cell_attr *attr, const psph_growth *g,
                                   const double *meteo_z, const double *meteo_rho, const double *meteo_p, const double *cell_mincrd,
                                   uint32_t add_step, psph_particle *out, int cap);
/* add_wall_ghost.cc:44-293; cell_count [ncells] = plist sizes at the time of the call (after add_pressure_ghost and
 * shift_brief_buck); -2: a bucket to fill has no bucket below it (the reference asserts) */
int psph_oracle_add_wall_ghost(const psph_config *cfg, psph_cell_attr *attr, const int32_t *cell_count, const psph_growth *g,
                               const double *meteo_z, const double *meteo_rho, const double *meteo_p, const double *cell_mincrd,
                               uint32_t add_step, psph_particle *out, int cap);
/* Hilbert key of a normalised position, hilbert.c:289-303 (HSFC3d) */
void psph_oracle_hsfc3d(const double coord[3], uint32_t key[2]);

#ifdef __cplusplus
}
#endif
#endif
