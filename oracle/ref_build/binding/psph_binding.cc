/* The reference-side binding of INTEGRATION.md, as a translation unit that COMPILES against the reference's own headers
 * (checked by tests/test_oracle_vs_ref.py::test_reference_side_binding_compiles in the build container; the reference
 * tree is not on the GPU box).  Nothing here is product code or oracle code: it is the glue a Plume-SPH maintainer
 * adds to src/plume/main.cc under -DUSE_PSPH_B200 — Particle / Bucket / BndImage objects to the plain records of
 * include/psph.h and back.  3-D builds. */
#include <vector>
#include <list>
using namespace std;

#include <mpi.h>
#include <hashtab.h>
#include <thashtab.h>
#include <bucket.h>
#include <particle.h>
#include <properties.h>
#include <bnd_image.h>
#include "constant.h"
#include "parameters.h"
#include "sph_header.h"

#include <psph.h>

/* Particle -> psph_particle: the fields pack_particles() copies (pack_data.cc:180-246) + the EOS outputs */
static void to_row(Particle *p, psph_particle *r)
{
  TKey k = p->getKey();
  for (int i = 0; i < TKEYLENGTH; i++) r->key[i] = k.key[i];
  r->bc_type = p->get_bc_type(); r->phase_num = p->which_phase();
  r->involved = p->get_involved(); r->guest = p->is_guest(); r->myprocess = p->get_my_processor();
  r->mass = p->get_mass(); r->smlen = p->get_smlen(); r->mass_frac = p->get_mass_frac();
  r->pressure = p->get_pressure(); r->gamma = p->get_gamma(); r->sound_speed = p->get_sound_speed();
  r->temperature = p->get_temperature(); r->specific_heat_p = p->get_specific_heat_p();
  r->phase_density[0] = p->get_phase_density()[0]; r->phase_density[1] = p->get_phase_density()[1];
  for (int i = 0; i < DIMENSION; i++) {
    r->coords[i] = p->get_coords()[i]; r->vel[i] = p->get_vel()[i];
    r->smoothed_v[i] = p->get_smoothed_velocity()[i];
  }
  r->rho = p->get_density(); r->engr = p->get_energy(); r->smoothed_e = p->get_smoothed_energy();
}

/* every particle of this rank, in THashTable iteration order */
size_t psph_bind_collect_rows(THashTable *P_table, vector<psph_particle> &rows)
{
  rows.clear();
  THTIterator itr(P_table);
  Particle *p;
  while ((p = (Particle *) itr.next())) {
    psph_particle r = psph_particle();
    to_row(p, &r);
    rows.push_back(r);
  }
  return rows.size();
}

/* bucket attributes over the whole lattice (x-major, z-minor); cells without a Bucket object stay BRIEF (type 0) */
void psph_bind_collect_cells(HashTable *BG_mesh, const psph_config *cfg, vector<psph_cell_attr> &attr)
{
  attr.assign((size_t) cfg->ncell[0] * cfg->ncell[1] * cfg->ncell[2], psph_cell_attr());
  HTIterator it(BG_mesh);
  void *t;
  while ((t = it.next())) {
    BriefBucket *bb = (BriefBucket *) t;
    if (bb->check_brief()) continue;
    Bucket *b = (Bucket *) t;
    int idx[3];
    for (int d = 0; d < 3; d++)
      idx[d] = (int) floor((*(b->get_mincrd() + d) * cfg->len_scale - cfg->mindom[d]) / cfg->del + 0.5);
    psph_cell_attr &a = attr[((size_t) idx[0] * cfg->ncell[1] + idx[1]) * cfg->ncell[2] + idx[2]];
    a.type = b->get_bucket_type(); a.has_involved = b->get_has_involved(); a.active = b->is_active();
    for (int k = 0; k < 6; k++) { a.bucket_index[k] = b->get_bucket_index()[k]; a.bnd[k] = b->get_bnd()[k]; }
    a.flags = (b->has_real_particles() ? PSPH_CELLF_REAL : 0) | (b->has_wall_ghost_particles() ? PSPH_CELLF_WALL : 0) |
              (b->has_pressure_ghost_particles() ? PSPH_CELLF_PRESS : 0) | (b->has_erupt_ghost_particles() ? PSPH_CELLF_ERUPT : 0) |
              ((int) b->get_erupt_flag() << 8);                       /* read by the domain-growth calls */
  }
}

/* Image_table -> psph_image records (psph_set_images) */
void psph_bind_collect_images(HashTable *BG_mesh, const psph_config *cfg, list<BndImage> &Image_table, vector<psph_image> &img)
{
  img.clear();
  for (list<BndImage>::iterator it = Image_table.begin(); it != Image_table.end(); ++it) {
    Bucket *b = (Bucket *) BG_mesh->lookup(it->bucket_key);
    if (!b) continue;
    psph_image m = psph_image();
    for (int k = 0; k < TKEYLENGTH; k++) m.ghost_key[k] = it->ghost_key[k];
    int idx[3];
    for (int d = 0; d < 3; d++) {
      idx[d] = (int) floor((*(b->get_mincrd() + d) * cfg->len_scale - cfg->mindom[d]) / cfg->del + 0.5);
      m.coord[d] = it->coord[d];
    }
    m.cell = (idx[0] * cfg->ncell[1] + idx[1]) * cfg->ncell[2] + idx[2];
    img.push_back(m);
  }
}

/* vent record + erupt buckets (psph_set_erupt), from the constants of parameters.h and SimProps */
void psph_bind_collect_erupt(THashTable *P_table, HashTable *BG_mesh, SimProps *simprops, psph_erupt *er, vector<double> &boxes)
{
  er->pos_v[0] = Pos_v_P[0]; er->pos_v[1] = Pos_v_P[1];
  er->rv = rv_P; er->Vv0 = Vv0_P; er->vel_prof = ERUPT_VELOCITY_PROF; er->pad = 0;
  er->bot = simprops->get_bot(); er->sml = simprops->sml_of_phase2; er->mass = simprops->mass_of_phase2;
  er->ratio_sml_dx = RATIO_SML_DX;
  er->ev0 = ev0_P; er->rhov = rhov_P; er->pv0 = pv0_P; er->gamma_v = gamma_v_P;
#if FLUID_COMPRESSIBILITY == 0
  er->sndspd = 340.;
#else
  er->sndspd = 1482.;
#endif
  for (int d = 0; d < 3; d++) { er->kmin[d] = *(P_table->get_minDom() + d); er->kmax[d] = *(P_table->get_maxDom() + d); }
  boxes.clear();
  HTIterator it(BG_mesh);
  void *t;
  while ((t = it.next())) {
    BriefBucket *bb = (BriefBucket *) t;
    if (bb->check_brief()) continue;
    Bucket *b = (Bucket *) t;
    if (!b->is_erupt()) continue;
    for (int d = 0; d < 3; d++) boxes.push_back(*(b->get_mincrd() + d));
    for (int d = 0; d < 3; d++) boxes.push_back(*(b->get_maxcrd() + d));
  }
}

/* what add_pressure_ghost reads besides the tables (psph_set_growth): lattice constants, key box of the particle table,
 * and the buckets' stored mincrd (the new particles sit on mincrd + dx / 2 + i dx, add_pressure_ghost.cc:161-163); the
 * sounding's columns (height [km], density, pressure [hPa]) are the driver's own Meteo table (src/datstr/meteo.h) */
void psph_bind_collect_growth(THashTable *P_table, HashTable *BG_mesh, MatProps *matprops, const psph_config *cfg, psph_growth *g,
                              vector<double> &mincrd)
{
  g->mass = matprops->particle_mass; g->smlen = matprops->smoothing_length; g->ratio_sml_dx = RATIO_SML_DX;
  g->gamma_air = gamma_P; g->sndspd = 340.; g->gmm = 1.4; g->Ta0 = Ta0_P;
  for (int d = 0; d < 3; d++) {
    g->kmin[d] = *(P_table->get_minDom() + d); g->kmax[d] = *(P_table->get_maxDom() + d);
    g->Ll[d] = Ll_P[d]; g->Lu[d] = Lu_P[d];
  }
  g->atmosphere_type = ATMOSPHERE_TYPE; g->nmeteo = 0;                /* nmeteo: rows of the sounding the driver hands over */
  mincrd.assign((size_t) cfg->ncell[0] * cfg->ncell[1] * cfg->ncell[2] * 3, 0.);
  for (int i = 0; i < cfg->ncell[0]; i++)
    for (int j = 0; j < cfg->ncell[1]; j++)
      for (int k = 0; k < cfg->ncell[2]; k++) {
        const int idx[3] = { i, j, k };
        for (int d = 0; d < 3; d++)
          mincrd[(((size_t) i * cfg->ncell[1] + j) * cfg->ncell[2] + k) * 3 + d] = (cfg->mindom[d] + idx[d] * cfg->del) / cfg->len_scale;
      }
  HTIterator it(BG_mesh);
  void *t;
  while ((t = it.next())) {
    BriefBucket *bb = (BriefBucket *) t;
    if (bb->check_brief()) continue;
    Bucket *b = (Bucket *) t;
    int idx[3];
    for (int d = 0; d < 3; d++)
      idx[d] = (int) floor((*(b->get_mincrd() + d) * cfg->len_scale - cfg->mindom[d]) / cfg->del + 0.5);
    for (int d = 0; d < 3; d++)
      mincrd[(((size_t) idx[0] * cfg->ncell[1] + idx[1]) * cfg->ncell[2] + idx[2]) * 3 + d] = *(b->get_mincrd() + d);
  }
}

/* the end of a step and the first half of the adapt block (main.cc:435, :256-271) on the device */
int psph_bind_grow(psph_ctx *gpu, TimeProps *timeprops, int *adapt_next)
{
  int turned = 0, added = 0;
  if (*adapt_next) {                                                  /* :259: adapt of the previous step's scan */
    if (psph_adapt_domain(gpu, &turned) < 0) return -1;               /* :265 */
    if (psph_add_pressure_ghost(gpu, (uint32_t) timeprops->step, &added) < 0) return -1; /* :271 */
    /* shift_brief_buck (:274) stays with the driver: psph_get_cells brings the attributes back, psph_set_cells takes the new
     * buckets down; then */
    if (psph_add_wall_ghost(gpu, (uint32_t) timeprops->step, &added) < 0) return -1;     /* :277 */
    /* update_bgmesh, search_bnd_images (:283-292) on the host again -> psph_set_cells / psph_set_images */
  }
  return psph_scan_outside_layer(gpu, adapt_next);                    /* :435 */
}

/* one time step of main.cc:217-471 with the stage calls forwarded (single rank) */
int psph_bind_step(psph_ctx *gpu, TimeProps *timeprops, int *lost)
{
  double dt = 0.;
  if (psph_timestep(gpu, &dt) < 0) return -1;                         /* main.cc:220 */
  timeprops->incrtime(&dt);                                           /* :231 */
  if (psph_set_dt(gpu, dt, timeprops->step, timeprops->iferupt()) < 0) return -1;
  int ierr = 0;
  if (psph_search_neighs(gpu) < 0) return -1;                         /* :318 */
  if (timeprops->step % SML_UPDATE_INT == 0 && psph_adaptive_sml(gpu) < 0) return -1; /* :322 */
  ierr += psph_mom_engr_update(gpu);                                  /* :358 */
  if (psph_smooth_density(gpu) < 0) return -1;                        /* :378 */
  if (psph_smooth_velocity(gpu) < 0) return -1;                       /* :392 */
  if (psph_update_pos(gpu) < 0) return -1;                            /* :405 */
  if (psph_rebin(gpu, lost) < 0) return -1;                           /* :408 */
  int added = 0;
  if (timeprops->iferupt() && psph_add_new_erupt(gpu, timeprops->time, dt, timeprops->step, &added) < 0) return -1; /* :414-417 */
  ierr += psph_apply_bcond(gpu);                                      /* :468 */
  return ierr;
}
