/* The UNMODIFIED reference driver with its per-step stage calls forwarded to libpsph.so — the drop-in of
 * INTEGRATION.md as a program (3-D SPH builds, single rank).
 *
 * Like ref_hooks.cc this translation unit textually includes /root/reference/src/plume/main.cc after renaming the
 * stage calls; here the wrappers do not call the reference functions but the C ABI of include/psph.h:
 *   timestep .. update_pos            -> psph_* on the device (the host table is not touched)
 *   update_particles_in_bucket        -> psph_rebin; positions, bc_type, smlen, involved come back to the host
 *                                        particles and the reference's own re-bin keeps its bucket lists in step
 *   add_new_erupt                     -> the reference's call (host tables) AND psph_add_new_erupt (device): same rows
 *   apply_bcond (end of step)         -> psph_apply_bcond, then the whole state comes back to the host particles
 * so that everything the driver keeps for itself (time bookkeeping, domain growth, ghost generation, image search,
 * output) sees current data, and whatever it changes is uploaded again before the next device call.
 * PSPH_LOCKSTEP=1: the host runs the reference's own stages as well (its table stays the pure reference trajectory), the
 * device is loaded with that table at the start of EVERY step and runs the same step beside it; at the end of each step
 * (after apply_bcond) the two tables are compared field by field and one "LOCKSTEP step field err" line per field goes to
 * stderr — one whole step from identical states, at every step, with domain growth, injection and wall ghosts in play
 * (end states of free-running programs part company for a reason that is not an error: see tests/test_gpu_dropin.py).
 * Input, set-up and output are the reference's.  PSPH_MAX_STEPS / PSPH_DUMP_DIR work as in ref_hooks.cc: the final
 * dump of this program against the final dump of particler_hooked is the end-to-end parity check of the boundary.
 *
 * Test infrastructure (needs a GPU to run; built by build_ref.py --b200 when libpsph.so exists). */
#define PSPH_HOOKS_NO_MAIN 1
#include "ref_hooks.cc"
#include <psph.h>

#if CODE_DIMENSION == 3
namespace {
psph_ctx *G = NULL;
psph_config g_cfg;
bool g_dirty = true;            /* host table changed since the last upload */
bool g_lock = false;            /* PSPH_LOCKSTEP: reference stages on the host beside the device's, compared every step */
bool g_in_loop = false;
TimeProps *g_tp = NULL;
MatProps *g_mp = NULL;
SimProps *g_sp = NULL;
vector<psph_particle> g_rows;

void die(const char *what)
{
  fprintf(stderr, "ref_b200: %s failed: %s\n", what, G ? psph_last_error(G) : "(no context)");
  exit(2);
}
#define CK(call) do { if ((call) < 0) die(#call); } while (0)

void to_row(Particle *p, psph_particle *r)
{
  memset(r, 0, sizeof *r);
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

/* device row -> host particle; geometry only (after the re-bin) or everything (end of step) */
void from_row(const psph_particle *r, Particle *p, bool all)
{
  double x[3] = { r->coords[0], r->coords[1], r->coords[2] };
  p->put_coords(x);
  p->put_bc_type(r->bc_type);
  if (r->smlen != p->get_smlen()) p->put_smlen(r->smlen / RATIO_SML_DX);
  p->set_involved_flag(r->involved);
  if (!all) return;
  double v[3] = { r->vel[0], r->vel[1], r->vel[2] }, sv[3] = { r->smoothed_v[0], r->smoothed_v[1], r->smoothed_v[2] };
  double ph[2] = { r->phase_density[0], r->phase_density[1] };
  p->put_vel(v);                                    /* (put_velocity() writes new_state_vars, particle.h:773-778) */
  p->put_smoothed_velocity(sv);
  p->put_density(r->rho); p->put_energy(r->engr); p->put_smoothed_energy(r->smoothed_e);
  p->put_mass_frac(r->mass_frac); p->put_pressure(r->pressure); p->put_gamma(r->gamma);
  p->put_sound_speed(r->sound_speed); p->put_temperature(r->temperature); p->put_specific_heat_p(r->specific_heat_p);
  p->put_phase_density(ph);
}

void download_to_host(THashTable *P, bool all)
{
  size_t n = 0;
  CK(psph_download(G, NULL, 0, &n, PSPH_ORDER_DEVICE));
  g_rows.resize(n);
  CK(psph_download(G, g_rows.data(), g_rows.size(), &n, PSPH_ORDER_DEVICE));
  for (size_t i = 0; i < n; i++) {
    Particle *p = (Particle *) P->lookup(g_rows[i].key);
    if (p) from_row(&g_rows[i], p, all);
  }
}

/* lock-step: device table against the host's (pure reference) table at the end of a step */
void lockstep_compare(THashTable *P)
{
  size_t n = 0, nhost = 0;
  CK(psph_download(G, NULL, 0, &n, PSPH_ORDER_DEVICE));
  g_rows.resize(n);
  CK(psph_download(G, g_rows.data(), g_rows.size(), &n, PSPH_ORDER_DEVICE));
  { THTIterator it(P); while (it.next()) nhost++; }
  const char *names[] = { "coord", "vel", "smoothed_v", "rho", "engr", "mass_frac", "pressure", "smlen", "sound_speed", "temperature",
                          "phase_density", "gamma", "specific_heat_p", "smoothed_e" };
  const int NF = sizeof names / sizeof names[0];
  double dmax[NF], smax[NF];
  for (int f = 0; f < NF; f++) { dmax[f] = 0.; smax[f] = 0.; }
  long long missing = 0, flags = 0;
  for (size_t i = 0; i < n; i++) {
    const psph_particle &r = g_rows[i];
    Particle *p = (Particle *) P->lookup(r.key);
    if (!p) { missing++; continue; }
    if (r.bc_type != p->get_bc_type() || r.involved != p->get_involved() || r.phase_num != p->which_phase()) flags++;
    psph_particle h;
    to_row(p, &h);
    const double *a[NF] = { r.coords, r.vel, r.smoothed_v, &r.rho, &r.engr, &r.mass_frac, &r.pressure, &r.smlen, &r.sound_speed,
                            &r.temperature, r.phase_density, &r.gamma, &r.specific_heat_p, &r.smoothed_e };
    const double *b[NF] = { h.coords, h.vel, h.smoothed_v, &h.rho, &h.engr, &h.mass_frac, &h.pressure, &h.smlen, &h.sound_speed,
                            &h.temperature, h.phase_density, &h.gamma, &h.specific_heat_p, &h.smoothed_e };
    const int len[NF] = { 3, 3, 3, 1, 1, 1, 1, 1, 1, 1, 2, 1, 1, 1 };
    for (int f = 0; f < NF; f++)
      for (int k = 0; k < len[f]; k++) {
        dmax[f] = std::max(dmax[f], fabs(a[f][k] - b[f][k]));
        smax[f] = std::max(smax[f], fabs(b[f][k]));
      }
  }
  fprintf(stderr, "LOCKSTEP %d rows device=%zu host=%zu missing=%lld flags=%lld\n", g_step, n, nhost, missing, flags);
  for (int f = 0; f < NF; f++)
    fprintf(stderr, "LOCKSTEP %d %s %.3e\n", g_step, names[f], smax[f] > 0. ? dmax[f] / smax[f] : dmax[f]);
}

void make_config(THashTable *P, HashTable *BG)
{
  psph_config_default_3d(&g_cfg);
  psph_config &c = g_cfg;
  c.use_gsph = USE_GSPH; c.have_turbulence_lans = HAVE_TURBULENCE_LANS; c.have_energy_smooth = HAVE_ENERGY_SMOOTH;
  c.pghost_contribute_des = PGHOST_CONTRIBUTE_DES; c.fluid_compressibility = FLUID_COMPRESSIBILITY;
  c.adaptive_sml = ADAPTIVE_SML >= 1; c.sml_update_int = psphhook_sml_update_int();
  c.cutoff = CUTOFF; c.cutoff2 = CUTOFF2; c.msfrc_thresh = MSFRC_THRESH; c.epsilon_half = EPSILON_HALF;
  c.prandtl_num = PRANDTL_NUM; c.heat_trans_scale_ratio = HEAT_TRANS_SCALE_RATIO; c.e_smooth_ratio = E_SMOOTH_RATIO;
  c.tiny = TINY;
  c.alf = alf_P; c.beta = beta_P; c.ata = ata_P; c.g = ::g_P;  /* (ref_hooks.cc has a g_P of its own in the unnamed namespace) */ c.cfl = CFL_P; c.eta_smooth = eta_smooth_P;
  c.c_smooth = C_smooth_P; c.thresh = thresh_P; c.num_loop = num_loop_P;
  c.ng0 = ng0_P; c.Cvs = Cvs_P; c.Cvg = Cvg_P; c.Cva = Cva_P; c.Rg = Rg_P; c.Ra = Ra_P; c.rhoa0 = rhoa0_P; c.pa0 = pa0_P;
  c.erupt_top = Ll_P[2]; c.smlen_real = g_mp->smoothing_length;              /* update_pos.cc:40,141 (sic: not simprops->sml_of_phase2) */
  c.len_scale = g_mp->LENGTH_SCALE;
  c.del = PARTICLE_DENSITY * g_mp->smoothing_length * g_mp->LENGTH_SCALE;      /* preprocess.cc:99 */
  for (int d = 0; d < 3; d++) {
    c.mindom[d] = *(BG->get_minDom() + d) * c.len_scale;                       /* dataread.cc:318-323 */
    c.ncell[d] = (int) floor((*(BG->get_maxDom() + d) - *(BG->get_minDom() + d)) * c.len_scale / c.del + 0.5);
  }
  c.ptable_size = 4000000; c.device = 0; c.rank = 0; c.nranks = 1;
  (void) P;
}

void upload_cells(HashTable *BG)
{
  const psph_config &c = g_cfg;
  vector<psph_cell_attr> attr((size_t) c.ncell[0] * c.ncell[1] * c.ncell[2], psph_cell_attr());
  HTIterator it(BG);
  void *t;
  while ((t = it.next())) {
    BriefBucket *bb = (BriefBucket *) t;
    if (bb->check_brief()) continue;
    Bucket *b = (Bucket *) t;
    int idx[3];
    for (int d = 0; d < 3; d++) idx[d] = (int) floor((*(b->get_mincrd() + d) * c.len_scale - c.mindom[d]) / c.del + 0.5);
    psph_cell_attr &a = attr[((size_t) idx[0] * c.ncell[1] + idx[1]) * c.ncell[2] + idx[2]];
    a.type = b->get_bucket_type(); a.has_involved = b->get_has_involved(); a.active = b->is_active();
    for (int k = 0; k < 6; k++) { a.bucket_index[k] = b->get_bucket_index()[k]; a.bnd[k] = b->get_bnd()[k]; }
  }
  CK(psph_set_cells(G, attr.data(), attr.size()));
}

void upload_erupt(THashTable *P, HashTable *BG)
{
#ifndef SIMULATE_ASH
  psph_erupt er;
  memset(&er, 0, sizeof er);
  er.pos_v[0] = Pos_v_P[0]; er.pos_v[1] = Pos_v_P[1];
  er.rv = rv_P; er.Vv0 = Vv0_P; er.vel_prof = ERUPT_VELOCITY_PROF;
  er.bot = g_sp->get_bot(); er.sml = g_sp->sml_of_phase2; er.mass = g_sp->mass_of_phase2; er.ratio_sml_dx = RATIO_SML_DX;
  er.ev0 = ev0_P; er.rhov = rhov_P; er.pv0 = pv0_P; er.gamma_v = gamma_v_P;
#if FLUID_COMPRESSIBILITY == 0
  er.sndspd = 340.;
#else
  er.sndspd = 1482.;
#endif
  for (int d = 0; d < 3; d++) { er.kmin[d] = *(P->get_minDom() + d); er.kmax[d] = *(P->get_maxDom() + d); }
  vector<double> boxes;
  HTIterator it(BG);
  void *t;
  while ((t = it.next())) {
    BriefBucket *bb = (BriefBucket *) t;
    if (bb->check_brief()) continue;
    Bucket *b = (Bucket *) t;
    if (!b->is_erupt()) continue;
    for (int d = 0; d < 3; d++) boxes.push_back(*(b->get_mincrd() + d));
    for (int d = 0; d < 3; d++) boxes.push_back(*(b->get_maxcrd() + d));
  }
  CK(psph_set_erupt(G, &er, boxes.data(), boxes.size() / 6));
#endif
}

void upload_images(HashTable *BG, list<BndImage> &Image_table)
{
  const psph_config &c = g_cfg;
  vector<psph_image> img;
  for (list<BndImage>::iterator it = Image_table.begin(); it != Image_table.end(); ++it) {
    Bucket *b = (Bucket *) BG->lookup(it->bucket_key);
    if (!b) continue;
    psph_image m;
    memset(&m, 0, sizeof m);
    for (int k = 0; k < TKEYLENGTH; k++) m.ghost_key[k] = it->ghost_key[k];
    int idx[3];
    for (int d = 0; d < 3; d++) {
      idx[d] = (int) floor((*(b->get_mincrd() + d) * c.len_scale - c.mindom[d]) / c.del + 0.5);
      m.coord[d] = it->coord[d];
    }
    m.cell = (idx[0] * c.ncell[1] + idx[1]) * c.ncell[2] + idx[2];
    img.push_back(m);
  }
  CK(psph_set_images(G, img.data(), img.size()));
}

/* host table -> device (context created at the first call) */
void sync_up(THashTable *P, HashTable *BG)
{
  if (!G) {
    make_config(P, BG);
    if (psph_create(&g_cfg, &G) != 0) { fprintf(stderr, "ref_b200: psph_create failed (no CUDA device?)\n"); exit(2); }
  }
  if (!g_dirty) return;
  upload_cells(BG);
  g_rows.clear();
  THTIterator itr(P);
  Particle *p;
  while ((p = (Particle *) itr.next())) { psph_particle r; to_row(p, &r); g_rows.push_back(r); }
  CK(psph_upload(G, g_rows.data(), g_rows.size()));
  upload_erupt(P, BG);
  g_dirty = false;
}
}

/* ---- set-up calls: run as they are, remember the property records ---- */
void psphb200_add_air(THashTable *P, HashTable *BG, MatProps *mp, SimProps *sp, int np, int myid)
{
  g_mp = mp; g_sp = sp;
  add_air(P, BG, mp, sp, np, myid);
}
/* domain growth changes the host table: upload again before the next device call */
void psphb200_adapt_domain(THashTable *P, HashTable *BG, MatProps *mp, int np, int myid)
{
  adapt_domain(P, BG, mp, np, myid);
  g_dirty = true;
}

/* ---- the time loop ---- */
double psphb200_timestep(HashTable *BG, THashTable *P, TimeProps *tp)
{
  init_once();
  begin_step(P, 0);
  g_tp = tp; g_in_loop = true;
  if (getenv("PSPH_LOCKSTEP")) { g_lock = true; g_dirty = true; }    /* the device restarts every step from the host's table */
  sync_up(P, BG);
  double dt = 0.;
  CK(psph_timestep(G, &dt));
  if (g_lock) {
    const double dth = timestep(BG, P, tp);
    if (dth != dt) fprintf(stderr, "LOCKSTEP %d dt device %.17g host %.17g\n", g_step + 1, dt, dth);
    dt = dth;
  }
  g_step++;
  return dt;
}

int psphb200_search_neighs_consth(int myid, THashTable *P, HashTable *BG)
{
  if (!g_in_loop) return search_neighs_consth(myid, P, BG);          /* call during initialisation */
  sync_up(P, BG);                                                     /* the adapt block may have changed the table */
  CK(psph_set_dt(G, g_tp->dtime, g_tp->step, g_tp->iferupt()));       /* incrtime has run (main.cc:231) */
  CK(psph_search_neighs(G));
  if (g_lock) return search_neighs_consth(myid, P, BG);
  return 0;
}
/* with PSPH_DUMP_DIR set and the step inside PSPH_DUMP_STEPS, the device state comes back after every stage and is dumped
 * under the same names particler_hooked uses (profiles/dropin_diff.py finds the first stage where the two programs part) */
void stage_dump(THashTable *P, const char *stage)
{
  if (!dumping() || g_lock) return;
  download_to_host(P, true);
  dump_table(P, stage, "post");
}
void psphb200_adaptive_sml(int myid, THashTable *P)
{
  CK(psph_adaptive_sml(G));
  if (g_lock) adaptive_sml(myid, P);
  stage_dump(P, "sml");
}
int psphb200_mom_engr_update(int myid, THashTable *P, TimeProps *tp, SimProps *sp)
{
  if (dumping() && !g_lock) { download_to_host(P, true); dump_table(P, "mom", "pre"); }
  int r = psph_mom_engr_update(G);
  if (r < 0) die("psph_mom_engr_update");
  if (g_lock) {
    const int rh = mom_engr_update(myid, P, tp, sp);
    if (rh != r) fprintf(stderr, "LOCKSTEP %d mom_engr_update device %d host %d\n", g_step, r, rh);
    return rh;
  }
  stage_dump(P, "mom");
  return r;
}
void psphb200_smooth_density(THashTable *P) { CK(psph_smooth_density(G)); if (g_lock) smooth_density(P); stage_dump(P, "dens"); }
void psphb200_smooth_velocity(THashTable *P) { CK(psph_smooth_velocity(G)); if (g_lock) smooth_velocity(P); stage_dump(P, "svel"); }
void psphb200_update_pos(THashTable *P, TimeProps *tp, MatProps *mp)
{
  CK(psph_update_pos(G));
  if (g_lock) update_pos(P, tp, mp);
  stage_dump(P, "pos");
}

void psphb200_update_particles_in_bucket(int myid, THashTable *P, HashTable *BG, int *lost)
{
  if (!g_lock) download_to_host(P, false);         /* positions, bc_type, smlen, involved of this step */
  update_particles_in_bucket(myid, P, BG, lost);   /* the reference's own bucket lists follow */
  int dlost = 0;
  CK(psph_rebin(G, &dlost));
  if (dlost != *lost) fprintf(stderr, "ref_b200: step %d: device lost %d particles, host %d\n", g_step, dlost, *lost);
}

#ifndef SIMULATE_ASH
void psphb200_add_new_erupt(int myid, THashTable *P, HashTable *BG, TimeProps *tp, MatProps *mp, SimProps *sp, double dt)
{
  size_t before = 0, after = 0;
  { THTIterator it(P); while (it.next()) before++; }
  add_new_erupt(myid, P, BG, tp, mp, sp, dt);
  { THTIterator it(P); while (it.next()) after++; }
  int added = 0;
  CK(psph_add_new_erupt(G, tp->time, dt, (uint32_t) tp->step, &added));
  if ((size_t) added != after - before)
    fprintf(stderr, "ref_b200: step %d: device added %d erupt particles, host %zu\n", g_step, added, after - before);
}
#endif

int psphb200_apply_bcond(int myid, THashTable *P, HashTable *BG, MatProps *mp, list<BndImage> &img)
{
  if (!g_in_loop || g_dirty) return apply_bcond(myid, P, BG, mp, img);   /* set-up and adapt block: host tables */
  upload_images(BG, img);
  int r = psph_apply_bcond(G);
  if (r < 0) die("psph_apply_bcond");
  if (g_lock) {
    const int rh = apply_bcond(myid, P, BG, mp, img);
    lockstep_compare(P);
    return rh;
  }
  download_to_host(P, true);                       /* end of the step: the host particles are current again */
  dump_table(P, "bcond", "post");
  return r;
}

#define add_air psphb200_add_air
#define adapt_domain psphb200_adapt_domain
#define timestep psphb200_timestep
#define search_neighs_consth psphb200_search_neighs_consth
#define adaptive_sml psphb200_adaptive_sml
#define mom_engr_update psphb200_mom_engr_update
#define smooth_density psphb200_smooth_density
#define smooth_velocity psphb200_smooth_velocity
#define update_pos psphb200_update_pos
#define update_particles_in_bucket psphb200_update_particles_in_bucket
#define add_new_erupt psphb200_add_new_erupt
#define apply_bcond psphb200_apply_bcond
#endif /* CODE_DIMENSION == 3 */
#include PSPH_REF_MAIN_CC
