/* psph.h — C ABI of the B200-native Plume-SPH particle hot path.
 *
 * The reference (Plume-SPH, C++/MPI, CPU only) has no plugin/FFI interface: its seam is
 * the set of free functions its driver calls once per time step
 * (/root/reference/src/plume/particler.h:22-357, src/multiproc/multiproc.h:39-110, call
 * sites in src/plume/main.cc:217-486 for 3-D and :676-751 for 1-D).  Each entry point
 * below replaces one of those calls; the citation on each says which.  The reference
 * driver keeps ownership of input parsing, atmosphere/eruption set-up, domain growth and
 * output; the particle table lives on the GPU between calls (INTEGRATION.md shows the
 * binding a maintainer adds to main.cc).
 *
 * Conventions (mirroring the reference's): one host thread per context; a context is not
 * thread-safe; every call is asynchronous on the context's CUDA stream unless it returns
 * data to the host; return value <0 = API/CUDA/NCCL failure (fatal, the driver should
 * MPI_Abort as main.cc:247-249 does), 0 = ok, >0 = number of particles that violated an
 * invariant the reference assert()s on (e.g. rho <= 0, mom_engr_update.cc:148,228).
 *
 * No torch types, no C++ types: plain pointers and sizes only.
 */
#ifndef PSPH_H
#define PSPH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSPH_ABI_VERSION 1

/* bc_type values (src/datstr/particle.h:137-145) */
#define PSPH_BC_ERUPT 0
#define PSPH_BC_PRESSURE 1
#define PSPH_BC_WALL 2
#define PSPH_BC_NOSAVE_REAL 99
#define PSPH_BC_REAL 100

/* bucket types (src/plume/constant.h:120-124) */
#define PSPH_CELL_BRIEF 0
#define PSPH_CELL_UNDERGROUND 10
#define PSPH_CELL_MIXED 11
#define PSPH_CELL_OVERGROUND 12
#define PSPH_CELL_PRESS_BC 13

/* Compile-time choices of the reference (src/plume/options.h, constant.h, parameters.h),
 * turned into one run-time record.  psph_config_default_3d()/_1d() fill in the reference's
 * defaults; SURVEY.md Appendix E lists which option set each BASELINE config needs. */
typedef struct psph_config
{
  int abi_version;           /* PSPH_ABI_VERSION */
  int dim;                   /* CODE_DIMENSION: 1 or 3 (config.h) */

  /* options.h */
  int use_gsph;              /* USE_GSPH: 0 artificial-viscosity SPH, 1 Godunov SPH */
  int riemann_solver;        /* RIEMANN_SOLVER: 0 Roe, 1 HLLC */
  int have_turbulence_lans;  /* HAVE_TURBULENCE_LANS: 0 | 1 */
  int have_energy_smooth;    /* HAVE_ENERGY_SMOOTH: 0 | 1 */
  int pghost_contribute_des; /* PGHOST_CONTRIBUTE_DES: 0 | 1 | 2 (particle.h:570-580) */
  int fluid_compressibility; /* FLUID_COMPRESSIBILITY: 0 ideal-gas mixture, 1 Tait */
  int adaptive_sml;          /* ADAPTIVE_SML: 0 | 1 */
  int sml_update_int;        /* SML_UPDATE_INT (constant.h:189); 1-D driver calls every step */

  /* constant.h */
  double cutoff;             /* CUTOFF   = 3 */
  double cutoff2;            /* CUTOFF2  = 8 (3-D) / 12 (1-D) */
  double neigh_radius;       /* 4.25, neighs.cc:84 */
  double msfrc_thresh;       /* MSFRC_THRESH = 1e-6 */
  double epsilon_half;       /* EPSILON_HALF = 0.4 */
  double prandtl_num;        /* PRANDTL_NUM = 0.85 */
  double heat_trans_scale_ratio; /* HEAT_TRANS_SCALE_RATIO = 1 */
  double e_smooth_ratio;     /* E_SMOOTH_RATIO = 4 */
  double tiny;               /* TINY = 1e-8 (Bucket::determine_escape) */

  /* parameters.h */
  double alf, beta, ata;     /* artificial viscosity alf_P, beta_P, ata_P */
  double g;                  /* g_P (0 when ATMOSPHERE_TYPE==2) */
  double cfl;                /* CFL_P */
  double eta_smooth;         /* eta_smooth_P */
  double c_smooth;           /* C_smooth_P = 2 */
  double thresh;             /* thresh_P = 1e-5 */
  int num_loop;              /* num_loop_P */
  int gsph_reconstruct;      /* Riemann states of the GSPH variant (options.h:141-262, :544-663):
                              *   0  piecewise constant: GSPH_SPECIFIC_VOL_APP 0, MINI_DERIVATIVE_COND 4 (the 1-D defaults)
                              *   3  from the gradients of calc_gradients: GSPH_SPECIFIC_VOL_APP 3 (cubic interface position),
                              *      MINI_DERIVATIVE_COND 10, GSPH_MODIFIED_MONOTONICITY 1, C_SHOCK = 3 (the 3-D defaults) */
  double ng0, Cvs, Cvg, Cva, Rg, Ra, rhoa0, pa0; /* EOS constants (particle.cc:528-604) */
  double erupt_top;          /* Ll_P[2]: erupt ghost turns real at z >= this (update_pos.cc:137) */
  double smlen_real;         /* matprops->smoothing_length, given to a particle that turns real */

  /* background mesh (src/preproc/preprocess.cc:99-179, dataread.cc:380-387):
   * cell (i,j,k) spans [ (mindom[d]+i*del)/len_scale , (mindom[d]+(i+1)*del)/len_scale ).
   * 1-D: one cell holding everything (ncell = {1,1,1}); the reference keeps no mesh there. */
  double mindom[3];
  double del;
  double len_scale;
  int ncell[3];

  /* particle-table iteration order (src/datstr/thashtab.cc:204-222): hash table size and
   * the key range it was built with.  Only used by psph_download(order=REFERENCE). */
  int ptable_size;           /* 4 000 000 (dataread.cc:254) / 10 000 (:506) */

  /* placement */
  int device;                /* CUDA device ordinal */
  int rank, nranks;          /* key-range slab owner / number of slabs */
  int pad1;
} psph_config;

/* One particle, host side (array of structs).  Field-for-field the reference's
 * ParticlePack (src/multiproc/pack_data.h:31-77) plus the EOS outputs that the reference
 * recomputes on unpack (pack_data.cc:459) — carried here so an upload is loss-free. */
typedef struct psph_particle
{
  uint32_t key[3];           /* TKey: birth Hilbert key + add_step (thashtab.h:18-85) */
  int32_t bc_type, phase_num, involved, guest;
  int32_t myprocess;
  double mass, smlen, mass_frac;
  double pressure, gamma, sound_speed, temperature, specific_heat_p;
  double phase_density[2];
  double coords[3], vel[3], smoothed_v[3];
  double rho, engr, smoothed_e;
  double grad[15];           /* d_rho,d_u,d_v,d_w,d_p (GSPH only) */
} psph_particle;

/* Per-cell attributes the driver owns (Bucket, src/datstr/bucket.h:212-300). */
typedef struct psph_cell_attr
{
  int32_t type;              /* PSPH_CELL_* */
  int32_t has_involved;      /* Bucket::get_has_involved() != 0 */
  int32_t active;
  int32_t bucket_index[6];   /* bucket.cc:434-452 */
  int32_t flags;             /* bits 0-3: Bucket::particles_type (bit 0 real, 1 wall ghost, 2 pressure ghost, 3 erupt ghost
                              * particles, bucket.h:364-399); bits 8-9: erupt_flag (1 erupt, 2 influx, :341-348).  Read by
                              * the domain-growth calls (psph_scan_outside_layer ..) only */
  double bnd[6];
} psph_cell_attr;
#define PSPH_CELLF_REAL 1
#define PSPH_CELLF_WALL 2
#define PSPH_CELLF_PRESS 4
#define PSPH_CELLF_ERUPT 8
#define PSPH_CELLF_ERUPT_FLAG(flags) (((flags) >> 8) & 3)

/* One wall-ghost image (BndImage, src/datstr/bnd_image.h:20-75): the mirror point of a wall ghost across the wall.
 * The driver keeps the table (search_bnd_images, main.cc:188,292,458) and hands it over whenever it changes. */
typedef struct psph_image
{
  uint32_t ghost_key[3];     /* TKey of the wall ghost (bc_type 2) */
  int32_t cell;              /* linear index (x-major, z-minor) of the bucket that contains the image point */
  double coord[3];           /* the image (reflection) point */
} psph_image;

/* What add_new_erupt (src/plume/setup_erupt.cc:390-670) reads besides the particle table: the vent
 * constants of parameters.h, the SimProps members it uses and the key range of the particle table. */
typedef struct psph_erupt
{
  double pos_v[2];           /* Pos_v_P[0], Pos_v_P[1]: centre of the duct's particle lattice */
  double rv;                 /* rv_P: duct radius (columns with x^2 + y^2 <= rv^2 inject) */
  double Vv0;                /* Vv0_P: average velocity */
  int32_t vel_prof;          /* ERUPT_VELOCITY_PROF: 0 uniform, 1 parabolic (sph_lib.cc:1223-1243) */
  int32_t pad;
  double bot;                /* SimProps::get_bot(): z where a column's particles are born */
  double sml;                /* SimProps::sml_of_phase2: lattice spacing and smoothing length (x ratio_sml_dx) */
  double mass;               /* SimProps::mass_of_phase2 */
  double ratio_sml_dx;       /* RATIO_SML_DX (constant.h:90) */
  double ev0, rhov, pv0, gamma_v, sndspd;   /* state of the erupted material (Particle ctor, particle.cc:210-316) */
  double kmin[3], kmax[3];   /* THashTable::get_minDom() / get_maxDom(): the box the birth keys are normalised to */
} psph_erupt;

/* What add_pressure_ghost (src/plume/add_pressure_ghost.cc:36-316) reads besides the tables: the lattice
 * constants, the key box of the particle table and the atmosphere its new particles are initialised from (initial_air,
 * sph_lib.cc:831-872 -> air_prop :770-800). */
typedef struct psph_growth
{
  double mass;               /* MatProps::particle_mass (overwritten by initial_air) */
  double smlen;              /* MatProps::smoothing_length: lattice spacing dx */
  double ratio_sml_dx;       /* RATIO_SML_DX (constant.h:190) */
  double gamma_air;          /* gamma_P: e = p / (rho (gamma_P - 1)) */
  double sndspd, gmm;        /* constructor values (add_pressure_ghost.cc:68-74): 340, 1.4 */
  double Ta0;                /* Ta0_P (uniform atmosphere only) */
  double kmin[3], kmax[3];   /* THashTable::get_minDom() / get_maxDom(): the box the birth keys are normalised to */
  double Ll[3], Lu[3];       /* Ll_P, Lu_P: an on-ground MIXED bucket is filled above Ll[2] only (:139) */
  int32_t atmosphere_type;   /* ATMOSPHERE_TYPE: 4 Meteo sounding (meteo.cc:64-110, linear in height), 2 uniform */
  int32_t nmeteo;            /* rows of the sounding handed to psph_set_growth */
} psph_growth;

typedef struct psph_ctx psph_ctx;

enum { PSPH_ORDER_DEVICE = 0, PSPH_ORDER_REFERENCE = 1 };

/* column ids for psph_column() (device pointers, SoA) */
enum {
  PSPH_COL_X = 0, PSPH_COL_Y, PSPH_COL_Z, PSPH_COL_VX, PSPH_COL_VY, PSPH_COL_VZ,
  PSPH_COL_SVX, PSPH_COL_SVY, PSPH_COL_SVZ, PSPH_COL_RHO, PSPH_COL_E, PSPH_COL_MASS,
  PSPH_COL_H, PSPH_COL_KS, PSPH_COL_P, PSPH_COL_GAMMA, PSPH_COL_C, PSPH_COL_T, PSPH_COL_CP,
  PSPH_COL_COUNT
};

void psph_config_default_3d(psph_config *cfg);   /* options.h:452-837, parameters.h:881-1046 */
void psph_config_default_1d(psph_config *cfg);   /* options.h:18-409,  parameters.h:1228-1450 */

/* after Read_Grid (main.cc:104) */
int psph_create(const psph_config *cfg, psph_ctx **out);
void psph_destroy(psph_ctx *ctx);
const char *psph_last_error(const psph_ctx *ctx);

/* Bucket attributes: after Read_Grid and whenever the driver changes them
 * (adapt_domain.cc, mesh_update.cc).  attr == NULL: every cell OVERGROUND, involved, active. */
int psph_set_cells(psph_ctx *ctx, const psph_cell_attr *attr, size_t ncells);
/* ... and back: the domain-growth calls below change has_involved, active and flags on the device */
int psph_get_cells(psph_ctx *ctx, psph_cell_attr *attr, size_t ncells);

/* Domain growth on the device (main.cc:435 and the first half of the adapt block, :256-271).
 * psph_scan_outside_layer (scan_outside_layer.cc:44-121), every step after the re-bin: a has_involved == 1 bucket that
 *   holds an involved particle becomes has_involved = 3; *adapt = how many of them touch a non-empty, not-involved,
 *   non-underground bucket (summed over the slabs).  No particle leaves the device for it.
 * psph_adapt_domain (adapt_domain.cc:41-194), when adapt != 0: the not-involved, non-empty buckets next to a
 *   has_involved >= 2 bucket become potential-involved and their pressure ghosts real (bc_type 100, involved 1).
 * psph_set_growth + psph_add_pressure_ghost (add_pressure_ghost.cc:36-316): the exposed buckets are filled with a lattice
 *   of static pressure ghosts (birth key = Hilbert index in the key box + add_step, state from the atmosphere);
 *   *added = rows appended on this rank.  meteo_z [km], meteo_rho, meteo_p [hPa]: columns 0, 1, 2 of the sounding;
 *   cell_mincrd [ncells][3]: Bucket::mincrd of every bucket as the driver stores it — the lattice points are
 *   mincrd + dx / 2 + i dx (:161-163) — or NULL: (mindom + i del) / len_scale.
 * The driver keeps shift_brief_buck, update_bgmesh and the image search (they run on bucket attributes: psph_get_cells). */
int psph_scan_outside_layer(psph_ctx *ctx, int *adapt);
int psph_adapt_domain(psph_ctx *ctx, int *turned);
int psph_set_growth(psph_ctx *ctx, const psph_growth *g, const double *meteo_z, const double *meteo_rho, const double *meteo_p,
                    const double *cell_mincrd);
int psph_add_pressure_ghost(psph_ctx *ctx, uint32_t add_step, int *added);
/* add_wall_ghost (add_wall_ghost.cc:44-293, main.cc:277), after the driver's shift_brief_buck (psph_set_cells with the new
 * buckets): every non-empty on-ground MIXED bucket without wall ghosts, outside the vent, gets the lattice layers below
 * Ll[2] as wall ghosts (bc_type 2, state of initial_air), and the bucket under it — marked active — a full lattice.
 * -2 when such a bucket has no bucket below it (the reference asserts).  The image search for the new ghosts stays with the
 * driver (search_bnd_images, main.cc:292; plume_sph_b200.mesh.wall_ghost_images for a flat ground). */
int psph_add_wall_ghost(psph_ctx *ctx, uint32_t add_step, int *added);

/* after initialisation (main.cc:199) / after add_new_erupt (:417) / ghost regeneration (:271-283) */
int psph_upload(psph_ctx *ctx, const psph_particle *rows, size_t n);
int psph_append(psph_ctx *ctx, const psph_particle *rows, size_t n);
/* before write_output (main.cc:474); order REFERENCE = THashTable iteration order */
int psph_download(psph_ctx *ctx, psph_particle *rows, size_t cap, size_t *n, int order);
/* The same three calls for rows `row_bytes` apart: sizeof(psph_particle), or PSPH_ROW_BYTES_SPH — psph_particle cut off
 * before its gradient block.  An SPH build of the reference (USE_GSPH == 0) does not carry gradients in what it
 * exchanges either: ParticlePack holds d_rho..d_p only #if USE_GSPH (src/multiproc/pack_data.h:68-75); 208 instead of
 * 328 bytes per particle over PCIe. */
#define PSPH_ROW_BYTES_SPH offsetof(psph_particle, grad)
int psph_upload_rows(psph_ctx *ctx, const void *rows, size_t n, size_t row_bytes);
int psph_append_rows(psph_ctx *ctx, const void *rows, size_t n, size_t row_bytes);
int psph_download_rows(psph_ctx *ctx, void *rows, size_t cap, size_t *n, int order, size_t row_bytes);
/* The output path as columns (write_h5part, h5part_writer.cc:40-269): the datasets of one "Step#k" group for the rows
 * the writer emits — non-guest real particles (:88-91), or every particle when with_ghosts != 0 (a WRITE_GHOSTS build) —
 * in THashTable iteration order (thashtab.cc:204-222, :86-134), selected, ordered and gathered on the device:
 *   dcols[10] = x, y, z, Rho, Vx, Vy, Vz, engr, mssfrc, prss     icols[3] = phase, bctp, guest      (ID is 0 .. n-1)
 * Each array holds `cap` entries (a NULL entry is skipped); *n = rows written.  dcols == NULL: count only. */
int psph_download_h5part(psph_ctx *ctx, double *const *dcols, int32_t *const *icols, size_t cap, size_t *n, int with_ghosts);
size_t psph_num_particles(const psph_ctx *ctx);
/* rows the summation stages update on this rank (need_neigh, particle.h:555-558, and the list rule of
 * neighs.cc:65-82); builds the cell list if needed and synchronises */
int64_t psph_num_updated(psph_ctx *ctx);
/* device-resident SoA column (valid until the next call that reorders particles) */
int psph_column(psph_ctx *ctx, int col, double **dptr);
/* Upload directly from device SoA columns (bench / device-side generators): cols[PSPH_COL_*],
 * bc/phase int32 device arrays; keys are assigned from the row index. */
int psph_upload_device_soa(psph_ctx *ctx, const double *const *cols, const int32_t *bc,
                           const int32_t *phase, size_t n);

/* timestep(), main.cc:220 — returns the unclipped CFL dt; TimeProps::incrtime stays on host */
int psph_timestep(psph_ctx *ctx, double *dt);
/* after incrtime (main.cc:231): dt actually used, step counter, TimeProps::iferupt() */
int psph_set_dt(psph_ctx *ctx, double dt, int step, int erupting);

int psph_search_neighs(psph_ctx *ctx);       /* search_neighs_consth, main.cc:318 */
int psph_adaptive_sml(psph_ctx *ctx);        /* adaptive_sml,         main.cc:322 */
int psph_calc_gradients(psph_ctx *ctx);      /* calc_gradients,       main.cc:343 */
int psph_mom_engr_update(psph_ctx *ctx);     /* mom_engr_update,      main.cc:358 */
int psph_smooth_density(psph_ctx *ctx);      /* smooth_density,       main.cc:378 */
int psph_smooth_velocity(psph_ctx *ctx);     /* smooth_velocity,      main.cc:392 */
int psph_update_pos(psph_ctx *ctx);          /* update_pos,           main.cc:405 */
int psph_rebin(psph_ctx *ctx, int *lost);    /* update_particles_in_bucket, main.cc:408 */
/* Wall boundary condition.  psph_set_images replaces the table search_bnd_images fills (main.cc:188,292,458) and is
 * called whenever that table changes; psph_apply_bcond replaces apply_bcond (bcond.cc:24-249; main.cc:196,298,468):
 * every wall ghost with an image takes the fluid state interpolated at its image point, velocity reflected about the
 * wall normal, secondary variables recomputed.  Single rank (nranks == 1) in this version. */
/* Eruption boundary: psph_set_erupt hands over the vent record and the boxes [mincrd[3], maxcrd[3]] of the buckets
 * flagged is_erupt(); psph_add_new_erupt replaces add_new_erupt (setup_erupt.cc:390-670; main.cc:414-417): every duct
 * column injects floor((fmod(t - dt, t_each) + dt) / t_each) erupt ghosts (bc_type 0, phase 2) at the bottom, born with a
 * Hilbert key of their position and add_step as third key word.  time = TimeProps::time after incrtime.  *added gets
 * the number of new particles (may be NULL; reading it synchronises).  Single rank in this version. */
int psph_set_erupt(psph_ctx *ctx, const psph_erupt *erupt, const double *cell_boxes, size_t ncells);
int psph_add_new_erupt(psph_ctx *ctx, double time, double dt, uint32_t add_step, int *added);
int psph_set_images(psph_ctx *ctx, const psph_image *images, size_t n);
int psph_apply_bcond(psph_ctx *ctx);

/* The whole per-step hot path in one call, in the reference's order (main.cc:217-408):
 * timestep -> [clip on host via cb] -> search -> [adaptive_sml] -> [gradients] -> mom/energy
 * -> density -> [velocity filter] -> positions -> re-bin.  dt_max caps the CFL dt (output/end
 * time clipping of properties.h:71-82); *dt_used returns the dt applied. */
int psph_step(psph_ctx *ctx, double dt_max, int erupting, double *dt_used, int *lost);
/* Same, nsteps times with no host round trip in between (dt stays on the device). */
int psph_run(psph_ctx *ctx, int nsteps, int erupting);

/* move_data (main.cc:326..471) / repartition (:440-464); no-ops for nranks == 1
 * (move_data.cc:53-54).  nccl_unique_id: 128 bytes from ncclGetUniqueId on rank 0, handed to every
 * rank by the driver (MPI_Bcast in the reference driver, torch.distributed in bench.py).
 * psph_exchange_ghosts sends what changed since the previous call: the columns a stage wrote, or
 * -- after update_pos / re-bin / upload / append -- migrates rows to their new owner and mirrors
 * the boundary cells anew. */
int psph_comm_unique_id(void *out, size_t cap);   /* rank 0: ncclGetUniqueId; returns the id size (128) */
int psph_comm_init(psph_ctx *ctx, const void *nccl_unique_id, size_t id_bytes);
int psph_exchange_ghosts(psph_ctx *ctx);
int psph_repartition(psph_ctx *ctx);
/* position of every background cell (linear index, x-major z-minor) along the bucket Hilbert curve
 * (preprocess.cc:196-205, hilbert.c:156-303): the order the slabs cut */
int psph_cell_ranks(psph_ctx *ctx, int32_t *rank_of_cell, size_t cap);
/* slab boundaries in cell-rank units (cuts[0..nranks]); returns nranks */
int psph_slab_cuts(psph_ctx *ctx, int *cuts, size_t cap);
int64_t psph_comm_bytes_sent(psph_ctx *ctx, int reset);

/* test-only: neighbour lists exactly as search_neighs_consth builds them (neighs.cc:37-211),
 * as CSR over device rows: offsets[n+1], idx = device row numbers, both in the row order that
 * psph_download(order = PSPH_ORDER_DEVICE) returns right afterwards (map through the keys).
 * Order inside one list is unspecified.  Two-call protocol: idx == NULL only fills offsets
 * and *nnz. */
int psph_emit_neighbours(psph_ctx *ctx, int64_t *offsets, int32_t *idx, size_t idx_cap, size_t *nnz);

/* Invariant violations the reference assert()s on (rho <= 0, empty normalisation sums),
 * counted on the device since the last reset; synchronises the stream. */
int psph_errors(psph_ctx *ctx, int reset);
/* diagnostics of the last step's pair path: out = { pair_mode (2 block tiles, 4 sub-tiled mixed-h lists, else legacy), rows
 * the summation stages update, rows among them that got no pair list (summed by the fused gather), rows whose
 * near-partner list overflowed (smooth_density sums those by itself) }; synchronises */
int psph_pair_stats(psph_ctx *ctx, int32_t out[4]);

/* diagnostic: linear background-cell index (x-major, z-minor) of every row, in the row order
 * psph_download(order = PSPH_ORDER_DEVICE) returns right afterwards; replaces walking the
 * Bucket plists (bucket.h:255-321) in the tests of update_particles_in_bucket. */
int psph_cell_of_rows(psph_ctx *ctx, int32_t *cells, size_t cap);

/* instrumentation: kernels launched since the last call (bench.py gpu_launches); with
 * psph_set_timing(on) every stage call is bracketed by CUDA events on the context stream
 * (and synchronised) and psph_stage_ms returns the accumulated device time of a stage:
 * 0 timestep, 1 search, 2 adaptive_sml, 3 gradients, 4 mom_engr, 5 density, 6 velocity,
 * 7 update_pos, 8 rebin. */
int64_t psph_launch_count(psph_ctx *ctx, int reset);
int psph_set_timing(psph_ctx *ctx, int on);
double psph_stage_ms(psph_ctx *ctx, int stage, int reset);
/* per-kernel device time (CUDA events on the context stream, only while psph_set_timing is on):
 * 0 gather mom_engr, 1 gather density, 2 gather velocity filter, 3 gather adaptive_sml,
 * 4 gather gradients, 5 gather GSPH mom_engr, 6 radix sort, 7 column reorder,
 * 8 pair search of the 3-D SPH stages (the scan that writes the per-row pair lists) */
double psph_kernel_ms(psph_ctx *ctx, int kernel, int reset, int64_t *calls);
/* pinned host memory for the rows passed to psph_upload / psph_append / psph_download
 * (optional: any host pointer works, pinned ones copy at full PCIe rate) */
void *psph_host_alloc(size_t bytes);
void psph_host_free(void *p);
int psph_sync(psph_ctx *ctx);
void *psph_stream(psph_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* PSPH_H */
