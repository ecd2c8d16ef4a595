// psph_grow.cu — dynamic domain growth on the device.
//
// Replaces  scan_outside_layer   src/plume/scan_outside_layer.cc:44-121   (main.cc:435, every step)
//           adapt_domain         src/plume/adapt_domain.cc:41-194         (main.cc:265, when adapt != 0)
//           add_pressure_ghost   src/plume/add_pressure_ghost.cc:36-316   (main.cc:271) with initial_air / air_prop
//                                (sph_lib.cc:831-872, :737-800), Meteo::interpolate (src/datstr/meteo.cc:64-147) and the
//                                initial-air Particle constructor (src/datstr/particle.cc:116-190).
// The reference walks its bucket hash table and, per bucket, the bucket's particle list; here the bucket attributes are one
// array over the whole lattice (psph_cell_attr, brief buckets = type 0) and the per-bucket facts the three calls need —
// "holds a row", "holds an involved row" — are two integer arrays filled by one pass over the particle rows (I_CELL is
// current after every re-bin and append), summed over the slabs when there are several.  Every rank then takes the same
// decisions for every bucket from identical inputs, so the attribute table stays replicated without being exchanged.
// None of the three calls moves a particle off the device; the reference's host loop downloads nothing for them.
#include "psph_internal.cuh"
#include <algorithm>

#define GRID(n) (unsigned) (((n) + 255) / 256), 256, 0, ctx->stream

// ---- per-bucket facts ---------------------------------------------------------------------
// cnt[cell] = live rows of the bucket owned by this rank; inv[cell] = 1 when one of them is involved (involved == 2)
__global__ void k_grow_count(Tab V, StepState *ss, int32_t *__restrict__ cnt, int32_t *__restrict__ inv)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ss->n) return;
  const int cell = V.i[I_CELL][p];
  if (cell < 0 || V.i[I_GUEST][p]) return;                           // lost in the last re-bin / a guest copy
  atomicAdd(&cnt[cell], 1);
  if (V.i[I_INVOLVED][p] == 2) inv[cell] = 1;
}

// neighbour q (0..26, 13 = the bucket itself) of cell: its linear index, or -1 when it lies outside the mesh
// (neigh_proc == -1) or is a brief bucket (check_brief())
__device__ __forceinline__ int grow_neigh(const GridView &G, const psph_cell_attr *attr, int cell, int q)
{
  if (q == 13) return -1;
  const int ck = cell % G.ncell[2], cj = (cell / G.ncell[2]) % G.ncell[1], ci = cell / (G.ncell[2] * G.ncell[1]);
  const int ni = ci + q / 9 - 1, nj = cj + (q / 3) % 3 - 1, nk = ck + q % 3 - 1;
  if (ni < 0 || nj < 0 || nk < 0 || ni >= G.ncell[0] || nj >= G.ncell[1] || nk >= G.ncell[2]) return -1;
  const int nl = (ni * G.ncell[1] + nj) * G.ncell[2] + nk;
  return attr[nl].type != PSPH_CELL_BRIEF ? nl : -1;
}

// ---- scan_outside_layer (scan_outside_layer.cc:44-121) ---------------------------------------
// A has_involved == 1 bucket that holds an involved row becomes HAS_BOTH (3).  Its neighbours are read with the flags they
// had BEFORE the call: the reference updates in place while it iterates, but the only value it writes (3) and the value the
// neighbour test looks for (0) never meet.  Two passes all the same, so that no thread reads what another writes.
__global__ void k_scan_outside(GridView G, psph_cell_attr *attr, const int32_t *__restrict__ cnt, const int32_t *__restrict__ inv,
                               int32_t *__restrict__ mark, int *adapt)
{
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= G.ncells) return;
  const psph_cell_attr &a = attr[cell];
  mark[cell] = 0;
  if (a.type == PSPH_CELL_BRIEF || a.has_involved != 1 || !inv[cell]) return;
  mark[cell] = 1;
  for (int q = 0; q < 27; q++) {
    const int nl = grow_neigh(G, attr, cell, q);
    if (nl < 0) continue;
    if (cnt[nl] > 0 && !attr[nl].has_involved && attr[nl].type != PSPH_CELL_UNDERGROUND) { atomicAdd(adapt, 1); break; }
  }
}
__global__ void k_scan_commit(long long ncells, psph_cell_attr *attr, const int32_t *__restrict__ mark)
{
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell < ncells && mark[cell]) attr[cell].has_involved = 3;
}

// ---- adapt_domain (adapt_domain.cc:41-194) ------------------------------------------------------
// mode[cell]: 0 untouched; 1 on-ground MIXED (its pressure ghosts turn real); 2 MIXED on a side or on top (the rows inside
// the domain turn real); 3 OVERGROUND (every row turns real)
__global__ void k_adapt_mark(GridView G, const psph_cell_attr *attr, const int32_t *__restrict__ cnt, int32_t *__restrict__ mode)
{
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= G.ncells) return;
  const psph_cell_attr &a = attr[cell];
  mode[cell] = 0;
  if (a.type == PSPH_CELL_BRIEF || a.has_involved || cnt[cell] == 0 || a.type == PSPH_CELL_UNDERGROUND || a.type == PSPH_CELL_PRESS_BC)
    return;
  bool hit = false;
  for (int q = 0; q < 27 && !hit; q++) {
    const int nl = grow_neigh(G, attr, cell, q);
    hit = nl >= 0 && (attr[nl].has_involved & 2);                     // is_has_involved(): untouched by this call
  }
  if (hit) mode[cell] = a.type == PSPH_CELL_MIXED ? (a.bucket_index[4] == -1 ? 1 : 2) : 3;
}
__global__ void k_adapt_cells(long long ncells, psph_cell_attr *attr, const int32_t *__restrict__ mode)
{
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ncells || !mode[cell]) return;
  psph_cell_attr &a = attr[cell];
  a.has_involved = 1;                                                 // set_has_involved(false); set_has_potential_involved(true)
  a.flags |= PSPH_CELLF_REAL;
  if (mode[cell] != 2) a.flags &= ~PSPH_CELLF_PRESS;
}
__device__ __forceinline__ bool grow_escape(const KParams &K, const psph_cell_attr &a, const double *pos)
{ // Bucket::determine_escape, bucket.cc:434-452
  for (int i = 0; i < K.dim; i++) {
    if (a.bucket_index[i * 2] == -1) {
      if ((pos[i] - a.bnd[i * 2]) < -K.tiny) return true;
    } else if (a.bucket_index[i * 2 + 1] == 1) {
      if ((pos[i] - a.bnd[i * 2 + 1]) > K.tiny) return true;
    }
  }
  return false;
}
__global__ void k_adapt_rows(Tab V, KParams K, StepState *ss, const psph_cell_attr *attr, const int32_t *__restrict__ mode, int *turned)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ss->n) return;
  const int cell = V.i[I_CELL][p];
  if (cell < 0) return;
  const int m = mode[cell];
  if (!m) return;
  bool turn = true;
  if (m == 1) turn = V.i[I_BC][p] == 1;                               // :113-118
  else if (m == 2) {                                                  // :133-141
    const double pos[3] = { V.d[D_X][p], V.d[D_Y][p], V.d[D_Z][p] };
    turn = !grow_escape(K, attr[cell], pos);
  }
  if (turn) {
    V.i[I_BC][p] = 100; V.i[I_INVOLVED][p] = 1;
    if (!V.i[I_GUEST][p]) atomicAdd(turned, 1);
  }
}

// ---- add_pressure_ghost (add_pressure_ghost.cc:36-316) -------------------------------------------
struct GrowPar
{
  psph_growth g;
  const double *mz, *mrho, *mp;    // the sounding: height [km], density, pressure [hPa]
  const double *mincrd;            // [ncells][3] Bucket::mincrd as the driver stores it, or nullptr: the lattice formula
  uint32_t add_step;
  int base, cap, rank, has_grad, nper;
  int bc;                          // bc_type of the new rows: 1 pressure ghost, 2 wall ghost
};

// fill[cell] = lattice points the call puts into the bucket (0: none); the bucket is marked active either way when a
// neighbour is involved (:108)
// layers of a bucket's lattice whose z lies above (mode 1) / below (mode 2) Ll[2]; every layer for mode 0
__device__ __forceinline__ int grow_layers(const GridView &G, const GrowPar &P, int cell, int mode)
{
  if (mode == 0) return P.nper;
  const int ck = cell % G.ncell[2];
  const double z0 = P.mincrd ? P.mincrd[3 * cell + 2] : (G.mindom[2] + ck * G.del) / G.len_scale;
  int n = 0;
  for (int k = 0; k < P.nper; k++) {
    const double z = __dadd_rn(__dadd_rn(z0, __dmul_rn(0.5, P.g.smlen)), __dmul_rn((double) k, P.g.smlen));
    n += mode == 1 ? z > P.g.Ll[2] : z < P.g.Ll[2];
  }
  return n;
}

__global__ void k_pghost_mark(GridView G, psph_cell_attr *attr, const int32_t *__restrict__ owner, GrowPar P, int32_t *__restrict__ fill,
                              int32_t *__restrict__ zmode)
{
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= G.ncells) return;
  psph_cell_attr &a = attr[cell];
  fill[cell] = 0;
  if (a.type == PSPH_CELL_BRIEF || a.has_involved || (a.flags & PSPH_CELLF_PRESS) || PSPH_CELLF_ERUPT_FLAG(a.flags) != 0) return;
  bool hit = false;
  for (int q = 0; q < 27 && !hit; q++) {
    const int nl = grow_neigh(G, attr, cell, q);
    hit = nl >= 0 && attr[nl].has_involved != 0;                      // (has_involved is not written by this call)
  }
  if (!hit) return;
  a.active = 1;
  if (!(a.type == PSPH_CELL_MIXED || a.type == PSPH_CELL_OVERGROUND || a.type == PSPH_CELL_PRESS_BC)) return;
  const int mode = (a.type == PSPH_CELL_MIXED && a.bucket_index[4] == -1) ? 1 : 0;   // on the ground: the layers above Ll[2] only (:139)
  zmode[cell] = mode;
  const int n = P.nper * P.nper * grow_layers(G, P, cell, mode);
  if (n > 0) a.flags |= PSPH_CELLF_PRESS;                             // add_pressure_ghost_particle (bucket.h:664-669)
  if (owner && owner[cell] != P.rank) return;                         // slabs: every rank fills the buckets it owns (!is_guest())
  fill[cell] = n;
}

// ---- add_wall_ghost (add_wall_ghost.cc:44-293) ---------------------------------------------------
// mark[cell] = 1: a non-empty on-ground MIXED bucket without wall ghosts, outside the vent (:136)
__global__ void k_wghost_pick(GridView G, const psph_cell_attr *attr, const int32_t *__restrict__ cnt, int32_t *__restrict__ mark)
{
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= G.ncells) return;
  const psph_cell_attr &a = attr[cell];
  mark[cell] = a.type == PSPH_CELL_MIXED && a.bucket_index[4] == -1 && cnt[cell] > 0 && !(a.flags & PSPH_CELLF_WALL) &&
               PSPH_CELLF_ERUPT_FLAG(a.flags) == 0;
}
// a picked bucket gets the lattice layers below Ll[2]; the bucket UNDER a picked bucket is marked active and gets the full
// lattice (:219-283).  bad: a picked bucket with no bucket below it (the reference asserts)
__global__ void k_wghost_mark(GridView G, psph_cell_attr *attr, const int32_t *__restrict__ owner, GrowPar P, const int32_t *__restrict__ mark,
                              int32_t *__restrict__ fill, int32_t *__restrict__ zmode, int *bad)
{
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= G.ncells) return;
  psph_cell_attr &a = attr[cell];
  const int ck = cell % G.ncell[2];
  const bool self = mark[cell] != 0, under = ck + 1 < G.ncell[2] && mark[cell + 1] != 0;   // (z-minor: cell + 1 is the bucket above)
  fill[cell] = 0; zmode[cell] = 0;
  if (self && (ck == 0 || attr[cell - 1].type == PSPH_CELL_BRIEF)) atomicAdd(bad, 1);
  if (!self && !under) return;
  if (under && a.type == PSPH_CELL_BRIEF) return;                     // (counted as bad by the bucket above)
  const int mode = under ? 0 : 2;                                     // (a bucket that is both would be filled twice: the reference exits)
  const int n = P.nper * P.nper * grow_layers(G, P, cell, mode);
  if (under) a.active = 1;
  if (n > 0) { a.flags |= PSPH_CELLF_WALL; a.active = 1; }            // add_wall_ghost_particle (bucket.h:656-661)
  const int own_cell = under ? cell + 1 : cell;                       // filled by the rank that owns the on-ground bucket (:136, :219)
  if (owner && owner[own_cell] != P.rank) return;
  zmode[cell] = mode;
  fill[cell] = n;
}

// Meteo::interpolate (meteo.cc:64-147): the sample itself, the end values outside the range, else linear between the
// neighbours
__device__ double dev_meteo(const double *__restrict__ z, const double *__restrict__ v, int n, double z0)
{
  if (z0 < z[0]) return v[0];
  if (z0 > z[n - 1]) return v[n - 1];
  int lo = 0, hi = n - 1;                                             // first index with z >= z0
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (z[mid] < z0) lo = mid + 1; else hi = mid; }
  if (z[lo] == z0) return v[lo];
  const int li = lo - 1, ui = lo;
  return __dadd_rn(v[li], __dmul_rn(__ddiv_rn(__dsub_rn(z0, z[li]), __dsub_rn(z[ui], z[li])), __dsub_rn(v[ui], v[li])));
}

// one thread per lattice point of a bucket that is filled; off[cell] = first new row of the bucket (exclusive scan of fill)
__global__ void k_pghost_fill(Tab V, GridView G, KParams K, GrowPar P, const int32_t *__restrict__ zmode,
                              const int32_t *__restrict__ off, const int32_t *__restrict__ cells, int ncell_fill)
{
  const int per = P.nper * P.nper * P.nper;
  const long long t = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if (t >= (long long) ncell_fill * per) return;
  const int cell = cells[t / per], l = (int) (t % per);
  const int i = l / (P.nper * P.nper), j = (l / P.nper) % P.nper, k = l % P.nper;
  const psph_growth &g = P.g;
  const int ck = cell % G.ncell[2], cj = (cell / G.ncell[2]) % G.ncell[1], ci = cell / (G.ncell[2] * G.ncell[1]);
  const int idx[3] = { ci, cj, ck }, ijk[3] = { i, j, k };
  const double dx = g.smlen, dx2 = __dmul_rn(0.5, dx);
  double crd[3];
  for (int d = 0; d < 3; d++) {
    const double mn = P.mincrd ? P.mincrd[3 * cell + d] : (G.mindom[d] + idx[d] * G.del) / G.len_scale;
    crd[d] = __dadd_rn(__dadd_rn(mn, dx2), __dmul_rn((double) ijk[d], dx));
  }
  int slot = l;
  const int mode = zmode[cell];
  if (mode) {                                                         // an on-ground bucket: the layers above (pressure ghosts) or
    if (!(mode == 1 ? crd[2] > g.Ll[2] : crd[2] < g.Ll[2])) return;   // below (wall ghosts) Ll[2], densely numbered
    const double z00 = __dadd_rn(P.mincrd ? P.mincrd[3 * cell + 2] : (G.mindom[2] + ck * G.del) / G.len_scale, dx2);
    int skipped = 0, nz = 0;
    for (int kk = 0; kk < P.nper; kk++) {
      const double z = __dadd_rn(z00, __dmul_rn((double) kk, dx));
      const bool in = mode == 1 ? z > g.Ll[2] : z < g.Ll[2];
      nz += in;
      skipped += (!in && kk < k);
    }
    slot = (i * P.nper + j) * nz + (k - skipped);
  }
  const int p = P.base + off[cell] + slot;
  if (p >= P.cap) return;
  double nc[3];
  for (int d = 0; d < 3; d++) nc[d] = __ddiv_rn(__dsub_rn(crd[d], g.kmin[d]), __dsub_rn(g.kmax[d], g.kmin[d]));
  uint32_t key[2];
  dev_hsfc3d(nc, key);
  // initial_air (sph_lib.cc:831-872): density, pressure, energy and the particle's mass from the atmosphere
  const double h = __dmul_rn(g.ratio_sml_dx, g.smlen), sml2 = __dmul_rn(0.5, h);
  double range[6];
  for (int d = 0; d < 3; d++) { range[2 * d] = __dsub_rn(crd[d], sml2); range[2 * d + 1] = __dadd_rn(crd[d], sml2); }
  const double vol = __dmul_rn(__dmul_rn(__dsub_rn(range[1], range[0]), __dsub_rn(range[3], range[2])), __dsub_rn(range[5], range[4]));
  double prss, dens, erg, mss;
  if (g.atmosphere_type == 4) {                                       // air_prop_meteo_based, sph_lib.cc:737-766
    const double hk = __dmul_rn(crd[2], 0.001);
    prss = __dmul_rn(dev_meteo(P.mz, P.mp, g.nmeteo, hk), 100.0);
    dens = dev_meteo(P.mz, P.mrho, g.nmeteo, hk);
    erg = __ddiv_rn(prss, __dmul_rn(dens, __dsub_rn(g.gamma_air, 1.0)));
    const double d1 = dev_meteo(P.mz, P.mrho, g.nmeteo, __dmul_rn(range[4], 0.001));
    const double d2 = dev_meteo(P.mz, P.mrho, g.nmeteo, __dmul_rn(range[5], 0.001));
    mss = __dmul_rn(vol, __dadd_rn(__dadd_rn(__dmul_rn(0.1666667, d1), __dmul_rn(0.6666666, dens)), __dmul_rn(0.1666667, d2)));
  } else {                                                            // air_prop_uniform, sph_lib.cc:682-712 (ideal gas)
    prss = K.pa0;
    dens = __ddiv_rn(prss, __dmul_rn(K.Ra, g.Ta0));
    erg = __ddiv_rn(prss, __dmul_rn(dens, __dsub_rn(g.gamma_air, 1.0)));
    mss = __dmul_rn(vol, dens);
  }
  // update_second_var (particle.cc:528-604) with ks = 0, ideal-gas branch, explicit IEEE operations in the reference's order
  const double ks = 0.0, na = __dsub_rn(1.0, ks), ng = __dmul_rn(K.ng0, ks), ns = __dsub_rn(ks, ng);
  const double Cvm = __dadd_rn(__dadd_rn(__dmul_rn(ns, K.Cvs), __dmul_rn(ng, K.Cvg)), __dmul_rn(na, K.Cva));
  const double Rm = __dadd_rn(__dmul_rn(ng, K.Rg), __dmul_rn(na, K.Ra)), Rmg = __ddiv_rn(Rm, __dadd_rn(ng, na));
  const double gam = __dadd_rn(1.0, __ddiv_rn(Rm, Cvm));
  const double desmg = __dmul_rn(dens, __dadd_rn(na, ng));
  const double pe = __ddiv_rn(__dmul_rn(__dmul_rn(Rmg, desmg), erg), Cvm);
  V.d[D_X][p] = crd[0]; V.d[D_Y][p] = crd[1]; V.d[D_Z][p] = crd[2];
  V.d[D_VX][p] = 0.0; V.d[D_VY][p] = 0.0; V.d[D_VZ][p] = 0.0;
  V.d[D_SVX][p] = 0.0; V.d[D_SVY][p] = 0.0; V.d[D_SVZ][p] = 0.0;
  V.d[D_RHO][p] = dens; V.d[D_E][p] = erg; V.d[D_SE][p] = 0.0;
  V.d[D_MASS][p] = mss;
  V.d[D_H][p] = h; V.d[D_HLIST][p] = h; V.d[D_KS][p] = ks;
  V.d[D_P][p] = pe; V.d[D_GAMMA][p] = gam; V.d[D_C][p] = sqrt(__ddiv_rn(__dmul_rn(gam, pe), desmg));   // pow(x, 0.5) == sqrt(x)
  V.d[D_T][p] = __ddiv_rn(erg, Cvm); V.d[D_CP][p] = K.lans ? Cvm : 0.0;
  V.d[D_RPH0][p] = __dmul_rn(dens, na); V.d[D_RPH1][p] = __dmul_rn(dens, ks);
  if (P.has_grad)
    for (int q = 0; q < 15; q++) V.d[D_GRAD0 + q][p] = 0.0;
  V.i[I_BC][p] = P.bc; V.i[I_PHASE][p] = 1; V.i[I_INVOLVED][p] = 0; V.i[I_GUEST][p] = 0; V.i[I_MYPROC][p] = P.rank;
  V.i[I_KEY0][p] = (int32_t) key[0]; V.i[I_KEY1][p] = (int32_t) key[1]; V.i[I_KEY2][p] = (int32_t) P.add_step;
  V.i[I_CELL][p] = cell;
}

// =========================================================================================
// host side
// =========================================================================================
struct GrowState
{
  int32_t *d_cnt = nullptr, *d_inv = nullptr, *d_mark = nullptr, *d_off = nullptr, *d_cells = nullptr;
  int *d_counter = nullptr;
  int *h_counter = nullptr;
  psph_growth g;
  bool has_growth = false;
  double *d_mz = nullptr, *d_mrho = nullptr, *d_mp = nullptr, *d_mincrd = nullptr;
  std::vector<int32_t> h_fill;
};

static GrowState *grow_of(psph_ctx *ctx)
{
  if (!ctx->grow) {
    GrowState *S = new GrowState();
    const size_t nb = (size_t) ctx->ncells * sizeof(int32_t);
    if (cudaMalloc(&S->d_cnt, nb) != cudaSuccess || cudaMalloc(&S->d_inv, nb) != cudaSuccess || cudaMalloc(&S->d_mark, nb) != cudaSuccess ||
        cudaMalloc(&S->d_off, nb) != cudaSuccess || cudaMalloc(&S->d_cells, nb) != cudaSuccess ||
        cudaMalloc(&S->d_counter, sizeof(int)) != cudaSuccess || cudaMallocHost(&S->h_counter, sizeof(int)) != cudaSuccess) {
      ctx->err = "psph_grow: out of memory"; delete S; return nullptr;
    }
    ctx->grow = S;
  }
  return static_cast<GrowState *>(ctx->grow);
}

void psph_grow_free(psph_ctx *ctx)
{
  GrowState *S = static_cast<GrowState *>(ctx->grow);
  if (!S) return;
  cudaFree(S->d_cnt); cudaFree(S->d_inv); cudaFree(S->d_mark); cudaFree(S->d_off); cudaFree(S->d_cells); cudaFree(S->d_counter);
  cudaFreeHost(S->h_counter);
  cudaFree(S->d_mz); cudaFree(S->d_mrho); cudaFree(S->d_mp); cudaFree(S->d_mincrd);
  delete S;
  ctx->grow = nullptr;
}

// the two per-bucket arrays, summed / or-ed over the slabs
static int grow_facts(psph_ctx *ctx, GrowState *S)
{
  if (psph_sync_n(ctx)) return -1;
  const size_t nb = (size_t) ctx->ncells * sizeof(int32_t);
  PSPH_CUDA(cudaMemsetAsync(S->d_cnt, 0, nb, ctx->stream));
  PSPH_CUDA(cudaMemsetAsync(S->d_inv, 0, nb, ctx->stream));
  if (ctx->n > 0) { k_grow_count<<<GRID(ctx->n)>>>(ctx->cur.t, ctx->d_ss, S->d_cnt, S->d_inv); PSPH_LAUNCH_CHECK(); }
  if (psph_comm_allreduce_i32(ctx, S->d_cnt, (size_t) ctx->ncells, 0)) return -1;
  if (psph_comm_allreduce_i32(ctx, S->d_inv, (size_t) ctx->ncells, 1)) return -1;
  return 0;
}

static int grow_read_counter(psph_ctx *ctx, GrowState *S, int *out)
{
  PSPH_CUDA(cudaMemcpyAsync(S->h_counter, S->d_counter, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = *S->h_counter;
  return 0;
}

extern "C" int psph_get_cells(psph_ctx *ctx, psph_cell_attr *attr, size_t ncells)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  if ((long long) ncells != ctx->ncells) { ctx->err = "psph_get_cells: ncells does not match the mesh"; return -2; }
  if (!ctx->d_attr) { ctx->err = "psph_get_cells: psph_set_cells has not been called"; return -2; }
  PSPH_CUDA(cudaMemcpyAsync(attr, ctx->d_attr, ncells * sizeof(psph_cell_attr), cudaMemcpyDeviceToHost, ctx->stream));
  PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int psph_scan_outside_layer(psph_ctx *ctx, int *adapt)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  if (adapt) *adapt = 0;
  if (ctx->cfg.dim != 3) return 0;
  if (!ctx->d_attr) { ctx->err = "psph_scan_outside_layer: psph_set_cells has not been called"; return -2; }
  GrowState *S = grow_of(ctx);
  if (!S) return -1;
  StageTimer tm(ctx, ST_GROW);
  if (grow_facts(ctx, S)) return -1;
  PSPH_CUDA(cudaMemsetAsync(S->d_counter, 0, sizeof(int), ctx->stream));
  k_scan_outside<<<GRID(ctx->ncells)>>>(psph_grid_view(ctx), ctx->d_attr, S->d_cnt, S->d_inv, S->d_mark, S->d_counter);
  PSPH_LAUNCH_CHECK();
  k_scan_commit<<<GRID(ctx->ncells)>>>(ctx->ncells, ctx->d_attr, S->d_mark);
  PSPH_LAUNCH_CHECK();
  // every rank scanned every bucket from the summed facts: the count is already the global one (main.cc:226-258 reduces it)
  return adapt ? grow_read_counter(ctx, S, adapt) : 0;
}

extern "C" int psph_adapt_domain(psph_ctx *ctx, int *turned)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  if (turned) *turned = 0;
  if (ctx->cfg.dim != 3) return 0;
  if (!ctx->d_attr) { ctx->err = "psph_adapt_domain: psph_set_cells has not been called"; return -2; }
  GrowState *S = grow_of(ctx);
  if (!S) return -1;
  StageTimer tm(ctx, ST_GROW);
  if (grow_facts(ctx, S)) return -1;
  PSPH_CUDA(cudaMemsetAsync(S->d_counter, 0, sizeof(int), ctx->stream));
  k_adapt_mark<<<GRID(ctx->ncells)>>>(psph_grid_view(ctx), ctx->d_attr, S->d_cnt, S->d_mark);
  PSPH_LAUNCH_CHECK();
  if (ctx->n > 0) {                                                   // (before the attributes change: MIXED rows read bnd only)
    k_adapt_rows<<<GRID(ctx->n)>>>(ctx->cur.t, ctx->kp, ctx->d_ss, ctx->d_attr, S->d_mark, S->d_counter);
    PSPH_LAUNCH_CHECK();
  }
  k_adapt_cells<<<GRID(ctx->ncells)>>>(ctx->ncells, ctx->d_attr, S->d_mark);
  PSPH_LAUNCH_CHECK();
  ctx->list_valid = false;                                            // bc_type changed: row flags and lists are stale
  psph_comm_mark(ctx, DIRTY_LAYOUT);
  return turned ? grow_read_counter(ctx, S, turned) : 0;
}

static int grow_fill_host(psph_ctx *ctx, GrowState *S, GrowPar &P, int *added)
{
  // the buckets to fill and their first rows: the list is short (one shell of buckets) and the call rare — scanned on the host
  S->h_fill.resize((size_t) ctx->ncells);
  PSPH_CUDA(cudaMemcpyAsync(S->h_fill.data(), S->d_mark, (size_t) ctx->ncells * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  if (psph_sync_n(ctx)) return -1;
  PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<int32_t> cells, off((size_t) ctx->ncells, 0);
  long long total = 0;
  for (long long c = 0; c < ctx->ncells; c++)
    if (S->h_fill[c] > 0) { cells.push_back((int32_t) c); off[c] = (int32_t) total; total += S->h_fill[c]; }
  if (total == 0) return 0;
  if ((long long) ctx->n + total > 0x7fffffffLL) { ctx->err = "psph_grow: table would exceed 2^31 rows"; return -2; }
  const int base = ctx->n;
  if (psph_reserve(ctx, (size_t) base + (size_t) total)) return -1;
  PSPH_CUDA(cudaMemcpyAsync(S->d_off, off.data(), (size_t) ctx->ncells * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  PSPH_CUDA(cudaMemcpyAsync(S->d_cells, cells.data(), cells.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  P.base = base; P.cap = (int) std::min<size_t>(ctx->cap, 0x7fffffff);
  const long long nthreads = (long long) cells.size() * P.nper * P.nper * P.nper;
  k_pghost_fill<<<(unsigned) ((nthreads + 255) / 256), 256, 0, ctx->stream>>>(ctx->cur.t, psph_grid_view(ctx), ctx->kp, P, S->d_inv, S->d_off,
                                                                             S->d_cells, (int) cells.size());
  PSPH_LAUNCH_CHECK();
  PSPH_CUDA(cudaStreamSynchronize(ctx->stream));                      // (off / cells are host vectors of this call)
  ctx->n = base + (int) total;
  ctx->h_ss->n = ctx->n;
  PSPH_CUDA(cudaMemcpyAsync(&ctx->d_ss->n, &ctx->h_ss->n, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->n_dirty = false;
  ctx->h_unknown = true;
  ctx->list_valid = false;
  psph_comm_mark(ctx, DIRTY_LAYOUT);
  if (added) *added = (int) total;
  return 0;
}

extern "C" int psph_set_growth(psph_ctx *ctx, const psph_growth *g, const double *mz, const double *mrho, const double *mp,
                               const double *cell_mincrd)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  GrowState *S = grow_of(ctx);
  if (!S) return -1;
  PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(S->d_mz); cudaFree(S->d_mrho); cudaFree(S->d_mp); cudaFree(S->d_mincrd);
  S->d_mz = S->d_mrho = S->d_mp = S->d_mincrd = nullptr;
  S->has_growth = false;
  if (!g) return 0;
  if (!(g->smlen > 0)) { ctx->err = "psph_set_growth: smlen must be positive"; return -2; }
  if (g->atmosphere_type == 4) {
    if (g->nmeteo < 2 || !mz || !mrho || !mp) { ctx->err = "psph_set_growth: atmosphere_type 4 needs the sounding"; return -2; }
    const size_t nb = (size_t) g->nmeteo * sizeof(double);
    PSPH_CUDA(cudaMalloc(&S->d_mz, nb)); PSPH_CUDA(cudaMalloc(&S->d_mrho, nb)); PSPH_CUDA(cudaMalloc(&S->d_mp, nb));
    PSPH_CUDA(cudaMemcpy(S->d_mz, mz, nb, cudaMemcpyHostToDevice));
    PSPH_CUDA(cudaMemcpy(S->d_mrho, mrho, nb, cudaMemcpyHostToDevice));
    PSPH_CUDA(cudaMemcpy(S->d_mp, mp, nb, cudaMemcpyHostToDevice));
  } else if (g->atmosphere_type != 2) { ctx->err = "psph_set_growth: atmosphere_type must be 4 (Meteo sounding) or 2 (uniform)"; return -2; }
  if (cell_mincrd) {
    const size_t nb = (size_t) ctx->ncells * 3 * sizeof(double);
    PSPH_CUDA(cudaMalloc(&S->d_mincrd, nb));
    PSPH_CUDA(cudaMemcpy(S->d_mincrd, cell_mincrd, nb, cudaMemcpyHostToDevice));
  }
  S->g = *g;
  S->has_growth = true;
  return 0;
}

extern "C" int psph_add_pressure_ghost(psph_ctx *ctx, uint32_t add_step, int *added)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  if (added) *added = 0;
  if (ctx->cfg.dim != 3) return 0;
  if (!ctx->d_attr) { ctx->err = "psph_add_pressure_ghost: psph_set_cells has not been called"; return -2; }
  GrowState *S = grow_of(ctx);
  if (!S) return -1;
  if (!S->has_growth) { ctx->err = "psph_add_pressure_ghost: psph_set_growth has not been called"; return -2; }
  const int32_t *owner = nullptr;
  if (ctx->cfg.nranks > 1 && !(owner = psph_comm_owner_table(ctx))) { ctx->err = "psph_add_pressure_ghost: psph_comm_init was not called"; return -2; }
  StageTimer tm(ctx, ST_GROW);
  GrowPar P;
  P.g = S->g; P.mz = S->d_mz; P.mrho = S->d_mrho; P.mp = S->d_mp; P.mincrd = S->d_mincrd;
  P.add_step = add_step; P.rank = ctx->cfg.rank; P.has_grad = ctx->has_grad;
  P.nper = (int) llround(ctx->cfg.del / ctx->cfg.len_scale / S->g.smlen);   // round((maxcrd - mincrd) / dx), :148-150
  if (P.nper < 1 || P.nper > 64) { ctx->err = "psph_add_pressure_ghost: bucket edge / smlen out of range"; return -2; }
  P.base = 0; P.cap = 0; P.bc = 1;
  k_pghost_mark<<<GRID(ctx->ncells)>>>(psph_grid_view(ctx), ctx->d_attr, owner, P, S->d_mark, S->d_inv);   // (d_inv: z mode per bucket)
  PSPH_LAUNCH_CHECK();
  return grow_fill_host(ctx, S, P, added);
}


extern "C" int psph_add_wall_ghost(psph_ctx *ctx, uint32_t add_step, int *added)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  if (added) *added = 0;
  if (ctx->cfg.dim != 3) return 0;
  if (!ctx->d_attr) { ctx->err = "psph_add_wall_ghost: psph_set_cells has not been called"; return -2; }
  GrowState *S = grow_of(ctx);
  if (!S) return -1;
  if (!S->has_growth) { ctx->err = "psph_add_wall_ghost: psph_set_growth has not been called"; return -2; }
  const int32_t *owner = nullptr;
  if (ctx->cfg.nranks > 1 && !(owner = psph_comm_owner_table(ctx))) { ctx->err = "psph_add_wall_ghost: psph_comm_init was not called"; return -2; }
  StageTimer tm(ctx, ST_GROW);
  if (grow_facts(ctx, S)) return -1;                                  // rows per bucket (the pressure ghosts just added count)
  GrowPar P;
  P.g = S->g; P.mz = S->d_mz; P.mrho = S->d_mrho; P.mp = S->d_mp; P.mincrd = S->d_mincrd;
  P.add_step = add_step; P.rank = ctx->cfg.rank; P.has_grad = ctx->has_grad;
  P.nper = (int) llround(ctx->cfg.del / ctx->cfg.len_scale / S->g.smlen);
  if (P.nper < 1 || P.nper > 64) { ctx->err = "psph_add_wall_ghost: bucket edge / smlen out of range"; return -2; }
  P.base = 0; P.cap = 0; P.bc = 2;
  PSPH_CUDA(cudaMemsetAsync(S->d_counter, 0, sizeof(int), ctx->stream));
  k_wghost_pick<<<GRID(ctx->ncells)>>>(psph_grid_view(ctx), ctx->d_attr, S->d_cnt, S->d_off);          // (d_off: picked buckets)
  PSPH_LAUNCH_CHECK();
  k_wghost_mark<<<GRID(ctx->ncells)>>>(psph_grid_view(ctx), ctx->d_attr, owner, P, S->d_off, S->d_mark, S->d_inv, S->d_counter);
  PSPH_LAUNCH_CHECK();
  int bad = 0;
  if (grow_read_counter(ctx, S, &bad)) return -1;
  if (bad) { ctx->err = "psph_add_wall_ghost: an on-ground bucket has no bucket below it (the reference asserts, add_wall_ghost.cc:222)"; return -2; }
  return grow_fill_host(ctx, S, P, added);
}
