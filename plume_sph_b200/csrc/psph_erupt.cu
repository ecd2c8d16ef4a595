// psph_erupt.cu — eruption boundary on the device.
//
// Replaces add_new_erupt (src/plume/setup_erupt.cc:390-670): the duct is a lattice of columns, spacing sml_of_phase2,
// inside the circle x^2 + y^2 <= rv^2; per time step a column injects n = floor((fmod(t - dt, t_each) + dt) / t_each)
// erupt ghosts (bc_type 0, phase 2) at z = bot + (t_add + i t_each) vel, t_each = sml / vel(r) — a closed form, so the
// reference's four nested while-loops become ONE thread per column.  The thread rebuilds its column's coordinates by
// the reference's own repeated additions and evaluates every expression with explicit IEEE operations (__dmul_rn,
// __dadd_rn, ... : no FMA contraction), so positions, velocities and the birth keys (Hilbert index of the position +
// add_step, THSFC3d, hilbert.c:329-359) are bit-identical to the reference's.  New rows are appended to the table in
// column order (slots from a block-wide scan of the per-column counts: deterministic), each with the state the
// erupted-material constructor gives it (Particle ctor, particle.cc:210-316).
#include "psph_internal.cuh"

struct EruptPar
{
  psph_erupt e;
  int nxp, nxm, nyp, nym;        // columns of the x+ / x- halves and of the y+ / y- runs (setup_erupt.cc:456-600)
  double t_prev, dt;
  uint32_t add_step;
  int base, cap, rank, has_grad, nbox;
};

// One block: the columns are few (tens to thousands) and the slots must not depend on scheduling — a row's place in
// the table decides its place among the rows of its subcell after the (stable) sort, hence summation orders.  Columns
// are taken 256 at a time; a block-wide exclusive scan of the per-column counts gives every column its slots.
constexpr int ERUPT_THREADS = 256;
constexpr int ERUPT_MAX_PER_COL = 64;         // particles one column can inject per call that are tracked in the mask
__global__ void __launch_bounds__(ERUPT_THREADS) k_add_erupt(Tab V, GridView G, KParams K, StepState *ss, EruptPar P,
                                                             const double *__restrict__ boxes, const int32_t *__restrict__ owner)
{
  __shared__ int s_warp[ERUPT_THREADS / 32], s_base;
  const int ny = P.nyp + P.nym, ncol = (P.nxp + P.nxm) * ny;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const psph_erupt &e = P.e;
  const double sml = e.sml, sml2 = __dmul_rn(0.5, sml);
  // state of the erupted material (Particle ctor, particle.cc:210-316; update_second_var with ks = 1, then pressure,
  // sound speed and gamma overwritten)
  const double ks = 1.0, na = __dsub_rn(1.0, ks), ng = __dmul_rn(K.ng0, ks), ns = __dsub_rn(ks, ng);
  const double Cvm = __dadd_rn(__dadd_rn(__dmul_rn(ns, K.Cvs), __dmul_rn(ng, K.Cvg)), __dmul_rn(na, K.Cva));
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int c0 = 0; c0 < ncol; c0 += ERUPT_THREADS) {
    const int t = c0 + threadIdx.x;
    int n = 0;
    double x = 0., y = 0., vel = 0., t_each = 0., t_add = 0.;
    if (t < ncol) {
      const int ix = t / ny, iy = t - ix * ny;
      // the column's coordinates, built as the loops build them: start + repeated addition (:456,:526,:530,:598)
      if (ix < P.nxp) { x = __dadd_rn(e.pos_v[0], sml2); for (int k = 0; k < ix; k++) x = __dadd_rn(x, sml); }
      else { x = __dsub_rn(e.pos_v[0], sml2); for (int k = 0; k < ix - P.nxp; k++) x = __dsub_rn(x, sml); }
      if (iy < P.nyp) { y = __dadd_rn(e.pos_v[1], sml2); for (int k = 0; k < iy; k++) y = __dadd_rn(y, sml); }
      else { y = __dsub_rn(e.pos_v[1], sml2); for (int k = 0; k < iy - P.nyp; k++) y = __dsub_rn(y, sml); }
      const double dist = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));               // :461-464
      if (dist <= __dmul_rn(e.rv, e.rv)) {
        vel = e.Vv0;                                                                 // vel_prof, sph_lib.cc:1223-1243
        if (e.vel_prof != 0) vel = __dmul_rn(__dmul_rn(e.Vv0, 2.0), __dsub_rn(1.0, __ddiv_rn(dist, __dmul_rn(e.rv, e.rv))));
        t_each = __ddiv_rn(sml, vel);
        t_add = __dadd_rn(fmod(P.t_prev, t_each), P.dt);                             // :469
        n = (int) floor(__ddiv_rn(t_add, t_each));
        t_add = __dsub_rn(t_add, __dmul_rn((double) n, t_each));
        n = max(0, min(n, ERUPT_MAX_PER_COL));
      }
    }
    // which of the column's n particles lie in a bucket flagged is_erupt() (in_bucket, :609-650; both faces inclusive)
    unsigned long long keep = 0ull;
    for (int i = 0; i < n; i++) {
      const double z = __dadd_rn(e.bot, __dmul_rn(__dadd_rn(t_add, __dmul_rn((double) i, t_each)), vel));
      bool inside = false;
      for (int b = 0; b < P.nbox && !inside; b++) {
        const double *mn = boxes + 6 * b, *mx = mn + 3;
        inside = x <= mx[0] && x >= mn[0] && y <= mx[1] && y >= mn[1] && z <= mx[2] && z >= mn[2];
      }
      // key-range slabs: every rank runs every column (as every MPI rank of the reference does, :456-600) and keeps the
      // particles born in a bucket it owns; the others' copies arrive with the next ghost refresh
      if (inside && owner) {
        const double crd[3] = { x, y, z };
        int idx[3];
        for (int d = 0; d < 3; d++) idx[d] = dev_cell_index(G, d, crd[d]);
        inside = owner[(idx[0] * G.ncell[1] + idx[1]) * G.ncell[2] + idx[2]] == P.rank;
      }
      if (inside) keep |= 1ull << i;
    }
    const int cnt = __popcll(keep);
    int inc = cnt;                               // block-wide exclusive scan of cnt
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int pre = s_base;
    for (int k = 0; k < warp; k++) pre += s_warp[k];
    int slot = pre + inc - cnt;
    for (int i = 0; i < n; i++) {
      if (!((keep >> i) & 1ull)) continue;
      const int p = P.base + slot++;
      if (p >= P.cap) continue;               // (the host reserves an upper bound: cannot happen)
      const double crd[3] = { x, y, __dadd_rn(e.bot, __dmul_rn(__dadd_rn(t_add, __dmul_rn((double) i, t_each)), vel)) };
      double nc[3];
      for (int k = 0; k < 3; k++) nc[k] = __ddiv_rn(__dsub_rn(crd[k], e.kmin[k]), __dsub_rn(e.kmax[k], e.kmin[k]));
      uint32_t key[2];
      dev_hsfc3d(nc, key);
      V.d[D_X][p] = crd[0]; V.d[D_Y][p] = crd[1]; V.d[D_Z][p] = crd[2];
      V.d[D_VX][p] = 0.0; V.d[D_VY][p] = 0.0; V.d[D_VZ][p] = vel;
      V.d[D_SVX][p] = 0.0; V.d[D_SVY][p] = 0.0; V.d[D_SVZ][p] = 0.0;
      V.d[D_RHO][p] = e.rhov; V.d[D_E][p] = e.ev0; V.d[D_SE][p] = e.ev0;
      V.d[D_MASS][p] = e.mass;
      const double h = __dmul_rn(e.ratio_sml_dx, sml);
      V.d[D_H][p] = h; V.d[D_HLIST][p] = h; V.d[D_KS][p] = ks;
      V.d[D_P][p] = e.pv0; V.d[D_GAMMA][p] = e.gamma_v; V.d[D_C][p] = e.sndspd;
      V.d[D_T][p] = __ddiv_rn(e.ev0, Cvm); V.d[D_CP][p] = K.lans ? Cvm : 0.0;
      V.d[D_RPH0][p] = __dmul_rn(e.rhov, na); V.d[D_RPH1][p] = __dmul_rn(e.rhov, ks);
      if (P.has_grad)
        for (int k = 0; k < 15; k++) V.d[D_GRAD0 + k][p] = 0.0;
      V.i[I_BC][p] = 0; V.i[I_PHASE][p] = 2; V.i[I_INVOLVED][p] = 0; V.i[I_GUEST][p] = 0; V.i[I_MYPROC][p] = P.rank;
      V.i[I_KEY0][p] = (int32_t) key[0]; V.i[I_KEY1][p] = (int32_t) key[1]; V.i[I_KEY2][p] = (int32_t) P.add_step;
      int idx[3];
      for (int d = 0; d < 3; d++) idx[d] = dev_cell_index(G, d, crd[d]);
      V.i[I_CELL][p] = (idx[0] * G.ncell[1] + idx[1]) * G.ncell[2] + idx[2];
    }
    __syncthreads();                            // every warp has read s_base and s_warp
    if (threadIdx.x == ERUPT_THREADS - 1) s_base = pre + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) ss->n = min(P.base + s_base, P.cap);
}

extern "C" int psph_set_erupt(psph_ctx *ctx, const psph_erupt *er, const double *boxes, size_t nbox)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(ctx->d_erupt_boxes); ctx->d_erupt_boxes = nullptr; ctx->n_erupt_boxes = 0; ctx->has_erupt = false;
  if (!er) return 0;
  if (!(er->sml > 0) || !(er->rv >= 0) || !(er->Vv0 > 0)) { ctx->err = "psph_set_erupt: sml, rv, Vv0 must be positive"; return -2; }
  ctx->erupt = *er;
  if (nbox) {
    PSPH_CUDA(cudaMalloc(&ctx->d_erupt_boxes, nbox * 6 * sizeof(double)));
    PSPH_CUDA(cudaMemcpy(ctx->d_erupt_boxes, boxes, nbox * 6 * sizeof(double), cudaMemcpyHostToDevice));
  }
  ctx->n_erupt_boxes = (int) nbox;
  ctx->has_erupt = true;
  return 0;
}

extern "C" int psph_add_new_erupt(psph_ctx *ctx, double time, double dt, uint32_t add_step, int *added)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  if (added) *added = 0;
  if (ctx->cfg.dim != 3) return 0;                                   // the 1-D loop has no eruption (main.cc:676-751)
  if (!ctx->has_erupt) { ctx->err = "psph_add_new_erupt: psph_set_erupt has not been called"; return -2; }
  const int32_t *owner = nullptr;                                    // slabs: owner rank of every cell
  if (ctx->cfg.nranks > 1 && !(owner = psph_comm_owner_table(ctx))) { ctx->err = "psph_add_new_erupt: psph_comm_init was not called"; return -2; }
  const psph_erupt &e = ctx->erupt;
  EruptPar P;
  P.e = e;
  // the extent of the reference's while-loops (setup_erupt.cc:446-449,456-600): columns per half / per run
  const double sml2 = 0.5 * e.sml;
  P.nxp = P.nxm = P.nyp = P.nym = 0;
  for (double x = e.pos_v[0] + sml2; x <= e.rv; x += e.sml) P.nxp++;
  for (double x = e.pos_v[0] - sml2; x >= -e.rv; x -= e.sml) P.nxm++;
  for (double y = e.pos_v[1] + sml2; y <= e.rv; y += e.sml) P.nyp++;
  for (double y = e.pos_v[1] - sml2; y >= -e.rv; y -= e.sml) P.nym++;
  const long long ncol = (long long) (P.nxp + P.nxm) * (P.nyp + P.nym);
  if (ncol == 0 || ctx->n_erupt_boxes == 0) return 0;
  // upper bound of the particles one call adds: every column at the fastest injection rate
  const double vmax = e.vel_prof ? 2.0 * e.Vv0 : e.Vv0;
  const long long per_col = (long long) (dt * vmax / e.sml) + 2;
  if (per_col > ERUPT_MAX_PER_COL) { ctx->err = "psph_add_new_erupt: more than 64 particles per column and call"; return -2; }
  if (ncol * per_col > 0x3fffffffLL) { ctx->err = "psph_add_new_erupt: time step too long for the duct lattice"; return -2; }
  if (psph_sync_n(ctx)) return -1;
  const int base = ctx->n;
  if (psph_reserve(ctx, (size_t) base + (size_t) (ncol * per_col))) return -1;
  StageTimer tm(ctx, ST_ERUPT);
  P.t_prev = time - dt;                                              // :454
  P.dt = dt; P.add_step = add_step;
  P.base = base; P.cap = (int) std::min<size_t>(ctx->cap, 0x7fffffff); P.rank = ctx->cfg.rank; P.has_grad = ctx->has_grad;
  P.nbox = ctx->n_erupt_boxes;
  k_add_erupt<<<1, ERUPT_THREADS, 0, ctx->stream>>>(ctx->cur.t, psph_grid_view(ctx), ctx->kp, ctx->d_ss, P, ctx->d_erupt_boxes, owner);
  PSPH_LAUNCH_CHECK();
  psph_comm_mark(ctx, DIRTY_LAYOUT);
  ctx->n = base + (int) (ncol * per_col);                            // upper bound until the count is read back
  ctx->n_dirty = true;
  ctx->list_valid = false;
  if (added) {
    PSPH_CUDA(cudaMemcpyAsync(ctx->h_ss, ctx->d_ss, sizeof(StepState), cudaMemcpyDeviceToHost, ctx->stream));
    PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
    *added = ctx->h_ss->n - base;
    ctx->n = ctx->h_ss->n;
    ctx->n_dirty = false;
  }
  return 0;
}
