// psph_ctx.cu — context life cycle, configuration defaults, host<->device particle traffic.
// C ABI declared in include/psph.h; reference call sites cited there.
#include "psph_internal.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

// ---------------------------------------------------------------------------------------
// defaults (src/plume/options.h, constant.h, parameters.h)
// ---------------------------------------------------------------------------------------
extern "C" void psph_config_default_3d(psph_config *c)
{
  memset(c, 0, sizeof *c);
  c->abi_version = PSPH_ABI_VERSION;
  c->dim = 3;
  c->use_gsph = 0;              // options.h:470
  c->riemann_solver = 0;        // options.h:544 (Roe)
  c->gsph_reconstruct = 3;      // options.h:552-596 (used when use_gsph is switched on)
  c->have_turbulence_lans = 1;  // options.h:700
  c->have_energy_smooth = 1;    // options.h:730
  c->pghost_contribute_des = 1; // options.h:760
  c->fluid_compressibility = 0;
  c->adaptive_sml = 1;
  c->sml_update_int = 50;       // constant.h:189
  c->cutoff = 3.0; c->cutoff2 = 8.0; c->neigh_radius = 4.25;
  c->msfrc_thresh = 1.0e-6; c->epsilon_half = 0.4; c->prandtl_num = 0.85;
  c->heat_trans_scale_ratio = 1.0; c->e_smooth_ratio = 4.0; c->tiny = 1.0e-8;
  c->alf = 0.3; c->beta = 0.6; c->ata = 0.01;   // parameters.h:996-998
  c->g = 9.80665;
  c->cfl = 0.8; c->eta_smooth = 1.5; c->c_smooth = 2.0; c->thresh = 1e-5; c->num_loop = 2;
  c->ng0 = 0.026; c->Cvs = 1100; c->Cvg = 1348; c->Cva = 717; c->Rg = 462; c->Ra = 287;
  c->rhoa0 = 1.10894; c->pa0 = 0.831e5;
  // Eyjafjallajokull domain, smoothing length 100 (bin/simulation.data), preprocess.cc:99-133
  const double Ll[3] = { -20800, -20800, 1666 }, Lu[3] = { 20800, 20800, 24000 };
  const double smlen = 100.0, del = 8 * smlen;
  for (int k = 0; k < 3; k++) {
    c->mindom[k] = Ll[k] - 1.5 * del;
    c->ncell[k] = (int) ceil((Lu[k] + 1.5 * del - c->mindom[k]) / del);
  }
  c->del = del; c->len_scale = 1.0;
  c->erupt_top = Ll[2]; c->smlen_real = smlen;
  c->ptable_size = 4000000;
  c->nranks = 1;
}

extern "C" void psph_config_default_1d(psph_config *c)
{
  memset(c, 0, sizeof *c);
  c->abi_version = PSPH_ABI_VERSION;
  c->dim = 1;
  c->use_gsph = 1;              // options.h:30
  c->riemann_solver = 1;        // HLLC
  c->have_turbulence_lans = 0;
  c->have_energy_smooth = 0;
  c->pghost_contribute_des = 2;
  c->fluid_compressibility = 0;
  c->adaptive_sml = 1;
  c->sml_update_int = 1;        // main.cc:690-694 calls adaptive_sml every step
  c->cutoff = 3.0; c->cutoff2 = 12.0; c->neigh_radius = 4.25;
  c->msfrc_thresh = 1.0e-6; c->epsilon_half = 0.4; c->prandtl_num = 0.85;
  c->heat_trans_scale_ratio = 1.0; c->e_smooth_ratio = 4.0; c->tiny = 1.0e-8;
  c->alf = 1.0; c->beta = 2.0; c->ata = 0.01;   // parameters.h:1401-1403
  c->g = 0.0;
  c->cfl = 0.1; c->eta_smooth = 1.0; c->c_smooth = 2.0; c->thresh = 1e-5; c->num_loop = 5;
  c->ng0 = 0.05; c->Cvs = 1100; c->Cvg = 1348; c->Cva = 717; c->Rg = 462; c->Ra = 287;
  c->rhoa0 = 1.104; c->pa0 = 1.0125e5;
  c->mindom[0] = -1.0; c->del = 2.0; c->len_scale = 1.0; // octant slicing range only (no mesh in 1-D)
  c->ncell[0] = c->ncell[1] = c->ncell[2] = 1;
  c->ptable_size = 10000;
  c->nranks = 1;
}

// ---------------------------------------------------------------------------------------
// Hilbert rank of the background cells.  The reference keys a bucket by the Hilbert index
// of its centre normalised to the mesh box (src/preproc/preprocess.cc:196-205 -> HSFC3d,
// src/datstr/hilbert.c:156-271,289-303) and partitions/iterates along that key.  Here the
// key only has to be computed once per cell, on the host, to get each cell's rank.
// ---------------------------------------------------------------------------------------
static uint64_t hilbert_key_3d(const double nc[3])
{
  // transition tables of the 3-D Hilbert curve in the orientation used by Zoltan/hilbert.c:
  // per octant, which old axis slot feeds each new slot and whether it is reflected
  static const unsigned char gray_inv[8] = { 0, 1, 3, 2, 7, 6, 4, 5 };
  static const unsigned char from[8][3] = { {2,1,0}, {0,2,1}, {0,1,2}, {2,0,1}, {2,0,1}, {0,1,2}, {0,2,1}, {2,1,0} };
  static const unsigned char refl[8][3] = { {0,0,0}, {0,0,0}, {0,0,0}, {1,1,0}, {0,1,1}, {0,0,0}, {0,1,1}, {1,0,1} };
  uint32_t q[3];
  for (int d = 0; d < 3; d++) {
    double v = nc[d] <= 0 ? 0.0 : 4294967295.0 * nc[d];
    q[d] = (uint32_t) (v > 4294967295.0 ? 4294967295.0 : v);
  }
  unsigned char ax[3] = { 0, 2, 4 };
  uint64_t H = 0;
  for (int lvl = 1; lvl <= 21; lvl++) {
    unsigned s = 32 - lvl, g = 0;
    for (int a = 0; a < 3; a++) g |= (((q[ax[a] >> 1] >> s) ^ ax[a]) & 1u) << a;
    unsigned oct = gray_inv[g];
    H = (H << 3) | oct;
    unsigned char o[3] = { ax[0], ax[1], ax[2] };
    for (int a = 0; a < 3; a++) ax[a] = o[from[oct][a]] ^ refl[oct][a];
  }
  return H >> 1; // the reference's 2-word key drops the last bit
}

static void build_cell_ranks(psph_ctx *ctx)
{
  const psph_config &c = ctx->cfg;
  long long nc = ctx->ncells;
  std::vector<uint64_t> key(nc);
  if (c.dim == 3) {
    for (int i = 0; i < c.ncell[0]; i++)
      for (int j = 0; j < c.ncell[1]; j++)
        for (int k = 0; k < c.ncell[2]; k++) {
          int idx[3] = { i, j, k };
          double nrm[3];
          for (int d = 0; d < 3; d++) {
            double lo = c.mindom[d] + idx[d] * c.del, hi = c.mindom[d] + (idx[d] + 1) * c.del;
            double maxdom = c.mindom[d] + c.ncell[d] * c.del;
            nrm[d] = ((lo + hi) * 0.5 - c.mindom[d]) / (maxdom - c.mindom[d]);
          }
          key[((long long) i * c.ncell[1] + j) * c.ncell[2] + k] = hilbert_key_3d(nrm);
        }
  } else
    for (long long q = 0; q < nc; q++) key[q] = q;
  std::vector<int32_t> order(nc);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return key[a] < key[b]; });
  ctx->h_cell2rank.assign(nc, 0);
  for (long long r = 0; r < nc; r++) ctx->h_cell2rank[order[r]] = (int32_t) r;
  cudaMemcpy(ctx->d_rank2cell, order.data(), nc * sizeof(int32_t), cudaMemcpyHostToDevice);
  cudaMemcpy(ctx->d_cell2rank, ctx->h_cell2rank.data(), nc * sizeof(int32_t), cudaMemcpyHostToDevice);
}

// ---------------------------------------------------------------------------------------
GridView psph_grid_view(const psph_ctx *ctx)
{
  GridView G;
  for (int d = 0; d < 3; d++) { G.ncell[d] = ctx->cfg.ncell[d]; G.mindom[d] = ctx->cfg.mindom[d]; }
  G.dim = ctx->cfg.dim;
  G.ncells = ctx->ncells;
  G.del = ctx->cfg.del; G.len_scale = ctx->cfg.len_scale;
  G.inv_subw = 1.0 / (ctx->cfg.del / (ctx->cfg.len_scale * SUBDIV));
  G.cell2rank = ctx->d_cell2rank; G.rank2cell = ctx->d_rank2cell; G.sub_start = ctx->d_sub_start;
  G.cell_has_ghost = ctx->d_cell_has_ghost;
  G.cell_hmax = ctx->d_cell_hmax;
  G.attr = ctx->d_attr;
  return G;
}

static void fill_kparams(psph_ctx *ctx)
{
  const psph_config &c = ctx->cfg;
  KParams &k = ctx->kp;
  k.dim = c.dim; k.use_gsph = c.use_gsph; k.riemann_solver = c.riemann_solver;
  k.lans = c.have_turbulence_lans; k.energy_smooth = c.have_energy_smooth;
  k.pghost = c.pghost_contribute_des; k.compress = c.fluid_compressibility; k.gsph_recon = c.gsph_reconstruct;
  k.cutoff = c.cutoff; k.cutoff2 = c.cutoff2; k.neigh_radius = c.neigh_radius;
  k.msfrc_thresh = c.msfrc_thresh; k.epsilon_half = c.epsilon_half; k.prandtl = c.prandtl_num;
  k.heat_ratio = c.heat_trans_scale_ratio; k.e_ratio = c.e_smooth_ratio; k.tiny = c.tiny;
  {
    const double ce[14] = PSPH_EXP_COEF;
    for (int q = 0; q < 14; q++) k.cexp[q] = ce[q];
  }
  k.heat_c = 5.0 * c.epsilon_half / c.prandtl_num;
  {
    const double b1 = c.cutoff * (1.0 + 0x1p-40);
    k.near_c = c.e_smooth_ratio * c.e_smooth_ratio / (c.dim * b1 * b1);
  }
  k.alf = c.alf; k.beta = c.beta; k.ata = c.ata; k.g = c.g; k.cfl = c.cfl;
  k.eta_smooth = c.eta_smooth; k.c_smooth = c.c_smooth; k.thresh = c.thresh; k.num_loop = c.num_loop;
  k.ng0 = c.ng0; k.Cvs = c.Cvs; k.Cvg = c.Cvg; k.Cva = c.Cva; k.Rg = c.Rg; k.Ra = c.Ra;
  k.rhoa0 = c.rhoa0; k.pa0 = c.pa0; k.erupt_top = c.erupt_top; k.smlen_real = c.smlen_real;
  for (int d = 0; d < 3; d++) k.org[d] = c.mindom[d] / c.len_scale;
}

extern "C" int psph_create(const psph_config *cfg, psph_ctx **out)
{
  if (!cfg || !out || cfg->abi_version != PSPH_ABI_VERSION) return -2;
  if (cfg->dim != 1 && cfg->dim != 3) return -2;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return -3; // no CUDA device: fail loudly, no fallback
  if (cudaSetDevice(cfg->device) != cudaSuccess) return -3;
  psph_ctx *ctx = new psph_ctx();
  ctx->cfg = *cfg;
  fill_kparams(ctx);
  ctx->n = 0; ctx->cap = 0; ctx->has_grad = cfg->use_gsph != 0;
  memset(&ctx->cur, 0, sizeof ctx->cur); memset(&ctx->alt, 0, sizeof ctx->alt);
  ctx->ncells = (long long) cfg->ncell[0] * cfg->ncell[1] * cfg->ncell[2];
  if (ctx->ncells <= 0 || ctx->ncells * SUBS >= 0x7fffffffLL) { delete ctx; return -2; }
  ctx->grow = nullptr;
  ctx->d_attr = nullptr; ctx->d_sortkey = ctx->d_sortkey_alt = ctx->d_perm = ctx->d_perm_alt = nullptr;
  ctx->d_cub_tmp = nullptr; ctx->cub_tmp_bytes = 0;
  ctx->list_valid = false; ctx->n_dirty = false; ctx->step = 0; ctx->launches = 0;
  ctx->d_aos = nullptr; ctx->aos_cap = 0; ctx->comm = nullptr; ctx->d_gspill = nullptr;
  {
    const char *gm = getenv("PSPH_GATHER"); // A/B switch for profiling; both paths are CUDA
    ctx->gather_mode = (gm && !strcmp(gm, "row")) ? 1 : 0;
    const char *gw = getenv("PSPH_GWARPS");
    ctx->gather_warps = gw ? atoi(gw) : 0;
    const char *gp = getenv("PSPH_GPF");
    ctx->gather_pf = gp ? atoi(gp) : -1;
    const char *lw = getenv("PSPH_LWARPS");
    ctx->list_warps = lw ? atoi(lw) : 0;
    const char *tw = getenv("PSPH_TWARPS");
    ctx->tile_shape = tw ? atoi(tw) : 0;
    const char *nb = getenv("PSPH_NBR");
    ctx->nbr_mode = nb ? atoi(nb) : 5;
    ctx->mix_rows = 0; ctx->mix_wanted = ctx->nbr_mode == 4; ctx->h_unknown = true;
    const char *sc = getenv("PSPH_SCAN");
    ctx->scan_mode = sc ? atoi(sc) : 2;
    const char *dn = getenv("PSPH_DENSNEAR");
    ctx->dens_near = dn ? atoi(dn) : 1;
    const char *mw = getenv("PSPH_MWARPS");
    ctx->mix_warps = mw ? atoi(mw) : 16;
    const char *wc = getenv("PSPH_WCACHE");
    ctx->wcache = wc ? atoi(wc) : 1;
    ctx->has_erupt = false; ctx->d_erupt_boxes = nullptr; ctx->n_erupt_boxes = 0;
    ctx->d_img_keys = nullptr; ctx->d_img_coord = nullptr; ctx->d_img_cell = nullptr; ctx->d_img_row = nullptr; ctx->n_img = 0;
    ctx->nbr_units = 0; ctx->nbr_want_tried = 0; ctx->nbr_valid = false; ctx->nbr_kind = 0;
  }
  ctx->timing = 0; memset(ctx->stage_ms, 0, sizeof ctx->stage_ms);
  *out = ctx;
  ctx->sms = 0;
  cudaDeviceGetAttribute(&ctx->sms, cudaDevAttrMultiProcessorCount, cfg->device);
  if (ctx->sms <= 0) ctx->sms = 148;
  PSPH_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  PSPH_CUDA(cudaEventCreate(&ctx->ev0));
  PSPH_CUDA(cudaEventCreate(&ctx->ev1));
  PSPH_CUDA(cudaEventCreate(&ctx->kev0));
  PSPH_CUDA(cudaEventCreate(&ctx->kev1));
  memset(ctx->kernel_ms, 0, sizeof ctx->kernel_ms); memset(ctx->kernel_calls, 0, sizeof ctx->kernel_calls);
  PSPH_CUDA(cudaMalloc(&ctx->d_cell2rank, ctx->ncells * sizeof(int32_t)));
  PSPH_CUDA(cudaMalloc(&ctx->d_rank2cell, ctx->ncells * sizeof(int32_t)));
  PSPH_CUDA(cudaMalloc(&ctx->d_sub_start, (ctx->ncells * SUBS + 1) * sizeof(int32_t)));
  PSPH_CUDA(cudaMemset(ctx->d_sub_start, 0, (ctx->ncells * SUBS + 1) * sizeof(int32_t)));
  PSPH_CUDA(cudaMalloc(&ctx->d_blk_any, ctx->ncells * 4));
  PSPH_CUDA(cudaMemset(ctx->d_blk_any, 0, ctx->ncells * 4));
  PSPH_CUDA(cudaMalloc(&ctx->d_cell_has_ghost, ctx->ncells));
  PSPH_CUDA(cudaMemset(ctx->d_cell_has_ghost, 0, ctx->ncells));
  PSPH_CUDA(cudaMalloc(&ctx->d_cell_hmax, ctx->ncells * sizeof(unsigned long long)));
  PSPH_CUDA(cudaMemset(ctx->d_cell_hmax, 0, ctx->ncells * sizeof(unsigned long long)));
  PSPH_CUDA(cudaMalloc(&ctx->d_ss, sizeof(StepState)));
  PSPH_CUDA(cudaMemset(ctx->d_ss, 0, sizeof(StepState)));
  PSPH_CUDA(cudaMallocHost(&ctx->h_ss, sizeof(StepState)));
  memset(ctx->h_ss, 0, sizeof(StepState));
  ctx->h_ss->dt_max = 1.0e300; ctx->h_ss->erupting = 1;
  PSPH_CUDA(cudaMemcpy(ctx->d_ss, ctx->h_ss, sizeof(StepState), cudaMemcpyHostToDevice));
  build_cell_ranks(ctx);
  PSPH_CUDA(cudaGetLastError());
  return 0;
}

static void free_tab(Tab &v)
{
  for (int c = 0; c < D_NCOL; c++) { cudaFree(v.d[c]); v.d[c] = nullptr; }
  for (int c = 0; c < I_NCOL; c++) { cudaFree(v.i[c]); v.i[c] = nullptr; }
}

extern "C" void psph_destroy(psph_ctx *ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->cfg.device);
  cudaStreamSynchronize(ctx->stream);
  psph_comm_destroy(ctx);
  free_tab(ctx->cur.t); free_tab(ctx->alt);
  for (int c = 0; c < NSCRATCH; c++) cudaFree(ctx->cur.scr[c]);
  for (int c = 0; c < NREC; c++) cudaFree(ctx->cur.rec[c]);
  cudaFree(ctx->cur.geo); cudaFree(ctx->cur.geo32); cudaFree(ctx->cur.nbr); cudaFree(ctx->cur.wts); cudaFree(ctx->cur.nbr_off); cudaFree(ctx->cur.nbr_cnt); cudaFree(ctx->cur.near); cudaFree(ctx->cur.mlist); cudaFree(ctx->cur.mseg); cudaFree(ctx->cur.flags); cudaFree(ctx->d_gspill);
  cudaFree(ctx->cur.gsub); cudaFree(ctx->cur.need_idx); cudaFree(ctx->cur.grad_idx);
  psph_grow_free(ctx);
  cudaFree(ctx->d_cell2rank); cudaFree(ctx->d_rank2cell); cudaFree(ctx->d_sub_start); cudaFree(ctx->d_attr);
  cudaFree(ctx->d_cell_has_ghost); cudaFree(ctx->d_cell_hmax); cudaFree(ctx->d_blk_any);
  cudaFree(ctx->d_erupt_boxes);
  cudaFree(ctx->d_img_keys); cudaFree(ctx->d_img_coord); cudaFree(ctx->d_img_cell); cudaFree(ctx->d_img_row);
  cudaFree(ctx->d_sortkey); cudaFree(ctx->d_sortkey_alt); cudaFree(ctx->d_perm); cudaFree(ctx->d_perm_alt);
  cudaFree(ctx->d_cub_tmp); cudaFree(ctx->d_ss); cudaFreeHost(ctx->h_ss);
  cudaFree(ctx->d_aos);
  cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1);
  cudaEventDestroy(ctx->kev0); cudaEventDestroy(ctx->kev1);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char *psph_last_error(const psph_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" size_t psph_num_particles(const psph_ctx *ctx) { return ctx ? (size_t) ctx->n : 0; }
extern "C" void *psph_stream(psph_ctx *ctx) { return ctx ? (void *) ctx->stream : nullptr; }
// pinned host staging for psph_upload / psph_download buffers (plain malloc semantics otherwise)
extern "C" void *psph_host_alloc(size_t bytes)
{
  void *p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
  return p;
}
extern "C" void psph_host_free(void *p) { if (p) cudaFreeHost(p); }
extern "C" int psph_sync(psph_ctx *ctx) { PSPH_CUDA(cudaStreamSynchronize(ctx->stream)); return 0; }
extern "C" int64_t psph_launch_count(psph_ctx *ctx, int reset)
{
  long long v = ctx->launches;
  if (reset) ctx->launches = 0;
  return v;
}

// grow one device array to ncap elements, keeping the first `keep` (0: contents dropped)
template <class T>
static int grow(psph_ctx *ctx, T *&p, size_t keep, size_t ncap)
{
  T *q = nullptr;
  PSPH_CUDA(cudaMalloc(&q, ncap * sizeof(T)));
  PSPH_CUDA(cudaMemsetAsync(q, 0, ncap * sizeof(T), ctx->stream));
  if (keep && p) PSPH_CUDA(cudaMemcpyAsync(q, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
  PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(p);
  p = q;
  return 0;
}

// grow every column (and the sort scratch) to hold `rows`; contents of cur are preserved
int psph_reserve(psph_ctx *ctx, size_t rows)
{
  if (rows <= ctx->cap) return 0;
  size_t ncap = std::max<size_t>(rows, ctx->cap + ctx->cap / 2 + 1024);
  ncap = (ncap + 63) & ~(size_t) 63;
  size_t keep = (size_t) ctx->n;
  int nd = ctx->has_grad ? D_NCOL : D_GRAD0;
  for (int c = 0; c < nd; c++) {
    if (grow(ctx, ctx->cur.t.d[c], keep, ncap)) return -1;
    if (grow(ctx, ctx->alt.d[c], 0, ncap)) return -1;
  }
  for (int c = 0; c < I_NCOL; c++) {
    if (grow(ctx, ctx->cur.t.i[c], keep, ncap)) return -1;
    if (grow(ctx, ctx->alt.i[c], 0, ncap)) return -1;
  }
  for (int c = 0; c < NSCRATCH; c++) if (grow(ctx, ctx->cur.scr[c], 0, ncap)) return -1;
  for (int c = 0; c < NREC; c++) if (grow(ctx, ctx->cur.rec[c], 0, ncap)) return -1;
  if (grow(ctx, ctx->cur.geo, 0, ncap)) return -1;
  if (grow(ctx, ctx->cur.geo32, 0, ncap)) return -1;
  if (grow(ctx, ctx->cur.nbr_off, 0, ncap)) return -1;
  if (grow(ctx, ctx->cur.nbr_cnt, 0, ncap)) return -1;
  if (grow(ctx, ctx->cur.near, 0, ncap * NEARW)) return -1;
  ctx->nbr_valid = false;
  if (grow(ctx, ctx->cur.flags, 0, ncap)) return -1;
  if (grow(ctx, ctx->cur.gsub, 0, ncap)) return -1;
  if (grow(ctx, ctx->cur.need_idx, 0, ncap)) return -1;
  if (ctx->has_grad && grow(ctx, ctx->cur.grad_idx, 0, ncap)) return -1;
  if (grow(ctx, ctx->d_sortkey, 0, ncap)) return -1;
  if (grow(ctx, ctx->d_sortkey_alt, 0, ncap)) return -1;
  if (grow(ctx, ctx->d_perm, 0, ncap)) return -1;
  if (grow(ctx, ctx->d_perm_alt, 0, ncap)) return -1;
  cudaFree(ctx->d_cub_tmp); ctx->d_cub_tmp = nullptr; ctx->cub_tmp_bytes = 0;
  ctx->cap = ncap;
  ctx->list_valid = false;
  return 0;
}

extern "C" int psph_set_cells(psph_ctx *ctx, const psph_cell_attr *attr, size_t ncells)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  if (!attr) { cudaFree(ctx->d_attr); ctx->d_attr = nullptr; return 0; }
  if ((long long) ncells != ctx->ncells) { ctx->err = "psph_set_cells: wrong cell count"; return -2; }
  if (!ctx->d_attr) PSPH_CUDA(cudaMalloc(&ctx->d_attr, ncells * sizeof(psph_cell_attr)));
  PSPH_CUDA(cudaMemcpyAsync(ctx->d_attr, attr, ncells * sizeof(psph_cell_attr), cudaMemcpyHostToDevice, ctx->stream));
  PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ---------------------------------------------------------------------------------------
// AoS (host, reference ParticlePack-like rows) <-> SoA (device columns)
// ---------------------------------------------------------------------------------------
// rows are `stride` bytes apart: sizeof(psph_particle), or PSPH_ROW_BYTES_SPH (the struct without its gradient block)
__global__ void k_aos_to_soa(const char *__restrict__ rows, size_t stride, int n, int base, Tab V, GridView G, int has_grad, int rank)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const psph_particle &r = *reinterpret_cast<const psph_particle *>(rows + (size_t) t * stride);
  has_grad = has_grad && stride >= sizeof(psph_particle);
  int p = base + t;
  V.d[D_X][p] = r.coords[0]; V.d[D_Y][p] = r.coords[1]; V.d[D_Z][p] = r.coords[2];
  V.d[D_VX][p] = r.vel[0]; V.d[D_VY][p] = r.vel[1]; V.d[D_VZ][p] = r.vel[2];
  V.d[D_SVX][p] = r.smoothed_v[0]; V.d[D_SVY][p] = r.smoothed_v[1]; V.d[D_SVZ][p] = r.smoothed_v[2];
  V.d[D_RHO][p] = r.rho; V.d[D_E][p] = r.engr; V.d[D_SE][p] = r.smoothed_e;
  V.d[D_MASS][p] = r.mass; V.d[D_H][p] = r.smlen; V.d[D_HLIST][p] = r.smlen; V.d[D_KS][p] = r.mass_frac;
  V.d[D_P][p] = r.pressure; V.d[D_GAMMA][p] = r.gamma; V.d[D_C][p] = r.sound_speed;
  V.d[D_T][p] = r.temperature; V.d[D_CP][p] = r.specific_heat_p;
  V.d[D_RPH0][p] = r.phase_density[0]; V.d[D_RPH1][p] = r.phase_density[1];
  if (has_grad)
    for (int k = 0; k < 15; k++) V.d[D_GRAD0 + k][p] = r.grad[k];
  else if (V.d[D_GRAD0])
    for (int k = 0; k < 15; k++) V.d[D_GRAD0 + k][p] = 0.0;
  V.i[I_BC][p] = r.bc_type; V.i[I_PHASE][p] = r.phase_num; V.i[I_INVOLVED][p] = r.involved;
  V.i[I_GUEST][p] = r.guest; V.i[I_MYPROC][p] = r.guest ? r.myprocess : rank;
  V.i[I_KEY0][p] = (int32_t) r.key[0]; V.i[I_KEY1][p] = (int32_t) r.key[1]; V.i[I_KEY2][p] = (int32_t) r.key[2];
  // bucket membership by position: [mincrd, maxcrd) against the mesh bounds exactly as the
  // reference stores them (dataread.cc:380-387; membership test update_pos.cc:257-271)
  int idx[3] = { 0, 0, 0 };
  if (G.dim == 3)
    for (int d = 0; d < 3; d++) idx[d] = dev_cell_index(G, d, r.coords[d]);
  V.i[I_CELL][p] = (idx[0] * G.ncell[1] + idx[1]) * G.ncell[2] + idx[2];
}

__global__ void k_soa_to_aos(char *__restrict__ rows, size_t stride, int n, Tab V, int has_grad)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  psph_particle &r = *reinterpret_cast<psph_particle *>(rows + (size_t) p * stride);
  r.key[0] = (uint32_t) V.i[I_KEY0][p]; r.key[1] = (uint32_t) V.i[I_KEY1][p]; r.key[2] = (uint32_t) V.i[I_KEY2][p];
  r.bc_type = V.i[I_BC][p]; r.phase_num = V.i[I_PHASE][p]; r.involved = V.i[I_INVOLVED][p];
  r.guest = V.i[I_GUEST][p]; r.myprocess = V.i[I_MYPROC][p];
  r.mass = V.d[D_MASS][p]; r.smlen = V.d[D_H][p]; r.mass_frac = V.d[D_KS][p];
  r.pressure = V.d[D_P][p]; r.gamma = V.d[D_GAMMA][p]; r.sound_speed = V.d[D_C][p];
  r.temperature = V.d[D_T][p]; r.specific_heat_p = V.d[D_CP][p];
  r.phase_density[0] = V.d[D_RPH0][p]; r.phase_density[1] = V.d[D_RPH1][p];
  r.coords[0] = V.d[D_X][p]; r.coords[1] = V.d[D_Y][p]; r.coords[2] = V.d[D_Z][p];
  r.vel[0] = V.d[D_VX][p]; r.vel[1] = V.d[D_VY][p]; r.vel[2] = V.d[D_VZ][p];
  r.smoothed_v[0] = V.d[D_SVX][p]; r.smoothed_v[1] = V.d[D_SVY][p]; r.smoothed_v[2] = V.d[D_SVZ][p];
  r.rho = V.d[D_RHO][p]; r.engr = V.d[D_E][p]; r.smoothed_e = V.d[D_SE][p];
  if (stride >= sizeof(psph_particle))
    for (int k = 0; k < 15; k++) r.grad[k] = has_grad ? V.d[D_GRAD0 + k][p] : 0.0;
}

int psph_ensure_aos(psph_ctx *ctx, size_t n)
{
  if (n > ctx->aos_cap) {
    cudaFree(ctx->d_aos);
    size_t c = n + n / 4 + 1024;
    PSPH_CUDA(cudaMalloc(&ctx->d_aos, c * sizeof(psph_particle)));
    ctx->aos_cap = c;
  }
  return 0;
}

// exact live row count: the re-bin flags lost rows and the next list build drops them, so the
// host mirror is only an upper bound until it is read back
int psph_sync_n(psph_ctx *ctx)
{
  if (!ctx->list_valid && ctx->n > 0 && psph_build_cell_list(ctx)) return -1;
  if (ctx->n_dirty) {
    PSPH_CUDA(cudaMemcpyAsync(ctx->h_ss, ctx->d_ss, sizeof(StepState), cudaMemcpyDeviceToHost, ctx->stream));
    PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->n = ctx->h_ss->n;
    ctx->n_dirty = false;
    if (ctx->h_ss->pair_mode == 4) ctx->mix_wanted = true;   // per-row smoothing lengths arrived with the rows
  }
  return 0;
}

static bool row_bytes_ok(psph_ctx *ctx, size_t stride, const char *who)
{
  if (stride == sizeof(psph_particle) || stride == PSPH_ROW_BYTES_SPH) return true;
  ctx->err = std::string(who) + ": row_bytes must be sizeof(psph_particle) or PSPH_ROW_BYTES_SPH";
  return false;
}

static int upload_at(psph_ctx *ctx, const void *rows, size_t n, size_t base, size_t stride = sizeof(psph_particle))
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  if (psph_reserve(ctx, base + n)) return -1;
  if (n) {
    if (psph_ensure_aos(ctx, n)) return -1;
    PSPH_CUDA(cudaMemcpyAsync(ctx->d_aos, rows, n * stride, cudaMemcpyHostToDevice, ctx->stream));
    k_aos_to_soa<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->stream>>>(reinterpret_cast<const char *>(ctx->d_aos), stride, (int) n,
                                                                         (int) base, ctx->cur.t, psph_grid_view(ctx), ctx->has_grad,
                                                                         ctx->cfg.rank);
    PSPH_LAUNCH_CHECK();
  }
  psph_comm_mark(ctx, DIRTY_LAYOUT);
  ctx->h_unknown = true;
  ctx->n = (int) (base + n);
  ctx->h_ss->n = ctx->n;
  PSPH_CUDA(cudaMemcpyAsync(&ctx->d_ss->n, &ctx->h_ss->n, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->list_valid = false;
  return 0;
}

extern "C" int psph_upload(psph_ctx *ctx, const psph_particle *rows, size_t n) { ctx->n = 0; return upload_at(ctx, rows, n, 0); }
extern "C" int psph_append(psph_ctx *ctx, const psph_particle *rows, size_t n)
{
  if (psph_sync_n(ctx)) return -1;
  return upload_at(ctx, rows, n, ctx->n);
}

extern "C" int psph_upload_rows(psph_ctx *ctx, const void *rows, size_t n, size_t row_bytes)
{
  if (!row_bytes_ok(ctx, row_bytes, "psph_upload_rows")) return -2;
  ctx->n = 0;
  return upload_at(ctx, rows, n, 0, row_bytes);
}
extern "C" int psph_append_rows(psph_ctx *ctx, const void *rows, size_t n, size_t row_bytes)
{
  if (!row_bytes_ok(ctx, row_bytes, "psph_append_rows")) return -2;
  if (psph_sync_n(ctx)) return -1;
  return upload_at(ctx, rows, n, ctx->n, row_bytes);
}

extern "C" int psph_download(psph_ctx *ctx, psph_particle *rows, size_t cap, size_t *n_out, int order)
{ return psph_download_rows(ctx, rows, cap, n_out, order, sizeof(psph_particle)); }

extern "C" int psph_download_rows(psph_ctx *ctx, void *rows_v, size_t cap, size_t *n_out, int order, size_t stride)
{
  if (!row_bytes_ok(ctx, stride, "psph_download_rows")) return -2;
  char *rows = static_cast<char *>(rows_v);
  auto key_of = [&](size_t i) { return reinterpret_cast<const psph_particle *>(rows + i * stride)->key; };
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  if (psph_sync_n(ctx)) return -1;
  size_t n = ctx->n;
  if (n_out) *n_out = n;
  if (!rows) return 0;
  if (cap < n) { ctx->err = "psph_download: buffer too small"; return -2; }
  if (n == 0) return 0;
  if (psph_ensure_aos(ctx, n)) return -1;
  k_soa_to_aos<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->stream>>>(reinterpret_cast<char *>(ctx->d_aos), stride, (int) n, ctx->cur.t,
                                                                       ctx->has_grad);
  PSPH_LAUNCH_CHECK();
  PSPH_CUDA(cudaMemcpyAsync(rows, ctx->d_aos, n * stride, cudaMemcpyDeviceToHost, ctx->stream));
  PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
  if (order == PSPH_ORDER_REFERENCE) {
    // THashTable iteration order: ascending hash slot, chains sorted by key
    // (src/datstr/thashtab.cc:204-222, :86-134, :246-262)
    std::vector<std::pair<long long, size_t>> slot(n);
    const double umax = 4294967295.0;
    int S02 = ctx->cfg.ptable_size / 10, S01 = ctx->cfg.ptable_size - S02;
    for (size_t i = 0; i < n; i++) {
      const uint32_t *ki = key_of(i);
      long long s = ctx->cfg.dim == 1 ? (long long) ki[0]
        : (long long) (int) ((double) ki[0] / umax * S01) + (int) ((double) ki[1] / umax * S02);
      slot[i] = std::make_pair(s, i);
    }
    std::sort(slot.begin(), slot.end(), [&](const std::pair<long long, size_t> &a, const std::pair<long long, size_t> &b) {
      if (a.first != b.first) return a.first < b.first;
      const uint32_t *ka = key_of(a.second), *kb = key_of(b.second);
      for (int k = 0; k < 3; k++) if (ka[k] != kb[k]) return ka[k] < kb[k];
      return a.second < b.second;
    });
    std::vector<char> tmp(rows, rows + n * stride);
    for (size_t i = 0; i < n; i++) memcpy(rows + i * stride, tmp.data() + slot[i].second * stride, stride);
  }
  return 0;
}

extern "C" int psph_cell_of_rows(psph_ctx *ctx, int32_t *cells, size_t cap)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  if (psph_sync_n(ctx)) return -1;
  if (cap < (size_t) ctx->n) { ctx->err = "psph_cell_of_rows: buffer too small"; return -2; }
  if (ctx->n) PSPH_CUDA(cudaMemcpyAsync(cells, ctx->cur.t.i[I_CELL], ctx->n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int psph_cell_ranks(psph_ctx *ctx, int32_t *rank_of_cell, size_t cap)
{
  if (cap < (size_t) ctx->ncells) { ctx->err = "psph_cell_ranks: buffer too small"; return -2; }
  memcpy(rank_of_cell, ctx->h_cell2rank.data(), ctx->ncells * sizeof(int32_t));
  return 0;
}

extern "C" int psph_column(psph_ctx *ctx, int col, double **dptr)
{
  if (col < 0 || col >= PSPH_COL_COUNT || !dptr) return -2;
  *dptr = ctx->cur.t.d[col];
  return 0;
}

__global__ void k_fill_from_soa(Tab V, GridView G, int n, const int32_t *bc, const int32_t *phase, int rank)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  V.i[I_BC][p] = bc ? bc[p] : 100; V.i[I_PHASE][p] = phase ? phase[p] : 1;
  V.i[I_INVOLVED][p] = 0; V.i[I_GUEST][p] = 0; V.i[I_MYPROC][p] = rank;
  V.i[I_KEY0][p] = p; V.i[I_KEY1][p] = rank; V.i[I_KEY2][p] = 0;
  V.d[D_HLIST][p] = V.d[D_H][p]; V.d[D_SE][p] = V.d[D_E][p];
  double ks = V.d[D_KS][p], rho = V.d[D_RHO][p];
  V.d[D_RPH0][p] = rho * (1 - ks); V.d[D_RPH1][p] = rho * ks;
  int idx[3] = { 0, 0, 0 };
  if (G.dim == 3)
    for (int d = 0; d < 3; d++) idx[d] = dev_cell_index(G, d, V.d[D_X + d][p]);
  V.i[I_CELL][p] = (idx[0] * G.ncell[1] + idx[1]) * G.ncell[2] + idx[2];
}

extern "C" int psph_upload_device_soa(psph_ctx *ctx, const double *const *cols, const int32_t *bc, const int32_t *phase, size_t n)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  ctx->n = 0;
  if (psph_reserve(ctx, n)) return -1;
  for (int c = 0; c < PSPH_COL_COUNT; c++) {
    if (!cols[c]) { ctx->err = "psph_upload_device_soa: missing column"; return -2; }
    PSPH_CUDA(cudaMemcpyAsync(ctx->cur.t.d[c], cols[c], n * sizeof(double), cudaMemcpyDefault, ctx->stream));
  }
  if (n) {
    k_fill_from_soa<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->cur.t, psph_grid_view(ctx), (int) n, bc, phase, ctx->cfg.rank);
    PSPH_LAUNCH_CHECK();
  }
  psph_comm_mark(ctx, DIRTY_LAYOUT);
  ctx->h_unknown = true;
  ctx->n = (int) n; ctx->h_ss->n = ctx->n;
  PSPH_CUDA(cudaMemcpyAsync(&ctx->d_ss->n, &ctx->h_ss->n, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->list_valid = false;
  return 0;
}
