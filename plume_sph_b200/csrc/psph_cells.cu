// psph_cells.cu — GPU cell-list build and re-binning.
//
// Replaces the reference's per-bucket particle lists (Bucket::particles, bucket.h:255-321,
// maintained by update_particles_in_bucket, update_pos.cc:172-462) and the 27-bucket walk of
// search_neighs_consth (neighs.cc:37-211): rows are radix-sorted by
//   key = rank(cell) * 512 + subcell(cell, position)      (subcell = sx*64 + sy*8 + sz)
// (rank = position of the cell along the reference's bucket Hilbert curve), every column is
// permuted into that order, and sub_start[key] gives the first row of every (cell, subcell).
// Rows flagged removed by the re-bin (cell == -1) sort to the end and are dropped.
// The sort is stable, so the row order (and with it every summation order) is a pure
// function of the input: results are reproducible run to run.
#include "psph_internal.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

__global__ void k_make_keys(Tab V, GridView G, StepState *ss, int n_cap, uint32_t inf_key,
                            uint32_t *__restrict__ key, uint32_t *__restrict__ perm)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  double h = 0.0;
  if (p < n_cap) {
    uint32_t k = inf_key;
    if (p < ss->n) {
      int cell = V.i[I_CELL][p];
      if (cell >= 0) {
        uint32_t sub = 0;
        if (G.dim == 3) {
          int ck = cell % G.ncell[2], cj = (cell / G.ncell[2]) % G.ncell[1], ci = cell / (G.ncell[2] * G.ncell[1]);
          sub = (uint32_t) (dev_sub_of(G, 0, ci, V.d[D_X][p]) * (SUBDIV * SUBDIV) +
                            dev_sub_of(G, 1, cj, V.d[D_Y][p]) * SUBDIV + dev_sub_of(G, 2, ck, V.d[D_Z][p]));
        }
        k = (uint32_t) G.cell2rank[cell] * (uint32_t) SUBS + sub;
        h = V.d[D_H][p];
      }
    }
    key[p] = k;
    perm[p] = p;
  }
  // max smoothing length (bounds the reach of the candidate boxes in the gathers)
  for (int o = 16; o; o >>= 1) h = fmax(h, __shfl_xor_sync(0xffffffffu, h, o));
  if ((threadIdx.x & 31) == 0 && h > 0.0) atomicMax(&ss->hmax_bits, (unsigned long long) __double_as_longlong(h));
}

__global__ void k_reset_list_scalars(StepState *ss) { ss->hmax_bits = 0ull; ss->n_need = 0; ss->n_grad = 0; ss->xmax_bits = 0u; ss->n_std = 0; }

// after the sort: number of live rows = first row whose key is inf_key
__global__ void k_count_live(const uint32_t *__restrict__ key, int n_cap, uint32_t inf_key, StepState *ss)
{
  int lo = 0, hi = n_cap;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (key[mid] < inf_key) lo = mid + 1; else hi = mid; }
  ss->n = lo;
  ss->hmax = __longlong_as_double((long long) ss->hmax_bits);
}

// permute every column into sorted order (coalesced writes; reads are near-coalesced because
// the order changes little from one step to the next)
__global__ void k_reorder(Tab src, Tab dst, const uint32_t *__restrict__ perm, const StepState *ss, int nd)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ss->n) return;
  uint32_t q = perm[p];
#pragma unroll 4
  for (int c = 0; c < nd; c++) dst.d[c][p] = src.d[c][q];
#pragma unroll
  for (int c = 0; c < I_NCOL; c++) dst.i[c][p] = src.i[c][q];
}

// sub_start[e] = lower_bound(sorted keys, e) for every (rank, subcell) entry e, plus the end sentinel
__global__ void k_sub_table(const uint32_t *__restrict__ key, const StepState *ss, long long nent,
                            int32_t *__restrict__ sub_start)
{
  long long e = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (e > nent) return;
  int lo = 0, hi = ss->n;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (key[mid] < (uint32_t) e) lo = mid + 1; else hi = mid; }
  sub_start[e] = lo;
}

// bucket "has a ghost particle" flags (neighs.cc:69) and the largest smoothing length of every bucket (the mixed-h pair
// search bounds a row's reach by the largest h a partner can have: a partner sits in one of the 27 buckets around it)
__global__ void k_ghost_cells(Tab V, const StepState *ss, uint8_t *__restrict__ has_ghost, unsigned long long *__restrict__ cell_hmax)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ss->n) return;
  int b = V.i[I_BC][p];
  const int cell = V.i[I_CELL][p];
  if (b == 0 || b == 1 || b == 2) has_ghost[cell] = 1;
  const unsigned long long hb = (unsigned long long) __double_as_longlong(V.d[D_H][p]);
  if (hb > cell_hmax[cell]) atomicMax(&cell_hmax[cell], hb);
}

// per-row flag word, geometry record, packed subcell coordinates, h at list-build time
__device__ __forceinline__ float row_flags(DevView &V, const GridView &G, const KParams &K, StepState *ss, int p)
{
  if (p >= ss->n) { V.flags[p] = 0; return 0.f; }
  int bc = V.t.i[I_BC][p], guest = V.t.i[I_GUEST][p], ph = V.t.i[I_PHASE][p];
  int f = 0;
  bool need = dev_need_neigh(bc, guest);
  // SPH build, 3-D: in buckets that hold a ghost only bc_type==100 rows get a list (neighs.cc:76-79)
  if (need && K.dim == 3 && !K.use_gsph && G.cell_has_ghost[V.t.i[I_CELL][p]] && bc != 100) need = false;
  if (need) f |= F_NEED;
  if (dev_contr_dens(K.pghost, bc)) f |= F_CDENS;
  if (dev_contr_vel(bc)) f |= F_CVEL;
  if (ph == 2) f |= F_PH2;
  if (ph == 1 || ph == 2) f |= F_PHOK;
  if (dev_is_real(bc)) f |= F_REAL;
  // rows that get gradients: every non-guest row that has a neighbour list (3-D GSPH build: all of them,
  // neighs.cc:142-201; 1-D: the need_neigh rows, add_air.cc:470-479)
  if (!guest && (K.dim == 3 || need)) f |= F_GRADROW;
  V.flags[p] = f;
  const double x[3] = { V.t.d[D_X][p], V.t.d[D_Y][p], V.t.d[D_Z][p] };
  double h = V.t.d[D_H][p];
  V.t.d[D_HLIST][p] = h;
  V.geo[p] = make_double4(x[0], x[1], x[2], h);
  const float xm = dev_store_geo32(V.geo32, K, p, x[0], x[1], x[2], h, ss->hmax);
  unsigned long long pk = 0;
  if (G.dim == 3) {
    // subcell of the row inside the bucket it is a member of (the same function the sort key uses),
    // plus a flag for rows within 1e-6 subcells of a subcell face (their candidate box is widened by one)
    const int cell = V.t.i[I_CELL][p];
    const int ci[3] = { cell / (G.ncell[2] * G.ncell[1]), (cell / G.ncell[2]) % G.ncell[1], cell % G.ncell[2] };
    unsigned long long near = 0;
    for (int d = 0; d < 3; d++) {
      const int sub = dev_sub_of(G, d, ci[d], x[d]);
      const double q = (x[d] - cell_min_d(G, d, ci[d])) * G.inv_subw - sub;
      if (!(q > 1e-6 && q < 1.0 - 1e-6)) near = 1;
      pk |= (unsigned long long) (ci[d] * SUBDIV + sub) << (20 * d);
    }
    pk |= near << 60;
  }
  V.gsub[p] = pk;
  return xm;
}
__global__ void k_row_flags(DevView V, GridView G, KParams K, StepState *ss, int n_cap)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  float xm = 0.f;
  if (p < n_cap) xm = row_flags(V, G, K, ss, p);
  // summed rows that carry the table's standard (largest) smoothing length: what k_pair_mode decides on
  const bool std_h = p < ss->n && (V.flags[p] & F_NEED) && V.t.d[D_H][p] == ss->hmax;
  const unsigned sb = __ballot_sync(0xffffffffu, std_h);
  if ((threadIdx.x & 31) == 0 && sb) atomicAdd(&ss->n_std, __popc(sb));
  // bound of the FP32 rounding error of the geo32 records (positive floats order like their bit patterns)
  for (int o = 16; o; o >>= 1) xm = fmaxf(xm, __shfl_xor_sync(0xffffffffu, xm, o));
  if ((threadIdx.x & 31) == 0 && xm > 0.f) atomicMax(&ss->xmax_bits, __float_as_uint(xm));
}

struct FlagSet
{
  const int32_t *flags; int bit;
  __device__ __forceinline__ bool operator()(int p) const { return (flags[p] & bit) != 0; }
};

// ascending list of the rows that carry `bit` -> idx, count -> *d_count (device)
static int select_rows(psph_ctx *ctx, int n_cap, int bit, int32_t *idx, int *d_count)
{
  cub::CountingInputIterator<int> it(0);
  FlagSet pred = { ctx->cur.flags, bit };
  size_t need = 0;
  cub::DeviceSelect::If(nullptr, need, it, idx, d_count, n_cap, pred, ctx->stream);
  if (need > ctx->cub_tmp_bytes) {
    PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_cub_tmp);
    PSPH_CUDA(cudaMalloc(&ctx->d_cub_tmp, need));
    ctx->cub_tmp_bytes = need;
  }
  cub::DeviceSelect::If(ctx->d_cub_tmp, need, it, idx, d_count, n_cap, pred, ctx->stream);
  ctx->launches += 2;
  PSPH_CUDA(cudaGetLastError());
  return 0;
}

int psph_build_cell_list(psph_ctx *ctx)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  int n_cap = ctx->n;
  long long nent = ctx->ncells * SUBS;
  uint32_t inf_key = (uint32_t) nent;
  GridView G = psph_grid_view(ctx);
  cudaStream_t s = ctx->stream;
  k_reset_list_scalars<<<1, 1, 0, s>>>(ctx->d_ss);
  PSPH_LAUNCH_CHECK();
  if (n_cap > 0) {
    k_make_keys<<<(n_cap + 255) / 256, 256, 0, s>>>(ctx->cur.t, G, ctx->d_ss, n_cap, inf_key, ctx->d_sortkey, ctx->d_perm);
    PSPH_LAUNCH_CHECK();
    int end_bit = 1;
    while ((1ull << end_bit) <= (unsigned long long) inf_key) end_bit++;
    size_t need = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, need, ctx->d_sortkey, ctx->d_sortkey_alt, ctx->d_perm, ctx->d_perm_alt,
                                    n_cap, 0, end_bit, s);
    if (need > ctx->cub_tmp_bytes) {
      PSPH_CUDA(cudaStreamSynchronize(s));
      cudaFree(ctx->d_cub_tmp);
      PSPH_CUDA(cudaMalloc(&ctx->d_cub_tmp, need));
      ctx->cub_tmp_bytes = need;
    }
    {
      KernelTimer kt(ctx, KN_SORT);
      cub::DeviceRadixSort::SortPairs(ctx->d_cub_tmp, need, ctx->d_sortkey, ctx->d_sortkey_alt, ctx->d_perm,
                                      ctx->d_perm_alt, n_cap, 0, end_bit, s);
    }
    ctx->launches += 1 + (end_bit + 7) / 8 * 2; // cub: histogram + (onesweep) one launch per 8-bit digit, approx.
    PSPH_CUDA(cudaGetLastError());
  }
  k_count_live<<<1, 1, 0, s>>>(ctx->d_sortkey_alt, n_cap, inf_key, ctx->d_ss);
  PSPH_LAUNCH_CHECK();
  if (n_cap > 0) {
    int nd = ctx->has_grad ? D_NCOL : D_GRAD0;
    {
      KernelTimer kt(ctx, KN_REORDER);
      k_reorder<<<(n_cap + 255) / 256, 256, 0, s>>>(ctx->cur.t, ctx->alt, ctx->d_perm_alt, ctx->d_ss, nd);
    }
    PSPH_LAUNCH_CHECK();
    Tab t = ctx->cur.t; ctx->cur.t = ctx->alt; ctx->alt = t;
  }
  k_sub_table<<<(unsigned) ((nent + 1 + 255) / 256), 256, 0, s>>>(ctx->d_sortkey_alt, ctx->d_ss, nent, ctx->d_sub_start);
  PSPH_LAUNCH_CHECK();
  PSPH_CUDA(cudaMemsetAsync(ctx->d_cell_has_ghost, 0, ctx->ncells, s));
  PSPH_CUDA(cudaMemsetAsync(ctx->d_cell_hmax, 0, ctx->ncells * sizeof(unsigned long long), s));
  if (n_cap > 0) {
    k_ghost_cells<<<(n_cap + 255) / 256, 256, 0, s>>>(ctx->cur.t, ctx->d_ss, ctx->d_cell_has_ghost, ctx->d_cell_hmax);
    PSPH_LAUNCH_CHECK();
    k_row_flags<<<(n_cap + 255) / 256, 256, 0, s>>>(ctx->cur, G, ctx->kp, ctx->d_ss, n_cap);
    PSPH_LAUNCH_CHECK();
    if (select_rows(ctx, n_cap, F_NEED, ctx->cur.need_idx, &ctx->d_ss->n_need)) return -1;
    if (ctx->has_grad && select_rows(ctx, n_cap, F_GRADROW, ctx->cur.grad_idx, &ctx->d_ss->n_grad)) return -1;
  }
  if (psph_pair_mode(ctx)) return -1;
  ctx->list_valid = true;
  ctx->nbr_valid = false;
  ctx->n_dirty = true;
  return 0;
}

// ---------------------------------------------------------------------------------------
// update_particles_in_bucket (update_pos.cc:172-462), single-rank semantics: every real or
// erupt-ghost row of an active, non-brief bucket is tested against the bounds of the bucket
// it is in; it moves to the adjacent bucket (one step per dimension, :257-271), is lost when
// it escapes a MIXED bucket through the ground (bucket.cc:434-452), leaves the mesh or
// enters a bucket that is not involved (:300-304).  Lost rows get cell = -1 and are
// dropped by the next list build.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ bool dev_escape(const KParams &K, const psph_cell_attr &a, const double *pos)
{
  for (int i = 0; i < K.dim; i++) {
    if (a.bucket_index[i * 2] == -1) {
      if ((pos[i] - a.bnd[i * 2]) < -K.tiny) return true;
    } else if (a.bucket_index[i * 2 + 1] == 1) {
      if ((pos[i] - a.bnd[i * 2 + 1]) > K.tiny) return true;
    }
  }
  return false;
}

__global__ void k_rebin(Tab V, GridView G, KParams K, StepState *ss)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ss->n) return;
  int bc = V.i[I_BC][p];
  if (!(dev_is_real(bc) || bc == 0)) return;
  int cell = V.i[I_CELL][p];
  if (cell < 0) return;
  const psph_cell_attr *a = G.attr ? &G.attr[cell] : nullptr;
  if (a && (a->type == PSPH_CELL_BRIEF || !a->active)) return;
  if (a && a->type == PSPH_CELL_PRESS_BC) return; // second loop keeps the old plist (:405-447)
  double pos[3] = { V.d[D_X][p], V.d[D_Y][p], V.d[D_Z][p] };
  bool lost = false;
  if (dev_is_real(bc) && a && a->type == PSPH_CELL_MIXED && dev_escape(K, *a, pos)) lost = true;
  int idx[3] = { cell / (G.ncell[2] * G.ncell[1]), (cell / G.ncell[2]) % G.ncell[1], cell % G.ncell[2] };
  int nidx[3];
  bool left = false;
  if (!lost) {
    for (int k = 0; k < 3; k++) {
      nidx[k] = idx[k];
      if (pos[k] < cell_min_d(G, k, idx[k])) { nidx[k] = idx[k] - 1; left = true; }
      else if (pos[k] >= cell_max_d(G, k, idx[k])) { nidx[k] = idx[k] + 1; left = true; }
    }
    if (left) {
      bool inside = true;
      for (int k = 0; k < 3; k++) if (nidx[k] < 0 || nidx[k] >= G.ncell[k]) inside = false;
      if (!inside) lost = true;
      else {
        int nl = (nidx[0] * G.ncell[1] + nidx[1]) * G.ncell[2] + nidx[2];
        if (G.attr && !G.attr[nl].has_involved) lost = true;
        else V.i[I_CELL][p] = nl;
      }
    }
  }
  if (lost) {
    V.i[I_CELL][p] = -1;
    atomicAdd(&ss->lost, 1);
  }
}

__global__ void k_reset_lost(StepState *ss) { ss->lost = 0; }

int psph_rebin_rows(psph_ctx *ctx)
{
  if (ctx->cfg.dim != 3) return 0; // the 1-D reference keeps no mesh (main.cc:676-751)
  cudaStream_t s = ctx->stream;
  k_reset_lost<<<1, 1, 0, s>>>(ctx->d_ss);
  PSPH_LAUNCH_CHECK();
  if (ctx->n > 0) {
    k_rebin<<<(ctx->n + 255) / 256, 256, 0, s>>>(ctx->cur.t, psph_grid_view(ctx), ctx->kp, ctx->d_ss);
    PSPH_LAUNCH_CHECK();
  }
  ctx->list_valid = false;
  return 0;
}

// ---------------------------------------------------------------------------------------
// test-only: the neighbour lists exactly as search_neighs_consth builds them
// (neighs.cc:37-211): all rows of the own bucket, plus the rows of the 26 surrounding
// buckets with |dx_k| <= 4.25*h_i for all k.  Device row indices; order inside a list is
// unspecified (the product never materialises lists — the gathers apply the same gate on
// the fly).
// ---------------------------------------------------------------------------------------
__global__ void k_emit_lists(DevView V, GridView G, KParams K, const StepState *ss, const long long *__restrict__ off,
                             int32_t *__restrict__ idx, int *__restrict__ cnt)
{
  const int lane = threadIdx.x & 31;
  const int n = ss->n;
  const int wglobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int i = wglobal; i < n; i += nw) {
    const int want = (K.use_gsph && K.dim == 3) ? F_GRADROW : F_NEED;
    int total = 0;
    if (V.flags[i] & want) {
      const double4 gi = V.geo[i];
      const double rl = K.neigh_radius * V.t.d[D_HLIST][i];
      long long w = off ? off[i] : 0;
      if (K.dim == 1) {
        for (int j0 = 0; j0 < n; j0 += 32) {
          int j = j0 + lane;
          bool ok = j < n;
          unsigned m = __ballot_sync(0xffffffffu, ok);
          if (ok && idx) idx[w + total + __popc(m & ((1u << lane) - 1))] = j;
          total += __popc(m);
        }
      } else {
        const int cell = V.t.i[I_CELL][i];
        const int ck = cell % G.ncell[2], cj = (cell / G.ncell[2]) % G.ncell[1], ci = cell / (G.ncell[2] * G.ncell[1]);
        for (int nb = 0; nb < 27; nb++) {
          const int ni = ci + nb / 9 - 1, nj = cj + (nb / 3) % 3 - 1, nk = ck + nb % 3 - 1;
          if (ni < 0 || nj < 0 || nk < 0 || ni >= G.ncell[0] || nj >= G.ncell[1] || nk >= G.ncell[2]) continue;
          const int nc = (ni * G.ncell[1] + nj) * G.ncell[2] + nk;
          const int r = G.cell2rank[nc];
          const int b = G.sub_start[r * SUBS], e = G.sub_start[r * SUBS + SUBS];
          for (int j0 = b; j0 < e; j0 += 32) {
            int j = j0 + lane;
            bool ok = j < e;
            if (ok && nc != cell) {
              const double4 gj = V.geo[j];
              ok = !(fabs(gi.x - gj.x) > rl || fabs(gi.y - gj.y) > rl || fabs(gi.z - gj.z) > rl);
            }
            unsigned m = __ballot_sync(0xffffffffu, ok);
            if (ok && idx) idx[w + total + __popc(m & ((1u << lane) - 1))] = j;
            total += __popc(m);
          }
        }
      }
    }
    if (lane == 0 && cnt) cnt[i] = total;
  }
}

extern "C" int psph_emit_neighbours(psph_ctx *ctx, int64_t *offsets, int32_t *idx, size_t idx_cap, size_t *nnz)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  if (!ctx->list_valid && psph_build_cell_list(ctx)) return -1;
  PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
  PSPH_CUDA(cudaMemcpy(ctx->h_ss, ctx->d_ss, sizeof(StepState), cudaMemcpyDeviceToHost));
  const int n = ctx->h_ss->n;
  ctx->n = n;
  if (!offsets) return -2;
  int *d_cnt = nullptr; long long *d_off = nullptr; int32_t *d_idx = nullptr;
  PSPH_CUDA(cudaMalloc(&d_cnt, (n + 1) * sizeof(int)));
  GridView G = psph_grid_view(ctx);
  const int grid = ctx->sms * 8;
  k_emit_lists<<<grid, 256, 0, ctx->stream>>>(ctx->cur, G, ctx->kp, ctx->d_ss, nullptr, nullptr, d_cnt);
  PSPH_LAUNCH_CHECK();
  std::vector<int> cnt(n + 1, 0);
  PSPH_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
  offsets[0] = 0;
  for (int i = 0; i < n; i++) offsets[i + 1] = offsets[i] + cnt[i];
  if (nnz) *nnz = (size_t) offsets[n];
  int rc = 0;
  if (idx) {
    if (idx_cap < (size_t) offsets[n]) { ctx->err = "psph_emit_neighbours: idx buffer too small"; rc = -2; }
    else if (offsets[n] > 0) {
      PSPH_CUDA(cudaMalloc(&d_off, (n + 1) * sizeof(long long)));
      PSPH_CUDA(cudaMalloc(&d_idx, offsets[n] * sizeof(int32_t)));
      PSPH_CUDA(cudaMemcpyAsync(d_off, offsets, (n + 1) * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
      k_emit_lists<<<grid, 256, 0, ctx->stream>>>(ctx->cur, G, ctx->kp, ctx->d_ss, d_off, d_idx, nullptr);
      PSPH_LAUNCH_CHECK();
      PSPH_CUDA(cudaMemcpyAsync(idx, d_idx, offsets[n] * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
      PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
    }
  }
  cudaFree(d_cnt); cudaFree(d_off); cudaFree(d_idx);
  return rc;
}
