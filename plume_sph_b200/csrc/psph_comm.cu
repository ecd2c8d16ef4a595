// psph_comm.cu — key-range slabs across the GPUs of one box and the ghost ("guest") exchange.
//
// Replaces the reference's MPI layer for this path:
//   move_data        src/multiproc/move_data.cc:48-427   -> psph_exchange_ghosts
//   repartition      src/multiproc/repartition_BSFC.cc:53-469 (+ BSFC_*.cc) -> psph_repartition
//   MPI_Allreduce    src/plume/main.cc:227 (dt)            -> ncclAllReduce(min) in psph_timestep
// One process (context) per GPU; NCCL point-to-point groups on the context stream.  Cells in
// Hilbert-rank order are cut into nranks contiguous key ranges ("slabs"); a rank owns the rows of
// its cells and mirrors, read-only, the rows of every cell that touches one of them (the
// reference's guest buckets, move_data.cc:236-299).  Ghost rows keep the sender's relative order
// through the stable sort on both sides, so refreshing a few columns after a stage needs no
// index exchange: sender packs `rows of boundary cells towards q`, receiver unpacks into `guest
// rows that came from q`, both in row order.
//
// NCCL is resolved at run time (dlopen "libnccl.so.2": the copy the process already has — e.g.
// torch's — or the system one); single-rank contexts never touch it.
#include "psph_internal.cuh"
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <dlfcn.h>
#include <nccl.h>
#include <algorithm>
#include <cstring>

namespace {

struct NcclApi
{
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string &err)
  {
    if (lib) return true;
    const char *names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char *nm : names) { lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) { err = std::string("NCCL not found: ") + dlerror(); return false; }
#define PSPH_SYM(field, name) *(void **) (&field) = dlsym(lib, name); if (!field) { err = std::string("NCCL symbol missing: ") + name; return false; }
    PSPH_SYM(GetUniqueId, "ncclGetUniqueId") PSPH_SYM(CommInitRank, "ncclCommInitRank") PSPH_SYM(CommDestroy, "ncclCommDestroy")
    PSPH_SYM(GroupStart, "ncclGroupStart") PSPH_SYM(GroupEnd, "ncclGroupEnd") PSPH_SYM(Send, "ncclSend") PSPH_SYM(Recv, "ncclRecv")
    PSPH_SYM(AllReduce, "ncclAllReduce") PSPH_SYM(AllGather, "ncclAllGather") PSPH_SYM(GetErrorString, "ncclGetErrorString")
#undef PSPH_SYM
    return true;
  }
};
NcclApi g_nccl;

struct ColList { int nd, ni; int d[D_NCOL]; int i[I_NCOL]; };

struct Comm
{
  ncclComm_t comm = nullptr;
  int P = 1, me = 0;
  std::vector<int> cuts;                 // [P+1] cell-rank boundaries of the slabs
  int32_t *d_owner = nullptr;            // [ncells] owner rank of every cell
  uint32_t *d_cellmask = nullptr;        // [ncells] ranks (bit q) that own a cell touching this (owned) cell
  int32_t *d_dest = nullptr;             // [cap] per-row scratch (destination rank / select key)
  int32_t *d_send_idx = nullptr, *d_recv_idx = nullptr; size_t idx_cap = 0;   // idx_cap: rows d_dest holds
  size_t send_idx_cap = 0, recv_idx_cap = 0;  // entries; a boundary row is listed once per peer that mirrors it
  std::vector<int> send_cnt, recv_cnt;   // per peer, of the current ghost layout
  bool lists_stale = true;               // idx lists do not match the current row order
  char *d_sendbuf = nullptr, *d_recvbuf = nullptr; size_t sendbuf_cap = 0, recvbuf_cap = 0;
  int *d_counts = nullptr, *d_allcounts = nullptr, *h_allcounts = nullptr; // [P], [P*P], pinned [P*P]
  unsigned long long *d_weights = nullptr; // [ncells] per-cell load weight, rank order
  int *d_cuts = nullptr;
  unsigned dirty = 0;
  long long bytes_sent = 0;
};

#define PSPH_NCCL(call)                                                                   \
  do {                                                                                    \
    ncclResult_t r_ = (call);                                                             \
    if (r_ != ncclSuccess) {                                                              \
      char b_[512];                                                                       \
      snprintf(b_, sizeof b_, "%s:%d %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
      ctx->err = b_;                                                                      \
      return -1;                                                                          \
    }                                                                                     \
  } while (0)

Comm *comm_of(psph_ctx *ctx) { return (Comm *) ctx->comm; }

// ---------------------------------------------------------------------------------------
// A bucket of the mesh's bottom layer (k == 0: underground) always goes with the bucket above it: the reference sorts it
// by the key of that bucket (PartiHead top_key, preprocess.cc:219-239, :264-275), so the two never part company.
__device__ __forceinline__ long long partition_cell(long long c, int nz) { return (nz > 1 && c % nz == 0) ? c + 1 : c; }

__global__ void k_owner(const int32_t *__restrict__ cell2rank, long long ncells, int nz, const int *__restrict__ cuts, int P,
                        int32_t *__restrict__ owner)
{
  long long c = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  const int r = cell2rank[partition_cell(c, nz)];
  int q = 0;
  while (q + 1 < P && r >= cuts[q + 1]) q++;
  owner[c] = q;
}

// which other ranks own a cell that touches cell c (27-neighbourhood, bucket.cc:23-26)
__global__ void k_cellmask(GridView G, const int32_t *__restrict__ owner, int me, uint32_t *__restrict__ mask)
{
  long long c = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= G.ncells) return;
  uint32_t m = 0;
  if (owner[c] == me) {
    const int ck = (int) (c % G.ncell[2]), cj = (int) ((c / G.ncell[2]) % G.ncell[1]), ci = (int) (c / ((long long) G.ncell[2] * G.ncell[1]));
    for (int nb = 0; nb < 27; nb++) {
      const int ni = ci + nb / 9 - 1, nj = cj + (nb / 3) % 3 - 1, nk = ck + nb % 3 - 1;
      if (ni < 0 || nj < 0 || nk < 0 || ni >= G.ncell[0] || nj >= G.ncell[1] || nk >= G.ncell[2]) continue;
      const int q = owner[((long long) ni * G.ncell[1] + nj) * G.ncell[2] + nk];
      if (q != me) m |= 1u << q;
    }
  }
  mask[c] = m;
}

// per-row destination of the migration: owner of the row's cell when that is another rank
__global__ void k_migration_dest(Tab V, const StepState *ss, const int32_t *__restrict__ owner, int me, int32_t *__restrict__ dest,
                                 int *__restrict__ counts)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ss->n) return;
  int d = -1;
  const int cell = V.i[I_CELL][p];
  if (cell >= 0 && !V.i[I_GUEST][p]) {
    const int q = owner[cell];
    if (q != me) { d = q; atomicAdd(&counts[q], 1); }
  }
  dest[p] = d;
}

__global__ void k_kill_rows(Tab V, int n, const int32_t *__restrict__ dest, int guests)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  if (guests ? V.i[I_GUEST][p] != 0 : dest[p] >= 0) V.i[I_CELL][p] = -1;
}

struct DestIs
{
  const int32_t *dest; int q;
  __device__ __forceinline__ bool operator()(int p) const { return dest[p] == q; }
};
// owned live rows of cells that touch a cell of rank q
struct BoundaryTo
{
  const int32_t *cell, *guest; const uint32_t *mask; int q;
  __device__ __forceinline__ bool operator()(int p) const
  {
    const int c = cell[p];
    return c >= 0 && !guest[p] && ((mask[c] >> q) & 1u);
  }
};
struct GuestFrom
{
  const int32_t *cell, *guest, *myproc; int q;
  __device__ __forceinline__ bool operator()(int p) const { return cell[p] >= 0 && guest[p] && myproc[p] == q; }
};

__global__ void k_pack_rows(Tab V, const int32_t *__restrict__ idx, int cnt, ColList cl, char *__restrict__ buf)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= cnt) return;
  const int r = idx[t];
  double *bd = (double *) buf;
  for (int c = 0; c < cl.nd; c++) bd[(size_t) c * cnt + t] = V.d[cl.d[c]][r];
  int32_t *bi = (int32_t *) (buf + (size_t) cl.nd * cnt * sizeof(double));
  for (int c = 0; c < cl.ni; c++) bi[(size_t) c * cnt + t] = V.i[cl.i[c]][r];
}

// idx == nullptr: rows base .. base+cnt-1 (append); set_guest >= 0: mark the rows guest of that rank,
// set_guest == -1: mark them owned by `me`; set_guest == -2: leave flags alone (column refresh)
__global__ void k_unpack_rows(Tab V, const int32_t *__restrict__ idx, int base, int cnt, ColList cl, const char *__restrict__ buf,
                              int set_guest, int me)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= cnt) return;
  const int r = idx ? idx[t] : base + t;
  const double *bd = (const double *) buf;
  for (int c = 0; c < cl.nd; c++) V.d[cl.d[c]][r] = bd[(size_t) c * cnt + t];
  const int32_t *bi = (const int32_t *) (buf + (size_t) cl.nd * cnt * sizeof(double));
  for (int c = 0; c < cl.ni; c++) V.i[cl.i[c]][r] = bi[(size_t) c * cnt + t];
  if (set_guest >= 0) { V.i[I_GUEST][r] = 1; V.i[I_MYPROC][r] = set_guest; }
  else if (set_guest == -1) { V.i[I_GUEST][r] = 0; V.i[I_MYPROC][r] = me; }
}

__global__ void k_set_n(StepState *ss, int n) { ss->n = n; }

// load weight of every cell in rank order, x100 so that the sums are exact integers
// (real 2.0, wall ghost 0.75, erupt ghost 0.02, pressure ghost 0: repartition_BSFC.cc:128-143, parameters.h:1035-1038)
__global__ void k_cell_weights(Tab V, const StepState *ss, const int32_t *__restrict__ cell2rank, int nz, unsigned long long *__restrict__ w)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ss->n) return;
  int cell = V.i[I_CELL][p];
  if (cell >= 0) cell = (int) partition_cell(cell, nz);           // an underground bucket's load counts where it goes
  if (cell < 0 || V.i[I_GUEST][p]) return;
  const int bc = V.i[I_BC][p];
  const unsigned long long v = (bc == 100 || bc == 99) ? 200ull : bc == 2 ? 75ull : bc == 0 ? 2ull : 0ull;
  if (v) atomicAdd(&w[cell2rank[cell]], v);
}

// cuts[p] = first cell rank whose inclusive weight prefix reaches p/P of the total (cuts[0] = 0, cuts[P] = ncells)
__global__ void k_find_cuts(const unsigned long long *__restrict__ prefix, long long ncells, int P, int *__restrict__ cuts)
{
  const int p = threadIdx.x;
  if (p > P) return;
  if (p == 0) { cuts[0] = 0; return; }
  if (p == P) { cuts[P] = (int) ncells; return; }
  const unsigned long long total = prefix[ncells - 1];
  // smallest r with prefix[r] * P >= total * p  ->  cut after the cells before r+? : cells [0, r] carry >= p/P
  long long lo = 0, hi = ncells;
  while (lo < hi) { long long mid = (lo + hi) >> 1; if (prefix[mid] * (unsigned long long) P < total * (unsigned long long) p) lo = mid + 1; else hi = mid; }
  cuts[p] = (int) min(lo + 1, ncells);
}

ColList all_columns(const psph_ctx *ctx)
{
  ColList cl;
  cl.nd = ctx->has_grad ? D_NCOL : D_GRAD0;
  for (int c = 0; c < cl.nd; c++) cl.d[c] = c;
  cl.ni = I_NCOL;
  for (int c = 0; c < I_NCOL; c++) cl.i[c] = c;
  return cl;
}
size_t row_bytes(const ColList &cl) { return (size_t) cl.nd * sizeof(double) + (size_t) cl.ni * sizeof(int32_t); }
// bytes a peer's segment occupies in the send/recv buffer: the payload (cnt * rb) padded so that the next
// peer's double columns start 16-byte aligned
size_t seg_bytes(int cnt, size_t rb) { return ((size_t) cnt * rb + 15) & ~(size_t) 15; }

template <class T>
int ensure_bytes(psph_ctx *ctx, T *&p, size_t &cap, size_t need)
{
  if (need <= cap) return 0;
  PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(p);
  size_t c = need + need / 4 + 4096;
  PSPH_CUDA(cudaMalloc((void **) &p, c));
  cap = c;
  return 0;
}

int ensure_row_scratch(psph_ctx *ctx)
{
  Comm *C = comm_of(ctx);
  if (C->idx_cap >= ctx->cap) return 0;
  PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
  int32_t *nd = nullptr; // the per-row destinations survive a growth (migrate() computes them before it knows the arrivals)
  PSPH_CUDA(cudaMalloc(&nd, ctx->cap * sizeof(int32_t)));
  if (C->d_dest && C->idx_cap) PSPH_CUDA(cudaMemcpy(nd, C->d_dest, C->idx_cap * sizeof(int32_t), cudaMemcpyDeviceToDevice));
  cudaFree(C->d_dest);
  C->d_dest = nd;
  C->idx_cap = ctx->cap;
  return 0;
}

// room for `ns` send-list and `nr` receive-list entries.  One peer's selection never exceeds the row count, but the
// lists of ALL peers are kept back to back (ensure_lists) and a boundary row appears once for every peer whose slab
// touches its cell: the sum over peers can exceed the table size.
int ensure_idx_lists(psph_ctx *ctx, size_t ns, size_t nr)
{
  Comm *C = comm_of(ctx);
  if (ns <= C->send_idx_cap && nr <= C->recv_idx_cap) return 0;
  PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ns > C->send_idx_cap) {
    cudaFree(C->d_send_idx); C->d_send_idx = nullptr;
    C->send_idx_cap = ns + ns / 4 + 1024;
    PSPH_CUDA(cudaMalloc(&C->d_send_idx, C->send_idx_cap * sizeof(int32_t)));
  }
  if (nr > C->recv_idx_cap) {
    cudaFree(C->d_recv_idx); C->d_recv_idx = nullptr;
    C->recv_idx_cap = nr + nr / 4 + 1024;
    PSPH_CUDA(cudaMalloc(&C->d_recv_idx, C->recv_idx_cap * sizeof(int32_t)));
  }
  C->lists_stale = true;
  return 0;
}

template <class Pred>
int select_into(psph_ctx *ctx, int n, Pred pred, int32_t *out, int *d_count)
{
  cub::CountingInputIterator<int> it(0);
  size_t need = 0;
  cub::DeviceSelect::If(nullptr, need, it, out, d_count, n, pred, ctx->stream);
  if (need > ctx->cub_tmp_bytes) {
    PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_cub_tmp);
    PSPH_CUDA(cudaMalloc(&ctx->d_cub_tmp, need));
    ctx->cub_tmp_bytes = need;
  }
  cub::DeviceSelect::If(ctx->d_cub_tmp, need, it, out, d_count, n, pred, ctx->stream);
  ctx->launches += 2;
  PSPH_CUDA(cudaGetLastError());
  return 0;
}

// every rank learns every rank's per-peer counts: all[src * P + dst]; also refreshes the host row count
int share_counts(psph_ctx *ctx)
{
  Comm *C = comm_of(ctx);
  PSPH_NCCL(g_nccl.AllGather(C->d_counts, C->d_allcounts, C->P, ncclInt32, C->comm, ctx->stream));
  PSPH_CUDA(cudaMemcpyAsync(C->h_allcounts, C->d_allcounts, (size_t) C->P * C->P * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PSPH_CUDA(cudaMemcpyAsync(ctx->h_ss, ctx->d_ss, sizeof(StepState), cudaMemcpyDeviceToHost, ctx->stream));
  PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->n = ctx->h_ss->n;
  return 0;
}

int update_owner_tables(psph_ctx *ctx)
{
  Comm *C = comm_of(ctx);
  PSPH_CUDA(cudaMemcpyAsync(C->d_cuts, C->cuts.data(), (C->P + 1) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  const unsigned grid = (unsigned) ((ctx->ncells + 255) / 256);
  k_owner<<<grid, 256, 0, ctx->stream>>>(ctx->d_cell2rank, ctx->ncells, ctx->cfg.ncell[2], C->d_cuts, C->P, C->d_owner);
  PSPH_LAUNCH_CHECK();
  k_cellmask<<<grid, 256, 0, ctx->stream>>>(psph_grid_view(ctx), C->d_owner, C->me, C->d_cellmask);
  PSPH_LAUNCH_CHECK();
  PSPH_CUDA(cudaStreamSynchronize(ctx->stream)); // cuts is host memory that may change
  return 0;
}

// send `send_cnt[q]` packed rows to every peer and receive `recv_cnt[q]`; buffers are laid out peer after peer
int exchange_buffers(psph_ctx *ctx, const std::vector<int> &send_cnt, const std::vector<int> &recv_cnt, size_t rb)
{
  Comm *C = comm_of(ctx);
  PSPH_NCCL(g_nccl.GroupStart());
  size_t so = 0, ro = 0;
  for (int q = 0; q < C->P; q++) {
    if (q == C->me) continue;
    if (send_cnt[q]) PSPH_NCCL(g_nccl.Send(C->d_sendbuf + so, (size_t) send_cnt[q] * rb, ncclInt8, q, C->comm, ctx->stream));
    if (recv_cnt[q]) PSPH_NCCL(g_nccl.Recv(C->d_recvbuf + ro, (size_t) recv_cnt[q] * rb, ncclInt8, q, C->comm, ctx->stream));
    so += seg_bytes(send_cnt[q], rb); ro += seg_bytes(recv_cnt[q], rb);
    C->bytes_sent += (long long) send_cnt[q] * (long long) rb;
  }
  PSPH_NCCL(g_nccl.GroupEnd());
  return 0;
}

// move the rows selected per peer by `make_pred(q)` (counts already in d_counts) as complete rows; received
// rows are appended.  as_guest: arrivals become guests of the sender, else owned rows.
template <class MakePred>
int ship_rows(psph_ctx *ctx, MakePred make_pred, bool as_guest, std::vector<int> *keep_send, std::vector<int> *keep_recv,
              int *n_before = nullptr)
{
  Comm *C = comm_of(ctx);
  const int P = C->P, me = C->me;
  if (share_counts(ctx)) return -1;
  std::vector<int> sc(P, 0), rc(P, 0);
  long long rtot = 0;
  size_t sbytes = 0, rbytes = 0;
  const ColList cl = all_columns(ctx);
  const size_t rb = row_bytes(cl);
  for (int q = 0; q < P; q++) {
    if (q == me) continue;
    sc[q] = C->h_allcounts[me * P + q]; rc[q] = C->h_allcounts[q * P + me];
    rtot += rc[q];
    sbytes += seg_bytes(sc[q], rb); rbytes += seg_bytes(rc[q], rb);
  }
  const int n0 = ctx->n;
  if (n_before) *n_before = n0;
  if (psph_reserve(ctx, (size_t) n0 + rtot)) return -1;
  if (ensure_row_scratch(ctx)) return -1;
  {
    int smax = 0;
    for (int q = 0; q < P; q++) smax = std::max(smax, sc[q]);
    if (ensure_idx_lists(ctx, (size_t) smax, 0)) return -1;      // one peer's selection at a time
  }
  if (ensure_bytes(ctx, C->d_sendbuf, C->sendbuf_cap, sbytes)) return -1;
  if (ensure_bytes(ctx, C->d_recvbuf, C->recvbuf_cap, rbytes)) return -1;
  size_t so = 0;
  for (int q = 0; q < P; q++) {
    if (q == me || !sc[q]) continue;
    if (select_into(ctx, n0, make_pred(q), C->d_send_idx, &C->d_counts[q])) return -1;
    k_pack_rows<<<(sc[q] + 255) / 256, 256, 0, ctx->stream>>>(ctx->cur.t, C->d_send_idx, sc[q], cl, C->d_sendbuf + so);
    PSPH_LAUNCH_CHECK();
    so += seg_bytes(sc[q], rb);
  }
  if (exchange_buffers(ctx, sc, rc, rb)) return -1;
  size_t ro = 0; int base = n0;
  for (int q = 0; q < P; q++) {
    if (q == me || !rc[q]) continue;
    k_unpack_rows<<<(rc[q] + 255) / 256, 256, 0, ctx->stream>>>(ctx->cur.t, nullptr, base, rc[q], cl, C->d_recvbuf + ro, as_guest ? q : -1, me);
    PSPH_LAUNCH_CHECK();
    ro += seg_bytes(rc[q], rb); base += rc[q];
  }
  ctx->n = base;
  k_set_n<<<1, 1, 0, ctx->stream>>>(ctx->d_ss, base);
  PSPH_LAUNCH_CHECK();
  ctx->list_valid = false;
  if (keep_send) *keep_send = sc;
  if (keep_recv) *keep_recv = rc;
  return 0;
}

// rows whose cell now belongs to another rank go there (replaces cases 1-3 of update_pos.cc:273-381
// and BSFC_update_and_send_elements.cc:131-850)
int migrate(psph_ctx *ctx)
{
  Comm *C = comm_of(ctx);
  if (ensure_row_scratch(ctx)) return -1;
  PSPH_CUDA(cudaMemsetAsync(C->d_counts, 0, C->P * sizeof(int), ctx->stream));
  if (ctx->n > 0) {
    k_migration_dest<<<(ctx->n + 255) / 256, 256, 0, ctx->stream>>>(ctx->cur.t, ctx->d_ss, C->d_owner, C->me, C->d_dest, C->d_counts);
    PSPH_LAUNCH_CHECK();
  }
  int n_before = 0;
  if (ship_rows(ctx, [C](int q) { DestIs p = { C->d_dest, q }; return p; }, false, nullptr, nullptr, &n_before)) return -1;
  if (n_before > 0) { // the rows that left (only rows < n_before carry a destination)
    k_kill_rows<<<(n_before + 255) / 256, 256, 0, ctx->stream>>>(ctx->cur.t, n_before, C->d_dest, 0);
    PSPH_LAUNCH_CHECK();
  }
  return 0;
}

__global__ void k_count_boundary(Tab V, const StepState *ss, const uint32_t *__restrict__ mask, int P, int *__restrict__ counts)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ss->n) return;
  const int c = V.i[I_CELL][p];
  if (c < 0 || V.i[I_GUEST][p]) return;
  uint32_t m = mask[c];
  while (m) { const int q = __ffs(m) - 1; m &= m - 1; atomicAdd(&counts[q], 1); }
}

// drop every guest row and mirror the boundary cells again (move_data.cc:150-170,236-398)
int rebuild_ghosts(psph_ctx *ctx)
{
  Comm *C = comm_of(ctx);
  if (ensure_row_scratch(ctx)) return -1;
  PSPH_CUDA(cudaMemsetAsync(C->d_counts, 0, C->P * sizeof(int), ctx->stream));
  if (ctx->n > 0) {
    k_kill_rows<<<(ctx->n + 255) / 256, 256, 0, ctx->stream>>>(ctx->cur.t, ctx->n, C->d_dest, 1);
    PSPH_LAUNCH_CHECK();
    k_count_boundary<<<(ctx->n + 255) / 256, 256, 0, ctx->stream>>>(ctx->cur.t, ctx->d_ss, C->d_cellmask, C->P, C->d_counts);
    PSPH_LAUNCH_CHECK();
  }
  auto pred = [ctx, C](int q) { BoundaryTo p = { ctx->cur.t.i[I_CELL], ctx->cur.t.i[I_GUEST], C->d_cellmask, q }; return p; };
  if (ship_rows(ctx, pred, true, &C->send_cnt, &C->recv_cnt)) return -1;
  C->lists_stale = true;
  return 0;
}

// send/recv row lists of the current row order (counts are those of the last ghost rebuild)
int ensure_lists(psph_ctx *ctx)
{
  Comm *C = comm_of(ctx);
  if (!C->lists_stale) return 0;
  if (psph_sync_n(ctx)) return -1; // builds the cell list if needed; exact row count on the host
  if (ensure_row_scratch(ctx)) return -1;
  {
    size_t ns = 0, nr = 0;
    for (int q = 0; q < C->P; q++) { if (q == C->me) continue; ns += (size_t) C->send_cnt[q]; nr += (size_t) C->recv_cnt[q]; }
    if (ensure_idx_lists(ctx, ns, nr)) return -1;
  }
  const int32_t *cell = ctx->cur.t.i[I_CELL], *guest = ctx->cur.t.i[I_GUEST], *myproc = ctx->cur.t.i[I_MYPROC];
  size_t so = 0, ro = 0;
  for (int q = 0; q < C->P; q++) {
    if (q == C->me) continue;
    if (C->send_cnt[q]) {
      BoundaryTo p = { cell, guest, C->d_cellmask, q };
      if (select_into(ctx, ctx->n, p, C->d_send_idx + so, &C->d_counts[q])) return -1;
      so += C->send_cnt[q];
    }
    if (C->recv_cnt[q]) {
      GuestFrom p = { cell, guest, myproc, q };
      if (select_into(ctx, ctx->n, p, C->d_recv_idx + ro, &C->d_counts[q])) return -1;
      ro += C->recv_cnt[q];
    }
  }
  C->lists_stale = false;
  return 0;
}

// refresh the given columns of every guest row from its owner
int refresh_columns(psph_ctx *ctx, const ColList &cl)
{
  Comm *C = comm_of(ctx);
  if (ensure_lists(ctx)) return -1;
  const size_t rb = row_bytes(cl);
  size_t sbytes = 0, rbytes = 0;
  for (int q = 0; q < C->P; q++) { sbytes += seg_bytes(C->send_cnt[q], rb); rbytes += seg_bytes(C->recv_cnt[q], rb); }
  if (ensure_bytes(ctx, C->d_sendbuf, C->sendbuf_cap, sbytes)) return -1;
  if (ensure_bytes(ctx, C->d_recvbuf, C->recvbuf_cap, rbytes)) return -1;
  size_t so = 0, si = 0;
  for (int q = 0; q < C->P; q++) {
    const int c = C->send_cnt[q];
    if (q == C->me || !c) continue;
    k_pack_rows<<<(c + 255) / 256, 256, 0, ctx->stream>>>(ctx->cur.t, C->d_send_idx + si, c, cl, C->d_sendbuf + so);
    PSPH_LAUNCH_CHECK();
    so += seg_bytes(c, rb); si += c;
  }
  if (exchange_buffers(ctx, C->send_cnt, C->recv_cnt, rb)) return -1;
  size_t ro = 0, ri = 0;
  for (int q = 0; q < C->P; q++) {
    const int c = C->recv_cnt[q];
    if (q == C->me || !c) continue;
    k_unpack_rows<<<(c + 255) / 256, 256, 0, ctx->stream>>>(ctx->cur.t, C->d_recv_idx + ri, 0, c, cl, C->d_recvbuf + ro, -2, C->me);
    PSPH_LAUNCH_CHECK();
    ro += seg_bytes(c, rb); ri += c;
  }
  return 0;
}

// ---- apply_bcond across slabs -------------------------------------------------------------------------------------
// image q is computed by the owner of its bucket (psph_bcond.cu); when the wall ghost it belongs to is another rank's row,
// the computing rank holds a guest copy and the state travels back to the owner.  Both sides enumerate the images in
// ascending q, so the payload needs no keys.
struct ImgToPeer     // computed here, ghost owned by rank r
{
  const int32_t *img_row, *img_cell, *guest, *myproc, *owner; int me, r;
  __device__ __forceinline__ bool operator()(int q) const
  {
    const int row = img_row[q];
    return row >= 0 && owner[img_cell[q]] == me && guest[row] && myproc[row] == r;
  }
};
struct ImgFromPeer   // ghost owned here, computed by rank r
{
  const int32_t *img_row, *img_cell, *guest, *owner; int r;
  __device__ __forceinline__ bool operator()(int q) const
  {
    const int row = img_row[q];
    return row >= 0 && !guest[row] && owner[img_cell[q]] == r;
  }
};
__global__ void k_img_counts(const int32_t *__restrict__ img_row, const int32_t *__restrict__ img_cell, const int32_t *__restrict__ guest,
                             const int32_t *__restrict__ myproc, const int32_t *__restrict__ owner, int me, int n_img, int *__restrict__ counts)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_img) return;
  const int row = img_row[q];
  if (row >= 0 && owner[img_cell[q]] == me && guest[row] && myproc[row] != me) atomicAdd(&counts[myproc[row]], 1);
}
__global__ void k_img_to_rows(int32_t *__restrict__ idx, int cnt, const int32_t *__restrict__ img_row)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < cnt) idx[t] = img_row[idx[t]];
}

} // namespace

int psph_comm_return_images(psph_ctx *ctx, const int32_t *img_row, const int32_t *img_cell, int n_img)
{
  Comm *C = comm_of(ctx);
  if (!C) return 0;
  const int P = C->P, me = C->me;
  const int32_t *guest = ctx->cur.t.i[I_GUEST], *myproc = ctx->cur.t.i[I_MYPROC];
  PSPH_CUDA(cudaMemsetAsync(C->d_counts, 0, P * sizeof(int), ctx->stream));
  k_img_counts<<<(n_img + 255) / 256, 256, 0, ctx->stream>>>(img_row, img_cell, guest, myproc, C->d_owner, me, n_img, C->d_counts);
  PSPH_LAUNCH_CHECK();
  if (share_counts(ctx)) return -1;
  std::vector<int> sc(P, 0), rc(P, 0);
  size_t ns = 0, nr = 0;
  for (int q = 0; q < P; q++) {
    if (q == me) continue;
    sc[q] = C->h_allcounts[me * P + q]; rc[q] = C->h_allcounts[q * P + me];
    ns += sc[q]; nr += rc[q];
  }
  bool any = false;
  for (int k = 0; k < P * P; k++) any = any || C->h_allcounts[k] != 0;
  if (!any) return 0;                                    // every rank sees the same table: all leave together
  ColList cl; cl.ni = 0; cl.nd = 0;
  const int cols[] = { D_RHO, D_VX, D_VY, D_VZ, D_E, D_KS, D_P, D_GAMMA, D_C, D_T, D_CP, D_RPH0, D_RPH1 };
  for (int c : cols) cl.d[cl.nd++] = c;
  const size_t rb = row_bytes(cl);
  if (ensure_idx_lists(ctx, ns, nr)) return -1;
  C->lists_stale = true;                                 // the ghost-refresh row lists shared these buffers
  size_t sbytes = 0, rbytes = 0;
  for (int q = 0; q < P; q++) { sbytes += seg_bytes(sc[q], rb); rbytes += seg_bytes(rc[q], rb); }
  if (ensure_bytes(ctx, C->d_sendbuf, C->sendbuf_cap, sbytes)) return -1;
  if (ensure_bytes(ctx, C->d_recvbuf, C->recvbuf_cap, rbytes)) return -1;
  size_t so = 0, si = 0, ri = 0;
  for (int q = 0; q < P; q++) {
    if (q == me) continue;
    if (sc[q]) {
      ImgToPeer p = { img_row, img_cell, guest, myproc, C->d_owner, me, q };
      if (select_into(ctx, n_img, p, C->d_send_idx + si, &C->d_counts[q])) return -1;
      k_img_to_rows<<<(sc[q] + 255) / 256, 256, 0, ctx->stream>>>(C->d_send_idx + si, sc[q], img_row);
      PSPH_LAUNCH_CHECK();
      k_pack_rows<<<(sc[q] + 255) / 256, 256, 0, ctx->stream>>>(ctx->cur.t, C->d_send_idx + si, sc[q], cl, C->d_sendbuf + so);
      PSPH_LAUNCH_CHECK();
      so += seg_bytes(sc[q], rb); si += sc[q];
    }
    if (rc[q]) {
      ImgFromPeer p = { img_row, img_cell, guest, C->d_owner, q };
      if (select_into(ctx, n_img, p, C->d_recv_idx + ri, &C->d_counts[q])) return -1;
      k_img_to_rows<<<(rc[q] + 255) / 256, 256, 0, ctx->stream>>>(C->d_recv_idx + ri, rc[q], img_row);
      PSPH_LAUNCH_CHECK();
      ri += rc[q];
    }
  }
  if (exchange_buffers(ctx, sc, rc, rb)) return -1;
  size_t ro = 0; ri = 0;
  for (int q = 0; q < P; q++) {
    if (q == me || !rc[q]) continue;
    k_unpack_rows<<<(rc[q] + 255) / 256, 256, 0, ctx->stream>>>(ctx->cur.t, C->d_recv_idx + ri, 0, rc[q], cl, C->d_recvbuf + ro, -2, me);
    PSPH_LAUNCH_CHECK();
    ro += seg_bytes(rc[q], rb); ri += rc[q];
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
void psph_comm_destroy(psph_ctx *ctx)
{
  Comm *C = comm_of(ctx);
  if (!C) return;
  if (C->comm) g_nccl.CommDestroy(C->comm);
  cudaFree(C->d_owner); cudaFree(C->d_cellmask); cudaFree(C->d_dest); cudaFree(C->d_send_idx); cudaFree(C->d_recv_idx);
  cudaFree(C->d_sendbuf); cudaFree(C->d_recvbuf); cudaFree(C->d_counts); cudaFree(C->d_allcounts); cudaFree(C->d_weights);
  cudaFree(C->d_cuts);
  cudaFreeHost(C->h_allcounts);
  delete C;
  ctx->comm = nullptr;
}

extern "C" int psph_comm_unique_id(void *out, size_t cap)
{
  std::string err;
  if (!out || cap < sizeof(ncclUniqueId) || !g_nccl.load(err)) return -1;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return -1;
  memcpy(out, &id, sizeof id);
  return (int) sizeof id;
}

extern "C" int psph_comm_init(psph_ctx *ctx, const void *nccl_unique_id, size_t id_bytes)
{
  if (ctx->cfg.nranks == 1) return 0;
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  if (ctx->cfg.nranks > 32) { ctx->err = "psph_comm_init: at most 32 ranks"; return -2; }
  if (!nccl_unique_id || id_bytes < sizeof(ncclUniqueId)) { ctx->err = "psph_comm_init: bad unique id"; return -2; }
  if (!g_nccl.load(ctx->err)) return -1;
  Comm *C = new Comm();
  ctx->comm = C;
  C->P = ctx->cfg.nranks; C->me = ctx->cfg.rank;
  ncclUniqueId id;
  memcpy(&id, nccl_unique_id, sizeof id);
  PSPH_NCCL(g_nccl.CommInitRank(&C->comm, C->P, id, C->me));
  PSPH_CUDA(cudaMalloc(&C->d_owner, ctx->ncells * sizeof(int32_t)));
  PSPH_CUDA(cudaMalloc(&C->d_cellmask, ctx->ncells * sizeof(uint32_t)));
  PSPH_CUDA(cudaMalloc(&C->d_weights, ctx->ncells * sizeof(unsigned long long)));
  PSPH_CUDA(cudaMalloc(&C->d_counts, C->P * sizeof(int)));
  PSPH_CUDA(cudaMalloc(&C->d_allcounts, (size_t) C->P * C->P * sizeof(int)));
  PSPH_CUDA(cudaMalloc(&C->d_cuts, (C->P + 1) * sizeof(int)));
  PSPH_CUDA(cudaMallocHost(&C->h_allcounts, (size_t) C->P * C->P * sizeof(int)));
  C->send_cnt.assign(C->P, 0); C->recv_cnt.assign(C->P, 0);
  // initial slabs: equal numbers of cells along the curve (preprocess.cc:249-279); psph_repartition balances by weight
  C->cuts.resize(C->P + 1);
  for (int p = 0; p <= C->P; p++) C->cuts[p] = (int) (ctx->ncells * p / C->P);
  C->dirty = DIRTY_LAYOUT;
  return update_owner_tables(ctx);
}

const int32_t *psph_comm_owner_table(psph_ctx *ctx) { return comm_of(ctx) ? comm_of(ctx)->d_owner : nullptr; }

// the stages tell the exchange what they changed
void psph_comm_mark(psph_ctx *ctx, unsigned what) { if (comm_of(ctx)) comm_of(ctx)->dirty |= what; }

extern "C" int psph_exchange_ghosts(psph_ctx *ctx)
{
  Comm *C = comm_of(ctx);
  if (ctx->cfg.nranks == 1) return 0; // move_data.cc:53-54
  if (!C) { ctx->err = "psph_exchange_ghosts: psph_comm_init was not called"; return -2; }
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  const unsigned dirty = C->dirty;
  C->dirty = 0;
  if (!dirty) return 0;
  StageTimer tm(ctx, ST_COMM);                // (timing pass only) migration + ghost layer / column refresh, host waits included
  if (dirty & DIRTY_LAYOUT) { // positions / membership / population changed: migrate, then mirror the boundary cells anew
    if (migrate(ctx)) return -1;
    return rebuild_ghosts(ctx);
  }
  if (!dirty) return 0;
  ColList cl; cl.nd = 0; cl.ni = 0;
  auto add = [&](int c) { for (int k = 0; k < cl.nd; k++) if (cl.d[k] == c) return; cl.d[cl.nd++] = c; };
  const int eos[] = { D_P, D_GAMMA, D_C, D_T, D_CP, D_RPH0, D_RPH1 };
  if (dirty & DIRTY_MOM) { add(D_VX); add(D_VY); add(D_VZ); add(D_E); for (int c : eos) add(c); }
  if (dirty & DIRTY_DENS) { add(D_RHO); add(D_KS); add(D_E); for (int c : eos) add(c); cl.i[cl.ni++] = I_INVOLVED; }
  if (dirty & DIRTY_VEL) { add(D_SVX); add(D_SVY); add(D_SVZ); }
  if (dirty & DIRTY_H) add(D_H);
  if (dirty & DIRTY_GRAD) for (int k = 0; k < 15; k++) add(D_GRAD0 + k);
  if (refresh_columns(ctx, cl)) return -1;
  if (dirty & DIRTY_H) return psph_refresh_geo(ctx);
  return 0;
}

extern "C" int psph_repartition(psph_ctx *ctx)
{
  Comm *C = comm_of(ctx);
  if (ctx->cfg.nranks == 1) return 0;
  if (!C) { ctx->err = "psph_repartition: psph_comm_init was not called"; return -2; }
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  PSPH_CUDA(cudaMemsetAsync(C->d_weights, 0, ctx->ncells * sizeof(unsigned long long), ctx->stream));
  if (ctx->n > 0) {
    k_cell_weights<<<(ctx->n + 255) / 256, 256, 0, ctx->stream>>>(ctx->cur.t, ctx->d_ss, ctx->d_cell2rank, ctx->cfg.ncell[2], C->d_weights);
    PSPH_LAUNCH_CHECK();
  }
  PSPH_NCCL(g_nccl.AllReduce(C->d_weights, C->d_weights, ctx->ncells, ncclUint64, ncclSum, C->comm, ctx->stream));
  size_t need = 0;
  cub::DeviceScan::InclusiveSum(nullptr, need, C->d_weights, C->d_weights, ctx->ncells, ctx->stream);
  if (need > ctx->cub_tmp_bytes) {
    PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_cub_tmp);
    PSPH_CUDA(cudaMalloc(&ctx->d_cub_tmp, need));
    ctx->cub_tmp_bytes = need;
  }
  cub::DeviceScan::InclusiveSum(ctx->d_cub_tmp, need, C->d_weights, C->d_weights, ctx->ncells, ctx->stream);
  k_find_cuts<<<1, 64, 0, ctx->stream>>>(C->d_weights, ctx->ncells, C->P, C->d_cuts);
  PSPH_LAUNCH_CHECK();
  PSPH_CUDA(cudaMemcpyAsync(C->cuts.data(), C->d_cuts, (C->P + 1) * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int p = 1; p <= C->P; p++) if (C->cuts[p] < C->cuts[p - 1]) C->cuts[p] = C->cuts[p - 1];
  if (update_owner_tables(ctx)) return -1;
  C->dirty = 0;
  if (migrate(ctx)) return -1;
  return rebuild_ghosts(ctx);
}

// dt = min over ranks (main.cc:224-231); ordered-bits minimum of positive doubles
int psph_comm_min_dt(psph_ctx *ctx)
{
  Comm *C = comm_of(ctx);
  if (!C) return 0;
  PSPH_NCCL(g_nccl.AllReduce(&ctx->d_ss->min_hc_bits, &ctx->d_ss->min_hc_bits, 1, ncclUint64, ncclMin, C->comm, ctx->stream));
  return 0;
}

// all-reduce of an int32 array over the slabs (op 0 sum, 1 max); no-op on one rank
int psph_comm_allreduce_i32(psph_ctx *ctx, int32_t *d, size_t n, int op)
{
  Comm *C = comm_of(ctx);
  if (!C || n == 0) return 0;
  PSPH_NCCL(g_nccl.AllReduce(d, d, n, ncclInt32, op ? ncclMax : ncclSum, C->comm, ctx->stream));
  return 0;
}

extern "C" int psph_slab_cuts(psph_ctx *ctx, int *cuts, size_t cap)
{
  Comm *C = comm_of(ctx);
  if (!C) { if (cap >= 2) { cuts[0] = 0; cuts[1] = (int) ctx->ncells; } return 1; }
  if (cap < (size_t) C->P + 1) return -2;
  for (int p = 0; p <= C->P; p++) cuts[p] = C->cuts[p];
  return C->P;
}
extern "C" int64_t psph_comm_bytes_sent(psph_ctx *ctx, int reset)
{
  Comm *C = comm_of(ctx);
  if (!C) return 0;
  long long v = C->bytes_sent;
  if (reset) C->bytes_sent = 0;
  return v;
}
