// psph_output.cu — the output path without the AoS round trip (output half).
//
// The reference's write_h5part (h5part_writer.cc:40-269) walks the particle table in THashTable order
// (thashtab.cc:204-222, :86-134, :246-262: ascending hash slot, chains sorted by key), keeps the non-guest real
// particles (every particle in a WRITE_GHOSTS build, :88-91) and writes one dataset per quantity:
//   ID (0 .. n-1), x, y, z, Rho, Vx, Vy, Vz, engr, mssfrc, prss (double), phase, bctp, guest (int).
// psph_download_h5part hands the driver exactly those arrays: the rows are selected and ordered on the device (four
// stable radix passes: key2, key1, key0, slot), the thirteen columns gathered into contiguous buffers and copied —
// 92 bytes per written particle instead of a 208 / 328-byte row for EVERY particle, ghosts included (3.5 ghosts per
// real particle in the default plume), and no host-side sort.
#include "psph_internal.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

namespace {

struct OutputRow   // the writer's filter (h5part_writer.cc:88-91)
{
  const int32_t *bc, *guest, *cell; int with_ghosts;
  __device__ __forceinline__ bool operator()(int p) const
  { return cell[p] >= 0 && (with_ghosts || (!guest[p] && bc[p] == 100)); }
};

// pass 0..2: key word 2, 1, 0 of the row; pass 3: the hash slot of the key (thashtab.cc:204-222)
__global__ void k_out_keys(Tab V, const uint32_t *__restrict__ perm, int n, int pass, int dim, int S01, int S02, uint32_t *__restrict__ key)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int p = (int) perm[t];
  uint32_t k;
  if (pass < 3) k = (uint32_t) V.i[I_KEY0 + (2 - pass)][p];
  else {
    const uint32_t k0 = (uint32_t) V.i[I_KEY0][p], k1 = (uint32_t) V.i[I_KEY1][p];
    const double umax = 4294967295.0;
    k = dim == 1 ? k0 : (uint32_t) ((int) __dmul_rn(__ddiv_rn((double) k0, umax), (double) S01) + (int) __dmul_rn(__ddiv_rn((double) k1, umax), (double) S02));
  }
  key[t] = k;
}

__global__ void k_out_gather(Tab V, const uint32_t *__restrict__ perm, int n, size_t cap, double *__restrict__ dout, int32_t *__restrict__ iout)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int p = (int) perm[t];
  const int dc[10] = { D_X, D_Y, D_Z, D_RHO, D_VX, D_VY, D_VZ, D_E, D_KS, D_P };
#pragma unroll
  for (int c = 0; c < 10; c++) dout[c * cap + t] = V.d[dc[c]][p];
  iout[t] = V.i[I_PHASE][p]; iout[cap + t] = V.i[I_BC][p]; iout[2 * cap + t] = V.i[I_GUEST][p];
}

} // namespace

static_assert(I_KEY1 == I_KEY0 + 1 && I_KEY2 == I_KEY0 + 2, "key words are consecutive columns");

extern "C" int psph_download_h5part(psph_ctx *ctx, double *const *dcols, int32_t *const *icols, size_t cap, size_t *n_out, int with_ghosts)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  if (psph_sync_n(ctx)) return -1;
  const int n = ctx->n;
  if (n_out) *n_out = 0;
  if (n == 0) return 0;
  cudaStream_t s = ctx->stream;
  // the sort scratch of the list build is free between steps: perm / perm_alt hold row indices, sortkey / sortkey_alt keys
  uint32_t *perm = ctx->d_perm, *perm_alt = ctx->d_perm_alt, *key = ctx->d_sortkey, *key_alt = ctx->d_sortkey_alt;
  int *d_count = &ctx->d_ss->work;
  OutputRow pred = { ctx->cur.t.i[I_BC], ctx->cur.t.i[I_GUEST], ctx->cur.t.i[I_CELL], with_ghosts };
  cub::CountingInputIterator<int> it(0);
  size_t need_sel = 0, need_sort = 0;
  cub::DeviceSelect::If(nullptr, need_sel, it, reinterpret_cast<int *>(perm), d_count, n, pred, s);
  cub::DeviceRadixSort::SortPairs(nullptr, need_sort, key, key_alt, perm, perm_alt, n, 0, 32, s);
  const size_t need = std::max(need_sel, need_sort);
  if (need > ctx->cub_tmp_bytes) {
    PSPH_CUDA(cudaStreamSynchronize(s));
    cudaFree(ctx->d_cub_tmp);
    PSPH_CUDA(cudaMalloc(&ctx->d_cub_tmp, need));
    ctx->cub_tmp_bytes = need;
  }
  size_t tmp = ctx->cub_tmp_bytes;
  cub::DeviceSelect::If(ctx->d_cub_tmp, tmp, it, reinterpret_cast<int *>(perm), d_count, n, pred, s);
  ctx->launches++;
  int m = 0;
  PSPH_CUDA(cudaMemcpyAsync(&m, d_count, sizeof(int), cudaMemcpyDeviceToHost, s));
  PSPH_CUDA(cudaStreamSynchronize(s));
  if (n_out) *n_out = (size_t) m;
  if (!dcols || !icols || m == 0) return 0;
  if (cap < (size_t) m) { ctx->err = "psph_download_h5part: buffers too small"; return -2; }
  const int S02 = ctx->cfg.ptable_size / 10, S01 = ctx->cfg.ptable_size - S02;
  for (int pass = 0; pass < 4; pass++) {
    k_out_keys<<<(m + 255) / 256, 256, 0, s>>>(ctx->cur.t, perm, m, pass, ctx->cfg.dim, S01, S02, key);
    PSPH_LAUNCH_CHECK();
    tmp = ctx->cub_tmp_bytes;
    cub::DeviceRadixSort::SortPairs(ctx->d_cub_tmp, tmp, key, key_alt, perm, perm_alt, m, 0, 32, s);
    ctx->launches++;
    PSPH_CUDA(cudaGetLastError());
    std::swap(perm, perm_alt);
  }
  // staging: 10 double + 3 int columns of m rows (the AoS staging buffer is at least that large once it exists)
  const size_t bytes = (size_t) m * (10 * sizeof(double) + 3 * sizeof(int32_t));
  if (psph_ensure_aos(ctx, (bytes + sizeof(psph_particle) - 1) / sizeof(psph_particle))) return -1;
  double *dout = reinterpret_cast<double *>(ctx->d_aos);
  int32_t *iout = reinterpret_cast<int32_t *>(dout + (size_t) 10 * m);
  k_out_gather<<<(m + 255) / 256, 256, 0, s>>>(ctx->cur.t, perm, m, (size_t) m, dout, iout);
  PSPH_LAUNCH_CHECK();
  for (int c = 0; c < 10; c++)
    if (dcols[c]) PSPH_CUDA(cudaMemcpyAsync(dcols[c], dout + (size_t) c * m, (size_t) m * sizeof(double), cudaMemcpyDeviceToHost, s));
  for (int c = 0; c < 3; c++)
    if (icols[c]) PSPH_CUDA(cudaMemcpyAsync(icols[c], iout + (size_t) c * m, (size_t) m * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  PSPH_CUDA(cudaStreamSynchronize(s));
  return 0;
}
