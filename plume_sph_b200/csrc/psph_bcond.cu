// psph_bcond.cu — wall boundary condition on the device.
//
// Replaces apply_bcond (src/plume/bcond.cc:24-249): every wall ghost (bc_type 2) that has an image — its mirror
// point across the wall, kept by the reference driver in its Image_table (BndImage, src/datstr/bnd_image.h) — takes
// the Shepard-interpolated state of the fluid at that point, with the velocity reflected about the wall normal.
// The reference walks the particle lists of the image's bucket and of its 26 neighbours; here one WARP takes one
// image and its lanes share out the subcell columns of those buckets that lie within kernel reach of the image
// point (contiguous row ranges of the sorted table).  Only real and
// pressure-ghost rows contribute (Particle::contr_image, particle.h:562-565), never wall ghosts, so the update is
// order independent.  Sums run in lane / shuffle order: fields agree with the reference to FP64 rounding.
#include "psph_gather.cuh"
#include <algorithm>

struct ImgKey { uint32_t k[3]; int32_t idx; };

// rows that carry an image: binary search of the row's birth key among the sorted image keys.  img_row[q] = the local
// row (owned or guest) of image q's wall ghost; `list` collects the images this rank computes: all of them on a single
// rank, with slabs those whose image point lies in a bucket this rank owns — the rank that holds every bucket around it
// (the reference computes an image on the owner of its bucket too: buckproc, bcond.cc:53).
__global__ void k_bc_match(Tab V, const StepState *ss, const ImgKey *__restrict__ keys, int nimg, int32_t *__restrict__ img_row,
                           int32_t *__restrict__ list, int *__restrict__ count, const int32_t *__restrict__ img_cell,
                           const int32_t *__restrict__ owner, int me)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ss->n) return;
  if (V.i[I_BC][p] != 2 || V.i[I_CELL][p] < 0) return;
  if (!owner && V.i[I_GUEST][p]) return;
  const uint32_t k0 = (uint32_t) V.i[I_KEY0][p], k1 = (uint32_t) V.i[I_KEY1][p], k2 = (uint32_t) V.i[I_KEY2][p];
  int lo = 0, hi = nimg;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const ImgKey &q = keys[mid];
    const bool less = q.k[0] < k0 || (q.k[0] == k0 && (q.k[1] < k1 || (q.k[1] == k1 && q.k[2] < k2)));
    if (less) lo = mid + 1; else hi = mid;
  }
  if (lo < nimg && keys[lo].k[0] == k0 && keys[lo].k[1] == k1 && keys[lo].k[2] == k2) {
    const int q = keys[lo].idx;
    img_row[q] = p;
    if (!owner || owner[img_cell[q]] == me) list[atomicAdd(count, 1)] = q;
  }
}

__global__ void __launch_bounds__(256) k_apply_bcond(DevView V, GridView G, KParams K, const StepState *ss, const int32_t *__restrict__ list,
                                                     const int *__restrict__ count, const int32_t *__restrict__ img_row,
                                                     const double *__restrict__ img_coord, const int32_t *__restrict__ img_cell)
{
  const int lane = threadIdx.x & 31;
  const int wglobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int n = *count;
  const double hmax = ss->hmax;
  for (int e = wglobal; e < n; e += nw) {
    const int q = list[e], g = img_row[q];
    const double rx = img_coord[3 * q], ry = img_coord[3 * q + 1], rz = img_coord[3 * q + 2];
    const int cl = img_cell[q];
    const int ck = cl % G.ncell[2], cj = (cl / G.ncell[2]) % G.ncell[1], ci = cl / (G.ncell[2] * G.ncell[1]);
    double u0 = 0., u1 = 0., u2 = 0., u3 = 0., u4 = 0., wn = 0., ks = 0.;
    // Rows that can carry weight sit within cutoff * hmax of the image point: only the subcell columns that box
    // crosses are walked (+1 subcell of slack for rows on a subcell face), clipped to the image's bucket and its 26
    // neighbours — the buckets bcond.cc:93-188 visits.  One lane per (x, y) subcell column; the z-run of a column
    // inside one bucket is one contiguous row range of the sorted table.
    const int R = (int) ceil(K.cutoff * hmax * G.inv_subw) + 1;
    const int cc[3] = { ci, cj, ck };
    const double rc[3] = { rx, ry, rz };
    int lo[3], hi[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
      const int gs = cc[d] * SUBDIV + dev_sub_of(G, d, cc[d], rc[d]);
      lo[d] = max(gs - R, max(cc[d] - 1, 0) * SUBDIV);
      hi[d] = min(gs + R, min(cc[d] + 2, G.ncell[d]) * SUBDIV - 1);
    }
    const int ny = hi[1] - lo[1] + 1, ncol = (hi[0] - lo[0] + 1) * ny;
    for (int c = lane; c < ncol; c += 32) {
      const int gx = lo[0] + c / ny, gy = lo[1] + c % ny;
      const int bx = gx / SUBDIV, by = gy / SUBDIV;
      const int colsub = (gx % SUBDIV) * (SUBDIV * SUBDIV) + (gy % SUBDIV) * SUBDIV;
      for (int bz = lo[2] / SUBDIV; bz <= hi[2] / SUBDIV; bz++) {
        const int z0 = max(lo[2], bz * SUBDIV) - bz * SUBDIV, z1 = min(hi[2], bz * SUBDIV + SUBDIV - 1) - bz * SUBDIV;
        const int r = __ldg(&G.cell2rank[(bx * G.ncell[1] + by) * G.ncell[2] + bz]);
        const int b0 = __ldg(&G.sub_start[r * SUBS + colsub + z0]), b1 = __ldg(&G.sub_start[r * SUBS + colsub + z1 + 1]);
        const bool own = bx == ci && by == cj && bz == ck;
        for (int j = b0; j < b1; j++) {
          const int bc = V.t.i[I_BC][j];
          if (!(bc == 100 || bc == 99 || bc == 1)) continue;          // contr_image, particle.h:562-565
          const double4 gj = ld_rec(V.geo, j);
          const double h = gj.w, supp = K.cutoff2 * h;
          const double dx = rx - gj.x, dy = ry - gj.y, dz = rz - gj.z;
          if (fabs(dx) > supp || fabs(dy) > supp || fabs(dz) > supp) continue;   // in_support, :104 / :156
          const double sx = dx / h, sy = dy / h, sz = dz / h;
          if (fabs(sx) > K.cutoff || fabs(sy) > K.cutoff || fabs(sz) > K.cutoff) continue;  // weight() = 0, sph_lib.cc:216-218
          const double t0 = 0.564189583547756 / h;
          const double w = t0 * t0 * t0 * exp(-(sx * sx + sy * sy + sz * sz));
          const double mw = V.t.d[D_MASS][j] * w, mwr = mw / V.t.d[D_RHO][j];
          u0 += mw;
          u1 += mwr * V.t.d[D_VX][j]; u2 += mwr * V.t.d[D_VY][j]; u3 += mwr * V.t.d[D_VZ][j];
          u4 += mwr * V.t.d[D_E][j];
          if (!own) ks += mwr * V.t.d[D_KS][j];                        // :172-174: neighbour buckets only (sic)
          wn += mwr;
        }
      }
    }
    u0 = warp_sum(u0); u1 = warp_sum(u1); u2 = warp_sum(u2); u3 = warp_sum(u3); u4 = warp_sum(u4);
    wn = warp_sum(wn); ks = warp_sum(ks);
    if (lane == 0) {
      if (wn > 0.) { u0 /= wn; u1 /= wn; u2 /= wn; u3 /= wn; u4 /= wn; ks /= wn; }   // :191-198
      else { u0 = 1.; u1 = u2 = u3 = u4 = 0.; ks = 0.; }
      // wall normal from the ghost to its image (:66-74), reflect() sph_lib.cc:75-83
      double nx = rx - V.t.d[D_X][g], ny = ry - V.t.d[D_Y][g], nz = rz - V.t.d[D_Z][g];
      const double d = sqrt(nx * nx + ny * ny + nz * nz);
      nx /= d; ny /= d; nz /= d;
      const double vdotn = u1 * nx + u2 * ny + u3 * nz;
      V.t.d[D_RHO][g] = u0;
      V.t.d[D_VX][g] = u1 - 2 * vdotn * nx; V.t.d[D_VY][g] = u2 - 2 * vdotn * ny; V.t.d[D_VZ][g] = u3 - 2 * vdotn * nz;
      V.t.d[D_E][g] = u4;
      V.t.d[D_KS][g] = ks;
      const double cprev = K.compress == 1 ? 1482. : V.t.d[D_C][g];  // :223-225
      const EosOut o = dev_eos(K, u0, ks, u4, cprev, V.t.d[D_CP][g], V.t.d[D_GAMMA][g]);
      V.t.d[D_P][g] = o.p; V.t.d[D_GAMMA][g] = o.gamma; V.t.d[D_C][g] = o.c; V.t.d[D_T][g] = o.T; V.t.d[D_CP][g] = o.Cp;
      V.t.d[D_RPH0][g] = o.rph0; V.t.d[D_RPH1][g] = o.rph1;
    }
  }
}

// an image whose wall ghost is not in the table: the reference stops with "ghost particle missing" (bcond.cc:59-66, return 3);
// here it is counted into the invariant-violation counter psph_errors() / psph_step() report (single rank: on slabs the ghost
// of an image this rank computes may legitimately live on another rank only)
__global__ void k_bc_unmatched(const int32_t *__restrict__ img_row, int n_img, StepState *ss)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_img && img_row[i] < 0) atomicAdd(&ss->bad, 1);
}

extern "C" int psph_set_images(psph_ctx *ctx, const psph_image *img, size_t n)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  PSPH_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(ctx->d_img_keys); cudaFree(ctx->d_img_coord); cudaFree(ctx->d_img_cell); cudaFree(ctx->d_img_row);
  ctx->d_img_keys = nullptr; ctx->d_img_coord = nullptr; ctx->d_img_cell = nullptr; ctx->d_img_row = nullptr;
  ctx->n_img = 0;
  if (!img || n == 0) return 0;
  std::vector<ImgKey> keys(n);
  std::vector<double> coord(3 * n);
  std::vector<int32_t> cell(n);
  for (size_t i = 0; i < n; i++) {
    if (img[i].cell < 0 || img[i].cell >= ctx->ncells) { ctx->err = "psph_set_images: image bucket outside the mesh"; return -2; }
    for (int k = 0; k < 3; k++) { keys[i].k[k] = img[i].ghost_key[k]; coord[3 * i + k] = img[i].coord[k]; }
    keys[i].idx = (int32_t) i;
    cell[i] = img[i].cell;
  }
  std::sort(keys.begin(), keys.end(), [](const ImgKey &a, const ImgKey &b) {
    return a.k[0] != b.k[0] ? a.k[0] < b.k[0] : (a.k[1] != b.k[1] ? a.k[1] < b.k[1] : a.k[2] < b.k[2]);
  });
  PSPH_CUDA(cudaMalloc(&ctx->d_img_keys, n * sizeof(ImgKey)));
  PSPH_CUDA(cudaMalloc(&ctx->d_img_coord, 3 * n * sizeof(double)));
  PSPH_CUDA(cudaMalloc(&ctx->d_img_cell, n * sizeof(int32_t)));
  PSPH_CUDA(cudaMalloc(&ctx->d_img_row, n * sizeof(int32_t)));
  PSPH_CUDA(cudaMemcpy(ctx->d_img_keys, keys.data(), n * sizeof(ImgKey), cudaMemcpyHostToDevice));
  PSPH_CUDA(cudaMemcpy(ctx->d_img_coord, coord.data(), 3 * n * sizeof(double), cudaMemcpyHostToDevice));
  PSPH_CUDA(cudaMemcpy(ctx->d_img_cell, cell.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice));
  ctx->n_img = (int) n;
  return 0;
}

extern "C" int psph_apply_bcond(psph_ctx *ctx)
{
  PSPH_CUDA(cudaSetDevice(ctx->cfg.device));
  if (ctx->cfg.dim != 3) return 0;                     // the 1-D loop has no wall ghosts (main.cc:676-751)
  const bool slabs = ctx->cfg.nranks > 1;
  const int32_t *owner = nullptr;
  if (slabs && !(owner = psph_comm_owner_table(ctx))) { ctx->err = "psph_apply_bcond: psph_comm_init was not called"; return -2; }
  if (ctx->n_img == 0) return 0;                       // (every rank holds the same image table: all return here together)
  StageTimer tm(ctx, ST_BCOND);
  // slabs: rows injected since the last exchange (add_new_erupt) reach their neighbours' ghost layers first (main.cc:430)
  if (slabs && psph_exchange_ghosts(ctx)) return -1;
  if (!ctx->list_valid && psph_build_cell_list(ctx)) return -1;   // positions moved since the last list build
  // scratch: local row of every image's ghost, the images this rank computes, their count
  int32_t *list = reinterpret_cast<int32_t *>(ctx->d_perm);
  if ((size_t) ctx->n_img > ctx->cap && psph_reserve(ctx, ctx->n_img)) return -1;       // (the list lives in the sort scratch)
  list = reinterpret_cast<int32_t *>(ctx->d_perm);
  int *count = &ctx->d_ss->work;
  PSPH_CUDA(cudaMemsetAsync(count, 0, sizeof(int), ctx->stream));
  PSPH_CUDA(cudaMemsetAsync(ctx->d_img_row, 0xff, (size_t) ctx->n_img * sizeof(int32_t), ctx->stream));
  if (ctx->n > 0) {
    k_bc_match<<<(ctx->n + 255) / 256, 256, 0, ctx->stream>>>(ctx->cur.t, ctx->d_ss, (const ImgKey *) ctx->d_img_keys, ctx->n_img,
                                                            ctx->d_img_row, list, count, ctx->d_img_cell, owner, ctx->cfg.rank);
    PSPH_LAUNCH_CHECK();
    if (!slabs) { k_bc_unmatched<<<(ctx->n_img + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_img_row, ctx->n_img, ctx->d_ss); PSPH_LAUNCH_CHECK(); }
    const int grid = std::min((ctx->n_img + 7) / 8, ctx->sms * 8);
    k_apply_bcond<<<grid, 256, 0, ctx->stream>>>(ctx->cur, psph_grid_view(ctx), ctx->kp, ctx->d_ss, list, count, ctx->d_img_row,
                                                 ctx->d_img_coord, ctx->d_img_cell);
    PSPH_LAUNCH_CHECK();
  }
  // slabs: a ghost whose image this rank computed may be another rank's row (here a guest copy): the state goes back to
  // its owner (move_bnd_images, move_bnd_img.cc:32-210), whose next ghost refresh hands it to everybody
  if (slabs && psph_comm_return_images(ctx, ctx->d_img_row, ctx->d_img_cell, ctx->n_img)) return -1;
  psph_comm_mark(ctx, DIRTY_MOM | DIRTY_DENS);
  return 0;
}
