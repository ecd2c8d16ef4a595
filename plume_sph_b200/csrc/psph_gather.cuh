// psph_gather.cuh — the pair gather that every SPH summation stage runs on.
//
// One WARP per particle i, lanes across the candidate rows j.  The candidates of i are the
// rows of the subcells (2h_air cubes, 64 per bucket) that overlap the box x_i +- reach; for
// every (x, y) subcell column and every bucket the column crosses, the z-run of subcells is
// ONE contiguous row range of the sorted table, so a warp reads its candidates as a few
// dozen fully coalesced 128-bit row-record loads ({x,y,z,h} double4 + flag word), and the
// records of the rows that pass the gates are read the same way.  Per pair the reference's
// own gates decide membership, bit for bit:
//   list gate   : j in i's bucket, or |dx_k| <= 4.25*h_i(list) for all k   (neighs.cc:84-92,130)
//   stage gates : in_support(dx, CUTOFF2*h_i) and the kernel cutoff |dx_k/h| <= 3 inside
//                 Stage::pair                                              (sph_header.h:49-60, sph_lib.cc:216-218)
// so the set of pairs that contribute is exactly the reference's; pairs that cannot contribute
// (zero kernel weight) are skipped.  Lane partial sums are combined with warp shuffles.
// CTAs are persistent and take chunks of consecutive (spatially adjacent) rows from an atomic
// counter, so the candidate rows of a chunk stay in L1/L2.
#pragma once
#include "psph_internal.cuh"

constexpr int GW = 8;          // warps per CTA
constexpr int GT = GW * 32;    // threads per CTA
constexpr int RPW = 4;         // consecutive rows per warp and chunk
constexpr int CHUNK = GW * RPW;

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v)
{
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double4 ld_rec(const double4 *__restrict__ a, int j)
{ // two 128-bit loads of one 32-byte record
  const double2 *p = reinterpret_cast<const double2 *>(a + j);
  double2 lo = __ldg(p), hi = __ldg(p + 1);
  return make_double4(lo.x, lo.y, hi.x, hi.y);
}

// fabs(dx / h) > cutoff, decided exactly as the reference's division would (sph_lib.cc:216-218)
// but without dividing unless the pair sits within a few ulp of the cutoff face
__device__ __forceinline__ bool beyond_cutoff(double dx, double h, double cutoff)
{
  double a = fabs(dx), b = cutoff * h;
  if (a > b * (1.0 + 0x1p-48)) return true;
  if (a < b * (1.0 - 0x1p-48)) return false;
  return fabs(dx / h) > cutoff;
}
template <int DIM>
__device__ __forceinline__ bool outside_kernel(double dx, double dy, double dz, double h, double cutoff)
{
  if (beyond_cutoff(dx, h, cutoff)) return true;
  if (DIM == 3) return beyond_cutoff(dy, h, cutoff) || beyond_cutoff(dz, h, cutoff);
  return false;
}
template <int DIM>
__device__ __forceinline__ bool dev_in_support(double dx, double dy, double dz, double supp)
{
  if (fabs(dx) > supp) return false;
  if (DIM == 3 && (fabs(dy) > supp || fabs(dz) > supp)) return false;
  return true;
}
// SPH kernel (sph_lib.cc:207-230) for a pair already known to be inside the cutoff:
// w = (pi^-1/2 / h)^DIM * exp(-|dx|^2/h^2)
template <int DIM>
__device__ __forceinline__ double dev_weight_in(double r2, double invh)
{
  const double t = 0.564189583547756 * invh;
  const double norm = DIM == 3 ? t * t * t : t;
  return norm * exp(-(r2 * invh * invh));
}

// Stage interface (static members):
//   struct I      per-row state, identical in every lane
//   struct A      per-lane accumulators
//   bool  init(V,K,ss,i,I&)             load row i; false = nothing to do
//   double reach(K,I,hmax)              half-width of the candidate box
//   int   passes(K,I)                   how many sweeps over the candidates (adaptive_sml > 1)
//   void  begin(K,I&,A&,pass)           zero the accumulators
//   void  pair(V,K,I,A&,dx,dy,dz,hj,fj,j)  one candidate that passed the list gate
//   bool  end(V,K,ss,i,I&,A&,pass,lane) warp-reduce + write; true = run another pass
template <int DIM, class S>
__global__ void __launch_bounds__(GT) k_gather(DevView V, GridView G, KParams K, StepState *ss)
{
  __shared__ int s_chunk;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = ss->n;
  const double hmax = ss->hmax;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_chunk = atomicAdd(&ss->work, 1);
    __syncthreads();
    const int row0 = s_chunk * CHUNK + warp * RPW;
    if (s_chunk * CHUNK >= n) return;
    for (int r = 0; r < RPW; r++) {
      const int i = row0 + r;
      if (i >= n) break;
      typename S::I st;
      if (!S::init(V, K, ss, i, st)) continue;
      const double4 gi = V.geo[i];
      const double rlist = K.neigh_radius * V.t.d[D_HLIST][i];
      const int cell_i = V.t.i[I_CELL][i];
      typename S::A acc;
      int pass = 0;
      bool again = true;
      while (again) {
        S::begin(K, st, acc, pass);
        if (DIM == 1) {
          // the 1-D reference has no search: every particle lists every particle (add_air.cc:470-479)
          for (int j = lane; j < n; j += 32) {
            const double4 gj = ld_rec(V.geo, j);
            S::pair(V, K, st, acc, gi.x - gj.x, 0.0, 0.0, gj.w, V.flags[j], j);
          }
        } else {
          const double R = S::reach(K, st, hmax) * (1.0 + 1e-12);
          const int gx0 = dev_gsub(G, 0, gi.x - R), gx1 = dev_gsub(G, 0, gi.x + R);
          const int gy0 = dev_gsub(G, 1, gi.y - R), gy1 = dev_gsub(G, 1, gi.y + R);
          const int gz0 = dev_gsub(G, 2, gi.z - R), gz1 = dev_gsub(G, 2, gi.z + R);
          const int nx = gx1 - gx0 + 1, ny = gy1 - gy0 + 1;
          const int cz0 = gz0 >> 2, ncz = (gz1 >> 2) - cz0 + 1;
          const int nseg = nx * ny * ncz;
          for (int base = 0; base < nseg; base += 32) {
            // lane s of this batch resolves segment base+s to a row range
            int beg = 0, end = 0, own = 0;
            const int sg = base + lane;
            if (sg < nseg) {
              const int icz = sg % ncz, t = sg / ncz, iy = t % ny, ix = t / ny;
              const int gx = gx0 + ix, gy = gy0 + iy, cz = cz0 + icz;
              const int cell = ((gx >> 2) * G.ncell[1] + (gy >> 2)) * G.ncell[2] + cz;
              const int zlo = (cz == cz0) ? (gz0 & 3) : 0, zhi = (cz == (gz1 >> 2)) ? (gz1 & 3) : 3;
              const int e = __ldg(&G.cell2rank[cell]) * SUBS + (gx & 3) * 16 + (gy & 3) * 4;
              beg = __ldg(&G.sub_start[e + zlo]);
              end = __ldg(&G.sub_start[e + zhi + 1]);
              own = cell == cell_i;
            }
            const int cnt = min(32, nseg - base);
            for (int s = 0; s < cnt; s++) {
              const int b = __shfl_sync(0xffffffffu, beg, s), e = __shfl_sync(0xffffffffu, end, s);
              const int o = __shfl_sync(0xffffffffu, own, s);
              for (int j = b + lane; j < e; j += 32) {
                const double4 gj = ld_rec(V.geo, j);
                const double dx = gi.x - gj.x, dy = gi.y - gj.y, dz = gi.z - gj.z;
                // neighbour-list membership (neighs.cc:88-92 own bucket, :130 other buckets)
                if (!o && (fabs(dx) > rlist || fabs(dy) > rlist || fabs(dz) > rlist)) continue;
                S::pair(V, K, st, acc, dx, dy, dz, gj.w, V.flags[j], j);
              }
            }
          }
        }
        again = S::end(V, K, ss, i, st, acc, pass, lane);
        pass++;
      }
    }
  }
}

static __global__ void k_reset_work(StepState *ss) { ss->work = 0; }

template <int DIM, class S>
int launch_gather_d(psph_ctx *ctx, int kid)
{
  k_reset_work<<<1, 1, 0, ctx->stream>>>(ctx->d_ss);
  PSPH_LAUNCH_CHECK();
  KernelTimer kt(ctx, kid);
  int nchunks = (ctx->n + CHUNK - 1) / CHUNK;
  int grid = min(max(nchunks, 1), ctx->sms * 8);
  k_gather<DIM, S><<<grid, GT, 0, ctx->stream>>>(ctx->cur, psph_grid_view(ctx), ctx->kp, ctx->d_ss);
  PSPH_LAUNCH_CHECK();
  return 0;
}
template <template <int> class S>
int launch_gather(psph_ctx *ctx, int kid)
{
  if (ctx->cfg.dim == 3) return launch_gather_d<3, S<3> >(ctx, kid);
  return launch_gather_d<1, S<1> >(ctx, kid);
}
