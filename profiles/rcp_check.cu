// rcp_check.cu — accuracy of the FP64 reciprocal seed (rcp.approx.ftz.f64 = MUFU.RCP64H) and of 1, 2, 3 Newton steps / one third-order step on it,
// over 2^24 operands spread across [2^-20, 2^20]: how many steps dev_rcp (psph_gather.cuh) needs.  Prints one JSON line.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__global__ void k(double *err)
{
  const unsigned long long id = blockIdx.x * (unsigned long long) blockDim.x + threadIdx.x;
  unsigned long long s = id * 6364136223846793005ull + 1442695040888963407ull;
  s ^= s >> 29; s *= 0xbf58476d1ce4e5b9ull; s ^= s >> 32;
  const double m = 1.0 + (double) (s >> 11) * 0x1p-53;               // [1, 2)
  const double x = ldexp(m, (int) (id % 41) - 20);
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double exact = 1.0 / x;
  double e[6];
  e[0] = fabs(r - exact) / exact;
  {                                                                  // one third-order step r (1 + e + e^2): what dev_rcp does
    const double t = fma(-x, r, 1.0), rh = fma(r, fma(t, t, t), r);
    e[4] = fabs(rh - exact) / exact;
    e[5] = rh != exact ? 1.0 : 0.0;                                  // (max -> 1 if any operand is not correctly rounded)
  }
  for (int k = 1; k < 4; k++) { const double t = fma(-x, r, 1.0); r = fma(r, t, r); e[k] = fabs(r - exact) / exact; }
  for (int k = 0; k < 6; k++) {
    double v = e[k];
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long *) &err[k], (unsigned long long) __double_as_longlong(v));
  }
}
int main()
{
  double *d, h[6];
  cudaMalloc(&d, 48); cudaMemset(d, 0, 48);
  k<<<65536, 256>>>(d);
  cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
  printf("{\"seed_relerr\": %.3e, \"newton1\": %.3e, \"newton2\": %.3e, \"newton3\": %.3e, \"halley\": %.3e, \"halley_any_misrounded\": %d, \"ulp\": %.3e}\n", h[0], h[1], h[2], h[3], h[4], (int) h[5], 0x1p-53);
  return 0;
}
