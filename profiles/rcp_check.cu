// rcp_check.cu — accuracy of the FP64 reciprocal seed (rcp.approx.ftz.f64 = MUFU.RCP64H) and of 1, 2, 3 Newton steps on it,
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
  double e[4];
  e[0] = fabs(r - exact) / exact;
  for (int k = 1; k < 4; k++) { const double t = fma(-x, r, 1.0); r = fma(r, t, r); e[k] = fabs(r - exact) / exact; }
  for (int k = 0; k < 4; k++) {
    double v = e[k];
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long *) &err[k], (unsigned long long) __double_as_longlong(v));
  }
}
int main()
{
  double *d, h[4];
  cudaMalloc(&d, 32); cudaMemset(d, 0, 32);
  k<<<65536, 256>>>(d);
  cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
  printf("{\"seed_relerr\": %.3e, \"newton1\": %.3e, \"newton2\": %.3e, \"newton3\": %.3e, \"ulp\": %.3e}\n", h[0], h[1], h[2], h[3], 0x1p-53);
  return 0;
}
