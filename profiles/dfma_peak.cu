// dfma_peak.cu — measured FP64 FMA peak of this GPU: the denominator for "FP64 pipe utilisation" of the pair kernels
// (BASELINE.md §3 asks for a measured figure before quoting one).  Every thread runs CHAINS independent dependent-FMA
// chains; enough warps per SM and enough chains per warp to cover the DFMA latency.  Prints one JSON line.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dfma_peak profiles/dfma_peak.cu && ./dfma_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void __launch_bounds__(1024, 1) k_dfma(double *out, int iters, double a, double b)
{
  double x[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) x[c] = 1.0 + 1e-9 * (threadIdx.x + c);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) x[c] = fma(x[c], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) s += x[c];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true: keeps the chains alive
}

template <int CHAINS>
double run(int sms, int threads, int iters, double *d_out)
{
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_dfma<CHAINS><<<sms, threads>>>(d_out, iters / 10, 0.999999, 1e-7);   // warm-up
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    k_dfma<CHAINS><<<sms, threads>>>(d_out, iters, 0.999999, 1e-7);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double fma_count = (double) sms * threads * (double) iters * CHAINS;
  return 2.0 * fma_count / (best * 1e-3) / 1e12;   // TFLOP/s
}

int main()
{
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double *d_out;
  cudaMalloc(&d_out, (size_t) p.multiProcessorCount * 1024 * sizeof(double));
  const int iters = 200000;
  double best = 0, t;
  int bt = 0, bc = 0;
  for (int threads : { 256, 512, 1024 }) {
    t = run<4>(p.multiProcessorCount, threads, iters, d_out); if (t > best) { best = t; bt = threads; bc = 4; }
    t = run<8>(p.multiProcessorCount, threads, iters, d_out); if (t > best) { best = t; bt = threads; bc = 8; }
    t = run<16>(p.multiProcessorCount, threads, iters, d_out); if (t > best) { best = t; bt = threads; bc = 16; }
  }
  // FMA lanes per SM and clock implied by the measurement
  const double per_sm_clk = best * 1e12 / 2.0 / p.multiProcessorCount / (clk_khz * 1e3);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp64_fma_tflops\": %.3f, \"dfma_per_s\": %.4e, \"threads_per_sm\": %d, \"chains\": %d, "
         "\"fma_per_sm_per_clk_at_max_clock\": %.2f, \"max_clock_mhz\": %d, \"how\": \"dependent FMA chains, best of 5, CUDA events\"}\n",
         p.name, p.multiProcessorCount, best, best * 1e12 / 2.0, bt, bc, per_sm_clk, clk_khz / 1000);
  return 0;
}
