// fp64_peaks.cu -- step 0 of SURVEY.md section 7: measure the B200 FP64 denominators.
//   DFMA issue peak, DMMA (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4) issue peak at several
//   occupancies / accumulator counts.  Prints one JSON object.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int CHAINS>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double acc[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

template <int NACC>
__global__ void dmma_kernel(double* out, int iters, double a0, double b0) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

template <typename F>
float time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  double* out;
  CK(cudaMalloc(&out, 8));
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, sms, p.clockRate);
  {
    const int iters = 20000, threads = 512, blocks = sms * 4;
    float ms = time_ms([&] { dfma_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
    double flops = 2.0 * 8 * iters * (double)threads * blocks;
    printf(", \"dfma_tflops\": %.3f", flops / ms * 1e-9);
  }
#define RUN_DMMA(NACC, WARPS, BPS)                                                                     \
  {                                                                                                    \
    const int iters = 4000, threads = WARPS * 32, blocks = sms * BPS;                                  \
    float ms = time_ms([&] { dmma_kernel<NACC><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5); \
    double flops = 2.0 * 256 * NACC * iters * (double)WARPS * blocks;                                  \
    printf(", \"dmma_acc%d_w%d_b%d_tflops\": %.3f", NACC, WARPS, BPS, flops / ms * 1e-9);              \
  }
  RUN_DMMA(1, 4, 1)
  RUN_DMMA(2, 4, 1)
  RUN_DMMA(4, 4, 1)
  RUN_DMMA(8, 4, 1)
  RUN_DMMA(32, 4, 1)
  RUN_DMMA(8, 8, 1)
  RUN_DMMA(32, 8, 1)
  RUN_DMMA(8, 16, 1)
  RUN_DMMA(8, 8, 2)
  // sustained: ~2 s of back-to-back DMMA to see the power-capped clock
  {
    const int iters = 4000, threads = 256, blocks = sms;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int launches = 400;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < launches; ++i) dmma_kernel<32><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double flops = 2.0 * 256 * 32 * iters * 8.0 * blocks * launches;
    printf(", \"dmma_sustained_tflops\": %.3f, \"dmma_sustained_s\": %.2f", flops / ms * 1e-9, ms * 1e-3);
  }
  printf("}\n");
  return 0;
}
