// fp64_rate.cu — microbenchmark: sustained DMMA.884 and DFMA issue rates per SM on this GPU, alone and
// mixed.  The CD kernels' ceiling depends on these two numbers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rate fp64_rate.cu && ./fp64_rate
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE>   // 0: DMMA only, 1: DFMA only, 2: 1 DMMA + 8 DFMA interleaved
__global__ void k(double* out, int iters, double a0, double b0) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; i++) c[i] = threadIdx.x * 1e-3 + i;
  double a = a0 + threadIdx.x, b = b0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0 || MODE == 2) dmma(c[2 * i], c[2 * i + 1], a, b);
      if (MODE == 1) { c[2 * i] = fma(a, b, c[2 * i]); c[2 * i + 1] = fma(a, b, c[2 * i + 1]); }
      if (MODE == 2) {
#pragma unroll
        for (int q = 0; q < 8; q++) c[(2 * i + 2 + q) & 15] = fma(a, b, c[(2 * i + 2 + q) & 15]);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += c[i];
  if (s == 123.456) out[0] = s;
}

template <int MODE>
void run(const char* name, int threads, int ctas_per_sm, int sms, double clk_ghz) {
  double* out; cudaMalloc(&out, 8);
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms * ctas_per_sm, threads>>>(out, 100, 1.0, 1e-9);
  cudaEventRecord(e0);
  k<MODE><<<sms * ctas_per_sm, threads>>>(out, iters, 1.0, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double warps = (double)threads / 32 * ctas_per_sm;
  const double clk = ms * 1e-3 * clk_ghz * 1e9;
  const double dmma_per_sm = (MODE == 1) ? 0 : warps * iters * 8;
  const double dfma_warp_per_sm = (MODE == 0) ? 0 : warps * iters * (MODE == 1 ? 16 : 64);
  printf("%-10s warps/SM %4.0f  %8.3f ms  ", name, warps, ms);
  if (dmma_per_sm) printf("clk per DMMA per SM %6.2f (per SMSP %6.2f)  ", clk / dmma_per_sm, 4 * clk / dmma_per_sm);
  if (dfma_warp_per_sm) printf("clk per warp-DFMA per SM %5.2f (per SMSP %5.2f)", clk / dfma_warp_per_sm, 4 * clk / dfma_warp_per_sm);
  printf("\n");
  cudaFree(out);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const double ghz = p.clockRate * 1e-6;
  printf("%s, %d SMs, clock %.3f GHz (nominal; rates assume it)\n", p.name, p.multiProcessorCount, ghz);
  for (int w : {1, 2, 4, 8, 16, 32}) {
    const int threads = w * 32 > 1024 ? 1024 : w * 32;
    const int ctas = w * 32 / threads;
    run<0>("DMMA", threads, ctas, p.multiProcessorCount, ghz);
  }
  for (int w : {4, 8, 16, 32}) run<1>("DFMA", w * 32, 1, p.multiProcessorCount, ghz);
  for (int w : {4, 8, 16, 32}) run<2>("DMMA+8DFMA", w * 32, 1, p.multiProcessorCount, ghz);
  return 0;
}
