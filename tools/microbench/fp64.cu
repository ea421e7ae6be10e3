// FP64 micro-benchmark: DFMA throughput per SM (independent chains, full occupancy) and dependent-chain latency.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_tp(double* out, int iters) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void ffma_tp(float* out, int iters) {
  float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const float m = 1.0000001f, c = 1e-9f;
  for (int i = 0; i < iters; ++i) {
    a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
    a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void dfma_lat(double* out, long long* cyc, int iters) {
  double a = threadIdx.x;
  const double m = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) a = fma(a, m, c);
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void ddiv_lat(double* out, long long* cyc, int iters) {
  double a = 1e10 + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) a = __ddiv_rn(a, 1.0000001);
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs, %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  double* out; cudaMalloc(&out, 148 * 8 * 1024 * 8);
  long long* cyc; cudaMallocManaged(&cyc, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, blocks = p.multiProcessorCount * 2, threads = 1024;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); dfma_tp<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = (double)blocks * threads * iters * 8;
    printf("DFMA throughput: %.2f TFLOP/s fp64 (%.1f FMA/clk/SM at %.0f MHz)\n", 2 * fmas / ms / 1e9, fmas / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3), p.clockRate / 1e3);
    cudaEventRecord(e0); ffma_tp<<<blocks, threads>>>((float*)out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("FFMA throughput: %.2f TFLOP/s fp32 (%.1f FMA/clk/SM)\n", 2 * fmas / ms / 1e9, fmas / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3));
  }
  dfma_lat<<<1, 32>>>(out, cyc, 10000); cudaDeviceSynchronize();
  printf("DFMA dependent latency, 1 warp: %.1f cycles\n", cyc[0] / 10000.0);
  dfma_lat<<<1, 256>>>(out, cyc, 10000); cudaDeviceSynchronize();
  printf("DFMA dependent latency, 8 warps on one SM: %.1f cycles per op per warp\n", cyc[0] / 10000.0);
  dfma_lat<<<1, 1024>>>(out, cyc, 10000); cudaDeviceSynchronize();
  printf("DFMA dependent latency, 32 warps on one SM: %.1f cycles per op per warp\n", cyc[0] / 10000.0);
  ddiv_lat<<<1, 32>>>(out, cyc, 2000); cudaDeviceSynchronize();
  printf("__ddiv_rn dependent latency, 1 warp: %.1f cycles\n", cyc[0] / 2000.0);
  ddiv_lat<<<1, 1024>>>(out, cyc, 2000); cudaDeviceSynchronize();
  printf("__ddiv_rn dependent latency, 32 warps: %.1f cycles\n", cyc[0] / 2000.0);
  return 0;
}
