// Throughput of 64-bit conversion / compare instructions vs DFMA (full occupancy, independent chains).
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double* out, int iters) {
  double a[4];
  float f[4];
  int n[4];
  for (int j = 0; j < 4; ++j) { a[j] = threadIdx.x + j + 0.5; f[j] = threadIdx.x * 0.25f + j; n[j] = threadIdx.x + j; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (OP == 0) a[j] = fma(a[j], 1.0000001, 1e-9);                         // DFMA
      if (OP == 1) { a[j] = (double)n[j]; n[j] = n[j] * 3 + (int)__double2loint(a[j]); }  // I2F.F64 (+ int ops)
      if (OP == 2) { a[j] = (double)f[j]; f[j] = __int_as_float(__double2hiint(a[j]) ^ 0x00100000); }  // F2F.F64.F32
      if (OP == 3) { f[j] = (float)a[j]; a[j] = __hiloint2double(__float_as_int(f[j]) | 0x3ff00000, i); }  // F2F.F32.F64
      if (OP == 4) { n[j] += (a[j] > (double)0.5 + n[j] * 0.0) ? 1 : 2; a[j] = __hiloint2double(0x3ff00000 + (n[j] & 0xff), n[j]); }  // DSETP (+ DMUL/DADD)
      if (OP == 5) { a[j] = __hiloint2double(0x43300000, n[j]) - 4503599627370496.0; n[j] = n[j] * 3 + __double2loint(a[j]); }  // magic int->double
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a[0] + a[1] + a[2] + a[3] + f[0] + f[1] + f[2] + f[3] + n[0] + n[1] + n[2] + n[3];
}
template <int OP>
void run(const char* name, double* out, int sms) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4000, blocks = sms * 2, threads = 1024;
  k<OP><<<blocks, threads>>>(out, 10);
  cudaEventRecord(e0); k<OP><<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)blocks * threads * iters * 4;
  printf("%-28s %.1f lane-ops/clk/SM\n", name, ops / (ms * 1e-3) / sms / 1.965e9);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 2 * 1024 * 8);
  run<0>("DFMA", out, p.multiProcessorCount);
  run<1>("I2F.F64.S32 (+IMAD)", out, p.multiProcessorCount);
  run<2>("F2F.F64.F32 (+LOP)", out, p.multiProcessorCount);
  run<3>("F2F.F32.F64 (+LOP)", out, p.multiProcessorCount);
  run<4>("DSETP+I2F+DMUL+DADD", out, p.multiProcessorCount);
  run<5>("magic int->double (DADD)", out, p.multiProcessorCount);
  return 0;
}
