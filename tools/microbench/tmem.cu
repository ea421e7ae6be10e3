// TMEM read / write port micro-benchmark (one CTA on one SM): bytes per clock of tcgen05.ld for W warps x shape
// .32x32b.xN, with one or two loads issued per tcgen05.wait::ld, and of the hidden-drain pattern of net_tc
// (ld -> cvt.rn.relu.bf16x2 -> st in place).  Run on the GPU box: gpurun -- ./tools/microbench/tmem
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int N>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&r)[N]);
#define LD_BODY(N, REGS, ...)                                                                             \
  template <>                                                                                             \
  __device__ __forceinline__ void ld<N>(uint32_t taddr, uint32_t (&r)[N]) {                               \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x" #N ".b32 {" REGS "}, [%" #N "];" : __VA_ARGS__ : "r"(taddr) : "memory"); \
  }
#define O4(b) "=r"(r[b]), "=r"(r[b + 1]), "=r"(r[b + 2]), "=r"(r[b + 3])
#define O16(b) O4(b), O4(b + 4), O4(b + 8), O4(b + 12)
LD_BODY(16, "%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15", O16(0))
LD_BODY(32,
        "%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31",
        O16(0), O16(16))
LD_BODY(64,
        "%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
        "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63",
        O16(0), O16(16), O16(32), O16(48))

__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ uint32_t pack_relu(uint32_t lo, uint32_t hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  return r;
}

// mode 0: ld + wait; 1: two lds per wait; 2: drain (ld, cvt, st in place, x16 double-buffered as net_tc v4); 3: drain with xN per wait
template <int N, int MODE>
__global__ void __launch_bounds__(512, 1) k(long long* cyc, uint32_t* sink, int iters) {
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // warp w reads lanes 32 (w & 3); its column window of 512 / (warps / 4) columns
  const int groups = (blockDim.x >> 5) >> 2, g = warp >> 2;
  const int cols = 512 / (groups < 1 ? 1 : groups);
  const uint32_t T = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(g * cols);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      for (int c = 0; c + N <= cols; c += N) {
        uint32_t r[N];
        ld<N>(T + c, r);
        wait_ld();
        acc ^= r[0] ^ r[N - 1];
      }
    } else if (MODE == 1) {
      for (int c = 0; c + 2 * N <= cols; c += 2 * N) {
        uint32_t r[N], q[N];
        ld<N>(T + c, r);
        ld<N>(T + c + N, q);
        wait_ld();
        acc ^= r[0] ^ q[N - 1];
      }
    } else if (MODE == 3) {  // N columns per wait, converted and stored in place behind the read pointer
      for (int c = 0; c + N <= cols; c += N) {
        uint32_t r[N];
        ld<N>(T + c, r);
        wait_ld();
        uint32_t pk[N / 2];
#pragma unroll
        for (int j = 0; j < N / 2; ++j) pk[j] = pack_relu(r[2 * j], r[2 * j + 1]);
#pragma unroll
        for (int j = 0; j < N / 16; ++j) st8(T + c / 2 + 8 * j, pk + 8 * j);
      }
      wait_st();
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  sink[threadIdx.x] = acc;
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

template <int N, int MODE>
void run(int warps, long long* cyc, uint32_t* sink, const char* what) {
  const int iters = 200;
  k<N, MODE><<<1, warps * 32>>>(cyc, sink, iters);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s x%d warps %d: %s\n", what, N, warps, cudaGetErrorString(e)); return; }
  const int groups = warps / 4, cols = 512 / groups;
  const int per = MODE == 1 ? (cols / (2 * N)) * 2 * N : (cols / N) * N;
  const double bytes = (double)iters * warps * 32.0 * per * 4.0;
  printf("%-28s x%-3d %2d warps: %8.1f B/clk  (%lld clk for %d x %d columns per warp)\n", what, N, warps, bytes / (double)cyc[0], cyc[0] / iters, 1, per);
}

int main() {
  long long* cyc; cudaMallocManaged(&cyc, 8);
  uint32_t* sink; cudaMalloc(&sink, 4096);
  for (int warps : {4, 8, 16}) {
    run<16, 0>(warps, cyc, sink, "ld + wait");
    run<32, 0>(warps, cyc, sink, "ld + wait");
    run<64, 0>(warps, cyc, sink, "ld + wait");
    run<16, 1>(warps, cyc, sink, "2 lds per wait");
    run<32, 1>(warps, cyc, sink, "2 lds per wait");
    run<16, 3>(warps, cyc, sink, "drain ld/cvt/st");
    run<32, 3>(warps, cyc, sink, "drain ld/cvt/st");
    run<64, 3>(warps, cyc, sink, "drain ld/cvt/st");
  }
  return 0;
}
