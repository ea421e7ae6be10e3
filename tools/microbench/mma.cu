// tcgen05.mma issue / completion timing of the shapes net_tc uses (one CTA, one issuing lane, zeroed operands):
// first layer = 5 x (M128 N256 K16, A and B from shared memory); second layer = 17 x (M128 N<=64 K16, A from TMEM)
// accumulating into ONE tile, or split over two accumulators, or interleaved with another tile's instructions.
// Run on the GPU box: gpurun -- ./tools/microbench/mma
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  const uint64_t hi = 64ull | (1ull << 14) | (2ull << 29);
  return (uint64_t)(((a >> 4) & 0x3FFFu) | (1u << 16)) | (hi << 32);
}
__device__ __forceinline__ uint32_t idesc(uint32_t n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24); }
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// test ids
//  0: L1  = 5 x ss N256                         1: L2 ts N = n, 17 in one accumulator
//  2: L2 ts, 17 split over two accumulators (even / odd k-steps)     3: L2 as ss (A from shared memory), one accumulator
//  4: two tiles' L2 interleaved instruction by instruction (2 x 17)  5: two tiles' L2 back to back (2 x 17)
//  6: L1 of one tile (5) interleaved with L2 of another (17)         7: L1 then L2 back to back (5 + 17)
__global__ void __launch_bounds__(128, 1) k(long long* out, int test, int n, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_base;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t T = tmem_base;
  if (threadIdx.x == 0 && test < 10) {
    const uint32_t a = smem_u32(smem), w1 = a + 32 * 1024, w2 = a + 80 * 1024;
    const uint64_t A = desc_sw128(a), W1 = desc_sw128(w1), W2 = desc_sw128(w2);
    const uint32_t id256 = idesc(256), idn = idesc((uint32_t)n);
    uint32_t ph = 0;
    long long issue = 0, total = 0;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      auto L1 = [&](uint32_t tile, int kk) { mma_ss(T + 256 * tile, A + (uint64_t)(kk * 2), W1 + (uint64_t)(kk * 2), id256, kk ? 1u : 0u); };
      auto L2 = [&](uint32_t tile, int j, uint32_t col) {
        mma_ts(T + 256 * tile + col, T + 256 * tile + (uint32_t)(((j & 15) >> 3) * 128 + (j & 7) * 8),
               W2 + (uint64_t)((((j & 15) >> 2) * n * 128 + (j & 3) * 32) >> 4), idn, j ? 1u : 0u);
      };
      if (test == 0) for (int kk = 0; kk < 5; ++kk) L1(0, kk & 3);
      if (test == 1) for (int j = 0; j < 17; ++j) L2(0, j, 64);
      if (test == 2) for (int j = 0; j < 17; ++j) mma_ts(T + 64 + (uint32_t)(j & 1) * 128, T + (uint32_t)(((j & 15) >> 3) * 128 + (j & 7) * 8),
                                                           W2 + (uint64_t)((((j & 15) >> 2) * n * 128 + (j & 3) * 32) >> 4), idn, j > 1 ? 1u : 0u);
      if (test == 3) for (int j = 0; j < 17; ++j) mma_ss(T + 64, A + (uint64_t)((j & 3) * 2), W2 + (uint64_t)((((j & 15) >> 2) * n * 128 + (j & 3) * 32) >> 4), idn, j ? 1u : 0u);
      if (test == 4) for (int j = 0; j < 17; ++j) { L2(0, j, 64); L2(1, j, 64); }
      if (test == 5) { for (int j = 0; j < 17; ++j) L2(0, j, 64); for (int j = 0; j < 17; ++j) L2(1, j, 64); }
      if (test == 6) for (int j = 0; j < 17; ++j) { L2(0, j, 64); if (j % 3 == 0 && j / 3 < 5) L1(1, (j / 3) & 3); }
      if (test == 7) { for (int kk = 0; kk < 5; ++kk) L1(1, kk & 3); for (int j = 0; j < 17; ++j) L2(0, j, 64); }
      if (test == 8) {
#pragma unroll
        for (int j = 0; j < 17; ++j) mma_ts(T + 64, T + (uint32_t)(((j & 15) >> 3) * 128 + (j & 7) * 8), W2 + (uint64_t)((j & 3) * 2), idn, j ? 1u : 0u);
      }
      if (test == 9) for (int j = 0; j < 17; ++j) { L2(0, j, 64); if (j == 8) commit(&bar2); }
      commit(&bar);
      const long long t1 = clock64();
      mbar_wait(&bar, ph);
      ph ^= 1;
      const long long t2 = clock64();
      if (r > 0) { issue += t1 - t0; total += t2 - t0; }
    }
    out[0] = issue / (reps - 1);
    out[1] = total / (reps - 1);
  }
  if (test >= 10 && threadIdx.x >= 32 && threadIdx.x < 96) {
    const int w = threadIdx.x >> 5;  // 1 or 2: tile w - 1
    const uint32_t a = smem_u32(smem), w1 = a + 32 * 1024, w2 = a + 80 * 1024;
    const uint64_t A = desc_sw128(a), W1 = desc_sw128(w1), W2 = desc_sw128(w2);
    const uint32_t id256 = idesc(256), idn = idesc((uint32_t)n);
    uint64_t* mybar = w == 1 ? &bar : &bar2;
    const uint32_t tile = (uint32_t)(w - 1);
    uint32_t ph = 0;
    long long total = 0;
    for (int r = 0; r < reps; ++r) {
      asm volatile("bar.sync 1, 64;" ::: "memory");
      if ((threadIdx.x & 31) != 0) continue;
      const long long t0 = clock64();
      const bool l1 = (test == 12) || (test == 11 && w == 2);
      if (l1) { for (int kk = 0; kk < 5; ++kk) mma_ss(T + 256 * tile, A + (uint64_t)((kk & 3) * 2), W1 + (uint64_t)((kk & 3) * 2), id256, kk ? 1u : 0u); }
      else {
#pragma unroll
        for (int j = 0; j < 17; ++j) mma_ts(T + 256 * tile + 64, T + 256 * tile + (uint32_t)(((j & 15) >> 3) * 128 + (j & 7) * 8), W2 + (uint64_t)((j & 3) * 2), idn, j ? 1u : 0u);
      }
      commit(mybar);
      mbar_wait(mybar, ph);
      ph ^= 1;
      const long long t2 = clock64();
      if (r > 0) total += t2 - t0;
    }
    if ((threadIdx.x & 31) == 0) out[w - 1] = total / (reps - 1);
  } else if (test >= 10 && (threadIdx.x & 31) == 0 && threadIdx.x >= 32 && threadIdx.x < 96) {
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(T), "r"(512u) : "memory");
}

int main() {
  long long* out; cudaMallocManaged(&out, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 161 * 1024);
  const char* names[13] = {"L1: 5 x ss N256", "L2: 17 x ts, one accumulator", "L2: 17 x ts, two accumulators", "L2: 17 x ss, one accumulator",
                          "2 tiles' L2 interleaved (34)", "2 tiles' L2 back to back (34)", "L2 (17) with L1 (5) interleaved", "L1 (5) then L2 (17)", "L2: 17 x ts unrolled", "L2: 17 x ts + extra commit", "2 warps: L2 | L2", "2 warps: L2 | L1", "2 warps: L1 | L1"};
  for (int test = 0; test < 13; ++test)
    for (int n : {64, 16, 128}) {
      if (n > 64 && test != 1 && test != 3) continue;
      if (test == 0 && n != 64) continue;
      k<<<1, 128, 161 * 1024>>>(out, test, n, 50);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s N=%d: %s\n", names[test], n, cudaGetErrorString(e)); return 1; }
      if (test >= 10) printf("%-34s N2=%2d: warp A start -> complete %5lld clk, warp B %5lld clk\n", names[test], n, out[0], out[1]);
      else printf("%-34s N2=%2d: issue %5lld clk, issue -> complete %5lld clk\n", names[test], n, out[0], out[1]);
    }
  return 0;
}
