// MuZeroNet recurrent inference on the 5th-generation tensor cores (HMZ_MODE_BF16).
//
// One CTA = 128 searches = one UMMA M=128 tile (cta_group::1).  The whole g + f chain of
// networks.py:96-116 (dynamics :129-138, prediction :140-150, support transform :152-189,
// normalize_h_state :191-196) runs inside the CTA:
//
//   gather parent latents -> smem A0 (bf16, K-major, SWIZZLE_128B)
//   D[0:256)   = A0  x Wg1^T      tcgen05.mma kind::f16, accumulators in TMEM
//   A1         = relu(D + b1 + W1[:,64+a])  (tcgen05.ld -> regs -> bf16 -> smem)
//   D[256:320) = A1  x Wg2^T  -> raw latent, min-max normalised -> A_raw, A_hn (smem) + HBM
//   reward / policy / value heads: D[0:256) = A_{raw|hn} x W1^T -> relu -> A1 -> D[256:..) = A1 x W2^T
//   softmax-expectation + signed-parabolic epilogues in registers
//
// Weights (216 KB bf16, pre-swizzled into the exact shared-memory image by hmz_weights_pack) are
// streamed from L2 per layer with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx)
// into two 32 KB buffers; every copy is issued as soon as the MMAs that read the buffer's previous
// content have committed, so it overlaps the epilogue of the current layer.
//
// 256 threads: thread t owns row (t & 127) — TMEM lane — and column half (t >> 7) of the wide
// epilogues.  One elected thread issues TMA and tcgen05.mma.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "hmz_net.cuh"

namespace hmz {

namespace tc {
constexpr int kM = 128;          // rows per CTA / UMMA M
constexpr int kThreads = 288;  // 8 epilogue warps + 1 control warp
constexpr uint32_t kAtomA = kM * 128;  // one K-atom (64 bf16) of a 128-row A tile: 16 KB

// byte offsets inside the tensor-core section of the weight blob (all multiples of 1024)
constexpr uint32_t kWg1 = 0, kWg2 = 32768, kWr1 = 65536, kWr2 = 98304, kWp1 = 122880, kWp2 = 155648, kWv1 = 163840,
                   kWv2 = 196608, kTables = 221184;
constexpr uint32_t kBytesW1 = 32768;                      // [256 out][64 in] bf16
constexpr uint32_t kBytesWg2 = 32768;                     // [64 out][256 in]
constexpr uint32_t kBytesW48 = 48 * 256 * 2, kBytesW16 = 16 * 256 * 2;
// float tables (element offsets from kTables)
constexpr int kBiasARow = 260;     // row pitch of the per-action bias table: 4-bank skew -> conflict-free LDS.128
constexpr int tBiasA = 0;          // [8][260]: b_g1[n] + W_g1[n][64 + a]  (the one-hot action column folded in)
constexpr int tBg2 = 2080;         // [64]
constexpr int tBr1 = 2144, tBr2 = 2400;  // [256], [48]
constexpr int tBp1 = 2448, tBp2 = 2704;  // [256], [16]
constexpr int tBv1 = 2720, tBv2 = 2976;  // [256], [48]
constexpr int kTableFloats = 3024;
constexpr uint32_t kTableBytes = kTableFloats * 4;  // 12096, multiple of 16
constexpr uint32_t kSectionBytes = kTables + kTableBytes;

// Debug timeline: block 0 records clock64() at phase boundaries when HMZ_TC_TIMELINE=1 (tools only).
__device__ unsigned long long g_timeline[96];
#define TL(slot) do { if (timeline && blockIdx.x == 0) g_timeline[slot] = clock64(); } while (0)

struct __align__(1024) Smem {
  uint8_t a0[kAtomA];       // input latent tile, later the raw (un-normalised) new latent
  uint8_t ahn[kAtomA];      // normalised new latent
  uint8_t a1[4 * kAtomA];   // hidden activations, 4 K-atoms
  uint8_t wf[32768];        // first-layer weights of the running MLP
  uint8_t ws[32768];        // second-layer weights of the running MLP
  float tables[kTableFloats];
  uint64_t bar_wf, bar_ws, bar_tab;  // TMA landed (tx-count barriers)
  uint64_t bar_a[2];                 // column-half h of the next A operand written (128 arrivals)
  uint64_t bar_d[2];                 // hidden accumulator columns [128h, 128h+128) complete (tcgen05.commit)
  uint64_t bar_s;                    // second-layer (small) accumulator complete
  uint64_t bar_lat;                  // raw + normalised latent tiles written (128 arrivals)
  uint64_t bar_fin;                  // small accumulator consumed by its epilogue (128 arrivals)
  uint32_t tmem_base;
  int action[kM];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  mbar_expect_tx(bar, bytes);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm100):
// start>>4 | LBO=1 (unused for swizzled K-major) | SBO = 1024 B between 8-row groups | version 1 | SW128
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  const uint64_t hi = 64ull | (1ull << 14) | (2ull << 29);
  return (uint64_t)(((smem_addr >> 4) & 0x3FFFu) | (1u << 16)) | (hi << 32);
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem_d .. +n) (+)= A[128 x 64 per atom] x B[n x 64 per atom]^T over K-atoms [ka0, ka1);
// A atoms are 16 KB apart, B atoms b_atom_stride bytes apart.  `first` clears the accumulator.
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, const uint8_t* a, const uint8_t* b, int ka0, int ka1,
                                           uint32_t b_atom_stride, uint32_t n, bool first) {
  const uint32_t idesc = umma_idesc(n);
  const uint32_t a0 = smem_u32(a), b0 = smem_u32(b);
  for (int ka = ka0; ka < ka1; ++ka)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)  // UMMA_K = 16 bf16 = 32 bytes inside the 128-byte swizzle row
      umma(tmem_d, umma_desc(a0 + ka * kAtomA + kk * 32), umma_desc(b0 + ka * b_atom_stride + kk * 32), idesc,
           (first && ka == ka0 && kk == 0) ? 0u : 1u);
}

// 32 lanes x 32 consecutive columns of TMEM -> 32 registers per thread (thread = lane = row).
// Issue and wait are separate so that the next chunk's load overlaps the current chunk's math.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Same wait, but naming the destination registers as in/out operands so that the compiler cannot
// schedule any use of them above the wait.
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  tmem_ld32_issue(taddr, r);
  tmem_ld_wait(r);
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// 16-byte chunk `chunk` (8 bf16) of row `row` inside a [rows][128 B] SWIZZLE_128B K-atom
__device__ __forceinline__ uint32_t sw128(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// bias + ReLU + bf16 of 32 accumulator columns starting at hidden column n0, stored into A1
__device__ __forceinline__ void hidden_chunk(Smem& s, const uint32_t (&acc)[32], int row, int n0, const float* bias) {
  uint32_t pk[16];
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 b = *reinterpret_cast<const float4*>(bias + n0 + j);  // 16-byte aligned: n0, j multiples of 4
    pk[(j >> 1)] = pack_bf16(fmaxf(__uint_as_float(acc[j]) + b.x, 0.f), fmaxf(__uint_as_float(acc[j + 1]) + b.y, 0.f));
    pk[(j >> 1) + 1] = pack_bf16(fmaxf(__uint_as_float(acc[j + 2]) + b.z, 0.f), fmaxf(__uint_as_float(acc[j + 3]) + b.w, 0.f));
  }
  uint8_t* atom = s.a1 + (n0 >> 6) * kAtomA;
  const int c0 = (n0 & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    *reinterpret_cast<uint4*>(atom + sw128(row, c0 + q)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
}

// Hidden-layer epilogue: D[128*half .. +128) -> relu(D + bias) -> bf16 -> A1 atoms 2*half, 2*half+1.
// The TMEM load of chunk c+1 is in flight while chunk c is converted.
__device__ __forceinline__ void hidden_epilogue(Smem& s, uint32_t tmem_row, int row, int half, const float* bias) {
  uint32_t va[32], vb[32];
  const int n0 = half * 128;
  tmem_ld32_issue(tmem_row + n0, va);
  tmem_ld_wait(va);
  tmem_ld32_issue(tmem_row + n0 + 32, vb);
  hidden_chunk(s, va, row, n0, bias);
  tmem_ld_wait(vb);
  tmem_ld32_issue(tmem_row + n0 + 64, va);
  hidden_chunk(s, vb, row, n0 + 32, bias);
  tmem_ld_wait(va);
  tmem_ld32_issue(tmem_row + n0 + 96, vb);
  hidden_chunk(s, va, row, n0 + 64, bias);
  tmem_ld_wait(vb);
  hidden_chunk(s, vb, row, n0 + 96, bias);
}

// softmax expectation over the 33 support logits in D[256:304) + signed parabolic (networks.py:152-189)
__device__ __forceinline__ float support_epilogue(uint32_t tmem_row, const float* bias) {
  float a[32], b[16];
  tmem_ld32(tmem_row + 256, a);
  tmem_ld16(tmem_row + 288, b);
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    a[i] += bias[i];
    mx = fmaxf(mx, a[i]);
  }
  b[0] += bias[32];
  mx = fmaxf(mx, b[0]);
  float den = 0.f, num = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float e = exp2f((a[i] - mx) * 1.4426950408889634f);
    den += e;
    num = fmaf(e, (float)(i - 16), num);
  }
  const float e = exp2f((b[0] - mx) * 1.4426950408889634f);
  den += e;
  num = fmaf(e, 16.f, num);
  return signed_parabolic(__fdividef(num, den));
}

// Warp roles: warps 0-7 = epilogue (thread -> row tid & 127, column half tid >> 7), warp 8 = control
// (one lane issues every TMA copy and every tcgen05.mma).  All hand-offs go through mbarriers:
//   bar_a[h]  epilogue -> control : column half h of the next A operand is in shared memory
//   bar_d[h]  control  -> epilogue: hidden accumulator columns [128h, 128h+128) are complete
//   bar_s     control  -> epilogue: the second-layer accumulator D[256:..) is complete
//   bar_lat / bar_fin  epilogue -> control: latent tiles written / small accumulator consumed
// so the MMA of one column half overlaps the epilogue of the other, the second-layer MMA starts as
// soon as its first two K-atoms exist, and the next head's first-layer MMA runs under the current
// head's final epilogue.
__global__ void __launch_bounds__(kThreads, 1)
net_recurrent_tc(const uint8_t* __restrict__ wsec, const void* __restrict__ lat_in, int64_t in_rows_per_item,
                 const uint16_t* __restrict__ in_row, const uint8_t* __restrict__ actions, void* lat_out,
                 int64_t out_rows_per_item, int64_t out_row, int latent_dtype, float* __restrict__ r_out,
                 float* __restrict__ p_out, float* __restrict__ v_out, int64_t n, int timeline) {
  extern __shared__ uint8_t smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int64_t row0 = (int64_t)blockIdx.x * kM;

  if (tid == 0) {
    mbar_init(&s.bar_wf, 1);
    mbar_init(&s.bar_ws, 1);
    mbar_init(&s.bar_tab, 1);
    mbar_init(&s.bar_a[0], 128);
    mbar_init(&s.bar_a[1], 128);
    mbar_init(&s.bar_d[0], 1);
    mbar_init(&s.bar_d[1], 1);
    mbar_init(&s.bar_s, 1);
    mbar_init(&s.bar_lat, 128);
    mbar_init(&s.bar_fin, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {  // TMEM: all 512 columns (one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s.tmem_base;

  if (warp == 8) {
    // ================================= control warp =================================
    if ((tid & 31) == 0) {
      uint32_t ph_wf = 0, ph_ws = 0, ph_a = 0, ph_d = 0, ph_s = 0, ph_fin = 0;
      TL(0);
      tma_load(s.tables, wsec + kTables, kTableBytes, &s.bar_tab);
      tma_load(s.wf, wsec + kWg1, kBytesW1, &s.bar_wf);
      tma_load(s.ws, wsec + kWg2, kBytesWg2, &s.bar_ws);
      // dynamics layer 1: D[0:256) = A0 x Wg1^T, issued as two N = 128 halves
      mbar_wait(&s.bar_a[0], ph_a);
      mbar_wait(&s.bar_a[1], ph_a);
      ph_a ^= 1;
      TL(1);  // A0 gathered
      mbar_wait(&s.bar_wf, ph_wf);
      ph_wf ^= 1;
      TL(2);  // Wg1 landed
      tc_fence_after();
      issue_gemm(tmem, s.a0, s.wf, 0, 1, 0, 128, true);
      umma_commit(&s.bar_d[0]);
      issue_gemm(tmem + 128, s.a0, s.wf + 16384, 0, 1, 0, 128, true);
      umma_commit(&s.bar_d[1]);
      TL(3);  // L1 issued
      mbar_wait(&s.bar_d[1], ph_d);
      ph_d ^= 1;
      TL(4);  // L1 complete
      tma_load(s.wf, wsec + kWr1, kBytesW1, &s.bar_wf);  // wf is free again: prefetch the reward head
      // dynamics layer 2: D[256:320) = A1 x Wg2^T, K-atoms consumed as the epilogue halves deliver them
      mbar_wait(&s.bar_ws, ph_ws);
      ph_ws ^= 1;
      TL(5);  // Wg2 landed
      mbar_wait(&s.bar_a[0], ph_a);
      TL(6);  // g-hidden half 0 written
      tc_fence_after();
      issue_gemm(tmem + 256, s.a1, s.ws, 0, 2, 64 * 128, 64, true);
      mbar_wait(&s.bar_a[1], ph_a);
      ph_a ^= 1;
      TL(7);  // g-hidden half 1 written
      tc_fence_after();
      issue_gemm(tmem + 256, s.a1, s.ws, 2, 4, 64 * 128, 64, false);
      umma_commit(&s.bar_s);
      mbar_wait(&s.bar_s, ph_s);
      ph_s ^= 1;
      TL(8);  // L2 complete
      tma_load(s.ws, wsec + kWr2, kBytesW48, &s.bar_ws);
      mbar_wait(&s.bar_lat, 0);  // raw + normalised latent tiles are in shared memory
      TL(9);  // E2 done
#pragma unroll 1
      for (int head = 0; head < 3; ++head) {
        const uint8_t* a_in = head == 0 ? s.a0 : s.ahn;
        const uint32_t n2 = head == 1 ? 16u : 48u;
        mbar_wait(&s.bar_wf, ph_wf);
        ph_wf ^= 1;
        TL(10 + head * 6);  // head first-layer weights landed
        tc_fence_after();
        issue_gemm(tmem, a_in, s.wf, 0, 1, 0, 128, true);
        umma_commit(&s.bar_d[0]);
        issue_gemm(tmem + 128, a_in, s.wf + 16384, 0, 1, 0, 128, true);
        umma_commit(&s.bar_d[1]);
        mbar_wait(&s.bar_d[1], ph_d);
        ph_d ^= 1;
        TL(11 + head * 6);  // head first-layer MMA complete
        if (head < 2) tma_load(s.wf, wsec + (head == 0 ? kWp1 : kWv1), kBytesW1, &s.bar_wf);
        mbar_wait(&s.bar_ws, ph_ws);
        ph_ws ^= 1;
        TL(12 + head * 6);  // head second-layer weights landed
        if (head > 0) {  // D[256:..) must have been drained by the previous head's final epilogue
          mbar_wait(&s.bar_fin, ph_fin);
          ph_fin ^= 1;
        }
        mbar_wait(&s.bar_a[0], ph_a);
        TL(13 + head * 6);  // head hidden half 0 written
        tc_fence_after();
        issue_gemm(tmem + 256, s.a1, s.ws, 0, 2, n2 * 128, n2, true);
        mbar_wait(&s.bar_a[1], ph_a);
        ph_a ^= 1;
        TL(14 + head * 6);  // head hidden half 1 written
        tc_fence_after();
        issue_gemm(tmem + 256, s.a1, s.ws, 2, 4, n2 * 128, n2, false);
        umma_commit(&s.bar_s);
        mbar_wait(&s.bar_s, ph_s);
        ph_s ^= 1;
        TL(15 + head * 6);  // head second-layer MMA complete
        if (head < 2) tma_load(s.ws, wsec + (head == 0 ? kWp2 : kWv2), head == 0 ? kBytesW16 : kBytesW48, &s.bar_ws);
      }
    }
  } else {
    // ================================ epilogue warps ================================
    const int row = tid & 127, half = tid >> 7;
    const uint32_t tmem_row = tmem + ((uint32_t)((warp & 3) * 32) << 16);  // this warp's TMEM lane quarter
    const int64_t item = row0 + row;
    uint32_t ph_d = 0, ph_s = 0;
    {  // gather the parent latent: thread (row, half) moves 32 of the row's 64 values
      const int64_t it = item < n ? item : n - 1;
      const int64_t irow = it * in_rows_per_item + (in_row ? (int64_t)in_row[it] : 0);
      uint32_t pk[16];
      if (latent_dtype == HMZ_LATENT_F32) {
        const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(lat_in) + irow * kLatent + half * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 t = __ldcs(src + q);  // streaming: do not displace the tree records in L2
          pk[2 * q] = pack_bf16(t.x, t.y);
          pk[2 * q + 1] = pack_bf16(t.z, t.w);
        }
      } else {
        const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(lat_in) + irow * kLatent + half * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 t = __ldcs(src + q);
          pk[4 * q] = t.x; pk[4 * q + 1] = t.y; pk[4 * q + 2] = t.z; pk[4 * q + 3] = t.w;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(s.a0 + sw128(row, half * 4 + q)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      mbar_arrive(&s.bar_a[half]);
      if (tid == 0) TL(32);  // gather done
    }
    int act = actions[item < n ? item : n - 1];
    act = act < kActions ? act : kActions - 1;
    mbar_wait(&s.bar_tab, 0);

    // ---- dynamics hidden layer: relu(D + b1 + W1[:, 64 + a]) -> A1
    mbar_wait(&s.bar_d[half], ph_d);
    ph_d ^= 1;
    tc_fence_after();
    if (tid == 0) TL(33);  // saw L1 half 0
    hidden_epilogue(s, tmem_row, row, half, s.tables + tBiasA + act * kBiasARow);
    if (tid == 0) TL(34);  // hidden epilogue math done
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(&s.bar_a[half]);
    if (tid == 0) TL(35);  // fenced + arrived

    // ---- new latent: normalize_h_state (networks.py:191-196) and its three copies
    mbar_wait(&s.bar_s, ph_s);
    ph_s ^= 1;
    tc_fence_after();
    if (tid == 0) TL(36);  // saw L2
    if (half == 0) {
      float raw[64];
      {
        float t[32];
        tmem_ld32(tmem_row + 256, t);
#pragma unroll
        for (int i = 0; i < 32; ++i) raw[i] = t[i] + s.tables[tBg2 + i];
        tmem_ld32(tmem_row + 288, t);
#pragma unroll
        for (int i = 0; i < 32; ++i) raw[32 + i] = t[i] + s.tables[tBg2 + 32 + i];
      }
      float mn = raw[0], mx = raw[0];
#pragma unroll
      for (int i = 1; i < 64; ++i) {
        mn = fminf(mn, raw[i]);
        mx = fmaxf(mx, raw[i]);
      }
      const float inv = 1.0f / ((mx - mn) + 1e-8f);
      const int64_t orow = item * out_rows_per_item + out_row;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float hn[8];
        uint32_t pr[4], ph[4];
#pragma unroll
        for (int j = 0; j < 8; ++j) hn[j] = (raw[c * 8 + j] - mn) * inv;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          pr[j] = pack_bf16(raw[c * 8 + 2 * j], raw[c * 8 + 2 * j + 1]);
          ph[j] = pack_bf16(hn[2 * j], hn[2 * j + 1]);
        }
        *reinterpret_cast<uint4*>(s.a0 + sw128(row, c)) = make_uint4(pr[0], pr[1], pr[2], pr[3]);
        *reinterpret_cast<uint4*>(s.ahn + sw128(row, c)) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        if (item < n) {
          if (latent_dtype == HMZ_LATENT_F32) {
            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(lat_out) + orow * kLatent + c * 8);
            __stcs(dst, make_float4(hn[0], hn[1], hn[2], hn[3]));
            __stcs(dst + 1, make_float4(hn[4], hn[5], hn[6], hn[7]));
          } else {
            __stcs(reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(lat_out) + orow * kLatent + c * 8),
                   make_uint4(ph[0], ph[1], ph[2], ph[3]));
          }
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&s.bar_lat);
      if (tid == 0) TL(37);  // E2 done
    }

    // ---- the three heads: {reward on the raw latent, policy and value on the normalised latent}
#pragma unroll 1
    for (int head = 0; head < 3; ++head) {
      const float* b1 = s.tables + (head == 0 ? tBr1 : (head == 1 ? tBp1 : tBv1));
      const float* b2 = s.tables + (head == 0 ? tBr2 : (head == 1 ? tBp2 : tBv2));
      mbar_wait(&s.bar_d[half], ph_d);
      ph_d ^= 1;
      tc_fence_after();
      if (tid == 0) TL(38 + head * 4);  // saw head first layer
      // A1 is rewritten here: the previous second-layer MMA (the last reader of A1) has completed —
      // both halves observed bar_s for it (dynamics: above; heads: at the end of the previous iteration).
      hidden_epilogue(s, tmem_row, row, half, b1);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&s.bar_a[half]);
      if (tid == 0) TL(39 + head * 4);  // head hidden epilogue done
      mbar_wait(&s.bar_s, ph_s);
      ph_s ^= 1;
      tc_fence_after();
      if (tid == 0) TL(40 + head * 4);  // saw head second layer
      if (half == 0) {
        if (head == 1) {  // F.softmax(pi_logits) (networks.py:109)
          float lg[16];
          tmem_ld16(tmem_row + 256, lg);
          float mx = -INFINITY, den = 0.f;
#pragma unroll
          for (int a = 0; a < kActions; ++a) {
            lg[a] += b2[a];
            mx = fmaxf(mx, lg[a]);
          }
#pragma unroll
          for (int a = 0; a < kActions; ++a) {
            lg[a] = exp2f((lg[a] - mx) * 1.4426950408889634f);
            den += lg[a];
          }
          const float inv = 1.0f / den;
          if (item < n) {
#pragma unroll
            for (int a = 0; a < kActions; ++a) p_out[item * kActions + a] = lg[a] * inv;
          }
        } else {
          const float x = support_epilogue(tmem_row, b2);
          if (item < n) (head == 0 ? r_out : v_out)[item] = x;
        }
        tc_fence_before();
        if (head < 2) mbar_arrive(&s.bar_fin);
        if (tid == 0) TL(41 + head * 4);  // head final epilogue done
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (tid == 0) TL(63);
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ---- host-side packing ----------------------------------------------------------------
static uint16_t f2bf(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((u >> 16) | 0x40u);  // NaN
  return (uint16_t)((u + 0x7FFFu + ((u >> 16) & 1u)) >> 16);
}

// torch Linear weight w [out][in] (row-major) restricted to input columns [k0, k0 + 64*k_atoms) ->
// k_atoms SWIZZLE_128B K-atoms of [out_pad rows][128 B]; rows >= out are zero.
static void pack_kmajor_sw128(uint8_t* dst, const float* w, int out, int in, int out_pad, int k0, int k_atoms) {
  for (int ka = 0; ka < k_atoms; ++ka)
    for (int n = 0; n < out_pad; ++n)
      for (int c = 0; c < 8; ++c) {
        uint16_t* chunk = reinterpret_cast<uint16_t*>(dst + (size_t)ka * out_pad * 128 + (size_t)n * 128 + ((c ^ (n & 7)) << 4));
        for (int j = 0; j < 8; ++j) {
          const int k = k0 + ka * 64 + c * 8 + j;
          chunk[j] = (n < out && k < in) ? f2bf(w[(size_t)n * in + k]) : (uint16_t)0;
        }
      }
}
}  // namespace tc

// bf16 blob = [tensor-core section, fixed size][float32 blob of hmz_net.cu]: the section comes first
// so that the recurrent path needs no disk count to find it.
int64_t tc_fp32_offset_bytes() { return ((int64_t)tc::kSectionBytes + 1023) / 1024 * 1024; }

int64_t tc_packed_bytes(int n_disks) { return tc_fp32_offset_bytes() + (int64_t)Fp32Layout::total(n_disks) * 4; }

void tc_pack(const float* const* t, int n_disks, void* out) {
  using namespace tc;
  std::memset(out, 0, (size_t)tc_packed_bytes(n_disks));
  // the root inference (1/S of the work) runs on the float32 copy behind the section
  pack_fp32(t, n_disks, (float*)((uint8_t*)out + tc_fp32_offset_bytes()));
  uint8_t* sec = (uint8_t*)out;
  // state_dict order: rep(0-3) dyn(4-7) rwd(8-11) pol(12-15) val(16-19); each {w1, b1, w2, b2}
  pack_kmajor_sw128(sec + kWg1, t[4], kHidden, kLatent + kActions, 256, 0, 1);
  pack_kmajor_sw128(sec + kWg2, t[6], kLatent, kHidden, 64, 0, 4);
  pack_kmajor_sw128(sec + kWr1, t[8], kHidden, kLatent, 256, 0, 1);
  pack_kmajor_sw128(sec + kWr2, t[10], kSupport, kHidden, 48, 0, 4);
  pack_kmajor_sw128(sec + kWp1, t[12], kHidden, kLatent, 256, 0, 1);
  pack_kmajor_sw128(sec + kWp2, t[14], kActions, kHidden, 16, 0, 4);
  pack_kmajor_sw128(sec + kWv1, t[16], kHidden, kLatent, 256, 0, 1);
  pack_kmajor_sw128(sec + kWv2, t[18], kSupport, kHidden, 48, 0, 4);
  float* tab = reinterpret_cast<float*>(sec + kTables);
  const int in_g1 = kLatent + kActions;
  for (int a = 0; a < 8; ++a)
    for (int n = 0; n < kHidden; ++n)
      tab[tBiasA + a * kBiasARow + n] = t[5][n] + (a < kActions ? t[4][(size_t)n * in_g1 + kLatent + a] : 0.f);
  for (int i = 0; i < kLatent; ++i) tab[tBg2 + i] = t[7][i];
  for (int i = 0; i < kHidden; ++i) {
    tab[tBr1 + i] = t[9][i];
    tab[tBp1 + i] = t[13][i];
    tab[tBv1 + i] = t[17][i];
  }
  for (int i = 0; i < kSupport; ++i) {
    tab[tBr2 + i] = t[11][i];
    tab[tBv2 + i] = t[19][i];
  }
  for (int i = 0; i < kActions; ++i) tab[tBp2 + i] = t[15][i];
}

static int tc_timeline_enabled() {
  static const int on = getenv("HMZ_TC_TIMELINE") ? atoi(getenv("HMZ_TC_TIMELINE")) : 0;
  return on;
}

int tc_debug_read_timeline(unsigned long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, tc::g_timeline, sizeof(unsigned long long) * 96) == cudaSuccess ? HMZ_OK : HMZ_ERR_CUDA;
}

int tc_net_recurrent(const void* weights, const void* lat_in, int64_t in_rows_per_item, const uint16_t* in_row,
                     const uint8_t* actions, void* lat_out, int64_t out_rows_per_item, int64_t out_row,
                     int latent_dtype, float* r, float* p, float* v, int64_t n, cudaStream_t stream) {
  static thread_local int done_dev = -1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(HMZ_ERR_CUDA, "cudaGetDevice failed (no CUDA device?)");
  const int smem_bytes = (int)sizeof(tc::Smem) + 1024;
  if (done_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(tc::net_recurrent_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return fail(HMZ_ERR_CUDA, "cudaFuncSetAttribute(net_recurrent_tc): %s", cudaGetErrorString(e));
    done_dev = dev;
  }
  const unsigned grid = (unsigned)((n + tc::kM - 1) / tc::kM);
  tc::net_recurrent_tc<<<grid, tc::kThreads, smem_bytes, stream>>>((const uint8_t*)weights, lat_in, in_rows_per_item, in_row,
                                                                   actions, lat_out, out_rows_per_item, out_row,
                                                                   latent_dtype, r, p, v, n, tc_timeline_enabled());
  return check_launch("net_recurrent_tc");
}

}  // namespace hmz
