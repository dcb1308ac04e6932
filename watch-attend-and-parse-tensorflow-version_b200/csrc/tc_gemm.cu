// tcgen05 TF32 GEMM / 3x3 implicit-GEMM kernels (see tc_gemm.cuh for the design).
#include "tc_gemm.cuh"

#include <cudaTypedefs.h>
#include <cuda_bf16.h>

#include <algorithm>

namespace wap {

constexpr int kTcBM = 128;          // UMMA M (cta_group::1)
#ifndef WAP_TC_BK
#define WAP_TC_BK 32
#endif
constexpr int kTcBK = WAP_TC_BK;    // floats per stage row: 32 (128 B rows, SWIZZLE_128B, 1 CTA/SM in 3xTF32 mode) or
                                    // 16 (64 B rows, SWIZZLE_64B: half-size stages so that TWO CTAs stay resident per SM)
constexpr int kTcRowB = kTcBK * 4;
static_assert(kTcBK == 32 || kTcBK == 16, "WAP_TC_BK must be 16 or 32");
constexpr int kTcThreads = 192;
constexpr uint32_t kSpinLimit = 1u << 24;

// ------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a pipeline bug traps (launch failure) instead of hanging the GPU.  Before trapping the waiter leaves a note in
// host-mapped memory (wap_debug_trap_info): {site, blockIdx.x, blockIdx.y, warp, parity, gridDim.x, gridDim.y, blockDim.x}.
__device__ int* g_trap_note = nullptr;
__device__ __noinline__ void mbar_timeout(int site, uint32_t parity) {
  if (g_trap_note != nullptr) {
    const int slot = atomicAdd_system(g_trap_note + 1, 1);            // [0] unused, [1] count, then 8 ints per note
    if (slot < 30) {
      volatile int* n = g_trap_note + 8 + 8 * slot;
      n[1] = blockIdx.x; n[2] = blockIdx.y; n[3] = threadIdx.x >> 5; n[4] = (int)parity; n[5] = gridDim.x; n[6] = gridDim.y; n[7] = blockDim.x;
      n[0] = site;
    }
    __threadfence_system();
    for (volatile int d = 0; d < 2000000; ++d) {}                      // let the other stuck waiters leave their notes too
  }
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int site = 0) {
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) {
#ifdef WAP_TRAP_NOTES          // debug builds only (build.sh -DWAP_TRAP_NOTES): the note call costs 0.66 ms per training step
    if (++spins > kSpinLimit) mbar_timeout(site, parity);
#else
    if (++spins > kSpinLimit) __trap();
#endif
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]   (kind::tf32, issued by ONE thread)
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]   (kind::f16 with bf16 operands, fp32 accumulate; K = 16 per instruction)
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem: lane = row, column = k] * B[smem desc]   (A operand staged in tensor memory)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 registers per thread -> 32 lanes x 32 consecutive fp32 columns (thread i = TMEM lane base+i)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
        "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
        "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
        "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// arrives on the mbarrier once all previously issued MMAs have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i = TMEM lane base+i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor: K-major, rows of kTcRowB bytes = one swizzle span (SWIZZLE_128B or SWIZZLE_64B),
// 8-row groups 8*kTcRowB bytes apart
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);       // start address        bits [0,14)
  d |= (uint64_t)0 << 16;                            // leading byte offset  (unused for swizzled K-major)
  d |= (uint64_t)((8 * kTcRowB) >> 4) << 32;         // stride byte offset   bits [32,46)
  d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
  d |= (uint64_t)(kTcRowB == 128 ? 2 : 4) << 61;     // layout: SWIZZLE_128B / SWIZZLE_64B
  return d;
}
// instruction descriptor: D=f32, A=B=tf32, both K-major, M=128, N=BN
__host__ __device__ constexpr uint32_t make_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
}

// instruction descriptor: D=f32, A=B=bf16 (kind::f16 format 1), both K-major, M=128, N=BN
__host__ __device__ constexpr uint32_t make_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
}
// 8 fp32 -> 8 bf16 (round to nearest even) as one 16-byte store
// (two 8-byte stores: rows of 300 bf16 channels are only 8-byte aligned)
__device__ __forceinline__ void store_bf16x8(void* dst, const float* v) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[i]) : "f"(v[2 * i + 1]), "f"(v[2 * i]));
  reinterpret_cast<uint2*>(dst)[0] = make_uint2(w[0], w[1]);
  reinterpret_cast<uint2*>(dst)[1] = make_uint2(w[2], w[3]);
}

// ------------------------------------------------------------------------------------ kernel
struct TcShape {
  int M, N, K;            // logical GEMM (K = total reduction length; for conv: 9*C)
  int nkb;                // number of 32-wide K blocks
  int kps;                // split-K: k-blocks per grid.z slice (0 = no split, every CTA runs all nkb)
  // im2col mode
  int B, H, W, kb_per_tap, TW, TH, tiles_x, tiles_y;
  long long* trace;       // debug: CTA 0 records clock64() per k-block: [kb][0 TMA issue, 1 data landed (splitter/MMA saw full), 2 split done, 3 MMA issued]
};

// X3 = error-compensated "3xTF32": D = a_hi*b_hi + a_lo*b_hi + a_hi*b_lo (fp32-grade accuracy).
// The three MMAs of a k-block would read A twice and B three times from shared memory, and the hi/lo split of the raw fp32
// A tile would add a read and two writes: 172 KB of shared-memory traffic per 128x96x32 k-block, which is what bounded the
// X3 kernels (~1390 cycles per k-block at 128 B/clk).  So the splitter warps write a_hi / a_lo into TENSOR MEMORY (lane =
// tile row, 32 columns each) and the MMAs take A from TMEM: shared memory then only sees the TMA write, one read of A by the
// splitter and the B reads.  Stage layout in smem: [A raw] [B hi] [B lo] (B lo only when X3).
template <int BN, bool X3, bool BF16 = false> struct TcSmem {
  // These kernels are bound by bytes in flight per SM (Little's law over the ~1.5 us L2->smem TMA latency), so the X3 variants
  // (one CTA per SM) take as many stages as the 227 KB of shared memory hold; single-pass variants are sized for 2 CTAs/SM.
#ifndef WAP_TC_X3_DEEP
#define WAP_TC_X3_DEEP 1
#endif
  // (bf16 with SIX stages and one CTA per SM was measured in round 2: 3 % slower end to end than 3-4 stages x 2 CTAs per SM - the second CTA's
  // epilogue overlapping the first one's main loop matters more in block 0 than deeper prefetch does in blocks 1-2)
  static constexpr int kStages = kTcBK == 16 ? (X3 ? 3 : 4)
                               : (X3 ? (BN <= 96 ? 4 : 3) : (BN <= 32 ? 4 : 3));     // X3: 4 x 64 TMEM columns of A ring + 2 x 128 accumulator columns = 512
                                                                       // (six stages for the BN = 48 N-split variant were measured: no gain)
  static constexpr int kABytes = kTcBM * kTcBK * 4;                 // 16384
  static constexpr int kBBytes = BN * kTcBK * 4;
  static constexpr int kStage = kABytes + (X3 ? 2 : 1) * kBBytes;
  static constexpr int kBarOff = kStages * kStage;
  static constexpr int kStatOff = kBarOff + 256;                    // barriers + tmem slot
  static constexpr int kTotal = kStatOff + 6 * BN * 4 + 1024;       // statistics (2 BN) + fused BN-backward constants (4 BN) + alignment slack
};
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float tf32_rna_dev(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// (A second group of splitter warps taking alternate k-blocks was measured: no gain - the X3 main loop is bound by the tensor
// pipe plus the TMEM stores of the split, ~1100 cycles per k-block, see DESIGN 3.2 - and with 3 stages a group that skips
// k-blocks can alias mbarrier parities.  One group of four warps observes every phase of every stage.)
template <bool X3> struct TcThreads { static constexpr int kN = kTcThreads; };
// IM2COL: 0 plain 2-D A operand; 1 3x3 SAME im2col (9 taps in K); 2 "row tile" 3x3 (3 vertical taps in K, the horizontal
// taps are 3 x 24 output columns whose results the epilogue adds with a +-1 pixel shift, see tc_conv3x3_rowtile)
// BF16: operands are bf16 in global / shared memory (a stage row is still one 128-byte swizzle span = 64 elements), the MMAs are
// kind::f16 with K = 16 (again four per stage), accumulation fp32 in TMEM; epi.out_bf16 stores the result as bf16.
// EPI selects which epilogue features are COMPILED IN (the all-in-one epilogue was ~100 KB of SASS per instantiation; the small per-time-step
// GEMMs of the scans spent more time fetching it than executing it):  0 = store / accumulate / bias / folded inference BN,  1 = store + BN
// statistics,  2 = transposed store, split-K atomics (+ bias),  3 = statistics + fused BN-backward reduce (debug-mask bit 10).
template <int BN, int IM2COL, bool X3, bool BF16, int EPI>
__global__ void __launch_bounds__(TcThreads<X3>::kN, (X3 && kTcBK == 32) ? 1 : 2)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo,
               const TcEpi epi, const TcShape sh) {
  pdl_sync();
  const bool tr0 = sh.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  if (tr0 && threadIdx.x == 0) sh.trace[1024] = clock64();
  static_assert(!(X3 && BF16), "3xTF32 splitting applies to fp32 operands only");
  constexpr bool kEStats = EPI == 1 || EPI == 3, kEBnb = EPI == 3, kETr = EPI == 2, kEPlain = EPI == 0;
  using S = TcSmem<BN, X3, BF16>;
  constexpr int kTcStages = S::kStages;
  constexpr int kElK = BF16 ? kTcRowB / 2 : kTcBK;            // elements of K per stage row (64 bf16 or 32 fp32)
  // X3 keeps the two small correction products in a SECOND accumulator (columns [kAcc2, kAcc2+BN)) so that their
  // fp32 accumulation error is relative to their own 2^-11-times smaller magnitude; the epilogue adds the two.
  constexpr uint32_t kAcc2 = BN <= 32 ? 32 : BN <= 64 ? 64 : 128;
  constexpr uint32_t kAccCols = (X3 ? 2 : 1) * (BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256);
  constexpr uint32_t kARing = kAccCols;                           // X3: stage s keeps a_hi at kARing + 64 s, a_lo 32 columns further
  constexpr uint32_t kNeed = kAccCols + (X3 ? 64u * kTcStages : 0u);
  constexpr uint32_t kTmemCols = kNeed <= 32 ? 32 : kNeed <= 64 ? 64 : kNeed <= 128 ? 128 : kNeed <= 256 ? 256 : 512;
  static_assert(kNeed <= 512, "TMEM budget");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
  uint64_t* empty_bar = full_bar + kTcStages;
  uint64_t* split_bar = empty_bar + kTcStages;
  uint64_t* tmem_full = split_bar + kTcStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* s_stat = reinterpret_cast<float*>(smem + S::kStatOff);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile coordinates
  int m0 = 0, bimg = 0, x0 = 0, y0 = 0;
  if (IM2COL) {
    int t = blockIdx.x;
    const int tx = t % sh.tiles_x; t /= sh.tiles_x;
    const int ty = t % sh.tiles_y; bimg = t / sh.tiles_y;
    x0 = tx * sh.TW; y0 = ty * sh.TH;
  } else {
    m0 = blockIdx.x * kTcBM;
  }
  const int n0 = blockIdx.y * BN;
  const int kbb = (kETr && sh.kps > 0) ? (int)blockIdx.z * sh.kps : 0;                       // split-K: this CTA reduces k-blocks [kbb, kbb + nloc)
  const int nloc = (kETr && sh.kps > 0) ? min(sh.kps, sh.nkb - kbb) : sh.nkb;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB);
    if (X3) prefetch_tmap(&tmBlo);
    for (int s = 0; s < kTcStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&split_bar[s], 4); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (kEStats && epi.stats != nullptr)
    for (int i = threadIdx.x; i < 2 * BN; i += blockDim.x) s_stat[i] = 0.f;
  if (kEBnb && epi.bnb_raw != nullptr)       // fused BN-backward reduce: scale | shift | mean | invstd of this tile's channels behind the two sums
    for (int i = threadIdx.x; i < BN; i += blockDim.x) {
      const bool ok = n0 + i < sh.N;
      const float mean = ok ? epi.bnb_saved[n0 + i] : 0.f, inv = ok ? epi.bnb_saved[kMaxBnC + n0 + i] : 0.f;
      const float sc = ok ? __ldg(epi.bnb_gamma + n0 + i) * inv : 0.f;
      s_stat[2 * BN + i] = sc; s_stat[3 * BN + i] = ok ? __ldg(epi.bnb_beta + n0 + i) - mean * sc : 0.f;
      s_stat[4 * BN + i] = mean; s_stat[5 * BN + i] = inv;
    }
  if ((kEPlain || kETr) && epi.bias != nullptr)      // bias row of this tile's columns (zero beyond N)
    for (int i = threadIdx.x; i < BN; i += blockDim.x) s_stat[2 * BN + i] = (n0 + i < sh.N) ? __ldg(epi.bias + n0 + i) : 0.f;
  if (kEPlain && epi.bn_scale != nullptr)      // folded inference batch norm: scale | shift of this tile's columns
    for (int i = threadIdx.x; i < BN; i += blockDim.x) {
      const bool ok = n0 + i < sh.N;
      s_stat[i] = ok ? __ldg(epi.bn_scale + n0 + i) : 0.f; s_stat[BN + i] = ok ? __ldg(epi.bn_shift + n0 + i) : 0.f;
    }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tr0 && threadIdx.x == 0) sh.trace[1025] = clock64();

  if (warp == 0) {
    // ================================================= TMA producer
    if (lane == 0) {
      for (int it = 0; it < nloc; ++it) {
        const int kb = kbb + it;
        const int s = it % kTcStages;
        const uint32_t ph = (it / kTcStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1, 1 + 100 * kb);
        if (sh.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0) sh.trace[kb * 4 + 0] = clock64();
        uint8_t* sa = smem + s * S::kStage;
        uint8_t* sb = sa + S::kABytes;
        mbar_expect_tx(&full_bar[s], S::kABytes + (X3 ? 2 : 1) * S::kBBytes);
        if (IM2COL) {
          const int tap = kb / sh.kb_per_tap, cb = kb - tap * sh.kb_per_tap;
          if (IM2COL == 2) tma_load_4d(sa, &tmA, &full_bar[s], cb * kElK, x0, y0 + tap - 1, bimg);
          else tma_load_4d(sa, &tmA, &full_bar[s], cb * kElK, x0 + tap % 3 - 1, y0 + tap / 3 - 1, bimg);
        } else {
          tma_load_2d(sa, &tmA, &full_bar[s], kb * kElK, m0);
        }
        tma_load_2d(sb, &tmB, &full_bar[s], kb * kElK, n0);
        if (X3) tma_load_2d(sb + S::kBBytes, &tmBlo, &full_bar[s], kb * kElK, n0);
      }
    }
  } else if (warp == 1) {
    // ================================================= MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = BF16 ? make_idesc_bf16(BN) : make_idesc_tf32(BN);
      for (int it = 0; it < nloc; ++it) {
        const int kb = it;                   // (only "first k-block of this CTA" matters below)
        const int s = it % kTcStages;
        const uint32_t ph = (it / kTcStages) & 1;
        mbar_wait(X3 ? &split_bar[s] : &full_bar[s], ph, 2 + 100 * kb);
        tc_fence_after();
        if (sh.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0) sh.trace[kb * 4 + 3] = clock64();
        const uint32_t sa = smem_u32(smem + s * S::kStage);
        const uint32_t sb = sa + S::kABytes;
#pragma unroll
        for (int k = 0; k < kTcBK / 8; ++k) {
          const uint64_t bd = make_kmajor_sw128_desc(sb + k * 32);
          if (X3) {
            const uint32_t ah = tmem_base + kARing + (uint32_t)s * 64u + (uint32_t)k * 8u, al = ah + 32u;
            const uint64_t bld = make_kmajor_sw128_desc(sb + S::kBBytes + k * 32);
            umma_tf32_ts(tmem_base, ah, bd, idesc, (kb | k) != 0 ? 1u : 0u);            // a_hi * b_hi
            umma_tf32_ts(tmem_base + kAcc2, al, bd, idesc, (kb | k) != 0 ? 1u : 0u);    // a_lo * b_hi
            umma_tf32_ts(tmem_base + kAcc2, ah, bld, idesc, 1u);                        // a_hi * b_lo
          } else {
            const uint64_t ad = make_kmajor_sw128_desc(sa + k * 32);
            if (BF16) umma_bf16(tmem_base, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_tf32(tmem_base, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[s]);          // frees the smem slot once these MMAs have read it
      }
      umma_commit(tmem_full);                // accumulator complete
    }
  } else {
    // ================================================= epilogue (warps 2..5)
    if (X3) {
      // ---- operand splitter (main loop): thread = tile row.  Reads its 128-byte row of the raw fp32 A tile (undoing the
      // 128B swizzle: 16-byte chunk c of row r sits at chunk c ^ (r & 7)), splits it into hi = rna_tf32(a) and lo = a - hi
      // and stores both into the TMEM ring of this stage; the MMAs take A from there.
      const int qq = warp & 3, rr = qq * 32 + lane;
      for (int kb = 0; kb < nloc; ++kb) {
        const int s = kb % kTcStages;
        const uint32_t ph = (kb / kTcStages) & 1;
        mbar_wait(&full_bar[s], ph, 3 + 100 * kb);
        if (sh.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 127) == 64) sh.trace[kb * 4 + 1] = clock64();
        const float4* row = reinterpret_cast<const float4*>(smem + s * S::kStage + rr * 128);
        float hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = row[c ^ (rr & 7)];
          hi[4 * c + 0] = tf32_rna_dev(v.x); hi[4 * c + 1] = tf32_rna_dev(v.y); hi[4 * c + 2] = tf32_rna_dev(v.z); hi[4 * c + 3] = tf32_rna_dev(v.w);
          lo[4 * c + 0] = v.x - hi[4 * c + 0]; lo[4 * c + 1] = v.y - hi[4 * c + 1]; lo[4 * c + 2] = v.z - hi[4 * c + 2]; lo[4 * c + 3] = v.w - hi[4 * c + 3];
        }
        const uint32_t ta = tmem_base + ((uint32_t)(qq * 32) << 16) + kARing + (uint32_t)s * 64u;
        tmem_st32(ta, hi);
        tmem_st32(ta + 32u, lo);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&split_bar[s]);
        if (sh.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 127) == 64) sh.trace[kb * 4 + 2] = clock64();
      }
    }
    if (warp >= 6) goto tail;                // (no such warps with 192 threads; kept for experiments with extra splitter warps)
    {
    const int q = warp & 3;                  // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;             // output row within the tile
    bool valid; size_t mrow;
    if (IM2COL) {
      const int tw = r % sh.TW, th = r / sh.TW;
      const int x = x0 + tw, y = y0 + th;
      valid = (x < sh.W) && (y < sh.H);
      mrow = ((size_t)bimg * sh.H + y) * sh.W + x;
    } else {
      valid = (m0 + r) < sh.M;
      mrow = (size_t)(m0 + r);
    }
    mbar_wait(tmem_full, 0, 4);
    tc_fence_after();
    if (tr0 && threadIdx.x == 64) sh.trace[1026] = clock64();
    if constexpr (IM2COL == 2) {
      // Row-tile 3x3 epilogue.  Accumulator column 24 j + c of tile row q holds the contribution of INPUT pixel q through the
      // horizontal tap dx = j - 1 (all three vertical taps already summed by the K loop), i.e. it belongs to OUTPUT pixel
      // q - dx:  D[p, c] = acc[p-1, c] + acc[p, 24 + c] + acc[p+1, 48 + c], neighbours taken within the image row (the tile is
      // TW = pow2 >= W pixels wide, so a row never continues in another tile; columns >= W are zero A rows).
      constexpr int kG = 24;
      float v[96];
#pragma unroll
      for (int c = 0; c < 96; c += 32) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v + c);
        if (X3) {
          float v2[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + kAcc2 + (uint32_t)c, v2);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[c + i] += v2[i];
        }
      }
      float* s_bnd = reinterpret_cast<float*>(smem);          // [4 warps][2][24]: pipeline buffers are free once tmem_full has fired
      if (lane == 31) {
#pragma unroll
        for (int c = 0; c < kG; ++c) s_bnd[(q * 2 + 0) * kG + c] = v[c];              // slot 0 of my last row -> first row of the next warp
      }
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < kG; ++c) s_bnd[(q * 2 + 1) * kG + c] = v[2 * kG + c];     // slot 2 of my first row -> last row of the previous warp
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int tw = r % sh.TW;
      float o[32];
#pragma unroll
      for (int c = 0; c < kG; ++c) {
        float left = __shfl_up_sync(0xffffffffu, v[c], 1);             // acc[p-1, slot 0]
        float right = __shfl_down_sync(0xffffffffu, v[2 * kG + c], 1); // acc[p+1, slot 2]
        if (lane == 0) left = q > 0 ? s_bnd[((q - 1) * 2 + 0) * kG + c] : 0.f;
        if (lane == 31) right = q < 3 ? s_bnd[((q + 1) * 2 + 1) * kG + c] : 0.f;
        if (tw == 0) left = 0.f;
        if (tw == sh.TW - 1) right = 0.f;
        o[c] = left + v[kG + c] + right;
      }
#pragma unroll
      for (int c = kG; c < 32; ++c) o[c] = 0.f;
      if (kEPlain && epi.bn_scale != nullptr) {
#pragma unroll
        for (int c = 0; c < kG; ++c) o[c] = fmaxf(fmaf(o[c], s_stat[c], s_stat[BN + c]), 0.f);
      }
      if (valid) {
        if (epi.out_bf16) {
          uint8_t* dst = reinterpret_cast<uint8_t*>(epi.out) + (mrow * epi.ld) * 2;
#pragma unroll
          for (int j = 0; j < kG; j += 8) store_bf16x8(dst + j * 2, o + j);
        } else {
          float* dst = epi.out + mrow * epi.ld;
#pragma unroll
          for (int j = 0; j < kG; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
        }
      }
      if (kEStats && epi.stats != nullptr) {
        float a[32], b[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) { a[i] = valid ? o[i] : 0.f; b[i] = a[i] * a[i]; }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
          const bool upper = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < off; ++i) {
            const float sa_ = upper ? a[i] : a[i + off], ka = upper ? a[i + off] : a[i];
            const float sb_ = upper ? b[i] : b[i + off], kb_ = upper ? b[i + off] : b[i];
            a[i] = ka + __shfl_xor_sync(0xffffffffu, sa_, off);
            b[i] = kb_ + __shfl_xor_sync(0xffffffffu, sb_, off);
          }
        }
        if (lane < kG) { atomicAdd(&s_stat[lane], a[0]); atomicAdd(&s_stat[BN + lane], b[0]); }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int t = threadIdx.x - 64;
        if (t < kG) {
          atomicAdd(epi.stats + t, (double)s_stat[t]);
          atomicAdd(epi.stats + sh.N + t, (double)s_stat[BN + t]);
        }
      }
    } else {
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      if (tr0 && threadIdx.x == 64) sh.trace[1028 + c / 32] = clock64();
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      if (X3) {
        float v2[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + kAcc2 + (uint32_t)c, v2);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += v2[i];
      }
      const int lim = min(sh.N - n0, BN) - c;          // columns of this chunk that belong to this tile AND to the matrix (BN = 48: the second chunk is half)
      const bool full = lim >= 32;                   // whole chunk valid: straight-line stores (the guarded forms cost a branch per group)
      if ((kEPlain || kETr) && epi.bias != nullptr && kbb == 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += s_stat[2 * BN + c + j];
      }
      if (kEPlain && epi.bn_scale != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(fmaf(v[j], s_stat[c + j], s_stat[BN + c + j]), 0.f);
      }
      float xh[32];                      // (fused BN backward only) normalised raw output of the element, for the second sum
      if (kEBnb && epi.bnb_raw != nullptr) {
        const uint32_t seed = (epi.bnb_dropout && epi.bnb_seed_src != nullptr) ? __ldg(epi.bnb_seed_src) : epi.bnb_seed;
        const float* rawp = epi.bnb_raw + mrow * epi.bnb_ld + n0 + c;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid && n0 + c + j + 3 < sh.N) rv = __ldg(reinterpret_cast<const float4*>(rawp + j));
          const float rr[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int i = c + j + jj;
            const float pre = fmaf(rr[jj], s_stat[2 * BN + i], s_stat[3 * BN + i]);
            float g = (pre > 0.f) ? v[j + jj] : 0.f;
            if (epi.bnb_dropout) g = drop_keep(seed, epi.bnb_site, (uint32_t)(mrow * (size_t)sh.N + n0 + i)) ? g * (1.f / (1.f - kDropRate)) : 0.f;
            v[j + jj] = g;
            xh[j + jj] = (rr[jj] - s_stat[4 * BN + i]) * s_stat[5 * BN + i];
          }
        }
      }
      if (kETr && valid && sh.kps > 0) {                   // split-K: partial sums added into the zero-filled (or pre-loaded) output
        float* dst = epi.transpose_out ? epi.out + (size_t)(n0 + c) * epi.ld + mrow : epi.out + mrow * epi.ld + n0 + c;
        const size_t step = epi.transpose_out ? (size_t)epi.ld : 1;
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(dst + j * step, v[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (j < lim) atomicAdd(dst + j * step, v[j]);
        }
      } else if (kETr && valid && epi.transpose_out) {
        float* dst = epi.out + (size_t)(n0 + c) * epi.ld + mrow;
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[(size_t)j * epi.ld] = v[j];
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (j < lim) dst[(size_t)j * epi.ld] = v[j];
        }
      } else if (valid && epi.out_bf16) {         // bf16 result (no accumulate): N is a multiple of 8 for every caller
        uint8_t* dst = reinterpret_cast<uint8_t*>(epi.out) + (mrow * epi.ld + n0 + c) * 2;
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) store_bf16x8(dst + j * 2, v + j);
        } else {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          if (j + 7 < lim) store_bf16x8(dst + j * 2, v + j);
          else {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj)
              if (j + jj < lim) reinterpret_cast<__nv_bfloat16*>(dst)[j + jj] = __float2bfloat16_rn(v[j + jj]);
          }
        }
        }
      } else if (valid) {
        float* dst = epi.out + mrow * epi.ld + n0 + c;
        if (full && !(kEPlain && epi.accumulate)) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else if (full) {
          float4 pv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) pv[j] = *reinterpret_cast<const float4*>(dst + 4 * j);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(v[4 * j] + pv[j].x, v[4 * j + 1] + pv[j].y, v[4 * j + 2] + pv[j].z, v[4 * j + 3] + pv[j].w);
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (j + 3 < lim) {
              float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              if (kEPlain && epi.accumulate) { const float4 p = *reinterpret_cast<const float4*>(dst + j); o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
              *reinterpret_cast<float4*>(dst + j) = o;
            } else {
#pragma unroll
              for (int jj = 0; jj < 4; ++jj)
                if (j + jj < lim) dst[j + jj] = (kEPlain && epi.accumulate) ? dst[j + jj] + v[j + jj] : v[j + jj];
            }
          }
        }
      }
      if (kEStats && epi.stats != nullptr) {
        float a[32], b[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) { a[i] = valid ? v[i] : 0.f; b[i] = a[i] * ((kEBnb && epi.bnb_raw != nullptr) ? xh[i] : a[i]); }
        // transposing butterfly: after the 5 rounds lane L holds the sum over the 32 rows of column L
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
          const bool upper = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < off; ++i) {
            const float sa_ = upper ? a[i] : a[i + off], ka = upper ? a[i + off] : a[i];
            const float sb_ = upper ? b[i] : b[i + off], kb_ = upper ? b[i + off] : b[i];
            a[i] = ka + __shfl_xor_sync(0xffffffffu, sa_, off);
            b[i] = kb_ + __shfl_xor_sync(0xffffffffu, sb_, off);
          }
        }
        if (c + lane < BN) {
          atomicAdd(&s_stat[c + lane], a[0]);
          atomicAdd(&s_stat[BN + c + lane], b[0]);
        }
      }
    }
    if (kEStats && epi.stats != nullptr) {
      asm volatile("bar.sync 1, 128;" ::: "memory");       // the four epilogue warps
      const int t = threadIdx.x - 64;
      for (int i = t; i < BN; i += 128) {
        const int n = n0 + i;
        if (n < sh.N) {
          atomicAdd(epi.stats + n, (double)s_stat[i]);
          atomicAdd(epi.stats + sh.N + n, (double)s_stat[BN + i]);
        }
      }
    }
    }
    }
  }
tail:
  tc_fence_before();
  __syncthreads();
  if (tr0 && threadIdx.x == 64) sh.trace[1027] = clock64();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}


// =====================================================================================
// Weight-gradient GEMM on tensor cores:  dW[M, N] += sum over pixels p of  X[p, m] * dY[p, n]
// Both operands have the reduction index (pixels) as their SLOW memory dimension, i.e. they are
// "MN-major" for the tensor core.  Canonical MN-major SWIZZLE_128B layout (tf32): one 1024-byte atom =
// 8 pixel rows x 32 channels (128 B); atoms of consecutive 8-row groups are SBO = 1024 B apart, the next
// block of 32 channels is LBO apart.  A stage holds kWgRows = 32 pixel rows of NBA + NBB channel blocks,
// each block written by one TMA box {32 channels, 32 pixels}.  For the 3x3 conv the pixel rows of a stage
// are a TW x TH rectangle fetched as a 4-D box, shifted by the tap of the channel block (zero-filled halo).
// 3xTF32: BOTH operands are activations, so the splitter warps produce hi/lo tiles of both.
// The K (pixel) range is split across grid.z; partial tiles are accumulated with RED.ADD.F32.
// =====================================================================================
constexpr int kWgRows = 32;                       // pixel rows per stage
constexpr int kWgBlk = kWgRows * 128;             // bytes of one 32-channel block of a stage (4096)

struct WgShape {
  int M, N;                 // output rows (channels of X / taps*channels) and columns
  int nstages;              // total pixel stages (all of K)
  int stages_per_split;
  int ldo;                  // output row stride
  // im2col mode
  int C, H, W, TW, TH, tiles_x, tiles_y;
  float* dbg; int dbg_idesc_xor;   // debugging aids
};

template <int NBB, bool X3> struct WgSmem {
  static constexpr int kNBA = 4;
  static constexpr int kStages = (X3 && WAP_TC_X3_DEEP) ? (NBB <= 1 ? 5 : 4) : 3;
  static constexpr int kABytes = kNBA * kWgBlk;                    // 16384
  static constexpr int kBBytes = NBB * kWgBlk;
  static constexpr int kStage = (X3 ? 2 : 1) * (kABytes + kBBytes);
  static constexpr int kBarOff = kStages * kStage;
  static constexpr int kTotal = kBarOff + 256 + 1024;
};

// MN-major 32-bit operands must use the "128-byte swizzle with 32-byte atoms" layout (UMMA layout type 1,
// TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): an atom is 4 K-rows x 128 B (32 channels); 32-byte chunk c of
// row r is stored at chunk c ^ (r & 3).  SBO = next 4-row group (512 B), LBO = next 32-channel block.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;             // leading byte offset: next 32-channel block
  d |= (uint64_t)(512 >> 4) << 32;                   // stride byte offset: next 4-row (K) group
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                            // SWIZZLE_128B_BASE32B
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_tf32_mn(int n) {   // both operands MN-major
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
}

template <int NBB, bool IM2COL, bool X3>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, float* __restrict__ out, const WgShape sh) {
  pdl_sync();
  using S = WgSmem<NBB, X3>;
  constexpr int BN = NBB * 32;
  constexpr uint32_t kAcc2 = BN <= 32 ? 32 : 128;
  constexpr uint32_t kTmemCols = (X3 ? 2 : 1) * (BN <= 32 ? 32 : 128);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
  uint64_t* empty_bar = full_bar + S::kStages;
  uint64_t* split_bar = empty_bar + S::kStages;
  uint64_t* tmem_full = split_bar + S::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kTcBM;
  const int s_beg = blockIdx.z * sh.stages_per_split;
  const int s_end = min(sh.nstages, s_beg + sh.stages_per_split);
  const int ns = s_end - s_beg;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX); prefetch_tmap(&tmY);
    for (int s = 0; s < S::kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&split_bar[s], 4); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (ns <= 0) { __syncthreads(); if (warp == 1) tmem_dealloc(tmem_base, kTmemCols); return; }

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < ns; ++i) {
        const int st = s_beg + i, s = i % S::kStages;
        const uint32_t ph = (i / S::kStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1, 11 + 100 * i);
        uint8_t* sa = smem + s * S::kStage;
        uint8_t* sb = sa + (X3 ? 2 : 1) * S::kABytes;
        mbar_expect_tx(&full_bar[s], S::kABytes + S::kBBytes);
        if (IM2COL) {
          int t = st;
          const int tx = t % sh.tiles_x; t /= sh.tiles_x;
          const int ty = t % sh.tiles_y; const int bimg = t / sh.tiles_y;
          const int x0 = tx * sh.TW, y0 = ty * sh.TH;
          const int cpb = sh.C / 32;                          // channel blocks per tap
#pragma unroll
          for (int blk = 0; blk < 4; ++blk) {
            const int gb = m0 / 32 + blk, tap = gb / cpb, cb = gb - tap * cpb;
            if (tap < 9) tma_load_4d(sa + blk * kWgBlk, &tmX, &full_bar[s], cb * 32, x0 + tap % 3 - 1, y0 + tap / 3 - 1, bimg);
            else tma_load_4d(sa + blk * kWgBlk, &tmX, &full_bar[s], sh.C, x0, y0, bimg);       // fully out of bounds: zeros
          }
#pragma unroll
          for (int blk = 0; blk < NBB; ++blk) tma_load_4d(sb + blk * kWgBlk, &tmY, &full_bar[s], blk * 32, x0, y0, bimg);
        } else {
#pragma unroll
          for (int blk = 0; blk < 4; ++blk) tma_load_2d(sa + blk * kWgBlk, &tmX, &full_bar[s], m0 + blk * 32, st * kWgRows);
#pragma unroll
          for (int blk = 0; blk < NBB; ++blk) tma_load_2d(sb + blk * kWgBlk, &tmY, &full_bar[s], blk * 32, st * kWgRows);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32_mn(BN) ^ (uint32_t)sh.dbg_idesc_xor;
      for (int i = 0; i < ns; ++i) {
        const int s = i % S::kStages;
        const uint32_t ph = (i / S::kStages) & 1;
        mbar_wait(X3 ? &split_bar[s] : &full_bar[s], ph, 12 + 100 * i);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * S::kStage);
        const uint32_t sb = sa + (X3 ? 2 : 1) * S::kABytes;
#pragma unroll
        for (int k = 0; k < kWgRows / 8; ++k) {
          const uint64_t ad = make_mnmajor_sw128_desc(sa + k * 1024, kWgBlk);
          const uint64_t bd = make_mnmajor_sw128_desc(sb + k * 1024, kWgBlk);
          umma_tf32(tmem_base, ad, bd, idesc, (i | k) != 0 ? 1u : 0u);
          if (X3) {
            const uint64_t ald = make_mnmajor_sw128_desc(sa + S::kABytes + k * 1024, kWgBlk);
            const uint64_t bld = make_mnmajor_sw128_desc(sb + S::kBBytes + k * 1024, kWgBlk);
            umma_tf32(tmem_base + kAcc2, ald, bd, idesc, (i | k) != 0 ? 1u : 0u);
            umma_tf32(tmem_base + kAcc2, ad, bld, idesc, 1u);
          }
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(tmem_full);
    }
  } else {
    if (X3) {      // split BOTH operand tiles: hi = rna_tf32(v) in place, lo = v - hi
      const int t = threadIdx.x - 64;
      for (int i = 0; i < ns; ++i) {
        const int s = i % S::kStages;
        const uint32_t ph = (i / S::kStages) & 1;
        mbar_wait(&full_bar[s], ph, 13 + 100 * i);
        uint8_t* base = smem + s * S::kStage;
        float4* ahi = reinterpret_cast<float4*>(base);
        float4* alo = reinterpret_cast<float4*>(base + S::kABytes);
        float4* bhi = reinterpret_cast<float4*>(base + 2 * S::kABytes);
        float4* blo = reinterpret_cast<float4*>(base + 2 * S::kABytes + S::kBBytes);
        auto split4 = [](float4* hi, float4* lo, int idx) {
          const float4 v = hi[idx];
          float4 h, l;
          h.x = tf32_rna_dev(v.x); h.y = tf32_rna_dev(v.y); h.z = tf32_rna_dev(v.z); h.w = tf32_rna_dev(v.w);
          l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
          hi[idx] = h; lo[idx] = l;
        };
#pragma unroll
        for (int j = 0; j < S::kABytes / 16 / 128; ++j) split4(ahi, alo, t + j * 128);
#pragma unroll
        for (int j = 0; j < S::kBBytes / 16 / 128; ++j) split4(bhi, blo, t + j * 128);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&split_bar[s]);
      }
    }
    const int q = warp & 3, r = q * 32 + lane, m = m0 + r;
    mbar_wait(tmem_full, 0, 14);
    tc_fence_after();
    if (sh.dbg != nullptr && blockIdx.x == 0 && blockIdx.z == 0)
      for (int i = threadIdx.x - 64; i < S::kStage / 4; i += 128) sh.dbg[i] = reinterpret_cast<const float*>(smem)[i];
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      if (X3) {
        float v2[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + kAcc2 + (uint32_t)c, v2);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += v2[i];
      }
      if (m < sh.M) {
        float* dst = out + (size_t)m * sh.ldo + c;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c + j < sh.N) atomicAdd(dst + j, v[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

// =====================================================================================
// bf16 weight-gradient GEMM (precision mode 3):  dW[M, N] += sum over pixels p of X[p, m] * dY[p, n], bf16 operands, fp32 result.
// 16-bit MN-major operands use the plain SWIZZLE_128B canonical layout: an atom is 8 pixel rows x 64 channels (128 B per row,
// 16-byte chunk c of row r stored at chunk c ^ (r & 7) - what TMA SWIZZLE_128B writes); SBO = next 8-row group (1024 B),
// LBO = next 64-channel block.  A stage holds 64 pixel rows: A = 2 channel blocks (M = 128), B = NBLK blocks; four K = 16 MMAs.
// IM2COL (3x3 conv): one M tile per tap, its two blocks are channels [0,64) and [64,128) of the tap-shifted input box (channels
// >= C are zero-filled by TMA); the result row (tap, ch) goes to dW row tap*C + ch.
// =====================================================================================
constexpr int kWbRows = 64;                       // pixel rows per stage
constexpr int kWbBlk = kWbRows * 128;             // bytes of one 64-channel block of a stage (8192)
constexpr int kWbStages = 4;

template <int NBLK> struct WbSmem {
  static constexpr int kABytes = 2 * kWbBlk, kBBytes = NBLK * kWbBlk, kStage = kABytes + kBBytes;
  static constexpr int kBarOff = kWbStages * kStage, kTotal = kBarOff + 256 + 1024;
};
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc16(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;             // leading byte offset: next 64-channel block
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset: next 8-row (K) group
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(int n) {   // both operands MN-major
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
}

template <int NMMA, bool IM2COL>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_wgrad_bf16_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, float* __restrict__ out, const WgShape sh) {
  pdl_sync();
  constexpr int NBLK = (NMMA + 63) / 64;
  using S = WbSmem<NBLK>;
  constexpr uint32_t kTmemCols = NMMA <= 32 ? 32 : NMMA <= 64 ? 64 : 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
  uint64_t* empty_bar = full_bar + kWbStages;
  uint64_t* tmem_full = empty_bar + kWbStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kTcBM;                 // IM2COL: blockIdx.x = tap
  const int s_beg = blockIdx.z * sh.stages_per_split;
  const int s_end = min(sh.nstages, s_beg + sh.stages_per_split);
  const int ns = s_end - s_beg;
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX); prefetch_tmap(&tmY);
    for (int s = 0; s < kWbStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (ns <= 0) { __syncthreads(); if (warp == 1) tmem_dealloc(tmem_base, kTmemCols); return; }

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < ns; ++i) {
        const int st = s_beg + i, s = i % kWbStages;
        const uint32_t ph = (i / kWbStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1, 21 + 100 * i);
        uint8_t* sa = smem + s * S::kStage;
        uint8_t* sb = sa + S::kABytes;
        mbar_expect_tx(&full_bar[s], S::kABytes + S::kBBytes);
        if (IM2COL) {
          int t = st;
          const int tx = t % sh.tiles_x; t /= sh.tiles_x;
          const int ty = t % sh.tiles_y; const int bimg = t / sh.tiles_y;
          const int x0 = tx * sh.TW, y0 = ty * sh.TH;
          const int tap = blockIdx.x;
#pragma unroll
          for (int blk = 0; blk < 2; ++blk) tma_load_4d(sa + blk * kWbBlk, &tmX, &full_bar[s], blk * 64, x0 + tap % 3 - 1, y0 + tap / 3 - 1, bimg);
#pragma unroll
          for (int blk = 0; blk < NBLK; ++blk) tma_load_4d(sb + blk * kWbBlk, &tmY, &full_bar[s], blk * 64, x0, y0, bimg);
        } else {
#pragma unroll
          for (int blk = 0; blk < 2; ++blk) tma_load_2d(sa + blk * kWbBlk, &tmX, &full_bar[s], m0 + blk * 64, st * kWbRows);
#pragma unroll
          for (int blk = 0; blk < NBLK; ++blk) tma_load_2d(sb + blk * kWbBlk, &tmY, &full_bar[s], blk * 64, st * kWbRows);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16_mn(NMMA) ^ (uint32_t)sh.dbg_idesc_xor;
      for (int i = 0; i < ns; ++i) {
        const int s = i % kWbStages;
        const uint32_t ph = (i / kWbStages) & 1;
        mbar_wait(&full_bar[s], ph, 22 + 100 * i);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * S::kStage);
        const uint32_t sb = sa + S::kABytes;
#pragma unroll
        for (int k = 0; k < kWbRows / 16; ++k) {
          const uint64_t ad = make_mnmajor_sw128_desc16(sa + k * 2048, kWbBlk);
          const uint64_t bd = make_mnmajor_sw128_desc16(sb + k * 2048, kWbBlk);
          umma_bf16(tmem_base, ad, bd, idesc, (i | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(tmem_full);
    }
  } else {
    const int q = warp & 3, r = q * 32 + lane;
    int orow; bool ok;
    if (IM2COL) { ok = r < sh.C; orow = blockIdx.x * sh.C + r; }
    else { orow = m0 + r; ok = orow < sh.M; }
    mbar_wait(tmem_full, 0, 24);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < NMMA; c += 32) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      if (ok) {
        float* dst = out + (size_t)orow * sh.ldo + c;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c + j < sh.N) atomicAdd(dst + j, v[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

// ------------------------------------------------------------------------------------ host side
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static bool g_encode_tried = false;

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  if (!g_encode_tried) {
    g_encode_tried = true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  return g_encode;
}
bool tc_available() { return get_encode() != nullptr; }

static int make_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                    CUtensorMapSwizzle swz = (kTcRowB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B), bool bf16 = false) {
  PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode();
  WAP_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t gdim[5]; cuuint64_t gstr[4]; cuuint32_t bdim[5]; cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  WAP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): rank=%d dims=%llu,%llu stride0=%llu box=%u,%u", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)strides_bytes[0], box[0], box[1]);
  return 0;
}

long long* g_tc_trace = nullptr;
static int* g_trap_host = nullptr;
const int* trap_note_host() {
  if (g_trap_host == nullptr) {
    int* dptr = nullptr;
    if (cudaHostAlloc(&g_trap_host, 1024, cudaHostAllocMapped) != cudaSuccess) return nullptr;
    memset(g_trap_host, 0, 1024);
    if (cudaHostGetDevicePointer(&dptr, g_trap_host, 0) != cudaSuccess) return nullptr;
    cudaMemcpyToSymbol(g_trap_note, &dptr, sizeof(dptr));
  }
  return g_trap_host;
}
template <int BN, int IM2COL, bool X3, bool BF16, int EPI>
static int launch_tc2(cudaStream_t st, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tblo, const TcEpi& epi,
                      const TcShape& sh, dim3 grid, int prof_cls, double flops) {
  static bool configured = false;
  if (!configured) {
    WAP_CHECK(cudaFuncSetAttribute(tc_gemm_kernel<BN, IM2COL, X3, BF16, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<BN, X3, BF16>::kTotal));
    configured = true;
  }
  TcShape sh2 = sh; sh2.trace = g_tc_trace;
  prof_begin(st, prof_cls, flops);
  WAP_CHECK(launch_pdl(tc_gemm_kernel<BN, IM2COL, X3, BF16, EPI>, grid, TcThreads<X3>::kN, TcSmem<BN, X3, BF16>::kTotal, st, ta, tb, tblo, epi, sh2));
  prof_end(st, prof_cls);
  WAP_LAUNCH_CHECK();
  return 0;
}
// picks the epilogue variant the call needs (see tc_gemm_kernel); only the combinations the library uses are instantiated
template <int BN, int IM2COL, bool X3, bool BF16 = false>
static int launch_tc1(cudaStream_t st, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tblo, const TcEpi& epi,
                      const TcShape& sh, dim3 grid, int prof_cls, double flops) {
  if (epi.bnb_raw != nullptr) {
    if constexpr (BN == 96 && IM2COL == 1) return launch_tc2<BN, IM2COL, X3, BF16, 3>(st, ta, tb, tblo, epi, sh, grid, prof_cls, flops);
    else WAP_REQUIRE(false, "tc_gemm: the fused BN-backward epilogue exists for the 3x3 im2col kernel with one 96-column tile only");
  }
  if (epi.stats != nullptr) {
    WAP_REQUIRE(!epi.accumulate && epi.bias == nullptr && epi.bn_scale == nullptr && !epi.transpose_out && sh.kps == 0,
                "tc_gemm: the statistics epilogue is a plain store (no accumulate / bias / folded BN / transposed / split-K)");
    return launch_tc2<BN, IM2COL, X3, BF16, 1>(st, ta, tb, tblo, epi, sh, grid, prof_cls, flops);
  }
  if (epi.transpose_out || sh.kps > 0) {
    if constexpr (IM2COL == 0 && !BF16) return launch_tc2<BN, IM2COL, X3, BF16, 2>(st, ta, tb, tblo, epi, sh, grid, prof_cls, flops);
    else WAP_REQUIRE(false, "tc_gemm: transposed / split-K output exists for the plain fp32 GEMM only");
  }
  return launch_tc2<BN, IM2COL, X3, BF16, 0>(st, ta, tb, tblo, epi, sh, grid, prof_cls, flops);
}
template <int BN, int IM2COL>
static int launch_tc(cudaStream_t st, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap* tblo, const TcEpi& epi,
                     const TcShape& sh, dim3 grid, int prof_cls, double flops, bool bf16 = false) {
  if (bf16) return launch_tc1<BN, IM2COL, false, true>(st, ta, tb, tb, epi, sh, grid, prof_cls, flops);
  if (tblo != nullptr) return launch_tc1<BN, IM2COL, true>(st, ta, tb, *tblo, epi, sh, grid, prof_cls, flops);
  return launch_tc1<BN, IM2COL, false>(st, ta, tb, tb, epi, sh, grid, prof_cls, flops);
}

__global__ void tc_prep_bt_kernel(const float* __restrict__ W, int K, int N, int transpose, float* __restrict__ hi, float* __restrict__ lo) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * K) return;
  const int n = i / K, k = i - n * K;
  const float v = transpose ? __ldg(W + (size_t)k * N + n) : __ldg(W + i);
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  const float h = __uint_as_float(u);
  hi[i] = h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v - h));
  lo[i] = __uint_as_float(u);
}
__global__ void tc_prep_multi_kernel(const TcPrepBatch b) {
  pdl_sync();
  const TcPrepEntry e = b.e[blockIdx.y];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < e.N * e.K; i += gridDim.x * blockDim.x) {
    const int n = i / e.K, k = i - n * e.K;
    const float v = e.transpose ? __ldg(e.W + (size_t)k * e.N + n) : __ldg(e.W + i);
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    const float h = __uint_as_float(u);
    e.hi[i] = h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v - h));
    e.lo[i] = __uint_as_float(u);
  }
}
int tc_prep_multi(cudaStream_t st, const TcPrepBatch& b) {
  WAP_REQUIRE(b.n >= 1 && b.n <= 8, "tc_prep_multi: n=%d", b.n);
  int mx = 1;
  for (int i = 0; i < b.n; ++i) mx = std::max(mx, b.e[i].N * b.e[i].K);
  WAP_CHECK(launch_pdl(tc_prep_multi_kernel, dim3(std::min(cdiv(mx, 256), 256), b.n), 256, 0, st, b));
  WAP_LAUNCH_CHECK();
  return 0;
}
int tc_prep_bt(cudaStream_t st, const float* W, int K, int N, int transpose, float* hi, float* lo) {
  WAP_CHECK(launch_pdl(tc_prep_bt_kernel, cdiv(N * K, 256), 256, 0, st, W, K, N, transpose, hi, lo));
  WAP_LAUNCH_CHECK();
  return 0;
}

int tc_gemm(cudaStream_t st, const float* A, int lda, const float* Bt, const float* Bt_lo, int ldb, int M, int N, int K, const TcEpi& epi, int prof_cls, int splitk) {
  WAP_REQUIRE(epi.bias == nullptr || epi.stats == nullptr, "tc_gemm: bias and statistics cannot be combined");
  WAP_REQUIRE(!epi.transpose_out || !epi.accumulate, "tc_gemm: transposed store cannot accumulate");
  WAP_REQUIRE(epi.bn_scale == nullptr || (epi.stats == nullptr && !epi.accumulate && !epi.transpose_out && epi.bias == nullptr), "tc_gemm: folded batch norm is store-only");
  WAP_REQUIRE((lda & 3) == 0 && (ldb & 3) == 0 && ((epi.ld & 3) == 0 || epi.transpose_out), "tc_gemm: strides must be multiples of 4 floats");
  WAP_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)Bt & 15) == 0 && ((uintptr_t)epi.out & 15) == 0, "tc_gemm: pointers must be 16-byte aligned");
  CUtensorMap ta, tb, tbl;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M}; const uint64_t str[1] = {(uint64_t)lda * 4}; const uint32_t box[2] = {kTcBK, kTcBM};
    WAP_TRY(make_map(&ta, A, 2, dims, str, box));
  }
  // 3xTF32 is tensor-pipe bound per CTA (one CTA per SM): when the row tiles fill less than half of the machine (block 2 of the encoder: 64
  // tiles), two 48-column CTAs per row tile halve the main loop; the BN statistics are per column, nothing has to be reduced
  const bool split_n = Bt_lo != nullptr && N > 64 && N <= 96 && 2 * cdiv(M, kTcBM) <= 148 && !epi.transpose_out && splitk <= 1;
  const int BN = N <= 32 ? 32 : split_n ? 48 : N <= 96 ? 96 : 128;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N}; const uint64_t str[1] = {(uint64_t)ldb * 4}; const uint32_t box[2] = {kTcBK, (uint32_t)BN};
    WAP_TRY(make_map(&tb, Bt, 2, dims, str, box));
    if (Bt_lo != nullptr) WAP_TRY(make_map(&tbl, Bt_lo, 2, dims, str, box));
  }
  TcShape sh{}; sh.M = M; sh.N = N; sh.K = K; sh.nkb = cdiv(K, kTcBK);
  dim3 grid(cdiv(M, kTcBM), cdiv(N, BN));
  if (splitk > 1) {      // latency-critical GEMMs with few output tiles: the reduction is spread over grid.z, partial sums meet by fp32 atomics
    WAP_REQUIRE(epi.stats == nullptr && !epi.accumulate && !epi.out_bf16 && epi.bn_scale == nullptr, "tc_gemm: split-K is store-only (onto a zero-filled output)");
    sh.kps = cdiv(sh.nkb, std::min(splitk, sh.nkb));
    grid.z = cdiv(sh.nkb, sh.kps);
  }
  const double flops = 2.0 * M * N * K;
  const CUtensorMap* pl = Bt_lo != nullptr ? &tbl : nullptr;
  prof_set_bytes(4.0 * ((double)M * K + (double)N * K * (Bt_lo != nullptr ? 2 : 1) + (double)M * N * (epi.accumulate ? 2 : 1)));
  if (BN == 32) return launch_tc<32, 0>(st, ta, tb, pl, epi, sh, grid, prof_cls, flops);
  if (BN == 48) return launch_tc1<48, 0, true>(st, ta, tb, tbl, epi, sh, grid, prof_cls, flops);
  if (BN == 96) return launch_tc<96, 0>(st, ta, tb, pl, epi, sh, grid, prof_cls, flops);
  return launch_tc<128, 0>(st, ta, tb, pl, epi, sh, grid, prof_cls, flops);
}

int tc_conv3x3(cudaStream_t st, const float* x, int ldx, int B, int H, int W, int C, const float* Wt, const float* Wt_lo, int Nrows, int N, const TcEpi& epi, int prof_cls) {
  WAP_REQUIRE(C % kTcBK == 0 && (ldx & 3) == 0 && (epi.ld & 3) == 0, "tc_conv3x3: C=%d must be a multiple of %d", C, kTcBK);
  WAP_REQUIRE(epi.bnb_raw == nullptr || (epi.stats != nullptr && N <= 96 && !epi.accumulate && !epi.transpose_out && epi.bias == nullptr && (epi.bnb_ld & 3) == 0),
              "tc_conv3x3: the fused BN-backward epilogue needs statistics, one column tile (N <= 96) and a plain store");
  int TW = 1; while (TW < W && TW < kTcBM) TW <<= 1;     // power of two >= W, at most 128
  const int TH = kTcBM / TW;
  CUtensorMap ta, tb, tbl;
  {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)ldx * 4, (uint64_t)W * ldx * 4, (uint64_t)H * W * ldx * 4};
    const uint32_t box[4] = {kTcBK, (uint32_t)TW, (uint32_t)TH, 1};
    WAP_TRY(make_map(&ta, x, 4, dims, str, box));
  }
  const bool split_n = Wt_lo != nullptr && N > 64 && N <= 96 && epi.bnb_raw == nullptr && 2 * cdiv(W, TW) * cdiv(H, TH) * B <= 148;
  const int BN = N <= 32 ? 32 : split_n ? 48 : N <= 96 ? 96 : 128;
  const int K = 9 * C;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)Nrows}; const uint64_t str[1] = {(uint64_t)K * 4}; const uint32_t box[2] = {kTcBK, (uint32_t)BN};
    WAP_TRY(make_map(&tb, Wt, 2, dims, str, box));
    if (Wt_lo != nullptr) WAP_TRY(make_map(&tbl, Wt_lo, 2, dims, str, box));
  }
  const CUtensorMap* pl = Wt_lo != nullptr ? &tbl : nullptr;
  TcShape sh{}; sh.M = B * H * W; sh.N = N; sh.K = K; sh.nkb = K / kTcBK;
  sh.B = B; sh.H = H; sh.W = W; sh.kb_per_tap = C / kTcBK; sh.TW = TW; sh.TH = TH; sh.tiles_x = cdiv(W, TW); sh.tiles_y = cdiv(H, TH);
  dim3 grid(sh.tiles_x * sh.tiles_y * B, cdiv(N, BN));
  const double flops = 2.0 * sh.M * N * K;
  prof_set_bytes(4.0 * ((double)sh.M * C + (double)Nrows * K * (Wt_lo != nullptr ? 2 : 1) + (double)sh.M * N));   // every input element once
  if (BN == 32) return launch_tc<32, 1>(st, ta, tb, pl, epi, sh, grid, prof_cls, flops);
  if (BN == 48) return launch_tc1<48, 1, true>(st, ta, tb, tbl, epi, sh, grid, prof_cls, flops);
  if (BN == 96) return launch_tc<96, 1>(st, ta, tb, pl, epi, sh, grid, prof_cls, flops);
  return launch_tc<128, 1>(st, ta, tb, pl, epi, sh, grid, prof_cls, flops);
}


// 3x3 SAME conv of a [B,H,W,96] activation with 24 output channels as a "row tile" GEMM: K runs over the 3 vertical taps x 96
// channels (9 k-blocks instead of 27), N = 3 horizontal taps x 24 channels (72, padded to 96 rows of Wt3[96][288]:
// Wt3[24 j + co][96 dy + ci] = w[dy][j][ci][co]); the epilogue adds the three 24-column groups with a -1/0/+1 pixel shift.
// One third of the MMAs, A-operand splits and L2->SM traffic of the 9-tap im2col form.  Requires W <= 128 (a tile row must
// hold a whole image row) - the caller falls back to tc_conv3x3 otherwise.
int tc_conv3x3_rowtile(cudaStream_t st, const float* x, int ldx, int B, int H, int W, const float* Wt3, const float* Wt3_lo, const TcEpi& epi, int prof_cls) {
  constexpr int C = 96, N = 24;
  WAP_REQUIRE(W <= kTcBM && (ldx & 3) == 0 && (epi.ld & 3) == 0, "tc_conv3x3_rowtile: W=%d must be <= %d", W, kTcBM);
  int TW = 1; while (TW < W) TW <<= 1;
  const int TH = kTcBM / TW;
  CUtensorMap ta, tb, tbl;
  {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)ldx * 4, (uint64_t)W * ldx * 4, (uint64_t)H * W * ldx * 4};
    const uint32_t box[4] = {kTcBK, (uint32_t)TW, (uint32_t)TH, 1};
    WAP_TRY(make_map(&ta, x, 4, dims, str, box));
  }
  const int K = 3 * C;
  {
    const uint64_t dims[2] = {(uint64_t)K, 96}; const uint64_t str[1] = {(uint64_t)K * 4}; const uint32_t box[2] = {kTcBK, 96};
    WAP_TRY(make_map(&tb, Wt3, 2, dims, str, box));
    if (Wt3_lo != nullptr) WAP_TRY(make_map(&tbl, Wt3_lo, 2, dims, str, box));
  }
  const CUtensorMap* pl = Wt3_lo != nullptr ? &tbl : nullptr;
  TcShape sh{}; sh.M = B * H * W; sh.N = N; sh.K = K; sh.nkb = K / kTcBK;
  sh.B = B; sh.H = H; sh.W = W; sh.kb_per_tap = C / kTcBK; sh.TW = TW; sh.TH = TH; sh.tiles_x = 1; sh.tiles_y = cdiv(H, TH);
  dim3 grid(sh.tiles_y * B, 1);
  prof_set_bytes(4.0 * ((double)sh.M * C + 96.0 * K * (Wt3_lo != nullptr ? 2 : 1) + (double)sh.M * N));
  return launch_tc<96, 2>(st, ta, tb, pl, epi, sh, grid, prof_cls, 2.0 * sh.M * N * 9 * C);
}

// ------------------------------------------------------------------------------------ bf16 variants (precision mode 3)
// Same kernels with BF16 = true: operands are bf16 in global memory (strides in ELEMENTS, byte strides must be multiples of 16),
// a stage row holds 64 elements of K, kind::f16 MMAs, fp32 accumulation; epi.out_bf16 selects a bf16 result.
constexpr int kBfK = kTcRowB / 2;      // 64

int tc_gemm_bf16(cudaStream_t st, const void* A, int lda, const void* Bt, int ldb, int M, int N, int K, const TcEpi& epi, int prof_cls) {
  WAP_REQUIRE(kTcRowB == 128, "bf16 kernels need 128-byte stage rows");
  WAP_REQUIRE((lda & 7) == 0 && (ldb & 7) == 0 && (epi.ld & 3) == 0, "tc_gemm_bf16: operand strides must be multiples of 16 bytes, the output stride of 4 elements (lda=%d ldb=%d ldd=%d)", lda, ldb, epi.ld);
  WAP_REQUIRE(!(epi.out_bf16 && (epi.accumulate || epi.transpose_out || epi.bias)), "tc_gemm_bf16: bf16 output is store-only");
  WAP_REQUIRE(epi.bn_scale == nullptr || (epi.stats == nullptr && !epi.accumulate && !epi.transpose_out && epi.bias == nullptr), "tc_gemm_bf16: folded batch norm is store-only");
  CUtensorMap ta, tb;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M}; const uint64_t str[1] = {(uint64_t)lda * 2}; const uint32_t box[2] = {kBfK, kTcBM};
    WAP_TRY(make_map(&ta, A, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, true));
  }
  const int BN = N <= 32 ? 32 : N <= 96 ? 96 : 128;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N}; const uint64_t str[1] = {(uint64_t)ldb * 2}; const uint32_t box[2] = {kBfK, (uint32_t)BN};
    WAP_TRY(make_map(&tb, Bt, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, true));
  }
  TcShape sh{}; sh.M = M; sh.N = N; sh.K = K; sh.nkb = cdiv(K, kBfK);
  dim3 grid(cdiv(M, kTcBM), cdiv(N, BN));
  const double flops = 2.0 * M * N * K;
  prof_set_bytes(2.0 * ((double)M * K + (double)N * K) + (epi.out_bf16 ? 2.0 : 4.0) * (double)M * N * (epi.accumulate ? 2 : 1));
  if (BN == 32) return launch_tc<32, 0>(st, ta, tb, nullptr, epi, sh, grid, prof_cls, flops, true);
  if (BN == 96) return launch_tc<96, 0>(st, ta, tb, nullptr, epi, sh, grid, prof_cls, flops, true);
  return launch_tc<128, 0>(st, ta, tb, nullptr, epi, sh, grid, prof_cls, flops, true);
}

// 9-tap 3x3 SAME conv of bf16 x[B,H,W,C] (pixel stride ldx elements) with Wt[Nrows][9*Cpad] (k = tap*Cpad + ci, Cpad = C rounded up to 64:
// every tap starts a fresh 64-element k-block, the tail of a partial block is zero in Wt and zero-filled by TMA in x)
int tc_conv3x3_bf16(cudaStream_t st, const void* x, int ldx, int B, int H, int W, int C, const void* Wt, int Nrows, int N, const TcEpi& epi, int prof_cls) {
  WAP_REQUIRE((ldx & 7) == 0 && (C & 7) == 0, "tc_conv3x3_bf16: ldx=%d C=%d must be multiples of 8", ldx, C);
  int TW = 1; while (TW < W && TW < kTcBM) TW <<= 1;
  const int TH = kTcBM / TW;
  const int kbt = cdiv(C, kBfK), Cpad = kbt * kBfK;
  CUtensorMap ta, tb;
  {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)ldx * 2, (uint64_t)W * ldx * 2, (uint64_t)H * W * ldx * 2};
    const uint32_t box[4] = {kBfK, (uint32_t)TW, (uint32_t)TH, 1};
    WAP_TRY(make_map(&ta, x, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, true));
  }
  const int BN = N <= 32 ? 32 : N <= 96 ? 96 : 128;
  const int K = 9 * Cpad;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)Nrows}; const uint64_t str[1] = {(uint64_t)K * 2}; const uint32_t box[2] = {kBfK, (uint32_t)BN};
    WAP_TRY(make_map(&tb, Wt, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, true));
  }
  TcShape sh{}; sh.M = B * H * W; sh.N = N; sh.K = K; sh.nkb = 9 * kbt;
  sh.B = B; sh.H = H; sh.W = W; sh.kb_per_tap = kbt; sh.TW = TW; sh.TH = TH; sh.tiles_x = cdiv(W, TW); sh.tiles_y = cdiv(H, TH);
  dim3 grid(sh.tiles_x * sh.tiles_y * B, cdiv(N, BN));
  prof_set_bytes(2.0 * ((double)sh.M * C + (double)Nrows * K) + (epi.out_bf16 ? 2.0 : 4.0) * (double)sh.M * N);
  const double flops = 2.0 * sh.M * N * 9 * C;
  if (BN == 32) return launch_tc<32, 1>(st, ta, tb, nullptr, epi, sh, grid, prof_cls, flops, true);
  if (BN == 96) return launch_tc<96, 1>(st, ta, tb, nullptr, epi, sh, grid, prof_cls, flops, true);
  return launch_tc<128, 1>(st, ta, tb, nullptr, epi, sh, grid, prof_cls, flops, true);
}

// row-tile 3x3 conv 96 -> 24 (W <= 128) in bf16: Wt3[96][3 * 128] with Wt3[24 j + co][128 dy + ci] = w[dy][j][ci][co] (ci >= 96 zero)
int tc_conv3x3_rowtile_bf16(cudaStream_t st, const void* x, int ldx, int B, int H, int W, const void* Wt3, const TcEpi& epi, int prof_cls) {
  constexpr int C = 96, N = 24, Cpad = 128;
  WAP_REQUIRE(W <= kTcBM && (ldx & 7) == 0, "tc_conv3x3_rowtile_bf16: W=%d must be <= %d", W, kTcBM);
  int TW = 1; while (TW < W) TW <<= 1;
  const int TH = kTcBM / TW;
  CUtensorMap ta, tb;
  {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)ldx * 2, (uint64_t)W * ldx * 2, (uint64_t)H * W * ldx * 2};
    const uint32_t box[4] = {kBfK, (uint32_t)TW, (uint32_t)TH, 1};
    WAP_TRY(make_map(&ta, x, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, true));
  }
  const int K = 3 * Cpad;
  {
    const uint64_t dims[2] = {(uint64_t)K, 96}; const uint64_t str[1] = {(uint64_t)K * 2}; const uint32_t box[2] = {kBfK, 96};
    WAP_TRY(make_map(&tb, Wt3, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, true));
  }
  TcShape sh{}; sh.M = B * H * W; sh.N = N; sh.K = K; sh.nkb = K / kBfK;
  sh.B = B; sh.H = H; sh.W = W; sh.kb_per_tap = Cpad / kBfK; sh.TW = TW; sh.TH = TH; sh.tiles_x = 1; sh.tiles_y = cdiv(H, TH);
  dim3 grid(sh.tiles_y * B, 1);
  prof_set_bytes(2.0 * ((double)sh.M * C + 96.0 * K) + (epi.out_bf16 ? 2.0 : 4.0) * (double)sh.M * N);
  return launch_tc<96, 2>(st, ta, tb, nullptr, epi, sh, grid, prof_cls, 2.0 * sh.M * N * 9 * C, true);
}

extern float* g_wg_dbg; extern int g_wg_xor;
template <int NBB, bool IM2COL>
static int launch_wg(cudaStream_t st, const CUtensorMap& tx, const CUtensorMap& ty, float* out, const WgShape& sh, dim3 grid, bool x3,
                     int prof_cls, double flops) {
  static bool configured = false;
  if (!configured) {
    WAP_CHECK(cudaFuncSetAttribute(tc_wgrad_kernel<NBB, IM2COL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgSmem<NBB, true>::kTotal));
    WAP_CHECK(cudaFuncSetAttribute(tc_wgrad_kernel<NBB, IM2COL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgSmem<NBB, false>::kTotal));
    configured = true;
  }
  WgShape sh2 = sh; sh2.dbg = g_wg_dbg; sh2.dbg_idesc_xor = g_wg_xor;
  prof_begin(st, prof_cls, flops);
  if (x3) WAP_CHECK(launch_pdl(tc_wgrad_kernel<NBB, IM2COL, true>, grid, kTcThreads, WgSmem<NBB, true>::kTotal, st, tx, ty, out, sh2));
  else WAP_CHECK(launch_pdl(tc_wgrad_kernel<NBB, IM2COL, false>, grid, kTcThreads, WgSmem<NBB, false>::kTotal, st, tx, ty, out, sh2));
  prof_end(st, prof_cls);
  WAP_LAUNCH_CHECK();
  return 0;
}

static int pick_splits(int nstages, int mtiles) {
  int splits = std::max(1, std::min(nstages / 6, 148 / std::max(mtiles, 1)));
  return std::max(1, splits);
}

// dW[C, N] (ldo) += X[Mpix, C](ldx)^T * dY[Mpix, N](ldy)      N in {96 (NBB=3)}: 1x1 conv weight gradient
int tc_wgrad_1x1(cudaStream_t st, const float* X, int ldx, const float* dY, int ldy, int Mpix, int C, int N, float* dW, int ldo, bool x3, int prof_cls) {
  WAP_REQUIRE(N <= 96 && (ldx & 3) == 0 && (ldy & 3) == 0, "tc_wgrad_1x1: N=%d must be <= 96", N);
  CUtensorMap tx, ty;
  {
    const uint64_t dims[2] = {(uint64_t)C, (uint64_t)Mpix}; const uint64_t str[1] = {(uint64_t)ldx * 4}; const uint32_t box[2] = {32, kWgRows};
    WAP_TRY(make_map(&tx, X, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  }
  {
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)Mpix}; const uint64_t str[1] = {(uint64_t)ldy * 4}; const uint32_t box[2] = {32, kWgRows};
    WAP_TRY(make_map(&ty, dY, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  }
  WgShape sh{}; sh.M = C; sh.N = N; sh.nstages = cdiv(Mpix, kWgRows); sh.ldo = ldo;
  const int mt = cdiv(C, kTcBM), splits = pick_splits(sh.nstages, mt);
  sh.stages_per_split = cdiv(sh.nstages, splits);
  dim3 grid(mt, 1, cdiv(sh.nstages, sh.stages_per_split));
  prof_set_bytes(4.0 * ((double)Mpix * C + (double)Mpix * N + (double)C * N));
  return launch_wg<3, false>(st, tx, ty, dW, sh, grid, x3, prof_cls, 2.0 * Mpix * C * N);
}

// dW[9*C, N] (ldo) += im2col3x3(X[B,H,W,C] (ldx))^T * dY[B,H,W,32-padded] (ldy = 32)   N <= 32: 3x3 conv weight gradient
int tc_wgrad_3x3(cudaStream_t st, const float* X, int ldx, const float* dY, int ldy, int B, int H, int W, int C, int N, float* dW, int ldo, bool x3,
                 int prof_cls) {
  WAP_REQUIRE(C % 32 == 0 && N <= 32 && ldy == 32, "tc_wgrad_3x3: C=%d N=%d ldy=%d", C, N, ldy);
  int TW = 1; while (TW < W && TW < kWgRows) TW <<= 1;
  const int TH = kWgRows / TW;
  CUtensorMap tx, ty;
  {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)ldx * 4, (uint64_t)W * ldx * 4, (uint64_t)H * W * ldx * 4};
    const uint32_t box[4] = {32, (uint32_t)TW, (uint32_t)TH, 1};
    WAP_TRY(make_map(&tx, X, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  }
  {
    const uint64_t dims[4] = {(uint64_t)ldy, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)ldy * 4, (uint64_t)W * ldy * 4, (uint64_t)H * W * ldy * 4};
    const uint32_t box[4] = {32, (uint32_t)TW, (uint32_t)TH, 1};
    WAP_TRY(make_map(&ty, dY, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  }
  WgShape sh{}; sh.M = 9 * C; sh.N = N; sh.ldo = ldo; sh.C = C; sh.H = H; sh.W = W; sh.TW = TW; sh.TH = TH;
  sh.tiles_x = cdiv(W, TW); sh.tiles_y = cdiv(H, TH); sh.nstages = sh.tiles_x * sh.tiles_y * B;
  const int mt = cdiv(9 * C, kTcBM), splits = pick_splits(sh.nstages, mt);
  sh.stages_per_split = cdiv(sh.nstages, splits);
  dim3 grid(mt, 1, cdiv(sh.nstages, sh.stages_per_split));
  prof_set_bytes(4.0 * ((double)B * H * W * (C + 32) + 9.0 * C * N));
  return launch_wg<1, true>(st, tx, ty, dW, sh, grid, x3, prof_cls, 2.0 * B * H * W * 9 * C * N);
}

template <int NMMA, bool IM2COL>
static int launch_wb(cudaStream_t st, const CUtensorMap& tx, const CUtensorMap& ty, float* out, const WgShape& sh, dim3 grid, int prof_cls, double flops) {
  constexpr int NBLK = (NMMA + 63) / 64;
  static bool configured = false;
  if (!configured) {
    WAP_CHECK(cudaFuncSetAttribute(tc_wgrad_bf16_kernel<NMMA, IM2COL>, cudaFuncAttributeMaxDynamicSharedMemorySize, WbSmem<NBLK>::kTotal));
    configured = true;
  }
  WgShape sh2 = sh; sh2.dbg = nullptr; sh2.dbg_idesc_xor = g_wg_xor;
  prof_begin(st, prof_cls, flops);
  WAP_CHECK(launch_pdl(tc_wgrad_bf16_kernel<NMMA, IM2COL>, grid, kTcThreads, WbSmem<NBLK>::kTotal, st, tx, ty, out, sh2));
  prof_end(st, prof_cls);
  WAP_LAUNCH_CHECK();
  return 0;
}

// dW[C, N] (ldo, fp32) += X[Mpix, C](ldx)^T * dY[Mpix, N](ldy), bf16 operands, N <= 128
int tc_wgrad_1x1_bf16(cudaStream_t st, const void* X, int ldx, const void* dY, int ldy, int Mpix, int C, int N, float* dW, int ldo, int prof_cls) {
  WAP_REQUIRE(N <= 128 && (ldx & 7) == 0 && (ldy & 7) == 0, "tc_wgrad_1x1_bf16: N=%d ldx=%d ldy=%d", N, ldx, ldy);
  CUtensorMap tx, ty;
  {
    const uint64_t dims[2] = {(uint64_t)C, (uint64_t)Mpix}; const uint64_t str[1] = {(uint64_t)ldx * 2}; const uint32_t box[2] = {64, kWbRows};
    WAP_TRY(make_map(&tx, X, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, true));
  }
  {
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)Mpix}; const uint64_t str[1] = {(uint64_t)ldy * 2}; const uint32_t box[2] = {64, kWbRows};
    WAP_TRY(make_map(&ty, dY, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, true));
  }
  WgShape sh{}; sh.M = C; sh.N = N; sh.nstages = cdiv(Mpix, kWbRows); sh.ldo = ldo;
  const int mt = cdiv(C, kTcBM), splits = pick_splits(sh.nstages, mt);
  sh.stages_per_split = cdiv(sh.nstages, splits);
  dim3 grid(mt, 1, cdiv(sh.nstages, sh.stages_per_split));
  prof_set_bytes(2.0 * ((double)Mpix * C + (double)Mpix * N) + 4.0 * (double)C * N);
  if (N <= 32) return launch_wb<32, false>(st, tx, ty, dW, sh, grid, prof_cls, 2.0 * Mpix * C * N);
  if (N <= 96) return launch_wb<96, false>(st, tx, ty, dW, sh, grid, prof_cls, 2.0 * Mpix * C * N);
  return launch_wb<128, false>(st, tx, ty, dW, sh, grid, prof_cls, 2.0 * Mpix * C * N);
}

// dW[9*C, N] (ldo, fp32) += im2col3x3(X[B,H,W,C] (ldx))^T * dY[B,H,W,:] (ldy elements per pixel, N <= 32 valid), bf16 operands, C <= 128
int tc_wgrad_3x3_bf16(cudaStream_t st, const void* X, int ldx, const void* dY, int ldy, int B, int H, int W, int C, int N, float* dW, int ldo, int prof_cls) {
  WAP_REQUIRE(C <= 128 && (C & 7) == 0 && N <= 32 && (ldx & 7) == 0 && (ldy & 7) == 0 && ldy >= N, "tc_wgrad_3x3_bf16: C=%d N=%d ldx=%d ldy=%d", C, N, ldx, ldy);
  int TW = 1; while (TW < W && TW < kWbRows) TW <<= 1;
  const int TH = kWbRows / TW;
  CUtensorMap tx, ty;
  {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)ldx * 2, (uint64_t)W * ldx * 2, (uint64_t)H * W * ldx * 2};
    const uint32_t box[4] = {64, (uint32_t)TW, (uint32_t)TH, 1};
    WAP_TRY(make_map(&tx, X, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, true));
  }
  {
    const uint64_t dims[4] = {(uint64_t)ldy, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)ldy * 2, (uint64_t)W * ldy * 2, (uint64_t)H * W * ldy * 2};
    const uint32_t box[4] = {64, (uint32_t)TW, (uint32_t)TH, 1};
    WAP_TRY(make_map(&ty, dY, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, true));
  }
  WgShape sh{}; sh.M = 9 * C; sh.N = N; sh.ldo = ldo; sh.C = C; sh.H = H; sh.W = W; sh.TW = TW; sh.TH = TH;
  sh.tiles_x = cdiv(W, TW); sh.tiles_y = cdiv(H, TH); sh.nstages = sh.tiles_x * sh.tiles_y * B;
  const int splits = pick_splits(sh.nstages, 9);
  sh.stages_per_split = cdiv(sh.nstages, splits);
  dim3 grid(9, 1, cdiv(sh.nstages, sh.stages_per_split));
  prof_set_bytes(2.0 * ((double)B * H * W * (C + ldy)) + 4.0 * 9.0 * C * N);
  return launch_wb<32, true>(st, tx, ty, dW, sh, grid, prof_cls, 2.0 * B * H * W * 9 * C * N);
}

}  // namespace wap
