// Fused 3xTF32 projection GEMMs on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
//   ghscn_gemm3x     C[M, N]  = A[M, K] . B[N, K]^T (+ bias) (ReLU)     forward projection and dX = dY . W
//   ghscn_gemm3x_tn  dW[M, N] = P[R, M]^T . Q[R, N]                     weight gradient dY^T x
//
// fp32 in, fp32 out, fp32-level accuracy.  They replace, for the h x h projections of GCNConv / GATConv / Linear
// (reference call sites model/mpnn.py:52,59 and model/hscn.py:109 -> PyG `Linear` inside the convs; SURVEY 8a rows
// a2/a9), the round-1 pipeline "split_tf32_cat (writes 3x the activations) + library TF32 GEMM over a 3K-long
// reduction".  Activations are read from HBM once as fp32, split into TF32 hi/lo parts in registers and written
// straight into the swizzled shared-memory operand tiles of tcgen05.mma; only the result goes back.
//
//   x = x_hi + x_lo  (hi: low 13 mantissa bits cleared; lo = x - hi, exact)          same split as ghscn_split_tf32
//   x.w ~= x_hi.w_hi + (x_lo.w_hi + x_hi.w_lo)                                       (lo.lo dropped, ~2^-22)
//
// Accuracy: the tensor core adds every 8-term partial sum into its fp32 accumulator with TRUNCATION, a bias that is
// coherent for same-sign data (post-ReLU features) and grows with the number of accumulations made at full
// magnitude.  Each output tile therefore owns THREE TMEM accumulators: the main term x_hi.w_hi alternates between
// two of them, the small cross terms go to the third, and the epilogue adds the three in fp32 (round to nearest).
// Three accumulators of <= 160 columns fill the 512 TMEM columns, so a CTA walks its 128-row tile once per N half.
//
// Operand layouts (cute::UMMA canonical forms, verified on hardware by scripts/gemm3x_check.py --probe):
//   K-major  (gemm3x A and B):  SWIZZLE_128B, rows of 32 tf32 (128 B), 8-row atoms of 1024 B (SBO), 16-byte chunk
//                               index XOR (row & 7); descriptor start advanced by 32 B per K = 8 step.
//   MN-major (gemm3x_tn P, Q):  SWIZZLE_128B_BASE32B -- the only legal layout for 32-bit MN-major operands: atoms of
//                               4 reduction rows x 32 tf32 (512 B), 32-byte unit index XOR (row & 3), LBO between
//                               32-wide blocks, SBO between 4-row groups; one MMA (K = 8) reads two groups.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"

namespace ghscn {
namespace {

// Pipeline timeline of CTA (0,0) for scripts/gemm3x_trace.py: clock() stamps kept in shared memory (a stamp is one
// CS2R + STS) and dumped at kernel exit.  Compiled in only with -DGHSCN_GEMM3X_TRACE.
#ifdef GHSCN_GEMM3X_TRACE
__device__ long long* g_trace = nullptr;
constexpr int kTraceSlots = 1024;
#define GHSCN_TR_DECL __shared__ unsigned int s_trace[kTraceSlots];
#define GHSCN_TR_INIT                                                                          \
  do {                                                                                         \
    for (int i_ = threadIdx.x; i_ < kTraceSlots; i_ += blockDim.x) s_trace[i_] = 0u;           \
  } while (0)
#define GHSCN_TR(slot)                                                                         \
  do {                                                                                         \
    if ((threadIdx.x & 31) == 0 && (slot) < kTraceSlots) s_trace[(slot)] = (unsigned int)clock(); \
  } while (0)
#define GHSCN_TR_DUMP                                                                          \
  do {                                                                                         \
    __syncthreads();                                                                           \
    if (g_trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0)                              \
      for (int i_ = threadIdx.x; i_ < kTraceSlots; i_ += blockDim.x) g_trace[i_] = s_trace[i_]; \
  } while (0)
#define GHSCN_TR1(slot)                                                                        \
  do {                                                                                         \
    if ((slot) < kTraceSlots) s_trace[(slot)] = (unsigned int)clock();                         \
  } while (0)
#else
#define GHSCN_TR1(slot) do { } while (0)
#define GHSCN_TR_DECL
#define GHSCN_TR_INIT do { } while (0)
#define GHSCN_TR(slot) do { } while (0)
#define GHSCN_TR_DUMP do { } while (0)
#endif

constexpr int kTileM = 128;
constexpr int kHalfMax = 160;                     // columns per accumulator; 3 accumulators = 480 of 512 TMEM columns
constexpr int kMaxN = 2 * kHalfMax;
constexpr uint32_t kColMain0 = 0, kColMain1 = kHalfMax, kColCross = 2 * kHalfMax;
constexpr int kSmemLimit = 227 * 1024;
constexpr unsigned kSpinLimit = 1u << 22;         // a broken pipeline traps instead of hanging the GPU

// ---------------------------------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void bar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  unsigned spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins > kSpinLimit) __trap();
  }
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]^T, kind::tf32.  Called by ALL 32 lanes of the (convergent) MMA warp with
// warp-uniform operands; elect.sync picks the issuing lane.  Written this way the compiler keeps the descriptors in
// uniform registers and emits back-to-back UTCHMMA; issuing from inside `if (lane == 0)` wraps every MMA in an
// ELECT / BRA.U.ANY lane loop (~90 cycles per MMA, more than an N = 160 MMA takes to execute).
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred pe, pa;\n\telect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, pa;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All previously issued MMAs arrive on `bar` once they have completed (implies fence::before).  Convergent warp.
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns (no wait): thread l of the warp receives TMEM lane (quadrant*32 + l).
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// Waits for the outstanding tcgen05.ld of this thread; the registers are tied to the statement so that the compiler
// cannot schedule their consumers above the wait.
__device__ __forceinline__ void tmem_ld_wait(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// main0 + main1 + cross of 16 columns -> v (fp32, round to nearest)
__device__ __forceinline__ void load_sum16(uint32_t tbase, uint32_t col, bool use_main1, float* v) {
  uint32_t m0[16], m1[16], cr[16];
  tmem_ld16_nowait(tbase + kColMain0 + col, m0);
  tmem_ld16_nowait(tbase + kColCross + col, cr);
  if (use_main1) tmem_ld16_nowait(tbase + kColMain1 + col, m1);
  tmem_ld_wait(m0);
  tmem_ld_wait(cr);
  if (use_main1) tmem_ld_wait(m1);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float s = __uint_as_float(m0[i]);
    if (use_main1) s += __uint_as_float(m1[i]);
    v[i] = s + __uint_as_float(cr[i]);
  }
}

__device__ __forceinline__ void split_store(unsigned char* hi_ptr, unsigned char* lo_ptr, const float4 v) {
  float4 h, l;
  h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
  h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
  h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
  h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
  *reinterpret_cast<float4*>(hi_ptr) = h;
  *reinterpret_cast<float4*>(lo_ptr) = l;
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): [0,14) start >> 4 | [16,30) LBO >> 4 |
// [32,46) SBO >> 4 | [46,48) version = 1 | [49,52) base offset = 0 | [61,64) layout type.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
constexpr uint32_t kLayoutSw128 = 2, kLayoutSw128Base32 = 1;
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t saddr) { return make_desc(saddr, 16, 1024, kLayoutSw128); }

// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 [4,6)=1, A = B = TF32 [7,10)=[10,13)=2,
// A/B major at bits 15/16 (0 = K-major, 1 = MN-major), N >> 3 at [17,23), M >> 4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n, bool mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (mn_major ? ((1u << 15) | (1u << 16)) : 0u) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// byte offset of element (row, k) inside one K-major SW128 block of 32 columns
__host__ __device__ __forceinline__ uint32_t sw128_offset(int row, int k) {
  const int chunk = k >> 2, e = k & 3, r8 = row & 7;
  return (uint32_t)((row >> 3) * 1024 + r8 * 128 + ((chunk ^ r8) << 4) + e * 4);
}

// The N range of a problem is cut into one or two "halves" of <= 160 padded columns.
struct Halves {
  int count;
  int pad[2];     // accumulator columns (multiple of 16)
  int valid[2];   // real columns
  int col[2];     // first column
};
__host__ __device__ inline Halves make_halves(int n_out) {
  Halves h;
  const int npad = (n_out + 15) / 16 * 16;
  if (npad <= kHalfMax) {
    h.count = 1; h.pad[0] = npad; h.valid[0] = n_out; h.col[0] = 0;
    h.pad[1] = 0; h.valid[1] = 0; h.col[1] = 0;
  } else {
    h.count = 2;
    h.pad[0] = ((npad / 2) + 15) / 16 * 16;
    if (h.pad[0] % 32) h.pad[0] += 16;            // MN-major operands start the second half on a 32-column block
    if (h.pad[0] > kHalfMax) h.pad[0] = kHalfMax;
    h.pad[1] = npad - h.pad[0];
    h.col[0] = 0; h.col[1] = h.pad[0];
    h.valid[0] = h.pad[0];
    h.valid[1] = n_out - h.pad[0];
  }
  return h;
}

// =====================================================================================================================
// gemm3x:  C = A . B^T.  One CTA per 128-row tile of A; for each N half: K loop over 32-wide chunks.
//   warp 0 / lane 0   weight-image producer: one cp.async.bulk per chunk into a 3-stage B ring (L2 resident image,
//                     pre-split and pre-swizzled by gemm3x_prep_b_kernel)
//   warp 1            MMA issuer, convergent (owns TMEM): per K = 8 step  cross += a_lo.b_hi, cross += a_hi.b_lo,
//                     main[chunk & 1] += a_hi.b_hi; commits the stage barriers, then the accumulator barrier
//   warps 2..7        A producers, each owning whole K chunks: 128-bit global loads of the chunk (all in flight) ->
//                     hi/lo split -> swizzled smem (3-stage A ring) -> fence.proxy.async -> mbarrier arrive
//   warps 8..11       epilogue (own TMEM lane quadrant = warp & 3): tcgen05.ld of 32 columns of the three accumulators
//                     per round -> fp32 sum -> bias / ReLU -> 128-bit global stores (one full line per lane and round);
//                     the stores of half h overlap the main loop of half h + 1
// =====================================================================================================================
constexpr int kChunkK = 32;
constexpr int kABytes = kTileM * kChunkK * 4;               // one part (hi or lo) of an A stage: 16 KB
constexpr int kAStages = 3, kBStages = 3;
constexpr int kBStageBytes = 2 * kHalfMax * kChunkK * 4;    // hi block + lo block: 40 KB
constexpr int kNnSmem = kAStages * 2 * kABytes + kBStages * kBStageBytes + 1024;
constexpr int kNnProducerWarps = 6;
constexpr int kNnThreads = 32 * (2 + kNnProducerWarps + 4);
static_assert(kNnSmem <= kSmemLimit - 1024, "gemm3x shared memory budget");

struct NnPlan {
  Halves hv;
  int kchunks;
  int64_t img_off[2];       // byte offset of each half's weight image
};
__host__ __device__ inline NnPlan make_nn_plan(int n_out, int k) {
  NnPlan p;
  p.hv = make_halves(n_out);
  p.kchunks = (k + kChunkK - 1) / kChunkK;
  p.img_off[0] = 0;
  p.img_off[1] = (int64_t)p.kchunks * 2 * p.hv.pad[0] * kChunkK * 4;
  return p;
}
__host__ __device__ inline int64_t nn_image_bytes(const NnPlan& p) {
  return p.img_off[1] + (int64_t)p.kchunks * 2 * p.hv.pad[1] * kChunkK * 4;
}

// Weight image: for each N half, for each K chunk: [pad rows x 128 B] hi parts then the same of lo parts, in the
// swizzled smem layout, zero padded in N and K.  b[n, k] = transpose ? w[k*ldw + n] : w[n*ldw + k].
__global__ void __launch_bounds__(256) gemm3x_prep_b_kernel(const float* __restrict__ w, int64_t ldw, int n_out,
                                                            int k_dim, int transpose, NnPlan plan,
                                                            unsigned char* __restrict__ image) {
  const int npad = plan.hv.pad[0] + plan.hv.pad[1];
  const int64_t total = (int64_t)plan.kchunks * npad * kChunkK;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int kk, n, kc;
    if (transpose) {          // consecutive threads walk n (contiguous in w when transposed)
      n = (int)(i % npad);
      const int64_t t = i / npad;
      kk = (int)(t % kChunkK);
      kc = (int)(t / kChunkK);
    } else {
      kk = (int)(i % kChunkK);
      const int64_t t = i / kChunkK;
      n = (int)(t % npad);
      kc = (int)(t / npad);
    }
    const int k = kc * kChunkK + kk;
    float v = 0.f;
    if (n < n_out && k < k_dim) v = transpose ? __ldg(w + (int64_t)k * ldw + n) : __ldg(w + (int64_t)n * ldw + k);
    const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    const float lo = v - hi;
    const int h = (plan.hv.count == 2 && n >= plan.hv.pad[0]) ? 1 : 0;
    const int nl = n - plan.hv.col[h];
    const int64_t block_bytes = (int64_t)plan.hv.pad[h] * kChunkK * 4;
    unsigned char* base = image + plan.img_off[h] + (int64_t)kc * 2 * block_bytes + sw128_offset(nl, kk);
    *reinterpret_cast<float*>(base) = hi;
    *reinterpret_cast<float*>(base + block_bytes) = lo;
  }
}

__global__ void __launch_bounds__(kNnThreads, 1)
gemm3x_kernel(const float* __restrict__ a, int64_t lda, int m_rows, int k_dim, const unsigned char* __restrict__ b_image,
              int n_out, const float* __restrict__ bias, int relu, float* __restrict__ c, NnPlan plan, int split) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long a_full[kAStages], a_empty[kAStages];
  __shared__ __align__(8) unsigned long long b_full[kBStages], b_empty[kBStages];
  __shared__ __align__(8) unsigned long long acc_full, acc_empty;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int a_turn;                   // next chunk whose producer may test its stage's empty barrier
  GHSCN_TR_DECL
  GHSCN_TR_INIT;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = (smem_addr(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_addr(smem_raw));
  const uint32_t a_ring = smem_base, b_ring = a_ring + kAStages * 2 * kABytes;
  const int m0 = blockIdx.x * kTileM;
  // split != 0: gridDim.y = number of N halves and this CTA computes only half blockIdx.y of its row tile (half the
  // work per CTA, twice the CTAs): used when the row tiles do not fill one wave of the 148 SMs exactly -- a problem of
  // 156 tiles runs as 312 half CTAs in 3 waves of ~11 us instead of one wave of ~23 us plus a tail GEMM, and a small
  // one (<= 74 tiles) finishes in half the time.  `hoff` maps the CTA-local half index to the problem's.
  const int hoff = split ? (int)blockIdx.y : 0;
  const int kchunks = plan.kchunks, nh = split ? 1 : plan.hv.count;
  const int total_chunks = nh * kchunks;
  // Every CTA streams the SAME weight image; walking it in lock step makes ~148 SMs hit the same few L2 slices at
  // once.  CTA b therefore starts at K chunk b mod kchunks (a sum: any chunk order is valid, and it is fixed per tile).
  const int rot = (int)(blockIdx.x % (unsigned)kchunks);

  if (tid == 0) {
    for (int s = 0; s < kAStages; ++s) { bar_init(smem_addr(&a_full[s]), 1); bar_init(smem_addr(&a_empty[s]), 1); }
    for (int s = 0; s < kBStages; ++s) { bar_init(smem_addr(&b_full[s]), 1); bar_init(smem_addr(&b_empty[s]), 1); }
    bar_init(smem_addr(&acc_full), 1);
    bar_init(smem_addr(&acc_empty), 4);
    a_turn = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_addr(&tmem_slot), 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===== weight-image producer =====
    if (lane == 0) {
      for (int g = 0; g < total_chunks; ++g) {
        const int hl = g >= kchunks ? 1 : 0, h = hl + hoff, kc = (g - hl * kchunks + rot) % kchunks;
        const int s = g % kBStages;
        const uint32_t ph = (uint32_t)(g / kBStages) & 1u;
        if (g < 50) GHSCN_TR1(300 + 2 * g);
        bar_wait(smem_addr(&b_empty[s]), ph ^ 1u);
        if (g < 50) GHSCN_TR1(301 + 2 * g);
        const uint32_t bytes = (uint32_t)(2 * plan.hv.pad[h] * kChunkK * 4);
        const uint32_t fb = smem_addr(&b_full[s]);
        bar_arrive_expect_tx(fb, bytes);
        bulk_load(b_ring + (uint32_t)s * kBStageBytes, b_image + plan.img_off[h] + (int64_t)kc * bytes, bytes, fb);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer (all 32 lanes run the loop; elect.sync inside mma_tf32 / mma_commit) =====
    {
      for (int hl = 0; hl < nh; ++hl) {
        const int h = hl + hoff;
        const uint32_t idesc = make_idesc_tf32(kTileM, plan.hv.pad[h], false);
        const uint32_t lo_off = (uint32_t)plan.hv.pad[h] * kChunkK * 4;
        if (hl > 0) {                                  // the epilogue has drained the accumulators of half hl-1
          bar_wait(smem_addr(&acc_empty), (uint32_t)(hl - 1) & 1u);
          tc_fence_after();
        }
        for (int pos = 0; pos < kchunks; ++pos) {
          const int g = hl * kchunks + pos;
          const int kc = (pos + rot) % kchunks;
          const int sa = g % kAStages, sb = g % kBStages;
          GHSCN_TR(3 * g);
          bar_wait(smem_addr(&a_full[sa]), (uint32_t)(g / kAStages) & 1u);
          GHSCN_TR(3 * g + 1);
          bar_wait(smem_addr(&b_full[sb]), (uint32_t)(g / kBStages) & 1u);
          GHSCN_TR(3 * g + 2);
          tc_fence_after();
          const uint32_t as = a_ring + (uint32_t)sa * 2 * kABytes, bs = b_ring + (uint32_t)sb * kBStageBytes;
          const uint64_t a_hi = desc_k_sw128(as), a_lo = desc_k_sw128(as + kABytes);
          const uint64_t b_hi = desc_k_sw128(bs), b_lo = desc_k_sw128(bs + lo_off);
          const int kleft = k_dim - kc * kChunkK;
          const int ksteps = kleft >= kChunkK ? kChunkK / 8 : (kleft + 7) / 8;
          const uint32_t d_main = tmem_base + ((pos & 1) ? kColMain1 : kColMain0);
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t adv = (uint64_t)(ks * 2);   // 8 tf32 = 32 bytes = 2 x 16-byte units
            mma_tf32(tmem_base + kColCross, a_lo + adv, b_hi + adv, idesc, (pos > 0 || ks > 0) ? 1u : 0u);
            mma_tf32(tmem_base + kColCross, a_hi + adv, b_lo + adv, idesc, 1u);
            mma_tf32(d_main, a_hi + adv, b_hi + adv, idesc, (pos > 1 || ks > 0) ? 1u : 0u);
          }
          mma_commit(smem_addr(&a_empty[sa]));
          mma_commit(smem_addr(&b_empty[sb]));
        }
        mma_commit(smem_addr(&acc_full));
      }
    }
    __syncwarp();
  } else if (warp < 2 + kNnProducerWarps) {
    // ===== A producers (warps 2..7): every warp owns whole K chunks (g = pw, pw + 6, ...) =====
    // All 32 float4 of the chunk a lane is responsible for are loaded at once; after the split + stores no global
    // load of this thread is outstanding, so fence.proxy.async is cheap, and the other five warps keep their
    // chunks' loads in flight meanwhile (see the note in gemm3x_tn_kernel).
    const int pw = warp - 2;
    const int cq = lane & 7;                // 16-byte chunk of the 128-byte row
    const int rsub = lane >> 3;             // rows rsub + 4 j
    for (int g = pw; g < total_chunks; g += kNnProducerWarps) {
      const int kc = ((g >= kchunks ? g - kchunks : g) + rot) % kchunks;
      const int k = kc * kChunkK + cq * 4;
      const bool kvalid = k < k_dim;
      const float* ap = a + (int64_t)(m0 + rsub) * lda + k;
      float4 v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j)
        v[j] = (kvalid && m0 + rsub + 4 * j < m_rows) ? ldg_f4(ap + (int64_t)(4 * j) * lda)
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
      const int s = g % kAStages;
      // Producers test the empty barriers in chunk order (a_turn): a parity wait is only meaningful when the waiter
      // is at most one phase behind, and the owner of chunk g - kAStages has passed its wait by now.
      GHSCN_TR(100 + 4 * g);
      while (a_turn != g) { }
      bar_wait(smem_addr(&a_empty[s]), ((uint32_t)(g / kAStages) & 1u) ^ 1u);
      GHSCN_TR(100 + 4 * g + 1);
      __syncwarp();
      if (lane == 0) a_turn = g + 1;
      unsigned char* st = smem_gen + (size_t)s * 2 * kABytes;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const uint32_t off = sw128_offset(rsub + 4 * j, cq * 4);
        split_store(st + off, st + kABytes + off, v[j]);
      }
      GHSCN_TR(100 + 4 * g + 2);
      fence_proxy_async();                             // generic-proxy stores -> async proxy (tcgen05.mma)
      __syncwarp();
      if (lane == 0) bar_arrive(smem_addr(&a_full[s]));
      GHSCN_TR(100 + 4 * g + 3);
    }
  } else {
    // ===== epilogue (warps 8..11) =====
    // Each lane owns one tile row (= TMEM lane).  Per round: the tcgen05.ld of 32 columns of all three accumulators are
    // issued together and waited for once (one ld + wait round trip costs ~300 cycles whatever its width), summed in
    // fp32, and written straight to global memory: 32 columns are one full 128-byte line per lane.  The accumulators
    // are handed back to the MMA warp right after the last wait, before the last stores.
    const int quad = warp & 3;                                  // TMEM lane quadrant this warp may read
    const int grow = m0 + quad * 32 + lane;                     // tile row == TMEM lane
    const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16);
    const bool use_main1 = kchunks > 1;
    float* crow = c + (int64_t)grow * n_out;
    for (int hl = 0; hl < nh; ++hl) {
      const int h = hl + hoff;
      if (quad == 0) GHSCN_TR(400 + 4 * hl);
      bar_wait(smem_addr(&acc_full), (uint32_t)hl & 1u);
      if (quad == 0) GHSCN_TR(401 + 4 * hl);
      tc_fence_after();
      const int hpad = plan.hv.pad[h], hvalid = plan.hv.valid[h], hcol = plan.hv.col[h];
      for (int c0 = 0; c0 < hpad; c0 += 32) {
        const bool second = c0 + 16 < hpad;
        uint32_t a0[16], a1[16], ac[16], b0[16], b1[16], bc[16];
        tmem_ld16_nowait(tbase + kColMain0 + (uint32_t)c0, a0);
        tmem_ld16_nowait(tbase + kColCross + (uint32_t)c0, ac);
        if (use_main1) tmem_ld16_nowait(tbase + kColMain1 + (uint32_t)c0, a1);
        if (second) {
          tmem_ld16_nowait(tbase + kColMain0 + (uint32_t)(c0 + 16), b0);
          tmem_ld16_nowait(tbase + kColCross + (uint32_t)(c0 + 16), bc);
          if (use_main1) tmem_ld16_nowait(tbase + kColMain1 + (uint32_t)(c0 + 16), b1);
        }
        tmem_ld_wait(a0);
        tmem_ld_wait(ac);
        if (use_main1) tmem_ld_wait(a1);
        if (second) {
          tmem_ld_wait(b0);
          tmem_ld_wait(bc);
          if (use_main1) tmem_ld_wait(b1);
        }
        if (c0 + 32 >= hpad) {                                  // accumulators fully read: release them to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) bar_arrive(smem_addr(&acc_empty));
          if (quad == 0) GHSCN_TR(402 + 4 * hl);
        }
#pragma unroll
        for (int part = 0; part < 2; ++part) {
          if (part == 1 && !second) break;
          const uint32_t* pm0 = part ? b0 : a0;
          const uint32_t* pm1 = part ? b1 : a1;
          const uint32_t* pcr = part ? bc : ac;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int lc = c0 + 16 * part + 4 * q;              // column inside the half
            if (lc < hvalid) {
              float v[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float t = __uint_as_float(pm0[4 * q + i]);
                if (use_main1) t += __uint_as_float(pm1[4 * q + i]);
                v[i] = t + __uint_as_float(pcr[4 * q + i]);
              }
              float4 o = make_float4(v[0], v[1], v[2], v[3]);
              if (bias != nullptr) {
                const float4 b = ldg_f4(bias + hcol + lc);
                o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
              }
              if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
              if (grow < m_rows) *reinterpret_cast<float4*>(crow + hcol + lc) = o;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) GHSCN_TR(410);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
  GHSCN_TR_DUMP;
}

// =====================================================================================================================
// gemm3x_tn:  dW[M, N] = P[R, M]^T . Q[R, N]  (P = dY, Q = x; the reduction runs over the R node rows).
// Grid = (M tiles of 128 x N halves, row slabs).  Every CTA reduces its slab into the three accumulators and writes an
// fp32 partial; gemm3x_tn_reduce_kernel adds the slab partials in slab order (deterministic).
//   warp 0            MMA issuer, convergent (owns TMEM)
//   warps 1..16       producers: P slice [16 rows x 128] and Q slice [16 rows x <=160] per stage, 128-bit loads along
//                     M/N land on 16-byte halves of the swizzle units (no transposition).  A chunk is shared by TWO
//                     warps (8 rows of P and half of Q each): the pipeline trace showed the MMA warp waiting for
//                     its producers two thirds of the time -- a warp needs ~10 k cycles per chunk, almost all of it
//                     the latency of its 18 KB of loads -- so eight chunks in flight of 9 KB per warp were replaced
//                     by eight chunks in flight of 2 x 4.5 KB: twice the loads in the air per SM.  Afterwards the
//                     same warps drain TMEM.
// =====================================================================================================================
constexpr int kTnRows = 16;                        // reduction rows per pipeline stage (two K = 8 MMAs)
constexpr int kTnSplit = 2;                        // producer warps that share one chunk (each stages half of its items)
constexpr int kTnChunkWarps = 8;                   // chunks in flight = producer warps / kTnSplit
constexpr int kTnProducers = 32 * kTnChunkWarps * kTnSplit;   // warps 1..16
constexpr int kTnThreads = 32 + kTnProducers;
constexpr int kTnABlocks = 4, kTnBBlocks = kHalfMax / 32;                 // 32-wide M/N blocks per operand
constexpr int kTnAPart = (kTnRows / 4) * kTnABlocks * 512;                 // 8 KB  (hi or lo)
constexpr int kTnBPart = (kTnRows / 4) * kTnBBlocks * 512;                 // 10 KB
constexpr int kTnStageBytes = 2 * kTnAPart + 2 * kTnBPart;                 // 36 KB
constexpr int kTnStages = 6;
constexpr int kTnSmem = kTnStages * kTnStageBytes + 1024;
constexpr int kSlabRows = 1024;                    // upper bound of rows reduced inside one set of accumulators
static_assert(kTnSmem <= kSmemLimit - 1024, "gemm3x_tn shared memory budget");
static_assert(kTileM * (kHalfMax + 4) * 4 <= kTnStages * kTnStageBytes, "epilogue staging reuses the stage ring");

struct TnPlan {
  Halves hv;
  int mtiles, nslabs, chunks_per_slab;
};
__host__ __device__ inline TnPlan make_tn_plan(int64_t rows, int m_out, int n_out) {
  TnPlan p;
  p.hv = make_halves(n_out);
  p.mtiles = (m_out + kTileM - 1) / kTileM;
  const int64_t chunks = (rows + kTnRows - 1) / kTnRows;
  int64_t nslabs = kNumSMs / (p.mtiles * p.hv.count);                      // one wave when the slabs are short enough
  const int64_t min_slabs = (rows + kSlabRows - 1) / kSlabRows;
  if (nslabs < min_slabs) nslabs = min_slabs;
  if (nslabs > chunks) nslabs = chunks;
  if (nslabs < 1) nslabs = 1;
  p.chunks_per_slab = (int)((chunks + nslabs - 1) / nslabs);
  p.nslabs = (int)((chunks + p.chunks_per_slab - 1) / p.chunks_per_slab);
  return p;
}

// byte offset of the 16-byte unit `f4` (4 consecutive M/N values) of reduction row `rr` inside an operand part laid
// out as [4-row group][32-wide block][512 B atom]
__device__ __forceinline__ uint32_t mn_offset(int rr, int f4, int nblocks) {
  const int kg = rr >> 2, k4 = rr & 3, c16 = f4 & 7;
  return (uint32_t)((kg * nblocks + (f4 >> 3)) * 512 + k4 * 128 + (((c16 >> 1) ^ k4) << 5) + ((c16 & 1) << 4));
}

template <bool kSegmented>
__global__ void __launch_bounds__(kTnThreads, 1)
gemm3x_tn_kernel(const float* __restrict__ pmat, int64_t ldp, const float* __restrict__ qmat, int64_t ldq, int rows,
                 int m_out, int n_out, float* __restrict__ partial, TnPlan plan, const int* __restrict__ seg_ptr,
                 int64_t out_slab_stride, int64_t ldo) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full_bar[kTnStages], empty_bar[kTnStages];
  __shared__ __align__(8) unsigned long long accum_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int turn;                     // next chunk whose producer may test its stage's empty barrier
  GHSCN_TR_DECL
  GHSCN_TR_INIT;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = (smem_addr(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_addr(smem_raw));
  const int nh = plan.hv.count;
  const int mt = blockIdx.x / nh, h = blockIdx.x - mt * nh;
  const int m0 = mt * kTileM;
  const int hpad = plan.hv.pad[h], hvalid = plan.hv.valid[h], hcol = plan.hv.col[h];
  // blockIdx.y is a row slab of ONE product (dW: equal slabs, partials reduced afterwards) or, with `seg_ptr`, one
  // segment of a batch of independent products (MinCUT S^T X per graph: rows seg_ptr[g] .. seg_ptr[g+1]-1).
  const int slab = blockIdx.y;
  int row_begin, row_end, nchunks;
  if (kSegmented) {
    row_begin = seg_ptr[slab];
    row_end = seg_ptr[slab + 1];
    nchunks = max(0, (row_end - row_begin + kTnRows - 1) / kTnRows);
  } else {
    const int chunk0 = slab * plan.chunks_per_slab;
    row_begin = chunk0 * kTnRows;
    row_end = rows;
    nchunks = min((rows + kTnRows - 1) / kTnRows - chunk0, plan.chunks_per_slab);
  }

  if (tid == 0) {
    for (int s = 0; s < kTnStages; ++s) {
      bar_init(smem_addr(&full_bar[s]), kTnSplit);     // the producer warps that share the stage's chunk
      bar_init(smem_addr(&empty_bar[s]), 1);
    }
    bar_init(smem_addr(&accum_bar), 1);
    turn = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_addr(&tmem_slot), 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===== MMA issuer (all 32 lanes run the loop; elect.sync inside mma_tf32 / mma_commit) =====
    {
      const uint32_t idesc = make_idesc_tf32(kTileM, hpad, true);
      constexpr uint32_t a_sbo = kTnABlocks * 512, b_sbo = kTnBBlocks * 512;   // one 4-row group of all blocks
      for (int c = 0; c < nchunks; ++c) {
        const int s = c % kTnStages;
        GHSCN_TR(2 * c);
        bar_wait(smem_addr(&full_bar[s]), (uint32_t)(c / kTnStages) & 1u);
        GHSCN_TR(2 * c + 1);
        tc_fence_after();
        const uint32_t st = smem_base + (uint32_t)s * kTnStageBytes;
        const uint32_t a_hi = st, a_lo = st + kTnAPart, b_hi = st + 2 * kTnAPart, b_lo = b_hi + kTnBPart;
#pragma unroll
        for (int g = 0; g < kTnRows / 8; ++g) {        // 8 reduction rows = two 4-row groups per MMA
          const uint32_t ao = g * 2 * a_sbo, bo = g * 2 * b_sbo;
          const uint64_t dah = make_desc(a_hi + ao, 512, a_sbo, kLayoutSw128Base32);
          const uint64_t dal = make_desc(a_lo + ao, 512, a_sbo, kLayoutSw128Base32);
          const uint64_t dbh = make_desc(b_hi + bo, 512, b_sbo, kLayoutSw128Base32);
          const uint64_t dbl = make_desc(b_lo + bo, 512, b_sbo, kLayoutSw128Base32);
          mma_tf32(tmem_base + kColCross, dal, dbh, idesc, (c > 0 || g > 0) ? 1u : 0u);
          mma_tf32(tmem_base + kColCross, dah, dbl, idesc, 1u);
          mma_tf32(tmem_base + (g ? kColMain1 : kColMain0), dah, dbh, idesc, c > 0 ? 1u : 0u);
        }
        mma_commit(smem_addr(&empty_bar[s]));
      }
      mma_commit(smem_addr(&accum_bar));
    }
    __syncwarp();
  } else {
    // ===== producers: every warp owns whole stages (chunk pw, pw + 8, ...) =====
    // A warp loads ALL operand data of its chunk into registers (36 float4 per lane, every load in flight at once),
    // then splits and stores it and only then executes fence.proxy.async: the fence waits for the thread's
    // outstanding global loads, and here there are none left, while the other seven warps have their chunks'
    // loads in flight.  (A register-prefetch ring inside the fencing thread is drained by every fence: that
    // version ran at one memory latency per chunk.)
    const int pw = (warp - 1) / kTnSplit;            // chunk slot 0..7
    const int part = (warp - 1) % kTnSplit;          // which half of the chunk's items
    constexpr int kQ4 = kHalfMax / 4;                // float4 slots per Q row (padded to 160 columns)
    constexpr int kAItems = kTnRows / kTnSplit;      // P: one row per item, lane = float4 column     (8 per warp)
    constexpr int kBItems = kTnRows * kQ4 / 32 / kTnSplit;   // Q: item = lane + 32 i -> (row, float4 column) (10)
    const int a0 = part * kAItems, b0 = part * kBItems;
    const bool a_col_ok = m0 + lane * 4 < m_out;
    for (int c = pw; c < nchunks; c += kTnChunkWarps) {
      const int r_base = row_begin + c * kTnRows;
      const float* pb = pmat + (int64_t)r_base * ldp + m0 + lane * 4;
      const float* qb = qmat + (int64_t)r_base * ldq + hcol;
      float4 va[kAItems], vb[kBItems];
#pragma unroll
      for (int i = 0; i < kAItems; ++i)
        va[i] = (a_col_ok && r_base + a0 + i < row_end) ? ldg_f4(pb + (int64_t)(a0 + i) * ldp)
                                                        : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < kBItems; ++i) {
        const int item = lane + 32 * (b0 + i), rr = item / kQ4, f4 = item - rr * kQ4;
        vb[i] = (f4 * 4 < hvalid && r_base + rr < row_end) ? ldg_f4(qb + (int64_t)rr * ldq + f4 * 4)
                                                        : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const int s = c % kTnStages;
      if (warp == 1) GHSCN_TR(200 + 5 * (c / 8));
      while (turn != c * kTnSplit + part) { }        // empty barriers are tested in chunk order (parity waits must
      bar_wait(smem_addr(&empty_bar[s]), ((uint32_t)(c / kTnStages) & 1u) ^ 1u);   // be at most one phase behind)
      __syncwarp();
      if (lane == 0) turn = c * kTnSplit + part + 1;
      if (warp == 1) GHSCN_TR(200 + 5 * (c / 8) + 1);
      unsigned char* st = smem_gen + (size_t)s * kTnStageBytes;
#pragma unroll
      for (int i = 0; i < kAItems; ++i) {
        const uint32_t off = mn_offset(a0 + i, lane, kTnABlocks);
        split_store(st + off, st + kTnAPart + off, va[i]);
      }
#pragma unroll
      for (int i = 0; i < kBItems; ++i) {
        const int item = lane + 32 * (b0 + i), rr = item / kQ4, f4 = item - rr * kQ4;
        if (f4 * 4 < hpad) {
          const uint32_t off = 2 * kTnAPart + mn_offset(rr, f4, kTnBBlocks);
          split_store(st + off, st + kTnBPart + off, vb[i]);
        }
      }
      if (warp == 1) GHSCN_TR(200 + 5 * (c / 8) + 2);
      fence_proxy_async();
      if (warp == 1) GHSCN_TR(200 + 5 * (c / 8) + 3);
      __syncwarp();
      if (lane == 0) bar_arrive(smem_addr(&full_bar[s]));
      if (warp == 1) GHSCN_TR(200 + 5 * (c / 8) + 4);
    }

    // ----- epilogue: kEpi warps per TMEM lane quadrant, each takes every kEpi-th 16-column group -----
    constexpr int kEpi = kTnProducers / 32 / 4;      // 4
    bar_wait(smem_addr(&accum_bar), 0);
    if (warp == 1) GHSCN_TR(500);
    tc_fence_after();
    const int quad = warp & 3;
    const int half = (warp - 1) >> 2;                // warps 1..4 -> 0, warps 5..8 -> 1, ...
    const int row = quad * 32 + lane;
    const int sstride = hvalid + ((hvalid & 4) ? 0 : 4);       // floats; odd multiple of 4 words: conflict-free float4
    float* srow = reinterpret_cast<float*>(smem_gen) + (size_t)row * sstride;
    const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int groups = (hvalid + 15) / 16;
    for (int gi = half; gi < groups; gi += kEpi) {
      const int c0 = gi * 16;
      float v[16];
      if (nchunks > 0) {
        load_sum16(tbase, (uint32_t)c0, true, v);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col = c0 + 4 * q;
        if (col < hvalid)
          *reinterpret_cast<float4*>(srow + col) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kTnProducers) : "memory");      // both column sets of every row are staged
    // coalesced stores: each warp writes 32 / kEpi rows of its quadrant, hvalid contiguous floats per row
    const int nf4 = hvalid / 4;
    for (int r = 0; r < 32 / kEpi; ++r) {
      const int trow = quad * 32 + half * (32 / kEpi) + r;
      if (m0 + trow < m_out) {
        float* dst = kSegmented ? partial + (int64_t)slab * out_slab_stride + (int64_t)(m0 + trow) * ldo + hcol
                                : partial + ((int64_t)slab * m_out + (m0 + trow)) * n_out + hcol;
        const float* src = reinterpret_cast<const float*>(smem_gen) + (size_t)trow * sstride;
        for (int f = lane; f < nf4; f += 32)
          *reinterpret_cast<float4*>(dst + f * 4) = *reinterpret_cast<const float4*>(src + f * 4);
      }
    }
  }

  if (warp == 1) GHSCN_TR(510);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
  GHSCN_TR_DUMP;
}

// out[e] = sum_s partial[s][e], slabs added in index order (deterministic).
__global__ void __launch_bounds__(256) gemm3x_tn_reduce_kernel(const float* __restrict__ partial, int64_t elems4,
                                                               int nslabs, float* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= elems4) return;
  const float4* p = reinterpret_cast<const float4*>(partial) + i;
  float4 acc = __ldg(p);
  for (int s = 1; s < nslabs; ++s) {
    const float4 v = __ldg(p + (int64_t)s * elems4);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  reinterpret_cast<float4*>(out)[i] = acc;
}

}  // namespace
}  // namespace ghscn

using namespace ghscn;

extern "C" {

int ghscn_gemm3x_supported(int64_t m, int64_t n_out, int64_t k) {
  if (m <= 0 || n_out < 16 || k < 8) return 0;
  if ((n_out % 4) != 0 || (k % 4) != 0) return 0;
  if (n_out > kMaxN) return 0;
  if (m > (int64_t)INT32_MAX - kTileM || k > (1 << 20)) return 0;
  return 1;
}

size_t ghscn_gemm3x_b_image_bytes(int64_t n_out, int64_t k) {
  if (!ghscn_gemm3x_supported(kTileM, n_out, k)) return 0;
  return (size_t)nn_image_bytes(make_nn_plan((int)n_out, (int)k));
}

int ghscn_gemm3x_prep_b(const float* w, int64_t ldw, int64_t n_out, int64_t k, int32_t transpose, void* image,
                        ghscn_stream_t stream) {
  GHSCN_REQUIRE(w != nullptr && image != nullptr);
  if (!ghscn_gemm3x_supported(kTileM, n_out, k)) return GHSCN_E_UNSUPPORTED;
  GHSCN_REQUIRE(ldw >= (transpose ? n_out : k));
  const NnPlan p = make_nn_plan((int)n_out, (int)k);
  const int64_t total = (int64_t)p.kchunks * (p.hv.pad[0] + p.hv.pad[1]) * kChunkK;
  const int64_t blocks = ceil_div<int64_t>(total, 256);
  gemm3x_prep_b_kernel<<<(unsigned)(blocks < 4 * kNumSMs ? blocks : 4 * kNumSMs), 256, 0, as_stream(stream)>>>(
      w, ldw, (int)n_out, (int)k, transpose, p, static_cast<unsigned char*>(image));
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

int ghscn_gemm3x_set_trace(void* device_buffer) {
#ifdef GHSCN_GEMM3X_TRACE
  long long* p = static_cast<long long*>(device_buffer);
  return (int)cudaMemcpyToSymbol(g_trace, &p, sizeof(p));
#else
  (void)device_buffer;
  return GHSCN_E_UNSUPPORTED;
#endif
}

int ghscn_gemm3x(const float* a, int64_t lda, int64_t m, int64_t k, const void* b_image, int64_t n_out,
                 const float* bias, int32_t relu, float* c, int64_t ldc, ghscn_stream_t stream) {
  GHSCN_REQUIRE(a != nullptr && b_image != nullptr && c != nullptr);
  if (!ghscn_gemm3x_supported(m, n_out, k)) return GHSCN_E_UNSUPPORTED;
  if (ldc != n_out || (lda % 4) != 0 || lda < k) return GHSCN_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(c) & 15) ||
      (reinterpret_cast<uintptr_t>(b_image) & 15) || (bias && (reinterpret_cast<uintptr_t>(bias) & 15)))
    return GHSCN_E_UNSUPPORTED;
  const NnPlan p = make_nn_plan((int)n_out, (int)k);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm3x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kNnSmem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const unsigned tiles = (unsigned)ceil_div<int64_t>(m, kTileM);
  // one CTA per row tile walks both N halves (A is re-read from L2 for the second) when the tiles fill exactly one
  // wave; otherwise one CTA per (tile, half): see the note in the kernel
  static const int force_split = [] {
    const char* e = getenv("GHSCN_GEMM3X_SPLIT");       // tuning: 0 = never, 1 = always, unset = by wave count
    return e ? atoi(e) : -1;
  }();
  // a half CTA costs ~0.54 of a full one (its prologue and drain do not shrink): split only where the finer grid
  // saves enough waves -- 156 tiles: 3 x 0.54 vs 2 waves; 1 224 tiles (B = 1024): 17 x 0.54 vs 9 waves -> keep full
  const unsigned waves_full = ceil_div<unsigned>(tiles, (unsigned)kNumSMs);
  const unsigned waves_half = ceil_div<unsigned>(2 * tiles, (unsigned)kNumSMs);
  int split = (p.hv.count == 2) && (0.54 * waves_half < waves_full - 0.05);
  if (force_split >= 0) split = force_split && p.hv.count == 2;
  dim3 grid(tiles, split ? 2u : 1u);
  gemm3x_kernel<<<grid, kNnThreads, kNnSmem, as_stream(stream)>>>(
      a, lda, (int)m, (int)k, static_cast<const unsigned char*>(b_image), (int)n_out, bias, relu, c, p, split);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

int ghscn_gemm3x_tn_supported(int64_t rows, int64_t m_out, int64_t n_out) {
  if (rows <= 0 || rows > (int64_t)INT32_MAX - 64 || m_out < 4 || n_out < 16) return 0;
  if ((m_out % 4) != 0 || (n_out % 4) != 0) return 0;
  if (n_out > kMaxN || m_out > 4096) return 0;
  return 1;
}

size_t ghscn_gemm3x_tn_workspace_bytes(int64_t rows, int64_t m_out, int64_t n_out) {
  if (!ghscn_gemm3x_tn_supported(rows, m_out, n_out)) return 0;
  const TnPlan p = make_tn_plan(rows, (int)m_out, (int)n_out);
  return (size_t)p.nslabs * (size_t)m_out * (size_t)n_out * sizeof(float);
}

int ghscn_gemm3x_tn(const float* p_mat, int64_t ldp, const float* q_mat, int64_t ldq, int64_t rows, int64_t m_out,
                    int64_t n_out, float* out, void* workspace, size_t workspace_bytes, ghscn_stream_t stream) {
  GHSCN_REQUIRE(p_mat != nullptr && q_mat != nullptr && out != nullptr && workspace != nullptr);
  if (!ghscn_gemm3x_tn_supported(rows, m_out, n_out)) return GHSCN_E_UNSUPPORTED;
  if ((ldp % 4) != 0 || (ldq % 4) != 0 || ldp < m_out || ldq < n_out) return GHSCN_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(p_mat) & 15) || (reinterpret_cast<uintptr_t>(q_mat) & 15) ||
      (reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(workspace) & 15))
    return GHSCN_E_UNSUPPORTED;
  const TnPlan p = make_tn_plan(rows, (int)m_out, (int)n_out);
  if (workspace_bytes < (size_t)p.nslabs * (size_t)m_out * (size_t)n_out * sizeof(float)) return GHSCN_E_WORKSPACE;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm3x_tn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTnSmem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  float* partial = p.nslabs == 1 ? out : static_cast<float*>(workspace);
  gemm3x_tn_kernel<false><<<dim3((unsigned)(p.mtiles * p.hv.count), (unsigned)p.nslabs), kTnThreads, kTnSmem,
                            as_stream(stream)>>>(p_mat, ldp, q_mat, ldq, (int)rows, (int)m_out, (int)n_out, partial, p,
                                                 nullptr, 0, 0);
  GHSCN_LAUNCH_CHECK();
  if (p.nslabs > 1) {
    const int64_t elems4 = m_out * n_out / 4;
    gemm3x_tn_reduce_kernel<<<(unsigned)ceil_div<int64_t>(elems4, 256), 256, 0, as_stream(stream)>>>(
        partial, elems4, p.nslabs, out);
    GHSCN_LAUNCH_CHECK();
  }
  return GHSCN_OK;
}

int ghscn_gemm3x_tn_segmented(const float* p_mat, int64_t ldp, const float* q_mat, int64_t ldq,
                              const int32_t* seg_ptr, int64_t num_segments, int64_t max_segment_rows, int64_t m_out,
                              int64_t n_out, float* out, int64_t ldo, int64_t out_segment_stride,
                              ghscn_stream_t stream) {
  GHSCN_REQUIRE(p_mat != nullptr && q_mat != nullptr && out != nullptr && seg_ptr != nullptr);
  GHSCN_REQUIRE(num_segments >= 0 && max_segment_rows >= 0);
  if (num_segments == 0) return GHSCN_OK;
  if (!ghscn_gemm3x_tn_supported(1, m_out, n_out) || num_segments > 65535) return GHSCN_E_UNSUPPORTED;
  if (max_segment_rows > kSlabRows) return GHSCN_E_UNSUPPORTED;      // accuracy: accumulations per TMEM accumulator
  if ((ldp % 4) != 0 || (ldq % 4) != 0 || (ldo % 4) != 0 || (out_segment_stride % 4) != 0 || ldp < m_out ||
      ldq < n_out || ldo < n_out)
    return GHSCN_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(p_mat) & 15) || (reinterpret_cast<uintptr_t>(q_mat) & 15) ||
      (reinterpret_cast<uintptr_t>(out) & 15))
    return GHSCN_E_UNSUPPORTED;
  TnPlan p;
  p.hv = make_halves((int)n_out);
  p.mtiles = (int)((m_out + kTileM - 1) / kTileM);
  p.nslabs = (int)num_segments;
  p.chunks_per_slab = 0;
  cudaError_t e = cudaFuncSetAttribute(gemm3x_tn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTnSmem);
  if (e != cudaSuccess) return (int)e;
  gemm3x_tn_kernel<true><<<dim3((unsigned)(p.mtiles * p.hv.count), (unsigned)num_segments), kTnThreads, kTnSmem,
                           as_stream(stream)>>>(p_mat, ldp, q_mat, ldq, 0, (int)m_out, (int)n_out, out, p, seg_ptr,
                                                out_segment_stride, ldo);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

}  // extern "C"
