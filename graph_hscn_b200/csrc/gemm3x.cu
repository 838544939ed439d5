// Fused 3xTF32 projection GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
//   C[M, N] = A[M, K] . B[N, K]^T (+ bias) (ReLU)          fp32 in, fp32 out, fp32-level accuracy
//
// Replaces, for the h x h projections of GCNConv/GATConv (reference call sites model/mpnn.py:52,59 and
// model/hscn.py:109 -> PyG `Linear` inside the conv), the round-1 pipeline
// "split_tf32_cat (writes 3x the activations) + one library TF32 GEMM over the 3K-long reduction".
// Here the activations are read from HBM exactly once as fp32, split into TF32 hi/lo parts in registers
// and written straight into the swizzled shared-memory operand tiles of tcgen05.mma; nothing but C goes back.
//
//   x = x_hi + x_lo  (hi: low 13 mantissa bits cleared; lo = x - hi, exact)       same split as ghscn_split_tf32
//   C = sum_k  x_lo.w_hi + x_hi.w_lo + x_hi.w_hi                                  (lo.lo dropped, ~2^-22)
//
// One CTA per 128-row tile of A, all N (<= 304 after padding to 16) columns: the fp32 accumulator
// [128 lanes x NPAD columns] lives in TMEM.  Warp roles (192 threads):
//   warp 0 / lane 0   streams the weight image (pre-split, pre-swizzled by gemm3x_prep_b_kernel; L2 resident)
//                     with ONE cp.async.bulk per K chunk into the stage's B buffer (mbarrier complete_tx)
//   warp 1 / lane 0   issues tcgen05.mma.kind::tf32 (3 products x N halves x K steps per chunk), commits the
//                     stage's "empty" mbarrier, finally the accumulator barrier; warp 1 owns the TMEM allocation
//   warps 2..5        A producers: 128-bit global loads (register-prefetched one chunk ahead) -> hi/lo split ->
//                     128B-swizzled K-major smem tiles -> fence.proxy.async -> mbarrier arrive; afterwards the
//                     same four warps are the epilogue: tcgen05.ld (their TMEM lane quadrant) -> bias/ReLU ->
//                     packed row-major staging tile in smem -> one bulk store per warp (32 contiguous rows of C)
// Shared-memory operand layout: the canonical K-major SWIZZLE_128B UMMA layout -- rows of 32 tf32 (128 bytes),
// atoms of 8 rows (1024 bytes, SBO), the 16-byte chunk index XORed with (row & 7).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"

namespace ghscn {
namespace {

constexpr int kTileM = 128;
constexpr int kChunkK = 32;                       // tf32 elements per smem row (128 bytes)
constexpr int kABytes = kTileM * kChunkK * 4;     // one A part (hi or lo) of one stage: 16 KB
constexpr int kMaxStages = 6;
constexpr int kMaxNPad = 304;
constexpr int kSmemLimit = 227 * 1024;
constexpr int kThreads = 192;
constexpr unsigned kSpinLimit = 1u << 22;         // a broken pipeline traps instead of hanging the GPU

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
__device__ __forceinline__ void bulk_store_commit_and_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
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
// D[tmem] (+)= A[smem desc] . B[smem desc]^T, kind::tf32, issued by ONE thread for the CTA.
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All previously issued MMAs of this thread arrive on `bar` once they have completed (implies fence::before).
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread l of the warp receives TMEM lane (quadrant*32 + l).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major: 1) | [32,46) SBO >> 4 (8 rows = 1024 B)
//   [46,48) version = 1 (Blackwell) | [49,52) base offset = 0 (tiles are 1024-byte aligned) | [61,64) layout = 2
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 [4,6)=1, A = B = TF32 [7,10)=[10,13)=2,
// both K-major (bits 15,16 = 0), N >> 3 at [17,23), M >> 4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// byte offset of element (row, k) inside one K-major SW128 block of `kChunkK` columns
__host__ __device__ __forceinline__ uint32_t sw128_offset(int row, int k) {
  const int chunk = k >> 2, e = k & 3, r8 = row & 7;
  return (uint32_t)((row >> 3) * 1024 + r8 * 128 + ((chunk ^ r8) << 4) + e * 4);
}

struct Plan {
  int npad, n0, n1, kchunks, stages, stage_bytes, b_bytes, tmem_cols, smem_bytes;
};

__host__ __device__ inline Plan make_plan(int n_out, int k) {
  Plan p;
  p.npad = (n_out + 15) / 16 * 16;
  if (p.npad <= 256) { p.n0 = p.npad; p.n1 = 0; }
  else { p.n0 = ((p.npad / 2) + 15) / 16 * 16; p.n1 = p.npad - p.n0; }
  p.kchunks = (k + kChunkK - 1) / kChunkK;
  p.b_bytes = 2 * p.npad * kChunkK * 4;            // hi block + lo block of one K chunk
  p.stage_bytes = 2 * kABytes + p.b_bytes;
  int st = (kSmemLimit - 2048) / p.stage_bytes;
  if (st > kMaxStages) st = kMaxStages;
  if (st > p.kchunks) st = p.kchunks;
  p.stages = st;
  int staging = kTileM * n_out * 4;                // epilogue tile, reuses the stage buffers
  int body = p.stages * p.stage_bytes;
  if (body < staging) body = staging;
  p.smem_bytes = body + 1024;                      // slack for the manual 1024-byte alignment
  p.tmem_cols = p.npad <= 32 ? 32 : p.npad <= 64 ? 64 : p.npad <= 128 ? 128 : p.npad <= 256 ? 256 : 512;
  return p;
}

// Weight image: for every K chunk, [NPAD rows x 128 B] of hi parts then the same of lo parts, already in the
// swizzled shared-memory layout, zero padded in N and K.  b[n, k] = transpose ? w[k*ldw + n] : w[n*ldw + k].
__global__ void __launch_bounds__(256) gemm3x_prep_b_kernel(const float* __restrict__ w, int64_t ldw, int n_out,
                                                            int k_dim, int transpose, int npad, int kchunks,
                                                            unsigned char* __restrict__ image) {
  const int64_t total = (int64_t)kchunks * npad * kChunkK;
  const int64_t block_bytes = (int64_t)npad * kChunkK * 4;
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
    unsigned char* base = image + (int64_t)kc * 2 * block_bytes + sw128_offset(n, kk);
    *reinterpret_cast<float*>(base) = hi;
    *reinterpret_cast<float*>(base + block_bytes) = lo;
  }
}

__global__ void __launch_bounds__(kThreads, 1)
gemm3x_kernel(const float* __restrict__ a, int64_t lda, int m_rows, int k_dim, const unsigned char* __restrict__ b_image,
              int n_out, const float* __restrict__ bias, int relu, float* __restrict__ c, Plan plan, int dbg) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full_bar[kMaxStages];
  __shared__ __align__(8) unsigned long long empty_bar[kMaxStages];
  __shared__ __align__(8) unsigned long long accum_bar;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = (smem_addr(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_addr(smem_raw));
  const int m0 = blockIdx.x * kTileM;
  const int stages = plan.stages, kchunks = (dbg & 16) ? 0 : plan.kchunks;

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      bar_init(smem_addr(&full_bar[s]), 4 + 1);     // 4 A-producer warps + the B producer's expect_tx arrive
      bar_init(smem_addr(&empty_bar[s]), 1);        // one tcgen05.commit
    }
    bar_init(smem_addr(&accum_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_addr(&tmem_slot), (uint32_t)plan.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===== weight-image producer =====
    if (lane == 0) {
      for (int kc = 0; kc < kchunks; ++kc) {
        const int s = kc % stages;
        const uint32_t ph = (uint32_t)(kc / stages) & 1u;
        bar_wait(smem_addr(&empty_bar[s]), ph ^ 1u);
        const uint32_t fb = smem_addr(&full_bar[s]);
        if (dbg & 4) { bar_arrive(fb); continue; }
        bar_arrive_expect_tx(fb, (uint32_t)plan.b_bytes);
        bulk_load(smem_base + (uint32_t)s * plan.stage_bytes + 2 * kABytes,
                  b_image + (int64_t)kc * plan.b_bytes, (uint32_t)plan.b_bytes, fb);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc0 = make_idesc_tf32(kTileM, plan.n0);
      const uint32_t idesc1 = make_idesc_tf32(kTileM, plan.n1 > 0 ? plan.n1 : 16);
      const uint32_t b_half = (uint32_t)plan.npad * kChunkK * 4;   // bytes of the hi block (lo block follows)
      const uint32_t n1_off = (uint32_t)plan.n0 * kChunkK * 4;     // rows n0.. of a block
      uint32_t acc = 0;
      for (int kc = 0; kc < kchunks; ++kc) {
        const int s = kc % stages;
        const uint32_t ph = (uint32_t)(kc / stages) & 1u;
        bar_wait(smem_addr(&full_bar[s]), ph);
        tc_fence_after();
        const uint32_t st = smem_base + (uint32_t)s * plan.stage_bytes;
        const uint64_t a_hi = make_desc_k_sw128(st), a_lo = make_desc_k_sw128(st + kABytes);
        const uint64_t b_hi = make_desc_k_sw128(st + 2 * kABytes), b_lo = make_desc_k_sw128(st + 2 * kABytes + b_half);
        const uint64_t b_hi1 = make_desc_k_sw128(st + 2 * kABytes + n1_off);
        const uint64_t b_lo1 = make_desc_k_sw128(st + 2 * kABytes + b_half + n1_off);
        const int kleft = k_dim - kc * kChunkK;
        const int ksteps = kleft >= kChunkK ? kChunkK / 8 : (kleft + 7) / 8;
        for (int ks = 0; ks < ((dbg & 1) ? 0 : ksteps); ++ks) {
          const uint64_t adv = (uint64_t)(ks * 2);       // 8 tf32 = 32 bytes = 2 x 16-byte units
          // small cross terms first, then the main term
          mma_tf32(tmem_base, a_lo + adv, b_hi + adv, idesc0, acc);
          acc = 1;
          mma_tf32(tmem_base, a_hi + adv, b_lo + adv, idesc0, 1);
          mma_tf32(tmem_base, a_hi + adv, b_hi + adv, idesc0, 1);
          if (plan.n1 > 0) {
            const uint32_t d1 = tmem_base + (uint32_t)plan.n0;
            mma_tf32(d1, a_lo + adv, b_hi1 + adv, idesc1, kc > 0 || ks > 0);
            mma_tf32(d1, a_hi + adv, b_lo1 + adv, idesc1, 1);
            mma_tf32(d1, a_hi + adv, b_hi1 + adv, idesc1, 1);
          }
        }
        mma_commit(smem_addr(&empty_bar[s]));
      }
      mma_commit(smem_addr(&accum_bar));
    }
    __syncwarp();
  } else {
    // ===== A producers (warps 2..5), then epilogue =====
    const int t = tid - 64;                 // 0..127
    const int cq = t & 7;                   // 16-byte chunk of the 128-byte row
    const int r0 = t >> 3;                  // rows r0 + 16 j
    const int r8 = r0 & 7;
    const uint32_t row_off = (uint32_t)((r0 >> 3) * 1024 + r8 * 128 + ((cq ^ r8) << 4));
    constexpr int kDepth = 3;               // K chunks of A in flight per thread (register ring)
    float4 ring[kDepth][8];
    auto load_chunk = [&](int kc, float4* dst) {
      const int k = kc * kChunkK + cq * 4;
      const bool kvalid = k < k_dim && !(dbg & 2);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = m0 + r0 + 16 * j;
        dst[j] = (kvalid && r < m_rows) ? ldg_f4(a + (int64_t)r * lda + k) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
#pragma unroll
    for (int d = 0; d < kDepth; ++d)
      if (d < kchunks) load_chunk(d, ring[d]);
    for (int kb = 0; kb < kchunks; kb += kDepth) {
#pragma unroll
      for (int d = 0; d < kDepth; ++d) {
        const int kc = kb + d;
        if (kc < kchunks) {
          const int s = kc % stages;
          const uint32_t ph = (uint32_t)(kc / stages) & 1u;
          bar_wait(smem_addr(&empty_bar[s]), ph ^ 1u);
          unsigned char* st = smem_gen + (size_t)s * plan.stage_bytes + row_off;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 v = ring[d][j];
            float4 h, l;
            h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
            h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
            h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
            h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
            *reinterpret_cast<float4*>(st + j * 2048) = h;              // rows +16 = two 8-row atoms
            *reinterpret_cast<float4*>(st + kABytes + j * 2048) = l;
          }
          fence_proxy_async();                      // every writer: generic-proxy stores -> async proxy (tcgen05.mma)
          __syncwarp();
          if (lane == 0) bar_arrive(smem_addr(&full_bar[s]));   // one arrive per warp: 128 arrives cost ~1 us/chunk
          if (kc + kDepth < kchunks) load_chunk(kc + kDepth, ring[d]);
        }
      }
    }

    // ----- epilogue: TMEM -> registers -> packed [rows, n_out] tile in smem -> bulk store -----
    bar_wait(smem_addr(&accum_bar), 0);
    tc_fence_after();
    const int quad = warp & 3;                                  // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;                           // tile row == TMEM lane
    float* srow = reinterpret_cast<float*>(smem_gen) + (size_t)row * n_out;
    const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16);
    for (int c0 = 0; c0 < ((dbg & 8) ? 16 : n_out); c0 += 16) {
      float v[16];
      tmem_ld16(tbase + (uint32_t)c0, v);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col = c0 + 4 * q;
        if (col < n_out) {                                       // n_out % 4 == 0
          float4 o = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          if (bias != nullptr) {
            const float4 b = ldg_f4(bias + col);
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
          }
          if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          *reinterpret_cast<float4*>(srow + col) = o;
        }
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      int rows = m_rows - (m0 + quad * 32);
      rows = rows > 32 ? 32 : rows;
      if (rows > 0 && !(dbg & 32)) {
        bulk_store(c + (int64_t)(m0 + quad * 32) * n_out, smem_base + (uint32_t)(quad * 32 * n_out * 4),
                   (uint32_t)(rows * n_out * 4));
        bulk_store_commit_and_wait();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)plan.tmem_cols);
  }
}


// =====================================================================================================================
// Weight-gradient GEMM  dW[M, N] = P[R, M]^T . Q[R, N]   (P = dY, Q = X; the reduction runs over the R node rows).
//
// Both operands are "MN-major" for the tensor core (the reduction index is the slow one in memory).  For 32-bit
// MN-major operands the only UMMA layout is SWIZZLE_128B_BASE32B (cute Layout_MN_SW128_32B_Atom): an atom is
// 4 reduction rows x 32 tf32 (128 B) of M/N = 512 B, the 32-byte unit index of a row is XORed with (row & 3);
// atoms of consecutive 32-wide M/N blocks are LBO apart, atoms of consecutive 4-row groups SBO apart, and one
// tcgen05.mma (K = 8) reads two row groups.  128-bit global loads along M/N map onto 16-byte halves of the
// swizzle units, so no transposition is needed.
// Grid = (M tiles of 128) x (row slabs): every CTA reduces its slab of rows into a [128 x NPAD] TMEM accumulator and
// writes an fp32 partial; gemm3x_tn_reduce_kernel adds the slab partials in fixed order (deterministic, and the
// tensor core's truncating accumulator never sees more than kSlabRows rows).
constexpr int kTnGroups = 2;                       // 8-row reduction groups per pipeline stage (16 rows)
constexpr int kTnRows = 8 * kTnGroups;
constexpr int kTnProducers = 256;                  // warps 1..8
constexpr int kTnThreads = 32 + kTnProducers;
constexpr int kTnMaxStages = 8;
constexpr int kSlabRows = 384;                     // upper bound of rows reduced inside one accumulator

struct TnPlan {
  int npad, nblocks, n0, n1, stages, a_part, b_part, stage_bytes, smem_bytes, mtiles, nslabs, chunks_per_slab;
};

__host__ __device__ inline TnPlan make_tn_plan(int64_t rows, int m_out, int n_out) {
  TnPlan p;
  p.npad = (n_out + 15) / 16 * 16;
  p.nblocks = (p.npad + 31) / 32;
  if (p.npad <= 256) { p.n0 = p.npad; p.n1 = 0; }
  else { p.n0 = 160; p.n1 = p.npad - 160; }
  p.a_part = kTnGroups * 2 * 4 * 512;              // per 4-row group: 128 M values = 4 blocks of 512 B
  p.b_part = kTnGroups * 2 * p.nblocks * 512;
  p.stage_bytes = 2 * p.a_part + 2 * p.b_part;
  int st = (kSmemLimit - 2048) / p.stage_bytes;
  if (st > kTnMaxStages) st = kTnMaxStages;
  p.stages = st;
  int staging = kTileM * n_out * 4;
  int body = p.stages * p.stage_bytes;
  if (body < staging) body = staging;
  p.smem_bytes = body + 1024;
  p.mtiles = (m_out + kTileM - 1) / kTileM;
  const int64_t chunks = (rows + kTnRows - 1) / kTnRows;
  int64_t nslabs = kNumSMs / p.mtiles;             // one wave when the slabs are short enough
  const int64_t min_slabs = (rows + kSlabRows - 1) / kSlabRows;
  if (nslabs < min_slabs) nslabs = min_slabs;
  if (nslabs > chunks) nslabs = chunks;
  if (nslabs < 1) nslabs = 1;
  p.chunks_per_slab = (int)((chunks + nslabs - 1) / nslabs);
  p.nslabs = (int)((chunks + p.chunks_per_slab - 1) / p.chunks_per_slab);
  return p;
}

// MN-major SWIZZLE_128B_BASE32B descriptor (layout type 1): LBO = byte distance between 32-element M/N blocks,
// SBO = between 4-row groups of the reduction dimension.
__device__ __forceinline__ uint64_t make_desc_mn_sw128_32b(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
// byte offset of the 16-byte unit `f4` (4 consecutive M/N values) of reduction row `rr` inside an operand part
// laid out as [4-row group][32-wide block][512 B atom]
__device__ __forceinline__ uint32_t mn_offset(int rr, int f4, int nblocks) {
  const int kg = rr >> 2, k4 = rr & 3, c16 = f4 & 7;
  return (uint32_t)((kg * nblocks + (f4 >> 3)) * 512 + k4 * 128 + (((c16 >> 1) ^ k4) << 5) + ((c16 & 1) << 4));
}

__global__ void __launch_bounds__(kTnThreads, 1)
gemm3x_tn_kernel(const float* __restrict__ pmat, int64_t ldp, const float* __restrict__ qmat, int64_t ldq, int rows,
                 int m_out, int n_out, float* __restrict__ partial, TnPlan plan) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full_bar[kTnMaxStages];
  __shared__ __align__(8) unsigned long long empty_bar[kTnMaxStages];
  __shared__ __align__(8) unsigned long long accum_bar;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_base = (smem_addr(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem_base - smem_addr(smem_raw));
  const int m0 = blockIdx.x * kTileM;
  const int slab = blockIdx.y;
  const int stages = plan.stages;
  const int chunk0 = slab * plan.chunks_per_slab;
  const int total_chunks = (rows + kTnRows - 1) / kTnRows;
  int nchunks = total_chunks - chunk0;
  if (nchunks > plan.chunks_per_slab) nchunks = plan.chunks_per_slab;

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      bar_init(smem_addr(&full_bar[s]), kTnProducers / 32);
      bar_init(smem_addr(&empty_bar[s]), 1);
    }
    bar_init(smem_addr(&accum_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_addr(&tmem_slot), 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t mn = (1u << 15) | (1u << 16);             // A and B are MN-major
      const uint32_t idesc0 = make_idesc_tf32(kTileM, plan.n0) | mn;
      const uint32_t idesc1 = make_idesc_tf32(kTileM, plan.n1 > 0 ? plan.n1 : 16) | mn;
      const uint32_t a_sbo = 4 * 512, b_sbo = (uint32_t)plan.nblocks * 512;   // one 4-row group of all blocks
      const uint32_t n1_off = (uint32_t)(plan.n0 / 32) * 512;
      for (int c = 0; c < nchunks; ++c) {
        const int s = c % stages;
        const uint32_t ph = (uint32_t)(c / stages) & 1u;
        bar_wait(smem_addr(&full_bar[s]), ph);
        tc_fence_after();
        const uint32_t st = smem_base + (uint32_t)s * plan.stage_bytes;
        const uint32_t a_hi = st, a_lo = st + plan.a_part;
        const uint32_t b_hi = st + 2 * plan.a_part, b_lo = b_hi + plan.b_part;
#pragma unroll
        for (int g = 0; g < kTnGroups; ++g) {          // 8 reduction rows = two 4-row groups per MMA
          const uint32_t ao = g * 2 * a_sbo, bo = g * 2 * b_sbo;
          const uint64_t dah = make_desc_mn_sw128_32b(a_hi + ao, 512, a_sbo);
          const uint64_t dal = make_desc_mn_sw128_32b(a_lo + ao, 512, a_sbo);
          const uint64_t dbh = make_desc_mn_sw128_32b(b_hi + bo, 512, b_sbo);
          const uint64_t dbl = make_desc_mn_sw128_32b(b_lo + bo, 512, b_sbo);
          const uint32_t acc = (c > 0 || g > 0) ? 1u : 0u;
          mma_tf32(tmem_base, dal, dbh, idesc0, acc);
          mma_tf32(tmem_base, dah, dbl, idesc0, 1);
          mma_tf32(tmem_base, dah, dbh, idesc0, 1);
          if (plan.n1 > 0) {
            const uint32_t d1 = tmem_base + (uint32_t)plan.n0;
            const uint64_t dbh1 = make_desc_mn_sw128_32b(b_hi + bo + n1_off, 512, b_sbo);
            const uint64_t dbl1 = make_desc_mn_sw128_32b(b_lo + bo + n1_off, 512, b_sbo);
            mma_tf32(d1, dal, dbh1, idesc1, acc);
            mma_tf32(d1, dah, dbl1, idesc1, 1);
            mma_tf32(d1, dah, dbh1, idesc1, 1);
          }
        }
        mma_commit(smem_addr(&empty_bar[s]));
      }
      mma_commit(smem_addr(&accum_bar));
    }
    __syncwarp();
  } else {
    // ===== producers (warps 1..8): P slice [16 rows x 128] and Q [16 rows x NPAD] per chunk, then epilogue =====
    const int t = tid - 32;                          // 0..255
    constexpr int kAItems = kTnRows * 32 / kTnProducers;          // float4 items of P per thread and chunk (2)
    const int q4 = plan.npad / 4;                    // float4 per row of Q (zero padded)
    const int b_items = kTnRows * q4;                // <= 16 * 76
    constexpr int kBMax = (kTnRows * (kMaxNPad / 4) + kTnProducers - 1) / kTnProducers;   // 5
    constexpr int kDepth = 4;                        // chunks of global loads in flight per thread (register ring)
    float4 ra[kDepth][kAItems], rb[kDepth][kBMax];
    auto load_chunk = [&](int c, float4* da, float4* db) {
      const int r_base = (chunk0 + c) * kTnRows;
#pragma unroll
      for (int i = 0; i < kAItems; ++i) {
        const int item = t + i * kTnProducers;       // row = item / 32, f4 = item % 32
        const int r = r_base + (item >> 5), col = m0 + (item & 31) * 4;
        da[i] = (r < rows && col < m_out) ? ldg_f4(pmat + (int64_t)r * ldp + col) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < kBMax; ++i) {
        const int item = t + i * kTnProducers;
        const int rr = item / q4, f4 = item - rr * q4;
        const int r = r_base + rr, col = f4 * 4;
        db[i] = (item < b_items && r < rows && col < n_out) ? ldg_f4(qmat + (int64_t)r * ldq + col)
                                                            : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto split_store = [&](unsigned char* hi_base, uint32_t part, uint32_t off, const float4 v) {
      float4 h, l;
      h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
      h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
      h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
      h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
      *reinterpret_cast<float4*>(hi_base + off) = h;
      *reinterpret_cast<float4*>(hi_base + part + off) = l;
    };
#pragma unroll
    for (int d = 0; d < kDepth; ++d)
      if (d < nchunks) load_chunk(d, ra[d], rb[d]);
    for (int cb = 0; cb < nchunks; cb += kDepth) {
#pragma unroll
      for (int d = 0; d < kDepth; ++d) {
        const int c = cb + d;
        if (c < nchunks) {
          const int s = c % stages;
          const uint32_t ph = (uint32_t)(c / stages) & 1u;
          bar_wait(smem_addr(&empty_bar[s]), ph ^ 1u);
          unsigned char* st = smem_gen + (size_t)s * plan.stage_bytes;
#pragma unroll
          for (int i = 0; i < kAItems; ++i) {
            const int item = t + i * kTnProducers;
            const int rr = item >> 5, f4 = item & 31;    // row in chunk, 16-byte unit along M
            split_store(st, (uint32_t)plan.a_part, mn_offset(rr, f4, 4), ra[d][i]);
          }
#pragma unroll
          for (int i = 0; i < kBMax; ++i) {
            const int item = t + i * kTnProducers;
            if (item < b_items) {
              const int rr = item / q4, f4 = item - rr * q4;
              split_store(st + 2 * plan.a_part, (uint32_t)plan.b_part, mn_offset(rr, f4, plan.nblocks), rb[d][i]);
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) bar_arrive(smem_addr(&full_bar[s]));
          if (c + kDepth < nchunks) load_chunk(c + kDepth, ra[d], rb[d]);
        }
      }
    }

    // ----- epilogue: two warps per TMEM lane quadrant, each takes half of the 16-column groups -----
    bar_wait(smem_addr(&accum_bar), 0);
    tc_fence_after();
    const int quad = warp & 3;
    const int half = (warp - 1) >> 2;                // warps 1..4 -> 0, warps 5..8 -> 1
    const int row = quad * 32 + lane;
    float* srow = reinterpret_cast<float*>(smem_gen) + (size_t)row * n_out;
    const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int groups = (n_out + 15) / 16;
    for (int gi = half; gi < groups; gi += 2) {
      const int c0 = gi * 16;
      float v[16];
      if (nchunks > 0) {
        tmem_ld16(tbase + (uint32_t)c0, v);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col = c0 + 4 * q;
        if (col < n_out)
          *reinterpret_cast<float4*>(srow + col) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
    }
    fence_proxy_async();
    asm volatile("bar.sync 1, %0;" ::"n"(kTnProducers) : "memory");      // both column halves of every row are staged
    if (half == 0 && lane == 0) {
      int nrow = m_out - (m0 + quad * 32);
      nrow = nrow > 32 ? 32 : nrow;
      if (nrow > 0) {
        float* dst = partial + ((int64_t)slab * m_out + (m0 + quad * 32)) * n_out;
        bulk_store(dst, smem_base + (uint32_t)(quad * 32 * n_out * 4), (uint32_t)(nrow * n_out * 4));
        bulk_store_commit_and_wait();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
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
  if ((n_out + 15) / 16 * 16 > kMaxNPad) return 0;
  if (m > (int64_t)INT32_MAX - kTileM || k > (1 << 20)) return 0;
  return 1;
}

size_t ghscn_gemm3x_b_image_bytes(int64_t n_out, int64_t k) {
  if (!ghscn_gemm3x_supported(kTileM, n_out, k)) return 0;
  const Plan p = make_plan((int)n_out, (int)k);
  return (size_t)p.kchunks * (size_t)p.b_bytes;
}

int ghscn_gemm3x_prep_b(const float* w, int64_t ldw, int64_t n_out, int64_t k, int32_t transpose, void* image,
                        ghscn_stream_t stream) {
  GHSCN_REQUIRE(w != nullptr && image != nullptr);
  if (!ghscn_gemm3x_supported(kTileM, n_out, k)) return GHSCN_E_UNSUPPORTED;
  GHSCN_REQUIRE(ldw >= (transpose ? n_out : k));
  const Plan p = make_plan((int)n_out, (int)k);
  const int64_t total = (int64_t)p.kchunks * p.npad * kChunkK;
  const int64_t blocks = ceil_div<int64_t>(total, 256);
  gemm3x_prep_b_kernel<<<(unsigned)(blocks < 4 * kNumSMs ? blocks : 4 * kNumSMs), 256, 0, as_stream(stream)>>>(
      w, ldw, (int)n_out, (int)k, transpose, p.npad, p.kchunks, static_cast<unsigned char*>(image));
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

int ghscn_gemm3x(const float* a, int64_t lda, int64_t m, int64_t k, const void* b_image, int64_t n_out,
                 const float* bias, int32_t relu, float* c, int64_t ldc, ghscn_stream_t stream) {
  GHSCN_REQUIRE(a != nullptr && b_image != nullptr && c != nullptr);
  if (!ghscn_gemm3x_supported(m, n_out, k)) return GHSCN_E_UNSUPPORTED;
  if (ldc != n_out || (lda % 4) != 0 || lda < k) return GHSCN_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(c) & 15) ||
      (reinterpret_cast<uintptr_t>(b_image) & 15) || (bias && (reinterpret_cast<uintptr_t>(bias) & 15)))
    return GHSCN_E_UNSUPPORTED;
  const Plan p = make_plan((int)n_out, (int)k);
  static bool attr_set = false;
  static int dbg = 0;   // GHSCN_GEMM3X_DEBUG: timing experiments only (1 no MMA, 2 no A loads, 4 no B loads, 8 short epilogue)
  if (!attr_set) {
    const char* ev = getenv("GHSCN_GEMM3X_DEBUG");
    dbg = ev ? atoi(ev) : 0;
    cudaError_t e = cudaFuncSetAttribute(gemm3x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit - 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const unsigned grid = (unsigned)ceil_div<int64_t>(m, kTileM);
  gemm3x_kernel<<<grid, kThreads, p.smem_bytes, as_stream(stream)>>>(
      a, lda, (int)m, (int)k, static_cast<const unsigned char*>(b_image), (int)n_out, bias, relu, c, p, dbg);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

int ghscn_gemm3x_tn_supported(int64_t rows, int64_t m_out, int64_t n_out) {
  if (rows <= 0 || rows > (int64_t)INT32_MAX - 64 || m_out < 4 || n_out < 16) return 0;
  if ((m_out % 4) != 0 || (n_out % 4) != 0) return 0;
  if ((n_out + 15) / 16 * 16 > kMaxNPad || m_out > 4096) return 0;
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
    cudaError_t e = cudaFuncSetAttribute(gemm3x_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit - 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  float* partial = p.nslabs == 1 ? out : static_cast<float*>(workspace);
  gemm3x_tn_kernel<<<dim3((unsigned)p.mtiles, (unsigned)p.nslabs), kTnThreads, p.smem_bytes, as_stream(stream)>>>(
      p_mat, ldp, q_mat, ldq, (int)rows, (int)m_out, (int)n_out, partial, p);
  GHSCN_LAUNCH_CHECK();
  if (p.nslabs > 1) {
    const int64_t elems4 = m_out * n_out / 4;
    gemm3x_tn_reduce_kernel<<<(unsigned)ceil_div<int64_t>(elems4, 256), 256, 0, as_stream(stream)>>>(
        partial, elems4, p.nslabs, out);
    GHSCN_LAUNCH_CHECK();
  }
  return GHSCN_OK;
}

}  // extern "C"
