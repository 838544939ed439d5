// K2/K3: row-parallel SpMM over the batch CSR (see include/ghscn.h).
//
// Roofline: HBM-bound.  Algorithmic bytes per call = 4F(N_src + N_dst) + 8 nnz + 4(N+1)
// (SURVEY.md 8d).  A group of LPR lanes owns one destination row; each lane keeps VEC-wide
// register accumulators for ITERS column chunks, so every feature row touched is read with
// coalesced 128-bit loads and the per-row neighbour sum is sequential in slot (= edge) order with
// separate multiply and add roundings -- the order and rounding of CPU scatter_add_.
#include <stdlib.h>

#include "common.cuh"

namespace ghscn {

template <int VEC>
struct Vec;
template <>
struct Vec<4> {
  float4 v;
  __device__ __forceinline__ void zero() { v = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void load(const float* p) { v = __ldg(reinterpret_cast<const float4*>(p)); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = v; }
  __device__ __forceinline__ void fma_unfused(float w, const Vec<4>& x) {
    v.x = mul_then_add(v.x, w, x.v.x);
    v.y = mul_then_add(v.y, w, x.v.y);
    v.z = mul_then_add(v.z, w, x.v.z);
    v.w = mul_then_add(v.w, w, x.v.w);
  }
  __device__ __forceinline__ void add(const Vec<4>& x) {
    v.x = __fadd_rn(v.x, x.v.x);
    v.y = __fadd_rn(v.y, x.v.y);
    v.z = __fadd_rn(v.z, x.v.z);
    v.w = __fadd_rn(v.w, x.v.w);
  }
  __device__ __forceinline__ void relu() {
    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
  }
  __device__ __forceinline__ float dot(const Vec<4>& x) const {
    return v.x * x.v.x + v.y * x.v.y + v.z * x.v.z + v.w * x.v.w;
  }
};
template <>
struct Vec<1> {
  float v;
  __device__ __forceinline__ void zero() { v = 0.f; }
  __device__ __forceinline__ void load(const float* p) { v = __ldg(p); }
  __device__ __forceinline__ void store(float* p) const { *p = v; }
  __device__ __forceinline__ void fma_unfused(float w, const Vec<1>& x) { v = mul_then_add(v, w, x.v); }
  __device__ __forceinline__ void add(const Vec<1>& x) { v = __fadd_rn(v, x.v); }
  __device__ __forceinline__ void relu() { v = fmaxf(v, 0.f); }
  __device__ __forceinline__ float dot(const Vec<1>& x) const { return v * x.v; }
};

// grid.x covers rows (32/LPR rows per warp), grid.y covers feature tiles of LPR*VEC*ITERS columns.
template <int VEC, int LPR, int ITERS, bool WEIGHTED>
__global__ void __launch_bounds__(256) spmm_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                                                   const float* __restrict__ w, const float* __restrict__ x,
                                                   int64_t ldx, float* __restrict__ y, int64_t ldy,
                                                   const float* __restrict__ bias, int num_rows, int num_feat,
                                                   int relu) {
  constexpr int kRowsPerWarp = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int warp = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  const int row = warp * kRowsPerWarp + lane / LPR;
  const int sub = lane % LPR;
  if (row >= num_rows) return;
  const int f0 = blockIdx.y * (LPR * VEC * ITERS) + sub * VEC;

  Vec<VEC> acc[ITERS];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) acc[it].zero();

  const int beg = rowptr[row], end = rowptr[row + 1];
  // The row's lane group fetches up to LPR (col, w) pairs with ONE coalesced load each and hands them out
  // by shuffle: the dependent chain per row is rowptr -> col/w -> features instead of one hop per slot.
  const unsigned group_mask = (LPR == 32) ? kFullMask : (((1u << LPR) - 1u) << ((lane / LPR) * LPR));
  for (int base = beg; base < end; base += LPR) {
    int my_c = 0;
    float my_w = 1.f;
    if (base + sub < end) {
      my_c = col[base + sub];
      if (WEIGHTED) my_w = w[base + sub];
    }
    const int cnt = min(LPR, end - base);
    int j = 0;
    // two slots per trip: all feature loads first, then the ordered (slot-order) accumulation
    for (; j + 1 < cnt; j += 2) {
      const int c0 = __shfl_sync(group_mask, my_c, j, LPR), c1 = __shfl_sync(group_mask, my_c, j + 1, LPR);
      const float w0 = __shfl_sync(group_mask, my_w, j, LPR), w1 = __shfl_sync(group_mask, my_w, j + 1, LPR);
      const float* x0 = x + (int64_t)c0 * ldx + f0;
      const float* x1 = x + (int64_t)c1 * ldx + f0;
      Vec<VEC> a[ITERS], b[ITERS];
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int f = f0 + it * LPR * VEC;
        if (f < num_feat) { a[it].load(x0 + it * LPR * VEC); b[it].load(x1 + it * LPR * VEC); }
      }
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int f = f0 + it * LPR * VEC;
        if (f < num_feat) {
          if (WEIGHTED) { acc[it].fma_unfused(w0, a[it]); acc[it].fma_unfused(w1, b[it]); }
          else { acc[it].add(a[it]); acc[it].add(b[it]); }
        }
      }
    }
    if (j < cnt) {
      const int c0 = __shfl_sync(group_mask, my_c, j, LPR);
      const float w0 = __shfl_sync(group_mask, my_w, j, LPR);
      const float* x0 = x + (int64_t)c0 * ldx + f0;
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int f = f0 + it * LPR * VEC;
        if (f < num_feat) {
          Vec<VEC> a;
          a.load(x0 + it * LPR * VEC);
          if (WEIGHTED) acc[it].fma_unfused(w0, a); else acc[it].add(a);
        }
      }
    }
  }
  float* yrow = y + (int64_t)row * ldy + f0;
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int f = f0 + it * LPR * VEC;
    if (f < num_feat) {
      if (bias) { Vec<VEC> bv; bv.load(bias + f); acc[it].add(bv); }
      if (relu) acc[it].relu();
      acc[it].store(yrow + it * LPR * VEC);
    }
  }
}

// ---- wide rows (F >= 128, 128-bit path): CTA-staged indices ------------------------------------------
// A CTA owns kRowsPerCta consecutive rows.  Their rowptr entries and the whole contiguous (col, w) slot range
// are staged in shared memory with two coalesced passes, so no warp has an index load on its critical path;
// each warp then walks its rows issuing the feature loads of up to 4 slots (4 x ITERS 128-bit loads per lane)
// before the ordered, unfused accumulation.  This keeps several KB per warp in flight, which is what a
// gather with ~3 slots per row needs to approach HBM/L2 bandwidth.
constexpr int kRowsPerCta = 64;
constexpr int kSlotCap = 1024;
// Rows per CTA of the register-gather kernel: 8 rows per warp, ROWS / 8 warps per CTA.  Small CTAs (32 rows = 4
// warps) let the register file hold 5 CTAs = 20 warps per SM instead of 2 x 8, and cut the grid into pieces fine
// enough that a batch a few per cent larger does not fall off a wave boundary (64-row CTAs at 2 per SM: 296 slots,
// so N = 18.3 k ran in one wave and N = 19.6 k in two -- 12.2 us vs 18.5 us, scripts/spmm_probe.py).
#ifndef GHSCN_WIDE_ROWS
#define GHSCN_WIDE_ROWS 32
#endif
static inline int wide_rows() {
  static const int v = [] {
    const char* e = getenv("GHSCN_SPMM_ROWS");          // tuning experiments: 16 | 32 | 64
    const int r = e ? atoi(e) : 0;
    return (r == 16 || r == 32 || r == 64) ? r : GHSCN_WIDE_ROWS;
  }();
  return v;
}

// MASKED: every gathered row is multiplied by (mask[row] > 0) as it is read -- the backward of a ReLU that was fused
// into the forward aggregation (dy (.) [y > 0]) without a separate pass over dy.
template <int ITERS, bool WEIGHTED, int ROWS, bool MASKED = false>
__global__ void __launch_bounds__(ROWS * 4, ROWS == 64 ? 2 : (ROWS == 32 ? 5 : 10))   // 16 / 20 / 20 warps per SM
    spmm_wide_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const float* __restrict__ w,
                     const float* __restrict__ x, int64_t ldx, float* __restrict__ y, int64_t ldy,
                     const float* __restrict__ bias, int num_rows, int num_feat, int relu,
                     const float* __restrict__ mask = nullptr, int64_t ldm = 0) {
  constexpr int kWideRows = ROWS, kThreads = ROWS * 4, kSlotCap = ROWS * 16;
  __shared__ int s_rowptr[kWideRows + 1];
  __shared__ int s_col[kSlotCap];
  __shared__ float s_w[kSlotCap];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int row0 = blockIdx.x * kWideRows;
  const int nrows = min(kWideRows, num_rows - row0);
  if (tid <= nrows) s_rowptr[tid] = rowptr[row0 + tid];
  __syncthreads();
  const int sbeg = s_rowptr[0];
  const int staged = min(s_rowptr[nrows] - sbeg, kSlotCap);
  for (int i = tid; i < staged; i += kThreads) {
    s_col[i] = col[sbeg + i];
    if (WEIGHTED) s_w[i] = w[sbeg + i];
  }
  __syncthreads();
  const int f0 = blockIdx.y * (128 * ITERS) + lane * 4;

  // Every warp owns kWideRows / 8 CONSECUTIVE rows, i.e. one contiguous slot range, and walks it four slots at a time
  // whatever rows they belong to: molecule-like graphs have ~2 slots per row, so batching per row would leave half of
  // the loads unissued.  Slots are still accumulated strictly in slot order into their own row (a row is written out
  // when the walk leaves it), so the arithmetic per row is unchanged.
  constexpr int kRowsPerWarp = 8;                             // ROWS / 8 warps per CTA
  const int rbeg = wid * kRowsPerWarp, rend = min(nrows, rbeg + kRowsPerWarp);
  if (rbeg >= nrows) return;
  float4 acc[ITERS];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) acc[it] = make_float4(0.f, 0.f, 0.f, 0.f);
  auto flush = [&](int r) {
    float* yrow = y + (int64_t)(row0 + r) * ldy + f0;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      if (f0 + it * 128 < num_feat) {
        float4 o = acc[it];
        if (bias) {
          const float4 b = ldg_f4(bias + f0 + it * 128);
          o.x = __fadd_rn(o.x, b.x); o.y = __fadd_rn(o.y, b.y); o.z = __fadd_rn(o.z, b.z); o.w = __fadd_rn(o.w, b.w);
        }
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        *reinterpret_cast<float4*>(yrow + it * 128) = o;
      }
      acc[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  int r = rbeg;
  int row_end = s_rowptr[r + 1] - sbeg;                       // first slot after the current row
  const int s_first = s_rowptr[rbeg] - sbeg, s_last = s_rowptr[rend] - sbeg;
  for (int s = s_first; s < s_last; s += 4) {
    const int n4 = min(4, s_last - s);
    float4 v[4][ITERS];
    float wv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < n4) {
        const int slot = s + k;
        const int c = slot < kSlotCap ? s_col[slot] : col[sbeg + slot];
        wv[k] = WEIGHTED ? (slot < kSlotCap ? s_w[slot] : w[sbeg + slot]) : 1.f;
        const float* xr = x + (int64_t)c * ldx + f0;
#pragma unroll
        for (int it = 0; it < ITERS; ++it)
          if (f0 + it * 128 < num_feat) {
            v[k][it] = ldg_f4(xr + it * 128);
            if (MASKED) {
              const float4 m = ldg_f4(mask + (int64_t)c * ldm + f0 + it * 128);
              v[k][it].x = m.x > 0.f ? v[k][it].x : 0.f;
              v[k][it].y = m.y > 0.f ? v[k][it].y : 0.f;
              v[k][it].z = m.z > 0.f ? v[k][it].z : 0.f;
              v[k][it].w = m.w > 0.f ? v[k][it].w : 0.f;
            }
          }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < n4) {
        while (s + k >= row_end) {                            // the walk leaves row r (and any empty rows after it)
          flush(r);
          ++r;
          row_end = s_rowptr[r + 1] - sbeg;
        }
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
          if (f0 + it * 128 < num_feat) {
            if (WEIGHTED) {
              acc[it].x = mul_then_add(acc[it].x, wv[k], v[k][it].x);
              acc[it].y = mul_then_add(acc[it].y, wv[k], v[k][it].y);
              acc[it].z = mul_then_add(acc[it].z, wv[k], v[k][it].z);
              acc[it].w = mul_then_add(acc[it].w, wv[k], v[k][it].w);
            } else {
              acc[it].x = __fadd_rn(acc[it].x, v[k][it].x);
              acc[it].y = __fadd_rn(acc[it].y, v[k][it].y);
              acc[it].z = __fadd_rn(acc[it].z, v[k][it].z);
              acc[it].w = __fadd_rn(acc[it].w, v[k][it].w);
            }
          }
        }
      }
    }
  }
  for (; r < rend; ++r) flush(r);                             // the last row with slots and trailing empty rows
}

// ---- wide rows, TMA bulk-copy gather (sm_100a: cp.async.bulk + mbarrier, SASS UBLKCP) -----------------------------
// Same CTA-staged indices as above, but the feature rows themselves are fetched by the TMA engine: every warp owns
// a private ring of kBulkStages shared-memory row buffers, each guarded by an mbarrier.  A source row x[col[s],:]
// is one contiguous, 16-byte aligned span of 4F bytes, so one `cp.async.bulk.shared::cluster.global` per slot
// moves it without touching registers; the warp keeps kBulkStages rows (~9.6 KB at F=300) in flight, reads a
// landed row with conflict-free 128-bit shared loads, accumulates in slot order with unfused mul/add (identical
// arithmetic to the register path), then re-arms the same stage for the slot kBulkStages ahead.  No cross-warp
// synchronisation after the index staging; ~150 KB of gather traffic in flight per SM at 2 CTAs/SM.
constexpr int kBulkStages = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}

template <int ITERS, bool WEIGHTED>
__global__ void __launch_bounds__(256, 2) spmm_bulk_kernel(const int* __restrict__ rowptr,
                                                           const int* __restrict__ col,
                                                           const float* __restrict__ w,
                                                           const float* __restrict__ x, int64_t ldx,
                                                           float* __restrict__ y, int64_t ldy,
                                                           const float* __restrict__ bias, int num_rows,
                                                           int num_feat, int relu) {
  extern __shared__ __align__(128) unsigned char ring[];  // [8 warps][kBulkStages][row_bytes]
  __shared__ __align__(8) unsigned long long bars[8 * kBulkStages];
  __shared__ int s_rowptr[kRowsPerCta + 1];
  __shared__ int s_col[kSlotCap];
  __shared__ float s_w[kSlotCap];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int row0 = blockIdx.x * kRowsPerCta;
  const int nrows = min(kRowsPerCta, num_rows - row0);
  const uint32_t row_bytes = (uint32_t)num_feat * 4u;
  if (tid < 8 * kBulkStages) mbar_init(smem_u32(&bars[tid]), 1);
  if (tid <= nrows) s_rowptr[tid] = rowptr[row0 + tid];
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const int sbeg = s_rowptr[0];
  const int staged = min(s_rowptr[nrows] - sbeg, kSlotCap);
  for (int i = tid; i < staged; i += 256) {
    s_col[i] = col[sbeg + i];
    if (WEIGHTED) s_w[i] = w[sbeg + i];
  }
  __syncthreads();

  unsigned char* my_ring = ring + (size_t)wid * kBulkStages * row_bytes;
  const uint32_t bar0 = smem_u32(&bars[wid * kBulkStages]);
  // issue cursor (runs kBulkStages slots ahead of the consume cursor); warp-uniform
  int ir = wid, is = (ir < nrows) ? s_rowptr[ir] - sbeg : 0;
  int issued = 0;
  auto issue_next = [&]() -> bool {   // lane 0 arms the stage and launches one row copy
    while (ir < nrows && is >= s_rowptr[ir + 1] - sbeg) {
      ir += 8;
      if (ir < nrows) is = s_rowptr[ir] - sbeg;
    }
    if (ir >= nrows) return false;
    const int stage = issued % kBulkStages;
    if (lane == 0) {
      const int c = is < kSlotCap ? s_col[is] : col[sbeg + is];
      const uint32_t bar = bar0 + stage * 8;
      mbar_arrive_expect_tx(bar, row_bytes);
      bulk_g2s(smem_u32(my_ring + (size_t)stage * row_bytes), x + (int64_t)c * ldx, row_bytes, bar);
    }
    ++is;
    ++issued;
    return true;
  };
  for (int k = 0; k < kBulkStages; ++k)
    if (!issue_next()) break;

  int consumed = 0;
  for (int r = wid; r < nrows; r += 8) {
    const int beg = s_rowptr[r] - sbeg, end = s_rowptr[r + 1] - sbeg;
    float4 acc[ITERS];
#pragma unroll
    for (int it = 0; it < ITERS; ++it) acc[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = beg; s < end; ++s) {
      const int stage = consumed % kBulkStages;
      mbar_wait(bar0 + stage * 8, (uint32_t)(consumed / kBulkStages) & 1u);
      const float wv = WEIGHTED ? (s < kSlotCap ? s_w[s] : w[sbeg + s]) : 1.f;
      const float4* src = reinterpret_cast<const float4*>(my_ring + (size_t)stage * row_bytes);
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        if ((it * 32 + lane) * 4 < num_feat) {
          const float4 v = src[it * 32 + lane];
          if (WEIGHTED) {
            acc[it].x = mul_then_add(acc[it].x, wv, v.x); acc[it].y = mul_then_add(acc[it].y, wv, v.y);
            acc[it].z = mul_then_add(acc[it].z, wv, v.z); acc[it].w = mul_then_add(acc[it].w, wv, v.w);
          } else {
            acc[it].x = __fadd_rn(acc[it].x, v.x); acc[it].y = __fadd_rn(acc[it].y, v.y);
            acc[it].z = __fadd_rn(acc[it].z, v.z); acc[it].w = __fadd_rn(acc[it].w, v.w);
          }
        }
      }
      ++consumed;
      __syncwarp();     // every lane has read this stage before the TMA engine may overwrite it
      issue_next();
    }
    float* yrow = y + (int64_t)(row0 + r) * ldy;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int f = (it * 32 + lane) * 4;
      if (f < num_feat) {
        float4 o = acc[it];
        if (bias) {
          const float4 b = ldg_f4(bias + f);
          o.x = __fadd_rn(o.x, b.x); o.y = __fadd_rn(o.y, b.y); o.z = __fadd_rn(o.z, b.z); o.w = __fadd_rn(o.w, b.w);
        }
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        *reinterpret_cast<float4*>(yrow + f) = o;
      }
    }
  }
}

template <int ITERS>
static int launch_spmm_bulk(const int* rowptr, const int* col, const float* w, const float* x, int64_t ldx,
                            float* y, int64_t ldy, const float* bias, int64_t num_rows, int64_t num_feat, int relu,
                            cudaStream_t stream) {
  const size_t shm = (size_t)8 * kBulkStages * num_feat * 4;
  dim3 grid((unsigned)ceil_div<int64_t>(num_rows, kRowsPerCta));
  if (w) {
    cudaFuncSetAttribute(spmm_bulk_kernel<ITERS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
    spmm_bulk_kernel<ITERS, true><<<grid, 256, shm, stream>>>(rowptr, col, w, x, ldx, y, ldy, bias, (int)num_rows,
                                                               (int)num_feat, relu);
  } else {
    cudaFuncSetAttribute(spmm_bulk_kernel<ITERS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
    spmm_bulk_kernel<ITERS, false><<<grid, 256, shm, stream>>>(rowptr, col, w, x, ldx, y, ldy, bias, (int)num_rows,
                                                                (int)num_feat, relu);
  }
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

// ---- wide rows, cp.async (LDGSTS) gather: deep per-lane prefetch without registers ----------------------------------
// Each lane copies exactly the 16-byte chunks it will later read (chunk it*32+lane of a row), one commit group per
// slot, kAsyncStages slots ahead, into a warp-private shared-memory ring.  Consumption only needs
// cp.async.wait_group (a lane waits for its own copies), no barrier and no shuffle; arithmetic and order are the
// same unfused slot-order accumulation as the register path.
constexpr int kAsyncStages = 10;

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int ITERS, bool WEIGHTED>
__global__ void __launch_bounds__(256, 2) spmm_async_kernel(const int* __restrict__ rowptr,
                                                            const int* __restrict__ col,
                                                            const float* __restrict__ w,
                                                            const float* __restrict__ x, int64_t ldx,
                                                            float* __restrict__ y, int64_t ldy,
                                                            const float* __restrict__ bias, int num_rows,
                                                            int num_feat, int relu) {
  extern __shared__ __align__(128) unsigned char ring[];  // [8 warps][kAsyncStages][row_bytes]
  __shared__ int s_rowptr[kRowsPerCta + 1];
  __shared__ int s_col[kSlotCap];
  __shared__ float s_w[kSlotCap];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int row0 = blockIdx.x * kRowsPerCta;
  const int nrows = min(kRowsPerCta, num_rows - row0);
  const uint32_t row_bytes = (uint32_t)num_feat * 4u;
  if (tid <= nrows) s_rowptr[tid] = rowptr[row0 + tid];
  __syncthreads();
  const int sbeg = s_rowptr[0];
  const int staged = min(s_rowptr[nrows] - sbeg, kSlotCap);
  for (int i = tid; i < staged; i += 256) {
    s_col[i] = col[sbeg + i];
    if (WEIGHTED) s_w[i] = w[sbeg + i];
  }
  __syncthreads();

  unsigned char* my_ring = ring + (size_t)wid * kAsyncStages * row_bytes;
  const uint32_t ring_u32 = smem_u32(my_ring);
  int ir = wid, is = (ir < nrows) ? s_rowptr[ir] - sbeg : 0;   // issue cursor, warp-uniform
  int issued = 0;
  auto issue_next = [&]() {   // always commits one group (possibly empty) so the pending count stays constant
    while (ir < nrows && is >= s_rowptr[ir + 1] - sbeg) {
      ir += 8;
      if (ir < nrows) is = s_rowptr[ir] - sbeg;
    }
    if (ir < nrows) {
      const int c = is < kSlotCap ? s_col[is] : col[sbeg + is];
      const float* src = x + (int64_t)c * ldx;
      const uint32_t dst = ring_u32 + (uint32_t)(issued % kAsyncStages) * row_bytes;
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int f = (it * 32 + lane) * 4;
        if (f < num_feat) cp_async_16(dst + f * 4, src + f);
      }
      ++is;
      ++issued;
    }
    cp_async_commit();
  };
#pragma unroll 1
  for (int k = 0; k < kAsyncStages - 1; ++k) issue_next();

  int consumed = 0;
  for (int r = wid; r < nrows; r += 8) {
    const int beg = s_rowptr[r] - sbeg, end = s_rowptr[r + 1] - sbeg;
    float4 acc[ITERS];
#pragma unroll
    for (int it = 0; it < ITERS; ++it) acc[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = beg; s < end; ++s) {
      issue_next();                             // keep kAsyncStages groups outstanding ...
      cp_async_wait<kAsyncStages - 1>();        // ... and wait for the oldest one (this slot)
      const float wv = WEIGHTED ? (s < kSlotCap ? s_w[s] : w[sbeg + s]) : 1.f;
      const float4* src = reinterpret_cast<const float4*>(my_ring + (size_t)(consumed % kAsyncStages) * row_bytes);
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        if ((it * 32 + lane) * 4 < num_feat) {
          const float4 v = src[it * 32 + lane];   // the chunk this very lane copied
          if (WEIGHTED) {
            acc[it].x = mul_then_add(acc[it].x, wv, v.x); acc[it].y = mul_then_add(acc[it].y, wv, v.y);
            acc[it].z = mul_then_add(acc[it].z, wv, v.z); acc[it].w = mul_then_add(acc[it].w, wv, v.w);
          } else {
            acc[it].x = __fadd_rn(acc[it].x, v.x); acc[it].y = __fadd_rn(acc[it].y, v.y);
            acc[it].z = __fadd_rn(acc[it].z, v.z); acc[it].w = __fadd_rn(acc[it].w, v.w);
          }
        }
      }
      ++consumed;
    }
    float* yrow = y + (int64_t)(row0 + r) * ldy;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int f = (it * 32 + lane) * 4;
      if (f < num_feat) {
        float4 o = acc[it];
        if (bias) {
          const float4 b = ldg_f4(bias + f);
          o.x = __fadd_rn(o.x, b.x); o.y = __fadd_rn(o.y, b.y); o.z = __fadd_rn(o.z, b.z); o.w = __fadd_rn(o.w, b.w);
        }
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        *reinterpret_cast<float4*>(yrow + f) = o;
      }
    }
  }
  cp_async_wait<0>();
}

template <int ITERS>
static int launch_spmm_async(const int* rowptr, const int* col, const float* w, const float* x, int64_t ldx,
                             float* y, int64_t ldy, const float* bias, int64_t num_rows, int64_t num_feat, int relu,
                             cudaStream_t stream) {
  const size_t shm = (size_t)8 * kAsyncStages * num_feat * 4;
  dim3 grid((unsigned)ceil_div<int64_t>(num_rows, kRowsPerCta));
  if (w) {
    cudaFuncSetAttribute(spmm_async_kernel<ITERS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
    spmm_async_kernel<ITERS, true><<<grid, 256, shm, stream>>>(rowptr, col, w, x, ldx, y, ldy, bias, (int)num_rows,
                                                                (int)num_feat, relu);
  } else {
    cudaFuncSetAttribute(spmm_async_kernel<ITERS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
    spmm_async_kernel<ITERS, false><<<grid, 256, shm, stream>>>(rowptr, col, w, x, ldx, y, ldy, bias, (int)num_rows,
                                                                 (int)num_feat, relu);
  }
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

template <int ITERS, int ROWS>
static int launch_spmm_wide_rows(const int* rowptr, const int* col, const float* w, const float* x, int64_t ldx,
                                 float* y, int64_t ldy, const float* bias, int64_t num_rows, int64_t num_feat,
                                 int relu, cudaStream_t stream) {
  dim3 grid((unsigned)ceil_div<int64_t>(num_rows, ROWS), (unsigned)ceil_div<int64_t>(num_feat, 128 * ITERS));
  if (w)
    spmm_wide_kernel<ITERS, true, ROWS><<<grid, ROWS * 4, 0, stream>>>(rowptr, col, w, x, ldx, y, ldy, bias,
                                                                       (int)num_rows, (int)num_feat, relu);
  else
    spmm_wide_kernel<ITERS, false, ROWS><<<grid, ROWS * 4, 0, stream>>>(rowptr, col, w, x, ldx, y, ldy, bias,
                                                                        (int)num_rows, (int)num_feat, relu);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

template <int ITERS>
static int launch_spmm_wide(const int* rowptr, const int* col, const float* w, const float* x, int64_t ldx,
                            float* y, int64_t ldy, const float* bias, int64_t num_rows, int64_t num_feat, int relu,
                            cudaStream_t stream) {
  switch (wide_rows()) {
    case 16: return launch_spmm_wide_rows<ITERS, 16>(rowptr, col, w, x, ldx, y, ldy, bias, num_rows, num_feat, relu, stream);
    case 64: return launch_spmm_wide_rows<ITERS, 64>(rowptr, col, w, x, ldx, y, ldy, bias, num_rows, num_feat, relu, stream);
    default: return launch_spmm_wide_rows<ITERS, 32>(rowptr, col, w, x, ldx, y, ldy, bias, num_rows, num_feat, relu, stream);
  }
}

template <int ITERS>
static int launch_spmm_wide_masked(const int* rowptr, const int* col, const float* w, const float* x, int64_t ldx,
                                   const float* mask, int64_t ldm, float* y, int64_t ldy, int64_t num_rows,
                                   int64_t num_feat, cudaStream_t stream) {
  dim3 grid((unsigned)ceil_div<int64_t>(num_rows, 32), (unsigned)ceil_div<int64_t>(num_feat, 128 * ITERS));
  if (w)
    spmm_wide_kernel<ITERS, true, 32, true><<<grid, 128, 0, stream>>>(rowptr, col, w, x, ldx, y, ldy, nullptr,
                                                                      (int)num_rows, (int)num_feat, 0, mask, ldm);
  else
    spmm_wide_kernel<ITERS, false, 32, true><<<grid, 128, 0, stream>>>(rowptr, col, w, x, ldx, y, ldy, nullptr,
                                                                       (int)num_rows, (int)num_feat, 0, mask, ldm);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

// ---- few long rows (pooling relations: every destination has many sources) ------------------------------
// One CTA per row: its 8 warps stride the row's slots, each keeps full-width partial sums, and the partials are
// combined in fixed warp order (deterministic; not the sequential CPU order, which pooled rows do not need).
template <int ITERS, bool WEIGHTED>
__global__ void __launch_bounds__(256) spmm_longrow_kernel(const int* __restrict__ rowptr,
                                                           const int* __restrict__ col,
                                                           const float* __restrict__ w,
                                                           const float* __restrict__ x, int64_t ldx,
                                                           float* __restrict__ y, int64_t ldy,
                                                           const float* __restrict__ bias, int num_feat) {
  __shared__ float4 part[8][ITERS * 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int row = blockIdx.x;
  const int f0 = blockIdx.y * (128 * ITERS) + lane * 4;
  const int beg = rowptr[row], end = rowptr[row + 1];
  float4 acc[ITERS];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) acc[it] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = beg + wid; s < end; s += 16) {
    const bool two = s + 8 < end;
    const int c0 = col[s], c1 = two ? col[s + 8] : c0;
    const float w0 = WEIGHTED ? w[s] : 1.f, w1 = (WEIGHTED && two) ? w[s + 8] : 1.f;
    float4 a[ITERS], b[ITERS];
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      if (f0 + it * 128 < num_feat) {
        a[it] = ldg_f4(x + (int64_t)c0 * ldx + f0 + it * 128);
        b[it] = ldg_f4(x + (int64_t)c1 * ldx + f0 + it * 128);
      }
    }
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      if (f0 + it * 128 < num_feat) {
        acc[it].x += w0 * a[it].x; acc[it].y += w0 * a[it].y; acc[it].z += w0 * a[it].z; acc[it].w += w0 * a[it].w;
        if (two) {
          acc[it].x += w1 * b[it].x; acc[it].y += w1 * b[it].y; acc[it].z += w1 * b[it].z; acc[it].w += w1 * b[it].w;
        }
      }
    }
  }
#pragma unroll
  for (int it = 0; it < ITERS; ++it) part[wid][it * 32 + lane] = acc[it];
  __syncthreads();
  for (int i = threadIdx.x; i < ITERS * 32; i += 256) {
    const int f = blockIdx.y * (128 * ITERS) + (i / 32) * 128 + (i % 32) * 4;
    if (f < num_feat) {
      float4 t = part[0][i];
#pragma unroll
      for (int k = 1; k < 8; ++k) { t.x += part[k][i].x; t.y += part[k][i].y; t.z += part[k][i].z; t.w += part[k][i].w; }
      if (bias) { const float4 bb = ldg_f4(bias + f); t.x += bb.x; t.y += bb.y; t.z += bb.z; t.w += bb.w; }
      *reinterpret_cast<float4*>(y + (int64_t)row * ldy + f) = t;
    }
  }
}

int spmm_long_rows(const int* rowptr, const int* col, const float* w, const float* x, int64_t ldx, float* y,
                   int64_t ldy, const float* bias, int64_t num_rows, int64_t num_feat, cudaStream_t stream) {
  // caller guarantees the 128-bit path (num_feat % 4 == 0, 16-byte aligned rows)
  if (num_rows == 0 || num_feat == 0) return GHSCN_OK;
#define GHSCN_LONG(ITERS)                                                                                        \
  do {                                                                                                           \
    dim3 grid((unsigned)num_rows, (unsigned)ceil_div<int64_t>(num_feat, 128 * ITERS));                           \
    if (w) spmm_longrow_kernel<ITERS, true><<<grid, 256, 0, stream>>>(rowptr, col, w, x, ldx, y, ldy, bias,      \
                                                                      (int)num_feat);                            \
    else spmm_longrow_kernel<ITERS, false><<<grid, 256, 0, stream>>>(rowptr, col, w, x, ldx, y, ldy, bias,       \
                                                                     (int)num_feat);                             \
  } while (0)
  if (num_feat <= 128) GHSCN_LONG(1);
  else if (num_feat <= 256) GHSCN_LONG(2);
  else GHSCN_LONG(3);
#undef GHSCN_LONG
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

template <int VEC, int LPR, int ITERS>
static int launch_spmm(const int* rowptr, const int* col, const float* w, const float* x, int64_t ldx, float* y,
                       int64_t ldy, const float* bias, int64_t num_rows, int64_t num_feat, int relu,
                       cudaStream_t stream) {
  constexpr int kRowsPerWarp = 32 / LPR;
  constexpr int kThreads = 256;
  const int64_t warps = ceil_div<int64_t>(num_rows, kRowsPerWarp);
  dim3 grid((unsigned)ceil_div<int64_t>(warps, kThreads / 32),
            (unsigned)ceil_div<int64_t>(num_feat, LPR * VEC * ITERS));
  if (w)
    spmm_kernel<VEC, LPR, ITERS, true><<<grid, kThreads, 0, stream>>>(rowptr, col, w, x, ldx, y, ldy, bias,
                                                                       (int)num_rows, (int)num_feat, relu);
  else
    spmm_kernel<VEC, LPR, ITERS, false><<<grid, kThreads, 0, stream>>>(rowptr, col, w, x, ldx, y, ldy, bias,
                                                                        (int)num_rows, (int)num_feat, relu);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

template <int VEC>
static int dispatch_spmm(const int* rowptr, const int* col, const float* w, const float* x, int64_t ldx, float* y,
                         int64_t ldy, const float* bias, int64_t num_rows, int64_t num_feat, int relu,
                         cudaStream_t stream) {
  const int64_t nvec = ceil_div<int64_t>(num_feat, VEC);
  // gather engine for 128 <= F <= 512, selectable for A/B measurements (scripts/spmm_ceiling.py):
  //   'w' register gather with CTA-staged indices (default: fastest on molecule-like graphs, 13.5 us at
  //       N=18,269 F=300 vs 15.0 us for 'a' and 18.3 us for 'b', CUDA-graph replay over 438 MB of operands)
  //   'a' cp.async ring, 'b' TMA bulk copies (both win only at ~1 slot per row: 10.0 / 9.7 us vs 12.8 us)
  static const char path = [] {
    const char* e = getenv("GHSCN_SPMM_PATH");
    return (e && (e[0] == 'a' || e[0] == 'b')) ? e[0] : 'w';
  }();
  if (VEC == 4 && nvec >= 32 && nvec <= 128 && path == 'a') {
    if (nvec <= 32) return launch_spmm_async<1>(rowptr, col, w, x, ldx, y, ldy, bias, num_rows, num_feat, relu, stream);
    if (nvec <= 64) return launch_spmm_async<2>(rowptr, col, w, x, ldx, y, ldy, bias, num_rows, num_feat, relu, stream);
    if (nvec <= 96) return launch_spmm_async<3>(rowptr, col, w, x, ldx, y, ldy, bias, num_rows, num_feat, relu, stream);
    return launch_spmm_async<4>(rowptr, col, w, x, ldx, y, ldy, bias, num_rows, num_feat, relu, stream);
  }
  if (VEC == 4 && nvec >= 32 && nvec <= 128 && path == 'b') {
    if (nvec <= 32) return launch_spmm_bulk<1>(rowptr, col, w, x, ldx, y, ldy, bias, num_rows, num_feat, relu, stream);
    if (nvec <= 64) return launch_spmm_bulk<2>(rowptr, col, w, x, ldx, y, ldy, bias, num_rows, num_feat, relu, stream);
    if (nvec <= 96) return launch_spmm_bulk<3>(rowptr, col, w, x, ldx, y, ldy, bias, num_rows, num_feat, relu, stream);
    return launch_spmm_bulk<4>(rowptr, col, w, x, ldx, y, ldy, bias, num_rows, num_feat, relu, stream);
  }
  if (VEC == 4 && nvec >= 32) {  // F >= 128 on the 128-bit path: CTA-staged indices, 4 slots in flight
    if (nvec <= 32) return launch_spmm_wide<1>(rowptr, col, w, x, ldx, y, ldy, bias, num_rows, num_feat, relu, stream);
    if (nvec <= 64) return launch_spmm_wide<2>(rowptr, col, w, x, ldx, y, ldy, bias, num_rows, num_feat, relu, stream);
    if (nvec <= 96) return launch_spmm_wide<3>(rowptr, col, w, x, ldx, y, ldy, bias, num_rows, num_feat, relu, stream);
    return launch_spmm_wide<4>(rowptr, col, w, x, ldx, y, ldy, bias, num_rows, num_feat, relu, stream);
  }
#define GHSCN_SPMM(LPR, ITERS) \
  return launch_spmm<VEC, LPR, ITERS>(rowptr, col, w, x, ldx, y, ldy, bias, num_rows, num_feat, relu, stream)
  if (nvec <= 4) GHSCN_SPMM(4, 1);
  if (nvec <= 8) GHSCN_SPMM(8, 1);
  if (nvec <= 16) GHSCN_SPMM(16, 1);
  if (nvec <= 32) GHSCN_SPMM(32, 1);
  if (nvec <= 64) GHSCN_SPMM(32, 2);
  if (nvec <= 96) GHSCN_SPMM(32, 3);
  GHSCN_SPMM(32, 4);  // wider rows: tiled over grid.y in chunks of 128 vectors
#undef GHSCN_SPMM
}

// One warp per row; per slot a warp-wide dot product <dy[row], x[col]>.
__global__ void __launch_bounds__(256) spmm_edge_grad_kernel(const int* __restrict__ rowptr,
                                                             const int* __restrict__ col,
                                                             const int* __restrict__ perm,
                                                             const float* __restrict__ x, int64_t ldx,
                                                             const float* __restrict__ dy, int64_t lddy,
                                                             int num_rows, int num_feat, int64_t num_edges,
                                                             float* __restrict__ dw_edge) {
  const int lane = threadIdx.x & 31;
  const int row = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  if (row >= num_rows) return;
  const float* dyr = dy + (int64_t)row * lddy;
  for (int s = rowptr[row]; s < rowptr[row + 1]; ++s) {
    const float* xr = x + (int64_t)col[s] * ldx;
    float acc = 0.f;
    for (int f = lane; f < num_feat; f += 32) acc += dyr[f] * __ldg(xr + f);
    acc = warp_sum(acc);
    const int e = perm ? perm[s] : s;  // perm == NULL: per-slot gradient
    if (lane == 0 && e < num_edges) dw_edge[e] = acc;
  }
}

}  // namespace ghscn

using namespace ghscn;

extern "C" {

int ghscn_spmm(const int32_t* rowptr, const int32_t* col, const float* w, const float* x, int64_t ldx, float* y,
               int64_t ldy, const float* bias, int64_t num_rows, int64_t num_feat, int32_t relu,
               ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_rows >= 0 && num_feat >= 0 && ldx >= num_feat && ldy >= num_feat);
  GHSCN_REQUIRE(num_rows < ((int64_t)1 << 31) && num_feat < ((int64_t)1 << 24));
  if (num_rows == 0 || num_feat == 0) return GHSCN_OK;
  GHSCN_REQUIRE(rowptr && x && y);  // col may be null for an empty relation (zero slots)
  const bool vec4 = (num_feat % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) &&
                    ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) |
                      reinterpret_cast<uintptr_t>(bias)) % 16 == 0);
  if (vec4)
    return dispatch_spmm<4>(rowptr, col, w, x, ldx, y, ldy, bias, num_rows, num_feat, relu, as_stream(stream));
  return dispatch_spmm<1>(rowptr, col, w, x, ldx, y, ldy, bias, num_rows, num_feat, relu, as_stream(stream));
}

int ghscn_spmm_masked_supported(int64_t num_feat, int64_t ldx, int64_t ldm, int64_t ldy) {
  return num_feat >= 128 && num_feat <= 512 && num_feat % 4 == 0 && ldx % 4 == 0 && ldm % 4 == 0 && ldy % 4 == 0;
}

int ghscn_spmm_masked(const int32_t* rowptr, const int32_t* col, const float* w, const float* x, int64_t ldx,
                      const float* mask, int64_t ldm, float* y, int64_t ldy, int64_t num_rows, int64_t num_feat,
                      ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_rows >= 0 && num_feat > 0 && ldx >= num_feat && ldy >= num_feat && ldm >= num_feat);
  GHSCN_REQUIRE(num_rows < ((int64_t)1 << 31));
  if (!ghscn_spmm_masked_supported(num_feat, ldx, ldm, ldy)) return GHSCN_E_UNSUPPORTED;
  if (num_rows == 0) return GHSCN_OK;
  GHSCN_REQUIRE(rowptr && x && y && mask);
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(mask)) % 16 != 0)
    return GHSCN_E_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  const int64_t nvec = num_feat / 4;
  if (nvec <= 32) return launch_spmm_wide_masked<1>(rowptr, col, w, x, ldx, mask, ldm, y, ldy, num_rows, num_feat, st);
  if (nvec <= 64) return launch_spmm_wide_masked<2>(rowptr, col, w, x, ldx, mask, ldm, y, ldy, num_rows, num_feat, st);
  if (nvec <= 96) return launch_spmm_wide_masked<3>(rowptr, col, w, x, ldx, mask, ldm, y, ldy, num_rows, num_feat, st);
  return launch_spmm_wide_masked<4>(rowptr, col, w, x, ldx, mask, ldm, y, ldy, num_rows, num_feat, st);
}

int ghscn_spmm_pool(const int32_t* rowptr, const int32_t* col, const float* w, const float* x, int64_t ldx, float* y,
                    int64_t ldy, const float* bias, int64_t num_rows, int64_t num_feat, ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_rows >= 0 && num_feat >= 0 && ldx >= num_feat && ldy >= num_feat);
  if (num_rows == 0 || num_feat == 0) return GHSCN_OK;
  GHSCN_REQUIRE(rowptr && x && y);  // col may be null for an empty relation (zero slots)
  const bool vec4 = (num_feat % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) &&
                    ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) |
                      reinterpret_cast<uintptr_t>(bias)) % 16 == 0);
  if (vec4 && num_feat >= 64 && num_rows <= 65535)
    return spmm_long_rows(rowptr, col, w, x, ldx, y, ldy, bias, num_rows, num_feat, as_stream(stream));
  return ghscn_spmm(rowptr, col, w, x, ldx, y, ldy, bias, num_rows, num_feat, 0, stream);
}

int ghscn_spmm_edge_grad(const int32_t* rowptr, const int32_t* col, const int32_t* perm, const float* x,
                         int64_t ldx, const float* dy, int64_t lddy, int64_t num_rows, int64_t num_feat,
                         int64_t num_edges, float* dw_edge, ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_rows >= 0 && num_feat >= 0 && num_edges >= 0);
  if (num_rows == 0) return GHSCN_OK;
  GHSCN_REQUIRE(rowptr && x && dy && (dw_edge || num_edges == 0));
  spmm_edge_grad_kernel<<<(unsigned)ceil_div<int64_t>(num_rows, 8), 256, 0, as_stream(stream)>>>(
      rowptr, col, perm, x, ldx, dy, lddy, (int)num_rows, (int)num_feat, num_edges, dw_edge);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

}  // extern "C"
