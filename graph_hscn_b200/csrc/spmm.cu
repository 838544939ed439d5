// K2/K3: row-parallel SpMM over the batch CSR (see include/ghscn.h).
//
// Roofline: HBM-bound.  Algorithmic bytes per call = 4F(N_src + N_dst) + 8 nnz + 4(N+1)
// (SURVEY.md 8d).  A group of LPR lanes owns one destination row; each lane keeps VEC-wide
// register accumulators for ITERS column chunks, so every feature row touched is read with
// coalesced 128-bit loads and the per-row neighbour sum is sequential in slot (= edge) order with
// separate multiply and add roundings -- the order and rounding of CPU scatter_add_.
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
  GHSCN_REQUIRE(rowptr && col && x && y);
  const bool vec4 = (num_feat % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) &&
                    ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) |
                      reinterpret_cast<uintptr_t>(bias)) % 16 == 0);
  if (vec4)
    return dispatch_spmm<4>(rowptr, col, w, x, ldx, y, ldy, bias, num_rows, num_feat, relu, as_stream(stream));
  return dispatch_spmm<1>(rowptr, col, w, x, ldx, y, ldy, bias, num_rows, num_feat, relu, as_stream(stream));
}

int ghscn_spmm_edge_grad(const int32_t* rowptr, const int32_t* col, const int32_t* perm, const float* x,
                         int64_t ldx, const float* dy, int64_t lddy, int64_t num_rows, int64_t num_feat,
                         int64_t num_edges, float* dw_edge, ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_rows >= 0 && num_feat >= 0 && num_edges >= 0);
  if (num_rows == 0) return GHSCN_OK;
  GHSCN_REQUIRE(rowptr && col && x && dy && dw_edge);
  spmm_edge_grad_kernel<<<(unsigned)ceil_div<int64_t>(num_rows, 8), 256, 0, as_stream(stream)>>>(
      rowptr, col, perm, x, ldx, dy, lddy, (int)num_rows, (int)num_feat, num_edges, dw_edge);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

}  // extern "C"
