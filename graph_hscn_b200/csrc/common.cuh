// Shared device/host helpers for libghscn (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ghscn.h"

#define GHSCN_REQUIRE(cond) \
  do {                      \
    if (!(cond)) return GHSCN_E_INVALID; \
  } while (0)

// Checks the launch status and accounts `n` kernel launches in the library-wide counter
// (ghscn_launch_count(); bench.py reports it as gpu_launches).
#define GHSCN_LAUNCH_CHECK_N(n)                \
  do {                                         \
    cudaError_t e__ = cudaPeekAtLastError();   \
    if (e__ != cudaSuccess) return (int)e__;   \
    ghscn::note_launches(n);                   \
  } while (0)
#define GHSCN_LAUNCH_CHECK() GHSCN_LAUNCH_CHECK_N(1)

namespace ghscn {

void note_launches(int n);  // defined in csr.cu
// spmm.cu: y[r,:] = sum_s w[s] x[col[s],:] (+ bias) for relations with few, long rows (128-bit path only)
int spmm_long_rows(const int* rowptr, const int* col, const float* w, const float* x, int64_t ldx, float* y,
                   int64_t ldy, const float* bias, int64_t num_rows, int64_t num_feat, cudaStream_t stream);

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs
constexpr unsigned kFullMask = 0xffffffffu;

static inline cudaStream_t as_stream(ghscn_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFullMask, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}

// Block-wide sum; `red` must hold >= 32 floats of shared memory. All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}

// 128-bit read-only global load (streaming operand: keep out of L1).
__device__ __forceinline__ float4 ldg_f4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

// Non-fused multiply-add with the rounding sequence of CPU `msg = w * x; out += msg`.
__device__ __forceinline__ float mul_then_add(float acc, float w, float x) {
  return __fadd_rn(acc, __fmul_rn(w, x));
}

}  // namespace ghscn
