// Tall-skinny dense projections: y = x W^T + b with a tiny input width (K <= 32), N in the tens of thousands.
// These are the raw-feature layers of the reference (9 OGB atom features -> hidden: GCNConv layer 1 at
// model/hscn.py:88-93, GraphConv lin_rel / lin_root at model/hscn.py:32-34, the SCN cluster MLP at
// model/hscn.py:50-54).  They are pure HBM streaming problems (read or write one [N, M] matrix once, ~50 MFLOP),
// which a tiled SIMT GEMM handles poorly (7-36 us each in profiles/r1d); here each is one pass at memory speed.
//   fwd : y[n,m]  = sum_k x[n,k] W[m,k] + b[m]          W^T staged in shared memory, warp per row, lanes over m
//   dW  : dW[m,k] = sum_n dY[n,m] x[n,k]                two-stage fixed-order reduction over 512-row chunks
//   dX  : dx[n,k] = sum_m dY[n,m] W[m,k]                warp per row, lanes over k
#include "common.cuh"

namespace ghscn {

constexpr int kSkinnyMaxK = 32;
constexpr int kDwRowFloats = 8192;                            // smem floats for the staged x rows of one chunk
__host__ __device__ constexpr int dw_rows(int kmax) { return kDwRowFloats / kmax; }   // 512 rows (K <= 16) / 256

__global__ void __launch_bounds__(256) skinny_fwd_kernel(const float* __restrict__ x, int64_t ldx,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         int num_rows, int K, int M, float* __restrict__ y,
                                                         int64_t ldy) {
  extern __shared__ float wt[];  // [K][M] (transposed: consecutive lanes read consecutive m)
  for (int i = threadIdx.x; i < K * M; i += blockDim.x) {
    const int m = i / K, k = i - m * K;
    wt[k * M + m] = w[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < num_rows; n += warps) {
    const float xv = lane < K ? __ldg(x + (int64_t)n * ldx + lane) : 0.f;
    for (int m0 = 0; m0 < M; m0 += 32) {  // warp-uniform trip count: every lane takes part in the shuffles
      const int m = m0 + lane;
      const bool on = m < M;
      float acc = (on && bias) ? __ldg(bias + m) : 0.f;
#pragma unroll 4
      for (int k = 0; k < K; ++k) {
        const float xk = __shfl_sync(kFullMask, xv, k);
        if (on) acc = fmaf(xk, wt[k * M + m], acc);
      }
      if (on) y[(int64_t)n * ldy + m] = acc;
    }
  }
}

// Register-resident variant for K <= 12 (the 9 raw atom features): lane l owns the outputs m = l + 32 j and keeps
// their J x K weights in registers, so a row costs K shuffles + J*K FMAs and no shared-memory traffic (the smem
// version issues one LDS per FMA and sits on the LSU limit at ~2.5 TB/s of output).
template <int J>
__global__ void __launch_bounds__(256) skinny_fwd_reg_kernel(const float* __restrict__ x, int64_t ldx,
                                                             const float* __restrict__ w,
                                                             const float* __restrict__ bias, int num_rows, int K,
                                                             int M, float* __restrict__ y, int64_t ldy) {
  constexpr int KR = 12;
  extern __shared__ float wt[];  // [K][M]: staged once per CTA with coalesced loads, then read conflict-free
  for (int i = threadIdx.x; i < K * M; i += blockDim.x) {
    const int m = i / K, k = i - m * K;
    wt[k * M + m] = w[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float wr[J][KR], br[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int m = lane + 32 * j;
    br[j] = (m < M && bias) ? __ldg(bias + m) : 0.f;
#pragma unroll
    for (int k = 0; k < KR; ++k) wr[j][k] = (m < M && k < K) ? wt[k * M + m] : 0.f;
  }
  const int warps = (gridDim.x * blockDim.x) >> 5;     // row stride between a warp's consecutive rows
  // x rows are fetched in groups of 4, one group ahead of the FMAs: with a single row per warp in flight the
  // kernel is latency-bound (1.6 TB/s of output)
  constexpr int G = 4;
  int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  float xa[G], xb[G];
  auto load_group = [&](int base, float* xx) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int r = base + g * warps;
      xx[g] = (r < num_rows && lane < K) ? __ldg(x + (int64_t)r * ldx + lane) : 0.f;
    }
  };
  auto compute_group = [&](int base, const float* xx) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int r = base + g * warps;
      if (r < num_rows) {
        float acc[J];
#pragma unroll
        for (int j = 0; j < J; ++j) acc[j] = br[j];
#pragma unroll
        for (int k = 0; k < KR; ++k) {
          const float xk = __shfl_sync(kFullMask, xx[g], k);    // lanes >= K hold 0: padded k contribute nothing
#pragma unroll
          for (int j = 0; j < J; ++j) acc[j] = fmaf(xk, wr[j][k], acc[j]);
        }
        float* yr = y + (int64_t)r * ldy + lane;
#pragma unroll
        for (int j = 0; j < J; ++j)
          if (lane + 32 * j < M) yr[32 * j] = acc[j];
      }
    }
  };
  load_group(n, xa);
  for (; n < num_rows; n += 2 * G * warps) {
    load_group(n + G * warps, xb);
    compute_group(n, xa);
    load_group(n + 2 * G * warps, xa);
    compute_group(n + G * warps, xb);
  }
}

// partial[chunk][m][k] = sum over the chunk's rows of dY[n,m] x[n,k].  CTA = (32-column tile of m, chunk of
// kDwRows rows): lanes own m, the 8 warps stride the rows (8 dY loads in flight per lane), K register
// accumulators per thread, fixed-order combine of the 8 warps in shared memory.
template <int KMAX>
__global__ void __launch_bounds__(256) skinny_dw_partial_kernel(const float* __restrict__ dy, int64_t lddy,
                                                                const float* __restrict__ x, int64_t ldx,
                                                                int num_rows, int K, int M,
                                                                float* __restrict__ partial) {
  constexpr int kDwRows = dw_rows(KMAX);
  constexpr int kBuf = (kDwRows * KMAX > 8 * 32 * (KMAX + 1)) ? kDwRows * KMAX : 8 * 32 * (KMAX + 1);
  __shared__ __align__(16) float buf[kBuf];                    // x rows first, then reused for the combine
  float* xs = buf;
  float (*red)[32][KMAX + 1] = reinterpret_cast<float (*)[32][KMAX + 1]>(buf);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int r0 = blockIdx.y * kDwRows;
  const int rows = min(kDwRows, num_rows - r0);
  for (int i = threadIdx.x; i < rows * KMAX; i += blockDim.x) {   // columns K..KMAX-1 are zero padding
    const int r = i / KMAX, k = i - r * KMAX;
    xs[r * KMAX + k] = k < K ? __ldg(x + (int64_t)(r0 + r) * ldx + k) : 0.f;
  }
  __syncthreads();
  const int m = blockIdx.x * 32 + lane;
  const bool on = m < M;
  float acc[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) acc[k] = 0.f;
  if (on) {
    const float* dp = dy + (int64_t)r0 * lddy + m;
    int r = wid;
    for (; r + 56 < rows; r += 64) {
      float d[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) d[u] = __ldg(dp + (int64_t)(r + 8 * u) * lddy);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        // the x row is a warp-wide broadcast: 128-bit shared loads (ceil(K/4) instead of K per row) keep the
        // kernel off the LSU limit (9 scalar LDS per 128 B of dY capped it at ~1.2 TB/s)
        const float4* xr = reinterpret_cast<const float4*>(xs + (r + 8 * u) * KMAX);
#pragma unroll
        for (int q = 0; q < KMAX / 4; ++q) {
          if (4 * q < K) {
            const float4 xv = xr[q];
            acc[4 * q + 0] = fmaf(d[u], xv.x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(d[u], xv.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(d[u], xv.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(d[u], xv.w, acc[4 * q + 3]);
          }
        }
      }
    }
    for (; r < rows; r += 8) {
      const float d0 = __ldg(dp + (int64_t)r * lddy);
      const float4* xr = reinterpret_cast<const float4*>(xs + r * KMAX);
#pragma unroll
      for (int q = 0; q < KMAX / 4; ++q) {
        if (4 * q < K) {
          const float4 xv = xr[q];
          acc[4 * q + 0] = fmaf(d0, xv.x, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(d0, xv.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(d0, xv.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(d0, xv.w, acc[4 * q + 3]);
        }
      }
    }
  }
  __syncthreads();  // every warp is done with xs before the buffer is reused
#pragma unroll
  for (int k = 0; k < KMAX; ++k) red[wid][lane][k] = acc[k];
  __syncthreads();
  // 32 x K outputs of this CTA, combined over the 8 warps in fixed order
  for (int i = threadIdx.x; i < 32 * K; i += blockDim.x) {
    const int l = i / K, k = i - l * K;
    const int mm = blockIdx.x * 32 + l;
    if (mm < M) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w][l][k];
      partial[((int64_t)blockIdx.y * M + mm) * K + k] = t;
    }
  }
}

// Register-resident dW for K <= 16 (M <= 128) / K <= 12 (M <= 320): one CTA per chunk of rows and per 160 columns
// (lane l owns m = m0 + l + 32 j, J x K accumulators in registers), so x is staged nowhere and read once, dY rows are read as full
// contiguous rows (J coalesced 128-byte loads per warp and row, the next row prefetched), and the only shared-memory
// traffic is the fixed-order combine of the 8 warps at the end.  partial[chunk][m][k].
template <int J, int KR>
__global__ void __launch_bounds__(256) skinny_dw_rows_kernel(const float* __restrict__ dy, int64_t lddy,
                                                             const float* __restrict__ x, int64_t ldx,
                                                             int num_rows, int K, int M, int rows_per_cta,
                                                             float* __restrict__ partial) {
  __shared__ float red[8][32][KR + 1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int r0 = blockIdx.x * rows_per_cta;
  const int rows = min(rows_per_cta, num_rows - r0);
  const int m0 = blockIdx.y * (32 * J);                       // this CTA's column range: m0 + lane + 32 j
  float acc[J][KR];
#pragma unroll
  for (int j = 0; j < J; ++j)
#pragma unroll
    for (int k = 0; k < KR; ++k) acc[j][k] = 0.f;
  constexpr int D = 2;                                        // rows in flight ahead of the FMAs
  float dq[D][J], xq[D];
  auto load_row = [&](int r, float* dd, float& xx) {
    if (r < rows) {
      const float* dp = dy + (int64_t)(r0 + r) * lddy + m0 + lane;
#pragma unroll
      for (int j = 0; j < J; ++j) dd[j] = (m0 + lane + 32 * j < M) ? __ldg(dp + 32 * j) : 0.f;
      xx = lane < K ? __ldg(x + (int64_t)(r0 + r) * ldx + lane) : 0.f;
    } else {
#pragma unroll
      for (int j = 0; j < J; ++j) dd[j] = 0.f;
      xx = 0.f;
    }
  };
#pragma unroll
  for (int q = 0; q < D; ++q) load_row(wid + 8 * q, dq[q], xq[q]);
  for (int rb = wid; rb < rows; rb += 8 * D) {
#pragma unroll
    for (int q = 0; q < D; ++q) {
      float d[J];
      const float xv = xq[q];
#pragma unroll
      for (int j = 0; j < J; ++j) d[j] = dq[q][j];
      load_row(rb + 8 * (q + D), dq[q], xq[q]);
#pragma unroll
      for (int k = 0; k < KR; ++k) {
        const float xk = __shfl_sync(kFullMask, xv, k);     // lanes >= K hold 0; rows past the end are zeros
#pragma unroll
        for (int j = 0; j < J; ++j) acc[j][k] = fmaf(d[j], xk, acc[j][k]);
      }
    }
  }
  // fixed-order combine of the 8 warps, one 32-column group at a time
#pragma unroll
  for (int j = 0; j < J; ++j) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < KR; ++k) red[wid][lane][k] = acc[j][k];
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * K; i += blockDim.x) {
      const int l = i / K, k = i - l * K;
      const int m = m0 + l + 32 * j;
      if (m < M) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][l][k];
        partial[((int64_t)blockIdx.x * M + m) * K + k] = t;
      }
    }
  }
}

// out[i] = sum_c partial[c][i]; one CTA per 32 outputs, 8 warps stride the chunks, fixed-order combine.
__global__ void __launch_bounds__(256) chunk_sum_kernel(const float* __restrict__ partial, int num_chunks,
                                                        int64_t width, float* __restrict__ out) {
  __shared__ float part[8][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + lane;
  float t0 = 0.f, t1 = 0.f;
  if (i < width) {
    int c = wid;
    for (; c + 8 < num_chunks; c += 16) {
      t0 += partial[(int64_t)c * width + i];
      t1 += partial[(int64_t)(c + 8) * width + i];
    }
    if (c < num_chunks) t0 += partial[(int64_t)c * width + i];
  }
  part[wid][lane] = t0 + t1;
  __syncthreads();
  if (wid == 0 && i < width) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += part[w][lane];
    out[i] = t;
  }
}

__global__ void __launch_bounds__(256) skinny_dx_kernel(const float* __restrict__ dy, int64_t lddy,
                                                        const float* __restrict__ w, int num_rows, int K, int M,
                                                        float* __restrict__ dx, int64_t lddx) {
  extern __shared__ float ws[];  // [M][K] as given
  for (int i = threadIdx.x; i < K * M; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < num_rows; n += warps) {
    float acc = 0.f;
    for (int m0 = 0; m0 < M; m0 += 32) {
      const float dv = (m0 + lane < M) ? __ldg(dy + (int64_t)n * lddy + m0 + lane) : 0.f;
      const int mm = min(32, M - m0);
      for (int j = 0; j < mm; ++j) {
        const float d = __shfl_sync(kFullMask, dv, j);
        if (lane < K) acc = fmaf(d, ws[(m0 + j) * K + lane], acc);
      }
    }
    if (lane < K) dx[(int64_t)n * lddx + lane] = acc;
  }
}

}  // namespace ghscn

using namespace ghscn;

extern "C" {

int ghscn_skinny_linear_fwd(const float* x, int64_t ldx, const float* w, const float* bias, int64_t num_rows,
                            int64_t in_feat, int64_t out_feat, float* y, int64_t ldy, ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_rows >= 0 && in_feat > 0 && out_feat > 0 && num_rows < ((int64_t)1 << 31));
  if (in_feat > kSkinnyMaxK || in_feat * out_feat * 4 > 96 * 1024) return GHSCN_E_UNSUPPORTED;
  if (num_rows == 0) return GHSCN_OK;
  GHSCN_REQUIRE(x && w && y && ldx >= in_feat && ldy >= out_feat);
  const int64_t want = ceil_div<int64_t>(num_rows, 8 * 4);  // ~4 rows per warp
  const int64_t cap = (int64_t)kNumSMs * 8;
  const unsigned grid = (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
  if (in_feat <= 12 && out_feat <= 320) {
    const int j = (int)ceil_div<int64_t>(out_feat, 32);
    // the weights live in registers: keep the grid at the resident CTA count so that they are loaded once per SM
    const int64_t resident = (int64_t)kNumSMs * (j <= 2 ? 4 : (j <= 4 ? 2 : 1));
    const int64_t want_r = ceil_div<int64_t>(num_rows, 8 * 8);
    const unsigned grid_r = (unsigned)(want_r < resident ? (want_r > 0 ? want_r : 1) : resident);
    const size_t shm_r = (size_t)in_feat * out_feat * 4;
#define GHSCN_SKINNY_REG(J)                                                                              \
  skinny_fwd_reg_kernel<J><<<grid_r, 256, shm_r, as_stream(stream)>>>(x, ldx, w, bias, (int)num_rows,    \
                                                                      (int)in_feat, (int)out_feat, y, ldy)
    if (j <= 1) GHSCN_SKINNY_REG(1);
    else if (j <= 2) GHSCN_SKINNY_REG(2);
    else if (j <= 4) GHSCN_SKINNY_REG(4);
    else GHSCN_SKINNY_REG(10);
#undef GHSCN_SKINNY_REG
    GHSCN_LAUNCH_CHECK();
    return GHSCN_OK;
  }
  const size_t shm = (size_t)in_feat * out_feat * 4;
  if (shm > 48 * 1024)
    cudaFuncSetAttribute(skinny_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  skinny_fwd_kernel<<<grid, 256, shm, as_stream(stream)>>>(
      x, ldx, w, bias, (int)num_rows, (int)in_feat, (int)out_feat, y, ldy);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

size_t ghscn_skinny_dw_workspace_bytes(int64_t num_rows, int64_t in_feat, int64_t out_feat) {
  if (num_rows < 0 || in_feat < 0 || out_feat < 0) return 0;
  int64_t chunks = ceil_div<int64_t>(num_rows > 0 ? num_rows : 1, dw_rows(in_feat <= 16 ? 16 : 32));
  if (chunks < kNumSMs) chunks = kNumSMs;                    // the register-resident variant uses up to one chunk per SM
  return (size_t)chunks * in_feat * out_feat * 4 + 256;
}

int ghscn_skinny_linear_dw(const float* dy, int64_t lddy, const float* x, int64_t ldx, int64_t num_rows,
                           int64_t in_feat, int64_t out_feat, float* dw, void* workspace, size_t workspace_bytes,
                           ghscn_stream_t stream_) {
  GHSCN_REQUIRE(num_rows >= 0 && in_feat > 0 && out_feat > 0 && num_rows < ((int64_t)1 << 31));
  if (in_feat > kSkinnyMaxK) return GHSCN_E_UNSUPPORTED;
  GHSCN_REQUIRE(dw && (num_rows == 0 || (dy && x && lddy >= out_feat && ldx >= in_feat)));
  if (workspace == nullptr || workspace_bytes < ghscn_skinny_dw_workspace_bytes(num_rows, in_feat, out_feat))
    return GHSCN_E_WORKSPACE;
  cudaStream_t stream = as_stream(stream_);
  float* partial = static_cast<float*>(workspace);
  const int K = (int)in_feat, M = (int)out_feat;
  int chunks;
  if (K <= 16 && num_rows > 0 && (M <= 128 || (K <= 12 && M <= 320))) {
    int rows_per_cta = (int)ceil_div<int64_t>(num_rows, kNumSMs);
    rows_per_cta = (rows_per_cta + 15) / 16 * 16;
    if (rows_per_cta < 64) rows_per_cta = 64;
    chunks = (int)ceil_div<int64_t>(num_rows, rows_per_cta);
    const int j = ceil_div(M, 32);
#define GHSCN_SKINNY_DW(J, KR, NY)                                                                       \
  skinny_dw_rows_kernel<J, KR><<<dim3((unsigned)chunks, NY), 256, 0, stream>>>(dy, lddy, x, ldx, (int)num_rows, K, M, \
                                                                             rows_per_cta, partial)
    if (j <= 1) GHSCN_SKINNY_DW(1, 16, 1);
    else if (j <= 2) GHSCN_SKINNY_DW(2, 16, 1);
    else if (j <= 4) GHSCN_SKINNY_DW(4, 16, 1);
    else GHSCN_SKINNY_DW(5, 12, (unsigned)ceil_div(j, 5));     // 160 columns per CTA: ~100 registers, 2 CTAs per SM
#undef GHSCN_SKINNY_DW
  } else {
    chunks = (int)ceil_div<int64_t>(num_rows, dw_rows(in_feat <= 16 ? 16 : 32));
    if (chunks > 65535) return GHSCN_E_UNSUPPORTED;
    if (chunks > 0) {
      dim3 grid((unsigned)ceil_div(M, 32), (unsigned)chunks);
      if (K <= 16)
        skinny_dw_partial_kernel<16><<<grid, 256, 0, stream>>>(dy, lddy, x, ldx, (int)num_rows, K, M, partial);
      else
        skinny_dw_partial_kernel<32><<<grid, 256, 0, stream>>>(dy, lddy, x, ldx, (int)num_rows, K, M, partial);
    }
  }
  const int64_t width = (int64_t)M * K;
  chunk_sum_kernel<<<(unsigned)ceil_div<int64_t>(width, 32), 256, 0, stream>>>(partial, chunks, width, dw);
  GHSCN_LAUNCH_CHECK_N(chunks > 0 ? 2 : 1);
  return GHSCN_OK;
}

int ghscn_skinny_linear_dx(const float* dy, int64_t lddy, const float* w, int64_t num_rows, int64_t in_feat,
                           int64_t out_feat, float* dx, int64_t lddx, ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_rows >= 0 && in_feat > 0 && out_feat > 0 && num_rows < ((int64_t)1 << 31));
  if (in_feat > kSkinnyMaxK || in_feat * out_feat * 4 > 96 * 1024) return GHSCN_E_UNSUPPORTED;
  if (num_rows == 0) return GHSCN_OK;
  GHSCN_REQUIRE(dy && w && dx && lddy >= out_feat && lddx >= in_feat);
  const size_t shm = (size_t)in_feat * out_feat * 4;
  if (shm > 48 * 1024)
    cudaFuncSetAttribute(skinny_dx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const int64_t want = ceil_div<int64_t>(num_rows, 8 * 4);
  const int64_t cap = (int64_t)kNumSMs * 8;
  skinny_dx_kernel<<<(unsigned)(want < cap ? (want > 0 ? want : 1) : cap), 256, shm, as_stream(stream)>>>(
      dy, lddy, w, (int)num_rows, (int)in_feat, (int)out_feat, dx, lddx);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

}  // extern "C"

// ---- AdamW on one flat fp32 parameter buffer (train/train.py:94 `optimizer.step()`, OPTIM_DICT["adamW"]) -----------
// torch.optim.AdamW's (non-capturable, single-tensor) update, elementwise, with the step count kept on the device so
// the launch is CUDA-graph capturable:  p *= 1 - lr*wd;  m = lerp(m, g, 1-b1);  v = b2*v + (1-b2)*g*g;
//              p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
namespace ghscn {
// state[0] = steps taken so far, state[1] = lr / (1 - b1^t), state[2] = sqrt(1 - b2^t) for the step being taken.
// One thread advances the counter and derives the two bias-correction scalars in double precision, exactly as
// torch.optim.AdamW's Python code does on the host; the elementwise kernel then only reads them.
__global__ void adamw_prepare_kernel(float* __restrict__ state, double lr, double b1, double b2) {
  const double t = (double)state[0] + 1.0;
  state[0] = (float)t;
  state[1] = (float)(lr / (1.0 - pow(b1, t)));
  state[2] = (float)sqrt(1.0 - pow(b2, t));
}
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                    float decay, float omb1, float b2, float omb2, float eps,
                                                    const float* __restrict__ state,
                                                    const float* __restrict__ grad_scale) {
  const float step_size = state[1], bc2_sqrt = state[2];
  const float gs = grad_scale ? grad_scale[0] : 1.0f;   // clip_grad_norm coefficient (== g.mul_(coef) beforehand)
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = grad_scale ? g[i] * gs : g[i];
    float pi = p[i] * decay;                                    // param.mul_(1 - lr * wd)
    const float mi = m[i] + (gi - m[i]) * omb1;                 // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = v[i] * b2 + omb2 * (gi * gi);              // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= step_size * (mi / denom);                             // param.addcdiv_(exp_avg, denom, value=-step_size)
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}

// ---- clip_grad_norm (train/train.py:92-93) over the flat gradient buffer ------------------------------------------
// total = ||g||_2 as a fixed-order two-stage sum of squares (deterministic); coef = min(1, max_norm / (total + 1e-6)),
// torch.nn.utils.clip_grad_norm_'s formula.  out[0] = total, out[1] = coef; the AdamW kernel applies coef.
constexpr int kClipCtas = 148;
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, int64_t n,
                                                            float* __restrict__ partial) {
  __shared__ float red[32];
  float acc = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) acc = fmaf(g[i], g[i], acc);
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
__global__ void __launch_bounds__(256) clip_coef_kernel(const float* __restrict__ partial, int count, float max_norm,
                                                        float* __restrict__ out) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < count; i += blockDim.x) acc += partial[i];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) {
    const float total = sqrtf(acc);
    const float coef = max_norm / (total + 1e-6f);
    out[0] = total;
    out[1] = coef < 1.0f ? coef : 1.0f;
  }
}
}  // namespace ghscn

extern "C" size_t ghscn_grad_clip_workspace_bytes(int64_t n) {
  (void)n;
  return (size_t)ghscn::kClipCtas * sizeof(float);
}

extern "C" int ghscn_grad_clip_scale(const float* grad, int64_t n, float max_norm, void* workspace,
                                     size_t workspace_bytes, float* out, ghscn_stream_t stream_) {
  GHSCN_REQUIRE(n >= 0 && out && max_norm > 0.f && (n == 0 || grad));
  if (workspace == nullptr || workspace_bytes < ghscn_grad_clip_workspace_bytes(n)) return GHSCN_E_WORKSPACE;
  cudaStream_t stream = ghscn::as_stream(stream_);
  float* partial = static_cast<float*>(workspace);
  ghscn::sumsq_partial_kernel<<<ghscn::kClipCtas, 256, 0, stream>>>(grad, n, partial);
  ghscn::clip_coef_kernel<<<1, 256, 0, stream>>>(partial, ghscn::kClipCtas, max_norm, out);
  GHSCN_LAUNCH_CHECK_N(2);
  return GHSCN_OK;
}

extern "C" int ghscn_adamw_step_scaled(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                                       double lr, double beta1, double beta2, double eps, double weight_decay,
                                       float* state, const float* grad_scale, ghscn_stream_t stream_) {
  GHSCN_REQUIRE(n >= 0 && state && (n == 0 || (param && grad && exp_avg && exp_avg_sq)));
  cudaStream_t stream = ghscn::as_stream(stream_);
  ghscn::adamw_prepare_kernel<<<1, 1, 0, stream>>>(state, lr, beta1, beta2);
  if (n > 0) {
    const int64_t blocks = ghscn::ceil_div<int64_t>(n, 256 * 4), cap = (int64_t)ghscn::kNumSMs * 8;
    // the hyper-parameter combinations are formed in double and rounded once, like the Python scalars torch passes
    ghscn::adamw_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, stream>>>(
        param, grad, exp_avg, exp_avg_sq, n, (float)(1.0 - lr * weight_decay), (float)(1.0 - beta1), (float)beta2,
        (float)(1.0 - beta2), (float)eps, state, grad_scale);
  }
  GHSCN_LAUNCH_CHECK_N(n > 0 ? 2 : 1);
  return GHSCN_OK;
}

extern "C" int ghscn_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                                double lr, double beta1, double beta2, double eps, double weight_decay, float* state,
                                ghscn_stream_t stream_) {
  return ghscn_adamw_step_scaled(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, state,
                                 nullptr, stream_);
}

// ---- gradients of many parameters -> one flat buffer ---------------------------------------------------------------
// `torch.autograd.grad` hands the step its gradients as separate tensors; the flat-buffer optimizer / all-reduce wants
// them back to back.  One launch copies up to 64 tensors: their device pointers and sizes travel as kernel arguments
// (no device-side table to upload, so the call is CUDA-graph capturable and needs no host synchronisation).
namespace ghscn {
constexpr int kGatherMax = 64;
struct GatherArgs {
  const float* src[kGatherMax];
  long long off[kGatherMax + 1];       // element offsets inside the flat destination (prefix sums)
  int count;
};
__global__ void __launch_bounds__(256) gather_flat_kernel(GatherArgs a, float* __restrict__ dst) {
  const long long total = a.off[a.count];
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4; i < total;
       i += (long long)gridDim.x * blockDim.x * 4) {
    int lo = 0, hi = a.count - 1;                     // tensor holding element i
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (a.off[mid] <= i) lo = mid; else hi = mid - 1;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long j = i + u;
      if (j >= total) break;
      while (j >= a.off[lo + 1]) ++lo;
      dst[j] = a.src[lo][j - a.off[lo]];
    }
  }
}
}  // namespace ghscn

extern "C" int ghscn_gather_flat(const float* const* srcs_host, const int64_t* numels_host, int32_t count, float* dst,
                                 ghscn_stream_t stream_) {
  GHSCN_REQUIRE(count >= 0 && (count == 0 || (srcs_host && numels_host && dst)));
  cudaStream_t stream = ghscn::as_stream(stream_);
  int launches = 0;
  int64_t base = 0;
  for (int first = 0; first < count; first += ghscn::kGatherMax) {
    ghscn::GatherArgs a;
    a.count = count - first < ghscn::kGatherMax ? count - first : ghscn::kGatherMax;
    a.off[0] = 0;
    for (int i = 0; i < a.count; ++i) {
      GHSCN_REQUIRE(numels_host[first + i] >= 0 && (numels_host[first + i] == 0 || srcs_host[first + i]));
      a.src[i] = srcs_host[first + i];
      a.off[i + 1] = a.off[i] + numels_host[first + i];
    }
    const long long total = a.off[a.count];
    if (total > 0) {
      const long long blocks = (total + 1023) / 1024;
      ghscn::gather_flat_kernel<<<(unsigned)(blocks < 4096 ? blocks : 4096), 256, 0, stream>>>(a, dst + base);
      ++launches;
    }
    base += total;
  }
  if (launches) GHSCN_LAUNCH_CHECK_N(launches);
  return GHSCN_OK;
}
