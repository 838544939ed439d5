// Small dense layers of the readout head and the virtual-node branch, plus the fused task loss.
//
// Reference call sites: model/hscn.py:99-100,112 (`lin_1`, activation, `lin_2` on the [B, H] graph embeddings),
// loss.py:6-19 (`binary_cross_entropy_with_logits` / `l1_loss`, mean reduction, sigmoid score), and the GCN/GAT
// projections of the ~10 virtual nodes per graph (model/hscn.py:85-93; M ~ 1.3 k rows).
//
// These GEMMs have 10^2..10^3 rows: far too small for a 128-row tcgen05 tile per SM to pay off (one tile costs
// ~23 us of pipeline latency whatever its work) and latency-bound as library calls (a cuBLAS SIMT sgemm + a bias
// epilogue + an activation kernel per layer; ~30 tiny kernels ~ 130 us for the head's forward and backward).  Plain
// fp32 FMA tiles, one launch per product, activation / bias / activation-derivative folded into the loads and the
// epilogue; every reduction runs in a fixed order (deterministic).  fp32 throughout: same accuracy class as the
// cuBLAS fp32 path they replace.
#include <math.h>

#include "common.cuh"

namespace ghscn {

enum { kActNone = 0, kActElu = 1, kActRelu = 2, kActTanh = 3 };

__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case kActElu: return v > 0.f ? v : expm1f(v);
    case kActRelu: return fmaxf(v, 0.f);
    case kActTanh: return tanhf(v);
    default: return v;
  }
}
// derivative of the activation expressed through its OUTPUT y (what the forward saved)
__device__ __forceinline__ float act_grad_from_output(float y, int act) {
  switch (act) {
    case kActElu: return y > 0.f ? 1.f : y + 1.f;      // alpha = 1: d/dv expm1(v) = exp(v) = y + 1
    case kActRelu: return y > 0.f ? 1.f : 0.f;
    case kActTanh: return 1.f - y * y;
    default: return 1.f;
  }
}

constexpr int kBK = 16;

// C[m, n] = epi( sum_k A'(m, k) * B(k, n) ),  A' = A (.) act'(Yref) when Yref != nullptr.
//   B_T = true : B(k, n) = W[n * ldw + k]   (y = x W^T: W is [N, K] row-major)
//   B_T = false: B(k, n) = W[k * ldw + n]   (dx = dy W:  W is [K, N] row-major)
// An optional second operand pair continues the reduction (k >= K reads A2 / W2 at k - K; B_T form only):
//   C = epi(A W^T + A2 W2^T + bias + bias2), the sum of two layers that share their destination rows.
// CTA tile BM x 64, 256 threads, thread tile (BM / 16) x 4, K chunks of 16 staged k-major in shared memory.
template <int BM, bool B_T>
__global__ void __launch_bounds__(256) small_gemm_kernel(const float* __restrict__ A, int64_t lda,
                                                         const float* __restrict__ Yref, int64_t ldy, int mask_act,
                                                         const float* __restrict__ W, int64_t ldw,
                                                         const float* __restrict__ bias, int act, int M, int N, int K,
                                                         float* __restrict__ C, int64_t ldc,
                                                         const float* __restrict__ A2 = nullptr, int64_t lda2 = 0,
                                                         const float* __restrict__ W2 = nullptr, int64_t ldw2 = 0,
                                                         int K2 = 0, const float* __restrict__ bias2 = nullptr) {
  constexpr int BN = 64, TM = BM / 16, TN = 4;
  __shared__ __align__(16) float As[kBK][BM + 4];
  __shared__ __align__(16) float Bs[kBK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tx = tid & 15, ty = tid >> 4;            // tx: column group (4 columns), ty: row group (TM rows)
  float acc[TM][TN];
#pragma unroll
  for (int a = 0; a < TM; ++a)
#pragma unroll
    for (int b = 0; b < TN; ++b) acc[a][b] = 0.f;

  const int KT = K + K2;
  for (int k0 = 0; k0 < KT; k0 += kBK) {
    // ---- A tile: BM rows x 16 k, k contiguous in memory -> As[k][m]
    for (int e = tid; e < BM * 4; e += 256) {
      const int r = e >> 2, kq = (e & 3) * 4;
      const int m = m0 + r;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + kq + u;
        float v = 0.f;
        if (m < M && k < K) {
          v = A[(int64_t)m * lda + k];
          if (Yref) v *= act_grad_from_output(Yref[(int64_t)m * ldy + k], mask_act);
        } else if (m < M && k < KT) {
          v = A2[(int64_t)m * lda2 + (k - K)];
        }
        As[kq + u][r] = v;
      }
    }
    // ---- B tile -> Bs[k][n]
    if (B_T) {
      for (int e = tid; e < BN * 4; e += 256) {
        const int c = e >> 2, kq = (e & 3) * 4;
        const int n = n0 + c;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int k = k0 + kq + u;
          float v = 0.f;
          if (n < N && k < K) v = W[(int64_t)n * ldw + k];
          else if (n < N && k < KT) v = W2[(int64_t)n * ldw2 + (k - K)];
          Bs[kq + u][c] = v;
        }
      }
    } else {
      for (int e = tid; e < kBK * 16; e += 256) {
        const int kk = e >> 4, cq = (e & 15) * 4;
        const int k = k0 + kk;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int n = n0 + cq + u;
          Bs[kk][cq + u] = (n < N && k < K) ? W[(int64_t)k * ldw + n] : 0.f;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      float av[TM], bv[TN];
#pragma unroll
      for (int a = 0; a < TM; ++a) av[a] = As[kk][ty * TM + a];
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      bv[0] = b4.x; bv[1] = b4.y; bv[2] = b4.z; bv[3] = b4.w;
#pragma unroll
      for (int a = 0; a < TM; ++a)
#pragma unroll
        for (int b = 0; b < TN; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < TM; ++a) {
    const int m = m0 + ty * TM + a;
    if (m >= M) continue;
#pragma unroll
    for (int b = 0; b < TN; ++b) {
      const int n = n0 + tx * 4 + b;
      if (n >= N) continue;
      float v = acc[a][b];
      if (bias) v += bias[n];
      if (bias2) v += bias2[n];
      C[(int64_t)m * ldc + n] = act_apply(v, act);
    }
  }
}

// dW[n, k] = sum_g dY'(g, n) X(g, k),  db[n] = sum_g dY'(g, n),  dY' = dY (.) act'(Yref) when Yref != nullptr.
// CTA tile 64 (n) x 64 (k); rows g are reduced in chunks of 16; gridDim.z splits the rows (partials are written to
// `part` [z][N*K + N] and added in split order by small_reduce_kernel; z == 1 writes the results directly).
__global__ void __launch_bounds__(256) small_gemm_tn_kernel(const float* __restrict__ dY, int64_t lddy,
                                                            const float* __restrict__ Yref, int64_t ldy, int mask_act,
                                                            const float* __restrict__ X, int64_t ldx, int G, int N,
                                                            int K, int rows_per_split, float* __restrict__ dW,
                                                            float* __restrict__ db, float* __restrict__ part) {
  constexpr int BN = 64, BKO = 64;
  __shared__ __align__(16) float As[kBK][BN + 4];    // [g][n]
  __shared__ __align__(16) float Bs[kBK][BKO + 4];   // [g][k]
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * BN, c0 = blockIdx.y * BKO;
  const int g_beg = blockIdx.z * rows_per_split, g_end = min(G, g_beg + rows_per_split);
  const int tx = tid & 15, ty = tid >> 4;            // ty: 4 n-rows of the tile, tx: 4 k-columns
  float acc[4][4], bacc[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    bacc[a] = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  }
  for (int g0 = g_beg; g0 < g_end; g0 += kBK) {
    for (int e = tid; e < kBK * 16; e += 256) {
      const int gg = e >> 4, q = (e & 15) * 4;
      const int g = g0 + gg;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int n = n0 + q + u, k = c0 + q + u;
        float a = 0.f, x = 0.f;
        if (g < g_end) {
          if (n < N) {
            a = dY[(int64_t)g * lddy + n];
            if (Yref) a *= act_grad_from_output(Yref[(int64_t)g * ldy + n], mask_act);
          }
          if (k < K) x = X[(int64_t)g * ldx + k];
        }
        As[gg][q + u] = a;
        Bs[gg][q + u] = x;
      }
    }
    __syncthreads();
#pragma unroll
    for (int gg = 0; gg < kBK; ++gg) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[gg][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[gg][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        bacc[a] += av[a];
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
      }
    }
    __syncthreads();
  }
  const bool direct = gridDim.z == 1;
  float* wout = direct ? dW : part + (int64_t)blockIdx.z * ((int64_t)N * K + N);
  float* bout = direct ? db : wout + (int64_t)N * K;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int n = n0 + ty * 4 + a;
    if (n >= N) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int k = c0 + tx * 4 + b;
      if (k < K) wout[(int64_t)n * K + k] = acc[a][b];
    }
    if (bout && blockIdx.y == 0 && tx == 0) bout[n] = bacc[a];
  }
}

__global__ void __launch_bounds__(256) small_reduce_kernel(const float* __restrict__ part, int splits, int64_t n_w,
                                                           int64_t n_b, float* __restrict__ dW,
                                                           float* __restrict__ db) {
  const int64_t stride = n_w + n_b;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < stride; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[z * stride + i];
    if (i < n_w) dW[i] = s;
    else if (db) db[i - n_w] = s;
  }
}

// Task loss over the first `rows` rows of pred / target [*, C] (mean reduction), its gradient and the sigmoid score.
//   mode 0: binary_cross_entropy_with_logits   l = (1 - t) x - log_sigmoid(x),  dl/dx = sigmoid(x) - t
//   mode 1: l1_loss                            l = |x - t|,                    dl/dx = sign(x - t)
// One CTA: fixed-order block reduction (deterministic).  d_pred rows >= rows are zero (padding graphs).
__global__ void __launch_bounds__(1024) graph_loss_kernel(const float* __restrict__ pred, int64_t ldp,
                                                          const float* __restrict__ target, int64_t ldt, int rows,
                                                          int total_rows, int C, int mode, float* __restrict__ loss,
                                                          float* __restrict__ d_pred, float* __restrict__ score) {
  __shared__ float red[32];
  const int n = rows * C, n_all = total_rows * C;
  const float inv = 1.0f / (float)max(n, 1);
  float acc = 0.f;
  for (int e = threadIdx.x; e < n_all; e += blockDim.x) {
    const int r = e / C, c = e - r * C;
    const float x = pred[(int64_t)r * ldp + c];
    const float sg = 1.0f / (1.0f + expf(-x));
    if (score) score[e] = sg;
    if (r >= rows) {
      d_pred[e] = 0.f;
      continue;
    }
    const float t = target[(int64_t)r * ldt + c];
    if (mode == 0) {
      const float log_sig = fminf(x, 0.f) - log1pf(expf(-fabsf(x)));
      acc += (1.0f - t) * x - log_sig;
      d_pred[e] = (sg - t) * inv;
    } else {
      const float d = x - t;
      acc += fabsf(d);
      d_pred[e] = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * inv;
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) loss[0] = acc * inv;
}

}  // namespace ghscn

using namespace ghscn;

extern "C" {

int ghscn_small_linear_fwd(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, int32_t act,
                           int64_t num_rows, int64_t in_feat, int64_t out_feat, float* y, int64_t ldy,
                           ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_rows >= 0 && in_feat > 0 && out_feat > 0 && act >= 0 && act <= 3);
  GHSCN_REQUIRE(num_rows < ((int64_t)1 << 31) && in_feat < ((int64_t)1 << 31) && out_feat < ((int64_t)1 << 31));
  if (num_rows == 0) return GHSCN_OK;
  GHSCN_REQUIRE(x && w && y && ldx >= in_feat && ldw >= in_feat && ldy >= out_feat);
  const int M = (int)num_rows, N = (int)out_feat, K = (int)in_feat;
  const unsigned gy = (unsigned)ceil_div(N, 64);
  if ((int64_t)ceil_div(M, 64) * gy >= kNumSMs) {
    dim3 grid((unsigned)ceil_div(M, 64), gy);
    small_gemm_kernel<64, true><<<grid, 256, 0, as_stream(stream)>>>(x, ldx, nullptr, 0, 0, w, ldw, bias, act, M, N, K,
                                                                     y, ldy);
  } else {
    dim3 grid((unsigned)ceil_div(M, 32), gy);
    small_gemm_kernel<32, true><<<grid, 256, 0, as_stream(stream)>>>(x, ldx, nullptr, 0, 0, w, ldw, bias, act, M, N, K,
                                                                     y, ldy);
  }
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

int ghscn_small_linear2_fwd(const float* x1, int64_t ldx1, const float* w1, int64_t ldw1, const float* bias1,
                            int64_t in_feat1, const float* x2, int64_t ldx2, const float* w2, int64_t ldw2,
                            const float* bias2, int64_t in_feat2, int32_t act, int64_t num_rows, int64_t out_feat,
                            float* y, int64_t ldy, ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_rows >= 0 && in_feat1 > 0 && in_feat2 > 0 && out_feat > 0 && act >= 0 && act <= 3);
  GHSCN_REQUIRE(num_rows < ((int64_t)1 << 31) && in_feat1 + in_feat2 < ((int64_t)1 << 31) && out_feat < ((int64_t)1 << 31));
  if (num_rows == 0) return GHSCN_OK;
  GHSCN_REQUIRE(x1 && w1 && x2 && w2 && y && ldx1 >= in_feat1 && ldw1 >= in_feat1 && ldx2 >= in_feat2 &&
                ldw2 >= in_feat2 && ldy >= out_feat);
  const int M = (int)num_rows, N = (int)out_feat;
  const unsigned gy = (unsigned)ceil_div(N, 64);
  if ((int64_t)ceil_div(M, 64) * gy >= kNumSMs) {
    dim3 grid((unsigned)ceil_div(M, 64), gy);
    small_gemm_kernel<64, true><<<grid, 256, 0, as_stream(stream)>>>(x1, ldx1, nullptr, 0, 0, w1, ldw1, bias1, act, M, N,
                                                                     (int)in_feat1, y, ldy, x2, ldx2, w2, ldw2,
                                                                     (int)in_feat2, bias2);
  } else {
    dim3 grid((unsigned)ceil_div(M, 32), gy);
    small_gemm_kernel<32, true><<<grid, 256, 0, as_stream(stream)>>>(x1, ldx1, nullptr, 0, 0, w1, ldw1, bias1, act, M, N,
                                                                     (int)in_feat1, y, ldy, x2, ldx2, w2, ldw2,
                                                                     (int)in_feat2, bias2);
  }
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

int ghscn_small_linear_dx(const float* dy, int64_t lddy, const float* y_ref, int64_t ldy, int32_t act, const float* w,
                          int64_t ldw, int64_t num_rows, int64_t in_feat, int64_t out_feat, float* dx, int64_t lddx,
                          ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_rows >= 0 && in_feat > 0 && out_feat > 0 && act >= 0 && act <= 3);
  GHSCN_REQUIRE(num_rows < ((int64_t)1 << 31) && in_feat < ((int64_t)1 << 31) && out_feat < ((int64_t)1 << 31));
  if (num_rows == 0) return GHSCN_OK;
  GHSCN_REQUIRE(dy && w && dx && lddy >= out_feat && ldw >= in_feat && lddx >= in_feat);
  GHSCN_REQUIRE(act == 0 || (y_ref && ldy >= out_feat));
  const int M = (int)num_rows, N = (int)in_feat, K = (int)out_feat;       // dx [M, in] = dy' [M, out] . W [out, in]
  const float* yr = act == 0 ? nullptr : y_ref;
  const unsigned gy = (unsigned)ceil_div(N, 64);
  if ((int64_t)ceil_div(M, 64) * gy >= kNumSMs) {
    dim3 grid((unsigned)ceil_div(M, 64), gy);
    small_gemm_kernel<64, false><<<grid, 256, 0, as_stream(stream)>>>(dy, lddy, yr, ldy, act, w, ldw, nullptr, 0, M, N,
                                                                      K, dx, lddx);
  } else {
    dim3 grid((unsigned)ceil_div(M, 32), gy);
    small_gemm_kernel<32, false><<<grid, 256, 0, as_stream(stream)>>>(dy, lddy, yr, ldy, act, w, ldw, nullptr, 0, M, N,
                                                                      K, dx, lddx);
  }
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

static int small_dw_splits(int64_t num_rows) {
  const int64_t z = ceil_div<int64_t>(num_rows, 256);
  return (int)(z < 1 ? 1 : (z > 8 ? 8 : z));
}

size_t ghscn_small_linear_dw_workspace_bytes(int64_t num_rows, int64_t in_feat, int64_t out_feat) {
  if (num_rows < 0 || in_feat <= 0 || out_feat <= 0) return 0;
  const int z = small_dw_splits(num_rows);
  return z == 1 ? 0 : (size_t)z * (size_t)(out_feat * in_feat + out_feat) * 4;
}

int ghscn_small_linear_dw(const float* dy, int64_t lddy, const float* y_ref, int64_t ldy, int32_t act, const float* x,
                          int64_t ldx, int64_t num_rows, int64_t in_feat, int64_t out_feat, float* dw, float* db,
                          void* workspace, size_t workspace_bytes, ghscn_stream_t stream_) {
  GHSCN_REQUIRE(num_rows >= 0 && in_feat > 0 && out_feat > 0 && act >= 0 && act <= 3 && dw);
  GHSCN_REQUIRE(num_rows < ((int64_t)1 << 31) && in_feat < ((int64_t)1 << 31) && out_feat < ((int64_t)1 << 31));
  GHSCN_REQUIRE(num_rows == 0 || (dy && x && lddy >= out_feat && ldx >= in_feat));
  GHSCN_REQUIRE(act == 0 || num_rows == 0 || (y_ref && ldy >= out_feat));
  cudaStream_t stream = as_stream(stream_);
  const int G = (int)num_rows, N = (int)out_feat, K = (int)in_feat;
  const int z = small_dw_splits(num_rows);
  if (z > 1 && (workspace == nullptr || workspace_bytes < ghscn_small_linear_dw_workspace_bytes(num_rows, in_feat, out_feat)))
    return GHSCN_E_WORKSPACE;
  const int rows_per_split = ceil_div(ceil_div(G, z), kBK) * kBK;
  dim3 grid((unsigned)ceil_div(N, 64), (unsigned)ceil_div(K, 64), (unsigned)z);
  small_gemm_tn_kernel<<<grid, 256, 0, stream>>>(dy, lddy, act == 0 ? nullptr : y_ref, ldy, act, x, ldx, G, N, K,
                                                 rows_per_split > 0 ? rows_per_split : kBK, dw, db,
                                                 static_cast<float*>(workspace));
  if (z > 1) {
    const int64_t total = (int64_t)N * K + N;
    small_reduce_kernel<<<(unsigned)ceil_div<int64_t>(total, 256), 256, 0, stream>>>(
        static_cast<const float*>(workspace), z, (int64_t)N * K, N, dw, db);
  }
  GHSCN_LAUNCH_CHECK_N(z > 1 ? 2 : 1);
  return GHSCN_OK;
}

int ghscn_graph_loss(const float* pred, int64_t ldp, const float* target, int64_t ldt, int64_t rows,
                     int64_t total_rows, int64_t num_targets, int32_t mode, float* loss, float* d_pred, float* score,
                     ghscn_stream_t stream) {
  GHSCN_REQUIRE(rows >= 0 && total_rows >= rows && num_targets > 0 && (mode == 0 || mode == 1) && loss && d_pred);
  GHSCN_REQUIRE(total_rows * num_targets < ((int64_t)1 << 30));
  GHSCN_REQUIRE(total_rows == 0 || (pred && ldp >= num_targets));
  GHSCN_REQUIRE(rows == 0 || (target && ldt >= num_targets));
  graph_loss_kernel<<<1, 1024, 0, as_stream(stream)>>>(pred, ldp, target, ldt, (int)rows, (int)total_rows,
                                                      (int)num_targets, mode, loss, d_pred, score);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

}  // extern "C"
