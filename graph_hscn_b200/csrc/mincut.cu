// K6: fused MinCUT pool, one CTA per graph, forward and analytic backward.
// Replaces to_dense_adj + dense_mincut_pool (model/hscn.py:61-63).
//
// The [n,n] dense adjacency is never materialised: the graph's CSR slice is consumed directly
// ((A S)[r] = sum_{e: row_e = r} val_e S[col_e]).  Per graph the CTA stages S = softmax(logits)
// and A S in shared memory (falls back to an HBM workspace when n*K does not fit), and produces
// S^T X, S^T A S, S^T S, both traces and both losses without any other round trip through HBM.
// HBM-bound at the reference's sizes (K <= 32): algorithmic bytes per graph
//   4nK (logits) + 4nH (X, only if `out` is requested) + 4(n+1) + 4 nnz  ->  4nK (S) + 4KH + 8K^2 + 32.
#include <math.h>

#include <cstdlib>

#include "common.cuh"

namespace ghscn {

// few graphs: one big CTA per graph keeps an SM busy; many graphs: several small CTAs per SM
static inline int mincut_threads(int64_t num_graphs) {
  static const int forced = [] {                       // tuning experiments: GHSCN_MINCUT_THREADS=256|512|1024
    const char* e = getenv("GHSCN_MINCUT_THREADS");
    const int v = e ? atoi(e) : 0;
    return (v == 128 || v == 256 || v == 512 || v == 1024) ? v : 0;
  }();
  if (forced) return forced;
  return num_graphs >= 4 * kNumSMs ? 256 : (num_graphs >= 2 * kNumSMs ? 512 : 1024);
}
constexpr int kMaxClusters = 128;
constexpr int kStatsStride = 8;  // per graph: num, den, ||SS||_F, ||R||_F (= ortho_g), mc_g, -, -, -
constexpr size_t kSmemBudget = 200 * 1024;

__device__ __forceinline__ float adj_value(const float* __restrict__ adj_val, int s) {
  return adj_val ? adj_val[s] : 1.0f;
}

// rows of `buf` (n x K) <- (A S) using the CSR slice whose rows are the graph's nodes.
__device__ __forceinline__ void csr_times_s(const int* __restrict__ rowptr, const int* __restrict__ col,
                                            const float* __restrict__ adj_val, const float* __restrict__ S,
                                            int base, int n, int K, float* __restrict__ buf) {
  for (int e = threadIdx.x; e < n * K; e += blockDim.x) {
    const int i = e / K, k = e - i * K;
    float acc = 0.f;
    const int beg = rowptr[base + i], end = rowptr[base + i + 1];
    for (int s = beg; s < end; ++s) {
      const int c = col[s] - base;
      if (c >= 0 && c < n) acc += adj_value(adj_val, s) * S[c * K + k];
    }
    buf[e] = acc;
  }
}

// ---- register-tiled contractions over the graph's nodes -----------------------------------------------------------
// C[m, c] (+)= sum_i A[i*lda + m] * B[i*ldb + c],  m < M, c < Nc, i < n.   A (the S tile) usually lives in shared
// memory, B is S / AS (shared) or the graph's X rows (global, coalesced along c).  Each thread owns a TM x TN tile of
// C: per node it reads TM + TN values for TM*TN FMAs (vs 2 reads per FMA in a scalar loop); consecutive threads
// take consecutive column tiles of the same row tile, so A reads are warp broadcasts and B reads are contiguous.
template <int TM, int TN>
__device__ __forceinline__ void atb_tiled(const float* __restrict__ A, int lda, int M, const float* __restrict__ B,
                                          int64_t ldb, int Nc, int n, float* __restrict__ C, int64_t ldc) {
  const int mt = ceil_div(M, TM), ct = ceil_div(Nc, TN);
  const int tiles = mt * ct;
  // few tiles (small K): `split` adjacent lanes share one tile, each taking every split-th node, then a shuffle
  // reduction -- keeps the whole CTA busy instead of a handful of threads walking all n nodes alone
  int split = 1;
  while (split < 32 && tiles * split * 2 <= (int)blockDim.x) split *= 2;
  const bool vec_b = (TN == 4) && (ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
  const int work = tiles * split;
  for (int t0 = 0; t0 < work; t0 += blockDim.x) {
    const int t = t0 + threadIdx.x;
    const bool active = t < work;
    const int tile = active ? t / split : 0, sub = t % split;
    const int m0 = (tile / ct) * TM, c0 = (tile % ct) * TN;
    float acc[TM][TN];
#pragma unroll
    for (int a = 0; a < TM; ++a)
#pragma unroll
      for (int b = 0; b < TN; ++b) acc[a][b] = 0.f;
    const bool full = (m0 + TM <= M) && (c0 + TN <= Nc);
    if (active) {
      if (full && vec_b) {
        int i = sub;
        for (; i + 3 * split < n; i += 4 * split) {   // 4 nodes in flight per lane
          float4 q[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) q[u] = *reinterpret_cast<const float4*>(B + (int64_t)(i + u * split) * ldb + c0);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float bv[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
            for (int a = 0; a < TM; ++a) {
              const float av = A[(i + u * split) * lda + m0 + a];
#pragma unroll
              for (int b = 0; b < TN; ++b) acc[a][b] = fmaf(av, bv[b % 4], acc[a][b]);
            }
          }
        }
        for (; i < n; i += split) {
          const float4 q = *reinterpret_cast<const float4*>(B + (int64_t)i * ldb + c0);
          const float bv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int a = 0; a < TM; ++a) {
            const float av = A[i * lda + m0 + a];
#pragma unroll
            for (int b = 0; b < TN; ++b) acc[a][b] = fmaf(av, bv[b % 4], acc[a][b]);
          }
        }
      } else {
        for (int i = sub; i < n; i += split) {
          float av[TM], bv[TN];
#pragma unroll
          for (int a = 0; a < TM; ++a) av[a] = (m0 + a < M) ? A[i * lda + m0 + a] : 0.f;
#pragma unroll
          for (int b = 0; b < TN; ++b) bv[b] = (c0 + b < Nc) ? B[(int64_t)i * ldb + c0 + b] : 0.f;
#pragma unroll
          for (int a = 0; a < TM; ++a)
#pragma unroll
            for (int b = 0; b < TN; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
        }
      }
    }
    for (int off = split >> 1; off > 0; off >>= 1) {   // warp-uniform: split is the same for every thread
#pragma unroll
      for (int a = 0; a < TM; ++a)
#pragma unroll
        for (int b = 0; b < TN; ++b) acc[a][b] += __shfl_xor_sync(kFullMask, acc[a][b], off);
    }
    if (active && sub == 0) {
#pragma unroll
      for (int a = 0; a < TM; ++a)
#pragma unroll
        for (int b = 0; b < TN; ++b)
          if (m0 + a < M && c0 + b < Nc) C[(int64_t)(m0 + a) * ldc + c0 + b] = acc[a][b];
    }
  }
}

// D[i, c] = sum_l P[i*ldp + l] * Q[l*ldq + c] (+ D if accumulate); the [n,K] x [K,K] products of the backward.
template <int TR, int TN>
__device__ __forceinline__ void ab_tiled(const float* __restrict__ P, int ldp, int n, const float* __restrict__ Q,
                                         int ldq, int Kred, int Nc, float scale, float* __restrict__ D, int ldd,
                                         bool accumulate) {
  const int rt = ceil_div(n, TR), ct = ceil_div(Nc, TN);
  for (int t = threadIdx.x; t < rt * ct; t += blockDim.x) {
    const int r0 = (t / ct) * TR, c0 = (t % ct) * TN;
    float acc[TR][TN];
#pragma unroll
    for (int a = 0; a < TR; ++a)
#pragma unroll
      for (int b = 0; b < TN; ++b) acc[a][b] = 0.f;
    for (int l = 0; l < Kred; ++l) {
      float pv[TR], qv[TN];
#pragma unroll
      for (int a = 0; a < TR; ++a) pv[a] = (r0 + a < n) ? P[(r0 + a) * ldp + l] : 0.f;
#pragma unroll
      for (int b = 0; b < TN; ++b) qv[b] = (c0 + b < Nc) ? Q[l * ldq + c0 + b] : 0.f;
#pragma unroll
      for (int a = 0; a < TR; ++a)
#pragma unroll
        for (int b = 0; b < TN; ++b) acc[a][b] = fmaf(pv[a], qv[b], acc[a][b]);
    }
#pragma unroll
    for (int a = 0; a < TR; ++a)
#pragma unroll
      for (int b = 0; b < TN; ++b)
        if (r0 + a < n && c0 + b < Nc) {
          float* d = D + (r0 + a) * ldd + c0 + b;
          *d = accumulate ? (*d + scale * acc[a][b]) : scale * acc[a][b];
        }
  }
}

// D[i, c] (+)= scale * sum_h P[i*ldp + h] * Q[c*ldq + h]: both operands reduce along their contiguous dimension
// (x g_out^T over the features, AS Gamma^T over the clusters); 128-bit loads along h when alignment allows.
template <int TR, int TN>
__device__ __forceinline__ void abt_tiled(const float* __restrict__ P, int64_t ldp, int n, const float* __restrict__ Q,
                                          int64_t ldq, int Nc, int Kred, float scale, float* __restrict__ D, int ldd,
                                          bool accumulate) {
  const int rt = ceil_div(n, TR), ct = ceil_div(Nc, TN);
  const bool vec = (Kred % 4 == 0) && (ldp % 4 == 0) && (ldq % 4 == 0) &&
                   (((reinterpret_cast<uintptr_t>(P) | reinterpret_cast<uintptr_t>(Q)) & 15) == 0);
  for (int t = threadIdx.x; t < rt * ct; t += blockDim.x) {
    const int r0 = (t / ct) * TR, c0 = (t % ct) * TN;
    float acc[TR][TN];
#pragma unroll
    for (int a = 0; a < TR; ++a)
#pragma unroll
      for (int b = 0; b < TN; ++b) acc[a][b] = 0.f;
    if (vec) {
      for (int h = 0; h < Kred; h += 4) {
        float4 pv[TR], qv[TN];
#pragma unroll
        for (int a = 0; a < TR; ++a)
          pv[a] = (r0 + a < n) ? *reinterpret_cast<const float4*>(P + (int64_t)(r0 + a) * ldp + h)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int b = 0; b < TN; ++b)
          qv[b] = (c0 + b < Nc) ? *reinterpret_cast<const float4*>(Q + (int64_t)(c0 + b) * ldq + h)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int a = 0; a < TR; ++a)
#pragma unroll
          for (int b = 0; b < TN; ++b)
            acc[a][b] += pv[a].x * qv[b].x + pv[a].y * qv[b].y + pv[a].z * qv[b].z + pv[a].w * qv[b].w;
      }
    } else {
      for (int h = 0; h < Kred; ++h) {
        float pv[TR], qv[TN];
#pragma unroll
        for (int a = 0; a < TR; ++a) pv[a] = (r0 + a < n) ? P[(int64_t)(r0 + a) * ldp + h] : 0.f;
#pragma unroll
        for (int b = 0; b < TN; ++b) qv[b] = (c0 + b < Nc) ? Q[(int64_t)(c0 + b) * ldq + h] : 0.f;
#pragma unroll
        for (int a = 0; a < TR; ++a)
#pragma unroll
          for (int b = 0; b < TN; ++b) acc[a][b] = fmaf(pv[a], qv[b], acc[a][b]);
      }
    }
#pragma unroll
    for (int a = 0; a < TR; ++a)
#pragma unroll
      for (int b = 0; b < TN; ++b)
        if (r0 + a < n && c0 + b < Nc) {
          float* d = D + (r0 + a) * ldd + c0 + b;
          *d = accumulate ? (*d + scale * acc[a][b]) : scale * acc[a][b];
        }
  }
}

template <bool SMEM>
__global__ void __launch_bounds__(1024) mincut_fwd_kernel(
    const float* __restrict__ logits, int64_t ldz, const float* __restrict__ x, int64_t ldx,
    const int* __restrict__ ptr, const int* __restrict__ rowptr, const int* __restrict__ col,
    const float* __restrict__ adj_val, float temp, int K, int H, int n_cap, float* __restrict__ s_soft,
    float* __restrict__ out, float* __restrict__ out_adj, float* __restrict__ ss_raw,
    float* __restrict__ adj_raw, float* __restrict__ stats, float* __restrict__ as_ws, int phase) {
  // phase 0: everything.  phase 1: S, A S, the two traces (num, den -> stats) and nothing else: the K x K and K x H
  // contractions are then done by ghscn_gemm3x_tn_segmented on the tensor cores (dense-bound for K >= 64), and
  // phase 2 finishes from ss_raw / adj_raw: norms, orthogonality loss, normalised coarse adjacency.
  extern __shared__ float smem[];
  __shared__ float red[32];
  __shared__ float dk[kMaxClusters];
  const int g = blockIdx.x;
  const int base = ptr[g];
  const int n = ptr[g + 1] - base;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = blockDim.x / 32;
  if (n > n_cap || n < 0) {
    // the caller's max-nodes-per-graph hint sized the shared tiles and is too small for this graph: poison the
    // graph's losses (NaN propagates to the batch means) instead of writing past the tiles
    if (tid == 0 && phase != 1) {
      float* st = stats + (int64_t)g * kStatsStride;
      st[3] = NAN; st[4] = NAN;
    }
    return;
  }

  float* Sg = s_soft + (int64_t)base * K;
  float* S = SMEM ? smem : Sg;
  float* AS = SMEM ? smem + (size_t)n_cap * K : as_ws + (int64_t)base * K;
  float* deg = SMEM ? smem + 2 * (size_t)n_cap * K : smem;

  float num = 0.f, den = 0.f;
  if (phase != 2) {
  // A. S = softmax(logits / temp).  Small K: one thread per node row (all rows of the graph in flight at once);
  //    large K: one warp per row.  Row degrees (row sums of A) by one thread per row.
  for (int i = tid; i < n; i += blockDim.x) {
    const int beg = rowptr[base + i], end = rowptr[base + i + 1];
    float d = (float)(end - beg);
    if (adj_val) {
      d = 0.f;
      for (int s = beg; s < end; ++s) d += adj_val[s];
    }
    deg[i] = d;
  }
  if (K <= 32) {
    for (int i = tid; i < n; i += blockDim.x) {
      const float* zr = logits + (int64_t)(base + i) * ldz;
      float z[32];
      float m = -INFINITY;
#pragma unroll
      for (int k = 0; k < 32; ++k)
        if (k < K) { z[k] = temp != 1.0f ? __fdiv_rn(zr[k], temp) : zr[k]; m = fmaxf(m, z[k]); }
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k)
        if (k < K) { z[k] = expf(z[k] - m); sum += z[k]; }
#pragma unroll
      for (int k = 0; k < 32; ++k)
        if (k < K) {
          const float p = __fdiv_rn(z[k], sum);
          S[i * K + k] = p;
          if (SMEM) Sg[i * K + k] = p;
        }
    }
  } else {
    for (int i = wid; i < n; i += nwarps) {
      const float* zr = logits + (int64_t)(base + i) * ldz;
      float m = -INFINITY;
      for (int k = lane; k < K; k += 32) {
        const float z = temp != 1.0f ? __fdiv_rn(zr[k], temp) : zr[k];
        m = fmaxf(m, z);
      }
      m = warp_max(m);
      float sum = 0.f;
      for (int k = lane; k < K; k += 32) {
        const float z = temp != 1.0f ? __fdiv_rn(zr[k], temp) : zr[k];
        const float p = expf(z - m);
        S[i * K + k] = p;
        sum += p;
      }
      sum = warp_sum(sum);
      for (int k = lane; k < K; k += 32) {
        const float p = __fdiv_rn(S[i * K + k], sum);
        S[i * K + k] = p;
        if (SMEM) Sg[i * K + k] = p;
      }
    }
  }
  __syncthreads();

  // B. AS = A S
  csr_times_s(rowptr, col, adj_val, S, base, n, K, AS);
  __syncthreads();

  // C. traces, S^T S, S^T A S, S^T X
  float pnum = 0.f, pden = 0.f;
  for (int e = tid; e < n * K; e += blockDim.x) {
    const float sv = S[e];
    pnum += sv * AS[e];
    pden += deg[e / K] * sv * sv;
  }
  num = block_sum(pnum, red);
  den = block_sum(pden, red);
  }  // phase != 2
  if (phase == 1) {
    if (tid == 0) {
      float* st = stats + (int64_t)g * kStatsStride;
      st[0] = num; st[1] = den;
    }
    return;
  }
  if (phase == 2) {
    num = stats[(int64_t)g * kStatsStride];
    den = stats[(int64_t)g * kStatsStride + 1];
  }

  float* ssg = ss_raw + (int64_t)g * K * K;
  float* oag = adj_raw + (int64_t)g * K * K;
  if (phase == 0) {
    atb_tiled<4, 4>(S, K, K, S, K, K, n, ssg, K);    // S^T S
    atb_tiled<4, 4>(S, K, K, AS, K, K, n, oag, K);   // S^T (A S)
    if (out != nullptr)                               // S^T X, X streamed from global memory
      atb_tiled<8, 4>(S, K, K, x + (int64_t)base * ldx, ldx, H, n, out + (int64_t)g * K * H, H);
    __syncthreads();  // ssg / oag visible to the whole CTA
  }

  // D. losses and the normalised coarse adjacency
  float pf = 0.f;
  for (int p = tid; p < K * K; p += blockDim.x) pf += ssg[p] * ssg[p];
  const float fro = sqrtf(block_sum(pf, red));
  const float inv_sqrt_k = __fdiv_rn(1.0f, sqrtf((float)K));
  float pr = 0.f;
  for (int p = tid; p < K * K; p += blockDim.x) {
    const int k = p / K, l = p - k * K;
    const float r = __fdiv_rn(ssg[p], fro) - (k == l ? inv_sqrt_k : 0.f);
    pr += r * r;
  }
  const float ortho = sqrtf(block_sum(pr, red));
  if (tid == 0) {
    float* st = stats + (int64_t)g * kStatsStride;
    st[0] = num; st[1] = den; st[2] = fro; st[3] = ortho; st[4] = -__fdiv_rn(num, den);
    st[5] = 0.f; st[6] = 0.f; st[7] = 0.f;
  }
  if (out_adj != nullptr) {
    for (int k = tid; k < K; k += blockDim.x) {
      float r = 0.f;
      for (int l = 0; l < K; ++l)
        if (l != k) r += oag[k * K + l];
      dk[k] = sqrtf(r) + 1e-15f;
    }
    __syncthreads();
    float* ng = out_adj + (int64_t)g * K * K;
    for (int p = tid; p < K * K; p += blockDim.x) {
      const int k = p / K, l = p - k * K;
      ng[p] = (k == l) ? 0.f : __fdiv_rn(__fdiv_rn(oag[p], dk[l]), dk[k]);
    }
  }
}

// losses[0] = mean_g(-num/den), losses[1] = mean_g(ortho_g); fixed reduction order.
__global__ void __launch_bounds__(256) mincut_reduce_losses_kernel(const float* __restrict__ stats, int B,
                                                                   float* __restrict__ losses) {
  __shared__ float red[32];
  float a = 0.f, b = 0.f;
  for (int g = threadIdx.x; g < B; g += blockDim.x) {
    a += stats[(int64_t)g * kStatsStride + 4];
    b += stats[(int64_t)g * kStatsStride + 3];
  }
  a = block_sum(a, red);
  b = block_sum(b, red);
  if (threadIdx.x == 0) {
    losses[0] = __fdiv_rn(a, (float)B);
    losses[1] = __fdiv_rn(b, (float)B);
  }
}

template <bool SMEM>
__global__ void __launch_bounds__(1024) mincut_bwd_kernel(
    const float* __restrict__ s_soft, const float* __restrict__ x, int64_t ldx, const int* __restrict__ ptr,
    const int* __restrict__ rowptr, const int* __restrict__ col, const float* __restrict__ adj_val,
    const int* __restrict__ rowptr_t, const int* __restrict__ col_t, const float* __restrict__ adj_val_t,
    float temp, int B, int N, int K, int H, int n_cap, const float* __restrict__ ss_raw,
    const float* __restrict__ adj_raw, const float* __restrict__ stats, const float* __restrict__ g_out,
    const float* __restrict__ g_out_adj, const float* __restrict__ g_losses, float* __restrict__ d_logits,
    int64_t lddz, float* __restrict__ d_x, int64_t lddx, float* __restrict__ ws) {
  extern __shared__ float smem[];
  __shared__ float red[32];
  __shared__ float dk[kMaxClusters], dr[kMaxClusters];
  const int g = blockIdx.x;
  const int base = ptr[g];
  const int n = ptr[g + 1] - base;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = blockDim.x / 32;
  const size_t nk_cap = (size_t)n_cap * K;
  if (g == B - 1) {
    // rows past the last graph (padding rows of a bucketed batch: `ptr` may cover fewer than N rows) get zero
    // gradients, so callers never see unwritten memory
    const int tail0 = ptr[B];
    for (int64_t e = (int64_t)tail0 * K + tid; e < (int64_t)N * K; e += blockDim.x)
      d_logits[(e / K) * lddz + e % K] = 0.f;
    if (d_x != nullptr)
      for (int64_t e = (int64_t)tail0 * H + tid; e < (int64_t)N * H; e += blockDim.x) d_x[(e / H) * lddx + e % H] = 0.f;
  }
  if (n > n_cap || n < 0) {                       // stale max-nodes hint: poison the gradient, touch nothing else
    if (tid == 0 && n > 0) d_logits[(int64_t)base * lddz] = NAN;
    return;
  }

  // shared layout: [Gsym K*K][Gam K*K][deg n_cap] (+ [S][AS][ATS][dS] when SMEM)
  float* Gsym = smem;
  float* Gam = smem + K * K;
  float* deg = smem + 2 * K * K;
  const float* Sg = s_soft + (int64_t)base * K;
  float* S = SMEM ? deg + n_cap : nullptr;
  float* AS = SMEM ? S + nk_cap : ws + (int64_t)base * K;
  float* ATS = SMEM ? AS + nk_cap : ws + (int64_t)(ptr[B] + base) * K;
  float* dS = SMEM ? ATS + nk_cap : ws + (int64_t)(2 * ptr[B] + base) * K;
  if (SMEM) {
    for (int e = tid; e < n * K; e += blockDim.x) S[e] = Sg[e];
  }
  const float* Sr = SMEM ? S : Sg;
  for (int i = tid; i < n; i += blockDim.x) {
    float d = 0.f;
    for (int s = rowptr[base + i]; s < rowptr[base + i + 1]; ++s) d += adj_value(adj_val, s);
    deg[i] = d;
  }
  __syncthreads();
  csr_times_s(rowptr, col, adj_val, Sr, base, n, K, AS);
  csr_times_s(rowptr_t, col_t, adj_val_t, Sr, base, n, K, ATS);

  const float* st = stats + (int64_t)g * kStatsStride;
  const float num = st[0], den = st[1], fro = st[2], nrm = st[3];
  const float gmc = g_losses ? __fdiv_rn(g_losses[0], (float)B) : 0.f;
  const float go = g_losses ? __fdiv_rn(g_losses[1], (float)B) : 0.f;
  const float* ssg = ss_raw + (int64_t)g * K * K;
  const float* oag = adj_raw + (int64_t)g * K * K;
  const float inv_sqrt_k = __fdiv_rn(1.0f, sqrtf((float)K));

  // ortho: G = R/||R||, G' = (G - M <G,M>)/F, Gsym = go * (G' + G'^T)
  float pin = 0.f;
  for (int p = tid; p < K * K; p += blockDim.x) {
    const int k = p / K, l = p - k * K;
    const float M = ssg[p] / fro;
    const float G = nrm > 0.f ? (M - (k == l ? inv_sqrt_k : 0.f)) / nrm : 0.f;
    pin += G * M;
  }
  const float inner = block_sum(pin, red);
  for (int p = tid; p < K * K; p += blockDim.x) {
    const int k = p / K, l = p - k * K;
    const float M = ssg[p] / fro, Mt = ssg[l * K + k] / fro;
    const float G = nrm > 0.f ? (M - (k == l ? inv_sqrt_k : 0.f)) / nrm : 0.f;
    const float Gt = nrm > 0.f ? (Mt - (k == l ? inv_sqrt_k : 0.f)) / nrm : 0.f;
    Gsym[p] = go * ((G - M * inner) + (Gt - Mt * inner)) / fro;
  }
  // Gamma = dL/d(S^T A S): trace term of the mincut loss + chain through the normalised out_adj
  if (g_out_adj != nullptr) {
    const float* gb = g_out_adj + (int64_t)g * K * K;
    for (int k = tid; k < K; k += blockDim.x) {
      float r = 0.f;
      for (int l = 0; l < K; ++l)
        if (l != k) r += oag[k * K + l];
      dk[k] = sqrtf(r) + 1e-15f;
      dr[k] = r;
    }
    __syncthreads();
    for (int k = tid; k < K; k += blockDim.x) {
      float acc = 0.f;  // sum_l Gbar[k][l] N[k][l] + Gbar[l][k] N[l][k]
      for (int l = 0; l < K; ++l) {
        if (l == k) continue;
        const float nkl = (oag[k * K + l] / dk[l]) / dk[k];
        const float nlk = (oag[l * K + k] / dk[k]) / dk[l];
        acc += gb[k * K + l] * nkl + gb[l * K + k] * nlk;
      }
      const float ddk = -acc / dk[k];
      const float sq = sqrtf(dr[k]);
      dr[k] = sq > 0.f ? ddk / (2.f * sq) : 0.f;  // dL/d(rowsum_k)
    }
    __syncthreads();
    for (int p = tid; p < K * K; p += blockDim.x) {
      const int k = p / K, l = p - k * K;
      Gam[p] = (k == l) ? 0.f : gb[p] / (dk[k] * dk[l]) + dr[k];
    }
  } else {
    for (int p = tid; p < K * K; p += blockDim.x) Gam[p] = 0.f;
  }
  __syncthreads();
  for (int k = tid; k < K; k += blockDim.x) Gam[k * K + k] += -gmc / den;
  __syncthreads();

  const float cden = gmc * num / (den * den);
  const bool diag_only = (g_out_adj == nullptr);
  const float gdiag = -gmc / den;
  const float* gog = g_out ? g_out + (int64_t)g * K * H : nullptr;
  // dS = [mincut trace / out_adj chain] + [den term] + [ortho] + [out term], as tiled small GEMMs
  if (diag_only) {
    for (int e = tid; e < n * K; e += blockDim.x)
      dS[e] = gdiag * (AS[e] + ATS[e]) + cden * 2.f * deg[e / K] * Sr[e];
  } else {
    abt_tiled<4, 4>(AS, K, n, Gam, K, K, K, 1.f, dS, K, false);          // AS  Gamma^T
    __syncthreads();
    ab_tiled<4, 4>(ATS, K, n, Gam, K, K, K, 1.f, dS, K, true);           // A^T S Gamma
    __syncthreads();
    for (int e = tid; e < n * K; e += blockDim.x) dS[e] += cden * 2.f * deg[e / K] * Sr[e];
  }
  __syncthreads();
  ab_tiled<4, 4>(Sr, K, n, Gsym, K, K, K, 1.f, dS, K, true);             // S (G' + G'^T) go
  if (gog) {
    __syncthreads();
    abt_tiled<4, 4>(x + (int64_t)base * ldx, ldx, n, gog, H, K, H, 1.f, dS, K, true);   // x g_out^T
  }
  __syncthreads();
  // softmax backward, one warp per node row
  for (int i = wid; i < n; i += nwarps) {
    float dot = 0.f;
    for (int k = lane; k < K; k += 32) dot += dS[i * K + k] * Sr[i * K + k];
    dot = warp_sum(dot);
    for (int k = lane; k < K; k += 32) {
      float dz = Sr[i * K + k] * (dS[i * K + k] - dot);
      if (temp != 1.0f) dz = dz / temp;
      d_logits[(int64_t)(base + i) * lddz + k] = dz;
    }
  }
  if (d_x != nullptr) {
    float* dxg = d_x + (int64_t)base * lddx;
    if (gog) {
      // dX = S g_out: [n,K] x [K,H]; thread tile 4 rows x 4 columns, columns contiguous
      const int rt = ceil_div(n, 4), ct = ceil_div(H, 4);
      for (int t = tid; t < rt * ct; t += blockDim.x) {
        const int r0 = (t / ct) * 4, c0 = (t % ct) * 4;
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
        for (int k = 0; k < K; ++k) {
          float sv[4], gv[4];
#pragma unroll
          for (int a = 0; a < 4; ++a) sv[a] = (r0 + a < n) ? Sr[(r0 + a) * K + k] : 0.f;
#pragma unroll
          for (int b = 0; b < 4; ++b) gv[b] = (c0 + b < H) ? __ldg(gog + (int64_t)k * H + c0 + b) : 0.f;
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(sv[a], gv[b], acc[a][b]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b)
            if (r0 + a < n && c0 + b < H) dxg[(int64_t)(r0 + a) * lddx + c0 + b] = acc[a][b];
      }
    } else {
      for (int e = tid; e < n * H; e += blockDim.x) dxg[(int64_t)(e / H) * lddx + e % H] = 0.f;
    }
  }
}

static inline size_t fwd_smem_bytes(int n_cap, int K, bool smem) {
  return smem ? (2 * (size_t)n_cap * K + n_cap) * 4 : (size_t)n_cap * 4;
}
static inline size_t bwd_smem_bytes(int n_cap, int K, bool smem) {
  return (2 * (size_t)K * K + n_cap + (smem ? 4 * (size_t)n_cap * K : 0)) * 4;
}

}  // namespace ghscn

using namespace ghscn;

extern "C" {

size_t ghscn_mincut_workspace_bytes(int64_t num_nodes, int64_t num_graphs, int64_t num_clusters) {
  (void)num_graphs;
  if (num_nodes < 0 || num_clusters < 0) return 0;
  return (size_t)3 * num_nodes * num_clusters * 4 + 256;
}

static int mincut_fwd_impl(const float* logits, int64_t ldz, const float* x, int64_t ldx, const int32_t* ptr,
                     const int32_t* rowptr, const int32_t* col, const float* adj_val, float temp,
                     int64_t num_graphs, int64_t num_nodes, int64_t num_clusters, int64_t num_feat,
                     int32_t max_nodes_per_graph, float* s_soft, float* out, float* out_adj, float* ss_raw,
                     float* adj_raw, float* stats, float* losses, void* workspace, size_t workspace_bytes,
                     ghscn_stream_t stream_, int phase) {
  GHSCN_REQUIRE(num_graphs >= 0 && num_nodes >= 0 && num_clusters > 0 && num_feat >= 0);
  GHSCN_REQUIRE(num_graphs < ((int64_t)1 << 31) && num_nodes < ((int64_t)1 << 31));
  if (num_clusters > kMaxClusters) return GHSCN_E_UNSUPPORTED;
  if (num_graphs == 0) return GHSCN_OK;
  GHSCN_REQUIRE(logits && ptr && rowptr && s_soft && ss_raw && adj_raw && stats && losses);
  GHSCN_REQUIRE(ldz >= num_clusters && (out == nullptr || (x != nullptr && ldx >= num_feat)));
  GHSCN_REQUIRE(max_nodes_per_graph > 0);
  cudaStream_t stream = as_stream(stream_);
  const int K = (int)num_clusters, H = (int)num_feat, n_cap = max_nodes_per_graph;
  const int threads = mincut_threads(num_graphs);
  // the split phases exchange S and A S through HBM (s_soft, workspace): they use the workspace variant
  const bool smem = phase == 0 && fwd_smem_bytes(n_cap, K, true) <= kSmemBudget;
  const size_t shm = fwd_smem_bytes(n_cap, K, smem);
  if (shm > kSmemBudget) return GHSCN_E_UNSUPPORTED;
  float* as_ws = nullptr;
  if (!smem) {
    if (workspace == nullptr || workspace_bytes < (size_t)num_nodes * K * 4) return GHSCN_E_WORKSPACE;
    as_ws = static_cast<float*>(workspace);
  }
  if (smem) {
    cudaFuncSetAttribute(mincut_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget);
    mincut_fwd_kernel<true><<<(unsigned)num_graphs, threads, shm, stream>>>(
        logits, ldz, x, ldx, ptr, rowptr, col, adj_val, temp, K, H, n_cap, s_soft, out, out_adj, ss_raw, adj_raw,
        stats, as_ws, phase);
  } else {
    mincut_fwd_kernel<false><<<(unsigned)num_graphs, threads, shm, stream>>>(
        logits, ldz, x, ldx, ptr, rowptr, col, adj_val, temp, K, H, n_cap, s_soft, out, out_adj, ss_raw, adj_raw,
        stats, as_ws, phase);
  }
  if (phase == 1) {
    GHSCN_LAUNCH_CHECK();
    return GHSCN_OK;
  }
  mincut_reduce_losses_kernel<<<1, 256, 0, stream>>>(stats, (int)num_graphs, losses);
  GHSCN_LAUNCH_CHECK_N(2);
  return GHSCN_OK;
}

int ghscn_mincut_fwd(const float* logits, int64_t ldz, const float* x, int64_t ldx, const int32_t* ptr,
                     const int32_t* rowptr, const int32_t* col, const float* adj_val, float temp,
                     int64_t num_graphs, int64_t num_nodes, int64_t num_clusters, int64_t num_feat,
                     int32_t max_nodes_per_graph, float* s_soft, float* out, float* out_adj, float* ss_raw,
                     float* adj_raw, float* stats, float* losses, void* workspace, size_t workspace_bytes,
                     ghscn_stream_t stream_) {
  return mincut_fwd_impl(logits, ldz, x, ldx, ptr, rowptr, col, adj_val, temp, num_graphs, num_nodes, num_clusters, num_feat, max_nodes_per_graph, s_soft, out, out_adj, ss_raw, adj_raw, stats, losses, workspace, workspace_bytes, stream_, 0);
}

int ghscn_mincut_fwd_phase(const float* logits, int64_t ldz, const float* x, int64_t ldx, const int32_t* ptr,
                     const int32_t* rowptr, const int32_t* col, const float* adj_val, float temp,
                     int64_t num_graphs, int64_t num_nodes, int64_t num_clusters, int64_t num_feat,
                     int32_t max_nodes_per_graph, float* s_soft, float* out, float* out_adj, float* ss_raw,
                     float* adj_raw, float* stats, float* losses, void* workspace, size_t workspace_bytes,
                     ghscn_stream_t stream_, int32_t phase) {
  if (phase < 0 || phase > 2) return GHSCN_E_INVALID;
  return mincut_fwd_impl(logits, ldz, x, ldx, ptr, rowptr, col, adj_val, temp, num_graphs, num_nodes, num_clusters, num_feat, max_nodes_per_graph, s_soft, out, out_adj, ss_raw, adj_raw, stats, losses, workspace, workspace_bytes, stream_, phase);
}

int ghscn_mincut_bwd(const float* s_soft, const float* x, int64_t ldx, const int32_t* ptr, const int32_t* rowptr,
                     const int32_t* col, const float* adj_val, const int32_t* rowptr_t, const int32_t* col_t,
                     const float* adj_val_t, float temp, int64_t num_graphs, int64_t num_nodes,
                     int64_t num_clusters, int64_t num_feat, int32_t max_nodes_per_graph, const float* ss_raw,
                     const float* adj_raw, const float* stats, const float* g_out, const float* g_out_adj,
                     const float* g_losses, float* d_logits, int64_t lddz, float* d_x, int64_t lddx,
                     void* workspace, size_t workspace_bytes, ghscn_stream_t stream_) {
  GHSCN_REQUIRE(num_graphs >= 0 && num_nodes >= 0 && num_clusters > 0 && num_feat >= 0);
  if (num_clusters > kMaxClusters) return GHSCN_E_UNSUPPORTED;
  if (num_graphs == 0) return GHSCN_OK;
  GHSCN_REQUIRE(s_soft && ptr && rowptr && rowptr_t && ss_raw && adj_raw && stats && d_logits);
  GHSCN_REQUIRE(lddz >= num_clusters && max_nodes_per_graph > 0);
  GHSCN_REQUIRE((g_out == nullptr && d_x == nullptr) || x != nullptr);
  cudaStream_t stream = as_stream(stream_);
  const int K = (int)num_clusters, H = (int)num_feat, n_cap = max_nodes_per_graph;
  const int threads = mincut_threads(num_graphs);
  const bool smem = bwd_smem_bytes(n_cap, K, true) <= kSmemBudget;
  const size_t shm = bwd_smem_bytes(n_cap, K, smem);
  if (shm > kSmemBudget) return GHSCN_E_UNSUPPORTED;
  float* ws = nullptr;
  if (!smem) {
    if (workspace == nullptr || workspace_bytes < (size_t)3 * num_nodes * K * 4) return GHSCN_E_WORKSPACE;
    ws = static_cast<float*>(workspace);
  }
  if (smem) {
    cudaFuncSetAttribute(mincut_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget);
    mincut_bwd_kernel<true><<<(unsigned)num_graphs, threads, shm, stream>>>(
        s_soft, x, ldx, ptr, rowptr, col, adj_val, rowptr_t, col_t, adj_val_t, temp, (int)num_graphs, (int)num_nodes, K,
        H, n_cap, ss_raw, adj_raw, stats, g_out, g_out_adj, g_losses, d_logits, lddz, d_x, lddx, ws);
  } else {
    cudaFuncSetAttribute(mincut_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget);
    mincut_bwd_kernel<false><<<(unsigned)num_graphs, threads, shm, stream>>>(
        s_soft, x, ldx, ptr, rowptr, col, adj_val, rowptr_t, col_t, adj_val_t, temp, (int)num_graphs, (int)num_nodes, K,
        H, n_cap, ss_raw, adj_raw, stats, g_out, g_out_adj, g_losses, d_logits, lddz, d_x, lddx, ws);
  }
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

}  // extern "C"
