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
                                            int base, int n, int K, float* __restrict__ buf, int ldb = 0) {
  if (ldb == 0) ldb = K;
  for (int e = threadIdx.x; e < n * K; e += blockDim.x) {
    const int i = e / K, k = e - i * K;
    float acc = 0.f;
    const int beg = rowptr[base + i], end = rowptr[base + i + 1];
    for (int s = beg; s < end; ++s) {
      const int c = col[s] - base;
      if (c >= 0 && c < n) acc += adj_value(adj_val, s) * S[c * K + k];
    }
    buf[(size_t)i * ldb + k] = acc;
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

  // B. AS = A S (phase 3: the caller runs the K2 SpMM over the whole batch instead)
  if (phase != 3) {
    csr_times_s(rowptr, col, adj_val, S, base, n, K, AS);
    __syncthreads();
  }

  // C. traces, S^T S, S^T A S, S^T X
  float pnum = 0.f, pden = 0.f;
#pragma unroll 4
  for (int e = tid; e < n * K; e += blockDim.x) {
    const float sv = S[e];
    if (phase != 3) pnum += sv * AS[e];
    pden += deg[e / K] * sv * sv;
  }
  num = block_sum(pnum, red);
  den = block_sum(pden, red);
  }  // phase != 2
  if (phase == 1 || phase == 3) {
    if (tid == 0) {
      float* st = stats + (int64_t)g * kStatsStride;
      st[0] = num; st[1] = den;
    }
    return;
  }
  if (phase == 2) {
    // num = Tr(S^T A S), literally as dense_mincut_pool takes it (`_rank3_trace(out_adj)`) from the contraction result
    const float* oa = adj_raw + (int64_t)g * K * K;
    float pt = 0.f;
    for (int k = tid; k < K; k += blockDim.x) pt += oa[k * K + k];
    num = block_sum(pt, red);
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
#pragma unroll 8
  for (int p = tid; p < K * K; p += blockDim.x) pf += ssg[p] * ssg[p];
  const float fro = sqrtf(block_sum(pf, red));
  const float inv_sqrt_k = __fdiv_rn(1.0f, sqrtf((float)K));
  float pr = 0.f;
#pragma unroll 8
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
    if (K <= 32) {
      for (int k = tid; k < K; k += blockDim.x) {
        float r = 0.f;
        for (int l = 0; l < K; ++l)
          if (l != k) r += oag[k * K + l];
        dk[k] = sqrtf(r) + 1e-15f;
      }
    } else {
      for (int k = wid; k < K; k += nwarps) {      // a warp per row: coalesced
        float r = 0.f;
#pragma unroll 4
        for (int l = lane; l < K; l += 32)
          if (l != k) r += oag[k * K + l];
        r = warp_sum(r);
        if (lane == 0) dk[k] = sqrtf(r) + 1e-15f;
      }
    }
    __syncthreads();
    float* ng = out_adj + (int64_t)g * K * K;
#pragma unroll 8
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

// Gsym = d(ortho loss)/d(S^T S) symmetrised, Gam = dL/d(S^T A S) (trace term of the mincut loss + the chain through the
// normalised out_adj); both [K,K], in shared or global memory.  Called by the whole CTA; ends with a __syncthreads().
// Every pass over a K x K matrix is coalesced and keeps several loads in flight per thread: row reductions take a warp
// per row, column reductions a thread per (column, row slice), and M^T for the symmetrisation goes through a 32 x 33
// shared tile per warp -- at K = 128 the matrices are 64 KB per graph, and strided, one-load-at-a-time passes over
// them were the whole cost of the split backward.
struct MincutBwdScratch {
  float red[32];
  float dk[kMaxClusters], dr[kMaxClusters], ra[kMaxClusters];
  float colpart[1024];
};
typedef float MincutTile[32][33];

// `tiles`: one 32 x 33 tile for each of the first `tile_warps` warps (the other warps skip the transposing pass)
__device__ __forceinline__ void mincut_bwd_coefficients(int g, int B, int K, const float* __restrict__ ss_raw,
                                                        const float* __restrict__ adj_raw,
                                                        const float* __restrict__ stats,
                                                        const float* __restrict__ g_out_adj,
                                                        const float* __restrict__ g_losses, float* __restrict__ Gsym,
                                                        float* __restrict__ Gam, MincutBwdScratch& sc, MincutTile* tiles,
                                                        int tile_warps) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = blockDim.x >> 5;
  float* dk = sc.dk;
  float* dr = sc.dr;
  float* ra = sc.ra;
  const float* st = stats + (int64_t)g * kStatsStride;
  const float den = st[1], fro = st[2], nrm = st[3];
  const float gmc = g_losses ? __fdiv_rn(g_losses[0], (float)B) : 0.f;
  const float go = g_losses ? __fdiv_rn(g_losses[1], (float)B) : 0.f;
  const float* __restrict__ ssg = ss_raw + (int64_t)g * K * K;
  const float* __restrict__ oag = adj_raw + (int64_t)g * K * K;
  const float inv_sqrt_k = __fdiv_rn(1.0f, sqrtf((float)K));
  const int KK = K * K;

  // ortho: G = R/||R||, G' = (G - M <G,M>)/F, Gsym = go * (G' + G'^T)
  float pin = 0.f;
#pragma unroll 8
  for (int p = tid; p < KK; p += blockDim.x) {
    const int k = p / K, l = p - k * K;
    const float M = __ldg(ssg + p) / fro;
    const float G = nrm > 0.f ? (M - (k == l ? inv_sqrt_k : 0.f)) / nrm : 0.f;
    pin += G * M;
  }
  const float inner = block_sum(pin, sc.red);
  if (K <= 32) {
    for (int p = tid; p < KK; p += blockDim.x) {
      const int k = p / K, l = p - k * K;
      const float M = ssg[p] / fro, Mt = ssg[l * K + k] / fro;
      const float G = nrm > 0.f ? (M - (k == l ? inv_sqrt_k : 0.f)) / nrm : 0.f;
      const float Gt = nrm > 0.f ? (Mt - (k == l ? inv_sqrt_k : 0.f)) / nrm : 0.f;
      Gsym[p] = go * ((G - M * inner) + (Gt - Mt * inner)) / fro;
    }
  } else if (wid < tile_warps) {
    const int kt = ceil_div(K, 32);
    float (*tile)[33] = tiles[wid];
    for (int t = wid; t < kt * kt; t += min(nwarps, tile_warps)) {
      const int kb = (t / kt) * 32, lb = (t % kt) * 32;
      __syncwarp();
#pragma unroll 8
      for (int a = 0; a < 32; ++a)                 // tile[a][b] = M_raw[lb + a][kb + b], rows read coalesced
        tile[a][lane] = (lb + a < K && kb + lane < K) ? __ldg(ssg + (lb + a) * K + kb + lane) : 0.f;
      __syncwarp();
      const int l = lb + lane;
#pragma unroll 8
      for (int r = 0; r < 32; ++r) {
        const int k = kb + r;
        if (k < K && l < K) {
          const float M = __ldg(ssg + k * K + l) / fro, Mt = tile[lane][r] / fro;
          const float G = nrm > 0.f ? (M - (k == l ? inv_sqrt_k : 0.f)) / nrm : 0.f;
          const float Gt = nrm > 0.f ? (Mt - (k == l ? inv_sqrt_k : 0.f)) / nrm : 0.f;
          Gsym[k * K + l] = go * ((G - M * inner) + (Gt - Mt * inner)) / fro;
        }
      }
    }
  }
  // Gamma = dL/d(S^T A S): trace term of the mincut loss + chain through the normalised out_adj
  if (g_out_adj != nullptr) {
    const float* __restrict__ gb = g_out_adj + (int64_t)g * K * K;
    for (int k = wid; k < K; k += nwarps) {      // row sums without the diagonal, a warp per row
      float r = 0.f;
#pragma unroll 4
      for (int l = lane; l < K; l += 32)
        if (l != k) r += __ldg(oag + k * K + l);
      r = warp_sum(r);
      if (lane == 0) { dk[k] = sqrtf(r) + 1e-15f; dr[k] = r; }
    }
    __syncthreads();
    // acc_k = sum_l Gbar[k][l] N[k][l] + Gbar[l][k] N[l][k]: the first sum a warp per row, the second a thread per
    // (column, row slice) with the slices combined in order
    for (int k = wid; k < K; k += nwarps) {
      float acc = 0.f;
#pragma unroll 4
      for (int l = lane; l < K; l += 32)
        if (l != k) acc += __ldg(gb + k * K + l) * ((__ldg(oag + k * K + l) / dk[l]) / dk[k]);
      acc = warp_sum(acc);
      if (lane == 0) ra[k] = acc;
    }
    const int parts = max(1, min((int)blockDim.x / K, 1024 / K));
    if (tid < parts * K) {
      const int k = tid % K, part = tid / K;
      float acc = 0.f;
#pragma unroll 4
      for (int l = part; l < K; l += parts)
        if (l != k) acc += __ldg(gb + l * K + k) * ((__ldg(oag + l * K + k) / dk[k]) / dk[l]);
      sc.colpart[part * K + k] = acc;
    }
    __syncthreads();
    for (int k = tid; k < K; k += blockDim.x) {
      float acc = ra[k];
      for (int q = 0; q < parts; ++q) acc += sc.colpart[q * K + k];
      const float ddk = -acc / dk[k];
      const float sq = sqrtf(dr[k]);
      ra[k] = sq > 0.f ? ddk / (2.f * sq) : 0.f;  // dL/d(rowsum_k)
    }
    __syncthreads();
#pragma unroll 8
    for (int p = tid; p < KK; p += blockDim.x) {
      const int k = p / K, l = p - k * K;
      Gam[p] = (k == l) ? -gmc / den : __ldg(gb + p) / (dk[k] * dk[l]) + ra[k];
    }
  } else {
    for (int p = tid; p < KK; p += blockDim.x) {
      const int k = p / K, l = p - k * K;
      Gam[p] = (k == l) ? -gmc / den : 0.f;
    }
  }
  __syncthreads();
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
  __shared__ MincutBwdScratch sc;
  __shared__ MincutTile tiles[4];
  float* red = sc.red;
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

  mincut_bwd_coefficients(g, B, K, ss_raw, adj_raw, stats, g_out_adj, g_losses, Gsym, Gam, sc, tiles, 4);
  const float* st = stats + (int64_t)g * kStatsStride;
  const float num = st[0], den = st[1];
  const float gmc = g_losses ? __fdiv_rn(g_losses[0], (float)B) : 0.f;
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

// ---- split backward (K >= 64): prepare / softmax passes around the segment GEMMs below -----------------------------------
// prepare: the graph's stacked [Gamma^T; Gamma; Gsym] (built in place in global memory: no K x K shared tiles, so several
// CTAs share an SM) and the elementwise part of dS.  [A S | A^T S] come from two launches of the K2 SpMM over the whole
// batch (a collated batch has no edge between graphs, so the block-diagonal product is the per-graph one).
constexpr int kPrepareTileWarps = 16;

__global__ void __launch_bounds__(1024) mincut_bwd_prepare_kernel(
    const float* __restrict__ s_soft, const int* __restrict__ ptr, const int* __restrict__ rowptr,
    const float* __restrict__ adj_val, int B, int N, int K, int n_cap,
    const float* __restrict__ ss_raw, const float* __restrict__ adj_raw, const float* __restrict__ stats,
    const float* __restrict__ g_out_adj, const float* __restrict__ g_losses, float* __restrict__ ws) {
  extern __shared__ float smem[];
  __shared__ MincutBwdScratch sc;
  MincutTile* tiles = reinterpret_cast<MincutTile*>(smem + n_cap);      // kPrepareTileWarps tiles
  const int g = blockIdx.x, tid = threadIdx.x;
  const int base = ptr[g], n = ptr[g + 1] - base;
  if (n > n_cap || n < 0) return;                                   // the softmax pass poisons this graph's gradient
  float* dS = ws + (size_t)2 * N * K + (size_t)base * K;            // [n][K]
  float* gstack = ws + (size_t)3 * N * K + (size_t)g * 2 * K * K;   // [2K][K] = [Gamma; Gsym]
  float* Gam = gstack;
  float* Gsym = gstack + K * K;
  float* deg = smem;
  const float* Sg = s_soft + (int64_t)base * K;
  for (int i = tid; i < n; i += blockDim.x) {
    float d = 0.f;
    for (int sidx = rowptr[base + i]; sidx < rowptr[base + i + 1]; ++sidx) d += adj_value(adj_val, sidx);
    deg[i] = d;
  }
  mincut_bwd_coefficients(g, B, K, ss_raw, adj_raw, stats, g_out_adj, g_losses, Gsym, Gam, sc, tiles,
                          kPrepareTileWarps);
  const float* st = stats + (int64_t)g * kStatsStride;
  const float gmc = g_losses ? __fdiv_rn(g_losses[0], (float)B) : 0.f;
  const float cden = gmc * st[0] / (st[1] * st[1]);
  for (int e = tid; e < n * K; e += blockDim.x) dS[e] = cden * 2.f * deg[e / K] * Sg[e];
}

// softmax backward from the accumulated dS, one warp per node row; zero rows behind the last graph / without g_out
__global__ void __launch_bounds__(256) mincut_bwd_softmax_kernel(
    const float* __restrict__ s_soft, const int* __restrict__ ptr, float temp, int B, int N, int K, int H, int n_cap,
    bool has_g_out, float* __restrict__ d_logits, int64_t lddz, float* __restrict__ d_x, int64_t lddx,
    const float* __restrict__ ws) {
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = blockDim.x / 32;
  const int base = ptr[g], n = ptr[g + 1] - base;
  if (g == B - 1) {
    const int tail0 = ptr[B];
    for (int64_t e = (int64_t)tail0 * K + tid; e < (int64_t)N * K; e += blockDim.x)
      d_logits[(e / K) * lddz + e % K] = 0.f;
    if (d_x != nullptr)
      for (int64_t e = (int64_t)tail0 * H + tid; e < (int64_t)N * H; e += blockDim.x) d_x[(e / H) * lddx + e % H] = 0.f;
  }
  if (n > n_cap || n < 0) {
    if (tid == 0 && n > 0) d_logits[(int64_t)base * lddz] = NAN;
    return;
  }
  const float* dS = ws + (size_t)2 * N * K + (size_t)base * K;
  const float* Sg = s_soft + (int64_t)base * K;
  for (int i = wid; i < n; i += nwarps) {
    float sv[4], dv[4];                         // K <= 128: four values per lane
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = lane + 32 * j;
      sv[j] = k < K ? Sg[i * K + k] : 0.f;
      dv[j] = k < K ? dS[i * K + k] : 0.f;
      dot = fmaf(dv[j], sv[j], dot);
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = lane + 32 * j;
      if (k < K) {
        float dz = sv[j] * (dv[j] - dot);
        if (temp != 1.0f) dz = dz / temp;
        d_logits[(int64_t)(base + i) * lddz + k] = dz;
      }
    }
  }
  if (d_x != nullptr && !has_g_out) {
    float* dxg = d_x + (int64_t)base * lddx;
    for (int e = tid; e < n * H; e += blockDim.x) dxg[(int64_t)(e / H) * lddx + e % H] = 0.f;
  }
}

// ---- tiled segment GEMM for the split backward ----------------------------------------------------------------------
// C_g[n_g, Nc] (+)= A_g[n_g, Kr] . B_g   for every graph g: A_g / C_g are the graph's node rows (ptr), B_g its own small
// matrix ([Kr, Nc] row-major, or [Nc, Kr] when B_T).  64 x 64 output tile per CTA, 16-deep K steps through shared
// memory, 4 x 4 register tile per thread (2 LDS.128 per 16 FMA).  Kr, Nc and all row strides are multiples of 4.
constexpr int kSegTile = 64, kSegK = 16;

// TN = output columns per thread (4 or 8): tile width 16 * TN.  The 128-wide tile does 32 FMAs per 3 LDS.128.
template <bool B_T, bool ACC, int TN>
__global__ void __launch_bounds__(256) mincut_seg_gemm_kernel(const float* __restrict__ A, int64_t lda,
                                                              const int* __restrict__ ptr,
                                                              const float* __restrict__ Bm, int64_t ldb,
                                                              int64_t b_stride, int Kr, int Nc,
                                                              float* __restrict__ C, int64_t ldc, int n_cap) {
  constexpr int WN = 16 * TN, NB = TN / 4;       // tile width, float4 loads of B per thread and K step
  __shared__ __align__(16) float As[kSegK][kSegTile + 4];
  __shared__ __align__(16) float Bs[kSegK][WN + 4];
  const int g = blockIdx.z;
  const int base = ptr[g], n = ptr[g + 1] - base;
  const int r0 = blockIdx.y * kSegTile, c0 = blockIdx.x * WN;
  if (r0 >= n || n > n_cap) return;
  const float* Ag = A + (int64_t)base * lda;
  const float* Bg = Bm + (int64_t)g * b_stride;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int lr = tid >> 2, lk = (tid & 3) * 4;   // A (and B when B_T): row = tid / 4, k4 = tid % 4
  float acc[4][TN];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < TN; ++b) acc[a][b] = 0.f;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  auto load_a = [&](int k0) -> float4 {
    return (r0 + lr < n && k0 + lk < Kr) ? *reinterpret_cast<const float4*>(Ag + (int64_t)(r0 + lr) * lda + k0 + lk)
                                         : zero;
  };
  auto load_b = [&](int k0, int j) -> float4 {
    if (B_T) {                                    // B [Nc, Kr]: rows c0 + lr + 64 j
      const int c = c0 + lr + 64 * j;
      return (c < Nc && k0 + lk < Kr) ? *reinterpret_cast<const float4*>(Bg + (int64_t)c * ldb + k0 + lk) : zero;
    }
    const int f = tid + 256 * j, bk = f / (WN / 4), bc = (f % (WN / 4)) * 4;   // B [Kr, Nc]
    return (k0 + bk < Kr && c0 + bc < Nc) ? *reinterpret_cast<const float4*>(Bg + (int64_t)(k0 + bk) * ldb + c0 + bc)
                                          : zero;
  };
  float4 ra = load_a(0), rb[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) rb[j] = load_b(0, j);
  for (int k0 = 0; k0 < Kr; k0 += kSegK) {
    As[lk + 0][lr] = ra.x; As[lk + 1][lr] = ra.y; As[lk + 2][lr] = ra.z; As[lk + 3][lr] = ra.w;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      if (B_T) {
        const int c = lr + 64 * j;
        Bs[lk + 0][c] = rb[j].x; Bs[lk + 1][c] = rb[j].y; Bs[lk + 2][c] = rb[j].z; Bs[lk + 3][c] = rb[j].w;
      } else {
        const int f = tid + 256 * j;
        *reinterpret_cast<float4*>(&Bs[f / (WN / 4)][(f % (WN / 4)) * 4]) = rb[j];
      }
    }
    __syncthreads();
    if (k0 + kSegK < Kr) {                       // next tile's global loads fly during this tile's FMAs
      ra = load_a(k0 + kSegK);
#pragma unroll
      for (int j = 0; j < NB; ++j) rb[j] = load_b(k0 + kSegK, j);
    }
#pragma unroll
    for (int k = 0; k < kSegK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float a4[4] = {av.x, av.y, av.z, av.w};
      float b4[TN];
#pragma unroll
      for (int j = 0; j < NB; ++j) {               // columns tx*4 + 64 j .. +3: conflict-free float4 reads
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4 + 64 * j]);
        b4[4 * j] = bv.x; b4[4 * j + 1] = bv.y; b4[4 * j + 2] = bv.z; b4[4 * j + 3] = bv.w;
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < TN; ++b) acc[a][b] = fmaf(a4[a], b4[b], acc[a][b]);
    }
    __syncthreads();
  }
  float* Cg = C + (int64_t)base * ldc;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int r = r0 + ty * 4 + a;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int c = c0 + tx * 4 + 64 * j;
      if (r < n && c < Nc) {
        float4* dst = reinterpret_cast<float4*>(Cg + (int64_t)r * ldc + c);
        float4 v = make_float4(acc[a][4 * j], acc[a][4 * j + 1], acc[a][4 * j + 2], acc[a][4 * j + 3]);
        if (ACC) {
          const float4 o = *dst;
          v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
        }
        *dst = v;
      }
    }
  }
}

template <bool B_T, bool ACC>
static void launch_seg_gemm(const float* A, int64_t lda, const int* ptr, const float* Bm, int64_t ldb,
                            int64_t b_stride, int Kr, int Nc, float* C, int64_t ldc, int n_cap, int B,
                            cudaStream_t stream) {
  const bool wide = Nc > 64;
  for (int g0 = 0; g0 < B; g0 += 65535) {        // grid.z limit
    const int gb = min(65535, B - g0);
    dim3 grid((unsigned)ceil_div(Nc, wide ? 128 : 64), (unsigned)ceil_div(n_cap, kSegTile), (unsigned)gb);
    if (wide)
      mincut_seg_gemm_kernel<B_T, ACC, 8><<<grid, 256, 0, stream>>>(A, lda, ptr + g0, Bm + (int64_t)g0 * b_stride, ldb,
                                                                   b_stride, Kr, Nc, C, ldc, n_cap);
    else
      mincut_seg_gemm_kernel<B_T, ACC, 4><<<grid, 256, 0, stream>>>(A, lda, ptr + g0, Bm + (int64_t)g0 * b_stride, ldb,
                                                                   b_stride, Kr, Nc, C, ldc, n_cap);
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
  if (phase == 3) {                                // S, degrees, den; A S = one SpMM over the whole batch
    GHSCN_REQUIRE(workspace != nullptr && workspace_bytes >= (size_t)num_nodes * num_clusters * 4);
  }
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
  if (phase == 3)
    return ghscn_spmm(rowptr, col, adj_val, s_soft, K, as_ws, K, nullptr, num_nodes, K, 0, stream_);
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
  if (phase < 0 || phase > 3) return GHSCN_E_INVALID;
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

size_t ghscn_mincut_bwd_split_workspace_bytes(int64_t num_nodes, int64_t num_graphs, int64_t num_clusters) {
  if (num_nodes < 0 || num_graphs < 0 || num_clusters < 0) return 0;
  return ((size_t)4 * num_nodes * num_clusters + (size_t)3 * num_graphs * num_clusters * num_clusters) * 4 + 256;
}

int ghscn_mincut_bwd_split_supported(int64_t num_clusters, int64_t num_feat, int64_t ldx, int64_t lddx,
                                     int32_t max_nodes_per_graph) {
  if (num_clusters <= 0 || num_clusters > kMaxClusters || num_clusters % 4 != 0 || max_nodes_per_graph <= 0) return 0;
  if (num_feat % 4 != 0 || ldx % 4 != 0 || lddx % 4 != 0) return 0;
  return (size_t)max_nodes_per_graph * 4 + kPrepareTileWarps * sizeof(MincutTile) <= kSmemBudget;
}

int ghscn_mincut_bwd_split(const float* s_soft, const float* x, int64_t ldx, const int32_t* ptr,
                           const int32_t* rowptr, const int32_t* col, const float* adj_val, const int32_t* rowptr_t,
                           const int32_t* col_t, const float* adj_val_t, float temp, int64_t num_graphs,
                           int64_t num_nodes, int64_t num_clusters, int64_t num_feat, int32_t max_nodes_per_graph,
                           const float* ss_raw, const float* adj_raw, const float* stats, const float* g_out,
                           const float* g_out_adj, const float* g_losses, float* d_logits, int64_t lddz, float* d_x,
                           int64_t lddx, void* workspace, size_t workspace_bytes, ghscn_stream_t stream_) {
  GHSCN_REQUIRE(num_graphs >= 0 && num_nodes >= 0 && num_clusters > 0 && num_feat >= 0);
  GHSCN_REQUIRE(num_graphs < ((int64_t)1 << 31) && num_nodes < ((int64_t)1 << 31));
  if (!ghscn_mincut_bwd_split_supported(num_clusters, num_feat, ldx, d_x ? lddx : 0, max_nodes_per_graph))
    return GHSCN_E_UNSUPPORTED;
  if (num_graphs == 0) return GHSCN_OK;
  GHSCN_REQUIRE(s_soft && ptr && rowptr && rowptr_t && ss_raw && adj_raw && stats && d_logits);
  GHSCN_REQUIRE(lddz >= num_clusters);
  GHSCN_REQUIRE((g_out == nullptr && d_x == nullptr) || x != nullptr);
  GHSCN_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(g_out) |
                  reinterpret_cast<uintptr_t>(d_x) | reinterpret_cast<uintptr_t>(s_soft) |
                  reinterpret_cast<uintptr_t>(workspace)) & 15) == 0);
  if (workspace == nullptr ||
      workspace_bytes < ghscn_mincut_bwd_split_workspace_bytes(num_nodes, num_graphs, num_clusters))
    return GHSCN_E_WORKSPACE;
  cudaStream_t stream = as_stream(stream_);
  const int K = (int)num_clusters, H = (int)num_feat, n_cap = max_nodes_per_graph, B = (int)num_graphs;
  float* ws = static_cast<float*>(workspace);
  float* stack = ws;                                            // [N][2K] = [A S | A^T S]
  float* dS = ws + (size_t)2 * num_nodes * K;                   // [N][K]
  float* gstack = ws + (size_t)3 * num_nodes * K;               // [B][2K][K] = [Gamma; Gsym]
  int rc = ghscn_spmm(rowptr, col, adj_val, s_soft, K, stack, 2 * K, nullptr, num_nodes, K, 0, stream_);
  if (rc != GHSCN_OK) return rc;
  rc = ghscn_spmm(rowptr_t, col_t, adj_val_t, s_soft, K, stack + K, 2 * K, nullptr, num_nodes, K, 0, stream_);
  if (rc != GHSCN_OK) return rc;
  const size_t prep_shm = (size_t)n_cap * 4 + kPrepareTileWarps * sizeof(MincutTile);
  cudaFuncSetAttribute(mincut_bwd_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget);
  mincut_bwd_prepare_kernel<<<(unsigned)num_graphs, 1024, prep_shm, stream>>>(
      s_soft, ptr, rowptr, adj_val, B, (int)num_nodes, K, n_cap, ss_raw, adj_raw, stats, g_out_adj, g_losses, ws);
  const int64_t gs = (int64_t)2 * K * K;
  // dS += (A S) Gamma^T + (A^T S) Gamma + S Gsym
  launch_seg_gemm<true, true>(stack, 2 * K, ptr, gstack, K, gs, K, K, dS, K, n_cap, B, stream);
  launch_seg_gemm<false, true>(stack + K, 2 * K, ptr, gstack, K, gs, K, K, dS, K, n_cap, B, stream);
  launch_seg_gemm<false, true>(s_soft, K, ptr, gstack + (size_t)K * K, K, gs, K, K, dS, K, n_cap, B, stream);
  if (g_out != nullptr)                                          // dS += x g_out^T
    launch_seg_gemm<true, true>(x, ldx, ptr, g_out, H, (int64_t)K * H, H, K, dS, K, n_cap, B, stream);
  mincut_bwd_softmax_kernel<<<(unsigned)num_graphs, 256, 0, stream>>>(s_soft, ptr, temp, B, (int)num_nodes, K, H,
                                                                     n_cap, g_out != nullptr, d_logits, lddz, d_x,
                                                                     lddx, ws);
  if (g_out != nullptr && d_x != nullptr)                        // dX = S g_out
    launch_seg_gemm<false, false>(s_soft, K, ptr, g_out, H, (int64_t)K * H, K, H, d_x, lddx, n_cap, B, stream);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

}  // extern "C"
