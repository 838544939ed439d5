// K6: fused MinCUT pool, one CTA per graph, forward and analytic backward.
// Replaces to_dense_adj + dense_mincut_pool (model/hscn.py:61-63).
//
// The [n,n] dense adjacency is never materialised: the graph's CSR slice is consumed directly
// ((A S)[r] = sum_{e: row_e = r} val_e S[col_e]).  Per graph the CTA stages S = softmax(logits)
// and A S in shared memory (falls back to an HBM workspace when n*K does not fit), and produces
// S^T X, S^T A S, S^T S, both traces and both losses without any other round trip through HBM.
// HBM-bound at the reference's sizes (K <= 32): algorithmic bytes per graph
//   4nK (logits) + 4nH (X, only if `out` is requested) + 4(n+1) + 4 nnz  ->  4nK (S) + 4KH + 8K^2 + 32.
#include <algorithm>
#include <cooperative_groups.h>
#include <math.h>

#include <cstdlib>

#include "common.cuh"

namespace ghscn {

// few graphs: one big CTA per graph keeps an SM busy; many graphs: several small CTAs per SM
static inline int mincut_threads(int64_t num_graphs, int64_t num_clusters) {
  static const int forced = [] {                       // tuning experiments: GHSCN_MINCUT_THREADS=256|512|1024
    const char* e = getenv("GHSCN_MINCUT_THREADS");
    const int v = e ? atoi(e) : 0;
    return (v == 128 || v == 256 || v == 512 || v == 1024) ? v : 0;
  }();
  if (forced) return forced;
  // K <= 12 (the row-wise code paths of the <= 512-thread regime, late round 2), losses only, K = 10, forward /
  // forward + backward: B = 128: 16.6 / 38.6 us (256 threads), 13.0 / 30.3 (512), 16.5 / 38.8 (1 024);
  // B = 300: 19.1 / 40.7 (256), 22.0 / 45.7 (512); B = 600: 24.2 / 55.8 (256), 32.3 / 64.1 (512)
  if (num_clusters <= 12) return num_graphs >= 2 * kNumSMs ? 256 : 512;
  return num_graphs >= 4 * kNumSMs ? 256 : (num_graphs >= 2 * kNumSMs ? 512 : 1024);
}
// GHSCN_MINCUT_STAGE=1 stages the graph's logits tile (one TMA bulk copy) and CSR slice in shared memory before the
// fused forward's phases.  Off by default: measured 18.3 vs 17.0 us (B = 128, K = 10), 90.6 vs 89.5 us (B = 1024,
// K = 16) -- the kernel is bound by its instruction count and barrier phases, not by the global round trips.
static inline bool mincut_stage_enabled() {
  static const bool on = [] {
    const char* e = getenv("GHSCN_MINCUT_STAGE");
    return e && e[0] == '1';
  }();
  return on;
}
// GHSCN_MINCUT_STREAM_X=0 keeps S^T X / x g_out^T / S g_out inside the per-graph kernels (A/B measurements)
static inline bool mincut_stream_x_enabled() {
  static const bool on = [] {
    const char* e = getenv("GHSCN_MINCUT_STREAM_X");
    return !(e && e[0] == '0');
  }();
  return on;
}
constexpr int kMaxClusters = 128;
constexpr int kStatsStride = 8;  // per graph: num, den, ||SS||_F, ||R||_F (= ortho_g), mc_g, -, -, -
constexpr size_t kSmemBudget = 200 * 1024;

// floor(e / K) for 0 <= e < 2^20 and 1 <= K <= 128 by one multiply-high: kinv = ceil(2^32 / K) (exact while
// e * (kinv * K - 2^32) < 2^32; a hardware integer division is ~20 instructions -- 7 - 12 % of the backward's warp
// instructions at B = 1 024 in ncu's source counters)
__device__ __forceinline__ unsigned div_magic(int K) { return (unsigned)((0x100000000ull + (unsigned)K - 1) / (unsigned)K); }
__device__ __forceinline__ int fast_div(int e, unsigned kinv) { return kinv ? (int)__umulhi((unsigned)e, kinv) : e; }   // K = 1: 2^32 wraps to 0

__device__ __forceinline__ float adj_value(const float* __restrict__ adj_val, int s) {
  return adj_val ? adj_val[s] : 1.0f;
}

// ---- mbarrier / TMA bulk-copy helpers for the tile staging of the forward kernel --------------------------------------
__device__ __forceinline__ uint32_t mc_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mc_bar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mc_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mc_bar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}

// (A S) from a CSR slice staged in shared memory: rp = the graph's n + 1 row pointers, cs / vs = its slots (global slot
// e0 + s), same slot order and roundings as csr_times_s.
__device__ __forceinline__ void csr_times_s_staged(const int* __restrict__ rp, const int* __restrict__ cs,
                                                   const float* __restrict__ vs, int e0, const float* __restrict__ S,
                                                   int base, int n, int K, float* __restrict__ buf) {
  for (int e = threadIdx.x; e < n * K; e += blockDim.x) {
    const int i = e / K, k = e - i * K;
    float acc = 0.f;
    const int beg = rp[i] - e0, end = rp[i + 1] - e0;
    for (int s = beg; s < end; ++s) {
      const int c = cs[s] - base;
      if (c >= 0 && c < n) acc += (vs ? vs[s] : 1.0f) * S[c * K + k];
    }
    buf[e] = acc;
  }
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

// The same product with one THREAD per node row (many-graph regime, K <= 32): the row's slots are read once and
// applied to all K columns held in registers -- 3 + 2 KT instructions per slot instead of ~8 per slot and column
// (the element-wise walk above was 38 % of the backward's and 25 % of the forward's warp instructions at B = 1 024,
// K = 10).  Per element the slots are still added in slot order.  TRACES: also accumulates the thread's share of
// num = sum S (.) A S and den = sum deg (.) S^2 (forward).
template <int KT, bool TRACES>
__device__ __forceinline__ void csr_times_s_rows_kt(const int* __restrict__ rowptr, const int* __restrict__ col,
                                                    const float* __restrict__ adj_val, const float* __restrict__ S,
                                                    int base, int n, int K, float* __restrict__ buf,
                                                    const float* __restrict__ deg, float& pnum, float& pden) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float acc[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) acc[k] = 0.f;
    const int beg = rowptr[base + i], end = rowptr[base + i + 1];
    for (int s = beg; s < end; ++s) {
      const int c = col[s] - base;
      if (c < 0 || c >= n) continue;
      const float w = adj_value(adj_val, s);
      const float* sr = S + c * K;
#pragma unroll
      for (int k = 0; k < KT; ++k)
        if (k < K) acc[k] += w * sr[k];
    }
    float* br = buf + (size_t)i * K;
    const float* si = S + i * K;
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int k = 0; k < KT; ++k)
      if (k < K) {
        br[k] = acc[k];
        if (TRACES) { const float sv = si[k]; a += sv * acc[k]; b += sv * sv; }
      }
    if (TRACES) { pnum += a; pden += deg[i] * b; }
  }
}
template <bool TRACES>
__device__ __forceinline__ void csr_times_s_rows(const int* __restrict__ rowptr, const int* __restrict__ col,
                                                 const float* __restrict__ adj_val, const float* __restrict__ S,
                                                 int base, int n, int K, float* __restrict__ buf,
                                                 const float* __restrict__ deg, float& pnum, float& pden) {
  if (K <= 4) csr_times_s_rows_kt<4, TRACES>(rowptr, col, adj_val, S, base, n, K, buf, deg, pnum, pden);
  else if (K <= 8) csr_times_s_rows_kt<8, TRACES>(rowptr, col, adj_val, S, base, n, K, buf, deg, pnum, pden);
  else if (K <= 12) csr_times_s_rows_kt<12, TRACES>(rowptr, col, adj_val, S, base, n, K, buf, deg, pnum, pden);
  else if (K <= 16) csr_times_s_rows_kt<16, TRACES>(rowptr, col, adj_val, S, base, n, K, buf, deg, pnum, pden);
  else csr_times_s_rows_kt<32, TRACES>(rowptr, col, adj_val, S, base, n, K, buf, deg, pnum, pden);
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
    float* __restrict__ adj_raw, float* __restrict__ stats, float* __restrict__ as_ws, int phase, int e_cap) {
  // phase 0: everything.  phase 1: S, A S, the two traces (num, den -> stats) and nothing else: the K x K and K x H
  // contractions are then done by ghscn_gemm3x_tn_segmented on the tensor cores (dense-bound for K >= 64), and
  // phase 2 finishes from ss_raw / adj_raw: norms, orthogonality loss, normalised coarse adjacency.
  extern __shared__ float smem[];
  __shared__ float red[32];
  __shared__ float dk[kMaxClusters];
  __shared__ __align__(8) unsigned long long tile_bar;
  __shared__ float kpart[1024];
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

  // Staging (fused path with room for it, e_cap > 0): the graph's logits tile -> the S tile (ONE TMA bulk copy when
  // the rows are contiguous and 16-byte aligned, else coalesced loads with every load of the CTA in flight at once),
  // its n + 1 row pointers and -- one dependent round trip later -- its column / value slots.  Every later phase
  // then runs out of shared memory: the round-1 kernel walked rowptr -> values, logits, rowptr -> col -> S as
  // separate dependent global round trips per thread.  Graphs with more than e_cap slots keep the global CSR reads.
  int* rp = nullptr;
  int* cs = nullptr;
  float* vs = nullptr;
  float* kk = nullptr;                               // [2][K*K] S^T S and S^T A S (K <= 32)
  bool tile_staged = false, csr_staged = false;
  int e0 = 0;
  if (SMEM && phase == 0 && e_cap > 0) {
    rp = reinterpret_cast<int*>(deg + n_cap);
    cs = rp + ((n_cap + 4) & ~3);
    vs = reinterpret_cast<float*>(cs + e_cap);
    if (K <= 32) kk = vs + e_cap;
    const float* zt = logits + (int64_t)base * ldz;
    const uint32_t bytes = (uint32_t)n * K * 4u;
    const bool bulk = ldz == K && n > 0 && ((reinterpret_cast<uintptr_t>(zt) | bytes) & 15u) == 0;   // CTA-uniform
    const uint32_t bar = mc_smem_addr(&tile_bar);
    if (bulk) {
      if (tid == 0) mc_bar_init(bar, 1);
      __syncthreads();
      if (tid == 0) mc_bulk_load(mc_smem_addr(S), zt, bytes, bar);
    } else {
      for (int e = tid; e < n * K; e += blockDim.x) S[e] = zt[(int64_t)(e / K) * ldz + e % K];
    }
    for (int i = tid; i <= n; i += blockDim.x) rp[i] = rowptr[base + i];
    if (bulk) mc_bar_wait(bar, 0);
    __syncthreads();
    tile_staged = true;
    e0 = rp[0];
    const int ne = rp[n] - e0;
    csr_staged = ne >= 0 && ne <= e_cap;             // CTA-uniform
    if (csr_staged) {
      for (int s2 = tid; s2 < ne; s2 += blockDim.x) {
        cs[s2] = col[e0 + s2];
        if (adj_val) vs[s2] = adj_val[e0 + s2];
      }
      if (!adj_val) vs = nullptr;
    }
    __syncthreads();
  }

  // lane-group geometry for K <= 32: G = 4 / 8 / 16 / 32 lanes per node row, 32 / G rows per warp pass
  const int gsh = K <= 4 ? 2 : (K <= 8 ? 3 : (K <= 16 ? 4 : 5));
  const int G = 1 << gsh, sub = lane >> gsh, kl = lane & (G - 1), rpw = 32 >> gsh;
  float num = 0.f, den = 0.f;
  if (phase != 2) {
  // A. S = softmax(logits / temp).  Small K: one thread per node row (all rows of the graph in flight at once);
  //    large K: one warp per row.  Row degrees (row sums of A) by one thread per row.
  for (int i = tid; i < n; i += blockDim.x) {
    float d;
    if (csr_staged) {
      const int beg = rp[i] - e0, end = rp[i + 1] - e0;
      d = (float)(end - beg);
      if (vs) {
        d = 0.f;
        for (int s = beg; s < end; ++s) d += vs[s];
      }
    } else {
      const int beg = rowptr[base + i], end = rowptr[base + i + 1];
      d = (float)(end - beg);
      if (adj_val) {
        d = 0.f;
        for (int s = beg; s < end; ++s) d += adj_val[s];
      }
    }
    deg[i] = d;
  }
  if (K <= 32 && blockDim.x <= 512 && 8 * K <= 5 * G) {
    // Many graphs (256 / 512 threads per graph, several CTAs per SM): the kernel is bound by instruction issue.  Where
    // the lane groups below would leave more than 3/8 of their lanes idle (K = 5, 9-10, 17-20), one thread per node
    // row with the row in registers needs fewer warp instructions (ncu source counters, B = 1 024, K = 10: 10 k of the
    // 38 k warp instructions per graph were this softmax; forward 65.3 -> 55.3 us).  Measured slower at K = 4 and
    // K = 16, where the groups are full.
    for (int i = tid; i < n; i += blockDim.x) {
      const float* zr = tile_staged ? S + i * K : logits + (int64_t)(base + i) * ldz;
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
          const float pv = __fdiv_rn(z[k], sum);
          S[i * K + k] = pv;
          if (SMEM) Sg[i * K + k] = pv;
        }
    }
  } else if (K <= 32) {
    // Few graphs (1 024 threads per graph, one CTA per SM: latency-bound): G lanes per node row, max / sum by shuffles
    // inside the lane group, every warp busy.  (One thread per row ran ~800 dependent instructions on 5 of the CTA's
    // 32 warps while the others waited at the barrier: 19.5 vs 17.0 us at B = 128.)
    for (int i0 = wid * rpw; i0 < n; i0 += nwarps * rpw) {
      const int i = i0 + sub;
      const bool on = i < n && kl < K;
      float z = -INFINITY;
      if (on) {
        z = tile_staged ? S[i * K + kl] : logits[(int64_t)(base + i) * ldz + kl];
        if (temp != 1.0f) z = __fdiv_rn(z, temp);
      }
      float m = z;
      for (int o = G >> 1; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFullMask, m, o));
      float pz = on ? expf(z - m) : 0.f;
      float sum = pz;
      for (int o = G >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(kFullMask, sum, o);
      if (on) {
        pz = __fdiv_rn(pz, sum);
        S[i * K + kl] = pz;
        if (SMEM) Sg[i * K + kl] = pz;
      }
    }
  } else {
    for (int i = wid; i < n; i += nwarps) {
      const float* zr = tile_staged ? S + i * K : logits + (int64_t)(base + i) * ldz;
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

  // B. AS = A S (phase 3: the caller runs the K2 SpMM over the whole batch instead) and C. the two traces
  float pnum = 0.f, pden = 0.f;
  if (K <= 12 && phase != 3 && blockDim.x <= 512 && !csr_staged) {
    // many graphs: thread per row (B = 1 024 forward: K = 4 23.3 -> 19.4 us, K = 10 52.4 -> 36.3; K = 16 slower, 87.7 -> 95.3)
    csr_times_s_rows<true>(rowptr, col, adj_val, S, base, n, K, AS, deg, pnum, pden);
  } else if (K <= 32 && phase != 3) {
    // lane group per row again: the row's slots are walked once by its K lanes (no e / K divisions), and the trace
    // terms are taken while S and A S of the element are in registers
    for (int i0 = wid * rpw; i0 < n; i0 += nwarps * rpw) {
      const int i = i0 + sub;
      if (i < n && kl < K) {
        int beg, end;
        const int* cp;
        const float* vp;
        if (csr_staged) { beg = rp[i] - e0; end = rp[i + 1] - e0; cp = cs; vp = vs; }
        else { beg = rowptr[base + i]; end = rowptr[base + i + 1]; cp = col; vp = adj_val; }
        float acc = 0.f;
        for (int s2 = beg; s2 < end; ++s2) {
          const int c = cp[s2] - base;
          if (c >= 0 && c < n) acc += (vp ? vp[s2] : 1.0f) * S[c * K + kl];
        }
        AS[i * K + kl] = acc;
        const float sv = S[i * K + kl];
        pnum += sv * acc;
        pden += deg[i] * sv * sv;
      }
    }
  } else {
    if (phase != 3) {
      if (csr_staged) csr_times_s_staged(rp, cs, vs, e0, S, base, n, K, AS);
      else csr_times_s(rowptr, col, adj_val, S, base, n, K, AS);
      __syncthreads();
    }
#pragma unroll 4
    for (int e = tid; e < n * K; e += blockDim.x) {
      const float sv = S[e];
      if (phase != 3) pnum += sv * AS[e];
      pden += deg[e / K] * sv * sv;
    }
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

  float* ssg_g = ss_raw + (int64_t)g * K * K;
  float* oag_g = adj_raw + (int64_t)g * K * K;
  // the K x K blocks stay in shared memory for the norms below when there is room (no write -> barrier -> re-read
  // through L2); the global copies are for the backward
  float* ssg = kk ? kk : ssg_g;
  float* oag = kk ? kk + K * K : oag_g;
  if (phase == 0) {
    const int items = 2 * K * K;
    if (K > 4 && K <= 12 && blockDim.x <= 512 && 8 * K <= (int)blockDim.x) {
      // [S^T S | S^T A S] for 4 < K <= 12, many graphs (issue-bound regime): a group of SL consecutive lanes per output ROW (k of either matrix), each lane
      // walking every SL-th node with the row's K accumulators in registers: 1 + K shared loads for K FMAs per node
      // (the one-thread-per-output-element form below spends 2 loads per FMA: 9.6 k of 38 k warp instructions per
      // graph at B = 1 024, K = 10), slices combined by shuffles inside the group
      int SL = 32;
      while (2 * K * SL > (int)blockDim.x) SL >>= 1;
      const int row_id = tid / SL, sl = tid - row_id * SL;
      const bool act = row_id < 2 * K;
      const bool second = row_id >= K;
      const int k = row_id - (second ? K : 0);
      const float* Bm = second ? AS : S;
      float acc[12];
#pragma unroll
      for (int l = 0; l < 12; ++l) acc[l] = 0.f;
      if (act) {
        for (int i = sl; i < n; i += SL) {
          const float a = S[i * K + k];
          const float* br = Bm + i * K;
#pragma unroll
          for (int l = 0; l < 12; ++l)
            if (l < K) acc[l] = fmaf(a, br[l], acc[l]);
        }
      }
#pragma unroll
      for (int l = 0; l < 12; ++l)
        if (l < K)
          for (int o = SL >> 1; o > 0; o >>= 1) acc[l] += __shfl_xor_sync(kFullMask, acc[l], o);
      if (act && sl == 0) {
        float* dst = (second ? oag : ssg) + k * K;
#pragma unroll
        for (int l = 0; l < 12; ++l)
          if (l < K) dst[l] = acc[l];
      }
    } else if (items <= (int)blockDim.x) {
      // small K: one thread per (output element, node slice) of [S^T S | S^T A S], slices combined in order
      const int parts = blockDim.x / items;
      if (tid < parts * items) {
        const int part = tid / items, item = tid - part * items;
        const bool second = item >= K * K;
        const int pe = item - (second ? K * K : 0), k = pe / K, l = pe - k * K;
        const float* Bm = second ? AS : S;
        float a = 0.f;
        for (int i = part; i < n; i += parts) a = fmaf(S[i * K + k], Bm[i * K + l], a);
        kpart[tid] = a;
      }
      __syncthreads();
      if (tid < items) {
        float t = 0.f;
        for (int q = 0; q < parts; ++q) t += kpart[q * items + tid];
        if (tid < K * K) ssg[tid] = t;
        else oag[tid - K * K] = t;
      }
    } else {
      atb_tiled<4, 4>(S, K, K, S, K, K, n, ssg, K);    // S^T S
      atb_tiled<4, 4>(S, K, K, AS, K, K, n, oag, K);   // S^T (A S)
    }
    if (out != nullptr)                               // S^T X, X streamed from global memory
      atb_tiled<8, 4>(S, K, K, x + (int64_t)base * ldx, ldx, H, n, out + (int64_t)g * K * H, H);
    __syncthreads();  // ssg / oag visible to the whole CTA
    if (kk)
      for (int p2 = tid; p2 < K * K; p2 += blockDim.x) { ssg_g[p2] = ssg[p2]; oag_g[p2] = oag[p2]; }
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
    int64_t lddz, float* __restrict__ d_x, int64_t lddx, float* __restrict__ ws, const float* __restrict__ xg) {
  // xg != nullptr: x g_out^T ([N, K]) and d_x = S g_out were already produced by mincut_pool_x_bwd_kernel
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
  if (K <= 32 && blockDim.x <= 512) {               // many graphs: thread per row
    float unused_a = 0.f, unused_b = 0.f;
    csr_times_s_rows<false>(rowptr, col, adj_val, Sr, base, n, K, AS, nullptr, unused_a, unused_b);
    csr_times_s_rows<false>(rowptr_t, col_t, adj_val_t, Sr, base, n, K, ATS, nullptr, unused_a, unused_b);
  } else {
    csr_times_s(rowptr, col, adj_val, Sr, base, n, K, AS);
    csr_times_s(rowptr_t, col_t, adj_val_t, Sr, base, n, K, ATS);
  }

  mincut_bwd_coefficients(g, B, K, ss_raw, adj_raw, stats, g_out_adj, g_losses, Gsym, Gam, sc, tiles, 4);
  const float* st = stats + (int64_t)g * kStatsStride;
  const float num = st[0], den = st[1];
  const float gmc = g_losses ? __fdiv_rn(g_losses[0], (float)B) : 0.f;
  const float cden = gmc * num / (den * den);
  const bool diag_only = (g_out_adj == nullptr);
  const float gdiag = -gmc / den;
  const unsigned kinv = div_magic(K);
  const float* gog = g_out ? g_out + (int64_t)g * K * H : nullptr;
  // dS = [mincut trace / out_adj chain] + [den term] + [ortho] + [out term], as tiled small GEMMs
  if (diag_only) {
    for (int e = tid; e < n * K; e += blockDim.x)
      dS[e] = gdiag * (AS[e] + ATS[e]) + cden * 2.f * deg[fast_div(e, kinv)] * Sr[e];
  } else {
    abt_tiled<4, 4>(AS, K, n, Gam, K, K, K, 1.f, dS, K, false);          // AS  Gamma^T
    __syncthreads();
    ab_tiled<4, 4>(ATS, K, n, Gam, K, K, K, 1.f, dS, K, true);           // A^T S Gamma
    __syncthreads();
    for (int e = tid; e < n * K; e += blockDim.x) dS[e] += cden * 2.f * deg[fast_div(e, kinv)] * Sr[e];
  }
  __syncthreads();
  ab_tiled<4, 4>(Sr, K, n, Gsym, K, K, K, 1.f, dS, K, true);             // S (G' + G'^T) go
  if (gog) {
    __syncthreads();
    if (xg) {
      const float* xgg = xg + (int64_t)base * K;
      for (int e = tid; e < n * K; e += blockDim.x) dS[e] += xgg[e];
    } else {
      abt_tiled<4, 4>(x + (int64_t)base * ldx, ldx, n, gog, H, K, H, 1.f, dS, K, true);   // x g_out^T
    }
  }
  __syncthreads();
  // softmax backward
  if (K <= 32) {
    // small K: row dot products by one thread per row (into the degree array, which is no longer needed), then one
    // thread per element with coalesced stores.  (A warp per row kept 10 of 32 lanes busy at K = 10 and was 36 % of
    // the kernel's warp instructions.)
    for (int i = tid; i < n; i += blockDim.x) {
      float dot = 0.f;
      for (int k = 0; k < K; ++k) dot = fmaf(dS[i * K + k], Sr[i * K + k], dot);
      deg[i] = dot;
    }
    __syncthreads();
    for (int e = tid; e < n * K; e += blockDim.x) {
      const int i = fast_div(e, kinv);
      float dz = Sr[e] * (dS[e] - deg[i]);
      if (temp != 1.0f) dz = dz / temp;
      d_logits[(int64_t)(base + i) * lddz + (e - i * K)] = dz;
    }
  } else {
    for (int i = wid; i < n; i += nwarps) {      // one warp per node row
      float dot = 0.f;
      for (int k = lane; k < K; k += 32) dot += dS[i * K + k] * Sr[i * K + k];
      dot = warp_sum(dot);
      for (int k = lane; k < K; k += 32) {
        float dz = Sr[i * K + k] * (dS[i * K + k] - dot);
        if (temp != 1.0f) dz = dz / temp;
        d_logits[(int64_t)(base + i) * lddz + k] = dz;
      }
    }
  }
  if (d_x != nullptr && !(xg && gog)) {
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

// ---- pooled features for K <= 32: streaming kernels over the whole batch --------------------------------------------
// S^T X, x g_out^T and S g_out read / write the graph's [n, H] feature rows exactly once and are bandwidth work: one
// CTA per graph walking them alone is held to one SM's share of HBM (round 1 / early round 2: 46 us of the 65 us
// pooled forward at the bench shape for 23 MB).  These kernels spread the rows of every graph over the whole device.

// S rows are short (K floats, 4- / 8- / 16-byte aligned depending on K): widest aligned load that K allows
template <int KT>
__device__ __forceinline__ void load_s_row(const float* __restrict__ sr, int K, float (&sv)[KT]) {
  if ((K & 3) == 0) {
#pragma unroll
    for (int k = 0; k < KT; k += 4)
      if (k < K) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(sr + k));
        sv[k] = v.x; sv[k + 1] = v.y; sv[k + 2] = v.z; sv[k + 3] = v.w;
      }
  } else if ((K & 1) == 0) {
#pragma unroll
    for (int k = 0; k < KT; k += 2)
      if (k < K) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(sr + k));
        sv[k] = v.x; sv[k + 1] = v.y;
      }
  } else {
#pragma unroll
    for (int k = 0; k < KT; ++k)
      if (k < K) sv[k] = __ldg(sr + k);
  }
}

// out[g, k, c] = sum_i S[i, k] X[i, c].  One CTA per graph over the full feature width.  The graph's feature rows
// are contiguous ([n, H], ldx == H), so they stream through a ring of kPoolStages shared-memory stages filled by ONE
// TMA bulk copy per chunk of rows (12 - 16 KB): every byte of the next stages is in flight while the current chunk is
// consumed, at no register cost.  (Measured alternatives: register-staged rows, 4 in flight per thread: 0.6 - 0.8 TB/s;
// column-tiled CTAs with one 400-byte bulk copy per row: 0.7 TB/s -- the copy engine is bound by the number of bulk
// requests, not by their bytes.)  The graph's S tile sits in shared memory too (one coalesced load, overlapped with
// the ring's prologue).  Thread = (float4 column group, row slice) with a K x 4 accumulator tile; a chunk is 4 rows
// per slice.  Slices are combined by a fixed-order tree through shared memory (deterministic).
// KT = K rounded up to a multiple of 4.
constexpr int kPoolMaxStages = 8;
static inline int pool_stages() {                  // GHSCN_POOL_STAGES=2..8: ring depth (A/B measurements)
  static const int v = [] {
    const char* e = getenv("GHSCN_POOL_STAGES");
    const int q = e ? atoi(e) : 0;
    return (q >= 2 && q <= kPoolMaxStages) ? q : 4;
  }();
  return v;
}

template <int KT>
__global__ void __launch_bounds__(256, (KT <= 16 ? 2 : 1))
    mincut_pool_x_kernel(const float* __restrict__ s_soft, const float* __restrict__ x, const int* __restrict__ ptr,
                         int K, int H, int s_cap, int s_off, int segs, int kPoolStages,
                         float* __restrict__ out) {
  // s_off: float4 offset of the S tile = max(ring, reduction tiles).  segs > 1: the graph's rows are cut into `segs`
  // equal ranges, one per CTA of a thread-block cluster (gridDim.x = cluster size = segs); the partial tiles are
  // combined over distributed shared memory in rank order.
  extern __shared__ __align__(128) float4 pool_smem[];
  __shared__ __align__(8) unsigned long long full[kPoolMaxStages];
  constexpr int KQ = KT / 4;
  const int g = blockIdx.y, rank = blockIdx.x;
  int base = ptr[g], n = ptr[g + 1] - base;
  if (segs > 1) {
    const int per = ceil_div(max(n, 0), segs), lo = min(n, rank * per);
    n = max(0, min(n, lo + per) - lo);
    base += lo;
  }
  const bool bad = n > s_cap;                                      // stale max-nodes hint: poison the output tile
  if (bad) n = 0;
  const int H4 = H >> 2;                                           // <= blockDim.x
  const int slices = blockDim.x / H4, rows = 4 * slices;           // rows per chunk
  const int grp = threadIdx.x % H4, slice = threadIdx.x / H4;
  const bool on = slice < slices;
  float4* ring = pool_smem;                                        // [kPoolStages][rows][H4]
  float4* s_tile = pool_smem + s_off;                              // [s_cap][KT / 4], columns K .. KT - 1 are zero
  const int nchunks = ceil_div(n, rows);
  const float* xt = x + (int64_t)base * H;

  auto issue = [&](int c) {                                        // thread 0 fills stage c % kPoolStages with chunk c
    const int st = c % kPoolStages, r0 = c * rows, rc = min(rows, n - r0);
    mc_bulk_load(mc_smem_addr(ring + (size_t)st * rows * H4), xt + (int64_t)r0 * H, (uint32_t)rc * H * 4u,
                 mc_smem_addr(&full[st]));
  };

  if (threadIdx.x == 0) {
    for (int st = 0; st < kPoolStages; ++st) mc_bar_init(mc_smem_addr(&full[st]), 1);
    for (int c = 0; c < min(nchunks, kPoolStages); ++c) issue(c);
  }
  {
    const float* sp = s_soft + (int64_t)base * K;
    float* st = reinterpret_cast<float*>(s_tile);
    const int total = n * KT;                                      // four loads in flight per thread and round
    for (int e0 = threadIdx.x; e0 < total; e0 += 4 * blockDim.x) {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * blockDim.x, i = e / KT, k = e - i * KT;
        v[u] = (e < total && k < K) ? __ldg(sp + i * K + k) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * blockDim.x;
        if (e < total) st[e] = v[u];
      }
    }
  }
  __syncthreads();

  float4 acc[KT];
#pragma unroll
  for (int k = 0; k < KT; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = 0; c < nchunks; ++c) {
    const int st = c % kPoolStages, r0 = c * rows, rc = min(rows, n - r0);
    mc_bar_wait(mc_smem_addr(&full[st]), (uint32_t)(c / kPoolStages) & 1u);
    if (on) {
      const float4* src = ring + (size_t)st * rows * H4 + grp;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = slice + u * slices;
        if (r < rc) {
          const float4 q = src[(size_t)r * H4];
          const float4* srow = s_tile + (size_t)(r0 + r) * KQ;
#pragma unroll
          for (int j = 0; j < KQ; ++j) {
            const float4 s4 = srow[j];
            const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              float4& a = acc[4 * j + t];
              a.x = fmaf(sv[t], q.x, a.x); a.y = fmaf(sv[t], q.y, a.y);
              a.z = fmaf(sv[t], q.z, a.z); a.w = fmaf(sv[t], q.w, a.w);
            }
          }
        }
      }
    }
    __syncthreads();                                               // the stage is free again
    if (threadIdx.x == 0 && c + kPoolStages < nchunks) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(c + kPoolStages);
    }
  }
  float4* part = ring;                                             // [ceil(slices / 2)][KT][H4], the ring is drained
  for (int cnt = slices; cnt > 1;) {               // slices [half, cnt) hand their tiles to slices [0, cnt - half)
    const int half = (cnt + 1) >> 1;
    if (on && slice >= half && slice < cnt) {
#pragma unroll
      for (int k = 0; k < KT; ++k) part[((slice - half) * KT + k) * H4 + grp] = acc[k];
    }
    __syncthreads();
    if (on && slice + half < cnt) {
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        const float4 v = part[(slice * KT + k) * H4 + grp];
        acc[k].x += v.x; acc[k].y += v.y; acc[k].z += v.z; acc[k].w += v.w;
      }
    }
    __syncthreads();
    cnt = half;
  }
  if (bad) {
#pragma unroll
    for (int k = 0; k < KT; ++k) acc[k] = make_float4(NAN, NAN, NAN, NAN);
  }
  if (segs == 1) {
    if (on && slice == 0) {
      float* og = out + (int64_t)g * K * H + grp * 4;
#pragma unroll
      for (int k = 0; k < KT; ++k)
        if (k < K) *reinterpret_cast<float4*>(og + (int64_t)k * H) = acc[k];
    }
  } else {
    namespace cgr = cooperative_groups;
    cgr::cluster_group cluster = cgr::this_cluster();
    float4* cpart = ring;                                          // [K][H4]; the tree above is done with `part`
    if (on && slice == 0) {
#pragma unroll
      for (int k = 0; k < KT; ++k)
        if (k < K) cpart[k * H4 + grp] = acc[k];
    }
    cluster.sync();
    float4* og = reinterpret_cast<float4*>(out + (int64_t)g * K * H);
    for (int e = rank * blockDim.x + threadIdx.x; e < K * H4; e += segs * blockDim.x) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int q = 0; q < segs; ++q) {
        const float4 v = cluster.map_shared_rank(cpart, q)[e];
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
      og[e] = a;
    }
    cluster.sync();                                                // nobody leaves while a peer still reads its tile
  }
}

// Sums of N (= 4, 8 or 16) per-lane values over the 32 lanes of a warp, "transposed": every halving step exchanges
// half of the values with the partner lane (the lanes with the step's bit clear keep the lower half), so N values cost
// N shuffles instead of 5 N; lane l ends up with the total of value index rev(l): bit 4 of l selects the upper half
// of 16, bit 3 the upper half of those 8, ... -- `warp_transpose_index` gives that index (-1: no value on this lane).
template <int N>
__device__ __forceinline__ float warp_transpose_sum(float (&v)[N], int lane) {
  static_assert(N == 4 || N == 8 || N == 16, "N must be 4, 8 or 16");
  int off = 16;
#pragma unroll
  for (int half = N / 2; half >= 1; half >>= 1, off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      const float send = up ? v[j] : v[j + half];
      const float keep = up ? v[j + half] : v[j];
      v[j] = keep + __shfl_xor_sync(kFullMask, send, off);
    }
  }
  float r = v[0];
  for (; off >= 1; off >>= 1) r += __shfl_xor_sync(kFullMask, r, off);   // remaining bits hold the same index
  return r;
}
template <int N>
__device__ __forceinline__ int warp_transpose_index(int lane) {
  int idx = 0, off = 16;
#pragma unroll
  for (int half = N / 2; half >= 1; half >>= 1, off >>= 1)
    if (lane & off) idx += half;
  return idx;
}

// Backward of the pooled features in one pass over the rows: xg[i, :] = x[i, :] g_out[g]^T (the term of dS) and, when
// WANT_DX, d_x[i, :] = S[i, :] g_out[g].  CTA = (chunk of `rows_per_cta` rows, graph): g_out[g] ([K, H], <= 64 KB) and
// the chunk's feature rows land in shared memory by TMA bulk copies issued up front (one each when the rows are
// contiguous, else one per row), so everything the CTA reads is in flight at once; a warp then owns two rows at a
// time: its lanes walk the float4 column groups, d_x is written once (coalesced 128-bit stores), the K partial dot
// products are combined by shuffles.
template <int KT, bool WANT_DX>
__global__ void __launch_bounds__(256, (KT <= 16 ? 2 : 1)) mincut_pool_x_bwd_kernel(const float* __restrict__ s_soft,
                                                                const float* __restrict__ x, int64_t ldx,
                                                                const int* __restrict__ ptr,
                                                                const float* __restrict__ g_out, int K, int H,
                                                                int rows_per_cta, int n_cap, float* __restrict__ xg,
                                                                float* __restrict__ d_x, int64_t lddx) {
  extern __shared__ __align__(128) float4 pool_smem[];
  __shared__ __align__(8) unsigned long long bar_mem;
  const int g = blockIdx.y;
  const int base = ptr[g], n = ptr[g + 1] - base;
  const int row0 = blockIdx.x * rows_per_cta;
  if (n > n_cap) {                                 // stale max-nodes hint (the grid does not cover the graph): poison
    if (WANT_DX && blockIdx.x == 0 && threadIdx.x == 0) d_x[(int64_t)base * lddx] = NAN;
    return;
  }
  if (row0 >= n) return;
  const int rc = min(n - row0, rows_per_cta);
  const int H4 = H >> 2;
  float4* gs = pool_smem;                          // [KT][H4], rows K .. KT - 1 zero
  float4* xs = pool_smem + (size_t)KT * H4;        // [rows_per_cta][H4]
  const uint32_t bar = mc_smem_addr(&bar_mem);
  const float* xc = x + (int64_t)(base + row0) * ldx;
  if (threadIdx.x == 0) mc_bar_init(bar, 1);
  __syncthreads();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const uint32_t gbytes = (uint32_t)K * H * 4u, rbytes = (uint32_t)H * 4u;
    if (lane == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                   "r"(gbytes + rbytes * (uint32_t)rc)
                   : "memory");
    __syncwarp();
    if (lane == 0)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       mc_smem_addr(gs)),
                   "l"(g_out + (int64_t)g * K * H), "r"(gbytes), "r"(bar)
                   : "memory");
    if (ldx == H) {
      if (lane == 1)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         mc_smem_addr(xs)),
                     "l"(xc), "r"(rbytes * (uint32_t)rc), "r"(bar)
                     : "memory");
    } else {
      for (int r = lane; r < rc; r += 32)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         mc_smem_addr(xs + (size_t)r * H4)),
                     "l"(xc + (int64_t)r * ldx), "r"(rbytes), "r"(bar)
                     : "memory");
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int e = K * H4 + threadIdx.x; e < KT * H4; e += blockDim.x) gs[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  // the S rows of this warp's first pair travel while the bulk copies land
  float sn0[KT], sn1[KT];
#pragma unroll
  for (int k = 0; k < KT; ++k) { sn0[k] = 0.f; sn1[k] = 0.f; }
  if (WANT_DX && 2 * wid < rc) {
    load_s_row<KT>(s_soft + (int64_t)(base + row0 + 2 * wid) * K, K, sn0);
    load_s_row<KT>(s_soft + (int64_t)(base + row0 + min(2 * wid + 1, rc - 1)) * K, K, sn1);
  }
  __syncthreads();
  mc_bar_wait(bar, 0);
  for (int r0 = 2 * wid; r0 < rc; r0 += 2 * nwarps) {
    const bool two = r0 + 1 < rc;
    const int r1 = two ? r0 + 1 : r0;
    const int64_t i0 = base + row0 + r0, i1 = base + row0 + r1;
    float sv0[KT], sv1[KT], t0[KT], t1[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) { t0[k] = 0.f; t1[k] = 0.f; sv0[k] = sn0[k]; sv1[k] = sn1[k]; }
    if (WANT_DX && r0 + 2 * nwarps < rc) {           // next pair's S rows during this pair's arithmetic
      load_s_row<KT>(s_soft + (int64_t)(base + row0 + r0 + 2 * nwarps) * K, K, sn0);
      load_s_row<KT>(s_soft + (int64_t)(base + row0 + min(r0 + 2 * nwarps + 1, rc - 1)) * K, K, sn1);
    }
    const float4* x0 = xs + (size_t)r0 * H4;
    const float4* x1 = xs + (size_t)r1 * H4;
    for (int c = lane; c < H4; c += 32) {
      const float4 q0 = x0[c], q1 = x1[c];
      float4 d0 = make_float4(0.f, 0.f, 0.f, 0.f), d1 = d0;
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        const float4 gv = gs[k * H4 + c];
        t0[k] = fmaf(q0.x, gv.x, fmaf(q0.y, gv.y, fmaf(q0.z, gv.z, fmaf(q0.w, gv.w, t0[k]))));
        t1[k] = fmaf(q1.x, gv.x, fmaf(q1.y, gv.y, fmaf(q1.z, gv.z, fmaf(q1.w, gv.w, t1[k]))));
        if (WANT_DX) {
          d0.x = fmaf(sv0[k], gv.x, d0.x); d0.y = fmaf(sv0[k], gv.y, d0.y);
          d0.z = fmaf(sv0[k], gv.z, d0.z); d0.w = fmaf(sv0[k], gv.w, d0.w);
          d1.x = fmaf(sv1[k], gv.x, d1.x); d1.y = fmaf(sv1[k], gv.y, d1.y);
          d1.z = fmaf(sv1[k], gv.z, d1.z); d1.w = fmaf(sv1[k], gv.w, d1.w);
        }
      }
      if (WANT_DX) {
        reinterpret_cast<float4*>(d_x + i0 * lddx)[c] = d0;
        if (two) reinterpret_cast<float4*>(d_x + i1 * lddx)[c] = d1;
      }
    }
    if (KT <= 16) {
      // transposed reduction: KT <= 16 partial sums per row over the warp in N shuffles (the plain butterflies were 28 %
      // of the kernel's warp instructions at K = 10)
      constexpr int N = KT <= 4 ? 4 : (KT <= 8 ? 8 : 16);
      float a[N], b[N];
#pragma unroll
      for (int k = 0; k < N; ++k) { a[k] = k < KT ? t0[k] : 0.f; b[k] = k < KT ? t1[k] : 0.f; }
      const float w0 = warp_transpose_sum<N>(a, lane), w1 = warp_transpose_sum<N>(b, lane);
      const int k = warp_transpose_index<N>(lane);
      constexpr int kRest = 32 / N;                 // lanes holding the same index: the lowest of them writes
      if ((lane & (kRest - 1)) == 0 && k < K) {
        xg[i0 * K + k] = w0;
        if (two) xg[i1 * K + k] = w1;
      }
    } else {
      float w0 = 0.f, w1 = 0.f;
#pragma unroll
      for (int k = 0; k < KT; ++k)
        if (k < K) {
          const float a = warp_sum(t0[k]), b = warp_sum(t1[k]);
          if (lane == k) { w0 = a; w1 = b; }
        }
      if (lane < K) {
        xg[i0 * K + lane] = w0;
        if (two) xg[i1 * K + lane] = w1;
      }
    }
  }
}

static inline bool pool_x_supported(int K, int H, const float* x, int64_t ldx, const float* other, int n_cap) {
  return K <= 32 && H >= 4 && H % 4 == 0 && H <= 2048 && ldx % 4 == 0 && n_cap > 0 &&
         ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(other)) & 15) == 0;
}
constexpr int kPoolBwdRows = 32;
static inline size_t pool_x_bwd_smem(int K, int H) { return ((size_t)((K + 3) & ~3) + kPoolBwdRows) * H * 4; }

// forward: feature rows back to back (ldx == H), one pass over the width (H <= 1024), ring + padded S tile in smem
constexpr size_t kPoolSmemLimit = 160 * 1024;
static inline int pool_x_segs(int B, int n_cap) {
  // few graphs: a cluster of 2 / 4 / 8 CTAs per graph (>= 2 CTAs per SM in total, >= 32 rows per CTA at the hint) --
  // B CTAs of 8 warps with the largest graph (3x the mean) as the critical path would leave most of the device idle
  static const int forced = [] {                   // GHSCN_POOL_SEGS=1|2|4|8: A/B measurements
    const char* e = getenv("GHSCN_POOL_SEGS");
    const int v = e ? atoi(e) : 0;
    return (v == 1 || v == 2 || v == 4 || v == 8) ? v : 0;
  }();
  if (forced) return forced;
  int segs = 1;
  while (segs < 8 && (int64_t)B * segs < 2 * kNumSMs && n_cap / (2 * segs) >= 32) segs *= 2;
  return segs;
}
static inline size_t pool_x_ring_bytes(int KT, int H) {
  const int H4 = H / 4, slices = 256 / H4, rows = 4 * slices;
  return std::max((size_t)pool_stages() * rows * H4, (size_t)((slices + 1) / 2) * KT * H4) * sizeof(float4);
}
static inline bool pool_x_fwd_supported(int K, int H, int64_t ldx, int B, int n_cap) {
  if (ldx != H || H > 1024) return false;
  const int KT = (K + 3) & ~3, s_cap = ceil_div(n_cap, pool_x_segs(B, n_cap));
  return pool_x_ring_bytes(KT, H) + (size_t)s_cap * KT * 4 <= kPoolSmemLimit;
}

template <int KT>
static void launch_pool_x(const float* s_soft, const float* x, const int* ptr, int K, int H, int B, int n_cap,
                          float* out, cudaStream_t stream) {
  const size_t ring = pool_x_ring_bytes(KT, H);
  const int segs = pool_x_segs(B, n_cap), s_cap = ceil_div(n_cap, segs);
  const size_t shm = ring + (size_t)s_cap * KT * 4;   // the reduction tree and the cluster tiles reuse the ring
  cudaFuncSetAttribute(mincut_pool_x_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPoolSmemLimit);
  for (int g0 = 0; g0 < B; g0 += 65535) {
    const int gb = std::min(65535, B - g0);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)segs, (unsigned)gb);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = shm;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)segs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, mincut_pool_x_kernel<KT>, s_soft, x, ptr + g0, K, H, s_cap, (int)(ring / sizeof(float4)),
                       segs, pool_stages(), out + (int64_t)g0 * K * H);
  }
}

template <int KT>
static void launch_pool_x_bwd(const float* s_soft, const float* x, int64_t ldx, const int* ptr, const float* g_out,
                              int K, int H, int B, int n_cap, float* xg, float* d_x, int64_t lddx,
                              cudaStream_t stream) {
  const size_t shm = pool_x_bwd_smem(K, H);
  cudaFuncSetAttribute(mincut_pool_x_bwd_kernel<KT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)kSmemBudget);
  cudaFuncSetAttribute(mincut_pool_x_bwd_kernel<KT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)kSmemBudget);
  for (int g0 = 0; g0 < B; g0 += 65535) {
    const int gb = min(65535, B - g0);
    const dim3 grid((unsigned)ceil_div(n_cap, kPoolBwdRows), (unsigned)gb);
    if (d_x)
      mincut_pool_x_bwd_kernel<KT, true><<<grid, 256, shm, stream>>>(s_soft, x, ldx, ptr + g0,
                                                                      g_out + (int64_t)g0 * K * H, K, H, kPoolBwdRows,
                                                                      n_cap, xg, d_x, lddx);
    else
      mincut_pool_x_bwd_kernel<KT, false><<<grid, 256, shm, stream>>>(s_soft, x, ldx, ptr + g0,
                                                                       g_out + (int64_t)g0 * K * H, K, H, kPoolBwdRows,
                                                                       n_cap, xg, nullptr, 0);
  }
}

#define GHSCN_POOL_X_DISPATCH(K, CALL)                                           \
  switch (((K) + 3) / 4) {                                                       \
    case 1: CALL(4); break;                                                      \
    case 2: CALL(8); break;                                                      \
    case 3: CALL(12); break;                                                     \
    case 4: CALL(16); break;                                                     \
    case 5: CALL(20); break;                                                     \
    case 6: CALL(24); break;                                                     \
    case 7: CALL(28); break;                                                     \
    default: CALL(32); break;                                                    \
  }

static inline size_t fwd_smem_bytes(int n_cap, int K, bool smem) {
  return smem ? (2 * (size_t)n_cap * K + n_cap) * 4 : (size_t)n_cap * 4;
}
// slots of a graph's CSR slice staged in shared memory by the fused forward (0 = no staging): 8 per node when the CTA
// then still fits three times on an SM, else 4 per node, else none; graphs with more slots read the global CSR
static inline int fwd_stage_slots(int n_cap, int K, size_t* total_bytes) {
  const size_t base = fwd_smem_bytes(n_cap, K, true);
  const size_t fixed = (size_t)((n_cap + 4) & ~3) * 4 + (K <= 32 ? 2 * (size_t)K * K * 4 : 0);
  int e_cap = 0;
  if (base + fixed + (size_t)8 * n_cap * 8 <= 72 * 1024) e_cap = 8 * n_cap;
  else if (base + fixed + (size_t)4 * n_cap * 8 <= kSmemBudget) e_cap = 4 * n_cap;
  *total_bytes = e_cap ? base + fixed + (size_t)e_cap * 8 : base;
  return e_cap;
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
  return (size_t)4 * num_nodes * num_clusters * 4 + 256;   // [A S | A^T S | dS] (no-smem variant) + x g_out^T
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
  const int threads = mincut_threads(num_graphs, num_clusters);
  // the split phases exchange S and A S through HBM (s_soft, workspace): they use the workspace variant
  const bool smem = phase == 0 && fwd_smem_bytes(n_cap, K, true) <= kSmemBudget;
  size_t staged_bytes = 0;
  const int e_cap = (smem && mincut_stage_enabled()) ? fwd_stage_slots(n_cap, K, &staged_bytes) : 0;
  if (phase == 3) {                                // S, degrees, den; A S = one SpMM over the whole batch
    GHSCN_REQUIRE(workspace != nullptr && workspace_bytes >= (size_t)num_nodes * num_clusters * 4);
  }
  const size_t shm = e_cap ? staged_bytes : fwd_smem_bytes(n_cap, K, smem);
  if (shm > kSmemBudget) return GHSCN_E_UNSUPPORTED;
  float* as_ws = nullptr;
  if (!smem) {
    if (workspace == nullptr || workspace_bytes < (size_t)num_nodes * K * 4) return GHSCN_E_WORKSPACE;
    as_ws = static_cast<float*>(workspace);
  }
  // pooled features S^T X (K <= 32): streamed by a batch-wide kernel behind the fused one, not by the graph's CTA
  const bool stream_out = out != nullptr && phase == 0 && mincut_stream_x_enabled() &&
                          pool_x_supported(K, H, x, ldx, out, n_cap) &&
                          pool_x_fwd_supported(K, H, ldx, (int)num_graphs, n_cap);
  float* out_all = out;
  if (stream_out) out = nullptr;
  if (smem) {
    cudaFuncSetAttribute(mincut_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget);
    mincut_fwd_kernel<true><<<(unsigned)num_graphs, threads, shm, stream>>>(
        logits, ldz, x, ldx, ptr, rowptr, col, adj_val, temp, K, H, n_cap, s_soft, out, out_adj, ss_raw, adj_raw,
        stats, as_ws, phase, e_cap);
  } else {
    mincut_fwd_kernel<false><<<(unsigned)num_graphs, threads, shm, stream>>>(
        logits, ldz, x, ldx, ptr, rowptr, col, adj_val, temp, K, H, n_cap, s_soft, out, out_adj, ss_raw, adj_raw,
        stats, as_ws, phase, e_cap);
  }
  if (phase == 3)
    return ghscn_spmm(rowptr, col, adj_val, s_soft, K, as_ws, K, nullptr, num_nodes, K, 0, stream_);
  if (phase == 1) {
    GHSCN_LAUNCH_CHECK();
    return GHSCN_OK;
  }
  mincut_reduce_losses_kernel<<<1, 256, 0, stream>>>(stats, (int)num_graphs, losses);
  if (stream_out) {
#define GHSCN_CALL(KT) launch_pool_x<KT>(s_soft, x, ptr, K, H, (int)num_graphs, n_cap, out_all, stream)
    GHSCN_POOL_X_DISPATCH(K, GHSCN_CALL);
#undef GHSCN_CALL
    GHSCN_LAUNCH_CHECK_N(3);
    return GHSCN_OK;
  }
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
  const int threads = mincut_threads(num_graphs, num_clusters);
  const bool smem = bwd_smem_bytes(n_cap, K, true) <= kSmemBudget;
  const size_t shm = bwd_smem_bytes(n_cap, K, smem);
  if (shm > kSmemBudget) return GHSCN_E_UNSUPPORTED;
  float* ws = nullptr;
  if (!smem) {
    if (workspace == nullptr || workspace_bytes < (size_t)3 * num_nodes * K * 4) return GHSCN_E_WORKSPACE;
    ws = static_cast<float*>(workspace);
  }
  // gradients through the pooled features (K <= 32): one streaming pass over x / d_x for the whole batch; the
  // per-graph kernel then only adds the [n, K] block it left in the workspace
  float* xg = nullptr;
  // (narrow features, H < 128, stay in the per-graph kernel: half of a warp's lanes would idle in the row pass --
  // K = 16, H = 64, 1 024 graphs: forward + backward 434 us in-kernel vs 471 us streamed)
  if (g_out != nullptr && H >= 128 && mincut_stream_x_enabled() && workspace != nullptr &&
      workspace_bytes >= (size_t)4 * num_nodes * K * 4 && pool_x_supported(K, H, x, ldx, g_out, n_cap) &&
      (d_x == nullptr || (lddx % 4 == 0 && (reinterpret_cast<uintptr_t>(d_x) & 15) == 0)) &&
      pool_x_bwd_smem(K, H) <= kSmemBudget) {
    xg = static_cast<float*>(workspace) + (size_t)3 * num_nodes * K;
#define GHSCN_CALL(KT) \
  launch_pool_x_bwd<KT>(s_soft, x, ldx, ptr, g_out, K, H, (int)num_graphs, n_cap, xg, d_x, lddx, stream)
    GHSCN_POOL_X_DISPATCH(K, GHSCN_CALL);
#undef GHSCN_CALL
    ghscn::note_launches(1);
  }
  if (smem) {
    cudaFuncSetAttribute(mincut_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget);
    mincut_bwd_kernel<true><<<(unsigned)num_graphs, threads, shm, stream>>>(
        s_soft, x, ldx, ptr, rowptr, col, adj_val, rowptr_t, col_t, adj_val_t, temp, (int)num_graphs, (int)num_nodes, K,
        H, n_cap, ss_raw, adj_raw, stats, g_out, g_out_adj, g_losses, d_logits, lddz, d_x, lddx, ws, xg);
  } else {
    cudaFuncSetAttribute(mincut_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget);
    mincut_bwd_kernel<false><<<(unsigned)num_graphs, threads, shm, stream>>>(
        s_soft, x, ldx, ptr, rowptr, col, adj_val, rowptr_t, col_t, adj_val_t, temp, (int)num_graphs, (int)num_nodes, K,
        H, n_cap, ss_raw, adj_raw, stats, g_out, g_out_adj, g_losses, d_logits, lddz, d_x, lddx, ws, xg);
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
