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
#include <cooperative_groups.h>
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

constexpr int kBK = 16;       // rows per chunk of the dW reduction
constexpr int kGK = 32;       // reduction chunk of the forward / dx tiles

// C[m, n] = epi( sum_k A'(m, k) * B(k, n) ),  A' = A (.) act'(Yref) when Yref != nullptr.
//   B_T = true : B(k, n) = W[n * ldw + k]   (y = x W^T: W is [N, K] row-major)
//   B_T = false: B(k, n) = W[k * ldw + n]   (dx = dy W:  W is [K, N] row-major)
// An optional second operand pair continues the reduction (k >= K reads A2 / W2 at k - K; B_T form only):
//   C = epi(A W^T + A2 W2^T + bias + bias2), the sum of two layers that share their destination rows.
// CTA tile BM x 64, 256 threads, thread tile (BM / 16) x 4.  The reduction runs in chunks of 32 staged k-major in
// shared memory; the global loads of chunk i+1 are issued (into registers) before the FMAs of chunk i, so one
// chunk's DRAM/L2 latency hides behind the previous chunk's arithmetic -- these problems give each SM one or two
// CTAs, so there is no other warp to hide it.  Elements are fetched as 4-wide groups along the contiguous dimension
// (one 128-bit load when the group is whole and aligned, guarded scalars otherwise).
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
  constexpr int A_GROUPS = BM * kGK / 4 / 256;       // 4-wide groups per thread: 1 (BM = 32) or 2 (BM = 64)
  constexpr int B_GROUPS = BN * kGK / 4 / 256;       // 2
  __shared__ __align__(16) float As[kGK][BM + 1];    // odd stride: the transposing stores hit 32 different banks
  __shared__ __align__(16) float Bs[kGK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tx = tid & 15, ty = tid >> 4;            // tx: column group (4 columns), ty: row group (TM rows)
  const int KT = K + K2;
  float acc[TM][TN];
#pragma unroll
  for (int a = 0; a < TM; ++a)
#pragma unroll
    for (int b = 0; b < TN; ++b) acc[a][b] = 0.f;

  // 4 consecutive reduction indices k .. k+3 of row `r` of a k-contiguous operand (A, A2, or W / W2 when B_T)
  auto load_k4 = [&](const float* P, int64_t ld, const float* P2, int64_t ld2, int r, int rmax, int k, float (&v)[4]) {
    v[0] = v[1] = v[2] = v[3] = 0.f;
    if (r >= rmax || k >= KT) return;
    const float* src;
    int kk, klim;
    if (k < K) { src = P + (int64_t)r * ld; kk = k; klim = K; }
    else       { src = P2 + (int64_t)r * ld2; kk = k - K; klim = K2; }
    if (kk + 3 < klim && ((reinterpret_cast<uintptr_t>(src + kk) & 15) == 0)) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(src + kk));
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        // a group may straddle the boundary between the two operand pairs: resolve every element on its own
        const int ku = k + u;
        if (ku < K) v[u] = P[(int64_t)r * ld + ku];
        else if (ku < KT) v[u] = P2[(int64_t)r * ld2 + (ku - K)];
      }
    }
  };
  float ra[A_GROUPS][4], rb[B_GROUPS][4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int gI = 0; gI < A_GROUPS; ++gI) {
      const int e = tid + gI * 256;
      const int r = e >> 3, kq = (e & 7) * 4;                    // 8 groups of 4 per 32-wide row chunk
      load_k4(A, lda, A2, lda2, m0 + r, M, k0 + kq, ra[gI]);
      if (Yref != nullptr && m0 + r < M) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (k0 + kq + u < K) ra[gI][u] *= act_grad_from_output(Yref[(int64_t)(m0 + r) * ldy + k0 + kq + u], mask_act);
      }
    }
#pragma unroll
    for (int gI = 0; gI < B_GROUPS; ++gI) {
      const int e = tid + gI * 256;
      if (B_T) {
        const int c = e >> 3, kq = (e & 7) * 4;
        load_k4(W, ldw, W2, ldw2, n0 + c, N, k0 + kq, rb[gI]);
      } else {
        const int kk = e >> 4, cq = (e & 15) * 4;                // 16 groups of 4 columns per reduction row
        const int k = k0 + kk, n = n0 + cq;
        rb[gI][0] = rb[gI][1] = rb[gI][2] = rb[gI][3] = 0.f;
        if (k < K) {
          const float* src = W + (int64_t)k * ldw + n;
          if (n + 3 < N && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(src));
            rb[gI][0] = q.x; rb[gI][1] = q.y; rb[gI][2] = q.z; rb[gI][3] = q.w;
          } else {
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (n + u < N) rb[gI][u] = src[u];
          }
        }
      }
    }
  };
  auto stage = [&]() {
#pragma unroll
    for (int gI = 0; gI < A_GROUPS; ++gI) {
      const int e = tid + gI * 256;
      const int r = e >> 3, kq = (e & 7) * 4;
#pragma unroll
      for (int u = 0; u < 4; ++u) As[kq + u][r] = ra[gI][u];
    }
#pragma unroll
    for (int gI = 0; gI < B_GROUPS; ++gI) {
      const int e = tid + gI * 256;
      if (B_T) {
        const int c = e >> 3, kq = (e & 7) * 4;
#pragma unroll
        for (int u = 0; u < 4; ++u) Bs[kq + u][c] = rb[gI][u];
      } else {
        const int kk = e >> 4, cq = (e & 15) * 4;
        *reinterpret_cast<float4*>(&Bs[kk][cq]) = make_float4(rb[gI][0], rb[gI][1], rb[gI][2], rb[gI][3]);
      }
    }
  };

  fetch(0);
  for (int k0 = 0; k0 < KT; k0 += kGK) {
    stage();
    __syncthreads();
    if (k0 + kGK < KT) fetch(k0 + kGK);               // in flight while this chunk is multiplied
#pragma unroll
    for (int kk = 0; kk < kGK; ++kk) {
      float av[TM];
#pragma unroll
      for (int a = 0; a < TM; ++a) av[a] = As[kk][ty * TM + a];
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
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

// ---- the same products, pipelined and with the reduction split across the warps of a CTA --------------------------
// (the path taken whenever rows are 16-byte aligned)
// The scalar tile kernels above are latency-bound three ways at these sizes: one exposed memory latency per chunk,
// 3 bytes of shared memory per FMA, and -- with one thread accumulating a whole output over all of K -- a serial
// chain of K x 16 FMA issue slots per warp.  Here
//   * a CTA (8 warps) owns a 32 x 32 output tile; inside every 32-wide reduction chunk warp w multiplies only the
//     4 reduction indices 4w .. 4w+3 (for dW: the 4 rows g = 4w .. 4w+3), so the per-warp chain is K/8 long; the 8
//     partial tiles are added at the end in warp order through shared memory (deterministic);
//   * a thread holds an 8 x 4 register tile (rows ty + 4a, columns tx + 8b): 12 float4 shared loads per 128 FMAs;
//   * the chunks of the next three stages are in flight (cp.async, 16 bytes per request, zero-filled past the edges)
//     while the current one is multiplied (a ring of 8 measured no faster: what is left at 129 rows is the ~3 us
//     dependent-launch floor, one first-load latency and ~0.15 us of arithmetic per chunk); operands stay
//     k-contiguous in shared memory with a row stride of 36 floats, which makes the float4 reads along k
//     bank-conflict free.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kStages = 4;
constexpr int kT = 32;             // output tile edge of the pipelined kernels
constexpr int kLdK = kGK + 4;      // row stride (floats) of a k-contiguous tile: 36 -> 144 B, 16-byte aligned
constexpr int kLdN = kT + 4;       // row stride of an n-contiguous tile
constexpr int kRedLd = kT * kT + 8;  // one warp's partial tile in the final cross-warp sum

// adds the 8 warps' partial 32 x 32 tiles (held as acc[a][b] for rows ty + 4a, columns tx + 8b) in warp order and
// returns, for thread t, the 4 consecutive outputs 4t .. 4t+3 of the row-major tile.
__device__ __forceinline__ float4 cross_warp_sum(float (&acc)[8][4], float* red, int wid, int ty, int tx, int tid) {
  __syncthreads();                                     // every warp is done with the pipeline buffers (reused here)
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) red[wid * kRedLd + (ty + 4 * a) * kT + tx + 8 * b] = acc[a][b];
  __syncthreads();
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    const float4 p = *reinterpret_cast<const float4*>(red + w * kRedLd + tid * 4);
    s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
  }
  return s;
}

template <bool B_T, bool MASK>
__global__ void __launch_bounds__(256) small_gemm_async_kernel(const float* __restrict__ A, int64_t lda,
                                                               const float* __restrict__ Yref, int64_t ldy,
                                                               int mask_act, const float* __restrict__ W, int64_t ldw,
                                                               const float* __restrict__ bias, int act, int M, int N,
                                                               int K, float* __restrict__ C, int64_t ldc,
                                                               const float* __restrict__ A2, int64_t lda2,
                                                               const float* __restrict__ W2, int64_t ldw2, int K2,
                                                               const float* __restrict__ bias2) {
  constexpr int A_FLOATS = kT * kLdK, B_FLOATS = B_T ? kT * kLdK : kGK * kLdN;
  constexpr int STAGE_FLOATS = A_FLOATS * (MASK ? 2 : 1) + B_FLOATS;
  constexpr int SMEM_FLOATS = kStages * STAGE_FLOATS > 8 * kRedLd ? kStages * STAGE_FLOATS : 8 * kRedLd;
  extern __shared__ __align__(16) float smem[];
  (void)SMEM_FLOATS;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int m0 = blockIdx.x * kT, n0 = blockIdx.y * kT;
  const int tx = lane & 7, ty = lane >> 3;             // inside the warp: 4 row groups x 8 column groups
  const int nc1 = ceil_div(K, kGK), nc2 = ceil_div(K2, kGK), nchunks = nc1 + nc2;
  float acc[8][4];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  auto issue = [&](int c) {
    if (c < nchunks) {
      const bool second = c >= nc1;
      const float* Ap = second ? A2 : A;
      const float* Wp = second ? W2 : W;
      const int64_t la = second ? lda2 : lda, lw = second ? ldw2 : ldw;
      const int Ks = second ? K2 : K, k0 = (second ? c - nc1 : c) * kGK;
      float* st = smem + (c % kStages) * STAGE_FLOATS;
      float* As = st;
      float* Ys = st + A_FLOATS;
      float* Bs = st + A_FLOATS * (MASK ? 2 : 1);
      {                                                // 32 rows x 8 groups of 4 = 256 requests per operand
        const int r = tid >> 3, kq = (tid & 7) * 4;
        const bool ok = (m0 + r < M) && (k0 + kq < Ks);
        cp_async16(As + r * kLdK + kq, ok ? Ap + (int64_t)(m0 + r) * la + k0 + kq : Ap, ok ? 16 : 0);
        if (MASK) cp_async16(Ys + r * kLdK + kq, ok ? Yref + (int64_t)(m0 + r) * ldy + k0 + kq : Yref, ok ? 16 : 0);
      }
      if (B_T) {
        const int c2 = tid >> 3, kq = (tid & 7) * 4;
        const bool ok = (n0 + c2 < N) && (k0 + kq < Ks);
        cp_async16(Bs + c2 * kLdK + kq, ok ? Wp + (int64_t)(n0 + c2) * lw + k0 + kq : Wp, ok ? 16 : 0);
      } else {
        const int kk = tid >> 3, cq = (tid & 7) * 4;
        const int nb = n0 + cq;
        const int bytes = (k0 + kk < Ks && nb < N) ? min(16, (N - nb) * 4) : 0;
        cp_async16(Bs + kk * kLdN + cq, bytes ? Wp + (int64_t)(k0 + kk) * lw + nb : Wp, bytes);
      }
    }
    cp_async_commit();
  };

  for (int c = 0; c < kStages - 1; ++c) issue(c);
  const int kk = wid * 4;                              // this warp's 4 reduction indices inside every chunk
  for (int c = 0; c < nchunks; ++c) {
    cp_async_wait<kStages - 2>();
    __syncthreads();
    issue(c + kStages - 1);
    const float* st = smem + (c % kStages) * STAGE_FLOATS;
    const float* As = st;
    const float* Ys = st + A_FLOATS;
    const float* Bs = st + A_FLOATS * (MASK ? 2 : 1);
    float4 a4[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      a4[a] = *reinterpret_cast<const float4*>(As + (ty + 4 * a) * kLdK + kk);
      if (MASK) {
        const float4 y4 = *reinterpret_cast<const float4*>(Ys + (ty + 4 * a) * kLdK + kk);
        a4[a].x *= act_grad_from_output(y4.x, mask_act);
        a4[a].y *= act_grad_from_output(y4.y, mask_act);
        a4[a].z *= act_grad_from_output(y4.z, mask_act);
        a4[a].w *= act_grad_from_output(y4.w, mask_act);
      }
    }
    if (B_T) {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const float4 b4 = *reinterpret_cast<const float4*>(Bs + (tx + 8 * b) * kLdK + kk);
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          acc[a][b] = fmaf(a4[a].x, b4.x, acc[a][b]);
          acc[a][b] = fmaf(a4[a].y, b4.y, acc[a][b]);
          acc[a][b] = fmaf(a4[a].z, b4.z, acc[a][b]);
          acc[a][b] = fmaf(a4[a].w, b4.w, acc[a][b]);
        }
      }
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float bv[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) bv[b] = Bs[(kk + u) * kLdN + tx + 8 * b];
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          const float av = u == 0 ? a4[a].x : (u == 1 ? a4[a].y : (u == 2 ? a4[a].z : a4[a].w));
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av, bv[b], acc[a][b]);
        }
      }
    }
  }
  cp_async_wait<0>();
  const float4 sum = cross_warp_sum(acc, smem, wid, ty, tx, tid);
  const int m = m0 + (tid * 4) / kT, nb = n0 + (tid * 4) % kT;
  if (m < M) {
    const float v[4] = {sum.x, sum.y, sum.z, sum.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int n = nb + u;
      if (n < N) {
        float o = v[u];
        if (bias) o += bias[n];
        if (bias2) o += bias2[n];
        C[(int64_t)m * ldc + n] = act_apply(o, act);
      }
    }
  }
}

template <bool B_T, bool MASK>
static constexpr size_t small_gemm_async_smem() {
  constexpr size_t pipe = (size_t)kStages * (kT * kLdK * (MASK ? 2 : 1) + (B_T ? kT * kLdK : kGK * kLdN));
  constexpr size_t red = (size_t)8 * kRedLd;
  return (pipe > red ? pipe : red) * sizeof(float);
}

// dW / db with the same structure: chunks of 32 rows g, warp w multiplies rows 4w .. 4w+3 of every chunk; tiles
// As[g][n], Ys[g][n], Bs[g][k] are contiguous along their output dimension; a thread's register tile is rows (n)
// ty + 4a, columns (k) tx + 8b like above, read as scalars (conflict free: consecutive tx, broadcast ty).
__global__ void __launch_bounds__(256) small_gemm_tn_async_kernel(const float* __restrict__ dY, int64_t lddy,
                                                                  const float* __restrict__ Yref, int64_t ldy,
                                                                  int mask_act, const float* __restrict__ X,
                                                                  int64_t ldx, int G, int N, int K, int rows_per_split,
                                                                  float* __restrict__ dW, float* __restrict__ db,
                                                                  float* __restrict__ part) {
  constexpr int TILE = kGK * kLdN;
  extern __shared__ __align__(16) float smem[];
  __shared__ float bred[8][kT];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n0 = blockIdx.x * kT, c0 = blockIdx.y * kT;
  const int g_beg = blockIdx.z * rows_per_split, g_end = min(G, g_beg + rows_per_split);
  const int nchunks = g_end > g_beg ? ceil_div(g_end - g_beg, kGK) : 0;
  const int tx = lane & 7, ty = lane >> 3;
  const bool mask = Yref != nullptr;
  float acc[8][4], bacc[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    bacc[a] = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  }
  auto issue = [&](int c) {
    if (c < nchunks) {
      float* st = smem + (c % kStages) * 3 * TILE;
      const int g0 = g_beg + c * kGK;
      const int gg = tid >> 3, q = (tid & 7) * 4;
      const int g = g0 + gg;
      const int nb = n0 + q, kb = c0 + q;
      const int bn = (g < g_end && nb < N) ? min(16, (N - nb) * 4) : 0;
      const int bk = (g < g_end && kb < K) ? min(16, (K - kb) * 4) : 0;
      cp_async16(st + gg * kLdN + q, bn ? dY + (int64_t)g * lddy + nb : dY, bn);
      if (mask) cp_async16(st + TILE + gg * kLdN + q, bn ? Yref + (int64_t)g * ldy + nb : Yref, bn);
      cp_async16(st + 2 * TILE + gg * kLdN + q, bk ? X + (int64_t)g * ldx + kb : X, bk);
    }
    cp_async_commit();
  };
  for (int c = 0; c < kStages - 1; ++c) issue(c);
  for (int c = 0; c < nchunks; ++c) {
    cp_async_wait<kStages - 2>();
    __syncthreads();
    issue(c + kStages - 1);
    const float* st = smem + (c % kStages) * 3 * TILE;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int gg = wid * 4 + u;
      float av[8], bv[4];
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        av[a] = st[gg * kLdN + ty + 4 * a];
        if (mask) av[a] *= act_grad_from_output(st[TILE + gg * kLdN + ty + 4 * a], mask_act);
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) bv[b] = st[2 * TILE + gg * kLdN + tx + 8 * b];
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        bacc[a] += av[a];
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
      }
    }
  }
  cp_async_wait<0>();
  const float4 sum = cross_warp_sum(acc, smem, wid, ty, tx, tid);
  if (tx == 0) {
#pragma unroll
    for (int a = 0; a < 8; ++a) bred[wid][ty + 4 * a] = bacc[a];
  }
  __syncthreads();
  const bool direct = gridDim.z == 1;
  float* wout = direct ? dW : part + (int64_t)blockIdx.z * ((int64_t)N * K + N);
  float* bout = direct ? db : wout + (int64_t)N * K;
  const int n = n0 + (tid * 4) / kT, kb = c0 + (tid * 4) % kT;
  if (n < N) {
    const float v[4] = {sum.x, sum.y, sum.z, sum.w};
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (kb + u < K) wout[(int64_t)n * K + kb + u] = v[u];
  }
  if (bout && blockIdx.y == 0 && tid < kT && n0 + tid < N) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += bred[w][tid];
    bout[n0 + tid] = t;
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// forward / dx dispatch: pipelined kernel when every row start is 16-byte aligned, scalar tiles otherwise
template <bool B_T>
static void launch_small_gemm(const float* A, int64_t lda, const float* Yref, int64_t ldy, int mask_act,
                              const float* W, int64_t ldw, const float* bias, int act, int M, int N, int K, float* C,
                              int64_t ldc, const float* A2, int64_t lda2, const float* W2, int64_t ldw2, int K2,
                              const float* bias2, cudaStream_t stream) {
  const unsigned gy = (unsigned)ceil_div(N, 64);
  const bool big = (int64_t)ceil_div(M, 64) * gy >= kNumSMs;
  bool fast = aligned16(A) && aligned16(W) && lda % 4 == 0 && ldw % 4 == 0 && K % 4 == 0 &&
              (Yref == nullptr || (aligned16(Yref) && ldy % 4 == 0));
  if (K2 > 0) fast = fast && aligned16(A2) && aligned16(W2) && lda2 % 4 == 0 && ldw2 % 4 == 0 && K2 % 4 == 0;
  if (!B_T) fast = fast && true;                       // W rows [k][n]: partial 16-byte groups are size-limited
#define GHSCN_SG_ASYNC(MASK)                                                                                      \
  do {                                                                                                            \
    constexpr size_t shm = small_gemm_async_smem<B_T, MASK>();                                                    \
    cudaFuncSetAttribute(small_gemm_async_kernel<B_T, MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                         (int)shm);                                                                               \
    dim3 grid((unsigned)ceil_div(M, kT), (unsigned)ceil_div(N, kT));                                              \
    small_gemm_async_kernel<B_T, MASK><<<grid, 256, shm, stream>>>(A, lda, Yref, ldy, mask_act, W, ldw, bias,     \
                                                                   act, M, N, K, C, ldc, A2, lda2, W2, ldw2, K2,  \
                                                                   bias2);                                        \
  } while (0)
  if (fast) {
    if (Yref) GHSCN_SG_ASYNC(true); else GHSCN_SG_ASYNC(false);
#undef GHSCN_SG_ASYNC
    return;
  }
  if (big) {
    dim3 grid((unsigned)ceil_div(M, 64), gy);
    small_gemm_kernel<64, B_T><<<grid, 256, 0, stream>>>(A, lda, Yref, ldy, mask_act, W, ldw, bias, act, M, N, K, C,
                                                         ldc, A2, lda2, W2, ldw2, K2, bias2);
  } else {
    dim3 grid((unsigned)ceil_div(M, 32), gy);
    small_gemm_kernel<32, B_T><<<grid, 256, 0, stream>>>(A, lda, Yref, ldy, mask_act, W, ldw, bias, act, M, N, K, C,
                                                         ldc, A2, lda2, W2, ldw2, K2, bias2);
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

// ---- output layer + task loss + their backward in ONE launch (small batches) ---------------------------------------
// model/hscn.py:112 `self.lin_2(...)`, loss.py:6-19 `criterion`, and the backward of both w.r.t. lin_2's parameters and
// its input: pred = h W2^T + b2, loss / score / d_pred as in graph_loss_kernel, dW2 = d_pred^T h, db2 = colsum(d_pred),
// dh = d_pred W2.  ONE cluster of 8 CTAs: each CTA owns total_rows/8 rows (their h rows, W2 and d_pred live in its
// shared memory), the loss and dW2 / db2 partials are combined through distributed shared memory in rank order, so
// every reduction runs in a fixed order.  Five launches of ~3 us dependent-launch latency each become one.
constexpr int kHeadCluster = 8;

__global__ void __launch_bounds__(1024) head_out_loss_kernel(const float* __restrict__ h, int64_t ldh,
                                                             const float* __restrict__ w2, int64_t ldw,
                                                             const float* __restrict__ b2,
                                                             const float* __restrict__ target, int64_t ldt, int rows,
                                                             int total_rows, int H, int C, int mode,
                                                             float* __restrict__ pred, float* __restrict__ loss,
                                                             float* __restrict__ score, float* __restrict__ d_w2,
                                                             float* __restrict__ d_b2, float* __restrict__ d_h) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ float hs_sm[];
  __shared__ float red[32];
  const int rank = (int)cluster.block_rank();
  const int per = (total_rows + kHeadCluster - 1) / kHeadCluster;
  const int g0 = min(rank * per, total_rows), nloc = min(total_rows, g0 + per) - g0;
  float* s_w = hs_sm;                         // [C][H]
  float* s_h = s_w + (size_t)C * H;           // [per][H]
  float* s_dp = s_h + (size_t)per * H;        // [per][C]: pred, then d_pred
  float* s_dw = s_dp + (size_t)per * C;       // [C*H + C + 1]: this CTA's dW2 | db2 | loss partial
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = blockDim.x >> 5;
  const int CH = C * H;
  // staging: every thread issues all of its loads before the first store (one memory round trip, not one per element)
  if (H % 4 == 0 && ldw % 4 == 0 && ldh % 4 == 0 && (((uintptr_t)w2 | (uintptr_t)h) & 15) == 0) {
    const int H4 = H / 4, nw = C * H4, nh = nloc * H4;
    float4 v[6];
    for (int i0 = 0; i0 < nw + nh; i0 += 6 * blockDim.x) {
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        const int i = i0 + u * blockDim.x + tid;
        if (i < nw) v[u] = *reinterpret_cast<const float4*>(w2 + (int64_t)(i / H4) * ldw + (i % H4) * 4);
        else if (i < nw + nh)
          v[u] = *reinterpret_cast<const float4*>(h + (int64_t)(g0 + (i - nw) / H4) * ldh + ((i - nw) % H4) * 4);
      }
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        const int i = i0 + u * blockDim.x + tid;
        if (i < nw) *reinterpret_cast<float4*>(s_w + (size_t)i * 4) = v[u];
        else if (i < nw + nh) *reinterpret_cast<float4*>(s_h + (size_t)(i - nw) * 4) = v[u];
      }
    }
  } else {
    for (int i = tid; i < CH; i += blockDim.x) s_w[i] = w2[(int64_t)(i / H) * ldw + i % H];
    for (int i = tid; i < nloc * H; i += blockDim.x) s_h[i] = h[(int64_t)(g0 + i / H) * ldh + i % H];
  }
  __syncthreads();
  // 1. pred: one warp per row
  for (int g = wid; g < nloc; g += nwarps) {
    float hv[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int k = lane + 32 * j;
      hv[j] = k < H ? s_h[g * H + k] : 0.f;
    }
    for (int c = 0; c < C; ++c) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int k = lane + 32 * j;
        if (k < H) acc = fmaf(hv[j], s_w[c * H + k], acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) s_dp[g * C + c] = acc + (b2 ? b2[c] : 0.f);
    }
  }
  __syncthreads();
  // 2. loss partial, score, d_pred (in place)
  const float inv = 1.0f / (float)max(rows * C, 1);
  float acc = 0.f;
  for (int e = tid; e < nloc * C; e += blockDim.x) {
    const int r = g0 + e / C, c = e % C;
    const int64_t o = (int64_t)r * C + c;
    const float x = s_dp[e];
    pred[o] = x;
    const float sg = 1.0f / (1.0f + expf(-x));
    if (score) score[o] = sg;
    float d = 0.f;
    if (r < rows) {
      const float t = target[(int64_t)r * ldt + c];
      if (mode == 0) {
        acc += (1.0f - t) * x - (fminf(x, 0.f) - log1pf(expf(-fabsf(x))));
        d = (sg - t) * inv;
      } else {
        const float df = x - t;
        acc += fabsf(df);
        d = (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f)) * inv;
      }
    }
    s_dp[e] = d;
  }
  acc = block_sum(acc, red);
  if (tid == 0) s_dw[CH + C] = acc;
  __syncthreads();
  // 3. this CTA's partial dW2 [C, H] | db2 [C] over its rows, in row order
  for (int o = tid; o < CH + C; o += blockDim.x) {
    float a = 0.f;
    if (o < CH) {
      const int c = o / H, k = o - c * H;
      for (int g = 0; g < nloc; ++g) a = fmaf(s_dp[g * C + c], s_h[g * H + k], a);
    } else {
      const int c = o - CH;
      for (int g = 0; g < nloc; ++g) a += s_dp[g * C + c];
    }
    s_dw[o] = a;
  }
  // 4. dh = d_pred W2 for this CTA's rows (zero for the padding rows: their d_pred is zero)
  for (int o = tid; o < nloc * H; o += blockDim.x) {
    const int g = o / H, k = o - g * H;
    float a = 0.f;
    for (int c = 0; c < C; ++c) a = fmaf(s_dp[g * C + c], s_w[c * H + k], a);
    d_h[(int64_t)(g0 + g) * H + k] = a;
  }
  cluster.sync();
  // 5. combine the partials in rank order: CTA r owns a slice of dW2 | db2 | loss
  const int n_out = CH + C + 1, chunk = (n_out + kHeadCluster - 1) / kHeadCluster;
  for (int o = rank * chunk + tid; o < min(n_out, (rank + 1) * chunk); o += blockDim.x) {
    float a = 0.f;
#pragma unroll
    for (int q = 0; q < kHeadCluster; ++q) a += cluster.map_shared_rank(s_dw, q)[o];
    if (o < CH) d_w2[o] = a;
    else if (o < CH + C) { if (d_b2 != nullptr) d_b2[o - CH] = a; }
    else loss[0] = a * inv;
  }
  cluster.sync();                              // nobody leaves while a peer still reads its shared memory
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
  launch_small_gemm<true>(x, ldx, nullptr, 0, 0, w, ldw, bias, act, (int)num_rows, (int)out_feat, (int)in_feat, y, ldy,
                          nullptr, 0, nullptr, 0, 0, nullptr, as_stream(stream));
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
  launch_small_gemm<true>(x1, ldx1, nullptr, 0, 0, w1, ldw1, bias1, act, (int)num_rows, (int)out_feat, (int)in_feat1, y,
                          ldy, x2, ldx2, w2, ldw2, (int)in_feat2, bias2, as_stream(stream));
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
  // dx [M, in] = dy' [M, out] . W [out, in]
  launch_small_gemm<false>(dy, lddy, act == 0 ? nullptr : y_ref, ldy, act, w, ldw, nullptr, 0, (int)num_rows,
                           (int)in_feat, (int)out_feat, dx, lddx, nullptr, 0, nullptr, 0, 0, nullptr, as_stream(stream));
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
  const int rows_per_split = ceil_div(ceil_div(G, z), kGK) * kGK;
  dim3 grid((unsigned)ceil_div(N, 64), (unsigned)ceil_div(K, 64), (unsigned)z);
  const float* yr = act == 0 ? nullptr : y_ref;
  const int rps = rows_per_split > 0 ? rows_per_split : kGK;
  if (aligned16(dy) && aligned16(x) && lddy % 4 == 0 && ldx % 4 == 0 && (yr == nullptr || (aligned16(yr) && ldy % 4 == 0))) {
    dim3 grid32((unsigned)ceil_div(N, kT), (unsigned)ceil_div(K, kT), (unsigned)z);
    constexpr size_t pipe = (size_t)kStages * 3 * kGK * kLdN, red = (size_t)8 * kRedLd;
    constexpr size_t shm = (pipe > red ? pipe : red) * sizeof(float);
    cudaFuncSetAttribute(small_gemm_tn_async_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
    small_gemm_tn_async_kernel<<<grid32, 256, shm, stream>>>(dy, lddy, yr, ldy, act, x, ldx, G, N, K, rps, dw, db,
                                                             static_cast<float*>(workspace));
  } else {
    small_gemm_tn_kernel<<<grid, 256, 0, stream>>>(dy, lddy, yr, ldy, act, x, ldx, G, N, K, rps, dw, db,
                                                   static_cast<float*>(workspace));
  }
  if (z > 1) {
    const int64_t total = (int64_t)N * K + N;
    small_reduce_kernel<<<(unsigned)ceil_div<int64_t>(total, 256), 256, 0, stream>>>(
        static_cast<const float*>(workspace), z, (int64_t)N * K, N, dw, db);
  }
  GHSCN_LAUNCH_CHECK_N(z > 1 ? 2 : 1);
  return GHSCN_OK;
}

static size_t head_out_loss_smem(int64_t total_rows, int64_t hidden, int64_t num_targets) {
  const int64_t per = (total_rows + kHeadCluster - 1) / kHeadCluster;
  return (size_t)(2 * num_targets * hidden + per * hidden + per * num_targets + num_targets + 1) * sizeof(float);
}

int ghscn_head_out_loss_supported(int64_t total_rows, int64_t hidden, int64_t num_targets) {
  if (total_rows <= 0 || total_rows > 256 || hidden <= 0 || hidden > 512 || num_targets <= 0 || num_targets > 32) return 0;
  return head_out_loss_smem(total_rows, hidden, num_targets) <= 200 * 1024;
}

int ghscn_head_out_loss(const float* h, int64_t ldh, const float* w2, int64_t ldw, const float* b2, const float* target,
                        int64_t ldt, int64_t rows, int64_t total_rows, int64_t hidden, int64_t num_targets,
                        int32_t mode, float* pred, float* loss, float* score, float* d_w2, float* d_b2, float* d_h,
                        ghscn_stream_t stream) {
  GHSCN_REQUIRE(rows >= 0 && total_rows >= rows && (mode == 0 || mode == 1));
  if (!ghscn_head_out_loss_supported(total_rows, hidden, num_targets)) return GHSCN_E_UNSUPPORTED;
  GHSCN_REQUIRE(h && w2 && pred && loss && d_w2 && d_h && ldh >= hidden && ldw >= hidden);
  GHSCN_REQUIRE(rows == 0 || (target && ldt >= num_targets));
  const size_t shm = head_out_loss_smem(total_rows, hidden, num_targets);
  if (shm > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(head_out_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
    if (e != cudaSuccess) return (int)e;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kHeadCluster);
  cfg.blockDim = dim3(1024);
  cfg.dynamicSmemBytes = shm;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kHeadCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, head_out_loss_kernel, h, ldh, w2, ldw, b2, target, ldt, (int)rows,
                                     (int)total_rows, (int)hidden, (int)num_targets, (int)mode, pred, loss, score,
                                     d_w2, d_b2, d_h);
  if (e != cudaSuccess) return (int)e;
  GHSCN_LAUNCH_CHECK();
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
