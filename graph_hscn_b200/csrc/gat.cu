// K5: bipartite single-head GAT pool (local -> virtual), forward and backward.
// Replaces GATConv((-1,-1), H, add_self_loops=False) on ("local","to","virtual")
// (model/hscn.py:85-87,118-125).
//
// HBM-bound: reads hs [N,H] once (gathered by cluster membership), 8N score terms, writes [V,H].
// Forward = one warp per destination (virtual) row for the segment softmax of the scores, then the
// generic K2 SpMM with w = alpha for the weighted feature sum (slot = edge order, like CPU scatter_add_).
#include <math.h>

#include "common.cuh"

namespace ghscn {

__global__ void __launch_bounds__(256) row_dot_kernel(const float* __restrict__ x, int64_t ldx,
                                                      const float* __restrict__ v, int num_rows, int num_feat,
                                                      float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int row = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  if (row >= num_rows) return;
  const float* xr = x + (int64_t)row * ldx;
  float acc = 0.f;
  for (int f = lane; f < num_feat; f += 32) acc += __ldg(xr + f) * __ldg(v + f);
  acc = warp_sum(acc);
  if (lane == 0) out[row] = acc;
}

__device__ __forceinline__ float leaky(float z, float slope) { return z > 0.f ? z : z * slope; }

// Attention coefficients of one destination row per warp: alpha[s] = softmax_s(leaky_relu(a_src[col[s]] + a_dst[row])).
// The weighted feature sum itself is the generic K2 SpMM with w = alpha (slot order, unfused mul/add).
__global__ void __launch_bounds__(256) gat_scores_kernel(const int* __restrict__ rowptr,
                                                         const int* __restrict__ col,
                                                         const float* __restrict__ a_src,
                                                         const float* __restrict__ a_dst, float slope,
                                                         int num_rows, float* __restrict__ alpha) {
  const int lane = threadIdx.x & 31;
  const int row = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  if (row >= num_rows) return;
  const int beg = rowptr[row], end = rowptr[row + 1];
  const float ad = a_dst ? a_dst[row] : 0.f;
  if (end - beg <= 32) {  // common case: the whole row fits one lane each, scores stay in registers
    const int s = beg + lane;
    const bool on = s < end;
    const float e = on ? leaky(a_src[col[s]] + ad, slope) : -INFINITY;
    const float m = warp_max(e);
    const float p = on ? expf(e - m) : 0.f;
    const float sum = warp_sum(p) + 1e-16f;   // PyG softmax: + 1e-16 in the denominator
    if (on) alpha[s] = __fdiv_rn(p, sum);
    return;
  }
  float m = -INFINITY;
  for (int s = beg + lane; s < end; s += 32) m = fmaxf(m, leaky(a_src[col[s]] + ad, slope));
  m = warp_max(m);
  float sum = 0.f;
  for (int s = beg + lane; s < end; s += 32) {
    const float p = expf(leaky(a_src[col[s]] + ad, slope) - m);
    alpha[s] = p;
    sum += p;
  }
  sum = warp_sum(sum) + 1e-16f;
  for (int s = beg + lane; s < end; s += 32) alpha[s] = __fdiv_rn(alpha[s], sum);
}

// Backward, destination side.  Per row: dalpha_s = <dout[row], hs[col_s]>; softmax + leaky-relu
// backward give dz_s (gradient of the pre-activation score) and da_dst[row] = sum_s dz_s.
__global__ void __launch_bounds__(256) gat_pool_bwd_scores_kernel(const int* __restrict__ rowptr,
                                                                  const int* __restrict__ col,
                                                                  const float* __restrict__ hs, int64_t ldhs,
                                                                  const float* __restrict__ a_src,
                                                                  const float* __restrict__ a_dst,
                                                                  const float* __restrict__ alpha,
                                                                  const float* __restrict__ dout, int64_t lddout,
                                                                  float slope, int num_rows, int num_feat,
                                                                  float* __restrict__ dz,
                                                                  float* __restrict__ da_dst) {
  const int lane = threadIdx.x & 31;
  const int row = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  if (row >= num_rows) return;
  const int beg = rowptr[row], end = rowptr[row + 1];
  const float* dr = dout + (int64_t)row * lddout;
  float t = 0.f;  // sum_s alpha_s * dalpha_s
  for (int s = beg; s < end; ++s) {
    const float* hr = hs + (int64_t)col[s] * ldhs;
    float acc = 0.f;
    for (int f = lane; f < num_feat; f += 32) acc += __ldg(dr + f) * __ldg(hr + f);
    acc = warp_sum(acc);
    if (lane == 0) dz[s] = acc;
    t += alpha[s] * acc;
  }
  __syncwarp();
  const float ad = a_dst ? a_dst[row] : 0.f;
  float dsum = 0.f;
  for (int s = beg + lane; s < end; s += 32) {
    const float de = alpha[s] * (dz[s] - t);
    const float z = a_src[col[s]] + ad;
    const float g = z > 0.f ? de : de * slope;
    dz[s] = g;
    dsum += g;
  }
  dsum = warp_sum(dsum);
  if (lane == 0 && da_dst) da_dst[row] = dsum;
}

// Backward, source side, on the transposed structure (rows = source nodes):
//   da_src[j] = sum_t dz[map[t]];  dhs[j,:] = sum_t alpha[map[t]] * dout[col_t[t],:] + da_src[j] * att_src
__global__ void __launch_bounds__(256) gat_pool_bwd_src_kernel(const int* __restrict__ rowptr_t,
                                                               const int* __restrict__ col_t,
                                                               const int* __restrict__ map_t,
                                                               const float* __restrict__ alpha,
                                                               const float* __restrict__ dz,
                                                               const float* __restrict__ dout, int64_t lddout,
                                                               const float* __restrict__ att_src, int num_rows,
                                                               int num_feat, float* __restrict__ dhs,
                                                               int64_t lddhs, float* __restrict__ da_src) {
  const int lane = threadIdx.x & 31;
  const int row = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  if (row >= num_rows) return;
  const int beg = rowptr_t[row], end = rowptr_t[row + 1];
  float das = 0.f;
  for (int t = beg; t < end; ++t) das += dz[map_t[t]];
  if (lane == 0 && da_src) da_src[row] = das;
  for (int f = lane; f < num_feat; f += 32) {
    float acc = 0.f;
    for (int t = beg; t < end; ++t) acc += alpha[map_t[t]] * __ldg(dout + (int64_t)col_t[t] * lddout + f);
    dhs[(int64_t)row * lddhs + f] = acc + das * __ldg(att_src + f);
  }
}

// ---- fused attention pool at the input width (forward) -------------------------------------------------------------
// GATConv((-1,-1), H, heads=1, add_self_loops=False) on local -> virtual is linear in the source features once the
// attention coefficients are known:  out_v = (sum_s alpha_s x_s) W_src^T + b,  a_src = x . (W_src^T att_src),
// a_dst = x_dst . (W_dst^T att_dst).  One warp per destination (virtual) row does the whole pool in ONE pass over its
// members' rows: score, leaky-relu, ONLINE softmax (running max / rescaled sums, exact in exact arithmetic) and the
// weighted feature sum, with the feature loads of 4 members in flight.  Replaces row_dot x2, gat_scores, a zero fill
// and the long-row SpMM (5 launches, ~36 us per layer at the bench shape); reads x_src once (HBM-bound: 4 F N bytes).
// u_src / u_dst are the two folded attention vectors (ghscn_gat_fold_attention).
template <int VEC, int ITERS>
__global__ void __launch_bounds__(256) gat_pool_fused_kernel(const int* __restrict__ rowptr,
                                                             const int* __restrict__ col,
                                                             const float* __restrict__ x_src, int64_t ldxs,
                                                             const float* __restrict__ x_dst, int64_t ldxd,
                                                             const float* __restrict__ u_src,
                                                             const float* __restrict__ u_dst, float slope,
                                                             int num_rows, int num_feat, float* __restrict__ pooled,
                                                             int64_t ldp, int wpr) {
  constexpr int W = VEC * ITERS;
  constexpr int kLong = 32;                              // rows with more members are pooled by the whole CTA
  extern __shared__ float part[];                        // [8 warps][2 + 32 * W]: (m, ssum, acc...) per warp
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // wpr = warps per (short) row: 1, 2, 4 or 8.  The grid has only ~num_rows warps of work (1.3 k virtual nodes at the
  // bench shape = 9 warps per SM), so rows of ~15 members are split over several warps whose partials are merged
  // like those of the long rows: 4x the loads in flight, a quarter of the dependent batches per warp.
  const int rpc = 8 / wpr;                               // rows per CTA
  const int row0 = blockIdx.x * rpc;
  // attention head = blockIdx.y: its own folded vectors and its own [num_rows, num_feat] slab of `pooled`
  u_src += (int64_t)blockIdx.y * num_feat;
  if (u_dst != nullptr) u_dst += (int64_t)blockIdx.y * num_feat;
  pooled += (int64_t)blockIdx.y * num_rows * ldp;
  auto load_row = [&](const float* p, float (&dst)[W]) {
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c = (it * 32 + lane) * VEC;
      if constexpr (VEC == 4) {
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < num_feat) q = __ldg(reinterpret_cast<const float4*>(p + c));
        dst[it * 4 + 0] = q.x; dst[it * 4 + 1] = q.y; dst[it * 4 + 2] = q.z; dst[it * 4 + 3] = q.w;
      } else {
        dst[it] = c < num_feat ? __ldg(p + c) : 0.f;
      }
    }
  };
  float us[W];
  load_row(u_src, us);
  // online-softmax pool of the members s = s_beg, s_beg + stride, ... < s_end of destination `row`
  float m, ssum, acc[W];
  auto pool = [&](int row, int s_beg, int s_end, int stride) {
    float ad = 0.f;
    if (x_dst != nullptr && u_dst != nullptr) {
      float ud[W], xd[W];
      load_row(u_dst, ud);
      load_row(x_dst + (int64_t)row * ldxd, xd);
#pragma unroll
      for (int j = 0; j < W; ++j) ad = fmaf(xd[j], ud[j], ad);
      ad = warp_sum(ad);
    }
    m = -INFINITY;
    ssum = 0.f;
#pragma unroll
    for (int j = 0; j < W; ++j) acc[j] = 0.f;
    for (int s = s_beg; s < s_end; s += 4 * stride) {
      float v[4][W];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (s + k * stride < s_end) load_row(x_src + (int64_t)col[s + k * stride] * ldxs, v[k]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (s + k * stride < s_end) {
          float d = 0.f;
#pragma unroll
          for (int j = 0; j < W; ++j) d = fmaf(v[k][j], us[j], d);
          d = warp_sum(d);
          const float z = d + ad;
          const float e = z > 0.f ? z : z * slope;
          if (e > m) {                                   // warp-uniform: every lane holds the same e and m
            const float sc = expf(m - e);                // m = -inf on the first member: exp(-inf) = 0
            ssum *= sc;
#pragma unroll
            for (int j = 0; j < W; ++j) acc[j] *= sc;
            m = e;
          }
          const float p = expf(e - m);
          ssum += p;
#pragma unroll
          for (int j = 0; j < W; ++j) acc[j] = fmaf(p, v[k][j], acc[j]);
        }
      }
    }
  };
  auto store = [&](int row, float inv) {
    float* out = pooled + (int64_t)row * ldp;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c = (it * 32 + lane) * VEC;
      if (c < num_feat) {
        if constexpr (VEC == 4) {
          *reinterpret_cast<float4*>(out + c) = make_float4(acc[it * 4] * inv, acc[it * 4 + 1] * inv,
                                                            acc[it * 4 + 2] * inv, acc[it * 4 + 3] * inv);
        } else {
          out[c] = acc[it] * inv;
        }
      }
    }
  };
  auto park = [&]() {                                    // this warp's (max, sum, weighted sum) partial -> shared memory
    float* mine = part + wid * (2 + 32 * W);
    if (lane == 0) { mine[0] = m; mine[1] = ssum; }
#pragma unroll
    for (int j = 0; j < W; ++j) mine[2 + j * 32 + lane] = acc[j];
  };
  // the log-sum-exp merge of the partials of warps w0 .. w0 + nw - 1, in warp order (deterministic)
  auto merge_store = [&](int row, int w0, int nw) {
    float mm = -INFINITY;
    for (int w = w0; w < w0 + nw; ++w) mm = fmaxf(mm, part[w * (2 + 32 * W)]);
    float tot = 0.f;
#pragma unroll
    for (int j = 0; j < W; ++j) acc[j] = 0.f;
    for (int w = w0; w < w0 + nw; ++w) {
      const float* pw = part + w * (2 + 32 * W);
      const float sc = pw[1] > 0.f ? expf(pw[0] - mm) : 0.f;   // a warp without members holds (m = -inf, sum = 0)
      tot = fmaf(pw[1], sc, tot);
#pragma unroll
      for (int j = 0; j < W; ++j) acc[j] = fmaf(pw[2 + j * 32 + lane], sc, acc[j]);
    }
    store(row, 1.0f / (tot + 1e-16f));                   // PyG softmax: + 1e-16 in the denominator
  };
  // ---- short rows: `wpr` warps each, in one pass
  {
    const int row = row0 + wid / wpr, sub = wid % wpr;
    bool mine_short = false;
    if (row < num_rows) {
      const int beg = rowptr[row], end = rowptr[row + 1];
      if (end - beg <= kLong) {
        mine_short = true;
        pool(row, beg + sub, end, wpr);
        if (wpr == 1) store(row, 1.0f / (ssum + 1e-16f));
        else park();
      }
    }
    if (wpr > 1) {
      __syncthreads();
      if (mine_short && sub == 0) merge_store(row, wid, wpr);
      __syncthreads();
    }
  }
  // ---- long rows (a cluster that swallowed most of its graph): the 8 warps take every 8th member each
  for (int r = 0; r < rpc; ++r) {
    const int row = row0 + r;
    if (row >= num_rows) break;
    const int beg = rowptr[row], end = rowptr[row + 1];
    if (end - beg <= kLong) continue;                    // CTA-uniform
    pool(row, beg + wid, end, 8);
    park();
    __syncthreads();
    if (wid == 0) merge_store(row, 0, 8);
    __syncthreads();
  }
}

// u_src[f] = sum_h att_src[h] W_src[h, f],  u_dst[f] = sum_h att_dst[h] W_dst[h, f]   (W [H, F] row-major)
// CTA = 32 columns x 8 row groups: the 8 partial sums of a column are added in group order (deterministic).
__global__ void __launch_bounds__(256) gat_fold_attention_kernel(const float* __restrict__ w_src, int64_t ldws,
                                                                 const float* __restrict__ att_src,
                                                                 const float* __restrict__ w_dst, int64_t ldwd,
                                                                 const float* __restrict__ att_dst, int H, int Fs,
                                                                 int Fd, float* __restrict__ u_src,
                                                                 float* __restrict__ u_dst) {
  __shared__ float red[8][33];
  const int fx = threadIdx.x & 31, hg = threadIdx.x >> 5;
  const int f = blockIdx.x * 32 + fx;
  const bool dst_side = blockIdx.y == 1;
  const int head = blockIdx.z;                       // head h uses rows h*H .. h*H+H-1 of W and of att
  const int64_t ld = dst_side ? ldwd : ldws;
  const float* w = dst_side ? w_dst : w_src;
  if (w != nullptr) w += (int64_t)head * H * ld;
  const float* att = (dst_side ? att_dst : att_src) + (int64_t)head * H;
  const int F = dst_side ? Fd : Fs;
  if (u_src != nullptr) u_src += (int64_t)head * Fs;
  if (u_dst != nullptr) u_dst += (int64_t)head * Fd;
  float acc = 0.f;
  if (w != nullptr && f < F)
    for (int h = hg; h < H; h += 8) acc = fmaf(__ldg(att + h), __ldg(w + (int64_t)h * ld + f), acc);
  red[hg][fx] = acc;
  __syncthreads();
  if (hg == 0 && w != nullptr && f < F) {
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) t += red[g][fx];
    (dst_side ? u_dst : u_src)[f] = t;
  }
}

// map_t[t] = slot (in the by-destination structure) of the edge sitting in transposed slot t.
__global__ void slot_pos_kernel(const int* __restrict__ perm, int64_t nnz, int* __restrict__ pos) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nnz) pos[perm[s]] = (int)s;
}
__global__ void slot_map_kernel(const int* __restrict__ perm_t, const int* __restrict__ pos, int64_t nnz,
                                int* __restrict__ map_t) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nnz) map_t[t] = pos[perm_t[t]];
}

}  // namespace ghscn

using namespace ghscn;

extern "C" {

int ghscn_row_dot(const float* x, int64_t ldx, const float* v, int64_t num_rows, int64_t num_feat, float* out,
                  ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_rows >= 0 && num_feat >= 0 && num_rows < ((int64_t)1 << 31));
  if (num_rows == 0) return GHSCN_OK;
  GHSCN_REQUIRE(x && v && out && ldx >= num_feat);
  row_dot_kernel<<<(unsigned)ceil_div<int64_t>(num_rows, 8), 256, 0, as_stream(stream)>>>(
      x, ldx, v, (int)num_rows, (int)num_feat, out);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

int ghscn_gat_scores(const int32_t* rowptr, const int32_t* col, const float* a_src, const float* a_dst,
                     float negative_slope, int64_t num_rows, float* alpha, ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_rows >= 0 && num_rows < ((int64_t)1 << 31));
  if (num_rows == 0) return GHSCN_OK;
  GHSCN_REQUIRE(rowptr && a_src);  // col / alpha are zero-sized for an empty relation
  gat_scores_kernel<<<(unsigned)ceil_div<int64_t>(num_rows, 8), 256, 0, as_stream(stream)>>>(
      rowptr, col, a_src, a_dst, negative_slope, (int)num_rows, alpha);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

int ghscn_gat_pool_fwd(const int32_t* rowptr, const int32_t* col, const float* hs, int64_t ldhs,
                       const float* a_src, const float* a_dst, const float* bias, float negative_slope,
                       int64_t num_rows, int64_t num_feat, float* alpha, float* out, int64_t ldout,
                       ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_rows >= 0 && num_feat >= 0 && num_rows < ((int64_t)1 << 31));
  if (num_rows == 0) return GHSCN_OK;
  GHSCN_REQUIRE(rowptr && hs && a_src && out && ldhs >= num_feat && ldout >= num_feat);
  gat_scores_kernel<<<(unsigned)ceil_div<int64_t>(num_rows, 8), 256, 0, as_stream(stream)>>>(
      rowptr, col, a_src, a_dst, negative_slope, (int)num_rows, alpha);
  GHSCN_LAUNCH_CHECK();
  // pooling relation: few destination rows with many members each -> one CTA per row
  const bool vec4 = num_feat % 4 == 0 && ldhs % 4 == 0 && ldout % 4 == 0 &&
                    ((reinterpret_cast<uintptr_t>(hs) | reinterpret_cast<uintptr_t>(out) |
                      reinterpret_cast<uintptr_t>(bias)) % 16 == 0);
  if (vec4 && num_feat >= 64 && num_rows <= 65535)
    return spmm_long_rows(rowptr, col, alpha, hs, ldhs, out, ldout, bias, num_rows, num_feat, as_stream(stream));
  return ghscn_spmm(rowptr, col, alpha, hs, ldhs, out, ldout, bias, num_rows, num_feat, 0, stream);
}

int ghscn_gat_fold_attention(const float* w_src, int64_t ldws, const float* att_src, const float* w_dst, int64_t ldwd,
                             const float* att_dst, int64_t out_feat, int64_t src_feat, int64_t dst_feat, int64_t heads,
                             float* u_src, float* u_dst, ghscn_stream_t stream) {
  GHSCN_REQUIRE(out_feat > 0 && src_feat > 0 && dst_feat >= 0 && w_src && att_src && u_src && ldws >= src_feat);
  GHSCN_REQUIRE(heads >= 1 && heads <= 64);
  GHSCN_REQUIRE(w_dst == nullptr || (att_dst && u_dst && ldwd >= dst_feat && dst_feat > 0));
  const int64_t fmax = src_feat > dst_feat ? src_feat : dst_feat;
  dim3 grid((unsigned)ceil_div<int64_t>(fmax, 32), w_dst ? 2u : 1u, (unsigned)heads);
  gat_fold_attention_kernel<<<grid, 256, 0, as_stream(stream)>>>(w_src, ldws, att_src, w_dst, ldwd, att_dst,
                                                                 (int)out_feat, (int)src_feat, (int)dst_feat, u_src,
                                                                 u_dst);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

int ghscn_gat_pool_fused_supported(int64_t num_feat, int64_t ldxs, int64_t ldp) {
  if (num_feat <= 0) return 0;
  if (num_feat <= 32) return 1;
  return num_feat % 4 == 0 && ldxs % 4 == 0 && ldp % 4 == 0 && num_feat <= 512;
}

int ghscn_gat_pool_fused_fwd(const int32_t* rowptr, const int32_t* col, const float* x_src, int64_t ldxs,
                             const float* x_dst, int64_t ldxd, const float* u_src, const float* u_dst,
                             float negative_slope, int64_t num_rows, int64_t num_feat, int64_t heads, float* pooled,
                             int64_t ldp, int32_t warps_per_row, ghscn_stream_t stream_) {
  GHSCN_REQUIRE(num_rows >= 0 && num_feat > 0 && num_rows < ((int64_t)1 << 31) && heads >= 1 && heads <= 64);
  GHSCN_REQUIRE(warps_per_row == 0 || warps_per_row == 1 || warps_per_row == 2 || warps_per_row == 4 ||
                warps_per_row == 8);
  const int wpr = warps_per_row == 0 ? 1 : warps_per_row;
  if (num_rows == 0) return GHSCN_OK;
  GHSCN_REQUIRE(rowptr && x_src && u_src && pooled && ldxs >= num_feat && ldp >= num_feat);
  GHSCN_REQUIRE((x_dst == nullptr) == (u_dst == nullptr) && (x_dst == nullptr || ldxd >= num_feat));
  cudaStream_t stream = as_stream(stream_);
  const dim3 blocks((unsigned)ceil_div<int64_t>(num_rows, 8 / wpr), (unsigned)heads);
#define GHSCN_POOL_LAUNCH(VEC, ITERS)                                                                          \
  gat_pool_fused_kernel<VEC, ITERS><<<blocks, 256, 8 * (2 + 32 * (VEC) * (ITERS)) * 4, stream>>>(              \
      rowptr, col, x_src, ldxs, x_dst, ldxd, u_src, u_dst, negative_slope, (int)num_rows, (int)num_feat, pooled, ldp, \
      wpr)
  const bool vec4 = num_feat % 4 == 0 && ldxs % 4 == 0 && ldp % 4 == 0 && (x_dst == nullptr || ldxd % 4 == 0) &&
                    ((reinterpret_cast<uintptr_t>(x_src) | reinterpret_cast<uintptr_t>(x_dst) |
                      reinterpret_cast<uintptr_t>(pooled) | reinterpret_cast<uintptr_t>(u_src) |
                      reinterpret_cast<uintptr_t>(u_dst)) % 16 == 0);
  if (num_feat <= 32) GHSCN_POOL_LAUNCH(1, 1);
  else if (!vec4) return GHSCN_E_UNSUPPORTED;
  else if (num_feat <= 128) GHSCN_POOL_LAUNCH(4, 1);
  else if (num_feat <= 256) GHSCN_POOL_LAUNCH(4, 2);
  else if (num_feat <= 384) GHSCN_POOL_LAUNCH(4, 3);
  else if (num_feat <= 512) GHSCN_POOL_LAUNCH(4, 4);
  else return GHSCN_E_UNSUPPORTED;
#undef GHSCN_POOL_LAUNCH
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

int ghscn_gat_pool_bwd_scores(const int32_t* rowptr, const int32_t* col, const float* hs, int64_t ldhs,
                              const float* a_src, const float* a_dst, const float* alpha, const float* dout,
                              int64_t lddout, float negative_slope, int64_t num_rows, int64_t num_feat, float* dz,
                              float* da_dst, ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_rows >= 0 && num_feat >= 0 && num_rows < ((int64_t)1 << 31));
  if (num_rows == 0) return GHSCN_OK;
  GHSCN_REQUIRE(rowptr && col && hs && a_src && alpha && dout && dz);
  gat_pool_bwd_scores_kernel<<<(unsigned)ceil_div<int64_t>(num_rows, 8), 256, 0, as_stream(stream)>>>(
      rowptr, col, hs, ldhs, a_src, a_dst, alpha, dout, lddout, negative_slope, (int)num_rows, (int)num_feat, dz,
      da_dst);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

int ghscn_gat_pool_bwd_src(const int32_t* rowptr_t, const int32_t* col_t, const int32_t* map_t, const float* alpha,
                           const float* dz, const float* dout, int64_t lddout, const float* att_src,
                           int64_t num_rows, int64_t num_feat, float* dhs, int64_t lddhs, float* da_src,
                           ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_rows >= 0 && num_feat >= 0 && num_rows < ((int64_t)1 << 31));
  if (num_rows == 0) return GHSCN_OK;
  GHSCN_REQUIRE(rowptr_t && col_t && map_t && alpha && dz && dout && att_src && dhs);
  gat_pool_bwd_src_kernel<<<(unsigned)ceil_div<int64_t>(num_rows, 8), 256, 0, as_stream(stream)>>>(
      rowptr_t, col_t, map_t, alpha, dz, dout, lddout, att_src, (int)num_rows, (int)num_feat, dhs, lddhs, da_src);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

int ghscn_slot_map(const int32_t* perm, const int32_t* perm_t, int64_t nnz, int64_t num_items, int32_t* scratch_pos,
                   int32_t* map_t, ghscn_stream_t stream_) {
  GHSCN_REQUIRE(nnz >= 0 && num_items >= nnz);
  if (nnz == 0) return GHSCN_OK;
  GHSCN_REQUIRE(perm && perm_t && scratch_pos && map_t);
  cudaStream_t stream = as_stream(stream_);
  slot_pos_kernel<<<(unsigned)ceil_div<int64_t>(nnz, 256), 256, 0, stream>>>(perm, nnz, scratch_pos);
  slot_map_kernel<<<(unsigned)ceil_div<int64_t>(nnz, 256), 256, 0, stream>>>(perm_t, scratch_pos, nnz, map_t);
  GHSCN_LAUNCH_CHECK_N(2);
  return GHSCN_OK;
}

}  // extern "C"
