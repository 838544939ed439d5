// K4: segment mean / sum over contiguous row ranges (graph readout), its broadcast backward, and the
// column sum used for bias gradients.
// Replaces torch_scatter.scatter_mean (model/mpnn.py:60) / global_mean_pool (model/hscn.py:111).
// HBM-bound: 4F(N + B) bytes.  One CTA owns one (segment, 128-column tile): its 8 warps stride the
// segment's rows (coalesced 128-bit loads, 2 rows in flight per warp) and the 8 partial sums are combined
// in fixed warp order in shared memory, so the result is deterministic (run-to-run bit-identical).
#include "common.cuh"

namespace ghscn {

constexpr int kSegWarps = 8;

template <int VEC>
__global__ void __launch_bounds__(kSegWarps * 32) segment_reduce_kernel(const float* __restrict__ x, int64_t ldx,
                                                                        const int* __restrict__ ptr,
                                                                        const int* __restrict__ perm, int num_feat,
                                                                        int mean, float* __restrict__ y,
                                                                        int64_t ldy) {
  __shared__ float part[kSegWarps][32 * VEC];
  const int g = blockIdx.y;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int f = (blockIdx.x * 32 + lane) * VEC;
  const bool on = f < num_feat;
  const int beg = ptr[g], end = ptr[g + 1];
  float acc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
  if (on) {
    int i = beg + wid;
    if (VEC == 4) {      // four rows in flight per warp, added in row order (the order of the two-row loop below)
      for (; i + 3 * kSegWarps < end; i += 4 * kSegWarps) {
        float4 q[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t r = perm ? perm[i + j * kSegWarps] : (i + j * kSegWarps);
          q[j] = ldg_f4(x + r * ldx + f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[0] += q[j].x; acc[1 % VEC] += q[j].y; acc[2 % VEC] += q[j].z; acc[3 % VEC] += q[j].w;
        }
      }
    }
    for (; i + kSegWarps < end; i += 2 * kSegWarps) {
      const int64_t r0 = perm ? perm[i] : i, r1 = perm ? perm[i + kSegWarps] : (i + kSegWarps);
      if (VEC == 4) {
        const float4 a = ldg_f4(x + r0 * ldx + f), b = ldg_f4(x + r1 * ldx + f);
        acc[0] += a.x; acc[1 % VEC] += a.y; acc[2 % VEC] += a.z; acc[3 % VEC] += a.w;
        acc[0] += b.x; acc[1 % VEC] += b.y; acc[2 % VEC] += b.z; acc[3 % VEC] += b.w;
      } else {
        const float a = __ldg(x + r0 * ldx + f), b = __ldg(x + r1 * ldx + f);
        acc[0] += a;
        acc[0] += b;
      }
    }
    for (; i < end; i += kSegWarps) {
      const int64_t r0 = perm ? perm[i] : i;
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[v] += __ldg(x + r0 * ldx + f + v);
    }
  }
#pragma unroll
  for (int v = 0; v < VEC; ++v) part[wid][lane * VEC + v] = acc[v];
  __syncthreads();
  if (wid == 0 && on) {
    float out[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kSegWarps; ++w) t += part[w][lane * VEC + v];
      out[v] = mean ? __fdiv_rn(t, (float)max(end - beg, 1)) : t;
    }
    float* yo = y + (int64_t)g * ldy + f;
    if (VEC == 4) *reinterpret_cast<float4*>(yo) = make_float4(out[0], out[1 % VEC], out[2 % VEC], out[3 % VEC]);
    else yo[0] = out[0];
  }
}

// dx[row(p), :] = dy[seg(p), :] / count(seg) (mean) ; one warp-sized column tile per thread group.
template <int VEC>
__global__ void __launch_bounds__(256) segment_broadcast_kernel(const float* __restrict__ dy, int64_t lddy,
                                                                const int* __restrict__ ptr,
                                                                const int* __restrict__ perm, int num_segments,
                                                                int num_feat, int mean, float* __restrict__ dx,
                                                                int64_t lddx) {
  // warp per output row: the segment search runs once per row (not once per element), the lanes stream the row
  const int nvec = num_feat / VEC;
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int total_rows = ptr[num_segments];
  for (int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < total_rows; p += nwarps) {
    int lo = 0, hi = num_segments;  // largest g with ptr[g] <= p
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (ptr[mid] <= p) lo = mid; else hi = mid;
    }
    const float cnt = (float)max(ptr[lo + 1] - ptr[lo], 1);  // autograd of true_divide_: g / count
    const int64_t r = perm ? perm[p] : p;
    const float* src = dy + (int64_t)lo * lddy;
    float* dst = dx + r * lddx;
    for (int v = lane; v < nvec; v += 32) {
      if (VEC == 4) {
        float4 q = ldg_f4(src + v * 4);
        if (mean) {
          q.x = __fdiv_rn(q.x, cnt); q.y = __fdiv_rn(q.y, cnt); q.z = __fdiv_rn(q.z, cnt); q.w = __fdiv_rn(q.w, cnt);
        }
        *reinterpret_cast<float4*>(dst + v * 4) = q;
      } else {
        dst[v] = mean ? __fdiv_rn(__ldg(src + v), cnt) : __ldg(src + v);
      }
    }
  }
}

// ---- column sum (bias gradients: db = sum_rows dY) ---------------------------------------------------
// Stage 1: CTA (rows chunk, 128-column tile) -> partial[chunk, :];  stage 2: fixed-order sum over chunks.
constexpr int kColsumRows = 64;  // rows per CTA of the 128-column tile kernel (and the workspace granularity)
// rows per CTA of the row-coalesced kernel: one wave of CTAs for a 20 k-row batch (128 rows), and a second stage that
// stays short (it walks one partial row per chunk)
__host__ __device__ inline int colsum_chunk_rows(int64_t num_rows) {
  return num_rows > 65536 ? 256 : (num_rows > 16384 ? 128 : 64);
}

template <int VEC>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ x, int64_t ldx,
                                                             int num_rows, int num_feat,
                                                             float* __restrict__ partial,
                                                             const float* __restrict__ mask = nullptr,
                                                             int64_t ldm = 0, float* __restrict__ masked = nullptr,
                                                             int64_t ldo = 0, int* __restrict__ chunk_count = nullptr) {
  __shared__ float part[8][32 * VEC];
  if (chunk_count != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *chunk_count = gridDim.y;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int f = (blockIdx.x * 32 + lane) * VEC;
  const bool on = f < num_feat;
  const int r0 = blockIdx.y * kColsumRows;
  float acc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
  if (on) {
#pragma unroll 4
    for (int k = 0; k < kColsumRows / 8; ++k) {
      const int r = r0 + k * 8 + wid;
      if (r < num_rows) {
        if (VEC == 4) {
          float4 a = ldg_f4(x + (int64_t)r * ldx + f);
          if (mask != nullptr) {                    // ReLU backward folded in: rows of x count where mask > 0
            const float4 m = ldg_f4(mask + (int64_t)r * ldm + f);
            a.x = m.x > 0.f ? a.x : 0.f; a.y = m.y > 0.f ? a.y : 0.f;
            a.z = m.z > 0.f ? a.z : 0.f; a.w = m.w > 0.f ? a.w : 0.f;
            if (masked != nullptr) *reinterpret_cast<float4*>(masked + (int64_t)r * ldo + f) = a;
          }
          acc[0] += a.x; acc[1 % VEC] += a.y; acc[2 % VEC] += a.z; acc[3 % VEC] += a.w;
        } else {
          float a = __ldg(x + (int64_t)r * ldx + f);
          if (mask != nullptr) {
            if (!(__ldg(mask + (int64_t)r * ldm + f) > 0.f)) a = 0.f;
            if (masked != nullptr) masked[(int64_t)r * ldo + f] = a;
          }
          acc[0] += a;
        }
      }
    }
  }
#pragma unroll
  for (int v = 0; v < VEC; ++v) part[wid][lane * VEC + v] = acc[v];
  __syncthreads();
  if (wid == 0 && on) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += part[w][lane * VEC + v];
      partial[(int64_t)blockIdx.y * num_feat + f + v] = t;
    }
  }
}

// Row-coalesced stage 1 for feature widths that are a multiple of 4 (<= 1280 columns): a CTA owns kColsumRows rows, its
// threads tile (rows-per-pass x float4 columns), so every warp reads whole contiguous row segments (the 128-column tile
// kernel above leaves 27 % of its lanes idle at 300 columns).  Optional ReLU mask, optional write of the masked rows.
template <bool MASK, bool WRITE>
__global__ void __launch_bounds__(320) colsum_rows_kernel(const float* __restrict__ x, int64_t ldx, int num_rows,
                                                          int chunk_rows, int f4, int rows_per_pass,
                                                          float* __restrict__ partial,
                                                          const float* __restrict__ mask, int64_t ldm,
                                                          float* __restrict__ masked, int64_t ldo,
                                                          int* __restrict__ chunk_count) {
  extern __shared__ float4 cs_part[];            // [rows_per_pass][f4]
  if (blockIdx.x == 0 && threadIdx.x == 0) *chunk_count = gridDim.x;
  const int rp = threadIdx.x / f4, c4 = threadIdx.x - rp * f4;
  const int r0 = blockIdx.x * chunk_rows, r1 = min(num_rows, r0 + chunk_rows);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rp < rows_per_pass) {
#pragma unroll 4
    for (int r = r0 + rp; r < r1; r += rows_per_pass) {
      float4 a = ldg_f4(x + (int64_t)r * ldx + c4 * 4);
      if (MASK) {
        const float4 m = ldg_f4(mask + (int64_t)r * ldm + c4 * 4);
        a.x = m.x > 0.f ? a.x : 0.f; a.y = m.y > 0.f ? a.y : 0.f;
        a.z = m.z > 0.f ? a.z : 0.f; a.w = m.w > 0.f ? a.w : 0.f;
        if (WRITE) *reinterpret_cast<float4*>(masked + (int64_t)r * ldo + c4 * 4) = a;
      }
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
    cs_part[rp * f4 + c4] = acc;
  }
  __syncthreads();
  if (rp == 0) {
    float4 t = cs_part[c4];
    for (int q = 1; q < rows_per_pass; ++q) {
      const float4 u = cs_part[q * f4 + c4];
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    *reinterpret_cast<float4*>(partial + (int64_t)blockIdx.x * f4 * 4 + c4 * 4) = t;
  }
}

// One CTA per 32 columns: 8 warps stride the chunk partials (4 loads in flight each), fixed-order combine.
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ partial,
                                                           const int* __restrict__ chunk_count, int num_feat,
                                                           float* __restrict__ out) {
  const int num_chunks = *chunk_count;           // written by the first stage (its row-chunk size depends on the kernel)
  __shared__ float part[8][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int f = blockIdx.x * 32 + lane;
  float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
  if (f < num_feat) {
    int c = wid;
    for (; c + 24 < num_chunks; c += 32) {
      t0 += partial[(int64_t)c * num_feat + f];
      t1 += partial[(int64_t)(c + 8) * num_feat + f];
      t2 += partial[(int64_t)(c + 16) * num_feat + f];
      t3 += partial[(int64_t)(c + 24) * num_feat + f];
    }
    for (; c < num_chunks; c += 8) t0 += partial[(int64_t)c * num_feat + f];
  }
  part[wid][lane] = (t0 + t1) + (t2 + t3);
  __syncthreads();
  if (wid == 0 && f < num_feat) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += part[w][lane];
    out[f] = t;
  }
}

// x = hi + lo with hi exactly representable in TF32 (low 13 mantissa bits cleared) and lo = x - hi (exact in fp32).
// Feeds the 3xTF32 tensor-core GEMM scheme: x.w ~= hi.hi + hi.lo + lo.hi, error ~2^-21 relative.
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ x, int64_t n,
                                                         float* __restrict__ hi, float* __restrict__ lo) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = n >> 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = ldg_f4(x + 4 * i);
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
    *reinterpret_cast<float4*>(hi + 4 * i) = h;
    *reinterpret_cast<float4*>(lo + 4 * i) = l;
  }
  for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = x[i];
    const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    hi[i] = h;
    lo[i] = v - h;
  }
}

// K-concatenated split: out[r, :] = [lo | hi | hi] (mode 0) or [hi | lo | hi] (mode 1) of x[r, :], rows
// num_rows..num_rows_padded-1 zero.  With A in mode 0 and W in mode 1, ONE TF32 GEMM over the 3K-long reduction
// computes x_lo.W_hi + x_hi.W_lo + x_hi.W_hi, i.e. the 3xTF32 product, in a single pass over the output.  The
// two small cross terms come FIRST in the reduction so that the tensor core's accumulator is still small when
// they are added (its additions are not correctly rounded; adding them after the big term loses them).
template <int VEC>
__global__ void __launch_bounds__(256) split_tf32_cat_kernel(const float* __restrict__ x, int64_t ldx,
                                                             int64_t num_rows, int64_t num_rows_padded, int K,
                                                             int mode, float* __restrict__ out) {
  const int kv = K / VEC;
  const int64_t total = num_rows_padded * kv;
  const int64_t ldo = 3 * (int64_t)K;
  const int o1 = K, o2 = 2 * K;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / kv;
    const int c = (int)(idx - r * kv) * VEC;
    float v[VEC], h[VEC], l[VEC];
    if (r < num_rows) {
      if (VEC == 4) {
        const float4 q = ldg_f4(x + r * ldx + c);
        v[0] = q.x; v[1 % VEC] = q.y; v[2 % VEC] = q.z; v[3 % VEC] = q.w;
      } else {
        v[0] = __ldg(x + r * ldx + c);
      }
    } else {
#pragma unroll
      for (int k = 0; k < VEC; ++k) v[k] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      h[k] = __uint_as_float(__float_as_uint(v[k]) & 0xffffe000u);
      l[k] = v[k] - h[k];
    }
    float* o = out + r * ldo + c;
    if (VEC == 4) {
      const float4 hh = make_float4(h[0], h[1 % VEC], h[2 % VEC], h[3 % VEC]);
      const float4 ll = make_float4(l[0], l[1 % VEC], l[2 % VEC], l[3 % VEC]);
      *reinterpret_cast<float4*>(o) = mode == 0 ? ll : hh;
      *reinterpret_cast<float4*>(o + o1) = mode == 0 ? hh : ll;
      *reinterpret_cast<float4*>(o + o2) = hh;
    } else {
      o[0] = mode == 0 ? l[0] : h[0];
      o[o1] = mode == 0 ? h[0] : l[0];
      o[o2] = h[0];
    }
  }
}

static inline bool aligned16(const void* a, const void* b) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) % 16) == 0;
}

// ---- device-side collate (loader/loader.py:48-60 -> PyG `Batch.from_data_list`) ----------------------------------------
// The host packs the graphs of a mini-batch back to back WITHOUT touching their indices (edge_index stays local to
// each graph); one CTA per graph then derives what collate adds: the graph's node offset (prefix sum of the node
// counts), `batch` (graph id per node) and the offset edge_index.  Replaces B small tensor adds + concatenations on
// the host per batch.
__global__ void __launch_bounds__(256) collate_batch_kernel(const int* __restrict__ node_counts,
                                                            const int* __restrict__ edge_counts, int num_graphs,
                                                            const long long* __restrict__ ei_local, long long e_cap,
                                                            long long* __restrict__ ei_global,
                                                            long long* __restrict__ batch, long long n_cap) {
  __shared__ int red[32];
  __shared__ int s_noff, s_eoff;
  const int g = blockIdx.x;
  int pn = 0, pe = 0;
  for (int j = threadIdx.x; j < g; j += blockDim.x) { pn += node_counts[j]; pe += edge_counts[j]; }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  pn = warp_sum_i(pn);
  if (lane == 0) red[wid] = pn;
  __syncthreads();
  if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += red[w]; s_noff = t; }
  __syncthreads();
  pe = warp_sum_i(pe);
  if (lane == 0) red[wid] = pe;
  __syncthreads();
  if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += red[w]; s_eoff = t; }
  __syncthreads();
  const long long noff = s_noff, eoff = s_eoff;
  const int n = node_counts[g], e = edge_counts[g];
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    if (noff + i < n_cap) batch[noff + i] = g;
  for (int i = threadIdx.x; i < e; i += blockDim.x) {
    if (eoff + i < e_cap) {
      ei_global[eoff + i] = ei_local[eoff + i] + noff;
      ei_global[e_cap + eoff + i] = ei_local[e_cap + eoff + i] + noff;
    }
  }
}

// ---- ReLU + dropout in one pass (model/mpnn.py:57-58: `F.dropout(self.activation(x), p, training)`) -------------------
// Philox4x32-10 keyed by a (seed, call counter) pair that lives on the device, so a captured CUDA graph draws a fresh
// mask on every replay; the element index is the Philox counter.  No mask tensor: a dropped element and a negative
// pre-activation both leave y = 0, so the backward is  dx = (y > 0) ? dy / (1 - p) : 0  from the output alone.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__global__ void __launch_bounds__(256) relu_dropout_fwd_kernel(const float* __restrict__ x, long long n, float p,
                                                               const unsigned long long* __restrict__ state,
                                                               float* __restrict__ y) {
  const unsigned long long seed = state[0], call = state[1];
  const uint2 key = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
  const float scale = 1.0f / (1.0f - p);
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g * 4 < n;
       g += (long long)gridDim.x * blockDim.x) {
    const uint4 r = philox4x32_10(make_uint4((unsigned)g, (unsigned)(g >> 32), (unsigned)call, (unsigned)(call >> 32)), key);
    const unsigned rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = g * 4 + u;
      if (i < n) {
        const float uni = (float)(rr[u] >> 8) * (1.0f / 16777216.0f);      // [0, 1)
        const float v = fmaxf(x[i], 0.f);
        y[i] = uni >= p ? v * scale : 0.f;
      }
    }
  }
}
__global__ void relu_dropout_bump_kernel(unsigned long long* state) { state[1] += 1ull; }
__global__ void __launch_bounds__(256) relu_dropout_bwd_kernel(const float* __restrict__ dy,
                                                               const float* __restrict__ y, long long n, float scale,
                                                               float* __restrict__ dx) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dx[i] = y[i] > 0.f ? dy[i] * scale : 0.f;
}

}  // namespace ghscn

using namespace ghscn;

extern "C" {

int ghscn_segment_reduce(const float* x, int64_t ldx, const int32_t* ptr, const int32_t* perm, int64_t num_segments,
                         int64_t num_feat, int32_t mean, float* y, int64_t ldy, ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_segments >= 0 && num_feat >= 0);
  if (num_segments == 0 || num_feat == 0) return GHSCN_OK;
  GHSCN_REQUIRE(x && ptr && y && ldx >= num_feat && ldy >= num_feat);
  GHSCN_REQUIRE(num_segments <= 65535);  // grid.y; readout batches are far below this
  const bool vec4 = num_feat % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && aligned16(x, y);
  if (vec4) {
    dim3 grid((unsigned)ceil_div<int64_t>(num_feat, 128), (unsigned)num_segments);
    segment_reduce_kernel<4><<<grid, kSegWarps * 32, 0, as_stream(stream)>>>(x, ldx, ptr, perm, (int)num_feat, mean,
                                                                            y, ldy);
  } else {
    dim3 grid((unsigned)ceil_div<int64_t>(num_feat, 32), (unsigned)num_segments);
    segment_reduce_kernel<1><<<grid, kSegWarps * 32, 0, as_stream(stream)>>>(x, ldx, ptr, perm, (int)num_feat, mean,
                                                                            y, ldy);
  }
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

int ghscn_segment_broadcast(const float* dy, int64_t lddy, const int32_t* ptr, const int32_t* perm,
                            int64_t num_segments, int64_t num_feat, int32_t mean, float* dx, int64_t lddx,
                            ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_segments >= 0 && num_feat >= 0 && num_segments < ((int64_t)1 << 31));
  if (num_segments == 0 || num_feat == 0) return GHSCN_OK;
  GHSCN_REQUIRE(dy && ptr && dx && lddy >= num_feat && lddx >= num_feat);
  const bool vec4 = num_feat % 4 == 0 && lddy % 4 == 0 && lddx % 4 == 0 && aligned16(dy, dx);
  const int blocks = kNumSMs * 8;
  if (vec4)
    segment_broadcast_kernel<4><<<blocks, 256, 0, as_stream(stream)>>>(dy, lddy, ptr, perm, (int)num_segments,
                                                                        (int)num_feat, mean, dx, lddx);
  else
    segment_broadcast_kernel<1><<<blocks, 256, 0, as_stream(stream)>>>(dy, lddy, ptr, perm, (int)num_segments,
                                                                        (int)num_feat, mean, dx, lddx);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

int ghscn_split_tf32(const float* x, int64_t n, float* hi, float* lo, ghscn_stream_t stream) {
  GHSCN_REQUIRE(n >= 0 && (n == 0 || (x && hi && lo)));
  if (n == 0) return GHSCN_OK;
  GHSCN_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(hi) |
                  reinterpret_cast<uintptr_t>(lo)) % 16) == 0);
  const int64_t blocks = ceil_div<int64_t>(ceil_div<int64_t>(n, 4), 256);
  const int64_t cap = (int64_t)kNumSMs * 16;
  split_tf32_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, as_stream(stream)>>>(x, n, hi, lo);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

int ghscn_split_tf32_cat(const float* x, int64_t ldx, int64_t num_rows, int64_t num_rows_padded, int64_t num_cols,
                         int32_t mode, float* out, ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_rows >= 0 && num_rows_padded >= num_rows && num_cols >= 0 && (mode == 0 || mode == 1));
  GHSCN_REQUIRE(num_cols < ((int64_t)1 << 24));
  if (num_rows_padded == 0 || num_cols == 0) return GHSCN_OK;
  GHSCN_REQUIRE(out && (num_rows == 0 || (x && ldx >= num_cols)));
  const bool vec4 = num_cols % 4 == 0 && ldx % 4 == 0 && aligned16(x, out);
  const int64_t work = num_rows_padded * (num_cols / (vec4 ? 4 : 1));
  const int64_t cap = (int64_t)kNumSMs * 16, blocks = ceil_div<int64_t>(work, 256);
  const unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
  if (vec4)
    split_tf32_cat_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(x, ldx, num_rows, num_rows_padded, (int)num_cols,
                                                                   mode, out);
  else
    split_tf32_cat_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(x, ldx, num_rows, num_rows_padded, (int)num_cols,
                                                                   mode, out);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

int ghscn_collate_batch(const int32_t* node_counts, const int32_t* edge_counts, int64_t num_graphs,
                        const int64_t* edge_index_local, int64_t edge_capacity, int64_t* edge_index,
                        int64_t* batch, int64_t node_capacity, ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_graphs >= 0 && num_graphs <= 65535 && edge_capacity >= 0 && node_capacity >= 0);
  if (num_graphs == 0) return GHSCN_OK;
  GHSCN_REQUIRE(node_counts && edge_counts && (edge_capacity == 0 || (edge_index_local && edge_index)) &&
                (node_capacity == 0 || batch));
  collate_batch_kernel<<<(unsigned)num_graphs, 256, 0, as_stream(stream)>>>(
      node_counts, edge_counts, (int)num_graphs, reinterpret_cast<const long long*>(edge_index_local),
      (long long)edge_capacity, reinterpret_cast<long long*>(edge_index), reinterpret_cast<long long*>(batch),
      (long long)node_capacity);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

int ghscn_relu_dropout_fwd(const float* x, int64_t n, float p, uint64_t* state, float* y, ghscn_stream_t stream_) {
  GHSCN_REQUIRE(n >= 0 && p >= 0.f && p < 1.f && state && (n == 0 || (x && y)));
  cudaStream_t stream = as_stream(stream_);
  if (n > 0) {
    const int64_t want = ceil_div<int64_t>(ceil_div<int64_t>(n, 4), 256), cap = (int64_t)kNumSMs * 16;
    relu_dropout_fwd_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, stream>>>(
        x, (long long)n, p, reinterpret_cast<const unsigned long long*>(state), y);
  }
  relu_dropout_bump_kernel<<<1, 1, 0, stream>>>(reinterpret_cast<unsigned long long*>(state));
  GHSCN_LAUNCH_CHECK_N(n > 0 ? 2 : 1);
  return GHSCN_OK;
}

int ghscn_relu_dropout_bwd(const float* dy, const float* y, int64_t n, float p, float* dx, ghscn_stream_t stream) {
  GHSCN_REQUIRE(n >= 0 && p >= 0.f && p < 1.f && (n == 0 || (dy && y && dx)));
  if (n == 0) return GHSCN_OK;
  const int64_t want = ceil_div<int64_t>(n, 256 * 4), cap = (int64_t)kNumSMs * 16;
  relu_dropout_bwd_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, as_stream(stream)>>>(
      dy, y, (long long)n, 1.0f / (1.0f - p), dx);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

// the first stage leaves its number of row chunks behind the partials (inside the 256 bytes of slack)
static inline int* colsum_chunk_count(const void* workspace, int64_t num_rows, int64_t num_feat) {
  const size_t off = (size_t)ceil_div<int64_t>(num_rows > 0 ? num_rows : 1, kColsumRows) * num_feat * 4;
  return reinterpret_cast<int*>(const_cast<char*>(static_cast<const char*>(workspace)) + off);
}

size_t ghscn_colsum_workspace_bytes(int64_t num_rows, int64_t num_feat) {
  if (num_rows < 0 || num_feat < 0) return 0;
  return (size_t)ceil_div<int64_t>(num_rows > 0 ? num_rows : 1, kColsumRows) * num_feat * 4 + 256;
}

int ghscn_colsum(const float* x, int64_t ldx, int64_t num_rows, int64_t num_feat, float* out, void* workspace,
                 size_t workspace_bytes, ghscn_stream_t stream_) {
  return ghscn_colsum_masked(x, ldx, nullptr, 0, num_rows, num_feat, out, workspace, workspace_bytes, stream_);
}

int ghscn_colsum_masked(const float* x, int64_t ldx, const float* mask, int64_t ldm, int64_t num_rows,
                        int64_t num_feat, float* out, void* workspace, size_t workspace_bytes,
                        ghscn_stream_t stream_) {
  const int rc = ghscn_relu_grad_colsum_partial(x, ldx, mask, ldm, num_rows, num_feat, nullptr, 0, workspace,
                                                workspace_bytes, stream_);
  if (rc != GHSCN_OK) return rc;
  return ghscn_colsum_finish(workspace, workspace_bytes, num_rows, num_feat, out, stream_);
}

int ghscn_colsum_finish(const void* workspace, size_t workspace_bytes, int64_t num_rows, int64_t num_feat, float* out,
                        ghscn_stream_t stream_) {
  GHSCN_REQUIRE(num_rows >= 0 && num_feat >= 0 && num_rows < ((int64_t)1 << 31) && num_feat < ((int64_t)1 << 24));
  if (num_feat == 0) return GHSCN_OK;
  GHSCN_REQUIRE(out != nullptr);
  if (workspace_bytes < ghscn_colsum_workspace_bytes(num_rows, num_feat) || workspace == nullptr)
    return GHSCN_E_WORKSPACE;
  colsum_final_kernel<<<(unsigned)ceil_div<int64_t>(num_feat, 32), 256, 0, as_stream(stream_)>>>(
      static_cast<const float*>(workspace), colsum_chunk_count(workspace, num_rows, num_feat), (int)num_feat, out);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

int ghscn_relu_grad_colsum_partial(const float* x, int64_t ldx, const float* mask, int64_t ldm, int64_t num_rows,
                                   int64_t num_feat, float* masked, int64_t ldo, void* workspace,
                                   size_t workspace_bytes, ghscn_stream_t stream_) {
  GHSCN_REQUIRE(num_rows >= 0 && num_feat >= 0 && num_rows < ((int64_t)1 << 31) && num_feat < ((int64_t)1 << 24));
  if (num_feat == 0) return GHSCN_OK;
  GHSCN_REQUIRE(num_rows == 0 || (x && ldx >= num_feat));
  GHSCN_REQUIRE(mask == nullptr || ldm >= num_feat);
  GHSCN_REQUIRE(masked == nullptr || (mask != nullptr && ldo >= num_feat));
  if (workspace_bytes < ghscn_colsum_workspace_bytes(num_rows, num_feat) || workspace == nullptr)
    return GHSCN_E_WORKSPACE;
  cudaStream_t stream = as_stream(stream_);
  float* partial = static_cast<float*>(workspace);
  int* count = colsum_chunk_count(workspace, num_rows, num_feat);
  int chunks = (int)ceil_div<int64_t>(num_rows, kColsumRows);
  if (chunks == 0) {
    cudaError_t e = cudaMemsetAsync(count, 0, sizeof(int), stream);
    if (e != cudaSuccess) return (int)e;
  } else {
    const bool vec4 = num_feat % 4 == 0 && ldx % 4 == 0 && aligned16(x, x) &&
                      (mask == nullptr || (ldm % 4 == 0 && aligned16(mask, mask))) &&
                      (masked == nullptr || (ldo % 4 == 0 && aligned16(masked, masked)));
    if (!(vec4 && num_feat <= 1280) && chunks > 65535) return GHSCN_E_UNSUPPORTED;      // grid.y of the tile kernel
    if (vec4 && num_feat <= 1280) {
      const int f4 = (int)num_feat / 4;
      const int rpp = f4 >= 320 ? 1 : 320 / f4;
      const int threads = (rpp * f4 + 31) / 32 * 32;
      const size_t shm = (size_t)rpp * f4 * sizeof(float4);
      const int cr = colsum_chunk_rows(num_rows);
      chunks = (int)ceil_div<int64_t>(num_rows, cr);
      if (mask == nullptr)
        colsum_rows_kernel<false, false><<<chunks, threads, shm, stream>>>(x, ldx, (int)num_rows, cr, f4, rpp, partial,
                                                                          nullptr, 0, nullptr, 0, count);
      else if (masked == nullptr)
        colsum_rows_kernel<true, false><<<chunks, threads, shm, stream>>>(x, ldx, (int)num_rows, cr, f4, rpp, partial,
                                                                         mask, ldm, nullptr, 0, count);
      else
        colsum_rows_kernel<true, true><<<chunks, threads, shm, stream>>>(x, ldx, (int)num_rows, cr, f4, rpp, partial,
                                                                        mask, ldm, masked, ldo, count);
    } else if (vec4) {
      dim3 grid((unsigned)ceil_div<int64_t>(num_feat, 128), (unsigned)chunks);
      colsum_partial_kernel<4><<<grid, 256, 0, stream>>>(x, ldx, (int)num_rows, (int)num_feat, partial, mask, ldm,
                                                         masked, ldo, count);
    } else {
      dim3 grid((unsigned)ceil_div<int64_t>(num_feat, 32), (unsigned)chunks);
      colsum_partial_kernel<1><<<grid, 256, 0, stream>>>(x, ldx, (int)num_rows, (int)num_feat, partial, mask, ldm,
                                                         masked, ldo, count);
    }
    GHSCN_LAUNCH_CHECK();
  }
  return GHSCN_OK;
}

}  // extern "C"
