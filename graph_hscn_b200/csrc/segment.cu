// K4: segment mean / sum over contiguous row ranges (graph readout) and its broadcast backward.
// Replaces torch_scatter.scatter_mean (model/mpnn.py:60) / global_mean_pool (model/hscn.py:111).
// HBM-bound: 4F(N + B) bytes.  One thread owns one (segment, column-vector) pair and walks the
// segment's rows in order (CPU scatter_add_ order); consecutive threads read consecutive columns so
// every row is fetched with coalesced 128-bit loads; 4 rows are kept in flight per thread.
#include "common.cuh"

namespace ghscn {

template <int VEC>
__global__ void __launch_bounds__(128) segment_reduce_kernel(const float* __restrict__ x, int64_t ldx,
                                                             const int* __restrict__ ptr,
                                                             const int* __restrict__ perm, int num_feat,
                                                             int mean, float* __restrict__ y, int64_t ldy) {
  const int g = blockIdx.y;
  const int f = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (f >= num_feat) return;
  const int beg = ptr[g], end = ptr[g + 1];
  float acc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
  int i = beg;
  for (; i + 3 < end; i += 4) {
    float t[4][VEC];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t r = perm ? perm[i + u] : (i + u);
      if (VEC == 4) {
        const float4 q = ldg_f4(x + r * ldx + f);
        t[u][0] = q.x; t[u][1 % VEC] = q.y; t[u][2 % VEC] = q.z; t[u][3 % VEC] = q.w;
      } else {
        t[u][0] = __ldg(x + r * ldx + f);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[v] = __fadd_rn(acc[v], t[u][v]);
  }
  for (; i < end; ++i) {
    const int64_t r = perm ? perm[i] : i;
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = __fadd_rn(acc[v], __ldg(x + r * ldx + f + v));
  }
  if (mean) {
    const float cnt = (float)max(end - beg, 1);
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = __fdiv_rn(acc[v], cnt);
  }
  float* yo = y + (int64_t)g * ldy + f;
  if (VEC == 4) *reinterpret_cast<float4*>(yo) = make_float4(acc[0], acc[1 % VEC], acc[2 % VEC], acc[3 % VEC]);
  else yo[0] = acc[0];
}

// dx[row(p), :] = dy[seg(p), :] * scale(seg);  grid.y tiles rows, grid.x tiles columns.
template <int VEC>
__global__ void __launch_bounds__(256) segment_broadcast_kernel(const float* __restrict__ dy, int64_t lddy,
                                                                const int* __restrict__ ptr,
                                                                const int* __restrict__ perm, int num_segments,
                                                                int num_feat, int mean, float* __restrict__ dx,
                                                                int64_t lddx) {
  const int nvec = num_feat / VEC;
  const int64_t total = (int64_t)ptr[num_segments] * nvec;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(idx / nvec);
    const int f = (int)(idx % nvec) * VEC;
    int lo = 0, hi = num_segments;  // largest g with ptr[g] <= p
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (ptr[mid] <= p) lo = mid; else hi = mid;
    }
    const float cnt = (float)max(ptr[lo + 1] - ptr[lo], 1);  // autograd of true_divide_: g / count
    const int64_t r = perm ? perm[p] : p;
    const float* src = dy + (int64_t)lo * lddy + f;
    float* dst = dx + r * lddx + f;
    if (VEC == 4) {
      float4 q = ldg_f4(src);
      if (mean) {
        q.x = __fdiv_rn(q.x, cnt); q.y = __fdiv_rn(q.y, cnt); q.z = __fdiv_rn(q.z, cnt); q.w = __fdiv_rn(q.w, cnt);
      }
      *reinterpret_cast<float4*>(dst) = q;
    } else {
      dst[0] = mean ? __fdiv_rn(__ldg(src), cnt) : __ldg(src);
    }
  }
}

static inline bool aligned16(const void* a, const void* b) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) % 16) == 0;
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
    dim3 grid((unsigned)ceil_div<int64_t>(num_feat / 4, 128), (unsigned)num_segments);
    segment_reduce_kernel<4><<<grid, 128, 0, as_stream(stream)>>>(x, ldx, ptr, perm, (int)num_feat, mean, y, ldy);
  } else {
    dim3 grid((unsigned)ceil_div<int64_t>(num_feat, 128), (unsigned)num_segments);
    segment_reduce_kernel<1><<<grid, 128, 0, as_stream(stream)>>>(x, ldx, ptr, perm, (int)num_feat, mean, y, ldy);
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

}  // extern "C"
