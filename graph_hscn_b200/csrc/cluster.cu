// K7: cluster argmax and cluster -> virtual-node construction, bit-exact integer work.
// Replaces `clust.max(1)[1]` (train/train_clustering.py:68) and the per-node Python/numpy loops
// of loader/hetero_data.py:44-86, including the reference's quirks:
//   * cluster ids are remapped to 0..U-1 by sorted unique value            (hetero_data.py:46-51)
//   * node i is bucketed into slot (c_i - 1) with Python negative indexing (hetero_data.py:52-54)
//     so, after empty slots are dropped, virtual row j holds the mean of cluster (j+1) mod U
//   * means are taken in float64 and rounded once to fp32                  (hetero_data.py:55-59,66)
//   * l->v edges are (i, c_i); v->v edges are {(a,b): a+b <= U-1} listed a-major (hetero_data.py:68-86)
#include <math.h>

#include "common.cuh"

namespace ghscn {

constexpr int kMaxK = 1024;

__global__ void cluster_argmax_kernel(const float* __restrict__ s, int64_t lds, int64_t n, int K,
                                      int* __restrict__ cluster) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* r = s + i * lds;
  float best = r[0];
  int arg = 0;
  for (int k = 1; k < K; ++k) {
    const float v = r[k];
    // first maximal value wins; NaN counts as maximal (torch.max semantics)
    if (v > best || (v != v && best == best)) { best = v; arg = k; }
  }
  cluster[i] = arg;
}

template <typename T>
__global__ void __launch_bounds__(256) virtual_build_kernel(const int* __restrict__ cluster,
                                                            const int* __restrict__ ptr, const T* __restrict__ x,
                                                            int64_t ldx, int K, int F,
                                                            int* __restrict__ cluster_remapped,
                                                            int* __restrict__ num_virtual,
                                                            float* __restrict__ virt_x, int dyn_smem_ints) {
  __shared__ int present[kMaxK];
  __shared__ int rank[kMaxK];
  __shared__ int U_s;
  const int g = blockIdx.x, tid = threadIdx.x;
  const int base = ptr[g], n = ptr[g + 1] - base;
  for (int k = tid; k < K; k += blockDim.x) present[k] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) {
    const int c = cluster[base + i];
    if (c >= 0 && c < K) present[c] = 1;
  }
  __syncthreads();
  if (tid == 0) {  // K is tiny: serial exclusive scan == np.unique order
    int run = 0;
    for (int k = 0; k < K; ++k) { rank[k] = run; run += present[k]; }
    U_s = run;
    num_virtual[g] = run;
  }
  __syncthreads();
  const int U = U_s;
  for (int i = tid; i < n; i += blockDim.x) {
    const int c = cluster[base + i];
    cluster_remapped[base + i] = (c >= 0 && c < K) ? rank[c] : 0;
  }
  __syncthreads();
  if (sizeof(T) == 8 && dyn_smem_ints > 0) {
    // integer (OGB atom) features: exact int64 sums with shared-memory atomics, node-parallel; the result does
    // not depend on the order, so it equals numpy's float64 mean of integers bit for bit.
    extern __shared__ unsigned long long sums[];           // [U][F] then counts [U] (as ull)
    unsigned long long* cnts = sums + (size_t)K * F;
    for (int i = tid; i < K * F + K; i += blockDim.x) sums[i] = 0ull;
    __syncthreads();
    for (int item = tid; item < n * F; item += blockDim.x) {
      const int i = item / F, f = item - i * F;
      const int c = cluster_remapped[base + i];
      atomicAdd(&sums[(size_t)c * F + f], (unsigned long long)(long long)x[(int64_t)(base + i) * ldx + f]);
      if (f == 0) atomicAdd(&cnts[c], 1ull);
    }
    __syncthreads();
    for (int item = tid; item < K * F; item += blockDim.x) {
      const int j = item / F, f = item - j * F;
      float r = 0.f;
      if (j < U) {
        const int want = (j + 1) % U;
        r = (float)((double)(long long)sums[(size_t)want * F + f] / (double)(long long)cnts[want]);
      }
      virt_x[((int64_t)g * K + j) * F + f] = r;
    }
    return;
  }
  // virtual row j <- mean over nodes whose remapped cluster is (j+1) mod U, in node order
  for (int item = tid; item < K * F; item += blockDim.x) {
    const int j = item / F, f = item - j * F;
    float r = 0.f;
    if (j < U) {
      const int want = (j + 1) % U;
      double sum = 0.0;
      int cnt = 0;
      for (int i = 0; i < n; ++i) {
        if (cluster_remapped[base + i] == want) {
          sum += (double)x[(int64_t)(base + i) * ldx + f];
          ++cnt;
        }
      }
      r = (float)(sum / (double)cnt);
    }
    virt_x[((int64_t)g * K + j) * F + f] = r;
  }
}

__global__ void __launch_bounds__(1024) virtual_offsets_kernel(const int* __restrict__ num_virtual, int B,
                                                               int* __restrict__ virt_offset,
                                                               int* __restrict__ vv_offset) {
  // single CTA, chunked serial-in-thread + warp/block scan; B is the batch size (<= a few thousand)
  __shared__ int tot_a[32], tot_b[32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int per = ceil_div(B, 1024);
  const int beg = min(B, tid * per), end = min(B, beg + per);
  int sa = 0, sb = 0;
  for (int g = beg; g < end; ++g) { const int u = num_virtual[g]; sa += u; sb += u * (u + 1) / 2; }
  int ia = sa, ib = sb;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int va = __shfl_up_sync(kFullMask, ia, o), vb = __shfl_up_sync(kFullMask, ib, o);
    if (lane >= o) { ia += va; ib += vb; }
  }
  if (lane == 31) { tot_a[wid] = ia; tot_b[wid] = ib; }
  __syncthreads();
  if (wid == 0) {
    int ta = tot_a[lane], tb = tot_b[lane], xa = ta, xb = tb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int va = __shfl_up_sync(kFullMask, xa, o), vb = __shfl_up_sync(kFullMask, xb, o);
      if (lane >= o) { xa += va; xb += vb; }
    }
    tot_a[lane] = xa - ta;
    tot_b[lane] = xb - tb;
  }
  __syncthreads();
  int ra = tot_a[wid] + ia - sa, rb = tot_b[wid] + ib - sb;
  for (int g = beg; g < end; ++g) {
    const int u = num_virtual[g];
    virt_offset[g] = ra;
    vv_offset[g] = rb;
    ra += u;
    rb += u * (u + 1) / 2;
  }
  if (end == B && beg < end || (B == 0 && tid == 0) ) { virt_offset[B] = ra; vv_offset[B] = rb; }
}

__global__ void __launch_bounds__(256) virtual_edges_kernel(const int* __restrict__ cluster_remapped,
                                                            const int* __restrict__ ptr,
                                                            const int* __restrict__ num_virtual,
                                                            const int* __restrict__ virt_offset,
                                                            const int* __restrict__ vv_offset, int K,
                                                            int64_t num_nodes, int64_t* __restrict__ lv,
                                                            int64_t* __restrict__ vv, int64_t vv_cap) {
  const int g = blockIdx.x, tid = threadIdx.x;
  const int base = ptr[g], n = ptr[g + 1] - base;
  const int U = num_virtual[g];
  const int64_t voff = virt_offset ? virt_offset[g] : (int64_t)g * K;
  for (int i = tid; i < n; i += blockDim.x) {
    lv[base + i] = base + i;
    lv[num_nodes + base + i] = voff + cluster_remapped[base + i];
  }
  const int64_t eoff = vv_offset ? vv_offset[g] : (int64_t)g * (K * (K + 1) / 2);
  const int cap = vv_offset ? U * (U + 1) / 2 : K * (K + 1) / 2;
  // position p enumerates (a, b): a = 0..U-1, b = 0..U-1-a; row 0 = a ("col" list), row 1 = b ("row" list)
  for (int p = tid; p < cap; p += blockDim.x) {
    int64_t r0 = -1, r1 = -1;
    if (p < U * (U + 1) / 2) {
      int a = 0, rem = p;
      while (rem >= U - a) { rem -= U - a; ++a; }
      r0 = voff + a;
      r1 = voff + rem;
    }
    if (eoff + p < vv_cap) {
      vv[eoff + p] = r0;
      vv[vv_cap + eoff + p] = r1;
    }
  }
}

__global__ void virtual_compact_kernel(const float* __restrict__ virt_x_padded,
                                       const int* __restrict__ num_virtual, const int* __restrict__ virt_offset,
                                       int K, int F, float* __restrict__ virt_x, int64_t* __restrict__ virt_batch) {
  const int g = blockIdx.x;
  const int U = num_virtual[g];
  const int off = virt_offset[g];
  for (int item = threadIdx.x; item < U * F; item += blockDim.x) {
    const int j = item / F, f = item - j * F;
    virt_x[(int64_t)(off + j) * F + f] = virt_x_padded[((int64_t)g * K + j) * F + f];
  }
  if (virt_batch)
    for (int j = threadIdx.x; j < U; j += blockDim.x) virt_batch[off + j] = g;
}

// Both orientations of the l->v and v->v relations straight from the cluster assignment: no sort.
//   l->v  by source: node i owns exactly slot i;  by destination: members of each virtual node in node order
//   v->v  edge p = (a, b), a-major, b = 0..U-1-a: by source slots are consecutive in p; by destination row b
//         holds sources a = 0..U-1-b in that order (closed-form offsets b*U - b(b-1)/2)
__global__ void __launch_bounds__(256) virtual_csr_kernel(
    const int* __restrict__ cluster_remapped, const int* __restrict__ ptr, const int* __restrict__ num_virtual,
    const int* __restrict__ virt_offset, const int* __restrict__ vv_offset, int B, int K, int padded,
    int* __restrict__ lvd_rowptr, int* __restrict__ lvd_col, int* __restrict__ lvd_perm,
    int* __restrict__ lvs_rowptr, int* __restrict__ lvs_col, int* __restrict__ lvs_perm,
    int* __restrict__ vvd_rowptr, int* __restrict__ vvd_col, int* __restrict__ vvd_perm,
    int* __restrict__ vvs_rowptr, int* __restrict__ vvs_col, int* __restrict__ vvs_perm) {
  __shared__ int cnt[kMaxK];
  __shared__ int start[kMaxK + 1];
  const int g = blockIdx.x, tid = threadIdx.x;
  const int base = ptr[g], n = ptr[g + 1] - base;
  const int U = num_virtual[g];
  const int voff = padded ? g * K : virt_offset[g];
  const int rows = padded ? K : U;
  const int eslot = vv_offset[g];                                  // compact slot offset in both layouts
  const int eid0 = padded ? g * (K * (K + 1) / 2) : vv_offset[g];  // id of the graph's first v->v edge
  for (int k = tid; k < K; k += blockDim.x) cnt[k] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) atomicAdd(&cnt[cluster_remapped[base + i]], 1);
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int k = 0; k < K; ++k) { start[k] = run; run += cnt[k]; }
    start[K] = run;
  }
  __syncthreads();
  for (int j = tid; j < rows; j += blockDim.x) {
    lvd_rowptr[voff + j] = base + start[j];
    const int jj = min(j, U);
    vvd_rowptr[voff + j] = eslot + jj * U - jj * (jj - 1) / 2;
    vvs_rowptr[voff + j] = eslot + jj * U - jj * (jj - 1) / 2;
  }
  if (g == B - 1 && tid == 0) {
    const int vrows = padded ? B * K : virt_offset[B];
    lvd_rowptr[vrows] = ptr[B];
    vvd_rowptr[vrows] = vv_offset[B];
    vvs_rowptr[vrows] = vv_offset[B];
    lvs_rowptr[ptr[B]] = ptr[B];
  }
  // l->v by source (one slot per node) and by destination (stable: node order inside each cluster)
  for (int i = tid; i < n; i += blockDim.x) {
    lvs_rowptr[base + i] = base + i;
    lvs_col[base + i] = voff + cluster_remapped[base + i];
    lvs_perm[base + i] = base + i;
  }
  for (int j = tid; j < U; j += blockDim.x) {
    int o = base + start[j];
    for (int i = 0; i < n; ++i)
      if (cluster_remapped[base + i] == j) { lvd_col[o] = base + i; lvd_perm[o] = base + i; ++o; }
  }
  const int ne = U * (U + 1) / 2;
  for (int p = tid; p < ne; p += blockDim.x) {
    int a = 0, b = p;
    while (b >= U - a) { b -= U - a; ++a; }
    vvs_col[eslot + p] = voff + b;
    vvs_perm[eslot + p] = eid0 + p;
    const int o = eslot + b * U - b * (b - 1) / 2 + a;
    vvd_col[o] = voff + a;
    vvd_perm[o] = eid0 + p;
  }
}

}  // namespace ghscn

using namespace ghscn;

extern "C" {

int ghscn_cluster_argmax(const float* s_soft, int64_t lds, int64_t num_nodes, int64_t num_clusters, int32_t* cluster,
                         ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_nodes >= 0 && num_clusters > 0 && lds >= num_clusters);
  if (num_nodes == 0) return GHSCN_OK;
  GHSCN_REQUIRE(s_soft && cluster);
  cluster_argmax_kernel<<<(unsigned)ceil_div<int64_t>(num_nodes, 256), 256, 0, as_stream(stream)>>>(
      s_soft, lds, num_nodes, (int)num_clusters, cluster);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

int ghscn_virtual_build(const int32_t* cluster, const int32_t* ptr, const void* x_raw, int32_t x_is_int64,
                        int64_t ldx, int64_t num_graphs, int64_t num_clusters, int64_t num_feat,
                        int32_t* cluster_remapped, int32_t* num_virtual, float* virt_x_padded,
                        ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_graphs >= 0 && num_clusters > 0 && num_feat >= 0);
  if (num_clusters > kMaxK) return GHSCN_E_UNSUPPORTED;
  if (num_graphs == 0) return GHSCN_OK;
  GHSCN_REQUIRE(cluster && ptr && x_raw && cluster_remapped && num_virtual && virt_x_padded && ldx >= num_feat);
  if (x_is_int64) {
    size_t dyn = ((size_t)num_clusters * num_feat + num_clusters) * 8;
    if (dyn > 40 * 1024) dyn = 0;  // too many cluster x feature sums for shared memory: sequential path
    virtual_build_kernel<int64_t><<<(unsigned)num_graphs, 256, dyn, as_stream(stream)>>>(
        cluster, ptr, static_cast<const int64_t*>(x_raw), ldx, (int)num_clusters, (int)num_feat, cluster_remapped,
        num_virtual, virt_x_padded, (int)(dyn / 4));
  } else {
    virtual_build_kernel<float><<<(unsigned)num_graphs, 256, 0, as_stream(stream)>>>(
        cluster, ptr, static_cast<const float*>(x_raw), ldx, (int)num_clusters, (int)num_feat, cluster_remapped,
        num_virtual, virt_x_padded, 0);
  }
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

int ghscn_virtual_offsets(const int32_t* num_virtual, int64_t num_graphs, int32_t* virt_offset, int32_t* vv_offset,
                          ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_graphs >= 0 && virt_offset && vv_offset && (num_graphs == 0 || num_virtual));
  GHSCN_REQUIRE(num_graphs < ((int64_t)1 << 30));
  virtual_offsets_kernel<<<1, 1024, 0, as_stream(stream)>>>(num_virtual, (int)num_graphs, virt_offset, vv_offset);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

int ghscn_virtual_edges(const int32_t* cluster_remapped, const int32_t* ptr, const int32_t* num_virtual,
                        const int32_t* virt_offset, int64_t num_graphs, int64_t num_nodes, int64_t num_clusters,
                        int64_t* lv_edge_index, int64_t* vv_edge_index, int64_t vv_cap, const int32_t* vv_offset,
                        ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_graphs >= 0 && num_nodes >= 0 && num_clusters > 0 && vv_cap >= 0);
  if (num_graphs == 0) return GHSCN_OK;
  GHSCN_REQUIRE(cluster_remapped && ptr && num_virtual && lv_edge_index && (vv_cap == 0 || vv_edge_index));
  GHSCN_REQUIRE((virt_offset == nullptr) == (vv_offset == nullptr));
  virtual_edges_kernel<<<(unsigned)num_graphs, 256, 0, as_stream(stream)>>>(
      cluster_remapped, ptr, num_virtual, virt_offset, vv_offset, (int)num_clusters, num_nodes, lv_edge_index,
      vv_edge_index, vv_cap);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

int ghscn_virtual_csr(const int32_t* cluster_remapped, const int32_t* ptr, const int32_t* num_virtual,
                      const int32_t* virt_offset, const int32_t* vv_offset, int64_t num_graphs,
                      int64_t num_clusters, int32_t padded, int32_t* lvd_rowptr, int32_t* lvd_col,
                      int32_t* lvd_perm, int32_t* lvs_rowptr, int32_t* lvs_col, int32_t* lvs_perm,
                      int32_t* vvd_rowptr, int32_t* vvd_col, int32_t* vvd_perm, int32_t* vvs_rowptr,
                      int32_t* vvs_col, int32_t* vvs_perm, ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_graphs >= 0 && num_clusters > 0);
  if (num_clusters > kMaxK) return GHSCN_E_UNSUPPORTED;
  if (num_graphs == 0) return GHSCN_OK;
  GHSCN_REQUIRE(cluster_remapped && ptr && num_virtual && vv_offset && (padded || virt_offset));
  GHSCN_REQUIRE(lvd_rowptr && lvd_col && lvd_perm && lvs_rowptr && lvs_col && lvs_perm);
  GHSCN_REQUIRE(vvd_rowptr && vvd_col && vvd_perm && vvs_rowptr && vvs_col && vvs_perm);
  virtual_csr_kernel<<<(unsigned)num_graphs, 256, 0, as_stream(stream)>>>(
      cluster_remapped, ptr, num_virtual, virt_offset, vv_offset, (int)num_graphs, (int)num_clusters, padded,
      lvd_rowptr, lvd_col, lvd_perm, lvs_rowptr, lvs_col, lvs_perm, vvd_rowptr, vvd_col, vvd_perm, vvs_rowptr,
      vvs_col, vvs_perm);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

int ghscn_virtual_compact(const float* virt_x_padded, const int32_t* num_virtual, const int32_t* virt_offset,
                          int64_t num_graphs, int64_t num_clusters, int64_t num_feat, float* virt_x,
                          int64_t* virt_batch, ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_graphs >= 0 && num_clusters > 0 && num_feat >= 0);
  if (num_graphs == 0) return GHSCN_OK;
  GHSCN_REQUIRE(virt_x_padded && num_virtual && virt_offset && virt_x);
  virtual_compact_kernel<<<(unsigned)num_graphs, 128, 0, as_stream(stream)>>>(
      virt_x_padded, num_virtual, virt_offset, (int)num_clusters, (int)num_feat, virt_x, virt_batch);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

}  // extern "C"
