// K1: CSR construction by a stable LSD radix sort over (key, item id) + boundary scan,
// K1b: gcn_norm folded onto the CSR.  See include/ghscn.h for the contract and the
// reference call sites (train/train_clustering.py:37-42, GCNConv.forward).
//
// HBM-bound integer work: nothing here is GEMM-shaped.  One radix pass = histogram
// (smem atomics) -> single-CTA exclusive scan of the [256 x tiles] matrix -> stable scatter
// (warp match-any ranking, per-warp digit counters in shared memory).  Stability of every pass
// makes the final order "by key, then by original edge id", i.e. torch.sort(stable=True).
#include <atomic>

#include "common.cuh"

namespace ghscn {

static std::atomic<unsigned long long> g_launches{0};
void note_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kItemsPerThread = 8;
constexpr int kTile = kSortThreads * kItemsPerThread;  // 2048 items per CTA
constexpr int kItemsPerWarp = 32 * kItemsPerThread;

struct SortSource {
  const int64_t* key;    // [E]
  const int64_t* other;  // [E]
  int64_t num_edges;
  int32_t num_rows;      // sentinel key value == num_rows
  int32_t add_loops;
};

// Key of item t for the first pass.  Items >= E are the appended self loops.
__device__ __forceinline__ int first_pass_key(const SortSource& src, int64_t t) {
  if (t >= src.num_edges) return (int)(t - src.num_edges);
  const int64_t k = src.key[t];
  if (k < 0 || k >= src.num_rows) return src.num_rows;
  if (src.add_loops && k == src.other[t]) return src.num_rows;
  return (int)k;
}

template <bool FIRST>
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(SortSource src, const int* __restrict__ keys_in,
                                                                  int64_t num_items, int shift,
                                                                  int* __restrict__ hist, int num_tiles) {
  __shared__ int h[kRadix];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kTile;
#pragma unroll
  for (int r = 0; r < kItemsPerThread; ++r) {
    const int64_t t = base + r * kSortThreads + threadIdx.x;
    if (t < num_items) {
      const int k = FIRST ? first_pass_key(src, t) : keys_in[t];
      atomicAdd(&h[(k >> shift) & (kRadix - 1)], 1);
    }
  }
  __syncthreads();
  hist[threadIdx.x * num_tiles + blockIdx.x] = h[threadIdx.x];
}

// In-place exclusive scan of n ints by one CTA (n = 256 * tiles; tiles <= a few thousand).
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(int* __restrict__ data, int n) {
  __shared__ int warp_tot[32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int per_thread = ceil_div(n, 1024);
  const int beg = min(n, tid * per_thread), end = min(n, beg + per_thread);
  int sum = 0;
  for (int i = beg; i < end; ++i) sum += data[i];
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(kFullMask, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int t = warp_tot[lane];
    int ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(kFullMask, ti, o);
      if (lane >= o) ti += v;
    }
    warp_tot[lane] = ti - t;
  }
  __syncthreads();
  int run = warp_tot[wid] + incl - sum;
  for (int i = beg; i < end; ++i) {
    const int v = data[i];
    data[i] = run;
    run += v;
  }
}

template <bool FIRST>
__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(SortSource src, const int* __restrict__ keys_in,
                                                                     const int* __restrict__ vals_in,
                                                                     int64_t num_items, int shift,
                                                                     const int* __restrict__ hist_scanned,
                                                                     int num_tiles, int* __restrict__ keys_out,
                                                                     int* __restrict__ vals_out) {
  __shared__ int warp_cnt[kSortWarps][kRadix];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int i = tid; i < kSortWarps * kRadix; i += kSortThreads) (&warp_cnt[0][0])[i] = 0;
  __syncthreads();

  int keys[kItemsPerThread], vals[kItemsPerThread], rank[kItemsPerThread];
  const int64_t warp_base = (int64_t)blockIdx.x * kTile + (int64_t)wid * kItemsPerWarp;
  const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < kItemsPerThread; ++r) {
    const int64_t t = warp_base + r * 32 + lane;
    const bool valid = t < num_items;
    int k = 0, v = 0;
    if (valid) {
      k = FIRST ? first_pass_key(src, t) : keys_in[t];
      v = FIRST ? (int)t : vals_in[t];
    }
    keys[r] = k;
    vals[r] = v;
    const int d = valid ? ((k >> shift) & (kRadix - 1)) : kRadix;  // invalid lanes form their own group
    const unsigned peers = __match_any_sync(kFullMask, d);
    const int in_group = __popc(peers & lt_mask);
    int base = 0;
    if (valid) base = warp_cnt[wid][d];
    __syncwarp();
    if (valid && in_group == 0) warp_cnt[wid][d] = base + __popc(peers);
    __syncwarp();
    rank[r] = base + in_group;
  }
  __syncthreads();
  {  // one thread per digit: exclusive scan over the CTA's warps, seeded with the global base
    const int d = tid;
    int run = hist_scanned[d * num_tiles + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const int c = warp_cnt[w][d];
      warp_cnt[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kItemsPerThread; ++r) {
    const int64_t t = warp_base + r * 32 + lane;
    if (t < num_items) {
      const int d = (keys[r] >> shift) & (kRadix - 1);
      const int pos = warp_cnt[wid][d] + rank[r];
      keys_out[pos] = keys[r];
      vals_out[pos] = vals[r];
    }
  }
}

// Sorted (key, item) -> rowptr by boundary detection (no atomics, no scan), col, perm.
__global__ void csr_finalize_kernel(const int* __restrict__ skeys, const int* __restrict__ sperm,
                                    const int64_t* __restrict__ other, int64_t num_edges, int num_rows,
                                    int64_t num_items, int* __restrict__ rowptr, int* __restrict__ col) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (num_items == 0) {
    for (int64_t r = s; r <= num_rows; r += (int64_t)gridDim.x * blockDim.x) rowptr[r] = 0;
    return;
  }
  if (s >= num_items) return;
  const int k = skeys[s];
  const int v = sperm[s];
  col[s] = (k >= num_rows) ? -1 : (v < num_edges ? (int)other[v] : (int)(v - num_edges));
  const int prev = (s == 0) ? -1 : skeys[s - 1];
  if (k != prev)
    for (int r = prev + 1; r <= k; ++r) rowptr[r] = (int)s;
  if (s == num_items - 1)
    for (int r = k + 1; r <= num_rows; ++r) rowptr[r] = (int)num_items;
}

__global__ void batch_to_ptr_kernel(const int64_t* __restrict__ batch, int64_t n, int num_graphs,
                                    int* __restrict__ ptr) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n == 0) {
    for (int64_t g = i; g <= num_graphs; g += (int64_t)gridDim.x * blockDim.x) ptr[g] = 0;
    return;
  }
  if (i >= n) return;
  const int b = (int)min((int64_t)num_graphs, max((int64_t)0, batch[i]));
  const int prev = (i == 0) ? -1 : (int)min((int64_t)num_graphs, max((int64_t)0, batch[i - 1]));
  if (b != prev)
    for (int g = prev + 1; g <= b; ++g) ptr[g] = (int)i;
  if (i == n - 1)
    for (int g = b + 1; g <= num_graphs; ++g) ptr[g] = (int)n;
}

// ---- K1b ---------------------------------------------------------------------------------
__device__ __forceinline__ float slot_weight(int item, int row, const float* __restrict__ ew,
                                             const float* __restrict__ loop_w, int64_t num_edges) {
  if (item < num_edges) return ew ? ew[item] : 1.0f;
  return loop_w ? loop_w[row] : 1.0f;
}

__global__ void gcn_dis_kernel(const int* __restrict__ rowptr, const int* __restrict__ perm,
                               const float* __restrict__ ew, const float* __restrict__ loop_w, int64_t num_edges,
                               int num_rows, float* __restrict__ dis) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= num_rows) return;
  const int beg = rowptr[r], end = rowptr[r + 1];
  float deg;
  if (ew == nullptr && loop_w == nullptr) {
    deg = (float)(end - beg);  // == sequential sum of ones (exact below 2^24)
  } else {
    deg = 0.f;
    for (int s = beg; s < end; ++s) deg = __fadd_rn(deg, slot_weight(perm[s], r, ew, loop_w, num_edges));
  }
  float d = __fdiv_rn(1.0f, __fsqrt_rn(deg));  // ATen CPU pow(-0.5) == 1/sqrt(x), both IEEE-rounded
  if (d == __int_as_float(0x7f800000)) d = 0.f;  // masked_fill_(== +inf, 0)
  dis[r] = d;
}

__global__ void edge_weights_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                                    const int* __restrict__ perm, const float* __restrict__ ew,
                                    const float* __restrict__ loop_w, const float* __restrict__ dis,
                                    int64_t num_edges, int num_rows, int normalize, int rows_are_dst,
                                    float* __restrict__ w) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= num_rows) return;
  const int beg = rowptr[r], end = rowptr[r + 1];
  const float dr = normalize ? dis[r] : 1.f;
  for (int s = beg; s < end; ++s) {
    const int item = perm[s];
    // appended loop of node i has item id E + i and i == r == col[s] in both orientations
    float v = slot_weight(item, r, ew, loop_w, num_edges);
    if (normalize) {
      const float dc = dis[col[s]];
      const float d_src = rows_are_dst ? dc : dr;
      const float d_dst = rows_are_dst ? dr : dc;
      v = __fmul_rn(__fmul_rn(d_src, v), d_dst);  // deg_inv_sqrt[row] * w * deg_inv_sqrt[col]
    }
    w[s] = v;
  }
}

__global__ void fill_i32_kernel(int* p, int64_t n, int v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void loop_last_kernel(const int64_t* __restrict__ row, const int64_t* __restrict__ colidx,
                                 int64_t num_edges, int num_rows, int* __restrict__ last) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= num_edges) return;
  const int64_t r = row[e];
  if (r == colidx[e] && r >= 0 && r < num_rows) atomicMax(&last[r], (int)e);
}
__global__ void loop_weight_kernel(const int* __restrict__ last, const float* __restrict__ ew, int num_rows,
                                   float fill, float* __restrict__ loop_w) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= num_rows) return;
  const int e = last[r];
  loop_w[r] = (e >= 0 && ew) ? ew[e] : fill;
}

__global__ void cast_i64_f32_kernel(const int64_t* __restrict__ in, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// CSR of (edges + one appended loop per row) from the plain CSR of the same edges; valid when the edge list
// holds no self loop (nothing to drop).  The loop (r,r) has item id E + r and is the last slot of row r,
// exactly where the stable sort of PyG's appended edge list puts it.
__global__ void csr_add_loops_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                                     const int* __restrict__ perm, int num_rows, int num_edges,
                                     int* __restrict__ rowptr2, int* __restrict__ col2, int* __restrict__ perm2) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > num_rows) return;
  if (r == num_rows) { rowptr2[r] = rowptr[r] + r; return; }
  const int beg = rowptr[r], end = rowptr[r + 1];
  int o = beg + r;
  rowptr2[r] = o;
  for (int s = beg; s < end; ++s, ++o) { col2[o] = col[s]; perm2[o] = perm[s]; }
  col2[o] = r;
  perm2[o] = num_edges + r;
}


// ---- K1 fast path: block-diagonal edge lists (PyG collate output) ---------------------------------------------
// A collated mini-batch stores its edges graph-major and no edge leaves its graph, so the stable sort
// by (key, edge id) decomposes into independent per-graph counting sorts.  One CTA per graph builds BOTH
// orientations in shared memory: 32-ary search for the graph's edge range -> per-node counts (smem atomics) ->
// block scan -> unordered bucket fill -> rank of every edge inside its bucket by edge id (what stability means) ->
// final slots.  Output is bit-identical to two ghscn_csr_build calls (tests/test_gpu_structures.py).
constexpr int kBlkThreads = 256;

// First e in [0, num_edges) with src[e] >= target (src is graph-major, so the predicate is monotone); one warp.
__device__ int64_t warp_lower_bound(const int64_t* __restrict__ src, int64_t num_edges, int64_t target) {
  const int lane = threadIdx.x & 31;
  int64_t lo = 0, hi = num_edges;
  while (hi > lo) {
    const int64_t span = hi - lo, step = (span + 31) / 32;
    const int64_t probe = lo + min(span, (int64_t)(lane + 1) * step) - 1;
    const int below = __popc(__ballot_sync(kFullMask, src[probe] < target));   // monotone: a prefix of the lanes
    const int64_t new_lo = lo + min(span, (int64_t)below * step);
    if (below < 32) hi = lo + min(span, (int64_t)(below + 1) * step) - 1;
    lo = new_lo;
  }
  return lo;
}

// Exclusive scan of cnt[0..n) into off[0..n] (off[n] = total) by the whole CTA; `wtot` holds >= 32 ints.
__device__ void block_exclusive_scan(const int* __restrict__ cnt, int* __restrict__ off, int n, int* wtot) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  const int per = ceil_div(n, (int)blockDim.x);
  const int beg = min(n, tid * per), end = min(n, beg + per);
  int sum = 0;
  for (int i = beg; i < end; ++i) sum += cnt[i];
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(kFullMask, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) wtot[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    const int t = lane < nw ? wtot[lane] : 0;
    int ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(kFullMask, ti, o);
      if (lane >= o) ti += v;
    }
    wtot[lane] = ti - t;
  }
  __syncthreads();
  int run = wtot[wid] + incl - sum;
  for (int i = beg; i < end; ++i) { off[i] = run; run += cnt[i]; }
  if (tid == blockDim.x - 1) off[n] = run;
  __syncthreads();
}

__global__ void __launch_bounds__(kBlkThreads)
csr_blocked_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t num_edges,
                   const int* __restrict__ ptr, int num_graphs, int num_nodes, int max_nodes, int max_edges,
                   int* __restrict__ rowptr_d, int* __restrict__ col_d, int* __restrict__ perm_d,
                   int* __restrict__ rowptr_s, int* __restrict__ col_s, int* __restrict__ perm_s,
                   int* __restrict__ status) {
  extern __shared__ int blk_sm[];
  int* off_d = blk_sm;                    // [max_nodes + 1] row offsets by destination (graph-local)
  int* off_s = off_d + max_nodes + 1;     // [max_nodes + 1] ... by source
  int* cur_d = off_s + max_nodes + 1;     // [max_nodes] counts, then fill cursors
  int* cur_s = cur_d + max_nodes;
  int* key_d = cur_s + max_nodes;         // [max_edges] local destination of every edge of the graph
  int* key_s = key_d + max_edges;         // [max_edges] local source
  int* tmp_d = key_s + max_edges;         // [max_edges] bucket-ordered (unstable) local edge ids
  int* tmp_s = tmp_d + max_edges;
  __shared__ int wtot[32];
  __shared__ int64_t range[2];

  const int g = blockIdx.x, tid = threadIdx.x, warp = tid >> 5;
  const int n0 = ptr[g], n1 = ptr[g + 1], n = n1 - n0;
  if (warp == 0) {
    const int64_t v = warp_lower_bound(src, num_edges, n0);
    if (tid == 0) range[0] = v;
  } else if (warp == 1) {
    const int64_t v = warp_lower_bound(src, num_edges, n1);
    if (tid == 32) range[1] = v;
  }
  for (int i = tid; i < n && i < max_nodes; i += kBlkThreads) { cur_d[i] = 0; cur_s[i] = 0; }
  __syncthreads();
  const int64_t e0 = range[0];
  const int ne = (int)(range[1] - e0);
  if (g == num_graphs - 1)                 // rows after the last graph (none in a collated batch) are empty
    for (int r = n1 + tid; r <= num_nodes; r += kBlkThreads) { rowptr_d[r] = (int)num_edges; rowptr_s[r] = (int)num_edges; }
  if (n > max_nodes || ne > max_edges || n < 0) {      // the caller's bounds do not hold: report, write nothing
    if (tid == 0) atomicOr(status, 1);
    return;
  }
  bool bad = false;
  for (int e = tid; e < ne; e += kBlkThreads) {
    const int64_t s = src[e0 + e] - n0, d = dst[e0 + e] - n0;
    const bool ok = s >= 0 && s < n && d >= 0 && d < n;          // an edge leaving its graph: not block diagonal
    bad |= !ok;
    key_s[e] = ok ? (int)s : -1;
    key_d[e] = ok ? (int)d : -1;
    if (ok) { atomicAdd(&cur_s[(int)s], 1); atomicAdd(&cur_d[(int)d], 1); }
  }
  if (bad) atomicOr(status, 2);
  __syncthreads();
  block_exclusive_scan(cur_d, off_d, n, wtot);
  block_exclusive_scan(cur_s, off_s, n, wtot);
  for (int i = tid; i < n; i += kBlkThreads) {
    cur_d[i] = off_d[i];
    cur_s[i] = off_s[i];
    rowptr_d[n0 + i] = (int)e0 + off_d[i];
    rowptr_s[n0 + i] = (int)e0 + off_s[i];
  }
  __syncthreads();
  for (int e = tid; e < ne; e += kBlkThreads) {
    const int d = key_d[e], s = key_s[e];
    if (d >= 0) { tmp_d[atomicAdd(&cur_d[d], 1)] = e; tmp_s[atomicAdd(&cur_s[s], 1)] = e; }
  }
  __syncthreads();
  const int placed = off_d[n];            // == ne unless an edge was rejected
  for (int t = tid; t < placed; t += kBlkThreads) {
    {
      const int e = tmp_d[t], r = key_d[e], a = off_d[r], b = off_d[r + 1];
      int rank = 0;
      for (int u = a; u < b; ++u) rank += (tmp_d[u] < e) ? 1 : 0;
      const int64_t o = e0 + a + rank;
      perm_d[o] = (int)e0 + e;
      col_d[o] = n0 + key_s[e];
    }
    {
      const int e = tmp_s[t], r = key_s[e], a = off_s[r], b = off_s[r + 1];
      int rank = 0;
      for (int u = a; u < b; ++u) rank += (tmp_s[u] < e) ? 1 : 0;
      const int64_t o = e0 + a + rank;
      perm_s[o] = (int)e0 + e;
      col_s[o] = n0 + key_d[e];
    }
  }
}

}  // namespace ghscn

using namespace ghscn;

extern "C" {

int ghscn_abi_version(void) { return 2; }

unsigned long long ghscn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char* ghscn_error_string(int code) {
  switch (code) {
    case GHSCN_OK: return "ok";
    case GHSCN_E_INVALID: return "invalid argument";
    case GHSCN_E_WORKSPACE: return "workspace too small";
    case GHSCN_E_UNSUPPORTED: return "unsupported shape";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
  }
}

size_t ghscn_csr_workspace_bytes(int64_t num_edges, int64_t num_rows, int32_t add_self_loops) {
  if (num_edges < 0 || num_rows < 0) return 0;
  const int64_t m = num_edges + (add_self_loops ? num_rows : 0);
  const int64_t tiles = ceil_div<int64_t>(m > 0 ? m : 1, kTile);
  return 3 * align_up((size_t)m * 4, 256) + align_up((size_t)tiles * kRadix * 4, 256) + 256;
}

int ghscn_csr_build(const int64_t* key, const int64_t* other, int64_t num_edges, int64_t num_rows,
                    int32_t add_self_loops, int32_t* rowptr, int32_t* col, int32_t* perm, void* workspace,
                    size_t workspace_bytes, ghscn_stream_t stream_) {
  GHSCN_REQUIRE(num_edges >= 0 && num_rows >= 0 && rowptr != nullptr);
  GHSCN_REQUIRE(num_edges == 0 || (key != nullptr && other != nullptr));
  const int64_t m = num_edges + (add_self_loops ? num_rows : 0);
  GHSCN_REQUIRE(m < (int64_t)1 << 31 && num_rows < ((int64_t)1 << 31) - 1);
  GHSCN_REQUIRE(m == 0 || (col != nullptr && perm != nullptr));
  if (workspace_bytes < ghscn_csr_workspace_bytes(num_edges, num_rows, add_self_loops)) return GHSCN_E_WORKSPACE;
  cudaStream_t stream = as_stream(stream_);

  if (m == 0) {
    csr_finalize_kernel<<<ceil_div<int64_t>(num_rows + 1, 256), 256, 0, stream>>>(nullptr, nullptr, other, 0,
                                                                                   (int)num_rows, 0, rowptr, col);
    GHSCN_LAUNCH_CHECK();
    return GHSCN_OK;
  }
  GHSCN_REQUIRE(workspace != nullptr);
  const int tiles = (int)ceil_div<int64_t>(m, kTile);
  char* ws = static_cast<char*>(workspace);
  const size_t seg = align_up((size_t)m * 4, 256);
  int* keys_a = reinterpret_cast<int*>(ws);
  int* keys_b = reinterpret_cast<int*>(ws + seg);
  int* vals_b = reinterpret_cast<int*>(ws + 2 * seg);
  int* hist = reinterpret_cast<int*>(ws + 3 * seg);
  // vals ping-pong between `perm` (caller's output) and vals_b so that the last pass lands in perm.
  int bits = 0;
  while (((int64_t)1 << bits) <= num_rows) ++bits;  // keys span [0, num_rows]
  const int passes = bits == 0 ? 1 : ceil_div(bits, kRadixBits);

  SortSource src{key, other, num_edges, (int32_t)num_rows, add_self_loops};
  int* k_in = nullptr;
  int* v_in = nullptr;
  for (int p = 0; p < passes; ++p) {
    const bool to_perm = ((passes - 1 - p) % 2) == 0;  // last pass writes perm
    int* k_out = to_perm ? keys_a : keys_b;
    int* v_out = to_perm ? perm : vals_b;
    const int shift = p * kRadixBits;
    if (p == 0) {
      radix_hist_kernel<true><<<tiles, kSortThreads, 0, stream>>>(src, nullptr, m, shift, hist, tiles);
      exclusive_scan_kernel<<<1, 1024, 0, stream>>>(hist, tiles * kRadix);
      radix_scatter_kernel<true><<<tiles, kSortThreads, 0, stream>>>(src, nullptr, nullptr, m, shift, hist, tiles,
                                                                      k_out, v_out);
    } else {
      radix_hist_kernel<false><<<tiles, kSortThreads, 0, stream>>>(src, k_in, m, shift, hist, tiles);
      exclusive_scan_kernel<<<1, 1024, 0, stream>>>(hist, tiles * kRadix);
      radix_scatter_kernel<false><<<tiles, kSortThreads, 0, stream>>>(src, k_in, v_in, m, shift, hist, tiles,
                                                                       k_out, v_out);
    }
    k_in = k_out;
    v_in = v_out;
  }
  csr_finalize_kernel<<<ceil_div<int64_t>(m, 256), 256, 0, stream>>>(k_in, perm, other, num_edges, (int)num_rows,
                                                                     m, rowptr, col);
  GHSCN_LAUNCH_CHECK_N(3 * passes + 1);
  return GHSCN_OK;
}

int ghscn_batch_to_ptr(const int64_t* batch, int64_t num_nodes, int64_t num_graphs, int32_t* ptr,
                       ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_nodes >= 0 && num_graphs >= 0 && ptr != nullptr && (num_nodes == 0 || batch != nullptr));
  GHSCN_REQUIRE(num_nodes < (int64_t)1 << 31 && num_graphs < (int64_t)1 << 31);
  const int64_t work = num_nodes > 0 ? num_nodes : num_graphs + 1;
  batch_to_ptr_kernel<<<ceil_div<int64_t>(work, 256), 256, 0, as_stream(stream)>>>(batch, num_nodes,
                                                                                   (int)num_graphs, ptr);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

int ghscn_gcn_deg_inv_sqrt(const int32_t* rowptr, const int32_t* perm, const float* edge_weight,
                           const float* loop_weight, int64_t num_edges, int64_t num_rows, float* dis,
                           ghscn_stream_t stream) {
  GHSCN_REQUIRE(rowptr && dis && num_rows >= 0 && num_edges >= 0);
  GHSCN_REQUIRE(perm != nullptr || (edge_weight == nullptr && loop_weight == nullptr));
  if (num_rows == 0) return GHSCN_OK;
  gcn_dis_kernel<<<ceil_div<int64_t>(num_rows, 256), 256, 0, as_stream(stream)>>>(
      rowptr, perm, edge_weight, loop_weight, num_edges, (int)num_rows, dis);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

int ghscn_edge_weights(const int32_t* rowptr, const int32_t* col, const int32_t* perm, const float* edge_weight,
                       const float* loop_weight, const float* dis, int64_t num_edges, int64_t num_rows,
                       int32_t normalize, int32_t rows_are_dst, float* w, ghscn_stream_t stream) {
  GHSCN_REQUIRE(rowptr && num_rows >= 0 && num_edges >= 0);
  if (col == nullptr || perm == nullptr || w == nullptr) {
    GHSCN_REQUIRE(num_edges == 0);  // an empty relation has zero-sized slot arrays (null device pointers)
    return GHSCN_OK;
  }
  GHSCN_REQUIRE(!normalize || dis != nullptr);
  if (num_rows == 0) return GHSCN_OK;
  edge_weights_kernel<<<ceil_div<int64_t>(num_rows, 256), 256, 0, as_stream(stream)>>>(
      rowptr, col, perm, edge_weight, loop_weight, dis, num_edges, (int)num_rows, normalize, rows_are_dst, w);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

int ghscn_loop_weights(const int64_t* row, const int64_t* colidx, const float* edge_weight, int64_t num_edges,
                       int64_t num_rows, float fill, int32_t* scratch_last, float* loop_weight,
                       ghscn_stream_t stream_) {
  GHSCN_REQUIRE(num_edges >= 0 && num_rows >= 0 && scratch_last && loop_weight);
  GHSCN_REQUIRE(num_edges == 0 || (row && colidx));
  if (num_rows == 0) return GHSCN_OK;
  cudaStream_t stream = as_stream(stream_);
  fill_i32_kernel<<<ceil_div<int64_t>(num_rows, 256), 256, 0, stream>>>(scratch_last, num_rows, -1);
  if (num_edges > 0)
    loop_last_kernel<<<ceil_div<int64_t>(num_edges, 256), 256, 0, stream>>>(row, colidx, num_edges, (int)num_rows,
                                                                            scratch_last);
  loop_weight_kernel<<<ceil_div<int64_t>(num_rows, 256), 256, 0, stream>>>(scratch_last, edge_weight,
                                                                           (int)num_rows, fill, loop_weight);
  GHSCN_LAUNCH_CHECK_N(num_edges > 0 ? 3 : 2);
  return GHSCN_OK;
}

int ghscn_csr_add_loops(const int32_t* rowptr, const int32_t* col, const int32_t* perm, int64_t num_rows,
                        int64_t num_edges, int32_t* rowptr2, int32_t* col2, int32_t* perm2, ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_rows >= 0 && num_edges >= 0 && num_rows + num_edges < ((int64_t)1 << 31));
  GHSCN_REQUIRE(rowptr && rowptr2 && (num_rows + num_edges == 0 || (col2 && perm2)) && (num_edges == 0 || (col && perm)));
  csr_add_loops_kernel<<<(unsigned)ceil_div<int64_t>(num_rows + 1, 256), 256, 0, as_stream(stream)>>>(
      rowptr, col, perm, (int)num_rows, (int)num_edges, rowptr2, col2, perm2);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

size_t ghscn_csr_blocked_smem_bytes(int64_t max_nodes_per_graph, int64_t max_edges_per_graph) {
  if (max_nodes_per_graph < 0 || max_edges_per_graph < 0) return 0;
  return (size_t)(4 * max_nodes_per_graph + 2 + 4 * max_edges_per_graph) * sizeof(int);
}

int ghscn_csr_build_blocked(const int64_t* src, const int64_t* dst, int64_t num_edges, const int32_t* ptr,
                            int64_t num_graphs, int64_t num_nodes, int64_t max_nodes_per_graph,
                            int64_t max_edges_per_graph, int32_t* rowptr_dst, int32_t* col_dst, int32_t* perm_dst,
                            int32_t* rowptr_src, int32_t* col_src, int32_t* perm_src, int32_t* status,
                            ghscn_stream_t stream) {
  GHSCN_REQUIRE(src && dst && ptr && rowptr_dst && col_dst && perm_dst && rowptr_src && col_src && perm_src && status);
  GHSCN_REQUIRE(num_edges > 0 && num_graphs > 0 && num_nodes > 0 && max_nodes_per_graph > 0 && max_edges_per_graph > 0);
  GHSCN_REQUIRE(num_edges < ((int64_t)1 << 31) && num_nodes < ((int64_t)1 << 31) - 1 && num_graphs < ((int64_t)1 << 31));
  const size_t smem = ghscn_csr_blocked_smem_bytes(max_nodes_per_graph, max_edges_per_graph);
  if (smem > 200 * 1024) return GHSCN_E_UNSUPPORTED;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(csr_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  csr_blocked_kernel<<<(unsigned)num_graphs, kBlkThreads, smem, as_stream(stream)>>>(
      src, dst, num_edges, ptr, (int)num_graphs, (int)num_nodes, (int)max_nodes_per_graph, (int)max_edges_per_graph,
      rowptr_dst, col_dst, perm_dst, rowptr_src, col_src, perm_src, status);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

int ghscn_cast_i64_f32(const int64_t* in, int64_t n, float* out, ghscn_stream_t stream) {
  GHSCN_REQUIRE(n >= 0 && (n == 0 || (in && out)));
  if (n == 0) return GHSCN_OK;
  cast_i64_f32_kernel<<<ceil_div<int64_t>(n, 256), 256, 0, as_stream(stream)>>>(in, n, out);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

}  // extern "C"
