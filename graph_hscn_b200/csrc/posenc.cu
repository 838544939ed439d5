// Laplacian positional encoding, batched on the device.
// Replaces the per-graph host loop of graph_hscn/transform/posenc.py:14-82:
//     L = to_scipy_sparse_matrix(*get_laplacian(edge_index, normalization)).toarray()
//     evals, evects = np.linalg.eigh(L);  k smallest -> normalise -> NaN padding
// One CTA per graph.  Kernel 1 builds the dense Laplacian from the graph's CSR slice with the reference's float32
// entries (PyG forms the weights in float32 and scipy keeps the dtype), mirrored from the lower triangle like LAPACK's
// UPLO = 'L'.  Kernel 2 diagonalises it with a one-sided (Hestenes) Jacobi iteration in float64: G = (L + sigma I) V
// is rotated column pair by column pair until its columns are orthogonal -- then G_j = (lambda_j + sigma) v_j (V itself
// is never formed).  The
// n/2 disjoint pairs of a round-robin step run on the CTA's warps in parallel (a pair touches only its own two
// columns, so one barrier per step is all the synchronisation there is); the shift sigma >= 1 keeps the
// matrix positive definite and well conditioned (condition number <= 3 for the default normalisation), so the zero modes of the Laplacian are ordinary columns and
// every eigenpair comes out with absolute error ~1e-13 (the reference's LAPACK ssyevd: ~1e-6).  Kernel 3 ranks the
// eigenvalues, keeps the `max_freqs` smallest, normalises the float32 eigenvectors and writes the per-node rows.
// Eigenvector signs / the basis inside a repeated eigenvalue are not defined by the reference either (LAPACK's
// choice); the encoder that consumes them (encoder/signnet.py) is sign invariant by construction.
#include <cooperative_groups.h>
#include <math.h>

#include "common.cuh"

namespace ghscn {

constexpr int kEigThreads = 1024;
constexpr int kEigMaxSweeps = 40;
constexpr double kEigTol = 1e-14;

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}

// ---- kernel 1: dense shifted Laplacian G (column-major n x n, float64) -------------------------------------------------
// norm: 0 = D - A, 1 = I - D^-1/2 A D^-1/2, 2 = I - D^-1 A.  symmetrize != 0: to_undirected + coalesce (an edge in
// either direction counts once); else duplicate slots add up as scipy's toarray() does.  Self loops are dropped.
__global__ void __launch_bounds__(kEigThreads) laplacian_build_kernel(
    const int* __restrict__ ptr, const int* __restrict__ rowptr, const int* __restrict__ col, int n_cap, int norm,
    int symmetrize, double* __restrict__ gmat, double* __restrict__ vmat, double* __restrict__ shift) {
  extern __shared__ float deg[];                   // [n_cap]
  __shared__ float red[32];
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = blockDim.x >> 5;
  const int base = ptr[g], n = ptr[g + 1] - base;
  if (n <= 0 || n > n_cap) {
    if (tid == 0) shift[g] = 0.0;
    return;
  }
  double* G = gmat + (size_t)base * n_cap;
  double* V = vmat + (size_t)base * n_cap;         // scratch: the edge counts (row-major)
  const int nn = n * n;
  for (int e = tid; e < nn; e += blockDim.x) { G[e] = 0.0; V[e] = 0.0; }
  __syncthreads();
  if (symmetrize) {
    for (int i = tid; i < n; i += blockDim.x)
      for (int s = rowptr[base + i]; s < rowptr[base + i + 1]; ++s) {
        const int j = col[s] - base;
        if (j >= 0 && j < n && j != i) { V[i * n + j] = 1.0; V[j * n + i] = 1.0; }   // idempotent stores
      }
  } else {
    for (int i = tid; i < n; i += blockDim.x)      // a row belongs to one thread: no atomics
      for (int s = rowptr[base + i]; s < rowptr[base + i + 1]; ++s) {
        const int j = col[s] - base;
        if (j >= 0 && j < n && j != i) V[i * n + j] += 1.0;
      }
  }
  __syncthreads();
  for (int i = wid; i < n; i += nwarps) {          // deg = scatter_add(edge_weight, row): exact small integers
    float d = 0.f;
    for (int j = lane; j < n; j += 32) d += (float)V[i * n + j];
    d = warp_sum(d);
    if (lane == 0) deg[i] = d;
  }
  __syncthreads();
  // shift: Gershgorin keeps the spectrum of D - A inside [0, 2 dmax] and that of the mirrored lower triangle of
  // I - D^-1 A inside [1 - dmax, 1 + dmax]; the symmetric normalisation lives in [0, 2]
  float dmax = 1.0f;
  if (norm != 1)
    for (int i = tid; i < n; i += blockDim.x) dmax = fmaxf(dmax, deg[i]);
  // block max
  for (int o = 16; o > 0; o >>= 1) dmax = fmaxf(dmax, __shfl_xor_sync(kFullMask, dmax, o));
  if (lane == 0) red[wid] = dmax;
  __syncthreads();
  dmax = red[0];
  for (int w = 1; w < nwarps; ++w) dmax = fmaxf(dmax, red[w]);
  const double sigma = (double)dmax;
  if (tid == 0) shift[g] = sigma;
  // lower triangle i > j from row i of the counts, mirrored (np.linalg.eigh reads UPLO = 'L')
  for (int e = tid; e < nn; e += blockDim.x) {
    const int i = e / n, j = e - i * n;
    if (j > i) continue;
    float v;
    if (i == j) {
      v = norm == 0 ? deg[i] : 1.0f;
      G[(size_t)i * n + i] = (double)v + sigma;
      continue;
    }
    const float a = (float)V[i * n + j];
    if (a == 0.f) continue;
    if (norm == 0) {
      v = -a;
    } else if (norm == 1) {
      const float di = deg[i] > 0.f ? __fdiv_rn(1.0f, __fsqrt_rn(deg[i])) : 0.f;
      const float dj = deg[j] > 0.f ? __fdiv_rn(1.0f, __fsqrt_rn(deg[j])) : 0.f;
      v = -(__fmul_rn(__fmul_rn(di, 1.0f), dj)) * a;
    } else {
      const float di = deg[i] > 0.f ? __fdiv_rn(1.0f, deg[i]) : 0.f;
      v = -di * a;
    }
    G[(size_t)j * n + i] = (double)v;              // column-major (i, j)
    G[(size_t)i * n + j] = (double)v;              // and its mirror (j, i)
  }
}

// ---- kernel 2: one-sided Jacobi: rotate column pairs of G until all columns are orthogonal ----------------------------
// G starts as the (shifted, positive definite) matrix itself, i.e. G = A' V with V = I; every rotation is applied to
// the columns of G only.  At convergence G = A' V has orthogonal columns, so G_j = lambda'_j v_j: the eigenvalue is
// the column norm and the eigenvector the normalised column -- V is never stored (half the memory and 4 instead of
// 10 column transfers per pair).  The matrix lives in shared memory when n^2 doubles fit (n <= 168), else in the
// L2-resident workspace; a pair's two columns stay in registers between the dot products and the rotation (n <= 256).
// One column pair on one warp.  R > 0: the pair's 2 * 32 * R elements stay in registers between the dot products and
// the rotation (n <= 32 R); R = 0: columns of any length are read twice.  Returns whether the pair was rotated.
// CG: the columns are shared with the other CTAs of a cluster through L2, so loads bypass L1 (ld.global.cg).
template <int R, bool CG>
__device__ __forceinline__ bool jacobi_pair(double* gp, double* gq, int n, int lane) {
  constexpr int RR = R > 0 ? R : 1;
  double x[RR], y[RR];
  double a = 0.0, b = 0.0, c = 0.0;
  if (R > 0) {
#pragma unroll
    for (int u = 0; u < RR; ++u) {
      const int i = lane + 32 * u;
      x[u] = i < n ? (CG ? __ldcg(gp + i) : gp[i]) : 0.0;
      y[u] = i < n ? (CG ? __ldcg(gq + i) : gq[i]) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < RR; ++u) { a = fma(x[u], x[u], a); b = fma(y[u], y[u], b); c = fma(x[u], y[u], c); }
  } else {
    for (int i = lane; i < n; i += 32) {
      const double xv = CG ? __ldcg(gp + i) : gp[i], yv = CG ? __ldcg(gq + i) : gq[i];
      a = fma(xv, xv, a); b = fma(yv, yv, b); c = fma(xv, yv, c);
    }
  }
  a = warp_sum_f64(a); b = warp_sum_f64(b); c = warp_sum_f64(c);
  if (fabs(c) <= kEigTol * sqrt(a * b)) return false;
  const double z = (b - a) / (2.0 * c);
  const double t = (z >= 0.0 ? 1.0 : -1.0) / (fabs(z) + sqrt(1.0 + z * z));
  const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
  if (R > 0) {
#pragma unroll
    for (int u = 0; u < RR; ++u) {
      const int i = lane + 32 * u;
      if (i < n) {
        gp[i] = cs * x[u] - sn * y[u];
        gq[i] = sn * x[u] + cs * y[u];
      }
    }
  } else {
    for (int i = lane; i < n; i += 32) {
      const double xv = CG ? __ldcg(gp + i) : gp[i], yv = CG ? __ldcg(gq + i) : gq[i];
      gp[i] = cs * xv - sn * yv;
      gq[i] = sn * xv + cs * yv;
    }
  }
  return true;
}

template <bool CG>
__device__ __forceinline__ bool jacobi_step_pairs(double* G, int n, int m, int half, int s, int first, int stride,
                                                  int lane) {
  const int mode = n <= 128 ? 4 : (n <= 256 ? 8 : 0);
  bool mine = false;
  for (int k = first; k < half; k += stride) {     // a warp per pair; warp-uniform control flow
    int p, q;
    if (k == 0) { p = m - 1; q = s; }
    else { p = (s + k) % (m - 1); q = (s - k + (m - 1)) % (m - 1); }
    if (p >= n || q >= n) continue;
    double* gp = G + (size_t)p * n;
    double* gq = G + (size_t)q * n;
    if (mode == 4) mine |= jacobi_pair<4, CG>(gp, gq, n, lane);
    else if (mode == 8) mine |= jacobi_pair<8, CG>(gp, gq, n, lane);
    else mine |= jacobi_pair<0, CG>(gp, gq, n, lane);
  }
  return mine;
}

// grid = (cluster size C, graphs), cluster dims (C, 1, 1).  A graph whose matrix fits in shared memory is handled by
// rank 0 alone (the other ranks leave at once).  A larger graph -- the tail of a small batch: its matrix streams from
// L2 at ONE SM's bandwidth -- is shared by the C CTAs of the cluster: each takes every C-th warp's worth of a step's
// pairs on the L2-resident matrix, with a cluster barrier (release / acquire) between steps.
template <bool CLUSTER>
__global__ void __launch_bounds__(kEigThreads) jacobi_eig_kernel(const int* __restrict__ ptr, int n_cap, int smem_n,
                                                                 double* gmat, int* flags, int* __restrict__ sweeps_out) {
  namespace cgr = cooperative_groups;
  extern __shared__ double gs[];                   // [smem_n * smem_n] when the graph fits
  __shared__ int rotated;
  const int g = blockIdx.y, rank = blockIdx.x, csize = gridDim.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = blockDim.x >> 5;
  const int base = ptr[g], n = ptr[g + 1] - base;
  if (n <= 1 || n > n_cap) {
    if (rank == 0 && tid == 0 && sweeps_out) sweeps_out[g] = 0;
    return;
  }
  double* Gg = gmat + (size_t)base * n_cap;
  const bool in_smem = n <= smem_n;
  const int m = n + (n & 1), half = m >> 1;        // round-robin over m players; player n (odd n) sits out
  int sweep = 0;
  if (!CLUSTER || in_smem || csize == 1) {
    if (rank != 0) return;
    double* G = in_smem ? gs : Gg;
    const int nn = n * n;
    if (in_smem) {
      for (int e = tid; e < nn; e += blockDim.x) gs[e] = Gg[e];
      __syncthreads();
    }
    for (; sweep < kEigMaxSweeps; ++sweep) {
      if (tid == 0) rotated = 0;
      __syncthreads();
      bool mine = false;
      for (int s = 0; s < m - 1; ++s) {
        mine |= jacobi_step_pairs<false>(G, n, m, half, s, wid, nwarps, lane);
        __syncthreads();                           // the next step pairs the columns differently
      }
      if (mine) rotated = 1;
      __syncthreads();
      const int any = rotated;
      __syncthreads();
      if (!any) { ++sweep; break; }
    }
    if (in_smem)
      for (int e = tid; e < nn; e += blockDim.x) Gg[e] = gs[e];
  } else if (CLUSTER) {
    cgr::cluster_group cluster = cgr::this_cluster();
    volatile int* flag = flags + g;
    for (; sweep < kEigMaxSweeps; ++sweep) {
      if (rank == 0 && tid == 0) *flag = 0;
      __threadfence();
      cluster.sync();
      bool mine = false;
      for (int s = 0; s < m - 1; ++s) {
        mine |= jacobi_step_pairs<true>(Gg, n, m, half, s, rank * nwarps + wid, csize * nwarps, lane);
        __threadfence();
        cluster.sync();
      }
      if (mine && lane == 0) *flag = 1;
      __threadfence();
      cluster.sync();
      const int any = *flag;
      cluster.sync();                              // everyone has read the flag before rank 0 clears it
      if (!any) { ++sweep; break; }
    }
  }
  if (rank == 0 && tid == 0 && sweeps_out) sweeps_out[g] = sweep;
}

// ---- kernel 3: eigenvalues = column norms - shift, the max_freqs smallest, normalised float32 rows ---------------------
// eigvec_norm: 0 = L1, 1 = L2, 2 = abs-max (posenc.py:85-108, eps = 1e-12).  Outputs [N, max_freqs] each; graphs with
// fewer nodes than max_freqs get NaN in the missing columns (posenc.py:66-76).
__global__ void __launch_bounds__(kEigThreads) eig_select_kernel(const int* __restrict__ ptr, int n_cap, int max_freqs,
                                                                 int eigvec_norm, const double* __restrict__ gmat,
                                                                 const double* __restrict__ shift,
                                                                 float* __restrict__ eigvals,
                                                                 float* __restrict__ eigvecs) {
  extern __shared__ double lam[];                  // [n_cap] column norms, then int rank[n_cap]
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = blockDim.x >> 5;
  const int base = ptr[g], n = ptr[g + 1] - base;
  if (n <= 0) return;
  float* ev = eigvals + (size_t)base * max_freqs;
  float* ex = eigvecs + (size_t)base * max_freqs;
  if (n > n_cap) {                                 // stale hint: poison, like the other per-graph kernels
    for (int e = tid; e < n * max_freqs; e += blockDim.x) { ev[e] = NAN; ex[e] = NAN; }
    return;
  }
  int* rank = reinterpret_cast<int*>(lam + n_cap);
  const double* G = gmat + (size_t)base * n_cap;
  const double sigma = shift[g];
  for (int j = wid; j < n; j += nwarps) {
    double r = 0.0;
    for (int i = lane; i < n; i += 32) r = fma(G[(size_t)j * n + i], G[(size_t)j * n + i], r);
    r = warp_sum_f64(r);
    if (lane == 0) lam[j] = sqrt(r);
  }
  __syncthreads();
  for (int j = tid; j < n; j += blockDim.x) {      // ascending rank, ties by column index
    const double lj = lam[j];
    int r = 0;
    for (int i = 0; i < n; ++i) r += (lam[i] < lj) || (lam[i] == lj && i < j);
    rank[j] = r;
  }
  __syncthreads();
  for (int j = wid; j < n; j += nwarps) {
    const int r = rank[j];
    if (r >= max_freqs) continue;
    const double* v = G + (size_t)j * n;
    const double inv = 1.0 / lam[j];               // > 0: the shifted matrix is positive definite
    float denom = 0.f;
    for (int i = lane; i < n; i += 32) {
      const float x = fabsf((float)(v[i] * inv));
      denom = eigvec_norm == 0 ? denom + x : (eigvec_norm == 1 ? fmaf(x, x, denom) : fmaxf(denom, x));
    }
    if (eigvec_norm == 2) {
      for (int o = 16; o > 0; o >>= 1) denom = fmaxf(denom, __shfl_xor_sync(kFullMask, denom, o));
    } else {
      denom = warp_sum(denom);
      if (eigvec_norm == 1) denom = sqrtf(denom);
    }
    denom = fmaxf(denom, 1e-12f);
    const float val = fmaxf((float)(lam[j] - sigma), 0.f);   // clamp_min(0) (posenc.py:60)
    for (int i = lane; i < n; i += 32) {
      ex[(size_t)i * max_freqs + r] = __fdiv_rn((float)(v[i] * inv), denom);
      ev[(size_t)i * max_freqs + r] = val;
    }
  }
  for (int e = tid; e < n * max_freqs; e += blockDim.x) {
    const int r = e % max_freqs;
    if (r >= n) { ev[e] = NAN; ex[e] = NAN; }
  }
}

}  // namespace ghscn

using namespace ghscn;

extern "C" {

size_t ghscn_laplacian_eig_workspace_bytes(int64_t num_nodes, int64_t num_graphs, int32_t max_nodes_per_graph) {
  if (num_nodes < 0 || num_graphs < 0 || max_nodes_per_graph < 0) return 0;
  // G and V: a graph's n x n block starts at ptr[g] * n_cap (n^2 <= n * n_cap); + one shift per graph
  // ... + per graph: one shift (double) and one convergence flag (int, 8-byte slot)
  return ((size_t)2 * num_nodes * max_nodes_per_graph + (size_t)2 * num_graphs) * sizeof(double) + 256;
}

int ghscn_laplacian_eig(const int32_t* ptr, const int32_t* rowptr, const int32_t* col, int64_t num_graphs,
                        int64_t num_nodes, int32_t max_nodes_per_graph, int32_t laplacian_norm, int32_t symmetrize,
                        int32_t max_freqs, int32_t eigvec_norm, float* eigvals, float* eigvecs, int32_t* sweeps,
                        void* workspace, size_t workspace_bytes, ghscn_stream_t stream_) {
  GHSCN_REQUIRE(num_graphs >= 0 && num_nodes >= 0 && max_freqs > 0);
  GHSCN_REQUIRE(laplacian_norm >= 0 && laplacian_norm <= 2 && eigvec_norm >= 0 && eigvec_norm <= 2);
  GHSCN_REQUIRE(num_graphs < 65536 && num_nodes < ((int64_t)1 << 31));   // grid.y
  if (num_graphs == 0 || num_nodes == 0) return GHSCN_OK;
  GHSCN_REQUIRE(ptr && rowptr && col && eigvals && eigvecs && max_nodes_per_graph > 0);
  const int n_cap = max_nodes_per_graph;
  if (n_cap > 4096) return GHSCN_E_UNSUPPORTED;    // shared arrays of the select pass: 12 bytes per node
  if (workspace == nullptr ||
      workspace_bytes < ghscn_laplacian_eig_workspace_bytes(num_nodes, num_graphs, max_nodes_per_graph))
    return GHSCN_E_WORKSPACE;
  GHSCN_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7) == 0);
  cudaStream_t stream = as_stream(stream_);
  double* gmat = static_cast<double*>(workspace);
  double* vmat = gmat + (size_t)num_nodes * n_cap;
  double* shift = vmat + (size_t)num_nodes * n_cap;
  laplacian_build_kernel<<<(unsigned)num_graphs, kEigThreads, (size_t)n_cap * 4, stream>>>(
      ptr, rowptr, col, n_cap, laplacian_norm, symmetrize, gmat, vmat, shift);
  // the Jacobi kernel keeps a graph's matrix in shared memory when it fits: room for min(n_cap, 168)^2 doubles
  const int smem_n = n_cap < 168 ? n_cap : 168;
  const size_t jac_shm = (size_t)smem_n * smem_n * sizeof(double);
  cudaFuncSetAttribute(jacobi_eig_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(168 * 168 * sizeof(double)));
  cudaFuncSetAttribute(jacobi_eig_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(168 * 168 * sizeof(double)));
  // small batches with graphs beyond the shared-memory size: a cluster of 4 CTAs shares each of those (the tail)
  const int csize = (n_cap > smem_n && num_graphs <= 2 * kNumSMs) ? 4 : 1;
  int* flags = reinterpret_cast<int*>(shift + num_graphs);
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)csize, (unsigned)num_graphs);
    cfg.blockDim = dim3(kEigThreads);
    cfg.dynamicSmemBytes = jac_shm;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = csize > 1 ? cudaLaunchKernelEx(&cfg, jacobi_eig_kernel<true>, ptr, n_cap, smem_n, gmat, flags, sweeps)
                              : cudaLaunchKernelEx(&cfg, jacobi_eig_kernel<false>, ptr, n_cap, smem_n, gmat, flags, sweeps);
    if (e != cudaSuccess) return (int)e;
  }
  eig_select_kernel<<<(unsigned)num_graphs, kEigThreads, (size_t)n_cap * 12, stream>>>(
      ptr, n_cap, max_freqs, eigvec_norm, gmat, shift, eigvals, eigvecs);
  GHSCN_LAUNCH_CHECK_N(3);
  return GHSCN_OK;
}

}  // extern "C"
