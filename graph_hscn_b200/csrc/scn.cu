// Fused node pipeline of the spectral-clustering net (model/hscn.py:30-45,57-60 with mp_units = [U]):
//   agg[i,:]    = sum_{s in row i} w[s] * x[col[s],:]            GraphConv aggregation at the INPUT width (F <= 16)
//   pre[i,:]    = W_rel agg[i,:] + b_rel + W_root x[i,:]          GraphConv.lin_rel / lin_root (U <= 32)
//   h[i,:]      = act(pre[i,:])                                   ELU / ReLU / tanh / identity (config/config.py:13-18)
//   logits[i,:] = W_out h[i,:] + b_out                            the cluster MLP's Linear (K <= 32)
// In the reference these are a gather/mul/scatter, three Linear layers, an add and an activation -- six launches of
// work on 9..16 floats per node, twice per step (training forward and cluster assignment).  Here one thread walks one
// node's CSR row and keeps the 9 + 16 + 10 intermediate values in registers; the weights (< 4 KB) sit in shared memory.
// HBM traffic: x rows of the neighbours (L1/L2 hits: collated molecules are index-local) + one write of each output.
#include "common.cuh"

namespace ghscn {

constexpr int kScnMaxF = 16, kScnMaxU = 32, kScnMaxK = 32;

__device__ __forceinline__ float scn_act(float v, int act) {
  switch (act) {
    case 1: return v > 0.f ? v : expm1f(v);       // ELU, alpha = 1
    case 2: return fmaxf(v, 0.f);
    case 3: return tanhf(v);
    default: return v;
  }
}

template <int F, int U>
__global__ void __launch_bounds__(128)
scn_forward_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const float* __restrict__ w,
                   const float* __restrict__ x, int64_t ldx, int num_nodes, int f_in, int units, int clusters,
                   const float* __restrict__ w_rel, const float* __restrict__ b_rel, const float* __restrict__ w_root,
                   const float* __restrict__ w_out, const float* __restrict__ b_out, int act,
                   float* __restrict__ agg_out, float* __restrict__ pre_out, float* __restrict__ h_out,
                   float* __restrict__ logits) {
  __shared__ float s_rel[U * F], s_root[U * F], s_out[kScnMaxK * U], s_brel[U], s_bout[kScnMaxK];
  for (int i = threadIdx.x; i < U * F; i += blockDim.x) {
    const int u = i / F, k = i - u * F;
    const bool on = u < units && k < f_in;
    s_rel[i] = on ? w_rel[u * f_in + k] : 0.f;
    s_root[i] = on ? w_root[u * f_in + k] : 0.f;
  }
  for (int i = threadIdx.x; i < kScnMaxK * U; i += blockDim.x) {
    const int c = i / U, u = i - c * U;
    s_out[i] = (c < clusters && u < units) ? w_out[c * units + u] : 0.f;
  }
  for (int i = threadIdx.x; i < U; i += blockDim.x) s_brel[i] = (i < units && b_rel) ? b_rel[i] : 0.f;
  for (int i = threadIdx.x; i < kScnMaxK; i += blockDim.x) s_bout[i] = (i < clusters && b_out) ? b_out[i] : 0.f;
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= num_nodes) return;

  float agg[F], xi[F];
#pragma unroll
  for (int k = 0; k < F; ++k) { agg[k] = 0.f; xi[k] = k < f_in ? __ldg(x + (int64_t)n * ldx + k) : 0.f; }
  const int beg = rowptr[n], end = rowptr[n + 1];
  for (int s = beg; s < end; ++s) {                       // slot order, unfused mul + add (= CPU scatter_add_)
    const float ws = w ? w[s] : 1.f;
    const float* xr = x + (int64_t)col[s] * ldx;
#pragma unroll
    for (int k = 0; k < F; ++k)
      if (k < f_in) agg[k] = mul_then_add(agg[k], ws, __ldg(xr + k));
  }
  if (agg_out) {
#pragma unroll
    for (int k = 0; k < F; ++k)
      if (k < f_in) agg_out[(int64_t)n * f_in + k] = agg[k];
  }
  float h[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    float rel = 0.f, root = 0.f;
#pragma unroll
    for (int k = 0; k < F; ++k) {
      rel = fmaf(agg[k], s_rel[u * F + k], rel);
      root = fmaf(xi[k], s_root[u * F + k], root);
    }
    const float pre = (rel + s_brel[u]) + root;            // lin_rel(agg) (+ bias) first, then + lin_root(x)
    h[u] = scn_act(pre, act);
    if (u < units) {
      if (pre_out) pre_out[(int64_t)n * units + u] = pre;
      if (h_out) h_out[(int64_t)n * units + u] = h[u];
    }
  }
  for (int c = 0; c < clusters; ++c) {
    float acc = 0.f;
#pragma unroll
    for (int u = 0; u < U; ++u) acc = fmaf(h[u], s_out[c * U + u], acc);
    logits[(int64_t)n * clusters + c] = acc + s_bout[c];
  }
}


// ---- backward of the node pipeline w.r.t. its parameters ---------------------------------------------------------
//   dh = ds W_out,  dpre = dh * act'(pre),
//   dW_out = ds^T h, db_out = colsum(ds), dW_rel = dpre^T agg, db_rel = colsum(dpre), dW_root = dpre^T x.
// A CTA owns 128 consecutive nodes.  Phase 1: one thread per node computes dpre and parks the node's five small rows
// (ds, h, dpre, agg, x: <= 96 floats) in shared memory.  Phase 2: one thread per GRADIENT ELEMENT walks the CTA's
// nodes in order and accumulates its product from shared memory (two loads + one FMA per node; consecutive threads
// read consecutive columns of the right-hand row, the left-hand value is a broadcast).  CTAs are added in CTA order by
// the second kernel: a fixed order, hence deterministic.  (The first version formed the outer products per thread and
// reduced each of the 474 elements with warp shuffles: 2 370 dependent shuffles per warp, 78 CTAs, 59 us.)
// Replaces three two-stage dW kernels, two column sums, a dX kernel and the activation backward (12 launches).
constexpr int kScnBwdThreads = 256;
constexpr int kScnBwdNodes = 128;

__host__ __device__ inline int scn_grad_count(int f_in, int units, int clusters) {
  return clusters * units + clusters + 2 * units * f_in + units;
}

template <int F, int U, int KC>
__global__ void __launch_bounds__(kScnBwdThreads)
scn_backward_kernel(const float* __restrict__ ds, const float* __restrict__ h, const float* __restrict__ pre,
                    const float* __restrict__ agg, const float* __restrict__ x, int64_t ldx, int num_nodes, int f_in,
                    int units, int clusters, const float* __restrict__ w_out, int act, float* __restrict__ partial) {
  extern __shared__ float scn_sm[];
  float* s_ds = scn_sm;                                  // [nodes][KC]
  float* s_h = s_ds + kScnBwdNodes * KC;                 // [nodes][U]
  float* s_dp = s_h + kScnBwdNodes * U;                  // [nodes][U]
  float* s_ag = s_dp + kScnBwdNodes * U;                 // [nodes][F]
  float* s_x = s_ag + kScnBwdNodes * F;                  // [nodes][F]
  float* s_wout = s_x + kScnBwdNodes * F;                // [KC][U]
  const int count = scn_grad_count(f_in, units, clusters);
  for (int i = threadIdx.x; i < KC * U; i += blockDim.x) {
    const int c = i / U, u = i - c * U;
    s_wout[i] = (c < clusters && u < units) ? w_out[c * units + u] : 0.f;
  }
  __syncthreads();
  const int n0 = blockIdx.x * kScnBwdNodes;
  const int nodes = min(kScnBwdNodes, num_nodes - n0);
  if (threadIdx.x < kScnBwdNodes) {
    const int ln = threadIdx.x, n = n0 + ln;
    const bool on = ln < nodes;
    float dsv[KC];
#pragma unroll
    for (int c = 0; c < KC; ++c) {
      dsv[c] = (on && c < clusters) ? ds[(int64_t)n * clusters + c] : 0.f;
      s_ds[ln * KC + c] = dsv[c];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = on && u < units;
      const float hv = ok ? h[(int64_t)n * units + u] : 0.f;
      const float p = ok ? pre[(int64_t)n * units + u] : 0.f;
      float dh = 0.f;
#pragma unroll
      for (int c = 0; c < KC; ++c) dh = fmaf(dsv[c], s_wout[c * U + u], dh);
      float g;
      switch (act) {
        case 1: g = p > 0.f ? dh : dh * expf(p); break;      // ELU'(p) = exp(p) for p <= 0
        case 2: g = p > 0.f ? dh : 0.f; break;
        case 3: g = dh * (1.f - hv * hv); break;
        default: g = dh;
      }
      s_h[ln * U + u] = hv;
      s_dp[ln * U + u] = ok ? g : 0.f;
    }
#pragma unroll
    for (int k = 0; k < F; ++k) {
      const bool ok = on && k < f_in;
      s_ag[ln * F + k] = ok ? agg[(int64_t)n * f_in + k] : 0.f;
      s_x[ln * F + k] = ok ? __ldg(x + (int64_t)n * ldx + k) : 0.f;
    }
  }
  __syncthreads();
  // layout of the gradient vector: dW_out [K,U] | db_out [K] | dW_rel [U,F] | db_rel [U] | dW_root [U,F]
  const int o1 = clusters * units, o2 = o1 + clusters, o3 = o2 + units * f_in, o4 = o3 + units;
  for (int o = threadIdx.x; o < count; o += blockDim.x) {
    const float* a;                    // left factor: element i of a row of stride sa
    const float* b = nullptr;          // right factor: element j of a row of stride sb (nullptr: column sum of a)
    int sa, sb = 0;
    if (o < o1)      { a = s_ds + o / units;            sa = KC; b = s_h + o % units;            sb = U; }
    else if (o < o2) { a = s_ds + (o - o1);             sa = KC; }
    else if (o < o3) { a = s_dp + (o - o2) / f_in;      sa = U;  b = s_ag + (o - o2) % f_in;     sb = F; }
    else if (o < o4) { a = s_dp + (o - o3);             sa = U; }
    else             { a = s_dp + (o - o4) / f_in;      sa = U;  b = s_x + (o - o4) % f_in;      sb = F; }
    float acc = 0.f;
    if (b != nullptr) {
#pragma unroll 4
      for (int ln = 0; ln < nodes; ++ln) acc = fmaf(a[ln * sa], b[ln * sb], acc);
    } else {
#pragma unroll 4
      for (int ln = 0; ln < nodes; ++ln) acc += a[ln * sa];
    }
    partial[(int64_t)blockIdx.x * count + o] = acc;
  }
}

// grads[i] = sum over CTAs, in CTA order
__global__ void __launch_bounds__(256) scn_backward_reduce_kernel(const float* __restrict__ partial, int num_ctas,
                                                                  int count, float* __restrict__ grads) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float t = 0.f;
  for (int b = 0; b < num_ctas; ++b) t += partial[(int64_t)b * count + i];
  grads[i] = t;
}

}  // namespace ghscn

using namespace ghscn;

extern "C" {

int ghscn_scn_forward(const int32_t* rowptr, const int32_t* col, const float* w, const float* x, int64_t ldx,
                      int64_t num_nodes, int64_t f_in, int64_t units, int64_t clusters, const float* w_rel,
                      const float* b_rel, const float* w_root, const float* w_out, const float* b_out, int32_t act,
                      float* agg, float* pre, float* h, float* logits, ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_nodes >= 0 && num_nodes < ((int64_t)1 << 31));
  if (f_in < 1 || f_in > kScnMaxF || units < 1 || units > kScnMaxU || clusters < 1 || clusters > kScnMaxK ||
      act < 0 || act > 3)
    return GHSCN_E_UNSUPPORTED;
  if (num_nodes == 0) return GHSCN_OK;
  GHSCN_REQUIRE(rowptr && col && x && w_rel && w_root && w_out && logits && ldx >= f_in);
  const unsigned grid = (unsigned)ceil_div<int64_t>(num_nodes, 128);
  cudaStream_t st = as_stream(stream);
#define GHSCN_SCN_LAUNCH(F, U)                                                                                       \
  scn_forward_kernel<F, U><<<grid, 128, 0, st>>>(rowptr, col, w, x, ldx, (int)num_nodes, (int)f_in, (int)units,       \
                                                 (int)clusters, w_rel, b_rel, w_root, w_out, b_out, act, agg, pre, h, \
                                                 logits)
  if (f_in <= 9 && units <= 16) GHSCN_SCN_LAUNCH(9, 16);
  else if (units <= 16) GHSCN_SCN_LAUNCH(16, 16);
  else GHSCN_SCN_LAUNCH(16, 32);
#undef GHSCN_SCN_LAUNCH
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

size_t ghscn_scn_backward_workspace_bytes(int64_t num_nodes, int64_t f_in, int64_t units, int64_t clusters) {
  if (num_nodes < 0 || f_in < 1 || units < 1 || clusters < 1) return 0;
  const int64_t ctas = ceil_div<int64_t>(num_nodes > 0 ? num_nodes : 1, kScnBwdNodes);
  return (size_t)ctas * (size_t)scn_grad_count((int)f_in, (int)units, (int)clusters) * sizeof(float);
}

int ghscn_scn_backward(const float* ds, const float* h, const float* pre, const float* agg, const float* x, int64_t ldx,
                       int64_t num_nodes, int64_t f_in, int64_t units, int64_t clusters, const float* w_out,
                       int32_t act, float* grads, void* workspace, size_t workspace_bytes, ghscn_stream_t stream) {
  GHSCN_REQUIRE(num_nodes >= 0 && num_nodes < ((int64_t)1 << 31));
  if (f_in < 1 || f_in > kScnMaxF || units < 1 || units > kScnMaxU || clusters < 1 || clusters > kScnMaxK ||
      act < 0 || act > 3)
    return GHSCN_E_UNSUPPORTED;
  GHSCN_REQUIRE(grads != nullptr);
  const int count = scn_grad_count((int)f_in, (int)units, (int)clusters);
  cudaStream_t st = as_stream(stream);
  if (num_nodes == 0) {
    cudaError_t e = cudaMemsetAsync(grads, 0, (size_t)count * sizeof(float), st);
    return e == cudaSuccess ? GHSCN_OK : (int)e;
  }
  GHSCN_REQUIRE(ds && h && pre && agg && x && w_out && workspace && ldx >= f_in);
  if (workspace_bytes < ghscn_scn_backward_workspace_bytes(num_nodes, f_in, units, clusters)) return GHSCN_E_WORKSPACE;
  const unsigned ctas = (unsigned)ceil_div<int64_t>(num_nodes, kScnBwdNodes);
  float* partial = static_cast<float*>(workspace);
#define GHSCN_SCN_BWD(F, U, KC)                                                                                      \
  do {                                                                                                               \
    const size_t shm = ((size_t)kScnBwdNodes * ((KC) + 2 * (U) + 2 * (F)) + (size_t)(KC) * (U)) * sizeof(float);     \
    if (shm > 48 * 1024) {                                                                                           \
      cudaError_t e = cudaFuncSetAttribute(scn_backward_kernel<F, U, KC>,                                            \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);                   \
      if (e != cudaSuccess) return (int)e;                                                                           \
    }                                                                                                                \
    scn_backward_kernel<F, U, KC><<<ctas, kScnBwdThreads, shm, st>>>(ds, h, pre, agg, x, ldx, (int)num_nodes,         \
                                                                     (int)f_in, (int)units, (int)clusters, w_out,    \
                                                                     act, partial);                                  \
  } while (0)
  if (f_in <= 9 && units <= 16 && clusters <= 16) GHSCN_SCN_BWD(9, 16, 16);
  else if (units <= 16) GHSCN_SCN_BWD(16, 16, 32);
  else GHSCN_SCN_BWD(16, 32, 32);
#undef GHSCN_SCN_BWD
  GHSCN_LAUNCH_CHECK();
  scn_backward_reduce_kernel<<<(unsigned)ceil_div(count, 256), 256, 0, st>>>(partial, (int)ctas, count, grads);
  GHSCN_LAUNCH_CHECK();
  return GHSCN_OK;
}

}  // extern "C"
