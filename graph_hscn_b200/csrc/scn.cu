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

}  // extern "C"
