/*
 * ghscn.h -- C ABI of libghscn.so: the B200 (sm_100a) kernels behind the
 * Graph-HSCN hot path.
 *
 * The reference (camille-004/Graph-HSCN) has no FFI layer of its own: its
 * hot path is reached by importing PyTorch Geometric / torch_scatter
 * operators.  Each entry point below therefore cites the *reference call
 * site* whose operator it replaces (paths relative to the reference root) and
 * the SURVEY.md section 8a row.  The Python host
 * (graph_hscn_b200/_lib.py -> ops.py -> pyg/*) binds these with ctypes and
 * wraps them as torch custom ops; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller (PyTorch) owns every buffer; the library never allocates,
 *     frees or synchronises, and keeps no state that a result depends on; all
 *     work is enqueued on `stream` and is CUDA-graph-capture safe.  The only
 *     process-wide state is instrumentation: the monotonic launch counter
 *     (ghscn_launch_count) and the optional pipeline-trace buffer of debug
 *     builds (ghscn_gemm3x_set_trace);
 *   - feature matrices are row-major fp32 with an explicit leading dimension
 *     (elements); index arrays produced by this library are int32, index
 *     arrays received from PyTorch (edge_index, batch) are int64;
 *   - return value: 0 = ok, <0 = GHSCN_E_* argument error (nothing was
 *     launched), >0 = cudaError_t reported by the launch.
 */
#ifndef GHSCN_H_
#define GHSCN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GHSCN_API __attribute__((visibility("default")))

typedef void* ghscn_stream_t; /* cudaStream_t */

enum {
  GHSCN_OK = 0,
  GHSCN_E_INVALID = -1,    /* null pointer / negative size / bad flag */
  GHSCN_E_WORKSPACE = -2,  /* workspace smaller than *_workspace_bytes() */
  GHSCN_E_UNSUPPORTED = -3 /* shape outside what the kernel supports */
};

GHSCN_API int ghscn_abi_version(void);
GHSCN_API const char* ghscn_error_string(int code);
/* number of CUDA kernels this library has launched in this process (monotonic; bench accounting) */
GHSCN_API unsigned long long ghscn_launch_count(void);

/* ---- K1: CSR construction (stable integer radix sort + boundary scan) ------------------------
 * Replaces the implicit COO handling of PyG MessagePassing and `add_remaining_self_loops`
 * inside gcn_norm (train/train_clustering.py:37-42; every GCNConv.forward at model/mpnn.py:52,59
 * and model/hscn.py:109).  SURVEY 8a rows a1, a2.
 * Sorts the E edges stably by key[] (destination for the forward structure, source for the
 * transposed one).  With add_self_loops != 0, edges with key == other are dropped and one loop
 * (i,i) per row is appended after all original edges, exactly PyG's edge order.  Negative keys are
 * padding and are dropped.
 *   rowptr [num_rows+1]  row r owns slots rowptr[r] .. rowptr[r+1]-1; rowptr[num_rows] = nnz
 *   col    [E (+num_rows)] other endpoint of the edge in each slot (-1 in unused tail slots)
 *   perm   [E (+num_rows)] original edge id of each slot; ids >= E are the appended loops
 */
GHSCN_API size_t ghscn_csr_workspace_bytes(int64_t num_edges, int64_t num_rows, int32_t add_self_loops);
GHSCN_API int ghscn_csr_build(const int64_t* key, const int64_t* other, int64_t num_edges, int64_t num_rows,
                              int32_t add_self_loops, int32_t* rowptr, int32_t* col, int32_t* perm,
                              void* workspace, size_t workspace_bytes, ghscn_stream_t stream);

/* CSR of the same edges plus one appended loop per row, derived from the plain CSR without sorting again.
 * Only valid when the edge list holds no self loop (nothing for add_remaining_self_loops to drop);
 * outputs are sized num_edges + num_rows and equal ghscn_csr_build(..., add_self_loops = 1). */
GHSCN_API int ghscn_csr_add_loops(const int32_t* rowptr, const int32_t* col, const int32_t* perm, int64_t num_rows,
                                  int64_t num_edges, int32_t* rowptr2, int32_t* col2, int32_t* perm2,
                                  ghscn_stream_t stream);

/* K1 fast path for collated mini-batches (edges are stored graph-major and never leave their graph).
 * One CTA per graph builds BOTH orientations of the plain (loop-free handling: none added, none dropped) CSR in
 * shared memory; results are bit-identical to ghscn_csr_build(dst, src, ...) and ghscn_csr_build(src, dst, ...).
 *   ptr [num_graphs+1]   node ranges of the graphs (ghscn_batch_to_ptr)
 *   max_*_per_graph      upper bounds the caller knows from collate time (they size the shared memory)
 *   status               device int32, caller-zeroed: bit 0 = a bound was exceeded, bit 1 = an edge leaves its
 *                        graph; outputs are unspecified when it is non-zero
 * GHSCN_E_UNSUPPORTED when the per-graph working set exceeds 200 KB of shared memory (use ghscn_csr_build). */
GHSCN_API size_t ghscn_csr_blocked_smem_bytes(int64_t max_nodes_per_graph, int64_t max_edges_per_graph);
GHSCN_API int ghscn_csr_build_blocked(const int64_t* src, const int64_t* dst, int64_t num_edges, const int32_t* ptr,
                                      int64_t num_graphs, int64_t num_nodes, int64_t max_nodes_per_graph,
                                      int64_t max_edges_per_graph, int32_t* rowptr_dst, int32_t* col_dst,
                                      int32_t* perm_dst, int32_t* rowptr_src, int32_t* col_src, int32_t* perm_src,
                                      int32_t* status, ghscn_stream_t stream);

/* Device-side collate (loader/loader.py:48-60, `Batch.from_data_list`): the host packs the graphs of a mini-batch back
 * to back with their LOCAL edge indices ([2, edge_capacity], row r at edge_index_local + r * edge_capacity) and one
 * node / edge count per graph; this call writes `batch` (graph id of every node) and the offset edge_index.  The counts
 * must sum to node_capacity / edge_capacity (padding graphs included). */
GHSCN_API int ghscn_collate_batch(const int32_t* node_counts, const int32_t* edge_counts, int64_t num_graphs,
                                  const int64_t* edge_index_local, int64_t edge_capacity, int64_t* edge_index,
                                  int64_t* batch, int64_t node_capacity, ghscn_stream_t stream);

/* sorted `batch` vector -> ptr[num_graphs+1] (PyG collate convention). */
GHSCN_API int ghscn_batch_to_ptr(const int64_t* batch, int64_t num_nodes, int64_t num_graphs, int32_t* ptr,
                                 ghscn_stream_t stream);

/* ---- K1b: gcn_norm folded onto the CSR ------------------------------------------------------
 * Replaces gcn_norm's scatter_add degree + deg^-1/2[row] * w * deg^-1/2[col]
 * (train/train_clustering.py:37-42; GCNConv.forward).  SURVEY 8a row a1, Appendix A.2.
 * `rowptr/perm` must be the BY-DESTINATION structure.  edge_weight may be NULL (all ones).
 * loop_weight[num_rows] (nullable) carries the weight of appended loops (fill value, or the
 * weight of a pre-existing loop); NULL means 1.0.  dis[r] = deg^-1/2 with inf -> 0.
 */
GHSCN_API int ghscn_gcn_deg_inv_sqrt(const int32_t* rowptr, const int32_t* perm, const float* edge_weight,
                                     const float* loop_weight, int64_t num_edges, int64_t num_rows, float* dis,
                                     ghscn_stream_t stream);
/* Per-slot weights for either orientation of the same edge set:
 *   normalize != 0: w[s] = (dis[src] * ew) * dis[dst]   (PyG multiplication order)
 *   normalize == 0: w[s] = ew
 * rows_are_dst tells which endpoint the structure's rows are. */
GHSCN_API int ghscn_edge_weights(const int32_t* rowptr, const int32_t* col, const int32_t* perm,
                                 const float* edge_weight, const float* loop_weight, const float* dis,
                                 int64_t num_edges, int64_t num_rows, int32_t normalize, int32_t rows_are_dst,
                                 float* w, ghscn_stream_t stream);
/* loop_weight[i] = weight of the LAST edge (i,i) in edge order, else fill (PyG add_remaining_self_loops). */
GHSCN_API int ghscn_loop_weights(const int64_t* row, const int64_t* colidx, const float* edge_weight,
                                 int64_t num_edges, int64_t num_rows, float fill, int32_t* scratch_last,
                                 float* loop_weight, ghscn_stream_t stream);

/* ---- K2/K3: SpMM  y[r,:] = sum_{s in row r} w[s] * x[col[s],:]  (+ bias) (+ relu) ----------------
 * Replaces MessagePassing.propagate = index_select -> mul -> scatter_add_ of GCNConv
 * (model/mpnn.py:52,59; model/hscn.py:88-93) and GraphConv (model/hscn.py:32-34,40-41).
 * SURVEY 8a rows a2, a3.  The sum runs sequentially in slot (= edge) order with separate
 * multiply and add roundings, i.e. the CPU scatter_add_ order.  The same entry computes the
 * backward w.r.t. x when given the transposed structure.  w == NULL means unit weights.
 */
GHSCN_API int ghscn_spmm(const int32_t* rowptr, const int32_t* col, const float* w, const float* x, int64_t ldx,
                         float* y, int64_t ldy, const float* bias, int64_t num_rows, int64_t num_feat,
                         int32_t relu, ghscn_stream_t stream);
/* Same contraction for pooling relations (few destination rows, many sources each, e.g. local -> virtual):
 * one CTA per row, fixed-order (deterministic) combination of 8 partial sums instead of the sequential order. */
/* y = A_w (x (.) [mask > 0]): ghscn_spmm over rows of x that are masked as they are gathered -- the input gradient of
 * `relu(GCNConv(...))` (model/mpnn.py:52, model/hscn.py:110) computed from dy and the layer's output in one pass
 * (x = dy, mask = y, (rowptr, col, w) = the transposed structure).  128 <= num_feat <= 512, multiples of 4. */
GHSCN_API int ghscn_spmm_masked_supported(int64_t num_feat, int64_t ldx, int64_t ldm, int64_t ldy);
GHSCN_API int ghscn_spmm_masked(const int32_t* rowptr, const int32_t* col, const float* w, const float* x, int64_t ldx,
                                const float* mask, int64_t ldm, float* y, int64_t ldy, int64_t num_rows,
                                int64_t num_feat, ghscn_stream_t stream);
GHSCN_API int ghscn_spmm_pool(const int32_t* rowptr, const int32_t* col, const float* w, const float* x, int64_t ldx,
                              float* y, int64_t ldy, const float* bias, int64_t num_rows, int64_t num_feat,
                              ghscn_stream_t stream);
/* d(edge weight)[s] = <dy[row(s),:], x[col[s],:]> written at the ORIGINAL edge position perm[s]
 * (entries for appended loops, perm >= num_edges, are skipped).  perm == NULL writes dw_edge[s]. */
GHSCN_API int ghscn_spmm_edge_grad(const int32_t* rowptr, const int32_t* col, const int32_t* perm, const float* x,
                                   int64_t ldx, const float* dy, int64_t lddy, int64_t num_rows, int64_t num_feat,
                                   int64_t num_edges, float* dw_edge, ghscn_stream_t stream);

/* ---- K4: segment mean / sum over contiguous row ranges ---------------------------------------
 * Replaces torch_scatter.scatter_mean(x, batch, dim=0) (model/mpnn.py:60) and
 * global_mean_pool (model/hscn.py:111).  SURVEY 8a row a10.  mean != 0 divides by max(count,1).
 * perm (nullable) gathers rows (x[perm[i]]) for unsorted indices. */
GHSCN_API int ghscn_segment_reduce(const float* x, int64_t ldx, const int32_t* ptr, const int32_t* perm,
                                   int64_t num_segments, int64_t num_feat, int32_t mean, float* y, int64_t ldy,
                                   ghscn_stream_t stream);
/* backward: dx[i,:] = dy[seg(i),:] * (mean ? 1/max(count,1) : 1) */
GHSCN_API int ghscn_segment_broadcast(const float* dy, int64_t lddy, const int32_t* ptr, const int32_t* perm,
                                      int64_t num_segments, int64_t num_feat, int32_t mean, float* dx, int64_t lddx,
                                      ghscn_stream_t stream);

/* y = dropout(relu(x), p) in one pass, training mode (model/mpnn.py:57-58, `F.dropout(self.activation(x), ...)` with
 * the ReLU activation of config/config.py:13-18).  `state` = two device uint64 {seed, call counter}: the mask is
 * Philox4x32-10(counter = element index, call counter; key = seed), and the call increments the counter on the device,
 * so replays of a captured CUDA graph draw fresh masks.  The backward needs only the output:
 * dx = y > 0 ? dy / (1 - p) : 0 (a dropped element and a negative input both give y = 0). */
GHSCN_API int ghscn_relu_dropout_fwd(const float* x, int64_t n, float p, uint64_t* state, float* y,
                                     ghscn_stream_t stream);
GHSCN_API int ghscn_relu_dropout_bwd(const float* dy, const float* y, int64_t n, float p, float* dx,
                                     ghscn_stream_t stream);

/* Column sum out[f] = sum_r x[r,f] (bias gradients db = sum_rows dY of GCNConv / GATConv); two-stage,
 * fixed order => deterministic.  workspace >= ghscn_colsum_workspace_bytes(). */
GHSCN_API size_t ghscn_colsum_workspace_bytes(int64_t num_rows, int64_t num_feat);
GHSCN_API int ghscn_colsum(const float* x, int64_t ldx, int64_t num_rows, int64_t num_feat, float* out,
                           void* workspace, size_t workspace_bytes, ghscn_stream_t stream);
/* Column sums of x (.) [mask > 0]: the bias gradient behind a ReLU that was fused into the forward aggregation
 * (mask = that layer's output), without materialising the masked gradient.  mask NULL = ghscn_colsum. */
GHSCN_API int ghscn_colsum_masked(const float* x, int64_t ldx, const float* mask, int64_t ldm, int64_t num_rows,
                                  int64_t num_feat, float* out, void* workspace, size_t workspace_bytes,
                                  ghscn_stream_t stream);

/* ReLU backward and the first stage of the bias gradient in ONE pass over dY (the backward of `.relu()` behind
 * GCNConv, model/hscn.py:110, followed by GCNConv's db = sum_rows dY): masked[r,f] = mask[r,f] > 0 ? x[r,f] : 0 (skipped
 * when masked is NULL) and the per-row-chunk column partials of the masked values in `workspace`;
 * ghscn_colsum_finish adds the partials in chunk order.  ghscn_colsum_masked = partial (masked NULL) + finish. */
GHSCN_API int ghscn_relu_grad_colsum_partial(const float* x, int64_t ldx, const float* mask, int64_t ldm,
                                             int64_t num_rows, int64_t num_feat, float* masked, int64_t ldo,
                                             void* workspace, size_t workspace_bytes, ghscn_stream_t stream);
GHSCN_API int ghscn_colsum_finish(const void* workspace, size_t workspace_bytes, int64_t num_rows, int64_t num_feat,
                                  float* out, ghscn_stream_t stream);

/* hi/lo split for the 3xTF32 GEMM scheme used by the layers' dense projections (x W^T of GCNConv / GATConv /
 * Linear): hi = x with the low 13 mantissa bits cleared (exact in TF32), lo = x - hi.  16-byte aligned buffers. */
GHSCN_API int ghscn_split_tf32(const float* x, int64_t n, float* hi, float* lo, ghscn_stream_t stream);
/* K-concatenated form: out[r,:] (3*num_cols wide) = [lo | hi | hi] (mode 0) or [hi | lo | hi] (mode 1) of row r;
 * rows num_rows..num_rows_padded-1 are zero.  A in mode 0 times W in mode 1 over the 3K-long reduction is the
 * 3xTF32 product x_lo.W_hi + x_hi.W_lo + x_hi.W_hi (small terms first) in ONE tensor-core GEMM. */
GHSCN_API int ghscn_split_tf32_cat(const float* x, int64_t ldx, int64_t num_rows, int64_t num_rows_padded,
                                   int64_t num_cols, int32_t mode, float* out, ghscn_stream_t stream);

/* ---- fused 3xTF32 projection GEMM on tcgen05 tensor cores (csrc/gemm3x.cu) -----------------------------------
 * C[m, n_out] = A[m, k] . B[n_out, k]^T (+ bias) (ReLU) with fp32-level accuracy: the h x h projections of
 * GCNConv / GATConv / Linear (PyG `Linear` inside the convs called at model/mpnn.py:52,59 and model/hscn.py:109;
 * SURVEY 8a rows a2, a9) and their input gradients.  A is read once as fp32 and split into TF32 hi/lo parts
 * in-kernel; the weights are split and laid out once per step by ghscn_gemm3x_prep_b:
 *   image = for each N half (<= 160 columns), for each 32-wide K chunk: [pad x 32] hi parts, [pad x 32] lo parts,
 *           K-major, 128B-swizzled, zero padded, ghscn_gemm3x_b_image_bytes() bytes, 16-byte aligned;
 *   b[n, k] = transpose ? w[k*ldw + n] : w[n*ldw + k]   (transpose = 1 gives dX = dY . W from W[out,in]).
 * Supported (ghscn_gemm3x_supported() != 0): n_out % 4 == 0, 16 <= n_out <= 320, k % 4 == 0, k >= 8;
 * ghscn_gemm3x additionally needs ldc == n_out, lda % 4 == 0 and 16-byte aligned a, c, bias. */
GHSCN_API int ghscn_gemm3x_supported(int64_t m, int64_t n_out, int64_t k);
GHSCN_API size_t ghscn_gemm3x_b_image_bytes(int64_t n_out, int64_t k);
GHSCN_API int ghscn_gemm3x_prep_b(const float* w, int64_t ldw, int64_t n_out, int64_t k, int32_t transpose,
                                  void* image, ghscn_stream_t stream);
/* debug builds only (-DGHSCN_GEMM3X_TRACE): device buffer (>= 1024 int64) receiving clock() stamps of the pipeline
 * events of CTA (0,0) of the tcgen05 kernels; NULL switches it off.  GHSCN_E_UNSUPPORTED in normal builds. */
GHSCN_API int ghscn_gemm3x_set_trace(void* device_buffer);
GHSCN_API int ghscn_gemm3x(const float* a, int64_t lda, int64_t m, int64_t k, const void* b_image, int64_t n_out,
                           const float* bias, int32_t relu, float* c, int64_t ldc, ghscn_stream_t stream);

/* Weight gradient of the same projections: out[m_out, n_out] = P[rows, m_out]^T . Q[rows, n_out] (P = dY, Q = x),
 * 3xTF32 on tcgen05 with both operands MN-major, split in-kernel.  The rows are cut into slabs of <= 1024; each
 * (128 x <=160 tile of out, slab) CTA writes an fp32 partial into `workspace`, then the partials are added in slab
 * order (deterministic).  Supported: m_out % 4 == 0, n_out % 4 == 0, 16 <= n_out <= 320, ld % 4 == 0, 16-byte
 * aligned pointers. */
GHSCN_API int ghscn_gemm3x_tn_supported(int64_t rows, int64_t m_out, int64_t n_out);
GHSCN_API size_t ghscn_gemm3x_tn_workspace_bytes(int64_t rows, int64_t m_out, int64_t n_out);
GHSCN_API int ghscn_gemm3x_tn(const float* p_mat, int64_t ldp, const float* q_mat, int64_t ldq, int64_t rows,
                              int64_t m_out, int64_t n_out, float* out, void* workspace, size_t workspace_bytes,
                              ghscn_stream_t stream);

/* Batch of independent products out[g] = P[rows_g]^T . Q[rows_g] over the row segments seg_ptr[g] .. seg_ptr[g+1]-1
 * (one CTA per segment x M tile x N half, same kernel as ghscn_gemm3x_tn).  This is the pooled-feature contraction
 * S^T X of dense_mincut_pool (model/hscn.py:63) per graph of a `ptr` batch, used where it is
 * dense-bound (K >= 64, profiles/r1_sweeps.md).  out[g] is [m_out, n_out] with row stride ldo at
 * out + g * out_segment_stride; empty segments give zeros.  max_segment_rows <= 1024 (accuracy bound per accumulator). */
GHSCN_API int ghscn_gemm3x_tn_segmented(const float* p_mat, int64_t ldp, const float* q_mat, int64_t ldq,
                                        const int32_t* seg_ptr, int64_t num_segments, int64_t max_segment_rows,
                                        int64_t m_out, int64_t n_out, float* out, int64_t ldo,
                                        int64_t out_segment_stride, ghscn_stream_t stream);

/* Tall-skinny projections y = x W^T + b with in_feat <= 32 (the 9 raw atom features -> hidden layers:
 * GCNConv layer 1, GraphConv lin_rel/lin_root, the SCN cluster MLP).  One streaming pass each, fp32 FMA.
 *   fwd: y [N,out];  dw: dW[out,in] = dY^T x (two-stage fixed-order reduction);  dx: dx [N,in] = dY W. */
GHSCN_API int ghscn_skinny_linear_fwd(const float* x, int64_t ldx, const float* w, const float* bias,
                                      int64_t num_rows, int64_t in_feat, int64_t out_feat, float* y, int64_t ldy,
                                      ghscn_stream_t stream);
GHSCN_API size_t ghscn_skinny_dw_workspace_bytes(int64_t num_rows, int64_t in_feat, int64_t out_feat);
GHSCN_API int ghscn_skinny_linear_dw(const float* dy, int64_t lddy, const float* x, int64_t ldx, int64_t num_rows,
                                     int64_t in_feat, int64_t out_feat, float* dw, void* workspace,
                                     size_t workspace_bytes, ghscn_stream_t stream);
GHSCN_API int ghscn_skinny_linear_dx(const float* dy, int64_t lddy, const float* w, int64_t num_rows,
                                     int64_t in_feat, int64_t out_feat, float* dx, int64_t lddx,
                                     ghscn_stream_t stream);

/* AdamW (torch.optim.AdamW semantics) on one flat fp32 parameter buffer.  `state` is THREE device floats owned by
 * the caller and zero-initialised: state[0] = number of steps taken so far (incremented by the call), state[1..2] =
 * the bias-correction scalars of the step being taken (derived on the device in double precision), so the call can be
 * captured in a CUDA graph.  Hyper-parameters are doubles, like the Python scalars torch.optim forms them from
 * (1 - beta2 in fp32 would be off by 1.3e-5 relative).
 * Replaces the per-tensor optimizer launches of train/train.py:94 for the flat-buffer training step. */
GHSCN_API int ghscn_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                               double lr, double beta1, double beta2, double eps, double weight_decay, float* state,
                               ghscn_stream_t stream);
/* Same update with every gradient multiplied by the device scalar grad_scale[0] as it is read (NULL = 1): the
 * clip coefficient of ghscn_grad_clip_scale, i.e. `clip_grad_norm_` followed by `optimizer.step()`. */
GHSCN_API int ghscn_adamw_step_scaled(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                                      double lr, double beta1, double beta2, double eps, double weight_decay,
                                      float* state, const float* grad_scale, ghscn_stream_t stream);
/* dst = concatenation of `count` fp32 device tensors (pointer and element-count arrays live on the HOST and are
 * consumed before the call returns): the gradients `loss.backward()` produced (train/train.py:87), laid out as the flat
 * buffer that the one-kernel optimizer step and the data-parallel all-reduce work on. */
GHSCN_API int ghscn_gather_flat(const float* const* srcs_host, const int64_t* numels_host, int32_t count, float* dst,
                                ghscn_stream_t stream);
/* nn.utils.clip_grad_norm(model.parameters(), max_norm) of train/train.py:92-93 over the flat gradient buffer:
 * out[0] = total 2-norm (fixed-order two-stage reduction, deterministic), out[1] = min(1, max_norm / (out[0] + 1e-6)).
 * The gradients themselves are not modified; pass out + 1 as grad_scale to ghscn_adamw_step_scaled. */
GHSCN_API size_t ghscn_grad_clip_workspace_bytes(int64_t n);
GHSCN_API int ghscn_grad_clip_scale(const float* grad, int64_t n, float max_norm, void* workspace,
                                    size_t workspace_bytes, float* out, ghscn_stream_t stream);

/* Fused forward of the same operator at the INPUT width (linearity of the pool in the source features):
 *   u_src = W_src^T att_src, u_dst = W_dst^T att_dst                     (ghscn_gat_fold_attention, W [out, in])
 *   pooled[r,:] = sum_s softmax_s(leaky_relu(<x_src[col[s]], u_src> + <x_dst[r], u_dst>)) x_src[col[s],:]
 * so that out = pooled W_src^T + bias is one small GEMM.  One warp per destination row, one pass over its members
 * (online softmax), replaces ghscn_row_dot x2 + ghscn_gat_scores + ghscn_spmm_pool.  x_dst / u_dst may both be NULL.
 * Supported widths: <= 32 (any alignment) or a multiple of 4 up to 512 with 16-byte aligned rows.
 * heads > 1 (GATConv(heads=h)): head i uses rows i*out_feat.. of W / att, u_src / u_dst are
 * [heads, feat] and pooled is [heads, num_rows, ldp]; all heads run in the same two launches.
 * warps_per_row (0 = 1, 2, 4, 8): warps that share a destination row of <= 32 members (their partial (max, sum,
 * weighted sum) are merged in warp order); callers pick it from the mean row length col.numel() / num_rows. */
GHSCN_API int ghscn_gat_fold_attention(const float* w_src, int64_t ldws, const float* att_src, const float* w_dst,
                                       int64_t ldwd, const float* att_dst, int64_t out_feat, int64_t src_feat,
                                       int64_t dst_feat, int64_t heads, float* u_src, float* u_dst,
                                       ghscn_stream_t stream);
GHSCN_API int ghscn_gat_pool_fused_supported(int64_t num_feat, int64_t ldxs, int64_t ldp);
GHSCN_API int ghscn_gat_pool_fused_fwd(const int32_t* rowptr, const int32_t* col, const float* x_src, int64_t ldxs,
                                       const float* x_dst, int64_t ldxd, const float* u_src, const float* u_dst,
                                       float negative_slope, int64_t num_rows, int64_t num_feat, int64_t heads,
                                       float* pooled, int64_t ldp, int32_t warps_per_row, ghscn_stream_t stream);

/* ---- small dense layers + task loss (readout head, virtual-node projections) --------------------------------------
 * Replaces the library GEMM + bias + activation kernels of `lin_1`, activation, `lin_2` on the [B, H] graph embeddings
 * (model/hscn.py:99-100,112) and of the virtual-node projections (model/hscn.py:85-93), problems of 10^2..10^3 rows.
 * fp32 FMA tiles, fixed reduction order.  act: 0 none, 1 ELU, 2 ReLU, 3 tanh (config/config.py:13-18).
 *   fwd: y = act(x W^T + bias)                     x [rows, in], W [out, in], y [rows, out]
 *   dx : dx = (dy (.) act'(y_ref)) W               y_ref = the forward's output (NULL with act 0)
 *   dw : dW = (dy (.) act'(y_ref))^T x, db = column sums of the same (db may be NULL); rows are split into <= 8 slabs
 *        whose partials are added in slab order (workspace needed only then) */
GHSCN_API int ghscn_small_linear_fwd(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias,
                                     int32_t act, int64_t num_rows, int64_t in_feat, int64_t out_feat, float* y,
                                     int64_t ldy, ghscn_stream_t stream);
/* y = act(x1 W1^T + bias1 + x2 W2^T + bias2): the sum of two layers with the same destination rows in one launch
 * (HeteroConv's `sum` over the v->v GCN and the l->v GAT pool, model/hscn.py:83-96; biases may be NULL). */
GHSCN_API int ghscn_small_linear2_fwd(const float* x1, int64_t ldx1, const float* w1, int64_t ldw1, const float* bias1,
                                      int64_t in_feat1, const float* x2, int64_t ldx2, const float* w2, int64_t ldw2,
                                      const float* bias2, int64_t in_feat2, int32_t act, int64_t num_rows,
                                      int64_t out_feat, float* y, int64_t ldy, ghscn_stream_t stream);
GHSCN_API int ghscn_small_linear_dx(const float* dy, int64_t lddy, const float* y_ref, int64_t ldy, int32_t act,
                                    const float* w, int64_t ldw, int64_t num_rows, int64_t in_feat, int64_t out_feat,
                                    float* dx, int64_t lddx, ghscn_stream_t stream);
GHSCN_API size_t ghscn_small_linear_dw_workspace_bytes(int64_t num_rows, int64_t in_feat, int64_t out_feat);
GHSCN_API int ghscn_small_linear_dw(const float* dy, int64_t lddy, const float* y_ref, int64_t ldy, int32_t act,
                                    const float* x, int64_t ldx, int64_t num_rows, int64_t in_feat, int64_t out_feat,
                                    float* dw, float* db, void* workspace, size_t workspace_bytes,
                                    ghscn_stream_t stream);
/* criterion(loss_fn, pred, true) of loss.py:6-19 over the first `rows` of `total_rows` rows (the rest are padding
 * graphs): mode 0 = binary_cross_entropy_with_logits, mode 1 = l1_loss, mean reduction.  loss [1]; d_pred
 * [total_rows, num_targets] = d loss / d pred (zero in the padding rows); score (may be NULL) = sigmoid(pred). */
GHSCN_API int ghscn_graph_loss(const float* pred, int64_t ldp, const float* target, int64_t ldt, int64_t rows,
                               int64_t total_rows, int64_t num_targets, int32_t mode, float* loss, float* d_pred,
                               float* score, ghscn_stream_t stream);

/* Output layer + task loss + their backward in ONE launch, for batches of <= 256 graphs (model/hscn.py:112 `lin_2`,
 * loss.py:6-19, train/train.py:82-87): pred = h W2^T + b2 [total_rows, C]; loss / score / d_pred as ghscn_graph_loss over
 * the first `rows` rows; d_w2 [C, H] = d_pred^T h, d_b2 [C], d_h [total_rows, H] = d_pred W2 -- all for d loss = 1 (the
 * caller scales by the incoming gradient).  hidden <= 512, num_targets <= 32. */
GHSCN_API int ghscn_head_out_loss_supported(int64_t total_rows, int64_t hidden, int64_t num_targets);
GHSCN_API int ghscn_head_out_loss(const float* h, int64_t ldh, const float* w2, int64_t ldw, const float* b2,
                                  const float* target, int64_t ldt, int64_t rows, int64_t total_rows, int64_t hidden,
                                  int64_t num_targets, int32_t mode, float* pred, float* loss, float* score,
                                  float* d_w2, float* d_b2, float* d_h, ghscn_stream_t stream);

/* ---- K5: bipartite GAT cluster pool (local -> virtual) ---------------------------------------
 * Replaces GATConv((-1,-1), H, add_self_loops=False) on ("local","to","virtual")
 * (model/hscn.py:85-87,118-125).  SURVEY 8a row a9, Appendix A.8.  heads = 1.
 *   a_src[j] = <hs[j,:], att_src>, a_dst[i] = <hd[i,:], att_dst>          (ghscn_row_dot)
 *   e = leaky_relu(a_src[col[s]] + a_dst[r]); alpha = segment softmax over row r (+1e-16)
 *   out[r,:] = sum_s alpha[s] * hs[col[s],:] + bias
 * alpha[nnz] is written per slot for the backward. a_dst may be NULL (no destination term).
 */
GHSCN_API int ghscn_row_dot(const float* x, int64_t ldx, const float* v, int64_t num_rows, int64_t num_feat,
                            float* out, ghscn_stream_t stream);
/* alpha[s] only (segment softmax of the leaky-relu scores); the pooled sum can then run at the INPUT width:
 * sum_s alpha_s (x_s W^T) = (sum_s alpha_s x_s) W^T, which removes the [N,F]x[F,H] projection of all local nodes. */
GHSCN_API int ghscn_gat_scores(const int32_t* rowptr, const int32_t* col, const float* a_src, const float* a_dst,
                               float negative_slope, int64_t num_rows, float* alpha, ghscn_stream_t stream);
GHSCN_API int ghscn_gat_pool_fwd(const int32_t* rowptr, const int32_t* col, const float* hs, int64_t ldhs,
                                 const float* a_src, const float* a_dst, const float* bias, float negative_slope,
                                 int64_t num_rows, int64_t num_feat, float* alpha, float* out, int64_t ldout,
                                 ghscn_stream_t stream);
/* backward part 1 (per destination row): dz[s] = d(loss)/d(pre-activation score of slot s),
 * da_dst[r] = sum_s dz[s].   part 2 is ghscn_spmm on the transposed structure with w = alpha_t,
 * plus the rank-1 term handled by the host. */
GHSCN_API int ghscn_gat_pool_bwd_scores(const int32_t* rowptr, const int32_t* col, const float* hs, int64_t ldhs,
                                        const float* a_src, const float* a_dst, const float* alpha,
                                        const float* dout, int64_t lddout, float negative_slope, int64_t num_rows,
                                        int64_t num_feat, float* dz, float* da_dst, ghscn_stream_t stream);
/* backward part 2 (per source row, transposed structure):
 *   da_src[j] = sum_t dz[map_t[t]];  dhs[j,:] = sum_t alpha[map_t[t]] * dout[col_t[t],:] + da_src[j] * att_src */
GHSCN_API int ghscn_gat_pool_bwd_src(const int32_t* rowptr_t, const int32_t* col_t, const int32_t* map_t,
                                     const float* alpha, const float* dz, const float* dout, int64_t lddout,
                                     const float* att_src, int64_t num_rows, int64_t num_feat, float* dhs,
                                     int64_t lddhs, float* da_src, ghscn_stream_t stream);
/* map_t[t] = slot, in the by-destination structure, of the edge stored in transposed slot t
 * (perm / perm_t are the slot -> edge-id arrays of the two orientations; scratch_pos[num_items]). */
GHSCN_API int ghscn_slot_map(const int32_t* perm, const int32_t* perm_t, int64_t nnz, int64_t num_items,
                             int32_t* scratch_pos, int32_t* map_t, ghscn_stream_t stream);

/* ---- fused node pipeline of the spectral-clustering net ----------------------------------------
 * Replaces, for SCN(mp_units=[U]) (model/hscn.py:30-45,57-60; main.py:101-106), the chain
 * GraphConv.propagate -> lin_rel -> + lin_root -> activation -> Linear that the reference runs as six launches,
 * once per training forward and once per cluster assignment (train/train_clustering.py:44,64-66):
 *   agg = A_w x (rows of (rowptr, col) are destinations; w NULL = unit weights; slot order, unfused mul+add),
 *   pre = W_rel agg + b_rel + W_root x,  h = act(pre) (0 identity, 1 ELU, 2 ReLU, 3 tanh),  logits = W_out h + b_out.
 * agg [N,f_in], pre [N,units], h [N,units] are written for the backward when non-NULL; logits [N,clusters].
 * GHSCN_E_UNSUPPORTED beyond f_in <= 16, units <= 32, clusters <= 32 (callers then run the separate operators). */
GHSCN_API int ghscn_scn_forward(const int32_t* rowptr, const int32_t* col, const float* w, const float* x, int64_t ldx,
                                int64_t num_nodes, int64_t f_in, int64_t units, int64_t clusters, const float* w_rel,
                                const float* b_rel, const float* w_root, const float* w_out, const float* b_out,
                                int32_t act, float* agg, float* pre, float* h, float* logits, ghscn_stream_t stream);

/* Parameter gradients of the pipeline above from ds = d loss / d logits and the saved agg / pre / h (the reference's
 * autograd chain through Linear, the activation and GraphConv's two Linears; train/train_clustering.py:48-50):
 *   grads = [ dW_out [clusters,units] | db_out [clusters] | dW_rel [units,f_in] | db_rel [units] | dW_root [units,f_in] ]
 * summed over the nodes in a fixed order (warp shuffles, warps in order, CTAs in order): deterministic.  x and the
 * graph structure get no gradient.  workspace: ghscn_scn_backward_workspace_bytes(). */
GHSCN_API size_t ghscn_scn_backward_workspace_bytes(int64_t num_nodes, int64_t f_in, int64_t units, int64_t clusters);
GHSCN_API int ghscn_scn_backward(const float* ds, const float* h, const float* pre, const float* agg, const float* x,
                                 int64_t ldx, int64_t num_nodes, int64_t f_in, int64_t units, int64_t clusters,
                                 const float* w_out, int32_t act, float* grads, void* workspace, size_t workspace_bytes,
                                 ghscn_stream_t stream);

/* ---- K6: fused MinCUT pool, one CTA per graph --------------------------------------------------
 * Replaces to_dense_adj + dense_mincut_pool (model/hscn.py:61-63).  SURVEY 8a rows a4, a5,
 * Appendix A.6/A.7.  The adjacency is consumed as the batch CSR whose ROWS ARE edge_index[0]
 * (A[r,c] = multiplicity or adj_val of edge (r,c)); adj_val == NULL means 1 per stored edge.
 * The backward also needs the other orientation (rows = edge_index[1]) for A^T S.
 * graph g owns nodes ptr[g]..ptr[g+1]-1; column ids are global node ids of the same graph.
 *   s_soft   [N,K]   softmax(logits)                     (written; saved for backward)
 *   out      [B,K,H] S^T X            (nullable: skipped)
 *   out_adj  [B,K,K] normalised, zero-diagonal S^T A S   (nullable)
 *   stats    [B,8]   per graph: num, den, ||SS||_F, ortho_g, mc_g, reserved...
 *   mc_loss / ortho_loss scalars are the batch means, reduced deterministically by
 *   ghscn_mincut_reduce_losses.
 * Pooled features for K <= 32 (round 2): when `out` is requested and x is a contiguous, 16-byte aligned [N,H] matrix
 * with H a multiple of 4 (H <= 1024), S^T X is computed behind the per-graph kernel by a batch-wide streaming kernel
 * (TMA bulk-copy ring; a thread-block cluster of 2/4/8 CTAs per graph when graphs are few); in the backward,
 * x g_out^T and d_x = S g_out come from one streaming pass over the rows (H >= 128) that parks x g_out^T in the
 * fourth [N,K] block of the workspace.  Same results up to summation order; without a workspace of
 * ghscn_mincut_workspace_bytes() the per-graph kernels do that work themselves.  GHSCN_MINCUT_STREAM_X=0 turns the
 * streaming kernels off.  A max_nodes_per_graph smaller than a graph's node count poisons that graph's losses,
 * pooled features and gradients with NaN instead of overrunning the shared tiles.
 */
GHSCN_API size_t ghscn_mincut_workspace_bytes(int64_t num_nodes, int64_t num_graphs, int64_t num_clusters);
GHSCN_API int ghscn_mincut_fwd(const float* logits, int64_t ldz, const float* x, int64_t ldx, const int32_t* ptr,
                               const int32_t* rowptr, const int32_t* col, const float* adj_val, float temp,
                               int64_t num_graphs, int64_t num_nodes, int64_t num_clusters, int64_t num_feat,
                               int32_t max_nodes_per_graph, float* s_soft, float* out, float* out_adj,
                               float* ss_raw, float* adj_raw, float* stats, float* losses /*[2]*/,
                               void* workspace, size_t workspace_bytes, ghscn_stream_t stream);
/* The same forward in two phases around ghscn_gemm3x_tn_segmented, for the dense-bound corner (K >= 64):
 *   phase 1  S = softmax, A S (-> workspace, [N,K] fp32), the traces num / den (-> stats); no contraction;
 *   caller   ss_raw = S^T S, adj_raw = S^T (A S), out = S^T X per graph on the tensor cores;
 *   phase 2  norms, orthogonality loss, normalised coarse adjacency from ss_raw / adj_raw, then the loss means.
 *   phase 3  = phase 1 with A S computed by one launch of the K2 SpMM over the whole batch (valid for collated
 *            batches: no edge between graphs) and num left to phase 2, which takes it as Tr(adj_raw) -- literally
 *            dense_mincut_pool's `_rank3_trace(out_adj)`.
 * phase 0 == ghscn_mincut_fwd.  Arguments as above; out is ignored by phases 1, 2 and 3. */
GHSCN_API int ghscn_mincut_fwd_phase(const float* logits, int64_t ldz, const float* x, int64_t ldx, const int32_t* ptr,
                                     const int32_t* rowptr, const int32_t* col, const float* adj_val, float temp,
                                     int64_t num_graphs, int64_t num_nodes, int64_t num_clusters, int64_t num_feat,
                                     int32_t max_nodes_per_graph, float* s_soft, float* out, float* out_adj,
                                     float* ss_raw, float* adj_raw, float* stats, float* losses /*[2]*/,
                                     void* workspace, size_t workspace_bytes, ghscn_stream_t stream, int32_t phase);
GHSCN_API int ghscn_mincut_bwd(const float* s_soft, const float* x, int64_t ldx, const int32_t* ptr,
                               const int32_t* rowptr, const int32_t* col, const float* adj_val,
                               const int32_t* rowptr_t, const int32_t* col_t, const float* adj_val_t, float temp,
                               int64_t num_graphs, int64_t num_nodes, int64_t num_clusters, int64_t num_feat,
                               int32_t max_nodes_per_graph, const float* ss_raw, const float* adj_raw,
                               const float* stats, const float* g_out /*[B,K,H]|NULL*/,
                               const float* g_out_adj /*[B,K,K]|NULL*/, const float* g_losses /*[2] device*/,
                               float* d_logits, int64_t lddz, float* d_x /*nullable*/, int64_t lddx,
                               void* workspace, size_t workspace_bytes, ghscn_stream_t stream);

/* The same backward in split form for the dense-bound corner (K >= 64, K % 4 == 0, num_feat % 4 == 0, 16-byte aligned
 * rows): one pass per graph leaves [A S | A^T S | S], the stacked [Gamma^T; Gamma; Gsym] and the elementwise part of
 * dS in the workspace; the [n,3K]x[3K,K], [n,H]x[H,K] (x g_out^T) and [n,K]x[K,H] (dX = S g_out) products run as tiled
 * per-graph GEMMs (64 x 64 tiles over every graph at once instead of one CTA walking a graph's products serially);
 * a last pass is the softmax backward.  Same arguments and results as ghscn_mincut_bwd. */
GHSCN_API size_t ghscn_mincut_bwd_split_workspace_bytes(int64_t num_nodes, int64_t num_graphs, int64_t num_clusters);
GHSCN_API int ghscn_mincut_bwd_split_supported(int64_t num_clusters, int64_t num_feat, int64_t ldx, int64_t lddx,
                                               int32_t max_nodes_per_graph);
GHSCN_API int ghscn_mincut_bwd_split(const float* s_soft, const float* x, int64_t ldx, const int32_t* ptr,
                                     const int32_t* rowptr, const int32_t* col, const float* adj_val,
                                     const int32_t* rowptr_t, const int32_t* col_t, const float* adj_val_t, float temp,
                                     int64_t num_graphs, int64_t num_nodes, int64_t num_clusters, int64_t num_feat,
                                     int32_t max_nodes_per_graph, const float* ss_raw, const float* adj_raw,
                                     const float* stats, const float* g_out /*[B,K,H]|NULL*/,
                                     const float* g_out_adj /*[B,K,K]|NULL*/, const float* g_losses /*[2] device*/,
                                     float* d_logits, int64_t lddz, float* d_x /*nullable*/, int64_t lddx,
                                     void* workspace, size_t workspace_bytes, ghscn_stream_t stream);

/* ---- K7: cluster assignment -> virtual-node construction --------------------------------------
 * Replaces softmax(s).max(1)[1] (train/train_clustering.py:68) and the Python/numpy loops of
 * loader/hetero_data.py:44-86.  SURVEY 8a rows a6, a7.  One CTA per graph, integer work bit-exact:
 *   cluster[i]   first-max index of row i of s_soft (int32)
 *   remap to 0..U-1 by sorted unique; bucket slot (c-1) mod K (negative index quirk, :53);
 *   non-empty slots compacted -> virtual feature rows (float64 mean of raw features -> fp32);
 *   lv edge (i, c_i); vv edges {(i,j): i+j <= U-1} in the reference's [col;row] order.
 * Padded outputs (static shapes, K virtual slots per graph) + per-graph U; the compaction to the
 * reference's ragged layout is ghscn_virtual_compact.
 */
GHSCN_API int ghscn_cluster_argmax(const float* s_soft, int64_t lds, int64_t num_nodes, int64_t num_clusters,
                                   int32_t* cluster, ghscn_stream_t stream);
GHSCN_API int ghscn_virtual_build(const int32_t* cluster, const int32_t* ptr, const void* x_raw, int32_t x_is_int64,
                                  int64_t ldx, int64_t num_graphs, int64_t num_clusters, int64_t num_feat,
                                  int32_t* cluster_remapped /*[N] 0..U-1*/, int32_t* num_virtual /*[B] U_g*/,
                                  float* virt_x_padded /*[B*K,F]*/, ghscn_stream_t stream);
GHSCN_API int ghscn_virtual_edges(const int32_t* cluster_remapped, const int32_t* ptr, const int32_t* num_virtual,
                                  const int32_t* virt_offset /*[B+1] exclusive scan of U_g, or NULL => g*K*/,
                                  int64_t num_graphs, int64_t num_nodes, int64_t num_clusters,
                                  int64_t* lv_edge_index /*[2,N]*/, int64_t* vv_edge_index /*[2,vv_cap]*/,
                                  int64_t vv_cap, const int32_t* vv_offset /*[B+1] or NULL => g*K(K+1)/2*/,
                                  ghscn_stream_t stream);
GHSCN_API int ghscn_virtual_offsets(const int32_t* num_virtual, int64_t num_graphs, int32_t* virt_offset /*[B+1]*/,
                                    int32_t* vv_offset /*[B+1]*/, ghscn_stream_t stream);
/* Both orientations of the l->v and v->v relations straight from the cluster assignment (no sort):
 * lvd_* rows = virtual nodes, lvs_* rows = local nodes, vvd_* / vvs_* rows = virtual nodes (by destination /
 * by source).  vv_offset[B+1] is the exclusive scan of U_g(U_g+1)/2 (ghscn_virtual_offsets); with padded != 0
 * virtual ids are g*K + j and v->v edge ids g*K(K+1)/2 + p, otherwise the compact virt_offset / vv_offset ids.
 * Equals ghscn_csr_build on the edge lists of ghscn_virtual_edges. */
GHSCN_API int ghscn_virtual_csr(const int32_t* cluster_remapped, const int32_t* ptr, const int32_t* num_virtual,
                                const int32_t* virt_offset, const int32_t* vv_offset, int64_t num_graphs,
                                int64_t num_clusters, int32_t padded, int32_t* lvd_rowptr, int32_t* lvd_col,
                                int32_t* lvd_perm, int32_t* lvs_rowptr, int32_t* lvs_col, int32_t* lvs_perm,
                                int32_t* vvd_rowptr, int32_t* vvd_col, int32_t* vvd_perm, int32_t* vvs_rowptr,
                                int32_t* vvs_col, int32_t* vvs_perm, ghscn_stream_t stream);
GHSCN_API int ghscn_virtual_compact(const float* virt_x_padded, const int32_t* num_virtual,
                                    const int32_t* virt_offset, int64_t num_graphs, int64_t num_clusters,
                                    int64_t num_feat, float* virt_x /*[V,F]*/, int64_t* virt_batch /*[V]*/,
                                    ghscn_stream_t stream);

/* ---- Laplacian positional encoding -------------------------------------------
 * Replaces the per-graph host loop of transform/posenc.py:14-82 (get_laplacian -> scipy toarray ->
 * np.linalg.eigh -> get_lap_decomp_stats :50-82 -> eigvec_normalizer :85-108) for a whole collated batch:
 * one CTA per graph builds the dense Laplacian from the graph's CSR slice (rows = edge_index[0], float32
 * entries like PyG's, lower triangle mirrored like LAPACK's UPLO='L'), diagonalises it with a float64
 * one-sided Jacobi iteration and writes, per node, the `max_freqs` smallest eigenvalues (clamped at 0) and
 * the matching normalised eigenvector entries; graphs with fewer nodes than max_freqs get NaN padding.
 *   laplacian_norm: 0 = none (D - A), 1 = sym, 2 = rw (posenc.py:33-41);  symmetrize != 0 = to_undirected
 *   (posenc.py:28-31) with coalescing;  eigvec_norm: 0 = L1, 1 = L2, 2 = abs-max.
 * Eigenvector signs / bases of repeated eigenvalues are LAPACK's choice in the reference and not reproduced.
 * sweeps [B] (nullable) receives the Jacobi sweep count per graph (40 = not converged).  B <= 65535 per call. */
GHSCN_API size_t ghscn_laplacian_eig_workspace_bytes(int64_t num_nodes, int64_t num_graphs,
                                                    int32_t max_nodes_per_graph);
GHSCN_API int ghscn_laplacian_eig(const int32_t* ptr, const int32_t* rowptr, const int32_t* col, int64_t num_graphs,
                                  int64_t num_nodes, int32_t max_nodes_per_graph, int32_t laplacian_norm,
                                  int32_t symmetrize, int32_t max_freqs, int32_t eigvec_norm,
                                  float* eigvals /*[N,max_freqs]*/, float* eigvecs /*[N,max_freqs]*/,
                                  int32_t* sweeps /*[B]|NULL*/, void* workspace, size_t workspace_bytes,
                                  ghscn_stream_t stream);

/* ---- small fused elementwise epilogues ------------------------------------ */
GHSCN_API int ghscn_cast_i64_f32(const int64_t* in, int64_t n, float* out, ghscn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GHSCN_H_ */
