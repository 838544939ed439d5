"""torch custom ops (`torch.ops.ghscn.*`) over the C-ABI kernels of libghscn.so.

Each op is a thin marshalling layer: it allocates outputs with PyTorch (caller-owns-memory contract
of include/ghscn.h), passes raw device pointers + the current CUDA stream through ctypes, and wires
the analytic backward with `torch.library.register_autograd`.  CUDA only -- there is no CPU
implementation registered, so a CPU tensor fails loudly in the dispatcher.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import os

import torch
from torch import Tensor

from ._lib import lib
from .structure import _p, _stream

_custom_op = torch.library.custom_op
_STATS_STRIDE = 8


PARALLEL_AUX = os.environ.get("GHSCN_PARALLEL_AUX", "1") != "0"
# measured in the step: 0.641 ms fused vs 0.637 ms separate (the step is bound by SM occupancy, not by the length of
# the backward chain; the separate column sum already overlaps the transposed aggregation) -- kept selectable
FUSED_RELU_COLSUM = os.environ.get("GHSCN_FUSED_RELU_COLSUM", "0") != "0"
_AUX_STREAMS: dict = {}


def _aux_stream(device: torch.device) -> "torch.cuda.Stream":
    """One auxiliary stream per (device, current stream): independent side work inside a backward function."""
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    st = _AUX_STREAMS.get(key)
    if st is None:
        st = _AUX_STREAMS[key] = torch.cuda.Stream(device=device)
    return st


def _rowmajor(x: Tensor) -> Tensor:
    if x.dim() != 2:
        raise ValueError("expected a 2-D feature matrix")
    if x.dtype != torch.float32:
        raise TypeError(f"ghscn kernels compute in fp32, got {x.dtype}")
    if x.stride(1) != 1 or x.stride(0) < x.size(1):
        x = x.contiguous()
    return x


# =============================================================================================
# column sum (bias gradients)
# =============================================================================================
@_custom_op("ghscn::colsum", mutates_args=(), device_types="cuda")
def colsum(x: Tensor) -> Tensor:
    x = _rowmajor(x)
    N, F = x.shape
    out = torch.empty(F, dtype=torch.float32, device=x.device)
    L = lib()
    ws_bytes = L.query("ghscn_colsum_workspace_bytes", N, F)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    L.call("ghscn_colsum", _p(x), x.stride(0), N, F, _p(out), _p(ws), ws_bytes, _stream())
    return out


@colsum.register_fake
def _(x):
    return x.new_empty((x.size(1),))


@_custom_op("ghscn::colsum_masked", mutates_args=(), device_types="cuda")
def colsum_masked(x: Tensor, mask: Tensor) -> Tensor:
    """Column sums of x (.) [mask > 0] (bias gradient behind a fused ReLU) without materialising the product."""
    x, mask = _rowmajor(x), _rowmajor(mask)
    N, F = x.shape
    out = torch.empty(F, dtype=torch.float32, device=x.device)
    L = lib()
    ws_bytes = L.query("ghscn_colsum_workspace_bytes", N, F)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    L.call("ghscn_colsum_masked", _p(x), x.stride(0), _p(mask), mask.stride(0), N, F, _p(out), _p(ws), ws_bytes,
           _stream())
    return out


@colsum_masked.register_fake
def _(x, mask):
    return x.new_empty((x.size(1),))


def relu_grad_colsum(dy: Tensor, y: Tensor, fork: bool = True):
    """-> (dy (.) [y > 0], its column sums): ReLU backward and GCNConv's bias gradient from ONE pass over dy (the
    separate threshold + column-sum kernels read dy twice).  The tiny second reduction stage runs on the auxiliary
    stream when `fork` is set; the caller joins with `join()` after launching what follows on its own stream."""
    dy, y = _rowmajor(dy), _rowmajor(y)
    N, F = dy.shape
    dev = dy.device
    out = torch.empty((N, F), dtype=torch.float32, device=dev)
    L = lib()
    ws_bytes = L.query("ghscn_colsum_workspace_bytes", N, F)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    L.call("ghscn_relu_grad_colsum_partial", _p(dy), dy.stride(0), _p(y), y.stride(0), N, F, _p(out), F, _p(ws),
           ws_bytes, _stream())
    main = torch.cuda.current_stream()
    side = _aux_stream(dev) if fork and PARALLEL_AUX else None
    if side is not None:
        side.wait_stream(main)
    with torch.cuda.stream(side if side is not None else main):
        db = torch.empty(F, dtype=torch.float32, device=dev)
        L.call("ghscn_colsum_finish", _p(ws), ws_bytes, N, F, _p(db), _stream())
    if side is not None:
        db.record_stream(main)
        ws.record_stream(side)
    return out, db, (lambda: main.wait_stream(side)) if side is not None else (lambda: None)


# =============================================================================================
# K2/K3  SpMM
# =============================================================================================
@_custom_op("ghscn::spmm_raw", mutates_args=(), device_types="cuda")
def spmm_raw(rowptr: Tensor, col: Tensor, w: Optional[Tensor], x: Tensor, bias: Optional[Tensor],
             num_rows: int, relu: bool) -> Tensor:
    x = _rowmajor(x)
    F = x.size(1)
    y = torch.empty((num_rows, F), dtype=torch.float32, device=x.device)
    lib().call("ghscn_spmm", _p(rowptr), _p(col), _p(w), _p(x), x.stride(0), _p(y), F, _p(bias), num_rows, F,
               int(relu), _stream())
    return y


@spmm_raw.register_fake
def _(rowptr, col, w, x, bias, num_rows, relu):
    return x.new_empty((num_rows, x.size(1)))


@_custom_op("ghscn::spmm_masked", mutates_args=(), device_types="cuda")
def spmm_masked(rowptr: Tensor, col: Tensor, w: Optional[Tensor], x: Tensor, mask: Tensor, num_rows: int) -> Tensor:
    """A_w (x (.) [mask > 0]): the rows of x are masked as they are gathered (ReLU backward fused into the transposed
    aggregation)."""
    x, mask = _rowmajor(x), _rowmajor(mask)
    F = x.size(1)
    y = torch.empty((num_rows, F), dtype=torch.float32, device=x.device)
    lib().call("ghscn_spmm_masked", _p(rowptr), _p(col), _p(w), _p(x), x.stride(0), _p(mask), mask.stride(0), _p(y), F,
               num_rows, F, _stream())
    return y


@spmm_masked.register_fake
def _(rowptr, col, w, x, mask, num_rows):
    return x.new_empty((num_rows, x.size(1)))


# measured slower than the separate threshold pass at degree ~2 (the mask rows double the gather traffic, which is what
# bounds the aggregation kernel: +17 us per step), so it is off by default; kept for graphs where rows are read once
FUSED_RELU_BACKWARD = os.environ.get("GHSCN_FUSED_RELU_BACKWARD", "0") != "0"


@_custom_op("ghscn::spmm_edge_grad", mutates_args=(), device_types="cuda")
def spmm_edge_grad(rowptr: Tensor, col: Tensor, x: Tensor, dy: Tensor, num_slots: int) -> Tensor:
    x, dy = _rowmajor(x), _rowmajor(dy)
    dw = torch.zeros(num_slots, dtype=torch.float32, device=x.device)
    lib().call("ghscn_spmm_edge_grad", _p(rowptr), _p(col), None, _p(x), x.stride(0), _p(dy), dy.stride(0),
               dy.size(0), x.size(1), num_slots, _p(dw), _stream())
    return dw


@spmm_edge_grad.register_fake
def _(rowptr, col, x, dy, num_slots):
    return x.new_empty((num_slots,))


@_custom_op("ghscn::spmm", mutates_args=(), device_types="cuda")
def spmm(rowptr: Tensor, col: Tensor, w: Optional[Tensor], rowptr_t: Tensor, col_t: Tensor,
         w_t: Optional[Tensor], x: Tensor, bias: Optional[Tensor], relu: bool = False) -> Tensor:
    """y = act(A_w x + bias); rows of (rowptr, col) are destinations, (rowptr_t, col_t) is the transpose;
    act = ReLU when `relu` (fused into the kernel's epilogue), identity otherwise."""
    return spmm_raw(rowptr, col, w, x, bias, rowptr.numel() - 1, relu)


@spmm.register_fake
def _(rowptr, col, w, rowptr_t, col_t, w_t, x, bias, relu=False):
    return x.new_empty((rowptr.numel() - 1, x.size(1)))


def _spmm_setup(ctx, inputs, output):
    rowptr, col, w, rowptr_t, col_t, w_t, x, bias, relu = inputs
    ctx.has_bias = bias is not None
    ctx.relu = relu
    ctx.w_needs_grad = w is not None and w.requires_grad
    ctx.save_for_backward(rowptr, col, rowptr_t, col_t, w_t, x if ctx.w_needs_grad else None,
                          output if relu else None)


def _spmm_backward(ctx, dy):
    rowptr, col, rowptr_t, col_t, w_t, x, y = ctx.saved_tensors
    dy = dy.contiguous()
    # ReLU backward: folded into the loads of the two consumers of dy (transposed aggregation, bias column sum) where
    # the kernels support it, instead of a separate pass that reads dy and y and writes dy (.) [y > 0]
    masked = (ctx.relu and FUSED_RELU_BACKWARD and not ctx.w_needs_grad and dy.is_cuda and dy.dim() == 2
              and bool(lib().query("ghscn_spmm_masked_supported", dy.size(1), dy.stride(0), y.stride(0), dy.size(1))))
    dx = dw = dbias = None
    want_bias = ctx.has_bias and ctx.needs_input_grad[7]
    join = None
    if ctx.relu and not masked:
        if want_bias and FUSED_RELU_COLSUM and dy.is_cuda and dy.dim() == 2:
            dy, dbias, join = relu_grad_colsum(dy, y, fork=bool(ctx.needs_input_grad[6]))
            want_bias = False
        else:
            dy = torch.ops.aten.threshold_backward(dy, y, 0.0)      # dy * (y > 0), one vectorised pass
    # The bias gradient (a column sum of dy) and the transposed aggregation both only read dy: the column sum runs on
    # an auxiliary stream, forked here and joined before returning (graph-capturable; same kernels, same results).
    side = main = None
    if want_bias and ctx.needs_input_grad[6] and PARALLEL_AUX and dy.is_cuda:
        main = torch.cuda.current_stream()
        side = _aux_stream(dy.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            dbias = colsum_masked(dy, y) if masked else colsum(dy)
        dbias.record_stream(main)           # allocated on the auxiliary stream, consumed on the caller's
    if ctx.needs_input_grad[6]:
        if masked:
            dx = spmm_masked(rowptr_t, col_t, w_t, dy, y, rowptr_t.numel() - 1)
        else:
            dx = spmm_raw(rowptr_t, col_t, w_t, dy, None, rowptr_t.numel() - 1, False)
    if ctx.w_needs_grad:
        dw = spmm_edge_grad(rowptr, col, x, dy, col.numel())
    if side is not None:
        main.wait_stream(side)
    elif want_bias:
        dbias = colsum_masked(dy, y) if masked else colsum(dy)
    if join is not None:
        join()
    return None, None, dw, None, None, None, dx, dbias, None


torch.library.register_autograd("ghscn::spmm", _spmm_backward, setup_context=_spmm_setup)


# =============================================================================================
# fused SCN node pipeline: GraphConv aggregation + lin_rel + lin_root + activation + cluster Linear
# =============================================================================================
MINCUT_TC_MIN_K = int(os.environ.get("GHSCN_MINCUT_TC_MIN_K", "32"))
MINCUT_TC_PHASE1 = int(os.environ.get("GHSCN_MINCUT_TC_PHASE1", "3"))      # 3: A S by the batch SpMM; 1: per-graph CTAs
MINCUT_SPLIT_MIN_K = int(os.environ.get("GHSCN_MINCUT_SPLIT_MIN_K", "32"))   # backward as per-graph tiled GEMMs
MINCUT_TC_KK = os.environ.get("GHSCN_MINCUT_TC_KK", "1") != "0"      # S^T S and S^T A S on the tensor cores too
FUSED_SCN_BACKWARD = os.environ.get("GHSCN_FUSED_SCN_BACKWARD", "1") != "0"
SCN_ACTS = {"identity": 0, "elu": 1, "relu": 2, "tanh": 3}
SCN_LIMITS = (16, 32, 32)          # f_in, units, clusters handled by ghscn_scn_forward


class ScnNodeForward(torch.autograd.Function):
    """logits = W_out act(W_rel (A_w x) + b_rel + W_root x) + b_out in ONE launch (csrc/scn.cu); `x` and the structure
    carry no gradient (raw atom features, gcn_norm weights: train/train_clustering.py:37-47).  The backward is the
    chain of the separate operators' backwards, on the saved agg / pre / h."""

    @staticmethod
    def forward(ctx, x, rowptr, col, w, w_rel, b_rel, w_root, w_out, b_out, act: int):
        n, f = x.shape
        u, k = w_rel.size(0), w_out.size(0)
        dev = x.device
        need = any(ctx.needs_input_grad[4:9])          # (grad mode is off inside Function.forward: ask the context)
        agg = torch.empty((n, f), dtype=torch.float32, device=dev) if need else None
        pre = torch.empty((n, u), dtype=torch.float32, device=dev) if need else None
        h = torch.empty((n, u), dtype=torch.float32, device=dev) if need else None
        logits = torch.empty((n, k), dtype=torch.float32, device=dev)
        lib().call("ghscn_scn_forward", _p(rowptr), _p(col), _p(w), _p(x), x.stride(0), n, f, u, k,
                   _p(w_rel), _p(b_rel), _p(w_root), _p(w_out), _p(b_out), int(act),
                   _p(agg), _p(pre), _p(h), _p(logits), _stream())
        ctx.save_for_backward(x, agg, pre, h, w_out)
        ctx.act = int(act)
        ctx.has = (b_rel is not None, b_out is not None)
        return logits

    @staticmethod
    def backward(ctx, ds):
        x, agg, pre, h, w_out = ctx.saved_tensors
        ds = ds.contiguous()
        if FUSED_SCN_BACKWARD:
            # one launch + a fixed-order reduction instead of three dW kernels, two column sums, dX and act'
            n, f = x.shape
            u, k = w_out.size(1), w_out.size(0)
            L = lib()
            grads = torch.empty(k * u + k + 2 * u * f + u, dtype=torch.float32, device=x.device)
            ws_bytes = L.query("ghscn_scn_backward_workspace_bytes", n, f, u, k)
            ws = torch.empty(max(ws_bytes, 4), dtype=torch.uint8, device=x.device)
            L.call("ghscn_scn_backward", _p(ds), _p(h), _p(pre), _p(agg), _p(x), x.stride(0), n, f, u, k,
                   _p(w_out.contiguous()), ctx.act, _p(grads), _p(ws), ws_bytes, _stream())
            d_wout, d_bout, d_wrel, d_brel, d_wroot = torch.split(grads, [k * u, k, u * f, u, u * f])
            return (None, None, None, None, d_wrel.view(u, f), d_brel if ctx.has[0] else None, d_wroot.view(u, f),
                    d_wout.view(k, u), d_bout if ctx.has[1] else None, None)
        from .gemm import skinny_dw, skinny_dx
        d_wout = skinny_dw(ds, h)
        d_bout = colsum(ds) if ctx.has[1] else None
        dh = skinny_dx(ds, w_out.contiguous())
        if ctx.act == 1:
            dpre = torch.ops.aten.elu_backward(dh, 1.0, 1.0, 1.0, False, pre)
        elif ctx.act == 2:
            dpre = torch.ops.aten.threshold_backward(dh, pre, 0.0)
        elif ctx.act == 3:
            dpre = dh * (1.0 - h * h)
        else:
            dpre = dh
        d_wrel = skinny_dw(dpre, agg)
        d_brel = colsum(dpre) if ctx.has[0] else None
        d_wroot = skinny_dw(dpre, x)
        return None, None, None, None, d_wrel, d_brel, d_wroot, d_wout, d_bout, None


def scn_node_forward(x: Tensor, rowptr: Tensor, col: Tensor, w: Optional[Tensor], w_rel: Tensor,
                     b_rel: Optional[Tensor], w_root: Tensor, w_out: Tensor, b_out: Optional[Tensor], act: str) -> Tensor:
    return ScnNodeForward.apply(x, rowptr, col, w, w_rel.contiguous(), b_rel, w_root.contiguous(),
                                w_out.contiguous(), b_out, SCN_ACTS[act])


# =============================================================================================
# K4  segment mean / sum
# =============================================================================================
@_custom_op("ghscn::segment_broadcast", mutates_args=(), device_types="cuda")
def segment_broadcast(dy: Tensor, ptr: Tensor, perm: Optional[Tensor], num_rows: int, mean: bool) -> Tensor:
    dy = _rowmajor(dy)
    F = dy.size(1)
    if perm is None:
        dx = torch.empty((num_rows, F), dtype=torch.float32, device=dy.device)
    else:
        dx = torch.zeros((num_rows, F), dtype=torch.float32, device=dy.device)
    lib().call("ghscn_segment_broadcast", _p(dy), dy.stride(0), _p(ptr), _p(perm), ptr.numel() - 1, F, int(mean),
               _p(dx), F, _stream())
    return dx


@segment_broadcast.register_fake
def _(dy, ptr, perm, num_rows, mean):
    return dy.new_empty((num_rows, dy.size(1)))


@_custom_op("ghscn::segment_reduce", mutates_args=(), device_types="cuda")
def segment_reduce(x: Tensor, ptr: Tensor, perm: Optional[Tensor], mean: bool) -> Tensor:
    x = _rowmajor(x)
    B, F = ptr.numel() - 1, x.size(1)
    y = torch.empty((B, F), dtype=torch.float32, device=x.device)
    lib().call("ghscn_segment_reduce", _p(x), x.stride(0), _p(ptr), _p(perm), B, F, int(mean), _p(y), F, _stream())
    return y


@segment_reduce.register_fake
def _(x, ptr, perm, mean):
    return x.new_empty((ptr.numel() - 1, x.size(1)))


def _seg_setup(ctx, inputs, output):
    x, ptr, perm, mean = inputs
    ctx.mean, ctx.num_rows = mean, x.size(0)
    ctx.save_for_backward(ptr, perm)


def _seg_backward(ctx, dy):
    ptr, perm = ctx.saved_tensors
    return segment_broadcast(dy.contiguous(), ptr, perm, ctx.num_rows, ctx.mean), None, None, None


torch.library.register_autograd("ghscn::segment_reduce", _seg_backward, setup_context=_seg_setup)


# =============================================================================================
# K5  bipartite GAT pool
# =============================================================================================
@_custom_op("ghscn::row_dot", mutates_args=(), device_types="cuda")
def row_dot(x: Tensor, v: Tensor) -> Tensor:
    x = _rowmajor(x)
    v = v.contiguous()
    out = torch.empty(x.size(0), dtype=torch.float32, device=x.device)
    lib().call("ghscn_row_dot", _p(x), x.stride(0), _p(v), x.size(0), x.size(1), _p(out), _stream())
    return out


@row_dot.register_fake
def _(x, v):
    return x.new_empty((x.size(0),))


@_custom_op("ghscn::gat_pool_fwd", mutates_args=(), device_types="cuda")
def gat_pool_fwd(rowptr: Tensor, col: Tensor, hs: Tensor, a_src: Tensor, a_dst: Optional[Tensor],
                 bias: Optional[Tensor], slope: float, num_slots: int) -> Tuple[Tensor, Tensor]:
    hs = _rowmajor(hs)
    V, H = rowptr.numel() - 1, hs.size(1)
    alpha = torch.zeros(num_slots, dtype=torch.float32, device=hs.device)
    out = torch.empty((V, H), dtype=torch.float32, device=hs.device)
    lib().call("ghscn_gat_pool_fwd", _p(rowptr), _p(col), _p(hs), hs.stride(0), _p(a_src), _p(a_dst), _p(bias),
               float(slope), V, H, _p(alpha), _p(out), H, _stream())
    return out, alpha


@gat_pool_fwd.register_fake
def _(rowptr, col, hs, a_src, a_dst, bias, slope, num_slots):
    return hs.new_empty((rowptr.numel() - 1, hs.size(1))), hs.new_empty((num_slots,))


@_custom_op("ghscn::gat_pool_bwd", mutates_args=(), device_types="cuda")
def gat_pool_bwd(rowptr: Tensor, col: Tensor, rowptr_t: Tensor, col_t: Tensor, map_t: Tensor, hs: Tensor,
                 a_src: Tensor, a_dst: Optional[Tensor], alpha: Tensor, att_src: Tensor, dout: Tensor,
                 slope: float) -> Tuple[Tensor, Tensor, Tensor]:
    """-> (dhs [N,H], da_src [N], da_dst [V])"""
    hs, dout = _rowmajor(hs), _rowmajor(dout)
    N, H, V = hs.size(0), hs.size(1), rowptr.numel() - 1
    dev = hs.device
    dz = torch.zeros(alpha.numel(), dtype=torch.float32, device=dev)
    da_dst = torch.empty(V, dtype=torch.float32, device=dev)
    L, st = lib(), _stream()
    L.call("ghscn_gat_pool_bwd_scores", _p(rowptr), _p(col), _p(hs), hs.stride(0), _p(a_src), _p(a_dst), _p(alpha),
           _p(dout), dout.stride(0), float(slope), V, H, _p(dz), _p(da_dst), st)
    dhs = torch.empty((N, H), dtype=torch.float32, device=dev)
    da_src = torch.empty(N, dtype=torch.float32, device=dev)
    L.call("ghscn_gat_pool_bwd_src", _p(rowptr_t), _p(col_t), _p(map_t), _p(alpha), _p(dz), _p(dout),
           dout.stride(0), _p(att_src.contiguous()), N, H, _p(dhs), H, _p(da_src), st)
    return dhs, da_src, da_dst


@gat_pool_bwd.register_fake
def _(rowptr, col, rowptr_t, col_t, map_t, hs, a_src, a_dst, alpha, att_src, dout, slope):
    return hs.new_empty(hs.shape), hs.new_empty((hs.size(0),)), hs.new_empty((rowptr.numel() - 1,))


@_custom_op("ghscn::gat_pool", mutates_args=(), device_types="cuda")
def gat_pool(rowptr: Tensor, col: Tensor, rowptr_t: Tensor, col_t: Tensor, map_t: Tensor, hs: Tensor,
             hd: Optional[Tensor], att_src: Tensor, att_dst: Optional[Tensor], bias: Optional[Tensor],
             slope: float) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Bipartite single-head GAT aggregation. -> (out [V,H], alpha [slots], a_src [N], a_dst [V])"""
    a_src = row_dot(hs, att_src)
    a_dst = row_dot(hd, att_dst) if hd is not None else torch.zeros(rowptr.numel() - 1, device=hs.device)
    out, alpha = gat_pool_fwd(rowptr, col, hs, a_src, a_dst, bias, slope, col.numel())
    return out, alpha, a_src, a_dst


@gat_pool.register_fake
def _(rowptr, col, rowptr_t, col_t, map_t, hs, hd, att_src, att_dst, bias, slope):
    V = rowptr.numel() - 1
    return (hs.new_empty((V, hs.size(1))), hs.new_empty((col.numel(),)), hs.new_empty((hs.size(0),)),
            hs.new_empty((V,)))


def _gat_setup(ctx, inputs, output):
    rowptr, col, rowptr_t, col_t, map_t, hs, hd, att_src, att_dst, bias, slope = inputs
    out, alpha, a_src, a_dst = output
    ctx.slope = slope
    ctx.has = (hd is not None, att_dst is not None, bias is not None)
    ctx.save_for_backward(rowptr, col, rowptr_t, col_t, map_t, hs, hd, att_src, att_dst, alpha, a_src, a_dst)


def _gat_backward(ctx, dout, _dalpha, _das, _dad):
    rowptr, col, rowptr_t, col_t, map_t, hs, hd, att_src, att_dst, alpha, a_src, a_dst = ctx.saved_tensors
    has_hd, has_ad, has_bias = ctx.has
    dout = dout.contiguous()
    dhs, da_src, da_dst = gat_pool_bwd(rowptr, col, rowptr_t, col_t, map_t, hs, a_src, a_dst, alpha, att_src, dout,
                                       ctx.slope)
    datt_src = da_src @ hs
    dhd = datt_dst = dbias = None
    if has_hd:
        dhd = da_dst.unsqueeze(1) * att_dst.unsqueeze(0)
        datt_dst = da_dst @ hd
    if has_bias:
        dbias = colsum(dout)
    return None, None, None, None, None, dhs, dhd, datt_src, datt_dst, dbias, None


torch.library.register_autograd("ghscn::gat_pool", _gat_backward, setup_context=_gat_setup)


class GatPoolInputWidth(torch.autograd.Function):
    """GATConv forward with the pooled sum taken at the INPUT width:
        out_i = sum_s alpha_s (x_s W_src^T) + b = (sum_s alpha_s x_s) W_src^T + b,   a_src = x (W_src^T att_src)
    so the [N,F]x[F,H] projection of every source node disappears from the forward (N >> number of destination
    rows for the local -> virtual pool).  The backward -- never reached in HSCN training, where the virtual branch
    is dead w.r.t. the loss -- materialises hs = x W_src^T and reuses the K5 backward kernels."""

    @staticmethod
    def forward(ctx, x_src, x_dst, w_src, w_dst, att_src, att_dst, bias, slope, rowptr, col, rowptr_t, col_t, map_t):
        from . import gemm
        x_src = _rowmajor(x_src)
        V = rowptr.numel() - 1
        dev = x_src.device
        L, st = lib(), _stream()
        side = None
        if x_dst is not None:
            x_dst = _rowmajor(x_dst)
            if PARALLEL_AUX and x_src.is_cuda:      # the two score halves are independent: a_dst on an aux stream
                main = torch.cuda.current_stream()
                side = _aux_stream(dev)
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    a_dst = row_dot(x_dst, att_dst @ w_dst)
                a_dst.record_stream(main)
            else:
                a_dst = row_dot(x_dst, att_dst @ w_dst)
        else:
            a_dst = None
        a_src = row_dot(x_src, att_src @ w_src)
        if side is not None:
            main.wait_stream(side)
        alpha = torch.zeros(col.numel(), dtype=torch.float32, device=dev)
        L.call("ghscn_gat_scores", _p(rowptr), _p(col), _p(a_src), _p(a_dst), float(slope), V, _p(alpha), st)
        Fin = x_src.size(1)
        pooled = torch.empty((V, Fin), dtype=torch.float32, device=dev)
        L.call("ghscn_spmm_pool", _p(rowptr), _p(col), _p(alpha), _p(x_src), x_src.stride(0), _p(pooled), Fin, None,
               V, Fin, st)
        out = gemm.linear(pooled, w_src, bias)
        ctx.slope = slope
        ctx.save_for_backward(x_src, x_dst, w_src, w_dst, att_src, att_dst, alpha, a_src, a_dst, rowptr, col, rowptr_t,
                              col_t, map_t)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        (x_src, x_dst, w_src, w_dst, att_src, att_dst, alpha, a_src, a_dst, rowptr, col, rowptr_t, col_t,
         map_t) = ctx.saved_tensors
        dout = dout.contiguous()
        hs = x_src @ w_src.t()
        dhs, da_src, da_dst = gat_pool_bwd(rowptr, col, rowptr_t, col_t, map_t, hs, a_src, a_dst, alpha, att_src, dout,
                                           ctx.slope)
        dx_src = dhs @ w_src
        dw_src = dhs.t() @ x_src
        datt_src = da_src @ hs
        dx_dst = dw_dst = datt_dst = None
        if x_dst is not None:
            hd = x_dst @ w_dst.t()
            dhd = da_dst.unsqueeze(1) * att_dst.unsqueeze(0)
            dx_dst = dhd @ w_dst
            dw_dst = dhd.t() @ x_dst
            datt_dst = da_dst @ hd
        dbias = colsum(dout) if ctx.has_bias else None
        return (dx_src, dx_dst, dw_src, dw_dst, datt_src, datt_dst, dbias, None, None, None, None, None, None)


# =============================================================================================
# K6  fused MinCUT pool
# =============================================================================================
@_custom_op("ghscn::mincut_bwd", mutates_args=(), device_types="cuda")
def mincut_bwd(s_soft: Tensor, x: Optional[Tensor], ptr: Tensor, rowptr: Tensor, col: Tensor,
               adj_val: Optional[Tensor], rowptr_t: Tensor, col_t: Tensor, adj_val_t: Optional[Tensor],
               ss_raw: Tensor, adj_raw: Tensor, stats: Tensor, g_out: Optional[Tensor],
               g_out_adj: Optional[Tensor], g_losses: Tensor, temp: float, max_nodes: int, num_feat: int,
               want_dx: bool) -> Tuple[Tensor, Tensor]:
    N, K = s_soft.shape
    B = ptr.numel() - 1
    dev = s_soft.device
    dz = torch.empty((N, K), dtype=torch.float32, device=dev)
    dx = torch.empty((N, num_feat), dtype=torch.float32, device=dev) if want_dx else torch.empty(0, device=dev)
    L = lib()
    if x is not None:
        x = _rowmajor(x)
    if g_out is not None:
        g_out = g_out.contiguous()
    ldx = x.stride(0) if x is not None else 0
    # dense-bound corner: the per-graph products as tiled GEMMs over all graphs at once
    split = (K >= MINCUT_SPLIT_MIN_K and s_soft.is_contiguous()
             and bool(L.query("ghscn_mincut_bwd_split_supported", K, num_feat, ldx, num_feat, max_nodes)))
    ws_bytes = L.query("ghscn_mincut_bwd_split_workspace_bytes" if split else "ghscn_mincut_workspace_bytes", N, B, K)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    L.call("ghscn_mincut_bwd_split" if split else "ghscn_mincut_bwd", _p(s_soft), _p(x), x.stride(0) if x is not None else 0, _p(ptr), _p(rowptr), _p(col),
           _p(adj_val), _p(rowptr_t), _p(col_t), _p(adj_val_t), float(temp), B, N, K, num_feat, max_nodes,
           _p(ss_raw), _p(adj_raw), _p(stats), _p(g_out), _p(g_out_adj), _p(g_losses), _p(dz), K,
           _p(dx) if want_dx else None, num_feat, _p(ws), ws_bytes, _stream())
    return dz, dx


@mincut_bwd.register_fake
def _(s_soft, x, ptr, rowptr, col, adj_val, rowptr_t, col_t, adj_val_t, ss_raw, adj_raw, stats, g_out, g_out_adj,
      g_losses, temp, max_nodes, num_feat, want_dx):
    return s_soft.new_empty(s_soft.shape), s_soft.new_empty((s_soft.size(0), num_feat) if want_dx else (0,))


@_custom_op("ghscn::mincut_pool", mutates_args=(), device_types="cuda")
def mincut_pool(logits: Tensor, x: Tensor, ptr: Tensor, rowptr: Tensor, col: Tensor, adj_val: Optional[Tensor],
                rowptr_t: Tensor, col_t: Tensor, adj_val_t: Optional[Tensor], temp: float, max_nodes: int,
                want_out: bool, want_adj: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """Ragged-batch MinCUT pool.  (rowptr, col) rows are edge_index[0] (rows of A); *_t the transpose.
    -> (out [B,K,H], out_adj [B,K,K], losses [2] = (mincut, ortho), s_soft [N,K], ss_raw, adj_raw, stats)"""
    logits, x = _rowmajor(logits), _rowmajor(x)
    N, K = logits.shape
    H = x.size(1)
    B = ptr.numel() - 1
    dev = logits.device
    f32 = dict(dtype=torch.float32, device=dev)
    s_soft = torch.empty((N, K), **f32)
    out = torch.empty((B, K, H), **f32) if want_out else torch.empty(0, **f32)
    out_adj = torch.empty((B, K, K), **f32) if want_adj else torch.empty(0, **f32)
    ss_raw = torch.empty((B, K, K), **f32)
    adj_raw = torch.empty((B, K, K), **f32)
    stats = torch.empty((B, _STATS_STRIDE), **f32)
    losses = torch.empty(2, **f32)
    L = lib()
    ws_bytes = L.query("ghscn_mincut_workspace_bytes", N, B, K)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    # The pooled features S^T X are dense-bound once K >= 64 (profiles/r1_sweeps.md: ~9 TFLOP/s on the SIMT pipes):
    # there the fused kernel skips them and the per-graph contraction runs on the tcgen05 tensor cores (3xTF32,
    # fp32-level accuracy) over the same `ptr` segments.  Below that the path is HBM/latency-bound and stays fused.
    from . import gemm
    big_k = K >= MINCUT_TC_MIN_K and N > 0
    tc_out = want_out and big_k and gemm.gemm3x_tn_segmented_supported(K, H, max_nodes, B)
    tc_kk = (big_k and MINCUT_TC_KK and gemm.gemm3x_tn_segmented_supported(K, K, max_nodes, B)
             and (tc_out or not want_out) and ws_bytes >= N * K * 4)
    args = (_p(logits), logits.stride(0), _p(x), x.stride(0), _p(ptr), _p(rowptr), _p(col), _p(adj_val), float(temp),
            B, N, K, H, max_nodes, _p(s_soft), _p(out) if (want_out and not tc_out) else None,
            _p(out_adj) if want_adj else None, _p(ss_raw), _p(adj_raw), _p(stats), _p(losses), _p(ws), ws_bytes,
            _stream())
    if tc_kk:                               # S, A S and the traces; then every contraction on the tensor cores
        L.call("ghscn_mincut_fwd_phase", *args, MINCUT_TC_PHASE1)
        a_s = ws[:N * K * 4].view(torch.float32).view(N, K)
        gemm.gemm3x_tn_segmented(s_soft, s_soft, ptr, max_nodes, out=ss_raw)
        gemm.gemm3x_tn_segmented(s_soft, a_s, ptr, max_nodes, out=adj_raw)
        if tc_out:
            gemm.gemm3x_tn_segmented(s_soft, x, ptr, max_nodes, out=out)
        L.call("ghscn_mincut_fwd_phase", *args, 2)
        return out, out_adj, losses, s_soft, ss_raw, adj_raw, stats
    L.call("ghscn_mincut_fwd", *args)
    if tc_out:
        gemm.gemm3x_tn_segmented(s_soft, x, ptr, max_nodes, out=out)
    return out, out_adj, losses, s_soft, ss_raw, adj_raw, stats


@mincut_pool.register_fake
def _(logits, x, ptr, rowptr, col, adj_val, rowptr_t, col_t, adj_val_t, temp, max_nodes, want_out, want_adj):
    N, K = logits.shape
    B, H = ptr.numel() - 1, x.size(1)
    e = logits.new_empty
    return (e((B, K, H) if want_out else (0,)), e((B, K, K) if want_adj else (0,)), e((2,)), e((N, K)),
            e((B, K, K)), e((B, K, K)), e((B, _STATS_STRIDE)))


def _mincut_setup(ctx, inputs, output):
    (logits, x, ptr, rowptr, col, adj_val, rowptr_t, col_t, adj_val_t, temp, max_nodes, want_out, want_adj) = inputs
    out, out_adj, losses, s_soft, ss_raw, adj_raw, stats = output
    ctx.temp, ctx.max_nodes, ctx.want = temp, max_nodes, (want_out, want_adj)
    ctx.num_feat = x.size(1)
    ctx.set_materialize_grads(False)       # unused outputs (s_soft, raw K x K blocks, stats) arrive as None, not as zero fills
    ctx.save_for_backward(s_soft, x, ptr, rowptr, col, adj_val, rowptr_t, col_t, adj_val_t, ss_raw, adj_raw, stats)


def _mincut_backward(ctx, g_out, g_adj, g_losses, *_unused):
    (s_soft, x, ptr, rowptr, col, adj_val, rowptr_t, col_t, adj_val_t, ss_raw, adj_raw, stats) = ctx.saved_tensors
    want_out, want_adj = ctx.want
    g_out = g_out.contiguous() if (want_out and g_out is not None) else None
    g_adj = g_adj.contiguous() if (want_adj and g_adj is not None) else None
    if g_losses is None:
        g_losses = torch.zeros(2, dtype=torch.float32, device=s_soft.device)
    want_dx = bool(ctx.needs_input_grad[1])
    dz, dx = mincut_bwd(s_soft, x, ptr, rowptr, col, adj_val, rowptr_t, col_t, adj_val_t, ss_raw, adj_raw, stats,
                        g_out, g_adj, g_losses.contiguous(), ctx.temp, ctx.max_nodes, ctx.num_feat, want_dx)
    return (dz, dx if want_dx else None) + (None,) * 11


torch.library.register_autograd("ghscn::mincut_pool", _mincut_backward, setup_context=_mincut_setup)


# =============================================================================================
# K7  cluster assignment -> virtual nodes (integer work, no autograd)
# =============================================================================================
@_custom_op("ghscn::cluster_argmax", mutates_args=(), device_types="cuda")
def cluster_argmax(s_soft: Tensor) -> Tensor:
    s_soft = _rowmajor(s_soft)
    out = torch.empty(s_soft.size(0), dtype=torch.int32, device=s_soft.device)
    lib().call("ghscn_cluster_argmax", _p(s_soft), s_soft.stride(0), s_soft.size(0), s_soft.size(1), _p(out),
               _stream())
    return out


@cluster_argmax.register_fake
def _(s_soft):
    return s_soft.new_empty((s_soft.size(0),), dtype=torch.int32)


@_custom_op("ghscn::virtual_build", mutates_args=(), device_types="cuda")
def virtual_build(cluster: Tensor, ptr: Tensor, x_raw: Tensor, num_clusters: int) -> Tuple[Tensor, Tensor, Tensor]:
    """-> (cluster_remapped int32 [N], num_virtual int32 [B], virt_x_padded fp32 [B*K, F])"""
    if x_raw.dtype not in (torch.int64, torch.float32):
        raise TypeError("raw node features must be int64 (OGB atoms) or float32")
    x_raw = x_raw.contiguous()
    N, F = x_raw.shape
    B = ptr.numel() - 1
    dev = x_raw.device
    remap = torch.empty(N, dtype=torch.int32, device=dev)
    nv = torch.empty(B, dtype=torch.int32, device=dev)
    vx = torch.empty((B * num_clusters, F), dtype=torch.float32, device=dev)
    lib().call("ghscn_virtual_build", _p(cluster), _p(ptr), _p(x_raw), int(x_raw.dtype == torch.int64), F, B,
               num_clusters, F, _p(remap), _p(nv), _p(vx), _stream())
    return remap, nv, vx


@virtual_build.register_fake
def _(cluster, ptr, x_raw, num_clusters):
    B = ptr.numel() - 1
    return (cluster.new_empty(cluster.shape), cluster.new_empty((B,)),
            x_raw.new_empty((B * num_clusters, x_raw.size(1)), dtype=torch.float32))


def cast_i64_f32(x: Tensor) -> Tensor:
    """`batch.x.float()` (train.py:79) for int64 OGB atom features."""
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    lib().call("ghscn_cast_i64_f32", _p(x), x.numel(), _p(out), _stream())
    return out


# =============================================================================================
# task loss (loss.py:6-19) over the first `rows` graphs, one launch forward, none backward
# =============================================================================================
LOSS_MODES = {"cross_entropy": 0, "l1": 1}


class GraphLoss(torch.autograd.Function):
    """criterion(loss_fn, pred, true) for [B, C] float targets: mean BCE-with-logits or mean L1 over the first `rows`
    rows (padding graphs behind them get zero gradient), the sigmoid score of train/train.py:82, and d loss / d pred,
    all from ONE kernel; the backward only scales the stored gradient by the incoming scalar."""

    @staticmethod
    def forward(ctx, pred: Tensor, target: Tensor, rows: int, mode: int):
        pred = _rowmajor(pred)
        target = _rowmajor(target.float())
        total, c = pred.shape
        dev = pred.device
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        d_pred = torch.empty((total, c), dtype=torch.float32, device=dev)
        score = torch.empty((total, c), dtype=torch.float32, device=dev)
        lib().call("ghscn_graph_loss", _p(pred), pred.stride(0), _p(target), target.stride(0), int(rows), total, c,
                   int(mode), _p(loss), _p(d_pred), _p(score), _stream())
        ctx.save_for_backward(d_pred)
        ctx.mark_non_differentiable(score)
        return loss.view(()), score

    @staticmethod
    def backward(ctx, g_loss, _g_score):
        (d_pred,) = ctx.saved_tensors
        return d_pred * g_loss, None, None, None


def graph_loss(loss_fn: str, pred: Tensor, target: Tensor, rows: Optional[int] = None):
    """-> (loss, sigmoid(pred[:rows])) like models.criterion / loss.py:6-19 (the [B, C] float-target forms)."""
    rows = pred.size(0) if rows is None else int(rows)
    loss, score = GraphLoss.apply(pred, target, rows, LOSS_MODES[loss_fn])
    return loss, score[:rows]


class HeadOutLoss(torch.autograd.Function):
    """lin_2 (model/hscn.py:112) + criterion (loss.py:6-19) + the backward of both in ONE launch (batches of <= 256
    graphs): the forward already produces d loss / d hidden, d loss / d W2 and d loss / d b2 for an incoming gradient
    of 1; the backward scales them by the incoming scalar (skipped with `unit_grad`, for callers that differentiate the
    loss itself -- `loss.backward()` -- and nothing derived from it)."""

    @staticmethod
    def forward(ctx, hidden: Tensor, weight: Tensor, bias: Optional[Tensor], target: Tensor, rows: int, mode: int,
                unit_grad: bool):
        hidden = _rowmajor(hidden)
        weight = _rowmajor(weight)
        target = _rowmajor(target.float())
        total, h = hidden.shape
        c = weight.size(0)
        dev = hidden.device
        out = torch.empty(1 + 2 * total * c, dtype=torch.float32, device=dev)
        loss, pred, score = out[:1], out[1:1 + total * c].view(total, c), out[1 + total * c:].view(total, c)
        grads = torch.empty(total * h + c * h + c, dtype=torch.float32, device=dev)
        d_h = grads[:total * h].view(total, h)
        d_w = grads[total * h:total * h + c * h].view(c, h)
        d_b = grads[total * h + c * h:]
        lib().call("ghscn_head_out_loss", _p(hidden), hidden.stride(0), _p(weight), weight.stride(0),
                   _p(bias) if bias is not None else None, _p(target), target.stride(0), int(rows), total, h, c,
                   int(mode), _p(pred), _p(loss), _p(score), _p(d_w), _p(d_b) if bias is not None else None, _p(d_h),
                   _stream())
        ctx.save_for_backward(grads)
        ctx.dims = (total, h, c, bias is not None, bool(unit_grad))
        ctx.mark_non_differentiable(pred, score)
        ctx.set_materialize_grads(False)        # no zero fills for the gradients of pred / score
        return loss.view(()), pred, score

    @staticmethod
    def backward(ctx, g_loss, _g_pred, _g_score):
        (grads,) = ctx.saved_tensors
        total, h, c, has_bias, unit = ctx.dims
        if g_loss is None:
            return None, None, None, None, None, None, None
        if not unit:
            grads = grads * g_loss
        d_h = grads[:total * h].view(total, h)
        d_w = grads[total * h:total * h + c * h].view(c, h)
        d_b = grads[total * h + c * h:] if has_bias else None
        return d_h, d_w, d_b, None, None, None, None


def head_out_loss_ok(total_rows: int, hidden: int, num_targets: int) -> bool:
    return bool(lib().query("ghscn_head_out_loss_supported", int(total_rows), int(hidden), int(num_targets)))


def head_out_loss(loss_fn: str, hidden: Tensor, weight: Tensor, bias: Optional[Tensor], target: Tensor,
                  rows: Optional[int] = None, unit_grad: bool = False):
    """-> (loss, pred, sigmoid(pred[:rows])): `criterion(loss_fn, lin_2(hidden), target)` of model/hscn.py:112 +
    loss.py:6-19 for [B, C] float targets, forward and backward in one kernel."""
    rows = hidden.size(0) if rows is None else int(rows)
    loss, pred, score = HeadOutLoss.apply(hidden, weight, bias, target, rows, LOSS_MODES[loss_fn], unit_grad)
    return loss, pred, score[:rows]


# =============================================================================================
# fused "virtual" destination of the HeteroConv: v->v GCN + l->v GAT pool, summed (model/hscn.py:83-96)
# =============================================================================================
# measured slower in the step (0.665 vs 0.653 ms): a tcgen05 CTA owns its SM outright for ~27 us, the fp32 tile kernel
# shares SMs with the local chain's kernels; kept selectable
VIRTUAL_TCGEN05 = os.environ.get("GHSCN_VIRTUAL_TCGEN05", "0") != "0"


# 0 = from the mean row length.  Timed alone the split wins (F = 300: 15.5 -> 12.8 us with 2 warps per row, F = 9: 13.4 ->
# 8.1 us with 4); inside the step, where the pool shares the SMs with the local chain's GEMMs, one warp per row is as
# fast or faster (0.641 vs 0.646-0.648 ms per step), so that is the default.
POOL_WARPS_PER_ROW = int(os.environ.get("GHSCN_POOL_WPR", "1"))


def pool_warps_per_row(nnz: int, num_rows: int) -> int:
    """Warps that share one destination row of the fused attention pool (a pooling relation has only ~num_rows warps
    of work: 9 per SM at the bench shape)."""
    if POOL_WARPS_PER_ROW:
        return POOL_WARPS_PER_ROW
    mean = nnz / max(num_rows, 1)
    return 1 if mean < 6 else (2 if mean < 12 else 4)


class VirtualLayerFused(torch.autograd.Function):
    """out = [relu]( GCNConv_vv(x_v) + GATConv_lv((x_l, x_v)) ) computed at the INPUT width:
        P = sum_s alpha_s x_l[s]            one-pass attention pool (ghscn_gat_pool_fused_fwd)
        Q = A_hat_vv x_v                    GCN aggregation before the projection (linearity)
        out = P W_src^T + b_gat + Q W_vv^T + b_vv          ONE small GEMM over the concatenated reduction
    4 launches instead of ~16 per layer.  The backward (never reached in HSCN training: nothing downstream of the
    virtual nodes feeds the loss, model/hscn.py:84-94) re-runs the two unfused, differentiable operators."""

    @staticmethod
    def forward(ctx, x_src, x_dst, w_src, w_dst, att_src, att_dst, b_gat, w_vv, b_vv, meta):
        x_src, x_dst = _rowmajor(x_src), _rowmajor(x_dst)
        V, F = x_dst.shape
        H = w_src.size(0)
        dev = x_src.device
        L, st = lib(), _stream()
        lvd, vvd, vv_w = meta["lv_by_dst"], meta["vv_by_dst"], meta["vv_w"]
        w_src_c, w_dst_c, w_vv_c = w_src.contiguous(), w_dst.contiguous(), w_vv.contiguous()
        u = torch.empty((2, F), dtype=torch.float32, device=dev)
        L.call("ghscn_gat_fold_attention", _p(w_src_c), F, _p(att_src.contiguous()), _p(w_dst_c), F,
               _p(att_dst.contiguous()), H, F, F, 1, _p(u[0]), _p(u[1]), st)
        from . import gemm
        # wide layers: the projection of [Q | P] by [W_vv | W_src] runs on the tcgen05 tensor cores, one CTA per
        # (128-row tile, N half) -- ~22 CTAs for 1.3 k virtual nodes, which leaves the other SMs to the local chain
        # that runs concurrently (the fp32 tile kernel spreads 410 CTAs over every SM and both chains stall)
        tc = (VIRTUAL_TCGEN05 and gemm.USE_TCGEN05 and F % 4 == 0 and F >= 64
              and gemm.gemm3x_supported(V, H, 2 * F))
        ld = 2 * F if tc else F
        both = torch.empty((V, 2 * F), dtype=torch.float32, device=dev) if tc else None
        agg = both[:, :F] if tc else torch.empty((V, F), dtype=torch.float32, device=dev)
        pooled = both[:, F:] if tc else torch.empty((V, F), dtype=torch.float32, device=dev)
        L.call("ghscn_gat_pool_fused_fwd", _p(lvd.rowptr), _p(lvd.col), _p(x_src), x_src.stride(0), _p(x_dst),
               x_dst.stride(0), _p(u[0]), _p(u[1]), float(meta["slope"]), V, F, 1, _p(pooled), ld,
               pool_warps_per_row(lvd.col.numel(), V), st)
        L.call("ghscn_spmm", _p(vvd.rowptr), _p(vvd.col), _p(vv_w), _p(x_dst), x_dst.stride(0), _p(agg), ld, None, V,
               F, 0, st)
        if tc:
            wcat = torch.cat([w_vv_c, w_src_c], dim=1)
            bias = b_vv if b_gat is None else (b_gat if b_vv is None else b_vv + b_gat)
            out = gemm.gemm3x(both, gemm.gemm3x_prep(wcat), H, bias, relu=bool(meta["relu"]))
        else:
            out = torch.empty((V, H), dtype=torch.float32, device=dev)
            L.call("ghscn_small_linear2_fwd", _p(agg), F, _p(w_vv_c), F, _p(b_vv), F, _p(pooled), F, _p(w_src_c), F,
                   _p(b_gat), F, 2 if meta["relu"] else 0, V, H, _p(out), H, st)
        ctx.meta = meta
        ctx.save_for_backward(x_src, x_dst)
        return out

    @staticmethod
    def backward(ctx, dout):
        x_src, x_dst = ctx.saved_tensors
        meta = ctx.meta
        gat, gcn = meta["gat"], meta["gcn"]
        with torch.enable_grad():
            xs, xd = x_src.detach().requires_grad_(), x_dst.detach().requires_grad_()
            out = gcn(xd, meta["vv_index"]) + gat._forward_cuda(xs, xd, meta["lv_index"], None, None)
            if meta["relu"]:
                out = out.relu()
            params = [gat.lin_src.weight, gat.lin_dst.weight, gat.att_src, gat.att_dst, gat.bias, gcn.lin.weight,
                      gcn.bias]
            live = [p for p in params if p is not None]
            grads = list(torch.autograd.grad(out, [xs, xd] + live, dout.contiguous(), allow_unused=True))
        dxs, dxd = grads[0], grads[1]
        it = iter(grads[2:])
        pg = [next(it) if p is not None else None for p in params]
        return (dxs, dxd, pg[0], pg[1], pg[2], pg[3], pg[4], pg[5], pg[6], None)


# =============================================================================================
# ReLU + dropout in one pass (model/mpnn.py:57-58), no mask tensor
# =============================================================================================
_DROPOUT_STATE: dict = {}


def dropout_state(device: torch.device) -> Tensor:
    """{seed, call counter} of the fused dropout on `device` (int64 [2] on the device; seeded from torch.initial_seed()
    on first use -- `manual_seed` before that makes runs reproducible; captured graphs advance the counter themselves)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    st = _DROPOUT_STATE.get(idx)
    if st is None:
        st = _DROPOUT_STATE[idx] = torch.tensor([torch.initial_seed() & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64,
                                                device=device)
    return st


class ReluDropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, p: float):
        x = x.contiguous()
        y = torch.empty_like(x)
        lib().call("ghscn_relu_dropout_fwd", _p(x), x.numel(), float(p), _p(dropout_state(x.device)), _p(y), _stream())
        ctx.save_for_backward(y)
        ctx.p = float(p)
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        (y,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        lib().call("ghscn_relu_dropout_bwd", _p(dy), _p(y), dy.numel(), ctx.p, _p(dx), _stream())
        return dx, None


def relu_dropout(x: Tensor, p: float, training: bool = True) -> Tensor:
    """F.dropout(F.relu(x), p, training) (model/mpnn.py:57-58 with the ReLU activation); one kernel, no mask tensor."""
    if not training or p <= 0.0:
        return torch.relu(x)
    if not (x.is_cuda and x.dtype == torch.float32) or p >= 1.0:
        return torch.nn.functional.dropout(torch.relu(x), p=p, training=True)
    return ReluDropout.apply(x, float(p))


# =============================================================================================
# multi-head GATConv forward in one pool launch for all heads
# =============================================================================================
class GatMultiHead(torch.autograd.Function):
    """GATConv(heads = h) forward: the attention vectors of all heads are folded in one launch, ALL heads are pooled at
    the input width in one launch (grid.y = head), then one small projection per head writes its column block of the
    concatenated output.  The backward re-runs the head-by-head differentiable operators (`meta["unfused"]`)."""

    @staticmethod
    def forward(ctx, x_src, x_dst, w_src, w_dst, att_src, att_dst, meta):
        x_src = _rowmajor(x_src)
        heads, C = meta["heads"], meta["out_channels"]
        V, F = meta["num_dst"], x_src.size(1)
        dev = x_src.device
        L, st = lib(), _stream()
        d = meta["by_dst"]
        w_src_c = w_src.contiguous()
        has_dst = x_dst is not None
        if has_dst:
            x_dst = _rowmajor(x_dst)
            w_dst_c = w_dst.contiguous()
        u = torch.empty((2, heads, F), dtype=torch.float32, device=dev)
        L.call("ghscn_gat_fold_attention", _p(w_src_c), F, _p(att_src.contiguous()),
               _p(w_dst_c) if has_dst else None, F, _p(att_dst.contiguous()) if has_dst else None, C, F,
               F if has_dst else 0, heads, _p(u[0]), _p(u[1]) if has_dst else None, st)
        pooled = torch.empty((heads, V, F), dtype=torch.float32, device=dev)
        L.call("ghscn_gat_pool_fused_fwd", _p(d.rowptr), _p(d.col), _p(x_src), x_src.stride(0),
               _p(x_dst) if has_dst else None, x_dst.stride(0) if has_dst else 0, _p(u[0]),
               _p(u[1]) if has_dst else None, float(meta["slope"]), V, F, heads, _p(pooled), F,
               pool_warps_per_row(d.col.numel(), V), st)
        out = torch.empty((V, heads * C), dtype=torch.float32, device=dev)
        for h in range(heads):
            L.call("ghscn_small_linear_fwd", _p(pooled[h]), F, w_src_c.data_ptr() + 4 * h * C * F, F, None, 0, V, F, C,
                   out.data_ptr() + 4 * h * C, heads * C, st)
        ctx.meta = meta
        ctx.save_for_backward(x_src, x_dst)
        return out

    @staticmethod
    def backward(ctx, dout):
        x_src, x_dst = ctx.saved_tensors
        meta = ctx.meta
        with torch.enable_grad():
            xs = x_src.detach().requires_grad_()
            xd = xs if meta["shared_x"] else (x_dst.detach().requires_grad_() if x_dst is not None else None)
            out = meta["unfused"](xs, xd)
            params = list(meta["params"])               # the tensors this function received (w_dst None if shared)
            wanted = [xs] + ([] if (xd is None or xd is xs) else [xd]) + [p for p in params if p is not None]
            grads = list(torch.autograd.grad(out, wanted, dout.contiguous(), allow_unused=True))
        it = iter(grads)
        dxs = next(it)
        dxd = None if (xd is None or xd is xs) else next(it)
        pg = [next(it) if p is not None else None for p in params]
        # att_* enter as [1, heads, C]; the function received them in that shape
        return dxs, dxd, pg[0], pg[1], pg[2], pg[3], None
