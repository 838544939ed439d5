"""Dense projections (x W^T + b) of the layers: fp32-accurate GEMMs on the tensor cores via 3xTF32.

The reference's `nn.Linear`s are host-side PyTorch code by the north_star's split (library GEMMs, not a
hand-written kernel), but at h=300 they dominate the step (profiles/r1b: 47 % of the kernel time as fp32 SIMT
sgemm).  Plain TF32 breaks the 1e-5 parity bar; the 3xTF32 split keeps fp32-level accuracy on the tensor pipe:

    x = x_hi + x_lo,  W = W_hi + W_lo        (hi: low 13 mantissa bits cleared, exact in TF32; lo = rest)
    x W^T ~= x_hi W_hi^T + x_hi W_lo^T + x_lo W_hi^T           (dropped lo.lo term ~ 2^-22 relative)

A hand-written kernel (ghscn_split_tf32_cat) writes the operands K-concatenated -- A_cat = [lo | hi | hi],
W_cat = [hi | lo | hi] -- so ONE library TF32 GEMM over the 3K-long reduction yields the three-term sum with a
single pass over the output; the two small cross terms lead the reduction so they are accumulated while the
tensor core's (not correctly rounded) accumulator is still small.  Weight gradients reduce over all N rows; tensor-core accumulation error grows
with the reduction length, so they are computed per chunk of `DW_CHUNK` rows (batched GEMM) and the chunk
results are added in fp32.  Mode "fp32" keeps plain cuBLAS fp32 (TF32 off).
"""
from __future__ import annotations

import os
from contextlib import contextmanager
from typing import Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

from ._lib import lib
from .structure import _p, _stream

_MODE = "3xtf32"
MIN_ROWS, MIN_DIM = 4096, 64      # below this the GEMM is latency/bandwidth-bound: plain fp32 is as fast
DW_CHUNK = 1024


def set_gemm_mode(mode: str) -> None:
    global _MODE
    if mode not in ("fp32", "3xtf32"):
        raise ValueError(mode)
    _MODE = mode


def gemm_mode() -> str:
    return _MODE


@contextmanager
def _tf32():
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        yield
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def split_tf32(x: Tensor) -> Tuple[Tensor, Tensor]:
    x = x.contiguous()
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    lib().call("ghscn_split_tf32", _p(x), x.numel(), _p(hi), _p(lo), _stream())
    return hi, lo


def split_cat(x: Tensor, mode: int, pad_rows_to: int = 1) -> Tensor:
    """[rows_padded, 3K]: mode 0 -> [lo | hi | hi], mode 1 -> [hi | lo | hi]; padding rows are zero."""
    if x.stride(1) != 1:
        x = x.contiguous()
    n, k = x.shape
    n_pad = (n + pad_rows_to - 1) // pad_rows_to * pad_rows_to
    out = torch.empty((n_pad, 3 * k), dtype=torch.float32, device=x.device)
    lib().call("ghscn_split_tf32_cat", _p(x), x.stride(0), n, n_pad, k, mode, _p(out), _stream())
    return out


def gemm3x_supported(m: int, n_out: int, k: int) -> bool:
    return bool(lib().query("ghscn_gemm3x_supported", m, n_out, k))


def gemm3x_prep(weight: Tensor, transpose: bool = False) -> Tensor:
    """Weight image of the fused tcgen05 3xTF32 GEMM (csrc/gemm3x.cu): B[n, k] = weight[n, k], or weight[k, n]
    with transpose=True (the operand of dX = dY . W).  Split into TF32 hi/lo, K-chunked, 128B-swizzled."""
    if weight.stride(1) != 1:
        weight = weight.contiguous()
    n_out, k = (weight.size(1), weight.size(0)) if transpose else (weight.size(0), weight.size(1))
    nbytes = lib().query("ghscn_gemm3x_b_image_bytes", n_out, k)
    if nbytes == 0:
        raise ValueError(f"gemm3x does not support n_out={n_out}, k={k}")
    image = torch.empty(nbytes, dtype=torch.uint8, device=weight.device)
    lib().call("ghscn_gemm3x_prep_b", _p(weight), weight.stride(0), n_out, k, int(transpose), _p(image), _stream())
    return image


def gemm3x(a: Tensor, image: Tensor, n_out: int, bias: Optional[Tensor] = None, relu: bool = False,
           out: Optional[Tensor] = None) -> Tensor:
    """C = A . B^T (+ bias) (ReLU) on the tcgen05 tensor cores with fp32-level accuracy (3xTF32, split in-kernel)."""
    if a.stride(1) != 1 or a.stride(0) % 4 != 0:
        a = a.contiguous()
    m, k = a.shape
    c = torch.empty((m, n_out), dtype=torch.float32, device=a.device) if out is None else out
    lib().call("ghscn_gemm3x", _p(a), a.stride(0), m, k, _p(image), n_out, _p(bias), int(relu), _p(c), n_out,
               _stream())
    return c


# ---- weight images computed ahead of their first use -----------------------------------------------------------
# The image of a weight (TF32 hi/lo split, K-chunked, swizzled) depends on the parameters only, not on the batch: a
# training step can build all of them on a side stream at its very beginning instead of ~4 us in front of every GEMM
# on the critical chain.  `_Linear3xTF32` notes which weights take the tcgen05 path; `prefetch_images` serves them.
_IMAGES: dict = {}            # (data_ptr, shape, transpose) -> (image, event)
_FUSED_WEIGHTS: dict = {}     # data_ptr -> weight tensor seen on the tcgen05 path (forward and/or transposed use)


def clear_images() -> None:
    _IMAGES.clear()


def prefetch_images(stream: "torch.cuda.Stream") -> int:
    """Builds forward and transposed images of every weight known to take the tcgen05 path, on `stream`."""
    if not _FUSED_WEIGHTS:
        return 0
    n = 0
    with torch.cuda.stream(stream):
        for w, uses in list(_FUSED_WEIGHTS.values()):
            for transpose in sorted(uses):
                img = gemm3x_prep(w, transpose=transpose)
                ev = torch.cuda.Event()
                ev.record(stream)
                _IMAGES[(w.data_ptr(), tuple(w.shape), transpose)] = (img, ev)
                n += 1
    return n


def _image(weight: Tensor, transpose: bool) -> Tensor:
    key = (weight.data_ptr(), tuple(weight.shape), transpose)
    ent = _FUSED_WEIGHTS.get(weight.data_ptr())
    if ent is None:
        _FUSED_WEIGHTS[weight.data_ptr()] = (weight.detach(), {transpose})
    else:
        ent[1].add(transpose)
    hit = _IMAGES.get(key)
    if hit is not None:
        cur = torch.cuda.current_stream()
        cur.wait_event(hit[1])
        hit[0].record_stream(cur)
        return hit[0]
    return gemm3x_prep(weight, transpose=transpose)


TILE_ROWS, NUM_SMS, MAX_TAIL_TILES = 128, 148, 24
# round 1 sent the tiles beyond whole waves to a library fp32 GEMM; the kernel now runs one CTA per (tile, N half) when
# the tiles do not fill one wave exactly (csrc/gemm3x.cu), which keeps every row on the tensor cores
TAIL_ON_LIBRARY = os.environ.get("GHSCN_GEMM_TAIL", "tcgen05") == "library"


def wave_rows(n: int) -> int:
    """Rows the tcgen05 kernel should take so that its 128-row tiles fill whole waves of the 148 SMs.  One CTA per
    SM and ~30 us per CTA regardless of its work: a last wave of a few tiles (the typical Peptides batch has
    19.3 k nodes = 151 tiles) would double the kernel time, so up to MAX_TAIL_TILES trailing tiles go to a
    plain fp32 GEMM instead (~9 us for ~1 k rows)."""
    tiles = (n + TILE_ROWS - 1) // TILE_ROWS
    waves, tail = divmod(tiles, NUM_SMS)
    if 1 <= waves <= 2 and 0 < tail <= MAX_TAIL_TILES:
        return waves * NUM_SMS * TILE_ROWS
    return n


def linear_rows(a: Tensor, weight_nk: Tensor, image: Tensor, n_out: int, bias: Optional[Tensor], transposed: bool):
    """a . B^T with B[n, k] = weight (or weight^T when `transposed`): whole waves on tcgen05, tail rows on cuBLAS fp32."""
    m = a.size(0)
    r = wave_rows(m) if TAIL_ON_LIBRARY else m
    if r == m:
        return gemm3x(a, image, n_out, bias)
    y = torch.empty((m, n_out), dtype=torch.float32, device=a.device)
    gemm3x(a[:r], image, n_out, bias, out=y[:r])
    b_kn = weight_nk if transposed else weight_nk.t()          # [k, n_out]
    if bias is not None:
        torch.addmm(bias, a[r:], b_kn, out=y[r:])
    else:
        torch.mm(a[r:], b_kn, out=y[r:])
    return y


def gemm3x_tn_supported(rows: int, m_out: int, n_out: int) -> bool:
    return bool(lib().query("ghscn_gemm3x_tn_supported", rows, m_out, n_out))


def gemm3x_tn(p: Tensor, q: Tensor) -> Tensor:
    """P^T . Q for row-major P [rows, m], Q [rows, n] (the weight gradient dY^T x) on tcgen05, 3xTF32, fp32 accuracy."""
    if p.stride(1) != 1 or p.stride(0) % 4 != 0:
        p = p.contiguous()
    if q.stride(1) != 1 or q.stride(0) % 4 != 0:
        q = q.contiguous()
    rows, m = p.shape
    n = q.size(1)
    out = torch.empty((m, n), dtype=torch.float32, device=p.device)
    ws_bytes = lib().query("ghscn_gemm3x_tn_workspace_bytes", rows, m, n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=p.device)
    lib().call("ghscn_gemm3x_tn", _p(p), p.stride(0), _p(q), q.stride(0), rows, m, n, _p(out), _p(ws), ws_bytes,
               _stream())
    return out


SEG_MAX_ROWS, SEG_COL_BLOCK = 1024, 256


def gemm3x_tn_segmented_supported(m_out: int, n_out: int, max_rows: int, num_segments: int) -> bool:
    return (USE_TCGEN05 and m_out % 4 == 0 and n_out % 4 == 0 and m_out >= 4 and n_out >= 16 and m_out <= 4096
            and 0 < max_rows <= SEG_MAX_ROWS and 0 < num_segments <= 65535)


def gemm3x_tn_segmented(p: Tensor, q: Tensor, ptr: Tensor, max_rows: int, out: Optional[Tensor] = None) -> Tensor:
    """out[g] = p[rows_g]^T q[rows_g] for the row segments of `ptr` (int32 [B+1]) on the tcgen05 tensor cores with
    fp32-level accuracy; q wider than 320 columns is processed in column blocks of 256."""
    if p.stride(1) != 1 or p.stride(0) % 4 != 0:
        p = p.contiguous()
    if q.stride(1) != 1 or q.stride(0) % 4 != 0:
        q = q.contiguous()
    B, m, n = ptr.numel() - 1, p.size(1), q.size(1)
    if out is None:
        out = torch.empty((B, m, n), dtype=torch.float32, device=p.device)
    L, st = lib(), _stream()
    blocks = [(0, n)] if n <= 320 else [(c, min(SEG_COL_BLOCK, n - c)) for c in range(0, n, SEG_COL_BLOCK)]
    if blocks[-1][1] < 16 and len(blocks) > 1:          # keep the last block >= 16 columns wide
        (c0, w0), (c1, w1) = blocks[-2], blocks[-1]
        blocks[-2:] = [(c0, w0 - 16), (c0 + w0 - 16, w1 + 16)]
    for c, wdt in blocks:
        L.call("ghscn_gemm3x_tn_segmented", _p(p), p.stride(0), q.data_ptr() + 4 * c, q.stride(0), _p(ptr), B,
               int(max_rows), m, wdt, out.data_ptr() + 4 * c, n, m * n, st)
    return out


USE_TCGEN05 = os.environ.get("GHSCN_TCGEN05", "1") != "0"      # fused tcgen05 kernel (csrc/gemm3x.cu) where the shape is supported; else split_cat + library GEMM


class _Linear3xTF32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Optional[Tensor]):
        n, k = x.shape
        m = weight.size(0)
        fused = USE_TCGEN05 and gemm3x_supported(n, m, k)
        if fused:
            y = linear_rows(x, weight, _image(weight, False), m, bias, transposed=False)
            ctx.save_for_backward(x, weight)
        else:
            a_cat = split_cat(x, 0, DW_CHUNK)                    # [Npad, 3K] = [xl | xh | xh]
            w_cat = split_cat(weight, 1)                         # [out, 3K]  = [wh | wl | wh]
            with _tf32():
                if bias is not None:
                    y = torch.addmm(bias, a_cat[:n], w_cat.t())
                else:
                    y = torch.mm(a_cat[:n], w_cat.t())
            ctx.save_for_backward(a_cat, weight)
        ctx.fused = fused
        ctx.n, ctx.k, ctx.has_bias = n, k, bias is not None
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        saved, weight = ctx.saved_tensors
        n, k = ctx.n, ctx.k
        m = weight.size(0)
        dx = dw = db = None
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dx_fused = need_dx and USE_TCGEN05 and gemm3x_supported(n, k, m)
        if dx_fused:
            dx = linear_rows(dy, weight, _image(weight, True), k, None, transposed=True)   # dX = dY . W
        dw_fused = need_dw and ctx.fused and USE_TCGEN05 and gemm3x_tn_supported(n, m, k)
        if dw_fused:
            dw = gemm3x_tn(dy, saved)                            # dW = dY^T x, slab partials added in fp32
        if (need_dw and not dw_fused) or (need_dx and not dx_fused):
            d_cat = split_cat(dy, 0, DW_CHUNK)                   # [Npad, 3m] = [dl | dh | dh]
        if need_dx and not dx_fused:
            wh, wl = split_tf32(weight)
            w_rows = torch.cat([wh, wl, wh], dim=0)              # [3m, K]: dX = [dl|dh|dh] . [wh; wl; wh]
            with _tf32():
                dx = torch.mm(d_cat[:n], w_rows)
        if need_dw and not dw_fused:
            a_cat = split_cat(saved, 0, DW_CHUNK) if ctx.fused else saved
            c = a_cat.size(0) // DW_CHUNK
            a3 = a_cat.view(c, DW_CHUNK, 3 * k)
            d3 = d_cat.view(c, DW_CHUNK, 3 * m)
            dl, dh = d3[:, :, :m].transpose(1, 2), d3[:, :, 2 * m:].transpose(1, 2)
            with _tf32():
                # dh^T [xl | xh] in ONE batched GEMM (a_cat = [xl | xh | xh]), then the dl^T xh cross term
                both = torch.bmm(dh, a3[:, :, :2 * k])               # [c, m, 2k] = [dh^T xl | dh^T xh]
                part = torch.bmm(dl, a3[:, :, 2 * k:])               # small cross term first
            part = part + both[:, :, :k]
            part = part + both[:, :, k:]
            dw = part.sum(0)                                     # fp32 adds across chunks
        if ctx.has_bias and ctx.needs_input_grad[2]:
            from . import ops
            db = ops.colsum(dy.contiguous())
        return dx, dw, db


def skinny_dx(dy: Tensor, weight: Tensor) -> Tensor:
    """dx[n,k] = sum_m dy[n,m] W[m,k] for a tiny k (hand-written streaming kernel)."""
    n, m = dy.shape
    k = weight.size(1)
    dx = torch.empty((n, k), dtype=torch.float32, device=dy.device)
    lib().call("ghscn_skinny_linear_dx", _p(dy), m, _p(weight), n, k, m, _p(dx), k, _stream())
    return dx


def skinny_dw(dy: Tensor, x: Tensor) -> Tensor:
    """dW[m,k] = sum_n dy[n,m] x[n,k] (two-stage fixed-order reduction: deterministic)."""
    n, m = dy.shape
    k = x.size(1)
    L = lib()
    dw = torch.empty((m, k), dtype=torch.float32, device=dy.device)
    ws_bytes = L.query("ghscn_skinny_dw_workspace_bytes", n, k, m)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dy.device)
    L.call("ghscn_skinny_linear_dw", _p(dy), m, _p(x), x.stride(0), n, k, m, _p(dw), _p(ws), ws_bytes, _stream())
    return dw


class _LinearSkinny(torch.autograd.Function):
    """y = x W^T + b for a tiny input width (<= 32) and many rows: hand-written streaming kernels (csrc/skinny.cu)."""

    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Optional[Tensor]):
        if x.stride(1) != 1:
            x = x.contiguous()
        weight = weight.contiguous()
        n, k = x.shape
        m = weight.size(0)
        y = torch.empty((n, m), dtype=torch.float32, device=x.device)
        lib().call("ghscn_skinny_linear_fwd", _p(x), x.stride(0), _p(weight), _p(bias), n, k, m, _p(y), m, _stream())
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        n, k = x.shape
        m = weight.size(0)
        L, st = lib(), _stream()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((n, k), dtype=torch.float32, device=x.device)
            L.call("ghscn_skinny_linear_dx", _p(dy), m, _p(weight), n, k, m, _p(dx), k, st)
        if ctx.needs_input_grad[1]:
            dw = torch.empty((m, k), dtype=torch.float32, device=x.device)
            ws_bytes = L.query("ghscn_skinny_dw_workspace_bytes", n, k, m)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
            L.call("ghscn_skinny_linear_dw", _p(dy), m, _p(x), x.stride(0), n, k, m, _p(dw), _p(ws), ws_bytes, st)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            from . import ops
            db = ops.colsum(dy)
        return dx, dw, db


SMALL_ACTS = {"identity": 0, None: 0, "elu": 1, "relu": 2, "tanh": 3}
SMALL_LINEAR = os.environ.get("GHSCN_SMALL_LINEAR", "1") != "0"


class _LinearSmall(torch.autograd.Function):
    """y = act(x W^T + b) for problems of 10^2..10^3 rows (graph-level head, virtual-node projections): one fp32 FMA
    tile kernel per product (csrc/dense_small.cu) instead of a library GEMM + bias epilogue + activation kernel each;
    the backward folds act'(y) into the operand loads of dW / db / dx."""

    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Optional[Tensor], act: int):
        if x.stride(1) != 1:
            x = x.contiguous()
        if weight.stride(1) != 1:
            weight = weight.contiguous()
        n, k = x.shape
        m = weight.size(0)
        y = torch.empty((n, m), dtype=torch.float32, device=x.device)
        lib().call("ghscn_small_linear_fwd", _p(x), x.stride(0), _p(weight), weight.stride(0), _p(bias), int(act),
                   n, k, m, _p(y), m, _stream())
        ctx.save_for_backward(x, weight, y if act else None)
        ctx.act, ctx.has_bias = int(act), bias is not None
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        x, weight, y = ctx.saved_tensors
        if dy.stride(1) != 1:
            dy = dy.contiguous()
        n, k = x.shape
        m = weight.size(0)
        L, st = lib(), _stream()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((n, k), dtype=torch.float32, device=x.device)
            L.call("ghscn_small_linear_dx", _p(dy), dy.stride(0), _p(y), m, ctx.act, _p(weight), weight.stride(0),
                   n, k, m, _p(dx), k, st)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dw = torch.empty((m, k), dtype=torch.float32, device=x.device)
            db = torch.empty(m, dtype=torch.float32, device=x.device) if ctx.has_bias else None
            ws_bytes = L.query("ghscn_small_linear_dw_workspace_bytes", n, k, m)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device) if ws_bytes else None
            L.call("ghscn_small_linear_dw", _p(dy), dy.stride(0), _p(y), m, ctx.act, _p(x), x.stride(0), n, k, m,
                   _p(dw), _p(db), _p(ws), ws_bytes, st)
        return dx, dw, db, None


SMALL_MAX_ROWS = 4096


def small_linear_ok(x: Tensor, weight: Tensor) -> bool:
    return (SMALL_LINEAR and x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and weight.is_cuda
            and 0 < x.size(0) < SMALL_MAX_ROWS and weight.dim() == 2)


def linear_act(x: Tensor, weight: Tensor, bias: Optional[Tensor], act: Optional[str]) -> Optional[Tensor]:
    """act(x W^T + b) in one launch when the problem is small enough for the tile kernels; None otherwise (the caller
    then applies the activation itself after `linear`)."""
    if act not in SMALL_ACTS or not small_linear_ok(x, weight):
        return None
    return _LinearSmall.apply(x, weight, bias, SMALL_ACTS[act])


SKINNY_MAX_IN, SKINNY_MIN_ROWS = 32, 2048


def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None) -> Tensor:
    """F.linear with fp32-level accuracy: tall-skinny problems (raw-feature layers) stream through hand-written
    kernels, large square ones run as 3xTF32 on the tensor cores, problems of < 4 k rows (graph-level head, virtual
    nodes) through the fp32 tile kernels, everything else is plain cuBLAS fp32."""
    if (x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and x.size(0) >= SKINNY_MIN_ROWS
            and weight.size(1) <= SKINNY_MAX_IN and weight.numel() * 4 <= 96 * 1024):
        return _LinearSkinny.apply(x, weight, bias)
    if (_MODE == "3xtf32" and x.is_cuda and x.dim() == 2 and x.dtype == torch.float32
            and x.size(0) >= MIN_ROWS and min(weight.shape) >= MIN_DIM):
        return _Linear3xTF32.apply(x, weight, bias)
    if small_linear_ok(x, weight):
        return _LinearSmall.apply(x, weight, bias, 0)
    return F.linear(x, weight, bias)
