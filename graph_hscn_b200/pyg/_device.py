"""Device staging for the drop-in route (`pyg.install()`): CPU tensors in, compute on the B200, results back.

The reference calls some PyG operators on CPU tensors although a GPU is present:
    train/train_clustering.py:37-43,58-64   gcn_norm(data.edge_index, ...) BEFORE data.to(device)
    train/train.py:78-81, main.py:117       the MPNN path never moves the model or the batch
With auto-device staging on (what `pyg.install()` selects) every operator / layer of this package accepts CPU
tensors: they are copied to the current CUDA device, the sm_100a kernels run there, and the result is returned
on the device of the operator's primary input, exactly where PyG would have left it.  Compute never runs on the
host: this is not a CPU fallback (without a CUDA device the copy itself raises).  With staging off (the
default for the mirror models / `GraphHSCNStep`) a CPU tensor raises as before.

Index tensors (edge_index, batch) are memoised by identity so the five layers of an MPNN forward share one
device copy -- and therefore one CSR in the structure cache -- like CUDA-resident batches do.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional, Tuple

import torch
from torch import Tensor

_AUTO = False
_MEMO: "OrderedDict[tuple, Tuple[Tensor, Tensor]]" = OrderedDict()
_MEMO_CAP = 32


def set_auto_device(on: bool) -> None:
    global _AUTO
    _AUTO = bool(on)
    if not on:
        _MEMO.clear()


def auto_device() -> bool:
    return _AUTO


def _compute_device(*tensors: Optional[Tensor]) -> torch.device:
    for t in tensors:
        if isinstance(t, Tensor) and t.is_cuda:
            return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("graph_hscn_b200 computes on a CUDA device (sm_100a) only and none is available; "
                           "there is no CPU fallback in the product path.")
    return torch.device("cuda", torch.cuda.current_device())


def plan(primary: Optional[Tensor], *others: Optional[Tensor]) -> Tuple[Optional[torch.device], Optional[torch.device]]:
    """-> (dev, back).  dev: CUDA device to stage CPU inputs to (None: everything is on CUDA already, nothing to do);
    back: device the results go back to (None: stay on the compute device)."""
    ts = [t for t in (primary,) + others if isinstance(t, Tensor)]
    if all(t.is_cuda for t in ts):
        return None, None
    if not _AUTO:
        raise RuntimeError("graph_hscn_b200 kernels are CUDA-only (sm_100a); got a CPU tensor. There is no CPU "
                           "fallback in the product path (pyg.install() / pyg.set_auto_device(True) stages CPU "
                           "tensors to the GPU for the reference's CPU-side call sites).")
    dev = _compute_device(*ts)
    back = primary.device if isinstance(primary, Tensor) and not primary.is_cuda else None
    return dev, back


def to_dev(t, dev: Optional[torch.device]):
    """Differentiable copy (plain `.to`) of a value tensor / parameter; no-op when dev is None or t is there."""
    if dev is None or not isinstance(t, Tensor) or t.device == dev:
        return t
    return t.to(dev)


def index_to_dev(t: Optional[Tensor], dev: Optional[torch.device]) -> Optional[Tensor]:
    """Memoised copy of an index tensor (no gradient): the same CPU tensor maps to the same device tensor until it
    is modified in place, so structures cached by device-tensor identity are reused across layers."""
    if dev is None or t is None or t.device == dev:
        return t
    key = (t.data_ptr(), t._version, tuple(t.shape), t.dtype, dev.index)
    hit = _MEMO.get(key)
    if hit is not None and hit[0] is t:
        _MEMO.move_to_end(key)
        return hit[1]
    d = t.to(dev)
    _MEMO[key] = (t, d)           # holds the CPU tensor: its address cannot be recycled while the entry lives
    while len(_MEMO) > _MEMO_CAP:
        _MEMO.popitem(last=False)
    return d


def back_to(out, back: Optional[torch.device]):
    if back is None:
        return out
    if isinstance(out, Tensor):
        return out.to(back)
    if isinstance(out, (tuple, list)):
        return type(out)(back_to(o, back) for o in out)
    return out
