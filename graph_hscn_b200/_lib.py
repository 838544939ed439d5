"""ctypes binding of libghscn.so -- the C-ABI boundary declared in include/ghscn.h.

There is NO fallback: if the shared library is missing or a symbol cannot be
resolved the import of the product path fails loudly (the CPU oracle under
oracle/ is test infrastructure and is never consulted here).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int32, c_int64, c_size_t, c_void_p
from pathlib import Path
from typing import Dict, List, Tuple

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libghscn.so"

P = c_void_p      # device pointer / stream
I64 = c_int64
I32 = c_int32
F32 = c_float
F64 = c_double
SZ = c_size_t

# name -> (restype, argtypes); must list every GHSCN_API symbol of include/ghscn.h
SIGNATURES: Dict[str, Tuple[object, List[object]]] = {
    "ghscn_abi_version": (I32, []),
    "ghscn_error_string": (c_char_p, [I32]),
    "ghscn_launch_count": (ctypes.c_uint64, []),
    "ghscn_csr_workspace_bytes": (SZ, [I64, I64, I32]),
    "ghscn_csr_build": (I32, [P, P, I64, I64, I32, P, P, P, P, SZ, P]),
    "ghscn_csr_add_loops": (I32, [P, P, P, I64, I64, P, P, P, P]),
    "ghscn_csr_blocked_smem_bytes": (SZ, [I64, I64]),
    "ghscn_csr_build_blocked": (I32, [P, P, I64, P, I64, I64, I64, I64, P, P, P, P, P, P, P, P]),
    "ghscn_split_tf32": (I32, [P, I64, P, P, P]),
    "ghscn_split_tf32_cat": (I32, [P, I64, I64, I64, I64, I32, P, P]),
    "ghscn_gemm3x_supported": (I32, [I64, I64, I64]),
    "ghscn_gemm3x_b_image_bytes": (SZ, [I64, I64]),
    "ghscn_gemm3x_prep_b": (I32, [P, I64, I64, I64, I32, P, P]),
    "ghscn_gemm3x_set_trace": (I32, [P]),
    "ghscn_gemm3x": (I32, [P, I64, I64, I64, P, I64, P, I32, P, I64, P]),
    "ghscn_gemm3x_tn_supported": (I32, [I64, I64, I64]),
    "ghscn_gemm3x_tn_segmented": (I32, [P, I64, P, I64, P, I64, I64, I64, I64, P, I64, I64, P]),
    "ghscn_gemm3x_tn_workspace_bytes": (SZ, [I64, I64, I64]),
    "ghscn_gemm3x_tn": (I32, [P, I64, P, I64, I64, I64, I64, P, P, SZ, P]),
    "ghscn_skinny_linear_fwd": (I32, [P, I64, P, P, I64, I64, I64, P, I64, P]),
    "ghscn_skinny_dw_workspace_bytes": (SZ, [I64, I64, I64]),
    "ghscn_skinny_linear_dw": (I32, [P, I64, P, I64, I64, I64, I64, P, P, SZ, P]),
    "ghscn_skinny_linear_dx": (I32, [P, I64, P, I64, I64, I64, P, I64, P]),
    "ghscn_adamw_step": (I32, [P, P, P, P, I64, F64, F64, F64, F64, F64, P, P]),
    "ghscn_adamw_step_scaled": (I32, [P, P, P, P, I64, F64, F64, F64, F64, F64, P, P, P]),
    "ghscn_gather_flat": (I32, [P, P, I32, P, P]),
    "ghscn_grad_clip_workspace_bytes": (SZ, [I64]),
    "ghscn_grad_clip_scale": (I32, [P, I64, F32, P, SZ, P, P]),
    "ghscn_small_linear_fwd": (I32, [P, I64, P, I64, P, I32, I64, I64, I64, P, I64, P]),
    "ghscn_small_linear2_fwd": (I32, [P, I64, P, I64, P, I64, P, I64, P, I64, P, I64, I32, I64, I64, P, I64, P]),
    "ghscn_small_linear_dx": (I32, [P, I64, P, I64, I32, P, I64, I64, I64, I64, P, I64, P]),
    "ghscn_small_linear_dw_workspace_bytes": (SZ, [I64, I64, I64]),
    "ghscn_small_linear_dw": (I32, [P, I64, P, I64, I32, P, I64, I64, I64, I64, P, P, P, SZ, P]),
    "ghscn_head_out_loss_supported": (I32, [I64, I64, I64]),
    "ghscn_head_out_loss": (I32, [P, I64, P, I64, P, P, I64, I64, I64, I64, I64, I32, P, P, P, P, P, P, P]),
    "ghscn_graph_loss": (I32, [P, I64, P, I64, I64, I64, I64, I32, P, P, P, P]),
    "ghscn_relu_dropout_fwd": (I32, [P, I64, F32, P, P, P]),
    "ghscn_relu_dropout_bwd": (I32, [P, P, I64, F32, P, P]),
    "ghscn_colsum_workspace_bytes": (SZ, [I64, I64]),
    "ghscn_colsum": (I32, [P, I64, I64, I64, P, P, SZ, P]),
    "ghscn_virtual_csr": (I32, [P, P, P, P, P, I64, I64, I32] + [P] * 12 + [P]),
    "ghscn_collate_batch": (I32, [P, P, I64, P, I64, P, P, I64, P]),
    "ghscn_batch_to_ptr": (I32, [P, I64, I64, P, P]),
    "ghscn_gcn_deg_inv_sqrt": (I32, [P, P, P, P, I64, I64, P, P]),
    "ghscn_edge_weights": (I32, [P, P, P, P, P, P, I64, I64, I32, I32, P, P]),
    "ghscn_loop_weights": (I32, [P, P, P, I64, I64, F32, P, P, P]),
    "ghscn_laplacian_eig_workspace_bytes": (SZ, [I64, I64, I32]),
    "ghscn_laplacian_eig": (I32, [P, P, P, I64, I64, I32, I32, I32, I32, I32, P, P, P, P, SZ, P]),
    "ghscn_spmm": (I32, [P, P, P, P, I64, P, I64, P, I64, I64, I32, P]),
    "ghscn_spmm_masked_supported": (I32, [I64, I64, I64, I64]),
    "ghscn_spmm_masked": (I32, [P, P, P, P, I64, P, I64, P, I64, I64, I64, P]),
    "ghscn_colsum_masked": (I32, [P, I64, P, I64, I64, I64, P, P, SZ, P]),
    "ghscn_relu_grad_colsum_partial": (I32, [P, I64, P, I64, I64, I64, P, I64, P, SZ, P]),
    "ghscn_colsum_finish": (I32, [P, SZ, I64, I64, P, P]),
    "ghscn_spmm_pool": (I32, [P, P, P, P, I64, P, I64, P, I64, I64, P]),
    "ghscn_gat_scores": (I32, [P, P, P, P, F32, I64, P, P]),
    "ghscn_spmm_edge_grad": (I32, [P, P, P, P, I64, P, I64, I64, I64, I64, P, P]),
    "ghscn_segment_reduce": (I32, [P, I64, P, P, I64, I64, I32, P, I64, P]),
    "ghscn_segment_broadcast": (I32, [P, I64, P, P, I64, I64, I32, P, I64, P]),
    "ghscn_row_dot": (I32, [P, I64, P, I64, I64, P, P]),
    "ghscn_gat_pool_fwd": (I32, [P, P, P, I64, P, P, P, F32, I64, I64, P, P, I64, P]),
    "ghscn_gat_pool_bwd_scores": (I32, [P, P, P, I64, P, P, P, P, I64, F32, I64, I64, P, P, P]),
    "ghscn_gat_pool_bwd_src": (I32, [P, P, P, P, P, P, I64, P, I64, I64, P, I64, P, P]),
    "ghscn_gat_fold_attention": (I32, [P, I64, P, P, I64, P, I64, I64, I64, I64, P, P, P]),
    "ghscn_gat_pool_fused_supported": (I32, [I64, I64, I64]),
    "ghscn_gat_pool_fused_fwd": (I32, [P, P, P, I64, P, I64, P, P, F32, I64, I64, I64, P, I64, I32, P]),
    "ghscn_slot_map": (I32, [P, P, I64, I64, P, P, P]),
    "ghscn_mincut_workspace_bytes": (SZ, [I64, I64, I64]),
    "ghscn_mincut_fwd": (I32, [P, I64, P, I64, P, P, P, P, F32, I64, I64, I64, I64, I32,
                               P, P, P, P, P, P, P, P, SZ, P]),
    "ghscn_mincut_fwd_phase": (I32, [P, I64, P, I64, P, P, P, P, F32, I64, I64, I64, I64, I32,
                                     P, P, P, P, P, P, P, P, SZ, P, I32]),
    "ghscn_mincut_bwd": (I32, [P, P, I64, P, P, P, P, P, P, P, F32, I64, I64, I64, I64, I32,
                               P, P, P, P, P, P, P, I64, P, I64, P, SZ, P]),
    "ghscn_mincut_bwd_split_workspace_bytes": (SZ, [I64, I64, I64]),
    "ghscn_mincut_bwd_split_supported": (I32, [I64, I64, I64, I64, I32]),
    "ghscn_mincut_bwd_split": (I32, [P, P, I64, P, P, P, P, P, P, P, F32, I64, I64, I64, I64, I32,
                                     P, P, P, P, P, P, P, I64, P, I64, P, SZ, P]),
    "ghscn_cluster_argmax": (I32, [P, I64, I64, I64, P, P]),
    "ghscn_virtual_build": (I32, [P, P, P, I32, I64, I64, I64, I64, P, P, P, P]),
    "ghscn_virtual_offsets": (I32, [P, I64, P, P, P]),
    "ghscn_virtual_edges": (I32, [P, P, P, P, I64, I64, I64, P, P, I64, P, P]),
    "ghscn_virtual_compact": (I32, [P, P, P, I64, I64, I64, P, P, P]),
    "ghscn_cast_i64_f32": (I32, [P, I64, P, P]),
    "ghscn_scn_forward": (I32, [P, P, P, P, I64, I64, I64, I64, I64, P, P, P, P, P, I32, P, P, P, P, P]),
    "ghscn_scn_backward_workspace_bytes": (SZ, [I64, I64, I64, I64]),
    "ghscn_scn_backward": (I32, [P, P, P, P, P, I64, I64, I64, I64, I64, P, I32, P, P, SZ, P]),
}


class GhscnError(RuntimeError):
    pass


class _Library:
    def __init__(self) -> None:
        if not LIB_PATH.exists():
            raise GhscnError(
                f"{LIB_PATH} is missing: build it with `python -m graph_hscn_b200.build` "
                "(or __graft_entry__.build()). graph_hscn_b200 has no CPU fallback.")
        self._dll = ctypes.CDLL(str(LIB_PATH), mode=os.RTLD_LOCAL | os.RTLD_NOW)
        self._fns = {}
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(self._dll, name)
            except AttributeError as e:
                raise GhscnError(f"libghscn.so does not export {name}; rebuild the library") from e
            fn.restype = res
            fn.argtypes = args
            self._fns[name] = fn
        self.launches = 0  # number of C-ABI compute calls issued (bench.py reports it)

    def error_string(self, code: int) -> str:
        return self._fns["ghscn_error_string"](code).decode()

    def query(self, name: str, *args):
        """Call a size/version query (no status code)."""
        return self._fns[name](*args)

    def call(self, name: str, *args) -> None:
        """Call a compute entry point; raise on a non-zero status."""
        self.launches += 1
        rc = self._fns[name](*args)
        if rc != 0:
            raise GhscnError(f"{name} failed: {self.error_string(rc)} (code {rc})")


_LIB = None


def lib() -> _Library:
    global _LIB
    if _LIB is None:
        _LIB = _Library()
    return _LIB
