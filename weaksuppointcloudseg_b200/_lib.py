"""ctypes binding of libwspc.so (the C ABI declared in include/wspc.h).

The shared library is built in-tree by ``__graft_entry__.build()`` /
``make -C weaksuppointcloudseg_b200/csrc``.  There is no CPU or PyTorch
fallback: if the library is missing, or the device is not sm_100, every
operator raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_longlong, c_size_t, c_uint64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwspc.so")

DIST_TFUTIL = 0
DIST_SMOOTH = 1


class WspcError(RuntimeError):
    pass


_lib = None


def _declare(lib):
    P = c_void_p
    lib.wspc_version.restype = c_int
    lib.wspc_last_error.restype = c_char_p
    lib.wspc_launch_count.restype = c_uint64
    lib.wspc_knn_workspace_bytes.restype = c_size_t
    lib.wspc_knn_workspace_bytes.argtypes = [c_int, c_int, c_int]
    lib.wspc_knn_fused.restype = c_int
    lib.wspc_knn_fused.argtypes = [P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, c_size_t, P]
    lib.wspc_pairwise_distance.restype = c_int
    lib.wspc_pairwise_distance.argtypes = [P, c_int, c_int, c_int, c_int, c_int, c_int, P, P, c_size_t, P]
    lib.wspc_topk_rows.restype = c_int
    lib.wspc_topk_rows.argtypes = [P, c_longlong, c_int, c_int, P, P, P]
    for name, spec in _EXTRA_DECLS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = spec


# filled by the other binding sections below (name -> (restype, argtypes))
_EXTRA_DECLS: dict = {}


def lib():
    """Load (once) and return the ctypes handle; raise loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise WspcError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU/PyTorch fallback for the wspc hot path)")
        h = ctypes.CDLL(LIB_PATH)
        _declare(h)
        _lib = h
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise WspcError(f"libwspc error {rc}: {lib().wspc_last_error().decode()}")


def launch_count() -> int:
    return int(lib().wspc_launch_count())


def ptr(t) -> c_void_p:
    return c_void_p(0 if t is None else t.data_ptr())


def stream() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise WspcError("wspc operators take CUDA tensors only (no CPU fallback)")
        if not t.is_contiguous():
            raise WspcError("wspc operators take contiguous tensors")


_ws = {}


def workspace(nbytes: int, device, slot: str = "default") -> torch.Tensor:
    """Grow-only per-(device, slot) scratch buffer; stream-ordered reuse."""
    key = (torch.device(device).index, slot)
    buf = _ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _ws[key] = buf
    return buf


# ---------------------------------------------------------------------------------------------
# struct mirrors of include/wspc.h
# ---------------------------------------------------------------------------------------------
c_double = ctypes.c_double
_P = c_void_p


class Operand(ctypes.Structure):
    """wspc_operand_t"""
    _fields_ = [("p", _P), ("ld", c_longlong), ("C", c_int), ("sc", _P), ("sh", _P), ("dmask", _P),
                ("dscale", c_float), ("idx", _P), ("k", c_int), ("npts", c_int), ("y", _P), ("ldy", c_longlong),
                ("c1", _P), ("c2", _P), ("c3", _P), ("dg", _P), ("amax", _P)]


class Epilogue(ctypes.Structure):
    """wspc_epilogue_t"""
    _fields_ = [("out", _P), ("ldo", c_longlong), ("bias", _P), ("rowbias", _P), ("rb_rows", c_int),
                ("ldrb", c_longlong), ("stats", _P), ("yprev", _P), ("ldyp", c_longlong), ("scp", _P), ("shp", _P),
                ("dmask", _P), ("dscale", c_float), ("dx", _P), ("lddx", c_longlong), ("idx", _P), ("k", c_int),
                ("npts", c_int)]


OP_PLAIN, OP_BNRELU, OP_EDGE, OP_DY, OP_DY_SPARSE, OP_DY_MAXK, OP_IMG = range(7)
EPI_STORE, EPI_STORE_STATS, EPI_RELUMASK_STATS, EPI_ACCUM, EPI_EDGE_SCATTER = range(5)

_OPP = ctypes.POINTER(Operand)
_EPP = ctypes.POINTER(Epilogue)
_EXTRA_DECLS.update({
    "wspc_conv1x1_rows": (c_int, [_OPP, c_int, _P, c_longlong, c_int, c_longlong, c_int, c_int, _EPP, c_int, _P]),
    "wspc_conv1x1_rows_workspace_bytes": (c_size_t, [c_int, c_int]),
    "wspc_conv1x1_rows_ws": (c_int, [_OPP, c_int, _P, c_longlong, c_int, c_longlong, c_int, c_int, _EPP, c_int, _P, c_size_t, _P]),
    "wspc_conv1x1_wgrad_workspace_bytes": (c_size_t, [c_int, c_int]),
    "wspc_conv1x1_wgrad": (c_int, [_OPP, c_int, _OPP, c_int, c_longlong, _P, _P, _P, c_size_t, _P]),
    "wspc_conv1x1_bwd_fused": (c_int, [_OPP, c_int, _OPP, c_int, c_longlong, _P, c_longlong, _EPP, _P, _P, _P, c_size_t, _P]),
    "wspc_bn_finalize": (c_int, [_P, c_int, c_double, _P, _P, c_float, c_float, c_int, _P, _P, _P, _P, _P, _P, _P]),
    "wspc_bn_bwd_coeffs": (c_int, [_P, c_int, c_double, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "wspc_maxk_bnrelu_fwd": (c_int, [_P, _P, _P, c_longlong, c_int, c_int, _P, c_longlong, _P]),
    "wspc_maxk_bnrelu_bwd": (c_int, [_P, _P, _P, _P, c_longlong, _P, c_longlong, c_longlong, c_int, c_int, _P, _P, _P]),
    "wspc_maxk_bnrelu_bwd_stats": (c_int, [_P, _P, _P, _P, c_longlong, _P, c_longlong, c_longlong, c_int, c_int, _P, _P, _P]),
    "wspc_edge_combine_bwd_maxk": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_longlong, c_int, c_int, c_int, _P, c_longlong, _P]),
    "wspc_maxn_bnrelu_fwd": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, _P, _P]),
    "wspc_maxn_bwd_gate": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P]),
    "wspc_conv1x1_pool_supported": (c_int, [c_longlong, c_int, c_int, c_int]),
    "wspc_conv1x1_pool_fwd": (c_int, [_OPP, c_int, _P, c_longlong, c_longlong, c_int, c_int, c_int, _P, _P, _P, _P, _P, c_size_t,
                                      _P]),
    "wspc_maxn_from_keys": (c_int, [_P, _P, _P, _P, c_int, c_int, _P, _P, _P, _P]),
    "wspc_rows_image_bytes": (c_size_t, [c_longlong, c_int]),
    "wspc_rows_image_supported": (c_int, [c_longlong, c_int, c_int]),
    "wspc_rows_image": (c_int, [_P, c_longlong, c_longlong, c_int, _P, _P]),
    "wspc_cloud_colsum": (c_int, [_OPP, c_int, c_int, _P, _P]),
    "wspc_head_losses_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "wspc_head_losses": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_float, c_float, c_int, c_int, _P, _P,
                                 _P, _P, c_size_t, _P]),
    "wspc_adam_tf": (c_int, [_P, _P, _P, _P, c_longlong, c_float, c_float, c_float, c_float, c_float, _P]),
    "wspc_dropout_mask": (c_int, [_P, c_longlong, c_float, c_uint64, c_uint64, _P]),
    "wspc_zero": (c_int, [_P, c_size_t, _P]),
    "wspc_set_gemm_path": (c_int, [c_int]),
    "wspc_set_knn_path": (c_int, [c_int]),
    "wspc_knn_fallback_rows": (c_int, [_P, c_int, c_int, c_int, ctypes.POINTER(c_int)]),
    "wspc_smooth_loss": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_float, _P, _P, _P, c_size_t, _P]),
    "wspc_smooth_loss_ex": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_float, c_int, _P, _P, _P, c_size_t, _P]),
    "wspc_laplacian_sym": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_float, c_float, _P, _P, _P]),
    "wspc_lp_blocks_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "wspc_lp_blocks": (c_int, [_P, _P, c_int, c_int, c_int, c_float, c_float, c_int, c_float, _P, _P, _P, _P, _P, _P, _P,
                               c_size_t, _P]),
    "wspc_lp_solve_workspace_bytes": (c_size_t, [c_int, c_int]),
    "wspc_lp_solve": (c_int, [_P, _P, c_int, c_int, c_float, c_float, c_int, c_float, _P, _P, _P, ctypes.POINTER(c_int), _P, c_size_t, _P]),
    "wspc_gather": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_longlong, c_int, _P, _P]),
    "wspc_bn_apply": (c_int, [_P, _P, _P, c_longlong, c_int, c_int, _P, _P]),
    "wspc_transform_points_fwd": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P]),
    "wspc_transform_points_bwd": (c_int, [_P, _P, c_int, c_int, _P, _P]),
    "wspc_edge_split_weights": (c_int, [_P, c_int, c_int, _P, _P]),
    "wspc_edge_combine_fwd": (c_int, [_P, c_longlong, _P, _P, c_longlong, c_int, c_int, c_int, _P, _P, _P]),
    "wspc_edge_combine_fwd_extrema": (c_int, [_P, c_longlong, _P, _P, c_longlong, c_int, c_int, c_int, _P, _P, _P, _P]),
    "wspc_maxk_from_extrema": (c_int, [_P, _P, _P, c_longlong, c_int, _P, c_longlong, _P]),
    "wspc_edge_combine_bwd": (c_int, [_P, _P, _P, _P, _P, _P, c_longlong, c_int, c_int, c_int, _P, c_longlong, _P]),
    "wspc_edge_merge_wgrad": (c_int, [_P, _P, c_int, c_int, _P, _P, _P]),
    "wspc_fill_rows": (c_int, [_P, c_longlong, _P, c_longlong, c_int, _P]),
    "wspc_poolconv_coeffs": (c_int, [_P, _P, _P, _P, c_int, c_int, _P, _P, _P]),
    "wspc_poolconv_sparse": (c_int, [_P, _P, _P, _P, _P, c_longlong, c_int, c_int, c_int, c_int, _P, c_longlong, _P, _P, _P]),
    "wspc_poolconv_finalize": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_double, _P, _P, _P]),
    "wspc_edge_gather_stats": (c_int, [_P, c_longlong, _P, _P, c_longlong, c_int, c_int, c_int, _P, _P, _P, _P, _P]),
    "wspc_edgeconv2_fwd": (c_int, [_P, c_longlong, _P, _P, _P, _P, _P, _P, c_longlong, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "wspc_maxk_extrema_bwd_prep": (c_int, [_P, _P, _P, c_longlong, _P, c_longlong, c_longlong, c_int, _P, _P, _P]),
    "wspc_edgeconv2_bwd_workspace_bytes": (c_size_t, []),
    "wspc_edgeconv2_bwd": (c_int, [_P, c_longlong, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_longlong, c_int, c_int,
                                   c_int, c_int, _P, _P, _P, c_size_t, _P]),
    "wspc_edge1_bwd": (c_int, [_P, c_longlong, _P, _P, _P, _P, _P, c_longlong, _P, c_longlong, c_longlong, c_int, c_int, c_int,
                               _P, _P]),
    "wspc_edgeconv2_bwd_ex": (c_int, [_P, c_longlong, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_longlong, c_int, c_int,
                                      c_int, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "wspc_edge1_bwd_ex": (c_int, [_P, c_longlong, _P, _P, _P, _P, _P, c_longlong, _P, c_longlong, c_longlong, c_int, c_int,
                                  c_int, _P, _P, _P]),
    "wspc_edge_bwd_stats": (c_int, [_P, _P, c_longlong, _P, c_longlong, c_int, _P, _P]),
    "wspc_edge_bwd_finalize": (c_int, [_P, _P, _P, _P, c_longlong, _P, _P, _P, _P, c_longlong, c_int, c_int, _P, c_longlong, _P]),
    "wspc_zero_cols": (c_int, [_P, c_longlong, c_int, c_int, c_longlong, _P]),
    "wspc_bn_bias_grad": (c_int, [_P, _P, _P, _P, _P, c_int, c_double, _P, _P]),
})


def dptr(t) -> int:
    """raw device address (0 for None) for struct fields"""
    return 0 if t is None else t.data_ptr()
