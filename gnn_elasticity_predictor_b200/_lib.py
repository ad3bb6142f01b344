"""ctypes binding of ``libalignn_b200.so`` -- the C ABI declared in ``include/alignn_b200.h``.

There is NO CPU fallback: if the library is missing and cannot be built, or a kernel is asked to
run on a non-CUDA tensor, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

from . import build as _build

_LOCK = threading.Lock()
_LIB = None

_P = c_void_p


class GraphStoreStruct(ctypes.Structure):
    """``alignn_graph_store`` (include/alignn_b200.h)."""
    _fields_ = ([(n, c_void_p) for n in ("x", "edge_attr", "lg_edge_attr", "global_x", "sg_one_hot", "y", "edge_index",
                                         "lg_edge_index", "node_ptr", "bond_ptr", "angle_ptr")]
                + [(n, c_int64) for n in ("n_graphs", "n_nodes", "n_bonds", "n_angles")]
                + [(n, c_int32) for n in ("node_dim", "edge_dim", "angle_dim", "global_dim", "sg_dim", "target_dim")])


class BatchOutStruct(ctypes.Structure):
    """``alignn_batch_out`` (include/alignn_b200.h)."""
    _fields_ = ([(n, c_void_p) for n in ("x", "edge_attr", "lg_edge_attr", "global_x", "sg_one_hot", "y", "edge_index",
                                         "lg_edge_index", "batch", "train_idx")]
                + [(n, c_int64) for n in ("n_graphs", "n_nodes", "n_bonds", "n_angles")])

# name -> (restype, argtypes); mirrors include/alignn_b200.h one to one
SIGNATURES = {
    "alignn_abi_version": (c_int, []),
    "alignn_error_string": (c_char_p, [c_int]),
    "alignn_plan_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "alignn_build_plan": (c_int, [_P, c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "alignn_build_plan_ex": (c_int, [_P, c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, c_int, _P]),
    "alignn_build_plan_bounded": (c_int, [_P, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, c_int, _P]),
    "alignn_plan_csc_positions": (c_int, [_P, _P, c_int64, _P, _P, _P]),
    "alignn_conv_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int, c_int, c_int,
                                c_float, c_uint64, c_uint64, _P]),
    "alignn_conv_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                c_int64, c_int64, c_int, c_int, c_int, c_float, c_uint64, c_uint64, _P]),
    "alignn_gate_ln_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_float,
                                   c_float, c_uint64, c_uint64, _P]),
    "alignn_gate_ln_bwd_partial_rows": (c_int64, []),
    "alignn_gate_ln_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int,
                                   c_float, c_uint64, c_uint64, _P]),
    "alignn_edgeattn_supported": (c_int, [c_int, c_int]),
    "alignn_edgeattn_fwd": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                    c_int64, c_int64, c_int, c_int, c_int, c_float, c_uint64, c_uint64, _P]),
    "alignn_edgeattn_bwd_dst": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P,
                                        _P, _P, _P, _P, c_int64, _P, _P, _P, _P, c_int,
                                        c_int64, c_int64, c_int, c_int, c_int, c_float, c_uint64, c_uint64, _P]),
    "alignn_edgeattn_bwd_src": (c_int, [_P, _P, c_int64, _P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int64,
                                        c_int, c_int, c_int, _P]),
    "alignn_edgeattn_bwd_src_lp": (c_int, [_P, _P, c_int64, _P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int64,
                                           c_int, c_int, c_int, _P]),
    "alignn_edgeattn_mma_supported": (c_int, [c_int, c_int, c_int]),
    "alignn_edgeattn_mma_fwd": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                        c_int64, c_int64, c_int, c_int, c_int, c_float, c_uint64, c_uint64, _P]),
    "alignn_edgeattn_mma_bwd_dst": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P,
                                            _P, _P, _P, _P, c_int64, _P, _P, _P, _P, c_int,
                                            c_int64, c_int64, c_int, c_int, c_int, c_float, c_uint64, c_uint64, _P]),
    "alignn_edgeattn_mma_fwd_s": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, _P, c_int64, c_int64, _P, _P, _P, _P,
                                          _P, _P, c_int64, c_int64, _P, _P, _P, c_int64, c_int64, c_int, c_int, c_int,
                                          c_float, c_uint64, c_uint64, _P, _P]),
    "alignn_edgeattn_mma_bwd_dst_s": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int64, _P, c_int64, c_int64,
                                              _P, c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, _P, c_int64,
                                              c_int64, _P, _P, _P, c_int64, c_int, c_int64, c_int64, c_int, c_int, c_int,
                                              c_float, c_uint64, c_uint64, _P, _P]),
    "alignn_lgattn_supported": (c_int, [c_int, c_int, c_int, c_int]),
    "alignn_lg_pack_angles": (c_int, [_P, _P, _P, c_int64, c_int, _P]),
    "alignn_lgattn_fwd": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, _P, c_int64, c_int64, _P, _P, _P, c_int,
                                  _P, _P, _P, _P, c_int64, c_int64, _P, _P, _P,
                                  c_int64, c_int64, c_int, c_int, c_int, c_float, c_uint64, c_uint64, _P, _P, _P]),
    "alignn_lgattn_fwd_tc": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, _P, c_int64, c_int64, _P, _P, _P, c_int,
                                     _P, _P, _P, _P, c_int64, c_int64, _P, _P, _P,
                                     c_int64, c_int64, c_int, c_int, c_int, c_float, c_uint64, c_uint64, _P, _P]),
    "alignn_lgattn_bwd_dst": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int64, _P, c_int64, c_int64, _P, c_int64,
                                      c_int64, _P, _P, _P, _P, c_int, _P, _P, _P, _P, _P, c_int64, _P, c_int64, c_int64,
                                      _P, c_int64, c_int64, c_int, c_int, c_int, c_float, c_uint64, c_uint64, _P, _P, _P]),
    "alignn_lg_angle_grad_partial_floats": (c_int64, [c_int64, c_int64]),
    "alignn_lg_angle_grad": (c_int, [_P, _P, _P, c_int, _P, c_int, _P, _P, _P, c_int64, c_int64, c_int64, c_int64,
                                     _P, _P, c_int64, c_int64, _P]),
    "alignn_lg_angle_grad2": (c_int, [_P, _P, _P, c_int, _P, c_int, _P, _P, _P, c_int64, c_int64, c_int64, c_int64,
                                     _P, _P, c_int64, c_int64, c_int, _P]),
    "alignn_adamw_partial_floats": (c_int64, []),
    "alignn_clip_adamw_step": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int64,
                                       c_float, c_float, c_float, c_float, c_float, c_float, _P]),
    "alignn_gate_ln_fwd2": (c_int, [_P, _P, _P, _P, c_int, _P, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                    c_int64, c_int, c_int, c_float, c_float, c_uint64, c_uint64, _P, _P]),
    "alignn_gate_ln_bwd2": (c_int, [_P, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, _P, _P,
                                    c_int64, c_int, c_int, c_float, c_uint64, c_uint64, _P, _P]),
    "alignn_gate_ln_fwd3": (c_int, [_P, _P, _P, _P, c_int, c_int64, _P, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                    c_int64, c_int, c_int, c_float, c_float, c_uint64, c_uint64, _P, _P]),
    "alignn_gate_ln_fwd4": (c_int, [_P, _P, _P, _P, c_int, c_int64, _P, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                    c_int64, c_int, c_int, c_float, c_float, c_uint64, c_uint64, _P, _P]),
    "alignn_gate_ln_bwd3": (c_int, [_P, _P, c_int64, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, _P, c_int, c_int64, _P, _P, _P, c_int64, _P, _P,
                                    c_int64, c_int, c_int, c_float, c_uint64, c_uint64, _P, _P]),
    "alignn_colsum_supported": (c_int, [c_int]),
    "alignn_colsum_partial_floats": (c_int64, [c_int]),
    "alignn_colsum": (c_int, [_P, c_int64, c_int64, c_int, c_int, _P, _P, _P]),
    "alignn_angle_supported": (c_int, [c_int, c_int]),
    "alignn_angle_h1_fwd": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P]),
    "alignn_angle_partial_floats": (c_int64, [c_int]),
    "alignn_angle_h1_bwd": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P]),
    "alignn_collate": (c_int, [_P, _P, c_int64, c_int, _P, _P, _P, _P, _P, _P]),
    "alignn_bond_features": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int64, _P, c_int, c_double, _P, _P, _P]),
    "alignn_linegraph_count": (c_int, [_P, _P, _P, _P, c_int64, _P, _P]),
    "alignn_linegraph_fill": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, _P, c_int, c_double, _P, c_int64,
                                      _P, _P]),
    "alignn_wgrad_supported": (c_int, [c_int, c_int]),
    "alignn_wgrad_partial_floats": (c_int64, [c_int64, c_int]),
    "alignn_wgrad": (c_int, [_P, c_int64, _P, c_int64, c_int64, c_int, c_int, c_int, _P, _P, _P, _P]),
    "alignn_proj_tc_supported": (c_int, [c_int, c_int, c_int]),
    "alignn_proj_tc": (c_int, [_P, c_int64, _P, c_int64, _P, _P, c_int64, c_int64, c_int, c_int, c_int, _P]),
    "alignn_proj_tc2": (c_int, [_P, c_int64, _P, c_int64, c_int64, _P, _P, c_int64, c_int, c_int, _P, c_int64, c_int, c_int,
                                c_int64, c_int64, c_int, c_int, _P]),
    "alignn_gaussian_nll": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, c_float, c_float, _P, _P, _P, _P]),
    "alignn_ensemble_post": (c_int, [_P, _P, c_int, c_int64, c_int, c_float, _P, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "alignn_segment_mean_fwd": (c_int, [_P, _P, _P, _P, c_int64, c_int, _P]),
    "alignn_segment_mean_bwd": (c_int, [_P, _P, _P, _P, c_int64, c_int, _P]),
}

ABI_VERSION = 22
F32, BF16 = 0, 1


def lib_path() -> str:
    return _build.LIB_PATH


def load(rebuild_if_stale: bool = True):
    """Load (building first if needed) the shared library; raises ``RuntimeError`` on failure."""
    global _LIB
    if _LIB is not None:
        return _LIB
    with _LOCK:
        if _LIB is not None:
            return _LIB
        path = _build.LIB_PATH
        need_build = not os.path.exists(path)
        if not need_build and rebuild_if_stale and not _build.is_current():
            need_build = True
        if need_build:
            try:
                _build.build()
            except Exception as exc:  # noqa: BLE001
                # never fall back to a stale binary: a library that does not match the sources on disk would pass the
                # ABI integer check and silently run old kernels
                what = "is missing" if not os.path.exists(path) else "is older than csrc/ (stale)"
                raise RuntimeError(
                    f"libalignn_b200.so {what} and could not be rebuilt; the ALIGNN hot path has no CPU or PyTorch "
                    f"fallback.  Fix the build (`python -m gnn_elasticity_predictor_b200.build --force`) or pass "
                    f"rebuild_if_stale=False to load the existing binary on purpose ({exc})") from exc
        try:
            lib = ctypes.CDLL(path)
        except OSError as exc:
            raise RuntimeError(f"cannot load {path}: {exc} (no fallback path exists)") from exc
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as exc:
                raise RuntimeError(f"{path} does not export {name}; rebuild with "
                                   "`python -m gnn_elasticity_predictor_b200.build --force`") from exc
            fn.restype = res
            fn.argtypes = args
        got = lib.alignn_abi_version()
        if got != ABI_VERSION:
            raise RuntimeError(f"{path}: ABI version {got}, binding expects {ABI_VERSION}; rebuild the library")
        _LIB = lib
        return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().alignn_error_string(rc)
        raise RuntimeError(f"{what} failed: {msg.decode() if msg else rc} (code {rc})")
