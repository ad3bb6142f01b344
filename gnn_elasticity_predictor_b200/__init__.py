"""B200-native ALIGNN message-passing hot path (drop-in for conorjmoran/gnn-elasticity-predictor).

Public surface mirrors the reference's model classes (``scripts/train.py:303-401,528-586``); the
arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of ``include/alignn_b200.h``.
"""
from .modules import (AlignnRegressor, EdgeUpdateBlock, HeteroAlignnRegressor, NodeUpdateBlock,  # noqa: F401
                      TransformerConv, fused_gaussian_nll, gaussian_nll_loss)
from .ops import (GraphPlan, build_plan, build_pool_plan, conv_core, gate_ln_relu_residual,  # noqa: F401
                  segment_mean)
from .synthetic import GraphBatch, collate, make_crystal, synthetic_batch, zscore_targets  # noqa: F401
from .batching import bucket_shape, pad_batch, round_up_bucket  # noqa: F401

__all__ = [
    "AlignnRegressor", "EdgeUpdateBlock", "HeteroAlignnRegressor", "NodeUpdateBlock", "TransformerConv",
    "gaussian_nll_loss", "fused_gaussian_nll", "GraphPlan", "build_plan", "build_pool_plan", "conv_core", "gate_ln_relu_residual",
    "segment_mean", "GraphBatch", "collate", "make_crystal", "synthetic_batch", "zscore_targets", "bucket_shape", "pad_batch",
    "round_up_bucket",
]
__version__ = "0.1.0"
