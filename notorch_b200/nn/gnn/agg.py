"""Atom -> molecule read-outs — drop-in for ``Sum`` / ``Mean`` of ``notorch/nn/gnn/agg.py:23-38``
(K3: deterministic segmented reduction over the CSR of ``batch_node_index``, no atomics), plus the
builder-defined ``Norm`` extension BASELINE config 4 names (not in the
reference tree, parity unpinned). ``Max``, ``Gated`` and ``SDPAttention`` (agg.py:41-86) are the
"next" row N2 and raise until their kernels exist — by design there is no fallback.
"""
from __future__ import annotations

from abc import abstractmethod

import torch.nn as nn
from torch import Tensor

from ... import ops
from ...data.models.graph import BatchedGraph


def _mol_csr(G: BatchedGraph) -> ops.SegmentCSR:
    return ops.segment_csr_for(G, "batch_node_index", len(G))


class Aggregation(nn.Module):
    @abstractmethod
    def forward(self, G: BatchedGraph, **kwargs) -> Tensor:
        pass


class Sum(Aggregation):
    def forward(self, G: BatchedGraph, **kwargs) -> Tensor:
        return ops.readout(G.node_feats, _mol_csr(G), "sum")


class Mean(Aggregation):
    def forward(self, G: BatchedGraph, **kwargs) -> Tensor:
        return ops.readout(G.node_feats, _mol_csr(G), "mean")


class Norm(Aggregation):
    """``sum / norm`` (chemprop's NormAggregation semantics) — extension, not in the reference."""

    def __init__(self, norm: float = 100.0):
        super().__init__()
        self.norm = float(norm)

    def forward(self, G: BatchedGraph, **kwargs) -> Tensor:
        return ops.readout(G.node_feats, _mol_csr(G), "norm", self.norm)


class _NotYet(Aggregation):
    def forward(self, G: BatchedGraph, **kwargs) -> Tensor:
        raise NotImplementedError(f"notorch_b200: {type(self).__name__} read-out has no sm_100a kernel yet (SURVEY.md §8f N2); no fallback")


class Max(_NotYet):
    pass


class Gated(_NotYet):
    def __init__(self, input_dim: int = 256):
        super().__init__()
        self.a = nn.Linear(input_dim, 1)


class SDPAttention(_NotYet):
    def __init__(self, key_dim: int = 256):
        super().__init__()
        self.sqrt_key_dim = key_dim ** 0.5
