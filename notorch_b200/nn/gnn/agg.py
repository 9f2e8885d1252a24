"""Atom -> molecule read-outs — drop-in for ``Sum`` / ``Mean`` of ``notorch/nn/gnn/agg.py:23-38``
(K3: deterministic segmented reduction over the CSR of ``batch_node_index``, no atomics), plus the
builder-defined ``Norm`` extension BASELINE config 4 names (not in the
reference tree, parity unpinned) and the row-N2 read-outs ``Max`` (pinned), ``Gated`` and ``SDPAttention``
(agg.py:41-86; both are broken in the reference, see their docstrings).
"""
from __future__ import annotations

from abc import abstractmethod

import torch
import torch.nn as nn
from torch import Tensor

from ... import ops
from ...data.models.graph import BatchedGraph


def _mol_csr(G: BatchedGraph) -> ops.SegmentCSR:
    ptr = getattr(G, "_nt_mol_ptr", None)
    if ptr is not None and ptr.device == G.batch_node_index.device:
        # collated on the device (BatchedGraph.from_packed): molecules are contiguous atom ranges, no permutation needed
        cached = getattr(G, "_nt_mol_csr", None)
        if cached is None:
            cached = ops.SegmentCSR(ptr, None, G.batch_node_index.to(torch.int32), len(G))
            G._nt_mol_csr = cached
        return cached
    return ops.segment_csr_for(G, "batch_node_index", len(G))


def _readout_over_edges(G: BatchedGraph, kind: str, norm: float = 100.0) -> Tensor | None:
    """``H[b] = sum_{v in b} sum_{e: dst[e] = v} h_L[e] = sum_{e in b} h_L[e]``: when the node features are the still-uncomputed
    edge -> atom SUM of a block (``PendingFeats`` left by ``ChempropBlock``) and the graph was collated by ``BatchedGraph.from_packed``
    (every molecule is a contiguous range of edges whose atoms are its own), Sum / Mean / Norm run over the edge states in one
    contiguous segmented reduction - K1 + K3 become one pass forward, and one gather backward. Returns ``None`` when it does not apply."""
    from ...data.models.graph import PendingFeats

    x = ops.peek_feats(G, "node_feats")
    eptr = getattr(G, "_nt_mol_edge_ptr", None)
    if not (ops._fuse_readout and isinstance(x, PendingFeats) and not x.materialized and x.origin[0] == "edge_to_atom_sum" and eptr is not None):
        return None
    h, h_sum = x.origin[1], x.origin[2] if len(x.origin) > 2 else None
    if eptr.device != x.device:
        return None
    if h_sum is not None:  # the block deferred its last depth as well: sum_{e in b} h_L[e] straight from h_{L-1} (ops._LastDepthPooled)
        H_sum = h_sum()
    else:
        if isinstance(h, PendingFeats):
            h = h.materialize()
        H_sum = ops.seg_reduce(h, ops.mol_edge_csr(G), "sum", tag="K3e")
    if kind == "norm":
        return H_sum * (1.0 / norm)
    H = H_sum
    if kind == "mean":  # scatter_mean = sum / clamp(count, 1) (agg.py:36): the count is the molecule's ATOM count
        cnt = getattr(G, "_nt_mol_atom_count", None)
        if cnt is None:
            aptr = G._nt_mol_ptr
            cnt = (aptr[1:] - aptr[:-1]).clamp(min=1).to(h.dtype).unsqueeze(1)
            G._nt_mol_atom_count = cnt
        H = H / cnt
    return H


class Aggregation(nn.Module):
    @abstractmethod
    def forward(self, G: BatchedGraph, **kwargs) -> Tensor:
        pass


class Sum(Aggregation):
    def forward(self, G: BatchedGraph, **kwargs) -> Tensor:
        H = _readout_over_edges(G, "sum")
        return H if H is not None else ops.readout(G.node_feats, _mol_csr(G), "sum")


class Mean(Aggregation):
    def forward(self, G: BatchedGraph, **kwargs) -> Tensor:
        H = _readout_over_edges(G, "mean")
        return H if H is not None else ops.readout(G.node_feats, _mol_csr(G), "mean")


class Norm(Aggregation):
    """``sum / norm`` (chemprop's NormAggregation semantics) — extension, not in the reference."""

    def __init__(self, norm: float = 100.0):
        super().__init__()
        self.norm = float(norm)

    def forward(self, G: BatchedGraph, **kwargs) -> Tensor:
        H = _readout_over_edges(G, "norm", self.norm)
        return H if H is not None else ops.readout(G.node_feats, _mol_csr(G), "norm", self.norm)


class Max(Aggregation):
    """``scatter_max`` read-out (agg.py:41-47); pinned against the reference (``tests/golden/readout_max.npz``)."""

    def forward(self, G: BatchedGraph, **kwargs) -> Tensor:
        from ... import readouts

        hiddens, _ = readouts.seg_max(G.node_feats, _mol_csr(G))
        return hiddens


class Gated(Aggregation):
    """Softmax-gated sum: ``alpha = softmax_b(Linear(x))``, ``H[b] = sum_v alpha[v] x[v]`` (agg.py:50-63).

    DEVIATION, labelled: the reference's ``forward`` unsqueezes ``alpha`` ([V,1] -> [V,1,1]) before multiplying it with
    ``node_feats`` [V,d], which broadcasts to [V,V,d] and returns a ``[b, V, d]`` tensor (agg.py:60-61, checked by
    running it). This implements the evident intent (``alpha`` of shape [V,1]); parity is pinned to the builder's
    restatement only (``oracle.dmpnn_oracle.readout_gated``). Parameter names match (``a.weight`` [1,d], ``a.bias`` [1])."""

    def __init__(self, input_dim: int = 256):
        super().__init__()
        self.a = nn.Linear(input_dim, 1)

    def forward(self, G: BatchedGraph, **kwargs) -> Tensor:
        from ... import readouts

        csr = _mol_csr(G)
        scores = readouts._RowDotVector.apply(G.node_feats, self.a.weight, self.a.bias)
        return readouts.attention_readout(G.node_feats, scores, csr)


class SDPAttention(Aggregation):
    """Scaled dot-product attention read-out against a per-molecule query ``Q`` [b,d] (agg.py:66-86).

    DEVIATION, labelled: the reference's ``forward`` raises — its einsum string ``"V d_v, V d_v -> V"`` is not a
    valid equation (agg.py:80, checked by running it). This implements what the string means:
    ``scores[v] = <Q[batch[v]], x[v]> / sqrt(key_dim)``; parity is pinned to the builder's restatement only."""

    def __init__(self, key_dim: int = 256):
        super().__init__()
        self.sqrt_key_dim = key_dim ** 0.5

    def forward(self, G: BatchedGraph, *, Q: Tensor, **kwargs) -> Tensor:
        from ... import readouts

        csr = _mol_csr(G)
        scores = readouts._RowDotSegment.apply(G.node_feats, Q, csr, 1.0 / self.sqrt_key_dim)
        return readouts.attention_readout(G.node_feats, scores, csr)
