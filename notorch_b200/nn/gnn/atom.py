"""Atom-state message passing — the "atom-based message passing variant" BASELINE configs[3] names.

**Extension, parity unpinned**: the reference tree has no such module (its only trace is the name
``nn.AtomMessagePassing`` in the stale ``tests/integration/test_regression_rxn.py:40``); SURVEY.md §8a row A10 defines it
in the reference's own idiom (pre-activation, running residual, ``Sequential(Linear, Dropout)`` update, the constructor
surface of ``ChempropBlock`` — ``notorch/nn/gnn/chemprop.py:49-75``):

    h_0     = x_v
    a       = act(h_l)
    msg[e]  = a[src[e]] + x_e[e]
    n[v]    = reduce_{e: dst[e] = v} msg[e]                       (sum | mean)
    h_{l+1} = [h_l +] Dropout(Linear(n))
    out     = G.update(node_feats = h_L)                          (edge_feats pass through)

The oracle is ``oracle/atom_mp_oracle.py`` (a plain PyTorch module written against this definition).
"""
from __future__ import annotations

import torch.nn as nn
from torch import Tensor

from ... import ops
from ...data.models.graph import BatchedGraph, Graph
from ...types import Reduction
from ..residual import Residual


class AtomMessagePassingLayer(nn.Module):
    def __init__(self, hidden_dim: int, act: type[nn.Module] = nn.ReLU, bias: bool = True, dropout: float = 0.0,
                 reduce: Reduction = "sum"):
        super().__init__()
        self.act = act()
        self.reduce = reduce
        self.update = nn.Sequential(nn.Linear(hidden_dim, hidden_dim, bias), nn.Dropout(dropout))

    def forward(self, node_feats: Tensor, edge_aggregate: Tensor, acsr: ops.AtomCSR, *, _residual: bool = False) -> Tensor:
        linear, drop = self.update[0], self.update[1]
        return ops.atom_layer(node_feats, edge_aggregate, linear.weight, linear.bias, acsr, act=ops.act_code(self.act), reduce=self.reduce,
                              residual=_residual, dropout=drop.p, training=self.training and drop.training)

    def extra_repr(self):
        return f"(reduce): {self.reduce}"


class AtomMessagePassing(nn.Module):
    def __init__(self, hidden_dim: int = 256, act: type[nn.Module] = nn.ReLU, bias: bool = True, dropout: float = 0.0,
                 depth: int = 3, residual: bool = True, shared: bool = False, reduce: Reduction = "sum"):
        super().__init__()
        if shared:
            one = AtomMessagePassingLayer(hidden_dim, act, bias, dropout, reduce)
            layers = [one for _ in range(depth)]
        else:
            layers = [AtomMessagePassingLayer(hidden_dim, act, bias, dropout, reduce) for _ in range(depth)]
        if residual:
            layers = [Residual(layer) for layer in layers]
        self.layers = nn.ModuleList(layers)
        self.hidden_dim = hidden_dim
        self.reduce = reduce

    @property
    def depth(self) -> int:
        return len(self.layers)

    def forward(self, G: Graph | BatchedGraph):
        if self.reduce not in ("sum", "mean"):
            raise NotImplementedError(f"notorch_b200: reduce='{self.reduce}' is not implemented (sum and mean are); no fallback")
        csr = ops.graph_csr(G)
        acsr = ops.atom_csr(csr)
        s_e = ops.edge_to_atom(G.edge_feats, csr, self.reduce)  # reduce_dst(x_e): the same at every depth, computed once (K1)
        h = G.node_feats
        for entry in self.layers:
            if isinstance(entry, Residual):
                h = entry.module(h, s_e, acsr, _residual=True)
            else:
                h = entry(h, s_e, acsr)
        return G.update(node_feats=h)
