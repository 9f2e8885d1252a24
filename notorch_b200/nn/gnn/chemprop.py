"""Directed-bond message passing — drop-in for ``notorch/nn/gnn/chemprop.py`` (same class names,
constructor arguments, attributes, parameter names and ``forward`` signatures), computed by the
sm_100a kernels in ``libnotorch_b200.so``.

Arithmetic (the reference's, not textbook chemprop — SURVEY.md §0 item 2):

    h_0     = x_v[src] + x_e                                            chemprop.py:83
    a       = act(h_l);  n = scatter(a, dst, V, reduce)                 chemprop.py:37,39
    m       = n[src] - a[rev]                                           chemprop.py:40
    h_{l+1} = [h_l +] Dropout(Linear(m))                                chemprop.py:41, residual.py:28
    out     = (scatter(h_L, dst, V, reduce), h_L)                       chemprop.py:86,88
"""
from __future__ import annotations

import torch.nn as nn
from torch import Tensor

from ... import ops
from ...data.models.graph import BatchedGraph, Graph
from ...types import Reduction
from ..residual import Residual


class ChempropLayer(nn.Module):
    def __init__(self, hidden_dim: int, act: type[nn.Module] = nn.ReLU, bias: bool = True, dropout: float = 0.0,
                 reduce: Reduction = "sum"):
        super().__init__()
        self.act = act()
        self.reduce = reduce
        self.update = nn.Sequential(nn.Linear(hidden_dim, hidden_dim, bias), nn.Dropout(dropout))

    def forward(self, edge_feats: Tensor, node_feats: Tensor, edge_index: Tensor, rev_index: Tensor, *,
                _residual: bool = False, _csr: ops.GraphCSR | None = None) -> Tensor:
        # node_feats is used only for its length, exactly like the reference (chemprop.py:39)
        csr = _csr if _csr is not None else ops.graph_csr_from_tensors(edge_index, rev_index, len(node_feats))
        linear, drop = self.update[0], self.update[1]
        return ops.layer(edge_feats, linear.weight, linear.bias, csr, act=ops.act_code(self.act), reduce=self.reduce,
                         residual=_residual, dropout=drop.p, training=self.training and drop.training)

    def extra_repr(self):
        return f"(reduce): {self.reduce}"


class ChempropBlock(nn.Module):
    def __init__(self, hidden_dim: int = 256, act: type[nn.Module] = nn.ReLU, bias: bool = True, dropout: float = 0.0,
                 depth: int = 3, residual: bool = True, shared: bool = False, reduce: Reduction = "sum"):
        super().__init__()
        if shared:
            one = ChempropLayer(hidden_dim, act, bias, dropout, reduce)
            layers = [one for _ in range(depth)]  # the same module object at every depth (chemprop.py:65-66)
        else:
            layers = [ChempropLayer(hidden_dim, act, bias, dropout, reduce) for _ in range(depth)]
        if residual:
            layers = [Residual(layer) for layer in layers]
        self.layers = nn.ModuleList(layers)
        self.hidden_dim = hidden_dim
        self.reduce = reduce

    @property
    def depth(self) -> int:
        return len(self.layers)

    def forward(self, G: Graph | BatchedGraph):
        csr = ops.graph_csr(G)
        h = ops.edge_init(G.node_feats, G.edge_feats, csr)  # K0
        for entry in self.layers:
            if isinstance(entry, Residual):
                h = entry.module(h, G.node_feats, G.edge_index, G.rev_index, _residual=True, _csr=csr)
            else:
                h = entry(h, G.node_feats, G.edge_index, G.rev_index, _csr=csr)
        node_hiddens = ops.edge_to_atom(h, csr, self.reduce)  # K1, no activation (chemprop.py:86)
        return G.update(node_feats=node_hiddens, edge_feats=h)
