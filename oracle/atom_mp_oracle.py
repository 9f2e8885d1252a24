"""CPU oracle of the atom-state message-passing variant. TEST INFRASTRUCTURE ONLY: nothing under
``notorch_b200/`` imports this file; ``tests/`` and ``__graft_entry__.smoke()`` do.

**Parity unpinned**: the reference tree contains no implementation of this variant (only the name ``nn.AtomMessagePassing``
in ``tests/integration/test_regression_rxn.py:40``), so there is nothing to pin against. This module is a plain-PyTorch
statement of the definition in SURVEY.md §8a row A10, written in the reference's idiom:

* ``scatter`` semantics = ``torch_scatter`` as composed in ``oracle/ref_shims/torch_scatter`` (sum = ``scatter_add_`` over the
  broadcast index; mean = sum / clamp(count, 1)), the call shape of ``notorch/nn/gnn/chemprop.py:39``;
* pre-activation and the running residual of ``notorch/nn/gnn/chemprop.py:37`` / ``notorch/nn/residual.py:27-28``;
* update = ``Sequential(Linear, Dropout)`` as in ``notorch/nn/gnn/chemprop.py:26``.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch import Tensor


def _scatter(x: Tensor, index: Tensor, dim_size: int, reduce: str) -> Tensor:
    out = torch.zeros((dim_size, x.shape[1]), dtype=x.dtype).scatter_add_(0, index[:, None].expand_as(x), x)
    if reduce == "mean":
        count = torch.zeros(dim_size, dtype=x.dtype).scatter_add_(0, index, torch.ones(len(index), dtype=x.dtype)).clamp_(min=1)
        out = out / count[:, None]
    return out


class AtomMessagePassingOracle(nn.Module):
    """Same constructor surface and parameter names as ``notorch_b200.nn.AtomMessagePassing`` (``layers.{i}[.module].update.0.*``)."""

    def __init__(self, hidden_dim: int = 256, act: type[nn.Module] = nn.ReLU, bias: bool = True, dropout: float = 0.0, depth: int = 3,
                 residual: bool = True, shared: bool = False, reduce: str = "sum"):
        super().__init__()

        def make():
            layer = nn.Module()
            layer.act = act()
            layer.update = nn.Sequential(nn.Linear(hidden_dim, hidden_dim, bias), nn.Dropout(dropout))
            return layer

        one = make()
        layers = [one if shared else make() for _ in range(depth)]
        if residual:
            wrapped = []
            for layer in layers:
                w = nn.Module()
                w.module = layer
                wrapped.append(w)
            layers = wrapped
        self.layers = nn.ModuleList(layers)
        self.residual, self.reduce = residual, reduce

    def forward(self, node_feats: Tensor, edge_feats: Tensor, edge_index: Tensor) -> Tensor:
        src, dst = edge_index[0], edge_index[1]
        h = node_feats
        for entry in self.layers:
            layer = entry.module if self.residual else entry
            a = layer.act(h)
            msg = a[src] + edge_feats
            n = _scatter(msg, dst, len(h), self.reduce)
            u = layer.update(n)
            h = h + u if self.residual else u
        return h
