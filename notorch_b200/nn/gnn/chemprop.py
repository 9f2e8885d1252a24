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

import torch
import torch.nn as nn
from torch import Tensor

from ... import ops
from ...data.models.graph import BatchedGraph, Graph, PendingFeats
from ...types import Reduction
from ..residual import Residual


class _Once:
    """A value computed on first use (the collapsed last depth: only a sum read-out asks for it)."""

    __slots__ = ("_fn", "_value")

    def __init__(self, fn):
        self._fn, self._value = fn, None

    def __call__(self):
        if self._fn is not None:
            self._value, self._fn = self._fn(), None
        return self._value


class ChempropLayer(nn.Module):
    """One message-passing depth. Parameters live in ``update = Sequential(Linear(d, d), Dropout(p))`` so that a reference
    checkpoint loads with ``strict=True``; the modules in ``update`` are containers only — K1 + K2 compute the layer."""

    def __init__(self, hidden_dim: int, act: type[nn.Module] = nn.ReLU, bias: bool = True, dropout: float = 0.0,
                 reduce: Reduction = "sum"):
        super().__init__()
        w_h = nn.Linear(hidden_dim, hidden_dim, bias=bias)
        self.act, self.reduce = act(), reduce
        self.update = nn.Sequential(w_h, nn.Dropout(p=dropout))

    def forward(self, edge_feats: Tensor, node_feats: Tensor, edge_index: Tensor, rev_index: Tensor, *,
                _residual: bool = False, _csr: ops.GraphCSR | None = None) -> Tensor:
        # only len(node_feats) matters, as in the reference (chemprop.py:39); a block passes its cached CSR bundle and asks for the
        # residual to be added inside K2
        if _csr is None:
            _csr = ops.graph_csr_from_tensors(edge_index, rev_index, len(node_feats))
        w_h, drop = self.update
        return ops.layer(edge_feats, w_h.weight, w_h.bias, _csr, act=ops.act_code(self.act), reduce=self.reduce, residual=_residual,
                         dropout=drop.p, training=self.training and drop.training)

    def _pooled(self, edge_feats: Tensor, csr: ops.GraphCSR, pool: ops.SegmentCSR, residual: bool) -> Tensor:
        """``sum_{e in b} h'[e]`` of this depth without computing h' (the last depth of a block under a sum read-out)."""
        w_h, _ = self.update
        return ops.last_depth_pooled(edge_feats, w_h.weight, w_h.bias, csr, pool, act=ops.act_code(self.act), reduce=self.reduce, residual=residual)

    def extra_repr(self):
        return f"(reduce): {self.reduce}"


class ChempropBlock(nn.Module):
    """``depth`` message-passing layers between the edge initialisation and the final edge -> atom reduction. With ``shared=True``
    every depth holds the SAME layer object (chemprop.py:65-66), each wrapped in its own ``Residual`` when ``residual=True``."""

    def __init__(self, hidden_dim: int = 256, act: type[nn.Module] = nn.ReLU, bias: bool = True, dropout: float = 0.0,
                 depth: int = 3, residual: bool = True, shared: bool = False, reduce: Reduction = "sum"):
        super().__init__()
        stack: list[nn.Module] = []
        core = None
        for _ in range(depth):  # created in depth order: the same RNG consumption as the reference under torch.manual_seed
            if core is None or not shared:
                core = ChempropLayer(hidden_dim, act=act, bias=bias, dropout=dropout, reduce=reduce)
            stack.append(Residual(core) if residual else core)
        self.layers = nn.ModuleList(stack)
        self.hidden_dim, self.reduce = hidden_dim, reduce

    @property
    def depth(self) -> int:
        return len(self.layers)

    def forward(self, G: Graph | BatchedGraph):
        csr = ops.graph_csr(G)  # int32 CSR bundle, built once per batch and cached on the graph object
        xv, xe = ops.peek_feats(G, "node_feats"), ops.peek_feats(G, "edge_feats")
        if (isinstance(xv, PendingFeats) and isinstance(xe, PendingFeats) and xv.origin[0] is xe.origin[0]
                and not (xv.materialized or xe.materialized)):
            # the features are a GraphEmbedding that has not been computed: table look-ups + K0 in one kernel (row N1)
            h = ops.embed_edge_init(xv.origin[1], xe.origin[1], xv.origin[2], xe.origin[2], csr)
        else:
            xv = G.node_feats
            h = ops.edge_init(xv, G.edge_feats, csr)  # K0
        # A sum-reduced block on a device-collated batch (every molecule = a contiguous range of edges between its own atoms) defers
        # the final edge -> atom reduction: a Sum / Mean / Norm read-out right behind it needs only sum_{e in b} h_L[e] (§5.9).
        defer = self.reduce == "sum" and ops._fuse_readout and isinstance(G, Graph) and getattr(G, "_nt_mol_edge_ptr", None) is not None
        n_dense = len(self.layers)
        last = None
        if defer and n_dense > 0:
            entry = self.layers[-1]
            core = entry.module if isinstance(entry, Residual) else entry
            drop = core.update[1] if isinstance(core, ChempropLayer) else None
            if drop is not None and ops.last_depth_pooled_supported(h, drop.p, core.training and drop.training, core.reduce):
                # ... and it does not need h_L itself: the last depth is deferred as well. The read-out gets sum_{e in b} h_L[e] from
                # h_{L-1} on the molecules (§5.10); h_L = edge_feats (and node_feats) are computed only if something reads them.
                last, n_dense = (core, isinstance(entry, Residual)), n_dense - 1
        for entry in self.layers[:n_dense]:
            fused_residual = isinstance(entry, Residual)
            layer = entry.module if fused_residual else entry
            h = layer(h, xv, G.edge_index, G.rev_index, _residual=fused_residual, _csr=csr)  # xv: only its length is used
        if defer:
            # K1 without activation (chemprop.py:86), deferred: a Sum / Mean / Norm read-out right behind the block sums h_L over each
            # molecule's edges directly (one pass instead of K1 + K3); any other reader of node_feats computes them on first access
            if last is not None:
                core, fused_residual, h_prev, ei, ri = last[0], last[1], h, G.edge_index, G.rev_index
                grad_mode = torch.is_grad_enabled()

                def deferred(fn):  # whenever it runs, it runs in the autograd mode this forward was called in
                    def run():
                        with torch.set_grad_enabled(grad_mode):
                            return fn()
                    return run

                edges = PendingFeats(deferred(lambda: core(h_prev, xv, ei, ri, _residual=fused_residual, _csr=csr)), tuple(h.shape), h.dtype,
                                     h.device, ("last_depth", h_prev))
                pool = ops.mol_edge_csr(G)
                h_sum = _Once(deferred(lambda: core._pooled(h_prev, csr, pool, fused_residual)))
                atoms = PendingFeats(deferred(lambda: ops.edge_to_atom(edges.materialize(), csr, "sum")), (csr.V, h.shape[1]), h.dtype, h.device,
                                     ("edge_to_atom_sum", edges, h_sum))
                return G.update(node_feats=atoms, edge_feats=edges)
            final_h = h
            atoms = PendingFeats(lambda: ops.edge_to_atom(final_h, csr, "sum"), (csr.V, h.shape[1]), h.dtype, h.device, ("edge_to_atom_sum", final_h, None))
        else:
            atoms = ops.edge_to_atom(h, csr, self.reduce)  # K1 without activation (chemprop.py:86)
        return G.update(node_feats=atoms, edge_feats=h)
