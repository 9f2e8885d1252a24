"""``GraphEmbedding`` — drop-in for ``notorch/nn/gnn/embed.py:11-36`` (row N1 of SURVEY.md §8f, the step
immediately before the message-passing block): two ``nn.EmbeddingBag(mode="sum")`` tables
(``node.weight`` / ``edge.weight``, same state-dict keys) looked up with the graph's integer type
features ``[V, t_v]`` / ``[E, t_e]``; computed by ``nt_embedding_bag_sum`` with a deterministic
hand-written backward. With it the on-the-wire input of a training step is the integer type ids, not
``[V, d]`` / ``[E, d]`` floats.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch import Tensor

from ... import _lib, ops
from ...data.models.graph import Graph, PendingFeats

# notorch/transforms/conf.py:35-44 (needs RDKit to import, so restated): 45 atom types, 13 bond types
DEFAULT_NUM_ATOM_TYPES = 45
DEFAULT_NUM_BOND_TYPES = 13
DEFAULT_HIDDEN_DIM = 256  # notorch/conf.py:11


def _embedding_bag_sum_raw(table: Tensor, idx: Tensor) -> Tensor:
    table = ops._require_float(table, "embedding table")
    idx = ops._require(idx, "type indices", torch.int64, 2)
    n, bag = idx.shape
    T, d = table.shape
    with torch.cuda.device(table.device):
        out = torch.empty((n, d), dtype=table.dtype, device=table.device)
        status = torch.zeros(1, dtype=torch.int32, device=table.device)
        ops._run("emb:nt_embedding_bag_sum", _lib.lib().nt_embedding_bag_sum, ops._p(table), T, ops._p(idx), n, bag, d, ops._p(out), ops._p(status),
                 _lib.NT_F32, ops._stream())
    ops._check_status(status, "GraphEmbedding type indices")
    return out


def _embedding_bag_backward_raw(g: Tensor, idx: Tensor, T: int) -> Tensor:
    n, bag = idx.shape
    d = g.shape[1]
    L = _lib.lib()
    with torch.cuda.device(g.device):
        gt = torch.empty((T, d), dtype=g.dtype, device=g.device)
        ws = ops._workspace(g.device, L.nt_embedding_bag_backward_workspace_bytes(n, T, d), slot=2)
        ops._run("embbwd:nt_embedding_bag_backward", L.nt_embedding_bag_backward, ops._p(g), ops._p(idx), n, bag, T, d, ops._p(gt), ops._p(ws),
                 ws.numel(), _lib.NT_F32, ops._stream())
    return gt


class _EmbeddingBagSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table: Tensor, idx: Tensor):
        out = _embedding_bag_sum_raw(table, idx)
        ctx.save_for_backward(idx.contiguous())
        ctx.shape = tuple(table.shape)
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        (idx,) = ctx.saved_tensors
        return _embedding_bag_backward_raw(g.contiguous(), idx, ctx.shape[0]), None


def embedding_bag_sum(table: Tensor, idx: Tensor) -> Tensor:
    if ops._via_ops(table, idx):
        ops._torch_ops()
        return torch.ops.notorch_b200.embedding_bag_sum(table, idx)
    return _EmbeddingBagSum.apply(table, idx)


class GraphEmbedding(nn.Module):
    def __init__(self, num_node_types: int = DEFAULT_NUM_ATOM_TYPES, num_edge_types: int = DEFAULT_NUM_BOND_TYPES,
                 hidden_dim: int = DEFAULT_HIDDEN_DIM):
        super().__init__()
        self.node = nn.EmbeddingBag(num_node_types, hidden_dim, mode="sum")
        self.edge = nn.EmbeddingBag(num_edge_types, hidden_dim, mode="sum")

    def forward(self, G: Graph) -> Graph:
        node_types, edge_types = G.node_feats, G.edge_feats
        wv, we = self.node.weight, self.edge.weight
        if isinstance(G, Graph) and ops.embed_edge_init_supported(wv, we, node_types, edge_types):
            # deferred: ChempropBlock fuses both look-ups into its edge initialisation (nt_embed_edge_init) and never needs
            # x_v / x_e; any other reader of G.node_feats / G.edge_feats triggers the unfused kernel below (same bits)
            call = object()  # the two placeholders of ONE forward call recognise each other by this token
            d, dev = wv.shape[1], wv.device
            return G.update(
                node_feats=PendingFeats(lambda: embedding_bag_sum(wv, node_types), (node_types.shape[0], d), wv.dtype, dev, (call, wv, node_types)),
                edge_feats=PendingFeats(lambda: embedding_bag_sum(we, edge_types), (edge_types.shape[0], d), we.dtype, dev, (call, we, edge_types)))
        return G.update(node_feats=embedding_bag_sum(wv, node_types), edge_feats=embedding_bag_sum(we, edge_types))

    @property
    def num_node_types(self) -> int:
        return self.node.num_embeddings

    @property
    def num_edge_types(self) -> int:
        return self.edge.num_embeddings

    @classmethod
    def from_transform(cls, transform, **kwargs) -> "GraphEmbedding":
        return cls(transform.num_node_types, transform.num_edge_types, **kwargs)
