"""TEST INFRASTRUCTURE ONLY — CPU restatement (oracle) of the reference's D-MPNN hot path.

This file is the *checker*, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it. The shipped path
(``notorch_b200``) never routes through it and has no CPU fallback.

Parity pinning: the reference's own tests hold **no** golden vectors for this path
("parity unpinned" by the reference's tests). The oracle is therefore pinned
against *outputs of the reference itself*: ``oracle/make_golden.py`` imports the unmodified
reference from ``/root/reference`` (with the two shims in ``oracle/ref_shims``), runs it on seeded
synthetic graphs and commits inputs + outputs under ``tests/golden/``;
``tests/test_oracle.py`` checks every function below against those fixtures (bit-exact for index
tensors and for the fp32 forward, which uses the same ATen ops in the same order) and, when
``/root/reference`` is present, against the live reference.

What each function follows (paths relative to ``/root/reference``):

* ``collate``        — ``notorch/data/models/graph.py:186-223`` (``BatchedGraph.from_graphs``),
                       including the node-offset ``rev_index`` (``:199-200``).
* ``edge_init``      — ``notorch/nn/gnn/chemprop.py:83``.
* ``seg_reduce``     — ``torch_scatter.scatter`` as called at ``chemprop.py:39,86`` and
                       ``nn/gnn/agg.py:27,36`` (sum = ``zeros.scatter_add_``; mean = sum / clamp(count, 1)).
* ``seg_extreme``    — ``torch_scatter.scatter_max`` / ``scatter_min`` (``reduce="max" | "min"`` at ``chemprop.py:39,86``,
                       ``agg.py:45``): published torch-scatter 2.1 semantics, pinned by ``tests/golden/{max,min}_reduce.npz`` and
                       ``tests/golden_readouts/readout_max.npz`` (reference run on the ``torch_scatter`` shim) and by a brute-force loop.
* ``layer_forward``  — ``notorch/nn/gnn/chemprop.py:28-43`` + ``notorch/nn/residual.py:27-28``.
* ``block_forward``  — ``notorch/nn/gnn/chemprop.py:81-88``.
* ``readout``        — ``notorch/nn/gnn/agg.py:23-38``.
* ``block_backward`` — hand-derived reverse of the above (SURVEY.md §8a "Backward"), checked
                       against autograd in ``tests/test_oracle.py``.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor

# --------------------------------------------------------------------------------------------
# integer work: collation and CSR (numpy)
# --------------------------------------------------------------------------------------------


def collate(mols: Sequence[tuple[int, np.ndarray, np.ndarray]]) -> dict[str, np.ndarray]:
    """``BatchedGraph.from_graphs`` on index tensors only (graph.py:186-223).

    ``mols[i] = (n_atoms, local edge_index [2, e_i], local rev_index [e_i])``. The running offset is
    the cumulative **atom** count and it is added to *both* ``edge_index`` and ``rev_index``
    (graph.py:199-200, 204) — the quirk SURVEY.md §0 item 3 documents; reproduced on purpose.
    """
    eis, revs, bni, bei = [], [], [], []
    offset = 0
    for i, (n, ei, rev) in enumerate(mols):
        eis.append(np.asarray(ei, dtype=np.int64).reshape(2, -1) + offset)
        revs.append(np.asarray(rev, dtype=np.int64) + offset)
        bni.append(np.full(n, i, dtype=np.int64))
        bei.append(np.full(len(rev), i, dtype=np.int64))
        offset += n
    return {
        "edge_index": np.concatenate(eis, axis=1) if eis else np.zeros((2, 0), np.int64),
        "rev_index": np.concatenate(revs) if revs else np.zeros((0,), np.int64),
        "batch_node_index": np.concatenate(bni) if bni else np.zeros((0,), np.int64),
        "batch_edge_index": np.concatenate(bei) if bei else np.zeros((0,), np.int64),
        "size": len(mols),
    }


def collate_fixed(mols: Sequence[tuple[int, np.ndarray, np.ndarray]]) -> dict[str, np.ndarray]:
    """Deviation, clearly labelled: same as :func:`collate` but ``rev_index`` gets the cumulative
    **edge** offset (the structurally correct reverse edge). Not reference behaviour."""
    out = collate(mols)
    revs, eoff = [], 0
    for _, _, rev in mols:
        revs.append(np.asarray(rev, dtype=np.int64) + eoff)
        eoff += len(rev)
    out["rev_index"] = np.concatenate(revs) if revs else np.zeros((0,), np.int64)
    return out


def build_csr(keys: np.ndarray, num_segments: int) -> tuple[np.ndarray, np.ndarray]:
    """Stable counting sort of item ids by ``keys``: ``(rowptr [S+1] int32, perm [n] int32)`` with
    ``perm[rowptr[s]:rowptr[s+1]]`` = the items whose key is ``s`` in ascending item id."""
    keys = np.asarray(keys, dtype=np.int64)
    perm = np.argsort(keys, kind="stable").astype(np.int32)
    counts = np.bincount(keys, minlength=num_segments)[:num_segments]
    rowptr = np.zeros(num_segments + 1, dtype=np.int32)
    np.cumsum(counts, out=rowptr[1:])
    return rowptr, perm


# --------------------------------------------------------------------------------------------
# floating-point work (torch, CPU; dtype follows the inputs: fp32 for parity, fp64 for gradcheck)
# --------------------------------------------------------------------------------------------

ACTS = {
    "relu": (F.relu, {}),
    "leaky_relu": (F.leaky_relu, {"negative_slope": 0.01}),
    "elu": (F.elu, {"alpha": 1.0}),
    "silu": (F.silu, {}),
    "gelu": (F.gelu, {}),
    "tanh": (torch.tanh, {}),
    "identity": (lambda x: x, {}),
}


def apply_act(x: Tensor, act: str = "relu", param: float | None = None) -> Tensor:
    fn, kw = ACTS[act]
    if param is not None and kw:
        kw = {next(iter(kw)): param}
    return fn(x, **kw)


def seg_extreme(x: Tensor, index: Tensor, size: int, reduce: str = "max") -> tuple[Tensor, Tensor]:
    """``torch_scatter.scatter_max`` / ``scatter_min`` along dim 0 (published semantics of torch-scatter 2.1: the FIRST row attaining
    the extreme is the argument, an empty segment yields value 0 and argument ``len(x)``; ``scatter(..., reduce="max")`` returns the
    value only and autograd routes the gradient to the argument row). Returns (value ``[size, d]``, arg ``[size, d]`` int64)."""
    n, d = x.shape
    sign = 1.0 if reduce == "max" else -1.0
    xs = x * sign  # exact
    idx = index.view(-1, 1).expand(n, d)
    best = torch.full((size, d), float("-inf"), dtype=x.dtype).scatter_reduce_(0, idx, xs, "amax", include_self=True)
    rows = torch.arange(n).view(-1, 1).expand(n, d)
    cand = torch.where(xs == best[index], rows, torch.full_like(rows, n))
    arg = torch.full((size, d), n, dtype=torch.long).scatter_reduce_(0, idx, cand, "amin", include_self=True)
    empty = arg == n
    return torch.where(empty, torch.zeros_like(best), best * sign), arg


def seg_reduce(x: Tensor, index: Tensor, size: int, reduce: str = "sum") -> Tensor:
    """``torch_scatter.scatter(x, index, dim=0, dim_size=size, reduce=...)`` for sum / mean / max / min."""
    if reduce in ("max", "min"):
        return seg_extreme(x, index, size, reduce)[0]
    idx = index.view(-1, 1).expand_as(x)
    out = torch.zeros((size, x.shape[1]), dtype=x.dtype, device=x.device).scatter_add_(0, idx, x)  # device: the port also runs as the eager-CUDA baseline
    if reduce == "sum":
        return out
    if reduce == "mean":
        count = torch.zeros(size, dtype=x.dtype, device=x.device).scatter_add_(0, index, torch.ones(len(index), dtype=x.dtype, device=x.device))
        count = count.clamp(min=1)
        return out / count.view(-1, 1)
    raise NotImplementedError(reduce)


def edge_init(x_v: Tensor, x_e: Tensor, src: Tensor) -> Tensor:
    """``h_0 = x_v[src] + x_e`` (chemprop.py:83) — no ``W_i``, no activation."""
    return x_v[src] + x_e


def layer_forward(
    h: Tensor,
    num_nodes: int,
    src: Tensor,
    dst: Tensor,
    rev: Tensor,
    weight: Tensor,
    bias: Tensor | None,
    *,
    act: str = "relu",
    act_param: float | None = None,
    reduce: str = "sum",
    residual: bool = True,
    keep_mask: Tensor | None = None,
    p: float = 0.0,
) -> tuple[Tensor, dict[str, Tensor]]:
    """One ``Residual(ChempropLayer)`` (chemprop.py:36-41, residual.py:28).

    ``keep_mask`` (bool ``[E, d]``) makes dropout reproducible: ``u = mask * u / (1 - p)``; the
    reference's Philox stream cannot be matched, so parity tests inject the mask.
    """
    a = apply_act(h, act, act_param)  # chemprop.py:37 (pre-activation)
    n = seg_reduce(a, dst, num_nodes, reduce)  # :39
    m = n[src] - a[rev]  # :40
    u = F.linear(m, weight, bias)  # :41 Linear
    if keep_mask is not None and p > 0.0:
        u = u * keep_mask.to(u.dtype) / (1.0 - p)  # :41 Dropout
    out = h + u if residual else u  # residual.py:28
    return out, {"a": a, "n": n, "m": m}


def block_forward(
    x_v: Tensor,
    x_e: Tensor,
    edge_index: Tensor,
    rev_index: Tensor,
    weights: Sequence[Tensor],
    biases: Sequence[Tensor | None],
    *,
    act: str = "relu",
    act_param: float | None = None,
    reduce: str = "sum",
    residual: bool = True,
    keep_masks: Sequence[Tensor] | None = None,
    p: float = 0.0,
) -> tuple[Tensor, Tensor, list[Tensor]]:
    """``ChempropBlock.forward`` (chemprop.py:81-88): returns ``(node_out, edge_out, [h_0..h_L])``."""
    src, dst = edge_index[0], edge_index[1]
    V = x_v.shape[0]
    h = edge_init(x_v, x_e, src)
    hs = [h]
    for l, (W, b) in enumerate(zip(weights, biases)):
        km = keep_masks[l] if keep_masks is not None else None
        h, _ = layer_forward(h, V, src, dst, rev_index, W, b, act=act, act_param=act_param,
                             reduce=reduce, residual=residual, keep_mask=km, p=p)
        hs.append(h)
    node_out = seg_reduce(h, dst, V, reduce)  # :86 — no final activation
    return node_out, h, hs


def readout(x: Tensor, batch_node_index: Tensor, size: int, kind: str = "sum", norm: float = 100.0) -> Tensor:
    """``agg.Sum`` / ``agg.Mean`` (agg.py:23-38); ``norm`` is the builder-defined extension of
    SURVEY.md §8a row A9 (sum / constant; parity unpinned — not in the reference tree)."""
    if kind == "sum":
        return seg_reduce(x, batch_node_index, size, "sum")
    if kind == "mean":
        return seg_reduce(x, batch_node_index, size, "mean")
    if kind == "norm":
        return seg_reduce(x, batch_node_index, size, "sum") / norm
    raise NotImplementedError(kind)


def readout_max(x: Tensor, batch_node_index: Tensor, size: int) -> Tensor:
    """``agg.Max`` (agg.py:41-47): ``torch_scatter.scatter_max`` values; empty molecules give 0."""
    idx = batch_node_index.view(-1, 1).expand_as(x)
    out = torch.full((size, x.shape[1]), float("-inf"), dtype=x.dtype).scatter_reduce(0, idx, x, reduce="amax", include_self=True)
    return torch.where(torch.isinf(out) & (out < 0), torch.zeros_like(out), out)


def _seg_softmax(scores: Tensor, index: Tensor, size: int) -> Tensor:
    """``torch_scatter.scatter_softmax`` over a vector (composite/softmax.py): max -> exp -> sum -> divide."""
    mx = torch.full((size,), float("-inf"), dtype=scores.dtype).scatter_reduce(0, index, scores, reduce="amax", include_self=True)
    ex = (scores - mx[index]).exp()
    return ex / torch.zeros(size, dtype=scores.dtype).scatter_add_(0, index, ex)[index]


def readout_gated(x: Tensor, weight: Tensor, bias: Tensor | None, batch_node_index: Tensor, size: int) -> Tensor:
    """INTENDED semantics of ``agg.Gated`` (agg.py:50-63): ``alpha = softmax_b(x w^T + b)`` of shape [V], ``H[b] = sum alpha x``.
    The reference itself mis-broadcasts (returns [b, V, d]); this restatement is the builder's, parity unpinned."""
    scores = torch.nn.functional.linear(x, weight, bias).squeeze(1)
    alpha = _seg_softmax(scores, batch_node_index, size)
    return seg_reduce(alpha.unsqueeze(1) * x, batch_node_index, size, "sum")


def readout_sdpa(x: Tensor, Q: Tensor, batch_node_index: Tensor, size: int, key_dim: int) -> Tensor:
    """INTENDED semantics of ``agg.SDPAttention`` (agg.py:66-86): ``scores = <Q[batch], x> / sqrt(key_dim)``. The reference
    raises on its einsum string; this restatement is the builder's, parity unpinned."""
    scores = (Q[batch_node_index] * x).sum(1) / key_dim ** 0.5
    alpha = _seg_softmax(scores, batch_node_index, size)
    return seg_reduce(alpha.unsqueeze(1) * x, batch_node_index, size, "sum")


def _act_grad(h: Tensor, act: str, act_param: float | None) -> Tensor:
    hh = h.detach().clone().requires_grad_(True)
    apply_act(hh, act, act_param).sum().backward()
    return hh.grad


def block_backward(
    x_v: Tensor,
    x_e: Tensor,
    edge_index: Tensor,
    rev_index: Tensor,
    weights: Sequence[Tensor],
    biases: Sequence[Tensor | None],
    g_node: Tensor | None,
    g_edge: Tensor | None,
    *,
    act: str = "relu",
    act_param: float | None = None,
    reduce: str = "sum",
    residual: bool = True,
    keep_masks: Sequence[Tensor] | None = None,
    p: float = 0.0,
    act_grad_at: Sequence[Tensor] | None = None,
) -> dict[str, object]:
    """Hand-derived backward of :func:`block_forward`. This is the arithmetic
    the CUDA backward kernels implement; autograd on :func:`block_forward` is the referee.

    ``act_grad_at`` (optional ``[h_0..h_{L-1}]``): evaluate ``act'`` at these states instead of the
    oracle's own. ReLU's derivative is discontinuous at 0, so two correct fp32 implementations can
    disagree on the sign of an element with ``|h| ~ 1e-7 * max|h|`` and then differ by O(1) in a few
    gradient entries; the parity tests pass the CUDA path's own ``h_l`` here when (and only when)
    such a sign flip is detected, so that the comparison is of the same piecewise-linear branch."""
    src, dst = edge_index[0], edge_index[1]
    V, E = x_v.shape[0], x_e.shape[0]
    dt = x_v.dtype
    _, _, hs = block_forward(x_v, x_e, edge_index, rev_index, weights, biases, act=act,
                             act_param=act_param, reduce=reduce, residual=residual,
                             keep_masks=keep_masks, p=p)
    indeg = torch.zeros(V, dtype=dt).index_add_(0, dst, torch.ones(E, dtype=dt)).clamp(min=1)
    g = torch.zeros((E, x_e.shape[1]), dtype=dt)
    eids = torch.arange(E).view(-1, 1)
    extreme = reduce in ("max", "min")

    def through_reduce(g_atoms: Tensor, reduced: Tensor) -> Tensor:
        """Gradient of ``scatter(reduced, dst, V, reduce)`` w.r.t. its [E, d] input, given the [V, d] cotangent."""
        if extreme:  # only the argument row of every (atom, channel) receives the gradient
            arg = seg_extreme(reduced, dst, V, reduce)[1]
            return torch.where(arg[dst] == eids, g_atoms[dst], torch.zeros((), dtype=dt))
        return (g_atoms / indeg.view(-1, 1) if reduce == "mean" else g_atoms)[dst]

    if g_edge is not None:
        g = g + g_edge
    if g_node is not None:
        g = g + through_reduce(g_node, hs[-1])
    gWs, gbs = [], []
    for l in reversed(range(len(weights))):
        W, h = weights[l], hs[l]
        a = apply_act(h, act, act_param)
        n = seg_reduce(a, dst, V, reduce)
        m = n[src] - a[rev_index]
        g_u = g
        if keep_masks is not None and p > 0.0:
            g_u = g * keep_masks[l].to(dt) / (1.0 - p)
        gWs.append(g_u.t() @ m)
        gbs.append(g_u.sum(0) if biases[l] is not None else None)
        g_m = g_u @ W
        g_n = torch.zeros((V, W.shape[1]), dtype=dt).index_add_(0, src, g_m)
        g_a = through_reduce(g_n, a) - torch.zeros_like(g_m).index_add_(0, rev_index, g_m)
        h_for_mask = act_grad_at[l].to(dt) if act_grad_at is not None else h
        g = (g if residual else 0) + _act_grad(h_for_mask, act, act_param) * g_a
    g_xv = torch.zeros_like(x_v).index_add_(0, src, g)
    return {"x_v": g_xv, "x_e": g, "weights": gWs[::-1], "biases": gbs[::-1]}


def readout_backward(g_out: Tensor, batch_node_index: Tensor, num_nodes: int, kind: str = "sum",
                     norm: float = 100.0) -> Tensor:
    B = g_out.shape[0]
    if kind == "sum":
        return g_out[batch_node_index]
    if kind == "norm":
        return g_out[batch_node_index] / norm
    if kind == "mean":
        cnt = torch.zeros(B, dtype=g_out.dtype).index_add_(
            0, batch_node_index, torch.ones(num_nodes, dtype=g_out.dtype)).clamp(min=1)
        return (g_out / cnt.view(-1, 1))[batch_node_index]
    raise NotImplementedError(kind)


# --------------------------------------------------------------------------------------------
# the timed CPU leg (bench.py cpu_baseline / --impl reference): same ATen op sequence the
# reference dispatches (index, add, relu, scatter_add_, index, sub, addmm, add; autograd backward)
# --------------------------------------------------------------------------------------------


class CpuPort(torch.nn.Module):
    """``ChempropBlock`` + ``agg.Sum``/``agg.Mean`` restated as an ``nn.Module`` with the
    reference's parameter names (``layers.{i}.module.update.0.{weight,bias}``), CPU only."""

    def __init__(self, hidden_dim: int = 300, depth: int = 3, bias: bool = True, residual: bool = True,
                 reduce: str = "sum", agg: str = "sum", act: str = "relu", embed: tuple[int, int] | None = None):
        super().__init__()
        # optional GraphEmbedding in front (notorch/nn/gnn/embed.py:20-24): then x_v / x_e are int64 type ids
        self.node = torch.nn.EmbeddingBag(embed[0], hidden_dim, mode="sum") if embed else None
        self.edge = torch.nn.EmbeddingBag(embed[1], hidden_dim, mode="sum") if embed else None
        self.hidden_dim, self.depth, self.residual, self.reduce, self.agg, self.act = (
            hidden_dim, depth, residual, reduce, agg, act)
        self.linears = torch.nn.ModuleList(torch.nn.Linear(hidden_dim, hidden_dim, bias) for _ in range(depth))

    def reference_state_dict(self) -> dict[str, Tensor]:
        mid = "module." if self.residual else ""
        out = {}
        for i, lin in enumerate(self.linears):
            out[f"layers.{i}.{mid}update.0.weight"] = lin.weight.detach().clone()
            if lin.bias is not None:
                out[f"layers.{i}.{mid}update.0.bias"] = lin.bias.detach().clone()
        return out

    def forward(self, x_v, x_e, edge_index, rev_index, batch_node_index, size):
        if self.node is not None:
            x_v, x_e = self.node(x_v), self.edge(x_e)  # embed.py:24
        node_out, edge_out, _ = block_forward(
            x_v, x_e, edge_index, rev_index,
            [l.weight for l in self.linears], [l.bias for l in self.linears],
            act=self.act, reduce=self.reduce, residual=self.residual)
        return readout(node_out, batch_node_index, size, self.agg), node_out, edge_out


def train_step_cpu(model: CpuPort, x_v, x_e, edge_index, rev_index, batch_node_index, size) -> float:
    """zero_grad -> forward -> ``loss = H.square().mean()`` -> backward (BASELINE.md §3 protocol)."""
    model.zero_grad(set_to_none=True)
    H, _, _ = model(x_v, x_e, edge_index, rev_index, batch_node_index, size)
    loss = H.square().mean()
    loss.backward()
    return float(loss.detach())
