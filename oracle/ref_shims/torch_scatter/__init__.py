"""TEST INFRASTRUCTURE ONLY — a stand-in for the third-party ``torch-scatter`` wheel so that the
*unmodified* reference modules under ``/root/reference`` can be imported and run on CPU.

``torch-scatter`` is a dependency of the reference (``pyproject.toml:35``, unpinned; the only
version evidence is the ``torch_scatter-2.1.2`` wheel URL at ``pyproject.toml:78``) and is absent
from this image. torch-scatter 2.1.2 composes its reductions as restated here (its published
``torch_scatter/scatter.py``):

* ``scatter_sum``  : broadcast ``index`` to ``src``'s shape, ``zeros(size).scatter_add_(dim, index, src)``
* ``scatter_mean`` : ``scatter_sum`` divided by ``scatter_sum(ones)`` with counts < 1 clamped to 1
  (``true_divide_`` for floating outputs — a division, not a reciprocal multiply)
* ``scatter_max``  : (values, argmax); empty segments give 0 / ``dim_size`` arg  (custom op upstream)
* ``scatter_softmax`` (``torch_scatter/composite/softmax.py``): max -> gather -> exp -> sum -> gather -> div

Reference call sites: ``notorch/nn/gnn/chemprop.py:6,39,86`` and ``notorch/nn/gnn/agg.py:9,27,36,45,60,83``.
"""
from __future__ import annotations

import torch
from torch import Tensor


def _broadcast(index: Tensor, src: Tensor, dim: int) -> Tensor:
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        for _ in range(dim):
            index = index.unsqueeze(0)
    for _ in range(index.dim(), src.dim()):
        index = index.unsqueeze(-1)
    return index.expand(src.size())


def _out_size(src: Tensor, index: Tensor, dim: int, dim_size: int | None) -> list[int]:
    size = list(src.size())
    if dim_size is not None:
        size[dim] = dim_size
    elif index.numel() == 0:
        size[dim] = 0
    else:
        size[dim] = int(index.max()) + 1
    return size


def scatter_sum(src: Tensor, index: Tensor, dim: int = -1, out: Tensor | None = None,
                dim_size: int | None = None) -> Tensor:
    index = _broadcast(index, src, dim)
    if out is None:
        out = torch.zeros(_out_size(src, index, dim, dim_size), dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, index, src)


scatter_add = scatter_sum


def scatter_mean(src: Tensor, index: Tensor, dim: int = -1, out: Tensor | None = None,
                 dim_size: int | None = None) -> Tensor:
    out = scatter_sum(src, index, dim, out, dim_size)
    dim_size = out.size(dim)
    index_dim = dim
    if index_dim < 0:
        index_dim = index_dim + src.dim()
    if index.dim() <= index_dim:
        index_dim = index.dim() - 1
    ones = torch.ones(index.size(), dtype=src.dtype, device=src.device)
    count = scatter_sum(ones, index, index_dim, None, dim_size)
    count[count < 1] = 1
    count = _broadcast(count, out, dim)
    if out.is_floating_point():
        out.true_divide_(count)
    else:
        out.div_(count, rounding_mode="floor")
    return out


def scatter_max(src: Tensor, index: Tensor, dim: int = -1, out: Tensor | None = None,
                dim_size: int | None = None) -> tuple[Tensor, Tensor]:
    """Upstream is a custom C++/CUDA op: values + ``arg`` (index of the maximum; ``src.size(dim)`` for an empty
    segment, whose value is 0), and its backward routes the gradient to the ``arg`` element only. On CPU the reducer
    updates on a strict ``>``, so the FIRST maximum wins on ties. Restated with that forward and that backward."""
    index_b = _broadcast(index, src, dim)
    size = _out_size(src, index_b, dim, dim_size)
    n = src.size(dim)
    with torch.no_grad():
        vals = torch.full(size, float("-inf"), dtype=src.dtype, device=src.device)
        vals = vals.scatter_reduce(dim, index_b, src, reduce="amax", include_self=True)
        pos_shape = [1] * src.dim()
        pos_shape[dim] = n
        pos = torch.arange(n, device=src.device).view(pos_shape).expand(src.size())
        hit = src == vals.gather(dim, index_b)
        cand = torch.where(hit, pos, torch.full_like(pos, n))
        arg = torch.full(size, n, dtype=torch.long, device=src.device)
        arg = arg.scatter_reduce(dim, index_b, cand, reduce="amin", include_self=True)
    empty = arg == n
    picked = src.gather(dim, arg.clamp(max=max(n - 1, 0))) if n > 0 else torch.zeros(size, dtype=src.dtype, device=src.device)
    vals = torch.where(empty, torch.zeros_like(picked), picked)  # gradient flows to the arg element only
    return vals, arg


def scatter_min(src, index, dim=-1, out=None, dim_size=None):
    v, a = scatter_max(-src, index, dim, out, dim_size)
    return -v, a


def scatter(src: Tensor, index: Tensor, dim: int = -1, out: Tensor | None = None,
            dim_size: int | None = None, reduce: str = "sum") -> Tensor:
    if reduce in ("sum", "add"):
        return scatter_sum(src, index, dim, out, dim_size)
    if reduce == "mean":
        return scatter_mean(src, index, dim, out, dim_size)
    if reduce == "max":
        return scatter_max(src, index, dim, out, dim_size)[0]
    if reduce == "min":
        return scatter_min(src, index, dim, out, dim_size)[0]
    raise ValueError(reduce)


def scatter_softmax(src: Tensor, index: Tensor, dim: int = -1, dim_size: int | None = None) -> Tensor:
    index = _broadcast(index, src, dim)
    max_per = scatter_max(src.detach(), index, dim, dim_size=dim_size)[0]
    recentered = src - max_per.gather(dim, index)
    ex = recentered.exp()
    denom = scatter_sum(ex, index, dim, dim_size=dim_size).gather(dim, index)
    return ex / denom
