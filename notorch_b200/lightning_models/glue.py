"""Glue between the drop-in modules and the reference's Lightning shell.

The reference's ``NotorchModel`` (``notorch/lightning_models/model.py:145-222``) is a ``TensorDictSequential`` of
``TensorDictModule``-wrapped ``nn.Module``s; every wrapped module is called with the values of its ``in_keys`` pulled out of the
batch TensorDict — positionally, or by keyword when ``in_keys`` is a mapping (``model.py:160-166``; ``SDPAttention.forward(G, *, Q)``,
``agg.py:72-78``, needs the keyword form). The batch itself is a TensorDict whose ``"inputs.G"`` entry is the ``BatchedGraph``
*object* (``data/dataset.py:65-66``), moved to the GPU by Lightning's ``transfer_batch_to_device`` calling ``.to(device)`` on it
(``data/models/graph.py:229-239``).

Two things a maintainer needs from this side, both here, neither needing ``lightning`` or ``tensordict`` to be installed:

* :func:`transfer_batch_to_device` — a body for ``LightningModule.transfer_batch_to_device``: moves tensors and graph objects of an
  arbitrarily nested batch and, for graphs, **keeps or rebuilds the per-batch CSR bundle at transfer time** (``prepare_graph``), so
  that the index preprocessing of batch t+1 is issued while batch t still computes and is never thrown away by ``.to()``.
* :class:`TensorDictModuleLite` / :class:`TensorDictSequentialLite` — the calling convention of ``tensordict.nn`` restated in 40
  lines over a plain ``dict`` (test double: ``tests/test_lightning_glue.py`` drives GraphEmbedding -> ChempropBlock -> read-out ->
  MLP through it positionally and by keyword, the way ``NotorchModel.forward`` would).
"""
from __future__ import annotations

from collections.abc import Mapping, Sequence
from typing import Any

import torch
import torch.nn as nn
from torch import Tensor

from ..data.models.graph import Graph

__all__ = ["transfer_batch_to_device", "prepare_graph", "TensorDictModuleLite", "TensorDictSequentialLite"]


def prepare_graph(G, num_segments: int | None = None):
    """Build (or re-key) everything the kernels cache on a graph object — the int32 CSR bundle by dst / src / rev and, for a
    batched graph, the read-out CSR — now, on the current stream, instead of inside the first module that needs it. A no-op for
    CPU graphs and for graphs whose caches are current. Returns ``G``."""
    from .. import ops

    ei = getattr(G, "edge_index", None)
    if not isinstance(ei, Tensor) or not ei.is_cuda or ei.dim() != 2:
        return G
    ops.graph_csr(G)
    bni = getattr(G, "batch_node_index", None)
    if isinstance(bni, Tensor) and getattr(G, "_nt_mol_ptr", None) is None:
        ops.segment_csr_for(G, "batch_node_index", len(G) if num_segments is None else num_segments)
    return G


def transfer_batch_to_device(batch: Any, device, dataloader_idx: int = 0, *, prepare: bool = True, non_blocking: bool = True) -> Any:
    """Drop-in body for ``LightningModule.transfer_batch_to_device(self, batch, device, dataloader_idx)``.

    Walks dicts / TensorDict-likes (anything with ``items()`` and item assignment), lists and tuples; tensors go through
    ``.to(device, non_blocking=True)``; graph objects (ours or the reference's: anything with ``edge_index`` and ``to``) through
    their own ``.to(device)`` — which for :class:`notorch_b200.Graph` carries the cached CSR bundle along — followed by
    :func:`prepare_graph`. Everything else is returned untouched."""
    del dataloader_idx
    dev = torch.device(device) if not isinstance(device, torch.device) else device

    def move(x):
        if isinstance(x, Tensor):
            return x.to(dev, non_blocking=non_blocking)
        if isinstance(x, Graph) or (hasattr(x, "edge_index") and hasattr(x, "rev_index") and callable(getattr(x, "to", None))):
            moved = x.to(dev)
            moved = x if moved is None else moved
            return prepare_graph(moved) if prepare and dev.type == "cuda" else moved
        if isinstance(x, Mapping):
            try:
                out = type(x)()
            except Exception:
                out = {}
            for k, v in x.items():
                out[k] = move(v)
            return out
        if hasattr(x, "items") and hasattr(x, "__setitem__") and not isinstance(x, (str, bytes)):  # TensorDict-like
            for k, v in list(x.items()):
                x[k] = move(v)
            return x
        if isinstance(x, tuple) and hasattr(x, "_fields"):
            return type(x)(*(move(v) for v in x))
        if isinstance(x, Sequence) and not isinstance(x, (str, bytes)):
            return type(x)(move(v) for v in x)
        return x

    return move(batch)


class TensorDictModuleLite(nn.Module):
    """``tensordict.nn.TensorDictModule(module, in_keys, out_keys)`` over a plain dict: ``in_keys`` a list -> positional call,
    a mapping ``{batch key: keyword}`` -> keyword call (model.py:160-166). Outputs are written under ``out_keys``."""

    def __init__(self, module: nn.Module, in_keys, out_keys: Sequence[str]):
        super().__init__()
        self.module, self.in_keys, self.out_keys = module, in_keys, list(out_keys)

    def forward(self, td: dict) -> dict:
        if isinstance(self.in_keys, Mapping):
            out = self.module(**{kw: td[key] for key, kw in self.in_keys.items()})
        else:
            out = self.module(*[td[key] for key in self.in_keys])
        outs = out if isinstance(out, tuple) else (out,)
        if len(outs) != len(self.out_keys):
            raise RuntimeError(f"{type(self.module).__name__} returned {len(outs)} values for out_keys {self.out_keys}")
        td = dict(td)
        td.update(zip(self.out_keys, outs))
        return td


class TensorDictSequentialLite(nn.Module):
    """``tensordict.nn.TensorDictSequential(*modules, selected_out_keys=...)`` over a plain dict (model.py:212,221-222)."""

    def __init__(self, *modules: TensorDictModuleLite, selected_out_keys: Sequence[str] | None = None):
        super().__init__()
        self.steps = nn.ModuleList(modules)
        self.selected_out_keys = None if selected_out_keys is None else list(selected_out_keys)

    def forward(self, td: dict) -> dict:
        for step in self.steps:
            td = step(td)
        if self.selected_out_keys is not None:
            keep = set(self.selected_out_keys)
            produced = {k for s in self.steps for k in s.out_keys}
            td = {k: v for k, v in td.items() if k in keep or k not in produced}
        return td
