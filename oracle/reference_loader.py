"""TEST INFRASTRUCTURE ONLY — imports the *unmodified* reference (davidegraff/notorch) from
``/root/reference`` so that the oracle restatement and the golden fixtures can be pinned against
the reference's own code executed on CPU.

``/root/reference`` exists only in the authoring container; on the GPU box ``available()`` is
False and everything that needs the live reference is skipped — the committed fixtures under
``tests/golden/`` (made by ``oracle/make_golden.py``) stand in for it.

Two shims go on ``sys.path`` (``oracle/ref_shims``): ``torch_scatter`` (absent third-party wheel,
restated as upstream composes it) and a stub ``rdkit.Chem.Mol``. No reference source is copied.
"""
from __future__ import annotations

import importlib
import os
import sys
from types import SimpleNamespace

REFERENCE_ROOT = os.environ.get("NOTORCH_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shims")
_cached: SimpleNamespace | None = None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "notorch", "nn", "gnn", "chemprop.py"))


def load() -> SimpleNamespace:
    """Return the reference's hot-path symbols. Raises ``ImportError`` if the tree is absent."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise ImportError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (_SHIMS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    chemprop = importlib.import_module("notorch.nn.gnn.chemprop")
    agg = importlib.import_module("notorch.nn.gnn.agg")
    graph = importlib.import_module("notorch.data.models.graph")
    residual = importlib.import_module("notorch.nn.residual")
    _cached = SimpleNamespace(
        ChempropLayer=chemprop.ChempropLayer,
        ChempropBlock=chemprop.ChempropBlock,
        agg=agg,
        Sum=agg.Sum,
        Mean=agg.Mean,
        Graph=graph.Graph,
        BatchedGraph=graph.BatchedGraph,
        Residual=residual.Residual,
    )
    return _cached
