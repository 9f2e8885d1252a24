"""notorch_b200 — B200-native (sm_100a) implementation of notorch's D-MPNN message-passing hot path.

Drop-in surface (same names as the reference):
``notorch_b200.nn.gnn.chemprop.{ChempropLayer, ChempropBlock}``, ``notorch_b200.nn.residual.Residual``,
``notorch_b200.nn.gnn.agg.{Sum, Mean}``, ``notorch_b200.data.models.graph.{Graph, BatchedGraph}``.
"""
from . import synth
from .data.models.graph import BatchedGraph, Graph

__all__ = ["Graph", "BatchedGraph", "synth", "__version__"]
__version__ = "0.1.0"
