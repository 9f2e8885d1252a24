from .graph import BatchedGraph, Graph

__all__ = ["Graph", "BatchedGraph"]
