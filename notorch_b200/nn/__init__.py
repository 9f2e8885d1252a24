from .gnn.agg import Aggregation, Mean, Norm, Sum
from .gnn.chemprop import ChempropBlock, ChempropLayer
from .residual import Residual

__all__ = ["ChempropLayer", "ChempropBlock", "Residual", "Aggregation", "Sum", "Mean", "Norm"]
