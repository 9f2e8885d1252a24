from .gnn.agg import Aggregation, Gated, Max, Mean, Norm, SDPAttention, Sum
from .gnn.chemprop import ChempropBlock, ChempropLayer
from .gnn.embed import GraphEmbedding
from .residual import Residual

__all__ = ["GraphEmbedding", "ChempropLayer", "ChempropBlock", "Residual", "Aggregation", "Sum", "Mean", "Norm", "Max", "Gated", "SDPAttention"]
