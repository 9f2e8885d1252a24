from .gnn.agg import Aggregation, Gated, Max, Mean, Norm, SDPAttention, Sum
from .gnn.atom import AtomMessagePassing, AtomMessagePassingLayer
from .gnn.chemprop import ChempropBlock, ChempropLayer
from .gnn.embed import GraphEmbedding
from .mlp import MLP, Linear
from .residual import Residual

__all__ = ["GraphEmbedding", "ChempropLayer", "ChempropBlock", "AtomMessagePassing", "AtomMessagePassingLayer", "Residual", "MLP", "Linear", "Aggregation", "Sum", "Mean", "Norm", "Max", "Gated", "SDPAttention"]
