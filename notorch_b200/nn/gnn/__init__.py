from . import agg
from .chemprop import ChempropBlock, ChempropLayer

__all__ = ["agg", "ChempropLayer", "ChempropBlock"]
