from . import agg
from .atom import AtomMessagePassing, AtomMessagePassingLayer
from .chemprop import ChempropBlock, ChempropLayer

__all__ = ["agg", "ChempropLayer", "ChempropBlock", "AtomMessagePassing", "AtomMessagePassingLayer"]
