"""``Residual`` — same contract as ``notorch/nn/residual.py:21-28``: ``inputs[0] + module(*inputs)``.

Kept as a real module so reference checkpoints load with ``strict=True`` (keys
``layers.{i}.module.update.0.*``). ``ChempropBlock`` does not call this ``forward``: it fuses the
residual add into the layer kernel's epilogue (K2) and reads the wrapped layer's parameters.
Calling a ``Residual(ChempropLayer)`` on its own still works and gives the same result.
"""
from __future__ import annotations

import torch.nn as nn


class Residual(nn.Module):
    def __init__(self, module: nn.Module):
        super().__init__()
        self.module = module

    def forward(self, *inputs):
        from .gnn.chemprop import ChempropLayer

        if isinstance(self.module, ChempropLayer):
            return self.module(*inputs, _residual=True)  # residual add fused into the kernel epilogue
        return inputs[0] + self.module(*inputs)
