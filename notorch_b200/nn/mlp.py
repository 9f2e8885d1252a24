"""Prediction head: ``MLP`` of ``notorch/nn/mlp.py:9-68`` with the ``nn.Linear`` layers on hand-written
strict-fp32 kernels (``nt_linear_forward`` / ``nt_linear_backward_input`` / ``nt_linear_backward_weight``).

The factory keeps the reference's module sequence — ``Linear, act, dropout, Linear, ..., Linear[, Unflatten]`` with ONE shared
activation and ONE shared dropout instance — so ``state_dict`` keys (``0.weight``, ``3.weight``, ...) and ``hydra`` constructor
kwargs are those of the reference. The activation / dropout modules between the layers are the stock torch modules: they act on
``[B, hidden]`` molecule vectors, three orders of magnitude smaller than the edge tensors of the encoder.
"""
from __future__ import annotations

from collections.abc import Sequence
from math import prod

import torch
import torch.nn as nn
from torch import Tensor

from .. import _lib, ops
from .._lib import NT_F32
from ..ops import _p, _run, _stream, _workspace

DEFAULT_HIDDEN_DIM = 256  # notorch/conf.py:11


def _linear_forward_raw(x: Tensor, W: Tensor, b: Tensor | None) -> Tensor:
    x, W = ops._require_float(x, "input"), ops._require_float(W, "weight")
    rows, k = x.shape
    n = W.shape[0]
    if W.shape[1] != k:
        raise RuntimeError(f"notorch_b200: Linear weight {tuple(W.shape)} does not match input features {k}")
    if b is not None:
        b = ops._require(b, "bias", torch.float32, 1)
    with torch.cuda.device(x.device):
        out = torch.empty((rows, n), dtype=x.dtype, device=x.device)
        _run("head:nt_linear_forward", _lib.lib().nt_linear_forward, _p(x), _p(W), _p(b), rows, n, k, _p(out), NT_F32, _stream())
    return out


def _linear_backward_raw(g: Tensor, x: Tensor, W: Tensor, has_bias: bool, need_x: bool, need_w: bool):
    rows, k = x.shape
    n = W.shape[0]
    L = _lib.lib()
    gx = gW = gb = None
    with torch.cuda.device(g.device):
        if need_x:
            gx = torch.empty_like(x)
            _run("head:nt_linear_backward_input", L.nt_linear_backward_input, _p(g), _p(W), rows, n, k, _p(gx), NT_F32, _stream())
        if need_w:
            gW = torch.empty_like(W)
            gb = torch.empty(n, dtype=W.dtype, device=W.device) if has_bias else None
            ws = _workspace(g.device, L.nt_linear_backward_weight_workspace_bytes(rows, n, k), slot=4)
            _run("head:nt_linear_backward_weight", L.nt_linear_backward_weight, _p(g), _p(x), rows, n, k, _p(gW), _p(gb), _p(ws), ws.numel(),
                 NT_F32, _stream())
    return gx, gW, gb


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, W: Tensor, b: Tensor | None):
        out = _linear_forward_raw(x, W, b)
        ctx.save_for_backward(x.contiguous(), W.contiguous())
        ctx.has_bias = b is not None
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        x, W = ctx.saved_tensors
        return _linear_backward_raw(g.contiguous(), x, W, ctx.has_bias, ctx.needs_input_grad[0],
                                    ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]))


class Linear(nn.Linear):
    """``nn.Linear`` (same parameters, init and ``state_dict``) whose forward / backward run on the library's fp32 kernels.
    CUDA float32 only; leading dimensions are flattened like ``F.linear`` does."""

    def forward(self, input: Tensor) -> Tensor:  # noqa: A002 (nn.Linear's own argument name)
        lead = input.shape[:-1]
        x = input.reshape(-1, input.shape[-1])
        if not x.is_contiguous():
            x = x.contiguous()
        if ops._via_ops(x, self.weight):
            ops._torch_ops()
            out = torch.ops.notorch_b200.linear(x, self.weight, self.bias)
        else:
            out = _LinearFn.apply(x, self.weight, self.bias)
        return out.reshape(*lead, self.out_features)


def MLP(input_dim: int, output_size: int | Sequence[int], hidden_dim: int = DEFAULT_HIDDEN_DIM, num_layers: int = 1, dropout: float = 0.0,
        activation: type[nn.Module] = nn.ReLU) -> nn.Sequential:
    """``h_0 = x W_0 + b_0``; ``h_l = dropout(act(h_{l-1})) W_l + b_l`` — ``num_layers`` hidden layers, then the output layer; a
    sequence ``output_size`` unflattens the last dimension (mlp.py:9-68)."""
    if isinstance(output_size, int):
        output_dim, unflatten = output_size, None
    else:
        output_dim, unflatten = prod(output_size), nn.Unflatten(-1, tuple(output_size))
    drop, act = nn.Dropout(dropout), activation()
    dims = [input_dim] + [hidden_dim] * num_layers + [output_dim]
    modules: list[nn.Module] = []
    for i, (d_in, d_out) in enumerate(zip(dims[:-1], dims[1:])):
        if i > 0:
            modules += [act, drop]
        modules.append(Linear(d_in, d_out))
    mlp = nn.Sequential(*modules)
    if unflatten is not None:
        mlp.append(unflatten)
    return mlp
