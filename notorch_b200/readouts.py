"""Autograd primitives behind the N2 read-outs (``Max``, ``Gated``, ``SDPAttention`` of
``notorch/nn/gnn/agg.py:41-86``): thin wrappers over the kernels in ``csrc/readout_kernels.cu`` with
hand-written backward passes. Everything is deterministic (no atomics)."""
from __future__ import annotations

import torch
from torch import Tensor

from . import _lib, ops
from ._lib import NT_F32
from .ops import SegmentCSR, _p, _run, _stream


def _vec(t: Tensor, name: str) -> Tensor:
    return ops._require(t, name, torch.float32, 1)


class _SegMax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, csr: SegmentCSR):
        x = ops._require_float(x, "node_feats")
        n, d = x.shape
        with torch.cuda.device(x.device):
            out = torch.empty((csr.num_segments, d), dtype=x.dtype, device=x.device)
            arg = torch.empty((csr.num_segments, d), dtype=torch.int32, device=x.device)
            _run("K3max:nt_seg_max", _lib.lib().nt_seg_max, _p(x), d, _p(csr.rowptr), _p(csr.perm), csr.num_segments, _p(out), _p(arg), NT_F32, _stream())
        ctx.save_for_backward(arg)
        ctx.csr, ctx.n = csr, n
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    def backward(ctx, g: Tensor, _garg):
        (arg,) = ctx.saved_tensors
        g = g.contiguous()
        d = g.shape[1]
        with torch.cuda.device(g.device):
            gx = torch.empty((ctx.n, d), dtype=g.dtype, device=g.device)
            _run("K3maxbwd:nt_seg_max_backward", _lib.lib().nt_seg_max_backward, _p(g), _p(arg), _p(ctx.csr.keys32), ctx.n, d, _p(gx), NT_F32, _stream())
        return gx, None


def _row_dot_raw(x: Tensor, y: Tensor, y_index: Tensor | None, scale: float, bias: float) -> Tensor:
    n, d = x.shape
    out = torch.empty(n, dtype=x.dtype, device=x.device)
    _run("rowdot:nt_row_dot", _lib.lib().nt_row_dot, _p(x), _p(y), _p(y_index), y.shape[0], n, d, scale, bias, _p(out), NT_F32, _stream())
    return out


def _row_scale_gather_raw(y: Tensor, w: Tensor, y_index: Tensor | None, n: int, scale: float) -> Tensor:
    d = y.shape[1]
    out = torch.empty((n, d), dtype=y.dtype, device=y.device)
    _run("rowscale:nt_row_scale_gather", _lib.lib().nt_row_scale_gather, _p(y), _p(w), _p(y_index), n, d, scale, _p(out), NT_F32, _stream())
    return out


def _seg_weighted_sum_raw(x: Tensor, w: Tensor, csr: SegmentCSR, scale: float) -> Tensor:
    d = x.shape[1]
    out = torch.empty((csr.num_segments, d), dtype=x.dtype, device=x.device)
    _run("wsum:nt_seg_weighted_sum", _lib.lib().nt_seg_weighted_sum, _p(x), _p(w), d, _p(csr.rowptr), _p(csr.perm), csr.num_segments, scale, _p(out),
         NT_F32, _stream())
    return out


class _RowDotSegment(torch.autograd.Function):
    """s[i] = scale * <x[i], y[seg(i)]>   (y: one row per segment; SDPAttention's Q[batch] . x, agg.py:80)"""

    @staticmethod
    def forward(ctx, x: Tensor, y: Tensor, csr: SegmentCSR, scale: float):
        x, y = ops._require_float(x, "x"), ops._require_float(y, "Q")
        if y.shape != (csr.num_segments, x.shape[1]):
            raise RuntimeError(f"notorch_b200: Q must have shape {(csr.num_segments, x.shape[1])}, got {tuple(y.shape)}")
        with torch.cuda.device(x.device):
            s = _row_dot_raw(x, y, csr.keys32, scale, 0.0)
        ctx.save_for_backward(x, y)
        ctx.csr, ctx.scale = csr, scale
        return s

    @staticmethod
    def backward(ctx, gs: Tensor):
        x, y = ctx.saved_tensors
        gs = gs.contiguous()
        with torch.cuda.device(gs.device):
            gx = _row_scale_gather_raw(y, gs, ctx.csr.keys32, x.shape[0], ctx.scale) if ctx.needs_input_grad[0] else None
            gy = _seg_weighted_sum_raw(x, gs, ctx.csr, ctx.scale) if ctx.needs_input_grad[1] else None
        return gx, gy, None, None


class _RowDotVector(torch.autograd.Function):
    """s[i] = <x[i], w> + b   (Gated's Linear(d, 1), agg.py:54,59)"""

    @staticmethod
    def forward(ctx, x: Tensor, w: Tensor, b: Tensor | None):
        x, w = ops._require_float(x, "x"), ops._require_float(w, "weight")
        if w.shape != (1, x.shape[1]):
            raise RuntimeError(f"notorch_b200: gate weight must have shape {(1, x.shape[1])}, got {tuple(w.shape)}")
        with torch.cuda.device(x.device):
            s = _row_dot_raw(x, w, None, 1.0, 0.0)
            if b is not None:
                s = s + b  # [1] bias added on the device (a one-element broadcast)
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return s

    @staticmethod
    def backward(ctx, gs: Tensor):
        x, w = ctx.saved_tensors
        gs = gs.contiguous()
        n, d = x.shape
        L = _lib.lib()
        gx = gw = gb = None
        with torch.cuda.device(gs.device):
            if ctx.needs_input_grad[0]:
                gx = _row_scale_gather_raw(w, gs, None, n, 1.0)
            if ctx.needs_input_grad[1]:
                gw = torch.empty((1, d), dtype=x.dtype, device=x.device)
                ws = ops._workspace(x.device, L.nt_weighted_col_sum_workspace_bytes(n, d), slot=3)
                _run("wcolsum:nt_weighted_col_sum", L.nt_weighted_col_sum, _p(x), _p(gs), n, d, 1.0, _p(gw), _p(ws), ws.numel(), NT_F32, _stream())
            if ctx.has_bias and ctx.needs_input_grad[2]:
                gb = torch.empty(1, dtype=x.dtype, device=x.device)
                ws = ops._workspace(x.device, L.nt_weighted_col_sum_workspace_bytes(n, 1), slot=3)
                _run("wcolsum:nt_weighted_col_sum", L.nt_weighted_col_sum, _p(gs), None, n, 1, 1.0, _p(gb), _p(ws), ws.numel(), NT_F32, _stream())
        return gx, gw, gb


class _SegSoftmax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, s: Tensor, csr: SegmentCSR):
        s = _vec(s, "scores")
        with torch.cuda.device(s.device):
            alpha = torch.empty_like(s)
            _run("softmax:nt_seg_softmax", _lib.lib().nt_seg_softmax, _p(s), _p(csr.rowptr), _p(csr.perm), csr.num_segments, _p(alpha), NT_F32, _stream())
        ctx.save_for_backward(alpha)
        ctx.csr = csr
        return alpha

    @staticmethod
    def backward(ctx, ga: Tensor):
        (alpha,) = ctx.saved_tensors
        ga = ga.contiguous()
        csr = ctx.csr
        with torch.cuda.device(ga.device):
            gs = torch.empty_like(alpha)
            _run("softmaxbwd:nt_seg_softmax_backward", _lib.lib().nt_seg_softmax_backward, _p(alpha), _p(ga), _p(csr.rowptr), _p(csr.perm),
                 csr.num_segments, _p(gs), NT_F32, _stream())
        return gs, None


class _SegWeightedSum(torch.autograd.Function):
    """H[b] = sum_{v in b} alpha[v] * x[v]   (scatter_sum(alpha * x), agg.py:61,84)"""

    @staticmethod
    def forward(ctx, x: Tensor, alpha: Tensor, csr: SegmentCSR):
        x, alpha = ops._require_float(x, "x"), _vec(alpha, "alpha")
        with torch.cuda.device(x.device):
            out = _seg_weighted_sum_raw(x, alpha, csr, 1.0)
        ctx.save_for_backward(x, alpha)
        ctx.csr = csr
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        x, alpha = ctx.saved_tensors
        g = g.contiguous()
        csr = ctx.csr
        with torch.cuda.device(g.device):
            gx = _row_scale_gather_raw(g, alpha, csr.keys32, x.shape[0], 1.0) if ctx.needs_input_grad[0] else None
            ga = _row_dot_raw(x, g, csr.keys32, 1.0, 0.0) if ctx.needs_input_grad[1] else None
        return gx, ga, None


def seg_max(x: Tensor, csr: SegmentCSR) -> tuple[Tensor, Tensor]:
    from . import ops

    if ops._via_ops(x):  # tracing / fake tensors / ops.set_dispatch("ops"): the registered op (nt_seg_max is nt_seg_extreme with identity, max)
        ops._torch_ops()
        return torch.ops.notorch_b200.seg_extreme(x, csr.rowptr, csr.perm, csr.keys32, csr.num_segments, False)
    return _SegMax.apply(x, csr)


def attention_readout(x: Tensor, scores: Tensor, csr: SegmentCSR) -> Tensor:
    return _SegWeightedSum.apply(x, _SegSoftmax.apply(scores, csr), csr)
