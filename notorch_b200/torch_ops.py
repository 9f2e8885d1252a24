"""``torch.library`` registration of the hot-path kernels (SURVEY.md §8b "Torch registration").

The ``nn`` modules call the kernels through ``torch.autograd.Function`` wrappers in ``ops.py``; this module additionally exposes the same
launchers as dispatcher ops, ``torch.ops.notorch_b200.*``, with

* ``register_fake`` shape / dtype inference, so that meta / fake tensors (``torch.compile`` tracing, ``FakeTensorMode``) work without a
  GPU and without running a kernel, and
* ``register_autograd`` formulas that launch the hand-written backward kernels.

Every op is registered for CUDA only: calling one with CPU tensors raises (no CPU fallback, by design). Arguments are plain tensors
and scalars — the int32 CSR bundle of a graph (``ops.GraphCSR``) is passed as its individual tensors.

    seg_reduce(x, rowptr, perm?, keys, num_segments, mean, scale) -> [S, d]      K1 (no activation) / K3 / K5; chemprop.py:86, agg.py:27,36
    gather_add(base?, x, idx, mean_rowptr?, scale) -> [n, d]                     K0 and the backward of the reductions; chemprop.py:83
    chemprop_layer(h, W, b?, <9 CSR tensors>, V, act, act_param, mean, residual, p, seed, offset, gemm_mode) -> (h', saved)
                                                                                   K1 + K2 of one depth; chemprop.py:36-41, residual.py:28
    chemprop_layer_backward(g, h, saved, W, <9 CSR tensors>, ...) -> (g_h, g_W, g_b)   K4b, K4a, K5 + K6
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import ops
from ._lib import GEMM_FP32

__all__ = ["seg_reduce", "gather_add", "chemprop_layer", "layer_from_csr"]


def _seg(rowptr: Tensor, perm: Optional[Tensor], keys: Tensor, n: int) -> ops.SegmentCSR:
    return ops.SegmentCSR(rowptr, perm, keys, n)


def _graph(V: int, src: Tensor, dst: Tensor, rev: Tensor, dst_rowptr: Tensor, dst_perm: Tensor, src_rowptr: Tensor, src_perm: Tensor,
           rev_rowptr: Tensor, rev_perm: Tensor) -> ops.GraphCSR:
    E = src.numel()
    return ops.GraphCSR(V, E, _seg(dst_rowptr, dst_perm, dst, V), _seg(src_rowptr, src_perm, src, V), _seg(rev_rowptr, rev_perm, rev, E))


# ------------------------------------------------------------------------------------------------ seg_reduce
@torch.library.custom_op("notorch_b200::seg_reduce", mutates_args=(), device_types="cuda")
def seg_reduce(x: Tensor, rowptr: Tensor, perm: Optional[Tensor], keys: Tensor, num_segments: int, mean: bool, scale: float) -> Tensor:
    return ops._seg_reduce_raw(ops._require_float(x, "x"), _seg(rowptr, perm, keys, num_segments), 0, 0.0, mean, scale)


@seg_reduce.register_fake
def _(x, rowptr, perm, keys, num_segments, mean, scale):
    return x.new_empty((num_segments, x.shape[1]))


def _seg_reduce_setup(ctx, inputs, output):
    _, rowptr, _, keys, _, mean, scale = inputs
    ctx.save_for_backward(rowptr, keys)
    ctx.mean, ctx.scale = mean, scale


def _seg_reduce_backward(ctx, g):
    rowptr, keys = ctx.saved_tensors
    gx = torch.ops.notorch_b200.gather_add(None, g.contiguous(), keys, rowptr if ctx.mean else None, ctx.scale)
    return gx, None, None, None, None, None, None


seg_reduce.register_autograd(_seg_reduce_backward, setup_context=_seg_reduce_setup)


# ------------------------------------------------------------------------------------------------ gather_add
@torch.library.custom_op("notorch_b200::gather_add", mutates_args=(), device_types="cuda")
def gather_add(base: Optional[Tensor], x: Tensor, idx: Tensor, mean_rowptr: Optional[Tensor], scale: float) -> Tensor:
    return ops._gather_add_raw(base, ops._require_float(x, "x"), idx, mean_rowptr, scale)


@gather_add.register_fake
def _(base, x, idx, mean_rowptr, scale):
    return x.new_empty((idx.shape[0], x.shape[1]))


# ------------------------------------------------------------------------------------------------ one message-passing depth
@torch.library.custom_op("notorch_b200::chemprop_layer", mutates_args=(), device_types="cuda")
def chemprop_layer(h: Tensor, W: Tensor, b: Optional[Tensor], src: Tensor, dst: Tensor, rev: Tensor, dst_rowptr: Tensor, dst_perm: Tensor,
                   src_rowptr: Tensor, src_perm: Tensor, rev_rowptr: Tensor, rev_perm: Tensor, V: int, act: int, act_param: float,
                   mean: bool, residual: bool, p: float, seed: int, offset: int, gemm_mode: int) -> tuple[Tensor, Tensor]:
    """Returns (h', saved): ``saved`` is the message tensor m (tensor-core path) or n (strict-fp32 path), what the backward needs."""
    csr = _graph(V, src, dst, rev, dst_rowptr, dst_perm, src_rowptr, src_perm, rev_rowptr, rev_perm)
    h, W = ops._require_float(h, "edge_feats"), ops._require_float(W, "weight")
    save_m = gemm_mode != GEMM_FP32 and h.shape[1] % 4 == 0
    out, m, n, _ = ops._layer_forward_raw(h, W, b, csr, act, act_param, mean, residual, p, seed, offset, gemm_mode, save_m)
    return out, (m if save_m else n)


@chemprop_layer.register_fake
def _(h, W, b, src, dst, rev, dst_rowptr, dst_perm, src_rowptr, src_perm, rev_rowptr, rev_perm, V, act, act_param, mean, residual, p, seed,
      offset, gemm_mode):
    saved_rows = h.shape[0] if (gemm_mode != GEMM_FP32 and h.shape[1] % 4 == 0) else V
    return torch.empty_like(h), h.new_empty((saved_rows, h.shape[1]))


@torch.library.custom_op("notorch_b200::chemprop_layer_backward", mutates_args=(), device_types="cuda")
def chemprop_layer_backward(g: Tensor, h: Tensor, saved: Tensor, W: Tensor, has_bias: bool, src: Tensor, dst: Tensor, rev: Tensor,
                            dst_rowptr: Tensor, dst_perm: Tensor, src_rowptr: Tensor, src_perm: Tensor, rev_rowptr: Tensor, rev_perm: Tensor,
                            V: int, act: int, act_param: float, mean: bool, residual: bool, p: float, seed: int, offset: int,
                            gemm_mode: int) -> tuple[Tensor, Tensor, Tensor]:
    csr = _graph(V, src, dst, rev, dst_rowptr, dst_perm, src_rowptr, src_perm, rev_rowptr, rev_perm)
    has_m = gemm_mode != GEMM_FP32 and h.shape[1] % 4 == 0
    gh, gW, gb = ops._layer_backward_raw(g, h, saved if has_m else None, None if has_m else saved, W, has_bias, csr, act, act_param, mean,
                                         residual, p, seed, offset, gemm_mode, True, True)
    return gh, gW, gb if gb is not None else W.new_zeros(W.shape[0])


@chemprop_layer_backward.register_fake
def _(g, h, saved, W, has_bias, *rest):
    return torch.empty_like(h), torch.empty_like(W), W.new_empty(W.shape[0])


def _layer_setup(ctx, inputs, output):
    h, W, b = inputs[0], inputs[1], inputs[2]
    ctx.save_for_backward(h, output[1], W, *inputs[3:12])
    ctx.has_bias = b is not None
    ctx.cfg = tuple(inputs[12:])


def _layer_backward(ctx, g_out, g_saved):
    h, saved, W, *csr_tensors = ctx.saved_tensors
    gh, gW, gb = torch.ops.notorch_b200.chemprop_layer_backward(g_out.contiguous(), h, saved, W, ctx.has_bias, *csr_tensors, *ctx.cfg)
    return (gh, gW, gb if ctx.has_bias else None) + (None,) * 18


chemprop_layer.register_autograd(_layer_backward, setup_context=_layer_setup)


def layer_from_csr(h: Tensor, W: Tensor, b: Optional[Tensor], csr: ops.GraphCSR, act: int = 1, act_param: float = 0.0, mean: bool = False,
                   residual: bool = True, p: float = 0.0, seed: int = 0, offset: int = 0, gemm_mode: int = 0) -> Tensor:
    """Convenience wrapper: ``torch.ops.notorch_b200.chemprop_layer`` with the CSR bundle unpacked."""
    out, _ = torch.ops.notorch_b200.chemprop_layer(h, W, b, csr.src, csr.dst, csr.rev, csr.by_dst.rowptr, csr.by_dst.perm, csr.by_src.rowptr,
                                                   csr.by_src.perm, csr.by_rev.rowptr, csr.by_rev.perm, csr.V, act, act_param, mean, residual,
                                                   p, seed, offset, gemm_mode)
    return out
