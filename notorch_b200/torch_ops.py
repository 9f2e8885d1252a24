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
    build_csr(keys, num_segments) -> (rowptr, perm, keys32, status)             K-l; graph.py:186-223 (index preprocessing)
    csr_to_ell(rowptr, perm, num_segments) -> [S, 4]
    collate(num_atoms, num_edges, local_edge_index, local_rev_index, V, E, fixed_rev) -> 6 index tensors   BatchedGraph.from_graphs
    embedding_bag_sum(table, idx) -> [n, d]                                      GraphEmbedding, embed.py:20-24
    embed_edge_init(table_v, table_e, node_types, edge_types, src, V) -> [E, d]  GraphEmbedding fused into K0 (row N1)
    seg_extreme(x, rowptr, perm, keys, num_segments, is_min) -> (out, arg)       scatter max / min, chemprop.py:39,86, agg.py:45
    linear(x, W, b?) -> [rows, out]                                              the MLP head's Linear, mlp.py:58-62
    atom_layer(h, s_e, W, b?, <5 neighbour-list tensors>, ...) -> (h', n)        atom-state message passing (extension A10)

``ops.set_dispatch("ops")`` (or ``NOTORCH_B200_DISPATCH=ops``) makes the ``nn`` modules call these instead of their
``autograd.Function`` twins; they do so on their own whenever they are being traced (``torch.compile``) or see fake tensors, so
``FakeTensorMode`` / ``torch.compile`` work on the modules themselves. The eager default stays on the ``autograd.Function`` path:
a ``custom_op`` call costs ~20 us of dispatcher time per launch, which a CUDA graph hides and an eager loop does not.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import ops
from ._lib import GEMM_FP32

__all__ = ["seg_reduce", "gather_add", "chemprop_layer", "layer_from_csr", "build_csr", "csr_to_ell", "collate", "embedding_bag_sum",
           "embed_edge_init", "seg_extreme", "linear", "atom_layer"]


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


def edge_init_op(node_feats: Tensor, edge_feats: Tensor, csr: ops.GraphCSR) -> Tensor:
    """K0 through the dispatcher: ``gather_add`` forward; its backward is the identity on ``edge_feats`` and K5 (``seg_reduce`` by src)
    on ``node_feats`` — composed from registered ops, so autograd and tracing see only dispatcher calls."""
    return _EdgeInitViaOps.apply(node_feats, edge_feats, csr.src, csr.by_src.rowptr, csr.by_src.perm, csr.V)


class _EdgeInitViaOps(torch.autograd.Function):
    @staticmethod
    def forward(ctx, node_feats, edge_feats, src, src_rowptr, src_perm, V):
        ctx.save_for_backward(src, src_rowptr, src_perm)
        ctx.V = V
        return torch.ops.notorch_b200.gather_add(edge_feats, node_feats, src, None, 1.0)

    @staticmethod
    def backward(ctx, g):
        src, rowptr, perm = ctx.saved_tensors
        gx = torch.ops.notorch_b200.seg_reduce(g.contiguous(), rowptr, perm, src, ctx.V, False, 1.0) if ctx.needs_input_grad[0] else None
        return gx, (g if ctx.needs_input_grad[1] else None), None, None, None, None


def layer_from_csr(h: Tensor, W: Tensor, b: Optional[Tensor], csr: ops.GraphCSR, act: int = 1, act_param: float = 0.0, mean: bool = False,
                   residual: bool = True, p: float = 0.0, seed: int = 0, offset: int = 0, gemm_mode: int = 0) -> Tensor:
    """Convenience wrapper: ``torch.ops.notorch_b200.chemprop_layer`` with the CSR bundle unpacked."""
    out, _ = torch.ops.notorch_b200.chemprop_layer(h, W, b, csr.src, csr.dst, csr.rev, csr.by_dst.rowptr, csr.by_dst.perm, csr.by_src.rowptr,
                                                   csr.by_src.perm, csr.by_rev.rowptr, csr.by_rev.perm, csr.V, act, act_param, mean, residual,
                                                   p, seed, offset, gemm_mode)
    return out


# ------------------------------------------------------------------------------------------------ index preprocessing (K-l)
@torch.library.custom_op("notorch_b200::build_csr", mutates_args=(), device_types="cuda")
def build_csr(keys: Tensor, num_segments: int) -> tuple[Tensor, Tensor, Tensor, Tensor]:
    """Stable CSR of int64 keys: (rowptr [S+1], perm [n], keys32 [n], status [1]; bit 0 of status = a key out of range)."""
    keys = ops._require(keys, "index", torch.int64, 1)
    with torch.cuda.device(keys.device):
        outs = ops._csr_outputs(keys.numel(), num_segments, keys.device)
        status = torch.zeros(1, dtype=torch.int32, device=keys.device)
        ops._launch_build_csr(keys, num_segments, outs, status, slot=0)
    return outs[0], outs[1], outs[2], status


@build_csr.register_fake
def _(keys, num_segments):
    i32 = lambda n: keys.new_empty((n,), dtype=torch.int32)  # noqa: E731
    return i32(num_segments + 1), i32(keys.shape[0]), i32(keys.shape[0]), i32(1)


@torch.library.custom_op("notorch_b200::csr_to_ell", mutates_args=(), device_types="cuda")
def csr_to_ell(rowptr: Tensor, perm: Tensor, num_segments: int) -> Tensor:
    return ops._ell_of(ops.SegmentCSR(rowptr, perm, perm, num_segments))


@csr_to_ell.register_fake
def _(rowptr, perm, num_segments):
    return rowptr.new_empty((num_segments, 4))


@torch.library.custom_op("notorch_b200::collate", mutates_args=(), device_types="cuda")
def collate(num_atoms: Tensor, num_edges: Tensor, local_edge_index: Tensor, local_rev_index: Tensor, V: int, E: int,
            fixed_rev: bool) -> tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    out = ops._collate_packed_raw(num_atoms, num_edges, local_edge_index, local_rev_index, V, E, fixed_rev)
    return (out["edge_index"], out["rev_index"], out["batch_node_index"], out["batch_edge_index"], out["mol_atom_ptr"], out["mol_edge_ptr"])


@collate.register_fake
def _(num_atoms, num_edges, local_edge_index, local_rev_index, V, E, fixed_rev):
    i64 = lambda *shape: num_atoms.new_empty(shape, dtype=torch.int64)  # noqa: E731
    B = num_atoms.shape[0]
    return i64(2, E), i64(E), i64(V), i64(E), num_atoms.new_empty((B + 1,)), num_atoms.new_empty((B + 1,))


# ------------------------------------------------------------------------------------------------ GraphEmbedding (row N1)
@torch.library.custom_op("notorch_b200::embedding_bag_sum", mutates_args=(), device_types="cuda")
def embedding_bag_sum(table: Tensor, idx: Tensor) -> Tensor:
    from .nn.gnn.embed import _embedding_bag_sum_raw

    return _embedding_bag_sum_raw(table, idx)


@embedding_bag_sum.register_fake
def _(table, idx):
    return table.new_empty((idx.shape[0], table.shape[1]))


@torch.library.custom_op("notorch_b200::embedding_bag_backward", mutates_args=(), device_types="cuda")
def embedding_bag_backward(g: Tensor, idx: Tensor, num_types: int) -> Tensor:
    from .nn.gnn.embed import _embedding_bag_backward_raw

    return _embedding_bag_backward_raw(g.contiguous(), idx, num_types)


@embedding_bag_backward.register_fake
def _(g, idx, num_types):
    return g.new_empty((num_types, g.shape[1]))


def _emb_setup(ctx, inputs, output):
    table, idx = inputs
    ctx.save_for_backward(idx)
    ctx.num_types = table.shape[0]


def _emb_backward(ctx, g):
    (idx,) = ctx.saved_tensors
    return torch.ops.notorch_b200.embedding_bag_backward(g, idx, ctx.num_types), None


embedding_bag_sum.register_autograd(_emb_backward, setup_context=_emb_setup)


@torch.library.custom_op("notorch_b200::embed_edge_init", mutates_args=(), device_types="cuda")
def embed_edge_init(table_v: Tensor, table_e: Tensor, node_types: Tensor, edge_types: Tensor, src: Tensor, V: int) -> Tensor:
    return ops._embed_edge_init_raw(table_v, table_e, node_types, edge_types, src, V)


@embed_edge_init.register_fake
def _(table_v, table_e, node_types, edge_types, src, V):
    return table_v.new_empty((edge_types.shape[0], table_v.shape[1]))


@torch.library.custom_op("notorch_b200::embed_edge_init_backward", mutates_args=(), device_types="cuda")
def embed_edge_init_backward(g: Tensor, node_types: Tensor, edge_types: Tensor, src: Tensor, V: int, num_node_types: int,
                             num_edge_types: int) -> tuple[Tensor, Tensor]:
    return ops._embed_edge_init_backward_raw(g.contiguous(), node_types, edge_types, src, V, num_node_types, num_edge_types)


@embed_edge_init_backward.register_fake
def _(g, node_types, edge_types, src, V, num_node_types, num_edge_types):
    return g.new_empty((num_node_types, g.shape[1])), g.new_empty((num_edge_types, g.shape[1]))


def _eei_setup(ctx, inputs, output):
    table_v, table_e, node_types, edge_types, src, V = inputs
    ctx.save_for_backward(node_types, edge_types, src)
    ctx.meta = (V, table_v.shape[0], table_e.shape[0])


def _eei_backward(ctx, g):
    node_types, edge_types, src = ctx.saved_tensors
    gv, ge = torch.ops.notorch_b200.embed_edge_init_backward(g, node_types, edge_types, src, *ctx.meta)
    return gv, ge, None, None, None, None


embed_edge_init.register_autograd(_eei_backward, setup_context=_eei_setup)


# ------------------------------------------------------------------------------------------------ scatter max / min
@torch.library.custom_op("notorch_b200::seg_extreme", mutates_args=(), device_types="cuda")
def seg_extreme(x: Tensor, rowptr: Tensor, perm: Optional[Tensor], keys: Tensor, num_segments: int, is_min: bool) -> tuple[Tensor, Tensor]:
    from ._lib import ACT_IDENTITY

    return ops._seg_extreme_raw(ops._require_float(x, "x"), _seg(rowptr, perm, keys, num_segments), ACT_IDENTITY, 0.0, is_min)


@seg_extreme.register_fake
def _(x, rowptr, perm, keys, num_segments, is_min):
    return x.new_empty((num_segments, x.shape[1])), x.new_empty((num_segments, x.shape[1]), dtype=torch.int32)


@torch.library.custom_op("notorch_b200::seg_extreme_backward", mutates_args=(), device_types="cuda")
def seg_extreme_backward(g: Tensor, arg: Tensor, keys: Tensor, n: int) -> Tensor:
    return ops._seg_extreme_backward_raw(g.contiguous(), arg, keys, n)


@seg_extreme_backward.register_fake
def _(g, arg, keys, n):
    return g.new_empty((n, g.shape[1]))


def _sx_setup(ctx, inputs, output):
    x, _, _, keys, _, _ = inputs
    ctx.save_for_backward(output[1], keys)
    ctx.n = x.shape[0]
    ctx.set_materialize_grads(False)


def _sx_backward(ctx, g_out, g_arg):
    arg, keys = ctx.saved_tensors
    gx = None if g_out is None else torch.ops.notorch_b200.seg_extreme_backward(g_out, arg, keys, ctx.n)
    return gx, None, None, None, None, None


seg_extreme.register_autograd(_sx_backward, setup_context=_sx_setup)


# ------------------------------------------------------------------------------------------------ MLP head
@torch.library.custom_op("notorch_b200::linear", mutates_args=(), device_types="cuda")
def linear(x: Tensor, W: Tensor, b: Optional[Tensor]) -> Tensor:
    from .nn.mlp import _linear_forward_raw

    return _linear_forward_raw(x, W, b)


@linear.register_fake
def _(x, W, b):
    return x.new_empty((x.shape[0], W.shape[0]))


@torch.library.custom_op("notorch_b200::linear_backward", mutates_args=(), device_types="cuda")
def linear_backward(g: Tensor, x: Tensor, W: Tensor, has_bias: bool) -> tuple[Tensor, Tensor, Tensor]:
    from .nn.mlp import _linear_backward_raw

    gx, gW, gb = _linear_backward_raw(g.contiguous(), x, W, has_bias, True, True)
    return gx, gW, gb if gb is not None else W.new_zeros(W.shape[0])


@linear_backward.register_fake
def _(g, x, W, has_bias):
    return torch.empty_like(x), torch.empty_like(W), W.new_empty(W.shape[0])


def _lin_setup(ctx, inputs, output):
    x, W, b = inputs
    ctx.save_for_backward(x, W)
    ctx.has_bias = b is not None


def _lin_backward(ctx, g):
    x, W = ctx.saved_tensors
    gx, gW, gb = torch.ops.notorch_b200.linear_backward(g, x, W, ctx.has_bias)
    return gx, gW, (gb if ctx.has_bias else None)


linear.register_autograd(_lin_backward, setup_context=_lin_setup)


# ------------------------------------------------------------------------------------------------ atom-state message passing (A10)
def _acsr(V: int, in_rowptr: Tensor, in_perm: Tensor, out_rowptr: Tensor, out_perm: Tensor, ident: Tensor) -> ops.AtomCSR:
    return ops.AtomCSR(V, in_perm.numel(), ops.SegmentCSR(in_rowptr, in_perm, in_perm, V), ops.SegmentCSR(out_rowptr, out_perm, out_perm, V), ident)


@torch.library.custom_op("notorch_b200::atom_layer", mutates_args=(), device_types="cuda")
def atom_layer(h: Tensor, s_e: Tensor, W: Tensor, b: Optional[Tensor], in_rowptr: Tensor, in_perm: Tensor, out_rowptr: Tensor, out_perm: Tensor,
               ident: Tensor, act: int, act_param: float, mean: bool, residual: bool, p: float, seed: int, offset: int,
               gemm_mode: int) -> tuple[Tensor, Tensor]:
    """Returns (h', n): n = s_e + reduce_in(act(h)[src]) is what the backward needs."""
    acsr = _acsr(h.shape[0], in_rowptr, in_perm, out_rowptr, out_perm, ident)
    return ops._atom_layer_forward_raw(h, s_e, W, b, acsr, act, act_param, mean, residual, p, seed, offset, gemm_mode)


@atom_layer.register_fake
def _(h, s_e, W, b, *rest):
    return torch.empty_like(h), torch.empty_like(h)


@torch.library.custom_op("notorch_b200::atom_layer_backward", mutates_args=(), device_types="cuda")
def atom_layer_backward(g: Tensor, h: Tensor, n: Tensor, W: Tensor, has_bias: bool, in_rowptr: Tensor, in_perm: Tensor, out_rowptr: Tensor,
                        out_perm: Tensor, ident: Tensor, act: int, act_param: float, mean: bool, residual: bool, p: float, seed: int,
                        offset: int, gemm_mode: int) -> tuple[Tensor, Tensor, Tensor, Tensor]:
    acsr = _acsr(h.shape[0], in_rowptr, in_perm, out_rowptr, out_perm, ident)
    gh, gs, gW, gb = ops._atom_layer_backward_raw(g.contiguous(), h, n, W, has_bias, acsr, act, act_param, mean, residual, p, seed, offset, gemm_mode,
                                                  True, True)
    return gh, gs, gW, gb if gb is not None else W.new_zeros(W.shape[0])


@atom_layer_backward.register_fake
def _(g, h, n, W, has_bias, *rest):
    return torch.empty_like(h), torch.empty_like(h), torch.empty_like(W), W.new_empty(W.shape[0])


def _al_setup(ctx, inputs, output):
    h, s_e, W, b = inputs[:4]
    ctx.save_for_backward(h, output[1], W, *inputs[4:9])
    ctx.has_bias = b is not None
    ctx.cfg = tuple(inputs[9:])


def _al_backward(ctx, g_out, g_n):
    h, n, W, *lists = ctx.saved_tensors
    gh, gs, gW, gb = torch.ops.notorch_b200.atom_layer_backward(g_out.contiguous(), h, n, W, ctx.has_bias, *lists, *ctx.cfg)
    return (gh, gs, gW, gb if ctx.has_bias else None) + (None,) * 13


atom_layer.register_autograd(_al_backward, setup_context=_al_setup)
