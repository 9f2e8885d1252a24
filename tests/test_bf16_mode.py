"""bf16 operand mode (BASELINE configs[4]: "bf16 W_h with fp32 accumulation"): W_h and the message operand are rounded to bf16 and
multiplied by one tcgen05 kind::f16 pass with fp32 accumulation; activations stay fp32 in HBM, the weight gradient runs as a
single TF32 pass.

Stated bounds (the north star asks for "a looser stated bound for a bf16 mode"; SURVEY.md §8c suggests these):
  * embeddings: rel-to-max 2e-2 against the exact fp64 oracle;
  * gradients:  relative L2 error (|x - ref|_2 / |ref|_2) 5e-2 against the exact fp64 oracle for every activation, and rel-to-max
    5e-2 for smooth activations. Through ReLU the element-wise (rel-to-max) comparison is dominated by derivative flips of elements
    with |h| below the bf16 error - an O(1) change of single gradient entries, 12-30 % rel-to-max even for an fp64 EMULATION of the
    mode (same arithmetic, GEMM operands rounded to bf16), independent of any kernel - while the L2 error stays at 1-3 %.
  The kernels themselves are pinned much tighter against that emulation (rel L2 2e-2)."""
from __future__ import annotations

import pytest
import torch

from helpers import ACT_MODULES, assert_close, oracle_inputs, rel_err

REL_BF16_FWD, REL_BF16_BWD, L2_VS_EMULATION = 2e-2, 5e-2, 2e-2


def _l2_err(x, ref) -> float:
    x, ref = x.detach().double().cpu(), ref.detach().double().cpu()
    return float((x - ref).norm() / ref.norm())


def _assert_l2(x, ref, what: str, tol: float):
    err = _l2_err(x, ref)
    assert err <= tol, f"{what}: relative L2 error {err:.3e} > {tol:.1e}"


def _bf(x):
    return x.float().bfloat16().double()


class _Bf16Linear(torch.autograd.Function):
    """u = bf16(m) . bf16(W)^T in fp64; backward: g_m = bf16(g) . bf16(W), g_W = g^T . m (the weight gradient is a TF32 pass of the
    unrounded fp32 operands: exact at this resolution)."""

    @staticmethod
    def forward(ctx, m, W, emulate):
        ctx.save_for_backward(m, W)
        ctx.emulate = emulate
        return (_bf(m) @ _bf(W).T) if emulate else m @ W.T

    @staticmethod
    def backward(ctx, g):
        m, W = ctx.saved_tensors
        return ((_bf(g) @ _bf(W)) if ctx.emulate else g @ W), g.T @ m, None


def _reference(inp, depth, act, emulate, seed=5):
    f64 = torch.float64
    src, dst = inp["edge_index"]
    rev, V, d, B = inp["rev_index"], inp["V"], inp["d"], inp["B"]
    xv = inp["x_v"].to(f64).requires_grad_(True)
    xe = inp["x_e"].to(f64).requires_grad_(True)
    Ws = [w.to(f64).requires_grad_(True) for w in inp["weights"]]
    bs = [b.to(f64).requires_grad_(True) for b in inp["biases"]]
    fn = ACT_MODULES[act]()
    h = xv[src] + xe
    for W, b in zip(Ws, bs):
        a = fn(h)
        n = torch.zeros(V, d, dtype=f64).index_add(0, dst, a)
        h = h + _Bf16Linear.apply(n[src] - a[rev], W, emulate) + b
    node = torch.zeros(V, d, dtype=f64).index_add(0, dst, h)
    H = torch.zeros(B, d, dtype=f64).index_add(0, inp["batch_node_index"], node)
    gH = torch.randn(H.shape, generator=torch.Generator().manual_seed(seed), dtype=f64)
    (H * gH).sum().backward()
    return dict(h=h.detach(), H=H.detach(), gH=gH, gxv=xv.grad, gxe=xe.grad, gW=[w.grad for w in Ws], gb=[b.grad for b in bs])


@pytest.fixture
def bf16_mode():
    from notorch_b200 import ops

    old = ops.get_gemm_mode()
    ops.set_gemm_mode("bf16")
    yield
    ops.set_gemm_mode(old)


def test_emulation_reproduces_the_exact_reference_when_switched_off():
    inp = oracle_inputs(4, 16, 2, config=1, seed=2)
    a, b = _reference(inp, 2, "relu", False), _reference(inp, 2, "relu", False)
    assert torch.equal(a["H"], b["H"])
    e = _reference(inp, 2, "silu", True)
    x = _reference(inp, 2, "silu", False)
    assert 1e-5 < rel_err(e["h"], x["h"]) < REL_BF16_FWD  # the emulated mode differs from exact arithmetic by bf16 rounding only


@pytest.mark.gpu
@pytest.mark.parametrize("act,d,depth,batch", [("relu", 300, 3, 64), ("relu", 64, 2, 32), ("relu", 256, 2, 48), ("silu", 300, 3, 64),
                                               ("tanh", 128, 2, 32), ("relu", 512, 1, 24), ("silu", 2048, 1, 8)])
def test_block_bf16_mode(bf16_mode, act, d, depth, batch):
    from notorch_b200 import BatchedGraph
    from notorch_b200.nn import ChempropBlock, Sum

    inp = oracle_inputs(batch, d, depth, config=1, seed=21)
    exact, emul = _reference(inp, depth, act, False), _reference(inp, depth, act, True)
    blk = ChempropBlock(hidden_dim=d, depth=depth, act=ACT_MODULES[act]).cuda()
    with torch.no_grad():
        for i, layer in enumerate(blk.layers):
            layer.module.update[0].weight.copy_(inp["weights"][i])
            layer.module.update[0].bias.copy_(inp["biases"][i])
    xv, xe = inp["x_v"].cuda().requires_grad_(True), inp["x_e"].cuda().requires_grad_(True)
    G = BatchedGraph(xv, xe, inp["edge_index"].cuda(), inp["rev_index"].cuda(), batch_node_index=inp["batch_node_index"].cuda(),
                     batch_edge_index=inp["batch_edge_index"].cuda(), size=inp["B"])
    out = blk(G)
    H = Sum()(out)
    assert_close(out.edge_feats, exact["h"], "h_L (bf16 mode vs exact)", REL_BF16_FWD)
    assert_close(H, exact["H"], "H (bf16 mode vs exact)", REL_BF16_FWD)
    assert rel_err(out.edge_feats, exact["h"]) > 1e-5, "bf16 mode is suspiciously exact: is the tf32x3 kernel running instead?"
    (H * exact["gH"].float().cuda()).sum().backward()
    got = dict(gxv=xv.grad, gxe=xe.grad, gW=[l.module.update[0].weight.grad for l in blk.layers],
               gb=[l.module.update[0].bias.grad for l in blk.layers])
    keys = [("gxv", got["gxv"], emul["gxv"], exact["gxv"]), ("gxe", got["gxe"], emul["gxe"], exact["gxe"])]
    keys += [(f"gW{i}", got["gW"][i], emul["gW"][i], exact["gW"][i]) for i in range(depth)]
    keys += [(f"gb{i}", got["gb"][i], emul["gb"][i], exact["gb"][i]) for i in range(depth)]
    for name, x, em, ex in keys:
        _assert_l2(x, em, f"{name} (bf16 mode vs its fp64 emulation)", L2_VS_EMULATION)  # pins the kernels
        _assert_l2(x, ex, f"{name} (bf16 mode vs exact)", REL_BF16_BWD)                    # the stated bound
        if act != "relu":
            assert_close(x, ex, f"{name} (bf16 mode vs exact, element-wise)", REL_BF16_BWD)


@pytest.mark.gpu
def test_bf16_layer_is_deterministic_and_equals_the_product_of_rounded_operands(bf16_mode):
    """One layer: the kernel's result is the fp64 product of the bf16-ROUNDED operands up to (a) fp32 accumulation and (b) the few
    message elements that sit on a bf16 rounding boundary (the kernel rounds the fp32 message, the reference an fp64 one)."""
    from notorch_b200 import ops

    E, d = 4096, 300
    gen = torch.Generator().manual_seed(9)
    V = E // 2
    src, dst, rev = (torch.randint(0, n, (E,), generator=gen) for n in (V, V, E))
    h = torch.randn(E, d, generator=gen)
    W = (torch.rand(d, d, generator=gen) * 2 - 1) / d ** 0.5
    b = torch.randn(d, generator=gen) * 0.1
    csr = ops.build_graph_csr(torch.stack([src, dst]).cuda(), rev.cuda(), V)
    out1 = ops.layer(h.cuda(), W.cuda(), b.cuda(), csr, residual=True)
    out2 = ops.layer(h.cuda(), W.cuda(), b.cuda(), csr, residual=True)
    assert torch.equal(out1, out2)
    a = torch.relu(h.double())
    n = torch.zeros(V, d, dtype=torch.float64).index_add_(0, dst, a)
    m = (n[src] - a[rev]).float()
    ref = h.double() + m.bfloat16().double() @ W.bfloat16().double().T + b.double()
    assert_close(out1, ref, "bf16-rounded operands, exact product", 1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("d", [64, 300])
def test_bf16_mode_with_dropout_uses_the_same_keep_mask_forward_and_backward(bf16_mode, d):
    """Dropout in bf16 mode: K2's epilogue mask, K4a's operand mask and K4b's mask are the same function of (seed, offset); checked
    against the fp64 oracle evaluated with the kernel's own keep-mask (smooth comparison: single layer, relative L2)."""
    from notorch_b200 import ops
    from oracle import dmpnn_oracle as O

    p = oracle_inputs(32, d, 1, seed=4)
    E, V, pr = p["E"], p["V"], 0.25
    csr = ops.build_graph_csr(p["edge_index"].cuda(), p["rev_index"].cuda(), V)
    gen = torch.Generator().manual_seed(1)
    h, g = torch.randn(E, d, generator=gen), torch.randn(E, d, generator=gen)
    W, b = p["weights"][0], p["biases"][0]
    seed, offset = 7654321, 5
    mask = ops.dropout_mask(E, d, pr, seed, offset, "cuda").cpu()
    hc, Wc, bc = h.cuda().requires_grad_(True), W.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    out = ops._Layer.apply(hc, Wc, bc, csr, 1, 0.0, False, True, pr, seed, offset, ops._gemm_mode)
    (out * g.cuda()).sum().backward()
    h64, W64, b64 = h.double().requires_grad_(True), W.double().requires_grad_(True), b.double().requires_grad_(True)
    ref, _ = O.layer_forward(h64, V, p["edge_index"][0], p["edge_index"][1], p["rev_index"], W64, b64, keep_mask=mask.bool(), p=pr)
    (ref * g.double()).sum().backward()
    dropped = ~mask.bool()
    assert torch.equal((out.cpu() - h)[dropped], torch.zeros(int(dropped.sum())))  # dropped entries are exactly the residual
    _assert_l2(out, ref.detach(), "dropout out (bf16 mode)", REL_BF16_FWD)
    _assert_l2(hc.grad, h64.grad, "dropout grad h (bf16 mode)", REL_BF16_BWD)
    _assert_l2(Wc.grad, W64.grad, "dropout grad W (bf16 mode)", REL_BF16_BWD)
    _assert_l2(bc.grad, b64.grad, "dropout grad b (bf16 mode)", REL_BF16_BWD)
