"""Atom-state message passing (extension, parity unpinned): the CUDA module against the plain-PyTorch
oracle ``oracle/atom_mp_oracle.py`` in fp64, forward and backward, plus the CPU-only checks of the oracle itself."""
from __future__ import annotations

import pytest
import torch

from helpers import ACT_MODULES, REL_F32, assert_close, oracle_inputs
from oracle.atom_mp_oracle import AtomMessagePassingOracle


def _oracle_run(inp, depth, act, reduce, residual, bias, shared, seed=0):
    torch.manual_seed(seed)
    ref = AtomMessagePassingOracle(hidden_dim=inp["d"], act=ACT_MODULES[act], bias=bias, depth=depth, residual=residual, shared=shared,
                                   reduce=reduce).double()
    xv = inp["x_v"].double().requires_grad_(True)
    xe = inp["x_e"].double().requires_grad_(True)
    out = ref(xv, xe, inp["edge_index"])
    gen = torch.Generator().manual_seed(seed + 1)
    cot = torch.randn(out.shape, generator=gen, dtype=torch.float64)
    (out * cot).sum().backward()
    return ref, out, cot, xv.grad, xe.grad


def test_oracle_matches_a_loop_restatement():
    """The oracle against the definition written as explicit Python loops (tiny case)."""
    inp = oracle_inputs(3, 8, 2, config=1, seed=5)
    ref, out, *_ = _oracle_run(inp, 2, "relu", "mean", True, True, False)
    src, dst = inp["edge_index"]
    h = inp["x_v"].double()
    for entry in ref.layers:
        lin = entry.module.update[0]
        a = torch.relu(h)
        n = torch.zeros_like(h)
        cnt = torch.zeros(len(h), dtype=torch.float64)
        for e in range(len(src)):
            n[dst[e]] += a[src[e]] + inp["x_e"].double()[e]
            cnt[dst[e]] += 1
        n = n / cnt.clamp(min=1)[:, None]
        h = h + n @ lin.weight.T + lin.bias
    assert torch.allclose(h, out.detach(), rtol=1e-12, atol=1e-12)


def test_oracle_state_dict_keys_follow_the_reference_convention():
    ref = AtomMessagePassingOracle(hidden_dim=8, depth=2)
    assert sorted(ref.state_dict()) == ["layers.0.module.update.0.bias", "layers.0.module.update.0.weight", "layers.1.module.update.0.bias",
                                        "layers.1.module.update.0.weight"]


@pytest.mark.gpu
@pytest.mark.parametrize("act,reduce,residual,bias,shared,d,depth", [
    ("relu", "sum", True, True, False, 300, 3),    # BASELINE configs[3] shape
    ("relu", "mean", True, True, False, 64, 2),
    ("silu", "sum", False, True, False, 128, 2),
    ("tanh", "mean", True, False, True, 96, 3),
    ("gelu", "sum", True, True, False, 320, 1),
])
def test_atom_message_passing_matches_oracle(act, reduce, residual, bias, shared, d, depth):
    from notorch_b200 import BatchedGraph
    from notorch_b200.nn import AtomMessagePassing

    inp = oracle_inputs(48, d, depth, config=2, seed=11)
    ref, out_ref, cot, gxv_ref, gxe_ref = _oracle_run(inp, depth, act, reduce, residual, bias, shared)
    blk = AtomMessagePassing(hidden_dim=d, act=ACT_MODULES[act], bias=bias, depth=depth, residual=residual, shared=shared, reduce=reduce)
    blk.load_state_dict({k: v.float() for k, v in ref.state_dict().items()}, strict=True)
    blk = blk.cuda()
    xv = inp["x_v"].cuda().requires_grad_(True)
    xe = inp["x_e"].cuda().requires_grad_(True)
    G = BatchedGraph(xv, xe, inp["edge_index"].cuda(), inp["rev_index"].cuda(), batch_node_index=inp["batch_node_index"].cuda(),
                     batch_edge_index=inp["batch_edge_index"].cuda(), size=inp["B"])
    out = blk(G)
    assert out.edge_feats is xe  # edge features pass through untouched
    assert_close(out.node_feats, out_ref, "node_feats", REL_F32)
    (out.node_feats * cot.float().cuda()).sum().backward()
    # smooth activations only: a ReLU sign flip on an element with |h| ~ 1e-7 would change single gradient entries by O(1)
    tol = 3e-5 if act == "relu" else REL_F32 * 2
    assert_close(xv.grad, gxv_ref, "grad x_v", tol)
    assert_close(xe.grad, gxe_ref, "grad x_e", tol)
    for (k, p_), (_, q) in zip(sorted(blk.named_parameters()), sorted(ref.named_parameters())):
        assert_close(p_.grad, q.grad, f"grad {k}", tol)


@pytest.mark.gpu
def test_atom_message_passing_inference_with_norm_readout_is_deterministic():
    """BASELINE configs[3]: atom message passing + Norm pooling, inference only; two runs are bit-identical."""
    from notorch_b200 import BatchedGraph
    from notorch_b200.nn import AtomMessagePassing, Norm

    inp = oracle_inputs(256, 300, 3, config=2, seed=3)
    torch.manual_seed(0)
    blk = AtomMessagePassing(hidden_dim=300, depth=3).cuda().eval()
    agg = Norm(100.0)
    G = BatchedGraph(inp["x_v"].cuda(), inp["x_e"].cuda(), inp["edge_index"].cuda(), inp["rev_index"].cuda(),
                     batch_node_index=inp["batch_node_index"].cuda(), batch_edge_index=inp["batch_edge_index"].cuda(), size=inp["B"])
    with torch.no_grad():
        H1 = agg(blk(G))
        H2 = agg(blk(G))
    assert H1.shape == (256, 300) and torch.equal(H1, H2)
    ref = AtomMessagePassingOracle(hidden_dim=300, depth=3).double()
    ref.load_state_dict({k: v.double().cpu() for k, v in blk.state_dict().items()})
    with torch.no_grad():
        h = ref(inp["x_v"].double(), inp["x_e"].double(), inp["edge_index"])
        Href = torch.zeros(256, 300, dtype=torch.float64).index_add_(0, inp["batch_node_index"], h) / 100.0
    assert_close(H1, Href, "Norm read-out of atom message passing", REL_F32)
