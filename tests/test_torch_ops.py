"""torch.library registration: fake-tensor shape inference without a GPU, and on the GPU the dispatcher ops against the
autograd.Function path the nn modules use (same kernels, so bit-identical) including the registered backward formulas."""
from __future__ import annotations

import pytest
import torch

from helpers import oracle_inputs


def test_ops_are_registered_with_fake_kernels():
    import notorch_b200.torch_ops  # noqa: F401  (registers torch.ops.notorch_b200.*)
    from torch._subclasses.fake_tensor import FakeTensorMode

    for name in ("seg_reduce", "gather_add", "chemprop_layer", "chemprop_layer_backward"):
        assert hasattr(torch.ops.notorch_b200, name)
    E, V, d = 10, 6, 16
    with FakeTensorMode():
        x = torch.empty(E, d, device="cuda")
        i32 = lambda n: torch.empty(n, dtype=torch.int32, device="cuda")  # noqa: E731
        out = torch.ops.notorch_b200.seg_reduce(x, i32(V + 1), i32(E), i32(E), V, False, 1.0)
        assert out.shape == (V, d) and out.dtype == torch.float32
        out = torch.ops.notorch_b200.gather_add(None, torch.empty(V, d, device="cuda"), i32(E), None, 1.0)
        assert out.shape == (E, d)
        W = torch.empty(d, d, device="cuda")
        h2, saved = torch.ops.notorch_b200.chemprop_layer(x, W, None, i32(E), i32(E), i32(E), i32(V + 1), i32(E), i32(V + 1), i32(E), i32(E + 1),
                                                          i32(E), V, 1, 0.0, False, True, 0.0, 0, 0, 0)
        assert h2.shape == (E, d) and saved.shape == (E, d)
        _, saved32 = torch.ops.notorch_b200.chemprop_layer(x, W, None, i32(E), i32(E), i32(E), i32(V + 1), i32(E), i32(V + 1), i32(E), i32(E + 1),
                                                           i32(E), V, 1, 0.0, False, True, 0.0, 0, 0, 1)
        assert saved32.shape == (V, d)  # strict-fp32 mode saves n [V, d] instead of m [E, d]


def test_cpu_tensors_have_no_kernel():
    import notorch_b200.torch_ops  # noqa: F401

    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.notorch_b200.gather_add(None, torch.zeros(4, 8), torch.zeros(3, dtype=torch.int32), None, 1.0)


@pytest.mark.gpu
def test_dispatcher_ops_match_the_module_path():
    from notorch_b200 import ops, torch_ops

    inp = oracle_inputs(24, 300, 1, config=1, seed=13)
    csr = ops.build_graph_csr(inp["edge_index"].cuda(), inp["rev_index"].cuda(), inp["V"])
    gen = torch.Generator().manual_seed(3)
    h = torch.randn(inp["E"], 300, generator=gen).cuda()
    g = torch.randn(inp["E"], 300, generator=gen).cuda()
    W, b = inp["weights"][0].cuda(), inp["biases"][0].cuda()

    def run(fn):
        hh, WW, bb = h.clone().requires_grad_(True), W.clone().requires_grad_(True), b.clone().requires_grad_(True)
        out = fn(hh, WW, bb)
        (out * g).sum().backward()
        return out.detach(), hh.grad, WW.grad, bb.grad

    ref = run(lambda hh, WW, bb: ops.layer(hh, WW, bb, csr, residual=True))
    got = run(lambda hh, WW, bb: torch_ops.layer_from_csr(hh, WW, bb, csr, residual=True))
    for a, r, name in zip(got, ref, ("out", "grad h", "grad W", "grad b")):
        assert torch.equal(a, r), name

    # seg_reduce + its registered backward (a gather) vs the functional API
    x = h.clone().requires_grad_(True)
    y = torch.ops.notorch_b200.seg_reduce(x, csr.by_dst.rowptr, csr.by_dst.perm, csr.dst, csr.V, True, 1.0)
    (y * y).sum().backward()
    x2 = h.clone().requires_grad_(True)
    y2 = ops.edge_to_atom(x2, csr, "mean")
    (y2 * y2).sum().backward()
    assert torch.equal(y, y2) and torch.equal(x.grad, x2.grad)


@pytest.mark.gpu
def test_opcheck_of_the_registrations():
    from notorch_b200 import ops, torch_ops  # noqa: F401

    inp = oracle_inputs(8, 64, 1, config=1, seed=1)
    csr = ops.build_graph_csr(inp["edge_index"].cuda(), inp["rev_index"].cuda(), inp["V"])
    x = torch.randn(inp["E"], 64, device="cuda", requires_grad=True)
    torch.library.opcheck(torch.ops.notorch_b200.seg_reduce.default, (x, csr.by_dst.rowptr, csr.by_dst.perm, csr.dst, csr.V, False, 1.0),
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))
    torch.library.opcheck(torch.ops.notorch_b200.gather_add.default, (None, torch.randn(inp["V"], 64, device="cuda"), csr.src, None, 1.0),
                          test_utils=("test_schema", "test_faketensor"))
