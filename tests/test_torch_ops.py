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


def _fake_model_pass(backward: bool):
    import types

    import notorch_b200.torch_ops  # noqa: F401
    from torch._subclasses.fake_tensor import FakeTensorMode

    from notorch_b200 import BatchedGraph, _lib
    from notorch_b200.nn import MLP, AtomMessagePassing, ChempropBlock, GraphEmbedding, Max, Mean, Norm

    B, V, E, d = 4, 30, 64, 32
    n0 = _lib.lib().nt_kernel_launch_count()
    with FakeTensorMode():
        dev = "cuda"
        i32 = lambda *s: torch.empty(s, dtype=torch.int32, device=dev)  # noqa: E731
        packed = types.SimpleNamespace(num_atoms=i32(B), num_edges=i32(B), edge_index=i32(2, E), rev_index=i32(E))
        node_types = torch.empty(V, 7, dtype=torch.int64, device=dev)
        edge_types = torch.empty(E, 2, dtype=torch.int64, device=dev)
        G = BatchedGraph.from_packed(packed, node_types, edge_types, device=dev)
        assert G.edge_index.shape == (2, E) and G.edge_index.dtype == torch.int64 and G.batch_node_index.shape == (V,)
        with torch.device(dev):  # parameters are created as fake CUDA tensors (Module.to() cannot swap fake parameters)
            embed, block, head = GraphEmbedding(hidden_dim=d), ChempropBlock(hidden_dim=d, depth=2), MLP(d, 3, hidden_dim=16)
            block_max, atom = ChempropBlock(hidden_dim=d, depth=0, reduce="max"), AtomMessagePassing(hidden_dim=d, depth=2)
        G1 = block(embed(G))
        assert G1.node_feats.shape == (V, d) and G1.edge_feats.shape == (E, d)
        y = head(Mean()(G1))
        assert y.shape == (B, 3) and y.requires_grad and Max()(G1).shape == (B, d)
        if backward:  # the autograd engine binds the CUDA device of its worker thread: needs a real GPU even for fake tensors
            (y.sum() + Max()(G1).sum()).backward()
            assert embed.node.weight.grad.shape == (45, d) and block.layers[0].module.update[0].weight.grad.shape == (d, d)
            assert head[0].weight.grad.shape == (16, d)
        # the final reduction of a reduce = max block routes through seg_extreme (its LAYERS with max / min stay autograd.Function-only);
        # the atom-state variant through atom_layer
        G2 = block_max(embed(G))
        assert G2.node_feats.shape == (V, d)
        H = Norm()(atom(embed(G)))
        assert H.shape == (B, d)
    assert _lib.lib().nt_kernel_launch_count() == n0  # nothing ran


def test_whole_model_traces_with_fake_tensors_on_a_cpu_box():
    """The nn modules route through torch.ops.notorch_b200.* by themselves when they see fake tensors: device collation -> GraphEmbedding
    (deferred, fused into the edge initialisation) -> ChempropBlock -> Mean / Max -> MLP head, block reduce = max, atom message passing,
    under FakeTensorMode with fake CUDA tensors - no GPU, no kernel launch: shapes, dtypes and autograd wiring only."""
    _fake_model_pass(backward=False)


@pytest.mark.gpu
def test_whole_model_traces_with_fake_tensors_including_backward():
    _fake_model_pass(backward=True)


def test_every_dispatcher_op_is_registered():
    import notorch_b200.torch_ops as T

    for name in T.__all__:
        if name in ("layer_from_csr",):
            continue
        assert hasattr(torch.ops.notorch_b200, name), name
    for name in ("embedding_bag_backward", "embed_edge_init_backward", "seg_extreme_backward", "linear_backward", "atom_layer_backward"):
        assert hasattr(torch.ops.notorch_b200, name), name


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


@pytest.mark.gpu
def test_opcheck_of_the_round2_registrations():
    """torch.library.opcheck (schema, fake-tensor agreement, autograd registration) of the ops added in round 2."""
    from notorch_b200 import ops, torch_ops  # noqa: F401

    ops.set_index_validation("sync")
    inp = oracle_inputs(8, 64, 1, config=1, seed=1)
    V, E, d = inp["V"], inp["E"], 64
    ei, rev = inp["edge_index"].cuda(), inp["rev_index"].cuda()
    csr = ops.build_graph_csr(ei, rev, V)
    ns = torch.ops.notorch_b200
    full = ("test_schema", "test_faketensor", "test_autograd_registration")
    torch.library.opcheck(ns.build_csr.default, (ei[1].contiguous(), V), test_utils=("test_schema", "test_faketensor"))
    torch.library.opcheck(ns.csr_to_ell.default, (csr.by_dst.rowptr, csr.by_dst.perm, V), test_utils=("test_schema", "test_faketensor"))
    mols = inp["mols"]
    i32 = lambda a: torch.from_numpy(a).cuda()  # noqa: E731
    torch.library.opcheck(ns.collate.default, (i32(mols.num_atoms), i32(mols.num_edges), i32(mols.edge_index), i32(mols.rev_index), V, E, False),
                          test_utils=("test_schema", "test_faketensor"))
    gen = torch.Generator().manual_seed(0)
    nv, ne = torch.randint(0, 45, (V, 7), generator=gen).cuda(), torch.randint(0, 13, (E, 2), generator=gen).cuda()
    tv, te = torch.randn(45, d, device="cuda", requires_grad=True), torch.randn(13, d, device="cuda", requires_grad=True)
    torch.library.opcheck(ns.embedding_bag_sum.default, (tv, nv), test_utils=full)
    torch.library.opcheck(ns.embed_edge_init.default, (tv, te, nv, ne, csr.src, V), test_utils=full)
    x = torch.randn(E, d, device="cuda", requires_grad=True)
    torch.library.opcheck(ns.seg_extreme.default, (x, csr.by_dst.rowptr, csr.by_dst.perm, csr.dst, V, False), test_utils=full)
    W, b = torch.randn(16, d, device="cuda", requires_grad=True), torch.randn(16, device="cuda", requires_grad=True)
    torch.library.opcheck(ns.linear.default, (torch.randn(8, d, device="cuda", requires_grad=True), W, b), test_utils=full)
    acsr = ops.atom_csr(csr)
    h, s_e = torch.randn(V, d, device="cuda", requires_grad=True), torch.randn(V, d, device="cuda")
    Wd = (torch.randn(d, d, device="cuda") / 8).requires_grad_(True)
    torch.library.opcheck(ns.atom_layer.default, (h, s_e, Wd, None, acsr.nbr_in.rowptr, acsr.nbr_in.perm, acsr.nbr_out.rowptr, acsr.nbr_out.perm,
                                                  acsr.ident, 1, 0.0, False, True, 0.0, 0, 0, 0), test_utils=full)


@pytest.mark.gpu
def test_modules_through_the_dispatcher_equal_the_function_path():
    """ops.set_dispatch("ops"): every module call goes through torch.ops.notorch_b200.* - same kernels, so the same bits, forward and
    backward, for GraphEmbedding (fused) -> ChempropBlock -> Sum -> MLP."""
    from notorch_b200 import BatchedGraph, ops
    from notorch_b200.nn import MLP, ChempropBlock, GraphEmbedding, Sum

    ops.set_index_validation("sync")
    inp = oracle_inputs(24, 8, 0, config=1, seed=3)
    gen = torch.Generator().manual_seed(5)
    nv, ne = torch.randint(0, 45, (inp["V"], 7), generator=gen), torch.randint(0, 13, (inp["E"], 2), generator=gen)
    torch.manual_seed(0)
    embed, block, head = GraphEmbedding(hidden_dim=64).cuda(), ChempropBlock(hidden_dim=64, depth=2).cuda(), MLP(64, 2, hidden_dim=32).cuda()
    params = list(embed.parameters()) + list(block.parameters()) + list(head.parameters())
    res = []
    for mode in ("function", "ops"):
        ops.set_dispatch(mode)
        try:
            for p in params:
                p.grad = None
            G = BatchedGraph(nv.cuda(), ne.cuda(), inp["edge_index"].cuda(), inp["rev_index"].cuda(), batch_node_index=inp["batch_node_index"].cuda(),
                             batch_edge_index=inp["batch_edge_index"].cuda(), size=24)
            y = head(Sum()(block(embed(G))))
            y.square().mean().backward()
            res.append([y.detach().clone()] + [p.grad.clone() for p in params])
        finally:
            ops.set_dispatch("function")
    for a, b in zip(*res):
        assert torch.equal(a, b)


@pytest.mark.gpu
def test_torch_compile_traces_the_functional_api_and_the_modules():
    """torch.compile (dynamo + AOT autograd, backend aot_eager: tracing, fake tensors, functionalisation, joint forward/backward graph -
    everything but a code generator): the functional layer API compiles as ONE graph (fullgraph=True) through
    torch.ops.notorch_b200.chemprop_layer and differentiates; ChempropBlock + read-out compile with graph breaks only at copy.copy
    (Graph.update is a shallow copy by contract, utils.py:34-40). Same kernels underneath, so the same bits as eager."""
    import torch._dynamo

    from notorch_b200 import BatchedGraph, ops
    from notorch_b200.nn import ChempropBlock, Sum

    ops.set_index_validation("off")
    try:
        p = oracle_inputs(16, 64, 2, seed=3)
        csr = ops.build_graph_csr(p["edge_index"].cuda(), p["rev_index"].cuda(), p["V"])
        h = torch.randn(p["E"], 64, device="cuda", requires_grad=True)
        W = (torch.randn(64, 64, device="cuda") / 8).requires_grad_(True)
        b = torch.zeros(64, device="cuda", requires_grad=True)

        def f(h, W, b):
            return ops.layer(ops.layer(h, W, b, csr), W, b, csr).square().mean()

        ref = f(h, W, b)
        ref.backward()
        want = (h.grad.clone(), W.grad.clone())
        h.grad = W.grad = b.grad = None
        torch._dynamo.reset()
        out = torch.compile(f, backend="aot_eager", fullgraph=True)(h, W, b)
        out.backward()
        assert torch.equal(out, ref) and torch.equal(h.grad, want[0]) and torch.equal(W.grad, want[1])

        blk = ChempropBlock(hidden_dim=64, depth=2).cuda()
        G = BatchedGraph(p["x_v"].cuda(), p["x_e"].cuda(), p["edge_index"].cuda(), p["rev_index"].cuda(), batch_node_index=p["batch_node_index"].cuda(),
                         batch_edge_index=p["batch_edge_index"].cuda(), size=16)

        def m(G):
            return Sum()(blk(G))

        ref = m(G)
        torch._dynamo.reset()
        assert torch.equal(torch.compile(m, backend="aot_eager")(G), ref)
    finally:
        torch._dynamo.reset()
        ops.set_index_validation("sync")
