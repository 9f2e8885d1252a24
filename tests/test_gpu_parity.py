"""GPU parity tests proper: the CUDA path (through the C ABI) vs the golden vectors produced by the
reference and vs the CPU oracle on seeded inputs. Bit-exact for index / CSR work and for the pure
gather / segmented-sum kernels; rel 1e-5 (to the tensor max) for everything that goes through W."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from helpers import ACT_MODULES, REL_F32, assert_close, block_from_golden, graph_from_golden, oracle_inputs, rel_err
from oracle import dmpnn_oracle as O

pytestmark = pytest.mark.gpu

GEMM_MODES = ["tf32x3", "fp32"]


@pytest.fixture(autouse=True)
def _sync_validation():
    from notorch_b200 import ops

    ops.set_index_validation("sync")
    yield
    ops.set_gemm_mode("tf32x3")


def _agg(kind):
    from notorch_b200.nn import Mean, Sum

    return {"sum": Sum, "mean": Mean}[kind]()


# ---------------------------------------------------------------- integer work: bit-exact
def test_device_collation_matches_reference_golden(golden):
    from notorch_b200 import ops

    dev = "cuda"
    t = lambda k: torch.from_numpy(golden[k]).to(dev)
    V, E = int(golden["num_atoms"].sum()), int(golden["num_edges"].sum())
    out = ops.collate_packed(t("num_atoms"), t("num_edges"), t("local_edge_index"), t("local_rev_index"), V, E)
    for k in ("edge_index", "rev_index", "batch_node_index", "batch_edge_index"):
        assert out[k].dtype == torch.int64
        assert torch.equal(out[k].cpu(), torch.from_numpy(golden[k])), k
    assert np.array_equal(out["mol_atom_ptr"].cpu().numpy(), np.concatenate([[0], np.cumsum(golden["num_atoms"])]))
    assert np.array_equal(out["mol_edge_ptr"].cpu().numpy(), np.concatenate([[0], np.cumsum(golden["num_edges"])]))
    fixed = ops.collate_packed(t("num_atoms"), t("num_edges"), t("local_edge_index"), t("local_rev_index"), V, E, fixed_rev=True)
    assert np.array_equal(fixed["rev_index"].cpu().numpy(), O.collate_fixed(golden.mols())["rev_index"])


@pytest.mark.parametrize("n,S,seed", [(0, 5, 0), (1, 1, 1), (1000, 37, 2), (5000, 5000, 3), (20000, 3, 4), (70000, 9000, 5)])
def test_csr_matches_oracle(n, S, seed):
    from notorch_b200 import ops

    rng = np.random.default_rng(seed)
    keys = rng.integers(0, S, size=n).astype(np.int64)
    if n > 10:
        keys[rng.integers(0, n, size=n // 3)] = S - 1  # one long segment, many empty ones
    rowptr, perm = O.build_csr(keys, S)
    csr = ops.build_segment_csr(torch.from_numpy(keys).cuda(), S)
    assert np.array_equal(csr.rowptr.cpu().numpy(), rowptr)
    assert np.array_equal(csr.perm.cpu().numpy(), perm)
    assert np.array_equal(csr.keys32.cpu().numpy(), keys.astype(np.int32))


def test_csr_out_of_range_raises():
    from notorch_b200 import ops

    keys = torch.tensor([0, 1, 7, 2], dtype=torch.int64, device="cuda")
    with pytest.raises(IndexError):
        ops.build_segment_csr(keys, 4)
    with pytest.raises(IndexError):
        ops.build_segment_csr(torch.tensor([0, -1], dtype=torch.int64, device="cuda"), 4)


def test_graph_csr_of_golden(golden):
    from notorch_b200 import ops

    V, E = int(golden["num_atoms"].sum()), int(golden["num_edges"].sum())
    csr = ops.build_graph_csr(torch.from_numpy(golden["edge_index"]).cuda(), torch.from_numpy(golden["rev_index"]).cuda(), V)
    for seg, keys, S in ((csr.by_src, golden["edge_index"][0], V), (csr.by_dst, golden["edge_index"][1], V), (csr.by_rev, golden["rev_index"], E)):
        rowptr, perm = O.build_csr(keys, S)
        assert np.array_equal(seg.rowptr.cpu().numpy(), rowptr) and np.array_equal(seg.perm.cpu().numpy(), perm)


# ---------------------------------------------------------------- K0 / K1 / K3: bit-exact fp32
@pytest.mark.parametrize("d", [300, 37, 8, 1024])
@pytest.mark.parametrize("reduce", ["sum", "mean"])
def test_segmented_reductions_bit_exact(d, reduce):
    from notorch_b200 import ops

    p = oracle_inputs(48, d, 0, config=1, seed=11)
    csr = ops.build_graph_csr(p["edge_index"].cuda(), p["rev_index"].cuda(), p["V"])
    xe = p["x_e"].cuda()
    got = ops.edge_to_atom(xe, csr, reduce).cpu()
    assert torch.equal(got, O.seg_reduce(p["x_e"], p["edge_index"][1], p["V"], reduce))  # K1 (chemprop.py:86)
    mol = ops.build_segment_csr(p["batch_node_index"].cuda(), p["B"])
    got = ops.readout(p["x_v"].cuda(), mol, reduce).cpu()
    assert torch.equal(got, O.readout(p["x_v"], p["batch_node_index"], p["B"], reduce))  # K3 (agg.py:27,36)
    h0 = ops.edge_init(p["x_v"].cuda(), xe, csr).cpu()
    assert torch.equal(h0, O.edge_init(p["x_v"], p["x_e"], p["edge_index"][0]))  # K0 (chemprop.py:83)


@pytest.mark.parametrize("d", [300, 37, 8])
@pytest.mark.parametrize("reduce", ["max", "min"])
def test_arg_reductions_bit_exact(d, reduce):
    """scatter(..., reduce="max"|"min") (chemprop.py:39,86): values, tie-breaking (first row wins), empty segments -> 0, and the gradient
    routed to the argument row, all bit-exact against the oracle's restatement of torch_scatter."""
    from notorch_b200 import _lib, ops

    p = oracle_inputs(48, d, 0, config=1, seed=13)
    E, V, dst = p["E"], p["V"], p["edge_index"][1]
    x = p["x_e"].clone()
    x[:, : min(4, d)] = torch.randint(-1, 2, (E, min(4, d))).float()  # plenty of exact ties
    csr = ops.build_graph_csr(p["edge_index"].cuda(), p["rev_index"].cuda(), V)
    xc = x.cuda().requires_grad_(True)
    got = ops.edge_to_atom(xc, csr, reduce)
    want, arg = O.seg_extreme(x, dst, V, reduce)
    assert torch.equal(got.detach().cpu(), want)
    g = torch.randn(V, d)
    got.backward(g.cuda())
    want_g = torch.where(arg[dst] == torch.arange(E).view(-1, 1), g[dst], torch.zeros(()))
    assert torch.equal(xc.grad.cpu(), want_g)
    # with an activation prologue (K1 of a max / min layer)
    n, a = ops._seg_extreme_raw(x.cuda(), csr.by_dst, _lib.ACT_RELU, 0.0, reduce == "min")
    w, wa = O.seg_extreme(torch.relu(x), dst, V, reduce)
    assert torch.equal(n.cpu(), w)
    assert torch.equal(torch.where(a < 0, E, a).long().cpu(), wa)  # empty segment: -1 here, len(x) in torch_scatter


@pytest.mark.parametrize("reduce,act", [("max", "relu"), ("min", "silu"), ("max", "tanh")])
@pytest.mark.parametrize("mode", ["tf32x3", "fp32"])
@pytest.mark.parametrize("d", [64, 37])
def test_block_arg_reductions_vs_oracle(reduce, act, mode, d):
    """ChempropBlock with reduce in {max, min}: forward and all gradients vs the oracle (pinned to the reference's own max / min
    goldens in tests/golden/{max,min}_reduce.npz) on a larger seeded batch."""
    from notorch_b200 import BatchedGraph, ops
    from notorch_b200.nn import ChempropBlock

    depth = 3
    p = oracle_inputs(40, d, depth, config=1, seed=21)
    blk = ChempropBlock(hidden_dim=d, act=ACT_MODULES[act], depth=depth, reduce=reduce).cuda()
    with torch.no_grad():
        for l, layer in enumerate(blk.layers):
            layer.module.update[0].weight.copy_(p["weights"][l])
            layer.module.update[0].bias.copy_(p["biases"][l])
    xv, xe = p["x_v"].cuda().requires_grad_(True), p["x_e"].cuda().requires_grad_(True)
    G = BatchedGraph(xv, xe, p["edge_index"].cuda(), p["rev_index"].cuda(), batch_node_index=p["batch_node_index"].cuda(),
                     batch_edge_index=p["batch_edge_index"].cuda(), size=p["B"])
    gN, gE = torch.randn(p["V"], d), torch.randn(p["E"], d)
    old = ops.get_gemm_mode()
    ops.set_gemm_mode(mode)
    try:
        G1 = blk(G)
        ((G1.node_feats * gN.cuda()).sum() + (G1.edge_feats * gE.cuda()).sum()).backward()
    finally:
        ops.set_gemm_mode(old)
    f64 = lambda t: t.double()
    Ws, bs = [f64(w) for w in p["weights"]], [f64(b) for b in p["biases"]]
    node, edge, _ = O.block_forward(f64(p["x_v"]), f64(p["x_e"]), p["edge_index"], p["rev_index"], Ws, bs, act=act, reduce=reduce)
    ref = O.block_backward(f64(p["x_v"]), f64(p["x_e"]), p["edge_index"], p["rev_index"], Ws, bs, f64(gN), f64(gE), act=act, reduce=reduce)
    assert_close(G1.node_feats.detach().cpu().double(), node, "node_out")
    assert_close(G1.edge_feats.detach().cpu().double(), edge, "edge_out")
    assert_close(xv.grad.cpu().double(), ref["x_v"], "grad x_v")
    assert_close(xe.grad.cpu().double(), ref["x_e"], "grad x_e")
    for l, layer in enumerate(blk.layers):
        assert_close(layer.module.update[0].weight.grad.cpu().double(), ref["weights"][l], f"grad W{l}")
        assert_close(layer.module.update[0].bias.grad.cpu().double(), ref["biases"][l], f"grad b{l}")


def test_norm_readout_extension():
    from notorch_b200 import ops

    p = oracle_inputs(16, 64, 0, seed=3)
    mol = ops.build_segment_csr(p["batch_node_index"].cuda(), p["B"])
    got = ops.readout(p["x_v"].cuda(), mol, "norm", 100.0).cpu()
    assert_close(got, O.readout(p["x_v"], p["batch_node_index"], p["B"], "norm", 100.0), "norm readout", 1e-6)


# ---------------------------------------------------------------- whole block vs reference golden
@pytest.mark.parametrize("mode", GEMM_MODES)
def test_block_forward_backward_matches_reference_golden(golden, mode):
    from notorch_b200 import ops

    ops.set_gemm_mode(mode)
    blk = block_from_golden(golden)
    G, xv, xe = graph_from_golden(golden)
    G1 = blk(G)
    H = _agg(golden.meta.get("agg", "sum"))(G1)
    assert G1.edge_index is G.edge_index and G1.rev_index is G.rev_index  # shallow copy, index tensors shared
    assert_close(G1.node_feats, golden["f32/node_out"], "node_out")
    assert_close(G1.edge_feats, golden["f32/edge_out"], "edge_out")
    assert_close(H, golden["f32/H"], "H")
    # vs the fp64 run of the reference as well (accumulation-order noise must stay inside the bound)
    assert_close(G1.edge_feats, golden["f64/edge_out"], "edge_out vs fp64")

    dev = "cuda"
    loss = (H * torch.from_numpy(golden["gH"]).to(dev)).sum() + (G1.edge_feats * torch.from_numpy(golden["gE"]).to(dev)).sum() \
        + (G1.node_feats * torch.from_numpy(golden["gN"]).to(dev)).sum()
    loss.backward()
    assert_close(xv.grad, golden["f32/g_x_v"], "grad x_v")
    assert_close(xe.grad, golden["f32/g_x_e"], "grad x_e")
    ref_grads = golden.grads("f32")
    seen = set()
    for name, prm in blk.named_parameters():
        if id(prm) in seen:
            continue
        seen.add(id(prm))
        assert prm.grad is not None, name
        assert_close(prm.grad, ref_grads[name], f"grad {name}")


@pytest.mark.parametrize("mode", GEMM_MODES)
@pytest.mark.parametrize("batch,d,depth,config", [(64, 300, 3, 1), (256, 300, 3, 2), (32, 256, 2, 1), (16, 1024, 2, 2), (24, 72, 5, 1)])
def test_block_vs_oracle_fresh_inputs(mode, batch, d, depth, config):
    """BASELINE config 1 exactly (B=64, d=300, L=3) and neighbours, vs the CPU oracle (fp32 forward,
    fp64 forward and backward). ReLU sign flips (|h| below the fp32 noise floor) are counted; when
    one occurs the fp64 backward is evaluated on the same piecewise-linear branch (see
    ``dmpnn_oracle.block_backward(act_grad_at=...)``)."""
    from notorch_b200 import ops

    ops.set_gemm_mode(mode)
    p = oracle_inputs(batch, d, depth, config=config, seed=100 + batch)
    gen = torch.Generator().manual_seed(5)
    gH, gE = torch.randn(batch, d, generator=gen), torch.randn(p["E"], d, generator=gen)
    f64 = torch.float64
    ei, rev, bni = p["edge_index"], p["rev_index"], p["batch_node_index"]

    # ---- CUDA path, layer by layer through the public functional API (same calls ChempropBlock makes)
    csr = ops.build_graph_csr(ei.cuda(), rev.cuda(), p["V"])
    mol = ops.build_segment_csr(bni.cuda(), batch)
    xv, xe = p["x_v"].cuda().requires_grad_(True), p["x_e"].cuda().requires_grad_(True)
    Ws = [w.cuda().requires_grad_(True) for w in p["weights"]]
    bs = [b.cuda().requires_grad_(True) for b in p["biases"]]
    hs = [ops.edge_init(xv, xe, csr)]
    for W, b in zip(Ws, bs):
        hs.append(ops.layer(hs[-1], W, b, csr))
    node = ops.edge_to_atom(hs[-1], csr)
    H = ops.readout(node, mol, "mean")
    ((H * gH.cuda()).sum() + (hs[-1] * gE.cuda()).sum()).backward()

    # ---- oracle
    node32, edge32, _ = O.block_forward(p["x_v"], p["x_e"], ei, rev, p["weights"], p["biases"])
    W64, b64 = [w.to(f64) for w in p["weights"]], [b.to(f64) for b in p["biases"]]
    node64, edge64, hs64 = O.block_forward(p["x_v"].to(f64), p["x_e"].to(f64), ei, rev, W64, b64)
    for ref_node, ref_edge, tag in ((node32, edge32, "fp32 oracle"), (node64, edge64, "fp64 oracle")):
        assert_close(node, ref_node, f"node_out vs {tag}")
        assert_close(hs[-1], ref_edge, f"edge_out vs {tag}")
        assert_close(H, O.readout(ref_node, bni, batch, "mean"), f"H vs {tag}")
    flips = sum(int(((hs[l].detach().cpu() > 0) != (hs64[l] > 0)).sum()) for l in range(depth))
    g_node = O.readout_backward(gH.to(f64), bni, p["V"], "mean")
    ref = O.block_backward(p["x_v"].to(f64), p["x_e"].to(f64), ei, rev, W64, b64, g_node, gE.to(f64),
                           act_grad_at=[h.detach().cpu() for h in hs[:depth]] if flips else None)
    print(f"relu sign flips vs fp64 oracle: {flips} of {depth * p['E'] * d}")
    assert flips <= 1e-4 * depth * p["E"] * d
    assert_close(xv.grad, ref["x_v"], "grad x_v")
    assert_close(xe.grad, ref["x_e"], "grad x_e")
    for i in range(depth):
        assert_close(Ws[i].grad, ref["weights"][i], f"grad W{i}")
        assert_close(bs[i].grad, ref["biases"][i], f"grad b{i}")


@pytest.mark.parametrize("mode", GEMM_MODES)
@pytest.mark.parametrize("E,d", [(1, 16), (127, 64), (128, 300), (129, 304), (1000, 256), (777, 512), (300, 1024), (5000, 300), (2500, 100), (200, 320), (260, 2048), (513, 8), (700, 332), (400, 576)])
def test_layer_kernels_vs_fp64(mode, E, d):
    """K2 / K4a / K4b in isolation on random (adversarial: arbitrary src / rev) indices vs an fp64 restatement."""
    from notorch_b200 import ops

    ops.set_gemm_mode(mode)
    gen = torch.Generator().manual_seed(E * 7 + d)
    V = max(1, E // 2)
    src = torch.randint(0, V, (E,), generator=gen)
    dst = torch.randint(0, V, (E,), generator=gen)
    rev = torch.randint(0, E, (E,), generator=gen)
    h = torch.randn(E, d, generator=gen)
    W = (torch.rand(d, d, generator=gen) * 2 - 1) / d ** 0.5
    b = torch.randn(d, generator=gen) * 0.1
    g = torch.randn(E, d, generator=gen)
    ei = torch.stack([src, dst])

    h64, W64, b64 = h.double().requires_grad_(True), W.double().requires_grad_(True), b.double().requires_grad_(True)
    ref, _ = O.layer_forward(h64, V, src, dst, rev, W64, b64, residual=True)
    (ref * g.double()).sum().backward()

    csr = ops.build_graph_csr(ei.cuda(), rev.cuda(), V)
    hc, Wc, bc = h.cuda().requires_grad_(True), W.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    out = ops.layer(hc, Wc, bc, csr, residual=True)
    (out * g.cuda()).sum().backward()
    # The tensor core truncates (does not round) when it adds into its fp32 accumulator, so the 3xTF32 residue grows with the
    # reduction length: 0.7-6e-6 for d = 64 ... 1024; with ONE accumulator it reached 1.2e-5 at d = 2048, so reductions longer than
    # 1024 alternate between two tensor-memory accumulators that the epilogue adds (5.4e-6 at d = 2048).
    tol = REL_F32
    assert_close(out, ref.detach(), "layer out", tol)
    assert_close(hc.grad, h64.grad, "grad h", tol)
    assert_close(Wc.grad, W64.grad, "grad W", tol)
    assert_close(bc.grad, b64.grad, "grad b", tol)


def test_single_pass_tf32_is_outside_the_fp32_bound_but_close():
    """Documents why the default is 3xTF32: one TF32 pass is ~1e-3, not 1e-5."""
    from notorch_b200 import ops

    gen = torch.Generator().manual_seed(0)
    E, d, V = 2048, 256, 900
    src, dst, rev = torch.randint(0, V, (E,), generator=gen), torch.randint(0, V, (E,), generator=gen), torch.randint(0, E, (E,), generator=gen)
    h, W = torch.randn(E, d, generator=gen), torch.randn(d, d, generator=gen) / 16
    ref, _ = O.layer_forward(h.double(), V, src, dst, rev, W.double(), None, residual=False)
    csr = ops.build_graph_csr(torch.stack([src, dst]).cuda(), rev.cuda(), V)
    ops.set_gemm_mode("tf32")
    out = ops.layer(h.cuda(), W.cuda(), None, csr, residual=False)
    err = rel_err(out, ref)
    assert 1e-5 < err < 5e-3, err


# ---------------------------------------------------------------- behaviour
def test_determinism_bitwise():
    from notorch_b200 import BatchedGraph
    from notorch_b200.nn import ChempropBlock, Sum

    p = oracle_inputs(128, 300, 3, config=2, seed=9)
    torch.manual_seed(0)
    blk = ChempropBlock(hidden_dim=300, depth=3).cuda()
    outs = []
    for _ in range(2):
        xv, xe = p["x_v"].cuda().requires_grad_(True), p["x_e"].cuda().requires_grad_(True)
        G = BatchedGraph(xv, xe, p["edge_index"].cuda(), p["rev_index"].cuda(), batch_node_index=p["batch_node_index"].cuda(),
                         batch_edge_index=p["batch_edge_index"].cuda(), size=128)
        blk.zero_grad()
        H = Sum()(blk(G))
        H.square().mean().backward()
        outs.append((H.detach().clone(), xv.grad.clone(), xe.grad.clone(), [q.grad.clone() for q in blk.parameters()]))
    a, b = outs
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    assert all(torch.equal(x, y) for x, y in zip(a[3], b[3]))


@pytest.mark.parametrize("mode", GEMM_MODES)
def test_dropout_mask_consistency(mode):
    """p > 0: forward and backward use the same Philox mask; keep rate ~ 1 - p (the reference's
    Philox stream cannot be matched)."""
    from notorch_b200 import ops

    ops.set_gemm_mode(mode)
    p = oracle_inputs(32, 64, 1, seed=4)
    E, d, V, pr = p["E"], 64, p["V"], 0.25
    csr = ops.build_graph_csr(p["edge_index"].cuda(), p["rev_index"].cuda(), V)
    gen = torch.Generator().manual_seed(1)
    h, g = torch.randn(E, d, generator=gen), torch.randn(E, d, generator=gen)
    W, b = p["weights"][0], p["biases"][0]
    seed, offset = 1234567, 3
    mask = ops.dropout_mask(E, d, pr, seed, offset, "cuda").cpu()
    assert abs(float(mask.mean()) - (1 - pr)) < 0.02
    hc, Wc, bc = h.cuda().requires_grad_(True), W.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    out = ops._Layer.apply(hc, Wc, bc, csr, 1, 0.0, False, True, pr, seed, offset, ops._gemm_mode)
    (out * g.cuda()).sum().backward()
    h64, W64, b64 = h.double().requires_grad_(True), W.double().requires_grad_(True), b.double().requires_grad_(True)
    ref, _ = O.layer_forward(h64, V, p["edge_index"][0], p["edge_index"][1], p["rev_index"], W64, b64, keep_mask=mask.bool(), p=pr)
    (ref * g.double()).sum().backward()
    assert_close(out, ref.detach(), "dropout out")
    assert_close(hc.grad, h64.grad, "dropout grad h")
    assert_close(Wc.grad, W64.grad, "dropout grad W")
    assert_close(bc.grad, b64.grad, "dropout grad b")


def test_module_dropout_train_eval():
    from notorch_b200 import BatchedGraph
    from notorch_b200.nn import ChempropBlock

    p = oracle_inputs(8, 32, 2, seed=2)
    blk = ChempropBlock(hidden_dim=32, depth=2, dropout=0.5).cuda()
    G = BatchedGraph(p["x_v"].cuda(), p["x_e"].cuda(), p["edge_index"].cuda(), p["rev_index"].cuda(),
                     batch_node_index=p["batch_node_index"].cuda(), batch_edge_index=p["batch_edge_index"].cuda(), size=8)
    blk.eval()
    a, b = blk(G).edge_feats, blk(G).edge_feats
    assert torch.equal(a, b)
    blk.train()
    torch.manual_seed(1)
    c = blk(G).edge_feats
    torch.manual_seed(1)
    c2 = blk(G).edge_feats
    d_ = blk(G).edge_feats
    assert not torch.equal(a, c) and not torch.equal(c, d_)
    # same torch seed -> same dropout seed, but the per-call offset differs: masks are fresh per call
    assert c.shape == c2.shape


def test_standalone_layer_and_residual_match_block():
    from notorch_b200 import BatchedGraph
    from notorch_b200.nn import ChempropBlock

    p = oracle_inputs(8, 48, 1, seed=6)
    blk = ChempropBlock(hidden_dim=48, depth=1).cuda()
    ei, rev = p["edge_index"].cuda(), p["rev_index"].cuda()
    xv, xe = p["x_v"].cuda(), p["x_e"].cuda()
    G = BatchedGraph(xv, xe, ei, rev, batch_node_index=p["batch_node_index"].cuda(), batch_edge_index=p["batch_edge_index"].cuda(), size=8)
    want = blk(G).edge_feats
    h0 = xv[ei[0]] + xe
    got = blk.layers[0](h0, xv, ei, rev)  # Residual(ChempropLayer).forward(*inputs)
    assert torch.equal(got, want)
    inner = blk.layers[0].module(h0, xv, ei, rev)  # bare ChempropLayer: no residual
    assert_close(h0 + inner, want, "h + layer(h)", 1e-6)


def test_unsupported_inputs_raise():
    from notorch_b200 import BatchedGraph, ops
    from notorch_b200.nn import ChempropBlock, Sum
    from notorch_b200.nn.gnn import agg

    p = oracle_inputs(4, 16, 1, seed=1)
    G_cpu = BatchedGraph(p["x_v"], p["x_e"], p["edge_index"], p["rev_index"], batch_node_index=p["batch_node_index"],
                         batch_edge_index=p["batch_edge_index"], size=4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ChempropBlock(hidden_dim=16, depth=1)(G_cpu)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Sum()(G_cpu)
    G = G_cpu.to("cuda")
    with pytest.raises(ValueError, match="unknown reduce"):
        ChempropBlock(hidden_dim=16, depth=1, reduce="prod").cuda()(G)
    with pytest.raises(NotImplementedError):
        ChempropBlock(hidden_dim=16, depth=1, act=torch.nn.Softplus).cuda()(G)
    with pytest.raises(RuntimeError, match="float32 only"):
        ChempropBlock(hidden_dim=16, depth=1).cuda().double()(G.update(node_feats=G.node_feats.double(), edge_feats=G.edge_feats.double()))


def test_from_packed_device_collation_end_to_end():
    from notorch_b200 import BatchedGraph
    from notorch_b200.nn import ChempropBlock, Sum

    p = oracle_inputs(32, 64, 2, seed=8)
    Gd = BatchedGraph.from_packed(p["mols"], p["x_v"], p["x_e"], device="cuda")
    torch.cuda.synchronize()
    for k in ("edge_index", "rev_index", "batch_node_index", "batch_edge_index"):
        assert torch.equal(getattr(Gd, k).cpu(), p[k]), k
    blk = ChempropBlock(hidden_dim=64, depth=2).cuda()
    H = Sum()(blk(Gd))
    Ws = [l.module.update[0].weight.detach().cpu() for l in blk.layers]
    bs = [l.module.update[0].bias.detach().cpu() for l in blk.layers]
    node, _, _ = O.block_forward(p["x_v"], p["x_e"], p["edge_index"], p["rev_index"], Ws, bs)
    assert_close(H, O.readout(node, p["batch_node_index"], 32, "sum"), "H from device collation")


def test_full_size_properties_config2():
    """BASELINE config 2 at full size (B=4096, d=300, L=3): size-independent properties — bitwise
    determinism, and the checksum of checksums sum_b H[b] == sum_v node_out[v]."""
    from notorch_b200 import BatchedGraph
    from notorch_b200.nn import ChempropBlock, Sum
    from notorch_b200.synth import make_molecules

    mols = make_molecules(4096, 2)
    V, E, d = mols.total_atoms, mols.total_edges, 300
    gen = torch.Generator().manual_seed(0)
    xv, xe = torch.randn(V, d, generator=gen), torch.randn(E, d, generator=gen)
    torch.manual_seed(0)
    blk = ChempropBlock(hidden_dim=d, depth=3).cuda()
    Hs = []
    for _ in range(2):
        G = BatchedGraph.from_packed(mols, xv, xe, device="cuda")
        G.node_feats.requires_grad_(True)
        G1 = blk(G)
        H = Sum()(G1)
        H.square().mean().backward()
        Hs.append((H.detach(), G1.node_feats.detach(), G.node_feats.grad.clone()))
    assert torch.equal(Hs[0][0], Hs[1][0]) and torch.equal(Hs[0][2], Hs[1][2])
    total_H, total_nodes = Hs[0][0].double().sum(0), Hs[0][1].double().sum(0)
    assert float((total_H - total_nodes).abs().max()) <= 1e-6 * float(total_nodes.abs().max())
    assert torch.isfinite(Hs[0][0]).all()


@pytest.mark.parametrize("E,d", [(300, 64), (2000, 300), (500, 256)])
def test_wgrad_gather_kernel_without_saved_messages(E, d):
    """K4b has two tensor-core kernels: TMA-streamed (m saved by K2; what autograd uses) and gather-based (m == NULL,
    recomputed from n / h / src / rev). This drives the second one through the C ABI directly."""
    from notorch_b200 import _lib, ops

    gen = torch.Generator().manual_seed(E + d)
    V = max(1, E // 2)
    src, dst, rev = torch.randint(0, V, (E,), generator=gen), torch.randint(0, V, (E,), generator=gen), torch.randint(0, E, (E,), generator=gen)
    h, g = torch.randn(E, d, generator=gen), torch.randn(E, d, generator=gen)
    a = torch.relu(h.double())
    n64 = torch.zeros(V, d, dtype=torch.float64).index_add_(0, dst, a)
    m64 = n64[src] - a[rev]
    csr = ops.build_graph_csr(torch.stack([src, dst]).cuda(), rev.cuda(), V)
    hc, gc = h.cuda(), g.cuda()
    n = ops._seg_reduce_raw(hc, csr.by_dst, _lib.ACT_RELU, 0.0, False)
    L = _lib.lib()
    gW, gb = torch.empty(d, d, device="cuda"), torch.empty(d, device="cuda")
    ws = torch.empty(L.nt_layer_backward_wgrad_workspace_bytes(E, d), dtype=torch.uint8, device="cuda")
    p = lambda t: t.data_ptr()
    _lib.check(L.nt_layer_backward_wgrad(p(gc), None, p(hc), p(n), p(csr.src), p(csr.rev), E, V, d, _lib.ACT_RELU, 0.0, 0.0, 0, 0, p(gW), p(gb),
                                         p(ws), ws.numel(), _lib.NT_F32, _lib.GEMM_TF32X3, torch.cuda.current_stream().cuda_stream), "wgrad")
    assert_close(gW, g.double().t() @ m64, "gW (gather kernel)")
    assert_close(gb, g.double().sum(0), "gb (gather kernel)")


def test_foreign_graph_object_duck_typing():
    """the modules only need node_feats / edge_feats / edge_index / rev_index /
    batch_node_index / update() / len() — i.e. the reference's own BatchedGraph works unchanged."""
    import copy

    from notorch_b200.nn import ChempropBlock, Sum

    class ForeignGraph:  # mimics notorch.data.models.graph.BatchedGraph's surface (graph.py:167-227, utils.py:34-40)
        def __init__(self, **kw):
            self.__dict__.update(kw)

        def update(self, in_place=False, **kw):
            other = self if in_place else copy.copy(self)
            for k, v in kw.items():
                setattr(other, k, v)
            return other

        def __len__(self):
            return self._size

    p = oracle_inputs(8, 32, 2, seed=12)
    G = ForeignGraph(node_feats=p["x_v"].cuda(), edge_feats=p["x_e"].cuda(), edge_index=p["edge_index"].cuda(), rev_index=p["rev_index"].cuda(),
                     batch_node_index=p["batch_node_index"].cuda(), batch_edge_index=p["batch_edge_index"].cuda(), _size=8)
    blk = ChempropBlock(hidden_dim=32, depth=2).cuda()
    G1 = blk(G)
    H = Sum()(G1)
    Ws = [l.module.update[0].weight.detach().cpu() for l in blk.layers]
    bs = [l.module.update[0].bias.detach().cpu() for l in blk.layers]
    node, edge, _ = O.block_forward(p["x_v"], p["x_e"], p["edge_index"], p["rev_index"], Ws, bs)
    assert isinstance(G1, ForeignGraph) and G1.edge_index is G.edge_index
    assert_close(G1.edge_feats, edge, "edge_out (foreign graph)")
    assert_close(H, O.readout(node, p["batch_node_index"], 8, "sum"), "H (foreign graph)")


@pytest.mark.parametrize("d", [300, 37, 64])
def test_graph_embedding_matches_torch_embedding_bag(d):
    """Row N1: GraphEmbedding (embed.py:20-24 = two nn.EmbeddingBag(mode='sum')) forward and table gradients."""
    from notorch_b200 import BatchedGraph
    from notorch_b200.nn import GraphEmbedding

    p = oracle_inputs(64, 8, 0, seed=21)
    gen = torch.Generator().manual_seed(3)
    V, E = p["V"], p["E"]
    nv = torch.randint(0, 45, (V, 7), generator=gen)
    ne = torch.randint(0, 13, (E, 2), generator=gen)
    gv, ge = torch.randn(V, d, generator=gen), torch.randn(E, d, generator=gen)
    torch.manual_seed(0)
    emb = GraphEmbedding(hidden_dim=d)
    ref_v = torch.nn.functional.embedding_bag(nv, emb.node.weight.detach().double().requires_grad_(True), mode="sum")
    wv64 = emb.node.weight.detach().double().requires_grad_(True)
    we64 = emb.edge.weight.detach().double().requires_grad_(True)
    ref_v = torch.nn.functional.embedding_bag(nv, wv64, mode="sum")
    ref_e = torch.nn.functional.embedding_bag(ne, we64, mode="sum")
    ((ref_v * gv.double()).sum() + (ref_e * ge.double()).sum()).backward()
    emb = emb.cuda()
    G = BatchedGraph(nv.cuda(), ne.cuda(), p["edge_index"].cuda(), p["rev_index"].cuda(), batch_node_index=p["batch_node_index"].cuda(),
                     batch_edge_index=p["batch_edge_index"].cuda(), size=64)
    G1 = emb(G)
    ((G1.node_feats * gv.cuda()).sum() + (G1.edge_feats * ge.cuda()).sum()).backward()
    assert_close(G1.node_feats, ref_v.detach(), "node embedding", 1e-6)
    assert_close(G1.edge_feats, ref_e.detach(), "edge embedding", 1e-6)
    assert_close(emb.node.weight.grad, wv64.grad, "grad node table")
    assert_close(emb.edge.weight.grad, we64.grad, "grad edge table")
    assert list(emb.state_dict()) == ["node.weight", "edge.weight"]
    with pytest.raises(IndexError):
        emb(G.update(node_feats=nv.cuda() + 45)).node_feats  # the look-up is deferred until the features are read (or fused into K0)


def _mol_graph(x, batch, B, Q=None):
    from notorch_b200 import BatchedGraph

    E0 = torch.zeros(0, x.shape[1], device="cuda")
    return BatchedGraph(x, E0, torch.zeros(2, 0, dtype=torch.long, device="cuda"), torch.zeros(0, dtype=torch.long, device="cuda"),
                        batch_node_index=batch.cuda(), batch_edge_index=torch.zeros(0, dtype=torch.long, device="cuda"), size=B)


def test_max_readout_matches_reference_golden():
    """Row N2: agg.Max (agg.py:41-47) vs the reference's own output, incl. empty molecules and tied maxima (bit-exact)."""
    import os

    from notorch_b200.nn import Max

    z = np.load(os.path.join(os.path.dirname(__file__), "golden_readouts", "readout_max.npz"))
    x = torch.from_numpy(z["x"]).cuda().requires_grad_(True)
    H = Max()(_mol_graph(x, torch.from_numpy(z["batch_node_index"]), 7))
    (H * torch.from_numpy(z["gH"]).cuda()).sum().backward()
    assert torch.equal(H.cpu(), torch.from_numpy(z["H"]))
    assert torch.equal(x.grad.cpu(), torch.from_numpy(z["g_x"]))


@pytest.mark.parametrize("B,d", [(16, 300), (5, 37)])
def test_gated_and_sdpa_readouts_vs_restatement(B, d):
    """Row N2: Gated / SDPAttention with their intended semantics (both are broken in the reference, see agg.py docstrings
    here) vs the builder's fp64 restatement, forward and all gradients."""
    from notorch_b200.nn import Gated, SDPAttention

    p = oracle_inputs(B, d, 0, seed=31)
    V, batch = p["V"], p["batch_node_index"]
    gen = torch.Generator().manual_seed(2)
    gH, Q = torch.randn(B, d, generator=gen), torch.randn(B, d, generator=gen)
    x = p["x_v"]

    torch.manual_seed(0)
    gate = Gated(d)
    w64, b64 = gate.a.weight.detach().double().requires_grad_(True), gate.a.bias.detach().double().requires_grad_(True)
    x64 = x.double().requires_grad_(True)
    ref = O.readout_gated(x64, w64, b64, batch, B)
    (ref * gH.double()).sum().backward()
    gate = gate.cuda()
    xc = x.cuda().requires_grad_(True)
    H = gate(_mol_graph(xc, batch, B))
    (H * gH.cuda()).sum().backward()
    assert_close(H, ref.detach(), "Gated H")
    assert_close(xc.grad, x64.grad, "Gated grad x")
    assert_close(gate.a.weight.grad, w64.grad, "Gated grad weight")
    # softmax is shift invariant: the bias gradient is exactly 0 in exact arithmetic
    assert float(gate.a.bias.grad.abs().max()) <= 1e-5 * float(w64.grad.abs().max()) and float(b64.grad.abs().max()) < 1e-12

    x64 = x.double().requires_grad_(True)
    Q64 = Q.double().requires_grad_(True)
    ref = O.readout_sdpa(x64, Q64, batch, B, d)
    (ref * gH.double()).sum().backward()
    xc, Qc = x.cuda().requires_grad_(True), Q.cuda().requires_grad_(True)
    H = SDPAttention(d)(_mol_graph(xc, batch, B), Q=Qc)
    (H * gH.cuda()).sum().backward()
    assert_close(H, ref.detach(), "SDPA H")
    assert_close(xc.grad, x64.grad, "SDPA grad x")
    assert_close(Qc.grad, Q64.grad, "SDPA grad Q")
