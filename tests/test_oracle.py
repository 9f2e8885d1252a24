"""CPU: pin the oracle (``oracle/dmpnn_oracle.py``) against the golden vectors the *reference
itself* produced (``oracle/make_golden.py``), and against the live reference when it is mounted."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import dmpnn_oracle as O
from oracle import reference_loader


def _layers(g, dt):
    meta = g.meta
    params = g.params()
    mid = "module." if meta.get("residual", True) else ""
    Ws, bs = [], []
    for i in range(meta["depth"]):
        Ws.append(torch.from_numpy(params[f"layers.{i}.{mid}update.0.weight"]).to(dt))
        bk = f"layers.{i}.{mid}update.0.bias"
        bs.append(torch.from_numpy(params[bk]).to(dt) if bk in params else None)
    return Ws, bs


def test_collate_matches_reference_golden(golden):
    out = O.collate(golden.mols())
    for k in ("edge_index", "rev_index", "batch_node_index", "batch_edge_index"):
        assert out[k].dtype == np.int64
        assert np.array_equal(out[k], golden[k]), k
    assert out["size"] == len(golden["num_atoms"])


def test_rev_index_is_node_offset_quirk():
    """graph.py:199-200 adds the *node* offset to rev_index; keep that visible in a test."""
    from conftest import load_golden

    g = load_golden("base_d24")
    node_off = np.concatenate([[0], np.cumsum(g["num_atoms"])])[:-1]
    edge_mol = np.repeat(np.arange(len(g["num_atoms"])), g["num_edges"])
    assert np.array_equal(g["rev_index"], g["local_rev_index"] + node_off[edge_mol])
    fixed = O.collate_fixed(g.mols())["rev_index"]
    assert not np.array_equal(fixed, g["rev_index"])
    assert np.array_equal(fixed[fixed], np.arange(len(fixed)))


@pytest.mark.parametrize("tag,dt", [("f32", torch.float32), ("f64", torch.float64)])
def test_forward_matches_reference_golden(golden, tag, dt):
    meta = golden.meta
    Ws, bs = _layers(golden, dt)
    ei = torch.from_numpy(golden["edge_index"])
    rev = torch.from_numpy(golden["rev_index"])
    bni = torch.from_numpy(golden["batch_node_index"])
    node_out, edge_out, _ = O.block_forward(
        torch.from_numpy(golden["x_v"]).to(dt), torch.from_numpy(golden["x_e"]).to(dt), ei, rev, Ws, bs,
        act=meta.get("act", "relu"), reduce=meta.get("reduce", "sum"), residual=meta.get("residual", True))
    H = O.readout(node_out, bni, len(golden["num_atoms"]), meta.get("agg", "sum"))
    # same ATen ops in the same order as the reference => bit-identical on CPU
    assert torch.equal(node_out, torch.from_numpy(golden[f"{tag}/node_out"]))
    assert torch.equal(edge_out, torch.from_numpy(golden[f"{tag}/edge_out"]))
    assert torch.equal(H, torch.from_numpy(golden[f"{tag}/H"]))


def test_manual_backward_matches_reference_golden(golden):
    """The hand-derived backward (what the CUDA kernels implement) vs the reference's autograd, fp64."""
    dt = torch.float64
    meta = golden.meta
    if "f64/grad/" not in "".join(golden.z.files):
        pytest.skip("fixture stores no fp64 parameter grads")
    Ws, bs = _layers(golden, dt)
    ei = torch.from_numpy(golden["edge_index"])
    rev = torch.from_numpy(golden["rev_index"])
    bni = torch.from_numpy(golden["batch_node_index"])
    V = len(bni)
    kind = meta.get("agg", "sum")
    g_node = torch.from_numpy(golden["gN"]).to(dt) + O.readout_backward(
        torch.from_numpy(golden["gH"]).to(dt), bni, V, kind)
    out = O.block_backward(
        torch.from_numpy(golden["x_v"]).to(dt), torch.from_numpy(golden["x_e"]).to(dt), ei, rev, Ws, bs,
        g_node, torch.from_numpy(golden["gE"]).to(dt),
        act=meta.get("act", "relu"), reduce=meta.get("reduce", "sum"), residual=meta.get("residual", True))
    tol = dict(rtol=1e-10, atol=1e-10)
    assert torch.allclose(out["x_v"], torch.from_numpy(golden["f64/g_x_v"]), **tol)
    assert torch.allclose(out["x_e"], torch.from_numpy(golden["f64/g_x_e"]), **tol)
    grads = golden.grads("f64")
    mid = "module." if meta.get("residual", True) else ""
    if meta.get("shared"):
        gW = sum(out["weights"])
        assert torch.allclose(gW, torch.from_numpy(grads[f"layers.0.{mid}update.0.weight"]), **tol)
        return
    for i in range(meta["depth"]):
        assert torch.allclose(out["weights"][i], torch.from_numpy(grads[f"layers.{i}.{mid}update.0.weight"]), **tol)
        if bs[i] is not None:
            assert torch.allclose(out["biases"][i], torch.from_numpy(grads[f"layers.{i}.{mid}update.0.bias"]), **tol)


def test_build_csr_invariants():
    rng = np.random.default_rng(0)
    keys = rng.integers(0, 17, size=200)
    rowptr, perm = O.build_csr(keys, 17)
    assert rowptr[0] == 0 and rowptr[-1] == 200 and np.all(np.diff(rowptr) >= 0)
    assert np.array_equal(np.sort(perm), np.arange(200))
    for s in range(17):
        seg = perm[rowptr[s]:rowptr[s + 1]]
        assert np.all(keys[seg] == s) and np.all(np.diff(seg) > 0)


def test_dropout_mask_path_autograd_consistency():
    """keep-mask dropout: manual backward == autograd of the oracle forward (fp64)."""
    from notorch_b200.synth import MolSpec, make_molecules

    mols = make_molecules(4, MolSpec(8, 2, 3, 12), seed=5)
    c = O.collate(mols.split())
    ei, rev = torch.from_numpy(c["edge_index"]), torch.from_numpy(c["rev_index"])
    V, E, d, p = mols.total_atoms, mols.total_edges, 10, 0.3
    g = torch.Generator().manual_seed(1)
    xv = torch.randn(V, d, generator=g, dtype=torch.float64, requires_grad=True)
    xe = torch.randn(E, d, generator=g, dtype=torch.float64, requires_grad=True)
    Ws = [torch.randn(d, d, generator=g, dtype=torch.float64, requires_grad=True) for _ in range(2)]
    bs = [torch.randn(d, generator=g, dtype=torch.float64, requires_grad=True) for _ in range(2)]
    masks = [torch.rand(E, d, generator=g) > p for _ in range(2)]
    gN = torch.randn(V, d, generator=g, dtype=torch.float64)
    gE = torch.randn(E, d, generator=g, dtype=torch.float64)
    node_out, edge_out, _ = O.block_forward(xv, xe, ei, rev, Ws, bs, keep_masks=masks, p=p, reduce="mean")
    ((node_out * gN).sum() + (edge_out * gE).sum()).backward()
    out = O.block_backward(xv.detach(), xe.detach(), ei, rev, [w.detach() for w in Ws], [b.detach() for b in bs],
                           gN, gE, keep_masks=masks, p=p, reduce="mean")
    assert torch.allclose(out["x_v"], xv.grad, atol=1e-10)
    assert torch.allclose(out["x_e"], xe.grad, atol=1e-10)
    for i in range(2):
        assert torch.allclose(out["weights"][i], Ws[i].grad, atol=1e-10)
        assert torch.allclose(out["biases"][i], bs[i].grad, atol=1e-10)


@pytest.mark.reference
@pytest.mark.skipif(not reference_loader.available(), reason="live reference tree not mounted")
def test_oracle_matches_live_reference_fresh_seed():
    """Fresh seed, not in the fixtures: oracle == reference code executed now (bit-exact fp32)."""
    from notorch_b200.synth import make_molecules

    ref = reference_loader.load()
    mols = make_molecules(16, 1, seed=777)
    graphs = [ref.Graph(torch.zeros(n, 1, dtype=torch.long), torch.zeros(len(rev), 1, dtype=torch.long),
                        torch.from_numpy(ei.astype(np.int64)), torch.from_numpy(rev.astype(np.int64)))
              for n, ei, rev in mols.split()]
    G0 = ref.BatchedGraph.from_graphs(graphs)
    c = O.collate(mols.split())
    assert np.array_equal(c["edge_index"], G0.edge_index.numpy())
    assert np.array_equal(c["rev_index"], G0.rev_index.numpy())
    d = 48
    torch.manual_seed(3)
    blk = ref.ChempropBlock(hidden_dim=d, depth=3)
    xv, xe = torch.randn(mols.total_atoms, d), torch.randn(mols.total_edges, d)
    G = ref.BatchedGraph(xv, xe, G0.edge_index, G0.rev_index, batch_node_index=G0.batch_node_index,
                         batch_edge_index=G0.batch_edge_index, size=16)
    G1 = blk(G)
    Ws = [l.module.update[0].weight.detach() for l in blk.layers]
    bs = [l.module.update[0].bias.detach() for l in blk.layers]
    node_out, edge_out, _ = O.block_forward(xv, xe, G0.edge_index, G0.rev_index, Ws, bs)
    assert torch.equal(node_out, G1.node_feats) and torch.equal(edge_out, G1.edge_feats)
    assert torch.equal(O.readout(node_out, G0.batch_node_index, 16, "mean"), ref.Mean()(G1))


def test_readout_max_matches_reference_golden():
    import os

    z = np.load(os.path.join(os.path.dirname(__file__), "golden_readouts", "readout_max.npz"))
    H = O.readout_max(torch.from_numpy(z["x"]), torch.from_numpy(z["batch_node_index"]), 7)
    assert torch.equal(H, torch.from_numpy(z["H"]))
    assert (H[2] == 0).all() and (H[5] == 0).all()  # empty molecules give 0 (torch_scatter.scatter_max)


@pytest.mark.parametrize("reduce", ["max", "min"])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_seg_extreme_restatement_vs_bruteforce(reduce, seed):
    """The vectorised arg-reduction of the oracle against a literal loop over rows in ascending order (torch-scatter's CPU kernel:
    strict comparison, so the first extreme wins; untouched outputs become 0 with argument len(x))."""
    gen = torch.Generator().manual_seed(seed)
    n, S, d = 57, 9, 5
    index = torch.randint(0, S - 2, (n,), generator=gen)  # segments S-2, S-1 stay empty
    x = torch.randint(-2, 3, (n, d), generator=gen).double()  # many ties
    x[:, -1] = torch.randn(n, generator=gen, dtype=torch.float64)
    val, arg = O.seg_extreme(x, index, S, reduce)
    want_v, want_a = torch.zeros(S, d, dtype=torch.float64), torch.full((S, d), n, dtype=torch.long)
    better = (lambda a, b: a > b) if reduce == "max" else (lambda a, b: a < b)
    for r in range(n):
        s = int(index[r])
        for c in range(d):
            if want_a[s, c] == n or better(float(x[r, c]), float(want_v[s, c])):
                want_v[s, c], want_a[s, c] = x[r, c], r
    assert torch.equal(val, want_v) and torch.equal(arg, want_a)
    assert torch.equal(O.seg_reduce(x, index, S, reduce), want_v)


def test_constructor_surface_matches_live_reference():
    """hydra builds the reference's modules from constructor kwargs (cli/train.py:24-26), so the drop-in classes must take exactly the
    same arguments with the same defaults. Compared against the live reference when it is mounted (authoring container only)."""
    import importlib
    import inspect

    from oracle import reference_loader

    if not reference_loader.available():
        pytest.skip("reference tree not mounted")
    reference_loader.load()
    ours = importlib.import_module("notorch_b200.nn")

    def surface(fn):
        return [(n, p.default) for n, p in inspect.signature(fn).parameters.items() if n != "self"]

    for module, names in (("notorch.nn.gnn.chemprop", ["ChempropLayer", "ChempropBlock"]),
                          ("notorch.nn.gnn.agg", ["Sum", "Mean", "Max", "Gated", "SDPAttention"]), ("notorch.nn.residual", ["Residual"])):
        ref = importlib.import_module(module)
        for name in names:
            assert surface(getattr(ref, name).__init__) == surface(getattr(ours, name).__init__), name
    assert surface(importlib.import_module("notorch.nn.mlp").MLP) == surface(ours.MLP)
