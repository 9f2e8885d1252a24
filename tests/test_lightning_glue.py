"""Row N4: the container-testable half of the Lightning glue — batch transfer that keeps the CSR caches, and the
reference's TensorDictModule calling conventions (lightning_models/model.py:160-166,212,221-222; agg.py:72-78) driven through a
TensorDict-free double."""
from __future__ import annotations

import collections

import pytest
import torch

from helpers import assert_close, oracle_inputs
from oracle import dmpnn_oracle as O


def _graph(p, device="cpu"):
    from notorch_b200 import BatchedGraph

    return BatchedGraph(p["x_v"].to(device), p["x_e"].to(device), p["edge_index"].to(device), p["rev_index"].to(device),
                        batch_node_index=p["batch_node_index"].to(device), batch_edge_index=p["batch_edge_index"].to(device), size=p["B"])


def test_transfer_batch_walks_containers_on_cpu():
    from notorch_b200 import BatchedGraph
    from notorch_b200.lightning_models import transfer_batch_to_device

    p = oracle_inputs(4, 8, 0, seed=1)
    Pair = collections.namedtuple("Pair", "a b")
    batch = {"inputs.G": _graph(p), "targets.y": torch.ones(4, 1), "meta": ["smiles", 3], "nested": {"w": torch.zeros(2)},
             "pair": Pair(torch.zeros(1), "x"), "tup": (torch.ones(2), None)}
    out = transfer_batch_to_device(batch, "cpu", 0)
    assert isinstance(out["inputs.G"], BatchedGraph) and out["inputs.G"] is batch["inputs.G"]  # .to() returns the same object
    assert out["meta"] == ["smiles", 3] and out["nested"]["w"].device.type == "cpu"
    assert isinstance(out["pair"], Pair) and out["pair"].b == "x" and isinstance(out["tup"], tuple) and out["tup"][1] is None

    class TD:  # TensorDict-like: items() + item assignment, not a Mapping
        def __init__(self):
            self.d = {"inputs.G": _graph(p), "x": torch.ones(3)}

        def items(self):
            return self.d.items()

        def __setitem__(self, k, v):
            self.d[k] = v

    td = TD()
    assert transfer_batch_to_device(td, torch.device("cpu")) is td


def test_tensordict_module_lite_calling_conventions():
    from notorch_b200.lightning_models import TensorDictModuleLite, TensorDictSequentialLite

    class AddMul(torch.nn.Module):
        def forward(self, x, y, *, scale=1.0):
            return (x + y) * scale, x - y

    pos = TensorDictModuleLite(AddMul(), ["a", "b"], ["m.sum", "m.diff"])
    kw = TensorDictModuleLite(AddMul(), {"m.sum": "x", "b": "y", "s": "scale"}, ["k.sum", "k.diff"])
    seq = TensorDictSequentialLite(pos, kw, selected_out_keys=["k.sum"])
    td = {"a": torch.tensor(3.0), "b": torch.tensor(1.0), "s": 2.0}
    out = seq(td)
    assert float(out["k.sum"]) == 10.0 and "m.sum" not in out and "k.diff" not in out and "a" in out
    assert "m.sum" not in td  # the input mapping is not mutated
    with pytest.raises(RuntimeError, match="out_keys"):
        TensorDictModuleLite(AddMul(), ["a", "b"], ["only_one"])(td)


@pytest.mark.gpu
def test_model_shell_double_runs_the_drop_in_modules_positionally_and_by_keyword():
    """What NotorchModel.forward does (model.py:159-166,212,221-222): GraphEmbedding -> ChempropBlock -> Sum (positional) and
    SDPAttention(G, Q=...) (keyword) -> MLP head, on a batch dict moved by transfer_batch_to_device; the CSR bundle is built once at
    transfer time and every module re-uses it (no further CSR kernel); the result equals the oracle."""
    from notorch_b200 import _lib, ops
    from notorch_b200.lightning_models import TensorDictModuleLite, TensorDictSequentialLite, transfer_batch_to_device
    from notorch_b200.nn import MLP, ChempropBlock, GraphEmbedding, SDPAttention, Sum

    ops.set_index_validation("sync")
    B, d = 12, 64
    p = oracle_inputs(B, d, 2, seed=4)
    gen = torch.Generator().manual_seed(2)
    p["x_v"], p["x_e"] = torch.randint(0, 45, (p["V"], 7), generator=gen), torch.randint(0, 13, (p["E"], 2), generator=gen)
    torch.manual_seed(0)
    embed, block, head = GraphEmbedding(hidden_dim=d).cuda(), ChempropBlock(hidden_dim=d, depth=2).cuda(), MLP(d, 1, hidden_dim=32).cuda()
    model = TensorDictSequentialLite(
        TensorDictModuleLite(embed, ["inputs.G"], ["embed.G"]),
        TensorDictModuleLite(block, ["embed.G"], ["encoder.G"]),
        TensorDictModuleLite(Sum(), ["encoder.G"], ["agg.H"]),
        TensorDictModuleLite(SDPAttention(d), {"encoder.G": "G", "inputs.Q": "Q"}, ["attn.H"]),
        TensorDictModuleLite(head, ["agg.H"], ["preds.y"]),
    )
    batch = {"inputs.G": _graph(p), "inputs.Q": torch.randn(B, d, generator=gen), "targets.y": torch.zeros(B, 1)}
    batch = transfer_batch_to_device(batch, torch.device("cuda", 0), 0)
    G = batch["inputs.G"]
    assert G.edge_index.is_cuda and getattr(G, "_nt_csr", None) is not None  # prepared at transfer time
    csr = G._nt_csr
    n_csr = []
    real = ops.build_graph_csr
    ops.build_graph_csr = lambda *a, **k: (n_csr.append(1), real(*a, **k))[1]
    try:
        out = model(batch)
    finally:
        ops.build_graph_csr = real
    assert not n_csr and out["encoder.G"]._nt_csr is csr  # one bundle for the whole forward
    loss = torch.nn.functional.mse_loss(out["preds.y"], batch["targets.y"]) + out["attn.H"].square().mean()
    loss.backward()
    assert embed.node.weight.grad is not None and block.layers[0].module.update[0].weight.grad is not None

    # oracle: embedding (torch, CPU) -> block -> sum -> head (torch Linear with the same parameters)
    F = torch.nn.functional
    xv = F.embedding_bag(p["x_v"], embed.node.weight.detach().cpu(), mode="sum")
    xe = F.embedding_bag(p["x_e"], embed.edge.weight.detach().cpu(), mode="sum")
    Ws = [l.module.update[0].weight.detach().cpu() for l in block.layers]
    bs = [l.module.update[0].bias.detach().cpu() for l in block.layers]
    node, _, _ = O.block_forward(xv, xe, p["edge_index"], p["rev_index"], Ws, bs)
    H = O.readout(node, p["batch_node_index"], B, "sum")
    assert_close(out["agg.H"], H, "H through the model shell")
    y = H
    for m in head:
        y = F.linear(y, m.weight.detach().cpu(), m.bias.detach().cpu()) if isinstance(m, torch.nn.Linear) else m.cpu()(y)
    assert_close(out["preds.y"], y, "prediction through the model shell", 1e-5)
    assert _lib.lib().nt_kernel_launch_count() > 0
