"""TEST INFRASTRUCTURE ONLY — generate ``tests/golden/*.npz`` by running the UNMODIFIED reference.

Run in the authoring container (where ``/root/reference`` exists)::

    python oracle/make_golden.py

For each case it builds seeded synthetic molecules (``notorch_b200.synth``), wraps them in the
reference's own ``Graph`` objects, collates them with the reference's ``BatchedGraph.from_graphs``
(``notorch/data/models/graph.py:186-223``), runs the reference's ``ChempropBlock`` and
``agg.Sum`` / ``agg.Mean`` (``notorch/nn/gnn/chemprop.py``, ``notorch/nn/gnn/agg.py``) forward and
backward on CPU in fp32 and again in fp64, and stores inputs, parameters, cotangents, outputs and
gradients. The fixtures travel to the GPU box; the reference does not.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from notorch_b200.synth import MolSpec, make_molecules  # noqa: E402
from oracle import reference_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

ACT_MODULES = {
    "relu": nn.ReLU, "leaky_relu": nn.LeakyReLU, "elu": nn.ELU, "silu": nn.SiLU,
    "gelu": nn.GELU, "tanh": nn.Tanh, "identity": nn.Identity,
}

SMALL = MolSpec(9.0, 3.0, 2, 16)

# name -> dict(batch, spec, d, block kwargs, agg, extras)
CASES = {
    "base_d24": dict(batch=6, d=24, depth=3),
    "headline_d300": dict(batch=3, d=300, depth=2, spec=MolSpec(12.0, 2.0, 8, 16), skip_f64_param_grads=True),
    "mean_reduce": dict(batch=5, d=16, depth=2, reduce="mean", agg="mean", bondless_every=3),
    "no_residual": dict(batch=4, d=20, depth=2, residual=False),
    "no_bias": dict(batch=4, d=20, depth=2, bias=False),
    "shared": dict(batch=4, d=12, depth=3, shared=True),
    "depth0": dict(batch=3, d=8, depth=0),
    "single_mol": dict(batch=1, d=32, depth=2),
    "bondless": dict(batch=6, d=8, depth=2, bondless_every=2, agg="mean"),
    "adversarial_rev": dict(batch=4, d=16, depth=2, adversarial_rev=True),
    "act_silu": dict(batch=3, d=16, depth=2, act="silu"),
    "act_tanh": dict(batch=3, d=16, depth=2, act="tanh"),
    "act_elu": dict(batch=3, d=16, depth=2, act="elu"),
    "act_gelu": dict(batch=3, d=16, depth=2, act="gelu"),
    "act_leaky": dict(batch=3, d=16, depth=2, act="leaky_relu"),
    "odd_d37": dict(batch=4, d=37, depth=2),
    # torch_scatter's arg-reductions inside the block (chemprop.py:39,86 with reduce in {"max","min"}); bondless molecules give
    # empty segments (value 0, no gradient)
    "max_reduce": dict(batch=5, d=16, depth=2, reduce="max", bondless_every=3),
    "min_reduce": dict(batch=5, d=20, depth=2, reduce="min", act="tanh", bondless_every=4),
}


def run_case(name: str, cfg: dict, seed: int) -> dict[str, np.ndarray]:
    ref = reference_loader.load()
    batch, d, depth = cfg["batch"], cfg["d"], cfg["depth"]
    spec = cfg.get("spec", SMALL)
    mols = make_molecules(batch, spec, seed=seed, bondless_every=cfg.get("bondless_every", 0))
    V, E = mols.total_atoms, mols.total_edges
    g = torch.Generator().manual_seed(seed)

    if cfg.get("adversarial_rev"):
        rng = np.random.default_rng(seed)
        # arbitrary in-range local rev (non-involutive, non-injective) — SURVEY.md §0 item 3
        offs = np.concatenate([[0], np.cumsum(mols.num_edges)])
        for i in range(batch):
            e = int(mols.num_edges[i])
            if e:
                mols.rev_index[offs[i]:offs[i + 1]] = rng.integers(0, e, size=e)

    # per-molecule reference Graphs with (dummy) integer type features, collated by the reference
    graphs = []
    for n, ei, rev in mols.split():
        ei_t = torch.from_numpy(ei.astype(np.int64)) if ei.shape[1] else torch.empty(0)
        graphs.append(ref.Graph(torch.zeros(n, 1, dtype=torch.long), torch.zeros(len(rev), 1, dtype=torch.long),
                                ei_t, torch.from_numpy(rev.astype(np.int64))))
    G0 = ref.BatchedGraph.from_graphs(graphs)
    assert G0.edge_index.shape == (2, E) and len(G0.batch_node_index) == V

    x_v = torch.randn(V, d, generator=g)
    x_e = torch.randn(E, d, generator=g)
    gH = torch.randn(batch, d, generator=g)
    gE = torch.randn(E, d, generator=g)
    gN = torch.randn(V, d, generator=g)

    kw = dict(hidden_dim=d, act=ACT_MODULES[cfg.get("act", "relu")], bias=cfg.get("bias", True),
              dropout=0.0, depth=depth, residual=cfg.get("residual", True),
              shared=cfg.get("shared", False), reduce=cfg.get("reduce", "sum"))
    torch.manual_seed(seed)
    block = ref.ChempropBlock(**kw)
    agg = {"sum": ref.Sum, "mean": ref.Mean}[cfg.get("agg", "sum")]()

    out: dict[str, np.ndarray] = {}
    meta = {k: v for k, v in cfg.items() if k != "spec"}
    meta.update(name=name, seed=seed, V=V, E=E, state_keys=list(block.state_dict().keys()))
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    out["num_atoms"], out["num_edges"] = mols.num_atoms, mols.num_edges
    out["local_edge_index"], out["local_rev_index"] = mols.edge_index, mols.rev_index
    out["edge_index"] = G0.edge_index.numpy()
    out["rev_index"] = G0.rev_index.numpy()
    out["batch_node_index"] = G0.batch_node_index.numpy()
    out["batch_edge_index"] = G0.batch_edge_index.numpy()
    out["x_v"], out["x_e"] = x_v.numpy(), x_e.numpy()
    out["gH"], out["gE"], out["gN"] = gH.numpy(), gE.numpy(), gN.numpy()
    for k, v in block.state_dict().items():
        out["param/" + k] = v.numpy().copy()

    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        blk = ref.ChempropBlock(**kw).to(dt)
        blk.load_state_dict({k: v.to(dt) for k, v in block.state_dict().items()})
        xv = x_v.detach().clone().to(dt).requires_grad_(True)
        xe = x_e.detach().clone().to(dt).requires_grad_(True)
        G = ref.BatchedGraph(xv, xe, G0.edge_index, G0.rev_index, batch_node_index=G0.batch_node_index,
                             batch_edge_index=G0.batch_edge_index, size=batch)
        G1 = blk(G)
        H = agg(G1)
        # cotangents on all three outputs
        loss = (H * gH.to(dt)).sum() + (G1.edge_feats * gE.to(dt)).sum() + (G1.node_feats * gN.to(dt)).sum()
        loss.backward()
        out[f"{tag}/node_out"] = G1.node_feats.detach().numpy()
        out[f"{tag}/edge_out"] = G1.edge_feats.detach().numpy()
        out[f"{tag}/H"] = H.detach().numpy()
        out[f"{tag}/g_x_v"] = xv.grad.numpy()
        out[f"{tag}/g_x_e"] = xe.grad.numpy()
        seen = set()
        for k, p_ in blk.named_parameters():
            if id(p_) in seen:
                continue
            seen.add(id(p_))
            if tag == "f64" and cfg.get("skip_f64_param_grads"):
                continue  # keeps the d=300 fixture small
            out[f"{tag}/grad/" + k] = (p_.grad if p_.grad is not None else torch.zeros_like(p_)).numpy()
    return out


def run_readout_max(seed: int) -> dict[str, np.ndarray]:
    """``agg.Max`` of the reference (agg.py:41-47) incl. an empty molecule and tied maxima, forward + backward."""
    ref = reference_loader.load()
    g = torch.Generator().manual_seed(seed)
    B, d = 7, 24
    counts = [5, 1, 0, 9, 3, 0, 6]  # molecules 2 and 5 own no atom
    batch = torch.repeat_interleave(torch.arange(B), torch.tensor(counts))
    V = int(batch.numel())
    x = torch.randn(V, d, generator=g)
    x[0, :4] = x[1, :4] = 3.0  # ties: the first maximum wins
    gH = torch.randn(B, d, generator=g)
    xr = x.clone().requires_grad_(True)
    G = ref.BatchedGraph(xr, torch.zeros(0, d), torch.zeros(2, 0, dtype=torch.long), torch.zeros(0, dtype=torch.long),
                         batch_node_index=batch, batch_edge_index=torch.zeros(0, dtype=torch.long), size=B)
    H = ref.agg.Max()(G)
    (H * gH).sum().backward()
    meta = dict(name="readout_max", seed=seed, B=B, d=d, V=V)
    return {"meta": np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), "x": x.numpy(), "batch_node_index": batch.numpy(),
            "gH": gH.numpy(), "H": H.detach().numpy(), "g_x": xr.grad.numpy()}


def run_mlp_head(seed: int) -> dict[str, np.ndarray]:
    """The reference's ``MLP`` factory (nn/mlp.py:9-68) on molecule vectors: two hidden layers, SiLU, unflattened ``(2, 3)`` output;
    forward + all gradients in fp32 and fp64."""
    import importlib

    reference_loader.load()
    RefMLP = importlib.import_module("notorch.nn.mlp").MLP
    g = torch.Generator().manual_seed(seed)
    B, d = 37, 52
    kw = dict(input_dim=d, output_size=(2, 3), hidden_dim=24, num_layers=2, dropout=0.0)
    torch.manual_seed(seed)
    mlp = RefMLP(activation=nn.SiLU, **kw)
    x = torch.randn(B, d, generator=g)
    gY = torch.randn(B, 2, 3, generator=g)
    out = {"meta": np.frombuffer(json.dumps(dict(kw, output_size=[2, 3], activation="silu", seed=seed)).encode(), dtype=np.uint8),
           "x": x.numpy(), "gY": gY.numpy()}
    for k, v in mlp.state_dict().items():
        out["param/" + k] = v.numpy().copy()
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        m = RefMLP(activation=nn.SiLU, **kw).to(dt)
        m.load_state_dict({k: v.to(dt) for k, v in mlp.state_dict().items()})
        xr = x.detach().clone().to(dt).requires_grad_(True)
        y = m(xr)
        (y * gY.to(dt)).sum().backward()
        out[f"{tag}/y"], out[f"{tag}/g_x"] = y.detach().numpy(), xr.grad.numpy()
        for k, p_ in m.named_parameters():
            out[f"{tag}/grad/" + k] = p_.grad.numpy()
    return out


EXTRAS = {"readout_max": (run_readout_max, 9001), "mlp_head": (run_mlp_head, 9002)}


def main() -> None:
    if not reference_loader.available():
        raise SystemExit("reference tree not found; golden fixtures can only be made in the authoring container")
    os.makedirs(OUT, exist_ok=True)
    total = 0
    only = set(sys.argv[1:])  # optional: regenerate just the named cases (seeds depend on the position in CASES, not on the selection)
    for i, (name, cfg) in enumerate(CASES.items()):
        if only and name not in only:
            continue
        data = run_case(name, cfg, seed=4242 + i)
        path = os.path.join(OUT, f"{name}.npz")
        np.savez_compressed(path, **data)
        total += os.path.getsize(path)
        print(f"{name:18s} V={int(data['num_atoms'].sum()):4d} E={int(data['num_edges'].sum()):4d} "
              f"{os.path.getsize(path) / 1024:.0f} KiB")
    os.makedirs(os.path.join(OUT, "..", "golden_readouts"), exist_ok=True)
    for name, (fn, seed) in EXTRAS.items():
        if only and name not in only:
            continue
        np.savez_compressed(os.path.join(OUT, "..", "golden_readouts", f"{name}.npz"), **fn(seed=seed))
        print(f"{name:18s} -> tests/golden_readouts/{name}.npz")
    print(f"total {total / 1024:.0f} KiB -> {OUT}")


if __name__ == "__main__":
    main()
