"""Shared helpers for the parity tests."""
from __future__ import annotations

import numpy as np
import torch

ACT_MODULES = {
    "relu": torch.nn.ReLU, "leaky_relu": torch.nn.LeakyReLU, "elu": torch.nn.ELU, "silu": torch.nn.SiLU,
    "gelu": torch.nn.GELU, "tanh": torch.nn.Tanh, "identity": torch.nn.Identity,
}

# fp32 parity bound of BASELINE.json's north_star: rel 1e-5, measured per tensor as
# max|x - x_ref| <= REL * max|x_ref|.
REL_F32 = 1e-5


def rel_err(x: torch.Tensor, ref: torch.Tensor) -> float:
    x, ref = x.detach().double().cpu(), ref.detach().double().cpu()
    denom = float(ref.abs().max())
    if denom == 0.0:
        return float((x - ref).abs().max())
    return float((x - ref).abs().max()) / denom


def assert_close(x, ref, what: str, rel: float = REL_F32):
    if isinstance(ref, np.ndarray):
        ref = torch.from_numpy(ref)
    assert tuple(x.shape) == tuple(ref.shape), f"{what}: shape {tuple(x.shape)} vs {tuple(ref.shape)}"
    err = rel_err(x, ref)
    assert err <= rel, f"{what}: rel-to-max error {err:.3e} > {rel:.1e}"


def block_from_golden(g, device="cuda"):
    """Our ChempropBlock with the reference's parameters loaded strict=True."""
    from notorch_b200.nn import ChempropBlock

    m = g.meta
    blk = ChempropBlock(hidden_dim=m["d"], act=ACT_MODULES[m.get("act", "relu")], bias=m.get("bias", True), dropout=0.0,
                        depth=m["depth"], residual=m.get("residual", True), shared=m.get("shared", False),
                        reduce=m.get("reduce", "sum"))
    sd = {k: torch.from_numpy(v) for k, v in g.params().items()}
    blk.load_state_dict(sd, strict=True)
    return blk.to(device)


def graph_from_golden(g, device="cuda", requires_grad=True):
    from notorch_b200 import BatchedGraph

    xv = torch.from_numpy(g["x_v"]).to(device).requires_grad_(requires_grad)
    xe = torch.from_numpy(g["x_e"]).to(device).requires_grad_(requires_grad)
    G = BatchedGraph(xv, xe, torch.from_numpy(g["edge_index"]).to(device), torch.from_numpy(g["rev_index"]).to(device),
                     batch_node_index=torch.from_numpy(g["batch_node_index"]).to(device),
                     batch_edge_index=torch.from_numpy(g["batch_edge_index"]).to(device), size=len(g["num_atoms"]))
    return G, xv, xe


def oracle_inputs(batch, d, depth, config=1, seed=0, bias=True, dtype=torch.float32):
    """Seeded synthetic problem: packed molecules, collated indices (oracle), features, weights."""
    from notorch_b200.synth import make_molecules
    from oracle import dmpnn_oracle as O

    mols = make_molecules(batch, config, seed=seed)
    c = O.collate(mols.split())
    gen = torch.Generator().manual_seed(seed)
    V, E = mols.total_atoms, mols.total_edges
    out = dict(mols=mols, V=V, E=E, B=batch, d=d,
               edge_index=torch.from_numpy(c["edge_index"]), rev_index=torch.from_numpy(c["rev_index"]),
               batch_node_index=torch.from_numpy(c["batch_node_index"]), batch_edge_index=torch.from_numpy(c["batch_edge_index"]),
               x_v=torch.randn(V, d, generator=gen, dtype=dtype), x_e=torch.randn(E, d, generator=gen, dtype=dtype))
    bound = 1.0 / d ** 0.5  # nn.Linear default init range
    out["weights"] = [(torch.rand(d, d, generator=gen, dtype=dtype) * 2 - 1) * bound for _ in range(depth)]
    out["biases"] = [((torch.rand(d, generator=gen, dtype=dtype) * 2 - 1) * bound) if bias else None for _ in range(depth)]
    return out
