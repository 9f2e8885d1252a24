"""Diagnostic: error of each GEMM mode for one layer, and relu-mask flips through a block."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from notorch_b200 import ops, BatchedGraph
from oracle import dmpnn_oracle as O
from helpers import oracle_inputs, rel_err

def one_layer(E, d, mode):
    ops.set_gemm_mode(mode)
    gen = torch.Generator().manual_seed(E * 7 + d)
    V = max(1, E // 2)
    src = torch.randint(0, V, (E,), generator=gen); dst = torch.randint(0, V, (E,), generator=gen); rev = torch.randint(0, E, (E,), generator=gen)
    h = torch.randn(E, d, generator=gen); W = (torch.rand(d, d, generator=gen) * 2 - 1) / d ** 0.5
    b = torch.randn(d, generator=gen) * 0.1; g = torch.randn(E, d, generator=gen)
    h64, W64, b64 = h.double().requires_grad_(True), W.double().requires_grad_(True), b.double().requires_grad_(True)
    ref, _ = O.layer_forward(h64, V, src, dst, rev, W64, b64, residual=True)
    (ref * g.double()).sum().backward()
    csr = ops.build_graph_csr(torch.stack([src, dst]).cuda(), rev.cuda(), V)
    hc, Wc, bc = h.cuda().requires_grad_(True), W.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    out = ops.layer(hc, Wc, bc, csr, residual=True)
    (out * g.cuda()).sum().backward()
    return rel_err(out, ref), rel_err(hc.grad, h64.grad), rel_err(Wc.grad, W64.grad)

for d in (64, 256, 300, 1024):
    for mode in ("fp32", "tf32x3", "tf32"):
        print(f"layer E=4000 d={d} {mode:7s} out/gh/gW rel err: " + " ".join(f"{x:.2e}" for x in one_layer(4000, d, mode)))

# block: per-layer error and relu flips
for mode in ("fp32", "tf32x3"):
    ops.set_gemm_mode(mode)
    p = oracle_inputs(64, 300, 3, config=1, seed=164)
    node64, edge64, hs64 = O.block_forward(p["x_v"].double(), p["x_e"].double(), p["edge_index"], p["rev_index"], [w.double() for w in p["weights"]], [b.double() for b in p["biases"]])
    node32, edge32, hs32 = O.block_forward(p["x_v"], p["x_e"], p["edge_index"], p["rev_index"], p["weights"], p["biases"])
    csr = ops.build_graph_csr(p["edge_index"].cuda(), p["rev_index"].cuda(), p["V"])
    h = ops.edge_init(p["x_v"].cuda(), p["x_e"].cuda(), csr)
    for l in range(3):
        h = ops.layer(h, p["weights"][l].cuda(), p["biases"][l].cuda(), csr)
        ref = hs64[l + 1]
        flips = int(((h.cpu() > 0) != (ref > 0)).sum())
        flips32 = int(((hs32[l + 1] > 0) != (ref > 0)).sum())
        print(f"{mode} layer {l}: rel err vs f64 {rel_err(h, ref):.2e} (cpu f32 oracle: {rel_err(hs32[l+1], ref):.2e}); relu sign flips vs f64: ours {flips}, cpu f32 {flips32}, of {h.numel()}")
