"""Diagnostic: tensor-core wgrad vs fp64, with a per-block error map when it is off."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from notorch_b200 import ops, _lib

def run(E, d, mode):
    ops.set_gemm_mode(mode)
    gen = torch.Generator().manual_seed(E * 7 + d)
    V = max(1, E // 2)
    src = torch.randint(0, V, (E,), generator=gen); dst = torch.randint(0, V, (E,), generator=gen); rev = torch.randint(0, E, (E,), generator=gen)
    h = torch.randn(E, d, generator=gen); g = torch.randn(E, d, generator=gen)
    W = (torch.rand(d, d, generator=gen) * 2 - 1) / d ** 0.5; b = torch.randn(d, generator=gen)
    csr = ops.build_graph_csr(torch.stack([src, dst]).cuda(), rev.cuda(), V)
    hc, Wc, bc = h.cuda().requires_grad_(True), W.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    out = ops.layer(hc, Wc, bc, csr, residual=True)
    (out * g.cuda()).sum().backward()
    a = torch.relu(h.double()); n = torch.zeros(V, d, dtype=torch.float64).index_add_(0, dst, a)
    m = n[src] - a[rev]
    gW = g.double().t() @ m; gb = g.double().sum(0)
    eW = (Wc.grad.cpu().double() - gW).abs() / gW.abs().max(); eb = (bc.grad.cpu().double() - gb).abs() / gb.abs().max()
    print(f"E={E} d={d} {mode}: gW err {eW.max():.2e} gb err {eb.max():.2e}")
    if eW.max() > 1e-5:
        nb = (d + 31) // 32
        for bi in range(nb):
            print("  o-block", bi, " ".join(f"{eW[bi*32:(bi+1)*32, bj*32:(bj+1)*32].max():.0e}" for bj in range(nb)))

for E, d in [(1, 16), (8, 32), (32, 32), (40, 64), (127, 64), (128, 300), (1000, 256), (5000, 300), (300, 1024)]:
    run(E, d, "tf32x3")
