"""rel-to-max error of one message-passing depth (forward, dgrad, wgrad) vs an fp64 evaluation, per hidden size. usage: err_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from notorch_b200 import ops
from oracle import dmpnn_oracle as O

def rel(x, ref):
    return float((x.double().cpu() - ref).abs().max() / ref.abs().max())

for d, E, V in ((300, 4096, 1800), (1024, 2048, 900), (2048, 1024, 500)):
    gen = torch.Generator().manual_seed(d)
    src, dst, rev = torch.randint(0, V, (E,), generator=gen), torch.randint(0, V, (E,), generator=gen), torch.randint(0, E, (E,), generator=gen)
    h, W, b = torch.randn(E, d, generator=gen), (torch.rand(d, d, generator=gen) * 2 - 1) / d ** 0.5, torch.randn(d, generator=gen) / 10
    g = torch.randn(E, d, generator=gen)
    hd, Wd, bd = h.double().requires_grad_(True), W.double().requires_grad_(True), b.double().requires_grad_(True)
    ref, _ = O.layer_forward(hd, V, src, dst, rev, Wd, bd)
    ref.backward(g.double())
    csr = ops.build_graph_csr(torch.stack([src, dst]).cuda(), rev.cuda(), V)
    hc, Wc, bc = h.cuda().requires_grad_(True), W.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    out = ops.layer(hc, Wc, bc, csr)
    out.backward(g.cuda())
    print(f"d={d:5d}  fwd {rel(out.detach(), ref.detach()):.2e}  g_h {rel(hc.grad, hd.grad):.2e}  g_W {rel(Wc.grad, Wd.grad):.2e}  g_b {rel(bc.grad, bd.grad):.2e}")
