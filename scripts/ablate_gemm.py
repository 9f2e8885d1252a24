"""Timing experiment: which role bounds the fused tcgen05 kernel? Run with NOTORCH_B200_ABLATE=<bits>."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from notorch_b200 import ops, _lib, BatchedGraph
from notorch_b200.synth import make_molecules

mols = make_molecules(4096, 2)
V, E, d = mols.total_atoms, mols.total_edges, 300
xv, xe = torch.randn(V, d, device="cuda"), torch.randn(E, d, device="cuda")
G = BatchedGraph.from_packed(mols, xv, xe, device="cuda")
csr = ops.graph_csr(G)
W = torch.randn(d, d, device="cuda") / 17; b = torch.zeros(d, device="cuda")
h = torch.randn(E, d, device="cuda"); g = torch.randn(E, d, device="cuda")
L = _lib.lib(); p = lambda t: None if t is None else t.data_ptr()
st = torch.cuda.current_stream().cuda_stream
mode = {"tf32x3": 0, "tf32": 2}[os.environ.get("GEMM", "tf32x3")]
n = ops._seg_reduce_raw(h, csr.by_dst, 1, 0.0, False)
img = ops._weight_image(W, False); imgt = ops._weight_image(W, True)
out = torch.empty_like(h); m = torch.empty_like(h)
def fwd(): _lib.check(L.nt_layer_forward(p(h), p(n), p(csr.src), p(csr.rev), p(W), p(img), p(b), E, V, d, 1, 0.0, 1, 0.0, 0, 0, p(out), p(m), 0, mode, st), "fwd")
def fwd_nom(): _lib.check(L.nt_layer_forward(p(h), p(n), p(csr.src), p(csr.rev), p(W), p(img), p(b), E, V, d, 1, 0.0, 1, 0.0, 0, 0, p(out), None, 0, mode, st), "fwd")
def fwd_nores(): _lib.check(L.nt_layer_forward(p(h), p(n), p(csr.src), p(csr.rev), p(W), p(img), p(b), E, V, d, 1, 0.0, 0, 0.0, 0, 0, p(out), None, 0, mode, st), "fwd")
def fwd_nores_nobias(): _lib.check(L.nt_layer_forward(p(h), p(n), p(csr.src), p(csr.rev), p(W), p(img), None, E, V, d, 1, 0.0, 0, 0.0, 0, 0, p(out), None, 0, mode, st), "fwd")
def dgrad(): _lib.check(L.nt_layer_backward_dgrad(p(g), p(W), p(imgt), E, d, 0.0, 0, 0, p(out), 0, mode, st), "dgrad")
def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize(); s, e = torch.cuda.Event(True), torch.cuda.Event(True)
    s.record()
    for _ in range(n): f()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n * 1e3
print(f"ABLATE={os.environ.get('NOTORCH_B200_ABLATE','0'):>2s} GEMM={os.environ.get('GEMM','tf32x3'):7s} K2 {timeit(fwd):7.1f} us   K2(no m_out) {timeit(fwd_nom):7.1f} us   K4a {timeit(dgrad):7.1f} us   K2(no m_out, no resid) {timeit(fwd_nores):7.1f}   K2(no m_out/resid/bias) {timeit(fwd_nores_nobias):7.1f}")
