"""Times the two forms of the backward epilogue (K5 + K6) at BASELINE configs[1] shapes and checks they agree bit for bit.
usage: python scripts/bench_k6.py [d]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from notorch_b200 import ops, _lib, BatchedGraph
from notorch_b200.synth import make_molecules

d = int(sys.argv[1]) if len(sys.argv) > 1 else 300
mols = make_molecules(4096, 2)
V, E = mols.total_atoms, mols.total_edges
G = BatchedGraph.from_packed(mols, torch.randn(V, d, device="cuda"), torch.randn(E, d, device="cuda"), device="cuda")
csr = ops.graph_csr(G)
g, h, g_m = (torch.randn(E, d, device="cuda") for _ in range(3))
L = _lib.lib(); p = lambda t: None if t is None else t.data_ptr(); st = torch.cuda.current_stream().cuda_stream
se, de = ops._ell_of(csr.by_src), ops._ell_of(csr.by_dst)
outs = [torch.empty_like(h) for _ in range(2)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def unfused(o, mean):
    g_n = ops._seg_reduce_raw(g_m, csr.by_src, tag="K5")
    _lib.check(L.nt_layer_backward_epilogue(p(g), p(h), p(g_n), p(g_m), p(csr.dst), p(csr.by_rev.rowptr), p(csr.by_rev.perm), p(csr.by_dst.rowptr),
                                            E, d, 1, 0.0, 1, mean, p(o), _lib.NT_F32, st), "k6")
def edges(o, mean):
    _lib.check(L.nt_layer_backward_epilogue_fused(p(g), p(h), p(g_m), p(csr.dst), p(csr.by_src.rowptr), p(csr.by_src.perm), p(se), p(csr.by_rev.rowptr),
                                                  p(csr.by_rev.perm), p(csr.by_dst.rowptr), E, d, 1, 0.0, 1, mean, p(o), _lib.NT_F32, st), "k6f")
for mean in (0, 1):
    for f, o in zip((unfused, edges), outs):
        o.fill_(float("nan")); f(o, mean)
    torch.cuda.synchronize()
    print(f"mean={mean}: fused == unfused {torch.equal(outs[0], outs[1])}")
alg = (V + 4 * E) * d * 4 + 12 * E
for name, f in (("unfused K5+K6", unfused), ("fused (edge-major)", edges)):
    ts = []
    for _ in range(12):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(outs[0], 0); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = sorted(ts[2:])[len(ts[2:]) // 2]
    print(f"{name:18s} d={d} E={E}: {t * 1e3:7.1f} us  ({alg / t / 1e6:6.0f} GB/s algorithmic, L2 flushed between runs)")
