"""Do K4b (weight gradient, persistent tcgen05 kernel) and K6 (memory-bound backward epilogue) overlap when launched on two streams?
Times each alone and both together at BASELINE configs[1] shapes. usage: python scripts/probe_overlap.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from notorch_b200 import ops, _lib, BatchedGraph
from notorch_b200.synth import make_molecules

d = 300
mols = make_molecules(4096, 2)
V, E = mols.total_atoms, mols.total_edges
G = BatchedGraph.from_packed(mols, torch.randn(V, d, device="cuda"), torch.randn(E, d, device="cuda"), device="cuda")
csr = ops.graph_csr(G)
g, h, g_m, m = (torch.randn(E, d, device="cuda") for _ in range(4))
W = torch.randn(d, d, device="cuda") / 17
L = _lib.lib(); p = lambda t: None if t is None else t.data_ptr()
se = ops._ell_of(csr.by_src)
out = torch.empty_like(h); gW = torch.empty_like(W); gb = torch.empty(d, device="cuda")
ws = torch.empty(L.nt_layer_backward_wgrad_workspace_bytes(E, d), dtype=torch.uint8, device="cuda")
imgt = ops._weight_image(W, True); g_m2 = torch.empty_like(h)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def k6(st):
    _lib.check(L.nt_layer_backward_epilogue_fused(p(g), p(h), p(g_m), p(csr.dst), p(csr.by_src.rowptr), p(csr.by_src.perm), p(se), p(csr.by_rev.rowptr),
                                                  p(csr.by_rev.perm), p(csr.by_dst.rowptr), E, d, 1, 0.0, 1, 0, p(out), _lib.NT_F32, st.cuda_stream), "k6")
def k4b(st):
    _lib.check(L.nt_layer_backward_wgrad(p(g), p(m), None, None, None, None, E, V, d, 1, 0.0, 0.0, 0, 0, p(gW), p(gb), p(ws), ws.numel(), _lib.NT_F32,
                                         _lib.GEMM_TF32X3, st.cuda_stream), "k4b")
def k4a(st):
    _lib.check(L.nt_layer_backward_dgrad(p(g), p(W), p(imgt), E, d, 0.0, 0, 0, p(g_m2), _lib.NT_F32, _lib.GEMM_TF32X3, st.cuda_stream), "k4a")

def timed(fn, reps=8):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s1); s2.wait_event(a)
        fn()
        j = torch.cuda.Event(); j.record(s2); s1.wait_event(j); b.record(s1)
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts[2:])[len(ts[2:]) // 2]

print(f"K6 alone            {timed(lambda: k6(s1)):7.1f} us")
print(f"K4b alone           {timed(lambda: k4b(s1)):7.1f} us")
print(f"K4a alone           {timed(lambda: k4a(s1)):7.1f} us")
print(f"K4b then K6, 1 strm {timed(lambda: (k4b(s1), k6(s1))):7.1f} us")
print(f"K4b | K6  (K4b 1st) {timed(lambda: (k4b(s2), k6(s1))):7.1f} us")
print(f"K6 | K4b  (K6 1st)  {timed(lambda: (k6(s1), k4b(s2))):7.1f} us")
print(f"K4a then K4b | K6   {timed(lambda: (k4a(s1), k4b(s1), k6(s1))):7.1f} us (serial)")
print(f"K4a | K4b           {timed(lambda: (k4a(s1), k4b(s2))):7.1f} us")
