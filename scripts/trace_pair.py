"""Role timeline of CTA 0 of the CTA-pair tcgen05 kernel (gemm_pair.cu debug trace buffer). usage: trace_pair.py fwd|dgrad"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from notorch_b200 import ops, _lib, BatchedGraph
from notorch_b200.synth import make_molecules

which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
mols = make_molecules(4096, 2)
V, E, d = mols.total_atoms, mols.total_edges, 300
xv, xe = torch.randn(V, d, device="cuda"), torch.randn(E, d, device="cuda")
G = BatchedGraph.from_packed(mols, xv, xe, device="cuda"); csr = ops.graph_csr(G)
W = torch.randn(d, d, device="cuda") / 17; b = torch.zeros(d, device="cuda")
h = torch.randn(E, d, device="cuda"); g = torch.randn(E, d, device="cuda")
L = _lib.lib(); p = lambda t: None if t is None else t.data_ptr(); st = torch.cuda.current_stream().cuda_stream
n = ops._seg_reduce_raw(h, csr.by_dst, 1, 0.0, False)
img = ops._weight_image(W, False); imgt = ops._weight_image(W, True)
out = torch.empty_like(h); m = torch.empty_like(h)
def fwd(): _lib.check(L.nt_layer_forward(p(h), p(n), p(csr.src), p(csr.rev), p(W), p(img), p(b), E, V, d, 1, 0.0, 1, 0.0, 0, 0, p(out), p(m), 0, 0, st), "fwd")
def dgrad(): _lib.check(L.nt_layer_backward_dgrad(p(g), p(W), p(imgt), E, d, 0.0, 0, 0, p(out), 0, 0, st), "dgrad")
f = fwd if which == "fwd" else dgrad
for _ in range(3): f()
torch.cuda.synchronize()
R = 16000
buf = torch.zeros(1 + 4 * R, dtype=torch.int64, device="cuda")
L.nt_debug_set_trace_buffer(buf.data_ptr()); f(); torch.cuda.synchronize(); L.nt_debug_set_trace_buffer(None)
raw = buf[1:].cpu().numpy().astype("uint64")
recs = []
for region in range(4):
    r = raw[region * R:(region + 1) * R]; r = r[r != 0]
    for x in r:
        recs.append((region, int((x >> 56) & 0xFF), int((x >> 40) & 0xFFFF), int((x >> 32) & 0xFF), int(x & 0xFFFFFFFF)))
t0 = min(c for *_, c in recs)
tiles = sorted({t for _, _, t, _, _ in recs})
print(f"{which}: {len(recs)} records, tiles {tiles[:5]} ... ({len(tiles)} tiles on CTA 0)")
def ev(e, t, a=None):
    c = [c for _, e2, t2, a2, c in recs if e2 == e and t2 == t and (a is None or a2 == a)]
    return min(c) - t0 if c else None
sel = tiles[3] if len(tiles) > 4 else tiles[0]
kbs = (d + 31) // 32
print(f"--- tile {sel}: per K-block  [transform: wait raw -> raw landed -> arrived]   [MMA: ready seen -> issued]")
for kb in range(kbs):
    print(f"kb {kb:2d}  TMA issue {ev(30, sel, kb)}  X wait {ev(20, sel, kb)}  raw {ev(21, sel, kb)}  loop {ev(23, sel, kb)}  fence {ev(24, sel, kb)}  W+arrive {ev(22, sel, kb)}   MMA ready {ev(12, sel, kb)}  issued {ev(14, sel, kb)}")
print("--- per tile: MMA wait for TMEM, mainloop (TMEM granted -> last K-block issued), epilogue (accumulator ready -> handback -> done), tile period")
prev = None
for t in tiles[1:9]:
    a, b_, c = ev(10, t), ev(11, t), ev(14, t, kbs - 1)
    e1, e2, e3, e4 = ev(1, t), ev(2, t), ev(3, t), ev(4, t)
    if None in (a, b_, c, e2, e4): continue
    print(f"tile {t}: tmem wait {b_ - a:6d}  mainloop {c - b_:6d}  epi wait {e2 - e1:6d}  epi ready->handback {(e3 - e2) if e3 else -1:6d}  epi ready->done {e4 - e2:6d}  period {'' if prev is None else b_ - prev}")
    prev = b_
