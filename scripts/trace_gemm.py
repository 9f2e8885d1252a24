"""Role timeline of CTA 0 of the fused tcgen05 kernel (debug trace buffer)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from notorch_b200 import ops, _lib, BatchedGraph
from notorch_b200.synth import make_molecules

which = sys.argv[1] if len(sys.argv) > 1 else "dgrad"
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
buf = torch.zeros(65001, dtype=torch.int64, device="cuda")
L.nt_debug_set_trace_buffer(buf.data_ptr()); f(); torch.cuda.synchronize(); L.nt_debug_set_trace_buffer(None)
rec = buf[1:].cpu().numpy().astype("uint64"); rec = rec[rec != 0]; n_rec = len(rec)
ev = (rec >> 56) & 0xFF; tile = (rec >> 40) & 0xFFFF; aux = (rec >> 32) & 0xFF; clk = (rec & 0xFFFFFFFF).astype("int64")
t0 = clk.min(); names = {7: "epi chunk stored", 8: "epi refill issued", 5: "epi chunk loaded", 6: "epi chunk staged", 1: "epi wait", 2: "epi ready", 3: "epi handback", 4: "epi done", 10: "mma wait tmem", 11: "mma tmem ok", 12: "mma A ready", 13: "mma W ready", 14: "mma issued", 20: "prod loads issued", 21: "prod stage free", 22: "prod stage written"}
print(f"{which}: {n_rec} records; tiles seen {sorted(set(tile.tolist()))[:4]}...")
tiles = sorted(set(tile.tolist()))
order = sorted(range(len(rec)), key=lambda i: clk[i])
sel = tiles[2] if len(tiles) > 3 else tiles[0]
print(f"--- timeline around tile {sel} (cycles since kernel start; only events of tiles {sel} and next)")
nxt = tiles[tiles.index(sel) + 1] if tiles.index(sel) + 1 < len(tiles) else sel
for i in order:
    if tile[i] == sel and ev[i] < 10 and aux[i] <= 7:
        print(f"{clk[i]-t0:9d}  tile {tile[i]:5d} kb {aux[i]:2d}  {names.get(int(ev[i]), ev[i])}")
# per-tile summaries
def first(evid, t, a=None):
    c = [clk[i] for i in range(len(rec)) if ev[i] == evid and tile[i] == t and (a is None or aux[i] == a)]
    return min(c) if c else None
print("--- per tile: mma tmem wait, mainloop (tmem ok -> last issue), epilogue (ready -> done), tile period")
prev = None
for t in tiles[1:8]:
    a, b_, c = first(10, t), first(11, t), first(14, t, 9)
    e2, e4 = first(2, t), first(4, t)
    if None in (a, b_, c, e2, e4): continue
    print(f"tile {t}: tmem wait {b_-a:7d}  mainloop {c-b_:7d}  epilogue {e4-e2:7d}  period {'' if prev is None else b_-prev}")
    prev = b_
