"""Round-2b probes on one B200: K4b with alternating accumulator windows (error vs fp64 + time at several shapes, positive operands so
that the truncation error of long chains is coherent) and the K6 variants (NOTORCH_B200_K6_VARIANT 0..3: bit 0 = hoisted index loads,
bit 1 = reversed block order; equality + time with the producer's output left in L2 and with L2 flushed).
usage: python scripts/probe_r02b.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from notorch_b200 import ops, _lib, BatchedGraph
from notorch_b200.synth import make_molecules

L = _lib.lib()
p = lambda t: None if t is None else t.data_ptr()
st = lambda: torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(f, n=10, do_flush=True):
    ts = []
    for _ in range(n):
        if do_flush:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts = sorted(ts[2:])
    return ts[len(ts) // 2] * 1e3


def wgrad(E, d, positive):
    gen = torch.Generator(device="cuda").manual_seed(E + d)
    m = torch.randn(E, d, device="cuda", generator=gen)
    g = torch.randn(E, d, device="cuda", generator=gen)
    if positive:  # every product has the same sign: the truncation residue of a long accumulation chain adds up coherently
        m, g = m.abs() + 0.5, g.abs() + 0.5
    gW, gb = torch.empty(d, d, device="cuda"), torch.empty(d, device="cuda")
    ws = torch.empty(max(L.nt_layer_backward_wgrad_workspace_bytes(E, d), 256), dtype=torch.uint8, device="cuda")
    def run():
        _lib.check(L.nt_layer_backward_wgrad(p(g), p(m), None, None, None, None, E, 1, d, 1, 0.0, 0.0, 0, 0, p(gW), p(gb), p(ws), ws.numel(),
                                             _lib.NT_F32, _lib.GEMM_TF32X3, st()), "wgrad")
    run(); torch.cuda.synchronize()
    ref = g.double().t() @ m.double()
    refb = g.double().sum(0)
    eW = float((gW.double() - ref).abs().max() / ref.abs().max())
    eb = float((gb.double() - refb).abs().max() / refb.abs().max())
    first = gW.clone(); run(); torch.cuda.synchronize()
    t = timed(run)
    print(f"K4b E={E:7d} d={d:5d} {'pos' if positive else 'rnd'}: gW {eW:.2e}  gb {eb:.2e}  deterministic {torch.equal(first, gW)}  {t:8.1f} us"
          f"  ({2.0 * E * d * (d + 1) / t / 1e6:6.1f} alg TFLOP/s)", flush=True)


for E, d in ((100, 300), (600, 300), (1100, 300), (33000, 300), (205166, 300), (300000, 300), (37, 64), (5000, 256), (70000, 256),
             (40000, 1024), (819000 // 4, 1024), (20000, 2048), (9000, 332), (9000, 576)):
    for positive in (False, True):
        wgrad(E, d, positive)

# ---------------------------------------------------------------- K6
d = 300
mols = make_molecules(4096, 2)
V, E = mols.total_atoms, mols.total_edges
G = BatchedGraph.from_packed(mols, torch.randn(V, d, device="cuda"), torch.randn(E, d, device="cuda"), device="cuda")
csr = ops.graph_csr(G)
g, h, g_m = (torch.randn(E, d, device="cuda") for _ in range(3))
se = ops._ell_of(csr.by_src)
W = torch.randn(d, d, device="cuda") / d ** 0.5
outs = {}


def k6(o, mean=0):
    _lib.check(L.nt_layer_backward_epilogue_fused(p(g), p(h), p(g_m), p(csr.dst), p(csr.by_src.rowptr), p(csr.by_src.perm), p(se), p(csr.by_rev.rowptr),
                                                  p(csr.by_rev.perm), p(csr.by_dst.rowptr), E, d, 1, 0.0, 1, mean, p(o), _lib.NT_F32, st()), "k6f")


img_t = ops._weight_image(W, True)


def k4a_then_k6(o):  # the real sequence: K4a writes g_m front to back, K6 follows
    _lib.check(L.nt_layer_backward_dgrad(p(g), p(W), p(img_t), E, d, 0.0, 0, 0, p(g_m), _lib.NT_F32, _lib.GEMM_TF32X3, st()), "k4a")
    k6(o)


alg = (V + 4 * E) * d * 4 + 12 * E
for mean in (0, 1):
    for v in (0, 1, 2, 3):
        os.environ["NOTORCH_B200_K6_VARIANT"] = str(v)
        outs[v] = torch.full_like(h, float("nan"))
        k6(outs[v], mean)
    torch.cuda.synchronize()
    print(f"K6 mean={mean}: variants equal {all(torch.equal(outs[0], outs[v]) for v in (1, 2, 3))}")
o = torch.empty_like(h)
t4a = timed(lambda: _lib.check(L.nt_layer_backward_dgrad(p(g), p(W), p(img_t), E, d, 0.0, 0, 0, p(g_m), _lib.NT_F32, _lib.GEMM_TF32X3, st()), "k4a"))
print(f"K4a alone {t4a:7.1f} us")
for v in (0, 1, 2, 3):
    os.environ["NOTORCH_B200_K6_VARIANT"] = str(v)
    t = timed(lambda: k6(o))
    t2 = timed(lambda: k4a_then_k6(o))
    print(f"K6 variant {v}: {t:7.1f} us L2 flushed ({alg / t / 1e3:6.0f} GB/s algorithmic);  after K4a: {t2 - t4a:7.1f} us", flush=True)
os.environ.pop("NOTORCH_B200_K6_VARIANT")
