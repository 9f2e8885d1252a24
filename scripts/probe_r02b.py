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


WG_SHAPES = ((100, 300), (600, 300), (1100, 300), (33000, 300), (205166, 300), (300000, 300), (37, 64), (5000, 256), (70000, 256),
             (40000, 1024), (819000 // 4, 1024), (20000, 2048), (9000, 332), (9000, 576))
if "wgrad1" in sys.argv:  # the headline shape only (e.g. under NOTORCH_B200_WGRAD_HALF_COST / _SEG sweeps: those are read once per process)
    wgrad(205166, 300, False)
if "wgrad" in sys.argv or len(sys.argv) == 1:
    for E, d in WG_SHAPES:
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
RUN_K6 = "k6" in sys.argv or len(sys.argv) == 1
for mean in (0, 1) if RUN_K6 else ():
    for v in (0, 1, 2, 3):
        os.environ["NOTORCH_B200_K6_VARIANT"] = str(v)
        outs[v] = torch.full_like(h, float("nan"))
        k6(outs[v], mean)
    torch.cuda.synchronize()
    print(f"K6 mean={mean}: variants equal {all(torch.equal(outs[0], outs[v]) for v in (1, 2, 3))}")
o = torch.empty_like(h)
if not RUN_K6:
    t4a = 0.0
else:
  t4a = timed(lambda: _lib.check(L.nt_layer_backward_dgrad(p(g), p(W), p(img_t), E, d, 0.0, 0, 0, p(g_m), _lib.NT_F32, _lib.GEMM_TF32X3, st()), "k4a"))
if RUN_K6:
    print(f"K4a alone {t4a:7.1f} us")
for v in (0, 3) if RUN_K6 else ():
    os.environ["NOTORCH_B200_K6_VARIANT"] = str(v)
    t = timed(lambda: k6(o))
    t2 = timed(lambda: k4a_then_k6(o))
    print(f"K6 variant {v}: {t:7.1f} us L2 flushed ({alg / t / 1e3:6.0f} GB/s algorithmic);  after K4a: {t2 - t4a:7.1f} us", flush=True)
os.environ.pop("NOTORCH_B200_K6_VARIANT", None)

# ---------------------------------------------------------------- embedding-table gradient: mma.sync kernel vs the tcgen05 pair kernel
if "embbwd" in sys.argv or len(sys.argv) == 1:
    gen = torch.Generator(device="cuda").manual_seed(5)
    for dd, Tv, Te in ((300, 45, 13), (64, 45, 13), (1024, 100, 20), (300, 10, 3)):
        nt_ = torch.randint(0, Tv, (V, 7), device="cuda", generator=gen)
        et_ = torch.randint(0, Te, (E, 2), device="cuda", generator=gen)
        gg = torch.randn(E, dd, device="cuda", generator=gen)
        src64 = csr.src.long()
        cnt_v = torch.zeros(E, Tv, dtype=torch.float64, device="cuda").scatter_add_(1, nt_[src64], torch.ones(E, 7, dtype=torch.float64, device="cuda"))
        cnt_e = torch.zeros(E, Te, dtype=torch.float64, device="cuda").scatter_add_(1, et_, torch.ones(E, 2, dtype=torch.float64, device="cuda"))
        ref_v, ref_e = cnt_v.t() @ gg.double(), cnt_e.t() @ gg.double()
        for kind in ("mma", "tc"):
            if kind == "mma" and Tv + Te > 64:
                continue
            os.environ["NOTORCH_B200_EMBBWD"] = kind
            f = lambda: ops._embed_edge_init_backward_raw(gg, nt_, et_, csr.src, V, Tv, Te)
            gv, ge = f(); torch.cuda.synchronize()
            gv2, ge2 = f(); torch.cuda.synchronize()
            ev = float((gv.double() - ref_v).abs().max() / ref_v.abs().max()); ee = float((ge.double() - ref_e).abs().max() / ref_e.abs().max())
            t = timed(f)
            print(f"embbwd {kind:3s} d={dd:5d} T={Tv}+{Te}: gTv {ev:.2e} gTe {ee:.2e} deterministic {torch.equal(gv, gv2) and torch.equal(ge, ge2)} {t:7.1f} us", flush=True)
    os.environ.pop("NOTORCH_B200_EMBBWD")

# ---------------------------------------------------------------- L2 look-ahead of the pair weight-gradient kernel (NOTORCH_B200_WGRAD_PF)
if "pf" in sys.argv:
    for pf in (0, 2, 4, 6, 10):
        os.environ["NOTORCH_B200_WGRAD_PF"] = str(pf)
        print(f"--- look-ahead {pf} K-blocks")
        for E_, d_ in ((205166, 300), (204750, 1024), (20000, 2048)):
            wgrad(E_, d_, False)
        gen = torch.Generator(device="cuda").manual_seed(5)
        Tv, Te, dd = 45, 13, 300
        nt_ = torch.randint(0, Tv, (V, 7), device="cuda", generator=gen); et_ = torch.randint(0, Te, (E, 2), device="cuda", generator=gen)
        gg = torch.randn(E, dd, device="cuda", generator=gen)
        ws = torch.empty(L.nt_embed_edge_init_backward_workspace_bytes(E, Tv, Te, dd), dtype=torch.uint8, device="cuda")
        gv, ge = torch.empty(Tv, dd, device="cuda"), torch.empty(Te, dd, device="cuda")
        f = lambda: _lib.check(L.nt_embed_edge_init_backward(p(gg), p(nt_), 7, p(et_), 2, p(csr.src), E, V, Tv, Te, dd, p(gv), p(ge), p(ws), ws.numel(),
                                                             _lib.NT_F32, st()), "embbwd")
        print(f"embbwd tc d=300: {timed(f):7.1f} us", flush=True)
    os.environ.pop("NOTORCH_B200_WGRAD_PF")

# ---------------------------------------------------------------- pooled epilogue variants (NOTORCH_B200_K6P_ITEMS: 0 = runs, 1 | 2 | 4 items)
if "k6p" in sys.argv:
    Bm = len(G)
    pool = ops.mol_edge_csr(G)
    GH, GHW = torch.randn(Bm, d, device="cuda"), torch.randn(Bm, d, device="cuda")
    ws6 = torch.empty(L.nt_layer_backward_epilogue_pooled_workspace_bytes(E), dtype=torch.uint8, device="cuda")
    outs6 = {}
    def k6p(o, mean=0, act=1):
        _lib.check(L.nt_layer_backward_epilogue_pooled(p(GH), p(GHW), p(h), p(pool.keys32), p(csr.dst), p(csr.by_src.rowptr), p(csr.by_rev.rowptr),
                                                       p(csr.by_rev.perm), p(csr.by_dst.rowptr), E, Bm, d, act, 0.0, 1, mean, p(o), p(ws6), ws6.numel(),
                                                       _lib.NT_F32, st()), "k6p")
    for v in (1, 0, 2, 4):
        os.environ["NOTORCH_B200_K6P_ITEMS"] = str(v)
        outs6[v] = torch.full_like(h, float("nan")); k6p(outs6[v], 1, 3)
    torch.cuda.synchronize()
    print("K6p variants equal:", all(torch.equal(outs6[1], outs6[v]) for v in (0, 2, 4)))
    o6 = torch.empty_like(h)
    for v in (1, 0, 2, 4):
        os.environ["NOTORCH_B200_K6P_ITEMS"] = str(v)
        print(f"K6p variant {v}: {timed(lambda: k6p(o6)):7.1f} us", flush=True)
    os.environ.pop("NOTORCH_B200_K6P_ITEMS")
