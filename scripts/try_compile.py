"""Probe: does torch.compile trace the functional layer API / the modules through the dispatcher ops? (run on a GPU box)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import oracle_inputs
from notorch_b200 import ops, BatchedGraph
from notorch_b200.nn import ChempropBlock, Sum

ops.set_index_validation("off")
p = oracle_inputs(16, 64, 2, seed=3)
csr = ops.build_graph_csr(p["edge_index"].cuda(), p["rev_index"].cuda(), p["V"])
h = torch.randn(p["E"], 64, device="cuda", requires_grad=True)
W = (torch.randn(64, 64, device="cuda") / 8).requires_grad_(True)
b = torch.zeros(64, device="cuda", requires_grad=True)

def f(h, W, b):
    return ops.layer(ops.layer(h, W, b, csr), W, b, csr).square().mean()

ref = f(h, W, b)
for backend in ("eager", "aot_eager"):
    try:
        torch._dynamo.reset()
        g = torch.compile(f, backend=backend, fullgraph=True)
        out = g(h, W, b)
        out.backward()
        print(backend, "functional fullgraph OK", float(out), float(ref), torch.equal(out, ref))
    except Exception as e:
        print(backend, "functional FAILED:", type(e).__name__, str(e)[:300])

blk = ChempropBlock(hidden_dim=64, depth=2).cuda()
G = BatchedGraph(p["x_v"].cuda(), p["x_e"].cuda(), p["edge_index"].cuda(), p["rev_index"].cuda(), batch_node_index=p["batch_node_index"].cuda(),
                 batch_edge_index=p["batch_edge_index"].cuda(), size=16)
ops.graph_csr(G); ops.segment_csr_for(G, "batch_node_index", 16)
def m(G):
    return Sum()(blk(G))
ref = m(G)
for backend in ("eager", "aot_eager"):
    try:
        torch._dynamo.reset()
        g = torch.compile(m, backend=backend)
        out = g(G)
        print(backend, "module OK", torch.equal(out, ref), "graph breaks:", torch._dynamo.utils.counters.get("graph_break", {}))
    except Exception as e:
        import traceback
        print(backend, "module FAILED:", type(e).__name__, str(e)[:300])
        traceback.print_exc(limit=40)
