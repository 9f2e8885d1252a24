"""CPU-only checks: the C-ABI library loads and exports every symbol ``include/notorch_b200.h``
declares (no compute without a GPU), argument errors come back as status codes, and the host-side
mirror of the reference interface behaves like the reference."""
from __future__ import annotations

import os
import re

import numpy as np
import pytest
import torch

from oracle import dmpnn_oracle as O
from oracle import reference_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols(name: str = "notorch_b200.h") -> set[str]:
    text = open(os.path.join(ROOT, "include", name)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(nt_[a-z0-9_]+)\s*\(", text))


def test_library_exports_every_header_symbol():
    from notorch_b200 import _lib

    lib = _lib.lib()
    declared = _header_symbols()
    assert declared, "no symbols parsed from the header"
    assert not any(n.startswith("nt_debug_") for n in declared), "debug entry points belong in notorch_b200_debug.h"
    debug = _header_symbols("notorch_b200_debug.h")
    assert debug and all(n.startswith("nt_debug_") for n in debug)
    assert declared | debug == set(_lib.SIGNATURES), ((declared | debug) ^ set(_lib.SIGNATURES))
    for name in declared | debug:
        assert hasattr(lib, name), name
    header = open(os.path.join(ROOT, "include", "notorch_b200.h")).read()
    assert int(re.search(r"#define NT_ABI_VERSION (\d+)", header).group(1)) == _lib.ABI_VERSION == lib.nt_version()


def test_argument_errors_are_status_codes_not_crashes():
    from notorch_b200 import _lib

    lib = _lib.lib()
    # bad dtype / sizes are rejected before any CUDA call (safe without a GPU)
    rc = lib.nt_seg_reduce(None, 300, None, None, 10, 1, 0.0, 0, 1.0, None, 7, None)
    assert rc == 1 and b"dtype" in lib.nt_last_error_string()
    rc = lib.nt_seg_reduce(None, 300, None, None, 10, 1, 0.0, 0, 1.0, None, _lib.NT_BF16, None)
    assert rc == 3
    rc = lib.nt_layer_forward(None, None, None, None, None, None, None, 5, 5, 300, 1, 0.0, 1, 1.5, 0, 0, None, None, 0, 0, None)
    assert rc == 1 and b"dropout_p" in lib.nt_last_error_string()
    rc = lib.nt_build_csr(None, -1, 4, None, None, None, None, None, 0, None)
    assert rc == 1
    with pytest.raises(RuntimeError, match="status 1"):
        _lib.check(rc, "nt_build_csr")
    assert lib.nt_weight_image_bytes(300) == 2 * 10 * 304 * 128
    assert lib.nt_layer_backward_wgrad_workspace_bytes(205000, 300) > 0


def test_host_from_graphs_matches_reference_golden(golden):
    from notorch_b200 import BatchedGraph, Graph

    Gs = []
    for n, ei, rev in golden.mols():
        ei_t = torch.from_numpy(ei.astype(np.int64)) if ei.shape[1] else torch.empty(0)
        Gs.append(Graph(torch.zeros(n, 1, dtype=torch.long), torch.zeros(len(rev), 1, dtype=torch.long), ei_t,
                        torch.from_numpy(rev.astype(np.int64))))
    G = BatchedGraph.from_graphs(Gs)
    for k in ("edge_index", "rev_index", "batch_node_index", "batch_edge_index"):
        t = getattr(G, k)
        assert t.dtype == torch.int64 and t.is_contiguous()
        assert np.array_equal(t.numpy(), golden[k]), k
    assert len(G) == len(golden["num_atoms"]) and G.num_nodes == int(golden["num_atoms"].sum())
    fixed = BatchedGraph.from_graphs(Gs, fixed_rev=True)
    assert np.array_equal(fixed.rev_index.numpy(), O.collate_fixed(golden.mols())["rev_index"])


def test_graph_update_is_shallow_copy_and_carries_cache():
    from notorch_b200 import BatchedGraph

    G = BatchedGraph(torch.zeros(3, 2), torch.zeros(4, 2), torch.zeros(2, 4, dtype=torch.long), torch.zeros(4, dtype=torch.long),
                     batch_node_index=torch.zeros(3, dtype=torch.long), batch_edge_index=torch.zeros(4, dtype=torch.long), size=1)
    G._nt_csr = "cache"
    G2 = G.update(node_feats=torch.ones(3, 2))
    assert G2 is not G and G2.edge_index is G.edge_index and G2._nt_csr == "cache"
    assert G.node_feats.sum() == 0 and G2.node_feats.sum() == 6 and len(G2) == 1
    assert G.update(in_place=True, node_feats=G2.node_feats) is G
    G.to("cpu")
    assert G._nt_csr == "cache"  # a move that moves nothing keeps the per-batch preprocessing (N4: Lightning's transfer_batch_to_device)
    G.to(torch.device("meta"))
    assert "_nt_csr" not in G.__dict__  # off the GPU the caches are dropped: they only serve the CUDA kernels
    G3 = BatchedGraph(torch.zeros(3, 2), torch.zeros(0, 2), torch.zeros(2, 0, dtype=torch.long), torch.zeros(0, dtype=torch.long),
                      batch_node_index=torch.tensor([0, 0, 2]), batch_edge_index=torch.zeros(0, dtype=torch.long))
    assert len(G3) == 3  # size inferred like the reference (graph.py:184)


def test_module_surface_matches_reference():
    """Same ctor arguments, attributes, parameter names and shapes as the reference modules."""
    import inspect

    from notorch_b200.nn import ChempropBlock, ChempropLayer, Residual

    blk = ChempropBlock(hidden_dim=12, depth=2)
    assert blk.depth == 2 and blk.hidden_dim == 12 and blk.reduce == "sum"
    assert isinstance(blk.layers[0], Residual) and isinstance(blk.layers[0].module, ChempropLayer)
    assert list(blk.state_dict()) == [f"layers.{i}.module.update.0.{w}" for i in range(2) for w in ("weight", "bias")]
    assert list(ChempropBlock(hidden_dim=8, depth=1, residual=False, bias=False).state_dict()) == ["layers.0.update.0.weight"]
    shared = ChempropBlock(hidden_dim=8, depth=3, shared=True)
    assert shared.layers[0].module is shared.layers[2].module
    assert len(list(shared.parameters())) == 2
    if reference_loader.available():
        ref = reference_loader.load()
        for ours, theirs in ((ChempropBlock, ref.ChempropBlock), (ChempropLayer, ref.ChempropLayer)):
            po, pt = inspect.signature(ours.__init__).parameters, inspect.signature(theirs.__init__).parameters
            assert list(po) == list(pt)
            assert all(po[k].default == pt[k].default for k in po)
        rb = ref.ChempropBlock(hidden_dim=12, depth=2)
        assert {k: tuple(v.shape) for k, v in rb.state_dict().items()} == {k: tuple(v.shape) for k, v in blk.state_dict().items()}
        blk.load_state_dict(rb.state_dict(), strict=True)


def test_act_codes_closed_set():
    from notorch_b200 import _lib, ops

    assert ops.act_code(torch.nn.ReLU()) == (_lib.ACT_RELU, 0.0)
    assert ops.act_code(torch.nn.LeakyReLU(0.2)) == (_lib.ACT_LEAKY_RELU, pytest.approx(0.2))
    assert ops.act_code(torch.nn.ELU(1.5)) == (_lib.ACT_ELU, 1.5)
    with pytest.raises(NotImplementedError):
        ops.act_code(torch.nn.Softplus())
    with pytest.raises(NotImplementedError):
        ops.act_code(torch.nn.GELU(approximate="tanh"))


def test_cpu_tensors_raise_no_fallback():
    from notorch_b200 import BatchedGraph
    from notorch_b200.nn import ChempropBlock, Sum

    G = BatchedGraph(torch.zeros(3, 8), torch.zeros(4, 8), torch.zeros(2, 4, dtype=torch.long), torch.zeros(4, dtype=torch.long),
                     batch_node_index=torch.zeros(3, dtype=torch.long), batch_edge_index=torch.zeros(4, dtype=torch.long), size=1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ChempropBlock(hidden_dim=8, depth=1)(G)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Sum()(G)
    from notorch_b200.nn import AtomMessagePassing, GraphEmbedding, Norm

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        AtomMessagePassing(hidden_dim=8, depth=1)(G)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Norm()(G)


def test_gemm_modes_and_new_entry_points_reject_bad_arguments():
    """Argument validation of the entry points added for the atom variant, the ELL reductions and the bf16 mode (no GPU needed: every
    call is rejected before a launch)."""
    from notorch_b200 import _lib, ops

    lib = _lib.lib()
    assert lib.nt_seg_reduce_ell(None, 300, None, None, None, 10, 1, 0.0, 0, 1.0, None, None, None, _lib.NT_F32, None) != 0  # null ell
    assert lib.nt_csr_to_ell(None, None, 10, None, None) != 0
    assert lib.nt_dense_forward(None, None, None, None, 10, 300, 0.0, 0, 0, None, _lib.NT_F32, 7, None) != 0  # bad gemm_mode
    assert lib.nt_dense_forward(None, None, None, None, 10, 300, 0.0, 0, 0, None, _lib.NT_F32, _lib.GEMM_TF32X3, None) != 0  # null pointers
    assert lib.nt_layer_backward_epilogue_fused(None, None, None, None, None, None, None, None, None, None, 10, 300, 1, 0.0, 1, 0, None,
                                                _lib.NT_F32, None) != 0
    assert lib.nt_weight_prepare(None, 300, 0, None, 5, None) != 0  # bad dtype
    old = ops.get_gemm_mode()
    try:
        for mode in ("tf32x3", "tf32", "bf16", "fp32"):
            ops.set_gemm_mode(mode)
            assert ops.get_gemm_mode() == mode
        with pytest.raises(KeyError):
            ops.set_gemm_mode("fp8")
    finally:
        ops.set_gemm_mode(old)


def test_synth_generator_statistics_and_contract():
    from notorch_b200.synth import make_molecules

    mols = make_molecules(512, 2)
    assert 2.0 < mols.total_edges / mols.total_atoms < 2.3  # E/V ~= 2.18 like tests/data/lipo.csv
    n, ei, rev = mols.molecule(5)
    assert np.array_equal(rev, np.arange(len(rev)).reshape(-1, 2)[:, ::-1].ravel())  # [1,0,3,2,...]
    assert np.array_equal(ei[:, 0::2], ei[::-1, 1::2])  # (u->v, v->u) per bond
    assert ei.max() < n
    again = make_molecules(512, 2)
    assert np.array_equal(again.edge_index, mols.edge_index)
    sh = [mols.shard(r, 4) for r in range(4)]
    assert sum(len(s) for s in sh) == 512 and sum(s.total_edges for s in sh) == mols.total_edges


def test_wgrad_pair_geometry_invariants():
    """Host-side work decomposition of the CTA-pair weight-gradient kernel (wgrad_pair.cu: make_geometry), no GPU needed: every unit
    covers all K-blocks, the grid never exceeds the SM count unless there are more units than CTA pairs, the half-height unit is
    used exactly when at most 128 feature rows (incl. the all-ones bias row) remain, and the two MMAs tile the N tile."""
    import ctypes

    from notorch_b200 import _lib

    L = _lib.lib()
    out = (ctypes.c_int64 * 12)()
    for sms in (148, 132, 2):
        for d in (4, 8, 64, 100, 128, 252, 256, 260, 300, 304, 320, 332, 384, 512, 576, 1024, 2048, 2348, 4096):
            for E in (0, 1, 31, 32, 33, 1000, 205166, 819000):
                assert L.nt_debug_wgrad_geometry(E, d, sms, out) == 0
                m_units, n_tiles, n_tile, n_a, n_b, half, full_u, half_u, sf, sl, per, per_l = list(out)
                kb_total = (E + 31) // 32
                ones = 1 if d % 256 else 0
                assert m_units == (d + 255) // 256 and n_tiles * n_tile >= d and n_tile % 64 == 0 and n_tile <= 320
                assert n_a + n_b == n_tile and n_a % 64 == 0 and n_b % 64 == 0 and 0 < n_a <= 256 and 0 <= n_b <= 256
                last_rows = d + ones - (m_units - 1) * 256
                assert half == (1 if last_rows <= 128 else 0)
                assert full_u == (m_units - half) * n_tiles and half_u == half * n_tiles
                clusters = max(sms // 2, 1)
                if full_u:
                    assert sf >= 1 and per * sf >= kb_total and (kb_total == 0 or sf <= kb_total)
                if half_u:
                    assert sl >= 1 and per_l * sl >= kb_total and (kb_total == 0 or sl <= kb_total)
                pairs = full_u * sf + half_u * sl
                assert pairs >= full_u + half_u
                if full_u + half_u <= clusters and kb_total >= clusters:
                    assert pairs <= clusters, (sms, d, E, list(out))
    assert L.nt_debug_wgrad_geometry(10, 6, 148, out) != 0  # d % 4 != 0 is not the tensor-core path


def test_chunk_split_by_multiply_high_is_exact():
    """Every HBM-bound row kernel turns its flat work-item index t into (row, 16-byte chunk) with a multiply-high by
    ceil(2^64 / chunks) instead of an emulated 64-bit division (common.cuh: split_item). Host restatement of that arithmetic
    (nt_debug_split_item) against Python's divmod: item counts up to 2^32 - 1, chunk counts incl. powers of two, 75 (d = 300),
    primes and the largest supported, t at every boundary; above 2^32 items (and for chunks == 1) the general division is used."""
    import ctypes
    import random

    from notorch_b200 import _lib

    L = _lib.lib()
    out = (ctypes.c_int64 * 3)()
    rng = random.Random(7)
    for chunks in (1, 2, 3, 4, 7, 16, 64, 75, 83, 256, 512, 1000, 4099, 65536, (1 << 20) - 1):
        for total in (chunks, 12345 * chunks, (1 << 31) - 1, (1 << 32) - 1, (1 << 32), (1 << 33) + 5):
            if total < chunks:
                continue
            ts = {0, chunks - 1, chunks, total - 1, total // 2, (total // chunks) * chunks - 1, max((total // chunks) * chunks - chunks, 0)}
            ts |= {rng.randrange(total) for _ in range(40)}
            for t in ts:
                if not 0 <= t < total:
                    continue
                assert L.nt_debug_split_item(total, chunks, t, out) == 0
                fast, row, chunk = list(out)
                assert fast == (1 if (chunks > 1 and total < (1 << 32)) else 0)
                assert (row, chunk) == divmod(t, chunks), (total, chunks, t, row, chunk)
