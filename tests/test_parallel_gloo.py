"""world_size-2 gloo tests (CPU) of the host-side data-parallel logic: molecule sharding before
collation and the single flat-gradient all-reduce."""
from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from notorch_b200.parallel import FlatGradients, shard_range
        from notorch_b200.synth import make_molecules
        from oracle import dmpnn_oracle as O

        torch.manual_seed(0)  # same init on every rank, like DDP's broadcast
        d = 12
        model = O.CpuPort(hidden_dim=d, depth=2)
        flat = FlatGradients(model.parameters())
        mols = make_molecules(8, 1, seed=3)  # the GLOBAL batch
        lo, hi = shard_range(len(mols), rank, world)
        mine = mols.shard(rank, world)  # shard BEFORE collation
        assert len(mine) == hi - lo
        c = O.collate(mine.split())
        gen = torch.Generator().manual_seed(100 + rank)
        xv, xe = torch.randn(mine.total_atoms, d, generator=gen), torch.randn(mine.total_edges, d, generator=gen)
        flat.zero()
        H, _, _ = model(xv, xe, torch.from_numpy(c["edge_index"]), torch.from_numpy(c["rev_index"]),
                        torch.from_numpy(c["batch_node_index"]), len(mine))
        H.square().mean().backward()
        local = flat.flat.clone()
        flat.all_reduce_mean()
        reduced = flat.flat.clone()

        # bucketed + overlapped form (what bench.py uses for N > 1): one bucket per parameter group, each all-reduce issued from the
        # post-accumulate-grad hooks as soon as the bucket is complete, joined by finish(); same numbers as the single collective
        torch.manual_seed(0)
        model2 = O.CpuPort(hidden_dim=d, depth=2)
        ps = list(model2.parameters())
        flat2 = FlatGradients([ps[:2], ps[2:]], overlap=True)
        assert len(flat2.buckets) == 2 and flat2.slices[0][1] == flat2.slices[1][0]
        for _ in range(2):  # two steps: the per-step hook state resets in finish()
            flat2.zero()
            H2, _, _ = model2(xv, xe, torch.from_numpy(c["edge_index"]), torch.from_numpy(c["rev_index"]),
                              torch.from_numpy(c["batch_node_index"]), len(mine))
            H2.square().mean().backward()
            flat2.finish()
        overlapped = flat2.flat.clone()

        # the default zero_grad() (set_to_none=True) detaches the views: the exchange must refuse, rebind() must repair
        for p in ps:
            p.grad = None
        H2, _, _ = model2(xv, xe, torch.from_numpy(c["edge_index"]), torch.from_numpy(c["rev_index"]), torch.from_numpy(c["batch_node_index"]), len(mine))
        try:
            H2.square().mean().backward()  # the bucket hook refuses before anything is exchanged
            refused = False
        except RuntimeError as exc:
            refused = "no longer aliases" in str(exc)
        flat2.reset()
        for p in ps:
            p.grad = None
        H2, _, _ = model2(xv, xe, torch.from_numpy(c["edge_index"]), torch.from_numpy(c["rev_index"]), torch.from_numpy(c["batch_node_index"]), len(mine))
        for h in flat2._hooks:
            h.remove()
        H2.square().mean().backward()
        moved = flat2.rebind()
        flat2.overlap = False
        flat2.all_reduce_mean()
        torch.save({"local": local, "reduced": reduced, "overlapped": overlapped, "rebound": flat2.flat.clone(), "refused": refused, "moved": moved,
                    "views_ok": all(p.grad.data_ptr() >= flat.flat.data_ptr() for p in flat.params)},
                   os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_flat_gradient_allreduce_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(tmp_path / f"rank{i}.pt") for i in range(world)]
    want = (r[0]["local"] + r[1]["local"]) / world
    for i in range(world):
        assert r[i]["views_ok"]
        assert torch.allclose(r[i]["reduced"], want, rtol=0, atol=1e-7)
        assert torch.allclose(r[i]["overlapped"], want, rtol=0, atol=1e-7)
        assert r[i]["refused"] and r[i]["moved"] == 4
        assert torch.allclose(r[i]["rebound"], want, rtol=0, atol=1e-7)
    assert not torch.equal(r[0]["local"], r[1]["local"])  # ranks really saw different molecules


def test_shard_range_partitions():
    from notorch_b200.parallel import shard_list, shard_range

    for n in (0, 1, 7, 64, 4097):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    assert list(shard_list(list(range(10)), 1, 3)) == [3, 4, 5]
    with pytest.raises(ValueError):
        shard_range(4, 4, 4)


def test_packed_shard_matches_collation_of_subset():
    from notorch_b200.synth import make_molecules
    from oracle import dmpnn_oracle as O

    mols = make_molecules(9, 1, seed=1)
    parts = [mols.shard(r, 2) for r in range(2)]
    allm = mols.split()
    lo = 0
    for part in parts:
        want = O.collate(allm[lo:lo + len(part)])
        got = O.collate(part.split())
        assert all(np.array_equal(want[k], got[k]) for k in ("edge_index", "rev_index", "batch_node_index"))
        lo += len(part)
