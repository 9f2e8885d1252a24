"""``Graph`` / ``BatchedGraph`` containers — drop-in for ``notorch/data/models/graph.py`` (fields,
constructor arguments, ``to``, ``update``, ``from_graphs``, ``len``), minus the dense V x V helpers
(``A``, ``P``, ``dense2sparse``, ``random_walk``: unused by the hot path).

Collation comes in two forms with identical results:

* ``BatchedGraph.from_graphs(graphs)`` — host tensors in, same device out, like the reference
  (graph.py:186-223) but vectorised instead of a Python loop over atoms/edges;
* ``BatchedGraph.from_packed(packed, node_feats, edge_feats, device)`` — ships the packed int32
  molecule arrays to the GPU once and runs the collation kernel there (``nt_collate``).

Both reproduce the reference's index tensors bit for bit, *including* ``rev_index`` being offset by
the cumulative atom count (graph.py:199-200); ``fixed_rev=True`` is the labelled deviation that
offsets it by the cumulative edge count instead.
"""
from __future__ import annotations

from dataclasses import InitVar, dataclass, field
from typing import Iterable, Self

import numpy as np
import torch
from torch import Tensor

from ...utils.utils import UpdateMixin


class PendingFeats:
    """A float feature matrix that has been *described* but not computed yet.

    ``GraphEmbedding`` stores one of these in ``node_feats`` / ``edge_feats`` so that the message-passing block can fuse the
    embedding-table look-ups into its edge initialisation (``nt_embed_edge_init``: neither ``[V, d]`` nor ``[E, d]`` is ever
    written). Anything else that READS ``G.node_feats`` / ``G.edge_feats`` gets a real tensor: the attribute access computes it
    through the unfused kernel (same bits, autograd-connected), caches it here and returns it — so the graph stays a drop-in for
    consumers that know nothing about this class. ``Graph.peek(name)`` returns the placeholder without computing."""

    __slots__ = ("_compute", "shape", "dtype", "device", "origin", "_value")

    def __init__(self, compute, shape: tuple[int, int], dtype: torch.dtype, device: torch.device, origin):
        self._compute, self.shape, self.dtype, self.device, self.origin = compute, tuple(shape), dtype, device, origin
        self._value: Tensor | None = None

    def __len__(self) -> int:
        return self.shape[0]

    @property
    def materialized(self) -> bool:
        return self._value is not None

    def materialize(self) -> Tensor:
        if self._value is None:
            self._value = self._compute()
        return self._value


def _feats_property(name: str) -> property:
    slot = "_" + name  # the value lives under another name than the property (self.__dict__[name] would re-enter the property under torch.compile)

    def getter(self):
        v = getattr(self, slot)
        if isinstance(v, PendingFeats):
            v = v.materialize()
            object.__setattr__(self, slot, v)
        return v

    def setter(self, value):
        object.__setattr__(self, slot, value)

    return property(getter, setter)


@dataclass(repr=False, eq=False)
class Graph(UpdateMixin):
    node_feats: Tensor  # [V, t_v] integer types before embedding, [V, d] float after
    edge_feats: Tensor  # [E, t_e] / [E, d]
    edge_index: Tensor  # [2, E] int64 COO, row 0 = source atom, row 1 = destination atom
    rev_index: Tensor  # [E] int64, index of each edge's reverse edge
    device_: InitVar[torch.device | str | int | None] = field(default=None, kw_only=True)

    def __post_init__(self, device_):
        self._device = device_
        self.to(device_)

    def peek(self, name: str):
        """``node_feats`` / ``edge_feats`` WITHOUT materialising a :class:`PendingFeats` placeholder."""
        return getattr(self, "_" + name)

    @property
    def num_nodes(self) -> int:
        return len(self.peek("node_feats"))

    @property
    def num_edges(self) -> int:
        return len(self.peek("edge_feats"))

    @property
    def device(self):
        return self._device

    _TENSOR_FIELDS = ("node_feats", "edge_feats", "edge_index", "rev_index")

    _CACHE_ATTRS = ("_nt_csr", "_nt_seg_csr", "_nt_mol_ptr", "_nt_mol_csr", "_nt_mol_edge_ptr", "_nt_mol_edge_csr", "_nt_mol_atom_count")

    def to(self, device, non_blocking: bool = False) -> Self:
        """Move the tensors (graph.py:45-53 / :229-239 of the reference) AND whatever the kernels cached on this object: the
        int32 CSR bundle, the molecule row pointers and the read-out CSR follow the graph to another CUDA device (small int32
        copies instead of a rebuild) and survive a no-op move untouched, so ``transfer_batch_to_device`` of a Lightning loop
        does not throw the per-batch preprocessing away. Moving to the CPU drops them (they only serve the CUDA kernels)."""
        self._device = device
        before = {name: getattr(self, name) for name in self._TENSOR_FIELDS}
        for name, t in before.items():
            setattr(self, name, t.to(device, non_blocking=non_blocking))
        if all(getattr(self, name) is t for name, t in before.items()):
            return self  # nothing moved: every cache key is still valid
        caches = {a: self.__dict__.pop(a) for a in self._CACHE_ATTRS if a in self.__dict__}
        if caches and self.edge_index.is_cuda:
            from ... import ops

            ops.carry_graph_caches(self, caches, before)
        return self

    def _field_lines(self) -> list[str]:
        return [
            f"node_feats: Tensor(shape={tuple(self.peek('node_feats').shape)})",
            f"edge_feats: Tensor(shape={tuple(self.peek('edge_feats').shape)})",
            f"device={self._device}",
        ]

    def __repr__(self) -> str:
        body = "\n".join("  " + line for line in self._field_lines())
        return f"{type(self).__name__}(\n{body}\n)"


# installed after @dataclass has read the field list (a property in the class body would be taken for a default value)
Graph.node_feats = _feats_property("node_feats")
Graph.edge_feats = _feats_property("edge_feats")


@dataclass(repr=False, eq=False, kw_only=True)
class BatchedGraph(Graph):
    batch_node_index: Tensor  # [V] int64, molecule id of each atom (non-decreasing from from_graphs)
    batch_edge_index: Tensor  # [E] int64, molecule id of each edge
    size: InitVar[int | None] = None

    _TENSOR_FIELDS = Graph._TENSOR_FIELDS + ("batch_node_index", "batch_edge_index")

    def __post_init__(self, device_, size):
        super().__post_init__(device_)
        # like the reference (graph.py:181-184): without `size` this costs a device sync
        self._size = int(self.batch_node_index.max()) + 1 if size is None else int(size)

    def __len__(self) -> int:
        return self._size

    def _field_lines(self) -> list[str]:
        return super()._field_lines() + [f"batch_size={len(self)}"]

    # ---- collation -------------------------------------------------------------------------
    @classmethod
    def from_graphs(cls, Gs: Iterable[Graph], fixed_rev: bool = False) -> "BatchedGraph":
        Gs = list(Gs)
        if not Gs:
            raise ValueError("from_graphs needs at least one graph")
        n_atoms = torch.tensor([len(G.node_feats) for G in Gs], dtype=torch.long)
        n_edges = torch.tensor([len(G.edge_feats) for G in Gs], dtype=torch.long)
        atom_off = torch.cumsum(n_atoms, 0) - n_atoms
        edge_off = torch.cumsum(n_edges, 0) - n_edges
        mol_ids = torch.arange(len(Gs), dtype=torch.long)
        batch_node_index = torch.repeat_interleave(mol_ids, n_atoms)
        batch_edge_index = torch.repeat_interleave(mol_ids, n_edges)
        dev = Gs[-1].device
        idx_dev = Gs[-1].edge_index.device

        node_feats = torch.cat([G.node_feats for G in Gs], dim=0)
        edge_feats = torch.cat([G.edge_feats for G in Gs], dim=0)
        # bond-less molecules carry an empty (possibly 1-D) edge_index; skip them like torch.cat does
        local_ei = [G.edge_index.reshape(2, -1).long() for G in Gs if G.edge_index.numel() > 0]
        edge_index = torch.cat(local_ei, dim=1) if local_ei else torch.zeros((2, 0), dtype=torch.long, device=idx_dev)
        rev_index = torch.cat([G.rev_index.long() for G in Gs], dim=0)
        per_edge_atom_off = atom_off[batch_edge_index].to(edge_index.device)
        edge_index = edge_index + per_edge_atom_off  # graph.py:199
        # graph.py:200 adds the ATOM offset to rev_index as well (reproduced); fixed_rev uses the edge offset
        rev_index = rev_index + (edge_off[batch_edge_index].to(rev_index.device) if fixed_rev else per_edge_atom_off)
        return cls(node_feats, edge_feats, edge_index, rev_index, device_=dev,
                   batch_node_index=batch_node_index, batch_edge_index=batch_edge_index, size=len(Gs))

    @classmethod
    def from_packed(cls, packed, node_feats: Tensor, edge_feats: Tensor, device="cuda", fixed_rev: bool = False,
                    non_blocking: bool = True) -> "BatchedGraph":
        """Device-side collation of a :class:`notorch_b200.synth.PackedMolecules`-like object
        (``num_atoms``, ``num_edges``, local ``edge_index`` [2,E], local ``rev_index`` [E], int32)."""
        from ... import ops

        device = torch.device(device)

        def dev(a):
            t = torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a
            if t.device.type == "cpu" and device.type == "cuda" and non_blocking:
                t = t.pin_memory()
            return t.to(device, non_blocking=non_blocking)

        na, ne = dev(packed.num_atoms), dev(packed.num_edges)
        lei, lrev = dev(packed.edge_index), dev(packed.rev_index)
        V, E = node_feats.shape[0], edge_feats.shape[0]
        if lei.shape[1] != E:
            raise ValueError(f"edge_feats has {E} rows but the packed molecules have {lei.shape[1]} edges")
        out = ops.collate_packed(na, ne, lei, lrev, V, E, fixed_rev)
        G = cls(node_feats.to(device, non_blocking=non_blocking), edge_feats.to(device, non_blocking=non_blocking),
                out["edge_index"], out["rev_index"], device_=device,
                batch_node_index=out["batch_node_index"], batch_edge_index=out["batch_edge_index"], size=len(packed.num_atoms))
        # molecules are contiguous atom ranges: the read-out CSR needs no permutation
        G._nt_mol_ptr = out["mol_atom_ptr"]
        # ... and contiguous edge ranges whose destination atoms all lie in the same molecule (this collation made them so): lets a Sum /
        # Mean / Norm read-out of a sum-reduced block run straight over the edge states (agg.py here)
        G._nt_mol_edge_ptr = out["mol_edge_ptr"]
        return G
