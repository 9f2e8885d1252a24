"""Synthetic molecular graphs (SURVEY.md §8d "synthetic molecule generator").

The reference builds per-molecule graphs from RDKit molecules in
``notorch/transforms/graph.py:32-43``: bond ``k`` becomes the two directed edges ``2k: u->v`` and
``2k+1: v->u`` and the local reverse map is ``[1, 0, 3, 2, ...]``. RDKit is not available on the
build or GPU boxes, so every test and benchmark uses this generator instead; it honours exactly
that edge-order contract and the size statistics of the reference's own fixtures
(``tests/data/lipo.csv``: bonds ~= 1.09 x heavy atoms, so directed edges ~= 2.18 x atoms).

Everything here is host-side numpy; nothing touches the GPU.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

__all__ = ["MolSpec", "CONFIGS", "PackedMolecules", "make_molecules", "config_seed"]


@dataclass(frozen=True)
class MolSpec:
    """Size distribution of one BASELINE.json config (atoms per molecule)."""

    mu: float
    sigma: float
    n_min: int
    n_max: int
    uniform: bool = False
    bond_ratio: float = 1.09
    max_degree: int = 4


# One entry per BASELINE.json config.
CONFIGS: dict[int, MolSpec] = {
    1: MolSpec(25.0, 3.0, 15, 35),
    2: MolSpec(23.0, 4.5, 6, 38),
    3: MolSpec(23.0, 4.5, 6, 38),
    4: MolSpec(23.0, 4.5, 6, 38),
    5: MolSpec(200.0, 0.0, 100, 300, uniform=True),
}


def config_seed(config: int, rank: int = 0) -> int:
    """Seed convention of SURVEY.md §8d: ``1234 + 1000 * config + rank``."""
    return 1234 + 1000 * config + rank


@dataclass
class PackedMolecules:
    """A batch of molecules in packed (pre-collation) form.

    ``edge_index`` / ``rev_index`` hold *molecule-local* indices, concatenated over molecules in
    batch order; ``num_atoms[i]`` / ``num_edges[i]`` give molecule ``i``'s slice lengths. This is
    the on-the-wire input of the device collation kernel (``nt_collate``).
    """

    num_atoms: np.ndarray  # [B] int32
    num_edges: np.ndarray  # [B] int32 (directed edges = 2 x bonds)
    edge_index: np.ndarray  # [2, E] int32, molecule-local atom ids
    rev_index: np.ndarray  # [E] int32, molecule-local edge ids

    def __len__(self) -> int:
        return len(self.num_atoms)

    @property
    def total_atoms(self) -> int:
        return int(self.num_atoms.sum())

    @property
    def total_edges(self) -> int:
        return int(self.num_edges.sum())

    def molecule(self, i: int) -> tuple[int, np.ndarray, np.ndarray]:
        """``(n_atoms, local edge_index [2, e], local rev_index [e])`` of molecule ``i``."""
        off = int(self.num_edges[:i].sum())
        e = int(self.num_edges[i])
        return int(self.num_atoms[i]), self.edge_index[:, off : off + e], self.rev_index[off : off + e]

    def split(self) -> list[tuple[int, np.ndarray, np.ndarray]]:
        offs = np.concatenate([[0], np.cumsum(self.num_edges)])
        return [
            (
                int(self.num_atoms[i]),
                self.edge_index[:, offs[i] : offs[i + 1]],
                self.rev_index[offs[i] : offs[i + 1]],
            )
            for i in range(len(self))
        ]

    def shard(self, rank: int, world_size: int) -> "PackedMolecules":
        """Contiguous molecule shard for data parallelism (shard *before*
        collation; a molecule's result depends on its local batch)."""
        B = len(self)
        lo = (B * rank) // world_size
        hi = (B * (rank + 1)) // world_size
        offs = np.concatenate([[0], np.cumsum(self.num_edges)])
        return PackedMolecules(
            self.num_atoms[lo:hi].copy(),
            self.num_edges[lo:hi].copy(),
            self.edge_index[:, offs[lo] : offs[hi]].copy(),
            self.rev_index[offs[lo] : offs[hi]].copy(),
        )


def _one_molecule(rng: np.random.Generator, n: int, spec: MolSpec) -> np.ndarray:
    """Bond list ``[nb, 2]`` of one synthetic molecule: a random spanning tree in which atom ``i``
    attaches to a uniformly random earlier atom of degree < ``max_degree``, plus ring closures
    between non-adjacent atoms of degree < ``max_degree`` until bonds ~= ``bond_ratio * n``."""
    deg = np.zeros(n, dtype=np.int64)
    bonds: list[tuple[int, int]] = []
    adj: set[tuple[int, int]] = set()
    for i in range(1, n):
        cand = np.flatnonzero(deg[:i] < spec.max_degree)
        p = int(cand[rng.integers(len(cand))]) if len(cand) else int(rng.integers(i))
        bonds.append((p, i))
        adj.add((p, i))
        deg[p] += 1
        deg[i] += 1
    target = int(round(spec.bond_ratio * n))
    tries = 0
    while len(bonds) < target and tries < 16 * n:
        tries += 1
        u, v = (int(x) for x in rng.integers(n, size=2))
        if u == v:
            continue
        a, b = (u, v) if u < v else (v, u)
        if (a, b) in adj or deg[u] >= spec.max_degree or deg[v] >= spec.max_degree:
            continue
        bonds.append((u, v))
        adj.add((a, b))
        deg[u] += 1
        deg[v] += 1
    return np.asarray(bonds, dtype=np.int32).reshape(-1, 2)


def make_molecules(
    batch: int, spec: MolSpec | int = 2, seed: int | None = None, bondless_every: int = 0
) -> PackedMolecules:
    """Generate ``batch`` synthetic molecules.

    ``bondless_every > 0`` replaces every ``bondless_every``-th molecule by a single bond-less
    atom (the edge case of SURVEY.md §3.4 / §4 item 5).
    """
    if isinstance(spec, int):
        if seed is None:
            seed = config_seed(spec)
        spec = CONFIGS[spec]
    rng = np.random.default_rng(1234 if seed is None else seed)
    if spec.uniform:
        sizes = rng.integers(spec.n_min, spec.n_max + 1, size=batch)
    else:
        sizes = np.clip(np.rint(rng.normal(spec.mu, spec.sigma, size=batch)), spec.n_min, spec.n_max)
    sizes = sizes.astype(np.int64)

    num_atoms = np.empty(batch, dtype=np.int32)
    num_edges = np.empty(batch, dtype=np.int32)
    eis: list[np.ndarray] = []
    revs: list[np.ndarray] = []
    for i in range(batch):
        n = int(sizes[i])
        if bondless_every and (i % bondless_every) == bondless_every - 1:
            n, bonds = 1, np.zeros((0, 2), dtype=np.int32)
        else:
            bonds = _one_molecule(rng, n, spec)
        nb = len(bonds)
        # bond-major directed edges: (u->v, v->u) per bond  [transforms/graph.py:36-40]
        ei = np.empty((2, 2 * nb), dtype=np.int32)
        ei[0, 0::2], ei[1, 0::2] = bonds[:, 0], bonds[:, 1]
        ei[0, 1::2], ei[1, 1::2] = bonds[:, 1], bonds[:, 0]
        # local reverse map [1, 0, 3, 2, ...]  [transforms/graph.py:41]
        rev = (np.arange(2 * nb, dtype=np.int32).reshape(-1, 2)[:, ::-1]).ravel()
        num_atoms[i], num_edges[i] = n, 2 * nb
        eis.append(ei)
        revs.append(rev)
    edge_index = np.concatenate(eis, axis=1) if eis else np.zeros((2, 0), np.int32)
    rev_index = np.concatenate(revs) if revs else np.zeros((0,), np.int32)
    return PackedMolecules(num_atoms, num_edges, np.ascontiguousarray(edge_index), rev_index)
