"""pytest configuration: the ``gpu`` marker, golden-fixture loading, and path set-up.

``-m "not gpu"`` runs on a CPU-only box (oracle vs golden vectors, host logic, C-ABI symbol
table, gloo world-size-2); ``-m gpu`` runs the parity tests proper on a B200 through the C-ABI.
"""
from __future__ import annotations

import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the live reference tree at /root/reference")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """One ``tests/golden/*.npz`` fixture (made by ``oracle/make_golden.py`` from the reference)."""

    def __init__(self, path: str):
        self.path = path
        self.z = np.load(path)
        self.meta = json.loads(bytes(self.z["meta"]).decode())

    def __getitem__(self, k: str) -> np.ndarray:
        return self.z[k]

    def __contains__(self, k: str) -> bool:
        return k in self.z.files

    @property
    def name(self) -> str:
        return self.meta["name"]

    def params(self) -> dict[str, np.ndarray]:
        return {k[len("param/"):]: self.z[k] for k in self.z.files if k.startswith("param/")}

    def grads(self, tag: str = "f32") -> dict[str, np.ndarray]:
        pre = f"{tag}/grad/"
        return {k[len(pre):]: self.z[k] for k in self.z.files if k.startswith(pre)}

    def mols(self):
        offs = np.concatenate([[0], np.cumsum(self.z["num_edges"])])
        return [
            (int(self.z["num_atoms"][i]), self.z["local_edge_index"][:, offs[i]:offs[i + 1]],
             self.z["local_rev_index"][offs[i]:offs[i + 1]])
            for i in range(len(self.z["num_atoms"]))
        ]


def golden_names() -> list[str]:
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name: str) -> Golden:
    return Golden(os.path.join(GOLDEN_DIR, name + ".npz"))


@pytest.fixture(params=golden_names())
def golden(request) -> Golden:
    return load_golden(request.param)
