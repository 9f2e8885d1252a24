"""Row N3: the prediction head ``MLP`` (notorch/nn/mlp.py:9-68). CPU: module structure and state-dict layout; GPU: the Linear
kernels against the reference's own outputs (tests/golden_readouts/mlp_head.npz, made by oracle/make_golden.py) and an fp64 restatement."""
from __future__ import annotations

import json
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from helpers import REL_F32, assert_close

GOLD = os.path.join(os.path.dirname(__file__), "golden_readouts", "mlp_head.npz")


def test_mlp_factory_structure_matches_reference_layout():
    from notorch_b200.nn import MLP, Linear

    z = np.load(GOLD)
    meta = json.loads(bytes(z["meta"]).decode())
    mlp = MLP(meta["input_dim"], tuple(meta["output_size"]), meta["hidden_dim"], meta["num_layers"], meta["dropout"], nn.SiLU)
    ref_keys = [k[len("param/"):] for k in z.files if k.startswith("param/")]
    assert list(mlp.state_dict()) == ref_keys  # "0.weight", "0.bias", "3.weight", ... : the reference's Sequential indices
    assert [tuple(v.shape) for v in mlp.state_dict().values()] == [z["param/" + k].shape for k in ref_keys]
    kinds = [type(m).__name__ for m in mlp]
    assert kinds == ["Linear", "SiLU", "Dropout", "Linear", "SiLU", "Dropout", "Linear", "Unflatten"]
    assert mlp[1] is mlp[4] and mlp[2] is mlp[5]  # one shared activation / dropout instance, as in the reference
    assert all(isinstance(m, Linear) for m in mlp if isinstance(m, nn.Linear))
    single = MLP(8, 4, num_layers=0)
    assert len(single) == 1 and single[0].in_features == 8 and single[0].out_features == 4
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        single(torch.zeros(2, 8))


@pytest.mark.gpu
def test_mlp_matches_reference_golden():
    from notorch_b200.nn import MLP

    z = np.load(GOLD)
    meta = json.loads(bytes(z["meta"]).decode())
    mlp = MLP(meta["input_dim"], tuple(meta["output_size"]), meta["hidden_dim"], meta["num_layers"], meta["dropout"], nn.SiLU)
    mlp.load_state_dict({k[len("param/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}, strict=True)
    mlp = mlp.cuda()
    x = torch.from_numpy(z["x"]).cuda().requires_grad_(True)
    y = mlp(x)
    assert tuple(y.shape) == z["f32/y"].shape
    (y * torch.from_numpy(z["gY"]).cuda()).sum().backward()
    for tag in ("f32", "f64"):
        assert_close(y.detach().cpu().double(), torch.from_numpy(z[f"{tag}/y"]).double(), f"y vs {tag}")
        assert_close(x.grad.cpu().double(), torch.from_numpy(z[f"{tag}/g_x"]).double(), f"g_x vs {tag}")
        for k, p in mlp.named_parameters():
            assert_close(p.grad.cpu().double(), torch.from_numpy(z[f"{tag}/grad/{k}"]).double(), f"grad {k} vs {tag}")


@pytest.mark.gpu
@pytest.mark.parametrize("rows,k,n,bias", [(4096, 300, 300, True), (4096, 300, 1, True), (37, 5, 129, False), (1, 64, 64, True), (0, 16, 8, True),
                                           (1000, 1024, 7, True)])
def test_linear_kernels_vs_fp64(rows, k, n, bias):
    from notorch_b200.nn import Linear

    torch.manual_seed(rows + k + n)
    lin = Linear(k, n, bias=bias).cuda()
    x = torch.randn(rows, k, device="cuda", requires_grad=True)
    g = torch.randn(rows, n, device="cuda")
    y = lin(x)
    y.backward(g)
    W, b = lin.weight.detach().double(), (lin.bias.detach().double() if bias else None)
    xd = x.detach().double()
    assert tuple(y.shape) == (rows, n)
    if rows:
        assert_close(y.detach().double(), torch.nn.functional.linear(xd, W, b), "y", REL_F32)
        assert_close(x.grad.double(), g.double() @ W, "g_x", REL_F32)
        assert_close(lin.weight.grad.double(), g.double().t() @ xd, "g_W", REL_F32)
        if bias:
            assert_close(lin.bias.grad.double(), g.double().sum(0), "g_b", REL_F32)
    else:
        assert float(lin.weight.grad.abs().sum()) == 0.0


@pytest.mark.gpu
def test_linear_leading_dims_and_determinism():
    from notorch_b200.nn import Linear

    torch.manual_seed(0)
    lin = Linear(48, 20).cuda()
    x = torch.randn(3, 5, 48, device="cuda")
    y1, y2 = lin(x), lin(x)
    assert y1.shape == (3, 5, 20) and torch.equal(y1, y2)
    assert_close(y1.double(), torch.nn.functional.linear(x.double(), lin.weight.double(), lin.bias.double()), "y", REL_F32)
