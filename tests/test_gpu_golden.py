"""The CUDA path against the committed fixtures (tests/golden/*.npz)."""
import os

import numpy as np
import pytest
import torch

from tests._util import REL, rel_err, to_dev

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["deepfm_cfg1_small.npz", "deepfm_k16_skew.npz"])
@pytest.mark.parametrize("sharded", [False, True])
def test_deepfm_step_matches_fixture(pkg, cuda, name, sharded):
    g = dict(np.load(os.path.join(GOLD, name)))
    rows = [int(r) for r in g["rows"]]
    F, K = len(rows), g["table"].shape[1]
    cls = pkg.ShardedEmbeddingFM if sharded else pkg.EmbeddingFM
    layer = cls(F, K, rows, optimizer="adagrad", lr=float(g["lr"])).train()
    layer.load_tables(g["table"], g["w1"])
    first, fm, emb = layer(to_dev(g["idx"]), to_dev(g["val"]))
    B = g["idx"].shape[0]
    assert np.array_equal(emb.detach().cpu().numpy().reshape(B, F, K), g["e"].astype(np.float32)), \
        "gathered rows must be bit-exact"
    logits = (first + fm).detach()
    floor = 0.5 * (g["e"] ** 2).sum((1, 2))[:, None] + np.abs(g["first"])
    assert rel_err(logits.cpu().numpy(), g["logits"], floor) <= REL
    gy = (torch.sigmoid(logits) - to_dev(g["labels"]).unsqueeze(1))
    torch.autograd.backward((first, fm, emb), (gy, gy, to_dev(g["u"].reshape(B, -1))))
    torch.cuda.synchronize()
    t = g["touched"]
    assert rel_err(layer.table.cpu().numpy(), g["table_after"], np.abs(g["table"]).max()) <= REL
    assert rel_err(layer.w1.cpu().numpy(), g["w1_after"], np.abs(g["w1"]).max() + 1e-3) <= REL
    assert rel_err(layer.accum.cpu().numpy()[t], g["accum_after"][t], 0.1 + np.abs(g["accum_after"][t])) <= 1e-4
    untouched = np.ones(g["table"].shape[0], bool)
    untouched[t] = False
    assert np.array_equal(layer.table.cpu().numpy()[untouched], g["table"][untouched])


@pytest.mark.parametrize("name", ["dcn_d312_l3.npz", "dcn_d51_l2.npz"])
def test_cross_matches_fixture(pkg, cuda, name):
    g = dict(np.load(os.path.join(GOLD, name)))
    L, d = g["cross_w"].shape
    net = pkg.CrossNetwork(d, L).train()
    with torch.no_grad():
        net.cross_w.copy_(to_dev(g["cross_w"]))
        net.cross_b.copy_(to_dev(g["cross_b"]))
    x0 = to_dev(g["x0"]).requires_grad_(True)
    xL = net(x0)
    xL.backward(to_dev(g["dy"]))
    torch.cuda.synchronize()
    assert rel_err(xL.detach().cpu().numpy(), g["xL"], np.abs(g["xL"]).max()) <= REL
    assert rel_err(x0.grad.cpu().numpy(), g["dx0"], np.abs(g["dx0"]).max()) <= REL
    assert rel_err(net.cross_w.grad.cpu().numpy(), g["dw"], np.abs(g["dw"]).max()) <= REL
    assert rel_err(net.cross_b.grad.cpu().numpy(), g["db"], np.abs(g["db"]).max()) <= REL
